#!/bin/bash
# Build an A/B variant of librlb.so: tools/build_variant.sh NAME "-DFLAG=1 ..." tu1 [tu2 ...]
# Recompiles only the named translation units (e.g. rlb_inst_cliff_walking) with the extra flags and links them with the
# main build's other objects into rl-rust_b200/ab/librlb_NAME.so (run with RLB_LIB=<that file>).
set -e
cd "$(dirname "$0")/../rl-rust_b200/csrc"
NAME=$1; FLAGS=$2; shift 2
ARCH="-gencode arch=compute_100a,code=sm_100a"
NV="$ARCH -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-Wall -Xptxas -v"
D=../build_ab/$NAME; mkdir -p $D ../ab
OBJS=""
for o in ../build/*.o; do
  b=$(basename $o .o); use=$o
  for tu in "$@"; do [ "$tu" = "$b" ] && use=$D/$b.o; done
  OBJS="$OBJS $use"
done
for tu in "$@"; do nvcc $NV $FLAGS -c $tu.cu -o $D/$tu.o 2> $D/$tu.ptxas.log & done
wait
nvcc $ARCH -shared -o ../ab/librlb_$NAME.so $OBJS -cudart static -ldl
echo built ../ab/librlb_$NAME.so
