#!/bin/bash
# r03d: UCB one-step kernels at 6 (5) CTAs per SM, with and without the Double carry: C3 and the C5 cells.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
B="--steps 6 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
for v in main all_c6 all_n6 c3_c5 main all_c6 all_n6 c3_c5; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  eval RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c3 $B >> $O/r03d_ab_c3_$v.json 2>> $O/r03d_err.log
  tail -1 $O/r03d_ab_c3_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3 $v', d['value'], d['ms_per_step'])"
done
for v in main all_c6 all_n6; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  RLB_LIB=$PWD/$lib timeout 600 python tools/c5_cells.py 102400 > $O/r03d_c5_cells_$v.txt 2>> $O/r03d_err.log; echo "== $v"; grep "onestep.*ucb" $O/r03d_c5_cells_$v.txt | sort | cut -c1-140; tail -1 $O/r03d_c5_cells_$v.txt | cut -c1-160
done
tail -3 $O/r03d_err.log
