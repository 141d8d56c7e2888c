#!/bin/bash
# r02p: Taxi rows padded to 32 bytes and read with ONE 256-bit load (round 1 preferred 24-byte rows read with three 64-bit loads).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
B="--steps 6 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
for v in main taxi_apad8 main taxi_apad8; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  eval RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c4 $B >> $O/r02p_ab_c4_$v.json 2>> $O/r02p_err.log
  tail -1 $O/r02p_ab_c4_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 $v', d['value'], d['ms_per_step'])"
done
eval RLB_LIB=$PWD/rl-rust_b200/ab/librlb_taxi_apad8.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k taxi > $O/r02p_pytest.log 2>&1; tail -2 $O/r02p_pytest.log
tail -3 $O/r02p_err.log
