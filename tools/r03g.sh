#!/bin/bash
# r03g: C4 (Taxi Q-learning) CTAs per SM re-measured on the final kernel: 6 / 7 / 8 (main) / 10.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
B="--steps 6 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
for v in main taxi_mb6 taxi_mb7 taxi_mb10 main taxi_mb6 taxi_mb7 taxi_mb10; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  eval RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c4 $B >> $O/r03g_ab_c4_$v.json 2>> $O/r03g_err.log
  tail -1 $O/r03g_ab_c4_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 $v', d['value'], d['ms_per_step'])"
done
tail -3 $O/r03g_err.log
