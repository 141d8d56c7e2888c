#!/bin/bash
# A/B (r01h): hybrid sweep with the split touch (copy vs ping-pong prefetch, 4 / 6 rows), Taxi occupancy 8 / 9 / 10 CTAs per SM.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r01h_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r01h_pytest.log
tail -3 $O/r01h_pytest.log
B="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
for v in fl_notouch main fl_c6 fl_p4 fl_p6 fl_notouch main fl_c6 fl_p4 fl_p6; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c2 $B >> $O/r01h_ab_c2_$v.json 2>> $O/r01h_ab_err.log
  tail -1 $O/r01h_ab_c2_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2 $v', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
for v in main taxi_mb9 taxi_mb10 main taxi_mb9 taxi_mb10; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c4 $B >> $O/r01h_ab_c4_$v.json 2>> $O/r01h_ab_err.log
  tail -1 $O/r01h_ab_c4_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 $v', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
ls -la $O | grep r01h
