#!/bin/bash
# A/B (r01n): Taxi start thresholds out of shared memory + L1 carve-out hint for the HBM-store kernels; C5 cells spread over
# host threads / CUDA streams.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r01n_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r01n_pytest.log
tail -3 $O/r01n_pytest.log
B="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
for v in taxi_old main taxi_nocarve taxi_thrs taxi_old main taxi_nocarve taxi_thrs; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c4 $B >> $O/r01n_ab_c4_$v.json 2>> $O/r01n_err.log
  tail -1 $O/r01n_ab_c4_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 $v', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
for wl in c1 c3; do
  timeout 300 python bench.py --workload $wl $B >> $O/r01n_main_$wl.json 2>> $O/r01n_err.log
  tail -1 $O/r01n_main_$wl.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl main', d['value'], d['ms_per_step'])"
done
for k in 1 8; do
  timeout 400 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --cell-streams $k > $O/r01n_c5_streams$k.json 2>> $O/r01n_err.log
  tail -1 $O/r01n_c5_streams$k.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5 8192/cell streams $k', d['value'], d['ms_per_step'])"
done
timeout 500 python bench.py --workload c5 --agents-per-gpu 102400 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --cell-streams 8 > $O/r01n_c5_full_streams8.json 2>> $O/r01n_err.log
tail -1 $O/r01n_c5_full_streams8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5 102400/cell streams 8', d['value'], d['ms_per_step'])"
tail -5 $O/r01n_err.log
