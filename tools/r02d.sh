#!/bin/bash
# r02d: C3 after the count store from registers: main (8 CTAs/SM) vs 10 / 12 CTAs per SM; UCB parity; C5 sweep driven by ONE
# host thread through the asynchronous train call (8 streams) at the quick and the BASELINE size; C1 for reference.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random_configs.py tests/test_gpu_abi2.py -m gpu -q -x > $O/r02d_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r02d_pytest.log
tail -3 $O/r02d_pytest.log
B="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
for v in main ucb_mb10 ucb_mb12 main ucb_mb10 ucb_mb12; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  eval RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c3 $B >> $O/r02d_ab_c3_$v.json 2>> $O/r02d_err.log
  tail -1 $O/r02d_ab_c3_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3 $v', d['value'], d['ms_per_step'])"
done
for k in 1 8 16; do
  eval timeout 400 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --cell-streams $k --sub "''" > $O/r02d_c5_streams$k.json 2>> $O/r02d_err.log
  tail -1 $O/r02d_c5_streams$k.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5 8192/cell streams $k', d['value'], d['ms_per_step'])"
done
eval timeout 600 python bench.py --workload c5 --agents-per-gpu 102400 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --cell-streams 8 --sub "''" > $O/r02d_c5_full_streams8.json 2>> $O/r02d_err.log
tail -1 $O/r02d_c5_full_streams8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5 102400/cell streams 8', d['value'], d['ms_per_step'])"
tail -5 $O/r02d_err.log
