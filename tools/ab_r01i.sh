#!/bin/bash
# A/B (r01i): hybrid sweep ring depth / trip size, lazy RNG for the hybrid store, packed 24-byte Taxi rows.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r01i_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r01i_pytest.log
tail -3 $O/r01i_pytest.log
RLB_LIB=$PWD/rl-rust_b200/ab/librlb_taxi_apad6.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random_configs.py tests/test_gpu_snapshot.py -m gpu -x -q > $O/r01i_pytest_apad6.log 2>&1; echo "pytest apad6 exit $?" | tee -a $O/r01i_pytest_apad6.log
tail -3 $O/r01i_pytest_apad6.log
B="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
for v in main fl_p8 fl_t4 fl_t6 fl_q4 fl_p6lazy main fl_p8 fl_t4 fl_t6 fl_q4 fl_p6lazy; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c2 $B >> $O/r01i_ab_c2_$v.json 2>> $O/r01i_ab_err.log
  tail -1 $O/r01i_ab_c2_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2 $v', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
for v in main taxi_apad6 main taxi_apad6; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c4 $B >> $O/r01i_ab_c4_$v.json 2>> $O/r01i_ab_err.log
  tail -1 $O/r01i_ab_c4_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 $v', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
ls -la $O | grep r01i
