#!/bin/bash
# A/B (r01j): new defaults (4 sets x 4 rows, 24-byte Taxi rows) parity; hoisting the sweep's first trips to the top of the step.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r01j_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r01j_pytest.log
tail -3 $O/r01j_pytest.log
RLB_LIB=$PWD/rl-rust_b200/ab/librlb_fl_q4h.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random_configs.py -m gpu -x -q > $O/r01j_pytest_q4h.log 2>&1; echo "pytest q4h exit $?" | tee -a $O/r01j_pytest_q4h.log
tail -3 $O/r01j_pytest_q4h.log
B="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
for v in main fl_q4h fl_t4h fl_p6h fl_q5 main fl_q4h fl_t4h fl_p6h fl_q5; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c2 $B >> $O/r01j_ab_c2_$v.json 2>> $O/r01j_ab_err.log
  tail -1 $O/r01j_ab_c2_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2 $v', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
for wl in c4 c1 c3; do
  timeout 300 python bench.py --workload $wl $B >> $O/r01j_main_$wl.json 2>> $O/r01j_ab_err.log
  tail -1 $O/r01j_main_$wl.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl main', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
ls -la $O | grep r01j
