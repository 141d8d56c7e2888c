#!/bin/bash
# r03b: one-step Double agents carry the (s, a) cell of both tables in registers: parity, then C3 A/B same box.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random_configs.py tests/test_gpu_mirror.py -m gpu -q -x > $O/r03b_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/r03b_pytest.log | cut -c1-200
B="--steps 6 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
for v in main nocarry2 main nocarry2; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  eval RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c3 $B >> $O/r03b_ab_c3_$v.json 2>> $O/r03b_err.log
  tail -1 $O/r03b_ab_c3_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3 $v', d['value'], d['ms_per_step'])"
done
for v in main nocarry2; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  RLB_LIB=$PWD/$lib timeout 600 python tools/c5_cells.py 102400 > $O/r03b_c5_cells_$v.txt 2>> $O/r03b_err.log; grep "onestep.*double" $O/r03b_c5_cells_$v.txt | head -12 | cut -c1-140; tail -1 $O/r03b_c5_cells_$v.txt | cut -c1-160
done
tail -3 $O/r03b_err.log
