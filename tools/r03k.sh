#!/bin/bash
# r03k: lazy trace store with 64-thread CTAs (7 instead of 3 CTAs/SM at 144 registers: +17 % resident threads, finer waves).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
RLB_LIB=$PWD/rl-rust_b200/ab/librlb_lz64.so timeout 300 python -m pytest tests/test_gpu_abi2.py -m gpu -q -x -k lazy > $O/r03k_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $O/r03k_pytest.log | cut -c1-200
for v in main lz64; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  echo "== $v"
  RLB_LIB=$PWD/$lib timeout 300 python tools/lazy_phase.py 102400 500 0 2 > $O/r03k_lazy_phase_$v.txt 2>> $O/r03k_err.log; grep -v '^{' $O/r03k_lazy_phase_$v.txt | cut -c1-200 | awk 'NR%2==1 || /taxi/'
done
tail -3 $O/r03k_err.log
