#!/bin/bash
# r02w: lazy store, the step's row materialised cooperatively too: parity, A/B against the flush-only build, ncu regions.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_abi2.py tests/test_gpu_parity.py -m gpu -q -x -k "lazy or traces" > $O/r02w_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/r02w_pytest.log | cut -c1-200
for v in main lz_norow; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  echo "== $v"
  RLB_LIB=$PWD/$lib timeout 600 python tools/lazy_phase.py 102400 1000 0 2 > $O/r02w_lazy_phase_$v.txt 2>> $O/r02w_err.log; grep -v '^{' $O/r02w_lazy_phase_$v.txt | cut -c1-200
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_run -s 1 -c 1 -f -o $O/r02w_taxi_lazy_k_run python tools/lazy_phase.py 32768 100 0 1 > $O/r02w_ncu.log 2>&1
echo "ncu exit $?"
tail -3 $O/r02w_err.log
