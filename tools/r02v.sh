#!/bin/bash
# r02v: lazy store with the warp-cooperative flush: parity, then eager vs lazy along a run (main; 4 CTAs/SM; the per-lane flush).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_abi2.py tests/test_gpu_parity.py -m gpu -q -x -k "lazy or traces" > $O/r02v_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/r02v_pytest.log | cut -c1-200
for v in main lz_mb4 lz_nocoop; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  echo "== $v"
  RLB_LIB=$PWD/$lib timeout 600 python tools/lazy_phase.py 102400 1000 0 2 > $O/r02v_lazy_phase_$v.txt 2>> $O/r02v_err.log; grep -v '^{' $O/r02v_lazy_phase_$v.txt | cut -c1-200
done
tail -3 $O/r02v_err.log
