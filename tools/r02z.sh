#!/bin/bash
# r02z: lazy store: Q rows requested ahead of the slot lookup and handed on in registers (A/B against the previous form).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_abi2.py tests/test_gpu_parity.py tests/test_gpu_random.py -m gpu -q -x -k "lazy or traces or random" > $O/r02z_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/r02z_pytest.log | cut -c1-200
for v in main lz_nofused main lz_nofused; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  echo "== $v"
  RLB_LIB=$PWD/$lib timeout 600 python tools/lazy_phase.py 102400 1000 0 2 >> $O/r02z_lazy_phase_$v.txt 2>> $O/r02z_err.log; grep -v '^{' $O/r02z_lazy_phase_$v.txt | tail -22 | cut -c1-200 | awk 'NR%3==1 || /taxi/'
done
tail -3 $O/r02z_err.log
