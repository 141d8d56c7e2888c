#!/bin/bash
# r02k: per-cell cost of the C5 sweep at BASELINE's size; GPU tests of the custom-map mirrors.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python tools/c5_cells.py 102400 > $O/r02k_c5_cells.txt 2> $O/r02k_err.log; head -45 $O/r02k_c5_cells.txt | cut -c1-140
timeout 600 python -m pytest tests/test_gpu_mirror.py tests/test_cpp_mirror.py tests/test_gpu_abi2.py -m gpu -q -x > $O/r02k_pytest.log 2>&1; tail -2 $O/r02k_pytest.log
tail -3 $O/r02k_err.log
