#!/bin/bash
# r02t: lazy store: how many sweeps between two full updates of the rows (RLB_LZ_CAP), with and without the warp-voted flush.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_abi2.py -m gpu -q -x -k lazy > $O/r02t_pytest_lazy.log 2>&1; echo "lazy exit $?"; tail -3 $O/r02t_pytest_lazy.log | cut -c1-200
timeout 600 python tools/lazy_phase.py 102400 1000 4,8,16,32,0 2 > $O/r02t_lazy_phase_joint.txt 2> $O/r02t_err.log; grep -v '^{' $O/r02t_lazy_phase_joint.txt | cut -c1-220
RLB_LIB=$PWD/rl-rust_b200/ab/librlb_lz_nojoint.so timeout 600 python tools/lazy_phase.py 102400 1000 8,16 1 > $O/r02t_lazy_phase_nojoint.txt 2>> $O/r02t_err.log; grep -v '^{' $O/r02t_lazy_phase_nojoint.txt | cut -c1-220
tail -3 $O/r02t_err.log
