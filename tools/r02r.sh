#!/bin/bash
# r02r: lazy trace sweeps (store 4): parity, then the C5 cells and the sweep.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_abi2.py -m gpu -q -x -k lazy > $O/r02r_pytest_lazy.log 2>&1; echo "lazy exit $?"; tail -15 $O/r02r_pytest_lazy.log | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -q > $O/r02r_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r02r_pytest.log
tail -6 $O/r02r_pytest.log | cut -c1-200
timeout 900 python tools/c5_cells.py 102400 > $O/r02r_c5_cells.txt 2> $O/r02r_err.log; head -24 $O/r02r_c5_cells.txt | cut -c1-140; tail -1 $O/r02r_c5_cells.txt | cut -c1-200
eval timeout 600 python bench.py --workload c5 --agents-per-gpu 102400 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --cell-streams 8 --sub "''" > $O/r02r_c5_full_streams8.json 2>> $O/r02r_err.log
tail -1 $O/r02r_c5_full_streams8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5 102400/cell streams 8', d['value'], d['ms_per_step'])"
tail -3 $O/r02r_err.log
