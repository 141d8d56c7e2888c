#!/bin/bash
# r03a: full GPU suite, C5 cells and sweep on the lazy store with rows handed on in registers.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r03a_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r03a_pytest.log
tail -6 $O/r03a_pytest.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python tools/c5_cells.py 102400 > $O/r03a_c5_cells.txt 2> $O/r03a_err.log; head -14 $O/r03a_c5_cells.txt | cut -c1-140; tail -1 $O/r03a_c5_cells.txt | cut -c1-200
eval timeout 600 python bench.py --workload c5 --agents-per-gpu 102400 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --cell-streams 8 --sub "''" > $O/r03a_c5_full_streams8.json 2>> $O/r03a_err.log
tail -1 $O/r03a_c5_full_streams8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5 102400/cell streams 8', d['value'], d['ms_per_step'])"
eval timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --cell-streams 8 --sub "''" > $O/r03a_c5_8192_streams8.json 2>> $O/r03a_err.log
tail -1 $O/r03a_c5_8192_streams8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5 8192/cell streams 8', d['value'], d['ms_per_step'])"
tail -3 $O/r03a_err.log
