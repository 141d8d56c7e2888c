#!/bin/bash
# r03m: last check of the committed build: lazy-store parity tests and smoke().
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 100 python -m pytest tests/test_gpu_abi2.py -m gpu -q -x -k lazy > $O/r03m_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $O/r03m_pytest.log | cut -c1-200
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
