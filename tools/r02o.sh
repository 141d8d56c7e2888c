#!/bin/bash
# r02o: the whole GPU suite on the final tree (golden v2, GPU f32-vs-f64 bounds, Blackjack row order) and smoke().
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r02o_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r02o_pytest.log
tail -4 $O/r02o_pytest.log
timeout 600 python __graft_entry__.py smoke > $O/r02o_smoke.log 2>&1; echo "smoke exit $?"; tail -4 $O/r02o_smoke.log
timeout 300 python -m pytest tests/test_gpu_f32_vs_f64.py -m gpu -q -s 2>&1 | grep "updates" > $O/r02o_f32_vs_f64.txt; cat $O/r02o_f32_vs_f64.txt
