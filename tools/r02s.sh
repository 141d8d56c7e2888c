#!/bin/bash
# r02s: lazy store after the row-limit fallback: failing tests, eager vs lazy along a run, ncu lane/issue counters of both.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_abi2.py tests/test_gpu_dyna.py -m gpu -q > $O/r02s_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $O/r02s_pytest.log | cut -c1-200
timeout 900 python tools/lazy_phase.py 102400 1000 > $O/r02s_lazy_phase.txt 2> $O/r02s_err.log; grep -v '^{' $O/r02s_lazy_phase.txt | cut -c1-160
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes.sum,lts__t_bytes.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:k_run -c 24 --csv --log-file $O/r02s_ncu_lazy_phase.csv python tools/lazy_phase.py 102400 200 > $O/r02s_ncu.log 2>&1
echo "ncu exit $?"
tail -3 $O/r02s_err.log
