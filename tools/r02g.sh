#!/bin/bash
# r02g: device->host copies of asynchronous calls off the main stream (e2e), probes; default bench line; counters and full
# captures of the new C1 / C3 kernels.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_abi2.py tests/test_gpu_api.py tests/test_gpu_mirror.py tests/test_driver.py -m gpu -q -x > $O/r02g_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r02g_pytest.log
tail -3 $O/r02g_pytest.log
timeout 300 python tools/e2e_probe.py c2 > $O/r02g_e2e_probe_c2.json 2>> $O/r02g_err.log; cat $O/r02g_e2e_probe_c2.json
timeout 300 python tools/e2e_probe.py c4 > $O/r02g_e2e_probe_c4.json 2>> $O/r02g_err.log; cat $O/r02g_e2e_probe_c4.json
timeout 900 python bench.py --steps 8 --warmup 3 > $O/r02g_bench.json 2> $O/r02g_bench.err; echo "bench exit $?"; cut -c1-200 $O/r02g_bench.json; tail -3 $O/r02g_bench.err
T=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
for w in c3 c1; do
  A="--workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
  eval timeout 300 python bench.py $A > $O/r02g_${w}_step1.json 2>> $O/r02g_err.log
  eval timeout 400 ncu --metrics $T --clock-control none -k regex:k_run -s 3 -c 1 --csv --log-file $O/r02g_${w}_counters.csv python bench.py $A > /dev/null 2>> $O/r02g_err.log
  eval timeout 500 ncu --set full --import-source on --clock-control none -k regex:k_run -s 3 -c 1 -f -o $O/r02g_${w}_k_run python bench.py --workload $w --agents-per-gpu 524288 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sub "''" > $O/r02g_ncu_$w.log 2>&1
done
tail -5 $O/r02g_err.log
