#!/bin/bash
# r02b: ABI v2 on the GPU — the whole parity suite, the default bench line with its sub-records, per-configuration ncu
# counters for the roofline fractions (profiles/counters.json), the N = 1 step latency.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02b_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r02b_pytest.log
tail -5 $O/r02b_pytest.log
timeout 600 python bench.py --steps 8 --warmup 3 > $O/r02b_bench.json 2> $O/r02b_bench.err; echo "bench exit $?"; cut -c1-250 $O/r02b_bench.json; tail -3 $O/r02b_bench.err
T=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
for w in c2 c4 c3 c1 c2_f64; do
  base=${w%%_*}; real=f32; [ "$w" != "$base" ] && real=f64
  A="--workload $base --real $real --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
  eval timeout 300 python bench.py $A > $O/r02b_${w}_step1.json 2>> $O/r02b_err.log
  eval timeout 400 ncu --metrics $T --clock-control none -k regex:k_run -s 3 -c 1 --csv --log-file $O/r02b_${w}_counters.csv python bench.py $A > /dev/null 2>> $O/r02b_err.log
  cut -c1-120 $O/r02b_${w}_step1.json
done
timeout 300 python tools/step_latency.py > $O/r02b_step_latency.json 2>> $O/r02b_err.log; cat $O/r02b_step_latency.json
ls -la $O | grep r02b
