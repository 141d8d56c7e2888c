#!/bin/bash
# r02l: full GPU suite on the final build (NVTX ranges, single tap test, Blackjack at 8 CTAs/SM, custom maps through the
# mirrors, teacher-forced f32-vs-f64 on the GPU), then counters of all five configurations and the default bench line.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r02l_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r02l_pytest.log
tail -6 $O/r02l_pytest.log
timeout 300 python -m pytest tests/test_gpu_f32_vs_f64.py -m gpu -q -s 2>&1 | grep "updates" > $O/r02l_f32_vs_f64.txt; cat $O/r02l_f32_vs_f64.txt
T=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
for w in c2 c4 c3 c1 c2_f64; do
  base=${w%%_*}; real=f32; [ "$w" != "$base" ] && real=f64
  A="--workload $base --real $real --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
  eval timeout 300 python bench.py $A > $O/r02l_${w}_step1.json 2>> $O/r02l_err.log
  eval timeout 400 ncu --metrics $T --clock-control none -k regex:k_run -s 3 -c 1 --csv --log-file $O/r02l_${w}_counters.csv python bench.py $A > /dev/null 2>> $O/r02l_err.log
done
eval timeout 500 ncu --set full --import-source on --clock-control none -k regex:k_run -s 3 -c 1 -f -o $O/r02l_c1_k_run python bench.py --workload c1 --agents-per-gpu 524288 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sub "''" > $O/r02l_ncu_c1.log 2>&1
eval timeout 500 ncu --set full --import-source on --clock-control none -k regex:k_run -s 3 -c 1 -f -o $O/r02l_c2_k_run python bench.py --workload c2 --agents-per-gpu 262144 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sub "''" > $O/r02l_ncu_c2.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02l_bench.json 2> $O/r02l_bench.err; echo "bench exit $?"; cut -c1-200 $O/r02l_bench.json; tail -3 $O/r02l_bench.err
timeout 300 python bench.py --workload c1 --steps 10 --warmup 3 --sub '' > $O/r02l_bench_c1.json 2>> $O/r02l_err.log; cut -c1-160 $O/r02l_bench_c1.json
tail -5 $O/r02l_err.log
