#!/bin/bash
# r02a: baseline evidence for the kernels VERDICT r01 asked about, before any round-2 change —
#   C3 (CliffWalking, Expected Sarsa, Double, UCB) bench line, full-size traffic/instruction counters, a reduced-size
#   `ncu --set full` capture with source; C1 (Blackjack) the same; a compute-sanitizer memcheck of smoke().
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r02a_gpu.txt
B="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
T=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
for w in c3 c1 c4 c2; do
  timeout 300 python bench.py --workload $w $B > $O/r02a_bench_$w.json 2>> $O/r02a_err.log; cut -c1-200 $O/r02a_bench_$w.json
  timeout 400 ncu --metrics $T --clock-control none -k regex:k_run -s 3 -c 1 --csv --log-file $O/r02a_${w}_traffic_full_size.csv python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> $O/r02a_err.log
done
timeout 500 ncu --set full --import-source on --clock-control none -k regex:k_run -s 3 -c 1 -f -o $O/r02a_c3_k_run python bench.py --workload c3 --agents-per-gpu 524288 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02a_ncu_c3.log 2>&1
timeout 500 ncu --set full --import-source on --clock-control none -k regex:k_run -s 3 -c 1 -f -o $O/r02a_c1_k_run python bench.py --workload c1 --agents-per-gpu 524288 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02a_ncu_c1.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > $O/r02a_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> $O/r02a_sanitizer_memcheck.log
tail -5 $O/r02a_sanitizer_memcheck.log
ls -la $O | grep r02a
