#!/bin/bash
# r02n: Blackjack tables in (dealer, ace)-major row order: A/B against state order, the GPU suite, C1 counters.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r02n_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r02n_pytest.log
tail -3 $O/r02n_pytest.log
B="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
for v in main bj_staterows main bj_staterows; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  eval RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c1 $B >> $O/r02n_ab_c1_$v.json 2>> $O/r02n_err.log
  tail -1 $O/r02n_ab_c1_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c1 $v', d['value'], d['ms_per_step'])"
done
T=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
A="--workload c1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
eval timeout 300 python bench.py $A > $O/r02n_c1_step1.json 2>> $O/r02n_err.log
eval timeout 400 ncu --metrics $T --clock-control none -k regex:k_run -s 3 -c 1 --csv --log-file $O/r02n_c1_counters.csv python bench.py $A > /dev/null 2>> $O/r02n_err.log
timeout 300 python bench.py --workload c1 --steps 10 --warmup 3 --sub '' --no-e2e > $O/r02n_bench_c1.json 2>> $O/r02n_err.log; cut -c1-160 $O/r02n_bench_c1.json
tail -3 $O/r02n_err.log
