#!/bin/bash
# r02c: (1) the whole GPU suite on the ABI v2 + UCB-math build; (2) same-box A/B of the C3 kernel: r02a baseline numbers are
# in profiles/, here main (fast interleaved div/sqrt, ln table, counts-row reuse, 8 CTAs/SM) vs 6 / 5 CTAs per SM vs the
# compiler's own div/sqrt; (3) the default bench line (async e2e: one launch per call); (4) counters + a full capture of the
# new C3 kernel; (5) N = 1 step latency.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r02c_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r02c_pytest.log
tail -8 $O/r02c_pytest.log
B="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
for v in main ucb_mb6 ucb_mb5 ucb_slowmath main ucb_mb6 ucb_mb5 ucb_slowmath; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  eval RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c3 $B >> $O/r02c_ab_c3_$v.json 2>> $O/r02c_err.log
  tail -1 $O/r02c_ab_c3_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3 $v', d['value'], d['ms_per_step'])"
done
timeout 900 python bench.py --steps 8 --warmup 3 > $O/r02c_bench.json 2> $O/r02c_bench.err; echo "bench exit $?"; cut -c1-200 $O/r02c_bench.json; tail -3 $O/r02c_bench.err
T=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
for w in c3; do
  A="--workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
  eval timeout 300 python bench.py $A > $O/r02c_${w}_step1.json 2>> $O/r02c_err.log
  eval timeout 400 ncu --metrics $T --clock-control none -k regex:k_run -s 3 -c 1 --csv --log-file $O/r02c_${w}_counters.csv python bench.py $A > /dev/null 2>> $O/r02c_err.log
done
eval timeout 500 ncu --set full --import-source on --clock-control none -k regex:k_run -s 3 -c 1 -f -o $O/r02c_c3_k_run python bench.py --workload c3 --agents-per-gpu 524288 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sub "''" > $O/r02c_ncu_c3.log 2>&1
timeout 300 python tools/step_latency.py > $O/r02c_step_latency.json 2>> $O/r02c_err.log; cat $O/r02c_step_latency.json
tail -5 $O/r02c_err.log
