#!/bin/bash
# r02h: after the window-slide fix (the generalised RNG window had pushed w[] into local memory in the voted-refill kernels:
# Taxi 292 -> 1130 B of spills, C4 3.7x slower in r02e / r02g — those two runs' C4 / C1-variant numbers are void).
# Full GPU suite, Blackjack A/B (register window vs shared-memory ring), default bench line, counters of all five
# configurations with the final kernels, C1 full capture.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r02h_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r02h_pytest.log
tail -3 $O/r02h_pytest.log
B="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
for v in main bj_regwin main bj_regwin; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  eval RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c1 $B >> $O/r02h_ab_c1_$v.json 2>> $O/r02h_err.log
  tail -1 $O/r02h_ab_c1_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c1 $v', d['value'], d['ms_per_step'])"
done
timeout 900 python bench.py --steps 8 --warmup 3 > $O/r02h_bench.json 2> $O/r02h_bench.err; echo "bench exit $?"; cut -c1-200 $O/r02h_bench.json; tail -3 $O/r02h_bench.err
T=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
for w in c2 c4 c3 c1 c2_f64; do
  base=${w%%_*}; real=f32; [ "$w" != "$base" ] && real=f64
  A="--workload $base --real $real --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
  eval timeout 300 python bench.py $A > $O/r02h_${w}_step1.json 2>> $O/r02h_err.log
  eval timeout 400 ncu --metrics $T --clock-control none -k regex:k_run -s 3 -c 1 --csv --log-file $O/r02h_${w}_counters.csv python bench.py $A > /dev/null 2>> $O/r02h_err.log
done
eval timeout 500 ncu --set full --import-source on --clock-control none -k regex:k_run -s 3 -c 1 -f -o $O/r02h_c1_k_run python bench.py --workload c1 --agents-per-gpu 524288 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sub "''" > $O/r02h_ncu_c1.log 2>&1
eval timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02h_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r02h_ncu_launches.log 2>&1
tail -5 $O/r02h_err.log
