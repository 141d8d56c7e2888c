#!/bin/bash
# r03f: kernels of record (lazy trace store; UCB one-step kernels at 6 / 5 CTAs per SM, CliffWalking Double carry):
# full GPU suite, smoke, C3 counters, the default bench line and its reference arm, C5 cells and sweep.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r03f_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r03f_pytest.log
tail -4 $O/r03f_pytest.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
T=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
for w in c3; do
  A="--workload $w --real f32 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
  eval timeout 300 python bench.py $A > $O/r03f_${w}_step1.json 2>> $O/r03f_err.log
  eval timeout 400 ncu --metrics $T --clock-control none -k regex:k_run -s 3 -c 1 --csv --log-file $O/r03f_${w}_counters.csv python bench.py $A > /dev/null 2>> $O/r03f_err.log
done
python tools/make_counters.py r03f > /dev/null 2>> $O/r03f_err.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r03f_bench.json 2> $O/r03f_bench.err; echo "bench exit $?"; cut -c1-300 $O/r03f_bench.json; tail -3 $O/r03f_bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r03f_bench_ref.json 2> $O/r03f_bench_ref.err; echo "ref exit $?"; cut -c1-300 $O/r03f_bench_ref.json
timeout 600 python tools/c5_cells.py 102400 > $O/r03f_c5_cells.txt 2>> $O/r03f_err.log; head -10 $O/r03f_c5_cells.txt | cut -c1-140; tail -1 $O/r03f_c5_cells.txt | cut -c1-160
eval timeout 600 python bench.py --workload c5 --agents-per-gpu 102400 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --cell-streams 8 --sub "''" > $O/r03f_c5_full_streams8.json 2>> $O/r03f_err.log
tail -1 $O/r03f_c5_full_streams8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5 102400/cell streams 8', d['value'], d['ms_per_step'])"
tail -5 $O/r03f_err.log
