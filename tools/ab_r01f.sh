#!/bin/bash
# A/B of the RNG window rework (r01f): parity first, then same-box bench lines per library, instruction counts.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r01f_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r01f_pytest.log
tail -3 $O/r01f_pytest.log
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for v in base r01e main base r01e main; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c4 $B >> $O/r01f_ab_c4_$v.json 2>> $O/r01f_ab_err.log
  tail -1 $O/r01f_ab_c4_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 $v', d['value'], d['ms_per_step'])"
done
for wl in c2 c3 c1; do for v in r01e main; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-e2e >> $O/r01f_ab_${wl}_$v.json 2>> $O/r01f_ab_err.log
  tail -1 $O/r01f_ab_${wl}_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl $v', d['value'], d['ms_per_step'])"
done; done
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum
for v in r01e main; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  RLB_LIB=$PWD/$lib timeout 300 ncu --metrics $M --clock-control none -k regex:k_run -c 9 --csv --log-file $O/r01f_inst_$v.csv \
     python bench.py --workload c4 --agents-per-gpu 131072 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> $O/r01f_ab_err.log
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_run -s 7 -c 1 -f -o $O/r01f_c4_k_run \
     python bench.py --workload c4 --agents-per-gpu 1048576 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/r01f_ncu_full.log 2>&1
ls -la $O | head -40
