#!/bin/bash
# A/B of the one-step HBM-path changes (r01e): parity first, then same-box bench lines per library variant,
# instruction counts, and one full ncu capture with source.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r01e_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r01e_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r01e_pytest.log
tail -3 $O/r01e_pytest.log
B="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
for v in base main nolazy nocarry base main; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c4 $B >> $O/r01e_ab_c4_$v.json 2>> $O/r01e_ab_err.log
  tail -1 $O/r01e_ab_c4_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'])"
done
for wl in c2 c3; do
  timeout 300 python bench.py --workload $wl $B > $O/r01e_main_$wl.json 2>> $O/r01e_ab_err.log; tail -1 $O/r01e_main_$wl.json | cut -c1-200
done
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed_op_shared_ld.sum
for v in base main nolazy nocarry; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  RLB_LIB=$PWD/$lib timeout 300 ncu --metrics $M --clock-control none -k regex:k_run -c 3 --csv --log-file $O/r01e_inst_$v.csv \
     python bench.py --workload c4 --agents-per-gpu 262144 --steps 1 --warmup 2 --no-cpu-baseline --no-e2e > /dev/null 2>> $O/r01e_ab_err.log
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_run -s 2 -c 1 -f -o $O/r01e_c4_k_run \
     python bench.py --workload c4 --agents-per-gpu 1048576 --steps 1 --warmup 2 --no-cpu-baseline --no-e2e > $O/r01e_ncu_full.log 2>&1
ls -la $O
