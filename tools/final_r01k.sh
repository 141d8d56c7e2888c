#!/bin/bash
# r01k: final verification of the round's kernels — parity, bench lines (value / e2e / roofline / cpu_baseline), the ncu launch
# list of the default bench command, full-size DRAM traffic of the dominant kernel, reduced-size full captures with source.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r01k_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r01k_pytest.log
tail -3 $O/r01k_pytest.log
B="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
for v in main bj_mb10 bj_mb8 main bj_mb10 bj_mb8; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c1 $B >> $O/r01k_ab_c1_$v.json 2>> $O/r01k_err.log
  tail -1 $O/r01k_ab_c1_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c1 $v', d['value'], d['ms_per_step'])"
done
for v in main fl_s5u3h fl_s6u2h main fl_s5u3h fl_s6u2h; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c2 $B >> $O/r01k_ab_c2_$v.json 2>> $O/r01k_err.log
  tail -1 $O/r01k_ab_c2_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2 $v', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
timeout 600 python bench.py > $O/r01k_bench_c2.json 2>> $O/r01k_err.log; cut -c1-160 $O/r01k_bench_c2.json
timeout 600 python bench.py --workload c4 > $O/r01k_bench_c4.json 2>> $O/r01k_err.log; cut -c1-160 $O/r01k_bench_c4.json
timeout 600 python bench.py --workload c2 --real f64 --steps 4 --no-cpu-baseline > $O/r01k_bench_c2_f64.json 2>> $O/r01k_err.log; cut -c1-160 $O/r01k_bench_c2_f64.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01k_c2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r01k_ncu_launches.log 2>&1
T=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_issued.avg.pct_of_peak_sustained_active
timeout 300 ncu --metrics $T --clock-control none -k regex:k_run -s 3 -c 1 --csv --log-file $O/r01k_c2_traffic_full_size.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> $O/r01k_err.log
timeout 300 ncu --metrics $T --clock-control none -k regex:k_run -s 3 -c 1 --csv --log-file $O/r01k_c4_traffic_full_size.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>> $O/r01k_err.log
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_run -s 3 -c 1 -f -o $O/r01k_c2_k_run python bench.py --agents-per-gpu 262144 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/r01k_ncu_c2.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_run -s 3 -c 1 -f -o $O/r01k_c4_k_run python bench.py --workload c4 --agents-per-gpu 524288 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/r01k_ncu_c4.log 2>&1
ls -la $O | grep r01k
