#!/bin/bash
# A/B of the hybrid-store sweep rework (r01g): parity first, then same-box C2 lines per library, then ncu.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r01g_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r01g_pytest.log
tail -3 $O/r01g_pytest.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
for v in fl_notouch main fl_u6 fl_u8 fl_notouch main fl_u6 fl_u8; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c2 $B >> $O/r01g_ab_c2_$v.json 2>> $O/r01g_ab_err.log
  tail -1 $O/r01g_ab_c2_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2 $v', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
for v in fl_notouch main; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c2 --real f64 $B >> $O/r01g_ab_c2f64_$v.json 2>> $O/r01g_ab_err.log
  tail -1 $O/r01g_ab_c2f64_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2 f64 $v', d['value'], d['ms_per_step'])"
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_run -s 4 -c 1 -f -o $O/r01g_c2_k_run \
     python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r01g_ncu_full.log 2>&1
ls -la $O | grep r01g
