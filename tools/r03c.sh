#!/bin/bash
# r03c: C3 with the Double carry at 7 / 6 CTAs per SM (fewer spills) against the kernel of record.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
B="--steps 6 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
for v in main carry_mb7 carry_mb6 mb7 main carry_mb7 carry_mb6 mb7; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  eval RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c3 $B >> $O/r03c_ab_c3_$v.json 2>> $O/r03c_err.log
  tail -1 $O/r03c_ab_c3_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3 $v', d['value'], d['ms_per_step'])"
done
RLB_LIB=$PWD/rl-rust_b200/ab/librlb_carry_mb7.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k cliff > $O/r03c_pytest.log 2>&1; tail -2 $O/r03c_pytest.log
tail -3 $O/r03c_err.log
