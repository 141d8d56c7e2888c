#!/bin/bash
# r02i: device-resident asynchronous calls overlap (bench device leg back to back); Blackjack occupancy with the
# shared-memory RNG ring (12 / 10 / 8 CTAs per SM); default bench line.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_abi2.py tests/test_gpu_api.py tests/test_gpu_mirror.py tests/test_cpp_mirror.py tests/test_gpu_parity.py -m gpu -q -x > $O/r02i_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r02i_pytest.log
tail -3 $O/r02i_pytest.log
B="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
for v in main bj_mb10 bj_mb8 bj_regwin main bj_mb10 bj_mb8 bj_regwin; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  eval RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c1 $B >> $O/r02i_ab_c1_$v.json 2>> $O/r02i_err.log
  tail -1 $O/r02i_ab_c1_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c1 $v', d['value'], d['ms_per_step'])"
done
timeout 900 python bench.py --steps 8 --warmup 3 > $O/r02i_bench.json 2> $O/r02i_bench.err; echo "bench exit $?"; cut -c1-200 $O/r02i_bench.json; tail -3 $O/r02i_bench.err
tail -5 $O/r02i_err.log
