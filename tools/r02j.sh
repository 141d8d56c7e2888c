#!/bin/bash
# r02j: Blackjack occupancy around the new default (8 CTAs/SM with the shared-memory RNG ring): 6 / 7 / 8 / 12; default
# bench line with the e2e leg aligned to the device leg's chunks and one whole run per sub-record.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
B="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
for v in main bj_mb6 bj_mb7 bj_mb12 main bj_mb6 bj_mb7 bj_mb12; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  eval RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c1 $B >> $O/r02j_ab_c1_$v.json 2>> $O/r02j_err.log
  tail -1 $O/r02j_ab_c1_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c1 $v', d['value'], d['ms_per_step'])"
done
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02j_bench.json 2> $O/r02j_bench.err; echo "bench exit $?"; cut -c1-200 $O/r02j_bench.json; tail -3 $O/r02j_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02j_ref.json 2>> $O/r02j_err.log; cut -c1-200 $O/r02j_ref.json
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q -x -k "blackjack or Blackjack or trait or traj" > $O/r02j_pytest.log 2>&1; tail -2 $O/r02j_pytest.log
tail -5 $O/r02j_err.log
