#!/bin/bash
# r02e: shared-memory RNG window for Blackjack (A/B against the register window), 256-bit row loads (A/B on C3 and on the
# f64 headline), the whole GPU suite on this build, the e2e timing probe.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r02e_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r02e_pytest.log
tail -4 $O/r02e_pytest.log
B="--steps 4 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
for v in main bj_regwin main bj_regwin; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  eval RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c1 $B >> $O/r02e_ab_c1_$v.json 2>> $O/r02e_err.log
  tail -1 $O/r02e_ab_c1_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c1 $v', d['value'], d['ms_per_step'])"
done
for v in main no_ld256 main no_ld256; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  eval RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c3 $B >> $O/r02e_ab_c3_$v.json 2>> $O/r02e_err.log
  tail -1 $O/r02e_ab_c3_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3 $v', d['value'], d['ms_per_step'])"
  eval RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c3 --real f64 --agents-per-gpu 2097152 $B >> $O/r02e_ab_c3f64_$v.json 2>> $O/r02e_err.log
  tail -1 $O/r02e_ab_c3f64_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3 f64 $v', d['value'], d['ms_per_step'])"
done
timeout 300 python tools/e2e_probe.py c4 > $O/r02e_e2e_probe_c4.json 2>> $O/r02e_err.log; cat $O/r02e_e2e_probe_c4.json
timeout 300 python tools/e2e_probe.py c2 > $O/r02e_e2e_probe_c2.json 2>> $O/r02e_err.log; cat $O/r02e_e2e_probe_c2.json
tail -5 $O/r02e_err.log
