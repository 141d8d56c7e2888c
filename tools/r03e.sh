#!/bin/bash
# r03e: UCB one-step kernels at 5 / 4 CTAs per SM; CliffWalking Double carry at 5 / 4.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
B="--steps 6 --warmup 3 --no-cpu-baseline --no-e2e --sub ''"
for v in main all_n5 all_n4 c3_c5 c3_c4 main all_n5 all_n4 c3_c5 c3_c4; do
  lib=rl-rust_b200/ab/librlb_$v.so; [ $v = main ] && lib=rl-rust_b200/librlb.so
  eval RLB_LIB=$PWD/$lib timeout 300 python bench.py --workload c3 $B >> $O/r03e_ab_c3_$v.json 2>> $O/r03e_err.log
  tail -1 $O/r03e_ab_c3_$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3 $v', d['value'], d['ms_per_step'])"
done
for v in all_n5 all_n4; do
  lib=rl-rust_b200/ab/librlb_$v.so
  RLB_LIB=$PWD/$lib timeout 600 python tools/c5_cells.py 102400 onestep-ucb > $O/r03e_c5_cells_$v.txt 2>> $O/r03e_err.log; echo "== $v"; grep "onestep.*ucb" $O/r03e_c5_cells_$v.txt | grep -v blackjack | sort | cut -c1-140; tail -1 $O/r03e_c5_cells_$v.txt | cut -c1-160
done
tail -3 $O/r03e_err.log
