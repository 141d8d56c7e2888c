#!/bin/bash
# r03i: ncu --set full of the two kernels this round changed last: C3 (CliffWalking Double UCB, 5 CTAs/SM, carry) and the
# lazy-store Taxi Q(lambda) kernel of record (cooperative flush, rows handed on in registers).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
eval timeout 500 ncu --set full --import-source on --clock-control none -k regex:k_run -s 3 -c 1 -f -o $O/r03i_c3_k_run python bench.py --workload c3 --agents-per-gpu 524288 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sub "''" > $O/r03i_ncu_c3.log 2>&1; echo "ncu c3 exit $?"
timeout 500 ncu --set full --import-source on --clock-control none -k regex:k_run -s 1 -c 1 -f -o $O/r03i_taxi_lazy_k_run python tools/lazy_phase.py 102400 100 0 1 > $O/r03i_ncu_lazy.log 2>&1; echo "ncu lazy exit $?"
ls -la $O/r03i_*.ncu-rep
