#!/bin/bash
# r03j: ncu launch list of the default bench command on the kernels of record (per-launch times are cold-cache and
# serialised: the kernel's SHARE of the step is what must agree with the bench line).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r03j_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r03j_ncu_launches.log 2>&1; echo "ncu exit $?"
grep -c k_run $O/r03j_launches.csv
