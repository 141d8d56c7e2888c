#!/bin/bash
# r02u: ncu --set full of the lazy-store Taxi trace kernel (second k_run of lazy_phase.py: store 4, episodes 0..100).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 300 python tools/lazy_phase.py 32768 100 0 1 > $O/r02u_plain.txt 2> $O/r02u_err.log; grep -v '^{' $O/r02u_plain.txt | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_run -s 1 -c 1 -f -o $O/r02u_taxi_lazy_k_run python tools/lazy_phase.py 32768 100 0 1 > $O/r02u_ncu.log 2>&1
echo "ncu exit $?"; ls -la $O/r02u_taxi_lazy_k_run.ncu-rep
tail -3 $O/r02u_err.log
