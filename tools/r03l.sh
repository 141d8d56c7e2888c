#!/bin/bash
# r03l: lazy store, TD-history capacity (RLB_LZ_CAP) now that the history-full flush is cooperative too: a smaller history
# leaves more of the SM's 256 KB to L1 (101 sweeps x 128 threads x 4 B = 52 KB per CTA).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 200 python tools/lazy_phase.py 102400 300 0,80,64,48,32 2 > $O/r03l_lazy_caps.txt 2> $O/r03l_err.log; grep -v '^{' $O/r03l_lazy_caps.txt | cut -c1-230
tail -3 $O/r03l_err.log
