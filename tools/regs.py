#!/usr/bin/env python
"""Summarise ptxas -v logs under rl-rust_b200/build: registers / spills / smem per k_run variant."""
import glob, re, subprocess, sys
pat = sys.argv[1] if len(sys.argv) > 1 else "k_run"
for f in sorted(glob.glob("rl-rust_b200/build/*.ptxas.log")):
    txt = open(f).read().split("Compiling entry function")
    for blk in txt[1:]:
        m = re.search(r"'(\S+)'", blk)
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        if pat not in name:
            continue
        regs = re.search(r"Used (\d+) registers", blk)
        spill = re.search(r"(\d+) bytes spill stores", blk)
        print("%-70s regs=%s spill=%s" % (name.split("(")[0][-70:], regs.group(1) if regs else "?", spill.group(1) if spill else "?"))
