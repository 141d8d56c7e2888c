#!/usr/bin/env python
"""Condense same-box A/B bench lines (gpurun_out/<tag>_ab_<workload>_<variant>.json, one JSON line per run) into a table:
python tools/ab_summary.py TAG > profiles/TAG_ab_same_box.txt"""
import glob, json, os, re, sys
tag = sys.argv[1]
rows = {}
for f in sorted(glob.glob("gpurun_out/%s_ab_*.json" % tag)):
    m = re.match(r".*%s_ab_([a-z0-9]+?)_(.+)\.json" % tag, f)
    wl, var = m.group(1), m.group(2)
    for line in open(f):
        line = line.strip()
        if line.startswith("{"):
            d = json.loads(line)
            rows.setdefault((wl, var), []).append((d["value"], d["ms_per_step"], d["config"]["agents_per_gpu"], d["dtype"]))
print("# %s: same-box A/B (bench.py --steps 4 --warmup 3, one line per run, runs interleaved main / variant / main / variant)" % tag)
print("%-8s %-16s %-10s %-6s %s" % ("workload", "variant", "agents", "dtype", "training steps/s per run (ms per step)"))
for (wl, var), runs in sorted(rows.items()):
    print("%-8s %-16s %-10d %-6s %s" % (wl, var, runs[0][2], runs[0][3], "   ".join("%.3e (%.0f)" % (v, ms) for v, ms, _, _ in runs)))
