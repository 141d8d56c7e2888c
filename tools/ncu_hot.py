#!/usr/bin/env python
"""Top SASS instructions by stall samples: python tools/ncu_hot.py file.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
data = []
for idx, r in enumerate(rows[2:]):
    try: data.append((float(r[ci["# Samples"]]), idx, r))
    except Exception: pass
tot = sum(d[0] for d in data); ninst = sum(float(d[2][ci["Instructions Executed"]] or 0) for d in data)
print("instructions in kernel:", len(data), "total samples", tot, "warp-instr executed %.3e" % ninst)
for v, idx, r in sorted(data, key=lambda x: -x[0])[:N]:
    print("%5.1f%% #%4d exec=%10s thr=%5s | %s" % (100 * v / tot, idx, r[ci["Instructions Executed"]], r[ci["Avg. Threads Executed"]][:5], r[ci["Source"]][:100]))
