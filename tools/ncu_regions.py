#!/usr/bin/env python
"""Group a kernel's SASS into regions of equal execution count: python tools/ncu_regions.py file.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
ins = []
for idx, r in enumerate(rows[2:]):
    try: ins.append((idx, float(r[ci["Instructions Executed"]]), float(r[ci["Avg. Threads Executed"]] or 0), float(r[ci["# Samples"]]), r[ci["Source"]]))
    except Exception: pass
tot = sum(i[1] for i in ins); tots = sum(i[3] for i in ins)
print("total warp-instr %.3e over %d static instr; samples %d" % (tot, len(ins), tots))
regions = []; cur = None
for idx, ex, thr, smp, src in ins:
    if cur and abs(ex - cur["ex"]) <= 0.02 * max(ex, cur["ex"]):
        cur["n"] += 1; cur["sum"] += ex; cur["thr"] += thr * ex; cur["smp"] += smp; cur["end"] = idx
    else:
        if cur: regions.append(cur)
        cur = dict(start=idx, end=idx, ex=ex, n=1, sum=ex, thr=thr * ex, smp=smp, src=src)
regions.append(cur)
for r in sorted(regions, key=lambda r: -r["smp"])[:N]:
    print("#%4d-%4d n=%3d exec=%.2e instr%%=%4.1f thr=%4.1f time%%=%4.1f | %s" % (r["start"], r["end"], r["n"], r["ex"], 100 * r["sum"] / tot, r["thr"] / max(r["sum"], 1), 100 * r["smp"] / tots, r["src"][:50]))
