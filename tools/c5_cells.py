#!/usr/bin/env python
"""Per-cell cost of the C5 sweep: every one of the 80 cells alone on the GPU at BASELINE's size (102 400 agents), three
100-episode chunks after one warm-up chunk; prints training steps/s, env steps/s and the kernel's share of the sweep."""
import importlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
W = importlib.import_module("rl-rust_b200.workloads")
w = W.WORKLOADS["c5"]; N = int(sys.argv[1]) if len(sys.argv) > 1 else 102400
ONLY = sys.argv[2] if len(sys.argv) > 2 else ""      # substring of the cell id, e.g. "onestep-ucb"
rows = []
for cell in w["cells"]:
    c = dict(w, **cell)
    if ONLY not in W.combo_id(W.combo(c, 0)):
        continue
    eng = W.make_engine(W.combo(c, 0), W.workload_hyper(c), N)
    sums = torch.zeros((100, 4), dtype=torch.float64, device="cuda")
    eng.train(100, 100, ep_begin=0, sums_out=sums)
    ms = ts = es = 0
    for k in range(1, 4):
        r = eng.train((k + 1) * 100, 100, ep_begin=k * 100, sums_out=sums)
        ms += r["kernel_ms"]; ts += r["train_steps"]; es += r["eval_steps"]
    rows.append(dict(cell=W.combo_id(W.combo(c, 0)), store=eng.store_kind(), kernel_ms=ms, train_steps=ts, env_steps=ts + es,
                     train_steps_per_s=ts / ms * 1e3, env_steps_per_s=(ts + es) / ms * 1e3))
    eng.close()
tot = sum(r["kernel_ms"] for r in rows)
for r in sorted(rows, key=lambda r: -r["kernel_ms"]):
    print("%-52s store=%d %8.1f ms %5.1f%%  train %.2e/s  env %.2e/s" % (r["cell"], r["store"], r["kernel_ms"], 100 * r["kernel_ms"] / tot, r["train_steps_per_s"], r["env_steps_per_s"]))
print(json.dumps({"agents_per_cell": N, "total_kernel_ms": tot, "total_train_steps": sum(r["train_steps"] for r in rows),
                  "serial_train_steps_per_s": sum(r["train_steps"] for r in rows) / tot * 1e3, "cells": rows}))
