#!/usr/bin/env python
"""Eager (store 1) against lazy (store 4) trace sweeps along a training run: per 100-episode chunk, training steps/s, mean
episode length and rows swept per step — where in a run (long early episodes, short late ones) each one wins."""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
W = importlib.import_module("rl-rust_b200.workloads")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 102400
EPISODES = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
CAPS = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]     # RLB_LZ_CAP values for the lazy runs (0: the engine's own)
CELLS = [dict(env=3, agent=1, target=1, selector=0, policy=0), dict(env=3, agent=1, target=0, selector=1, policy=1),
         dict(env=0, agent=1, target=1, selector=0, policy=0)][:int(sys.argv[4]) if len(sys.argv) > 4 else 3]
out = []
for cell in CELLS:
    w = dict(W.WORKLOADS["c5"], **cell)
    for store, cap in [(1, 0)] + [(4, c) for c in CAPS]:
        os.environ.pop("RLB_LZ_CAP", None)
        if cap:
            os.environ["RLB_LZ_CAP"] = str(cap)
        eng = W.make_engine(W.combo(w, 0), W.hyper(EPISODES), N, store_kind=store)
        sums = torch.zeros((100, 4), dtype=torch.float64, device="cuda")
        chunks = []
        for k in range(EPISODES // 100):
            r = eng.train((k + 1) * 100, 100, ep_begin=k * 100, sums_out=sums)
            chunks.append(dict(ms=r["kernel_ms"], train_steps=r["train_steps"], rows=r["trace_rows"],
                               steps_per_s=r["train_steps"] / r["kernel_ms"] * 1e3, ep_len=r["train_steps"] / (100.0 * N)))
        out.append(dict(cell=W.combo_id(W.combo(w, 0)), store=store, cap=cap, chunks=chunks, total_ms=sum(c["ms"] for c in chunks)))
        eng.close()
per = 1 + len(CAPS)
for g in range(0, len(out), per):
    a, lazy = out[g], out[g + 1:g + per]
    print(a["cell"], "eager %.0f ms;" % a["total_ms"], " ".join("lazy(cap %d) %.0f ms" % (b["cap"], b["total_ms"]) for b in lazy))
    for k, x in enumerate(a["chunks"]):
        assert all(b["chunks"][k]["train_steps"] == x["train_steps"] for b in lazy)
        print("  ep %4d..  len %6.1f rows/step %5.1f  eager %.2e/s  lazy" % (k * 100, x["ep_len"], x["rows"] / max(1, x["train_steps"]), x["steps_per_s"]),
              " ".join("%.2e (x%.2f)" % (b["chunks"][k]["steps_per_s"], b["chunks"][k]["steps_per_s"] / x["steps_per_s"]) for b in lazy))
print(json.dumps(dict(agents=N, episodes=EPISODES, runs=out)))
