#!/usr/bin/env python
"""Where does the e2e leg lose time?  Per-step host timings of the asynchronous train call with pinned host record
buffers against the device-resident call (C4 by default)."""
import importlib, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
W = importlib.import_module("rl-rust_b200.workloads")
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
w = W.WORKLOADS[name]; N = w["agents_per_gpu"]; chunk = w["chunk"]; n_ep = w["n_episodes"]; eval_at = n_ep // 10
eng = W.make_engine(W.combo(w, 0), W.workload_hyper(w), N)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
sums_dev = torch.zeros((chunk, 4), dtype=torch.float64, device="cuda")
recs = [torch.zeros((chunk, N, 4), dtype=torch.int32).pin_memory() for _ in range(2)]
sums_host = torch.zeros((chunk, 4), dtype=torch.float64).pin_memory()
def run(mode, steps, k0):
    torch.cuda.synchronize(); t0 = time.perf_counter(); marks = []
    for k in range(k0, k0 + steps):
        c = k % (n_ep // chunk)
        if c == 0 and k > 0: eng.agent_reset()
        a = time.perf_counter()
        if mode == "device": eng.train((c + 1) * chunk, eval_at, ep_begin=c * chunk, sums_out=sums_dev)
        elif mode == "sync": eng.train((c + 1) * chunk, eval_at, ep_begin=c * chunk, sums_out=sums_host, episodes_out=recs[k & 1])
        else: eng.train((c + 1) * chunk, eval_at, ep_begin=c * chunk, sums_out=sums_host, episodes_out=recs[k & 1], wait=False)
        marks.append(round((time.perf_counter() - a) * 1e3, 1))
    if mode == "async":
        a = time.perf_counter(); res = eng.train_wait(); marks.append(round((time.perf_counter() - a) * 1e3, 1))
        kms = [round(r["kernel_ms"], 1) for r in res]
    else:
        kms = None
    torch.cuda.synchronize()
    return {"mode": mode, "ms_per_step": round((time.perf_counter() - t0) * 1e3 / steps, 1), "host_ms_in_each_call": marks, "kernel_ms": kms}
run("device", 3, 0)
out = [run("device", 6, 3), run("sync", 6, 3), run("async", 6, 3), run("device", 6, 3)]
print(json.dumps({"workload": name, "record_bytes_per_step": chunk * N * 16, "runs": out}))
