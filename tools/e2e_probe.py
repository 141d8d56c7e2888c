#!/usr/bin/env python
"""Where does the e2e leg lose time?  The same six chunks (episodes 300..899 of a fresh run) timed three ways on fresh
engines: device-resident outputs, the blocking call with pinned host record buffers, the asynchronous call with two
buffers taking turns.  Prints per-call host blocking times and the library's kernel times."""
import importlib, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
W = importlib.import_module("rl-rust_b200.workloads")
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
w = W.WORKLOADS[name]; N = w["agents_per_gpu"]; chunk = w["chunk"]; n_ep = w["n_episodes"]; eval_at = n_ep // 10
sums_dev = torch.zeros((chunk, 4), dtype=torch.float64, device="cuda")
recs = [torch.zeros((chunk, N, 4), dtype=torch.int32).pin_memory() for _ in range(2)]
sums_host = torch.zeros((chunk, 4), dtype=torch.float64).pin_memory()
def run(mode, steps=6, warm=3):
    eng = W.make_engine(W.combo(w, 0), W.workload_hyper(w), N)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    for k in range(warm):
        eng.train((k + 1) * chunk, eval_at, ep_begin=k * chunk, sums_out=sums_dev)
    torch.cuda.synchronize(); t0 = time.perf_counter(); marks = []; kms = []
    for k in range(warm, warm + steps):
        a = time.perf_counter()
        if mode == "device": kms.append(eng.train((k + 1) * chunk, eval_at, ep_begin=k * chunk, sums_out=sums_dev)["kernel_ms"])
        elif mode == "sync": kms.append(eng.train((k + 1) * chunk, eval_at, ep_begin=k * chunk, sums_out=sums_host, episodes_out=recs[k & 1])["kernel_ms"])
        else: eng.train((k + 1) * chunk, eval_at, ep_begin=k * chunk, sums_out=sums_host, episodes_out=recs[k & 1], wait=False)
        marks.append(round((time.perf_counter() - a) * 1e3, 1))
    if mode == "async":
        a = time.perf_counter(); res = eng.train_wait(); marks.append(round((time.perf_counter() - a) * 1e3, 1))
        kms = [r["kernel_ms"] for r in res]
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3 / steps
    eng.close()
    return {"mode": mode, "ms_per_step": round(dt, 1), "kernel_ms_per_step": round(sum(kms) / steps, 1), "host_ms_in_each_call": marks}
out = [run("device"), run("sync"), run("async"), run("device")]
print(json.dumps({"workload": name, "record_bytes_per_step": chunk * N * 16, "runs": out}))
