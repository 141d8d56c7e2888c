#!/usr/bin/env python
"""Why C4 (Taxi one-step) keeps its tables in HBM instead of staging per-phase slices in shared memory.

SURVEY H5(a) / VERDICT r01 #4 propose: an episode phase (passenger, destination fixed) touches only the 25 rows of its
taxi positions (600 B of f32), so stage that slice per agent on chip and reload it when the phase changes.  The cost of
that design is set by how often the phase changes.  This script measures it on the bench's own C4 run (oracle, same
seeds): env transitions per phase change over the 1000-episode run, by hundred-episode chunk, and the DRAM bytes per
env step the slice store would HAVE to move (600 B out + 600 B in per change; 24-byte rows) — to put beside the
measured DRAM traffic of the current HBM-store kernel (profiles/counters.json: c4 dram_bytes_per_env_step).
CPU only; writes profiles/taxi_phase_stats.json."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle_py as O
import parity as P

c = dict(env=3, agent=0, selector=0, policy=0, target=1, real=0)
n_ep, n_agents, chunk = 1000, 32, 100
h = P.hyper(n_ep)
cfg = P.oracle_config(c, h)
steps = np.zeros(n_ep // chunk); changes = np.zeros(n_ep // chunk)
for i in range(n_agents):
    s = O.Session(cfg, i)
    s.record()
    for k in range(n_ep // chunk):
        s.train((k + 1) * chunk, n_ep // 10, ep_begin=k * chunk)
        tr = s.trajectory()                      # every transition of the chunk: training and injected evaluate alike
        phase = (tr["obs"] % 20).astype(np.int64)    # state = ((row*5+col)*5+pass)*4+dest -> pass*4+dest
        fresh = tr["kind"] == 0                  # reset: a new episode's first observation = a (re)load of the slice
        prev = np.concatenate([[-1], phase[:-1]])
        # a truncated step reports observation 0 without moving (taxi.rs:148-151): its row is read in place, no reload
        chg = fresh | ((phase != prev) & ~fresh & ~((tr["terminated"] == 1) & (tr["obs"] == 0)))
        steps[k] += (~fresh).sum(); changes[k] += chg.sum()
    s.close()
per_change = steps / changes
slice_bytes = 2 * 25 * 24
cur = None
try:
    t = json.load(open(os.path.join(ROOT, "profiles", "counters.json")))["kernels"]
    cur = [x for x in t if x["workload"] == "c4" and x["dtype"] == "f32"][-1]["dram_bytes_per_env_step"]
except Exception:
    pass
out = {"workload": "c4: Taxi one-step Q-learning, eps-greedy, 1000-episode run (eval_at 100), %d agents" % n_agents,
       "env_steps_per_phase_change_by_chunk": [round(float(x), 2) for x in per_change],
       "env_steps_per_phase_change_overall": float(steps.sum() / changes.sum()),
       "slice_store_min_dram_bytes_per_env_step_by_chunk": [round(slice_bytes / float(x), 1) for x in per_change],
       "slice_store_min_dram_bytes_per_env_step_overall": slice_bytes / float(steps.sum() / changes.sum()),
       "hbm_store_measured_dram_bytes_per_env_step": cur,
       "note": "slice store: 25 rows x 24 B written back + 25 rows read per phase change (episode start, pick-up, drop-off elsewhere)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "taxi_phase_stats.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
