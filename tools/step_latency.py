#!/usr/bin/env python
"""Latency of ONE agent-step through the trait-object style entry points (N = 1 and a few batch sizes):
the fused rlb_agent_step (one launch, one wait) against the three step-level calls a loop written like agent.rs:86-106
makes (rlb_env_step + rlb_agent_get_action + rlb_agent_update).  Prints one JSON line."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rlb = importlib.import_module("rl-rust_b200")
W = importlib.import_module("rl-rust_b200.workloads")
out = {}
for n in (1, 1024, 1 << 20):
    c = dict(env=3, agent=0, selector=0, policy=0, target=1, real=0)
    h = W.hyper(1000)
    with W.make_engine(c, h, n) as eng:
        iters = 2000 if n <= 1024 else 200
        for _ in range(50):
            eng.agent_step()
        t0 = time.perf_counter()
        for _ in range(iters):
            eng.agent_step()
        fused = (time.perf_counter() - t0) / iters
    with W.make_engine(c, h, n) as eng:
        obs = eng.env_reset(); act = eng.get_action(obs)
        def one(obs, act):
            try:
                o2, r, t = eng.env_step(act)
            except rlb.EnvNotReady:
                o2 = eng.env_reset(); a2 = eng.get_action(o2); return o2, a2
            a2 = eng.get_action(o2)
            eng.update(obs, act, r, t, o2, a2)
            if t.any():
                o2 = eng.env_reset(); a2 = eng.get_action(o2)
            return o2, a2
        for _ in range(20):
            obs, act = one(obs, act)
        t0 = time.perf_counter()
        for _ in range(iters // 4):
            obs, act = one(obs, act)
        three = (time.perf_counter() - t0) / (iters // 4)
    out["n_agents_%d" % n] = {"fused_rlb_agent_step_us": fused * 1e6, "three_calls_us": three * 1e6,
                              "fused_agent_steps_per_s": n / fused}
print(json.dumps({"workload": "Taxi Q-learning eps-greedy f32, host (numpy) buffers, one Python call per transition", "latency": out}))
