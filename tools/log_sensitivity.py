#!/usr/bin/env python
"""How far does the choice of `ln` reach?  

The reference computes the UCB bonus with Rust's f64::ln = the platform libm's log (upper_confidence_bound.rs:36,56);
engine and oracle share one portable fdlibm-style routine so that CPU and GPU agree bit for bit.  This script measures,
on THIS machine's glibc: (1) for how many integer t <= T_MAX the two logs differ (and by how many ulps), and (2) what
fraction of UCB agents end a reference-style run with ANY different result (episode lengths, returns, Q tables, counts)
when the oracle is rebuilt with glibc's log (liboracle_libmlog.so) — per env x rule x policy.  CPU only; writes
profiles/log_sensitivity.json."""
import ctypes as C, json, math, os, struct, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle_py as O
import parity as P

T_MAX = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
L = O.lib()
def bits(x): return struct.unpack("<q", struct.pack("<d", x))[0]
diff, worst, first = 0, 0, []
for t in range(1, T_MAX + 1):
    a, b = L.oracle_log(float(t)), math.log(float(t))
    if a != b:
        diff += 1
        d = abs(bits(a) - bits(b)); worst = max(worst, d)
        if len(first) < 10: first.append(t)
out = {"t_max": T_MAX, "differing_t": diff, "fraction": diff / T_MAX, "max_ulps": worst, "first_differing_t": first,
       "libm": "glibc " + os.confstr("CS_GNU_LIBC_VERSION")}
print(json.dumps(out))

# (2) trajectories: swap the library under oracle_py and rerun
def run(libname, c, h, n_agents, n_ep, eval_at):
    O._LIB = None
    so = os.path.join(ROOT, "oracle", libname)
    real_build = O.build
    O.build = lambda force=False: so
    try:
        return O.batch_train(P.oracle_config(c, h), 0, n_agents, n_ep, eval_at, n_threads=8)
    finally:
        O.build = real_build; O._LIB = None
cells = []
n_agents, n_ep = 256, 200
for env in (0, 1, 2, 3):
    for agent, target in ((0, 0), (0, 1), (0, 2), (1, 0)):
        for pol in (0, 1):
            c = dict(env=env, agent=agent, selector=1, policy=pol, target=target, real=1)
            h = P.hyper(n_ep)
            a = run("liboracle.so", c, h, n_agents, n_ep, n_ep // 10)
            b = run("liboracle_libmlog.so", c, h, n_agents, n_ep, n_ep // 10)
            changed = 0
            for i in range(n_agents):
                same = (np.array_equal(a["len"][i], b["len"][i]) and P.bits_equal(a["ret"][i], b["ret"][i]) and P.bits_equal(a["q"][i], b["q"][i])
                        and np.array_equal(a["counts"][i], b["counts"][i]))
                changed += not same
            traj = sum(not np.array_equal(a["len"][i], b["len"][i]) or not np.array_equal(a["counts"][i], b["counts"][i]) for i in range(n_agents))
            cells.append({"cell": P.combo_id(c), "agents": n_agents, "episodes": n_ep, "agents_with_any_difference": changed,
                          "agents_with_a_different_trajectory": int(traj), "steps_per_agent": float(a["train_steps"]) / n_agents})
            print(cells[-1])
out["ucb_cells"] = cells
out["summary"] = {"cells": len(cells), "agents": n_agents * len(cells), "agents_with_any_difference": sum(c["agents_with_any_difference"] for c in cells),
                  "agents_with_a_different_trajectory": sum(c["agents_with_a_different_trajectory"] for c in cells)}
json.dump(out, open(os.path.join(ROOT, "profiles", "log_sensitivity.json"), "w"), indent=1)
print(json.dumps(out["summary"]))
