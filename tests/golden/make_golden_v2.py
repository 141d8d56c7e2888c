#!/usr/bin/env python
"""Generates tests/golden/golden_v2.npz from the CPU oracle: the round-2 features — FrozenLakeEnv::new on caller-supplied
rows (frozen_lake.rs:48: several start cells, non-square, a start that is not cell 0) and the per-step `training_error`
vector (agent.rs:98,117).  Same purpose as golden_v1: the oracle cannot drift silently (CPU test) and the engine is held to
committed data with no oracle in the loop (GPU test).

    python tests/golden/make_golden_v2.py       # rewrites golden_v2.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import parity as P            # noqa: E402
from oracle import oracle_py as O   # noqa: E402

MAPS = {
    "two_starts_5x7": ["SFFFFFH", "FFHFFFF", "FFFFHFS", "HFFFFFF", "FFFHFFG"],
    "start_not_first_4x4": ["FFFH", "FSFF", "HFFF", "FFFG"],
}
CASES = {
    "m_two_starts_q_onestep": dict(c=dict(env=1, agent=0, selector=0, policy=0, target=1), map="two_starts_5x7", slippery=True),
    "m_two_starts_sarsa_lambda_double": dict(c=dict(env=1, agent=1, selector=0, policy=1, target=0), map="two_starts_5x7", slippery=True),
    "m_start_not_first_expsarsa_ucb": dict(c=dict(env=1, agent=0, selector=1, policy=0, target=2), map="start_not_first_4x4", slippery=False),
    "t_taxi_q_training_error": dict(c=dict(env=3, agent=0, selector=0, policy=0, target=1), map=None, slippery=True),
    "t_cliff_sarsa_lambda_training_error": dict(c=dict(env=2, agent=1, selector=0, policy=0, target=0), map=None, slippery=True),
}
N_AGENTS, N_EPISODES, EVAL_AT, FIRST_AGENT, TD_AGENTS = 8, 24, 8, 500, 2


def hyper_of(case):
    return P.hyper(N_EPISODES, slippery=case["slippery"], max_steps=40, map_rows=MAPS[case["map"]] if case["map"] else None)


def run_case(case, real):
    c = dict(case["c"], real=real)
    h = hyper_of(case)
    o = O.batch_train(P.oracle_config(c, h), FIRST_AGENT, N_AGENTS, N_EPISODES, EVAL_AT, n_threads=4)
    tds = []
    for i in range(TD_AGENTS):   # the per-step vector of the first agents
        s = O.Session(P.oracle_config(c, h), FIRST_AGENT + i)
        s.train(N_EPISODES, EVAL_AT)
        tds.append(s.training_error())
        s.close()
    o["td"] = tds
    return o


def main():
    out = {}
    for name, case in CASES.items():
        for real in (0, 1):
            o = run_case(case, real)
            key = "%s_%s" % (name, "f64" if real else "f32")
            out[key + "/len"] = o["len"].astype(np.uint32)
            out[key + "/ret"] = o["ret"]
            out[key + "/q"] = o["q"][:2]
            out[key + "/rng_n"] = o["state"]["rng_n"]
            for i, t in enumerate(o["td"]):
                out[key + "/td%d" % i] = t
    np.savez_compressed(os.path.join(HERE, "golden_v2.npz"), **out)
    print("wrote golden_v2.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
