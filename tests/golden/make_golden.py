#!/usr/bin/env python
"""Generates tests/golden/golden_v1.npz from the CPU oracle (oracle/).  The reference ships no fixtures and cannot be
run here (no Rust toolchain), so these vectors are the ORACLE's outputs, pinned so that (a) the oracle cannot drift
silently (tests/test_golden.py::test_oracle_reproduces_golden, CPU) and (b) the CUDA engine is also checked against
committed data rather than only a live oracle (test_engine_reproduces_golden, GPU).

    python tests/golden/make_golden.py          # rewrites golden_v1.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import parity as P            # noqa: E402
from oracle import oracle_py as O   # noqa: E402

# BASELINE.json config families C1..C4 in both arithmetic modes, plus two cells that exercise every other switch
CASES = {
    "c1_blackjack_q_eps_basic": dict(env=0, agent=0, selector=0, policy=0, target=1),
    "c2_frozenlake_sarsa_lambda_eps_basic": dict(env=1, agent=1, selector=0, policy=0, target=0),
    "c3_cliff_expsarsa_double_ucb": dict(env=2, agent=0, selector=1, policy=1, target=2),
    "c4_taxi_q_eps_basic": dict(env=3, agent=0, selector=0, policy=0, target=1),
    "x_taxi_qlambda_ucb_double": dict(env=3, agent=1, selector=1, policy=1, target=1),
    "x_blackjack_expsarsa_lambda_eps_double": dict(env=0, agent=1, selector=0, policy=1, target=2),
    # Dyna (agent/internal_model_agent.rs): the bin's cell (bin/cliffwalking_model.rs) and one with every switch flipped
    "d_cliff_dynaq10_eps_basic": dict(env=2, agent=0, selector=0, policy=0, target=1, planning=10),
    "d_taxi_dyna2_sarsa_lambda_ucb_double": dict(env=3, agent=1, selector=1, policy=1, target=0, planning=2),
}
N_AGENTS, N_EPISODES, EVAL_AT, FIRST_AGENT = 16, 30, 10, 1000


def run_case(c):
    h = P.hyper(N_EPISODES, planning_steps=c.get("planning", 0))
    return O.batch_train(P.oracle_config(c, h), FIRST_AGENT, N_AGENTS, N_EPISODES, EVAL_AT, n_threads=4)


def main():
    out = {}
    for name, base in CASES.items():
        for real in (0, 1):
            c = dict(base, real=real)
            o = run_case(c)
            key = "%s_%s" % (name, "f64" if real else "f32")
            out[key + "/len"] = o["len"].astype(np.uint32)
            out[key + "/ret"] = o["ret"]
            out[key + "/tdsum"] = o["tdsum"]
            out[key + "/q"] = o["q"][:4]                     # tables of the first 4 agents
            out[key + "/rng_n"] = o["state"]["rng_n"]
            out[key + "/ucb_t"] = o["state"]["ucb_t"]
            out[key + "/epsilon"] = o["state"]["epsilon"]
            out[key + "/steps"] = np.array([o["train_steps"], o["eval_steps"]], np.uint64)
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_v1.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
