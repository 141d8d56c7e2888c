"""Committed golden vectors (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py from the CPU oracle).
CPU: the oracle still reproduces them.  GPU: the CUDA engine reproduces them through the C ABI, with no oracle in
the loop."""
import os
import sys

import numpy as np
import pytest

import parity as P

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as G   # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))
KEYS = sorted({k.split("/")[0] for k in GOLD.files})


def case_of(key):
    name, real = key.rsplit("_", 1)
    return dict(G.CASES[name], real=1 if real == "f64" else 0)


def check(key, res):
    assert np.array_equal(res["len"].astype(np.uint32), GOLD[key + "/len"])
    assert P.bits_equal(res["ret"], GOLD[key + "/ret"])
    assert P.bits_equal(res["tdsum"], GOLD[key + "/tdsum"]), P.first_diff(res["tdsum"], GOLD[key + "/tdsum"])
    assert P.bits_equal(res["q"][:4], GOLD[key + "/q"])
    assert np.array_equal(res["state"]["rng_n"], GOLD[key + "/rng_n"])
    c = case_of(key)
    if c["selector"] == 1:
        assert np.array_equal(res["state"]["ucb_t"], GOLD[key + "/ucb_t"])
    else:
        assert P.bits_equal(res["state"]["epsilon"], GOLD[key + "/epsilon"])
    assert [res["train_steps"], res["eval_steps"]] == list(GOLD[key + "/steps"])


@pytest.mark.parametrize("key", KEYS)
def test_oracle_reproduces_golden(key):
    check(key, G.run_case(case_of(key)))


@pytest.mark.gpu
@pytest.mark.parametrize("key", KEYS)
def test_engine_reproduces_golden(key):
    c = case_of(key)
    res = P.gpu_run(c, P.hyper(G.N_EPISODES, planning_steps=c.get("planning", 0)), G.N_AGENTS, G.N_EPISODES, G.EVAL_AT, first_agent_id=G.FIRST_AGENT)
    check(key, res)
