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


# ---- round-2 features (tests/golden/golden_v2.npz, made by tests/golden/make_golden_v2.py): caller-supplied FrozenLake
# maps and the per-step training_error vector
import make_golden_v2 as G2   # noqa: E402

GOLD2 = np.load(os.path.join(HERE, "golden", "golden_v2.npz"))
KEYS2 = sorted({k.split("/")[0] for k in GOLD2.files})


def check2(key, res, tds):
    assert np.array_equal(res["len"].astype(np.uint32), GOLD2[key + "/len"])
    assert P.bits_equal(res["ret"], GOLD2[key + "/ret"])
    assert P.bits_equal(res["q"][:2], GOLD2[key + "/q"])
    assert np.array_equal(res["state"]["rng_n"], GOLD2[key + "/rng_n"])
    for i, t in enumerate(tds):
        want = GOLD2[key + "/td%d" % i]
        assert len(t) == len(want) == int(res["len"][i].sum()) and P.bits_equal(np.asarray(t, np.float64), want)


@pytest.mark.parametrize("key", KEYS2)
def test_oracle_reproduces_golden_v2(key):
    name, real = key.rsplit("_", 1)
    o = G2.run_case(G2.CASES[name], 1 if real == "f64" else 0)
    check2(key, o, o["td"])


@pytest.mark.gpu
@pytest.mark.parametrize("key", KEYS2)
def test_engine_reproduces_golden_v2(key):
    name, real = key.rsplit("_", 1)
    case = G2.CASES[name]
    c = dict(case["c"], real=1 if real == "f64" else 0)
    h = G2.hyper_of(case)
    res = P.gpu_run(c, h, G2.N_AGENTS, G2.N_EPISODES, G2.EVAL_AT, first_agent_id=G2.FIRST_AGENT)
    with P.make_engine(c, h, G2.TD_AGENTS, G2.FIRST_AGENT) as eng:
        r = eng.train(G2.N_EPISODES, G2.EVAL_AT, td_capacity=G2.N_EPISODES * 41)
    tds = [r["td_steps"][i, :int(r["td_count"][i])] for i in range(G2.TD_AGENTS)]
    check2(key, res, tds)


# ---- `Agent::example` transcripts (tests/golden/example_v1.json, made by tests/golden/make_example_golden.py)
import json   # noqa: E402

import make_example_golden as GE   # noqa: E402

EXAMPLES = json.load(open(os.path.join(HERE, "golden", "example_v1.json")))


def test_oracle_reproduces_golden_example_transcripts():
    got = GE.transcripts()
    assert sorted(got) == sorted(EXAMPLES)
    for name in EXAMPLES:
        assert got[name] == EXAMPLES[name], name
    # what the four episodes must show whatever the stream: Taxi's random walk is truncated after max_steps + 1 steps
    # WITHOUT moving the taxi (taxi.rs:148-151), CliffWalking's fall costs -100 and ends the episode (cliff_walking.rs:26-28)
    assert EXAMPLES["taxi"][-1] == "terminated with 101 steps" and EXAMPLES["taxi"][-4] == "step reward 0.0"
    assert EXAMPLES["taxi"][-3] == EXAMPLES["taxi"][-6]            # the view after the truncated step == the view before it
    assert EXAMPLES["cliff_walking"][-2] == "episode reward -101.0"


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(EXAMPLES))
def test_engine_reproduces_golden_example_transcripts(rlb, name):
    """The mirror's example() on the GPU against the committed transcript, no oracle in the loop."""
    sel = rlb.UniformEpsilonGreed(1.0, ("sub", GE.DECAY), 0.0)
    agent = rlb.OneStepAgent(rlb.TabularPolicy(0.05, 0.0), 0.95, sel, rlb.qlearning, n_agents=1, seed=GE.SEED, real="f64")
    env = {"taxi": lambda: rlb.TaxiEnv(100), "frozen_lake": lambda: rlb.FrozenLakeEnv(rlb.FrozenLakeEnv.MAP_8X8, True, 100),
           "cliff_walking": lambda: rlb.CliffWalkingEnv(100), "blackjack": rlb.BlackJackEnv}[name]()
    assert agent.example(env, out=lambda _: None) == EXAMPLES[name]
