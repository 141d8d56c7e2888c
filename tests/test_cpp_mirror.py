"""The C++ host-side mirror (include/rlb.hpp): a client program written like code against the reference crate is compiled
with g++, linked with librlb.so and — on the GPU — its printed results compared bit for bit with the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

import parity as P
from oracle import oracle_py as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "mirror_main")


def build():
    libdir = os.path.join(ROOT, "rl-rust_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "mirror_main.cpp"), "-o", EXE, "-L", libdir, "-lrlb",
                           "-Wl,-rpath," + libdir])


def f64s(tokens):
    return np.array([struct.unpack("<d", struct.pack("<Q", int(t, 16)))[0] for t in tokens], np.float64)


def test_mirror_compiles_and_refuses_without_a_gpu(rlb):
    build()
    if rlb.abi.lib.rlb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    out = subprocess.check_output([EXE, "probe"], text=True)
    assert out.startswith("probe status=3") and "no CPU path" in out


@pytest.mark.gpu
def test_mirror_results_equal_oracle():
    build()
    out = subprocess.check_output([EXE], text=True)
    assert "FAILED" not in out and "no-throw" not in out
    assert "A eval_at==0 -> domain_error" in out and "A step after termination -> EnvNotReady" in out
    lines = {l.split()[0]: l.split()[1:] for l in out.splitlines() if l and l[0] in "ABCEF" and "." in l.split()[0]}
    n = 60
    cfg = O.make_config(O.ENV_TAXI, target=O.TARGET_QLEARNING, eps_decay=1.0 / (0.5 * n), seed=0xC0DE)
    sessions = [O.Session(cfg, i) for i in range(3)]
    rew, ln, err = f64s(lines["A.rewards"]).reshape(3, n), np.array(lines["A.lengths"], np.uint64).reshape(3, n), f64s(lines["A.errors"]).reshape(3, n)
    erew, eln = f64s(lines["A.eval_rewards"]).reshape(3, 10), np.array(lines["A.eval_lengths"], np.uint64).reshape(3, 10)
    step = lines["A.step"]
    for i, s in enumerate(sessions):
        r, l, t, _ = s.train(n, n // 10)
        assert np.array_equal(rew[i], r) and np.array_equal(ln[i], l) and P.bits_equal(err[i], t)
        r, l = s.evaluate(10)
        assert np.array_equal(erew[i], r) and np.array_equal(eln[i], l)
        o = s.env_reset()
        a = s.get_action(o)
        o2, r2, _ = s.env_step(a)
        assert int(step[i]) == o and int(step[4 + i]) == o2 and float(step[8 + i]) == r2
        s.agent_reset(); s.set_selector(1); s.set_target(2)
        r, l, t, _ = s.train(20, 2)
        assert np.array_equal(np.array(lines["A2.lengths"], np.uint64).reshape(3, 20)[i], l)
        assert P.bits_equal(f64s(lines["A2.errors"]).reshape(3, 20)[i], t)
        s.close()
    c = dict(env=1, agent=1, selector=0, policy=0, target=0, real=0)
    o = O.batch_train(P.oracle_config(c, P.hyper(20, seed=77)), 0, 40, 20, 2, n_threads=4)
    tot = next(l for l in out.splitlines() if l.startswith("B.totals")).split()
    assert int(tot[1]) == o["train_steps"] and int(tot[2]) == o["eval_steps"] and tot[3] == "store=3"
    assert np.array_equal(np.array(lines["B.lengths"], np.uint64).reshape(40, 20), o["len"])
    # C: Dyna-Q through InternalModelAgent / RandomModel, then the plain agent again once the borrow ended
    assert "C model unbound -> logic_error" in out
    cfg = O.make_config(O.ENV_CLIFF_WALKING, target=O.TARGET_QLEARNING, eps_decay=1.0 / (0.5 * 30), seed=0xD17A, planning_steps=10)
    lens, errs, after = (np.array(lines["C.lengths"], np.uint64).reshape(2, 30), f64s(lines["C.errors"]).reshape(2, 30),
                         np.array(lines["C.after"], np.uint64).reshape(2, 5))
    for i in range(2):
        s = O.Session(cfg, i)
        r, l, t, _ = s.train(30, 3)
        assert np.array_equal(lens[i], l) and P.bits_equal(errs[i], t)
        ms, ma, ms2, mr = s.model()
        assert int(lines["C.model"][i]) == len(ms)
        if i == 0:
            assert [int(x) for x in lines["C.model"][3:6]] == [ms[0], ma[0], ms2[0]] and float(lines["C.model"][6]) == mr[0]
            info = s.model_get_info()
            assert [int(x) for x in lines["C.info"][:3]] == list(info[:3]) and float(lines["C.info"][3]) == info[3]
        else:
            s.model_get_info()      # the batched call draws on every agent's stream
        s.set_planning(0)
        r, l, t, _ = s.train(5, 5)
        assert np.array_equal(after[i], l)
        s.close()
    # F: a caller-supplied FrozenLake map through the mirror's FrozenLakeEnv(rows, ..) constructor
    rows = ["SFFFFFH", "FFHFFFF", "FFFFHFS", "HFFFFFF", "FFFHFFG"]
    cf = dict(env=1, agent=0, selector=0, policy=0, target=1, real=1)
    of = O.batch_train(P.oracle_config(cf, P.hyper(30, seed=0xF1A6, max_steps=40, map_rows=rows)), 0, 4, 30, 10, n_threads=2)
    assert np.array_equal(np.array(lines["F.lengths"], np.uint64).reshape(4, 30), of["len"])
    assert P.bits_equal(f64s(lines["F.rewards"]).reshape(4, 30), of["ret"])
    # E: one agent -> training_error is the reference's per-step vector
    s = O.Session(O.make_config(O.ENV_TAXI, agent=O.AGENT_TRACES, target=O.TARGET_QLEARNING, eps_decay=1.0 / (0.5 * 25), seed=0x51), 0)
    r, l, t, _ = s.train(25, 5)
    assert np.array_equal(np.array(lines["E.lengths"], np.uint64), l)
    assert len(lines["E.errors"]) == int(l.sum()) and P.bits_equal(f64s(lines["E.errors"]), s.training_error())
    s.close()
    # D: Agent::example / Env::render through the C++ mirror equal the transcript built from the oracle (the same helper
    # the Python mirror's test uses) — four envs, one untrained episode each
    import importlib
    from test_gpu_example import oracle_transcript, SEED
    R = importlib.import_module("rl-rust_b200.render")
    rlb = importlib.import_module("rl-rust_b200")
    assert "D example with 2 agents -> logic_error" in out
    decay = 1.0 / 15.0
    m8 = rlb.FrozenLakeEnv.MAP_8X8

    def hands(n0, n1):
        return R.cards_from_words(rlb.abi.rng_words(SEED, 0, n0, n1 - n0))
    cases = {"taxi": (O.ENV_TAXI, {}, lambda pos, ready: R.render_taxi(pos), rlb.TaxiEnv.ACTIONS, None),
             "frozen_lake": (O.ENV_FROZEN_LAKE, dict(map_id=1, slippery=True), lambda pos, ready: R.render_frozen_lake(m8, pos),
                             rlb.FrozenLakeEnv.ACTIONS, None),
             "cliff_walking": (O.ENV_CLIFF_WALKING, {}, lambda pos, ready: R.render_cliff_walking(pos), rlb.CliffWalkingEnv.ACTIONS, None),
             "blackjack": (O.ENV_BLACKJACK, {}, lambda pos, ready, dealer, player: R.render_blackjack(ready, dealer, player),
                           rlb.BlackJackEnv.ACTIONS, hands)}
    for tag, (kind, extra, view, labels, hd) in cases.items():
        got = [l.split("|", 1)[1].replace("\\n", "\n") for l in out.splitlines() if l.startswith("D.%s|" % tag)]
        cfg = O.make_config(kind, target=O.TARGET_QLEARNING, eps_decay=decay, seed=SEED, **extra)
        want = oracle_transcript(cfg, kind, view, lambda a, labels=labels: labels[a], hd)
        assert got == want, tag

