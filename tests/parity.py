"""Shared helpers for the parity tests: run the same seeded configuration through the CPU
oracle (oracle/) and through the CUDA engine (C ABI, via rl-rust_b200), and compare bit for bit."""
import importlib
import itertools

import numpy as np

from oracle import oracle_py as O

import importlib as _il

_W = _il.import_module("rl-rust_b200.workloads")   # product-side helpers: the CLI defaults and engine construction
ENV_NAMES, TARGET_NAMES = _W.ENV_NAMES, _W.TARGET_NAMES
combo_id, hyper, make_engine = _W.combo_id, _W.hyper, _W.make_engine


def all_combos(envs=(0, 1, 2, 3), agents=(0, 1), selectors=(0, 1), policies=(0, 1), targets=(0, 1, 2), reals=(0, 1)):
    out = []
    for env, agent, sel, pol, tgt, real in itertools.product(envs, agents, selectors, policies, targets, reals):
        out.append(dict(env=env, agent=agent, selector=sel, policy=pol, target=tgt, real=real))
    return out


def oracle_config(c, h):
    return O.make_config(c["env"], map_id=h["map_id"], slippery=h["slippery"], max_steps=h["max_steps"],
                         policy=c["policy"], selector=c["selector"], target=c["target"], agent=c["agent"],
                         real=c["real"], decay_kind=h["decay_kind"], lr=h["lr"], gamma=h["gamma"],
                         lambda_=h["lambda_"], eps0=h["eps0"], eps_decay=h["eps_decay"], eps_final=h["eps_final"],
                         ucb_c=h["ucb_c"], default_q=h["default_q"], seed=h["seed"],
                         planning_steps=h.get("planning_steps", 0), map_rows=h.get("map_rows"))


def bits_equal(a, b):
    """Bit-exact float comparison; NaN == NaN (any payload)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.shape != b.shape:
        return False
    same = a.view(np.uint64) == b.view(np.uint64)
    both_nan = np.isnan(a) & np.isnan(b)
    return bool(np.all(same | both_nan))


def first_diff(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    bad = ~((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b)))
    idx = np.argwhere(bad)
    if len(idx) == 0:
        return None
    i = tuple(idx[0])
    return i, a[i], b[i], int(bad.sum())


def gpu_run(c, h, n_agents, n_episodes, eval_at, first_agent_id=0, traj_capacity=0, chunks=None, store_kind=0):
    """Train through the C ABI; returns dict of numpy arrays shaped like the oracle's batch output."""
    with make_engine(c, h, n_agents, first_agent_id, store_kind=store_kind) as eng:
        if chunks is None:
            res = eng.train(n_episodes, eval_at, sums=True, episodes=True, traj_capacity=traj_capacity)
            eps, sums = res["episodes"], res["sums"]
            train_steps, eval_steps = res["train_steps"], res["eval_steps"]
        else:
            parts, sparts = [], []
            train_steps = eval_steps = 0
            b = 0
            for e in chunks:
                r = eng.train(e, eval_at, ep_begin=b, sums=True, episodes=True)
                parts.append(r["episodes"]); sparts.append(r["sums"])
                train_steps += r["train_steps"]; eval_steps += r["eval_steps"]
                b = e
            eps, sums = np.concatenate(parts, 0), np.concatenate(sparts, 0)
            res = {}
        q, counts = eng.download_tables()
        st = eng.states()
        model = eng.download_model() if h.get("planning_steps", 0) else None
    out = dict(ret=eps["ret"].T.astype(np.float64), len=eps["length"].T.astype(np.uint64),
               tdsum=eps["td_sum"].T.astype(np.float64), tdabs=eps["td_abs_sum"].T.astype(np.float64),
               q=q.astype(np.float64), counts=counts.astype(np.uint64), state=st, sums=sums,
               train_steps=train_steps, eval_steps=eval_steps)
    if traj_capacity:
        out["traj"], out["traj_count"] = res["traj"], res["traj_count"]
    if model is not None:
        out["model_len"], out["model"] = model
    return out


def compare(g, o, c, check_counts=True):
    """Assert the engine's results equal the oracle's: integers and rewards exactly, floats bit for bit."""
    tag = combo_id(c)
    assert np.array_equal(g["len"], o["len"]), "%s: episode lengths differ" % tag
    assert bits_equal(g["ret"], o["ret"]), "%s: returns differ %r" % (tag, first_diff(g["ret"], o["ret"]))
    assert bits_equal(g["tdsum"], o["tdsum"]), "%s: td sums differ %r" % (tag, first_diff(g["tdsum"], o["tdsum"]))
    assert bits_equal(g["tdabs"], o["tdabs"]), "%s: |td| sums differ %r" % (tag, first_diff(g["tdabs"], o["tdabs"]))
    assert bits_equal(g["q"], o["q"]), "%s: Q tables differ %r" % (tag, first_diff(g["q"], o["q"]))
    assert np.array_equal(g["state"]["rng_n"], o["state"]["rng_n"]), "%s: RNG word counters differ" % tag
    assert np.array_equal(g["state"]["policy_flag"], o["state"]["policy_flag"]), "%s: Double flags differ" % tag
    if c["selector"] == 0:
        assert bits_equal(g["state"]["epsilon"], o["state"]["epsilon"]), "%s: epsilon differs" % tag
    else:
        assert np.array_equal(g["state"]["ucb_t"], o["state"]["ucb_t"]), "%s: UCB t differs" % tag
        if check_counts:
            assert np.array_equal(g["counts"], o["counts"]), "%s: UCB counts differ" % tag
    assert g["train_steps"] == o["train_steps"], "%s: train step totals differ" % tag
    assert g["eval_steps"] == o["eval_steps"], "%s: eval step totals differ" % tag
