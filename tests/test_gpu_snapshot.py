"""Snapshot / resume (SURVEY.md §5 "checkpoint / resume", next row N3): tables + per-agent scalars are the complete
resumable state at an episode boundary."""
import os
import tempfile

import numpy as np
import pytest

import parity as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("c", [dict(env=3, agent=0, selector=1, policy=1, target=2, real=1),
                               dict(env=1, agent=1, selector=0, policy=0, target=0, real=0),
                               dict(env=0, agent=1, selector=0, policy=1, target=1, real=0)], ids=P.combo_id)
def test_resume_from_snapshot_equals_uninterrupted_run(c, rlb):
    n_agents, n_ep, eval_at = 50, 16, 4
    h = P.hyper(n_ep)
    whole = P.gpu_run(c, h, n_agents, n_ep, eval_at)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "snap.npz")
        with P.make_engine(c, h, n_agents) as eng:
            eng.train(7, eval_at)
            rlb.save_snapshot(eng, path)
        with P.make_engine(c, h, n_agents) as eng2:           # a new engine (fresh tables, word 0 of every stream)
            rlb.load_snapshot(eng2, path)
            r = eng2.train(n_ep, eval_at, ep_begin=7, sums=False, episodes=True)
            q, counts = eng2.download_tables()
            st = eng2.states()
    assert np.array_equal(r["episodes"]["length"].T, whole["len"][:, 7:])
    assert P.bits_equal(r["episodes"]["td_sum"].T.astype(np.float64), whole["tdsum"][:, 7:])
    assert P.bits_equal(q.astype(np.float64), whole["q"]) and np.array_equal(counts.astype(np.uint64), whole["counts"])
    assert np.array_equal(st["rng_n"], whole["state"]["rng_n"]) and np.array_equal(st["ucb_t"], whole["state"]["ucb_t"])
    assert P.bits_equal(st["epsilon"], whole["state"]["epsilon"])


def test_resume_a_dyna_agent_from_snapshot(rlb):
    """The Dyna model (model/random_model.rs) is part of the resumable state; the snapshot re-attaches it."""
    c = dict(env=3, agent=0, selector=0, policy=1, target=1, real=1)
    n_agents, n_ep, eval_at = 21, 12, 5
    h = P.hyper(n_ep, planning_steps=3)
    whole = P.gpu_run(c, h, n_agents, n_ep, eval_at)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "snap.npz")
        with P.make_engine(c, h, n_agents) as eng:
            eng.train(5, eval_at)
            rlb.save_snapshot(eng, path)
        with P.make_engine(c, P.hyper(n_ep), n_agents) as eng2:     # created without a model
            rlb.load_snapshot(eng2, path)
            r = eng2.train(n_ep, eval_at, ep_begin=5, sums=False, episodes=True)
            q, _ = eng2.download_tables()
            st = eng2.states()
            ln, ent = eng2.download_model()
        with P.make_engine(dict(c, agent=1), h, 4) as eng3:
            with pytest.raises(NotImplementedError):
                rlb.save_snapshot(eng3, path)
    assert np.array_equal(r["episodes"]["length"].T, whole["len"][:, 5:])
    assert P.bits_equal(q.astype(np.float64), whole["q"]) and np.array_equal(st["rng_n"], whole["state"]["rng_n"])
    assert np.array_equal(ln, whole["model_len"]) and np.array_equal(ent, whole["model"])
