"""Dyna parity on the GPU: InternalModelAgent + RandomModel (agent/internal_model_agent.rs, model/random_model.rs) through
the C ABI against the CPU oracle on the same Philox streams — the fused training path, the step-level Model / Agent
methods, wrapping an agent that has already trained, chunked runs and model snapshots."""
import numpy as np
import pytest

import parity as P
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu

N_AGENTS, N_EPISODES, EVAL_AT = 40, 20, 7

# the bin's cell (bin/cliffwalking_model.rs: CliffWalking, one-step Q-learning, eps-greedy, Basic, 10 planning steps)
CLI_CELL = dict(env=2, agent=0, selector=0, policy=0, target=1)
CELLS = [
    (CLI_CELL, 10),
    (dict(env=2, agent=0, selector=1, policy=1, target=0), 3),     # UCB counts / t advance inside planning, Double flag flips
    (dict(env=2, agent=1, selector=0, policy=0, target=1), 2),     # planning sweeps leave traces behind across episodes
    (dict(env=1, agent=0, selector=0, policy=1, target=2), 5),     # stochastic env: first-seen next_obs is kept
    (dict(env=1, agent=1, selector=1, policy=0, target=0), 2),
    (dict(env=3, agent=0, selector=0, policy=0, target=1), 4),     # 6 actions: rejections in both samplers
    (dict(env=3, agent=1, selector=0, policy=1, target=2), 1),
    (dict(env=0, agent=0, selector=0, policy=0, target=1), 3),     # Blackjack ids, odd word alignment
    (dict(env=0, agent=1, selector=1, policy=1, target=0), 2),
]
CASES = [dict(c, real=r, planning=k) for c, k in CELLS for r in (0, 1)]


def case_id(c):
    return "%s-plan%d" % (P.combo_id(c), c["planning"])


def oracle_models(cfg, n_agents, n_episodes, eval_at):
    out = []
    for i in range(n_agents):
        s = O.Session(cfg, i)
        s.train(n_episodes, eval_at)
        out.append(s.model())
        s.close()
    return out


def assert_models_equal(g_len, g_ent, models, tag):
    for i, (s, a, s2, r) in enumerate(models):
        n = len(s)
        assert int(g_len[i]) == n, "%s: agent %d model size %d != %d" % (tag, i, g_len[i], n)
        e = g_ent[i, :n]
        assert np.array_equal(e["obs"], s) and np.array_equal(e["action"], a), "%s: agent %d model keys differ" % (tag, i)
        assert np.array_equal(e["next_obs"], s2), "%s: agent %d model next_obs differ" % (tag, i)
        assert np.array_equal(e["reward"].astype(np.float64), r), "%s: agent %d model rewards differ" % (tag, i)


@pytest.mark.parametrize("c", CASES, ids=case_id)
def test_dyna_training_matches_oracle(c):
    h = P.hyper(N_EPISODES, planning_steps=c["planning"])
    cfg = P.oracle_config(c, h)
    o = O.batch_train(cfg, 0, N_AGENTS, N_EPISODES, EVAL_AT, n_threads=8)
    g = P.gpu_run(c, h, N_AGENTS, N_EPISODES, EVAL_AT)
    P.compare(g, o, c)
    assert_models_equal(g["model_len"], g["model"], oracle_models(cfg, 6, N_EPISODES, EVAL_AT), case_id(c))


def test_dyna_learns_faster_than_plain_q_learning():
    """The bin's comparison (bin/cliffwalking_model.rs:158-160): same cell with and without the model."""
    c = dict(CLI_CELL, real=0)
    n_ep = 150
    plain = P.gpu_run(c, P.hyper(n_ep), 256, n_ep, 50)
    dyna = P.gpu_run(c, P.hyper(n_ep, planning_steps=10), 256, n_ep, 50)
    assert dyna["len"][:, -20:].mean() < 0.6 * plain["len"][:, -20:].mean()
    assert dyna["ret"][:, -20:].mean() > plain["ret"][:, -20:].mean()


def test_chunked_dyna_run_equals_single_call():
    c = dict(env=2, agent=1, selector=0, policy=1, target=1, real=1)
    h = P.hyper(18, planning_steps=3)
    one = P.gpu_run(c, h, 33, 18, 5)
    parts = P.gpu_run(c, h, 33, 18, 5, chunks=[4, 11, 18])
    P.compare(parts, one, c)
    assert np.array_equal(parts["model_len"], one["model_len"]) and np.array_equal(parts["model"], one["model"])


@pytest.mark.parametrize("c", [dict(CLI_CELL, real=1), dict(env=3, agent=1, selector=1, policy=1, target=2, real=0),
                               dict(env=1, agent=1, selector=0, policy=0, target=0, real=0)], ids=P.combo_id)
def test_wrap_a_trained_agent_then_unwrap(c):
    """InternalModelAgent::new borrows the agent as it is (tables, epsilon, counts kept); dropping the wrapper leaves it
    as planning left it."""
    h = P.hyper(30)
    n = 9
    sess = [O.Session(P.oracle_config(c, h), i) for i in range(n)]
    with P.make_engine(c, h, n) as eng:
        def both(ep_begin, ep_end):
            r = eng.train(ep_end, 6, ep_begin=ep_begin, episodes=True)
            for i, s in enumerate(sess):
                ret, ln, tds, tda = s.train(ep_end, 6, ep_begin=ep_begin)
                assert np.array_equal(r["episodes"]["length"][:, i].astype(np.uint64), ln)
                assert P.bits_equal(r["episodes"]["ret"][:, i], ret)
                assert P.bits_equal(r["episodes"]["td_sum"][:, i], tds)
        both(0, 10)
        eng.set_model(4)
        for s in sess:
            s.set_planning(4)
        both(10, 20)
        ln, ent = eng.download_model()
        assert_models_equal(ln, ent, [s.model() for s in sess], "wrapped")
        eng.set_model(0)
        for s in sess:
            s.set_planning(0)
        both(20, 30)
        q, counts = eng.download_tables()
        st = eng.states()
        for i, s in enumerate(sess):
            oq, oc, ost = s.export()
            assert P.bits_equal(q[i], oq), "agent %d: %r" % (i, P.first_diff(q[i], oq))
            assert int(st["rng_n"][i]) == ost.rng_n


def test_step_level_model_and_agent_methods(rlb):
    """Model::{add_info,get_info,reset} and InternalModelAgent::{update,reset} one call at a time."""
    c = dict(env=2, agent=0, selector=0, policy=1, target=0, real=1)
    h = P.hyper(50, planning_steps=3)
    n = 7
    sess = [O.Session(P.oracle_config(c, h), i) for i in range(n)]
    rng = np.random.default_rng(5)
    with P.make_engine(c, h, n) as eng:
        with pytest.raises(rlb.RlbError) as ei:   # gen_range over an empty range panics in the reference
            eng.model_get_info()
        assert ei.value.status == 2
        assert np.array_equal(eng.states()["rng_n"], np.zeros(n, np.uint64))
        for step in range(60):
            s = rng.integers(0, 37, n).astype(np.uint32)
            a = rng.integers(0, 4, n).astype(np.uint32)
            s2 = rng.integers(0, 48, n).astype(np.uint32)
            r = rng.choice([-1.0, -100.0], n)
            if step % 3 == 0:
                eng.model_add_info(s, a, r, s2)
                for i, o in enumerate(sess):
                    o.model_add_info(int(s[i]), int(a[i]), float(r[i]), int(s2[i]))
                gs, ga, gs2, gr = eng.model_get_info()
                for i, o in enumerate(sess):
                    assert (int(gs[i]), int(ga[i]), int(gs2[i]), float(gr[i])) == o.model_get_info()
            else:
                term = (rng.random(n) < 0.2).astype(np.uint8)
                a2 = eng.get_action(s2)
                td = eng.update(s, a, r, term, s2, a2)
                for i, o in enumerate(sess):
                    assert int(a2[i]) == o.get_action(int(s2[i]))
                    otd = o.update(int(s[i]), int(a[i]), float(r[i]), bool(term[i]), int(s2[i]), int(a2[i]))
                    assert P.bits_equal(td[i], otd)
        ln, ent = eng.download_model()
        assert_models_equal(ln, ent, [o.model() for o in sess], "step-level")
        q, _ = eng.download_tables()
        st = eng.states()
        for i, o in enumerate(sess):
            oq, _, ost = o.export()
            assert P.bits_equal(q[i], oq) and int(st["rng_n"][i]) == ost.rng_n and int(st["policy_flag"][i]) == ost.policy_flag
        # snapshot round trip of the model, then Agent::reset empties it (internal_model_agent.rs:81-84)
        eng.model_reset()
        assert not eng.download_model()[0].any()
        eng.upload_model(ln, ent)
        ln2, ent2 = eng.download_model()
        assert np.array_equal(ln, ln2) and np.array_equal(ent, ent2)
        s = np.full(n, 3, np.uint32)
        eng.model_add_info(s, s, np.zeros(n), s)     # a key every model may or may not hold: membership survived the upload
        for o in sess:
            o.model_add_info(3, 3, 0.0, 3)
        assert_models_equal(*eng.download_model(), [o.model() for o in sess], "after upload")
        eng.agent_reset()
        for o in sess:
            o.agent_reset()
        assert not eng.download_model()[0].any() and all(len(o.model()[0]) == 0 for o in sess)
        assert not eng.download_tables()[0].any()


def test_model_needs_the_hbm_store(rlb):
    c = dict(env=1, agent=1, selector=0, policy=0, target=0, real=0)
    with pytest.raises(rlb.RlbError) as ei:
        P.make_engine(c, P.hyper(10, planning_steps=2), 8, store_kind=3)
    assert ei.value.status == 5
    with P.make_engine(c, P.hyper(10), 8, store_kind=3) as eng:
        with pytest.raises(rlb.RlbError) as ei:
            eng.set_model(2)
        assert ei.value.status == 5
    with P.make_engine(c, P.hyper(10), 8) as eng:     # auto: hybrid without a model, HBM with one
        assert eng.store_kind() == 3
        eng.set_model(2)
        assert eng.store_kind() == 1
        eng.set_model(0)
        assert eng.store_kind() == 3
