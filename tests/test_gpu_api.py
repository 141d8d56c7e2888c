"""GPU tests of the C ABI beyond the parity matrix: step-level trajectories, the individual trait
methods, chunked runs, the CLI's train -> evaluate -> reset sequence, sharding invariance, error
behaviour, snapshots — and size-independent properties at BASELINE.json's full sizes."""
import numpy as np
import pytest

import parity as P
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu

TRAJ_CASES = [
    dict(env=0, agent=0, selector=0, policy=0, target=1, real=1),   # C1 family: Blackjack Q-learning
    dict(env=0, agent=1, selector=1, policy=1, target=2, real=1),
    dict(env=1, agent=1, selector=0, policy=0, target=0, real=0),   # C2 family
    dict(env=1, agent=1, selector=0, policy=0, target=0, real=1),
    dict(env=2, agent=0, selector=1, policy=1, target=2, real=1),   # C3 family (NaN-poisoned by the reference's semantics)
    dict(env=2, agent=0, selector=1, policy=1, target=1, real=0),
    dict(env=3, agent=0, selector=0, policy=0, target=1, real=0),   # C4 family
    dict(env=3, agent=1, selector=1, policy=0, target=2, real=1),
]


@pytest.mark.parametrize("c", TRAJ_CASES, ids=P.combo_id)
def test_step_trajectories_bit_exact(c):
    """Every env transition, chosen action, reward, termination flag and TD of every train / evaluate step."""
    n_agents, n_ep, eval_at = 6, 6, 3
    h = P.hyper(n_ep)
    g = P.gpu_run(c, h, n_agents, n_ep, eval_at, traj_capacity=60000, store_kind=2 if c["env"] in (1, 2) else 1)
    cfg = P.oracle_config(c, h)
    for i in range(n_agents):
        s = O.Session(cfg, i)
        s.record()
        s.train(n_ep, eval_at)
        tr = s.trajectory()
        s.close()
        n = int(g["traj_count"][i])
        assert n == len(tr) and n <= 60000
        gt = g["traj"][i, :n]
        for f in ("kind", "action", "terminated", "obs"):
            assert np.array_equal(gt[f], tr[f]), "%s agent %d field %s" % (P.combo_id(c), i, f)
        assert P.bits_equal(gt["reward"], tr["reward"])
        assert P.bits_equal(gt["td"], tr["td"]), P.first_diff(gt["td"], tr["td"])


@pytest.mark.parametrize("c", [TRAJ_CASES[2], TRAJ_CASES[6], TRAJ_CASES[4]], ids=P.combo_id)
def test_chunked_run_equals_single_call(c):
    """rlb_agent_train_range over [0,5) [5,6) [6,17) [17,20) == rlb_agent_train(20): eval injection uses the global episode index."""
    n_agents, n_ep, eval_at = 40, 20, 5
    h = P.hyper(n_ep)
    one = P.gpu_run(c, h, n_agents, n_ep, eval_at)
    parts = P.gpu_run(c, h, n_agents, n_ep, eval_at, chunks=[5, 6, 17, 20])
    o = O.batch_train(P.oracle_config(c, h), 0, n_agents, n_ep, eval_at, n_threads=4)
    P.compare(one, o, c)
    P.compare(parts, o, c)


@pytest.mark.parametrize("c", [TRAJ_CASES[0], TRAJ_CASES[3], TRAJ_CASES[5], TRAJ_CASES[7]], ids=P.combo_id)
def test_step_level_trait_methods(c):
    """Drive env.reset / agent.get_action / env.step / agent.update from the host, as Agent::train does
    (agent.rs:80-106), on the engine and on the oracle; every returned value must agree."""
    n_agents = 5
    h = P.hyper(10, max_steps=12)
    cfg = P.oracle_config(c, h)
    sessions = [O.Session(cfg, i) for i in range(n_agents)]
    with P.make_engine(c, h, n_agents) as eng:
        with pytest.raises(Exception) as ei:
            eng.env_step(np.zeros(n_agents, np.uint32))               # step before reset -> EnvNotReady (env.rs:16-17)
        assert ei.value.status == 1
        for episode in range(3):
            obs = eng.env_reset()
            assert list(obs) == [s.env_reset() for s in sessions]
            act = eng.get_action(obs)
            assert list(act) == [s.get_action(int(o)) for s, o in zip(sessions, obs)]
            # lock-step until every agent terminated; finished agents idle (their env would be NotReady)
            alive = np.ones(n_agents, bool)
            o_obs, o_act = obs.copy(), act.copy()
            for _ in range(40):
                if not alive.any():
                    break
                if not alive.all():
                    break                                              # batched step needs every env ready; stop this episode here
                obs2, rew, term = eng.env_step(o_act)
                ref = [s.env_step(int(a)) for s, a in zip(sessions, o_act)]
                assert list(obs2) == [r[0] for r in ref] and list(rew) == [r[1] for r in ref] and list(term) == [r[2] for r in ref]
                act2 = eng.get_action(obs2)
                assert list(act2) == [s.get_action(int(o)) for s, o in zip(sessions, obs2)]
                td = eng.update(o_obs, o_act, rew, term, obs2, act2)
                rtd = [s.update(int(a), int(b), float(r), bool(t), int(cc), int(d)) for s, a, b, r, t, cc, d in
                       zip(sessions, o_obs, o_act, rew, term, obs2, act2)]
                assert P.bits_equal(td.astype(np.float64), np.array(rtd))
                alive &= ~term
                o_obs, o_act = obs2, act2
        q, counts = eng.download_tables()
        st = eng.states()
    for i, s in enumerate(sessions):
        oq, oc, ost = s.export()
        assert P.bits_equal(q[i].astype(np.float64), oq)
        assert np.array_equal(counts[i].astype(np.uint64), oc)
        assert st["rng_n"][i] == ost.rng_n and st["policy_flag"][i] == ost.policy_flag
        s.close()


def test_cli_sequence_train_evaluate_reset():
    """bin/taxi.rs:158-203: for each selector, for each target fn: train -> evaluate -> agent.reset(); the env and the
    RNG stream carry on, the Double flag survives reset."""
    c = dict(env=3, agent=0, selector=0, policy=1, target=0, real=1)
    n_agents, n_ep = 12, 12
    h = P.hyper(n_ep)
    cfg = P.oracle_config(c, h)
    sessions = [O.Session(cfg, i) for i in range(n_agents)]
    with P.make_engine(c, h, n_agents) as eng:
        for sel in (0, 1):
            eng.set_selector(sel)
            for s in sessions:
                s.set_selector(sel)
            for tgt in (0, 1, 2):
                eng.set_target(tgt)
                r = eng.train(n_ep, max(1, n_ep // 10), sums=False, episodes=True)
                ev = eng.evaluate(n_ep, episodes=True)
                for i, s in enumerate(sessions):
                    s.set_target(tgt)
                    ret, ln, tds, tda = s.train(n_ep, max(1, n_ep // 10))
                    eret, eln = s.evaluate(n_ep)
                    assert np.array_equal(r["episodes"]["length"][:, i], ln)
                    assert P.bits_equal(r["episodes"]["td_sum"][:, i].astype(np.float64), tds)
                    assert np.array_equal(ev["episodes"]["length"][:, i], eln)
                    assert P.bits_equal(ev["episodes"]["ret"][:, i].astype(np.float64), eret)
                q, counts = eng.download_tables()
                st = eng.states()
                for i, s in enumerate(sessions):
                    oq, oc, ost = s.export()
                    assert P.bits_equal(q[i].astype(np.float64), oq)
                    assert st["rng_n"][i] == ost.rng_n and st["policy_flag"][i] == ost.policy_flag
                    if sel == 1:
                        assert np.array_equal(counts[i].astype(np.uint64), oc) and st["ucb_t"][i] == ost.ucb_t
                eng.agent_reset()
                for s in sessions:
                    s.agent_reset()
    for s in sessions:
        s.close()


def test_sharding_invariance():
    """Agents are keyed by GLOBAL id: one engine with 96 agents == three engines of 32 with first_agent_id 0/32/64."""
    c = dict(env=1, agent=1, selector=0, policy=0, target=0, real=0)
    h = P.hyper(10)
    whole = P.gpu_run(c, h, 96, 10, 5)
    for k in range(3):
        part = P.gpu_run(c, h, 32, 10, 5, first_agent_id=32 * k)
        sl = slice(32 * k, 32 * (k + 1))
        assert np.array_equal(part["len"], whole["len"][sl])
        assert P.bits_equal(part["q"], whole["q"][sl])
        assert np.array_equal(part["state"]["rng_n"], whole["state"]["rng_n"][sl])


def test_errors_and_snapshots(rlb):
    with pytest.raises(rlb.RlbError) as ei:
        rlb.Engine(3, n_agents=0)
    assert ei.value.status == 2
    with rlb.Engine(3, n_agents=4, real=1, policy=1, selector=1) as eng:
        with pytest.raises(rlb.RlbError) as ei:
            eng.train(5, 0)                                            # eval_at == 0: the reference divides by it (agent.rs:107)
        assert ei.value.status == 2
        q = np.random.default_rng(0).normal(size=(4, 2, 500, 6))
        cnt = np.random.default_rng(1).integers(0, 9, size=(4, 500, 6)).astype(np.uint32)
        eng.upload_tables(q, cnt)
        q2, c2 = eng.download_tables()
        assert np.array_equal(q, q2) and np.array_equal(cnt, c2)
        st = eng.states()
        assert np.all(st["epsilon"] == 1.0) and np.all(st["ucb_t"] == 1) and np.all(st["policy_flag"] == 1) and np.all(st["env_ready"] == 0)
        st["epsilon"] = 0.25; st["rng_n"] = 1234; st["policy_flag"] = 0
        eng.set_states(st)
        assert np.array_equal(eng.states()[["epsilon", "rng_n", "policy_flag"]], st[["epsilon", "rng_n", "policy_flag"]])
        # Policy::predict = (alpha+beta)/2, get_values = alpha if flag else beta (double_tabular_policy.rs:31-48)
        obs = np.array([3, 77, 499, 0], np.uint32)
        pred, vals = eng.policy_predict(obs), eng.policy_get_values(obs)
        for i, o in enumerate(obs):
            assert np.array_equal(pred[i], (q[i, 0, o] + q[i, 1, o]) / 2.0) and np.array_equal(vals[i], q[i, 1, o])
        eng.policy_after_update()
        assert np.array_equal(eng.policy_get_values(obs)[2], q[2, 0, 499])
        # Policy::update writes the other table: flag is now true -> beta += lr * td
        eng.policy_update(obs, np.array([1, 2, 3, 4], np.uint32), obs, np.array([1.0, -2.0, 0.5, 4.0]))
        q3, _ = eng.download_tables()
        assert q3[0, 1, 3, 1] == q[0, 1, 3, 1] + 0.05 * 1.0 and q3[3, 1, 0, 4] == q[3, 1, 0, 4] + 0.05 * 4.0
        assert np.array_equal(q3[:, 0], q[:, 0])
        eng.policy_reset()
        assert np.all(eng.download_tables()[0] == 0.0)


def test_selector_methods(rlb):
    with rlb.Engine(2, n_agents=3, real=1, selector=0, initial_epsilon=0.2, epsilon_decay=0.05) as eng:
        vals = np.array([[0.0, 2.0, 2.0, 1.0], [5.0, 1.0, 1.0, 1.0], [0.0, 0.0, 0.0, 0.0]])
        obs = np.zeros(3, np.uint32)
        pr = eng.selector_get_exploration_probs(obs, vals)            # uniform_epsilon_greed.rs:72-76 (sums to 1 - eps/A ... not 1)
        assert np.array_equal(pr[0], [0.05, 0.8, 0.05, 0.05]) and np.array_equal(pr[1], [0.8, 0.05, 0.05, 0.05])
        eng.selector_update()                                          # eps <- eps - 0.05
        assert np.all(eng.states()["epsilon"] == 0.2 - 0.05)
        for _ in range(10):
            eng.selector_update()                                      # would go below final_epsilon=0 -> keeps the old value (Q8)
        eps = eng.states()["epsilon"][0]
        assert 0.0 <= eps < 0.05
        eng.selector_reset()
        assert np.all(eng.states()["epsilon"] == 0.2)
    with rlb.Engine(2, n_agents=2, real=1, selector=1, confidence_level=0.5) as eng:
        vals = np.array([[0.0, 1.0, 3.0, 2.0], [1.0, 1.0, 1.0, 1.0]])
        obs = np.array([7, 7], np.uint32)
        assert list(eng.selector_get_action(obs, vals)) == [2, 0]     # t = 1: ln(1) = 0 -> pure argmax (Q9)
        a = eng.selector_get_action(obs, vals)                        # t = 2: unvisited actions get a huge bonus, index order
        assert list(a) == [0, 1]
        _, counts = eng.download_tables()
        assert list(counts[0, 7]) == [1, 0, 1, 0] and list(counts[1, 7]) == [1, 1, 0, 0]
        assert np.all(eng.states()["ucb_t"] == 3)


# ------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties + a slice pinned to the oracle
# ------------------------------------------------------------------------------------------------
FULL = [
    ("c2", dict(env=1, agent=1, selector=0, policy=0, target=0, real=0), 1 << 20, 4),
    ("c3", dict(env=2, agent=0, selector=1, policy=1, target=2, real=0), 1 << 22, 2),
    ("c4", dict(env=3, agent=0, selector=0, policy=0, target=1, real=0), 1 << 21, 4),
]


@pytest.mark.parametrize("name,c,n_agents,n_ep", FULL, ids=[f[0] for f in FULL])
def test_full_size_properties(name, c, n_agents, n_ep):
    h = P.hyper(1000)
    eval_at = 2
    with P.make_engine(c, h, n_agents) as eng:
        r = eng.train(n_ep, eval_at, sums=True, episodes=True)
        ep = r["episodes"]
        st = eng.states()
        # a slice of the batch: tables of the first and last 64 agents
    length = ep["length"]
    assert length.min() >= 1 and length.max() <= h["max_steps"] + 1                      # Q2: truncation pseudo-step
    assert int(length.sum()) == r["train_steps"]
    assert np.array_equal(r["sums"][:, 0], length.sum(1).astype(np.float64))
    assert np.array_equal(r["sums"][:, 1], ep["ret"].astype(np.float64).sum(1))
    if c["env"] == 1:
        assert set(np.unique(ep["ret"])) <= {0.0, 1.0}
    if c["env"] == 3:
        assert ep["ret"].max() <= 20.0 and np.all(ep["ret"] == np.round(ep["ret"]))
    assert r["eval_episodes"] == n_agents * 100 * len([e for e in range(n_ep) if e % eval_at == 0])
    # per-agent results do not depend on the batch they ran in: re-run two 64-agent windows alone and against the oracle
    for first in (0, n_agents - 64):
        part = P.gpu_run(c, h, 64, n_ep, eval_at, first_agent_id=first)
        assert np.array_equal(part["len"], length[:, first:first + 64].T)
        assert P.bits_equal(part["tdsum"], ep["td_sum"][:, first:first + 64].T.astype(np.float64))
        assert np.array_equal(part["state"]["rng_n"], st["rng_n"][first:first + 64])
        o = O.batch_train(P.oracle_config(c, h), first, 64, n_ep, eval_at, n_threads=8)
        P.compare(part, o, c)


# ------------------------------------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("env", [1, 2, 3])
@pytest.mark.parametrize("max_steps", [0, 1, 2])
def test_truncation_edges(env, max_steps):
    """max_steps = 0: every episode is the truncation pseudo-step alone (obs 0, terminated; Q2).  Trace agent, both
    arithmetic modes, every table store the env supports."""
    n_agents, n_ep = 33, 6            # one lane past a warp
    for real in (0, 1):
        c = dict(env=env, agent=1, selector=0, policy=0, target=1, real=real)
        h = P.hyper(n_ep, max_steps=max_steps)
        o = O.batch_train(P.oracle_config(c, h), 0, n_agents, n_ep, 2, n_threads=4)
        assert o["len"].max() <= max_steps + 1
        for store in ((1, 2, 3, 4) if env in (1, 2) else (1, 4)):
            g = P.gpu_run(c, h, n_agents, n_ep, 2, store_kind=store)
            P.compare(g, o, c)


def test_zero_episodes_and_single_agent(rlb):
    with rlb.Engine(1, n_agents=1, agent=1, slippery=True) as eng:
        r = eng.train(0, 5, sums=True, episodes=True)
        assert r["train_steps"] == 0 and r["eval_steps"] == 0 and r["sums"].shape == (0, 4) and r["episodes"].shape == (0, 1)
        assert eng.states()["rng_n"][0] == 0                      # nothing drawn
        ev = eng.evaluate(0)
        assert ev["steps"] == 0
        r = eng.train(3, 2, sums=True, episodes=True)             # evaluate(100) after episodes 0 and 2
        assert r["eval_episodes"] == 200 and r["train_steps"] == int(r["episodes"]["length"].sum())


def test_blackjack_heavy_rng_paths():
    """Blackjack draws u32 cards with rejection, so the stream goes odd and straddles the 8-word window: many agents,
    both selectors, to hit the mid-step refill paths."""
    for sel in (0, 1):
        c = dict(env=0, agent=0, selector=sel, policy=0, target=1, real=1)
        h = P.hyper(60)
        o = O.batch_train(P.oracle_config(c, h), 5000, 512, 60, 6, n_threads=8)
        g = P.gpu_run(c, h, 512, 60, 6, first_agent_id=5000)
        P.compare(g, o, c)


@pytest.mark.parametrize("name,c,n_ep,chunk", [("c2", dict(env=1, agent=1, selector=0, policy=0, target=0, real=0), 1000, 100),
                                                ("c4", dict(env=3, agent=0, selector=0, policy=0, target=1, real=0), 1000, 100)])
def test_bench_workload_full_length_sample(name, c, n_ep, chunk):
    """Exactly what bench.py runs — the whole 1000-episode run (epsilon 1 -> 0, then greedy), driven in 100-episode
    chunks with eval_at = n/10 — on 4096 agents; the first and last 24 agents are replayed by the oracle."""
    n_agents = 4096
    h = P.hyper(n_ep, slippery=True)
    with P.make_engine(c, h, n_agents) as eng:
        lens, tds = [], []
        for b in range(0, n_ep, chunk):
            r = eng.train(b + chunk, n_ep // 10, ep_begin=b, sums=False, episodes=True)
            lens.append(r["episodes"]["length"]); tds.append(r["episodes"]["td_sum"])
        length, tdsum = np.concatenate(lens, 0), np.concatenate(tds, 0)
        st = eng.states()
        q, _ = eng.download_tables(counts=False)
    for first in (0, n_agents - 24):
        o = O.batch_train(P.oracle_config(c, h), first, 24, n_ep, n_ep // 10, n_threads=8)
        sl = slice(first, first + 24)
        assert np.array_equal(length[:, sl].T, o["len"])
        assert P.bits_equal(tdsum[:, sl].T.astype(np.float64), o["tdsum"])
        assert P.bits_equal(q[sl].astype(np.float64), o["q"])
        assert np.array_equal(st["rng_n"][sl], o["state"]["rng_n"]) and P.bits_equal(st["epsilon"][sl], o["state"]["epsilon"])
    assert np.all(st["epsilon"] < 0.01)   # the schedule ran out: the last half of the run was (almost) greedy
