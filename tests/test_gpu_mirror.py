"""GPU test of the host-side mirror of the reference's trait surface (rl-rust_b200/api.py): the calls a user of the
reference would make — constructors, Agent.train / evaluate / set_* / reset, Env.reset / step — against the oracle."""
import numpy as np
import pytest

import parity as P
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu


def test_cli_like_run_through_the_mirror(rlb):
    """src/bin/taxi.rs:86-203 in miniature: env, TabularPolicy, both selectors, OneStepAgent; train -> evaluate -> reset."""
    n_agents, n_ep, seed = 6, 15, 0xABCDEF
    decay = 1.0 / (0.5 * n_ep)
    env = rlb.TaxiEnv(100)
    policy = rlb.TabularPolicy(0.05, 0.0)
    eg = rlb.UniformEpsilonGreed(1.0, ("sub", decay), 0.0)
    ucb = rlb.UpperConfidenceBound(0.5)
    agent = rlb.OneStepAgent(policy, 0.95, eg, rlb.sarsa, n_agents=n_agents, seed=seed, real="f64")
    agent.register_selector(ucb)
    cfg = O.make_config(O.ENV_TAXI, target=O.TARGET_SARSA, eps_decay=decay, seed=seed)
    sessions = [O.Session(cfg, i) for i in range(n_agents)]
    for sel_obj, sel_kind in ((eg, 0), (ucb, 1)):
        agent.set_action_selector(sel_obj)
        for func, tgt in ((rlb.sarsa, 0), (rlb.qlearning, 1), (rlb.expected_sarsa, 2)):
            agent.set_future_q_value_func(func)
            rewards, lengths, errors = agent.train(env, n_ep, n_ep // 10)
            ev_rewards, ev_lengths = agent.evaluate(env, n_ep)
            assert rewards.shape == (n_agents, n_ep) and lengths.dtype == np.uint64
            for i, s in enumerate(sessions):
                s.set_selector(sel_kind) if tgt == 0 else None
                s.set_target(tgt)
                ret, ln, tds, _ = s.train(n_ep, n_ep // 10)
                eret, eln = s.evaluate(n_ep)
                assert np.array_equal(lengths[i], ln) and np.array_equal(rewards[i], ret)
                assert P.bits_equal(errors[i], tds)
                assert np.array_equal(ev_lengths[i], eln) and np.array_equal(ev_rewards[i], eret)
            agent.reset()
            for s in sessions:
                s.agent_reset()
    for s in sessions:
        s.close()


def test_single_agent_trait_object(rlb):
    """n_agents = 1 is the reference's trait object: BASELINE config C1 (Blackjack Q-learning, eps-greedy, Basic) and the
    env used directly (reset / step / EnvNotReady / observation ids)."""
    env = rlb.BlackJackEnv()
    agent = rlb.OneStepAgent(rlb.TabularPolicy(0.05, 0.0), 0.95, rlb.UniformEpsilonGreed(1.0, ("sub", 2e-5), 0.0), rlb.qlearning,
                             n_agents=1, seed=0x5EED0001, real="f64")
    rewards, lengths, errors = agent.train(env, 2000, 200)
    s = O.Session(O.make_config(O.ENV_BLACKJACK, seed=0x5EED0001), 0)
    ret, ln, tds, _ = s.train(2000, 200)
    assert rewards.shape == (2000,) and np.array_equal(rewards, ret) and np.array_equal(lengths, ln)
    # one agent: training_error is the reference's per-STEP vector (agent.rs:98,117)
    assert errors.shape == (int(ln.sum()),) and P.bits_equal(errors, s.training_error())
    assert set(np.unique(rewards)) <= {-1.0, 0.0, 1.0}
    # the env trait on its own
    with pytest.raises(rlb.EnvNotReady):
        env.step(0)                                   # terminated by train(): Err(EnvNotReady)
    obs = env.reset()
    assert obs[0] == s.env_reset()
    assert rlb.BlackJackEnv.dense_index(rlb.BlackJackEnv.obs_id(obs[0])) == obs[0]
    o2, r, t = env.step(1)                            # STICK
    ref = s.env_step(1)
    assert (int(o2[0]), float(r[0]), bool(t[0])) == ref and ref[2] is True
    a = agent.get_action(o2)
    assert int(a[0]) == s.get_action(int(o2[0]))
    s.close()


def test_trace_agent_through_the_mirror(rlb):
    """BASELINE config C2's objects: FrozenLakeEnv(MAP_8X8, slippery), ElegibilityTracesAgent + sarsa."""
    n_agents, n_ep = 40, 20
    decay = 1.0 / (0.5 * n_ep)
    env = rlb.FrozenLakeEnv(rlb.FrozenLakeEnv.MAP_8X8, True, 100)
    agent = rlb.ElegibilityTracesAgent(rlb.TabularPolicy(0.05, 0.0), 0.95, rlb.UniformEpsilonGreed(1.0, ("sub", decay), 0.0), 0.5,
                                       rlb.sarsa, n_agents=n_agents, seed=77)
    res = agent.train(env, n_ep, n_ep // 10, raw=True)
    c = dict(env=1, agent=1, selector=0, policy=0, target=0, real=0)
    o = O.batch_train(P.oracle_config(c, P.hyper(n_ep, seed=77)), 0, n_agents, n_ep, n_ep // 10, n_threads=4)
    assert res["train_steps"] == o["train_steps"] and res["eval_steps"] == o["eval_steps"]
    assert np.array_equal(res["sums"][:, 0], o["len"].sum(0).astype(np.float64))
    assert agent.engine.cfg.store_kind == 0 and rlb.abi.lib.rlb_engine_store_kind(agent.engine.h) == 3   # auto -> hybrid store
    q, _ = agent.engine.download_tables()
    assert P.bits_equal(q.astype(np.float64), o["q"])


def test_dyna_through_the_mirror(rlb):
    """bin/cliffwalking_model.rs:136-203: `other` wrapped as InternalModelAgent(other, RandomModel, 10); train, evaluate,
    the Model trait on its own, reset, and the end of the borrow."""
    n_agents, n_ep, seed = 5, 24, 0xD17A
    decay = 1.0 / (0.5 * n_ep)
    env = rlb.CliffWalkingEnv(100)
    other = rlb.OneStepAgent(rlb.TabularPolicy(0.05, 0.0), 0.95, rlb.UniformEpsilonGreed(1.0, ("sub", decay), 0.0), rlb.qlearning,
                             n_agents=n_agents, seed=seed, real="f64")
    model = rlb.RandomModel()
    with pytest.raises(RuntimeError):
        model.reset()                                   # not bound yet
    with pytest.raises(ValueError):
        rlb.InternalModelAgent(other, model, 0)
    model_agent = rlb.InternalModelAgent(other, model, 10)
    rewards, lengths, errors = model_agent.train(env, n_ep, n_ep // 10)
    ev_rewards, ev_lengths = model_agent.evaluate(env, 6)
    sessions = [O.Session(O.make_config(O.ENV_CLIFF_WALKING, target=O.TARGET_QLEARNING, eps_decay=decay, seed=seed, planning_steps=10), i)
                for i in range(n_agents)]
    ln_m, ent = model.entries()
    info = model.get_info()
    for i, s in enumerate(sessions):
        ret, ln, tds, _ = s.train(n_ep, n_ep // 10)
        eret, eln = s.evaluate(6)
        assert np.array_equal(lengths[i], ln) and np.array_equal(rewards[i], ret) and P.bits_equal(errors[i], tds)
        assert np.array_equal(ev_lengths[i], eln) and np.array_equal(ev_rewards[i], eret)
        ms, ma, ms2, mr = s.model()
        assert int(ln_m[i]) == len(ms) and np.array_equal(ent[i, :len(ms)]["next_obs"], ms2)
        assert tuple(float(x[i]) for x in info) == tuple(float(x) for x in s.model_get_info())
    model_agent.reset()
    assert not model.entries()[0].any()
    model_agent.release()
    rewards, lengths, errors = other.train(env, 4, 2)   # plain Q-learning again, tables reset, stream carried on
    for i, s in enumerate(sessions):
        s.agent_reset()
        s.set_planning(0)
        ret, ln, tds, _ = s.train(4, 2)
        assert np.array_equal(lengths[i], ln) and P.bits_equal(errors[i], tds)
        s.close()


def test_frozen_lake_on_custom_rows(rlb):
    """api.FrozenLakeEnv(map, ..) with rows of the caller's own (frozen_lake.rs:48): trained through the mirror, equal to
    the oracle built on the same rows; render() draws the same map."""
    rows = ["FFFH", "FSFF", "HFFF", "FFFG"]
    env = rlb.FrozenLakeEnv(rows, True, 30)
    n_agents, n_ep, seed = 3, 25, 0xAB
    decay = 1.0 / (0.5 * n_ep)
    agent = rlb.ElegibilityTracesAgent(rlb.TabularPolicy(0.05, 0.0), 0.95, rlb.UniformEpsilonGreed(1.0, ("sub", decay), 0.0), 0.5, rlb.sarsa,
                                       n_agents=n_agents, seed=seed, real="f64")
    rewards, lengths, errors = agent.train(env, n_ep, 5)
    c = dict(env=1, agent=1, selector=0, policy=0, target=0, real=1)
    o = O.batch_train(P.oracle_config(c, P.hyper(n_ep, seed=seed, max_steps=30, map_rows=rows)), 0, n_agents, n_ep, 5, n_threads=2)
    assert np.array_equal(lengths, o["len"]) and P.bits_equal(rewards, o["ret"]) and P.bits_equal(errors, o["tdsum"])
    with pytest.raises(ValueError):
        rlb.FrozenLakeEnv(["SF", "F"], False, 10)
