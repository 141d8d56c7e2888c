"""GPU tests of the ABI v2 additions: the per-step TD stream (`training_error`, agent.rs:98), the asynchronous train
call, the fused one-transition call, argument validation, caller-supplied FrozenLake maps (frozen_lake.rs:48), the
NCCL gather behind rlb_comm_*, and the evaluate-return total."""
import numpy as np
import pytest

import parity as P
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu

CASES = [
    dict(env=3, agent=0, selector=0, policy=0, target=1, real=1),   # C4 family, reference arithmetic
    dict(env=3, agent=0, selector=0, policy=0, target=1, real=0),
    dict(env=1, agent=1, selector=0, policy=0, target=0, real=0),   # C2 family (hybrid store)
    dict(env=2, agent=0, selector=1, policy=1, target=2, real=1),   # C3 family
    dict(env=0, agent=1, selector=1, policy=1, target=0, real=0),
]


@pytest.mark.parametrize("c", CASES, ids=P.combo_id)
def test_td_stream_is_training_error(c):
    """rlb_train_out.td_steps == the reference's `training_error` vector (one TD per training step, episodes
    concatenated, agent.rs:98,117), value for value, also across chunked calls and past the capacity."""
    n_agents, n_ep, eval_at = 7, 12, 4
    h = P.hyper(n_ep)
    cfg = P.oracle_config(c, h)
    refs = []
    for i in range(n_agents):
        s = O.Session(cfg, i)
        s.train(n_ep, eval_at)
        refs.append(s.training_error())
        s.close()
    cap = max(len(r) for r in refs)
    with P.make_engine(c, h, n_agents) as eng:
        r = eng.train(n_ep, eval_at, td_capacity=cap)
    for i in range(n_agents):
        n = int(r["td_count"][i])
        assert n == len(refs[i])
        assert P.bits_equal(r["td_steps"][i, :n].astype(np.float64), refs[i]), P.first_diff(r["td_steps"][i, :n].astype(np.float64), refs[i])
    # chunked: every call returns the TDs of its own episodes; a capacity that is too small truncates, the count does not
    with P.make_engine(c, h, n_agents) as eng:
        a = eng.train(5, eval_at, td_capacity=cap)
        b = eng.train(n_ep, eval_at, ep_begin=5, td_capacity=3)
    for i in range(n_agents):
        na, nb = int(a["td_count"][i]), int(b["td_count"][i])
        assert na + nb == len(refs[i])
        assert P.bits_equal(a["td_steps"][i, :na].astype(np.float64), refs[i][:na])
        assert P.bits_equal(b["td_steps"][i, :min(nb, 3)].astype(np.float64), refs[i][na:na + min(nb, 3)])


@pytest.mark.parametrize("chunk", [16, 4])   # 16: records stream through the two scratch halves; 4: too short to pipeline
@pytest.mark.parametrize("c", [CASES[1], CASES[2]], ids=P.combo_id)
def test_async_train_equals_blocking(c, chunk):
    """rlb_agent_train_range_async + rlb_agent_train_wait: same records, sums, totals and tables as the blocking call,
    with two calls in flight and pinned host record buffers (the pipelined copy path)."""
    import torch
    n_agents, n_ep, eval_at = 300, 48, 12
    h = P.hyper(n_ep)
    ref = P.gpu_run(c, h, n_agents, n_ep, eval_at)
    with P.make_engine(c, h, n_agents) as eng:
        recs = [torch.zeros((chunk, n_agents, 4), dtype=torch.int32).pin_memory() for _ in range(n_ep // chunk)]
        sums = [torch.zeros((chunk, 4), dtype=torch.float64).pin_memory() for _ in range(n_ep // chunk)]
        for k in range(n_ep // chunk):
            eng.train((k + 1) * chunk, eval_at, ep_begin=k * chunk, sums_out=sums[k], episodes_out=recs[k], wait=False)
        done = eng.train_wait()
        q, counts = eng.download_tables()
        st = eng.states()
    assert len(done) == n_ep // chunk
    assert sum(d["train_steps"] for d in done) == ref["train_steps"] and sum(d["eval_steps"] for d in done) == ref["eval_steps"]
    eps = np.concatenate([r.numpy().view(eng.episode_dtype)[..., 0] for r in recs], 0)
    assert np.array_equal(eps["length"].T.astype(np.uint64), ref["len"])
    assert P.bits_equal(eps["ret"].T.astype(np.float64), ref["ret"])
    assert P.bits_equal(eps["td_sum"].T.astype(np.float64), ref["tdsum"])
    assert P.bits_equal(np.concatenate([s.numpy() for s in sums], 0), ref["sums"])
    assert P.bits_equal(q.astype(np.float64), ref["q"])
    assert np.array_equal(st["rng_n"], ref["state"]["rng_n"])


@pytest.mark.parametrize("c", [CASES[1], CASES[3]], ids=P.combo_id)
def test_async_device_resident_calls_overlap(c):
    """Asynchronous calls whose outputs stay on the device are queued back to back (two in flight, per-call blocks of
    totals and sums): same sums, totals and tables as the blocking calls."""
    import torch
    n_agents, n_ep, eval_at, chunk = 500, 40, 10, 8
    h = P.hyper(n_ep)
    ref = P.gpu_run(c, h, n_agents, n_ep, eval_at)
    with P.make_engine(c, h, n_agents) as eng:
        sums = [torch.zeros((chunk, 4), dtype=torch.float64, device="cuda") for _ in range(n_ep // chunk)]
        for k in range(n_ep // chunk):
            eng.train((k + 1) * chunk, eval_at, ep_begin=k * chunk, sums_out=sums[k], wait=False)
        done = eng.train_wait()
        torch.cuda.synchronize()
        q, counts = eng.download_tables()
        st = eng.states()
    assert [d["train_steps"] for d in done] == [int(ref["len"][:, k * chunk:(k + 1) * chunk].sum()) for k in range(n_ep // chunk)]
    assert sum(d["eval_steps"] for d in done) == ref["eval_steps"]
    assert P.bits_equal(torch.cat(sums).cpu().numpy(), ref["sums"])
    assert P.bits_equal(q.astype(np.float64), ref["q"]) and np.array_equal(st["rng_n"], ref["state"]["rng_n"])


@pytest.mark.parametrize("c", [CASES[0], CASES[2], CASES[3], CASES[4]], ids=P.combo_id)
def test_agent_step_reproduces_the_loop(c):
    """A host loop over rlb_agent_step == env.reset / get_action / env.step / get_action / update driven through the
    oracle in the order of agent.rs:83-106 — every observation, action, reward, flag and TD, then tables and streams."""
    n_agents, n_calls = 6, 260
    h = P.hyper(10, max_steps=15)
    cfg = P.oracle_config(c, h)
    sessions = [O.Session(cfg, i) for i in range(n_agents)]
    ready = [False] * n_agents
    cur = [(0, 0)] * n_agents
    with P.make_engine(c, h, n_agents) as eng:
        for _ in range(n_calls):
            r = eng.agent_step()
            for i, s in enumerate(sessions):
                if not ready[i]:
                    o = s.env_reset()
                    a = s.get_action(o)
                    assert (r["kind"][i], r["obs"][i], r["action"][i]) == (0, o, a)
                    ready[i], cur[i] = True, (o, a)
                else:
                    o, rew, term = s.env_step(cur[i][1])
                    a = s.get_action(o)
                    td = s.update(cur[i][0], cur[i][1], rew, term, o, a)
                    assert (r["kind"][i], r["obs"][i], r["action"][i], r["reward"][i], bool(r["terminated"][i])) == (1, o, a, rew, term)
                    assert P.bits_equal(np.float64(r["td"][i]), np.float64(td)), (i, r["td"][i], td)
                    ready[i], cur[i] = not term, (o, a)
        q, counts = eng.download_tables()
        st = eng.states()
    for i, s in enumerate(sessions):
        oq, oc, ost = s.export()
        assert P.bits_equal(q[i].astype(np.float64), oq)
        assert np.array_equal(counts[i].astype(np.uint64), oc)
        assert st["rng_n"][i] == ost.rng_n and st["policy_flag"][i] == ost.policy_flag
        s.close()


def test_step_level_calls_validate_their_arguments(rlb):
    """obs >= S / action >= A: RLB_ERR_INVALID_ARG and the agents are left untouched (the reference would panic)."""
    c = dict(env=3, agent=0, selector=1, policy=1, target=0, real=1)
    h = P.hyper(10)
    n = 4
    with P.make_engine(c, h, n) as eng:
        eng.train(5, 5)
        q0, c0 = eng.download_tables()
        st0 = eng.states()
        ok = np.zeros(n, np.uint32)
        bad_obs = np.array([0, 500, 0, 0], np.uint32)
        bad_act = np.array([0, 0, 6, 0], np.uint32)
        def snap(i):
            q, cnt = eng.download_tables()
            st = eng.states()
            return q[i].copy(), cnt[i].copy(), int(st["ucb_t"][i]), int(st["policy_flag"][i])

        def same(a, b):
            return P.bits_equal(a[0].astype(np.float64), b[0].astype(np.float64)) and np.array_equal(a[1], b[1]) and a[2:] == b[2:]

        def expect_invalid(call):
            with pytest.raises(rlb.RlbError) as ei:
                call()
            assert ei.value.status == rlb.abi.ERR_INVALID_ARG

        before = snap(1)   # agent 1 is the one with the bad observation
        for call in [lambda: eng.get_action(bad_obs), lambda: eng.policy_predict(bad_obs), lambda: eng.policy_get_values(bad_obs),
                     lambda: eng.update(bad_obs, ok, np.zeros(n), np.zeros(n, np.uint8), ok, ok),
                     lambda: eng.update(ok, ok, np.zeros(n), np.zeros(n, np.uint8), bad_obs, ok),
                     lambda: eng.policy_update(bad_obs, ok, ok, np.zeros(n)),
                     lambda: eng.selector_get_action(bad_obs, np.zeros((n, 6))), lambda: eng.selector_get_exploration_probs(bad_obs, np.zeros((n, 6)))]:
            expect_invalid(call)
        assert same(before, snap(1))
        assert not same((q0[0], c0[0], int(st0["ucb_t"][0]), int(st0["policy_flag"][0])), snap(0))   # the well-formed agents did run
        before = snap(2)   # agent 2 is the one with the bad action
        for call in [lambda: eng.update(ok, bad_act, np.zeros(n), np.zeros(n, np.uint8), ok, ok),
                     lambda: eng.update(ok, ok, np.zeros(n), np.zeros(n, np.uint8), ok, bad_act),
                     lambda: eng.policy_update(ok, bad_act, ok, np.zeros(n))]:
            expect_invalid(call)
        eng.env_reset()
        expect_invalid(lambda: eng.env_step(bad_act))
        assert same(before, snap(2))
        eng.set_model(3)
        with pytest.raises(rlb.RlbError) as ei:
            eng.model_add_info(bad_obs, ok, np.zeros(n), ok)
        assert ei.value.status == rlb.abi.ERR_INVALID_ARG
        ln = np.array([1, 1, 1, 1], np.uint32)
        ent = np.zeros((n, 3000), rlb.abi.MODEL_ENTRY)
        ent["obs"][2, 0] = 777
        with pytest.raises(rlb.RlbError) as ei:
            eng.upload_model(ln, ent)
        assert ei.value.status == rlb.abi.ERR_INVALID_ARG
    # odd stream positions are refused on envs that only draw 64-bit values
    with P.make_engine(c, h, n) as eng:
        st = eng.states()
        st["rng_n"][1] = 7
        with pytest.raises(rlb.RlbError) as ei:
            eng.set_states(st)
        assert ei.value.status == rlb.abi.ERR_INVALID_ARG


MAPS = {
    "two_starts_5x7": ["SFFFFFH", "FFHFFFF", "FFFFHFS", "HFFFFFF", "FFFHFFG"],                  # 35 cells: on-chip stores; start drawn among 2 cells
    "three_starts_9x9": ["SFFFFFFFF", "FFFHFFFFF", "FFFFFFHFF", "FSFFFFFFF", "FFFFHFFFF", "FFHFFFFFH", "FFFFFFFFS", "FHFFFFHFF", "FFFFFFFFG"],   # 81 cells: past the 64-state row lut
    "start_not_first_4x4": ["FFFH", "FSFF", "HFFF", "FFFG"],
    "no_start_3x3": ["FFF", "FHF", "FFG"],                                                       # all-zero start distribution -> cell 0
}


@pytest.mark.parametrize("map_name", sorted(MAPS))
@pytest.mark.parametrize("agent,real,slippery", [(0, 1, True), (1, 0, True), (1, 1, False)])
def test_custom_frozen_lake_maps(map_name, agent, real, slippery):
    """FrozenLakeEnv::new(map, ..) on caller-supplied rows (frozen_lake.rs:48): transition table, rewards, holes, the
    categorical start draw over every 'S' cell (:54-66,106-109) — bit-exact against the oracle's own constructor."""
    c = dict(env=1, agent=agent, selector=0, policy=0, target=1, real=real)
    n_agents, n_ep, eval_at = 64, 30, 10
    h = P.hyper(n_ep, slippery=slippery, max_steps=40, map_rows=MAPS[map_name])
    o = O.batch_train(P.oracle_config(c, h), 0, n_agents, n_ep, eval_at, n_threads=4)
    g = P.gpu_run(c, h, n_agents, n_ep, eval_at)
    P.compare(g, o, c)
    if map_name == "two_starts_5x7":   # both start cells are really used
        with P.make_engine(c, h, 256) as eng:
            obs = eng.env_reset()
        assert set(obs.tolist()) == {0, 20}


def test_custom_map_errors(rlb):
    for rows in (["SFX", "FFG"], []):
        with pytest.raises((rlb.RlbError, ValueError)):
            rlb.Engine(1, n_agents=1, map_rows=rows)
    with pytest.raises(rlb.RlbError):
        rlb.Engine(1, n_agents=1, map_rows=["F" * 40] * 40)   # 1600 cells > 1024


def test_evaluate_return_total_is_exact_and_repeatable():
    """rlb_train_out.eval_return_sum: an exact integer total (every reward is an integer), identical run to run and equal
    to the oracle's sum over the injected evaluate episodes."""
    c = dict(env=3, agent=0, selector=0, policy=0, target=1, real=0)
    n_agents, n_ep, eval_at = 4096, 8, 4
    h = P.hyper(n_ep)
    vals = []
    for _ in range(3):
        with P.make_engine(c, h, n_agents) as eng:
            vals.append(eng.train(n_ep, eval_at, sums=False)["eval_return_sum"])
    assert vals[0] == vals[1] == vals[2] and float(vals[0]).is_integer()


def test_comm_gather_single_process(rlb):
    """rlb_comm_*: one communicator per visible GPU in ONE process (rlb_comm_init_all), grouped gather of [E,4] sums to
    rank 0 — the ctypes form of what a Rust host would call.  With one GPU this is the world-size-1 path."""
    import torch
    n_dev = min(rlb.abi.lib.rlb_device_count(), 2)
    comms = rlb.abi.Comm.init_all(list(range(n_dev)))
    assert [c.rank for c in comms] == list(range(n_dev)) and all(c.world_size == n_dev for c in comms)
    E = 37
    local = [torch.arange(E * 4, dtype=torch.float64, device="cuda:%d" % d).reshape(E, 4) + 1000.0 * d for d in range(n_dev)]
    out = torch.zeros((n_dev, E, 4), dtype=torch.float64, device="cuda:0")
    for d in range(n_dev):
        torch.cuda.synchronize(d)
    rlb.abi.check(rlb.abi.lib.rlb_comm_group_begin())
    for d, cm in enumerate(comms):
        cm.gather_episode_sums(local[d], out if d == 0 else None, root=0, stream=torch.cuda.current_stream(d).cuda_stream)
    rlb.abi.check(rlb.abi.lib.rlb_comm_group_end())
    for d in range(n_dev):
        torch.cuda.synchronize(d)
    for d in range(n_dev):
        assert torch.equal(out[d].cpu(), local[d].cpu())
    tot = [torch.full((3,), float(d + 1), dtype=torch.float64, device="cuda:%d" % d) for d in range(n_dev)]
    rlb.abi.check(rlb.abi.lib.rlb_comm_group_begin())
    for d, cm in enumerate(comms):
        cm.allreduce_sum(tot[d], stream=torch.cuda.current_stream(d).cuda_stream)
    rlb.abi.check(rlb.abi.lib.rlb_comm_group_end())
    for d in range(n_dev):
        torch.cuda.synchronize(d)
        assert tot[d].tolist() == [float(sum(range(1, n_dev + 1)))] * 3
    with pytest.raises(rlb.RlbError):
        comms[0].gather_episode_sums(local[0], out, root=5)
    for cm in comms:
        cm.close()


@pytest.mark.parametrize("env,real,max_steps", [(3, 1, 150), (3, 0, 250), (1, 1, 100), (0, 0, 100), (2, 1, 90)])
@pytest.mark.parametrize("policy,selector,target", [(0, 0, 0), (1, 0, 1), (1, 1, 2), (0, 1, 1)])
def test_lazy_trace_sweeps_equal_eager(env, real, max_steps, policy, selector, target):
    """store_kind 4 — a trace agent's sweeps recorded and applied to a row only when it is next read or the trace is
    cleared — against the oracle, with episodes longer than the shared-memory TD history (f64: 64 sweeps, f32: 128: the
    history fills and every row is brought up to date mid-episode), Double tables (the write table alternates per sweep)
    and UCB; then the step-level API leaves a trace behind and the fused kernel picks it up."""
    c = dict(env=env, agent=1, selector=selector, policy=policy, target=target, real=real)
    n_agents, n_ep, eval_at = 37, 16, 4
    h = P.hyper(n_ep, max_steps=max_steps, lambda_=0.9)
    o = O.batch_train(P.oracle_config(c, h), 0, n_agents, n_ep, eval_at, n_threads=4)
    g = P.gpu_run(c, h, n_agents, n_ep, eval_at, store_kind=4)
    P.compare(g, o, c)
    if env == 3 and policy == 1 and selector == 0:
        # a trace left by step-level update() calls carries into the fused kernel (the map survives between calls)
        sessions = [O.Session(P.oracle_config(c, h), i) for i in range(4)]
        with P.make_engine(c, h, 4, store_kind=4) as eng:
            obs = eng.env_reset(); act = eng.get_action(obs)
            ref_o = [s.env_reset() for s in sessions]; ref_a = [s.get_action(x) for s, x in zip(sessions, ref_o)]
            for _ in range(5):
                o2, r, t = eng.env_step(act); a2 = eng.get_action(o2)
                eng.update(obs, act, r, t, o2, a2)
                for k, s in enumerate(sessions):
                    x, rr, tt = s.env_step(ref_a[k]); y = s.get_action(x)
                    s.update(ref_o[k], ref_a[k], rr, tt, x, y)
                    ref_o[k], ref_a[k] = x, y
                obs, act = o2, a2
            res = eng.train(6, 3, sums=False, episodes=True)
            q, counts = eng.download_tables()
        for k, s in enumerate(sessions):
            ret, ln, tds, _ = s.train(6, 3)
            assert np.array_equal(res["episodes"]["length"][:, k], ln)
            assert P.bits_equal(res["episodes"]["td_sum"][:, k].astype(np.float64), tds)
            assert P.bits_equal(q[k].astype(np.float64), s.export()[0])
            s.close()


def test_lazy_store_row_limit(rlb):
    """The lazy store keeps an agent's visited-row slots in 8 bits: more than 255 possible rows per episode is refused when
    asked for explicitly, and the automatic choice falls back to the eager HBM store."""
    c = dict(env=3, agent=1, selector=0, policy=0, target=0, real=0)
    h = P.hyper(8, max_steps=300)
    with pytest.raises(rlb.RlbError) as ei:
        P.make_engine(c, h, 8, store_kind=4)
    assert ei.value.status == rlb.abi.ERR_UNSUPPORTED
    with P.make_engine(c, h, 8) as eng:
        assert eng.store_kind() == 1
    with P.make_engine(c, P.hyper(8, max_steps=200), 8) as eng:
        assert eng.store_kind() == 4
