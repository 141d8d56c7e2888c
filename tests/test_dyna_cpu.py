"""CPU tests of the Dyna restatement (agent/internal_model_agent.rs, model/random_model.rs): the one-shot integer sampler
behind `gen_range`, a hand-worked planning case, and the C++ oracle against the independent Python restatement."""
import ctypes as C
import itertools

import numpy as np
import pytest

from oracle import oracle_py as O
from oracle import pyref as R


def test_gen_range_zone_and_samples(rlb):
    """rand 0.8.5 `UniformInt::sample_single_inclusive`: zone = (range << lzcnt(range)) - 1, accept iff lo <= zone."""
    seed, agent = 0xABCDEF12345, 3
    for rng_size in (1, 2, 3, 5, 6, 148, 192, 2912, 3000, (1 << 40) + 12345):
        n = 200
        out = np.zeros(n, np.float64)
        used = O.lib().oracle_sample(seed, agent, 0, 3, rng_size, n, out.ctypes.data_as(C.c_void_p))
        s = R.Stream(seed, agent)
        want = [s.gen_range(rng_size) for _ in range(n)]
        assert list(out.astype(np.uint64)) == want and used == s.n
        assert max(want) < rng_size
        idx = C.c_uint64(0)
        got = [rlb._abi.lib.rlb_rng_gen_range(seed, agent, C.byref(idx), rng_size) for _ in range(n)]
        assert got == want and idx.value == s.n
    # the conservative zone rejects: a power of two (1 included) drops half of the words, range 3 a quarter
    words = R.Stream(seed, agent)
    v = [words.u64() for _ in range(4000)]
    for rng_size, zone in ((1, (1 << 63) - 1), (3, (3 << 62) - 1), (4, (1 << 63) - 1), (148, (148 << 56) - 1)):
        s = R.Stream(seed, agent)
        for _ in range(500):
            s.gen_range(rng_size)
        accepted = 0
        k = 0
        while accepted < 500:
            accepted += ((v[k] * rng_size) & R.M64) <= zone
            k += 1
        assert s.n == 2 * k


def test_hand_worked_planning_steps():
    """CliffWalking, Q-learning, lr 0.05, gamma 0.95, eps-greedy with eps = 0 (no selector draws), 2 planning steps.
    update(36, up, -1, 24): td = -1, Q[36][up] = -0.05; the model then holds that one transition and replays it twice:
    td = -1 + 0.95*max(Q[24]) - Q[36][up]."""
    cfg = O.make_config(O.ENV_CLIFF_WALKING, real=O.REAL_F64, eps0=0.0, planning_steps=2)
    s = O.Session(cfg, 0)
    td = s.update(36, 3, -1.0, False, 24, 0)
    assert td == -1.0
    q = -0.05
    for _ in range(2):
        q = q + 0.05 * (-1.0 + 0.95 * 0.0 - q)
    got, _, st = s.export()
    assert got[0, 36, 3] == q and np.count_nonzero(got) == 1
    assert [list(x) for x in s.model()] == [[36], [3], [24], [-1.0]]
    # each replay drew gen_range(0..1): u64 words until one is below 2^63; nothing else consumed (eps == 0)
    ref = R.Stream(cfg.seed, 0)
    ref.gen_range(1), ref.gen_range(1)
    assert st.rng_n == ref.n
    # a second sighting of (36, up) with another outcome is ignored: `.entry().or_insert()`
    s.update(36, 3, -100.0, True, 37, 0)
    assert [list(x) for x in s.model()] == [[36], [3], [24], [-1.0]]
    s.update(24, 2, -1.0, False, 25, 1)
    assert [list(x) for x in s.model()] == [[36, 24], [3, 2], [24, 25], [-1.0, -1.0]]
    s.agent_reset()
    assert len(s.model()[0]) == 0 and not s.export()[0].any()


CASES = [(env, traces, sel, (env + traces + sel) % 2, (env + 2 * traces + sel) % 3, 1 + (env + sel) % 3)
         for env, traces, sel in itertools.product(range(4), (0, 1), (0, 1))]


@pytest.mark.parametrize("env,traces,sel,pol,tgt,planning", CASES)
def test_oracle_dyna_matches_python_restatement(env, traces, sel, pol, tgt, planning):
    n_ep, eval_at, agent_id, seed, max_steps = 4, 3, 11 + env, 0xD1A, 30
    eps_decay = 1.0 / (0.5 * n_ep)
    cfg = O.make_config(env, map_id=1, slippery=1, max_steps=max_steps, policy=pol, selector=sel, target=tgt, agent=traces,
                        real=O.REAL_F64, eps_decay=eps_decay, seed=seed, planning_steps=planning)
    s = O.Session(cfg, agent_id)
    s.record()
    ret, ln, tds, tda = s.train(n_ep, eval_at)
    tr = s.trajectory()
    q, counts, st = s.export()
    ms, ma, ms2, mr = s.model()
    s.close()

    penv, pagent, prng = R.build(env, agent_id=agent_id, seed=seed, map_id=1, slippery=True, max_steps=max_steps, policy=pol,
                                 selector=sel, target=tgt, traces=bool(traces), eps_decay=eps_decay, planning=planning)
    log = []
    rewards, lengths, errors = pagent.train(penv, n_ep, eval_at, log)
    assert len(log) == len(tr) and list(ln) == lengths and prng.n == st.rng_n
    assert np.array_equal(np.array([l[1] for l in log]), tr["obs"]) and np.array_equal(np.array([l[2] for l in log]), tr["action"])
    td = np.array([l[5] for l in log], np.float64)
    assert ((td.view(np.uint64) == tr["td"].view(np.uint64)) | (np.isnan(td) & np.isnan(tr["td"]))).all()
    dense_of = {}
    if env == 0:
        for p, d, a in itertools.product(range(4, 32), range(1, 27), (0, 1)):
            dense_of[R.fxhash3(p, d, a)] = ((p - 4) * 26 + (d - 1)) * 2 + a
    dense = (lambda o: dense_of[o]) if env == 0 else (lambda o: o)
    for ti, tab in enumerate(pagent.inner.policy.tables()):
        full = np.zeros_like(q[ti])
        for o, row in tab.items():
            full[dense(o)] = row
        assert ((full.view(np.uint64) == q[ti].view(np.uint64)) | (np.isnan(full) & np.isnan(q[ti]))).all()
    items = list(pagent.model.items())
    assert [dense(k[0]) for k, _ in items] == list(ms) and [k[1] for k, _ in items] == list(ma)
    assert [dense(v[0]) for _, v in items] == list(ms2) and [v[1] for _, v in items] == list(mr)
