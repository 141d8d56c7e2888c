"""CPU test: the C++ oracle against the independent pure-Python restatement (oracle/pyref.py), f64 mode,
step by step — observation, chosen action, reward, termination flag, TD — and final tables.  Two
restatements written separately from the reference sources must agree before either is trusted."""
import itertools

import numpy as np
import pytest

from oracle import oracle_py as O
from oracle import pyref as R

CASES = []
for env in range(4):
    for traces, sel, pol, tgt in itertools.product((0, 1), (0, 1), (0, 1), (0, 1, 2)):
        CASES.append((env, traces, sel, pol, tgt))


@pytest.mark.parametrize("env,traces,sel,pol,tgt", CASES)
def test_oracle_matches_python_restatement(env, traces, sel, pol, tgt):
    n_ep, eval_at, agent_id, seed = 5, 3, 7 + env, 0xC0FFEE
    eps_decay = 1.0 / (0.5 * n_ep)
    max_steps = 40
    cfg = O.make_config(env, map_id=1, slippery=1, max_steps=max_steps, policy=pol, selector=sel, target=tgt, agent=traces,
                        real=O.REAL_F64, eps_decay=eps_decay, seed=seed)
    s = O.Session(cfg, agent_id)
    s.record()
    ret, ln, tds, tda = s.train(n_ep, eval_at)
    tr = s.trajectory()
    q, counts, st = s.export()
    s.close()

    penv, pagent, prng = R.build(env, agent_id=agent_id, seed=seed, map_id=1, slippery=True, max_steps=max_steps, policy=pol,
                                 selector=sel, target=tgt, traces=bool(traces), eps_decay=eps_decay)
    log = []
    rewards, lengths, errors = pagent.train(penv, n_ep, eval_at, log)
    assert len(log) == len(tr)
    kind = np.array([l[0] for l in log]); obs = np.array([l[1] for l in log]); act = np.array([l[2] for l in log])
    rew = np.array([l[3] for l in log]); term = np.array([l[4] for l in log]); td = np.array([l[5] for l in log], np.float64)
    assert np.array_equal(kind, tr["kind"]) and np.array_equal(obs, tr["obs"]) and np.array_equal(act, tr["action"])
    assert np.array_equal(rew, tr["reward"]) and np.array_equal(term, tr["terminated"].astype(bool))
    same = (td.view(np.uint64) == tr["td"].view(np.uint64)) | (np.isnan(td) & np.isnan(tr["td"]))
    assert same.all(), np.argwhere(~same)[:3]
    assert list(ln) == lengths and prng.n == st.rng_n and pagent.eval_steps == st.eval_steps
    # final tables
    dense_of = {}
    if env == 0:
        for p in range(4, 32):
            for d in range(1, 27):
                for a in (0, 1):
                    dense_of[R.fxhash3(p, d, a)] = ((p - 4) * 26 + (d - 1)) * 2 + a
    for ti, tab in enumerate(pagent.policy.tables()):
        full = np.zeros_like(q[ti])
        for o, row in tab.items():
            full[dense_of.get(o, o) if env == 0 else o] = row
        a, b = full, q[ti]
        ok = (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))
        assert ok.all()
    if sel == 1:
        assert pagent.selector.t == st.ucb_t
        for o, row in pagent.selector.n.items():
            assert list(counts[dense_of.get(o, o) if env == 0 else o]) == row
    else:
        assert pagent.selector.eps == st.epsilon
    if pol == 1:
        assert int(pagent.policy.flag) == st.policy_flag
