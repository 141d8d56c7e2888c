"""The argument behind `store_kind` 4 (rlb_device.cuh `AgentCore::lz_*`), checked on the CPU in numpy f32/f64:
the reference sweeps EVERY row of the trace map at EVERY update (elegibility_traces_agent.rs:86-96); recording one TD per
update and replaying a row's pending sweeps only when that row is read (or the trace is cleared) performs the same
operations on the same cells in the same order — so the tables come out bit-identical, Double tables included
(double_tabular_policy.rs:50-67: the written table alternates per update).  This is a model of the bookkeeping
(stamps, history, flag parity), not of the CUDA code; the GPU tests compare the kernels themselves with the oracle."""
import numpy as np
import pytest


def run(real, double, lazy, seed, n_states=12, n_actions=3, n_updates=400):
    rng = np.random.default_rng(seed)
    lr, gamma, lam = real(0.37), real(0.93), real(0.81)
    gl = real(gamma * lam)
    tables = 2 if double else 1
    q = rng.standard_normal((tables, n_states, n_actions)).astype(real)
    flag = True
    trace = {}                                    # state -> e row, insertion-ordered like the visit list
    hist, flag0, stamp = [], flag, {}             # lazy: recorded TDs, flag at sweep 0, sweeps already applied per row
    td_log = []

    def replay(s):
        e = trace[s]
        for u in range(stamp[s], len(hist)):
            second = double and (flag0 != bool(u & 1))
            for k in range(n_actions):
                d = real(lr * real(hist[u] * e[k]))
                q[1 if second else 0, s, k] = real(q[1 if second else 0, s, k] + d)
                e[k] = real(e[k] * gl)
        stamp[s] = len(hist)

    s = int(rng.integers(n_states))
    for _ in range(n_updates):
        a, o = int(rng.integers(n_actions)), int(rng.integers(n_states))
        terminated = rng.random() < 0.05
        reward = real(rng.integers(-10, 3))
        read_tbl = 1 if (double and not flag) else 0
        write_tbl = 1 if (double and flag) else 0
        if lazy:
            if o in trace:
                replay(o)                          # next_q_values(o) is about to be read
            if s in trace:
                replay(s)                          # and so is Q[s][a]
        future = real(q[read_tbl, o].max())
        td = real(real(reward + real(gamma * future)) - q[read_tbl, s, a])
        td_log.append(td)
        if s not in trace:
            trace[s] = np.zeros(n_actions, real)
            stamp[s] = len(hist)
        trace[s][a] = real(trace[s][a] + real(1.0))
        if lazy:
            if not hist:
                flag0 = flag
            stamp[s] = len(hist)
            hist.append(td)
        else:
            for st, e in trace.items():            # the reference's sweep
                for k in range(n_actions):
                    q[write_tbl, st, k] = real(q[write_tbl, st, k] + real(lr * real(td * e[k])))
                    e[k] = real(e[k] * gl)
        if double:
            flag = not flag
        if terminated:
            if lazy:
                for st in list(trace):
                    replay(st)
                hist, stamp = [], {}
            trace = {}
            s = int(rng.integers(n_states))
        else:
            s = o
    if lazy:
        for st in list(trace):
            replay(st)
    return q, np.array(td_log, real)


@pytest.mark.parametrize("real", [np.float32, np.float64])
@pytest.mark.parametrize("double", [False, True])
def test_lazy_replay_equals_eager_sweeps(real, double):
    for seed in range(6):
        qe, te = run(real, double, False, seed)
        ql, tl = run(real, double, True, seed)
        assert qe.tobytes() == ql.tobytes()
        assert te.tobytes() == tl.tobytes()
