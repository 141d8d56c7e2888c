"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  The product package (rl-rust_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ENV_BLACKJACK, ENV_FROZEN_LAKE, ENV_CLIFF_WALKING, ENV_TAXI = 0, 1, 2, 3
POLICY_BASIC, POLICY_DOUBLE = 0, 1
SEL_EPS_GREEDY, SEL_UCB = 0, 1
TARGET_SARSA, TARGET_QLEARNING, TARGET_EXPECTED_SARSA = 0, 1, 2
AGENT_ONE_STEP, AGENT_TRACES = 0, 1
REAL_F32, REAL_F64 = 0, 1
DECAY_SUB, DECAY_MUL = 0, 1


class OracleConfig(C.Structure):
    _fields_ = [
        ("env_kind", C.c_int32), ("map_id", C.c_int32), ("slippery", C.c_int32), ("max_steps", C.c_uint32),
        ("policy_kind", C.c_int32), ("selector_kind", C.c_int32), ("target_kind", C.c_int32),
        ("agent_kind", C.c_int32), ("real_kind", C.c_int32), ("decay_kind", C.c_int32),
        ("lr", C.c_double), ("gamma", C.c_double), ("lambda_", C.c_double), ("eps0", C.c_double),
        ("eps_decay", C.c_double), ("eps_final", C.c_double), ("ucb_c", C.c_double), ("default_q", C.c_double),
        ("seed", C.c_uint64), ("planning_steps", C.c_uint32), ("pad", C.c_uint32),
        ("map_rows", C.c_uint32), ("map_cols", C.c_uint32), ("map", C.c_char_p),
    ]


class OracleState(C.Structure):
    _fields_ = [("epsilon", C.c_double), ("ucb_t", C.c_uint64), ("rng_n", C.c_uint64), ("eval_steps", C.c_uint64),
                ("policy_flag", C.c_int32), ("pad", C.c_int32)]


TRAJ_DTYPE = np.dtype([("kind", "u1"), ("action", "u1"), ("terminated", "u1"), ("pad", "u1"), ("obs", "<u4"),
                       ("reward", "<f8"), ("td", "<f8")])
STATE_DTYPE = np.dtype([("epsilon", "<f8"), ("ucb_t", "<u8"), ("rng_n", "<u8"), ("eval_steps", "<u8"),
                        ("policy_flag", "<i4"), ("pad", "<i4")])


def build(force=False):
    """Compile oracle/liboracle.so with the committed Makefile (g++ only)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle.hpp", "oracle_capi.cpp", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = C.CDLL(so)
        vp, u64, u32, i32, dbl = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_double
        P = C.POINTER
        L.oracle_create.restype = vp
        L.oracle_create.argtypes = [P(OracleConfig), u64]
        L.oracle_destroy.argtypes = [vp]
        for f in ("oracle_n_states", "oracle_n_actions", "oracle_n_tables"):
            getattr(L, f).restype = u32
            getattr(L, f).argtypes = [vp]
        L.oracle_train.restype = i32
        L.oracle_train.argtypes = [vp, u64, u64, u64, vp, vp, vp, vp]
        L.oracle_evaluate.restype = i32
        L.oracle_evaluate.argtypes = [vp, u64, vp, vp]
        L.oracle_agent_reset.argtypes = [vp]
        L.oracle_set_target.argtypes = [vp, i32]
        L.oracle_set_selector.argtypes = [vp, i32]
        L.oracle_set_agent_kind.argtypes = [vp, i32]
        L.oracle_export.argtypes = [vp, vp, vp, P(OracleState)]
        L.oracle_training_error_len.restype = u64
        L.oracle_training_error_len.argtypes = [vp]
        L.oracle_training_error_copy.argtypes = [vp, vp]
        L.oracle_record.argtypes = [vp, i32]
        L.oracle_traj_len.restype = u64
        L.oracle_traj_len.argtypes = [vp]
        L.oracle_traj_copy.argtypes = [vp, vp]
        L.oracle_traj_clear.argtypes = [vp]
        L.oracle_env_reset.restype = u32
        L.oracle_env_reset.argtypes = [vp]
        L.oracle_env_step.restype = i32
        L.oracle_env_step.argtypes = [vp, u32, P(u32), P(dbl), P(i32)]
        L.oracle_get_action.restype = u32
        L.oracle_get_action.argtypes = [vp, u32]
        L.oracle_update.restype = dbl
        L.oracle_update.argtypes = [vp, u32, u32, dbl, i32, u32, u32]
        L.oracle_set_planning.argtypes = [vp, u32]
        L.oracle_model_len.restype = u64
        L.oracle_model_len.argtypes = [vp]
        L.oracle_model_copy.argtypes = [vp, vp, vp, vp, vp]
        L.oracle_model_add_info.argtypes = [vp, u32, u32, dbl, u32]
        L.oracle_model_get_info.restype = i32
        L.oracle_model_get_info.argtypes = [vp, P(u32), P(u32), P(u32), P(dbl)]
        L.oracle_model_reset.argtypes = [vp]
        L.oracle_batch_train.restype = i32
        L.oracle_batch_train.argtypes = [P(OracleConfig), u64, u64, u64, u64, i32, vp, vp, vp, vp, vp, vp, vp,
                                         P(dbl), P(u64), P(u64)]
        L.oracle_philox4x32_10.argtypes = [vp, vp, vp]
        L.oracle_stream_words.argtypes = [u64, u64, u64, u64, vp]
        L.oracle_sample.restype = u64
        L.oracle_sample.argtypes = [u64, u64, u64, i32, u64, u64, vp]
        L.oracle_fxhash_blackjack.restype = u64
        L.oracle_fxhash_blackjack.argtypes = [u32, u32, i32]
        L.oracle_blackjack_dense.restype = u32
        L.oracle_blackjack_dense.argtypes = [u32, u32, i32]
        L.oracle_log.restype = dbl
        L.oracle_log.argtypes = [dbl]
        L.oracle_categorical_sample.restype = u64
        L.oracle_categorical_sample.argtypes = [vp, u64, dbl]
        L.oracle_argmax.restype = u64
        L.oracle_argmax.argtypes = [vp, u64]
        L.oracle_taxi_table.argtypes = [vp, vp, vp, vp]
        L.oracle_cliff_table.argtypes = [vp, vp, vp]
        L.oracle_frozen_lake_table.argtypes = [i32, i32, vp, vp, vp, vp]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


ENV_DIMS = {ENV_BLACKJACK: (1456, 2), ENV_CLIFF_WALKING: (48, 4), ENV_TAXI: (500, 6)}


def env_dims(cfg):
    if cfg.env_kind == ENV_FROZEN_LAKE:
        if cfg.map_id == 2:
            return (cfg.map_rows * cfg.map_cols, 4)
        return (16 if cfg.map_id == 0 else 64, 4)
    return ENV_DIMS[cfg.env_kind]


def make_config(env_kind, *, map_id=1, slippery=0, max_steps=100, policy=POLICY_BASIC, selector=SEL_EPS_GREEDY,
                target=TARGET_QLEARNING, agent=AGENT_ONE_STEP, real=REAL_F64, decay_kind=DECAY_SUB, lr=0.05,
                gamma=0.95, lambda_=0.5, eps0=1.0, eps_decay=2e-5, eps_final=0.0, ucb_c=0.5, default_q=0.0,
                seed=0x5EED0001, planning_steps=0, map_rows=None):
    """Defaults = the reference CLI's (bin/taxi.rs:22-68) with n_episodes=100000 -> decay 2e-5.
    map_rows: a caller-supplied FrozenLake map (list of equally long S/F/H/G strings, frozen_lake.rs:48)."""
    rows, cols, flat = 0, 0, None
    if map_rows is not None:
        rows, cols, flat, map_id = len(map_rows), len(map_rows[0]), "".join(map_rows).encode("ascii"), 2
    return OracleConfig(env_kind, map_id, int(slippery), max_steps, policy, selector, target, agent, real,
                        decay_kind, lr, gamma, lambda_, eps0, eps_decay, eps_final, ucb_c, default_q, seed,
                        planning_steps, 0, rows, cols, flat)


class Session:
    """One reference agent + env pair on one RNG stream (what the reference CLI builds)."""

    def __init__(self, cfg, agent_id=0):
        self.L = lib()
        self.cfg = cfg
        self.h = self.L.oracle_create(C.byref(cfg), agent_id)
        self.S = self.L.oracle_n_states(self.h)
        self.A = self.L.oracle_n_actions(self.h)
        self.T = self.L.oracle_n_tables(self.h)

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def train(self, n_episodes, eval_at, ep_begin=0):
        n = n_episodes - ep_begin
        ret = np.zeros(n, np.float64)
        ln = np.zeros(n, np.uint64)
        tds = np.zeros(n, np.float64)
        tda = np.zeros(n, np.float64)
        rc = self.L.oracle_train(self.h, ep_begin, n_episodes, eval_at, _p(ret), _p(ln), _p(tds), _p(tda))
        if rc:
            raise RuntimeError("oracle_train rc=%d" % rc)
        return ret, ln, tds, tda

    def training_error(self):
        n = self.L.oracle_training_error_len(self.h)
        out = np.zeros(n, np.float64)
        self.L.oracle_training_error_copy(self.h, _p(out))
        return out

    def evaluate(self, n):
        ret = np.zeros(n, np.float64)
        ln = np.zeros(n, np.uint64)
        rc = self.L.oracle_evaluate(self.h, n, _p(ret), _p(ln))
        if rc:
            raise RuntimeError("oracle_evaluate rc=%d" % rc)
        return ret, ln

    def agent_reset(self):
        self.L.oracle_agent_reset(self.h)

    def set_target(self, k):
        self.L.oracle_set_target(self.h, k)

    def set_selector(self, k):
        self.L.oracle_set_selector(self.h, k)

    def set_agent_kind(self, k):
        self.L.oracle_set_agent_kind(self.h, k)

    def export(self):
        q = np.zeros((self.T, self.S, self.A), np.float64)
        counts = np.zeros((self.S, self.A), np.uint64)
        st = OracleState()
        self.L.oracle_export(self.h, _p(q), _p(counts), C.byref(st))
        return q, counts, st

    def record(self, on=True):
        self.L.oracle_record(self.h, 1 if on else 0)

    def trajectory(self, clear=True):
        n = self.L.oracle_traj_len(self.h)
        out = np.zeros(n, TRAJ_DTYPE)
        if n:
            self.L.oracle_traj_copy(self.h, _p(out))
        if clear:
            self.L.oracle_traj_clear(self.h)
        return out

    def env_reset(self):
        return self.L.oracle_env_reset(self.h)

    def env_step(self, action):
        obs, rew, term = C.c_uint32(), C.c_double(), C.c_int()
        rc = self.L.oracle_env_step(self.h, action, C.byref(obs), C.byref(rew), C.byref(term))
        if rc:
            return None   # EnvNotReady
        return obs.value, rew.value, bool(term.value)

    def get_action(self, obs):
        return self.L.oracle_get_action(self.h, obs)

    def update(self, s, a, r, term, s2, a2):
        return self.L.oracle_update(self.h, s, a, r, int(term), s2, a2)

    # InternalModelAgent / RandomModel (agent/internal_model_agent.rs, model/random_model.rs)
    def set_planning(self, steps):
        self.L.oracle_set_planning(self.h, steps)

    def model(self):
        """The model's entries in insertion order: arrays (obs, action, next_obs, reward)."""
        n = self.L.oracle_model_len(self.h)
        s, a, s2 = (np.zeros(n, np.uint32) for _ in range(3))
        r = np.zeros(n, np.float64)
        if n:
            self.L.oracle_model_copy(self.h, _p(s), _p(a), _p(s2), _p(r))
        return s, a, s2, r

    def model_add_info(self, s, a, r, s2):
        self.L.oracle_model_add_info(self.h, s, a, r, s2)

    def model_get_info(self):
        s, a, s2, r = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_double()
        if self.L.oracle_model_get_info(self.h, C.byref(s), C.byref(a), C.byref(s2), C.byref(r)):
            return None
        return s.value, a.value, s2.value, r.value

    def model_reset(self):
        self.L.oracle_model_reset(self.h)


def batch_train(cfg, first_agent, count, n_episodes, eval_at, n_threads=1, want_tables=True, want_stats=True):
    """Run `count` independent reference agents; returns a dict of numpy arrays."""
    L = lib()
    S, A = env_dims(cfg)
    T = 2 if cfg.policy_kind == POLICY_DOUBLE else 1
    out = {}
    if want_stats:
        out["ret"] = np.zeros((count, n_episodes), np.float64)
        out["len"] = np.zeros((count, n_episodes), np.uint64)
        out["tdsum"] = np.zeros((count, n_episodes), np.float64)
        out["tdabs"] = np.zeros((count, n_episodes), np.float64)
    if want_tables:
        out["q"] = np.zeros((count, T, S, A), np.float64)
        out["counts"] = np.zeros((count, S, A), np.uint64)
    states = np.zeros(count, STATE_DTYPE)
    secs, ts, es = C.c_double(), C.c_uint64(), C.c_uint64()
    rc = L.oracle_batch_train(C.byref(cfg), first_agent, count, n_episodes, eval_at, n_threads, _p(out.get("ret")),
                              _p(out.get("len")), _p(out.get("tdsum")), _p(out.get("tdabs")), _p(out.get("q")),
                              _p(out.get("counts")), _p(states), C.byref(secs), C.byref(ts), C.byref(es))
    if rc:
        raise RuntimeError("oracle_batch_train rc=%d" % rc)
    out["state"] = states
    out["seconds"] = secs.value
    out["train_steps"] = ts.value
    out["eval_steps"] = es.value
    return out
