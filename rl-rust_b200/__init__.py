"""rl-rust_b200 — B200-native batched tabular-RL engine behind the trait surface of
JohnVithor/RL-Rust's hot path (Env / Agent / Policy / ActionSelection).

The package holds only what that path needs: `csrc/` (hand-written sm_100a kernels + the
C ABI of include/rlb.h, built into librlb.so), `_abi.py` (ctypes binding) and `api.py`
(host-side mirror of the reference's interface), plus the thin callers either side of the path: `driver.py` (the
bins), `render.py` (`Env::render` / `Agent::example`), `charts.py` (the bins' PNG charts), `snapshot.py`.  Import it with
`importlib.import_module("rl-rust_b200")` (the directory name is not a Python identifier).
"""
from . import _abi as abi
from ._abi import Engine, EnvNotReady, RlbError
from .snapshot import load_snapshot, save_snapshot
from .api import (BlackJackEnv, CliffWalkingEnv, DoubleTabularPolicy, ElegibilityTracesAgent, Env, FrozenLakeEnv,
                  InternalModelAgent, OneStepAgent, RandomModel, TabularPolicy, TaxiEnv, UniformEpsilonGreed, UpperConfidenceBound, expected_sarsa,
                  example_episode, qlearning, sarsa)

__all__ = ["abi", "Engine", "save_snapshot", "load_snapshot", "EnvNotReady", "RlbError", "BlackJackEnv", "CliffWalkingEnv", "DoubleTabularPolicy",
           "ElegibilityTracesAgent", "Env", "FrozenLakeEnv", "InternalModelAgent", "OneStepAgent", "RandomModel", "TabularPolicy", "TaxiEnv",
           "UniformEpsilonGreed", "UpperConfidenceBound", "expected_sarsa", "qlearning", "sarsa", "example_episode"]
