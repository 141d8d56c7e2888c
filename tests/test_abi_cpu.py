"""CPU tests of the product library's host side (no GPU, no compute calls): the C-ABI library
loads and exports every symbol include/rlb.h declares, its host-callable RNG contract and Blackjack
id bijection agree with the oracle, compute entry points fail loudly without a device, and the
host-side mirror of the reference interface has the reference's names."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import oracle_py as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(rlb):
    hdr = open(os.path.join(ROOT, "include", "rlb.h")).read()
    declared = set(re.findall(r"^(?:rlb_status|void|int|int32_t|const char\*|uint32_t|uint64_t|double)\s+(rlb_[a-z0-9_]+)\(", hdr, re.M))
    assert len(declared) >= 40
    assert declared == set(rlb.abi.EXPORTS), declared ^ set(rlb.abi.EXPORTS)
    for sym in declared:
        assert hasattr(rlb.abi.lib, sym), sym
    assert rlb.abi.lib.rlb_abi_version() == 2


def test_struct_layouts_match_header(rlb):
    # ask the C compiler what include/rlb.h says
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "rlb.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu", sizeof(rlb_config), ' \
          'sizeof(rlb_train_out), sizeof(rlb_episode_f32), sizeof(rlb_episode_f64), sizeof(rlb_traj_record), sizeof(rlb_agent_state));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(td, "s"), os.path.join(td, "s.c")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(td, "s")]).split()]
    assert sizes == [C.sizeof(rlb.abi.RlbConfig), C.sizeof(rlb.abi.RlbTrainOut), rlb.abi.EPISODE_F32.itemsize,
                     rlb.abi.EPISODE_F64.itemsize, rlb.abi.TRAJ_DTYPE.itemsize, rlb.abi.STATE_DTYPE.itemsize]
    assert rlb.abi.EPISODE_F32.itemsize == 16 and rlb.abi.EPISODE_F64.itemsize == 32
    assert rlb.abi.TRAJ_DTYPE.itemsize == 24 and O.TRAJ_DTYPE.itemsize == 24
    assert rlb.abi.STATE_DTYPE.itemsize == 32


def test_host_rng_contract_matches_oracle(rlb):
    L = O.lib()
    assert list(rlb.abi.philox4x32_10([0] * 4, [0] * 2)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    for seed, agent, start in ((0, 0, 0), (0x5EED0001, 123456789012, 5), (2 ** 64 - 1, 2 ** 64 - 1, 2 ** 40 + 3)):
        w = np.zeros(50, np.uint32)
        L.oracle_stream_words(seed, agent, start, 50, O._p(w))
        assert np.array_equal(rlb.abi.rng_words(seed, agent, start, 50), w)
    lib = rlb.abi.lib
    for kind, rng_range in ((0, 0), (1, 2), (1, 4), (1, 6), (2, 0)):
        ref = np.zeros(300, np.float64)
        used = L.oracle_sample(77, 9, 3, kind, rng_range, 300, O._p(ref))
        n = C.c_uint64(3)
        got = []
        for _ in range(300):
            if kind == 0:
                got.append(lib.rlb_rng_uniform_f64(77, 9, C.byref(n)))
            elif kind == 1:
                got.append(float(lib.rlb_rng_uniform_usize(77, 9, C.byref(n), rng_range)))
            else:
                got.append(float(lib.rlb_rng_card(77, 9, C.byref(n))))
        assert got == list(ref) and n.value - 3 == used


def test_blackjack_id_bijection(rlb):
    L = O.lib()
    p, d, a = C.c_uint32(), C.c_uint32(), C.c_uint32()
    seen = set()
    for dense in range(1456):
        rlb.abi.lib.rlb_blackjack_decode(dense, C.byref(p), C.byref(d), C.byref(a))
        assert 4 <= p.value <= 31 and 1 <= d.value <= 26 and a.value in (0, 1)
        assert L.oracle_blackjack_dense(p.value, d.value, a.value) == dense
        oid = rlb.abi.blackjack_obs_id(dense)
        assert oid == L.oracle_fxhash_blackjack(p.value, d.value, a.value)
        seen.add(oid)
    assert len(seen) == 1456
    assert rlb.abi.blackjack_dense_index(14677233788820433653) == L.oracle_blackjack_dense(12, 1, 1)
    assert rlb.abi.blackjack_dense_index(12345) == 0xffffffff


def test_no_cpu_fallback(rlb):
    """Without a CUDA device the engine refuses to exist; with one this test is vacuous."""
    if rlb.abi.lib.rlb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(rlb.RlbError) as ei:
        rlb.Engine(rlb.abi.ENV_TAXI, n_agents=4)
    assert ei.value.status == rlb.abi.ERR_CUDA
    assert "no CPU path" in str(ei.value)


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under rl-rust_b200/ or include/ may reference it."""
    for base in ("rl-rust_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                    txt = open(os.path.join(dirpath, f), errors="replace").read()
                    assert "oracle_py" not in txt and "liboracle" not in txt and "oracle.hpp" not in txt and "oracle_capi" not in txt, \
                        "%s references the oracle" % os.path.join(dirpath, f)


def test_mirror_has_the_reference_names(rlb):
    for name in ("BlackJackEnv", "FrozenLakeEnv", "CliffWalkingEnv", "TaxiEnv", "TabularPolicy", "DoubleTabularPolicy",
                 "UniformEpsilonGreed", "UpperConfidenceBound", "OneStepAgent", "ElegibilityTracesAgent", "sarsa", "qlearning",
                 "expected_sarsa", "EnvNotReady"):
        assert hasattr(rlb, name), name
    env = rlb.TaxiEnv(100)
    assert env.action_size() == 6 and env.get_action_label(4) == "PICKUP" and rlb.TaxiEnv.decode(499) == (4, 4, 4, 3)
    assert rlb.FrozenLakeEnv(rlb.FrozenLakeEnv.MAP_8X8, True, 100).action_size() == 4
    agent = rlb.OneStepAgent(rlb.TabularPolicy(0.05, 0.0), 0.95, rlb.UniformEpsilonGreed(1.0, ("sub", 2e-5), 0.0), rlb.qlearning)
    with pytest.raises(ZeroDivisionError):
        agent.train(env, 10, 0)                               # agent.rs:107 `episode % eval_at` with eval_at == 0 panics
    for m in ("set_future_q_value_func", "set_action_selector", "get_action", "update", "reset", "train", "evaluate"):
        assert callable(getattr(agent, m))


def test_comm_entry_points_without_a_gpu(rlb):
    """rlb_comm_*: NCCL is resolved at run time (dlopen); the unique id needs no device, a communicator does — and says so."""
    uid = rlb.abi.Comm.unique_id()
    assert len(uid) == rlb.abi.COMM_ID_BYTES and uid != bytes(128)
    if rlb.abi.lib.rlb_device_count() == 0:
        with pytest.raises(rlb.RlbError) as ei:
            rlb.abi.Comm.init_rank(uid, 1, 0, 0)
        assert ei.value.status in (rlb.abi.ERR_CUDA, rlb.abi.ERR_NCCL)
        with pytest.raises(rlb.RlbError):
            rlb.abi.Comm.init_all([0])
    with pytest.raises(rlb.RlbError) as ei:
        rlb.abi.Comm.init_rank(uid, 2, 5, 0)          # rank out of range
    assert ei.value.status == rlb.abi.ERR_INVALID_ARG
    assert rlb.abi.lib.rlb_comm_rank(None) == -1 and rlb.abi.lib.rlb_comm_world_size(None) == 0


def test_shard_sizes(rlb):
    import importlib
    sh = importlib.import_module("rl-rust_b200.sharding")
    for total, n in ((1 << 24, 8), (1 << 24, 2), (10, 3), (7, 7)):
        parts = sh.shard_sizes(total, n)
        assert sum(c for _, c in parts) == total and parts[0][0] == 0
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(n - 1))
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
