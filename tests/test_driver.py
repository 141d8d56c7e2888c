"""Host driver ("next" rows N1/N2): moving_average with the reference's short-last-chunk quirk (CPU), and the
12-run CLI matrix on one env + one RNG stream against the oracle driven the same way (GPU)."""
import importlib

import numpy as np
import pytest

from oracle import oracle_py as O


def sequential_moving_average(window, v):
    """utils.rs:78-93 restated with plain Python floats: left-to-right sums, the short last slice divided by the window."""
    out, aux = [], 0
    while window > 0 and aux < len(v):
        end = aux + window if aux + window < len(v) else len(v)
        acc = 0.0
        for x in v[aux:end]:
            acc += float(x)
        out.append(acc / float(window))
        aux = end
    return out


def P_bits_equal(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and bool(np.all((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))))


def LEGENDS_I(res, i):
    return "%s (%s)" % (res["legends"][i], res.get("train_errors_kind"))


def test_moving_average_is_sequential():
    drv = importlib.import_module("rl-rust_b200.driver")
    rng = np.random.default_rng(5)
    v = rng.standard_normal(10007) * 10.0 ** rng.integers(-8, 8, 10007)
    for w in (1, 7, 100, 4096, 10007, 20000):
        assert P_bits_equal(drv.moving_average(w, v), sequential_moving_average(w, v))
    assert drv.moving_average(0, v) == []      # the reference would loop forever on a zero window (utils.rs:81-90)


def test_training_error_curve_forms():
    """bin/taxi.rs:170-174 on the per-step stream (one agent: exact; several: mean of each agent's own curve), and the
    per-episode fallback when the stream was not taken."""
    drv = importlib.import_module("rl-rust_b200.driver")
    rng = np.random.default_rng(9)
    td = rng.standard_normal((2, 57))
    one = dict(td_steps=td[:1], td_count=np.array([50], np.uint64), sums=np.zeros((5, 4)))
    curve, kind = drv.training_error_curve(one, 1, 10, 1)
    assert kind.startswith("per-step (exact") and P_bits_equal(curve, sequential_moving_average(50 // 10, td[0, :50]))
    two = dict(td_steps=td, td_count=np.array([50, 57], np.uint64), sums=np.zeros((5, 4)))
    curve, kind = drv.training_error_curve(two, 2, 10, 1)
    a, b = sequential_moving_average(5, td[0, :50]), sequential_moving_average(5, td[1, :57])
    assert len(curve) == max(len(a), len(b)) == 12 and curve[0] == (a[0] + b[0]) / 2 and curve[-1] == b[-1]
    sums = np.array([[10.0, 0, 5.0, 0], [20.0, 0, -4.0, 0], [0.0, 0, 0.0, 0]])
    curve, kind = drv.training_error_curve(dict(sums=sums), 3, 10, 2)
    assert kind.startswith("per-episode") and curve[0] == (0.5 - 0.2) / 2 and np.isnan(curve[1])
    overflow = dict(td_steps=td[:1, :20], td_count=np.array([50], np.uint64), sums=sums)   # stream cut short: fall back
    assert drv.training_error_curve(overflow, 1, 10, 2)[1].startswith("per-episode")


def test_moving_average_quirk():
    drv = importlib.import_module("rl-rust_b200.driver")
    v = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0]
    assert drv.moving_average(3, v) == [2.0, 5.0, 7.0 / 3.0]          # utils.rs:78-93: the short last slice is divided by the full window
    assert drv.moving_average(7, v) == [4.0]
    assert drv.moving_average(8, v) == [28.0 / 8.0]
    assert drv.moving_average(2, [1.0, 3.0, 5.0, 7.0]) == [2.0, 6.0]
    assert drv.moving_average(3, []) == []


@pytest.mark.gpu
@pytest.mark.parametrize("env_name,env_kind", [("taxi", O.ENV_TAXI), ("frozen_lake", O.ENV_FROZEN_LAKE), ("blackjack", O.ENV_BLACKJACK),
                                               ("cliffwalking", O.ENV_CLIFF_WALKING)])
def test_twelve_run_matrix_matches_oracle(env_name, env_kind):
    """bin/taxi.rs:158-203 with n_agents = 1: every curve of every one of the 12 runs equals the oracle's, i.e. the two
    agent objects really share one env and one stream."""
    drv = importlib.import_module("rl-rust_b200.driver")
    n, seed = 40, 0xFACE
    flags = dict(n_episodes=n, moving_average_window=10, stochastic_env=True, map="8x8")
    res = drv.run_experiment(env_name, n_agents=1, seed=seed, real="f64", tally_games=300, verbose=False, **flags)
    cfg = O.make_config(env_kind, map_id=1, slippery=1, target=O.TARGET_SARSA, eps_decay=1.0 / (0.5 * n), seed=seed)
    s = O.Session(cfg, 0)
    window = n // 10
    i = 0
    for kind in (0, 1):
        s.set_agent_kind(kind)
        for sel in (0, 1):
            s.set_selector(sel)
            for tgt in (0, 1, 2):
                s.set_target(tgt)
                ret, ln, tds, _ = s.train(n, n // 10)
                assert res["train_steps"][i] == int(ln.sum())
                assert res["train_rewards"][i] == drv.moving_average(window, ret)
                assert res["train_episodes_length"][i] == drv.moving_average(window, ln.astype(np.float64))
                # "Training Error": windows of training_error.len() / moving_average_window raw per-step TDs (bin/taxi.rs:170-174)
                te = s.training_error()
                assert len(te) == int(ln.sum())
                assert P_bits_equal(res["train_errors"][i], sequential_moving_average(len(te) // 10, te)), LEGENDS_I(res, i)
                if env_name == "blackjack":
                    eret, _ = s.evaluate(300)
                    assert res["blackjack_rates"][i] == (float((eret == 1).sum()) / 300, float((eret == -1).sum()) / 300,
                                                         float(((eret != 1) & (eret != -1)).sum()) / 300)
                eret, eln = s.evaluate(n)
                assert res["test_rewards"][i] == drv.moving_average(window, eret)
                assert res["test_episodes_length"][i] == drv.moving_average(window, eln.astype(np.float64))
                s.agent_reset()
                i += 1
    assert res["final_rng_n"][0] == s.export()[2].rng_n
    s.close()


@pytest.mark.gpu
def test_cliffwalking_model_bin_matches_oracle():
    """bin/cliffwalking_model.rs:158-203 with n_agents = 1: Q-learning, then Dyna-Q (10 planning steps) around a second
    fresh agent, on one env and one stream."""
    drv = importlib.import_module("rl-rust_b200.driver")
    n, seed = 60, 0xD17A
    res = drv.run_model_experiment(n_agents=1, seed=seed, real="f64", verbose=False, n_episodes=n, moving_average_window=10)
    cfg = O.make_config(O.ENV_CLIFF_WALKING, target=O.TARGET_QLEARNING, eps_decay=1.0 / (0.5 * n), seed=seed)
    s = O.Session(cfg, 0)
    window = n // 10
    for i, planning in enumerate((0, 10)):
        s.set_agent_kind(0)
        s.set_planning(planning)
        s.set_selector(0)
        ret, ln, tds, _ = s.train(n, n // 10)
        assert res["train_steps"][i] == int(ln.sum())
        assert res["train_rewards"][i] == drv.moving_average(window, ret)
        assert res["train_episodes_length"][i] == drv.moving_average(window, ln.astype(np.float64))
        te = s.training_error()
        assert P_bits_equal(res["train_errors"][i], sequential_moving_average(len(te) // 10, te))
        eret, eln = s.evaluate(n)
        assert res["test_rewards"][i] == drv.moving_average(window, eret)
        assert res["test_episodes_length"][i] == drv.moving_average(window, eln.astype(np.float64))
        s.agent_reset()
    assert res["final_rng_n"][0] == s.export()[2].rng_n
    assert sum(res["train_episodes_length"][1]) < sum(res["train_episodes_length"][0])   # planning pays
    s.close()


@pytest.mark.gpu
def test_driver_multi_gpu_matches_single_gpu(rlb):
    """`--gpus 2`: the agents sharded over two GPUs (one engine each, asynchronous train calls, NCCL gather of the
    per-episode sums) give the curves of the same agents on one GPU — per-agent results do not depend on the sharding."""
    if rlb.abi.lib.rlb_device_count() < 2:
        pytest.skip("needs two GPUs")
    drv = importlib.import_module("rl-rust_b200.driver")
    kw = dict(n_agents=64, seed=0xBEEF, real="f64", tally_games=0, verbose=False, n_episodes=20, moving_average_window=10)
    one = drv.run_experiment("taxi", gpus=1, **kw)
    two = drv.run_experiment("taxi", gpus=2, **kw)
    for k in ("train_rewards", "train_episodes_length", "test_rewards", "test_episodes_length", "train_steps"):
        assert one[k] == two[k], k
    for a, b in zip(one["train_errors"], two["train_errors"]):
        assert P_bits_equal(a, b)
