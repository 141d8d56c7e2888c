"""`Agent::example` (agent.rs:143-163) and `Env::render` through the mirror on the GPU, against a transcript built from the
CPU oracle driven through the same calls."""
import importlib

import numpy as np
import pytest

from oracle import oracle_py as O

pytestmark = pytest.mark.gpu
R = importlib.import_module("rl-rust_b200.render")
SEED = 0xE8A3


def oracle_transcript(cfg, env_kind, view, label, hands=None):
    """The reference's example() over an oracle Session: same call order, views from the oracle's own observations."""
    s = O.Session(cfg, 0)

    def rng_n():
        return int(s.export()[2].rng_n)
    n0 = rng_n()
    obs = s.env_reset()
    if hands is not None:
        c = hands(n0, rng_n())
        player, dealer = [c[0], c[1]], [c[2], c[3]]
    pos, nsteps, ready = obs, 0, True
    action = s.get_action(obs)

    def transitions():
        nonlocal pos, nsteps, ready, action
        while True:
            before = view(pos, ready, *( (dealer, player) if hands is not None else ()))
            n0 = rng_n()
            truncated = env_kind != O.ENV_BLACKJACK and nsteps >= 100
            obs2, rew, term = s.env_step(action)
            if hands is not None:
                (player if action == 0 else dealer).extend(hands(n0, rng_n()))
            if not truncated:
                pos, nsteps = obs2, nsteps + 1
            ready = not term
            shown = action
            action = s.get_action(obs2)
            yield before, shown, rew, term, view(pos, ready, *((dealer, player) if hands is not None else ())) if term else None
            if term:
                return
    lines = R.example_lines(label, transitions())
    s.close()
    return lines


@pytest.mark.parametrize("name", ["taxi", "frozen_lake", "cliff_walking", "blackjack"])
def test_example_transcript_equals_the_oracles(rlb, name):
    n_train = 30
    decay = 1.0 / (0.5 * n_train)
    sel = rlb.UniformEpsilonGreed(1.0, ("sub", decay), 0.0)
    agent = rlb.OneStepAgent(rlb.TabularPolicy(0.05, 0.0), 0.95, sel, rlb.qlearning, n_agents=1, seed=SEED, real="f64")
    if name == "taxi":
        env, kind, view = rlb.TaxiEnv(100), O.ENV_TAXI, lambda pos, ready: R.render_taxi(pos)
    elif name == "frozen_lake":
        env, kind = rlb.FrozenLakeEnv(rlb.FrozenLakeEnv.MAP_8X8, True, 100), O.ENV_FROZEN_LAKE
        view = lambda pos, ready: R.render_frozen_lake(rlb.FrozenLakeEnv.MAP_8X8, pos)
    elif name == "cliff_walking":
        env, kind, view = rlb.CliffWalkingEnv(100), O.ENV_CLIFF_WALKING, lambda pos, ready: R.render_cliff_walking(pos)
    else:
        env, kind = rlb.BlackJackEnv(), O.ENV_BLACKJACK
        view = lambda pos, ready, dealer, player: R.render_blackjack(ready, dealer, player)
    kw = dict(target=O.TARGET_QLEARNING, eps_decay=decay, seed=SEED)
    if name == "frozen_lake":
        kw.update(map_id=1, slippery=True)
    cfg = O.make_config(kind, **kw)
    hands = None
    if name == "blackjack":
        def hands(n0, n1):   # an independent decode of the stream words the env call consumed (rand 0.8.5 Uniform<u8>(1..11))
            out = []
            for w in rlb.abi.rng_words(SEED, 0, n0, n1 - n0):
                m = int(w) * 10
                if (m & 0xffffffff) <= 0xfffffff9:
                    out.append(1 + (m >> 32))
            return out
    # untrained agent: epsilon = 1, a random walk (long for the grid envs, exercising truncation for Taxi)
    printed = []
    lines = agent.example(env, out=printed.append)
    want = oracle_transcript(cfg, kind, view, env.get_action_label, hands)
    assert lines == printed == want
    assert lines[-1].startswith("terminated with ") and lines[-2].startswith("episode reward ")
    # and after some training, from the stream position the example left (the bins call it between train and evaluate)
    agent.train(env, n_train, 10)
    lines2 = agent.example(env, out=lambda _: None)
    assert len(lines2) >= 6 and lines2[1].strip('"') in env.ACTIONS


def test_example_needs_the_single_env(rlb):
    sel = rlb.UniformEpsilonGreed(1.0, ("sub", 0.1), 0.0)
    agent = rlb.OneStepAgent(rlb.TabularPolicy(0.05, 0.0), 0.95, sel, rlb.qlearning, n_agents=2, seed=SEED)
    with pytest.raises(RuntimeError):
        agent.example(rlb.CliffWalkingEnv(100))


def test_driver_show_example_and_plots(rlb, tmp_path):
    """`taxi --show_example` prints an episode after each of the 12 training runs (bin/taxi.rs:184-186); `--plots` writes
    the five charts (bin/taxi.rs:205-223)."""
    drv = importlib.import_module("rl-rust_b200.driver")
    res = drv.main(["cliffwalking", "--n_episodes", "20", "--show_example", "--plots", str(tmp_path), "--seed", "7"])
    assert len(res["examples"]) == 12 and all(e[-1].startswith("terminated with ") for e in res["examples"])
    assert sorted(p.name for p in tmp_path.iterdir()) == sorted("%s.png" % t for _, t in importlib.import_module("rl-rust_b200.charts").TITLES)
