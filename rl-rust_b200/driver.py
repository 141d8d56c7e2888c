"""Host experiment driver — the reference's per-env bins over the engine ("next" row N1/N2 of SURVEY.md §8f).

`run_model_experiment()` is `main()` of src/bin/cliffwalking_model.rs (Q-learning against Dyna-Q).
`run_experiment()` is `main()` of src/bin/{taxi,frozen_lake,cliffwalking,blackjack}.rs with the same flags and
defaults (bin/taxi.rs:22-68): two agents (OneStepAgent, ElegibilityTracesAgent) x two selectors (eps-greedy, UCB) x
three bootstrap targets = 12 runs, each `train(n, n/10)` -> `evaluate(n)` -> `agent.reset()` (bin/taxi.rs:158-203), the
env shared and never re-created, Blackjack followed by its win / loss / draw tally (bin/blackjack.rs:179-207).  The
curves are `moving_average` (utils.rs:78-93, quirk included) of the per-episode MEAN over the batch's agents; they are
returned / written as JSON and, with `--plots DIR`, drawn as the bins' five PNG charts (charts.py; utils.rs:97-157).

With n_agents = 1 every curve is the reference's, value for value (same Philox stream) — including "Training Error",
which the bins window over the raw per-STEP temporal differences with a window of `training_error.len() / window`
(bin/taxi.rs:170-174): the engine streams them (rlb_train_out.td_steps) whenever n_agents x n_episodes x (max_steps+1)
values fit TD_STREAM_BYTES; with more agents each agent's own curve is formed that way and the chart shows their mean.
Past that budget the curve falls back to windows over episodes (mean TD per step of each episode) and says so in
`train_errors_kind`.

`gpus=G` (CLI `--gpus G`) shards the agents over G GPUs of the box by global agent id — one engine per device, driven
from this one process through the asynchronous train call — and gathers the per-episode sums to GPU 0 with the library's
NCCL gather (rlb_comm_init_all / rlb_comm_gather_episode_sums).
"""
import json
import time

import numpy as np

from . import _abi as abi
from . import api

LEGENDS = ["ε-Greedy One-Step Sarsa", "ε-Greedy One-Step Qlearning", "ε-Greedy One-Step Expected Sarsa", "UCB One-Step Sarsa",
           "UCB One-Step Qlearning", "UCB One-Step Expected Sarsa", "ε-Greedy Trace Sarsa", "ε-Greedy Trace Qlearning",
           "ε-Greedy Trace Expected Sarsa", "UCB Trace Sarsa", "UCB Trace Qlearning", "UCB Trace Expected Sarsa"]   # bin/taxi.rs:96-110

DEFAULTS = dict(n_episodes=100000, max_steps=100, learning_rate=0.05, initial_epsilon=1.0, exploration_time=0.5, final_epsilon=0.0,
                confidence_level=0.5, discount_factor=0.95, lambda_factor=0.5, moving_average_window=100, stochastic_env=False,
                map="4x4")


def moving_average(window, vector):
    """utils.rs:78-93: sums of consecutive `window`-long slices, each divided by `window` — including the last,
    possibly shorter, slice (the reference's quirk); when len % window == 0 there is no short slice.  The sums are
    sequential left-to-right f64 additions like `slice.iter().sum()` (np.cumsum; np.sum would add pairwise)."""
    vector = np.asarray(vector, np.float64)
    out = []
    aux = 0
    if window <= 0:
        return out
    while aux < len(vector):
        end = aux + window if aux + window < len(vector) else len(vector)
        out.append(float(np.cumsum(vector[aux:end])[-1]) / float(window))
        aux = end
    return out


TD_STREAM_BYTES = 1 << 31   # host + device budget for the exact per-step TD stream of one train() call


def training_error_curve(res, n_agents, ma_window, window_episodes):
    """The bins' "Training Error" series (bin/taxi.rs:170-174): moving_average(training_error.len() / ma_window,
    &training_error) per agent from the per-step TD stream, averaged over agents point by point (an agent contributes
    to the points it has); without the stream, windows over episodes of the mean TD per step."""
    if "td_steps" in res and int(res["td_count"].max()) <= res["td_steps"].shape[1]:
        curves = []
        for i in range(n_agents):
            n = int(res["td_count"][i])
            curves.append(moving_average(n // ma_window, res["td_steps"][i, :n].astype(np.float64)))
        if n_agents == 1:
            return curves[0], "per-step (exact: bin/taxi.rs:170-174)"
        width = max(len(c) for c in curves)
        acc, cnt = np.zeros(width), np.zeros(width)
        for c in curves:
            acc[:len(c)] += c
            cnt[:len(c)] += 1
        return (acc / np.maximum(cnt, 1)).tolist(), "per-step, mean over agents of each agent's own curve"
    s_ = res["sums"]
    with np.errstate(invalid="ignore", divide="ignore"):
        mean_td = s_[:, 2] / s_[:, 0]
    return moving_average(window_episodes, mean_td), "per-episode windows of the mean TD per step (the per-step stream exceeds TD_STREAM_BYTES)"


class EngineGroup:
    """The engines of one experiment: one per GPU, agents sharded by global id; a single engine when gpus == 1.
    Forwards the agent / selector / target switches to every engine, trains them concurrently (asynchronous train
    call) and gathers the per-episode sums to GPU 0 with the library's NCCL gather."""

    def __init__(self, gpus, n_agents, **kw):
        self.gpus = gpus
        base, extra = divmod(n_agents, gpus)
        self.sizes = [base + (1 if r < extra else 0) for r in range(gpus)]
        if min(self.sizes) < 1:
            raise ValueError("n_agents must be >= gpus")
        first = 0
        self.engines = []
        dev0 = kw.pop("device", 0)
        for r, n in enumerate(self.sizes):
            self.engines.append(abi.Engine(n_agents=n, first_agent_id=first, device=dev0 + r, **kw))
            first += n
        self.comms = abi.Comm.init_all([dev0 + r for r in range(gpus)]) if gpus > 1 else None
        self.dev0 = dev0

    def __getattr__(self, name):   # set_agent_kind, set_selector, set_target, agent_reset, set_model: applied to every engine
        def call(*a, **k):
            out = [getattr(e, name)(*a, **k) for e in self.engines]
            return out[0]
        return call

    def train(self, n, eval_at, td_capacity=0):
        if self.gpus == 1:
            return self.engines[0].train(n, eval_at, td_capacity=td_capacity)
        import torch
        sums = [torch.zeros((n, 4), dtype=torch.float64, device="cuda:%d" % (self.dev0 + r)) for r in range(self.gpus)]
        gathered = torch.zeros((self.gpus, n, 4), dtype=torch.float64, device="cuda:%d" % self.dev0)
        for r, e in enumerate(self.engines):
            torch.cuda.synchronize(self.dev0 + r)
            e.train(n, eval_at, sums_out=sums[r], td_capacity=td_capacity, wait=False)
        parts = [e.train_wait()[0] for e in self.engines]
        abi.check(abi.lib.rlb_comm_group_begin())
        for r, cm in enumerate(self.comms):
            cm.gather_episode_sums(sums[r], gathered if r == 0 else None, root=0, stream=torch.cuda.current_stream(self.dev0 + r).cuda_stream)
        abi.check(abi.lib.rlb_comm_group_end())
        for r in range(self.gpus):
            torch.cuda.synchronize(self.dev0 + r)
        res = dict(sums=gathered.sum(0).cpu().numpy(), train_steps=sum(p["train_steps"] for p in parts), eval_steps=sum(p["eval_steps"] for p in parts))
        if td_capacity:
            res["td_steps"] = np.concatenate([p["td_steps"] for p in parts], 0)
            res["td_count"] = np.concatenate([p["td_count"] for p in parts], 0)
        return res

    def evaluate(self, n, sums=True, episodes=False):
        parts = [e.evaluate(n, sums=sums, episodes=episodes) for e in self.engines]
        out = dict(steps=sum(p["steps"] for p in parts))
        out["sums"] = sum(p["sums"] for p in parts) if sums else None
        out["episodes"] = np.concatenate([p["episodes"] for p in parts], 1) if episodes else None
        return out

    def get_action(self, obs):
        return self.engines[0].get_action(obs)

    def states(self):
        return np.concatenate([e.states() for e in self.engines])

    def close(self):
        for e in self.engines:
            e.close()
        for c in self.comms or []:
            c.close()


def make_env(name, **flags):
    if name == "blackjack":
        return api.BlackJackEnv()                                                                    # bin/blackjack.rs:78
    if name == "frozen_lake":
        m = api.FrozenLakeEnv.MAP_4X4 if flags.get("map", "4x4") == "4x4" else api.FrozenLakeEnv.MAP_8X8   # bin/frozen_lake.rs:93-97
        return api.FrozenLakeEnv(m, flags.get("stochastic_env", False), flags["max_steps"])
    if name in ("cliffwalking", "cliff_walking"):
        return api.CliffWalkingEnv(flags["max_steps"])
    if name == "taxi":
        return api.TaxiEnv(flags["max_steps"])
    raise ValueError("unknown env %r" % name)


def run_experiment(env_name, *, n_agents=1, seed=0x5EED0001, real="f64", device=0, tally_games=1000000, policy="basic",
                   verbose=True, show_example=False, gpus=1, **flags):
    """Returns {'legends', 'train_rewards', 'train_episodes_length', 'train_errors', 'test_rewards',
    'test_episodes_length', 'seconds', ['blackjack_rates']} — the five chart series of bin/taxi.rs:205-223."""
    f = dict(DEFAULTS)
    f.update(flags)
    n = int(f["n_episodes"])
    epsilon_decay = f["initial_epsilon"] / (f["exploration_time"] * n)                               # bin/taxi.rs:78
    window = max(1, n // int(f["moving_average_window"]))                                            # first argument at every call site
    env = make_env(env_name, **f)
    # ONE engine = the bins' one env + one RNG stream per agent slot; the two agent objects of bin/taxi.rs:138-156 take
    # turns on it (rlb_agent_set_kind), so run 7 continues the stream where run 6 left it, as in the reference.
    eng = EngineGroup(gpus, n_agents, env_kind=env.kind, policy=abi.POLICY_BASIC if policy == "basic" else abi.POLICY_DOUBLE,   # bins: Basic (bin/taxi.rs:126)
                      selector=abi.SEL_EPS_GREEDY, target=abi.TARGET_SARSA, agent=abi.AGENT_ONE_STEP,
                      real=abi.REAL_F32 if real == "f32" else abi.REAL_F64, learning_rate=f["learning_rate"],
                      discount_factor=f["discount_factor"], lambda_factor=f["lambda_factor"], initial_epsilon=f["initial_epsilon"],
                      decay_kind=abi.DECAY_SUB, epsilon_decay=epsilon_decay, final_epsilon=f["final_epsilon"],
                      confidence_level=f["confidence_level"], default_value=0.0, seed=seed, device=device, **env._cfg())
    env.bind(eng.engines[0])
    # per-step TD stream: at most max_steps + 1 steps per episode (Blackjack: a hand holds 16 cards, blackjack.rs:32-35)
    td_cap = n * (32 if env_name == "blackjack" else int(f["max_steps"]) + 1)
    if n_agents * td_cap * 8 > TD_STREAM_BYTES:
        td_cap = 0
    out = dict(legends=LEGENDS, train_rewards=[], train_episodes_length=[], train_errors=[], test_rewards=[],
               test_episodes_length=[], seconds=[], train_steps=[], gpus=gpus)
    if env_name == "blackjack":
        out["blackjack_rates"] = []
    i = 0
    for agent_kind in (abi.AGENT_ONE_STEP, abi.AGENT_TRACES):                                        # bin/taxi.rs:154-159
        eng.set_agent_kind(agent_kind)
        for sel in (abi.SEL_EPS_GREEDY, abi.SEL_UCB):
            eng.set_selector(sel)                                                                    # bin/taxi.rs:161 (a fresh clone)
            for func in (api.sarsa, api.qlearning, api.expected_sarsa):
                eng.set_target(func)                                                                 # bin/taxi.rs:163
                t0 = time.perf_counter()
                res = eng.train(n, max(1, n // 10), td_capacity=td_cap)                              # bin/taxi.rs:165-166
                dt = time.perf_counter() - t0
                if verbose:
                    print("%s %.2fs (%d agents, %.3g train steps/s)" % (LEGENDS[i], dt, n_agents, res["train_steps"] / dt))
                s_ = res["sums"]                                                                     # [n,4]: sum len, ret, td, |td|
                mean_len, mean_ret = s_[:, 0] / n_agents, s_[:, 1] / n_agents
                curve, kind = training_error_curve(res, n_agents, int(f["moving_average_window"]), window)   # bin/taxi.rs:170-174
                out["train_errors"].append(curve)
                out["train_errors_kind"] = kind
                out["train_rewards"].append(moving_average(window, mean_ret))
                out["train_episodes_length"].append(moving_average(window, mean_len))
                out["seconds"].append(dt)
                out["train_steps"].append(int(res["train_steps"]))
                if env_name == "blackjack" and tally_games:                                          # bin/blackjack.rs:179-207
                    ev = eng.evaluate(int(tally_games), sums=False, episodes=True)["episodes"]["ret"]
                    tot = float(ev.size)
                    rates = (float((ev == 1.0).sum()) / tot, float((ev == -1.0).sum()) / tot, float(((ev != 1.0) & (ev != -1.0)).sum()) / tot)
                    out["blackjack_rates"].append(rates)
                    if verbose:
                        print("%s has win-rate of %s%%, loss-rate of %s%% and draw-rate %s%%" % ((LEGENDS[i],) + rates))
                if show_example:                                                                     # bin/taxi.rs:184-186
                    out.setdefault("examples", []).append(api.example_episode(env, eng.get_action, out=print if verbose else (lambda _: None)))
                ev = eng.evaluate(n, sums=True)["sums"]                                              # bin/taxi.rs:188
                out["test_rewards"].append(moving_average(window, ev[:, 1] / n_agents))
                out["test_episodes_length"].append(moving_average(window, ev[:, 0] / n_agents))
                i += 1
                eng.agent_reset()                                                                    # bin/taxi.rs:200
    out["final_rng_n"] = [int(x) for x in eng.states()["rng_n"][:8]]
    eng.close()
    return out


LEGENDS_MODEL = ["ε-Greedy One-Step Qlearning", "ε-Greedy One-Step Dyna-Qlearning"]   # bin/cliffwalking_model.rs:96-111


def run_model_experiment(*, n_agents=1, seed=0x5EED0001, real="f64", device=0, planning_steps=10, verbose=True, **flags):
    """`main()` of src/bin/cliffwalking_model.rs: CliffWalking, a OneStepAgent with qlearning and then an
    InternalModelAgent (RandomModel, 10 planning steps) around a second such agent, eps-greedy, Basic policy; each
    `train(n, n/10)` -> `evaluate(n)` -> `reset()` on the one shared env (:158-203).  Same output dict as run_experiment."""
    f = dict(DEFAULTS)
    f.update(flags)
    n = int(f["n_episodes"])
    epsilon_decay = f["initial_epsilon"] / (f["exploration_time"] * n)                               # :77
    window = max(1, n // int(f["moving_average_window"]))
    env = api.CliffWalkingEnv(f["max_steps"])                                                        # :86
    eng = abi.Engine(env.kind, n_agents=n_agents, policy=abi.POLICY_BASIC, selector=abi.SEL_EPS_GREEDY, target=abi.TARGET_QLEARNING,
                     agent=abi.AGENT_ONE_STEP, real=abi.REAL_F32 if real == "f32" else abi.REAL_F64, learning_rate=f["learning_rate"],
                     discount_factor=f["discount_factor"], lambda_factor=f["lambda_factor"], initial_epsilon=f["initial_epsilon"],
                     decay_kind=abi.DECAY_SUB, epsilon_decay=epsilon_decay, final_epsilon=f["final_epsilon"],
                     confidence_level=f["confidence_level"], default_value=0.0, seed=seed, device=device, **env._cfg())
    env.bind(eng)
    out = dict(legends=LEGENDS_MODEL, train_rewards=[], train_episodes_length=[], train_errors=[], test_rewards=[],
               test_episodes_length=[], seconds=[], train_steps=[])
    for i, planning in enumerate((0, int(planning_steps))):                                          # :158-160
        eng.set_agent_kind(abi.AGENT_ONE_STEP)                                                       # `one_step_agent` / `other` (:136-148): a fresh agent
        if planning:
            eng.set_model(planning)                                                                  # :150-156
        eng.set_selector(abi.SEL_EPS_GREEDY)                                                         # :164
        t0 = time.perf_counter()
        td_cap = n * (int(f["max_steps"]) + 1)
        res = eng.train(n, max(1, n // 10), td_capacity=td_cap if n_agents * td_cap * 8 <= TD_STREAM_BYTES else 0)   # :166-167
        dt = time.perf_counter() - t0
        if verbose:
            print("%s %.2fs (%d agents, %.3g train steps/s)" % (LEGENDS_MODEL[i], dt, n_agents, res["train_steps"] / dt))
        s_ = res["sums"]
        curve, kind = training_error_curve(res, n_agents, int(f["moving_average_window"]), window)   # :171-175; the model agent returns the wrapped agent's TD (internal_model_agent.rs:78)
        out["train_errors"].append(curve)
        out["train_errors_kind"] = kind
        out["train_rewards"].append(moving_average(window, s_[:, 1] / n_agents))
        out["train_episodes_length"].append(moving_average(window, s_[:, 0] / n_agents))
        out["seconds"].append(dt)
        out["train_steps"].append(int(res["train_steps"]))
        ev = eng.evaluate(n, sums=True)["sums"]                                                      # :189
        out["test_rewards"].append(moving_average(window, ev[:, 1] / n_agents))
        out["test_episodes_length"].append(moving_average(window, ev[:, 0] / n_agents))
        eng.agent_reset()                                                                            # :201
    out["final_rng_n"] = [int(x) for x in eng.states()["rng_n"][:8]]
    eng.close()
    return out


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="RL-Rust bins over the B200 engine: blackjack | frozen_lake | cliffwalking | taxi | cliffwalking_model")
    ap.add_argument("env", choices=["blackjack", "frozen_lake", "cliffwalking", "taxi", "cliffwalking_model"])
    ap.add_argument("--n_episodes", "-n", type=int, default=DEFAULTS["n_episodes"])
    for name in ("max_steps", "moving_average_window"):
        ap.add_argument("--" + name, type=int, default=DEFAULTS[name])
    for name in ("learning_rate", "initial_epsilon", "exploration_time", "final_epsilon", "confidence_level", "discount_factor",
                 "lambda_factor"):
        ap.add_argument("--" + name, type=float, default=DEFAULTS[name])
    ap.add_argument("--stochastic_env", action="store_true")
    ap.add_argument("--map", default="4x4")
    ap.add_argument("--show_example", action="store_true", help="print one rendered episode after each training run (needs --n_agents 1)")
    ap.add_argument("--n_agents", type=int, default=1)
    ap.add_argument("--gpus", type=int, default=1, help="shard the agents over this many GPUs of the box (NCCL gather of the curves)")
    ap.add_argument("--seed", type=lambda x: int(x, 0), default=0x5EED0001)
    ap.add_argument("--real", choices=["f32", "f64"], default="f64")
    ap.add_argument("--tally_games", type=int, default=1000000)
    ap.add_argument("--out", default=None, help="write the chart series as JSON here")
    ap.add_argument("--plots", default=None, help="directory for the five PNG charts of the bins (utils.rs:97-157)")
    a = vars(ap.parse_args(argv))
    env_name, outp, plots = a.pop("env"), a.pop("out"), a.pop("plots")
    if env_name == "cliffwalking_model":
        for k in ("tally_games", "stochastic_env", "map", "show_example", "gpus"):
            a.pop(k)
        res = run_model_experiment(**a)
    else:
        res = run_experiment(env_name, **a)
    if outp:
        with open(outp, "w") as fh:
            json.dump(res, fh)
    if plots:
        from . import charts
        charts.plot_experiment(res, plots)
    return res
