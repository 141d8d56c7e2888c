"""The reference CLI's defaults and the BASELINE.json configurations as engine arguments.

Product-side helpers shared by bench.py, the driver and the tests: nothing here touches oracle/.
"""
from . import _abi as abi

ENV_NAMES = {0: "blackjack", 1: "frozen_lake", 2: "cliff_walking", 3: "taxi"}
TARGET_NAMES = {0: "sarsa", 1: "qlearning", 2: "expected_sarsa"}
A_OF_ENV = {0: 2, 1: 4, 2: 4, 3: 6}


def hyper(n_episodes, **over):
    """The reference CLI's defaults (bin/taxi.rs:22-68) with the decay derived from n_episodes (:78)."""
    h = dict(map_id=1, slippery=True, max_steps=100, lr=0.05, gamma=0.95, lambda_=0.5, eps0=1.0,
             eps_decay=1.0 / (0.5 * n_episodes), eps_final=0.0, ucb_c=0.5, default_q=0.0, decay_kind=0, seed=0x5EED0001,
             planning_steps=0, map_rows=None)
    h.update(over)
    return h


def combo_id(c):
    return "%s-%s-%s-%s-%s-%s" % (ENV_NAMES[c["env"]], "traces" if c["agent"] else "onestep",
                                  "ucb" if c["selector"] else "eps", "double" if c["policy"] else "basic",
                                  TARGET_NAMES[c["target"]], "f64" if c["real"] else "f32")


def make_engine(c, h, n_agents, first_agent_id=0, **kw):
    """An Engine for the combination `c` = {env, agent, selector, policy, target, real} with hyper-parameters `h`."""
    return abi.Engine(c["env"], n_agents=n_agents, map_id=h["map_id"], slippery=h["slippery"], max_steps=h["max_steps"],
                      policy=c["policy"], selector=c["selector"], target=c["target"], agent=c["agent"], real=c["real"],
                      decay_kind=h["decay_kind"], learning_rate=h["lr"], discount_factor=h["gamma"],
                      lambda_factor=h["lambda_"], initial_epsilon=h["eps0"], epsilon_decay=h["eps_decay"],
                      final_epsilon=h["eps_final"], confidence_level=h["ucb_c"], default_value=h["default_q"],
                      seed=h["seed"], first_agent_id=first_agent_id, planning_steps=h.get("planning_steps", 0),
                      map_rows=h.get("map_rows"), **kw)


# BASELINE.json configs: C1 .. C5, plus the crate's Dyna bin.  `agents_per_gpu` is the single-GPU size;
# `agents_total` (C4) is the job size that is SHARDED over the GPUs of a multi-GPU run (strong scaling).
WORKLOADS = {
    "c1": dict(desc="Blackjack one-step Q-learning, eps-greedy, Basic", env=0, agent=0, selector=0, policy=0, target=1,
               agents_per_gpu=1 << 22, n_episodes=1000, chunk=100),
    "c2": dict(desc="FrozenLake 8x8 slippery, Sarsa(lambda) eligibility traces, eps-greedy, Basic", env=1, agent=1,
               selector=0, policy=0, target=0, agents_per_gpu=1 << 20, n_episodes=1000, chunk=100, slippery=True),
    "c3": dict(desc="CliffWalking Expected Sarsa, Double policy, UCB", env=2, agent=0, selector=1, policy=1, target=2,
               agents_per_gpu=1 << 22, n_episodes=200, chunk=20),
    "c4": dict(desc="Taxi one-step Q-learning, eps-greedy, Basic", env=3, agent=0, selector=0, policy=0, target=1,
               agents_per_gpu=1 << 21, agents_total=1 << 24, n_episodes=1000, chunk=100),
    # not a BASELINE config: the crate's Dyna bin (src/bin/cliffwalking_model.rs), "next" row N4 of SURVEY.md §8(f)
    "dyna": dict(desc="CliffWalking one-step Dyna-Q (InternalModelAgent, RandomModel, 10 planning steps), eps-greedy, Basic",
                 env=2, agent=0, selector=0, policy=0, target=1, planning_steps=10, agents_per_gpu=1 << 20, n_episodes=200, chunk=20),
    # C5, the full sweep: 4 envs x {Sarsa, Q, Expected Sarsa one-step; Sarsa(lambda), Q(lambda)} x {eps-greedy, UCB} x
    # {Basic, Double} = 80 cells, every cell an engine of its own on every GPU, one metric gather per cell per step.
    "c5": dict(desc="full sweep: 4 envs x 5 update rules x {eps-greedy, UCB} x {Basic, Double} = 80 cells", cells=[
        dict(env=env, agent=agent, target=target, selector=sel, policy=pol, slippery=True)
        for env in (0, 1, 2, 3) for (agent, target) in ((0, 0), (0, 1), (0, 2), (1, 0), (1, 1)) for sel in (0, 1) for pol in (0, 1)],
        agents_per_gpu=8192, n_episodes=1000, chunk=100),
}


def workload_hyper(w):
    return hyper(w["n_episodes"], slippery=w.get("slippery", False), planning_steps=w.get("planning_steps", 0))


def combo(w, real):
    return dict(env=w["env"], agent=w["agent"], selector=w["selector"], policy=w["policy"], target=w["target"], real=real)


def algorithmic_bytes(w, real_size, train_steps, trace_rows):
    """bytes the algorithm must move per agent-step between the table store and the SM.
    one-step Basic: read Q[s'][0..A) + read Q[s][a] + write Q[s][a] + 2 B packed transition = (A+2)*R + 2;
    Double: (2A+3)*R + 2; UCB adds 4A (counts row) + 8 (count RMW); traces add 4*R*A per swept row
    (read e, read Q, write Q, write e)."""
    A, R = A_OF_ENV[w["env"]], real_size
    per_step = ((2 * A + 3) * R + 2) if w["policy"] else ((A + 2) * R + 2)
    if w["selector"]:
        per_step += 4 * A + 8
    k = w.get("planning_steps", 0)
    if k:   # Dyna: membership word + k replays, each one 8-byte model entry and one more update
        per_step = per_step * (1 + k) + 8 * k + 4
    return train_steps * per_step + trace_rows * 4 * R * A
