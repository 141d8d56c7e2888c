"""Host-side mirror of the reference's trait surface for the hot path, over the C ABI.

Same names, argument order and error behaviour as JohnVithor/RL-Rust (paths below are under
its `src/`), batched: every object stands for `n_agents` independent reference objects, each
on its own Philox stream.  n_agents = 1 is the reference's single trait object.

    env    = TaxiEnv(100)                                         # env/taxi.rs:57
    policy = TabularPolicy(0.05, 0.0)                             # policy/tabular_policy.rs:15
    sel    = UniformEpsilonGreed(1.0, ("sub", 2e-5), 0.0)         # action_selection/uniform_epsilon_greed.rs:31
    agent  = OneStepAgent(policy, 0.95, sel, qlearning,           # agent/one_step_agent.rs:16
                          n_agents=1 << 20, seed=7)
    rewards, lengths, errors = agent.train(env, 100000, 10000)    # agent.rs:66-118

The Rust closure `epsilon_decay: Rc<dyn Fn(f64)->f64>` cannot run on the device; it is given
as ("sub", k) for `|a| a - k` (bin/taxi.rs:132) or ("mul", k) for `|a| a * k`
(bin/frozen_lake_neural.rs:181).
"""
import numpy as np

from . import _abi as abi
from . import render as _render

# agent.rs:19-45 — the three GetNextQValue functions are passed by name
sarsa = abi.TARGET_SARSA
qlearning = abi.TARGET_QLEARNING
expected_sarsa = abi.TARGET_EXPECTED_SARSA

EnvNotReady = abi.EnvNotReady


# --------------------------------------------------------------------------- Env<T, COUNT>
class Env:
    """env.rs:19-49.  `reset`/`step` act on all agents of the engine the env is bound to; an
    env is bound by the first agent that trains on it (or by `bind`)."""
    kind = None
    COUNT = 0
    ACTIONS = ()

    def __init__(self):
        self._engine = None

    def action_size(self):          # env.rs:20-22
        return self.COUNT

    def _cfg(self):
        return {}

    # render() shows what the step-level reset()/step() calls since the last reset() did: the mirror keeps each env's
    # position from the observations they return (a truncated step answers obs 0 WITHOUT moving, e.g. taxi.rs:148-151).
    # Engines of more than TRACK_LIMIT agents are not tracked (rendering millions of envs is nobody's use).
    TRACK_LIMIT = 4096

    def bind(self, engine):
        self._engine = engine
        self._pos = self._nsteps = None
        return self

    def reset(self):                # env.rs:23
        if self._engine is None:
            raise RuntimeError("env is not bound to an engine yet (train an agent on it, or call bind())")
        obs = self._engine.env_reset()
        if self._engine.N <= self.TRACK_LIMIT:
            self._pos, self._nsteps = obs.astype(np.int64), np.zeros(self._engine.N, np.int64)
        return obs

    def step(self, action):         # env.rs:24 — raises EnvNotReady where the reference returns Err(EnvNotReady)
        if self._engine is None:
            raise RuntimeError("env is not bound to an engine yet")
        a = np.broadcast_to(np.asarray(action, np.uint32), (self._engine.N,))
        out = self._engine.env_step(a)
        if self._pos is not None:
            truncated = self._nsteps >= getattr(self, "max_steps", np.iinfo(np.int64).max)
            self._pos = np.where(truncated, self._pos, out[0].astype(np.int64))
            self._nsteps = self._nsteps + (~truncated)
        return out

    def get_action_label(self, action):   # env.rs:48
        return self.ACTIONS[action]

    def render(self, agent=0):      # env.rs:47 — of ONE of the batched envs (the reference has one)
        """The reference's `render()` string for agent `agent`'s env, from the state the engine holds."""
        if self._engine is None:
            raise RuntimeError("env is not bound to an engine yet")
        if self._pos is None:
            raise RuntimeError("render() follows step-level reset()/step() calls (engines of <= %d agents): call reset() first"
                               % self.TRACK_LIMIT)
        return self._render_state(int(self._pos[agent]), agent)

    def _render_state(self, pos, agent):
        raise NotImplementedError


class BlackJackEnv(Env):
    """env/blackjack.rs:30-163.  Observations are dense indices here; `obs_id` gives the
    reference's fxhash id (blackjack.rs:25-27)."""
    kind = abi.ENV_BLACKJACK
    COUNT = 2
    ACTIONS = ("HIT", "STICK")

    # The engine keeps of a hand only what the rules read (two sums, the dealer's first card, two ace flags).  The
    # cards themselves — which only render() shows — are re-derived on the host from the agents' Philox streams: an env
    # call draws nothing but cards, so the stream words between the positions before and after it ARE its cards
    # (blackjack.rs:54-56,76).  Tracked for step-level use only (engines of <= TRACK_LIMIT agents).
    def bind(self, engine):
        super().bind(engine)
        self._hands = None
        return self

    def _cards_between(self, n0, n1, agent):
        e = self._engine
        words = abi.rng_words(e.seed, e.first_agent_id + agent, int(n0), int(n1 - n0)) if n1 > n0 else []
        return _render.cards_from_words(words)

    def reset(self):                # blackjack.rs:105-116: player gets cards 0 and 1, the dealer cards 2 and 3 (:60-66)
        if self._engine is None or self._engine.N > self.TRACK_LIMIT:
            self._hands = None
            return super().reset()
        n0 = self._engine.states()["rng_n"].copy()
        obs = super().reset()
        n1 = self._engine.states()["rng_n"]
        self._hands = []
        for i in range(self._engine.N):
            c = self._cards_between(n0[i], n1[i], i)
            self._hands.append(([c[0], c[1]], [c[2], c[3]]))
        return obs

    def step(self, action):         # blackjack.rs:118-163: HIT draws one player card, anything else the dealer's to >= 17
        if self._engine is None or self._hands is None:
            return super().step(action)
        a = np.broadcast_to(np.asarray(action, np.uint32), (self._engine.N,))
        n0 = self._engine.states()["rng_n"].copy()
        out = super().step(a)
        n1 = self._engine.states()["rng_n"]
        for i in range(self._engine.N):
            c = self._cards_between(n0[i], n1[i], i)
            (self._hands[i][0] if a[i] == 0 else self._hands[i][1]).extend(c)
        return out

    def _render_state(self, pos, agent):   # blackjack.rs:165-184
        if getattr(self, "_hands", None) is None:
            raise RuntimeError("BlackJackEnv.render shows the cards dealt by step-level reset()/step() calls since the "
                               "last reset (engines of <= %d agents); a fused train()/evaluate() keeps only the sums" % self.TRACK_LIMIT)
        player, dealer = self._hands[agent]
        return _render.render_blackjack(bool(self._engine.states()["env_ready"][agent]), dealer, player)

    @staticmethod
    def obs_id(dense_index):
        return abi.blackjack_obs_id(int(dense_index))

    @staticmethod
    def dense_index(obs_id):
        return abi.blackjack_dense_index(int(obs_id))


class FrozenLakeEnv(Env):
    """env/frozen_lake.rs:12-134.  `map` is any list of equally long rows of S / F / H / G cells, as
    `FrozenLakeEnv::new(map: &[&str], ..)` takes (:48) — FrozenLakeEnv.MAP_4X4 and MAP_8X8 are the crate's two constants
    (:23-28).  Every 'S' is a start cell (:54-66)."""
    kind = abi.ENV_FROZEN_LAKE
    COUNT = 4
    ACTIONS = ("LEFT", "DOWN", "RIGHT", "UP")
    MAP_4X4 = ("SFFF", "FHFH", "FFFH", "HFFG")
    MAP_8X8 = ("SFFFFFFF", "FFFFFFFF", "FFFHFFFF", "FFFFFHFF", "FFFHFFFF", "FHHFFFHF", "FHFFHFHF", "FFFHFFFG")

    def __init__(self, map, is_slippery, max_steps):
        super().__init__()
        self.map = tuple(str(r) for r in map)
        if not self.map or any(len(r) != len(self.map[0]) or set(r) - set("SFHG") for r in self.map):
            raise ValueError("map rows must be non-empty, equally long and made of S, F, H, G")
        self.map_id = 0 if self.map == self.MAP_4X4 else (1 if self.map == self.MAP_8X8 else abi.MAP_CUSTOM)
        self.is_slippery = bool(is_slippery)
        self.max_steps = int(max_steps)

    def _cfg(self):
        cfg = dict(map_id=self.map_id, slippery=self.is_slippery, max_steps=self.max_steps)
        if self.map_id == abi.MAP_CUSTOM:
            cfg["map_rows"] = list(self.map)
        return cfg

    def _render_state(self, pos, agent):   # frozen_lake.rs:136-149
        return _render.render_frozen_lake(self.map, pos)


class CliffWalkingEnv(Env):
    """env/cliff_walking.rs:6-89"""
    kind = abi.ENV_CLIFF_WALKING
    COUNT = 4
    ACTIONS = ("LEFT", "DOWN", "RIGHT", "UP")

    def __init__(self, max_steps):
        super().__init__()
        self.max_steps = int(max_steps)

    def _cfg(self):
        return dict(max_steps=self.max_steps)

    def _render_state(self, pos, agent):   # cliff_walking.rs:91-102
        return _render.render_cliff_walking(pos)


class TaxiEnv(Env):
    """env/taxi.rs:10-159"""
    kind = abi.ENV_TAXI
    COUNT = 6
    ACTIONS = ("DOWN", "UP", "RIGHT", "LEFT", "PICKUP", "DROPOFF")

    def __init__(self, max_steps):
        super().__init__()
        self.max_steps = int(max_steps)

    def _cfg(self):
        return dict(max_steps=self.max_steps)

    @staticmethod
    def decode(i):                  # taxi.rs:44-55
        return (i // 100, (i // 20) % 5, (i // 4) % 5, i % 4)

    def _render_state(self, pos, agent):   # taxi.rs:161-172
        return _render.render_taxi(pos)


def example_episode(env, get_action, out=print):
    """`Agent::example` (agent.rs:143-163) over a bound env and an agent's `get_action`: the env before each step, the
    action's label (`{:?}` of a &str: quoted), the step's reward; when the episode ends the last view, the episode's
    reward and its length.  The draws and UCB counts it consumes are the reference's (it sits between train and evaluate
    in the bins, bin/taxi.rs:184-186)."""
    if env._engine is None:
        raise RuntimeError("env is not bound to an engine yet")
    if env._engine.N != 1:
        raise RuntimeError("example() shows the reference's single env: use an engine of one agent")

    def transitions():
        curr_action = get_action(env.reset())
        while True:
            before = env.render(0)
            next_obs, reward, terminated = env.step(curr_action)
            shown, done = int(curr_action[0]), bool(terminated[0])
            curr_action = get_action(next_obs)           # also on the terminal observation (agent.rs:153)
            yield before, shown, float(reward[0]), done, env.render(0) if done else None
            if done:
                return

    lines = _render.example_lines(env.get_action_label, transitions())
    for ln in lines:
        out(ln)
    return lines


# --------------------------------------------------------------------------- Policy<T, COUNT>
class TabularPolicy:
    """policy/tabular_policy.rs:8-44 ("Basic")"""
    kind = abi.POLICY_BASIC

    def __init__(self, learning_rate, default_value):
        self.learning_rate = float(learning_rate)
        self.default_value = float(default_value)


class DoubleTabularPolicy(TabularPolicy):
    """policy/double_tabular_policy.rs:8-67"""
    kind = abi.POLICY_DOUBLE


# --------------------------------------------------------------------------- ActionSelection<T, COUNT>
class UniformEpsilonGreed:
    """action_selection/uniform_epsilon_greed.rs:8-80"""
    kind = abi.SEL_EPS_GREEDY

    def __init__(self, epsilon, epsilon_decay, final_epsilon):
        op, k = epsilon_decay
        if op not in ("sub", "mul"):
            raise ValueError("epsilon_decay must be ('sub', k) or ('mul', k)")
        self.initial_epsilon = float(epsilon)
        self.decay_kind = abi.DECAY_SUB if op == "sub" else abi.DECAY_MUL
        self.decay_param = float(k)
        self.final_epsilon = float(final_epsilon)


class UpperConfidenceBound:
    """action_selection/upper_confidence_bound.rs:9-68"""
    kind = abi.SEL_UCB

    def __init__(self, confidence_level):
        self.confidence_level = float(confidence_level)


# --------------------------------------------------------------------------- Agent<T, COUNT>
class _Agent:
    agent_kind = None

    def __init__(self, policy, discount_factor, action_selection, lambda_factor, get_next_q_value, *, n_agents=1,
                 seed=0x5EED0001, first_agent_id=0, real="f32", device=0):
        self.policy = policy
        self.discount_factor = float(discount_factor)
        self.lambda_factor = float(lambda_factor)
        self._selectors = {}
        self._remember(action_selection)
        self.selector_kind = action_selection.kind
        self.get_next_q_value = get_next_q_value
        self.n_agents, self.seed, self.first_agent_id, self.device = int(n_agents), int(seed), int(first_agent_id), device
        self.real = abi.REAL_F32 if real in ("f32", abi.REAL_F32) else abi.REAL_F64
        self.engine = None

    def _remember(self, sel):
        self._selectors[sel.kind] = sel

    def _bind(self, env):
        if self.engine is not None:
            if env._engine is not self.engine:
                raise RuntimeError("agent is already bound to another env")
            return
        eg = self._selectors.get(abi.SEL_EPS_GREEDY)
        ucb = self._selectors.get(abi.SEL_UCB)
        kw = dict(n_agents=self.n_agents, policy=self.policy.kind, selector=self.selector_kind,
                  target=self.get_next_q_value, agent=self.agent_kind, real=self.real,
                  learning_rate=self.policy.learning_rate, default_value=self.policy.default_value,
                  discount_factor=self.discount_factor, lambda_factor=self.lambda_factor, seed=self.seed,
                  first_agent_id=self.first_agent_id, device=self.device)
        if eg is not None:
            kw.update(initial_epsilon=eg.initial_epsilon, decay_kind=eg.decay_kind, epsilon_decay=eg.decay_param,
                      final_epsilon=eg.final_epsilon)
        if ucb is not None:
            kw.update(confidence_level=ucb.confidence_level)
        kw.update(env._cfg())
        self.engine = abi.Engine(env.kind, **kw)
        env.bind(self.engine)

    # agent.rs:48
    def set_future_q_value_func(self, func):
        self.get_next_q_value = func
        if self.engine is not None:
            self.engine.set_target(func)

    # agent.rs:50 — the engine holds one parameter set per selector kind, fixed at bind time
    def set_action_selector(self, action_selector):
        if self.engine is not None and action_selector.kind not in self._selectors:
            raise RuntimeError("selector parameters must be known before the agent is bound; pass every selector "
                               "you will use to register_selector() first")
        self._remember(action_selector)
        self.selector_kind = action_selector.kind
        if self.engine is not None:
            self.engine.set_selector(action_selector.kind)

    def register_selector(self, action_selector):
        """Make a selector's parameters known before binding (the CLI builds both up front, bin/taxi.rs:129-136)."""
        if self.engine is not None:
            raise RuntimeError("agent already bound")
        self._remember(action_selector)

    def get_action(self, obs):      # agent.rs:52
        return self.engine.get_action(obs)

    def update(self, curr_obs, curr_action, reward, terminated, next_obs, next_action):   # agent.rs:54-62
        return self.engine.update(curr_obs, curr_action, reward, terminated, next_obs, next_action)

    def reset(self):                # agent.rs:64
        if self.engine is not None:
            self.engine.agent_reset()

    def train(self, env, n_episodes, eval_at, *, raw=False):
        """agent.rs:66-118.  Returns (reward_history, episode_length, training_error).

        reward_history and episode_length are the reference's vectors ([n_episodes] for one
        agent, [n_agents, n_episodes] otherwise).  With ONE agent training_error is the reference's
        too: one temporal difference per training step, episodes concatenated (agent.rs:98,117;
        rlb_train_out.td_steps).  With a batch of agents it is per EPISODE — the sum of the
        episode's temporal differences, [n_agents, n_episodes] — because a per-step stream for
        millions of agents does not fit anywhere; Engine.train(td_capacity=...) returns the
        per-step streams of small batches.  raw=True returns the engine's dict instead
        (per-episode sums over agents, step counters, kernel time).
        """
        if eval_at == 0:
            raise ZeroDivisionError("attempt to calculate the remainder with a divisor of zero")   # agent.rs:107
        self._bind(env)
        if raw:
            return self.engine.train(n_episodes, eval_at)
        if self.n_agents == 1:
            # the most steps n episodes can take: max_steps + 1 each (Blackjack: a hand holds 16 cards, blackjack.rs:32-35)
            cap = int(n_episodes) * (32 if env.kind == abi.ENV_BLACKJACK else int(self.engine.cfg.max_steps) + 1)
            res = self.engine.train(n_episodes, eval_at, sums=False, episodes=True, td_capacity=cap if cap * 8 <= (1 << 31) else 0)
            ep = res["episodes"][:, 0]
            if "td_steps" in res:
                n = int(res["td_count"][0])
                return ep["ret"].astype(np.float64), ep["length"].astype(np.uint64), res["td_steps"][0, :n].astype(np.float64)
            return ep["ret"].astype(np.float64), ep["length"].astype(np.uint64), ep["td_sum"].astype(np.float64)
        res = self.engine.train(n_episodes, eval_at, sums=False, episodes=True)
        ep = res["episodes"]
        return (ep["ret"].T.astype(np.float64), ep["length"].T.astype(np.uint64), ep["td_sum"].T.astype(np.float64))

    def example(self, env, out=print):
        """agent.rs:143-163: one episode through the step-level calls, printing what the reference prints.  Needs an
        engine of ONE agent (the reference's single env).  Returns the printed lines."""
        self._bind(env)
        return example_episode(env, self.get_action, out=out)

    def evaluate(self, env, n_episodes):
        """agent.rs:120-141.  Returns (reward_history, episode_length)."""
        self._bind(env)
        res = self.engine.evaluate(n_episodes, sums=False, episodes=True)
        ep = res["episodes"]
        if self.n_agents == 1:
            return ep["ret"][:, 0].astype(np.float64), ep["length"][:, 0].astype(np.uint64)
        return ep["ret"].T.astype(np.float64), ep["length"].T.astype(np.uint64)


class OneStepAgent(_Agent):
    """agent/one_step_agent.rs:7-86"""
    agent_kind = abi.AGENT_ONE_STEP

    def __init__(self, policy, discount_factor, action_selection, get_next_q_value, **kw):
        super().__init__(policy, discount_factor, action_selection, 0.0, get_next_q_value, **kw)


class ElegibilityTracesAgent(_Agent):
    """agent/elegibility_traces_agent.rs:8-104"""
    agent_kind = abi.AGENT_TRACES

    def __init__(self, policy, discount_factor, action_selection, lambda_factor, get_next_q_value, **kw):
        super().__init__(policy, discount_factor, action_selection, lambda_factor, get_next_q_value, **kw)


# --------------------------------------------------------------------------- Model<T, COUNT> / Dyna
class RandomModel:
    """model/random_model.rs:9-45 — first-seen transitions `(obs, action) -> (next_obs, reward)` in insertion order.
    Lives on the device next to the agent it is given to (InternalModelAgent binds it)."""

    def __init__(self):
        self._engine = None

    def _need(self):
        if self._engine is None:
            raise RuntimeError("the model is bound when its InternalModelAgent first sees an env")
        return self._engine

    def get_info(self):             # model.rs:13 — (state, action, next_state, reward), index drawn with gen_range
        obs, action, next_obs, reward = self._need().model_get_info()
        return obs, action, next_obs, reward

    def add_info(self, obs, action, reward, next_obs):   # model.rs:14
        self._need().model_add_info(obs, action, reward, next_obs)

    def reset(self):                # model.rs:15
        self._need().model_reset()

    def entries(self):
        """(len [n_agents], entries [n_agents, capacity]) snapshot of the remembered transitions."""
        return self._need().download_model()


class InternalModelAgent:
    """agent/internal_model_agent.rs:9-85 — Dyna: `InternalModelAgent::new(agent, model, planning_length)` borrows an
    agent (which keeps whatever it has learned) and, after each of its updates, remembers the transition and replays
    `planning_length` remembered ones.  planning_length must be > 0 here (0 is the plain agent)."""

    def __init__(self, agent, model, planning_length):
        if int(planning_length) <= 0:
            raise ValueError("planning_length must be > 0")
        self.agent, self.model, self.planning_steps = agent, model, int(planning_length)
        if agent.engine is not None:
            self._attach()

    def _attach(self):
        if self.model._engine is not self.agent.engine:
            self.agent.engine.set_model(self.planning_steps)
            self.model._engine = self.agent.engine

    def set_future_q_value_func(self, func):             # :34-36
        self.agent.set_future_q_value_func(func)

    def set_action_selector(self, action_selector):      # :38-40
        self.agent.set_action_selector(action_selector)

    def get_action(self, obs):                           # :42-44
        return self.agent.get_action(obs)

    def update(self, curr_obs, curr_action, reward, terminated, next_obs, next_action):   # :46-79
        self._attach()
        return self.agent.update(curr_obs, curr_action, reward, terminated, next_obs, next_action)

    def reset(self):                                     # :81-84
        self.agent.reset()                               # the engine empties the attached model with it

    def train(self, env, n_episodes, eval_at, **kw):
        self.agent._bind(env)
        self._attach()
        return self.agent.train(env, n_episodes, eval_at, **kw)

    def evaluate(self, env, n_episodes):
        self.agent._bind(env)
        self._attach()
        return self.agent.evaluate(env, n_episodes)

    def release(self):
        """End of the borrow (the wrapper going out of scope in the reference): the agent carries on without a model."""
        if self.agent.engine is not None and self.model._engine is self.agent.engine:
            self.agent.engine.set_model(0)
        self.model._engine = None
