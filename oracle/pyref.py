"""oracle/pyref.py — a second, independent restatement of the reference hot path in pure Python.
TEST INFRASTRUCTURE ONLY (small cases; Python floats are the reference's f64).

Written from the reference sources (paths under /root/reference/src) separately from oracle.hpp, with dicts for
the FxHashMaps, so that two restatements have to agree before the C++ oracle is trusted
(tests/test_pyref_cross.py).  Shares only the RNG injection contract of SURVEY.md §8.2 with it.
"""
import math
import struct

M64 = (1 << 64) - 1
M32 = 0xFFFFFFFF


# ------------------------------------------------------------------ RNG contract
def philox(ctr, key):
    c = list(ctr)
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k0, p1 & M32, (p0 >> 32) ^ c[3] ^ k1, p0 & M32]
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c


class Stream:
    def __init__(self, seed, agent):
        self.seed, self.agent, self.n = seed, agent, 0
        self._blk, self._words = None, None

    def u32(self):
        b = self.n >> 2
        if b != self._blk:
            self._words = philox([b & M32, b >> 32, self.agent & M32, self.agent >> 32], (self.seed & M32, self.seed >> 32))
            self._blk = b
        w = self._words[self.n & 3]
        self.n += 1
        return w

    def u64(self):
        lo = self.u32()
        return lo | (self.u32() << 32)

    def unit(self):                     # rand 0.8.5 Uniform<f64>(0.0..1.0)
        return (self.u64() >> 12) * 2.0 ** -52

    def below(self, rng):               # rand 0.8.5 Uniform<usize>(0..rng)
        zone = M64 - ((M64 - rng + 1) % rng)
        while True:
            m = self.u64() * rng
            if (m & M64) <= zone:
                return m >> 64

    def gen_range(self, rng):           # rand 0.8.5 Rng::gen_range(0..rng) on usize: sample_single's cheap zone
        zone = ((rng << (64 - rng.bit_length())) - 1) & M64
        while True:
            m = self.u64() * rng
            if (m & M64) <= zone:
                return m >> 64

    def card(self):                     # rand 0.8.5 Uniform<u8>(1..11), via u32
        while True:
            m = self.u32() * 10
            if (m & M32) <= 0xFFFFFFF9:
                return 1 + (m >> 32)


def fxhash3(p, d, ace):                 # fxhash::hash(&BlackJackObservation) — env/blackjack.rs:25-27
    h = 0
    for b in (p, d, 1 if ace else 0):
        h = ((((h << 5) | (h >> 59)) & M64) ^ b) * 0x517CC1B727220A95 & M64
    return h


def portable_log(x):
    """Same portable ln as the contract (fdlibm reduction, no FMA); x finite >= 1."""
    ln2_hi, ln2_lo = 6.93147180369123816490e-01, 1.90821492927058770002e-10
    Lg = (6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
          1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01)
    bits = struct.unpack("<Q", struct.pack("<d", x))[0]
    hx, lx = bits >> 32, bits & M32
    k = (hx >> 20) - 1023
    hx &= 0x000FFFFF
    i = (hx + 0x95F64) & 0x100000
    x = struct.unpack("<d", struct.pack("<Q", ((hx | (i ^ 0x3FF00000)) << 32) | lx))[0]
    k += i >> 20
    f = x - 1.0
    dk = float(k)
    if (0x000FFFFF & (2 + hx)) < 3:
        if f == 0.0:
            return 0.0 if k == 0 else dk * ln2_hi + dk * ln2_lo
        R = f * f * (0.5 - 0.33333333333333333 * f)
        return f - R if k == 0 else dk * ln2_hi - ((R - dk * ln2_lo) - f)
    s = f / (2.0 + f)
    z = s * s
    w = z * z
    t1 = w * (Lg[1] + w * (Lg[3] + w * Lg[5]))
    t2 = z * (Lg[0] + w * (Lg[2] + w * (Lg[4] + w * Lg[6])))
    R = t2 + t1
    if ((hx - 0x6147A) | (0x6B851 - hx)) > 0:
        hfsq = 0.5 * f * f
        return f - (hfsq - s * (hfsq + R)) if k == 0 else dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f)
    return f - s * (f - R) if k == 0 else dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f)


# ------------------------------------------------------------------ utils.rs
def argmax(v):                          # utils.rs:1-11
    best, res = v[0], 0
    for i, x in enumerate(v):
        if x > best:
            best, res = x, i
    return res


def vmax(v):                            # utils.rs:13-21
    best = v[0]
    for x in v:
        if x > best:
            best = x
    return best


def categorical_sample(probs, random):  # utils.rs:33-43
    b, flags = 0.0, []
    for a in probs:
        b += a
        flags.append(b > random)
    return argmax(flags)


def inc(nrow, ncol, row, col, a):       # utils.rs:53-76
    if a == 0:
        return row, max(col - 1, 0)
    if a == 1:
        return min(row + 1, nrow - 1), col
    if a == 2:
        return row, min(col + 1, ncol - 1)
    if a == 3:
        return max(row - 1, 0), col
    return row, col


# ------------------------------------------------------------------ envs
class BlackJack:                        # env/blackjack.rs
    COUNT = 2

    def __init__(self, rng):
        self.rng, self.ready = rng, False
        self._deal()                    # new() :57

    def _deal(self):
        self.player = [self.rng.card(), self.rng.card()]
        self.dealer = [self.rng.card(), self.rng.card()]
        self.p_ace = 1 in self.player[:2]
        self.d_ace = 1 in self.dealer[:2]

    @staticmethod
    def _score(hand, ace):
        s = sum(hand)
        return s + 10 if ace and s + 10 <= 21 else s

    def dense(self, obs):
        return obs[3]

    def reset(self):
        self._deal()
        self.ready = True
        return self._obs(self._score(self.player, self.p_ace), self.dealer[0])

    def _obs(self, p, d):
        return (fxhash3(p, d, self.p_ace), p, d, ((p - 4) * 26 + (d - 1)) * 2 + (1 if self.p_ace else 0))

    def step(self, action):
        if not self.ready:
            return None
        if action == 0:
            self.player.append(self.rng.card())
            p = self._score(self.player, self.p_ace)
            if p > 21:
                self.ready = False
                return self._obs(p, self._score(self.dealer, self.d_ace)), -1.0, True
            return self._obs(p, self.dealer[0]), 0.0, False
        self.ready = False
        d = self._score(self.dealer, self.d_ace)
        while d < 17:
            self.dealer.append(self.rng.card())
            d = self._score(self.dealer, self.d_ace)
        p = self._score(self.player, self.p_ace)
        if d > 21:
            return self._obs(p, d), 1.0, True
        return self._obs(p, d), (1.0 if p > d else (-1.0 if p < d else 0.0)), True


class TableEnv:
    """Shared step/reset shape of frozen_lake.rs / cliff_walking.rs / taxi.rs."""
    truncation_reward = 0.0

    def dense(self, obs):
        return obs

    def step(self, action):
        if not self.ready:
            return None
        if self.curr_step >= self.max_steps:
            self.ready = False
            return 0, self.truncation_reward, True
        self.curr_step += 1
        s, r, t = self._transition(action)
        self.pos = s
        if t:
            self.ready = False
        return s, r, t


class FrozenLake(TableEnv):             # env/frozen_lake.rs
    COUNT = 4
    MAPS = {0: ["SFFF", "FHFH", "FFFH", "HFFG"],
            1: ["SFFFFFFF", "FFFFFFFF", "FFFHFFFF", "FFFFFHFF", "FFFHFFFF", "FHHFFFHF", "FHFFHFHF", "FFFHFFFG"]}

    def __init__(self, map_id, slippery, max_steps, rng):
        m = self.MAPS[map_id]
        self.rng, self.max_steps, self.ready, self.pos, self.curr_step = rng, max_steps, False, 0, 0
        n = len(m)
        flat = "".join(m)
        starts = [i for i, ch in enumerate(flat) if ch == "S"]
        self.init = [0.0] * len(flat)
        for i in starts:
            self.init[i] = 1.0 / len(starts)
        self.probs = {}
        for row in range(n):
            for col in range(n):
                s = row * n + col
                for a in range(4):
                    li = [(0.0, 0, 0.0, False)] * 3
                    if m[row][col] in "GH":
                        li[0] = (1.0, s, 0.0, True)
                    else:
                        cands = [(a - 1) % 4, a, (a + 1) % 4] if slippery else [a]
                        for i, b in enumerate(cands):
                            r2, c2 = inc(n, n, row, col, b)
                            ch = m[r2][c2]
                            li[i] = ((1.0 / 3.0) if slippery else 1.0, r2 * n + c2, 1.0 if ch == "G" else 0.0, ch in "GH")
                    self.probs[(s, a)] = li

    def reset(self):
        self.pos = categorical_sample(self.init, self.rng.unit())
        self.ready, self.curr_step = True, 0
        return self.pos

    def _transition(self, action):
        li = self.probs[(self.pos, action)]
        i = categorical_sample([t[0] for t in li], self.rng.unit())
        return li[i][1], li[i][2], li[i][3]


class CliffWalking(TableEnv):           # env/cliff_walking.rs
    COUNT = 4
    truncation_reward = -100.0

    def __init__(self, max_steps):
        self.max_steps, self.ready, self.pos, self.curr_step = max_steps, False, 0, 0

    def reset(self):
        self.pos, self.ready, self.curr_step = 36, True, 0
        return 36

    def _transition(self, action):
        r, c = inc(4, 12, self.pos // 12, self.pos % 12, action)
        ns = r * 12 + c
        lose = 37 <= ns <= 46
        return ns, (-100.0 if lose else -1.0), lose or ns == 47


class Taxi(TableEnv):                   # env/taxi.rs
    COUNT = 6
    MAP = ["+---------+", "|R: | : :G|", "| : | : : |", "| : : : : |", "| | : | : |", "|Y| : |B: |", "+---------+"]
    LOCS = [(0, 0), (0, 4), (4, 0), (4, 3)]

    def __init__(self, max_steps, rng):
        self.rng, self.max_steps, self.ready, self.pos, self.curr_step = rng, max_steps, False, 0, 0
        self.init = [0.0] * 500
        total = 0.0
        for st in range(500):
            d, p = st % 4, (st // 4) % 5
            if p < 4 and p != d:
                self.init[st] += 1.0
                total += 1.0
        self.init = [v / total for v in self.init]

    def reset(self):
        self.pos = categorical_sample(self.init, self.rng.unit())
        self.ready, self.curr_step = True, 0
        return self.pos

    def _transition(self, action):
        st = self.pos
        dest, pas, col, row = st % 4, (st // 4) % 5, (st // 20) % 5, st // 100
        nr, nc, np_, reward, term = row, col, pas, -1.0, False
        if action == 0:
            nr = min(row + 1, 4)
        elif action == 1:
            nr = max(row - 1, 0)
        if action == 2 and self.MAP[1 + row][2 * col + 2] == ":":
            nc = min(col + 1, 4)
        elif action == 3 and self.MAP[1 + row][2 * col] == ":":
            nc = max(col - 1, 0)
        elif action == 4:
            if pas < 4 and (row, col) == self.LOCS[pas]:
                np_ = 4
            else:
                reward = -10.0
        elif action == 5:
            if (row, col) == self.LOCS[dest] and pas == 4:
                np_, term, reward = dest, True, 20.0
            else:
                reward = -10.0
        return ((nr * 5 + nc) * 5 + np_) * 4 + dest, reward, term


# ------------------------------------------------------------------ policies / selectors / agents
class Tabular:                          # policy/tabular_policy.rs
    def __init__(self, lr, default, count):
        self.lr, self.default, self.q, self.count = lr, [default] * count, {}, count

    def predict(self, o):
        return list(self.q.get(o, self.default))

    get_values = predict

    def update(self, o, a, td):
        self.q.setdefault(o, list(self.default))[a] += self.lr * td

    def after_update(self):
        pass

    def reset(self):
        self.q = {}

    def tables(self):
        return [self.q]


class DoubleTabular:                    # policy/double_tabular_policy.rs
    def __init__(self, lr, default, count):
        self.lr, self.default, self.a, self.b, self.flag, self.count = lr, [default] * count, {}, {}, True, count

    def predict(self, o):
        av, bv = self.a.get(o, self.default), self.b.get(o, self.default)
        return [(x + y) / 2.0 for x, y in zip(av, bv)]

    def get_values(self, o):
        return list((self.a if self.flag else self.b).get(o, self.default))

    def update(self, o, a, td):
        (self.b if self.flag else self.a).setdefault(o, list(self.default))[a] += self.lr * td

    def after_update(self):
        self.flag = not self.flag

    def reset(self):
        self.a, self.b = {}, {}

    def tables(self):
        return [self.a, self.b]


class EpsGreedy:                        # action_selection/uniform_epsilon_greed.rs
    def __init__(self, eps, kind, param, final, count, rng):
        self.eps0 = self.eps = eps
        self.kind, self.param, self.final, self.count, self.rng = kind, param, final, count, rng

    def get_action(self, o, values):
        if self.eps != 0.0 and self.rng.unit() < self.eps:
            return self.rng.below(self.count)
        return argmax(values)

    def update(self):
        new = self.eps - self.param if self.kind == 0 else self.eps * self.param
        self.eps = self.eps if self.final > new else new

    def probs(self, o, values):
        p = [self.eps / self.count] * self.count
        p[argmax(values)] = 1.0 - self.eps
        return p

    def reset(self):
        self.eps = self.eps0


class UCB:                              # action_selection/upper_confidence_bound.rs
    MIN_POSITIVE = 2.2250738585072014e-308

    def __init__(self, c, count):
        self.c, self.count, self.n, self.t = c, count, {}, 1

    def _ucbs(self, o, values):
        n = self.n.setdefault(o, [0] * self.count)
        ln_t = portable_log(float(self.t))
        out = []
        for i in range(self.count):
            den = float(n[i]) + self.MIN_POSITIVE
            q = ln_t / den if not (den == 0.0) else math.inf
            out.append(values[i] + self.c * _sqrt(q))
        return out, n

    def get_action(self, o, values):
        u, n = self._ucbs(o, values)
        a = argmax(u)
        n[a] += 1
        self.t += 1
        return a

    def update(self):
        pass

    def probs(self, o, values):
        u, _ = self._ucbs(o, values)
        total = 0.0
        for x in u:
            total += x
        return [_div(x, total) for x in u]

    def reset(self):
        self.n, self.t = {}, 1


def _sqrt(x):
    return math.inf if x == math.inf else (math.nan if x != x else math.sqrt(x))


def _div(a, b):                         # IEEE division incl. inf/inf, x/0
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0.0:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)


def ieee_div(a, b):
    if b == 0.0:
        return _div(a, b)
    return a / b


def target_value(kind, next_q, next_action, probs):   # agent.rs:19-45
    if kind == 0:
        return next_q[next_action]
    if kind == 1:
        return vmax(next_q)
    acc = 0.0
    for p, q in zip(probs, next_q):
        acc += p * q
    return acc


class Agent:
    def __init__(self, policy, gamma, selector, target, traces=False, lam=0.0):
        self.policy, self.gamma, self.selector, self.target, self.traces, self.lam = policy, gamma, selector, target, traces, lam
        self.trace = {}
        self.eval_steps = 0

    def get_action(self, o):
        return self.selector.get_action(o, self.policy.predict(o))

    def update(self, s, a, r, term, s2, a2):          # one_step_agent.rs:53-86 / elegibility_traces_agent.rs:61-104
        nq = self.policy.get_values(s2)
        fut = target_value(self.target, nq, a2, self.selector.probs(s2, nq))
        td = r + self.gamma * fut - self.policy.get_values(s)[a]
        if not self.traces:
            self.policy.update(s, a, td)
        else:
            self.trace.setdefault(s, [0.0] * self.policy.count)[a] += 1.0
            for o, row in self.trace.items():
                for act in range(len(row)):
                    self.policy.update(o, act, td * row[act])
                    row[act] *= self.gamma * self.lam
        self.policy.after_update()
        if term:
            if self.traces:
                self.trace = {}
            self.selector.update()
        return td

    def reset(self):
        self.selector.reset()
        self.policy.reset()

    def train(self, env, n_episodes, eval_at, log=None):   # agent.rs:66-118
        rewards, lengths, errors = [], [], []
        for ep in range(n_episodes):
            steps, total = 0, 0.0
            o = env.reset()
            a = self.get_action(_key(o))
            if log is not None:
                log.append((0, env.dense(o), a, 0.0, False, 0.0))
            while True:
                steps += 1
                o2, r, term = env.step(a)
                a2 = self.get_action(_key(o2))
                td = self.update(_key(o), a, r, term, _key(o2), a2)
                errors.append(td)
                if log is not None:
                    log.append((1, env.dense(o2), a2, r, term, td))
                o, a = o2, a2
                total += r
                if term:
                    rewards.append(total)
                    break
            if ep % eval_at == 0:
                self.evaluate(env, 100, log)
            lengths.append(steps)
        return rewards, lengths, errors

    def evaluate(self, env, n, log=None):                  # agent.rs:120-141
        rewards, lengths = [], []
        for _ in range(n):
            steps, total = 0, 0.0
            o = env.reset()
            a = self.get_action(_key(o))
            if log is not None:
                log.append((0, env.dense(o), a, 0.0, False, 0.0))
            while True:
                steps += 1
                o, r, term = env.step(a)
                a = self.get_action(_key(o))
                if log is not None:
                    log.append((2, env.dense(o), a, r, term, 0.0))
                total += r
                if term:
                    rewards.append(total)
                    break
            self.eval_steps += steps
            lengths.append(steps)
        return rewards, lengths


class DynaAgent(Agent):                 # agent/internal_model_agent.rs around model/random_model.rs
    """Wraps `inner`: every real update is followed by add_info and `planning` replays of remembered transitions."""

    def __init__(self, inner, planning, rng):
        self.inner, self.planning, self.rng = inner, planning, rng
        self.model = {}                 # IndexMap: python dicts keep insertion order too
        self.eval_steps = 0

    def get_action(self, o):
        return self.inner.get_action(o)

    def update(self, s, a, r, term, s2, a2):          # internal_model_agent.rs:46-79
        td = self.inner.update(s, a, r, term, s2, a2)
        self.model.setdefault((s, a), (s2, r))        # random_model.rs:37-41
        for _ in range(self.planning):
            (ms, ma), (ms2, mr) = list(self.model.items())[self.rng.gen_range(len(self.model))]   # :27-35
            self.inner.update(ms, ma, mr, False, ms2, self.inner.get_action(ms2))
        return td

    def reset(self):                                  # :81-84
        self.inner.reset()
        self.model = {}


def _key(obs):
    return obs[0] if isinstance(obs, tuple) else obs


def build(env_kind, *, agent_id=0, seed=0x5EED0001, map_id=1, slippery=True, max_steps=100, policy=0, selector=0, target=1,
          traces=False, lr=0.05, gamma=0.95, lam=0.5, eps0=1.0, eps_decay=2e-5, eps_final=0.0, decay_kind=0, ucb_c=0.5,
          default_q=0.0, planning=0):
    rng = Stream(seed, agent_id)
    env = {0: lambda: BlackJack(rng), 1: lambda: FrozenLake(map_id, slippery, max_steps, rng),
           2: lambda: CliffWalking(max_steps), 3: lambda: Taxi(max_steps, rng)}[env_kind]()
    count = env.COUNT
    pol = (DoubleTabular if policy else Tabular)(lr, default_q, count)
    sel = UCB(ucb_c, count) if selector else EpsGreedy(eps0, decay_kind, eps_decay, eps_final, count, rng)
    agent = Agent(pol, gamma, sel, target, traces, lam)
    if planning:
        agent = DynaAgent(agent, planning, rng)
    return env, agent, rng
