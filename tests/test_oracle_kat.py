"""CPU tests that pin the oracle (no GPU): published KATs (Philox / Random123), the
[derived] constants of SURVEY.md §8.2, and hand-worked cases of the reference's rules.

The reference ships no tests or golden vectors (parity unpinned by the reference); these
constants are what anchors the restatement in oracle/oracle.hpp to the cited lines."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import oracle_py as O


@pytest.fixture(scope="module")
def L():
    return O.lib()


def philox(L, ctr, key):
    c = np.array(ctr, np.uint32)
    k = np.array(key, np.uint32)
    o = np.zeros(4, np.uint32)
    L.oracle_philox4x32_10(O._p(c), O._p(k), O._p(o))
    return [int(x) for x in o]


def test_philox_kat(L):
    # Random123 kat_vectors, philox4x32 10 rounds
    assert philox(L, [0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox(L, [0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox(L, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_stream_layout(L):
    """w[n] = Philox(key=seed, ctr=(n>>2 lo, hi, agent lo, hi))[n&3]; any start offset gives the same words."""
    seed, agent = 0x0123456789abcdef, 0xfedcba9876543210
    w = np.zeros(40, np.uint32)
    L.oracle_stream_words(seed, agent, 0, 40, O._p(w))
    for blk in range(10):
        ref = philox(L, [blk, 0, agent & 0xffffffff, agent >> 32], [seed & 0xffffffff, seed >> 32])
        assert [int(x) for x in w[4 * blk:4 * blk + 4]] == ref
    w2 = np.zeros(20, np.uint32)
    L.oracle_stream_words(seed, agent, 7, 20, O._p(w2))
    assert np.array_equal(w2, w[7:27])


def test_uniform_f64_mapping(L):
    """rand 0.8.5 UniformFloat<f64>: u = (next_u64 >> 12) * 2^-52, low word first."""
    seed, agent = 99, 3
    w = np.zeros(64, np.uint32)
    L.oracle_stream_words(seed, agent, 0, 64, O._p(w))
    out = np.zeros(20, np.float64)
    used = L.oracle_sample(seed, agent, 5, 0, 0, 20, O._p(out))   # odd start: u64s straddle Philox blocks
    assert used == 40
    for i in range(20):
        lo, hi = int(w[5 + 2 * i]), int(w[6 + 2 * i])
        k = ((hi << 32) | lo) >> 12
        assert out[i] == k * 2.0 ** -52
        assert 0.0 <= out[i] < 1.0


@pytest.mark.parametrize("rng_range", [2, 4, 6])
def test_uniform_usize_mapping(L, rng_range):
    """rand 0.8.5 UniformInt<usize>: hi word of v * range, rejecting lo > zone."""
    seed, agent = 7, 11
    n = 2000
    w = np.zeros(2 * n + 64, np.uint32)
    L.oracle_stream_words(seed, agent, 0, len(w), O._p(w))
    out = np.zeros(n, np.float64)
    used = L.oracle_sample(seed, agent, 0, 1, rng_range, n, O._p(out))
    zone = (1 << 64) - 1 - (((1 << 64) - rng_range) % rng_range)
    assert zone == {2: 0xffffffffffffffff, 4: 0xffffffffffffffff, 6: 0xfffffffffffffffb}[rng_range]
    pos, got = 0, []
    while len(got) < n:
        v = int(w[pos]) | (int(w[pos + 1]) << 32)
        pos += 2
        m = v * rng_range
        if (m & ((1 << 64) - 1)) <= zone:
            got.append(m >> 64)
    assert used == pos
    assert got == [int(x) for x in out]
    assert set(got) == set(range(rng_range))


def test_card_mapping(L):
    """rand 0.8.5 UniformInt<u8> through u32: card = 1 + hi32(v * 10), rejecting lo > 0xfffffff9."""
    seed, agent, n = 5, 1, 5000
    w = np.zeros(n + 64, np.uint32)
    L.oracle_stream_words(seed, agent, 0, len(w), O._p(w))
    out = np.zeros(n, np.float64)
    used = L.oracle_sample(seed, agent, 0, 2, 0, n, O._p(out))
    pos, got = 0, []
    while len(got) < n:
        m = int(w[pos]) * 10
        pos += 1
        if (m & 0xffffffff) <= 0xfffffff9:
            got.append(1 + (m >> 32))
    assert used == pos
    assert got == [int(x) for x in out]
    assert min(got) == 1 and max(got) == 10


def test_fxhash_ids(L):
    # SURVEY.md §8.2
    assert L.oracle_fxhash_blackjack(12, 1, 1) == 14677233788820433653
    assert L.oracle_fxhash_blackjack(20, 10, 0) == 10584062389149388782
    ids = {L.oracle_fxhash_blackjack(p, d, a) for p in range(32) for d in range(32) for a in range(2)}
    assert len(ids) == 2048   # collision-free over every (p<32, d<32, ace)
    dense = {L.oracle_blackjack_dense(p, d, a) for p in range(4, 32) for d in range(1, 27) for a in range(2)}
    assert dense == set(range(1456))


def test_portable_log(L):
    assert L.oracle_log(1.0) == 0.0
    assert L.oracle_log(2.0) == math.log(2.0)
    worst = 0.0
    mism = 0
    for t in list(range(1, 20001)) + [10 ** k for k in range(5, 16)] + [2 ** 52, 2 ** 53 - 1]:
        a, b = L.oracle_log(float(t)), math.log(float(t))
        if a != b:
            mism += 1
            worst = max(worst, abs(a - b) / math.ulp(b))
    assert worst <= 1.0          # never more than one ulp from the platform libm
    assert mism < 0.03 * 20000   # and identical for the vast majority of arguments


def test_ucb_overflow_constants(L):
    """upper_confidence_bound.rs:36 — ln(t)/f64::MIN_POSITIVE: 0 at t=1, finite up to t=54, +inf from t=55."""
    tiny = 2.2250738585072014e-308
    assert L.oracle_log(1.0) / tiny == 0.0
    assert math.isfinite(L.oracle_log(54.0) / tiny)
    assert abs(L.oracle_log(54.0) / tiny - 1.7927e308) / 1.7927e308 < 1e-4
    assert abs(L.oracle_log(2.0) / tiny - 3.115e307) / 3.115e307 < 1e-3
    assert math.isinf(L.oracle_log(55.0) / tiny)


def test_argmax_and_categorical(L):
    nan = float("nan")

    def am(v):
        a = np.array(v, np.float64)
        return L.oracle_argmax(O._p(a), len(a))

    assert am([1.0, 3.0, 3.0, 2.0]) == 1          # first strict max
    assert am([nan, 5.0, 7.0]) == 0               # NaN at index 0 is never beaten
    assert am([1.0, nan, 0.5]) == 0               # a NaN never wins
    assert am([0.0, 0.0, 0.0, 0.0]) == 0

    def cs(p, r):
        a = np.array(p, np.float64)
        return L.oracle_categorical_sample(O._p(a), len(a), r)

    assert cs([0.25, 0.25, 0.5], 0.0) == 0
    assert cs([0.25, 0.25, 0.5], 0.25) == 1       # b > random is strict
    assert cs([0.25, 0.25, 0.5], 0.9999) == 2
    assert cs([0.25, 0.25, 0.25], 0.9) == 0       # nothing exceeds the draw -> argmax of all-false = 0
    assert cs([0.0, 1.0, 0.0], 0.3) == 1


def test_frozen_lake_slip_thresholds(L):
    """cumulative [0x1.5555555555555p-2, 0x1.5555555555555p-1, 1.0]."""
    third = 1.0 / 3.0
    b1, b2, b3 = third, third + third, third + third + third
    assert b1.hex() == "0x1.5555555555555p-2" and b2.hex() == "0x1.5555555555555p-1" and b3 == 1.0
    assert math.ceil(b1 * 2 ** 52) == 1501199875790166
    assert math.ceil(b2 * 2 ** 52) == 3002399751580331
    p = np.zeros(64 * 4 * 3); s = np.zeros(64 * 4 * 3, np.uint32); r = np.zeros(64 * 4 * 3); t = np.zeros(64 * 4 * 3, np.uint8)
    L.oracle_frozen_lake_table(1, 1, O._p(p), O._p(s), O._p(r), O._p(t))
    p = p.reshape(64, 4, 3); s = s.reshape(64, 4, 3); r = r.reshape(64, 4, 3); t = t.reshape(64, 4, 3)
    # state 0, action 0 (LEFT): slip set [(0-1)%4 = 3 (UP), 0 (LEFT), 1 (DOWN)] -> stays, stays, moves to 8
    assert list(s[0, 0]) == [0, 0, 8] and np.all(p[0, 0] == third)
    # state 62 (row 7 col 6), action 2 (RIGHT): [DOWN, RIGHT, UP] -> 62, 63 (G, reward 1, terminal), 54 (H, terminal)
    assert list(s[62, 2]) == [62, 63, 54] and list(r[62, 2]) == [0.0, 1.0, 0.0] and list(t[62, 2]) == [0, 1, 1]
    # holes / goal: slot 0 = (1.0, s, 0, true)
    assert p[19, 1, 0] == 1.0 and s[19, 1, 0] == 19 and t[19, 1, 0] == 1 and p[19, 1, 1] == 0.0
    # non-slippery: one deterministic slot
    L.oracle_frozen_lake_table(1, 0, O._p(p.reshape(-1)), O._p(s.reshape(-1)), O._p(r.reshape(-1)), O._p(t.reshape(-1)))
    assert p[0, 2, 0] == 1.0 and s[0, 2, 0] == 1 and p[0, 2, 1] == 0.0


def test_taxi_table_and_start_distribution(L):
    s = np.zeros(3000, np.uint32); r = np.zeros(3000); t = np.zeros(3000, np.uint8); init = np.zeros(500)
    L.oracle_taxi_table(O._p(s), O._p(r), O._p(t), O._p(init))
    s = s.reshape(500, 6); r = r.reshape(500, 6); t = t.reshape(500, 6)
    valid = np.nonzero(init)[0]
    assert len(valid) == 300 and list(valid[:8]) == [1, 2, 3, 4, 6, 7, 8, 9] and valid[-1] == 494
    assert np.all(init[valid] == 1.0 / 300.0)
    cum = 0.0
    for v in init:
        cum += v
    assert cum.hex() == "0x1.fffffffffffddp-1"
    assert math.ceil(cum * 2 ** 52) == 4503599627370479
    assert 2 ** 52 - 4503599627370479 == 17                         # 17 draws fall through to state 0
    enc = lambda row, col, p, d: ((row * 5 + col) * 5 + p) * 4 + d
    # taxi at R (0,0), passenger at R (0), dest G (1): pickup succeeds
    st = enc(0, 0, 0, 1)
    assert s[st, 4] == enc(0, 0, 4, 1) and r[st, 4] == -1.0 and t[st, 4] == 0
    # carrying, at G (0,4), dest G: dropoff terminates with +20
    st = enc(0, 4, 4, 1)
    assert s[st, 5] == enc(0, 4, 1, 1) and r[st, 5] == 20.0 and t[st, 5] == 1
    # illegal pickup / dropoff: -10, no move
    st = enc(2, 2, 0, 1)
    assert s[st, 4] == st and r[st, 4] == -10.0 and s[st, 5] == st and r[st, 5] == -10.0
    # wall between (0,1) and (0,2): east from (0,1) blocked, west from (0,2) blocked
    assert s[enc(0, 1, 0, 1), 2] == enc(0, 1, 0, 1) and s[enc(0, 2, 0, 1), 3] == enc(0, 2, 0, 1)
    # free move east from (0,0)
    assert s[enc(0, 0, 0, 1), 2] == enc(0, 1, 0, 1)
    # south / north clamp
    assert s[enc(4, 0, 0, 1), 0] == enc(4, 0, 0, 1) and s[enc(0, 0, 0, 1), 1] == enc(0, 0, 0, 1)
    # dest_loc never changes along any transition
    assert np.all(s % 4 == (np.arange(500) % 4)[:, None])


def test_cliff_table(L):
    s = np.zeros(192, np.uint32); r = np.zeros(192); t = np.zeros(192, np.uint8)
    L.oracle_cliff_table(O._p(s), O._p(r), O._p(t))
    s = s.reshape(48, 4); r = r.reshape(48, 4); t = t.reshape(48, 4)
    assert s[36, 2] == 37 and r[36, 2] == -100.0 and t[36, 2] == 1     # stepping right from start: cliff, terminal (Q7)
    assert s[36, 3] == 24 and r[36, 3] == -1.0 and t[36, 3] == 0       # up
    assert s[36, 0] == 36 and s[36, 1] == 36                           # left/down clamp
    assert s[35, 1] == 47 and r[35, 1] == -1.0 and t[35, 1] == 1       # down into the goal
    assert s[25, 1] == 37 and r[25, 1] == -100.0 and t[25, 1] == 1


def test_epsilon_schedule():
    """SURVEY §8.2 / Q8 with the CLI defaults, through the oracle's selector (Taxi run, 100k episodes)."""
    cfg = O.make_config(O.ENV_TAXI)
    assert cfg.eps_decay.hex() == "0x1.4f8b588e368f1p-16"
    s = O.Session(cfg, 0)
    s.train(1, 10 ** 9)
    assert s.export()[2].epsilon == 0.99998
    s.close()


def test_taxi_default_run_profile():
    """C4-style run profile (SURVEY §6 [derived], Python RNG there, so statistical agreement only)."""
    cfg = O.make_config(O.ENV_TAXI)
    s = O.Session(cfg, 0)
    ret, ln, tds, tda = s.train(100000, 10000)
    q, counts, st = s.export()
    assert st.epsilon == 1.9999999281486693e-05            # sticks there, never reaches 0 (Q8)
    assert abs(ln.mean() - 27.3) < 0.5
    dec = [ln[i * 10000:(i + 1) * 10000].mean() for i in range(10)]
    assert abs(dec[0] - 91.6) < 2.0 and abs(dec[2] - 29.0) < 1.5 and abs(dec[9] - 13.1) < 0.3
    assert st.eval_steps > 0 and ln.max() <= 101            # Q2: truncation pseudo-step, max length max_steps + 1
    s.close()


def test_truncation_quirk():
    """Q2: at curr_step >= max_steps the env returns obs 0, terminated, reward 0 (-100 for Cliff) without moving."""
    for env, trunc_reward in ((O.ENV_TAXI, 0.0), (O.ENV_FROZEN_LAKE, 0.0), (O.ENV_CLIFF_WALKING, -100.0)):
        cfg = O.make_config(env, max_steps=3, slippery=0)
        s = O.Session(cfg, 0)
        assert s.env_step(0) is None                         # EnvNotReady before reset
        s.env_reset()
        safe = {O.ENV_TAXI: 1, O.ENV_FROZEN_LAKE: 0, O.ENV_CLIFF_WALKING: 0}[env]   # an action that cannot terminate here
        for _ in range(3):
            obs, rew, term = s.env_step(safe)
            assert not term
        obs, rew, term = s.env_step(safe)
        assert (obs, rew, term) == (0, trunc_reward, True)
        assert s.env_step(safe) is None                      # EnvNotReady after termination
        s.close()


def test_blackjack_rules():
    """Q5: scripted against the stream — cards are 1..10 uniform, ace flag from the first two cards only,
    bust obs uses the dealer's two-card score, `new()` deals 4 cards that reset() discards."""
    L = O.lib()
    seed = 0xB1AC
    for agent in range(40):
        cards = np.zeros(64, np.float64)
        L.oracle_sample(seed, agent, 0, 2, 0, 64, O._p(cards))
        cards = [int(c) for c in cards]
        cfg = O.make_config(O.ENV_BLACKJACK, seed=seed)
        s = O.Session(cfg, agent)
        obs = s.env_reset()
        c = cards[4:]                                        # the constructor's hand is discarded
        p, d = [c[0], c[1]], [c[2], c[3]]
        p_ace, d_ace = 1 in p, 1 in d
        score = lambda hand, ace: sum(hand) + (10 if ace and sum(hand) + 10 <= 21 else 0)
        assert obs == L.oracle_blackjack_dense(score(p, p_ace), d[0], int(p_ace))
        nxt = 4
        # hit until 17 or bust, then stick
        while True:
            if score(p, p_ace) < 17:
                p.append(c[nxt]); nxt += 1
                obs, rew, term = s.env_step(0)
                ps = score(p, p_ace)
                if ps > 21:
                    assert (obs, rew, term) == (L.oracle_blackjack_dense(ps, score(d, d_ace), int(p_ace)), -1.0, True)
                    break
                assert (obs, rew, term) == (L.oracle_blackjack_dense(ps, d[0], int(p_ace)), 0.0, False)
            else:
                while score(d, d_ace) < 17:
                    d.append(c[nxt]); nxt += 1
                obs, rew, term = s.env_step(1)
                ps, ds = score(p, p_ace), score(d, d_ace)
                want = 1.0 if ds > 21 else (1.0 if ps > ds else (-1.0 if ps < ds else 0.0))
                assert (obs, rew, term) == (L.oracle_blackjack_dense(ps, ds, int(p_ace)), want, True)
                break
        assert s.env_step(0) is None
        s.close()


def test_ucb_expected_sarsa_nan_quirk():
    """Q9: UCB + Expected Sarsa goes NaN once t >= 55 meets an unvisited action; sarsa/qlearning stay finite."""
    cfg = O.make_config(O.ENV_CLIFF_WALKING, selector=O.SEL_UCB, target=O.TARGET_EXPECTED_SARSA)
    s = O.Session(cfg, 0)
    s.train(8, 10 ** 9, ep_begin=1)                          # ep_begin=1 skips the evaluate(100) after episode 0
    te = s.training_error()
    first = int(np.nonzero(np.isnan(te))[0][0])
    assert first == 49                                       # global step 50 (t = 54 + 1 at that get_exploration_probs)
    s.close()
    for tgt in (O.TARGET_SARSA, O.TARGET_QLEARNING):
        cfg = O.make_config(O.ENV_CLIFF_WALKING, selector=O.SEL_UCB, target=tgt)
        s = O.Session(cfg, 0)
        s.train(30, 10)
        assert np.all(np.isfinite(s.training_error()))
        s.close()


def test_double_policy_alternation():
    """Q10: TD read from the flag-selected table, written to the other, flag flips every update; reset keeps it."""
    cfg = O.make_config(O.ENV_CLIFF_WALKING, policy=O.POLICY_DOUBLE, selector=O.SEL_UCB, target=O.TARGET_QLEARNING, lr=0.5)
    s = O.Session(cfg, 0)
    # first update: flag=true -> reads alpha (all 0), writes beta
    td = s.update(36, 3, -1.0, False, 24, 0)
    q, _, st = s.export()
    assert td == -1.0 and q[0, 36, 3] == 0.0 and q[1, 36, 3] == -0.5 and st.policy_flag == 0
    # second: flag=false -> reads beta, writes alpha
    td = s.update(36, 3, -1.0, False, 24, 0)
    q, _, st = s.export()
    assert td == -1.0 - (-0.5) and q[0, 36, 3] == 0.5 * td and st.policy_flag == 1
    s.update(36, 3, -1.0, False, 24, 0)
    s.agent_reset()
    q, _, st = s.export()
    assert np.all(q == 0.0) and st.policy_flag == 0         # reset() wipes the tables, not the flag
    s.close()


def test_trace_sweep_by_hand():
    """Q11: accumulating traces, every action of every visited state swept, decay gamma*lambda, cleared on termination."""
    lr, g, lam = 0.5, 0.9, 0.5
    cfg = O.make_config(O.ENV_CLIFF_WALKING, agent=O.AGENT_TRACES, target=O.TARGET_SARSA, lr=lr, gamma=g, lambda_=lam,
                        selector=O.SEL_UCB)
    s = O.Session(cfg, 0)
    td1 = s.update(36, 3, -1.0, False, 24, 2)               # e[36][3] = 1
    q, _, _ = s.export()
    assert td1 == -1.0 and q[0, 36, 3] == lr * td1 and np.count_nonzero(q) == 1
    td2 = s.update(24, 2, -1.0, False, 25, 2)               # e[24][2] = 1, e[36][3] = g*lam
    q, _, _ = s.export()
    assert td2 == -1.0
    assert q[0, 24, 2] == lr * td2
    assert q[0, 36, 3] == lr * td1 + lr * (td2 * (g * lam))
    td3 = s.update(25, 1, -100.0, True, 37, 0)              # terminal: traces cleared afterwards
    q3, _, _ = s.export()
    assert q3[0, 36, 3] == q[0, 36, 3] + lr * (td3 * ((g * lam) * (g * lam)))
    td4 = s.update(36, 3, -1.0, False, 24, 2)               # fresh episode: only (36,3) moves
    q4, _, _ = s.export()
    changed = np.argwhere(q4 != q3)
    assert [tuple(x) for x in changed] == [(0, 36, 3)]
    s.close()
