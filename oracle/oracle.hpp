// oracle/oracle.hpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A C++17 restatement of the hot path of JohnVithor/RL-Rust (paths below are under the
// reference's `src/`), with a counter-based Philox4x32-10 word stream injected at every
// `rand::thread_rng()` call site.  It is the checker for the CUDA engine in
// `rl-rust_b200/csrc/`: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
// cpu_baseline / `--impl reference` legs may build, load or call anything in `oracle/`.
// The product (`librlb.so`) never includes, links or calls this file.
//
// PARITY UNPINNED BY THE REFERENCE: the reference ships no tests, golden vectors or
// seedable RNG, and cannot be compiled in this image (no Rust toolchain).  The bits→value
// samplers restate rand 0.8.5 (`Uniform<f64>`, `Uniform<usize>`, `Uniform<u8>`), the
// Blackjack observation id restates fxhash 0.2.1 — both un-vendored dependencies
// (Cargo.toml:16,19).  What pins this oracle is (i) published KATs (Philox / Random123),
// (ii) the [derived] constants of SURVEY.md §8.2, (iii) an independent pure-Python
// restatement (`oracle/pyref.py`) cross-checked in tests/, (iv) NVIDIA's own Philox (cuRAND
// device API) producing the same word stream (tests/test_gpu_curand.py).  The reference-side
// half of a real pin — the Philox `RngCore` for the crate, the patch of its nine
// thread_rng() call sites, a dump binary and the test that compares — is in oracle/rust_ref/
// (source only: it needs `cargo`); tests/test_rust_ref.py runs when oracle/_ref/ is built.
//
// Arithmetic contract (compile with -ffp-contract=off; x86-64 SSE2, no x87):
//   Real = double  -> reference-faithful mode (the reference computes in f64 everywhere).
//   Real = float   -> fast mode: Q, traces, TD, lr, gamma, lambda, rewards in f32;
//                     epsilon state / explore test / UCB bonus math stay in f64.
//
// Every function cites the reference file:line it follows.
#pragma once
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace oracle {

using u8 = uint8_t;
using u16 = uint16_t;
using u32 = uint32_t;
using u64 = uint64_t;
using i64 = int64_t;

// ---------------------------------------------------------------------------------------
// RNG injection contract.  The reference draws everything from
// rand::thread_rng() (ChaCha12, OS-seeded, unseedable).  Replaced by: per-agent stream of
// 32-bit words w[n] = Philox4x32-10(key=(seed_lo,seed_hi), ctr=(blk_lo,blk_hi,agent_lo,
// agent_hi))[n&3], blk = n>>2; next_u64 = w[n] | w[n+1]<<32 (low word first, as
// rand_core::block::BlockRng::next_u64 does, including across a block boundary).
// ---------------------------------------------------------------------------------------
inline void philox4x32_10(const u32 ctr_in[4], const u32 key_in[2], u32 out[4]) {
    // Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3" (SC'11).
    const u32 M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    u32 c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    u32 k0 = key_in[0], k1 = key_in[1];
    for (int round = 0; round < 10; ++round) {
        u64 p0 = (u64)M0 * c0, p1 = (u64)M1 * c2;
        u32 n0 = (u32)(p1 >> 32) ^ c1 ^ k0;
        u32 n1 = (u32)p1;
        u32 n2 = (u32)(p0 >> 32) ^ c3 ^ k1;
        u32 n3 = (u32)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Stream {
    u64 seed = 0, agent = 0, n = 0;   // n = index of the next 32-bit word
    u32 blk[4] = {0, 0, 0, 0};
    u64 cur_blk = ~0ull;
    Stream() = default;
    Stream(u64 seed_, u64 agent_, u64 n_ = 0) : seed(seed_), agent(agent_), n(n_) {}
    u32 next_u32() {
        u64 b = n >> 2;
        if (b != cur_blk) {
            u32 ctr[4] = {(u32)b, (u32)(b >> 32), (u32)agent, (u32)(agent >> 32)};
            u32 key[2] = {(u32)seed, (u32)(seed >> 32)};
            philox4x32_10(ctr, key, blk);
            cur_blk = b;
        }
        return blk[(n++) & 3];
    }
    u64 next_u64() {
        u64 lo = next_u32();
        u64 hi = next_u32();
        return lo | (hi << 32);
    }
};

// rand 0.8.5 `UniformFloat<f64>::sample` for Uniform::from(0.0..1.0): 52 random mantissa
// bits -> [1,2) -> minus 1.0, times scale 1.0, plus low 0.0.  Call sites: env/taxi.rs:136-137,
// env/frozen_lake.rs:107-108,126, action_selection/uniform_epsilon_greed.rs:33,53.
inline double uniform_f64(Stream& rng) {
    u64 k = rng.next_u64() >> 12;
    return (double)k * 0x1p-52;   // exact: k < 2^52
}

// rand 0.8.5 `UniformInt<usize>::sample` for Uniform::from(0..range) on a 64-bit target:
// widening multiply + rejection zone.  Call site: uniform_epsilon_greed.rs:34,62.
inline u64 uniform_usize(Stream& rng, u64 range) {
    const u64 ints_to_reject = (UINT64_MAX - range + 1) % range;
    const u64 zone = UINT64_MAX - ints_to_reject;
    for (;;) {
        u64 v = rng.next_u64();
        unsigned __int128 m = (unsigned __int128)v * range;
        u64 hi = (u64)(m >> 64), lo = (u64)m;
        if (lo <= zone) return hi;
    }
}

// rand 0.8.5 `Rng::gen_range(0..range)` on usize (64-bit target) = `UniformInt<usize>::sample_single_inclusive(0,
// range-1)`: the same widening multiply, but the one-shot path skips the modulus and uses the conservative zone
// `(range << range.leading_zeros()) - 1` (rejects up to half of the draws; a 1-entry range still rejects v >= 2^63).
// Call site: model/random_model.rs:30.
inline u64 gen_range_usize(Stream& rng, u64 range) {
    const u64 zone = (range << __builtin_clzll(range)) - 1;
    for (;;) {
        u64 v = rng.next_u64();
        unsigned __int128 m = (unsigned __int128)v * range;
        u64 hi = (u64)(m >> 64), lo = (u64)m;
        if (lo <= zone) return hi;
    }
}

// rand 0.8.5 `UniformInt<u8>::sample` for Uniform::from(1..11): u8 is sampled through u32.
// Call site: env/blackjack.rs:55,76.
inline u8 uniform_card(Stream& rng) {
    const u32 range = 10;
    const u32 ints_to_reject = (UINT32_MAX - range + 1) % range;   // 6
    const u32 zone = UINT32_MAX - ints_to_reject;                  // 0xfffffff9
    for (;;) {
        u32 v = rng.next_u32();
        u64 m = (u64)v * range;
        u32 hi = (u32)(m >> 32), lo = (u32)m;
        if (lo <= zone) return (u8)(1 + hi);
    }
}

// fxhash 0.2.1, 64-bit FxHasher: h = (rotl(h,5) ^ word) * SEED per written word.  The
// derived Hash of BlackJackObservation writes p_score:u8, d_score:u8, p_ace:bool(as u8).
// env/blackjack.rs:10-28.
inline u64 fxhash_blackjack(u8 p_score, u8 d_score, bool p_ace) {
    const u64 SEED = 0x517cc1b727220a95ull;
    u64 h = 0;
    const u8 bytes[3] = {p_score, d_score, (u8)(p_ace ? 1 : 0)};
    for (u8 b : bytes) {
        h = ((h << 5) | (h >> 59)) ^ (u64)b;
        h *= SEED;
    }
    return h;
}

// Natural log used by UCB (upper_confidence_bound.rs:36,56 call f64::ln).  Rust's ln is the
// platform libm's; neither it nor CUDA's log can be pinned bit-for-bit here, so the
// contract fixes ONE algorithm for both sides: the classic fdlibm reduction
// (x = 2^k (1+f), s = f/(2+f), log(1+f) = 2s + s*R(s^2), 14-term minimax split in two
// chains), evaluated with plain round-to-nearest +,-,*,/ and no FMA.  < 1 ulp.  Only
// finite x >= 1 is needed (x = t as f64, t >= 1).
inline double portable_log(double x);
// `(self.t as f64).ln()` (upper_confidence_bound.rs:36,56).  Rust's f64::ln is the platform libm's log; the engine and
// the oracle share ONE portable routine instead so that host and device agree bit for bit.  Building
// with -DORACLE_LIBM_LOG swaps in this machine's libm (glibc) — used only by tools/log_sensitivity.py to measure how
// far the choice of log reaches into UCB trajectories.
inline double ucb_ln(double x) {
#ifdef ORACLE_LIBM_LOG
    return std::log(x);
#else
    return portable_log(x);
#endif
}
inline double portable_log(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                 Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                 Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    u64 bits;
    std::memcpy(&bits, &x, 8);
    int32_t hx = (int32_t)(bits >> 32);
    u32 lx = (u32)bits;
    int32_t k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int32_t i = (hx + 0x95f64) & 0x100000;
    u64 nb = ((u64)(u32)(hx | (i ^ 0x3ff00000)) << 32) | lx;   // normalise x or x/2
    std::memcpy(&x, &nb, 8);
    k += (i >> 20);
    double f = x - 1.0;
    double dk = (double)k;
    if ((0x000fffff & (2 + hx)) < 3) {   // |f| < 2^-20
        if (f == 0.0) {
            if (k == 0) return 0.0;
            return dk * ln2_hi + dk * ln2_lo;
        }
        double R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        return dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    double s = f / (2.0 + f);
    double z = s * s;
    i = hx - 0x6147a;
    double w = z * z;
    int32_t j = 0x6b851 - hx;
    double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    double R = t2 + t1;
    if (i > 0) {
        double hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    }
    if (k == 0) return f - s * (f - R);
    return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

// ---------------------------------------------------------------------------------------
// utils.rs
// ---------------------------------------------------------------------------------------
// utils.rs:1-11 — first index of the strict maximum (PartialOrd `>`; a NaN never wins,
// a NaN at index 0 is never beaten).
template <class V>
inline size_t argmax(const V* vec, size_t len) {
    size_t result = 0;
    V best = vec[0];
    for (size_t i = 0; i < len; ++i) {
        if (vec[i] > best) { best = vec[i]; result = i; }
    }
    return result;
}
// utils.rs:13-21
template <class V>
inline V max_of(const V* vec, size_t len) {
    V best = vec[0];
    for (size_t i = 0; i < len; ++i) {
        if (vec[i] > best) best = vec[i];
    }
    return best;
}
// utils.rs:33-43 — builds a Vec<bool> of "running sum > random" and returns argmax of it
// (first true; 0 if none).  The heap-allocated vector is kept: it is part of what the
// reference costs per call (per FrozenLake step, per Taxi reset).
inline size_t categorical_sample(const double* probs, size_t len, double random) {
    double b = 0.0;
    std::vector<u8> r;
    r.reserve(len);
    for (size_t i = 0; i < len; ++i) {
        b += probs[i];
        r.push_back(b > random ? 1 : 0);
    }
    return argmax<u8>(r.data(), r.size());
}
// utils.rs:45-47
inline size_t from_2d_to_1d(size_t ncol, size_t row, size_t col) { return row * ncol + col; }
// utils.rs:53-76 — 0 left, 1 down, 2 right, 3 up, clamped; anything else: stay.
inline void inc(size_t nrow, size_t ncol, size_t row, size_t col, size_t a, size_t& nr, size_t& nc) {
    if (a == 0) { nc = col != 0 ? col - 1 : 0; nr = row; }
    else if (a == 1) { nc = col; nr = (row + 1 < nrow - 1) ? row + 1 : nrow - 1; }
    else if (a == 2) { nc = (col + 1 < ncol - 1) ? col + 1 : ncol - 1; nr = row; }
    else if (a == 3) { nc = col; nr = row != 0 ? row - 1 : 0; }
    else { nr = row; nc = col; }
}

// ---------------------------------------------------------------------------------------
// env.rs:19-49 — trait Env<T, COUNT>.  T = usize for all four in-scope envs.
// step() returns false for Err(EnvNotReady) (env.rs:16-17).
// ---------------------------------------------------------------------------------------
struct StepResult { u64 obs; double reward; bool terminated; };

template <int A>
struct Env {
    virtual ~Env() {}
    size_t action_size() const { return A; }
    virtual u64 reset() = 0;
    virtual bool step(size_t action, StepResult& out) = 0;
    // dense index of an observation (identity except Blackjack) — oracle-side helper for
    // table export, not a reference method.
    virtual u32 dense_index(u64 obs) const { return (u32)obs; }
    virtual u32 n_states() const = 0;
};

// env/blackjack.rs:30-163
struct BlackJackEnv : Env<2> {
    bool ready = false;
    u8 player[16], dealer[16];
    size_t player_i = 0, dealer_i = 0;
    bool player_has_ace = false, dealer_has_ace = false;
    Stream* rng;
    explicit BlackJackEnv(Stream* rng_) : rng(rng_) {   // blackjack.rs:45-59: new() deals a hand
        std::memset(player, 0, 16);
        std::memset(dealer, 0, 16);
        initialize_hands();
    }
    u8 get_new_card() { return uniform_card(*rng); }   // :75-77
    void initialize_hands() {                          // :60-69
        player[0] = get_new_card();
        player[1] = get_new_card();
        player_i = 2;
        dealer[0] = get_new_card();
        dealer[1] = get_new_card();
        dealer_i = 2;
        player_has_ace = (player[0] == 1) || (player[1] == 1);
        dealer_has_ace = (dealer[0] == 1) || (dealer[1] == 1);
    }
    u8 compute_player_score() const {                  // :79-86
        u8 score = 0;
        for (u8 c : player) score = (u8)(score + c);
        return (player_has_ace && score + 10 <= 21) ? (u8)(score + 10) : score;
    }
    u8 compute_dealer_score() const {                  // :88-95
        u8 score = 0;
        for (u8 c : dealer) score = (u8)(score + c);
        return (dealer_has_ace && score + 10 <= 21) ? (u8)(score + 10) : score;
    }
    u64 reset() override {                             // :105-116
        std::memset(player, 0, 16);
        std::memset(dealer, 0, 16);
        initialize_hands();
        u64 id = fxhash_blackjack(compute_player_score(), dealer[0], player_has_ace);
        ready = true;
        return id;
    }
    bool step(size_t action, StepResult& out) override {   // :118-163
        if (!ready) return false;
        if (action == 0) {
            if (player_i >= 16) return false;   // reference would panic (index out of bounds); p ~ 1e-17
            player[player_i] = get_new_card();
            player_i += 1;
            u8 p_score = compute_player_score();
            if (p_score > 21) {
                ready = false;
                out = {fxhash_blackjack(p_score, compute_dealer_score(), player_has_ace), -1.0, true};
                return true;
            }
            out = {fxhash_blackjack(p_score, dealer[0], player_has_ace), 0.0, false};
            return true;
        }
        ready = false;
        u8 d_score = compute_dealer_score();
        while (d_score < 17) {
            if (dealer_i >= 16) return false;
            dealer[dealer_i] = get_new_card();
            dealer_i += 1;
            d_score = compute_dealer_score();
        }
        u8 p_score = compute_player_score();
        u64 id = fxhash_blackjack(p_score, d_score, player_has_ace);
        if (d_score > 21) { out = {id, 1.0, true}; return true; }
        double reward = p_score > d_score ? 1.0 : (p_score < d_score ? -1.0 : 0.0);
        out = {id, reward, true};
        return true;
    }
    // dense index: p in [4,31], d in [1,26], ace in {0,1} -> 28*26*2 = 1456
    static u32 dense_of(u8 p, u8 d, bool ace) { return ((u32)(p - 4) * 26u + (u32)(d - 1)) * 2u + (ace ? 1u : 0u); }
    static const std::vector<u64>& id_table() {
        static std::vector<u64> t = [] {
            std::vector<u64> v(1456);
            for (u32 p = 4; p <= 31; ++p) for (u32 d = 1; d <= 26; ++d) for (u32 a = 0; a < 2; ++a)
                v[dense_of((u8)p, (u8)d, a != 0)] = fxhash_blackjack((u8)p, (u8)d, a != 0);
            return v;
        }();
        return t;
    }
    u32 dense_index(u64 obs) const override {
        const auto& t = id_table();
        for (u32 i = 0; i < t.size(); ++i) if (t[i] == obs) return i;
        return 0xffffffffu;
    }
    u32 n_states() const override { return 1456; }
};

// env/frozen_lake.rs:12-134
struct FrozenLakeEnv : Env<4> {
    struct Transition { double p; size_t s; double r; bool t; };
    bool ready = false;
    std::vector<double> initial_state_distrib;
    std::vector<std::array<std::array<Transition, 3>, 4>> probs;
    size_t player_pos = 0;
    u64 max_steps, curr_step = 0;
    Stream* rng;
    static std::vector<std::string> map_4x4() { return {"SFFF", "FHFH", "FFFH", "HFFG"}; }   // :23
    static std::vector<std::string> map_8x8() {                                               // :25-28
        return {"SFFFFFFF", "FFFFFFFF", "FFFHFFFF", "FFFFFHFF", "FFFHFFFF", "FHHFFFHF", "FHFFHFHF", "FFFHFFFG"};
    }
    static void update_probability_matrix(const std::vector<std::string>& map, size_t nrow, size_t ncol,
                                          size_t row, size_t col, size_t action, size_t& ns, double& r, bool& t) {
        size_t newrow, newcol;                                                               // :32-46
        inc(nrow, ncol, row, col, action, newrow, newcol);
        ns = from_2d_to_1d(ncol, newrow, newcol);
        char letter = map[newrow][newcol];
        t = (letter == 'G' || letter == 'H');
        r = (letter == 'G') ? 1.0 : 0.0;
    }
    FrozenLakeEnv(const std::vector<std::string>& map, bool is_slippery, u64 max_steps_, Stream* rng_)
        : max_steps(max_steps_), rng(rng_) {                                                 // :48-102
        size_t nrow = map.size(), ncol = map[0].size();
        std::string flat;
        for (auto& row : map) flat += row;
        initial_state_distrib.assign(flat.size(), 0.0);
        size_t counter = 0;
        std::vector<size_t> pos;
        for (size_t i = 0; i < flat.size(); ++i) if (flat[i] == 'S') { counter += 1; pos.push_back(i); }
        for (size_t i : pos) initial_state_distrib[i] = 1.0 / (double)counter;
        Transition zero{0.0, 0, 0.0, false};
        std::array<Transition, 3> z3{zero, zero, zero};
        probs.assign(nrow * ncol, {z3, z3, z3, z3});
        for (size_t row = 0; row < nrow; ++row) for (size_t col = 0; col < ncol; ++col) {
            size_t s = from_2d_to_1d(ncol, row, col);
            for (size_t a = 0; a < 4; ++a) {
                auto& li = probs[s][a];
                char letter = map[row][col];
                if (letter == 'G' || letter == 'H') {
                    li[0] = {1.0, s, 0.0, true};
                } else if (is_slippery) {
                    // :78 — usize arithmetic, wrapping in release: (0-1)%4 == 3
                    size_t cand[3] = {(size_t)((a - (size_t)1) % 4), a, (a + 1) % 4};
                    for (size_t i = 0; i < 3; ++i) {
                        size_t ns; double r; bool t;
                        update_probability_matrix(map, nrow, ncol, row, col, cand[i], ns, r, t);
                        li[i] = {1.0 / 3.0, ns, r, t};
                    }
                } else {
                    size_t ns; double r; bool t;
                    update_probability_matrix(map, nrow, ncol, row, col, a, ns, r, t);
                    li[0] = {1.0, ns, r, t};
                }
            }
        }
    }
    u64 reset() override {                                                                   // :106-113
        double random = uniform_f64(*rng);
        std::vector<double> copy = initial_state_distrib;   // `.to_vec()` :109
        player_pos = categorical_sample(copy.data(), copy.size(), random);
        ready = true;
        curr_step = 0;
        return player_pos;
    }
    bool step(size_t action, StepResult& out) override {                                     // :115-134
        if (!ready) return false;
        if (curr_step >= max_steps) { ready = false; out = {0, 0.0, true}; return true; }
        curr_step += 1;
        const auto& transitions = probs[player_pos][action];
        double t_probs[3] = {transitions[0].p, transitions[1].p, transitions[2].p};
        double random = uniform_f64(*rng);
        size_t i = categorical_sample(t_probs, 3, random);
        const Transition& tr = transitions[i];
        player_pos = tr.s;
        if (tr.t) ready = false;
        out = {tr.s, tr.r, tr.t};
        return true;
    }
    u32 n_states() const override { return (u32)probs.size(); }
};

// env/cliff_walking.rs:6-89
struct CliffWalkingEnv : Env<4> {
    struct Cell { size_t s; double r; bool t; };
    bool ready = false;
    Cell obs[48][4];
    size_t player_pos = 0;
    u64 max_steps, curr_step = 0;
    static constexpr size_t START_POSITION = 36, GOAL_POSITION = 47;   // :16-18
    static Cell update_probability_matrix(size_t row, size_t col, size_t action) {   // :22-29
        size_t nr, nc;
        inc(4, 12, row, col, action, nr, nc);
        size_t ns = from_2d_to_1d(12, nr, nc);
        bool win = ns == GOAL_POSITION;
        bool lose = ns >= 37 && ns <= 46;   // CLIFF_POSITIONS :17
        return {ns, lose ? -100.0 : -1.0, lose || win};
    }
    explicit CliffWalkingEnv(u64 max_steps_) : max_steps(max_steps_) {   // :31-63
        for (size_t row = 0; row < 4; ++row) for (size_t col = 0; col < 12; ++col) for (size_t a = 0; a < 4; ++a)
            obs[from_2d_to_1d(12, row, col)][a] = update_probability_matrix(row, col, a);
    }
    u64 reset() override { player_pos = START_POSITION; ready = true; curr_step = 0; return player_pos; }   // :67-72
    bool step(size_t action, StepResult& out) override {   // :74-89
        if (!ready) return false;
        if (curr_step >= max_steps) { ready = false; out = {0, -100.0, true}; return true; }
        curr_step += 1;
        Cell c = obs[player_pos][action];
        player_pos = c.s;
        if (c.t) ready = false;
        out = {c.s, c.r, c.t};
        return true;
    }
    u32 n_states() const override { return 48; }
};

// env/taxi.rs:10-159
struct TaxiEnv : Env<6> {
    struct Cell { size_t s; double r; bool t; };
    bool ready = false;
    double initial_state_distrib[500];
    Cell obs[500][6];
    size_t curr_obs = 0;
    u64 max_steps, curr_step = 0;
    Stream* rng;
    static size_t encode(size_t taxi_row, size_t taxi_col, size_t pass_loc, size_t dest_loc) {   // :33-42
        return ((taxi_row * 5 + taxi_col) * 5 + pass_loc) * 4 + dest_loc;
    }
    TaxiEnv(u64 max_steps_, Stream* rng_) : max_steps(max_steps_), rng(rng_) {   // :57-131
        static const char* MAP[7] = {"+---------+", "|R: | : :G|", "| : | : : |", "| : : : : |",
                                     "| | : | : |", "|Y| : |B: |", "+---------+"};   // :21-29
        static const size_t LOCS[4][2] = {{0, 0}, {0, 4}, {4, 0}, {4, 3}};              // :30
        for (double& v : initial_state_distrib) v = 0.0;
        double sum = 0.0;
        for (size_t row = 0; row < 5; ++row) for (size_t col = 0; col < 5; ++col)
        for (size_t pass_loc = 0; pass_loc < 5; ++pass_loc) for (size_t dest_loc = 0; dest_loc < 4; ++dest_loc) {
            size_t state = encode(row, col, pass_loc, dest_loc);
            if (pass_loc < 4 && pass_loc != dest_loc) { initial_state_distrib[state] += 1.0; sum += 1.0; }
            for (size_t action = 0; action < 6; ++action) {
                size_t new_row = row, new_col = col, new_pass_loc = pass_loc;
                double reward = -1.0;
                bool terminated = false;
                if (action == 0) new_row = (row + 1 < 4) ? row + 1 : 4;
                else if (action == 1) new_row = row != 0 ? row - 1 : 0;
                if (action == 2 && MAP[1 + row][2 * col + 2] == ':') {
                    new_col = (col + 1 < 4) ? col + 1 : 4;
                } else if (action == 3 && MAP[1 + row][2 * col] == ':') {
                    new_col = col != 0 ? col - 1 : 0;
                } else if (action == 4) {
                    if (pass_loc < 4 && row == LOCS[pass_loc][0] && col == LOCS[pass_loc][1]) new_pass_loc = 4;
                    else reward = -10.0;
                } else if (action == 5) {
                    if (row == LOCS[dest_loc][0] && col == LOCS[dest_loc][1] && pass_loc == 4) {
                        new_pass_loc = dest_loc; terminated = true; reward = 20.0;
                    } else reward = -10.0;
                }
                obs[state][action] = {encode(new_row, new_col, new_pass_loc, dest_loc), reward, terminated};
            }
        }
        for (double& v : initial_state_distrib) v /= sum;
    }
    u64 reset() override {   // :135-142
        double random = uniform_f64(*rng);
        curr_obs = categorical_sample(initial_state_distrib, 500, random);
        ready = true;
        curr_step = 0;
        return curr_obs;
    }
    bool step(size_t action, StepResult& out) override {   // :144-159
        if (!ready) return false;
        if (curr_step >= max_steps) { ready = false; out = {0, 0.0, true}; return true; }
        curr_step += 1;
        Cell c = obs[curr_obs][action];
        curr_obs = c.s;
        if (c.t) ready = false;
        out = {c.s, c.r, c.t};
        return true;
    }
    u32 n_states() const override { return 500; }
};

// ---------------------------------------------------------------------------------------
// FxHashMap stand-in: open-addressing u64 -> V with the Fx multiply hash.  Iteration order
// differs from hashbrown's, which is unobservable: the only iterated map is the trace map
// (elegibility_traces_agent.rs:86) and each (obs,action) cell is touched once per sweep.
// ---------------------------------------------------------------------------------------
template <class V>
struct FxMap {
    std::vector<u64> keys;
    std::vector<V> vals;
    std::vector<u8> used;
    size_t count = 0, mask = 0;
    FxMap() { rehash(16); }
    static size_t hash(u64 k) { return (size_t)((k * 0x517cc1b727220a95ull) >> 20); }
    void rehash(size_t cap) {
        std::vector<u64> ok = std::move(keys);
        std::vector<V> ov = std::move(vals);
        std::vector<u8> ou = std::move(used);
        keys.assign(cap, 0); vals.assign(cap, V()); used.assign(cap, 0);
        mask = cap - 1; count = 0;
        for (size_t i = 0; i < ou.size(); ++i) if (ou[i]) *insert_slot(ok[i]) = ov[i];
    }
    V* insert_slot(u64 k) {
        size_t i = hash(k) & mask;
        while (used[i]) { if (keys[i] == k) return &vals[i]; i = (i + 1) & mask; }
        used[i] = 1; keys[i] = k; ++count;
        return &vals[i];
    }
    const V* get(u64 k) const {   // HashMap::get
        size_t i = hash(k) & mask;
        while (used[i]) { if (keys[i] == k) return &vals[i]; i = (i + 1) & mask; }
        return nullptr;
    }
    V& entry_or_insert(u64 k, const V& dflt) {   // .entry(k).or_insert(dflt)
        if ((count + 1) * 4 > (mask + 1) * 3) rehash((mask + 1) * 2);
        size_t before = count;
        V* p = insert_slot(k);
        if (count != before) *p = dflt;
        return *p;
    }
    void clear() { keys.clear(); vals.clear(); used.clear(); count = 0; rehash(16); }   // = FxHashMap::default()
    template <class F> void for_each(F&& f) { for (size_t i = 0; i <= mask; ++i) if (used[i]) f(keys[i], vals[i]); }
};

// ---------------------------------------------------------------------------------------
// policy.rs:15-25 — trait Policy<T, COUNT>
// ---------------------------------------------------------------------------------------
template <int A, class Real>
struct Policy {
    using Row = std::array<Real, A>;
    virtual ~Policy() {}
    virtual Row predict(u64 obs) = 0;
    virtual Row get_values(u64 obs) = 0;
    virtual Real update(u64 obs, size_t action, u64 next_obs, Real temporal_difference) = 0;
    virtual void reset() = 0;
    virtual void after_update() = 0;
};

// policy/tabular_policy.rs:8-44 — "Basic"
template <int A, class Real>
struct TabularPolicy : Policy<A, Real> {
    using Row = std::array<Real, A>;
    Real learning_rate;
    Row dflt;
    FxMap<Row> policy;
    TabularPolicy(Real lr, Real default_value) : learning_rate(lr) { dflt.fill(default_value); }   // :15-21
    Row predict(u64 obs) override { const Row* r = policy.get(obs); return r ? *r : dflt; }        // :27-29
    Row get_values(u64 obs) override { const Row* r = policy.get(obs); return r ? *r : dflt; }     // :31-33
    Real update(u64 obs, size_t action, u64, Real td) override {                                   // :35-38
        Row& row = policy.entry_or_insert(obs, dflt);
        row[action] += learning_rate * td;
        return learning_rate * td;
    }
    void reset() override { policy.clear(); }   // :40-42
    void after_update() override {}             // :44
};

// policy/double_tabular_policy.rs:8-67
template <int A, class Real>
struct DoubleTabularPolicy : Policy<A, Real> {
    using Row = std::array<Real, A>;
    Real learning_rate;
    Row dflt;
    FxMap<Row> alpha_policy, beta_policy;
    bool policy_flag = true;                                                                       // :23
    DoubleTabularPolicy(Real lr, Real default_value) : learning_rate(lr) { dflt.fill(default_value); }
    Row predict(u64 obs) override {                                                                // :31-39
        Row values;
        values.fill((Real)0);
        const Row* a = alpha_policy.get(obs);
        const Row* b = beta_policy.get(obs);
        const Row& av = a ? *a : dflt;
        const Row& bv = b ? *b : dflt;
        for (int i = 0; i < A; ++i) values[i] = (av[i] + bv[i]) / (Real)2.0;
        return values;
    }
    Row get_values(u64 obs) override {                                                             // :41-48
        const Row* r = (policy_flag ? alpha_policy : beta_policy).get(obs);
        return r ? *r : dflt;
    }
    Real update(u64 obs, size_t action, u64, Real td) override {                                   // :50-58
        Row& row = (policy_flag ? beta_policy : alpha_policy).entry_or_insert(obs, dflt);
        row[action] += learning_rate * td;
        return learning_rate * td;
    }
    void reset() override { alpha_policy.clear(); beta_policy.clear(); }   // :60-63 (flag kept)
    void after_update() override { policy_flag = !policy_flag; }           // :65-67
};

// ---------------------------------------------------------------------------------------
// action_selection.rs:10-15 — trait ActionSelection<T, COUNT>
// Probabilities are produced in f64 (the selector's own type) and narrowed to Real by
// the caller in float mode.
// ---------------------------------------------------------------------------------------
template <int A, class Real>
struct ActionSelection {
    using Row = std::array<Real, A>;
    virtual ~ActionSelection() {}
    virtual size_t get_action(u64 obs, const Row& values) = 0;
    virtual void update() = 0;
    virtual Row get_exploration_probs(u64 obs, const Row& values) = 0;
    virtual void reset() = 0;
    virtual std::unique_ptr<ActionSelection> clone() const = 0;   // #[derive(Clone)]
};

enum DecayKind { DECAY_SUB = 0, DECAY_MUL = 1 };

// action_selection/uniform_epsilon_greed.rs:8-80.  The `Rc<dyn Fn(f64)->f64>` decay
// closure is replaced by {kind,param}: the bins only ever pass `a - k` (bin/taxi.rs:132)
// or `a * k` (bin/frozen_lake_neural.rs:181).
template <int A, class Real>
struct UniformEpsilonGreed : ActionSelection<A, Real> {
    using Row = std::array<Real, A>;
    double initial_epsilon, epsilon;
    int decay_kind; double decay_param;
    double final_epsilon;
    Stream* rng;
    UniformEpsilonGreed(double eps, int kind, double param, double final_eps, Stream* rng_)
        : initial_epsilon(eps), epsilon(eps), decay_kind(kind), decay_param(param), final_epsilon(final_eps), rng(rng_) {}
    void decay_epsilon() {                                                                          // :42-49
        double new_epsilon = decay_kind == DECAY_SUB ? epsilon - decay_param : epsilon * decay_param;
        epsilon = (final_epsilon > new_epsilon) ? epsilon : new_epsilon;
    }
    bool should_explore() { return epsilon != 0.0 && uniform_f64(*rng) < epsilon; }               // :51-54
    size_t get_action(u64, const Row& values) override {                                            // :60-66
        if (should_explore()) return (size_t)uniform_usize(*rng, A);
        return argmax<Real>(values.data(), A);
    }
    void update() override { decay_epsilon(); }                                                     // :68-70
    Row get_exploration_probs(u64, const Row& values) override {                                    // :72-76
        Row policy_probs;
        policy_probs.fill((Real)(epsilon / (double)A));
        policy_probs[argmax<Real>(values.data(), A)] = (Real)(1.0 - epsilon);
        return policy_probs;
    }
    void reset() override { epsilon = initial_epsilon; }                                            // :78-80
    std::unique_ptr<ActionSelection<A, Real>> clone() const override {
        return std::unique_ptr<ActionSelection<A, Real>>(new UniformEpsilonGreed(*this));
    }
};

// action_selection/upper_confidence_bound.rs:9-68.  Counters are u128 in the reference;
// u64 here (a run would need > 1.8e19 steps to tell).
template <int A, class Real>
struct UpperConfidenceBound : ActionSelection<A, Real> {
    using Row = std::array<Real, A>;
    using Cnt = std::array<u64, A>;
    FxMap<Cnt> action_counter;
    u64 t = 1;
    double confidence_level;
    explicit UpperConfidenceBound(double c) : confidence_level(c) {}
    size_t get_action(u64 obs, const Row& values) override {                                        // :29-42
        Cnt zero; zero.fill(0);
        Cnt& obs_actions = action_counter.entry_or_insert(obs, zero);
        double ucbs[A];
        for (int i = 0; i < A; ++i)
            ucbs[i] = (double)values[i] +
                      confidence_level * std::sqrt(ucb_ln((double)t) / ((double)obs_actions[i] + DBL_MIN));
        size_t action = argmax<double>(ucbs, A);
        obs_actions[action] += 1;
        t += 1;
        return action;
    }
    void update() override {}                                                                       // :44-46
    Row get_exploration_probs(u64 obs, const Row& values) override {                                // :48-63
        Cnt zero; zero.fill(0);
        Cnt& obs_actions = action_counter.entry_or_insert(obs, zero);
        double ucbs[A];
        double sum = 0.0;
        for (int i = 0; i < A; ++i) {
            ucbs[i] = (double)values[i] +
                      confidence_level * std::sqrt(ucb_ln((double)t) / ((double)obs_actions[i] + DBL_MIN));
            sum += ucbs[i];
        }
        Row out;
        for (int i = 0; i < A; ++i) out[i] = (Real)(ucbs[i] / sum);
        return out;
    }
    void reset() override { action_counter.clear(); t = 1; }                                        // :65-68
    std::unique_ptr<ActionSelection<A, Real>> clone() const override {
        return std::unique_ptr<ActionSelection<A, Real>>(new UpperConfidenceBound(*this));
    }
};

// ---------------------------------------------------------------------------------------
// agent.rs:17-45 — GetNextQValue and the three bootstrap targets
// ---------------------------------------------------------------------------------------
enum TargetKind { TARGET_SARSA = 0, TARGET_QLEARNING = 1, TARGET_EXPECTED_SARSA = 2 };

template <int A, class Real>
using GetNextQValue = Real (*)(const std::array<Real, A>&, size_t, const std::array<Real, A>&);

template <int A, class Real>
Real sarsa(const std::array<Real, A>& next_q, size_t next_action, const std::array<Real, A>&) {   // :19-25
    return next_q[next_action];
}
template <int A, class Real>
Real qlearning(const std::array<Real, A>& next_q, size_t, const std::array<Real, A>&) {            // :27-33
    return max_of<Real>(next_q.data(), A);
}
template <int A, class Real>
Real expected_sarsa(const std::array<Real, A>& next_q, size_t, const std::array<Real, A>& probs) { // :35-45
    Real future_q_value = (Real)0.0;
    for (int i = 0; i < A; ++i) future_q_value += probs[i] * next_q[i];
    return future_q_value;
}

// One record per env transition, for step-level parity (not a reference structure).
struct TrajRecord {
    u8 kind;        // 0 = reset (+ first get_action), 1 = train step, 2 = evaluate step
    u8 action;      // action chosen on `obs`
    u8 terminated;
    u8 pad = 0;
    u32 obs;        // dense index of the observation returned by reset/step
    double reward;
    double td;      // temporal difference (train steps only)
};

// ---------------------------------------------------------------------------------------
// agent.rs:47-164 — trait Agent<T, COUNT> with its default train/evaluate loops
// ---------------------------------------------------------------------------------------
template <int A, class Real>
struct Agent {
    using Row = std::array<Real, A>;
    std::vector<TrajRecord>* recorder = nullptr;   // oracle-side tap
    u64 eval_steps = 0;                            // steps spent inside evaluate (reported separately)
    double last_td_ = 0.0;
    virtual ~Agent() {}
    virtual void set_future_q_value_func(GetNextQValue<A, Real> f) = 0;
    virtual void set_action_selector(std::unique_ptr<ActionSelection<A, Real>> s) = 0;
    virtual size_t get_action(u64 obs) = 0;
    virtual Real update(u64 curr_obs, size_t curr_action, Real reward, bool terminated, u64 next_obs, size_t next_action) = 0;
    virtual void reset() = 0;

    void rec(Env<A>& env, u8 kind, u64 obs, size_t action, bool term, double reward, double td) {
        if (recorder) recorder->push_back({kind, (u8)action, (u8)(term ? 1 : 0), 0, env.dense_index(obs), reward, td});
    }

    // agent.rs:66-118.  Returns false if the env reported EnvNotReady (`.unwrap()` panic).
    bool train(Env<A>& env, u64 n_episodes, u64 eval_at, std::vector<Real>& reward_history,
               std::vector<u64>& episode_length, std::vector<Real>& training_error) {
        return train_range(env, 0, n_episodes, eval_at, reward_history, episode_length, training_error);
    }
    // Episodes [ep_begin, ep_end) of a train() call: same loop, global episode index in the
    // `episode % eval_at` test, so a run can be executed in chunks.
    bool train_range(Env<A>& env, u64 ep_begin, u64 ep_end, u64 eval_at, std::vector<Real>& reward_history,
                     std::vector<u64>& episode_length, std::vector<Real>& training_error) {
        for (u64 episode = ep_begin; episode < ep_end; ++episode) {
            u64 action_counter = 0;
            Real epi_reward = (Real)0.0;
            u64 curr_obs = env.reset();
            size_t curr_action = get_action(curr_obs);
            rec(env, 0, curr_obs, curr_action, false, 0.0, 0.0);
            for (;;) {
                action_counter += 1;
                StepResult sr;
                if (!env.step(curr_action, sr)) return false;
                size_t next_action = get_action(sr.obs);
                Real td = update(curr_obs, curr_action, (Real)sr.reward, sr.terminated, sr.obs, next_action);
                training_error.push_back(td);
                rec(env, 1, sr.obs, next_action, sr.terminated, sr.reward, (double)td);
                curr_obs = sr.obs;
                curr_action = next_action;
                epi_reward += (Real)sr.reward;
                if (sr.terminated) { reward_history.push_back(epi_reward); break; }
            }
            if (episode % eval_at == 0) {   // :107-113 (eval_at == 0 panics in the reference)
                std::vector<Real> r; std::vector<u64> l;
                if (!evaluate(env, 100, r, l)) return false;
            }
            episode_length.push_back(action_counter);
        }
        return true;
    }
    // agent.rs:120-141
    bool evaluate(Env<A>& env, u64 n_episodes, std::vector<Real>& reward_history, std::vector<u64>& episode_length) {
        for (u64 ep = 0; ep < n_episodes; ++ep) {
            u64 action_counter = 0;
            Real epi_reward = (Real)0.0;
            u64 o0 = env.reset();
            size_t curr_action = get_action(o0);
            rec(env, 0, o0, curr_action, false, 0.0, 0.0);
            for (;;) {
                action_counter += 1;
                StepResult sr;
                if (!env.step(curr_action, sr)) return false;
                size_t next_action = get_action(sr.obs);
                rec(env, 2, sr.obs, next_action, sr.terminated, sr.reward, 0.0);
                curr_action = next_action;
                epi_reward += (Real)sr.reward;
                if (sr.terminated) { reward_history.push_back(epi_reward); break; }
            }
            eval_steps += action_counter;
            episode_length.push_back(action_counter);
        }
        return true;
    }
};

// agent/one_step_agent.rs:7-86
template <int A, class Real>
struct OneStepAgent : Agent<A, Real> {
    using Row = std::array<Real, A>;
    std::unique_ptr<Policy<A, Real>> policy;
    Real discount_factor;
    std::unique_ptr<ActionSelection<A, Real>> action_selection;
    GetNextQValue<A, Real> get_next_q_value;
    OneStepAgent(std::unique_ptr<Policy<A, Real>> p, Real gamma, std::unique_ptr<ActionSelection<A, Real>> s,
                 GetNextQValue<A, Real> f)
        : policy(std::move(p)), discount_factor(gamma), action_selection(std::move(s)), get_next_q_value(f) {}
    void set_future_q_value_func(GetNextQValue<A, Real> f) override { get_next_q_value = f; }                     // :35-37
    void set_action_selector(std::unique_ptr<ActionSelection<A, Real>> s) override { action_selection = std::move(s); }   // :39-41
    void reset() override { action_selection->reset(); policy->reset(); }                                         // :43-46
    size_t get_action(u64 obs) override { return action_selection->get_action(obs, policy->predict(obs)); }       // :48-51
    Real update(u64 curr_obs, size_t curr_action, Real reward, bool terminated, u64 next_obs, size_t next_action) override {   // :53-86
        Row next_q_values = policy->get_values(next_obs);
        Real future_q_value = get_next_q_value(next_q_values, next_action,
                                               action_selection->get_exploration_probs(next_obs, next_q_values));
        Row curr_q_values = policy->get_values(curr_obs);
        Real temporal_difference = reward + discount_factor * future_q_value - curr_q_values[curr_action];
        policy->update(curr_obs, curr_action, next_obs, temporal_difference);
        policy->after_update();
        if (terminated) action_selection->update();
        return temporal_difference;
    }
};

// agent/elegibility_traces_agent.rs:8-104
template <int A, class Real>
struct ElegibilityTracesAgent : Agent<A, Real> {
    using Row = std::array<Real, A>;
    std::unique_ptr<Policy<A, Real>> policy;
    Real discount_factor;
    std::unique_ptr<ActionSelection<A, Real>> action_selection;
    Real lambda_factor;
    FxMap<Row> trace;
    GetNextQValue<A, Real> get_next_q_value;
    ElegibilityTracesAgent(std::unique_ptr<Policy<A, Real>> p, Real gamma, std::unique_ptr<ActionSelection<A, Real>> s,
                           Real lambda, GetNextQValue<A, Real> f)
        : policy(std::move(p)), discount_factor(gamma), action_selection(std::move(s)), lambda_factor(lambda),
          get_next_q_value(f) {}
    void set_future_q_value_func(GetNextQValue<A, Real> f) override { get_next_q_value = f; }                     // :43-45
    void set_action_selector(std::unique_ptr<ActionSelection<A, Real>> s) override { action_selection = std::move(s); }   // :47-49
    void reset() override { action_selection->reset(); policy->reset(); }                                         // :51-54 (trace kept)
    size_t get_action(u64 obs) override { return action_selection->get_action(obs, policy->predict(obs)); }       // :56-59
    Real update(u64 curr_obs, size_t curr_action, Real reward, bool terminated, u64 next_obs, size_t next_action) override {   // :61-104
        Row next_q_values = policy->get_values(next_obs);
        Real future_q_value = get_next_q_value(next_q_values, next_action,
                                               action_selection->get_exploration_probs(next_obs, next_q_values));
        Row curr_q_values = policy->get_values(curr_obs);
        Real temporal_difference = reward + discount_factor * future_q_value - curr_q_values[curr_action];
        Row zero; zero.fill((Real)0.0);
        Row& curr_trace = trace.entry_or_insert(curr_obs, zero);
        curr_trace[curr_action] += (Real)1.0;
        Policy<A, Real>* pol = policy.get();
        const Real decay = discount_factor * lambda_factor;   // evaluated per cell in the reference; same value
        trace.for_each([&](u64 obs, Row& trace_values) {
            for (int action = 0; action < A; ++action) {
                pol->update(obs, (size_t)action, next_obs, temporal_difference * trace_values[action]);
                trace_values[action] *= decay;
            }
        });
        policy->after_update();
        if (terminated) { trace.clear(); action_selection->update(); }
        return temporal_difference;
    }
};

// model/random_model.rs:9-45 — `IndexMap<(T, usize), (T, f64)>`: first-seen transitions in insertion order
// (nothing is ever removed, so `get_index(i)` is the i-th inserted entry).
template <int A, class Real>
struct RandomModel {
    struct Info { u64 obs; size_t action; u64 next_obs; Real reward; };
    std::vector<Info> values;
    std::map<std::pair<u64, size_t>, size_t> slot;   // key -> index into `values`
    Stream* rng;
    explicit RandomModel(Stream* rng_) : rng(rng_) {}
    Info get_info() { return values[(size_t)gen_range_usize(*rng, (u64)values.size())]; }                       // :27-35 (panics when empty)
    void add_info(u64 obs, size_t action, Real reward, u64 next_obs) {                                            // :37-41 `.entry().or_insert()`
        auto key = std::make_pair(obs, action);
        if (slot.find(key) != slot.end()) return;
        slot.emplace(key, values.size());
        values.push_back({obs, action, next_obs, reward});
    }
    void reset() { values.clear(); slot.clear(); }                                                                // :43-45
};

// agent/internal_model_agent.rs:9-85 — Dyna: wraps a borrowed agent, learns a model of the transitions it sees and
// replays `planning_steps` sampled ones after every real update.
template <int A, class Real>
struct InternalModelAgent : Agent<A, Real> {
    Agent<A, Real>* agent;   // `Box<RefCell<&'a mut dyn Agent>>`
    RandomModel<A, Real> model;
    size_t planning_steps;
    InternalModelAgent(Agent<A, Real>* inner, RandomModel<A, Real> m, size_t planning_length)
        : agent(inner), model(std::move(m)), planning_steps(planning_length) {}
    void set_future_q_value_func(GetNextQValue<A, Real> f) override { agent->set_future_q_value_func(f); }                       // :34-36
    void set_action_selector(std::unique_ptr<ActionSelection<A, Real>> s) override { agent->set_action_selector(std::move(s)); }   // :38-40
    size_t get_action(u64 obs) override { return agent->get_action(obs); }                                                       // :42-44
    Real update(u64 curr_obs, size_t curr_action, Real reward, bool terminated, u64 next_obs, size_t next_action) override {     // :46-79
        Real td = agent->update(curr_obs, curr_action, reward, terminated, next_obs, next_action);
        model.add_info(curr_obs, curr_action, reward, next_obs);
        for (size_t i = 0; i < planning_steps; ++i) {
            auto info = model.get_info();
            size_t planned_action = agent->get_action(info.next_obs);
            agent->update(info.obs, info.action, info.reward, false, info.next_obs, planned_action);
        }
        return td;
    }
    void reset() override { agent->reset(); model.reset(); }                                                                     // :81-84
};

}   // namespace oracle
