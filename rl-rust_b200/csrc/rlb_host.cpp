// rlb_host.cpp — host-only parts of librlb: environment transition tables (built once per
// engine from the reference constructors' rules and uploaded to the device), the
// host-callable RNG contract, and the Blackjack observation-id bijection.
#include "rlb_host.h"
#include "rlb_taxi_start.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace rlb {

static thread_local std::string g_last_error;
void set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}
const char* last_error_cstr() { return g_last_error.c_str(); }

static inline uint16_t pack_tr(uint32_t s2, uint32_t rcode, bool term) {
    return (uint16_t)(s2 | (rcode << 10) | (term ? 0x8000u : 0u));
}

// u = k * 2^-52 with integer k in [0, 2^52);  u < b  <=>  k < ceil(b * 2^52)   (b * 2^52 is exact)
static inline uint64_t k_threshold(double b) {
    if (!(b > 0.0)) return 0;
    if (b >= 1.0) return 1ull << 52;
    return (uint64_t)std::ceil(std::ldexp(b, 52));
}

// Grid move shared by FrozenLake and CliffWalking: 0 left, 1 down, 2 right, 3 up, clamped
// at the border (reference utils.rs:53-76).
static inline void grid_move(int nrow, int ncol, int& row, int& col, int a) {
    switch (a) {
        case 0: if (col > 0) col -= 1; break;
        case 1: if (row < nrow - 1) row += 1; break;
        case 2: if (col < ncol - 1) col += 1; break;
        case 3: if (row > 0) row -= 1; break;
        default: break;
    }
}

// env/taxi.rs:57-131.  State = ((row*5+col)*5+pass)*4+dest (:33-42).  Reward codes:
// 0 -> -1, 1 -> -10, 2 -> +20.
static void build_taxi(EnvTables& t) {
    static const char* kMap[5] = {"|R: | : :G|", "| : | : : |", "| : : : : |", "| | : | : |", "|Y| : |B: |"};   // :21-29 rows 1..5
    static const int kLoc[4][2] = {{0, 0}, {0, 4}, {4, 0}, {4, 3}};                                             // :30
    t.S = 500; t.A = 6;
    t.trans.assign(3000, 0);
    for (int state = 0; state < 500; ++state) {
        const int dest = state % 4, pass = (state / 4) % 5, col = (state / 20) % 5, row = state / 100;
        for (int a = 0; a < 6; ++a) {
            int nr = row, nc = col, np = pass;
            uint32_t rcode = 0;
            bool term = false;
            if (a == 0) nr = row < 4 ? row + 1 : 4;                                  // south
            else if (a == 1) nr = row > 0 ? row - 1 : 0;                             // north
            else if (a == 2) { if (kMap[row][2 * col + 2] == ':') nc = col < 4 ? col + 1 : 4; }   // east unless a wall
            else if (a == 3) { if (kMap[row][2 * col] == ':') nc = col > 0 ? col - 1 : 0; }       // west unless a wall
            else if (a == 4) {                                                       // pickup
                if (pass < 4 && row == kLoc[pass][0] && col == kLoc[pass][1]) np = 4; else rcode = 1;
            } else {                                                                 // dropoff
                if (pass == 4 && row == kLoc[dest][0] && col == kLoc[dest][1]) { np = dest; term = true; rcode = 2; }
                else rcode = 1;
            }
            const int ns = ((nr * 5 + nc) * 5 + np) * 4 + dest;
            t.trans[state * 6 + a] = pack_tr((uint32_t)ns, rcode, term);
        }
    }
    // start distribution (:66-69,119-121): passenger not in the taxi and not at the destination,
    // each 1.0/300.0; reset() walks the running f64 sum over all 500 entries (:138, utils.rs:33-43).
    int valid = 0;
    for (int state = 0; state < 500; ++state) { int dest = state % 4, pass = (state / 4) % 5; if (pass < 4 && pass != dest) ++valid; }
    const double each = 1.0 / (double)valid;
    double running = 0.0;
    t.thr.clear(); t.thr_state.clear();
    for (int state = 0; state < 500; ++state) {
        int dest = state % 4, pass = (state / 4) % 5;
        if (pass < 4 && pass != dest) {
            running += each;
            t.thr.push_back(k_threshold(running));
            t.thr_state.push_back((uint16_t)state);
        }
    }
    t.thr_direct = start_index_is_direct(t.thr) ? 1u : 0u;
}

// Licence for start_index_direct() (rlb_taxi_start.h): compare it with the search at every breakpoint of either step
// function — each threshold and each k where floor(k * n / 2^52) steps, +-2 — and at both ends.  Between two
// consecutive breakpoints both functions are constant, so agreement there is agreement everywhere.
bool start_index_is_direct(const std::vector<uint64_t>& thr) {
    const uint64_t n = thr.size(), top = 1ull << 52;
    if (n == 0 || n > 0xffffu) return false;
    for (uint64_t i = 0; i + 1 < n; ++i) if (!(thr[i] < thr[i + 1])) return false;
    if (thr[n - 1] > top) return false;
    auto same = [&](uint64_t k) {
        if (k >= top) return true;   // also catches the wrap of `x - 2` below 0
        return start_index_direct(thr.data(), (uint32_t)n, k) == start_index_search(thr.data(), (uint32_t)n, k);
    };
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t b = (uint64_t)((((unsigned __int128)(i + 1)) << 52) / n);   // around where the guess steps to i + 1
        for (uint64_t d = 0; d < 5; ++d)
            if (!same(thr[i] + d - 2) || !same(b + d - 2)) return false;
    }
    return same(0) && same(1) && same(top - 1) && same(top - 2);
}

// env/cliff_walking.rs:22-63.  Reward codes: 0 -> -1, 1 -> -100.
static void build_cliff(EnvTables& t) {
    t.S = 48; t.A = 4;
    t.trans.assign(48 * 4, 0);
    for (int s = 0; s < 48; ++s) for (int a = 0; a < 4; ++a) {
        int row = s / 12, col = s % 12;
        grid_move(4, 12, row, col, a);
        const int ns = row * 12 + col;
        const bool lose = ns >= 37 && ns <= 46, win = ns == 47;
        t.trans[s * 4 + a] = pack_tr((uint32_t)ns, lose ? 1u : 0u, lose || win);
    }
}

// env/frozen_lake.rs:23-102.  [S][4][3] slots; reward code 1 -> 1.0 (entering G).  `map` is MAP_4X4 (:23), MAP_8X8
// (:25-28) or the caller's own rows (FrozenLakeEnv::new takes any `&[&str]`, :48).
static bool build_frozen_lake(const rlb_config& cfg, EnvTables& t, std::string& err) {
    static const char* k4 = "SFFF" "FHFH" "FFFH" "HFFG";
    static const char* k8 = "SFFFFFFF" "FFFFFFFF" "FFFHFFFF" "FFFFFHFF" "FFFHFFFF" "FHHFFFHF" "FHFFHFHF" "FFFHFFFG";
    int nrow, ncol;
    std::string map;
    if (cfg.map_id == RLB_MAP_4X4) { nrow = ncol = 4; map = k4; }
    else if (cfg.map_id == RLB_MAP_8X8) { nrow = ncol = 8; map = k8; }
    else if (cfg.map_id == RLB_MAP_CUSTOM) {
        if (!cfg.map || cfg.map_rows == 0 || cfg.map_cols == 0) { err = "frozen lake: map_id = RLB_MAP_CUSTOM needs map, map_rows, map_cols"; return false; }
        if ((uint64_t)cfg.map_rows * cfg.map_cols > 1024) { err = "frozen lake: at most 1024 cells"; return false; }
        nrow = (int)cfg.map_rows; ncol = (int)cfg.map_cols;
        if (strnlen(cfg.map, (size_t)nrow * ncol + 1) != (size_t)nrow * ncol) { err = "frozen lake: map must hold exactly map_rows * map_cols cells"; return false; }
        map.assign(cfg.map, (size_t)nrow * ncol);
        for (char c : map) if (c != 'S' && c != 'F' && c != 'H' && c != 'G') { err = "frozen lake: map cells must be S, F, H or G"; return false; }
    } else { err = "frozen lake: map_id must be 0 (4x4), 1 (8x8) or 2 (custom)"; return false; }
    const int n_cells = nrow * ncol;
    t.S = (uint32_t)n_cells; t.A = 4;
    t.trans.assign((size_t)t.S * 12, 0);
    t.dead_cell.assign(t.S, 0);
    for (int s = 0; s < n_cells; ++s) {
        const int row = s / ncol, col = s % ncol;
        const char here = map[s];
        for (int a = 0; a < 4; ++a) {
            uint16_t* slot = &t.trans[((size_t)s * 4 + a) * 3];
            if (here == 'G' || here == 'H') { slot[0] = pack_tr((uint32_t)s, 0, true); t.dead_cell[s] = 1; continue; }   // :75-76, never stepped from
            const int cand[3] = {(a + 3) % 4, a, (a + 1) % 4};   // :78 (usize wrap: (0-1)%4 == 3)
            const int n_slots = cfg.slippery ? 3 : 1;
            for (int i = 0; i < n_slots; ++i) {
                int r2 = row, c2 = col;
                grid_move(nrow, ncol, r2, c2, cfg.slippery ? cand[i] : a);
                const char c = map[r2 * ncol + c2];
                slot[i] = pack_tr((uint32_t)(r2 * ncol + c2), c == 'G' ? 1u : 0u, c == 'G' || c == 'H');
            }
        }
    }
    // categorical_sample over [1/3,1/3,1/3] with a running f64 sum (:81,125-127)
    const double third = 1.0 / 3.0;
    const double b1 = third, b2 = b1 + third, b3 = b2 + third;
    if (b3 < 1.0) { err = "frozen lake: slip distribution does not reach 1.0"; return false; }
    t.slip_thr0 = k_threshold(b1);
    t.slip_thr1 = k_threshold(b2);
    // start distribution (:54-66): 1/count on every 'S' cell; reset() walks the running f64 sum over all cells
    // (:106-109, utils.rs:33-43) — the first 'S' whose cumulative exceeds the draw, cell 0 if none does (also the
    // answer for a map without 'S': the distribution is all zeros).
    int count = 0;
    for (char c : map) count += c == 'S';
    t.thr.clear(); t.thr_state.clear();
    const double each = count ? 1.0 / (double)count : 0.0;
    double running = 0.0;
    for (int s = 0; s < n_cells; ++s) {
        if (map[s] != 'S') continue;
        running += each;
        t.thr.push_back(k_threshold(running));
        t.thr_state.push_back((uint16_t)s);
    }
    t.fl_start = 0;
    if (count == 1) {   // one start cell: cumulative 1.0 exceeds every draw
        t.fl_start = t.thr_state[0];
        t.thr.clear(); t.thr_state.clear();
    }
    return true;
}

// Which states can be a step's `curr_obs`: everything but the terminal cells (FrozenLake holes / goal, Cliff cells /
// goal).  Terminal states are only ever OBSERVED (Agent::get_action on the last observation, agent.rs:89), never
// updated, so a table store may keep just the live rows on chip.
static void build_row_lut(const rlb_config& cfg, EnvTables& t) {
    std::memset(t.row_lut, 0xFF, sizeof t.row_lut);
    t.n_live = t.S;
    if (t.S > 64) return;
    uint32_t next = 0;
    for (uint32_t s = 0; s < t.S; ++s) {
        bool dead = false;
        if (cfg.env_kind == RLB_ENV_FROZEN_LAKE) dead = t.dead_cell[s] != 0;
        if (cfg.env_kind == RLB_ENV_CLIFF_WALKING) dead = s >= 37;
        if (!dead) t.row_lut[s] = (uint8_t)next++;
    }
    t.n_live = next;
}

bool build_env_tables(const rlb_config& cfg, EnvTables& t, std::string& err) {
    t = EnvTables();
    switch (cfg.env_kind) {
        case RLB_ENV_BLACKJACK: t.S = 1456; t.A = 2; t.n_live = t.S; std::memset(t.row_lut, 0xFF, sizeof t.row_lut); return true;
        case RLB_ENV_FROZEN_LAKE:
            if (!build_frozen_lake(cfg, t, err)) return false;
            build_row_lut(cfg, t);
            return true;
        case RLB_ENV_CLIFF_WALKING: build_cliff(t); build_row_lut(cfg, t); return true;
        case RLB_ENV_TAXI: build_taxi(t); t.n_live = t.S; std::memset(t.row_lut, 0xFF, sizeof t.row_lut); return true;
    }
    err = "unknown env_kind";
    return false;
}

// ---------------------------------------------------------------- host RNG contract
void host_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        const uint64_t m0 = 0xD2511F53ull * c[0], m1 = 0xCD9E8D57ull * c[2];
        const uint32_t t0 = (uint32_t)(m1 >> 32) ^ c[1] ^ k[0], t2 = (uint32_t)(m0 >> 32) ^ c[3] ^ k[1];
        c[1] = (uint32_t)m1; c[3] = (uint32_t)m0; c[0] = t0; c[2] = t2;
        k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
    std::memcpy(out, c, 16);
}
uint32_t host_word(uint64_t seed, uint64_t agent, uint64_t n) {
    const uint64_t blk = n >> 2;
    const uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)agent, (uint32_t)(agent >> 32)};
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    host_philox(ctr, key, out);
    return out[n & 3];
}

// ---------------------------------------------------------------- Blackjack ids
// fxhash 0.2.1 64-bit: h = (rotl(h,5) ^ byte) * 0x517cc1b727220a95 for p_score, d_score, p_ace.
uint64_t blackjack_id(uint32_t p, uint32_t d, uint32_t ace) {
    uint64_t h = 0;
    const uint64_t in[3] = {p & 0xffu, d & 0xffu, ace ? 1u : 0u};
    for (uint64_t b : in) h = (((h << 5) | (h >> 59)) ^ b) * 0x517cc1b727220a95ull;
    return h;
}

}   // namespace rlb
