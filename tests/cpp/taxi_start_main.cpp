// CPU check of the Taxi start-state draw (reference src/env/taxi.rs:135-142, utils.rs:33-43): the direct form the
// device uses (rlb_taxi_start.h) against a plain left-to-right scan of the thresholds, at every breakpoint, in a band
// around each, and on a few million pseudo-random draws; and that the host licenses the table it built.
#include <cstdio>
#include <cstdlib>
#include "../../rl-rust_b200/csrc/rlb_host.h"
#include "../../rl-rust_b200/csrc/rlb_taxi_start.h"

static uint32_t scan(const std::vector<uint64_t>& thr, uint64_t k) {   // categorical_sample: first running sum that exceeds the draw
    uint32_t i = 0;
    while (i < thr.size() && !(thr[i] > k)) ++i;
    return i;
}

int main() {
    rlb_config cfg{};
    cfg.env_kind = RLB_ENV_TAXI;
    rlb::EnvTables t;
    std::string err;
    if (!rlb::build_env_tables(cfg, t, err)) { std::printf("FAILED build: %s\n", err.c_str()); return 1; }
    const uint32_t n = (uint32_t)t.thr.size();
    const uint64_t top = 1ull << 52;
    std::printf("n_thr %u direct %u last %llu\n", n, t.thr_direct, (unsigned long long)t.thr[n - 1]);
    if (n != 300 || t.thr_direct != 1 || !rlb::start_index_is_direct(t.thr)) { std::printf("FAILED licence\n"); return 1; }
    unsigned long long checked = 0, fall_through = 0;
    auto check = [&](uint64_t k) {
        if (k >= top) return true;
        const uint32_t want = scan(t.thr, k);
        ++checked;
        if (want == n) ++fall_through;
        return rlb::start_index_direct(t.thr.data(), n, k) == want && rlb::start_index_search(t.thr.data(), n, k) == want;
    };
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t b = (uint64_t)((((unsigned __int128)(i + 1)) << 52) / n);
        for (uint64_t d = 0; d < 4097; ++d)
            if (!check(t.thr[i] + d - 2048) || !check(b + d - 2048)) { std::printf("FAILED at i=%u d=%llu\n", i, (unsigned long long)d); return 1; }
    }
    for (uint64_t k = 0; k < 4096; ++k) if (!check(k) || !check(top - 1 - k)) { std::printf("FAILED at the ends\n"); return 1; }
    uint64_t x = 0x9E3779B97F4A7C15ull;
    for (int it = 0; it < 4000000; ++it) {   // splitmix64
        x += 0x9E3779B97F4A7C15ull;
        uint64_t z = x; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        if (!check(z >> 12)) { std::printf("FAILED random k=%llu\n", (unsigned long long)(z >> 12)); return 1; }
    }
    // 17 of the 2^52 uniforms fall through to state 0
    unsigned long long tail = 0;
    for (uint64_t k = top - 64; k < top; ++k) if (scan(t.thr, k) == n) ++tail;
    // a table that is NOT equal-weight must be refused
    std::vector<uint64_t> skew = t.thr;
    for (uint32_t i = 0; i < n / 2; ++i) skew[i] = t.thr[i] / 4;
    std::printf("checked %llu tail %llu skew_direct %d\n", checked, tail, (int)rlb::start_index_is_direct(skew));
    if (tail != 17 || rlb::start_index_is_direct(skew)) { std::printf("FAILED tail/skew\n"); return 1; }
    std::printf("OK\n");
    return 0;
}
