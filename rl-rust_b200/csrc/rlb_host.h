// rlb_host.h — host-only helpers of librlb (see rlb_host.cpp).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rlb.h"

namespace rlb {

struct EnvTables {
    uint32_t S = 0, A = 0;
    std::vector<uint16_t> trans;       // s' | rcode << 10 | terminated << 15
    std::vector<uint64_t> thr;         // start thresholds (k-space): Taxi's 300 valid starts; a FrozenLake map's 'S' cells when it has several
    std::vector<uint16_t> thr_state;
    uint32_t fl_start = 0;             // FrozenLake with exactly one 'S': that cell
    uint32_t thr_direct = 0;           // 1: the start index is floor(k * n_thr / 2^52) or its neighbour (verified when built)
    uint64_t slip_thr0 = 0, slip_thr1 = 0;
    uint32_t n_live = 0;               // states an action is ever taken from
    uint8_t row_lut[64];               // state -> compact live-row index, 0xFF for terminal states (S <= 64 only)
    std::vector<uint8_t> dead_cell;    // FrozenLake: hole / goal cells
};

// the calling thread's last failure (rlb_last_error_string)
void set_error(const char* fmt, ...);
const char* last_error_cstr();

bool start_index_is_direct(const std::vector<uint64_t>& thr);   // licence for start_index_direct(), rlb_taxi_start.h
bool build_env_tables(const rlb_config& cfg, EnvTables& out, std::string& err);
void host_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
uint32_t host_word(uint64_t seed, uint64_t agent, uint64_t n);
uint64_t blackjack_id(uint32_t p, uint32_t d, uint32_t ace);

}   // namespace rlb
