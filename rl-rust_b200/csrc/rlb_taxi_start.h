// rlb_taxi_start.h — Taxi's start-state draw (reference src/env/taxi.rs:135-142 through utils.rs:33-43
// `categorical_sample`): the index of the first cumulative threshold that exceeds the uniform draw, in integer
// "k-space" (u = k * 2^-52, thr[i] = ceil(cumulative_i * 2^52)).  Shared by the device code (rlb_device.cuh), the
// host-side check that licenses the direct form (rlb_host.cpp) and its CPU test (tests/cpp/taxi_start_main.cpp).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define RLB_HD __host__ __device__ __forceinline__
#else
#define RLB_HD inline
#endif

namespace rlb {

RLB_HD uint64_t mulhi_u64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// first index with thr[idx] > k, or n if none: binary search (any increasing thresholds)
RLB_HD uint32_t start_index_search(const uint64_t* thr, uint32_t n, uint64_t k) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (thr[mid] > k) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// The thresholds are the running sum of n EQUAL weights, i.e. within a few units of (i+1) * 2^52 / n, so the index can
// only be g = floor(k * n / 2^52) or a neighbour of it: one multiply and two compares.  Valid only for tables that
// passed start_index_is_direct() (rlb_host.cpp), which checks every breakpoint of both step functions.  k < 2^52.
RLB_HD uint32_t start_index_direct(const uint64_t* thr, uint32_t n, uint64_t k) {
    const uint32_t g = (uint32_t)mulhi_u64(k << 12, (uint64_t)n);   // <= n - 1
    const uint64_t below = thr[g > 0u ? g - 1u : 0u], at = thr[g];
    return (g > 0u && below > k) ? g - 1u : (at > k ? g : g + 1u);
}

}   // namespace rlb
