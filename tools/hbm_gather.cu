// tools/hbm_gather.cu — what can HBM deliver for the access pattern of the one-step kernels?
//
// k_run's HBM store reads ONE random 32-byte sector per agent-step (the Q row of the next state: every agent owns a
// private table, so consecutive steps of one agent land in unrelated sectors) and dirties 4 bytes of a sector it read
// one step earlier.  This microbenchmark measures the chip's ceiling for exactly that: every thread walks its own
// `stride`-byte region (one "agent table"), each iteration loading one 32-byte row at a pseudo-random row index that
// DEPENDS on the previously loaded data (as the next state depends on the chosen action) and optionally storing 4 bytes
// back into the previous row.  Reported: sectors/s and GB/s of useful 32-byte rows, for several occupancies.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hbm_gather tools/hbm_gather.cu
//   tools/hbm_gather [rows_per_table=500] [tables=2097152] [loads_per_thread=2000] [float4s_per_row=2|4|8]
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e__ = (x);                                                                 \
        if (e__ != cudaSuccess) { std::printf("%s: %s\n", #x, cudaGetErrorString(e__)); std::exit(1); } \
    } while (0)

// ROWV = float4s per row: 2 = one 32-byte sector (a Taxi f32 row), 4 = 64 B, 8 = a whole 128-byte line.  Only the first
// 32 bytes of a row are consumed; wider rows show what else a DRAM access brings along.
template <bool WRITE, int ROWV>
__global__ void __launch_bounds__(128) k_walk(float* base, uint64_t n_threads, uint32_t rows_per_table, uint32_t iters, float* sink) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_threads) return;
    float4* table = reinterpret_cast<float4*>(base) + i * (uint64_t)rows_per_table * ROWV;
    uint32_t x = (uint32_t)i * 2654435761u + 12345u;
    uint32_t prev = x % rows_per_table;
    float acc = 0.f;
    for (uint32_t k = 0; k < iters; ++k) {
        x = x * 1664525u + 1013904223u;
        const uint32_t row = (x >> 8) % rows_per_table;
        float m = -1e30f;
#pragma unroll
        for (int v = 0; v < ROWV; ++v) {
            const float4 a = table[row * ROWV + v];
            m = fmaxf(m, fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)));
        }
        acc += m;
        if (WRITE) reinterpret_cast<float*>(table + prev * ROWV)[x & 3u] = m * 0.95f;
        x ^= __float_as_uint(m) & 1u;   // the next row depends on what was loaded
        prev = row;
    }
    if (acc == 123.456f) sink[0] = acc;
}

int main(int argc, char** argv) {
    const uint32_t rows = argc > 1 ? (uint32_t)std::atoi(argv[1]) : 500;         // Taxi: 500 states
    const uint64_t agents = argc > 2 ? (uint64_t)std::atoll(argv[2]) : (1ull << 21);
    const uint32_t iters = argc > 3 ? (uint32_t)std::atoi(argv[3]) : 2000;
    float *buf, *sink;
    const int rowv = argc > 4 ? std::atoi(argv[4]) : 2;
    const size_t bytes = (size_t)agents * rows * 16 * rowv;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(buf, 0, bytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    std::printf("tables: %llu x %u rows x %d B = %.1f GB; %u dependent row loads per thread\n", (unsigned long long)agents, rows, 16 * rowv, bytes / 1e9, iters);
    for (int write = 0; write < 2; ++write) {
        for (int ctas_per_sm : {2, 8}) {
            // occupancy is capped by dynamic shared memory: 228 KB / ctas
            const size_t smem = (size_t)(227 * 1024 / ctas_per_sm) - 1024;
            auto kern = write ? (rowv == 2 ? k_walk<true, 2> : rowv == 4 ? k_walk<true, 4> : k_walk<true, 8>)
                              : (rowv == 2 ? k_walk<false, 2> : rowv == 4 ? k_walk<false, 4> : k_walk<false, 8>);
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const unsigned grid = (unsigned)((agents + 127) / 128);
            kern<<<grid, 128, smem>>>(buf, agents, rows, 50, sink);   // warm-up
            CK(cudaEventRecord(e0));
            kern<<<grid, 128, smem>>>(buf, agents, rows, iters, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double loads = (double)agents * iters;
            std::printf("%s  %2d CTAs/SM (%4d thr/SM): %7.1f ms  %.3e rows/s  %.0f GB/s of rows\n", write ? "load+store" : "load only ", ctas_per_sm,
                        ctas_per_sm * 128, ms, loads / (ms * 1e-3), loads * 16 * rowv / (ms * 1e-3) / 1e9);
        }
    }
    return 0;
}
