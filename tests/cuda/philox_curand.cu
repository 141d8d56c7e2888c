// tests/cuda/philox_curand.cu — TEST INFRASTRUCTURE: an independent pin of the engine's RNG contract.
//
// Prints words of the per-agent stream  w[n] = Philox4x32-10(key = seed, ctr = (n >> 2, agent))[n & 3]  computed by
// NVIDIA's own implementation in two ways:
//   raw : curand_Philox4x32_10(ctr, key) from <curand_philox4x32_x.h>
//   api : curand_init(seed, subsequence = agent, offset = first_word, &state) + curand() of a curandStatePhilox4_32_10_t
// tests/test_gpu_curand.py compares both with librlb's host stream (rlb_rng_words), which the parity tests in turn
// tie to the device kernels and to the oracle.  Usage: philox_curand SEED AGENT FIRST_WORD COUNT   (FIRST_WORD % 4 == 0)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <curand_kernel.h>

__global__ void k_words(unsigned long long seed, unsigned long long agent, unsigned long long first, unsigned count, uint32_t* raw, uint32_t* api) {
    if (threadIdx.x || blockIdx.x) return;
    for (unsigned i = 0; i < count; i += 4) {
        const unsigned long long blk = (first + i) >> 2;
        uint4 c = make_uint4((unsigned)blk, (unsigned)(blk >> 32), (unsigned)agent, (unsigned)(agent >> 32));
        uint2 k = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
        uint4 r = curand_Philox4x32_10(c, k);
        raw[i] = r.x; if (i + 1 < count) raw[i + 1] = r.y; if (i + 2 < count) raw[i + 2] = r.z; if (i + 3 < count) raw[i + 3] = r.w;
    }
    curandStatePhilox4_32_10_t st;
    curand_init(seed, agent, first, &st);
    for (unsigned i = 0; i < count; ++i) api[i] = curand(&st);
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s SEED AGENT FIRST_WORD COUNT\n", argv[0]); return 2; }
    const unsigned long long seed = strtoull(argv[1], nullptr, 0), agent = strtoull(argv[2], nullptr, 0), first = strtoull(argv[3], nullptr, 0);
    const unsigned count = (unsigned)strtoul(argv[4], nullptr, 0);
    uint32_t *d_raw, *d_api;
    if (cudaMalloc(&d_raw, count * 4) != cudaSuccess || cudaMalloc(&d_api, count * 4) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 3; }
    k_words<<<1, 1>>>(seed, agent, first, count, d_raw, d_api);
    std::vector<uint32_t> raw(count), api(count);
    if (cudaMemcpy(raw.data(), d_raw, count * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 4; }
    cudaMemcpy(api.data(), d_api, count * 4, cudaMemcpyDeviceToHost);
    for (unsigned i = 0; i < count; ++i) printf("%08x %08x\n", raw[i], api[i]);
    return 0;
}
