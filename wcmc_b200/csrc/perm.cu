// Pseudo-random permutations on the device without a sort.
//
// The path-disentangling loss pairs every row i with row pi(i) of a fresh random permutation, twice per call
// (/root/reference/support/losses.py:35 `torch.randperm(s*h*w)`, :50 `torch.randperm(b*s*h*w)`).  The reference draws
// them on the CPU; torch.randperm on the GPU sorts random keys (a radix sort: 3.4 % of the round-1 step for 4 x 541,696
// indices).  In the throughput mode (`FeatureMSE(rng="device")`) the pairing only has to be a uniform-looking random
// bijection, so here pi is a keyed FEISTEL network over the next power-of-four domain >= n with cycle walking
// (re-encrypt until the value falls below n: a bijection of [0, n) for any n): every thread computes its own index,
// no communication, one 8-byte store.  The key comes from a device-resident counter that the launch itself advances
// (the last CTA to finish bumps it), so a CUDA graph that replays the launch draws a new permutation every time.
// The parity mode (`rng="cpu"`) never comes here: it consumes the reference's CPU generator stream.
#include <algorithm>

#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {   // murmur3 finaliser
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}

constexpr int kPermThreads = 256;
constexpr int kRounds = 6;

__global__ void __launch_bounds__(kPermThreads)
feistel_perm_kernel(int64_t* __restrict__ out, long n, int half_bits, unsigned long long* __restrict__ state, uint32_t salt) {
    wcmc::pdl_start();
    __shared__ uint32_t keys[kRounds];
    if (threadIdx.x < kRounds) {
        const unsigned long long c = state[0];
        keys[threadIdx.x] = mix32(static_cast<uint32_t>(c) ^ mix32(static_cast<uint32_t>(c >> 32) + 0x9E3779B9u * (threadIdx.x + 1)) ^ salt);
    }
    __syncthreads();
    const uint32_t mask = (1u << half_bits) - 1u;
    for (long i = blockIdx.x * static_cast<long>(kPermThreads) + threadIdx.x; i < n;
         i += static_cast<long>(gridDim.x) * kPermThreads) {
        unsigned long long x = static_cast<unsigned long long>(i);
        do {
            uint32_t l = static_cast<uint32_t>(x >> half_bits) & mask, r = static_cast<uint32_t>(x) & mask;
#pragma unroll
            for (int k = 0; k < kRounds; ++k) {
                const uint32_t f = mix32(r ^ keys[k]) & mask;
                const uint32_t t = l ^ f;
                l = r;
                r = t;
            }
            x = (static_cast<unsigned long long>(l) << half_bits) | r;
        } while (x >= static_cast<unsigned long long>(n));
        out[i] = static_cast<int64_t>(x);
    }
    // the last CTA to finish advances the counter (every CTA has read it by then) and re-arms the ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long done = atomicAdd(&state[1], 1ull);
        if (done == gridDim.x - 1) {
            state[1] = 0ull;
            state[0] = state[0] + 1ull;
            __threadfence();
        }
    }
}

}  // namespace

extern "C" int wcmc_random_permutation(int64_t* out, long n, unsigned long long* state, unsigned salt, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(out != nullptr && state != nullptr && n > 0 && n < (1L << 40), WCMC_ESHAPE,
                 "random_permutation: bad arguments (n = %ld)", n);
    int bits = 2;
    while ((1L << bits) < n) bits += 2;          // even number of bits: two equal Feistel halves; domain < 4 n
    const int grid = static_cast<int>(std::min<long>((n + kPermThreads - 1) / kPermThreads, 4L * wcmc_num_sms()));
    WCMC_LAUNCH(feistel_perm_kernel, grid, kPermThreads, 0, stream, out, n, bits / 2, state, salt);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}
