// tcgen05.mma issue-rate probe (measurement tool, not part of libwcmc.so).
//
// Question it answers: what does ONE SM retire per clock for the operand shapes the
// conv kernels use -- M128 x N112 x K16 with both operands in shared memory, the A descriptor a shifted
// window of a halo (stride-byte-offset = halo row pitch, start at any 128-byte row) -- and what would the
// alternatives give (canonical SBO = 1024, N = 128 / 224 / 256, cta_group::2 with M = 256)?
//
// Every variant: all SMs busy (one CTA per SM, 200 KB of dynamic shared memory), shared memory filled with
// small random fp16 values (realistic switching power), one elected thread issues `rounds` groups of
// `group` MMAs, each group followed by a tcgen05.commit on a ring of mbarriers; the issuer stays at most
// `depth` groups ahead.  Reported per variant: clocks per MMA on SM 0 (clock64), wall time per MMA
// (globaltimer) -> effective SM clock, and the aggregate TFLOP/s over all CTAs.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/mma_probe tools/mma_probe.cu
// Run:    build/mma_probe            (prints one JSON object per line)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded: a mistake in the probe must trap, never hang the box
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 2000000000LL) return false;
    }
    return true;
}
__device__ __forceinline__ uint64_t globaltimer() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CG>
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
                     : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint64_t* bar) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else  // arrive on the leader CTA's barrier only (mask bit 0)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         smem_u32(bar)),
                     "h"(static_cast<uint16_t>(1))
                     : "memory");
}

__host__ __device__ inline uint64_t sdesc_sw128(uint32_t saddr, uint32_t sbo) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;                      // LBO = 16 B (unused for K-major SW128)
    d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}
__host__ __device__ inline uint32_t idesc_f16(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                                // fp32 accumulate, fp16 x fp16, both K-major
    d |= static_cast<uint32_t>(N >> 3) << 17;
    d |= static_cast<uint32_t>(M >> 4) << 24;
    return d;
}

struct ProbeParams {
    int n;        // MMA N
    int a_sbo;    // stride-byte-offset of the A descriptor (1024 canonical, 2560 = 20-pixel halo row)
    int tiles;    // M tiles per group (accumulators at d + t*128 columns ... or t*N if it fits)
    int ksteps;   // K16 steps per tile in a group
    int rounds;   // groups issued
    int depth;    // groups in flight
    int walk;     // 1: move the A start by one 128-byte row per group (conv tap walk), 0: fixed
    int inter;    // 1: K step outer, tile inner (consecutive MMAs hit different accumulators)
    unsigned long long* out;  // per CTA: clocks, ns, ok
};

template <int CG>
__global__ void __launch_bounds__(128, 1) probe_kernel(const ProbeParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);           // 16 barriers
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + 256);
    uint8_t* a_base = smem + 1024;                                 // 96 KB: halo-like region
    uint8_t* b_base = smem + 1024 + 96 * 1024;                     // 64 KB: weight-like region
    // pseudo-random small fp16 values in both regions
    {
        uint32_t* w = reinterpret_cast<uint32_t*>(smem + 1024);
        uint32_t s = 0x9E3779B9u * (threadIdx.x + 1) + blockIdx.x;
        for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) {
            s = s * 1664525u + 1013904223u;
            // two fp16 in [-1, 1): sign | exponent 01110/01101.. | mantissa
            uint32_t lo = ((s >> 3) & 0x83FFu) | 0x3800u;
            uint32_t hi = ((s >> 17) & 0x83FFu) | 0x3400u;
            w[i] = lo | (hi << 16);
        }
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (threadIdx.x < 32) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_ptr;
    const bool leader = (CG == 1) || cluster_ctarank() == 0;

    unsigned long long clocks = 0, ns = 0, ok = 1;
    if (threadIdx.x == 0 && leader) {
        const uint32_t idesc = idesc_f16(CG == 2 ? 256 : 128, p.n);
        const uint64_t a0 = sdesc_sw128(smem_u32(a_base), p.a_sbo);
        const uint64_t b0 = sdesc_sw128(smem_u32(b_base), 1024);
        const int dstep = (p.tiles * p.n <= 512) ? p.n : 128;   // accumulator column step between tiles
        long long c0 = clock64();
        uint64_t t0 = globaltimer();
        for (int r = 0; r < p.rounds; ++r) {
            const int slot = r % p.depth;
            if (r >= p.depth) {
                if (!mbar_wait(&bars[slot], ((r / p.depth) - 1) & 1)) { ok = 0; break; }
            }
            // A start walks over 25 "taps" (5 rows of 5 pixels) like the conv kernel when walk = 1
            const int tap = p.walk ? (r % 25) : 0;
            const uint32_t a_off = p.walk ? static_cast<uint32_t>(((tap / 5) * (p.a_sbo / 128) + tap % 5) * 8) : 0u;
            const uint32_t b_off = static_cast<uint32_t>((r & 3) * (16384 >> 4));   // 4 weight stages of 16 KB
            if (p.inter) {
                for (int j = 0; j < p.ksteps; ++j)
                    for (int t = 0; t < p.tiles; ++t)
                        umma<CG>(tmem + t * dstep, a0 + a_off + 64 * t + 2 * j, b0 + b_off + 2 * j, idesc, (r | j) ? 1u : 0u);
            } else {
                for (int t = 0; t < p.tiles; ++t)
                    for (int j = 0; j < p.ksteps; ++j)
                        umma<CG>(tmem + t * dstep, a0 + a_off + 64 * t + 2 * j, b0 + b_off + 2 * j, idesc, (r | j) ? 1u : 0u);
            }
            commit<CG>(&bars[slot]);
        }
        // drain: every slot's last phase
        if (ok) {
            for (int s = 0; s < p.depth && s < p.rounds; ++s) {
                int last_r = ((p.rounds - 1 - s) / p.depth) * p.depth + s;   // last round that used slot s
                if (!mbar_wait(&bars[s], (last_r / p.depth) & 1)) { ok = 0; break; }
            }
        }
        uint64_t t1 = globaltimer();
        long long c1 = clock64();
        clocks = static_cast<unsigned long long>(c1 - c0);
        ns = t1 - t0;
        p.out[3 * blockIdx.x + 0] = clocks;
        p.out[3 * blockIdx.x + 1] = ns;
        p.out[3 * blockIdx.x + 2] = ok;
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (threadIdx.x < 32) {
        if (CG == 1)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

static int run(const char* name, int cg, int n, int a_sbo, int tiles, int ksteps, int walk, int rounds, int depth, int num_sms, int inter = 0) {
    const int smem_bytes = 1024 + 1024 + 160 * 1024 + 40 * 1024;  // > 113 KB: one CTA per SM
    int grid = num_sms;
    if (cg == 2) grid &= ~1;
    unsigned long long* out;
    CK(cudaMalloc(&out, sizeof(unsigned long long) * 3 * grid));
    CK(cudaMemset(out, 0, sizeof(unsigned long long) * 3 * grid));
    ProbeParams p{n, a_sbo, tiles, ksteps, rounds, depth, walk, inter, out};
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        if (cg == 1) {
            CK(cudaFuncSetAttribute(probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
            probe_kernel<1><<<grid, 128, smem_bytes>>>(p);
        } else {
            CK(cudaFuncSetAttribute(probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid);
            cfg.blockDim = dim3(128);
            cfg.dynamicSmemBytes = smem_bytes;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            CK(cudaLaunchKernelEx(&cfg, probe_kernel<2>, p));
        }
        CK(cudaEventRecord(e1));
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
            printf("{\"variant\": \"%s\", \"error\": \"%s\"}\n", name, cudaGetErrorString(e));
            return 1;
        }
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best_ms) best_ms = ms;
    }
    unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long) * 3 * grid);
    CK(cudaMemcpy(h, out, sizeof(unsigned long long) * 3 * grid, cudaMemcpyDeviceToHost));
    int issuers = 0, bad = 0;
    double clk_sum = 0, ns_sum = 0, clk_max = 0;
    for (int i = 0; i < grid; ++i) {
        if (h[3 * i + 0] == 0) continue;
        ++issuers;
        if (!h[3 * i + 2]) ++bad;
        clk_sum += (double)h[3 * i + 0];
        ns_sum += (double)h[3 * i + 1];
        if ((double)h[3 * i + 0] > clk_max) clk_max = (double)h[3 * i + 0];
    }
    const double mmas = (double)rounds * tiles * ksteps;
    const int M = cg == 2 ? 256 : 128;
    const double flop_per_mma = 2.0 * M * n * 16;
    const double clk_per = clk_sum / issuers / mmas, ns_per = ns_sum / issuers / mmas;
    const double floor_clk = 128.0 * n / 256.0;   // per SM (cta_group::2: same clocks, two SMs)
    printf("{\"variant\": \"%s\", \"cta_group\": %d, \"M\": %d, \"N\": %d, \"a_sbo\": %d, \"tiles\": %d, \"ksteps\": %d, \"walk\": %d, \"interleaved\": %d, "
           "\"rounds\": %d, \"depth\": %d, \"issuers\": %d, \"timeouts\": %d, \"clk_per_mma\": %.2f, \"floor_clk\": %.1f, "
           "\"ns_per_mma\": %.3f, \"sm_ghz\": %.3f, \"tflops_in_kernel\": %.1f, \"kernel_ms\": %.4f}\n",
           name, cg, M, n, a_sbo, tiles, ksteps, walk, inter, rounds, depth, issuers, bad, clk_per, floor_clk, ns_per, clk_per / ns_per,
           issuers * flop_per_mma / ns_per * 1e-3, best_ms);
    fflush(stdout);
    free(h);
    CK(cudaFree(out));
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        printf("{\"error\": \"needs sm_100, found sm_%d%d\"}\n", prop.major, prop.minor);
        return 1;
    }
    const int sms = prop.multiProcessorCount;
    const int R = 4000;   // groups of 8 MMAs: ~1.8 M clocks at the floor for N = 112
    const bool with_pair = !(argc > 1 && strcmp(argv[1], "--no-pair") == 0);
    // the conv kernel's shape: 2 tiles x 4 K steps per tap, halo pitch 20 pixels, start walks over the taps
    run("conv_like_n112_sbo2560_walk", 1, 112, 2560, 2, 4, 1, R, 4, sms);
    run("conv_like_n112_sbo2560_fixed", 1, 112, 2560, 2, 4, 0, R, 4, sms);
    run("canonical_n112_sbo1024", 1, 112, 1024, 2, 4, 0, R, 4, sms);
    run("canonical_n112_one_tile", 1, 112, 1024, 1, 4, 0, R, 4, sms);
    run("canonical_n64", 1, 64, 1024, 2, 4, 0, R, 4, sms);
    run("canonical_n128", 1, 128, 1024, 2, 4, 0, R, 4, sms);
    run("canonical_n224", 1, 224, 1024, 2, 4, 0, R, 4, sms);
    run("canonical_n256", 1, 256, 1024, 2, 4, 0, R, 4, sms);
    run("conv_like_n112_interleaved", 1, 112, 2560, 2, 4, 1, R, 4, sms, 1);
    run("conv_like_n112_4tiles_interleaved", 1, 112, 2560, 4, 4, 1, R / 2, 4, sms, 1);
    run("canonical_n112_4tiles", 1, 112, 1024, 4, 4, 0, R / 2, 4, sms, 0);
    run("conv_like_n112_depth8", 1, 112, 2560, 2, 4, 1, R, 8, sms);
    run("conv_like_n112_group32", 1, 112, 2560, 2, 16, 0, R / 4, 4, sms);
    if (with_pair) {
        run("pair_m256_n112_sbo2560_walk", 2, 112, 2560, 2, 4, 1, R, 4, sms);
        run("pair_m256_n112_canonical", 2, 112, 1024, 2, 4, 0, R, 4, sms);
        run("pair_m256_n224_canonical", 2, 224, 1024, 2, 4, 0, R, 4, sms);
        run("pair_m256_n256_canonical", 2, 256, 1024, 1, 4, 0, R, 4, sms);
    }
    return 0;
}
