// Data-parallel gradient exchange over NVSwitch peer memory: sum (x scale) of one region of a SYMMETRIC buffer
// over the ranks of a node, in place, in ONE kernel per rank -- no NCCL on the data path.
//
// The reference's multi-GPU mode is single-process nn.DataParallel (/root/reference/train_kpcn.py:266-269): the
// replicas' gradients are reduced onto GPU 0 by torch's scatter/gather every step.  Here every rank owns a replica;
// its flat fp32 gradients sit in a buffer that every rank maps at the same offsets (torch symmetric memory supplies
// the mappings: plumbing) and, where the switch supports it, in one MULTICAST mapping.  The kernel is a two-shot
// all-reduce:
//     rank r owns slice r of the region (16-byte units, contiguous):
//       multicast:  v = multimem.ld_reduce.add(mc + i)      the switch sums the W replicas' copies on the way in
//                   multimem.st(mc + i, v * scale)           and broadcasts the result on the way out
//       peer loads: v = sum over ranks in rank order of ld(peer[q] + i); st(peer[q] + i, v * scale) for every q
//     bracketed by two cross-rank barriers (flags in the symmetric buffer itself, one slot per (channel, CTA, source
//     rank), monotonic epochs: nothing to reset, so the launch replays inside a CUDA graph).
// Every slice is summed exactly once and broadcast, so the replicas stay bit-identical; the peer-load variant also
// fixes the summation order (rank 0 first), i.e. it is run-to-run deterministic.
//
// Why not NCCL for this step: the exchange overlaps the path-embedding networks' backward pass, whose persistent
// tcgen05 kernels want every SM.  This kernel uses no shared memory to speak of and a few dozen registers per
// thread, so its CTAs CO-RESIDE with the 200 KB convolution CTAs instead of displacing them; two channels keep the
// early (dncnn) and the late (path networks + "all finite" flag) exchange independent while both are in flight.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kXThreads = 256;   // small CTAs with few registers: they fit next to a resident convolution CTA
constexpr int kXMaxBlocks = 128;
constexpr int kXMaxWorld = 16;
constexpr int kXChannels = 2;
// uint32 words at the flag offset: flags [channel][block][source rank], then this rank's launch counters [channel][block]
constexpr int kXFlagWords = kXChannels * kXMaxBlocks * kXMaxWorld + kXChannels * kXMaxBlocks;

int g_blocks = 0;   // 0 = per transport: 64 CTAs feed the multicast path, the peer-load path wants every SM (128)

struct XParams {
    float* local;          // this rank's mapping
    float* mc;             // multicast mapping of the same buffer, or nullptr
    float* const* peers;   // device array [world]: every rank's mapping as seen from this rank (peers[rank] == local)
    long flag_word;        // offset of the flag area in 4-byte words
    long lo4, hi4;         // this rank's slice, in float4 units from the start of the buffer
    int rank, world, chan;
    float scale;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 mc_ld_reduce(const float* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float4 peer_ld(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ void peer_st(float* p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// CTA b of every rank meets CTA b of every other rank.  Thread q < world tells rank q "rank `rank` reached `epoch`"
// and waits until rank q has said the same here.  The release / acquire pair at system scope together with the two
// CTA barriers orders every thread's earlier stores (to any rank) before every thread's later loads on every rank.
__device__ __forceinline__ void cross_rank_barrier(const XParams& P, uint32_t epoch) {
    __syncthreads();
    if (threadIdx.x < P.world) {
        const long slot = P.flag_word + (static_cast<long>(P.chan) * kXMaxBlocks + blockIdx.x) * kXMaxWorld;
        st_release_sys(reinterpret_cast<uint32_t*>(P.peers[threadIdx.x]) + slot + P.rank, epoch);
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(P.local) + slot + threadIdx.x;
        const long long t0 = clock64();
        while (static_cast<int>(ld_acquire_sys(mine) - epoch) < 0) {
            if (clock64() - t0 > 8000000000LL) {   // ~4 s: a rank that never arrives must not hang the box
                printf("wcmc: gradient exchange timed out (rank %d waits for rank %d, channel %d, block %d, epoch %u)\n",
                       P.rank, threadIdx.x, P.chan, blockIdx.x, epoch);
                __trap();
            }
        }
    }
    __syncthreads();
}

// W > 0: compile-time world size (all W x U peer loads of a thread are issued before the first add); W = 0: any world.
template <bool MC, int U, int W>
__global__ void __launch_bounds__(kXThreads, 4) grad_exchange_kernel(const XParams P) {
    __shared__ uint32_t s_epoch;
    if (threadIdx.x == 0) {
        uint32_t* counter = reinterpret_cast<uint32_t*>(P.local) + P.flag_word + kXChannels * kXMaxBlocks * kXMaxWorld +
                            P.chan * kXMaxBlocks + blockIdx.x;
        s_epoch = *counter;
        *counter = s_epoch + 2u;   // two barriers per launch; every rank launches the same sequence of grids
    }
    __syncthreads();
    const uint32_t epoch = s_epoch;
    cross_rank_barrier(P, epoch + 1u);   // every rank's gradients are in its buffer

    const long step = static_cast<long>(gridDim.x) * kXThreads * U;
    for (long base = P.lo4 + static_cast<long>(blockIdx.x) * kXThreads * U + threadIdx.x; base < P.hi4; base += step) {
        float4 v[U];
        if (MC) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long i = base + static_cast<long>(u) * kXThreads;
                if (i < P.hi4) v[u] = mc_ld_reduce(P.mc + 4 * i);
            }
        } else if (W > 0) {
            float4 t[W > 0 ? W : 1][U];
#pragma unroll
            for (int q = 0; q < W; ++q) {
                const float* src = P.peers[q];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const long i = base + static_cast<long>(u) * kXThreads;
                    t[q][u] = i < P.hi4 ? peer_ld(src + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                v[u] = t[0][u];
#pragma unroll
                for (int q = 1; q < W; ++q) {   // rank order: the sum does not depend on who computes it
                    v[u].x += t[q][u].x; v[u].y += t[q][u].y; v[u].z += t[q][u].z; v[u].w += t[q][u].w;
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int q = 0; q < P.world; ++q) {
                const float* src = P.peers[q];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const long i = base + static_cast<long>(u) * kXThreads;
                    if (i < P.hi4) {
                        const float4 t = peer_ld(src + 4 * i);
                        v[u].x += t.x; v[u].y += t.y; v[u].z += t.z; v[u].w += t.w;
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            v[u].x *= P.scale; v[u].y *= P.scale; v[u].z *= P.scale; v[u].w *= P.scale;
        }
        if (MC) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long i = base + static_cast<long>(u) * kXThreads;
                if (i < P.hi4) mc_st(P.mc + 4 * i, v[u]);
            }
        } else {
            for (int q = 0; q < P.world; ++q) {
                float* dst = P.peers[(P.rank + q) % P.world];   // start at home: spreads the ranks over the links
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const long i = base + static_cast<long>(u) * kXThreads;
                    if (i < P.hi4) peer_st(dst + 4 * i, v[u]);
                }
            }
        }
    }
    cross_rank_barrier(P, epoch + 2u);   // every rank's slice has landed here, and nobody still reads this rank's
}

}  // namespace

int wcmc_exchange_set_blocks(int v) {
    if (v < 0 || v > kXMaxBlocks) return 1;
    g_blocks = v;
    return 0;
}

extern "C" size_t wcmc_grad_exchange_flag_bytes(void) { return static_cast<size_t>(kXFlagWords) * 4; }

extern "C" int wcmc_grad_exchange(float* local, float* multicast, float* const* peers_dev, long flag_offset_bytes,
                                  long offset_floats, long n_floats, int rank, int world, int channel, float scale,
                                  void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(local != nullptr && peers_dev != nullptr, WCMC_ESHAPE, "grad_exchange: null buffer");
    WCMC_REQUIRE(world >= 1 && world <= kXMaxWorld && rank >= 0 && rank < world, WCMC_ESHAPE,
                 "grad_exchange: rank %d of %d (at most %d ranks)", rank, world, kXMaxWorld);
    WCMC_REQUIRE(channel >= 0 && channel < kXChannels, WCMC_ESHAPE, "grad_exchange: channel %d", channel);
    WCMC_REQUIRE(n_floats > 0 && n_floats % 4 == 0 && offset_floats >= 0 && offset_floats % 4 == 0, WCMC_ESHAPE,
                 "grad_exchange: region [%ld, +%ld) floats must be a multiple of 4 on both ends", offset_floats, n_floats);
    WCMC_REQUIRE(flag_offset_bytes % 16 == 0 && flag_offset_bytes >= (offset_floats + n_floats) * 4, WCMC_ESHAPE,
                 "grad_exchange: the flag area (byte %ld) must lie behind the region and be 16-byte aligned",
                 flag_offset_bytes);
    WCMC_REQUIRE((reinterpret_cast<uintptr_t>(local) & 15) == 0 && (reinterpret_cast<uintptr_t>(multicast) & 15) == 0,
                 WCMC_EALIGN, "grad_exchange: buffers must be 16-byte aligned");
    XParams P;
    P.local = local;
    P.mc = multicast;
    P.peers = peers_dev;
    P.flag_word = flag_offset_bytes / 4;
    const long n4 = n_floats / 4, off4 = offset_floats / 4;
    const long per = (n4 + world - 1) / world;
    P.lo4 = off4 + std::min(n4, per * rank);
    P.hi4 = off4 + std::min(n4, per * (rank + 1));
    P.rank = rank;
    P.world = world;
    P.chan = channel;
    P.scale = scale;
    // every rank must launch the SAME grid (the barriers pair CTAs by index): it depends on the region and the world only
    const int want = g_blocks > 0 ? g_blocks : (multicast != nullptr ? 64 : kXMaxBlocks);
    const long per_cta = static_cast<long>(kXThreads) * 4;
    const int grid = static_cast<int>(std::max<long>(1, std::min<long>(want, (per + per_cta - 1) / per_cta)));
    if (multicast != nullptr)
        grad_exchange_kernel<true, 4, 0><<<grid, kXThreads, 0, stream>>>(P);
    else if (world == 2)          // peer loads: W x U 16-byte loads in flight per thread
        grad_exchange_kernel<false, 4, 2><<<grid, kXThreads, 0, stream>>>(P);
    else if (world == 4)
        grad_exchange_kernel<false, 2, 4><<<grid, kXThreads, 0, stream>>>(P);
    else if (world == 8)
        grad_exchange_kernel<false, 1, 8><<<grid, kXThreads, 0, stream>>>(P);
    else
        grad_exchange_kernel<false, 4, 0><<<grid, kXThreads, 0, stream>>>(P);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}
