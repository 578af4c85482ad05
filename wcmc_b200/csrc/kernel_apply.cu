// K4/K5: softmax over the k*k predicted weights + weighted gather of the noisy radiance
// neighbourhood (forward and backward).  B200-native replacement of
// sbmc.modules.KernelApply(softmax=True, splat=False) and of the Halide `kernel_weighting`
// / `kernel_weighting_grad` ops it wraps (called from
// sbmc.KPCN.forward, which /root/reference/support/interfaces.py:203-204 runs every step).
//
// HBM-bound on paper (1764 B of logits per pixel against 24 B of everything else) but at
// ~7 B per issued instruction the first version was issue-bound (ncu: 400 warp instructions per
// pixel, IPC 1.9, 30 % of DRAM peak), so this version is written for instruction count:
//   * one warp owns one pixel, lane l handles taps k = l, l+32, ...: logit reads are 128-byte
//     coalesced streaming loads and max / sum-exp / weighted sums are warp reductions;
//   * the halo of the radiance buffer sits in shared memory as one float4 {c0,c1,c2,c3} per
//     position, so a tap costs ONE 16-byte shared load (consecutive taps -> consecutive positions:
//     conflict free) instead of one load per channel; the row pitch is == k (mod 8), see ka_pitch;
//   * the k = 21 instantiation has 13 full tap slots + one 25-lane tail: no per-slot predicates;
//   * exp via ex2.approx; the four warp sums share a 7-shuffle transposing butterfly;
//   * 16 x 8 pixel tiles and <= 64 registers: 4 CTAs (32 warps) per SM, each warp with its 14 loads
//     in flight while it waits, which is what Little's law asks for at HBM3e latency.
// The logits are read exactly once (softmax is fused; probabilities never reach HBM); the
// backward recomputes them from the saved (max, 1/sum) pair.
#include "common.cuh"

namespace wcmc {

constexpr int kKaTileH = 8;
constexpr int kKaThreads = 256;  // 8 warps; warp w owns tile row w
constexpr int kKaMaxSlots = 14;  // ceil(441 / 32)

// Row pitch of the halo in float4 positions.  A 16-byte shared load is served a quarter warp at a
// time (8 lanes x 16 B = all 32 banks), so 8 consecutive taps must hit 8 distinct positions mod 8;
// inside a kernel row they are consecutive, across a row wrap the position jumps by pitch - (ks-1):
// pitch == ks (mod 8) keeps that jump == 1 (mod 8).
__host__ __device__ inline int ka_pitch(int ks, int tile_w) {
    int p = tile_w + ks - 1;
    while ((p - ks) & 7) ++p;
    return p;
}

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Halo as float4 per position; rows by warp, columns by lane: coalesced, no div / mod.
template <int C, int TW>
__device__ __forceinline__ void ka_load_halo(float4* sm, const float* __restrict__ data, int n, int H,
                                             int W, int y0, int x0, int ks, int pitch) {
    const int r = ks >> 1;
    const int hh = kKaTileH + ks - 1, hw = TW + ks - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t plane = static_cast<size_t>(H) * W;
    const float* base = data + static_cast<size_t>(n) * C * plane;
    for (int yy = warp; yy < hh; yy += kKaThreads / 32) {
        const int gy = y0 + yy - r;
        const bool yin = gy >= 0 && gy < H;
        for (int xx = lane; xx < hw; xx += 32) {
            const int gx = x0 + xx - r;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (yin && gx >= 0 && gx < W) {
#pragma unroll
                for (int c = 0; c < C; ++c) v[c] = __ldg(base + c * plane + static_cast<size_t>(gy) * W + gx);
            }
            sm[yy * pitch + xx] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// Sums four per-lane values over the warp with 7 shuffles (instead of 20): afterwards lane l holds
// the warp total of value number (l & 3).
__device__ __forceinline__ float warp_sum4(float a, float b, float c, float d) {
    const int lane = threadIdx.x & 31;
    const bool o1 = lane & 1;
    float k0 = o1 ? b : a, s0 = o1 ? a : b;
    float k1 = o1 ? d : c, s1 = o1 ? c : d;
    k0 += __shfl_xor_sync(0xffffffffu, s0, 1);   // even lanes: a summed over the pair; odd lanes: b
    k1 += __shfl_xor_sync(0xffffffffu, s1, 1);   // even lanes: c; odd lanes: d
    const bool o2 = lane & 2;
    float k = o2 ? k1 : k0, s = o2 ? k0 : k1;
    k += __shfl_xor_sync(0xffffffffu, s, 2);     // lane & 3 = 0:a 1:b 2:c 3:d, summed over 4 lanes
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) k += __shfl_xor_sync(0xffffffffu, k, o);
    return k;
}

// KS > 0: compile-time kernel size (no slot predicates); KS == 0: runtime `ks`.
template <int C, int KS, int TW>
__global__ void __launch_bounds__(kKaThreads, 4)
kernel_apply_fwd_kernel(const float* __restrict__ logits, int l_cs, const float* __restrict__ data,
                        float* __restrict__ out, float* __restrict__ stats, int N, int H, int W, int ks_rt) {
    wcmc::pdl_start();
    extern __shared__ float4 sm4[];
    const int ks = KS > 0 ? KS : ks_rt;
    const int taps = ks * ks;
    constexpr int NS = KS > 0 ? (KS * KS + 31) / 32 : kKaMaxSlots;   // tap slots per lane
    const int full = taps >> 5;                                         // slots valid for every lane
    const int pitch = ka_pitch(ks, TW);
    const int n = blockIdx.z, y0 = blockIdx.y * kKaTileH, x0 = blockIdx.x * TW;
    ka_load_halo<C, TW>(sm4, data, n, H, W, y0, x0, ks, pitch);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int y = y0 + warp;
    if (y >= H) return;
    int off[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        int k = lane + 32 * j;
        int dy = k / ks, dx = k - dy * ks;
        off[j] = (k < taps) ? (warp + dy) * pitch + dx : 0;
    }
    const float kLog2e = 1.4426950408889634f;
    const int xe = min(TW, W - x0);
    const float* row = logits + ((static_cast<size_t>(n) * H + y) * W + x0) * l_cs;
    const size_t cplane = static_cast<size_t>(H) * W;
    float* orow = out + static_cast<size_t>(n) * C * cplane + static_cast<size_t>(y) * W + x0;
    auto load = [&](float (&z)[NS], int tx) {
        const float* lp = row + static_cast<size_t>(tx) * l_cs + lane;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if (j < full) z[j] = __ldcs(lp + 32 * j);
            else z[j] = (lane + 32 * j < taps) ? __ldcs(lp + 32 * j) : -INFINITY;
        }
    };
    auto compute = [&](const float (&z)[NS], int tx) {
        float mx = z[0];
#pragma unroll
        for (int j = 1; j < NS; ++j) mx = fmaxf(mx, z[j]);
        mx = warp_max(mx);
        float s = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float mxs = mx * kLog2e;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const float e = ex2(fmaf(z[j], kLog2e, -mxs));  // exp(z - max); 0 for the padded taps
            const float4 v = sm4[off[j] + tx];
            s += e;
            a0 = fmaf(e, v.x, a0);
            if (C > 1) a1 = fmaf(e, v.y, a1);
            if (C > 2) a2 = fmaf(e, v.z, a2);
            if (C > 3) a3 = fmaf(e, v.w, a3);
        }
        const float r = warp_sum4(a0, a1, a2, s);   // lane l: total of (a0, a1, a2, s)[l & 3]
        const float inv = 1.f / __shfl_sync(0xffffffffu, r, 3);
        if (C > 3) a3 = warp_sum(a3);
        if (lane < 3 && lane < C) orow[lane * cplane + tx] = r * inv;
        if (C > 3 && lane == 3) orow[3 * cplane + tx] = a3 * inv;
        if (stats != nullptr && lane == 0)
            reinterpret_cast<float2*>(stats)[(static_cast<size_t>(n) * H + y) * W + x0 + tx] = make_float2(mx, inv);
    };
    // No software prefetch: two register sets cost an occupancy step (126 vs 64 registers); 4 CTAs x
    // 8 warps per SM, each with its 14 loads in flight while it waits, cover the HBM latency instead.
    float z[NS];
#pragma unroll 1
    for (int tx = 0; tx < xe; ++tx) {
        load(z, tx);
        compute(z, tx);
    }
}

template <int C, int DT, int KS, int TW>
__global__ void __launch_bounds__(kKaThreads, 4)
kernel_apply_bwd_kernel(const float* __restrict__ logits, int l_cs, const float* __restrict__ data,
                        const float* __restrict__ out, const float* __restrict__ stats,
                        const float* __restrict__ gout, void* __restrict__ dlogits, int dl_cs, int N, int H,
                        int W, int ks_rt, const float* __restrict__ scale) {
    wcmc::pdl_start();
    constexpr bool H16 = DT != WCMC_F32;
    const float sc = scale != nullptr ? __ldg(scale) : 1.f;
    extern __shared__ float4 sm4[];
    const int ks = KS > 0 ? KS : ks_rt;
    const int taps = ks * ks;
    constexpr int NS = KS > 0 ? (KS * KS + 31) / 32 : kKaMaxSlots;
    const int full = taps >> 5;
    const int pitch = ka_pitch(ks, TW);
    const int n = blockIdx.z, y0 = blockIdx.y * kKaTileH, x0 = blockIdx.x * TW;
    ka_load_halo<C, TW>(sm4, data, n, H, W, y0, x0, ks, pitch);
    // per-warp staging row for coalesced 16-byte stores of d_logits
    uint32_t* stage_all = reinterpret_cast<uint32_t*>(sm4 + (kKaTileH + ks - 1) * pitch);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int y = y0 + warp;
    if (y >= H) return;
    const int stage_elems = H16 ? dl_cs / 2 : dl_cs;  // in 4-byte words
    uint32_t* stage = stage_all + warp * stage_elems;
    // channels k*k .. dl_cs-1 of d_logits are zero: written to the staging row once
    for (int k = taps + lane; k < dl_cs; k += 32) {
        if (DT == WCMC_F32) reinterpret_cast<float*>(stage)[k] = 0.f;
        else reinterpret_cast<uint16_t*>(stage)[k] = 0;
    }
    int off[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        int k = lane + 32 * j;
        int dy = k / ks, dx = k - dy * ks;
        off[j] = (k < taps) ? (warp + dy) * pitch + dx : 0;
    }
    const float kLog2e = 1.4426950408889634f;
    const int xe = min(TW, W - x0);
    const size_t pix0 = (static_cast<size_t>(n) * H + y) * W + x0;
    const size_t cplane = static_cast<size_t>(H) * W;
    const size_t o0 = static_cast<size_t>(n) * C * cplane + static_cast<size_t>(y) * W + x0;
    // per pixel: logits z, (max, 1 / sum), upstream gradient g (float4) and g . out
    auto load = [&](float (&z)[NS], float2& st, float4& g, float& go, int tx) {
        const float* lp = logits + (pix0 + tx) * l_cs + lane;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if (j < full) z[j] = __ldcs(lp + 32 * j);
            else z[j] = (lane + 32 * j < taps) ? __ldcs(lp + 32 * j) : -INFINITY;
        }
        st = __ldg(reinterpret_cast<const float2*>(stats) + pix0 + tx);
        float gg[4] = {0.f, 0.f, 0.f, 0.f};
        go = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            gg[c] = __ldg(gout + o0 + c * cplane + tx);
            go = fmaf(gg[c], __ldg(out + o0 + c * cplane + tx), go);
        }
        g = make_float4(gg[0], gg[1], gg[2], gg[3]);
    };
    auto compute = [&](const float (&z)[NS], const float2 st, const float4 g, const float go, int tx) {
        const float mxs = st.x * kLog2e;
        const float w = sc * st.y;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const float pk = ex2(fmaf(z[j], kLog2e, -mxs)) * w;   // 0 for the padded taps (z = -inf)
            const float4 v = sm4[off[j] + tx];
            float a = fmaf(g.x, v.x, -go);
            if (C > 1) a = fmaf(g.y, v.y, a);
            if (C > 2) a = fmaf(g.z, v.z, a);
            if (C > 3) a = fmaf(g.w, v.w, a);
            const float d = pk * a;
            const int k = lane + 32 * j;
            if (j < full || k < taps) {
                if (DT == WCMC_BF16) reinterpret_cast<__nv_bfloat16*>(stage)[k] = __float2bfloat16_rn(d);
                else if (DT == WCMC_F16) reinterpret_cast<__half*>(stage)[k] = __float2half_rn(d);
                else reinterpret_cast<float*>(stage)[k] = d;
            }
        }
        __syncwarp();
        {
            const int nvec = stage_elems / 4;  // uint4 per pixel row
            uint4* dst = reinterpret_cast<uint4*>(static_cast<uint8_t*>(dlogits) +
                                                  (pix0 + tx) * static_cast<size_t>(dl_cs) * (H16 ? 2 : 4));
            const uint4* src = reinterpret_cast<const uint4*>(stage);
            for (int i = lane; i < nvec; i += 32) __stcs(dst + i, src[i]);
        }
        __syncwarp();
    };
    float z[NS], go;
    float2 st;
    float4 g;
#pragma unroll 1
    for (int tx = 0; tx < xe; ++tx) {
        load(z, st, g, go, tx);
        compute(z, st, g, go, tx);
    }
}

}  // namespace wcmc

using namespace wcmc;

static int g_ka_tile_w = 16;   // tuning hook: wcmc_tuning_set("ka_tile_w", 16 | 32)

int wcmc_ka_set_tile(int w) {
    if (w != 8 && w != 16 && w != 32) return -1;
    g_ka_tile_w = w;
    return 0;
}

static size_t ka_smem_bytes(int ks, int tile_w, int stage_words_per_warp) {
    int pitch = ka_pitch(ks, tile_w);
    return static_cast<size_t>(kKaTileH + ks - 1) * pitch * sizeof(float4) +
           static_cast<size_t>(8) * stage_words_per_warp * sizeof(float);
}

template <int C, int KS, int TW>
static int launch_fwd2(const float* logits, int l_cs, const float* data, float* out, float* stats, int N,
                       int H, int W, int ks, cudaStream_t stream) {
    dim3 grid((W + TW - 1) / TW, (H + kKaTileH - 1) / kKaTileH, N);
    size_t smem = ka_smem_bytes(ks, TW, 0);
    WCMC_FUNC_SMEM((kernel_apply_fwd_kernel<C, KS, TW>), static_cast<int>(smem));
    WCMC_LAUNCH((kernel_apply_fwd_kernel<C, KS, TW>), grid, kKaThreads, smem, stream, logits, l_cs, data, out, stats, N, H, W, ks);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

template <int C>
static int launch_fwd(const float* logits, int l_cs, const float* data, float* out, float* stats, int N,
                      int H, int W, int ks, cudaStream_t stream) {
    if (ks == 21) {
        if (g_ka_tile_w == 32) return launch_fwd2<C, 21, 32>(logits, l_cs, data, out, stats, N, H, W, ks, stream);
        if (g_ka_tile_w == 8) return launch_fwd2<C, 21, 8>(logits, l_cs, data, out, stats, N, H, W, ks, stream);
        return launch_fwd2<C, 21, 16>(logits, l_cs, data, out, stats, N, H, W, ks, stream);
    }
    return launch_fwd2<C, 0, 16>(logits, l_cs, data, out, stats, N, H, W, ks, stream);
}

extern "C" int wcmc_kernel_apply_fwd(const float* logits, int l_cs, const float* data, float* out,
                                     float* stats, int N, int C, int H, int W, int ksize, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(ksize >= 1 && ksize <= 21 && (ksize & 1), WCMC_ESHAPE, "kernel_apply: ksize %d must be odd and <= 21", ksize);
    WCMC_REQUIRE(l_cs >= ksize * ksize, WCMC_ESHAPE, "kernel_apply: logits channel stride %d < k*k", l_cs);
    WCMC_REQUIRE(C >= 1 && C <= 4, WCMC_ESHAPE, "kernel_apply: C=%d not in [1,4]", C);
    WCMC_REQUIRE(N > 0 && H > 0 && W > 0 && N <= 65535, WCMC_ESHAPE, "kernel_apply: bad N/H/W");
    switch (C) {
        case 1: return launch_fwd<1>(logits, l_cs, data, out, stats, N, H, W, ksize, stream);
        case 2: return launch_fwd<2>(logits, l_cs, data, out, stats, N, H, W, ksize, stream);
        case 3: return launch_fwd<3>(logits, l_cs, data, out, stats, N, H, W, ksize, stream);
        default: return launch_fwd<4>(logits, l_cs, data, out, stats, N, H, W, ksize, stream);
    }
}

template <int C, int DT, int KS, int TW>
static int launch_bwd2(const float* logits, int l_cs, const float* data, const float* out, const float* stats,
                       const float* gout, void* dl, int dl_cs, int N, int H, int W, int ks, const float* scale,
                       cudaStream_t stream) {
    dim3 grid((W + TW - 1) / TW, (H + kKaTileH - 1) / kKaTileH, N);
    size_t smem = ka_smem_bytes(ks, TW, DT != WCMC_F32 ? dl_cs / 2 : dl_cs);
    WCMC_FUNC_SMEM((kernel_apply_bwd_kernel<C, DT, KS, TW>), static_cast<int>(smem));
    WCMC_LAUNCH((kernel_apply_bwd_kernel<C, DT, KS, TW>), grid, kKaThreads, smem, stream, logits, l_cs, data, out, stats, gout,
                                                                               dl, dl_cs, N, H, W, ks, scale);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

template <int C, int DT>
static int launch_bwd(const float* logits, int l_cs, const float* data, const float* out, const float* stats,
                      const float* gout, void* dl, int dl_cs, int N, int H, int W, int ks, const float* scale,
                      cudaStream_t stream) {
    if (ks == 21) {
        if (g_ka_tile_w == 32)
            return launch_bwd2<C, DT, 21, 32>(logits, l_cs, data, out, stats, gout, dl, dl_cs, N, H, W, ks, scale, stream);
        if (g_ka_tile_w == 8)
            return launch_bwd2<C, DT, 21, 8>(logits, l_cs, data, out, stats, gout, dl, dl_cs, N, H, W, ks, scale, stream);
        return launch_bwd2<C, DT, 21, 16>(logits, l_cs, data, out, stats, gout, dl, dl_cs, N, H, W, ks, scale, stream);
    }
    return launch_bwd2<C, DT, 0, 16>(logits, l_cs, data, out, stats, gout, dl, dl_cs, N, H, W, ks, scale, stream);
}

extern "C" int wcmc_kernel_apply_bwd(const float* logits, int l_cs, const float* data, const float* out,
                                     const float* stats, const float* grad_out, void* d_logits, int dl_cs,
                                     int dl_dtype, int N, int C, int H, int W, int ksize, const float* scale,
                                     void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(ksize >= 1 && ksize <= 21 && (ksize & 1), WCMC_ESHAPE, "kernel_apply: ksize %d must be odd and <= 21", ksize);
    WCMC_REQUIRE(l_cs >= ksize * ksize && dl_cs >= ksize * ksize && dl_cs <= 448, WCMC_ESHAPE,
                 "kernel_apply_bwd: channel strides (%d,%d) must cover k*k and be <= 448", l_cs, dl_cs);
    WCMC_REQUIRE(dl_cs % 8 == 0, WCMC_ESHAPE, "kernel_apply_bwd: dl_cs %d must be a multiple of 8", dl_cs);
    WCMC_REQUIRE(C >= 1 && C <= 4, WCMC_ESHAPE, "kernel_apply: C=%d not in [1,4]", C);
    WCMC_REQUIRE(N > 0 && H > 0 && W > 0 && N <= 65535, WCMC_ESHAPE, "kernel_apply: bad N/H/W");
    WCMC_REQUIRE(stats != nullptr, WCMC_ESHAPE, "kernel_apply_bwd: stats from the forward pass are required");
    WCMC_REQUIRE(dl_dtype >= 0 && dl_dtype <= 2, WCMC_ESHAPE, "kernel_apply_bwd: bad dl_dtype %d", dl_dtype);
#define WCMC_KA_BWD(CC)                                                                                       \
    switch (dl_dtype) {                                                                                       \
        case WCMC_BF16:                                                                                       \
            return launch_bwd<CC, WCMC_BF16>(logits, l_cs, data, out, stats, grad_out, d_logits, dl_cs, N, H, W,  \
                                             ksize, scale, stream);                                           \
        case WCMC_F16:                                                                                        \
            return launch_bwd<CC, WCMC_F16>(logits, l_cs, data, out, stats, grad_out, d_logits, dl_cs, N, H, W,   \
                                            ksize, scale, stream);                                            \
        default:                                                                                              \
            return launch_bwd<CC, WCMC_F32>(logits, l_cs, data, out, stats, grad_out, d_logits, dl_cs, N, H, W,   \
                                            ksize, scale, stream);                                            \
    }
    switch (C) {
        case 1: WCMC_KA_BWD(1);
        case 2: WCMC_KA_BWD(2);
        case 3: WCMC_KA_BWD(3);
        default: WCMC_KA_BWD(4);
    }
#undef WCMC_KA_BWD
}
