// K4/K5: softmax over the k*k predicted weights + weighted gather of the noisy radiance
// neighbourhood (forward and backward).  B200-native replacement of
// sbmc.modules.KernelApply(softmax=True, splat=False) and of the Halide `kernel_weighting`
// / `kernel_weighting_grad` ops it wraps (called from
// sbmc.KPCN.forward, which /root/reference/support/interfaces.py:203-204 runs every step).
//
// HBM-bound: 1764 B of logits per pixel against 24 B of everything else.  One warp owns one
// pixel: lane l handles taps k = l, l+32, ... so the warp's logit reads are 128-byte coalesced
// and max / sum-exp / the three weighted sums are warp-shuffle reductions.  The (8+k-1) x
// (32+k-1) halo of the radiance buffer is staged once per CTA in shared memory with a row pitch
// == k (mod 32), which makes tap k hit bank (k mod 32): conflict free for consecutive taps.
// The logits are read exactly once (softmax is fused; probabilities never reach HBM); the
// backward recomputes them from the saved (max, 1/sum) pair.  Each warp keeps the NEXT pixel's
// 14 coalesced 128-byte logit loads in flight while it reduces the current one (two register
// sets), and 3 CTAs are resident per SM: ~24 warps x 3.5 KB of loads in flight per SM, which is what
// Little's law asks for at HBM3e latency.  Logits use streaming loads (read once, never reused).
#include "common.cuh"

namespace wcmc {

constexpr int kKaTileW = 16;     // 16 x 8 pixel tiles: 576 CTAs at 8 x 92 x 92 (~4 resident per SM)
constexpr int kKaTileH = 8;
constexpr int kKaThreads = 256;  // 8 warps; warp w owns tile row w
constexpr int kKaMaxSlots = 14;  // ceil(441 / 32)

__host__ __device__ inline int ka_pitch(int ks) {
    int p = ks;  // pitch == ks (mod 32) and >= tile_w + ks - 1
    while (p < kKaTileW + ks - 1) p += 32;
    return p;
}

template <int C>
__device__ __forceinline__ void ka_load_halo(float* sm, const float* __restrict__ data, int n, int H,
                                             int W, int y0, int x0, int ks, int pitch) {
    const int r = ks >> 1;
    const int hh = kKaTileH + ks - 1, hw = kKaTileW + ks - 1;
    const int plane = hh * pitch;
    for (int i = threadIdx.x; i < C * hh * hw; i += kKaThreads) {
        int c = i / (hh * hw);
        int rem = i - c * hh * hw;
        int yy = rem / hw, xx = rem - yy * hw;
        int gy = y0 + yy - r, gx = x0 + xx - r;
        float v = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W)
            v = __ldg(data + ((static_cast<size_t>(n) * C + c) * H + gy) * W + gx);
        sm[c * plane + yy * pitch + xx] = v;
    }
}

template <int C>
__global__ void __launch_bounds__(kKaThreads, 3)
kernel_apply_fwd_kernel(const float* __restrict__ logits, int l_cs, const float* __restrict__ data,
                        float* __restrict__ out, float* __restrict__ stats, int N, int H, int W, int ks) {
    extern __shared__ float sm[];
    const int taps = ks * ks;
    const int pitch = ka_pitch(ks);
    const int plane = (kKaTileH + ks - 1) * pitch;
    const int n = blockIdx.z, y0 = blockIdx.y * kKaTileH, x0 = blockIdx.x * kKaTileW;
    ka_load_halo<C>(sm, data, n, H, W, y0, x0, ks, pitch);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int y = y0 + warp;
    if (y >= H) return;
    int off[kKaMaxSlots];
#pragma unroll
    for (int j = 0; j < kKaMaxSlots; ++j) {
        int k = lane + 32 * j;
        int dy = k / ks, dx = k - dy * ks;
        off[j] = (k < taps) ? (warp + dy) * pitch + dx : 0;
    }
    const float kLog2e = 1.4426950408889634f;
    const int xe = min(kKaTileW, W - x0);
    const float* row = logits + ((static_cast<size_t>(n) * H + y) * W + x0) * l_cs;
    auto load = [&](float (&z)[kKaMaxSlots], int tx) {
        const float* lp = row + static_cast<size_t>(tx) * l_cs;
#pragma unroll
        for (int j = 0; j < kKaMaxSlots; ++j) {
            int k = lane + 32 * j;
            z[j] = (k < taps) ? __ldcs(lp + k) : -INFINITY;
        }
    };
    auto compute = [&](const float (&z)[kKaMaxSlots], int tx) {
        float mx = z[0];
#pragma unroll
        for (int j = 1; j < kKaMaxSlots; ++j) mx = fmaxf(mx, z[j]);
        mx = warp_max(mx);
        float s = 0.f, acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 0.f;
        const float mxs = mx * kLog2e;
#pragma unroll
        for (int j = 0; j < kKaMaxSlots; ++j) {
            float e = exp2f(fmaf(z[j], kLog2e, -mxs));  // exp(z - max); 0 for the padded taps
            s += e;
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = fmaf(e, sm[c * plane + off[j] + tx], acc[c]);
        }
        s = warp_sum(s);
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = warp_sum(acc[c]);
        const float inv = 1.f / s;
        if (lane < C) {
            float v = acc[0];
#pragma unroll
            for (int c = 1; c < C; ++c) v = (lane == c) ? acc[c] : v;
            out[((static_cast<size_t>(n) * C + lane) * H + y) * W + x0 + tx] = v * inv;
        }
        if (stats != nullptr && lane == 0) {
            float2* sp = reinterpret_cast<float2*>(stats) + (static_cast<size_t>(n) * H + y) * W + x0 + tx;
            *sp = make_float2(mx, inv);
        }
    };
    float za[kKaMaxSlots], zb[kKaMaxSlots];
    load(za, 0);
    for (int tx = 0; tx < xe; tx += 2) {
        if (tx + 1 < xe) load(zb, tx + 1);
        compute(za, tx);
        if (tx + 2 < xe) load(za, tx + 2);
        if (tx + 1 < xe) compute(zb, tx + 1);
    }
}

template <int C, int DT>
__global__ void __launch_bounds__(kKaThreads, 2)
kernel_apply_bwd_kernel(const float* __restrict__ logits, int l_cs, const float* __restrict__ data,
                        const float* __restrict__ out, const float* __restrict__ stats,
                        const float* __restrict__ gout, void* __restrict__ dlogits, int dl_cs, int N, int H,
                        int W, int ks, const float* __restrict__ scale) {
    constexpr bool H16 = DT != WCMC_F32;
    const float sc = scale != nullptr ? __ldg(scale) : 1.f;
    extern __shared__ float sm[];
    const int taps = ks * ks;
    const int pitch = ka_pitch(ks);
    const int plane = (kKaTileH + ks - 1) * pitch;
    const int n = blockIdx.z, y0 = blockIdx.y * kKaTileH, x0 = blockIdx.x * kKaTileW;
    ka_load_halo<C>(sm, data, n, H, W, y0, x0, ks, pitch);
    // per-warp staging row for coalesced 16-byte stores of d_logits
    float* stage_all = sm + C * plane;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int y = y0 + warp;
    if (y >= H) return;
    const int stage_elems = H16 ? dl_cs / 2 : dl_cs;  // in 4-byte words
    uint32_t* stage = reinterpret_cast<uint32_t*>(stage_all) + warp * stage_elems;
    int off[kKaMaxSlots];
#pragma unroll
    for (int j = 0; j < kKaMaxSlots; ++j) {
        int k = lane + 32 * j;
        int dy = k / ks, dx = k - dy * ks;
        off[j] = (k < taps) ? (warp + dy) * pitch + dx : 0;
    }
    const float kLog2e = 1.4426950408889634f;
    const int xe = min(kKaTileW, W - x0);
    const int nslots = (dl_cs + 31) / 32;
    const size_t pix0 = (static_cast<size_t>(n) * H + y) * W + x0;
    // per pixel: logits z, (max, 1/sum), upstream gradient g and its dot product with the output
    auto load = [&](float (&z)[kKaMaxSlots], float2& st, float (&g)[C], float& go, int tx) {
        const float* lp = logits + (pix0 + tx) * l_cs;
#pragma unroll
        for (int j = 0; j < kKaMaxSlots; ++j) {
            int k = lane + 32 * j;
            z[j] = (k < taps) ? __ldcs(lp + k) : -INFINITY;
        }
        st = __ldg(reinterpret_cast<const float2*>(stats) + pix0 + tx);
        go = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            size_t o = ((static_cast<size_t>(n) * C + c) * H + y) * W + x0 + tx;
            g[c] = __ldg(gout + o);
            go = fmaf(g[c], __ldg(out + o), go);
        }
    };
    auto compute = [&](const float (&z)[kKaMaxSlots], const float2 st, const float (&g)[C], const float go, int tx) {
        const float mxs = st.x * kLog2e;
#pragma unroll
        for (int j = 0; j < kKaMaxSlots; ++j) {
            if (j < nslots) {
                int k = lane + 32 * j;
                float pk = exp2f(fmaf(z[j], kLog2e, -mxs)) * st.y;
                float a = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) a = fmaf(g[c], sm[c * plane + off[j] + tx], a);
                float d = (k < taps) ? sc * pk * (a - go) : 0.f;
                if (k < dl_cs) {
                    if (DT == WCMC_BF16)
                        reinterpret_cast<__nv_bfloat16*>(stage)[k] = __float2bfloat16_rn(d);
                    else if (DT == WCMC_F16)
                        reinterpret_cast<__half*>(stage)[k] = __float2half_rn(d);
                    else
                        reinterpret_cast<float*>(stage)[k] = d;
                }
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
    float za[kKaMaxSlots], zb[kKaMaxSlots], ga[C], gb[C], goa, gob;
    float2 sta, stb;
    load(za, sta, ga, goa, 0);
    for (int tx = 0; tx < xe; tx += 2) {
        if (tx + 1 < xe) load(zb, stb, gb, gob, tx + 1);
        compute(za, sta, ga, goa, tx);
        if (tx + 2 < xe) load(za, sta, ga, goa, tx + 2);
        if (tx + 1 < xe) compute(zb, stb, gb, gob, tx + 1);
    }
}

}  // namespace wcmc

using namespace wcmc;

static size_t ka_smem_bytes(int C, int ks, int stage_words_per_warp) {
    int pitch = ka_pitch(ks);
    return (static_cast<size_t>(C) * (kKaTileH + ks - 1) * pitch + 8 * stage_words_per_warp) * sizeof(float);
}

template <int C>
static int launch_fwd(const float* logits, int l_cs, const float* data, float* out, float* stats, int N,
                      int H, int W, int ks, cudaStream_t stream) {
    dim3 grid((W + kKaTileW - 1) / kKaTileW, (H + kKaTileH - 1) / kKaTileH, N);
    size_t smem = ka_smem_bytes(C, ks, 0);
    WCMC_CHECK_CUDA(cudaFuncSetAttribute(kernel_apply_fwd_kernel<C>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kernel_apply_fwd_kernel<C><<<grid, kKaThreads, smem, stream>>>(logits, l_cs, data, out, stats, N, H, W, ks);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
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

template <int C, int DT>
static int launch_bwd(const float* logits, int l_cs, const float* data, const float* out, const float* stats,
                      const float* gout, void* dl, int dl_cs, int N, int H, int W, int ks, const float* scale,
                      cudaStream_t stream) {
    dim3 grid((W + kKaTileW - 1) / kKaTileW, (H + kKaTileH - 1) / kKaTileH, N);
    size_t smem = ka_smem_bytes(C, ks, DT != WCMC_F32 ? dl_cs / 2 : dl_cs);
    WCMC_CHECK_CUDA(cudaFuncSetAttribute(kernel_apply_bwd_kernel<C, DT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kernel_apply_bwd_kernel<C, DT><<<grid, kKaThreads, smem, stream>>>(logits, l_cs, data, out, stats, gout, dl,
                                                                       dl_cs, N, H, W, ks, scale);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
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
    WCMC_REQUIRE(dl_dtype >= 0 && dl_dtype <= 2, WCMC_ESHAPE, "kernel_apply_bwd: bad dl_dtype %d", dl_dtype);
    switch (C) {
        case 1: WCMC_KA_BWD(1);
        case 2: WCMC_KA_BWD(2);
        case 3: WCMC_KA_BWD(3);
        default: WCMC_KA_BWD(4);
    }
#undef WCMC_KA_BWD
}
