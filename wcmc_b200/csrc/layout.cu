// Layout conversion at the boundary between torch's NCHW fp32 tensors (the reference's tensor
// contract, /root/reference/support/datasets.py:760-793, :1078-1126) and the NHWC bf16
// pipeline of the tcgen05 convolutions; weight packing; wgrad finalisation; bias gradient.
// All bandwidth-bound: transposes go through a padded shared-memory tile so both the global
// reads (pixel-contiguous) and the global writes (channel-contiguous) are coalesced.
#include <algorithm>

#include "common.cuh"

namespace wcmc {

constexpr int kTrPix = 32;
constexpr int kTrThreads = 256;

// grid: (ceil(HW/32), N).  smem: c_fill x 33 floats.
__global__ void __launch_bounds__(kTrThreads)
nchw_to_nhwc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C, int HW,
                    int dst_cs, int dst_coff, int c_fill, int dtype, const float* __restrict__ scale) {
    wcmc::pdl_start();
    extern __shared__ float tile[];
    const float sc = scale != nullptr ? __ldg(scale) : 1.f;
    const int n = blockIdx.y, p0 = blockIdx.x * kTrPix;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = p0 + lane;
    for (int c = warp; c < c_fill; c += kTrThreads / 32) {
        float v = 0.f;
        if (c < C && p < HW) v = sc * __ldg(src + (static_cast<size_t>(n) * C + c) * HW + p);
        tile[c * 33 + lane] = v;
    }
    __syncthreads();
    const int npair = c_fill >> 1;
    for (int pp = warp; pp < kTrPix; pp += kTrThreads / 32) {
        if (p0 + pp >= HW) break;
        uint32_t* drow = reinterpret_cast<uint32_t*>(dst + (static_cast<size_t>(n) * HW + p0 + pp) * dst_cs +
                                                     dst_coff);
        for (int l = lane; l < npair; l += 32)
            drow[l] = pack_h2(tile[(2 * l) * 33 + pp], tile[(2 * l + 1) * 33 + pp], dtype);
    }
}

__global__ void __launch_bounds__(kTrThreads)
nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int C, int HW,
                    int src_cs, int src_coff, int accumulate, int dtype, const float* __restrict__ scale) {
    wcmc::pdl_start();
    extern __shared__ float tile[];
    const float sc = scale != nullptr ? __ldg(scale) : 1.f;
    const int n = blockIdx.y, p0 = blockIdx.x * kTrPix;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int pp = warp; pp < kTrPix; pp += kTrThreads / 32) {
        if (p0 + pp >= HW) break;
        const __nv_bfloat16* srow = src + (static_cast<size_t>(n) * HW + p0 + pp) * src_cs + src_coff;
        for (int c = lane; c < C; c += 32)
            tile[c * 33 + pp] = dtype == WCMC_F16 ? __half2float(reinterpret_cast<const __half*>(srow)[c])
                                                  : __bfloat162float(srow[c]);
    }
    __syncthreads();
    const int p = p0 + lane;
    if (p >= HW) return;
    for (int c = warp; c < C; c += kTrThreads / 32) {
        float* d = dst + (static_cast<size_t>(n) * C + c) * HW + p;
        float v = sc * tile[c * 33 + lane];
        *d = accumulate ? (*d + v) : v;
    }
}

__global__ void pack_weights_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                    __nv_bfloat16* __restrict__ fwd, __nv_bfloat16* __restrict__ dgrad,
                                    float* __restrict__ bias_p, int cout, int cin, int ks, int cout_p,
                                    int cin_p, int dtype) {
    wcmc::pdl_start();
    const int taps = ks * ks;
    const long total = static_cast<long>(cout_p) * taps * cin_p;
    if (bias_p != nullptr && blockIdx.x == 0)
        for (int c = threadIdx.x; c < cout_p; c += blockDim.x)
            bias_p[c] = (bias != nullptr && c < cout) ? bias[c] : 0.f;
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
        if (fwd != nullptr) {  // fwd[co][tap][ci]
            int ci = static_cast<int>(i % cin_p);
            int tap = static_cast<int>((i / cin_p) % taps);
            int co = static_cast<int>(i / (static_cast<long>(cin_p) * taps));
            float v = (co < cout && ci < cin) ? w[(static_cast<long>(co) * cin + ci) * taps + tap] : 0.f;
            if (dtype == WCMC_F16) reinterpret_cast<__half*>(fwd)[i] = __float2half_rn(v);
            else fwd[i] = __float2bfloat16_rn(v);
        }
        if (dgrad != nullptr) {  // dgrad[ci][taps-1-tap][co]
            int co = static_cast<int>(i % cout_p);
            int tapf = static_cast<int>((i / cout_p) % taps);
            int ci = static_cast<int>(i / (static_cast<long>(cout_p) * taps));
            int tap = taps - 1 - tapf;
            float v = (co < cout && ci < cin) ? w[(static_cast<long>(co) * cin + ci) * taps + tap] : 0.f;
            if (dtype == WCMC_F16) reinterpret_cast<__half*>(dgrad)[i] = __float2half_rn(v);
            else dgrad[i] = __float2bfloat16_rn(v);
        }
    }
}

struct PackBatch {
    wcmc_pack_desc d[WCMC_PACK_BATCH_MAX];
};

// blockIdx.y = layer.  Round 1 gathered every packed element from torch's (cout,cin,k,k) layout with stride-k^2
// 4-byte reads (398 GB/s, 0.24 ms of every step).  Here a CTA owns one OUTPUT ROW and stages its source through
// shared memory, so both sides move whole sectors:
//   blockIdx.x <  cout_p : fwd row co   = w[co][:][:]   one contiguous cin*taps run       -> fwd[co][tap][ci]
//   blockIdx.x >= cout_p : dgrad row ci = w[:][ci][:]   cout runs of taps floats          -> dgrad[ci][taps-1-tap][co]
// Rows / columns beyond the logical channel counts are written as zeros (the convolutions rely on it).
__global__ void __launch_bounds__(256) pack_weights_batch_kernel(const PackBatch pb, int dtype) {
    wcmc::pdl_start();
    extern __shared__ float row[];
    const wcmc_pack_desc& L = pb.d[blockIdx.y];
    const int taps = L.ksize * L.ksize;
    const int r = blockIdx.x;
    if (r >= L.cout_p + L.cin_p) return;
    if (r == 0 && L.dst_bias != nullptr)
        for (int c = threadIdx.x; c < L.cout_p; c += blockDim.x)
            L.dst_bias[c] = (L.bias != nullptr && c < L.cout) ? L.bias[c] : 0.f;
    if (r < L.cout_p) {
        if (L.dst_fwd == nullptr) return;
        const int co = r;
        const int n = L.cin * taps;
        if (co < L.cout) {
            const float* src = L.w + static_cast<long>(co) * n;
            for (int i = threadIdx.x; i < n; i += blockDim.x) row[i] = __ldg(src + i);     // [ci][tap]
        }
        __syncthreads();
        const long base = static_cast<long>(co) * taps * L.cin_p;
        const int total = taps * L.cin_p;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int tap = i / L.cin_p, ci = i - tap * L.cin_p;
            const float v = (co < L.cout && ci < L.cin) ? row[ci * taps + tap] : 0.f;
            if (dtype == WCMC_F16) static_cast<__half*>(L.dst_fwd)[base + i] = __float2half_rn(v);
            else static_cast<__nv_bfloat16*>(L.dst_fwd)[base + i] = __float2bfloat16_rn(v);
        }
    } else {
        if (L.dst_dgrad == nullptr) return;
        const int ci = r - L.cout_p;
        if (ci < L.cin) {
            const int n = L.cout * taps;
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int co = i / taps, tap = i - co * taps;
                row[i] = __ldg(L.w + (static_cast<long>(co) * L.cin + ci) * taps + tap);     // [co][tap]
            }
        }
        __syncthreads();
        const long base = static_cast<long>(ci) * taps * L.cout_p;
        const int total = taps * L.cout_p;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int tapf = i / L.cout_p, co = i - tapf * L.cout_p;
            const float v = (co < L.cout && ci < L.cin) ? row[co * taps + (taps - 1 - tapf)] : 0.f;
            if (dtype == WCMC_F16) static_cast<__half*>(L.dst_dgrad)[base + i] = __float2half_rn(v);
            else static_cast<__nv_bfloat16*>(L.dst_dgrad)[base + i] = __float2bfloat16_rn(v);
        }
    }
}

// db[c] (+)= scale * sum_pix dy[pix][coff + c].  One thread owns an 8-channel group (16-byte loads,
// consecutive threads = consecutive groups of one pixel row: coalesced), the remaining thread
// dimension strides over the CTA's slab of pixels; shared-memory reduce, then one atomic per channel.
__global__ void __launch_bounds__(256)
bias_grad_kernel(const __nv_bfloat16* __restrict__ dy, long npix, int cs, int coff, int cout,
                 float* __restrict__ db, int dtype, const float* __restrict__ scale) {
    wcmc::pdl_start();
    extern __shared__ float red[];  // [PL][G*8]
    const int G = (cout + 7) >> 3;
    const int PL = 256 / G;
    const int g = threadIdx.x % G, pl = threadIdx.x / G;
    const long per = (npix + gridDim.x - 1) / gridDim.x;
    const long b = blockIdx.x * per, e = min(npix, b + per);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    if (pl < PL) {
        for (long p = b + pl; p < e; p += PL) {
            uint4 u = __ldg(reinterpret_cast<const uint4*>(dy + p * cs + coff + g * 8));
            uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float2 f = unpack_h2(w[i], dtype);
                acc[2 * i] += f.x;
                acc[2 * i + 1] += f.y;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) red[pl * G * 8 + g * 8 + k] = acc[k];
    }
    __syncthreads();
    const float sc = scale != nullptr ? __ldg(scale) : 1.f;
    for (int c = threadIdx.x; c < cout; c += 256) {
        float s = 0.f;
        for (int q = 0; q < PL; ++q) s += red[q * G * 8 + c];
        atomicAdd(db + c, s * sc);
    }
}

__global__ void zero_f32_kernel(float* p, long n) {
    wcmc::pdl_start();
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long>(gridDim.x) * blockDim.x)
        p[i] = 0.f;
}

}  // namespace wcmc

using namespace wcmc;

extern "C" int wcmc_nchw_f32_to_nhwc(const float* src, void* dst, int dst_dtype, int N, int C, int H, int W,
                                     int dst_cs, int dst_coff, int c_fill, const float* scale, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && N <= 65535, WCMC_ESHAPE, "nchw_to_nhwc: bad shape");
    WCMC_REQUIRE(c_fill >= C && c_fill % 2 == 0 && dst_coff % 2 == 0 && dst_cs % 2 == 0 &&
                     dst_coff + c_fill <= dst_cs,
                 WCMC_ESHAPE, "nchw_to_nhwc: c_fill %d / dst_coff %d / dst_cs %d inconsistent", c_fill, dst_coff,
                 dst_cs);
    const int HW = H * W;
    dim3 grid((HW + kTrPix - 1) / kTrPix, N);
    size_t smem = static_cast<size_t>(c_fill) * 33 * sizeof(float);
    WCMC_REQUIRE(smem <= 200 * 1024, WCMC_ESHAPE, "nchw_to_nhwc: too many channels (%d)", c_fill);
    if (smem > 48 * 1024) WCMC_FUNC_SMEM(nchw_to_nhwc_kernel, static_cast<int>(smem));
    WCMC_LAUNCH(nchw_to_nhwc_kernel, grid, kTrThreads, smem, stream, src, static_cast<__nv_bfloat16*>(dst), C, HW, dst_cs,
                                                            dst_coff, c_fill, dst_dtype, scale);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" int wcmc_nhwc_to_nchw_f32(const void* src, int src_dtype, float* dst, int N, int C, int H, int W,
                                     int src_cs, int src_coff, int accumulate, const float* scale, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && N <= 65535, WCMC_ESHAPE, "nhwc_to_nchw: bad shape");
    WCMC_REQUIRE(src_coff + C <= src_cs, WCMC_ESHAPE, "nhwc_to_nchw: channel slice out of range");
    const int HW = H * W;
    dim3 grid((HW + kTrPix - 1) / kTrPix, N);
    size_t smem = static_cast<size_t>(C) * 33 * sizeof(float);
    WCMC_REQUIRE(smem <= 200 * 1024, WCMC_ESHAPE, "nhwc_to_nchw: too many channels (%d)", C);
    if (smem > 48 * 1024) WCMC_FUNC_SMEM(nhwc_to_nchw_kernel, static_cast<int>(smem));
    WCMC_LAUNCH(nhwc_to_nchw_kernel, grid, kTrThreads, smem, stream, static_cast<const __nv_bfloat16*>(src), dst, C, HW,
                                                            src_cs, src_coff, accumulate, src_dtype, scale);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" int wcmc_pack_weights(const float* w, const float* bias, void* dst_fwd, void* dst_dgrad,
                                 float* dst_bias, int dtype, int cout, int cin, int ksize, int cout_p, int cin_p,
                                 void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(cout > 0 && cin > 0 && cout_p >= cout && cin_p >= cin && ksize > 0, WCMC_ESHAPE,
                 "pack_weights: bad shape");
    long total = static_cast<long>(cout_p) * cin_p * ksize * ksize;
    int blocks = static_cast<int>(std::min<long>((total + 255) / 256, 148 * 8));
    WCMC_LAUNCH(pack_weights_kernel, blocks, 256, 0, stream, w, bias, static_cast<__nv_bfloat16*>(dst_fwd),
                                                    static_cast<__nv_bfloat16*>(dst_dgrad), dst_bias, cout, cin,
                                                    ksize, cout_p, cin_p, dtype);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" int wcmc_pack_weights_batch(const wcmc_pack_desc* descs, int n, int dtype, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(descs != nullptr && n > 0, WCMC_ESHAPE, "pack_weights_batch: empty batch");
    for (int base = 0; base < n; base += WCMC_PACK_BATCH_MAX) {
        PackBatch pb;
        const int m = std::min(WCMC_PACK_BATCH_MAX, n - base);
        int max_rows = 0, max_floats = 0;
        for (int i = 0; i < m; ++i) {
            pb.d[i] = descs[base + i];
            const wcmc_pack_desc& L = pb.d[i];
            WCMC_REQUIRE(L.cout > 0 && L.cin > 0 && L.cout_p >= L.cout && L.cin_p >= L.cin && L.ksize > 0 &&
                             L.w != nullptr,
                         WCMC_ESHAPE, "pack_weights_batch: bad layer %d", base + i);
            max_rows = std::max(max_rows, L.cout_p + L.cin_p);
            max_floats = std::max(max_floats, std::max(L.cout, L.cin) * L.ksize * L.ksize);
        }
        WCMC_REQUIRE(max_floats <= 12 * 1024, WCMC_ESHAPE, "pack_weights_batch: a weight row of %d floats does not fit "
                     "the staging tile", max_floats);
        const int smem = max_floats * static_cast<int>(sizeof(float));
        if (smem > 48 * 1024) WCMC_FUNC_SMEM(pack_weights_batch_kernel, smem);
        dim3 grid(static_cast<unsigned>(max_rows), m);
        WCMC_LAUNCH(pack_weights_batch_kernel, grid, 256, smem, stream, pb, dtype);
        WCMC_LAUNCH_CHECK();
    }
    return WCMC_OK;
}

extern "C" int wcmc_bias_grad(const void* dy, int dy_dtype, int npix, int dy_cs, int dy_coff, int cout, float* db,
                              int accumulate, const float* scale, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(npix > 0 && cout > 0 && dy_coff + cout <= dy_cs && cout <= 1024, WCMC_ESHAPE,
                 "bias_grad: bad shape");
    if (!accumulate) {
        WCMC_LAUNCH(zero_f32_kernel, 1, 256, 0, stream, db, cout);
        WCMC_LAUNCH_CHECK();
    }
    WCMC_REQUIRE(dy_cs % 8 == 0 && dy_coff % 8 == 0 && dy_coff + ((cout + 7) / 8) * 8 <= dy_cs, WCMC_EALIGN,
                 "bias_grad: channel stride/offset must be multiples of 8 and cover cout rounded up to 8");
    int blocks = std::min(592, (npix + 63) / 64);
    const int G = (cout + 7) / 8;
    WCMC_LAUNCH(bias_grad_kernel, blocks, 256, static_cast<size_t>(256 / G) * G * 8 * sizeof(float), stream, static_cast<const __nv_bfloat16*>(dy),
                                                                        npix, dy_cs, dy_coff, cout, db, dy_dtype, scale);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}
