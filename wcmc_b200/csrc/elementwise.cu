// Bandwidth-bound NHWC bf16 glue kernels of the path-embedding network (PathNet,
// /root/reference/support/networks.py:29-42, and the sbmc.modules.Autoencoder U-Net it wraps,
// SURVEY.md Appendix A.3): 2x2 max-pool, 2x bilinear up-sampling (align_corners=False), the
// mean over samples-per-pixel, the broadcast of the propagated feature back to every sample and
// the matching backward passes.  Every tensor is addressed as (pixels, channel stride, channel
// offset) so "concatenation" is just writing into a channel slice of the consumer's buffer --
// the reference's `repeat` + `cat` (networks.py:39-40) is never materialised separately.
// One thread moves 8 channels (16 bytes); consecutive threads take consecutive channel groups
// of the same pixel, so accesses are fully coalesced.
#include <algorithm>

#include "common.cuh"

namespace wcmc {

struct V8 {
    float f[8];
};
__device__ __forceinline__ V8 ld8(const __nv_bfloat16* p, int dt = WCMC_BF16) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    V8 v;
    uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = unpack_h2(w[i], dt);
        v.f[2 * i] = f.x;
        v.f[2 * i + 1] = f.y;
    }
    return v;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const V8& v, int dt = WCMC_BF16) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_h2(v.f[0], v.f[1], dt), pack_h2(v.f[2], v.f[3], dt),
                                              pack_h2(v.f[4], v.f[5], dt), pack_h2(v.f[6], v.f[7], dt));
}

#define WCMC_GRID_STRIDE(i, total)                                                          \
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < (total);    \
         i += static_cast<long>(gridDim.x) * blockDim.x)

// y[n,oy,ox,c] = max over the 2x2 window of x
__global__ void maxpool2_fwd_kernel(const __nv_bfloat16* __restrict__ x, int x_cs, int x_coff,
                                    __nv_bfloat16* __restrict__ y, int y_cs, int y_coff, int N, int H, int W,
                                    int C8, int dt) {
    wcmc::pdl_start();
    const int Ho = H / 2, Wo = W / 2;
    const long total = static_cast<long>(N) * Ho * Wo * C8;
    WCMC_GRID_STRIDE(i, total) {
        int c = static_cast<int>(i % C8) * 8;
        long pix = i / C8;
        int ox = static_cast<int>(pix % Wo);
        int oy = static_cast<int>((pix / Wo) % Ho);
        int n = static_cast<int>(pix / (static_cast<long>(Wo) * Ho));
        const __nv_bfloat16* p = x + ((static_cast<long>(n) * H + 2 * oy) * W + 2 * ox) * x_cs + x_coff + c;
        V8 a = ld8(p, dt), b = ld8(p + x_cs, dt), d = ld8(p + static_cast<long>(W) * x_cs, dt),
           e = ld8(p + static_cast<long>(W) * x_cs + x_cs, dt);
        V8 r;
#pragma unroll
        for (int k = 0; k < 8; ++k) r.f[k] = fmaxf(fmaxf(a.f[k], b.f[k]), fmaxf(d.f[k], e.f[k]));
        st8(y + pix * y_cs + y_coff + c, r, dt);
    }
}

// dx[n,h,w,c] = (add ? add[n,h,w,c] : 0) + (x[n,h,w,c] is the first max of its window ? dy : 0)
__global__ void maxpool2_bwd_kernel(const __nv_bfloat16* __restrict__ x, int x_cs, int x_coff,
                                    const __nv_bfloat16* __restrict__ dy, int dy_cs, int dy_coff,
                                    const __nv_bfloat16* __restrict__ add, int add_cs, int add_coff,
                                    __nv_bfloat16* __restrict__ dx, int dx_cs, int dx_coff, int N, int H, int W,
                                    int C8, int dt) {
    wcmc::pdl_start();
    const int Ho = H / 2, Wo = W / 2;
    const long total = static_cast<long>(N) * Ho * Wo * C8;
    WCMC_GRID_STRIDE(i, total) {
        int c = static_cast<int>(i % C8) * 8;
        long pix = i / C8;
        int ox = static_cast<int>(pix % Wo);
        int oy = static_cast<int>((pix / Wo) % Ho);
        int n = static_cast<int>(pix / (static_cast<long>(Wo) * Ho));
        const long base = (static_cast<long>(n) * H + 2 * oy) * W + 2 * ox;
        const long offs[4] = {base, base + 1, base + W, base + W + 1};
        V8 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = ld8(x + offs[q] * x_cs + x_coff + c, dt);
        V8 g = ld8(dy + pix * dy_cs + dy_coff + c, dt);
        int arg[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float m = v[0].f[k];
            int a = 0;
#pragma unroll
            for (int q = 1; q < 4; ++q)
                if (v[q].f[k] > m) { m = v[q].f[k]; a = q; }
            arg[k] = a;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            V8 o;
            if (add != nullptr) o = ld8(add + offs[q] * add_cs + add_coff + c, dt);
#pragma unroll
            for (int k = 0; k < 8; ++k) o.f[k] = (add != nullptr ? o.f[k] : 0.f) + (arg[k] == q ? g.f[k] : 0.f);
            st8(dx + offs[q] * dx_cs + dx_coff + c, o, dt);
        }
    }
}

__device__ __forceinline__ void up2_src(int o, int n_in, int& i0, int& i1, float& f) {
    float s = fmaxf(0.5f * (o + 0.5f) - 0.5f, 0.f);  // align_corners=False, scale 1/2
    i0 = static_cast<int>(s);
    f = s - i0;
    i1 = min(i0 + 1, n_in - 1);
}

// y (N,2h,2w) slice = bilinear 2x of x (N,h,w)
__global__ void upsample2_fwd_kernel(const __nv_bfloat16* __restrict__ x, int x_cs, int x_coff,
                                     __nv_bfloat16* __restrict__ y, int y_cs, int y_coff, int N, int h, int w,
                                     int C8, int dt) {
    wcmc::pdl_start();
    const int H = 2 * h, W = 2 * w;
    const long total = static_cast<long>(N) * H * W * C8;
    WCMC_GRID_STRIDE(i, total) {
        int c = static_cast<int>(i % C8) * 8;
        long pix = i / C8;
        int ox = static_cast<int>(pix % W);
        int oy = static_cast<int>((pix / W) % H);
        int n = static_cast<int>(pix / (static_cast<long>(W) * H));
        int y0, y1, x0, x1;
        float fy, fx;
        up2_src(oy, h, y0, y1, fy);
        up2_src(ox, w, x0, x1, fx);
        const __nv_bfloat16* b = x + static_cast<long>(n) * h * w * x_cs + x_coff + c;
        V8 a00 = ld8(b + (static_cast<long>(y0) * w + x0) * x_cs, dt), a01 = ld8(b + (static_cast<long>(y0) * w + x1) * x_cs, dt),
           a10 = ld8(b + (static_cast<long>(y1) * w + x0) * x_cs, dt), a11 = ld8(b + (static_cast<long>(y1) * w + x1) * x_cs, dt);
        V8 r;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            r.f[k] = (1.f - fy) * ((1.f - fx) * a00.f[k] + fx * a01.f[k]) + fy * ((1.f - fx) * a10.f[k] + fx * a11.f[k]);
        st8(y + pix * y_cs + y_coff + c, r, dt);
    }
}

// dx (N,h,w) = adjoint of the 2x bilinear up-sampling applied to dy (N,2h,2w) slice
__global__ void upsample2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int dy_cs, int dy_coff,
                                     __nv_bfloat16* __restrict__ dx, int dx_cs, int dx_coff, int N, int h, int w,
                                     int C8, int dt) {
    wcmc::pdl_start();
    const int H = 2 * h, W = 2 * w;
    const long total = static_cast<long>(N) * h * w * C8;
    WCMC_GRID_STRIDE(i, total) {
        int c = static_cast<int>(i % C8) * 8;
        long pix = i / C8;
        int ix = static_cast<int>(pix % w);
        int iy = static_cast<int>((pix / w) % h);
        int n = static_cast<int>(pix / (static_cast<long>(w) * h));
        V8 acc;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc.f[k] = 0.f;
        for (int oy = max(0, 2 * iy - 2); oy <= min(H - 1, 2 * iy + 2); ++oy) {
            int y0, y1;
            float fy;
            up2_src(oy, h, y0, y1, fy);
            float wy = (y0 == iy ? 1.f - fy : 0.f) + (y1 == iy ? fy : 0.f);
            if (wy == 0.f) continue;
            for (int ox = max(0, 2 * ix - 2); ox <= min(W - 1, 2 * ix + 2); ++ox) {
                int x0, x1;
                float fx;
                up2_src(ox, w, x0, x1, fx);
                float wx = (x0 == ix ? 1.f - fx : 0.f) + (x1 == ix ? fx : 0.f);
                if (wx == 0.f) continue;
                V8 g = ld8(dy + ((static_cast<long>(n) * H + oy) * W + ox) * dy_cs + dy_coff + c, dt);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc.f[k] = fmaf(wy * wx, g.f[k], acc.f[k]);
            }
        }
        st8(dx + pix * dx_cs + dx_coff + c, acc, dt);
    }
}

// y[b,pix,c] = scale * sum_s x[b,s,pix,c]          (scale = 1/S: mean over samples, networks.py:36)
__global__ void spp_reduce_kernel(const __nv_bfloat16* __restrict__ x, int x_cs, int x_coff,
                                  __nv_bfloat16* __restrict__ y, int y_cs, int y_coff, int B, int S, long HW,
                                  int C8, float scale, int dt) {
    wcmc::pdl_start();
    const long total = static_cast<long>(B) * HW * C8;
    WCMC_GRID_STRIDE(i, total) {
        int c = static_cast<int>(i % C8) * 8;
        long pix = i / C8;
        long hw = pix % HW;
        long b = pix / HW;
        V8 acc;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc.f[k] = 0.f;
        for (int s = 0; s < S; ++s) {
            V8 v = ld8(x + ((b * S + s) * HW + hw) * x_cs + x_coff + c, dt);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc.f[k] += v.f[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc.f[k] *= scale;
        st8(y + pix * y_cs + y_coff + c, acc, dt);
    }
}

// y[b,s,pix,c] = (add ? add[b,s,pix,c] : 0) + scale * x[b,pix,c]   (broadcast over samples)
__global__ void spp_broadcast_kernel(const __nv_bfloat16* __restrict__ x, int x_cs, int x_coff,
                                     const __nv_bfloat16* __restrict__ add, int add_cs, int add_coff,
                                     __nv_bfloat16* __restrict__ y, int y_cs, int y_coff, int B, int S, long HW,
                                     int C8, float scale, int dt) {
    wcmc::pdl_start();
    const long total = static_cast<long>(B) * S * HW * C8;
    WCMC_GRID_STRIDE(i, total) {
        int c = static_cast<int>(i % C8) * 8;
        long pix = i / C8;  // over (b,s,hw)
        long hw = pix % HW;
        long b = pix / (HW * S);
        V8 v = ld8(x + (b * HW + hw) * x_cs + x_coff + c, dt);
        V8 o;
        if (add != nullptr) o = ld8(add + pix * add_cs + add_coff + c, dt);
#pragma unroll
        for (int k = 0; k < 8; ++k) o.f[k] = (add != nullptr ? o.f[k] : 0.f) + scale * v.f[k];
        st8(y + pix * y_cs + y_coff + c, o, dt);
    }
}

// dz = dy * act'(y)    (act: 1 relu, 2 leaky with `slope`), y = the activation OUTPUT
__global__ void act_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int dy_cs, int dy_coff,
                               const __nv_bfloat16* __restrict__ y, int y_cs, int y_coff,
                               __nv_bfloat16* __restrict__ dz, int dz_cs, int dz_coff, long npix, int C8,
                               float slope, int dt) {
    wcmc::pdl_start();
    const long total = npix * C8;
    WCMC_GRID_STRIDE(i, total) {
        int c = static_cast<int>(i % C8) * 8;
        long pix = i / C8;
        V8 g = ld8(dy + pix * dy_cs + dy_coff + c, dt), v = ld8(y + pix * y_cs + y_coff + c, dt);
#pragma unroll
        for (int k = 0; k < 8; ++k) g.f[k] *= (v.f[k] > 0.f) ? 1.f : slope;
        st8(dz + pix * dz_cs + dz_coff + c, g, dt);
    }
}

}  // namespace wcmc

using namespace wcmc;

static int ew_blocks(long total) { return static_cast<int>(std::min<long>((total + 255) / 256, 148L * 16)); }

#define WCMC_EW_CHECK(name, C, ...)                                                                    \
    WCMC_REQUIRE((C) > 0 && (C) % 8 == 0, WCMC_ESHAPE, name ": channel count %d must be a multiple of 8", C); \
    {                                                                                                  \
        const int _v[] = {__VA_ARGS__};                                                                \
        for (int _x : _v)                                                                              \
            WCMC_REQUIRE(_x % 8 == 0, WCMC_EALIGN, name ": channel strides/offsets must be multiples of 8"); \
    }

extern "C" int wcmc_maxpool2_fwd(const void* x, int x_cs, int x_coff, void* y, int y_cs, int y_coff, int N, int H,
                                 int W, int C, int dtype, void* stream) {
    WCMC_EW_CHECK("maxpool2_fwd", C, x_cs, x_coff, y_cs, y_coff);
    WCMC_REQUIRE(N > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, WCMC_ESHAPE, "maxpool2: H, W must be even");
    long total = static_cast<long>(N) * (H / 2) * (W / 2) * (C / 8);
    WCMC_LAUNCH(maxpool2_fwd_kernel, ew_blocks(total), 256, 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(x), x_cs, x_coff, static_cast<__nv_bfloat16*>(y), y_cs, y_coff, N, H, W,
        C / 8, dtype);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" int wcmc_maxpool2_bwd(const void* x, int x_cs, int x_coff, const void* dy, int dy_cs, int dy_coff,
                                 const void* add, int add_cs, int add_coff, void* dx, int dx_cs, int dx_coff, int N,
                                 int H, int W, int C, int dtype, void* stream) {
    WCMC_EW_CHECK("maxpool2_bwd", C, x_cs, x_coff, dy_cs, dy_coff, add_cs, add_coff, dx_cs, dx_coff);
    WCMC_REQUIRE(N > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, WCMC_ESHAPE, "maxpool2: H, W must be even");
    long total = static_cast<long>(N) * (H / 2) * (W / 2) * (C / 8);
    WCMC_LAUNCH(maxpool2_bwd_kernel, ew_blocks(total), 256, 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(x), x_cs, x_coff, static_cast<const __nv_bfloat16*>(dy), dy_cs, dy_coff,
        static_cast<const __nv_bfloat16*>(add), add_cs, add_coff, static_cast<__nv_bfloat16*>(dx), dx_cs, dx_coff,
        N, H, W, C / 8, dtype);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" int wcmc_upsample2_fwd(const void* x, int x_cs, int x_coff, void* y, int y_cs, int y_coff, int N, int h,
                                  int w, int C, int dtype, void* stream) {
    WCMC_EW_CHECK("upsample2_fwd", C, x_cs, x_coff, y_cs, y_coff);
    WCMC_REQUIRE(N > 0 && h > 0 && w > 0, WCMC_ESHAPE, "upsample2: bad shape");
    long total = static_cast<long>(N) * h * w * 4 * (C / 8);
    WCMC_LAUNCH(upsample2_fwd_kernel, ew_blocks(total), 256, 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(x), x_cs, x_coff, static_cast<__nv_bfloat16*>(y), y_cs, y_coff, N, h, w,
        C / 8, dtype);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" int wcmc_upsample2_bwd(const void* dy, int dy_cs, int dy_coff, void* dx, int dx_cs, int dx_coff, int N,
                                  int h, int w, int C, int dtype, void* stream) {
    WCMC_EW_CHECK("upsample2_bwd", C, dy_cs, dy_coff, dx_cs, dx_coff);
    WCMC_REQUIRE(N > 0 && h > 0 && w > 0, WCMC_ESHAPE, "upsample2: bad shape");
    long total = static_cast<long>(N) * h * w * (C / 8);
    WCMC_LAUNCH(upsample2_bwd_kernel, ew_blocks(total), 256, 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(dy), dy_cs, dy_coff, static_cast<__nv_bfloat16*>(dx), dx_cs, dx_coff, N, h,
        w, C / 8, dtype);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" int wcmc_spp_reduce(const void* x, int x_cs, int x_coff, void* y, int y_cs, int y_coff, int B, int S,
                               int HW, int C, float scale, int dtype, void* stream) {
    WCMC_EW_CHECK("spp_reduce", C, x_cs, x_coff, y_cs, y_coff);
    WCMC_REQUIRE(B > 0 && S > 0 && HW > 0, WCMC_ESHAPE, "spp_reduce: bad shape");
    long total = static_cast<long>(B) * HW * (C / 8);
    WCMC_LAUNCH(spp_reduce_kernel, ew_blocks(total), 256, 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(x), x_cs, x_coff, static_cast<__nv_bfloat16*>(y), y_cs, y_coff, B, S, HW,
        C / 8, scale, dtype);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" int wcmc_spp_broadcast(const void* x, int x_cs, int x_coff, const void* add, int add_cs, int add_coff,
                                  void* y, int y_cs, int y_coff, int B, int S, int HW, int C, float scale,
                                  int dtype, void* stream) {
    WCMC_EW_CHECK("spp_broadcast", C, x_cs, x_coff, add_cs, add_coff, y_cs, y_coff);
    WCMC_REQUIRE(B > 0 && S > 0 && HW > 0, WCMC_ESHAPE, "spp_broadcast: bad shape");
    long total = static_cast<long>(B) * S * HW * (C / 8);
    WCMC_LAUNCH(spp_broadcast_kernel, ew_blocks(total), 256, 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(x), x_cs, x_coff, static_cast<const __nv_bfloat16*>(add), add_cs, add_coff,
        static_cast<__nv_bfloat16*>(y), y_cs, y_coff, B, S, HW, C / 8, scale, dtype);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" int wcmc_act_bwd(const void* dy, int dy_cs, int dy_coff, const void* y, int y_cs, int y_coff, void* dz,
                            int dz_cs, int dz_coff, long npix, int C, int act, float slope, int dtype, void* stream) {
    WCMC_EW_CHECK("act_bwd", C, dy_cs, dy_coff, y_cs, y_coff, dz_cs, dz_coff);
    WCMC_REQUIRE(npix > 0 && (act == WCMC_ACT_RELU || act == WCMC_ACT_LEAKY), WCMC_ESHAPE, "act_bwd: bad args");
    long total = npix * (C / 8);
    WCMC_LAUNCH(act_bwd_kernel, ew_blocks(total), 256, 0, static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(dy), dy_cs, dy_coff, static_cast<const __nv_bfloat16*>(y), y_cs, y_coff,
        static_cast<__nv_bfloat16*>(dz), dz_cs, dz_coff, npix, C / 8, act == WCMC_ACT_RELU ? 0.f : slope, dtype);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}
