// K9 / K12: the element-wise and reduction glue of the KPCN+WCMC step that round 1 left to ~50 eager ATen launches.
//
// K9  p-buffer statistics + concatenation (/root/reference/support/interfaces.py:165-180):
//       kpcn_in' = cat[kpcn_in, mean_S(p), var_S(p).mean(C) / S]      (var unbiased, detached; mean carries gradient)
//     one launch forward (the reference: var, mean, mean, div, two cats), one launch backward (d p = d mean / S).
// K12 radiance recombination of sbmc.KPCN.forward (radiance = albedo * diffuse + exp(specular) - 1) and
//     the image losses of the step in one reduction: L1 on the diffuse / specular / recombined images
//     (nn.L1Loss, /root/reference/train_kpcn.py:300-302) and RelativeMSE (support/losses.py:255-264:
//     0.5 * mean((x - y)^2 / (y^2 + eps))).  Targets are read as centred crops of the full-size tensors
//     (crop_like, support/utils.py:24-42): no sliced copies.  The launch also leaves sign(pred - target) / N for the
//     two L1 branches, so their backward is one multiply.  Sums are deterministic: fixed per-CTA partials, summed in
//     order by the last CTA to finish.
// All HBM-bound streaming kernels (fp32, 4-byte coalesced accesses along x).
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kGlueThreads = 256;

// grid (ceil(HW/256), B)
__global__ void __launch_bounds__(kGlueThreads)
pbuffer_concat_fwd_kernel(const float* __restrict__ kin, const float* __restrict__ p, float* __restrict__ out, int S,
                          int C, int c0, int cr, int Cin, int HW) {
    wcmc::pdl_start();
    const int b = blockIdx.y;
    const int i = blockIdx.x * kGlueThreads + threadIdx.x;
    if (i >= HW) return;
    const int Cout = Cin + cr + 1;
    const float* ki = kin + static_cast<long>(b) * Cin * HW + i;
    float* o = out + static_cast<long>(b) * Cout * HW + i;
    for (int c = 0; c < Cin; ++c) o[static_cast<long>(c) * HW] = __ldg(ki + static_cast<long>(c) * HW);
    const float* pb = p + static_cast<long>(b) * S * C * HW + i;
    float vsum = 0.f;
    for (int j = 0; j < cr; ++j) {
        const float* pc = pb + static_cast<long>(c0 + j) * HW;
        float m = 0.f;
        for (int s = 0; s < S; ++s) m += __ldg(pc + static_cast<long>(s) * C * HW);
        m /= static_cast<float>(S);
        float v = 0.f;
        for (int s = 0; s < S; ++s) {
            const float d = __ldg(pc + static_cast<long>(s) * C * HW) - m;
            v = fmaf(d, d, v);
        }
        o[static_cast<long>(Cin + j) * HW] = m;
        vsum += v / static_cast<float>(S - 1);          // unbiased, like torch.var
    }
    o[static_cast<long>(Cin + cr) * HW] = vsum / static_cast<float>(cr) / static_cast<float>(S);
}

// dp[b,s,c,:] = (c0 <= c < c0+cr) ? g[b, Cin + c - c0, :] / S : 0          grid (ceil(HW/256), B*S)
__global__ void __launch_bounds__(kGlueThreads)
pbuffer_concat_bwd_kernel(const float* __restrict__ g, float* __restrict__ dp, int S, int C, int c0, int cr, int Cin,
                          int HW) {
    wcmc::pdl_start();
    const int bs = blockIdx.y, b = bs / S;
    const int i = blockIdx.x * kGlueThreads + threadIdx.x;
    if (i >= HW) return;
    const int Cout = Cin + cr + 1;
    const float inv = 1.f / static_cast<float>(S);
    float* d = dp + static_cast<long>(bs) * C * HW + i;
    const float* gb = g + (static_cast<long>(b) * Cout + Cin) * HW + i;
    for (int c = 0; c < C; ++c) {
        const int j = c - c0;
        d[static_cast<long>(c) * HW] = (j >= 0 && j < cr) ? __ldg(gb + static_cast<long>(j) * HW) * inv : 0.f;
    }
}

struct CropView {       // (B, 3, H, W) fp32 with unit x stride, read at [y0 + y][x0 + x]
    const float* p;
    long sb, sc, sh;
};

// radiance = albedo * r_d + exp(r_s) - 1      (r_d, r_s, radiance contiguous (B,3,h,w); albedo a crop view)
__global__ void __launch_bounds__(kGlueThreads)
recombine_kernel(CropView alb, const float* __restrict__ rd, const float* __restrict__ rs, float* __restrict__ rad,
                 int h, int w, long n) {
    wcmc::pdl_start();
    const long i = blockIdx.x * static_cast<long>(kGlueThreads) + threadIdx.x;
    if (i >= n) return;
    const int x = static_cast<int>(i % w);
    const int y = static_cast<int>((i / w) % h);
    const int c = static_cast<int>((i / (static_cast<long>(w) * h)) % 3);
    const long b = i / (static_cast<long>(w) * h * 3);
    const float a = __ldg(alb.p + b * alb.sb + c * alb.sc + y * alb.sh + x);
    rad[i] = a * rd[i] + expf(rs[i]) - 1.0f;
}

constexpr int kLossSums = 4;

// sums[0..3] = mean|rd - td|, mean|rs - ts|, mean|rad - tt|, 0.5 mean((rad - tt)^2 / (tt^2 + eps)); any prediction may be
// null (its sum is 0).  sgn_d / sgn_s (optional) = sign(pred - target) / n.
__global__ void __launch_bounds__(kGlueThreads)
image_losses_kernel(const float* __restrict__ rd, const float* __restrict__ rs, const float* __restrict__ rad,
                    CropView td, CropView ts, CropView tt, float* __restrict__ sgn_d, float* __restrict__ sgn_s,
                    int h, int w, long n, float eps, float* __restrict__ partial, unsigned* __restrict__ ticket,
                    float* __restrict__ sums) {
    wcmc::pdl_start();
    float acc[kLossSums] = {0.f, 0.f, 0.f, 0.f};
    const float inv_n = 1.0f / static_cast<float>(n);
    for (long i = blockIdx.x * static_cast<long>(kGlueThreads) + threadIdx.x; i < n;
         i += static_cast<long>(gridDim.x) * kGlueThreads) {
        const int x = static_cast<int>(i % w);
        const int y = static_cast<int>((i / w) % h);
        const int c = static_cast<int>((i / (static_cast<long>(w) * h)) % 3);
        const long b = i / (static_cast<long>(w) * h * 3);
        if (rd != nullptr) {
            const float d = rd[i] - __ldg(td.p + b * td.sb + c * td.sc + y * td.sh + x);
            acc[0] += fabsf(d);
            if (sgn_d != nullptr) sgn_d[i] = d > 0.f ? inv_n : (d < 0.f ? -inv_n : 0.f);
        }
        if (rs != nullptr) {
            const float d = rs[i] - __ldg(ts.p + b * ts.sb + c * ts.sc + y * ts.sh + x);
            acc[1] += fabsf(d);
            if (sgn_s != nullptr) sgn_s[i] = d > 0.f ? inv_n : (d < 0.f ? -inv_n : 0.f);
        }
        if (rad != nullptr) {
            const float t = __ldg(tt.p + b * tt.sb + c * tt.sc + y * tt.sh + x);
            const float d = rad[i] - t;
            acc[2] += fabsf(d);
            acc[3] += d * d / (t * t + eps);
        }
    }
    __shared__ float red[kLossSums][kGlueThreads / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < kLossSums; ++k) {
        const float v = wcmc::warp_sum(acc[k]);
        if (lane == 0) red[k][warp] = v;
    }
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x < kLossSums) {
        float v = 0.f;
        for (int j = 0; j < kGlueThreads / 32; ++j) v += red[threadIdx.x][j];
        partial[blockIdx.x * kLossSums + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    // the last CTA to finish sums the per-CTA partials: warp k takes sum k, lane-strided in a fixed order, then the
    // fixed shuffle tree -- the result does not depend on which CTA came last
    if (warp < kLossSums) {
        float v = 0.f;
        for (unsigned j = lane; j < gridDim.x; j += 32) v += __ldcg(partial + j * kLossSums + warp);
        v = wcmc::warp_sum(v);
        if (lane == 0) sums[warp] = v * inv_n * (warp == 3 ? 0.5f : 1.0f);
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

// out[0] = target / max|g|, out[1] = max|g| / target   (the per-pass loss scale of the 16-bit backward, ops.grad_scale;
// round 1: abs + amax + clamp + reciprocal + two multiplies = six ATen launches and a temporary the size of g)
__global__ void __launch_bounds__(kGlueThreads)
absmax_scale_kernel(const float* __restrict__ g, long n, float target, float* __restrict__ partial,
                    unsigned* __restrict__ ticket, float* __restrict__ out) {
    wcmc::pdl_start();
    float m = 0.f;
    const long n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (long i = blockIdx.x * static_cast<long>(kGlueThreads) + threadIdx.x; i < n4;
         i += static_cast<long>(gridDim.x) * kGlueThreads) {
        const float4 v = __ldg(g4 + i);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) m = fmaxf(m, fabsf(__ldg(g + (n4 << 2) + threadIdx.x)));
    __shared__ float red[kGlueThreads / 32];
    __shared__ bool last;
    m = wcmc::warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int j = 1; j < kGlueThreads / 32; ++j) m = fmaxf(m, red[j]);
        partial[blockIdx.x] = m;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x < 32) {
        float v = 0.f;
        for (unsigned j = threadIdx.x; j < gridDim.x; j += 32) v = fmaxf(v, __ldcg(partial + j));
        v = wcmc::warp_max(v);
        if (threadIdx.x == 0) {
            v = fmaxf(v, 1e-30f);
            out[0] = target / v;
            out[1] = v / target;
            *ticket = 0u;
        }
    }
}

}  // namespace

extern "C" size_t wcmc_absmax_scale_workspace(void) { return (2 * 148 + 4) * sizeof(float); }

extern "C" int wcmc_absmax_scale(const float* g, long n, float target, float* out2, void* workspace,
                                 size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(g && out2 && n > 0 && target > 0.f, WCMC_ESHAPE, "absmax_scale: bad arguments");
    WCMC_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, WCMC_EALIGN, "absmax_scale: g must be 16-byte aligned");
    WCMC_REQUIRE(workspace && workspace_bytes >= wcmc_absmax_scale_workspace(), WCMC_EWORKSPACE,
                 "absmax_scale: workspace too small");
    float* partial = static_cast<float*>(workspace);
    unsigned* ticket = reinterpret_cast<unsigned*>(partial + 2 * 148);       // caller zero-fills the workspace ONCE
    const int grid = static_cast<int>(std::min<long>((n / 4 + kGlueThreads - 1) / kGlueThreads + 1, 2L * 148));
    WCMC_LAUNCH(absmax_scale_kernel, grid, kGlueThreads, 0, stream, g, n, target, partial, ticket, out2);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" int wcmc_pbuffer_concat_fwd(const float* kpcn_in, const float* p, float* out, int B, int S, int C, int c0,
                                       int cr, int Cin, int HW, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(kpcn_in && p && out && B > 0 && S > 0 && C > 0 && Cin > 0 && HW > 0 && B * S <= 65535, WCMC_ESHAPE,
                 "pbuffer_concat_fwd: bad shape");
    WCMC_REQUIRE(c0 >= 0 && cr > 0 && c0 + cr <= C, WCMC_ESHAPE, "pbuffer_concat_fwd: channel range [%d, %d) of %d", c0,
                 c0 + cr, C);
    dim3 grid((HW + kGlueThreads - 1) / kGlueThreads, B);
    WCMC_LAUNCH(pbuffer_concat_fwd_kernel, grid, kGlueThreads, 0, stream, kpcn_in, p, out, S, C, c0, cr, Cin, HW);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" int wcmc_pbuffer_concat_bwd(const float* grad_out, float* dp, int B, int S, int C, int c0, int cr, int Cin,
                                       int HW, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(grad_out && dp && B > 0 && S > 0 && C > 0 && Cin > 0 && HW > 0 && B * S <= 65535, WCMC_ESHAPE,
                 "pbuffer_concat_bwd: bad shape");
    WCMC_REQUIRE(c0 >= 0 && cr > 0 && c0 + cr <= C, WCMC_ESHAPE, "pbuffer_concat_bwd: bad channel range");
    dim3 grid((HW + kGlueThreads - 1) / kGlueThreads, B * S);
    WCMC_LAUNCH(pbuffer_concat_bwd_kernel, grid, kGlueThreads, 0, stream, grad_out, dp, S, C, c0, cr, Cin, HW);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" int wcmc_recombine(const float* albedo, long a_sb, long a_sc, long a_sh, const float* r_d, const float* r_s,
                              float* radiance, int B, int h, int w, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(albedo && r_d && r_s && radiance && B > 0 && h > 0 && w > 0, WCMC_ESHAPE, "recombine: bad arguments");
    const long n = static_cast<long>(B) * 3 * h * w;
    CropView a{albedo, a_sb, a_sc, a_sh};
    WCMC_LAUNCH(recombine_kernel, static_cast<unsigned>((n + kGlueThreads - 1) / kGlueThreads), kGlueThreads, 0, stream, a, r_d, r_s, radiance, h, w, n);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" size_t wcmc_image_losses_workspace(void) { return (4 * 148 * kLossSums + 4) * sizeof(float); }

extern "C" int wcmc_image_losses(const float* r_d, const float* r_s, const float* radiance, const float* t_d,
                                 const float* t_s, const float* t_t, const long* strides9, int B, int h, int w,
                                 float eps, float* sgn_d, float* sgn_s, float* sums4, void* workspace,
                                 size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(sums4 && strides9 && B > 0 && h > 0 && w > 0, WCMC_ESHAPE, "image_losses: bad arguments");
    WCMC_REQUIRE((!r_d || t_d) && (!r_s || t_s) && (!radiance || t_t), WCMC_ESHAPE, "image_losses: prediction without target");
    WCMC_REQUIRE(workspace && workspace_bytes >= wcmc_image_losses_workspace(), WCMC_EWORKSPACE,
                 "image_losses: workspace too small");
    const long n = static_cast<long>(B) * 3 * h * w;
    const int grid = static_cast<int>(std::min<long>((n + 4 * kGlueThreads - 1) / (4 * kGlueThreads), 2L * 148));
    CropView td{t_d, strides9[0], strides9[1], strides9[2]}, ts{t_s, strides9[3], strides9[4], strides9[5]},
        tt{t_t, strides9[6], strides9[7], strides9[8]};
    float* partial = static_cast<float*>(workspace);
    unsigned* ticket = reinterpret_cast<unsigned*>(partial + 4 * 148 * kLossSums);   // caller zero-initialises it ONCE
    WCMC_LAUNCH(image_losses_kernel, grid, kGlueThreads, 0, stream, r_d, r_s, radiance, td, ts, tt, sgn_d, sgn_s, h, w, n, eps,
                                                           partial, ticket, sums4);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}
