// raw OptaGen sample buffer (H,W,S,104) fp32 -> the KPCN feature buffer (H,W,44) and the path
// descriptor buffer (H,W,S,37), i.e. DenoiseDataset._preprocess_kpcn / _preprocess_llpm of
// /root/reference/support/datasets.py:487-582 / :301-361 (numpy on one CPU core there), with the NaN / inf clamp of
// :621-624 folded into the loads.  HBM-bound: 416 B per sample read, 176 B per pixel + 148 B per sample written.
//
//   llpm   : purely elementwise -- one thread per output element, consecutive threads = consecutive output channels of
//            one sample (reads inside one 176-byte run of the raw row, fully coalesced writes).
//   kpcn   : pass 1, one thread per pixel: mean / variance over the S samples of the 13 raw channels it needs (two
//            28-byte runs per sample: every 32-byte sector is fetched once and reused through L1 by the next loads of
//            the same sample), the albedo factorisation / log transform, 18 per-pixel values to a workspace and an
//            atomic max of the mean depth;  pass 2, one thread per (pixel, output channel): depth normalisation by
//            the image maximum, left / top finite differences, the 44-channel layout.
// Status: written in round 1 after the GPU budget was spent -- compiled, NOT yet run on a GPU; nothing on the product
// path calls it and its GPU tests are gated (tests/test_gpu_preprocess.py).
#include <algorithm>

#include "common.cuh"

namespace wcmc {

constexpr int kRawC = 104;
constexpr float kClamp = 1.0e+38f;

__device__ __forceinline__ float sane(float v) {   // np.where(isfinite(v), v, 1e38); np.where(v < 1e38, v, 1e38)
    return (isfinite(v) && v < kClamp) ? v : kClamp;
}

// out (npix*S, 37): [log(pw+1e-6)/90 | log(rad+1e-6)/30 x3 | log(light+1e-8)/10 x3 | log(thr+1e-6)/30 x18 |
//                    bounce/19 x6 | sqrt(rough) x6]            raw channels: 73 | 74..76 | 77..79 | 80..97 | 60..65 | 98..103
__global__ void __launch_bounds__(256) preprocess_llpm_kernel(const float* __restrict__ raw, long nrows,
                                                             float* __restrict__ out) {
    const long total = nrows * 37;
    for (long i = blockIdx.x * 256L + threadIdx.x; i < total; i += gridDim.x * 256L) {
        const long row = i / 37;
        const int c = static_cast<int>(i - row * 37);
        const float* r = raw + row * kRawC;
        float v;
        if (c == 0) v = logf(sane(__ldg(r + 73)) + 1e-6f) / 90.0f;
        else if (c < 4) v = logf(sane(__ldg(r + 73 + c)) + 1e-6f) / 30.0f;
        else if (c < 7) v = logf(sane(__ldg(r + 73 + c)) + 1e-8f) / 10.0f;
        else if (c < 25) v = logf(sane(__ldg(r + 73 + c)) + 1e-6f) / 30.0f;
        else if (c < 31) v = sane(__ldg(r + 60 + (c - 25))) / 19.0f;
        else v = sqrtf(sane(__ldg(r + 98 + (c - 31))));
        out[i] = v;
    }
}

// workspace per pixel: 0..2 diffuse | 3 diffuse_v | 4..6 specular | 7 specular_v | 8..10 normal | 11 normal_v |
//                      12 depth (mean, un-normalised) | 13 depth_v (un-normalised) | 14..16 albedo | 17 albedo_v
constexpr int kStats = 18;

template <int C>
__device__ __forceinline__ void mean_var(const float (&x)[8][C], int S, float (&mean)[C], float (&var)[C]) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
        float m = 0.f;
        for (int s = 0; s < S; ++s) m += x[s][c];
        m /= S;
        float v = 0.f;
        for (int s = 0; s < S; ++s) {
            const float d = x[s][c] - m;
            v += d * d;
        }
        mean[c] = m;
        var[c] = v / S;
    }
}

__global__ void __launch_bounds__(128) preprocess_kpcn_stats_kernel(const float* __restrict__ raw, long npix, int S,
                                                                   float* __restrict__ ws, int* __restrict__ max_depth_bits) {
    const float eps = 0.00316f;
    float local_max = 0.f;
    for (long p = blockIdx.x * 128L + threadIdx.x; p < npix; p += gridDim.x * 128L) {
        const float* r = raw + p * S * kRawC;
        float rad[8][3], dif[8][3], alb[8][3], nrm[8][3], dep[8][1];
        for (int s = 0; s < S; ++s) {
            const float* q = r + s * kRawC;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float d = fmaxf(sane(__ldg(q + 5 + c)), 0.f);                  // np.maximum(diffuse, 0)
                dif[s][c] = d;
                rad[s][c] = fmaxf(fmaxf(sane(__ldg(q + 2 + c)), 0.f) - d, 0.f);      // specular sample (:541-543)
                alb[s][c] = sane(__ldg(q + 66 + c));
                nrm[s][c] = sane(__ldg(q + 69 + c));
            }
            dep[s][0] = sane(__ldg(q + 72));
        }
        float m3[3], v3[3], m1[1], v1[1];
        float* o = ws + p * kStats;
        // albedo first: the diffuse factorisation needs it
        mean_var<3>(alb, S, m3, v3);
        const float a0 = m3[0], a1 = m3[1], a2 = m3[2];
        o[14] = a0; o[15] = a1; o[16] = a2;
        o[17] = ((v3[0] + v3[1] + v3[2]) / 3.f) / S;
        const float albedo_sqr = ((a0 + eps) * (a0 + eps) + (a1 + eps) * (a1 + eps) + (a2 + eps) * (a2 + eps)) / 3.f;
        mean_var<3>(dif, S, m3, v3);
        o[0] = m3[0] / (a0 + eps); o[1] = m3[1] / (a1 + eps); o[2] = m3[2] / (a2 + eps);
        o[3] = (((v3[0] + v3[1] + v3[2]) / 3.f) / S) / albedo_sqr;
        mean_var<3>(rad, S, m3, v3);
        const float spec_sqr = ((1.f + m3[0]) * (1.f + m3[0]) + (1.f + m3[1]) * (1.f + m3[1]) + (1.f + m3[2]) * (1.f + m3[2])) / 3.f;
        o[4] = logf(1.f + m3[0]); o[5] = logf(1.f + m3[1]); o[6] = logf(1.f + m3[2]);
        o[7] = (((v3[0] + v3[1] + v3[2]) / 3.f) / S) / spec_sqr;
        mean_var<3>(nrm, S, m3, v3);
        o[8] = m3[0]; o[9] = m3[1]; o[10] = m3[2];
        o[11] = ((v3[0] + v3[1] + v3[2]) / 3.f) / S;
        mean_var<1>(dep, S, m1, v1);
        o[12] = m1[0];
        o[13] = v1[0];
        local_max = fmaxf(local_max, m1[0]);
    }
    // depth.max(): non-negative floats order like their bit patterns; a negative maximum never beats the initial 0.0,
    // which reproduces `if max_depth > 0` (:519-521)
    local_max = warp_max(local_max);
    if ((threadIdx.x & 31) == 0 && local_max > 0.f) atomicMax(max_depth_bits, __float_as_int(local_max));
}

// out (npix, 44): diffuse 3 | v | dx 3 | dy 3 | specular 3 | v | dx 3 | dy 3 | normal 3 | v | dx 3 | dy 3 | depth | v | dx | dy |
//                 albedo 3 | v | dx 3 | dy 3
__global__ void __launch_bounds__(256) preprocess_kpcn_finish_kernel(const float* __restrict__ ws, int H, int W, int S,
                                                                    const int* __restrict__ max_depth_bits,
                                                                    float* __restrict__ out) {
    const float md = __int_as_float(__ldg(max_depth_bits));
    const long total = static_cast<long>(H) * W * 44;
    for (long i = blockIdx.x * 256L + threadIdx.x; i < total; i += gridDim.x * 256L) {
        const long p = i / 44;
        const int c = static_cast<int>(i - p * 44);
        const int y = static_cast<int>(p / W), x = static_cast<int>(p - static_cast<long>(y) * W);
        // group layout: (base output channel, base stats channel, number of value channels)
        int g0, s0, nc;
        if (c < 10) { g0 = 0; s0 = 0; nc = 3; }
        else if (c < 20) { g0 = 10; s0 = 4; nc = 3; }
        else if (c < 30) { g0 = 20; s0 = 8; nc = 3; }
        else if (c < 34) { g0 = 30; s0 = 12; nc = 1; }
        else { g0 = 34; s0 = 14; nc = 3; }
        const int k = c - g0;
        const bool is_depth = (s0 == 12);
        auto value = [&](long pix, int ch) {   // final value channel `ch` of the group at pixel `pix`
            float v = ws[pix * kStats + s0 + ch];
            if (is_depth) {
                if (md > 0.f) v = v / md;
                v = fminf(fmaxf(v, 0.f), 1.f);
            }
            return v;
        };
        float v;
        if (k < nc) {
            v = value(p, k);
        } else if (k == nc) {                    // variance channel
            v = ws[p * kStats + s0 + nc];
            if (is_depth && md > 0.f) v = v / (md * md * S);
        } else if (k < 2 * nc + 1) {             // dx: zero in the first column
            const int ch = k - nc - 1;
            v = x > 0 ? value(p, ch) - value(p - 1, ch) : 0.f;
        } else {                                 // dy: zero in the first row
            const int ch = k - 2 * nc - 1;
            v = y > 0 ? value(p, ch) - value(p - W, ch) : 0.f;
        }
        out[i] = v;
    }
}

}  // namespace wcmc

using namespace wcmc;

extern "C" size_t wcmc_preprocess_kpcn_workspace(int H, int W) {
    return (static_cast<size_t>(H) * W * kStats + 4) * sizeof(float);
}

extern "C" int wcmc_preprocess_kpcn(const float* raw, int H, int W, int S, float* out44, void* workspace,
                                    size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(raw && out44 && workspace && H > 0 && W > 0 && S >= 1 && S <= 8, WCMC_ESHAPE,
                 "preprocess_kpcn: need H, W > 0 and 1 <= spp <= 8 (got %d x %d, %d spp)", H, W, S);
    WCMC_REQUIRE(workspace_bytes >= wcmc_preprocess_kpcn_workspace(H, W), WCMC_EWORKSPACE,
                 "preprocess_kpcn: workspace too small");
    const long npix = static_cast<long>(H) * W;
    float* ws = static_cast<float*>(workspace);
    int* md = reinterpret_cast<int*>(ws + npix * kStats);
    WCMC_CHECK_CUDA(cudaMemsetAsync(md, 0, sizeof(int), stream));
    const int sms = wcmc_num_sms();
    const int b1 = static_cast<int>(std::min<long>((npix + 127) / 128, 8L * sms));
    preprocess_kpcn_stats_kernel<<<b1, 128, 0, stream>>>(raw, npix, S, ws, md);
    WCMC_LAUNCH_CHECK();
    const int b2 = static_cast<int>(std::min<long>((npix * 44 + 255) / 256, 16L * sms));
    preprocess_kpcn_finish_kernel<<<b2, 256, 0, stream>>>(ws, H, W, S, md, out44);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" int wcmc_preprocess_llpm(const float* raw, long nrows, float* out37, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(raw && out37 && nrows > 0, WCMC_ESHAPE, "preprocess_llpm: bad arguments");
    const int sms = wcmc_num_sms();
    const int blocks = static_cast<int>(std::min<long>((nrows * 37 + 255) / 256, 16L * sms));
    preprocess_llpm_kernel<<<blocks, 256, 0, stream>>>(raw, nrows, out37);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}
