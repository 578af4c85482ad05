// Per-element arithmetic of the N3 preprocessing kernels (preprocess.cu), written as __host__ __device__ functions so
// that tests/native/preprocess_host_check.cu can run EXACTLY this code on the CPU against the reference-generated
// vectors (the kernels themselves had not run on a GPU when round 1 ended).  Follows
// /root/reference/support/datasets.py:301-361 (_preprocess_llpm), :487-582 (_preprocess_kpcn), :286-299 (_gradients),
// :621-624 (NaN / inf clamp).  Raw channel indices: datasets.py:223-266 with MAX_DEPTH = 5.
#pragma once
#include <math.h>

#if defined(__CUDA_ARCH__)
#define WCMC_PREP_LD(p) __ldg(p)
#else
#define WCMC_PREP_LD(p) (*(p))
#endif
#if defined(__CUDACC__)
#define WCMC_HD __host__ __device__ __forceinline__
#else
#define WCMC_HD inline
#endif

namespace wcmc {
namespace prep {

constexpr int kRawC = 104;
constexpr int kStats = 18;
constexpr int kMaxSpp = 8;

WCMC_HD float sane(float v) {   // np.where(isfinite(v), v, 1e38); np.where(v < 1e38, v, 1e38)
    const float c = 1.0e+38f;
    return (isfinite(v) && v < c) ? v : c;
}

// out channel c of one sample row: [log(pw+1e-6)/90 | log(rad+1e-6)/30 x3 | log(light+1e-8)/10 x3 | log(thr+1e-6)/30 x18 |
// bounce/19 x6 | sqrt(rough) x6];  raw channels 73 | 74..76 | 77..79 | 80..97 | 60..65 | 98..103
WCMC_HD float llpm_value(const float* r, int c) {
    if (c == 0) return logf(sane(WCMC_PREP_LD(r + 73)) + 1e-6f) / 90.0f;
    if (c < 4) return logf(sane(WCMC_PREP_LD(r + 73 + c)) + 1e-6f) / 30.0f;
    if (c < 7) return logf(sane(WCMC_PREP_LD(r + 73 + c)) + 1e-8f) / 10.0f;
    if (c < 25) return logf(sane(WCMC_PREP_LD(r + 73 + c)) + 1e-6f) / 30.0f;
    if (c < 31) return sane(WCMC_PREP_LD(r + 60 + (c - 25))) / 19.0f;
    return sqrtf(sane(WCMC_PREP_LD(r + 98 + (c - 31))));
}

template <int C>
WCMC_HD void mean_var(const float (&x)[kMaxSpp][C], int S, float (&mean)[C], float (&var)[C]) {
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

// The 13 raw channels the KPCN statistics read, per sample: total radiance 2..4, diffuse 5..7, albedo 66..68,
// normal 69..71, depth 72 (datasets.py:223-266).  kRawIdx[j] = raw channel of value j.
constexpr int kNeed = 13;
WCMC_HD int raw_index(int j) { return j < 6 ? 2 + j : 60 + j; }    // 0..5 -> 2..7, 6..12 -> 66..72

// The 18 per-pixel values of one pixel from the (sanitised) 13 needed values of each of its S samples -> o; returns the
// mean depth (for the image maximum).  v[s] = {total r,g,b, diffuse r,g,b, albedo r,g,b, normal x,y,z, depth}.
//   o: 0..2 diffuse | 3 diffuse_v | 4..6 specular | 7 specular_v | 8..10 normal | 11 normal_v | 12 depth (mean, not yet
//      normalised) | 13 depth_v (not yet normalised) | 14..16 albedo | 17 albedo_v
WCMC_HD float kpcn_stats_from_values(const float (*v)[kNeed], int S, float* o) {
    const float eps = 0.00316f;
    float spc[kMaxSpp][3], dif[kMaxSpp][3], alb[kMaxSpp][3], nrm[kMaxSpp][3], dep[kMaxSpp][1];
    for (int s = 0; s < S; ++s) {
        for (int c = 0; c < 3; ++c) {
            const float d = fmaxf(v[s][3 + c], 0.f);                                         // np.maximum(diffuse, 0)
            dif[s][c] = d;
            spc[s][c] = fmaxf(fmaxf(v[s][c], 0.f) - d, 0.f);                                 // specular sample (:541-543)
            alb[s][c] = v[s][6 + c];
            nrm[s][c] = v[s][9 + c];
        }
        dep[s][0] = v[s][12];
    }
    float m3[3], v3[3], m1[1], v1[1];
    mean_var<3>(alb, S, m3, v3);   // albedo first: the diffuse factorisation needs it
    const float a0 = m3[0], a1 = m3[1], a2 = m3[2];
    o[14] = a0; o[15] = a1; o[16] = a2;
    o[17] = ((v3[0] + v3[1] + v3[2]) / 3.f) / S;
    const float albedo_sqr = ((a0 + eps) * (a0 + eps) + (a1 + eps) * (a1 + eps) + (a2 + eps) * (a2 + eps)) / 3.f;
    mean_var<3>(dif, S, m3, v3);
    o[0] = m3[0] / (a0 + eps); o[1] = m3[1] / (a1 + eps); o[2] = m3[2] / (a2 + eps);
    o[3] = (((v3[0] + v3[1] + v3[2]) / 3.f) / S) / albedo_sqr;
    mean_var<3>(spc, S, m3, v3);
    const float spec_sqr =
        ((1.f + m3[0]) * (1.f + m3[0]) + (1.f + m3[1]) * (1.f + m3[1]) + (1.f + m3[2]) * (1.f + m3[2])) / 3.f;
    o[4] = logf(1.f + m3[0]); o[5] = logf(1.f + m3[1]); o[6] = logf(1.f + m3[2]);
    o[7] = (((v3[0] + v3[1] + v3[2]) / 3.f) / S) / spec_sqr;
    mean_var<3>(nrm, S, m3, v3);
    o[8] = m3[0]; o[9] = m3[1]; o[10] = m3[2];
    o[11] = ((v3[0] + v3[1] + v3[2]) / 3.f) / S;
    mean_var<1>(dep, S, m1, v1);
    o[12] = m1[0];
    o[13] = v1[0];
    return m1[0];
}

// The same from the pixel's S x 104 raw floats (r).
WCMC_HD float kpcn_pixel_stats(const float* r, int S, float* o) {
    float v[kMaxSpp][kNeed];
    for (int s = 0; s < S; ++s)
        for (int j = 0; j < kNeed; ++j) v[s][j] = sane(WCMC_PREP_LD(r + s * kRawC + raw_index(j)));
    return kpcn_stats_from_values(v, S, o);
}

// Output channel c (0..43) of pixel p from the per-pixel workspace: depth normalisation by the image maximum `md`
// (only if md > 0), left / top zero-padded finite differences, the reference's 44-channel order
//   diffuse 3 | v | dx 3 | dy 3 | specular 3 | v | dx 3 | dy 3 | normal 3 | v | dx 3 | dy 3 | depth | v | dx | dy |
//   albedo 3 | v | dx 3 | dy 3
WCMC_HD float kpcn_finish_value(const float* ws, int W, int S, float md, long p, int c) {
    const int y = static_cast<int>(p / W), x = static_cast<int>(p - static_cast<long>(y) * W);
    int g0, s0, nc;   // group: base output channel, base workspace channel, number of value channels
    if (c < 10) { g0 = 0; s0 = 0; nc = 3; }
    else if (c < 20) { g0 = 10; s0 = 4; nc = 3; }
    else if (c < 30) { g0 = 20; s0 = 8; nc = 3; }
    else if (c < 34) { g0 = 30; s0 = 12; nc = 1; }
    else { g0 = 34; s0 = 14; nc = 3; }
    const int k = c - g0;
    const bool is_depth = (s0 == 12);
    auto value = [&](long pix, int ch) {
        float v = ws[pix * kStats + s0 + ch];
        if (is_depth) {
            if (md > 0.f) v = v / md;
            v = fminf(fmaxf(v, 0.f), 1.f);
        }
        return v;
    };
    if (k < nc) return value(p, k);
    if (k == nc) {   // variance channel
        float v = ws[p * kStats + s0 + nc];
        if (is_depth && md > 0.f) v = v / (md * md * S);
        return v;
    }
    if (k < 2 * nc + 1) {   // dx: zero in the first column
        const int ch = k - nc - 1;
        return x > 0 ? value(p, ch) - value(p - 1, ch) : 0.f;
    }
    const int ch = k - 2 * nc - 1;   // dy: zero in the first row
    return y > 0 ? value(p, ch) - value(p - W, ch) : 0.f;
}

}  // namespace prep
}  // namespace wcmc
