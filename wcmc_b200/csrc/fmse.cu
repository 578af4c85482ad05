// K10: path-disentangling loss, permutation-paired form
// (/root/reference/support/losses.py:33-61 `intra_patch_dist` / `intra_batch_dist`, :82-113 forward).
//
// Row i = sample-pixel (b, s, y, x) carries the embedding P_i (C floats, strided NCHW view of the
// p-buffer crop) and the tone-mapped weak label t_i = (max(ref,0)/(1+max(ref,0)))^0.454545 of its
// pixel (losses.py:63-65, :94-97).  For a permutation pi:
//     e_i = 1/2 |P_i - P_pi(i)|^2 - 1/2 |t_i - t_pi(i)|^2 ,   L = 1/2 mean_i e_i^2
// once with one permutation of the S*H*W rows shared by every batch item ("patch") and once with a
// permutation of all B*S*H*W rows ("batch").  HBM/L2-bound gather: one thread per row does both
// modes, so P_i / t_i are read once; partner rows are 4-byte gathers that hit L2 (the p-buffer is
// 12.6 MB at B=8,S=8,C=3).  The reference runs ~20 ATen kernels with permute->reshape copies.
//
// Backward (no grad to ref): dL/dP_i = w_i (P_i - P_pi(i)) + w_inv(i) (P_i - P_inv(i)) per mode,
// w = row weights (FeatureMSE: e * g / N).  inv = pi^-1 is produced by the forward launch, so the
// backward is a pure gather: deterministic, no atomics.
#include "common.cuh"

namespace {

struct PView {           // strided (B,S,C,H,W) fp32 view, innermost (x) stride 1
    const float* p;
    long sb, ss, sc, sh;
};
struct RView {           // strided (B,3,H,W) fp32 view
    const float* p;
    long sb, sc, sh;
};

__device__ __forceinline__ float tonemap(float v) {
    v = fmaxf(v, 0.0f);
    return powf(v / (1.0f + v), 0.454545f);
}

constexpr int kThreads = 256;

// grid: (ceil(n / 256), B); n = S*H*W rows per batch item
__global__ void __launch_bounds__(kThreads)
fmse_perm_fwd_kernel(PView pv, RView rv, const int64_t* __restrict__ idx_patch,
                     const int64_t* __restrict__ idx_batch, int B, int S, int C, int H, int W,
                     float* __restrict__ e_patch, float* __restrict__ e_batch,
                     int32_t* __restrict__ inv_patch, int32_t* __restrict__ inv_batch,
                     float* __restrict__ partial, int* __restrict__ nonfinite) {
    wcmc::pdl_start();
    const int hw = H * W;
    const int n = S * hw;
    const int b = blockIdx.y;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    float ep = 0.0f, eb = 0.0f;
    if (i < n) {
        const int s = i / hw, r = i - s * hw, y = r / W, x = r - y * W;
        const float* pi = pv.p + b * pv.sb + s * pv.ss + y * pv.sh + x;
        const float* ri = rv.p + b * rv.sb + y * rv.sh + x;
        const float r0 = ri[0], r1 = ri[rv.sc], r2 = ri[2 * rv.sc];
        const float t0 = tonemap(r0), t1 = tonemap(r1), t2 = tonemap(r2);
        // torch.clamp propagates NaN and inf/(1+inf) is NaN, fmaxf does not: test the raw values
        // (-inf clamps to 0 and is fine, as in the reference)
        bool bad = isnan(r0) || isnan(r1) || isnan(r2) || r0 == INFINITY || r1 == INFINITY || r2 == INFINITY;
        // partner inside the patch (same b)
        const int j = static_cast<int>(idx_patch[i]);
        const int js = j / hw, jr = j - js * hw, jy = jr / W, jx = jr - jy * W;
        const float* pj = pv.p + b * pv.sb + js * pv.ss + jy * pv.sh + jx;
        const float* rj = rv.p + b * rv.sb + jy * rv.sh + jx;
        // partner anywhere in the batch
        const float* pk = nullptr;
        const float* rk = nullptr;
        const long gi = static_cast<long>(b) * n + i;
        if (idx_batch != nullptr) {
            const long k = idx_batch[gi];
            const int kb = static_cast<int>(k / n);
            const int kl = static_cast<int>(k - static_cast<long>(kb) * n);
            const int ks = kl / hw, kr = kl - ks * hw, ky = kr / W, kx = kr - ky * W;
            pk = pv.p + kb * pv.sb + ks * pv.ss + ky * pv.sh + kx;
            rk = rv.p + kb * rv.sb + ky * rv.sh + kx;
            inv_batch[k] = static_cast<int32_t>(gi);
        }
        if (b == 0) inv_patch[j] = i;
        float dp = 0.0f, db = 0.0f;
        for (int c = 0; c < C; ++c) {
            const float v = pi[c * pv.sc];
            bad |= !isfinite(v);
            const float a = v - pj[c * pv.sc];
            dp = fmaf(a, a, dp);
            if (pk != nullptr) {
                const float g = v - pk[c * pv.sc];
                db = fmaf(g, g, db);
            }
        }
        {
            const float a0 = t0 - tonemap(rj[0]), a1 = t1 - tonemap(rj[rv.sc]), a2 = t2 - tonemap(rj[2 * rv.sc]);
            ep = 0.5f * dp - 0.5f * (a0 * a0 + a1 * a1 + a2 * a2);
            e_patch[gi] = ep;
        }
        if (pk != nullptr) {
            const float a0 = t0 - tonemap(rk[0]), a1 = t1 - tonemap(rk[rv.sc]), a2 = t2 - tonemap(rk[2 * rv.sc]);
            eb = 0.5f * db - 0.5f * (a0 * a0 + a1 * a1 + a2 * a2);
            e_batch[gi] = eb;
        }
        if (bad) atomicOr(nonfinite, 1);
    }
    // block partial sums of e^2 (fixed order -> deterministic)
    __shared__ float sm[2][kThreads / 32];
    float sp = wcmc::warp_sum(ep * ep), sb = wcmc::warp_sum(eb * eb);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        sm[0][warp] = sp;
        sm[1][warp] = sb;
    }
    __syncthreads();
    if (warp == 0) {
        sp = lane < kThreads / 32 ? sm[0][lane] : 0.0f;
        sb = lane < kThreads / 32 ? sm[1][lane] : 0.0f;
        sp = wcmc::warp_sum(sp);
        sb = wcmc::warp_sum(sb);
        if (lane == 0) {
            const int blk = blockIdx.y * gridDim.x + blockIdx.x;
            partial[2 * blk] = sp;
            partial[2 * blk + 1] = sb;
        }
    }
}

// loss[0] = 1/2 mean e_patch^2, loss[1] = 1/2 mean e_batch^2 (both means over B*n rows)
__global__ void __launch_bounds__(kThreads)
fmse_finish_kernel(const float* __restrict__ partial, int nblocks, double inv_rows, float* __restrict__ loss) {
    wcmc::pdl_start();
    double sp = 0.0, sb = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += kThreads) {
        sp += partial[2 * i];
        sb += partial[2 * i + 1];
    }
    __shared__ double sm[2][kThreads];
    sm[0][threadIdx.x] = sp;
    sm[1][threadIdx.x] = sb;
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sm[0][threadIdx.x] += sm[0][threadIdx.x + o];
            sm[1][threadIdx.x] += sm[1][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        loss[0] = static_cast<float>(0.5 * sm[0][0] * inv_rows);
        loss[1] = static_cast<float>(0.5 * sm[1][0] * inv_rows);
    }
}

// dp (B,S,C,H,W) contiguous
__global__ void __launch_bounds__(kThreads)
fmse_perm_bwd_kernel(PView pv, const int64_t* __restrict__ idx_patch, const int64_t* __restrict__ idx_batch,
                     const int32_t* __restrict__ inv_patch, const int32_t* __restrict__ inv_batch,
                     const float* __restrict__ w_patch, const float* __restrict__ w_batch,
                     const float* __restrict__ scale, float coef_patch, float coef_batch, int B, int S, int C,
                     int H, int W, float* __restrict__ dp, long d_sb, long d_ss, long d_sc, long d_sh) {
    wcmc::pdl_start();
    const int hw = H * W;
    const int n = S * hw;
    const int b = blockIdx.y;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const float g = scale != nullptr ? *scale : 1.0f;
    const long gi = static_cast<long>(b) * n + i;
    const int s = i / hw, r = i - s * hw, y = r / W, x = r - y * W;
    const float* pi = pv.p + b * pv.sb + s * pv.ss + y * pv.sh + x;
    auto row = [&](int bb, int l) {
        const int ls = l / hw, lr = l - ls * hw, ly = lr / W, lx = lr - ly * W;
        return pv.p + bb * pv.sb + ls * pv.ss + ly * pv.sh + lx;
    };
    // patch mode: partners pi(i) and pi^-1(i), weights of rows i and pi^-1(i) (same batch item)
    const int j = static_cast<int>(idx_patch[i]);
    const int ji = inv_patch[i];
    const float* pj = row(b, j);
    const float* pji = row(b, ji);
    const float w_i = w_patch[gi] * coef_patch * g;
    const float w_ji = w_patch[static_cast<long>(b) * n + ji] * coef_patch * g;
    const float* pk = nullptr;
    const float* pki = nullptr;
    float w_bi = 0.0f, w_ki = 0.0f;
    if (idx_batch != nullptr) {
        const long k = idx_batch[gi];
        const long ki = inv_batch[gi];
        const int kb = static_cast<int>(k / n), kib = static_cast<int>(ki / n);
        pk = row(kb, static_cast<int>(k - static_cast<long>(kb) * n));
        pki = row(kib, static_cast<int>(ki - static_cast<long>(kib) * n));
        w_bi = w_batch[gi] * coef_batch * g;
        w_ki = w_batch[ki] * coef_batch * g;
    }
    float* out = dp + b * d_sb + s * d_ss + y * d_sh + x;
    for (int c = 0; c < C; ++c) {
        const float v = pi[c * pv.sc];
        float d = w_i * (v - pj[c * pv.sc]) + w_ji * (v - pji[c * pv.sc]);
        if (pk != nullptr) d += w_bi * (v - pk[c * pv.sc]) + w_ki * (v - pki[c * pv.sc]);
        out[c * d_sc] = d;
    }
}

}  // namespace

extern "C" size_t wcmc_fmse_perm_workspace(int B, int S, int H, int W) {
    const long n = static_cast<long>(S) * H * W;
    const long blocks = (n + kThreads - 1) / kThreads * B;
    return static_cast<size_t>(blocks) * 2 * sizeof(float);
}

extern "C" int wcmc_fmse_perm_fwd(const float* p, long p_sb, long p_ss, long p_sc, long p_sh, const float* ref,
                                  long r_sb, long r_sc, long r_sh, const int64_t* idx_patch,
                                  const int64_t* idx_batch, int B, int S, int C, int H, int W, float* e_patch,
                                  float* e_batch, int32_t* inv_patch, int32_t* inv_batch, float* loss,
                                  int* nonfinite, void* workspace, size_t workspace_bytes, void* stream) {
    WCMC_REQUIRE(B > 0 && S > 0 && C > 0 && H > 0 && W > 0 && B <= 65535, WCMC_ESHAPE,
                 "fmse_perm_fwd: bad shape B=%d S=%d C=%d H=%d W=%d", B, S, C, H, W);
    WCMC_REQUIRE(static_cast<long>(B) * S * H * W < (1l << 31), WCMC_ESHAPE, "fmse_perm_fwd: more than 2^31 rows");
    WCMC_REQUIRE(p && ref && idx_patch && e_patch && inv_patch && loss && nonfinite, WCMC_ESHAPE,
                 "fmse_perm_fwd: null pointer");
    WCMC_REQUIRE(idx_batch == nullptr || (e_batch && inv_batch), WCMC_ESHAPE,
                 "fmse_perm_fwd: idx_batch given without e_batch / inv_batch");
    WCMC_REQUIRE(workspace_bytes >= wcmc_fmse_perm_workspace(B, S, H, W), WCMC_EWORKSPACE,
                 "fmse_perm_fwd: workspace too small");
    const int n = S * H * W;
    dim3 grid((n + kThreads - 1) / kThreads, B);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PView pv{p, p_sb, p_ss, p_sc, p_sh};
    RView rv{ref, r_sb, r_sc, r_sh};
    WCMC_LAUNCH(fmse_perm_fwd_kernel, grid, kThreads, 0, st, pv, rv, idx_patch, idx_batch, B, S, C, H, W, e_patch, e_batch,
                                                    inv_patch, inv_batch, static_cast<float*>(workspace),
                                                    nonfinite);
    WCMC_LAUNCH_CHECK();
    WCMC_LAUNCH(fmse_finish_kernel, 1, kThreads, 0, st, static_cast<const float*>(workspace), grid.x * grid.y,
                                               1.0 / (static_cast<double>(B) * n), loss);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" int wcmc_fmse_perm_bwd_strided(const float* p, long p_sb, long p_ss, long p_sc, long p_sh,
                                          const int64_t* idx_patch, const int64_t* idx_batch, const int32_t* inv_patch,
                                          const int32_t* inv_batch, const float* w_patch, const float* w_batch,
                                          const float* scale, float coef_patch, float coef_batch, int B, int S, int C,
                                          int H, int W, float* dp, long d_sb, long d_ss, long d_sc, long d_sh,
                                          void* stream) {
    WCMC_REQUIRE(B > 0 && S > 0 && C > 0 && H > 0 && W > 0 && B <= 65535, WCMC_ESHAPE,
                 "fmse_perm_bwd: bad shape B=%d S=%d C=%d H=%d W=%d", B, S, C, H, W);
    WCMC_REQUIRE(p && idx_patch && inv_patch && w_patch && dp, WCMC_ESHAPE, "fmse_perm_bwd: null pointer");
    WCMC_REQUIRE(idx_batch == nullptr || (inv_batch && w_batch), WCMC_ESHAPE,
                 "fmse_perm_bwd: idx_batch given without inv_batch / w_batch");
    const int n = S * H * W;
    dim3 grid((n + kThreads - 1) / kThreads, B);
    PView pv{p, p_sb, p_ss, p_sc, p_sh};
    WCMC_LAUNCH(fmse_perm_bwd_kernel, grid, kThreads, 0, static_cast<cudaStream_t>(stream), pv, idx_patch, idx_batch, inv_patch, inv_batch, w_patch, w_batch, scale, coef_patch, coef_batch, B, S, C, H,
        W, dp, d_sb, d_ss, d_sc, d_sh);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" int wcmc_fmse_perm_bwd(const float* p, long p_sb, long p_ss, long p_sc, long p_sh,
                                  const int64_t* idx_patch, const int64_t* idx_batch, const int32_t* inv_patch,
                                  const int32_t* inv_batch, const float* w_patch, const float* w_batch,
                                  const float* scale, float coef_patch, float coef_batch, int B, int S, int C,
                                  int H, int W, float* dp, void* stream) {
    const long hw = static_cast<long>(H) * W;
    return wcmc_fmse_perm_bwd_strided(p, p_sb, p_ss, p_sc, p_sh, idx_patch, idx_batch, inv_patch, inv_batch, w_patch,
                                      w_batch, scale, coef_patch, coef_batch, B, S, C, H, W, dp, S * C * hw, C * hw, hw, W,
                                      stream);
}
