// K11: ALL-PAIRS form of the path-disentangling loss on the tensor cores (extension; the reference's
// FeatureMSE pairs each row with one random partner, /root/reference/support/losses.py:33-61 -- see
// SURVEY.md 0.3 and oracle/allpairs_ref.py for the definition this kernel is checked against).
//
//   d_p(i,j) = 1/2 |P_i - P_j|^2 = hp_i + hp_j - P_i.P_j        hp = 1/2 |P|^2
//   d_t(i,j) = 1/2 |t_i - t_j|^2 = ht_i + ht_j - t_i.t_j        t  = tone-mapped reference colour
//   e = d_p - d_t ;  mse: sum 1/2 e^2 / N^2 ;  lse: logsumexp(alpha [e, -e, 0]) ;  optional weak-label
//   mask keep(i,j) = d_t < tau.
//
// The Gram matrices come from tcgen05.mma (M = 128 rows i, N = 128 rows j, fp32 accumulation in TMEM),
// everything else happens in the epilogue on the accumulator tile, so the N x N matrix never exists in
// memory.  fp32 accuracy from 16-bit operands: every value is split v = hi + lo (two fp16, 22
// significant bits) and the K axis carries [hi | hi | lo] against [hi | lo | hi], i.e. one GEMM adds
// hi.hi + hi.lo + lo.hi.  Only tiles on or above the diagonal are visited (e is symmetric, e_ii = 0).
// Unmasked: the label columns of the B operand are negated so ONE accumulator holds P.P - t.t; masked:
// a second accumulator holds t.t so d_t is available per pair.
//
// Bound: K <= 112 per 128 x 128 tile means 4-7 MMAs (<= 450 cycles) against a 16 K-element epilogue that
// must pull 64 KB (128 KB masked) out of TMEM and spend ~5 FP32 operations per pair: TMEM-read /
// issue-bound, NOT tensor-pipe-bound, as SURVEY.md 7.3-8 predicted.  Warp roles (320 threads): TMA
// producer | MMA issuer | 8 epilogue warps; A block resident per work item, 3-stage B ring,
// double-buffered accumulators.
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace wcmc {

constexpr int kApThreads = 64 + 256;
constexpr int kApBStages = 3;
constexpr int kApChunk = 16;          // B tiles per work item (the A block is loaded once per item)
constexpr int kApTile = 16384;        // 128 rows x 64 halves
constexpr int kApSmem = 1024 + 2 * kApTile + kApBStages * 2 * kApTile + 8 * 64 * 8 + 256;

__device__ __forceinline__ float ap_tonemap(float v) {
    v = fmaxf(v, 0.0f);
    return powf(v / (1.0f + v), 0.454545f);
}
__device__ __forceinline__ float ap_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// one thread per (padded) row
__global__ void __launch_bounds__(256)
allpairs_prep_kernel(const float* __restrict__ P, const float* __restrict__ R, int N, int Npad, int D, int Kp1,
                     int Kp, float tsign, __half* __restrict__ A, __half* __restrict__ B, float2* __restrict__ h,
                     int* __restrict__ nonfinite) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Npad) return;
    __half* a = A + static_cast<size_t>(i) * Kp;
    __half* b = B + static_cast<size_t>(i) * Kp;
    for (int k = 0; k < Kp; k += 8) {
        *reinterpret_cast<uint4*>(a + k) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(b + k) = make_uint4(0, 0, 0, 0);
    }
    if (i >= N) {
        h[i] = make_float2(0.f, 0.f);
        return;
    }
    float hp = 0.f, ht = 0.f;
    bool bad = false;
    for (int c = 0; c < D; ++c) {
        const float v = P[static_cast<size_t>(i) * D + c];
        bad |= !isfinite(v);
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        a[c] = hi; a[D + c] = hi; a[2 * D + c] = lo;
        b[c] = hi; b[D + c] = lo; b[2 * D + c] = hi;
        hp = fmaf(v, v, hp);
    }
    for (int c = 0; c < 3; ++c) {
        const float r = R[static_cast<size_t>(i) * 3 + c];
        bad |= isnan(r) || r == INFINITY;
        const float t = ap_tonemap(r);
        const __half hi = __float2half_rn(t);
        const __half lo = __float2half_rn(t - __half2float(hi));
        a[Kp1 + c] = hi; a[Kp1 + 3 + c] = hi; a[Kp1 + 6 + c] = lo;
        b[Kp1 + c] = __float2half_rn(tsign * __half2float(hi));
        b[Kp1 + 3 + c] = __float2half_rn(tsign * __half2float(lo));
        b[Kp1 + 6 + c] = __float2half_rn(tsign * __half2float(hi));
        ht = fmaf(t, t, ht);
    }
    h[i] = make_float2(0.5f * hp, 0.5f * ht);
    if (bad) atomicOr(nonfinite, 1);
}

struct ApParams {
    int N, NT;            // rows, 128-row tiles
    int nk1, kchunks;     // K16 steps of the embedding part; 64-half chunks of the operand rows
    int mode, masked;     // 0 mse / 1 lse
    float alpha_l2e, tau;
    const float2* h;      // (hp, ht) per row
    double* partial;      // per CTA: sum e^2 | lse max (log2 domain) | lse sum | kept unordered pairs
};

struct ApCursor {
    int ti, c;
    __device__ __forceinline__ void advance(int steps, int NT) {
        while (ti < NT) {
            const int rc = (NT - ti + kApChunk - 1) / kApChunk;
            if (c + steps < rc) { c += steps; return; }
            steps -= rc - c;
            ++ti;
            c = 0;
        }
    }
};

template <int MODE, bool MASKED>
__global__ void __launch_bounds__(kApThreads, 1)
allpairs_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const ApParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint8_t* at = smem;                           // 2 chunk tiles
    uint8_t* bt = smem + 2 * kApTile;             // stages x 2 chunk tiles
    float2* hcol = reinterpret_cast<float2*>(bt + kApBStages * 2 * kApTile);   // 8 warps x 64 columns
    uint64_t* bars = reinterpret_cast<uint64_t*>(hcol + 8 * 64);
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + 1;
    uint64_t* b_full = bars + 2;
    uint64_t* b_empty = b_full + kApBStages;
    uint64_t* acc_full = b_empty + kApBStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
    __shared__ double red[256][4];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int i = 0; i < kApBStages; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 256);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int NT = p.NT;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            ApCursor cur{0, 0};
            cur.advance(blockIdx.x, NT);
            int it = 0, bs = 0, bph = 0;
            for (; cur.ti < NT; cur.advance(gridDim.x, NT), ++it) {
                mbar_wait(a_empty, (it & 1) ^ 1);
                mbar_expect_tx(a_full, p.kchunks * kApTile);
                for (int k = 0; k < p.kchunks; ++k) tma_load_2d(at + k * kApTile, &tmA, a_full, k * 64, cur.ti * 128);
                const int tj0 = cur.ti + cur.c * kApChunk;
                const int ntj = min(kApChunk, NT - tj0);
                for (int t = 0; t < ntj; ++t) {
                    mbar_wait(&b_empty[bs], bph ^ 1);
                    mbar_expect_tx(&b_full[bs], p.kchunks * kApTile);
                    for (int k = 0; k < p.kchunks; ++k)
                        tma_load_2d(bt + (bs * 2 + k) * kApTile, &tmB, &b_full[bs], k * 64, (tj0 + t) * 128);
                    if (++bs == kApBStages) { bs = 0; bph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = make_idesc_f16(128, 128, 0, 0, WCMC_F16, WCMC_F16);
        const uint32_t a_lo = smem_u32(at) >> 4, b_lo0 = smem_u32(bt) >> 4;
        const uint64_t hi = make_sdesc_sw128(0, 16, 1024) & 0xFFFFFFFF00000000ull;
        const uint32_t lo_fixed = 1u << 16;
        auto koff = [](int k) { return static_cast<uint32_t>((k >> 2) * (kApTile >> 4) + (k & 3) * 2); };
        ApCursor cur{0, 0};
        cur.advance(blockIdx.x, NT);
        int it = 0, bs = 0, bph = 0, tile = 0;
        for (; cur.ti < NT; cur.advance(gridDim.x, NT), ++it) {
            mbar_wait(a_full, it & 1);
            const int tj0 = cur.ti + cur.c * kApChunk;
            const int ntj = min(kApChunk, NT - tj0);
            for (int t = 0; t < ntj; ++t, ++tile) {
                const int buf = tile & 1;
                mbar_wait(&acc_empty[buf], ((tile >> 1) & 1) ^ 1);
                mbar_wait(&b_full[bs], bph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d = tmem_base + buf * 256;
                    const uint32_t b_lo = b_lo0 + bs * 2 * (kApTile >> 4);
                    for (int k = 0; k < p.nk1; ++k)
                        umma_bf16(d, hi | (lo_fixed | (a_lo + koff(k))), hi | (lo_fixed | (b_lo + koff(k))), idesc,
                                  k > 0 ? 1u : 0u);
                    const int kt = p.nk1;   // the label columns: their own K16 step
                    umma_bf16(p.masked ? d + 128 : d, hi | (lo_fixed | (a_lo + koff(kt))),
                              hi | (lo_fixed | (b_lo + koff(kt))), idesc, p.masked ? 0u : 1u);
                    umma_commit(&b_empty[bs]);
                    umma_commit(&acc_full[buf]);
                }
                __syncwarp();
                if (++bs == kApBStages) { bs = 0; bph ^= 1; }
            }
            if (elect_one()) umma_commit(a_empty);
            __syncwarp();
        }
    } else {
        // ---------------- epilogue: 8 warps, lane quarter q, column half hs ----------------
        // Compile-time MODE / MASKED and a predicate-free path for tiles strictly above the diagonal keep
        // the per-pair cost at ~3 FP32 instructions (mse): add, subtract, fused multiply-add.
        const int ew = warp - 2;
        const int q = warp & 3, hs = ew >> 2;
        float2* hw = hcol + ew * 64;               // per-warp copy of (hp, ht) [masked] or (hp - ht, -) of its columns
        double sum = 0.0;          // sum of e^2 over kept pairs j > i
        double cnt = 0.0;
        float lm = 0.f, ls = 0.f;  // running logsumexp in the log2 domain: sum 2^(x - lm)
        ApCursor cur{0, 0};
        cur.advance(blockIdx.x, NT);
        int tile = 0;
        auto lse_add = [&](float e) {
            const float x = e * p.alpha_l2e, a = fabsf(x);
            if (a > lm) {
                ls *= ap_ex2(lm - a);
                lm = a;
            }
            ls += ap_ex2(x - lm) + ap_ex2(-x - lm);
        };
        for (; cur.ti < NT; cur.advance(gridDim.x, NT)) {
            const int i = cur.ti * 128 + q * 32 + lane;
            const float2 hi2 = p.h[i];
            const float hd_i = hi2.x - hi2.y;
            const int tj0 = cur.ti + cur.c * kApChunk;
            const int ntj = min(kApChunk, NT - tj0);
            for (int t = 0; t < ntj; ++t, ++tile) {
                const int buf = tile & 1;
                const int j0 = (tj0 + t) * 128 + hs * 64;
                __syncwarp();
                {
                    const float2 a = p.h[j0 + lane], b = p.h[j0 + lane + 32];
                    hw[lane] = MASKED ? a : make_float2(a.x - a.y, 0.f);
                    hw[lane + 32] = MASKED ? b : make_float2(b.x - b.y, 0.f);
                }
                __syncwarp();
                mbar_wait(&acc_full[buf], (tile >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 256 + hs * 64;
                const bool interior = (tj0 + t > cur.ti) && (j0 + 64 <= p.N) && (cur.ti * 128 + 128 <= p.N);
                float ts0 = 0.f, ts1 = 0.f;
                int tcnt = 0;
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    uint32_t gp[16], gt[16];
                    tmem_ld16(taddr + cc * 16, gp);
                    if (MASKED) tmem_ld16(taddr + 128 + cc * 16, gt);
                    tmem_ld_wait16(gp);
                    if (MASKED) tmem_ld_wait16(gt);
                    if (interior && !MASKED) {
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            const float e = (hd_i + hw[cc * 16 + c].x) - __uint_as_float(gp[c]);
                            if (MODE == 0) {
                                if (c & 1) ts1 = fmaf(e, e, ts1); else ts0 = fmaf(e, e, ts0);
                            } else {
                                lse_add(e);
                            }
                        }
                        tcnt += 16;
                    } else {
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            const int j = j0 + cc * 16 + c;
                            const float2 hj = hw[cc * 16 + c];
                            bool keep = interior || ((j > i) && (j < p.N) && (i < p.N));
                            float e;
                            if (MASKED) {
                                const float dt = (hi2.y + hj.y) - __uint_as_float(gt[c]);
                                e = ((hi2.x + hj.x) - __uint_as_float(gp[c])) - dt;
                                keep = keep && (dt < p.tau);
                            } else {
                                e = (hd_i + hj.x) - __uint_as_float(gp[c]);
                            }
                            if (MODE == 0) {
                                e = keep ? e : 0.f;
                                ts0 = fmaf(e, e, ts0);
                            } else if (keep) {
                                lse_add(e);
                            }
                            tcnt += keep ? 1 : 0;
                        }
                    }
                }
                sum += static_cast<double>(ts0 + ts1);
                cnt += static_cast<double>(tcnt);
                tc_fence_before();
                mbar_arrive(&acc_empty[buf]);
            }
        }
        const int e = threadIdx.x - 64;
        red[e][0] = sum;
        red[e][1] = static_cast<double>(lm);
        red[e][2] = static_cast<double>(ls);
        red[e][3] = cnt;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
    if (threadIdx.x == 64) {
        double sum = 0.0, cnt = 0.0, m = 0.0, s = 0.0;
        for (int k = 0; k < 256; ++k) {
            sum += red[k][0];
            cnt += red[k][3];
            const double mk = red[k][1], sk = red[k][2];
            if (sk > 0.0) {
                const double mn = fmax(m, mk);
                s = s * exp2(m - mn) + sk * exp2(mk - mn);
                m = mn;
            }
        }
        double* o = p.partial + 4 * blockIdx.x;
        o[0] = sum; o[1] = m; o[2] = s; o[3] = cnt;
    }
}

// out[0] = loss, out[1] = number of kept ORDERED pairs (i != j)
__global__ void allpairs_finish_kernel(const double* __restrict__ partial, int nparts, int N, int mode,
                                       float alpha, float* __restrict__ out) {
    if (threadIdx.x != 0) return;
    double sum = 0.0, cnt = 0.0, m = 0.0, s = 0.0;
    for (int k = 0; k < nparts; ++k) {
        sum += partial[4 * k];
        cnt += partial[4 * k + 3];
        const double mk = partial[4 * k + 1], sk = partial[4 * k + 2];
        if (sk > 0.0) {
            const double mn = fmax(m, mk);
            s = s * exp2(m - mn) + sk * exp2(mk - mn);
            m = mn;
        }
    }
    if (mode == 0) {
        out[0] = static_cast<float>(sum / (static_cast<double>(N) * N));   // 2 * sum_{j>i} 1/2 e^2 / N^2
    } else {
        // multiset over ordered pairs: every unordered pair contributes its two terms twice; plus the 0 term
        const double tot = 2.0 * s + exp2(-m);           // in units of 2^m
        const double lse = (m + log2(tot)) * 0.6931471805599453;
        out[0] = static_cast<float>((lse - log(1.0 + 4.0 * cnt)) / sqrt(static_cast<double>(alpha)));
    }
    out[1] = static_cast<float>(2.0 * cnt);
}

}  // namespace wcmc

using namespace wcmc;

static void ap_dims(int N, int D, int* Npad, int* Kp1, int* Kp) {
    *Npad = (N + 127) / 128 * 128;
    *Kp1 = (3 * D + 15) / 16 * 16;
    *Kp = (*Kp1 + 16 + 63) / 64 * 64;
}

extern "C" size_t wcmc_fmse_allpairs_workspace(int N, int D) {
    int Npad, Kp1, Kp;
    ap_dims(N, D, &Npad, &Kp1, &Kp);
    return static_cast<size_t>(Npad) * Kp * 2 * 2 + static_cast<size_t>(Npad) * 8 + 4096 * 8 + 256;
}

extern "C" int wcmc_fmse_allpairs_fwd(const float* p_rows, const float* ref_rows, int N, int D, int mode,
                                      float alpha, float tau, float* out, int* nonfinite, void* workspace,
                                      size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(N > 1 && D >= 1 && 3 * D + 16 <= 128 + 0 && N <= (1 << 24), WCMC_ESHAPE,
                 "fmse_allpairs_fwd: need 1 < N <= 2^24 and D <= 37 (got N=%d D=%d)", N, D);
    WCMC_REQUIRE(mode == 0 || mode == 1, WCMC_ESHAPE, "fmse_allpairs_fwd: mode %d not in {0 mse, 1 lse}", mode);
    WCMC_REQUIRE(p_rows && ref_rows && out && nonfinite && workspace, WCMC_ESHAPE, "fmse_allpairs_fwd: null pointer");
    WCMC_REQUIRE(workspace_bytes >= wcmc_fmse_allpairs_workspace(N, D), WCMC_EWORKSPACE,
                 "fmse_allpairs_fwd: workspace too small");
    WCMC_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, WCMC_EALIGN,
                 "fmse_allpairs_fwd: workspace must be 256-byte aligned");
    int Npad, Kp1, Kp;
    ap_dims(N, D, &Npad, &Kp1, &Kp);
    const bool masked = tau > 0.f;
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    __half* A = reinterpret_cast<__half*>(ws);
    __half* B = A + static_cast<size_t>(Npad) * Kp;
    float2* h = reinterpret_cast<float2*>(B + static_cast<size_t>(Npad) * Kp);
    double* partial = reinterpret_cast<double*>(h + Npad);
    allpairs_prep_kernel<<<(Npad + 255) / 256, 256, 0, stream>>>(p_rows, ref_rows, N, Npad, D, Kp1, Kp,
                                                                 masked ? 1.f : -1.f, A, B, h, nonfinite);
    WCMC_LAUNCH_CHECK();
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[2] = {static_cast<uint64_t>(Kp), static_cast<uint64_t>(Npad)};
        uint64_t strides[1] = {static_cast<uint64_t>(Kp) * 2};
        uint32_t box[2] = {64, 128};
        int rc = wcmc_encode_tmap(&tmA, WCMC_F16, A, 2, dims, strides, box, 1);
        if (rc) return rc;
        rc = wcmc_encode_tmap(&tmB, WCMC_F16, B, 2, dims, strides, box, 1);
        if (rc) return rc;
    }
    ApParams p;
    p.N = N; p.NT = Npad / 128; p.nk1 = Kp1 / 16; p.kchunks = Kp / 64;
    p.mode = mode; p.masked = masked ? 1 : 0;
    p.alpha_l2e = alpha * 1.4426950408889634f; p.tau = tau;
    p.h = h; p.partial = partial;
    long items = 0;
    for (int ti = 0; ti < p.NT; ++ti) items += (p.NT - ti + kApChunk - 1) / kApChunk;
    int grid = static_cast<int>(std::min<long>(items, wcmc_num_sms()));
#define WCMC_AP_LAUNCH(M, K)                                                                                  \
    do {                                                                                                     \
        WCMC_FUNC_SMEM((allpairs_kernel<M, K>), kApSmem);                                                    \
        allpairs_kernel<M, K><<<grid, kApThreads, kApSmem, stream>>>(tmA, tmB, p);                           \
    } while (0)
    if (mode == 0 && !masked) WCMC_AP_LAUNCH(0, false);
    else if (mode == 0) WCMC_AP_LAUNCH(0, true);
    else if (!masked) WCMC_AP_LAUNCH(1, false);
    else WCMC_AP_LAUNCH(1, true);
#undef WCMC_AP_LAUNCH
    WCMC_LAUNCH_CHECK();
    allpairs_finish_kernel<<<1, 32, 0, stream>>>(partial, grid, N, mode, alpha, out);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}
