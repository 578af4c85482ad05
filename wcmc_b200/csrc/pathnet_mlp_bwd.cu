// K8/K9: backward passes of the two per-sample MLPs of the path-embedding network, each fused into ONE kernel
// (autograd of /root/reference/support/networks.py:29-42 `embedding` / `final`; the reference runs them as
// 1x1 nn.Conv2d backward passes over the materialised (B*S,128,H,W) tensors).
//
//   K8  final bwd : g = dL/dout (B,S,outc,H,W) fp32, out, hfin, [emb | prop]  ->  d_emb (per sample), d_prop (summed
//                   over spp), dW1, db1, dW2, db2
//   K9  embed bwd : d_emb, d_red (gradient of the spp mean, from the U-Net), emb, h2, h1, x16  ->  dW3, db3, dW2,
//                   db2, dW1, db1
//
// Both are HBM-bound once the products run on the tensor cores: the generic path (1x1 conv dgrad / wgrad / bias /
// activation kernels) moves ~2.6 GB + ~1.5 GB per network at B=8, S=8; fused, every saved activation is read
// exactly once and only d_emb / d_prop are written (0.56 GB + 0.64 GB).
//
// Structure.  A CTA is two independent groups of 256 threads (two threads per sample-pixel row = TMEM lane, each
// owning one 64-column half of the row: the epilogues are latency chains, so warps, not instructions, are what was
// missing -- ncu: 9 warps per SM issued one instruction every 14 clocks) plus one MMA-issuing warp.  A group owns its own shared-memory tiles and 128 TMEM columns for the data gradients and walks
// (pixel tile, image) items, looping over the spp samples of a tile; while one group waits for its TMA loads or runs
// an epilogue, the other group's MMAs run.  Every tile is a [128 rows][64 channels] 16-bit SWIZZLE_128B tile, which
// is at once the K-major A operand of the data-gradient MMA  dX[row][cin] = sum_cout dZ[row][cout] W[cout][cin]
// and the MN-major operand of the weight-gradient MMA  dW[cout][cin] = sum_row dZ[row][cout] X[row][cin]
// (same bytes, other major-ness bit in the instruction descriptor -- cf. conv_wgrad.cu).  Activation derivatives are
// applied in place in the epilogue (the saved post-activation tile becomes the dZ tile of the layer below).  Bias
// gradients come out of the same weight-gradient MMA through a constant tile of ones appended to X.  Weight
// gradients accumulate in TMEM over all the tiles of the CTA (one single thread issues every MMA of both groups, so
// accumulation order is program order); each CTA writes ONE partial slab, and a small second kernel sums the slabs
// in a fixed order (deterministic, no atomics), applies 1/loss-scale and scatters into torch's parameter layout.
#include "common.cuh"

namespace wcmc {

constexpr int kBwdThreads = 544;             // 2 groups x 256 + issuer warp
constexpr int kGroupThreads = 256;
constexpr int kTile = 128 * 128;             // bytes of a [128][64] 16-bit tile

__device__ __forceinline__ uint32_t swz(int row, int chunk) {   // byte offset of 16-byte chunk `chunk` of row `row`
    return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ float dact(uint32_t h16, int act, float slope) {   // act'(.) from the saved 16-bit output
    if (act == WCMC_ACT_LINEAR) return 1.f;
    return h16_pos(h16) ? 1.f : (act == WCMC_ACT_LEAKY ? slope : 0.f);
}
__device__ __forceinline__ float dact_f(float y, int act, float slope) {
    if (act == WCMC_ACT_LINEAR) return 1.f;
    return y > 0.f ? 1.f : (act == WCMC_ACT_LEAKY ? slope : 0.f);
}
// rows x (k_total 16-bit) row-major global matrix -> swizzled [rows][64] tiles, one per 64-column block, `tile_stride`
// apart, starting at 16-byte chunk `chunk0` of every row (chunk0 + k_total/8 <= 8 when there is one tile)
__device__ __forceinline__ void load_rows_swizzled(uint8_t* tiles, int tile_stride, const void* w, int rows, int k_total,
                                                   int chunk0, int tid, int nthreads) {
    const int cpr = k_total >> 3;
    const uint4* src = static_cast<const uint4*>(w);
    for (int i = tid; i < rows * cpr; i += nthreads) {
        const int row = i / cpr, ch = i - row * cpr + chunk0;
        *reinterpret_cast<uint4*>(tiles + (ch >> 3) * tile_stride + swz(row, ch & 7)) = __ldg(src + i);
    }
}
// K-major operand descriptor of a dense swizzled tile; MN-major one with `lbo` bytes between 64-element planes
__device__ __forceinline__ uint64_t kdesc(uint32_t addr) { return make_sdesc_sw128(addr, 16, 1024); }
__device__ __forceinline__ uint64_t mndesc(uint32_t addr, uint32_t lbo) { return make_sdesc_sw128(addr, lbo, 1024); }

struct GroupSync {
    uint64_t tma_full, req, done;
};

// Issuer: bounded spin over the two groups' request barriers.
struct IssuerState {
    int step[2], total[2];
    uint32_t req_ph[2], tma_ph[2];
};
__device__ __forceinline__ void issuer_trap(const char* what) {
    printf("wcmc: %s issuer made no progress for 4 s (block %d)\n", what, blockIdx.x);
    __trap();
}

// ================================================================================================================
// K8: final MLP backward
// ================================================================================================================
struct FinalBwdParams {
    const float* g;        // (B,S,outc,HW) fp32
    const float* out;      // (B,S,outc,HW) fp32, post-activation forward output
    const float* gscale;   // device scalar (loss scale) or null
    void* d_emb;           // (B*S*HW, 64) 16-bit
    void* d_prop;          // (B*HW, 64) 16-bit
    float* partial;        // [gridDim.x][kFinSlab]
    const void* w1t;       // [128 cin][128 cout] 16-bit (the data-gradient packing of layer 1)
    const void* w2t;       // [128 cin][outc_p cout]
    int B, S, HW, outc, outc_p, dtype, act1, act2;
    float slope;
    int tiles, items;
};
constexpr int kFinSlab = 128 * 128 + 128 + 32 * 128 + 32;   // dW1 | db1 | dW2 [32][128] | db2 [32]
// shared: W1T (2 planes) | ONES | per group: H0 H1 EMB PROP ZW | barriers
constexpr int kFinGroupBytes = 5 * kTile;
constexpr int kFinSmem = 1024 + 3 * kTile + 2 * kFinGroupBytes + 2048;   // tail: barriers, TMEM pointer, db2 staging

template <int OP>   // OP = outc_p (16 or 32): sizes the per-thread dz2 / db2 registers
__global__ void __launch_bounds__(kBwdThreads, 1)
pathnet_final_bwd_kernel(const __grid_constant__ CUtensorMap tmh, const __grid_constant__ CUtensorMap tme,
                         const __grid_constant__ CUtensorMap tmpr, const FinalBwdParams p) {
    pdl_wait();                // the prologue below reads packed weights / issues TMA loads
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* w1t = smem;                 // plane c: cout in [64c, 64c+64)
    uint8_t* ones = smem + 2 * kTile;
    uint8_t* gbase = smem + 3 * kTile;
    GroupSync* sync = reinterpret_cast<GroupSync*>(gbase + 2 * kFinGroupBytes);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sync + 2);
    float* db2_smem = reinterpret_cast<float*>(tmem_ptr + 2);   // [8 warps][32]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = tid >> 8;            // 0, 1, or 2 (issuer warp)

    if (tid == 0) {
        tma_prefetch_desc(&tmh);
        tma_prefetch_desc(&tme);
        tma_prefetch_desc(&tmpr);
        for (int g = 0; g < 2; ++g) {
            mbar_init(&sync[g].tma_full, 1);
            mbar_init(&sync[g].req, kGroupThreads);
            mbar_init(&sync[g].done, 1);
        }
        fence_barrier_init();
    }
    // constant tiles: W1T, ones, and per group the W2T part of ZW (rest of ZW zero)
    {
        const uint32_t one2 = p.dtype == WCMC_F16 ? 0x3C003C00u : 0x3F803F80u;
        for (int i = tid; i < kTile / 16; i += kBwdThreads) {
            reinterpret_cast<uint4*>(ones)[i] = make_uint4(one2, one2, one2, one2);
            reinterpret_cast<uint4*>(gbase + 4 * kTile)[i] = make_uint4(0, 0, 0, 0);
            reinterpret_cast<uint4*>(gbase + kFinGroupBytes + 4 * kTile)[i] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
        load_rows_swizzled(w1t, kTile, p.w1t, 128, 128, 0, tid, kBwdThreads);
        load_rows_swizzled(gbase + 4 * kTile, 0, p.w2t, 128, p.outc_p, 4, tid, kBwdThreads);
        load_rows_swizzled(gbase + kFinGroupBytes + 4 * kTile, 0, p.w2t, 128, p.outc_p, 4, tid, kBwdThreads);
    }
    if (warp == 16) tmem_alloc(tmem_ptr, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_launch_dependents();   // after the TMEM allocation (common.cuh: PDL rules)
    const uint32_t tmem = *tmem_ptr;
    // TMEM columns: [0,128) / [128,256) data gradients of group 0 / 1; [256,384) dW1; [384,448) db1 (ones block);
    // [448,512) dW2^T
    constexpr uint32_t kColW1 = 256, kColB1 = 384, kColW2 = 448;
    const int ngroups = 2 * gridDim.x;

    if (grp == 2) {
        // ------------------------------------------ MMA issuer ------------------------------------------
        if (lane == 0) {
            IssuerState st;
            for (int g = 0; g < 2; ++g) {
                const int gi = blockIdx.x * 2 + g;
                const int n = gi < p.items ? (p.items - gi + ngroups - 1) / ngroups : 0;
                st.step[g] = 0; st.total[g] = n * p.S * 2; st.req_ph[g] = 0; st.tma_ph[g] = 0;
            }
            const uint32_t id_dh = make_idesc_f16(128, 128, 0, 0, p.dtype, p.dtype);      // K-major x K-major
            const uint32_t id_w = make_idesc_f16(128, 128, 1, 1, p.dtype, p.dtype);       // MN-major x MN-major
            const uint32_t id_w64 = make_idesc_f16(128, 64, 1, 1, p.dtype, p.dtype);
            // Every descriptor is built ONCE: this single thread feeds the tensor pipe for both groups, and with
            // descriptors rebuilt per MMA it was the bottleneck of the kernel (one warp retires an instruction every
            // ~14 clocks; ~45 instructions per MMA x 33 MMAs per sample tile = 12 us).
            const uint64_t w1k0 = kdesc(smem_u32(w1t)), w1k1 = kdesc(smem_u32(w1t) + kTile);
            const uint64_t ones_mn = mndesc(smem_u32(ones), kTile);
            const int k2 = p.outc_p >> 4;
            bool first_w2 = true, first_w1 = true;
            long long last = clock64();
            auto serve = [&](const int g) {
                const uint32_t base = smem_u32(gbase + g * kFinGroupBytes);
                const uint64_t hk0 = kdesc(base), hk1 = kdesc(base + kTile), zk = kdesc(base + 4 * kTile);
                const uint64_t h_mn = mndesc(base, kTile), e_mn = mndesc(base + 2 * kTile, kTile),
                               z_mn = mndesc(base + 4 * kTile, kTile);
                const uint32_t dcol = tmem + g * 128;
                if ((st.step[g] & 1) == 0) {
                    // R1: dH = Z2 . W2 (K = outc_p);  dW2^T += H^T . Z2  (K = 128 rows).  The group posts the request
                    // only after its TMA tiles have landed: this thread serves both groups and must never block.
                    tc_fence_after();
                    umma_bf16(dcol, zk, zk + 4, id_dh, 0u);
                    if (k2 > 1) umma_bf16(dcol, zk + 2, zk + 6, id_dh, 1u);
                    umma_bf16(tmem + kColW2, h_mn, z_mn, id_w64, first_w2 ? 0u : 1u);
#pragma unroll
                    for (int j = 1; j < 8; ++j) umma_bf16(tmem + kColW2, h_mn + j * 128, z_mn + j * 128, id_w64, 1u);
                    first_w2 = false;
                } else {
                    // R2: dBoth = Z1 . W1 (K = 128 channels);  dW1 += Z1^T . [emb | prop];  db1 += Z1^T . 1
                    tc_fence_after();
                    umma_bf16(dcol, hk0, w1k0, id_dh, 0u);
#pragma unroll
                    for (int k = 1; k < 4; ++k) umma_bf16(dcol, hk0 + 2 * k, w1k0 + 2 * k, id_dh, 1u);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(dcol, hk1 + 2 * k, w1k1 + 2 * k, id_dh, 1u);
                    const uint32_t acc0 = first_w1 ? 0u : 1u;
                    umma_bf16(tmem + kColW1, h_mn, e_mn, id_w, acc0);
                    umma_bf16(tmem + kColB1, h_mn, ones_mn, id_w64, acc0);
#pragma unroll
                    for (int j = 1; j < 8; ++j) {
                        umma_bf16(tmem + kColW1, h_mn + j * 128, e_mn + j * 128, id_w, 1u);
                        umma_bf16(tmem + kColB1, h_mn + j * 128, ones_mn + j * 128, id_w64, 1u);
                    }
                    first_w1 = false;
                }
                umma_commit(&sync[g].done);
                ++st.step[g];
            };
            while (st.step[0] < st.total[0] || st.step[1] < st.total[1]) {
                if (st.step[0] < st.total[0] && mbar_test_wait(&sync[0].req, st.req_ph[0])) {
                    st.req_ph[0] ^= 1;
                    last = clock64();
                    serve(0);
                }
                if (st.step[1] < st.total[1] && mbar_test_wait(&sync[1].req, st.req_ph[1])) {
                    st.req_ph[1] ^= 1;
                    last = clock64();
                    serve(1);
                }
                if (clock64() - last > 8000000000LL) issuer_trap("pathnet_final_bwd");
            }
        }
    } else {
        // ------------------------------------------ row groups ------------------------------------------
        const int r = tid & 127;                 // row of the tile = TMEM lane
        const int half = (tid >> 7) & 1;         // which 64-column half of the row this thread owns
        uint8_t* base = gbase + grp * kFinGroupBytes;
        uint8_t* h0 = base;
        uint8_t* et = base + 2 * kTile;
        uint8_t* pt = base + 3 * kTile;
        uint8_t* zw = base + 4 * kTile;
        GroupSync& sy = sync[grp];
        const uint32_t tlane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + grp * 128;
        const float gs = p.gscale != nullptr ? __ldg(p.gscale) : 1.f;
        const int dt = p.dtype;
        uint32_t tma_ph = 0, done_ph = 0;
        float db2[OP];
#pragma unroll
        for (int c = 0; c < OP; ++c) db2[c] = 0.f;

        for (int item = blockIdx.x * 2 + grp; item < p.items; item += ngroups) {
            const int b = item / p.tiles, pix0 = (item - b * p.tiles) * 128;
            const int pix = pix0 + r;
            const bool valid = pix < p.HW;
            float acc[32];   // running sum over spp of this thread's 32 prop-gradient columns
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = 0.f;
            if (r == 0 && half == 0) {
                mbar_expect_tx(&sy.tma_full, 4 * kTile);
                tma_load_3d(pt, &tmpr, &sy.tma_full, 0, pix0, b);
                tma_load_3d(h0, &tmh, &sy.tma_full, 0, pix0, b * p.S);
                tma_load_3d(h0 + kTile, &tmh, &sy.tma_full, 64, pix0, b * p.S);
                tma_load_3d(et, &tme, &sy.tma_full, 0, pix0, b * p.S);
            }
            for (int s = 0; s < p.S; ++s) {
                const int img = b * p.S + s;
                // ---- (a) dz2 = scale * g * act2'(out) -> ZW chunks [0, outc_p/8) ----
                if (half == 0) {
                    const size_t o = static_cast<size_t>(img) * p.outc * p.HW + pix;
                    float z[OP];
#pragma unroll
                    for (int c = 0; c < OP; ++c) {
                        z[c] = 0.f;
                        if (c < p.outc && valid) {
                            const float gv = __ldcs(p.g + o + static_cast<size_t>(c) * p.HW);
                            const float ov = __ldcs(p.out + o + static_cast<size_t>(c) * p.HW);
                            z[c] = gs * gv * dact_f(ov, p.act2, p.slope);
                        }
                        db2[c] += z[c];
                    }
                    if (s + 1 < p.S) {   // pull the next sample's lines towards L2 while this one is processed
#pragma unroll
                        for (int c = 0; c < OP; ++c) {
                            if (c < p.outc && valid && lane == 0) {
                                const size_t o2 = o + static_cast<size_t>(p.outc) * p.HW + static_cast<size_t>(c) * p.HW;
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.g + o2));
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.out + o2));
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < OP / 8; ++q) {
                        *reinterpret_cast<uint4*>(zw + swz(r, q)) =
                                make_uint4(pack_h2(z[8 * q], z[8 * q + 1], dt), pack_h2(z[8 * q + 2], z[8 * q + 3], dt),
                                           pack_h2(z[8 * q + 4], z[8 * q + 5], dt), pack_h2(z[8 * q + 6], z[8 * q + 7], dt));
                    }
                }
                mbar_wait(&sy.tma_full, tma_ph);   // the request below promises the issuer that the tiles are there
                tma_ph ^= 1;
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(&sy.req);
                // ---- (c) dz1 = dH * act1'(h), in place over the h tiles ----
                mbar_wait(&sy.done, done_ph);
                done_ph ^= 1;
                tc_fence_after();
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    uint32_t v[2][16];
                    tmem_ld16(tlane + half * 64 + pass * 32, v[0]);
                    tmem_ld16(tlane + half * 64 + pass * 32 + 16, v[1]);
                    tmem_ld_wait16(v[0]);
                    tmem_ld_wait16(v[1]);
                    uint8_t* tile = h0 + half * kTile;
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const int ch = pass * 4 + c4;
                        uint4* ptr = reinterpret_cast<uint4*>(tile + swz(r, ch));
                        const uint4 hv = *ptr;
                        const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
                        uint32_t o[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int col = (c4 & 1) * 8 + 2 * i;
                            const float a = __uint_as_float(v[c4 >> 1][col]) * dact(hw[i] & 0xFFFFu, p.act1, p.slope);
                            const float c2 = __uint_as_float(v[c4 >> 1][col + 1]) * dact(hw[i] >> 16, p.act1, p.slope);
                            o[i] = pack_h2(a, c2, dt);
                        }
                        *ptr = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(&sy.req);
                // ---- (e) dBoth: [0,64) -> d_emb (this sample), [64,128) -> running sum over spp ----
                mbar_wait(&sy.done, done_ph);
                done_ph ^= 1;
                tc_fence_after();
                if (r == 0 && half == 0 && s + 1 < p.S) {   // the h / emb tiles are free: next sample's
                    mbar_expect_tx(&sy.tma_full, 3 * kTile);
                    tma_load_3d(h0, &tmh, &sy.tma_full, 0, pix0, img + 1);
                    tma_load_3d(h0 + kTile, &tmh, &sy.tma_full, 64, pix0, img + 1);
                    tma_load_3d(et, &tme, &sy.tma_full, 0, pix0, img + 1);
                }
                {
                    // d_emb: this thread's 32 of the row's 64 channels
                    uint32_t v[2][16];
                    tmem_ld16(tlane + half * 32, v[0]);
                    tmem_ld16(tlane + half * 32 + 16, v[1]);
                    tmem_ld_wait16(v[0]);
                    tmem_ld_wait16(v[1]);
                    if (valid) {
                        uint4* dst = reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.d_emb) +
                                                              (static_cast<size_t>(img) * p.HW + pix) * 128) + half * 4;
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            const uint32_t* q = v[c4 >> 1] + (c4 & 1) * 8;
                            dst[c4] = make_uint4(pack_h2(__uint_as_float(q[0]), __uint_as_float(q[1]), dt),
                                                 pack_h2(__uint_as_float(q[2]), __uint_as_float(q[3]), dt),
                                                 pack_h2(__uint_as_float(q[4]), __uint_as_float(q[5]), dt),
                                                 pack_h2(__uint_as_float(q[6]), __uint_as_float(q[7]), dt));
                        }
                    }
                    // prop half of dBoth: summed over the samples (networks.py:39 repeats prop for every sample)
                    tmem_ld16(tlane + 64 + half * 32, v[0]);
                    tmem_ld16(tlane + 64 + half * 32 + 16, v[1]);
                    tmem_ld_wait16(v[0]);
                    tmem_ld_wait16(v[1]);
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] += __uint_as_float(v[i >> 4][i & 15]);
                }
                tc_fence_before();
            }
            if (valid) {
                uint4* dst = reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.d_prop) + (static_cast<size_t>(b) * p.HW + pix) * 128) +
                             half * 4;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4)
                    dst[c4] = make_uint4(pack_h2(acc[8 * c4], acc[8 * c4 + 1], dt), pack_h2(acc[8 * c4 + 2], acc[8 * c4 + 3], dt),
                                         pack_h2(acc[8 * c4 + 4], acc[8 * c4 + 5], dt), pack_h2(acc[8 * c4 + 6], acc[8 * c4 + 7], dt));
            }
        }
        // db2: warp sums of the half-0 warps (4 per group) -> shared
        if (half == 0) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float v = c < OP ? warp_sum(db2[c < OP ? c : 0]) : 0.f;
                if (lane == 0) db2_smem[(grp * 4 + (warp & 3)) * 32 + c] = v;
            }
        }
    }
    // ---------------------------------------------- slab ----------------------------------------------
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    float* slab = p.partial + static_cast<size_t>(blockIdx.x) * kFinSlab;
    const bool any = blockIdx.x * 2 < p.items;   // a CTA without items never initialised its accumulators
    if (tid < 128) {
        // lanes = cout of layer 1 (row tid), columns [256, 384) = cin
        const uint32_t tl = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + kColW1;
        for (int cc = 0; cc < 8; ++cc) {
            uint32_t v[16];
            tmem_ld16(tl + cc * 16, v);
            tmem_ld_wait16(v);
            float4* o = reinterpret_cast<float4*>(slab + tid * 128 + cc * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                o[i] = any ? make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                         __uint_as_float(v[4 * i + 3]))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else if (tid >= 256 && tid < 384) {
        const int row = tid & 127;
        const uint32_t tl = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        uint32_t v[16];
        tmem_ld16(tl + kColB1, v);
        tmem_ld_wait16(v);
        slab[128 * 128 + row] = any ? __uint_as_float(v[0]) : 0.f;
        // dW2^T: lane = h channel (row), column c = output channel -> dW2[c][row]
        for (int cc = 0; cc < 2; ++cc) {
            tmem_ld16(tl + kColW2 + cc * 16, v);
            tmem_ld_wait16(v);
#pragma unroll
            for (int i = 0; i < 16; ++i)
                slab[128 * 128 + 128 + (cc * 16 + i) * 128 + row] = (any && cc * 16 + i < p.outc_p) ? __uint_as_float(v[i]) : 0.f;
        }
        if (row < 32) {
            float sum = 0.f;
            for (int w8 = 0; w8 < 8; ++w8) sum += db2_smem[w8 * 32 + row];
            slab[128 * 128 + 128 + 32 * 128 + row] = sum;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tmem, 512);
}

// ================================================================================================================
// K9: embedding MLP backward
// ================================================================================================================
struct EmbedBwdParams {
    const void* d_red;     // (B*HW, 64) 16-bit: gradient w.r.t. the spp mean (already loss-scaled), or null
    float inv_s;           // 1 / spp
    float* partial;        // [gridDim.x][kEmbSlab]
    const void* w3t;       // [64 cin][64 cout] data-gradient packing of layer 3
    const void* w2t;       // [64 cin][64 cout] of layer 2
    int B, S, HW, dtype, act1, act2, act3;
    float slope;
    int tiles, items;
};
constexpr int kEmbSlab = 3 * (64 * 64 + 64);   // per layer (3, 2, 1): dW [64 cout][64 cin] | db [64]
// shared: W3T W2T (8 KB each, one 16 KB slot) | ONES | per group: DE EM H2 H1 X | barriers
constexpr int kEmbGroupBytes = 5 * kTile;
constexpr int kEmbSmem = 1024 + 2 * kTile + 2 * kEmbGroupBytes + 256;

__global__ void __launch_bounds__(kBwdThreads, 1)
pathnet_embed_bwd_kernel(const __grid_constant__ CUtensorMap tmde, const __grid_constant__ CUtensorMap tmem_map,
                         const __grid_constant__ CUtensorMap tmh2, const __grid_constant__ CUtensorMap tmh1,
                         const __grid_constant__ CUtensorMap tmx, const EmbedBwdParams p) {
    pdl_wait();                // the prologue below reads packed weights / issues TMA loads
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* w3t = smem;                 // [64 rows][64] = 8 KB
    uint8_t* w2t = smem + 8192;
    uint8_t* ones = smem + kTile;
    uint8_t* gbase = smem + 2 * kTile;
    GroupSync* sync = reinterpret_cast<GroupSync*>(gbase + 2 * kEmbGroupBytes);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sync + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = tid >> 8;

    if (tid == 0) {
        tma_prefetch_desc(&tmde);
        tma_prefetch_desc(&tmem_map);
        tma_prefetch_desc(&tmh2);
        tma_prefetch_desc(&tmh1);
        tma_prefetch_desc(&tmx);
        for (int g = 0; g < 2; ++g) {
            mbar_init(&sync[g].tma_full, 1);
            mbar_init(&sync[g].req, kGroupThreads);
            mbar_init(&sync[g].done, 1);
        }
        fence_barrier_init();
    }
    {
        const uint32_t one2 = p.dtype == WCMC_F16 ? 0x3C003C00u : 0x3F803F80u;
        for (int i = tid; i < kTile / 16; i += kBwdThreads) reinterpret_cast<uint4*>(ones)[i] = make_uint4(one2, one2, one2, one2);
        load_rows_swizzled(w3t, 0, p.w3t, 64, 64, 0, tid, kBwdThreads);
        load_rows_swizzled(w2t, 0, p.w2t, 64, 64, 0, tid, kBwdThreads);
    }
    if (warp == 16) tmem_alloc(tmem_ptr, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_launch_dependents();   // after the TMEM allocation (common.cuh: PDL rules)
    const uint32_t tmem = *tmem_ptr;
    // TMEM columns: [0,64) / [64,128) data gradients of group 0 / 1; then per layer L = 3, 2, 1 a 128-column block
    // [128 + 128 i, +64) dW_L (lanes 0..63 = cout), [+64, +128) ones block (column 0 = db_L)
    const int ngroups = 2 * gridDim.x;

    if (grp == 2) {
        if (lane == 0) {
            IssuerState st;
            for (int g = 0; g < 2; ++g) {
                const int gi = blockIdx.x * 2 + g;
                const int n = gi < p.items ? (p.items - gi + ngroups - 1) / ngroups : 0;
                st.step[g] = 0; st.total[g] = n * p.S * 3; st.req_ph[g] = 0; st.tma_ph[g] = 0;
            }
            const uint32_t id_dh = make_idesc_f16(128, 64, 0, 0, p.dtype, p.dtype);
            const uint32_t id_w = make_idesc_f16(128, 64, 1, 1, p.dtype, p.dtype);
            const uint64_t ones_mn = mndesc(smem_u32(ones), kTile);
            const uint64_t w3k = kdesc(smem_u32(w3t)), w2k = kdesc(smem_u32(w2t));
            bool first[3] = {true, true, true};
            int kind[2] = {0, 0};
            long long last = clock64();
            // one request: Z tile `z` (K-major for the data gradient with weights `wk`, MN-major for the weight
            // gradient against the layer input `in_mn`), accumulators of layer block `blk`
            auto layer = [&](uint32_t dcol, uint32_t z, uint64_t in_mn, uint64_t wk, bool dgrad, int blk) {
                const uint64_t zk = kdesc(z), z_mn = mndesc(z, kTile);   // A = (Z tile, the tile after it): rows 64.. unused
                if (dgrad) {
                    umma_bf16(dcol, zk, wk, id_dh, 0u);
#pragma unroll
                    for (int k = 1; k < 4; ++k) umma_bf16(dcol, zk + 2 * k, wk + 2 * k, id_dh, 1u);
                }
                const uint32_t wcol = tmem + 128 + blk * 128;
                const uint32_t acc0 = first[blk] ? 0u : 1u;
                umma_bf16(wcol, z_mn, in_mn, id_w, acc0);
                umma_bf16(wcol + 64, z_mn, ones_mn, id_w, acc0);
#pragma unroll
                for (int j = 1; j < 8; ++j) {
                    umma_bf16(wcol, z_mn + j * 128, in_mn + j * 128, id_w, 1u);
                    umma_bf16(wcol + 64, z_mn + j * 128, ones_mn + j * 128, id_w, 1u);
                }
                first[blk] = false;
            };
            auto serve = [&](const int g) {
                const uint32_t base = smem_u32(gbase + g * kEmbGroupBytes);
                const uint32_t de = base, em = base + kTile;
                const uint32_t dcol = tmem + g * 64;
                const int k3 = kind[g];
                // request 0: Z3 in DE, input h2, weights W3;  1: Z2 in EM, input h1, weights W2;  2: Z1 in DE, input x
                if (k3 == 0) {   // (the group waited for its TMA tiles before posting the request)
                    tc_fence_after();
                    layer(dcol, de, mndesc(base + 2 * kTile, kTile), w3k, true, 0);
                } else if (k3 == 1) {
                    tc_fence_after();
                    layer(dcol, em, mndesc(base + 3 * kTile, kTile), w2k, true, 1);
                } else {
                    tc_fence_after();
                    layer(dcol, de, mndesc(base + 4 * kTile, kTile), w2k, false, 2);
                }
                umma_commit(&sync[g].done);
                kind[g] = k3 == 2 ? 0 : k3 + 1;
                ++st.step[g];
            };
            while (st.step[0] < st.total[0] || st.step[1] < st.total[1]) {
                if (st.step[0] < st.total[0] && mbar_test_wait(&sync[0].req, st.req_ph[0])) {
                    st.req_ph[0] ^= 1;
                    last = clock64();
                    serve(0);
                }
                if (st.step[1] < st.total[1] && mbar_test_wait(&sync[1].req, st.req_ph[1])) {
                    st.req_ph[1] ^= 1;
                    last = clock64();
                    serve(1);
                }
                if (clock64() - last > 8000000000LL) issuer_trap("pathnet_embed_bwd");
            }
        }
    } else {
        const int r = tid & 127;                 // row of the tile = TMEM lane
        const int half = (tid >> 7) & 1;         // this thread's 32-column half of the row
        uint8_t* base = gbase + grp * kEmbGroupBytes;
        uint8_t* de = base;
        uint8_t* em = base + kTile;
        uint8_t* h2 = base + 2 * kTile;
        uint8_t* h1 = base + 3 * kTile;
        uint8_t* xt = base + 4 * kTile;
        GroupSync& sy = sync[grp];
        const uint32_t tlane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + grp * 64;
        const int dt = p.dtype;
        uint32_t tma_ph = 0, done_ph = 0;

        // The five tiles of a sample are requested in two pieces, each as soon as its buffers are dead: h2 (last read
        // by every thread's layer-2 mask pass) and emb (holds dz2, last read by the layer-2 MMAs) once the layer-2
        // request has retired; h1, d_emb and x after the layer-1 MMAs.  Only the last piece arrives on the barrier,
        // so the phase completes when all five tiles have landed.
        auto load_piece = [&](int piece, int pix0, int img) {
            if (piece == 0) {
                mbar_add_tx(&sy.tma_full, 2 * kTile);
                tma_load_3d(h2, &tmh2, &sy.tma_full, 0, pix0, img);
                tma_load_3d(em, &tmem_map, &sy.tma_full, 0, pix0, img);
            } else {
                mbar_expect_tx(&sy.tma_full, 3 * kTile);
                tma_load_3d(h1, &tmh1, &sy.tma_full, 0, pix0, img);
                tma_load_3d(de, &tmde, &sy.tma_full, 0, pix0, img);
                tma_load_3d(xt, &tmx, &sy.tma_full, 0, pix0, img);
            }
        };
        // dz = dH * act'(saved output), read from TMEM columns [0,64) of the group, written to tile `dst` (row r);
        // `saved` is the tile that holds the layer's post-activation output
        auto mask_epilogue = [&](const uint8_t* saved, uint8_t* dst, int act) {
            uint32_t v[2][16];
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) tmem_ld16(tlane + half * 32 + cc * 16, v[cc]);
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) tmem_ld_wait16(v[cc]);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const int ch = half * 4 + c4;
                const uint4 hv = *reinterpret_cast<const uint4*>(saved + swz(r, ch));
                const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
                uint32_t o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int col = (c4 & 1) * 8 + 2 * i;
                    o[i] = pack_h2(__uint_as_float(v[c4 >> 1][col]) * dact(hw[i] & 0xFFFFu, act, p.slope),
                                   __uint_as_float(v[c4 >> 1][col + 1]) * dact(hw[i] >> 16, act, p.slope), dt);
                }
                *reinterpret_cast<uint4*>(dst + swz(r, ch)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        };

        for (int item = blockIdx.x * 2 + grp; item < p.items; item += ngroups) {
            const int b = item / p.tiles, pix0 = (item - b * p.tiles) * 128;
            const int pix = pix0 + r;
            const bool valid = pix < p.HW;
            if (r == 0 && half == 0) {
                load_piece(0, pix0, b * p.S);
                load_piece(1, pix0, b * p.S);
            }
            for (int s = 0; s < p.S; ++s) {
                const int img = b * p.S + s;
                // ---- (a) dz3 = (d_emb + d_red / S) * act3'(emb), in place over the d_emb tile ----
                mbar_wait(&sy.tma_full, tma_ph);
                tma_ph ^= 1;
                {
                    const uint4* dr = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(p.d_red) +
                                                                     (static_cast<size_t>(b) * p.HW + pix) * 128);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const int ch = half * 4 + c4;
                        uint4* ptr = reinterpret_cast<uint4*>(de + swz(r, ch));
                        const uint4 dv = *ptr;
                        const uint4 ev = *reinterpret_cast<const uint4*>(em + swz(r, ch));
                        uint4 rv = make_uint4(0, 0, 0, 0);
                        if (p.d_red != nullptr && valid) rv = __ldg(dr + ch);
                        const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w}, ew[4] = {ev.x, ev.y, ev.z, ev.w},
                                       rw[4] = {rv.x, rv.y, rv.z, rv.w};
                        uint32_t o[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 d2 = unpack_h2(dw[i], dt), r2 = unpack_h2(rw[i], dt);
                            o[i] = pack_h2((d2.x + r2.x * p.inv_s) * dact(ew[i] & 0xFFFFu, p.act3, p.slope),
                                           (d2.y + r2.y * p.inv_s) * dact(ew[i] >> 16, p.act3, p.slope), dt);
                        }
                        *ptr = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(&sy.req);
                // ---- (c) dz2 = dH2 * act2'(h2) -> EM tile (emb is dead) ----
                mbar_wait(&sy.done, done_ph);
                done_ph ^= 1;
                tc_fence_after();
                mask_epilogue(h2, em, p.act2);
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(&sy.req);
                // ---- (e) dz1 = dH1 * act1'(h1) -> DE tile (dz3 is dead) ----
                mbar_wait(&sy.done, done_ph);
                done_ph ^= 1;
                tc_fence_after();
                // the layer-2 request was posted by all 256 threads after their pass over h2 and has retired
                if (r == 0 && half == 0 && s + 1 < p.S) load_piece(0, pix0, img + 1);
                mask_epilogue(h1, de, p.act1);
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(&sy.req);
                // ---- (g) layer-1 weight gradient retired: all five tiles are free ----
                mbar_wait(&sy.done, done_ph);
                done_ph ^= 1;
                if (r == 0 && half == 0 && s + 1 < p.S) load_piece(1, pix0, img + 1);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    float* slab = p.partial + static_cast<size_t>(blockIdx.x) * kEmbSlab;
    const bool any = blockIdx.x * 2 < p.items;
    if (tid < 64) {
        // lanes 0..63 = cout; per layer block: columns [0,64) dW, column 64 = db
        const uint32_t tl = tmem + (static_cast<uint32_t>(warp * 32) << 16) + 128;
        for (int L = 0; L < 3; ++L) {
            float* o = slab + L * (64 * 64 + 64);
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t v[16];
                tmem_ld16(tl + L * 128 + cc * 16, v);
                tmem_ld_wait16(v);
#pragma unroll
                for (int i = 0; i < 16; ++i) o[tid * 64 + cc * 16 + i] = any ? __uint_as_float(v[i]) : 0.f;
            }
            uint32_t v[16];
            tmem_ld16(tl + L * 128 + 64, v);
            tmem_ld_wait16(v);
            o[64 * 64 + tid] = any ? __uint_as_float(v[0]) : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tmem, 512);
}

// Sums the per-CTA slabs in a fixed order, multiplies by *scale (1 / loss scale) and scatters into up to 8 output
// segments: segment i covers slab elements [seg.off, seg.off + rows*cols_src) read as a [rows][cols_src] matrix of
// which the first cols_dst columns go to dst (row-major [rows][cols_dst]).
struct SlabSeg {
    float* dst;
    int off, rows, cols_src, cols_dst;
};
struct SlabReduceParams {
    const float* partial;
    int nslabs, slab;
    const float* scale;
    int nseg;
    SlabSeg seg[8];
};
// 64 output elements per CTA, 4 threads per element (each sums every 4th slab with 4 loads in flight), fixed-order
// combine through shared memory.  (The first version -- 32 CTAs, one thread per element walking all 148 slabs --
// took 196 us per call: 8 k threads cannot cover 12 MB of latency-bound strided reads.)
__global__ void __launch_bounds__(256) slab_reduce_kernel(const SlabReduceParams p) {
    pdl_start();
    __shared__ float red[4][64];
    const float sc = p.scale != nullptr ? __ldg(p.scale) : 1.f;
    const int e = threadIdx.x & 63, q = threadIdx.x >> 6;
    for (int s = 0; s < p.nseg; ++s) {
        const SlabSeg& sg = p.seg[s];
        const int n = sg.rows * sg.cols_dst;
        for (int base = blockIdx.x * 64; base < n; base += gridDim.x * 64) {
            const int i = base + e;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            if (i < n) {
                const int row = i / sg.cols_dst, col = i - row * sg.cols_dst;
                const float* src = p.partial + sg.off + row * sg.cols_src + col;
                int k = q;
                for (; k + 12 < p.nslabs; k += 16) {
                    a0 += __ldg(src + static_cast<size_t>(k) * p.slab);
                    a1 += __ldg(src + static_cast<size_t>(k + 4) * p.slab);
                    a2 += __ldg(src + static_cast<size_t>(k + 8) * p.slab);
                    a3 += __ldg(src + static_cast<size_t>(k + 12) * p.slab);
                }
                for (; k < p.nslabs; k += 4) a0 += __ldg(src + static_cast<size_t>(k) * p.slab);
            }
            red[q][e] = (a0 + a1) + (a2 + a3);
            __syncthreads();
            if (q == 0 && i < n) sg.dst[i] = ((red[0][e] + red[1][e]) + (red[2][e] + red[3][e])) * sc;
            __syncthreads();
        }
    }
}

}  // namespace wcmc

using namespace wcmc;

static bool bwd_act_ok(int a) { return a == WCMC_ACT_LINEAR || a == WCMC_ACT_RELU || a == WCMC_ACT_LEAKY; }
static bool bwd_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static int bwd_act_tmap(CUtensorMap* m, const void* base, int cs, int HW, long images) {
    uint64_t dims[3] = {static_cast<uint64_t>(cs), static_cast<uint64_t>(HW), static_cast<uint64_t>(images)};
    uint64_t strides[2] = {static_cast<uint64_t>(cs) * 2, static_cast<uint64_t>(cs) * 2 * HW};
    uint32_t box[3] = {64, 128, 1};
    return wcmc_encode_tmap(m, WCMC_BF16, base, 3, dims, strides, box, 1);
}
static int bwd_grid(int items) {
    const int sms = wcmc_num_sms();
    const int want = (items + 1) / 2;
    return want < sms ? want : sms;
}

extern "C" size_t wcmc_pathnet_bwd_workspace(int which) {
    const size_t slab = which == 0 ? kFinSlab : kEmbSlab;
    return slab * sizeof(float) * static_cast<size_t>(wcmc_num_sms());
}

extern "C" int wcmc_pathnet_final_bwd(const float* g, const float* out, const float* gscale, const float* inv_scale,
                                      const void* emb, int emb_cs, const void* prop, int prop_cs, const void* hfin,
                                      const void* w1t, const void* w2t, int outc, int outc_p, int dtype, int act1, int act2,
                                      float slope, void* d_emb, void* d_prop, float* dw1, float* db1, float* dw2,
                                      float* db2, int B, int S, int HW, void* workspace, size_t workspace_bytes,
                                      void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(B > 0 && S > 0 && HW > 0, WCMC_ESHAPE, "pathnet_final_bwd: bad shape");
    WCMC_REQUIRE(outc > 0 && outc_p % 16 == 0 && outc_p >= outc && outc_p <= 32, WCMC_ESHAPE,
                 "pathnet_final_bwd: outc_p %d must be a multiple of 16 in [outc, 32]", outc_p);
    WCMC_REQUIRE(dtype == WCMC_BF16 || dtype == WCMC_F16, WCMC_ESHAPE, "pathnet_final_bwd: dtype must be 16-bit");
    WCMC_REQUIRE(bwd_act_ok(act1) && bwd_act_ok(act2), WCMC_ESHAPE, "pathnet_final_bwd: bad activation");
    WCMC_REQUIRE(g && out && emb && prop && hfin && w1t && w2t && d_emb && d_prop && dw1 && db1 && dw2 && db2 && workspace,
                 WCMC_ESHAPE, "pathnet_final_bwd: null pointer");
    WCMC_REQUIRE(emb_cs % 8 == 0 && emb_cs >= 64 && prop_cs % 8 == 0 && prop_cs >= 64, WCMC_ESHAPE,
                 "pathnet_final_bwd: channel strides");
    WCMC_REQUIRE(bwd_al16(emb) && bwd_al16(prop) && bwd_al16(hfin) && bwd_al16(w1t) && bwd_al16(w2t) && bwd_al16(d_emb) &&
                     bwd_al16(d_prop) && bwd_al16(workspace),
                 WCMC_EALIGN, "pathnet_final_bwd: pointers must be 16-byte aligned");
    const int tiles = (HW + 127) / 128;
    const int items = B * tiles;
    const int grid = bwd_grid(items);
    WCMC_REQUIRE(workspace_bytes >= static_cast<size_t>(grid) * kFinSlab * sizeof(float), WCMC_EWORKSPACE,
                 "pathnet_final_bwd: workspace too small");
    FinalBwdParams p{g, out, gscale, d_emb, d_prop, static_cast<float*>(workspace), w1t, w2t, B, S, HW, outc, outc_p,
                     dtype, act1, act2, slope, tiles, items};
    CUtensorMap tmh, tme, tmpr;
    int rc = bwd_act_tmap(&tmh, hfin, 128, HW, static_cast<long>(B) * S);
    if (rc) return rc;
    if ((rc = bwd_act_tmap(&tme, emb, emb_cs, HW, static_cast<long>(B) * S))) return rc;
    if ((rc = bwd_act_tmap(&tmpr, prop, prop_cs, HW, B))) return rc;
    WCMC_FUNC_SMEM(pathnet_final_bwd_kernel<16>, kFinSmem);
    WCMC_FUNC_SMEM(pathnet_final_bwd_kernel<32>, kFinSmem);
    if (outc_p == 16) WCMC_LAUNCH((pathnet_final_bwd_kernel<16>), grid, kBwdThreads, kFinSmem, stream, tmh, tme, tmpr, p);
    else WCMC_LAUNCH((pathnet_final_bwd_kernel<32>), grid, kBwdThreads, kFinSmem, stream, tmh, tme, tmpr, p);
    WCMC_LAUNCH_CHECK();
    SlabReduceParams r;
    r.partial = p.partial; r.nslabs = grid; r.slab = kFinSlab; r.scale = inv_scale; r.nseg = 4;
    r.seg[0] = SlabSeg{dw1, 0, 128, 128, 128};
    r.seg[1] = SlabSeg{db1, 128 * 128, 1, 128, 128};
    r.seg[2] = SlabSeg{dw2, 128 * 128 + 128, outc, 128, 128};
    r.seg[3] = SlabSeg{db2, 128 * 128 + 128 + 32 * 128, 1, 32, outc};
    WCMC_LAUNCH(slab_reduce_kernel, 296, 256, 0, stream, r);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" int wcmc_pathnet_embed_bwd(const void* d_emb, const void* d_red, const float* inv_scale, const void* emb,
                                      int emb_cs, const void* h2, const void* h1, const void* x16, const void* w3t,
                                      const void* w2t, int cin, int dtype, int act1, int act2, int act3, float slope,
                                      float* dw3, float* db3, float* dw2, float* db2, float* dw1, float* db1, int B,
                                      int S, int HW, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(B > 0 && S > 0 && HW > 0 && cin > 0 && cin <= 64, WCMC_ESHAPE, "pathnet_embed_bwd: bad shape");
    WCMC_REQUIRE(dtype == WCMC_BF16 || dtype == WCMC_F16, WCMC_ESHAPE, "pathnet_embed_bwd: dtype must be 16-bit");
    WCMC_REQUIRE(bwd_act_ok(act1) && bwd_act_ok(act2) && bwd_act_ok(act3), WCMC_ESHAPE, "pathnet_embed_bwd: bad activation");
    WCMC_REQUIRE(d_emb && emb && h2 && h1 && x16 && w3t && w2t && dw3 && db3 && dw2 && db2 && dw1 && db1 && workspace,
                 WCMC_ESHAPE, "pathnet_embed_bwd: null pointer");
    WCMC_REQUIRE(emb_cs % 8 == 0 && emb_cs >= 64, WCMC_ESHAPE, "pathnet_embed_bwd: emb channel stride");
    WCMC_REQUIRE(bwd_al16(d_emb) && bwd_al16(d_red) && bwd_al16(emb) && bwd_al16(h2) && bwd_al16(h1) && bwd_al16(x16) &&
                     bwd_al16(w3t) && bwd_al16(w2t) && bwd_al16(workspace),
                 WCMC_EALIGN, "pathnet_embed_bwd: pointers must be 16-byte aligned");
    const int tiles = (HW + 127) / 128;
    const int items = B * tiles;
    const int grid = bwd_grid(items);
    WCMC_REQUIRE(workspace_bytes >= static_cast<size_t>(grid) * kEmbSlab * sizeof(float), WCMC_EWORKSPACE,
                 "pathnet_embed_bwd: workspace too small");
    EmbedBwdParams p{d_red, 1.f / S, static_cast<float*>(workspace), w3t, w2t, B, S, HW, dtype, act1, act2, act3, slope,
                     tiles, items};
    CUtensorMap tmde, tmem_map, tmh2, tmh1, tmx;
    const long images = static_cast<long>(B) * S;
    int rc = bwd_act_tmap(&tmde, d_emb, 64, HW, images);
    if (rc) return rc;
    if ((rc = bwd_act_tmap(&tmem_map, emb, emb_cs, HW, images))) return rc;
    if ((rc = bwd_act_tmap(&tmh2, h2, 64, HW, images))) return rc;
    if ((rc = bwd_act_tmap(&tmh1, h1, 64, HW, images))) return rc;
    if ((rc = bwd_act_tmap(&tmx, x16, 64, HW, images))) return rc;
    WCMC_FUNC_SMEM(pathnet_embed_bwd_kernel, kEmbSmem);
    WCMC_LAUNCH(pathnet_embed_bwd_kernel, grid, kBwdThreads, kEmbSmem, stream, tmde, tmem_map, tmh2, tmh1, tmx, p);
    WCMC_LAUNCH_CHECK();
    SlabReduceParams r;
    r.partial = p.partial; r.nslabs = grid; r.slab = kEmbSlab; r.scale = inv_scale; r.nseg = 6;
    const int blk = 64 * 64 + 64;
    r.seg[0] = SlabSeg{dw3, 0, 64, 64, 64};
    r.seg[1] = SlabSeg{db3, 64 * 64, 1, 64, 64};
    r.seg[2] = SlabSeg{dw2, blk, 64, 64, 64};
    r.seg[3] = SlabSeg{db2, blk + 64 * 64, 1, 64, 64};
    r.seg[4] = SlabSeg{dw1, 2 * blk, 64, 64, cin};
    r.seg[5] = SlabSeg{db1, 2 * blk + 64 * 64, 1, 64, 64};
    WCMC_LAUNCH(slab_reduce_kernel, 296, 256, 0, stream, r);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}
