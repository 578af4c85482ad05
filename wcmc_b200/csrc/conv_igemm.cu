// K1/K2: halo-resident implicit-GEMM convolution on tcgen05 (forward and data-gradient).
//
// Replaces the cuDNN calls behind `nn.Conv2d` in sbmc.modules.ConvChain (the un-vendored
// dependency of /root/reference/support/networks.py:18-24 and train_kpcn.py:213); see
// SURVEY.md Appendix A.1/A.2 for the layer shapes.
//
//   y[n, oy, ox, co] = act( bias[co] + sum_{ky,kx,ci} x[n, oy+ky-pad, ox+kx-pad, ci] * w[co, ky, kx, ci] )
//
// Data layout: activations NHWC bf16 with the channel count padded to a multiple of 16;
// weights packed as bf16 [cout_p][k*k][cin_p].  The data gradient is the same kernel run on
// dY with pad' = k-1-pad and tap-flipped / transposed weights, with the ReLU mask of the
// layer input fused in the epilogue.
//
// Mapping to the GEMM:  M = 128 output pixels (an 8-wide x 16-tall box), N = cout tile
// (<= 128), K = cin per tap.  A CTA owns a region of `mt` M-tiles side by side.  For every
// 64-channel chunk the (16+k-1) x (8*mt+k-1) input halo is brought into shared memory ONCE
// by a single 4-D TMA box (128-byte swizzle, zero fill outside the image = conv padding).
// Every tap's A operand is then just a shifted window of that halo: the UMMA shared-memory
// descriptor starts at pixel (ky, kx+8t) of the halo and uses the halo row pitch as its
// stride-byte-offset, so the k*k-fold re-read of the activations never leaves the SM.
// Weights stream through a TMA ring whose stages carry one tap or a whole kernel row of taps (TPS).  Accumulators
// live in TMEM (double buffered: the epilogue of one region overlaps the MMAs of the next).
//
// Warp roles (352 threads): w0 halo producer | w1 weight producer | w2 TMEM alloc + MMA issuer | w3..w10 epilogue
// (TMEM -> regs -> bias / act / mask / column sums -> global).
//
// PAIR = true: the kernel runs as clusters of two CTAs on tcgen05 cta_group::2 (M = 256) -- see decode_item below and
// DESIGN.md section 5 for why (the shared-memory operand feed, not the tensor pipe, bounds the single-CTA kernel).
#include <algorithm>

#include "common.cuh"

namespace wcmc {

constexpr int kMaxPlaneSlots = 4;
constexpr int kMaxBStages = 12;
constexpr int kBarBytes = 1024;      // mbarriers + TMEM pointer live in the first KB
constexpr int kConvThreads = 96 + 256;  // 3 single-role warps + 8 epilogue warps
constexpr int kConvSmemMax = 232448 - 1024;  // opt-in limit minus the slack used to 1024-align the base

struct ConvParams {
    int N, Ho, Wo;
    int ksize, pad;
    int cin_p, nch;
    int cout_p, nt, n_tiles;
    int mt, regions_x, regions_y, total_items;
    int total_regions, pair_items;   // CTA-pair mode: regions over the batch; (region pair, n tile) work items
    int halo_w, halo_h;
    int plane_stride, b_stride, b_stages, plane_slots;  // shared-memory carve-up
    int tap_stride;                                      // bytes between the taps of one weight stage
    int b_resident;   // 1: the weight ring holds ALL taps / chunks of the (single) n tile: loaded once per CTA, then reused
    void* out;
    int out_cs, out_coff, out_dtype;
    int x_dtype, w_dtype;
    const float* bias;
    int act;
    const __nv_bfloat16* mask;
    int mask_cs, mask_coff;
    float slope;
    int flags;
    float* colsum;             // optional: colsum[ch] += colsum_scale * sum over valid pixels of the output
    const float* colsum_scale;  // device scalar or null
    // fused softmax + kernel-apply epilogue (KA instantiations): the logits never reach HBM
    const float* ka_data;      // (N, 3, Ho, Wo) fp32: the radiance buffer the predicted kernels filter
    float* ka_out;             // (N, 3, Ho, Wo) fp32
    int ka_k, ka_taps;         // 21, 441
    int ka_off;                // byte offset of the staging area (data halo + merge buffer) in shared memory
};

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v > 0.f ? v : v * slope;
    return v;
}

// One tap = MT x NK tcgen05.mma (compile-time counts: no branches between the MMAs, descriptor
// deltas are immediates).  Executed by the single elected lane.
// PAIR = true issues tcgen05.mma.cta_group::2 (M = 256: this CTA's 128 pixels and the peer's).
template <bool PAIR>
__device__ __forceinline__ void umma_any(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t accum) {
    if (PAIR) umma_bf16_pair(d, ad, bd, idesc, accum);
    else umma_bf16(d, ad, bd, idesc, accum);
}
template <int MT, int NK, bool PAIR>
__device__ __forceinline__ void issue_tap(uint32_t d_base, uint64_t ad, uint64_t bd, uint32_t idesc,
                                          uint32_t accum) {
#pragma unroll
    for (int t = 0; t < MT; ++t) {
#pragma unroll
        for (int j = 0; j < NK; ++j)
            umma_any<PAIR>(d_base + t * 128, ad + (64 * t + 2 * j), bd + 2 * j, idesc, j > 0 ? 1u : accum);
    }
}

// MMA-issue state that lives across chunks / regions (all in registers of the issuing thread).
struct BRing {
    uint64_t* full;
    uint64_t* empty;
    uint32_t base_lo;   // (address of stage 0) >> 4
    uint32_t stride_lo; // stage stride >> 4
    uint32_t cur_lo;    // (address of the current stage) >> 4
    uint32_t tap_lo;    // stride between the taps of one stage >> 4
    int stages, idx, phase;
    bool loaded;   // resident weights: every stage has been waited for once, later items skip the barrier
};

// All taps of one 64-channel chunk.  A weight stage carries TPS taps (1, or a whole kernel row: TPS = ksize), so
// the issuing thread pays one mbarrier wait and one tcgen05.commit per TPS*MT*NK MMAs -- with one tap per stage
// that fixed cost (~200 clocks of uniform-register traffic around the barrier) is serialised with the 8 MMAs of
// a tap often enough to matter (tools/mma_probe.cu: the issue rate, not the tensor pipe, sets the pace).
// MT / NK / TPS are compile-time, the body is branch-free and the descriptor deltas are immediates.
template <int MT, int NK, bool PAIR, int TPS>
__device__ __forceinline__ void mma_chunk(BRing& br, int taps, int ksize, int row_step, uint32_t a_lo,
                                          uint32_t a_hi, uint32_t b_hi, uint32_t lo_fixed, uint32_t d_base,
                                          uint32_t idesc, uint32_t& accum) {
    int kx = 0;
    for (int tap = 0; tap < taps; tap += TPS) {
        if (!br.loaded) {
            mbar_wait(&br.full[br.idx], br.phase);
            tc_fence_after();
        }
        if (elect_one()) {
#pragma unroll
            for (int u = 0; u < TPS; ++u)
                issue_tap<MT, NK, PAIR>(d_base, (static_cast<uint64_t>(a_hi) << 32) | (lo_fixed | (a_lo + 8 * u)),
                                        (static_cast<uint64_t>(b_hi) << 32) | (lo_fixed | (br.cur_lo + u * br.tap_lo)),
                                        idesc, u > 0 ? 1u : accum);
            if (PAIR) umma_commit_pair(&br.empty[br.idx], 3);
            else umma_commit(&br.empty[br.idx]);
        }
        __syncwarp();
        accum = 1;
        br.cur_lo += br.stride_lo;
        if (++br.idx == br.stages) { br.idx = 0; br.phase ^= 1; br.cur_lo = br.base_lo; }
        a_lo += 8 * TPS;                                // next taps: TPS halo pixels (128 B each) to the right ...
        kx += TPS;
        if (kx == ksize) { kx = 0; a_lo += row_step; }  // ... or down to the next halo row
    }
}

// Work decomposition.  Single CTA: item -> (region, n tile).  CTA pair (cluster of 2, cta_group::2): the two
// CTAs take two DIFFERENT regions (2*rp, 2*rp + 1) of the SAME n tile -- each loads its own halo into its own
// shared memory, and one M = 256 MMA of the leader covers both (the A descriptor addresses the same offsets in
// both CTAs); each CTA holds only half of the weight tile (N/2 rows), so the B operand traffic per CTA and the
// shared-memory reads per MMA drop from 4 KB + nt*32 B to 4 KB + nt*16 B.  An odd region count leaves the last
// pair with a dead second region (computed on the clamped region, never stored).
struct Item {
    int n, rx, ry, n0;
    bool live;
};
template <bool PAIR>
__device__ __forceinline__ Item decode_item(const ConvParams& p, int item, int rank) {
    Item it;
    int r = item / p.n_tiles;
    it.n0 = (item - r * p.n_tiles) * p.nt;
    it.live = true;
    if (PAIR) {
        r = 2 * r + rank;
        if (r >= p.total_regions) { r = p.total_regions - 1; it.live = false; }
    }
    it.rx = r % p.regions_x;
    it.ry = (r / p.regions_x) % p.regions_y;
    it.n = r / (p.regions_x * p.regions_y);
    return it;
}

// k-th work item of a CTA.  Plain launches stride the items over the grid.  KA launches keep ALL n tiles of a region
// on one CTA, back to back: the fused epilogue carries a running softmax over the 441 logits of its pixels across
// the four 112-column tiles (the flash-attention recurrence), so the tiles of a region must arrive in order.
template <bool KA>
__device__ __forceinline__ int item_at(const ConvParams& p, int item0, int item_step, int k) {
    if (!KA) return item0 + k * item_step;
    const int g = k / p.n_tiles;
    return (item0 + g * item_step) * p.n_tiles + (k - g * p.n_tiles);
}

constexpr int kKaHaloRows = 16 + 20;     // region rows + the 21x21 footprint

template <bool PAIR, int TPS, bool KA = false>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmw,
                  const ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint8_t* planes = smem + kBarBytes;
    const int kPlaneSlots = p.plane_slots;
    uint8_t* bst = planes + kPlaneSlots * p.plane_stride;
    const int kBStages = p.b_stages;
    uint64_t* plane_full = bars;
    uint64_t* plane_empty = bars + kMaxPlaneSlots;
    uint64_t* b_full = bars + 2 * kMaxPlaneSlots;
    uint64_t* b_empty = b_full + kMaxBStages;
    uint64_t* acc_full = b_empty + kMaxBStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // pair mode: cluster rank 0 is the leader (issues the MMAs, owns the full / acc_empty barriers)
    const int rank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
    const int item0 = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int item_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const int n_items = PAIR ? p.pair_items : p.total_items;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmx);
        tma_prefetch_desc(&tmw);
        for (int i = 0; i < kPlaneSlots; ++i) {
            mbar_init(&plane_full[i], 1);
            mbar_init(&plane_empty[i], 1);
        }
        for (int i = 0; i < kBStages; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], PAIR ? 512 : 256);   // the epilogue threads of both CTAs release a buffer
        }
        fence_barrier_init();
    }
    __syncwarp();
    if (warp == 2) {
        if (PAIR) tmem_alloc_pair(tmem_ptr, 512);
        else tmem_alloc(tmem_ptr, 512);
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();   // the peer's barriers and TMEM exist before anything is signalled to it
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // PDL: barriers, tensor memory and descriptors are set up in the shadow of the previous kernel's tail; the
    // dependents may be scheduled now that this CTA owns its TMEM columns; nothing below runs before the previous
    // grid has completed
    pdl_launch_dependents();
    pdl_wait();

    const int taps = p.ksize * p.ksize;
    const int region_w = 8 * p.mt;
    const uint32_t plane_bytes = static_cast<uint32_t>(p.halo_w * p.halo_h * 128);
    const uint32_t b_bytes = static_cast<uint32_t>(p.nt * 128);   // whole n tile (a pair CTA loads half of it)

    if (warp == 0) {
        // ---------------- halo producer ----------------
        if (lane == 0) {
            int ps = 0, ph = 0, loaded = 0;
            const bool dry = (p.flags & (1 << 17)) != 0;   // tuning knob: planes loaded once (wrong results)
            for (int k = 0, item; (item = item_at<KA>(p, item0, item_step, k)) < n_items; ++k) {
                const Item w = decode_item<PAIR>(p, item, rank);
                for (int c = 0; c < p.nch; ++c) {
                    mbar_wait(&plane_empty[ps], ph ^ 1);
                    if (dry && loaded >= kPlaneSlots) {
                        if (rank == 0) mbar_arrive(&plane_full[ps]);
                    } else if (PAIR) {
                        // both halos are signalled on the leader's barrier, which expects the two of them
                        if (rank == 0) mbar_expect_tx(&plane_full[ps], 2 * plane_bytes);
                        tma_load_4d_pair(planes + ps * p.plane_stride, &tmx, mapa_u32(smem_u32(&plane_full[ps]), 0),
                                         c * 64, w.rx * region_w - p.pad, w.ry * 16 - p.pad, w.n);
                        ++loaded;
                    } else {
                        mbar_expect_tx(&plane_full[ps], plane_bytes);
                        tma_load_4d(planes + ps * p.plane_stride, &tmx, &plane_full[ps], c * 64,
                                    w.rx * region_w - p.pad, w.ry * 16 - p.pad, w.n);
                        ++loaded;
                    }
                    if (++ps == kPlaneSlots) { ps = 0; ph ^= 1; }
                }
            }
            // tail: every slot released by the MMAs (a pair CTA must not exit while commits can still arrive)
            for (int i = 0; i < kPlaneSlots; ++i) {
                mbar_wait(&plane_empty[ps], ph ^ 1);
                if (++ps == kPlaneSlots) { ps = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ---------------- weight producer ----------------
        if (lane == 0) {
            int bs = 0, ph = 0, loaded = 0;
            const bool dry = (p.flags & (1 << 16)) != 0;   // tuning knob: weight ring filled once (wrong results)
            for (int k = 0, item; (item = item_at<KA>(p, item0, item_step, k)) < n_items; ++k) {
                if (p.b_resident && k > 0) break;      // the ring already holds every tap of every chunk
                const int n0 = (item % p.n_tiles) * p.nt;
                for (int c = 0; c < p.nch; ++c) {
                    for (int tap = 0; tap < taps; tap += TPS) {   // one stage = TPS taps, one box each
                        mbar_wait(&b_empty[bs], ph ^ 1);
                        uint8_t* dst = bst + bs * p.b_stride;
                        if (dry && loaded >= kBStages) {
                            if (rank == 0) mbar_arrive(&b_full[bs]);
                        } else if (PAIR) {
                            // this CTA's half of the n tile (the box of tmw is nt/2 rows tall in pair mode)
                            if (rank == 0) mbar_expect_tx(&b_full[bs], TPS * b_bytes);
                            const uint32_t bar = mapa_u32(smem_u32(&b_full[bs]), 0);
#pragma unroll
                            for (int u = 0; u < TPS; ++u)
                                tma_load_3d_pair(dst + u * p.tap_stride, &tmw, bar, c * 64, tap + u,
                                                 n0 + rank * (p.nt >> 1));
                            ++loaded;
                        } else {
                            mbar_expect_tx(&b_full[bs], TPS * b_bytes);
#pragma unroll
                            for (int u = 0; u < TPS; ++u)
                                tma_load_3d(dst + u * p.tap_stride, &tmw, &b_full[bs], c * 64, tap + u, n0);
                            ++loaded;
                        }
                        if (++bs == kBStages) { bs = 0; ph ^= 1; }
                    }
                }
            }
            for (int i = 0; i < kBStages && !p.b_resident; ++i) {   // tail, as for the halos
                mbar_wait(&b_empty[bs], ph ^ 1);
                if (++bs == kBStages) { bs = 0; ph ^= 1; }
            }
        }
    } else if (warp == 2) {
        // ---------------- MMA issuer ----------------
        // The whole warp runs the warp-uniform control flow (so descriptors stay in uniform registers
        // and the tcgen05.mma of a tap issue back to back); one elected lane issues them.  Descriptors
        // are built once per plane / stage; per MMA only the 14-bit start-address field moves, by
        // compile-time constants (issue_tap).  The (mt, nk) dispatch happens once per chunk.
        if (rank == 0) {
            const uint32_t idesc = make_idesc_f16(PAIR ? 256 : 128, p.nt, 0, 0, p.x_dtype, p.w_dtype);
            const uint32_t sbo = static_cast<uint32_t>(p.halo_w * 128);
            const uint32_t a_hi = static_cast<uint32_t>(make_sdesc_sw128(0, 16, sbo, 0) >> 32);
            const uint32_t b_hi = static_cast<uint32_t>(make_sdesc_sw128(0, 16, 1024, 0) >> 32);
            const uint32_t lo_fixed = 1u << 16;  // LBO field (16 B >> 4) sits in the low word
            const int row_step = (p.halo_w - p.ksize) * 8;
            const int ksize = p.ksize, nch = p.nch, cin_p = p.cin_p, mt = p.mt;
            const uint32_t plane0_lo = smem_u32(planes) >> 4;
            const uint32_t plane_stride_lo = static_cast<uint32_t>(p.plane_stride) >> 4;
            BRing br;
            br.full = b_full; br.empty = b_empty;
            br.base_lo = smem_u32(bst) >> 4; br.stride_lo = static_cast<uint32_t>(p.b_stride) >> 4;
            br.cur_lo = br.base_lo; br.tap_lo = static_cast<uint32_t>(p.tap_stride) >> 4;
            br.stages = p.b_stages; br.idx = 0; br.phase = 0; br.loaded = false;
            int ps = 0, pph = 0;
            for (int it = 0; item_at<KA>(p, item0, item_step, it) < n_items; ++it) {
                const int buf = it & 1;
                if (PAIR) mbar_wait_cluster(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
                else mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_base = tmem_base + buf * 256;
                uint32_t accum = 0;
                for (int c = 0; c < nch; ++c) {
                    mbar_wait(&plane_full[ps], pph);
                    tc_fence_after();
                    const uint32_t a_lo = plane0_lo + ps * plane_stride_lo;
                    int nk = (cin_p - c * 64) >> 4;
                    if (nk > 4) nk = 4;
#define WCMC_CHUNK(MT, NK) \
    mma_chunk<MT, NK, PAIR, TPS>(br, taps, ksize, row_step, a_lo, a_hi, b_hi, lo_fixed, d_base, idesc, accum)
                    switch (mt * 8 + nk) {
                        case 8 + 1: WCMC_CHUNK(1, 1); break;
                        case 8 + 2: WCMC_CHUNK(1, 2); break;
                        case 8 + 3: WCMC_CHUNK(1, 3); break;
                        case 8 + 4: WCMC_CHUNK(1, 4); break;
                        case 16 + 1: WCMC_CHUNK(2, 1); break;
                        case 16 + 2: WCMC_CHUNK(2, 2); break;
                        case 16 + 3: WCMC_CHUNK(2, 3); break;
                        default: WCMC_CHUNK(2, 4); break;
                    }
#undef WCMC_CHUNK
                    if (elect_one()) {
                        if (PAIR) umma_commit_pair(&plane_empty[ps], 3);
                        else umma_commit(&plane_empty[ps]);
                    }
                    __syncwarp();
                    if (++ps == kPlaneSlots) { ps = 0; pph ^= 1; }
                }
                if (elect_one()) {
                    if (PAIR) umma_commit_pair(&acc_full[buf], 3);
                    else umma_commit(&acc_full[buf]);
                }
                __syncwarp();
                if (p.b_resident) br.loaded = true;     // one item has walked the whole (non-recycling) ring
            }
        }
    } else {
        // ---------------- epilogue (8 warps) ----------------
        // warp w: TMEM lane quarter q = w & 3 (hardware rule), half h = (w - 3) >> 2.  With two M tiles
        // each half owns one tile; with one tile the halves interleave 16-column chunks.  TMEM loads are
        // software pipelined (the load of chunk i+1 is in flight while chunk i is processed).
        const int q = warp & 3;
        const int half = (warp - 3) >> 2;
        const int m = q * 32 + lane;
        const int ty = m >> 3, tx = m & 7;
        const int cc0 = (p.mt == 2) ? 0 : half;
        const int ccs = (p.mt == 2) ? 1 : 2;
        const int t = (p.mt == 2) ? half : 0;
        const uint32_t acc_empty_leader = PAIR ? mapa_u32(smem_u32(&acc_empty[0]), 0) : 0u;
        // ---- KA: running softmax + weighted gather state of this thread's pixel (carried over the n tiles of a region) ----
        float4* ka_halo = reinterpret_cast<float4*>(smem + p.ka_off);          // [36][region_w + 20] (c0, c1, c2, -)
        const int ka_pitch = region_w + 20;
        float* ka_merge = reinterpret_cast<float*>(smem + p.ka_off) + 4 * kKaHaloRows * ka_pitch;   // [128][5]
        float ka_m = -INFINITY, ka_l = 0.f, ka_a0 = 0.f, ka_a1 = 0.f, ka_a2 = 0.f;
        const int epi_tid = threadIdx.x - 96;      // 0..255 among the epilogue warps
        // The bias of the item's n tile is staged in shared memory (second half of the barrier KB) while the item's
        // MMAs still run: fetched per 16-column chunk from global memory it put one L2 round trip (~700 clocks) into
        // the dependent chain of every chunk -- 3-5 k clocks per item, the "fixed per-item cost" of the small layers.
        float* bias_s = reinterpret_cast<float*>(smem + 512);
        for (int it = 0, item; (item = item_at<KA>(p, item0, item_step, it)) < n_items; ++it) {
            const int buf = it & 1;
            const Item w = decode_item<PAIR>(p, item, rank);
            const int n0 = w.n0, rx = w.rx, ry = w.ry, n = w.n;
            int ncc = (p.cout_p - n0) >> 4;
            if (ncc > (p.nt >> 4)) ncc = p.nt >> 4;
            const int oy = ry * 16 + ty;
            const int ox = rx * region_w + 8 * t + tx;
            const bool valid = w.live && (oy < p.Ho) && (ox < p.Wo);
            const size_t pix = (static_cast<size_t>(n) * p.Ho + oy) * p.Wo + ox;
            if (p.bias != nullptr) {
                asm volatile("bar.sync 3, 256;" ::: "memory");      // nobody still reads the previous item's bias
                if (epi_tid < p.nt) bias_s[epi_tid] = (n0 + epi_tid < p.cout_p) ? __ldg(p.bias + n0 + epi_tid) : 0.f;
                asm volatile("bar.sync 3, 256;" ::: "memory");
            }
            if (p.mask != nullptr && valid) {
                // pull this pixel's mask row (the dgrad's ReLU mask, 2 bytes per channel) towards L1 while the MMAs run
                const __nv_bfloat16* mrow = p.mask + pix * p.mask_cs + p.mask_coff + n0;
                for (int cc = cc0; cc < ncc; cc += ccs)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(mrow + cc * 16));
            }
            mbar_wait(&acc_full[buf], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 256 + t * 128;

            if (KA && n0 == 0) {
                // first n tile of a region: stage the (16+20) x (region_w+20) neighbourhood of the radiance buffer
                // (zero outside the image, like the Halide op's constant_exterior) and reset the running softmax
                asm volatile("bar.sync 1, 256;" ::: "memory");      // everyone is done with the previous region's halo
                const int gy0 = ry * 16 - (p.ka_k >> 1), gx0 = rx * region_w - (p.ka_k >> 1);
                const size_t plane = static_cast<size_t>(p.Ho) * p.Wo;
                const float* dn = p.ka_data + static_cast<size_t>(n) * 3 * plane;
                for (int i = epi_tid; i < kKaHaloRows * ka_pitch; i += 256) {
                    const int hy = i / ka_pitch, hx = i - hy * ka_pitch;
                    const int gy = gy0 + hy, gx = gx0 + hx;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (gy >= 0 && gy < p.Ho && gx >= 0 && gx < p.Wo) {
                        const float* d = dn + static_cast<size_t>(gy) * p.Wo + gx;
                        v = make_float4(__ldg(d), __ldg(d + plane), __ldg(d + 2 * plane), 0.f);
                    }
                    ka_halo[i] = v;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                ka_m = -INFINITY; ka_l = 0.f; ka_a0 = ka_a1 = ka_a2 = 0.f;
            }
            // one 16-logit chunk of the fused epilogue: z = acc + bias (base-2 scaled), running max / sum / weighted sums
            auto process_ka = [&](const uint32_t (&v)[16], int cc) {
                const int ch = n0 + cc * 16;
                float z[16];
                const float4* b4 = reinterpret_cast<const float4*>(bias_s + cc * 16);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 b = b4[i];
                    z[4 * i + 0] = (__uint_as_float(v[4 * i + 0]) + b.x) * 1.4426950408889634f;
                    z[4 * i + 1] = (__uint_as_float(v[4 * i + 1]) + b.y) * 1.4426950408889634f;
                    z[4 * i + 2] = (__uint_as_float(v[4 * i + 2]) + b.z) * 1.4426950408889634f;
                    z[4 * i + 3] = (__uint_as_float(v[4 * i + 3]) + b.w) * 1.4426950408889634f;
                }
                float mx = ka_m;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (ch + i < p.ka_taps) mx = fmaxf(mx, z[i]);
                const float sc = exp2f(ka_m - mx);         // 0 on the first chunk (ka_m = -inf)
                ka_l *= sc; ka_a0 *= sc; ka_a1 *= sc; ka_a2 *= sc;
                ka_m = mx;
                int dy = ch / p.ka_k, dx = ch - dy * p.ka_k;
                const float4* hrow = ka_halo + (ty + dy) * ka_pitch + 8 * t + tx;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (ch + i < p.ka_taps) {
                        const float e = exp2f(z[i] - mx);
                        const float4 nb = hrow[dx];
                        ka_l += e;
                        ka_a0 = fmaf(e, nb.x, ka_a0);
                        ka_a1 = fmaf(e, nb.y, ka_a1);
                        ka_a2 = fmaf(e, nb.z, ka_a2);
                    }
                    if (++dx == p.ka_k) { dx = 0; hrow += ka_pitch; }
                }
            };

            auto process = [&](const uint32_t (&v)[16], int cc) {
                if (KA) { process_ka(v, cc); return; }
                if (p.flags & (1 << 18)) return;               // tuning knob: epilogue without math / stores
                if (!valid && p.colsum == nullptr) return;
                const int ch = n0 + cc * 16;
                float f[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
                if (p.bias != nullptr) {
                    const float4* b4 = reinterpret_cast<const float4*>(bias_s + cc * 16);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 b = b4[i];
                        f[4 * i + 0] += b.x;
                        f[4 * i + 1] += b.y;
                        f[4 * i + 2] += b.z;
                        f[4 * i + 3] += b.w;
                    }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = apply_act(f[i], p.act, p.slope);
                if (p.mask != nullptr && valid) {
                    const uint4* mp = reinterpret_cast<const uint4*>(p.mask + pix * p.mask_cs + p.mask_coff + ch);
                    uint4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
                    uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        f[2 * i] *= h16_pos(mw[i] & 0xFFFFu) ? 1.f : p.slope;
                        f[2 * i + 1] *= h16_pos(mw[i] >> 16) ? 1.f : p.slope;
                    }
                }
                if (p.colsum != nullptr) {
                    // per-channel sum over the 32 pixels of this warp: butterfly that halves the number of
                    // live values at every step (16 shuffles instead of 80), then one atomic per channel
                    float w8[8], w4[4], w2[2];
                    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float keep = valid ? (h16 ? f[i + 8] : f[i]) : 0.f;
                        float send = valid ? (h16 ? f[i] : f[i + 8]) : 0.f;
                        w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        w4[i] = (h8 ? w8[i + 4] : w8[i]) + __shfl_xor_sync(0xffffffffu, h8 ? w8[i] : w8[i + 4], 8);
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        w2[i] = (h4 ? w4[i + 2] : w4[i]) + __shfl_xor_sync(0xffffffffu, h4 ? w4[i] : w4[i + 2], 4);
                    float w1 = (h2 ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, h2 ? w2[0] : w2[1], 2);
                    w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
                    if ((lane & 1) == 0) {
                        const int c = (h16 ? 8 : 0) + (h8 ? 4 : 0) + (h4 ? 2 : 0) + (h2 ? 1 : 0);
                        const float sc = p.colsum_scale != nullptr ? __ldg(p.colsum_scale) : 1.f;
                        atomicAdd(p.colsum + ch + c, w1 * sc);
                    }
                    if (!valid) return;
                }
                if (p.out_dtype == WCMC_F32) {
                    float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + pix * p.out_cs + p.out_coff + ch);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        op[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                } else {
                    uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + pix * p.out_cs +
                                                         p.out_coff + ch);
                    const int dt = p.out_dtype;
                    op[0] = make_uint4(pack_h2(f[0], f[1], dt), pack_h2(f[2], f[3], dt), pack_h2(f[4], f[5], dt),
                                       pack_h2(f[6], f[7], dt));
                    op[1] = make_uint4(pack_h2(f[8], f[9], dt), pack_h2(f[10], f[11], dt),
                                       pack_h2(f[12], f[13], dt), pack_h2(f[14], f[15], dt));
                }
            };

            uint32_t va[16], vb[16];
            __syncwarp();
            if (cc0 < ncc) tmem_ld16(taddr + cc0 * 16, va);
            for (int cc = cc0; cc < ncc; cc += 2 * ccs) {
                tmem_ld_wait16(va);
                const int c1 = cc + ccs;
                __syncwarp();
                if (c1 < ncc) tmem_ld16(taddr + c1 * 16, vb);
                process(va, cc);
                if (c1 < ncc) {
                    tmem_ld_wait16(vb);
                    const int c2 = c1 + ccs;
                    __syncwarp();
                    if (c2 < ncc) tmem_ld16(taddr + c2 * 16, va);
                    process(vb, c1);
                }
            }
            tc_fence_before();
            if (PAIR) mbar_arrive_cluster(acc_empty_leader + buf * 8);
            else mbar_arrive(&acc_empty[buf]);
            if (KA && n0 + p.nt >= p.cout_p) {
                // last n tile of the region: out_c = sum_k softmax_k * neighbour_c,k.  With one M tile the two warp
                // halves split the columns of the same pixels: merge the two partial softmax states first.
                if (p.mt == 1) {
                    if (half == 1) {
                        float* mg = ka_merge + m * 5;
                        mg[0] = ka_m; mg[1] = ka_l; mg[2] = ka_a0; mg[3] = ka_a1; mg[4] = ka_a2;
                    }
                    asm volatile("bar.sync 2, 256;" ::: "memory");
                    if (half == 0) {
                        const float* mg = ka_merge + m * 5;
                        const float mo = mg[0], mx = fmaxf(ka_m, mo);
                        const float s0 = exp2f(ka_m - mx), s1 = exp2f(mo - mx);
                        ka_l = ka_l * s0 + mg[1] * s1;
                        ka_a0 = ka_a0 * s0 + mg[2] * s1;
                        ka_a1 = ka_a1 * s0 + mg[3] * s1;
                        ka_a2 = ka_a2 * s0 + mg[4] * s1;
                    }
                }
                if (valid && (p.mt == 2 || half == 0)) {
                    const float inv = 1.f / ka_l;
                    const size_t plane = static_cast<size_t>(p.Ho) * p.Wo;
                    float* o = p.ka_out + static_cast<size_t>(n) * 3 * plane + static_cast<size_t>(oy) * p.Wo + ox;
                    o[0] = ka_a0 * inv;
                    o[plane] = ka_a1 * inv;
                    o[2 * plane] = ka_a2 * inv;
                }
            }
        }
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all();   // nobody leaves while the peer may still read its shared memory / TMEM
    else __syncthreads();
    if (warp == 2) {
        if (PAIR) tmem_dealloc_pair(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace wcmc

using namespace wcmc;

static int pick_nt(int cout_p) {
    if (cout_p <= 128) return cout_p;
    for (int nt = 128; nt >= 64; nt -= 16)
        if (cout_p % nt == 0) return nt;
    return 128;
}

// tuning hooks: wcmc_tuning_set("conv_pair", 0 | 1): CTA-pair (cta_group::2, M = 256) launches for the k > 1
// layers that fill the machine (flags bit 20 forces a pair launch, bit 21 forbids it);
// wcmc_tuning_set("conv_row_stages", 0 | 1): a weight stage carries a whole kernel row of taps when at least
// three such stages fit (flags bit 22 forbids it)
static int g_conv_pair = 1;
int wcmc_conv_set_pair(int v) { g_conv_pair = v ? 1 : 0; return 0; }
static int g_conv_row_stages = 1;
int wcmc_conv_set_row_stages(int v) { g_conv_row_stages = v ? 1 : 0; return 0; }
static int g_conv_share = 1;       // wcmc_tuning_set("conv_share", n): every launch plans for 1/n of the SMs
static int g_conv_share_ks = 7;
int wcmc_conv_set_share(int v) { if (v < 1 || v > 8) return 1; g_conv_share = v; return 0; }
int wcmc_conv_set_share_ks(int v) { g_conv_share_ks = v & 7; return 0; }
static int g_conv_resident = 1;    // wcmc_tuning_set("conv_resident", 0 | 1): weights resident in shared memory when they fit
int wcmc_conv_set_resident(int v) { g_conv_resident = v ? 1 : 0; return 0; }
// measurement knobs of the launch-shape model (defaults = the values measured in round 1):
//   "conv_pair_min_clk"  single-CTA MMA clocks above which a layer is launched on CTA pairs (20000)
//   "conv_item_clk"      fixed per-item cost of the M-tiles-per-region model (1000)
//   "conv_plane_slots"   halo ring depth for k > 1 layers (0 = 2 slots; 3 or 4 trade weight stages for a deeper halo
//                        prefetch -- untested idea for the short 3x3 U-Net layers, which are latency-bound per item)
static int g_conv_pair_min_clk = 20000, g_conv_item_clk = 1000, g_conv_plane_slots = 0;
int wcmc_conv_set_model(int which, int v) {
    if (v < 0) return -1;
    if (which == 0) g_conv_pair_min_clk = v;
    else if (which == 1) g_conv_item_clk = v;
    else if (which == 2 && (v == 0 || (v >= 2 && v <= kMaxPlaneSlots))) g_conv_plane_slots = v;
    else return -1;
    return 0;
}

template <bool PAIR, int TPS, bool KA = false>
static int launch_conv(const CUtensorMap& tmx, const CUtensorMap& tmw, const ConvParams& p, int grid, int smem_bytes,
                       cudaStream_t stream) {
    WCMC_FUNC_SMEM((conv_igemm_kernel<PAIR, TPS, KA>), kConvSmemMax + 1024);
    if (PAIR) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kConvThreads);
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = wcmc_pdl_enabled() ? 2 : 1;
        WCMC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<PAIR, TPS, KA>, tmx, tmw, p));
        return WCMC_OK;
    }
    WCMC_LAUNCH((conv_igemm_kernel<PAIR, TPS, KA>), grid, kConvThreads, smem_bytes, stream, tmx, tmw, p);
    return WCMC_OK;
}

static int conv2d_impl(const void* x, int x_dtype, int N, int H, int W, int x_cs, int x_coff, int cin_p,
                       const void* w_packed, int w_dtype, int cout_p, const float* bias, int ksize, int pad,
                       void* y, int y_dtype, int y_cs, int y_coff, int act, const void* mask,
                       int mask_cs, int mask_coff, float slope, float* colsum, const float* colsum_scale,
                       int flags, const float* ka_data, float* ka_out, int ka_k, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const bool ka = ka_out != nullptr;
    WCMC_REQUIRE(ksize == 1 || ksize == 3 || ksize == 5, WCMC_ESHAPE, "conv2d: ksize %d not in {1,3,5}", ksize);
    WCMC_REQUIRE((x_dtype == WCMC_BF16 || x_dtype == WCMC_F16) && (w_dtype == WCMC_BF16 || w_dtype == WCMC_F16) &&
                     (y_dtype == WCMC_BF16 || y_dtype == WCMC_F16 || y_dtype == WCMC_F32),
                 WCMC_ESHAPE, "conv2d: bad dtypes (x %d, w %d, y %d)", x_dtype, w_dtype, y_dtype);
    WCMC_REQUIRE(x_dtype == w_dtype, WCMC_ESHAPE,
                 "conv2d: x and w must share one 16-bit format (tcgen05.mma kind::f16 traps on f16 x bf16)");
    WCMC_REQUIRE(pad >= 0 && pad < ksize, WCMC_ESHAPE, "conv2d: bad pad %d", pad);
    WCMC_REQUIRE(cin_p % 16 == 0 && cout_p % 16 == 0 && cin_p > 0 && cout_p > 0, WCMC_ESHAPE,
                 "conv2d: cin_p (%d) and cout_p (%d) must be positive multiples of 16", cin_p, cout_p);
    WCMC_REQUIRE(x_cs % 8 == 0 && x_coff % 8 == 0 && x_coff + cin_p <= x_cs, WCMC_ESHAPE,
                 "conv2d: input channel stride/offset (%d,%d) must be multiples of 8", x_cs, x_coff);
    WCMC_REQUIRE(ka || (y_cs % 8 == 0 && y_coff % 8 == 0 && y_coff + cout_p <= y_cs), WCMC_ESHAPE,
                 "conv2d: output channel stride/offset (%d,%d) invalid", y_cs, y_coff);
    WCMC_REQUIRE(mask == nullptr || (mask_cs % 8 == 0 && mask_coff % 8 == 0), WCMC_ESHAPE,
                 "conv2d: mask channel stride/offset must be multiples of 8");
    WCMC_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (ka || (reinterpret_cast<uintptr_t>(y) & 15) == 0) &&
                     (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0,
                 WCMC_EALIGN, "conv2d: pointers must be 16-byte aligned");
    const int Ho = H + 2 * pad - ksize + 1, Wo = W + 2 * pad - ksize + 1;
    WCMC_REQUIRE(Ho > 0 && Wo > 0 && N > 0, WCMC_ESHAPE, "conv2d: empty output");

    ConvParams p;
    p.N = N; p.Ho = Ho; p.Wo = Wo; p.ksize = ksize; p.pad = pad;
    p.cin_p = cin_p; p.nch = (cin_p + 63) / 64;
    p.cout_p = cout_p;
    p.nt = pick_nt(cout_p);
    if ((flags >> 8) & 0xFF) p.nt = ((flags >> 8) & 0xFF);   // test override: n tile
    WCMC_REQUIRE(p.nt % 16 == 0 && p.nt >= 16 && p.nt <= 128, WCMC_ESHAPE, "conv2d: bad n tile %d", p.nt);
    p.n_tiles = (cout_p + p.nt - 1) / p.nt;
    // SMs this launch plans for: all of them, or (flags bits 24..27 = share, or the "conv_share" knob) 1/share of them
    // when the caller runs `share` such launches side by side on different streams (the two branches / the two
    // path-embedding networks of a step): two half-machine launches overlap each other's prologue, tail and partial
    // last wave instead of serialising them.
    const bool share_k = (g_conv_share_ks >> (ksize >> 1)) & 1;      // knob: which kernel sizes share (bit 0: 1x1, 1: 3x3, 2: 5x5)
    const int share_req = ((flags >> 24) & 15) > 0 ? ((flags >> 24) & 15) : (share_k ? g_conv_share : 1);
    const int sms = share_req > 1 ? std::max(2, (wcmc_num_sms() / share_req) & ~1) : wcmc_num_sms();
    // Launch shape.  Single CTA: two M tiles per region (B streamed once for 256 pixels) unless that leaves SMs
    // idle.  CTA pair (measured, profiles/r01e_pair_check.txt): worth it for the MMA-bound layers that fill the
    // machine; the 64- and 128-channel 3x3 U-Net layers are epilogue-bound and lose ~2 us to the cluster launch.  With the B operand already halved per CTA, one M
    // tile per region often wins: less quantisation of the 8-pixel-wide tiles and a fuller last wave, so mt
    // minimises waves x (mt x MMA clocks per tile + a fixed per-item cost).
    const int ksteps = (cin_p + 15) / 16;
    const long mmas_per_tile = static_cast<long>(ksize) * ksize * ksteps;
    const int ry16 = (Ho + 15) / 16;
    int mt = 2;
    {
        long items2 = static_cast<long>(N) * ((Wo + 15) / 16) * ry16 * p.n_tiles;
        if (items2 < 2L * sms) mt = 1;
    }
    bool pair = false;
    if (g_conv_pair && ksize > 1) {
        // MMA clocks of the single-CTA launch: below ~20k (about 10 us) the kernel is launch / epilogue bound
        const long items1 = static_cast<long>(N) * ((Wo + 8 * mt - 1) / (8 * mt)) * ry16 * p.n_tiles;
        const long clk1 = ((items1 + sms - 1) / sms) * mt * mmas_per_tile * (p.nt / 2);
        pair = items1 >= sms && clk1 >= g_conv_pair_min_clk;
    }
    if (flags & (1 << 20)) pair = true;
    if (flags & (1 << 21)) pair = false;
    if (pair) {
        const long tile_clk = mmas_per_tile * (p.nt / 2), fixed_clk = g_conv_item_clk;
        long best = -1;
        for (int m = 1; m <= 2; ++m) {
            const long regions = static_cast<long>(N) * ((Wo + 8 * m - 1) / (8 * m)) * ry16;
            const long items = ((regions + 1) / 2) * p.n_tiles;
            const long workers = sms / 2;
            const long cost = ((items + workers - 1) / workers) * (m * tile_clk + fixed_clk);
            if (best < 0 || cost < best) { best = cost; mt = m; }
        }
    }
    if ((flags >> 4) & 3) mt = (flags >> 4) & 3;              // test override: m tiles
    p.mt = mt;
    p.regions_x = (Wo + 8 * mt - 1) / (8 * mt);
    p.regions_y = ry16;
    p.total_items = N * p.regions_x * p.regions_y * p.n_tiles;
    p.total_regions = N * p.regions_x * p.regions_y;
    p.pair_items = ((p.total_regions + 1) / 2) * p.n_tiles;
    p.halo_w = 8 * mt + ksize - 1;
    p.halo_h = 16 + ksize - 1;
    p.out = y; p.out_cs = y_cs; p.out_coff = y_coff; p.out_dtype = y_dtype;
    p.x_dtype = x_dtype; p.w_dtype = w_dtype;
    p.bias = bias; p.act = act;
    p.mask = static_cast<const __nv_bfloat16*>(mask); p.mask_cs = mask_cs; p.mask_coff = mask_coff;
    p.slope = slope; p.flags = flags;
    p.colsum = colsum; p.colsum_scale = colsum_scale;
    p.ka_data = ka_data; p.ka_out = ka_out; p.ka_k = ka_k; p.ka_taps = ka_k * ka_k; p.ka_off = 0;
    // fused epilogue: data neighbourhood (16+20) x (8 mt + 20) float4 + the 128 x 5 merge buffer of the two warp halves
    const int ka_bytes = ka ? ((kKaHaloRows * (8 * mt + 20) * 16 + 128 * 5 * 4 + 1023) / 1024) * 1024 : 0;

    CUtensorMap tmx, tmw;
    {
        uint64_t dims[4] = {static_cast<uint64_t>(cin_p), static_cast<uint64_t>(W),
                            static_cast<uint64_t>(H), static_cast<uint64_t>(N)};
        uint64_t strides[3] = {static_cast<uint64_t>(x_cs) * 2, static_cast<uint64_t>(x_cs) * 2 * W,
                               static_cast<uint64_t>(x_cs) * 2 * W * H};
        uint32_t box[4] = {64, static_cast<uint32_t>(p.halo_w), static_cast<uint32_t>(p.halo_h), 1};
        int rc = wcmc_encode_tmap_bf16(&tmx, static_cast<const __nv_bfloat16*>(x) + x_coff, 4, dims,
                                       strides, box, 1);
        if (rc) return rc;
    }
    {
        const int taps = ksize * ksize;
        uint64_t dims[3] = {static_cast<uint64_t>(cin_p), static_cast<uint64_t>(taps),
                            static_cast<uint64_t>(cout_p)};
        uint64_t strides[2] = {static_cast<uint64_t>(cin_p) * 2, static_cast<uint64_t>(cin_p) * 2 * taps};
        uint32_t box[3] = {64, 1, static_cast<uint32_t>(pair ? p.nt / 2 : p.nt)};   // a pair CTA holds half the tile
        int rc = wcmc_encode_tmap_bf16(&tmw, w_packed, 3, dims, strides, box, 1);
        if (rc) return rc;
    }
    p.plane_stride = ((p.halo_w * p.halo_h * 128 + 1023) / 1024) * 1024;
    p.tap_stride = (((pair ? p.nt / 2 : p.nt) * 128 + 1023) / 1024) * 1024;
    // Two halo slots are enough when a chunk carries k*k taps of MMAs (the next plane loads during a whole
    // chunk); 1x1 convolutions have one tap per region and are HBM-bound: give them a deeper plane ring.
    p.plane_slots = (ksize == 1) ? kMaxPlaneSlots : (g_conv_plane_slots ? g_conv_plane_slots : 2);
    if (ksize > 1 && p.plane_slots > 2 &&
        (kConvSmemMax - kBarBytes - p.plane_slots * p.plane_stride) / p.tap_stride < 4)
        p.plane_slots = 2;   // the deeper halo ring must leave room for a useful weight ring
    const int b_room = kConvSmemMax - kBarBytes - p.plane_slots * p.plane_stride - ka_bytes;
    int tps = 1;
    if (g_conv_row_stages && !(flags & (1 << 22)) && ksize > 1 && b_room / (ksize * p.tap_stride) >= 3) tps = ksize;
    p.b_stride = tps * p.tap_stride;
    p.b_stages = b_room / p.b_stride;
    if (p.b_stages > kMaxBStages) p.b_stages = kMaxBStages;
    if (ksize == 1 && p.b_stages > 4) p.b_stages = 4;
    // Resident weights: a layer whose whole weight tensor (one n tile, every tap of every 64-channel chunk) fits the ring
    // loads it ONCE per CTA.  The 64-channel 3x3 U-Net layers re-streamed 73 KB of weights per 2-tile item -- with the
    // 41 KB halo that is 33 B/clk per SM, the TMA delivery limit (the same ~32 B/clk seen in the weight-gradient
    // kernel), for 3.4 k clocks of MMAs: bandwidth-bound on data that never changes.  Exactly `need` stages are kept so
    // that every item starts at stage 0 again.
    p.b_resident = 0;
    {
        const int need = ((ksize * ksize + tps - 1) / tps) * p.nch;
        if (g_conv_resident && !pair && !ka && ksize > 1 && p.n_tiles == 1 && (ksize * ksize) % tps == 0 &&
            need <= p.b_stages && !(flags & (1 << 16))) {
            p.b_resident = 1;
            p.b_stages = need;
        }
    }
    WCMC_REQUIRE(p.b_stages >= 2, WCMC_ESHAPE, "conv2d: shared memory carve-up failed");
    p.ka_off = kBarBytes + p.plane_slots * p.plane_stride + p.b_stages * p.b_stride;
    const int smem_bytes = 1024 + p.ka_off + ka_bytes;
    if (ka) {
        // all n tiles of a region stay on one CTA (pair): the grid is sized in regions, not items
        if (pair) {
            const int groups = p.pair_items / p.n_tiles;
            const int grid = 2 * groups < (sms & ~1) ? 2 * groups : (sms & ~1);
            if (tps == 5) return launch_conv<true, 5, true>(tmx, tmw, p, grid, smem_bytes, stream);
            return launch_conv<true, 1, true>(tmx, tmw, p, grid, smem_bytes, stream);
        }
        const int groups = p.total_items / p.n_tiles;
        if (tps == 5) return launch_conv<false, 5, true>(tmx, tmw, p, groups < sms ? groups : sms, smem_bytes, stream);
        return launch_conv<false, 1, true>(tmx, tmw, p, groups < sms ? groups : sms, smem_bytes, stream);
    }
    if (pair) {
        WCMC_REQUIRE(p.nt % 16 == 0, WCMC_ESHAPE, "conv2d: pair launch needs an n tile that is a multiple of 16");
        const int grid = 2 * p.pair_items < (sms & ~1) ? 2 * p.pair_items : (sms & ~1);
        if (tps == 5) return launch_conv<true, 5>(tmx, tmw, p, grid, smem_bytes, stream);
        if (tps == 3) return launch_conv<true, 3>(tmx, tmw, p, grid, smem_bytes, stream);
        return launch_conv<true, 1>(tmx, tmw, p, grid, smem_bytes, stream);
    }
    const int grid = p.total_items < sms ? p.total_items : sms;
    if (tps == 5) return launch_conv<false, 5>(tmx, tmw, p, grid, smem_bytes, stream);
    if (tps == 3) return launch_conv<false, 3>(tmx, tmw, p, grid, smem_bytes, stream);
    return launch_conv<false, 1>(tmx, tmw, p, grid, smem_bytes, stream);
}

extern "C" int wcmc_conv2d(const void* x, int x_dtype, int N, int H, int W, int x_cs, int x_coff, int cin_p,
                           const void* w_packed, int w_dtype, int cout_p, const float* bias, int ksize, int pad,
                           void* y, int y_dtype, int y_cs, int y_coff, int act, const void* mask,
                           int mask_cs, int mask_coff, float slope, float* colsum, const float* colsum_scale,
                           int flags, void* stream_) {
    WCMC_REQUIRE(y != nullptr, WCMC_ESHAPE, "conv2d: null output");
    return conv2d_impl(x, x_dtype, N, H, W, x_cs, x_coff, cin_p, w_packed, w_dtype, cout_p, bias, ksize, pad, y, y_dtype,
                       y_cs, y_coff, act, mask, mask_cs, mask_coff, slope, colsum, colsum_scale, flags, nullptr, nullptr,
                       0, stream_);
}

// the last layer of a KPCN branch fused with softmax + 21x21 kernel-apply -- the 441 logits of a
// pixel live in TMEM / registers only (1.65 GB of fp32 logits per 720p branch neither written nor re-read).
extern "C" int wcmc_conv2d_kernel_apply(const void* x, int x_dtype, int N, int H, int W, int x_cs, int x_coff, int cin_p,
                                        const void* w_packed, int w_dtype, int cout_p, const float* bias, int ksize,
                                        int pad, const float* data, float* out, int ka_ksize, int flags, void* stream_) {
    WCMC_REQUIRE(data != nullptr && out != nullptr && bias != nullptr, WCMC_ESHAPE, "conv2d_kernel_apply: null pointer");
    WCMC_REQUIRE(ksize == 5 && ka_ksize == 21 && cout_p == 448, WCMC_ESHAPE,
                 "conv2d_kernel_apply: built for the KPCN head (5x5 conv -> 441 logits -> 21x21 kernels); got k %d, "
                 "kernel %d, cout_p %d", ksize, ka_ksize, cout_p);
    return conv2d_impl(x, x_dtype, N, H, W, x_cs, x_coff, cin_p, w_packed, w_dtype, cout_p, bias, ksize, pad, nullptr,
                       WCMC_F32, 0, 0, WCMC_ACT_LINEAR, nullptr, 0, 0, 0.f, nullptr, nullptr, flags, data, out, ka_ksize,
                       stream_);
}
