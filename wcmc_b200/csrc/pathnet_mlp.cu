// K6/K7: the two per-sample 1x1-conv MLPs of the path-embedding network, each fused into ONE kernel
// (/root/reference/support/networks.py:18-19 `embedding`, :23-24 `final`, forward :29-42).
//
//   K6  embed : paths (B,S,Cin,H,W) fp32 NCHW  ->  64 -> 64 -> 64  (+ mean over spp, networks.py:36)
//   K7  final : [emb (B*S px, 64) | prop (B px, 64)]  ->  128 -> outc   -> (B,S,outc,H,W) fp32
//
// Both are HBM-bound (54 kFLOP per 144 B of input) as long as the matrix products run
// on the tensor cores, so: a CTA owns a tile of 128 pixels (thread t <-> pixel row t <-> TMEM lane
// t), the layer weights stay resident in shared memory as canonical K-major SWIZZLE_128B operands
// for the whole CTA, every layer is a handful of tcgen05.mma (M = 128 pixels, N = layer width, K = 16
// per instruction) into TMEM, and the epilogue (tcgen05.ld -> bias -> activation -> 16-bit) writes the
// next layer's A operand straight back into shared memory.  Intermediate activations reach HBM only
// when the backward pass needs them (training); the NCHW fp32 -> NHWC conversion of the path
// descriptors, the mean over spp, the repeat + cat of networks.py:39-40 (prop is a second K chunk of
// the first `final` layer, read once per pixel tile, not once per sample) and the NHWC -> NCHW fp32
// conversion of the output are all folded in.  Several CTAs per SM overlap each other's load / MMA /
// epilogue phases; inside a CTA the phases are sequential (one mbarrier, __syncthreads).
#include "common.cuh"

namespace wcmc {

constexpr int kMlpThreads = 128;
constexpr int kTileBytes = 128 * 128;  // 128 rows x 64 16-bit channels

// 16-byte chunk `chunk` (8 channels) of row `row` of a K-major SWIZZLE_128B tile with 128-byte rows;
// the tile base is 1024-byte aligned, so the hardware's address-bit swizzle is chunk ^ (row & 7).
__device__ __forceinline__ void st_tile_chunk(uint8_t* tile, int row, int chunk, uint4 v) {
    *reinterpret_cast<uint4*>(tile + row * 128 + ((chunk ^ (row & 7)) << 4)) = v;
}

__device__ __forceinline__ float mlp_act(float v, int act, float slope) {
    if (act == WCMC_ACT_RELU) return fmaxf(v, 0.f);
    if (act == WCMC_ACT_LEAKY) return v > 0.f ? v : v * slope;
    return v;
}

// Copies packed weights [rows][k_total] (16-bit, row-major) into `nchunk` swizzled tiles of
// [rows][64 K] each (K zero-padded to a multiple of 64 by the caller's packing or by this loop).
__device__ __forceinline__ void load_weight_tiles(uint8_t* tiles, int tile_stride, const void* w, int rows,
                                                  int k_total) {
    const int cpr = k_total >> 3;                      // 16-byte chunks per row
    const uint4* src = static_cast<const uint4*>(w);
    for (int i = threadIdx.x; i < rows * cpr; i += kMlpThreads) {
        const int row = i / cpr, ch = i - row * cpr;
        st_tile_chunk(tiles + (ch >> 3) * tile_stride, row, ch & 7, __ldg(src + i));
    }
}

// One layer's MMAs: D[tmem] = sum over `nk` K16 steps of A (128 x 16) * B^T (N x 16); issued by one thread.
__device__ __forceinline__ void issue_layer(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, int nk,
                                            uint32_t idesc, bool first) {
    const uint64_t ad = make_sdesc_sw128(a_addr, 16, 1024);
    const uint64_t bd = make_sdesc_sw128(b_addr, 16, 1024);
    for (int k = 0; k < nk; ++k) umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
}

struct EmbedParams {
    int B, S, Cin, HW;
    const void *w1, *w2, *w3;       // packed [64][cin_p], [64][64], [64][64]
    const float *b1, *b2, *b3;      // fp32 [64]
    int cin_p, dtype, act1, act2, act3;
    float slope;
    int has_x16, has_h1, has_h2;    // optional outputs (training)
    int emb_coff;
    void* mean; int mean_cs, mean_coff;   // (B*HW, mean_cs)
};

// Four 16-column tcgen05.ld in flight, one wait.
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[4][16]) {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) tmem_ld16(taddr + cc * 16, v[cc]);
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) tmem_ld_wait16(v[cc]);
}

// smem: activation tile (x, then h1, h2, emb: each layer's MMAs have retired before its epilogue
// overwrites their A operand) | w1 | w2 | w3 (8 KB each) | biases 3 x 64 fp32 | 2 mbarriers | tmem ptr |
// fp32 staging [Cin][128 px] filled by ONE TMA box per sample (issued a whole sample ahead).
// Every 16-bit output leaves through a TMA store of the activation tile (it already has the
// 128-byte-swizzled layout): full-line writes, no per-thread partial-sector stores.
constexpr int kEmbStageOff = 41 * 1024;
static int emb_smem_bytes(int cin) { return 1024 + kEmbStageOff + cin * 512; }

__global__ void __launch_bounds__(kMlpThreads, 3)
pathnet_embed_fwd_kernel(const __grid_constant__ CUtensorMap tmp, const __grid_constant__ CUtensorMap tmx16,
                         const __grid_constant__ CUtensorMap tmh1, const __grid_constant__ CUtensorMap tmh2,
                         const __grid_constant__ CUtensorMap tmemb, const EmbedParams p) {
    pdl_wait();                // the prologue below reads packed weights / issues TMA loads
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint8_t* at = smem;
    uint8_t* wt = smem + kTileBytes;                 // w1, w2, w3 at 8 KB steps
    float* bias = reinterpret_cast<float*>(wt + 3 * 8192);
    uint64_t* bar = reinterpret_cast<uint64_t*>(bias + 3 * 64);   // [0] MMA done, [1] staging full
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
    const float* stage = reinterpret_cast<const float*>(smem + kEmbStageOff);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int b = blockIdx.y;
    const int pix0 = blockIdx.x * 128;
    const uint32_t stage_bytes = static_cast<uint32_t>(p.Cin) * 512u;

    if (tid == 0) {
        tma_prefetch_desc(&tmp);
        tma_prefetch_desc(&tmemb);
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
        mbar_expect_tx(&bar[1], stage_bytes);
        tma_load_2d(smem + kEmbStageOff, &tmp, &bar[1], pix0, b * p.S * p.Cin);
    }
    for (int i = tid; i < 8192 / 16; i += kMlpThreads) reinterpret_cast<uint4*>(wt)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    load_weight_tiles(wt, 0, p.w1, 64, p.cin_p);
    load_weight_tiles(wt + 8192, 0, p.w2, 64, 64);
    load_weight_tiles(wt + 16384, 0, p.w3, 64, 64);
    for (int i = tid; i < 64; i += kMlpThreads) {
        bias[i] = __ldg(p.b1 + i);
        bias[64 + i] = __ldg(p.b2 + i);
        bias[128 + i] = __ldg(p.b3 + i);
    }
    if (warp == 0) tmem_alloc(tmem_ptr, 64);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_launch_dependents();   // after the TMEM allocation (common.cuh: PDL rules)
    const uint32_t tmem = *tmem_ptr;
    const uint32_t tlane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t idesc = make_idesc_f16(128, 64, 0, 0, p.dtype, p.dtype);
    const uint32_t aa = smem_u32(at), wa = smem_u32(wt);

    const int pix = pix0 + tid;
    const bool valid = pix < p.HW;
    const int ngrp = p.cin_p >> 3;
    uint32_t phase = 0, tphase = 0;
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;

    // The MMAs of the layer just issued have retired and the TMA store that was reading the tile has
    // finished with it: the tile may be overwritten.
    auto tile_free = [&]() {
        mbar_wait(&bar[0], phase);
        phase ^= 1;
        if (tid == 0) tma_store_wait_read();
        tc_fence_after();
        __syncthreads();
    };
    // TMEM -> bias -> act -> 16-bit -> activation tile (+ fp32 running sum of the last layer)
    auto epilogue = [&](const float* bs, int act, bool accumulate) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t v[2][16];
            tmem_ld16(tlane + half * 32, v[0]);
            tmem_ld16(tlane + half * 32 + 16, v[1]);
            tmem_ld_wait16(v[0]);
            tmem_ld_wait16(v[1]);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int cc = half * 2 + q;
                float f[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = mlp_act(__uint_as_float(v[q][i]) + bs[cc * 16 + i], act, p.slope);
                if (accumulate) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[cc * 16 + i] += f[i];
                }
                const int dt = p.dtype;
                st_tile_chunk(at, tid, 2 * cc, make_uint4(pack_h2(f[0], f[1], dt), pack_h2(f[2], f[3], dt),
                                                          pack_h2(f[4], f[5], dt), pack_h2(f[6], f[7], dt)));
                st_tile_chunk(at, tid, 2 * cc + 1, make_uint4(pack_h2(f[8], f[9], dt), pack_h2(f[10], f[11], dt),
                                                              pack_h2(f[12], f[13], dt), pack_h2(f[14], f[15], dt)));
            }
        }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
    };

    for (int s = 0; s < p.S; ++s) {
        const int img = b * p.S + s;
        // ---- input: fp32 [c][pixel] staging -> 16-bit K-major tile (64 channels, zero padded) ----
        mbar_wait(&bar[1], tphase);
        tphase ^= 1;
        if (s > 0) {            // the emb store of the previous sample must have read the tile
            if (tid == 0) tma_store_wait_read();
            __syncthreads();
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (g < ngrp) {
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c = 8 * g + i;
                    f[i] = c < p.Cin ? stage[c * 128 + tid] : 0.f;
                }
                const int dt = p.dtype;
                v = make_uint4(pack_h2(f[0], f[1], dt), pack_h2(f[2], f[3], dt), pack_h2(f[4], f[5], dt),
                               pack_h2(f[6], f[7], dt));
            }
            st_tile_chunk(at, tid, g, v);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        // ---- layer 1 (and the next sample's TMA: the staging buffer has been consumed by every thread) ----
        if (tid == 0) {
            if (p.has_x16) {
                tma_store_3d(&tmx16, at, 0, pix0, img);
                tma_store_commit();
            }
            if (s + 1 < p.S) {
                mbar_expect_tx(&bar[1], stage_bytes);
                tma_load_2d(smem + kEmbStageOff, &tmp, &bar[1], pix0, (img + 1) * p.Cin);
            }
            tc_fence_after();
            issue_layer(tmem, aa, wa, p.cin_p >> 4, idesc, true);
            umma_commit(&bar[0]);
        }
        tile_free();
        epilogue(bias, p.act1, false);
        // ---- layer 2 ----
        if (tid == 0) {
            if (p.has_h1) {
                tma_store_3d(&tmh1, at, 0, pix0, img);
                tma_store_commit();
            }
            tc_fence_after();
            issue_layer(tmem, aa, wa + 8192, 4, idesc, true);
            umma_commit(&bar[0]);
        }
        tile_free();
        epilogue(bias + 64, p.act2, false);
        // ---- layer 3 ----
        if (tid == 0) {
            if (p.has_h2) {
                tma_store_3d(&tmh2, at, 0, pix0, img);
                tma_store_commit();
            }
            tc_fence_after();
            issue_layer(tmem, aa, wa + 16384, 4, idesc, true);
            umma_commit(&bar[0]);
        }
        tile_free();
        epilogue(bias + 128, p.act3, true);
        if (tid == 0) {
            tma_store_3d(&tmemb, at, p.emb_coff, pix0, img);
            tma_store_commit();
        }
    }
    if (p.mean != nullptr && valid) {
        const float inv = 1.f / p.S;
        const int dt = p.dtype;
        uint4* g = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.mean) +
                                            (static_cast<size_t>(b) * p.HW + pix) * p.mean_cs + p.mean_coff);
#pragma unroll
        for (int c = 0; c < 8; ++c)
            g[c] = make_uint4(pack_h2(acc[8 * c] * inv, acc[8 * c + 1] * inv, dt),
                              pack_h2(acc[8 * c + 2] * inv, acc[8 * c + 3] * inv, dt),
                              pack_h2(acc[8 * c + 4] * inv, acc[8 * c + 5] * inv, dt),
                              pack_h2(acc[8 * c + 6] * inv, acc[8 * c + 7] * inv, dt));
    }
    if (tid == 0) tma_store_wait_all();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

struct FinalParams {
    const void *w1, *w2;                        // packed [128][128], [outc_p][128]
    const float *b1, *b2;                       // fp32 [128], [outc_p]
    int has_h;                                  // optional (B*S*HW, 128) hidden activation (training)
    float* out;                                 // (B,S,outc,HW) fp32
    int B, S, HW, outc, outc_p, dtype, act1, act2;
    int emb_coff, prop_coff;
    float slope;
};

// smem: emb tile | h (2 x 64-channel chunks) | prop tile | w1 (2 x 16 KB) | w2 (2 x 4 KB, outc_p <= 32) |
// biases (128 + 32) | 2 mbarriers | tmem ptr;  ~107 KB: 2 CTAs / SM.  emb / prop tiles arrive by TMA
// (128-byte swizzle = the UMMA operand layout); the next sample's emb tile is requested as soon as the
// layer-1 MMAs of the current one have retired; h leaves by TMA store.
constexpr int kFinSmem = 4 * kTileBytes + 2 * 16384 + 2 * 4096 + 160 * 4 + 64 + 1024;

__global__ void __launch_bounds__(kMlpThreads, 2)
pathnet_final_fwd_kernel(const __grid_constant__ CUtensorMap tme, const __grid_constant__ CUtensorMap tmpr,
                         const __grid_constant__ CUtensorMap tmh, const FinalParams p) {
    pdl_wait();                // the prologue below reads packed weights / issues TMA loads
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint8_t* et = smem;
    uint8_t* ht = smem + kTileBytes;                 // two 64-channel chunks
    uint8_t* pt = smem + 3 * kTileBytes;
    uint8_t* w1t = smem + 4 * kTileBytes;            // two K chunks of [128][64]
    uint8_t* w2t = w1t + 2 * 16384;                  // two K chunks of [outc_p <= 32][64]
    float* bias = reinterpret_cast<float*>(w2t + 2 * 4096);
    uint64_t* bar = reinterpret_cast<uint64_t*>(bias + 160);    // [0] MMA done, [1] tiles full
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int b = blockIdx.y;
    const int pix0 = blockIdx.x * 128;
    const int pix = pix0 + tid;
    const bool valid = pix < p.HW;

    if (tid == 0) {
        tma_prefetch_desc(&tme);
        tma_prefetch_desc(&tmpr);
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
        mbar_expect_tx(&bar[1], 2 * kTileBytes);
        tma_load_3d(pt, &tmpr, &bar[1], p.prop_coff, pix0, b);
        tma_load_3d(et, &tme, &bar[1], p.emb_coff, pix0, b * p.S);
    }
    load_weight_tiles(w1t, 16384, p.w1, 128, 128);
    load_weight_tiles(w2t, 4096, p.w2, p.outc_p, 128);
    for (int i = tid; i < 128; i += kMlpThreads) bias[i] = __ldg(p.b1 + i);
    for (int i = tid; i < p.outc_p; i += kMlpThreads) bias[128 + i] = __ldg(p.b2 + i);
    if (warp == 0) tmem_alloc(tmem_ptr, 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_launch_dependents();   // after the TMEM allocation (common.cuh: PDL rules)
    const uint32_t tmem = *tmem_ptr;
    const uint32_t tlane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t idesc1 = make_idesc_f16(128, 128, 0, 0, p.dtype, p.dtype);
    const uint32_t idesc2 = make_idesc_f16(128, p.outc_p, 0, 0, p.dtype, p.dtype);
    const uint32_t ea = smem_u32(et), pa = smem_u32(pt), ha = smem_u32(ht), w1a = smem_u32(w1t),
                   w2a = smem_u32(w2t);
    uint32_t phase = 0, tphase = 0;

    for (int s = 0; s < p.S; ++s) {
        const int img = b * p.S + s;
        // ---- layer 1: [emb | prop] (K = 128) -> 128 ----
        if (tid == 0) {
            mbar_wait(&bar[1], tphase);
            tc_fence_after();
            issue_layer(tmem, ea, w1a, 4, idesc1, true);
            issue_layer(tmem, pa, w1a + 16384, 4, idesc1, false);
            umma_commit(&bar[0]);
        }
        tphase ^= 1;
        mbar_wait(&bar[0], phase);
        phase ^= 1;
        if (tid == 0) {
            if (s + 1 < p.S) {   // the emb tile is free again: request the next sample's
                mbar_expect_tx(&bar[1], kTileBytes);
                tma_load_3d(et, &tme, &bar[1], p.emb_coff, pix0, img + 1);
            }
            tma_store_wait_read();   // the previous sample's h store has read the h tiles
        }
        tc_fence_after();
        __syncthreads();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t v[4][16];
            tmem_ld64(tlane + half * 64, v);
            uint8_t* tile = ht + half * kTileBytes;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                float f[16];
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    f[i] = mlp_act(__uint_as_float(v[cc][i]) + bias[half * 64 + cc * 16 + i], p.act1, p.slope);
                const int dt = p.dtype;
                st_tile_chunk(tile, tid, 2 * cc, make_uint4(pack_h2(f[0], f[1], dt), pack_h2(f[2], f[3], dt),
                                                            pack_h2(f[4], f[5], dt), pack_h2(f[6], f[7], dt)));
                st_tile_chunk(tile, tid, 2 * cc + 1,
                              make_uint4(pack_h2(f[8], f[9], dt), pack_h2(f[10], f[11], dt),
                                         pack_h2(f[12], f[13], dt), pack_h2(f[14], f[15], dt)));
            }
        }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
        // ---- layer 2: 128 -> outc ----
        if (tid == 0) {
            if (p.has_h) {
                tma_store_3d(&tmh, ht, 0, pix0, img);
                tma_store_3d(&tmh, ht + kTileBytes, 64, pix0, img);
                tma_store_commit();
            }
            tc_fence_after();
            issue_layer(tmem + 128, ha, w2a, 4, idesc2, true);
            issue_layer(tmem + 128, ha + kTileBytes, w2a + 4096, 4, idesc2, false);
            umma_commit(&bar[0]);
        }
        mbar_wait(&bar[0], phase);
        phase ^= 1;
        tc_fence_after();
        float* orow = p.out + static_cast<size_t>(img) * p.outc * p.HW + pix;
        for (int cc = 0; cc < (p.outc_p >> 4); ++cc) {
            uint32_t v[16];
            tmem_ld16(tlane + 128 + cc * 16, v);
            tmem_ld_wait16(v);
            if (valid) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int c = cc * 16 + i;
                    if (c < p.outc)
                        __stcs(orow + static_cast<size_t>(c) * p.HW,
                               mlp_act(__uint_as_float(v[i]) + bias[128 + c], p.act2, p.slope));
                }
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    if (tid == 0) tma_store_wait_all();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace wcmc

using namespace wcmc;

static bool act_ok(int a) { return a == WCMC_ACT_LINEAR || a == WCMC_ACT_RELU || a == WCMC_ACT_LEAKY; }
static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// (npix-per-image HW, images) view of a 16-bit NHWC tensor with channel stride cs: box = 64 ch x 128 px x 1
static int act_tmap(CUtensorMap* m, const void* base, int cs, int HW, long images) {
    uint64_t dims[3] = {static_cast<uint64_t>(cs), static_cast<uint64_t>(HW), static_cast<uint64_t>(images)};
    uint64_t strides[2] = {static_cast<uint64_t>(cs) * 2, static_cast<uint64_t>(cs) * 2 * HW};
    uint32_t box[3] = {64, 128, 1};
    return wcmc_encode_tmap(m, WCMC_BF16, base, 3, dims, strides, box, 1);
}

extern "C" int wcmc_pathnet_embed_fwd(const float* paths, int B, int S, int Cin, int HW, const void* w1,
                                      const float* b1, const void* w2, const float* b2, const void* w3,
                                      const float* b3, int cin_p, int dtype, int act1, int act2, int act3,
                                      float slope, void* x16, void* h1, void* h2, void* emb, int emb_cs,
                                      int emb_coff, void* mean, int mean_cs, int mean_coff, void* stream) {
    WCMC_REQUIRE(B > 0 && S > 0 && HW > 0 && Cin > 0 && B <= 65535, WCMC_ESHAPE, "pathnet_embed_fwd: bad shape");
    WCMC_REQUIRE(cin_p % 16 == 0 && cin_p >= Cin && cin_p <= 64, WCMC_ESHAPE,
                 "pathnet_embed_fwd: cin_p %d must be a multiple of 16 in [Cin, 64]", cin_p);
    WCMC_REQUIRE(dtype == WCMC_BF16 || dtype == WCMC_F16, WCMC_ESHAPE, "pathnet_embed_fwd: dtype must be 16-bit");
    WCMC_REQUIRE(act_ok(act1) && act_ok(act2) && act_ok(act3), WCMC_ESHAPE, "pathnet_embed_fwd: bad activation");
    WCMC_REQUIRE(paths && w1 && w2 && w3 && b1 && b2 && b3 && emb, WCMC_ESHAPE, "pathnet_embed_fwd: null pointer");
    WCMC_REQUIRE(emb_cs % 8 == 0 && emb_coff % 8 == 0 && emb_coff + 64 <= emb_cs, WCMC_ESHAPE,
                 "pathnet_embed_fwd: emb channel stride/offset (%d,%d)", emb_cs, emb_coff);
    WCMC_REQUIRE(mean == nullptr || (mean_cs % 8 == 0 && mean_coff % 8 == 0 && mean_coff + 64 <= mean_cs),
                 WCMC_ESHAPE, "pathnet_embed_fwd: mean channel stride/offset (%d,%d)", mean_cs, mean_coff);
    WCMC_REQUIRE(al16(w1) && al16(w2) && al16(w3) && al16(emb) && al16(mean) && al16(x16) && al16(h1) && al16(h2),
                 WCMC_EALIGN, "pathnet_embed_fwd: pointers must be 16-byte aligned");
    WCMC_REQUIRE(HW % 4 == 0 && (reinterpret_cast<uintptr_t>(paths) & 15) == 0, WCMC_EALIGN,
                 "pathnet_embed_fwd: H*W must be a multiple of 4 and paths 16-byte aligned (TMA row stride)");
    WCMC_REQUIRE(Cin <= 64, WCMC_ESHAPE, "pathnet_embed_fwd: Cin %d > 64", Cin);
    EmbedParams p{B, S, Cin, HW, w1, w2, w3, b1, b2, b3, cin_p, dtype, act1, act2, act3, slope,
                  x16 != nullptr, h1 != nullptr, h2 != nullptr, emb_coff, mean, mean_cs, mean_coff};
    CUtensorMap tmp, tmx16, tmh1, tmh2, tmemb;
    {
        uint64_t dims[2] = {static_cast<uint64_t>(HW), static_cast<uint64_t>(B) * S * Cin};
        uint64_t strides[1] = {static_cast<uint64_t>(HW) * 4};
        uint32_t box[2] = {128, static_cast<uint32_t>(Cin)};
        int rc = wcmc_encode_tmap(&tmp, WCMC_F32, paths, 2, dims, strides, box, 0);
        if (rc) return rc;
    }
    const long images = static_cast<long>(B) * S;
    int rc = act_tmap(&tmemb, emb, emb_cs, HW, images);
    if (rc) return rc;
    tmx16 = tmh1 = tmh2 = tmemb;   // placeholders when the optional outputs are absent (never dereferenced)
    if (x16 && (rc = act_tmap(&tmx16, x16, 64, HW, images))) return rc;
    if (h1 && (rc = act_tmap(&tmh1, h1, 64, HW, images))) return rc;
    if (h2 && (rc = act_tmap(&tmh2, h2, 64, HW, images))) return rc;
    const int smem_bytes = emb_smem_bytes(Cin);
    WCMC_FUNC_SMEM(pathnet_embed_fwd_kernel, emb_smem_bytes(64));
    dim3 grid((HW + 127) / 128, B);
    WCMC_LAUNCH(pathnet_embed_fwd_kernel, grid, kMlpThreads, smem_bytes, static_cast<cudaStream_t>(stream), tmp, tmx16, tmh1, tmh2, tmemb, p);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" int wcmc_pathnet_final_fwd(const void* emb, int emb_cs, int emb_coff, const void* prop, int prop_cs,
                                      int prop_coff, const void* w1, const float* b1, const void* w2,
                                      const float* b2, int outc, int outc_p, int dtype, int act1, int act2,
                                      float slope, void* h, float* out, int B, int S, int HW, void* stream) {
    WCMC_REQUIRE(B > 0 && S > 0 && HW > 0 && B <= 65535, WCMC_ESHAPE, "pathnet_final_fwd: bad shape");
    WCMC_REQUIRE(outc > 0 && outc_p % 16 == 0 && outc_p >= outc && outc_p <= 32, WCMC_ESHAPE,
                 "pathnet_final_fwd: outc_p %d must be a multiple of 16 in [outc, 32]", outc_p);
    WCMC_REQUIRE(dtype == WCMC_BF16 || dtype == WCMC_F16, WCMC_ESHAPE, "pathnet_final_fwd: dtype must be 16-bit");
    WCMC_REQUIRE(act_ok(act1) && act_ok(act2), WCMC_ESHAPE, "pathnet_final_fwd: bad activation");
    WCMC_REQUIRE(emb && prop && w1 && w2 && b1 && b2 && out, WCMC_ESHAPE, "pathnet_final_fwd: null pointer");
    WCMC_REQUIRE(emb_cs % 8 == 0 && emb_coff % 8 == 0 && emb_coff + 64 <= emb_cs && prop_cs % 8 == 0 &&
                     prop_coff % 8 == 0 && prop_coff + 64 <= prop_cs,
                 WCMC_ESHAPE, "pathnet_final_fwd: channel strides/offsets");
    WCMC_REQUIRE(al16(emb) && al16(prop) && al16(w1) && al16(w2) && al16(h), WCMC_EALIGN,
                 "pathnet_final_fwd: pointers must be 16-byte aligned");
    FinalParams p{w1, w2, b1, b2, h != nullptr, out, B, S, HW, outc, outc_p, dtype, act1, act2, emb_coff, prop_coff,
                  slope};
    CUtensorMap tme, tmpr, tmh;
    int rc = act_tmap(&tme, emb, emb_cs, HW, static_cast<long>(B) * S);
    if (rc) return rc;
    if ((rc = act_tmap(&tmpr, prop, prop_cs, HW, B))) return rc;
    tmh = tme;
    if (h && (rc = act_tmap(&tmh, h, 128, HW, static_cast<long>(B) * S))) return rc;
    WCMC_FUNC_SMEM(pathnet_final_fwd_kernel, kFinSmem);
    dim3 grid((HW + 127) / 128, B);
    WCMC_LAUNCH(pathnet_final_fwd_kernel, grid, kMlpThreads, kFinSmem, static_cast<cudaStream_t>(stream), tme, tmpr, tmh, p);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}
