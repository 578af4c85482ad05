// K3: convolution weight gradient on tcgen05 with MN-major operands and split-K.
//
//   dW[tap][co][ci] = sum_{n,oy,ox} dy[n,oy,ox,co] * x[n, oy+ky-pad, ox+kx-pad, ci]
//
// (autograd of nn.Conv2d inside sbmc.modules.ConvChain; reference call sites
// /root/reference/support/interfaces.py:237-238 `L_diffuse.backward(); L_specular.backward()`.)
//
// GEMM view: M = cout (128 per tile), N = cin tile (<= 128), K = pixels.  Both operands are
// "MN-major": dy and x are NHWC, i.e. the M / N index (channel) is the contiguous one, so a TMA
// box [pixels][64 channels] with 128-byte swizzle is directly a canonical MN-major SW128 UMMA
// operand (8 pixels = one 1024-byte swizzle atom; 64-channel blocks are LBO apart).  One
// K-step (16 pixels) = two rows of the 8-pixel-wide tile.  As in the forward kernel the input
// halo is loaded once per pixel tile and every tap addresses a shifted window of it through the
// descriptor start address, so one halo load feeds `tpg` taps (as many as fit in 512 TMEM
// columns).  Work item = (cout tile, cin tile, tap group, K split); partial sums go to a
// workspace [split][tap][cout_p][cin_p] with plain coalesced stores and are reduced (and
// transposed to torch's (cout,cin,k,k) layout) by wcmc_wgrad_finalize -- deterministic, no atomics.
#include <algorithm>

#include "common.cuh"

namespace wcmc {

constexpr int kWgStages = 2;
constexpr int kWgAPlane = 16384;  // 16 x 8 pixels x 128 B
constexpr int kWgBPlaneMax = 30720;  // 20 x 12 pixels x 128 B (k = 5)
constexpr int kWgStageBytes = 2 * kWgAPlane + 2 * kWgBPlaneMax;  // 94208
constexpr int kWgThreads = 192;
constexpr int kWgSmem = kWgStages * kWgStageBytes + 1024 + 128;

struct WgradParams {
    int N, Ho, Wo;
    int ksize, pad, taps;
    int cin_p, cout_p;
    int nt, ci_tiles, m_tiles;
    int cstride, tpg, tap_groups;
    int nsplit_a, nsplit_b, taps_a;   // K splits of the full tap groups / of the last (short) group; taps in full groups
    int tiles_x, tiles_y, total_tiles;
    int halo_w, halo_h, b_plane;  // b_plane: bytes between the two 64-channel halo planes
    int a_planes, b_planes;
    int x_dtype, dy_dtype;
    float* ws;
};

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmdy, const __grid_constant__ CUtensorMap tmx,
                  const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * kWgStageBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kWgStages;
    uint64_t* acc_full = empty + kWgStages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // decode the work item
    // tap groups 0 .. G-2 carry tpg taps and nsplit_a K splits each, the last group the remaining taps and nsplit_b
    const int per_mc = (p.tap_groups - 1) * p.nsplit_a + p.nsplit_b;
    int item = blockIdx.x;
    const int rr = item % per_mc; item /= per_mc;
    const int cit = item % p.ci_tiles; item /= p.ci_tiles;
    const int mtile = item;
    const bool last_group = rr >= (p.tap_groups - 1) * p.nsplit_a;
    const int tg = last_group ? p.tap_groups - 1 : rr / p.nsplit_a;
    const int split = last_group ? rr - (p.tap_groups - 1) * p.nsplit_a : rr % p.nsplit_a;
    const int nsplit = last_group ? p.nsplit_b : p.nsplit_a;
    const int tap0 = tg * p.tpg;
    const int ntap = min(p.tpg, p.taps - tap0);
    const int ci0 = cit * p.nt, co0 = mtile * 128;
    const int my_tiles = (p.total_tiles - split + nsplit - 1) / nsplit;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmdy);
        tma_prefetch_desc(&tmx);
        for (int i = 0; i < kWgStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();   // after the TMEM allocation (common.cuh: PDL rules)
    pdl_wait();

    const uint32_t stage_tx = static_cast<uint32_t>(p.a_planes * kWgAPlane +
                                                    p.b_planes * p.halo_w * p.halo_h * 128);

    if (warp == 0) {
        if (lane == 0) {
            int s = 0, ph = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int tile = split + i * nsplit;
                const int tx = tile % p.tiles_x;
                const int ty = (tile / p.tiles_x) % p.tiles_y;
                const int n = tile / (p.tiles_x * p.tiles_y);
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], stage_tx);
                uint8_t* st = smem + s * kWgStageBytes;
                for (int pl = 0; pl < p.a_planes; ++pl)
                    tma_load_4d(st + pl * kWgAPlane, &tmdy, &full[s], co0 + pl * 64, tx * 8, ty * 16, n);
                for (int pl = 0; pl < p.b_planes; ++pl)
                    tma_load_4d(st + 2 * kWgAPlane + pl * p.b_plane, &tmx, &full[s], ci0 + pl * 64,
                                tx * 8 - p.pad, ty * 16 - p.pad, n);
                if (++s == kWgStages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: warp-uniform control flow, one elected lane issues (see conv_igemm.cu).
        const uint32_t idesc = make_idesc_f16(128, p.nt, 1, 1, p.dy_dtype, p.x_dtype);
        const uint32_t b_sbo = static_cast<uint32_t>(p.halo_w * 128);
        const uint32_t a_hi = static_cast<uint32_t>(make_sdesc_sw128(0, kWgAPlane, 1024, 0) >> 32);
        const uint32_t b_hi = static_cast<uint32_t>(make_sdesc_sw128(0, 0, b_sbo, 0) >> 32);
        const uint32_t a_lbo = static_cast<uint32_t>(kWgAPlane >> 4) << 16;
        const uint32_t b_lbo = static_cast<uint32_t>(p.b_plane >> 4) << 16;
        const uint32_t b_kstep = static_cast<uint32_t>(2 * p.halo_w * 8);  // two halo rows per K16 step
        const int row_step = (p.halo_w - p.ksize) * 8;
        const int ky0 = tap0 / p.ksize, kx0 = tap0 % p.ksize;
        int s = 0, ph = 0;
        for (int i = 0; i < my_tiles; ++i) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t a_base = smem_u32(smem + s * kWgStageBytes);
            const uint32_t a_lo = a_lbo | (a_base >> 4);
            uint32_t b_lo = b_lbo | ((a_base + 2 * kWgAPlane + static_cast<uint32_t>((ky0 * p.halo_w + kx0) * 128)) >> 4);
            int kx = kx0;
            if (elect_one()) {
                for (int tl = 0; tl < ntap; ++tl) {
                    const uint32_t d = tmem_base + tl * p.cstride;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (a_lo + j * 128);
                        const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | (b_lo + j * b_kstep);
                        umma_bf16(d, ad, bd, idesc, (i > 0 || j > 0) ? 1u : 0u);
                    }
                    b_lo += 8;
                    if (++kx == p.ksize) { kx = 0; b_lo += row_step; }
                }
                umma_commit(&empty[s]);
            }
            __syncwarp();
            if (++s == kWgStages) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit(acc_full);
        __syncwarp();
    } else {
        // epilogue: TMEM lane = cout row, columns = (tap, cin)
        const int q = warp & 3;
        const int co = co0 + q * 32 + lane;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        int ncc = (p.cin_p - ci0) >> 4;
        if (ncc > (p.nt >> 4)) ncc = p.nt >> 4;
        for (int tl = 0; tl < ntap; ++tl) {
            const int tap = tap0 + tl;
            const size_t slab = static_cast<size_t>(p.cout_p) * p.cin_p;
            float* dst = last_group
                             ? p.ws + (static_cast<size_t>(p.nsplit_a) * p.taps_a +
                                       static_cast<size_t>(split) * (p.taps - p.taps_a) + (tap - p.taps_a)) * slab
                             : p.ws + (static_cast<size_t>(split) * p.taps_a + tap) * slab;
            dst += static_cast<size_t>(co) * p.cin_p + ci0;
            for (int cc = 0; cc < ncc; ++cc) {
                uint32_t v[16];
                if (my_tiles > 0) {
                    tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + tl * p.cstride + cc * 16, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0;
                }
                if (co < p.cout_p) {
                    float4* o = reinterpret_cast<float4*>(dst + cc * 16);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        o[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                           __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

struct ReduceBatch {
    wcmc_wgrad_reduce_desc d[WCMC_WGRAD_BATCH_MAX];
};

// Finalisation of many layers in one launch: blockIdx.y = layer, blockIdx.x walks (output channel co, block of
// kRedCi input channels).  For its block a CTA sums the split-K slabs with coalesced reads (ci fastest, as the
// workspace is laid out; 8 independent loads in flight per thread), transposes (tap, ci) -> (ci, tap) through
// shared memory and writes torch's dw[co][ci0 .. ci0+kRedCi)[tap] as one contiguous run (full-sector stores).
constexpr int kRedCi = 32;

__global__ void __launch_bounds__(256) wgrad_reduce_batch_kernel(const ReduceBatch rb) {
    pdl_start();
    __shared__ float tile[kRedCi * 25 + 8];
    const wcmc_wgrad_reduce_desc& L = rb.d[blockIdx.y];
    const float sc = L.scale != nullptr ? __ldg(L.scale) : 1.f;
    const long mat = static_cast<long>(L.cout_p) * L.cin_p;
    const int taps_b = L.taps - L.taps_a;
    const float* ws_b = L.ws + static_cast<long>(L.nsplit) * L.taps_a * mat;
    const int nblk = (L.cin + kRedCi - 1) / kRedCi;
    for (int item = blockIdx.x; item < L.cout * nblk; item += gridDim.x) {
        const int co = item / nblk, ci0 = (item - co * nblk) * kRedCi;
        const int nci = min(kRedCi, L.cin - ci0);
        const int per = nci * L.taps;
        for (int idx = threadIdx.x; idx < per; idx += blockDim.x) {
            const int tap = idx / nci, ci = idx - tap * nci;
            const bool a = tap < L.taps_a;
            const float* src = (a ? L.ws + tap * mat : ws_b + (tap - L.taps_a) * mat) + static_cast<long>(co) * L.cin_p +
                               ci0 + ci;
            const long slab = (a ? L.taps_a : taps_b) * mat;
            const int ns = a ? L.nsplit : L.nsplit_b;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            int k = 0;
            for (; k + 8 <= ns; k += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) acc[u] += __ldcs(src + (k + u) * slab);
            }
            for (; k < ns; ++k) acc[0] += __ldcs(src + k * slab);
            tile[ci * L.taps + tap] = (((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]))) * sc;
        }
        __syncthreads();
        float* dst = L.dw + (static_cast<long>(co) * L.cin + ci0) * L.taps;
        for (int idx = threadIdx.x; idx < per; idx += blockDim.x)
            dst[idx] = L.accumulate ? dst[idx] + tile[idx] : tile[idx];
        __syncthreads();
    }
}

}  // namespace wcmc

using namespace wcmc;

// Same K splits for every tap group by default: the kernel is bound by the L2 -> SM stream of the pixel tiles
// (94 KB per tile for <= 4 taps of MMAs), not by the MMAs, so a group with fewer taps is NOT cheaper per tile, and
// equal split counts keep the groups' CTAs walking the same tiles at the same time (measured: splits
// proportional to the taps made the 5x5 layers 1.9x slower; wcmc_tuning_set("wgrad_uniform", 0) selects that).
static int g_wgrad_uniform = 1;
int wcmc_wgrad_set_uniform(int v) { g_wgrad_uniform = v; return 0; }

static int wgrad_plan(int N, int Ho, int Wo, int cin_p, int cout_p, int ksize, WgradParams* p) {
    p->taps = ksize * ksize;
    p->cin_p = cin_p;
    p->cout_p = cout_p;
    int nt = cin_p;
    if (nt > 128) {
        nt = 128;
        for (int c = 128; c >= 64; c -= 16)
            if (cin_p % c == 0) { nt = c; break; }
    }
    p->nt = nt;
    p->ci_tiles = (cin_p + nt - 1) / nt;
    p->m_tiles = (cout_p + 127) / 128;
    p->cstride = nt <= 32 ? 32 : (nt <= 64 ? 64 : 128);
    p->tpg = std::min(p->taps, 512 / p->cstride);
    p->tap_groups = (p->taps + p->tpg - 1) / p->tpg;
    p->tiles_x = (Wo + 7) / 8;
    p->tiles_y = (Ho + 15) / 16;
    p->total_tiles = N * p->tiles_x * p->tiles_y;
    {
        const int budget = std::max(1, wcmc_num_sms() / (p->m_tiles * p->ci_tiles));   // CTAs per (cout, cin) tile
        const int G = p->tap_groups;
        const int rem = p->taps - (G - 1) * p->tpg;
        p->taps_a = (G - 1) * p->tpg;
        int nb, na;
        if (G == 1) {
            nb = na = budget;
        } else if (rem == p->tpg || g_wgrad_uniform) {
            nb = na = std::max(1, budget / G);
        } else {
            nb = std::max(1, (budget * rem + p->taps / 2) / p->taps);
            na = std::max(1, (budget - nb) / (G - 1));
        }
        p->nsplit_a = std::max(1, std::min(na, p->total_tiles));
        p->nsplit_b = std::max(1, std::min(nb, p->total_tiles));
    }
    p->halo_w = 8 + ksize - 1;
    p->halo_h = 16 + ksize - 1;
    p->b_plane = ((p->halo_w * p->halo_h * 128 + 1023) / 1024) * 1024;
    p->a_planes = 2;
    p->b_planes = (nt + 63) / 64;
    return 0;
}

extern "C" size_t wcmc_conv2d_wgrad_workspace(int N, int H, int W, int cin_p, int cout_p, int ksize, int pad) {
    WgradParams p;
    const int Ho = H + 2 * pad - ksize + 1, Wo = W + 2 * pad - ksize + 1;
    wgrad_plan(N, Ho, Wo, cin_p, cout_p, ksize, &p);
    return (static_cast<size_t>(p.nsplit_a) * p.taps_a + static_cast<size_t>(p.nsplit_b) * (p.taps - p.taps_a)) *
           cout_p * cin_p * sizeof(float);
}

extern "C" int wcmc_conv2d_wgrad_partial(const void* x, int x_dtype, int N, int H, int W, int x_cs, int x_coff,
                                         int cin_p, const void* dy, int dy_dtype, int dy_cs, int dy_coff,
                                         int cout_p, int ksize, int pad, float* dw, int cout, int cin,
                                         int accumulate, const float* scale, void* workspace,
                                         size_t workspace_bytes, wcmc_wgrad_reduce_desc* desc_out, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(desc_out != nullptr && dw != nullptr, WCMC_ESHAPE, "wgrad_partial: null desc_out / dw");
    WCMC_REQUIRE(ksize == 1 || ksize == 3 || ksize == 5, WCMC_ESHAPE, "wgrad: ksize %d not in {1,3,5}", ksize);
    WCMC_REQUIRE(pad >= 0 && pad < ksize, WCMC_ESHAPE, "wgrad: bad pad %d", pad);
    WCMC_REQUIRE((x_dtype == WCMC_BF16 || x_dtype == WCMC_F16) && (dy_dtype == WCMC_BF16 || dy_dtype == WCMC_F16),
                 WCMC_ESHAPE, "wgrad: bad dtypes (x %d, dy %d)", x_dtype, dy_dtype);
    WCMC_REQUIRE(x_dtype == dy_dtype, WCMC_ESHAPE,
                 "wgrad: x and dy must share one 16-bit format (tcgen05.mma kind::f16 traps on f16 x bf16)");
    WCMC_REQUIRE(cin_p % 16 == 0 && cout_p % 16 == 0 && cin_p > 0 && cout_p > 0, WCMC_ESHAPE,
                 "wgrad: cin_p (%d) / cout_p (%d) must be positive multiples of 16", cin_p, cout_p);
    WCMC_REQUIRE(x_cs % 8 == 0 && x_coff % 8 == 0 && dy_cs % 8 == 0 && dy_coff % 8 == 0, WCMC_ESHAPE,
                 "wgrad: channel strides/offsets must be multiples of 8");
    WCMC_REQUIRE(cout <= cout_p && cin <= cin_p, WCMC_ESHAPE, "wgrad: logical channels exceed padded");
    const int Ho = H + 2 * pad - ksize + 1, Wo = W + 2 * pad - ksize + 1;
    WCMC_REQUIRE(Ho > 0 && Wo > 0 && N > 0, WCMC_ESHAPE, "wgrad: empty output");
    WgradParams p;
    p.N = N; p.Ho = Ho; p.Wo = Wo; p.ksize = ksize; p.pad = pad;
    wgrad_plan(N, Ho, Wo, cin_p, cout_p, ksize, &p);
    size_t need = (static_cast<size_t>(p.nsplit_a) * p.taps_a + static_cast<size_t>(p.nsplit_b) * (p.taps - p.taps_a)) *
                  cout_p * cin_p * sizeof(float);
    WCMC_REQUIRE(workspace != nullptr && workspace_bytes >= need, WCMC_EWORKSPACE,
                 "wgrad: workspace too small (%zu < %zu)", workspace_bytes, need);
    p.ws = static_cast<float*>(workspace);
    p.x_dtype = x_dtype; p.dy_dtype = dy_dtype;

    CUtensorMap tmdy, tmx;
    {
        uint64_t dims[4] = {static_cast<uint64_t>(cout_p), static_cast<uint64_t>(Wo), static_cast<uint64_t>(Ho),
                            static_cast<uint64_t>(N)};
        uint64_t strides[3] = {static_cast<uint64_t>(dy_cs) * 2, static_cast<uint64_t>(dy_cs) * 2 * Wo,
                               static_cast<uint64_t>(dy_cs) * 2 * Wo * Ho};
        uint32_t box[4] = {64, 8, 16, 1};
        int rc = wcmc_encode_tmap_bf16(&tmdy, static_cast<const __nv_bfloat16*>(dy) + dy_coff, 4, dims, strides,
                                       box, 1);
        if (rc) return rc;
    }
    {
        uint64_t dims[4] = {static_cast<uint64_t>(cin_p), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                            static_cast<uint64_t>(N)};
        uint64_t strides[3] = {static_cast<uint64_t>(x_cs) * 2, static_cast<uint64_t>(x_cs) * 2 * W,
                               static_cast<uint64_t>(x_cs) * 2 * W * H};
        uint32_t box[4] = {64, static_cast<uint32_t>(p.halo_w), static_cast<uint32_t>(p.halo_h), 1};
        int rc = wcmc_encode_tmap_bf16(&tmx, static_cast<const __nv_bfloat16*>(x) + x_coff, 4, dims, strides, box,
                                       1);
        if (rc) return rc;
    }
    WCMC_FUNC_SMEM(conv_wgrad_kernel, kWgSmem);
    const int grid = p.m_tiles * p.ci_tiles * ((p.tap_groups - 1) * p.nsplit_a + p.nsplit_b);
    WCMC_LAUNCH(conv_wgrad_kernel, grid, kWgThreads, kWgSmem, stream, tmdy, tmx, p);
    WCMC_LAUNCH_CHECK();
    desc_out->ws = p.ws;
    desc_out->dw = dw;
    desc_out->scale = scale;
    desc_out->nsplit = p.nsplit_a;
    desc_out->nsplit_b = p.nsplit_b;
    desc_out->taps_a = p.taps_a;
    desc_out->cout = cout;
    desc_out->cin = cin;
    desc_out->taps = p.taps;
    desc_out->cout_p = cout_p;
    desc_out->cin_p = cin_p;
    desc_out->accumulate = accumulate;
    return WCMC_OK;
}

extern "C" int wcmc_wgrad_reduce_batch(const wcmc_wgrad_reduce_desc* host_descs, int n, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(host_descs != nullptr && n >= 0, WCMC_ESHAPE, "wgrad_reduce_batch: bad arguments");
    for (int base = 0; base < n; base += WCMC_WGRAD_BATCH_MAX) {
        const int m = std::min(WCMC_WGRAD_BATCH_MAX, n - base);
        ReduceBatch rb;
        int max_items = 1;
        for (int i = 0; i < m; ++i) {
            rb.d[i] = host_descs[base + i];
            const wcmc_wgrad_reduce_desc& d = rb.d[i];
            WCMC_REQUIRE(d.ws && d.dw && d.nsplit > 0 && d.nsplit_b > 0 && d.cout > 0 && d.cin > 0 && d.taps > 0 &&
                             d.taps_a >= 0 && d.taps_a < d.taps && d.cout <= d.cout_p && d.cin <= d.cin_p,
                         WCMC_ESHAPE, "wgrad_reduce_batch: bad descriptor %d", base + i);
            WCMC_REQUIRE(d.taps <= 25, WCMC_ESHAPE, "wgrad_reduce_batch: more than 25 taps");
            max_items = std::max(max_items, d.cout * ((d.cin + kRedCi - 1) / kRedCi));
        }
        dim3 grid(std::min(max_items, 148 * 4), m);
        WCMC_LAUNCH(wgrad_reduce_batch_kernel, grid, 256, 0, stream, rb);
        WCMC_LAUNCH_CHECK();
    }
    return WCMC_OK;
}

extern "C" int wcmc_conv2d_wgrad(const void* x, int x_dtype, int N, int H, int W, int x_cs, int x_coff, int cin_p,
                                 const void* dy, int dy_dtype, int dy_cs, int dy_coff, int cout_p, int ksize,
                                 int pad, float* dw, int cout, int cin, int accumulate, const float* scale,
                                 void* workspace, size_t workspace_bytes, void* stream_) {
    wcmc_wgrad_reduce_desc d;
    int rc = wcmc_conv2d_wgrad_partial(x, x_dtype, N, H, W, x_cs, x_coff, cin_p, dy, dy_dtype, dy_cs, dy_coff, cout_p,
                                       ksize, pad, dw, cout, cin, accumulate, scale, workspace, workspace_bytes, &d,
                                       stream_);
    if (rc) return rc;
    return wcmc_wgrad_reduce_batch(&d, 1, stream_);
}
