// K3 (round 2): weight gradients of EVERY layer of a conv chain in one launch.
//
//   dW_l[tap][co][ci] = sum_{n,oy,ox} dy_l[n,oy,ox,co] * x_l[n, oy+ky-pad, ox+kx-pad, ci]        for every layer l
//
// (autograd of the nn.Conv2d modules of sbmc.modules.ConvChain; reference call site
// /root/reference/support/interfaces.py:237-238 `L_diffuse.backward(); L_specular.backward()`.)
//
// Why grouped.  One layer's weight gradient is a GEMM with a tiny output (25 x 100 x 100) and a huge reduction
// dimension (8 x 116^2 pixels): launched alone (round 1, conv_wgrad.cu) it needs ~21 K splits per tap group to fill
// 148 SMs, every split writes a full fp32 slab (33 MB per layer, 1.07 GB per step re-read by the reduction), and
// the 32^2 / 64^2 U-Net levels cannot fill the machine at all.  A backward pass has all its (x_l, dy_l) pairs alive
// at its end, so here the host hands the whole chain to ONE launch: the 148 CTAs are dealt out to the layers in
// proportion to their cost, a layer gets a few TEAMS (K splits, typically 2-4), and slab traffic drops ~8x.
//
// GEMM view per CTA (as in conv_wgrad.cu): M = cout tile (128), N = cin tile (<= 128), K = pixels; both operands
// MN-major straight from the NHWC tensors (a TMA box [pixels][64 channels] with 128-byte swizzle IS the canonical
// MN-major SW128 operand); taps are shifted windows of one x halo, addressed through the descriptor start address.
// New here:
//   * a tap group is a run of whole KERNEL ROWS, so the halo only carries the extra columns (8+k-1) and not the extra
//     rows: 40 KB per 8x8-pixel stage instead of 94 KB per 8x16, five stages in flight instead of two;
//   * a 5x5 row of a 100-channel layer (5 taps x N = 112 columns = 560 > 512 TMEM columns) is packed at a column
//     stride of 100: accumulator t owns columns [100 t, 100 t + 112), its last 12 columns -- the zero-padded input
//     channels 100..111, whose products are exactly 0 -- overlap the first columns of accumulator t+1.  The
//     initialising (non-accumulating) MMAs run in ascending tap order, so every accumulator's own columns are
//     written last; afterwards the overlap only ever receives += 0.   4 * 100 + 112 = 512.
// A team = the (cout tile, cin tile, tap group) members of one layer and one K split; its members walk the same pixel
// tiles at the same time, so the L2 serves each tile to all of them from one fill.  Partial sums go to
// ws[team][tap][cout_p][cin_p] with plain stores; wgrad_reduce_batch_kernel (conv_wgrad.cu) sums the teams in a
// fixed order and writes torch's (cout,cin,k,k) layout -- deterministic, no atomics.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace wcmc {

constexpr int kGgMaxLayers = 16;
constexpr int kGgThreads = 192;
constexpr int kGgTileW = 8, kGgTileH = 8;           // 64 output pixels = 4 K16 steps per stage
constexpr int kGgAPlane = kGgTileW * kGgTileH * 128;  // 8 KB: 64 pixels x 64 channels
constexpr int kGgMaxStages = 6;
constexpr int kGgSmemBudget = 208 * 1024;

struct GgLayer {
    CUtensorMap tmdy, tmx;
    float* ws;                       // slabs [slab][tap][cout_p][cin_p]
    int tiles_x, tiles_y, total_tiles;
    int ksize, pad, taps;
    int cin_p, cout_p;
    int nt, ci_tiles, m_tiles;
    int cstride, tpg, groups;        // TMEM column stride of a tap, taps per group, groups
    int halo_w, box_h, b_plane, b_planes;
    int stage_bytes, stages;
    int dtype;
};

// A CHAIN = layers of identical structure (filter size, channel counts), e.g. the seven 100->100 layers of a KPCN
// stack: their pixel tiles form one line that the chain's teams cut into equal contiguous pieces, whatever the layer
// boundaries -- a team that crosses a boundary drains its accumulators into the finished layer's slab and goes on
// with the next layer's tensor maps.  Balance is then limited by three or four chains, not by nine layers.
struct GgSeg { int layer, t0, t1, slab; };            // tiles [t0, t1) of `layer`, partial sums -> slab `slab`
struct GgChain { int first_cta, per_team, teams, first_team; };
constexpr int kGgMaxSeg = 4;
constexpr int kGgMaxTeams = 148;

struct GgParams {
    int n_layers, n_chains;
    GgChain C[kGgMaxLayers];
    GgLayer L[kGgMaxLayers];
    GgSeg S[kGgMaxTeams][kGgMaxSeg];
    unsigned char nseg[kGgMaxTeams];
};

__global__ void __launch_bounds__(kGgThreads, 1) conv_wgrad_group_kernel(const __grid_constant__ GgParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    __shared__ uint64_t full[kGgMaxStages], empty[kGgMaxStages], acc_full, acc_empty;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- which chain, which member of which team ----
    int ci_ = 0;
    while (ci_ + 1 < P.n_chains && static_cast<int>(blockIdx.x) >= P.C[ci_ + 1].first_cta) ++ci_;
    const GgChain& C = P.C[ci_];
    const int local = blockIdx.x - C.first_cta;
    const int team_l = local / C.per_team;
    const int team = C.first_team + team_l;              // row of the segment table
    const int nseg = P.nseg[team];
    const GgLayer& L0 = P.L[P.S[team][0].layer];         // structure (identical for every layer of the chain)
    int r = local - team_l * C.per_team;
    const int grp = r % L0.groups; r /= L0.groups;
    const int cit = r % L0.ci_tiles;
    const int mtile = r / L0.ci_tiles;
    const int tap0 = grp * L0.tpg;
    const int ntap = min(L0.tpg, L0.taps - tap0);
    const int ky_first = tap0 / L0.ksize;
    const int ci0 = cit * L0.nt, co0 = mtile * 128;
    const int stages = L0.stages;
    const int stage_bytes = L0.stage_bytes;

    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(&acc_full, 1);
        mbar_init(&acc_empty, 4);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_launch_dependents();   // after the TMEM allocation (common.cuh: PDL rules)
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t stage_tx = static_cast<uint32_t>(2 * kGgAPlane + L0.b_planes * L0.halo_w * L0.box_h * 128);
            int s = 0, ph = 0;
            for (int sg = 0; sg < nseg; ++sg) {
                const GgSeg seg = P.S[team][sg];
                const GgLayer& L = P.L[seg.layer];
                tma_prefetch_desc(&L.tmdy);
                tma_prefetch_desc(&L.tmx);
                const int per_img = L.tiles_x * L.tiles_y;
                for (int tile = seg.t0; tile < seg.t1; ++tile) {
                    const int n = tile / per_img;
                    const int rr = tile - n * per_img;
                    const int ty = rr / L.tiles_x, tx = rr - ty * L.tiles_x;
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], stage_tx);
                    uint8_t* st = smem + s * stage_bytes;
                    tma_load_4d(st, &L.tmdy, &full[s], co0, tx * kGgTileW, ty * kGgTileH, n);
                    tma_load_4d(st + kGgAPlane, &L.tmdy, &full[s], co0 + 64, tx * kGgTileW, ty * kGgTileH, n);
                    for (int pl = 0; pl < L.b_planes; ++pl)
                        tma_load_4d(st + 2 * kGgAPlane + pl * L.b_plane, &L.tmx, &full[s], ci0 + pl * 64,
                                    tx * kGgTileW - L.pad, ty * kGgTileH - L.pad + ky_first, n);
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: warp-uniform control flow, one elected lane issues.
        const uint32_t idesc = make_idesc_f16(128, L0.nt, 1, 1, L0.dtype, L0.dtype);
        const uint32_t b_sbo = static_cast<uint32_t>(L0.halo_w * 128);
        const uint32_t a_hi = static_cast<uint32_t>(make_sdesc_sw128(0, kGgAPlane, 1024, 0) >> 32);
        const uint32_t b_hi = static_cast<uint32_t>(make_sdesc_sw128(0, 0, b_sbo, 0) >> 32);
        const uint32_t a_lbo = static_cast<uint32_t>(kGgAPlane >> 4) << 16;
        const uint32_t b_lbo = static_cast<uint32_t>(L0.b_plane >> 4) << 16;
        const uint32_t b_kstep = static_cast<uint32_t>(2 * L0.halo_w * 8);     // two halo rows per K16 step (16-byte units)
        const int row_step = (L0.halo_w - L0.ksize) * 8;
        const int kx0 = tap0 - ky_first * L0.ksize;
        const int ksize = L0.ksize, cstride = L0.cstride;
        int s = 0, ph = 0;
        for (int sg = 0; sg < nseg; ++sg) {
            const int ntiles = P.S[team][sg].t1 - P.S[team][sg].t0;
            if (sg > 0) {                      // the epilogue warps have drained the previous segment's accumulators
                mbar_wait(&acc_empty, (sg - 1) & 1);
                tc_fence_after();
            }
            for (int i = 0; i < ntiles; ++i) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a_base = smem_u32(smem + s * stage_bytes);
                const uint32_t a_lo = a_lbo | (a_base >> 4);
                uint32_t b_lo = b_lbo | ((a_base + 2 * kGgAPlane + static_cast<uint32_t>(kx0 * 128)) >> 4);
                int kx = kx0;
                if (elect_one()) {
                    for (int tl = 0; tl < ntap; ++tl) {
                        const uint32_t d = tmem_base + tl * cstride;
#pragma unroll
                        for (int j = 0; j < kGgTileH / 2; ++j) {
                            const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (a_lo + j * 128);
                            const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | (b_lo + j * b_kstep);
                            umma_bf16(d, ad, bd, idesc, (i > 0 || j > 0) ? 1u : 0u);
                        }
                        b_lo += 8;
                        if (++kx == ksize) { kx = 0; b_lo += row_step; }
                    }
                    umma_commit(&empty[s]);
                }
                __syncwarp();
                if (++s == stages) { s = 0; ph ^= 1; }
            }
            if (elect_one()) umma_commit(&acc_full);
            __syncwarp();
        }
    } else {
        // epilogue: TMEM lane = cout row, columns = (tap, cin)
        const int q = warp & 3;
        const int co = co0 + q * 32 + lane;
        int ncc = (L0.cin_p - ci0) >> 4;
        if (ncc > (L0.nt >> 4)) ncc = L0.nt >> 4;
        const size_t mat = static_cast<size_t>(L0.cout_p) * L0.cin_p;
        for (int sg = 0; sg < nseg; ++sg) {
            const GgSeg seg = P.S[team][sg];
            const GgLayer& L = P.L[seg.layer];
            const bool any = seg.t1 > seg.t0;
            mbar_wait(&acc_full, sg & 1);
            tc_fence_after();
            for (int tl = 0; tl < ntap; ++tl) {
                float* dst = L.ws + (static_cast<size_t>(seg.slab) * L.taps + (tap0 + tl)) * mat +
                             static_cast<size_t>(co) * L.cin_p + ci0;
                for (int cc = 0; cc < ncc; ++cc) {
                    uint32_t v[16];
                    if (any) {
                        tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + tl * L.cstride + cc * 16, v);
                        tmem_ld_wait16(v);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = 0;
                    }
                    if (co < L.cout_p) {
                        float4* o = reinterpret_cast<float4*>(dst + cc * 16);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            o[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                               __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace wcmc

using namespace wcmc;

// knobs (wcmc_tuning_set): column packing of 5x5 rows, minimum teams, pixel-tile cost model
static int g_gg_pack = 1;        // 1: column stride = logical cin when that lets a whole kernel row fit in TMEM
static int g_gg_enable = 1;      // 0: wcmc_conv2d_wgrad_group falls back to one round-1 launch per layer
int wcmc_wgrad_group_set(int which, int v) {
    if (which == 0) g_gg_pack = v;
    else if (which == 1) g_gg_enable = v;
    else return -1;
    return 0;
}

namespace {

struct GgPlan {
    int Ho, Wo;
    int nt, ci_tiles, m_tiles, cstride, tpg, groups, halo_w, box_h, b_plane, b_planes, stage_bytes, stages;
    int tiles_x, tiles_y, total_tiles, per_team;
    double tile_clk;   // tensor-pipe clocks of the slowest member per pixel tile
    int teams;
};

int gg_plan_layer(const wcmc_wgrad_layer& l, GgPlan* p) {
    const int k = l.ksize, taps = k * k;
    p->Ho = l.H + 2 * l.pad - k + 1;
    p->Wo = l.W + 2 * l.pad - k + 1;
    int nt = l.cin_p;
    if (nt > 128) {
        nt = 128;
        for (int c = 128; c >= 64; c -= 16)
            if (l.cin_p % c == 0) { nt = c; break; }
    }
    p->nt = nt;
    p->ci_tiles = (l.cin_p + nt - 1) / nt;
    p->m_tiles = (l.cout_p + 127) / 128;
    // columns per tap: a multiple of 16 is always safe; packing a whole kernel row at the logical channel count
    // (see the header comment) needs the padded channels to be zeros, which the NHWC convention guarantees
    int cstride = nt;
    int tpg = std::min(taps, 512 / cstride);
    if (g_gg_pack && p->ci_tiles == 1 && l.cin < nt && l.cin % 4 == 0 && tpg < k && (k - 1) * l.cin + nt <= 512) {
        cstride = l.cin;
        tpg = k;
    } else {
        // as few groups as TMEM allows, of equal size (the members of a team walk the same tiles: the largest
        // group sets the pace)
        const int groups = (taps + tpg - 1) / tpg;
        tpg = (taps + groups - 1) / groups;
    }
    p->cstride = cstride;
    p->tpg = tpg;
    p->groups = (taps + tpg - 1) / tpg;
    // kernel rows spanned by a group (a group that is not row aligned touches one more)
    int span = 1;
    for (int g = 0; g < p->groups; ++g) {
        const int t0 = g * tpg, t1 = std::min(taps, t0 + tpg) - 1;
        span = std::max(span, t1 / k - t0 / k + 1);
    }
    p->halo_w = kGgTileW + k - 1;
    p->box_h = kGgTileH + span - 1;
    p->b_plane = ((p->halo_w * p->box_h * 128 + 1023) / 1024) * 1024;
    p->b_planes = (nt + 63) / 64;
    p->stage_bytes = 2 * kGgAPlane + p->b_planes * p->b_plane;
    p->stages = std::max(2, std::min(kGgMaxStages, kGgSmemBudget / p->stage_bytes));
    p->tiles_x = (p->Wo + kGgTileW - 1) / kGgTileW;
    p->tiles_y = (p->Ho + kGgTileH - 1) / kGgTileH;
    p->total_tiles = l.N * p->tiles_x * p->tiles_y;
    p->per_team = p->m_tiles * p->ci_tiles * p->groups;
    // per pixel tile: taps x 4 MMAs of M128 x N(nt); the operand feed (4 KB of A + nt*32 B of B per MMA at 128 B/clk)
    // is a little above the tensor floor nt/2
    const double mma_clk = std::max(nt / 2.0, (4096.0 + nt * 32.0) / 128.0);
    p->tile_clk = std::min(p->tpg, taps) * (kGgTileH / 2) * mma_clk + 200.0;
    p->teams = 1;
    return 0;
}

// ---- chains, teams, segments --------------------------------------------------------------------------------
struct GgChainPlan {
    std::vector<int> layers;       // indices into the launch's layer list, in order
    int per_team = 0, teams = 1;
    long tiles = 0;                // pixel tiles of all its layers
    double tile_clk = 0.0;
};

bool gg_same_structure(const wcmc_wgrad_layer& a, const GgPlan& pa, const wcmc_wgrad_layer& b, const GgPlan& pb) {
    return a.ksize == b.ksize && a.cin == b.cin && a.cin_p == b.cin_p && a.cout_p == b.cout_p && pa.nt == pb.nt &&
           pa.cstride == pb.cstride && pa.tpg == pb.tpg && pa.stage_bytes == pb.stage_bytes && pa.stages == pb.stages;
}

double gg_chain_cost(const GgChainPlan& c) {
    const long tiles = (c.tiles + c.teams - 1) / c.teams;
    return tiles * c.tile_clk + 6000.0 * (1.0 + c.layers.size() / static_cast<double>(c.teams));   // + prologue / drains
}

// One launch: chains of structurally identical layers; the SMs are dealt out one team at a time to the chain whose
// CTAs currently carry the most work; each chain's tile line is then cut into equal contiguous pieces.
struct GgLaunchPlan {
    std::vector<GgChainPlan> chains;
    std::vector<std::vector<GgSeg>> segs;     // per team (global order: chain by chain)
    std::vector<int> slabs;                   // per layer: number of partial-sum slabs
    int ctas = 0;
};

int gg_plan_launch(const wcmc_wgrad_layer* layers, const std::vector<GgPlan>& plans, int first, int n, int sms,
                   GgLaunchPlan* out) {
    std::vector<GgChainPlan>& ch = out->chains;
    for (int k = 0; k < n; ++k) {
        const int i = first + k;
        int found = -1;
        for (size_t c = 0; c < ch.size(); ++c)
            if (gg_same_structure(layers[first + ch[c].layers[0]], plans[first + ch[c].layers[0]], layers[i], plans[i])) {
                found = static_cast<int>(c);
                break;
            }
        if (found < 0) {
            GgChainPlan c;
            c.per_team = plans[i].per_team;
            c.tile_clk = plans[i].tile_clk;
            ch.push_back(c);
            found = static_cast<int>(ch.size()) - 1;
        }
        ch[found].layers.push_back(k);
        ch[found].tiles += plans[i].total_tiles;
    }
    int used = 0;
    for (auto& c : ch) used += c.per_team;
    if (used > sms) return -1;
    for (;;) {
        int best = -1;
        double worst = 0.0;
        for (size_t i = 0; i < ch.size(); ++i) {
            if (ch[i].teams >= ch[i].tiles || used + ch[i].per_team > sms) continue;
            const double c = gg_chain_cost(ch[i]);
            if (c > worst) { worst = c; best = static_cast<int>(i); }
        }
        if (best < 0) break;
        double top = 0.0;
        for (const auto& c : ch) top = std::max(top, gg_chain_cost(c));
        if (gg_chain_cost(ch[best]) < 0.85 * top) break;     // the most loaded chain cannot be helped: stop adding slabs
        ch[best].teams += 1;
        used += ch[best].per_team;
    }
    out->ctas = used;
    out->slabs.assign(n, 0);
    for (auto& c : ch) {
        // a team that would cross more than kGgMaxSeg layers (many tiny layers, one team) cannot happen with the
        // shapes of the path; guard anyway by giving such a chain one team per layer at least
        for (int t = 0; t < c.teams; ++t) {
            const long b = c.tiles * t / c.teams, e = c.tiles * (t + 1) / c.teams;
            std::vector<GgSeg> sg;
            long base = 0;
            for (int k : c.layers) {
                const long nt = plans[first + k].total_tiles;
                const long lo = std::max(b, base), hi = std::min(e, base + nt);
                if (hi > lo) {
                    GgSeg s;
                    s.layer = k;
                    s.t0 = static_cast<int>(lo - base);
                    s.t1 = static_cast<int>(hi - base);
                    s.slab = out->slabs[k]++;
                    sg.push_back(s);
                }
                base += nt;
            }
            if (sg.empty() || sg.size() > static_cast<size_t>(kGgMaxSeg)) return -2;
            out->segs.push_back(sg);
        }
    }
    return 0;
}

int gg_check_layer(const wcmc_wgrad_layer& l, int i) {
    WCMC_REQUIRE(l.x != nullptr && l.dy != nullptr && l.dw != nullptr, WCMC_ESHAPE, "wgrad_group: layer %d: null pointer", i);
    WCMC_REQUIRE(l.ksize == 1 || l.ksize == 3 || l.ksize == 5, WCMC_ESHAPE, "wgrad_group: layer %d: ksize %d not in {1,3,5}", i, l.ksize);
    WCMC_REQUIRE(l.pad >= 0 && l.pad < l.ksize, WCMC_ESHAPE, "wgrad_group: layer %d: bad pad %d", i, l.pad);
    WCMC_REQUIRE(l.cin_p % 16 == 0 && l.cout_p % 16 == 0 && l.cin_p > 0 && l.cout_p > 0, WCMC_ESHAPE,
                 "wgrad_group: layer %d: cin_p (%d) / cout_p (%d) must be positive multiples of 16", i, l.cin_p, l.cout_p);
    WCMC_REQUIRE(l.x_cs % 8 == 0 && l.x_coff % 8 == 0 && l.dy_cs % 8 == 0 && l.dy_coff % 8 == 0, WCMC_ESHAPE,
                 "wgrad_group: layer %d: channel strides / offsets must be multiples of 8", i);
    WCMC_REQUIRE(l.cout <= l.cout_p && l.cin <= l.cin_p && l.cout > 0 && l.cin > 0, WCMC_ESHAPE,
                 "wgrad_group: layer %d: logical channels exceed padded", i);
    WCMC_REQUIRE(l.N > 0 && l.H + 2 * l.pad - l.ksize + 1 > 0 && l.W + 2 * l.pad - l.ksize + 1 > 0, WCMC_ESHAPE,
                 "wgrad_group: layer %d: empty output", i);
    return WCMC_OK;
}

size_t gg_layer_ws(const wcmc_wgrad_layer& l, int slabs) {
    return static_cast<size_t>(slabs) * l.ksize * l.ksize * l.cout_p * l.cin_p * sizeof(float);
}

// splits the layer list into launches of at most kGgMaxLayers layers whose one-team-per-chain footprint fits the SMs
struct GgLaunch { int first, n; GgLaunchPlan plan; };

int gg_chunks(const wcmc_wgrad_layer* layers, int n, std::vector<GgLaunch>* out, std::vector<GgPlan>* plans_out) {
    const int sms = std::min(wcmc_num_sms(), kGgMaxTeams);
    plans_out->resize(n);
    for (int i = 0; i < n; ++i) {
        gg_plan_layer(layers[i], &(*plans_out)[i]);
        if ((*plans_out)[i].per_team > sms) {
            wcmc_set_error("wgrad_group: layer %d needs %d CTAs per team (> %d SMs)", i, (*plans_out)[i].per_team, sms);
            return WCMC_ESHAPE;
        }
    }
    int i = 0;
    while (i < n) {
        int cnt = std::min(kGgMaxLayers, n - i);
        for (;;) {
            GgLaunch L;
            L.first = i;
            L.n = cnt;
            const int rc = gg_plan_launch(layers, *plans_out, i, cnt, sms, &L.plan);
            if (rc == 0) {
                out->push_back(L);
                break;
            }
            if (cnt == 1) {
                wcmc_set_error("wgrad_group: cannot plan layer %d (%d)", i, rc);
                return WCMC_ESHAPE;
            }
            cnt = (cnt + 1) / 2;     // fewer layers per launch
        }
        i += cnt;
    }
    return WCMC_OK;
}

}  // namespace

int wcmc_wgrad_group_enabled() { return g_gg_enable; }

extern "C" size_t wcmc_conv2d_wgrad_group_workspace(const wcmc_wgrad_layer* layers, int n) {
    if (layers == nullptr || n <= 0) return 0;
    if (!g_gg_enable) {
        size_t tot = 0;
        for (int i = 0; i < n; ++i) {
            const wcmc_wgrad_layer& l = layers[i];
            tot += (wcmc_conv2d_wgrad_workspace(l.N, l.H, l.W, l.cin_p, l.cout_p, l.ksize, l.pad) + 255) / 256 * 256;
        }
        return tot;
    }
    for (int i = 0; i < n; ++i)
        if (gg_check_layer(layers[i], i) != WCMC_OK) return 0;
    std::vector<GgLaunch> launches;
    std::vector<GgPlan> plans;
    if (gg_chunks(layers, n, &launches, &plans) != WCMC_OK) return 0;
    size_t tot = 0;
    for (const auto& L : launches)
        for (int k = 0; k < L.n; ++k) tot += (gg_layer_ws(layers[L.first + k], L.plan.slabs[k]) + 255) / 256 * 256;
    return tot;
}

extern "C" int wcmc_conv2d_wgrad_group(const wcmc_wgrad_layer* layers, int n, int dtype, void* workspace,
                                       size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(layers != nullptr && n > 0, WCMC_ESHAPE, "wgrad_group: no layers");
    WCMC_REQUIRE(dtype == WCMC_BF16 || dtype == WCMC_F16, WCMC_ESHAPE, "wgrad_group: dtype must be a 16-bit format");
    WCMC_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, WCMC_EALIGN,
                 "wgrad_group: workspace must be 256-byte aligned");
    if (!g_gg_enable) {   // round-1 path: one launch per layer, one reduction for all of them
        std::vector<wcmc_wgrad_reduce_desc> red(n);
        uint8_t* ws = static_cast<uint8_t*>(workspace);
        size_t off = 0;
        for (int i = 0; i < n; ++i) {
            const wcmc_wgrad_layer& l = layers[i];
            const size_t need = wcmc_conv2d_wgrad_workspace(l.N, l.H, l.W, l.cin_p, l.cout_p, l.ksize, l.pad);
            WCMC_REQUIRE(off + need <= workspace_bytes, WCMC_EWORKSPACE, "wgrad_group: workspace too small");
            int rc = wcmc_conv2d_wgrad_partial(l.x, dtype, l.N, l.H, l.W, l.x_cs, l.x_coff, l.cin_p, l.dy, dtype, l.dy_cs,
                                               l.dy_coff, l.cout_p, l.ksize, l.pad, l.dw, l.cout, l.cin, l.accumulate,
                                               l.scale, ws + off, need, &red[i], stream_);
            if (rc) return rc;
            off += (need + 255) / 256 * 256;
        }
        return wcmc_wgrad_reduce_batch(red.data(), n, stream_);
    }
    for (int i = 0; i < n; ++i) {
        int rc = gg_check_layer(layers[i], i);
        if (rc) return rc;
    }
    std::vector<GgLaunch> launches;
    std::vector<GgPlan> plans;
    int rc = gg_chunks(layers, n, &launches, &plans);
    if (rc) return rc;
    std::vector<wcmc_wgrad_reduce_desc> red(n);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    size_t off = 0;
    static thread_local GgParams P;     // 16 KB: kept off the stack; a launch copies it into the kernel parameters
    for (const auto& LN : launches) {
        P.n_layers = LN.n;
        int smem_need = 0;
        for (int k = 0; k < LN.n; ++k) {
            const int i = LN.first + k;
            const wcmc_wgrad_layer& l = layers[i];
            const GgPlan& p = plans[i];
            GgLayer& G = P.L[k];
            const int slabs = LN.plan.slabs[k];
            const size_t need = gg_layer_ws(l, slabs);
            WCMC_REQUIRE(off + need <= workspace_bytes, WCMC_EWORKSPACE, "wgrad_group: workspace too small (%zu < %zu)",
                         workspace_bytes, off + need);
            G.ws = reinterpret_cast<float*>(ws + off);
            off += (need + 255) / 256 * 256;
            {
                uint64_t dims[4] = {static_cast<uint64_t>(l.cout_p), static_cast<uint64_t>(p.Wo),
                                    static_cast<uint64_t>(p.Ho), static_cast<uint64_t>(l.N)};
                uint64_t strides[3] = {static_cast<uint64_t>(l.dy_cs) * 2, static_cast<uint64_t>(l.dy_cs) * 2 * p.Wo,
                                       static_cast<uint64_t>(l.dy_cs) * 2 * p.Wo * p.Ho};
                uint32_t box[4] = {64, kGgTileW, kGgTileH, 1};
                rc = wcmc_encode_tmap_bf16(&G.tmdy, static_cast<const __nv_bfloat16*>(l.dy) + l.dy_coff, 4, dims, strides,
                                           box, 1);
                if (rc) return rc;
            }
            {
                uint64_t dims[4] = {static_cast<uint64_t>(l.cin_p), static_cast<uint64_t>(l.W), static_cast<uint64_t>(l.H),
                                    static_cast<uint64_t>(l.N)};
                uint64_t strides[3] = {static_cast<uint64_t>(l.x_cs) * 2, static_cast<uint64_t>(l.x_cs) * 2 * l.W,
                                       static_cast<uint64_t>(l.x_cs) * 2 * l.W * l.H};
                uint32_t box[4] = {64, static_cast<uint32_t>(p.halo_w), static_cast<uint32_t>(p.box_h), 1};
                rc = wcmc_encode_tmap_bf16(&G.tmx, static_cast<const __nv_bfloat16*>(l.x) + l.x_coff, 4, dims, strides, box,
                                           1);
                if (rc) return rc;
            }
            G.tiles_x = p.tiles_x; G.tiles_y = p.tiles_y; G.total_tiles = p.total_tiles;
            G.ksize = l.ksize; G.pad = l.pad; G.taps = l.ksize * l.ksize;
            G.cin_p = l.cin_p; G.cout_p = l.cout_p;
            G.nt = p.nt; G.ci_tiles = p.ci_tiles; G.m_tiles = p.m_tiles;
            G.cstride = p.cstride; G.tpg = p.tpg; G.groups = p.groups;
            G.halo_w = p.halo_w; G.box_h = p.box_h; G.b_plane = p.b_plane; G.b_planes = p.b_planes;
            G.stage_bytes = p.stage_bytes; G.stages = p.stages;
            G.dtype = dtype;
            smem_need = std::max(smem_need, p.stage_bytes * p.stages);
            wcmc_wgrad_reduce_desc& d = red[i];
            d.ws = G.ws; d.dw = l.dw; d.scale = l.scale;
            d.nsplit = slabs; d.nsplit_b = slabs; d.taps_a = 0;
            d.cout = l.cout; d.cin = l.cin; d.taps = G.taps; d.cout_p = l.cout_p; d.cin_p = l.cin_p;
            d.accumulate = l.accumulate;
        }
        P.n_chains = static_cast<int>(LN.plan.chains.size());
        int cta = 0, team = 0;
        for (int c = 0; c < P.n_chains; ++c) {
            const GgChainPlan& cp = LN.plan.chains[c];
            P.C[c].first_cta = cta;
            P.C[c].per_team = cp.per_team;
            P.C[c].teams = cp.teams;
            P.C[c].first_team = team;
            cta += cp.per_team * cp.teams;
            team += cp.teams;
        }
        WCMC_REQUIRE(team <= kGgMaxTeams && team == static_cast<int>(LN.plan.segs.size()), WCMC_ESHAPE,
                     "wgrad_group: internal plan error (%d teams)", team);
        for (int t = 0; t < team; ++t) {
            const auto& sg = LN.plan.segs[t];
            P.nseg[t] = static_cast<unsigned char>(sg.size());
            for (size_t k = 0; k < sg.size(); ++k) P.S[t][k] = sg[k];
        }
        const int smem_bytes = smem_need + 1024;
        WCMC_FUNC_SMEM(conv_wgrad_group_kernel, kGgSmemBudget + 1024);
        WCMC_LAUNCH(conv_wgrad_group_kernel, cta, kGgThreads, smem_bytes, stream, P);
        WCMC_LAUNCH_CHECK();
    }
    return wcmc_wgrad_reduce_batch(red.data(), n, stream_);
}

// Plan for tools / tests (host only): per layer the number of partial-sum slabs (= K-split teams that touch it), the
// CTAs of its chain, taps per group and the TMEM column stride.
extern "C" int wcmc_conv2d_wgrad_group_plan(const wcmc_wgrad_layer* layers, int n, int* teams_out, int* ctas_out,
                                            int* tpg_out, int* cstride_out) {
    WCMC_REQUIRE(layers != nullptr && n > 0, WCMC_ESHAPE, "wgrad_group_plan: no layers");
    for (int i = 0; i < n; ++i) {
        int rc = gg_check_layer(layers[i], i);
        if (rc) return rc;
    }
    std::vector<GgLaunch> launches;
    std::vector<GgPlan> plans;
    int rc = gg_chunks(layers, n, &launches, &plans);
    if (rc) return rc;
    for (const auto& L : launches) {
        for (const auto& c : L.plan.chains)
            for (int k : c.layers)
                if (ctas_out) ctas_out[L.first + k] = c.teams * c.per_team;
        for (int k = 0; k < L.n; ++k) {
            const int i = L.first + k;
            if (teams_out) teams_out[i] = L.plan.slabs[k];
            if (tpg_out) tpg_out[i] = plans[i].tpg;
            if (cstride_out) cstride_out[i] = plans[i].cstride;
        }
    }
    return static_cast<int>(launches.size());
}
