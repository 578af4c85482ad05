/* libwcmc.so -- C ABI of the B200-native (sm_100a) backend for WCMC's KPCN hot path.
 *
 * The reference (Mephisto405/WCMC) is pure Python; its only native boundary on this path is the
 * Halide `kernel_weighting` extension of the un-vendored `sbmc` package plus the cuDNN / ATen
 * calls behind torch.nn.  This header is the boundary that replaces them: every
 * entry point cites the reference call site whose arithmetic it takes over.  The Python host
 * side (wcmc_b200/, loaded with ctypes) mirrors sbmc.KPCN / sbmc.modules / support.networks /
 * support.losses / support.interfaces on top of it.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless named host_*;
 *   - the library never allocates, never synchronises and only touches the stream passed in
 *     (`stream` is a cudaStream_t passed as void*);
 *   - return value 0 = WCMC_OK, negative = error; wcmc_last_error() gives the message
 *     (thread local);
 *   - "NHWC bf16" = activations stored (N, H, W, Cs) in bfloat16 with a channel stride Cs that
 *     is a multiple of 8 and a logical channel count padded with zeros to a multiple of 16;
 *   - sm_100a only: wcmc_init() fails with WCMC_EARCH elsewhere.  There is no CPU fallback.
 */
#ifndef WCMC_H_
#define WCMC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WCMC_OK 0
#define WCMC_ESHAPE (-1)  /* unsupported / inconsistent shape argument */
#define WCMC_EALIGN (-2)  /* pointer or stride alignment */
#define WCMC_EARCH (-3)   /* device is not sm_100 */
#define WCMC_ECUDA (-4)   /* CUDA runtime / driver error (message has the details) */
#define WCMC_EWORKSPACE (-5)

/* storage types of activations / packed weights / gradients */
#define WCMC_BF16 0
#define WCMC_F16 1
#define WCMC_F32 2

#define WCMC_ACT_LINEAR 0
#define WCMC_ACT_RELU 1
#define WCMC_ACT_LEAKY 2

const char* wcmc_last_error(void);
const char* wcmc_version(void);
/* Checks that `device` is a compute-capability-10.x GPU and prepares per-process state. */
int wcmc_init(int device);
/* Tuning knobs for the micro-benchmarks under tools/ ("ka_tile_w" = 16 | 32: kernel-apply tile width). */
int wcmc_tuning_set(const char* name, int value);

/* ---- layout conversion (boundary between torch NCHW fp32 tensors and the NHWC bf16 pipeline) --
 * dst[n,h,w,dst_coff+c] = half(src[n,c,h,w]) for c < C; channels C..c_fill-1 are written as 0.
 * dst_dtype / src_dtype: WCMC_BF16 or WCMC_F16.  `scale` (device pointer to one float, may be
 * NULL) multiplies every value: the backward pass keeps its 16-bit gradients multiplied by a
 * per-call loss scale s (s on the way in, 1/s on the way out) so fp16 gradients never underflow. */
int wcmc_nchw_f32_to_nhwc(const float* src, void* dst, int dst_dtype, int N, int C, int H, int W,
                          int dst_cs, int dst_coff, int c_fill, const float* scale, void* stream);
/* dst[n,c,h,w] (fp32, contiguous) (+)= src[n,h,w,src_coff+c] */
int wcmc_nhwc_to_nchw_f32(const void* src, int src_dtype, float* dst, int N, int C, int H, int W,
                          int src_cs, int src_coff, int accumulate, const float* scale, void* stream);

/* ---- weights: torch (Cout, Cin, k, k) fp32  ->  packed bf16 operands of the conv kernels ------
 * fwd  : dst[co][ky*k+kx][ci]            = w[co][ci][ky][kx]   (cout_p x k*k x cin_p, zero padded)
 * dgrad: dst[ci][(k-1-ky)*k+(k-1-kx)][co] = w[co][ci][ky][kx]  (cin_p  x k*k x cout_p)           */
/* dst_bias[cout_p] = bias zero padded (bias may be NULL -> zeros); any dst may be NULL.        */
int wcmc_pack_weights(const float* w, const float* bias, void* dst_fwd, void* dst_dgrad,
                      float* dst_bias, int dtype, int cout, int cin, int ksize, int cout_p, int cin_p,
                      void* stream);

/* All layers of a conv chain in one launch. */
#define WCMC_PACK_BATCH_MAX 24
typedef struct {
    const float* w;     /* (cout, cin, k, k) fp32 */
    const float* bias;  /* (cout) fp32 or NULL */
    void* dst_fwd;      /* [cout_p][k*k][cin_p] or NULL */
    void* dst_dgrad;    /* [cin_p][k*k][cout_p] or NULL */
    float* dst_bias;    /* [cout_p] or NULL */
    int cout, cin, ksize, cout_p, cin_p;
} wcmc_pack_desc;
int wcmc_pack_weights_batch(const wcmc_pack_desc* host_descs, int n, int dtype, void* stream);

/* ---- K1/K2: convolution forward / data gradient (tcgen05 implicit GEMM) -----------------------
 * Replaces nn.Conv2d(+ReLU) inside sbmc.modules.ConvChain (used at
 * /root/reference/support/networks.py:18-24 and by sbmc.KPCN, train_kpcn.py:213).
 *   y = act(conv(x, w) + bias) [* (mask > 0 ? 1 : slope)]
 * x: NHWC bf16 (N,H,W,x_cs), channels [x_coff, x_coff+cin_p); w_packed: see wcmc_pack_weights;
 * x_dtype / w_dtype: WCMC_BF16 or WCMC_F16 and equal (tcgen05.mma kind::f16 traps on mixed formats);
 * bias: fp32[cout_p] or NULL; y: NHWC (N,Ho,Wo,y_cs) in y_dtype (bf16 / f16 / f32), channels
 * [y_coff, y_coff+cout_p), Ho = H + 2*pad - ksize + 1.  mask (optional, NHWC bf16 with the
 * spatial size of y) fuses the activation derivative of the previous layer into a dgrad.
 * colsum (optional, fp32[cout_p], caller zero-initialises): colsum[c] += *colsum_scale * sum over all
 * output pixels of y[.., c] -- a data-gradient launch thereby also produces the bias gradient of
 * the layer below it.
 * flags: 0 for production (test knobs: bits 4-5 force m tiles per region, bits 8-15 force the
 * n tile; tuning knobs that give WRONG results, used by tools/conv_bench.py to find the bound:
 * bit 16 weights loaded once, bit 17 halos loaded once, bit 18 epilogue skipped.  Launch-shape knobs with
 * identical results: bit 20 forces / bit 21 forbids the CTA-pair launch (cluster of 2, tcgen05
 * cta_group::2, M = 256), bit 22 forbids weight stages that carry a whole kernel row of taps).             */
int wcmc_conv2d(const void* x, int x_dtype, int N, int H, int W, int x_cs, int x_coff, int cin_p,
                const void* w_packed, int w_dtype, int cout_p, const float* bias, int ksize, int pad,
                void* y, int y_dtype, int y_cs, int y_coff, int act, const void* mask, int mask_cs,
                int mask_coff, float slope, float* colsum, const float* colsum_scale, int flags,
                void* stream);

/* the LAST layer of a KPCN branch (5x5, 100 -> 441) fused with softmax + the 21x21 kernel-apply
 * (sbmc.KPCN.forward: ConvChain -> KernelApply, called at /root/reference/support/interfaces.py:310): same
 * implicit GEMM, but the epilogue keeps a running softmax over the 441 logits of its pixels across the four n tiles
 * and gathers the radiance neighbourhood from a shared-memory tile, so the logits (1764 B per pixel, written once
 * and read once by the unfused pair of launches) never reach HBM.  Inference only (no statistics for a backward).
 *   out[n,c,y,x] = sum_k softmax_k(conv(x, w)[n,y,x,:] + bias) * data0[n,c,y+k/21-10,x+k%21-10]
 * data, out: (N,3,Ho,Wo) fp32 NCHW (data zero outside).  cout_p must be 448, ksize 5, ka_ksize 21.            */
int wcmc_conv2d_kernel_apply(const void* x, int x_dtype, int N, int H, int W, int x_cs, int x_coff, int cin_p,
                             const void* w_packed, int w_dtype, int cout_p, const float* bias, int ksize, int pad,
                             const float* data, float* out, int ka_ksize, int flags, void* stream);

/* ---- K3: convolution weight gradient (tcgen05, MN-major operands, split-K) --------------------
 * Autograd of nn.Conv2d w.r.t. its weight (reference: `L_diffuse.backward()`,
 * /root/reference/support/interfaces.py:237-238).
 *   dw[co][ci][ky][kx] (torch layout, fp32) (+)= sum_{n,oy,ox} dy[n,oy,ox,co] * x[n,oy+ky-pad,ox+kx-pad,ci]
 * x, dy: NHWC bf16 as for wcmc_conv2d (dy has the conv's OUTPUT spatial size).  workspace holds
 * the split-K partial sums; size it with wcmc_conv2d_wgrad_workspace().  `scale` (device float
 * or NULL) multiplies the result (un-does the loss scale carried by dy).                        */
size_t wcmc_conv2d_wgrad_workspace(int N, int H, int W, int cin_p, int cout_p, int ksize, int pad);
int wcmc_conv2d_wgrad(const void* x, int x_dtype, int N, int H, int W, int x_cs, int x_coff, int cin_p,
                      const void* dy, int dy_dtype, int dy_cs, int dy_coff, int cout_p, int ksize, int pad,
                      float* dw, int cout, int cin, int accumulate, const float* scale, void* workspace,
                      size_t workspace_bytes, void* stream);
/* The same in two steps, so that one launch finalises many layers: `_partial` runs the tensor-core
 * kernel only (split-K partial sums stay in `workspace`, which must stay alive and un-aliased until
 * the reduction) and fills *desc_out; wcmc_wgrad_reduce_batch sums the splits of n layers, transposes
 * to torch's (cout,cin,k,k) layout with full-sector stores and applies accumulate / scale.        */
typedef struct {
    const float* ws;     /* partial sums: [nsplit][taps_a][cout_p][cin_p] for taps < taps_a, followed by
                            [nsplit_b][taps - taps_a][cout_p][cin_p] for the last (short) tap group */
    float* dw;           /* (cout, cin, k, k) fp32 */
    const float* scale;  /* device float or NULL */
    int nsplit, cout, cin, taps, cout_p, cin_p, accumulate;
    int nsplit_b, taps_a;
} wcmc_wgrad_reduce_desc;
#define WCMC_WGRAD_BATCH_MAX 32
int wcmc_conv2d_wgrad_partial(const void* x, int x_dtype, int N, int H, int W, int x_cs, int x_coff, int cin_p,
                              const void* dy, int dy_dtype, int dy_cs, int dy_coff, int cout_p, int ksize,
                              int pad, float* dw, int cout, int cin, int accumulate, const float* scale,
                              void* workspace, size_t workspace_bytes, wcmc_wgrad_reduce_desc* desc_out,
                              void* stream);
int wcmc_wgrad_reduce_batch(const wcmc_wgrad_reduce_desc* host_descs, int n, void* stream);
/* Weight gradients of EVERY layer of a backward pass in one call (round 2; conv_wgrad_group.cu): the SMs are dealt
 * out to the layers in proportion to their cost, each layer gets a few K-split teams instead of ~21 splits per tap
 * group, one batched reduction follows.  Same arithmetic and result layout as n wcmc_conv2d_wgrad calls; all layers
 * share one 16-bit format `dtype`.  host_layers is a HOST array (n <= 64); x / dy / dw / scale are device pointers
 * that must stay alive until the stream has run the call.  workspace: wcmc_conv2d_wgrad_group_workspace(layers, n)
 * bytes, 256-byte aligned.                                                                                    */
typedef struct {
    const void* x;       /* NHWC 16-bit (N,H,W,x_cs), channels [x_coff, x_coff + cin_p), padded channels zero */
    const void* dy;      /* NHWC 16-bit (N,Ho,Wo,dy_cs), channels [dy_coff, dy_coff + cout_p) */
    float* dw;           /* (cout, cin, k, k) fp32 */
    const float* scale;  /* device float or NULL: multiplies the result */
    int N, H, W, x_cs, x_coff, cin_p, cin;
    int dy_cs, dy_coff, cout_p, cout;
    int ksize, pad, accumulate;
} wcmc_wgrad_layer;
size_t wcmc_conv2d_wgrad_group_workspace(const wcmc_wgrad_layer* host_layers, int n);
int wcmc_conv2d_wgrad_group(const wcmc_wgrad_layer* host_layers, int n, int dtype, void* workspace,
                            size_t workspace_bytes, void* stream);
/* Host-only: the launch plan (K-split teams, CTAs, taps per group, TMEM column stride per layer; any output may be
 * NULL); returns the number of kernel launches or a negative error.                                        */
int wcmc_conv2d_wgrad_group_plan(const wcmc_wgrad_layer* host_layers, int n, int* teams_out, int* ctas_out,
                                 int* tpg_out, int* cstride_out);
/* db[co] (+)= scale * sum over all pixels of dy[pix][dy_coff+co]   (scale: device float or NULL) */
int wcmc_bias_grad(const void* dy, int dy_dtype, int npix, int dy_cs, int dy_coff, int cout, float* db,
                   int accumulate, const float* scale, void* stream);

/* ---- K4/K5: softmax + 21x21 kernel-apply (replaces sbmc.modules.KernelApply and the Halide
 * `kernel_weighting` op; called by sbmc.KPCN.forward) ------------
 * logits: NHWC fp32 (N,H,W,l_cs), first k*k channels used (k = 21: l_cs = 448)
 * data  : NCHW fp32 (N,C,H,W), C <= 4;   out: NCHW fp32 (N,C,H,W)
 *   out[n,c,y,x] = sum_{dy,dx} softmax(logits[n,y,x,:])[dy*k+dx] * data0[n,c,y+dy-k/2,x+dx-k/2]
 * stats (N,H,W,2) fp32 receives (row max, 1/sum exp) for the backward pass (may be NULL).   */
int wcmc_kernel_apply_fwd(const float* logits, int l_cs, const float* data, float* out,
                          float* stats, int N, int C, int H, int W, int ksize, void* stream);
/* d_logits[n,y,x,k] = scale * p_k * (sum_c g_c v_{c,k} - sum_c g_c out_c); written as NHWC
 * (N,H,W,dl_cs) in dl_dtype (bf16 / f16 / f32), channels k*k..dl_cs-1 zeroed.                  */
int wcmc_kernel_apply_bwd(const float* logits, int l_cs, const float* data, const float* out,
                          const float* stats, const float* grad_out, void* d_logits, int dl_cs,
                          int dl_dtype, int N, int C, int H, int W, int ksize, const float* scale,
                          void* stream);

/* ---- NHWC bf16 glue of the path-embedding network (PathNet, /root/reference/support/
 * networks.py:29-42; U-Net = sbmc.modules.Autoencoder) ------------------
 * Every tensor is (pixels, channel stride cs, channel offset coff); C, cs, coff multiples of 8.
 * `dtype` (WCMC_BF16 / WCMC_F16) is the storage type of every 16-bit tensor of the call
 * (activations and gradients use the same type: the tensor cores do not mix f16 with bf16).
 * Writing into a channel slice of the consumer's buffer replaces torch.cat.                     */
/* y (N,H/2,W/2) = 2x2 max pool of x (N,H,W)            (nn.MaxPool2d(2,2)) */
int wcmc_maxpool2_fwd(const void* x, int x_cs, int x_coff, void* y, int y_cs, int y_coff, int N, int H,
                      int W, int C, int dtype, void* stream);
/* dx (N,H,W) = (add or 0) + dy routed to the first maximum of each window (x = forward input) */
int wcmc_maxpool2_bwd(const void* x, int x_cs, int x_coff, const void* dy, int dy_cs, int dy_coff,
                      const void* add, int add_cs, int add_coff, void* dx, int dx_cs, int dx_coff, int N,
                      int H, int W, int C, int dtype, void* stream);
/* y (N,2h,2w) = F.interpolate(x (N,h,w), scale 2, mode='bilinear', align_corners=False) */
int wcmc_upsample2_fwd(const void* x, int x_cs, int x_coff, void* y, int y_cs, int y_coff, int N, int h,
                       int w, int C, int dtype, void* stream);
int wcmc_upsample2_bwd(const void* dy, int dy_cs, int dy_coff, void* dx, int dx_cs, int dx_coff, int N,
                       int h, int w, int C, int dtype, void* stream);
/* y[b,hw,:] = scale * sum_s x[b,s,hw,:]      (mean over samples per pixel, networks.py:36) */
int wcmc_spp_reduce(const void* x, int x_cs, int x_coff, void* y, int y_cs, int y_coff, int B, int S,
                    int HW, int C, float scale, int dtype, void* stream);
/* y[b,s,hw,:] = (add or 0)[b,s,hw,:] + scale * x[b,hw,:]   (broadcast over samples, networks.py:39) */
int wcmc_spp_broadcast(const void* x, int x_cs, int x_coff, const void* add, int add_cs, int add_coff,
                       void* y, int y_cs, int y_coff, int B, int S, int HW, int C, float scale,
                       int dtype, void* stream);
/* dz = dy * act'(y), y = activation output (act = WCMC_ACT_RELU / WCMC_ACT_LEAKY) */
int wcmc_act_bwd(const void* dy, int dy_cs, int dy_coff, const void* y, int y_cs, int y_coff, void* dz,
                 int dz_cs, int dz_coff, long npix, int C, int act, float slope, int dtype, void* stream);

/* ---- K10: path-disentangling loss, permutation-paired form (FeatureMSE,
 * /root/reference/support/losses.py:33-61 intra_patch_dist / intra_batch_dist, :82-113 forward;
 * the displacement vectors also feed GlobalRelativeSimilarityLoss, :185-211) -----------------
 * p  : strided fp32 view (B,S,C,H,W) of the p-buffer (element strides p_sb,p_ss,p_sc,p_sh; x stride 1)
 * ref: strided fp32 view (B,3,H,W) of the target radiance (strides r_sb,r_sc,r_sh); it is
 *      tone-mapped inside: t = (max(ref,0)/(1+max(ref,0)))^0.454545 (losses.py:63-65)
 * idx_patch: int64[S*H*W] permutation shared by every batch item (losses.py:35);
 * idx_batch: int64[B*S*H*W] permutation of all rows (losses.py:50), NULL for non_local=False.
 * Row order is (s,y,x) inside a batch item, (b,s,y,x) globally (losses.py:88-93).
 *   e_patch[b*n+i] = 1/2|P_i-P_j|^2 - 1/2|t_i-t_j|^2, j = idx_patch[i] (same b);  e_batch likewise
 *   loss[0] = 1/2 mean(e_patch^2), loss[1] = 1/2 mean(e_batch^2)   (fp32[2])
 *   inv_patch int32[n], inv_batch int32[B*n] = inverse permutations (consumed by the backward)
 *   *nonfinite |= 1 when any P or t value is not finite (losses.py:99-102); caller zero-initialises.
 * The indices must be permutations (the reference always draws torch.randperm).              */
size_t wcmc_fmse_perm_workspace(int B, int S, int H, int W);
int wcmc_fmse_perm_fwd(const float* p, long p_sb, long p_ss, long p_sc, long p_sh, const float* ref,
                       long r_sb, long r_sc, long r_sh, const int64_t* idx_patch, const int64_t* idx_batch,
                       int B, int S, int C, int H, int W, float* e_patch, float* e_batch, int32_t* inv_patch,
                       int32_t* inv_batch, float* loss, int* nonfinite, void* workspace,
                       size_t workspace_bytes, void* stream);
/* dp (B,S,C,H,W) contiguous fp32 = sum over modes of
 *   g*coef*( w[i] (P_i - P_idx(i)) + w[inv(i)] (P_i - P_inv(i)) ),  g = *scale (device float) or 1.
 * FeatureMSE: w = e, coef = 1/(B*n), scale = upstream gradient.  GRS: w = dL/de, coef = 1.     */
/* the same, writing d p through element strides (batch, sample, channel, row; unit x stride): the gradient lands in
 * the interior of a zero-filled tensor of the UNcropped p-buffer's shape, so no slice-backward copies are needed */
int wcmc_fmse_perm_bwd_strided(const float* p, long p_sb, long p_ss, long p_sc, long p_sh, const int64_t* idx_patch,
                               const int64_t* idx_batch, const int32_t* inv_patch, const int32_t* inv_batch,
                               const float* w_patch, const float* w_batch, const float* scale, float coef_patch,
                               float coef_batch, int B, int S, int C, int H, int W, float* dp, long d_sb, long d_ss,
                               long d_sc, long d_sh, void* stream);
int wcmc_fmse_perm_bwd(const float* p, long p_sb, long p_ss, long p_sc, long p_sh, const int64_t* idx_patch,
                       const int64_t* idx_batch, const int32_t* inv_patch, const int32_t* inv_batch,
                       const float* w_patch, const float* w_batch, const float* scale, float coef_patch,
                       float coef_batch, int B, int S, int C, int H, int W, float* dp, void* stream);

/* ---- K6/K7: the per-sample 1x1 MLPs of the path-embedding network, one kernel each
 * (PathNet.embedding / PathNet.final, /root/reference/support/networks.py:18-19, :23-24, :29-42) ----
 * K6: paths (B,S,Cin,HW) fp32 (the dataset's layout) -> act1(W1 x + b1) -> act2(W2 . + b2) ->
 *     act3(W3 . + b3), all 64 wide.  w*: 16-bit packed [64][cin_p] / [64][64] (wcmc_pack_weights,
 *     fwd form); b*: fp32[64].  Outputs, all 16-bit NHWC over B*S*HW pixels (optional ones may be NULL):
 *       emb  (.., emb_cs) channels [emb_coff, emb_coff+64)   layer-3 output
 *       mean (B*HW, mean_cs) channels [mean_coff, +64)       mean of emb over the S samples (networks.py:36)
 *       x16, h1, h2 (.., 64)                                 input (zero padded to 64 channels) / hidden
 *                                                            activations for the backward pass
 *     Requires H*W % 4 == 0 (TMA row stride).  16-bit outputs leave by TMA store.                 */
int wcmc_pathnet_embed_fwd(const float* paths, int B, int S, int Cin, int HW, const void* w1, const float* b1,
                           const void* w2, const float* b2, const void* w3, const float* b3, int cin_p,
                           int dtype, int act1, int act2, int act3, float slope, void* x16, void* h1, void* h2,
                           void* emb, int emb_cs, int emb_coff, void* mean, int mean_cs, int mean_coff,
                           void* stream);
/* K7: out (B,S,outc,HW) fp32 = act2(W2 act1(W1 [emb_s | prop] + b1) + b2): emb (B*S*HW px, 64 ch) per
 *     sample, prop (B*HW px, 64 ch) per pixel (the reference repeats it S times and concatenates,
 *     networks.py:39-40).  w1: [128][128], w2: [outc_p][128] packed 16-bit; outc_p <= 32.
 *     h (optional): (B*S*HW, 128) 16-bit hidden activation for the backward pass.               */
int wcmc_pathnet_final_fwd(const void* emb, int emb_cs, int emb_coff, const void* prop, int prop_cs,
                           int prop_coff, const void* w1, const float* b1, const void* w2, const float* b2,
                           int outc, int outc_p, int dtype, int act1, int act2, float slope, void* h,
                           float* out, int B, int S, int HW, void* stream);

/* ---- batched weight normalisation (torch._weight_norm over dim 0 for every convolution of a network in one
 * launch; sbmc.modules.ConvChain's `weight_norm=True` default, kept by networks.py:18-24) ----------------------
 * forward (backward = 0): w[r,:] = g[r] * v[r,:] / ||v[r,:]||, norm[r] = ||v[r,:]|| saved.
 * backward (backward = 1): dg[r] = <dw[r,:], v[r,:]> / norm[r];  dv = (g/norm) (dw - v <dw,v> / norm^2).
 * All tensors fp32, row-major (rows = output channels, cols = cin*k*k).                                       */
#define WCMC_WN_BATCH_MAX 32
typedef struct {
    const float* v;
    const float* g;
    float* w;          /* forward output */
    float* norm;       /* forward output, backward input */
    const float* dw;   /* backward input */
    float* dv;         /* backward outputs */
    float* dg;
    int rows, cols;
} wcmc_wn_desc;
int wcmc_weight_norm_batch(const wcmc_wn_desc* host_descs, int n, int backward, void* stream);

/* ---- K8 / K9: the backward passes of the two MLPs, each ONE tensor-core kernel + a slab reduction
 * (autograd of networks.py:29-42; replaces the 1x1-conv dgrad / wgrad / bias / activation launches) ----
 * Gradients between kernels stay 16-bit and loss-scaled: *gscale multiplies the incoming fp32 gradient,
 * *inv_scale multiplies the fp32 parameter gradients on the way out (either may be NULL = 1).
 * wcmc_pathnet_bwd_workspace(which): bytes of per-CTA partial slabs, which = 0 (final) / 1 (embed).
 * K8: g, out (B,S,outc,HW) fp32 (dL/dout and the forward output); emb (B*S*HW, emb_cs) channels [0,64),
 *     prop (B*HW, prop_cs) channels [0,64), hfin (B*S*HW,128) 16-bit; w1t [128 cin][128 cout],
 *     w2t [128 cin][outc_p cout] = the data-gradient packing of wcmc_pack_weights.  Outputs: d_emb
 *     (B*S*HW,64), d_prop (B*HW,64) = sum over spp, 16-bit; dw1 (128,128), db1 (128), dw2 (outc,128),
 *     db2 (outc) fp32 (torch layout, overwritten).
 * K9: d_emb as above, d_red (B*HW,64) gradient of the spp mean (NULL = none; the kernel applies 1/S),
 *     saved activations emb, h2, h1, x16 (64 channels each; x16 zero padded beyond cin), w3t / w2t
 *     [64 cin][64 cout].  Outputs dw3, dw2 (64,64), dw1 (64,cin), db3, db2, db1 (64) fp32.          */
size_t wcmc_pathnet_bwd_workspace(int which);
int wcmc_pathnet_final_bwd(const float* g, const float* out, const float* gscale, const float* inv_scale,
                           const void* emb, int emb_cs, const void* prop, int prop_cs, const void* hfin,
                           const void* w1t, const void* w2t, int outc, int outc_p, int dtype, int act1,
                           int act2, float slope, void* d_emb, void* d_prop, float* dw1, float* db1,
                           float* dw2, float* db2, int B, int S, int HW, void* workspace,
                           size_t workspace_bytes, void* stream);
int wcmc_pathnet_embed_bwd(const void* d_emb, const void* d_red, const float* inv_scale, const void* emb,
                           int emb_cs, const void* h2, const void* h1, const void* x16, const void* w3t,
                           const void* w2t, int cin, int dtype, int act1, int act2, int act3, float slope,
                           float* dw3, float* db3, float* dw2, float* db2, float* dw1, float* db1, int B,
                           int S, int HW, void* workspace, size_t workspace_bytes, void* stream);

/* ---- preprocessing of raw OptaGen sample buffers (DenoiseDataset._preprocess_kpcn /
 * _preprocess_llpm, /root/reference/support/datasets.py:487-582, :301-361, NaN clamp :621-624 folded in) ----
 * raw: (H,W,S,104) fp32, S <= 8.  out44: (H,W,44) fp32 in the reference's channel order; out37: (H*W*S,37) fp32.
 * Written after round 1's GPU budget was spent: compiled, not yet validated on a GPU, not on the product path.   */
size_t wcmc_preprocess_kpcn_workspace(int H, int W);
int wcmc_preprocess_kpcn(const float* raw, int H, int W, int S, float* out44, void* workspace,
                         size_t workspace_bytes, void* stream);
int wcmc_preprocess_llpm(const float* raw, long nrows, float* out37, void* stream);

/* ---- K12: clip_grad_value_ + Adam for all parameter tensors in one launch
 * (/root/reference/support/interfaces.py:261 clip, :269-271 three optimiser steps; Adam with torch's
 * defaults as constructed at /root/reference/train_kpcn.py:277) -----------------------------------
 * dev_tensors: DEVICE array of descriptors; dev_blocks: DEVICE array of nblocks (tensor index,
 * chunk index) pairs, one per CTA, chunk = wcmc_adam_chunk() elements; *dev_step = number of steps
 * taken so far (the kernel uses t = *dev_step + 1 for the bias corrections and the call increments
 * it); dev_ok_flag (optional): when it points at 0 nothing is modified (non-finite loss: the host
 * raises after its one sync of the step).  clip <= 0 disables clipping; otherwise the clipped
 * gradient is written back, as clip_grad_value_ does.                                            */
typedef struct {
    float* p;   /* parameter */
    float* g;   /* gradient */
    float* m;   /* exp_avg */
    float* v;   /* exp_avg_sq */
    long n;
    float lr, beta1, beta2, eps;
} wcmc_adam_tensor;
int wcmc_adam_chunk(void);
/* dev_nonfinite_count (optional, device uint64): gradient elements that were NaN / inf -- an fp16 overflow in the 16-bit
 * backward pass -- are skipped (parameter and moments untouched) and counted here instead of poisoning the weights.   */
int wcmc_adam_clip_step(const wcmc_adam_tensor* dev_tensors, const int* dev_blocks, int nblocks, int* dev_step,
                        const int* dev_ok_flag, float clip, unsigned long long* dev_nonfinite_count, void* stream);

/* ---- K9: p-buffer statistics + concatenation (/root/reference/support/interfaces.py:165-180) ----------------
 * out (B, Cin+cr+1, HW) = cat[kpcn_in (B,Cin,HW), mean_S p[:, :, c0:c0+cr], var_S(p[:, :, c0:c0+cr]).mean(C) / S]
 * with p (B,S,C,HW) fp32 contiguous, var unbiased (torch.var).  The backward writes the whole d p (B,S,C,HW):
 * d p[b,s,c] = grad_out[b, Cin + c - c0] / S inside the channel range, 0 outside (the variance is detached).      */
int wcmc_pbuffer_concat_fwd(const float* kpcn_in, const float* p, float* out, int B, int S, int C, int c0, int cr,
                            int Cin, int HW, void* stream);
int wcmc_pbuffer_concat_bwd(const float* grad_out, float* dp, int B, int S, int C, int c0, int cr, int Cin, int HW,
                            void* stream);

/* ---- K12: radiance recombination (sbmc.KPCN.forward) and the image losses of the step ----------------------------
 * radiance = albedo * r_d + exp(r_s) - 1; r_d, r_s, radiance (B,3,h,w) fp32 contiguous; albedo is read through
 * strides (batch, channel, row; unit x stride) at the crop origin the pointer already includes.                     */
int wcmc_recombine(const float* albedo, long a_sb, long a_sc, long a_sh, const float* r_d, const float* r_s,
                   float* radiance, int B, int h, int w, void* stream);
/* sums4 = { mean|r_d - t_d|, mean|r_s - t_s|, mean|radiance - t_t|, 0.5 mean((radiance - t_t)^2 / (t_t^2 + eps)) }
 * (nn.L1Loss x3, /root/reference/train_kpcn.py:300-302; RelativeMSE, support/losses.py:255-264) in ONE launch;
 * a NULL prediction skips its term.  Targets are crop views: host_strides9 = (batch, channel, row) strides of t_d,
 * t_s, t_t.  sgn_d / sgn_s (optional, (B,3,h,w)) receive sign(pred - target) / n: the L1 gradients.  workspace:
 * wcmc_image_losses_workspace() bytes that the caller zero-fills ONCE and then keeps (per-CTA partials + a ticket
 * the launch re-arms); sums are accumulated in a fixed order.                                                        */
size_t wcmc_image_losses_workspace(void);
int wcmc_image_losses(const float* r_d, const float* r_s, const float* radiance, const float* t_d, const float* t_s,
                      const float* t_t, const long* host_strides9, int B, int h, int w, float eps, float* sgn_d,
                      float* sgn_s, float* sums4, void* workspace, size_t workspace_bytes, void* stream);

/* out2 = { target / max|g|, max|g| / target } over n fp32 values in one launch: the per-pass loss scale that keeps the
 * fp16 gradients of a backward pass in range (wcmc_b200/ops.py::grad_scale).  workspace: wcmc_absmax_scale_workspace()
 * bytes, zero-filled once by the caller and then kept.                                                            */
size_t wcmc_absmax_scale_workspace(void);
int wcmc_absmax_scale(const float* g, long n, float target, float* out2, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ---- device-side pairing permutations for the throughput mode of the loss (the reference draws
 * torch.randperm on the CPU, /root/reference/support/losses.py:35, :50; the parity mode keeps that) ----
 * out[i] = pi(i), a keyed Feistel bijection of [0, n) (cycle walking), no sort.  state: 2 x uint64 on the device,
 * state[0] = draw counter (seed it; the launch advances it by one), state[1] = 0.  salt distinguishes launches
 * that read the same counter value.                                                                        */
int wcmc_random_permutation(int64_t* out, long n, unsigned long long* state, unsigned salt, void* stream);

/* ---- data-parallel gradient exchange over NVSwitch peer memory (replaces nn.DataParallel's per-step gradient
 * reduction onto GPU 0, /root/reference/train_kpcn.py:266-269; one process per GPU here) -----------------------
 * In-place sum x scale over `world` ranks of the region [offset_floats, +n_floats) of a buffer that every rank has
 * allocated symmetrically: `local` = this rank's mapping, `peers_dev` = DEVICE array of the `world` mappings as seen
 * from this rank (peers_dev[rank] == local), `multicast` = multicast mapping of the same buffer or NULL (then the
 * kernel sums peer loads in rank order: deterministic).  Two-shot: rank r reduces slice r and broadcasts it; the
 * replicas end bit-identical.  The kernel synchronises the ranks itself through flags at byte `flag_offset_bytes`
 * of the buffer (wcmc_grad_exchange_flag_bytes() bytes, zero-filled ONCE by the caller before the first launch on
 * any rank, behind every region ever exchanged); `channel` 0 or 1 selects an independent set of flags so that two
 * exchanges may be in flight at once.  Every rank must issue the same sequence of calls per channel (same region
 * sizes); a rank that never arrives makes the others trap after ~4 s instead of hanging.  Safe to capture in a
 * CUDA graph.  Regions and offsets in floats, multiples of 4.                                                  */
size_t wcmc_grad_exchange_flag_bytes(void);
int wcmc_grad_exchange(float* local, float* multicast, float* const* peers_dev, long flag_offset_bytes,
                       long offset_floats, long n_floats, int rank, int world, int channel, float scale, void* stream);

/* ---- K11: all-pairs form of the path-disentangling loss on the tensor cores (EXTENSION: the reference
 * pairs each row with one random partner, /root/reference/support/losses.py:33-61; this is the quantity
 * that estimator samples, BASELINE.json north_star (4) / configs[4]; oracle/allpairs_ref.py) ---------
 * p_rows (N,D) fp32 row-major embeddings, ref_rows (N,3) fp32 reference radiance (tone-mapped inside).
 *   e_ij = 1/2|P_i-P_j|^2 - 1/2|t_i-t_j|^2 over all i != j [with 1/2|t_i-t_j|^2 < tau when tau > 0]
 *   mode 0: out[0] = sum 1/2 e^2 / N^2          mode 1: out[0] = (logsumexp(alpha [e,-e,0]) - log(1+2 kept)) / sqrt(alpha)
 *   out[1] = number of kept ordered pairs.  *nonfinite |= 1 on non-finite input (caller zero-initialises).
 * D <= 37.  workspace: wcmc_fmse_allpairs_workspace(N, D) bytes, 256-byte aligned.                 */
size_t wcmc_fmse_allpairs_workspace(int N, int D);
int wcmc_fmse_allpairs_fwd(const float* p_rows, const float* ref_rows, int N, int D, int mode, float alpha,
                           float tau, float* out, int* nonfinite, void* workspace, size_t workspace_bytes,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WCMC_H_ */
