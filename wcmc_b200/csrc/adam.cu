// K12: gradient clipping + Adam for every parameter tensor of the step in ONE launch.
//
// Replaces `nn.utils.clip_grad_value_(params, 1.0)` (/root/reference/support/interfaces.py:261) followed
// by `optim.step()` of three `torch.optim.Adam` instances (:269-271; constructed at
// /root/reference/train_kpcn.py:277 with default betas / eps, no weight decay, no amsgrad), which
// torch runs as ~50 multi-tensor launches making 6-7 passes over the 47 MB of parameters, gradients
// and moments.  Here every element is read once and written once: 8 streams x 4 B per parameter.
// The arithmetic follows torch's single-tensor Adam step operation by operation:
//     g   = clamp(g, -clip, clip)                       (written back: the reference leaves clipped grads)
//     m   = m + (g - m) * (1 - beta1)                   (lerp_)
//     v   = v * beta2 + (1 - beta2) * g * g             (mul_().addcmul_())
//     p   = p - (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kAdamThreads = 256;
constexpr int kAdamChunk = 256 * 16;  // elements per CTA

__global__ void __launch_bounds__(kAdamThreads)
adam_clip_kernel(const wcmc_adam_tensor* __restrict__ tensors, const int2* __restrict__ blocks,
                 const int* __restrict__ step, const int* __restrict__ ok_flag, float clip,
                 unsigned long long* __restrict__ nonfinite) {
    wcmc::pdl_start();
    if (ok_flag != nullptr && *ok_flag == 0) return;
    const int2 bk = blocks[blockIdx.x];
    const wcmc_adam_tensor T = tensors[bk.x];
    __shared__ float s_step_size, s_inv_bc2_sqrt;
    if (threadIdx.x == 0) {
        const double t = static_cast<double>(*step + 1);
        const double bc1 = 1.0 - pow(static_cast<double>(T.beta1), t);
        const double bc2 = 1.0 - pow(static_cast<double>(T.beta2), t);
        s_step_size = static_cast<float>(static_cast<double>(T.lr) / bc1);
        s_inv_bc2_sqrt = static_cast<float>(1.0 / sqrt(bc2));
    }
    __syncthreads();
    const float step_size = s_step_size, inv_bc2_sqrt = s_inv_bc2_sqrt;
    const float w1 = 1.f - T.beta1, w2 = 1.f - T.beta2, b2 = T.beta2, eps = T.eps;
    const long base = static_cast<long>(bk.y) * kAdamChunk;
    const long end = min(T.n, base + kAdamChunk);
    const bool vec = ((reinterpret_cast<uintptr_t>(T.p) | reinterpret_cast<uintptr_t>(T.g) |
                       reinterpret_cast<uintptr_t>(T.m) | reinterpret_cast<uintptr_t>(T.v)) & 15) == 0;
    // A non-finite gradient element (an fp16 overflow somewhere in a 16-bit backward pass, which fp32 training would not
    // have) must not reach the weights: torch's Adam would write NaN into p, m and v for good.  Such an element is
    // left alone -- no update, moments untouched -- and counted; the host reports the count (KPCNInterface).
    unsigned bad = 0;
    auto upd = [&](float& p, float& g, float& m, float& v) {
        if (!isfinite(g)) { ++bad; return; }
        if (clip > 0.f) g = fminf(fmaxf(g, -clip), clip);
        m = fmaf(g - m, w1, m);
        v = fmaf(w2 * g, g, v * b2);
        const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
        p = p - step_size * (m / denom);
    };
    if (vec) {
        for (long i = base + threadIdx.x * 4; i < end; i += kAdamThreads * 4) {
            if (i + 4 <= end) {
                float4 p = *reinterpret_cast<float4*>(T.p + i), g = *reinterpret_cast<float4*>(T.g + i);
                float4 m = *reinterpret_cast<float4*>(T.m + i), v = *reinterpret_cast<float4*>(T.v + i);
                upd(p.x, g.x, m.x, v.x);
                upd(p.y, g.y, m.y, v.y);
                upd(p.z, g.z, m.z, v.z);
                upd(p.w, g.w, m.w, v.w);
                *reinterpret_cast<float4*>(T.p + i) = p;
                *reinterpret_cast<float4*>(T.m + i) = m;
                *reinterpret_cast<float4*>(T.v + i) = v;
                if (clip > 0.f) *reinterpret_cast<float4*>(T.g + i) = g;
            } else {
                for (long j = i; j < end; ++j) {
                    float p = T.p[j], g = T.g[j], m = T.m[j], v = T.v[j];
                    upd(p, g, m, v);
                    T.p[j] = p; T.m[j] = m; T.v[j] = v;
                    if (clip > 0.f) T.g[j] = g;
                }
            }
        }
    } else {
        for (long j = base + threadIdx.x; j < end; j += kAdamThreads) {
            float p = T.p[j], g = T.g[j], m = T.m[j], v = T.v[j];
            upd(p, g, m, v);
            T.p[j] = p; T.m[j] = m; T.v[j] = v;
            if (clip > 0.f) T.g[j] = g;
        }
    }
    if (bad != 0 && nonfinite != nullptr) atomicAdd(nonfinite, static_cast<unsigned long long>(bad));
}

__global__ void adam_tick_kernel(int* step, const int* ok_flag) {
    wcmc::pdl_start();
    if (ok_flag == nullptr || *ok_flag != 0) *step += 1;
}

}  // namespace

extern "C" int wcmc_adam_chunk(void) { return kAdamChunk; }

extern "C" int wcmc_adam_clip_step(const wcmc_adam_tensor* dev_tensors, const int* dev_blocks, int nblocks,
                                   int* dev_step, const int* dev_ok_flag, float clip,
                                   unsigned long long* dev_nonfinite_count, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(dev_tensors != nullptr && dev_blocks != nullptr && dev_step != nullptr && nblocks >= 0,
                 WCMC_ESHAPE, "adam_clip_step: null pointer");
    if (nblocks > 0) {
        WCMC_LAUNCH(adam_clip_kernel, nblocks, kAdamThreads, 0, stream, dev_tensors, reinterpret_cast<const int2*>(dev_blocks),
                                                               dev_step, dev_ok_flag, clip, dev_nonfinite_count);
        WCMC_LAUNCH_CHECK();
    }
    WCMC_LAUNCH(adam_tick_kernel, 1, 1, 0, stream, dev_step, dev_ok_flag);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}
