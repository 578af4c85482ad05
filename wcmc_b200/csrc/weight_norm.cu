// Batched weight normalisation: w = g * v / ||v||  (norm over everything but the output channel), forward and
// backward for ALL convolutions of a network in one launch each.
//
// sbmc.modules.ConvChain builds its convolutions with `weight_norm=True` by default, and
// /root/reference/support/networks.py:18-24 keeps that default for the three sub-networks of PathNet: 20 layers per
// network.  torch runs one `_weight_norm` kernel per layer in the forward pass and one per layer in the backward
// pass (80 launches of 3-6 us per training step for the two path-embedding networks); here blockIdx.y = layer.
#include <algorithm>

#include "common.cuh"

namespace wcmc {

struct WnBatch {
    wcmc_wn_desc d[WCMC_WN_BATCH_MAX];
};

__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];   // same order in every thread: deterministic
    __syncthreads();
    return t;
}

// forward: w[row, :] = g[row] * v[row, :] / ||v[row, :]||, norm[row] saved for the backward pass
__global__ void __launch_bounds__(256) weight_norm_fwd_kernel(const WnBatch wb) {
    wcmc::pdl_start();
    __shared__ float red[8];
    const wcmc_wn_desc& L = wb.d[blockIdx.y];
    for (int row = blockIdx.x; row < L.rows; row += gridDim.x) {
        const float* v = L.v + static_cast<size_t>(row) * L.cols;
        float s = 0.f;
        for (int i = threadIdx.x; i < L.cols; i += 256) {
            const float x = __ldg(v + i);
            s += x * x;
        }
        const float nrm = sqrtf(block_sum_256(s, red));
        const float k = __ldg(L.g + row) / nrm;
        float* w = L.w + static_cast<size_t>(row) * L.cols;
        for (int i = threadIdx.x; i < L.cols; i += 256) w[i] = __ldg(v + i) * k;
        if (threadIdx.x == 0) L.norm[row] = nrm;
    }
}

// backward: dg[row] = <dw, v> / ||v||;  dv = (g / ||v||) * (dw - v * <dw, v> / ||v||^2)
__global__ void __launch_bounds__(256) weight_norm_bwd_kernel(const WnBatch wb) {
    wcmc::pdl_start();
    __shared__ float red[8];
    const wcmc_wn_desc& L = wb.d[blockIdx.y];
    for (int row = blockIdx.x; row < L.rows; row += gridDim.x) {
        const float* v = L.v + static_cast<size_t>(row) * L.cols;
        const float* dw = L.dw + static_cast<size_t>(row) * L.cols;
        float s = 0.f;
        for (int i = threadIdx.x; i < L.cols; i += 256) s += __ldg(dw + i) * __ldg(v + i);
        const float dot = block_sum_256(s, red);
        const float nrm = __ldg(L.norm + row), g = __ldg(L.g + row);
        const float a = g / nrm, b = dot / (nrm * nrm);
        float* dv = L.dv + static_cast<size_t>(row) * L.cols;
        for (int i = threadIdx.x; i < L.cols; i += 256) dv[i] = a * (__ldg(dw + i) - __ldg(v + i) * b);
        if (threadIdx.x == 0) L.dg[row] = dot / nrm;
    }
}

}  // namespace wcmc

using namespace wcmc;

extern "C" int wcmc_weight_norm_batch(const wcmc_wn_desc* host_descs, int n, int backward, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(host_descs != nullptr && n > 0, WCMC_ESHAPE, "weight_norm_batch: no layers");
    for (int base = 0; base < n; base += WCMC_WN_BATCH_MAX) {
        const int m = std::min(WCMC_WN_BATCH_MAX, n - base);
        WnBatch wb;
        int max_rows = 1;
        for (int i = 0; i < m; ++i) {
            const wcmc_wn_desc& L = host_descs[base + i];
            WCMC_REQUIRE(L.rows > 0 && L.cols > 0 && L.v && L.g && L.norm, WCMC_ESHAPE, "weight_norm_batch: bad layer %d",
                         base + i);
            WCMC_REQUIRE(backward ? (L.dw && L.dv && L.dg) : (L.w != nullptr), WCMC_ESHAPE,
                         "weight_norm_batch: layer %d lacks the %s pointers", base + i, backward ? "backward" : "forward");
            wb.d[i] = L;
            max_rows = std::max(max_rows, L.rows);
        }
        dim3 grid(static_cast<unsigned>(std::min(max_rows, 128)), m);
        if (backward) WCMC_LAUNCH(weight_norm_bwd_kernel, grid, 256, 0, stream, wb);
        else WCMC_LAUNCH(weight_norm_fwd_kernel, grid, 256, 0, stream, wb);
        WCMC_LAUNCH_CHECK();
    }
    return WCMC_OK;
}
