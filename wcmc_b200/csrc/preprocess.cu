// raw OptaGen sample buffer (H,W,S,104) fp32 -> the KPCN feature buffer (H,W,44) and the path
// descriptor buffer (H,W,S,37), i.e. DenoiseDataset._preprocess_kpcn / _preprocess_llpm of
// /root/reference/support/datasets.py:487-582 / :301-361 (numpy on one CPU core there), with the NaN / inf clamp of
// :621-624 folded into the loads.  HBM-bound: 416 B per sample read, 176 B per pixel + 148 B per sample written.
//
//   llpm   : purely elementwise -- one thread per output element, consecutive threads = consecutive output channels of
//            one sample (reads inside one 176-byte run of the raw row, fully coalesced writes).
//   kpcn   : pass 1, one thread per pixel: mean / variance over the S samples of the 13 raw channels it needs (two
//            28-byte runs per sample: every 32-byte sector is fetched once and reused through L1 by the next loads of
//            the same sample), the albedo factorisation / log transform, 18 per-pixel values to a workspace and an
//            atomic max of the mean depth;  pass 2, one thread per (pixel, output channel): depth normalisation by
//            the image maximum, left / top finite differences, the 44-channel layout.
// Status: written in round 1 after the GPU budget was spent.  The per-element arithmetic lives in preprocess_math.cuh
// (__host__ __device__) and is checked on the CPU against the reference-generated vectors
// (tests/test_host_logic.py::test_preprocess_kernel_arithmetic_on_host); the kernels themselves have NOT yet run on a
// GPU, nothing on the product path calls them and their GPU tests run as non-strict xfail (tests/test_gpu_preprocess.py).
#include <algorithm>

#include "common.cuh"
#include "preprocess_math.cuh"

namespace wcmc {

__global__ void __launch_bounds__(256) preprocess_llpm_kernel(const float* __restrict__ raw, long nrows,
                                                             float* __restrict__ out) {
    const long total = nrows * 37;
    for (long i = blockIdx.x * 256L + threadIdx.x; i < total; i += gridDim.x * 256L) {
        const long row = i / 37;
        out[i] = prep::llpm_value(raw + row * prep::kRawC, static_cast<int>(i - row * 37));
    }
}

__global__ void __launch_bounds__(128) preprocess_kpcn_stats_kernel(const float* __restrict__ raw, long npix, int S,
                                                                   float* __restrict__ ws, int* __restrict__ max_depth_bits) {
    float local_max = 0.f;
    for (long p = blockIdx.x * 128L + threadIdx.x; p < npix; p += gridDim.x * 128L)
        local_max = fmaxf(local_max, prep::kpcn_pixel_stats(raw + p * S * prep::kRawC, S, ws + p * prep::kStats));
    // depth.max(): non-negative floats order like their bit patterns; a negative maximum never beats the initial 0.0,
    // which reproduces `if max_depth > 0` (datasets.py:519-521)
    local_max = warp_max(local_max);
    if ((threadIdx.x & 31) == 0 && local_max > 0.f) atomicMax(max_depth_bits, __float_as_int(local_max));
}

__global__ void __launch_bounds__(256) preprocess_kpcn_finish_kernel(const float* __restrict__ ws, int H, int W, int S,
                                                                    const int* __restrict__ max_depth_bits,
                                                                    float* __restrict__ out) {
    const float md = __int_as_float(__ldg(max_depth_bits));
    const long total = static_cast<long>(H) * W * 44;
    for (long i = blockIdx.x * 256L + threadIdx.x; i < total; i += gridDim.x * 256L) {
        const long p = i / 44;
        out[i] = prep::kpcn_finish_value(ws, W, S, md, p, static_cast<int>(i - p * 44));
    }
}

}  // namespace wcmc

using namespace wcmc;

extern "C" size_t wcmc_preprocess_kpcn_workspace(int H, int W) {
    return (static_cast<size_t>(H) * W * prep::kStats + 4) * sizeof(float);
}

extern "C" int wcmc_preprocess_kpcn(const float* raw, int H, int W, int S, float* out44, void* workspace,
                                    size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(raw && out44 && workspace && H > 0 && W > 0 && S >= 1 && S <= 8, WCMC_ESHAPE,
                 "preprocess_kpcn: need H, W > 0 and 1 <= spp <= 8 (got %d x %d, %d spp)", H, W, S);
    WCMC_REQUIRE(workspace_bytes >= wcmc_preprocess_kpcn_workspace(H, W), WCMC_EWORKSPACE,
                 "preprocess_kpcn: workspace too small");
    const long npix = static_cast<long>(H) * W;
    float* ws = static_cast<float*>(workspace);
    int* md = reinterpret_cast<int*>(ws + npix * prep::kStats);
    WCMC_CHECK_CUDA(cudaMemsetAsync(md, 0, sizeof(int), stream));
    const int sms = wcmc_num_sms();
    const int b1 = static_cast<int>(std::min<long>((npix + 127) / 128, 8L * sms));
    preprocess_kpcn_stats_kernel<<<b1, 128, 0, stream>>>(raw, npix, S, ws, md);
    WCMC_LAUNCH_CHECK();
    const int b2 = static_cast<int>(std::min<long>((npix * 44 + 255) / 256, 16L * sms));
    preprocess_kpcn_finish_kernel<<<b2, 256, 0, stream>>>(ws, H, W, S, md, out44);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}

extern "C" int wcmc_preprocess_llpm(const float* raw, long nrows, float* out37, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WCMC_REQUIRE(raw && out37 && nrows > 0, WCMC_ESHAPE, "preprocess_llpm: bad arguments");
    const int sms = wcmc_num_sms();
    const int blocks = static_cast<int>(std::min<long>((nrows * 37 + 255) / 256, 16L * sms));
    preprocess_llpm_kernel<<<blocks, 256, 0, stream>>>(raw, nrows, out37);
    WCMC_LAUNCH_CHECK();
    return WCMC_OK;
}
