// TEST INFRASTRUCTURE (never linked into libwcmc.so): runs the per-element arithmetic of the N3 preprocessing kernels
// (wcmc_b200/csrc/preprocess_math.cuh -- the very functions the CUDA kernels call) in plain host loops, so that
// tests/test_host_logic.py can check it against the reference-generated vectors without a GPU.
#include <algorithm>

#include "../../wcmc_b200/csrc/preprocess_math.cuh"

using namespace wcmc::prep;

extern "C" void host_preprocess_llpm(const float* raw, long nrows, float* out) {
    for (long i = 0; i < nrows * 37; ++i) out[i] = llpm_value(raw + (i / 37) * kRawC, static_cast<int>(i % 37));
}

extern "C" void host_preprocess_kpcn(const float* raw, int H, int W, int S, float* ws, float* out44) {
    const long npix = static_cast<long>(H) * W;
    float md = 0.f;   // the kernel's atomicMax over int bits starts at 0.0f as well
    for (long p = 0; p < npix; ++p) md = std::max(md, kpcn_pixel_stats(raw + p * S * kRawC, S, ws + p * kStats));
    for (long i = 0; i < npix * 44; ++i) out44[i] = kpcn_finish_value(ws, W, S, md, i / 44, static_cast<int>(i % 44));
}
