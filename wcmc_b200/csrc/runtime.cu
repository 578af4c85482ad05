// Host-side runtime of libwcmc.so: error state, device check, TMA tensor-map encoding.
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void wcmc_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* wcmc_last_error(void) { return g_err; }
extern "C" const char* wcmc_version(void) { return "wcmc-b200 0.1 (sm_100a)"; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
constexpr int kMaxDevices = 64;
static int g_sms[kMaxDevices] = {0};                               // per device, written by wcmc_init(device)
static std::map<std::pair<const void*, int>, int> g_func_smem;     // (kernel, device) -> opted-in dynamic smem bytes
static std::mutex g_mu;

extern "C" int wcmc_init(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    cudaDeviceProp prop;
    WCMC_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    WCMC_REQUIRE(prop.major == 10, WCMC_EARCH,
                 "wcmc_init: device %d is sm_%d%d; libwcmc.so is sm_100a only (no fallback)", device,
                 prop.major, prop.minor);
    WCMC_REQUIRE(device >= 0 && device < kMaxDevices, WCMC_ESHAPE, "wcmc_init: device index %d out of range", device);
    g_sms[device] = prop.multiProcessorCount;
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        WCMC_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        WCMC_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, WCMC_ECUDA,
                     "wcmc_init: cuTensorMapEncodeTiled not available from the driver");
        g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    return WCMC_OK;
}

// SM count of the CURRENT device (the launches of this library go to the current device).
int wcmc_num_sms() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
    return g_sms[dev] > 0 ? g_sms[dev] : 148;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: every (kernel, device) pair opts in once,
// behind the library mutex (nn.DataParallel drives one replica per Python thread and per device through this
// library; /root/reference/train_kpcn.py:266-269).
int wcmc_func_smem(const void* kernel, int bytes) {
    int dev = 0;
    WCMC_CHECK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    const auto key = std::make_pair(kernel, dev);
    auto it = g_func_smem.find(key);
    if (it != g_func_smem.end() && it->second >= bytes) return WCMC_OK;
    WCMC_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    g_func_smem[key] = bytes;
    return WCMC_OK;
}

int wcmc_encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, int swizzle128) {
    return wcmc_encode_tmap(map, WCMC_BF16, base, rank, dims, strides_bytes, box, swizzle128);
}

int wcmc_encode_tmap(CUtensorMap* map, int dtype, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle128) {
    WCMC_REQUIRE(g_encode != nullptr, WCMC_ECUDA, "wcmc_init() has not been called");
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = g_encode(map, dtype == WCMC_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                          static_cast<cuuint32_t>(rank),
                          const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        wcmc_set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box "
                       "[%u,%u,%u,%u] stride0 %llu base %p",
                       static_cast<int>(r), rank, (unsigned long long)dims[0],
                       (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
                       (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], rank > 1 ? box[1] : 0,
                       rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0,
                       (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), base);
        return WCMC_ECUDA;
    }
    return WCMC_OK;
}

static int g_pdl = 0;   // measured: slower inside the two-stream step (profiles/r02_pdl_ab.txt); wcmc_tuning_set("pdl", 1) enables it
int wcmc_pdl_enabled() { return g_pdl; }

// Tuning hooks for the micro-benchmarks under tools/ (never used by the product path).
int wcmc_ka_set_tile(int w);         // kernel_apply.cu
int wcmc_wgrad_set_uniform(int v);   // conv_wgrad.cu
int wcmc_conv_set_row_stages(int v); // conv_igemm.cu
int wcmc_conv_set_pair(int v);       // conv_igemm.cu
int wcmc_conv_set_model(int which, int v);  // conv_igemm.cu
int wcmc_wgrad_group_set(int which, int v); // conv_wgrad_group.cu
int wcmc_conv_set_resident(int v);          // conv_igemm.cu
int wcmc_exchange_set_blocks(int v);        // grad_exchange.cu
int wcmc_conv_set_share(int v);             // conv_igemm.cu
int wcmc_conv_set_share_ks(int v);          // conv_igemm.cu
extern "C" int wcmc_tuning_set(const char* name, int value) {
    if (name != nullptr && strcmp(name, "ka_tile_w") == 0 && wcmc_ka_set_tile(value) == 0) return WCMC_OK;
    if (name != nullptr && strcmp(name, "wgrad_uniform") == 0) return wcmc_wgrad_set_uniform(value);
    if (name != nullptr && strcmp(name, "conv_row_stages") == 0) return wcmc_conv_set_row_stages(value);
    if (name != nullptr && strcmp(name, "conv_pair") == 0) return wcmc_conv_set_pair(value);
    if (name != nullptr && strcmp(name, "conv_pair_min_clk") == 0 && wcmc_conv_set_model(0, value) == 0) return WCMC_OK;
    if (name != nullptr && strcmp(name, "conv_item_clk") == 0 && wcmc_conv_set_model(1, value) == 0) return WCMC_OK;
    if (name != nullptr && strcmp(name, "conv_resident") == 0) return wcmc_conv_set_resident(value);
    if (name != nullptr && strcmp(name, "conv_share") == 0 && wcmc_conv_set_share(value) == 0) return WCMC_OK;
    if (name != nullptr && strcmp(name, "conv_share_ks") == 0) return wcmc_conv_set_share_ks(value);
    if (name != nullptr && strcmp(name, "pdl") == 0) { g_pdl = value ? 1 : 0; return WCMC_OK; }
    if (name != nullptr && strcmp(name, "exchange_blocks") == 0 && wcmc_exchange_set_blocks(value) == 0) return WCMC_OK;
    if (name != nullptr && strcmp(name, "wgrad_group_pack") == 0) return wcmc_wgrad_group_set(0, value);
    if (name != nullptr && strcmp(name, "wgrad_group") == 0) return wcmc_wgrad_group_set(1, value);
    if (name != nullptr && strcmp(name, "conv_plane_slots") == 0 && wcmc_conv_set_model(2, value) == 0) return WCMC_OK;
    wcmc_set_error("wcmc_tuning_set: unknown knob or bad value (%s = %d)", name ? name : "(null)", value);
    return WCMC_ESHAPE;
}
