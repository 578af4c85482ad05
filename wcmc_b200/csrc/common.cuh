// Shared device/host helpers for the sm_100a kernels of libwcmc.so.
// Inline-PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld)
// plus the host-side error plumbing of the C ABI (include/wcmc.h).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/wcmc.h"

// ------------------------------------------------------------------------------------------
// host side: error handling
// ------------------------------------------------------------------------------------------
void wcmc_set_error(const char* fmt, ...);

#define WCMC_CHECK_CUDA(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            wcmc_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,           \
                           cudaGetErrorString(_e));                                        \
            return WCMC_ECUDA;                                                             \
        }                                                                                  \
    } while (0)

#define WCMC_REQUIRE(cond, code, ...)                                                      \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            wcmc_set_error(__VA_ARGS__);                                                   \
            return (code);                                                                 \
        }                                                                                  \
    } while (0)

#define WCMC_LAUNCH_CHECK()                                                                \
    do {                                                                                   \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            wcmc_set_error("%s:%d kernel launch failed: %s", __FILE__, __LINE__,           \
                           cudaGetErrorString(_e));                                        \
            return WCMC_ECUDA;                                                             \
        }                                                                                  \
    } while (0)

// Encodes a tiled bf16 tensor map (rank <= 5).  dims/strides innermost first; strides in bytes
// for dims 1..rank-1 (stride of dim 0 is the element size).  Returns a WCMC_E* code.
int wcmc_encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, int swizzle128);

// Same for any element type: dtype = WCMC_F32 (4-byte elements) or a 16-bit code.
int wcmc_encode_tmap(CUtensorMap* map, int dtype, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle128);

int wcmc_num_sms();   // of the current device

// Opts `kernel` in to `bytes` of dynamic shared memory on the CURRENT device (once per kernel and device,
// mutex-guarded; the attribute is per device).  Returns a WCMC_E* code.
int wcmc_func_smem(const void* kernel, int bytes);
#define WCMC_FUNC_SMEM(kernel, bytes)                                                      \
    do {                                                                                   \
        int _rc = wcmc_func_smem(reinterpret_cast<const void*>(kernel), (bytes));          \
        if (_rc) return _rc;                                                               \
    } while (0)

// ------------------------------------------------------------------------------------------
// launches: programmatic dependent launch (PDL)
// ------------------------------------------------------------------------------------------
// Every kernel of this library starts with `pdl_launch_dependents(); pdl_wait();` (after its TMEM allocation, if it has
// one) and is launched through WCMC_LAUNCH with cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel
// of the stream may become resident as soon as every CTA of this one runs (or has left), set up its barriers /
// tensor memory / descriptors in the shadow of this kernel's tail, and then blocks in `griddepcontrol.wait` until
// this grid has COMPLETED and its memory is visible.  Rules that keep this safe:
//   * nothing touches global memory before pdl_wait() (kernel parameters, tensor maps and shared memory are fine);
//   * a kernel that allocates tensor memory triggers only AFTER the allocation (otherwise an early dependent could
//     take the columns a late CTA of this grid still needs, while waiting for this grid: deadlock).
// A kernel launched without the attribute, or after a kernel that never triggers (torch's), behaves as usual.
// OFF by default (wcmc_tuning_set("pdl", 1) turns the attribute on): in the two-stream step an early-resident
// dependent holds the shared memory of its SM while it waits, which keeps the OTHER stream's ready kernel off that
// SM -- measured 7.51 -> 7.72 ms per step (profiles/r02_pdl_ab.txt).  Without the attribute the two instructions
// are no-ops.
int wcmc_pdl_enabled();
#ifdef __CUDACC__
namespace wcmc {
template <typename... P, typename... A>
inline cudaError_t launch_pdl(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = wcmc_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}
}  // namespace wcmc
#define WCMC_LAUNCH(kernel, grid, block, smem, stream, ...)                                          \
    do {                                                                                           \
        cudaError_t _e = wcmc::launch_pdl(kernel, dim3(grid), dim3(block), (smem), (stream), __VA_ARGS__); \
        if (_e != cudaSuccess) {                                                                   \
            wcmc_set_error("%s:%d kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return WCMC_ECUDA;                                                                     \
        }                                                                                          \
    } while (0)
#endif

// ------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

namespace wcmc {

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_start() {   // kernels without tensor memory: first statement of the kernel
    pdl_launch_dependents();
    pdl_wait();
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// Adds `bytes` to the transaction count of the current phase WITHOUT arriving (a phase whose loads are issued
// in several pieces: the last piece arrives with mbar_expect_tx).
__device__ __forceinline__ void mbar_add_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Non-blocking poll (try_wait may suspend the thread for a hardware time slice before it answers "not yet":
// a thread that polls SEVERAL barriers must use this one, or a ready barrier waits behind an idle one).
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must never hang the GPU box.  After ~4 s of spinning the
// kernel traps (the launch then reports an error through the C ABI).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) {
            printf("wcmc: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n",
                   blockIdx.x, threadIdx.x, smem_u32(bar), parity);
            __trap();
        }
    }
}

// ---- TMA ---------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
        "{%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
        "{%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
        "{%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// TMA store (shared -> global) of one box; bulk-group completion.  The source tile must have been made
// visible to the async proxy (fence.proxy.async + barrier) and must stay intact until
// tma_store_wait_read() returns in the issuing thread.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- tcgen05 -----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrives (count 1) on `bar` once every tcgen05.mma issued so far by this thread has retired.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
            smem_u32(bar))
        : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
// Waits for all outstanding tcgen05.ld of this thread; the registers of the load being consumed are
// threaded through the asm so the compiler cannot read (or move) them before the wait.
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]),
                   "+r"(v[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- CTA pairs (thread-block cluster of 2, tcgen05 cta_group::2) ---------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// All threads of both CTAs (reconverges the warp first: barrier.cluster is .aligned).
__device__ __forceinline__ void cluster_sync_all() {
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// Arrive (count 1, release at cluster scope) on an mbarrier given by its shared::cluster address.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded like mbar_wait; acquires at cluster scope (the arrivals come from the peer CTA as well).
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) {
            printf("wcmc: cluster mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x,
                   threadIdx.x, smem_u32(bar), parity);
            __trap();
        }
    }
}
// TMA loads of a CTA pair: the box lands in THIS CTA's shared memory, the bytes are signalled on the mbarrier
// `bar_cluster` (a shared::cluster address: normally the leader CTA's barrier, which expects both halves).
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1,
                                                 int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
        "{%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1,
                                                 int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
        "{%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// TMEM of both CTAs of the pair; one warp of EACH CTA executes it.
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs, M = 256] (+)= A * B; each CTA supplies 128 rows of A and N/2 rows of B from the same
// shared-memory offsets.  Issued by ONE thread of the leader CTA (cluster rank 0).
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrives on the barrier at this shared-memory offset in every CTA of `cta_mask` once the MMAs issued so far
// have retired.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(cta_mask)
        : "memory");
}

// ---- descriptors -------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B (cf. PTX ISA "tcgen05 matrix descriptor"):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [49,52) base offset
//   | [61,64) layout type (2 = 128B swizzle)
__device__ __forceinline__ uint64_t make_sdesc_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                     uint32_t sbo_bytes, uint32_t base_offset = 0) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;
    d |= static_cast<uint64_t>(base_offset & 7) << 49;
    d |= 2ull << 61;
    return d;
}
// Instruction descriptor for kind::f16 (fp16 / bf16 inputs, independently per operand), fp32
// accumulate.
//   [4,6) D fmt (1=f32) | [7,10) A fmt (0=f16, 1=bf16) | [10,13) B fmt | 15 A major | 16 B major
//   | [17,23) N>>3 | [24,29) M>>4      (major: 0 = K-major, 1 = MN-major)
// a_dtype / b_dtype use the WCMC_BF16 / WCMC_F16 codes of include/wcmc.h.
__host__ __device__ __forceinline__ uint32_t make_idesc_f16(int M, int N, int a_mn_major, int b_mn_major,
                                                            int a_dtype, int b_dtype) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= static_cast<uint32_t>(a_dtype == WCMC_BF16 ? 1 : 0) << 7;
    d |= static_cast<uint32_t>(b_dtype == WCMC_BF16 ? 1 : 0) << 10;
    d |= static_cast<uint32_t>(a_mn_major & 1) << 15;
    d |= static_cast<uint32_t>(b_mn_major & 1) << 16;
    d |= static_cast<uint32_t>(N >> 3) << 17;
    d |= static_cast<uint32_t>(M >> 4) << 24;
    return d;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
// dtype: WCMC_BF16 or WCMC_F16 (16-bit storage types)
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi, int dtype) {
    return dtype == WCMC_F16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v, int dtype) {
    if (dtype == WCMC_F16) return __half22float2(*reinterpret_cast<__half2*>(&v));
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u));
}
// sign test that is valid for both 16-bit float formats: value > 0
__device__ __forceinline__ bool h16_pos(uint32_t h) { return (h & 0x8000u) == 0 && (h & 0x7FFFu) != 0; }
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace wcmc
#endif  // __CUDACC__
