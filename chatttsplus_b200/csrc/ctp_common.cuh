// Shared device/host helpers for libctp (sm_100a only): error handling, mbarrier / TMA / tcgen05 PTX wrappers.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

// ---------------------------------------------------------------------------------------------------------
// Host-side error plumbing: every C-ABI entry returns ctp_status and records a thread-local message.
// ---------------------------------------------------------------------------------------------------------
void ctp_set_error(const char* fmt, ...);
// Kernel-launch accounting (bench.py's gpu_launches): every launch site calls ctp_count_launch(); while a step graph is
// being captured the launches are tallied per graph and added once per replay.
void ctp_count_launch(int n = 1);
void ctp_count_capture_begin();
long long ctp_count_capture_end();

#define CTP_CUDA_OK(expr)                                                                          \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ctp_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));   \
            return CTP_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

#define CTP_REQUIRE(cond, ...)                                                                     \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            ctp_set_error(__VA_ARGS__);                                                            \
            return CTP_ERR_INVALID;                                                                \
        }                                                                                          \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

namespace ctp {
// launch with (optionally) the programmatic-stream-serialization attribute; works inside stream capture
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
// ... and as thread-block clusters of `cluster_x` CTAs along x
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kc(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, unsigned cluster_x, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_x; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
}  // namespace ctp

namespace ctp {

// Programmatic dependent launch (PDL): let the next kernel in the stream start its prologue now; block before the first
// access to data the previous kernel produces.  Both are no-ops when the kernel was not launched with the PDL attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- in-graph timeline (bring-up: CTP_TRACE=1): per kernel launch a record of 4 globaltimer stamps (ns) -------------------
// [0] CTA 0 enters, [1] CTA 0 returns from griddepcontrol.wait, [2] CTA 0 leaves, [3] latest exit over all CTAs.
// Stamps are absolute, so after a run the records describe the LAST replay of the step graph.
__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool trace_cta0() { return blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0; }
__device__ __forceinline__ void trace_mark(unsigned long long* rec, int slot) {
    if (rec && trace_cta0()) rec[slot] = gtime_ns();
}
__device__ __forceinline__ void trace_end(unsigned long long* rec) {
    if (rec) {
        const unsigned long long t = gtime_ns();
        if (trace_cta0()) rec[2] = t;
        atomicMax(rec + 3, t);
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier -------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must never hang the GPU box (a hang is a strike).  ~2 s at 2 GHz, then trap.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("ctp: mbarrier wait timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
                   threadIdx.x);
            __trap();
        }
    }
}

// ---- TMA ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
// 2-D tiled load global -> shared (this CTA), completion on an mbarrier via complete_tx.
__device__ __forceinline__ void tma_load_2d(const void* desc, uint64_t* bar, void* smem_dst, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// 1-D bulk copy global -> shared (contiguous bytes, multiple of 16).
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :
                 : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- tcgen05 / TMEM -------------------------------------------------------------------------------------
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
                 "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], fp16/bf16 inputs, fp32 accumulate. Issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 columns of fp32: thread t of the warp receives lane (base_lane + t), columns [c, c+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major operand tile in shared memory, 128-byte swizzle (the layout TMA SWIZZLE_128B writes):
// rows of 128 B (64 fp16), 8-row atoms of 1024 B, 16-byte chunk index XORed with (row & 7).
// Descriptor fields (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=2 (SWIZZLE_128B) [61,64).
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr, uint32_t lbo16, uint32_t sbo16, uint32_t layout) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo16 & 0x3FFF) << 16;   // LBO>>4 (unused for swizzled K-major; canonical value 1)
    d |= static_cast<uint64_t>(sbo16 & 0x3FFF) << 32;   // SBO>>4: 8 rows * 128 B = 1024 B between row groups -> 64
    d |= static_cast<uint64_t>(1) << 46;                // descriptor version for sm_100
    d |= static_cast<uint64_t>(layout & 7) << 61;       // 2 = SWIZZLE_128B
    return d;
}
// Instruction descriptor for kind::f16, A/B fp16 K-major, D fp32, shape M x N (M in {64,128}, N%16==0).
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | (uint32_t(N >> 3) << 17) |
           (uint32_t(M >> 4) << 24);
}

// Exact-erf GELU (nn.GELU(), dvae.py:38 / vocos ConvNeXtBlock) = 0.5 x (1 + erf(x / sqrt 2)).  erf through Abramowitz-Stegun
// 7.1.26 (|error| <= 1.5e-7, far below the fp16 rounding of the value this feeds): ~12 instructions instead of erff's ~40 —
// the GELU epilogue of the 128x256 tiles (256 values per thread) was 3x longer than the tile's MMAs.
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float erf_abs = 1.0f - p * t * __expf(-z * z);
    return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}
// Same function with ONE special-function instruction instead of two (no reciprocal): erf(z) = 1 - 2^(z Q(z)) on [0, 4.2], Q a degree-6
// polynomial fitted to -log2(1 - erf z) / z (max |error| 1.7e-7 on erf, 5.7e-7 on the GELU in fp32; erf saturates to 1 - 3e-9 at 4.2).
// The 128x256 GELU tiles spend 2 MUFU per value = 4096 SFU cycles per tile, as long as the tile's MMAs.
__device__ __forceinline__ float gelu_erf_poly(float x) {
    const float z = fminf(fabsf(x) * 0.70710678118654752440f, 4.2f);
    float q = 1.0019626643e-4f;
    q = fmaf(q, z, -4.6142147039e-4f);
    q = fmaf(q, z, -2.3025872651e-3f);
    q = fmaf(q, z, 2.9452895746e-2f);
    q = fmaf(q, z, -1.4896386862e-1f);
    q = fmaf(q, z, -9.1832858324e-1f);
    q = fmaf(q, z, -1.6279137135f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * q));
    return 0.5f * x * (1.0f + copysignf(1.0f - e, x));
}
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

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

}  // namespace ctp
#endif  // __CUDACC__
