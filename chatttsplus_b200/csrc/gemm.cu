// Host side of the tcgen05 GEMM: tensor-map construction (driver entry point resolved at run time, so the
// library links against nothing but the static CUDA runtime) and launch dispatch over the N-tile width.
#include "gemm.cuh"
#include "../../include/ctp.h"

#include <cudaTypedefs.h>
#include <algorithm>
#include <mutex>
#include <stdlib.h>

namespace ctp {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_once;
static int g_init_status = 0;
long long* g_dbg = nullptr;  // set by ctp_debug_gemm_stamps
static int g_sm_count = 148;
static int g_pair = 1;             // cta_group::2 pair per 256x256 tile when the problem fills a wave of pairs (CTP_GEMM_2CTA=0: never, 2: whenever eligible)
static bool g_persistent = true;   // CTP_GEMM_PERSISTENT=0 selects the one-tile-per-CTA kernel for the large GEMMs too
static uint32_t g_desc[4] = {1, 64, 2, 2};  // LBO>>4, SBO>>4, layout type, K-advance per UMMA_K (in 16-byte units)

template <int BN>
static cudaError_t set_smem_attr() {
    return cudaFuncSetAttribute(gemm_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<BN>::TOTAL);
}

int gemm_init() {
    std::call_once(g_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
            ctp_set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
            g_init_status = CTP_ERR_NO_DEVICE;
            return;
        }
        g_encode = reinterpret_cast<EncodeTiledFn>(fn);
        if (const char* pe = getenv("CTP_GEMM_PERSISTENT")) g_persistent = atoi(pe) != 0;
        if (const char* pe = getenv("CTP_GEMM_2CTA")) g_pair = atoi(pe);
        if (const char* d = getenv("CTP_DESC")) {  // bring-up diagnostics only
            unsigned a, b, c, e;
            if (sscanf(d, "%u,%u,%u,%u", &a, &b, &c, &e) == 4) { g_desc[0] = a; g_desc[1] = b; g_desc[2] = c; g_desc[3] = e; }
        }
        cudaError_t pa = cudaFuncSetAttribute(gemm_tcgen05_persistent<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmPSmem<256>::TOTAL);
        if (pa == cudaSuccess) pa = cudaFuncSetAttribute(gemm_tcgen05_persistent<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmPSmem<128>::TOTAL);
        if (pa == cudaSuccess) pa = cudaFuncSetAttribute(gemm_tcgen05_persistent2<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmP2Smem::TOTAL);
        if (pa == cudaSuccess) { cudaDeviceProp prop; int dev = 0; cudaGetDevice(&dev); if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) g_sm_count = prop.multiProcessorCount; }
        if (pa != cudaSuccess) { ctp_set_error("cudaFuncSetAttribute(persistent gemm smem): %s", cudaGetErrorString(pa)); g_init_status = CTP_ERR_CUDA; return; }
        cudaError_t a = set_smem_attr<32>();
        if (a == cudaSuccess) a = set_smem_attr<64>();
        if (a == cudaSuccess) a = set_smem_attr<128>();
        if (a == cudaSuccess) a = set_smem_attr<256>();
        if (a == cudaSuccess) a = cudaFuncSetAttribute(gemm_tcgen05_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<32, 1>::TOTAL);
        if (a == cudaSuccess) a = cudaFuncSetAttribute(gemm_tcgen05_kernel<32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<32, 2>::TOTAL);
        if (a != cudaSuccess) {
            ctp_set_error("cudaFuncSetAttribute(gemm smem): %s", cudaGetErrorString(a));
            g_init_status = CTP_ERR_CUDA;
        }
    });
    return g_init_status;
}

int make_tmap_kmajor(CUtensorMap* out, const void* base, long long rows, long long K, long long ld_elems, int box_rows) {
    int st = gemm_init();
    if (st) return st;
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld_elems * 2) & 15) != 0) {
        ctp_set_error("tensor map: base %p / row pitch %lld elements must be 16-byte aligned", base, ld_elems);
        return CTP_ERR_INVALID;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)(ld_elems * 2)};
    cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ctp_set_error("cuTensorMapEncodeTiled failed (%d): rows=%lld K=%lld ld=%lld box_rows=%d", (int)r, rows, K, ld_elems,
                      box_rows);
        return CTP_ERR_CUDA;
    }
    return CTP_OK;
}

// K-major fp16 matrix [rows, K]: box = 32 k x box_rows, 64-byte swizzle (the 32-wide K slices of W_down in the fused MLP kernel)
int make_tmap_k32_sw64(CUtensorMap* out, const void* base, long long rows, long long K, long long ld_elems, int box_rows) {
    int st = gemm_init();
    if (st) return st;
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld_elems * 2) & 15) != 0) {
        ctp_set_error("tensor map: base %p / row pitch %lld elements must be 16-byte aligned", base, ld_elems);
        return CTP_ERR_INVALID;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)(ld_elems * 2)};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ctp_set_error("cuTensorMapEncodeTiled (64B swizzle) failed (%d): rows=%lld K=%lld ld=%lld box_rows=%d", (int)r, rows, K, ld_elems, box_rows);
        return CTP_ERR_CUDA;
    }
    return CTP_OK;
}

int make_tmap_f32(CUtensorMap* out, const void* base, long long rows, long long K, long long ld_elems, int box_rows) {
    int st = gemm_init();
    if (st) return st;
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld_elems * 4) & 15) != 0) {
        ctp_set_error("tensor map: base %p / row pitch %lld elements must be 16-byte aligned", base, ld_elems);
        return CTP_ERR_INVALID;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)(ld_elems * 4)};
    cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ctp_set_error("cuTensorMapEncodeTiled (fp32) failed (%d): rows=%lld K=%lld ld=%lld box_rows=%d", (int)r, rows, K, ld_elems, box_rows);
        return CTP_ERR_CUDA;
    }
    return CTP_OK;
}

int gemm_launch_x(int xmode, const CUtensorMap& tmA, const CUtensorMap& tmX, long long a_rows, long long T, long long K, int split_k,
                  const GemmEpilogue& epi, const GemmShape& extra, cudaStream_t stream, bool pdl) {
    int st = gemm_init();
    if (st) return st;
    if (T > 32 || (K % GEMM_BK) != 0 || !epi.swap || !epi.atomic || (xmode == 1 && !extra.norm_w) || (xmode != 1 && xmode != 2)) {
        ctp_set_error("gemm_launch_x: needs <= 32 token rows, K %% 64 == 0, the swap/atomic decode epilogue and mode 1 (norm_w) or 2");
        return CTP_ERR_INVALID;
    }
    GemmShape shp = extra;
    shp.a_independent = pdl ? 1 : 0;
    shp.dbg = nullptr;
    shp.k_blocks = (int)(K / GEMM_BK);
    shp.desc_lbo = g_desc[0]; shp.desc_sbo = g_desc[1]; shp.desc_layout = g_desc[2]; shp.desc_kadv = g_desc[3];
    if (split_k < 1) split_k = 1;
    if (split_k > shp.k_blocks) split_k = shp.k_blocks;
    dim3 grid(1, (unsigned)((a_rows + GEMM_BM - 1) / GEMM_BM), (unsigned)split_k);
    cudaError_t e;
    if (xmode == 1) e = launch_k(gemm_tcgen05_kernel<32, 1>, grid, dim3(GEMM_THREADS), (size_t)GemmSmem<32, 1>::TOTAL, stream, pdl, tmA, tmX, shp, epi);
    else e = launch_k(gemm_tcgen05_kernel<32, 2>, grid, dim3(GEMM_THREADS), (size_t)GemmSmem<32, 2>::TOTAL, stream, pdl, tmA, tmX, shp, epi);
    ctp_count_launch();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { ctp_set_error("decode gemm (in-kernel operand) launch failed: %s", cudaGetErrorString(e)); return CTP_ERR_CUDA; }
    return CTP_OK;
}

template <int BN>
static int launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, long long a_rows, long long b_rows, long long K,
                     int split_k, const GemmEpilogue& epi, cudaStream_t stream, const void* pf_ptr, unsigned long long pf_bytes, bool pdl,
                     float* zero_ptr, unsigned long long zero_f4, unsigned long long* trace) {
    GemmShape shp{};
    shp.pf_ptr = pf_ptr; shp.pf_bytes = pf_bytes; shp.a_independent = pdl ? 1 : 0;
    shp.zero_ptr = zero_ptr; shp.zero_f4 = zero_f4; shp.trace = trace;
    shp.k_blocks = (int)((K + GEMM_BK - 1) / GEMM_BK);
    shp.dbg = g_dbg;
    shp.desc_lbo = g_desc[0]; shp.desc_sbo = g_desc[1]; shp.desc_layout = g_desc[2]; shp.desc_kadv = g_desc[3];
    if (split_k < 1) split_k = 1;
    if (split_k > shp.k_blocks) split_k = shp.k_blocks;
    if (split_k > 1 && !(epi.atomic && !epi.out_f16)) {
        ctp_set_error("gemm: split_k > 1 needs fp32 atomic output");
        return CTP_ERR_INVALID;
    }
    dim3 grid((unsigned)((b_rows + BN - 1) / BN), (unsigned)((a_rows + GEMM_BM - 1) / GEMM_BM), (unsigned)split_k);
    cudaError_t e = launch_k(gemm_tcgen05_kernel<BN>, grid, dim3(GEMM_THREADS), (size_t)GemmSmem<BN>::TOTAL, stream, pdl, tmA, tmB, shp, epi);
    ctp_count_launch();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        ctp_set_error("gemm launch failed: %s", cudaGetErrorString(e));
        return CTP_ERR_CUDA;
    }
    return CTP_OK;
}

int gemm_launch_maps(const CUtensorMap& tmA, const CUtensorMap& tmB, long long a_rows, long long b_rows, long long K,
                     int block_n, int split_k, const GemmEpilogue& epi, cudaStream_t stream, const void* pf_ptr, unsigned long long pf_bytes, bool pdl,
                     float* zero_ptr, unsigned long long zero_f4, unsigned long long* trace) {
    int st = gemm_init();
    if (st) return st;
    if (g_persistent && !epi.swap && !epi.atomic && split_k <= 1 && a_rows >= 512 && (block_n == 256 || block_n == 128)) {
        GemmShape shp{};
        shp.k_blocks = (int)((K + GEMM_BK - 1) / GEMM_BK);
        shp.desc_lbo = g_desc[0]; shp.desc_sbo = g_desc[1]; shp.desc_layout = g_desc[2]; shp.desc_kadv = g_desc[3];
        const int tiles_m = (int)((a_rows + GEMM_BM - 1) / GEMM_BM), tiles_n = (int)((b_rows + block_n - 1) / block_n);
        const int grid = std::min(tiles_m * tiles_n, g_sm_count);
        cudaError_t e;
        if (block_n == 256) e = launch_k(gemm_tcgen05_persistent<256>, dim3(grid), dim3(GEMM_P_THREADS), (size_t)GemmPSmem<256>::TOTAL, stream, false, tmA, tmB, shp, epi, tiles_m, tiles_n);
        else e = launch_k(gemm_tcgen05_persistent<128>, dim3(grid), dim3(GEMM_P_THREADS), (size_t)GemmPSmem<128>::TOTAL, stream, false, tmA, tmB, shp, epi, tiles_m, tiles_n);
        ctp_count_launch();
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { ctp_set_error("persistent gemm launch failed: %s", cudaGetErrorString(e)); return CTP_ERR_CUDA; }
        return CTP_OK;
    }
    switch (block_n) {
        case 32: return launch_bn<32>(tmA, tmB, a_rows, b_rows, K, split_k, epi, stream, pf_ptr, pf_bytes, pdl, zero_ptr, zero_f4, trace);
        case 64: return launch_bn<64>(tmA, tmB, a_rows, b_rows, K, split_k, epi, stream, pf_ptr, pf_bytes, pdl, zero_ptr, zero_f4, trace);
        case 128: return launch_bn<128>(tmA, tmB, a_rows, b_rows, K, split_k, epi, stream, pf_ptr, pf_bytes, pdl, zero_ptr, zero_f4, trace);
        case 256: return launch_bn<256>(tmA, tmB, a_rows, b_rows, K, split_k, epi, stream, pf_ptr, pf_bytes, pdl, zero_ptr, zero_f4, trace);
        default: ctp_set_error("gemm: unsupported block_n %d", block_n); return CTP_ERR_INVALID;
    }
}

// The pair tile pays once the problem holds at least one full wave of 256 x 256 tiles (measured: +4-8 % at M = 131072, 1316 -> 1376 TFLOP/s at
// 16384 x 4096 x 4096; -7 % at M = 4096, N = 768 where 48 pairs leave a third of the SMs idle)
static bool pair_wins(long long a_rows, long long n_cols) {
    if (g_pair == 0) return false;
    const long long pairs = ((a_rows + 2 * GEMM_BM - 1) / (2 * GEMM_BM)) * ((n_cols + 255) / 256);
    return g_pair == 2 || pairs >= g_sm_count / 2;
}

// one cta_group::2 pair per 256 x 256 tile; tmB must have been built with 128-row boxes
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, long long a_rows, long long n_cols, GemmShape shp, const GemmEpilogue& epi, cudaStream_t stream) {
    const int tiles_m2 = (int)((a_rows + 2 * GEMM_BM - 1) / (2 * GEMM_BM)), tiles_n = (int)((n_cols + 255) / 256);
    const int pairs = std::min(tiles_m2 * tiles_n, g_sm_count / 2);
    cudaError_t e = launch_kc(gemm_tcgen05_persistent2<256>, dim3(2 * pairs), dim3(GEMM_P_THREADS), (size_t)GemmP2Smem::TOTAL, stream, false, 2u, tmA, tmB, shp, epi, tiles_m2,
                              tiles_n);
    ctp_count_launch();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { ctp_set_error("cta-pair gemm launch failed: %s", cudaGetErrorString(e)); return CTP_ERR_CUDA; }
    return CTP_OK;
}

int gemm_launch_swiglu(const void* A, long long a_rows, long long lda, const void* Wgu, long long I, long long K, __half* out, long long ldo,
                       cudaStream_t stream) {
    int st = gemm_init();
    if (st) return st;
    if ((I % 128) != 0 || (ldo % 8) != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0) {
        ctp_set_error("gemm_launch_swiglu: I %% 128, output pitch %% 8 and a 16-byte aligned output are required");
        return CTP_ERR_INVALID;
    }
    CUtensorMap tmA, tmB;
    if ((st = make_tmap_kmajor(&tmA, A, a_rows, K, lda, GEMM_BM))) return st;
    if ((st = make_tmap_kmajor(&tmB, Wgu, 2 * I, K, K, 128))) return st;   // two 128-row boxes per 256-column tile
    GemmShape shp{};
    shp.k_blocks = (int)((K + GEMM_BK - 1) / GEMM_BK);
    shp.desc_lbo = g_desc[0]; shp.desc_sbo = g_desc[1]; shp.desc_layout = g_desc[2]; shp.desc_kadv = g_desc[3];
    shp.swiglu_up_row = (int)I;
    GemmEpilogue epi{};
    epi.out = out; epi.ldo = ldo; epi.out_f16 = 1; epi.T = (int)a_rows; epi.F = (int)I;
    if (pair_wins(a_rows, 2 * I)) return launch_pair(tmA, tmB, a_rows, 2 * I, shp, epi, stream);
    const int tiles_m = (int)((a_rows + GEMM_BM - 1) / GEMM_BM), tiles_n = (int)(I / 128);
    const int grid = std::min(tiles_m * tiles_n, g_sm_count);
    cudaError_t e = launch_k(gemm_tcgen05_persistent<256>, dim3(grid), dim3(GEMM_P_THREADS), (size_t)GemmPSmem<256>::TOTAL, stream, false, tmA, tmB, shp, epi,
                             tiles_m, tiles_n);
    ctp_count_launch();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { ctp_set_error("swiglu gemm launch failed: %s", cudaGetErrorString(e)); return CTP_ERR_CUDA; }
    return CTP_OK;
}

int gemm_launch(const GemmLaunch& g, cudaStream_t stream) {
    CUtensorMap tmA, tmB;
    int st = make_tmap_kmajor(&tmA, g.A, g.a_rows, g.K, g.lda, GEMM_BM);
    if (st) return st;
    const bool pair = g_persistent && !g.epi.swap && !g.epi.atomic && g.split_k <= 1 && g.a_rows >= 512 && g.block_n == 256 && pair_wins(g.a_rows, g.b_rows);
    st = make_tmap_kmajor(&tmB, g.B, g.b_rows, g.K, g.ldb, pair ? 128 : g.block_n);
    if (st) return st;
    if (pair) {
        GemmShape shp{};
        shp.k_blocks = (int)((g.K + GEMM_BK - 1) / GEMM_BK);
        shp.desc_lbo = g_desc[0]; shp.desc_sbo = g_desc[1]; shp.desc_layout = g_desc[2]; shp.desc_kadv = g_desc[3];
        return launch_pair(tmA, tmB, g.a_rows, g.b_rows, shp, g.epi, stream);
    }
    return gemm_launch_maps(tmA, tmB, g.a_rows, g.b_rows, g.K, g.block_n, g.split_k, g.epi, stream);
}

}  // namespace ctp
