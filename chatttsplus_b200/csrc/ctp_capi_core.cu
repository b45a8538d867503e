// C-ABI core: error string, version, device check, and the GEMM building-block entry point.
#include "../../include/ctp.h"
#include "gemm.cuh"

#include <stdarg.h>

namespace ctp { extern long long* g_dbg; }
static thread_local char g_err[1024] = "";

void ctp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static long long g_launches = 0;
static thread_local bool g_capturing = false;
static thread_local long long g_capture_launches = 0;
void ctp_count_launch(int n) { if (g_capturing) g_capture_launches += n; else g_launches += n; }
void ctp_count_capture_begin() { g_capturing = true; g_capture_launches = 0; }
long long ctp_count_capture_end() { g_capturing = false; return g_capture_launches; }

extern "C" {

long long ctp_launch_count(int reset) { long long v = g_launches; if (reset) g_launches = 0; return v; }

const char* ctp_last_error(void) { return g_err; }
int ctp_version(void) { return 100; }

ctp_status ctp_device_check(int dev) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= dev) {
        ctp_set_error("no CUDA device %d (%s)", dev, cudaGetErrorString(e));
        return CTP_ERR_NO_DEVICE;
    }
    cudaDeviceProp p;
    CTP_CUDA_OK(cudaGetDeviceProperties(&p, dev));
    if (p.major != 10) {
        ctp_set_error("device %d is sm_%d%d; libctp is built for sm_100a only", dev, p.major, p.minor);
        return CTP_ERR_UNSUPPORTED;
    }
    return CTP_OK;
}

// bring-up hook (not part of include/ctp.h): device buffer of per-CTA clock64 stamps for subsequent GEMM launches
__attribute__((visibility("default"))) void ctp_debug_gemm_stamps(long long* buf) { ctp::g_dbg = buf; }

ctp_status ctp_gemm_f16(int32_t M, int32_t N, int32_t K, const void* A, int64_t lda, const void* B, int64_t ldb, void* out,
                        int64_t ldo, const float* bias, int32_t flags, int32_t block_n, int32_t split_k, ctp_stream stream) {
    CTP_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: bad shape %d %d %d", M, N, K);
    ctp::GemmLaunch g{};
    g.A = A; g.a_rows = M; g.lda = lda;
    g.B = B; g.b_rows = N; g.ldb = ldb;
    g.K = K; g.block_n = block_n; g.split_k = split_k;
    ctp::GemmEpilogue& e = g.epi;
    e.out = out; e.ldo = ldo;
    e.out_f16 = flags & 1; e.act_gelu = (flags >> 1) & 1; e.atomic = (flags >> 2) & 1; e.swap = (flags >> 3) & 1;
    e.bias = bias; e.gamma = nullptr; e.residual = nullptr; e.ldr = 0; e.row_valid = nullptr;
    if (e.swap) { e.F = M; e.T = N; } else { e.T = M; e.F = N; }
    return (ctp_status)ctp::gemm_launch(g, (cudaStream_t)stream);
}

}  // extern "C"
