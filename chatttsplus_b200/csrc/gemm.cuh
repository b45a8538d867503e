// tcgen05 + TMA GEMM for sm_100a:  D[Mrows, Ncols] = sum_k A[Mrows,k] * B[Ncols,k]   (both operands K-major fp16)
//
//   * one 128 x BN output tile per CTA, K in blocks of 64 (one 128-byte swizzle row), STAGES-deep TMA ring
//   * warp 4 lane 0: TMA producer; warp 5 lane 0: tcgen05.mma issuer (accumulator in TMEM, fp32);
//     warps 0-3: epilogue (tcgen05.ld 32x32b -> registers -> fused epilogue -> global)
//   * split-K over gridDim.z with fp32 atomic accumulation (decode path: the weight stream is spread over SMs)
//   * "swap" orientation: the M operand is the weight matrix [features, K] and the N operand the (few) tokens, so
//     a batch-32 decode step still fills the 128-row MMA (rows = output features, columns = batch rows).
//
// Operand tiles in shared memory use the 128B-swizzled K-major canonical layout (see ctp_common.cuh).
#pragma once
#include "ctp_common.cuh"

namespace ctp {

struct GemmEpilogue {
    void* out;               // fp32 or fp16, element (t, f) at out[t*ldo + f]
    long long ldo;
    int out_f16;             // 1: __half output, 0: float
    int atomic;              // 1: atomicAdd into fp32 out (split-K or residual accumulate)
    int swap;                // 0: tile rows = tokens t, tile cols = features f;  1: rows = f, cols = t
    int act_gelu;            // exact-erf GELU after bias: 1 = Abramowitz-Stegun form (rcp + ex2), 2 = one-MUFU form (ctp_common.cuh)
    const float* bias;       // [F] or null
    const float* gamma;      // [F] or null (applied after activation)
    const float* residual;   // fp32 [T][ldr] or null (added last)
    long long ldr;
    const unsigned char* row_valid;  // [T] or null: invalid tokens are written as 0
    int T, F;                // logical extents (bounds for partial tiles)
};

struct GemmShape {
    int k_blocks;  // ceil(K / 64)
    // shared-memory matrix descriptor fields (defaults = 128B-swizzled K-major canonical layout); overridable through
    // the CTP_DESC environment variable for bring-up diagnostics only
    uint32_t desc_lbo, desc_sbo, desc_layout, desc_kadv;
    long long* dbg;  // bring-up only: per-CTA clock64 stamps [cta][8], or null
    unsigned long long* trace;   // bring-up only: in-graph timeline record of this launch (ctp_common.cuh), or null
    int a_independent;        // PDL: the M operand (weights in the decode path) does not depend on the previous kernel
    const void* pf_ptr;       // optional: region the NEXT kernel will stream (its weights); every CTA prefetches a share into L2
    unsigned long long pf_bytes;
    // XNORM variant (decode): the token operand is built in the kernel from the fp32 residual stream, B[t][k] = fp16(x[t][k] * w[k])
    // (RMSNorm folded into the GEMM, llama.py:82-87).  tmB is then an fp32 tensor map over x; the contraction is linear in the row
    // factor rsqrt(mean(x[t]^2) + eps), which the CONSUMER of the accumulator applies (it recomputes it from x[t], 3 KB per row).
    const float* norm_w;      // [K]
    float* ss_out;            // XNORM, optional: [T] += sum over this CTA's k-slice of x[t][k]^2 (m-tile 0 only, fp32 atomics)
    // XSILU variant (decode down_proj): B[t][k] = fp16(silu(r*g[t][k]) * (r*u[t][k])) (llama.py:214) built from the fp32 gate|up
    // accumulator (tmB: fp32 map over [T][2I], gate tile at column k, up tile at column up_off + k); r[t] = rsqrt(ss_in[t]/ss_dim + eps)
    // is the RMSNorm row factor deferred from the gate|up GEMM (null: r = 1)
    const float* ss_in;
    float ss_dim, eps;
    int up_off;
    // any mode: a 16-byte aligned scratch region this kernel re-arms to zero after griddepcontrol.wait (dealt over the CTAs);
    // its last readers ran in an earlier kernel of the step
    float* zero_ptr;
    unsigned long long zero_f4;
    // persistent kernel, SwiGLU mode (prefill gate|up GEMM, llama.py:214): the 256-column weight tile is TWO boxes of 128 rows — gate rows
    // [n0/2, n0/2+128) and up rows [swiglu_up_row + n0/2, ...) of the packed [gate | up] matrix — so that an accumulator row holds gate(f) in
    // column j and up(f) in column 128 + j, and the epilogue writes fp16 silu(gate) * up [T][I] directly (no fp32 [T][2I] round trip)
    int swiglu_up_row;        // 0: off
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_P_THREADS = 320;   // persistent kernel: 8 epilogue warps (0-3, 6-9) + TMA warp 4 + MMA warp 5

// XMODE: 0 = token operand by TMA (fp16 matrix), 1 = XNORM, 2 = XSILU (see GemmShape)
template <int BN, int XMODE = 0>
struct GemmSmem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int X_BYTES = XMODE * BN * GEMM_BK * 4;   // fp32 landing tile(s): x (XNORM) or gate and up (XSILU); TMA, no swizzle
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES + X_BYTES;
    // <= 200 KB of operand ring; leaves room for the barrier block and 1 KB alignment slack under 227 KB.
    // BN=32 is the decode path (<= 4 k-blocks per CTA after split-K): 4 stages = 83 KB so that two CTAs — e.g. of two
    // independent half-batch chains — fit on one SM.
    // XNORM: 3 stages of 28 KB, XSILU: 2 stages of 36 KB (<= 2 k-blocks per CTA after split-K): two CTAs per SM as well.
    static constexpr int STAGES = XMODE == 1 ? 3 : XMODE == 2 ? 2 : (BN == 32 ? 4 : ((200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES));
    static constexpr int BAR_BYTES = 256;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
    static_assert(XMODE == 0 || BN == 32, "the in-kernel token operands are built for the 32-token decode tile");
};

template <int BN>
__device__ __forceinline__ void gemm_epilogue_store(const GemmEpilogue& e, int row_g, int col_g0, const float (&acc)[32],
                                                    bool add_bias) {
    // row_g: global index along the M operand; col_g0..+31: along the N operand
    if (!e.swap) {
        const int t = row_g;
        if (t >= e.T) return;
        const bool valid = e.row_valid ? (e.row_valid[t] != 0) : true;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int f = col_g0 + j;
            float x = acc[j];
            if (f < e.F) {
                if (e.bias && add_bias) x += e.bias[f];
                if (e.act_gelu) x = e.act_gelu == 2 ? gelu_erf_poly(x) : gelu_erf(x);
                if (e.gamma) x *= e.gamma[f];
                if (e.residual && add_bias) x += e.residual[(long long)t * e.ldr + f];
            }
            v[j] = valid ? x : 0.0f;
        }
        if (e.out_f16) {
            __half* o = reinterpret_cast<__half*>(e.out) + (long long)t * e.ldo + col_g0;
            if (col_g0 + 32 <= e.F && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    __half2 h0 = __floats2half2_rn(v[j], v[j + 1]);
                    __half2 h1 = __floats2half2_rn(v[j + 2], v[j + 3]);
                    __half2 h2 = __floats2half2_rn(v[j + 4], v[j + 5]);
                    __half2 h3 = __floats2half2_rn(v[j + 6], v[j + 7]);
                    uint4 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&h0);
                    pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    pk.z = *reinterpret_cast<uint32_t*>(&h2);
                    pk.w = *reinterpret_cast<uint32_t*>(&h3);
                    *reinterpret_cast<uint4*>(o + j) = pk;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col_g0 + j < e.F) o[j] = __float2half_rn(v[j]);
            }
        } else {
            float* o = reinterpret_cast<float*>(e.out) + (long long)t * e.ldo + col_g0;
            if (e.atomic) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col_g0 + j < e.F) atomicAdd(o + j, v[j]);
            } else if (col_g0 + 32 <= e.F && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col_g0 + j < e.F) o[j] = v[j];
            }
        }
    } else {
        // rows are features: consecutive lanes hold consecutive f for a fixed token -> coalesced along f
        const int f = row_g;
        if (f >= e.F) return;
        const float b = (e.bias && add_bias) ? e.bias[f] : 0.0f;
        const float g = e.gamma ? e.gamma[f] : 1.0f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int t = col_g0 + j;
            if (t >= e.T) break;
            float x = acc[j] + b;
            if (e.act_gelu) x = e.act_gelu == 2 ? gelu_erf_poly(x) : gelu_erf(x);
            x *= g;
            if (e.residual && add_bias) x += e.residual[(long long)t * e.ldr + f];
            if (e.row_valid && !e.row_valid[t]) x = 0.0f;
            if (e.out_f16) {
                reinterpret_cast<__half*>(e.out)[(long long)t * e.ldo + f] = __float2half_rn(x);
            } else {
                float* o = reinterpret_cast<float*>(e.out) + (long long)t * e.ldo + f;
                if (e.atomic) atomicAdd(o, x);
                else *o = x;
            }
        }
    }
}

template <int BN, int XMODE = 0>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmShape shp, const GemmEpilogue epi) {
    using S = GemmSmem<BN, XMODE>;
    constexpr bool XNORM = XMODE == 1, XSILU = XMODE == 2;
    constexpr int STAGES = S::STAGES;
    constexpr uint32_t FULL_COUNT = XMODE ? 5 : 1;                       // weight TMA (arrive.expect_tx) + 4 converter warps
    constexpr uint32_t TX_BYTES = XMODE ? S::A_BYTES : S::A_BYTES + S::B_BYTES;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B-swizzle atoms
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint64_t* x_bar = accum_bar + 1;                                     // [STAGES] fp32 tile landed (XNORM only)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_bar + STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN;
    const int m0 = blockIdx.y * GEMM_BM;
    const int nsplit = gridDim.z;
    const int kb0 = (int)(((long long)shp.k_blocks * blockIdx.z) / nsplit);
    const int kb1 = (int)(((long long)shp.k_blocks * (blockIdx.z + 1)) / nsplit);
    const int nkb = kb1 - kb0;
    long long* dbg = shp.dbg ? shp.dbg + ((long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 : nullptr;
    if (dbg && threadIdx.x == 0) dbg[0] = clock64();
    if (threadIdx.x == 0) trace_mark(shp.trace, 0);
    pdl_launch_dependents();

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], FULL_COUNT);
            mbar_init(&empty_bar[s], 1);
            if (XMODE) mbar_init(&x_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 5) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (dbg && threadIdx.x == 0) dbg[1] = clock64();
    const bool early_a = shp.a_independent != 0;
    if (!(early_a && warp == 4 && lane == 0)) pdl_wait();   // the producer thread waits after issuing the weight loads
    if (threadIdx.x == 0) trace_mark(shp.trace, 1);

    if (warp == 4) {
        if (lane == 0) {
            int pre = 0;
            if (early_a) {
                // weights first (they do not depend on the previous kernel), then wait, then the activation halves
                pre = nkb < STAGES ? nkb : STAGES;
                for (int i = 0; i < pre; ++i) {
                    mbar_expect_tx(&full_bar[i], TX_BYTES);
                    tma_load_2d(&tmA, &full_bar[i], ring + i * S::STAGE_BYTES, (kb0 + i) * GEMM_BK, m0);
                }
                pdl_wait();
                for (int i = 0; i < pre; ++i) {
                    if (XMODE) {
                        uint8_t* xt = ring + i * S::STAGE_BYTES + S::A_BYTES + S::B_BYTES;
                        mbar_expect_tx(&x_bar[i], S::X_BYTES);
                        tma_load_2d(&tmB, &x_bar[i], xt, (kb0 + i) * GEMM_BK, n0);
                        if (XSILU) tma_load_2d(&tmB, &x_bar[i], xt + S::X_BYTES / 2, shp.up_off + (kb0 + i) * GEMM_BK, n0);
                    } else {
                        tma_load_2d(&tmB, &full_bar[i], ring + i * S::STAGE_BYTES + S::A_BYTES, (kb0 + i) * GEMM_BK, n0);
                    }
                }
            }
            for (int i = pre; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* a = ring + s * S::STAGE_BYTES;
                uint8_t* b = a + S::A_BYTES;
                mbar_expect_tx(&full_bar[s], TX_BYTES);
                tma_load_2d(&tmA, &full_bar[s], a, (kb0 + i) * GEMM_BK, m0);
                if (XMODE) {
                    mbar_expect_tx(&x_bar[s], S::X_BYTES);
                    tma_load_2d(&tmB, &x_bar[s], b + S::B_BYTES, (kb0 + i) * GEMM_BK, n0);
                    if (XSILU) tma_load_2d(&tmB, &x_bar[s], b + S::B_BYTES + S::X_BYTES / 2, shp.up_off + (kb0 + i) * GEMM_BK, n0);
                } else {
                    tma_load_2d(&tmB, &full_bar[s], b, (kb0 + i) * GEMM_BK, n0);
                }
            }
            if (shp.pf_ptr) {   // after our own loads are in flight: warm L2 with the next GEMM's weights
                const unsigned long long n_cta = (unsigned long long)gridDim.x * gridDim.y * gridDim.z;
                const unsigned long long cta = ((unsigned long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
                unsigned long long per = ((shp.pf_bytes + n_cta - 1) / n_cta + 127) & ~127ULL;
                unsigned long long off = cta * per;
                if (off < shp.pf_bytes) {
                    unsigned long long n = shp.pf_bytes - off < per ? shp.pf_bytes - off : per;
                    n &= ~15ULL;
                    const char* src = reinterpret_cast<const char*>(shp.pf_ptr) + off;
                    while (n > 0) {
                        const unsigned int chunk = n > 32768ULL ? 32768u : (unsigned int)n;
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(chunk) : "memory");
                        src += chunk;
                        n -= chunk;
                    }
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(GEMM_BM, BN);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                if (dbg && i == 0) dbg[2] = clock64();
                tc_fence_after();
                const uint32_t a_addr = smem_u32(ring + s * S::STAGE_BYTES);
                const uint64_t da = make_kmajor_desc(a_addr, shp.desc_lbo, shp.desc_sbo, shp.desc_layout);
                const uint64_t db = make_kmajor_desc(a_addr + S::A_BYTES, shp.desc_lbo, shp.desc_sbo, shp.desc_layout);
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k) {
                    // advance 16 fp16 = 32 bytes along K inside the swizzle row: +2 in the (addr>>4) field
                    umma_f16(tmem_base, da + (uint64_t)(shp.desc_kadv * k), db + (uint64_t)(shp.desc_kadv * k), idesc,
                             (i > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);  // frees this smem stage once the MMAs above retire
            }
            umma_commit(accum_bar);  // accumulator complete
            if (dbg) dbg[3] = clock64();
        }
    } else {
        // epilogue warps 0..3: TMEM lanes [32*warp, 32*warp+32)
        if (shp.zero_ptr) {
            const unsigned long long n_cta = (unsigned long long)gridDim.x * gridDim.y * gridDim.z;
            const unsigned long long cta = ((unsigned long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
            const unsigned long long per = (shp.zero_f4 + n_cta - 1) / n_cta;
            const unsigned long long lo = cta * per, hi = (lo + per < shp.zero_f4) ? lo + per : shp.zero_f4;
            float4* z = reinterpret_cast<float4*>(shp.zero_ptr);
            for (unsigned long long q = lo + threadIdx.x; q < hi; q += 128) z[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (XMODE) {
            // converter: fp32 tile(s) landed by TMA ([32 rows][64 k] row-major) -> fp16 operand in the 128B-swizzled K-major layout
            // the MMA reads (16-byte chunk c of row r at r*128 + ((c ^ (r & 7)) << 4)).  Thread (c16 = et & 15, r0 = et >> 4) owns k
            // columns 4*c16..+3 of rows r0 + 8j: shared-memory reads are conflict-free, each thread fills half a chunk.
            const int et = threadIdx.x, c16 = et & 15, r0 = et >> 4;
            float rf[BN / 8], ssacc[BN / 8];
#pragma unroll
            for (int rr = 0; rr < BN / 8; ++rr) {
                ssacc[rr] = 0.f;
                rf[rr] = 1.f;
                const int t = n0 + rr * 8 + r0;
                if (XSILU && shp.ss_in && t < epi.T) rf[rr] = rsqrtf(__ldcg(shp.ss_in + t) / shp.ss_dim + shp.eps);
            }
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                float4 w4 = make_float4(1.f, 1.f, 1.f, 1.f);
                if (XNORM) w4 = __ldg(reinterpret_cast<const float4*>(shp.norm_w + (kb0 + i) * GEMM_BK + c16 * 4));
                if (i >= STAGES) mbar_wait(&empty_bar[s], ((i / STAGES) & 1) ^ 1);   // the MMAs that read this B tile have retired
                mbar_wait(&x_bar[s], (i / STAGES) & 1);
                if (i == 0 && threadIdx.x == 0) trace_mark(shp.trace, 4);
                uint8_t* bt = ring + s * S::STAGE_BYTES + S::A_BYTES;
                const float* xt = reinterpret_cast<const float*>(bt + S::B_BYTES);
#pragma unroll
                for (int rr = 0; rr < BN / 8; ++rr) {
                    const int r = rr * 8 + r0;
                    const float4 a4 = *reinterpret_cast<const float4*>(xt + r * GEMM_BK + c16 * 4);
                    float v0, v1, v2, v3;
                    if (XNORM) {
                        ssacc[rr] += a4.x * a4.x + a4.y * a4.y + a4.z * a4.z + a4.w * a4.w;
                        // un-normalised operand: saturate instead of overflowing fp16 should a checkpoint carry a massive activation
                        v0 = fminf(fmaxf(a4.x * w4.x, -65504.f), 65504.f); v1 = fminf(fmaxf(a4.y * w4.y, -65504.f), 65504.f);
                        v2 = fminf(fmaxf(a4.z * w4.z, -65504.f), 65504.f); v3 = fminf(fmaxf(a4.w * w4.w, -65504.f), 65504.f);
                    } else {
                        const float4 u4 = *reinterpret_cast<const float4*>(xt + BN * GEMM_BK + r * GEMM_BK + c16 * 4);
                        const float q = rf[rr];
                        v0 = silu(a4.x * q) * (u4.x * q); v1 = silu(a4.y * q) * (u4.y * q);
                        v2 = silu(a4.z * q) * (u4.z * q); v3 = silu(a4.w * q) * (u4.w * q);
                    }
                    __half2 h0 = __floats2half2_rn(v0, v1), h1 = __floats2half2_rn(v2, v3);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    *reinterpret_cast<uint2*>(bt + r * 128 + (((c16 >> 1) ^ (r & 7)) << 4) + ((c16 & 1) << 3)) = pk;
                }
                fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[s]);
            }
            if (threadIdx.x == 0) trace_mark(shp.trace, 5);
            if (XNORM && shp.ss_out && blockIdx.y == 0) {   // every split adds its k-slice of sum(x^2) once
#pragma unroll
                for (int rr = 0; rr < BN / 8; ++rr) {
                    float v = ssacc[rr];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    const int t = n0 + rr * 8 + r0;
                    if (c16 == 0 && t < epi.T) atomicAdd(shp.ss_out + t, v);
                }
            }
        }
        if (nkb > 0) {
            mbar_wait(accum_bar, 0);
            // The epilogue reuses the operand ring as transpose scratch.  Its writes are ordered after the converter warps' operand writes
            // through full_bar -> tcgen05.mma -> tcgen05.commit -> accum_bar; compute-sanitizer's racecheck does not model that chain
            // (profiles/r2_sanitizer_racecheck.log reported the pair as a write-write hazard), so the same four warps also meet at a named
            // barrier the tool does understand (a few cycles per launch)
            if (XMODE) asm volatile("bar.sync 1, 128;" ::: "memory");
            if (threadIdx.x == 0) trace_mark(shp.trace, 6);
            if (dbg && threadIdx.x == 0) dbg[4] = clock64();
            tc_fence_after();
            const int row_g = m0 + warp * 32 + lane;
            const bool add_bias = (blockIdx.z == 0);
            // decode fast path: swap orientation + fp32 atomic accumulate and nothing else.  The accumulator tile is transposed
            // through shared memory (the operand ring is idle once the accumulator is complete) so that each lane owns 4
            // consecutive features of one token and issues vector reductions: 4x fewer RED operations through the LSU
            // (scalar: ~6.9k cycles per 128x32 tile, measured with tests/prof_gemm_stamps.py).
            const bool fast = epi.swap && epi.atomic && !epi.out_f16 && !epi.bias && !epi.gamma && !epi.residual && !epi.row_valid &&
                              !epi.act_gelu && ((epi.ldo & 3) == 0) && ((epi.F & 3) == 0);   // (partial last m-tile: per-lane feature guard below)
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                float acc[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, acc);
                if (fast) {
                    float* tile = reinterpret_cast<float*>(ring) + warp * (32 * 36);   // [32 tokens][32 features + 4 pad]
#pragma unroll
                    for (int j = 0; j < 32; ++j) tile[j * 36 + lane] = acc[j];
                    __syncwarp();
                    const int f4 = (lane & 7) * 4;
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const int tl = 4 * r + (lane >> 3);
                        const int t = n0 + c + tl;
                        if (t < epi.T && m0 + warp * 32 + f4 < epi.F) {
                            const float4 v = *reinterpret_cast<const float4*>(tile + tl * 36 + f4);
                            float* o = reinterpret_cast<float*>(epi.out) + (long long)t * epi.ldo + (m0 + warp * 32 + f4);
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                        }
                    }
                    __syncwarp();
                } else {
                    gemm_epilogue_store<BN>(epi, row_g, n0 + c, acc, add_bias);
                }
            }
        }
    }
    if (dbg && threadIdx.x == 0) dbg[5] = clock64();
    if (threadIdx.x == 0) trace_mark(shp.trace, 7);
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc<TMEM_COLS>(tmem_base);
    if (dbg && threadIdx.x == 160) dbg[6] = clock64();
    if (threadIdx.x == 0) trace_end(shp.trace);
}

// ---------------------------------------------------------------------------------------------------------
// Persistent variant for the large, compute-bound GEMMs (prefill M = B*L0, vocoder M = mel frames): one CTA per SM loops
// over output tiles (grouped order, gemm_p_tile below); TWO accumulator tiles in TMEM
// so the epilogue of tile i (TMEM -> registers -> bias / GELU / layer-scale / residual -> vectorised global stores) overlaps
// the tcgen05.mma stream of tile i+1.  Normal orientation only (rows = tokens), no split-K.
// ---------------------------------------------------------------------------------------------------------
template <int BN>
struct GemmPSmem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;
    static constexpr int VEC_BYTES = 2 * BN * 4;            // per-tile bias / gamma staging
    static constexpr int TOTAL = STAGES * STAGE_BYTES + VEC_BYTES + 256 + 1024;
};

// Epilogue of one accumulator tile for one of the eight epilogue warps of the persistent kernels: TMEM lane quarter q (tile rows), column
// half `half`; bias / GELU / layer-scale / residual / row mask / fp16-or-fp32 store, or the SwiGLU form.  The accumulator is handed back
// to the MMA thread (arrival on the shared::cluster address tempty_addr: the CTA's own barrier, or the pair leader's) right after this
// warp's last TMEM read.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

template <int BN>
__device__ __forceinline__ void gemm_p_epilogue_tile(const GemmShape& shp, const GemmEpilogue& epi, uint32_t tmem_base, int acc, int m0, int n0, int q,
                                                     int half, int lane, const float* s_bias, const float* s_gamma, uint32_t tempty_addr) {
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < epi.T;
    const bool valid = row_ok && (epi.row_valid ? (epi.row_valid[row] != 0) : true);
    if (shp.swiglu_up_row) {
        // SwiGLU epilogue: this warp owns features [n0/2 + half * BN/4, + BN/4) of its 32 token rows: gate from columns
        // [half * BN/4, ...), up from the same columns of the second half of the tile
#pragma unroll 1
        for (int c = half * (BN / 4); c < (half + 1) * (BN / 4); c += 32) {
            float g[32], u[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c), g);
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + BN / 2 + c), u);
            if (c + 32 >= (half + 1) * (BN / 4)) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_addr);
            }
            const int f0 = n0 / 2 + c;
            if (row_ok && f0 < epi.F) {
                __half* o = reinterpret_cast<__half*>(epi.out) + (long long)row * epi.ldo + f0;
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 pk;
                    __half2 h0 = __floats2half2_rn(silu(g[i]) * u[i], silu(g[i + 1]) * u[i + 1]);
                    __half2 h1 = __floats2half2_rn(silu(g[i + 2]) * u[i + 2], silu(g[i + 3]) * u[i + 3]);
                    __half2 h2 = __floats2half2_rn(silu(g[i + 4]) * u[i + 4], silu(g[i + 5]) * u[i + 5]);
                    __half2 h3 = __floats2half2_rn(silu(g[i + 6]) * u[i + 6], silu(g[i + 7]) * u[i + 7]);
                    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                    *reinterpret_cast<uint4*>(o + i) = pk;
                }
            }
        }
        return;
    }
    const int c_lo = half * (BN / 2), c_hi = c_lo + BN / 2;
#pragma unroll 1
    for (int c = c_lo; c < c_hi; c += 32) {
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c), v);
        if (c + 32 >= c_hi) {   // last read of this accumulator by this warp: hand it back to the MMA warp as early as possible
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_addr);
        }
        const int f0 = n0 + c;
        if (row_ok && f0 < epi.F) {   // (a block, not `continue`: the warp reconverges before the next aligned tcgen05.ld)
        if (epi.bias) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += s_bias[c + i];
        }
        if (epi.act_gelu == 2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_erf_poly(v[i]);
        } else if (epi.act_gelu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
        }
        if (epi.gamma) {   // (the K = 512 GELU tiles are epilogue-bound: 2 warps per scheduler x 128 values x ~17 instructions against 4096 MMA cycles)
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= s_gamma[c + i];
        }
        const bool full = (f0 + 32 <= epi.F);
        if (epi.residual) {
            const float* rp = epi.residual + (long long)row * epi.ldr + f0;
            if (full && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 r4 = *reinterpret_cast<const float4*>(rp + i);
                    v[i] += r4.x; v[i + 1] += r4.y; v[i + 2] += r4.z; v[i + 3] += r4.w;
                }
            } else {
                for (int i = 0; i < 32; ++i) if (f0 + i < epi.F) v[i] += rp[i];
            }
        }
        if (!valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        if (epi.out_f16) {
            __half* o = reinterpret_cast<__half*>(epi.out) + (long long)row * epi.ldo + f0;
            if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    __half2 h0 = __floats2half2_rn(v[i], v[i + 1]), h1 = __floats2half2_rn(v[i + 2], v[i + 3]);
                    __half2 h2 = __floats2half2_rn(v[i + 4], v[i + 5]), h3 = __floats2half2_rn(v[i + 6], v[i + 7]);
                    uint4 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                    *reinterpret_cast<uint4*>(o + i) = pk;
                }
            } else {
                for (int i = 0; i < 32; ++i) if (f0 + i < epi.F) o[i] = __float2half_rn(v[i]);
            }
        } else {
            float* o = reinterpret_cast<float*>(epi.out) + (long long)row * epi.ldo + f0;
            if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
                for (int i = 0; i < 32; ++i) if (f0 + i < epi.F) o[i] = v[i];
            }
        }
        }
    }
}

// Tile order of the persistent kernel: groups of G = gridDim.x / tiles_n m-tiles, and inside a group all n-tiles of those m-tiles.
// One wave of CTAs then covers every n-tile of G m-tiles: an activation tile is pulled from HBM once and shared through L2 by the
// tiles_n CTAs that need it at the same time (n-major order re-read the whole activation matrix from HBM once per n-tile: 800 MB
// instead of 134 MB for the [131072, 512] x [1536, 512] pointwise conv, ncu dram__bytes_read), the weight matrix stays L2-resident.
__device__ __forceinline__ void gemm_p_tile(int t, int tiles_m, int tiles_n, int G, int& tm, int& tn) {
    const int per = G * tiles_n;
    const int g = t / per, r = t - g * per;
    const int m_base = g * G;
    const int ge = tiles_m - m_base < G ? tiles_m - m_base : G;   // last group may be short
    tn = r / ge;
    tm = m_base + (r - tn * ge);
}

template <int BN>
__global__ void __launch_bounds__(GEMM_P_THREADS, 1)
gemm_tcgen05_persistent(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmShape shp,
                        const GemmEpilogue epi, const int tiles_m, const int tiles_n) {
    using S = GemmPSmem<BN>;
    constexpr int STAGES = S::STAGES;
    constexpr uint32_t TMEM_COLS = 2 * BN;   // 512 (BN=256) or 256 (BN=128)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    float* s_bias = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES);
    float* s_gamma = s_bias + BN;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES + S::VEC_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull = empty_bar + STAGES;    // [2] accumulator ready
    uint64_t* tempty = tfull + 2;            // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = tiles_m * tiles_n;
    const int nkb = shp.k_blocks;
    const int G = (int)gridDim.x / tiles_n > 0 ? (int)gridDim.x / tiles_n : 1;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 5) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                int tm, tn;
                gemm_p_tile(t, tiles_m, tiles_n, G, tm, tn);
                const int m0 = tm * GEMM_BM, n0 = tn * BN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                    uint8_t* a = ring + s * S::STAGE_BYTES;
                    mbar_expect_tx(&full_bar[s], S::STAGE_BYTES);
                    tma_load_2d(&tmA, &full_bar[s], a, kb * GEMM_BK, m0);
                    if (shp.swiglu_up_row) {
                        tma_load_2d(&tmB, &full_bar[s], a + S::A_BYTES, kb * GEMM_BK, n0 / 2);
                        tma_load_2d(&tmB, &full_bar[s], a + S::A_BYTES + S::B_BYTES / 2, kb * GEMM_BK, shp.swiglu_up_row + n0 / 2);
                    } else {
                        tma_load_2d(&tmB, &full_bar[s], a + S::A_BYTES, kb * GEMM_BK, n0);
                    }
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(GEMM_BM, BN);
            uint32_t it = 0;
            int j = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++j) {
                const int acc = j & 1;
                mbar_wait(&tempty[acc], ((j >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(ring + s * S::STAGE_BYTES);
                    const uint64_t da = make_kmajor_desc(a_addr, shp.desc_lbo, shp.desc_sbo, shp.desc_layout);
                    const uint64_t db = make_kmajor_desc(a_addr + S::A_BYTES, shp.desc_lbo, shp.desc_sbo, shp.desc_layout);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k)
                        umma_f16(d_tmem, da + (uint64_t)(shp.desc_kadv * k), db + (uint64_t)(shp.desc_kadv * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tfull[acc]);
            }
        }
    } else {
        // epilogue: EIGHT warps (0-3 and 6-9).  A warp reads the TMEM lane quarter (warp % 4); the two warps of a quarter split the
        // tile's columns in halves, so the per-tile epilogue (up to 256 GELUs per accumulator row, which outlasted the MMAs of a
        // K = 512 tile with four warps: tensor pipe 27 %, profiles/r2_voc_kernels_full_summary.csv) takes half as long.
        const int q = warp & 3;                       // TMEM lanes [32 q, 32 q + 32) = tile rows
        const int half = warp < 4 ? 0 : 1;            // columns [half * BN/2, (half + 1) * BN/2)
        const int et = (warp < 4 ? warp : warp - 2) * 32 + lane;   // 0..255
        int j = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++j) {
            const int acc = j & 1;
            int tm, tn;
            gemm_p_tile(t, tiles_m, tiles_n, G, tm, tn);
            const int m0 = tm * GEMM_BM, n0 = tn * BN;
            // stage the per-feature vectors of this tile (previous tile's readers are past their last use: see bar below)
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (int i = et; i < BN; i += 256) {
                const int f = n0 + i;
                s_bias[i] = (epi.bias && f < epi.F) ? epi.bias[f] : 0.f;
                s_gamma[i] = (epi.gamma && f < epi.F) ? epi.gamma[f] : 1.f;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(&tfull[acc], (j >> 1) & 1);
            tc_fence_after();
            gemm_p_epilogue_tile<BN>(shp, epi, tmem_base, acc, m0, n0, q, half, lane, s_bias, s_gamma, smem_u32(&tempty[acc]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster — two SMs of one TPC — work on ONE 256 x 256 output tile.  Each CTA loads its
// own 128 rows of the activation tile and HALF of the weight tile (128 of the 256 weight rows) per k-block: 32 KB per SM per k-block
// instead of 48 KB, which is what bounds the K = 512 pointwise convolutions of the vocoder (94 B/clk per SM demanded by the one-CTA
// tile against ~64 B/clk an SM can take in).  Protocol (the CUTLASS 2-SM scheme, written out in PTX):
//   * both CTAs run a TMA producer; every load is cp.async.bulk.tensor...cta_group::2 and completes on the LEADER's (rank 0) full
//     barrier, for which the leader's producer posts the bytes of both CTAs;
//   * the leader's MMA thread alone issues tcgen05.mma.cta_group::2 (M = 256: rows 0-127 accumulate in its own TMEM, 128-255 in the
//     peer's, same column addresses; the 256-row N operand is read half from each CTA's shared memory) and signals stage-free and
//     accumulator-ready with tcgen05.commit.cta_group::2 ... multicast::cluster to the barriers of BOTH CTAs;
//   * the eight epilogue warps of each CTA drain their own TMEM and arrive on the leader's accumulator-free barrier (16 arrivals).
// In SwiGLU mode CTA 0 holds the gate rows of the weight tile and CTA 1 the up rows.
// ---------------------------------------------------------------------------------------------------------
struct GemmP2Smem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;      // this CTA's 128 activation rows
    static constexpr int B_BYTES = 128 * GEMM_BK * 2;          // this CTA's half of the 256 weight rows
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;      // 32 KB
    static constexpr int STAGES = 6;
    static constexpr int VEC_BYTES = 2 * 256 * 4;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + VEC_BYTES + 256 + 1024;
};

__device__ __forceinline__ void tma_load_2d_2sm(const void* desc, uint32_t bar_cluster_addr, void* smem_dst, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once all MMAs issued so far have completed) on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}

template <int BN>   // 256
__global__ void __launch_bounds__(GEMM_P_THREADS, 1)
gemm_tcgen05_persistent2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmShape shp,
                         const GemmEpilogue epi, const int tiles_m2 /* pairs of m-tiles */, const int tiles_n) {
    using S = GemmP2Smem;
    static_assert(BN == 256, "the pair tile is 256 x 256");
    constexpr int STAGES = S::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    float* s_bias = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES);
    float* s_gamma = s_bias + BN;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES + S::VEC_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull = empty_bar + STAGES;    // [2] accumulator ready (both CTAs, multicast commit)
    uint64_t* tempty = tfull + 2;            // [2] accumulator drained (leader's: 16 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const bool leader = rank == 0;
    const int pair = (int)blockIdx.x >> 1, n_pairs = (int)gridDim.x >> 1;
    const int n_tiles = tiles_m2 * tiles_n;
    const int nkb = shp.k_blocks;
    const int G = n_pairs / tiles_n > 0 ? n_pairs / tiles_n : 1;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 16); }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    // both CTAs' barriers are initialised and both allocations are done before anything is signalled across the pair
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = pair; t < n_tiles; t += n_pairs) {
                int tm, tn;
                gemm_p_tile(t, tiles_m2, tiles_n, G, tm, tn);
                const int m0 = (2 * tm + (int)rank) * GEMM_BM, n0 = tn * BN;
                const int brow = shp.swiglu_up_row ? (leader ? n0 / 2 : shp.swiglu_up_row + n0 / 2) : n0 + 128 * (int)rank;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                    uint8_t* a = ring + s * S::STAGE_BYTES;
                    uint32_t lbar;   // the leader's full barrier of this stage, as a shared::cluster address
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(lbar) : "r"(smem_u32(&full_bar[s])), "r"(0u));
                    if (leader) mbar_expect_tx(&full_bar[s], 2 * S::STAGE_BYTES);
                    tma_load_2d_2sm(&tmA, lbar, a, kb * GEMM_BK, m0);
                    tma_load_2d_2sm(&tmB, lbar, a + S::A_BYTES, kb * GEMM_BK, brow);
                }
            }
        }
    } else if (warp == 5) {
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(256, 256);
            uint32_t it = 0;
            int j = 0;
            for (int t = pair; t < n_tiles; t += n_pairs, ++j) {
                const int acc = j & 1;
                mbar_wait(&tempty[acc], ((j >> 1) & 1) ^ 1);   // both CTAs' epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(ring + s * S::STAGE_BYTES);
                    const uint64_t da = make_kmajor_desc(a_addr, shp.desc_lbo, shp.desc_sbo, shp.desc_layout);
                    const uint64_t db = make_kmajor_desc(a_addr + S::A_BYTES, shp.desc_lbo, shp.desc_sbo, shp.desc_layout);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k)
                        umma_f16_2sm(d_tmem, da + (uint64_t)(shp.desc_kadv * k), db + (uint64_t)(shp.desc_kadv * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit_2sm(&empty_bar[s]);
                }
                umma_commit_2sm(&tfull[acc]);
            }
        }
    } else {
        const int q = warp & 3;
        const int half = warp < 4 ? 0 : 1;
        const int et = (warp < 4 ? warp : warp - 2) * 32 + lane;   // 0..255
        uint32_t lte[2];   // the leader's accumulator-drained barriers
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(lte[0]) : "r"(smem_u32(&tempty[0])), "r"(0u));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(lte[1]) : "r"(smem_u32(&tempty[1])), "r"(0u));
        int j = 0;
        for (int t = pair; t < n_tiles; t += n_pairs, ++j) {
            const int acc = j & 1;
            int tm, tn;
            gemm_p_tile(t, tiles_m2, tiles_n, G, tm, tn);
            const int m0 = (2 * tm + (int)rank) * GEMM_BM, n0 = tn * BN;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const bool sw = shp.swiglu_up_row != 0;
            for (int i = et; i < BN; i += 256) {
                const int f = n0 + i;
                s_bias[i] = (!sw && epi.bias && f < epi.F) ? epi.bias[f] : 0.f;
                s_gamma[i] = (!sw && epi.gamma && f < epi.F) ? epi.gamma[f] : 1.f;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(&tfull[acc], (j >> 1) & 1);
            tc_fence_after();
            gemm_p_epilogue_tile<BN>(shp, epi, tmem_base, acc, m0, n0, q, half, lane, s_bias, s_gamma, lte[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    // neither CTA may release its half of the pair's tensor memory (or exit, while the peer may still signal its barriers) before both are done
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------------
// K-major fp16 matrix [rows, K] with row pitch `ld` elements -> 2-D tensor map, box = 64 x box_rows, 128B swizzle.
// Overlapping rows (ld < K, used for the k3/k7 convolutions as im2col windows) are legal: strides only need to be
// multiples of 16 bytes.
int make_tmap_kmajor(CUtensorMap* out, const void* base, long long rows, long long K, long long ld_elems, int box_rows);

struct GemmLaunch {
    const void* A; long long a_rows; long long lda;   // M operand
    const void* B; long long b_rows; long long ldb;   // N operand
    long long K;
    int block_n;   // 32, 64, 128, 256
    int split_k;
    GemmEpilogue epi;
};
int gemm_launch(const GemmLaunch& g, cudaStream_t stream);
// Prefill gate|up projection with the SwiGLU epilogue: out[t][f] = fp16(silu(A[t] . Wgu[f]) * (A[t] . Wgu[I + f])), Wgu = [2I][K] packed
// gate rows then up rows (I % 128 == 0, out pitch ldo % 8 == 0).  Persistent 128x256 tcgen05 kernel.
int gemm_launch_swiglu(const void* A, long long a_rows, long long lda, const void* Wgu, long long I, long long K, __half* out, long long ldo,
                       cudaStream_t stream);
// Variant with prebuilt tensor maps (decode path: maps are created once at bind time).
int gemm_launch_maps(const CUtensorMap& tmA, const CUtensorMap& tmB, long long a_rows, long long b_rows, long long K,
                     int block_n, int split_k, const GemmEpilogue& epi, cudaStream_t stream, const void* pf_ptr = nullptr,
                     unsigned long long pf_bytes = 0, bool pdl = false, float* zero_ptr = nullptr, unsigned long long zero_f4 = 0,
                     unsigned long long* trace = nullptr);
// Decode QKV / gate|up GEMM with RMSNorm folded in (XNORM kernel): tmX is an fp32 map over the residual stream (make_tmap_f32).
// `extra` carries norm_w (+ ss_out) for xmode 1, ss_in / ss_dim / eps / up_off for xmode 2 (XSILU: tmX maps the gate|up
// accumulator), zero_ptr / zero_f4 for either, and the L2 prefetch region.
int gemm_launch_x(int xmode, const CUtensorMap& tmA, const CUtensorMap& tmX, long long a_rows, long long T, long long K, int split_k,
                  const GemmEpilogue& epi, const GemmShape& extra, cudaStream_t stream, bool pdl);
// fp32 matrix [rows, K] with row pitch `ld` elements -> 2-D tensor map, box = 64 x box_rows, no swizzle.
int make_tmap_f32(CUtensorMap* out, const void* base, long long rows, long long K, long long ld_elems, int box_rows);
int make_tmap_k32_sw64(CUtensorMap* out, const void* base, long long rows, long long K, long long ld_elems, int box_rows);
int gemm_init();  // resolves cuTensorMapEncodeTiled, sets max dynamic smem attributes

}  // namespace ctp
