// Decode "layer chain" kernel (sm_100a): the four dependent skinny GEMMs between two attention calls of the decode step
//
//     o_proj (+ residual)  ->  gate|up (post_attention_layernorm folded)  ->  down (SiLU gate folded, + residual)
//                          ->  q|k|v of the NEXT layer (its input_layernorm folded)            (llama.py:663-666,689-749,196-216)
//
// as ONE persistent launch instead of four (each of which cost ~4.5 us for 1-9 MB of weights: PDL hand-off, TMEM allocation and
// barrier set-up, first read of RED-produced lines, tail).  One CTA per SM; every CTA owns at most one (128-row weight tile,
// k-slice) unit per GEMM:
//   * ALL weight tiles of the CTA's four units (<= 9 x 16 KB) are requested by TMA before griddepcontrol.wait — they do not depend
//     on anything — and stay in shared memory until their phase runs: the weight stream never sits on the dependent path;
//   * same swap-AB tcgen05 form as gemm.cuh (weights = 128-row M operand, the <= 32 tokens = N), one 32-column TMEM accumulator
//     per phase, TMEM allocated and mbarriers initialised once per launch, every mbarrier used exactly once (parity 0);
//   * the token operand of each phase is built in the kernel from fp32 data fetched by TMA (XNORM / XSILU converters of gemm.cuh);
//   * split-K partial tiles are reduced with red.global.add.v4.f32 into L2-resident fp32 accumulators (the residual stream itself
//     for o_proj / down);
//   * phases are separated by grid-wide counters, not kernel boundaries: after its REDs a CTA does  bar.sync ; fence.acq_rel.gpu ;
//     red.add flag[p]  and the TMA thread of every CTA spins on  ld.acquire.gpu flag[p] >= (epoch+1) * gridDim  before it requests
//     the next phase's operand (fence.proxy.async between the acquire and the TMA reads).  The counters are monotonic; the epoch
//     lives next to them in device memory, so the captured step graph replays unchanged.
// All waits are bounded (trap instead of hanging the GPU).
#pragma once
#include "gemm.cuh"

namespace ctp {

constexpr int LK_THREADS = 192;
constexpr int LK_MAX_KB = 4;                                  // k-blocks per unit (token operand / fp32 landing tiles)
constexpr int LK_W_SLOTS = 9;                                 // weight k-blocks per CTA over the four phases (1 + 4 + 2 + 2)
constexpr int LK_A_BYTES = GEMM_BM * GEMM_BK * 2;             // 16 KB weight tile
constexpr int LK_B_BYTES = 32 * GEMM_BK * 2;                  // 4 KB fp16 token operand tile
constexpr int LK_X_TILE = 32 * GEMM_BK * 4;                   // 8 KB fp32 landing tile
constexpr int LK_X_BYTES = 4 * LK_X_TILE;                     // 32 KB (also the epilogue's transpose scratch)
constexpr int LK_SMEM = LK_W_SLOTS * LK_A_BYTES + LK_MAX_KB * LK_B_BYTES + LK_X_BYTES + 256 + 1024;
constexpr unsigned LK_TMEM_COLS = 128;                        // four 32-column fp32 accumulators

struct LkPhase {      // one GEMM of the chain: out[t][f] += sum_k W[f][k] * operand[t][k]
    int m_tiles, k_blocks, splits;   // units = m_tiles * splits (<= gridDim); unit u: m = u / splits, k-slice s = u % splits
    int w_slot0;                     // first weight slot of this phase in shared memory
    float* out;                      // RED target, element (t, f) at out[t * ldo + f]
    int ldo;
};

struct LayerArgs {
    LkPhase ph[4];
    int n_phases;               // 4, or 3 for the last layer (no next q|k|v)
    int T;                      // live token rows (<= 32)
    int I;                      // MLP width: the up half of the gate|up accumulator starts at column I
    float ss_dim, eps;          // RMSNorm: rsqrt(sum(x^2) / ss_dim + eps)
    const float* ln_post;       // [H] post_attention_layernorm weight (phase 1 operand)
    const float* ln_next;       // [H] next layer's input_layernorm weight (phase 3 operand)
    float* ss1;                 // [64] sum(x^2) per row as seen by phase 3 (read by the next attention); re-armed in phase 0
    float* ss2;                 // [64] ... as seen by phase 1 (read by phase 2); re-armed, with the gate|up accumulator, in phase 3
    float* rearm_ptr;           // 16-byte aligned region zeroed in phase 3: ss2 | gate|up rows of the live batch
    unsigned long long rearm_f4;
    unsigned int* flags;        // [0..2] phase counters (monotonic), [3] epoch
    unsigned long long* trace;  // bring-up timeline record or null
};

struct LkUnit { int m, kb0, nkb; };   // nkb == 0: this CTA has no unit in the phase

__device__ __forceinline__ LkUnit lk_unit(const LkPhase& p, int cta) {
    LkUnit u{0, 0, 0};
    if (cta < p.m_tiles * p.splits) {
        u.m = cta / p.splits;
        const int s = cta - u.m * p.splits;
        u.kb0 = (p.k_blocks * s) / p.splits;
        u.nkb = (p.k_blocks * (s + 1)) / p.splits - u.kb0;
    }
    return u;
}

// spin until *f >= target (wrap-safe), acquire at gpu scope; bounded
__device__ __forceinline__ void lk_flag_wait(const unsigned int* f, unsigned int target) {
    const long long t0 = clock64();
    for (;;) {
        unsigned int v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if ((int)(v - target) >= 0) return;
        if (clock64() - t0 > 2000000000LL) {
            printf("ctp: layer-chain flag wait timeout (block %d, flag value %u, target %u)\n", blockIdx.x, v, target);
            __trap();
        }
    }
}

__global__ void __launch_bounds__(LK_THREADS, 1)
k_layer_chain(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW3,
              const __grid_constant__ CUtensorMap tmAttn, const __grid_constant__ CUtensorMap tmX,
              const __grid_constant__ CUtensorMap tmGU, const LayerArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* wreg = smem;                                       // [LK_W_SLOTS][16 KB] weight tiles, 128B-swizzled (TMA)
    uint8_t* breg = wreg + LK_W_SLOTS * LK_A_BYTES;             // [LK_MAX_KB][4 KB] fp16 token operand of the running phase
    uint8_t* xreg = breg + LK_MAX_KB * LK_B_BYTES;              // fp32 landing tiles of the running phase / epilogue scratch
    uint64_t* wfull = reinterpret_cast<uint64_t*>(xreg + LK_X_BYTES);   // [4] weights of phase p landed
    uint64_t* bfull = wfull + 4;                                // [4] token operand of phase p complete
    uint64_t* xfull = bfull + 4;                                // [4] phase p may start: predecessor complete grid-wide, fp32 tiles landed
    uint64_t* accum = xfull + 4;                                // [4] accumulator of phase p complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    if (threadIdx.x == 0) trace_mark(a.trace, 0);
    pdl_launch_dependents();

    LkUnit un[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) un[p] = (p < a.n_phases) ? lk_unit(a.ph[p], cta) : LkUnit{0, 0, 0};

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmW0); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmW3);
        tma_prefetch_desc(&tmAttn); tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmGU);
        for (int p = 0; p < 4; ++p) {
            mbar_init(&wfull[p], 1);
            mbar_init(&bfull[p], p == 0 ? 1 : 4);   // phase 0: attention tile by TMA; later phases: the four converter warps
            mbar_init(&xfull[p], 1);
            mbar_init(&accum[p], 1);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 5) tmem_alloc<LK_TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            // ---- TMA thread: weights first (independent of everything), then one operand request per phase
            const CUtensorMap* tmW[4] = {&tmW0, &tmW1, &tmW2, &tmW3};
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                if (un[p].nkb > 0) {
                    mbar_expect_tx(&wfull[p], (uint32_t)un[p].nkb * LK_A_BYTES);
                    for (int i = 0; i < un[p].nkb; ++i)
                        tma_load_2d(tmW[p], &wfull[p], wreg + (a.ph[p].w_slot0 + i) * LK_A_BYTES, (un[p].kb0 + i) * GEMM_BK, un[p].m * GEMM_BM);
                }
            }
            pdl_wait();
            trace_mark(a.trace, 1);
            const unsigned int epoch = *reinterpret_cast<volatile const unsigned int*>(a.flags + 3);
            const unsigned int target = (epoch + 1u) * gridDim.x;
            // phase 0: attention output rows (fp16, plain-stored by the attention kernel) straight into the operand tiles
            if (un[0].nkb > 0) {
                mbar_expect_tx(&bfull[0], (uint32_t)un[0].nkb * LK_B_BYTES);
                for (int i = 0; i < un[0].nkb; ++i) tma_load_2d(&tmAttn, &bfull[0], breg + i * LK_B_BYTES, (un[0].kb0 + i) * GEMM_BK, 0);
            }
            mbar_arrive(&xfull[0]);   // the epoch has been read: this CTA may arrive on the phase counters
#pragma unroll
            for (int p = 1; p < 4; ++p) {
                lk_flag_wait(a.flags + (p - 1), target);
                trace_mark(a.trace, 3 + p);
                if (p == 1 && cta == 0) *reinterpret_cast<volatile unsigned int*>(a.flags + 3) = epoch + 1u;   // every CTA has read the epoch
                asm volatile("fence.proxy.async;" ::: "memory");   // acquired generic-proxy writes (REDs) -> visible to the TMA reads below
                if (un[p].nkb > 0) {
                    const int nkb = un[p].nkb, kb0 = un[p].kb0;
                    if (p == 2) {   // gate and up tiles of the fp32 gate|up accumulator
                        mbar_expect_tx(&xfull[p], (uint32_t)nkb * 2 * LK_X_TILE);
                        for (int i = 0; i < nkb; ++i) {
                            tma_load_2d(&tmGU, &xfull[p], xreg + (2 * i) * LK_X_TILE, (kb0 + i) * GEMM_BK, 0);
                            tma_load_2d(&tmGU, &xfull[p], xreg + (2 * i + 1) * LK_X_TILE, a.I + (kb0 + i) * GEMM_BK, 0);
                        }
                    } else {        // residual stream tiles
                        mbar_expect_tx(&xfull[p], (uint32_t)nkb * LK_X_TILE);
                        for (int i = 0; i < nkb; ++i) tma_load_2d(&tmX, &xfull[p], xreg + i * LK_X_TILE, (kb0 + i) * GEMM_BK, 0);
                    }
                } else {
                    mbar_arrive(&xfull[p]);
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            // ---- MMA thread
            constexpr uint32_t idesc = make_idesc_f16(GEMM_BM, 32);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                if (un[p].nkb > 0) {
                    mbar_wait(&wfull[p], 0);
                    mbar_wait(&bfull[p], 0);
                    tc_fence_after();
                    for (int i = 0; i < un[p].nkb; ++i) {
                        const uint64_t da = make_kmajor_desc(smem_u32(wreg + (a.ph[p].w_slot0 + i) * LK_A_BYTES), 1, 64, 2);
                        const uint64_t db = make_kmajor_desc(smem_u32(breg + i * LK_B_BYTES), 1, 64, 2);
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k)
                            umma_f16(tmem_base + (uint32_t)(p * 32), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (i > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&accum[p]);
                }
            }
        }
    } else {
        // ---- converter / epilogue warps 0..3 (128 threads); TMEM lanes [32*warp, 32*warp + 32)
        const int et = threadIdx.x, c16 = et & 15, r0 = et >> 4;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            mbar_wait(&xfull[p], 0);
            const bool has = un[p].nkb > 0;
            if (p == 0 && cta == (int)gridDim.x - 1 && et < 16)   // re-arm ss1: its reader (this layer's attention) has completed
                reinterpret_cast<float4*>(a.ss1)[et] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p == 3 && a.rearm_ptr) {   // every phase-2 unit has read ss2 / gate|up: re-arm them for the next layer
                const unsigned long long per = (a.rearm_f4 + gridDim.x - 1) / gridDim.x;
                const unsigned long long lo = (unsigned long long)cta * per, hi = (lo + per < a.rearm_f4) ? lo + per : a.rearm_f4;
                float4* z = reinterpret_cast<float4*>(a.rearm_ptr);
                for (unsigned long long q = lo + et; q < hi; q += 128) z[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (has && p > 0) {
                // fp32 tile(s) landed by TMA ([32 rows][64 k] row-major) -> fp16 operand in the 128B-swizzled K-major layout (16-byte
                // chunk c of row r at r*128 + ((c ^ (r & 7)) << 4)); thread (c16, r0) owns k columns 4*c16..+3 of rows r0 + 8j
                const float* nw = (p == 1) ? a.ln_post : a.ln_next;
                float rf[4], ssacc[4];
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    ssacc[rr] = 0.f;
                    rf[rr] = 1.f;
                    const int t = rr * 8 + r0;
                    if (p == 2 && t < a.T) rf[rr] = rsqrtf(__ldcg(a.ss2 + t) / a.ss_dim + a.eps);   // deferred post_attention_layernorm row factor
                }
                for (int i = 0; i < un[p].nkb; ++i) {
                    uint8_t* bt = breg + i * LK_B_BYTES;
                    if (p == 2) {
                        const float* gt = reinterpret_cast<const float*>(xreg + (2 * i) * LK_X_TILE);
                        const float* ut = reinterpret_cast<const float*>(xreg + (2 * i + 1) * LK_X_TILE);
#pragma unroll
                        for (int rr = 0; rr < 4; ++rr) {
                            const int r = rr * 8 + r0;
                            const float4 g4 = *reinterpret_cast<const float4*>(gt + r * GEMM_BK + c16 * 4);
                            const float4 u4 = *reinterpret_cast<const float4*>(ut + r * GEMM_BK + c16 * 4);
                            const float q = rf[rr];
                            __half2 h0 = __floats2half2_rn(silu(g4.x * q) * (u4.x * q), silu(g4.y * q) * (u4.y * q));
                            __half2 h1 = __floats2half2_rn(silu(g4.z * q) * (u4.z * q), silu(g4.w * q) * (u4.w * q));
                            uint2 pk;
                            pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                            *reinterpret_cast<uint2*>(bt + r * 128 + (((c16 >> 1) ^ (r & 7)) << 4) + ((c16 & 1) << 3)) = pk;
                        }
                    } else {
                        const float* xt = reinterpret_cast<const float*>(xreg + i * LK_X_TILE);
                        const float4 w4 = __ldg(reinterpret_cast<const float4*>(nw + (un[p].kb0 + i) * GEMM_BK + c16 * 4));
#pragma unroll
                        for (int rr = 0; rr < 4; ++rr) {
                            const int r = rr * 8 + r0;
                            const float4 a4 = *reinterpret_cast<const float4*>(xt + r * GEMM_BK + c16 * 4);
                            ssacc[rr] += a4.x * a4.x + a4.y * a4.y + a4.z * a4.z + a4.w * a4.w;
                            // un-normalised operand: saturate instead of overflowing fp16 should a checkpoint carry a massive activation
                            const float v0 = fminf(fmaxf(a4.x * w4.x, -65504.f), 65504.f), v1 = fminf(fmaxf(a4.y * w4.y, -65504.f), 65504.f);
                            const float v2 = fminf(fmaxf(a4.z * w4.z, -65504.f), 65504.f), v3 = fminf(fmaxf(a4.w * w4.w, -65504.f), 65504.f);
                            __half2 h0 = __floats2half2_rn(v0, v1), h1 = __floats2half2_rn(v2, v3);
                            uint2 pk;
                            pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                            *reinterpret_cast<uint2*>(bt + r * 128 + (((c16 >> 1) ^ (r & 7)) << 4) + ((c16 & 1) << 3)) = pk;
                        }
                    }
                }
                fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&bfull[p]);
                if (p != 2 && un[p].m == 0) {   // every k-slice of m-tile 0 adds its share of sum(x^2) once
                    float* ss_out = (p == 1) ? a.ss2 : a.ss1;
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        float v = ssacc[rr];
                        v += __shfl_xor_sync(0xffffffffu, v, 1);
                        v += __shfl_xor_sync(0xffffffffu, v, 2);
                        v += __shfl_xor_sync(0xffffffffu, v, 4);
                        v += __shfl_xor_sync(0xffffffffu, v, 8);
                        const int t = rr * 8 + r0;
                        if (c16 == 0 && t < a.T) atomicAdd(ss_out + t, v);
                    }
                }
            }
            if (has) {
                // accumulator tile [128 features][32 tokens] -> transposed through shared memory so that each lane owns 4 consecutive
                // features of one token -> vector REDs into the L2-resident fp32 target
                mbar_wait(&accum[p], 0);
                tc_fence_after();
                float acc[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(p * 32), acc);
                float* tile = reinterpret_cast<float*>(xreg) + warp * (32 * 36);   // [32 tokens][32 features + 4 pad]
#pragma unroll
                for (int j = 0; j < 32; ++j) tile[j * 36 + lane] = acc[j];
                __syncwarp();
                const int f4 = (lane & 7) * 4;
                float* out = a.ph[p].out + (un[p].m * GEMM_BM + warp * 32 + f4);
                const int ldo = a.ph[p].ldo;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int tl = 4 * r + (lane >> 3);
                    if (tl < a.T) {
                        const float4 v = *reinterpret_cast<const float4*>(tile + tl * 36 + f4);
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + (long long)tl * ldo), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                    }
                }
            }
            if (p < 3) {
                fence_proxy_async();                               // scratch writes above vs the next phase's TMA writes into xreg
                asm volatile("bar.sync 1, 128;" ::: "memory");     // all REDs / re-arm stores of this CTA issued
                if (et == 0) {
                    __threadfence();                               // cumulative: orders the whole CTA's writes before the arrival
                    atomicAdd(a.flags + p, 1u);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc<LK_TMEM_COLS>(tmem_base);
    if (threadIdx.x == 0) trace_end(a.trace);
}

}  // namespace ctp
