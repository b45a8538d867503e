// Decode "layer chain" kernel (sm_100a): the four dependent skinny GEMMs between two attention calls of the decode step
//
//     o_proj (+ residual)  ->  gate|up (post_attention_layernorm folded)  ->  down (SiLU gate folded, + residual)
//                          ->  q|k|v of the NEXT layer (its input_layernorm folded)            (llama.py:663-666,689-749,196-216)
//
// as ONE persistent launch instead of four.  One CTA per SM; every CTA owns at most one (128-row weight tile, k-slice) unit per
// GEMM:
//   * ALL weight tiles of the CTA's four units (<= 9 x 16 KB) are requested by TMA before griddepcontrol.wait — they do not depend
//     on anything — and stay in shared memory until their phase runs: the weight stream never sits on the dependent path;
//   * same swap-AB tcgen05 form as gemm.cuh (weights = 128-row M operand, the <= 32 tokens = N), one 32-column TMEM accumulator
//     per phase, TMEM allocated and mbarriers initialised once per launch, every mbarrier used exactly once (parity 0);
//   * the token operand of phases 1-3 is built by the four converter warps straight from the L2-resident fp32 data (coalesced
//     ld.global.cg, every load of the phase in flight at once, converted in registers to the swizzled fp16 MMA operand; the norm
//     weights were fetched before the wait).  Measured with the in-kernel stamps (profiles/README.md): a TMA fetch of the same
//     tiles costs 0.85 us of latency and the separate shared-memory conversion pass with its dependent loads another 1.9 us;
//   * split-K partial tiles are reduced with red.global.add.v4.f32 into L2-resident fp32 accumulators (the residual stream itself
//     for o_proj / down);
//   * phases are separated by grid-wide counters, not kernel boundaries: after its REDs a CTA does  bar.sync ; fence.acq_rel.gpu ;
//     red.add flag[p]  and one thread of every CTA spins on  ld.relaxed.gpu flag[p] >= (epoch+1) * gridDim  (+ one acquire fence)
//     before it releases the CTA's converter warps.  The counters are monotonic; the epoch lives next to them in device memory,
//     so the captured step graph replays unchanged;
//   * before the wait the kernel also warms L2 with the K/V streams the NEXT attention launch reads (HBM is otherwise idle while
//     the chain runs out of shared memory).
// All waits are bounded (trap instead of hanging the GPU).
#pragma once
#include "gemm.cuh"

namespace ctp {

constexpr int LK_THREADS = 192;
constexpr int LK_MAX_KB = 4;                                  // k-blocks per unit (token operand / fp32 landing tiles)
constexpr int LK_W_SLOTS = 9;                                 // weight k-blocks per CTA over the four phases (1 + 4 + 2 + 2)
constexpr int LK_A_BYTES = GEMM_BM * GEMM_BK * 2;             // 16 KB weight tile
constexpr int LK_B_BYTES = 32 * GEMM_BK * 2;                  // 4 KB fp16 token operand tile
constexpr int LK_X_BYTES = 4 * 32 * 36 * 4;                   // epilogue transpose scratch: [4 warps][32 tokens][32 features + 4 pad] fp32
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
    // L2 warm-up, issued once this launch's own weight tiles have landed (phase 1 running; HBM is idle while the chain works out
    // of shared memory): the weight stream of the NEXT layer-chain launch, dealt over the CTAs
    const void* pf_ptr[5];
    unsigned long long pf_bytes[5];
    // optional L2 warm-up of the K/V streams of the next attention launch: stream s of plane kv_k / kv_v starts at s * kv_stream_bytes,
    // its live slots are [pad_len[s / nH], *cur_len) x 128 B; at most kv_cap bytes per stream
    const char* kv_k; const char* kv_v;
    unsigned long long kv_stream_bytes, kv_cap;
    int kv_streams, nH;
    const int* pad_len; const int* cur_len;
    unsigned long long* trace;  // bring-up timeline record or null
    unsigned long long* dbg;    // bring-up: [2][64] fine-grained stamps of CTA 0 and the last CTA, or null
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
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if ((int)(v - target) >= 0) break;
        if (clock64() - t0 > 2000000000LL) {
            printf("ctp: layer-chain flag wait timeout (block %d, flag value %u, target %u)\n", blockIdx.x, v, target);
            __trap();
        }
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");   // one acquire fence after the relaxed polls (an acquire load per poll invalidates L1 every time)
}

// bring-up: fine-grained stamps of one CTA (dbg[slot] = globaltimer ns); null in production
__device__ __forceinline__ void lk_stamp(unsigned long long* dbg, int slot) {
    if (dbg) dbg[slot] = gtime_ns();
}

// Token operand of phases 1 and 3 (XNORM): B[t][k] = fp16(x[t][k] * w[k]) for this unit's k-blocks, plus sum(x^2) per row over the
// slice.  Thread (c16 = et & 15, r0 = et >> 4) owns k columns 4*c16..+3 of rows r0 + 8j of every k-block: 16 consecutive threads
// read 256 contiguous bytes.  fp16 operand layout = 128B-swizzled K-major (16-byte chunk c of row r at r*128 + ((c ^ (r & 7)) << 4)).
__device__ __forceinline__ void lk_operand_norm(const float* __restrict__ x, int ldx, int T, const LkUnit& u, const float4 (&nw)[LK_MAX_KB],
                                                uint8_t* breg, int et, float (&ssacc)[4]) {
    const int c16 = et & 15, r0 = et >> 4;
    float4 v[LK_MAX_KB][4];
#pragma unroll
    for (int i = 0; i < LK_MAX_KB; ++i) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const int r = rr * 8 + r0;
            v[i][rr] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < u.nkb && r < T) v[i][rr] = __ldcg(reinterpret_cast<const float4*>(x + (long long)r * ldx + (u.kb0 + i) * GEMM_BK + c16 * 4));
        }
    }
#pragma unroll
    for (int i = 0; i < LK_MAX_KB; ++i) {
        if (i < u.nkb) {
            uint8_t* bt = breg + i * LK_B_BYTES;
            const float4 w4 = nw[i];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int r = rr * 8 + r0;
                const float4 a4 = v[i][rr];
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
}

// Token operand of phase 2 (XSILU): B[t][k] = fp16(silu(r*g[t][k]) * (r*u[t][k])) (llama.py:214) from the fp32 gate|up accumulator;
// r[t] = rsqrt(ss2[t] / ss_dim + eps) is the post_attention_layernorm row factor deferred from phase 1.
__device__ __forceinline__ void lk_operand_silu(const float* __restrict__ gu, int ldg, int up_off, int T, const LkUnit& u, const float* ss2,
                                                float ss_dim, float eps, uint8_t* breg, int et) {
    constexpr int KB = LK_MAX_KB / 2;
    const int c16 = et & 15, r0 = et >> 4;
    float ss[4];
    float4 g[KB][4], up[KB][4];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int t = rr * 8 + r0;
        ss[rr] = t < T ? __ldcg(ss2 + t) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < KB; ++i) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const int r = rr * 8 + r0;
            g[i][rr] = make_float4(0.f, 0.f, 0.f, 0.f);
            up[i][rr] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < u.nkb && r < T) {
                const float* src = gu + (long long)r * ldg + (u.kb0 + i) * GEMM_BK + c16 * 4;
                g[i][rr] = __ldcg(reinterpret_cast<const float4*>(src));
                up[i][rr] = __ldcg(reinterpret_cast<const float4*>(src + up_off));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < KB; ++i) {
        if (i < u.nkb) {
            uint8_t* bt = breg + i * LK_B_BYTES;
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int r = rr * 8 + r0;
                const float q = rsqrtf(ss[rr] / ss_dim + eps);
                const float4 g4 = g[i][rr], u4 = up[i][rr];
                __half2 h0 = __floats2half2_rn(silu(g4.x * q) * (u4.x * q), silu(g4.y * q) * (u4.y * q));
                __half2 h1 = __floats2half2_rn(silu(g4.z * q) * (u4.z * q), silu(g4.w * q) * (u4.w * q));
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                *reinterpret_cast<uint2*>(bt + r * 128 + (((c16 >> 1) ^ (r & 7)) << 4) + ((c16 & 1) << 3)) = pk;
            }
        }
    }
}

__global__ void __launch_bounds__(LK_THREADS, 1)
k_layer_chain(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
              const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW3,
              const __grid_constant__ CUtensorMap tmAttn, const LayerArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* wreg = smem;                                       // [LK_W_SLOTS][16 KB] weight tiles, 128B-swizzled (TMA)
    uint8_t* breg = wreg + LK_W_SLOTS * LK_A_BYTES;             // [LK_MAX_KB][4 KB] fp16 token operand of the running phase
    uint8_t* xreg = breg + LK_MAX_KB * LK_B_BYTES;              // epilogue transpose scratch
    uint64_t* wfull = reinterpret_cast<uint64_t*>(xreg + LK_X_BYTES);   // [4] weights of phase p landed
    uint64_t* bfull = wfull + 4;                                // [4] token operand of phase p complete
    uint64_t* go = bfull + 4;                                   // [4] phase p may start: its predecessor is complete grid-wide
    uint64_t* accum = go + 4;                                   // [4] accumulator of phase p complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    unsigned long long* dbg = (a.dbg && (cta == 0 || cta == (int)gridDim.x - 1)) ? a.dbg + (cta == 0 ? 0 : 64) : nullptr;
    if (threadIdx.x == 0) { trace_mark(a.trace, 0); lk_stamp(dbg, 0); }
    pdl_launch_dependents();

    LkUnit un[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) un[p] = (p < a.n_phases) ? lk_unit(a.ph[p], cta) : LkUnit{0, 0, 0};

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmW0); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmW3);
        tma_prefetch_desc(&tmAttn);
        for (int p = 0; p < 4; ++p) {
            mbar_init(&wfull[p], 1);
            mbar_init(&bfull[p], p == 0 ? 1 : 4);   // phase 0: attention tile by TMA; later phases: the four converter warps
            mbar_init(&go[p], 1);
            mbar_init(&accum[p], 1);
        }
        fence_barrier_init();
        fence_proxy_async();
        // weights first: they depend on nothing (requested before the CTA-wide set-up barrier and before griddepcontrol.wait)
        const CUtensorMap* tmW[4] = {&tmW0, &tmW1, &tmW2, &tmW3};
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            if (un[p].nkb > 0) {
                mbar_expect_tx(&wfull[p], (uint32_t)un[p].nkb * LK_A_BYTES);
                for (int i = 0; i < un[p].nkb; ++i)
                    tma_load_2d(tmW[p], &wfull[p], wreg + (a.ph[p].w_slot0 + i) * LK_A_BYTES, (un[p].kb0 + i) * GEMM_BK, un[p].m * GEMM_BM);
            }
        }
        lk_stamp(dbg, 1);   // weights requested
    }
    if (warp == 5) tmem_alloc<LK_TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            // ---- control thread: one "go" per phase
            pdl_wait();
            trace_mark(a.trace, 1);
            lk_stamp(dbg, 2);
            const unsigned int epoch = *reinterpret_cast<volatile const unsigned int*>(a.flags + 3);
            const unsigned int target = (epoch + 1u) * gridDim.x;
            // phase 0: attention output rows (fp16, plain-stored by the attention kernel) straight into the operand tiles
            if (un[0].nkb > 0) {
                mbar_expect_tx(&bfull[0], (uint32_t)un[0].nkb * LK_B_BYTES);
                for (int i = 0; i < un[0].nkb; ++i) tma_load_2d(&tmAttn, &bfull[0], breg + i * LK_B_BYTES, (un[0].kb0 + i) * GEMM_BK, 0);
            }
            mbar_arrive(&go[0]);   // the epoch has been read: this CTA may arrive on the phase counters
#pragma unroll
            for (int p = 1; p < 4; ++p) {
                lk_flag_wait(a.flags + (p - 1), target);
                mbar_arrive(&go[p]);
                trace_mark(a.trace, 3 + p);
                lk_stamp(dbg, 8 + 8 * p);       // predecessor observed complete
                if (p == 1 && cta == 0) *reinterpret_cast<volatile unsigned int*>(a.flags + 3) = epoch + 1u;   // every CTA has read the epoch
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
                    lk_stamp(dbg, 8 + 8 * p + 1);   // weights of the phase in shared memory
                    mbar_wait(&bfull[p], 0);
                    lk_stamp(dbg, 8 + 8 * p + 3);   // token operand complete
                    tc_fence_after();
                    for (int i = 0; i < un[p].nkb; ++i) {
                        const uint64_t da = make_kmajor_desc(smem_u32(wreg + (a.ph[p].w_slot0 + i) * LK_A_BYTES), 1, 64, 2);
                        const uint64_t db = make_kmajor_desc(smem_u32(breg + i * LK_B_BYTES), 1, 64, 2);
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k)
                            umma_f16(tmem_base + (uint32_t)(p * 32), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (i > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&accum[p]);
                    lk_stamp(dbg, 8 + 8 * p + 4);   // MMAs issued
                }
            }
        } else if (lane == 1) {
            // ---- an idle lane of the MMA warp: L2 warm-up for the next launches (no dependency on the previous kernel)
            mbar_wait(&go[1], 0);   // phase 0 is done everywhere: every CTA's own weight tiles were requested long ago
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                if (a.pf_ptr[r] == nullptr) continue;
                const unsigned long long per = ((a.pf_bytes[r] + gridDim.x - 1) / gridDim.x + 127) & ~127ULL;
                const unsigned long long off = (unsigned long long)cta * per;
                if (off >= a.pf_bytes[r]) continue;
                unsigned long long n = a.pf_bytes[r] - off < per ? a.pf_bytes[r] - off : per;
                n &= ~15ULL;
                const char* src = reinterpret_cast<const char*>(a.pf_ptr[r]) + off;
                while (n > 0) {
                    const unsigned int chunk = n > 32768ULL ? 32768u : (unsigned int)n;
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(chunk) : "memory");
                    src += chunk;
                    n -= chunk;
                }
            }
            if (a.kv_k) {
                const int cur = *reinterpret_cast<const volatile int*>(a.cur_len);   // cached slots only
                for (int s = cta; s < 2 * a.kv_streams; s += gridDim.x) {
                    const int which = s >= a.kv_streams ? 1 : 0;
                    const int idx = s - which * a.kv_streams;
                    const int pad = a.pad_len[idx / a.nH];
                    unsigned long long n = (unsigned long long)(cur > pad ? cur - pad : 0) * 128ULL;
                    if (n > a.kv_cap) n = a.kv_cap;
                    const char* src = (which ? a.kv_v : a.kv_k) + (unsigned long long)idx * a.kv_stream_bytes + (unsigned long long)pad * 128ULL;
                    while (n > 0) {
                        const unsigned int chunk = n > 32768ULL ? 32768u : (unsigned int)n;
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(chunk) : "memory");
                        src += chunk;
                        n -= chunk;
                    }
                }
            }
        }
    } else {
        // ---- converter / epilogue warps 0..3 (128 threads); TMEM lanes [32*warp, 32*warp + 32)
        const int et = threadIdx.x, c16 = et & 15, r0 = et >> 4;
        // norm weights of this CTA's phase-1 / phase-3 k-blocks: constants, fetched before anything is waited for
        float4 nw1[LK_MAX_KB], nw3[LK_MAX_KB];
#pragma unroll
        for (int i = 0; i < LK_MAX_KB; ++i) {
            nw1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            nw3[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < un[1].nkb) nw1[i] = __ldg(reinterpret_cast<const float4*>(a.ln_post + (un[1].kb0 + i) * GEMM_BK + c16 * 4));
            if (i < un[3].nkb) nw3[i] = __ldg(reinterpret_cast<const float4*>(a.ln_next + (un[3].kb0 + i) * GEMM_BK + c16 * 4));
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            mbar_wait(&go[p], 0);
            if (et == 0) lk_stamp(dbg, 8 + 8 * p + 2);   // go observed by the converter warps
            const bool has = un[p].nkb > 0;
            if (p == 0 && cta == (int)gridDim.x - 1 && et < 16)   // re-arm ss1: its reader (this layer's attention) has completed
                reinterpret_cast<float4*>(a.ss1)[et] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has && p > 0) {
                float ssacc[4] = {0.f, 0.f, 0.f, 0.f};
                if (p == 1) lk_operand_norm(a.ph[0].out, a.ph[0].ldo, a.T, un[1], nw1, breg, et, ssacc);
                else if (p == 2) lk_operand_silu(a.ph[1].out, a.ph[1].ldo, a.I, a.T, un[2], a.ss2, a.ss_dim, a.eps, breg, et);
                else lk_operand_norm(a.ph[0].out, a.ph[0].ldo, a.T, un[3], nw3, breg, et, ssacc);
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
            if (p == 3 && a.rearm_ptr) {   // every phase-2 unit has read ss2 / gate|up: re-arm them for the next layer (off the MMA's path)
                const unsigned long long per = (a.rearm_f4 + gridDim.x - 1) / gridDim.x;
                const unsigned long long lo = (unsigned long long)cta * per, hi = (lo + per < a.rearm_f4) ? lo + per : a.rearm_f4;
                float4* z = reinterpret_cast<float4*>(a.rearm_ptr);
                for (unsigned long long q = lo + et; q < hi; q += 128) z[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (has) {
                // accumulator tile [128 features][32 tokens] -> transposed through shared memory so that each lane owns 4 consecutive
                // features of one token -> vector REDs into the L2-resident fp32 target
                mbar_wait(&accum[p], 0);
                if (et == 0) lk_stamp(dbg, 8 + 8 * p + 5);   // accumulator complete
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
                __syncwarp();   // the scratch is reused by the next phase
            }
            if (p < 3) {
                asm volatile("bar.sync 1, 128;" ::: "memory");     // all REDs / re-arm stores of this CTA issued
                if (et == 0) {
                    lk_stamp(dbg, 8 + 8 * p + 6);                  // REDs issued
                    asm volatile("fence.acq_rel.gpu;" ::: "memory");   // cumulative: orders the whole CTA's writes before the arrival
                    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(a.flags + p) : "memory");
                    lk_stamp(dbg, 8 + 8 * p + 7);                  // arrived
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc<LK_TMEM_COLS>(tmem_base);
    if (threadIdx.x == 0) { trace_end(a.trace); lk_stamp(dbg, 3); }
}

}  // namespace ctp
