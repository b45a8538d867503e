// Decode-step GEMM for sm_100a:  out[t][f] (+)= sum_k W[f][k] * B[t][k]   for T <= 32 token rows.
//
// Measured on B200 (tests/prof_trace.py, profiles/README.md): at batch 32 the decode step is a chain of short dependent
// kernels, and data produced by fp32 RED atomics costs its consumer ~1 us extra to read (a plain-stored line: ~0.25 us).
// So this kernel never accumulates through L2:
//
//   * weights are the 128-row M operand of tcgen05.mma (features on TMEM lanes), the <= 32 tokens are the N operand;
//   * split-K runs over the CTAs of ONE thread-block cluster (grid.z == cluster.z == S).  Every CTA stages its fp32 partial
//     tile in its own shared memory; after a cluster barrier CTA r owns 1/S of the tile, sums the S partials through
//     distributed shared memory, adds the residual stream when asked, and writes the FINAL values with plain stores;
//   * the token operand is either a TMA tile of an fp16 matrix (BMODE 0) or produced in place by the epilogue warps:
//       BMODE 1  RMSNorm prologue (llama.py:82-87):   B[t][k] = fp16(x[t][k] * w[k])
//                The contraction is linear in the row factor r[t] = rsqrt(mean(x[t]^2) + eps), so the factor is applied by
//                the CONSUMER of this GEMM's output; sum(x^2) per row is left behind by the GEMM that produced x (ss_out).
//       BMODE 2  SiLU-gate prologue (llama.py:214):    B[t][k] = fp16(silu(r*g[t][k]) * (r*u[t][k])),  r from ss_in
//   * PDL: barrier init, TMEM allocation and the weight TMA loads of the first ring pass are issued before
//     griddepcontrol.wait; the producing CTA also warms L2 with the next kernel's weights / K,V streams.
#pragma once
#include "ctp_common.cuh"

namespace ctp {

constexpr int DG_BM = 128, DG_BN = 32, DG_BK = 64, DG_STAGES = 4, DG_THREADS = 192;
constexpr int DG_A_BYTES = DG_BM * DG_BK * 2;          // 16 KB
constexpr int DG_B_BYTES = DG_BN * DG_BK * 2;          //  4 KB
constexpr int DG_STAGE_BYTES = DG_A_BYTES + DG_B_BYTES;
constexpr int DG_STAGE_LD = DG_BM + 4;                 // fp32 staging tile [32 tokens][128 features + 4]
constexpr int DG_SMEM = DG_STAGES * DG_STAGE_BYTES + 256 + 1024;
constexpr int DG_MAX_SS_PARTS = 8;
#ifndef DG_PROBE
#define DG_PROBE 0   // bring-up: 1 moves the timeline stamps 4..6 inside the token-operand producer
#endif

struct DecGemmArgs {
    int k_blocks;            // K / 64
    int T;                   // live token rows (<= 32)
    // ---- token operand (BMODE != 0) ----
    const float* bsrc;       // BMODE 1: residual stream x [T][ldbs];  BMODE 2: gate|up [T][ldbs] (gate at k, up at bI + k)
    long long ldbs;
    const float* bw;         // BMODE 1: norm weight [K]
    int bI;
    const float* ss_in;      // BMODE 2: [ss_parts][ss_stride] partial sums of squares of the rows the gate|up GEMM contracted; null: r = 1
    int ss_parts, ss_stride;
    float ss_dim, eps;
    // ---- output ----
    float* out;              // [T][ldo]; column m0 + f
    long long ldo;
    const float* residual;   // [T][ldr] or null (may alias out: each element is read and written by one thread)
    long long ldr;
    float* ss_out;           // [gridDim.y][ss_out_stride] or null: sum over this m-tile's 128 features of out[t][.]^2
    int ss_out_stride;
    // ---- L2 warm-up for the kernels that follow ----
    const void* pf_ptr;      // weights of the next GEMM
    unsigned long long pf_bytes;
    const void* kvpf_base;   // K plane of the layer whose attention runs next (V plane at + kvpf_plane); null = off
    unsigned long long kvpf_plane, kvpf_stream, kvpf_cap;
    int kvpf_streams, kvpf_slot_bytes;
    const int* kvpf_len;
    unsigned long long* trace;   // bring-up timeline record or null
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

template <int BMODE>
__global__ void __launch_bounds__(DG_THREADS, 2)   // <= 168 registers: the next kernel's CTA can co-reside (PDL overlap)
k_dec_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const DecGemmArgs a) {
    constexpr int STAGES = DG_STAGES;
    constexpr uint32_t TMEM_COLS = 32;
    constexpr uint32_t FULL_COUNT = BMODE ? 5 : 1;       // weight TMA (arrive.expect_tx) + 4 producer warps
    constexpr uint32_t TX_BYTES = BMODE ? DG_A_BYTES : DG_STAGE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * DG_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * DG_BM;
    const int S = gridDim.z;                              // == cluster size along z
    const int kb0 = (int)(((long long)a.k_blocks * blockIdx.z) / S);
    const int kb1 = (int)(((long long)a.k_blocks * (blockIdx.z + 1)) / S);
    const int nkb = kb1 - kb0;
    if (threadIdx.x == 0) trace_mark(a.trace, 0);
    pdl_launch_dependents();

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmA);
        if (BMODE == 0) tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], FULL_COUNT); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
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
            // weights do not depend on the previous kernel: first ring pass before the wait
            const int pre = nkb < STAGES ? nkb : STAGES;
            for (int i = 0; i < pre; ++i) {
                mbar_expect_tx(&full_bar[i], TX_BYTES);
                tma_load_2d(&tmA, &full_bar[i], ring + i * DG_STAGE_BYTES, (kb0 + i) * DG_BK, m0);
            }
            pdl_wait();
            if (BMODE == 0) {
                for (int i = 0; i < pre; ++i)
                    tma_load_2d(&tmB, &full_bar[i], ring + i * DG_STAGE_BYTES + DG_A_BYTES, (kb0 + i) * DG_BK, 0);
            }
            for (int i = pre; i < nkb; ++i) {
                const int s = i % STAGES;
                mbar_wait(&empty_bar[s], ((i / STAGES) & 1) ^ 1);
                uint8_t* st = ring + s * DG_STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], TX_BYTES);
                tma_load_2d(&tmA, &full_bar[s], st, (kb0 + i) * DG_BK, m0);
                if (BMODE == 0) tma_load_2d(&tmB, &full_bar[s], st + DG_A_BYTES, (kb0 + i) * DG_BK, 0);
            }
            const unsigned long long n_cta = (unsigned long long)gridDim.y * gridDim.z;
            const unsigned long long cta = (unsigned long long)blockIdx.z * gridDim.y + blockIdx.y;
            if (a.pf_ptr) {   // warm L2 with the next GEMM's weights: the region is dealt evenly to the CTAs
                const unsigned long long per = ((a.pf_bytes + n_cta - 1) / n_cta + 127) & ~127ULL;
                const unsigned long long off = cta * per;
                if (off < a.pf_bytes) {
                    unsigned long long n = a.pf_bytes - off < per ? a.pf_bytes - off : per;
                    n &= ~15ULL;
                    const char* src = reinterpret_cast<const char*>(a.pf_ptr) + off;
                    while (n > 0) {
                        const unsigned int chunk = n > 32768ULL ? 32768u : (unsigned int)n;
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(chunk) : "memory");
                        src += chunk;
                        n -= chunk;
                    }
                }
            }
            if (a.kvpf_base) {   // K/V streams of the next attention kernel, (b, head) streams dealt round-robin
                unsigned long long len = (unsigned long long)(*a.kvpf_len) * (unsigned long long)a.kvpf_slot_bytes;
                if (len > a.kvpf_cap) len = a.kvpf_cap;
                len &= ~15ULL;
                for (int sidx = (int)cta; sidx < 2 * a.kvpf_streams; sidx += (int)n_cta) {
                    const int plane = sidx >= a.kvpf_streams ? 1 : 0;
                    const char* src = reinterpret_cast<const char*>(a.kvpf_base) + (unsigned long long)plane * a.kvpf_plane +
                                      (unsigned long long)(sidx - plane * a.kvpf_streams) * a.kvpf_stream;
                    unsigned long long n = len;
                    while (n > 0) {
                        const unsigned int chunk = n > 32768ULL ? 32768u : (unsigned int)n;
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(chunk) : "memory");
                        src += chunk;
                        n -= chunk;
                    }
                }
            }
        } else {
            pdl_wait();
        }
    } else if (warp == 5) {
        pdl_wait();
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(DG_BM, DG_BN);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                mbar_wait(&full_bar[s], (i / STAGES) & 1);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(ring + s * DG_STAGE_BYTES);
                const uint64_t da = make_kmajor_desc(a_addr, 1, 64, 2);
                const uint64_t db = make_kmajor_desc(a_addr + DG_A_BYTES, 1, 64, 2);
#pragma unroll
                for (int k = 0; k < DG_BK / 16; ++k) umma_f16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (i > 0 || k > 0) ? 1u : 0u);
                umma_commit(&empty_bar[s]);
            }
            if (nkb > 0) umma_commit(accum_bar);
        }
    } else {
        pdl_wait();
        if (threadIdx.x == 0) trace_mark(a.trace, 1);
        const int et = threadIdx.x;   // 0..127
        // this thread's share of the reduced tile (see below): issue the residual loads now, they are consumed after the MMAs
        const int n_items = (DG_BN * DG_BM / 4) / S;      // float4 per CTA
        const int q_base = (int)cluster_ctarank() * n_items;
        constexpr int MAX_IT = (DG_BN * DG_BM / 4) / 128;   // 8 items per thread when the cluster is a single CTA
        float4 res[MAX_IT];
#pragma unroll
        for (int i = 0; i < MAX_IT; ++i) {
            res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int q = q_base + et + 128 * i;
            if (a.residual && et + 128 * i < n_items && (q >> 5) < a.T)
                res[i] = __ldcg(reinterpret_cast<const float4*>(a.residual + (long long)(q >> 5) * a.ldr + m0 + (q & 31) * 4));
        }
        if (BMODE != 0) {
            // token operand tiles: 32 rows x 64 k.  Thread (c16 = et & 15, r0 = et >> 4) owns k columns 4*c16..+3 of rows
            // r0 + 8j: a warp-wide 16-byte load covers two full 256-byte row segments (every 32-byte sector is requested once).
            // The fp16 results land in the 128B-swizzled K-major layout the TMA path would have written:
            // 16-byte chunk c of row r at r*128 + ((c ^ (r & 7)) << 4); this thread fills half a chunk (8 bytes).
            const int c16 = et & 15, r0 = et >> 4;
            constexpr int RR = DG_BN / 8;          // 4 rows per thread per k-block
            constexpr int GROUP = BMODE == 1 ? 4 : 2;
            float rs[RR];
#pragma unroll
            for (int rr = 0; rr < RR; ++rr) {
                rs[rr] = 1.f;
                const int t = rr * 8 + r0;
                if (BMODE == 2 && a.ss_in && t < a.T) {
                    float part[DG_MAX_SS_PARTS];   // independent loads: one L2 round trip, not ss_parts of them
#pragma unroll
                    for (int p = 0; p < DG_MAX_SS_PARTS; ++p) part[p] = p < a.ss_parts ? __ldcg(a.ss_in + p * a.ss_stride + t) : 0.f;
                    float ss = 0.f;
#pragma unroll
                    for (int p = 0; p < DG_MAX_SS_PARTS; ++p) ss += part[p];
                    rs[rr] = rsqrtf(ss / a.ss_dim + a.eps);
                }
            }
            for (int i0 = 0; i0 < nkb; i0 += GROUP) {
                float4 va[GROUP][RR], vb[GROUP][BMODE == 1 ? 1 : RR];
#pragma unroll
                for (int g = 0; g < GROUP; ++g) {
                    const int k = (kb0 + i0 + g) * DG_BK + c16 * 4;
                    const bool kok = (i0 + g) < nkb;
                    if (BMODE == 1) vb[g][0] = kok ? __ldg(reinterpret_cast<const float4*>(a.bw + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int rr = 0; rr < RR; ++rr) {
                        const int t = rr * 8 + r0;
                        const bool ok = kok && t < a.T;
                        const float* src = a.bsrc + (long long)t * a.ldbs + k;
                        va[g][rr] = ok ? *reinterpret_cast<const float4*>(src) : make_float4(0.f, 0.f, 0.f, 0.f);
                        if (BMODE == 2) vb[g][rr] = ok ? *reinterpret_cast<const float4*>(src + a.bI) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int g = 0; g < GROUP; ++g) {
                    const int i = i0 + g;
                    if (i < nkb) {
                        const int s = i % STAGES;
                        if (i >= STAGES) mbar_wait(&empty_bar[s], ((i / STAGES) & 1) ^ 1);
                        uint8_t* bt = ring + s * DG_STAGE_BYTES + DG_A_BYTES;
#pragma unroll
                        for (int rr = 0; rr < RR; ++rr) {
                            const int r = rr * 8 + r0;
                            const float4 a0 = va[g][rr];
                            float v0, v1, v2, v3;
                            if (BMODE == 1) {
                                const float4 w0 = vb[g][0];
                                v0 = a0.x * w0.x; v1 = a0.y * w0.y; v2 = a0.z * w0.z; v3 = a0.w * w0.w;
                            } else {
                                const float4 u0 = vb[g][rr];
                                const float q = rs[rr];
                                v0 = silu(a0.x * q) * (u0.x * q); v1 = silu(a0.y * q) * (u0.y * q);
                                v2 = silu(a0.z * q) * (u0.z * q); v3 = silu(a0.w * q) * (u0.w * q);
                            }
                            __half2 h0 = __floats2half2_rn(v0, v1), h1 = __floats2half2_rn(v2, v3);
                            uint2 pk;
                            pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                            *reinterpret_cast<uint2*>(bt + r * 128 + (((c16 >> 1) ^ (r & 7)) << 4) + ((c16 & 1) << 3)) = pk;
                        }
                        if (DG_PROBE && g == 0 && i0 == 0 && threadIdx.x == 0) trace_mark(a.trace, 4);
                    }
                }
                if (DG_PROBE && i0 == 0 && threadIdx.x == 0) trace_mark(a.trace, 5);
                fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int g = 0; g < GROUP; ++g)
                        if (i0 + g < nkb) mbar_arrive(&full_bar[(i0 + g) % STAGES]);
                }
            }
            if (threadIdx.x == 0) trace_mark(a.trace, DG_PROBE ? 6 : 4);
        }
        // ---- stage this CTA's partial tile, fp32 [32 tokens][128 features (+4)], in the (now idle) operand ring
        float acc[32];
        if (nkb > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
            tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16), acc);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = 0.f;
        }
        if (!DG_PROBE && threadIdx.x == 0) trace_mark(a.trace, 5);
        float* stage = reinterpret_cast<float*>(ring);
#pragma unroll
        for (int j = 0; j < 32; ++j) stage[j * DG_STAGE_LD + warp * 32 + lane] = acc[j];
        // (res[] stays live across the cluster barrier below)
        tc_fence_before();
        __syncwarp();
        cluster_sync_all();   // #1: every partial tile of the cluster is staged
        if (!DG_PROBE && threadIdx.x == 0) trace_mark(a.trace, 6);
        // ---- reduce-scatter through distributed shared memory: CTA r owns float4 items [r*n_items, (r+1)*n_items) of the
        //      tile (item q = token q/32, features 4*(q%32)..+3), i.e. whole token rows -> 32 consecutive items share a token
        const uint32_t stage_addr = smem_u32(stage);
#pragma unroll
        for (int i = 0; i < MAX_IT; ++i) {
            const int qi = et + 128 * i;
            if (qi < n_items) {          // warp-uniform (n_items is a multiple of 32)
                const int q = q_base + qi;
                const int t = q >> 5, f4 = q & 31;
                const uint32_t local = stage_addr + (uint32_t)((t * DG_STAGE_LD + f4 * 4) * 4);
                float4 sum = res[i];
                for (int s0 = 0; s0 < S; s0 += 8) {   // 8 remote loads in flight at a time
                    float4 p[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) p[j] = (s0 + j < S) ? ld_dsmem_f4(dsmem_addr(local, (uint32_t)(s0 + j))) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { sum.x += p[j].x; sum.y += p[j].y; sum.z += p[j].z; sum.w += p[j].w; }
                }
                if (t < a.T) *reinterpret_cast<float4*>(a.out + (long long)t * a.ldo + m0 + f4 * 4) = sum;
                if (a.ss_out) {
                    float ss = sum.x * sum.x + sum.y * sum.y + sum.z * sum.z + sum.w * sum.w;
                    ss = warp_sum(ss);
                    if (lane == 0 && t < a.T) a.ss_out[blockIdx.y * a.ss_out_stride + t] = ss;
                }
            }
        }
        if (threadIdx.x == 0) trace_mark(a.trace, 7);
    }
    if (warp >= 4) {
        tc_fence_before();
        __syncwarp();
        cluster_sync_all();   // #1 (TMA / MMA warps take part in the cluster barrier as whole warps)
    }
    __syncwarp();
    cluster_sync_all();       // #2: nobody leaves while a peer may still read its staged tile
    if (warp == 5) tmem_dealloc<TMEM_COLS>(tmem_base);
    if (threadIdx.x == 0) trace_end(a.trace);
}

// Host launcher (gpt.cu): grid (1, F/128, S), cluster (1, 1, S), optional programmatic stream serialization.
template <int BMODE>
inline cudaError_t launch_dec_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const DecGemmArgs& a, int m_tiles, int S,
                                   cudaStream_t stream, bool pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1, (unsigned)m_tiles, (unsigned)S);
    cfg.blockDim = dim3(DG_THREADS);
    cfg.dynamicSmemBytes = DG_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n = 0;
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = 1; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = (unsigned)S;
    ++n;
    if (pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr; cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, k_dec_gemm<BMODE>, tmA, tmB, a);
}

}  // namespace ctp
