// Decode-step SwiGLU MLP block of one decoder layer in ONE launch (batch <= 32), partitioned over the intermediate dimension
// the way tensor parallelism partitions it over devices (column-parallel gate|up, row-parallel down_proj), here over SMs:
//
//     x_out = x_in + down_proj( silu(gate_proj(n)) * up_proj(n) ),   n = post_attention_layernorm(x_in)      (llama.py:82-87,214,741-745)
//
// CTA u owns 32 intermediate features [32u, 32u+32): the 32 gate rows and 32 up rows of W_gu over the full K = H.  Thread-block
// clusters of C = H/128 CTAs exchange the small things through distributed shared memory instead of L2, as bulk copies
// (cp.async.bulk.shared::cluster.shared::cta, completion on the receiver's mbarrier — per-thread st.shared::cluster stores of the
// same bytes took 2.4 us, measured):
//   * the token operand: CTA c of a cluster loads columns [128c, 128c+128) of the fp32 residual tile (16 KB instead of the whole
//     96 KB), converts them to the fp16 operand w*x (two 64-wide k-blocks, 8 KB) and copies them into the other C-1 operand
//     buffers; the RMSNorm row factor is linear, so the per-row partial sums of squares travel the same way and are applied in
//     epilogue 1;
//   * silu(g)*u: each CTA's [32 tokens x 32 features] slice (2 KB) is copied into all C CTAs, so every CTA holds the cluster's
//     [32 x 32C] slice and contracts it against ITS 128 output features of W_down: the partial sums that leave the SM are
//     [32 tokens x 128 features] per CTA (16 KB of vector REDs) instead of [32 x H].
// The gate|up -> down hand-off therefore never goes through L2; the only L2 exchange of the block is the final fp32 RED into
// the residual stream (same-address depth I / (32 C)).
//
//   before griddepcontrol.wait   all of this CTA's weights (112 KB) by TMA: they do not depend on the previous kernel
//   MMA 1 (tcgen05, swap-AB)     D1[64 weight rows (gate | up), 32 tokens] over K = H, M = 64      (TMEM columns [0, 128): 4 partials)
//   epilogue 1 (8 warps)         silu(r g) * (r u) -> fp16 slice in the 64B-swizzled K-major layout, all-gathered over the cluster
//   MMA 2                        D2[128 features, 32 tokens] over K = 32 C (C slices of 32)         (TMEM columns [128, 192): 2 partials)
//   epilogue 2 (8 warps)         TMEM -> registers -> shared-memory transpose -> red.global.add.v4.f32 into x_out
//
// Measured: the 48 MMAs of the K = 768 loop take 1.75-1.9 us (~50 cycles each on top of ~0.55 us of issue/commit latency) whether
// M = 64 or 128 and whether they accumulate into one TMEM tile or into four independent ones (one per 16-wide k-step index, added
// up by the epilogue: kept, 1.9 -> 1.75 us).
//
// Tried on top of this and dropped: extra CTAs of the same launch (on the 52 SMs the MLP leaves idle) that prefetch the next
// attention launch's K/V streams into L2 — cp.async.bulk.prefetch.L2 drains at ~50 GB/s per SM (12 CTAs: 60 us for 38 MB), per-thread
// prefetch.global.L2 from 48 CTAs outlasts the MLP by 4 us, and the attention kernel was no faster with its streams L2-resident
// (7.6 vs 7.8 us body: it is bound by its dependent per-tile chain, not by HBM): 615-810 vs 569 us/step.
//
// x_out must be all-zero when the first RED lands (the launch sequence has an earlier kernel clear it); cluster g adds the rows
// r = g (mod clusters) of x_in, which carries the residual connection.
#pragma once
#include "ctp_common.cuh"

namespace ctp {

struct MlpArgs {
    const float* x_in;       // [>= 32][H] residual stream after the attention block (complete when griddepcontrol.wait returns)
    float* x_out;            // [>= 32][H] zeroed by an earlier kernel
    const float* ln_w;       // [H] post_attention_layernorm weight
    int T, H, I;
    float eps;
    int m64;                 // 1: MMA 1 as M = 64 (rows r -> TMEM lane (r & 15) + 32 (r >> 4)), 0: M = 128 with 64 idle rows
    const void* pf_ptr;      // optional: region the next GEMM streams (its weights); every CTA prefetches a share into L2
    unsigned long long pf_bytes;
    unsigned long long* trace;
};

constexpr int MLP_THREADS = 320;          // warps 0-7: operand build + epilogues, warp 8: TMA / bulk copies, warp 9: TMEM + MMA issue
constexpr int MLP_UW = 32;                // intermediate features per CTA
constexpr int MLP_XA_KB = 32 * 128;       // x operand bytes per 64-wide k-block: 32 token rows x 128 B (128B swizzle)
constexpr int MLP_GU_KB = 64 * 128;       // gate rows [0, 32), up rows [32, 64) of a 64-row tile (128B swizzle)
constexpr int MLP_D_TILE = 128 * 64;      // W_down: 128 features x 32 k (64B swizzle)
constexpr int MLP_M_SLICE = 32 * 64;      // silu*up: 32 tokens x 32 k (64B swizzle)
// cluster size C = H / 128 (<= 8); shared memory: x operand | W_gu | W_down tiles | m slices | row sums | barriers
__host__ __device__ constexpr int mlp_smem_bytes(int H) {
    return (H / 64) * (MLP_XA_KB + MLP_GU_KB) + (H / 128) * (MLP_D_TILE + MLP_M_SLICE) + 8 * 32 * 4 + 320 + 1024;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
    uint32_t d;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(d) : "r"(smem_addr), "r"(rank));
    return d;
}
// shared memory of this CTA -> shared memory of a peer (addresses from mapa), completion counted on the PEER's mbarrier
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :
                 : "r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster)
                 : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// 32 lanes x 16 columns of fp32
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(MLP_THREADS, 1)
k_mlp_fused(const __grid_constant__ CUtensorMap tmGU, const __grid_constant__ CUtensorMap tmD, const MlpArgs a) {
    extern __shared__ uint8_t mlp_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mlp_smem_raw) + 1023) & ~uintptr_t(1023));
    const int kbH = a.H >> 6;          // k-blocks of MMA 1
    const int C = a.H >> 7;            // cluster size = 128-feature tiles of the output = 32-wide K slices of MMA 2
    uint8_t* xa = smem;
    uint8_t* wgu = xa + kbH * MLP_XA_KB;
    uint8_t* wd = wgu + kbH * MLP_GU_KB;
    uint8_t* mop = wd + C * MLP_D_TILE;
    float* ss_part = reinterpret_cast<float*>(mop + C * MLP_M_SLICE);   // [C][32] per-rank partial sums of squares
    uint64_t* bars = reinterpret_cast<uint64_t*>(ss_part + 8 * 32);     // [0] weights, [1] D1, [2] D2, [3] x operand + row sums, [4] m slices
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = (int)cluster_ctarank();
    const int cl = (int)blockIdx.x / C, n_cl = (int)gridDim.x / C;
    const int i0 = (int)blockIdx.x * MLP_UW;
    if (threadIdx.x == 0) trace_mark(a.trace, 0);
    pdl_launch_dependents();
    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&tmGU);
        tma_prefetch_desc(&tmD);
        for (int i = 0; i < 3; ++i) mbar_init(&bars[i], 1);
        mbar_init(&bars[3], 2);   // the expect_tx arrival below + the arrival that publishes this CTA's own slice
        mbar_init(&bars[4], 2);
        fence_barrier_init();
        fence_proxy_async();
        // peers copy (C-1) slices into this CTA; bytes may land before or after these expectations are posted (the phase cannot
        // complete before the second arrival either way).  One barrier per exchange: a barrier per source rank, so that the MMAs
        // start on whatever has landed, measured slower (9.4-9.9 vs 8.5-9.0 us per launch)
        mbar_expect_tx(&bars[3], (uint32_t)((C - 1) * (2 * MLP_XA_KB + 128)));
        mbar_expect_tx(&bars[4], (uint32_t)((C - 1) * MLP_M_SLICE));
    }
    if (warp == 9) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    cluster_arrive();   // once the matching wait returns every CTA of the cluster is resident and its barriers are initialised

    if (warp == 8 && lane == 0) {
        mbar_expect_tx(&bars[0], (uint32_t)(kbH * MLP_GU_KB + C * MLP_D_TILE));
        for (int kb = 0; kb < kbH; ++kb) {
            tma_load_2d(&tmGU, &bars[0], wgu + kb * MLP_GU_KB, kb * 64, i0);                 // gate rows
            tma_load_2d(&tmGU, &bars[0], wgu + kb * MLP_GU_KB + 32 * 128, kb * 64, a.I + i0);  // up rows
        }
        for (int j = 0; j < C; ++j) tma_load_2d(&tmD, &bars[0], wd + j * MLP_D_TILE, (cl * C + j) * MLP_UW, c * 128);
        if (a.pf_ptr) {
            const unsigned long long n_cta = gridDim.x, cta = blockIdx.x;
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
    }
    __syncwarp();
    float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp < 8) w4 = __ldg(reinterpret_cast<const float4*>(a.ln_w + c * 128) + lane);
    pdl_wait();
    if (threadIdx.x == 0) trace_mark(a.trace, 1);

    if (warp < 8) {
        // ---- operand build: this CTA converts columns [128c, 128c+128) of the residual tile; warp w owns token rows 4w .. 4w+3
        float4 xv[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) xv[r] = __ldcg(reinterpret_cast<const float4*>(a.x_in + (long long)(warp * 4 + r) * a.H + c * 128) + lane);
        cluster_wait();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = warp * 4 + r;
            const float4 x = xv[r];
            const float ss = warp_sum(x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w);
            if (lane == 0) ss_part[c * 32 + row] = ss;
            if (row < a.T && (row % n_cl) == cl) {   // the residual connection: this cluster adds row `row` of x_in
                float* o = a.x_out + (long long)row * a.H + c * 128 + 4 * lane;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
            }
            // un-normalised operand w*x (the row factor is applied in epilogue 1): saturate instead of overflowing fp16
            const __half2 h0 = __floats2half2_rn(fminf(fmaxf(w4.x * x.x, -65504.f), 65504.f), fminf(fmaxf(w4.y * x.y, -65504.f), 65504.f));
            const __half2 h1 = __floats2half2_rn(fminf(fmaxf(w4.z * x.z, -65504.f), 65504.f), fminf(fmaxf(w4.w * x.w, -65504.f), 65504.f));
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&h0);
            pk.y = *reinterpret_cast<const uint32_t*>(&h1);
            // k = 128c + 4 lane: k-block 2c + (lane >> 4), 16-byte chunk (lane & 15) >> 1, half (lane & 1)
            *reinterpret_cast<uint2*>(xa + (2 * c + (lane >> 4)) * MLP_XA_KB + row * 128 + ((((lane & 15) >> 1) ^ (row & 7)) << 4) + ((lane & 1) << 3)) = pk;
        }
        fence_proxy_async();   // generic-proxy writes -> visible to the async proxy (bulk copies, tensor cores)
    } else {
        cluster_wait();
    }
    if (warp <= 8) asm volatile("bar.sync 4, 288;" ::: "memory");
    if (warp == 8 && lane == 0) {
        // all-gather of the operand slice and the row sums: one 8 KB and one 128 B bulk copy per peer
        const uint32_t src_x = smem_u32(xa + 2 * c * MLP_XA_KB), src_s = smem_u32(ss_part + c * 32), bar = smem_u32(&bars[3]);
        for (int d = 1; d < C; ++d) {
            const uint32_t rk = (uint32_t)((c + d) % C);
            const uint32_t rbar = mapa_u32(bar, rk);
            bulk_copy_to_peer(mapa_u32(src_x, rk), src_x, 2 * MLP_XA_KB, rbar);
            bulk_copy_to_peer(mapa_u32(src_s, rk), src_s, 128, rbar);
        }
        mbar_arrive(&bars[3]);
    }

    if (warp == 9 && lane == 0) {
        const uint32_t idesc1 = a.m64 ? make_idesc_f16(64, 32) : make_idesc_f16(128, 32);
        mbar_wait(&bars[0], 0);
        mbar_wait(&bars[3], 0);
        tc_fence_after();
        for (int kb = 0; kb < kbH; ++kb) {
            const uint64_t da = make_kmajor_desc(smem_u32(wgu + kb * MLP_GU_KB), 1, 64, 2);
            const uint64_t db = make_kmajor_desc(smem_u32(xa + kb * MLP_XA_KB), 1, 64, 2);
#pragma unroll
            for (int k = 0; k < 4; ++k)   // one accumulator per k-step index (see header)
                umma_f16(tmem_base + 32u * (uint32_t)k, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc1, kb > 0 ? 1u : 0u);
        }
        umma_commit(&bars[1]);
    }
    if (warp < 8) {
        // ---- epilogue 1: D1 rows (gate 0-31, up 32-63) -> shared memory [64][33] fp32, then 256 threads finish 4 values each
        // scratch lives in the W_gu region: dead once D1 is complete, and (unlike this CTA's x operand slice) never the source of an
        // outgoing bulk copy that a peer may still be draining
        float* gu = reinterpret_cast<float*>(wgu);
        mbar_wait(&bars[3], 0);                       // row sums of all ranks have landed
        if (threadIdx.x == 0) trace_mark(a.trace, 4);
        mbar_wait(&bars[1], 0);
        tc_fence_after();
        if (threadIdx.x == 0) trace_mark(a.trace, 5);
        if (warp < (a.m64 ? 4 : 2)) {
            float v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16), v);
#pragma unroll
            for (int k = 1; k < 4; ++k) {
                float w[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + 32u * (uint32_t)k, w);
#pragma unroll
                for (int t = 0; t < 32; ++t) v[t] += w[t];
            }
            const int row = a.m64 ? warp * 16 + lane : warp * 32 + lane;
            if (!a.m64 || lane < 16) {
#pragma unroll
                for (int t = 0; t < 32; ++t) gu[row * 33 + t] = v[t];
            }
        }
        asm volatile("bar.sync 5, 256;" ::: "memory");
        {
            const int e = threadIdx.x, t = e >> 3, fg = (e & 7) * 4;
            float ssum = 0.f;
            for (int rk = 0; rk < C; ++rk) ssum += ss_part[rk * 32 + t];
            const float rs = rsqrtf(ssum / (float)a.H + a.eps);
            float m[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) m[i] = silu(gu[(fg + i) * 33 + t] * rs) * (gu[(32 + fg + i) * 33 + t] * rs);
            const __half2 h0 = __floats2half2_rn(m[0], m[1]), h1 = __floats2half2_rn(m[2], m[3]);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&h0);
            pk.y = *reinterpret_cast<const uint32_t*>(&h1);
            // slice c, token row t (64 B), k = fg: 16-byte chunk (fg >> 3) XOR ((t >> 1) & 3), half (fg >> 2) & 1
            *reinterpret_cast<uint2*>(mop + c * MLP_M_SLICE + t * 64 + ((((fg >> 3) ^ ((t >> 1) & 3))) << 4) + (((fg >> 2) & 1) << 3)) = pk;
        }
        fence_proxy_async();
    }
    if (warp <= 8) asm volatile("bar.sync 4, 288;" ::: "memory");
    if (threadIdx.x == 0) trace_mark(a.trace, 6);
    if (warp == 8 && lane == 0) {
        const uint32_t src_m = smem_u32(mop + c * MLP_M_SLICE), bar = smem_u32(&bars[4]);
        for (int d = 1; d < C; ++d) {
            const uint32_t rk = (uint32_t)((c + d) % C);
            bulk_copy_to_peer(mapa_u32(src_m, rk), src_m, MLP_M_SLICE, mapa_u32(bar, rk));
        }
        mbar_arrive(&bars[4]);
    }
    if (warp == 9 && lane == 0) {
        constexpr uint32_t idesc2 = make_idesc_f16(128, 32);
        mbar_wait(&bars[4], 0);
        tc_fence_after();
        for (int j = 0; j < C; ++j) {
            const uint64_t dw = make_kmajor_desc(smem_u32(wd + j * MLP_D_TILE), 1, 32, 4);   // 64B swizzle: 8-row groups 512 B apart
            const uint64_t dm = make_kmajor_desc(smem_u32(mop + j * MLP_M_SLICE), 1, 32, 4);
#pragma unroll
            for (int k = 0; k < 2; ++k) umma_f16(tmem_base + 128u + 32u * (uint32_t)k, dw + (uint64_t)(2 * k), dm + (uint64_t)(2 * k), idesc2, j > 0 ? 1u : 0u);
        }
        umma_commit(&bars[2]);
    }
    __syncwarp();
    // Every slice a peer copied into this CTA has landed before this CTA's MMA thread arrives here, so once the matching wait returns
    // all of THIS CTA's outgoing copies have been read out of its shared memory and it may exit.
    cluster_arrive();
    if (warp < 8) {
        // ---- epilogue 2: warp w reads TMEM lane quarter (w & 3), token half (w >> 2), of this CTA's 128 output features
        mbar_wait(&bars[2], 0);
        tc_fence_after();
        if (threadIdx.x == 0) trace_mark(a.trace, 7);
        const int q = warp & 3, half = warp >> 2;
        float* tile = reinterpret_cast<float*>(wgu + 16384) + warp * (16 * 36);   // [16 tokens][32 features + 4 pad]
        float acc[16], acc1[16];
        tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + 128u + 16u * (uint32_t)half, acc);
        tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + 160u + 16u * (uint32_t)half, acc1);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] += acc1[j];
#pragma unroll
        for (int j = 0; j < 16; ++j) tile[j * 36 + lane] = acc[j];
        __syncwarp();
        const int f4 = (lane & 7) * 4;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int tl = 4 * r + (lane >> 3), tok = half * 16 + tl;
            if (tok < a.T) {
                const float4 v = *reinterpret_cast<const float4*>(tile + tl * 36 + f4);
                float* o = a.x_out + (long long)tok * a.H + c * 128 + q * 32 + f4;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc<256>(tmem_base);
    cluster_wait();
    if (threadIdx.x == 0) trace_end(a.trace);
}

}  // namespace ctp
