// Persistent fused decode-step kernel (sm_100a): ONE cooperative launch runs the whole decode step
//   code-embedding sum -> n_layers x { RMSNorm + QKV + RoPE + KV append | attention | o_proj + residual |
//   RMSNorm + gate/up + SiLU*mul | down_proj + residual } -> final norm -> heads -> fused sampler
// (reference: gpt.py:389-549 loop body; llama.py:689-749 LlamaDecoderLayer; processors.py).
//
// One CTA per SM, 6 warps:
//   warp 4 (one lane)  PRODUCER  walks a static per-CTA schedule of cp.async.bulk loads — pre-packed weight blobs and
//                      KV-cache tiles — into a ring of 16 KB shared-memory slots.  Weights and old KV do not depend on
//                      this step's activations, so the producer runs ahead across phase boundaries: HBM keeps
//                      streaming while the other warps sit in a grid barrier.
//   warps 0-3          per-CTA GEMM tiles D[32 batch rows x 16..48 weight rows] over the full K with warp-level
//                      mma.sync (m16n8k16, fp16 in / fp32 accumulate): at M=32 the 128-row tcgen05 atom costs ~128 cycles per
//                      instruction for 16-48 useful weight rows (measured, profiles/README.md), so the decode step uses HMMA
//                      and tcgen05 is kept for the M>=4096 GEMMs (prefill, vocoder).  Also: norm prologues, RoPE / SiLU /
//                      residual epilogues, decode attention over the KV tiles in the ring, the sampler; thread 0 runs
//                      the grid barrier between phases.
//
// Weight blobs are pre-packed at bind time into the exact shared-memory image the UMMA descriptor expects
// (K-major, 128-byte swizzle, one contiguous blob per work item), so a slot is filled by a single bulk copy.
#pragma once
#include "gpt_kernels.cuh"

namespace ctp {

constexpr int NCOMP = 256;                 // compute threads (8 warps); warp 8 is the producer
constexpr int STEP_THREADS = NCOMP + 32;
constexpr int SLOT_BYTES = 16384;
constexpr int MAX_RING = 12;
constexpr int A_KB_BYTES = 4096;      // one activation k-block: 32 rows x 128 B (the MMA reads 128 rows = 16 KB from it)
constexpr int A_OVERRUN = 0;
constexpr int DBUF_BYTES = 2 * 32 * 52 * 4;   // two K-half partial tiles [32][48 + 4 pad] fp32 between MMA fragments and epilogue
constexpr int DN_KSPLIT = 3;
constexpr int TMEM_COLS_STEP = 256;
// consecutive tcgen05.mma into ONE accumulator tile serialise on the accumulate dependency (~125+ cycles each, measured);
// the K loop therefore rotates over NCHAIN independent accumulators that the epilogue sums
__host__ __device__ constexpr int n_chains(int N) { return N <= 16 ? 8 : 4; }

struct StepParams {
    // dims
    int L, H, nH, I, num_vq, num_audio, B, max_seq;
    int b0;                // first global batch row of this lane (B = rows of the lane); KV / ids / logits use b0 + b
    float eps;
    int ring_slots;        // S
    int do_sample;         // 1: run the sampler phase (generate loop); 0: trunk + heads only (teacher-forced step)
    // packed weights (see pack_* kernels): one contiguous blob per work item
    const __half* wqkv_p;  // [L][3*nH*4 items][H/64][16][64]
    const __half* wo_p;    // [L][H/16 items][H/64][16][64]
    const __half* wgu_p;   // [L][I/24 items][H/64][48][64]
    const __half* wdn_p;   // [L][(H/16)*KS items][I/64/KS][16][64]
    const __half* whead_p; // [ceil(F/16) items][H/64][16][64]
    const float* ln1; const float* ln2; const float* norm_f;
    const __half* emb_code;
    // activations / state (global)
    float* x;              // [32][H] residual stream fp32
    float* q;              // [32][H] rotated queries fp32
    __half* attn_p;        // packed A operand of o_proj   [H/64][32][64] swizzled
    __half* h_p;           // packed A operand of down_proj [I/64][32][64] swizzled
    __half* kv;            // [L][2][maxB][nH][maxS][64]
    long long kv_plane;    // elements per plane
    float* logits;         // [B][num_vq*num_audio]
    float* hidden;         // [B][H]
    GenState* st;
    const int* pad_len;
    const float* inv_freq;
    const int* ids_ext;    // teacher-forced ids [B][num_vq] or null
    unsigned long long* bar;        // grid barrier counter (monotonic)
    unsigned long long* bar_epoch;  // number of barriers completed by all previous launches
    long long* dbg;                 // bring-up: [cta][barrier][2] clock64 at arrive / leave, or null
};

__device__ __forceinline__ uint64_t ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct StepCtx {
    uint8_t* ring;
    uint8_t* areg;           // activation operand region (after the ring)
    uint64_t* full;
    uint64_t* empty;
    int S;
    int cta, G;
    unsigned long long bar_base;  // counter value at launch
    int dbg_e;                    // bring-up: barrier id the current phase ends with
};

__device__ __forceinline__ uint64_t ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// poll with relaxed loads (no L1 invalidate per iteration), then one acquire fence
__device__ __forceinline__ void wait_counter(const unsigned long long* bar, unsigned long long target) {
    long long t0 = clock64();
    while (ld_relaxed_u64(bar) < target) {
        if (clock64() - t0 > 4000000000LL) {
            printf("ctp: grid barrier timeout (cta %d thread %d target %llu have %llu)\n", blockIdx.x, threadIdx.x, target,
                   ld_acquire_u64(bar));
            __trap();
        }
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

// grid barrier #e of this launch, executed by the 128 compute threads
__device__ __forceinline__ void grid_barrier(const StepParams& p, const StepCtx& c, int e) {
    compute_sync();
    if (threadIdx.x == 0) {
        long long* d = p.dbg ? p.dbg + ((size_t)c.cta * 128 + e) * 8 : nullptr;
        if (d) d[0] = clock64();
        fence_proxy_async_all();   // generic global writes of this phase -> visible to other CTAs' bulk copies
        asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p.bar), "l"(1ULL) : "memory");   // release: orders the CTA's writes
        wait_counter(p.bar, c.bar_base + (unsigned long long)(e + 1) * c.G);
        if (d) d[1] = clock64();
    }
    compute_sync();
}

// ---- ring helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t* slot_ptr(const StepCtx& c, uint32_t idx) { return c.ring + (size_t)(idx % c.S) * SLOT_BYTES; }
__device__ __forceinline__ void wait_full(const StepCtx& c, uint32_t idx) { mbar_wait(&c.full[idx % c.S], (idx / c.S) & 1); }
__device__ __forceinline__ void wait_empty(const StepCtx& c, uint32_t idx) { mbar_wait(&c.empty[idx % c.S], ((idx / c.S) & 1) ^ 1); }

// ---- work decomposition (identical in all three roles) --------------------------------------------------------------
struct GemmItem {
    const __half* w;      // packed weight blob of this item
    int N;                // weight rows (MMA N)
    int nkb;              // k-blocks
    int kb_per_wslot;     // k-blocks per weight slot
    int n_wslots;
    const __half* a_glob; // packed activation k-blocks in global (stream phases) or null (activation region in smem)
    int n_aslots;         // 4 k-blocks per activation slot
};

enum Phase { PH_QKV = 0, PH_ATT = 1, PH_O = 2, PH_GU = 3, PH_DN = 4, PH_HEAD = 5 };

__device__ __forceinline__ int phase_items(const StepParams& p, int ph) {
    switch (ph) {
        case PH_QKV: return 3 * p.nH * 4;
        case PH_O: return p.H / 16;
        case PH_GU: return p.I / 24;
        case PH_DN: return (p.H / 16) * DN_KSPLIT;
        case PH_HEAD: return (p.num_vq * p.num_audio + 15) / 16;
        default: return 0;
    }
}

__device__ __forceinline__ GemmItem make_item(const StepParams& p, int layer, int ph, int it) {
    GemmItem g;
    const int nkbH = p.H / 64;
    g.a_glob = nullptr;
    g.n_aslots = 0;
    switch (ph) {
        case PH_QKV:
            g.N = 16; g.nkb = nkbH;
            g.w = p.wqkv_p + ((size_t)layer * (3 * p.nH * 4) + it) * ((size_t)16 * p.H);
            break;
        case PH_O:
            g.N = 16; g.nkb = nkbH;
            g.w = p.wo_p + ((size_t)layer * (p.H / 16) + it) * ((size_t)16 * p.H);
            g.a_glob = p.attn_p;
            break;
        case PH_GU:
            g.N = 48; g.nkb = nkbH;
            g.w = p.wgu_p + ((size_t)layer * (p.I / 24) + it) * ((size_t)48 * p.H);
            break;
        case PH_DN: {
            const int nkbI = p.I / 64, per = nkbI / DN_KSPLIT;
            const int ks = it % DN_KSPLIT;
            g.N = 16; g.nkb = per;
            g.w = p.wdn_p + ((size_t)layer * (p.H / 16) * DN_KSPLIT + it) * ((size_t)16 * per * 64);
            g.a_glob = p.h_p + (size_t)ks * per * (A_KB_BYTES / 2);
            break;
        }
        default:  // PH_HEAD
            g.N = 16; g.nkb = nkbH;
            g.w = p.whead_p + (size_t)it * ((size_t)16 * p.H);
            break;
    }
    g.kb_per_wslot = SLOT_BYTES / (g.N * 128);
    g.n_wslots = (g.nkb + g.kb_per_wslot - 1) / g.kb_per_wslot;
    if (g.a_glob) g.n_aslots = (g.nkb + 3) / 4;
    return g;
}

// attention: unit u = b*nH + h is handled by CTA (u % G); old slots [pad_b, cur) stream through the ring in tiles of 64
__device__ __forceinline__ int att_tiles(int n_old) { return (n_old + 63) >> 6; }

// barrier ids within one launch
__device__ __forceinline__ int bar_id(int layer, int k) { return 1 + 5 * layer + k; }  // k: 0 after QKV .. 4 after DN

// =====================================================================================================================
// PRODUCER (warp 4, one lane)
// =====================================================================================================================
__device__ void step_producer(const StepParams& p, const StepCtx& c) {
    uint32_t idx = 0;
    const int cur = p.st->cur_len;
    auto load = [&](const void* src, uint32_t bytes) {
        wait_empty(c, idx);
        uint64_t* fb = &c.full[idx % c.S];
        mbar_expect_tx(fb, bytes);
        bulk_load_1d(slot_ptr(c, idx), src, bytes, fb);
        ++idx;
    };
    auto gemm_loads = [&](int layer, int ph, int need_bar) {
        const int n_it = phase_items(p, ph);
        for (int it = c.cta; it < n_it; it += c.G) {
            const GemmItem g = make_item(p, layer, ph, it);
            for (int ws = 0; ws < g.n_wslots; ++ws) {
                const int kb0 = ws * g.kb_per_wslot;
                const int nk = min(g.kb_per_wslot, g.nkb - kb0);
                load(g.w + (size_t)kb0 * g.N * 64, (uint32_t)(nk * g.N * 128));
            }
            if (g.a_glob) {
                // activations are produced by the previous phase: wait for its grid barrier, then order the async proxy
                wait_counter(p.bar, c.bar_base + (unsigned long long)(need_bar + 1) * c.G);
                fence_proxy_async_all();
                for (int as = 0; as < g.n_aslots; ++as) {
                    const int kb0 = as * 4;
                    const int nk = min(4, g.nkb - kb0);
                    load(g.a_glob + (size_t)kb0 * (A_KB_BYTES / 2), (uint32_t)(nk * A_KB_BYTES));
                }
            }
        }
    };
    for (int l = 0; l < p.L; ++l) {
        gemm_loads(l, PH_QKV, -1);
        {   // attention KV tiles (old slots only; the new token is read from global after the QKV barrier)
            const __half* kc = p.kv + (size_t)(2 * l) * p.kv_plane;
            const __half* vc = p.kv + (size_t)(2 * l + 1) * p.kv_plane;
            const int n_units = p.B * p.nH;
            for (int u = c.cta; u < n_units; u += c.G) {
                const int b = u / p.nH, h = u % p.nH;
                const int pad = p.pad_len[p.b0 + b];
                const int n_old = cur - pad;
                const size_t head_off = ((size_t)(p.b0 + b) * p.nH + h) * p.max_seq * HEAD_DIM;
                for (int t = 0; t < att_tiles(n_old); ++t) {
                    const int j0 = pad + t * 64;
                    const int n = min(64, cur - j0);
                    wait_empty(c, idx);
                    uint64_t* fb = &c.full[idx % c.S];
                    mbar_expect_tx(fb, (uint32_t)(2 * n * 128));
                    bulk_load_1d(slot_ptr(c, idx), kc + head_off + (size_t)j0 * HEAD_DIM, (uint32_t)(n * 128), fb);
                    bulk_load_1d(slot_ptr(c, idx) + 8192, vc + head_off + (size_t)j0 * HEAD_DIM, (uint32_t)(n * 128), fb);
                    ++idx;
                }
            }
        }
        gemm_loads(l, PH_O, bar_id(l, 1));
        gemm_loads(l, PH_GU, -1);
        gemm_loads(l, PH_DN, bar_id(l, 3));
    }
    gemm_loads(0, PH_HEAD, -1);
}

// number of attention tiles this CTA streams in one layer
__device__ __forceinline__ uint32_t att_slot_count(const StepParams& p, const StepCtx& c, int cur) {
    uint32_t n = 0;
    const int n_units = p.B * p.nH;
    for (int u = c.cta; u < n_units; u += c.G) n += att_tiles(cur - p.pad_len[p.b0 + u / p.nH]);
    return n;
}

// =====================================================================================================================
// COMPUTE warps 0-3
// =====================================================================================================================
template <int NT>
__device__ __forceinline__ void norm_rows_batched(const StepParams& p, const StepCtx& c, const float* w, float* out32, bool write_hid) {
    // warp handles rows warp, warp+8, ...; 4 rows per batch so that 4*NT*2 float4 loads are in flight per lane
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunk = p.H / 8;  // 16-byte fp16 chunks per row
    float4 wa[NT], wb[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const int ch = lane + 32 * t;
        if (ch < nchunk) {
            wa[t] = *reinterpret_cast<const float4*>(w + ch * 8);
            wb[t] = *reinterpret_cast<const float4*>(w + ch * 8 + 4);
        }
    }
    for (int r0 = warp; r0 < 32; r0 += 16) {
        float4 va[2][NT], vb[2][NT];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = (r0 + 8 * i + c.cta) & 31;   // rotate the row order per CTA: all SMs read x, spread them over the L2 slices
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int ch = lane + 32 * t;
                if (r < p.B && ch < nchunk) {
                    const float* xr = p.x + (size_t)r * p.H + ch * 8;
                    va[i][t] = __ldcg(reinterpret_cast<const float4*>(xr));       // x is updated by other CTAs inside
                    vb[i][t] = __ldcg(reinterpret_cast<const float4*>(xr + 4));   // this launch: bypass L1
                } else {
                    va[i][t] = make_float4(0.f, 0.f, 0.f, 0.f);
                    vb[i][t] = va[i][t];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = (r0 + 8 * i + c.cta) & 31;
            float ss = 0.f;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                ss += va[i][t].x * va[i][t].x + va[i][t].y * va[i][t].y + va[i][t].z * va[i][t].z + va[i][t].w * va[i][t].w;
                ss += vb[i][t].x * vb[i][t].x + vb[i][t].y * vb[i][t].y + vb[i][t].z * vb[i][t].z + vb[i][t].w * vb[i][t].w;
            }
            ss = warp_sum(ss);
            if (r >= p.B) continue;   // warp-uniform
            const float rstd = rsqrtf(ss / (float)p.H + p.eps);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int ch = lane + 32 * t;
                if (ch < nchunk) {
                    float y[8];
                    y[0] = wa[t].x * (va[i][t].x * rstd); y[1] = wa[t].y * (va[i][t].y * rstd);
                    y[2] = wa[t].z * (va[i][t].z * rstd); y[3] = wa[t].w * (va[i][t].w * rstd);
                    y[4] = wb[t].x * (vb[i][t].x * rstd); y[5] = wb[t].y * (vb[i][t].y * rstd);
                    y[6] = wb[t].z * (vb[i][t].z * rstd); y[7] = wb[t].w * (vb[i][t].w * rstd);
                    __half2 h0 = __floats2half2_rn(y[0], y[1]), h1 = __floats2half2_rn(y[2], y[3]);
                    __half2 h2 = __floats2half2_rn(y[4], y[5]), h3 = __floats2half2_rn(y[6], y[7]);
                    uint4 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                    const int kb = ch >> 3, cc = ch & 7;
                    *reinterpret_cast<uint4*>(c.areg + kb * A_KB_BYTES + r * 128 + ((cc ^ (r & 7)) << 4)) = pk;
                    if (out32) {
                        float* o = out32 + (size_t)(p.b0 + r) * p.H + ch * 8;
                        *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
                        *reinterpret_cast<float4*>(o + 4) = make_float4(y[4], y[5], y[6], y[7]);
                        if (write_hid && p.st->hid_buf && p.st->step < p.st->max_new) {
                            float* hb = p.st->hid_buf + ((size_t)(p.b0 + r) * p.st->max_new + p.st->step) * p.H + ch * 8;
                            *reinterpret_cast<float4*>(hb) = make_float4(y[0], y[1], y[2], y[3]);
                            *reinterpret_cast<float4*>(hb + 4) = make_float4(y[4], y[5], y[6], y[7]);
                        }
                    }
                }
            }
        }
    }
}

// RMSNorm of all B rows of x into the activation region (fp16, K-major 128B-swizzled k-blocks) — llama.py:82-87.
// out32 (CTA 0 in the heads phase): also emit the normalised rows in fp32 (hidden state) and into hid_buf.
__device__ void norm_prologue(const StepParams& p, const StepCtx& c, const float* w, float* out32, bool write_hid) {
    if (p.H <= 768) norm_rows_batched<3>(p, c, w, out32, write_hid);
    else norm_rows_batched<4>(p, c, w, out32, write_hid);
    compute_sync();
    if (threadIdx.x == 0 && p.dbg) p.dbg[((size_t)c.cta * 128 + c.dbg_e) * 8 + 2] = clock64();
}

// write 8 consecutive fp16 values (k0 multiple of 8) of batch row b into a packed activation buffer in global memory
__device__ __forceinline__ void store_packed8(__half* base, int b, int k0, const float* y) {
    __half2 h0 = __floats2half2_rn(y[0], y[1]), h1 = __floats2half2_rn(y[2], y[3]);
    __half2 h2 = __floats2half2_rn(y[4], y[5]), h3 = __floats2half2_rn(y[6], y[7]);
    uint4 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
    const int kb = k0 >> 6, cc = (k0 & 63) >> 3;
    uint8_t* dst = reinterpret_cast<uint8_t*>(base) + (size_t)kb * A_KB_BYTES + b * 128 + ((cc ^ (b & 7)) << 4);
    *reinterpret_cast<uint4*>(dst) = pk;
}

// ---- warp-level MMA helpers -------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t (&r)[2]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void hmma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 16-byte chunk `chunk` (0..7) of row `row` inside a 128B-swizzled K-major k-block that starts at `kb_base`
__device__ __forceinline__ uint32_t swz(uint32_t kb_base, int row, int chunk) {
    return kb_base + (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
}

// One GEMM work item: D[32 x N] = A[32 x K] . W[N x K]^T with K = 64*nkb, A and W as swizzled k-blocks in shared memory.
// 8 warps: warp w owns m-tile (w & 1), the n-tiles [NT*((w>>1)&1), +NT) with NT = N/16, and the k-blocks of parity (w >> 2).
// All fragments of a k-block are loaded before its MMAs (two accumulator sets break the accumulate dependency); the two
// K-half partial tiles are staged in `dbuf` ([2][32][N+4] fp32) and summed by the epilogue.
template <int N>
__device__ __forceinline__ void gemm_item_mma(const StepParams& p, const StepCtx& c, const GemmItem& g, uint32_t slot0, float* dbuf) {
    constexpr int NT = N / 16;           // n-tiles (of 8 rows) per warp: 1 (N=16) or 3 (N=48)
    constexpr int LD = N + 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = 16 * (warp & 1);
    const int n0 = 8 * NT * ((warp >> 1) & 1);
    const int khalf = warp >> 2;
    float acc[2][NT][4];
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int i = 0; i < NT; ++i) { acc[e][i][0] = acc[e][i][1] = acc[e][i][2] = acc[e][i][3] = 0.f; }
    const uint32_t w0 = slot0, a0 = slot0 + g.n_wslots;
    uint32_t w_seen = 0, a_seen = 0;
    const int a_row = m0 + (lane & 7) + ((lane >> 3) & 1) * 8, a_hi = lane >> 4;           // A: x4 = (m lo/hi) x (k lo/hi)
    const int b_row4 = (lane & 7) + ((lane >> 4) & 1) * 8, b_hi4 = (lane >> 3) & 1;          // B x4: (n-tile 0/1) x (k lo/hi)
    const int b_row2 = lane & 7, b_hi2 = (lane >> 3) & 1;                                    // B x2: one n-tile x (k lo/hi)
    for (int kb = khalf; kb < g.nkb; kb += 2) {
        constexpr int KPS = SLOT_BYTES / (N * 128);   // k-blocks per weight slot: 8 (N=16) or 2 (N=48)
        const uint32_t wneed = (uint32_t)(kb / KPS) + 1;
        while (w_seen < wneed) { wait_full(c, w0 + w_seen); ++w_seen; }
        uint32_t a_base;
        if (g.a_glob) {
            const uint32_t aneed = (uint32_t)(kb >> 2) + 1;
            while (a_seen < aneed) { wait_full(c, a0 + a_seen); ++a_seen; }
            a_base = smem_u32(slot_ptr(c, a0 + (kb >> 2))) + (kb & 3) * A_KB_BYTES;
        } else {
            a_base = smem_u32(c.areg) + kb * A_KB_BYTES;
        }
        const uint32_t b_base = smem_u32(slot_ptr(c, w0 + kb / KPS)) + (kb % KPS) * (N * 128);
        uint32_t a[4][4], bw[4][2 * NT];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            ldsm_x4(swz(a_base, a_row, 2 * ks + a_hi), a[ks]);
            if (NT == 1) {
                uint32_t t2[2];
                ldsm_x2(swz(b_base, n0 + b_row2, 2 * ks + b_hi2), t2);
                bw[ks][0] = t2[0]; bw[ks][1] = t2[1];
            } else {
                uint32_t t4[4], t2[2];
                ldsm_x4(swz(b_base, n0 + b_row4, 2 * ks + b_hi4), t4);
                ldsm_x2(swz(b_base, n0 + 16 + b_row2, 2 * ks + b_hi2), t2);
                bw[ks][0] = t4[0]; bw[ks][1] = t4[1]; bw[ks][2 % (2 * NT)] = t4[2]; bw[ks][3 % (2 * NT)] = t4[3];
                bw[ks][4 % (2 * NT)] = t2[0]; bw[ks][5 % (2 * NT)] = t2[1];
            }
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int i = 0; i < NT; ++i) hmma_16816(acc[ks & 1][i], a[ks], bw[ks][2 * i], bw[ks][2 * i + 1]);
    }
    // C fragment: c0,c1 -> row lane/4, cols 2*(lane%4)+{0,1}; c2,c3 -> row lane/4 + 8
    float* dh = dbuf + khalf * (32 * LD);
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        const int col = n0 + 8 * i + 2 * (lane & 3);
        const int r = m0 + (lane >> 2);
        *reinterpret_cast<float2*>(dh + r * LD + col) = make_float2(acc[0][i][0] + acc[1][i][0], acc[0][i][1] + acc[1][i][1]);
        *reinterpret_cast<float2*>(dh + (r + 8) * LD + col) = make_float2(acc[0][i][2] + acc[1][i][2], acc[0][i][3] + acc[1][i][3]);
    }
    compute_sync();   // both partial tiles complete; every warp is done reading this item's slots
    if (threadIdx.x == 0) {
        const uint32_t n_slots = g.n_wslots + g.n_aslots;
        for (uint32_t s2 = 0; s2 < n_slots; ++s2) mbar_arrive(&c.empty[(slot0 + s2) % c.S]);
    }
}

// 4 consecutive fp16 values (k0 multiple of 4) of batch row b into a packed activation buffer in global memory
__device__ __forceinline__ void store_packed4(__half* base, int b, int k0, const float* y) {
    __half2 h0 = __floats2half2_rn(y[0], y[1]), h1 = __floats2half2_rn(y[2], y[3]);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
    const int kb = k0 >> 6, cc = (k0 & 63) >> 3;
    uint8_t* dst = reinterpret_cast<uint8_t*>(base) + (size_t)kb * A_KB_BYTES + b * 128 + ((cc ^ (b & 7)) << 4) + (k0 & 7) * 2;
    *reinterpret_cast<uint2*>(dst) = pk;
}

// GEMM phase of this CTA: items it, it+G, ... ; 256-thread epilogues: thread = (batch row b = tid & 31, column group cg = tid >> 5)
__device__ void gemm_phase(const StepParams& p, const StepCtx& c, uint32_t& idx, int layer, int ph) {
    const int tid = threadIdx.x;
    const int b = tid & 31, cg = tid >> 5;   // cg in 0..7
    const bool live = b < p.B;
    const int n_it = phase_items(p, ph);
    const int cur = p.st->cur_len;
    float* dbuf = reinterpret_cast<float*>(c.areg + (p.H / 64) * A_KB_BYTES);
    for (int it = c.cta; it < n_it; it += c.G) {
        const GemmItem g = make_item(p, layer, ph, it);
        if (ph == PH_GU) gemm_item_mma<48>(p, c, g, idx, dbuf);
        else gemm_item_mma<16>(p, c, g, idx, dbuf);
        idx += g.n_wslots + g.n_aslots;
        if (p.dbg && tid == 0 && it == c.cta) p.dbg[((size_t)c.cta * 128 + c.dbg_e) * 8 + 3] = clock64();
        if (ph == PH_GU) {
            if (live && cg < 6) {   // h = silu(gate) * up (llama.py:214): gate cols [4cg, 4cg+4), up cols 24 + the same
                const float* d0 = dbuf + b * 52;
                const float* d1 = d0 + 32 * 52;
                float y[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float gte = d0[4 * cg + i] + d1[4 * cg + i];
                    const float up = d0[24 + 4 * cg + i] + d1[24 + 4 * cg + i];
                    y[i] = silu(gte) * up;
                }
                store_packed4(p.h_p, b, 24 * it + 4 * cg, y);
            }
        } else if (live) {
            const float* d0 = dbuf + b * 20;
            const float* d1 = d0 + 32 * 20;
            if (ph == PH_QKV) {
                const int type = it / (p.nH * 4), h = (it / 4) % p.nH, j = it % 4;
                if (type < 2) {   // RoPE on the pair (i, i+32), i = 8j + cg: llama.py:151-182; position = cur - pad (gpt.py:238-245)
                    const float pos = (float)(cur - p.pad_len[p.b0 + b]);
                    float sn, cs;
                    sincosf(pos * p.inv_freq[j * 8 + cg], &sn, &cs);
                    const float x1 = d0[cg] + d1[cg], x2 = d0[8 + cg] + d1[8 + cg];
                    const float lo = x1 * cs - x2 * sn, hi = x2 * cs + x1 * sn;
                    if (type == 0) {
                        float* qd = p.q + (size_t)b * p.H + h * 64 + j * 8 + cg;
                        qd[0] = lo;
                        qd[32] = hi;
                    } else {      // KV append, O(1) (replaces DynamicCache.update's torch.cat, llama.py:630-633)
                        __half* kd = p.kv + (size_t)(2 * layer) * p.kv_plane + (((size_t)(p.b0 + b) * p.nH + h) * p.max_seq + cur) * HEAD_DIM + j * 8 + cg;
                        kd[0] = __float2half_rn(lo);
                        kd[32] = __float2half_rn(hi);
                    }
                } else {
                    __half* vd = p.kv + (size_t)(2 * layer + 1) * p.kv_plane + (((size_t)(p.b0 + b) * p.nH + h) * p.max_seq + cur) * HEAD_DIM + j * 16 + 2 * cg;
                    *reinterpret_cast<__half2*>(vd) = __floats2half2_rn(d0[2 * cg] + d1[2 * cg], d0[2 * cg + 1] + d1[2 * cg + 1]);
                }
            } else {
                const float v0 = d0[2 * cg] + d1[2 * cg], v1 = d0[2 * cg + 1] + d1[2 * cg + 1];
                if (ph == PH_O) {          // residual add, this CTA owns features [16 it, 16 it + 16) (llama.py:737)
                    float* xd = p.x + (size_t)b * p.H + 16 * it + 2 * cg;
                    float2 o = __ldcg(reinterpret_cast<const float2*>(xd));
                    o.x += v0; o.y += v1;
                    *reinterpret_cast<float2*>(xd) = o;
                } else if (ph == PH_DN) {  // split-K partial of down_proj added into the residual (llama.py:745)
                    float* xd = p.x + (size_t)b * p.H + 16 * (it / DN_KSPLIT) + 2 * cg;
                    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(xd), "f"(v0), "f"(v1) : "memory");
                } else {                   // heads (gpt.py:424-439): logits row layout b*(num_vq*A) + q*A + a
                    const int F = p.num_vq * p.num_audio;
                    float* ld = p.logits + (size_t)(p.b0 + b) * F + 16 * it + 2 * cg;
                    if (16 * it + 2 * cg + 2 <= F) *reinterpret_cast<float2*>(ld) = make_float2(v0, v1);
                    else if (16 * it + 2 * cg < F) ld[0] = v0;
                }
            }
        }
        if (p.dbg && tid == 0) p.dbg[((size_t)c.cta * 128 + c.dbg_e) * 8 + 4] = clock64();
        compute_sync();   // dbuf is reused by the next item / phase
    }
}

// Decode attention (SDPA q_len = 1, llama.py:653-661) over the KV tiles in the ring.  Each warp owns whole units
// (b, h): the k-th unit of this CTA goes to warp k % 8, so units proceed independently (no CTA-wide syncs); inside a warp
// 4 groups of 8 lanes take positions round-robin, each lane holding 8 of the 64 head dims.
__device__ void attention_phase(const StepParams& p, const StepCtx& c, uint32_t& idx, int layer) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, sub = lane & 7;
    const int cur = p.st->cur_len;
    const __half* kc = p.kv + (size_t)(2 * layer) * p.kv_plane;
    const __half* vc = p.kv + (size_t)(2 * layer + 1) * p.kv_plane;
    const int n_units = p.B * p.nH;
    uint32_t base = idx;   // ring index of the first tile of the unit under consideration
    int k = 0;
    for (int u = c.cta; u < n_units; u += c.G, ++k) {
        const int b = u / p.nH, h = u % p.nH;
        const int pad = p.pad_len[p.b0 + b];
        const int n_old = cur - pad;
        const int n_tiles = att_tiles(n_old);
        if ((k & 7) == warp) {
            float q[8];
            {
                const float* qp = p.q + (size_t)b * p.H + h * 64 + sub * 8;
                const float4 a = __ldcg(reinterpret_cast<const float4*>(qp)), bq = __ldcg(reinterpret_cast<const float4*>(qp + 4));
                q[0] = a.x * 0.125f; q[1] = a.y * 0.125f; q[2] = a.z * 0.125f; q[3] = a.w * 0.125f;
                q[4] = bq.x * 0.125f; q[5] = bq.y * 0.125f; q[6] = bq.z * 0.125f; q[7] = bq.w * 0.125f;
            }
            // the new token (slot cur, appended by this step's QKV phase): issue its loads early, consumed by group 0 last
            const size_t off_new = (((size_t)(p.b0 + b) * p.nH + h) * p.max_seq + cur) * HEAD_DIM;
            uint4 k_new = make_uint4(0, 0, 0, 0), v_new = make_uint4(0, 0, 0, 0);
            if (grp == 0) {
                k_new = __ldcg(reinterpret_cast<const uint4*>(kc + off_new + sub * 8));
                v_new = __ldcg(reinterpret_cast<const uint4*>(vc + off_new + sub * 8));
            }
            float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = 0.f;
            auto dot8 = [&](const uint4& kr) {
                const __half2* k2 = reinterpret_cast<const __half2*>(&kr);
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(k2[i]);
                    s += q[2 * i] * f.x + q[2 * i + 1] * f.y;
                }
                return s;
            };
            // four positions per online-softmax rescale: the score chains are independent, one max/rescale per batch
            auto accum4 = [&](const uint4 (&kr)[4], const uint4 (&vr)[4], const bool (&valid)[4]) {
                float s[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) s[e] = dot8(kr[e]);
#pragma unroll
                for (int e = 0; e < 4; ++e) s[e] += __shfl_xor_sync(0xffffffffu, s[e], 1);
#pragma unroll
                for (int e = 0; e < 4; ++e) s[e] += __shfl_xor_sync(0xffffffffu, s[e], 2);
#pragma unroll
                for (int e = 0; e < 4; ++e) s[e] += __shfl_xor_sync(0xffffffffu, s[e], 4);
                float mn = m;
#pragma unroll
                for (int e = 0; e < 4; ++e) if (valid[e]) mn = fmaxf(mn, s[e]);
                if (mn == -INFINITY) return;   // nothing valid yet (uniform within the 8-lane group)
                const float corr = __expf(m - mn);
                float pr[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) pr[e] = valid[e] ? __expf(s[e] - mn) : 0.f;
                l = l * corr + (pr[0] + pr[1]) + (pr[2] + pr[3]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float a0 = o[2 * i] * corr, a1 = o[2 * i + 1] * corr;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = __half22float2(reinterpret_cast<const __half2*>(&vr[e])[i]);
                        a0 += pr[e] * f.x;
                        a1 += pr[e] * f.y;
                    }
                    o[2 * i] = a0;
                    o[2 * i + 1] = a1;
                }
                m = mn;
            };
            for (int t = 0; t < n_tiles; ++t) {
                const int n = min(64, n_old - t * 64);
                wait_full(c, base + t);
                const uint8_t* sl = slot_ptr(c, base + t);
#pragma unroll 2
                for (int r = 0; r < 4; ++r) {
                    uint4 kr[4], vr[4];
                    bool valid[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int j = grp + 4 * (4 * r + e);
                        valid[e] = j < n;
                        kr[e] = make_uint4(0, 0, 0, 0);
                        vr[e] = make_uint4(0, 0, 0, 0);
                        if (valid[e]) {
                            kr[e] = *reinterpret_cast<const uint4*>(sl + j * 128 + sub * 16);
                            vr[e] = *reinterpret_cast<const uint4*>(sl + 8192 + j * 128 + sub * 16);
                        }
                    }
                    accum4(kr, vr, valid);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&c.empty[(base + t) % c.S]);   // this warp was the tile's only reader
            }
            {   // the new token: group 0 only
                uint4 kr[4] = {k_new, make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
                uint4 vr[4] = {v_new, make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
                const bool valid[4] = {grp == 0, false, false, false};
                accum4(kr, vr, valid);
            }
            // merge the 4 position groups (lanes with equal `sub` hold the same 8 dims)
#pragma unroll
            for (int d = 8; d <= 16; d <<= 1) {
                const float m2 = __shfl_xor_sync(0xffffffffu, m, d);
                const float l2 = __shfl_xor_sync(0xffffffffu, l, d);
                const float mn = fmaxf(m, m2);
                const float w1 = (m == -INFINITY) ? 0.f : __expf(m - mn);
                const float w2 = (m2 == -INFINITY) ? 0.f : __expf(m2 - mn);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float o2 = __shfl_xor_sync(0xffffffffu, o[i], d);
                    o[i] = o[i] * w1 + o2 * w2;
                }
                l = l * w1 + l2 * w2;
                m = mn;
            }
            if (grp == 0) {   // packed A operand of o_proj: k-block h, row b, 16-byte chunk `sub`
                float y[8];
                const float inv = 1.0f / l;
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = o[i] * inv;
                store_packed8(p.attn_p, b, h * 64 + sub * 8, y);
            }
        }
        base += n_tiles;
    }
    idx = base;
}

__device__ void step_compute(const StepParams& p, StepCtx c) {
    uint32_t idx = 0;
    const int tid = threadIdx.x;
    const int cur = p.st->cur_len;
    // ---- code embedding (gpt.py:398-407): CTA b builds x[b]
    if (c.cta < p.B) {
        const int b = c.cta;
        __shared__ int sid[MAX_VQ];
        if (tid < p.num_vq)
            sid[tid] = p.ids_ext ? p.ids_ext[(p.b0 + b) * p.num_vq + tid]
                                 : p.st->ids_buf[((size_t)(p.b0 + b) * p.st->max_new + (p.st->step - 1)) * p.num_vq + tid];
        compute_sync();
        for (int k = tid; k < p.H; k += NCOMP) {
            float v = 0.f;
            for (int qv = 0; qv < p.num_vq; ++qv) v += __half2float(p.emb_code[((size_t)qv * p.num_audio + sid[qv]) * p.H + k]);
            p.x[(size_t)b * p.H + k] = v;
        }
    }
    grid_barrier(p, c, 0);
    auto count_att = [&]() { return att_slot_count(p, c, cur); };
    for (int l = 0; l < p.L; ++l) {
        // QKV
        c.dbg_e = bar_id(l, 0);
        if (c.cta < phase_items(p, PH_QKV)) norm_prologue(p, c, p.ln1 + (size_t)l * p.H, nullptr, false);
        gemm_phase(p, c, idx, l, PH_QKV);
        grid_barrier(p, c, bar_id(l, 0));
        // attention
        attention_phase(p, c, idx, l);
        grid_barrier(p, c, bar_id(l, 1));
        // o_proj
        c.dbg_e = bar_id(l, 2);
        gemm_phase(p, c, idx, l, PH_O);
        grid_barrier(p, c, bar_id(l, 2));
        // gate/up
        c.dbg_e = bar_id(l, 3);
        if (c.cta < phase_items(p, PH_GU)) norm_prologue(p, c, p.ln2 + (size_t)l * p.H, nullptr, false);
        gemm_phase(p, c, idx, l, PH_GU);
        grid_barrier(p, c, bar_id(l, 3));
        // down
        c.dbg_e = bar_id(l, 4);
        gemm_phase(p, c, idx, l, PH_DN);
        grid_barrier(p, c, bar_id(l, 4));
    }
    (void)count_att;
    c.dbg_e = 1 + 5 * p.L;
    // final norm (llama.py:1002) -> hidden state of this step (gpt.py:422-423) + heads
    if (c.cta < phase_items(p, PH_HEAD) || c.cta == 0) {
        if (c.cta < phase_items(p, PH_HEAD)) norm_prologue(p, c, p.norm_f, c.cta == 0 ? p.hidden : nullptr, true);
    }
    gemm_phase(p, c, idx, 0, PH_HEAD);
    const int last_bar = 1 + 5 * p.L;
    grid_barrier(p, c, last_bar);
    if (p.do_sample) {
        if (c.cta < p.B) {
            SampleArgs sa{};
            sa.logits = p.logits; sa.vocab = p.num_audio; sa.num_vq = p.num_vq; sa.rows = p.st->B * p.num_vq; sa.st = p.st;
            sa.b0 = p.b0; sa.ids_cols = p.num_vq;
            __shared__ int s_choice[MAX_VQ];
            sample_block<true>(sa, p.b0 + c.cta, p.B, reinterpret_cast<float*>(c.ring), s_choice);
        }
    } else if (c.cta == 0 && tid == 0) {
        p.st->cur_len = cur + 1;
    }
    if (c.cta == 0 && tid == 0) *p.bar_epoch = c.bar_base + (unsigned long long)(last_bar + 1) * c.G;
}

__global__ void __launch_bounds__(STEP_THREADS, 1) k_decode_step(const StepParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    StepCtx c;
    c.S = p.ring_slots;
    c.ring = smem;
    c.areg = smem + (size_t)c.S * SLOT_BYTES;
    const int areg_bytes = (p.H / 64) * A_KB_BYTES + DBUF_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(c.areg + areg_bytes);
    c.full = bars;
    c.empty = bars + MAX_RING;
    c.cta = blockIdx.x;
    c.G = gridDim.x;
    c.dbg_e = 0;
    c.bar_base = *p.bar_epoch;   // written only by the previous launch's last phase
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == NCOMP / 32 && lane == 0) {
        for (int s = 0; s < c.S; ++s) { mbar_init(&c.full[s], 1); mbar_init(&c.empty[s], 1); }
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();
    if (warp == NCOMP / 32) {
        if (lane == 0) step_producer(p, c);
    } else {
        step_compute(p, c);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Weight packing (bind time): natural [rows][K] fp16 -> per-item blobs [kb][N][64] with the 128-byte swizzle.
// row_of(item, n) gives the source row (or -1 for zero padding).
// ---------------------------------------------------------------------------------------------------------
enum PackKind { PACK_QKV = 0, PACK_PLAIN16 = 1, PACK_GU = 2, PACK_DN = 3 };

__global__ void k_pack_weights(const __half* __restrict__ src, __half* __restrict__ dst, int kind, int rows_total, int K, int nH, int I,
                               int n_items, int N, int nkb, int ksplit) {
    // one block per (item, kb); thread = (n, chunk)
    const int item = blockIdx.x / nkb, kb = blockIdx.x % nkb;
    for (int e = threadIdx.x; e < N * 8; e += blockDim.x) {
        const int n = e >> 3, cc = e & 7;
        int row = -1, k0 = kb * 64 + cc * 8;
        if (kind == PACK_QKV) {
            const int type = item / (nH * 4), h = (item / 4) % nH, j = item % 4;
            if (type < 2) row = type * (nH * 64) + h * 64 + j * 8 + (n < 8 ? n : 32 + (n - 8));
            else row = 2 * (nH * 64) + h * 64 + j * 16 + n;
        } else if (kind == PACK_PLAIN16) {
            row = item * 16 + n;
        } else if (kind == PACK_GU) {
            row = (n < 24) ? (24 * item + n) : (I + 24 * item + (n - 24));
        } else {  // PACK_DN: item = rs*ksplit + ks; k-blocks [ks*nkb, (ks+1)*nkb)
            row = (item / ksplit) * 16 + n;
            k0 += (item % ksplit) * nkb * 64;
        }
        uint4 val = make_uint4(0, 0, 0, 0);
        if (row >= 0 && row < rows_total) val = *reinterpret_cast<const uint4*>(src + (size_t)row * K + k0);
        uint8_t* d = reinterpret_cast<uint8_t*>(dst) + ((size_t)item * nkb + kb) * (N * 128) + n * 128 + ((cc ^ (n & 7)) << 4);
        *reinterpret_cast<uint4*>(d) = val;
    }
}

}  // namespace ctp
