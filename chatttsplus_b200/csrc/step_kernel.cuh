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
//   warp 5 (one lane)  MMA issuer: tcgen05.mma with the 32 batch rows as the (padded) M=128 operand and 16..48 weight
//                      rows as the N operand, full K per CTA -> no split-K atomics except the 512-element down_proj tile.
//   warps 0-3          norm prologues, RoPE / SiLU / residual epilogues (TMEM -> registers), decode attention over the
//                      KV tiles in the ring, the sampler; thread 0 runs the grid barrier between phases.
//
// Weight blobs are pre-packed at bind time into the exact shared-memory image the UMMA descriptor expects
// (K-major, 128-byte swizzle, one contiguous blob per work item), so a slot is filled by a single bulk copy.
#pragma once
#include "gpt_kernels.cuh"

namespace ctp {

constexpr int STEP_THREADS = 192;
constexpr int SLOT_BYTES = 16384;
constexpr int MAX_RING = 12;
constexpr int A_KB_BYTES = 4096;      // one activation k-block: 32 rows x 128 B (the MMA reads 128 rows = 16 KB from it)
constexpr int A_OVERRUN = 12288;
constexpr int DN_KSPLIT = 3;

struct StepParams {
    // dims
    int L, H, nH, I, num_vq, num_audio, B, max_seq;
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
};

__device__ __forceinline__ uint64_t ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct StepCtx {
    uint8_t* ring;
    uint8_t* areg;           // activation operand region (after the ring)
    uint64_t* full;
    uint64_t* empty;
    uint64_t* a_ready;       // compute warps -> MMA: activation operand written (norm phases)
    uint64_t* acc_ready;     // MMA -> epilogue
    uint64_t* acc_free;      // epilogue -> MMA
    uint32_t tmem;
    int S;
    int cta, G;
    unsigned long long bar_base;  // counter value at launch
};

__device__ __forceinline__ void wait_counter(const unsigned long long* bar, unsigned long long target) {
    long long t0 = clock64();
    while (ld_acquire_u64(bar) < target) {
        if (clock64() - t0 > 4000000000LL) {
            printf("ctp: grid barrier timeout (cta %d thread %d target %llu have %llu)\n", blockIdx.x, threadIdx.x, target,
                   ld_acquire_u64(bar));
            __trap();
        }
    }
}

// grid barrier #e of this launch, executed by the 128 compute threads
__device__ __forceinline__ void grid_barrier(const StepParams& p, const StepCtx& c, int e) {
    compute_sync();
    if (threadIdx.x == 0) {
        __threadfence();
        fence_proxy_async_all();   // generic global writes of this phase -> visible to other CTAs' bulk copies
        atomicAdd(p.bar, 1ULL);
        wait_counter(p.bar, c.bar_base + (unsigned long long)(e + 1) * c.G);
        __threadfence();
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
                const int pad = p.pad_len[b];
                const int n_old = cur - pad;
                const size_t head_off = ((size_t)b * p.nH + h) * p.max_seq * HEAD_DIM;
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

// =====================================================================================================================
// MMA issuer (warp 5, one lane)
// =====================================================================================================================
struct MmaState { uint32_t idx; uint32_t item; uint32_t a_phase; };

__device__ void mma_phase(const StepParams& p, const StepCtx& c, MmaState& m, int layer, int ph, bool norm_phase) {
    const int n_it = phase_items(p, ph);
    bool first = true;
    for (int it = c.cta; it < n_it; it += c.G) {
        const GemmItem g = make_item(p, layer, ph, it);
        if (norm_phase && first) {
            mbar_wait(c.a_ready, m.a_phase & 1);   // activation operand of this phase is in shared memory
            m.a_phase++;
        }
        first = false;
        if (m.item > 0) mbar_wait(c.acc_free, (m.item - 1) & 1);  // previous accumulator has been drained
        const uint32_t idesc = make_idesc_f16(128, g.N);
        const uint32_t w0 = m.idx, a0 = m.idx + g.n_wslots;
        const uint32_t n_slots = g.n_wslots + g.n_aslots;
        for (uint32_t s = 0; s < n_slots; ++s) wait_full(c, m.idx + s);
        tc_fence_after();
        for (int kb = 0; kb < g.nkb; ++kb) {
            uint32_t a_addr;
            if (g.a_glob) a_addr = smem_u32(slot_ptr(c, a0 + kb / 4)) + (kb % 4) * A_KB_BYTES;
            else a_addr = smem_u32(c.areg) + kb * A_KB_BYTES;
            const uint32_t b_addr = smem_u32(slot_ptr(c, w0 + kb / g.kb_per_wslot)) + (kb % g.kb_per_wslot) * (g.N * 128);
            const uint64_t da = make_kmajor_desc(a_addr, 1, 64, 2);
            const uint64_t db = make_kmajor_desc(b_addr, 1, 64, 2);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(c.tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        for (uint32_t s = 0; s < n_slots; ++s) umma_commit(&c.empty[(m.idx + s) % c.S]);
        umma_commit(c.acc_ready);
        m.idx += n_slots;
        m.item++;
    }
}

// the MMA warp only skips over the attention tiles in the slot sequence
__device__ __forceinline__ uint32_t att_slot_count(const StepParams& p, const StepCtx& c, int cur) {
    uint32_t n = 0;
    const int n_units = p.B * p.nH;
    for (int u = c.cta; u < n_units; u += c.G) n += att_tiles(cur - p.pad_len[u / p.nH]);
    return n;
}

__device__ void step_mma(const StepParams& p, const StepCtx& c) {
    MmaState m{0, 0, 0};
    const int cur = p.st->cur_len;
    for (int l = 0; l < p.L; ++l) {
        mma_phase(p, c, m, l, PH_QKV, true);
        m.idx += att_slot_count(p, c, cur);
        mma_phase(p, c, m, l, PH_O, false);
        mma_phase(p, c, m, l, PH_GU, true);
        mma_phase(p, c, m, l, PH_DN, false);
    }
    mma_phase(p, c, m, 0, PH_HEAD, true);
}

// =====================================================================================================================
// COMPUTE warps 0-3
// =====================================================================================================================
// RMSNorm of all B rows of x into the activation region (fp16, K-major 128B-swizzled k-blocks) — llama.py:82-87.
// out32 (CTA 0 in the heads phase): also emit the normalised rows in fp32 (hidden state) and into hid_buf.
__device__ void norm_prologue(const StepParams& p, const StepCtx& c, const float* w, float* out32, bool write_hid) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunk = p.H / 8;  // 16-byte chunks per row
    for (int r = warp; r < p.B; r += 4) {
        const float* xr = p.x + (size_t)r * p.H;
        float v[4][8];
        float ss = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int ch = lane + 32 * t;
            if (ch < nchunk) {
                const float4 a = __ldcg(reinterpret_cast<const float4*>(xr + ch * 8));       // x is updated by other CTAs
                const float4 b = __ldcg(reinterpret_cast<const float4*>(xr + ch * 8 + 4));   // inside this launch: skip L1
                v[t][0] = a.x; v[t][1] = a.y; v[t][2] = a.z; v[t][3] = a.w; v[t][4] = b.x; v[t][5] = b.y; v[t][6] = b.z; v[t][7] = b.w;
#pragma unroll
                for (int i = 0; i < 8; ++i) ss += v[t][i] * v[t][i];
            }
        }
        ss = warp_sum(ss);
        const float rstd = rsqrtf(ss / (float)p.H + p.eps);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int ch = lane + 32 * t;
            if (ch < nchunk) {
                const float4 wa = *reinterpret_cast<const float4*>(w + ch * 8);
                const float4 wb = *reinterpret_cast<const float4*>(w + ch * 8 + 4);
                const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                float y[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = ww[i] * (v[t][i] * rstd);
                __half2 h0 = __floats2half2_rn(y[0], y[1]), h1 = __floats2half2_rn(y[2], y[3]);
                __half2 h2 = __floats2half2_rn(y[4], y[5]), h3 = __floats2half2_rn(y[6], y[7]);
                uint4 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                const int kb = ch >> 3, cc = ch & 7;
                *reinterpret_cast<uint4*>(c.areg + kb * A_KB_BYTES + r * 128 + ((cc ^ (r & 7)) << 4)) = pk;
                if (out32) {
                    float* o = out32 + (size_t)r * p.H + ch * 8;
                    *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
                    *reinterpret_cast<float4*>(o + 4) = make_float4(y[4], y[5], y[6], y[7]);
                    if (write_hid && p.st->hid_buf && p.st->step < p.st->max_new) {
                        float* hb = p.st->hid_buf + ((size_t)r * p.st->max_new + p.st->step) * p.H + ch * 8;
                        *reinterpret_cast<float4*>(hb) = make_float4(y[0], y[1], y[2], y[3]);
                        *reinterpret_cast<float4*>(hb + 4) = make_float4(y[4], y[5], y[6], y[7]);
                    }
                }
            }
        }
    }
    fence_proxy_async();   // generic-proxy smem writes -> visible to tcgen05.mma (async proxy)
    compute_sync();
    if (threadIdx.x == 0) mbar_arrive(c.a_ready);
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

// read NC accumulator columns (NC in {16, 48}) of lane `lane` (= batch row) from TMEM
template <int NC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v) {
#pragma unroll
    for (int c0 = 0; c0 < NC; c0 += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr + (uint32_t)c0)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) v[c0 + i] = __uint_as_float(r[i]);
    }
}

// Epilogues: executed by warp 0 (TMEM lanes 0..31 = batch rows).  `item` counts this CTA's GEMM items (parity).
__device__ void gemm_epilogues(const StepParams& p, const StepCtx& c, uint32_t& item, int layer, int ph) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_it = phase_items(p, ph);
    const int cur = p.st->cur_len;
    for (int it = c.cta; it < n_it; it += c.G) {
        if (warp == 0) {
            mbar_wait(c.acc_ready, item & 1);
            tc_fence_after();
            const int b = lane;
            const bool live = b < p.B;
            if (ph == PH_GU) {
                float v[48];
                tmem_ld_cols<48>(c.tmem, v);
                if (live) {
                    float y[24];
#pragma unroll
                    for (int i = 0; i < 24; ++i) y[i] = silu(v[i]) * v[24 + i];   // llama.py:214
                    store_packed8(p.h_p, b, 24 * it, y);
                    store_packed8(p.h_p, b, 24 * it + 8, y + 8);
                    store_packed8(p.h_p, b, 24 * it + 16, y + 16);
                }
            } else {
                float v[16];
                tmem_ld_cols<16>(c.tmem, v);
                if (live) {
                    if (ph == PH_QKV) {
                        const int type = it / (p.nH * 4), h = (it / 4) % p.nH, j = it % 4;
                        if (type < 2) {   // RoPE on 8 (i, i+32) pairs: llama.py:151-182; position = cur - pad (gpt.py:238-245)
                            const float pos = (float)(cur - p.pad_len[b]);
                            float lo[8], hi[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                float sn, cs;
                                sincosf(pos * p.inv_freq[j * 8 + i], &sn, &cs);
                                lo[i] = v[i] * cs - v[8 + i] * sn;
                                hi[i] = v[8 + i] * cs + v[i] * sn;
                            }
                            if (type == 0) {
                                float* qd = p.q + (size_t)b * p.H + h * 64 + j * 8;
                                *reinterpret_cast<float4*>(qd) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                                *reinterpret_cast<float4*>(qd + 4) = make_float4(lo[4], lo[5], lo[6], lo[7]);
                                *reinterpret_cast<float4*>(qd + 32) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                                *reinterpret_cast<float4*>(qd + 36) = make_float4(hi[4], hi[5], hi[6], hi[7]);
                            } else {      // KV append, O(1) (replaces DynamicCache.update's torch.cat, llama.py:630-633)
                                __half* kd = p.kv + (size_t)(2 * layer) * p.kv_plane + (((size_t)b * p.nH + h) * p.max_seq + cur) * HEAD_DIM + j * 8;
                                __half2 t0 = __floats2half2_rn(lo[0], lo[1]), t1 = __floats2half2_rn(lo[2], lo[3]);
                                __half2 t2 = __floats2half2_rn(lo[4], lo[5]), t3 = __floats2half2_rn(lo[6], lo[7]);
                                uint4 pk;
                                pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                                pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
                                *reinterpret_cast<uint4*>(kd) = pk;
                                t0 = __floats2half2_rn(hi[0], hi[1]); t1 = __floats2half2_rn(hi[2], hi[3]);
                                t2 = __floats2half2_rn(hi[4], hi[5]); t3 = __floats2half2_rn(hi[6], hi[7]);
                                pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                                pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
                                *reinterpret_cast<uint4*>(kd + 32) = pk;
                            }
                        } else {
                            __half* vd = p.kv + (size_t)(2 * layer + 1) * p.kv_plane + (((size_t)b * p.nH + h) * p.max_seq + cur) * HEAD_DIM + j * 16;
                            __half2 t[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) t[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                            uint4 pk;
                            pk.x = *reinterpret_cast<uint32_t*>(&t[0]); pk.y = *reinterpret_cast<uint32_t*>(&t[1]);
                            pk.z = *reinterpret_cast<uint32_t*>(&t[2]); pk.w = *reinterpret_cast<uint32_t*>(&t[3]);
                            *reinterpret_cast<uint4*>(vd) = pk;
                            pk.x = *reinterpret_cast<uint32_t*>(&t[4]); pk.y = *reinterpret_cast<uint32_t*>(&t[5]);
                            pk.z = *reinterpret_cast<uint32_t*>(&t[6]); pk.w = *reinterpret_cast<uint32_t*>(&t[7]);
                            *reinterpret_cast<uint4*>(vd + 8) = pk;
                        }
                    } else if (ph == PH_O) {   // residual add, this CTA owns features [16 it, 16 it + 16) (llama.py:737)
                        float* xd = p.x + (size_t)b * p.H + 16 * it;
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            float4 o = __ldcg(reinterpret_cast<const float4*>(xd + i));
                            o.x += v[i]; o.y += v[i + 1]; o.z += v[i + 2]; o.w += v[i + 3];
                            *reinterpret_cast<float4*>(xd + i) = o;
                        }
                    } else if (ph == PH_DN) {  // split-K partial of down_proj added into the residual (llama.py:745)
                        float* xd = p.x + (size_t)b * p.H + 16 * (it / DN_KSPLIT);
#pragma unroll
                        for (int i = 0; i < 16; ++i) atomicAdd(xd + i, v[i]);
                    } else {                   // heads (gpt.py:424-439): logits row layout b*(num_vq*A) + q*A + a
                        const int F = p.num_vq * p.num_audio;
                        float* ld = p.logits + (size_t)b * F + 16 * it;
                        if (16 * it + 16 <= F) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(ld + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        } else {
                            for (int i = 0; i < 16; ++i)
                                if (16 * it + i < F) ld[i] = v[i];
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(c.acc_free);
        }
        item++;
    }
}

// Decode attention for this CTA's units over the KV tiles in the ring (SDPA q_len = 1, llama.py:653-661).
__device__ void attention_phase(const StepParams& p, const StepCtx& c, uint32_t& idx, int layer) {
    const int tid = threadIdx.x;   // 0..127
    const int grp = tid >> 3, sub = tid & 7;
    const int cur = p.st->cur_len;
    float* s_m = reinterpret_cast<float*>(c.areg);          // [16]
    float* s_l = s_m + 16;                                   // [16]
    float* s_o = s_l + 16;                                   // [16][65]
    const __half* kc = p.kv + (size_t)(2 * layer) * p.kv_plane;
    const __half* vc = p.kv + (size_t)(2 * layer + 1) * p.kv_plane;
    const int n_units = p.B * p.nH;
    for (int u = c.cta; u < n_units; u += c.G) {
        const int b = u / p.nH, h = u % p.nH;
        const int pad = p.pad_len[b];
        const int n_old = cur - pad;
        float q[8];
        {
            const float* qp = p.q + (size_t)b * p.H + h * 64 + sub * 8;
            const float4 a = __ldcg(reinterpret_cast<const float4*>(qp)), bq = __ldcg(reinterpret_cast<const float4*>(qp + 4));
            q[0] = a.x * 0.125f; q[1] = a.y * 0.125f; q[2] = a.z * 0.125f; q[3] = a.w * 0.125f;
            q[4] = bq.x * 0.125f; q[5] = bq.y * 0.125f; q[6] = bq.z * 0.125f; q[7] = bq.w * 0.125f;
        }
        float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = 0.f;
        auto accum = [&](const uint4& kr, const uint4& vr, bool valid) {
            const __half2* k2 = reinterpret_cast<const __half2*>(&kr);
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(k2[i]);
                s += q[2 * i] * f.x + q[2 * i + 1] * f.y;
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if (valid) {
                const float mn = fmaxf(m, s);
                const float corr = __expf(m - mn);
                const float pr = __expf(s - mn);
                l = l * corr + pr;
                const __half2* v2 = reinterpret_cast<const __half2*>(&vr);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(v2[i]);
                    o[2 * i] = o[2 * i] * corr + pr * f.x;
                    o[2 * i + 1] = o[2 * i + 1] * corr + pr * f.y;
                }
                m = mn;
            }
        };
        for (int t = 0; t < att_tiles(n_old); ++t) {
            const int n = min(64, n_old - t * 64);
            wait_full(c, idx);
            const uint8_t* sl = slot_ptr(c, idx);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int j = grp + 16 * r;
                const bool valid = j < n;
                uint4 kr = make_uint4(0, 0, 0, 0), vr = make_uint4(0, 0, 0, 0);
                if (valid) {
                    kr = *reinterpret_cast<const uint4*>(sl + j * 128 + sub * 16);
                    vr = *reinterpret_cast<const uint4*>(sl + 8192 + j * 128 + sub * 16);
                }
                accum(kr, vr, valid);
            }
            compute_sync();                       // all 128 threads are done with this slot
            if (tid == 0) mbar_arrive(&c.empty[idx % c.S]);
            ++idx;
        }
        {   // the new token (slot cur), appended by the QKV phase of this step: group 0 takes it from global memory
            const size_t off = (((size_t)b * p.nH + h) * p.max_seq + cur) * HEAD_DIM;
            const bool valid = grp == 0;
            uint4 kr = make_uint4(0, 0, 0, 0), vr = make_uint4(0, 0, 0, 0);
            if (valid) {
                kr = __ldcg(reinterpret_cast<const uint4*>(kc + off + sub * 8));
                vr = __ldcg(reinterpret_cast<const uint4*>(vc + off + sub * 8));
            }
            accum(kr, vr, valid);
        }
        if (sub == 0) { s_m[grp] = m; s_l[grp] = l; }
#pragma unroll
        for (int i = 0; i < 8; ++i) s_o[grp * 65 + sub * 8 + i] = o[i];
        compute_sync();
        if (tid < 64) {
            float M = -INFINITY;
#pragma unroll
            for (int g = 0; g < 16; ++g) M = fmaxf(M, s_m[g]);
            float Ls = 0.f, O = 0.f;
#pragma unroll
            for (int g = 0; g < 16; ++g) {
                const float w = (s_m[g] == -INFINITY) ? 0.f : __expf(s_m[g] - M);
                Ls += s_l[g] * w;
                O += s_o[g * 65 + tid] * w;
            }
            // packed A operand of o_proj: k-block h, row b, swizzled 16-byte chunks
            const int cc = tid >> 3;
            __half* dst = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(p.attn_p) + (size_t)h * A_KB_BYTES + b * 128 + ((cc ^ (b & 7)) << 4)) + (tid & 7);
            *dst = __float2half_rn(O / Ls);
        }
        compute_sync();
    }
}

__device__ void step_compute(const StepParams& p, const StepCtx& c) {
    uint32_t idx = 0, item = 0;
    const int tid = threadIdx.x;
    const int cur = p.st->cur_len;
    // ---- code embedding (gpt.py:398-407): CTA b builds x[b]
    if (c.cta < p.B) {
        const int b = c.cta;
        __shared__ int sid[MAX_VQ];
        if (tid < p.num_vq)
            sid[tid] = p.ids_ext ? p.ids_ext[b * p.num_vq + tid] : p.st->ids_buf[((size_t)b * p.st->max_new + (p.st->step - 1)) * p.num_vq + tid];
        compute_sync();
        for (int k = tid; k < p.H; k += 128) {
            float v = 0.f;
            for (int qv = 0; qv < p.num_vq; ++qv) v += __half2float(p.emb_code[((size_t)qv * p.num_audio + sid[qv]) * p.H + k]);
            p.x[(size_t)b * p.H + k] = v;
        }
    }
    grid_barrier(p, c, 0);
    auto count_att = [&]() { return att_slot_count(p, c, cur); };
    for (int l = 0; l < p.L; ++l) {
        // QKV
        if (c.cta < phase_items(p, PH_QKV)) norm_prologue(p, c, p.ln1 + (size_t)l * p.H, nullptr, false);
        gemm_epilogues(p, c, item, l, PH_QKV);
        { uint32_t n = 0; const int nq = phase_items(p, PH_QKV); for (int it = c.cta; it < nq; it += c.G) n += make_item(p, l, PH_QKV, it).n_wslots; idx += n; }
        grid_barrier(p, c, bar_id(l, 0));
        // attention
        attention_phase(p, c, idx, l);
        grid_barrier(p, c, bar_id(l, 1));
        // o_proj
        gemm_epilogues(p, c, item, l, PH_O);
        { uint32_t n = 0; const int nq = phase_items(p, PH_O); for (int it = c.cta; it < nq; it += c.G) { const GemmItem g = make_item(p, l, PH_O, it); n += g.n_wslots + g.n_aslots; } idx += n; }
        grid_barrier(p, c, bar_id(l, 2));
        // gate/up
        if (c.cta < phase_items(p, PH_GU)) norm_prologue(p, c, p.ln2 + (size_t)l * p.H, nullptr, false);
        gemm_epilogues(p, c, item, l, PH_GU);
        { uint32_t n = 0; const int nq = phase_items(p, PH_GU); for (int it = c.cta; it < nq; it += c.G) n += make_item(p, l, PH_GU, it).n_wslots; idx += n; }
        grid_barrier(p, c, bar_id(l, 3));
        // down
        gemm_epilogues(p, c, item, l, PH_DN);
        { uint32_t n = 0; const int nq = phase_items(p, PH_DN); for (int it = c.cta; it < nq; it += c.G) { const GemmItem g = make_item(p, l, PH_DN, it); n += g.n_wslots + g.n_aslots; } idx += n; }
        grid_barrier(p, c, bar_id(l, 4));
    }
    (void)count_att;
    // final norm (llama.py:1002) -> hidden state of this step (gpt.py:422-423) + heads
    if (c.cta < phase_items(p, PH_HEAD) || c.cta == 0) {
        if (c.cta < phase_items(p, PH_HEAD)) norm_prologue(p, c, p.norm_f, c.cta == 0 ? p.hidden : nullptr, true);
    }
    gemm_epilogues(p, c, item, 0, PH_HEAD);
    const int last_bar = 1 + 5 * p.L;
    grid_barrier(p, c, last_bar);
    if (p.do_sample) {
        if (c.cta < p.B) {
            SampleArgs sa{};
            sa.logits = p.logits; sa.vocab = p.num_audio; sa.num_vq = p.num_vq; sa.rows = p.B * p.num_vq; sa.st = p.st;
            __shared__ int s_choice[MAX_VQ];
            sample_block<true>(sa, c.cta, p.B, reinterpret_cast<float*>(c.ring), s_choice);
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
    const int areg_bytes = (p.H / 64) * A_KB_BYTES + A_OVERRUN;
    uint64_t* bars = reinterpret_cast<uint64_t*>(c.areg + areg_bytes);
    c.full = bars;
    c.empty = bars + MAX_RING;
    c.a_ready = bars + 2 * MAX_RING;
    c.acc_ready = c.a_ready + 1;
    c.acc_free = c.a_ready + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c.a_ready + 3);
    c.cta = blockIdx.x;
    c.G = gridDim.x;
    c.bar_base = *p.bar_epoch;   // written only by the previous launch's last phase
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 4 && lane == 0) {
        for (int s = 0; s < c.S; ++s) { mbar_init(&c.full[s], 1); mbar_init(&c.empty[s], 1); }
        mbar_init(c.a_ready, 1);
        mbar_init(c.acc_ready, 1);
        mbar_init(c.acc_free, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 5) tmem_alloc<64>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    c.tmem = *tmem_slot;
    if (warp == 4) {
        if (lane == 0) step_producer(p, c);
    } else if (warp == 5) {
        if (lane == 0) step_mma(p, c);
    } else {
        step_compute(p, c);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc<64>(c.tmem);
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
