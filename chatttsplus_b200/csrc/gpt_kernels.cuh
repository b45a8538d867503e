// Device kernels of the GPT decode / prefill path other than the GEMMs (sm_100a).
// Reference semantics cited per kernel (paths relative to the ChatTTSPlus checkout).
#pragma once
#include "ctp_common.cuh"
#include "../../include/ctp.h"

namespace ctp {

constexpr int HEAD_DIM = 64;
constexpr int MAX_VQ = 8;

// Generation state, resident in device memory so that a captured step graph can be replayed unchanged:
// every kernel reads the current cache length / step / buffer pointers from here.
struct GenState {
    int cur_len;     // KV slots filled (prompt + decoded); the next token is written at slot cur_len
    int step;        // number of sample steps taken so far
    int B;
    int max_new;
    int* ids_buf;            // [B][max_new][num_vq]
    float* hid_buf;          // [B][max_new][H] or null
    int* end_idx;            // [B]
    unsigned char* finish;   // [B]
    const float* u_base;     // [max_new][B*num_vq] or null
    int all_done;            // set by the sampler when every sequence has finished
    int ticket;              // scratch counter
    int text_mode;           // 1: refine-text pass (infer_text=True): text embedding in, head_text logits, one sampling column
    ctp_sample_cfg cfg;
};

struct GptDims {
    int L, H, nH, I, num_vq, num_audio, num_text, max_batch, max_seq;
    float eps;
};

// ---------------------------------------------------------------------------------------------------------
// Prompt embedding  (GPT.forward, gpt.py:125-149)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_embed_prompt(const int* __restrict__ ids, const unsigned char* __restrict__ text_mask,
                               const __half* __restrict__ emb_text, const __half* __restrict__ emb_code, float* __restrict__ out,
                               int H, int num_vq, int num_audio) {
    const int tok = blockIdx.x;
    const int* id = ids + (long long)tok * num_vq;
    const bool is_text = text_mask[tok] != 0;
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
        float v;
        if (is_text) {
            v = __half2float(emb_text[(long long)id[0] * H + c]);
        } else {
            v = 0.f;
            for (int q = 0; q < num_vq; ++q) v += __half2float(emb_code[((long long)q * num_audio + id[q]) * H + c]);
        }
        out[(long long)tok * H + c] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// RMSNorm (llama.py:82-87): xn = w * (x * rsqrt(mean(x^2) + eps)), fp32 math, fp16 operand for the next GEMM.
// Optional front end for the decode step: x = sum_q emb_code[q][ids[b][q]]  (gpt.py:398-407).
// Optional back end: also emit the normalised row in fp32 (final norm -> hidden state, gpt.py:422-423) and
// zero a scratch row (the fp32 split-K accumulator the following GEMM adds into).
// ---------------------------------------------------------------------------------------------------------
struct NormArgs {
    float* x;                 // [rows][H] residual stream (read; written when embedding)
    const float* w;           // [H]
    __half* xn;               // [rows][H]
    float* out_f32;           // [rows][H] or null
    float* zero_buf;          // [rows][zero_n] or null
    int zero_n;
    int H;
    float eps;
    // embedding front end (decode only)
    const GenState* st;       // null -> no embedding
    const int* ids_ext;       // [B][num_vq] or null (then ids_buf[b][step-1])
    const __half* emb_code;
    const __half* emb_text;   // text mode: x = emb_text[ids[b][0]]  (gpt.py:400-401)
    int num_vq, num_audio;
    int write_hid;            // 1: also copy out_f32 row into st->hid_buf[b][step]
    unsigned long long* trace;
};

__global__ void __launch_bounds__(256) k_rmsnorm(NormArgs a) {
    if (threadIdx.x == 0) trace_mark(a.trace, 0);
    pdl_launch_dependents();
    const int row = blockIdx.x;
    __shared__ float red[8];
    __shared__ int sid[MAX_VQ];
    float* x = a.x + (long long)row * a.H;
    const bool embed = (a.st != nullptr && a.emb_code != nullptr);
    float wv[4];   // the norm weight does not depend on the previous kernel: loaded before the wait (H <= 1024 with 256 threads)
    {
        int n = 0;
        for (int c = threadIdx.x; c < a.H; c += 256, ++n) wv[n] = __ldg(a.w + c);
    }
    pdl_wait();
    if (threadIdx.x == 0) trace_mark(a.trace, 1);
    if (embed && threadIdx.x < a.num_vq) {
        int id;
        if (a.ids_ext) id = a.ids_ext[row * a.num_vq + threadIdx.x];
        else id = a.st->ids_buf[((long long)row * a.st->max_new + (a.st->step - 1)) * a.num_vq + threadIdx.x];
        sid[threadIdx.x] = id;
    }
    if (embed) __syncthreads();
    float ss = 0.f;
    float vals[4];
    int n = 0;
    for (int c = threadIdx.x; c < a.H; c += 256, ++n) {
        float v;
        if (embed) {
            if (a.emb_text) {
                v = __half2float(a.emb_text[(long long)sid[0] * a.H + c]);
            } else {
                v = 0.f;
                for (int q = 0; q < a.num_vq; ++q) v += __half2float(a.emb_code[((long long)q * a.num_audio + sid[q]) * a.H + c]);
            }
            x[c] = v;
        } else {
            v = x[c];
        }
        vals[n] = v;
        ss += v * v;
    }
    ss = warp_sum(ss);
    if (threadIdx.x == 0) trace_mark(a.trace, 4);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i];
    const float rstd = rsqrtf(tot / (float)a.H + a.eps);
    n = 0;
    for (int c = threadIdx.x; c < a.H; c += 256, ++n) {
        const float y = wv[n] * (vals[n] * rstd);
        a.xn[(long long)row * a.H + c] = __float2half_rn(y);
        if (a.out_f32) a.out_f32[(long long)row * a.H + c] = y;
        if (a.write_hid && a.st->hid_buf && a.st->step < a.st->max_new)
            a.st->hid_buf[((long long)row * a.st->max_new + a.st->step) * a.H + c] = y;
    }
    if (threadIdx.x == 0) trace_mark(a.trace, 6);
    if (a.zero_buf) {
        float* z = a.zero_buf + (long long)row * a.zero_n;
        for (int c = threadIdx.x; c < a.zero_n; c += 256) z[c] = 0.f;
    }
    if (threadIdx.x == 0) trace_end(a.trace);
}

// RMSNorm of many rows (prefill: rows = B * L0): one warp per row, float4 loads, no block-level barrier.  H % 128 == 0, H <= 1024.
__global__ void __launch_bounds__(256) k_rmsnorm_rows(const float* __restrict__ x, const float* __restrict__ w, __half* __restrict__ xn,
                                                       long long rows, int H, float eps) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31, nf = H >> 7;
    const float4* xr = reinterpret_cast<const float4*>(x + row * H);
    float4 v[8];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (j < nf) {
            v[j] = xr[lane + 32 * j];
            ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
        }
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / (float)H + eps);
    uint2* o = reinterpret_cast<uint2*>(xn + row * H);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (j < nf) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * j);
            const __half2 h0 = __floats2half2_rn(g.x * (v[j].x * rstd), g.y * (v[j].y * rstd));
            const __half2 h1 = __floats2half2_rn(g.z * (v[j].z * rstd), g.w * (v[j].w * rstd));
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&h0);
            pk.y = *reinterpret_cast<const uint32_t*>(&h1);
            o[lane + 32 * j] = pk;
        }
    }
}

// h = silu(gate) * up   (llama.py:214), gate/up read from the fp32 accumulator [rows][2I]
// zero_after: the split-K accumulator is consumed exactly once per step, so its reader re-arms it for the next step
// (saves the 24 KB-per-row clear that used to sit in the RMSNorm kernel's critical path)
__global__ void k_silu_mul(float* __restrict__ gu, __half* __restrict__ out, int I, long long total, int zero_after) {
    pdl_launch_dependents();
    pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long r = i / I;
    const int c = (int)(i % I);
    const float g = gu[r * 2 * I + c];
    const float u = gu[r * 2 * I + I + c];
    out[i] = __float2half_rn(silu(g) * u);
    if (zero_after) { gu[r * 2 * I + c] = 0.f; gu[r * 2 * I + I + c] = 0.f; }
}

// ---------------------------------------------------------------------------------------------------------
// Decode attention: RoPE (llama.py:106-119,151-182) on the new q/k, KV append (replaces the torch.cat of
// DynamicCache.update, llama.py:630-633, with an O(1) in-place write), softmax(q K^T / 8) V over the cached slots
// of this sequence (SDPA q_len = 1, llama.py:653-661).  Left padding: slots [0, pad_len[b]) are masked.
// split partials are merged by the last CTA to arrive for each (b, h).
// ---------------------------------------------------------------------------------------------------------
struct AttnDecArgs {
    float* qkv;             // [B][3H] fp32 accumulators of the QKV GEMM (re-armed to zero by their last reader)
    __half* kcache;         // this layer: [maxB][nH][maxS][64]
    __half* vcache;
    __half* out;            // [B][H] fp16
    float* part;            // [B][nH][nsplit][66]
    int* counters;          // [B][nH]
    const int* pad_len;     // [B]
    const GenState* st;
    const float* inv_freq;  // [32]
    int H, nH, max_seq;
    // RMSNorm folded into the QKV GEMM (XNORM operand): the GEMM contracted x*w, the row factor rsqrt(sum(x^2)/H + eps)
    // (llama.py:85) is applied here; sum(x^2) per row arrives in ss
    const float* ss;        // [B] or null (rows already normalised)
    float eps;
    unsigned long long* trace;
};

// ---------------------------------------------------------------------------------------------------------
// Decode attention, TMA-staged: the cached K/V stream of this (b, head) is pulled
// into a 4-stage shared-memory ring by 1-D bulk copies (cp.async.bulk, 8 KB K + 8 KB V per 64-slot tile, one elected
// producer thread, mbarrier full/empty handshake) instead of per-thread 16-byte loads: few large requests keep HBM busy,
// and — because cached slots do not depend on this step's QKV GEMM — the first ring pass is issued BEFORE
// griddepcontrol.wait, so most of the KV stream overlaps the previous kernel (PDL).
//   * pre-wait loads use a possibly one-step-old cur_len as a hint and only touch whole tiles below it (slots written by
//     earlier steps / the prefill, complete long ago); the true length is read after the wait.
//   * grid (nH, B, nsplit), 160 threads: warps 0-3 consume (16 groups x 8 lanes, online softmax per group), warp 4 produces.
//   * KV splits are INTERLEAVED by tile (split sp owns tiles sp, sp + nsplit, ... counted from the first unmasked slot), so
//     which tiles a CTA owns does not depend on the current length and the pre-wait pass works for any nsplit.
// ---------------------------------------------------------------------------------------------------------
constexpr int AT_TILE = 64, AT_STAGES = 4, AT_THREADS = 160;
constexpr int AT_HALF_BYTES = AT_TILE * HEAD_DIM * 2;       // 8 KB: one K (or V) tile
constexpr int AT_STAGE_BYTES = 2 * AT_HALF_BYTES;
constexpr int AT_SMEM = AT_STAGES * AT_STAGE_BYTES + 128;

__global__ void __launch_bounds__(AT_THREADS) k_attn_decode_tma(AttnDecArgs a) {
    if (threadIdx.x == 0) trace_mark(a.trace, 0);
    pdl_launch_dependents();
    extern __shared__ uint8_t at_smem_raw[];
    uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 127) & ~uintptr_t(127));
    const int h = blockIdx.x, b = blockIdx.y, sp = blockIdx.z, nsplit = gridDim.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ __align__(8) uint64_t full_bar[AT_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[AT_STAGES];
    __shared__ float sq[HEAD_DIM];
    __shared__ __align__(16) __half sk_new[HEAD_DIM];
    __shared__ __align__(16) __half sv_new[HEAD_DIM];
    __shared__ __align__(16) float s_raw[3 * HEAD_DIM];   // this step's q | k | v accumulator rows of (b, head), landed by bulk copies
    __shared__ __align__(8) uint64_t raw_bar;
    __shared__ float s_m[16], s_l[16];
    __shared__ float s_o[16][HEAD_DIM + 1];
    __shared__ int s_last, s_pre;

    const int pad = a.pad_len[b];   // fixed for the whole generation (uploaded by the prefill call)
    const float my_inv_freq = a.inv_freq[tid & 31];   // constant table: fetched before the wait
    const long long head_off = (((long long)b * a.nH + h) * a.max_seq) * HEAD_DIM;
    __half* kc = a.kcache + head_off;
    __half* vc = a.vcache + head_off;
    if (tid == 0) {
        for (int s = 0; s < AT_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 4); }
        mbar_init(&raw_bar, 1);
        fence_barrier_init();
        fence_proxy_async();
        s_pre = 0;
    }
    __syncthreads();
    int pre = 0;
    if (warp == 4 && lane == 0) {
        const int hint = *reinterpret_cast<const volatile int*>(&a.st->cur_len);   // <= the true length (see header)
        const int n_full = (hint - pad) / AT_TILE;                                  // whole tiles below the hint
        const int mine = n_full > sp ? (n_full - sp + nsplit - 1) / nsplit : 0;     // ... that this split owns
        pre = mine > AT_STAGES ? AT_STAGES : mine;
        for (int i = 0; i < pre; ++i) {
            uint8_t* st = ring + i * AT_STAGE_BYTES;
            const long long p0 = pad + (long long)(sp + i * nsplit) * AT_TILE;
            mbar_expect_tx(&full_bar[i], AT_STAGE_BYTES);
            bulk_load_1d(st, kc + p0 * HEAD_DIM, AT_HALF_BYTES, &full_bar[i]);
            bulk_load_1d(st + AT_HALF_BYTES, vc + p0 * HEAD_DIM, AT_HALF_BYTES, &full_bar[i]);
        }
        s_pre = pre;
    }
    pdl_wait();
    if (tid == 0) trace_mark(a.trace, 1);
    if (warp == 4 && lane == 0) {
        // q | k | v of this step: three 256-byte rows of the fp32 accumulator the QKV GEMM just reduced into.  Lines produced by REDs
        // come back faster through the bulk-copy path than through per-thread loads (measured: 0.85 vs ~1.3 us)
        const float* src = a.qkv + (long long)b * 3 * a.H + h * HEAD_DIM;
        mbar_expect_tx(&raw_bar, 3 * HEAD_DIM * 4);
        bulk_load_1d(s_raw, src, HEAD_DIM * 4, &raw_bar);
        bulk_load_1d(s_raw + HEAD_DIM, src + a.H, HEAD_DIM * 4, &raw_bar);
        bulk_load_1d(s_raw + 2 * HEAD_DIM, src + 2 * a.H, HEAD_DIM * 4, &raw_bar);
    }
    const int cur = a.st->cur_len;  // new token's slot
    // cached slots [pad, cur) in tiles of 64 counted from pad; this split owns tiles sp, sp + nsplit, ...; the new token (slot cur)
    // is taken from shared memory by the last split
    const int nt_all = (cur - pad + AT_TILE - 1) / AT_TILE;
    int n_tiles = nt_all > sp ? (nt_all - sp + nsplit - 1) / nsplit : 0;

    if (warp == 4) {
        if (lane == 0) {
            for (int i = pre; i < n_tiles; ++i) {
                const int s = i % AT_STAGES;
                if (i >= AT_STAGES) mbar_wait(&empty_bar[s], ((i / AT_STAGES) & 1) ^ 1);
                const int p0 = pad + (sp + i * nsplit) * AT_TILE;
                const int cnt = min(AT_TILE, cur - p0);
                const uint32_t bytes = (uint32_t)cnt * HEAD_DIM * 2;
                uint8_t* st = ring + s * AT_STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], 2 * bytes);
                bulk_load_1d(st, kc + (long long)p0 * HEAD_DIM, bytes, &full_bar[s]);
                bulk_load_1d(st + AT_HALF_BYTES, vc + (long long)p0 * HEAD_DIM, bytes, &full_bar[s]);
            }
        }
        if (tid == 128) trace_end(a.trace);   // (producer exit; records only the max-exit stamp when it is the latest)
        return;
    }

    // ---- consumer warps (128 threads) ----
    float* qp = a.qkv + (long long)b * 3 * a.H + h * HEAD_DIM;
    // every load of the prologue is issued before the first use: one L2 round trip for q/k/v, the row factor and the length
    float rf = 1.f;   // deferred RMSNorm row factor (llama.py:85): the QKV GEMM contracted x*w, sum(x^2) arrives in ss
    if (a.ss) rf = rsqrtf(__ldcg(a.ss + b) / (float)a.H + a.eps);
    mbar_wait(&raw_bar, 0);
    float raw[4] = {0.f, 0.f, 0.f, 0.f};
    if (tid < 32) { raw[0] = s_raw[tid]; raw[1] = s_raw[tid + 32]; raw[2] = s_raw[HEAD_DIM + tid]; raw[3] = s_raw[HEAD_DIM + tid + 32]; }
    else if (tid < 96) raw[0] = s_raw[2 * HEAD_DIM + tid - 32];
    if (tid < 32) {
        const float pos = (float)(cur - pad);
        const float ang = pos * my_inv_freq;
        float sn, cs;
        sincosf(ang, &sn, &cs);
        const float q1 = raw[0] * rf, q2 = raw[1] * rf;
        const float k1 = raw[2] * rf, k2 = raw[3] * rf;
        sq[tid] = (q1 * cs - q2 * sn) * 0.125f;  // 1/sqrt(64) folded into q
        sq[tid + 32] = (q2 * cs + q1 * sn) * 0.125f;
        sk_new[tid] = __float2half_rn(k1 * cs - k2 * sn);
        sk_new[tid + 32] = __float2half_rn(k2 * cs + k1 * sn);
    } else if (tid < 96) {
        sv_new[tid - 32] = __float2half_rn(raw[0] * rf);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (tid == 0) trace_mark(a.trace, 4);
    if (sp == nsplit - 1 && tid < 16) {  // append (16 threads x 16 B = 64 halfs for K and for V)
        reinterpret_cast<uint2*>(kc + (long long)cur * HEAD_DIM)[tid] = reinterpret_cast<const uint2*>(sk_new)[tid];
        reinterpret_cast<uint2*>(vc + (long long)cur * HEAD_DIM)[tid] = reinterpret_cast<const uint2*>(sv_new)[tid];
    }
    const int pre_tiles = s_pre;                      // tiles whose bulk copies were issued before the wait
    if (n_tiles < pre_tiles) n_tiles = pre_tiles;     // (never leave a copy in flight; cannot happen within a generation)

    // 16 groups of 8 lanes; lane `sub` owns dims [8*sub, 8*sub+8)
    const int grp = tid >> 3, sub = tid & 7;
    float q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = sq[sub * 8 + i];
    float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;

    for (int it = 0; it < n_tiles; ++it) {
        const int s = it % AT_STAGES;
        mbar_wait(&full_bar[s], (it / AT_STAGES) & 1);
        const int cnt = min(AT_TILE, cur - (pad + (sp + it * nsplit) * AT_TILE));   // may be <= 0 only in the defensive case above
        const uint8_t* kt = ring + s * AT_STAGE_BYTES;
        const uint8_t* vt = kt + AT_HALF_BYTES;
        uint4 kr[4], vr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int p = grp + 16 * u;
            kr[u] = reinterpret_cast<const uint4*>(kt + p * 128)[sub];
            vr[u] = reinterpret_cast<const uint4*>(vt + p * 128)[sub];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);   // this warp's reads of the stage are in registers
        // four keys per group and tile: the four score chains (dot product + 8-lane butterfly) are independent, then ONE online-softmax
        // update for the four together (a per-key update serialises max -> exp -> rescale four times per tile)
        float sc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const __half2* k2 = reinterpret_cast<const __half2*>(&kr[u]);
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(k2[i]);
                acc += q[2 * i] * f.x + q[2 * i + 1] * f.y;
            }
            sc[u] = acc;
        }
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) {
#pragma unroll
            for (int u = 0; u < 4; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], off);
        }
        float mn = m;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if ((grp + 16 * u) >= cnt) sc[u] = -INFINITY;   // slots beyond the live range of a ragged last tile
            mn = fmaxf(mn, sc[u]);
        }
        if (mn > -INFINITY) {
            const float corr = __expf(m - mn);   // m = -inf on the first live tile: exp(-inf) = 0
            float p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) p[u] = __expf(sc[u] - mn);   // masked keys: exp(-inf) = 0
            l = l * corr + ((p[0] + p[1]) + (p[2] + p[3]));
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] *= corr;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if ((grp + 16 * u) < cnt) {   // (the bytes behind a ragged last tile are whatever the stage held before: never touch them)
                    const __half2* v2 = reinterpret_cast<const __half2*>(&vr[u]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 f = __half22float2(v2[i]);
                        o[2 * i] += p[u] * f.x;
                        o[2 * i + 1] += p[u] * f.y;
                    }
                }
            }
            m = mn;
        }
    }
    if (sp == nsplit - 1) {   // the new token (slot cur), group 0 takes it
        const __half2* k2 = reinterpret_cast<const __half2*>(sk_new) + sub * 4;
        float sc = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(k2[i]);
            sc += q[2 * i] * f.x + q[2 * i + 1] * f.y;
        }
        sc += __shfl_xor_sync(0xffffffffu, sc, 1);
        sc += __shfl_xor_sync(0xffffffffu, sc, 2);
        sc += __shfl_xor_sync(0xffffffffu, sc, 4);
        if (grp == 0) {
            const float mn = fmaxf(m, sc);
            const float corr = __expf(m - mn);
            const float p = __expf(sc - mn);
            l = l * corr + p;
            const __half2* v2 = reinterpret_cast<const __half2*>(sv_new) + sub * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(v2[i]);
                o[2 * i] = o[2 * i] * corr + p * f.x;
                o[2 * i + 1] = o[2 * i + 1] * corr + p * f.y;
            }
            m = mn;
        }
    }
    // merge the 16 groups
    if (tid == 0) trace_mark(a.trace, 5);
    if (sub == 0) { s_m[grp] = m; s_l[grp] = l; }
#pragma unroll
    for (int i = 0; i < 8; ++i) s_o[grp][sub * 8 + i] = o[i];
    asm volatile("bar.sync 1, 128;" ::: "memory");
    float M = -INFINITY, Lsum = 0.f, O = 0.f;
    if (tid < HEAD_DIM) {
#pragma unroll
        for (int g = 0; g < 16; ++g) M = fmaxf(M, s_m[g]);
#pragma unroll
        for (int g = 0; g < 16; ++g) {
            const float w = (s_m[g] == -INFINITY) ? 0.f : __expf(s_m[g] - M);
            Lsum += s_l[g] * w;
            O += s_o[g][tid] * w;
        }
    }
    if (nsplit == 1) {
        if (tid < HEAD_DIM) {
            a.out[(long long)b * a.H + h * HEAD_DIM + tid] = __float2half_rn(O / Lsum);
            { qp[tid] = 0.f; qp[a.H + tid] = 0.f; qp[2 * a.H + tid] = 0.f; }   // all reads of q,k,v happened before the first barrier
        }
        if (tid == 0) trace_end(a.trace);
        return;
    }
    float* pp = a.part + (((long long)b * a.nH + h) * nsplit + sp) * 66;
    if (tid < HEAD_DIM) pp[tid] = O;
    if (tid == 0) { pp[64] = M; pp[65] = Lsum; }
    __threadfence();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (tid == 0) {
        const int t = atomicAdd(&a.counters[b * a.nH + h], 1);
        s_last = (t == nsplit - 1);
        if (s_last) a.counters[b * a.nH + h] = 0;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (tid == 0) trace_end(a.trace);
    if (!s_last) return;
    __threadfence();
    if (tid < HEAD_DIM) {
        const float* p0 = a.part + (((long long)b * a.nH + h) * nsplit) * 66;
        float MM = -INFINITY;
        for (int s = 0; s < nsplit; ++s) MM = fmaxf(MM, p0[s * 66 + 64]);
        float LL = 0.f, OO = 0.f;
        for (int s = 0; s < nsplit; ++s) {
            const float ms = p0[s * 66 + 64];
            const float w = (ms == -INFINITY) ? 0.f : __expf(ms - MM);
            LL += p0[s * 66 + 65] * w;
            OO += p0[s * 66 + tid] * w;
        }
        a.out[(long long)b * a.H + h * HEAD_DIM + tid] = __float2half_rn(OO / LL);
        { qp[tid] = 0.f; qp[a.H + tid] = 0.f; qp[2 * a.H + tid] = 0.f; }   // every split has arrived (ticket): safe to re-arm
    }
}

// ---------------------------------------------------------------------------------------------------------
// Prefill: RoPE + KV write for all prompt tokens, then causal attention with left-padding mask.
// ---------------------------------------------------------------------------------------------------------
// grid (L0, B), block = nH * 32 threads: thread (h, i) rotates pair (i, i+32) of head h.
__global__ void k_rope_prefill(float* __restrict__ qkv, __half* __restrict__ kcache, __half* __restrict__ vcache,
                               const int* __restrict__ pad_len, const float* __restrict__ inv_freq, int H, int nH, int L0,
                               int max_seq) {
    const int j = blockIdx.x, b = blockIdx.y;
    const int h = threadIdx.x >> 5, i = threadIdx.x & 31;
    const int pad = pad_len[b];
    const float pos = (j >= pad) ? (float)(j - pad) : 1.0f;  // gpt.py:238-245: padded slots get position 1
    float sn, cs;
    sincosf(pos * inv_freq[i], &sn, &cs);
    float* row = qkv + ((long long)b * L0 + j) * 3 * H;
    const float q1 = row[h * 64 + i], q2 = row[h * 64 + i + 32];
    row[h * 64 + i] = q1 * cs - q2 * sn;
    row[h * 64 + i + 32] = q2 * cs + q1 * sn;
    const float k1 = row[H + h * 64 + i], k2 = row[H + h * 64 + i + 32];
    const long long off = ((((long long)b * nH + h) * max_seq) + j) * HEAD_DIM;
    kcache[off + i] = __float2half_rn(k1 * cs - k2 * sn);
    kcache[off + i + 32] = __float2half_rn(k2 * cs + k1 * sn);
    vcache[off + i] = __float2half_rn(row[2 * H + h * 64 + i]);
    vcache[off + i + 32] = __float2half_rn(row[2 * H + h * 64 + i + 32]);
}

// grid (nH, B, ceil(L0/8)), 256 threads = 8 warps, one query row per warp; keys strided over lanes.
// K/V for the (b,h) pair are read from the cache (fp16, L2-resident for prompt-sized L0).
__global__ void __launch_bounds__(256) k_attn_prefill(const float* __restrict__ qkv, const __half* __restrict__ kcache,
                                                       const __half* __restrict__ vcache, __half* __restrict__ out,
                                                       const int* __restrict__ pad_len, int H, int nH, int L0, int max_seq) {
    const int h = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.z * 8 + warp;  // query slot
    __shared__ float sq[8][HEAD_DIM];
    __shared__ float sp[8][32];
    const int pad = pad_len[b];
    if (i < L0) {
        const float* qr = qkv + ((long long)b * L0 + i) * 3 * H + h * HEAD_DIM;
        sq[warp][lane] = qr[lane] * 0.125f;
        sq[warp][lane + 32] = qr[lane + 32] * 0.125f;
    }
    __syncwarp();
    if (i >= L0) return;
    const long long head_off = (((long long)b * nH + h) * max_seq) * HEAD_DIM;
    const __half* kc = kcache + head_off;
    const __half* vc = vcache + head_off;
    // padded query rows (i < pad) attend to themselves only: their output is never consumed (left padding)
    const int jlo = (i >= pad) ? pad : i;
    const int jhi = i;  // inclusive
    float m = -INFINITY, l = 0.f;
    float o0 = 0.f, o1 = 0.f;  // this lane owns output dims lane and lane+32
    for (int jb = jlo; jb <= jhi; jb += 32) {
        const int j = jb + lane;
        float s = -INFINITY;
        if (j <= jhi) {
            const uint4* kr = reinterpret_cast<const uint4*>(kc + (long long)j * HEAD_DIM);
            s = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 kk = __ldg(kr + c);
                const __half2* k2 = reinterpret_cast<const __half2*>(&kk);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float2 f = __half22float2(k2[t]);
                    s += sq[warp][c * 8 + 2 * t] * f.x + sq[warp][c * 8 + 2 * t + 1] * f.y;
                }
            }
        }
        const float mb = warp_max(s);
        const float mn = fmaxf(m, mb);
        const float corr = __expf(m - mn);
        const float p = (j <= jhi) ? __expf(s - mn) : 0.f;
        l = l * corr + warp_sum(p);
        o0 *= corr;
        o1 *= corr;
        sp[warp][lane] = p;
        __syncwarp();
        const int cnt = min(32, jhi - jb + 1);
        for (int t = 0; t < cnt; ++t) {
            const float pt = sp[warp][t];
            const __half* vr = vc + (long long)(jb + t) * HEAD_DIM;
            o0 += pt * __half2float(vr[lane]);
            o1 += pt * __half2float(vr[lane + 32]);
        }
        __syncwarp();
        m = mn;
    }
    __half* orow = out + ((long long)b * L0 + i) * H + h * HEAD_DIM;
    orow[lane] = __float2half_rn(o0 / l);
    orow[lane + 32] = __float2half_rn(o1 / l);
}

// ---------------------------------------------------------------------------------------------------------
// Prefill attention on the tensor cores (default): flash-style causal attention with left-padding mask, one CTA per
// (64-query tile, head, sequence), 4 warps x 16 query rows.  S = Q K^T and O += P V run as mma.sync.m16n8k16 (fp16 in, fp32
// accumulate: a 64x64x64 tile pair per step is far too small for tcgen05, whose minimum M is 64 per CTA-wide instruction and
// whose operands must round-trip through shared memory descriptors); K is staged row-major and V transposed in shared memory
// so that every B fragment is one conflict-free 32-bit load; the score fragments are reused in place as the A operand of P V.
// Same semantics as k_attn_prefill: query slot i sees key slots [pad, i]; padded query rows (i < pad) see only themselves
// (their output is never consumed).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int PF_TILE = 64, PF_PITCH = 72;   // 64-slot tiles; shared-memory row pitch in halfs (144 B: conflict-free fragment loads)

__global__ void __launch_bounds__(128) k_attn_prefill_mma(const float* __restrict__ qkv, const __half* __restrict__ kcache,
                                                           const __half* __restrict__ vcache, __half* __restrict__ out,
                                                           const int* __restrict__ pad_len, int H, int nH, int L0, int max_seq) {
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    __shared__ __align__(16) __half sK[PF_TILE * PF_PITCH];    // [key][d]
    __shared__ __align__(16) __half sVt[HEAD_DIM * PF_PITCH];  // [d][key]
    const int pad = pad_len[b];
    const long long head_off = (((long long)b * nH + h) * max_seq) * HEAD_DIM;
    const __half* kc = kcache + head_off;
    const __half* vc = vcache + head_off;
    // this thread's two query rows
    const int i0 = qt * PF_TILE + warp * 16 + g, i1 = i0 + 8;
    // Q fragments (A operand, 16 x 64 per warp = 4 k-steps), 1/sqrt(64) folded in; rows beyond L0 read as zero
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const int d0 = ks * 16 + 2 * t;
        const float* r0 = qkv + ((long long)b * L0 + i0) * 3 * H + h * HEAD_DIM;
        const float* r1 = qkv + ((long long)b * L0 + i1) * 3 * H + h * HEAD_DIM;
        const bool ok0 = i0 < L0, ok1 = i1 < L0;
        qa[ks][0] = ok0 ? pack_h2(r0[d0] * 0.125f, r0[d0 + 1] * 0.125f) : 0u;
        qa[ks][1] = ok1 ? pack_h2(r1[d0] * 0.125f, r1[d0 + 1] * 0.125f) : 0u;
        qa[ks][2] = ok0 ? pack_h2(r0[d0 + 8] * 0.125f, r0[d0 + 9] * 0.125f) : 0u;
        qa[ks][3] = ok1 ? pack_h2(r1[d0 + 8] * 0.125f, r1[d0 + 9] * 0.125f) : 0u;
    }
    const int lo0 = i0 >= pad ? pad : i0, lo1 = i1 >= pad ? pad : i1;   // first visible key slot of each row
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }

    const int q_hi = min(qt * PF_TILE + PF_TILE - 1, L0 - 1);           // last query slot of the tile
    const int q_lo = qt * PF_TILE;
    const int k_first = (min(q_lo, pad) / PF_TILE);                      // rows >= pad start at pad; padded rows at themselves (>= q_lo)
    for (int kt = k_first; kt * PF_TILE <= q_hi; ++kt) {
        const int j0 = kt * PF_TILE;
        __syncthreads();   // previous tile fully consumed
        // stage K [key][d] and V^T [d][key]: 64 keys x 8 chunks of 8 halfs; keys beyond the prompt read as zero (masked below)
        for (int c = tid; c < PF_TILE * 8; c += 128) {
            const int key = c >> 3, ch = c & 7;
            uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
            if (j0 + key < L0) {
                kv = __ldg(reinterpret_cast<const uint4*>(kc + (long long)(j0 + key) * HEAD_DIM) + ch);
                vv = __ldg(reinterpret_cast<const uint4*>(vc + (long long)(j0 + key) * HEAD_DIM) + ch);
            }
            *reinterpret_cast<uint4*>(sK + key * PF_PITCH + ch * 8) = kv;
            const __half* vh = reinterpret_cast<const __half*>(&vv);
#pragma unroll
            for (int e = 0; e < 8; ++e) sVt[(ch * 8 + e) * PF_PITCH + key] = vh[e];
        }
        __syncthreads();
        // S = Q K^T for 16 rows x 64 keys
        float sc[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const __half* kp = sK + (n * 8 + g) * PF_PITCH + ks * 16 + 2 * t;
                mma_16816(sc[n], qa[ks], *reinterpret_cast<const uint32_t*>(kp), *reinterpret_cast<const uint32_t*>(kp + 8));
            }
        }
        // mask + online softmax (rows i0 and i1; this thread holds keys j0 + 8n + 2t, +1)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = j0 + n * 8 + 2 * t + e;
                if (j > i0 || j < lo0) sc[n][e] = -INFINITY;
                if (j > i1 || j < lo1) sc[n][2 + e] = -INFINITY;
                mx0 = fmaxf(mx0, sc[n][e]);
                mx1 = fmaxf(mx1, sc[n][2 + e]);
            }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        // rows with no visible key in this tile (and none before) keep m = -inf: use 0 as the reference to avoid inf - inf
        const float ref0 = mn0 == -INFINITY ? 0.f : mn0, ref1 = mn1 == -INFINITY ? 0.f : mn1;
        const float corr0 = __expf(m0 - ref0), corr1 = __expf(m1 - ref1);   // exp(-inf) = 0 on the first visible tile
        float ps0 = 0.f, ps1 = 0.f;
        uint32_t pa[4][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float p00 = __expf(sc[n][0] - ref0), p01 = __expf(sc[n][1] - ref0);
            const float p10 = __expf(sc[n][2] - ref1), p11 = __expf(sc[n][3] - ref1);
            ps0 += p00 + p01; ps1 += p10 + p11;
            // score fragment of key n-tile n -> A fragment of key k-step n/2 (low half: n even, high half: n odd)
            pa[n >> 1][(n & 1) * 2 + 0] = pack_h2(p00, p01);
            pa[n >> 1][(n & 1) * 2 + 1] = pack_h2(p10, p11);
        }
        ps0 += __shfl_xor_sync(0xffffffffu, ps0, 1); ps0 += __shfl_xor_sync(0xffffffffu, ps0, 2);
        ps1 += __shfl_xor_sync(0xffffffffu, ps1, 1); ps1 += __shfl_xor_sync(0xffffffffu, ps1, 2);
        l0 = l0 * corr0 + ps0; l1 = l1 * corr1 + ps1;
        m0 = mn0; m1 = mn1;
        // O = O * corr + P V
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            o[n][0] *= corr0; o[n][1] *= corr0; o[n][2] *= corr1; o[n][3] *= corr1;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const __half* vp = sVt + (n * 8 + g) * PF_PITCH + ks * 16 + 2 * t;
                mma_16816(o[n], pa[ks], *reinterpret_cast<const uint32_t*>(vp), *reinterpret_cast<const uint32_t*>(vp + 8));
            }
        }
    }
    // normalise and store (fp16 operand rows of the o_proj GEMM)
    const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        const int d = n * 8 + 2 * t;
        if (i0 < L0) *reinterpret_cast<uint32_t*>(out + ((long long)b * L0 + i0) * H + h * HEAD_DIM + d) = pack_h2(o[n][0] * inv0, o[n][1] * inv0);
        if (i1 < L0) *reinterpret_cast<uint32_t*>(out + ((long long)b * L0 + i1) * H + h * HEAD_DIM + d) = pack_h2(o[n][2] * inv1, o[n][3] * inv1);
    }
}

// x_last[b] = x[b][L0-1]  (only the last prompt position feeds the heads, gpt.py:442)
__global__ void k_gather_last(const float* __restrict__ x, float* __restrict__ out, int L0, int H) {
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < H; c += blockDim.x) out[(long long)b * H + c] = x[((long long)b * L0 + L0 - 1) * H + c];
}

// ---------------------------------------------------------------------------------------------------------
// Sampler: one warp per (b, q) logits row.  gpt.py:469-494 + processors.py:18-34 + TopP/TopK warpers.
//
// TopP (ascending cumulative prob <= 1-p removed, last min_keep kept) followed by TopK(k) leaves exactly the
// highest-scoring tokens, in rank order, while the probability mass strictly above a token is < top_p, capped at
// k tokens and never fewer than min_keep: so the 20 best scores plus the full-row softmax normaliser suffice and no
// 626-wide sort is needed.  (Exact ties at the k-th score are broken by token id instead of all being kept.)
// ---------------------------------------------------------------------------------------------------------
constexpr int SAMPLE_MAX_K = 32;

struct SampleArgs {
    const float* logits;   // [rows][vocab]
    int vocab, num_vq, rows;   // num_vq = sampling columns per batch row (1 in text mode)
    int ids_cols;              // columns of the ids / history buffers per token (the model's num_vq)
    // history for the repetition penalty: row r looks at hist[r_b][t][r_q], t in [hist_len - window, hist_len)
    const int* hist;       // element (b, t, q) at hist[(b*hist_stride + t)*num_vq + q]
    int hist_stride;       // tokens per batch row in the history buffer
    int hist_len;          // valid history length (ignored when st != null: then st->step)
    const float* u;        // [rows] or null
    int step;              // (ignored when st != null)
    ctp_sample_cfg cfg;    // (ignored when st != null: then st->cfg)
    int* next_ids;         // [rows] or null
    float* probs_out;      // [rows][vocab] or null
    GenState* st;          // generation mode: write ids_buf / finish / end_idx
    int advance_len;       // graph path: the sampler's last block also advances cur_len (one kernel less per step)
    unsigned long long* trace;
    int b0;                // lanes: first global row handled by this launch (ticket / all_done cover rows [b0, b0 + n_blocks))
};

__device__ __forceinline__ uint32_t philox_mix(uint64_t seed, uint32_t a, uint32_t b) {
    // Philox-2x32-10 style counter hash (counter = (a, b), key from seed) -> 32 random bits
    uint32_t k = (uint32_t)(seed ^ (seed >> 32));
    uint32_t c0 = a, c1 = b;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint64_t p = (uint64_t)0xD256D193u * c0;
        const uint32_t hi = (uint32_t)(p >> 32), lo = (uint32_t)p;
        c0 = hi ^ k ^ c1;
        c1 = lo;
        k += 0x9E3779B9u;
    }
    return c0;
}

// One warp per (b, q) row: warps 0..num_vq-1 of the calling block handle rows b*num_vq + warp.
// FUSED: called from the persistent step kernel (compute warps only: named barrier 1 over 256 threads, n_blocks = B).
// PDL_INSIDE: the function itself executes griddepcontrol.wait, after everything that does not depend on the logits (sampling
// configuration, history window and repetition counts, the uniform) has been loaded — that part overlaps the heads GEMM.
template <bool FUSED, bool PDL_INSIDE = false>
__device__ __forceinline__ void sample_block(const SampleArgs& a, const int b, const int n_blocks, float* s_scores, int* s_choice) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= a.num_vq) {
        if (PDL_INSIDE) pdl_wait();
        if (a.st) { if (FUSED) asm volatile("bar.sync 1, 256;" ::: "memory"); else __syncthreads(); }
        return;
    }
    const int row = b * a.num_vq + warp;
    const int V = a.vocab;
    const int vpad = (V + 31) & ~31;
    float* sc = s_scores + warp * vpad;
    const ctp_sample_cfg* cfgp = a.st ? &a.st->cfg : &a.cfg;
    // scalar copies: the configuration lives in device memory (GenState) and must not be re-read inside the loops below
    struct { float rep_penalty, top_p; int rep_window, rep_max_ids, top_k, min_keep, eos, min_new; unsigned long long seed; } cfg;
    cfg.rep_penalty = cfgp->rep_penalty; cfg.top_p = cfgp->top_p; cfg.rep_window = cfgp->rep_window; cfg.rep_max_ids = cfgp->rep_max_ids;
    cfg.top_k = cfgp->top_k; cfg.min_keep = cfgp->min_keep; cfg.eos = cfgp->eos; cfg.min_new = cfgp->min_new; cfg.seed = cfgp->seed;
    const float inv_t = 1.0f / cfgp->temperature[warp];
    const int step = a.st ? a.st->step : a.step;
    const int hist_len = a.st ? a.st->step : a.hist_len;
    const int* hist = a.st ? a.st->ids_buf : a.hist;
    const int hstride = a.st ? a.st->max_new : a.hist_stride;
    const float* u_ptr = a.st ? (a.st->u_base ? a.st->u_base + (long long)step * a.rows : nullptr) : a.u;
    float uu;
    if (u_ptr) uu = u_ptr[row];
    else uu = (float)(philox_mix(cfg.seed, (uint32_t)step, (uint32_t)row) >> 8) * (1.0f / 16777216.0f);
    // windowed repetition counts (processors.py:18-34): which token this lane penalises and by how much
    const bool rep_on = cfg.rep_penalty != 1.0f && row < cfg.rep_max_ids;
    int my = -1;
    float alpha = 1.f;
    bool rep_apply = false;
    if (rep_on) {
        const int w = min(hist_len, cfg.rep_window);
        if (lane < w) my = hist[((long long)b * hstride + (hist_len - w + lane)) * a.ids_cols + warp];
        int mult = 0;
        bool first = true;
        for (int t = 0; t < w; ++t) {
            const int other = __shfl_sync(0xffffffffu, my, t);
            if (other == my && lane < w) {
                ++mult;
                if (t < lane) first = false;
            }
        }
        rep_apply = lane < w && first && my >= 0 && my < V;
        if (rep_apply) alpha = powf(cfg.rep_penalty, (float)mult);
    }
    // bookkeeping state of the previous step (written by the previous sampler launch)
    int max_new_st = 0, end_prev = 0;
    bool fin_prev = false;
    if (a.st) { max_new_st = a.st->max_new; fin_prev = a.st->finish[b] != 0; end_prev = a.st->end_idx[b]; }
    if (PDL_INSIDE) {
        pdl_wait();
        if (threadIdx.x == 0) trace_mark(a.trace, 1);
    }

    // 1. temperature (gpt.py:469)
    const float* lg = a.logits + (long long)row * V;
    for (int v = lane; v < V; v += 32) sc[v] = (FUSED ? __ldcg(lg + v) : lg[v]) * inv_t;
    __syncwarp();
    if (!FUSED && threadIdx.x == 0) trace_mark(a.trace, 4);
    // 2. windowed repetition penalty (processors.py:18-34)
    if (rep_on) {
        if (rep_apply) {
            const float s = sc[my];
            sc[my] = (s < 0.f) ? s * alpha : s / alpha;
        }
        __syncwarp();
    }
    if (!FUSED && threadIdx.x == 0) trace_mark(a.trace, 7);
    // full-row softmax normaliser (TopP works on probabilities of the whole row)
    float mx = -INFINITY;
    for (int v = lane; v < V; v += 32) mx = fmaxf(mx, sc[v]);
    mx = warp_max(mx);
    float z = 0.f;
    for (int v = lane; v < V; v += 32) z += __expf(sc[v] - mx);
    z = warp_sum(z);
    // How many ranks TopK lets through: max(top_k, min_keep) (transformers TopKLogitsWarper), or the whole row when there is no TopK
    // warper (processors.py:43-47 with top_K=None).  Up to SAMPLE_MAX_K ranks are handled out of registers (lane k = rank k); wider
    // settings take the general path below.
    int K_req = cfg.top_k > 0 ? max(cfg.top_k, cfg.min_keep) : V;
    K_req = min(K_req, V);
    int chosen;
    if (K_req <= SAMPLE_MAX_K) {
        // 3. the K best scores in rank order (K = max(top_k, min_keep))
        const int K = K_req;
        float my_val = -INFINITY;  // lane k holds rank-k candidate
        int my_idx = -1;
        constexpr int REG_V = 20;   // audio vocab (626) fits 20 scores per lane: selection runs out of registers
        if (V <= 32 * REG_V) {
            // lane owns scores v = lane + 32*i.  Per round: warp arg-max over the lanes' local maxima with three redux.sync
            // (max of an order-preserving integer key, then the smallest token id among the ties), the owner retires its entry
            // and rescans its 20 registers.  ~2 us for K = 20 instead of ~13 us through shared memory (tests/prof_trace.py).
            float vals[REG_V];
    #pragma unroll
            for (int i = 0; i < REG_V; ++i) { const int v = lane + 32 * i; vals[i] = v < V ? sc[v] : -INFINITY; }
            const float inv_z = 1.0f / z;
            const bool stop_on_mass = cfg.top_p > 0.f && cfg.top_p < 1.f;
            float mass = 0.f;   // probability mass of the candidates selected so far (all lanes hold the same value)
            for (int k = 0; k < K; ++k) {
                // TopP removes every rank whose mass strictly above reaches top_p (unless rank < min_keep): once the selected candidates
                // hold that much, no further rank can survive and the remaining rounds are skipped (peaked rows need 2-4 rounds, not 20)
                // (margin: the decision itself is taken below from the prefix sums, exactly as before; this only prunes rounds)
                if (stop_on_mass && k >= cfg.min_keep && mass >= cfg.top_p + 1e-3f) break;
                // local arg-max as four independent chains of five, then merged (ties keep the smaller index)
                float cv[4];
                int ci[4];
    #pragma unroll
                for (int c = 0; c < 4; ++c) {
                    cv[c] = vals[5 * c]; ci[c] = 5 * c;
    #pragma unroll
                    for (int i = 1; i < 5; ++i) if (vals[5 * c + i] > cv[c]) { cv[c] = vals[5 * c + i]; ci[c] = 5 * c + i; }
                }
                if (cv[1] > cv[0]) { cv[0] = cv[1]; ci[0] = ci[1]; }
                if (cv[3] > cv[2]) { cv[2] = cv[3]; ci[2] = ci[3]; }
                float lv = cv[0];
                int li = ci[0];
                if (cv[2] > lv) { lv = cv[2]; li = ci[2]; }
                const unsigned u = __float_as_uint(lv);
                const unsigned key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
                const unsigned kmax = __reduce_max_sync(0xffffffffu, key);
                const int cand = (key == kmax) ? (lane + 32 * li) : 0x7fffffff;
                const int bi = __reduce_min_sync(0xffffffffu, cand);
                const float bv = __shfl_sync(0xffffffffu, lv, bi & 31);
                if (lane == k) { my_val = bv; my_idx = bi; }
                mass += __expf(bv - mx) * inv_z;
                if ((bi & 31) == lane) {
    #pragma unroll
                    for (int i = 0; i < REG_V; ++i) if (i == (bi >> 5)) vals[i] = -INFINITY;
                }
            }
        } else {
            for (int k = 0; k < K; ++k) {
                float bv = -INFINITY;
                int bi = 0x7fffffff;
                for (int v = lane; v < V; v += 32) {
                    const float s = sc[v];
                    if (s > bv) { bv = s; bi = v; }
                }
    #pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                if (lane == k) { my_val = bv; my_idx = bi; }
                if (lane == 0 && bi < V) sc[bi] = -INFINITY;  // remove from the pool
                __syncwarp();
            }
        }
        if (!FUSED && threadIdx.x == 0) trace_mark(a.trace, 5);
        // 4. TopP on rank order: token at rank k survives iff (mass strictly above it) < top_p, or k < min_keep.
        //    (reference: cumulative prob from the bottom <= 1 - top_p is removed.)
        const float pk = (lane < K && my_idx >= 0 && my_idx < V) ? __expf(my_val - mx) / z : 0.f;
        float above = pk;  // inclusive prefix sum over lanes, then make exclusive
    #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_up_sync(0xffffffffu, above, o);
            if (lane >= o) above += t;
        }
        above -= pk;
        bool keep = (lane < K) && (my_idx >= 0 && my_idx < V) && (my_val > -INFINITY);
        if (cfg.top_p > 0.f && cfg.top_p < 1.f) {
            // removed iff 1 - above <= 1 - top_p  (written like the reference's comparison on the ascending cumsum)
            const float cum_from_bottom = 1.0f - above;
            if (lane >= cfg.min_keep && cum_from_bottom <= (1.0f - cfg.top_p)) keep = false;
        }
        // 5. min-length EOS ban (gpt.py:477-478)
        if (step < cfg.min_new && my_idx == cfg.eos) keep = false;
        // 6. softmax over survivors and inverse-CDF draw in token-id order.  The softmax is taken relative to the SURVIVORS' maximum:
        //    when the min-length ban removed an EOS that was the row maximum, exp(s - row max) underflows to 0 for every survivor at
        //    near-greedy temperatures (the reference's softmax runs after the ban, gpt.py:477-480, and renormalises by itself)
        const float mx_s = warp_max(keep ? my_val : -INFINITY);
        const float e = keep ? __expf(my_val - mx_s) : 0.f;
        const float zs = warp_sum(e);
        const float p = e / zs;
        // rank of my token id among survivors
        float cdf_before = 0.f;  // mass of surviving tokens with a smaller id
        for (int t = 0; t < K; ++t) {
            const int oi = __shfl_sync(0xffffffffu, my_idx, t);
            const float op = __shfl_sync(0xffffffffu, p, t);
            if (oi < my_idx) cdf_before += op;
        }
        // chosen = survivor with cdf_before <= u < cdf_before + p; fall back to the largest id survivor
        const bool hit = keep && (cdf_before <= uu) && (uu < cdf_before + p);
        unsigned ballot = __ballot_sync(0xffffffffu, hit);
        if (ballot) {
            chosen = __shfl_sync(0xffffffffu, my_idx, __ffs(ballot) - 1);
        } else {
            int best = keep ? my_idx : -1;
    #pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
            chosen = best;
        }
        if (a.probs_out) {
            float* po = a.probs_out + (long long)row * V;
            for (int v = lane; v < V; v += 32) po[v] = 0.f;
            __syncwarp();
            if (keep) po[my_idx] = p;
        }
    } else {
        // ---- general path (top_K > 32 or no TopK at all): survivors are described by a cut-off key instead of being held one per lane.
        // Pass 1 walks the ranks in order (score descending, ties by token id ascending) WITHOUT removing anything: each round takes
        // the best key strictly after the previous one; rank n survives iff n < K_req and (n < min_keep or the mass strictly above it
        // is < top_p).  Pass 2 takes the softmax over {key <= cut-off} minus a banned EOS and draws by inverse CDF in token-id order.
        const bool use_p = cfg.top_p > 0.f && cfg.top_p < 1.f;
        float prev_s = INFINITY, cut_s = INFINITY, mass = 0.f;
        int prev_i = -1, cut_i = -1, n_sel = 0;
        while (n_sel < K_req) {
            float bs = -INFINITY;
            int bi = 0x7fffffff;
            for (int v = lane; v < V; v += 32) {
                const float sv = sc[v];
                const bool after = (sv < prev_s) || (sv == prev_s && v > prev_i);
                if (after && (sv > bs || (sv == bs && v < bi))) { bs = sv; bi = v; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bs, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bs || (ov == bs && oi < bi)) { bs = ov; bi = oi; }
            }
            if (bi == 0x7fffffff || bs == -INFINITY) break;                       // nothing finite is left
            if (use_p && n_sel >= cfg.min_keep && (1.0f - mass) <= (1.0f - cfg.top_p)) break;   // same comparison as the reference's cumsum test
            cut_s = bs; cut_i = bi;
            mass += __expf(bs - mx) / z;
            prev_s = bs; prev_i = bi;
            ++n_sel;
        }
        const bool ban = step < cfg.min_new;
        float mx_s = -INFINITY;
        for (int v = lane; v < V; v += 32) {
            const float sv = sc[v];
            const bool surv = n_sel > 0 && (sv > cut_s || (sv == cut_s && v <= cut_i)) && !(ban && v == cfg.eos);
            if (surv) mx_s = fmaxf(mx_s, sv);
        }
        mx_s = warp_max(mx_s);
        float zs = 0.f;
        for (int v = lane; v < V; v += 32) {
            const float sv = sc[v];
            const bool surv = n_sel > 0 && (sv > cut_s || (sv == cut_s && v <= cut_i)) && !(ban && v == cfg.eos);
            if (surv) zs += __expf(sv - mx_s);
        }
        zs = warp_sum(zs);
        float running = 0.f;
        int last_surv = -1;
        chosen = -1;
        float* po = a.probs_out ? a.probs_out + (long long)row * V : nullptr;
        for (int base = 0; base < V; base += 32) {
            const int v = base + lane;
            const float sv = v < V ? sc[v] : -INFINITY;
            const bool surv = v < V && n_sel > 0 && (sv > cut_s || (sv == cut_s && v <= cut_i)) && !(ban && v == cfg.eos);
            const float pv = surv ? __expf(sv - mx_s) / zs : 0.f;
            if (po && v < V) po[v] = pv;
            float incl = pv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const float before = running + incl - pv;
            const bool hit = surv && before <= uu && uu < before + pv;
            const unsigned hb = __ballot_sync(0xffffffffu, hit);
            const unsigned sb = __ballot_sync(0xffffffffu, surv);
            if (chosen < 0 && hb) chosen = base + __ffs(hb) - 1;
            if (sb) last_surv = base + 31 - __clz(sb);
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (chosen < 0) chosen = last_surv;   // u beyond the rounded total: the largest-id survivor, like the register path
    }
    if (a.next_ids && lane == 0) a.next_ids[row] = chosen;
    if (!FUSED && threadIdx.x == 0) trace_mark(a.trace, 6);
    if (a.st) {
        if (lane == 0) s_choice[warp] = chosen;
        if (FUSED) asm volatile("bar.sync 1, 256;" ::: "memory"); else __syncthreads();
        if (threadIdx.x == 0) {
            GenState* st = a.st;
            bool eos = false;
            for (int q = 0; q < a.num_vq; ++q) eos |= (s_choice[q] == cfg.eos);
            if (step < max_new_st) {   // text mode: the sampled id goes to every VQ column (gpt.py:489-494)
                for (int q = 0; q < a.ids_cols; ++q)
                    st->ids_buf[((long long)b * max_new_st + step) * a.ids_cols + q] = s_choice[a.num_vq == 1 ? 0 : q];
            }
            const bool fin = fin_prev || eos;                // gpt.py:486-487
            st->finish[b] = fin ? 1 : 0;
            if (!fin) st->end_idx[b] = end_prev + 1;         // gpt.py:530-531
            // one ticket per block; the high half counts the rows that have finished, so the last block to arrive knows
            // whether every sequence is done without re-reading the flags
            const int t = atomicAdd(&st->ticket, 1 + (fin ? 0x10000 : 0));
            if ((t & 0xffff) == n_blocks - 1) {              // last block: global bookkeeping
                st->ticket = 0;
                st->all_done = ((t >> 16) + (fin ? 1 : 0)) == n_blocks;
                st->step = step + 1;
                if (FUSED || a.advance_len) { st->cur_len += 1; }
            }
        }
    }
}

// blockDim = 32 * num_vq; block b handles rows b*num_vq .. b*num_vq + num_vq-1.
__global__ void k_sample(SampleArgs a) {
    if (threadIdx.x == 0) trace_mark(a.trace, 0);
    pdl_launch_dependents();
    extern __shared__ float s_scores_dyn[];  // [num_vq][vocab_pad]
    __shared__ int s_choice[MAX_VQ];
    sample_block<false, true>(a, blockIdx.x, gridDim.x, s_scores_dyn, s_choice);
    if (threadIdx.x == 0) trace_end(a.trace);
}

// cur_len += 1 after a trunk step (the new token's K/V now occupy slot cur_len)
__global__ void k_advance_len(GenState* st) {
    pdl_launch_dependents();
    pdl_wait();
    st->cur_len += 1;
}

}  // namespace ctp
