// GPT decoder handle: weight binding, prefill, decode step, sampler, device-resident generate loop.
// C-ABI entry points are declared in include/ctp.h (each cites the reference interface it replaces).
#include "gemm.cuh"
#include "gpt_kernels.cuh"
#include "mlp_kernel.cuh"

#include <map>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

using namespace ctp;

namespace {

struct LayerMaps {
    CUtensorMap wqkv, wo, wgu, wdown;
    CUtensorMap mlp_gu, mlp_d;   // fused MLP kernel: W_gu in boxes of 32 rows, W_down in [128 x 32] tiles (64B swizzle)
};

struct ActMaps {  // activation operands (N side) for one N-tile width
    CUtensorMap xn, attn, hmid;
};

struct GraphKey {
    int B, nsplit, text;
    bool operator<(const GraphKey& o) const { return B != o.B ? B < o.B : (nsplit != o.nsplit ? nsplit < o.nsplit : text < o.text); }
};

}  // namespace

static int g_split_target = 296;   // CTAs per decode GEMM: two per SM (83 KB of shared memory each) measured best; CTP_GEMM_CTAS, re-read at every handle creation

struct ctp_gpt {
    ctp_gpt_cfg cfg{};
    ctp_gpt_weights w{};
    bool bound = false;
    int dev = 0;

    // device workspace (rows = max(max_batch, prefill tokens))
    long long ws_rows = 0;
    float* x = nullptr;        // [rows][H] residual stream
    float* x2 = nullptr;       // [64][H] decode only: the fused MLP kernel reads one residual buffer and reduces into the other
    __half* xn = nullptr;      // [rows][H]
    float* acc_qkv = nullptr;  // [rows][3H]
    __half* attn = nullptr;    // [rows][H]
    float* acc_gu = nullptr;   // [rows][2I]
    __half* hmid = nullptr;    // [rows][I]
    float* x_last = nullptr;   // [maxB][H] (prefill: gathered last positions)
    float* hidden = nullptr;   // [maxB][H]
    float* logits = nullptr;   // [maxB][num_vq*num_audio]
    __half* kv = nullptr;      // [L][2][maxB][nH][maxS][64]
    float* attn_part = nullptr;
    int* attn_cnt = nullptr;
    int* pad_len = nullptr;    // [maxB]
    float* inv_freq = nullptr; // [32]
    GenState* st = nullptr;    // device
    GenState* st_pin = nullptr;  // pinned host mirror for polling
    int max_splits = 16;

    std::vector<LayerMaps> lmaps;
    CUtensorMap head_map{};
    CUtensorMap head_text_map{};
    bool have_head_text = false;
    int text_mode = 0;       // current generation is the refine-text pass
    ActMaps act32{}, act64{};
    // decode activation maps point at the first max_batch rows of xn / attn / hmid
    std::map<GraphKey, cudaGraphExec_t> graphs;
    std::map<GraphKey, long long> graph_nodes;
    cudaStream_t cap_stream = nullptr;
    cudaEvent_t poll_ev = nullptr;   // marks the lagged copy of the all_done flag in ctp_gpt_generate

    int sm_count = 0;
    bool use_pdl = true;       // programmatic dependent launch between the kernels of the decode step graph (CTP_PDL=0 disables)
    bool fuse_norm = true;     // RMSNorm folded into the QKV / gate|up GEMMs (XNORM kernel); CTP_FUSE_NORM=0: stand-alone norm kernels
    CUtensorMap x_map{};       // fp32 map over the first 64 rows of the residual stream
    CUtensorMap x2_map{};      // ... of the second residual buffer
    int mlp_m64 = 1;           // MMA 1 of the fused MLP kernel as M = 64 (CTP_MLP_M64=0: M = 128 with idle rows)
    int mlp_cluster = 0;       // fused MLP kernel (mlp_kernel.cuh): cluster size H/128; 0 = shape / device not eligible or CTP_MLP=0
    CUtensorMap gu_map{};      // fp32 map over the decode gate|up accumulator
    // decode only, re-armable scratch: ss1[64] | ss2[64] | gate|up accumulator [64][2I].  ss1 / ss2 = sum(x^2) per token row as seen by
    // the QKV / gate|up GEMM (input / post-attention RMSNorm folded in)
    float* dec_gu = nullptr;
    float* ss1() const { return dec_gu; }
    float* ss2() const { return dec_gu + 64; }
    float* gu_acc() const { return dec_gu + 128; }
    int attn_cta_target = 296;   // split the KV range until B*heads*nsplit reaches this many CTAs (CTP_ATTN_CTAS)
    // bring-up: in-graph timeline (CTP_TRACE=1): one 8-stamp record per kernel of the step graph, in launch order
    unsigned long long* trace = nullptr;
    int trace_n = 0;
    unsigned long long* trace_rec() { return trace ? trace + 8 * (size_t)(trace_n++ % 256) : nullptr; }

    // host mirror of the generation state
    int B = 0, cur_len = 0, step = 0, max_new = 0, prompt_len = 0;
    bool have_bufs = false;

    size_t kv_plane_elems() const { return (size_t)cfg.max_batch * cfg.n_heads * cfg.max_seq * HEAD_DIM; }
    __half* kplane(int l) const { return kv + (size_t)(2 * l) * kv_plane_elems(); }
    __half* vplane(int l) const { return kv + (size_t)(2 * l + 1) * kv_plane_elems(); }
};

static int ensure_workspace(ctp_gpt* h, long long rows) {
    if (rows < 64) rows = 64;  // decode activation tensor maps always span 64 rows (widest N tile)
    if (rows <= h->ws_rows) return CTP_OK;
    const int H = h->cfg.hidden, I = h->cfg.inter;
    cudaFree(h->x); cudaFree(h->xn); cudaFree(h->acc_qkv); cudaFree(h->attn); cudaFree(h->acc_gu); cudaFree(h->hmid);
    h->x = nullptr; h->xn = nullptr; h->acc_qkv = nullptr; h->attn = nullptr; h->acc_gu = nullptr; h->hmid = nullptr;
    h->ws_rows = 0;
    CTP_CUDA_OK(cudaMalloc(&h->x, sizeof(float) * rows * H));
    CTP_CUDA_OK(cudaMalloc(&h->xn, sizeof(__half) * rows * H));
    CTP_CUDA_OK(cudaMalloc(&h->acc_qkv, sizeof(float) * rows * 3 * H));
    CTP_CUDA_OK(cudaMalloc(&h->attn, sizeof(__half) * rows * H));
    CTP_CUDA_OK(cudaMalloc(&h->acc_gu, sizeof(float) * rows * 2 * I));
    CTP_CUDA_OK(cudaMalloc(&h->hmid, sizeof(__half) * rows * I));
    CTP_CUDA_OK(cudaMemset(h->xn, 0, sizeof(__half) * rows * H));
    CTP_CUDA_OK(cudaMemset(h->attn, 0, sizeof(__half) * rows * H));
    CTP_CUDA_OK(cudaMemset(h->hmid, 0, sizeof(__half) * rows * I));
    h->ws_rows = rows;
    // decode activation maps depend on the buffer addresses
    const int mb = 64;  // rows past the live batch hold zeros/stale data; their output columns are masked
    int st;
    for (int bn : {32, 64}) {
        ActMaps& am = (bn == 32) ? h->act32 : h->act64;
        if ((st = make_tmap_kmajor(&am.xn, h->xn, mb, H, H, bn))) return st;
        if ((st = make_tmap_kmajor(&am.attn, h->attn, mb, H, H, bn))) return st;
        if ((st = make_tmap_kmajor(&am.hmid, h->hmid, mb, I, I, bn))) return st;
    }
    if ((st = make_tmap_f32(&h->x_map, h->x, mb, H, H, 32))) return st;
    if (h->x2 && (st = make_tmap_f32(&h->x2_map, h->x2, mb, H, H, 32))) return st;
    if (h->dec_gu && (st = make_tmap_f32(&h->gu_map, h->gu_acc(), mb, 2 * I, 2 * I, 32))) return st;
    // captured graphs hold the old pointers
    for (auto& kvp : h->graphs) cudaGraphExecDestroy(kvp.second);
    h->graphs.clear();
    return CTP_OK;
}

extern "C" ctp_status ctp_gpt_create(ctp_gpt** out, const ctp_gpt_cfg* cfg) {
    CTP_REQUIRE(out && cfg, "ctp_gpt_create: null argument");
    CTP_REQUIRE(cfg->n_heads * HEAD_DIM == cfg->hidden, "head_dim must be 64 (hidden %d, heads %d)", cfg->hidden, cfg->n_heads);
    CTP_REQUIRE(cfg->hidden % 64 == 0 && cfg->hidden <= 1024 && cfg->inter % 64 == 0, "hidden/inter must be multiples of 64, hidden <= 1024");
    CTP_REQUIRE(cfg->n_heads * 32 <= 1024, "too many heads");
    CTP_REQUIRE(cfg->max_batch >= 1 && cfg->max_batch <= 64, "max_batch must be in [1,64]");
    CTP_REQUIRE(cfg->num_vq >= 1 && cfg->num_vq <= MAX_VQ, "num_vq must be in [1,%d]", MAX_VQ);
    CTP_REQUIRE(cfg->max_seq >= 2, "max_seq too small");
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        ctp_set_error("no CUDA device: libctp has no CPU fallback");
        return CTP_ERR_NO_DEVICE;
    }
    ctp_status ds = ctp_device_check(dev);
    if (ds != CTP_OK) return ds;
    int gi = gemm_init();
    if (gi) return (ctp_status)gi;
    ctp_gpt* h = new ctp_gpt();
    h->cfg = *cfg;
    h->dev = dev;
    const int H = cfg->hidden, mb = cfg->max_batch;
    const size_t kv_elems = (size_t)cfg->n_layers * 2 * h->kv_plane_elems();
#define CK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { ctp_set_error("%s -> %s", #expr, cudaGetErrorString(_e)); ctp_gpt_destroy(h); return CTP_ERR_CUDA; } } while (0)
    CK(cudaMalloc(&h->kv, kv_elems * sizeof(__half)));
    CK(cudaMemset(h->kv, 0, kv_elems * sizeof(__half)));
    CK(cudaMalloc(&h->x_last, sizeof(float) * mb * H));
    CK(cudaMalloc(&h->hidden, sizeof(float) * mb * H));
    const size_t lg = (size_t)mb * std::max(cfg->num_vq * cfg->num_audio, cfg->num_text);
    CK(cudaMalloc(&h->logits, sizeof(float) * lg));
    CK(cudaMalloc(&h->attn_part, sizeof(float) * (size_t)mb * cfg->n_heads * h->max_splits * 66));
    CK(cudaMalloc(&h->attn_cnt, sizeof(int) * mb * cfg->n_heads));
    CK(cudaMemset(h->attn_cnt, 0, sizeof(int) * mb * cfg->n_heads));
    CK(cudaMalloc(&h->pad_len, sizeof(int) * mb));
    CK(cudaMemset(h->pad_len, 0, sizeof(int) * mb));
    CK(cudaMalloc(&h->inv_freq, sizeof(float) * 32));
    CK(cudaMalloc(&h->st, sizeof(GenState)));
    CK(cudaMemset(h->st, 0, sizeof(GenState)));
    CK(cudaMallocHost(&h->st_pin, sizeof(GenState)));
    CK(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->poll_ev, cudaEventDisableTiming));
    CK(cudaFuncSetAttribute(k_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    {
        // llama.py:98 — inv_freq = 1 / base^(2i/d), evaluated in fp32 like torch does
        float f[32];
        for (int i = 0; i < 32; ++i) f[i] = 1.0f / powf(cfg->rope_theta, (float)(2 * i) / (float)HEAD_DIM);
        CK(cudaMemcpy(h->inv_freq, f, sizeof(f), cudaMemcpyHostToDevice));
    }
    {   // decode-path resources and bring-up switches (read once, here)
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, dev));
        h->sm_count = prop.multiProcessorCount;
        const int I = cfg->inter;
        if (const char* e = getenv("CTP_PDL")) h->use_pdl = atoi(e) != 0;
        g_split_target = getenv("CTP_GEMM_CTAS") ? atoi(getenv("CTP_GEMM_CTAS")) : 296;
        if (const char* e = getenv("CTP_FUSE_NORM")) h->fuse_norm = atoi(e) != 0;
        if (const char* e = getenv("CTP_ATTN_CTAS")) h->attn_cta_target = atoi(e);
        CK(cudaMalloc(&h->dec_gu, sizeof(float) * (128 + (size_t)64 * 2 * I)));
        CK(cudaMemset(h->dec_gu, 0, sizeof(float) * (128 + (size_t)64 * 2 * I)));
        CK(cudaMalloc(&h->x2, sizeof(float) * 64 * H));
        CK(cudaMemset(h->x2, 0, sizeof(float) * 64 * H));
        {   // fused MLP kernel: clusters of H/128 CTAs, 32 intermediate features per CTA (CTP_MLP=0 keeps the two-GEMM path)
            const int C = H / 128;
            if (const char* e = getenv("CTP_MLP_M64")) h->mlp_m64 = atoi(e) != 0;
            bool ok = H % 128 == 0 && C >= 2 && C <= 8 && I % (MLP_UW * C) == 0 && (int)prop.sharedMemPerBlockOptin >= mlp_smem_bytes(H);
            if (const char* e = getenv("CTP_MLP")) if (!atoi(e)) ok = false;
            if (ok) {
                CK(cudaFuncSetAttribute(k_mlp_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, mlp_smem_bytes(H)));
                // the cluster must be schedulable (C SMs of one GPC with this much shared memory each)
                cudaLaunchConfig_t qc{};
                qc.gridDim = dim3((unsigned)(I / MLP_UW)); qc.blockDim = dim3(MLP_THREADS); qc.dynamicSmemBytes = (size_t)mlp_smem_bytes(H);
                cudaLaunchAttribute qa[1];
                qa[0].id = cudaLaunchAttributeClusterDimension; qa[0].val.clusterDim.x = (unsigned)C; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
                qc.attrs = qa; qc.numAttrs = 1;
                int n_clusters = 0;
                if (cudaOccupancyMaxActiveClusters(&n_clusters, k_mlp_fused, &qc) != cudaSuccess || n_clusters < 1) { ok = false; cudaGetLastError(); }
            }
            h->mlp_cluster = ok ? C : 0;
        }
        CK(cudaFuncSetAttribute(k_attn_decode_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
        if (const char* e = getenv("CTP_TRACE")) if (atoi(e)) {
            CK(cudaMalloc(&h->trace, sizeof(unsigned long long) * (8 * 256 + 128)));
            CK(cudaMemset(h->trace, 0, sizeof(unsigned long long) * (8 * 256 + 128)));
        }
    }
#undef CK
    int st = ensure_workspace(h, mb);
    if (st) { ctp_gpt_destroy(h); return (ctp_status)st; }
    *out = h;
    return CTP_OK;
}

extern "C" void ctp_gpt_destroy(ctp_gpt* h) {
    if (!h) return;
    for (auto& kvp : h->graphs) cudaGraphExecDestroy(kvp.second);
    cudaFree(h->x); cudaFree(h->xn); cudaFree(h->acc_qkv); cudaFree(h->attn); cudaFree(h->acc_gu); cudaFree(h->hmid);
    cudaFree(h->trace); cudaFree(h->dec_gu); cudaFree(h->x2);
    cudaFree(h->x_last); cudaFree(h->hidden); cudaFree(h->logits); cudaFree(h->kv); cudaFree(h->attn_part);
    cudaFree(h->attn_cnt); cudaFree(h->pad_len); cudaFree(h->inv_freq); cudaFree(h->st);
    if (h->st_pin) cudaFreeHost(h->st_pin);
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    if (h->poll_ev) cudaEventDestroy(h->poll_ev);
    delete h;
}

extern "C" ctp_status ctp_gpt_bind_weights(ctp_gpt* h, const ctp_gpt_weights* w) {
    CTP_REQUIRE(h && w, "bind: null argument");
    CTP_REQUIRE(w->wqkv && w->wo && w->wgu && w->wdown && w->ln1 && w->ln2 && w->norm_f && w->emb_code && w->head_code,
                "bind: missing weight pointer");
    const ctp_gpt_cfg& c = h->cfg;
    const long long H = c.hidden, I = c.inter;
    h->lmaps.resize(c.n_layers);
    int st;
    for (int l = 0; l < c.n_layers; ++l) {
        LayerMaps& m = h->lmaps[l];
        const __half* wqkv = (const __half*)w->wqkv + (size_t)l * 3 * H * H;
        const __half* wo = (const __half*)w->wo + (size_t)l * H * H;
        const __half* wgu = (const __half*)w->wgu + (size_t)l * 2 * I * H;
        const __half* wd = (const __half*)w->wdown + (size_t)l * H * I;
        if ((st = make_tmap_kmajor(&m.wqkv, wqkv, 3 * H, H, H, GEMM_BM))) return (ctp_status)st;
        if ((st = make_tmap_kmajor(&m.wo, wo, H, H, H, GEMM_BM))) return (ctp_status)st;
        if ((st = make_tmap_kmajor(&m.wgu, wgu, 2 * I, H, H, GEMM_BM))) return (ctp_status)st;
        if ((st = make_tmap_kmajor(&m.wdown, wd, H, I, I, GEMM_BM))) return (ctp_status)st;
        if (h->mlp_cluster && (st = make_tmap_kmajor(&m.mlp_gu, wgu, 2 * I, H, H, MLP_UW))) return (ctp_status)st;
        if (h->mlp_cluster && (st = make_tmap_k32_sw64(&m.mlp_d, wd, H, I, I, 128))) return (ctp_status)st;
    }
    if ((st = make_tmap_kmajor(&h->head_map, w->head_code, (long long)c.num_vq * c.num_audio, H, H, GEMM_BM))) return (ctp_status)st;
    h->have_head_text = false;
    if (w->head_text && w->emb_text) {
        if ((st = make_tmap_kmajor(&h->head_text_map, w->head_text, (long long)c.num_text, H, H, GEMM_BM))) return (ctp_status)st;
        h->have_head_text = true;
    }
    h->w = *w;
    h->bound = true;
    // weights are baked into captured graphs as tensor maps
    for (auto& kvp : h->graphs) cudaGraphExecDestroy(kvp.second);
    h->graphs.clear();
    return CTP_OK;
}

extern "C" ctp_status ctp_gpt_embed_prompt(ctp_gpt* h, int32_t B, int32_t L0, const int32_t* ids, const uint8_t* text_mask,
                                           float* emb_out, ctp_stream stream) {
    CTP_REQUIRE(h && h->bound, "embed_prompt: weights not bound");
    CTP_REQUIRE(h->w.emb_text, "embed_prompt: emb_text not bound");
    CTP_REQUIRE(B >= 1 && L0 >= 1 && ids && text_mask && emb_out, "embed_prompt: bad argument");
    k_embed_prompt<<<B * L0, 256, 0, (cudaStream_t)stream>>>(ids, text_mask, (const __half*)h->w.emb_text,
                                                             (const __half*)h->w.emb_code, emb_out, h->cfg.hidden,
                                                             h->cfg.num_vq, h->cfg.num_audio);
    ctp_count_launch();
    CTP_CUDA_OK(cudaGetLastError());
    return CTP_OK;
}

// ---- shared launch helpers ------------------------------------------------------------------------------
static GemmEpilogue epi_swap_atomic(float* out, long long ldo, int T, int F) {
    GemmEpilogue e{};
    e.out = out; e.ldo = ldo; e.out_f16 = 0; e.atomic = 1; e.swap = 1; e.T = T; e.F = F;
    return e;
}

static int split_for(int k_blocks, int m_tiles, int target_ctas = 0) {
    if (target_ctas <= 0) target_ctas = g_split_target;
    int s = target_ctas / m_tiles;
    if (s < 1) s = 1;
    if (s > k_blocks) s = k_blocks;
    return s;
}

#define LAUNCH_OK() do { ctp_count_launch(); cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) { ctp_set_error("%s:%d launch: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); return CTP_ERR_CUDA; } } while (0)

// heads: logits[b][q*A + a] = hidden_n[b] . head_code[q*A + a]   (gpt.py:424-439; weight_norm folded at bind)
static int launch_heads(ctp_gpt* h, int B, cudaStream_t s, bool pdl = false, bool traced = false) {
    const ctp_gpt_cfg& c = h->cfg;
    const int F = h->text_mode ? c.num_text : c.num_vq * c.num_audio;   // head_text (gpt.py:425-426) or the 4 code heads
    const int bn = B <= 32 ? 32 : 64;
    const ActMaps& am = bn == 32 ? h->act32 : h->act64;
    const int m_tiles = (F + GEMM_BM - 1) / GEMM_BM;
    GemmEpilogue e = epi_swap_atomic(h->logits, F, B, F);
    return gemm_launch_maps(h->text_mode ? h->head_text_map : h->head_map, am.xn, F, B, c.hidden, bn, split_for(c.hidden / 64, m_tiles), e, s, nullptr, 0, pdl,
                            nullptr, 0, traced ? h->trace_rec() : nullptr);
}

// One decode trunk step for B sequences (ids_ext == nullptr -> codes of the previous sample step).
#define CTP_LAUNCH(kern, grid, block, smem, ...) do { cudaError_t _le = launch_k(kern, grid, block, (size_t)(smem), s, pdl, __VA_ARGS__); ctp_count_launch(); if (_le == cudaSuccess) _le = cudaGetLastError(); if (_le != cudaSuccess) { ctp_set_error("%s:%d launch %s: %s", __FILE__, __LINE__, #kern, cudaGetErrorString(_le)); return CTP_ERR_CUDA; } } while (0)

static int launch_attn(ctp_gpt* h, int l, int B, int nsplit, bool row_factor, cudaStream_t s, bool pdl) {
    const ctp_gpt_cfg& c = h->cfg;
    AttnDecArgs aa{};
    aa.qkv = h->acc_qkv; aa.kcache = h->kplane(l); aa.vcache = h->vplane(l); aa.out = h->attn; aa.part = h->attn_part;
    aa.counters = h->attn_cnt; aa.pad_len = h->pad_len; aa.st = h->st; aa.inv_freq = h->inv_freq;
    aa.H = c.hidden; aa.nH = c.n_heads; aa.max_seq = c.max_seq; aa.eps = c.rms_eps;
    if (row_factor) aa.ss = h->ss1();   // input_layernorm's row factor (llama.py:718), deferred from the QKV GEMM
    aa.trace = h->trace_rec();
    CTP_LAUNCH(k_attn_decode_tma, dim3(c.n_heads, B, nsplit), dim3(AT_THREADS), AT_SMEM, aa);
    return CTP_OK;
}

static int run_decode_trunk(ctp_gpt* h, int B, int nsplit, const int* ids_ext, cudaStream_t s, bool advance_in_sampler = false) {
    const ctp_gpt_cfg& c = h->cfg;
    const bool pdl = h->use_pdl;
    const int H = c.hidden, I = c.inter;
    const int bn = B <= 32 ? 32 : 64;
    const ActMaps& am = bn == 32 ? h->act32 : h->act64;
    int st;
    h->trace_n = 0;
    // Batch <= 32: RMSNorm is folded into the QKV / gate|up GEMMs and silu(gate)*up into down_proj (in-kernel token operands).
    const bool fuse = h->fuse_norm && B <= 32;
    float* gu = fuse ? h->gu_acc() : h->acc_gu;
    // fused MLP kernel (mlp_kernel.cuh): gate|up -> silu*up -> down in one launch per layer.  The residual stream then alternates between
    // two buffers: layer l reads / accumulates o_proj into cur, the MLP kernel reduces cur + mlp(cur) into the other buffer, which this
    // layer's QKV GEMM cleared (its last reader was the previous layer's MLP kernel).
    const bool mlp = fuse && h->mlp_cluster > 0;
    float* xcur = h->x;
    unsigned long long rearm_f4 = (unsigned long long)(64 + (size_t)B * 2 * I) / 4;   // ss2 | gate|up rows of the live batch (from ss2())
    if (mlp) rearm_f4 = (unsigned long long)((size_t)B * H) / 4;
    // default (batches of 33..64 rows use the 64-token N tile with stand-alone norm / SiLU kernels): one launch per op
    for (int l = 0; l < c.n_layers; ++l) {
        const bool f1 = fuse && l > 0;   // layer 0 keeps the norm kernel: it is also the code-embedding front end
        float* xoth = (xcur == h->x) ? h->x2 : h->x;
        float* rearm = mlp ? xoth : h->ss2();
        if (!f1) {
            NormArgs na{};
            na.x = xcur; na.w = h->w.ln1 + (size_t)l * H; na.xn = h->xn; na.H = H; na.eps = c.rms_eps;
            if (l == 0) { na.st = h->st; na.ids_ext = ids_ext; na.emb_code = (const __half*)h->w.emb_code; na.num_vq = c.num_vq; na.num_audio = c.num_audio;
                          na.emb_text = h->text_mode ? (const __half*)h->w.emb_text : nullptr; }
            na.trace = h->trace_rec();
            CTP_LAUNCH(k_rmsnorm, dim3(B), dim3(256), 0, na);
        }
        {   // q,k,v projections as one GEMM (llama.py:619-621), weights are the M operand; input_layernorm (llama.py:718) folded in
            // when f1 (attention applies the row factor).  Also re-arms ss2 / gate|up, last read by the previous layer's down GEMM.
            GemmEpilogue e = epi_swap_atomic(h->acc_qkv, 3 * H, B, 3 * H);
            const void* pf = (const __half*)h->w.wo + (size_t)l * H * H;
            const size_t pfb = sizeof(__half) * (size_t)H * H;
            if (f1) {
                GemmShape ex{};
                ex.pf_ptr = pf; ex.pf_bytes = pfb; ex.norm_w = h->w.ln1 + (size_t)l * H; ex.zero_ptr = rearm; ex.zero_f4 = rearm_f4;
                ex.ss_out = h->ss1(); ex.trace = h->trace_rec();
                st = gemm_launch_x(1, h->lmaps[l].wqkv, xcur == h->x ? h->x_map : h->x2_map, 3 * H, B, H, split_for(H / 64, 3 * H / GEMM_BM), e, ex, s, pdl);
            } else {
                st = gemm_launch_maps(h->lmaps[l].wqkv, am.xn, 3 * H, B, H, bn, split_for(H / 64, 3 * H / GEMM_BM), e, s, pf, pfb, pdl,
                                      fuse ? rearm : nullptr, fuse ? rearm_f4 : 0, h->trace_rec());
            }
            if (st) return st;
        }
        if ((st = launch_attn(h, l, B, nsplit, f1, s, pdl))) return st;
        {   // o_proj accumulated straight into the residual stream (llama.py:663-666,737)
            GemmEpilogue e = epi_swap_atomic(xcur, H, B, H);
            if ((st = gemm_launch_maps(h->lmaps[l].wo, am.attn, H, B, H, bn, split_for(H / 64, (H + GEMM_BM - 1) / GEMM_BM), e, s,
                                       (const __half*)h->w.wgu + (size_t)l * 2 * I * H, sizeof(__half) * (size_t)2 * I * H, pdl,
                                       fuse ? h->ss1() : nullptr, fuse ? 16 : 0, h->trace_rec()))) return st;   // re-arms ss1 (read by this layer's attention)
        }
        if (mlp) {
            MlpArgs ma{};
            ma.x_in = xcur; ma.x_out = xoth; ma.ln_w = h->w.ln2 + (size_t)l * H; ma.T = B; ma.H = H; ma.I = I; ma.eps = c.rms_eps;
            ma.pf_ptr = (l + 1 < c.n_layers) ? (const void*)((const __half*)h->w.wqkv + (size_t)(l + 1) * 3 * H * H) : h->w.head_code;
            ma.pf_bytes = (l + 1 < c.n_layers) ? sizeof(__half) * (size_t)3 * H * H : sizeof(__half) * (size_t)c.num_vq * c.num_audio * H;
            ma.trace = h->trace_rec();
            ma.m64 = h->mlp_m64;
            const LayerMaps& m = h->lmaps[l];
            {
                cudaError_t le = launch_kc(k_mlp_fused, dim3((unsigned)(I / MLP_UW)), dim3(MLP_THREADS), (size_t)mlp_smem_bytes(H), s, pdl, (unsigned)h->mlp_cluster,
                                           m.mlp_gu, m.mlp_d, ma);
                ctp_count_launch();
                if (le == cudaSuccess) le = cudaGetLastError();
                if (le != cudaSuccess) { ctp_set_error("%s:%d launch k_mlp_fused: %s", __FILE__, __LINE__, cudaGetErrorString(le)); return CTP_ERR_CUDA; }
            }
            xcur = xoth;
            continue;
        }
        if (!fuse) {
            NormArgs nb{};
            nb.x = h->x; nb.w = h->w.ln2 + (size_t)l * H; nb.xn = h->xn; nb.H = H; nb.eps = c.rms_eps;
            nb.trace = h->trace_rec();
            CTP_LAUNCH(k_rmsnorm, dim3(B), dim3(256), 0, nb);
        }
        {   // gate_proj | up_proj (llama.py:214); post_attention_layernorm (llama.py:741) folded in when fused: the GEMM contracts x*w
            // and leaves sum(x^2) per row in ss2 for the down GEMM's prologue
            GemmEpilogue e = epi_swap_atomic(gu, 2 * I, B, 2 * I);
            const void* pf = (const __half*)h->w.wdown + (size_t)l * H * I;
            const size_t pfb = sizeof(__half) * (size_t)H * I;
            if (fuse) {
                GemmShape ex{};
                ex.pf_ptr = pf; ex.pf_bytes = pfb; ex.norm_w = h->w.ln2 + (size_t)l * H; ex.ss_out = h->ss2(); ex.trace = h->trace_rec();
                st = gemm_launch_x(1, h->lmaps[l].wgu, h->x_map, 2 * I, B, H, split_for(H / 64, 2 * I / GEMM_BM), e, ex, s, pdl);
            } else {
                st = gemm_launch_maps(h->lmaps[l].wgu, am.xn, 2 * I, B, H, bn, split_for(H / 64, 2 * I / GEMM_BM), e, s, pf, pfb, pdl);
            }
            if (st) return st;
        }
        if (!fuse) {
            const long long total = (long long)B * I;
            CTP_LAUNCH(k_silu_mul, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, gu, h->hmid, I, total, 1);
        }
        {   // down_proj accumulated into the residual stream (llama.py:214,745); silu(gate)*up folded into its token operand when fused
            GemmEpilogue e = epi_swap_atomic(h->x, H, B, H);
            const void* nxt = (l + 1 < c.n_layers) ? (const void*)((const __half*)h->w.wqkv + (size_t)(l + 1) * 3 * H * H) : h->w.head_code;
            const size_t nxt_bytes = (l + 1 < c.n_layers) ? sizeof(__half) * (size_t)3 * H * H : sizeof(__half) * (size_t)c.num_vq * c.num_audio * H;
            if (fuse) {
                GemmShape ex{};
                ex.pf_ptr = nxt; ex.pf_bytes = nxt_bytes; ex.ss_in = h->ss2(); ex.ss_dim = (float)H; ex.eps = c.rms_eps; ex.up_off = I;
                ex.trace = h->trace_rec();
                st = gemm_launch_x(2, h->lmaps[l].wdown, h->gu_map, H, B, I, split_for(I / 64, (H + GEMM_BM - 1) / GEMM_BM), e, ex, s, pdl);
            } else {
                st = gemm_launch_maps(h->lmaps[l].wdown, am.hmid, H, B, I, bn, split_for(I / 64, (H + GEMM_BM - 1) / GEMM_BM), e, s, nxt, nxt_bytes, pdl);
            }
            if (st) return st;
        }
    }
    // final norm (llama.py:1002) -> hidden state of this step (gpt.py:422-423) + operand of the heads
    NormArgs nf{};
    nf.x = xcur; nf.w = h->w.norm_f; nf.xn = h->xn; nf.out_f32 = h->hidden; nf.H = H; nf.eps = c.rms_eps;
    nf.zero_buf = h->logits; nf.zero_n = h->text_mode ? c.num_text : c.num_vq * c.num_audio; nf.st = h->st; nf.write_hid = 1;
    nf.trace = h->trace_rec();
    CTP_LAUNCH(k_rmsnorm, dim3(B), dim3(256), 0, nf);
    if ((st = launch_heads(h, B, s, pdl, true))) return st;
    if (!advance_in_sampler) CTP_LAUNCH(k_advance_len, dim3(1), dim3(1), 0, h->st);
    return CTP_OK;
}

static int launch_sampler(ctp_gpt* h, int B, cudaStream_t s, bool pdl = false, bool advance_len = false) {
    const ctp_gpt_cfg& c = h->cfg;
    SampleArgs sa{};
    const int cols = h->text_mode ? 1 : c.num_vq;
    const int V = h->text_mode ? c.num_text : c.num_audio;
    sa.logits = h->logits; sa.vocab = V; sa.num_vq = cols; sa.ids_cols = c.num_vq; sa.rows = B * cols; sa.st = h->st; sa.advance_len = advance_len ? 1 : 0;
    sa.trace = h->trace_rec();
    const size_t smem = sizeof(float) * cols * ((V + 31) & ~31);
    CTP_LAUNCH(k_sample, dim3(B), dim3(32 * cols), smem, sa);
    return CTP_OK;
}

static int nsplit_for(const ctp_gpt* h, int B, int ctx_len) {
    // keep >= ~2 CTAs per SM worth of independent KV streams; one split per 512 slots beyond that
    int ns = 1;
    const int base = B * h->cfg.n_heads;
    while (ns < h->max_splits && (base * ns < h->attn_cta_target || ctx_len / ns > 1024)) ns *= 2;
    if (ctx_len < ns * 8) ns = 1;
    return ns;
}

extern "C" ctp_status ctp_gpt_prefill(ctp_gpt* h, int32_t B, int32_t L0, const float* emb, const int32_t* pad_len_host,
                                      const ctp_gen_buffers* bufs, int32_t infer_text, ctp_stream stream) {
    CTP_REQUIRE(h && h->bound, "prefill: weights not bound");
    CTP_REQUIRE(!infer_text || h->have_head_text, "prefill: infer_text needs emb_text / head_text bound");
    h->text_mode = infer_text ? 1 : 0;
    const ctp_gpt_cfg& c = h->cfg;
    CTP_REQUIRE(B >= 1 && B <= c.max_batch, "prefill: batch %d outside [1,%d]", B, c.max_batch);
    CTP_REQUIRE(L0 >= 1 && L0 < c.max_seq, "prefill: prompt length %d does not fit max_seq %d", L0, c.max_seq);
    CTP_REQUIRE(emb && bufs && bufs->ids && bufs->end_idx && bufs->finish && bufs->max_new >= 1, "prefill: bad buffers");
    cudaStream_t s = (cudaStream_t)stream;
    const int H = c.hidden, I = c.inter;
    const long long T = (long long)B * L0;
    int st = ensure_workspace(h, std::max<long long>(T, c.max_batch));
    if (st) return (ctp_status)st;
    std::vector<int> pads(B, 0);
    if (pad_len_host) {
        for (int b = 0; b < B; ++b) {
            CTP_REQUIRE(pad_len_host[b] >= 0 && pad_len_host[b] < L0, "prefill: pad_len[%d]=%d invalid", b, pad_len_host[b]);
            pads[b] = pad_len_host[b];
        }
    }
    CTP_CUDA_OK(cudaMemcpyAsync(h->pad_len, pads.data(), sizeof(int) * B, cudaMemcpyHostToDevice, s));
    // generation state
    GenState gs{};
    gs.cur_len = L0; gs.step = 0; gs.B = B; gs.max_new = bufs->max_new; gs.ids_buf = bufs->ids; gs.hid_buf = bufs->hiddens;
    gs.end_idx = bufs->end_idx; gs.finish = bufs->finish; gs.u_base = nullptr; gs.all_done = 0; gs.ticket = 0; gs.text_mode = infer_text ? 1 : 0;
    CTP_CUDA_OK(cudaMemcpyAsync(h->st, &gs, sizeof(gs), cudaMemcpyHostToDevice, s));
    CTP_CUDA_OK(cudaMemsetAsync(bufs->end_idx, 0, sizeof(int) * B, s));
    CTP_CUDA_OK(cudaMemsetAsync(bufs->finish, 0, B, s));
    CTP_CUDA_OK(cudaStreamSynchronize(s));  // pads / gs are stack/heap temporaries
    CTP_CUDA_OK(cudaMemcpyAsync(h->x, emb, sizeof(float) * T * H, cudaMemcpyDeviceToDevice, s));

    const bool fast_norm = H % 128 == 0;   // (k_rmsnorm: one 256-thread CTA per row, any H <= 1024)
    for (int l = 0; l < c.n_layers; ++l) {
        if (fast_norm) {
            k_rmsnorm_rows<<<(unsigned)((T + 7) / 8), 256, 0, s>>>(h->x, h->w.ln1 + (size_t)l * H, h->xn, T, H, c.rms_eps);
        } else {
            NormArgs na{};
            na.x = h->x; na.w = h->w.ln1 + (size_t)l * H; na.xn = h->xn; na.H = H; na.eps = c.rms_eps;
            k_rmsnorm<<<(unsigned)T, 256, 0, s>>>(na);
        }
        LAUNCH_OK();
        {   // tokens are the M operand here (T = B*L0 rows): compute-bound, 128x256 tiles
            GemmLaunch g{};
            g.A = h->xn; g.a_rows = T; g.lda = H;
            g.B = (const __half*)h->w.wqkv + (size_t)l * 3 * H * H; g.b_rows = 3 * H; g.ldb = H;
            g.K = H; g.block_n = 256; g.split_k = 1;
            g.epi.out = h->acc_qkv; g.epi.ldo = 3 * H; g.epi.T = (int)T; g.epi.F = 3 * H;
            if ((st = gemm_launch(g, s))) return (ctp_status)st;
        }
        k_rope_prefill<<<dim3(L0, B), c.n_heads * 32, 0, s>>>(h->acc_qkv, h->kplane(l), h->vplane(l), h->pad_len, h->inv_freq, H,
                                                               c.n_heads, L0, c.max_seq);
        LAUNCH_OK();
        static const bool scalar_attn = getenv("CTP_PREFILL_ATTN") && strcmp(getenv("CTP_PREFILL_ATTN"), "scalar") == 0;
        if (scalar_attn)
            k_attn_prefill<<<dim3(c.n_heads, B, (L0 + 7) / 8), 256, 0, s>>>(h->acc_qkv, h->kplane(l), h->vplane(l), h->attn, h->pad_len,
                                                                             H, c.n_heads, L0, c.max_seq);
        else
            k_attn_prefill_mma<<<dim3((L0 + PF_TILE - 1) / PF_TILE, c.n_heads, B), 128, 0, s>>>(h->acc_qkv, h->kplane(l), h->vplane(l), h->attn,
                                                                                                h->pad_len, H, c.n_heads, L0, c.max_seq);
        LAUNCH_OK();
        {
            GemmLaunch g{};
            g.A = h->attn; g.a_rows = T; g.lda = H;
            g.B = (const __half*)h->w.wo + (size_t)l * H * H; g.b_rows = H; g.ldb = H;
            g.K = H; g.block_n = 256; g.split_k = 1;
            g.epi.out = h->x; g.epi.ldo = H; g.epi.residual = h->x; g.epi.ldr = H; g.epi.T = (int)T; g.epi.F = H;
            if ((st = gemm_launch(g, s))) return (ctp_status)st;
        }
        if (fast_norm) {
            k_rmsnorm_rows<<<(unsigned)((T + 7) / 8), 256, 0, s>>>(h->x, h->w.ln2 + (size_t)l * H, h->xn, T, H, c.rms_eps);
        } else {
            NormArgs nb{};
            nb.x = h->x; nb.w = h->w.ln2 + (size_t)l * H; nb.xn = h->xn; nb.H = H; nb.eps = c.rms_eps;
            k_rmsnorm<<<(unsigned)T, 256, 0, s>>>(nb);
        }
        LAUNCH_OK();
        if (I % 128 == 0) {
            // gate_proj | up_proj with silu(gate) * up in the GEMM epilogue (llama.py:214): fp16 [T][I] out, no fp32 [T][2I] round trip
            if ((st = gemm_launch_swiglu(h->xn, T, H, (const __half*)h->w.wgu + (size_t)l * 2 * I * H, I, H, h->hmid, I, s))) return (ctp_status)st;
        } else {
            GemmLaunch g{};
            g.A = h->xn; g.a_rows = T; g.lda = H;
            g.B = (const __half*)h->w.wgu + (size_t)l * 2 * I * H; g.b_rows = 2 * I; g.ldb = H;
            g.K = H; g.block_n = 256; g.split_k = 1;
            g.epi.out = h->acc_gu; g.epi.ldo = 2 * I; g.epi.T = (int)T; g.epi.F = 2 * I;
            if ((st = gemm_launch(g, s))) return (ctp_status)st;
            const long long total = T * I;
            k_silu_mul<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(h->acc_gu, h->hmid, I, total, 0);
            LAUNCH_OK();
        }
        {
            GemmLaunch g{};
            g.A = h->hmid; g.a_rows = T; g.lda = I;
            g.B = (const __half*)h->w.wdown + (size_t)l * H * I; g.b_rows = H; g.ldb = I;
            g.K = I; g.block_n = 256; g.split_k = 1;
            g.epi.out = h->x; g.epi.ldo = H; g.epi.residual = h->x; g.epi.ldr = H; g.epi.T = (int)T; g.epi.F = H;
            if ((st = gemm_launch(g, s))) return (ctp_status)st;
        }
    }
    // last position of every sequence -> final norm -> heads; the decode residual stream continues in x[0..B)
    k_gather_last<<<B, 256, 0, s>>>(h->x, h->x_last, L0, H);
    LAUNCH_OK();
    CTP_CUDA_OK(cudaMemcpyAsync(h->x, h->x_last, sizeof(float) * B * H, cudaMemcpyDeviceToDevice, s));
    NormArgs nf{};
    nf.x = h->x; nf.w = h->w.norm_f; nf.xn = h->xn; nf.out_f32 = h->hidden; nf.H = H; nf.eps = c.rms_eps;
    nf.zero_buf = h->logits; nf.zero_n = infer_text ? c.num_text : c.num_vq * c.num_audio; nf.st = h->st; nf.write_hid = 1;
    k_rmsnorm<<<B, 256, 0, s>>>(nf);
    LAUNCH_OK();
    if ((st = launch_heads(h, B, s))) return (ctp_status)st;
    // the decode path accumulates into acc_qkv / acc_gu with fp32 atomics and their readers re-arm them: start from zero
    CTP_CUDA_OK(cudaMemsetAsync(h->acc_qkv, 0, sizeof(float) * (size_t)c.max_batch * 3 * H, s));
    CTP_CUDA_OK(cudaMemsetAsync(h->acc_gu, 0, sizeof(float) * (size_t)c.max_batch * 2 * I, s));
    CTP_CUDA_OK(cudaMemsetAsync(h->dec_gu, 0, sizeof(float) * (128 + (size_t)64 * 2 * I), s));
    h->B = B; h->cur_len = L0; h->prompt_len = L0; h->step = 0; h->max_new = bufs->max_new; h->have_bufs = true;
    return CTP_OK;
}

// device-side part of ctp_gpt_rewind: step 0, nothing finished (one thread)
__global__ void k_rewind_state(GenState* st, int B) {
    st->step = 0; st->all_done = 0; st->ticket = 0;
    for (int b = 0; b < B; ++b) { st->end_idx[b] = 0; st->finish[b] = 0; }
}

extern "C" ctp_status ctp_gpt_rewind(ctp_gpt* h, ctp_stream stream) {
    CTP_REQUIRE(h && h->bound && h->have_bufs, "rewind: call prefill first");
    CTP_REQUIRE(h->step <= 1 && h->cur_len == h->prompt_len, "rewind: a decode step has run since the prefill (step %d, cache length %d, prompt %d)",
                h->step, h->cur_len, h->prompt_len);
    k_rewind_state<<<1, 1, 0, (cudaStream_t)stream>>>(h->st, h->B);
    LAUNCH_OK();
    h->step = 0;
    return CTP_OK;
}

extern "C" ctp_status ctp_gpt_decode_step(ctp_gpt* h, const int32_t* ids, ctp_stream stream) {
    CTP_REQUIRE(h && h->bound && h->have_bufs, "decode_step: call prefill first");
    CTP_REQUIRE(h->cur_len + 1 <= h->cfg.max_seq, "decode_step: KV cache full (%d slots)", h->cfg.max_seq);
    CTP_REQUIRE(ids != nullptr || h->step >= 1, "decode_step: no sampled codes yet and no ids given");
    int st = run_decode_trunk(h, h->B, nsplit_for(h, h->B, h->cur_len + 1), ids, (cudaStream_t)stream);
    if (st) return (ctp_status)st;
    h->cur_len += 1;
    return CTP_OK;
}

static int check_sample_cfg(const ctp_gpt_cfg* c, const ctp_sample_cfg* cfg, int num_vq) {
    CTP_REQUIRE(cfg, "sample: null cfg");
    CTP_REQUIRE(cfg->top_k >= 0, "sample: top_k must be >= 0 (0 = no TopK warper; got %d)", cfg->top_k);
    CTP_REQUIRE(cfg->min_keep >= 1, "sample: min_keep must be >= 1");
    CTP_REQUIRE(cfg->rep_window >= 0 && cfg->rep_window <= 32, "sample: rep_window must be <= 32");
    CTP_REQUIRE(cfg->rep_penalty > 0.f, "sample: rep_penalty must be > 0");
    for (int q = 0; q < num_vq; ++q) CTP_REQUIRE(cfg->temperature[q] > 0.f, "sample: temperature[%d] must be > 0", q);
    (void)c;
    return CTP_OK;
}

static int upload_sample_state(ctp_gpt* h, const ctp_sample_cfg* cfg, const float* u, cudaStream_t s) {
    // cfg and u_base live inside the device GenState; update just those fields
    CTP_CUDA_OK(cudaMemcpyAsync(reinterpret_cast<char*>(h->st) + offsetof(GenState, cfg), cfg, sizeof(*cfg), cudaMemcpyHostToDevice, s));
    CTP_CUDA_OK(cudaMemcpyAsync(reinterpret_cast<char*>(h->st) + offsetof(GenState, u_base), &u, sizeof(u), cudaMemcpyHostToDevice, s));
    CTP_CUDA_OK(cudaStreamSynchronize(s));
    return CTP_OK;
}

extern "C" ctp_status ctp_gpt_sample_step(ctp_gpt* h, const ctp_sample_cfg* cfg, const float* u, ctp_stream stream) {
    CTP_REQUIRE(h && h->bound && h->have_bufs, "sample_step: call prefill first");
    CTP_REQUIRE(h->step < h->max_new, "sample_step: generation buffers full (%d steps)", h->max_new);
    int st = check_sample_cfg(&h->cfg, cfg, h->cfg.num_vq);
    if (st) return (ctp_status)st;
    cudaStream_t s = (cudaStream_t)stream;
    // single-step mode: u is [B*num_vq] for this step -> bias the base so that base + step*rows == u
    const float* ubase = u ? u - (long long)h->step * h->B * (h->text_mode ? 1 : h->cfg.num_vq) : nullptr;
    if ((st = upload_sample_state(h, cfg, ubase, s))) return (ctp_status)st;
    if ((st = launch_sampler(h, h->B, s))) return (ctp_status)st;
    h->step += 1;
    return CTP_OK;
}

static int get_graph(ctp_gpt* h, int B, int nsplit, cudaGraphExec_t* out) {
    GraphKey key{B, nsplit, h->text_mode};
    auto it = h->graphs.find(key);
    if (it != h->graphs.end()) { *out = it->second; return CTP_OK; }
    cudaGraph_t graph = nullptr;
    CTP_CUDA_OK(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
    ctp_count_capture_begin();
    int st = run_decode_trunk(h, B, nsplit, nullptr, h->cap_stream, true);
    if (!st) st = launch_sampler(h, B, h->cap_stream, h->use_pdl, true);
    const long long n_nodes = ctp_count_capture_end();
    cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
    if (st) { if (graph) cudaGraphDestroy(graph); return st; }
    if (e != cudaSuccess) { ctp_set_error("graph capture failed: %s", cudaGetErrorString(e)); return CTP_ERR_CUDA; }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { ctp_set_error("graph instantiate failed: %s", cudaGetErrorString(e)); return CTP_ERR_CUDA; }
    h->graphs[key] = exec;
    h->graph_nodes[key] = n_nodes;
    *out = exec;
    return CTP_OK;
}

extern "C" ctp_status ctp_gpt_generate(ctp_gpt* h, const ctp_sample_cfg* cfg, int32_t max_steps, const float* u,
                                       int32_t check_every, int32_t* steps_done, ctp_stream stream) {
    CTP_REQUIRE(h && h->bound && h->have_bufs, "generate: call prefill first");
    CTP_REQUIRE(h->step >= 1, "generate: sample the first step (ctp_gpt_sample_step) before the loop");
    int st = check_sample_cfg(&h->cfg, cfg, h->cfg.num_vq);
    if (st) return (ctp_status)st;
    cudaStream_t s = (cudaStream_t)stream;
    if ((st = upload_sample_state(h, cfg, u, s))) return (ctp_status)st;
    if (check_every < 1) check_every = 16;
    int done = 0;
    int iters = max_steps;
    iters = std::min(iters, h->max_new - h->step);
    iters = std::min(iters, h->cfg.max_seq - h->cur_len);
    cudaEvent_t ev = h->poll_ev;
    bool pending = false;
    h->st_pin->all_done = 0;
    for (int it = 0; it < iters; ++it) {
        {
            cudaGraphExec_t g;
            if ((st = get_graph(h, h->B, nsplit_for(h, h->B, h->cur_len + 1), &g))) return (ctp_status)st;
            ctp_count_launch((int)h->graph_nodes[GraphKey{h->B, nsplit_for(h, h->B, h->cur_len + 1), h->text_mode}]);
            cudaError_t e = cudaGraphLaunch(g, s);
            if (e != cudaSuccess) { ctp_set_error("graph launch: %s", cudaGetErrorString(e)); return CTP_ERR_CUDA; }
        }
        h->cur_len += 1; h->step += 1; done += 1;
        if ((it + 1) % check_every == 0 && it + 1 < iters) {
            // lagged poll: look at the flag copied after the PREVIOUS chunk while this chunk is already queued
            if (pending) {
                cudaEventSynchronize(ev);
                if (h->st_pin->all_done) break;
            }
            cudaMemcpyAsync(&h->st_pin->all_done, reinterpret_cast<char*>(h->st) + offsetof(GenState, all_done), sizeof(int),
                            cudaMemcpyDeviceToHost, s);
            cudaEventRecord(ev, s);
            pending = true;
        }
    }
    CTP_CUDA_OK(cudaGetLastError());
    if (steps_done) *steps_done = done;
    return CTP_OK;
}

// bring-up hook (not in include/ctp.h): copies the timeline records (8 x u64 per kernel, launch order) to `out`; returns the count
extern "C" __attribute__((visibility("default"))) int ctp_debug_trace(ctp_gpt* h, unsigned long long* out, int max_rec) {
    if (!h || !h->trace) return 0;
    const int n = h->trace_n < max_rec ? h->trace_n : max_rec;
    cudaDeviceSynchronize();
    cudaMemcpy(out, h->trace, sizeof(unsigned long long) * 8 * (size_t)n, cudaMemcpyDeviceToHost);
    return n;
}

// Bring-up / unit-test entry (not part of include/ctp.h): the fused MLP kernel of layer `layer` alone.  x_in, x_out: device fp32
// [>= 32][H]; x_out is cleared here.  Returns CTP_ERR_INVALID when the handle's shape is not eligible for the kernel.
extern "C" __attribute__((visibility("default"))) int ctp_debug_mlp(ctp_gpt* h, int layer, const float* x_in, float* x_out, int T, void* stream) {
    if (!h || !h->bound || !h->mlp_cluster || layer < 0 || layer >= h->cfg.n_layers || T < 1 || T > 32) { ctp_set_error("ctp_debug_mlp: not eligible"); return CTP_ERR_INVALID; }
    const ctp_gpt_cfg& c = h->cfg;
    const int H = c.hidden, I = c.inter;
    cudaStream_t s = (cudaStream_t)stream;
    CTP_CUDA_OK(cudaMemsetAsync(x_out, 0, sizeof(float) * 32 * H, s));
    MlpArgs ma{};
    ma.x_in = x_in; ma.x_out = x_out; ma.ln_w = h->w.ln2 + (size_t)layer * H; ma.T = T; ma.H = H; ma.I = I; ma.eps = c.rms_eps;
    ma.m64 = h->mlp_m64;
    const LayerMaps& m = h->lmaps[layer];
    cudaError_t le = launch_kc(k_mlp_fused, dim3((unsigned)(I / MLP_UW)), dim3(MLP_THREADS), (size_t)mlp_smem_bytes(H), s, false, (unsigned)h->mlp_cluster,
                               m.mlp_gu, m.mlp_d, ma);
    ctp_count_launch();
    if (le == cudaSuccess) le = cudaGetLastError();
    if (le != cudaSuccess) { ctp_set_error("%s:%d launch k_mlp_fused: %s", __FILE__, __LINE__, cudaGetErrorString(le)); return CTP_ERR_CUDA; }
    return CTP_OK;
}

extern "C" const float* ctp_gpt_logits(ctp_gpt* h) { return h ? h->logits : nullptr; }
extern "C" const float* ctp_gpt_hidden(ctp_gpt* h) { return h ? h->hidden : nullptr; }
extern "C" ctp_status ctp_gpt_copy_outputs(ctp_gpt* h, float* logits_out, float* hidden_out, ctp_stream stream) {
    CTP_REQUIRE(h && h->have_bufs, "copy_outputs: call prefill first");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t nl = (size_t)h->B * (h->text_mode ? h->cfg.num_text : h->cfg.num_vq * h->cfg.num_audio);
    if (logits_out) CTP_CUDA_OK(cudaMemcpyAsync(logits_out, h->logits, nl * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (hidden_out) CTP_CUDA_OK(cudaMemcpyAsync(hidden_out, h->hidden, (size_t)h->B * h->cfg.hidden * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return CTP_OK;
}
extern "C" ctp_status ctp_gpt_state(ctp_gpt* h, int32_t* cur_len, int32_t* step) {
    CTP_REQUIRE(h, "state: null handle");
    if (cur_len) *cur_len = h->cur_len;
    if (step) *step = h->step;
    return CTP_OK;
}
extern "C" const void* ctp_gpt_kv_plane(ctp_gpt* h, int32_t layer, int32_t which) {
    if (!h || layer < 0 || layer >= h->cfg.n_layers) return nullptr;
    return which ? (const void*)h->vplane(layer) : (const void*)h->kplane(layer);
}

extern "C" ctp_status ctp_sample(int32_t rows, int32_t vocab, int32_t num_vq, const float* logits, const int32_t* history,
                                 int32_t hist_len, int32_t hist_stride, const ctp_sample_cfg* cfg, int32_t step, const float* u,
                                 int32_t* next_ids, float* probs_out, ctp_stream stream) {
    CTP_REQUIRE(rows >= 1 && vocab >= 1 && num_vq >= 1 && num_vq <= MAX_VQ && rows % num_vq == 0, "sample: bad shape");
    CTP_REQUIRE(logits && (next_ids || probs_out), "sample: null pointer");
    CTP_REQUIRE(hist_len == 0 || history, "sample: history missing");
    int st = check_sample_cfg(nullptr, cfg, num_vq);
    if (st) return (ctp_status)st;
    SampleArgs sa{};
    sa.logits = logits; sa.vocab = vocab; sa.num_vq = num_vq; sa.rows = rows; sa.hist = history; sa.hist_stride = hist_stride;
    sa.hist_len = hist_len; sa.u = u; sa.step = step; sa.cfg = *cfg; sa.next_ids = next_ids; sa.probs_out = probs_out; sa.st = nullptr;
    sa.ids_cols = num_vq;
    const size_t smem = sizeof(float) * num_vq * ((vocab + 31) & ~31);
    CTP_REQUIRE(smem <= 100 * 1024, "sample: vocab %d too large for the warp sampler", vocab);
    CTP_CUDA_OK(cudaFuncSetAttribute(k_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    k_sample<<<rows / num_vq, 32 * num_vq, smem, (cudaStream_t)stream>>>(sa);
    ctp_count_launch();
    CTP_CUDA_OK(cudaGetLastError());
    return CTP_OK;
}
