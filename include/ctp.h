/*
 * libctp — C ABI of the B200-native ChatTTSPlus generation hot path (GPT decode loop + DVAE/Vocos vocoder).
 *
 * Boundary contract: plain pointers and sizes, no torch types; every device buffer passed in
 * is owned by the caller and only borrowed for the duration of the call (or, for bound weights / generation
 * buffers, until the next bind / prefill); work is issued on the caller's cudaStream_t (passed as void*);
 * no internal threads; one handle per device; every entry returns ctp_status and records a thread-local
 * message readable with ctp_last_error().
 *
 * Each entry cites the reference interface (path:line under the ChatTTSPlus checkout) that it replaces.
 * The reference-side binding (ctypes) a maintainer would add is shown in INTEGRATION.md and implemented in
 * chatttsplus_b200/_lib.py.
 */
#ifndef CTP_H_
#define CTP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define CTP_API __declspec(dllexport)
#else
#define CTP_API __attribute__((visibility("default")))
#endif

typedef enum ctp_status {
    CTP_OK = 0,
    CTP_ERR_INVALID = 1,     /* bad argument / shape / state */
    CTP_ERR_CUDA = 2,        /* a CUDA runtime or driver call failed (message has the detail) */
    CTP_ERR_UNSUPPORTED = 3, /* configuration outside what the kernels are built for */
    CTP_ERR_NO_DEVICE = 4    /* no sm_100 device / driver: the product path never falls back to the CPU */
} ctp_status;

typedef void* ctp_stream; /* cudaStream_t */

CTP_API const char* ctp_last_error(void);
CTP_API int ctp_version(void);
/* CTP_OK iff device `dev` exists and is compute capability 10.x. */
CTP_API ctp_status ctp_device_check(int dev);
/* Number of libctp kernels launched by this process so far (graph replays count their kernel nodes); reset != 0 zeroes it. */
CTP_API long long ctp_launch_count(int reset);

/* ======================================================================================================
 * GPT decoder (reference: chattts_plus/models/gpt.py GPT, chattts_plus/models/llama.py LlamaModel;
 * functional precedent for a native trunk plugin: chattts_plus/trt_models/llama_trt_model.py:25-81)
 * ====================================================================================================*/
typedef struct ctp_gpt ctp_gpt;

typedef struct ctp_gpt_cfg {
    int32_t n_layers;   /* 20  (configs/infer/chattts_plus.yaml:67-82) */
    int32_t hidden;     /* 768 */
    int32_t n_heads;    /* 12; head_dim is fixed at 64 */
    int32_t inter;      /* 3072 */
    int32_t num_vq;     /* 4 */
    int32_t num_audio;  /* 626 */
    int32_t num_text;   /* 21178 */
    int32_t max_batch;  /* KV cache rows (<= 64) */
    int32_t max_seq;    /* KV cache slots per sequence (prompt + generated) */
    float rms_eps;      /* 1e-6 */
    float rope_theta;   /* 10000 */
} ctp_gpt_cfg;

/* Device pointers to packed 16-bit weights; [out,in] row-major exactly as the reference's nn.Linear.weight.
 * Replaces GPT.from_pretrained / load_state_dict (gpt.py:84-85); weight_norm heads (gpt.py:57-77) arrive folded
 * (W = g*v/||v||); LoRA (chattts_plus_pipeline.py:420-434) arrives merged into wqkv / wo — call bind again to swap. */
typedef struct ctp_gpt_weights {
    const void* wqkv;      /* fp16 [L][3H][H]   rows: q_proj, k_proj, v_proj */
    const void* wo;        /* fp16 [L][H][H] */
    const void* wgu;       /* fp16 [L][2I][H]   rows: gate_proj then up_proj */
    const void* wdown;     /* fp16 [L][H][I] */
    const float* ln1;      /* fp32 [L][H] input_layernorm.weight */
    const float* ln2;      /* fp32 [L][H] post_attention_layernorm.weight */
    const float* norm_f;   /* fp32 [H]    gpt.norm.weight */
    const void* emb_code;  /* fp16 [num_vq][num_audio][H] */
    const void* head_code; /* fp16 [num_vq*num_audio][H] */
    const void* emb_text;  /* fp16 [num_text][H] */
    const void* head_text; /* fp16 [num_text][H] or NULL (refine-text pass) */
} ctp_gpt_weights;

/* Sampling parameters = gpt.py:346-351,469-481 + processors.py:6-57 (temperature per VQ head, windowed
 * repetition penalty, TopP then TopK with min_tokens_to_keep, min-length EOS ban, multinomial draw). */
typedef struct ctp_sample_cfg {
    float temperature[8]; /* per VQ head */
    float rep_penalty;    /* 1.0 disables */
    int32_t rep_window;   /* 16 */
    int32_t rep_max_ids;  /* processors.py:24-27 row-truncation quirk: rows >= this get no penalty */
    float top_p;          /* <= 0 disables */
    int32_t top_k;        /* <= 0 disables */
    int32_t min_keep;     /* 3 */
    int32_t eos;          /* 625 */
    int32_t min_new;      /* steps < min_new cannot emit eos */
    uint64_t seed;        /* Philox seed when no uniforms are supplied */
} ctp_sample_cfg;

/* Caller-owned generation buffers (replace inputs_ids_buf / hiddens / end_idx / finish of gpt.py:339-378). */
typedef struct ctp_gen_buffers {
    int32_t* ids;      /* dev [B][max_new][num_vq] sampled codes */
    float* hiddens;    /* dev [B][max_new][H] post-final-norm hidden per step (gpt.py:422-423) or NULL */
    int32_t* end_idx;  /* dev [B] number of frames before EOS */
    uint8_t* finish;   /* dev [B] */
    int32_t max_new;
} ctp_gen_buffers;

CTP_API ctp_status ctp_gpt_create(ctp_gpt** out, const ctp_gpt_cfg* cfg);
CTP_API void ctp_gpt_destroy(ctp_gpt* h);
CTP_API ctp_status ctp_gpt_bind_weights(ctp_gpt* h, const ctp_gpt_weights* w);

/* GPT.forward "get_emb" (gpt.py:125-149): ids int32 [B,L0,num_vq], text_mask u8 [B,L0] -> emb fp32 [B,L0,H]. */
CTP_API ctp_status ctp_gpt_embed_prompt(ctp_gpt* h, int32_t B, int32_t L0, const int32_t* ids, const uint8_t* text_mask,
                                        float* emb_out, ctp_stream stream);

/* First iteration of GPT.generate (gpt.py:389-457 with i == 0): runs the trunk over the prompt (left-padded:
 * pad_len[b] leading masked slots, host array), fills the KV cache, leaves hidden/logits of the last position
 * in the handle, resets the generation state and attaches `bufs`. */
CTP_API ctp_status ctp_gpt_prefill(ctp_gpt* h, int32_t B, int32_t L0, const float* emb, const int32_t* pad_len_host,
                                   const ctp_gen_buffers* bufs, int32_t infer_text, ctp_stream stream);

/* gpt.py:496-525 (a sequence ended on the very first sample: "regenerate in order to ensure non-empty"): put the generation
 * state back to "prefill done, nothing sampled" — step 0, end_idx / finish cleared.  The prompt's KV cache and the logits of
 * its last position are untouched, so the retry costs one sampler launch instead of a second prefill.  Only valid while no
 * decode step has run since the prefill. */
CTP_API ctp_status ctp_gpt_rewind(ctp_gpt* h, ctp_stream stream);

/* One trunk step on `ids` (dev int32 [B,num_vq]; NULL = the codes written by the last ctp_gpt_sample_step):
 * code embedding sum (gpt.py:398-407) -> 20 decoder layers with KV append (llama.py:689-749) -> final norm ->
 * heads (gpt.py:424-439).  This is LlamaTRTModel.predict's role (llama_trt_model.py:43-81). */
CTP_API ctp_status ctp_gpt_decode_step(ctp_gpt* h, const int32_t* ids, ctp_stream stream);

/* gpt.py:469-494,527-532 for the current step: processors, warpers, draw, EOS / finish bookkeeping, write ids.
 * u: dev fp32 [B*num_vq] uniforms in [0,1) or NULL (Philox).  Advances the step counter. */
CTP_API ctp_status ctp_gpt_sample_step(ctp_gpt* h, const ctp_sample_cfg* cfg, const float* u, ctp_stream stream);

/* Whole loop gpt.py:389-549 after the first sample: repeats decode_step + sample_step as one CUDA graph per step
 * until max steps or all finished (finish flags polled every `check_every` steps; no per-step host sync).
 * u: dev fp32 [max_new][B*num_vq] or NULL.  steps_done (host) receives the number of loop iterations executed. */
CTP_API ctp_status ctp_gpt_generate(ctp_gpt* h, const ctp_sample_cfg* cfg, int32_t max_steps, const float* u,
                                    int32_t check_every, int32_t* steps_done, ctp_stream stream);

/* Device views of the handle's current step outputs: logits fp32 [B*num_vq][num_audio] (row = b*num_vq+q, the
 * layout gpt.py:444-457 builds) or [B][num_text] in infer_text mode; hidden fp32 [B][H]. */
CTP_API const float* ctp_gpt_logits(ctp_gpt* h);
CTP_API const float* ctp_gpt_hidden(ctp_gpt* h);
/* Stream-ordered copies of the above into caller buffers (either may be NULL); B rows of the live batch. */
CTP_API ctp_status ctp_gpt_copy_outputs(ctp_gpt* h, float* logits_out, float* hidden_out, ctp_stream stream);
/* Current cache length (prompt + decoded so far) and number of sample steps taken. */
CTP_API ctp_status ctp_gpt_state(ctp_gpt* h, int32_t* cur_len, int32_t* step);
/* Debug/teaching hook: KV cache plane of one layer, fp16 [max_batch][n_heads][max_seq][64]; which = 0 K, 1 V. */
CTP_API const void* ctp_gpt_kv_plane(ctp_gpt* h, int32_t layer, int32_t which);

/* Stand-alone sampler over arbitrary logits (tests / other callers): logits fp32 [rows][vocab] (not modified),
 * history int32 [rows][hist_len] (last rep_window entries are used), row r uses temperature[r % num_vq].
 * Writes next_ids int32 [rows] and, if probs_out != NULL, the processed probabilities fp32 [rows][vocab]. */
CTP_API ctp_status ctp_sample(int32_t rows, int32_t vocab, int32_t num_vq, const float* logits, const int32_t* history,
                              int32_t hist_len, int32_t hist_stride, const ctp_sample_cfg* cfg, int32_t step,
                              const float* u, int32_t* next_ids, float* probs_out, ctp_stream stream);

/* ======================================================================================================
 * Vocoder: DVAE decode + Vocos (reference: chattts_plus/models/dvae.py:254-291, pip `vocos` Vocos.decode,
 * called per utterance by chattts_plus/pipelines/chattts_plus_pipeline.py:286-305)
 * ====================================================================================================*/
typedef struct ctp_voc ctp_voc;

typedef struct ctp_voc_cfg {
    /* DVAE decoder (configs/infer/chattts_plus.yaml:7-45) */
    int32_t dvae_idim;    /* 384 (Decoder.pt) / 512 (DVAE_full.pt) */
    int32_t dvae_bn;      /* 128 */
    int32_t dvae_hidden;  /* 512 / 256 */
    int32_t dvae_layers;  /* 12 */
    int32_t dvae_odim;    /* 384 / 512 */
    int32_t dvae_dilation;/* 2 */
    int32_t n_mels;       /* 100 */
    int32_t use_vq;       /* 1: input is codes [n,4] through GFSQ embed (dvae.py:84-94) */
    /* Vocos (configs/infer/chattts_plus.yaml:46-66) */
    int32_t voc_dim;      /* 512 */
    int32_t voc_inter;    /* 1536 */
    int32_t voc_layers;   /* 8 */
    int32_t n_fft;        /* 1024 */
    int32_t hop;          /* 256 */
    int32_t max_frames;   /* workspace: max total mel frames (2 per code frame) per decode call, incl. padding */
    /* Zero-shot speaker-prompt ENCODER handle (DVAE.forward(mode="encode"), dvae.py:263-270): set encoder = 1 and describe the
     * `encoder_config` stack with the dvae_* fields above (idim = dim of downsample_conv 512, bn 128, hidden 256, layers 12,
     * odim = vq dim 1024, configs/infer/chattts_plus.yaml:13); such a handle serves ctp_voc_encode only. */
    int32_t encoder;
} ctp_voc_cfg;

/* One ConvNeXt block (dvae.py:16-63 / vocos ConvNeXtBlock). */
typedef struct ctp_convnext_w {
    const float* dw_w;   /* fp32 [C][7] depthwise taps */
    const float* dw_b;   /* fp32 [C] */
    const float* ln_w;   /* fp32 [C] */
    const float* ln_b;   /* fp32 [C] */
    const void* pw1_w;   /* fp16 [4C or inter][C] */
    const float* pw1_b;  /* fp32 */
    const void* pw2_w;   /* fp16 [C][inter] */
    const float* pw2_b;  /* fp32 [C] */
    const float* gamma;  /* fp32 [C] */
} ctp_convnext_w;

typedef struct ctp_voc_weights {
    /* DVAE */
    const void* conv_in0_w;  /* fp16 [bn][3*idim]  im2col order: k = tap*idim + c */
    const float* conv_in0_b; /* fp32 [bn] */
    const void* conv_in2_w;  /* fp16 [hidden][3*bn] */
    const float* conv_in2_b; /* fp32 [hidden] */
    const ctp_convnext_w* dvae_blocks; /* host array [dvae_layers] of device pointers */
    const void* conv_out_w;  /* fp16 [odim][hidden] */
    const void* out_conv_w;  /* fp16 [n_mels][3*odim] */
    const float* coef;       /* fp32 [n_mels] */
    const float* vq_proj_w;  /* fp32 [G][dim/G][4]  GFSQ project_out (use_vq) or NULL */
    const float* vq_proj_b;  /* fp32 [G][dim/G] */
    /* Vocos */
    const void* embed_w;     /* fp16 [dim][7*mel_pad]  k = tap*mel_pad + c, mel_pad = 104 */
    const float* embed_b;    /* fp32 [dim] */
    const float* norm_w;     /* fp32 [dim] */
    const float* norm_b;
    const ctp_convnext_w* voc_blocks; /* host array [voc_layers] */
    const float* final_ln_w;
    const float* final_ln_b;
    const void* head_w;      /* fp16 [n_fft+2][dim] */
    const float* head_b;     /* fp32 [n_fft+2] */
    const float* window;     /* fp32 [n_fft]  (encoder handle: the analysis Hann window of MelSpectrogramFeatures, dvae.py:183-190) */
    /* prompt encoder (cfg.encoder = 1): conv_in0/2, dvae_blocks, conv_out above hold `encoder.*`; coef as above */
    const void* ds0_w;       /* fp16 [idim][3*mel_pad]  downsample_conv.0 (k3, p1), k = tap*mel_pad + c (dvae.py:227) */
    const float* ds0_b;      /* fp32 [idim] */
    const void* ds2_w;       /* fp16 [idim][4*idim]     downsample_conv.2 (k4, stride 2, p1), k = tap*idim + c (dvae.py:229) */
    const float* ds2_b;      /* fp32 [idim] */
    const float* mel_fb;     /* fp32 [n_fft/2+1][n_mels] mel filter bank (torchaudio melscale_fbanks, htk, norm=None) */
    const float* vq_in_w;    /* fp32 [G][4][odim/G]  GFSQ project_in (vector_quantize_pytorch ResidualFSQ) */
    const float* vq_in_b;    /* fp32 [G][4] */
} ctp_voc_weights;

CTP_API ctp_status ctp_voc_create(ctp_voc** out, const ctp_voc_cfg* cfg);
CTP_API void ctp_voc_destroy(ctp_voc* h);
CTP_API ctp_status ctp_voc_bind_weights(ctp_voc* h, const ctp_voc_weights* w); /* a DVAE-only or Vocos-only set may be bound (other pointers NULL) */

/* _decode_to_wavs for a whole batch (chattts_plus_pipeline.py:286-305): utterance i has lens_host[i] code frames;
 * src = concatenated hiddens fp32 [sum n_i][2*idim] (use_vq=0) or codes int32 [sum n_i][4] (use_vq=1);
 * wav_out = concatenated fp32 waveforms, utterance i at wav_offsets_host[i], length hop*(2*n_i-1).
 * mel_out (optional, may be NULL): concatenated fp32 mel [sum 2*n_i][n_mels] (the DVAE output, dvae.py:291). */
CTP_API ctp_status ctp_voc_decode(ctp_voc* h, int32_t n_utt, const int32_t* lens_host, const void* src, float* wav_out,
                                  const int64_t* wav_offsets_host, float* mel_out, ctp_stream stream);
/* wav_out may be NULL (DVAE only: DVAE.__call__(inp, "decode"), dvae.py:254-291, needs mel_out).
 *
 * vocos.Vocos.decode(mel) alone (chattts_plus_pipeline.py:303, tests/test_models.py:104-148): mel fp32
 * [sum T_i][n_mels] concatenated, utterance i has mel_lens_host[i] frames -> wav of hop*(T_i-1) samples. */
CTP_API ctp_status ctp_voc_decode_mel(ctp_voc* h, int32_t n_utt, const int32_t* mel_lens_host, const float* mel,
                                      float* wav_out, const int64_t* wav_offsets_host, ctp_stream stream);

/* DVAE.forward(inp, mode="encode") for ONE utterance (dvae.py:263-270; caller: ChatTTSPlusPipeline.sample_audio_speaker,
 * chattts_plus_pipeline.py:279-284): audio fp32 [n_samples] at 24 kHz -> log-mel (n_fft 1024, hop 256, 100 mels, centre
 * reflect padding) / coef -> downsample_conv -> encoder -> GFSQ indices.  ids_out: dev int32 [T2][G*R] with
 * T2 = ((n_samples / hop + 1) - 2) / 2 + 1 code frames (row t = the reference's ind[0, :, t]); feat_out (optional, may be
 * NULL): dev fp32 [T2][odim] encoder output before the quantiser; *n_frames_host receives T2. */
CTP_API ctp_status ctp_voc_encode(ctp_voc* h, int32_t n_samples, const float* audio, int32_t* ids_out, float* feat_out,
                                  int32_t* n_frames_host, ctp_stream stream);

/* GFSQ.forward alone (dvae.py:98-126 -> vector_quantize_pytorch GroupedResidualFSQ.forward): encoder features dev fp32
 * [n_frames][odim] (channels-last) -> indices dev int32 [n_frames][G*R].  Needs a prompt-encoder handle (project_in bound). */
CTP_API ctp_status ctp_voc_quantize(ctp_voc* h, int32_t n_frames, const float* feat, int32_t* ids_out, ctp_stream stream);

/* ======================================================================================================
 * Building block exposed for tests and profiling: C = epilogue(A[M,K] * B[N,K]^T), fp16 in, fp32 accumulate,
 * tcgen05 + TMA.  flags: bit0 = output fp16 (else fp32), bit1 = GELU, bit2 = atomic-add into fp32 out (split-K),
 * bit3 = swap (out[n*ldo + m], used by the decode path where the weight matrix is the M operand).
 * ====================================================================================================*/
CTP_API ctp_status ctp_gemm_f16(int32_t M, int32_t N, int32_t K, const void* A, int64_t lda, const void* B, int64_t ldb,
                                void* out, int64_t ldo, const float* bias, int32_t flags, int32_t block_n,
                                int32_t split_k, ctp_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* CTP_H_ */
