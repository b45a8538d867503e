// Vocoder entry points (placeholder until the DVAE/Vocos kernels land; every call fails loudly).
#include "../../include/ctp.h"
#include "ctp_common.cuh"
extern "C" ctp_status ctp_voc_create(ctp_voc** out, const ctp_voc_cfg* cfg) { ctp_set_error("vocoder kernels not built yet"); return CTP_ERR_UNSUPPORTED; }
extern "C" void ctp_voc_destroy(ctp_voc* h) {}
extern "C" ctp_status ctp_voc_bind_weights(ctp_voc* h, const ctp_voc_weights* w) { ctp_set_error("vocoder kernels not built yet"); return CTP_ERR_UNSUPPORTED; }
extern "C" ctp_status ctp_voc_decode(ctp_voc* h, int32_t n_utt, const int32_t* lens_host, const void* src, float* wav_out,
                                     const int64_t* wav_offsets_host, float* mel_out, ctp_stream stream) { ctp_set_error("vocoder kernels not built yet"); return CTP_ERR_UNSUPPORTED; }
