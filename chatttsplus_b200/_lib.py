"""ctypes binding of libctp.so (include/ctp.h).  This is the stub a maintainer of the reference would add
(see INTEGRATION.md): plain pointers and sizes cross the boundary, torch only supplies device memory
(``tensor.data_ptr()``) and the current stream (``torch.cuda.current_stream().cuda_stream``, exactly what the
reference's TensorRT plugin passes: chattts_plus/trt_models/llama_trt_model.py:18, predictor.py:165).

There is no CPU fallback: a missing library or a failing call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libctp.so")

CTP_OK = 0


class CtpError(RuntimeError):
    pass


class GptCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("n_layers", "hidden", "n_heads", "inter", "num_vq", "num_audio", "num_text", "max_batch", "max_seq")] + \
               [("rms_eps", C.c_float), ("rope_theta", C.c_float)]


class GptWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("wqkv", "wo", "wgu", "wdown", "ln1", "ln2", "norm_f", "emb_code", "head_code", "emb_text", "head_text")]


class SampleCfg(C.Structure):
    _fields_ = [("temperature", C.c_float * 8), ("rep_penalty", C.c_float), ("rep_window", C.c_int32),
                ("rep_max_ids", C.c_int32), ("top_p", C.c_float), ("top_k", C.c_int32), ("min_keep", C.c_int32),
                ("eos", C.c_int32), ("min_new", C.c_int32), ("seed", C.c_uint64)]


class GenBuffers(C.Structure):
    _fields_ = [("ids", C.c_void_p), ("hiddens", C.c_void_p), ("end_idx", C.c_void_p), ("finish", C.c_void_p),
                ("max_new", C.c_int32)]


class VocCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("dvae_idim", "dvae_bn", "dvae_hidden", "dvae_layers", "dvae_odim", "dvae_dilation", "n_mels", "use_vq",
                 "voc_dim", "voc_inter", "voc_layers", "n_fft", "hop", "max_frames", "encoder")]


class ConvNextW(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dw_w", "dw_b", "ln_w", "ln_b", "pw1_w", "pw1_b", "pw2_w", "pw2_b", "gamma")]


class VocWeights(C.Structure):
    _fields_ = [("conv_in0_w", C.c_void_p), ("conv_in0_b", C.c_void_p), ("conv_in2_w", C.c_void_p),
                ("conv_in2_b", C.c_void_p), ("dvae_blocks", C.POINTER(ConvNextW)), ("conv_out_w", C.c_void_p),
                ("out_conv_w", C.c_void_p), ("coef", C.c_void_p), ("vq_proj_w", C.c_void_p), ("vq_proj_b", C.c_void_p),
                ("embed_w", C.c_void_p), ("embed_b", C.c_void_p), ("norm_w", C.c_void_p), ("norm_b", C.c_void_p),
                ("voc_blocks", C.POINTER(ConvNextW)), ("final_ln_w", C.c_void_p), ("final_ln_b", C.c_void_p),
                ("head_w", C.c_void_p), ("head_b", C.c_void_p), ("window", C.c_void_p),
                ("ds0_w", C.c_void_p), ("ds0_b", C.c_void_p), ("ds2_w", C.c_void_p), ("ds2_b", C.c_void_p),
                ("mel_fb", C.c_void_p), ("vq_in_w", C.c_void_p), ("vq_in_b", C.c_void_p)]


# every symbol include/ctp.h declares: (restype, argtypes)
_VP, _I32, _I64 = C.c_void_p, C.c_int32, C.c_int64
SYMBOLS = {
    "ctp_last_error": (C.c_char_p, []),
    "ctp_version": (C.c_int, []),
    "ctp_device_check": (C.c_int, [C.c_int]),
    "ctp_launch_count": (C.c_longlong, [C.c_int]),
    "ctp_gpt_create": (C.c_int, [C.POINTER(_VP), C.POINTER(GptCfg)]),
    "ctp_gpt_destroy": (None, [_VP]),
    "ctp_gpt_bind_weights": (C.c_int, [_VP, C.POINTER(GptWeights)]),
    "ctp_gpt_embed_prompt": (C.c_int, [_VP, _I32, _I32, _VP, _VP, _VP, _VP]),
    "ctp_gpt_prefill": (C.c_int, [_VP, _I32, _I32, _VP, C.POINTER(_I32), C.POINTER(GenBuffers), _I32, _VP]),
    "ctp_gpt_rewind": (C.c_int, [_VP, _VP]),
    "ctp_gpt_decode_step": (C.c_int, [_VP, _VP, _VP]),
    "ctp_gpt_sample_step": (C.c_int, [_VP, C.POINTER(SampleCfg), _VP, _VP]),
    "ctp_gpt_generate": (C.c_int, [_VP, C.POINTER(SampleCfg), _I32, _VP, _I32, C.POINTER(_I32), _VP]),
    "ctp_gpt_logits": (_VP, [_VP]),
    "ctp_gpt_hidden": (_VP, [_VP]),
    "ctp_gpt_copy_outputs": (C.c_int, [_VP, _VP, _VP, _VP]),
    "ctp_gpt_state": (C.c_int, [_VP, C.POINTER(_I32), C.POINTER(_I32)]),
    "ctp_gpt_kv_plane": (_VP, [_VP, _I32, _I32]),
    "ctp_sample": (C.c_int, [_I32, _I32, _I32, _VP, _VP, _I32, _I32, C.POINTER(SampleCfg), _I32, _VP, _VP, _VP, _VP]),
    "ctp_voc_create": (C.c_int, [C.POINTER(_VP), C.POINTER(VocCfg)]),
    "ctp_voc_destroy": (None, [_VP]),
    "ctp_voc_bind_weights": (C.c_int, [_VP, C.POINTER(VocWeights)]),
    "ctp_voc_decode": (C.c_int, [_VP, _I32, C.POINTER(_I32), _VP, _VP, C.POINTER(_I64), _VP, _VP]),
    "ctp_voc_decode_mel": (C.c_int, [_VP, _I32, C.POINTER(_I32), _VP, _VP, C.POINTER(_I64), _VP]),
    "ctp_voc_encode": (C.c_int, [_VP, _I32, _VP, _VP, _VP, C.POINTER(_I32), _VP]),
    "ctp_voc_quantize": (C.c_int, [_VP, _I32, _VP, _VP, _VP]),
    "ctp_gemm_f16": (C.c_int, [_I32, _I32, _I32, _VP, _I64, _VP, _I64, _VP, _I64, _VP, _I32, _I32, _I32, _VP]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libctp.so (built by ``python -m chatttsplus_b200.build`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CtpError(f"{LIB_PATH} is missing: build it with `python -m chatttsplus_b200.build` "
                           "(the B200 path has no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)  # raises AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status: int, what: str = "") -> None:
    if status != CTP_OK:
        msg = lib().ctp_last_error()
        raise CtpError(f"{what} failed (status {status}): {msg.decode() if msg else ''}")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device/host pointer of a contiguous torch tensor (or None)."""
    if t is None:
        return C.c_void_p(0)
    assert t.is_contiguous(), "non-contiguous tensor passed to libctp"
    return C.c_void_p(t.data_ptr())
