"""``DVAE`` / ``Vocos`` host-side mirrors and the fused batch vocoder engine (libctp ``ctp_voc_*``).

Reference interfaces kept:
  * ``DVAE(decoder_config, encoder_config=None, vq_config=None, dim=512, coef=None, model_path=...)``,
    ``DVAE.__call__(inp, mode="decode")`` -> mel ``[B, 100, 2T]``        (chattts_plus/models/dvae.py:203-291)
  * ``Vocos`` object with ``.parameters()`` (dtype probe) and ``.decode(mel)`` -> wav ``[B, 256*(T-1)]``
    (pip ``vocos`` as assembled by chattts_plus_pipeline.py:93-111 and called at :300-303)
  * ``VocoderEngine.decode_batch`` = ``ChatTTSPlusPipeline._decode_to_wavs`` for the whole batch in one library
    call (the reference loops utterance by utterance at batch 1, chattts_plus_pipeline.py:298-304).
  * ``DVAE.__call__(audio, mode="encode")`` -> GFSQ indices ``[B, 4, T2]`` (dvae.py:263-270): the zero-shot speaker prompt
    encoder (scope row f3): log-mel front end, downsample_conv, encoder stack and quantiser run in ``ctp_voc_encode``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from .commons import b14
from .commons import logger as _logger
from .synth import DVAEConfig, VocosConfig

MEL_PAD = 104


def _dev_f16(t, dev):
    return t.detach().to(dev, torch.float16).contiguous()


def _dev_f32(t, dev):
    return t.detach().to(dev, torch.float32).contiguous()


def _pack_convnext(sd: Dict[str, torch.Tensor], prefix: str, dev, keep: list) -> _lib.ConvNextW:
    t = dict(
        dw_w=_dev_f32(sd[prefix + "dwconv.weight"].squeeze(1), dev), dw_b=_dev_f32(sd[prefix + "dwconv.bias"], dev),
        ln_w=_dev_f32(sd[prefix + "norm.weight"], dev), ln_b=_dev_f32(sd[prefix + "norm.bias"], dev),
        pw1_w=_dev_f16(sd[prefix + "pwconv1.weight"], dev), pw1_b=_dev_f32(sd[prefix + "pwconv1.bias"], dev),
        pw2_w=_dev_f16(sd[prefix + "pwconv2.weight"], dev), pw2_b=_dev_f32(sd[prefix + "pwconv2.bias"], dev),
        gamma=_dev_f32(sd[prefix + "gamma"], dev))
    keep.append(t)
    return _lib.ConvNextW(**{k: v.data_ptr() for k, v in t.items()})


def _im2col_weight(w: torch.Tensor, c_pad: Optional[int] = None) -> torch.Tensor:
    """Conv1d weight [O, C, K] -> [O, K*C'] with k = tap*C' + c (the order the overlapping-row TMA window has)."""
    O, Cc, K = w.shape
    if c_pad is not None and c_pad > Cc:
        w = F.pad(w, (0, 0, 0, c_pad - Cc))
    return w.permute(0, 2, 1).reshape(O, -1)


class DVAE:
    def __init__(self, decoder_config: dict, encoder_config: Optional[dict] = None, vq_config: Optional[dict] = None,
                 dim=512, coef: Optional[str] = None, **kwargs):
        self.logger = _logger.get_logger(self.__class__.__name__)
        d = dict(decoder_config)
        self.cfg = DVAEConfig(dim=int(dim), idim=int(d["idim"]), odim=int(d["odim"]), hidden=int(d.get("hidden", 256)),
                              n_layer=int(d.get("n_layer", 12)), bn_dim=int(d.get("bn_dim", 64)),
                              kernel=int(d.get("kernel", 7)), dilation=int(d.get("dilation", 2)), vq=vq_config is not None)
        if vq_config is not None:
            v = dict(vq_config)
            self.cfg.vq_dim = int(v["dim"]); self.cfg.vq_levels = tuple(v["levels"]); self.cfg.vq_G = int(v["G"]); self.cfg.vq_R = int(v["R"])
            if tuple(self.cfg.vq_levels) != (5, 5, 5, 5) or self.cfg.vq_G != 2 or self.cfg.vq_R != 2:
                raise _lib.CtpError("the GFSQ embed kernel is built for levels [5,5,5,5], G=2, R=2")
        self.has_encoder = encoder_config is not None
        if self.has_encoder:
            e = dict(encoder_config)
            self.cfg.encoder = True
            self.cfg.enc_hidden = int(e.get("hidden", 256)); self.cfg.enc_layers = int(e.get("n_layer", 12)); self.cfg.enc_bn = int(e.get("bn_dim", 64))
            if int(e["idim"]) != self.cfg.dim:
                raise _lib.CtpError("encoder_config.idim must equal dim (the width of downsample_conv, dvae.py:224-233)")
            self.cfg.enc_odim = int(e["odim"])
        self._enc = None          # packed encoder weights
        self._enc_handle = C.c_void_p(0)
        self._enc_frames = 0
        if coef is None:
            coef_t = torch.rand(100)
        else:
            coef_t = torch.from_numpy(np.copy(np.frombuffer(b14.decode_from_string(coef), dtype=np.float32)))
        self.coef = coef_t.view(1, 100, 1)  # dvae.py:213-222 (overwritten by load_state_dict, as in the reference)
        self.device = torch.device("cpu")
        self._state: Optional[Dict[str, torch.Tensor]] = None
        self._keep: list = []
        self._w: Optional[dict] = None
        self._engine: Optional["VocoderEngine"] = None
        self.model_path = kwargs.get("model_path", None)
        if self.model_path:
            self.logger.info(f"loading DVAE pretrained model: {self.model_path}")
            self.from_pretrained(self.model_path)

    def __repr__(self) -> str:
        return b14.encode_to_string(self.coef.cpu().numpy().astype(np.float32).tobytes())

    def eval(self):
        return self

    def from_pretrained(self, file_path: str):
        self.load_state_dict(torch.load(file_path, weights_only=True, mmap=True))

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        need = ["coef", "decoder.conv_in.0.weight", "decoder.conv_in.0.bias", "decoder.conv_in.2.weight", "decoder.conv_in.2.bias",
                "decoder.conv_out.weight", "out_conv.weight"]
        for l in range(self.cfg.n_layer):
            for n in ("dwconv.weight", "dwconv.bias", "norm.weight", "norm.bias", "pwconv1.weight", "pwconv1.bias",
                      "pwconv2.weight", "pwconv2.bias", "gamma"):
                need.append(f"decoder.decoder_block.{l}.{n}")
        if self.cfg.vq:
            for g in range(self.cfg.vq_G):
                need += [f"vq_layer.quantizer.rvqs.{g}.project_out.weight", f"vq_layer.quantizer.rvqs.{g}.project_out.bias"]
        missing = [k for k in need if k not in sd]
        if missing:
            raise RuntimeError(f"Error(s) in loading state_dict for DVAE: missing {missing[:5]}")
        if self.has_encoder and "encoder.conv_out.weight" in sd:   # the encode side is optional in a state dict
            need_e = ["downsample_conv.0.weight", "downsample_conv.0.bias", "downsample_conv.2.weight", "downsample_conv.2.bias",
                      "encoder.conv_in.0.weight", "encoder.conv_in.0.bias", "encoder.conv_in.2.weight", "encoder.conv_in.2.bias"]
            need_e += [f"encoder.decoder_block.{l}.{n}" for l in range(self.cfg.enc_layers) for n in
                       ("dwconv.weight", "dwconv.bias", "norm.weight", "norm.bias", "pwconv1.weight", "pwconv1.bias", "pwconv2.weight", "pwconv2.bias", "gamma")]
            if self.cfg.vq:
                need_e += [f"vq_layer.quantizer.rvqs.{g}.project_in.{n}" for g in range(self.cfg.vq_G) for n in ("weight", "bias")]
            miss_e = [k for k in need_e if k not in sd]
            if miss_e:
                raise RuntimeError(f"Error(s) in loading state_dict for DVAE (encoder side): missing {miss_e[:5]}")
        self._state = {k: sd[k].detach().cpu() for k in sd}
        self.coef = self._state["coef"].float().view(1, 100, 1)
        if self.device.type == "cuda":
            self._pack()
        return self

    def to(self, device=None, dtype=None, **kw):
        if device is not None:
            device = torch.device(device)
            if device.type == "cuda":
                if not torch.cuda.is_available():
                    raise _lib.CtpError("chatttsplus_b200.DVAE needs a CUDA (sm_100a) device; there is no CPU path")
                if device.index is None:
                    device = torch.device("cuda", torch.cuda.current_device())
                self.device = device
                if self._state is not None:
                    self._pack()
        return self

    def _pack(self):
        sd, dev, c = self._state, self.device, self.cfg
        self._keep = []
        w = dict(
            conv_in0_w=_dev_f16(_im2col_weight(sd["decoder.conv_in.0.weight"].float()), dev), conv_in0_b=_dev_f32(sd["decoder.conv_in.0.bias"], dev),
            conv_in2_w=_dev_f16(_im2col_weight(sd["decoder.conv_in.2.weight"].float()), dev), conv_in2_b=_dev_f32(sd["decoder.conv_in.2.bias"], dev),
            conv_out_w=_dev_f16(sd["decoder.conv_out.weight"].float().squeeze(-1), dev),
            out_conv_w=_dev_f16(_im2col_weight(sd["out_conv.weight"].float()), dev),
            coef=_dev_f32(sd["coef"].reshape(-1), dev))
        if c.vq:
            w["vq_proj_w"] = _dev_f32(torch.stack([sd[f"vq_layer.quantizer.rvqs.{g}.project_out.weight"] for g in range(c.vq_G)]), dev)
            w["vq_proj_b"] = _dev_f32(torch.stack([sd[f"vq_layer.quantizer.rvqs.{g}.project_out.bias"] for g in range(c.vq_G)]), dev)
        blocks = (_lib.ConvNextW * c.n_layer)(*[_pack_convnext(sd, f"decoder.decoder_block.{l}.", dev, self._keep) for l in range(c.n_layer)])
        self._w = dict(tensors=w, blocks=blocks)
        self._engine = None
        self._enc = None
        if self._enc_handle:
            _lib.lib().ctp_voc_destroy(self._enc_handle)
            self._enc_handle = C.c_void_p(0)
        if self.has_encoder and c.vq and "encoder.conv_out.weight" in sd:
            self._pack_encoder()

    def _pack_encoder(self):
        """Prompt-encoder weight set of ``ctp_voc_encode`` (include/ctp.h): encoder.* in the DVAE stack slots, downsample_conv as
        im2col matrices, the torchaudio mel filter bank (HTK, norm=None) and analysis window, GFSQ project_in."""
        sd, dev, c = self._state, self.device, self.cfg
        n_fft, sr = 1024, 24000
        all_freqs = torch.linspace(0, sr // 2, n_fft // 2 + 1)
        m_pts = torch.linspace(0.0, 2595.0 * float(np.log10(1.0 + (sr / 2.0) / 700.0)), c.n_mels + 2)
        f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
        f_diff = f_pts[1:] - f_pts[:-1]
        slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
        fb = torch.clamp(torch.min(-slopes[:, :-2] / f_diff[:-1], slopes[:, 2:] / f_diff[1:]), min=0.0)   # torchaudio.functional.melscale_fbanks
        keep: list = []
        w = dict(
            conv_in0_w=_dev_f16(_im2col_weight(sd["encoder.conv_in.0.weight"].float()), dev), conv_in0_b=_dev_f32(sd["encoder.conv_in.0.bias"], dev),
            conv_in2_w=_dev_f16(_im2col_weight(sd["encoder.conv_in.2.weight"].float()), dev), conv_in2_b=_dev_f32(sd["encoder.conv_in.2.bias"], dev),
            conv_out_w=_dev_f16(sd["encoder.conv_out.weight"].float().squeeze(-1), dev),
            coef=_dev_f32(sd["coef"].reshape(-1), dev),
            ds0_w=_dev_f16(_im2col_weight(sd["downsample_conv.0.weight"].float(), MEL_PAD), dev), ds0_b=_dev_f32(sd["downsample_conv.0.bias"], dev),
            ds2_w=_dev_f16(_im2col_weight(sd["downsample_conv.2.weight"].float()), dev), ds2_b=_dev_f32(sd["downsample_conv.2.bias"], dev),
            mel_fb=_dev_f32(fb, dev), window=_dev_f32(torch.hann_window(n_fft, periodic=True), dev),
            vq_in_w=_dev_f32(torch.stack([sd[f"vq_layer.quantizer.rvqs.{g}.project_in.weight"] for g in range(c.vq_G)]), dev),
            vq_in_b=_dev_f32(torch.stack([sd[f"vq_layer.quantizer.rvqs.{g}.project_in.bias"] for g in range(c.vq_G)]), dev))
        blocks = (_lib.ConvNextW * c.enc_layers)(*[_pack_convnext(sd, f"encoder.decoder_block.{l}.", dev, keep) for l in range(c.enc_layers)])
        self._enc = dict(tensors=w, blocks=blocks, keep=keep)

    def _ensure_encoder(self, mel_frames: int):
        if self._enc_handle and mel_frames + 32 <= self._enc_frames:
            return
        lib = _lib.lib()
        if self._enc_handle:
            lib.ctp_voc_destroy(self._enc_handle)
            self._enc_handle = C.c_void_p(0)
        c = self.cfg
        self._enc_frames = max(4096, mel_frames + 32)
        cfg = _lib.VocCfg(dvae_idim=c.dim, dvae_bn=c.enc_bn, dvae_hidden=c.enc_hidden, dvae_layers=c.enc_layers, dvae_odim=c.enc_odim,
                          dvae_dilation=c.dilation, n_mels=c.n_mels, use_vq=0, voc_dim=512, voc_inter=1536, voc_layers=0, n_fft=1024, hop=256,
                          max_frames=self._enc_frames, encoder=1)
        h = C.c_void_p(0)
        with torch.cuda.device(self.device):
            _lib.check(lib.ctp_voc_create(C.byref(h), C.byref(cfg)), "ctp_voc_create (prompt encoder)")
            self._enc_handle = h
            w = _lib.VocWeights()
            for k, t in self._enc["tensors"].items():
                setattr(w, k, t.data_ptr())
            w.dvae_blocks = C.cast(self._enc["blocks"], C.POINTER(_lib.ConvNextW))
            _lib.check(lib.ctp_voc_bind_weights(self._enc_handle, C.byref(w)), "ctp_voc_bind_weights (prompt encoder)")

    @torch.inference_mode()
    def encode(self, audio: torch.Tensor, return_features: bool = False):
        """audio ``[B, N]`` (24 kHz) -> GFSQ indices ``[B, G*R, T2]`` int64 (dvae.py:263-270); utterances are encoded one by one."""
        if self._enc is None:
            raise _lib.CtpError("this DVAE has no prompt encoder on a CUDA device (needs encoder_config, vq_config and encoder.* weights)")
        if audio.dim() == 1:
            audio = audio[None]
        outs, feats = [], []
        with torch.cuda.device(self.device):
            for b in range(audio.shape[0]):
                a = audio[b].to(self.device, torch.float32).contiguous()
                n = int(a.numel())
                T = n // 256 + 1
                T2 = (T - 2) // 2 + 1
                self._ensure_encoder(T)
                ids = torch.empty(T2, self.cfg.vq_G * self.cfg.vq_R, device=self.device, dtype=torch.int32)
                feat = torch.empty(T2, self.cfg.enc_odim, device=self.device, dtype=torch.float32) if return_features else None
                nf = C.c_int32(0)
                _lib.check(_lib.lib().ctp_voc_encode(self._enc_handle, n, _lib.ptr(a), _lib.ptr(ids), _lib.ptr(feat), C.byref(nf), _lib.stream_ptr()),
                           "ctp_voc_encode")
                assert nf.value == T2
                outs.append(ids.permute(1, 0).long())
                if return_features:
                    feats.append(feat.permute(1, 0))
        ind = torch.stack(outs)
        return (ind, torch.stack(feats)) if return_features else ind

    @torch.inference_mode()
    def quantize_features(self, x: torch.Tensor) -> torch.Tensor:
        """GFSQ.forward alone (dvae.py:98-126): encoder features ``[B, odim, T]`` -> indices ``[B, G*R, T]`` int64."""
        if self._enc is None:
            raise _lib.CtpError("this DVAE has no prompt encoder on a CUDA device")
        outs = []
        with torch.cuda.device(self.device):
            for b in range(x.shape[0]):
                f = x[b].to(self.device, torch.float32).t().contiguous()          # [T, odim]
                self._ensure_encoder(64)
                ids = torch.empty(f.shape[0], self.cfg.vq_G * self.cfg.vq_R, device=self.device, dtype=torch.int32)
                _lib.check(_lib.lib().ctp_voc_quantize(self._enc_handle, f.shape[0], _lib.ptr(f), _lib.ptr(ids), _lib.stream_ptr()),
                           "ctp_voc_quantize")
                outs.append(ids.permute(1, 0).long())
        return torch.stack(outs)

    def __del__(self):
        try:
            if self._enc_handle:
                _lib.lib().ctp_voc_destroy(self._enc_handle)
        except Exception:
            pass

    def __call__(self, inp: torch.Tensor, mode: str = "decode") -> torch.Tensor:
        return self.forward(inp, mode)

    @torch.inference_mode()
    def forward(self, inp: torch.Tensor, mode: str = "decode") -> torch.Tensor:
        if mode == "encode" and self.has_encoder and self.cfg.vq:   # dvae.py:263
            return self.encode(inp)
        if self._w is None:
            raise _lib.CtpError("DVAE weights are not on a CUDA device: call .to('cuda') after loading")
        if self._engine is None:
            self._engine = VocoderEngine(self, None)
        B = inp.shape[0]
        if self.cfg.vq:   # codes [B, 4, T]
            items = [inp[b].permute(1, 0) for b in range(B)]
        else:             # hidden [B, 768, T]
            items = [inp[b].permute(1, 0) for b in range(B)]
        mels = self._engine.decode_batch(items, want_wav=False, want_mel=True)[1]
        return torch.stack([m.permute(1, 0) for m in mels])  # [B, 100, 2T]


class Vocos:
    """vocos.Vocos-compatible surface used by the pipeline: ``parameters()`` and ``decode(mel)``."""

    def __init__(self, feature_extractor_config: Optional[dict] = None, backbone_config: Optional[dict] = None,
                 head_config: Optional[dict] = None, **kwargs):
        self.logger = _logger.get_logger(self.__class__.__name__)
        b = dict(backbone_config or {})
        hc = dict(head_config or {})
        self.cfg = VocosConfig(input_channels=int(b.get("input_channels", 100)), dim=int(b.get("dim", 512)),
                               intermediate_dim=int(b.get("intermediate_dim", 1536)), num_layers=int(b.get("num_layers", 8)),
                               n_fft=int(hc.get("n_fft", 1024)), hop_length=int(hc.get("hop_length", 256)))
        if hc.get("padding", "center") != "center":
            raise _lib.CtpError("only ISTFT padding='center' is implemented (configs/infer/chattts_plus.yaml:62-66)")
        self.device = torch.device("cpu")
        self._state = None
        self._keep: list = []
        self._w = None
        self._engine = None
        self._dtype_probe = torch.zeros(1, dtype=torch.float32)
        self.model_path = kwargs.get("model_path", None)
        if self.model_path:
            self.load_state_dict(torch.load(self.model_path, weights_only=True, mmap=True))

    def eval(self):
        return self

    def parameters(self):
        return iter([self._dtype_probe])

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        need = ["backbone.embed.weight", "backbone.embed.bias", "backbone.norm.weight", "backbone.norm.bias",
                "backbone.final_layer_norm.weight", "backbone.final_layer_norm.bias", "head.out.weight", "head.out.bias"]
        missing = [k for k in need if k not in sd]
        if missing:
            raise RuntimeError(f"Error(s) in loading state_dict for Vocos: missing {missing}")
        self._state = {k: sd[k].detach().cpu() for k in sd}
        if "head.istft.window" not in self._state:
            self._state["head.istft.window"] = torch.hann_window(self.cfg.n_fft, periodic=True)
        if self.device.type == "cuda":
            self._pack()
        return self

    def to(self, device=None, dtype=None, **kw):
        if device is not None:
            device = torch.device(device)
            if device.type == "cuda":
                if not torch.cuda.is_available():
                    raise _lib.CtpError("chatttsplus_b200.Vocos needs a CUDA (sm_100a) device; there is no CPU path")
                if device.index is None:
                    device = torch.device("cuda", torch.cuda.current_device())
                self.device = device
                self._dtype_probe = torch.zeros(1, device=device, dtype=torch.float32)
                if self._state is not None:
                    self._pack()
        return self

    def _pack(self):
        sd, dev, c = self._state, self.device, self.cfg
        self._keep = []
        w = dict(
            embed_w=_dev_f16(_im2col_weight(sd["backbone.embed.weight"].float(), MEL_PAD), dev), embed_b=_dev_f32(sd["backbone.embed.bias"], dev),
            norm_w=_dev_f32(sd["backbone.norm.weight"], dev), norm_b=_dev_f32(sd["backbone.norm.bias"], dev),
            final_ln_w=_dev_f32(sd["backbone.final_layer_norm.weight"], dev), final_ln_b=_dev_f32(sd["backbone.final_layer_norm.bias"], dev),
            head_w=_dev_f16(sd["head.out.weight"], dev), head_b=_dev_f32(sd["head.out.bias"], dev),
            window=_dev_f32(sd["head.istft.window"], dev))
        blocks = (_lib.ConvNextW * c.num_layers)(*[_pack_convnext(sd, f"backbone.convnext.{l}.", dev, self._keep) for l in range(c.num_layers)])
        self._w = dict(tensors=w, blocks=blocks)
        self._engine = None

    @torch.inference_mode()
    def decode(self, mel: torch.Tensor) -> torch.Tensor:
        """mel [B, 100, T] -> wav [B, hop*(T-1)]"""
        if self._w is None:
            raise _lib.CtpError("Vocos weights are not on a CUDA device")
        if self._engine is None:
            self._engine = VocoderEngine(None, self)
        wavs = self._engine.decode_mel([mel[b].permute(1, 0) for b in range(mel.shape[0])])
        return torch.stack(wavs)


class VocoderEngine:
    """One libctp vocoder handle bound to a DVAE and/or a Vocos weight set."""

    # Rows of one workspace group.  74 * 512: the persistent GEMMs then run in whole waves of 128- / 256-row tiles on 148 SMs, and a
    # configs[1] batch (32 utterances x 512 code frames = 33 000 mel rows with their guard rows) is ONE group — with 32 768 rows it
    # was 31 utterances plus a second ~190-launch pass for the last one.
    def __init__(self, dvae: Optional[DVAE], vocos: Optional[Vocos], max_frames: int = 74 * 512):
        assert dvae is not None or vocos is not None
        self.dvae, self.vocos = dvae, vocos
        self.device = (dvae or vocos).device
        self._handle = C.c_void_p(0)
        self._max_frames = 0
        self._want_frames = int(max_frames)

    def _ensure(self, frames_needed: int):
        if self._handle and frames_needed <= self._max_frames:
            return
        lib = _lib.lib()
        if self._handle:
            lib.ctp_voc_destroy(self._handle)
            self._handle = C.c_void_p(0)
        dc = self.dvae.cfg if self.dvae else DVAEConfig()
        vc = self.vocos.cfg if self.vocos else VocosConfig()
        self._max_frames = max(self._want_frames, frames_needed)
        cfg = _lib.VocCfg(dvae_idim=dc.idim, dvae_bn=dc.bn_dim, dvae_hidden=dc.hidden, dvae_layers=dc.n_layer, dvae_odim=dc.odim,
                          dvae_dilation=dc.dilation, n_mels=dc.n_mels, use_vq=1 if dc.vq else 0, voc_dim=vc.dim, voc_inter=vc.intermediate_dim,
                          voc_layers=vc.num_layers, n_fft=vc.n_fft, hop=vc.hop_length, max_frames=self._max_frames)
        h = C.c_void_p(0)
        with torch.cuda.device(self.device):
            _lib.check(lib.ctp_voc_create(C.byref(h), C.byref(cfg)), "ctp_voc_create")
            self._handle = h
            w = _lib.VocWeights()
            if self.dvae is not None:
                for k, t in self.dvae._w["tensors"].items():
                    setattr(w, k, t.data_ptr())
                w.dvae_blocks = C.cast(self.dvae._w["blocks"], C.POINTER(_lib.ConvNextW))
            if self.vocos is not None:
                for k, t in self.vocos._w["tensors"].items():
                    setattr(w, k, t.data_ptr())
                w.voc_blocks = C.cast(self.vocos._w["blocks"], C.POINTER(_lib.ConvNextW))
            _lib.check(lib.ctp_voc_bind_weights(self._handle, C.byref(w)), "ctp_voc_bind_weights")

    def __del__(self):
        try:
            if self._handle:
                _lib.lib().ctp_voc_destroy(self._handle)
        except Exception:
            pass

    @torch.inference_mode()
    def decode_batch(self, items: Sequence[torch.Tensor], want_wav: bool = True, want_mel: bool = False):
        """items[i]: hiddens ``[n_i, 2*idim]`` (float) or codes ``[n_i, 4]`` (int) -> (wavs, mels).

        wavs[i]: fp32 ``[hop*(2 n_i - 1)]``; mels[i]: fp32 ``[2 n_i, 100]``."""
        dev = self.device
        vq = bool(self.dvae.cfg.vq)
        lens = [int(t.shape[0]) for t in items]
        keep = [i for i, n in enumerate(lens) if n >= 1]
        if not keep:
            return [torch.zeros(0, device=dev) for _ in items], [torch.zeros(0, 100, device=dev) for _ in items]
        glens = [lens[i] for i in keep]
        hop = self.vocos.cfg.hop_length if self.vocos else 256
        with torch.cuda.device(dev):
            if vq:
                src = torch.cat([items[i].to(dev, torch.int32).reshape(lens[i], -1) for i in keep]).contiguous()
            else:
                src = torch.cat([items[i].to(dev, torch.float32) for i in keep]).contiguous()
            self._ensure(max(2 * n for n in glens) + 32)
            wlen = [hop * (2 * n - 1) for n in glens]
            offs = np.concatenate([[0], np.cumsum(wlen)]).astype(np.int64)
            wav = torch.empty(int(offs[-1]), device=dev, dtype=torch.float32) if want_wav else None
            mel = torch.empty(2 * sum(glens), 100, device=dev, dtype=torch.float32) if want_mel else None
            n = len(glens)
            lens_arr = (C.c_int32 * n)(*glens)
            offs_arr = (C.c_int64 * n)(*offs[:-1].tolist())
            torch.cuda.nvtx.range_push("ctp.voc.decode")   # (reference precedent for NVTX: trt_models/predictor.py:92,142,159,164)
            try:
                _lib.check(_lib.lib().ctp_voc_decode(self._handle, n, lens_arr, _lib.ptr(src), _lib.ptr(wav), offs_arr, _lib.ptr(mel),
                                                     _lib.stream_ptr()), "ctp_voc_decode")
            finally:
                torch.cuda.nvtx.range_pop()
        wavs: List[torch.Tensor] = [torch.zeros(0, device=dev) for _ in items]
        mels: List[torch.Tensor] = [torch.zeros(0, 100, device=dev) for _ in items]
        mrow = 0
        for k, i in enumerate(keep):
            if want_wav:
                wavs[i] = wav[int(offs[k]): int(offs[k + 1])]
            if want_mel:
                mels[i] = mel[mrow: mrow + 2 * glens[k]]
                mrow += 2 * glens[k]
        return wavs, mels

    @torch.inference_mode()
    def decode_mel(self, mels: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """mels[i]: ``[T_i, 100]`` -> wav ``[hop*(T_i-1)]``"""
        dev = self.device
        hop = self.vocos.cfg.hop_length
        lens = [int(m.shape[0]) for m in mels]
        with torch.cuda.device(dev):
            src = torch.cat([m.to(dev, torch.float32) for m in mels]).contiguous()
            self._ensure(max(lens) + 32)
            wlen = [hop * (n - 1) for n in lens]
            offs = np.concatenate([[0], np.cumsum(wlen)]).astype(np.int64)
            wav = torch.empty(int(offs[-1]), device=dev, dtype=torch.float32)
            n = len(lens)
            _lib.check(_lib.lib().ctp_voc_decode_mel(self._handle, n, (C.c_int32 * n)(*lens), _lib.ptr(src), _lib.ptr(wav),
                                                     (C.c_int64 * n)(*offs[:-1].tolist()), _lib.stream_ptr()), "ctp_voc_decode_mel")
        return [wav[int(offs[k]): int(offs[k + 1])] for k in range(n)]


# receptive field of one waveform sample in mel frames: DVAE conv_in 2 + 12 ConvNeXt blocks x (3 taps x dilation 2) + out_conv 1 = 75,
# Vocos embed 3 + 8 blocks x 3 = 27, ISTFT overlap (n_fft / hop = 4 frames, i.e. +-2) -> 104 mel frames = 52 code frames each side
def vocoder_halo_code_frames(dvae: DVAE, vocos: Vocos) -> int:
    dc, vc = dvae.cfg, vocos.cfg
    dvae_reach = 1 + 1 + dc.n_layer * (dc.kernel // 2) * dc.dilation + 1
    voc_reach = 3 + vc.num_layers * 3
    istft_reach = vc.n_fft // vc.hop_length // 2
    return (dvae_reach + voc_reach + istft_reach + 1) // 2 + 1


class StreamingVocoder:
    """Incremental ``_decode_to_wavs`` for stream mode (reference intent: chattts_plus_pipeline.py:441-464 yields the NEW slice
    ``wavs[:, a:b]`` per chunk; gpt.py:533-543 hands over the growing ids / hiddens every ``stream_batch`` steps).

    Each ``push`` decodes, per utterance, only a window of code frames — the frames that are new since the last push plus ``halo``
    frames of context on either side — and returns the samples that have just become FINAL: sample s is final once every frame
    inside its receptive field (+-halo code frames) has been generated, or the utterance has ended.  The window's own left edge is
    ``halo`` frames before the first sample it emits, so the zero padding the kernels apply there cannot reach an emitted sample.
    Concatenating the returned pieces reproduces the one-shot waveform; the work per chunk is bounded by
    ``stream_batch + 2 * halo`` frames per utterance, independent of how long the utterance already is."""

    def __init__(self, engine: VocoderEngine, n_utt: int, halo: Optional[int] = None):
        self.engine = engine
        self.halo = int(halo) if halo is not None else vocoder_halo_code_frames(engine.dvae, engine.vocos)
        self.emitted = [0] * n_utt          # samples already returned, per utterance
        self.frames_decoded = 0             # bookkeeping for tests: code frames sent through the kernels so far
        self.hop2 = 2 * engine.vocos.cfg.hop_length   # samples per code frame

    @torch.inference_mode()
    def push(self, items: Sequence[torch.Tensor], final: bool) -> List[torch.Tensor]:
        dev = self.engine.device
        hop2, halo = self.hop2, self.halo
        win, meta = [], []
        for i, t in enumerate(items):
            n = int(t.shape[0])
            total = hop2 * n - hop2 // 2 if n >= 1 else 0                      # hop * (2 n - 1)
            upto = total if final else max(0, min(total, hop2 * (n - halo)))
            if upto <= self.emitted[i]:
                meta.append(None)
                continue
            a = max(0, self.emitted[i] // hop2 - halo)                          # first code frame of the window
            win.append(t[a:n])
            meta.append((a, upto))
            self.frames_decoded += n - a
        out = [torch.zeros(0, device=dev) for _ in items]
        if win:
            wavs, _ = self.engine.decode_batch(win)
            k = 0
            for i, m in enumerate(meta):
                if m is None:
                    continue
                a, upto = m
                out[i] = wavs[k][self.emitted[i] - hop2 * a: upto - hop2 * a]
                self.emitted[i] = upto
                k += 1
        return out
