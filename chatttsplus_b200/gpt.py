"""``GPT`` — host-side mirror of reference ``chattts_plus/models/gpt.py`` whose trunk, heads, sampler and
generate loop run in libctp (hand-written sm_100a CUDA) behind the C ABI of include/ctp.h.

Kept from the reference (so the pipeline / webui / tests call it unchanged):
  * constructor ``GPT(gpt_config, num_audio_tokens, num_text_tokens, num_vq, use_flash_attn, model_path=...)``
    (gpt.py:25-81), ``from_pretrained`` (strict state-dict load, gpt.py:84-85), ``.eval()``, ``.to(device, dtype=)``
  * ``__call__(input_ids, text_mask) -> emb``            (gpt.py:117-149)
  * ``generate(emb, inputs_ids, temperature, eos_token, attention_mask, max_new_token, min_new_token,
    logits_warpers, logits_processors, infer_text, return_attn, return_hidden, stream, show_tqdm,
    ensure_non_empty, stream_batch, context)`` -> generator of ``GenerationOutputs``   (gpt.py:313-569)
  * ``.num_vq``, ``.emb_code[i].num_embeddings``, ``GPT.Context``, ``GPT.GenerationOutputs``
  * LoRA hook: the reference swaps ``self.gpt`` for a peft-merged copy (chattts_plus_pipeline.py:420-434); here
    ``load_lora(dir)`` / ``unload_lora()`` merge W += (alpha/r) B A into the packed q/k/v/o and re-bind.
Host code only packs weights and drives the C ABI; there is no PyTorch compute fallback.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple, Union

import torch

from . import _lib
from .commons import logger as _logger
from .processors import flatten as _flatten_processors
from .synth import GPTConfig


class _nvtx:
    """NVTX range (reference precedent: chattts_plus/trt_models/predictor.py:92,142,159,164) — shows up in nsys / ncu timelines."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *a):
        torch.cuda.nvtx.range_pop()
        return False


class _EmbInfo:
    """Stand-in for nn.Embedding where callers only read ``num_embeddings`` (chattts_plus_pipeline.py:209)."""

    def __init__(self, num_embeddings: int, embedding_dim: int):
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim


def expected_gpt_keys(cfg: GPTConfig) -> Dict[str, Tuple[int, ...]]:
    H, I = cfg.hidden_size, cfg.intermediate_size
    k: Dict[str, Tuple[int, ...]] = {}
    for l in range(cfg.num_hidden_layers):
        p = f"gpt.layers.{l}."
        k[p + "input_layernorm.weight"] = (H,)
        k[p + "post_attention_layernorm.weight"] = (H,)
        for nm in ("q_proj", "k_proj", "v_proj", "o_proj"):
            k[p + f"self_attn.{nm}.weight"] = (H, H)
        k[p + "mlp.gate_proj.weight"] = (I, H)
        k[p + "mlp.up_proj.weight"] = (I, H)
        k[p + "mlp.down_proj.weight"] = (H, I)
    k["gpt.norm.weight"] = (H,)
    for q in range(cfg.num_vq):
        k[f"emb_code.{q}.weight"] = (cfg.num_audio_tokens, H)
        k[f"head_code.{q}.parametrizations.weight.original0"] = (cfg.num_audio_tokens, 1)
        k[f"head_code.{q}.parametrizations.weight.original1"] = (cfg.num_audio_tokens, H)
    k["emb_text.weight"] = (cfg.num_text_tokens, H)
    k["head_text.parametrizations.weight.original0"] = (cfg.num_text_tokens, 1)
    k["head_text.parametrizations.weight.original1"] = (cfg.num_text_tokens, H)
    return k


def _fold_weight_norm(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    # torch weight_norm(dim=0): W = g * v / ||v|| per output row (gpt.py:57-77)
    return g.float() * v.float() / v.float().norm(dim=1, keepdim=True)


class _TrunkView:
    """Stand-in for ``GPT.gpt`` (reference: the LlamaModel instance)."""

    def __init__(self, owner: "GPT"):
        self._owner = owner
        self.config = owner.cfg

    def cpu(self):
        return self

    def to(self, *a, **k):
        return self

    def eval(self):
        return self


class GPT:
    class Context:
        def __init__(self):
            self._interrupt = False

        def set(self, v: bool):
            self._interrupt = v

        def get(self) -> bool:
            return self._interrupt

    @dataclass(repr=False, eq=False)
    class GenerationOutputs:
        ids: List[torch.Tensor]
        attentions: List[Optional[Tuple[torch.FloatTensor, ...]]]
        hiddens: List[torch.Tensor]

    def __init__(self, gpt_config: dict, num_audio_tokens: int = 626, num_text_tokens: int = 21178, num_vq=4,
                 use_flash_attn=False, **kwargs):
        self.logger = _logger.get_logger(self.__class__.__name__)
        g = dict(gpt_config)
        self.cfg = GPTConfig(
            hidden_size=int(g.get("hidden_size", 768)), intermediate_size=int(g.get("intermediate_size", 3072)),
            num_attention_heads=int(g.get("num_attention_heads", 12)), num_hidden_layers=int(g.get("num_hidden_layers", 20)),
            num_audio_tokens=int(num_audio_tokens), num_text_tokens=int(num_text_tokens), num_vq=int(num_vq),
            rms_norm_eps=float(g.get("rms_norm_eps", 1e-6)), rope_theta=float(g.get("rope_theta", 10000.0)),
            max_position_embeddings=int(g.get("max_position_embeddings", 4096)))
        self.num_vq = int(num_vq)
        self.num_audio_tokens = int(num_audio_tokens)
        self.model_dim = self.cfg.hidden_size
        self.use_flash_attn = use_flash_attn
        self.is_te_llama = False
        self.emb_code = [_EmbInfo(self.num_audio_tokens, self.model_dim) for _ in range(self.num_vq)]
        self.emb_text = _EmbInfo(self.cfg.num_text_tokens, self.model_dim)
        self.device = torch.device("cpu")
        self._state: Optional[Dict[str, torch.Tensor]] = None   # fp32 CPU state dict (reference key names)
        self._packed: Optional[Dict[str, torch.Tensor]] = None  # device tensors bound to the library
        self._base_w: Optional[Dict[str, torch.Tensor]] = None  # un-merged wqkv / wo / wgu / wdown while a LoRA adapter is merged
        self._handle = C.c_void_p(0)
        self._max_batch = int(kwargs.get("max_batch", 32))
        self._max_seq = 0
        self._max_batch_alloc = 0
        self.record_timing = False   # bench.py: CUDA events around prefill / decode loop on the launching stream
        self.timing: Dict[str, float] = {}
        self.model_path = kwargs.get("model_path", None)
        if self.model_path:
            self.logger.info(f"loading GPT pretrained model: {self.model_path}")
            self.from_pretrained(self.model_path)

    # ---- nn.Module-ish surface ---------------------------------------------------------------------------
    def eval(self):
        return self

    def parameters(self):
        return iter(self._packed.values()) if self._packed else iter(())

    def from_pretrained(self, file_path: str):
        self.load_state_dict(torch.load(file_path, weights_only=True, mmap=True))

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        exp = expected_gpt_keys(self.cfg)
        missing = [k for k in exp if k not in sd]
        unexpected = [k for k in sd if k not in exp]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for GPT: missing {missing[:5]}... unexpected {unexpected[:5]}...")
        for k, shp in exp.items():
            if k in sd and tuple(sd[k].shape) != shp:
                raise RuntimeError(f"size mismatch for {k}: {tuple(sd[k].shape)} vs {shp}")
        self._state = {k: sd[k].detach().to("cpu") for k in exp if k in sd}
        if self.device.type == "cuda":
            self._pack_and_bind()
        return self

    def to(self, device=None, dtype=None, **kw):
        if device is not None:
            device = torch.device(device)
            if device.type == "cuda":
                if not torch.cuda.is_available():
                    raise _lib.CtpError("chatttsplus_b200.GPT needs a CUDA (sm_100a) device; there is no CPU path")
                if device.index is None:
                    device = torch.device("cuda", torch.cuda.current_device())
                self.device = device
                if self._state is not None:
                    self._pack_and_bind()
            elif device.type != "cpu":
                raise _lib.CtpError(f"unsupported device {device}")
        return self

    def cpu(self):
        return self

    # ---- weights -----------------------------------------------------------------------------------------
    def _pack_and_bind(self):
        c, sd, dev = self.cfg, self._state, self.device
        L = c.num_hidden_layers
        f16 = torch.float16

        def stack(fmt):
            return torch.stack([sd[fmt.format(l)] for l in range(L)])

        with torch.cuda.device(dev):
            p: Dict[str, torch.Tensor] = {}
            q, k, v = (stack("gpt.layers.{}.self_attn.%s_proj.weight" % n) for n in ("q", "k", "v"))
            p["wqkv"] = torch.cat([q, k, v], dim=1).to(dev, f16).contiguous()          # [L, 3H, H]
            p["wo"] = stack("gpt.layers.{}.self_attn.o_proj.weight").to(dev, f16).contiguous()
            p["wgu"] = torch.cat([stack("gpt.layers.{}.mlp.gate_proj.weight"), stack("gpt.layers.{}.mlp.up_proj.weight")],
                                 dim=1).to(dev, f16).contiguous()                     # [L, 2I, H]
            p["wdown"] = stack("gpt.layers.{}.mlp.down_proj.weight").to(dev, f16).contiguous()
            p["ln1"] = stack("gpt.layers.{}.input_layernorm.weight").to(dev, torch.float32).contiguous()
            p["ln2"] = stack("gpt.layers.{}.post_attention_layernorm.weight").to(dev, torch.float32).contiguous()
            p["norm_f"] = sd["gpt.norm.weight"].to(dev, torch.float32).contiguous()
            p["emb_code"] = torch.stack([sd[f"emb_code.{i}.weight"] for i in range(c.num_vq)]).to(dev, f16).contiguous()
            p["head_code"] = torch.cat([_fold_weight_norm(sd[f"head_code.{i}.parametrizations.weight.original0"],
                                                          sd[f"head_code.{i}.parametrizations.weight.original1"])
                                        for i in range(c.num_vq)]).to(dev, f16).contiguous()
            p["emb_text"] = sd["emb_text.weight"].to(dev, f16).contiguous()
            p["head_text"] = _fold_weight_norm(sd["head_text.parametrizations.weight.original0"],
                                               sd["head_text.parametrizations.weight.original1"]).to(dev, f16).contiguous()
        self._packed = p
        self._base_w = None
        self._bind()

    def _ensure_handle(self, batch: int, seq: int):
        if self._handle and batch <= self._max_batch_alloc and seq <= self._max_seq:
            return
        if self._handle:
            _lib.lib().ctp_gpt_destroy(self._handle)
            self._handle = C.c_void_p(0)
        c = self.cfg
        self._max_batch_alloc = max(batch, self._max_batch)
        self._max_seq = max(seq, 256)
        cfg = _lib.GptCfg(n_layers=c.num_hidden_layers, hidden=c.hidden_size, n_heads=c.num_attention_heads,
                          inter=c.intermediate_size, num_vq=c.num_vq, num_audio=c.num_audio_tokens, num_text=c.num_text_tokens,
                          max_batch=self._max_batch_alloc, max_seq=self._max_seq, rms_eps=c.rms_norm_eps, rope_theta=c.rope_theta)
        h = C.c_void_p(0)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ctp_gpt_create(C.byref(h), C.byref(cfg)), "ctp_gpt_create")
        self._handle = h
        self._bind()

    def _bind(self):
        if not self._handle or self._packed is None:
            return
        p = self._packed
        w = _lib.GptWeights(**{k: p[k].data_ptr() for k in
                               ("wqkv", "wo", "wgu", "wdown", "ln1", "ln2", "norm_f", "emb_code", "head_code", "emb_text", "head_text")})
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ctp_gpt_bind_weights(self._handle, C.byref(w)), "ctp_gpt_bind_weights")

    def __del__(self):
        try:
            if self._handle:
                _lib.lib().ctp_gpt_destroy(self._handle)
        except Exception:
            pass

    # ---- LoRA (A12) --------------------------------------------------------------------------------------
    _LORA_TARGETS = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")

    @staticmethod
    def _pattern_value(pattern: Optional[dict], module_name: str, default):
        """peft ``rank_pattern`` / ``alpha_pattern`` lookup: the first key that equals the module name or matches its tail."""
        import re
        for key, val in (pattern or {}).items():
            if module_name == key or re.match(rf"(.*\.)?{key}$", module_name):
                return val
        return default

    def merge_lora(self, lora_sd: Dict[str, torch.Tensor], alpha: float, r: int, use_rslora: bool = False,
                   rank_pattern: Optional[dict] = None, alpha_pattern: Optional[dict] = None, fan_in_fan_out: bool = False):
        """peft ``merge_and_unload`` for LoRA adapters: W' = W + scale * B @ A on every adapted Linear of the trunk — q/k/v/o
        (configs/train/train_voice_clone_lora.yaml:72-80) and gate/up/down — with scale = alpha / r (alpha / sqrt(r) for rsLoRA),
        per-module r / alpha from ``rank_pattern`` / ``alpha_pattern``.  Every ``lora_*`` tensor of the adapter must be consumed:
        an adapter on a module this trunk does not have (or any DoRA / embedding adapter tensor) raises instead of being dropped."""
        if self._packed is None:
            raise _lib.CtpError("merge_lora: model not on device")
        c = self.cfg
        H, I = c.hidden_size, c.intermediate_size
        if self._base_w is None:
            self._base_w = {k: self._packed[k].clone() for k in ("wqkv", "wo", "wgu", "wdown")}
        w = {k: v.float() for k, v in self._base_w.items()}
        consumed = set()
        for l in range(c.num_hidden_layers):
            for nm in self._LORA_TARGETS:
                sub = "self_attn" if nm in ("q_proj", "k_proj", "v_proj", "o_proj") else "mlp"
                a = b = None
                for pre in ("base_model.model.", "base_model.model.model.", ""):
                    for mid in ("lora_A.weight", "lora_A.default.weight"):
                        ka = f"{pre}layers.{l}.{sub}.{nm}.{mid}"
                        if ka in lora_sd:
                            kb = ka.replace("lora_A", "lora_B")
                            if kb not in lora_sd:
                                raise _lib.CtpError(f"LoRA adapter has {ka} but not {kb}")
                            a, b = lora_sd[ka], lora_sd[kb]
                            consumed.update((ka, kb))
                            break
                    if a is not None:
                        break
                if a is None:
                    continue
                name = f"layers.{l}.{sub}.{nm}"
                r_m = int(a.shape[0])
                r_cfg = int(self._pattern_value(rank_pattern, name, r))
                if r_cfg != r_m:
                    raise _lib.CtpError(f"LoRA rank mismatch for {name}: adapter tensors have r={r_m}, config says {r_cfg}")
                alpha_m = float(self._pattern_value(alpha_pattern, name, alpha))
                scale = alpha_m / (r_m ** 0.5) if use_rslora else alpha_m / r_m
                delta = scale * (b.to(self.device, torch.float32) @ a.to(self.device, torch.float32))
                if fan_in_fan_out:
                    delta = delta.t()
                if nm in ("q_proj", "k_proj", "v_proj"):
                    j = ("q_proj", "k_proj", "v_proj").index(nm)
                    w["wqkv"][l, j * H:(j + 1) * H] += delta
                elif nm == "o_proj":
                    w["wo"][l] += delta
                elif nm == "gate_proj":
                    w["wgu"][l, :I] += delta
                elif nm == "up_proj":
                    w["wgu"][l, I:] += delta
                else:
                    w["wdown"][l] += delta
        unused = sorted(k for k in lora_sd if "lora_" in k and k not in consumed)
        if unused:
            raise _lib.CtpError(f"LoRA adapter tensors not consumed by the merge (unsupported target or variant): {unused[:4]}"
                                f"{' ...' if len(unused) > 4 else ''}")
        if not consumed:
            raise _lib.CtpError("LoRA adapter holds no lora_A / lora_B tensors for this trunk")
        for k in w:
            self._packed[k] = w[k].to(torch.float16).contiguous()
        self._bind()

    def load_lora(self, lora_path: str):
        """peft adapter directory: adapter_config.json + adapter_model.safetensors|.bin (webui.py:48-62)."""
        with open(os.path.join(lora_path, "adapter_config.json"), "r", encoding="utf-8") as f:
            ac = json.load(f)
        if str(ac.get("peft_type", "LORA")).upper() != "LORA":
            raise _lib.CtpError(f"adapter type {ac.get('peft_type')!r} is not supported (LoRA only)")
        for key in ("use_dora", "modules_to_save", "layers_to_transform", "layer_replication", "megatron_config"):
            if ac.get(key):   # peft's merge honours these; the native merge does not implement them
                raise _lib.CtpError(f"adapter_config.json: {key}={ac[key]!r} is not supported by the native merge")
        if ac.get("bias", "none") != "none":
            raise _lib.CtpError(f"adapter_config.json: bias={ac['bias']!r} is not supported (the trunk has no biases)")
        st_path = os.path.join(lora_path, "adapter_model.safetensors")
        if os.path.exists(st_path):
            from safetensors.torch import load_file
            sd = load_file(st_path)
        else:
            sd = torch.load(os.path.join(lora_path, "adapter_model.bin"), weights_only=True, map_location="cpu")
        self.merge_lora(sd, float(ac.get("lora_alpha", 16)), int(ac.get("r", 8)), use_rslora=bool(ac.get("use_rslora", False)),
                        rank_pattern=ac.get("rank_pattern") or None, alpha_pattern=ac.get("alpha_pattern") or None,
                        fan_in_fan_out=bool(ac.get("fan_in_fan_out", False)))

    def unload_lora(self):
        if self._base_w is not None:
            self._packed.update(self._base_w)
            self._base_w = None
            self._bind()

    @property
    def gpt(self):
        """The reference swaps / parks ``GPT.gpt`` (the LlamaModel) around a LoRA merge (chattts_plus_pipeline.py:425-432,468-470).
        Here the trunk lives in the library handle; this view keeps attribute access and ``.cpu()`` / ``.to()`` chains harmless.
        Use ``load_lora`` / ``unload_lora`` for adapters."""
        return _TrunkView(self)

    @gpt.setter
    def gpt(self, value):
        if not isinstance(value, _TrunkView):
            raise _lib.CtpError("GPT.gpt cannot be replaced by a torch module: the trunk runs in libctp; use GPT.load_lora(dir)")

    # ---- get_emb (A1) ------------------------------------------------------------------------------------
    def __call__(self, input_ids: torch.Tensor, text_mask: torch.Tensor) -> torch.Tensor:
        return self.forward(input_ids, text_mask)

    def forward(self, input_ids: torch.Tensor, text_mask: torch.Tensor) -> torch.Tensor:
        if self._packed is None:
            raise _lib.CtpError("GPT weights are not on a CUDA device: call .to('cuda') after loading")
        B, L0, nq = input_ids.shape
        assert nq == self.num_vq
        self._ensure_handle(B, L0 + 1)
        ids = input_ids.to(self.device, torch.int32).contiguous()
        tm = text_mask.to(self.device, torch.uint8).contiguous()
        emb = torch.empty(B, L0, self.model_dim, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ctp_gpt_embed_prompt(self._handle, B, L0, _lib.ptr(ids), _lib.ptr(tm), _lib.ptr(emb),
                                                       _lib.stream_ptr()), "ctp_gpt_embed_prompt")
        return emb

    # ---- generate (A3-A5, A13-A18) -----------------------------------------------------------------------
    @staticmethod
    def _pad_lens(attention_mask: Optional[torch.Tensor], B: int, L0: int) -> List[int]:
        if attention_mask is None:
            return [0] * B
        m = attention_mask.to("cpu").bool()
        pads = (~m).sum(1).tolist()
        for b, p in enumerate(pads):  # must be LEFT padding (tokenizer.py:87-121 pads on the left)
            if p and (bool(m[b, :p].any()) or not bool(m[b, p:].all())):
                raise ValueError("attention_mask must be left-padded (zeros then ones) for the B200 KV-cache layout")
            if p >= L0:
                raise ValueError("a sequence in the batch is entirely padding")
        return [int(p) for p in pads]

    def _sample_cfg(self, temperature: torch.Tensor, eos_token: int, min_new_token: int, logits_warpers, logits_processors, seed=0):
        sp = _flatten_processors(logits_warpers, logits_processors)
        cfg = _lib.SampleCfg()
        t = temperature.detach().float().cpu().reshape(-1).tolist()
        if len(t) == 1:
            t = t * self.num_vq
        for i in range(self.num_vq):
            cfg.temperature[i] = float(t[i])
        cfg.rep_penalty = sp.rep_penalty
        cfg.rep_window = sp.rep_window
        cfg.rep_max_ids = sp.rep_max_ids
        cfg.top_p = sp.top_p if 0.0 < sp.top_p < 1.0 else 0.0
        cfg.top_k = sp.top_k
        cfg.min_keep = sp.min_keep
        cfg.eos = int(eos_token)
        cfg.min_new = int(min_new_token)
        cfg.seed = int(seed)
        return cfg

    @torch.no_grad()
    def generate(self, emb: torch.Tensor, inputs_ids: torch.Tensor, temperature: torch.Tensor,
                 eos_token: Union[int, torch.Tensor], attention_mask: Optional[torch.Tensor] = None, max_new_token=2048,
                 min_new_token=0, logits_warpers=[], logits_processors=[], infer_text=False, return_attn=False,
                 return_hidden=False, stream=False, show_tqdm=True, ensure_non_empty=True, stream_batch=24,
                 context=None, uniforms: Optional[torch.Tensor] = None):
        """Generator of GenerationOutputs, like reference GPT.generate (gpt.py:313-569).

        ``uniforms`` (extra, optional): fp32 [max_new_token, B*num_vq] in [0,1) consumed by the inverse-CDF draw;
        default = ``torch.rand`` on the CUDA generator (so ``TorchSeedContext`` / ``torch.manual_seed`` seed it).
        """
        if return_attn:
            raise NotImplementedError("return_attn is not supported by the fused attention kernels")
        if self._packed is None:
            raise _lib.CtpError("GPT weights are not on a CUDA device")
        context = context or GPT.Context()
        lib = _lib.lib()
        dev = self.device
        B, L0, nq = inputs_ids.shape
        eos = int(eos_token)
        max_new = int(max_new_token)
        cols = 1 if infer_text else nq   # refine-text pass: one sampling column per sequence (gpt.py:459-467)
        self._ensure_handle(B, L0 + max_new + 1)
        pads = self._pad_lens(attention_mask, B, L0)
        H = self.model_dim
        with torch.cuda.device(dev):
            emb32 = emb.to(dev, torch.float32).contiguous()
            ids_buf = torch.zeros(B, max_new, nq, device=dev, dtype=torch.int32)
            hid_buf = torch.empty(B, max_new, H, device=dev, dtype=torch.float32) if return_hidden else None
            end_idx = torch.zeros(B, device=dev, dtype=torch.int32)
            finish = torch.zeros(B, device=dev, dtype=torch.uint8)
            bufs = _lib.GenBuffers(ids=ids_buf.data_ptr(), hiddens=hid_buf.data_ptr() if hid_buf is not None else None,
                                   end_idx=end_idx.data_ptr(), finish=finish.data_ptr(), max_new=max_new)
            pad_arr = (C.c_int32 * B)(*pads)
            cfg = self._sample_cfg(temperature, eos, min_new_token, logits_warpers, logits_processors)
            strm = _lib.stream_ptr()
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if self.record_timing else None
            if evs:
                evs[0].record()
            text_flag = 1 if infer_text else 0
            with _nvtx("ctp.gpt.prefill"):
                _lib.check(lib.ctp_gpt_prefill(self._handle, B, L0, _lib.ptr(emb32), pad_arr, C.byref(bufs), text_flag, strm), "ctp_gpt_prefill")

            def draw_uniforms():
                if uniforms is not None:
                    u = uniforms.to(dev, torch.float32).contiguous()
                    assert u.shape == (max_new, B * cols), f"uniforms must be [{max_new}, {B * cols}]"
                    return u
                return torch.rand(max_new, B * cols, device=dev, dtype=torch.float32)

            u = draw_uniforms()
            # step 0 (+ gpt.py:496-525: if any sequence ends immediately, draw again).  A retry only rewinds the generation state
            # (the prompt's KV cache and last-position logits are unchanged) and redraws; the draw of the last attempt is kept.
            attempts = 8 if (ensure_non_empty and uniforms is None) else 1
            for attempt in range(attempts):
                _lib.check(lib.ctp_gpt_sample_step(self._handle, C.byref(cfg), _lib.ptr(u[0]), strm), "ctp_gpt_sample_step")
                if attempt + 1 == attempts or not bool(finish.any().item()):
                    break
                self.logger.info("unexpected end at index %s; regenerate in order to ensure non-empty"
                                 % str(finish.nonzero().flatten().tolist()))
                _lib.check(lib.ctp_gpt_rewind(self._handle, strm), "ctp_gpt_rewind")
                u = draw_uniforms()

            pbar = None
            if show_tqdm:
                from tqdm import tqdm
                pbar = tqdm(total=max_new, desc="text" if infer_text else "code",
                            bar_format="{l_bar}{bar}| {n_fmt}/{total_fmt}(max) [{elapsed}, {rate_fmt}{postfix}]")
                pbar.update(1)

            def outputs():
                n = end_idx.to("cpu").tolist()
                ids = [ids_buf[b, : n[b]].to(inputs_ids.dtype) for b in range(B)]
                if infer_text:
                    ids = [i[:, 0] for i in ids]   # gpt.py:298-299
                hid = [hid_buf[b, : n[b]] for b in range(B)] if hid_buf is not None else []
                return GPT.GenerationOutputs(ids=ids, attentions=[], hiddens=hid)

            if evs:
                evs[1].record()
            remaining = max_new - 1
            chunk = int(stream_batch) if stream else remaining
            done_total = 1
            all_done = False
            while remaining > 0 and not all_done and not context.get():
                n_it = min(chunk, remaining)
                steps_done = C.c_int32(0)
                with _nvtx("ctp.gpt.decode_steps"):
                    _lib.check(lib.ctp_gpt_generate(self._handle, C.byref(cfg), n_it, _lib.ptr(u), 16, C.byref(steps_done), strm),
                               "ctp_gpt_generate")
                remaining -= n_it
                done_total += steps_done.value
                if pbar is not None:
                    pbar.update(steps_done.value)
                if steps_done.value < n_it:
                    all_done = True
                if stream and remaining > 0 and not all_done:
                    all_done = bool(finish.all().item())
                    yield outputs()
            if evs:
                evs[2].record()
            torch.cuda.current_stream().synchronize()
            # raw buffers of the finished run (parity tests compare them with the oracle; finished rows keep decoding, gpt.py:483-494)
            self._last_run = {"ids_buf": ids_buf, "hid_buf": hid_buf, "end_idx": end_idx, "finish": finish, "steps": done_total}
            if evs:
                self.timing = {"prefill_ms": evs[0].elapsed_time(evs[1]), "decode_ms": evs[1].elapsed_time(evs[2]),
                               "decode_steps": done_total - 1, "B": B, "L0": L0}
            if pbar is not None:
                pbar.close()
            if not bool(finish.all().item()):
                if context.get():
                    self.logger.info("generation is interrupted")
                else:
                    self.logger.info(f"incomplete result. hit max_new_token: {max_new_token}")
            last = outputs()
            last.is_final = True   # (extra attribute: stream consumers flush their tail on it)
            yield last

    # ---- low-level hooks used by the parity tests and bench ----------------------------------------------
    def logits_view(self, B: int, text: bool = False) -> torch.Tensor:
        """Copy of the handle's current logits as fp32 [B*num_vq, num_audio] (row = b*num_vq + q)."""
        shape = (B, self.cfg.num_text_tokens) if text else (B * self.num_vq, self.num_audio_tokens)
        out = torch.empty(*shape, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ctp_gpt_copy_outputs(self._handle, _lib.ptr(out), None, _lib.stream_ptr()), "ctp_gpt_copy_outputs")
        return out

    def hidden_view(self, B: int) -> torch.Tensor:
        out = torch.empty(B, self.model_dim, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ctp_gpt_copy_outputs(self._handle, None, _lib.ptr(out), _lib.stream_ptr()), "ctp_gpt_copy_outputs")
        return out
