"""``ChatTTSPlusPipeline`` — the inference API of reference ``chattts_plus/pipelines/chattts_plus_pipeline.py`` with
the generation hot path (``_infer_code`` -> ``GPT.generate``, ``_decode_to_wavs`` -> DVAE + Vocos) running in libctp.

Kept verbatim from the reference (so ``webui.py`` and ``tests/test_pipelines.py`` drive it unchanged):
  ``ChatTTSPlusPipeline(cfg, device=, dtype=, coef=)``, the ``cfg.MODELS[name].{name, infer_type, kwargs}`` model table
  (chattts_plus_pipeline.py:113-129), ``infer(text, stream, lang, skip_refine_text, refine_text_only, use_decoder,
  do_text_*, params_refine_text, params_infer_code, **kwargs{speaker_emb_path, speaker_audio_path,
  speaker_audio_text, lora_path, slice_size, speaker_save_dir})`` -> generator of ``List[Tensor]`` (:472-579),
  ``sample_random_speaker`` / ``_encode_spk_emb`` (:307-331).
Differences (all host-side): no hub downloads (offline image: missing checkpoints raise); LoRA is merged natively
(no peft); utterances are vocoded as one batch; the broken stream slicing (:445-464: a Python list indexed like an array) is
replaced by the evident intent: every yield carries the new samples only (incremental vocoder with a receptive-field halo).
``ChatTTSPlusPipeline.from_models`` builds a pipeline from already constructed models (synthetic-weight tests, bench).
"""
from __future__ import annotations

import os
import time
from typing import List, Optional, Union

import numpy as np
import torch

from . import gpt as _gpt_mod
from . import processors
from . import text as _text
from . import tokenizer as _tok_mod
from . import vocoder as _voc_mod
from .commons import constants, logger
from .commons.utils import InferCodeParams, RefineTextParams

# the plugin table the YAML's ``name`` fields resolve against (reference: getattr(models, name))
MODEL_REGISTRY = {"Tokenizer": _tok_mod.Tokenizer, "DVAE": _voc_mod.DVAE, "GPT": _gpt_mod.GPT, "Vocos": _voc_mod.Vocos}


class ChatTTSPlusPipeline:
    def __init__(self, cfg=None, **kwargs):
        self.logger = logger.get_logger(self.__class__.__name__)
        self.cfg = cfg
        self.device = torch.device(kwargs.get("device", "cuda" if torch.cuda.is_available() else "cpu"))
        self.dtype = kwargs.get("dtype", None) or (torch.float16 if self.device.type == "cuda" else torch.float32)
        self.logger.info(f"device: {str(self.device)}")
        self.logger.info(f"dtype: {str(self.dtype)} (kernels: fp16 operands, fp32 accumulate/residual)")
        self.models_dict = dict()
        self.infer_type = "pytorch"
        self.load_lora = False
        self.normalizer = _text.Normalizer(None)
        self.std = self.mean = None
        self._engines = {}
        if cfg is not None:
            self.load_models(**kwargs)

    # ---- construction ------------------------------------------------------------------------------------
    @classmethod
    def from_models(cls, tokenizer, gpt, dvae_decode, vocos, dvae_encode=None, spk_stat: Optional[torch.Tensor] = None,
                    device="cuda"):
        p = cls(None, device=device)
        p.models_dict = {"tokenizer": tokenizer, "gpt": gpt, "dvae_decode": dvae_decode, "vocos": vocos}
        if dvae_encode is not None:
            p.models_dict["dvae_encode"] = dvae_encode
        if spk_stat is not None:
            p.std, p.mean = spk_stat.to(p.device, torch.float32).chunk(2)
        return p

    def load_models(self, **kwargs):
        """chattts_plus_pipeline.py:54-155 without the hub downloads (this image is offline)."""
        coef = kwargs.get("coef", None)
        self.dave_coef = coef
        for dv in ("dvae_encode", "dvae_decode"):
            if dv in self.cfg.MODELS and coef is not None:
                self.cfg.MODELS[dv]["kwargs"]["coef"] = coef
        for model_name in self.cfg.MODELS:
            entry = self.cfg.MODELS[model_name]
            self.logger.info("loading model: {} >>>>".format(model_name))
            path_org = entry["kwargs"]["model_path"]
            path_new = os.path.join(constants.CHECKPOINT_DIR, path_org.replace("checkpoints/", ""))
            if not os.path.exists(path_new):
                raise FileNotFoundError(f"{path_new} not found (no network in this environment: place the ChatTTS assets under "
                                        f"{constants.CHECKPOINT_DIR} or build the pipeline with ChatTTSPlusPipeline.from_models)")
            entry["kwargs"]["model_path"] = path_new
            if entry["infer_type"] != "pytorch":
                raise ValueError(f"infer_type {entry['infer_type']!r}: this build has one backend (the sm_100a library); "
                                 "use configs/infer/chattts_plus.yaml")
            kw = dict(entry["kwargs"])
            model_ = MODEL_REGISTRY[entry["name"]](**kw)
            if model_name != "tokenizer":
                model_.eval().to(self.device, dtype=self.dtype)
            self.models_dict[model_name] = model_
        spk_stat_path = os.path.join(constants.CHECKPOINT_DIR, "asset/spk_stat.pt")
        assert os.path.exists(spk_stat_path), f"Missing spk_stat.pt: {spk_stat_path}"
        spk_stat = torch.load(spk_stat_path, weights_only=True, mmap=True).to(self.device, dtype=torch.float32)
        self.std, self.mean = spk_stat.chunk(2)
        normalizer_json = os.path.join(constants.CHECKPOINT_DIR, "homophones_map.json")
        self.normalizer = _text.Normalizer(normalizer_json if os.path.exists(normalizer_json) else None)

    def _engine(self, use_decoder: bool) -> _voc_mod.VocoderEngine:
        key = "dvae_decode" if use_decoder else "dvae_encode"
        if key not in self._engines:
            self._engines[key] = _voc_mod.VocoderEngine(self.models_dict[key], self.models_dict["vocos"])
        return self._engines[key]

    # ---- hot path ----------------------------------------------------------------------------------------
    @torch.no_grad()
    def infer_ids(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, text_mask: torch.Tensor,
                  params: InferCodeParams, *, stream: bool = False, use_decoder: bool = True, spk_emb_ids: Optional[int] = None,
                  uniforms: Optional[torch.Tensor] = None, return_codes: bool = False):
        """``_infer_code`` after tokenisation + ``_decode_to_wavs`` (chattts_plus_pipeline.py:195-235,286-305):
        host or device ``input_ids [B, L, num_vq]`` -> generator of ``List[wav]`` (and the GenerationOutputs if asked)."""
        gpt = self.models_dict["gpt"]
        dev = self.device
        input_ids = input_ids.to(dev, non_blocking=True)
        text_mask = text_mask.to(dev, non_blocking=True)
        temperature = params.temperature if isinstance(params.temperature, list) else [params.temperature] * gpt.num_vq
        emb = gpt(input_ids, text_mask)
        if params.spk_emb is not None:
            sid = spk_emb_ids if spk_emb_ids is not None else self.models_dict["tokenizer"].spk_emb_ids
            _tok_mod.apply_spk_emb(emb, params.spk_emb, input_ids, sid)
        num_code = int(gpt.emb_code[0].num_embeddings - 1)
        warpers, procs = processors.gen_logits(num_code=num_code, top_P=params.top_P, top_K=params.top_K,
                                               repetition_penalty=params.repetition_penalty)
        gen = gpt.generate(emb, input_ids, temperature=torch.tensor(temperature), eos_token=num_code,
                           attention_mask=attention_mask, max_new_token=params.max_new_token, min_new_token=params.min_new_token,
                           logits_warpers=warpers, logits_processors=procs, infer_text=False, return_hidden=use_decoder,
                           stream=stream, show_tqdm=params.show_tqdm, ensure_non_empty=params.ensure_non_empty,
                           stream_batch=params.stream_batch, uniforms=uniforms)
        if not stream:
            for result in gen:
                wavs = self._decode_to_wavs(result.hiddens if use_decoder else result.ids, use_decoder)
                yield (wavs, result) if return_codes else wavs
            return
        # stream mode (chattts_plus_pipeline.py:417-419,445-464): every yield carries only the NEW samples of each utterance, so a caller
        # that appends the chunks (webui.py:162-166: gr.Audio(streaming=True)) plays each sample once.  The first
        # ``pass_first_n_batches`` chunks are held back like the reference does and ride along with the next one.
        sv = _voc_mod.StreamingVocoder(self._engine(use_decoder), int(input_ids.shape[0]))
        held = None
        n_chunk = 0
        for result in gen:
            final = bool(getattr(result, "is_final", False))
            self.logger.info("Start decode to wavs >>>>")
            new = sv.push(result.hiddens if use_decoder else result.ids, final=final)
            n_chunk += 1
            if held is not None:
                new = [torch.cat([h, w]) for h, w in zip(held, new)]
                held = None
            if not final and n_chunk <= int(getattr(params, "pass_first_n_batches", 0) or 0):
                held = new
                continue
            yield (new, result) if return_codes else new

    @torch.no_grad()
    def _infer_code(self, text, stream: bool, return_hidden: bool, params: InferCodeParams):
        """Tokenise and run the decoder loop; yields GenerationOutputs (chattts_plus_pipeline.py:157-235)."""
        self.logger.info("Start inference audio code >>>>")
        if not isinstance(text, list):
            text = [text]
        assert len(text), "text should not be empty"
        gpt, tok = self.models_dict["gpt"], self.models_dict["tokenizer"]
        text = [t.replace("[Stts]", "").replace("[spk_emb]", "").replace("[empty_spk]", "").strip() for t in text]
        if params.prompt:
            text = [params.prompt + i for i in text]
        txt_smp = "" if params.txt_smp is None else params.txt_smp
        tag = "[spk_emb]" if params.spk_emb is not None else "[empty_spk]"
        text = [f"[Stts]{tag}{txt_smp}{i}[Ptts]" for i in text]
        input_ids, attention_mask, text_mask = tok.encode(text, gpt.num_vq, prompt_str=params.spk_smp, device="cpu")
        for wavs, result in self.infer_ids(input_ids, attention_mask, text_mask, params, stream=stream, use_decoder=return_hidden,
                                           return_codes=True):
            result._wavs = wavs
            yield result

    @torch.no_grad()
    def _refine_text(self, text, params: RefineTextParams):
        """Refine-text pass (chattts_plus_pipeline.py:237-277): text-token generation with head_text."""
        gpt, tok = self.models_dict["gpt"], self.models_dict["tokenizer"]
        text = [f"[Sbreak]{i}[Pbreak]{params.prompt}" for i in text]
        input_ids, attention_mask, text_mask = tok.encode(text, gpt.num_vq, device="cpu")
        warpers, procs = processors.gen_logits(num_code=tok.len, top_P=params.top_P, top_K=params.top_K,
                                               repetition_penalty=params.repetition_penalty)
        input_ids = input_ids.to(self.device)
        emb = gpt(input_ids, text_mask.to(self.device))
        result = None
        for result in gpt.generate(emb, input_ids, temperature=torch.tensor([params.temperature]), eos_token=tok.eos_token,
                                   attention_mask=attention_mask, max_new_token=params.max_new_token,
                                   min_new_token=params.min_new_token, logits_warpers=warpers, logits_processors=procs,
                                   infer_text=True, stream=False, show_tqdm=params.show_tqdm, ensure_non_empty=params.ensure_non_empty):
            pass
        return result

    @torch.inference_mode()
    def _decode_to_wavs(self, result_list, use_decoder: bool):
        self.logger.info("Start decode to wavs >>>>")
        if len(result_list) == 0:
            return []
        wavs, _ = self._engine(use_decoder).decode_batch(list(result_list))
        return wavs

    # ---- speakers ----------------------------------------------------------------------------------------
    def sample_random_speaker(self) -> str:
        return self._encode_spk_emb(self._sample_random_speaker())

    @staticmethod
    @torch.no_grad()
    def _encode_spk_emb(spk_emb: torch.Tensor) -> str:
        return _tok_mod.Tokenizer._encode_spk_emb(spk_emb)

    @torch.no_grad()
    def _sample_random_speaker(self) -> torch.Tensor:
        dim = self.std.shape[-1]
        return torch.randn(dim, device=self.std.device, dtype=self.std.dtype).mul_(self.std).add_(self.mean)

    def _load_speaker(self, path: str):
        """Accept what the shipped speaker files contain (a torch-saved b14 str) or a tensor ([768] or [1,768])."""
        try:
            spk = torch.load(path, weights_only=True, map_location="cpu")
        except Exception:
            if path.endswith(".safetensors"):
                import safetensors.torch
                spk = next(iter(safetensors.torch.load_file(path).values()))
            else:
                spk = torch.load(path, weights_only=False, map_location="cpu")
        if isinstance(spk, dict):
            spk = next(iter(spk.values()))
        if isinstance(spk, str):
            return spk
        if not isinstance(spk, torch.Tensor):
            raise ValueError(f"speaker embedding file holds {type(spk)}")
        return spk.reshape(-1).float()

    # ---- public API --------------------------------------------------------------------------------------
    def _infer(self, text_in, stream=False, lang=None, skip_refine_text=False, refine_text_only=False, use_decoder=True,
               do_text_normalization=True, do_text_optimization=True, do_homophone_replacement=True,
               params_refine_text=RefineTextParams(), params_infer_code=InferCodeParams(), **kwargs):
        if not isinstance(text_in, list):
            text_in = [text_in]
        if do_text_optimization:  # chattts_plus_pipeline.py:353-377
            text_list = []
            for t in text_in:
                text_list.extend([s.strip() for s in t.split("\n") if s.strip()])
            text_in = _text.merge_short_texts(_text.split_text(text_list))
        text_in = [self.normalizer(t, do_text_normalization, do_homophone_replacement, lang) for t in text_in]
        slice_size = kwargs.get("slice_size", 4)
        gpt = self.models_dict["gpt"]
        for ii in range(0, len(text_in), slice_size):
            text = text_in[ii:ii + slice_size].copy()
            if not skip_refine_text:   # chattts_plus_pipeline.py:399-411
                self.logger.info("Process Text Refinement >>>")
                tok = self.models_dict["tokenizer"]
                refined = self._refine_text(text, params_refine_text)
                text_tokens = [i[i.less(tok.break_0_ids)] for i in refined.ids]
                text = tok.decode(text_tokens)
                self.logger.info("Refine text: ")
                self.logger.info(text)
                if refine_text_only:
                    yield text
            if refine_text_only:
                continue
            for ti in range(len(text)):
                if not text[ti].strip().endswith("[uv_break]"):
                    text[ti] += " [uv_break]"
            lora_path = kwargs.get("lora_path", None)
            if lora_path:
                self.logger.info(f"load lora into gpt: {lora_path}")
                gpt.load_lora(lora_path)
            try:
                for result in self._infer_code(text, stream, use_decoder, params_infer_code):
                    yield result._wavs
            finally:
                if lora_path:
                    self.logger.info("unload lora !")
                    gpt.unload_lora()

    @staticmethod
    def _load_audio_24k(path: str) -> torch.Tensor:
        """torchaudio.load + resample to 24 kHz + channel mean (chattts_plus_pipeline.py:495-498).  torchaudio's file IO needs an
        optional codec package; plain PCM WAV files are read with the standard library when it is absent."""
        import torchaudio
        try:
            wav, sr = torchaudio.load(path)
        except Exception:
            import wave
            with wave.open(path, "rb") as f:
                sr, ch, sw, n = f.getframerate(), f.getnchannels(), f.getsampwidth(), f.getnframes()
                raw = f.readframes(n)
            if sw == 2:
                a = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
            elif sw == 4:
                a = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
            elif sw == 1:
                a = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
            else:
                raise ValueError(f"{path}: unsupported PCM sample width {sw}")
            wav = torch.from_numpy(a.reshape(-1, ch).T.copy())
        wav = torchaudio.functional.resample(wav, orig_freq=sr, new_freq=24000)
        return torch.mean(wav, 0)

    @torch.inference_mode()
    def sample_audio_speaker(self, wav) -> str:
        """chattts_plus_pipeline.py:279-284: waveform (24 kHz, 1-D) -> b14/LZMA string of the GFSQ prompt codes [num_vq, T]."""
        from .tokenizer import Tokenizer
        if isinstance(wav, np.ndarray):
            wav = torch.from_numpy(wav)
        if "dvae_encode" not in self.models_dict:
            raise RuntimeError("zero-shot speaker prompts need the dvae_encode model (configs/infer/chattts_plus.yaml)")
        ids = self.models_dict["dvae_encode"](wav.to(self.device, torch.float32)[None], "encode").squeeze_(0)
        return Tokenizer._encode_prompt(ids)

    @torch.no_grad()
    def infer(self, text, stream=False, lang=None, skip_refine_text=False, refine_text_only=False, use_decoder=True,
              do_text_normalization=True, do_text_optimization=True, do_homophone_replacement=True,
              params_refine_text=RefineTextParams(), params_infer_code=InferCodeParams(), **kwargs):
        if kwargs.get("speaker_audio_path", None):
            # zero-shot speaker prompt (chattts_plus_pipeline.py:486-500): the audio is encoded to GFSQ codes by the DVAE encoder
            # and rides in the prompt (tokenizer.encode(prompt_str=...)); no speaker embedding is applied
            speaker_audio_path = kwargs["speaker_audio_path"]
            assert os.path.exists(speaker_audio_path), f"speaker_audio_path {speaker_audio_path} not exists!"
            speaker_audio_text = kwargs.get("speaker_audio_text", "")
            self.logger.info("Use zero shot >>>")
            self.logger.info(f"speaker_audio_path is {speaker_audio_path}")
            self.logger.info(f"speaker_audio_text is {speaker_audio_text}")
            audio_wav = self._load_audio_24k(speaker_audio_path)
            params_infer_code.txt_smp = speaker_audio_text
            params_infer_code.spk_smp = self.sample_audio_speaker(audio_wav)
            params_infer_code.spk_emb = None
        elif kwargs.get("speaker_emb_path", None):
            p = kwargs["speaker_emb_path"]
            assert os.path.exists(p), f"speaker_emb_path {p} not exists!"
            self.logger.info(f"loading speaker_emb from {p}")
            params_infer_code.spk_emb = self._load_speaker(p)
        else:
            self.logger.info("speaker_emb is None, random select a speaker!")
            speaker_emb = self.sample_random_speaker()
            params_infer_code.spk_emb = speaker_emb
            spk_dir = kwargs.get("speaker_save_dir", os.path.join(constants.PROJECT_DIR, "results/speakers"))
            os.makedirs(spk_dir, exist_ok=True)
            torch.save(speaker_emb, f"{spk_dir}/{time.time()}.pt")
        return self._infer(text, stream, lang, skip_refine_text, refine_text_only, use_decoder, do_text_normalization,
                           do_text_optimization, do_homophone_replacement, params_refine_text, params_infer_code, **kwargs)
