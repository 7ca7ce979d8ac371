"""``gsv_tts.TTS`` on the B200 backend: the reference's constructor and model-management surface
(reference gsv_tts/TTS.py:38-147, 1164-1290) over the two native hot paths.

What is here: the constructor keywords, ``load_gpt_model`` / ``load_sovits_model`` / ``unload_*`` /
``get_*_list``, and ``infer_features`` / ``infer_features_batched``: the hot-path halves of ``infer`` /
``infer_batched`` (TTS.py:232-247, 695-764) for callers that already hold phoneme ids, BERT features, prompt
tokens and the vocoder latents.

What is not: the text front end (G2P, language segmentation, BERT / HuBERT / speaker-embedding featurisers)
and ``enc_p`` are outside the scope of this package.  ``infer`` /
``infer_stream`` / ``infer_batched`` therefore need a ``frontend`` object that supplies those pieces and raise a
clear error without one -- they never fall back to a CPU path.
"""
from __future__ import annotations

import logging
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import Loader
from .Config import Config
from .Player import AudioClip

log = logging.getLogger(__name__)


def cut_text(text: str, max_len: int = 50) -> List[str]:
    """Sentence cutting on terminal punctuation, merged up to ``max_len`` characters (the reference cuts with
    pysbd per language, TextProcessor.py:18-59; this is the dependency-free equivalent used by tests)."""
    import re
    parts = [p for p in re.split(r"(?<=[。！？!?.;；\n])", text) if p.strip()]
    out, cur = [], ""
    for p in parts:
        if cur and len(cur) + len(p) > max_len:
            out.append(cur)
            cur = ""
        cur += p
    if cur:
        out.append(cur)
    return out


class FrontendRequired(RuntimeError):
    pass


class TTS:
    def __init__(self, gpt_cache=None, sovits_cache=None, models_dir: Optional[str] = None, device=None, dtype=None,
                 use_flash_attn: bool = False, use_bert: bool = False, auto_bert: bool = True, use_jieba_fast: bool = True,
                 always_load_cnhubert: bool = False, always_load_sv: bool = False, frontend=None):
        cfg = Config()
        if gpt_cache is not None:
            cfg.gpt_cache = [tuple(x) for x in gpt_cache]
        if sovits_cache is not None:
            cfg.sovits_cache = list(sovits_cache)
        if device is not None:
            cfg.device = torch.device(device)
        if dtype is not None:
            cfg.dtype = getattr(torch, dtype) if isinstance(dtype, str) else dtype
        cfg.use_flash_attn, cfg.use_bert = use_flash_attn, use_bert
        if cfg.device.type != "cuda":
            raise RuntimeError("gsv_tts (B200 backend) needs an sm_100 CUDA device; there is no CPU path")
        self.tts_config = cfg
        self.models_dir = models_dir
        self.frontend = frontend
        self.gpt_models: Dict[str, Loader.Gpt] = {}
        self.sovits_models: Dict[str, Loader.Sovits] = {}
        self.samplerate, self.gpt_hz, self.sovits_hz = cfg.samplerate, cfg.gpt_hz, cfg.sovits_hz

    # ---- model management (TTS.py:1164-1290) -------------------------------------------------------------
    def load_gpt_model(self, *paths: str):
        for p in paths:
            if p not in self.gpt_models:
                self.gpt_models[p] = Loader.get_gpt_weights(p, self.tts_config)
                log.info("loaded GPT model %s", p)

    def load_sovits_model(self, *paths: str):
        for p in paths:
            if p not in self.sovits_models:
                self.sovits_models[p] = Loader.get_sovits_weights(p, self.tts_config)
                log.info("loaded SoVITS model %s", p)

    def unload_gpt_model(self, *paths: str):
        for p in paths:
            self.gpt_models.pop(p, None)

    def unload_sovits_model(self, *paths: str):
        for p in paths:
            self.sovits_models.pop(p, None)

    def get_gpt_list(self):
        return list(self.gpt_models)

    def get_sovits_list(self):
        return list(self.sovits_models)

    def _pick(self, table, path, what):
        if path is None:
            if not table:
                raise RuntimeError(f"no {what} model loaded")
            path = next(iter(table))
        if path not in table:
            raise KeyError(f"{what} model {path!r} is not loaded")
        return table[path]

    # ---- hot-path entry points ---------------------------------------------------------------------------------
    @torch.inference_mode()
    def infer_features(self, phoneme_ids, bert, prompt_tokens, z_p, y_mask, ge, gpt_model: Optional[str] = None,
                       sovits_model: Optional[str] = None, top_k=15, top_p=1.0, temperature=1.0, repetition_penalty=1.35):
        """GPT decode then flow + HiFi-GAN for ONE utterance whose features are already computed: returns
        ``(semantic tokens int64 [1,1,N], AudioClip)``.  ``z_p`` / ``y_mask`` / ``ge`` are what the reference's
        ``enc_p`` hands to ``flow_dec`` (models.py:404-406)."""
        gpt = self._pick(self.gpt_models, gpt_model, "GPT").t2s_model
        voc = self._pick(self.sovits_models, sovits_model, "SoVITS").vq_model
        tokens = gpt.infer(phoneme_ids, prompt_tokens, bert, top_k=top_k, top_p=top_p, temperature=temperature,
                           repetition_penalty=repetition_penalty)
        audio = voc.flow_dec(z_p, y_mask, ge)[0, 0].float().cpu().numpy()
        return tokens, self._clip(audio)

    def infer_features_stream(self, phoneme_ids, bert, prompt_tokens, features_of_chunk, gpt_model: Optional[str] = None,
                              sovits_model: Optional[str] = None, top_k=15, top_p=1.0, temperature=1.0,
                              repetition_penalty=1.35, stream_chunk: int = 25, decode_sms: int = 128,
                              force_steps: Optional[int] = None):
        """Streaming counterpart of ``infer_features`` (the GPT / vocoder part of ``infer_stream``, TTS.py:402-470):
        yields one ``AudioClip`` per chunk of ``stream_chunk`` semantic tokens.  ``features_of_chunk(tokens, final)``
        stands for ``enc_p`` (out of scope): it returns ``(z_p, y_mask, ge)`` for the frames of that chunk.
        The reference runs chunk c's vocoder and chunk c+1's decode back to back; here the decode of chunk c+1 is
        already in flight (``Text2SemanticDecoder.infer_stream`` launches ahead, on ``decode_sms`` SMs) while the
        vocoder of chunk c runs on a second stream."""
        gpt = self._pick(self.gpt_models, gpt_model, "GPT").t2s_model
        voc = self._pick(self.sovits_models, sovits_model, "SoVITS").vq_model
        dev = gpt._device
        side = torch.cuda.Stream(dev)
        gpt.set_decode_sms(decode_sms)
        try:
            it = gpt.infer_stream(phoneme_ids, prompt_tokens, bert, top_k=top_k, top_p=top_p, temperature=temperature,
                                  repetition_penalty=repetition_penalty, stream_chunk=stream_chunk, force_steps=force_steps)
            while True:
                with torch.inference_mode():
                    try:
                        tokens, final = next(it)          # host-synchronised: the tokens exist, the next chunk is decoding
                    except StopIteration:
                        break
                    with torch.cuda.stream(side):
                        side.wait_event(gpt.chunk_ready)                     # the token copy, not the decode behind it
                        tokens.record_stream(side)
                        z_p, y_mask, ge = features_of_chunk(tokens, final)
                        audio = voc.flow_dec(z_p, y_mask, ge)[0, 0].float().cpu()        # waits for `side` only
                yield self._clip(audio.numpy())
        finally:
            gpt.set_decode_sms(0)

    @torch.inference_mode()
    def infer_features_batched(self, phoneme_ids: Sequence, bert: Sequence, prompt_tokens: Sequence,
                               gpt_model: Optional[str] = None, top_k=15, top_p=1.0, temperature=1.0, max_new=None):
        """Continuous-batched GPT stage of ``infer_batched`` (TTS.py:695-703): token lists in request order."""
        gpt = self._pick(self.gpt_models, gpt_model, "GPT").t2s_model
        outs, order = gpt.infer_batched(list(phoneme_ids), list(prompt_tokens), list(bert), top_k=top_k, top_p=top_p,
                                        temperature=temperature, max_new=max_new)
        res = [None] * len(outs)
        for o, i in zip(outs, order.tolist()):
            res[i] = o
        return res

    @torch.inference_mode()
    def vocode_features_batched(self, z_p: Sequence[torch.Tensor], ge: Sequence[torch.Tensor], sovits_model: Optional[str] = None,
                                max_frames: int = 8192) -> List[AudioClip]:
        """Vocoder stage of ``infer_batched`` (TTS.py:705-764): utterances of different lengths ``z_p[i]`` [192, T_i] are
        sorted by length and run through flow + HiFi-GAN in padded groups of at most ``max_frames`` frames (padded frames
        are masked), so that a group is one native call; clips come back in request order."""
        voc = self._pick(self.sovits_models, sovits_model, "SoVITS").vq_model
        dev, dt = voc._device, voc._dtype
        order = sorted(range(len(z_p)), key=lambda i: -int(z_p[i].shape[-1]))
        clips: List[Optional[AudioClip]] = [None] * len(z_p)
        spf = voc.samples_per_frame
        i = 0
        while i < len(order):
            tmax = int(z_p[order[i]].shape[-1])
            n = max(1, min(len(order) - i, max_frames // max(tmax, 1)))
            grp = order[i:i + n]
            zb = torch.zeros(n, voc.inter_channels, tmax, device=dev, dtype=dt)
            mb = torch.zeros(n, 1, tmax, device=dev, dtype=dt)
            gb = torch.empty(n, voc.gin_channels, 1, device=dev, dtype=dt)
            for r, k in enumerate(grp):
                t = int(z_p[k].shape[-1])
                zb[r, :, :t] = z_p[k].to(device=dev, dtype=dt, non_blocking=True)
                mb[r, :, :t] = 1
                gb[r] = ge[k].to(device=dev, dtype=dt, non_blocking=True).view(voc.gin_channels, 1)
            audio = voc.flow_dec(zb, mb, gb).float().cpu().numpy()
            for r, k in enumerate(grp):
                clips[k] = self._clip(audio[r, 0, : int(z_p[k].shape[-1]) * spf])
            i += n
        return clips

    def _clip(self, audio: np.ndarray, text: str = "") -> AudioClip:
        peak = float(np.abs(audio).max()) if audio.size else 0.0
        if peak > 1.0:
            audio = audio / peak                                    # TTS.py:276-278
        return AudioClip(audio, self.samplerate, orig_text=text)

    # ---- text entry points need the front end -----------------------------------------------------------------------
    def _need_frontend(self, name):
        if self.frontend is None:
            raise FrontendRequired(
                f"TTS.{name} needs the text / audio front end (G2P, BERT, HuBERT, speaker embedding, enc_p), which is "
                "outside this package's scope (SURVEY.md 2 rows 7-9): pass frontend=..., or call infer_features().")
        return self.frontend

    def infer(self, *args, **kwargs):
        return self._need_frontend("infer").infer(self, *args, **kwargs)

    def infer_stream(self, *args, **kwargs):
        return self._need_frontend("infer_stream").infer_stream(self, *args, **kwargs)

    def infer_batched(self, *args, **kwargs):
        return self._need_frontend("infer_batched").infer_batched(self, *args, **kwargs)
