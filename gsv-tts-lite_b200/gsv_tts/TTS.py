"""``gsv_tts.TTS`` on the B200 backend: the reference's constructor and model-management surface
(reference gsv_tts/TTS.py:38-147, 1164-1290) over the two native hot paths.

What is here: the constructor keywords, ``load_gpt_model`` / ``load_sovits_model`` / ``unload_*`` /
``get_*_list``, and ``infer_features`` / ``infer_features_batched``: the hot-path halves of ``infer`` /
``infer_batched`` (TTS.py:232-247, 695-764) for callers that already hold phoneme ids, BERT features, prompt
tokens and the vocoder latents.

``infer_phones`` / ``infer_phones_stream`` are the reference's ``infer`` / ``infer_stream`` from the point where the text
front end has done its work (phoneme ids, BERT rows, prompt tokens, speaker embedding): GPT decode, ``vq_model.decode``
(prior encoder + flow + HiFi-GAN), monotonic alignment, silence trimming, SOLA splicing -- all on the device.

What is not here: the text front end itself (G2P, language segmentation, BERT / HuBERT / speaker-embedding featurisers;
SURVEY.md 2 rows 7-9).  ``infer`` / ``infer_stream`` / ``infer_batched`` keep the reference's signatures and take those
pieces from a ``frontend`` plug-in (three methods, see ``TTS.infer``); without one they raise -- they never fall back to
a CPU path.
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

    def to_safetensors(self, checkpoint_path: str, output_dir: Optional[str] = None):
        """TTS.py:1482-1523: convert a ``.ckpt`` / ``.pth`` checkpoint into the safetensors directory format the loaders read."""
        out = Loader.to_safetensors(checkpoint_path, output_dir)
        log.info("Successfully converted and saved to: %s", out)

    def get_gpt_list(self):
        return list(self.gpt_models)

    def get_sovits_list(self):
        return list(self.sovits_models)

    def _side_stream(self, dev) -> "torch.cuda.Stream":
        """The stream the SoVITS stage of a streaming call runs on, created ONCE per TTS object: torch's caching allocator keeps
        its free blocks per stream, so a fresh stream per call makes every tensor of the call's first chunk a cudaMalloc --
        which waits for the running decode kernel each time (measured: 20-60 ms on the first clip of an utterance, at random)."""
        cache = self.__dict__.setdefault("_side_streams", {})
        key = str(dev)
        if key not in cache:
            cache[key] = torch.cuda.Stream(dev)
        return cache[key]

    def _lane_streams(self, dev, n: int) -> List["torch.cuda.Stream"]:
        """Streams of the prior-encoder lanes of a batched SoVITS stage (created once, like ``_side_stream``)."""
        cache = self.__dict__.setdefault("_lane_stream_cache", {})
        lst = cache.setdefault(str(dev), [])
        while len(lst) < n:
            lst.append(torch.cuda.Stream(dev))
        return lst[:n]

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
        side = self._side_stream(dev)
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
                        gpt.hold_until_decode_resident(side)                 # ... but let that decode take its SMs first
                        tokens.record_stream(side)
                        z_p, y_mask, ge = features_of_chunk(tokens, final)
                        audio = voc.flow_dec(z_p, y_mask, ge)[0, 0].float().cpu()        # waits for `side` only
                yield self._clip(audio.numpy())
        finally:
            gpt.set_decode_sms(0)

    @torch.inference_mode()
    def infer_features_batched(self, phoneme_ids: Sequence, bert: Sequence, prompt_tokens: Sequence,
                               gpt_model: Optional[str] = None, top_k=15, top_p=1.0, temperature=1.0, max_new=None,
                               queue_order: Optional[str] = None):
        """Continuous-batched GPT stage of ``infer_batched`` (TTS.py:695-703): token lists in request order.
        ``queue_order="longest_first"`` hands the requests to the slot scheduler longest predicted first (``max_new[i]`` where
        given, else the phoneme count): the batch then drains with short requests, not with a few long ones holding a handful
        of slots while the others stand empty (longest-processing-time-first; on BASELINE config 3 the ideal schedule is 609
        decode steps instead of 732).  Results come back in request order either way; a request's noise stream is keyed by its
        place in the queue, so its sampled tokens differ between the two orders (both are samples of the same distribution)."""
        gpt = self._pick(self.gpt_models, gpt_model, "GPT").t2s_model
        n = len(phoneme_ids)
        perm = list(range(n))
        if queue_order == "longest_first":
            def predicted(i):
                if max_new is not None and int(max_new[i]) > 0:
                    return int(max_new[i])
                return 10 ** 9 + len(phoneme_ids[i])              # no cap given: the longer text first
            perm.sort(key=lambda i: -predicted(i))                # stable: ties keep the caller's order
        elif queue_order not in (None, "fifo"):
            raise ValueError(f"queue_order must be None, 'fifo' or 'longest_first', not {queue_order!r}")
        outs, order = gpt.infer_batched([phoneme_ids[i] for i in perm], [prompt_tokens[i] for i in perm], [bert[i] for i in perm],
                                        top_k=top_k, top_p=top_p, temperature=temperature,
                                        max_new=[max_new[i] for i in perm] if max_new is not None else None)
        res = [None] * n
        for o, i in zip(outs, order.tolist()):
            res[perm[i]] = o
        return res

    def _vocode_groups(self, voc, z_p: Sequence[torch.Tensor], ge: Sequence[torch.Tensor], max_frames: int) -> List[torch.Tensor]:
        """Utterances of different lengths ``z_p[i]`` [192, T_i] through flow + HiFi-GAN in length-sorted, padded groups of at
        most ``max_frames`` frames (padded frames are masked; a group is one native call).  Returns fp32 waveforms on the
        device, in input order -- no host synchronisation."""
        dev, dt = voc._device, voc._dtype
        order = sorted(range(len(z_p)), key=lambda i: -int(z_p[i].shape[-1]))
        out: List[Optional[torch.Tensor]] = [None] * len(z_p)
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
            audio = voc.flow_dec(zb, mb, gb).float()
            for r, k in enumerate(grp):
                out[k] = audio[r, 0, : int(z_p[k].shape[-1]) * spf]
            i += n
        return out

    @torch.inference_mode()
    def vocode_features_batched(self, z_p: Sequence[torch.Tensor], ge: Sequence[torch.Tensor], sovits_model: Optional[str] = None,
                                max_frames: int = 8192) -> List[AudioClip]:
        """Vocoder stage of ``infer_batched`` (TTS.py:705-764) for latents of different lengths; clips in request order."""
        voc = self._pick(self.sovits_models, sovits_model, "SoVITS").vq_model
        return [self._clip(a.cpu().numpy()) for a in self._vocode_groups(voc, z_p, ge, max_frames)]

    PRIOR_LANES = int(__import__('os').environ.get('GSV_PRIOR_LANES', '4'))

    def _sovits_stage_device(self, vq, tokens, phones2, ge, noise_scale, speed, max_frames, lanes: Optional[int] = None) -> List[torch.Tensor]:
        """SoVITS stage on the CURRENT stream, no host synchronisation: the prior encoder per utterance (its attention must not
        cross utterance boundaries), then flow + HiFi-GAN in padded, length-sorted groups; fp32 waveforms on the device.
        One prior-encoder call is ~100 small dependent launches (0.9-2 ms whatever the GPU could do beside it), so the
        utterances are dealt over ``lanes`` native contexts on as many streams, forked from and joined back into the current
        stream; per utterance the arithmetic is that of a single call (tested bit for bit against ``lanes=1``)."""
        dev = vq._device
        work = [i for i, t in enumerate(tokens) if t.numel() > 0]           # EOS as the first token: nothing to say
        n_lanes = max(1, min(self.PRIOR_LANES if lanes is None else int(lanes), len(work)))
        zs, gs = [], []
        if n_lanes == 1:
            for i in work:
                z_p, _mask, g2, _attn = vq.prior(tokens[i].view(1, 1, -1), phones2[i].view(1, -1), ge[i], noise_scale=noise_scale, speed=speed)
                zs.append(z_p[0])
                gs.append(g2[0])
        else:
            cur = torch.cuda.current_stream(dev)
            streams = self._lane_streams(dev, n_lanes)
            fork = torch.cuda.Event()
            fork.record(cur)
            for lane, st in enumerate(streams):
                st.wait_event(fork)
            for k, i in enumerate(work):
                lane = k % n_lanes
                with torch.cuda.stream(streams[lane]):
                    for t in (tokens[i], phones2[i], ge[i]):
                        if t.is_cuda:
                            t.record_stream(streams[lane])
                    z_p, _mask, g2, _attn = vq.prior(tokens[i].view(1, 1, -1), phones2[i].view(1, -1), ge[i], noise_scale=noise_scale, speed=speed,
                                                     lane=lane)
                    for t in (z_p, g2):
                        t.record_stream(cur)
                zs.append(z_p[0])
                gs.append(g2[0])
            for st in streams:
                cur.wait_stream(st)
        out = [torch.zeros(0, device=dev) for _ in tokens]
        for i, a in zip(work, self._vocode_groups(vq, zs, gs, max_frames) if work else []):
            out[i] = a
        return out

    @torch.inference_mode()
    def decode_batched(self, tokens: Sequence[torch.Tensor], phones2: Sequence[torch.Tensor], ge: Sequence[torch.Tensor],
                       noise_scale: float = 0.5, speed: float = 1.0, sovits_model: Optional[str] = None, max_frames: int = 8192,
                       lanes: Optional[int] = None):
        """SoVITS stage of ``infer_batched`` for a list of utterances.  tokens[i] int64 [N_i], phones2[i] int64 [Nt_i],
        ge[i] [1,gin,1] -> AudioClips."""
        vq = self._pick(self.sovits_models, sovits_model, "SoVITS").vq_model
        return [self._clip(a.cpu().numpy()) for a in self._sovits_stage_device(vq, tokens, phones2, ge, noise_scale, speed, max_frames, lanes)]

    @torch.inference_mode()
    def infer_phones_batched(self, phoneme_ids: Sequence, bert: Sequence, prompt_tokens: Sequence, phones2: Sequence, ge: Sequence,
                             top_k=15, top_p=1.0, temperature=1.0, noise_scale: float = 0.5, speed: float = 1.0, max_new=None,
                             gpt_model: Optional[str] = None, sovits_model: Optional[str] = None, max_frames: int = 8192,
                             overlap: bool = True, trim_silence: bool = False, texts: Optional[Sequence[str]] = None):
        """Everything of ``TTS.infer_batched`` after the text front end (TTS.py:695-764): continuous-batched GPT, then the SoVITS
        stage per utterance.  The reference runs the two stages back to back; here the SoVITS stage of the requests harvested
        at one read is enqueued on the second stream right behind the next decode launch (held until that launch is resident,
        ``hold_until_decode_resident``), so it runs on the SMs the batched decode kernel leaves free while the other slots keep
        decoding.  Returns ``(token lists, AudioClips)`` in request order.  ``overlap=False`` is the back-to-back order."""
        gpt = self._pick(self.gpt_models, gpt_model, "GPT").t2s_model
        vq = self._pick(self.sovits_models, sovits_model, "SoVITS").vq_model
        dev = self.tts_config.device
        n = len(phoneme_ids)
        tokens: List[Optional[torch.Tensor]] = [None] * n
        audio: List[Optional[torch.Tensor]] = [None] * n
        ready: List[int] = []
        side = self._side_stream(dev) if overlap else None

        def on_finish(r, toks):
            tokens[r] = toks
            ready.append(r)

        def run_ready(decoding: bool):
            if not ready or side is None:
                return
            batch = list(ready)
            del ready[:]
            side.wait_stream(torch.cuda.current_stream(dev))               # the harvested tokens were copied on the caller's stream
            with torch.cuda.stream(side):
                if decoding:
                    gpt.hold_until_decode_resident(side)
                for r in batch:
                    tokens[r].record_stream(side)
                outs = self._sovits_stage_device(vq, [tokens[r] for r in batch], [phones2[r] for r in batch], [ge[r] for r in batch],
                                                 noise_scale, speed, max_frames)
            for r, a in zip(batch, outs):
                audio[r] = a

        gpt.infer_batched(list(phoneme_ids), list(prompt_tokens), list(bert), top_k=top_k, top_p=top_p, temperature=temperature,
                          max_new=max_new, on_finish=on_finish, on_launch=run_ready if overlap else None)
        if overlap:
            run_ready(False)                                               # what finished at the last read
            torch.cuda.current_stream(dev).wait_stream(side)
        else:
            outs = self._sovits_stage_device(vq, tokens, phones2, ge, noise_scale, speed, max_frames)
            audio = list(outs)
        clips = []
        for i, a in enumerate(audio):
            if trim_silence:                                               # TTS.py:803-812: leading / trailing silence of every utterance
                a = a[self._find_head_threshold_offsets(a):a.numel() - self._find_tail_threshold_offsets(a)]
            clips.append(self._clip(a.cpu().numpy(), texts[i] if texts is not None else ""))
        return tokens, clips

    def _clip(self, audio: np.ndarray, text: str = "") -> AudioClip:
        peak = float(np.abs(audio).max()) if audio.size else 0.0
        if peak > 1.0:
            audio = audio / peak                                    # TTS.py:276-278
        return AudioClip(audio, self.samplerate, orig_text=text)

    # ---- host glue of infer / infer_stream, on the device (reference TTS.py:1612-1797; csrc/glue.cu) ---------------------------
    def _native_dtype(self, t: torch.Tensor):
        from . import _native as N
        return N, N.dtype_code(t.dtype)

    def _viterbi_monotonic(self, attn: torch.Tensor) -> torch.Tensor:
        """TTS.py:1744-1797: attn [H, T, N] -> int64 [T] monotonic text index per frame (-1 before the first aligned frame).
        One launch instead of a Python loop over the frames."""
        import ctypes as C
        from . import _native as N
        attn = attn.to(torch.float32).contiguous()
        H, T, Nn = attn.shape
        work = torch.empty(T * Nn * 5, dtype=torch.uint8, device=attn.device)
        assign = torch.empty(T, dtype=torch.int32, device=attn.device)
        st = C.c_void_p(torch.cuda.current_stream(attn.device).cuda_stream)
        N.check(N.lib().gsv_glue_viterbi_monotonic(attn.data_ptr(), H, T, Nn, work.data_ptr(), assign.data_ptr(), st))
        return assign.to(torch.int64)

    def _silence_offset_enqueue(self, audio: torch.Tensor, tail: bool, threshold, frame_length, hop_length, search_len, margin) -> torch.Tensor:
        """The search kernel enqueued on the current stream; the offset stays on the device (int32 [1])."""
        import ctypes as C
        N, code = self._native_dtype(audio)
        audio = audio.contiguous()
        work = torch.empty(2, dtype=torch.int32, device=audio.device)
        out = torch.empty(1, dtype=torch.int32, device=audio.device)
        st = C.c_void_p(torch.cuda.current_stream(audio.device).cuda_stream)
        N.check(N.lib().gsv_glue_silence_offset(audio.data_ptr(), audio.numel(), code, 1 if tail else 0, float(threshold), frame_length,
                                                hop_length, search_len, margin, work.data_ptr(), out.data_ptr(), st))
        return out

    def _silence_offset(self, audio: torch.Tensor, tail: bool, threshold, frame_length, hop_length, search_len, margin) -> int:
        return int(self._silence_offset_enqueue(audio, tail, threshold, frame_length, hop_length, search_len, margin).item())

    def _find_head_threshold_offsets(self, audio, threshold=0.02, frame_length=512, hop_length=256, search_len=64000, margin=3200):
        """TTS.py:1630-1645 (frame RMS in fp32 here; the reference squares and averages in the 16-bit storage type)."""
        return self._silence_offset(audio, False, threshold, frame_length, hop_length, search_len, margin)

    def _find_tail_threshold_offsets(self, audio, threshold=0.01, frame_length=512, hop_length=256, search_len=64000, margin=3200):
        """TTS.py:1647-1662."""
        return self._silence_offset(audio, True, threshold, frame_length, hop_length, search_len, margin)

    def _sola_enqueue(self, f1_overlap, f2, overlap_len, search_len: int = 320):
        """The SOLA kernels enqueued on the current stream, nothing read back: -> (buffer [n2] whose first n2 - offset samples
        are f2 aligned and cross-faded, offset int32 [1] on the device)."""
        import ctypes as C
        N, code = self._native_dtype(f2)
        f1, f2c = f1_overlap.reshape(-1).contiguous(), f2.reshape(-1).contiguous()
        n2 = f2c.numel()
        work = torch.empty(search_len + 1, dtype=torch.float32, device=f2.device)
        off = torch.empty(1, dtype=torch.int32, device=f2.device)
        out = torch.empty(n2, dtype=f2.dtype, device=f2.device)
        st = C.c_void_p(torch.cuda.current_stream(f2.device).cuda_stream)
        N.check(N.lib().gsv_glue_sola(f1.data_ptr(), f2c.data_ptr(), n2, int(overlap_len), int(search_len), code, work.data_ptr(),
                                      off.data_ptr(), out.data_ptr(), st))
        return out, off

    def _sola_algorithm(self, f1_overlap, f2, overlap_len, search_len: int = 320):
        """TTS.py:1612-1628: f1_overlap [1,1,ov], f2 [1,1,n] -> (f2 aligned and cross-faded [1,1,n - offset], offset tensor)."""
        out, off = self._sola_enqueue(f1_overlap, f2, overlap_len, search_len)
        k = int(off.item())
        return out[: out.numel() - k].view(1, 1, -1), off

    def _get_subtitles(self, word2ph, assign, speed, last_end_s=0):
        """TTS.py:1664-1707: frame alignment -> word timings (host arithmetic on T small integers)."""
        frame_time = (1 / self.sovits_hz) / speed
        assign = assign.tolist() if hasattr(assign, "tolist") else list(assign)
        ph_end_s, cur = [], int(assign[0])
        for f in range(1, len(assign)):
            if int(assign[f]) != cur:
                ph_end_s.append(f * frame_time)
                cur = int(assign[f])
        ph_end_s.append(len(assign) * frame_time)
        idx = -1
        end_s = last_end_s + ph_end_s.pop(0) if assign[0] == -1 else last_end_s
        subtitles, word = [], None
        for i in range(len(word2ph["word"])):
            word, ph = word2ph["word"][i], word2ph["ph"][i]
            idx += ph
            if idx >= len(ph_end_s):
                break
            start_s, end_s = end_s, ph_end_s[idx] + last_end_s
            subtitles.append({"text": word, "start_s": start_s, "end_s": end_s})
        if end_s - last_end_s != ph_end_s[-1]:
            start_s, end_s = end_s, ph_end_s[-1] + last_end_s
            subtitles.append({"text": word, "start_s": start_s, "end_s": end_s})
        return subtitles

    @staticmethod
    def _increment_subtitle_times(subtitles, increment):
        for sub in subtitles:
            sub["start_s"] += increment
            if sub["end_s"]:
                sub["end_s"] += increment

    # ---- the reference's infer / infer_stream from the point where the front end has produced its features ---------------------
    @torch.inference_mode()
    def infer_phones(self, phones1: Sequence[int], bert1: torch.Tensor, prompt: torch.Tensor, phones2: Sequence[int], bert2: torch.Tensor,
                     ge: torch.Tensor, word2ph: Optional[dict] = None, text: str = "", return_subtitles: bool = False, top_k: int = 15,
                     top_p: float = 1.0, temperature: float = 1.0, repetition_penalty: float = 1.35, noise_scale: float = 0.5,
                     speed: float = 1.0, gpt_model: Optional[str] = None, sovits_model: Optional[str] = None) -> AudioClip:
        """``TTS.infer`` (TTS.py:232-286) after ``_prepare_*_resources`` and ``get_phones_and_bert``: GPT decode ->
        ``vq_model.decode`` -> monotonic alignment (-> subtitles) -> leading-silence trim -> peak normalisation -> 0.2 s tail.
        phones1 / bert1 / prompt: the cached prompt's phoneme ids, BERT rows [N1,1024] and semantic tokens [1,Ny]; phones2 /
        bert2: the target text's; ge: the speaker embedding [1,gin,1]."""
        gpt = self._pick(self.gpt_models, gpt_model, "GPT").t2s_model
        vq = self._pick(self.sovits_models, sovits_model, "SoVITS").vq_model
        dev = self.tts_config.device
        ids = torch.tensor(list(phones1) + list(phones2), dtype=torch.int64, device=dev).unsqueeze(0)
        bert = torch.cat([bert1.to(dev), bert2.to(dev)]).unsqueeze(0)
        pred = gpt.infer(ids, prompt, bert, top_k=top_k, top_p=top_p, temperature=temperature, repetition_penalty=repetition_penalty)
        ph2 = torch.tensor(list(phones2), dtype=torch.int64, device=dev).unsqueeze(0)
        audio, attn = vq.decode(pred, ph2, ge, noise_scale=noise_scale, speed=speed)
        audio = audio[0, 0, :]
        assign = self._viterbi_monotonic(attn)
        subtitles = []
        if return_subtitles and word2ph is not None:
            subtitles = self._get_subtitles(word2ph, assign, speed)
            subtitles[-1]["end_s"] += 0.2
        head = self._find_head_threshold_offsets(audio)
        audio = audio[head:]
        if subtitles:
            self._increment_subtitle_times(subtitles, -head / self.samplerate)
            subtitles[0]["start_s"] = max(0, subtitles[0]["start_s"])
        audio = audio.float().cpu().numpy()
        peak = float(np.abs(audio).max()) if audio.size else 0.0
        if peak > 1:
            audio = audio / peak
        audio = np.concatenate([audio, np.zeros((int(0.2 * self.samplerate),), dtype=audio.dtype)])
        return AudioClip(audio, self.samplerate, subtitles=subtitles, orig_text=text)

    def infer_phones_stream(self, phones1: Sequence[int], bert1: torch.Tensor, prompt: torch.Tensor, phones2: Sequence[int],
                            bert2: torch.Tensor, ge: torch.Tensor, stream_chunk: int = 25, overlap_len: int = 5,
                            boost_first_chunk: bool = True, cut_mute: float = 0.4, top_k: int = 15, top_p: float = 1.0,
                            temperature: float = 1.0, repetition_penalty: float = 1.35, noise_scale: float = 0.5, speed: float = 1.0,
                            gpt_model: Optional[str] = None, sovits_model: Optional[str] = None, force_steps: Optional[int] = None,
                            decode_ahead: bool = True):
        """One text cut of ``TTS.infer_stream`` (TTS.py:402-498): every ``stream_chunk`` semantic tokens the whole prefix goes
        through ``vq_model.decode(stream_mode=True)``; chunks are spliced with SOLA over ``overlap_len`` frames, the first one
        loses its leading silence, the last one gets ``cut_mute`` seconds of silence.  Yields one ``AudioClip`` per chunk.
        The decode of chunk c+1 is already in flight on the GPT's SMs while chunk c goes through the prior encoder and the
        vocoder on a second stream.

        The reference hands a chunk to the SoVITS stage one chunk late (t2s_model.py:540-547: only when the next one exists,
        so that a short remainder is merged into the last full chunk).  ``decode_ahead`` keeps that order of clips but starts
        a held-back chunk's SoVITS stage the moment its tokens exist (``on_chunk_held`` of ``infer_stream``), without any host
        synchronisation; when the stream yields the chunk its clip is already in pinned host memory.  If the stream ends
        before the next boundary the chunk decoded ahead is dropped and the cross-chunk state (``enc_p.y_overlap``, the SOLA
        tail, ``valid_start_idx``) is rolled back -- the clips are the same as with ``decode_ahead=False`` (tested)."""
        gpt = self._pick(self.gpt_models, gpt_model, "GPT").t2s_model
        vq = self._pick(self.sovits_models, sovits_model, "SoVITS").vq_model
        dev = self.tts_config.device
        side = self._side_stream(dev)
        with torch.inference_mode():
            ids = torch.tensor(list(phones1) + list(phones2), dtype=torch.int64, device=dev).unsqueeze(0)
            bert = torch.cat([bert1.to(dev), bert2.to(dev)]).unsqueeze(0)
            ph2 = torch.tensor(list(phones2), dtype=torch.int64, device=dev).unsqueeze(0)
            ge = ge.to(dev)
        overlap_samples = overlap_len * vq.samples_per_frame
        st = {"last_overlap": None, "valid_start": 0, "chunk_idx": 0}       # cross-chunk state of the SoVITS stage
        held = {}                                                             # id(tokens) -> (tokens, staged work, state before it)

        def stage(pred, is_final):
            """The SoVITS stage of one chunk enqueued on the side stream; nothing is read back here.  Offsets (SOLA alignment,
            leading silence) stay on the device: the buffers travel to pinned host memory whole and are cut in ``collect``."""
            with torch.inference_mode(), torch.cuda.stream(side):
                side.wait_event(gpt.chunk_ready)
                gpt.hold_until_decode_resident(side)
                pred.record_stream(side)
                audio, attn = vq.decode(pred, ph2, ge, noise_scale=noise_scale, speed=speed, stream_mode=True,
                                        valid_start_idx=st["valid_start"], overlap_len=overlap_len,
                                        text_unchanged=st["chunk_idx"] > 0)       # one utterance: the text branch of chunk 0 serves them all
                flat = audio.reshape(-1)
                n2 = flat.numel()
                meta = torch.zeros(2, dtype=torch.int32, device=dev)          # SOLA offset, leading-silence offset
                if st["last_overlap"] is not None:
                    flat, off = self._sola_enqueue(st["last_overlap"], flat, overlap_samples)
                    meta[0:1].copy_(off)
                    tail = (n2 - overlap_samples) - off.to(torch.int64) + torch.arange(overlap_samples, device=dev)
                    st["last_overlap"] = flat[tail]                          # the last overlap_samples of the n2 - offset valid ones
                else:
                    st["last_overlap"] = flat[n2 - overlap_samples:].clone()
                if not is_final:
                    st["valid_start"] = attn.shape[1] - overlap_len
                if st["chunk_idx"] == 0:                                     # never spliced: its length is known here
                    body = flat if is_final else flat[: n2 - overlap_samples]
                    meta[1:2].copy_(self._silence_offset_enqueue(body, False, 0.02, 512, 256, 64000, 3200))
                st["chunk_idx"] += 1
                host = torch.empty(n2, dtype=torch.float32, pin_memory=True)
                host.copy_(flat, non_blocking=True)
                host_meta = torch.empty(2, dtype=torch.int32, pin_memory=True)
                host_meta.copy_(meta, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
            return host, host_meta, ev, n2, is_final

        def collect(work) -> np.ndarray:
            host, host_meta, ev, n2, is_final = work
            ev.synchronize()
            end = n2 - int(host_meta[0]) - (0 if is_final else overlap_samples)
            out = host[int(host_meta[1]):end].numpy().copy()
            if is_final:
                out = np.concatenate([out, np.zeros(int(cut_mute * self.samplerate), dtype=out.dtype)])
            return out

        def on_held(pred):
            before = dict(st)
            held[id(pred)] = (pred, stage(pred, False), before)

        audio_len_s = 0.0
        it = gpt.infer_stream(ids, prompt, bert, top_k=top_k, top_p=top_p, temperature=temperature, repetition_penalty=repetition_penalty,
                              stream_chunk=stream_chunk, boost_first_chunk=boost_first_chunk, force_steps=force_steps,
                              on_chunk_held=on_held if decode_ahead else None)
        try:
            while True:
                with torch.inference_mode():
                    try:
                        pred, is_final = next(it)
                    except StopIteration:
                        break
                ahead = held.pop(id(pred), None)
                if ahead is not None:
                    work = ahead[1]
                else:
                    if held:                                   # the stream ended before the next boundary: drop the chunk decoded ahead
                        _, (_p, _w, before) = held.popitem()
                        st.update(before)
                        vq.enc_p.rollback()
                    work = stage(pred, is_final)
                out = collect(work)
                audio_len_s += len(out) / self.samplerate
                yield AudioClip(out, self.samplerate, audio_len_s=audio_len_s)
        finally:
            vq.enc_p.y_overlap = None                      # TTS.py:498

    # ---- text entry points: the front end (G2P, BERT, HuBERT, speaker embedding) is a plug-in ----------------------------------
    def _need_frontend(self, name):
        if self.frontend is None:
            raise FrontendRequired(
                f"TTS.{name} needs the text / audio front end (G2P, BERT, HuBERT, speaker embedding), which is outside this "
                "package's scope (SURVEY.md 2 rows 7-9): pass frontend=..., or call infer_phones() / infer_phones_stream().")
        return self.frontend

    def infer(self, spk_audio_path, prompt_audio_path: str, prompt_audio_text: str, text: str, return_subtitles: bool = False,
              top_k: int = 15, top_p: float = 1.0, temperature: float = 1.0, repetition_penalty: float = 1.35, noise_scale: float = 0.5,
              speed: float = 1.0, gpt_model: Optional[str] = None, sovits_model: Optional[str] = None) -> AudioClip:
        """``TTS.infer`` (TTS.py:150-286).  The front end supplies ``speaker(tts, sovits_model, spk_audio_path) -> ge``,
        ``prompt(tts, gpt_model, prompt_audio_path, prompt_audio_text) -> (prompt tokens, phones1, bert1)`` and
        ``phones_and_bert(text) -> (phones2, word2ph, bert2, norm_text)`` (reference _prepare_sovits_resources,
        _prepare_gpt_resources, get_phones_and_bert); everything after that runs here."""
        fe = self._need_frontend("infer")
        ge = fe.speaker(self, sovits_model, spk_audio_path)
        prompt, phones1, bert1 = fe.prompt(self, gpt_model, prompt_audio_path, prompt_audio_text)
        phones2, word2ph, bert2, _norm = fe.phones_and_bert(text)
        return self.infer_phones(phones1, bert1, prompt, phones2, bert2, ge, word2ph=word2ph, text=text, return_subtitles=return_subtitles,
                                 top_k=top_k, top_p=top_p, temperature=temperature, repetition_penalty=repetition_penalty,
                                 noise_scale=noise_scale, speed=speed, gpt_model=gpt_model, sovits_model=sovits_model)

    def infer_stream(self, spk_audio_path, prompt_audio_path: str, prompt_audio_text: str, text: str, cut_minlen: int = 10,
                     cut_mute: float = 0.4, stream_chunk: int = 25, overlap_len: int = 5, boost_first_chunk: bool = True, top_k: int = 15,
                     top_p: float = 1.0, temperature: float = 1.0, repetition_penalty: float = 1.35, noise_scale: float = 0.5,
                     speed: float = 1.0, gpt_model: Optional[str] = None, sovits_model: Optional[str] = None, **_ignored):
        """``TTS.infer_stream`` (TTS.py:289-504): the text is cut, every cut streams through ``infer_phones_stream``."""
        fe = self._need_frontend("infer_stream")
        ge = fe.speaker(self, sovits_model, spk_audio_path)
        prompt, phones1, bert1 = fe.prompt(self, gpt_model, prompt_audio_path, prompt_audio_text)
        total = 0.0
        for i, cut in enumerate(cut_text(text, cut_minlen)):
            phones2, _w2p, bert2, _norm = fe.phones_and_bert(cut)
            for clip in self.infer_phones_stream(phones1, bert1, prompt, phones2, bert2, ge, stream_chunk=stream_chunk,
                                                 overlap_len=overlap_len, boost_first_chunk=boost_first_chunk if i == 0 else False,
                                                 cut_mute=cut_mute, top_k=top_k, top_p=top_p, temperature=temperature,
                                                 repetition_penalty=repetition_penalty, noise_scale=noise_scale, speed=speed,
                                                 gpt_model=gpt_model, sovits_model=sovits_model):
                total += len(clip.audio_data) / self.samplerate
                clip.audio_len_s = total
                clip.orig_text = text
                yield clip

    @torch.inference_mode()
    def infer_batched(self, spk_audio_paths, prompt_audio_paths, prompt_audio_texts, texts, top_k: int = 15, top_p: float = 1.0,
                      temperature: float = 1.0, repetition_penalty: float = 1.35, noise_scale: float = 0.5, speed: float = 1.0,
                      gpt_model: Optional[str] = None, sovits_model: Optional[str] = None, **_ignored):
        """``TTS.infer_batched`` (TTS.py:507-868): every text goes through the continuous-batched GPT, and through the SoVITS
        stage one utterance at a time (the reference concatenates a SoVITS batch into ONE sequence whose encoder attends across
        utterance boundaries, TTS.py:730-764; decoding them separately is the per-utterance result the single path gives), the
        two stages overlapped (``infer_phones_batched``).  Returns a tuple of ``AudioClip`` in input order."""
        fe = self._need_frontend("infer_batched")
        dev = self.tts_config.device
        one = lambda v, i: v[i] if isinstance(v, (list, tuple)) else v
        ids, prompts, berts, ph2s, ges = [], [], [], [], []
        for i, text in enumerate(texts):
            ge = fe.speaker(self, sovits_model, one(spk_audio_paths, i))
            prompt, phones1, bert1 = fe.prompt(self, gpt_model, one(prompt_audio_paths, i), one(prompt_audio_texts, i))
            phones2, _w2p, bert2, _norm = fe.phones_and_bert(text)
            ids.append(torch.tensor(list(phones1) + list(phones2), dtype=torch.int64, device=dev))
            prompts.append(prompt.reshape(-1))
            berts.append(torch.cat([bert1.to(dev), bert2.to(dev)]))
            ph2s.append(torch.tensor(list(phones2), dtype=torch.int64, device=dev).unsqueeze(0))
            ges.append(ge)
        _toks, clips = self.infer_phones_batched(ids, berts, prompts, ph2s, ges, top_k=top_k, top_p=top_p, temperature=temperature,
                                                 noise_scale=noise_scale, speed=speed, gpt_model=gpt_model, sovits_model=sovits_model,
                                                 trim_silence=True, texts=list(texts))
        return tuple(clips)
