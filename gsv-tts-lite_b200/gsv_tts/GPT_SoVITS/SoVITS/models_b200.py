"""B200 backend of the SoVITS reverse flow + HiFi-GAN generator.

Host-side mirror of the ``flow`` / ``dec`` part of ``SynthesizerTrn`` (reference
gsv_tts/GPT_SoVITS/SoVITS/models.py:235-429): same hyper-parameter names, same state-dict
keys, ``initialize_runtime(dtype, device, sovits_cache)``, ``flow_dec(z_p, y_mask, ge)`` and
``samples_per_frame``.  The arithmetic runs in ``libgsv_b200.so`` (csrc/vocoder.cu).

What the host does once, at load: fold weight norm (the reference leaves ``flow``'s weight
norm parametrised and re-evaluates it every forward, Loader.py:73,95), transpose every kernel
to tap-major ``[k][Cout][Cin]`` and cast to the storage dtype.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch
from torch import nn

from ... import _native as N


def fold_weight_norm(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """w = g * v / ||v||, norm over all dims but 0 (old-style ``weight_norm``)."""
    n = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return g * v / n


class _FoldedWeightNorm:
    """``vq_model.dec`` as the reference's loader touches it (Loader.py:73, 95: ``vq_model.dec.remove_weight_norm()``): the
    weight-norm pairs of a checkpoint are folded by ``load_state_dict`` here, so the call has nothing left to do."""

    def remove_weight_norm(self):
        return None


class FlowDecoder(nn.Module):
    """``flow`` + ``dec`` of SynthesizerTrn on the native path.  Parameters are kept as a plain
    name -> tensor table in the reference's key space (``flow.flows.0.pre.weight`` ...)."""

    def __init__(self, inter_channels=192, hidden_channels=192, resblock="1", resblock_kernel_sizes=(3, 7, 11),
                 resblock_dilation_sizes=((1, 3, 5), (1, 3, 5), (1, 3, 5)), upsample_rates=(10, 8, 2, 2, 2),
                 upsample_initial_channel=512, upsample_kernel_sizes=(16, 16, 8, 2, 2), gin_channels=512,
                 version="v2", **_ignored):
        super().__init__()
        self.dec = _FoldedWeightNorm()
        if str(resblock) != "1":
            raise ValueError("only ResBlock1 generators exist in the reference (models.py:85)")
        self.inter_channels = inter_channels
        self.hidden_channels = hidden_channels
        self.resblock_kernel_sizes = list(resblock_kernel_sizes)
        self.resblock_dilation_sizes = [list(d) for d in resblock_dilation_sizes]
        self.upsample_rates = list(upsample_rates)
        self.upsample_initial_channel = upsample_initial_channel
        self.upsample_kernel_sizes = list(upsample_kernel_sizes)
        self.gin_channels = gin_channels
        self.version = version
        self.samples_per_frame = math.prod(self.upsample_rates)       # models.py:279
        self.n_flows, self.wn_layers, self.wn_kernel = 4, 4, 5         # models.py:303
        self._raw: Dict[str, torch.Tensor] = {}
        self._dev: Dict[str, torch.Tensor] = {}
        self._ctx = None

    # ---- checkpoint ingestion --------------------------------------------------------------
    def load_state_dict(self, state_dict, strict: bool = False):
        """Keeps every ``flow.*`` / ``dec.*`` tensor; other keys (enc_p, quantizer, training-only
        modules) are ignored like ``strict=False`` does in the reference (Loader.py:94)."""
        self._raw = {k: v.detach().float() for k, v in state_dict.items()        # stay where they are (CPU file / GPU broadcast)
                     if k.startswith("flow.") or k.startswith("dec.")}
        return self

    def state_dict(self, *a, **k):
        return dict(self._raw)

    def _weight(self, prefix: str) -> torch.Tensor:
        if prefix + ".weight" in self._raw:
            return self._raw[prefix + ".weight"]
        return fold_weight_norm(self._raw[prefix + ".weight_g"], self._raw[prefix + ".weight_v"])

    def _conv_names(self):
        names = []
        for f in range(self.n_flows):
            p = f"flow.flows.{2 * f}."
            names.append((p + "pre", False))
            names.append((p + "post", False))
            names.append((p + "enc.cond_layer", False))
            for l in range(self.wn_layers):
                names.append((p + f"enc.in_layers.{l}", False))
                names.append((p + f"enc.res_skip_layers.{l}", False))
        names.append(("dec.conv_pre", False))
        names.append(("dec.cond", False))
        nk = len(self.resblock_kernel_sizes)
        for i in range(len(self.upsample_rates)):
            names.append((f"dec.ups.{i}", True))
            for j in range(nk):
                for c in range(3):
                    names.append((f"dec.resblocks.{i * nk + j}.convs1.{c}", False))
                    names.append((f"dec.resblocks.{i * nk + j}.convs2.{c}", False))
        names.append(("dec.conv_post", False))
        return names

    # ---- runtime ---------------------------------------------------------------------------------
    @torch.inference_mode()
    def initialize_runtime(self, dtype, device, sovits_cache=None):
        """Counterpart of models.py:322-369.  ``sovits_cache`` (graph bucket lengths there) is
        accepted for signature compatibility; the native path takes any T and true batches."""
        device = torch.device(device)
        if device.type != "cuda":
            raise N.NativeError("the B200 backend needs a CUDA device; there is no CPU path here")
        self._device, self._dtype = device, dtype
        d = N.VocDims()
        d.inter_channels, d.hidden_channels, d.gin_channels = self.inter_channels, self.hidden_channels, self.gin_channels
        d.n_flows, d.wn_layers, d.wn_kernel = self.n_flows, self.wn_layers, self.wn_kernel
        d.upsample_initial_channel, d.n_ups = self.upsample_initial_channel, len(self.upsample_rates)
        for i, (u, k) in enumerate(zip(self.upsample_rates, self.upsample_kernel_sizes)):
            d.upsample_rates[i], d.upsample_kernel_sizes[i] = u, k
        d.n_resblock_kernels = len(self.resblock_kernel_sizes)
        for j, k in enumerate(self.resblock_kernel_sizes):
            d.resblock_kernel_sizes[j] = k
            for c in range(3):
                d.resblock_dilations[j][c] = self.resblock_dilation_sizes[j][c]
        d.dtype = N.dtype_code(dtype)
        ctx = C.c_void_p()
        with torch.cuda.device(device):
            N.check(N.lib().gsv_voc_create(C.byref(d), C.byref(ctx)))
        self._ctx = ctx
        for name, transposed in self._conv_names():
            w = self._weight(name)
            # Conv1d [Cout][Cin][k] -> [k][Cout][Cin]; ConvTranspose1d [Cin][Cout][k] -> [k][Cout][Cin]
            w = w.permute(2, 1, 0) if transposed else w.permute(2, 0, 1)
            wd = w.to(device=device, dtype=dtype).contiguous()
            b = self._raw.get(name + ".bias")
            bd = b.to(device=device, dtype=dtype).contiguous() if b is not None else None
            self._dev[name] = wd
            if bd is not None:
                self._dev[name + "#b"] = bd
            N.check(N.lib().gsv_voc_set_weight(ctx, name.encode(), wd.data_ptr(), bd.data_ptr() if bd is not None else None))

    def __del__(self):
        try:
            if self._ctx is not None:
                N.lib().gsv_voc_destroy(self._ctx)
                self._ctx = None
        except Exception:
            pass

    @torch.inference_mode()
    def flow_dec(self, z_p: torch.Tensor, y_mask: torch.Tensor, ge: torch.Tensor,
                 return_z: bool = False):
        """models.py:380-383: ``dec(flow(z_p, y_mask, ge, reverse) * y_mask, g=ge)``.
        z_p [B,192,T], y_mask [B,1,T], ge [B,gin,1] or [B,gin,T] -> [B,1,T*samples_per_frame]."""
        B, Cc, T = z_p.shape
        if Cc != self.inter_channels or ge.shape[1] != self.gin_channels or ge.shape[-1] not in (1, T):
            raise ValueError(f"flow_dec: bad shapes z_p {tuple(z_p.shape)} ge {tuple(ge.shape)}")
        if ge.shape[0] != B:
            ge = ge.expand(B, -1, -1)
        z_p = z_p.to(device=self._device, dtype=self._dtype).contiguous()
        mask = y_mask.to(device=self._device, dtype=self._dtype).contiguous()
        ge = ge.to(device=self._device, dtype=self._dtype).contiguous()
        out = torch.empty(B, 1, T * self.samples_per_frame, device=self._device, dtype=self._dtype)
        z = torch.empty_like(z_p) if return_z else None
        N.check(N.lib().gsv_voc_set_debug_z(self._ctx, z.data_ptr() if return_z else None))
        st = C.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)
        N.check(N.lib().gsv_voc_flow_dec(self._ctx, z_p.data_ptr(), mask.data_ptr(), ge.data_ptr(), B, T,
                                         ge.shape[-1], out.data_ptr(), st))
        return (out, z) if return_z else out

    def launch_count(self) -> int:
        return int(N.lib().gsv_voc_launch_count(self._ctx))


class _PriorEncoderState:
    """What callers touch on ``vq_model.enc_p``: the cross-chunk state of streaming decode (reference models.py:194,
    213-215; ``TTS.infer_stream`` resets it with ``enc_p.y_overlap = None``, TTS.py:498).  The tensor itself lives in the
    native context; here it can only be forgotten."""

    def __init__(self, owner):
        self._owner = owner
        self.latent_channels = owner.inter_channels

    @property
    def y_overlap(self):
        return None

    @y_overlap.setter
    def y_overlap(self, value):
        if value is not None:
            raise ValueError("enc_p.y_overlap can only be reset to None on the B200 backend")
        if self._owner._enc_ctx is not None:
            N.check(N.lib().gsv_encp_reset_stream(self._owner._enc_ctx))

    def rollback(self):
        """Undo the last ``decode(stream_mode=True)``'s update of the cross-chunk state (one level): for a chunk that was
        decoded ahead of time and then dropped (``TTS.infer_phones_stream``)."""
        N.check(N.lib().gsv_encp_stream_rollback(self._owner._enc_ctx))


class SynthesizerTrn(FlowDecoder):
    """``SynthesizerTrn`` as ``Loader.get_sovits_weights`` builds it (reference SoVITS/models.py:235-429), inference half:
    ``decode`` (codes -> waveform: quantizer lookup, prior encoder ``enc_p``, prior sample, reverse flow, HiFi-GAN),
    ``flow_dec``, ``initialize_runtime``, ``samples_per_frame``, ``enc_p.y_overlap``.  Everything between the semantic
    tokens and the waveform runs in ``libgsv_b200.so`` (csrc/encp.cu, csrc/vocoder.cu).  ``get_ge`` / ``extract_latent``
    (reference-audio featurisers, run once per cached speaker / prompt, TTS.py:1374, 1567) are not on the hot path: they are
    kept for the drop-in as plain tensor expressions over the checkpoint's own ``ref_enc`` / ``sv_emb`` / ``ssl_proj`` /
    codebook tensors (no kernels of this library, nothing to measure)."""

    def __init__(self, spec_channels=1025, segment_size=32, inter_channels=192, hidden_channels=192, filter_channels=768, n_heads=2,
                 n_layers=6, kernel_size=3, p_dropout=0.0, n_speakers=0, gin_channels=512, semantic_frame_rate="25hz", version="v2",
                 **kw):
        super().__init__(inter_channels=inter_channels, hidden_channels=hidden_channels, gin_channels=gin_channels, version=version, **kw)
        self.filter_channels, self.n_heads, self.n_layers, self.kernel_size = filter_channels, n_heads, n_layers, kernel_size
        self.is_v2pro = version in ("v2Pro", "v2ProPlus")
        self._enc_raw: Dict[str, torch.Tensor] = {}
        self._enc_dev: Dict[str, torch.Tensor] = {}
        self._enc_ctx = None
        self.enc_p = _PriorEncoderState(self)
        self.debug_seed: Optional[int] = None
        self._noise = None          # test hook: [inter, T'] fp32 tensor standing for randn_like(m_p)

    def load_state_dict(self, state_dict, strict: bool = False):
        super().load_state_dict(state_dict, strict)
        keep = ("enc_p.", "ge_to512.", "quantizer.vq.layers.0._codebook.embed")
        self._enc_raw = {k: v.detach().float() for k, v in state_dict.items() if k.startswith(keep)}
        aux = ("ref_enc.", "sv_emb.", "prelu.", "ssl_proj.")
        self._aux_raw = {k: v.detach().float() for k, v in state_dict.items() if k.startswith(aux)}
        return self

    def state_dict(self, *a, **k):
        sd = super().state_dict()
        sd.update(self._enc_raw)
        sd.update(getattr(self, "_aux_raw", {}))
        return sd

    # ---- once per cached speaker / prompt (reference models.py:371-378, 431-434): tensor expressions, not kernels ------------
    def _aux(self, name: str, like: torch.Tensor) -> torch.Tensor:
        raw = getattr(self, "_aux_raw", {})
        if name not in raw:
            raise KeyError(f"this SoVITS checkpoint carries no '{name}': get_ge / extract_latent need ref_enc.*, sv_emb.*, ssl_proj.*")
        cache = self.__dict__.setdefault("_aux_cache", {})
        key = (name, like.device, like.dtype)
        if key not in cache:
            cache[key] = raw[name].to(device=like.device, dtype=like.dtype)
        return cache[key]

    @torch.inference_mode()
    def get_ge(self, refer: torch.Tensor, sv_emb: Optional[torch.Tensor] = None) -> torch.Tensor:
        """models.py:371-378: refer [1, spec, T] (linear spectrogram of the speaker reference) [, sv_emb [1, 20480]] ->
        ge [1, gin, 1].  ``ref_enc`` is MelStyleEncoder(704) (module/modules.py:367-446): two Linear + Mish, two Conv1dGLU
        (k = 5), one 2-head self-attention with a residual (scores / sqrt(d_model)), Linear, mean over time; for v2Pro the
        speaker-verification embedding is added through ``sv_emb`` and a PReLU."""
        F = torch.nn.functional
        w = lambda n: self._aux(n, refer)
        x = refer[:, :704].transpose(1, 2)                                     # [1, T, 704]
        mish = lambda t: t * torch.tanh(F.softplus(t))
        x = mish(F.linear(x, w("ref_enc.spectral.0.fc.weight"), w("ref_enc.spectral.0.fc.bias")))
        x = mish(F.linear(x, w("ref_enc.spectral.3.fc.weight"), w("ref_enc.spectral.3.fc.bias")))
        x = x.transpose(1, 2)                                                  # [1, 128, T]
        for i in range(2):
            cw, cb = w(f"ref_enc.temporal.{i}.conv1.conv.weight"), w(f"ref_enc.temporal.{i}.conv1.conv.bias")
            h = F.conv1d(x, cw, cb, padding=(cw.shape[-1] - 1) // 2)
            a, g = torch.split(h, x.shape[1], dim=1)
            x = x + a * torch.sigmoid(g)
        x = x.transpose(1, 2)                                                  # [1, T, 128]
        n_head = 2
        d_model = x.shape[-1]
        dk = d_model // n_head
        b, t, _ = x.shape
        q = F.linear(x, w("ref_enc.slf_attn.w_qs.weight"), w("ref_enc.slf_attn.w_qs.bias")).view(b, t, n_head, dk).permute(2, 0, 1, 3)
        k = F.linear(x, w("ref_enc.slf_attn.w_ks.weight"), w("ref_enc.slf_attn.w_ks.bias")).view(b, t, n_head, dk).permute(2, 0, 1, 3)
        v = F.linear(x, w("ref_enc.slf_attn.w_vs.weight"), w("ref_enc.slf_attn.w_vs.bias")).view(b, t, n_head, dk).permute(2, 0, 1, 3)
        attn = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) / (d_model ** 0.5), dim=-1)
        o = torch.matmul(attn, v).permute(1, 2, 0, 3).reshape(b, t, d_model)
        x = F.linear(o, w("ref_enc.slf_attn.fc.weight"), w("ref_enc.slf_attn.fc.bias")) + x
        x = F.linear(x, w("ref_enc.fc.fc.weight"), w("ref_enc.fc.fc.bias"))
        ge = x.float().div(t).sum(dim=1).to(x.dtype).unsqueeze(-1)             # temporal_avg_pool with an all-ones mask
        if self.is_v2pro and sv_emb is not None:
            ge = ge + F.linear(sv_emb.to(ge.dtype), w("sv_emb.weight"), w("sv_emb.bias")).unsqueeze(-1)
            ge = F.prelu(ge, w("prelu.weight"))
        return ge

    @torch.inference_mode()
    def extract_latent(self, x: torch.Tensor) -> torch.Tensor:
        """models.py:431-434: HuBERT features [1, 768, T] -> semantic codes [1, 1, T // 2] (the prompt tokens): ``ssl_proj``
        (Conv1d k = 2, stride 2) and the nearest codebook row (module/core_vq.py:124-128: arg-max of the negated squared
        distance, expanded as the reference expands it)."""
        F = torch.nn.functional
        ssl = F.conv1d(x, self._aux("ssl_proj.weight", x), self._aux("ssl_proj.bias", x), stride=2)
        embed = self._enc_raw["quantizer.vq.layers.0._codebook.embed"].to(device=x.device, dtype=x.dtype).t()    # [768, 1024]
        flat = ssl.transpose(1, 2).reshape(-1, ssl.shape[1])
        dist = -(flat.pow(2).sum(1, keepdim=True) - 2 * flat @ embed + embed.pow(2).sum(0, keepdim=True))
        codes = dist.max(dim=-1).indices.view(ssl.shape[0], -1)               # [B, T']
        return codes.unsqueeze(0).transpose(0, 1)                              # [n_q = 1, B, T'] -> transpose(0, 1), as the reference returns it

    @torch.inference_mode()
    def initialize_runtime(self, dtype, device, sovits_cache=None):
        super().initialize_runtime(dtype, device, sovits_cache)
        raw = self._enc_raw
        if "enc_p.proj.weight" not in raw:
            return                      # a flow + HiFi-GAN only checkpoint: decode() is unavailable, flow_dec works
        d = N.EncpDims(hidden_channels=self.hidden_channels, filter_channels=self.filter_channels, inter_channels=self.inter_channels,
                       n_heads=self.n_heads, n_layers=self.n_layers, kernel_size=self.kernel_size,
                       ssl_dim=raw["quantizer.vq.layers.0._codebook.embed"].shape[1],
                       n_codes=raw["quantizer.vq.layers.0._codebook.embed"].shape[0],
                       n_symbols=raw["enc_p.text_embedding.weight"].shape[0], mrte_channels=raw["enc_p.mrte.c_pre.weight"].shape[0],
                       mrte_heads=4, gin_channels=self.gin_channels, dtype=N.dtype_code(dtype))
        ctx = C.c_void_p()
        with torch.cuda.device(self._device):
            N.check(N.lib().gsv_encp_create(C.byref(d), C.byref(ctx)))
        self._enc_ctx = ctx

        self._enc_names = []
        self._enc_dims = d
        self._enc_lanes = {}

        def put(name, w, b=None):
            wd = w.to(device=self._device, dtype=dtype).contiguous()
            bd = b.to(device=self._device, dtype=dtype).contiguous() if b is not None else None
            self._enc_names.append(name)
            self._enc_dev[name] = wd
            if bd is not None:
                self._enc_dev[name + "#b"] = bd
            N.check(N.lib().gsv_encp_set_weight(ctx, name.encode(), wd.data_ptr(), bd.data_ptr() if bd is not None else None))

        def lin(name):                  # Conv1d k=1 [out][in][1] -> [out][in]
            put(name, raw[name + ".weight"][:, :, 0], raw[name + ".bias"])

        put("quantizer.codebook", raw["quantizer.vq.layers.0._codebook.embed"])
        put("enc_p.text_embedding", raw["enc_p.text_embedding.weight"])
        lin("enc_p.ssl_proj")
        lin("enc_p.proj")
        for nm in ("c_pre", "text_pre", "c_post"):
            lin("enc_p.mrte." + nm)
        ca = "enc_p.mrte.cross_attention."
        lin(ca + "conv_q")
        lin(ca + "conv_o")
        put(ca + "kv", torch.cat([raw[ca + "conv_k.weight"][:, :, 0], raw[ca + "conv_v.weight"][:, :, 0]], 0),
            torch.cat([raw[ca + "conv_k.bias"], raw[ca + "conv_v.bias"]], 0))
        for enc, n in (("encoder_ssl", self.n_layers // 2), ("encoder_text", self.n_layers), ("encoder2", self.n_layers // 2)):
            for i in range(n):
                a = f"enc_p.{enc}.attn_layers.{i}."
                put(a + "qkv", torch.cat([raw[a + f"conv_{c}.weight"][:, :, 0] for c in "qkv"], 0),
                    torch.cat([raw[a + f"conv_{c}.bias"] for c in "qkv"], 0))
                lin(a + "conv_o")
                put(a + "emb_rel_k", raw[a + "emb_rel_k"][0])
                put(a + "emb_rel_v", raw[a + "emb_rel_v"][0])
                for j in (1, 2):
                    put(f"enc_p.{enc}.norm_layers_{j}.{i}", raw[f"enc_p.{enc}.norm_layers_{j}.{i}.gamma"],
                        raw[f"enc_p.{enc}.norm_layers_{j}.{i}.beta"])
                    f = f"enc_p.{enc}.ffn_layers.{i}.conv_{j}"
                    put(f, raw[f + ".weight"].permute(2, 0, 1), raw[f + ".bias"])      # [out][in][k] -> [k][out][in]
        if self.is_v2pro and "ge_to512.weight" in raw:
            put("ge_to512", raw["ge_to512.weight"], raw["ge_to512.bias"])

    def _lane_ctx(self, lane: int):
        """Prior-encoder context ``lane`` (0 = the one streaming decode keeps its cross-chunk state in).  Further lanes share
        the device weights and own only their scratch: independent utterances can go through the prior encoder on several
        streams at once (``TTS._sovits_stage_device``) -- a single call is a chain of ~100 small dependent launches."""
        if lane == 0:
            return self._enc_ctx
        if lane not in self._enc_lanes:
            ctx = C.c_void_p()
            with torch.cuda.device(self._device):
                N.check(N.lib().gsv_encp_create(C.byref(self._enc_dims), C.byref(ctx)))
            for name in self._enc_names:
                b = self._enc_dev.get(name + "#b")
                N.check(N.lib().gsv_encp_set_weight(ctx, name.encode(), self._enc_dev[name].data_ptr(), b.data_ptr() if b is not None else None))
            self._enc_lanes[lane] = ctx
        return self._enc_lanes[lane]

    def enc_launch_count(self) -> int:
        ctxs = ([self._enc_ctx] if self._enc_ctx is not None else []) + list(getattr(self, "_enc_lanes", {}).values())
        return sum(int(N.lib().gsv_encp_launch_count(c)) for c in ctxs)

    def __del__(self):
        try:
            for ctx in list(getattr(self, "_enc_lanes", {}).values()):
                N.lib().gsv_encp_destroy(ctx)
            self._enc_lanes = {}
            if self._enc_ctx is not None:
                N.lib().gsv_encp_destroy(self._enc_ctx)
                self._enc_ctx = None
        except Exception:
            pass
        super().__del__()

    @torch.inference_mode()
    def prior(self, codes, text, ge, noise_scale=0.5, speed=1, stream_mode=False, valid_start_idx=None, overlap_len=None,
              slice_indices=None, return_stats=False, lane: int = 0, text_unchanged: bool = False):
        """The front of ``decode`` (models.py:387-404): -> (z_p [1,inter,T'], y_mask [1,1,T'], ge for flow_dec, attn [4,T,Nt]
        [, m_p, logs_p]).  ``lane`` picks the native context (``_lane_ctx``); streaming calls use lane 0.
        ``text_unchanged``: the caller vouches that ``text`` holds what the previous call on this lane was given (the chunks
        of one streaming utterance): the text branch of the encoder is not recomputed (``gsv_encp_reuse_text``)."""
        if self._enc_ctx is None:
            raise N.NativeError("this SoVITS checkpoint carries no enc_p weights: decode() needs them")
        if stream_mode and lane != 0:
            raise ValueError("streaming decode keeps its cross-chunk state in lane 0")
        enc_ctx = self._lane_ctx(lane)
        dev, dt = self._device, self._dtype
        codes = codes.to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
        text = text.to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
        n, nt = codes.numel(), text.numel()
        ge = ge.to(device=dev, dtype=dt)
        if ge.shape[-1] != 1:                                        # models.py:389
            ge = torch.nn.functional.interpolate(ge.float(), size=ge.shape[-1] * 2, mode="nearest").to(dt)
        ge_c = ge[0].contiguous()
        tg = ge_c.shape[-1]
        lib = N.lib()
        speed = float(speed)
        vs = int(valid_start_idx) if stream_mode else 0
        ov = int(overlap_len) if stream_mode else 0
        tp = lib.gsv_encp_output_frames(enc_ctx, n, speed, 1 if stream_mode else 0, vs)
        z_p = torch.empty(1, self.inter_channels, tp, device=dev, dtype=dt)
        attn = torch.empty(4, 2 * n, nt, device=dev, dtype=torch.float32)
        m_p = torch.empty(self.inter_channels, tp, device=dev, dtype=torch.float32) if return_stats else None
        logs_p = torch.empty_like(m_p) if return_stats else None
        sl = None
        if slice_indices is not None:                               # [1, 2] or one (start, end) row per frame
            sl = slice_indices.to(device=dev, dtype=torch.int32).reshape(-1, 2).contiguous()
            if sl.shape[0] not in (1, 2 * n):
                raise ValueError(f"slice_indices has {sl.shape[0]} rows; expected 1 or {2 * n}")
        noise = self._noise.to(device=dev, dtype=torch.float32).contiguous() if self._noise is not None else None
        if noise is not None and tuple(noise.shape[-2:]) != (self.inter_channels, tp):
            raise ValueError(f"injected noise has shape {tuple(noise.shape)}, expected [{self.inter_channels}, {tp}]")
        seed = self.debug_seed if self.debug_seed is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
        frames = C.c_int(0)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if text_unchanged:
            N.check(lib.gsv_encp_reuse_text(enc_ctx, 1))
        N.check(lib.gsv_encp_forward(enc_ctx, codes.data_ptr(), n, text.data_ptr(), nt, ge_c.data_ptr(), tg, speed,
                                     1 if stream_mode else 0, vs, ov, sl.data_ptr() if sl is not None else None,
                                     sl.shape[0] if sl is not None else 0, noise.data_ptr() if noise is not None else None,
                                     float(noise_scale), seed, z_p.data_ptr(), m_p.data_ptr() if return_stats else None,
                                     logs_p.data_ptr() if return_stats else None, attn.data_ptr(), C.byref(frames), st))
        assert frames.value == tp
        y_mask = torch.ones(1, 1, tp, device=dev, dtype=dt)
        if speed != 1 and ge.shape[-1] != 1:                         # models.py:402
            ge = torch.nn.functional.interpolate(ge.float(), size=tp, mode="nearest").to(dt)
        out = (z_p, y_mask, ge, attn)
        return out + (m_p, logs_p) if return_stats else out

    @torch.inference_mode()
    def decode(self, codes, text, ge, noise_scale=0.5, speed=1, cuda_graph=True, stream_mode=False, valid_start_idx=None,
               overlap_len=None, slice_indices=None, text_unchanged: bool = False):
        """models.py:385-429: codes [1,1,N] int64, text [1,Nt] int64, ge [1,gin,1|N] -> (audio [1,1,640 T'], attn [4,T,Nt]).
        ``cuda_graph`` is accepted for signature compatibility (the native path takes any length)."""
        z_p, y_mask, ge2, attn = self.prior(codes, text, ge, noise_scale, speed, stream_mode, valid_start_idx, overlap_len, slice_indices,
                                            text_unchanged=text_unchanged)
        return self.flow_dec(z_p, y_mask, ge2), attn
