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


class FlowDecoder(nn.Module):
    """``flow`` + ``dec`` of SynthesizerTrn on the native path.  Parameters are kept as a plain
    name -> tensor table in the reference's key space (``flow.flows.0.pre.weight`` ...)."""

    def __init__(self, inter_channels=192, hidden_channels=192, resblock="1", resblock_kernel_sizes=(3, 7, 11),
                 resblock_dilation_sizes=((1, 3, 5), (1, 3, 5), (1, 3, 5)), upsample_rates=(10, 8, 2, 2, 2),
                 upsample_initial_channel=512, upsample_kernel_sizes=(16, 16, 8, 2, 2), gin_channels=512,
                 version="v2", **_ignored):
        super().__init__()
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
