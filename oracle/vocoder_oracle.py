"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's
SoVITS reverse flow + HiFi-GAN generator (``SynthesizerTrn.flow_dec``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg may import this module.

Parity pin: as for ``gpt_oracle`` -- pinned against the reference's own modules run in the
build container (``oracle/make_golden.py`` -> ``tests/golden/vocoder_*.npz``).

Reference lines restated (relative to ``/root/reference/gsv_tts/GPT_SoVITS/SoVITS/``):
models.py:58-65 (block, reversed order), :113-132 (generator), :380-383 (flow_dec);
module/modules.py:80-104 (WN), :190-203 (ResBlock1), :482-501 (coupling, reverse),
:504-511 (Flip); module/commons.py:14-21 (gate).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def fold_weight_norm(g: Tensor, v: Tensor) -> Tensor:
    """Old-style weight norm: w = g * v / ||v||, norm over every dim except 0."""
    n = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return g * v / n


class VocoderOracle:
    def __init__(self, state_dict: Dict[str, Tensor], model: dict, dtype=torch.float32):
        self.m = model
        self.dtype = dtype
        self.w = {k: v.to(dtype) for k, v in state_dict.items()}
        self.hidden = model["hidden_channels"]
        self.half = model["inter_channels"] // 2
        self.n_wn = 4
        self.ups = list(zip(model["upsample_rates"], model["upsample_kernel_sizes"]))
        self.rk = model["resblock_kernel_sizes"]
        self.rd = model["resblock_dilation_sizes"]
        self.samples_per_frame = 1
        for u, _ in self.ups:
            self.samples_per_frame *= u

    def _wn_weight(self, prefix: str) -> Tensor:
        if prefix + "weight" in self.w:
            return self.w[prefix + "weight"]
        return fold_weight_norm(self.w[prefix + "weight_g"], self.w[prefix + "weight_v"])

    # ---- WaveNet body of one coupling layer (modules.py:80-104) -------------------------
    def wn(self, p: str, h: Tensor, mask: Tensor, g: Tensor) -> Tensor:
        Hc = self.hidden
        cond = F.conv1d(g, self._wn_weight(p + "cond_layer."), self.w[p + "cond_layer.bias"])
        out = torch.zeros_like(h)
        for l in range(self.n_wn):
            a = F.conv1d(h, self._wn_weight(p + f"in_layers.{l}."), self.w[p + f"in_layers.{l}.bias"], padding=2)
            a = a + cond[:, 2 * Hc * l: 2 * Hc * (l + 1)]
            u = torch.tanh(a[:, :Hc]) * torch.sigmoid(a[:, Hc:])
            r = F.conv1d(u, self._wn_weight(p + f"res_skip_layers.{l}."), self.w[p + f"res_skip_layers.{l}.bias"])
            if l < self.n_wn - 1:
                h = (h + r[:, :Hc]) * mask
                out = out + r[:, Hc:]
            else:
                out = out + r
        return out * mask

    # ---- reverse flow (models.py:58-65; modules.py:482-511) ---------------------------------
    def flow_reverse(self, z: Tensor, mask: Tensor, g: Tensor) -> Tensor:
        for fi in (6, 4, 2, 0):
            z = torch.flip(z, [1])
            p = f"flow.flows.{fi}."
            z0, z1 = z[:, : self.half], z[:, self.half:]
            h = F.conv1d(z0, self.w[p + "pre.weight"], self.w[p + "pre.bias"]) * mask
            h = self.wn(p + "enc.", h, mask, g)
            m = F.conv1d(h, self.w[p + "post.weight"], self.w[p + "post.bias"]) * mask
            z = torch.cat([z0, (z1 - m) * mask], 1)
        return z

    # ---- HiFi-GAN generator (models.py:113-132; modules.py:190-203) -------------------------
    def resblock(self, idx: int, x: Tensor) -> Tensor:
        k = self.rk[idx % len(self.rk)]
        dil = self.rd[idx % len(self.rk)]
        p = f"dec.resblocks.{idx}."
        for c, d in enumerate(dil):
            t = F.leaky_relu(x, 0.1)
            t = F.conv1d(t, self._wn_weight(p + f"convs1.{c}."), self.w[p + f"convs1.{c}.bias"],
                         dilation=d, padding=(k * d - d) // 2)
            t = F.leaky_relu(t, 0.1)
            t = F.conv1d(t, self._wn_weight(p + f"convs2.{c}."), self.w[p + f"convs2.{c}.bias"],
                         padding=(k - 1) // 2)
            x = t + x
        return x

    def generator(self, z: Tensor, g: Tensor) -> Tensor:
        w = self.w
        x = F.conv1d(z, w["dec.conv_pre.weight"], w["dec.conv_pre.bias"], padding=3)
        x = x + F.conv1d(g, w["dec.cond.weight"], w["dec.cond.bias"])
        nk = len(self.rk)
        for i, (u, k) in enumerate(self.ups):
            x = F.leaky_relu(x, 0.1)
            x = F.conv_transpose1d(x, self._ups_weight(i), w[f"dec.ups.{i}.bias"], stride=u, padding=(k - u) // 2)
            acc = None
            for j in range(nk):
                r = self.resblock(i * nk + j, x)
                acc = r if acc is None else acc + r
            x = acc / nk
        x = F.leaky_relu(x)                      # default slope 0.01 (models.py:128)
        x = F.conv1d(x, w["dec.conv_post.weight"], None, padding=3)
        return torch.tanh(x)

    def _ups_weight(self, i: int) -> Tensor:
        p = f"dec.ups.{i}."
        if p + "weight" in self.w:
            return self.w[p + "weight"]
        return fold_weight_norm(self.w[p + "weight_g"], self.w[p + "weight_v"])   # dim 0 = Cin

    def flow_dec(self, z_p: Tensor, mask: Tensor, ge: Tensor) -> Tensor:
        """models.py:380-383.  z_p [B,192,T], mask [B,1,T], ge [B,gin,1 or T] -> [B,1,640*T]."""
        z_p, mask, ge = z_p.to(self.dtype), mask.to(self.dtype), ge.to(self.dtype)
        z = self.flow_reverse(z_p, mask, ge)
        return self.generator(z * mask, ge)
