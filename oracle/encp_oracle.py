"""ORACLE (test infrastructure, never imported by the product path): CPU restatement of the stage BETWEEN the two hot
paths -- ``SynthesizerTrn.decode`` up to ``z_p`` (reference SoVITS/models.py:385-404): codebook lookup, x2 nearest
interpolation, ``ge_to512``, ``TextEncoder.infer`` (models.py:196-224: ssl_proj, three relative-position Transformer
encoders, MRTE cross attention, streaming cross-fade, speed interpolation, proj) and the prior sample.  SURVEY.md 8f row
f-1: the next component to move to sm_100a kernels; this file and tests/golden/encp_*.npz are its parity anchor.

Written from the published VITS / GPT-SoVITS algorithm as the reference implements it; every function cites the reference
lines it follows.  Pinned by tests/test_encp_oracle_cpu.py to outputs of the reference's own modules
(oracle/make_golden.py encp -> tests/golden/encp_*.npz, re-checked live when /root/reference exists).
The relative-position terms are computed by explicit index arithmetic (gathering the embedding of offset j - i), not by
the reference's pad-and-reshape skewing, so the two derivations check each other.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
WINDOW = 4          # attentions.Encoder default window_size (attentions.py:19)


def channel_layer_norm(x: Tensor, gamma: Tensor, beta: Tensor) -> Tensor:
    """modules.py:24-27: LayerNorm over the channel axis of [B, C, T], eps 1e-5."""
    return F.layer_norm(x.transpose(1, 2), (x.shape[1],), gamma, beta, 1e-5).transpose(1, 2)


def conv1x1(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    return torch.einsum("oc,bct->bot", w[:, :, 0], x) + b.view(1, -1, 1)


def rel_embedding_table(emb: Tensor, length: int) -> Tensor:
    """attentions.py:177-191: [1, 2W+1, dk] -> [2L-1, dk], row r <-> offset r - (L-1); offsets beyond the window are zero."""
    out = emb.new_zeros(2 * length - 1, emb.shape[-1])
    for r in range(2 * length - 1):
        off = r - (length - 1)
        if -WINDOW <= off <= WINDOW:
            out[r] = emb[0, off + WINDOW]
    return out


def attention(x: Tensor, c: Tensor, sd: Dict[str, Tensor], pre: str, n_heads: int, mask: Optional[Tensor], relative: bool) -> Tensor:
    """attentions.py:119-161 (MultiHeadAttention.forward / attention).  x [B,C,Tt] queries, c [B,C,Ts] keys/values,
    mask broadcastable to [B,1,Tt,Ts] (0 = masked with -1e4)."""
    q = conv1x1(x, sd[pre + "conv_q.weight"], sd[pre + "conv_q.bias"])
    k = conv1x1(c, sd[pre + "conv_k.weight"], sd[pre + "conv_k.bias"])
    v = conv1x1(c, sd[pre + "conv_v.weight"], sd[pre + "conv_v.bias"])
    B, C, Tt = q.shape
    Ts = k.shape[2]
    dk = C // n_heads
    q = q.view(B, n_heads, dk, Tt).transpose(2, 3) / math.sqrt(dk)
    k = k.view(B, n_heads, dk, Ts).transpose(2, 3)
    v = v.view(B, n_heads, dk, Ts).transpose(2, 3)
    scores = q @ k.transpose(-2, -1)
    if relative:
        assert Tt == Ts
        idx = torch.arange(Ts).view(1, -1) - torch.arange(Tt).view(-1, 1) + (Tt - 1)       # [Tt, Ts] -> row of the table
        ek = rel_embedding_table(sd[pre + "emb_rel_k"], Tt)                                # [2L-1, dk]
        scores = scores + torch.einsum("bhid,ijd->bhij", q, ek[idx])                       # q_i . E_k[j - i]
    if mask is not None:
        scores = scores.masked_fill(mask == 0, -1e4)
    p = torch.softmax(scores, dim=-1)
    out = p @ v
    if relative:
        ev = rel_embedding_table(sd[pre + "emb_rel_v"], Tt)
        out = out + torch.einsum("bhij,ijd->bhid", p, ev[idx])                             # sum_j p_ij E_v[j - i]
    out = out.transpose(2, 3).reshape(B, C, Tt)
    return conv1x1(out, sd[pre + "conv_o.weight"], sd[pre + "conv_o.bias"])


def ffn(x: Tensor, mask: Tensor, sd: Dict[str, Tensor], pre: str) -> Tensor:
    """attentions.py:244-252, 264-271: conv(k) -> relu -> conv(k), 'same' padding ((k-1)//2 left, k//2 right), masked."""
    w1, w2 = sd[pre + "conv_1.weight"], sd[pre + "conv_2.weight"]
    k = w1.shape[-1]
    pad = ((k - 1) // 2, k // 2)
    h = torch.relu(F.conv1d(F.pad(x * mask, pad), w1, sd[pre + "conv_1.bias"]))
    return F.conv1d(F.pad(h * mask, pad), w2, sd[pre + "conv_2.bias"]) * mask


def encoder(x: Tensor, mask: Tensor, sd: Dict[str, Tensor], pre: str, n_layers: int, n_heads: int) -> Tensor:
    """attentions.py:59-80 (no conditioning branch: g is None at every call site in models.py)."""
    attn_mask = mask.unsqueeze(2) * mask.unsqueeze(-1)
    x = x * mask
    for i in range(n_layers):
        y = attention(x, x, sd, f"{pre}attn_layers.{i}.", n_heads, attn_mask, True)
        x = channel_layer_norm(x + y, sd[f"{pre}norm_layers_1.{i}.gamma"], sd[f"{pre}norm_layers_1.{i}.beta"])
        y = ffn(x, mask, sd, f"{pre}ffn_layers.{i}.")
        x = channel_layer_norm(x + y, sd[f"{pre}norm_layers_2.{i}.gamma"], sd[f"{pre}norm_layers_2.{i}.beta"])
    return x * mask


def mrte(ssl_enc: Tensor, ssl_mask: Tensor, text: Tensor, text_mask: Tensor, ge, sd: Dict[str, Tensor], pre: str,
         slice_indices: Optional[Tensor] = None) -> Tensor:
    """mrte_model.py:19-38: 4-head cross attention of the content frames over the text, residual + ge, 1x1 convs."""
    if slice_indices is None:
        attn_mask = text_mask.unsqueeze(2) * ssl_mask.unsqueeze(-1)
    else:
        rng = torch.arange(text.shape[-1]).unsqueeze(0)
        m = (rng >= slice_indices[:, 0].unsqueeze(-1)) & (rng < slice_indices[:, 1].unsqueeze(-1))
        m[:, -1] = True
        attn_mask = m.unsqueeze(0).unsqueeze(0)
    s = conv1x1(ssl_enc * ssl_mask, sd[pre + "c_pre.weight"], sd[pre + "c_pre.bias"])
    t = conv1x1(text * text_mask, sd[pre + "text_pre.weight"], sd[pre + "text_pre.bias"])
    x = attention(s * ssl_mask, t * text_mask, sd, pre + "cross_attention.", 4, attn_mask, False) + s + (0 if ge is None else ge)
    return conv1x1(x * ssl_mask, sd[pre + "c_post.weight"], sd[pre + "c_post.bias"])


class EncPOracle:
    """State: ``y_overlap`` between streaming chunks (models.py:213-215; reset by TTS.py:498)."""

    def __init__(self, state_dict: Dict[str, Tensor], model: dict):
        self.sd = {k: v.float() for k, v in state_dict.items()}
        self.n_heads = model["n_heads"]
        self.n_layers = model["n_layers"]
        self.out_channels = model["inter_channels"]
        self.is_v2pro = model.get("version") in ("v2Pro", "v2ProPlus")
        self.y_overlap = None

    def infer(self, y: Tensor, text: Tensor, ge: Tensor, speed: float = 1, stream_mode: bool = False,
              valid_start_idx: Optional[int] = None, overlap_len: Optional[int] = None, slice_indices=None):
        """TextEncoder.infer, models.py:196-224.  y [1,768,T] quantized frames, text [1,Nt] int64, ge [1,512,1|T]."""
        sd, H, L = self.sd, self.n_heads, self.n_layers
        y_mask = torch.ones(1, 1, y.shape[2])
        y = conv1x1(y * y_mask, sd["enc_p.ssl_proj.weight"], sd["enc_p.ssl_proj.bias"]) * y_mask
        y = encoder(y * y_mask, y_mask, sd, "enc_p.encoder_ssl.", L // 2, H)
        t_mask = torch.ones(1, 1, text.shape[1])
        t = sd["enc_p.text_embedding.weight"][text].transpose(1, 2)
        t = encoder(t * t_mask, t_mask, sd, "enc_p.encoder_text.", L, H)
        y = mrte(y, y_mask, t, t_mask, ge, sd, "enc_p.mrte.", slice_indices)
        y = encoder(y * y_mask, y_mask, sd, "enc_p.encoder2.", L // 2, H)
        if stream_mode:
            y = y[:, :, valid_start_idx:].clone()
            y_mask = y_mask[:, :, valid_start_idx:]
            alpha = torch.linspace(0, 1, overlap_len).view(1, 1, -1)
            if self.y_overlap is not None:
                y[:, :, :overlap_len] = self.y_overlap * (1 - alpha) + y[:, :, :overlap_len] * alpha
            self.y_overlap = y[:, :, -overlap_len:].clone()
        if speed != 1:
            y = F.interpolate(y, size=int(y.shape[-1] / speed) + 1, mode="linear")
            y_mask = F.interpolate(y_mask, size=y.shape[-1], mode="nearest")
        stats = conv1x1(y, sd["enc_p.proj.weight"], sd["enc_p.proj.bias"]) * y_mask
        m, logs = torch.split(stats, self.out_channels, dim=1)
        return m, logs, y_mask

    def decode_front(self, codes: Tensor, text: Tensor, ge: Tensor, noise: Optional[Tensor] = None, noise_scale: float = 0.5,
                     speed: float = 1, **stream):
        """SynthesizerTrn.decode, models.py:387-404: codes [1,1,N] int64 -> (z_p, y_mask, m_p, logs_p, ge for flow_dec).
        ``noise`` stands for ``torch.randn_like(m_p)`` (:404)."""
        q = self.sd["quantizer.vq.layers.0._codebook.embed"][codes[0]].transpose(1, 2)          # core_vq.py:133-135, 222-226
        q = F.interpolate(q, size=q.shape[-1] * 2, mode="nearest")
        if ge.shape[-1] != 1:
            ge = F.interpolate(ge, size=ge.shape[-1] * 2, mode="nearest")
        ge_in = ge
        if self.is_v2pro:
            ge_in = (ge.transpose(2, 1) @ self.sd["ge_to512.weight"].t() + self.sd["ge_to512.bias"]).transpose(2, 1)
        m_p, logs_p, y_mask = self.infer(q, text, ge_in, speed, **stream)
        if speed != 1 and ge.shape[-1] != 1:
            ge = F.interpolate(ge, size=m_p.shape[-1], mode="nearest")
        z_p = m_p if noise is None else m_p + noise * torch.exp(logs_p) * noise_scale
        return z_p, y_mask, m_p, logs_p, ge
