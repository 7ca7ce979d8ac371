"""Synthetic checkpoints in the reference's state-dict layout.

There are no pretrained weights on the build or GPU boxes and no network, so every
test and benchmark runs on seeded random weights of the reference's architecture.
Key names and shapes follow what ``Loader.get_gpt_weights`` / ``get_sovits_weights``
hand to ``load_state_dict`` (reference gsv_tts/Loader.py:130-162, 80-95; SURVEY.md
A.1, A.6).  Values are chosen so the network is *non-degenerate*: LayerNorm gamma/beta,
the positional ``alpha`` scalars, weight-norm ``g`` and the zero-initialised coupling
``post`` convolution (reference modules.py:479-480) are all randomised, otherwise a
parity test against them would be vacuous.

All draws come from a CPU ``torch.Generator`` so the same seed gives the same tensors
on every machine with this torch build.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

GPT_CONFIG = {
    "model": {
        "hidden_dim": 512,
        "embedding_dim": 512,
        "head": 16,
        "n_layer": 24,
        "vocab_size": 1025,
        "phoneme_vocab_size": 732,
        "dropout": 0.0,
        "EOS": 1024,
    }
}

# a small config with the same head_dim (32) for fast CPU checks
GPT_CONFIG_TINY = {
    "model": {
        "hidden_dim": 256,
        "embedding_dim": 256,
        "head": 8,
        "n_layer": 3,
        "vocab_size": 1025,
        "phoneme_vocab_size": 732,
        "dropout": 0.0,
        "EOS": 1024,
    }
}

_SOVITS_COMMON = dict(
    inter_channels=192,
    hidden_channels=192,
    filter_channels=768,
    n_heads=2,
    n_layers=6,
    kernel_size=3,
    p_dropout=0.0,
    resblock="1",
    resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    upsample_rates=[10, 8, 2, 2, 2],
    upsample_kernel_sizes=[16, 16, 8, 2, 2],
)

SOVITS_MODEL = {
    "v2": dict(_SOVITS_COMMON, upsample_initial_channel=512, gin_channels=512, version="v2"),
    "v2Pro": dict(_SOVITS_COMMON, upsample_initial_channel=512, gin_channels=1024, version="v2Pro"),
    "v2ProPlus": dict(_SOVITS_COMMON, upsample_initial_channel=768, gin_channels=1024, version="v2ProPlus"),
    # reduced widths, same topology: CPU-fast parity case
    "tiny": dict(_SOVITS_COMMON, upsample_initial_channel=256, gin_channels=64, version="v2"),
}


def _randn(gen, *shape, std=1.0):
    return torch.randn(*shape, generator=gen, dtype=torch.float32) * std


def gpt_state_dict(config=GPT_CONFIG, seed: int = 0, eos_boost: float = 0.0) -> Dict[str, torch.Tensor]:
    """fp32 state dict with the key set of the reference ``Text2SemanticDecoder``.

    ``eos_boost`` adds a constant to the EOS logit (through the last LayerNorm's beta and
    the EOS row of ``ar_predict_layer``, which has no bias) so that sampled sequences
    terminate in tests of the EOS handling.
    """
    m = config["model"]
    d, L, V, P = m["hidden_dim"], m["n_layer"], m["vocab_size"], m["phoneme_vocab_size"]
    F = 4 * d
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    sd["bert_proj.weight"] = _randn(g, d, 1024, std=0.5 / math.sqrt(1024))
    sd["bert_proj.bias"] = _randn(g, d, std=0.02)
    sd["ar_text_embedding.word_embeddings.weight"] = _randn(g, P, d, std=0.7)
    sd["ar_text_position.alpha"] = torch.tensor([1.15])
    sd["ar_audio_embedding.word_embeddings.weight"] = _randn(g, V, d, std=0.7)
    sd["ar_audio_position.alpha"] = torch.tensor([0.85])
    sd["ar_predict_layer.weight"] = _randn(g, V, d, std=2.0 / math.sqrt(d))
    for i in range(L):
        p = f"t2s_transformer.blocks.{i}."
        sd[p + "qkv.weight"] = _randn(g, 3 * d, d, std=1.0 / math.sqrt(d))
        sd[p + "qkv.bias"] = _randn(g, 3 * d, std=0.05)
        sd[p + "out_proj.weight"] = _randn(g, d, d, std=1.0 / math.sqrt(d))
        sd[p + "out_proj.bias"] = _randn(g, d, std=0.05)
        sd[p + "mlp.0.weight"] = _randn(g, F, d, std=1.0 / math.sqrt(d))
        sd[p + "mlp.0.bias"] = _randn(g, F, std=0.05)
        sd[p + "mlp.2.weight"] = _randn(g, d, F, std=1.0 / math.sqrt(F))
        sd[p + "mlp.2.bias"] = _randn(g, d, std=0.05)
        sd[p + "norm1.weight"] = 1.0 + _randn(g, d, std=0.1)
        sd[p + "norm1.bias"] = _randn(g, d, std=0.05)
        sd[p + "norm2.weight"] = 1.0 + _randn(g, d, std=0.1)
        sd[p + "norm2.bias"] = _randn(g, d, std=0.05)
    if eos_boost != 0.0:
        u = _randn(g, d)
        u = u / u.norm()
        last = f"t2s_transformer.blocks.{L - 1}.norm2."
        sd[last + "bias"] = sd[last + "bias"] + u
        sd["ar_predict_layer.weight"][m["EOS"]] = eos_boost * u
    return sd


def _wn_pair(gen, cout, cin, k, gain):
    """(weight_g, weight_v) for old-style weight norm over dims (1,2)."""
    v = _randn(gen, cout, cin, k)
    target = gain * torch.ones(cout) * (1.0 + 0.1 * torch.randn(cout, generator=gen))
    return target.abs().view(cout, 1, 1), v


def sovits_flow_dec_state_dict(model: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    """fp32 state dict for the ``flow.*`` and ``dec.*`` sub-modules of ``SynthesizerTrn``.

    ``flow`` keeps ``weight_g``/``weight_v`` (the reference never strips flow's weight norm,
    Loader.py:73,95); ``dec`` uses plain ``weight`` as after ``dec.remove_weight_norm()``.
    """
    g = torch.Generator().manual_seed(seed + 1000)
    C = model["inter_channels"]
    Hc = model["hidden_channels"]
    gin = model["gin_channels"]
    half = C // 2
    nl = 4
    sd: Dict[str, torch.Tensor] = {}
    for fi in (0, 2, 4, 6):
        p = f"flow.flows.{fi}."
        sd[p + "pre.weight"] = _randn(g, Hc, half, 1, std=1.0 / math.sqrt(half))
        sd[p + "pre.bias"] = _randn(g, Hc, std=0.05)
        for l in range(nl):
            wg, wv = _wn_pair(g, 2 * Hc, Hc, 5, gain=1.0)
            sd[p + f"enc.in_layers.{l}.weight_g"] = wg
            sd[p + f"enc.in_layers.{l}.weight_v"] = wv
            sd[p + f"enc.in_layers.{l}.bias"] = _randn(g, 2 * Hc, std=0.05)
            rs = 2 * Hc if l < nl - 1 else Hc
            wg, wv = _wn_pair(g, rs, Hc, 1, gain=0.7)
            sd[p + f"enc.res_skip_layers.{l}.weight_g"] = wg
            sd[p + f"enc.res_skip_layers.{l}.weight_v"] = wv
            sd[p + f"enc.res_skip_layers.{l}.bias"] = _randn(g, rs, std=0.05)
        wg, wv = _wn_pair(g, 2 * Hc * nl, gin, 1, gain=0.5)
        sd[p + "enc.cond_layer.weight_g"] = wg
        sd[p + "enc.cond_layer.weight_v"] = wv
        sd[p + "enc.cond_layer.bias"] = _randn(g, 2 * Hc * nl, std=0.05)
        sd[p + "post.weight"] = _randn(g, half, Hc, 1, std=0.3 / math.sqrt(Hc))
        sd[p + "post.bias"] = _randn(g, half, std=0.02)

    C0 = model["upsample_initial_channel"]
    sd["dec.conv_pre.weight"] = _randn(g, C0, C, 7, std=1.0 / math.sqrt(C * 7))
    sd["dec.conv_pre.bias"] = _randn(g, C0, std=0.05)
    sd["dec.cond.weight"] = _randn(g, C0, gin, 1, std=0.5 / math.sqrt(gin))
    sd["dec.cond.bias"] = _randn(g, C0, std=0.05)
    ch = C0
    for i, (u, k) in enumerate(zip(model["upsample_rates"], model["upsample_kernel_sizes"])):
        cin, cout = ch, ch // 2
        taps = max(1.0, k / u)
        sd[f"dec.ups.{i}.weight"] = _randn(g, cin, cout, k, std=1.2 / math.sqrt(cin * taps))
        sd[f"dec.ups.{i}.bias"] = _randn(g, cout, std=0.05)
        ch = cout
        for j, kk in enumerate(model["resblock_kernel_sizes"]):
            rp = f"dec.resblocks.{i * len(model['resblock_kernel_sizes']) + j}."
            for c in range(3):
                sd[rp + f"convs1.{c}.weight"] = _randn(g, ch, ch, kk, std=0.8 / math.sqrt(ch * kk))
                sd[rp + f"convs1.{c}.bias"] = _randn(g, ch, std=0.05)
                sd[rp + f"convs2.{c}.weight"] = _randn(g, ch, ch, kk, std=0.5 / math.sqrt(ch * kk))
                sd[rp + f"convs2.{c}.bias"] = _randn(g, ch, std=0.05)
    sd["dec.conv_post.weight"] = _randn(g, 1, ch, 7, std=0.6 / math.sqrt(ch * 7))
    return sd


def sovits_encp_state_dict(model: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    """fp32 tensors with the key set and shapes of the reference's ``enc_p.*``, ``quantizer`` codebook and ``ge_to512``
    (SoVITS/models.py:141-194, 305-317; probed from the reference module, checked by oracle/make_golden.py).  Scales are
    chosen so that activations stay O(1) through the 12 encoder layers; LayerNorm gains / biases are randomised."""
    g = torch.Generator().manual_seed(1000 + seed)
    C, Fc, H, L = model["hidden_channels"], model["filter_channels"], model["n_heads"], model["n_layers"]
    k = model["kernel_size"]
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, out_c, in_c, kk=1):
        sd[name + ".weight"] = _randn(g, out_c, in_c, kk, std=(in_c * kk) ** -0.5)
        sd[name + ".bias"] = _randn(g, out_c, std=0.05)

    def enc(pre, n):
        for i in range(n):
            a = f"{pre}.attn_layers.{i}"
            sd[a + ".emb_rel_k"] = _randn(g, 1, 9, C // H, std=(C // H) ** -0.5)
            sd[a + ".emb_rel_v"] = _randn(g, 1, 9, C // H, std=(C // H) ** -0.5)
            for nm in ("conv_q", "conv_k", "conv_v", "conv_o"):
                lin(f"{a}.{nm}", C, C)
            for j in (1, 2):
                sd[f"{pre}.norm_layers_{j}.{i}.gamma"] = 1.0 + _randn(g, C, std=0.1)
                sd[f"{pre}.norm_layers_{j}.{i}.beta"] = _randn(g, C, std=0.1)
            lin(f"{pre}.ffn_layers.{i}.conv_1", Fc, C, k)
            lin(f"{pre}.ffn_layers.{i}.conv_2", C, Fc, k)

    lin("enc_p.ssl_proj", C, 768)
    enc("enc_p.encoder_ssl", L // 2)
    enc("enc_p.encoder_text", L)
    sd["enc_p.text_embedding.weight"] = _randn(g, 732, C, std=1.0)
    for nm in ("conv_q", "conv_k", "conv_v", "conv_o"):
        lin(f"enc_p.mrte.cross_attention.{nm}", 512, 512)
    lin("enc_p.mrte.c_pre", 512, C)
    lin("enc_p.mrte.text_pre", 512, C)
    lin("enc_p.mrte.c_post", C, 512)
    enc("enc_p.encoder2", L // 2)
    lin("enc_p.proj", 2 * model["inter_channels"], C)
    sd["enc_p.proj.bias"][model["inter_channels"]:] -= 1.0          # logs around -1: exp(logs) stays tame
    sd["quantizer.vq.layers.0._codebook.embed"] = _randn(g, 1024, 768, std=1.0)
    if model.get("version") in ("v2Pro", "v2ProPlus"):
        lin("ge_to512", 512, model["gin_channels"])
        sd["ge_to512.weight"] = sd["ge_to512.weight"][:, :, 0].contiguous()
    return sd


def sovits_aux_state_dict(model: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    """fp32 tensors with the key set and shapes of the reference's ``ref_enc.*`` (MelStyleEncoder(704), modules.py:367-409),
    ``sv_emb`` / ``prelu`` (v2Pro, models.py:316-318) and the top-level ``ssl_proj`` (Conv1d(768, 768, 2, stride 2), :310):
    what ``SynthesizerTrn.get_ge`` / ``extract_latent`` read."""
    g = torch.Generator().manual_seed(2000 + seed)
    gin = model["gin_channels"]
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, out_c, in_c):
        sd[name + ".weight"] = _randn(g, out_c, in_c, std=in_c ** -0.5)
        sd[name + ".bias"] = _randn(g, out_c, std=0.05)

    lin("ref_enc.spectral.0.fc", 128, 704)
    lin("ref_enc.spectral.3.fc", 128, 128)
    for i in range(2):
        sd[f"ref_enc.temporal.{i}.conv1.conv.weight"] = _randn(g, 256, 128, 5, std=(128 * 5) ** -0.5)
        sd[f"ref_enc.temporal.{i}.conv1.conv.bias"] = _randn(g, 256, std=0.05)
    for nm in ("w_qs", "w_ks", "w_vs", "fc"):
        lin(f"ref_enc.slf_attn.{nm}", 128, 128)
    lin("ref_enc.fc.fc", gin, 128)
    sd["ssl_proj.weight"] = _randn(g, 768, 768, 2, std=(768 * 2) ** -0.5)
    sd["ssl_proj.bias"] = _randn(g, 768, std=0.05)
    if model.get("version") in ("v2Pro", "v2ProPlus"):
        lin("sv_emb", gin, 20480)
        sd["prelu.weight"] = 0.25 + _randn(g, gin, std=0.05)
    return sd
