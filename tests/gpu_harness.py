"""Shared helpers for the GPU parity tests and ``__graft_entry__.smoke()``: run the CUDA path
through the C ABI and the oracle on the same seeded inputs.  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from gsv_tts import _native as N
from gsv_tts import _synthetic as syn
from gsv_tts.GPT_SoVITS.GPT.t2s_model_b200 import Text2SemanticDecoder
from gsv_tts.GPT_SoVITS.SoVITS.models_b200 import FlowDecoder
from oracle.gpt_oracle import GptOracle
from oracle.vocoder_oracle import VocoderOracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def rounded(sd, dtype):
    """The weights the kernels actually see: storage dtype, read back as fp32 for the oracle."""
    return {k: v.to(dtype).float() for k, v in sd.items()}


def build_gpt(cfg, sd, dtype, dev, gpt_cache):
    m = Text2SemanticDecoder(cfg)
    m.load_state_dict(sd)
    m.eval()
    m.initialize_runtime(dtype, dev, gpt_cache)
    return m


def gpt_teacher_forced_error(cfg, name, dtype, dev):
    """Prefill + forced decode steps through the product kernels with the raw-logits trace on;
    compare with (a) the oracle on the dtype-rounded weights and (b) the reference's golden."""
    g = golden(f"gpt_{name}.npz")
    sd = syn.gpt_state_dict(cfg, 0, float(g["eos_boost"]))
    max_seq = int(g["max_seq"])
    m = build_gpt(cfg, sd, dtype, dev, [(1, max_seq)])
    forced = torch.from_numpy(g["forced"]).to(torch.int32).to(dev)
    n = forced.numel()
    V = cfg["model"]["vocab_size"]
    trace = torch.zeros(n + 1, V, dtype=torch.float32, device=dev)
    lib = N.lib()
    N.check(lib.gsv_gpt_set_forced(m._ctx, forced.data_ptr(), n))
    N.check(lib.gsv_gpt_set_logits_trace(m._ctx, trace.data_ptr(), n + 1))
    x, y, bert = (torch.from_numpy(g[k]) for k in ("x", "y", "bert"))
    bert16 = bert.to(dtype)
    m._single_setup(x, y, bert16, 15, 1.0, 1.0, 1.35, 10, None)
    m._decode(n)
    torch.cuda.synchronize()
    got = trace.cpu().numpy()
    N.check(lib.gsv_gpt_set_forced(m._ctx, None, 0))
    N.check(lib.gsv_gpt_set_logits_trace(m._ctx, None, 0))
    # oracle on the same rounded weights / rounded bert features
    orc = GptOracle(rounded(sd, dtype), cfg)
    K, Vc, kv_len = orc.new_cache(1, max_seq)
    h = orc.prefill(x, y, bert16.float(), K, Vc, kv_len)
    rows = [orc.logits(h.unsqueeze(0))[0]]
    for t in g["forced"].tolist():
        xin = orc.embed_next(torch.tensor([t]), kv_len - x.shape[0])
        rows.append(orc.logits(orc.decode_step(xin, K, Vc, kv_len))[0])
    want = torch.stack(rows).numpy()
    return {"vs_oracle": float(np.abs(got - want).max()), "vs_golden": float(np.abs(got - g["tf_logits"]).max()),
            "per_row_vs_oracle": np.abs(got - want).max(1), "model": m, "launches": int(lib.gsv_gpt_launch_count(m._ctx))}


def reference_noise_rows(seed, n_rows, V):
    """The Exp(1) draws the reference consumes under torch.manual_seed(seed): one [1,V-1] draw for
    the first token, then [1,V] per step (GPT/utils.py:8; t2s_model.py:417, 447)."""
    torch.manual_seed(seed)
    rows = torch.ones(n_rows, V, dtype=torch.float32)
    rows[0, : V - 1] = torch.empty(1, V - 1).exponential_(1)[0]
    for i in range(1, n_rows):
        rows[i] = torch.empty(1, V).exponential_(1)[0]
    return rows


class RowNoise:
    """Oracle-side reader of the same rows."""

    def __init__(self, rows):
        self.rows, self.i = rows, 0

    def __call__(self, shape):
        r = self.rows[self.i, : shape[-1]].view(shape)
        self.i += 1
        return r


def build_vocoder(key, dtype, dev):
    model = syn.SOVITS_MODEL[key]
    sd = syn.sovits_flow_dec_state_dict(model, 0)
    fd = FlowDecoder(**model)
    fd.load_state_dict(sd)
    fd.initialize_runtime(dtype, dev, [])
    return fd, sd, model


def folded_rounded_vocoder_sd(sd, dtype):
    """Weight-norm folded in fp32 (as the host loader does), then rounded to the storage dtype."""
    from oracle.vocoder_oracle import fold_weight_norm
    out = {}
    for k, v in sd.items():
        if k.endswith("weight_v"):
            base = k[: -len("weight_v")]
            out[base + "weight"] = fold_weight_norm(sd[base + "weight_g"], v).to(dtype).float()
        elif k.endswith("weight_g"):
            continue
        else:
            out[k] = v.to(dtype).float()
    return out


def vocoder_error(name, key, dtype, dev):
    g = golden(f"vocoder_{name}.npz")
    fd, sd, model = build_vocoder(key, dtype, dev)
    z_p, mask, ge = (torch.from_numpy(g[k]) for k in ("z_p", "mask", "ge"))
    audio, z = fd.flow_dec(z_p.to(dev), mask.to(dev), ge.to(dev), return_z=True)
    torch.cuda.synchronize()
    audio, z = audio.float().cpu().numpy(), z.float().cpu().numpy()
    vo = VocoderOracle(folded_rounded_vocoder_sd(sd, dtype), model)
    zin, gin_ = z_p.to(dtype).float(), ge.to(dtype).float()
    z_o = vo.flow_reverse(zin, mask, gin_)
    a_o = vo.generator(z_o * mask, gin_)
    return {"z_vs_oracle": float(np.abs(z - z_o.numpy()).max()), "audio_vs_oracle": float(np.abs(audio - a_o.numpy()).max()),
            "z_vs_golden": float(np.abs(z - g["z"]).max()), "audio_vs_golden": float(np.abs(audio - g["audio"]).max()),
            "ref16_vs_golden": float(np.abs(g["audio_fp16"] - g["audio"]).max()) if "audio_fp16" in g else None,
            "launches": fd.launch_count()}
