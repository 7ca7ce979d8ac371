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


# ---------------------------------------------------------------------------------------------------------------
# Multi-sequence parity: every slot of the batched decode kernels held to the oracle (per-slot hooks of the C ABI)
# ---------------------------------------------------------------------------------------------------------------
def ragged_requests(n, seed, nx=(8, 40), ny=(8, 50), bert_std=1.0):
    g = torch.Generator().manual_seed(seed)
    xs = [torch.randint(0, 732, (int(torch.randint(nx[0], nx[1], (1,), generator=g)),), generator=g) for _ in range(n)]
    ys = [torch.randint(0, 1024, (int(torch.randint(ny[0], ny[1], (1,), generator=g)),), generator=g) for _ in range(n)]
    bs = [bert_std * torch.randn(len(x), 1024, generator=g) for x in xs]
    return xs, ys, bs


def set_slot_hooks(m, slot, noise=None, forced=None, trace=None):
    N.check(N.lib().gsv_gpt_set_slot_hooks(
        m._ctx, slot, noise.data_ptr() if noise is not None else None, noise.shape[0] if noise is not None else 0,
        forced.data_ptr() if forced is not None else None, forced.numel() if forced is not None else 0,
        trace.data_ptr() if trace is not None else None, trace.shape[0] if trace is not None else 0, m._stream()))


def clear_slot_hooks(m):
    torch.cuda.synchronize()
    for s in range(m._max_slots):
        set_slot_hooks(m, s)
    torch.cuda.synchronize()


def multi_sequence_teacher_forced_error(cfg, dtype, dev, n_seq, n_steps, seed=5, max_seq=192, eos_boost=0.0, nx=(8, 40), ny=(8, 50)):
    """n_seq DIFFERENT sequences (ragged prompts, their own forced tokens) live together in n_seq slots: every slot's
    raw logits (prefill row + n_steps decode rows) against the oracle run on the same rounded weights."""
    sd = syn.gpt_state_dict(cfg, 0, eos_boost)
    V = cfg["model"]["vocab_size"]
    m = build_gpt(cfg, sd, dtype, dev, [(n_seq, max_seq)])
    xs, ys, bs = ragged_requests(n_seq, seed, nx, ny)
    g = torch.Generator().manual_seed(seed + 1)
    forced = [torch.randint(0, V - 1, (n_steps,), generator=g, dtype=torch.int32) for _ in range(n_seq)]   # never EOS
    forced_d = [f.to(dev) for f in forced]
    traces = [torch.zeros(n_steps + 1, V, dtype=torch.float32, device=dev) for _ in range(n_seq)]
    m._release_all()
    for s in range(n_seq):
        set_slot_hooks(m, s, forced=forced_d[s], trace=traces[s])
        samp = N.GptSampling(top_k=15, top_p=1.0, temperature=1.0, repetition_penalty=1.0, suppress_steps=0,
                             max_new_tokens=0, mask_eos=0, max_kv=max_seq, suppress_first=0, seed=s + 1)
        m._prefill(s, xs[s], ys[s], bs[s].to(dtype), samp)
    m._decode(n_steps)
    torch.cuda.synchronize()
    m._read(n_seq)
    got_tokens = [m._h_tokens[s, : n_steps].tolist() for s in range(n_seq)]
    got = [t.cpu() for t in traces]
    clear_slot_hooks(m)
    # oracle, all sequences as one batch with per-sequence kv_len
    orc = GptOracle(rounded(sd, dtype), cfg)
    K, Vc, kv_len = orc.new_cache(n_seq, max_seq)
    rows = [[] for _ in range(n_seq)]
    for s in range(n_seq):
        h = orc.prefill(xs[s], ys[s], bs[s].to(dtype).float(), K, Vc, kv_len, slot=s)
        rows[s].append(orc.logits(h.unsqueeze(0))[0])
    nxs = torch.tensor([len(x) for x in xs])
    for i in range(n_steps):
        tok = torch.tensor([int(forced[s][i]) for s in range(n_seq)])
        xin = orc.embed_next(tok, kv_len - nxs)
        lg = orc.logits(orc.decode_step(xin, K, Vc, kv_len))
        for s in range(n_seq):
            rows[s].append(lg[s])
    err = [float((got[s] - torch.stack(rows[s])).abs().max()) for s in range(n_seq)]
    forced_ok = all(got_tokens[s] == forced[s].tolist() for s in range(n_seq))
    return {"per_slot": err, "max": max(err), "forced_ok": forced_ok, "model": m}


class BatchedAudit:
    """Per-request noise rows in, per-request raw-logit traces out, attached to whichever slot a request lands in
    (``Text2SemanticDecoder._slot_audit``).  ``check`` then proves, request by request, that
    (1) the sampler is exact: the oracle's sample() on the kernel's own logits and the same noise rows reproduces the
        kernel's tokens at every step (s0, every returned token, the stopping sample);
    (2) the model is right at every step of the free run: the oracle teacher-forced with those tokens gives the traced
        logits within the 16-bit tolerance (prefill row included, whatever slot / launch / co-runners the request had);
    (3) the returned list is what the reference returns: s0 dropped, cut before EOS / at max_new."""

    def __init__(self, n_req, max_rows, V, dev, seed):
        g = torch.Generator().manual_seed(seed)
        self.noise = [torch.empty(max_rows, V).exponential_(1, generator=g) for _ in range(n_req)]
        self.noise_d = [n.to(dev) for n in self.noise]
        self.trace = [torch.full((max_rows, V), float("nan"), dtype=torch.float32, device=dev) for _ in range(n_req)]
        self.placed = []

    def __call__(self, slot, r):
        self.placed.append((slot, r))
        return self.noise_d[r], None, self.trace[r]

    def check(self, orc, cfg, xs, ys, bs16, results, order, max_new, max_seq, tol, top_k=15):
        from oracle.gpt_oracle import sample_token
        EOS = cfg["model"]["EOS"]
        kw = dict(top_k=top_k, top_p=1.0, temperature=1.0, repetition_penalty=1.0)
        byreq = {r: t.cpu().tolist() for t, r in zip(results, order.cpu().tolist())}
        worst = 0.0
        for r in range(len(xs)):
            trace = self.trace[r].cpu()
            n_rows = int((~torch.isnan(trace[:, 0])).sum())
            assert n_rows >= 1, f"request {r}: nothing traced"
            # (1) sampler replay on the kernel's logits
            toks = []
            for i in range(n_rows):
                lg = trace[i:i + 1].clone()
                if i == 0:
                    lg = lg[:, :-1]
                noise = lambda shape, i=i: self.noise[r][i, : shape[-1]].view(shape)
                t = int(sample_token(lg, None, noise=noise, **kw)[0])
                if max_new is not None and i > max_new[r]:
                    t = EOS                                     # forced stop (gsv_gpt_sampling.max_new_tokens)
                toks.append(t)
            body = toks[1:]
            cut = body.index(EOS) if EOS in body else len(body)
            assert byreq[r] == body[:cut], f"request {r}: returned tokens are not the sampler's on the kernel's own logits"
            kv_end = len(xs[r]) + len(ys[r]) + n_rows - 1
            assert cut == len(body) - 1 or kv_end >= max_seq - 1, f"request {r}: stopped without EOS or a full cache"
            # (2) oracle teacher-forced with the kernel's tokens
            K, Vc, kv_len = orc.new_cache(1, max_seq)
            h = orc.prefill(xs[r], ys[r], bs16[r].float(), K, Vc, kv_len)
            rows = [orc.logits(h.unsqueeze(0))[0]]
            for i in range(n_rows - 1):
                xin = orc.embed_next(torch.tensor([toks[i]]), kv_len - len(xs[r]))
                rows.append(orc.logits(orc.decode_step(xin, K, Vc, kv_len))[0])
            e = (trace[:n_rows] - torch.stack(rows)).abs().max(1).values
            worst = max(worst, float(e.max()))
            assert float(e.max()) < tol, f"request {r}: logits off by {float(e.max()):.3e} at step {int(e.argmax())} (rows {e.tolist()})"
        return worst
