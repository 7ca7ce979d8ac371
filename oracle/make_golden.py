"""ORACLE support (test infrastructure): generate ``tests/golden/*.npz`` by running the
REFERENCE's own modules (imported from /root/reference through ``ref_shim``) on seeded
synthetic weights and inputs.  Run in the build container only:

    python oracle/make_golden.py

The reference cannot travel to the GPU box, so these small fixtures are what pins both
the oracle restatement (CPU tests) and the CUDA path (GPU tests).  Weights are not stored:
they are regenerated from the seed by ``gsv_tts._synthetic`` (same torch build everywhere).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gsv-tts-lite_b200"))

from gsv_tts import _synthetic as syn  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_grad_enabled(False)


def _inputs(seed, nx, ny):
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 732, (nx,), generator=g)
    y = torch.randint(0, 1024, (ny,), generator=g)
    bert = torch.randn(nx, 1024, generator=g)
    return x, y, bert


def gpt_teacher_forced(ref, x, y, bert, forced, max_seq):
    """Reference prefill + decode steps with the input token of every step forced, so a
    difference in one step's logits cannot cascade.  Returns logits [1+n, V] (row 0 =
    after prefill) and the hidden states."""
    xy_pos, mask = ref.process_single_data(x[None], y[None], bert[None])
    bucket = ref.cuda_graph_buckets[1][-1]
    bucket.kv_cache_len.fill_(0)
    # the reference allocates its cache with torch.empty (t2s_model.py:245-246); the SDPA path
    # multiplies masked garbage by zero weights, which is NaN if the garbage is NaN/Inf
    bucket.k_cache.zero_()
    bucket.v_cache.zero_()
    h = ref.t2s_transformer.process_prompt(xy_pos, bucket.k_cache, bucket.v_cache, bucket.kv_cache_len, mask)
    rows = [ref.ar_predict_layer(h[:, -1]).float()[0].clone()]
    hid = [h[0, -1].float().clone()]
    pe = (ref.ar_audio_position.alpha * ref.ar_audio_position.pe).transpose(0, 1)
    sdpa = hasattr(bucket, "decode_attn_mask") and bucket.decode_attn_mask is not None
    if sdpa:
        bucket.decode_attn_mask.fill_(False)
        bucket.decode_attn_mask[:, :, :, : int(bucket.kv_cache_len)] = True
    for t in forced.tolist():
        tok = torch.tensor([[t]])
        xin = ref.ar_audio_embedding(tok) * ref.ar_audio_position.x_scale + pe[bucket.kv_cache_len - x.shape[0]]
        if sdpa:
            bucket.decode_attn_mask[:, :, :, int(bucket.kv_cache_len)] = True
            hd = ref.t2s_transformer.decode_next_token(xin, bucket.k_cache, bucket.v_cache, bucket.kv_cache_len,
                                                       bucket.decode_attn_mask, bucket.batch_indices)
        else:
            hd = ref.t2s_transformer.decode_next_token(xin, bucket.k_cache, bucket.v_cache, bucket.kv_cache_len)
        rows.append(ref.ar_predict_layer(hd[:, -1]).float()[0].clone())
        hid.append(hd[0, -1].float().clone())
    return torch.stack(rows), torch.stack(hid)


def make_gpt(name, cfg, nx, ny, n_forced, max_seq, eos_boost, infer_seed, with_bf16):
    sd = syn.gpt_state_dict(cfg, seed=0, eos_boost=eos_boost)
    x, y, bert = _inputs(1234, nx, ny)
    forced = torch.randint(0, 1024, (n_forced,), generator=torch.Generator().manual_seed(99))
    ref = ref_shim.build_reference_gpt(sd, cfg, torch.float32, "cpu", [(1, max_seq)])
    logits, hidden = gpt_teacher_forced(ref, x, y, bert, forced, max_seq)
    out = dict(x=x.numpy(), y=y.numpy(), bert=bert.numpy(), forced=forced.numpy(),
               tf_logits=logits.numpy(), tf_hidden=hidden.numpy(),
               eos_boost=np.float32(eos_boost), max_seq=np.int64(max_seq), infer_seed=np.int64(infer_seed))
    # seeded free-running decode through the reference's own entry points
    torch.manual_seed(infer_seed)
    out["infer_tokens"] = ref.infer(x[None], y[None], bert[None])[0, 0].numpy()
    torch.manual_seed(infer_seed)
    chunks = [(c[0, 0].numpy().copy(), f) for c, f in
              ref.infer_stream(x[None], y[None], bert[None], stream_chunk=10, debug=False)]
    out["stream_lens"] = np.array([len(c) for c, _ in chunks])
    out["stream_final"] = chunks[-1][0]
    out["stream_first"] = chunks[0][0]
    if with_bf16:
        # the reference's own 16-bit CPU path, to calibrate tolerances (SURVEY.md 7 "rounding order")
        for dt, tag in ((torch.bfloat16, "bf16"), (torch.float16, "fp16")):
            r16 = ref_shim.build_reference_gpt(sd, cfg, dt, "cpu", [(1, max_seq)])
            lg16, _ = gpt_teacher_forced(r16, x, y, bert.to(dt), forced, max_seq)
            out[f"tf_logits_{tag}"] = lg16.numpy()
    np.savez_compressed(os.path.join(OUT, f"gpt_{name}.npz"), **out)
    print(name, "logits", logits.shape, "std", float(logits.std()), "infer n", len(out["infer_tokens"]),
          "stream", out["stream_lens"])


def make_gpt_batched(name, cfg, n_req, slots, max_seq, eos_boost):
    """Reference ``infer_batched`` end to end (continuous batching through ``slots`` slots).
    Only the token *lengths* depend on its 5-step check schedule; stored for the contract
    checks (s0 excluded, cut at first EOS, every request returned exactly once)."""
    sd = syn.gpt_state_dict(cfg, seed=0, eos_boost=eos_boost)
    ref = ref_shim.build_reference_gpt(sd, cfg, torch.float32, "cpu", [(slots, max_seq)])
    g = torch.Generator().manual_seed(4321)
    xs, ys, bs = [], [], []
    for r in range(n_req):
        nx = int(torch.randint(20, 50, (1,), generator=g))
        ny = int(torch.randint(20, 60, (1,), generator=g))
        xs.append(torch.randint(0, 732, (nx,), generator=g))
        ys.append(torch.randint(0, 1024, (ny,), generator=g))
        bs.append(torch.randn(nx, 1024, generator=g))
    torch.manual_seed(5)
    toks, order = ref.infer_batched(xs, ys, bs)
    out = dict(n_req=np.int64(n_req), slots=np.int64(slots), max_seq=np.int64(max_seq), eos_boost=np.float32(eos_boost),
               order=order.numpy(), lens=np.array([len(t) for t in toks]))
    for i, t in enumerate(toks):
        out[f"tok{i}"] = t.numpy()
    for r in range(n_req):
        out[f"x{r}"], out[f"y{r}"], out[f"b{r}"] = xs[r].numpy(), ys[r].numpy(), bs[r].numpy().astype(np.float16)
    np.savez_compressed(os.path.join(OUT, f"gpt_{name}.npz"), **out)
    print(name, "order", order.tolist(), "lens", out["lens"].tolist())


def make_vocoder(name, key, B, T, masked_tail, ge_per_frame=False, with16=True):
    model = syn.SOVITS_MODEL[key]
    sd = syn.sovits_flow_dec_state_dict(model, seed=0)
    flow, dec = ref_shim.build_reference_flow_dec(sd, model, torch.float32, "cpu")
    g = torch.Generator().manual_seed(777)
    z_p = torch.randn(B, 192, T, generator=g)
    mask = torch.ones(B, 1, T)
    if masked_tail and B > 1:
        mask[1, :, T - masked_tail:] = 0
    ge = torch.randn(B, model["gin_channels"], T if ge_per_frame else 1, generator=g)
    z = flow(z_p, mask, ge)
    o = dec(z * mask, g=ge)
    out = dict(z_p=z_p.numpy(), mask=mask.numpy(), ge=ge.numpy(), z=z.numpy(), audio=o.numpy())
    if with16:
        for dt, tag in ((torch.bfloat16, "bf16"), (torch.float16, "fp16")):
            f16, d16 = ref_shim.build_reference_flow_dec(sd, model, dt, "cpu")
            z16 = f16(z_p.to(dt), mask.to(dt), ge.to(dt))
            out[f"audio_{tag}"] = d16(z16 * mask.to(dt), g=ge.to(dt)).float().numpy()
            out[f"z_{tag}"] = z16.float().numpy()
    np.savez_compressed(os.path.join(OUT, f"vocoder_{name}.npz"), **out)
    msg = f"{name}: z std {float(z.std()):.3f} audio std {float(o.std()):.3f} max {float(o.abs().max()):.3f}"
    if with16:
        msg += f" | ref16-vs-32 audio err bf16 {np.abs(out['audio_bf16'] - out['audio']).max():.2e} fp16 {np.abs(out['audio_fp16'] - out['audio']).max():.2e}"
    print(msg)


def make_encp(name, key):
    """Row f-1: outputs of the reference's own quantizer / ge_to512 / TextEncoder on synthetic weights:
    plain call, speed != 1, two streaming chunks with the cross-fade state, and a slice_indices call (MRTE window)."""
    import torch.nn.functional as F
    model = dict(syn.SOVITS_MODEL[key])
    M = ref_shim.sovits_models()
    net = M.SynthesizerTrn(1025, 32, n_speakers=300, **model).eval()
    sd = syn.sovits_encp_state_dict(model, 0)
    ref = net.state_dict()
    want = {k for k in ref if k.startswith(("enc_p.", "ge_to512."))} | {"quantizer.vq.layers.0._codebook.embed"}
    assert set(sd) == want, set(sd) ^ want                       # the synthetic writer follows the reference's key set ...
    assert all(sd[k].shape == ref[k].shape for k in sd)           # ... and shapes
    net.load_state_dict(sd, strict=False)
    g = torch.Generator().manual_seed(4321)
    n, nt = 13, 9
    codes = torch.randint(0, 1024, (1, 1, n), generator=g)
    text = torch.randint(0, 732, (1, nt), generator=g)
    ge = torch.randn(1, model["gin_channels"], 1, generator=g)
    noise = torch.randn(1, model["inter_channels"], 2 * n, generator=g)
    q = F.interpolate(net.quantizer.decode(codes), size=2 * n, mode="nearest")
    gin = net.ge_to512(ge.transpose(2, 1)).transpose(2, 1) if net.is_v2pro else ge
    out = dict(codes=codes.numpy(), text=text.numpy(), ge=ge.numpy(), noise=noise.numpy(), quantized=q.numpy())
    m, logs, mask = net.enc_p.infer(q, text, gin, 1)
    out.update(m_p=m.numpy(), logs_p=logs.numpy(), z_p=(m + noise * torch.exp(logs) * 0.5).numpy())
    m, logs, mask = net.enc_p.infer(q, text, gin, 1.3)
    out.update(m_p_speed=m.numpy(), logs_p_speed=logs.numpy(), speed=np.float32(1.3))
    sl = torch.tensor([[2, 7]])
    m, logs, mask = net.enc_p.infer(q, text, gin, 1, slice_indices=sl)
    out.update(m_p_slice=m.numpy(), slice_indices=sl.numpy())
    net.enc_p.y_overlap = None
    chunks = [(8, 0), (13, 5)]                                   # (codes so far, valid_start_idx), overlap_len 5
    for i, (nn_, vs) in enumerate(chunks):
        m, logs, mask = net.enc_p.infer(q[:, :, : 2 * nn_], text, gin, 1, True, vs, 5)
        out[f"m_p_stream{i}"] = m.numpy()
    out["stream_chunks"] = np.array(chunks)
    np.savez_compressed(os.path.join(OUT, f"encp_{name}.npz"), **out)
    print(f"encp {name}: m_p std {float(out['m_p'].std()):.3f} logs mean {float(out['logs_p'].mean()):.3f}")


def reference_glue_methods():
    """The bodies of the reference's own glue methods, compiled from the source text of gsv_tts/TTS.py (the module cannot be
    imported here: it needs av, pysbd, ...).  Returns (namespace of plain functions taking ``self`` first, fake self)."""
    import ast
    import types
    import torch.nn.functional as F
    src = open(os.path.join(ref_shim.REF_ROOT, "gsv_tts", "TTS.py")).read()
    want = {"_viterbi_monotonic", "_sola_algorithm", "_find_head_threshold_offsets", "_find_tail_threshold_offsets", "_get_subtitles"}
    fns = [n for cls in ast.parse(src).body if isinstance(cls, ast.ClassDef) and cls.name == "TTS"
           for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in want]
    assert {f.name for f in fns} == want
    ns = {"torch": torch, "F": F, "np": np}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "reference_TTS_glue", "exec"), ns)
    me = types.SimpleNamespace(tts_config=types.SimpleNamespace(device=torch.device("cpu"), dtype=torch.float32), samplerate=32000, sovits_hz=50)
    return ns, me


def glue_cases(seed=0):
    """Seeded inputs for the glue functions: attention maps with a monotonic drift (three sizes, one with heads stuck on
    the null key), a waveform with leading / trailing silence, two overlapping chunks for SOLA."""
    g = torch.Generator().manual_seed(seed)
    attn = []
    for T_, N_ in ((40, 12), (7, 3), (123, 50), (400, 97)):
        logits = torch.randn(4, T_, N_, generator=g) * 2
        pos = torch.arange(T_).float()[:, None] / T_ * N_
        logits = logits - 0.5 * (torch.arange(N_)[None, :] - pos).abs()[None]
        if T_ == 123:
            logits[1:, 20:30, N_ - 1] += 30.0          # three heads look at the null key for ten frames
            logits[0, 24:27, N_ - 1] += 30.0           # ... all four for three of them: the default distribution is used
        attn.append(torch.softmax(logits, -1))
    n = 70000
    t = torch.arange(n)
    audio = torch.randn(n, generator=g) * 0.05 * ((t > 9000) & (t < 61000)).float() + torch.randn(n, generator=g) * 0.002
    f1 = torch.randn(3200, generator=g)
    f2 = torch.randn(9000, generator=g) * 0.5
    f2[137:3337] += 2 * f1
    return attn, audio, f1, f2


def make_glue():
    ns, me = reference_glue_methods()
    attn, audio, f1, f2 = glue_cases()
    out = {}
    for i, a in enumerate(attn):
        out[f"attn{i}"] = a.numpy()
        out[f"assign{i}"] = ns["_viterbi_monotonic"](me, a).numpy()
    out["audio"] = audio.numpy()
    out["head_offset"] = np.int64(ns["_find_head_threshold_offsets"](me, audio))
    out["tail_offset"] = np.int64(ns["_find_tail_threshold_offsets"](me, audio))
    silent = torch.zeros(5000)
    out["head_offset_silent"] = np.int64(ns["_find_head_threshold_offsets"](me, silent))
    out["tail_offset_silent"] = np.int64(ns["_find_tail_threshold_offsets"](me, silent))
    r, off = ns["_sola_algorithm"](me, f1.view(1, 1, -1), f2.view(1, 1, -1), 3200)
    out.update(sola_f1=f1.numpy(), sola_f2=f2.numpy(), sola_out=r[0, 0].numpy(), sola_offset=np.int64(int(off)))
    np.savez_compressed(os.path.join(OUT, "glue.npz"), **out)
    print("glue: assign sizes", [int(out[f"assign{i}"].shape[0]) for i in range(len(attn))], "head", int(out["head_offset"]),
          "tail", int(out["tail_offset"]), "sola offset", int(out["sola_offset"]))


def make_sovits_aux():
    """get_ge / extract_latent of the reference's own SynthesizerTrn (models.py:371-378, 431-434) on the tiny synthetic
    checkpoint (v2 and v2Pro): tests/golden/sovits_aux.npz, read by tests/test_aux_cpu.py."""
    M = ref_shim.sovits_models()
    g = torch.Generator().manual_seed(17)
    refer, sv, ssl = torch.randn(1, 1025, 37, generator=g), torch.randn(1, 20480, generator=g) * 0.1, torch.randn(1, 768, 46, generator=g)
    out = {}
    for version in ("v2", "v2Pro"):
        model = dict(syn.SOVITS_MODEL["tiny"], version=version)
        sd = dict(syn.sovits_flow_dec_state_dict(model, 0))
        sd.update(syn.sovits_encp_state_dict(model, 0))
        sd.update(syn.sovits_aux_state_dict(model, 0))
        sd["quantizer.vq.layers.0._codebook.inited"] = torch.Tensor([True])    # a trained checkpoint: no k-means re-initialisation
        ref = M.SynthesizerTrn(1025, 32, n_speakers=300, **model).eval()
        ref.load_state_dict(sd, strict=False)
        out[f"ge_{version}"] = ref.get_ge(refer, sv if version == "v2Pro" else None).numpy()
        out[f"codes_{version}"] = ref.extract_latent(ssl).numpy()
    np.savez_compressed(os.path.join(OUT, "sovits_aux.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["gpt", "batched", "voc", "encp", "glue", "aux"]
    if "aux" in which:
        make_sovits_aux()
    if "glue" in which:
        make_glue()
    if "gpt" in which:
        make_gpt("tiny", syn.GPT_CONFIG_TINY, nx=40, ny=30, n_forced=12, max_seq=256, eos_boost=6.0, infer_seed=7, with_bf16=True)
        make_gpt("full", syn.GPT_CONFIG, nx=48, ny=60, n_forced=8, max_seq=512, eos_boost=6.0, infer_seed=11, with_bf16=True)
    if "batched" in which:
        make_gpt_batched("tiny_batched", syn.GPT_CONFIG_TINY, n_req=7, slots=4, max_seq=256, eos_boost=6.0)
    if "encp" in which:
        make_encp("v2pro", "v2Pro")
        make_encp("v2", "v2")
    if "voc" in which:
        make_vocoder("tiny", "tiny", B=2, T=20, masked_tail=5)
        make_vocoder("tiny_ge_t", "tiny", B=1, T=16, masked_tail=0, ge_per_frame=True)
        make_vocoder("v2pro", "v2Pro", B=1, T=12, masked_tail=0)
        make_vocoder("v2proplus", "v2ProPlus", B=1, T=8, masked_tail=0)
        make_vocoder("v2", "v2", B=1, T=8, masked_tail=0)


if __name__ == "__main__":
    # the reference's static buffers are inference tensors
    with torch.inference_mode():
        main()
