"""Same-box GPU baseline: the REFERENCE's own CUDA path (FlashAttention decoder + CUDA graphs, and its SynthesizerTrn
flow + HiFi-GAN with the streaming CUDA graph) timed on this B200 with the same synthetic weights and inputs as
``bench.py`` (north_star: "beating the reference's own FlashAttn-on CUDA path on the same box").

Runs the UNMODIFIED reference modules from ``baseline/_ref`` (``baseline/install_ref.py``) through their public entry
points -- ``Text2SemanticDecoder.infer`` / ``infer_stream`` / ``infer_batched`` (t2s_model_flash_attn.py, SDPA variant
with ``--sdpa``) and ``Generator`` / ``ResidualCouplingBlock`` as ``SynthesizerTrn`` wires them -- and none of this
repo's kernels.  Prints one JSON object; ``bench.py`` does not import this file (the CPU reference arm stays the
driver's ``--impl reference``); its numbers are quoted in DESIGN.md section 6 and committed under ``profiles/``.

    python tools/ref_gpu_bench.py [--sdpa] [--dtype bf16|fp16] [--quick]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gsv-tts-lite_b200"))

from gsv_tts import _synthetic as syn  # noqa: E402
from oracle import ref_shim  # noqa: E402   (test infrastructure; this tool is not on the product path)


def no_eos_state_dict(cfg):
    """bench.py's GPT weights: EOS row zeroed so that the EOS logit (0) never reaches the top-15 of logits with std ~2
    and every sequence runs to the bucket length, as SURVEY.md 8d config 2 prescribes (EOS masked)."""
    sd = syn.gpt_state_dict(cfg, 0, 0.0)
    sd["ar_predict_layer.weight"][cfg["model"]["EOS"]] = 0.0
    return sd


def cuda_time(fn, n=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    out = None
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n, out


def bench_gpt(res, flash, dtype, quick):
    cfg = syn.GPT_CONFIG
    dev = torch.device("cuda:0")
    sd = no_eos_state_dict(cfg)
    NX, NY, S = 64, 100, 512
    tag = "flash" if flash else "sdpa"
    caches = [(1, S), (8, S), (32, S)]
    t0 = time.perf_counter()
    ref = ref_shim.build_reference_gpt(sd, cfg, dtype, "cuda:0", caches, flash=flash)
    torch.cuda.synchronize()
    res[f"{tag}_init_s"] = round(time.perf_counter() - t0, 2)
    g = torch.Generator().manual_seed(1234)
    x = torch.randint(0, 732, (1, NX), generator=g).to(dev)
    y = torch.randint(0, 1024, (1, NY), generator=g).to(dev)
    bert = torch.zeros(1, NX, 1024, dtype=dtype, device=dev)
    # ---- B = 1: infer() to the bucket length (512 - 164 = 348 steps)
    torch.manual_seed(1)
    ref.infer(x, y, bert)                                   # warm-up
    reps = 1 if quick else 3
    best = None
    for r in range(reps):
        torch.manual_seed(2 + r)
        ms, wall, out = cuda_time(lambda: ref.infer(x, y, bert))
        n = int(out.shape[-1])
        if best is None or wall < best[1]:
            best = (ms, wall, n)
    res[f"{tag}_b1_infer"] = {"tokens": best[2], "ms": round(best[1], 2), "tok_s": round(best[2] / best[1] * 1e3, 1),
                              "us_per_token": round(best[1] / max(best[2], 1) * 1e3, 1)}
    # ---- B = 1 streaming: time to the first 25-token chunk, and the whole stream
    torch.manual_seed(5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    first = None
    total = 0
    for chunk, final in ref.infer_stream(x, y, bert, stream_chunk=25, debug=False):
        torch.cuda.synchronize()
        if first is None:
            first = (time.perf_counter() - t0) * 1e3
        total = int(chunk.shape[-1]) if final else total
    res[f"{tag}_b1_stream"] = {"first_chunk_ms": round(first, 2), "total_ms": round((time.perf_counter() - t0) * 1e3, 2)}
    # ---- B = 8 / 32: infer_batched with B requests = B slots, every request to the bucket length
    for B in (8, 32):
        gg = torch.Generator().manual_seed(100 + B)
        xs = [torch.randint(0, 732, (NX,), generator=gg).to(dev) for _ in range(B)]
        ys = [torch.randint(0, 1024, (NY,), generator=gg).to(dev) for _ in range(B)]
        bs = [torch.zeros(NX, 1024, dtype=dtype, device=dev) for _ in range(B)]
        torch.manual_seed(7)
        if not quick:
            ref.infer_batched(xs, ys, bs)                    # warm-up
        torch.manual_seed(8)
        ms, wall, out = cuda_time(lambda: ref.infer_batched(xs, ys, bs))
        n = int(sum(len(t) for t in out[0]))
        res[f"{tag}_b{B}_infer_batched"] = {"tokens": n, "ms": round(wall, 2), "tok_s": round(n / wall * 1e3, 1),
                                            "us_per_step": round(wall / max(n / B, 1) * 1e3, 1)}
    # ---- config 3 in miniature: 128 mixed requests through 32 slots (same generator as tools/batched_throughput.py)
    if not quick:
        gg = torch.Generator().manual_seed(3)
        n_req = 128
        xs, ys, bs = [], [], []
        for _ in range(n_req):
            nx = int(torch.randint(40, 121, (1,), generator=gg))
            ny = int(torch.randint(75, 251, (1,), generator=gg))
            xs.append(torch.randint(0, 732, (nx,), generator=gg).to(dev))
            ys.append(torch.randint(0, 1024, (ny,), generator=gg).to(dev))
            bs.append(torch.zeros(nx, 1024, dtype=dtype, device=dev))
        torch.manual_seed(9)
        ms, wall, out = cuda_time(lambda: ref.infer_batched(xs, ys, bs))
        n = int(sum(len(t) for t in out[0]))
        res[f"{tag}_config3_128req_32slots"] = {"tokens": n, "ms": round(wall, 2), "tok_s": round(n / wall * 1e3, 1),
                                                "note": "every request runs to the 512 bucket (no EOS in synthetic weights)"}
    del ref
    torch.cuda.empty_cache()


def bench_vocoder(res, dtype, quick):
    dev = torch.device("cuda:0")
    for key in ("v2Pro",) if quick else ("v2Pro", "v2ProPlus"):
        model = syn.SOVITS_MODEL[key]
        sd = syn.sovits_flow_dec_state_dict(model, 0)
        flow, dec = ref_shim.build_reference_flow_dec(sd, model, dtype, "cuda:0")

        def run(z_p, mask, ge):
            z = flow(z_p, mask, ge)
            return dec(z * mask, g=ge)

        g = torch.Generator().manual_seed(777)
        for B, T, reps in ((1, 50, 20), (1, 500, 5), (16, 500, 3), (64, 500, 2)):
            if quick and B > 16:
                continue
            z_p = torch.randn(B, 192, T, generator=g).to(dev, dtype)
            mask = torch.ones(B, 1, T, device=dev, dtype=dtype)
            ge = torch.randn(B, model["gin_channels"], 1, generator=g).to(dev, dtype)
            try:
                run(z_p, mask, ge)
                ms, wall, _ = cuda_time(lambda: run(z_p, mask, ge), reps)
                entry = {"eager_ms": round(ms, 3)}
                if T <= 55:
                    # the reference replays a CUDA graph for T <= sovits_cache (models.py:322-369, 406-425)
                    s = torch.cuda.Stream()
                    s.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(s):
                        for _ in range(3):
                            run(z_p, mask, ge)
                    torch.cuda.current_stream().wait_stream(s)
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        o = run(z_p, mask, ge)
                    graph.replay()
                    ms, wall, _ = cuda_time(graph.replay, reps)
                    entry["graph_ms"] = round(ms, 3)
                audio_s = B * T / 50.0
                best = entry.get("graph_ms", entry["eager_ms"])
                entry["audio_s_per_s"] = round(audio_s / best * 1e3, 1)
                flops = (813.1e6 if key == "v2Pro" else 1828.4e6) + 14.2e6
                entry["tflops"] = round(flops * B * T / best / 1e9, 1)
            except torch.cuda.OutOfMemoryError:
                entry = {"oom": True}
                torch.cuda.empty_cache()
            res[f"voc_{key}_B{B}_T{T}"] = entry
        del flow, dec
        torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sdpa", action="store_true", help="also time the SDPA decoder (t2s_model.py)")
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--skip-voc", action="store_true")
    a = ap.parse_args()
    dtype = torch.bfloat16 if a.dtype == "bf16" else torch.float16
    res = {"what": "reference CUDA path on this box (unmodified modules from baseline/_ref)", "dtype": a.dtype,
           "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
    try:
        import flash_attn
        res["flash_attn"] = flash_attn.__version__
    except Exception as e:  # pragma: no cover
        res["flash_attn"] = f"unavailable: {e}"
    with torch.inference_mode():
        bench_gpt(res, True, dtype, a.quick)
        if a.sdpa:
            bench_gpt(res, False, dtype, a.quick)
        if not a.skip_voc:
            bench_vocoder(res, dtype, a.quick)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
