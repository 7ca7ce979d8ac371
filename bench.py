#!/usr/bin/env python
"""bench.py -- headline benchmark of the two hot paths (contract: see the task statement / DESIGN.md).

Workload at every N (weak scaling, one process per GPU, independent utterances per rank):
BASELINE.json configs[1] -- V2Pro, batch=1 streaming, seq_len <= 512:
    prompt Nx=64 phonemes + Ny=100 prompt tokens -> prefill; 200 generated semantic tokens (EOS masked)
    in 8 stream chunks of 25 (kv 164 -> 364, gpt_cache=[(1,512)]); after every chunk the SoVITS
    flow + HiFi-GAN vocoder turns 50 (first) / 55 (later) latent frames into 32 kHz audio.
One "step" = one such utterance.  metric = generated AR tokens per second over the whole step
(prefill + decode + vocoder), i.e. tokens / end-to-end time, with RTF = time / audio seconds.

Arms:
    (default)          the CUDA path through libgsv_b200.so
    --impl reference   the reference algorithm's CPU path (oracle port; /root/reference cannot travel to
                       the GPU box and the reference forces fp32 on CPU, gsv_tts/TTS.py:74-76)
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "gsv-tts-lite_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

NX, NY, N_TOK, CHUNK = 64, 100, 200, 25
DECODE_SMS = 128                           # SMs of the single-sequence decode kernel when the vocoder overlaps it
FRAMES_FIRST, FRAMES_NEXT = 50, 55        # sovits_cache=[50,55] (reference README_EN.md:214)
WORKLOAD = "V2Pro batch=1 streaming: prefill 64+100, 200 tokens in 8 chunks of 25, flow+HiFi-GAN 50/55 frames per chunk"
METRIC = "AR tokens/sec (V2Pro, batch=1 streaming, end-to-end incl. prefill + vocoder)"


def synth_inputs(seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 732, (1, NX), generator=g)
    y = torch.randint(0, 1024, (1, NY), generator=g)
    bert = torch.zeros(1, NX, 1024)                        # JA/EN text: bert features are zeros (TextProcessor.py:100)
    zs = [torch.randn(1, 192, FRAMES_FIRST if i == 0 else FRAMES_NEXT, generator=g) for i in range(N_TOK // CHUNK)]
    ge = torch.randn(1, 1024, 1, generator=g)
    return x, y, bert, zs, ge


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (via NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.05)
        except Exception as e:            # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_reference_step(n_tok=N_TOK, threads=None):
    """One bounded sample of the workload on the host cores with the oracle port of the reference
    algorithm (fp32, all cores).  Returns (seconds, tokens, description)."""
    from gsv_tts import _synthetic as syn
    from oracle.gpt_oracle import GptOracle
    from oracle.vocoder_oracle import VocoderOracle
    torch.set_num_threads(threads or os.cpu_count())
    torch.set_grad_enabled(False)
    orc = GptOracle(syn.gpt_state_dict(syn.GPT_CONFIG, 0), syn.GPT_CONFIG)
    model = syn.SOVITS_MODEL["v2Pro"]
    vo = VocoderOracle(syn.sovits_flow_dec_state_dict(model, 0), model)
    x, y, bert, zs, ge = synth_inputs(1234)
    torch.manual_seed(0)
    t0 = time.perf_counter()
    orc.infer(x[0], y[0], bert[0], max_seq=512, force_steps=n_tok)
    n_chunks = max(1, n_tok // CHUNK)
    for z in zs[:n_chunks]:
        vo.flow_dec(z, torch.ones(1, 1, z.shape[-1]), ge)
    dt = time.perf_counter() - t0
    return dt, n_tok, f"1 utterance: prefill {NX}+{NY}, {n_tok} tokens, {n_chunks} vocoder chunks, fp32, {torch.get_num_threads()} threads"


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    n_tok = 100      # bounded sample: half an utterance per step keeps K steps within minutes
    for _ in range(min(args.warmup, 1)):
        cpu_reference_step(n_tok)
    times = []
    for _ in range(args.steps):
        dt, ntok, desc = cpu_reference_step(n_tok)
        times.append(dt)
    total = sum(times)
    val = args.steps * n_tok / total
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": desc},
        "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="vocoder chunks on the decode stream (reference order) instead of a second stream")
    ap.add_argument("--no-extra", action="store_true", help="skip the batch 8/32 decode and vocoder-only extras")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch.distributed as dist
    from gsv_tts import _native as N
    from gsv_tts import _synthetic as syn
    from gsv_tts.GPT_SoVITS.GPT.t2s_model_b200 import Text2SemanticDecoder
    from gsv_tts.GPT_SoVITS.SoVITS.models_b200 import FlowDecoder

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback on the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float16

    # ---- weights: rank 0 materialises the synthetic checkpoint, NCCL broadcasts it ----
    gsd = syn.gpt_state_dict(syn.GPT_CONFIG, 0)
    model = syn.SOVITS_MODEL["v2Pro"]
    vsd = syn.sovits_flow_dec_state_dict(model, 0)
    if world > 1:
        from gsv_tts import _shard
        if rank != 0:       # only rank 0's copy counts: the others start from garbage and receive the broadcast
            gsd = {k: torch.empty_like(v) for k, v in gsd.items()}
            vsd = {k: torch.empty_like(v) for k, v in vsd.items()}
        gsd = _shard.broadcast_state_dict(gsd, 0, dev)
        vsd = _shard.broadcast_state_dict(vsd, 0, dev)
    gpt = Text2SemanticDecoder(syn.GPT_CONFIG)
    gpt.load_state_dict(gsd)
    gpt.initialize_runtime(dtype, dev, [(1, 512)])
    voc = FlowDecoder(**model)
    voc.load_state_dict(vsd)
    voc.initialize_runtime(dtype, dev, [FRAMES_FIRST, FRAMES_NEXT])
    lib = N.lib()

    x, y, bert, zs, ge = synth_inputs(1234 + rank)
    xd, yd, bertd = x.to(dev), y.to(dev), bert.to(dev, dtype)
    zsd = [z.to(dev, dtype) for z in zs]
    masks = [torch.ones(1, 1, z.shape[-1], device=dev, dtype=dtype) for z in zs]
    ged = ge.to(dev, dtype)
    # pinned host copies for the end-to-end arm
    xh, yh, berth = x.pin_memory(), y.pin_memory(), bert.to(dtype).pin_memory()
    zsh = [z.to(dtype).pin_memory() for z in zs]
    geh = ge.to(dtype).pin_memory()
    audio_host = torch.empty(N_TOK // CHUNK, FRAMES_NEXT * 640, dtype=dtype).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)       # > 126 MB L2

    gpt.debug_seed = 99
    stream = torch.cuda.current_stream(dev)
    overlap = not args.no_overlap
    side = torch.cuda.Stream(dev)
    if overlap:
        gpt.set_decode_sms(DECODE_SMS)
    dec_ms = []         # per decode launch (25 tokens), CUDA events on the launching stream

    ttft_marks = None

    def utterance(device_resident: bool, time_decode: bool):
        """prefill + 8 x (25-token persistent decode launch + vocoder chunk)."""
        nonlocal ttft_marks
        if device_resident:
            gx, gy, gb, gz, gg = xd, yd, bertd, zsd, ged
        else:
            gx, gy, gb = xh.to(dev, non_blocking=True), yh.to(dev, non_blocking=True), berth.to(dev, non_blocking=True)
            gz = [z.to(dev, non_blocking=True) for z in zsh]
            gg = geh.to(dev, non_blocking=True)
        gpt._single_setup(gx, gy, gb, 15, 1.0, 1.0, 1.35, 10, N_TOK)
        n_chunks = N_TOK // CHUNK

        def decode_chunk():
            if time_decode:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            gpt._decode(CHUNK)
            if time_decode:
                e1.record(stream)
                dec_ms.append((e0, e1))

        if overlap:
            side.wait_stream(stream)                       # the host->device copies above
        decode_chunk()
        for c in range(n_chunks):
            gpt._read(1)                                   # tokens of chunk c on the host (stream sync)
            if overlap:
                # gsv_tts.TTS.infer_features_stream: chunk c+1 decodes (on 128 SMs) while the vocoder of chunk c runs
                # on a second stream; the first chunk's vocoder is enqueued BEFORE the next decode
                if c + 1 < n_chunks and c > 0:
                    decode_chunk()
                ctx_mgr = torch.cuda.stream(side)
            else:
                ctx_mgr = contextlib.nullcontext()
            with ctx_mgr:
                audio = voc.flow_dec(gz[c], masks[c], gg)
                if not device_resident:
                    audio_host[c, : audio.shape[-1]].copy_(audio[0, 0], non_blocking=True)
                    if c == 0 and ttft_marks is not None:
                        # time to first audio: host call -> first 1.6 s chunk of samples in host memory
                        torch.cuda.current_stream(dev).synchronize()
                        ttft_marks.append(time.perf_counter())
            if (not overlap or c == 0) and c + 1 < n_chunks:
                decode_chunk()                             # behind the first chunk's vocoder: it gets the whole GPU (time to first audio)
        if overlap:
            stream.wait_stream(side)                       # the step ends when the last chunk of audio exists
        if not device_resident:
            stream.synchronize()
        return int(gpt._h_ngen[0]) - 1

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(device_resident, steps, collect):
        evs = []
        l0 = int(lib.gsv_gpt_launch_count(gpt._ctx)) + voc.launch_count()
        barrier()
        wall0 = time.perf_counter()
        for _ in range(steps):
            flush.zero_()                                    # evict L2 between timed steps (outside the events)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            n = utterance(device_resident, collect)
            b.record(stream)
            evs.append((a, b))
            assert n == N_TOK, n
        barrier()
        wall = time.perf_counter() - wall0
        ms = sum(a.elapsed_time(b) for a, b in evs)
        launches = int(lib.gsv_gpt_launch_count(gpt._ctx)) + voc.launch_count() - l0
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall, launches

    for _ in range(args.warmup):
        utterance(True, False)
    sampler = ClockSampler(local)
    sampler.start()
    ms_total, wall, launches = timed(True, args.steps, True)
    decode_launch_ms = [a.elapsed_time(b) for a, b in dec_ms]
    ms_e2e, wall_e2e, _ = timed(False, args.steps, False)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    # TTFT (BASELINE config 2): separate, untimed-region measurement so the extra sync does not perturb `e2e`
    ttfts = []
    for _ in range(5):
        flush.zero_()
        torch.cuda.synchronize(dev)
        ttft_marks = []
        t_call = time.perf_counter()
        utterance(False, False)
        ttfts.append((ttft_marks[0] - t_call) * 1e3)
    ttft_marks = None
    ttfts.sort()
    ttft_ms = ttfts[len(ttfts) // 2]

    tokens = args.steps * N_TOK * world
    value = tokens / (ms_total / 1e3)
    e2e_value = tokens / (ms_e2e / 1e3)
    audio_s = N_TOK * 0.04
    # ---- roofline of the dominant kernel: the small-batch persistent decode kernel (HBM bound) ----
    pk, pk_kind = peaks()
    w_bytes = (24 * (12 * 512 * 512 + 13 * 512) + 1025 * 512) * 2
    kv_mean = NX + NY + (N_TOK + 1) / 2.0
    bytes_per_token = w_bytes + 49152 * kv_mean                 # weights once + 49 152 B per live position
    mean_launch_ms = sum(decode_launch_ms) / len(decode_launch_ms)
    achieved = bytes_per_token * CHUNK / (mean_launch_ms / 1e3) / 1e9
    traffic = None      # dram bytes per launch of the same kernel from the committed ncu --set full capture
    try:
        with open(os.path.join(ROOT, "profiles", "r01_decode_traffic.json")) as f:
            tj = json.load(f)
        traffic = (tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]) * CHUNK / tj["tokens_per_launch"]
    except Exception:
        pass
    roofline = {"kernel": "gpt_decode_ll_kernel", "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk_kind,
                "bytes_per_launch": bytes_per_token * CHUNK, "launch_ms": mean_launch_ms,
                "decode_only_tok_s": CHUNK / (mean_launch_ms / 1e3)}

    extra = {}
    if rank == 0 and not args.no_extra:
        extra = extras(gpt, voc, dev, dtype, lib, N, syn)

    cpu_base = None
    if rank == 0 and not args.no_cpu_baseline:
        dt, ntok, desc = cpu_reference_step(50)
        cpu_base = {"value": ntok / dt, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port", "sample": desc}

    if rank == 0:
        h2d = sum(t.numel() * t.element_size() for t in [xh, yh, berth, geh] + zsh)
        d2h = sum(z.shape[-1] * 640 * 2 for z in zs) + (N_TOK // CHUNK) * (512 * 4 + 8)
        line = {
            "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "256 MB flush between timed steps; 152 MB of weights (> L2) streamed per token",
                       "parallelism": f"{world} independent utterance streams, NCCL weight broadcast at load only",
                       "overlap": (f"vocoder of chunk c on a second stream while chunk c+1 decodes on {DECODE_SMS} SMs" if overlap
                                   else "none: decode and vocoder chunks back to back on one stream")},
            "rtf": (ms_total / 1e3 / args.steps) / audio_s, "ttft_ms": ttft_ms,
            "roofline": roofline, "cpu_baseline": cpu_base,
            "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "rtf": (ms_e2e / 1e3 / args.steps) / audio_s},
            "gpu_launches": launches, "clocks": sampler.summary(), "wall_s": wall + wall_e2e, "extra": extra,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def extras(gpt, voc, dev, dtype, lib, N, syn):
    """Secondary numbers (not the headline): decode tok/s at batch 8 / 32 and vocoder-only throughput."""
    import ctypes as C
    from gsv_tts.GPT_SoVITS.GPT.t2s_model_b200 import Text2SemanticDecoder
    out = {}
    try:
        m = Text2SemanticDecoder(syn.GPT_CONFIG)
        m.load_state_dict(syn.gpt_state_dict(syn.GPT_CONFIG, 0))
        m.initialize_runtime(dtype, dev, [(32, 512)])
        g = torch.Generator().manual_seed(7)
        for B in (4, 8, 16, 32):
            m._release_all()
            for s in range(B):
                samp = N.GptSampling(top_k=15, top_p=1.0, temperature=1.0, repetition_penalty=1.0, suppress_steps=0, suppress_first=0,
                                     max_new_tokens=0, mask_eos=1, max_kv=512, seed=s + 1)
                m._prefill(s, torch.randint(0, 732, (NX,), generator=g), torch.randint(0, 1024, (NY + 100,), generator=g),
                           torch.zeros(NX, 1024), samp)
            m._decode(8)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            m._decode(64)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            out[f"decode_batch{B}_tok_s"] = B * 64 / (ms / 1e3)
            out[f"decode_batch{B}_us_per_step"] = ms * 1e3 / 64
        del m
        # vocoder only: 2 s of audio (100 frames) x batch 8
        z = torch.randn(8, 192, 100, device=dev, dtype=dtype)
        mk = torch.ones(8, 1, 100, device=dev, dtype=dtype)
        ge = torch.randn(8, 1024, 1, device=dev, dtype=dtype)
        voc.flow_dec(z, mk, ge)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        voc.flow_dec(z, mk, ge)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out["vocoder_b8_100f_ms"] = ms
        out["vocoder_audio_s_per_s"] = 8 * 2.0 / (ms / 1e3)
        out["vocoder_tflops"] = 8 * 100 * (813.1e6 + 14.2e6) / (ms / 1e3) / 1e12
        # a quarter of BASELINE config 5 (10 s of tokens, batch 64): batch 16 x 500 frames
        z = torch.randn(16, 192, 500, device=dev, dtype=dtype)
        mk = torch.ones(16, 1, 500, device=dev, dtype=dtype)
        ge = torch.randn(16, 1024, 1, device=dev, dtype=dtype)
        voc.flow_dec(z, mk, ge)
        torch.cuda.synchronize()
        e0.record()
        voc.flow_dec(z, mk, ge)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out["vocoder_b16_500f_ms"] = ms
        out["vocoder_b16_500f_tflops"] = 16 * 500 * (813.1e6 + 14.2e6) / (ms / 1e3) / 1e12
        out["vocoder_b16_500f_audio_s_per_s"] = 16 * 10.0 / (ms / 1e3)
    except Exception as e:      # extras must never break the headline line
        out["error"] = f"{type(e).__name__}: {e}"
    try:
        # SURVEY.md 8d config 3 (GPT stage): 128 mixed requests through 32 slots of infer_batched (continuous batching;
        # Nx ~ U{40..120}, Ny ~ U{75..250}, length ~ U{50..250} via max_new), wall clock, both refill modes
        import time as _t
        m = Text2SemanticDecoder(syn.GPT_CONFIG)
        m.load_state_dict(syn.gpt_state_dict(syn.GPT_CONFIG, 0))
        m.initialize_runtime(dtype, dev, [(32, 1024)])
        g = torch.Generator().manual_seed(1234)
        xs, ys, bs, mx = [], [], [], []
        for _ in range(128):
            nx = int(torch.randint(40, 121, (1,), generator=g))
            ny = int(torch.randint(75, 251, (1,), generator=g))
            xs.append(torch.randint(0, 732, (nx,), generator=g).to(dev))
            ys.append(torch.randint(0, 1024, (ny,), generator=g).to(dev))
            bs.append(torch.zeros(nx, 1024, device=dev, dtype=dtype))
            mx.append(int(torch.randint(50, 251, (1,), generator=g)))
        m.debug_seed = 5
        m.infer_batched(xs[:40], ys[:40], bs[:40], max_new=[20] * 40)        # warm-up (weight re-tiling, kernel attributes)
        for overlap, tag in ((True, "batched128_tok_s"), (False, "batched128_serial_refill_tok_s")):
            m.overlap_refill = overlap
            m.debug_seed = 5
            torch.cuda.synchronize()
            t0 = _t.perf_counter()
            outs, _ = m.infer_batched(xs, ys, bs, max_new=mx)
            torch.cuda.synchronize()
            out[tag] = sum(int(o.numel()) for o in outs) / (_t.perf_counter() - t0)
        del m
    except Exception as e:
        out["error_batched"] = f"{type(e).__name__}: {e}"
    return out


if __name__ == "__main__":
    main()
