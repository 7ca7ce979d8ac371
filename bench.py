#!/usr/bin/env python
"""bench.py -- headline benchmark of the two hot paths (contract: see the task statement / DESIGN.md section 6).

Headline workload at every N (weak scaling, one process per GPU, independent utterances per rank):
BASELINE.json configs[1] -- V2Pro, batch=1 streaming, seq_len <= 512:
    prompt Nx=64 phonemes + Ny=100 prompt tokens -> prefill; 200 generated semantic tokens (EOS masked)
    in 8 stream chunks of 25 (kv 164 -> 364, gpt_cache (1,512)); after every chunk ``vq_model.decode(stream_mode=True)``
    re-encodes the token prefix (prior encoder: 12 relative-position layers + MRTE), and the reverse flow + HiFi-GAN turn the
    new 45-55 latent frames into 32 kHz audio, spliced to the previous chunk with SOLA (the reference's infer_stream loop,
    TTS.py:402-498).
One "step" = one such utterance, driven through the PUBLIC entry ``gsv_tts.TTS.infer_phones_stream`` (models loaded
with ``TTS.load_gpt_model`` / ``load_sovits_model`` from checkpoint files in the reference's formats).
    value  = generated AR tokens / device time of the step, inputs resident in HBM;
    e2e    = the same call with HOST (pinned) inputs: host->device copies of the prompt, BERT features and latents and
             the device->host read of tokens and audio are inside the timed region.
The rest of BASELINE.json's metric rides in the same JSON line: ``batches`` (batch 1 / 8 / 32: decode-only roofline
and GPT + vocoder end to end), ``config3`` (128 mixed requests through 32 slots), ``config4`` (32 utterances per GPU,
dealt by length over the ranks, continuous batch + vocoder, audio-seconds per second of the whole job) and ``config5``
(vocoder only, 10 s x batch 64, V2Pro and V2ProPlus, against the tensor peak).

Arms:
    (default)          the CUDA path through libgsv_b200.so
    --impl reference   the reference algorithm's CPU path (oracle port; /root/reference cannot travel to
                       the GPU box and the reference forces fp32 on CPU, gsv_tts/TTS.py:74-76)
The reference's own CUDA path (FlashAttention decoder, CUDA graphs) on the same B200 is timed by
tools/ref_gpu_bench.py; its recorded numbers (profiles/r02_ref_gpu_baseline.json) are attached as ``ref_gpu``.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "gsv-tts-lite_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

NX, NY, N_TOK, CHUNK = 64, 100, 200, 25
FRAMES_FIRST, FRAMES_NEXT = 50, 55        # sovits_cache=[50,55] (reference README_EN.md:214)
WORKLOAD = "V2Pro batch=1 streaming: prefill 64+100, 200 tokens in 8 chunks of 25; per chunk prior encoder over the prefix + flow + HiFi-GAN on the new 45-55 frames + SOLA"
METRIC = "AR tokens/sec (V2Pro, batch=1 streaming, end-to-end incl. prefill + prior encoder + vocoder)"
REF_SAMPLE_TOKENS = 100                    # the CPU reference arm's bounded sample per step
W_BYTES = (24 * (12 * 512 * 512 + 13 * 512) + 1025 * 512) * 2      # 152.4 MB of 16-bit weights per step
KV_BYTES = 49152                                                   # per live position per sequence
DEC_FLOP = {"v2Pro": 813.1e6 + 14.2e6, "v2ProPlus": 1828.4e6 + 14.2e6}   # flow + dec FLOPs per 50 Hz frame


N_TEXT = 32                                 # target-text phonemes (the other NX - N_TEXT belong to the prompt)


def synth_inputs(seed):
    g = torch.Generator().manual_seed(seed)
    phones1 = torch.randint(0, 732, (NX - N_TEXT,), generator=g).tolist()
    phones2 = torch.randint(0, 732, (N_TEXT,), generator=g).tolist()
    y = torch.randint(0, 1024, (1, NY), generator=g)
    bert1 = torch.zeros(NX - N_TEXT, 1024)                 # JA/EN text: bert features are zeros (TextProcessor.py:100)
    bert2 = torch.zeros(N_TEXT, 1024)
    ge = torch.randn(1, 1024, 1, generator=g)
    return phones1, phones2, y, bert1, bert2, ge


def mixed_requests(n, seed, dev, dtype):
    """SURVEY.md 8d config 3 / 4 generator: Nx ~ U{40..120}, Ny ~ U{75..250}, target length ~ U{50..250}."""
    g = torch.Generator().manual_seed(seed)
    xs, ys, bs, mx = [], [], [], []
    for _ in range(n):
        nx = int(torch.randint(40, 121, (1,), generator=g))
        ny = int(torch.randint(75, 251, (1,), generator=g))
        xs.append(torch.randint(0, 732, (nx,), generator=g).to(dev))
        ys.append(torch.randint(0, 1024, (ny,), generator=g).to(dev))
        bs.append(torch.zeros(nx, 1024, device=dev, dtype=dtype))
        mx.append(int(torch.randint(50, 251, (1,), generator=g)))
    return xs, ys, bs, mx


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML DURING the timed region: one sample per utterance, taken by the
    consumer of the stream right after it received the middle clip -- the decode of the next chunk is running on the GPU at
    that moment and the host has nothing to launch.  (A sampling THREAD was measured to cost 60-100 ms outliers per 44 ms step:
    an NVML query that lands while the host is issuing the ~200 launches of a prefill stalls them.)"""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self.on, self.cost_ms = [], set(), None, False, []
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
        except Exception as e:            # pragma: no cover
            self.nv = None
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def sample(self):
        if not self.on or self.nv is None:
            return
        t0 = time.perf_counter()
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, nm in self.names.items():
                if r & bit:
                    self.reasons.add(nm)
        except Exception as e:            # pragma: no cover
            self.reasons.add(f"nvml_error:{type(e).__name__}")
        self.cost_ms.append(round((time.perf_counter() - t0) * 1e3, 2))

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_step(n_tok=N_TOK, threads=None):
    """One bounded sample of the workload on the host cores with the oracle port of the reference algorithm (fp32, all
    cores): GPT prefill + n_tok decode steps, then per 25-token chunk the prior encoder over the prefix (streaming cross-fade)
    and flow + HiFi-GAN over the chunk's new frames.  Returns (seconds, tokens, description)."""
    from gsv_tts import _synthetic as syn
    from oracle.encp_oracle import EncPOracle
    from oracle.gpt_oracle import GptOracle
    from oracle.vocoder_oracle import VocoderOracle
    torch.set_num_threads(threads or os.cpu_count())
    torch.set_grad_enabled(False)
    orc = GptOracle(syn.gpt_state_dict(syn.GPT_CONFIG, 0), syn.GPT_CONFIG)
    model = syn.SOVITS_MODEL["v2Pro"]
    vo = VocoderOracle(syn.sovits_flow_dec_state_dict(model, 0), model)
    enc = EncPOracle(syn.sovits_encp_state_dict(model, 0), model)
    phones1, phones2, y, bert1, bert2, ge = synth_inputs(1234)
    x = torch.tensor(phones1 + phones2)
    torch.manual_seed(0)
    t0 = time.perf_counter()
    toks = orc.infer(x, y[0], torch.cat([bert1, bert2]), max_seq=512, force_steps=n_tok)[0]          # [1, n]
    n_chunks = max(1, n_tok // CHUNK)
    vs = 0
    for c in range(n_chunks):
        codes = toks[:, : CHUNK * (c + 1)].unsqueeze(0)
        z_p, mask, _m, _l, ge2 = enc.decode_front(codes, torch.tensor(phones2).unsqueeze(0), ge, torch.randn(1, 192, 2 * codes.shape[-1] - vs),
                                                  stream_mode=True, valid_start_idx=vs, overlap_len=5)
        vo.flow_dec(z_p, mask, ge2)
        vs = 2 * codes.shape[-1] - 5
    dt = time.perf_counter() - t0
    return dt, n_tok, (f"1 utterance: prefill {NX}+{NY}, {n_tok} tokens, {n_chunks} chunks of prior encoder + vocoder, fp32, "
                       f"{torch.get_num_threads()} threads")


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    n_tok = REF_SAMPLE_TOKENS      # bounded sample: half an utterance per step keeps K steps within minutes
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
        "config": {"workload": WORKLOAD, "sample": desc,
                   "note": f"bounded sample of {n_tok} tokens / {n_tok // CHUNK} vocoder chunks per step (per-token rate; the shorter "
                           "KV makes the CPU side if anything faster per token than the full 200-token step of the CUDA arm)"},
        "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------------------------------
def write_checkpoints(tmp, sovits_keys=("v2Pro",)):
    """Synthetic checkpoints in the reference's on-disk formats (GPT: {"config","weight"} .ckpt; SoVITS: {"config","weight"}
    .pth), so that the bench loads its models the way a user does (TTS.load_gpt_model / load_sovits_model)."""
    from gsv_tts import _synthetic as syn
    gsd = syn.gpt_state_dict(syn.GPT_CONFIG, 0)
    gsd["ar_predict_layer.weight"][syn.GPT_CONFIG["model"]["EOS"]] = 0.0      # EOS never in the top-k (config 3/4 stop by max_new)
    gpath = os.path.join(tmp, "s1_synthetic.ckpt")
    torch.save({"config": syn.GPT_CONFIG, "weight": gsd}, gpath)
    spaths = {}
    for key in sovits_keys:
        model = dict(syn.SOVITS_MODEL[key])
        hps = {"data": {"filter_length": 2048, "hop_length": 640, "n_speakers": 300}, "train": {"segment_size": 20480}, "model": model}
        sp = os.path.join(tmp, f"s2_{key}_synthetic.pth")
        sd = dict(syn.sovits_flow_dec_state_dict(model, 0))
        sd.update(syn.sovits_encp_state_dict(model, 0))              # prior encoder, quantizer codebook, ge_to512
        torch.save({"config": hps, "weight": sd}, sp)
        spaths[key] = sp
    return gpath, spaths


class DecodeTimer:
    """CUDA events around every decode launch of a Text2SemanticDecoder (on the launching stream)."""

    def __init__(self, gpt):
        self.gpt, self.orig, self.ev, self.on = gpt, gpt._decode, [], False
        gpt._decode = self._decode

    def _decode(self, n):
        if not self.on:
            return self.orig(n)
        st = torch.cuda.current_stream(self.gpt._device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        self.orig(n)
        e1.record(st)
        self.ev.append((e0, e1, n))

    def take(self):
        out = [(a.elapsed_time(b), n) for a, b, n in self.ev]
        self.ev = []
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline only: skip batches / config3 / config4 / config5")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch.distributed as dist
    from gsv_tts import TTS
    from gsv_tts import _native as N
    from gsv_tts import _shard

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback on the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float16
    lib = N.lib()

    # ---- models through the public loaders; under torch.distributed only rank 0 reads the files and NCCL broadcasts the
    #      tensors GPU to GPU (Loader.get_*_weights)
    tmp = tempfile.mkdtemp(prefix="gsv_bench_")
    if rank == 0:
        gpath, spaths = write_checkpoints(tmp, ("v2Pro", "v2ProPlus") if not args.no_extra else ("v2Pro",))
    else:
        gpath, spaths = "s1_synthetic.ckpt", {"v2Pro": "s2_v2Pro_synthetic.pth", "v2ProPlus": "s2_v2ProPlus_synthetic.pth"}
    tts = TTS(gpt_cache=[(1, 512), (8, 512), (32, 512), (32, 1024)], sovits_cache=[FRAMES_FIRST, FRAMES_NEXT], device=dev, dtype=dtype)
    tts.load_gpt_model(gpath)
    tts.load_sovits_model(spaths["v2Pro"])
    gpt = tts.gpt_models[gpath].t2s_model
    voc = tts.sovits_models[spaths["v2Pro"]].vq_model
    timer = DecodeTimer(gpt)

    phones1, phones2, y, bert1, bert2, ge = synth_inputs(1234 + rank)
    dev_in = dict(y=y.to(dev), bert1=bert1.to(dev, dtype), bert2=bert2.to(dev, dtype), ge=ge.to(dev, dtype))
    host_in = dict(y=y.pin_memory(), bert1=bert1.to(dtype).pin_memory(), bert2=bert2.to(dtype).pin_memory(), ge=ge.to(dtype).pin_memory())
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)       # > 126 MB L2
    gpt.debug_seed = 99
    stream = torch.cuda.current_stream(dev)

    DEBUG_CLIPS = None
    sampler = ClockSampler(local)
    if os.environ.get("BENCH_DEBUG"):
        # host time spent inside the calls of a streaming utterance (where do first-clip outliers come from?)
        HOST_TRACE = []

        def traced(obj, name):
            fn = getattr(obj, name)

            def wrapper(*a, **k):
                t0 = time.perf_counter()
                try:
                    return fn(*a, **k)
                finally:
                    HOST_TRACE.append((name, round((time.perf_counter() - t0) * 1e3, 2), round(t0 * 1e3, 1)))
            setattr(obj, name, wrapper)
        for nm in ("_single_setup", "_decode", "_read"):
            traced(gpt, nm)
        traced(voc, "decode")

    def utterance(inp, first_clip_time=None):
        """One streaming utterance through the public API: 8 chunks of 25 tokens, one AudioClip each."""
        n_clips = 0
        samples = 0
        t_prev = time.perf_counter()
        for clip in tts.infer_phones_stream(phones1, inp["bert1"], inp["y"], phones2, inp["bert2"], inp["ge"], stream_chunk=CHUNK,
                                            overlap_len=5, force_steps=N_TOK):
            if n_clips == 0 and first_clip_time is not None:
                first_clip_time.append(time.perf_counter())
            n_clips += 1
            samples += clip.audio_data.shape[0]
            if n_clips == 4:
                sampler.sample()
            if DEBUG_CLIPS is not None:
                t_now = time.perf_counter()
                DEBUG_CLIPS.append(round((t_now - t_prev) * 1e3, 1))
                t_prev = t_now
        assert n_clips == N_TOK // CHUNK, n_clips
        return N_TOK, samples

    def launch_total():
        n = int(lib.gsv_gpt_launch_count(gpt._ctx)) + voc.launch_count()
        if voc._enc_ctx is not None:
            n += voc.enc_launch_count()
        return n

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(inp, steps, collect):
        nonlocal DEBUG_CLIPS
        evs = []
        DEBUG_CLIPS = [] if os.environ.get("BENCH_DEBUG") else None
        if DEBUG_CLIPS is not None:
            del HOST_TRACE[:]
        l0 = launch_total()
        timer.on = collect or bool(os.environ.get("BENCH_DEBUG"))
        barrier()
        gc.collect()
        gc.disable()                                         # a generation-2 collection inside a 44 ms step is a 30-90 ms outlier
        utterance(inp)                                       # one more untimed step in exactly the timed configuration (timer events, GC off)
        timer.take()
        if DEBUG_CLIPS is not None:
            del DEBUG_CLIPS[:]
            del HOST_TRACE[:]
        l0 = launch_total()
        barrier()
        wall0 = time.perf_counter()
        for _ in range(steps):
            flush.zero_()                                    # evict L2 between timed steps (outside the events)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            utterance(inp)
            b.record(stream)
            evs.append((a, b))
        barrier()
        wall = time.perf_counter() - wall0
        gc.enable()
        timer.on = False
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if os.environ.get("BENCH_DEBUG"):
            sys.stderr.write(f"[bench] collect={collect} per-step ms: {[round(a.elapsed_time(b), 1) for a, b in evs]}\n")
            sys.stderr.write(f"[bench] host ms between clips: {DEBUG_CLIPS}\n")
            slow = [i for i, (a, b) in enumerate(evs) if a.elapsed_time(b) > 50]
            per = len(HOST_TRACE) // max(1, steps)
            for i in slow[:3]:
                sys.stderr.write(f"[bench] slow step {i}: {HOST_TRACE[i * per:(i + 1) * per][:12]}\n")
            if slow:
                sys.stderr.write(f"[bench] normal step: {HOST_TRACE[0:per][:12] if 0 not in slow else HOST_TRACE[per:2 * per][:12]}\n")
            del HOST_TRACE[:]
            if not collect:
                sys.stderr.write(f"[bench] decode launch ms (host-input pass): {[round(m, 1) for m, _ in timer.take()]}\n")
        DEBUG_CLIPS = None
        launches = launch_total() - l0
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall, launches

    for _ in range(args.warmup):                             # both input kinds: host inputs allocate their own device copies
        flush.zero_()                                        # (first touch of the flush buffer belongs to the warm-up too)
        utterance(dev_in)
        utterance(host_in)
    sampler.on = not os.environ.get("BENCH_NO_SAMPLER")
    ms_total, wall, launches = timed(dev_in, args.steps, True)
    decode_launches = timer.take()
    ms_e2e, wall_e2e, _ = timed(host_in, args.steps, False)
    sampler.on = False
    if os.environ.get("BENCH_DEBUG"):
        sys.stderr.write(f"[bench] NVML sample cost ms: {sampler.cost_ms}\n")
    # TTFT (BASELINE config 2): host call -> first AudioClip in host memory, median of 5
    ttfts = []
    for _ in range(5):
        flush.zero_()
        torch.cuda.synchronize(dev)
        marks = []
        t_call = time.perf_counter()
        utterance(host_in, marks)
        ttfts.append((marks[0] - t_call) * 1e3)
    ttfts.sort()
    ttft_ms = ttfts[len(ttfts) // 2]

    tokens = args.steps * N_TOK * world
    value = tokens / (ms_total / 1e3)
    e2e_value = tokens / (ms_e2e / 1e3)
    audio_s = N_TOK * 0.04
    # ---- roofline of the dominant kernel: the single-sequence persistent decode kernel (HBM bound) ----
    pk, pk_kind = peaks()
    kv_mean = NX + NY + (N_TOK + 1) / 2.0
    bytes_per_token = W_BYTES + KV_BYTES * kv_mean              # weights once + 49 152 B per live position
    mean_launch_ms = sum(ms for ms, _ in decode_launches) / len(decode_launches)
    achieved = bytes_per_token * CHUNK / (mean_launch_ms / 1e3) / 1e9
    traffic = None      # dram bytes per launch of the same kernel from the committed ncu --set full capture
    for name in ("r02_decode_traffic.json", "r01_decode_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                tj = json.load(f)
            traffic = (tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]) * CHUNK / tj["tokens_per_launch"]
            break
        except Exception:
            pass
    roofline = {"kernel": "gpt_decode_hx_kernel", "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk_kind,
                "bytes_per_launch": bytes_per_token * CHUNK, "launch_ms": mean_launch_ms,
                "decode_only_tok_s": CHUNK / (mean_launch_ms / 1e3), "us_per_token": mean_launch_ms * 1e3 / CHUNK}

    blocks = {}
    if not args.no_extra:
        try:
            blocks = metric_blocks(tts, gpt, voc, timer, spaths, dev, dtype, pk, rank, world, dist, _shard, barrier)
        except Exception as e:      # the extra blocks must never break the headline line
            blocks = {"error": f"{type(e).__name__}: {e}"}

    cpu_base = None
    if rank == 0 and not args.no_cpu_baseline:
        dt, ntok, desc = cpu_reference_step(50)
        cpu_base = {"value": ntok / dt, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port", "sample": desc}

    if rank == 0:
        n_chunks = N_TOK // CHUNK
        h2d = sum(t.numel() * t.element_size() for t in host_in.values()) + (NX + N_TEXT) * 8
        d2h = (2 * N_TOK + 5 * (n_chunks - 1)) * 640 * 4 + n_chunks * (512 * 4 + 8) + n_chunks * 2 * 4   # audio buffers (fp32, whole, cut on the host), tokens + state per chunk, glue offsets
        ref_gpu = None
        try:
            with open(os.path.join(ROOT, "profiles", "r02_ref_gpu_baseline.json")) as f:
                rg = json.load(f)
            ref_gpu = {"source": "profiles/r02_ref_gpu_baseline.json (tools/ref_gpu_bench.py on this pool's B200, recorded, not re-run here)",
                       "flash_b1_tok_s": rg["flash_b1_infer"]["tok_s"], "flash_b8_tok_s": rg["flash_b8_infer_batched"]["tok_s"],
                       "flash_b32_tok_s": rg["flash_b32_infer_batched"]["tok_s"], "flash_first_chunk_ms": rg["flash_b1_stream"]["first_chunk_ms"],
                       "voc_v2Pro_B1_T50_graph_ms": rg["voc_v2Pro_B1_T50"]["graph_ms"], "voc_v2Pro_B64_T500_ms": rg["voc_v2Pro_B64_T500"]["eager_ms"]}
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "256 MB flush between timed steps; 152 MB of weights (> L2) streamed per token",
                       "api": "gsv_tts.TTS.infer_phones_stream (models from TTS.load_gpt_model / load_sovits_model)",
                       "parallelism": f"{world} independent utterance streams, rank 0 reads the checkpoints, NCCL broadcast GPU to GPU at load only",
                       "overlap": "prior encoder + vocoder of chunk c on a second stream while chunk c+1 decodes on 64 SMs; held-back chunks decoded ahead of their yield (same clips, same order)",
                       "reference_arm_sample": f"{REF_SAMPLE_TOKENS} tokens / {REF_SAMPLE_TOKENS // CHUNK} vocoder chunks per step (per-token rate)"},
            "rtf": (ms_total / 1e3 / args.steps) / audio_s, "ttft_ms": ttft_ms,
            "roofline": roofline, "cpu_baseline": cpu_base,
            "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "rtf": (ms_e2e / 1e3 / args.steps) / audio_s},
            "gpu_launches": launches, "clocks": sampler.summary(), "wall_s": wall + wall_e2e, "ref_gpu": ref_gpu,
        }
        line.update(blocks)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def metric_blocks(tts, gpt, voc, timer, spaths, dev, dtype, pk, rank, world, dist, _shard, barrier):
    """The rest of BASELINE.json's metric.  Every block times device work with CUDA events (GPT decode launches) or the
    wall clock around a public call bracketed by synchronisations, after one warm-up call."""
    out = {}
    g = torch.Generator().manual_seed(7)

    def sync_time(fn):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize(dev)
        return time.perf_counter() - t0, r

    # ---- batches 1 / 8 / 32: B identical-length requests through B slots, 200 tokens each (EOS never sampled), then the
    #      vocoder over the B utterances (400 frames each); decode launches timed for the roofline of the batch kernel
    batches = {}
    for B in (1, 8, 32):
        xs = [torch.randint(0, 732, (NX,), generator=g).to(dev) for _ in range(B)]
        ys = [torch.randint(0, 1024, (NY,), generator=g).to(dev) for _ in range(B)]
        bs = [torch.zeros(NX, 1024, device=dev, dtype=dtype) for _ in range(B)]
        ph2 = [torch.randint(0, 732, (N_TEXT,), generator=g).to(dev) for _ in range(B)]
        ges = [torch.randn(1, 1024, 1, generator=g).to(dev, dtype) for _ in range(B)]
        gpt.debug_seed = 11
        warm = tts.infer_features_batched(xs, bs, ys, max_new=[16] * B)      # warm-up (kernel choice, weight re-tiling)
        tts.decode_batched([torch.zeros(N_TOK, dtype=torch.int64, device=dev)] * B, ph2, ges)   # same shapes as the timed call
        timer.on = True
        gpt.debug_seed = 11
        t_gpt, toks = sync_time(lambda: tts.infer_features_batched(xs, bs, ys, max_new=[N_TOK] * B))
        launches = timer.take()
        timer.on = False
        t_voc, clips = sync_time(lambda: tts.decode_batched(toks, ph2, ges))
        n_tok = sum(int(t.numel()) for t in toks)
        steps = min(sum(n for _, n in launches), N_TOK + 1)      # the last launch stops early once every sequence is done
        ms_dec = sum(ms for ms, _ in launches)
        us_step = ms_dec * 1e3 / max(steps, 1)
        kv_mean = NX + NY + (N_TOK + 1) / 2.0
        bytes_step = W_BYTES + B * KV_BYTES * kv_mean
        ach = bytes_step / (us_step * 1e-6) / 1e9
        audio_s = n_tok * 0.04
        batches[str(B)] = {
            "decode_us_per_step": us_step, "decode_tok_s": B * 1e6 / us_step,
            "roofline": {"kernel": "gpt_decode_hx_kernel" if B == 1 else "gpt_decode_cl8_kernel", "bound": "hbm", "achieved": ach,
                         "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "bytes_per_step": bytes_step},
            "gpt_stage_tok_s": n_tok / t_gpt, "gpt_stage_ms": t_gpt * 1e3, "sovits_stage_ms": t_voc * 1e3,
            "e2e_tok_s": n_tok / (t_gpt + t_voc), "e2e_rtf": (t_gpt + t_voc) / audio_s, "audio_s_per_s": audio_s / (t_gpt + t_voc),
            "tokens": n_tok,
        }
    out["batches"] = batches

    # ---- config 3 (GPT stage): 128 mixed requests through 32 slots of infer_batched, refill prompts on a second stream
    xs, ys, bs, mx = mixed_requests(128, 1234, dev, dtype)
    gpt.debug_seed = 5
    tts.infer_features_batched(xs[:40], bs[:40], ys[:40], max_new=[20] * 40)
    gpt.debug_seed = 5
    t3, toks = sync_time(lambda: tts.infer_features_batched(xs, bs, ys, max_new=mx))
    n3 = sum(int(t.numel()) for t in toks)
    out["config3"] = {"workload": "128 mixed requests (Nx U{40..120}, Ny U{75..250}, length U{50..250}) through 32 slots, V2Pro-size GPT",
                      "tokens": n3, "ms": t3 * 1e3, "tok_s": n3 / t3, "audio_s_per_s": n3 * 0.04 / t3}

    # the same 128 requests handed to the slot scheduler longest predicted first (TTS.infer_features_batched(queue_order=...)):
    # the batch drains with short requests instead of a few long ones (ideal schedule 609 decode steps instead of 732).  The
    # predictor is max_new -- here the known target length, i.e. a perfect one; a text-length predictor would do less well.
    try:
        gpt.debug_seed = 5
        t3l, toksl = sync_time(lambda: tts.infer_features_batched(xs, bs, ys, max_new=mx, queue_order="longest_first"))
        n3l = sum(int(t.numel()) for t in toksl)
        out["config3_longest_first"] = {"workload": "config 3's 128 requests, queue ordered longest predicted first (predictor: the known target length)",
                                        "tokens": n3l, "ms": t3l * 1e3, "tok_s": n3l / t3l, "audio_s_per_s": n3l * 0.04 / t3l}
    except Exception as e:      # a side measurement: never in the way of the blocks below
        out["config3_longest_first"] = {"error": f"{type(e).__name__}: {e}"}

    # ---- config 4: 32 utterances per GPU, dealt over the ranks by predicted length, continuous batch + vocoder per rank
    n4 = 32 * world
    xs, ys, bs, mx = mixed_requests(n4, 4321, dev, dtype)
    mine = _shard.shard_by_length([len(a) + m for a, m in zip(xs, mx)], world)[rank]
    gz = torch.Generator().manual_seed(99 + rank)
    lx, ly, lb, lm = [xs[i] for i in mine], [ys[i] for i in mine], [bs[i] for i in mine], [mx[i] for i in mine]
    ges = [torch.randn(1, 1024, 1, generator=gz).to(dev, dtype) for _ in mine]
    ph2s = [torch.randint(0, 732, (max(8, len(a) // 2),), generator=gz).to(dev) for a in lx]

    def run4(overlap=True):
        # SoVITS stage (prior encoder per utterance, flow + HiFi-GAN in padded groups) of the requests harvested at one read
        # on the second stream while the other slots keep decoding
        _toks, clips = tts.infer_phones_batched(lx, lb, ly, ph2s, ges, max_new=lm, overlap=overlap)
        return sum(c.audio_data.shape[0] for c in clips) / 32000.0

    gpt.debug_seed = 6
    run4()
    gpt.debug_seed = 6
    t4_serial, _ = sync_time(lambda: run4(False))                                  # the reference's order: GPT stage, then SoVITS stage
    barrier()
    gpt.debug_seed = 6
    t4, audio4 = sync_time(run4)
    stat = torch.tensor([t4, audio4], device=dev, dtype=torch.float64)
    if world > 1:
        tmax = stat[:1].clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        asum = stat[1:].clone()
        dist.all_reduce(asum, op=dist.ReduceOp.SUM)
        t4, audio4 = float(tmax.item()), float(asum.item())
        got = _shard.gather_in_order({i: 1 for i in mine}, n4)                   # every request answered exactly once
        assert len(got) == n4
    out["config4"] = {"workload": f"{n4} utterances (32 per GPU) sharded by length over {world} rank(s); continuous batch + prior encoder + flow/HiFi-GAN per rank",
                      "audio_s": audio4, "ms": t4 * 1e3, "audio_s_per_s": audio4 / t4, "rtf": t4 / audio4,
                      "ms_back_to_back_rank0": t4_serial * 1e3}

    # ---- config 5: vocoder only, 10 s of latents x batch 64 (rank 0)
    if rank == 0:
        c5 = {}
        for key in ("v2Pro", "v2ProPlus"):
            try:
                if key == "v2Pro":
                    v = voc
                else:
                    tts.load_sovits_model(spaths[key]) if world == 1 else None
                    v = tts.sovits_models[spaths[key]].vq_model if spaths[key] in tts.sovits_models else None
                if v is None:
                    continue
                gin = v.gin_channels
                z = torch.randn(64, 192, 500, device=dev, dtype=dtype)
                mk = torch.ones(64, 1, 500, device=dev, dtype=dtype)
                gg = torch.randn(64, gin, 1, device=dev, dtype=dtype)
                v.flow_dec(z, mk, gg)
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                v.flow_dec(z, mk, gg)
                e1.record()
                torch.cuda.synchronize(dev)
                ms = e0.elapsed_time(e1)
                tf = 64 * 500 * DEC_FLOP[key] / (ms / 1e3) / 1e12
                c5[key] = {"ms": ms, "tflops": tf, "frac_of_tensor_peak": tf / pk["bf16_tflops"], "audio_s_per_s": 64 * 10.0 / (ms / 1e3),
                           "roofline": {"kernel": "conv_umma_ws_kernel + conv_umma_kernel (flow + HiFi-GAN call)", "bound": "tensor", "achieved": tf,
                                        "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": tf / pk["bf16_tflops"]}}
                del z, mk, gg
            except Exception as e:
                c5[key] = {"error": f"{type(e).__name__}: {e}"}
        out["config5"] = c5
    return out


if __name__ == "__main__":
    main()
