"""Where does the SoVITS stage of ONE streaming chunk spend its time?  (TTS.infer_phones_stream: prior encoder over the whole
prefix, flow + HiFi-GAN on the new frames, SOLA, copy to the host.)  Alone on the GPU, V2Pro-size synthetic weights, per
prefix length: device time of each part (CUDA events), host time to enqueue it, and the launch counts.

    gpurun -- python tools/sovits_chunk_latency.py > gpurun_out/sovits_chunk_latency.txt
"""
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gsv-tts-lite_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import bench
    from gsv_tts import TTS
    dev = torch.device("cuda:0")
    tmp = tempfile.mkdtemp()
    gpath, spaths = bench.write_checkpoints(tmp)
    tts = TTS(gpt_cache=[(1, 512)], sovits_cache=[50, 55], device=dev, dtype=torch.bfloat16)
    tts.load_sovits_model(spaths["v2Pro"])
    vq = tts.sovits_models[spaths["v2Pro"]].vq_model
    g = torch.Generator().manual_seed(3)
    ph2 = torch.randint(0, 732, (1, 32), generator=g).to(dev)
    ge = torch.randn(1, 1024, 1, generator=g).to(dev, torch.bfloat16)
    side = torch.cuda.Stream(dev)
    ov = 5 * vq.samples_per_frame
    lib_enc = lambda: int(__import__("gsv_tts._native", fromlist=["lib"]).lib().gsv_encp_launch_count(vq._enc_ctx))

    def timed(fn, reps=20):
        """(device ms, host ms) per call, median"""
        dts, hts = [], []
        for _ in range(reps):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record(side)
            fn()
            b.record(side)
            hts.append((time.perf_counter() - t0) * 1e3)
            torch.cuda.synchronize()
            dts.append(a.elapsed_time(b))
        dts.sort(); hts.sort()
        return dts[len(dts) // 2], hts[len(hts) // 2]

    print("prefix_tokens  prior_ms(host)  flow_dec_ms(host)  sola+copy_ms(host)  whole_stage_ms(host)  encp_launches voc_launches")
    with torch.inference_mode(), torch.cuda.stream(side):
        for n in (25, 50, 100, 200, 400):
            codes = torch.randint(0, 1024, (1, 1, n), generator=g).to(dev)
            vs = 0 if n == 25 else 2 * (n - 25) - 5
            keep = {}

            def prior():
                vq.enc_p.y_overlap = None
                keep["p"] = vq.prior(codes, ph2, ge, 0.5, 1.0, True, vs, 5)

            def flow():
                z_p, y_mask, ge2, attn = keep["p"]
                keep["a"] = vq.flow_dec(z_p, y_mask, ge2)

            def tail():
                flat = keep["a"].reshape(-1)
                n2 = flat.numel()
                out, off = tts._sola_enqueue(flat[:ov].clone(), flat, ov)
                host = torch.empty(n2, dtype=torch.float32, pin_memory=True)
                host.copy_(out, non_blocking=True)

            def whole():
                prior(); flow(); tail()

            for _ in range(3):
                whole()
            l0, v0 = lib_enc(), vq.launch_count()
            whole()
            l1, v1 = lib_enc(), vq.launch_count()
            r = [timed(prior), timed(flow), timed(tail), timed(whole)]
            print(f"{n:6d}   " + "   ".join(f"{d:6.3f} ({h:5.3f})" for d, h in r) + f"   {l1 - l0:4d} {v1 - v0:4d}")


if __name__ == "__main__":
    main()
