"""Teacher-forced logits error of the full-size model with the tensor-core and the CUDA-core prefill GEMMs."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    sys.path[:0] = [ROOT, os.path.join(ROOT, "gsv-tts-lite_b200")]
    import torch
    from gsv_tts import _synthetic as syn
    from tests import gpu_harness as H
    for dt in (torch.float16, torch.bfloat16):
        e = H.gpt_teacher_forced_error(syn.GPT_CONFIG, "full", dt, torch.device("cuda:0"))
        print(sys.argv[1], dt, "vs_oracle %.4e vs_golden %.4e row0 %.4e" % (e["vs_oracle"], e["vs_golden"], e["per_row_vs_oracle"][0]))
else:
    for mode in ("umma", "cuda"):
        env = dict(os.environ, GSV_GPT_GEMM=mode)
        subprocess.run([sys.executable, __file__, mode], env=env)
