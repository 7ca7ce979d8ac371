"""Make the reference's hot-path modules available on the GPU box.

The reference is pure Python; ``pip install --target baseline/_ref /root/reference`` is of no use because importing
the package needs ``av``, ``pysbd``, ... which are absent offline.  What the same-box GPU baseline
needs (``tools/ref_gpu_bench.py``: the reference's FlashAttention / SDPA decoders and its ``SynthesizerTrn``) is the
import closure of ``GPT/t2s_model*.py`` and ``SoVITS/models.py``: 18 files, copied UNMODIFIED into the git-ignored
``baseline/_ref/gsv_tts/`` (it travels with ``gpurun`` like a built ``.so``; it never enters the history).

    python baseline/install_ref.py          # build container only (needs /root/reference)
"""
import os
import shutil
import sys

REF = os.environ.get("GSV_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")

FILES = [
    "gsv_tts/Config.py",
    "gsv_tts/GPT_SoVITS/GPT/__init__.py",
    "gsv_tts/GPT_SoVITS/GPT/embedding.py",
    "gsv_tts/GPT_SoVITS/GPT/t2s_model.py",
    "gsv_tts/GPT_SoVITS/GPT/t2s_model_flash_attn.py",
    "gsv_tts/GPT_SoVITS/GPT/utils.py",
    "gsv_tts/GPT_SoVITS/SoVITS/models.py",
    "gsv_tts/GPT_SoVITS/SoVITS/module/__init__.py",
    "gsv_tts/GPT_SoVITS/SoVITS/module/attentions.py",
    "gsv_tts/GPT_SoVITS/SoVITS/module/commons.py",
    "gsv_tts/GPT_SoVITS/SoVITS/module/core_vq.py",
    "gsv_tts/GPT_SoVITS/SoVITS/module/modules.py",
    "gsv_tts/GPT_SoVITS/SoVITS/module/mrte_model.py",
    "gsv_tts/GPT_SoVITS/SoVITS/module/quantize.py",
    "gsv_tts/GPT_SoVITS/G2P/__init__.py",
    "gsv_tts/GPT_SoVITS/G2P/Symbols.py",
    "gsv_tts/GPT_SoVITS/G2P/Pause.py",
]


def main() -> int:
    if not os.path.isdir(os.path.join(REF, "gsv_tts")):
        print(f"reference tree not found at {REF}: nothing installed")
        return 0
    n = 0
    for rel in FILES:
        src = os.path.join(REF, rel)
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        n += 1
    print(f"copied {n} reference files into {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
