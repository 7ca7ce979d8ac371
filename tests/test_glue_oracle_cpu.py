"""CPU: the oracle of the host glue (monotonic alignment, silence offsets, SOLA, subtitle timing)
against outputs of the REFERENCE's own method bodies (tests/golden/glue.npz, written by oracle/make_golden.py glue) and,
when /root/reference exists, against those bodies run live on other inputs."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import glue_oracle as G      # noqa: E402
from oracle import ref_shim              # noqa: E402


def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "glue.npz"))


def test_glue_oracle_matches_reference_goldens():
    g = golden()
    i = 0
    while f"attn{i}" in g:
        assert (G.viterbi_monotonic(g[f"attn{i}"]) == g[f"assign{i}"]).all(), i
        i += 1
    assert i == 4
    assert G.head_offset(g["audio"]) == int(g["head_offset"]) and G.tail_offset(g["audio"]) == int(g["tail_offset"])
    assert G.head_offset(np.zeros(5000, np.float32)) == int(g["head_offset_silent"])
    assert G.tail_offset(np.zeros(5000, np.float32)) == int(g["tail_offset_silent"])
    out, off = G.sola(g["sola_f1"], g["sola_f2"], 3200)
    assert off == int(g["sola_offset"]) and np.abs(out - g["sola_out"]).max() < 1e-5


def test_alignment_is_monotonic_and_covers_the_text():
    g = golden()
    a = G.viterbi_monotonic(g["attn3"])
    body = a[a >= 0]
    assert (np.diff(body) >= 0).all() and (np.diff(body) <= 1).all()


@pytest.mark.skipif(not os.path.exists(os.path.join(ref_shim.REF_ROOT, "gsv_tts", "TTS.py")), reason="reference TTS.py not present")
def test_glue_oracle_matches_reference_bodies_live():
    sys.path.insert(0, os.path.join(ROOT, "gsv-tts-lite_b200"))
    from oracle.make_golden import glue_cases, reference_glue_methods
    ns, me = reference_glue_methods()
    attn, audio, f1, f2 = glue_cases(seed=7)
    for a in attn:
        assert (ns["_viterbi_monotonic"](me, a).numpy() == G.viterbi_monotonic(a.numpy())).all()
    assert ns["_find_head_threshold_offsets"](me, audio) == G.head_offset(audio.numpy())
    assert ns["_find_tail_threshold_offsets"](me, audio) == G.tail_offset(audio.numpy())
    w2p = {"word": ["he", "llo", " ", "wor", "ld"], "ph": [2, 3, 1, 3, 2]}
    assign = torch.tensor([-1, -1, 0, 0, 1, 1, 2, 3, 3, 4, 5, 5, 6, 7, 8, 8, 9, 10, 10, 10])
    for speed, last in ((1.0, 0.0), (1.3, 2.5)):
        assert ns["_get_subtitles"](me, w2p, assign, speed, last) == G.get_subtitles(w2p, assign.numpy(), speed, 50, last)
