"""ORACLE (test infrastructure, never imported by the product path): CPU restatement of the host glue the reference runs
around the two models in ``TTS.infer`` / ``infer_stream``: monotonic alignment of the MRTE
attention map (``_viterbi_monotonic``, reference gsv_tts/TTS.py:1744-1797), leading / trailing silence offsets
(``_find_head_threshold_offsets`` / ``_find_tail_threshold_offsets``, :1630-1662), the SOLA splice between streaming
chunks (``_sola_algorithm``, :1612-1628) and the frame -> word subtitle timing (``_get_subtitles``, :1664-1707).

numpy, one function per reference method; pinned by tests/test_glue_oracle_cpu.py to outputs of the reference's own
method bodies (oracle/make_golden.py glue compiles them from TTS.py's source text, since the module itself cannot be
imported here) and re-checked live when /root/reference exists.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np


def viterbi_monotonic(attn: np.ndarray) -> np.ndarray:
    """TTS.py:1744-1797.  attn [H, T, N] -> assign [T] int64 (text index per frame, -1 before the first frame whose averaged
    distribution peaks at index 0)."""
    attn = attn.astype(np.float32)
    H, T, N = attn.shape
    max_idx = attn.argmax(-1)
    mask = max_idx != N - 1                                   # heads that look at the null key are left out
    sum_attn = (attn * mask[..., None].astype(np.float32)).sum(0)
    count = mask.sum(0)[:, None]
    default = np.full((T, N), 1.0 / N, dtype=np.float32)
    default[:, N - 1] = 0.9 / N
    default[:, 1] = 1.1 / N
    default /= default.sum(-1, keepdims=True)
    normal = np.where(count > 0, sum_attn / (count.astype(np.float32) + np.float32(1e-9)), default).astype(np.float32)
    is_zero = normal.argmax(-1) == 0
    first_zero = int(np.nonzero(is_zero)[0][0]) if is_zero.any() else 0
    dp = np.zeros((T, N), dtype=np.float32)
    ptr = np.zeros((T, N), dtype=np.int64)
    dp[0] = normal[0]
    ar = np.arange(N)
    for t in range(1, T):
        prev = dp[t - 1]
        shifted = np.concatenate([[-np.inf], prev[:-1]]).astype(np.float32)
        rel = (shifted > prev).astype(np.int64)               # torch.max over the stacked pair keeps the first on ties
        dp[t] = normal[t] + np.maximum(prev, shifted)
        ptr[t] = ar - rel
    assign = np.zeros(T, dtype=np.int64)
    assign[-1] = int(dp[-1].argmax())
    for t in range(T - 2, -1, -1):
        assign[t] = ptr[t + 1, assign[t + 1]]
    assign[:first_zero] = -1
    return assign


def _frame_rms(x: np.ndarray, frame_length: int, hop_length: int) -> np.ndarray:
    n = (len(x) - frame_length) // hop_length + 1 if len(x) >= frame_length else 0
    if n <= 0:
        return np.zeros(0, dtype=np.float32)
    idx = np.arange(frame_length)[None, :] + hop_length * np.arange(n)[:, None]
    return np.sqrt((x[idx].astype(np.float32) ** 2).mean(1))


def head_offset(audio: np.ndarray, threshold=0.02, frame_length=512, hop_length=256, search_len=64000, margin=3200) -> int:
    """TTS.py:1630-1645."""
    head = audio[:search_len]
    hit = np.nonzero(_frame_rms(head, frame_length, hop_length) > threshold)[0]
    if hit.size:
        return max(0, int(hit[0]) * hop_length - margin)
    return int(head.shape[0])


def tail_offset(audio: np.ndarray, threshold=0.01, frame_length=512, hop_length=256, search_len=64000, margin=3200) -> int:
    """TTS.py:1647-1662."""
    tail = audio[-search_len:]
    hit = np.nonzero(_frame_rms(tail, frame_length, hop_length) > threshold)[0]
    if hit.size:
        return max(1, tail.shape[0] - int(hit[-1]) * hop_length - margin)
    return int(tail.shape[0])


def sola(f1_overlap: np.ndarray, f2: np.ndarray, overlap_len: int, search_len: int = 320):
    """TTS.py:1612-1628 on 1-D signals: -> (f2 spliced: cross-faded overlap + the rest after the best offset, offset)."""
    q = f1_overlap.astype(np.float32)
    key = f2[: overlap_len + search_len].astype(np.float32)
    n = len(key) - len(q) + 1
    idx = np.arange(len(q))[None, :] + np.arange(n)[:, None]
    corr = (key[idx] * q[None, :]).sum(1)
    energy = (key[idx] ** 2).sum(1) + np.float32(1e-8)
    offset = int((corr / np.sqrt(energy)).argmax())
    aligned = f2[offset:].astype(np.float32)
    alpha = np.linspace(0, 1, overlap_len, dtype=np.float32)
    faded = q * (1 - alpha) + aligned[:overlap_len] * alpha
    return np.concatenate([faded, aligned[overlap_len:]]), offset


def get_subtitles(word2ph: Dict[str, list], assign: np.ndarray, speed: float, sovits_hz: int = 50, last_end_s: float = 0.0) -> List[dict]:
    """TTS.py:1664-1707."""
    frame_time = (1 / sovits_hz) / speed
    ph_end_s = []
    cur = int(assign[0])
    for f in range(1, assign.shape[-1]):
        ph = int(assign[f])
        if ph != cur:
            ph_end_s.append(f * frame_time)
            cur = ph
    ph_end_s.append(assign.shape[-1] * frame_time)
    idx = -1
    end_s = last_end_s + ph_end_s.pop(0) if assign[0] == -1 else last_end_s
    subs = []
    word = None
    for i in range(len(word2ph["word"])):
        word, ph = word2ph["word"][i], word2ph["ph"][i]
        idx += ph
        if idx >= len(ph_end_s):
            break
        start_s = end_s
        end_s = ph_end_s[idx] + last_end_s
        subs.append({"text": word, "start_s": start_s, "end_s": end_s})
    if end_s - last_end_s != ph_end_s[-1]:
        start_s = end_s
        end_s = ph_end_s[-1] + last_end_s
        subs.append({"text": word, "start_s": start_s, "end_s": end_s})
    return subs
