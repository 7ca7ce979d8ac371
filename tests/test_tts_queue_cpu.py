"""CPU: ``TTS.infer_features_batched(queue_order=...)`` -- the permutation in front of the slot scheduler and the way back
(the scheduler itself is covered by test_batched_host_cpu.py; here it is a stub that records what it was given)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gsv-tts-lite_b200"))


class _StubGpt:
    def __init__(self):
        self.seen = None

    def infer_batched(self, x, y, bert, top_k=15, top_p=1.0, temperature=1.0, max_new=None):
        self.seen = dict(ids=[int(t[0]) for t in x], prompts=[int(t[0]) for t in y], berts=[int(b[0, 0]) for b in bert], max_new=max_new)
        n = len(x)
        finish = list(reversed(range(n)))                          # completion order: last queued first
        outs = [torch.full((3,), int(x[q][0]), dtype=torch.int64) for q in finish]
        return outs, torch.tensor(finish)


def _tts(stub):
    from gsv_tts import TTS
    t = TTS.__new__(TTS)
    t.gpt_models = {"m": type("W", (), {"t2s_model": stub})()}
    return t


def _requests(lengths):
    ids = [torch.full((n,), r, dtype=torch.int64) for r, n in enumerate(lengths)]       # element 0 names the request
    prompts = [torch.full((4,), r, dtype=torch.int64) for r in range(len(lengths))]
    berts = [torch.full((n, 2), float(r)) for r, n in enumerate(lengths)]
    return ids, berts, prompts


def test_longest_first_permutes_the_queue_and_restores_request_order():
    stub = _StubGpt()
    tts = _tts(stub)
    ids, berts, prompts = _requests([5, 9, 7, 9, 3])
    max_new = [50, 10, 0, 200, 50]                                 # 0 = no cap: ahead of every capped request
    res = tts.infer_features_batched(ids, berts, prompts, max_new=max_new, queue_order="longest_first")
    assert stub.seen["ids"] == [2, 3, 0, 4, 1]                     # uncapped, 200, 50 (request 0 before 4: stable), 10
    assert stub.seen["prompts"] == stub.seen["ids"] == stub.seen["berts"]
    assert stub.seen["max_new"] == [0, 200, 50, 50, 10]
    assert [int(t[0]) for t in res] == [0, 1, 2, 3, 4] and all(t.numel() == 3 for t in res)
    # without caps: by phoneme count, ties in the caller's order
    res = tts.infer_features_batched(ids, berts, prompts, queue_order="longest_first")
    assert stub.seen["ids"] == [1, 3, 2, 0, 4] and stub.seen["max_new"] is None
    assert [int(t[0]) for t in res] == [0, 1, 2, 3, 4]


def test_default_order_is_the_callers():
    stub = _StubGpt()
    tts = _tts(stub)
    ids, berts, prompts = _requests([5, 9, 7])
    for order in (None, "fifo"):
        res = tts.infer_features_batched(ids, berts, prompts, max_new=[9, 1, 5], queue_order=order)
        assert stub.seen["ids"] == [0, 1, 2] and stub.seen["max_new"] == [9, 1, 5]
        assert [int(t[0]) for t in res] == [0, 1, 2]
    with pytest.raises(ValueError):
        tts.infer_features_batched(ids, berts, prompts, queue_order="shortest_first")
