"""CPU: the scheduling side of ``Text2SemanticDecoder.infer_batched`` (t2s_model.py:555-734: slots, harvest, refill from
the queue, completion order) with the native library replaced by a scripted runtime that follows the kernels'
bookkeeping, in both refill modes: reference order (prefill between two decode launches) and overlapped
(``gsv_gpt_prefill_begin`` on a second stream, ``gsv_gpt_prefill_finish`` behind the next decode launch)."""
import contextlib
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gsv-tts-lite_b200"))

EOS = 1024
SPARE = 8            # Text2SemanticDecoder.SPARE_SLOTS


class FakeRuntime:
    """Slots with (request, generated tokens, active); a request r generates tokens r*1000 + i and ends with EOS after
    length[r] tokens, or is cut at max_new."""

    def __init__(self, m, lengths, slots, log):
        self.m, self.lengths, self.log = m, lengths, log
        self.slot = [None] * (slots + SPARE)        # dict(r, toks, active, limit) once live; physical slots incl. the spare pool
        self.begun = {}                             # slot -> request whose first half has run
        self.stream = "main"

    # --- what infer_batched calls -------------------------------------------------------------------------------
    def release_all(self):
        self.slot = [None] * len(self.slot)

    def prefill(self, slot, x, y, bert, samp):
        assert self.stream == "main"
        self._activate(slot, int(x[0]), samp)
        self.log.append(("prefill", slot, int(x[0])))

    def prefill_begin_many(self, items):
        """One stacked pass over several idle slots (gsv_gpt_prefill_begin_many)."""
        assert len({it[0] for it in items}) == len(items), "a slot named twice in one pass"
        self.log.append(("begin_many", len(items)))
        return [self.prefill_begin(*it) for it in items]

    def prefill_begin(self, slot, x, y, bert):
        first_wave = not any(e[0] == "decode" for e in self.log)
        assert self.stream == "side" or first_wave, "the first half of a REFILL belongs on the second stream"
        st = self.slot[slot]
        assert st is None or not st["active"], "begin on a slot that is still decoding"
        self.begun[slot] = int(x[0])
        self.log.append(("begin", slot, int(x[0])))
        return x, y, bert

    def prefill_finish(self, slot, y, samp):
        assert self.stream == "main" and slot in self.begun
        first_wave = not any(e[0] == "decode" for e in self.log)
        assert first_wave or (self.log and self.log[-1][0] in ("decode", "wait", "finish")), "finish goes behind a decode launch"
        self._activate(slot, self.begun.pop(slot), samp)
        self.log.append(("finish", slot))

    def decode(self, n):
        if not any(s is not None and s["active"] for s in self.slot):
            # pipelined loop: a launch enqueued before the previous results were read may find every slot finished (a no-op on
            # the device); the one-launch-at-a-time order never does that
            assert self.m.overlap_refill, "decode launched with nothing live"
            self.log.append(("idle_decode", n))
            return
        self.log.append(("decode", n))
        n_active = sum(1 for st in self.slot if st is not None and st["active"])
        assert n_active <= len(self.slot) - SPARE, "more live sequences than the configured batch"
        self.max_active = max(getattr(self, "max_active", 0), n_active)
        for _ in range(n):
            for st in self.slot:
                if st is None or not st["active"]:
                    continue
                i = len(st["toks"])                 # toks[0] is the first sampled token (not returned)
                if i - 1 >= self.lengths[st["r"]]:
                    st["toks"].append(EOS)
                    st["active"] = 0
                else:
                    st["toks"].append(st["r"] * 1000 + i)
                    if st["limit"] and len(st["toks"]) - 1 >= st["limit"]:
                        st["active"] = 0

    def read(self, n_slots, tokens=True):
        m = self.m
        for s, st in enumerate(self.slot):
            m._h_active[s] = 0 if st is None else st["active"]
            m._h_ngen[s] = 0 if st is None else len(st["toks"])
            if st is not None:
                m._h_tokens[s, : len(st["toks"])] = torch.tensor(st["toks"], dtype=torch.int32)
        self.log.append(("read",))

    def _activate(self, slot, r, samp):
        self.slot[slot] = dict(r=r, toks=[r * 1000], active=1, limit=int(samp.max_new_tokens))

    # --- stream plumbing ----------------------------------------------------------------------------------------
    @contextlib.contextmanager
    def on_stream(self, stream):
        prev, self.stream = self.stream, stream
        try:
            yield
        finally:
            self.stream = prev

    def record_event(self, stream):
        return ("event", len(self.log))

    def wait(self, ev):
        self.log.append(("wait",))


def make_model(lengths, slots, log, overlap):
    from gsv_tts.GPT_SoVITS.GPT.t2s_model_b200 import Text2SemanticDecoder
    from gsv_tts import _native as N
    m = Text2SemanticDecoder.__new__(Text2SemanticDecoder)
    torch.nn.Module.__init__(m)
    m._device = torch.device("cpu")
    m.EOS = EOS
    m._buckets = {slots: [256]}
    m._ctx = None
    m.debug_seed = 1
    m.overlap_refill = overlap
    m._max_slots = slots + SPARE
    m._h_ngen = torch.zeros(slots + SPARE, dtype=torch.int32)
    m._h_active = torch.zeros(slots + SPARE, dtype=torch.int32)
    m._h_tokens = torch.zeros(slots + SPARE, 256, dtype=torch.int32)
    rt = FakeRuntime(m, lengths, slots, log)
    m._release_all = rt.release_all
    m._prefill = rt.prefill
    m._prefill_begin = rt.prefill_begin
    m._prefill_begin_many = rt.prefill_begin_many
    m._prefill_finish = rt.prefill_finish
    m._decode = rt.decode
    m._read = rt.read
    m._read_enqueue = lambda n_slots: rt.read(n_slots)          # snapshot at its place in the stream order
    m._read_wait = lambda: log.append(("read_wait",))
    m._side_stream = lambda: "side"
    m._on_stream = rt.on_stream
    m._record_event = rt.record_event
    m._event_done = lambda ev: True
    m._wait_on_current_stream = rt.wait
    m._hold_side_stream = lambda stream: log.append(("hold",))       # the second stream waits for the decode launch to be resident
    return m, N


@pytest.mark.parametrize("overlap", [False, True])
@pytest.mark.parametrize("slots,n_req", [(4, 11), (2, 7), (8, 5), (3, 3)])
def test_every_request_completes_once_with_its_own_tokens(monkeypatch, overlap, slots, n_req):
    g = torch.Generator().manual_seed(slots * 100 + n_req)
    lengths = [int(torch.randint(1, 60, (1,), generator=g)) for _ in range(n_req)]
    max_new = [int(torch.randint(5, 50, (1,), generator=g)) for _ in range(n_req)]
    log = []
    m, N = make_model(lengths, slots, log, overlap)
    monkeypatch.setattr(N, "lib", lambda: type("L", (), {"gsv_gpt_set_noise": staticmethod(lambda *a: 0)})())
    xs = [torch.full((3,), r, dtype=torch.int64) for r in range(n_req)]         # x[0] carries the request id
    ys = [torch.zeros(4, dtype=torch.int64) for _ in range(n_req)]
    bs = [torch.zeros(3, 1024) for _ in range(n_req)]
    outs, order = m.infer_batched(xs, ys, bs, max_new=max_new)
    order = order.tolist()
    assert sorted(order) == list(range(n_req))
    for t, r in zip(outs, order):
        n = min(lengths[r], max_new[r])
        assert t.tolist() == [r * 1000 + i for i in range(1, n + 1)], (r, t.tolist())   # s0 dropped, EOS cut, max_new respected
    kinds = [e[0] for e in log]
    # speculation costs one empty launch each time every live slot finishes inside the same interval
    assert kinds.count("idle_decode") <= n_req
    if overlap:
        # every request goes through begin (stacked passes) + finish; nothing is prefilled synchronously
        assert kinds.count("begin") == kinds.count("finish") == n_req and "prefill" not in kinds
        first_decode = kinds.index("decode")
        wave = min(slots, n_req)
        assert kinds[:first_decode].count("begin_many") == 1 and kinds[:first_decode].count("begin") == wave    # first wave: ONE pass
        # prompts of refills start on the second stream only after it was told to wait for the decode launch in flight
        for i, k in enumerate(kinds):
            if k == "begin_many" and i > first_decode:
                j = max(q for q in range(i) if kinds[q] in ("hold", "decode"))
                assert kinds[j] == "hold"
                n_pass = log[i][1]
                assert kinds[i + 1:i + 1 + n_pass] == ["begin"] * n_pass         # the slots freed at one read share one pass
    else:
        assert "begin" not in kinds and kinds.count("prefill") == n_req
