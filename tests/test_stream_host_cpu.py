"""CPU: the host side of ``Text2SemanticDecoder.infer_stream`` (chunking, one-chunk-late yields, EOS cut, launch-ahead
order) against a restatement of the reference loop (GPT/t2s_model.py:466-553), with the native library replaced by a
scripted token source that follows the kernels' bookkeeping (tokens[0] = first sampled token, EOS appended and
counted, ``active`` cleared at EOS or when the cache is full)."""
import sys
import os

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gsv-tts-lite_b200"))

EOS = 1024


def reference_stream(script, n_iter, stream_chunk, boost_first_chunk, force_steps):
    """t2s_model.py:466-553 over a scripted sampler: script[0] is the first token (prefill), script[i] the token of
    decode step i.  Yields (list of tokens, is_final)."""
    prev = [script[0]]
    first, pre_chunk, idx = True, None, 0
    if force_steps is not None:
        n_iter = min(n_iter, force_steps)
    for idx in range(1, n_iter + 1):
        tok = script[idx]
        if tok == EOS:
            break
        prev.append(tok)
        if idx % stream_chunk == 0:
            if pre_chunk is not None:
                yield pre_chunk, False
            pre_chunk = prev[-idx:]
            if boost_first_chunk and first:
                first = False
                yield pre_chunk, False
                pre_chunk = None
    yield prev[-idx:] if idx > 0 else [], True


class FakeRuntime:
    """Stands for libgsv_b200 behind the three calls infer_stream makes."""

    def __init__(self, model, script, n_iter, log):
        self.m, self.script, self.n_iter, self.log = model, script, n_iter, log

    def single_setup(self, *a):
        m = self.m
        m._h_tokens.zero_()
        m._h_tokens[0, 0] = self.script[0]
        m._state = dict(n_gen=1, active=1)

    def decode(self, n):
        st = self.m._state
        self.log.append(("decode", n))
        for _ in range(n):
            if not st["active"]:
                break
            step = st["n_gen"]                       # decode step index (1-based)
            tok = self.script[step]
            self.m._h_tokens[0, st["n_gen"]] = tok
            st["n_gen"] += 1
            if tok == EOS or step >= self.n_iter:
                st["active"] = 0

    def read(self, n_slots, tokens=True):
        st = self.m._state
        self.m._h_ngen[0] = st["n_gen"]
        self.m._h_active[0] = st["active"]


def make_model(script, n_iter, log):
    from gsv_tts.GPT_SoVITS.GPT.t2s_model_b200 import Text2SemanticDecoder
    m = Text2SemanticDecoder.__new__(Text2SemanticDecoder)
    torch.nn.Module.__init__(m)
    m._device = torch.device("cpu")
    m.EOS = EOS
    m._h_ngen = torch.zeros(1, dtype=torch.int32)
    m._h_active = torch.zeros(1, dtype=torch.int32)
    m._h_tokens = torch.zeros(1, 2048, dtype=torch.int32)
    rt = FakeRuntime(m, script, n_iter, log)
    m._single_setup = rt.single_setup
    m._decode = rt.decode
    m._read = rt.read
    m._read_enqueue = lambda n_slots: (rt.read(n_slots), log.append(("read",)))      # snapshot at its place in the stream order
    m._read_wait = lambda: None
    m._mark_chunk_ready = lambda: log.append(("ready",))
    # stream plumbing of the real class (a high-priority CUDA stream for the decode launches): no-ops here
    import contextlib
    m._high_priority_stream = lambda: "hp"
    m._stream_waits_for_current = lambda stream: None
    m._on_stream = lambda stream: contextlib.nullcontext()
    return m


def scripted(n, eos_at, seed):
    g = torch.Generator().manual_seed(seed)
    s = torch.randint(0, 1024, (n + 2,), generator=g).tolist()
    if eos_at is not None:
        s[eos_at] = EOS
    return s


@pytest.mark.parametrize("boost", [True, False])
@pytest.mark.parametrize("eos_at,n_iter,force", [
    (None, 60, None),        # the cache fills first
    (3, 200, None),          # EOS inside the first chunk
    (10, 200, None),         # EOS on the step that would have closed chunk 1
    (11, 200, None),         # EOS right after a chunk boundary
    (37, 200, None),         # EOS in the fourth chunk
    (1, 200, None),          # EOS on the first decode step
    (None, 200, 25),         # forced length, not a multiple of the chunk
    (None, 200, 30),         # forced length on a chunk boundary
])
def test_infer_stream_yields_match_reference_loop(boost, eos_at, n_iter, force):
    chunk = 10
    script = scripted(max(n_iter, 64), eos_at, 7)
    log = []
    m = make_model(script, n_iter if force is None else min(n_iter, 10 ** 9), log)
    x = torch.zeros(1, 4, dtype=torch.int64)
    got = [(t.view(-1).tolist(), f) for t, f in
           m.infer_stream(x, x, torch.zeros(1, 4, 1024), stream_chunk=chunk, boost_first_chunk=boost, force_steps=force)]
    want = list(reference_stream(script, n_iter, chunk, boost, force))
    assert got == want


def test_two_launches_stay_in_flight_and_chunks_are_handed_out_behind_them():
    chunk = 10
    script = scripted(64, 45, 3)
    log = []
    m = make_model(script, 200, log)
    x = torch.zeros(1, 4, dtype=torch.int64)
    for t, f in m.infer_stream(x, x, torch.zeros(1, 4, 1024), stream_chunk=chunk):
        log.append(("yield", t.numel(), f))
    kinds = [e[0] for e in log]
    # stream order: decode 1, read 1, decode 2 | read 2, <chunk handed out>, decode 3 | ...: while launch k runs, launch k+1 is
    # queued behind the copy of k's results, so the GPU never waits for the host; a chunk is handed out with the next launch
    # running (its read already enqueued) and the one after that is queued when the caller comes back -- after the caller's
    # residency hold (gsv_gpt_wait_resident), which must not wait for a launch that cannot start yet
    assert kinds[:7] == ["decode", "read", "decode", "ready", "read", "yield", "decode"]
    ys = [i for i, k in enumerate(kinds) if k == "yield"]
    for i in ys[:-1]:
        n_dec = kinds[:i].count("decode")
        n_chunks_done = kinds[:i].count("ready")
        assert n_dec == n_chunks_done + 1                          # exactly one launch beyond the chunks already complete
        assert kinds[i + 1] == "decode"                            # the next one is queued as soon as the caller returns
    # a read is always enqueued between two launches (results of launch k are copied before launch k+1 may change them)
    for a, b in zip([i for i, k in enumerate(kinds) if k == "decode"][:-1], [i for i, k in enumerate(kinds) if k == "decode"][1:]):
        assert "read" in kinds[a:b]
    # EOS at step 45: five launches of 10 steps do the work, the sixth was queued speculatively and finds the sequence stopped
    assert sum(1 for k in kinds if k == "decode") == 6
    assert kinds[-2:] == ["ready", "yield"] and log[-1][2] is True


@pytest.mark.parametrize("boost", [True, False])
@pytest.mark.parametrize("eos_at,force", [(None, 40), (37, None), (31, None), (None, 34), (5, None), (None, 9)])
def test_held_chunks_are_reported_early_and_only_when_they_can_still_be_yielded(boost, eos_at, force):
    """``on_chunk_held``: called with a chunk the moment it is complete when the reference's order holds it back (never for
    the boosted first chunk, which is yielded at once, and never for a chunk already known to be dropped because the stream
    has stopped); the object reported is the object yielded later, or it is never yielded (the stream ended first and the
    final yield covers it) -- and the yields themselves are those of the reference loop."""
    chunk = 10
    script = scripted(64, eos_at, 11)
    log = []
    m = make_model(script, 200, log)
    x = torch.zeros(1, 4, dtype=torch.int64)
    held, events = [], []

    def on_held(t):
        held.append(t)
        events.append(("held", t.numel()))

    got = []
    for t, f in m.infer_stream(x, x, torch.zeros(1, 4, 1024), stream_chunk=chunk, boost_first_chunk=boost, force_steps=force,
                               on_chunk_held=on_held):
        got.append((t, f))
        events.append(("yield", t.numel(), f))
    want = list(reference_stream(script, 200, chunk, boost, force))
    assert [(t.view(-1).tolist(), f) for t, f in got] == want
    yielded_nonfinal = [t for t, f in got if not f]
    first_boosted = 1 if (boost and yielded_nonfinal) else 0
    # every non-final yield except the boosted first chunk was reported before, as the same object, in the same order
    assert len(held) >= len(yielded_nonfinal) - first_boosted
    for t, h in zip(yielded_nonfinal[first_boosted:], held):
        assert t is h
    # at most one reported chunk is never yielded: the one the final yield covers
    assert len(held) - (len(yielded_nonfinal) - first_boosted) in (0, 1)
    # a report always precedes its yield, with exactly one other yield (the previous chunk) allowed in between
    for h in held:
        i = events.index(("held", h.numel()))
        later = [e for e in events[i + 1:] if e[0] == "yield" and e[1] == h.numel() and not e[2]]
        assert len(later) <= 1
    # a stream that has stopped on a chunk boundary does not report the chunk it is about to drop
    if force is not None and force % chunk == 0 and eos_at is None:
        assert all(h.numel() < force for h in held)
