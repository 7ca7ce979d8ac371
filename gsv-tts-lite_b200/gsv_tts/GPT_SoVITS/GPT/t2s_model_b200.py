"""B200 backend of the text-to-semantic decoder.

Third backend next to the reference's SDPA and FlashAttention ones (reference
gsv_tts/GPT_SoVITS/GPT/t2s_model.py and t2s_model_flash_attn.py, selected in
gsv_tts/Loader.py:117-121, 156-160).  Same constructor, same state-dict keys, same
``initialize_runtime`` / ``infer`` / ``infer_stream`` / ``infer_batched`` signatures and return
shapes; the arithmetic runs in ``libgsv_b200.so`` (prefill kernels + one persistent decode
kernel per launch).  This class only owns parameters, packs them for the kernels and drives
the slot life cycle.

Differences a caller can observe (documented):
  * decode stops at the first EOS on the device; the reference tests EOS every 5th step and cuts
    the overshoot afterwards (t2s_model.py:451-464) -- same returned tokens.
  * the noise for ``argmax(p / Exp(1))`` comes from a per-slot Philox stream seeded from torch's
    global generator, not from ``Tensor.exponential_``; tests inject identical noise instead.
  * idle slots are skipped instead of stepped with kv_len reset to 0 (t2s_model.py:684-695).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterator, List, Optional, Tuple

import torch
from torch import nn

from ... import _native as N


class _TokenEmbedding(nn.Module):
    """Parameter holder with the reference's key ``word_embeddings.weight`` (embedding.py:7-32)."""

    def __init__(self, dim: int, vocab: int):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab, dim)


class _SinePositionalEmbedding(nn.Module):
    """Holds ``alpha`` (embedding.py:35-50); the table itself is built in initialize_runtime."""

    def __init__(self):
        super().__init__()
        self.alpha = nn.Parameter(torch.ones(1))


class _Block(nn.Module):
    def __init__(self, d: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(d)
        self.qkv = nn.Linear(d, 3 * d)
        self.out_proj = nn.Linear(d, d)
        self.norm2 = nn.LayerNorm(d)
        self.mlp = nn.Sequential(nn.Linear(d, 4 * d), nn.ReLU(inplace=True), nn.Linear(4 * d, d))


class _Transformer(nn.Module):
    def __init__(self, n: int, d: int):
        super().__init__()
        self.blocks = nn.ModuleList([_Block(d) for _ in range(n)])


def sine_table(n_pos: int, dim: int) -> torch.Tensor:
    """pe[p,2j]=sin(p w_j), pe[p,2j+1]=cos(p w_j), w_j=exp(-2j ln(1e4)/dim), fp32 (embedding.py:52-69)."""
    pos = torch.arange(n_pos, dtype=torch.float32).unsqueeze(1)
    w = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * -(math.log(10000.0) / dim))
    pe = torch.zeros(n_pos, dim, dtype=torch.float32)
    pe[:, 0::2] = torch.sin(pos * w)
    pe[:, 1::2] = torch.cos(pos * w)
    return pe


class Text2SemanticDecoder(nn.Module):
    N_POS = 4000                      # t2s_model.py:212-213
    DECODE_CHUNK = 32                 # decode steps per persistent launch in infer()
    BATCH_INTERVAL = 8                # decode steps between harvest/refill points in infer_batched()
    SPARE_SLOTS = 8                   # infer_batched: requests kept prefilled beside a full batch (slots beyond the batch size)
    _slot_audit = None                # test hook: callable(slot, request) -> (noise, forced, trace) device tensors or None
                                      # each; attached to the slot right before the request's first sample (infer_batched)

    def __init__(self, config):
        super().__init__()
        m = config["model"]
        self.model_dim = m["hidden_dim"]
        self.embedding_dim = m["embedding_dim"]
        self.num_head = m["head"]
        self.num_layers = m["n_layer"]
        self.vocab_size = m["vocab_size"]
        self.phoneme_vocab_size = m["phoneme_vocab_size"]
        self.p_dropout = m["dropout"]
        self.EOS = m["EOS"]
        if self.model_dim != self.embedding_dim:
            raise ValueError("hidden_dim must equal embedding_dim")
        self.bert_proj = nn.Linear(1024, self.embedding_dim)
        self.ar_text_embedding = _TokenEmbedding(self.embedding_dim, self.phoneme_vocab_size)
        self.ar_text_position = _SinePositionalEmbedding()
        self.ar_audio_embedding = _TokenEmbedding(self.embedding_dim, self.vocab_size)
        self.ar_audio_position = _SinePositionalEmbedding()
        self.ar_predict_layer = nn.Linear(self.model_dim, self.vocab_size, bias=False)
        self.t2s_transformer = _Transformer(self.num_layers, self.model_dim)
        self._ctx = None
        self._packed = {}
        self._buckets = {}            # batch size -> sorted list of lengths (gpt_cache)
        self._parity_noise = None     # test hook: [rows][V] fp32 Exp(1) noise for slot 0
        self.debug_seed: Optional[int] = None
        self.overlap_refill = True          # infer_batched: prompts of refills on a second stream (False: reference order)
        self._refill_stream = None
        self._decode_stream = None

    # ------------------------------------------------------------------ runtime set-up
    @torch.inference_mode()
    def initialize_runtime(self, dtype, device, gpt_cache):
        """Counterpart of t2s_model.py:210-298: size the KV cache for the largest (B, S) bucket and
        hand packed weights to the native context.  Nested buckets collapse to "attend over the
        live kv_len" (SURVEY.md A.5); only the per-batch-size length cap is kept."""
        device = torch.device(device)
        if device.type != "cuda":
            raise N.NativeError("the B200 backend needs a CUDA device; there is no CPU path here")
        self._device, self._dtype = device, dtype
        for b, s in gpt_cache:
            self._buckets.setdefault(int(b), []).append(int(s))
        for b in self._buckets:
            self._buckets[b].sort()
        max_slots = max(self._buckets)
        if max_slots >= 8:                                  # room for prefilled requests beside a full batch (infer_batched)
            max_slots = min(64, max_slots + self.SPARE_SLOTS)
        max_seq = max(max(v) for v in self._buckets.values())
        d, L = self.model_dim, self.num_layers
        blk = self.t2s_transformer.blocks

        def stack(get):
            return torch.stack([get(b).detach() for b in blk]).to(device=device, dtype=dtype).contiguous()

        pe = sine_table(self.N_POS, d).to(device=device, dtype=dtype)
        pk = {
            "w_qkv": stack(lambda b: b.qkv.weight), "b_qkv": stack(lambda b: b.qkv.bias),
            "w_o": stack(lambda b: b.out_proj.weight), "b_o": stack(lambda b: b.out_proj.bias),
            "w_1": stack(lambda b: b.mlp[0].weight), "b_1": stack(lambda b: b.mlp[0].bias),
            "w_2": stack(lambda b: b.mlp[2].weight), "b_2": stack(lambda b: b.mlp[2].bias),
            "ln1_g": stack(lambda b: b.norm1.weight), "ln1_b": stack(lambda b: b.norm1.bias),
            "ln2_g": stack(lambda b: b.norm2.weight), "ln2_b": stack(lambda b: b.norm2.bias),
        }

        def one(t):
            return t.detach().to(device=device, dtype=dtype).contiguous()

        pk["w_head"] = one(self.ar_predict_layer.weight)
        pk["emb_audio"] = one(self.ar_audio_embedding.word_embeddings.weight)
        pk["emb_text"] = one(self.ar_text_embedding.word_embeddings.weight)
        # alpha * pe in the storage dtype, as the reference computes pe_cache (t2s_model.py:409)
        pk["pe_audio"] = (one(self.ar_audio_position.alpha) * pe).contiguous()
        pk["pe_text"] = (one(self.ar_text_position.alpha) * pe).contiguous()
        pk["w_bert"] = one(self.bert_proj.weight)
        pk["b_bert"] = one(self.bert_proj.bias)
        self._packed = pk
        dims = N.GptDims(d_model=d, n_head=self.num_head, n_layer=L, d_ff=4 * d, vocab=self.vocab_size,
                         eos=self.EOS, n_phoneme=self.phoneme_vocab_size, d_bert=1024, n_pos=self.N_POS,
                         dtype=N.dtype_code(dtype), max_slots=max_slots, max_seq=max_seq)
        w = N.GptWeights(**{k: v.data_ptr() for k, v in pk.items()})
        ctx = C.c_void_p()
        with torch.cuda.device(device):
            N.check(N.lib().gsv_gpt_create(C.byref(dims), C.byref(w), C.byref(ctx)))
        self._ctx = ctx
        self._max_slots, self._max_seq = max_slots, max_seq
        # pinned read-back buffers
        self._h_ngen = torch.zeros(max_slots, dtype=torch.int32).pin_memory()
        self._h_active = torch.zeros(max_slots, dtype=torch.int32).pin_memory()
        self._h_tokens = torch.zeros(max_slots, max_seq, dtype=torch.int32).pin_memory()

    def __del__(self):
        try:
            if self._ctx is not None:
                N.lib().gsv_gpt_destroy(self._ctx)
                self._ctx = None
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)

    def _next_seed(self) -> int:
        if self.debug_seed is not None:
            return self.debug_seed
        return int(torch.randint(0, 2 ** 62, (1,)).item())      # follows torch.manual_seed

    def _prefill(self, slot, x, y, bert, samp: N.GptSampling):
        x = x.to(device=self._device, dtype=torch.int64).contiguous().view(-1)
        y = y.to(device=self._device, dtype=torch.int64).contiguous().view(-1)
        bert = bert.to(device=self._device, dtype=self._dtype).contiguous().view(x.numel(), -1)
        N.check(N.lib().gsv_gpt_prefill(self._ctx, slot, x.data_ptr(), x.numel(), y.data_ptr(), y.numel(),
                                        bert.data_ptr(), C.byref(samp), self._stream()))

    def _prefill_begin(self, slot, x, y, bert):
        """First half of a prefill on the CURRENT stream (``gsv_gpt_prefill_begin``); returns the tensors the second
        half needs alive."""
        x = x.to(device=self._device, dtype=torch.int64).contiguous().view(-1)
        y = y.to(device=self._device, dtype=torch.int64).contiguous().view(-1)
        bert = bert.to(device=self._device, dtype=self._dtype).contiguous().view(x.numel(), -1)
        N.check(N.lib().gsv_gpt_prefill_begin(self._ctx, slot, x.data_ptr(), x.numel(), y.data_ptr(), y.numel(),
                                              bert.data_ptr(), self._stream()))
        return x, y, bert

    def _prefill_begin_many(self, items):
        """First half of the prefills of several idle slots on the CURRENT stream, in as few passes as their stacked rows
        allow (``gsv_gpt_prefill_begin_many``: one tensor-core launch per linear over all prompts of a pass).
        ``items`` = [(slot, x, y, bert)]; returns the device tensors each second half needs alive, in order."""
        lib = N.lib()
        cap = int(lib.gsv_gpt_prefill_capacity(self._ctx))
        keep = []
        for slot, x, y, bert in items:
            x = x.to(device=self._device, dtype=torch.int64).contiguous().view(-1)
            y = y.to(device=self._device, dtype=torch.int64).contiguous().view(-1)
            bert = bert.to(device=self._device, dtype=self._dtype).contiguous().view(x.numel(), -1)
            keep.append((x, y, bert))
        i = 0
        while i < len(items):
            j, rows = i, 0
            while j < len(items) and j - i < self._max_slots and rows + keep[j][0].numel() + keep[j][1].numel() <= cap:
                rows += keep[j][0].numel() + keep[j][1].numel()
                j += 1
            j = max(j, i + 1)                      # a single over-long prompt is the library's argument error to report
            n = j - i
            slots_a = (C.c_int * n)(*[items[k][0] for k in range(i, j)])
            xs = (C.c_void_p * n)(*[keep[k][0].data_ptr() for k in range(i, j)])
            nxs = (C.c_int * n)(*[keep[k][0].numel() for k in range(i, j)])
            ys = (C.c_void_p * n)(*[keep[k][1].data_ptr() for k in range(i, j)])
            nys = (C.c_int * n)(*[keep[k][1].numel() for k in range(i, j)])
            bs = (C.c_void_p * n)(*[keep[k][2].data_ptr() for k in range(i, j)])
            N.check(lib.gsv_gpt_prefill_begin_many(self._ctx, n, slots_a, xs, nxs, ys, nys, bs, self._stream()))
            i = j
        return keep

    def _prefill_finish(self, slot, y, samp: N.GptSampling):
        N.check(N.lib().gsv_gpt_prefill_finish(self._ctx, slot, y.data_ptr(), y.numel(), C.byref(samp), self._stream()))

    def set_decode_sms(self, n_sms: int):
        """Leave ``num_sms - n_sms`` SMs free while one sequence decodes (0 = use all): room for the vocoder of the
        previous chunk on another stream (``TTS.infer_features_stream``)."""
        N.check(N.lib().gsv_gpt_set_decode_sms(self._ctx, int(n_sms)))

    # second-stream plumbing of infer_batched (separate methods so that the CPU tests can script them)
    def _side_stream(self):
        if self._refill_stream is None:
            self._refill_stream = torch.cuda.Stream(self._device)
        return self._refill_stream

    def _on_stream(self, stream):
        return torch.cuda.stream(stream)

    def _record_event(self, stream):
        ev = torch.cuda.Event()
        ev.record(stream)
        return ev

    def _event_done(self, ev) -> bool:
        return ev.query()

    def _hold_side_stream(self, stream):
        self.hold_until_decode_resident(stream)

    def _wait_on_current_stream(self, ev):
        torch.cuda.current_stream(self._device).wait_event(ev)

    def _high_priority_stream(self):
        """Stream of the streaming decode: higher priority than the caller's streams.  The single-sequence kernel needs all
        of its thread-block clusters resident at once; behind a second stream that keeps launching small vocoder kernels its
        clusters could wait tens of milliseconds for 8 free SMs in one GPC (measured: 46 -> 144 ms per utterance, at random).
        Pending blocks of a higher-priority stream are placed first."""
        if self._decode_stream is None:
            self._decode_stream = torch.cuda.Stream(self._device, priority=-1)
        return self._decode_stream

    def hold_until_decode_resident(self, stream=None):
        """Enqueue on ``stream`` (default: the current one) a wait for the decode launch in flight to occupy its SMs
        (``gsv_gpt_wait_resident``): call it on the vocoder stream before the chunk's work."""
        st = stream if stream is not None else torch.cuda.current_stream(self._device)
        N.check(N.lib().gsv_gpt_wait_resident(self._ctx, C.c_void_p(st.cuda_stream)))

    def _stream_waits_for_current(self, stream):
        stream.wait_stream(torch.cuda.current_stream(self._device))

    def _mark_chunk_ready(self):
        """Event behind the host->device copy of the chunk about to be handed out: a consumer on another stream waits
        for this, not for the decode launched after it."""
        self.chunk_ready = torch.cuda.Event()
        self.chunk_ready.record(torch.cuda.current_stream(self._device))

    def _decode(self, n_steps: int):
        N.check(N.lib().gsv_gpt_decode(self._ctx, n_steps, self._stream()))

    def _read(self, n_slots: int, tokens: bool = True):
        N.check(N.lib().gsv_gpt_read(self._ctx, self._h_ngen.data_ptr(), self._h_active.data_ptr(),
                                     self._h_tokens.data_ptr() if tokens else None, 0, n_slots, self._stream()))
        torch.cuda.current_stream(self._device).synchronize()

    def _read_enqueue(self, n_slots: int):
        """Asynchronous half of ``_read``: the copies into the pinned host buffers are enqueued on the current stream and an
        event is recorded behind them; ``_read_wait`` blocks on that event only, so launches enqueued after this call run on."""
        N.check(N.lib().gsv_gpt_read(self._ctx, self._h_ngen.data_ptr(), self._h_active.data_ptr(), self._h_tokens.data_ptr(),
                                     0, n_slots, self._stream()))
        self._read_event = torch.cuda.Event()
        self._read_event.record(torch.cuda.current_stream(self._device))

    def _read_wait(self):
        self._read_event.synchronize()

    def _release_all(self):
        N.check(N.lib().gsv_gpt_release_slot(self._ctx, -1, self._stream()))

    def _single_setup(self, x, y, bert, top_k, top_p, temperature, repetition_penalty, suppress_steps,
                      force_steps: Optional[int]):
        if 1 not in self._buckets:
            raise KeyError("gpt_cache has no batch-size-1 bucket")        # same failure as t2s_model.py:400
        self._release_all()
        samp = N.GptSampling(top_k=top_k if top_k is not None else 0, top_p=top_p if top_p is not None else 1.0,
                             temperature=temperature, repetition_penalty=repetition_penalty,
                             suppress_steps=suppress_steps, max_new_tokens=0,
                             mask_eos=1 if force_steps is not None else 0, max_kv=self._buckets[1][-1],
                             suppress_first=1, seed=self._next_seed())
        if self._parity_noise is not None:
            N.check(N.lib().gsv_gpt_set_noise(self._ctx, self._parity_noise.data_ptr(), self._parity_noise.shape[0]))
        else:
            N.check(N.lib().gsv_gpt_set_noise(self._ctx, None, 0))
        self._prefill(0, x, y, bert, samp)

    # ------------------------------------------------------------------ reference entry points
    @torch.inference_mode()
    def infer(self, x, y, bert_feature, top_k: int = 15, top_p: float = 1.0, temperature: float = 1.0,
              repetition_penalty: float = 1.35, initial_suppression_steps: int = 10, check_interval: int = 5,
              force_steps: Optional[int] = None):
        """t2s_model.py:385-464.  x [1,Nx] int64, y [1,Ny] int64, bert [1,Nx,1024] -> int64 [1,1,N]."""
        self._single_setup(x, y, bert_feature, top_k, top_p, temperature, repetition_penalty,
                           initial_suppression_steps, force_steps)
        done_steps = 0
        while True:
            n = self.DECODE_CHUNK
            if force_steps is not None:
                n = min(n, force_steps - done_steps)
                if n <= 0:
                    break
            self._decode(n)
            done_steps += n
            self._read(1, tokens=False)
            if not int(self._h_active[0]):
                break
        self._read(1)
        n_gen = int(self._h_ngen[0])
        toks = self._h_tokens[0, :n_gen].to(torch.int64)
        # tokens[0] is the first sampled token, which the reference never returns (:458); cut at EOS (:459-464)
        out = toks[1:]
        if out.numel() and int(out[-1]) == self.EOS:
            out = out[:-1]
        return out.to(self._device).view(1, 1, -1)

    @torch.inference_mode()
    def infer_stream(self, x, y, bert_feature, top_k: int = 15, top_p: float = 1.0, temperature: float = 1.0,
                     repetition_penalty: float = 1.35, initial_suppression_steps: int = 10, stream_chunk: int = 25,
                     boost_first_chunk: bool = True, debug: bool = True,
                     force_steps: Optional[int] = None, on_chunk_held=None) -> Iterator[Tuple[torch.Tensor, bool]]:
        """t2s_model.py:466-553: yields (tokens so far [1,1,n], is_final) every ``stream_chunk`` tokens,
        one chunk late unless ``boost_first_chunk``; the final yield after an EOS break carries the
        first sampled token (the reference slices ``[-idx:]`` with idx one past the appended count).
        ``on_chunk_held(tokens)`` (not in the reference) is called with a chunk the moment it is complete when the
        reference's order holds it back until the next one exists -- the same tensor object is yielded later, or never (the
        stream ended first and the final yield covers it); it is not called for a chunk that is already known to be
        dropped.  A consumer may start that chunk's work early, speculatively."""
        hp = self._high_priority_stream()
        self._stream_waits_for_current(hp)                                # inputs produced on the caller's stream
        with self._on_stream(hp):
            self._single_setup(x, y, bert_feature, top_k, top_p, temperature, repetition_penalty,
                               initial_suppression_steps, force_steps)
        first, pre_chunk, idx = True, None, 0
        launched = 0                                  # decode steps enqueued so far

        def launch() -> bool:
            nonlocal launched
            n = stream_chunk
            if force_steps is not None:
                n = min(n, force_steps - launched)
            if n <= 0:
                return False
            with self._on_stream(hp):
                self._decode(n)
            launched += n
            return True

        def enqueue_read():
            with self._on_stream(hp):
                self._read_enqueue(1)

        # Two launches are kept in the stream: while launch k runs, launch k+1 is already queued behind the copy of k's
        # results, so the GPU never waits for the host between chunks (a launch that finds its sequence stopped returns at
        # once).  Stream order: decode 1, read 1, decode 2 | read 2, decode 3 | ...
        in_flight = launch()
        if in_flight:
            enqueue_read()
            queued = launch()
        while in_flight:
            self._read_wait()
            n_gen = int(self._h_ngen[0])            # s0 + decode steps so far
            active = int(self._h_active[0])
            toks = self._h_tokens[0, :n_gen].to(torch.int64)
            steps = n_gen - 1
            if not active and int(toks[-1]) == self.EOS:
                # EOS sampled at decode step `steps`: loop broke before the append (:534-535)
                final = toks[0:steps].to(self._device).view(1, 1, -1)
                self._mark_chunk_ready()
                yield final, True
                return
            idx = steps
            chunk = None
            if idx % stream_chunk == 0 and idx > 0:
                chunk = toks[1:idx + 1].to(self._device).view(1, 1, -1)
                self._mark_chunk_ready()
            in_flight = bool(active) and queued
            if in_flight:
                enqueue_read()                        # behind the launch already queued
            # The chunk is handed out with the next launch running or queued: whatever the caller does with it (prior encoder
            # and vocoder, on its own stream after `chunk_ready` and `hold_until_decode_resident`) overlaps the decode instead
            # of delaying it.  The launch after that is queued when the caller comes back -- after it has taken its residency
            # hold, which must not wait for a launch that cannot start yet.  The single-sequence kernel occupies 64 SMs; the
            # other 84 are the caller's.
            if chunk is not None:
                if pre_chunk is not None:
                    yield pre_chunk, False
                pre_chunk = chunk
                if boost_first_chunk and first:
                    first = False
                    yield pre_chunk, False
                    pre_chunk = None
                elif in_flight and on_chunk_held is not None:
                    on_chunk_held(pre_chunk)
            if in_flight:
                queued = launch()
        if not launched:                              # nothing to decode (force_steps = 0): the first sampled token alone
            enqueue_read()
            self._read_wait()
        n_gen = int(self._h_ngen[0])
        toks = self._h_tokens[0, :n_gen].to(torch.int64)
        out = toks[1:].to(self._device).view(1, 1, -1)
        self._mark_chunk_ready()
        yield out, True

    @torch.inference_mode()
    def infer_batched(self, x: List[torch.Tensor], y: List[torch.Tensor], bert_feature: List[torch.Tensor],
                      top_k: int = 15, top_p: float = 1.0, temperature: float = 1.0, repetition_penalty: float = 1.35,
                      check_interval: int = 5, max_new: Optional[List[int]] = None, on_finish=None, on_launch=None):
        """t2s_model.py:555-734: continuous batching.  Slot count = smallest configured batch >= B,
        else the largest (:569-574); no repetition penalty, no token suppression (:613, :651);
        finished rows are harvested and their slot refilled from the queue (:672-722).
        Returns (list of int64 [n_i] in completion order, int64 [B] original indices).
        Two hooks let a caller overlap its next stage with the decode (``TTS.infer_phones_batched``): ``on_finish(r, tokens)``
        at the harvest of request r, ``on_launch(decoding)`` right after every decode launch -- the place to enqueue work
        on another stream behind ``hold_until_decode_resident``."""
        B = len(x)
        slots = None
        for b in sorted(self._buckets):
            slots = b
            if b >= B:
                break
        max_kv = self._buckets[slots][-1]
        self._release_all()
        N.check(N.lib().gsv_gpt_set_noise(self._ctx, None, 0))
        base_seed = self._next_seed()

        def sampling(r):
            return N.GptSampling(top_k=top_k if top_k is not None else 0, top_p=top_p if top_p is not None else 1.0,
                                 temperature=temperature, repetition_penalty=1.0, suppress_steps=0,
                                 max_new_tokens=(max_new[r] if max_new is not None else 0), mask_eos=0,
                                 max_kv=max_kv, suppress_first=0,
                                 seed=(base_seed + 0x9E3779B97F4A7C15 * (r + 1)) & (2 ** 64 - 1))

        def audit(slot, r):
            if self._slot_audit is not None:
                noise, forced, trace = self._slot_audit(slot, r)
                N.check(N.lib().gsv_gpt_set_slot_hooks(
                    self._ctx, slot, noise.data_ptr() if noise is not None else None, noise.shape[0] if noise is not None else 0,
                    forced.data_ptr() if forced is not None else None, forced.numel() if forced is not None else 0,
                    trace.data_ptr() if trace is not None else None, trace.shape[0] if trace is not None else 0, self._stream()))

        def start(slot, r):
            audit(slot, r)
            self._prefill(slot, x[r], y[r], bert_feature[r], sampling(r))

        FREE, REFILLING = -1, -2
        # Physical slots: `slots` of them decode at a time (the configured batch); with the overlapped refill up to SPARE_SLOTS
        # more hold requests whose prompts are already computed, so that a slot freed at one harvest is replaced at the next
        # launch boundary instead of after a prompt pass (~190 dependent launches beside the decode kernel).
        total = min(self._max_slots, slots + self.SPARE_SLOTS) if self.overlap_refill else slots
        owner = [FREE] * total
        nxt = 0
        n0 = min(slots, B)
        if self.overlap_refill:
            # the first wave in stacked passes (the reference prefills it as ONE padded batch, t2s_model.py:576-634)
            keeps = self._prefill_begin_many([(s, x[s], y[s], bert_feature[s]) for s in range(n0)])
            for s in range(n0):
                audit(s, s)
                self._prefill_finish(s, keeps[s][1], sampling(s))
                owner[s] = s
            del keeps
            nxt = n0
            # the passes above ran on THIS stream and share their scratch with the passes of the second stream: a real
            # dependency (the residency hold below is a bounded wait, a scheduling hint, not an ordering)
            first_wave = self._record_event(torch.cuda.current_stream(self._device) if self._device.type == "cuda" else None)
        else:
            for s in range(n0):
                start(s, nxt)
                owner[s] = nxt
                nxt += 1
        results, order = [], []
        interval = max(int(check_interval), self.BATCH_INTERVAL)
        side = self._side_stream() if self.overlap_refill else None
        pending = []                      # prompts being computed on `side`: (slot, request, tensors, event)
        ready = []                        # prompts computed, waiting for a place in the batch: (slot, request, tensors, event)

        def harvest(skip):
            nonlocal nxt
            for s in range(total):
                if s in skip or owner[s] < 0 or int(self._h_active[s]):
                    continue
                n_gen = int(self._h_ngen[s])
                toks = self._h_tokens[s, 1:n_gen].to(torch.int64)
                if toks.numel() and int(toks[-1]) == self.EOS:
                    toks = toks[:-1]
                results.append(toks.to(self._device))
                order.append(owner[s])
                if on_finish is not None:
                    on_finish(owner[s], results[-1])
                owner[s] = FREE
                if side is None and nxt < B:
                    start(s, nxt)
                    owner[s] = nxt
                    nxt += 1

        if side is None:
            # the reference's order: launch, wait, harvest, prefill the successors between two launches
            while any(o >= 0 for o in owner):
                self._decode(interval)
                if on_launch is not None:
                    on_launch(True)
                self._read(total)
                harvest(())
            return results, torch.tensor(order, device=self._device)

        # Overlapped refill, pipelined one launch deep: launch k+1 is enqueued BEFORE the results of launch k are waited for
        # (stream order: decode k, read k, decode k+1, read k+1, ...), so the GPU does not idle while the host harvests.  A
        # launch enqueued over slots that turn out to have finished simply skips them (the kernels look at `active` on the
        # device); a slot published behind launch k+1 is not in the copy of launch k's results (`fresh`).
        read_pending = False
        keep_alive = []                   # prompt tensors of published requests, until the publication has certainly run
        while any(o >= 0 for o in owner) or pending or ready or read_pending or nxt < B:
            # publication first -- behind the launch in flight, before the next one: computed prompts take the places the last
            # harvest freed and decode from the very next launch (a pass still running is not waited for, unless nothing else
            # is left to run)
            idle = not any(o >= 0 for o in owner) and not read_pending
            ready += [q for q in pending if idle or self._event_done(q[3])]
            pending = [q for q in pending if q not in ready]
            fresh, published = set(), []
            n_live = sum(1 for o in owner if o >= 0)
            while ready and n_live < slots:
                slot, r, keep, ev = ready.pop(0)
                self._wait_on_current_stream(ev)
                audit(slot, r)
                self._prefill_finish(slot, keep[1], sampling(r))
                owner[slot] = r
                fresh.add(slot)
                published.append(keep)
                n_live += 1
            launched = n_live > 0
            if launched:
                self._decode(interval)
            # new prompt passes on the second stream for the requests that will be needed next: enough to fill the batch plus
            # the spare pool -- held until the launch just enqueued has its clusters resident (hold_until_decode_resident): a stream
            # of small prompt kernels must not be what a 16-CTA cluster waits behind
            want = (slots - n_live) + (total - slots) - len(ready) - len(pending)
            free = [s for s in range(total) if owner[s] == FREE]
            n_new = min(want, len(free), B - nxt)
            if n_new > 0:
                batch = []
                for s in free[:n_new]:
                    batch.append((s, nxt))
                    owner[s] = REFILLING
                    nxt += 1
                with self._on_stream(side):
                    if first_wave is not None:
                        self._wait_on_current_stream(first_wave)
                        first_wave = None
                    self._hold_side_stream(side)
                    keeps = self._prefill_begin_many([(slot, x[r], y[r], bert_feature[r]) for slot, r in batch])
                    ev = self._record_event(side)
                    for (slot, r), keep in zip(batch, keeps):
                        pending.append((slot, r, keep, ev))
            if on_launch is not None:
                on_launch(launched)
            if read_pending:
                self._read_wait()         # results of the PREVIOUS launch; the one enqueued above is running or queued
                keep_alive = keep_alive[-1:]
                harvest(fresh)
            keep_alive.append(published)
            read_pending = launched
            if launched:
                self._read_enqueue(total)              # behind the launch and the publications enqueued above
        return results, torch.tensor(order, device=self._device)
