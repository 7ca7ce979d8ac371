"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's
GPT semantic-token decoder.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg may import this module; the product path never does.

Parity pin: the reference ships no tests or golden vectors ("parity unpinned" by the
reference itself).  This restatement is pinned instead against outputs of
the reference's own modules run in the build container (``oracle/make_golden.py`` ->
``tests/golden/*.npz``; ``tests/test_oracle_vs_reference.py`` re-checks live whenever
``/root/reference`` is present).

Each function cites the reference lines it restates (paths relative to
``/root/reference/gsv_tts/GPT_SoVITS/GPT/``).  It is written as plain functions over a
state dict -- no ``nn.Module`` -- in one floating-point dtype (fp32 by default, fp64 for
error budgeting).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor
NoiseFn = Callable[[Tuple[int, ...]], Tensor]


def torch_global_exponential(shape) -> Tensor:
    """Exp(1) noise drawn like utils.py:5-9 (``empty_like(probs).exponential_(1)``) from
    torch's global CPU generator, so a seeded run consumes the same stream as the reference."""
    return torch.empty(shape, dtype=torch.float32).exponential_(1)


def sine_table(n_pos: int, dim: int) -> Tensor:
    """embedding.py:52-69: pe[p,2j]=sin(p*w_j), pe[p,2j+1]=cos(p*w_j), w_j=exp(-2j ln(1e4)/dim), fp32."""
    pos = torch.arange(n_pos, dtype=torch.float32).unsqueeze(1)
    w = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * -(math.log(10000.0) / dim))
    pe = torch.zeros(n_pos, dim, dtype=torch.float32)
    pe[:, 0::2] = torch.sin(pos * w)
    pe[:, 1::2] = torch.cos(pos * w)
    return pe


def prompt_mask(nx: int, ny: int) -> Tensor:
    """t2s_model.py:365-381. True = may attend.  Text rows see all text and no audio;
    audio row m sees all text and audio columns <= m."""
    n = nx + ny
    m = torch.zeros(n, n, dtype=torch.bool)
    m[:nx, :nx] = True
    m[nx:, :nx] = True
    m[nx:, nx:] = torch.tril(torch.ones(ny, ny, dtype=torch.bool))
    return m


def layer_norm(x: Tensor, g: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def sample_token(
    logits: Tensor,
    previous_tokens: Optional[Tensor],
    top_k: Optional[int],
    top_p: Optional[float],
    temperature: float,
    repetition_penalty: float,
    noise: NoiseFn,
) -> Tuple[Tensor, Tensor]:
    """utils.py:12-59.  logits [B,V] (modified like the reference: the penalty scatter is in
    place).  Returns (token [B,1] int64, probs [B,V])."""
    if previous_tokens is not None and repetition_penalty != 1.0:
        score = torch.gather(logits, 1, previous_tokens)
        score = torch.where(score < 0, score * repetition_penalty, score / repetition_penalty)
        logits.scatter_(1, previous_tokens, score)
    if top_p is not None and top_p < 1.0:
        srt, idx = torch.sort(logits, descending=True)
        cum = torch.cumsum(torch.softmax(srt, dim=-1), dim=-1)
        drop_sorted = cum > top_p
        drop_sorted[:, 0] = False
        drop = drop_sorted.scatter(1, idx, drop_sorted)
        logits = logits.masked_fill(drop, -float("inf"))
    logits = logits / max(temperature, 1e-5)
    if top_k is not None:
        kth = torch.topk(logits, min(top_k, logits.size(-1)))[0][:, -1:]
        logits = torch.where(logits < kth, torch.full_like(logits, -float("inf")), logits)
    probs = torch.softmax(logits, dim=-1)
    q = noise(tuple(probs.shape)).to(probs.dtype)
    tok = torch.argmax(probs / q, dim=-1, keepdim=True)
    return tok, probs


class GptOracle:
    """Functional restatement of ``Text2SemanticDecoder`` (t2s_model.py:158-734)."""

    SUPPRESS_STEPS = 10      # t2s_model.py:395
    CHECK_INTERVAL = 5       # t2s_model.py:396

    def __init__(self, state_dict: Dict[str, Tensor], config: dict, dtype=torch.float32, n_pos: int = 4000):
        m = config["model"]
        self.d = m["hidden_dim"]
        self.H = m["head"]
        self.dh = self.d // self.H
        self.L = m["n_layer"]
        self.V = m["vocab_size"]
        self.EOS = m["EOS"]
        self.dtype = dtype
        self.w = {k: v.to(dtype) for k, v in state_dict.items()}
        self.suppressed = [280, 486, self.EOS]          # t2s_model.py:170
        pe = sine_table(n_pos, self.d).to(dtype)        # t2s_model.py:212-213
        self.pe_text = self.w["ar_text_position.alpha"] * pe
        self.pe_audio = self.w["ar_audio_position.alpha"] * pe   # "pe_cache", t2s_model.py:409

    # ---- embeddings (t2s_model.py:351-361; embedding.py:71-75) -------------------------
    def embed_text(self, x: Tensor, bert: Tensor) -> Tensor:
        w = self.w
        e = w["ar_text_embedding.word_embeddings.weight"][x]
        e = e + bert.to(self.dtype) @ w["bert_proj.weight"].T + w["bert_proj.bias"]
        return e + self.pe_text[: x.shape[0]]

    def embed_audio(self, y: Tensor) -> Tensor:
        return self.w["ar_audio_embedding.word_embeddings.weight"][y] + self.pe_audio[: y.shape[0]]

    def embed_next(self, tok: Tensor, pos: Tensor) -> Tensor:
        """t2s_model.py:455-456: emb(tok)*x_scale(=1) + (alpha*pe)[kv_len - Nx]."""
        return self.w["ar_audio_embedding.word_embeddings.weight"][tok] + self.pe_audio[pos]

    # ---- one layer ---------------------------------------------------------------------
    def _lw(self, i: int):
        p = f"t2s_transformer.blocks.{i}."
        w = self.w
        return (w[p + "qkv.weight"], w[p + "qkv.bias"], w[p + "out_proj.weight"], w[p + "out_proj.bias"],
                w[p + "mlp.0.weight"], w[p + "mlp.0.bias"], w[p + "mlp.2.weight"], w[p + "mlp.2.bias"],
                w[p + "norm1.weight"], w[p + "norm1.bias"], w[p + "norm2.weight"], w[p + "norm2.bias"])

    def _tail(self, x: Tensor, a: Tensor, lw) -> Tensor:
        """out_proj -> +res -> LN1 -> MLP(ReLU) -> +res -> LN2 (t2s_model.py:55-63 / 95-103)."""
        _, _, wo, bo, w1, b1, w2, b2, g1, be1, g2, be2 = lw
        x = layer_norm(x + a @ wo.T + bo, g1, be1)
        h = torch.relu(x @ w1.T + b1)
        return layer_norm(x + h @ w2.T + b2, g2, be2)

    def layer_prompt(self, i: int, x: Tensor, mask: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        """t2s_model.py:31-65 for one sequence x [n,d]; returns (x', k [n,d], v [n,d])."""
        lw = self._lw(i)
        n = x.shape[0]
        qkv = x @ lw[0].T + lw[1]
        q, k, v = qkv.view(n, 3, self.H, self.dh).unbind(1)          # [n,H,dh]
        s = torch.einsum("qhd,khd->hqk", q, k) / math.sqrt(self.dh)
        s = s.masked_fill(~mask.unsqueeze(0), -float("inf"))
        a = torch.einsum("hqk,khd->qhd", torch.softmax(s, dim=-1), v).reshape(n, self.d)
        return self._tail(x, a, lw), k.reshape(n, self.d), v.reshape(n, self.d)

    def layer_step(self, i: int, x: Tensor, K: Tensor, V: Tensor, kv_len: Tensor) -> Tensor:
        """t2s_model.py:67-105 (flash variant :64-101): x [B,d]; K,V [B,S,d] for layer i are
        updated in place at position kv_len[b]; attention covers 0..kv_len[b] inclusive."""
        lw = self._lw(i)
        B = x.shape[0]
        qkv = x @ lw[0].T + lw[1]
        q, k, v = qkv.view(B, 3, self.d).unbind(1)
        a = torch.empty_like(x)
        for b in range(B):
            n = int(kv_len[b])
            K[b, n] = k[b]
            V[b, n] = v[b]
            kk = K[b, : n + 1].view(n + 1, self.H, self.dh)
            vv = V[b, : n + 1].view(n + 1, self.H, self.dh)
            s = torch.einsum("hd,khd->hk", q[b].view(self.H, self.dh), kk) / math.sqrt(self.dh)
            a[b] = torch.einsum("hk,khd->hd", torch.softmax(s, dim=-1), vv).reshape(self.d)
        return self._tail(x, a, lw)

    # ---- multi-layer drivers -----------------------------------------------------------
    def new_cache(self, slots: int, max_seq: int):
        K = torch.zeros(self.L, slots, max_seq, self.d, dtype=self.dtype)
        return K, torch.zeros_like(K), torch.zeros(slots, dtype=torch.int64)

    def prefill(self, x: Tensor, y: Tensor, bert: Tensor, K: Tensor, V: Tensor, kv_len: Tensor, slot: int = 0):
        """process_single_data + process_prompt (t2s_model.py:351-383, 114-127) for one
        sequence into cache slot ``slot``.  Returns the last position's hidden state [d]."""
        h = torch.cat([self.embed_text(x, bert), self.embed_audio(y)], 0)
        n = h.shape[0]
        mask = prompt_mask(x.shape[0], y.shape[0])
        for i in range(self.L):
            h, k, v = self.layer_prompt(i, h, mask)
            K[i, slot, :n] = k
            V[i, slot, :n] = v
        kv_len[slot] = n
        return h[-1]

    def decode_step(self, x: Tensor, K: Tensor, V: Tensor, kv_len: Tensor) -> Tensor:
        """T2STransformer.decode_next_token (t2s_model.py:129-143): 24 layers then kv_len += 1."""
        for i in range(self.L):
            x = self.layer_step(i, x, K[i], V[i], kv_len)
        kv_len += 1
        return x

    def logits(self, h: Tensor) -> Tensor:
        return h @ self.w["ar_predict_layer.weight"].T           # no bias, t2s_model.py:196

    # ---- single-utterance entry points ---------------------------------------------------
    def _first_token(self, h_last, prev, kw, noise):
        """t2s_model.py:415-417: suppress, drop the EOS column, sample."""
        lg = self.logits(h_last.unsqueeze(0))
        lg[:, self.suppressed] = -float("inf")
        return sample_token(lg[:, :-1], prev, noise=noise, **kw)[0]

    def infer_stream(self, x, y, bert, top_k=15, top_p=1.0, temperature=1.0, repetition_penalty=1.35,
                     stream_chunk=25, boost_first_chunk=True, max_seq=1024,
                     noise: NoiseFn = torch_global_exponential, eos_every_step=True,
                     force_steps: Optional[int] = None) -> Iterator[Tuple[Tensor, bool]]:
        """t2s_model.py:466-553 (and, with eos_every_step=False and one final yield, :385-464).
        x [Nx] int64, y [Ny] int64, bert [Nx,1024].  Yields (tokens [1,1,n], is_final)."""
        kw = dict(top_k=top_k, top_p=top_p, temperature=temperature, repetition_penalty=repetition_penalty)
        nx = x.shape[0]
        K, V, kv_len = self.new_cache(1, max_seq)
        h_last = self.prefill(x, y, bert, K, V, kv_len)
        prev = y.view(1, -1).clone()
        tok = self._first_token(h_last, prev, kw, noise)
        prev = torch.cat([prev, tok], 1)
        xin = self.embed_next(tok[:, 0], kv_len - nx)
        first, pre_chunk, idx = True, None, 0
        n_iter = max_seq - int(kv_len[0])
        if force_steps is not None:
            n_iter = min(n_iter, force_steps)
        for idx in range(1, n_iter + 1):
            h = self.decode_step(xin, K, V, kv_len)
            lg = self.logits(h)
            if idx < self.SUPPRESS_STEPS:
                lg[:, self.suppressed] = -float("inf")
            if force_steps is not None:
                lg[:, self.EOS] = -float("inf")
            tok = sample_token(lg, prev, noise=noise, **kw)[0]
            if eos_every_step:
                if int(tok[0, 0]) == self.EOS:
                    break
                prev = torch.cat([prev, tok], 1)
                if idx % stream_chunk == 0:
                    if pre_chunk is not None:
                        yield pre_chunk, False
                    pre_chunk = prev[:, -idx:].unsqueeze(0)
                    if boost_first_chunk and first:
                        first = False
                        yield pre_chunk, False
                        pre_chunk = None
            else:
                prev = torch.cat([prev, tok], 1)
                if idx % self.CHECK_INTERVAL == 0 and int(tok[0, 0]) == self.EOS:
                    break
            xin = self.embed_next(tok[:, 0], kv_len - nx)
        yield prev[:, -idx:].unsqueeze(0), True

    def infer(self, x, y, bert, **kw) -> Tensor:
        """t2s_model.py:385-464: EOS tested every 5th step only, then the last ``idx`` tokens
        are cut at the first EOS; the very first sampled token is never returned."""
        out = None
        for out, _ in self.infer_stream(x, y, bert, eos_every_step=False, **kw):
            pass
        toks = out[0]
        eos = (toks[0] == self.EOS).nonzero()
        if eos.numel() > 0:
            toks = toks[:, : int(eos[0, 0])]
        return toks.unsqueeze(0)

    # ---- continuous batching -----------------------------------------------------------
    def infer_batched(self, xs: Sequence[Tensor], ys: Sequence[Tensor], berts: Sequence[Tensor],
                      slots: int, max_seq: int, top_k=15, top_p=1.0, temperature=1.0,
                      noise: NoiseFn = torch_global_exponential,
                      max_new: Optional[Sequence[int]] = None) -> Tuple[List[Tensor], List[int]]:
        """Restates the *results contract* of t2s_model.py:555-734 for one bucket length:
        every request is prefilled into a free slot, decoded with no repetition penalty and
        no token suppression (:613,:651), finishes at its first EOS or when the cache is
        full, and its returned tokens exclude the first sampled token and everything from
        the first EOS on (:674-678).  Finished slots are refilled from the queue in order.

        The reference only looks for finished rows every 5 steps and draws noise for idle
        rows too, so its RNG stream depends on that schedule; this restatement takes the
        per-request view (which tokens a request produces given *its own* noise), which is
        what a per-slot counter-based generator on the device reproduces.  ``noise`` is called
        as noise((1,V')) once per sampled token of a request, in request-major order of use.
        ``max_new[r]`` (optional) forces request r to stop after that many returned tokens
        by treating the next token as EOS (bench config 3)."""
        kw = dict(top_k=top_k, top_p=top_p, temperature=temperature, repetition_penalty=1.0)
        results: List[Tensor] = []
        order: List[int] = []
        for r in range(len(xs)):
            K, V, kv_len = self.new_cache(1, max_seq)
            nx = xs[r].shape[0]
            h_last = self.prefill(xs[r], ys[r], berts[r], K, V, kv_len)
            lg = self.logits(h_last.unsqueeze(0))
            tok = sample_token(lg[:, :-1], None, noise=noise, **kw)[0]
            gen: List[int] = []
            while int(kv_len[0]) < max_seq:
                xin = self.embed_next(tok[:, 0], kv_len - nx)
                h = self.decode_step(xin, K, V, kv_len)
                tok = sample_token(self.logits(h), None, noise=noise, **kw)[0]
                t = int(tok[0, 0])
                if t == self.EOS or (max_new is not None and len(gen) >= max_new[r]):
                    break
                gen.append(t)
            results.append(torch.tensor(gen, dtype=torch.int64))
            order.append(r)
        return results, order
