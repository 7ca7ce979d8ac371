// gpt_sample.cuh -- one CTA samples the next token of one slot and prepares the slot's next
// input vector.  Restates sample()/logits_to_probs() (reference GPT/utils.py:12-59) and the
// bookkeeping around it (t2s_model.py:415-420, 442-456, 649-653, 727-728) as device code.
#pragma once
#include "gpt_internal.cuh"

#define GSV_NEG_INF (-__int_as_float(0x7f800000))

struct ArgMax {
  float v;
  int i;
};
__device__ __forceinline__ ArgMax amax2(ArgMax a, ArgMax b) {
  // larger value wins; ties -> smaller index (torch.argmax returns the first maximum)
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ __forceinline__ ArgMax block_argmax(ArgMax a, float* red_v, int* red_i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax b;
    b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
    b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    a = amax2(a, b);
  }
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { red_v[w] = a.v; red_i[w] = a.i; }
  __syncthreads();
  ArgMax r;
  r.v = red_v[0]; r.i = red_i[0];
  for (int k = 1; k < nw; ++k) { ArgMax b; b.v = red_v[k]; b.i = red_i[k]; r = amax2(r, b); }
  return r;
}
__device__ __forceinline__ float block_sum(float a, float* red_v) {
  a = warp_sum(a);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red_v[w] = a;
  __syncthreads();
  float r = 0.f;
  for (int k = 0; k < nw; ++k) r += red_v[k];
  return r;
}

// Shared memory (in floats) needed by sample_slot; carved up at the top of the function.
#define GSV_SAMPLE_SMEM_FLOATS (3 * GSV_VOCAB_MAX + 64)

// Optional flag-in-data ("LL") plumbing used by the small-batch decode kernel (gpt_decode_ll.cu):
// logits were already polled into sm[0..V), the next input and the slot status are published as
// {value, tag} words, and kv_len comes from the caller instead of global memory.
struct SampleLL {
  bool preloaded;
  uint2* xin_ll;        // LL kernels: next input published as {value, tag} words (or null)
  uint2* status_ll;
  unsigned tag;
  int kv_len;
  float* xin_smem;      // cluster kernel: next input [d] and alive flag left in shared memory instead (or null)
  int* alive_smem;
};
// One naturally aligned 64-bit SCALAR access per word: {value (low 32 bits), tag (high 32 bits)}.  A
// vector access (st.v2.u32 / ld.v2.u32) is modelled by the PTX memory model as two scalar accesses in
// unspecified order, so a reader could see the new tag with the old value; a scalar b64 access is
// single-copy atomic.
__device__ __forceinline__ void ll_store(uint2* p, float val, unsigned tag) {
  const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(val);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ uint2 ll_peek(const uint2* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return make_uint2((unsigned)(w & 0xffffffffull), (unsigned)(w >> 32));
}
// Polls that need SEVERAL words per thread: issue every ll_peek first, then test every tag, and loop over the whole
// group (see gpt_decode_hx.cu).  Testing a word right after its load makes ptxas wait for each L2 round trip in turn
// (8 words = 8 x 0.24 us); and the loads must stay STRONG: with weak loads (ld.global.cg) ptxas may assume that a value
// cannot change between two iterations, deletes the loop and the tag test, and the kernel sums whatever was there.
__device__ __forceinline__ float ll_wait(const uint2* p, unsigned tag) {
  uint2 v;
  do { v = ll_peek(p); } while (v.y != tag);
  return __uint_as_float(v.x);
}

// What sample_slot reads from global memory before it can touch a logit: a caller that has to wait for the logits anyway
// (the single-sequence kernel polls them out of L2) issues these loads first and hands them over, so that their L2 round
// trips overlap the wait instead of following it.
struct SamplePre {
  gsv_gpt_sampling sp;
  int ngen;
  unsigned long long cnt;
  unsigned seen[GSV_VOCAB_MAX / GSV_DECODE_THREADS];      // bitmap word of each of this thread's columns (tid + i * 512)
  int x_len;
  int pe_pos;                                             // row of the positional table the next input will use (-1: not fetched)
  float pe;                                               // ... and its element threadIdx.x
};
// kv_len: the slot's cache length AFTER the step being sampled (what SampleLL.kv_len will carry), or -1 if unknown
template <typename T>
__device__ __forceinline__ void sample_prefetch(const GptParams& p, int slot, int kv_len, SamplePre& o) {
  o.x_len = ld_cg(p.x_len + slot);
  o.sp = p.samp[slot];
  o.ngen = ld_cg(p.n_gen + slot);
  o.cnt = __ldcg(p.samp_count + slot);
  const unsigned* seen = p.seen + (size_t)slot * (GSV_VOCAB_MAX / 32);
#pragma unroll
  for (int i = 0; i < GSV_VOCAB_MAX / GSV_DECODE_THREADS; ++i) {
    const int v = threadIdx.x + i * GSV_DECODE_THREADS;
    o.seen[i] = v < p.V ? __ldcg(seen + (v >> 5)) : 0u;
  }
  o.pe_pos = -1;
  o.pe = 0.f;
  if (kv_len >= 0 && (int)threadIdx.x < p.d) {
    o.pe_pos = max(0, min(kv_len - o.x_len, p.n_pos - 1));
    o.pe = Elem<T>::to_f(reinterpret_cast<const T*>(p.pe_audio)[(size_t)o.pe_pos * p.d + threadIdx.x]);
  }
}

// Timeline of the sampler (tuning builds, -DGSV_TIMELINE; tools/hx_timeline.py): {id, globaltimer} records of thread 0 in the
// region behind the per-CTA regions of the kernel's own markers.
__device__ __forceinline__ void mark_sampler(const GptParams& p, int id) {
#ifdef GSV_TIMELINE
  if (p.prof != nullptr && threadIdx.x == 0) {
    long long* rec = p.prof + (size_t)gridDim.x * 2 * p.prof_max;
    const long long n = rec[0];
    if (n + 1 < p.prof_max) {
      unsigned long long gt;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
      rec[2 * (n + 1)] = id;
      rec[2 * (n + 1) + 1] = (long long)gt;
      rec[0] = n + 1;
    }
  }
#else
  (void)p; (void)id;
#endif
}

// Candidates the fast top-k path keeps in shared memory before it falls back to the radix select
#define GSV_SAMPLE_CAND_MAX 256

template <typename T>
__device__ void sample_slot(const GptParams& p, int slot, float* sm, const SampleLL* ll = nullptr, const SamplePre* pre = nullptr) {
  float* lg = sm;                                         // [GSV_VOCAB_MAX] working logits
  int* sidx = reinterpret_cast<int*>(sm + GSV_VOCAB_MAX); // [GSV_VOCAB_MAX] sort indices (top-p only)
  float* kbuf = sm + 2 * GSV_VOCAB_MAX;                   // [GSV_VOCAB_MAX] exp() in sorted order (top-p only)
  float* red_v = sm + 3 * GSV_VOCAB_MAX;                  // [32]
  int* red_i = reinterpret_cast<int*>(red_v + 32);        // [32]
  const int tid = threadIdx.x, NT = blockDim.x;
  const gsv_gpt_sampling sp = pre ? pre->sp : p.samp[slot];
  const int ngen = pre ? pre->ngen : ld_cg(p.n_gen + slot);
  const bool first = ngen == 0;
  const int Vv = first ? p.V - 1 : p.V;                   // first sample: EOS column sliced off (:417)
  const unsigned long long cnt = pre ? pre->cnt : __ldcg(p.samp_count + slot);
  __syncthreads();
  mark_sampler(p, 60);

  // raw logits (+ optional trace of the slot for the teacher-forced / audited parity tests)
  GptSlotHooks* const hk = p.hooks ? p.hooks + slot : nullptr;     // CTA-uniform
  const float* hk_noise = nullptr; const int* hk_forced = nullptr; float* hk_trace = nullptr;
  int hk_noise_rows = 0, hk_n_forced = 0;
  int trow = -1;
  if (hk) {
    hk_noise = hk->noise; hk_noise_rows = hk->noise_rows;
    hk_forced = hk->forced; hk_n_forced = hk->n_forced;
    hk_trace = hk->trace;
    if (hk_trace) {
      trow = ld_cg(&hk->trace_pos);
      if (trow >= hk->trace_max) trow = -1;
    }
  }
  const bool preloaded = ll != nullptr && ll->preloaded;
  const bool suppress = first ? (sp.suppress_first != 0) : (ngen < sp.suppress_steps);
  const float rp = sp.repetition_penalty;
  const bool ext = hk_noise != nullptr && cnt < (unsigned long long)hk_noise_rows;
  const uint2 key = make_uint2((uint32_t)sp.seed, (uint32_t)(sp.seed >> 32));
  int tok = -1;

  // ---- fast path: no top-p, 1 <= top_k <= 32 (the reference's defaults: top_k 15, top_p 1) ------------------------------
  // Same arithmetic per column as the general path below, but nothing is sorted or histogrammed: every thread keeps its
  // columns in registers through the masks, the repetition penalty and the temperature; the 32 half-warps publish their
  // maxima, whose k-th largest L is a lower bound of the top-k pivot (k disjoint groups hold a value >= L); the few
  // columns >= L (about k ln(32/k) + k of them) are compacted into shared memory, the pivot is the candidate with fewer
  // than k candidates above it and at least k at or above it (ties counted, as torch.topk's k-th value), and ONE warp
  // finishes: exp, sum, Exp(1) noise for the survivors only (a column with probability 0 cannot win the arg-max: q > 0),
  // arg-max with ties to the lower index.  5 block barriers instead of ~25 and no contended shared-memory atomics (the
  // top byte of the radix keys -- sign and exponent -- put nearly all 1025 columns into 2-3 bins).
  const int kk = min(sp.top_k, Vv);
  if (sp.top_p >= 1.0f && sp.top_k > 0 && kk <= 32 && NT == GSV_DECODE_THREADS) {
    constexpr int PER = GSV_VOCAB_MAX / GSV_DECODE_THREADS;
    float* cand_v = kbuf;                                 // [GSV_SAMPLE_CAND_MAX]
    int* cand_i = sidx;                                   // [GSV_SAMPLE_CAND_MAX]
    int* n_cand = red_i;                                  // red_i[0]: candidates, red_i[1]: token
    float* pivot_s = red_v + 40;                          // (red_v[0..31]: group maxima; 32 floats past them belong to red_i)
    const int lane = tid & 31, warp = tid >> 5;
    const float temp = fmaxf(sp.temperature, 1e-5f);
    const unsigned* seen = p.seen + (size_t)slot * (GSV_VOCAB_MAX / 32);
    float x[PER];
    float gm = GSV_NEG_INF;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int v = tid + i * GSV_DECODE_THREADS;
      float l = GSV_NEG_INF;
      if (v < p.V) {
        l = preloaded ? lg[v] : ld_cg(p.logits + (size_t)slot * GSV_VOCAB_MAX + v);
        if (trow >= 0) hk_trace[(size_t)trow * p.V + v] = l;
        if (v >= Vv) l = GSV_NEG_INF;
        if (suppress && (v == 280 || v == 486 || v == p.eos)) l = GSV_NEG_INF;
        if (sp.mask_eos && v == p.eos) l = GSV_NEG_INF;
        if (rp != 1.0f && v < Vv) {
          const unsigned word = pre ? pre->seen[i] : __ldcg(seen + (v >> 5));
          if ((word >> (v & 31)) & 1u) l = l < 0.f ? l * rp : l / rp;
        }
        if (temp != 1.0f && v < Vv) l = l / temp;
      }
      x[i] = l;
      lg[v] = l;                                          // the general path's input, should the candidates overflow
      gm = fmaxf(gm, l);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
    if ((lane & 15) == 0) red_v[2 * warp + (lane >> 4)] = gm;
    if (tid == 0) {
      n_cand[0] = 0;
      if (trow >= 0) st_cg(&hk->trace_pos, trow + 1);
    }
    __syncthreads();
    mark_sampler(p, 61);
    // every warp: the k-th largest of the 32 group maxima (lane = group) and the overall maximum
    float L, mx;
    {
      const float g = red_v[lane];
      int gt = 0, ge = 0;
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 o = reinterpret_cast<const float4*>(red_v)[j4];
        gt += (o.x > g) + (o.y > g) + (o.z > g) + (o.w > g);
        ge += (o.x >= g) + (o.y >= g) + (o.z >= g) + (o.w >= g);
      }
      const unsigned hit = __ballot_sync(0xffffffffu, gt < kk && kk <= ge);
      L = __shfl_sync(0xffffffffu, g, __ffs(hit) - 1);
      mx = warp_max(g);
    }
    // compaction of the columns >= L: one shared-memory atomic per warp
    {
      unsigned m[PER];
      int n = 0;
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        m[i] = __ballot_sync(0xffffffffu, tid + i * GSV_DECODE_THREADS < Vv && x[i] >= L);
        n += __popc(m[i]);
      }
      int base = 0;
      if (lane == 0 && n > 0) base = atomicAdd(n_cand, n);
      base = __shfl_sync(0xffffffffu, base, 0);
      const unsigned below = (1u << lane) - 1u;
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        if ((m[i] >> lane) & 1u) {
          const int pos = base + __popc(m[i] & below);
          if (pos < GSV_SAMPLE_CAND_MAX) { cand_v[pos] = x[i]; cand_i[pos] = tid + i * GSV_DECODE_THREADS; }
        }
        base += __popc(m[i]);
      }
    }
    __syncthreads();
    mark_sampler(p, 62);
    const int C = n_cand[0];
    if (C <= GSV_SAMPLE_CAND_MAX && L > GSV_NEG_INF) {
      // pivot: candidate c with  #(candidates > c) < k <= #(candidates >= c)
      if (tid < C) {
        const float c = cand_v[tid];
        int gt = 0, ge = 0;
        for (int j = 0; j < C; ++j) {
          const float o = cand_v[j];
          gt += o > c;
          ge += o >= c;
        }
        if (gt < kk && kk <= ge) *pivot_s = c;            // every hit holds the same value
      }
      __syncthreads();
      mark_sampler(p, 63);
      if (warp == 0) {
        const float pivot = *pivot_s;
        float part = 0.f;
        for (int c = lane; c < C; c += 32) {
          const float xv = cand_v[c];
          if (xv >= pivot) part += __expf(xv - mx);
        }
        const float total = warp_sum(part);
        const float inv_total = 1.0f / total;
        ArgMax best;
        best.v = -1.0f; best.i = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {
          const float xv = cand_v[c];
          if (xv >= pivot) {
            const int v = cand_i[c];
            float q;
            if (ext) {
              q = hk_noise[(size_t)cnt * p.V + v];
            } else {
              uint4 r = philox4x32_10(make_uint4((uint32_t)cnt, (uint32_t)(cnt >> 32), (uint32_t)(v >> 2), 0u), key);
              uint32_t bits = (v & 3) == 0 ? r.x : (v & 3) == 1 ? r.y : (v & 3) == 2 ? r.z : r.w;
              q = exp1_from_bits(bits);
            }
            ArgMax b;
            b.v = (__expf(xv - mx) * inv_total) / q;
            b.i = v;
            best = amax2(best, b);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ArgMax b;
          b.v = __shfl_xor_sync(0xffffffffu, best.v, o);
          b.i = __shfl_xor_sync(0xffffffffu, best.i, o);
          best = amax2(best, b);
        }
        if (lane == 0) n_cand[1] = best.i;
      }
      __syncthreads();
      mark_sampler(p, 64);
      tok = n_cand[1];
    } else {
      tok = -2;                                           // overflow (or fewer than k finite columns): general path from the top-k on
    }
  }

  if (tok == -1) {
  for (int v = tid; v < GSV_VOCAB_MAX; v += NT) {
    float l = GSV_NEG_INF;
    if (v < p.V) {
      l = preloaded ? lg[v] : ld_cg(p.logits + (size_t)slot * GSV_VOCAB_MAX + v);
      if (trow >= 0) hk_trace[(size_t)trow * p.V + v] = l;
      if (v >= Vv) l = GSV_NEG_INF;
    }
    lg[v] = l;
  }
  __syncthreads();
  // the first sample of infer / infer_stream masks the suppressed tokens whatever initial_suppression_steps is
  // (t2s_model.py:415); infer_batched never does (:613)
  if (tid == 0) {
    if (trow >= 0) st_cg(&hk->trace_pos, trow + 1);
    if (suppress) {                                       // suppressed_tokens = [280, 486, EOS] (:170)
      if (280 < Vv) lg[280] = GSV_NEG_INF;
      if (486 < Vv) lg[486] = GSV_NEG_INF;
      if (p.eos < Vv) lg[p.eos] = GSV_NEG_INF;
    }
    if (sp.mask_eos && p.eos < Vv) lg[p.eos] = GSV_NEG_INF;
  }
  __syncthreads();

  // (1) repetition penalty over the set of previous tokens (utils.py:20-27; duplicates in
  //     previous_tokens gather the same original score, so a set is equivalent)
  if (rp != 1.0f) {
    const unsigned* seen = p.seen + (size_t)slot * (GSV_VOCAB_MAX / 32);
    for (int v = tid; v < Vv; v += NT) {
      if ((__ldcg(seen + (v >> 5)) >> (v & 31)) & 1u) {
        float l = lg[v];
        lg[v] = l < 0.f ? l * rp : l / rp;
      }
    }
    __syncthreads();
  }

  // (2) top-p (utils.py:29-39): sort descending, drop sorted entries whose inclusive cumulative
  //     probability exceeds top_p, except rank 0
  if (sp.top_p < 1.0f) {
    for (int v = tid; v < GSV_VOCAB_MAX; v += NT) sidx[v] = v;
    __syncthreads();
    // bitonic sort of the index array, keys looked up through lg[]; descending, ties by index
    for (int k = 2; k <= GSV_VOCAB_MAX; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = tid; t < GSV_VOCAB_MAX; t += NT) {
          int u = t ^ j;
          if (u > t) {
            int ia = sidx[t], ib = sidx[u];
            float a = lg[ia], b = lg[ib];
            bool desc = (t & k) == 0;
            bool a_before_b = (a > b) || (a == b && ia < ib);
            if (desc ? !a_before_b : a_before_b) { sidx[t] = ib; sidx[u] = ia; }
          }
        }
        __syncthreads();
      }
    }
    const float top = lg[sidx[0]];
    float part = 0.f;
    for (int t = tid; t < GSV_VOCAB_MAX; t += NT) {
      float e = __expf(lg[sidx[t]] - top);
      kbuf[t] = e;
      part += e;
    }
    const float total = block_sum(part, red_v);
    __syncthreads();
    // inclusive scan in sorted order: each thread owns a contiguous run of `per` entries
    const int per = GSV_VOCAB_MAX / NT;
    float run = 0.f;
    for (int i = 0; i < per; ++i) run += kbuf[tid * per + i];
    float incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float n = __shfl_up_sync(0xffffffffu, incl, o);
      if ((tid & 31) >= o) incl += n;
    }
    __syncthreads();
    if ((tid & 31) == 31) red_v[tid >> 5] = incl;
    __syncthreads();
    float base = 0.f;
    for (int w = 0; w < (tid >> 5); ++w) base += red_v[w];
    float cum = base + incl - run;
    for (int i = 0; i < per; ++i) {
      int t = tid * per + i;
      cum += kbuf[t];
      if (t > 0 && cum / total > sp.top_p) lg[sidx[t]] = GSV_NEG_INF;   // distinct targets: no race
    }
    __syncthreads();
  }

  // (3) temperature (utils.py:41)
  {
    const float t = fmaxf(sp.temperature, 1e-5f);
    if (t != 1.0f)
      for (int v = tid; v < Vv; v += NT) lg[v] = lg[v] / t;
    __syncthreads();
  }
  }   // tok == -1: the general path's steps before the top-k

  if (tok < 0) {
  // (4) top-k pivot = k-th largest value with multiplicity (utils.py:43-46); ties with the pivot survive.
  //     Exact radix select on an order-preserving integer image of the floats: 4 passes of 8 bits, each a
  //     256-bin block histogram (instead of k rounds of block arg-max).
  if (sp.top_k > 0) {
    int k = min(sp.top_k, Vv);
    constexpr int PER = GSV_VOCAB_MAX / GSV_DECODE_THREADS;   // blockDim.x == GSV_DECODE_THREADS
    unsigned key[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int v = tid + i * NT;
      const unsigned u = __float_as_uint(v < Vv ? lg[v] : GSV_NEG_INF);
      key[i] = (u & 0x80000000u) ? ~u : (u | 0x80000000u);    // larger float <=> larger key
    }
    unsigned* hist = reinterpret_cast<unsigned*>(kbuf);       // [256] (kbuf is only used by top-p, which is done)
    unsigned* sel = hist + 256;                               // [2]: chosen bin, remaining rank
    unsigned prefix = 0u, pmask = 0u;
#pragma unroll 1
    for (int shift = 24; shift >= 0; shift -= 8) {
      if (tid < 256) hist[tid] = 0u;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < PER; ++i)
        if (tid + i * NT < Vv && (key[i] & pmask) == prefix) atomicAdd(&hist[(key[i] >> shift) & 255u], 1u);
      __syncthreads();
      if (tid < 32) {
        // lane owns bins [8*lane, 8*lane+8); walk from the top bin down until the rank falls inside a bin
        unsigned c[8], tot = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) { c[j] = hist[tid * 8 + j]; tot += c[j]; }
        // inclusive suffix sum of tot over lanes: entries in this lane's bins and every higher bin
        unsigned suf = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned n = __shfl_down_sync(0xffffffffu, suf, o);
          if (tid + o < 32) suf += n;
        }
        const unsigned above = suf - tot;                     // entries in bins above this lane's
        if ((unsigned)k > above && (unsigned)k <= above + tot) {
          unsigned run = above;
#pragma unroll
          for (int j = 7; j >= 0; --j) {
            if ((unsigned)k > run && (unsigned)k <= run + c[j]) { sel[0] = (unsigned)(tid * 8 + j); sel[1] = (unsigned)k - run; }
            run += c[j];
          }
        }
      }
      __syncthreads();
      prefix |= sel[0] << shift;
      pmask |= 255u << shift;
      k = (int)sel[1];
    }
    const unsigned pk = prefix;                               // key of the k-th largest value
    const float pivot = __uint_as_float((pk & 0x80000000u) ? (pk & 0x7fffffffu) : ~pk);
    for (int v = tid; v < Vv; v += NT)
      if (lg[v] < pivot) lg[v] = GSV_NEG_INF;
    __syncthreads();
  }

  // (5) softmax (utils.py:48)
  float mx = GSV_NEG_INF;
  for (int v = tid; v < Vv; v += NT) mx = fmaxf(mx, lg[v]);
  mx = warp_max(mx);
  __syncthreads();
  if ((tid & 31) == 0) red_v[tid >> 5] = mx;
  __syncthreads();
  mx = red_v[0];
  for (int w = 1; w < (NT >> 5); ++w) mx = fmaxf(mx, red_v[w]);
  float part = 0.f;
  for (int v = tid; v < Vv; v += NT) {
    float e = __expf(lg[v] - mx);
    lg[v] = e;
    part += e;
  }
  const float total = block_sum(part, red_v);
  const float inv_total = 1.0f / total;

  // (6) token = argmax(p / q), q ~ Exp(1) i.i.d. per column (utils.py:5-9)
  ArgMax best;
  best.v = -1.0f; best.i = 0x7fffffff;
  for (int v = tid; v < Vv; v += NT) {
    float q;
    if (ext) {
      q = hk_noise[(size_t)cnt * p.V + v];
    } else {
      uint4 r = philox4x32_10(make_uint4((uint32_t)cnt, (uint32_t)(cnt >> 32), (uint32_t)(v >> 2), 0u), key);   // slot-independent: the seed names the request
      uint32_t bits = (v & 3) == 0 ? r.x : (v & 3) == 1 ? r.y : (v & 3) == 2 ? r.z : r.w;
      q = exp1_from_bits(bits);
    }
    ArgMax b;
    b.v = (lg[v] * inv_total) / q;
    b.i = v;
    best = amax2(best, b);
  }
  best = block_argmax(best, red_v, red_i);
  tok = best.i;
  }   // tok < 0: general top-k / softmax / arg-max

  // bookkeeping by one thread; every value other CTAs read later goes through st.cg
  const int kvl = ll ? ll->kv_len : ld_cg(p.kv_len + slot);
  if (sp.max_new_tokens > 0 && ngen > sp.max_new_tokens) tok = p.eos;    // ngen counts s0 too
  if (hk_forced != nullptr) {                    // CTA-uniform branch
    int fp = ld_cg(&hk->forced_pos);
    __syncthreads();                              // every thread has read the cursor before thread 0 moves it
    if (fp < hk_n_forced) {
      tok = hk_forced[fp];
      if (tid == 0) st_cg(&hk->forced_pos, fp + 1);
    }
  }
  const int kv_cap = (sp.max_kv > 0 && sp.max_kv < p.S) ? sp.max_kv : p.S;
  const bool stop = (tok == p.eos) || (kvl >= kv_cap);
  if (tid == 0) {
    st_cg(p.tokens + (size_t)slot * p.S + ngen, tok);
    st_cg(p.n_gen + slot, ngen + 1);
    __stcg(p.samp_count + slot, cnt + 1);
    if (tok < GSV_VOCAB_MAX) atomicOr(p.seen + (size_t)slot * (GSV_VOCAB_MAX / 32) + (tok >> 5), 1u << (tok & 31));
    if (stop) st_cg(p.active + slot, 0);
    mark_sampler(p, 65);
    if (ll) {
      st_cg(p.kv_len + slot, kvl);
      if (ll->status_ll) ll_store(ll->status_ll, stop ? 0.f : 1.f, ll->tag);
      if (ll->alive_smem) *ll->alive_smem = stop ? 0 : 1;
    }
  }
  // next input: emb_audio[tok] * x_scale(=1) + (alpha*pe)[kv_len - Nx]  (:455-456, :727-728)
  if (!stop) {
    const T* emb = reinterpret_cast<const T*>(p.emb_audio) + (size_t)tok * p.d;
    int pos = kvl - (pre ? pre->x_len : ld_cg(p.x_len + slot));
    pos = max(0, min(pos, p.n_pos - 1));
    const T* pe = reinterpret_cast<const T*>(p.pe_audio) + (size_t)pos * p.d;
    for (int c = tid; c < p.d; c += NT) {
      const float pe_c = (pre && pre->pe_pos == pos && c == tid) ? pre->pe : Elem<T>::to_f(pe[c]);
      // the reference adds two T values and rounds to T
      float s = Elem<T>::to_f(Elem<T>::from_f(Elem<T>::to_f(emb[c]) + pe_c));
      st_cg(p.xin + (size_t)slot * p.d + c, s);
      if (ll && ll->xin_ll) ll_store(ll->xin_ll + c, s, ll->tag);
      if (ll && ll->xin_smem) ll->xin_smem[c] = s;
    }
  }
  mark_sampler(p, 66);
  __syncthreads();
}
