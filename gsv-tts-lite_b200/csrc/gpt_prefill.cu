// gpt_prefill.cu -- prompt prefill of one or several requests into their KV-cache slots, and their first tokens.
//
// Replaces process_single_data + T2STransformer.process_prompt + the first sample()
// (reference gsv_tts/GPT_SoVITS/GPT/t2s_model.py:351-383, 31-65, 114-127, 414-420; the same
// sequence is the slot-refill path of infer_batched, :696-722, which the reference runs one request at a time).
//
// Several prompts share one pass: their rows are stacked into one [sum n_i][d] matrix, so the four linears of a layer
// are ONE tcgen05 launch each over all prompts (rows of the same 128-row tiles) and the per-row kernels (masked attention,
// which also files K/V in the cache, add + LayerNorm) find their prompt through a small segment table passed by value.  The
// launch count of a pass does not depend on the number of prompts (7 per layer).
#include "gpt_sample.cuh"

namespace {

// prompts of one pass: rows [row0[i], row0[i+1]) of the stacked matrices belong to prompt i
struct PfSegs {
  int n;
  int row0[GSV_MAX_SLOTS + 1];
  int nx[GSV_MAX_SLOTS];
  int slot[GSV_MAX_SLOTS];
};
__device__ __forceinline__ int pf_seg_of(const PfSegs& s, int row) {
  int i = 0;
  while (i + 1 < s.n && row >= s.row0[i + 1]) ++i;
  return i;
}

// ---- embeddings (A.2) ------------------------------------------------------------------------
// text rows:  round(round(emb_text[x] + bert_proj) + alpha_t*pe[i]);  audio rows: round(emb_audio[y] + alpha_a*pe[m])
template <typename T>
__global__ void embed_kernel(const GptParams p, const int64_t* __restrict__ x, int nx, const int64_t* __restrict__ y, int ny,
                             const T* __restrict__ bert_proj /*[nx][d]*/, T* __restrict__ out /*[nx+ny][d]*/) {
  const int row = blockIdx.x, d = p.d;
  const T* pe;
  const T* emb;
  if (row < nx) {
    long long id = x[row];
    id = id < 0 ? 0 : (id >= p.n_phoneme ? p.n_phoneme - 1 : id);
    emb = reinterpret_cast<const T*>(p.emb_text) + (size_t)id * d;
    pe = reinterpret_cast<const T*>(p.pe_text) + (size_t)row * d;
  } else {
    long long id = y[row - nx];
    id = id < 0 ? 0 : (id >= p.V ? p.V - 1 : id);
    emb = reinterpret_cast<const T*>(p.emb_audio) + (size_t)id * d;
    pe = reinterpret_cast<const T*>(p.pe_audio) + (size_t)(row - nx) * d;
  }
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float e = Elem<T>::to_f(emb[c]);
    if (row < nx) e = Elem<T>::to_f(Elem<T>::from_f(e + Elem<T>::to_f(bert_proj[(size_t)row * d + c])));
    out[(size_t)row * d + c] = Elem<T>::from_f(e + Elem<T>::to_f(pe[c]));
  }
}

// ---- C[M][N] = A[M][K] W[N][K]^T + bias  (nn.Linear), optional ReLU --------------------------------
template <typename T, bool RELU>
__global__ void __launch_bounds__(256) gemm_tn_kernel(const T* __restrict__ A, int lda, const T* __restrict__ W,
                                                      const T* __restrict__ bias, T* __restrict__ C, int ldc, int M, int N,
                                                      int K) {
  constexpr int BM = 64, BN = 64, BK = 32;
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lr = tid >> 2, lk = (tid & 3) * 8;
  for (int k0 = 0; k0 < K; k0 += BK) {
    float a8[8], w8[8];
    if (m0 + lr < M) unpack8<T>(*reinterpret_cast<const uint4*>(A + (size_t)(m0 + lr) * lda + k0 + lk), a8);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) a8[j] = 0.f;
    }
    if (n0 + lr < N) unpack8<T>(ld_weight(W + (size_t)(n0 + lr) * K + k0 + lk), w8);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) w8[j] = 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) { As[lk + j][lr] = a8[j]; Ws[lk + j][lr] = w8[j]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? Elem<T>::to_f(bias[n]) : 0.f);
      if (RELU) v = fmaxf(v, 0.f);
      C[(size_t)m * ldc + n] = Elem<T>::from_f(v);
    }
  }
}

// ---- masked prompt attention (A.2 mask): text row -> all text; audio row i -> keys 0..i -----------------
// one warp per (query row, head); 4 lanes x 8 dims per key, 8 keys per pass.  The warp also files its row's K and V of this
// head in the slot's cache, kc[l][slot][h][t][32] (process_prompt writes K/V [:, :, :L], t2s_model.py:44-47): no separate
// scatter launch.
template <typename T>
__global__ void __launch_bounds__(128) prefill_attn_kernel(const GptParams p, int layer, const T* __restrict__ qkv_all,
                                                           T* __restrict__ out_all, const PfSegs segs, int n_tot, int d, int H) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 4 + warp;
  if (item >= n_tot * H) return;
  const int gi = item / H, h = item - gi * H;
  const int sg = pf_seg_of(segs, gi);
  const int base = segs.row0[sg], nx = segs.nx[sg];
  const int i = gi - base;                                      // row within its prompt; keys are rows of the same prompt only
  const T* qkv = qkv_all + (size_t)base * 3 * d;
  T* out = out_all + (size_t)base * d;
  const int lim = i < nx ? nx : i + 1;
  if (lane < 8) {                                               // 4 lanes x 16 bytes = the 32 K values of (row, head); 4 more for V
    const int which = lane >> 2, part = lane & 3;
    const size_t a = (((size_t)(layer * p.slots + segs.slot[sg]) * H + h) * p.S + i) * GSV_HEAD_DIM + part * 8;
    const uint4 v = *reinterpret_cast<const uint4*>(qkv + (size_t)i * 3 * d + (1 + which) * d + h * GSV_HEAD_DIM + part * 8);
    *reinterpret_cast<uint4*>(reinterpret_cast<T*>(which ? p.vc : p.kc) + a) = v;
  }
  const int sub = lane & 3, pg = lane >> 2;
  const float qscale = rsqrtf((float)GSV_HEAD_DIM) * 1.4426950408889634f;
  float q[8];
  unpack8<T>(*reinterpret_cast<const uint4*>(qkv + (size_t)i * 3 * d + h * GSV_HEAD_DIM + sub * 8), q);
#pragma unroll
  for (int j = 0; j < 8; ++j) q[j] *= qscale;
  float m = GSV_NEG_INF, l = 0.f, o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = 0.f;
  for (int tt = 0; tt < lim; tt += 8) {
    const int t = tt + pg;
    const bool ok = t < lim;
    float s = GSV_NEG_INF, vf[8];
    if (ok) {
      float kf[8];
      const T* row = qkv + (size_t)t * 3 * d + h * GSV_HEAD_DIM + sub * 8;
      unpack8<T>(*reinterpret_cast<const uint4*>(row + d), kf);
      unpack8<T>(*reinterpret_cast<const uint4*>(row + 2 * d), vf);
      s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s = fmaf(q[j], kf[j], s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (ok) {
      const float mn = fmaxf(m, s);
      const float sc = exp2f(m - mn), pr = exp2f(s - mn);
      l = l * sc + pr;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(pr, vf[j], o[j] * sc);
      m = mn;
    }
  }
#pragma unroll
  for (int off = 4; off < 32; off <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, off);
    const float l2 = __shfl_xor_sync(0xffffffffu, l, off);
    const float mn = fmaxf(m, m2);
    const float a = mn > GSV_NEG_INF ? exp2f(m - mn) : 0.f;
    const float b = mn > GSV_NEG_INF ? exp2f(m2 - mn) : 0.f;
    l = l * a + l2 * b;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = o[j] * a + __shfl_xor_sync(0xffffffffu, o[j], off) * b;
    m = mn;
  }
  if (pg == 0) {
    float r[8];
    const float inv = 1.f / l;
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = o[j] * inv;
    *reinterpret_cast<uint4*>(out + (size_t)i * d + h * GSV_HEAD_DIM + sub * 8) = pack8<T>(r);
  }
}

// ---- x = LayerNorm(round(res + y)) (post-LN, t2s_model.py:57-58, 62-63); warp per row, in place on x ---------------
template <typename T>
__global__ void __launch_bounds__(128) add_ln_kernel(T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ g,
                                                     const T* __restrict__ b, int n, int d) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // a tensor-core linear follows: let it set up early
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp;
  if (row >= n) return;
  float v[32];
  const int per = d >> 5;     // d <= 1024
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < per) {
      const int c = i * 32 + lane;
      v[i] = Elem<T>::to_f(Elem<T>::from_f(Elem<T>::to_f(x[(size_t)row * d + c]) + Elem<T>::to_f(y[(size_t)row * d + c])));
      sum += v[i];
    }
  const float mean = warp_sum(sum) / (float)d;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < per) sq += (v[i] - mean) * (v[i] - mean);
  const float rstd = rsqrtf(warp_sum(sq) / (float)d + 1e-5f);
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < per) {
      const int c = i * 32 + lane;
      x[(size_t)row * d + c] = Elem<T>::from_f((v[i] - mean) * rstd * Elem<T>::to_f(g[c]) + Elem<T>::to_f(b[c]));
    }
}

// ---- logits of the last prompt row (fp32) + slot state reset ----------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) head_row_kernel(const GptParams p, int slot, const T* __restrict__ xrow) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= p.V) return;
  const T* w = reinterpret_cast<const T*>(p.w_head) + (size_t)row * p.d;
  float acc = 0.f;
  for (int c = lane * 8; c < p.d; c += 256) {
    float wf[8], xf[8];
    unpack8<T>(ld_weight(w + c), wf);
    unpack8<T>(*reinterpret_cast<const uint4*>(xrow + c), xf);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(wf[j], xf[j], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) p.logits[(size_t)slot * GSV_VOCAB_MAX + row] = acc;
}

__global__ void slot_reset_kernel(const GptParams p, int slot, int nx, int n, const int64_t* __restrict__ y, int ny,
                                  gsv_gpt_sampling samp) {
  unsigned* seen = p.seen + (size_t)slot * (GSV_VOCAB_MAX / 32);
  for (int i = threadIdx.x; i < GSV_VOCAB_MAX / 32; i += blockDim.x) seen[i] = 0u;
  __syncthreads();
  // previous_tokens starts as the prompt tokens y (t2s_model.py:412)
  for (int i = threadIdx.x; i < ny; i += blockDim.x) {
    long long t = y[i];
    if (t >= 0 && t < GSV_VOCAB_MAX) atomicOr(seen + (t >> 5), 1u << (t & 31));
  }
  if (threadIdx.x == 0) {
    p.kv_len[slot] = n;
    p.x_len[slot] = nx;
    p.n_gen[slot] = 0;
    p.samp_count[slot] = 0ull;
    p.active[slot] = 1;
    p.samp[slot] = samp;
    if (p.hooks) { p.hooks[slot].forced_pos = 0; p.hooks[slot].trace_pos = 0; }
  }
}

template <typename T>
__global__ void __launch_bounds__(GSV_DECODE_THREADS, 1) first_token_kernel(const GptParams p, int slot) {
  extern __shared__ __align__(16) float sm[];
  sample_slot<T>(p, slot, sm);
}

template <typename T>
int gemm(const T* A, int lda, const T* W, const T* bias, T* C, int ldc, int M, int N, int K, bool relu, cudaStream_t st,
         long long& launches, gsv_gpt_ctx* ctx = nullptr, size_t op = 0, int rows_cap = 0) {
  // tensor-core path: the tcgen05 implicit-GEMM kernel with one tap (conv_umma.cuh); tensor maps cached per call site
  if (ctx && ctx->use_umma_linear && ctx->umma && bias && lda == K && ldc == N && K >= 64 && K % 8 == 0 && N % 16 == 0) {
    launches += 1;
    return gsv_umma_linear(ctx->umma, op, ctx->dims.dtype, A, M, rows_cap, K, W, bias, N, C, relu ? 1 : 0, st);
  }
  if (K % 32 != 0) { gsv_set_error("gemm: K=%d must be a multiple of 32", K); return GSV_ERR_ARG; }
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  if (relu) gemm_tn_kernel<T, true><<<grid, 256, 0, st>>>(A, lda, W, bias, C, ldc, M, N, K);
  else gemm_tn_kernel<T, false><<<grid, 256, 0, st>>>(A, lda, W, bias, C, ldc, M, N, K);
  launches += 1;
  GSV_CHECK_LAUNCH();
  return GSV_OK;
}

// Body of a prefill pass: everything that does not touch the slots' decode state (embeddings, the L layers, the K/V rows of
// the slots).  It may run on another stream than the decode launches: decode kernels never look at an inactive slot.
// The last row of every prompt is parked in pf_last[slot] for prefill_tail().
template <typename T>
int prefill_body(gsv_gpt_ctx* ctx, int n_prompts, const int* slots, const int64_t* const* xs, const int* nxs, const int64_t* const* ys,
                 const int* nys, const void* const* berts, cudaStream_t st) {
  const GptParams& p = ctx->p;
  const int d = p.d, F = p.F, L = p.L;
  T* X = reinterpret_cast<T*>(ctx->pf_x);        // [n][d]
  T* QKV = reinterpret_cast<T*>(ctx->pf_qkv);    // [n][3d]
  T* ATT = reinterpret_cast<T*>(ctx->pf_attn);   // [n][d]
  T* Hh = reinterpret_cast<T*>(ctx->pf_h);       // [n][F]
  T* TMP = reinterpret_cast<T*>(ctx->pf_tmp);    // [n][d]
  PfSegs segs;
  memset(&segs, 0, sizeof(segs));
  segs.n = n_prompts;
  int n = 0;
  for (int i = 0; i < n_prompts; ++i) {
    segs.row0[i] = n; segs.nx[i] = nxs[i]; segs.slot[i] = slots[i];
    n += nxs[i] + nys[i];
  }
  segs.row0[n_prompts] = n;
  const int cap = ctx->pf_rows;
  int rc;
  for (int i = 0; i < n_prompts; ++i) {
    const int r0 = segs.row0[i];
    // bert_proj (nn.Linear(1024, d), t2s_model.py:172, 354); bert rows are caller-owned: CUDA cores
    if ((rc = gemm<T>(reinterpret_cast<const T*>(berts[i]), p.d_bert, reinterpret_cast<const T*>(p.w_bert),
                      reinterpret_cast<const T*>(p.b_bert), TMP + (size_t)r0 * d, d, nxs[i], d, p.d_bert, false, st, ctx->launches)))
      return rc;
    embed_kernel<T><<<nxs[i] + nys[i], 128, 0, st>>>(p, xs[i], nxs[i], ys[i], nys[i], TMP + (size_t)r0 * d, X + (size_t)r0 * d);
    ctx->launches += 1;
    GSV_CHECK_LAUNCH();
  }
  for (int l = 0; l < L; ++l) {
    const T* wqkv = reinterpret_cast<const T*>(p.w_qkv) + (size_t)l * 3 * d * d;
    const T* bqkv = reinterpret_cast<const T*>(p.b_qkv) + (size_t)l * 3 * d;
    if ((rc = gemm<T>(X, d, wqkv, bqkv, QKV, 3 * d, n, 3 * d, d, false, st, ctx->launches, ctx, (size_t)l * 4 + 0, cap))) return rc;
    prefill_attn_kernel<T><<<(n * p.H + 3) / 4, 128, 0, st>>>(p, l, QKV, ATT, segs, n, d, p.H);
    ctx->launches += 1;
    GSV_CHECK_LAUNCH();
    if ((rc = gemm<T>(ATT, d, reinterpret_cast<const T*>(p.w_o) + (size_t)l * d * d,
                      reinterpret_cast<const T*>(p.b_o) + (size_t)l * d, TMP, d, n, d, d, false, st, ctx->launches, ctx,
                      (size_t)l * 4 + 1, cap)))
      return rc;
    add_ln_kernel<T><<<(n + 3) / 4, 128, 0, st>>>(X, TMP, reinterpret_cast<const T*>(p.ln1_g) + (size_t)l * d,
                                                 reinterpret_cast<const T*>(p.ln1_b) + (size_t)l * d, n, d);
    ctx->launches += 1;
    GSV_CHECK_LAUNCH();
    if ((rc = gemm<T>(X, d, reinterpret_cast<const T*>(p.w_1) + (size_t)l * F * d,
                      reinterpret_cast<const T*>(p.b_1) + (size_t)l * F, Hh, F, n, F, d, true, st, ctx->launches, ctx,
                      (size_t)l * 4 + 2, cap)))
      return rc;
    if ((rc = gemm<T>(Hh, F, reinterpret_cast<const T*>(p.w_2) + (size_t)l * d * F,
                      reinterpret_cast<const T*>(p.b_2) + (size_t)l * d, TMP, d, n, d, F, false, st, ctx->launches, ctx,
                      (size_t)l * 4 + 3, cap)))
      return rc;
    add_ln_kernel<T><<<(n + 3) / 4, 128, 0, st>>>(X, TMP, reinterpret_cast<const T*>(p.ln2_g) + (size_t)l * d,
                                                 reinterpret_cast<const T*>(p.ln2_b) + (size_t)l * d, n, d);
    ctx->launches += 1;
    GSV_CHECK_LAUNCH();
  }
  for (int i = 0; i < n_prompts; ++i) {
    GSV_CUDA(cudaMemcpyAsync(reinterpret_cast<T*>(ctx->pf_last) + (size_t)slots[i] * d, X + (size_t)(segs.row0[i + 1] - 1) * d,
                             (size_t)d * sizeof(T), cudaMemcpyDeviceToDevice, st));
    ctx->pf_nx[slots[i]] = nxs[i];
    ctx->pf_n[slots[i]] = nxs[i] + nys[i];
  }
  return GSV_OK;
}

// Tail of a prefill: slot state reset (this is what makes the slot live), logits of the last prompt row, first sample.
// Must be ordered after the body of the same slot and must not overlap a decode launch (same stream as the decodes).
template <typename T>
int prefill_tail(gsv_gpt_ctx* ctx, int slot, const int64_t* y, int ny, const gsv_gpt_sampling& samp, cudaStream_t st) {
  const GptParams& p = ctx->p;
  slot_reset_kernel<<<1, 256, 0, st>>>(p, slot, ctx->pf_nx[slot], ctx->pf_n[slot], y, ny, samp);
  head_row_kernel<T><<<(p.V + 7) / 8, 256, 0, st>>>(p, slot, reinterpret_cast<const T*>(ctx->pf_last) + (size_t)slot * p.d);
  const size_t smem = GSV_SAMPLE_SMEM_FLOATS * sizeof(float);
  first_token_kernel<T><<<1, GSV_DECODE_THREADS, smem, st>>>(p, slot);
  ctx->launches += 3;
  GSV_CHECK_LAUNCH();
  return GSV_OK;
}


// =====================================================================================================================
// Batched decode step as a short sequence of kernels (5..32 live sequences; GSV_DECODE_IMPL=gemm forces it):
// the four linears of every layer run on the tcgen05 kernel with the live sequences as rows (M tile 128, rows past
// the slot count are zero-filled by TMA), attention / add+LayerNorm / head / sampling are small kernels over
// (slot, head) or slot.  One step = 7 L + 2 launches, captured once into a CUDA graph and replayed per token.
// Same arithmetic as T2STransformer.decode_next_token + sample (t2s_model.py:129-143, 637-653) with activations
// rounded to the storage type between ops, as the reference rounds them.
// =====================================================================================================================

// x rows of the step: Xd[slot] = T(xin[slot]) (xin is written by prefill / the sampler as fp32 values already rounded to T)
template <typename T>
__global__ void xin_to_x_kernel(const GptParams p, T* __restrict__ X) {
  const int slot = blockIdx.x;
  if (!p.active[slot]) return;
  for (int c = threadIdx.x; c < p.d; c += blockDim.x) X[(size_t)slot * p.d + c] = Elem<T>::from_f(p.xin[(size_t)slot * p.d + c]);
}

// KV append + attention of one (head, slot) over positions 0..kv_len inclusive; 4 warps, 8 positions per warp pass
template <typename T>
__global__ void __launch_bounds__(128) dec_attn_kernel(const GptParams p, int layer, const T* __restrict__ qkv, T* __restrict__ out) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int h = blockIdx.x, slot = blockIdx.y;
  if (!p.active[slot]) return;
  __shared__ float part[4][GSV_HEAD_DIM + 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = p.d, kv = p.kv_len[slot];
  const int sub = lane & 3, pg = lane >> 2;
  const size_t head_base = ((size_t)(layer * p.slots + slot) * p.H + h) * (size_t)p.S * GSV_HEAD_DIM;
  T* kc = reinterpret_cast<T*>(p.kc) + head_base;
  T* vc = reinterpret_cast<T*>(p.vc) + head_base;
  const T* row = qkv + (size_t)slot * 3 * d + h * GSV_HEAD_DIM;
  if (threadIdx.x < GSV_HEAD_DIM) {
    kc[(size_t)kv * GSV_HEAD_DIM + threadIdx.x] = row[d + threadIdx.x];
    vc[(size_t)kv * GSV_HEAD_DIM + threadIdx.x] = row[2 * d + threadIdx.x];
  }
  __syncthreads();
  const float qscale = rsqrtf((float)GSV_HEAD_DIM) * 1.4426950408889634f;
  float q[8];
  unpack8<T>(*reinterpret_cast<const uint4*>(row + sub * 8), q);
#pragma unroll
  for (int j = 0; j < 8; ++j) q[j] *= qscale;
  const int n = kv + 1;
  float m = GSV_NEG_INF, l = 0.f, o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = 0.f;
  for (int base = warp * 8; base < n; base += 32) {
    const int t = base + pg;
    const bool ok = t < n;
    float s = GSV_NEG_INF, vf[8];
    if (ok) {
      float kf[8];
      unpack8<T>(*reinterpret_cast<const uint4*>(kc + (size_t)t * GSV_HEAD_DIM + sub * 8), kf);
      unpack8<T>(*reinterpret_cast<const uint4*>(vc + (size_t)t * GSV_HEAD_DIM + sub * 8), vf);
      s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s = fmaf(q[j], kf[j], s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (ok) {
      const float mn = fmaxf(m, s);
      const float sc = exp2f(m - mn), pr = exp2f(s - mn);
      l = l * sc + pr;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(pr, vf[j], o[j] * sc);
      m = mn;
    }
  }
#pragma unroll
  for (int off = 4; off < 32; off <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, off);
    const float l2 = __shfl_xor_sync(0xffffffffu, l, off);
    const float mn = fmaxf(m, m2);
    const float a = mn > GSV_NEG_INF ? exp2f(m - mn) : 0.f;
    const float b = mn > GSV_NEG_INF ? exp2f(m2 - mn) : 0.f;
    l = l * a + l2 * b;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = o[j] * a + __shfl_xor_sync(0xffffffffu, o[j], off) * b;
    m = mn;
  }
  if (pg == 0) {
    if (sub == 0) { part[warp][0] = m; part[warp][1] = l; }
#pragma unroll
    for (int j = 0; j < 8; ++j) part[warp][2 + sub * 8 + j] = o[j];
  }
  __syncthreads();
  if (warp == 0) {
    float M = GSV_NEG_INF;
#pragma unroll
    for (int w = 0; w < 4; ++w) M = fmaxf(M, part[w][0]);
    float L = 0.f, acc = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float mw = part[w][0];
      if (mw > GSV_NEG_INF) {
        const float sc = exp2f(mw - M);
        L = fmaf(part[w][1], sc, L);
        acc = fmaf(part[w][2 + lane], sc, acc);
      }
    }
    out[(size_t)slot * d + h * GSV_HEAD_DIM + lane] = Elem<T>::from_f(acc / L);
  }
}

// logits[slot][row] = X[slot] . Whead[row] (fp32; ar_predict_layer has no bias, t2s_model.py:442)
template <typename T>
__global__ void __launch_bounds__(256) head_rows_kernel(const GptParams p, const T* __restrict__ X) {
  const int slot = blockIdx.y;
  if (!p.active[slot]) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= p.V) return;
  const T* w = reinterpret_cast<const T*>(p.w_head) + (size_t)row * p.d;
  const T* xrow = X + (size_t)slot * p.d;
  float acc = 0.f;
  for (int c = lane * 8; c < p.d; c += 256) {
    float wf[8], xf[8];
    unpack8<T>(ld_weight(w + c), wf);
    unpack8<T>(*reinterpret_cast<const uint4*>(xrow + c), xf);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(wf[j], xf[j], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) p.logits[(size_t)slot * GSV_VOCAB_MAX + row] = acc;
}

// kv_len += 1 (t2s_model.py:142), sample, stop flags, next input row
template <typename T>
__global__ void __launch_bounds__(GSV_DECODE_THREADS, 1) sample_step_kernel(const GptParams p, T* __restrict__ X) {
  extern __shared__ __align__(16) float sm[];
  const int slot = blockIdx.x;
  if (!p.active[slot]) return;
  if (threadIdx.x == 0) st_cg(p.kv_len + slot, p.kv_len[slot] + 1);
  __syncthreads();
  sample_slot<T>(p, slot, sm);
  for (int c = threadIdx.x; c < p.d; c += blockDim.x) X[(size_t)slot * p.d + c] = Elem<T>::from_f(ld_cg(p.xin + (size_t)slot * p.d + c));
}

template <typename T>
int decode_step_launches(gsv_gpt_ctx* ctx, cudaStream_t st) {
  const GptParams& p = ctx->p;
  const int d = p.d, F = p.F, L = p.L, nb = p.slots;
  T* X = reinterpret_cast<T*>(ctx->dx);
  T* QKV = reinterpret_cast<T*>(ctx->dqkv);
  T* ATT = reinterpret_cast<T*>(ctx->datt);
  T* Hh = reinterpret_cast<T*>(ctx->dh);
  T* TMP = reinterpret_cast<T*>(ctx->dtmp);
  const size_t op0 = (size_t)4 * L + 8;        // call sites of the decode linears (prefill uses [0, 4L))
  long long dummy = 0;
  int rc;
  for (int l = 0; l < L; ++l) {
    if ((rc = gemm<T>(X, d, reinterpret_cast<const T*>(p.w_qkv) + (size_t)l * 3 * d * d,
                      reinterpret_cast<const T*>(p.b_qkv) + (size_t)l * 3 * d, QKV, 3 * d, nb, 3 * d, d, false, st, dummy, ctx,
                      op0 + (size_t)l * 4 + 0, nb)))
      return rc;
    dec_attn_kernel<T><<<dim3(p.H, nb), 128, 0, st>>>(p, l, QKV, ATT);
    if ((rc = gemm<T>(ATT, d, reinterpret_cast<const T*>(p.w_o) + (size_t)l * d * d,
                      reinterpret_cast<const T*>(p.b_o) + (size_t)l * d, TMP, d, nb, d, d, false, st, dummy, ctx,
                      op0 + (size_t)l * 4 + 1, nb)))
      return rc;
    add_ln_kernel<T><<<(nb + 3) / 4, 128, 0, st>>>(X, TMP, reinterpret_cast<const T*>(p.ln1_g) + (size_t)l * d,
                                                  reinterpret_cast<const T*>(p.ln1_b) + (size_t)l * d, nb, d);
    if ((rc = gemm<T>(X, d, reinterpret_cast<const T*>(p.w_1) + (size_t)l * F * d,
                      reinterpret_cast<const T*>(p.b_1) + (size_t)l * F, Hh, F, nb, F, d, true, st, dummy, ctx,
                      op0 + (size_t)l * 4 + 2, nb)))
      return rc;
    if ((rc = gemm<T>(Hh, F, reinterpret_cast<const T*>(p.w_2) + (size_t)l * d * F,
                      reinterpret_cast<const T*>(p.b_2) + (size_t)l * d, TMP, d, nb, d, F, false, st, dummy, ctx,
                      op0 + (size_t)l * 4 + 3, nb)))
      return rc;
    add_ln_kernel<T><<<(nb + 3) / 4, 128, 0, st>>>(X, TMP, reinterpret_cast<const T*>(p.ln2_g) + (size_t)l * d,
                                                  reinterpret_cast<const T*>(p.ln2_b) + (size_t)l * d, nb, d);
  }
  head_rows_kernel<T><<<dim3((p.V + 7) / 8, nb), 256, 0, st>>>(p, X);
  sample_step_kernel<T><<<nb, GSV_DECODE_THREADS, GSV_SAMPLE_SMEM_FLOATS * sizeof(float), st>>>(p, X);
  GSV_CHECK_LAUNCH();
  return GSV_OK;
}

template <typename T>
int decode_gemm_impl(gsv_gpt_ctx* ctx, int n_steps, cudaStream_t st) {
  const GptParams& p = ctx->p;
  static unsigned long long attr_set = 0ull;        // one bit per device
  if (!((attr_set >> (ctx->device & 63)) & 1ull)) {
    GSV_CUDA(cudaFuncSetAttribute(sample_step_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(GSV_SAMPLE_SMEM_FLOATS * sizeof(float))));
    attr_set |= 1ull << (ctx->device & 63);
  }
  xin_to_x_kernel<T><<<p.slots, 128, 0, st>>>(p, reinterpret_cast<T*>(ctx->dx));
  ctx->launches += 1;
  GSV_CHECK_LAUNCH();
  const long long per_step = 7LL * p.L + 2;
  if (!ctx->step_graph_exec) {
    // one eager step first (it also encodes and caches every tensor map), then capture the same launches once
    int rc = decode_step_launches<T>(ctx, st);
    if (rc) return rc;
    ctx->launches += per_step;
    n_steps -= 1;
    if (n_steps == 0) return GSV_OK;
    cudaStream_t cs;
    GSV_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    cudaGraph_t g = nullptr;
    GSV_CUDA(cudaStreamSynchronize(st));
    GSV_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    rc = decode_step_launches<T>(ctx, cs);
    cudaError_t ce = cudaStreamEndCapture(cs, &g);
    cudaStreamDestroy(cs);
    if (rc) return rc;
    GSV_CUDA(ce);
    cudaGraphExec_t ge = nullptr;
    GSV_CUDA(cudaGraphInstantiate(&ge, g, 0));
    cudaGraphDestroy(g);
    ctx->step_graph_exec = ge;
  }
  for (int i = 0; i < n_steps; ++i) GSV_CUDA(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(ctx->step_graph_exec), st));
  ctx->launches += per_step * n_steps;
  return GSV_OK;
}

}  // namespace

int gsv_gpt_decode_gemm_launch(gsv_gpt_ctx* ctx, int n_steps, cudaStream_t st) {
  if (ctx->dims.dtype == GSV_F16) return decode_gemm_impl<__half>(ctx, n_steps, st);
  return decode_gemm_impl<__nv_bfloat16>(ctx, n_steps, st);
}

int gsv_gpt_prefill_body_many(gsv_gpt_ctx* ctx, int n_prompts, const int* slots, const int64_t* const* xs, const int* nxs,
                              const int64_t* const* ys, const int* nys, const void* const* berts, cudaStream_t st) {
  if (ctx->dims.dtype == GSV_F16) return prefill_body<__half>(ctx, n_prompts, slots, xs, nxs, ys, nys, berts, st);
  return prefill_body<__nv_bfloat16>(ctx, n_prompts, slots, xs, nxs, ys, nys, berts, st);
}

int gsv_gpt_prefill_body(gsv_gpt_ctx* ctx, int slot, const int64_t* x, int nx, const int64_t* y, int ny, const void* bert, cudaStream_t st) {
  return gsv_gpt_prefill_body_many(ctx, 1, &slot, &x, &nx, &y, &ny, &bert, st);
}

int gsv_gpt_prefill_tail(gsv_gpt_ctx* ctx, int slot, const int64_t* y, int ny, const gsv_gpt_sampling* samp, cudaStream_t st) {
  if (ctx->dims.dtype == GSV_F16) return prefill_tail<__half>(ctx, slot, y, ny, *samp, st);
  return prefill_tail<__nv_bfloat16>(ctx, slot, y, ny, *samp, st);
}
