// encp.cu -- the stage between the two hot paths: semantic tokens -> prior latent z_p.
//
// Replaces the front of SynthesizerTrn.decode (reference gsv_tts/GPT_SoVITS/SoVITS/models.py:385-404): codebook lookup
// (module/core_vq.py:133-135, 222-226) + x2 nearest interpolation, ge_to512, TextEncoder.infer (models.py:196-224):
// ssl_proj, three relative-position Transformer encoders (module/attentions.py:10-278), MRTE cross attention
// (module/mrte_model.py:19-38), streaming cross-fade with y_overlap, speed interpolation, proj, and the prior sample
// z_p = m_p + randn * exp(logs_p) * noise_scale (models.py:404).  The reference runs it as 488 eager ops, over the whole
// prefix for every streaming chunk.
//
// Layout: every activation is time-major / channels-last [T][C] in the storage type (the K-major operand of the
// tensor-core kernels); accumulation is fp32.  All 1x1 convolutions and projections run on the tcgen05 implicit-GEMM
// kernel of the vocoder (gsv_umma_linear: one tap), the k = 3 FFN convolutions on the same kernel with three taps
// (gsv_umma_conv, ReLU in the epilogue).  New here: the window-4 relative-position attention (one warp per query row:
// keys over lanes for the scores, output dimensions over lanes for P.V; the 9 relative key / value rows are folded in
// from shared memory), residual + channel LayerNorm, the small gather / combine / interpolation / prior kernels.
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "gpt_internal.cuh"

namespace {

constexpr int WIN = 4;                    // attentions.Encoder window_size (attentions.py:19)
constexpr int AT_WARPS = 8;

struct EW {
  const void* w;
  const void* b;
};

// ---- codebook lookup + x2 nearest interpolation: q[t][:] = codebook[codes[t / 2]][:] ---------------------------------
template <typename T>
__global__ void gather_codes_kernel(const int64_t* __restrict__ codes, int n_codes, const T* __restrict__ book, int n_book, int dim,
                                    T* __restrict__ out) {
  const int t = blockIdx.x;
  long long id = codes[t >> 1];
  id = id < 0 ? 0 : (id >= n_book ? n_book - 1 : id);
  const uint4* src = reinterpret_cast<const uint4*>(book + (size_t)id * dim);
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t)t * dim);
  for (int i = threadIdx.x; i < dim / 8; i += blockDim.x) dst[i] = src[i];
}
template <typename T>
__global__ void embed_text_kernel(const int64_t* __restrict__ text, const T* __restrict__ emb, int n_vocab, int dim, T* __restrict__ out) {
  const int t = blockIdx.x;
  long long id = text[t];
  id = id < 0 ? 0 : (id >= n_vocab ? n_vocab - 1 : id);
  for (int i = threadIdx.x; i < dim / 8; i += blockDim.x)
    reinterpret_cast<uint4*>(out + (size_t)t * dim)[i] = reinterpret_cast<const uint4*>(emb + (size_t)id * dim)[i];
}

// ---- x = LayerNorm_channels(x + y) * gamma + beta (attentions.py:69-79; modules.py:24-27), warp per row, in place -----
template <typename T>
__global__ void __launch_bounds__(128) enc_add_ln_kernel(T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ g,
                                                         const T* __restrict__ b, int n, int d) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp;
  if (row >= n) return;
  float v[16];
  const int per = d >> 5;     // d <= 512, multiple of 32
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (i < per) {
      const int c = i * 32 + lane;
      v[i] = Elem<T>::to_f(Elem<T>::from_f(Elem<T>::to_f(x[(size_t)row * d + c]) + Elem<T>::to_f(y[(size_t)row * d + c])));
      sum += v[i];
    }
  const float mean = warp_sum(sum) / (float)d;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (i < per) sq += (v[i] - mean) * (v[i] - mean);
  const float rstd = rsqrtf(warp_sum(sq) / (float)d + 1e-5f);
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (i < per) {
      const int c = i * 32 + lane;
      x[(size_t)row * d + c] = Elem<T>::from_f((v[i] - mean) * rstd * Elem<T>::to_f(g[c]) + Elem<T>::to_f(b[c]));
    }
}

// ---- multi-head attention with optional window-4 relative-position terms (attentions.py:119-161) ----------------------
// One warp per (query row, head).  q / k / v are rows of channels-last buffers (row strides ldq / ldk / ldv, head h at
// +h*DK).  win: [n_win][2] (start, end) per query row (n_win == Tq) or one for all (n_win == 1): keys outside are masked
// with -1e4 like the reference's masked_fill, key Tk-1 is always kept (MRTE's null key, mrte_model.py:31); null: no mask.  rel_k / rel_v: [2*WIN+1][DK] T or null.  probs: [H][Tq][Tk] fp32 or null.
template <typename T, int DK>
__global__ void __launch_bounds__(AT_WARPS * 32) attn_kernel(const T* __restrict__ q, int ldq, const T* __restrict__ k, int ldk,
                                                             const T* __restrict__ v, int ldv, T* __restrict__ out, int ldo, int Tq,
                                                             int Tk, int H, const T* __restrict__ rel_k, const T* __restrict__ rel_v,
                                                             const int* __restrict__ win, int n_win, float* __restrict__ probs) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * AT_WARPS + warp;
  if (item >= Tq * H) return;
  const int i = item / H, h = item - i * H;
  // key window of this query row: [klo, khi) plus the last key (MRTE), or everything
  int klo = 0, khi = Tk, keep_last = 0;
  if (win != nullptr) { const int r = n_win > 1 ? i : 0; klo = win[2 * r]; khi = win[2 * r + 1]; keep_last = 1; }
  float* sc = sm + (size_t)warp * (Tk + DK + 16);       // [Tk] scores / probabilities
  float* qs = sc + Tk;                                   // [DK] scaled query
  float* rk = qs + DK;                                   // [9] q . E_k[o]
  const float scale = rsqrtf((float)DK);
  for (int d = lane; d < DK; d += 32) qs[d] = Elem<T>::to_f(q[(size_t)i * ldq + h * DK + d]) * scale;
  __syncwarp();
  if (rel_k != nullptr && lane < 2 * WIN + 1) {
    float a = 0.f;
    for (int d = 0; d < DK; ++d) a = fmaf(qs[d], Elem<T>::to_f(rel_k[lane * DK + d]), a);
    rk[lane] = a;
  }
  __syncwarp();
  // pass 1: scores, keys over lanes (the query in registers: read through the shared array it would be re-loaded for every
  // product, the scores being written to the same array)
  float qr[DK];
#pragma unroll
  for (int d = 0; d < DK; ++d) qr[d] = qs[d];
  float mx = -3.0e38f;
  for (int j = lane; j < Tk; j += 32) {
    const uint4* kr = reinterpret_cast<const uint4*>(k + (size_t)j * ldk + h * DK);
    float a = 0.f;
#pragma unroll
    for (int c = 0; c < DK / 8; ++c) {
      float kf[8];
      unpack8<T>(kr[c], kf);
#pragma unroll
      for (int e = 0; e < 8; ++e) a = fmaf(qr[c * 8 + e], kf[e], a);
    }
    if (rel_k != nullptr) {
      const int off = j - i;
      if (off >= -WIN && off <= WIN) a += rk[off + WIN];
    }
    const bool keep = (j >= klo && j < khi) || (keep_last && j == Tk - 1);
    if (!keep) a = -1e4f;
    sc[j] = a;
    mx = fmaxf(mx, a);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < Tk; j += 32) {
    const float e = __expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int j = lane; j < Tk; j += 32) {
    const float pj = sc[j] * inv;
    sc[j] = pj;
    if (probs != nullptr) probs[((size_t)h * Tq + i) * Tk + j] = pj;
  }
  __syncwarp();
  // pass 2: P . V -- four output dimensions per lane (lanes 0 .. DK/4 - 1), the rows of eight keys in flight.  (One dimension
  // triple per lane and one key per iteration made this loop a chain of Tk dependent L2 round trips: 66 us per call at
  // 400 frames, the largest item of a chunk's prior encoder.)  Per dimension the keys are still summed in ascending order.
  constexpr int NL = DK / 4;
  if (lane < NL) {
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    const T* vbase = v + h * DK + lane * 4;
    auto accum = [&](const uint2& r, float pj) {
      const float2 a = Elem<T>::to_f2(r.x), b2 = Elem<T>::to_f2(r.y);
      o[0] = fmaf(pj, a.x, o[0]); o[1] = fmaf(pj, a.y, o[1]); o[2] = fmaf(pj, b2.x, o[2]); o[3] = fmaf(pj, b2.y, o[3]);
    };
    int j = 0;
    for (; j + 8 <= Tk; j += 8) {
      uint2 r[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) r[u] = *reinterpret_cast<const uint2*>(vbase + (size_t)(j + u) * ldv);
#pragma unroll
      for (int u = 0; u < 8; ++u) accum(r[u], sc[j + u]);
    }
    for (; j < Tk; ++j) accum(*reinterpret_cast<const uint2*>(vbase + (size_t)j * ldv), sc[j]);
    if (rel_v != nullptr) {
      for (int off = -WIN; off <= WIN; ++off) {
        const int jr = i + off;
        if (jr < 0 || jr >= Tk) continue;
        accum(*reinterpret_cast<const uint2*>(rel_v + (off + WIN) * DK + lane * 4), sc[jr]);
      }
    }
    __align__(8) T res[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) res[e] = Elem<T>::from_f(o[e]);
    *reinterpret_cast<uint2*>(out + (size_t)i * ldo + h * DK + lane * 4) = *reinterpret_cast<const uint2*>(res);
  }
}

// ---- MRTE: x = attn_out + s + ge (mrte_model.py:36); ge [C][Tg] in torch layout, Tg == 1 or T -------------------------
template <typename T>
__global__ void mrte_combine_kernel(const T* __restrict__ a, const T* __restrict__ s, const T* __restrict__ ge, int Tg, int Tn, int C,
                                    T* __restrict__ out) {
  const int t = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    // the reference adds three 16-bit tensors left to right, rounding after each add
    float v = Elem<T>::to_f(Elem<T>::from_f(Elem<T>::to_f(a[(size_t)t * C + c]) + Elem<T>::to_f(s[(size_t)t * C + c])));
    if (ge != nullptr) v = Elem<T>::to_f(Elem<T>::from_f(v + Elem<T>::to_f(ge[(size_t)c * Tg + (Tg > 1 ? t : 0)])));
    out[(size_t)t * C + c] = Elem<T>::from_f(v);
  }
}

// ---- ge_to512: out[n][t] = W[n][:] . ge[:][t] + b[n] (models.py:396), ge [K][Tg] torch layout -------------------------
template <typename T>
__global__ void ge_proj_kernel(const T* __restrict__ W, const T* __restrict__ b, const T* __restrict__ ge, int K, int Tg, int N,
                               T* __restrict__ out) {
  const int n = blockIdx.x, t = blockIdx.y;
  float a = 0.f;
  for (int kk = threadIdx.x; kk < K; kk += blockDim.x) a = fmaf(Elem<T>::to_f(W[(size_t)n * K + kk]), Elem<T>::to_f(ge[(size_t)kk * Tg + t]), a);
  __shared__ float red[8];
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    out[(size_t)n * Tg + t] = Elem<T>::from_f(s + Elem<T>::to_f(b[n]));
  }
}

// ---- streaming: y = y[valid_start:], cross-fade of the first `ov` frames with the previous chunk's tail, new tail kept --
template <typename T>
__global__ void stream_fade_kernel(const T* __restrict__ y, int valid_start, int Tn, int C, int ov, T* __restrict__ overlap, int has_prev,
                                   T* __restrict__ out) {
  const int t = blockIdx.x;          // 0 .. Tn - valid_start
  const T* src = y + (size_t)(valid_start + t) * C;
  const int Tout = Tn - valid_start;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = Elem<T>::to_f(src[c]);
    if (has_prev && t < ov) {
      // alpha = linspace(0, 1, ov) in the storage type; y_overlap * (1 - alpha) + y * alpha, each op rounded (models.py:211-214)
      const float alpha = Elem<T>::to_f(Elem<T>::from_f(ov > 1 ? (float)t / (float)(ov - 1) : 0.f));
      const float one_m = Elem<T>::to_f(Elem<T>::from_f(1.f - alpha));
      const float a = Elem<T>::to_f(Elem<T>::from_f(Elem<T>::to_f(overlap[(size_t)t * C + c]) * one_m));
      const float bb = Elem<T>::to_f(Elem<T>::from_f(v * alpha));
      v = Elem<T>::to_f(Elem<T>::from_f(a + bb));
    }
    out[(size_t)t * C + c] = Elem<T>::from_f(v);
  }
  (void)Tout;
}
template <typename T>
__global__ void keep_tail_kernel(const T* __restrict__ y, int Tn, int C, int ov, T* __restrict__ overlap) {
  const int t = blockIdx.x;          // 0 .. ov
  for (int c = threadIdx.x; c < C; c += blockDim.x) overlap[(size_t)t * C + c] = y[(size_t)(Tn - ov + t) * C + c];
}
// ---- speed: F.interpolate(mode="linear", align_corners=False) along time (models.py:217-219) -------------------------
template <typename T>
__global__ void interp_linear_kernel(const T* __restrict__ y, int Tin, int Tout, int C, T* __restrict__ out) {
  const int t = blockIdx.x;
  const float scale = (float)Tin / (float)Tout;
  float src = ((float)t + 0.5f) * scale - 0.5f;
  if (src < 0.f) src = 0.f;
  int i0 = (int)src;
  if (i0 > Tin - 1) i0 = Tin - 1;
  const int i1 = i0 + 1 < Tin ? i0 + 1 : Tin - 1;
  const float w1 = src - (float)i0, w0 = 1.f - w1;
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    out[(size_t)t * C + c] = Elem<T>::from_f(w0 * Elem<T>::to_f(y[(size_t)i0 * C + c]) + w1 * Elem<T>::to_f(y[(size_t)i1 * C + c]));
}

// ---- prior: z_p[c][t] = m + noise * exp(logs) * noise_scale (models.py:404); stats [T][2*Co] ---------------------------
template <typename T>
__global__ void prior_kernel(const T* __restrict__ stats, int Tn, int Co, const float* __restrict__ noise, float noise_scale,
                             unsigned long long seed, T* __restrict__ z_p, float* __restrict__ m_out, float* __restrict__ logs_out) {
  const int t = blockIdx.x;
  for (int c = threadIdx.x; c < Co; c += blockDim.x) {
    const float m = Elem<T>::to_f(stats[(size_t)t * 2 * Co + c]);
    const float lg = Elem<T>::to_f(stats[(size_t)t * 2 * Co + Co + c]);
    float n;
    if (noise != nullptr) n = noise[(size_t)c * Tn + t];
    else {
      // standard normal from the counter-based generator (Box-Muller on two uniforms)
      const uint4 r = philox4x32_10(make_uint4((uint32_t)t, (uint32_t)c, 0x6e6f6973u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      const float u1 = ((float)(r.x >> 8) + 1.0f) * (1.0f / 16777216.0f);
      const float u2 = (float)(r.y >> 8) * (1.0f / 16777216.0f);
      n = sqrtf(-2.f * __logf(u1)) * __cosf(6.283185307179586f * u2);
    }
    // randn_like(m_p) * exp(logs_p) * noise_scale + m_p in the storage type, each op rounded
    const float e = Elem<T>::to_f(Elem<T>::from_f(__expf(lg)));
    float v = Elem<T>::to_f(Elem<T>::from_f(Elem<T>::to_f(Elem<T>::from_f(n)) * e));
    v = Elem<T>::to_f(Elem<T>::from_f(v * noise_scale));
    z_p[(size_t)c * Tn + t] = Elem<T>::from_f(m + v);
    if (m_out) m_out[(size_t)c * Tn + t] = m;
    if (logs_out) logs_out[(size_t)c * Tn + t] = lg;
  }
}

}  // namespace

struct gsv_encp_ctx {
  gsv_encp_dims dims;
  std::map<std::string, EW> w;
  gsv_umma_cache* umma;
  int num_sms;
  void* scratch;
  size_t scratch_bytes;
  void* overlap;              // y_overlap: two buffers [2][64][C] T; a streaming call reads [ov_cur] and writes [ov_cur ^ 1]
  int ov_cur;
  int overlap_len;            // 0: none (first chunk / after reset)
  int overlap_len_prev;       // state before the last streaming call (gsv_encp_stream_rollback); -1: nothing to undo
  long long launches;
  size_t op;                  // call-site counter of the tensor-core launches of one forward
  std::vector<void*> owned;
  // text branch of the last forward (embedding, encoder_text, MRTE text_pre and k/v projections -> ckv [Nt][2 Cm]): kept in
  // its own buffer so that the next call on the same text can take it as is (gsv_encp_reuse_text)
  void* text_ckv;
  size_t text_ckv_bytes;
  int text_n;                 // rows held (-1: none)
  int text_reuse_next;
};

namespace {

template <typename T>
struct Fwd {
  gsv_encp_ctx* ctx;
  cudaStream_t st;
  int rc = GSV_OK;

  const EW* get(const std::string& name) {
    auto it = ctx->w.find(name);
    if (it == ctx->w.end()) { gsv_set_error("enc_p: weight '%s' was not set", name.c_str()); rc = GSV_ERR_STATE; return nullptr; }
    return &it->second;
  }
  void linear(const T* x, int rows, int K, const std::string& name, int N, T* out, bool relu = false) {
    if (rc) return;
    const EW* e = get(name);
    if (!e) return;
    rc = gsv_umma_linear(ctx->umma, ctx->op++, ctx->dims.dtype, x, rows, rows, K, e->w, e->b, N, out, relu ? 1 : 0, st);
    ctx->launches += 1;
  }
  void conv(const T* x, int rows, int K, const std::string& name, int N, int KW, T* out, bool relu) {
    if (rc) return;
    const EW* e = get(name);
    if (!e) return;
    rc = gsv_umma_conv(ctx->umma, ctx->op++, ctx->dims.dtype, x, rows, K, e->w, e->b, N, KW, out, relu ? 1 : 0, st);
    ctx->launches += 1;
  }
  template <int DK>
  void attention(const T* q, int ldq, const T* k, int ldk, const T* v, int ldv, T* out, int ldo, int Tq, int Tk, int H, const T* rel_k,
                 const T* rel_v, const int* win, int n_win, float* probs) {
    if (rc) return;
    const size_t smem = (size_t)AT_WARPS * (Tk + DK + 16) * sizeof(float);
    if (smem > 200 * 1024) { gsv_set_error("enc_p: %d keys exceed the attention kernel's shared memory", Tk); rc = GSV_ERR_ARG; return; }
    cudaFuncSetAttribute(attn_kernel<T, DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attn_kernel<T, DK><<<(Tq * H + AT_WARPS - 1) / AT_WARPS, AT_WARPS * 32, smem, st>>>(q, ldq, k, ldk, v, ldv, out, ldo, Tq, Tk, H, rel_k,
                                                                                       rel_v, win, n_win, probs);
    ctx->launches += 1;
  }
  // attentions.Encoder (attentions.py:59-80) on x [Tn][C], in place; buffers: qkv [Tn][3C], att [Tn][C], tmp [Tn][C], hid [Tn][F]
  void encoder(T* x, int Tn, const std::string& pre, int n_layers, T* qkv, T* att, T* tmp, T* hid) {
    const int C = ctx->dims.hidden_channels, F = ctx->dims.filter_channels, H = ctx->dims.n_heads, KW = ctx->dims.kernel_size;
    for (int l = 0; l < n_layers && !rc; ++l) {
      const std::string a = pre + "attn_layers." + std::to_string(l) + ".";
      linear(x, Tn, C, a + "qkv", 3 * C, qkv);
      const EW* ek = get(a + "emb_rel_k");
      const EW* ev = get(a + "emb_rel_v");
      if (!ek || !ev) return;
      attention<96>(qkv, 3 * C, qkv + C, 3 * C, qkv + 2 * C, 3 * C, att, C, Tn, Tn, H, reinterpret_cast<const T*>(ek->w),
                    reinterpret_cast<const T*>(ev->w), nullptr, 0, nullptr);
      linear(att, Tn, C, a + "conv_o", C, tmp);
      const EW* n1 = get(pre + "norm_layers_1." + std::to_string(l));
      if (!n1) return;
      enc_add_ln_kernel<T><<<(Tn + 3) / 4, 128, 0, st>>>(x, tmp, reinterpret_cast<const T*>(n1->w), reinterpret_cast<const T*>(n1->b), Tn, C);
      const std::string f = pre + "ffn_layers." + std::to_string(l) + ".";
      conv(x, Tn, C, f + "conv_1", F, KW, hid, true);
      conv(hid, Tn, F, f + "conv_2", C, KW, tmp, false);
      const EW* n2 = get(pre + "norm_layers_2." + std::to_string(l));
      if (!n2) return;
      enc_add_ln_kernel<T><<<(Tn + 3) / 4, 128, 0, st>>>(x, tmp, reinterpret_cast<const T*>(n2->w), reinterpret_cast<const T*>(n2->b), Tn, C);
      ctx->launches += 2;
    }
  }
};

template <typename T>
int encp_forward_t(gsv_encp_ctx* ctx, const int64_t* codes, int n_codes, const int64_t* text, int n_text, const void* ge_v, int Tg,
                   float speed, int stream_mode, int valid_start, int overlap_len, const int* slices, int n_slices, const float* noise,
                   float noise_scale, unsigned long long seed, void* z_p_v, float* m_out, float* logs_out, float* attn_out, int* out_T,
                   cudaStream_t st) {
  const gsv_encp_dims& d = ctx->dims;
  const int C = d.hidden_channels, F = d.filter_channels, Co = d.inter_channels, Cm = d.mrte_channels, L = d.n_layers;
  const int Tn = 2 * n_codes, Nt = n_text;
  GSV_ARG(C == 192 && d.n_heads == 2 && Cm == 512 && d.mrte_heads == 4);      // head sizes the attention kernel is built for (96, 128)
  GSV_ARG(d.kernel_size % 2 == 1);
  // scratch: q768 [Tn][768] | y [Tn][C] | qkv [Tm][3C] | att [Tm][C] | tmp [Tm][C] | hid [Tm][F] | t [Nt][C] | s [Tn][Cm] | tp [Nt][Cm] |
  //          cq [Tn][Cm] | ckv [Nt][2Cm] | ca [Tn][Cm] | cx [Tn][Cm] | y2 [Tn][C] | y3 [Tout][C] | stats [Tout][2Co] | ge512 [Cm][Tg]
  const int Tm = Tn > Nt ? Tn : Nt;
  int T_after = stream_mode ? Tn - valid_start : Tn;
  GSV_ARG(!stream_mode || (valid_start >= 0 && valid_start < Tn && overlap_len >= 1 && overlap_len <= T_after && overlap_len <= 64));
  int T_out = T_after;
  if (speed != 1.f) T_out = (int)((float)T_after / speed) + 1;
  const size_t el = sizeof(T);
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += al(bytes); return o; };
  const size_t o_q768 = take((size_t)Tn * d.ssl_dim * el), o_y = take((size_t)Tn * C * el), o_qkv = take((size_t)Tm * 3 * C * el),
               o_att = take((size_t)Tm * C * el), o_tmp = take((size_t)Tm * C * el), o_hid = take((size_t)Tm * F * el),
               o_t = take((size_t)Nt * C * el), o_s = take((size_t)Tn * Cm * el), o_tp = take((size_t)Nt * Cm * el),
               o_cq = take((size_t)Tn * Cm * el), o_ckv = take((size_t)Nt * 2 * Cm * el), o_ca = take((size_t)Tn * Cm * el),
               o_cx = take((size_t)Tn * Cm * el), o_y3 = take((size_t)(T_out > Tn ? T_out : Tn) * C * el),
               o_y4 = take((size_t)(T_out > Tn ? T_out : Tn) * C * el), o_stats = take((size_t)T_out * 2 * Co * el),
               o_ge = take((size_t)Cm * (Tg > 1 ? Tn : 1) * el);
  if (off > ctx->scratch_bytes) {
    GSV_CUDA(cudaStreamSynchronize(st));
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    GSV_CUDA(cudaMalloc(&ctx->scratch, off));
    ctx->scratch_bytes = off;
  }
  unsigned char* base = reinterpret_cast<unsigned char*>(ctx->scratch);
  T* q768 = reinterpret_cast<T*>(base + o_q768); T* y = reinterpret_cast<T*>(base + o_y); T* qkv = reinterpret_cast<T*>(base + o_qkv);
  T* att = reinterpret_cast<T*>(base + o_att); T* tmp = reinterpret_cast<T*>(base + o_tmp); T* hid = reinterpret_cast<T*>(base + o_hid);
  T* tx = reinterpret_cast<T*>(base + o_t); T* s = reinterpret_cast<T*>(base + o_s); T* tp = reinterpret_cast<T*>(base + o_tp);
  T* cq = reinterpret_cast<T*>(base + o_cq); T* ckv = reinterpret_cast<T*>(base + o_ckv); T* ca = reinterpret_cast<T*>(base + o_ca);
  T* cx = reinterpret_cast<T*>(base + o_cx); T* y3 = reinterpret_cast<T*>(base + o_y3); T* y4 = reinterpret_cast<T*>(base + o_y4);
  T* stats = reinterpret_cast<T*>(base + o_stats); T* ge512 = reinterpret_cast<T*>(base + o_ge);

  Fwd<T> f;
  f.ctx = ctx; f.st = st;
  ctx->op = 0;
  const T* ge = reinterpret_cast<const T*>(ge_v);
  GSV_ARG(Tg == 1 || Tg == Tn);
  // ge for the MRTE residual: through ge_to512 on v2Pro models (models.py:396), as given otherwise
  const T* ge_in = ge;
  if (ge != nullptr && ctx->w.count("ge_to512")) {
    const EW* e = f.get("ge_to512");
    ge_proj_kernel<T><<<dim3(Cm, Tg), 128, 0, st>>>(reinterpret_cast<const T*>(e->w), reinterpret_cast<const T*>(e->b), ge, d.gin_channels, Tg, Cm, ge512);
    ctx->launches += 1;
    ge_in = ge512;
  }
  // content branch
  const EW* book = f.get("quantizer.codebook");
  if (!book) return f.rc;
  gather_codes_kernel<T><<<Tn, 96, 0, st>>>(codes, n_codes, reinterpret_cast<const T*>(book->w), d.n_codes, d.ssl_dim, q768);
  ctx->launches += 1;
  f.linear(q768, Tn, d.ssl_dim, "enc_p.ssl_proj", C, y);
  f.encoder(y, Tn, "enc_p.encoder_ssl.", L / 2, qkv, att, tmp, hid);
  // text branch: text -> the MRTE's keys / values ckv [Nt][2 Cm] (6 encoder layers + 2 linears = 45 launches).  It does not
  // depend on the codes: the chunks of one streaming utterance after the first take it from the previous call.
  const bool reuse_text = ctx->text_reuse_next && ctx->text_n == Nt && ctx->text_ckv != nullptr;
  ctx->text_reuse_next = 0;
  {
    const size_t need = (size_t)Nt * 2 * Cm * el;
    if (need > ctx->text_ckv_bytes) {
      GSV_CUDA(cudaStreamSynchronize(st));
      if (ctx->text_ckv) cudaFree(ctx->text_ckv);
      ctx->text_ckv = nullptr; ctx->text_ckv_bytes = 0; ctx->text_n = -1;
      GSV_CUDA(cudaMalloc(&ctx->text_ckv, need));
      ctx->text_ckv_bytes = need;
    }
    ckv = reinterpret_cast<T*>(ctx->text_ckv);
  }
  if (!reuse_text || ctx->text_n != Nt) {
    ctx->text_n = -1;
    const EW* emb = f.get("enc_p.text_embedding");
    if (!emb) return f.rc;
    embed_text_kernel<T><<<Nt, 32, 0, st>>>(text, reinterpret_cast<const T*>(emb->w), d.n_symbols, C, tx);
    ctx->launches += 1;
    f.encoder(tx, Nt, "enc_p.encoder_text.", L, qkv, att, tmp, hid);
    f.linear(tx, Nt, C, "enc_p.mrte.text_pre", Cm, tp);
    f.linear(tp, Nt, Cm, "enc_p.mrte.cross_attention.kv", 2 * Cm, ckv);
    if (f.rc) return f.rc;
    ctx->text_n = Nt;
  } else {
    ctx->op += (size_t)4 * L + 2;              // the skipped call sites keep their tensor-map slots
  }
  // MRTE (mrte_model.py:19-38)
  f.linear(y, Tn, C, "enc_p.mrte.c_pre", Cm, s);
  f.linear(s, Tn, Cm, "enc_p.mrte.cross_attention.conv_q", Cm, cq);
  f.template attention<128>(cq, Cm, ckv, 2 * Cm, ckv + Cm, 2 * Cm, ca, Cm, Tn, Nt, d.mrte_heads, nullptr, nullptr, slices, n_slices, attn_out);
  f.linear(ca, Tn, Cm, "enc_p.mrte.cross_attention.conv_o", Cm, cx);
  if (f.rc) return f.rc;
  mrte_combine_kernel<T><<<Tn, 128, 0, st>>>(cx, s, ge_in, Tg, Tn, Cm, ca);
  ctx->launches += 1;
  f.linear(ca, Tn, Cm, "enc_p.mrte.c_post", C, y);
  f.encoder(y, Tn, "enc_p.encoder2.", L / 2, qkv, att, tmp, hid);
  if (f.rc) return f.rc;
  // streaming cross-fade (models.py:208-215)
  T* cur = y;
  int Tc = Tn;
  if (stream_mode) {
    const bool has_prev = ctx->overlap_len == overlap_len && ctx->overlap != nullptr;
    if (!ctx->overlap) GSV_CUDA(cudaMalloc(&ctx->overlap, (size_t)2 * 64 * C * el));
    Tc = Tn - valid_start;
    // the previous tail is left intact (the new one goes to the other buffer): one call can be undone, in stream order
    T* ov_prev = reinterpret_cast<T*>(ctx->overlap) + (size_t)ctx->ov_cur * 64 * C;
    T* ov_next = reinterpret_cast<T*>(ctx->overlap) + (size_t)(ctx->ov_cur ^ 1) * 64 * C;
    stream_fade_kernel<T><<<Tc, 64, 0, st>>>(y, valid_start, Tn, C, overlap_len, ov_prev, has_prev ? 1 : 0, y3);
    keep_tail_kernel<T><<<overlap_len, 64, 0, st>>>(y3, Tc, C, overlap_len, ov_next);
    ctx->overlap_len_prev = ctx->overlap_len;
    ctx->ov_cur ^= 1;
    ctx->overlap_len = overlap_len;
    ctx->launches += 2;
    cur = y3;
  }
  if (speed != 1.f) {
    interp_linear_kernel<T><<<T_out, 64, 0, st>>>(cur, Tc, T_out, C, y4);
    ctx->launches += 1;
    cur = y4;
    Tc = T_out;
  }
  f.linear(cur, Tc, C, "enc_p.proj", 2 * Co, stats);
  if (f.rc) return f.rc;
  prior_kernel<T><<<Tc, 64, 0, st>>>(stats, Tc, Co, noise, noise_scale, seed, reinterpret_cast<T*>(z_p_v), m_out, logs_out);
  ctx->launches += 1;
  GSV_CHECK_LAUNCH();
  if (out_T) *out_T = Tc;
  return GSV_OK;
}

}  // namespace

extern "C" int gsv_encp_create(const gsv_encp_dims* dims, gsv_encp_ctx** out) {
  GSV_ARG(dims && out);
  GSV_ARG(dims->dtype == GSV_F16 || dims->dtype == GSV_BF16);
  int dev = 0;
  GSV_CUDA(cudaGetDevice(&dev));
  int rc = gsv_device_check(dev);
  if (rc) return rc;
  gsv_encp_ctx* ctx = new (std::nothrow) gsv_encp_ctx();
  GSV_ARG(ctx != nullptr);
  ctx->dims = *dims;
  ctx->scratch = nullptr; ctx->scratch_bytes = 0; ctx->overlap = nullptr; ctx->ov_cur = 0; ctx->overlap_len = 0; ctx->overlap_len_prev = -1; ctx->launches = 0; ctx->op = 0;
  ctx->text_ckv = nullptr; ctx->text_ckv_bytes = 0; ctx->text_n = -1; ctx->text_reuse_next = 0;
  GSV_CUDA(cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, dev));
  ctx->umma = gsv_umma_cache_create(ctx->num_sms);
  *out = ctx;
  return GSV_OK;
}

extern "C" int gsv_encp_set_weight(gsv_encp_ctx* ctx, const char* name, const void* dev_weight, const void* dev_bias) {
  GSV_ARG(ctx && name && dev_weight);
  ctx->w[name] = EW{dev_weight, dev_bias};
  ctx->text_n = -1;                            // whatever was computed with the old weights is stale
  return GSV_OK;
}

// The next gsv_encp_forward call on this context carries the same text as the previous one: its text branch is reused.
extern "C" int gsv_encp_reuse_text(gsv_encp_ctx* ctx, int on) {
  GSV_ARG(ctx);
  ctx->text_reuse_next = on ? 1 : 0;
  return GSV_OK;
}

extern "C" int gsv_encp_destroy(gsv_encp_ctx* ctx) {
  if (!ctx) return GSV_OK;
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->overlap) cudaFree(ctx->overlap);
  if (ctx->text_ckv) cudaFree(ctx->text_ckv);
  for (void* p : ctx->owned) cudaFree(p);
  gsv_umma_cache_destroy(ctx->umma);
  delete ctx;
  return GSV_OK;
}

extern "C" int gsv_encp_reset_stream(gsv_encp_ctx* ctx) {
  GSV_ARG(ctx);
  ctx->overlap_len = 0;
  ctx->overlap_len_prev = -1;
  return GSV_OK;
}

// Undo the last streaming call's effect on the cross-chunk state (a chunk computed ahead of time that the stream then
// merged into its final chunk, TTS.infer_phones_stream): the previous tail is still in its buffer.  One level.
extern "C" int gsv_encp_stream_rollback(gsv_encp_ctx* ctx) {
  GSV_ARG(ctx);
  if (ctx->overlap_len_prev < 0) { gsv_set_error("gsv_encp_stream_rollback: no streaming call to undo"); return GSV_ERR_STATE; }
  ctx->ov_cur ^= 1;
  ctx->overlap_len = ctx->overlap_len_prev;
  ctx->overlap_len_prev = -1;
  return GSV_OK;
}

extern "C" int gsv_encp_output_frames(gsv_encp_ctx* ctx, int n_codes, float speed, int stream_mode, int valid_start) {
  (void)ctx;
  int t = 2 * n_codes;
  if (stream_mode) t -= valid_start;
  if (speed != 1.f) t = (int)((float)t / speed) + 1;
  return t;
}

extern "C" int gsv_encp_forward(gsv_encp_ctx* ctx, const int64_t* dev_codes, int n_codes, const int64_t* dev_text, int n_text,
                                const void* dev_ge, int Tg, float speed, int stream_mode, int valid_start, int overlap_len,
                                const int32_t* dev_slices, int n_slices, const float* dev_noise, float noise_scale, uint64_t seed, void* dev_z_p, float* dev_m_p,
                                float* dev_logs_p, float* dev_attn, int* out_frames, void* stream) {
  GSV_ARG(ctx && dev_codes && dev_text && dev_z_p && n_codes >= 1 && n_text >= 1);
  GSV_ARG(dev_slices == nullptr || n_slices == 1 || n_slices == 2 * n_codes);
  if (ctx->dims.dtype == GSV_F16)
    return encp_forward_t<__half>(ctx, dev_codes, n_codes, dev_text, n_text, dev_ge, Tg, speed, stream_mode, valid_start, overlap_len, dev_slices,
                                  n_slices, dev_noise, noise_scale, seed, dev_z_p, dev_m_p, dev_logs_p, dev_attn, out_frames, (cudaStream_t)stream);
  return encp_forward_t<__nv_bfloat16>(ctx, dev_codes, n_codes, dev_text, n_text, dev_ge, Tg, speed, stream_mode, valid_start, overlap_len,
                                       dev_slices, n_slices, dev_noise, noise_scale, seed, dev_z_p, dev_m_p, dev_logs_p, dev_attn, out_frames,
                                       (cudaStream_t)stream);
}

extern "C" int64_t gsv_encp_launch_count(gsv_encp_ctx* ctx) { return ctx ? ctx->launches : 0; }
