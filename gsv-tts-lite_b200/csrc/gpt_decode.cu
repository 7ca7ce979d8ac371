// gpt_decode.cu -- persistent decode kernel: n_steps tokens for every active slot in ONE launch.
//
// Replaces, per token, the reference's CUDA-graph replay of 24 x T2SBlock.decode_next_token
// (gsv_tts/GPT_SoVITS/GPT/t2s_model.py:67-105, 129-143; ~300 graph nodes) plus the ~25 eager
// ops of ar_predict_layer / sample / embedding (:442-456) and its per-token host sync (:426).
//
// Shape of the work: B <= 32 sequences, d = 512, so every matrix product is a skinny GEMV that
// is bound by streaming the 152 MB of weights from HBM, and the step is a chain of ~5 dependent
// phases per layer.  One CTA per SM stays resident (cooperative launch); phases are separated by
// a hand-rolled grid barrier; each warp prefetches the weight row of its *next* phase into
// registers before it waits on the barrier, so HBM latency overlaps the barrier instead of adding
// to it.  Activations cross CTAs through small fp32 buffers in L2 (ld.cg / st.cg); LayerNorm is
// recomputed redundantly by every CTA instead of costing another barrier.
//
// Phases of layer l:
//   P1  x = (l==0 ? xin : LN2(y2));  [q|k|v] = x Wqkv^T + b;  k,v appended to the cache at kv_len
//   P2  split-KV attention partials (m, l, o[32]) per (slot, head, split) over 0..kv_len inclusive
//   P3  a = combine(partials);  y1 = x + a Wo^T + bo
//   P4  x1 = LN1(y1);  h = relu(x1 W1^T + b1)
//   P5  y2 = x1 + h W2^T + b2
// then once per step:
//   P6  logits = LN2(y2) Whead^T;  kv_len += 1
//   P7  one CTA per slot: sample, append token, stop flags, next input embedding (gpt_sample.cuh)
#include "gpt_sample.cuh"

namespace {

constexpr int NT = GSV_DECODE_THREADS;
constexpr int NWARP = NT / 32;
constexpr int TILE = GSV_SLOT_TILE;

// Shared-memory layout of a staged activation row of length K (fp32): element k = 8*ch + j lives at
// ch*4 + j (j < 4) or K/2 + ch*4 + (j-4): a lane that owns weight chunk `ch` reads two float4 that
// are contiguous across lanes -> conflict-free LDS.128.
__device__ __forceinline__ int split_pos(int k, int K) {
  const int ch = k >> 3, j = k & 7;
  return j < 4 ? ch * 4 + j : (K >> 1) + ch * 4 + (j - 4);
}

template <typename T, int NCH>
__device__ __forceinline__ void load_row(const T* __restrict__ W, size_t row, int lane, uint4 (&w)[NCH]) {
  const uint4* src = reinterpret_cast<const uint4*>(W + row * (size_t)(NCH * 256));
#pragma unroll
  for (int c = 0; c < NCH; ++c) w[c] = ld_weight(src + c * 32 + lane);
}

template <typename T, int NCH>
__device__ __forceinline__ void dot_tile(const uint4 (&w)[NCH], const float* xs, int nt, int lane, float (&acc)[TILE]) {
  constexpr int K = NCH * 256;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    float wf[8];
    unpack8<T>(w[c], wf);
    const int ch = c * 32 + lane;
#pragma unroll
    for (int s = 0; s < TILE; ++s) {
      if (s < nt) {
        const float4 lo = *reinterpret_cast<const float4*>(xs + s * K + ch * 4);
        const float4 hi = *reinterpret_cast<const float4*>(xs + s * K + (K >> 1) + ch * 4);
        float a = acc[s];
        a = fmaf(wf[0], lo.x, a); a = fmaf(wf[1], lo.y, a); a = fmaf(wf[2], lo.z, a); a = fmaf(wf[3], lo.w, a);
        a = fmaf(wf[4], hi.x, a); a = fmaf(wf[5], hi.y, a); a = fmaf(wf[6], hi.z, a); a = fmaf(wf[7], hi.w, a);
        acc[s] = a;
      }
    }
  }
}

// Transposing butterfly: 8 per-lane partial sums -> lane holds the full sum of value
// idx = 4*bit4 + 2*bit3 + bit2 of its lane id (9 shuffles instead of 40).
__device__ __forceinline__ float reduce8(const float (&acc)[TILE], int lane) {
  float a4[4], a2[2], a1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = (lane & 16) ? acc[i] : acc[i + 4];
    const float keep = (lane & 16) ? acc[i + 4] : acc[i];
    a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = (lane & 8) ? a4[i] : a4[i + 2];
    const float keep = (lane & 8) ? a4[i + 2] : a4[i];
    a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const float send = (lane & 4) ? a2[0] : a2[1];
    const float keep = (lane & 4) ? a2[1] : a2[0];
    a1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
  a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
  return a1;
}

struct StepShared {
  int sl[GSV_MAX_SLOTS];   // active slot ids
  int kv[GSV_MAX_SLOTS];   // their kv_len at the start of the step
  int nb;
};

__device__ __forceinline__ void build_active(const GptParams& p, StepShared& ss) {
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const int flag = lane < p.slots ? ld_cg(p.active + lane) : 0;
    const unsigned m = __ballot_sync(0xffffffffu, flag != 0);
    if (flag) {
      const int pos = __popc(m & ((1u << lane) - 1u));
      ss.sl[pos] = lane;
      ss.kv[pos] = ld_cg(p.kv_len + lane);
    }
    if (lane == 0) ss.nb = __popc(m);
  }
  __syncthreads();
}

// ---- staging of one slot tile into shared memory ---------------------------------------------
// plain copy of src[slot][K] (fp32)
__device__ __forceinline__ void stage_plain(float* xs, const float* src, int K, const StepShared& ss, int t0, int nt,
                                            float* copy_to) {
  const int n4 = nt * (K >> 2);
  for (int i = threadIdx.x; i < n4; i += NT) {
    const int s = i / (K >> 2), k = (i - s * (K >> 2)) << 2;
    const int slot = ss.sl[t0 + s];
    const float4 v = ld_cg4(src + (size_t)slot * K + k);
    *reinterpret_cast<float4*>(xs + s * K + split_pos(k, K)) = v;
    if (copy_to) __stcg(reinterpret_cast<float4*>(copy_to + (size_t)slot * K + k), v);
  }
}

// LayerNorm(src[slot][d]) * gamma + beta, eps 1e-5 (nn.LayerNorm, t2s_model.py:20,24); warp per slot
template <typename T>
__device__ __forceinline__ void stage_ln(float* xs, const float* src, const T* gamma, const T* beta, int d,
                                         const StepShared& ss, int t0, int nt, float* copy_to) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < nt) {
    const int slot = ss.sl[t0 + warp];
    const int nf4 = d >> 7;            // float4 per lane (d multiple of 256 -> 2,4,6,8)
    float4 v[8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nf4) {
        v[i] = ld_cg4(src + (size_t)slot * d + ((i * 32 + lane) << 2));
        sum += v[i].x + v[i].y + v[i].z + v[i].w;
      }
    const float mean = warp_sum(sum) / (float)d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nf4) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
        sq += a * a + b * b + c * c + e * e;
      }
    const float rstd = rsqrtf(warp_sum(sq) / (float)d + 1e-5f);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nf4) {
        const int k = (i * 32 + lane) << 2;
        const uint2 g2 = *reinterpret_cast<const uint2*>(gamma + k);
        const uint2 b2 = *reinterpret_cast<const uint2*>(beta + k);
        const float2 g01 = Elem<T>::to_f2(g2.x), g23 = Elem<T>::to_f2(g2.y);
        const float2 b01 = Elem<T>::to_f2(b2.x), b23 = Elem<T>::to_f2(b2.y);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g01.x + b01.x;
        o.y = (v[i].y - mean) * rstd * g01.y + b01.y;
        o.z = (v[i].z - mean) * rstd * g23.x + b23.x;
        o.w = (v[i].w - mean) * rstd * g23.y + b23.y;
        *reinterpret_cast<float4*>(xs + warp * d + split_pos(k, d)) = o;
        if (copy_to) __stcg(reinterpret_cast<float4*>(copy_to + (size_t)slot * d + k), o);
      }
  }
}

// merge the split-KV partials of every (slot, head) of the tile into the attention output row
__device__ __forceinline__ void stage_attn(float* xs, const GptParams& p, const StepShared& ss, int t0, int nt, int NS) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = p.d, H = p.H;
  for (int pi = warp; pi < nt * H; pi += NWARP) {
    const int s = pi / H, h = pi - s * H;
    const int slot = ss.sl[t0 + s];
    const float* base = p.part + ((size_t)(slot * H + h) * GSV_NSPLIT_MAX) * GSV_PART_STRIDE;
    float M = GSV_NEG_INF;
    for (int k = 0; k < NS; ++k) M = fmaxf(M, ld_cg(base + k * GSV_PART_STRIDE));
    float L = 0.f, o = 0.f;
    for (int k = 0; k < NS; ++k) {
      const float m = ld_cg(base + k * GSV_PART_STRIDE);
      if (m > GSV_NEG_INF) {
        const float sc = exp2f(m - M);
        L += ld_cg(base + k * GSV_PART_STRIDE + 1) * sc;
        o += ld_cg(base + k * GSV_PART_STRIDE + 4 + lane) * sc;
      }
    }
    xs[s * d + split_pos(h * 32 + lane, d)] = o / L;
  }
}

// ---- one GEMV phase: out[slot][row] = epi(row, slot, W[row] . x[slot]) for all active slots ------------
template <typename T, int NCH, typename Stage, typename Epi>
__device__ __forceinline__ void gemv_phase(const T* __restrict__ W, int N, float* xs, const StepShared& ss,
                                           unsigned* barrier, unsigned& epoch, Stage stage, Epi epi) {
  constexpr int K = NCH * 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;
  const int row0 = blockIdx.x + G * warp;
  uint4 w[NCH];
  if (row0 < N) load_row<T, NCH>(W, row0, lane, w);     // in flight while we wait on the barrier
  grid_sync(barrier, epoch);
  const int nb = ss.nb;
  for (int t0 = 0; t0 < nb; t0 += TILE) {
    const int nt = min(TILE, nb - t0);
    if (t0 > 0) __syncthreads();
    stage(t0, nt);
    __syncthreads();
    for (int row = row0; row < N; row += G * NWARP) {
      if (row != row0 || t0 != 0) load_row<T, NCH>(W, row, lane, w);
      float acc[TILE];
#pragma unroll
      for (int s = 0; s < TILE; ++s) acc[s] = 0.f;
      dot_tile<T, NCH>(w, xs, nt, lane, acc);
      const float r = reduce8(acc, lane);
      const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
      if ((lane & 3) == 0 && idx < nt) epi(row, t0 + idx, r);
    }
  }
  (void)K;
}

// ---- P2: split-KV attention, one warp per (slot, head, split) ------------------------------------
template <typename T>
__device__ __forceinline__ void attention_phase(const GptParams& p, const StepShared& ss, int layer, int NS) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, H = p.H, d = p.d, S = p.S;
  const int sub = lane & 3, pg = lane >> 2;      // 4 lanes x 8 dims cover one position; 8 positions per pass
  const int items = ss.nb * H * NS;
  const float qscale = rsqrtf((float)GSV_HEAD_DIM) * 1.4426950408889634f;   // 1/sqrt(dh) * log2(e)
  for (int it = blockIdx.x + G * warp; it < items; it += G * NWARP) {
    const int si = it / (H * NS);
    const int rem = it - si * H * NS;
    const int h = rem / NS, sp = rem - h * NS;
    const int slot = ss.sl[si];
    const int n = ss.kv[si] + 1;                                    // positions 0..kv_len inclusive
    const int chunk = (((n + NS - 1) / NS) + 7) & ~7;
    const int tb = sp * chunk, te = min(n, tb + chunk);
    float q[8];
    {
      const float* qp = p.q + (size_t)slot * d + h * GSV_HEAD_DIM + sub * 8;
      const float4 a = ld_cg4(qp), b = ld_cg4(qp + 4);
      q[0] = a.x * qscale; q[1] = a.y * qscale; q[2] = a.z * qscale; q[3] = a.w * qscale;
      q[4] = b.x * qscale; q[5] = b.y * qscale; q[6] = b.z * qscale; q[7] = b.w * qscale;
    }
    const size_t head_base = ((size_t)(layer * p.slots + slot) * H + h) * (size_t)S * GSV_HEAD_DIM;
    const T* kb = reinterpret_cast<const T*>(p.kc) + head_base + sub * 8;
    const T* vb = reinterpret_cast<const T*>(p.vc) + head_base + sub * 8;
    float m = GSV_NEG_INF, l = 0.f, o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.f;
    for (int tt = tb; tt < te; tt += 8) {            // uniform trip count inside the warp
      const int t = tt + pg;
      const bool ok = t < te;
      float s = GSV_NEG_INF;
      float vf[8];
      if (ok) {
        float kf[8];
        unpack8<T>(ld_cg16(kb + (size_t)t * GSV_HEAD_DIM), kf);
        unpack8<T>(ld_cg16(vb + (size_t)t * GSV_HEAD_DIM), vf);
        s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s = fmaf(q[j], kf[j], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (ok) {
        const float mn = fmaxf(m, s);
        const float sc = exp2f(m - mn);       // m = -inf on the first hit -> 0
        const float pr = exp2f(s - mn);
        l = l * sc + pr;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(pr, vf[j], o[j] * sc);
        m = mn;
      }
    }
    // merge the 8 position groups
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
      float* out = p.part + ((size_t)(slot * H + h) * GSV_NSPLIT_MAX + sp) * GSV_PART_STRIDE;
      if (sub == 0) { st_cg(out, m); st_cg(out + 1, l); }
      __stcg(reinterpret_cast<float4*>(out + 4 + sub * 8), make_float4(o[0], o[1], o[2], o[3]));
      __stcg(reinterpret_cast<float4*>(out + 8 + sub * 8), make_float4(o[4], o[5], o[6], o[7]));
    }
  }
}

template <typename T, int NCH_D, int NCH_F>
__global__ void __launch_bounds__(NT, 1) gpt_decode_kernel(const GptParams p, const int n_steps) {
  extern __shared__ __align__(16) float xs[];
  __shared__ StepShared ss;
  unsigned epoch = 0;
  const int d = p.d, F = p.F, H = p.H, L = p.L;
  const T* const w_qkv = reinterpret_cast<const T*>(p.w_qkv);
  const T* const b_qkv = reinterpret_cast<const T*>(p.b_qkv);
  const T* const w_o = reinterpret_cast<const T*>(p.w_o);
  const T* const b_o = reinterpret_cast<const T*>(p.b_o);
  const T* const w_1 = reinterpret_cast<const T*>(p.w_1);
  const T* const b_1 = reinterpret_cast<const T*>(p.b_1);
  const T* const w_2 = reinterpret_cast<const T*>(p.w_2);
  const T* const b_2 = reinterpret_cast<const T*>(p.b_2);
  const T* const ln1_g = reinterpret_cast<const T*>(p.ln1_g);
  const T* const ln1_b = reinterpret_cast<const T*>(p.ln1_b);
  const T* const ln2_g = reinterpret_cast<const T*>(p.ln2_g);
  const T* const ln2_b = reinterpret_cast<const T*>(p.ln2_b);
  const T* const w_head = reinterpret_cast<const T*>(p.w_head);
  T* const kc = reinterpret_cast<T*>(p.kc);
  T* const vc = reinterpret_cast<T*>(p.vc);
  const bool lead = blockIdx.x == 0;

  build_active(p, ss);
  for (int step = 0; step < n_steps; ++step) {
    const int nb = ss.nb;
    if (nb == 0) break;
    const int NS = max(1, min(GSV_NSPLIT_MAX, (int)(gridDim.x * NWARP) / (nb * H)));

    for (int l = 0; l < L; ++l) {
      // ---- P1: QKV projection + KV append
      gemv_phase<T, NCH_D>(
          w_qkv + (size_t)l * 3 * d * d, 3 * d, xs, ss, p.barrier, epoch,
          [&](int t0, int nt) {
            if (l == 0) stage_plain(xs, p.xin, d, ss, t0, nt, lead ? p.xres : nullptr);
            else stage_ln<T>(xs, p.y2, ln2_g + (size_t)(l - 1) * d, ln2_b + (size_t)(l - 1) * d, d, ss, t0, nt,
                             lead ? p.xres : nullptr);
          },
          [&](int row, int si, float r) {
            const int slot = ss.sl[si];
            const float v = r + Elem<T>::to_f(b_qkv[(size_t)l * 3 * d + row]);
            if (row < d) {
              st_cg(p.q + (size_t)slot * d + row, v);
            } else {
              const int c = row < 2 * d ? row - d : row - 2 * d;
              T* cache = row < 2 * d ? kc : vc;
              const size_t a = (((size_t)(l * p.slots + slot) * H + (c >> 5)) * p.S + ss.kv[si]) * GSV_HEAD_DIM + (c & 31);
              cache[a] = Elem<T>::from_f(v);
            }
          });
      // ---- P2: attention partials
      grid_sync(p.barrier, epoch);
      attention_phase<T>(p, ss, l, NS);
      // ---- P3: out_proj + residual
      gemv_phase<T, NCH_D>(
          w_o + (size_t)l * d * d, d, xs, ss, p.barrier, epoch,
          [&](int t0, int nt) { stage_attn(xs, p, ss, t0, nt, NS); },
          [&](int row, int si, float r) {
            const int slot = ss.sl[si];
            st_cg(p.y1 + (size_t)slot * d + row,
                  r + Elem<T>::to_f(b_o[(size_t)l * d + row]) + ld_cg(p.xres + (size_t)slot * d + row));
          });
      // ---- P4: LN1 + MLP up + ReLU
      gemv_phase<T, NCH_D>(
          w_1 + (size_t)l * F * d, F, xs, ss, p.barrier, epoch,
          [&](int t0, int nt) {
            stage_ln<T>(xs, p.y1, ln1_g + (size_t)l * d, ln1_b + (size_t)l * d, d, ss, t0, nt, lead ? p.xres1 : nullptr);
          },
          [&](int row, int si, float r) {
            const int slot = ss.sl[si];
            st_cg(p.hbuf + (size_t)slot * F + row, fmaxf(0.f, r + Elem<T>::to_f(b_1[(size_t)l * F + row])));
          });
      // ---- P5: MLP down + residual
      gemv_phase<T, NCH_F>(
          w_2 + (size_t)l * d * F, d, xs, ss, p.barrier, epoch,
          [&](int t0, int nt) { stage_plain(xs, p.hbuf, F, ss, t0, nt, nullptr); },
          [&](int row, int si, float r) {
            const int slot = ss.sl[si];
            st_cg(p.y2 + (size_t)slot * d + row,
                  r + Elem<T>::to_f(b_2[(size_t)l * d + row]) + ld_cg(p.xres1 + (size_t)slot * d + row));
          });
    }
    // ---- P6: logits = LN2_last(y2) . Whead^T ; kv_len += 1 (t2s_model.py:142, 442)
    gemv_phase<T, NCH_D>(
        w_head, p.V, xs, ss, p.barrier, epoch,
        [&](int t0, int nt) {
          stage_ln<T>(xs, p.y2, ln2_g + (size_t)(L - 1) * d, ln2_b + (size_t)(L - 1) * d, d, ss, t0, nt, nullptr);
        },
        [&](int row, int si, float r) { st_cg(p.logits + (size_t)ss.sl[si] * GSV_VOCAB_MAX + row, r); });
    if (lead && threadIdx.x < nb) st_cg(p.kv_len + ss.sl[threadIdx.x], ss.kv[threadIdx.x] + 1);
    // ---- P7: sampling + next input, one CTA per slot
    grid_sync(p.barrier, epoch);
    for (int i = blockIdx.x; i < nb; i += gridDim.x) sample_slot<T>(p, ss.sl[i], xs);
    grid_sync(p.barrier, epoch);
    build_active(p, ss);
  }
}

template <typename T>
int launch_decode(gsv_gpt_ctx* ctx, int n_steps, cudaStream_t st) {
  const int nd = ctx->p.d / 256, nf = ctx->p.F / 256;
  void* fn = nullptr;
  if (nd == 2 && nf == 8) fn = (void*)gpt_decode_kernel<T, 2, 8>;
  else if (nd == 1 && nf == 4) fn = (void*)gpt_decode_kernel<T, 1, 4>;
  else if (nd == 4 && nf == 16) fn = (void*)gpt_decode_kernel<T, 4, 16>;
  else {
    gsv_set_error("decode kernel: unsupported d_model=%d d_ff=%d (supported: 256/1024, 512/2048, 1024/4096)", ctx->p.d, ctx->p.F);
    return GSV_ERR_ARG;
  }
  GSV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->decode_smem));
  if (ctx->decode_grid == 0) {
    int per_sm = 0;
    GSV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, NT, ctx->decode_smem));
    if (per_sm < 1) { gsv_set_error("decode kernel does not fit on an SM"); return GSV_ERR_CUDA; }
    ctx->decode_grid = ctx->num_sms;      // one persistent CTA per SM
  }
  GSV_CUDA(cudaMemsetAsync(ctx->p.barrier, 0, sizeof(unsigned), st));
  GptParams p = ctx->p;
  int ns = n_steps;
  void* args[] = {&p, &ns};
  GSV_CUDA(cudaLaunchCooperativeKernel(fn, dim3(ctx->decode_grid), dim3(NT), args, ctx->decode_smem, st));
  ctx->launches += 1;
  return GSV_OK;
}

}  // namespace

int gsv_gpt_decode_configure(gsv_gpt_ctx* ctx) {
  const int kmax = ctx->p.F > ctx->p.d ? ctx->p.F : ctx->p.d;
  size_t tile = (size_t)TILE * kmax * sizeof(float);
  size_t samp = (size_t)GSV_SAMPLE_SMEM_FLOATS * sizeof(float);
  ctx->decode_smem = tile > samp ? tile : samp;
  ctx->decode_grid = 0;
  return GSV_OK;
}

int gsv_gpt_decode_launch(gsv_gpt_ctx* ctx, int n_steps, cudaStream_t st) {
  if (ctx->dims.dtype == GSV_F16) return launch_decode<__half>(ctx, n_steps, st);
  return launch_decode<__nv_bfloat16>(ctx, n_steps, st);
}
