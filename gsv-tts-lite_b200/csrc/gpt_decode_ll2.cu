// gpt_decode_ll2.cu -- second-generation small-batch (1..4 live sequences) persistent decode kernel.
//
// Same arithmetic as gpt_decode.cu / gpt_decode_ll.cu (reference t2s_model.py:67-105, 129-143, 442-456) and the
// same flag-in-data ("LL") exchange, restructured after the per-phase timeline and ncu source profile of the
// first version (profiles/r01_ll_*): a batch-1 token is ~120 dependent exchanges, and what the first version
// spent between them was not memory time but hundreds of dependent scalar instructions (generic descriptors,
// runtime divisions, two block-wide reductions and 5-6 __syncthreads per phase).  Here:
//
//   * 4 exchanges per layer instead of 5.  The QKV projection of head h and the attention of head h run in the
//     SAME CTA ("attention CTA", one per (live sequence, head)): its 96 weight rows (q_h | k_h | v_h) sit in
//     shared memory (cp.async one layer ahead), so q, k, v never cross CTAs.
//   * Each GEMV phase has ONE __syncthreads: threads poll one word each into shared memory, sync, then every warp
//     reads the 16 elements per lane its dot product needs, computes the LayerNorm statistics redundantly with
//     warp shuffles, normalises in registers and multiplies.
//   * Every CTA owns a fixed, balanced set of output rows, one per warp and phase; the next layer's row is copied
//     with cp.async into the warp's private shared-memory slot right after the current one is consumed.  (Holding
//     it in registers instead was measured 25 % SLOWER: an outstanding HBM load shares one of the warp's six
//     scoreboards with the polling loads, so every spin inherited the HBM latency of the prefetch.)
//   * A CTA's outputs are gathered in shared memory and published by one warp with one coalesced store into a
//     128-byte line that only this CTA writes.
//
// Phases of layer l (tags t..t+3):
//   A   all CTAs: x = l == 0 ? xin : LN2(sum of the 4 y2 partials)  -> residual copy; attention CTAs: q,k,v of
//       their head, KV append, attention over 0..kv -> att[slot][h*32..]
//   O   y1 = x + att Wo^T + bo
//   M1  x1 = LN1(y1) -> residual copy; h = relu(x1 W1^T + b1)
//   M2  y2 = x1 + h W2^T + b2 (four K-quarter warps per row, summed inside the CTA)
// then the head (LN2 + ar_predict_layer) and the sampler CTAs (gpt_sample.cuh).
//
// Buffer-reuse safety (single-buffered exchange areas): every CTA polls the full input of every phase, and every
// CTA PRODUCES in at least one of any three consecutive phases (M1: all CTAs; M2: CTAs [0, D/4); O: the top D/16
// CTAs, which include every CTA without M2 rows).  The writers of an area's next version transitively wait for
// the outputs of those phases, hence for every reader of the previous version.
#include "gpt_sample.cuh"

namespace {

constexpr int NT = GSV_DECODE_THREADS;   // 512
constexpr int NWARP = NT / 32;           // 16
constexpr int MAXB = 4;
constexpr int QKV_ROWS = 3 * GSV_HEAD_DIM;   // 96 rows of Wqkv per head
constexpr int WQ_PAD = 8;                    // elements of padding per staged Wqkv row: conflict-free mma fragment loads

__device__ __forceinline__ int split_pos(int k, int K) {
  const int ch = k >> 3, j = k & 7;
  return j < 4 ? ch * 4 + j : (K >> 1) + ch * 4 + (j - 4);
}
__device__ __forceinline__ unsigned short ld_raw16(const void* p) {
#ifdef GSV_EXP_NOCONST
  return 0x3c00;
#endif
  unsigned short v;
  asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}
template <typename T> __device__ __forceinline__ float raw16_to_f(unsigned short v);
template <> __device__ __forceinline__ float raw16_to_f<__half>(unsigned short v) { return __half2float(__ushort_as_half(v)); }
template <> __device__ __forceinline__ float raw16_to_f<__nv_bfloat16>(unsigned short v) { return __uint_as_float((unsigned)v << 16); }
__device__ __forceinline__ void l2_prefetch(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// per-warp weight slot: lane copies / reads back its own 16-byte chunks (no cross-thread visibility needed)
template <typename T, int NCH>
__device__ __forceinline__ void slot_fetch(uint4* slot, const T* row, int lane) {
  const uint4* src = reinterpret_cast<const uint4*>(row);
#pragma unroll
  for (int c = 0; c < NCH; ++c) cp_async16(slot + c * 32 + lane, src + c * 32 + lane);
}
template <int NCH>
__device__ __forceinline__ void slot_read(const uint4* slot, int lane, uint4 (&w)[NCH]) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) w[c] = slot[c * 32 + lane];
}

__device__ __forceinline__ void mark(const GptParams& p, int id) {
#ifdef GSV_TIMELINE
  if (p.prof != nullptr && threadIdx.x == 0) {
    long long* rec = p.prof + (size_t)blockIdx.x * 2 * p.prof_max;
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

// the 8*NCH elements of a staged vector (split layout, K = 256*NCH) that lane `lane` multiplies
template <int NCH>
__device__ __forceinline__ void load_x(const float* xs, int lane, float (&x)[NCH * 8]) {
  constexpr int K = NCH * 256;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = c * 32 + lane;
    const float4 lo = *reinterpret_cast<const float4*>(xs + ch * 4);
    const float4 hi = *reinterpret_cast<const float4*>(xs + (K >> 1) + ch * 4);
    x[c * 8 + 0] = lo.x; x[c * 8 + 1] = lo.y; x[c * 8 + 2] = lo.z; x[c * 8 + 3] = lo.w;
    x[c * 8 + 4] = hi.x; x[c * 8 + 5] = hi.y; x[c * 8 + 6] = hi.z; x[c * 8 + 7] = hi.w;
  }
}
template <typename T, int NCH>
__device__ __forceinline__ float dot_regs(const uint4 (&w)[NCH], const float (&x)[NCH * 8]) {
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    float wf[8];
    unpack8<T>(w[c], wf);
    a = fmaf(wf[0], x[c * 8 + 0], a); a = fmaf(wf[1], x[c * 8 + 1], a); a = fmaf(wf[2], x[c * 8 + 2], a); a = fmaf(wf[3], x[c * 8 + 3], a);
    b = fmaf(wf[4], x[c * 8 + 4], b); b = fmaf(wf[5], x[c * 8 + 5], b); b = fmaf(wf[6], x[c * 8 + 6], b); b = fmaf(wf[7], x[c * 8 + 7], b);
  }
  return a + b;
}
// LayerNorm statistics of the whole D-vector from the lane's 8*NCH elements (every warp computes them redundantly)
template <int NCH>
__device__ __forceinline__ void ln_stats(const float (&x)[NCH * 8], float& mean, float& rstd) {
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int i = 0; i < NCH * 8; ++i) { s += x[i]; q = fmaf(x[i], x[i], q); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  constexpr float inv = 1.f / (float)(NCH * 256);
  mean = s * inv;
  rstd = rsqrtf(fmaxf(q * inv - mean * mean, 0.f) + 1e-5f);
}
// x <- (x - mean) * rstd * gamma + beta for the lane's elements; gamma/beta rows of D elements (type T)
template <typename T, int NCH>
__device__ __forceinline__ void ln_apply(float (&x)[NCH * 8], float mean, float rstd, const uint4 (&g)[NCH], const uint4 (&b)[NCH]) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    float gf[8], bf[8];
    unpack8<T>(g[c], gf);
    unpack8<T>(b[c], bf);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[c * 8 + j] = fmaf((x[c * 8 + j] - mean) * rstd, gf[j], bf[j]);
  }
}
template <typename T, int NCH>
__device__ __forceinline__ void load_row_regs(const T* row, int lane, uint4 (&w)[NCH]) {
  const uint4* src = reinterpret_cast<const uint4*>(row);
#ifdef GSV_EXP_NOCONST
#pragma unroll
  for (int c = 0; c < NCH; ++c) w[c] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  return;
#endif
#pragma unroll
  for (int c = 0; c < NCH; ++c) w[c] = ld_weight(src + c * 32 + lane);
}

// four words `stride` apart, polled together (one L2 round trip when they have all arrived)
__device__ __forceinline__ void ll_wait4(const uint2* p0, int stride, unsigned tag, float (&out)[4]) {
  uint2 w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) w[e] = make_uint2(0u, ~tag);
  bool ok;
  do {
    ok = true;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (w[e].y != tag) {
        w[e] = ll_peek(p0 + (size_t)e * stride);
        ok = ok && (w[e].y == tag);
      }
    }
  } while (!ok);
#pragma unroll
  for (int e = 0; e < 4; ++e) out[e] = __uint_as_float(w[e].x);
}

// four words at explicit offsets, polled together
__device__ __forceinline__ void ll_wait4p(const uint2* base, const int (&off)[4], unsigned tag, float (&out)[4]) {
  uint2 w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) w[e] = make_uint2(0u, ~tag);
  bool ok;
  do {
    ok = true;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (w[e].y != tag) {
        w[e] = ll_peek(base + off[e]);
        ok = ok && (w[e].y == tag);
      }
    }
  } while (!ok);
#pragma unroll
  for (int e = 0; e < 4; ++e) out[e] = __uint_as_float(w[e].x);
}

// m16n8k16 tensor-core tile, 16-bit inputs, fp32 accumulate: D = A(16x16, row) * B(16x8, col) + C
template <typename T> struct Mma16816;
template <> struct Mma16816<__nv_bfloat16> {
  static __device__ __forceinline__ void run(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
};
template <> struct Mma16816<__half> {
  static __device__ __forceinline__ void run(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
};

// Experiment (GSV_LL2_WARPPOLL=n): only n warps of a CTA poll an exchanged D-vector (D/n/32 words per lane, all in
// flight, re-polling only what is missing) and scatter it into shared memory; the other warps wait at the phase's
// __syncthreads.  Fewer polling warps = fewer requests queued on the lines being written.
#ifdef GSV_LL2_WARPPOLL
template <int D>
__device__ __forceinline__ void poll_vec(const uint2* src, unsigned tag, float* dst) {
  constexpr int PW = GSV_LL2_WARPPOLL, PER = D / PW / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= PW) return;
  uint2 w[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) w[i] = make_uint2(0u, ~tag);
  bool ok;
  do {
    ok = true;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      if (w[i].y != tag) {
        w[i] = ll_peek(src + warp * (D / PW) + lane + 32 * i);
        ok = ok && (w[i].y == tag);
      }
    }
  } while (!ok);
#pragma unroll
  for (int i = 0; i < PER; ++i) dst[split_pos(warp * (D / PW) + lane + 32 * i, D)] = __uint_as_float(w[i].x);
}
#endif

struct L2Shared {
  int sl[MAXB], kv[MAXB], nb;
  float q[GSV_HEAD_DIM], knew[GSV_HEAD_DIM], vnew[GSV_HEAD_DIM];
  float wpart[NWARP][GSV_HEAD_DIM + 2];
  float wscale[NWARP];
  float outv[MAXB][NWARP];      // this phase's outputs, gathered for one coalesced publication by warp 0
  float qkv_part[2][QKV_ROWS];  // attention CTAs: K-half partial sums of the tensor-core QKV projection
};

template <typename T, int NCH, int NB>
__global__ void __launch_bounds__(NT, 1) gpt_decode_ll2_kernel(const GptParams p, const int n_steps, const unsigned tag_base,
                                                               uint2* const ll_buf) {
  extern __shared__ __align__(16) float smem[];
  __shared__ L2Shared sh;
  constexpr int D = NCH * 256, F = 4 * D;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  const int H = p.H, L = p.L, V = p.V, S = p.S;
  // shared memory: xa[NB][D] | xb[NB][F] staged operands (split layout per D-wide piece), used by alternating phases
  // (A, M1, head: xa; O, M2: xb) so that the single __syncthreads of a phase also orders the previous phase's
  // reads against the next phase's writes | xres[NB][D] | xres1[NB][D] | wq[96][D] (T)
  float* xa = smem;
  float* xb = xa + NB * D;
  float* xres = xb + NB * F;
  float* xres1 = xres + NB * D;
  constexpr int kAct = NB * F + 3 * NB * D;
  constexpr int kScratch = kAct > GSV_SAMPLE_SMEM_FLOATS ? kAct : GSV_SAMPLE_SMEM_FLOATS;
  T* wq = reinterpret_cast<T*>(smem + ((kScratch + 3) & ~3));
  // per-warp weight slots [3 units][NWARP][D/8] uint4 (one D-wide row segment each)
  constexpr int LDW = D + WQ_PAD;
  T* xbf = wq + (size_t)QKV_ROWS * LDW;                      // the attention CTA's normalised input in the storage type [D]
  uint4* wslot = reinterpret_cast<uint4*>(xbf + D);
  uint4* const slot_o = wslot + (size_t)(0 * NWARP + warp) * (D / 8);
  uint4* const slot_1 = wslot + (size_t)(1 * NWARP + warp) * (D / 8);
  uint4* const slot_2 = wslot + (size_t)(2 * NWARP + warp) * (D / 8);
  // exchange areas ({value, tag} words), per slot
  const size_t slots = p.slots;
  uint2* ll_xin = ll_buf;                                   // [slots][D]
  uint2* ll_att = ll_xin + slots * D;                       // [slots][D]
  uint2* ll_y1 = ll_att + slots * D;                        // [slots][D]
  const size_t padw = (size_t)gridDim.x * 16;
  uint2* ll_h = ll_y1 + slots * D;                          // [slots][G][16]  MLP-up outputs, one line per producer CTA
  uint2* ll_y2 = ll_h + slots * padw;                       // [slots][G][16]  MLP-down partials (task = quarter*D + row)
  uint2* ll_logit = ll_y2 + slots * padw;                   // [slots][VOCAB_MAX]
  uint2* ll_stat = ll_logit + slots * GSV_VOCAB_MAX;        // [slots]

  const T* const Wqkv = reinterpret_cast<const T*>(p.w_qkv);
  const T* const Wo = reinterpret_cast<const T*>(p.w_o);
  const T* const W1 = reinterpret_cast<const T*>(p.w_1);
  const T* const W2 = reinterpret_cast<const T*>(p.w_2);
  const T* const Wh = reinterpret_cast<const T*>(p.w_head);
  const T* const Bqkv = reinterpret_cast<const T*>(p.b_qkv);
  const T* const Bo = reinterpret_cast<const T*>(p.b_o);
  const T* const B1 = reinterpret_cast<const T*>(p.b_1);
  const T* const B2 = reinterpret_cast<const T*>(p.b_2);
  const T* const G1 = reinterpret_cast<const T*>(p.ln1_g);
  const T* const Be1 = reinterpret_cast<const T*>(p.ln1_b);
  const T* const G2 = reinterpret_cast<const T*>(p.ln2_g);
  const T* const Be2 = reinterpret_cast<const T*>(p.ln2_b);

  // ---- this CTA's fixed share of every GEMV phase: tasks [c*N/G, (c+1)*N/G), warp w takes task begin + w (+16 j) ----
  // Every 128-byte line of an exchange area has exactly ONE writer CTA, which publishes it with one coalesced
  // warp store: measured, a line written word by word from several SMs while ~150 CTAs poll it takes 3-5 us to
  // settle, a single-writer line 0.4 us.
  //   O, head : 16 consecutive rows per producer CTA (D/16 resp. ceil(V/16) CTAs produce; the rest only consume)
  //   M1, M2  : all G CTAs produce (needed for the buffer-reuse argument above): CTA c owns tasks
  //             [c*F/G, (c+1)*F/G) (<= 16) and publishes them into ITS line: word index c*16 + (task - begin)
  const int o_cta = cta - (G - D / 16);                     // O producers: the top D/16 CTAs (they include every CTA without MLP-down rows)
  const int o_b = (o_cta >= 0 && o_cta < D / 16) ? o_cta * 16 : 0, o_e = (o_cta >= 0 && o_cta < D / 16) ? o_b + 16 : 0;
  const int m1_b = (int)((long long)cta * F / G), m1_e = (int)((long long)(cta + 1) * F / G);
  // MLP-down: CTA c < D/4 owns rows [4c, 4c+4); warp w = quarter (w & 3) of row 4c + (w >> 2); the four K-quarters are
  // summed inside the CTA, so consumers poll ONE word per element (4-word polls made this exchange 2x slower)
  const int hd_b = min(V, cta * 16), hd_e = min(V, cta * 16 + 16);
  const int o_t = o_b + warp, m1_t = m1_b + warp;
  const bool o_ok = o_t < o_e, m1_ok = m1_t < m1_e, m2_ok = cta < D / 4;
  const int m2_q = warp & 3, m2_row = cta * 4 + (warp >> 2);
#ifdef GSV_LL2_PADDED
  const int y2_off = (tid >> 2) * 16 + (tid & 3);           // where element `tid` of y2 lives: line of CTA tid/4, word tid%4
#else
  const int y2_off = tid;
#endif
  // where this thread finds element `tid` of the h vector / of each y2 partial in the padded layout
  int pad_off[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int t = e * D + tid;                              // task index (valid for tid < D)
    const int owner = (int)((((long long)t + 1) * G - 1) / F);
    (void)owner;
    pad_off[e] = t;                                         // dense layout (GSV_LL2_PADDED: one line per producer CTA)
#ifdef GSV_LL2_PADDED
    pad_off[e] = owner * 16 + (t - (int)((long long)owner * F / G));
#endif
  }
  const int PADW = G * 16;                                  // words per slot of a padded area

  // ---- launch prologue: live slots ----
  if (tid < 32) {
    const int flag = tid < p.slots ? ld_cg(p.active + tid) : 0;
    const unsigned m = __ballot_sync(0xffffffffu, flag != 0);
    if (flag) {
      const int pos = __popc(m & ((1u << tid) - 1u));
      if (pos < NB) { sh.sl[pos] = tid; sh.kv[pos] = ld_cg(p.kv_len + tid); }
    }
    if (tid == 0) sh.nb = min(NB, __popc(m));
  }
  __syncthreads();
  // attention role: CTA a < nb*H serves (slot index a / H, head a % H)
  const int att_s = cta / H, att_h = cta - att_s * H;
  const bool att_role = cta < NB * H;

  // Layer-0 rows in flight.  Every thread commits exactly one cp.async group per phase, in use order
  // (A: Wqkv rows of the attention CTAs, else empty; O; M1; M2), so "all but the 3 newest groups complete"
  // always means "the group of the phase being entered has landed".
  if (att_role) {
    for (int i = tid; i < QKV_ROWS * (D / 8); i += NT) {
      const int rr = i / (D / 8), c = i - rr * (D / 8);
      const int grow = (rr >> 5) * D + att_h * GSV_HEAD_DIM + (rr & 31);
      cp_async16(wq + (size_t)rr * LDW + c * 8, Wqkv + (size_t)grow * D + c * 8);
    }
  }
  cp_async_commit();
  if (o_ok) slot_fetch<T, NCH>(slot_o, Wo + (size_t)o_t * D, lane);
  cp_async_commit();
  if (m1_ok) slot_fetch<T, NCH>(slot_1, W1 + (size_t)m1_t * D, lane);
  cp_async_commit();
  if (m2_ok) slot_fetch<T, NCH>(slot_2, W2 + (size_t)m2_row * F + (size_t)m2_q * D, lane);
  cp_async_commit();

  unsigned tag = tag_base;                                  // tag of the most recent publication (xin)
  if (cta == 0) {
    // xin of every live slot was left as plain fp32 by prefill / the previous launch
    for (int i = tid; i < sh.nb * D; i += NT) {
      const int s = i / D, k = i - s * D;
      ll_store(ll_xin + (size_t)sh.sl[s] * D + k, ld_cg(p.xin + (size_t)sh.sl[s] * D + k), tag);
    }
  }

#pragma unroll 1
  for (int step = 0; step < n_steps; ++step) {
    const int nb = sh.nb;
    if (nb == 0) break;
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      const int ln = l == L - 1 ? 0 : l + 1;                // layer whose rows are fetched next (wraps to the next token)
      // ================= phase A: layer input (+LN2 of the previous layer), QKV + attention ==================
      mark(p, 1);
      {
        uint4 g[NCH], be[NCH];
        unsigned short gr = 0, br = 0;
        if (l > 0) {
          load_row_regs<T, NCH>(G2 + (size_t)(l - 1) * D, lane, g);
          load_row_regs<T, NCH>(Be2 + (size_t)(l - 1) * D, lane, be);
          if (tid < D) { gr = ld_raw16(G2 + (size_t)(l - 1) * D + tid); br = ld_raw16(Be2 + (size_t)(l - 1) * D + tid); }
        }
        // attention CTAs: first K/V rows of their stream, requested before the spin
        const bool att = att_role && att_s < nb;
        const int aslot = att ? sh.sl[att_s] : 0;
        const int kvn = att ? sh.kv[att_s] : 0;
        const int sub = lane & 3, pg = lane >> 2;
        const size_t head_base = ((size_t)(l * p.slots + aslot) * H + att_h) * (size_t)S * GSV_HEAD_DIM;
        const T* kb = reinterpret_cast<const T*>(p.kc) + head_base + sub * 8;
        const T* vb = reinterpret_cast<const T*>(p.vc) + head_base + sub * 8;
        uint4 kr0 = make_uint4(0, 0, 0, 0), vr0 = kr0, kr1 = kr0, vr1 = kr0;
        uint4 kr2 = kr0, vr2 = kr0;
        unsigned short bq = 0;                              // thread r < 96: bias of staged row r (q | k | v of this head)
        if (att) {
          if (tid < QKV_ROWS) bq = ld_raw16(Bqkv + (size_t)l * 3 * D + (tid >> 5) * D + att_h * GSV_HEAD_DIM + (tid & 31));
          const int p0 = warp * 8 + pg, p1 = p0 + NWARP * 8, p2 = p1 + NWARP * 8;
          if (p0 < kvn) { kr0 = ld_cg16(kb + (size_t)p0 * GSV_HEAD_DIM); vr0 = ld_cg16(vb + (size_t)p0 * GSV_HEAD_DIM); }
          if (p1 < kvn) { kr1 = ld_cg16(kb + (size_t)p1 * GSV_HEAD_DIM); vr1 = ld_cg16(vb + (size_t)p1 * GSV_HEAD_DIM); }
          if (p2 < kvn) { kr2 = ld_cg16(kb + (size_t)p2 * GSV_HEAD_DIM); vr2 = ld_cg16(vb + (size_t)p2 * GSV_HEAD_DIM); }
          if (tid == 0 && l + 1 < L && kvn > 0) {
            // next layer's K/V stream of this (slot, head) into L2
            const size_t nxt = (size_t)p.slots * H * S * GSV_HEAD_DIM;
            const unsigned bytes = (unsigned)kvn * GSV_HEAD_DIM * (unsigned)sizeof(T);
            l2_prefetch(reinterpret_cast<const T*>(p.kc) + head_base + nxt, bytes);
            l2_prefetch(reinterpret_cast<const T*>(p.vc) + head_base + nxt, bytes);
          }
        }
        if (warp == NWARP - 1 && lane == 0 && L > 1) {
          // small per-layer vectors of the next layer into L2 (they were evicted by the 152 MB weight stream)
          l2_prefetch(G2 + (size_t)l * D, D * (unsigned)sizeof(T));
          l2_prefetch(Be2 + (size_t)l * D, D * (unsigned)sizeof(T));
          l2_prefetch(G1 + (size_t)ln * D, D * (unsigned)sizeof(T));
          l2_prefetch(Be1 + (size_t)ln * D, D * (unsigned)sizeof(T));
          l2_prefetch(Bqkv + (size_t)ln * 3 * D, 3 * D * (unsigned)sizeof(T));
          l2_prefetch(Bo + (size_t)ln * D, D * (unsigned)sizeof(T));
          l2_prefetch(B1 + (size_t)ln * F, F * (unsigned)sizeof(T));
          l2_prefetch(B2 + (size_t)ln * D, D * (unsigned)sizeof(T));
        }
        // ---- poll the layer input: xin (l == 0) or the 4 MLP-down partials, summed in fixed order
        float raw[NB];
#pragma unroll
        for (int s = 0; s < NB; ++s) {
          raw[s] = 0.f;
#ifdef GSV_LL2_WARPPOLL
          if (s < nb) poll_vec<D>(l == 0 ? ll_xin + (size_t)sh.sl[s] * D : ll_y2 + (size_t)sh.sl[s] * PADW, tag, xa + s * D);
#else
          if (s < nb && tid < D) {
            if (l == 0) {
              raw[s] = ll_wait(ll_xin + (size_t)sh.sl[s] * D + tid, tag);
            } else {
              raw[s] = ll_wait(ll_y2 + (size_t)sh.sl[s] * PADW + y2_off, tag);
            }
            xa[s * D + split_pos(tid, D)] = raw[s];
          }
#endif
        }
        cp_async_wait<3>();                                 // this layer's Wqkv rows (requested one layer ago)
        __syncthreads();
#ifdef GSV_LL2_WARPPOLL
#pragma unroll
        for (int s = 0; s < NB; ++s)
          if (s < nb && tid < D) raw[s] = xa[s * D + split_pos(tid, D)];
#endif
        mark(p, 40);
        // ---- residual copy x = LN2(raw) (or raw for layer 0): every warp derives the statistics itself
        float xq[NCH * 8];                                  // attention CTAs keep their slot's normalised operand
#pragma unroll
        for (int s = 0; s < NB; ++s) {
          if (s < nb) {
            float xv[NCH * 8];
            load_x<NCH>(xa + s * D, lane, xv);
            float mean = 0.f, rstd = 1.f;
            if (l > 0) {
              ln_stats<NCH>(xv, mean, rstd);
              ln_apply<T, NCH>(xv, mean, rstd, g, be);
              if (tid < D) xres[s * D + tid] = fmaf((raw[s] - mean) * rstd, raw16_to_f<T>(gr), raw16_to_f<T>(br));
            } else if (tid < D) {
              xres[s * D + tid] = raw[s];
            }
            if (att && s == att_s) {
#pragma unroll
              for (int i = 0; i < NCH * 8; ++i) xq[i] = xv[i];
            }
          }
        }
        if (att) {
          // ---- q, k, v of this head on the tensor cores: [96 x D] staged weight rows times the normalised input
          //      (rounded to the storage type, as the reference's LayerNorm output is).  Warp w < 12: 16-row tile
          //      w % 6, K half w / 6; the input sits in column 0 of the 16x8 B operand (lanes 0..3).
          if (warp == 0) {
#pragma unroll
            for (int c = 0; c < NCH; ++c)
              *reinterpret_cast<uint4*>(xbf + c * 256 + lane * 8) = pack8<T>(&xq[c * 8]);
          }
          __syncthreads();
          if (warp < 12) {
            const int mt = warp % 6, kh = warp / 6;
            const int g = lane >> 2, t = lane & 3;
            const T* a_lo = wq + (size_t)(16 * mt + g) * LDW + 2 * t;
            const T* a_hi = a_lo + 8 * LDW;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
            for (int ks = 0; ks < D / 32; ++ks) {
              const int k0 = kh * (D / 2) + ks * 16;
              const unsigned a0 = *reinterpret_cast<const unsigned*>(a_lo + k0), a1 = *reinterpret_cast<const unsigned*>(a_hi + k0);
              const unsigned a2 = *reinterpret_cast<const unsigned*>(a_lo + k0 + 8), a3 = *reinterpret_cast<const unsigned*>(a_hi + k0 + 8);
              unsigned b0 = 0u, b1 = 0u;
              if (g == 0) {
                b0 = *reinterpret_cast<const unsigned*>(xbf + k0 + 2 * t);
                b1 = *reinterpret_cast<const unsigned*>(xbf + k0 + 8 + 2 * t);
              }
              Mma16816<T>::run(acc, a0, a1, a2, a3, b0, b1);
            }
            if (t == 0) {                                  // column 0 of the accumulator tile: rows g and g + 8
              sh.qkv_part[kh][16 * mt + g] = acc[0];
              sh.qkv_part[kh][16 * mt + g + 8] = acc[2];
            }
          }
          __syncthreads();
          if (tid < QKV_ROWS) {
            const int which = tid >> 5, c = tid & 31;
            const float v = (sh.qkv_part[0][tid] + sh.qkv_part[1][tid]) + raw16_to_f<T>(bq);
            if (which == 0) {
              sh.q[c] = v * (rsqrtf((float)GSV_HEAD_DIM) * 1.4426950408889634f);
            } else {
              // the reference attends over the 16-bit cache entry it has just written
              const T t16 = Elem<T>::from_f(v);
              (which == 1 ? sh.knew : sh.vnew)[c] = Elem<T>::to_f(t16);
              T* cache = reinterpret_cast<T*>(which == 1 ? p.kc : p.vc);
              cache[head_base + (size_t)kvn * GSV_HEAD_DIM + c] = t16;
            }
          }
          __syncthreads();
          mark(p, 50);
          // Wqkv rows of the next layer (shared-memory rows are free now)
          {
            const T* src = Wqkv + (size_t)ln * 3 * D * D;
            for (int i = tid; i < QKV_ROWS * (D / 8); i += NT) {
              const int rr = i / (D / 8), c = i - rr * (D / 8);
              const int grow = (rr >> 5) * D + att_h * GSV_HEAD_DIM + (rr & 31);
              cp_async16(wq + (size_t)rr * LDW + c * 8, src + (size_t)grow * D + c * 8);
            }
          }
          // ---- attention over cached positions [0, kvn) + the new one; 4 lanes per position, 8 positions per warp pass
          float q[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) q[j] = sh.q[sub * 8 + j];
          float mg = GSV_NEG_INF, lsum = 0.f, o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = 0.f;
          int pass = 0;
#pragma unroll 1
          for (int base = warp * 8; base < kvn; base += NWARP * 8, ++pass) {
            const int pos = base + pg;
            const bool ok = pos < kvn;
            uint4 kr = pass == 0 ? kr0 : (pass == 1 ? kr1 : kr2), vr = pass == 0 ? vr0 : (pass == 1 ? vr1 : vr2);
            if (pass > 2 && ok) {
              kr = ld_cg16(kb + (size_t)pos * GSV_HEAD_DIM);
              vr = ld_cg16(vb + (size_t)pos * GSV_HEAD_DIM);
            }
            float kf[8], vf[8], sc_ = 0.f;
            unpack8<T>(kr, kf);
            unpack8<T>(vr, vf);
#pragma unroll
            for (int j = 0; j < 8; ++j) sc_ = fmaf(q[j], kf[j], sc_);
            sc_ += __shfl_xor_sync(0xffffffffu, sc_, 1);
            sc_ += __shfl_xor_sync(0xffffffffu, sc_, 2);
            if (ok) {
              const float mn = fmaxf(mg, sc_);
              const float sc = exp2f(mg - mn), pr = exp2f(sc_ - mn);
              lsum = fmaf(lsum, sc, pr);
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = fmaf(pr, vf[j], o[j] * sc);
              mg = mn;
            }
          }
          mark(p, 51);
          float m = mg;
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
          const float rs = (mg > GSV_NEG_INF) ? exp2f(mg - m) : 0.f;
          lsum *= rs;
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] *= rs;
#pragma unroll
          for (int off = 4; off < 32; off <<= 1) {
            lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += __shfl_xor_sync(0xffffffffu, o[j], off);
          }
          if (lane < 4) {
            if (sub == 0) { sh.wpart[warp][0] = m; sh.wpart[warp][1] = lsum; }
#pragma unroll
            for (int j = 0; j < 8; ++j) sh.wpart[warp][2 + sub * 8 + j] = o[j];
          }
          __syncthreads();
          if (warp == 0) {
            const float mw = lane < NWARP ? sh.wpart[lane][0] : GSV_NEG_INF;
            const float snew = warp_sum(sh.q[lane] * sh.knew[lane]);
            const float M = fmaxf(warp_max(mw), snew);
            if (lane < NWARP) sh.wscale[lane] = mw > GSV_NEG_INF ? exp2f(mw - M) : 0.f;
            __syncwarp();
            float Ls = 0.f, Ls2 = 0.f, oa = 0.f, ob = 0.f;
#pragma unroll
            for (int w = 0; w < NWARP; w += 2) {
              const float sc0 = sh.wscale[w], sc1 = sh.wscale[w + 1];
              Ls = fmaf(sh.wpart[w][1], sc0, Ls);
              Ls2 = fmaf(sh.wpart[w + 1][1], sc1, Ls2);
              oa = fmaf(sh.wpart[w][2 + lane], sc0, oa);
              ob = fmaf(sh.wpart[w + 1][2 + lane], sc1, ob);
            }
            Ls += Ls2; oa += ob;
            const float pr = exp2f(snew - M);
            Ls += pr;
            oa = fmaf(pr, sh.vnew[lane], oa);
            ll_store(ll_att + (size_t)aslot * D + att_h * GSV_HEAD_DIM + lane, oa / Ls, tag + 1);
          }
          mark(p, 52);
        }
        cp_async_commit();                                  // phase-A group (empty unless this CTA refilled its Wqkv rows)
      }
      tag += 1;
      // ================= phase O: y1 = x + att Wo^T + bo ==================
      mark(p, 3);
      {
        unsigned short braw = 0;
        if (o_ok && lane == 0) braw = ld_raw16(Bo + (size_t)l * D + o_t);
#pragma unroll
        for (int s = 0; s < NB; ++s)
#ifdef GSV_LL2_WARPPOLL
          if (s < nb) poll_vec<D>(ll_att + (size_t)sh.sl[s] * D, tag, xb + s * F);
#else
          if (s < nb && tid < D) xb[s * F + split_pos(tid, D)] = ll_wait(ll_att + (size_t)sh.sl[s] * D + tid, tag);
#endif
        __syncthreads();
        mark(p, 41);
        cp_async_wait<3>();
        if (o_ok) {
          uint4 w_o[NCH];
          slot_read<NCH>(slot_o, lane, w_o);
#pragma unroll
          for (int s = 0; s < NB; ++s) {
            if (s < nb) {
              float xv[NCH * 8];
              load_x<NCH>(xb + s * F, lane, xv);
              float a = dot_regs<T, NCH>(w_o, xv);
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
              if (lane == 0) sh.outv[s][warp] = a + raw16_to_f<T>(braw) + xres[s * D + o_t];
            }
          }
        }
        __syncthreads();
        if (warp == 0 && o_b + lane < o_e) {
#pragma unroll
          for (int s = 0; s < NB; ++s)
            if (s < nb) ll_store(ll_y1 + (size_t)sh.sl[s] * D + o_b + lane, sh.outv[s][lane], tag + 1);
        }
        if (o_ok) slot_fetch<T, NCH>(slot_o, Wo + ((size_t)ln * D + o_t) * D, lane);      // next layer's row, in flight until then
        cp_async_commit();
      }
      tag += 1;
      // ================= phase M1: x1 = LN1(y1); h = relu(x1 W1^T + b1) ==================
      mark(p, 4);
      {
        uint4 g[NCH], be[NCH];
        load_row_regs<T, NCH>(G1 + (size_t)l * D, lane, g);
        load_row_regs<T, NCH>(Be1 + (size_t)l * D, lane, be);
        unsigned short gr = 0, br = 0, braw = 0;
        if (tid < D) { gr = ld_raw16(G1 + (size_t)l * D + tid); br = ld_raw16(Be1 + (size_t)l * D + tid); }
        if (m1_ok && lane == 0) braw = ld_raw16(B1 + (size_t)l * F + m1_t);
        float raw[NB];
#pragma unroll
        for (int s = 0; s < NB; ++s) {
          raw[s] = 0.f;
#ifdef GSV_LL2_WARPPOLL
          if (s < nb) poll_vec<D>(ll_y1 + (size_t)sh.sl[s] * D, tag, xa + s * D);
#else
          if (s < nb && tid < D) {
            raw[s] = ll_wait(ll_y1 + (size_t)sh.sl[s] * D + tid, tag);
            xa[s * D + split_pos(tid, D)] = raw[s];
          }
#endif
        }
        __syncthreads();
#ifdef GSV_LL2_WARPPOLL
#pragma unroll
        for (int s = 0; s < NB; ++s)
          if (s < nb && tid < D) raw[s] = xa[s * D + split_pos(tid, D)];
#endif
        mark(p, 42);
        cp_async_wait<3>();
        uint4 w_1[NCH];
        if (m1_ok) slot_read<NCH>(slot_1, lane, w_1);
#pragma unroll
        for (int s = 0; s < NB; ++s) {
          if (s < nb) {
            float xn[NCH * 8];
            load_x<NCH>(xa + s * D, lane, xn);
            float mean, rstd;
            ln_stats<NCH>(xn, mean, rstd);
            ln_apply<T, NCH>(xn, mean, rstd, g, be);
            if (tid < D) xres1[s * D + tid] = fmaf((raw[s] - mean) * rstd, raw16_to_f<T>(gr), raw16_to_f<T>(br));
            if (m1_ok) {
              float a = dot_regs<T, NCH>(w_1, xn);
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
              if (lane == 0) sh.outv[s][warp] = fmaxf(a + raw16_to_f<T>(braw), 0.f);
            }
          }
        }
        __syncthreads();
        if (warp == 0 && m1_b + lane < m1_e) {
#pragma unroll
          for (int s = 0; s < NB; ++s)
#ifdef GSV_LL2_PADDED
            if (s < nb) ll_store(ll_h + (size_t)sh.sl[s] * PADW + cta * 16 + lane, sh.outv[s][lane], tag + 1);
#else
            if (s < nb) ll_store(ll_h + (size_t)sh.sl[s] * PADW + m1_b + lane, sh.outv[s][lane], tag + 1);
#endif
        }
        if (m1_ok) slot_fetch<T, NCH>(slot_1, W1 + ((size_t)ln * F + m1_t) * D, lane);
        cp_async_commit();
      }
      tag += 1;
      // ================= phase M2: y2 partial[q] = h[q] W2[:, q]^T (+ b2 + x1 for q == 0) ==================
      mark(p, 5);
      {
        unsigned short braw = 0;
        if (m2_ok && lane == 0 && m2_q == 0) braw = ld_raw16(B2 + (size_t)l * D + m2_row);
#pragma unroll
        for (int s = 0; s < NB; ++s) {
          if (s < nb) {
            // F = 4 D elements: thread t takes element t of every D-quarter (tid < D)
            if (tid < D) {
              float hv[4];
              ll_wait4p(ll_h + (size_t)sh.sl[s] * PADW, pad_off, tag, hv);
#pragma unroll
              for (int e = 0; e < 4; ++e) xb[s * F + e * D + split_pos(tid, D)] = hv[e];
            }
          }
        }
        __syncthreads();
        mark(p, 43);
        cp_async_wait<3>();
        if (m2_ok) {
          uint4 w_2[NCH];
          slot_read<NCH>(slot_2, lane, w_2);
#pragma unroll
          for (int s = 0; s < NB; ++s) {
            if (s < nb) {
              float xv[NCH * 8];
              load_x<NCH>(xb + s * F + m2_q * D, lane, xv);
              float a = dot_regs<T, NCH>(w_2, xv);
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
              if (lane == 0) {
                if (m2_q == 0) a += raw16_to_f<T>(braw) + xres1[s * D + m2_row];
                sh.outv[s][warp] = a;
              }
            }
          }
        }
        __syncthreads();
        if (warp == 0 && m2_ok && lane < 4) {             // rows 4c .. 4c+3: sum of the four K-quarters, one 32-byte store
#pragma unroll
          for (int s = 0; s < NB; ++s)
            if (s < nb) {
              const float* q4 = &sh.outv[s][lane * 4];
#ifdef GSV_LL2_PADDED
              ll_store(ll_y2 + (size_t)sh.sl[s] * PADW + cta * 16 + lane, ((q4[0] + q4[1]) + q4[2]) + q4[3], tag + 1);
#else
              ll_store(ll_y2 + (size_t)sh.sl[s] * PADW + cta * 4 + lane, ((q4[0] + q4[1]) + q4[2]) + q4[3], tag + 1);
#endif
            }
        }
        if (m2_ok) slot_fetch<T, NCH>(slot_2, W2 + ((size_t)ln * D + m2_row) * F + (size_t)m2_q * D, lane);
        cp_async_commit();
      }
      tag += 1;
    }
    // ================= head: logits = LN2_last(y2) Whead^T ==================
    mark(p, 6);
    {
      uint4 g[NCH], be[NCH], wh[NCH];
      load_row_regs<T, NCH>(G2 + (size_t)(L - 1) * D, lane, g);
      load_row_regs<T, NCH>(Be2 + (size_t)(L - 1) * D, lane, be);
      const int h_t = hd_b + warp;
      if (h_t < hd_e) load_row_regs<T, NCH>(Wh + (size_t)h_t * D, lane, wh);
#pragma unroll
      for (int s = 0; s < NB; ++s) {
        if (s < nb && tid < D) {
          xa[s * D + split_pos(tid, D)] = ll_wait(ll_y2 + (size_t)sh.sl[s] * PADW + y2_off, tag);
        }
      }
      __syncthreads();
      mark(p, 44);
#pragma unroll
      for (int s = 0; s < NB; ++s) {
        if (s < nb) {
          float xn[NCH * 8];
          load_x<NCH>(xa + s * D, lane, xn);
          float mean, rstd;
          ln_stats<NCH>(xn, mean, rstd);
          ln_apply<T, NCH>(xn, mean, rstd, g, be);
          if (h_t < hd_e) {
            float a = dot_regs<T, NCH>(wh, xn);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) sh.outv[s][warp] = a;
          }
        }
      }
      __syncthreads();
      if (warp == 0 && hd_b + lane < hd_e) {
#pragma unroll
        for (int s = 0; s < NB; ++s)
          if (s < nb) ll_store(ll_logit + (size_t)sh.sl[s] * GSV_VOCAB_MAX + hd_b + lane, sh.outv[s][lane], tag + 1);
      }
    }
    tag += 1;
    // ================= sampling: the LAST nb CTAs take one slot each; everyone then learns who is still alive ==========
    {
      const unsigned tag_logits = tag;
      tag += 1;                                   // tag of xin / status published by the samplers
      mark(p, 20);
      __syncthreads();
      const int samp_i = G - 1 - cta;
      if (samp_i < nb) {
        const int slot = sh.sl[samp_i];
        for (int v = tid; v < V; v += NT) smem[v] = ll_wait(ll_logit + (size_t)slot * GSV_VOCAB_MAX + v, tag_logits);
        __syncthreads();
        SampleLL io;
        io.preloaded = true;
        io.xin_ll = ll_xin + (size_t)slot * D;
        io.status_ll = ll_stat + slot;
        io.tag = tag;
        io.kv_len = sh.kv[samp_i] + 1;
        io.xin_smem = nullptr;
        io.alive_smem = nullptr;
        sample_slot<T>(p, slot, smem, &io);
      }
      __syncthreads();
      if (tid == 0) {
        int n2 = 0;
        int sl2[MAXB], kv2[MAXB];
        for (int s = 0; s < nb; ++s) {
          const float alive = ll_wait(ll_stat + sh.sl[s], tag);
          if (alive != 0.f) { sl2[n2] = sh.sl[s]; kv2[n2] = sh.kv[s] + 1; ++n2; }
        }
        for (int s = 0; s < n2; ++s) { sh.sl[s] = sl2[s]; sh.kv[s] = kv2[s]; }
        sh.nb = n2;
      }
      __syncthreads();
      mark(p, 21);
    }
  }
  cp_async_wait_all();
}

template <typename T, int NB>
int launch_ll2_nb(gsv_gpt_ctx* ctx, int n_steps, cudaStream_t st) {
  const int nd = ctx->p.d / 256;
  void* fn = nullptr;
  if (nd == 2) fn = (void*)gpt_decode_ll2_kernel<T, 2, NB>;
  else if (nd == 1) fn = (void*)gpt_decode_ll2_kernel<T, 1, NB>;
  else return GSV_ERR_ARG;
  const size_t act = (size_t)(NB * ctx->p.F + 3 * NB * ctx->p.d);
  const size_t scratch = act > (size_t)GSV_SAMPLE_SMEM_FLOATS ? act : (size_t)GSV_SAMPLE_SMEM_FLOATS;
  const size_t bytes = ((scratch + 3) & ~(size_t)3) * sizeof(float) + (size_t)QKV_ROWS * (ctx->p.d + WQ_PAD) * 2 + (size_t)ctx->p.d * 2 +
                       (size_t)3 * NWARP * ctx->p.d * 2;       // + Wqkv rows + per-warp weight slots
  GSV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  GptParams p = ctx->p;
  int ns = n_steps;
  ctx->ll_seq += 1;
  if ((ctx->ll_seq & 0xffffull) == 0) {
    // the 16-bit launch sequence wraps every 65 536 launches: clear every cell (a rarely written one could still hold
    // a tag of the previous lap) and skip sequence 0, whose tags equal the cleared value
    GSV_CUDA(cudaMemsetAsync(ctx->ll_buf, 0, gsv_gpt_ll_buffer_bytes(ctx), st));
    ctx->ll_seq += 1;
  }
  unsigned tag_base = (unsigned)(ctx->ll_seq << 16);      // tags never repeat between launches
  uint2* buf = reinterpret_cast<uint2*>(ctx->ll_buf);
  void* args[] = {&p, &ns, &tag_base, &buf};
  GSV_CUDA(cudaLaunchCooperativeKernel(fn, dim3(ctx->num_sms), dim3(NT), args, bytes, st));
  ctx->launches += 1;
  return GSV_OK;
}

template <typename T>
int launch_ll2(gsv_gpt_ctx* ctx, int live, int n_steps, cudaStream_t st) {
  if (live <= 1) return launch_ll2_nb<T, 1>(ctx, n_steps, st);
  if (live <= 2) return launch_ll2_nb<T, 2>(ctx, n_steps, st);
  return launch_ll2_nb<T, 4>(ctx, n_steps, st);
}

}  // namespace

static size_t gsv_gpt_ll2_words_per_slot(const gsv_gpt_ctx* ctx) {
  return 3 * (size_t)ctx->p.d + 2 * (size_t)ctx->num_sms * 16 + GSV_VOCAB_MAX + 1;
}

bool gsv_gpt_ll2_supported(const gsv_gpt_ctx* ctx, int live_slots, int n_steps) {
  const int nd = ctx->p.d / 256;
  const bool shape = (nd == 1 || nd == 2) && ctx->p.F == 4 * ctx->p.d && ctx->p.F % NT == 0;
  // one attention CTA per (live sequence, head); every CTA needs at least one row of the narrowest phase and at
  // most one row per warp of the widest (no trips: each warp keeps ONE row per phase in registers);
  // phases per launch < 2^16 so that tags are unique
  return shape && live_slots >= 1 && live_slots <= MAXB && ctx->p.H * live_slots <= ctx->num_sms &&
         (ctx->p.F + ctx->num_sms - 1) / ctx->num_sms <= NWARP && ctx->p.F >= ctx->num_sms &&
         (ctx->p.V + 15) / 16 <= ctx->num_sms && ctx->p.d / 4 <= ctx->num_sms && ctx->p.d / 4 + ctx->p.d / 16 >= ctx->num_sms &&
         gsv_gpt_ll2_words_per_slot(ctx) * ctx->p.slots * sizeof(uint2) <= gsv_gpt_ll_buffer_bytes(ctx) && (long long)n_steps * (4 * ctx->p.L + 2) < 65000;
}

int gsv_gpt_decode_ll2_launch(gsv_gpt_ctx* ctx, int live_slots, int n_steps, cudaStream_t st) {
  if (ctx->dims.dtype == GSV_F16) return launch_ll2<__half>(ctx, live_slots, n_steps, st);
  return launch_ll2<__nv_bfloat16>(ctx, live_slots, n_steps, st);
}
