// gpt_decode_hx.cu -- single-sequence decode kernel, "head clusters": H clusters of 4 CTAs, 2 grid-wide exchanges per layer.
//
// Same arithmetic as the other decode kernels (reference t2s_model.py:67-105, 129-143, 442-456).  What the grid-wide
// flag-in-data kernel (gpt_decode_ll.cu, 299 us/token, 8 % of the HBM roofline) showed: a token is a chain of ~120
// dependent phases and every phase ends in an exchange through L2 between 148 polling CTAs (1.1-1.9 us each, 4-5 per
// layer) -- DRAM is 7 % busy.  This kernel cuts the chain the tensor-parallel way:
//   * cluster h (4 CTAs on one GPC) owns attention head h: its q/k/v rows, its K/V stream, its attention, and the
//     out-projection as a ROW-PARALLEL partial  Wo[:, 32h:32h+32] . att_h  (D outputs);
//   * CTA j = 4h + r owns 32 of the F = 128 H hidden units of the MLP: h_j = relu(W1[32j:32j+32] . x1 + b1) and the
//     down-projection again as a partial  W2[:, 32j:32j+32] . h_j, reduce-scattered over the cluster's 4 CTAs through
//     distributed shared memory (st.async pushes completing on the receiver's mbarrier, ~0.15 us per hop);
//   * the H cluster partials meet in L2 as {value, tag} words ("LL" protocol, gpt_sample.cuh): CTA (h, r) publishes rows
//     [R r, R r + R) of its cluster's partial (R = D/4) and READS those rows of all H clusters -- one L2 hop is an
//     all-reduce; the four quarter sums (+ residual + bias) are all-gathered inside the cluster.
// Per layer: 2 exchanges through L2 (after the out-projection, after the MLP) + 5 hops inside the cluster, against 5
// L2 exchanges before.  Summation orders are fixed (heads 0..H-1, ranks 0..3): results are bit-deterministic.
//
// Weights are re-tiled ONCE (first launch) into per-(layer, CTA) blobs of four sections -- everything a CTA touches in
// a layer, biases and LayerNorm parameters included, is contiguous: [q/k/v rows | Wo slice + bo + LN1 | W1 rows + b1 |
// W2 slice + b2 + LN2].  A section is ONE cp.async.bulk copy (8-34 KB: an SM ingests >= 8 KB bulk copies at ~200 GB/s,
// 1 KB ones at 20 GB/s, tools/ubench/bulk_bw.cu) into its own shared-memory area, re-requested for the next layer as
// soon as the phase that reads it is over: a copy has a whole layer time to land and never sits on the chain.  The head
// rows of a CTA (V / 4H, 17 KB) stay resident in shared memory for the whole launch.  K/V rows of the CTA's cached
// positions (p = r mod 4) are requested into registers before the q/k/v phase and prefetched into L2 a layer ahead.
// 64 of the 148 SMs are used (32 for an 8-head model): the rest stay free for the vocoder stream (TTS.infer_features_stream).
#include <cstring>
#include <type_traits>

#include "gpt_cluster_common.cuh"

namespace {

constexpr int CS = 4;                    // CTAs per cluster (= per head)
constexpr int ATT_W = 36;                // floats per attention partial: m, l, pad, pad, o[32]

// ---- blob layout (element offsets, T elements) -------------------------------------------------------------------
template <int D> struct HxLayout {
  static constexpr int R = D / CS;                       // rows of a D-vector one rank sums / owns
  static constexpr int S0 = 0;                           // 24 q/k/v rows [24][D], bias[32]
  static constexpr int S0_N = 24 * D + 32;
  static constexpr int S1 = S0 + S0_N;                   // Wo slice [R][32], bo[R], ln1 g[D], ln1 b[D]
  static constexpr int S1_N = R * 32 + R + 2 * D;
  static constexpr int S2 = S1 + S1_N;                   // W1 rows [32][D], b1[32]
  static constexpr int S2_N = 32 * D + 32;
  static constexpr int S3 = S2 + S2_N;                   // W2 slice [D][32], b2[R], ln2 g[D], ln2 b[D]
  static constexpr int S3_N = D * 32 + R + 2 * D;
  static constexpr int BLOB = S3 + S3_N;
};

// one thread per element of the packed copy; blob index = (layer * H + h) * 4 + r
template <typename T, int D>
__global__ void hx_pack_kernel(const GptParams p, T* __restrict__ out) {
  using Lo = HxLayout<D>;
  constexpr int R = Lo::R, F = 4 * D;
  const int H = p.H;
  const size_t total = (size_t)p.L * H * CS * Lo::BLOB;
  const T* Wqkv = reinterpret_cast<const T*>(p.w_qkv);
  const T* Wo = reinterpret_cast<const T*>(p.w_o);
  const T* W1 = reinterpret_cast<const T*>(p.w_1);
  const T* W2 = reinterpret_cast<const T*>(p.w_2);
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx % Lo::BLOB);
    size_t b = idx / Lo::BLOB;
    const int r = (int)(b % CS); b /= CS;
    const int h = (int)(b % H);
    const int l = (int)(b / H);
    const int j = h * CS + r;
    T v = Elem<T>::from_f(0.f);
    if (e < Lo::S1) {
      const int k = e - Lo::S0;
      if (k < 24 * D) {
        const int i = k / D, c = k - i * D, g = 24 * r + i;            // g-th of the head's 96 rows: q 0..31, k 32..63, v 64..95
        v = Wqkv[((size_t)l * 3 * D + (size_t)(g >> 5) * D + h * 32 + (g & 31)) * D + c];
      } else if (k - 24 * D < 24) {
        const int g = 24 * r + (k - 24 * D);
        v = reinterpret_cast<const T*>(p.b_qkv)[(size_t)l * 3 * D + (size_t)(g >> 5) * D + h * 32 + (g & 31)];
      }
    } else if (e < Lo::S2) {
      const int k = e - Lo::S1;
      if (k < R * 32) {
        const int row = R * r + (k >> 5), c = k & 31;
        v = Wo[((size_t)l * D + row) * D + h * 32 + c];
      } else if (k < R * 32 + R) v = reinterpret_cast<const T*>(p.b_o)[(size_t)l * D + R * r + (k - R * 32)];
      else if (k < R * 32 + R + D) v = reinterpret_cast<const T*>(p.ln1_g)[(size_t)l * D + (k - R * 32 - R)];
      else v = reinterpret_cast<const T*>(p.ln1_b)[(size_t)l * D + (k - R * 32 - R - D)];
    } else if (e < Lo::S3) {
      const int k = e - Lo::S2;
      if (k < 32 * D) {
        const int i = k / D, c = k - i * D;
        v = W1[((size_t)l * F + 32 * j + i) * D + c];
      } else v = reinterpret_cast<const T*>(p.b_1)[(size_t)l * F + 32 * j + (k - 32 * D)];
    } else {
      const int k = e - Lo::S3;
      if (k < D * 32) {
        const int row = k >> 5, c = k & 31;
        v = W2[((size_t)l * D + row) * F + 32 * j + c];
      } else if (k < D * 32 + R) v = reinterpret_cast<const T*>(p.b_2)[(size_t)l * D + R * r + (k - D * 32)];
      else if (k < D * 32 + R + D) v = reinterpret_cast<const T*>(p.ln2_g)[(size_t)l * D + (k - D * 32 - R)];
      else v = reinterpret_cast<const T*>(p.ln2_b)[(size_t)l * D + (k - D * 32 - R - D)];
    }
    out[idx] = v;
  }
}
// head rows of CTA j: global rows j, j + NC, j + 2 NC, ... (zero rows past V), [NC][HR][D]
template <typename T>
__global__ void hx_pack_head_kernel(const T* __restrict__ Wh, T* __restrict__ out, int NC, int HR, int V, int D) {
  const size_t total = (size_t)NC * HR * D;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % D);
    const int i = (int)((idx / D) % HR);
    const int j = (int)(idx / ((size_t)D * HR));
    const int g = j + NC * i;
    out[idx] = g < V ? Wh[(size_t)g * D + c] : Elem<T>::from_f(0.f);
  }
}

__device__ __forceinline__ void st_async_v4f(float* local_ptr, uint64_t* local_bar, unsigned rank, float4 v) {
  unsigned ra, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(local_bar)), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(ra), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)), "r"(rb)
               : "memory");
}

__device__ __forceinline__ void st_async_v2f(float* local_ptr, uint64_t* local_bar, unsigned rank, float a, float b) {
  unsigned ra, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(local_bar)), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
               ::"r"(ra), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(rb) : "memory");
}
__device__ __forceinline__ unsigned short ld_cg_u16(const void* p) { return __ldcg(reinterpret_cast<const unsigned short*>(p)); }
template <typename T> __device__ __forceinline__ float u16_to_f(unsigned short v);
template <> __device__ __forceinline__ float u16_to_f<__half>(unsigned short v) { return __half2float(__ushort_as_half(v)); }
template <> __device__ __forceinline__ float u16_to_f<__nv_bfloat16>(unsigned short v) { return __uint_as_float((unsigned)v << 16); }

constexpr int TILE = NWARP * 8;            // cached positions of one CTA scored per attention tile (128)

struct HxShared {
  __align__(16) float qkv_in[3 * GSV_HEAD_DIM];   // inbox: q | k | v of this head (24 values from each rank)
  __align__(16) float att_in[CS][ATT_W];          // inbox: attention partial of every rank
  __align__(16) float stat1[CS][2], stat2[CS][2]; // inboxes: (sum, sum of squares) of every rank's rows of y1 / y2
  __align__(16) float qkv_stage[24];              // this CTA's 24 q/k/v values before they are pushed
  __align__(16) float att_out[ATT_W];             // this CTA's attention partial (pushed)
  __align__(16) float sc[TILE];                   // scores of one tile
  float q[GSV_HEAD_DIM], kn[GSV_HEAD_DIM], vn[GSV_HEAD_DIM];
  float wsum[NWARP][GSV_HEAD_DIM + 1];            // per-warp attention sums: o[32], l
  float att[GSV_HEAD_DIM];                        // attention output of the head (after the 4-way merge)
  float hloc[GSV_HEAD_DIM];                       // this CTA's 32 hidden units
  float ypart[NWARP][2];                          // per-warp (sum, sum of squares) of this rank's rows
  int alive;
  int slot, kv;
  uint64_t wbar[4];                               // weight sections landed
  uint64_t hbar;                                  // head rows landed
  uint64_t xbar[5];                               // inboxes: 0 qkv, 1 att, 2 y1 (+stat1), 3 reduce-scatter, 4 y2 (+stat2)
};

template <typename T, int NCH>
__global__ void __launch_bounds__(NT, 1) gpt_decode_hx_kernel(const GptParams p, const int n_steps, const unsigned tag_base,
                                                               uint2* const ll_buf, const T* __restrict__ pack,
                                                               const T* __restrict__ hpack, const int HR) {
  using Lo = HxLayout<NCH * 256>;
  constexpr int D = NCH * 256, R = Lo::R;
  constexpr int G = NT / R;                 // lanes that share one row in the all-read (4 for D = 512, 8 for D = 256)
  constexpr unsigned YBYTES = D * 4u + CS * 8u;   // one fill of a y inbox: D values + 4 x (sum, sum of squares)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ HxShared sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, L = p.L, V = p.V, S = p.S;
  const int HPG = H / G;                    // heads summed per thread in the all-read
  const unsigned rank = cluster_rank();
  const int h = blockIdx.x / CS;            // head = cluster index
  const int j = blockIdx.x;                 // = 4 h + rank
  const int NC = gridDim.x;

  // shared memory: weight sections (T) | head rows (T) | xbuf[D] x1buf[D] ybox1[D] ybox2[D] rs_in[4][R] m2stage[D] ystage[R]
  //                | sampler scratch (CTA 0)
  T* wsec = reinterpret_cast<T*>(smem_raw);
  T* headw = wsec + Lo::BLOB;
  float* xbuf = reinterpret_cast<float*>(headw + (size_t)HR * D);   // layer input x (split layout): residual of the attention half
  float* x1buf = xbuf + D;                  // LN1 output (split layout): residual of the MLP half
  float* ybox1 = x1buf + D;                 // inbox: y1 (split layout)
  float* ybox2 = ybox1 + D;                 // inbox: y2
  float* rs_in = ybox2 + D;                 // inbox: [4 source ranks][R] MLP-down partial rows of this rank
  float* m2stage = rs_in + CS * R;          // this CTA's D partial outputs of the MLP-down before they are pushed
  float* ystage = m2stage + D;              // this rank's R summed rows before they are pushed
  float* samp = ystage + R;                 // sampler scratch (GSV_SAMPLE_SMEM_FLOATS), CTA 0 only

  // LL exchange areas ({value, tag} words): P1[H][D] | P2[H][D] | logits[VOCAB_MAX] | xin[D] | status
  uint2* P1 = ll_buf;
  uint2* P2 = P1 + (size_t)H * D;
  uint2* LLlogit = P2 + (size_t)H * D;
  uint2* LLxin = LLlogit + GSV_VOCAB_MAX;
  uint2* LLstat = LLxin + D;

  // ---- which sequence: the first active slot ----
  if (tid < 32) {
    const int flag = tid < p.slots ? ld_cg(p.active + tid) : 0;
    const unsigned m = __ballot_sync(0xffffffffu, flag != 0);
    if (tid == 0) {
      sh.slot = m ? (__ffs(m) - 1) : -1;
      sh.kv = m ? ld_cg(p.kv_len + (__ffs(m) - 1)) : 0;
    }
  }
  __syncthreads();
  const int slot = sh.slot;
  if (slot < 0) return;                     // uniform over the grid
  int kv = sh.kv;

  const T* const blob0 = pack + (size_t)j * Lo::BLOB;                       // layer 0 blob of this CTA
  const size_t blob_lstride = (size_t)H * CS * Lo::BLOB;
  // request section s of layer `layer` (thread 0, after the phase that read the previous contents is over)
  auto issue_sec = [&](int s, int layer) {
    const unsigned off = s == 0 ? Lo::S0 : (s == 1 ? Lo::S1 : (s == 2 ? Lo::S2 : Lo::S3));
    const unsigned bytes = 2u * (s == 0 ? Lo::S0_N : (s == 1 ? Lo::S1_N : (s == 2 ? Lo::S2_N : Lo::S3_N)));
    mbar_expect_tx(&sh.wbar[s], bytes);
    bulk_g2s(wsec + off, blob0 + (size_t)layer * blob_lstride + off, bytes, &sh.wbar[s]);
  };
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&sh.wbar[i], 1);
    mbar_init(&sh.hbar, 1);
    for (int i = 0; i < 5; ++i) mbar_init(&sh.xbar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < 4; ++s) issue_sec(s, 0);
    const unsigned hbytes = (unsigned)HR * D * 2u;
    mbar_expect_tx(&sh.hbar, hbytes);
    bulk_g2s(headw, hpack + (size_t)j * HR * D, hbytes, &sh.hbar);
    mbar_expect_tx(&sh.xbar[0], CS * 24 * 4u);
    mbar_expect_tx(&sh.xbar[1], CS * ATT_W * 4u);
    mbar_expect_tx(&sh.xbar[2], YBYTES);
    mbar_expect_tx(&sh.xbar[3], D * 4u);
    mbar_expect_tx(&sh.xbar[4], YBYTES);
  }
  unsigned wpar = 0;                        // parity of the weight sections' current fill (all four advance together, once per layer)
  unsigned xpar = 0;                        // parity of the inboxes (each used once per layer)
  // layer-0 input of the first step: xin left by prefill / the previous launch (plain fp32)
  for (int k = tid; k < D; k += NT) xbuf[split_pos(k, D)] = ld_cg(p.xin + (size_t)slot * D + k);
  __syncthreads();
  cluster_sync_all();                       // every CTA of the cluster has initialised its barriers

  unsigned tag = tag_base;
  const int sub = lane & 3, pg = lane >> 2;
  int gl = 0;                               // layers done in this launch (prefetch cursor)
  const int total_layers = n_steps * L;
  float xv[NCH * 8];                        // the layer input in dot-product order (every warp holds all of it)

  // LayerNorm of a gathered vector: every warp normalises all of it into its registers (statistics came with the
  // data: no reduction, no barrier); warp 0 keeps a copy in shared memory as the residual of the next half
  auto gathered_ln = [&](const float* ybox, const float (*stat)[2], const T* g, const T* b, float* dst) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int rr = 0; rr < CS; ++rr) { s += stat[rr][0]; q += stat[rr][1]; }
    const float mean = s * (1.f / (float)D);
    const float rstd = rsqrtf(fmaxf(q * (1.f / (float)D) - mean * mean, 0.f) + 1e-5f);
    load_x<NCH>(ybox, lane, xv);
    uint4 gv[NCH], bv[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      gv[c] = reinterpret_cast<const uint4*>(g)[c * 32 + lane];
      bv[c] = reinterpret_cast<const uint4*>(b)[c * 32 + lane];
    }
    ln_apply<T, NCH>(xv, mean, rstd, gv, bv);
    if (warp == 0) store_x<NCH>(dst, lane, xv);
  };
  // all-reduce over the clusters: this rank's R rows of P[0..H) summed in head order, + residual + bias, pushed to the 4
  // ranks together with (sum, sum of squares) of those rows.  G adjacent lanes share a row (HPG heads each).
  auto all_reduce = [&](const uint2* P, unsigned t, const float* resid, const T* bias, float* ybox, float (*stat)[2], uint64_t* bar) {
    const int i = tid / G, g = tid % G;
    const uint2* src = P + (size_t)(g * HPG) * D + R * rank + i;
    float acc = 0.f;
    uint2 w[4];
    for (int h0 = 0; h0 < HPG; h0 += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) w[u] = make_uint2(0u, ~t);
      bool ok;
      do {
        ok = true;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (h0 + u < HPG && w[u].y != t) { w[u] = ll_peek(src + (size_t)(h0 + u) * D); ok = ok && (w[u].y == t); }
        }
      } while (!ok);
#pragma unroll
      for (int u = 0; u < 4; ++u) if (h0 + u < HPG) acc += __uint_as_float(w[u].x);
    }
#pragma unroll
    for (int o = 1; o < G; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    float y = 0.f;
    if (g == 0) {
      y = acc + resid[split_pos(R * rank + i, D)] + Elem<T>::to_f(bias[i]);
      ystage[i] = y;
    }
    float s = y, q = y * y;                  // lanes with g != 0 contribute zeros
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (lane == 0) { sh.ypart[warp][0] = s; sh.ypart[warp][1] = q; }
    __syncthreads();
    if (tid < R) {                          // R/4 float4 per target x 4 targets
      const int tgt = tid / (R / 4), q4 = tid % (R / 4);
      st_async_v4f(ybox + split_pos(R * rank + q4 * 4, D), bar, (unsigned)tgt, *reinterpret_cast<const float4*>(ystage + q4 * 4));
    } else if (tid < R + CS) {
      float ss = 0.f, qq = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < NWARP; ++w2) { ss += sh.ypart[w2][0]; qq += sh.ypart[w2][1]; }
      st_async_v2f(&stat[rank][0], bar, (unsigned)(tid - R), ss, qq);
    }
  };

#pragma unroll 1
  for (int step = 0; step < n_steps; ++step) {
    load_x<NCH>(xbuf, lane, xv);
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      mark(p, 1);
      const size_t head_base = ((size_t)(l * p.slots + slot) * H + h) * (size_t)S * GSV_HEAD_DIM;
      const T* kc = reinterpret_cast<const T*>(p.kc) + head_base;
      const T* vc = reinterpret_cast<const T*>(p.vc) + head_base;
      // cached positions of this CTA: p = 4 m + rank, m < n_mine; the newest position (p = kv) belongs to rank kv mod 4
      const int n_mine = kv > (int)rank ? (kv - (int)rank + 3) >> 2 : 0;
      const bool owner = (kv & 3) == (int)rank;
      const int n_tot = n_mine + (owner ? 1 : 0);
      // ---- K rows (4 lanes per position) and V columns (warp = position mod 16, lane = dimension) of the first two
      //      tiles requested before anything is waited for
      uint4 kr0 = make_uint4(0, 0, 0, 0), kr1 = kr0;
      unsigned short vh0[8], vh1[8];
      {
        const int m0 = warp * 8 + pg, m1 = m0 + TILE;
        if (m0 < n_mine) kr0 = ld_cg16(kc + (size_t)(4 * m0 + (int)rank) * GSV_HEAD_DIM + sub * 8);
        if (m1 < n_mine) kr1 = ld_cg16(kc + (size_t)(4 * m1 + (int)rank) * GSV_HEAD_DIM + sub * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int a = warp + NWARP * i, b = a + TILE;
          vh0[i] = a < n_mine ? ld_cg_u16(vc + (size_t)(4 * a + (int)rank) * GSV_HEAD_DIM + lane) : (unsigned short)0;
          vh1[i] = b < n_mine ? ld_cg_u16(vc + (size_t)(4 * b + (int)rank) * GSV_HEAD_DIM + lane) : (unsigned short)0;
        }
      }
      if (tid == 32 && rank == 0 && l + 1 < L && kv > 0) {    // next layer's K/V of this head into L2
        const size_t nxt = (size_t)p.slots * H * S * GSV_HEAD_DIM;
        const unsigned bytes = (unsigned)kv * GSV_HEAD_DIM * (unsigned)sizeof(T);
        l2_prefetch(kc + nxt, bytes);
        l2_prefetch(vc + nxt, bytes);
      }
      // ================= q/k/v rows of this rank (24 of the head's 96), all-gathered in the cluster =================
      mbar_wait(&sh.wbar[0], wpar);
      mark(p, 30);
      {
        const T* w0 = wsec + Lo::S0;
        const int ib = warp + NWARP < 24 ? warp + NWARP : warp;       // warps 8..15 have one row: the second is a dummy
        uint4 wa[NCH], wb[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          wa[c] = reinterpret_cast<const uint4*>(w0 + (size_t)warp * D)[c * 32 + lane];
          wb[c] = reinterpret_cast<const uint4*>(w0 + (size_t)ib * D)[c * 32 + lane];
        }
        const float a = reduce2(dot_regs<T, NCH>(wa, xv), dot_regs<T, NCH>(wb, xv), lane);   // lanes 0..15: row a, 16..31: row b
        if (lane == 0) sh.qkv_stage[warp] = a + Elem<T>::to_f(w0[24 * D + warp]);
        if (lane == 16 && warp + NWARP < 24) sh.qkv_stage[warp + NWARP] = a + Elem<T>::to_f(w0[24 * D + warp + NWARP]);
      }
      __syncthreads();
      if (tid == 0 && gl + 1 < total_layers) issue_sec(0, (l + 1) % L);
      if (tid < 24) {                        // 6 float4 per target x 4 targets; value g = 24 rank + i lands at qkv_in[g]
        const int tgt = tid / 6, q4 = tid % 6;
        st_async_v4f(sh.qkv_in + 24 * rank + q4 * 4, &sh.xbar[0], (unsigned)tgt, *reinterpret_cast<const float4*>(sh.qkv_stage + q4 * 4));
      }
      mark(p, 2);
      mbar_wait(&sh.xbar[0], xpar);
      if (tid < 3 * GSV_HEAD_DIM) {
        const int which = tid >> 5, c = tid & 31;
        const float v = sh.qkv_in[tid];
        if (which == 0) sh.q[c] = v * (rsqrtf((float)GSV_HEAD_DIM) * 1.4426950408889634f);
        else {
          const T t16 = Elem<T>::from_f(v);             // the reference attends over the 16-bit cache entry it has just written
          (which == 1 ? sh.kn : sh.vn)[c] = Elem<T>::to_f(t16);
          if (owner) {
            T* cache = reinterpret_cast<T*>(which == 1 ? p.kc : p.vc);
            cache[head_base + (size_t)kv * GSV_HEAD_DIM + c] = t16;
          }
        }
      }
      __syncthreads();                        // q / k / v staged; qkv_in fully read
      if (tid == 0) mbar_expect_tx(&sh.xbar[0], CS * 24 * 4u);
      mark(p, 3);
      // ================= attention over this CTA's positions, tile by tile; one maximum per tile for the whole CTA =================
      {
        float q[8];
#pragma unroll
        for (int jq = 0; jq < 8; ++jq) q[jq] = sh.q[sub * 8 + jq];
        float Mrun = GSV_NEG_INF, lacc = 0.f, oacc = 0.f;       // oacc: output dimension `lane` over this warp's positions
        const int n_tiles = (n_tot + TILE - 1) / TILE;
#pragma unroll 1
        for (int tile = 0; tile < n_tiles; ++tile) {
          // scores: 4 lanes per position
          {
            const int m = tile * TILE + warp * 8 + pg;
            float kf[8];
            if (m < n_mine) {
              uint4 kr = tile == 0 ? kr0 : kr1;
              if (tile > 1) kr = ld_cg16(kc + (size_t)(4 * m + (int)rank) * GSV_HEAD_DIM + sub * 8);
              unpack8<T>(kr, kf);
            } else {
#pragma unroll
              for (int jq = 0; jq < 8; ++jq) kf[jq] = sh.kn[sub * 8 + jq];     // the newest position (or an empty slot)
            }
            float sc_ = 0.f;
#pragma unroll
            for (int jq = 0; jq < 8; ++jq) sc_ = fmaf(q[jq], kf[jq], sc_);
            sc_ += __shfl_xor_sync(0xffffffffu, sc_, 1);
            sc_ += __shfl_xor_sync(0xffffffffu, sc_, 2);
            if (sub == 0) sh.sc[warp * 8 + pg] = m < n_tot ? sc_ : GSV_NEG_INF;
          }
          __syncthreads();
          // tile maximum (every warp computes the same value)
          float tm;
          {
            const float4 s4 = *reinterpret_cast<const float4*>(&sh.sc[lane * 4]);
            tm = fmaxf(fmaxf(s4.x, s4.y), fmaxf(s4.z, s4.w));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, o));
          }
          const float Mnew = fmaxf(Mrun, tm);
          const float scale = Mrun > GSV_NEG_INF ? exp2f(Mrun - Mnew) : 0.f;
          lacc *= scale;
          oacc *= scale;
          // probabilities x V: warp w owns the tile's positions w, w + 16, ...; lane = output dimension
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int a = tile * TILE + warp + NWARP * i;
            if (a < n_tot) {
              const float pr = exp2f(sh.sc[warp + NWARP * i] - Mnew);
              float vv;
              if (a < n_mine) {
                unsigned short raw = tile == 0 ? vh0[i] : vh1[i];
                if (tile > 1) raw = ld_cg_u16(vc + (size_t)(4 * a + (int)rank) * GSV_HEAD_DIM + lane);
                vv = u16_to_f<T>(raw);
              } else vv = sh.vn[lane];
              lacc += pr;
              oacc = fmaf(pr, vv, oacc);
            }
          }
          Mrun = Mnew;
          if (tile + 1 < n_tiles) __syncthreads();          // sc is rewritten by the next tile
        }
        sh.wsum[warp][lane] = oacc;
        if (lane == 0) sh.wsum[warp][GSV_HEAD_DIM] = lacc;
        __syncthreads();
        if (warp == 0) {
          float oa = 0.f, Ls = 0.f;
#pragma unroll
          for (int w2 = 0; w2 < NWARP; ++w2) { oa += sh.wsum[w2][lane]; Ls += sh.wsum[w2][GSV_HEAD_DIM]; }
          if (lane == 0) { sh.att_out[0] = Mrun; sh.att_out[1] = Ls; sh.att_out[2] = 0.f; sh.att_out[3] = 0.f; }
          sh.att_out[4 + lane] = oa;
          __syncwarp();
          if (lane < CS * (ATT_W / 4)) {     // 9 float4 per target x 4 targets: two rounds of the 32 lanes
            const int tgt = lane / (ATT_W / 4), q4 = lane % (ATT_W / 4);
            st_async_v4f(&sh.att_in[rank][q4 * 4], &sh.xbar[1], (unsigned)tgt, *reinterpret_cast<const float4*>(sh.att_out + q4 * 4));
          }
          if (lane + 32 < CS * (ATT_W / 4)) {
            const int e = lane + 32, tgt = e / (ATT_W / 4), q4 = e % (ATT_W / 4);
            st_async_v4f(&sh.att_in[rank][q4 * 4], &sh.xbar[1], (unsigned)tgt, *reinterpret_cast<const float4*>(sh.att_out + q4 * 4));
          }
        }
      }
      mark(p, 4);
      mbar_wait(&sh.xbar[1], xpar);
      if (warp == 0) {
        float M = GSV_NEG_INF;
#pragma unroll
        for (int rr = 0; rr < CS; ++rr) M = fmaxf(M, sh.att_in[rr][0]);
        float Ls = 0.f, oa = 0.f;
#pragma unroll
        for (int rr = 0; rr < CS; ++rr) {
          const float mr = sh.att_in[rr][0];
          const float sc = mr > GSV_NEG_INF ? exp2f(mr - M) : 0.f;
          Ls = fmaf(sh.att_in[rr][1], sc, Ls);
          oa = fmaf(sh.att_in[rr][4 + lane], sc, oa);
        }
        sh.att[lane] = oa / Ls;
      }
      __syncthreads();                        // att staged; att_in fully read
      if (tid == 0) mbar_expect_tx(&sh.xbar[1], CS * ATT_W * 4u);
      mark(p, 5);
      // ================= out-projection partial of this head: rows [R rank, R rank + R) of Wo[:, 32h:32h+32] . att =================
      mbar_wait(&sh.wbar[1], wpar);
      tag += 1;
      {
        const T* w1s = wsec + Lo::S1;
        float af[8];
#pragma unroll
        for (int jq = 0; jq < 8; ++jq) af[jq] = sh.att[sub * 8 + jq];
        constexpr int RPW = R / NWARP;        // 8 (D = 512) or 4 (D = 256) rows per warp
        const bool valid = pg < RPW;
        const int row = warp * RPW + (valid ? pg : 0);
        float wf[8];
        unpack8<T>(reinterpret_cast<const uint4*>(w1s + (size_t)row * 32)[sub], wf);
        float a = 0.f;
#pragma unroll
        for (int jq = 0; jq < 8; ++jq) a = fmaf(wf[jq], af[jq], a);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        if (valid && sub == 0) ll_store(P1 + (size_t)h * D + R * rank + row, a, tag);
      }
      mark(p, 6);
      // ================= all-reduce 1: y1 = x + bo + sum_h partial_h ; x1 = LN1(y1) =================
      all_reduce(P1, tag, xbuf, wsec + Lo::S1 + R * 32, ybox1, sh.stat1, &sh.xbar[2]);
      mark(p, 7);
      mbar_wait(&sh.xbar[2], xpar);
      gathered_ln(ybox1, sh.stat1, wsec + Lo::S1 + R * 32 + R, wsec + Lo::S1 + R * 32 + R + D, x1buf);
      mark(p, 8);
      // ================= MLP-up: this CTA's 32 hidden units (rows warp and warp + 16) =================
      mbar_wait(&sh.wbar[2], wpar);
      mark(p, 31);
      {
        const T* w2s = wsec + Lo::S2;
        uint4 wa[NCH], wb[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          wa[c] = reinterpret_cast<const uint4*>(w2s + (size_t)warp * D)[c * 32 + lane];
          wb[c] = reinterpret_cast<const uint4*>(w2s + (size_t)(warp + NWARP) * D)[c * 32 + lane];
        }
        const float a = reduce2(dot_regs<T, NCH>(wa, xv), dot_regs<T, NCH>(wb, xv), lane);
        if ((lane & 15) == 0) {
          const int i = warp + (lane >> 4) * NWARP;
          sh.hloc[i] = fmaxf(a + Elem<T>::to_f(w2s[32 * D + i]), 0.f);
        }
      }
      __syncthreads();                        // hloc complete; ybox1 / stat1 / S1 / S2 fully read by every warp
      if (tid == 0) {
        mbar_expect_tx(&sh.xbar[2], YBYTES);
        if (gl + 1 < total_layers) { issue_sec(1, (l + 1) % L); issue_sec(2, (l + 1) % L); }
      }
      mark(p, 9);
      // ================= MLP-down partial over these 32 hidden units, reduce-scattered over the cluster =================
      mbar_wait(&sh.wbar[3], wpar);
      {
        const T* w3s = wsec + Lo::S3;
        float hf[8];
#pragma unroll
        for (int jq = 0; jq < 8; ++jq) hf[jq] = sh.hloc[sub * 8 + jq];
        constexpr int RPW = D / NWARP;        // 32 / 16 rows per warp
#pragma unroll
        for (int rr = 0; rr < RPW; rr += 8) {
          const int row = warp * RPW + rr + pg;
          float wf[8];
          unpack8<T>(reinterpret_cast<const uint4*>(w3s + (size_t)row * 32)[sub], wf);
          float a = 0.f;
#pragma unroll
          for (int jq = 0; jq < 8; ++jq) a = fmaf(wf[jq], hf[jq], a);
          a += __shfl_xor_sync(0xffffffffu, a, 1);
          a += __shfl_xor_sync(0xffffffffu, a, 2);
          if (sub == 0) m2stage[row] = a;
        }
      }
      __syncthreads();
      if (tid < D / 4) {                      // rows 4 tid .. 4 tid + 3 go to rank (4 tid) / R, slot [this rank][(4 tid) % R]
        const int row = 4 * tid, tgt = row / R;
        st_async_v4f(rs_in + rank * R + (row - tgt * R), &sh.xbar[3], (unsigned)tgt, *reinterpret_cast<const float4*>(m2stage + row));
      }
      mark(p, 10);
      mbar_wait(&sh.xbar[3], xpar);
      tag += 1;
      if (tid < R) {
        const float s4 = ((rs_in[tid] + rs_in[R + tid]) + rs_in[2 * R + tid]) + rs_in[3 * R + tid];
        ll_store(P2 + (size_t)h * D + R * rank + tid, s4, tag);
      }
      mark(p, 11);
      // ================= all-reduce 2: y2 = x1 + b2 + sum partials ; x = LN2(y2) =================
      // (the barrier inside follows the reads of rs_in: its inbox is re-armed right after)
      all_reduce(P2, tag, x1buf, wsec + Lo::S3 + D * 32, ybox2, sh.stat2, &sh.xbar[4]);
      if (tid == 0) mbar_expect_tx(&sh.xbar[3], D * 4u);
      mark(p, 12);
      mbar_wait(&sh.xbar[4], xpar);
      gathered_ln(ybox2, sh.stat2, wsec + Lo::S3 + D * 32 + R, wsec + Lo::S3 + D * 32 + R + D, xbuf);
      // ybox2 / stat2 / S3 are re-armed / re-requested after the next barrier every warp passes (q/k/v phase of the next
      // layer, or the head below): see `late_rearm`
      wpar ^= 1u;
      xpar ^= 1u;
      gl += 1;
      if (l + 1 < L) {
        // the next layer's first barrier is after its q/k/v rows; the re-arm must follow every warp's reads of ybox2 and
        // S3 (LayerNorm parameters), so it is done here behind a barrier of its own
        __syncthreads();
        if (tid == 0) {
          mbar_expect_tx(&sh.xbar[4], YBYTES);
          if (gl < total_layers) issue_sec(3, l + 1);
        }
      }
    }
    mark(p, 13);
    // ================= head: this CTA's vocabulary rows j, j + NC, ... =================
    if (step == 0) mbar_wait(&sh.hbar, 0u);
    tag += 1;
#pragma unroll 1
    for (int i = warp; i < HR; i += NWARP) {
      const int g = j + NC * i;
      if (g < V) {
        uint4 w[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) w[c] = reinterpret_cast<const uint4*>(headw + (size_t)i * D)[c * 32 + lane];
        const float a = warp_allsum(dot_regs<T, NCH>(w, xv));
        if (lane == 0) ll_store(LLlogit + g, a, tag);
      }
    }
    __syncthreads();                          // every warp is past the last layer's LayerNorm: ybox2 / S3 may be refilled
    if (tid == 0) {
      mbar_expect_tx(&sh.xbar[4], YBYTES);
      if (gl < total_layers) issue_sec(3, 0);
    }
    mark(p, 20);
    // ================= sampling in CTA 0; next input and status published for everyone =================
    const unsigned tag_logits = tag;
    tag += 1;
    if (j == 0) {
      for (int v = tid; v < V; v += NT) samp[v] = ll_wait(LLlogit + v, tag_logits);
      __syncthreads();
      SampleLL io;
      io.preloaded = true;
      io.xin_ll = LLxin;
      io.status_ll = LLstat;
      io.tag = tag;
      io.kv_len = kv + 1;
      io.xin_smem = nullptr;
      io.alive_smem = nullptr;
      sample_slot<T>(p, slot, samp, &io);
    }
    if (tid == 0) sh.alive = ll_wait(LLstat, tag) != 0.f ? 1 : 0;
    __syncthreads();
    const bool alive = sh.alive != 0;
    kv += 1;
    mark(p, 21);
    if (!alive) break;
    if (step + 1 < n_steps) {
      if (tid < D) xbuf[split_pos(tid, D)] = ll_wait(LLxin + tid, tag);
      __syncthreads();
    }
  }
  // outstanding weight requests (one per section at most), then leave together: no CTA exits while a peer may still push
  if (gl < total_layers) {
    for (int s = 0; s < 4; ++s) mbar_wait(&sh.wbar[s], wpar);
  }
  cluster_sync_all();
}

template <typename T, int NCH>
size_t hx_smem_bytes(int HR) {
  using Lo = HxLayout<NCH * 256>;
  constexpr int D = NCH * 256, R = Lo::R;
  return (size_t)Lo::BLOB * 2 + (size_t)HR * D * 2 + sizeof(float) * ((size_t)4 * D + CS * R + D + R + ((GSV_SAMPLE_SMEM_FLOATS + 3) & ~3));
}

template <typename T, int NCH>
int launch_hx_t(gsv_gpt_ctx* ctx, int n_steps, cudaStream_t st) {
  using Lo = HxLayout<NCH * 256>;
  constexpr int D = NCH * 256;
  const int H = ctx->p.H, L = ctx->p.L, V = ctx->p.V, NC = H * CS;
  const int HR = (V + NC - 1) / NC;
  void* fn = (void*)gpt_decode_hx_kernel<T, NCH>;
  const size_t bytes = hx_smem_bytes<T, NCH>(HR);
  if (!ctx->hx_pack) {
    void *pk = nullptr, *hp = nullptr;
    GSV_CUDA(cudaMalloc(&pk, (size_t)L * NC * Lo::BLOB * sizeof(T)));
    GSV_CUDA(cudaMalloc(&hp, (size_t)NC * HR * D * sizeof(T)));
    hx_pack_kernel<T, D><<<ctx->num_sms * 8, 256, 0, st>>>(ctx->p, reinterpret_cast<T*>(pk));
    hx_pack_head_kernel<T><<<ctx->num_sms, 256, 0, st>>>(reinterpret_cast<const T*>(ctx->p.w_head), reinterpret_cast<T*>(hp), NC, HR, V, D);
    GSV_CUDA(cudaGetLastError());
    ctx->hx_pack = pk;
    ctx->hx_head_pack = hp;
    ctx->launches += 2;
  }
  GSV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  GptParams p = ctx->p;
  int ns = n_steps;
  ctx->ll_seq += 1;
  if ((ctx->ll_seq & 0xffffull) == 0) {     // see gpt_decode_ll.cu: tags wrap every 65 536 launches
    GSV_CUDA(cudaMemsetAsync(ctx->ll_buf, 0, gsv_gpt_ll_buffer_bytes(ctx), st));
    ctx->ll_seq += 1;
  }
  unsigned tag_base = (unsigned)(ctx->ll_seq << 16);
  uint2* buf = reinterpret_cast<uint2*>(ctx->ll_buf);
  const T* pk = reinterpret_cast<const T*>(ctx->hx_pack);
  const T* hp = reinterpret_cast<const T*>(ctx->hx_head_pack);
  int hr = HR;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(NC); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = bytes; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (ctx->hx_clusters_ok == 0) {
    // every cluster must be co-resident (the clusters wait for each other): ask the occupancy calculator once
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, fn, &cfg);
    ctx->hx_clusters_ok = (e == cudaSuccess && n >= H) ? 1 : -1;
    if (e != cudaSuccess) cudaGetLastError();
  }
  if (ctx->hx_clusters_ok < 0) { gsv_set_error("hx decode kernel: %d clusters of %d CTAs are not co-resident on this device", H, CS); return GSV_ERR_STATE; }
  void* args[] = {&p, &ns, &tag_base, &buf, &pk, &hp, &hr};
  GSV_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  ctx->launches += 1;
  return GSV_OK;
}

template <typename T>
int launch_hx(gsv_gpt_ctx* ctx, int n_steps, cudaStream_t st) {
  if (ctx->p.d == 512) return launch_hx_t<T, 2>(ctx, n_steps, st);
  if (ctx->p.d == 256) return launch_hx_t<T, 1>(ctx, n_steps, st);
  return GSV_ERR_ARG;
}

}  // namespace

// words of the LL exchange areas this kernel needs: P1[H][D] + P2[H][D] + logits + xin + status
size_t gsv_gpt_hx_buffer_words(const gsv_gpt_ctx* ctx) {
  return (size_t)2 * ctx->p.H * ctx->p.d + GSV_VOCAB_MAX + ctx->p.d + 8;
}

bool gsv_gpt_hx_supported(const gsv_gpt_ctx* ctx, int live_slots, int n_steps) {
  const GptParams& p = ctx->p;
  const bool shape = (p.d == 256 || p.d == 512) && p.H * GSV_HEAD_DIM == p.d && p.F == 4 * p.d && p.V <= GSV_VOCAB_MAX;
  return shape && live_slots == 1 && p.H * CS <= ctx->num_sms && ctx->hx_clusters_ok >= 0 &&
         (long long)n_steps * (2 * p.L + 2) < 65000;
}

int gsv_gpt_decode_hx_launch(gsv_gpt_ctx* ctx, int n_steps, cudaStream_t st) {
  if (ctx->dims.dtype == GSV_F16) return launch_hx<__half>(ctx, n_steps, st);
  return launch_hx<__nv_bfloat16>(ctx, n_steps, st);
}
