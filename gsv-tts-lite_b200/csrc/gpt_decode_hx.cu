// gpt_decode_hx.cu -- single-sequence decode kernel, "head clusters": 2 grid-wide exchanges per layer.
//
// Same arithmetic as the other decode kernels (reference t2s_model.py:67-105, 129-143, 442-456).  What the grid-wide
// flag-in-data kernel (gpt_decode_ll.cu, 299 us/token, 8 % of the HBM roofline) showed: a token is a chain of ~120
// dependent phases and every phase ends in an exchange through L2 between 148 polling CTAs (1.1-1.9 us each, 4-5 per
// layer) -- DRAM is 7 % busy.  This kernel cuts the chain the tensor-parallel way, on 4 H CTAs in clusters of CS
// (CS = 4, 8 or 16 CTAs = 1, 2 or 4 attention heads per cluster):
//   * 4 CTAs own an attention head: its q/k/v rows (24 each, all-gathered between the four), its K/V stream (cached
//     positions dealt p mod 4), its attention (partials pushed to the whole cluster);
//   * the out-projection is ROW-PARALLEL over the cluster's heads: CTA r computes rows [Rc r, Rc r + Rc) (Rc = D / CS)
//     of  Wo[:, cluster's columns] . att_cluster  -- a partial of the D outputs;
//   * CTA j owns 32 of the F = 128 H hidden units of the MLP: h_j = relu(W1[32j:32j+32] . x1 + b1) and the down-projection
//     again as a partial  W2[:, 32j:32j+32] . h_j, reduce-scattered over the cluster through distributed shared memory
//     (st.async pushes completing on the receiver's mbarrier, ~0.2 us per hop);
//   * the cluster partials meet in L2 as {value, tag} words ("LL" protocol, gpt_sample.cuh): CTA (c, r) publishes rows
//     [Rc r, Rc r + Rc) of its cluster's partial and READS those rows of every cluster -- one L2 hop is an all-reduce;
//     the sums (+ residual + bias) and their LayerNorm statistics are all-gathered inside the cluster.
// Per layer: 2 exchanges through L2 (after the out-projection, after the MLP) + 5 hops inside the cluster, against 5
// L2 exchanges before.  The larger the cluster, the fewer words cross L2: CS = 16 polls 4 x 32 words per CTA and
// exchange, CS = 4 polls 16 x 128.  Summation orders are fixed: results are bit-deterministic.
//
// Weights are re-tiled ONCE (first launch) into per-(layer, CTA) blobs of four sections -- everything a CTA touches in
// a layer, biases and LayerNorm parameters included, is contiguous: [q/k/v rows | Wo slice + bo + LN1 | W1 rows + b1 |
// W2 slice + b2 + LN2].  A section is ONE cp.async.bulk copy (8-34 KB: an SM ingests >= 8 KB bulk copies at ~200 GB/s,
// 1 KB ones at 20 GB/s, tools/ubench/bulk_bw.cu) into its own shared-memory area, re-requested for the next layer as
// soon as the phase that reads it is over: a copy has a whole layer time to land and never sits on the chain.  The head
// rows of a CTA (V / 4H, 17 KB) stay resident in shared memory for the whole launch.  The head's cached positions are
// dealt to its 4 CTAs in contiguous quarters, so a CTA's K rows (and V rows) are one run of the cache: one thread
// requests them a layer ahead with two bulk copies.
// What the in-kernel timelines showed next (tools/hx_timeline.py): with 16 warps per CTA every instruction a thread
// executes costs 4 issue cycles per scheduler, so per-phase instruction counts matter as much as latencies -- loads
// that are not needed are branched over, not predicated; LayerNorm is one element per thread with statistics that
// travel with the data.
// 64 of the 148 SMs are used (32 for an 8-head model): the rest stay free for the vocoder stream (TTS.infer_features_stream).
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "gpt_cluster_common.cuh"

namespace {

constexpr int ATT_W = 36;                // floats per attention partial: m, l, pad, pad, o[32]
constexpr int CS_MAX = 16;
constexpr int TILE = NWARP * 8;          // cached positions of one CTA scored per attention pass (128)
constexpr int QMAX = 2 * TILE;           // cached positions per CTA whose K/V rows are prefetched into shared memory (kv <= 1024)

// ---- blob layout (element offsets, T elements) -------------------------------------------------------------------
template <int D, int CS> struct HxLayout {
  static constexpr int HC = CS / 4;                      // heads per cluster
  static constexpr int KO = 32 * HC;                     // columns of Wo a cluster owns
  static constexpr int RC = D / CS;                      // rows of a D-vector one rank sums / owns
  static constexpr int S0 = 0;                           // 24 q/k/v rows [24][D], bias[32]
  static constexpr int S0_N = 24 * D + 32;
  static constexpr int S1 = S0 + S0_N;                   // Wo slice [RC][KO], bo[RC], ln1 g[D], ln1 b[D]
  static constexpr int S1_N = RC * KO + RC + 2 * D;
  static constexpr int S2 = S1 + S1_N;                   // W1 rows [32][D], b1[32]
  static constexpr int S2_N = 32 * D + 32;
  static constexpr int S3 = S2 + S2_N;                   // W2 slice [D][32], b2[RC], ln2 g[D], ln2 b[D]
  static constexpr int S3_N = D * 32 + RC + 2 * D;
  static constexpr int BLOB = S3 + S3_N;
  static_assert(S1 % 8 == 0 && S2 % 8 == 0 && S3 % 8 == 0 && BLOB % 8 == 0, "sections are 16-byte aligned");
};

// one thread per element of the packed copy; blob index = layer * NC + j, j = cluster * CS + rank
template <typename T, int D, int CS>
__global__ void hx_pack_kernel(const GptParams p, T* __restrict__ out) {
  using Lo = HxLayout<D, CS>;
  constexpr int RC = Lo::RC, KO = Lo::KO, F = 4 * D;
  const int NC = 4 * p.H;
  const size_t total = (size_t)p.L * NC * Lo::BLOB;
  const T* Wqkv = reinterpret_cast<const T*>(p.w_qkv);
  const T* Wo = reinterpret_cast<const T*>(p.w_o);
  const T* W1 = reinterpret_cast<const T*>(p.w_1);
  const T* W2 = reinterpret_cast<const T*>(p.w_2);
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx % Lo::BLOB);
    const size_t b = idx / Lo::BLOB;
    const int j = (int)(b % NC);
    const int l = (int)(b / NC);
    const int c = j / CS, r = j % CS;
    const int h = c * Lo::HC + r / 4, s4 = r % 4;          // the CTA's head and its quarter of the head's work
    T v = Elem<T>::from_f(0.f);
    if (e < Lo::S1) {
      const int k = e - Lo::S0;
      if (k < 24 * D) {
        const int i = k / D, col = k - i * D, g = 24 * s4 + i;            // g-th of the head's 96 rows: q 0..31, k 32..63, v 64..95
        v = Wqkv[((size_t)l * 3 * D + (size_t)(g >> 5) * D + h * 32 + (g & 31)) * D + col];
      } else if (k - 24 * D < 24) {
        const int g = 24 * s4 + (k - 24 * D);
        v = reinterpret_cast<const T*>(p.b_qkv)[(size_t)l * 3 * D + (size_t)(g >> 5) * D + h * 32 + (g & 31)];
      }
    } else if (e < Lo::S2) {
      const int k = e - Lo::S1;
      if (k < RC * KO) {
        const int row = RC * r + k / KO, col = k % KO;
        v = Wo[((size_t)l * D + row) * D + c * KO + col];
      } else if (k < RC * KO + RC) v = reinterpret_cast<const T*>(p.b_o)[(size_t)l * D + RC * r + (k - RC * KO)];
      else if (k < RC * KO + RC + D) v = reinterpret_cast<const T*>(p.ln1_g)[(size_t)l * D + (k - RC * KO - RC)];
      else v = reinterpret_cast<const T*>(p.ln1_b)[(size_t)l * D + (k - RC * KO - RC - D)];
    } else if (e < Lo::S3) {
      const int k = e - Lo::S2;
      if (k < 32 * D) {
        const int i = k / D, col = k - i * D;
        v = W1[((size_t)l * F + 32 * j + i) * D + col];
      } else v = reinterpret_cast<const T*>(p.b_1)[(size_t)l * F + 32 * j + (k - 32 * D)];
    } else {
      const int k = e - Lo::S3;
      if (k < D * 32) {
        const int row = k >> 5, col = k & 31;
        v = W2[((size_t)l * D + row) * F + 32 * j + col];
      } else if (k < D * 32 + RC) v = reinterpret_cast<const T*>(p.b_2)[(size_t)l * D + RC * r + (k - D * 32)];
      else if (k < D * 32 + RC + D) v = reinterpret_cast<const T*>(p.ln2_g)[(size_t)l * D + (k - D * 32 - RC)];
      else v = reinterpret_cast<const T*>(p.ln2_b)[(size_t)l * D + (k - D * 32 - RC - D)];
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

struct HxShared {
  __align__(16) float qkv_in[3 * GSV_HEAD_DIM];   // inbox: q | k | v of this CTA's head (24 values from each of its 4 CTAs)
  __align__(16) float att_in[CS_MAX][ATT_W];      // inbox: attention partial of every rank of the cluster
  __align__(16) float stat1[CS_MAX][2], stat2[CS_MAX][2];   // inboxes: (sum, sum of squares) of every rank's rows of y1 / y2
  __align__(16) float qkv_stage[24];              // this CTA's 24 q/k/v values before they are pushed
  __align__(16) float att_out[ATT_W];             // this CTA's attention partial (pushed)
  float q[GSV_HEAD_DIM], kn[GSV_HEAD_DIM], vn[GSV_HEAD_DIM];
  float wpart[NWARP][GSV_HEAD_DIM + 2];           // per-warp attention partials: m, l, o[32]
  float att[CS_MAX / 4 * GSV_HEAD_DIM];           // attention output of the cluster's heads (after the 4-way merges)
  float hloc[GSV_HEAD_DIM];                       // this CTA's 32 hidden units
  float snew;                                     // score of the newest position (owner CTA)
  int alive;
  int slot, kv;
  uint64_t wbar[4];                               // weight sections landed
  uint64_t hbar;                                  // head rows landed
  uint64_t kbar;                                  // prefetched K/V rows landed
  uint64_t xbar[5];                               // inboxes: 0 qkv, 1 att, 2 y1 (+stat1), 3 reduce-scatter, 4 y2 (+stat2)
};

template <typename T, int NCH, int CS>
__global__ void __launch_bounds__(NT, 1) gpt_decode_hx_kernel(const GptParams p, const int n_steps, const unsigned tag_base,
                                                               uint2* const ll_buf, const T* __restrict__ pack,
                                                               const T* __restrict__ hpack, const int HR, unsigned* const resident) {
  using Lo = HxLayout<NCH * 256, CS>;
  constexpr int D = NCH * 256, RC = Lo::RC, HC = Lo::HC, KO = Lo::KO;
  constexpr int RW = (RC + 31) / 32;              // warps that own rows in the all-reduce (one row per thread)
  constexpr unsigned YBYTES = D * 4u + CS * RW * 8u;   // one fill of a y inbox: D values + (sum, sum of squares) of every rank's warps
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ HxShared sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, L = p.L, V = p.V, S = p.S;
  constexpr int NCL = D / (8 * CS);         // clusters (= H / HC) = partials that meet in L2
  const unsigned rank = cluster_rank();
  const int cl = blockIdx.x / CS;           // cluster index
  const int h = cl * HC + (int)rank / 4;    // this CTA's head
  const int s4 = (int)rank & 3;             // ... and which quarter of the head's work it does
  const int j = blockIdx.x;                 // = CS cl + rank: owner of hidden units [32 j, 32 j + 32)
  const int NC = gridDim.x;

  // shared memory: weight sections (T) | head rows (T) | xbuf[D] x1buf[D] ybox1[D] ybox2[D] rs_in[CS][RC] m2stage[D]
  //                | sampler scratch (CTA 0)
  T* wsec = reinterpret_cast<T*>(smem_raw);
  T* headw = wsec + Lo::BLOB;
  float* xbuf = reinterpret_cast<float*>(headw + (size_t)HR * D);   // layer input x (split layout): residual of the attention half
  float* x1buf = xbuf + D;                  // LN1 output (split layout): residual of the MLP half
  float* ybox1 = x1buf + D;                 // inbox: y1 (plain layout)
  float* ybox2 = ybox1 + D;                 // inbox: y2
  float* rs_in = ybox2 + D;                 // inbox: [CS source ranks][RC] MLP-down partial rows of this rank
  float* m2stage = rs_in + CS * RC;         // this CTA's D partial outputs of the MLP-down before they are pushed
  T* kbuf = reinterpret_cast<T*>(m2stage + D);   // prefetched K rows of this CTA's chunk of cached positions [QMAX][32] ...
  T* vbuf = kbuf + QMAX * GSV_HEAD_DIM;          // ... and V rows
  float* samp = reinterpret_cast<float*>(vbuf + QMAX * GSV_HEAD_DIM);   // sampler scratch (GSV_SAMPLE_SMEM_FLOATS), CTA 0 only

  // LL exchange areas ({value, tag} words): P1[NCL][D] | P2[NCL][D] | logits[VOCAB_MAX] | xin[D] | status
  uint2* P1 = ll_buf;
  uint2* P2 = P1 + (size_t)NCL * D;
  uint2* LLlogit = P2 + (size_t)NCL * D;
  uint2* LLxin = LLlogit + GSV_VOCAB_MAX;
  uint2* LLstat = LLxin + D;

  // ---- which sequence: the first active slot ----
  if (tid < 32) {
    const int flag = tid < p.slots ? ld_cg(p.active + tid) : 0;
    const int flag2 = tid + 32 < p.slots ? ld_cg(p.active + tid + 32) : 0;        // slot table of up to 64
    const unsigned m = __ballot_sync(0xffffffffu, flag != 0);
    const unsigned m2 = __ballot_sync(0xffffffffu, flag2 != 0);
    if (tid == 0) {
      const int first = m ? (__ffs(m) - 1) : (m2 ? 32 + __ffs(m2) - 1 : -1);
      sh.slot = first;
      sh.kv = first >= 0 ? ld_cg(p.kv_len + first) : 0;
    }
  }
  __syncthreads();
  const int slot = sh.slot;
  if (slot < 0) { if (tid == 0) atomicAdd(resident, 1u); return; }     // uniform over the grid (counted: gsv_gpt_wait_resident)
  int kv = sh.kv;

  const T* const blob0 = pack + (size_t)j * Lo::BLOB;                       // layer 0 blob of this CTA
  const size_t blob_lstride = (size_t)NC * Lo::BLOB;
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
    mbar_init(&sh.kbar, 1);
    for (int i = 0; i < 5; ++i) mbar_init(&sh.xbar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < 4; ++s) issue_sec(s, 0);
    const unsigned hbytes = (unsigned)HR * D * 2u;
    mbar_expect_tx(&sh.hbar, hbytes);
    bulk_g2s(headw, hpack + (size_t)j * HR * D, hbytes, &sh.hbar);
    mbar_expect_tx(&sh.xbar[0], 4 * 24 * 4u);
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
  if (tid == 0) atomicAdd(resident, 1u);    // gsv_gpt_wait_resident: work of other streams is held back until every CTA is here

  unsigned tag = tag_base;
  const int sub = lane & 3, pg = lane >> 2;
  int gl = 0;                               // layers done in this launch (prefetch cursor)
  const int total_layers = n_steps * L;
  float xv[NCH * 8];                        // the layer input in dot-product order (every warp holds all of it)

  // LayerNorm of a gathered vector, one element per thread; the statistics came with the data (no reduction)
  auto gathered_ln = [&](const float* ybox, const float (*stat)[2], const T* g, const T* b, float* dst) {
    if (tid < D) {
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int rr = 0; rr < CS * RW; ++rr) { s += stat[rr][0]; q += stat[rr][1]; }
      const float mean = s * (1.f / (float)D);
      const float rstd = rsqrtf(fmaxf(q * (1.f / (float)D) - mean * mean, 0.f) + 1e-5f);
      dst[split_pos(tid, D)] = (ybox[tid] - mean) * rstd * Elem<T>::to_f(g[tid]) + Elem<T>::to_f(b[tid]);
    }
    __syncthreads();
    load_x<NCH>(dst, lane, xv);
  };
  // all-reduce over the clusters: one thread per row of this rank's RC rows: the row's word of every cluster's partial is
  // polled (all in flight), summed in cluster order, + residual + bias; the rows travel to every rank of the cluster as
  // float4 (a quad of lanes gathers its 4 rows by shuffles, each lane pushes to CS/4 targets) together with each warp's
  // (sum, sum of squares).  No staging, no barrier.
  auto all_reduce = [&](const uint2* P, unsigned t, const float* resid, const T* bias, float* ybox, float (*stat)[2], uint64_t* bar) {
    if (warp < RW) {
      const int i = tid;
      const bool valid = i < RC;
      float y = 0.f;
      if (valid) {
        const uint2* src = P + RC * rank + i;
        float acc = 0.f;
        // every cluster's word of this row in flight (strong loads), then all tags are checked; words that have arrived
        // are simply read again (they stay put until the next layer)
        uint2 w[NCL];
        bool ok;
        do {
#pragma unroll
          for (int c = 0; c < NCL; ++c) w[c] = ll_peek(src + (size_t)c * D);
          ok = true;
#pragma unroll
          for (int c = 0; c < NCL; ++c) ok = ok && (w[c].y == t);
        } while (!ok);
#pragma unroll
        for (int c = 0; c < NCL; ++c) acc += __uint_as_float(w[c].x);
        y = acc + resid[split_pos(RC * rank + i, D)] + Elem<T>::to_f(bias[i]);
      }
      mark(p, 43);
      const int q0 = lane & ~3, tq = lane & 3;
      float4 v4;
      v4.x = __shfl_sync(0xffffffffu, y, q0);
      v4.y = __shfl_sync(0xffffffffu, y, q0 + 1);
      v4.z = __shfl_sync(0xffffffffu, y, q0 + 2);
      v4.w = __shfl_sync(0xffffffffu, y, q0 + 3);
      if (valid) {
#pragma unroll
        for (int tg = 0; tg < CS / 4; ++tg)
          st_async_v4f(ybox + RC * rank + (i & ~3), bar, (unsigned)(tq + 4 * tg), v4);
      }
      float s = y, q = y * y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
      if (lane < CS) st_async_v2f(&stat[rank * RW + warp][0], bar, (unsigned)lane, s, q);
    }
  };

  // The head's cached positions [0, kv) are dealt to its 4 CTAs in contiguous chunks of Q = ceil(kv / 4): CTA s4 owns
  // [s4 Q, min(kv, s4 Q + Q)), so its K rows (and its V rows) are ONE contiguous run of the cache: one thread requests
  // them a layer ahead with two bulk copies into shared memory.  (Requested by every lane into registers or with cp.async
  // they stalled whatever the requesting warps did next -- shuffles and shared-memory loads of a warp wait behind its
  // pending global loads: +2.5 us in the attention merge.)  Rows past QMAX (kv > 1024) are read from L2 in the loop.
  unsigned kpar = 0;
  auto kv_request = [&](int layer, int kv_then) {
    const int Qn = (kv_then + 3) >> 2;
    const int first = s4 * Qn;
    const int n = min(min(kv_then, first + Qn) - first, QMAX);
    if (n > 0) {
      const size_t hb = ((size_t)(layer * p.slots + slot) * H + h) * (size_t)S * GSV_HEAD_DIM + (size_t)first * GSV_HEAD_DIM;
      const unsigned bytes = (unsigned)n * GSV_HEAD_DIM * (unsigned)sizeof(T);
      mbar_expect_tx(&sh.kbar, 2u * bytes);
      bulk_g2s(kbuf, reinterpret_cast<const T*>(p.kc) + hb, bytes, &sh.kbar);
      bulk_g2s(vbuf, reinterpret_cast<const T*>(p.vc) + hb, bytes, &sh.kbar);
    } else {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sh.kbar)) : "memory");   // nothing to copy: complete the phase
    }
  };

#pragma unroll 1
  for (int step = 0; step < n_steps; ++step) {
    load_x<NCH>(xbuf, lane, xv);
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      mark(p, 1);
      const size_t head_base = ((size_t)(l * p.slots + slot) * H + h) * (size_t)S * GSV_HEAD_DIM;
      const T* kb = reinterpret_cast<const T*>(p.kc) + head_base + sub * 8;
      const T* vb = reinterpret_cast<const T*>(p.vc) + head_base + sub * 8;
      // cached positions of this CTA: [first, first + n_mine); the newest position (p = kv) is scored by the last quarter
      const int Qc = (kv + 3) >> 2;
      const int first = s4 * Qc;
      const int n_mine = max(0, min(kv, first + Qc) - first);
      const bool owner = s4 == 3;
      if (gl == 0 && tid == NT - 32) kv_request(l, kv);     // the first layer of a launch requests its own rows
      // ================= q/k/v rows of this CTA (24 of the head's 96), all-gathered between the head's 4 CTAs =================
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
      if (tid < 24) {                        // 6 float4 per target x the head's 4 CTAs; value g = 24 s4 + i lands at qkv_in[g]
        const int tgt = ((int)rank & ~3) + tid / 6, q4 = tid % 6;
        st_async_v4f(sh.qkv_in + 24 * s4 + q4 * 4, &sh.xbar[0], (unsigned)tgt, *reinterpret_cast<const float4*>(sh.qkv_stage + q4 * 4));
      }
      mark(p, 2);
      // (waiting warps park at the barrier below: a warp spinning on an mbarrier competes with the working warps for
      //  the shared-memory pipeline that also carries their LDS / SHFL / LDG)
      if (tid < 3 * GSV_HEAD_DIM) {
        mbar_wait(&sh.xbar[0], xpar);
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
      if (tid == 0) mbar_expect_tx(&sh.xbar[0], 4 * 24 * 4u);
      mark(p, 3);
      // ================= attention over this CTA's cached positions; partial pushed to the whole cluster =================
      {
        const int n_act = min(NWARP, (n_mine + 7) >> 3);      // warps that hold cached positions in the first pass
        if (warp == NWARP - 1 && owner) {                      // score of the newest position (k from shared memory)
          const float sn = warp_allsum(sh.q[lane] * sh.kn[lane]);
          if (lane == 0) sh.snew = sn;
        }
        mbar_wait(&sh.kbar, kpar);                            // K/V rows (requested a layer ago) have landed
        kpar ^= 1u;
        if (warp < n_act) {
          float q[8];
#pragma unroll
          for (int jq = 0; jq < 8; ++jq) q[jq] = sh.q[sub * 8 + jq];
          float mg = GSV_NEG_INF, lsum = 0.f, o[8];
#pragma unroll
          for (int jq = 0; jq < 8; ++jq) o[jq] = 0.f;
          int pass = 0;
#pragma unroll 1
          for (int m0 = warp * 8; m0 < n_mine; m0 += TILE, ++pass) {
            const int m = m0 + pg;
            const bool ok = m < n_mine;
            uint4 kr = make_uint4(0, 0, 0, 0), vr = kr;
            if (ok) {
              if (m < QMAX) {
                kr = reinterpret_cast<const uint4*>(kbuf)[m * 4 + sub];
                vr = reinterpret_cast<const uint4*>(vbuf)[m * 4 + sub];
              } else {
                kr = ld_cg16(kb + (size_t)(first + m) * GSV_HEAD_DIM);
                vr = ld_cg16(vb + (size_t)(first + m) * GSV_HEAD_DIM);
              }
            }
            float kf[8], vf[8], sc_ = 0.f;
            unpack8<T>(kr, kf);
            unpack8<T>(vr, vf);
#pragma unroll
            for (int jq = 0; jq < 8; ++jq) sc_ = fmaf(q[jq], kf[jq], sc_);
            sc_ += __shfl_xor_sync(0xffffffffu, sc_, 1);
            sc_ += __shfl_xor_sync(0xffffffffu, sc_, 2);
            if (ok) {
              const float mn = fmaxf(mg, sc_);
              const float sc = exp2f(mg - mn), pr = exp2f(sc_ - mn);
              lsum = fmaf(lsum, sc, pr);
#pragma unroll
              for (int jq = 0; jq < 8; ++jq) o[jq] = fmaf(pr, vf[jq], o[jq] * sc);
              mg = mn;
            }
          }
          float m = mg;
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
          const float rs = (mg > GSV_NEG_INF) ? exp2f(mg - m) : 0.f;
          lsum *= rs;
#pragma unroll
          for (int jq = 0; jq < 8; ++jq) o[jq] *= rs;
#pragma unroll
          for (int off = 4; off < 32; off <<= 1) {
            lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
#pragma unroll
            for (int jq = 0; jq < 8; ++jq) o[jq] += __shfl_xor_sync(0xffffffffu, o[jq], off);
          }
          if (lane < 4) {
            if (sub == 0) { sh.wpart[warp][0] = m; sh.wpart[warp][1] = lsum; }
#pragma unroll
            for (int jq = 0; jq < 8; ++jq) sh.wpart[warp][2 + sub * 8 + jq] = o[jq];
          }
        }
        mark(p, 40);
        // the next layer's rows (the next token's first layer sees one more cached position)
        __syncthreads();
        // every warp is done with the K/V rows: request the next layer's (the next token's first layer sees one more position)
        if (tid == NT - 32 && gl + 1 < total_layers) kv_request(l + 1 < L ? l + 1 : 0, l + 1 < L ? kv : kv + 1);
        mark(p, 41);
        if (warp == 0) {
          // merge the active warps' partials and, as one more partial (m = its score, l = 1, o = v), the newest position
          // if this CTA owns it: lane w < n_act holds partial w's scale factor, lane = output dimension in the sums
          float mw = GSV_NEG_INF, lw = 0.f;
          if (lane < n_act) { mw = sh.wpart[lane][0]; lw = sh.wpart[lane][1]; }
          else if (lane == n_act && owner) { mw = sh.snew; lw = 1.f; }
          const float M = warp_max(mw);
          const float scw = mw > GSV_NEG_INF ? exp2f(mw - M) : 0.f;
          const float Ls = warp_allsum(scw * lw);
          float oa = 0.f;
          for (int w2 = 0; w2 < n_act; ++w2) oa = fmaf(__shfl_sync(0xffffffffu, scw, w2), sh.wpart[w2][2 + lane], oa);
          if (owner) oa = fmaf(__shfl_sync(0xffffffffu, scw, n_act), sh.vn[lane], oa);
          if (lane == 0) { sh.att_out[0] = M; sh.att_out[1] = Ls; sh.att_out[2] = 0.f; sh.att_out[3] = 0.f; }
          sh.att_out[4 + lane] = oa;
          __syncwarp();
          mark(p, 42);
          for (int e = lane; e < CS * (ATT_W / 4); e += 32) {     // 9 float4 per target x CS targets
            const int tgt = e / (ATT_W / 4), q4 = e % (ATT_W / 4);
            st_async_v4f(&sh.att_in[rank][q4 * 4], &sh.xbar[1], (unsigned)tgt, *reinterpret_cast<const float4*>(sh.att_out + q4 * 4));
          }
        }
      }
      mark(p, 4);
      if (warp == 0) {
        mbar_wait(&sh.xbar[1], xpar);
#pragma unroll
        for (int hh = 0; hh < HC; ++hh) {     // the 4 partials of the cluster's hh-th head
          float M = GSV_NEG_INF;
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) M = fmaxf(M, sh.att_in[4 * hh + rr][0]);
          float Ls = 0.f, oa = 0.f;
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const float mr = sh.att_in[4 * hh + rr][0];
            const float sc = mr > GSV_NEG_INF ? exp2f(mr - M) : 0.f;
            Ls = fmaf(sh.att_in[4 * hh + rr][1], sc, Ls);
            oa = fmaf(sh.att_in[4 * hh + rr][4 + lane], sc, oa);
          }
          sh.att[hh * GSV_HEAD_DIM + lane] = oa / Ls;
        }
      }
      __syncthreads();                        // att staged; att_in fully read
      if (tid == 0) mbar_expect_tx(&sh.xbar[1], CS * ATT_W * 4u);
      mark(p, 5);
      // ================= out-projection partial: rows [RC rank, RC rank + RC) of Wo[:, cluster's columns] . att =================
      mbar_wait(&sh.wbar[1], wpar);
      tag += 1;
      {
        constexpr int LPR = KO / 8;           // lanes per row (4, 8, 16)
        constexpr int RPP = 32 / LPR;         // rows per warp pass (8, 4, 2)
        constexpr int RPW = RC / NWARP > 0 ? RC / NWARP : 1;   // rows per warp
        const T* w1s = wsec + Lo::S1;
        const int kc8 = lane % LPR, rin = lane / LPR;
        float af[8];
#pragma unroll
        for (int jq = 0; jq < 8; ++jq) af[jq] = sh.att[kc8 * 8 + jq];
#pragma unroll
        for (int r0 = 0; r0 < RPW; r0 += RPP) {
          const int rloc = r0 + rin;
          const int row = warp * RPW + rloc;
          const bool valid = rloc < RPW && row < RC;
          float wf[8];
          unpack8<T>(reinterpret_cast<const uint4*>(w1s + (size_t)(valid ? row : 0) * KO)[kc8], wf);
          float a = 0.f;
#pragma unroll
          for (int jq = 0; jq < 8; ++jq) a = fmaf(wf[jq], af[jq], a);
#pragma unroll
          for (int o = 1; o < LPR; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
          if (valid && kc8 == 0) ll_store(P1 + (size_t)cl * D + RC * rank + row, a, tag);
        }
      }
      mark(p, 6);
      // ================= all-reduce 1: y1 = x + bo + sum of the cluster partials ; x1 = LN1(y1) =================
      all_reduce(P1, tag, xbuf, wsec + Lo::S1 + RC * KO, ybox1, sh.stat1, &sh.xbar[2]);
      mark(p, 7);
      if (warp == 0) mbar_wait(&sh.xbar[2], xpar);
      __syncthreads();
      mark(p, 45);
      gathered_ln(ybox1, sh.stat1, wsec + Lo::S1 + RC * KO + RC, wsec + Lo::S1 + RC * KO + RC + D, x1buf);
      // (barrier inside: ybox1 / stat1 / the LayerNorm parameters of S1 fully read)
      if (tid == 0) {
        mbar_expect_tx(&sh.xbar[2], YBYTES);
        if (gl + 1 < total_layers) issue_sec(1, (l + 1) % L);
      }
      mark(p, 8);
      // ================= MLP-up: this CTA's 32 hidden units (rows warp and warp + 16) =================
      mbar_wait(&sh.wbar[2], wpar);
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
      __syncthreads();
      if (tid == 0 && gl + 1 < total_layers) issue_sec(2, (l + 1) % L);
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
      if (tid < D / 4) {                      // rows 4 tid .. 4 tid + 3 go to rank (4 tid) / RC, slot [this rank][(4 tid) % RC]
        const int row = 4 * tid, tgt = row / RC;
        st_async_v4f(rs_in + rank * RC + (row - tgt * RC), &sh.xbar[3], (unsigned)tgt, *reinterpret_cast<const float4*>(m2stage + row));
      }
      mark(p, 10);
      tag += 1;
      if (tid < RC) {
        mbar_wait(&sh.xbar[3], xpar);
        float s = rs_in[tid];
#pragma unroll
        for (int rr = 1; rr < CS; ++rr) s += rs_in[rr * RC + tid];
        ll_store(P2 + (size_t)cl * D + RC * rank + tid, s, tag);
      }
      mark(p, 11);
      // ================= all-reduce 2: y2 = x1 + b2 + sum of the cluster partials ; x = LN2(y2) =================
      all_reduce(P2, tag, x1buf, wsec + Lo::S3 + D * 32, ybox2, sh.stat2, &sh.xbar[4]);
      mark(p, 12);
      if (warp == 0) mbar_wait(&sh.xbar[4], xpar);
      __syncthreads();
      gathered_ln(ybox2, sh.stat2, wsec + Lo::S3 + D * 32 + RC, wsec + Lo::S3 + D * 32 + RC + D, xbuf);
      wpar ^= 1u;
      xpar ^= 1u;
      gl += 1;
      if (tid == 0) {                         // behind the barrier inside gathered_ln: rs_in / ybox2 / stat2 / S3 fully read
        mbar_expect_tx(&sh.xbar[3], D * 4u);
        mbar_expect_tx(&sh.xbar[4], YBYTES);
        if (gl < total_layers) issue_sec(3, (l + 1) % L);
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
    mark(p, 20);
    // ================= sampling in CTA 0; next input and status published for everyone =================
    const unsigned tag_logits = tag;
    tag += 1;
    if (j == 0) {
      mark_sampler(p, 58);
      SamplePre pre;                          // the sampler's own global reads, issued before the wait for the logits
      sample_prefetch<T>(p, slot, kv + 1, pre);
      {                                       // every word of this thread in flight (GSV_VOCAB_MAX / NT = 4)
        uint2 w[GSV_VOCAB_MAX / NT];
#pragma unroll
        for (int u = 0; u < GSV_VOCAB_MAX / NT; ++u) w[u] = make_uint2(0u, ~tag_logits);
        bool ok;
        do {
#pragma unroll
          for (int u = 0; u < GSV_VOCAB_MAX / NT; ++u)
            if (tid + u * NT < V) w[u] = ll_peek(LLlogit + tid + u * NT);
          ok = true;
#pragma unroll
          for (int u = 0; u < GSV_VOCAB_MAX / NT; ++u) ok = ok && (tid + u * NT >= V || w[u].y == tag_logits);
        } while (!ok);
#pragma unroll
        for (int u = 0; u < GSV_VOCAB_MAX / NT; ++u) if (tid + u * NT < V) samp[tid + u * NT] = __uint_as_float(w[u].x);
      }
      __syncthreads();
      mark_sampler(p, 59);
      SampleLL io;
      io.preloaded = true;
      io.xin_ll = LLxin;
      io.status_ll = LLstat;
      io.tag = tag;
      io.kv_len = kv + 1;
      io.xin_smem = nullptr;
      io.alive_smem = nullptr;
      sample_slot<T>(p, slot, samp, &io, &pre);
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
  if (gl < total_layers) mbar_wait(&sh.kbar, kpar);
  cluster_sync_all();
}

template <int NCH, int CS>
size_t hx_smem_bytes(int HR) {
  using Lo = HxLayout<NCH * 256, CS>;
  constexpr int D = NCH * 256;
  return (size_t)Lo::BLOB * 2 + (size_t)HR * D * 2 + sizeof(float) * ((size_t)6 * D + ((GSV_SAMPLE_SMEM_FLOATS + 3) & ~3)) + (size_t)2 * QMAX * GSV_HEAD_DIM * 2;
}

template <typename T, int NCH, int CS>
int launch_hx_t(gsv_gpt_ctx* ctx, int n_steps, cudaStream_t st) {
  using Lo = HxLayout<NCH * 256, CS>;
  constexpr int D = NCH * 256;
  const int H = ctx->p.H, L = ctx->p.L, V = ctx->p.V, NC = H * 4;
  const int HR = (V + NC - 1) / NC;
  void* fn = (void*)gpt_decode_hx_kernel<T, NCH, CS>;
  const size_t bytes = hx_smem_bytes<NCH, CS>(HR);
  if (!ctx->hx_pack) {
    void *pk = nullptr, *hp = nullptr;
    GSV_CUDA(cudaMalloc(&pk, (size_t)L * NC * Lo::BLOB * sizeof(T)));
    GSV_CUDA(cudaMalloc(&hp, (size_t)NC * HR * D * sizeof(T)));
    hx_pack_kernel<T, D, CS><<<ctx->num_sms * 8, 256, 0, st>>>(ctx->p, reinterpret_cast<T*>(pk));
    hx_pack_head_kernel<T><<<ctx->num_sms, 256, 0, st>>>(reinterpret_cast<const T*>(ctx->p.w_head), reinterpret_cast<T*>(hp), NC, HR, V, D);
    GSV_CUDA(cudaGetLastError());
    ctx->hx_pack = pk;
    ctx->hx_head_pack = hp;
    ctx->launches += 2;
  }
  GSV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  if (CS > 8) GSV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
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
    ctx->hx_clusters_ok = (e == cudaSuccess && n >= NC / CS) ? 1 : -1;
    if (e != cudaSuccess) cudaGetLastError();
  }
  if (ctx->hx_clusters_ok < 0) { gsv_set_error("hx decode kernel: %d clusters of %d CTAs are not co-resident on this device", NC / CS, CS); return GSV_ERR_STATE; }
  unsigned* resident = ctx->hx_resident;
  ctx->hx_resident_expected += (unsigned)NC;
  void* args[] = {&p, &ns, &tag_base, &buf, &pk, &hp, &hr, &resident};
  GSV_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  ctx->launches += 1;
  return GSV_OK;
}

// cluster size: 8 (two heads per cluster) measured best on B200 -- 168 us/token against 183 (16) and 200 (4): larger
// clusters poll fewer words in L2 but push every attention partial to more CTAs; GSV_HX_CS = 4 | 8 | 16 pins it (A/B runs)
int hx_cluster_size(gsv_gpt_ctx* ctx) {
  if (ctx->hx_cs == 0) {
    int cs = 8;
    const char* e = getenv("GSV_HX_CS");
    if (e) { const int v = atoi(e); if (v == 4 || v == 8 || v == 16) cs = v; }
    while (cs > 4 && (4 * ctx->p.H) % cs != 0) cs >>= 1;
    ctx->hx_cs = cs;
  }
  return ctx->hx_cs;
}

template <typename T>
int launch_hx(gsv_gpt_ctx* ctx, int n_steps, cudaStream_t st) {
  const int cs = hx_cluster_size(ctx);
  if (ctx->p.d == 512) {
    if (cs == 16) return launch_hx_t<T, 2, 16>(ctx, n_steps, st);
    if (cs == 8) return launch_hx_t<T, 2, 8>(ctx, n_steps, st);
    return launch_hx_t<T, 2, 4>(ctx, n_steps, st);
  }
  if (ctx->p.d == 256) {
    if (cs == 16) return launch_hx_t<T, 1, 16>(ctx, n_steps, st);
    if (cs == 8) return launch_hx_t<T, 1, 8>(ctx, n_steps, st);
    return launch_hx_t<T, 1, 4>(ctx, n_steps, st);
  }
  return GSV_ERR_ARG;
}

}  // namespace

// One thread spins (bounded) until every CTA of the last head-cluster launch is resident.
__global__ void hx_gate_kernel(const unsigned* __restrict__ resident, unsigned expected) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(resident) : "memory");
    if ((int)(v - expected) >= 0) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 3000000ull) break;          // 3 ms: never hold a stream hostage
    __nanosleep(200);
  }
}

int gsv_gpt_hx_gate(gsv_gpt_ctx* ctx, cudaStream_t st) {
  if (!ctx->hx_resident || ctx->hx_resident_expected == 0) return GSV_OK;
  hx_gate_kernel<<<1, 1, 0, st>>>(ctx->hx_resident, ctx->hx_resident_expected);
  GSV_CHECK_LAUNCH();
  return GSV_OK;
}

// words of the LL exchange areas this kernel needs: P1[<= H][D] + P2[<= H][D] + logits + xin + status
size_t gsv_gpt_hx_buffer_words(const gsv_gpt_ctx* ctx) {
  return (size_t)2 * ctx->p.H * ctx->p.d + GSV_VOCAB_MAX + ctx->p.d + 8;
}

bool gsv_gpt_hx_supported(const gsv_gpt_ctx* ctx, int live_slots, int n_steps) {
  const GptParams& p = ctx->p;
  const bool shape = (p.d == 256 || p.d == 512) && p.H * GSV_HEAD_DIM == p.d && p.F == 4 * p.d && p.V <= GSV_VOCAB_MAX;
  return shape && live_slots == 1 && p.H * 4 <= ctx->num_sms && ctx->hx_clusters_ok >= 0 &&
         (long long)n_steps * (2 * p.L + 2) < 65000;
}

int gsv_gpt_decode_hx_launch(gsv_gpt_ctx* ctx, int n_steps, cudaStream_t st) {
  if (ctx->dims.dtype == GSV_F16) return launch_hx<__half>(ctx, n_steps, st);
  return launch_hx<__nv_bfloat16>(ctx, n_steps, st);
}
