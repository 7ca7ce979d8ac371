// gpt_decode_cl8.cu -- cluster decode kernel for up to EIGHT sequences per cluster on the tensor cores.
//
// Same arithmetic as the other decode kernels (reference t2s_model.py:67-105, 129-143, 442-456) and the same
// skeleton as gpt_decode_cl.cu / gpt_decode_cln.cu (one CTA per attention head, push exchanges completing on the
// receiver's mbarrier, per-warp bulk-copy weight ring).  What changes: the cluster's live sequences are the N = 8
// columns of mma.sync m16n8k16 tiles whose M rows are the warp's weight rows, so a weight row that has been streamed
// from HBM is multiplied with all eight inputs by ONE instruction stream (gpt_decode_cln.cu pays the full CUDA-core
// dot product per sequence: 1.8x / 2.85x the step time for 2 / 4 sequences).  Consequences:
//   * activations cross CTAs in the storage type (bf16 / fp16 pairs, as the reference rounds them) and land directly
//     in the layout the B operand is read from ([sequence][k], rows padded by 8 elements: conflict-free fragment
//     loads); the residual stream (y1, y2, next input) stays fp32;
//   * weight units sit in padded ring slots (row stride D*2 + 16 bytes) so that the A fragments of the 2..4 rows of
//     a batch come from different banks;
//   * LayerNorm of sequence n is done once per CTA by warp n; attention of sequence n by warps 2n and 2n+1;
//   * results are staged per warp and pushed with 8- / 16-byte st.async (one per (target CTA, sequence)).
// The accumulator fragment gives thread (g = lane / 4, t = lane % 4) row g of the batch for sequences 2t and 2t + 1.
#include <type_traits>

#include "gpt_cluster_common.cuh"

namespace {

constexpr int NB8 = 8;                   // sequences per cluster
constexpr int RING8 = 6;                 // weight units in flight per warp (batches are at most 4 units)
constexpr int XPAD = 8;                  // padding elements per staged activation row

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

// 8 / 16 bytes into the same shared-memory location of CTA `rank`, completing that many bytes on its copy of `bar`
__device__ __forceinline__ void st_async_v2(void* local_ptr, uint64_t* local_bar, unsigned rank, unsigned x, unsigned y) {
  unsigned ra, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(local_bar)), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(ra), "r"(x), "r"(y), "r"(rb) : "memory");
}
__device__ __forceinline__ void st_async_v4(void* local_ptr, uint64_t* local_bar, unsigned rank, uint4 v) {
  unsigned ra, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(local_bar)), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(ra), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rb) : "memory");
}

struct Cl8Shared {
  float qs[NB8][GSV_HEAD_DIM], kn[NB8][GSV_HEAD_DIM], vn[NB8][GSV_HEAD_DIM];   // scaled q, new k / v of this head per sequence
  float apart[NWARP][GSV_HEAD_DIM + 2];     // attention partials per warp: m, l, o[32]
  float xres[NB8][GSV_HEAD_DIM];            // residual rows of this CTA for the out-proj (layer input x)
  float xres1[NB8][GSV_HEAD_DIM];           // ... and for the MLP-down (x1)
  __align__(16) float stage[NWARP][NB8][4]; // per-warp staging of a phase's results before they are pushed (16 B per sequence)
  float alive[NB8];                         // pushed by the sampler CTAs together with the next inputs
  int alive_i;
  int slot[NB8], kv[NB8];
  uint64_t wbar[NWARP][RING8];
  uint64_t xbar[4];                         // inboxes: 0 inA (fp32: xin / y1 / y2), 1 att, 2 h, 3 logits (CTA n for sequence n)
};

template <typename T, int NCH>
__global__ void __launch_bounds__(NT, 1) gpt_decode_cl8_kernel(const GptParams p, const int n_steps) {
  extern __shared__ __align__(16) float smem[];
  __shared__ Cl8Shared sh;
  constexpr int D = NCH * 256, F = 4 * D;
  constexpr int LDX = D + XPAD, LDH = F + XPAD;          // element strides of the staged activation rows
  constexpr int USTRIDE = D * 2 + 16;                     // bytes between weight ring slots
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;                  // mma fragment coordinates
  const int H = p.H, L = p.L, V = p.V, S = p.S;
  const unsigned rank = cluster_rank();                  // = head index
  const int cid = blockIdx.x / H;                         // cluster index: serves live sequences [cid*8, cid*8 + 8)

  // shared memory: inA[8][D] fp32 | xa[8][LDX] T | attb[8][LDX] T | hb[8][LDH] T | xin_s[D] | sampler scratch (+ logits) | weight ring
  float* inA = smem;
  T* xa = reinterpret_cast<T*>(inA + NB8 * D);
  T* attb = xa + NB8 * LDX;
  T* hb = attb + NB8 * LDX;
  float* xin_s = reinterpret_cast<float*>(hb + NB8 * LDH);
  float* samp = xin_s + D;
  unsigned char* ring = reinterpret_cast<unsigned char*>(samp + ((GSV_SAMPLE_SMEM_FLOATS + 3) & ~3)) + (size_t)warp * RING8 * USTRIDE;

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

  // ---- which sequences: the active slots number cid*8 .. cid*8 + 7 ----
  if (tid < 32) {
    const int flag = tid < p.slots ? ld_cg(p.active + tid) : 0;
    const unsigned m = __ballot_sync(0xffffffffu, flag != 0);
    const int pos = flag ? __popc(m & ((1u << tid) - 1u)) - cid * NB8 : -1;
    if (tid < NB8) { sh.slot[tid] = -1; sh.kv[tid] = 0; sh.alive[tid] = 0.f; }
    __syncwarp();
    if (flag && pos >= 0 && pos < NB8) { sh.slot[pos] = tid; sh.kv[pos] = ld_cg(p.kv_len + tid); sh.alive[pos] = 1.f; }
  }
  // zero the staged operands once: columns of sequences that are not live must hold finite values
  for (int i = tid; i < (NB8 * LDX * 2 + NB8 * LDH) / 2; i += NT) reinterpret_cast<unsigned*>(xa)[i] = 0u;
  __syncthreads();
  unsigned livemask = 0;                                  // bit n: sequence n of this cluster is live (uniform across its CTAs)
#pragma unroll
  for (int n = 0; n < NB8; ++n) livemask |= (sh.slot[n] >= 0 ? 1u : 0u) << n;
  if (livemask == 0) return;
  int na = __popc(livemask);

  // ---- weight unit sequence of this warp: 6 QKV rows, 2 O rows, 8 MLP-up rows, then the MLP-down rows quarter by
  //      quarter (row 0 quarter q, row 1 quarter q) so that a batch of 2 units is one K-quarter of both rows ----
  auto unit_src = [&](int l, int u) -> const T* {
    if (u < 6) {
      const int rr = warp + NWARP * u;
      const int row = (rr >> 5) * D + (int)rank * GSV_HEAD_DIM + (rr & 31);
      return Wqkv + ((size_t)l * 3 * D + row) * D;
    }
    u -= 6;
    if (u < 2) return Wo + ((size_t)l * D + rank * GSV_HEAD_DIM + warp * 2 + u) * D;
    u -= 2;
    if (u < 8) return W1 + ((size_t)l * F + rank * (4 * GSV_HEAD_DIM) + warp * 8 + u) * D;
    u -= 8;
    return W2 + ((size_t)l * D + rank * GSV_HEAD_DIM + warp * 2 + (u & 1)) * F + (size_t)(u >> 1) * D;
  };
  int iss_l = 0, iss_u = 0, use_i = 0;
  unsigned use_par = 0;
  constexpr unsigned UNIT_BYTES = D * (unsigned)sizeof(T);
  auto issue_at = [&](int slot_i, int ahead) {
    int u = iss_u + ahead, l = iss_l;
    if (u >= UNITS_PER_LAYER) { u -= UNITS_PER_LAYER; l = l + 1 == L ? 0 : l + 1; }
    mbar_expect_tx(&sh.wbar[warp][slot_i], UNIT_BYTES);
    bulk_g2s(ring + (size_t)slot_i * USTRIDE, unit_src(l, u), UNIT_BYTES, &sh.wbar[warp][slot_i]);
  };
  auto advance_cursor = [&](int n) {
    iss_u += n;
    if (iss_u >= UNITS_PER_LAYER) { iss_u -= UNITS_PER_LAYER; iss_l = iss_l + 1 == L ? 0 : iss_l + 1; }
  };
  // One batch of NV (<= 4) weight units against a staged operand: acc[0], acc[1] += row g of the batch (g < NV) times
  // sequences 2t, 2t+1 over K = D starting at element `koff` of every sequence's row (row stride `ldb`).
  auto mma_batch = [&](auto nv_tag, const T* xb, int ldb, int koff, float (&acc)[4]) {
    constexpr int NV = decltype(nv_tag)::value;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int si = use_i + i;
      unsigned par = use_par;
      if (si >= RING8) { si -= RING8; par ^= 1u; }
      mbar_wait(&sh.wbar[warp][si], par);
    }
    int sg = use_i + (g < NV ? g : 0);
    if (sg >= RING8) sg -= RING8;
    const T* arow = reinterpret_cast<const T*>(ring + (size_t)sg * USTRIDE) + 2 * t;
    const T* brow = xb + (size_t)g * ldb + koff + 2 * t;
#pragma unroll 4
    for (int ks = 0; ks < D / 16; ++ks) {
      unsigned a0 = 0u, a2 = 0u;
      if (g < NV) {
        a0 = *reinterpret_cast<const unsigned*>(arow + ks * 16);
        a2 = *reinterpret_cast<const unsigned*>(arow + ks * 16 + 8);
      }
      const unsigned b0 = *reinterpret_cast<const unsigned*>(brow + ks * 16);
      const unsigned b1 = *reinterpret_cast<const unsigned*>(brow + ks * 16 + 8);
      Mma16816<T>::run(acc, a0, 0u, a2, 0u, b0, b1);
    }
    // release: refill the NV slots with the units RING8 ahead
    __syncwarp();
    if (lane < NV) {
      int si = use_i + lane;
      if (si >= RING8) si -= RING8;
      issue_at(si, lane);
    }
    advance_cursor(NV);
    use_i += NV;
    if (use_i >= RING8) { use_i -= RING8; use_par ^= 1u; }
  };
  if (lane == 0) {
    for (int i = 0; i < RING8; ++i) mbar_init(&sh.wbar[warp][i], 1);
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&sh.xbar[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < RING8; ++i) issue_at(i, i);
  }
  advance_cursor(RING8);
  __syncwarp();
  unsigned parA = 0, parT = 0, parH = 0, parL = 0;
  if (tid == 0) {
    mbar_expect_tx(&sh.xbar[0], (unsigned)na * D * 4u);            // first fill of inA: y1 of layer 0
    mbar_expect_tx(&sh.xbar[1], (unsigned)na * D * 2u);            // att of layer 0
    mbar_expect_tx(&sh.xbar[2], (unsigned)na * F * 2u);            // h of layer 0
    if ((int)rank < NB8 && ((livemask >> rank) & 1u)) mbar_expect_tx(&sh.xbar[3], (unsigned)V * 4u);
  }
  // layer-0 inputs of the first step: xin left by prefill / the previous launch (fp32)
  for (int i = tid; i < NB8 * D; i += NT) {
    const int n = i / D, k = i - n * D;
    if ((livemask >> n) & 1u) inA[i] = ld_cg(p.xin + (size_t)sh.slot[n] * D + k);
  }
  __syncthreads();
  cluster_sync_all();

  // elements of a D-vector this lane handles in the LayerNorm warps: [c*256 + lane*8, +8) for c < NCH
  const int own_c = ((int)rank * GSV_HEAD_DIM) >> 8, own_l0 = (((int)rank * GSV_HEAD_DIM) & 255) >> 3;   // where this CTA's 32 rows sit
  uint4 gv[NCH], bv[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) { gv[c] = make_uint4(0, 0, 0, 0); bv[c] = gv[c]; }

  // LayerNorm (or plain copy) of sequence `warp` from inA into the staged operand xa, and its residual rows
  auto stage_ln = [&](bool do_ln, float* res /* [8][32] */) {
    if (warp < NB8 && ((livemask >> warp) & 1u)) {
      const int n = warp;
      float xv[NCH * 8];
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const float4 lo = *reinterpret_cast<const float4*>(inA + n * D + c * 256 + lane * 8);
        const float4 hi = *reinterpret_cast<const float4*>(inA + n * D + c * 256 + lane * 8 + 4);
        xv[c * 8 + 0] = lo.x; xv[c * 8 + 1] = lo.y; xv[c * 8 + 2] = lo.z; xv[c * 8 + 3] = lo.w;
        xv[c * 8 + 4] = hi.x; xv[c * 8 + 5] = hi.y; xv[c * 8 + 6] = hi.z; xv[c * 8 + 7] = hi.w;
      }
      if (do_ln) {
        float mean, rstd;
        ln_stats<NCH>(xv, mean, rstd);
        ln_apply<T, NCH>(xv, mean, rstd, gv, bv);
      }
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        *reinterpret_cast<uint4*>(xa + n * LDX + c * 256 + lane * 8) = pack8<T>(&xv[c * 8]);
        if (c == own_c && lane >= own_l0 && lane < own_l0 + 4) {
#pragma unroll
          for (int j = 0; j < 8; ++j) res[n * GSV_HEAD_DIM + (lane - own_l0) * 8 + j] = xv[c * 8 + j];
        }
      }
    }
  };

#pragma unroll 1
  for (int step = 0; step < n_steps; ++step) {
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      const int ln = l + 1 == L ? 0 : l + 1;
      if (warp == NWARP - 1 && lane == 0 && L > 1) {
        l2_prefetch(G2 + (size_t)l * D, D * (unsigned)sizeof(T));
        l2_prefetch(Be2 + (size_t)l * D, D * (unsigned)sizeof(T));
        l2_prefetch(G1 + (size_t)ln * D, D * (unsigned)sizeof(T));
        l2_prefetch(Be1 + (size_t)ln * D, D * (unsigned)sizeof(T));
        l2_prefetch(Bqkv + (size_t)ln * 3 * D, 3 * D * (unsigned)sizeof(T));
        l2_prefetch(Bo + (size_t)ln * D, D * (unsigned)sizeof(T));
        l2_prefetch(B1 + (size_t)ln * F, F * (unsigned)sizeof(T));
        l2_prefetch(B2 + (size_t)ln * D, D * (unsigned)sizeof(T));
      }
      // ================= A: x = l == 0 ? xin : LN2(y2); q,k,v of this head for every sequence; attention -> att ==========
      mark(p, 1);
      {
        if (tid < NB8 && ((livemask >> tid) & 1u) && l + 1 < L && sh.kv[tid] > 0) {
          const size_t hb2 = ((size_t)((l + 1) * p.slots + sh.slot[tid]) * H + rank) * (size_t)S * GSV_HEAD_DIM;
          const unsigned bytes = (unsigned)sh.kv[tid] * GSV_HEAD_DIM * (unsigned)sizeof(T);
          l2_prefetch(reinterpret_cast<const T*>(p.kc) + hb2, bytes);
          l2_prefetch(reinterpret_cast<const T*>(p.vc) + hb2, bytes);
        }
        // biases of the rows this thread finishes: batch 0 -> it = g (g < 4), batch 1 -> it = 4 + g (g < 2)
        float bq0 = 0.f, bq1 = 0.f;
        {
          const int rr0 = warp + NWARP * (g < 4 ? g : 0), rr1 = warp + NWARP * (4 + (g < 2 ? g : 0));
          bq0 = Elem<T>::to_f(Bqkv[(size_t)l * 3 * D + (rr0 >> 5) * D + rank * GSV_HEAD_DIM + (rr0 & 31)]);
          bq1 = Elem<T>::to_f(Bqkv[(size_t)l * 3 * D + (rr1 >> 5) * D + rank * GSV_HEAD_DIM + (rr1 & 31)]);
        }
        if (l > 0) { mbar_wait(&sh.xbar[0], parA); parA ^= 1u; }      // y2 of the previous layer has arrived
        stage_ln(l > 0, &sh.xres[0][0]);
        __syncthreads();                                // xa staged by the LayerNorm warps; inA fully read
        if (l > 0 && tid == 0) mbar_expect_tx(&sh.xbar[0], (unsigned)na * D * 4u);     // re-arm inA for this layer's y1
        // q, k, v rows: batch 0 = rows it 0..3, batch 1 = rows it 4..5 of this warp
#pragma unroll
        for (int bi = 0; bi < 2; ++bi) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          if (bi == 0) mma_batch(std::integral_constant<int, 4>{}, xa, LDX, 0, acc);
          else mma_batch(std::integral_constant<int, 2>{}, xa, LDX, 0, acc);
          const int nv = bi == 0 ? 4 : 2;
          if (g < nv) {
            const int rr = warp + NWARP * (bi * 4 + g), which = rr >> 5, c = rr & 31;
            const float bias = bi == 0 ? bq0 : bq1;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int n = 2 * t + j;
              if (!((livemask >> n) & 1u)) continue;
              const float v = acc[j] + bias;
              if (which == 0) {
                sh.qs[n][c] = v * (rsqrtf((float)GSV_HEAD_DIM) * 1.4426950408889634f);
              } else {
                const T t16 = Elem<T>::from_f(v);       // the reference attends over the 16-bit cache entry it has just written
                (which == 1 ? sh.kn : sh.vn)[n][c] = Elem<T>::to_f(t16);
                T* cache = reinterpret_cast<T*>(which == 1 ? p.kc : p.vc);
                const size_t hbase = ((size_t)(l * p.slots + sh.slot[n]) * H + rank) * (size_t)S * GSV_HEAD_DIM;
                cache[hbase + (size_t)sh.kv[n] * GSV_HEAD_DIM + c] = t16;
              }
            }
          }
        }
        __syncthreads();                                // q / k / v of every sequence staged
        mark(p, 50);
        // ---- attention: warps 2n and 2n+1 share the cached positions of sequence n
        {
          const int n = warp >> 1, half = warp & 1;
          const bool on = (livemask >> n) & 1u;
          const int kvn = on ? sh.kv[n] : 0;
          const int hl = (((kvn + 1) >> 1) + 7) & ~7;
          const int pb = half * hl, pe = min(kvn, pb + hl);
          const int sub = lane & 3, pg = lane >> 2;
          float mg = GSV_NEG_INF, lsum = 0.f, o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = 0.f;
          if (on) {
            const size_t head_base = ((size_t)(l * p.slots + sh.slot[n]) * H + rank) * (size_t)S * GSV_HEAD_DIM;
            const T* kb = reinterpret_cast<const T*>(p.kc) + head_base + sub * 8;
            const T* vb = reinterpret_cast<const T*>(p.vc) + head_base + sub * 8;
            float q[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) q[j] = sh.qs[n][sub * 8 + j];
            uint4 krn = make_uint4(0, 0, 0, 0), vrn = krn;
            if (pb + pg < pe) {
              krn = ld_cg16(kb + (size_t)(pb + pg) * GSV_HEAD_DIM);
              vrn = ld_cg16(vb + (size_t)(pb + pg) * GSV_HEAD_DIM);
            }
#pragma unroll 1
            for (int base = pb; base < pe; base += 8) {
              const int pos = base + pg;
              const bool ok = pos < pe;
              const uint4 kr = krn, vr = vrn;
              if (pos + 8 < pe) {
                krn = ld_cg16(kb + (size_t)(pos + 8) * GSV_HEAD_DIM);
                vrn = ld_cg16(vb + (size_t)(pos + 8) * GSV_HEAD_DIM);
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
          }
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
            if (sub == 0) { sh.apart[warp][0] = m; sh.apart[warp][1] = lsum; }
#pragma unroll
            for (int j = 0; j < 8; ++j) sh.apart[warp][2 + sub * 8 + j] = o[j];
          }
          __syncthreads();
          if (on && half == 0) {
            // merge the two halves and the new position; lane = output dimension
            const float m0 = sh.apart[warp][0], m1 = sh.apart[warp + 1][0];
            const float snew = warp_allsum(sh.qs[n][lane] * sh.kn[n][lane]);
            const float M = fmaxf(fmaxf(m0, m1), snew);
            const float s0 = m0 > GSV_NEG_INF ? exp2f(m0 - M) : 0.f, s1 = m1 > GSV_NEG_INF ? exp2f(m1 - M) : 0.f;
            const float pr = exp2f(snew - M);
            const float Ls = sh.apart[warp][1] * s0 + sh.apart[warp + 1][1] * s1 + pr;
            const float oa = sh.apart[warp][2 + lane] * s0 + sh.apart[warp + 1][2 + lane] * s1 + pr * sh.vn[n][lane];
            // stage the 32 outputs of (sequence n, this head) in the storage type, then 16 lanes push 64 bytes each
            T* stg = reinterpret_cast<T*>(&sh.stage[warp][0][0]);
            stg[lane] = Elem<T>::from_f(oa / Ls);
            __syncwarp();
            if (lane < H) {
              const uint4* s4 = reinterpret_cast<const uint4*>(stg);
              T* dst = attb + n * LDX + (int)rank * GSV_HEAD_DIM;
#pragma unroll
              for (int i = 0; i < 4; ++i) st_async_v4(dst + i * 8, &sh.xbar[1], (unsigned)lane, s4[i]);
            }
          }
        }
      }
      mark(p, 52);
      mbar_wait(&sh.xbar[1], parT); parT ^= 1u;          // att of every head and sequence has arrived
      mark(p, 3);
      // ================= O: y1 = x + att Wo^T + bo ==================
      {
        load_vec<T, NCH>(G1 + (size_t)l * D, lane, gv);
        load_vec<T, NCH>(Be1 + (size_t)l * D, lane, bv);
        const int o_loc = warp * 2 + (g < 2 ? g : 0);   // row of this CTA's 32 that thread (g < 2) finishes
        const float o_bias = Elem<T>::to_f(Bo[(size_t)l * D + rank * GSV_HEAD_DIM + o_loc]);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        mma_batch(std::integral_constant<int, 2>{}, attb, LDX, 0, acc);
        __syncthreads();                                // every warp has read att: re-arm its inbox for the next layer
        if (tid == 0 && l + 1 < L) mbar_expect_tx(&sh.xbar[1], (unsigned)na * D * 2u);
        if (g < 2) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int n = 2 * t + j;
            sh.stage[warp][n][g] = acc[j] + o_bias + sh.xres[n][o_loc];
          }
        }
        __syncwarp();
        {
          // lane -> target CTA lane & 15, sequences (lane >> 4) * 4 .. + 3; 8 bytes (rows 2w, 2w+1) per (target, sequence)
          const int tgt = lane & 15;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int n = (lane >> 4) * 4 + i;
            if (tgt < H && ((livemask >> n) & 1u))
              st_async_v2(inA + n * D + (int)rank * GSV_HEAD_DIM + warp * 2, &sh.xbar[0], (unsigned)tgt,
                          __float_as_uint(sh.stage[warp][n][0]), __float_as_uint(sh.stage[warp][n][1]));
          }
        }
        __syncwarp();
      }
      mark(p, 53);
      mbar_wait(&sh.xbar[0], parA); parA ^= 1u;          // y1 has arrived
      mark(p, 4);
      // ================= M1: x1 = LN1(y1); h = relu(x1 W1^T + b1) ==================
      {
        const int r_lo = warp * 8 + (g < 4 ? g : 0), r_hi = r_lo + 4;       // rows of this CTA's 128 that thread (g < 4) finishes
        const float b_lo = Elem<T>::to_f(B1[(size_t)l * F + rank * (4 * GSV_HEAD_DIM) + r_lo]);
        const float b_hi = Elem<T>::to_f(B1[(size_t)l * F + rank * (4 * GSV_HEAD_DIM) + r_hi]);
        stage_ln(true, &sh.xres1[0][0]);
        __syncthreads();                                // xa staged; inA fully read
        if (tid == 0) mbar_expect_tx(&sh.xbar[0], (unsigned)na * D * 4u);             // re-arm inA for y2
        float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
        mma_batch(std::integral_constant<int, 4>{}, xa, LDX, 0, acc0);
        mma_batch(std::integral_constant<int, 4>{}, xa, LDX, 0, acc1);
        // stage h of rows 8w .. 8w+7 for every sequence in the storage type: 16 bytes per sequence
        if (g < 4) {
          T* stg = reinterpret_cast<T*>(&sh.stage[warp][0][0]);               // [8 sequences][8 rows] T
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int n = 2 * t + j;
            stg[n * 8 + g] = Elem<T>::from_f(fmaxf(acc0[j] + b_lo, 0.f));
            stg[n * 8 + 4 + g] = Elem<T>::from_f(fmaxf(acc1[j] + b_hi, 0.f));
          }
        }
        __syncwarp();
        {
          const int tgt = lane & 15;
          const uint4* s4 = reinterpret_cast<const uint4*>(&sh.stage[warp][0][0]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int n = (lane >> 4) * 4 + i;
            if (tgt < H && ((livemask >> n) & 1u))
              st_async_v4(hb + n * LDH + (int)rank * (4 * GSV_HEAD_DIM) + warp * 8, &sh.xbar[2], (unsigned)tgt, s4[n]);
          }
        }
        __syncwarp();
      }
      mark(p, 54);
      mbar_wait(&sh.xbar[2], parH); parH ^= 1u;          // h has arrived
      mark(p, 5);
      // ================= M2: y2 = x1 + h W2^T + b2 ==================
      {
        load_vec<T, NCH>(G2 + (size_t)l * D, lane, gv);
        load_vec<T, NCH>(Be2 + (size_t)l * D, lane, bv);
        const int m_loc = warp * 2 + (g < 2 ? g : 0);
        const float m_bias = Elem<T>::to_f(B2[(size_t)l * D + rank * GSV_HEAD_DIM + m_loc]);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) mma_batch(std::integral_constant<int, 2>{}, hb, LDH, q4 * D, acc);
        __syncthreads();                                // every warp has read h: re-arm its inbox for the next layer
        if (tid == 0 && l + 1 < L) mbar_expect_tx(&sh.xbar[2], (unsigned)na * F * 2u);
        if (g < 2) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int n = 2 * t + j;
            sh.stage[warp][n][g] = acc[j] + m_bias + sh.xres1[n][m_loc];
          }
        }
        __syncwarp();
        {
          const int tgt = lane & 15;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int n = (lane >> 4) * 4 + i;
            if (tgt < H && ((livemask >> n) & 1u))
              st_async_v2(inA + n * D + (int)rank * GSV_HEAD_DIM + warp * 2, &sh.xbar[0], (unsigned)tgt,
                          __float_as_uint(sh.stage[warp][n][0]), __float_as_uint(sh.stage[warp][n][1]));
          }
        }
        __syncwarp();
      }
      mark(p, 55);
    }
    mark(p, 6);
    // ================= head: logits of sequence n = LN2_last(y2_n) Whead^T, pushed to CTA n ==================
    {
      mbar_wait(&sh.xbar[0], parA); parA ^= 1u;          // y2 of the last layer has arrived
      stage_ln(true, &sh.xres[0][0]);                    // (residual copy unused here)
      __syncthreads();
      if (tid == 0) mbar_expect_tx(&sh.xbar[0], (unsigned)na * (D * 4u + 4u));        // re-arm inA for the next inputs (+ alive flags)
      uint4 w[NCH], wn[NCH];
      int j = warp;
      if ((int)rank + j * H < V) load_vec<T, NCH>(Wh + (size_t)((int)rank + j * H) * D, lane, w);
#pragma unroll 1
      for (; (int)rank + j * H < V; j += NWARP) {
        const int gg = (int)rank + j * H, gn = gg + NWARP * H;
        if (gn < V) load_vec<T, NCH>(Wh + (size_t)gn * D, lane, wn);
#pragma unroll 1
        for (int n = 0; n < NB8; ++n) {
          if (!((livemask >> n) & 1u)) continue;
          float xv[NCH * 8];
#pragma unroll
          for (int c = 0; c < NCH; ++c) unpack8<T>(*reinterpret_cast<const uint4*>(xa + n * LDX + c * 256 + lane * 8), &xv[c * 8]);
          const float a = warp_allsum(dot_regs<T, NCH>(w, xv));
          if (lane == 0) st_async(samp + gg, &sh.xbar[3], (unsigned)n, a);
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) w[c] = wn[c];
      }
    }
    mark(p, 20);
    // ================= sampling: CTA n samples sequence n; next inputs and alive flags pushed to every CTA ==================
    if ((int)rank < NB8 && ((livemask >> rank) & 1u)) {
      const int n = (int)rank;
      mbar_wait(&sh.xbar[3], parL); parL ^= 1u;          // all V logits of sequence n have arrived
      SampleLL io;
      io.preloaded = true;
      io.xin_ll = nullptr;
      io.status_ll = nullptr;
      io.tag = 0;
      io.kv_len = sh.kv[n] + 1;
      io.xin_smem = xin_s;
      io.alive_smem = &sh.alive_i;
      sample_slot<T>(p, sh.slot[n], samp, &io);
      __syncthreads();
      const bool still = sh.alive_i != 0;
      if (tid == 0 && still) mbar_expect_tx(&sh.xbar[3], (unsigned)V * 4u);           // re-armed before anyone can refill it
      for (int i = tid; i < H * (D / 4); i += NT) {
        const int tgt = i / (D / 4), k4 = i - tgt * (D / 4);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (still) v = *reinterpret_cast<const uint4*>(xin_s + k4 * 4);
        st_async_v4(inA + n * D + k4 * 4, &sh.xbar[0], (unsigned)tgt, v);
      }
      if (tid < H) st_async(&sh.alive[n], &sh.xbar[0], (unsigned)tid, still ? 1.f : 0.f);
    }
    mbar_wait(&sh.xbar[0], parA); parA ^= 1u;            // next inputs + alive flags of every live sequence have arrived
    {
      unsigned nm = 0;
#pragma unroll
      for (int n = 0; n < NB8; ++n)
        if (((livemask >> n) & 1u) && sh.alive[n] != 0.f) nm |= 1u << n;
      __syncthreads();                                  // everyone has read the flags and kv
      if (tid < NB8 && ((livemask >> tid) & 1u)) sh.kv[tid] += 1;
      livemask = nm;
      na = __popc(livemask);
    }
    __syncthreads();
    if (tid == 0 && na > 0) {
      mbar_expect_tx(&sh.xbar[0], (unsigned)na * D * 4u);          // inA: y1 of the next token's layer 0
      mbar_expect_tx(&sh.xbar[1], (unsigned)na * D * 2u);          // att (armed here because the live count may have changed)
      mbar_expect_tx(&sh.xbar[2], (unsigned)na * F * 2u);          // h
    }
    mark(p, 21);
    if (na == 0) break;
  }
  for (int i = 0; i < RING8; ++i) {
    mbar_wait(&sh.wbar[warp][use_i], use_par);
    if (++use_i == RING8) { use_i = 0; use_par ^= 1u; }
  }
  cluster_sync_all();
}

template <typename T>
int launch_cl8(gsv_gpt_ctx* ctx, int live, int n_steps, cudaStream_t st) {
  const int nd = ctx->p.d / 256;
  void* fn = nullptr;
  if (nd == 2) fn = (void*)gpt_decode_cl8_kernel<T, 2>;
  else if (nd == 1) fn = (void*)gpt_decode_cl8_kernel<T, 1>;
  else return GSV_ERR_ARG;
  const int D = ctx->p.d, F = ctx->p.F, H = ctx->p.H;
  const size_t bytes = (size_t)NB8 * D * 4 + (size_t)2 * NB8 * (D + XPAD) * 2 + (size_t)NB8 * (F + XPAD) * 2 + (size_t)D * 4 +
                       (size_t)((GSV_SAMPLE_SMEM_FLOATS + 3) & ~3) * 4 + (size_t)NWARP * RING8 * (D * 2 + 16);
  GSV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  GSV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  const int n_clusters = (live + NB8 - 1) / NB8;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(n_clusters * H); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = bytes; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = H; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  GptParams p = ctx->p;
  int ns = n_steps;
  void* args[] = {&p, &ns};
  GSV_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  ctx->launches += 1;
  return GSV_OK;
}

}  // namespace

int gsv_gpt_decode_cl8_launch(gsv_gpt_ctx* ctx, int live_slots, int n_steps, cudaStream_t st) {
  if (ctx->dims.dtype == GSV_F16) return launch_cl8<__half>(ctx, live_slots, n_steps, st);
  return launch_cl8<__nv_bfloat16>(ctx, live_slots, n_steps, st);
}
