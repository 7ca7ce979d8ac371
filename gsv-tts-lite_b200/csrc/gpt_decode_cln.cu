// gpt_decode_cln.cu -- cluster decode kernel for SEVERAL sequences per cluster (NB = 2 or 4).
//
// Same structure and arithmetic as gpt_decode_cl.cu (one CTA per attention head, st.async inboxes, per-warp
// bulk-copy weight ring; reference t2s_model.py:67-105, 129-143, 442-456).  What changes: a cluster serves NB live
// sequences, so every weight row that is streamed from HBM is multiplied with NB input vectors.  At most 7
// sixteen-CTA clusters are co-resident on a B200 (one per GPC) and each GPC's memory port sustains ~0.8 TB/s, so one
// sequence per cluster tops out at ~7 sequences x 2.8 k tok/s; sharing the weight stream is what lifts the
// aggregate beyond that.  Sequence s of the cluster is sampled by CTA s, all sequences advance in lock step, and a
// sequence that finishes simply drops out of the per-phase loops (inbox byte counts follow the live count).
#include "gpt_cluster_common.cuh"

namespace {

constexpr int MAXNB = 4;
template <int NB> struct RingOf { static constexpr int v = NB >= 4 ? 9 : 10; };   // shared memory: NB = 4 leaves room for 9 units per warp

struct ClnShared {
  float q[MAXNB][GSV_HEAD_DIM], knew[MAXNB][GSV_HEAD_DIM], vnew[MAXNB][GSV_HEAD_DIM];
  float wpart[NWARP][GSV_HEAD_DIM + 2];
  float alive[MAXNB];      // pushed by the sampler CTAs together with the next inputs (inbox A)
  int alive_i;             // sampler CTA: written by sample_slot
  int slot[MAXNB], kv[MAXNB];
  uint64_t wbar[NWARP][RING];   // weight ring: one mbarrier per warp and slot
  uint64_t xbar[3];             // inboxes: 0 bufA (xin / y1 / y2), 1 bufB (att / h), 2 logits (CTA s for sequence s)
};

template <typename T, int NCH, int NB>
__global__ void __launch_bounds__(NT, 1) gpt_decode_cln_kernel(const GptParams p, const int n_steps) {
  extern __shared__ __align__(16) float smem[];
  __shared__ ClnShared sh;
  constexpr int D = NCH * 256, F = 4 * D;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, L = p.L, V = p.V, S = p.S;
  const unsigned rank = cluster_rank();               // = head index
  const int cid = blockIdx.x / H;                      // cluster index: serves live sequences [cid*NB, cid*NB + NB)

  // shared memory: bufA[NB][D] (xin / y1 / y2) | bufB[NB][F] (att / h) inboxes | xn[NB][32] | x1n[NB][32] residual rows of this CTA |
  // hx[NB][D] normalised head input | xin_s[D] | sampler scratch (+ logits) | per-warp weight ring
  constexpr int RINGN = RingOf<NB>::v;
  float* bufA = smem;
  float* bufB = bufA + NB * D;
  float* xn = bufB + NB * F;
  float* x1n = xn + NB * GSV_HEAD_DIM;
  float* xin_s = x1n + NB * GSV_HEAD_DIM;
  float* samp = xin_s + D;
  uint4* ring = reinterpret_cast<uint4*>(samp + ((GSV_SAMPLE_SMEM_FLOATS + 3) & ~3)) + (size_t)warp * RINGN * (D / 8);

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

  // ---- which sequences: the active slots number cid*NB .. cid*NB + NB - 1 ----
  if (tid < 32) {
    const int flag = tid < p.slots ? ld_cg(p.active + tid) : 0;
    const unsigned m = __ballot_sync(0xffffffffu, flag != 0);
    const int pos = flag ? __popc(m & ((1u << tid) - 1u)) - cid * NB : -1;
    if (tid < MAXNB) { sh.slot[tid] = -1; sh.kv[tid] = 0; sh.alive[tid] = 0.f; }
    __syncwarp();
    if (flag && pos >= 0 && pos < NB) { sh.slot[pos] = tid; sh.kv[pos] = ld_cg(p.kv_len + tid); sh.alive[pos] = 1.f; }
  }
  __syncthreads();
  int slot[NB], kv[NB];
  bool live[NB];
  int na = 0;                                           // live sequences of this cluster (uniform across its CTAs)
#pragma unroll
  for (int s = 0; s < NB; ++s) { slot[s] = sh.slot[s]; kv[s] = sh.kv[s]; live[s] = slot[s] >= 0; na += live[s] ? 1 : 0; }
  if (na == 0) return;

  // ---- weight unit sequence of this warp (as in gpt_decode_cl.cu) ----
  auto unit_src = [&](int l, int u) -> const T* {
    if (u < QKV_PER_WARP) {
      const int rr = warp + NWARP * u;
      const int row = (rr >> 5) * D + (int)rank * GSV_HEAD_DIM + (rr & 31);
      return Wqkv + ((size_t)l * 3 * D + row) * D;
    }
    u -= QKV_PER_WARP;
    if (u < O_PER_WARP) return Wo + ((size_t)l * D + rank * GSV_HEAD_DIM + warp * O_PER_WARP + u) * D;
    u -= O_PER_WARP;
    if (u < M1_PER_WARP) return W1 + ((size_t)l * F + rank * (4 * GSV_HEAD_DIM) + warp * M1_PER_WARP + u) * D;
    u -= M1_PER_WARP;
    return W2 + ((size_t)l * D + rank * GSV_HEAD_DIM + warp * M2_PER_WARP + (u >> 2)) * F + (size_t)(u & 3) * D;
  };
  int iss_l = 0, iss_u = 0, use_i = 0;
  unsigned use_par = 0;
  constexpr unsigned UNIT_BYTES = D * (unsigned)sizeof(T);
  auto issue_at = [&](int slot_i, int ahead) {
    int u = iss_u + ahead, l = iss_l;
    if (u >= UNITS_PER_LAYER) { u -= UNITS_PER_LAYER; l = l + 1 == L ? 0 : l + 1; }
    mbar_expect_tx(&sh.wbar[warp][slot_i], UNIT_BYTES);
    bulk_g2s(ring + (size_t)slot_i * (D / 8), unit_src(l, u), UNIT_BYTES, &sh.wbar[warp][slot_i]);
  };
  auto advance_cursor = [&](int n) {
    iss_u += n;
    if (iss_u >= UNITS_PER_LAYER) { iss_u -= UNITS_PER_LAYER; iss_l = iss_l + 1 == L ? 0 : iss_l + 1; }
  };
  auto take_n = [&](auto& w) {
    constexpr int N = (int)(sizeof(w) / sizeof(w[0]));
#pragma unroll
    for (int i = 0; i < N; ++i) {
      int si = use_i + i;
      unsigned par = use_par;
      if (si >= RINGN) { si -= RINGN; par ^= 1u; }
      mbar_wait(&sh.wbar[warp][si], par);
      const uint4* srcs = ring + (size_t)si * (D / 8);
#pragma unroll
      for (int c = 0; c < NCH; ++c) w[i][c] = srcs[c * 32 + lane];
    }
  };
  auto release_n = [&](int n) {
    __syncwarp();
    if (lane < n) {
      int si = use_i + lane;
      if (si >= RINGN) si -= RINGN;
      issue_at(si, lane);
    }
    advance_cursor(n);
    use_i += n;
    if (use_i >= RINGN) { use_i -= RINGN; use_par ^= 1u; }
  };
  if (lane == 0) {
    for (int i = 0; i < RINGN; ++i) mbar_init(&sh.wbar[warp][i], 1);
    if (tid == 0) { mbar_init(&sh.xbar[0], 1); mbar_init(&sh.xbar[1], 1); mbar_init(&sh.xbar[2], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < RINGN; ++i) issue_at(i, i);
  }
  advance_cursor(RINGN);
  __syncwarp();
  unsigned parA = 0, parB = 0, parL = 0;
  if (tid == 0) {
    mbar_expect_tx(&sh.xbar[0], (unsigned)na * D * 4u);            // first fill of inbox A: y1 of layer 0
    mbar_expect_tx(&sh.xbar[1], (unsigned)na * D * 4u);            // first fill of inbox B: att of layer 0
#pragma unroll
    for (int s = 0; s < NB; ++s)
      if ((int)rank == s && live[s]) mbar_expect_tx(&sh.xbar[2], (unsigned)V * 4u);
  }
  // layer-0 inputs of the first step: xin left by prefill / the previous launch (fp32)
#pragma unroll
  for (int s = 0; s < NB; ++s)
    if (live[s])
      for (int k = tid; k < D; k += NT) bufA[s * D + split_pos(k, D)] = ld_cg(p.xin + (size_t)slot[s] * D + k);
  __syncthreads();
  cluster_sync_all();

  const int sub = lane & 3, pg = lane >> 2;
  uint4 gv[NCH], bv[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) { gv[c] = make_uint4(0, 0, 0, 0); bv[c] = gv[c]; }
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
      // ================= A: inputs (bufA), q,k,v of this head for every sequence, attention -> att (bufB) ==============
      mark(p, 1);
      {
        if (tid == 0 && l + 1 < L) {
#pragma unroll
          for (int s = 0; s < NB; ++s)
            if (live[s] && kv[s] > 0) {
              const size_t hb = ((size_t)((l + 1) * p.slots + slot[s]) * H + rank) * (size_t)S * GSV_HEAD_DIM;
              const unsigned bytes = (unsigned)kv[s] * GSV_HEAD_DIM * (unsigned)sizeof(T);
              l2_prefetch(reinterpret_cast<const T*>(p.kc) + hb, bytes);
              l2_prefetch(reinterpret_cast<const T*>(p.vc) + hb, bytes);
            }
        }
        float bq = 0.f;
        if (lane < QKV_PER_WARP) {
          const int rr = warp + NWARP * lane;
          bq = Elem<T>::to_f(Bqkv[(size_t)l * 3 * D + (rr >> 5) * D + rank * GSV_HEAD_DIM + (rr & 31)]);
        }
        if (l > 0) { mbar_wait(&sh.xbar[0], parA); parA ^= 1u; }      // y2 of the previous layer has arrived (all sequences)
        uint4 w[QKV_PER_WARP][NCH];
        take_n(w);
        const int it = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        const float bias = __shfl_sync(0xffffffffu, bq, it);
#pragma unroll
        for (int s = 0; s < NB; ++s) {
          if (!live[s]) continue;
          float xv[NCH * 8];
          load_x<NCH>(bufA + s * D, lane, xv);
          {
            float mean = 0.f, rstd = 1.f;
            if (l > 0) ln_stats<NCH>(xv, mean, rstd);
            if (warp == 0) {                            // residual rows of this CTA (out-proj adds them): row rank*32 + lane
              const int r = (int)rank * GSV_HEAD_DIM + lane;
              const float raw = bufA[s * D + split_pos(r, D)];
              xn[s * GSV_HEAD_DIM + lane] = l > 0 ? fmaf((raw - mean) * rstd, Elem<T>::to_f(G2[(size_t)(l - 1) * D + r]),
                                                         Elem<T>::to_f(Be2[(size_t)(l - 1) * D + r])) : raw;
            }
            if (l > 0) ln_apply<T, NCH>(xv, mean, rstd, gv, bv);
          }
          float part[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) part[i] = i < QKV_PER_WARP ? dot_regs<T, NCH>(w[i < QKV_PER_WARP ? i : 0], xv) : 0.f;
          const float sum = reduce8(part, lane);
          if ((lane & 3) == 0 && it < QKV_PER_WARP) {
            const int rr = warp + NWARP * it, which = rr >> 5, c = rr & 31;
            const float v = sum + bias;
            if (which == 0) {
              sh.q[s][c] = v * (rsqrtf((float)GSV_HEAD_DIM) * 1.4426950408889634f);
            } else {
              const T t16 = Elem<T>::from_f(v);
              (which == 1 ? sh.knew : sh.vnew)[s][c] = Elem<T>::to_f(t16);
              T* cache = reinterpret_cast<T*>(which == 1 ? p.kc : p.vc);
              const size_t hb = ((size_t)(l * p.slots + slot[s]) * H + rank) * (size_t)S * GSV_HEAD_DIM;
              cache[hb + (size_t)kv[s] * GSV_HEAD_DIM + c] = t16;
            }
          }
        }
        release_n(QKV_PER_WARP);
        __syncthreads();                                // q/k/v of every sequence staged; every warp has read inbox A
        if (l > 0 && tid == 0) mbar_expect_tx(&sh.xbar[0], (unsigned)na * D * 4u);      // re-arm inbox A for this layer's y1
        mark(p, 50);
#pragma unroll 1
        for (int s = 0; s < NB; ++s) {
          if (!live[s]) continue;
          const size_t head_base = ((size_t)(l * p.slots + slot[s]) * H + rank) * (size_t)S * GSV_HEAD_DIM;
          const T* kb = reinterpret_cast<const T*>(p.kc) + head_base + sub * 8;
          const T* vb = reinterpret_cast<const T*>(p.vc) + head_base + sub * 8;
          const int kvs = kv[s];
          float q[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) q[j] = sh.q[s][sub * 8 + j];
          float mg = GSV_NEG_INF, lsum = 0.f, o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = 0.f;
          uint4 krn = make_uint4(0, 0, 0, 0), vrn = krn;
          if (warp * 8 + pg < kvs) {
            krn = ld_cg16(kb + (size_t)(warp * 8 + pg) * GSV_HEAD_DIM);
            vrn = ld_cg16(vb + (size_t)(warp * 8 + pg) * GSV_HEAD_DIM);
          }
#pragma unroll 1
          for (int base = warp * 8; base < kvs; base += NWARP * 8) {
            const int pos = base + pg;
            const bool ok = pos < kvs;
            const uint4 kr = krn, vr = vrn;
            if (pos + NWARP * 8 < kvs) {                 // next pass in flight while this one is reduced
              krn = ld_cg16(kb + (size_t)(pos + NWARP * 8) * GSV_HEAD_DIM);
              vrn = ld_cg16(vb + (size_t)(pos + NWARP * 8) * GSV_HEAD_DIM);
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
          __syncthreads();                              // the previous sequence's partials have been merged
          if (lane < 4) {
            if (sub == 0) { sh.wpart[warp][0] = m; sh.wpart[warp][1] = lsum; }
#pragma unroll
            for (int j = 0; j < 8; ++j) sh.wpart[warp][2 + sub * 8 + j] = o[j];
          }
          __syncthreads();
          {
            const float mw = lane < NWARP ? sh.wpart[lane][0] : GSV_NEG_INF;
            const float snew = warp_allsum(sh.q[s][lane] * sh.knew[s][lane]);
            const float M = fmaxf(warp_max(mw), snew);
            float Ls = 0.f, oa = 0.f;
#pragma unroll
            for (int w2 = 0; w2 < NWARP; ++w2) {
              const float mwv = sh.wpart[w2][0];
              const float sc = mwv > GSV_NEG_INF ? exp2f(mwv - M) : 0.f;
              Ls = fmaf(sh.wpart[w2][1], sc, Ls);
              oa = fmaf(sh.wpart[w2][2 + lane], sc, oa);
            }
            const float pr = exp2f(snew - M);
            Ls += pr;
            oa = fmaf(pr, sh.vnew[s][lane], oa);
            if (warp < H) st_async(bufB + s * F + split_pos((int)rank * GSV_HEAD_DIM + lane, D), &sh.xbar[1], (unsigned)warp, oa / Ls);
          }
        }
      }
      mark(p, 52);
      mbar_wait(&sh.xbar[1], parB); parB ^= 1u;          // att of every head and sequence has arrived
      mark(p, 3);
      // ================= O: y1 = x + att Wo^T + bo (reads bufB, writes bufA) ==================
      {
        load_vec<T, NCH>(G1 + (size_t)l * D, lane, gv);
        load_vec<T, NCH>(Be1 + (size_t)l * D, lane, bv);
        const int o_row = (int)rank * GSV_HEAD_DIM + warp * O_PER_WARP + (lane >> 4);
        const float o_bias = Elem<T>::to_f(Bo[(size_t)l * D + o_row]);
        uint4 w[O_PER_WARP][NCH];
        take_n(w);
        float sum[NB];
#pragma unroll
        for (int s = 0; s < NB; ++s) {
          sum[s] = 0.f;
          if (!live[s]) continue;
          float xv[NCH * 8];
          load_x<NCH>(bufB + s * F, lane, xv);
          sum[s] = reduce2(dot_regs<T, NCH>(w[0], xv), dot_regs<T, NCH>(w[1], xv), lane);
        }
        release_n(O_PER_WARP);
        __syncthreads();                                // every warp has read inbox B: re-arm it for h
        if (tid == 0) mbar_expect_tx(&sh.xbar[1], (unsigned)na * F * 4u);
        const int tgt = lane & 15;
#pragma unroll
        for (int s = 0; s < NB; ++s) {
          if (!live[s]) continue;
          const float v = sum[s] + o_bias + xn[s * GSV_HEAD_DIM + (o_row - (int)rank * GSV_HEAD_DIM)];
          if (tgt < H) st_async(bufA + s * D + split_pos(o_row, D), &sh.xbar[0], (unsigned)tgt, v);
        }
      }
      mark(p, 53);
      mbar_wait(&sh.xbar[0], parA); parA ^= 1u;          // y1 has arrived
      mark(p, 4);
      // ================= M1: x1 = LN1(y1); h = relu(x1 W1^T + b1) (reads bufA, writes bufB) ==================
      {
        const int m1_i = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        const int m1_row = (int)rank * (4 * GSV_HEAD_DIM) + warp * M1_PER_WARP + m1_i;
        const float m1_bias = Elem<T>::to_f(B1[(size_t)l * F + m1_row]);
        const int qd = m1_row / D, kk = m1_row - qd * D;
        uint4 w[M1_PER_WARP][NCH];
        take_n(w);
        float hv[NB];
#pragma unroll
        for (int s = 0; s < NB; ++s) {
          hv[s] = 0.f;
          if (!live[s]) continue;
          float xv[NCH * 8];
          load_x<NCH>(bufA + s * D, lane, xv);
          float mean, rstd;
          ln_stats<NCH>(xv, mean, rstd);
          if (warp == 0) {                              // residual rows of this CTA (MLP-down adds them): row rank*32 + lane
            const int r = (int)rank * GSV_HEAD_DIM + lane;
            x1n[s * GSV_HEAD_DIM + lane] = fmaf((bufA[s * D + split_pos(r, D)] - mean) * rstd, Elem<T>::to_f(G1[(size_t)l * D + r]),
                                                Elem<T>::to_f(Be1[(size_t)l * D + r]));
          }
          ln_apply<T, NCH>(xv, mean, rstd, gv, bv);
          float part[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) part[i] = dot_regs<T, NCH>(w[i], xv);
          hv[s] = fmaxf(reduce8(part, lane) + m1_bias, 0.f);
        }
        release_n(M1_PER_WARP);
        __syncthreads();                                // re-arm inbox A for y2
        if (tid == 0) mbar_expect_tx(&sh.xbar[0], (unsigned)na * D * 4u);
#pragma unroll
        for (int s = 0; s < NB; ++s) {
          if (!live[s]) continue;
          float* dst = bufB + s * F + qd * D + split_pos(kk, D);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int tgt = (lane & 3) * 4 + j;
            if (tgt < H) st_async(dst, &sh.xbar[1], (unsigned)tgt, hv[s]);
          }
        }
      }
      mark(p, 54);
      mbar_wait(&sh.xbar[1], parB); parB ^= 1u;          // h has arrived
      mark(p, 5);
      // ================= M2: y2 = x1 + h W2^T + b2 (reads bufB, writes bufA) ==================
      {
        load_vec<T, NCH>(G2 + (size_t)l * D, lane, gv);
        load_vec<T, NCH>(Be2 + (size_t)l * D, lane, bv);
        const int m2_row = (int)rank * GSV_HEAD_DIM + warp * M2_PER_WARP + (lane >> 4);
        const float m2_bias = Elem<T>::to_f(B2[(size_t)l * D + m2_row]);
        uint4 w[4 * M2_PER_WARP][NCH];
        take_n(w);
        float sum[NB];
#pragma unroll
        for (int s = 0; s < NB; ++s) {
          sum[s] = 0.f;
          if (!live[s]) continue;
          float p0 = 0.f, p1 = 0.f;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            float xv[NCH * 8];
            load_x<NCH>(bufB + s * F + q4 * D, lane, xv);
            p0 += dot_regs<T, NCH>(w[q4], xv);
            p1 += dot_regs<T, NCH>(w[4 + q4], xv);
          }
          sum[s] = reduce2(p0, p1, lane);
        }
        release_n(4 * M2_PER_WARP);
        __syncthreads();                                // every warp has read inbox B: re-arm it for the next layer's att
        // (after the last layer the live count may change: inbox B is re-armed at the end of the step instead)
        if (tid == 0 && l + 1 < L) mbar_expect_tx(&sh.xbar[1], (unsigned)na * D * 4u);
        const int tgt = lane & 15;
#pragma unroll
        for (int s = 0; s < NB; ++s) {
          if (!live[s]) continue;
          const float v = sum[s] + m2_bias + x1n[s * GSV_HEAD_DIM + (m2_row - (int)rank * GSV_HEAD_DIM)];
          if (tgt < H) st_async(bufA + s * D + split_pos(m2_row, D), &sh.xbar[0], (unsigned)tgt, v);
        }
      }
      mark(p, 55);
    }
    mark(p, 6);
    // ================= head: logits of sequence s = LN2_last(y2_s) Whead^T, pushed to CTA s ==================
    {
      mbar_wait(&sh.xbar[0], parA); parA ^= 1u;          // y2 of the last layer has arrived
      float* hx = bufB;                                 // inbox B is idle here (re-armed at the end of the step): normalised head inputs [NB][D]
#pragma unroll
      for (int s = 0; s < NB; ++s) {
        if (!live[s]) continue;
        float xv[NCH * 8];
        load_x<NCH>(bufA + s * D, lane, xv);
        float mean, rstd;
        ln_stats<NCH>(xv, mean, rstd);
        ln_apply<T, NCH>(xv, mean, rstd, gv, bv);
        if (warp == 0) store_x<NCH>(hx + s * D, lane, xv);
      }
      __syncthreads();                                  // head inputs staged; re-arm inbox A for the next inputs (+ the alive flags)
      if (tid == 0) mbar_expect_tx(&sh.xbar[0], (unsigned)na * (D * 4u + 4u));
      uint4 w[NCH], wn[NCH];
      int j = warp;
      if ((int)rank + j * H < V) load_vec<T, NCH>(Wh + (size_t)((int)rank + j * H) * D, lane, w);
#pragma unroll 1
      for (; (int)rank + j * H < V; j += NWARP) {
        const int g = (int)rank + j * H, gn = g + NWARP * H;
        if (gn < V) load_vec<T, NCH>(Wh + (size_t)gn * D, lane, wn);
#pragma unroll
        for (int s = 0; s < NB; ++s) {
          if (!live[s]) continue;
          float xv[NCH * 8];
          load_x<NCH>(hx + s * D, lane, xv);
          const float a = warp_allsum(dot_regs<T, NCH>(w, xv));
          if (lane == 0) st_async(samp + g, &sh.xbar[2], (unsigned)s, a);
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) w[c] = wn[c];
      }
    }
    mark(p, 20);
    // ================= sampling: CTA s samples sequence s; next inputs and alive flags pushed to every CTA ==================
#pragma unroll
    for (int s = 0; s < NB; ++s) {
      if ((int)rank != s || !live[s]) continue;         // uniform per CTA
      mbar_wait(&sh.xbar[2], parL); parL ^= 1u;          // all V logits of sequence s have arrived
      SampleLL io;
      io.preloaded = true;
      io.xin_ll = nullptr;
      io.status_ll = nullptr;
      io.tag = 0;
      io.kv_len = kv[s] + 1;
      io.xin_smem = xin_s;
      io.alive_smem = &sh.alive_i;
      sample_slot<T>(p, slot[s], samp, &io);
      __syncthreads();
      const bool still = sh.alive_i != 0;
      if (tid == 0 && still) mbar_expect_tx(&sh.xbar[2], (unsigned)V * 4u);           // re-armed before anyone can refill it
      for (int i = tid; i < H * D; i += NT) {
        const int tgt = i / D, k = i - tgt * D;
        st_async(bufA + s * D + split_pos(k, D), &sh.xbar[0], (unsigned)tgt, still ? xin_s[k] : 0.f);
      }
      if (tid < H) st_async(&sh.alive[s], &sh.xbar[0], (unsigned)tid, still ? 1.f : 0.f);
    }
    mbar_wait(&sh.xbar[0], parA); parA ^= 1u;            // next inputs + alive flags of every live sequence have arrived
    int na2 = 0;
#pragma unroll
    for (int s = 0; s < NB; ++s) {
      if (live[s]) { kv[s] += 1; live[s] = sh.alive[s] != 0.f; }
      na2 += live[s] ? 1 : 0;
    }
    na = na2;
    __syncthreads();
    if (tid == 0 && na > 0) {
      mbar_expect_tx(&sh.xbar[0], (unsigned)na * D * 4u);          // inbox A: y1 of the next token's layer 0
      mbar_expect_tx(&sh.xbar[1], (unsigned)na * D * 4u);          // inbox B: its att vectors (armed here because the live count may have changed)
    }
    mark(p, 21);
    if (na == 0) break;
  }
  for (int i = 0; i < RINGN; ++i) {
    mbar_wait(&sh.wbar[warp][use_i], use_par);
    if (++use_i == RINGN) { use_i = 0; use_par ^= 1u; }
  }
  cluster_sync_all();
}

template <typename T, int NB>
int launch_cln(gsv_gpt_ctx* ctx, int live, int n_steps, cudaStream_t st) {
  const int nd = ctx->p.d / 256;
  void* fn = nullptr;
  if (nd == 2) fn = (void*)gpt_decode_cln_kernel<T, 2, NB>;
  else if (nd == 1) fn = (void*)gpt_decode_cln_kernel<T, 1, NB>;
  else return GSV_ERR_ARG;
  const int D = ctx->p.d, F = ctx->p.F, H = ctx->p.H;
  const size_t floats = (size_t)NB * D + (size_t)NB * F + 2 * NB * GSV_HEAD_DIM + D + ((GSV_SAMPLE_SMEM_FLOATS + 3) & ~3);
  const size_t bytes = floats * sizeof(float) + (size_t)NWARP * RingOf<NB>::v * D * 2;
  GSV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  GSV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  const int n_clusters = (live + NB - 1) / NB;
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

int gsv_gpt_decode_cln_launch(gsv_gpt_ctx* ctx, int live_slots, int nb, int n_steps, cudaStream_t st) {
  const bool f16 = ctx->dims.dtype == GSV_F16;
  if (nb <= 2) return f16 ? launch_cln<__half, 2>(ctx, live_slots, n_steps, st) : launch_cln<__nv_bfloat16, 2>(ctx, live_slots, n_steps, st);
  return f16 ? launch_cln<__half, 4>(ctx, live_slots, n_steps, st) : launch_cln<__nv_bfloat16, 4>(ctx, live_slots, n_steps, st);
}
