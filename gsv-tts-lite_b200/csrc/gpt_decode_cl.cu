// gpt_decode_cl.cu -- cluster decode kernel: ONE thread-block cluster per live sequence, one CTA per attention head.
//
// Same arithmetic as the other decode kernels (reference t2s_model.py:67-105, 129-143, 442-456).  What the
// timelines of the grid-wide kernels showed (profiles/r01_*): at 1..8 sequences a token is ~100 dependent
// exchanges, and an exchange through L2 between 148 polling CTAs costs 1.0-2.1 us, half of the token.  Here a
// sequence lives inside one cluster of H (= 16) CTAs on one GPC:
//   * every vector that crosses CTAs is PUSHED into the shared memory of all H CTAs with st.async, each store
//     completing 4 transaction bytes on the RECEIVER's mbarrier: a CTA waits only for its own inbox to fill (no
//     cluster-wide barrier on the chain; barrier.cluster measured 1.1 us per phase here, the inbox wait ~0.3);
//   * CTA h owns head h: its 96 Wqkv rows, its K/V stream and its attention never leave the SM; O / MLP rows are
//     split H ways (32 / 128 / 32 rows per CTA = 2 / 8 / 2 per warp);
//   * weights stream from HBM through a per-warp ring of bulk copies (cp.async.bulk, one mbarrier per slot; R units
//     of one D-wide row segment, consumed in a fixed cyclic order: 6 QKV + 2 O + 8 MLP-up + 2x4 MLP-down = 24 units
//     per layer), ~10 KB in flight per warp (a 16-byte cp.async ring sustained only 36 GB/s per SM);
//   * sequences do not interact at all: B live sequences = B clusters (grid = B x H), no cross-cluster traffic; the
//     152 MB weight stream of concurrent clusters is shared through L2.
// Measured limits of this design (B200): the cluster lives on ONE GPC, whose memory port sustains ~0.7-0.8 TB/s
// (16 SMs x ~45 GB/s; tools/ubench/bulk_bw.cu reaches 200 GB/s per SM only when the 16 CTAs sit on different GPCs),
// so one sequence cannot go below ~190 us/token here however the copies are shaped (per-warp 1 KB units, below, and
// CTA-wide 32 KB chunks were both tried: 355 vs 400 us/token).  A single sequence is therefore still served by
// the grid-wide flag-in-data kernel (299 us); from two sequences up the clusters win because they scale with the
// number of GPCs (6 co-resident clusters: 16.8k tok/s) and never exchange anything through L2.
// Two inboxes alternate between consecutive phases.  A CTA re-arms an inbox right after reading it and before it
// pushes its own outputs; a peer can only start the next fill of that inbox after it has received those outputs,
// so a fill never overtakes the read of the previous one.
#include "gpt_cluster_common.cuh"

namespace {

struct ClShared {
  float q[GSV_HEAD_DIM], knew[GSV_HEAD_DIM], vnew[GSV_HEAD_DIM];
  float wpart[NWARP][GSV_HEAD_DIM + 2];
  float wscale[NWARP];
  float alive;             // pushed by the sampler CTA together with the next input (inbox A)
  int alive_i;             // sampler CTA: written by sample_slot
  int slot, kv;
  uint64_t wbar[NWARP][RING];   // weight ring: one mbarrier per warp and slot
  uint64_t xbar[3];             // inboxes: 0 bufA (xin / y1 / y2), 1 bufB (att / h), 2 logits (CTA 0)
};

template <typename T, int NCH>
__global__ void __launch_bounds__(NT, 1) gpt_decode_cl_kernel(const GptParams p, const int n_steps, unsigned* const resident) {
  extern __shared__ __align__(16) float smem[];
  __shared__ ClShared sh;
  constexpr int D = NCH * 256, F = 4 * D;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, L = p.L, V = p.V, S = p.S;
  const unsigned rank = cluster_rank();               // = head index
  const int cid = blockIdx.x / H;                      // cluster index = index among live slots

  // shared memory: bufA[F] | bufB[F] staging (alternating phases; vectors in split layout per D-wide piece) |
  // xn[D] normalised layer input (residual of the O phase) | x1n[D] (residual of the MLP-down phase) |
  // logits[VOCAB_MAX] + xin_s[D] + sampler scratch (CTA 0) | per-warp weight ring [NWARP][RING][D/8] uint4
  float* bufA = smem;
  float* bufB = bufA + F;
  float* xn = bufB + F;
  float* x1n = xn + D;
  float* xin_s = x1n + D;
  float* samp = xin_s + D;                              // GSV_SAMPLE_SMEM_FLOATS (logits live in samp[0..V))
  uint4* ring = reinterpret_cast<uint4*>(samp + ((GSV_SAMPLE_SMEM_FLOATS + 3) & ~3)) + (size_t)warp * RING * (D / 8);

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

  // ---- which sequence: the cid-th active slot ----
  if (tid < 32) {
    const int flag = tid < p.slots ? ld_cg(p.active + tid) : 0;
    const unsigned m = __ballot_sync(0xffffffffu, flag != 0);
    int mine = -1;
    if (flag && __popc(m & ((1u << tid) - 1u)) == cid) mine = tid;
    const unsigned who = __ballot_sync(0xffffffffu, mine >= 0);
    if (tid == 0) {
      sh.slot = who ? (__ffs(who) - 1) : -1;
      sh.kv = who ? ld_cg(p.kv_len + (__ffs(who) - 1)) : 0;
      sh.alive = who ? 1.f : 0.f;
    }
  }
  __syncthreads();
  const int slot = sh.slot;
  if (slot < 0) { if (tid == 0) atomicAdd(resident, 1u); return; }   // whole cluster: no such live sequence (uniform across its CTAs)
  int kv = sh.kv;

  // ---- weight unit sequence of this warp (consumed strictly in this order, layer after layer) ----
  auto unit_src = [&](int l, int u) -> const T* {
#ifdef GSV_EXP_L0
    l = 0;      // experiment: every layer reads layer 0's weights (L2-resident) -- timing only
#endif
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
  int iss_l = 0, iss_u = 0;                             // next unit to request (every lane tracks the cursor)
  int use_i = 0;                                        // ring slot of the next unit to consume
  unsigned use_par = 0;                                 // parity of that slot's current fill
  constexpr unsigned UNIT_BYTES = D * (unsigned)sizeof(T);
  // request the unit `ahead` positions after the cursor into ring slot slot_i (one lane)
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
  // take the next N units (N <= RING): wait for their bulk copies, read them into registers
  auto take_n = [&](auto& w) {
    constexpr int N = (int)(sizeof(w) / sizeof(w[0]));
#pragma unroll
    for (int i = 0; i < N; ++i) {
      int si = use_i + i;
      unsigned par = use_par;
      if (si >= RING) { si -= RING; par ^= 1u; }
#ifndef GSV_EXP_NOW
      mbar_wait(&sh.wbar[warp][si], par);
#endif
      const uint4* srcs = ring + (size_t)si * (D / 8);
#pragma unroll
      for (int c = 0; c < NCH; ++c) w[i][c] = srcs[c * 32 + lane];
    }
  };
  // ... and once every lane has consumed them, lanes 0..N-1 refill the N slots with the units RING ahead
  auto release_n = [&](int n) {
    __syncwarp();
#ifndef GSV_EXP_NOW
    if (lane < n) {
      int si = use_i + lane;
      if (si >= RING) si -= RING;
      issue_at(si, lane);
    }
#endif
    advance_cursor(n);
    use_i += n;
    if (use_i >= RING) { use_i -= RING; use_par ^= 1u; }
  };
  if (lane == 0) {
    for (int i = 0; i < RING; ++i) mbar_init(&sh.wbar[warp][i], 1);
    if (tid == 0) { mbar_init(&sh.xbar[0], 1); mbar_init(&sh.xbar[1], 1); mbar_init(&sh.xbar[2], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < RING; ++i) issue_at(i, i);
  }
  advance_cursor(RING);
  __syncwarp();
  unsigned parA = 0, parB = 0, parL = 0;                // parities of the next fill of each inbox
  if (tid == 0) {
    mbar_expect_tx(&sh.xbar[0], D * 4u);                // first fill of inbox A: y1 of layer 0 (its input is loaded below)
    mbar_expect_tx(&sh.xbar[1], D * 4u);                // first fill of inbox B: att of layer 0
    if (rank == 0) mbar_expect_tx(&sh.xbar[2], (unsigned)V * 4u);
  }

  // layer-0 input of the first step: xin left by prefill / the previous launch (fp32)
  for (int k = tid; k < D; k += NT) bufA[split_pos(k, D)] = ld_cg(p.xin + (size_t)slot * D + k);
  __syncthreads();
  cluster_sync_all();                                   // every CTA of the cluster is resident and initialised
  if (tid == 0) atomicAdd(resident, 1u);                // gsv_gpt_wait_resident

  const int sub = lane & 3, pg = lane >> 2;
  uint4 gv[NCH], bv[NCH];                                // LayerNorm gamma / beta of the NEXT LayerNorm, requested one phase early
#pragma unroll
  for (int c = 0; c < NCH; ++c) { gv[c] = make_uint4(0, 0, 0, 0); bv[c] = gv[c]; }
#pragma unroll 1
  for (int step = 0; step < n_steps; ++step) {
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      const int ln = l + 1 == L ? 0 : l + 1;
      if (warp == NWARP - 1 && lane == 0 && L > 1) {
        // small per-layer vectors of the next layer into L2 (evicted by the weight stream since the last token)
        l2_prefetch(G2 + (size_t)l * D, D * (unsigned)sizeof(T));
        l2_prefetch(Be2 + (size_t)l * D, D * (unsigned)sizeof(T));
        l2_prefetch(G1 + (size_t)ln * D, D * (unsigned)sizeof(T));
        l2_prefetch(Be1 + (size_t)ln * D, D * (unsigned)sizeof(T));
        l2_prefetch(Bqkv + (size_t)ln * 3 * D, 3 * D * (unsigned)sizeof(T));
        l2_prefetch(Bo + (size_t)ln * D, D * (unsigned)sizeof(T));
        l2_prefetch(B1 + (size_t)ln * F, F * (unsigned)sizeof(T));
        l2_prefetch(B2 + (size_t)ln * D, D * (unsigned)sizeof(T));
      }
      // ================= A: x = l == 0 ? xin : LN2(y2) (bufA) ; q,k,v of this head ; attention -> att (bufB) =========
      mark(p, 1);
      {
        const size_t head_base = ((size_t)(l * p.slots + slot) * H + rank) * (size_t)S * GSV_HEAD_DIM;
        const T* kb = reinterpret_cast<const T*>(p.kc) + head_base + sub * 8;
        const T* vb = reinterpret_cast<const T*>(p.vc) + head_base + sub * 8;
        uint4 kr0 = make_uint4(0, 0, 0, 0), vr0 = kr0, kr1 = kr0, vr1 = kr0;
        {
          const int p0 = warp * 8 + pg, p1 = p0 + NWARP * 8;
          if (p0 < kv) { kr0 = ld_cg16(kb + (size_t)p0 * GSV_HEAD_DIM); vr0 = ld_cg16(vb + (size_t)p0 * GSV_HEAD_DIM); }
          if (p1 < kv) { kr1 = ld_cg16(kb + (size_t)p1 * GSV_HEAD_DIM); vr1 = ld_cg16(vb + (size_t)p1 * GSV_HEAD_DIM); }
        }
        if (tid == 0 && l + 1 < L && kv > 0) {
          const size_t nxt = (size_t)p.slots * H * S * GSV_HEAD_DIM;
          const unsigned bytes = (unsigned)kv * GSV_HEAD_DIM * (unsigned)sizeof(T);
          l2_prefetch(reinterpret_cast<const T*>(p.kc) + head_base + nxt, bytes);
          l2_prefetch(reinterpret_cast<const T*>(p.vc) + head_base + nxt, bytes);
        }
        float bq = 0.f;                                 // lane it < 6 holds the bias of the warp's it-th q/k/v row
        if (lane < QKV_PER_WARP) {
          const int rr = warp + NWARP * lane;
          bq = Elem<T>::to_f(Bqkv[(size_t)l * 3 * D + (rr >> 5) * D + rank * GSV_HEAD_DIM + (rr & 31)]);
        }
        if (l > 0) { mbar_wait(&sh.xbar[0], parA); parA ^= 1u; }      // y2 of the previous layer has arrived
        float xv[NCH * 8];
        load_x<NCH>(bufA, lane, xv);
        if (l > 0) {
          __syncthreads();                              // every warp has read inbox A: re-arm it for this layer's y1
          if (tid == 0) mbar_expect_tx(&sh.xbar[0], D * 4u);
        }
        if (l > 0) {
          float mean, rstd;
          ln_stats<NCH>(xv, mean, rstd);
          ln_apply<T, NCH>(xv, mean, rstd, gv, bv);       // requested during the previous MLP-down phase
        }
        if (warp == 0) store_x<NCH>(xn, lane, xv);      // residual copy (read after the next barrier)
        {
          uint4 w[QKV_PER_WARP][NCH];
          take_n(w);
          float part[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) part[it] = it < QKV_PER_WARP ? dot_regs<T, NCH>(w[it < QKV_PER_WARP ? it : 0], xv) : 0.f;
          release_n(QKV_PER_WARP);
          const float sum = reduce8(part, lane);          // lane holds row idx(lane)
          const int it = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
          const float bias = __shfl_sync(0xffffffffu, bq, it);
          if ((lane & 3) == 0 && it < QKV_PER_WARP) {
            const int rr = warp + NWARP * it, which = rr >> 5, c = rr & 31;
            const float v = sum + bias;
            if (which == 0) {
              sh.q[c] = v * (rsqrtf((float)GSV_HEAD_DIM) * 1.4426950408889634f);
            } else {
              const T t16 = Elem<T>::from_f(v);         // the reference attends over the 16-bit cache entry it has just written
              (which == 1 ? sh.knew : sh.vnew)[c] = Elem<T>::to_f(t16);
              T* cache = reinterpret_cast<T*>(which == 1 ? p.kc : p.vc);
              cache[head_base + (size_t)kv * GSV_HEAD_DIM + c] = t16;
            }
          }
        }
        __syncthreads();
        mark(p, 50);
        float q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) q[j] = sh.q[sub * 8 + j];
        float mg = GSV_NEG_INF, lsum = 0.f, o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
        int pass = 0;
#pragma unroll 1
        for (int base = warp * 8; base < kv; base += NWARP * 8, ++pass) {
          const int pos = base + pg;
          const bool ok = pos < kv;
          uint4 kr = pass == 0 ? kr0 : kr1, vr = pass == 0 ? vr0 : vr1;
          if (pass > 1 && ok) {
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
        // every warp merges the 16 partials (cheap) so that warp t can push the head's output to CTA t
        {
          const float mw = lane < NWARP ? sh.wpart[lane][0] : GSV_NEG_INF;
          const float snew = warp_allsum(sh.q[lane] * sh.knew[lane]);
          const float M = fmaxf(warp_max(mw), snew);
          float Ls = 0.f, oa = 0.f;
#pragma unroll
          for (int w = 0; w < NWARP; ++w) {
            const float mwv = sh.wpart[w][0];
            const float sc = mwv > GSV_NEG_INF ? exp2f(mwv - M) : 0.f;
            Ls = fmaf(sh.wpart[w][1], sc, Ls);
            oa = fmaf(sh.wpart[w][2 + lane], sc, oa);
          }
          const float pr = exp2f(snew - M);
          Ls += pr;
          oa = fmaf(pr, sh.vnew[lane], oa);
          if (warp < H) st_async(bufB + split_pos((int)rank * GSV_HEAD_DIM + lane, D), &sh.xbar[1], (unsigned)warp, oa / Ls);
        }
      }
      mark(p, 52);
      mbar_wait(&sh.xbar[1], parB); parB ^= 1u;          // att of every head has arrived
      mark(p, 3);
      // ================= O: y1 = x + att Wo^T + bo (reads bufB, writes bufA) ==================
      {
        load_vec<T, NCH>(G1 + (size_t)l * D, lane, gv);      // for the LayerNorm of the next phase
        load_vec<T, NCH>(Be1 + (size_t)l * D, lane, bv);
        const int o_row = (int)rank * GSV_HEAD_DIM + warp * O_PER_WARP + (lane >> 4);
        const float o_bias = Elem<T>::to_f(Bo[(size_t)l * D + o_row]);
        float xv[NCH * 8];
        load_x<NCH>(bufB, lane, xv);
        __syncthreads();                                // every warp has read inbox B: re-arm it for h
        if (tid == 0) mbar_expect_tx(&sh.xbar[1], F * 4u);
        {
          uint4 w[O_PER_WARP][NCH];
          take_n(w);
          const float p0 = dot_regs<T, NCH>(w[0], xv), p1 = dot_regs<T, NCH>(w[1], xv);
          release_n(O_PER_WARP);
          const float sum = reduce2(p0, p1, lane);       // lanes 0..15: row 0, lanes 16..31: row 1
          const int tgt = lane & 15;
          const float v = sum + o_bias + xn[split_pos(o_row, D)];
          if (tgt < H) st_async(bufA + split_pos(o_row, D), &sh.xbar[0], (unsigned)tgt, v);
        }
      }
      mark(p, 53);
      mbar_wait(&sh.xbar[0], parA); parA ^= 1u;          // y1 has arrived
      mark(p, 4);
      // ================= M1: x1 = LN1(y1); h = relu(x1 W1^T + b1) (reads bufA, writes bufB) ==================
      {
        float xv[NCH * 8];
        load_x<NCH>(bufA, lane, xv);
        __syncthreads();                                // re-arm inbox A for y2 (or, after the last layer's head, the next input)
        if (tid == 0) mbar_expect_tx(&sh.xbar[0], D * 4u);
        const int m1_i = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);   // this lane's row after reduce8
        const int m1_row = (int)rank * (4 * GSV_HEAD_DIM) + warp * M1_PER_WARP + m1_i;
        const float m1_bias = Elem<T>::to_f(B1[(size_t)l * F + m1_row]);
        float mean, rstd;
        ln_stats<NCH>(xv, mean, rstd);
        ln_apply<T, NCH>(xv, mean, rstd, gv, bv);
        if (warp == 0) store_x<NCH>(x1n, lane, xv);
        {
          uint4 w[M1_PER_WARP][NCH];
          take_n(w);
          float part[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) part[i] = dot_regs<T, NCH>(w[i], xv);
          release_n(M1_PER_WARP);
          const float v = fmaxf(reduce8(part, lane) + m1_bias, 0.f);
          const int qd = m1_row / D, kk = m1_row - qd * D;
          float* dst = bufB + qd * D + split_pos(kk, D);
          // the 4 lanes that hold this row cover the 16 target CTAs: targets (lane & 3) * 4 + j
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int tgt = (lane & 3) * 4 + j;
            if (tgt < H) st_async(dst, &sh.xbar[1], (unsigned)tgt, v);
          }
        }
      }
      mark(p, 54);
      mbar_wait(&sh.xbar[1], parB); parB ^= 1u;          // h has arrived
      mark(p, 5);
      // ================= M2: y2 = x1 + h W2^T + b2 (reads bufB, writes bufA) ==================
      {
        load_vec<T, NCH>(G2 + (size_t)l * D, lane, gv);      // for the LayerNorm of the next layer / the head
        load_vec<T, NCH>(Be2 + (size_t)l * D, lane, bv);
        const int m2_row = (int)rank * GSV_HEAD_DIM + warp * M2_PER_WARP + (lane >> 4);
        const float m2_bias = Elem<T>::to_f(B2[(size_t)l * D + m2_row]);
        float p0 = 0.f, p1 = 0.f;
        {
          uint4 w[4 * M2_PER_WARP][NCH];
          take_n(w);
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            float xv[NCH * 8];
            load_x<NCH>(bufB + qd * D, lane, xv);
            p0 += dot_regs<T, NCH>(w[qd], xv);
            p1 += dot_regs<T, NCH>(w[4 + qd], xv);
          }
          release_n(4 * M2_PER_WARP);
        }
        __syncthreads();                                // every warp has read inbox B: re-arm it for the next layer's att
        if (tid == 0) mbar_expect_tx(&sh.xbar[1], D * 4u);
        {
          const float sum = reduce2(p0, p1, lane);
          const int tgt = lane & 15;
          const float v = sum + m2_bias + x1n[split_pos(m2_row, D)];
          if (tgt < H) st_async(bufA + split_pos(m2_row, D), &sh.xbar[0], (unsigned)tgt, v);
        }
      }
      mark(p, 55);
    }
    mark(p, 6);
    // ================= head: logits = LN2_last(y2) Whead^T, pushed to CTA 0 ==================
    {
      mbar_wait(&sh.xbar[0], parA); parA ^= 1u;          // y2 of the last layer has arrived
      float xv[NCH * 8];
      load_x<NCH>(bufA, lane, xv);
      __syncthreads();                                  // re-arm inbox A for the next input (+ the alive flag)
      if (tid == 0) mbar_expect_tx(&sh.xbar[0], D * 4u + 4u);
      float mean, rstd;
      ln_stats<NCH>(xv, mean, rstd);
      ln_apply<T, NCH>(xv, mean, rstd, gv, bv);             // G2/Be2 of the last layer, requested in its MLP-down phase
      // row g -> CTA g % H, warp (g / H) % 16, trip (g / H) / 16; the next row is in flight while this one is reduced
      uint4 w[NCH], wn[NCH];
      int j = warp;
      if ((int)rank + j * H < V) load_vec<T, NCH>(Wh + (size_t)((int)rank + j * H) * D, lane, w);
#pragma unroll 1
      for (; (int)rank + j * H < V; j += NWARP) {
        const int g = (int)rank + j * H, gn = g + NWARP * H;
        if (gn < V) load_vec<T, NCH>(Wh + (size_t)gn * D, lane, wn);
        const float a = warp_allsum(dot_regs<T, NCH>(w, xv));
        if (lane == 0) st_async(samp + g, &sh.xbar[2], 0u, a);
#pragma unroll
        for (int c = 0; c < NCH; ++c) w[c] = wn[c];
      }
    }
    mark(p, 20);
    // ================= sampling in CTA 0; next input and alive flag pushed to every CTA ==================
    if (rank == 0) {
      mbar_wait(&sh.xbar[2], parL); parL ^= 1u;          // all V logits have arrived
      SampleLL io;
      io.preloaded = true;
      io.xin_ll = nullptr;
      io.status_ll = nullptr;
      io.tag = 0;
      io.kv_len = kv + 1;
      io.xin_smem = xin_s;
      io.alive_smem = &sh.alive_i;
      sample_slot<T>(p, slot, samp, &io);
      __syncthreads();
      if (tid == 0) mbar_expect_tx(&sh.xbar[2], (unsigned)V * 4u);     // logits inbox re-armed before anyone can refill it
      const float alive = sh.alive_i ? 1.f : 0.f;
      for (int i = tid; i < H * D; i += NT) {
        const int tgt = i / D, k = i - tgt * D;
        st_async(bufA + split_pos(k, D), &sh.xbar[0], (unsigned)tgt, xin_s[k]);
      }
      if (tid < H) st_async(&sh.alive, &sh.xbar[0], (unsigned)tid, alive);
    }
    mbar_wait(&sh.xbar[0], parA); parA ^= 1u;            // next input + alive flag have arrived
    const bool alive_now = sh.alive != 0.f;
    __syncthreads();
    if (tid == 0) mbar_expect_tx(&sh.xbar[0], D * 4u);    // inbox A re-armed for y1 of the next token's layer 0
    mark(p, 21);
    kv += 1;
    if (!alive_now) break;                              // uniform across the cluster
  }
  // drain this warp's outstanding bulk copies, then leave together: no CTA exits while a peer may still push into it
  for (int i = 0; i < RING; ++i) {
#ifndef GSV_EXP_NOW
    mbar_wait(&sh.wbar[warp][use_i], use_par);
#endif
    if (++use_i == RING) { use_i = 0; use_par ^= 1u; }
  }
  cluster_sync_all();                                   // no CTA exits while a peer may still push into its memory
}

template <typename T>
int launch_cl(gsv_gpt_ctx* ctx, int live, int n_steps, cudaStream_t st) {
  const int nd = ctx->p.d / 256;
  void* fn = nullptr;
  if (nd == 2) fn = (void*)gpt_decode_cl_kernel<T, 2>;
  else if (nd == 1) fn = (void*)gpt_decode_cl_kernel<T, 1>;
  else return GSV_ERR_ARG;
  const int D = ctx->p.d, F = ctx->p.F, H = ctx->p.H;
  const size_t floats = (size_t)2 * F + 3 * D + ((GSV_SAMPLE_SMEM_FLOATS + 3) & ~3);
  const size_t bytes = floats * sizeof(float) + (size_t)NWARP * RING * D * 2;
  GSV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  GSV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(live * H); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = bytes; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = H; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  GptParams p = ctx->p;
  int ns = n_steps;
  unsigned* resident = ctx->hx_resident;
  ctx->hx_resident_expected += (unsigned)(live * H);
  void* args[] = {&p, &ns, &resident};
  GSV_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  ctx->launches += 1;
  return GSV_OK;
}

}  // namespace

bool gsv_gpt_cl_supported(const gsv_gpt_ctx* ctx, int live_slots) {
  const int nd = ctx->p.d / 256;
  return (nd == 1 || nd == 2) && ctx->p.F == 4 * ctx->p.d && ctx->p.H >= 2 && ctx->p.H <= 16 && live_slots >= 1 &&
         ctx->p.V <= GSV_VOCAB_MAX;
}

int gsv_gpt_decode_cl_launch(gsv_gpt_ctx* ctx, int live_slots, int n_steps, cudaStream_t st) {
  if (ctx->dims.dtype == GSV_F16) return launch_cl<__half>(ctx, live_slots, n_steps, st);
  return launch_cl<__nv_bfloat16>(ctx, live_slots, n_steps, st);
}
