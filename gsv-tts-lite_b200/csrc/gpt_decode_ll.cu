// gpt_decode_ll.cu -- latency-optimised persistent decode kernel for 1..4 live sequences.
//
// Same arithmetic as gpt_decode.cu (reference t2s_model.py:67-105, 129-143, 442-456), different
// plumbing.  At batch 1 a token is a chain of ~120 dependent skinny GEMVs; what limits it is not HBM
// bandwidth but (a) how fast one phase's result reaches all SMs, (b) the tail latency of HBM loads
// sitting on that chain, and (c) instruction fetch.  What was measured on B200 and what follows from it:
//
// (1) Exchange.  Every value that crosses CTAs travels as an 8-byte {value, tag} word (the "LL"
//     protocol of NCCL's low-latency path): producers store it with one st.relaxed.gpu.v2, consumers
//     spin on the data itself until the tag of the expected phase shows up.  tools/ubench/ll_chain.cu:
//     ~1.5-1.7 k cycles per all-to-all exchange of 512 words, independent of the CTA count, against
//     ~3.5-3.9 k for grid barrier + load.
// (2) HBM off the chain.  Each warp bulk-prefetches into L2 (cp.async.bulk.prefetch.L2) the weight row
//     it will need one layer later; the row itself is loaded into registers before the spin of its phase
//     and only converted after it (a warp stalls at the first *use* of an outstanding load).
// (3) Instruction fetch.  Code that runs once per phase runs at I-cache-miss speed when the per-layer
//     body exceeds the L1 I-cache: the first versions (48-100 KB of SASS per layer, 66-80 % icc hit
//     rate) spent ~54 k cycles per layer, i.e. ~150 cycles per 128-byte line of straight-line code
//     (profiles/r01_*).  So ALL GEMV phases (QKV, out-proj, MLP up, MLP down, head) execute the same
//     routine, driven by a descriptor table; MLP-down rows are cut into 4 K-quarters so that every unit
//     of work is the same D-wide dot product.
//
// Per layer, 5 phases:  QKV gemv | split-KV attention | out-proj (+residual) | LN1 + MLP up + ReLU |
// MLP down (+residual); LN2 is applied by the consumers of y2.  Per token: + head, + sampling.
//
// Safety of single-buffered exchange buffers: every CTA produces rows in every GEMV phase and consumes
// the full vector of the previous one, so no CTA can run more than one phase ahead of the slowest
// (DESIGN.md "LL hazards").
#include <cstdlib>

#include "gpt_sample.cuh"

namespace {

constexpr int NT = GSV_DECODE_THREADS;   // 512
constexpr int NWARP = NT / 32;           // 16
constexpr int MAXB = 4;                  // live sequences this kernel family handles (NB = 1, 2 or 4 compiled)
constexpr int NSMAX = GSV_NSPLIT_MAX;    // stride of the partial buffer
constexpr int LL_NS = 4;                 // split-KV factor cap (more positions -> more passes per CTA)
constexpr int POS_PER_CTA = NWARP * 8;   // 128 positions per pass

__device__ __forceinline__ int split_pos(int k, int K) {
  const int ch = k >> 3, j = k & 7;
  return j < 4 ? ch * 4 + j : (K >> 1) + ch * 4 + (j - 4);
}

// raw 16-bit load that the compiler cannot consume early (asm volatile keeps program order with
// the polling loads; the conversion happens where the value is used)
__device__ __forceinline__ unsigned short ld_raw16(const void* p) {
  unsigned short v;
  asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}
template <typename T> __device__ __forceinline__ float raw16_to_f(unsigned short v);
template <> __device__ __forceinline__ float raw16_to_f<__half>(unsigned short v) { return __half2float(__ushort_as_half(v)); }
template <> __device__ __forceinline__ float raw16_to_f<__nv_bfloat16>(unsigned short v) { return __uint_as_float((unsigned)v << 16); }

// Bulk prefetch of `bytes` (multiple of 16, 16-byte aligned) into L2.
__device__ __forceinline__ void l2_prefetch(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---- per-warp weight ring in shared memory: cp.async two layers ahead (no global burst at phase start) ----
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
constexpr int RING_LAYERS = 2;           // layers in flight
constexpr int RING_UNITS = 4;            // QKV row, out-proj row, MLP-up row, MLP-down quarter row

// partial dot of one D-wide weight row segment (registers) with one staged activation row (shared,
// split layout); two independent accumulators halve the dependent FMA chain
template <typename T, int NCH>
__device__ __forceinline__ float dot_row(const uint4 (&w)[NCH], const float* xs, int lane) {
  constexpr int K = NCH * 256;
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    float wf[8];
    unpack8<T>(w[c], wf);
    const int ch = c * 32 + lane;
    const float4 lo = *reinterpret_cast<const float4*>(xs + ch * 4);
    const float4 hi = *reinterpret_cast<const float4*>(xs + (K >> 1) + ch * 4);
    a = fmaf(wf[0], lo.x, a); a = fmaf(wf[1], lo.y, a); a = fmaf(wf[2], lo.z, a); a = fmaf(wf[3], lo.w, a);
    b = fmaf(wf[4], hi.x, b); b = fmaf(wf[5], hi.y, b); b = fmaf(wf[6], hi.z, b); b = fmaf(wf[7], hi.w, b);
  }
  return a + b;
}

// timeline marker (tuning only: compiled in with -DGSV_TIMELINE, see tools/decode_timeline.py)
__device__ __forceinline__ void mark(const GptParams& p, int id) {
#ifdef GSV_TIMELINE
  if (p.prof != nullptr && threadIdx.x == 0) {            // every CTA records into its own region
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

// One GEMV phase, described at run time so that every phase of every layer runs the same instructions.
struct GemvDesc {
  const void* W;          // weights of layer 0, rows of w_ld elements
  long long w_lstride;    // elements between layers
  int w_ld;               // D or F
  const void* B;          // bias of layer 0 or null
  int b_lstride;
  int N;                  // output rows
  int kq;                 // 1, or 4: rows are cut into 4 K-quarters of D elements (w_ld = 4 D)
  int mode;               // staging: 0 LL vector of D (+LayerNorm), 1 split-KV partials, 2 LL vector of F = 4 D
  const uint2* in;        // LL input (modes 0 and 2)
  const void* gamma;      // LayerNorm parameters of layer 0 (stride D per layer) or null
  const void* beta;
  float* res_save;        // shared [NB][D]: staged input kept as a later residual, or null
  const float* res_add;   // shared [NB][D]: residual added to the output, or null
  int relu;
  uint2* out;             // LL output [slots][out_ld]
  int out_ld;
};

struct LLShared {
  int sl[MAXB];            // active slot ids
  int kv[MAXB];            // kv_len of each at the start of the step
  int nb;
  int NS;                  // split-KV factor of this step
  float red[NWARP][2 * MAXB];              // LayerNorm partial sums / K-quarter partial sums
  float qkv[3][GSV_HEAD_DIM];              // q (pre-scaled), k_new, v_new of this CTA's head
  float wpart[NWARP][GSV_HEAD_DIM + 2];    // per-warp attention partials
  float wscale[NWARP];
  float outv[MAXB][NWARP];                 // this trip's outputs, gathered for one coalesced LL store
  GemvDesc desc[5];        // 0 QKV, 1 out-proj, 2 MLP up, 3 MLP down, 4 head
  const GptParams* pp;     // timeline builds only
};

// ---- staging (shared by all GEMV phases) ----------------------------------------------------------------
template <typename T, int NB>
__device__ __forceinline__ void stage_input(const GemvDesc& d, const uint2* in, unsigned tag, bool ln, unsigned short g_raw,
                                            unsigned short b_raw, float* xs, int D, int H, const uint2* ll_part, LLShared& sh) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nb = sh.nb;
  const int F = 4 * D;
  if (d.mode == 0) {
    // one element of the D-vector per thread and slot; optional LayerNorm over D
    const bool mine = tid < D;
    float v[NB];
#pragma unroll
    for (int s = 0; s < NB; ++s) {
      v[s] = 0.f;
      if (s < nb && mine) v[s] = ll_wait(in + (size_t)sh.sl[s] * D + tid, tag);
    }
#ifdef GSV_TIMELINE
    __syncthreads();
    mark(*sh.pp, 50);
#endif
    if (ln) {
      // one pass: sum and sum of squares, reduced by all warps (values are O(1), D <= 512: E[v^2]-mean^2
      // in fp32 is accurate to ~1e-6 relative)
      float ps[2 * NB];
#pragma unroll
      for (int s = 0; s < NB; ++s) { ps[2 * s] = v[s]; ps[2 * s + 1] = v[s] * v[s]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < 2 * NB; ++i) ps[i] += __shfl_xor_sync(0xffffffffu, ps[i], o);
      }
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 2 * NB; ++i) sh.red[warp][i] = ps[i];
      }
      __syncthreads();
      const float g = raw16_to_f<T>(g_raw), b = raw16_to_f<T>(b_raw);
#pragma unroll
      for (int s = 0; s < NB; ++s) {
        float a = 0.f, q = 0.f;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) { a += sh.red[w][2 * s]; q += sh.red[w][2 * s + 1]; }
        const float mean = a / (float)D;
        const float rstd = rsqrtf(fmaxf(q / (float)D - mean * mean, 0.f) + 1e-5f);
        v[s] = (v[s] - mean) * rstd * g + b;
      }
    }
    if (mine) {
      const int pos = split_pos(tid, D);
#pragma unroll
      for (int s = 0; s < NB; ++s) {
        if (s < nb) {
          xs[s * F + pos] = v[s];
          if (d.res_save) d.res_save[s * D + tid] = v[s];
        }
      }
    }
  } else if (d.mode == 1) {
    // merge the split-KV partials of every (slot, head): thread = (head, dim)
    const int NS = sh.NS;
#pragma unroll 1
    for (int s = 0; s < nb; ++s) {
#pragma unroll 1
      for (int e = tid; e < D; e += NT) {
        const int h = e >> 5, dim = e & 31;
        const uint2* base = ll_part + ((size_t)(sh.sl[s] * H + h) * NSMAX) * GSV_PART_STRIDE;
        float mm[LL_NS], lw[LL_NS], oo[LL_NS];
        unsigned rdy = 0;
        const unsigned all = (1u << NS) - 1u;
        do {
#pragma unroll
          for (int k = 0; k < LL_NS; ++k) {
            if (k < NS && !((rdy >> k) & 1u)) {
              const uint2 a = ll_peek(base + k * GSV_PART_STRIDE);
              const uint2 b = ll_peek(base + k * GSV_PART_STRIDE + 1);
              const uint2 c = ll_peek(base + k * GSV_PART_STRIDE + 4 + dim);
              if (a.y == tag && b.y == tag && c.y == tag) {
                mm[k] = __uint_as_float(a.x); lw[k] = __uint_as_float(b.x); oo[k] = __uint_as_float(c.x);
                rdy |= 1u << k;
              }
            }
          }
        } while (rdy != all);
        float M = GSV_NEG_INF;
#pragma unroll
        for (int k = 0; k < LL_NS; ++k) if (k < NS) M = fmaxf(M, mm[k]);
        float Ls = 0.f, o = 0.f;
#pragma unroll
        for (int k = 0; k < LL_NS; ++k) {
          if (k < NS && mm[k] > GSV_NEG_INF) {
            const float sc = exp2f(mm[k] - M);
            Ls = fmaf(lw[k], sc, Ls);
            o = fmaf(oo[k], sc, o);
          }
        }
        xs[s * F + split_pos(e, D)] = o / Ls;
      }
    }
  } else {
    // F = 4 D words per slot: 4 quarters, each staged in the split layout of a D-wide operand
#pragma unroll 1
    for (int s = 0; s < nb; ++s) {
      const uint2* row = in + (size_t)sh.sl[s] * F;
      uint2 w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) w[e] = make_uint2(0u, ~tag);
      bool ok;
      do {
        ok = true;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = tid + e * NT;
          if (k < F && w[e].y != tag) {
            w[e] = ll_peek(row + k);
            ok = ok && (w[e].y == tag);
          }
        }
      } while (!ok);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = tid + e * NT;
        if (k < F) {
          const int qd = k / D, kk = k - qd * D;
          xs[s * F + qd * D + split_pos(kk, D)] = __uint_as_float(w[e].x);
        }
      }
    }
  }
  __syncthreads();
}

// ---- the one GEMV phase routine -----------------------------------------------------------------------------
// unit of work = (output row r, K-quarter qd): a D-wide dot product per live slot.  CTA c owns the
// CONTIGUOUS rows [c*rpc, c*rpc + rpc), rpc = ceil(N / G): its outputs are adjacent words of the LL buffer,
// gathered in shared memory and published by one warp with one coalesced store, so every 128-byte line of
// an exchange buffer has a single writer.  kq == 1: warp w computes row c*rpc + w (+16 j).  kq == 4: warp w
// computes quarter (w & 3) of row c*rpc + (w >> 2) (+4 j); the quarters are summed through shared memory.
template <typename T, int NCH_D, int NB>
__device__ __noinline__ void gemv_phase(const GptParams& p, int di, int layer, int ln_layer, const uint2* in_override,
                                        unsigned in_tag, unsigned out_tag, float* xs, const uint2* ll_part, int n_layers,
                                        uint4* ring, int gl, LLShared& sh) {
  constexpr int D = NCH_D * 256;
  const GemvDesc& d = sh.desc[di];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, nb = sh.nb;
  const int kq = d.kq;
  const int rpc = (d.N + G - 1) / G;
  const int row_begin = blockIdx.x * rpc, row_end = min(d.N, row_begin + rpc);
  const int rsub = kq == 4 ? (warp >> 2) : warp;             // row of this warp within a trip
  const int rstep = kq == 4 ? 4 : NWARP;                      // rows per trip
  const int r0 = row_begin + rsub;
  const int qd = kq == 4 ? (warp & 3) : 0;
  const T* W = reinterpret_cast<const T*>(d.W) + (size_t)layer * d.w_lstride + (size_t)qd * D;
  const T* B = d.B ? reinterpret_cast<const T*>(d.B) + (size_t)layer * d.b_lstride : nullptr;
  // ---- requests that do not depend on the input: issued before the spin.  Layer phases (di < 4) take
  //      their row segment from the warp's shared-memory ring (requested two layers ago); the head
  //      loads it directly.
  const bool use_ring = di < RING_UNITS;
  uint4* slot = ring + ((size_t)((gl % RING_LAYERS) * RING_UNITS + (use_ring ? di : 0)) * NWARP + warp) * (D / 8);
  uint4 w[NCH_D];
  unsigned short braw = 0, g_raw = 0, b_raw = 0;
  if (r0 < row_end) {
    if (!use_ring) {
      const uint4* src = reinterpret_cast<const uint4*>(W + (size_t)r0 * d.w_ld);
#pragma unroll
      for (int c = 0; c < NCH_D; ++c) w[c] = ld_weight(src + c * 32 + lane);
    }
    if (B) braw = ld_raw16(B + r0);
  }
  if (row_begin >= d.N) {
    // this CTA owns no row of this phase: it neither consumes the input nor produces output (nobody waits
    // on it); it only keeps the ring's commit-group count in step
    if (use_ring) cp_async_commit();
    return;
  }
  const bool ln = d.gamma != nullptr && ln_layer >= 0;
  if (ln && tid < D) {
    g_raw = ld_raw16(reinterpret_cast<const T*>(d.gamma) + (size_t)ln_layer * D + tid);
    b_raw = ld_raw16(reinterpret_cast<const T*>(d.beta) + (size_t)ln_layer * D + tid);
  }
#ifndef GSV_NO_L2PF
  if (warp == NWARP - 1 && lane == 0 && n_layers > 1) {
    const int ln2 = layer + 1 == n_layers ? 0 : layer + 1;
    if (d.B) l2_prefetch(reinterpret_cast<const T*>(d.B) + (size_t)ln2 * d.b_lstride, ((unsigned)d.N * (unsigned)sizeof(T) + 15u) & ~15u);
    if (d.gamma) {
      l2_prefetch(reinterpret_cast<const T*>(d.gamma) + (size_t)(ln_layer + 1 == n_layers ? 0 : ln_layer + 1) * D, D * (unsigned)sizeof(T));
      l2_prefetch(reinterpret_cast<const T*>(d.beta) + (size_t)(ln_layer + 1 == n_layers ? 0 : ln_layer + 1) * D, D * (unsigned)sizeof(T));
    }
  }
#endif
  // ---- wait for the input, stage it
  stage_input<T, NB>(d, in_override ? in_override : d.in, in_tag, ln, g_raw, b_raw, xs, D, p.H, ll_part, sh);
  mark(p, 40 + di);
#ifdef GSV_HASHTRACE
  // tuning: bit-exact fingerprint of the staged operand (CTA 0 only) to find the first phase that differs between runs
  if (p.prof != nullptr && blockIdx.x == 0 && warp == 0) {
    const int n = (d.mode == 2 ? 4 * D : D);
    unsigned h = 0;
    for (int k = lane; k < n; k += 32) h = h * 31u + __float_as_uint(xs[k]);
    h ^= __shfl_xor_sync(0xffffffffu, h, 16) * 3u; h ^= __shfl_xor_sync(0xffffffffu, h, 8) * 5u;
    h ^= __shfl_xor_sync(0xffffffffu, h, 4) * 7u; h ^= __shfl_xor_sync(0xffffffffu, h, 2) * 11u; h ^= __shfl_xor_sync(0xffffffffu, h, 1) * 13u;
    if (lane == 0) {
      const long long nrec = p.prof[0];
      if (nrec + 1 < p.prof_max) { p.prof[2 * (nrec + 1)] = di; p.prof[2 * (nrec + 1) + 1] = h; p.prof[0] = nrec + 1; }
    }
  }
#endif
  if (use_ring) {
    cp_async_wait<RING_LAYERS * RING_UNITS - 1>();       // this unit's group has landed (each lane reads back its own chunks)
    if (r0 < row_end) {
#ifdef GSV_NO_RING
      const uint4* src = reinterpret_cast<const uint4*>(W + (size_t)r0 * d.w_ld);
#pragma unroll
      for (int c = 0; c < NCH_D; ++c) w[c] = ld_weight(src + c * 32 + lane);
#else
#pragma unroll
      for (int c = 0; c < NCH_D; ++c) w[c] = slot[c * 32 + lane];
#endif
    }
  }
  // ---- dot products, epilogue, publish
  const int F = 4 * D;
  const int trips = (rpc + rstep - 1) / rstep;            // CTA-uniform (shared-memory syncs inside)
#pragma unroll 1
  for (int j = 0; j < trips; ++j) {
    const int r = r0 + j * rstep;
    const bool valid = r < row_end;
    if (j > 0 && valid) {
      const uint4* src = reinterpret_cast<const uint4*>(W + (size_t)r * d.w_ld);
#pragma unroll
      for (int c = 0; c < NCH_D; ++c) w[c] = ld_weight(src + c * 32 + lane);
      braw = B ? ld_raw16(B + r) : (unsigned short)0;
    }
    float acc[NB];
#pragma unroll
    for (int s = 0; s < NB; ++s) {
      acc[s] = 0.f;
      if (valid && s < nb) acc[s] = dot_row<T, NCH_D>(w, xs + s * F + qd * D, lane);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int s = 0; s < NB; ++s) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], o);
    }
    mark(p, 60 + di);
    if (kq == 4) {
      if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NB; ++s) sh.red[warp][s] = acc[s];
      }
      __syncthreads();
      if (qd == 0 && lane == 0) {
#pragma unroll
        for (int s = 0; s < NB; ++s) acc[s] = sh.red[warp][s] + sh.red[warp + 1][s] + sh.red[warp + 2][s] + sh.red[warp + 3][s];
      }
      __syncthreads();
    }
    if (valid && qd == 0 && lane == 0) {
      const float bias = B ? raw16_to_f<T>(braw) : 0.f;
#pragma unroll
      for (int s = 0; s < NB; ++s) {
        if (s < nb) {
          float v = acc[s] + bias;
          if (d.res_add) v += d.res_add[s * D + r];
          if (d.relu) v = fmaxf(v, 0.f);
          sh.outv[s][rsub] = v;
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      const int rr = row_begin + j * rstep + lane;            // one coalesced store of this trip's rows per slot
      if (lane < rstep && rr < row_end) {
#pragma unroll
        for (int s = 0; s < NB; ++s)
          if (s < nb) ll_store(d.out + (size_t)sh.sl[s] * d.out_ld + rr, sh.outv[s][lane], out_tag);
      }
    }
    if (j + 1 < trips) __syncthreads();
  }
  mark(p, 70 + di);
  if (use_ring) {
    // same unit, RING_LAYERS layers later, into the slot just consumed (every thread commits a group,
    // valid row or not, so that wait_group counts line up)
    if (r0 < row_end) {
      const int nl = (layer + RING_LAYERS) % n_layers;
      const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(d.W) + (size_t)nl * d.w_lstride +
                                                        (size_t)qd * D + (size_t)r0 * d.w_ld);
#pragma unroll
      for (int c = 0; c < NCH_D; ++c) cp_async16(slot + c * 32 + lane, src + c * 32 + lane);
    }
    cp_async_commit();
  }
}

// ---- attention phase for this CTA's item (slot, head, split) ------------------------------------------------------
template <typename T, int NB>
__device__ __noinline__ void attention_phase(const GptParams& p, int l, int D, const uint2* ll_qkv, uint2* ll_part,
                                             unsigned in_tag, unsigned out_tag, LLShared& sh) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x;
  const int H = p.H, S = p.S, NS = sh.NS;
  const int n_items = sh.nb * H * NS;
  if (cta >= n_items) return;                                        // whole CTA: no item, nothing to sync on
  const int it_s = cta / (H * NS), it_h = (cta / NS) % H, it_sp = cta % NS;
  const int slot = sh.sl[it_s];
  const int kvn = sh.kv[it_s];
  const int n = kvn + 1;                                            // positions 0..kv inclusive
  const int chunk = (((n + NS - 1) / NS) + 7) & ~7;
  const int tb = it_sp * chunk, te = min(n, tb + chunk);
  const bool newest = (te == n) && (tb < n);                        // this split holds position kv
  const int tec = min(te, kvn);                                     // cached positions of this split: [tb, tec)
  const int sub = lane & 3, pg = lane >> 2;
  const size_t head_base = ((size_t)(l * p.slots + slot) * H + it_h) * (size_t)S * GSV_HEAD_DIM;
  const T* kb = reinterpret_cast<const T*>(p.kc) + head_base + sub * 8;
  const T* vb = reinterpret_cast<const T*>(p.vc) + head_base + sub * 8;
  // first pass's K/V rows: requested before the spin
  const int pos0 = tb + warp * 8 + pg;
  uint4 kraw = make_uint4(0, 0, 0, 0), vraw = make_uint4(0, 0, 0, 0);
#ifndef GSV_LATE_KV
  if (pos0 < tec) {
    kraw = ld_cg16(kb + (size_t)pos0 * GSV_HEAD_DIM);
    vraw = ld_cg16(vb + (size_t)pos0 * GSV_HEAD_DIM);
  }
#endif
#ifndef GSV_NO_KVPF
  if (tid == 0 && l + 1 < p.L && tec > tb) {
    // next layer's K/V rows of this split into L2
    const size_t nxt = (size_t)p.slots * H * S * GSV_HEAD_DIM;
    const unsigned bytes = (unsigned)(tec - tb) * GSV_HEAD_DIM * (unsigned)sizeof(T);
    l2_prefetch(reinterpret_cast<const T*>(p.kc) + head_base + nxt + (size_t)tb * GSV_HEAD_DIM, bytes);
    l2_prefetch(reinterpret_cast<const T*>(p.vc) + head_base + nxt + (size_t)tb * GSV_HEAD_DIM, bytes);
  }
#endif
  // q_h (and k_h, v_h of the new token if this split holds it) from the QKV phase
  const float qscale = rsqrtf((float)GSV_HEAD_DIM) * 1.4426950408889634f;
  if (tid < 3 * GSV_HEAD_DIM && (tid < GSV_HEAD_DIM || newest)) {
    const int which = tid >> 5, c = tid & 31;                       // 0:q 1:k 2:v
    const float v = ll_wait(ll_qkv + (size_t)slot * 3 * D + which * D + it_h * GSV_HEAD_DIM + c, in_tag);
    if (which == 0) sh.qkv[0][c] = v * qscale;
    else {
      // the reference attends over the 16-bit cache entry it has just written
      const T t16 = Elem<T>::from_f(v);
      sh.qkv[which][c] = Elem<T>::to_f(t16);
      T* cache = reinterpret_cast<T*>(which == 1 ? p.kc : p.vc);
      cache[head_base + (size_t)kvn * GSV_HEAD_DIM + c] = t16;
    }
  }
  __syncthreads();
#ifdef GSV_LATE_KV
  if (pos0 < tec) {
    kraw = ld_cg16(kb + (size_t)pos0 * GSV_HEAD_DIM);
    vraw = ld_cg16(vb + (size_t)pos0 * GSV_HEAD_DIM);
  }
#endif
  // one cached position per 4 lanes and pass, 8 positions per warp, POS_PER_CTA per pass; online softmax per lane group
  float q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) q[j] = sh.qkv[0][sub * 8 + j];
  float mg = GSV_NEG_INF, lsum = 0.f, o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = 0.f;
  int pass = 0;
#pragma unroll 1
  for (int base = tb + warp * 8; base < tec; base += POS_PER_CTA, ++pass) {
    const int pos = base + pg;
    const bool ok = pos < tec;
    uint4 kr = kraw, vr = vraw;
    if (pass > 0 && ok) {
      kr = ld_cg16(kb + (size_t)pos * GSV_HEAD_DIM);
      vr = ld_cg16(vb + (size_t)pos * GSV_HEAD_DIM);
    }
    float kf[8], vf[8], s = 0.f;
    unpack8<T>(kr, kf);
    unpack8<T>(vr, vf);
#pragma unroll
    for (int j = 0; j < 8; ++j) s = fmaf(q[j], kf[j], s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (ok) {
      const float mn = fmaxf(mg, s);
      const float sc = exp2f(mg - mn), pr = exp2f(s - mn);      // mg = -inf on the first hit -> sc = 0
      lsum = fmaf(lsum, sc, pr);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(pr, vf[j], o[j] * sc);
      mg = mn;
    }
  }
  // warp maximum, then one rescale and plain sums over the 8 position groups
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
    // merge the 16 warp partials (+ the newest position from shared memory); lane = output dim
    const float mw = lane < NWARP ? sh.wpart[lane][0] : GSV_NEG_INF;
    float snew = GSV_NEG_INF;
    if (newest) snew = warp_sum(sh.qkv[0][lane] * sh.qkv[1][lane]);
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
    if (newest) {
      const float pr = exp2f(snew - M);
      Ls += pr;
      oa = fmaf(pr, sh.qkv[2][lane], oa);
    }
    uint2* out = ll_part + ((size_t)(slot * H + it_h) * NSMAX + it_sp) * GSV_PART_STRIDE;
    if (lane == 0) { ll_store(out, M, out_tag); ll_store(out + 1, Ls, out_tag); }
    ll_store(out + 4 + lane, oa, out_tag);
  }
}

template <typename T, int NCH_D, int NB>
__global__ void __launch_bounds__(NT, 1) gpt_decode_ll_kernel(const GptParams p, const int n_steps, const unsigned tag_base,
                                                              uint2* const ll_buf) {
  extern __shared__ __align__(16) float smem[];
  __shared__ LLShared sh;
  constexpr int D = NCH_D * 256, F = 4 * D;
  const int tid = threadIdx.x;
  const int cta = blockIdx.x, G = gridDim.x;
  const int H = p.H, L = p.L, V = p.V;
  // shared memory: xs[NB][F] (GEMV operand, split layout) | xres[NB][D] | xres1[NB][D]
  float* xs = smem;
  float* xres = smem + NB * F;
  float* xres1 = xres + NB * D;
  // the sampler's scratch aliases xs/xres (dead during sampling) but must not touch the ring (copies in flight)
  constexpr int kScratchFloats = (NB * F + 2 * NB * D) > GSV_SAMPLE_SMEM_FLOATS ? (NB * F + 2 * NB * D) : GSV_SAMPLE_SMEM_FLOATS;
  uint4* ring = reinterpret_cast<uint4*>(smem + ((kScratchFloats + 3) & ~3));   // [RING_LAYERS][RING_UNITS][NWARP][D/8] uint4
  // LL exchange buffers ({value,tag} words), per slot
  const size_t slots = p.slots;
  uint2* ll_xin = ll_buf;                                   // [slots][D]
  uint2* ll_qkv = ll_xin + slots * D;                       // [slots][3D]
  uint2* ll_y1 = ll_qkv + slots * 3 * D;                    // [slots][D]
  uint2* ll_y2 = ll_y1 + slots * D;                         // [slots][D]
  uint2* ll_h = ll_y2 + slots * D;                          // [slots][F]
  uint2* ll_part = ll_h + slots * F;                        // [slots][H][NSMAX][36]
  uint2* ll_logit = ll_part + slots * H * NSMAX * GSV_PART_STRIDE;   // [slots][VOCAB_MAX]
  uint2* ll_stat = ll_logit + slots * GSV_VOCAB_MAX;        // [slots]

  // ---- launch prologue: descriptor table, active slots, xin republished in LL form ----
  if (tid == 0) {
    sh.pp = &p;
    GemvDesc* d = sh.desc;
    d[0] = GemvDesc{p.w_qkv, (long long)3 * D * D, D, p.b_qkv, 3 * D, 3 * D, 1, 0, ll_y2, p.ln2_g, p.ln2_b, xres, nullptr, 0, ll_qkv, 3 * D};
    d[1] = GemvDesc{p.w_o, (long long)D * D, D, p.b_o, D, D, 1, 1, nullptr, nullptr, nullptr, nullptr, xres, 0, ll_y1, D};
    d[2] = GemvDesc{p.w_1, (long long)F * D, D, p.b_1, F, F, 1, 0, ll_y1, p.ln1_g, p.ln1_b, xres1, nullptr, 1, ll_h, F};
    d[3] = GemvDesc{p.w_2, (long long)D * F, F, p.b_2, D, D, 4, 2, ll_h, nullptr, nullptr, nullptr, xres1, 0, ll_y2, D};
    d[4] = GemvDesc{p.w_head, 0, D, nullptr, 0, V, 1, 0, ll_y2, p.ln2_g, p.ln2_b, nullptr, nullptr, 0, ll_logit, GSV_VOCAB_MAX};
  }
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
  // weight ring prologue: layers 0 and 1 of this warp's four units (one commit group per unit, in use order)
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int gl0 = 0; gl0 < RING_LAYERS; ++gl0) {
      for (int u = 0; u < RING_UNITS; ++u) {
        const GemvDesc& d = sh.desc[u];
        const int rpc = (d.N + G - 1) / G;
        const int r0 = cta * rpc + (d.kq == 4 ? (warp >> 2) : warp);
        const int qd = d.kq == 4 ? (warp & 3) : 0;
        if (r0 < min(d.N, cta * rpc + rpc)) {
          const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(d.W) + (size_t)(gl0 % L) * d.w_lstride +
                                                            (size_t)qd * D + (size_t)r0 * d.w_ld);
          uint4* slot = ring + ((size_t)(gl0 * RING_UNITS + u) * NWARP + warp) * (D / 8);
          for (int c = 0; c < NCH_D; ++c) cp_async16(slot + c * 32 + lane, src + c * 32 + lane);
        }
        cp_async_commit();
      }
    }
  }
  unsigned tag = tag_base;                                  // tag of the most recent publication (xin)
  int gl = 0;                                               // global layer counter (ring slot parity)
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
    if (tid == 0) {
      int kvmax = 0;
      for (int s = 0; s < nb; ++s) kvmax = max(kvmax, sh.kv[s]);
      sh.NS = max(1, min(min(LL_NS, G / (nb * H)), (kvmax + POS_PER_CTA) / POS_PER_CTA));
    }
    __syncthreads();
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      mark(p, 1);
      gemv_phase<T, NCH_D, NB>(p, 0, l, l - 1, l == 0 ? ll_xin : nullptr, tag, tag + 1, xs, ll_part, L, ring, gl, sh);   // QKV
      tag += 1;
      mark(p, 2);
      attention_phase<T, NB>(p, l, D, ll_qkv, ll_part, tag, tag + 1, sh);
      tag += 1;
      mark(p, 3);
      gemv_phase<T, NCH_D, NB>(p, 1, l, -1, nullptr, tag, tag + 1, xs, ll_part, L, ring, gl, sh);             // out-proj
      tag += 1;
      mark(p, 4);
      gemv_phase<T, NCH_D, NB>(p, 2, l, l, nullptr, tag, tag + 1, xs, ll_part, L, ring, gl, sh);              // MLP up
      tag += 1;
      mark(p, 5);
      gemv_phase<T, NCH_D, NB>(p, 3, l, -1, nullptr, tag, tag + 1, xs, ll_part, L, ring, gl, sh);             // MLP down
      tag += 1;
      gl += 1;
    }
    mark(p, 6);
    gemv_phase<T, NCH_D, NB>(p, 4, 0, L - 1, nullptr, tag, tag + 1, xs, ll_part, 1, ring, gl, sh);            // head
    tag += 1;
    // ---- sampling: one CTA per slot; everyone then learns who is still alive
    {
      const unsigned tag_logits = tag;
      tag += 1;                                   // tag of xin / status published by the samplers
      mark(p, 20);
      __syncthreads();
      if (cta < nb) {
        const int slot = sh.sl[cta];
        for (int v = tid; v < V; v += NT) smem[v] = ll_wait(ll_logit + (size_t)slot * GSV_VOCAB_MAX + v, tag_logits);
        __syncthreads();
        SampleLL io;
        io.preloaded = true;
        io.xin_ll = ll_xin + (size_t)slot * D;
        io.status_ll = ll_stat + slot;
        io.tag = tag;
        io.kv_len = sh.kv[cta] + 1;
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
}

template <typename T, int NB>
int launch_ll_nb(gsv_gpt_ctx* ctx, int n_steps, cudaStream_t st) {
  const int nd = ctx->p.d / 256;
  void* fn = nullptr;
  if (nd == 2) fn = (void*)gpt_decode_ll_kernel<T, 2, NB>;
  else if (nd == 1) fn = (void*)gpt_decode_ll_kernel<T, 1, NB>;
  else return GSV_ERR_ARG;
  const size_t act = (size_t)(NB * ctx->p.F + 2 * NB * ctx->p.d);
  const size_t scratch = act > (size_t)GSV_SAMPLE_SMEM_FLOATS ? act : (size_t)GSV_SAMPLE_SMEM_FLOATS;
  const size_t bytes = ((scratch + 3) & ~(size_t)3) * sizeof(float) +
                       (size_t)RING_LAYERS * RING_UNITS * NWARP * ctx->p.d * 2;   // + weight ring (16-bit rows)
  GSV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  GptParams p = ctx->p;
  int ns = n_steps;
  // tags never repeat between launches: launch sequence in the high 16 bits, phase counter below
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
  int grid = ctx->num_sms;
  if (ctx->decode_sms >= ctx->p.H * NB && ctx->decode_sms < ctx->num_sms) grid = ctx->decode_sms;   // the rest stay free for a concurrent stream
  {
    // tuning hook: fewer CTAs = fewer pollers per exchange, more rows per CTA
    static int env_grid = -1;
    if (env_grid < 0) { const char* e = getenv("GSV_LL_GRID"); env_grid = e ? atoi(e) : 0; }
    if (env_grid >= ctx->p.H * NB && env_grid <= ctx->num_sms) grid = env_grid;
  }
  GSV_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(NT), args, bytes, st));
  ctx->launches += 1;
  return GSV_OK;
}

template <typename T>
int launch_ll(gsv_gpt_ctx* ctx, int live, int n_steps, cudaStream_t st) {
  if (live <= 1) return launch_ll_nb<T, 1>(ctx, n_steps, st);
  if (live <= 2) return launch_ll_nb<T, 2>(ctx, n_steps, st);
  return launch_ll_nb<T, 4>(ctx, n_steps, st);
}

}  // namespace

size_t gsv_gpt_ll_buffer_bytes(const gsv_gpt_ctx* ctx) {
  const GptParams& p = ctx->p;
  size_t words = (size_t)p.slots * (6 * (size_t)p.d + p.F + (size_t)p.H * NSMAX * GSV_PART_STRIDE + GSV_VOCAB_MAX + 1);
  const size_t hx = gsv_gpt_hx_buffer_words(ctx);            // the head-cluster kernel lays its own areas over the same buffer
  if (hx > words) words = hx;
  return words * sizeof(uint2) + 256;
}

bool gsv_gpt_ll_supported(const gsv_gpt_ctx* ctx, int live_slots, int n_steps) {
  const int nd = ctx->p.d / 256;
  const bool shape = (nd == 1 || nd == 2) && ctx->p.F == 4 * ctx->p.d;
  // every (slot, head) needs a CTA; the 4 K-quarters of every MLP-down row must fit one CTA pass
  // structure (any grid works: rows loop); phases per launch < 2^16 so that tags are unique
  return shape && live_slots >= 1 && live_slots <= MAXB && ctx->p.H * live_slots <= ctx->num_sms &&
         (long long)n_steps * (5 * ctx->p.L + 2) < 65000;
}

int gsv_gpt_decode_ll_launch(gsv_gpt_ctx* ctx, int live_slots, int n_steps, cudaStream_t st) {
  if (ctx->dims.dtype == GSV_F16) return launch_ll<__half>(ctx, live_slots, n_steps, st);
  return launch_ll<__nv_bfloat16>(ctx, live_slots, n_steps, st);
}
