// conv_umma.cuh -- Conv1d / ConvTranspose1d as an implicit GEMM on the 5th-generation tensor cores.
//
// (included by vocoder.cu after ConvArgs / conv_epilogue_row8)
//
// GEMM view of  out[t][co] = sum_k sum_ci x[t + k*dil - pad][ci] * w[k][co][ci]  (reference F.conv1d call
// sites: models.py:114-130, modules.py:84-103, 190-203):
//     M = 128 consecutive time steps of one batch element        (TMEM lanes)
//     N = BN output channels                                     (TMEM columns, fp32 accumulators)
//     K = BK (64/32/16) input channels of one tap per MMA group  (BK/16 x tcgen05.mma M128 N(BN) K16)
// Activations are channels-last [B][T][C], so a tap is just a ROW SHIFT of the same matrix.  Per BK-channel
// chunk ONE TMA box of 128 + halo rows starting at row t0 - pad of the 3-D tensor map (C, T, B) is loaded
// (rows outside [0, T) are zero-filled by TMA = the convolution's zero padding, and a box never crosses into
// the next batch element); tap k reads it through a shared-memory descriptor whose start address is advanced
// by k*dil rows -- the 128/64/32-byte swizzle is a function of the shared-memory address, so a row-shifted
// view of a swizzled tile is still a valid K-major operand (verified on B200 against the oracle).  A bytes per
// tile therefore do not scale with the kernel width.
// Weights are tap-major [k][Cout][Cin], i.e. K-major N x K tiles: B tile = TMA box (BK, BN) of tap k, streamed
// through its own, deeper ring.
// ConvTranspose1d (models.py:120) runs in polyphase form: output phase r = (t + pad) mod s only sees the taps
// k = r + j*s, and frame index q = (t + pad) div s, so phase r is an ordinary convolution over input frames
// with row shift -j whose results land on the strided rows t = q*s + r - pad.  One CTA = one (M tile, N tile,
// batch element, phase).
//
// Warp roles (192 threads): warp 0 lane 0 = TMA producer, warp 1 = TMEM allocator + (lane 0) MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> fused conv epilogue -> global).  Two mbarrier rings: activation
// chunks (sa stages) and weight taps (sw stages).
#pragma once
#include <cuda.h>

namespace umma {

constexpr int BM = 128;          // time steps per tile
constexpr int kMaxSA = 16;       // activation ring depth (upper bound)
constexpr int kMaxSW = 12;       // weight ring depth (upper bound)
constexpr int kThreads = 192;
constexpr int kSmemBudget = 200 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const uint32_t a = smem_u32(bar);
  unsigned done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(a), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// Programmatic dependent launch: a kernel launched with the programmatic-serialization attribute may start while
// its predecessor is still running; everything that does not depend on the predecessor's output (barrier set-up,
// TMEM allocation, tensor-map fetch, the first WEIGHT tiles) runs before pdl_wait().
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], both K-major, 16-bit inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start address and
// offsets in 16-byte units, LBO unused (1), SBO = 8 rows, version 1, layout type 2 / 4 / 6 = 128 / 64 / 32-byte
// swizzle (row length BK * 2 bytes = swizzle span).
template <int BK>
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
  constexpr uint32_t row_bytes = BK * 2;
  constexpr uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t hi = ((8u * row_bytes) >> 4) | (1u << 14) | (layout << 29);
  return ((uint64_t)hi << 32) | lo;
}

template <typename T> struct UmmaFmt;
template <> struct UmmaFmt<__half> { static constexpr uint32_t v = 0; };
template <> struct UmmaFmt<__nv_bfloat16> { static constexpr uint32_t v = 1; };

template <typename T>
struct Params {
  CUtensorMap tm_a;      // activations (C, Tin, B), box (BK, a_rows, 1)
  CUtensorMap tm_w;      // weights (Cin, Cout, KW), box (BK, BN, 1)
  int kchunks;           // ceil(Cin / BK)
  int in_off;            // first input channel
  int KW, n_phase;       // taps; output phases (1: convolution, s: transposed convolution with stride s)
  int dil, pad;          // convolution: row shift of tap k is k*dil - pad
  int t_pad;             // transposed convolution: t = q*s + r - t_pad
  int m_ext;             // rows q in [0, m_ext) are computed
  int a_rows;            // rows of the activation box = 128 + halo
  int a_stage_bytes;     // a_rows * BK * 2 rounded up to 1024
  int sa, sw;            // ring depths
  int pdl_early;         // small grid (latency-bound): let the next kernel start its set-up as soon as this one is set up
  ConvArgs<T> ep;
};

template <typename T, int BN, int BK>
__global__ void __launch_bounds__(kThreads, 2) conv_umma_kernel(const __grid_constant__ Params<T> P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int ROW_BYTES = BK * 2, W_BYTES = BN * ROW_BYTES;
  constexpr int W_STAGE = (W_BYTES + 1023) & ~1023;
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  __shared__ uint64_t a_full[kMaxSA], a_empty[kMaxSA], w_full[kMaxSW], w_empty[kMaxSW], acc_bar;
  __shared__ uint32_t tmem_base_s;
  // carve (1024-byte aligned): [sa][A stage] | [sw][W stage]
  const uint32_t base_u = smem_u32(smem_raw);
  const uint32_t tiles_a = base_u + ((1024u - (base_u & 1023u)) & 1023u);
  const uint32_t tiles_w = tiles_a + (uint32_t)(P.sa * P.a_stage_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int b = blockIdx.z / P.n_phase, r = blockIdx.z - b * P.n_phase;
  const bool transposed = P.n_phase > 1;
  const int n_taps = transposed ? (P.KW - r + P.n_phase - 1) / P.n_phase : P.KW;
  const int SA = P.sa, SW = P.sw;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < SW; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(&acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tm_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tm_w) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  // Small grids are latency-bound: the next kernel may begin its own set-up right away.  Large grids keep the SMs
  // to themselves until their accumulators are done (a waiting successor CTA would only hold shared memory).
  if (P.pdl_early && threadIdx.x == 0) pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer: per channel chunk one activation box (with halo), then one weight box per tap ----
      const int row0 = q0 - (transposed ? n_taps - 1 : P.pad);
      // weights do not depend on the previous kernel: fill the weight ring first, then wait for the predecessor
      const int n_w = P.kchunks * n_taps;
      const int pre = n_w < SW ? n_w : SW;
      for (int i = 0; i < pre; ++i) {
        const int c = i / n_taps, j = i - c * n_taps;
        mbar_expect_tx(&w_full[i], W_BYTES);
        tma_load_3d(tiles_w + (uint32_t)(i * W_STAGE), &P.tm_w, &w_full[i], c * BK, n0, transposed ? r + j * P.n_phase : j);
      }
      pdl_wait();
      int i = 0;
      for (int c = 0; c < P.kchunks; ++c) {
        const int s = c % SA;
        if (c >= SA) mbar_wait(&a_empty[s], ((c / SA) - 1) & 1);
        mbar_expect_tx(&a_full[s], (unsigned)(P.a_rows * ROW_BYTES));
        tma_load_3d(tiles_a + (uint32_t)(s * P.a_stage_bytes), &P.tm_a, &a_full[s], P.in_off + c * BK, row0, b);
        for (int j = 0; j < n_taps; ++j, ++i) {
          if (i < pre) continue;                            // already in flight
          const int w = i % SW;
          mbar_wait(&w_empty[w], ((i / SW) - 1) & 1);
          mbar_expect_tx(&w_full[w], W_BYTES);
          tma_load_3d(tiles_w + (uint32_t)(w * W_STAGE), &P.tm_w, &w_full[w], c * BK, n0, transposed ? r + j * P.n_phase : j);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- MMA issuer ----
      constexpr uint32_t idesc = (1u << 4) | (UmmaFmt<T>::v << 7) | (UmmaFmt<T>::v << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(BM >> 4) << 24);
      int i = 0;
      for (int c = 0; c < P.kchunks; ++c) {
        const int s = c % SA;
        mbar_wait(&a_full[s], (c / SA) & 1);
        const uint32_t a_base = tiles_a + (uint32_t)(s * P.a_stage_bytes);
        for (int j = 0; j < n_taps; ++j, ++i) {
          const int w = i % SW;
          mbar_wait(&w_full[w], (i / SW) & 1);
          tc_fence_after();
          const int row_off = transposed ? n_taps - 1 - j : j * P.dil;
          const uint64_t ad = smem_desc<BK>(a_base + (uint32_t)(row_off * ROW_BYTES));
          const uint64_t bd = smem_desc<BK>(tiles_w + (uint32_t)(w * W_STAGE));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc_mma(tmem_d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (i > 0 || k > 0) ? 1u : 0u);
          tc_commit(&w_empty[w]);          // frees the weight slot when these MMAs have read it
        }
        tc_commit(&a_empty[s]);            // all taps of this chunk issued: frees the activation slot
      }
      tc_commit(&acc_bar);                 // accumulator complete
    }
  } else {
    // ---- epilogue: warp w owns TMEM lanes [32*(w%4), +32) = rows q0 + 32*(w%4) + lane ----
    const int quarter = warp & 3;
    pdl_wait();                            // the epilogue reads residual / accumulator streams of earlier kernels
    mbar_wait(&acc_bar, 0);
    tc_fence_after();
    if (!P.pdl_early && threadIdx.x == 64) pdl_launch_dependents();
    const int q = q0 + quarter * 32 + lane;
    const int t = transposed ? q * P.n_phase + r - P.t_pad : q;
    const bool row_ok = q < P.m_ext && t >= 0 && t < P.ep.Tout;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t v[16];
      tc_ld16(tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
      tc_ld_wait();
      if (row_ok) {
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
        conv_epilogue_row8<T>(P.ep, b, t, n0 + c0, f);
        conv_epilogue_row8<T>(P.ep, b, t, n0 + c0 + 8, f + 8);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
  }
}

// ---- weight-stationary, persistent variant (stride-1 convolutions on large grids) -------------------------------------
// What the launch lists of a B = 16..64 call showed (profiles/r02_voc_shares_*.txt): the one-tile-per-CTA kernel above spends
// 12-24 us per 128 x 128 tile of which 1-3 us are MMAs -- fill latency, the epilogue and the drain run one after the other,
// and every tile re-streams the whole weight tensor of its N tile from L2 (229 KB for a 128-channel, 7-tap layer against
// 34 KB of activations).  Here a CTA
//   * keeps ALL taps of its N tile resident in shared memory (loaded once, before the predecessor kernel has finished),
//   * walks over M tiles (blockIdx.x, += gridDim.x); only the activation boxes travel per tile, through a ring that holds
//     about two tiles,
//   * owns NACC accumulators in tensor memory: while the epilogue warps of set s drain accumulator s (tile i), the MMA
//     thread fills accumulator s+1 (tile i+1),
//   * has 4 NACC epilogue warps; a thread owns one time step of the tile and walks over its channels in groups of 16, with
//     the fp32 residual / MRF-accumulator values of the NEXT group already requested (the first group's before the
//     accumulator is even complete) and the rows of the next unit requested into L2.
// The N tile is chosen by the host so that weights + ring fit (conv_ws_plan); CTAs with the same blockIdx.x and different
// N tiles read the same activation boxes at about the same time (L2 hits).
constexpr int kMaxAcc = 4;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <typename T>
struct ParamsWS {
  CUtensorMap tm_a;      // activations (C, Tin, B), box (BK, a_rows, 1)
  CUtensorMap tm_w;      // weights (Cin, Cout, KW), box (BK, BN, 1)
  int kchunks;           // ceil(Cin / BK)
  int in_off;            // first input channel
  int KW, dil, pad;
  int mt;                // M tiles per batch element
  int n_mtiles;          // B * mt
  int a_rows;            // rows of the activation box = 128 + halo
  int a_stage_bytes;     // a_rows * BK * 2 rounded up to 1024
  int sa;                // activation ring depth (>= sub-tiles per unit)
  int w_resident;        // 1: all (chunk, tap) weight tiles of the N tile stay in shared memory; 0: streamed through a ring
  int sw;                // weight ring depth when streamed
  ConvArgs<T> ep;
};

// 256-bit global accesses (sm_100): a thread's 8 fp32 / 16 sixteen-bit channels of a row are one full 32-byte sector
struct F8 {
  float v[8];
};
__device__ __forceinline__ F8 ldg256(const float* p) {
  F8 r;
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg256(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void stg256u(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void ldg256u(const void* p, uint32_t* v) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
}

// residual / MRF-accumulator inputs of 16 channels of one row, requested ahead of their use
struct EpiPre {
  F8 r[2], a[2];
};
template <typename T>
__device__ __forceinline__ void epi_prefetch16(const ConvArgs<T>& a, size_t o, int n8, EpiPre& p) {
  if (a.res32) {
    p.r[0] = ldg256(a.res32 + o);
    if (n8 > 1) p.r[1] = ldg256(a.res32 + o + 8);
  }
  if (a.acc32 && !a.acc_init) {
    p.a[0] = ldg256(a.acc32 + o);
    if (n8 > 1) p.a[1] = ldg256(a.acc32 + o + 8);
  }
}
// conv_epilogue_row8 for 8 * n8 channels (n8 = 1, 2) whose residual / accumulator inputs were prefetched
template <typename T>
__device__ __forceinline__ void epi_apply16(const ConvArgs<T>& a, int b, int t, int co, size_t o, const uint32_t* acc, const EpiPre& p, int n8) {
  float v[16];
  {
    uint32_t bw[8];
    if (n8 > 1) ldg256u(a.bias + co, bw);
    else { const uint4 u = *reinterpret_cast<const uint4*>(a.bias + co); bw[0] = u.x; bw[1] = u.y; bw[2] = u.z; bw[3] = u.w; bw[4] = bw[5] = bw[6] = bw[7] = 0u; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 f = Elem<T>::to_f2(bw[j]);
      v[2 * j] = __uint_as_float(acc[2 * j]) + f.x;
      v[2 * j + 1] = __uint_as_float(acc[2 * j + 1]) + f.y;
    }
  }
  if (a.add) {
    const float* ap = a.add + ((size_t)b * a.add_tg + (a.add_tg > 1 ? t : 0)) * a.add_ld + co;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h < n8) {
        const F8 x = ldg256(ap + 8 * h);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[8 * h + j] += x.v[j];
      }
    }
  }
  if (a.res32) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = p.r[j >> 3].v[j & 7] + a.res_sign * v[j];
  }
  if (a.acc32) {
    if (!a.acc_init) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += p.a[j >> 3].v[j & 7];
    }
    stg256(a.acc32 + o, v);
    if (n8 > 1) stg256(a.acc32 + o + 8, v + 8);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] *= a.acc_scale;
  }
  if (a.mask) {
    const float m = Elem<T>::to_f(a.mask[(size_t)b * a.Tout + t]);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] *= m;
  }
  if (a.out32) {
    stg256(a.out32 + o, v);
    if (n8 > 1) stg256(a.out32 + o + 8, v + 8);
  }
  if (a.outT) {
    const Act act(a.act);
    uint32_t u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] = Elem<T>::from_f2(act(v[2 * j]), act(v[2 * j + 1]));
    if (n8 > 1 && ((a.o_ld | a.o_off) & 15) == 0) stg256u(a.outT + o, u);   // 16-bit rows of 24 channels are only 16-byte aligned
    else {
      *reinterpret_cast<uint4*>(a.outT + o) = make_uint4(u[0], u[1], u[2], u[3]);
      if (n8 > 1) *reinterpret_cast<uint4*>(a.outT + o + 8) = make_uint4(u[4], u[5], u[6], u[7]);
    }
  }
}

template <int COLS> struct WsTmem {
  static constexpr int cols = COLS <= 32 ? 32 : (COLS <= 64 ? 64 : (COLS <= 128 ? 128 : (COLS <= 256 ? 256 : 512)));
};

__device__ __forceinline__ void l2_prefetch(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// MS consecutive M tiles form one unit: one accumulator hand-over (two mbarrier round trips) per unit, so that narrow tiles
// (16 / 32 channels: a tile is 2-4 K outputs) do not pay the hand-over latency per tile.
template <typename T, int BN, int BK, int NACC, int MS>
__global__ void __launch_bounds__(64 + 128 * NACC, 1) conv_umma_ws_kernel(const __grid_constant__ ParamsWS<T> P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int ROW_BYTES = BK * 2, W_BYTES = BN * ROW_BYTES;
  constexpr int W_STAGE = (W_BYTES + 1023) & ~1023;
  constexpr int ACC_COLS = MS * BN;
  constexpr int G = BN / 16;
  constexpr int SU = (G % 2) ? 2 : 1;
  static_assert(MS % SU == 0, "sub-tiles pair up when a tile has an odd number of 16-channel groups");
  constexpr int TMEM_COLS = WsTmem<ACC_COLS * NACC>::cols;
  static_assert(ACC_COLS * NACC <= 512 && NACC <= kMaxAcc, "accumulators fit tensor memory");
  __shared__ uint64_t a_full[kMaxSA], a_empty[kMaxSA], acc_full[kMaxAcc], acc_empty[kMaxAcc], w_full[kMaxSW], w_empty[kMaxSW];
  __shared__ uint32_t tmem_base_s;
  const uint32_t base_u = smem_u32(smem_raw);
  const uint32_t tiles_w = base_u + ((1024u - (base_u & 1023u)) & 1023u);
  const int n_w = P.kchunks * P.KW;
  const uint32_t tiles_a = tiles_w + (uint32_t)((P.w_resident ? n_w : P.sw) * W_STAGE);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BN;
  const int SA = P.sa;
  const int n_units = (P.n_mtiles + MS - 1) / MS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < NACC; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    for (int s = 0; s < kMaxSW; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tm_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tm_w) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  // Loop order of a unit: channel chunk outermost, then tap, then sub-tile -- a weight tile (chunk, tap) meets the MS
  // activation boxes of its chunk back to back, so when the weights are STREAMED through a ring (they do not fit: 256-channel
  // layers, 11-tap 128-channel layers) every weight byte fetched from L2 serves MS tiles.
  const bool resident = P.w_resident != 0;
  const int SW = P.sw;                       // streamed: ring depth (resident: every (chunk, tap) has its own place)
  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer ----
      if (resident) {                        // weights do not depend on the previous kernel: before the wait
        mbar_expect_tx(&w_full[0], (unsigned)(n_w * W_BYTES));
        for (int i = 0; i < n_w; ++i) {
          const int c = i / P.KW, j = i - c * P.KW;
          tma_load_3d(tiles_w + (uint32_t)(i * W_STAGE), &P.tm_w, &w_full[0], c * BK, n0, j);
        }
      }
      pdl_wait();
      int it = 0, iw = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int nsub = P.n_mtiles - unit * MS < MS ? P.n_mtiles - unit * MS : MS;
        for (int c = 0; c < P.kchunks; ++c) {
          for (int sub = 0; sub < nsub; ++sub, ++it) {
            const int tile = unit * MS + sub;
            const int b = tile / P.mt, q0 = (tile - b * P.mt) * BM;
            const int s = it % SA;
            if (it >= SA) mbar_wait(&a_empty[s], ((it / SA) - 1) & 1);
            mbar_expect_tx(&a_full[s], (unsigned)(P.a_rows * ROW_BYTES));
            tma_load_3d(tiles_a + (uint32_t)(s * P.a_stage_bytes), &P.tm_a, &a_full[s], P.in_off + c * BK, q0 - P.pad, b);
          }
          if (!resident) {
            for (int j = 0; j < P.KW; ++j, ++iw) {
              const int w = iw % SW;
              if (iw >= SW) mbar_wait(&w_empty[w], ((iw / SW) - 1) & 1);
              mbar_expect_tx(&w_full[w], W_BYTES);
              tma_load_3d(tiles_w + (uint32_t)(w * W_STAGE), &P.tm_w, &w_full[w], c * BK, n0, j);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- MMA issuer ----
      constexpr uint32_t idesc = (1u << 4) | (UmmaFmt<T>::v << 7) | (UmmaFmt<T>::v << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(BM >> 4) << 24);
      if (resident) mbar_wait(&w_full[0], 0);
      int it = 0, iw = 0, lt = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++lt) {
        const int buf = lt % NACC;
        const int nsub = P.n_mtiles - unit * MS < MS ? P.n_mtiles - unit * MS : MS;
        if (lt >= NACC) mbar_wait(&acc_empty[buf], ((lt / NACC) - 1) & 1);
        tc_fence_after();
        const uint32_t d0 = tmem_d + (uint32_t)(buf * ACC_COLS);
        for (int c = 0; c < P.kchunks; ++c) {
          uint32_t a_base[MS];
          int a_slot[MS];
#pragma unroll
          for (int sub = 0; sub < MS; ++sub) {
            if (sub < nsub) {
              a_slot[sub] = (it + sub) % SA;
              a_base[sub] = tiles_a + (uint32_t)(a_slot[sub] * P.a_stage_bytes);
              mbar_wait(&a_full[a_slot[sub]], ((it + sub) / SA) & 1);
            }
          }
          tc_fence_after();
          uint32_t tap_off = 0;
          for (int j = 0; j < P.KW; ++j, tap_off += (uint32_t)(P.dil * ROW_BYTES)) {
            int w = c * P.KW + j;
            if (!resident) {
              w = iw % SW;
              mbar_wait(&w_full[w], (iw / SW) & 1);
              tc_fence_after();
            }
            const uint64_t bd = smem_desc<BK>(tiles_w + (uint32_t)(w * W_STAGE));
            const uint32_t acc = (c > 0 || j > 0) ? 1u : 0u;
#pragma unroll
            for (int sub = 0; sub < MS; ++sub) {
              if (sub < nsub) {
                const uint64_t ad = smem_desc<BK>(a_base[sub] + tap_off);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                  tc_mma(d0 + (uint32_t)(sub * BN), ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, k > 0 ? 1u : acc);
              }
            }
            if (!resident) {
              tc_commit(&w_empty[w]);        // frees the weight slot when these MMAs have read it
              ++iw;
            }
          }
#pragma unroll
          for (int sub = 0; sub < MS; ++sub)
            if (sub < nsub) tc_commit(&a_empty[a_slot[sub]]);    // frees the activation slots of this chunk
          it += nsub;
        }
        tc_commit(&acc_full[buf]);           // accumulators of this unit complete
      }
      pdl_launch_dependents();
    }
  } else {
    // ---- epilogue: set = which accumulator; warp owns TMEM lanes [32*(warp%4), +32) = rows q0 + 32*(warp%4) + lane ----
    const int set = (warp - 2) >> 2, quarter = warp & 3;
    const ConvArgs<T>& ep = P.ep;
    const int n_valid = ep.Cout - n0 < BN ? ep.Cout - n0 : BN;     // channels of this N tile that exist
    pdl_wait();                              // the epilogue reads residual / accumulator streams of earlier kernels
    // residual / MRF-accumulator rows of a unit, requested into L2 one unit ahead of their use
    auto l2_rows = [&](int unit) {
      if (!ep.res32 && !(ep.acc32 && !ep.acc_init)) return;
      for (int sub = 0; sub < MS; ++sub) {
        const int tile = unit * MS + sub;
        if (tile >= P.n_mtiles) break;
        const int b = tile / P.mt, t = (tile - b * P.mt) * BM + quarter * 32 + lane;
        if (t >= ep.Tout) continue;
        const size_t o = ((size_t)b * ep.Tout + t) * ep.o_ld + ep.o_off + n0;
        for (int c = 0; c < n_valid; c += 32) {
          if (ep.res32) l2_prefetch(ep.res32 + o + c);
          if (ep.acc32 && !ep.acc_init) l2_prefetch(ep.acc32 + o + c);
        }
      }
    };
    if (blockIdx.x + set * (int)gridDim.x < n_units) l2_rows(blockIdx.x + set * gridDim.x);
    int use = 0;
    for (int lt = set;; lt += NACC, ++use) {
      const int unit = blockIdx.x + lt * gridDim.x;
      if (unit >= n_units) break;
      if (unit + NACC * (int)gridDim.x < n_units) l2_rows(unit + NACC * gridDim.x);
      // row of this thread in sub-tile `sub` of the unit: batch element, time step, validity, output offset
      const int tile0 = unit * MS, b0 = tile0 / P.mt, m0 = tile0 - b0 * P.mt;      // one division per unit
      auto row_of = [&](int sub, int& b, int& t, size_t& o) -> bool {
        int m = m0 + sub;
        b = b0;
        while (m >= P.mt) { m -= P.mt; ++b; }
        t = m * BM + quarter * 32 + lane;
        o = ((size_t)b * ep.Tout + t) * ep.o_ld + ep.o_off + n0;
        return tile0 + sub < P.n_mtiles && t < ep.Tout;
      };
      EpiPre pre[2];
      {
        int b, t;
        size_t o;
        if (row_of(0, b, t, o)) epi_prefetch16<T>(ep, o, n_valid > 8 ? 2 : 1, pre[0]);
      }
      mbar_wait(&acc_full[set], use & 1);
      tc_fence_after();
      const uint32_t tbase = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * ACC_COLS);
      // SU sub-tiles per iteration = an even number of 16-channel groups, so the two prefetch buffers alternate at compile time
      // while the code stays small (the fully unrolled unit did not fit the instruction cache: "no instruction" stalls)
#pragma unroll 1
      for (int s0 = 0; s0 < MS; s0 += SU) {
#pragma unroll
        for (int u = 0; u < SU * G; ++u) {
          const int sub = s0 + u / G, c0 = (u % G) * 16;
          {
            const int sub1 = (u + 1 < SU * G) ? s0 + (u + 1) / G : s0 + SU, c1 = (u + 1 < SU * G) ? ((u + 1) % G) * 16 : 0;
            int b1, t1;
            size_t o1;
            if (sub1 < MS && c1 < n_valid && row_of(sub1, b1, t1, o1))
              epi_prefetch16<T>(ep, o1 + c1, c1 + 8 < n_valid ? 2 : 1, pre[(u + 1) & 1]);
          }
          uint32_t v[16];
          tc_ld16(tbase + (uint32_t)(sub * BN + c0), v);
          tc_ld_wait();
          int b, t;
          size_t o;
          if (sub < MS && c0 < n_valid && row_of(sub, b, t, o)) epi_apply16<T>(ep, b, t, n0 + c0, o + c0, v, pre[u & 1], c0 + 8 < n_valid ? 2 : 1);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[set]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
  }
}

// ---- one ResBlock unit in one persistent kernel ----------------------------------------------------------------------------
//   xt = lrelu(conv_d(x_act) + b1);  x = conv_1(xt) + b2 + x        (modules.py:190-203, one iteration of the loop)
// The persistent kernel above runs the two convolutions of a unit as two launches: the first one (16-bit in, 16-bit out) is
// bound by accumulator hand-overs, and its output makes a round trip through HBM (4 of the 16 bytes a unit moves per
// value).  Here the intermediate never leaves the SM: a tile computes 128 intermediate rows (MMA 1 over the dilated taps),
// the four E1 warps add the bias, apply the leaky ReLU, zero the rows outside the signal (the second convolution's zero
// padding) and write them as 16-bit K-major rows -- in the shared-memory swizzle the tensor core expects -- into one of two
// staging tiles; MMA 2 reads that tile through row-shifted descriptors (taps of the dilation-1 convolution) and yields
// R = 128 - (KW - 1) output rows, which the eight E2 warps (two accumulator sets) finish like any second convolution
// (bias, fp32 residual, fp32 + activated 16-bit outputs).  Both weight sets stay resident; the MMA thread issues MMA 1 of
// tile i + 1 before MMA 2 of tile i, so E1 of tile i overlaps it.  Tiles advance by R rows; channels C = Cin = Cout <= 64.
template <typename T>
struct ParamsRU {
  CUtensorMap tm_a;      // activated input (C, T, B), box (BK, a_rows, 1)
  CUtensorMap tm_w1;     // first convolution's weights (Cin, Cout, KW), box (BK, C, 1)
  CUtensorMap tm_w2;     // second convolution's
  int KW, dil;           // taps of both; dilation of the first (the second has dilation 1)
  int Tn;                // time steps per batch element
  int R;                 // output rows per tile = 128 - (KW - 1)
  int mt, n_tiles;       // tiles per batch element, B * mt
  int a_rows, a_stage_bytes, sa;
  int a2_bytes;          // one staging tile: (128 + 16) rows of C channels, rounded up to 1024
  const T* bias1;
  int act1;              // activation between the two convolutions
  ConvArgs<T> ep;        // the second convolution's epilogue
};

// BK = channels per K chunk = row width of the operand tiles: C for 64 / 32 / 16 channels; 64 for C = 48 (TMA zero-fills the
// missing input channels, the staging tiles keep theirs at the zero they are initialised with).  MS consecutive tiles form
// one unit (one hand-over chain MMA 1 -> E1 -> MMA 2 -> E2 per unit): with 16 / 32 channels a tile is 2-4 K outputs and
// the chain's fixed latencies, not its work, set the pace.
template <typename T, int C, int BK, int MS>
__global__ void __launch_bounds__(64 + 128 * 3, 1) resunit_umma_kernel(const __grid_constant__ ParamsRU<T> P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int ROW_BYTES = BK * 2, W_BYTES = C * ROW_BYTES;
  constexpr int W_STAGE = (W_BYTES + 1023) & ~1023;
  constexpr int G = C / 16;
  constexpr int UC = MS * C;                               // accumulator columns of one unit
  constexpr int TMEM_COLS = WsTmem<4 * UC>::cols;
  static_assert(4 * UC <= 512, "two first-convolution and two second-convolution accumulator sets fit tensor memory");
  constexpr uint32_t SWZ = ROW_BYTES == 128 ? 7u : (ROW_BYTES == 64 ? 3u : 1u);     // 16-byte chunk index ^= (offset >> 7) & SWZ
  __shared__ uint64_t a_full[kMaxSA], a_empty[kMaxSA], acc1_full[2], acc1_empty[2], a2_full[2], a2_empty[2], acc2_full[2], acc2_empty[2],
      w_full;
  __shared__ uint32_t tmem_base_s;
  const uint32_t base_u = smem_u32(smem_raw);
  const uint32_t tiles_w1 = base_u + ((1024u - (base_u & 1023u)) & 1023u);
  const uint32_t tiles_w2 = tiles_w1 + (uint32_t)(P.KW * W_STAGE);
  const uint32_t tiles_a = tiles_w2 + (uint32_t)(P.KW * W_STAGE);
  const uint32_t tiles_a2 = tiles_a + (uint32_t)(P.sa * P.a_stage_bytes);      // [2 buffers][MS tiles][a2_bytes]
  uint8_t* const a2_ptr = smem_raw + (tiles_a2 - base_u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int SA = P.sa;
  const int p2 = (P.KW - 1) / 2, p1 = (P.KW - 1) * P.dil / 2;
  const int n_units = (P.n_tiles + MS - 1) / MS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc1_full[s], 1); mbar_init(&acc1_empty[s], 4);
      mbar_init(&a2_full[s], 4); mbar_init(&a2_empty[s], 1);
      mbar_init(&acc2_full[s], 1); mbar_init(&acc2_empty[s], 4);
    }
    mbar_init(&w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tm_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tm_w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P.tm_w2) : "memory");
  }
  // rows 128.. of the staging tiles are read by the last taps of MMA 2 (their products only reach discarded output rows):
  // keep them finite
  for (int i = threadIdx.x; i < 2 * MS * P.a2_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(a2_ptr)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  auto n_sub = [&](int unit) { return P.n_tiles - unit * MS < MS ? P.n_tiles - unit * MS : MS; };

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer: both weight sets once, then one activation box per tile ----
      mbar_expect_tx(&w_full, (unsigned)(2 * P.KW * W_BYTES));
      for (int j = 0; j < P.KW; ++j) {
        tma_load_3d(tiles_w1 + (uint32_t)(j * W_STAGE), &P.tm_w1, &w_full, 0, 0, j);
        tma_load_3d(tiles_w2 + (uint32_t)(j * W_STAGE), &P.tm_w2, &w_full, 0, 0, j);
      }
      pdl_wait();
      int it = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int ns = n_sub(unit);
        for (int sub = 0; sub < ns; ++sub, ++it) {
          const int tile = unit * MS + sub;
          const int b = tile / P.mt, m = tile - b * P.mt;
          const int s = it % SA;
          if (it >= SA) mbar_wait(&a_empty[s], ((it / SA) - 1) & 1);
          mbar_expect_tx(&a_full[s], (unsigned)(P.a_rows * ROW_BYTES));
          tma_load_3d(tiles_a + (uint32_t)(s * P.a_stage_bytes), &P.tm_a, &a_full[s], P.ep.in_off, m * P.R - p2 - p1, b);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- MMA issuer: MMA 1 of unit lt, then MMA 2 of unit lt - 1 ----
      constexpr uint32_t idesc = (1u << 4) | (UmmaFmt<T>::v << 7) | (UmmaFmt<T>::v << 10) | ((uint32_t)(C >> 3) << 17) |
                                 ((uint32_t)(BM >> 4) << 24);
      auto mma2 = [&](int u, int ns) {
        const int buf = u & 1;
        mbar_wait(&a2_full[buf], (u >> 1) & 1);
        if (u >= 2) mbar_wait(&acc2_empty[buf], ((u >> 1) - 1) & 1);
        tc_fence_after();
        for (int sub = 0; sub < ns; ++sub) {
          const uint32_t a_base = tiles_a2 + (uint32_t)((buf * MS + sub) * P.a2_bytes);
          const uint32_t d = tmem_d + (uint32_t)(2 * UC + buf * UC + sub * C);
          for (int j = 0; j < P.KW; ++j) {
            const uint64_t ad = smem_desc<BK>(a_base + (uint32_t)(j * ROW_BYTES));
            const uint64_t bd = smem_desc<BK>(tiles_w2 + (uint32_t)(j * W_STAGE));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) tc_mma(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (j > 0 || k > 0) ? 1u : 0u);
          }
        }
        tc_commit(&a2_empty[buf]);
        tc_commit(&acc2_full[buf]);
      };
      mbar_wait(&w_full, 0);
      int lt = 0, it = 0, ns_prev = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++lt) {
        const int buf = lt & 1, ns = n_sub(unit);
        if (lt >= 2) mbar_wait(&acc1_empty[buf], ((lt >> 1) - 1) & 1);
        for (int sub = 0; sub < ns; ++sub, ++it) {
          const int s = it % SA;
          mbar_wait(&a_full[s], (it / SA) & 1);
          tc_fence_after();
          const uint32_t a_base = tiles_a + (uint32_t)(s * P.a_stage_bytes);
          const uint32_t d = tmem_d + (uint32_t)(buf * UC + sub * C);
          for (int j = 0; j < P.KW; ++j) {
            const uint64_t ad = smem_desc<BK>(a_base + (uint32_t)(j * P.dil * ROW_BYTES));
            const uint64_t bd = smem_desc<BK>(tiles_w1 + (uint32_t)(j * W_STAGE));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) tc_mma(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (j > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&a_empty[s]);
        }
        tc_commit(&acc1_full[buf]);
        if (lt >= 1) mma2(lt - 1, ns_prev);
        ns_prev = ns;
      }
      if (lt >= 1) mma2(lt - 1, ns_prev);
      pdl_launch_dependents();
    }
  } else if (warp < 6) {
    // ---- E1: intermediate rows -> bias, leaky ReLU, zero outside the signal, 16-bit, swizzled K-major staging tile ----
    const int quarter = warp & 3, r = quarter * 32 + lane;
    const Act act(P.act1);
    int lt = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++lt) {
      const int buf = lt & 1, ns = n_sub(unit);
      mbar_wait(&acc1_full[buf], (lt >> 1) & 1);
      if (lt >= 2) mbar_wait(&a2_empty[buf], ((lt >> 1) - 1) & 1);      // MMA 2 of unit lt - 2 has read these staging tiles
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < ns; ++sub) {
        const int tile = unit * MS + sub;
        const int b = tile / P.mt, m = tile - b * P.mt;
        const int t_int = m * P.R - p2 + r;
        const bool inside = t_int >= 0 && t_int < P.Tn;
        uint8_t* const tile_ptr = a2_ptr + (size_t)(buf * MS + sub) * P.a2_bytes;
        const uint32_t tbase = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * UC + sub * C);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          uint32_t v[16];
          tc_ld16(tbase + (uint32_t)(g * 16), v);
          tc_ld_wait();
          uint32_t bw[8], u[8];
          ldg256u(P.bias1 + g * 16, bw);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 f = Elem<T>::to_f2(bw[j]);
            const float x0 = act(__uint_as_float(v[2 * j]) + f.x), x1 = act(__uint_as_float(v[2 * j + 1]) + f.y);
            u[j] = inside ? Elem<T>::from_f2(x0, x1) : 0u;
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t off = (uint32_t)(r * ROW_BYTES + (g * 2 + h) * 16);
            off ^= ((off >> 7) & SWZ) << 4;
            *reinterpret_cast<uint4*>(tile_ptr + off) = make_uint4(u[4 * h], u[4 * h + 1], u[4 * h + 2], u[4 * h + 3]);
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&a2_full[buf]); mbar_arrive(&acc1_empty[buf]); }
    }
  } else {
    // ---- E2: the second convolution's epilogue on R output rows per tile; two accumulator sets ----
    const int set = (warp - 6) >> 2, quarter = warp & 3, r = quarter * 32 + lane;
    const ConvArgs<T>& ep = P.ep;
    pdl_wait();
    int use = 0;
    for (int lt = set;; lt += 2, ++use) {
      const int unit = blockIdx.x + lt * gridDim.x;
      if (unit >= n_units) break;
      const int ns = n_sub(unit);
      auto row_of = [&](int sub, int& b, int& t, size_t& o) -> bool {
        const int tile = unit * MS + sub;
        b = tile / P.mt;
        t = (tile - b * P.mt) * P.R + r;
        o = ((size_t)b * ep.Tout + t) * ep.o_ld + ep.o_off;
        return sub < ns && r < P.R && t < P.Tn;
      };
      EpiPre pre[2];
      {
        int b, t;
        size_t o;
        if (row_of(0, b, t, o)) epi_prefetch16<T>(ep, o, 2, pre[0]);
      }
      mbar_wait(&acc2_full[set], use & 1);
      tc_fence_after();
      const uint32_t tbase = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(2 * UC + set * UC);
      // an even number of 16-channel groups per iteration, so that the two prefetch buffers alternate at compile time across
      // iterations (a single-tile unit has one iteration: nothing to alternate across)
      constexpr int SU = (G % 2 == 1 && MS > 1) ? 2 : 1;
      static_assert(MS % SU == 0, "sub-tiles pair up when a tile has an odd number of 16-channel groups");
#pragma unroll 1
      for (int s0 = 0; s0 < MS; s0 += SU) {
        if (s0 >= ns) break;
#pragma unroll
        for (int u = 0; u < SU * G; ++u) {
          const int sub = s0 + u / G, c0 = (u % G) * 16;
          {
            const int sub1 = (u + 1 < SU * G) ? s0 + (u + 1) / G : s0 + SU, c1 = (u + 1 < SU * G) ? ((u + 1) % G) * 16 : 0;
            int b1, t1;
            size_t o1;
            if (sub1 < MS && row_of(sub1, b1, t1, o1)) epi_prefetch16<T>(ep, o1 + c1, 2, pre[(u + 1) & 1]);
          }
          uint32_t v[16];
          tc_ld16(tbase + (uint32_t)(sub * C + c0), v);
          tc_ld_wait();
          int b, t;
          size_t o;
          if (row_of(sub, b, t, o)) epi_apply16<T>(ep, b, t, c0, o + c0, v, pre[u & 1], 2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc2_empty[set]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host: tensor-map encoding through the driver entry point (no link-time libcuda dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// 3-D map over 16-bit elements: dims (d0, d1, d2) with byte strides (s1, s2) for dims 1 and 2; box (b0, b1, 1)
inline int make_map(CUtensorMap* m, bool bf16, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                    uint32_t b0, uint32_t b1) {
  const CUtensorMapSwizzle swz = b0 == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (b0 == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  EncodeTiledFn fn = encode_fn();
  if (!fn) { gsv_set_error("cuTensorMapEncodeTiled entry point not available"); return GSV_ERR_CUDA; }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1, s2};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult rc = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims,
                   strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    gsv_set_error("cuTensorMapEncodeTiled failed (%d): dims %llu,%llu,%llu strides %llu,%llu box %u,%u", (int)rc,
                  (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)s1,
                  (unsigned long long)s2, b0, b1);
    return GSV_ERR_CUDA;
  }
  return GSV_OK;
}

}  // namespace umma
