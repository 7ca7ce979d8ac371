// gpt_cluster_common.cuh -- device helpers shared by the cluster decode kernels (gpt_decode_cl.cu: one sequence per
// cluster; gpt_decode_cl8.cu: eight sequences per cluster; gpt_decode_hx.cu: one sequence over every cluster): mbarrier / bulk-copy / st.async wrappers, the staged
// vector layout, register dot products, transposing-butterfly reductions, LayerNorm pieces, timeline markers.
#pragma once
#include <cstdlib>

#include "gpt_sample.cuh"

#ifndef GSV_CL_RING
#define GSV_CL_RING 10
#endif

namespace {

constexpr int NT = GSV_DECODE_THREADS;   // 512
constexpr int NWARP = NT / 32;           // 16
constexpr int RING = GSV_CL_RING;                // weight units in flight per warp
constexpr int QKV_PER_WARP = 3 * GSV_HEAD_DIM / NWARP;   // 6
constexpr int O_PER_WARP = GSV_HEAD_DIM / NWARP;          // 2   (D/H = 32 rows per CTA)
constexpr int M1_PER_WARP = 4 * GSV_HEAD_DIM / NWARP;     // 8   (F/H = 128 rows per CTA)
constexpr int M2_PER_WARP = GSV_HEAD_DIM / NWARP;         // 2 rows x 4 K-quarters
constexpr int UNITS_PER_LAYER = QKV_PER_WARP + O_PER_WARP + M1_PER_WARP + 4 * M2_PER_WARP;   // 24

__device__ __forceinline__ int split_pos(int k, int K) {
  const int ch = k >> 3, j = k & 7;
  return j < 4 ? ch * 4 + j : (K >> 1) + ch * 4 + (j - 4);
}
__device__ __forceinline__ void l2_prefetch(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned a = smem_u32(bar);
  unsigned done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(a), "r"(parity) : "memory");
  } while (!done);
}
// bulk copy global -> this CTA's shared memory, completing `bytes` on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// one float into the same shared-memory location of CTA `rank`, completing 4 bytes on that CTA's copy of `bar`
__device__ __forceinline__ void st_async(float* local_ptr, uint64_t* local_bar, unsigned rank, float v) {
  unsigned ra, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(local_bar)), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(ra), "r"(__float_as_uint(v)), "r"(rb) : "memory");
}
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store one float into the same shared-memory variable of CTA `rank` of this cluster
__device__ __forceinline__ void st_remote(float* local_ptr, unsigned rank, float v) {
  const unsigned la = (unsigned)__cvta_generic_to_shared(local_ptr);
  unsigned ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}
__device__ __forceinline__ void st_remote_i(int* local_ptr, unsigned rank, int v) {
  const unsigned la = (unsigned)__cvta_generic_to_shared(local_ptr);
  unsigned ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ra), "r"(v) : "memory");
}

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
__device__ __forceinline__ float warp_allsum(float a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  return a;
}
// Transposing butterflies: per-lane partial sums of 8 (2) rows -> lane holds the FULL sum of row
// idx = 4*bit4 + 2*bit3 + bit2 of its lane id (row = bit4), in 9 (5) shuffles instead of 40 (10).
__device__ __forceinline__ float reduce8(const float (&acc)[8], int lane) {
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
__device__ __forceinline__ float reduce2(float a0, float a1v, int lane) {
  const float send = (lane & 16) ? a0 : a1v;
  const float keep = (lane & 16) ? a1v : a0;
  float a = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  a += __shfl_xor_sync(0xffffffffu, a, 8);
  a += __shfl_xor_sync(0xffffffffu, a, 4);
  a += __shfl_xor_sync(0xffffffffu, a, 2);
  a += __shfl_xor_sync(0xffffffffu, a, 1);
  return a;
}
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
template <typename T, int NCH>
__device__ __forceinline__ void load_vec(const T* row, int lane, uint4 (&w)[NCH]) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) w[c] = ld_weight(reinterpret_cast<const uint4*>(row) + c * 32 + lane);
}
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
// write the lane's 8*NCH elements back in split layout (inverse of load_x)
template <int NCH>
__device__ __forceinline__ void store_x(float* xs, int lane, const float (&x)[NCH * 8]) {
  constexpr int K = NCH * 256;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = c * 32 + lane;
    *reinterpret_cast<float4*>(xs + ch * 4) = make_float4(x[c * 8 + 0], x[c * 8 + 1], x[c * 8 + 2], x[c * 8 + 3]);
    *reinterpret_cast<float4*>(xs + (K >> 1) + ch * 4) = make_float4(x[c * 8 + 4], x[c * 8 + 5], x[c * 8 + 6], x[c * 8 + 7]);
  }
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

}  // namespace
