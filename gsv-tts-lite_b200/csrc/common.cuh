// common.cuh -- shared device helpers for libgsv_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gsv_b200.h"

#define GSV_WARP 32

void gsv_set_error(const char* fmt, ...);
#define GSV_CUDA(call)                                                                 \
  do {                                                                                 \
    cudaError_t _e = (call);                                                           \
    if (_e != cudaSuccess) {                                                           \
      gsv_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return GSV_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)
#define GSV_CHECK_LAUNCH() GSV_CUDA(cudaGetLastError())
#define GSV_ARG(cond)                                                      \
  do {                                                                     \
    if (!(cond)) {                                                         \
      gsv_set_error("%s:%d bad argument: %s", __FILE__, __LINE__, #cond);  \
      return GSV_ERR_ARG;                                                  \
    }                                                                      \
  } while (0)

// ---- 16-bit element traits -----------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<__half> {
  using T2 = __half2;
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
  static __device__ __forceinline__ float2 to_f2(uint32_t u) {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
  }
  static __device__ __forceinline__ uint32_t from_f2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};
template <> struct Elem<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float2 to_f2(uint32_t u) {
    // bf16 -> fp32 is a 16-bit shift
    float2 r;
    r.x = __uint_as_float(u << 16);
    r.y = __uint_as_float(u & 0xffff0000u);
    return r;
  }
  static __device__ __forceinline__ uint32_t from_f2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};

// 8 packed 16-bit elements -> 8 floats
template <typename T> __device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = Elem<T>::to_f2(u.x), b = Elem<T>::to_f2(u.y), c = Elem<T>::to_f2(u.z), d = Elem<T>::to_f2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
template <typename T> __device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = Elem<T>::from_f2(f[0], f[1]); u.y = Elem<T>::from_f2(f[2], f[3]);
  u.z = Elem<T>::from_f2(f[4], f[5]); u.w = Elem<T>::from_f2(f[6], f[7]);
  return u;
}

// ---- memory access flavours ----------------------------------------------------------------
// weights: read-only, streamed once per step -> non-coherent path, do not pollute L1
__device__ __forceinline__ uint4 ld_weight(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// data exchanged between CTAs inside one launch: always served from L2 (L1 is not coherent)
__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ uint4 ld_cg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ int ld_cg(const int* p) { return __ldcg(p); }
__device__ __forceinline__ void st_cg(float* p, float v) { __stcg(p, v); }
__device__ __forceinline__ void st_cg(int* p, int v) { __stcg(p, v); }

// ---- warp / block reductions ---------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- grid-wide barrier for the persistent decode kernel ----------------------------------------
// All CTAs are co-resident (cooperative launch).  `counter` is zeroed by the host before the
// launch; `*target` is this thread's private running target (only thread 0's copy is used).
__device__ __forceinline__ void grid_arrive(unsigned* counter) {
  // caller has already done __syncthreads(); executed by one thread
  __threadfence();
  atomicAdd(counter, 1u);
}
__device__ __forceinline__ void grid_wait(unsigned* counter, unsigned target) {
  unsigned v;
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
  } while (v < target);
}
__device__ __forceinline__ void grid_sync(unsigned* counter, unsigned& epoch) {
  __syncthreads();
  epoch += gridDim.x;
  if (threadIdx.x == 0) {
    grid_arrive(counter);
    grid_wait(counter, epoch);
  }
  __syncthreads();
}

// ---- Philox4x32-10 (counter-based RNG for the sampling noise) -----------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
// uniform in (0,1] from 32 random bits, then Exp(1)
__device__ __forceinline__ float exp1_from_bits(uint32_t b) {
  float u = (static_cast<float>(b >> 8) + 1.0f) * (1.0f / 16777216.0f);
  return -__logf(u) + 1e-30f;
}
