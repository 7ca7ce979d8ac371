// Per-SM ingest rate of cp.async.bulk (global -> shared) as a function of copy size and issuers per CTA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_bw bulk_bw.cu && ./bulk_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned par) {
  unsigned done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(par) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// each of `nw` warps streams `per_warp_bytes` from its own region in copies of `sz` bytes, `depth` in flight
__global__ void k(const char* src, size_t cta_stride, int nw, unsigned sz, int depth, size_t per_warp_bytes, long long* out) {
  extern __shared__ __align__(128) char sm[];
  __shared__ uint64_t bar[32][16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= nw) return;
  char* mine = sm + (size_t)warp * depth * sz;
  const char* g = src + blockIdx.x * cta_stride + (size_t)warp * per_warp_bytes;
  if (lane == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(&bar[warp][i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int n = (int)(per_warp_bytes / sz);
  long long t0 = clock64();
  if (lane == 0) for (int i = 0; i < depth && i < n; ++i) { mbar_expect_tx(&bar[warp][i], sz); bulk_g2s(mine + (size_t)i * sz, g + (size_t)i * sz, sz, &bar[warp][i]); }
  unsigned acc = 0;
  for (int i = 0; i < n; ++i) {
    const int s = i % depth;
    mbar_wait(&bar[warp][s], (i / depth) & 1);
    acc += *reinterpret_cast<unsigned*>(mine + (size_t)s * sz + lane * 4);
    __syncwarp();
    if (lane == 0 && i + depth < n) { mbar_expect_tx(&bar[warp][s], sz); bulk_g2s(mine + (size_t)s * sz, g + (size_t)(i + depth) * sz, sz, &bar[warp][s]); }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678) out[0] = 0;
}

int main() {
  const size_t total = 1ull << 30;
  char* src; cudaMalloc(&src, total); cudaMemset(src, 1, total);
  long long* out; cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int grids[] = {16, 148};
  struct Cfg { int nw; unsigned sz; int depth; } cfgs[] = {{16, 1024, 10}, {16, 2048, 5}, {16, 4096, 3}, {16, 8192, 1}, {4, 8192, 5}, {4, 32768, 1}, {1, 32768, 5}, {2, 32768, 3}, {1, 65536, 3}};
  for (int gi = 0; gi < 2; ++gi)
    for (auto c : cfgs) {
      const size_t per_cta = 6ull << 20, per_warp = per_cta / c.nw / c.sz * c.sz;
      long long h[148];
      for (int rep = 0; rep < 2; ++rep) {
        k<<<grids[gi], c.nw * 32, (size_t)c.nw * c.depth * c.sz>>>(src, per_cta, c.nw, c.sz, c.depth, per_warp, out);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, out, grids[gi] * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < grids[gi]; ++i) mx = h[i] > mx ? h[i] : mx;
      const double sec = (double)mx / (clk * 1e3);
      printf("grid %3d warps %2d copy %6u B depth %2d : %.1f GB/s per SM, %.2f TB/s total (%s)\n", grids[gi], c.nw, c.sz, c.depth,
             per_warp * c.nw / sec / 1e9, per_warp * c.nw * (double)grids[gi] / sec / 1e12, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
