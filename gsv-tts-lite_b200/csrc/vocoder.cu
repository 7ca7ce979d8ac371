// vocoder.cu -- SoVITS reverse flow (4 x WaveNet residual coupling) + HiFi-GAN generator.
//
// Replaces SynthesizerTrn.flow_dec / the CUDA-graph bucket path of decode()
// (reference gsv_tts/GPT_SoVITS/SoVITS/models.py:380-383, 406-425): ResidualCouplingBlock
// (models.py:23-65, modules.py:447-511), WN (modules.py:30-104, commons.py:14-21), Generator
// (models.py:68-132) and ResBlock1 (modules.py:115-203) -- ~504 library launches per call there.
//
// Data layout: every activation is time-major / channels-last, [B][T][C], so channels are the
// contiguous (reduction) axis of every convolution tap: coalesced 16-byte loads now, and the
// K-major operand layout an implicit-GEMM tensor-core kernel needs next.  Each convolution
// consumes a 16-bit, already-activated copy of its input and carries residual streams in fp32;
// every elementwise op of the reference (bias, conditioning add, mask, leaky-ReLU, residual add,
// MRF average, channel flip) is folded into a convolution epilogue or an index map:
//   * Flip (modules.py:504-511) never moves data: flows at odd flip parity read/write the
//     physical channels through a reversed index.
//   * weight-norm is folded once at load by the host (the reference re-evaluates g*v/||v|| on
//     every call because flow's weight-norm is never removed, Loader.py:73,95).
//   * xs/3 (models.py:127) and the following leaky-ReLU are applied in the epilogue of the last
//     convolution of the third ResBlock.
#include <cuda.h>

#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <vector>
#include <new>
#include <string>
#include <type_traits>

#include "common.cuh"

namespace {

enum { ACT_NONE = 0, ACT_LRELU_01 = 1, ACT_LRELU_001 = 2, ACT_RELU = 3 };

template <typename T>
struct ConvArgs {
  // input: [B][Tin][in_ld] T, channels [in_off, in_off+Cin); in_rev reads them in reversed order
  const T* in;
  int in_ld, in_off, in_rev;
  int B, Tin, Tout, Cin, Cout, KW, dil, stride;   // stride > 1: transposed convolution (upsampling)
  // weights [KW][rows][Cin] T: pointer already at the first output row, w_tap = rows_total*Cin
  const T* w;
  long long w_tap;
  const T* bias;
  // v = conv + bias (+ add[b][tg][co])
  const float* add;
  int add_ld, add_tg;
  // if res32: v = res32 + res_sign * v          (fp32 residual stream)
  const float* res32;
  float res_sign;
  // if acc32: v = (acc_init ? v : acc32 + v); acc32 = v; v *= acc_scale
  float* acc32;
  int acc_init;
  float acc_scale;
  // if mask: v *= mask[b][t]
  const T* mask;
  // if out32: out32 = v ; if outT: outT = T(act(v))
  float* out32;
  T* outT;
  int act;
  // res32/acc32/out32/outT share one channel geometry: row stride o_ld, first channel o_off,
  // o_rev writes channel (Cout-1-co)
  int o_ld, o_off, o_rev;
};

// Every activation here is max(v, slope * v) with 0 <= slope <= 1 (identity 1, leaky-ReLU 0.1 / 0.01, ReLU 0): two
// instructions per value with slope decoded ONCE per epilogue call, instead of a chain of compares on the activation code
// per value (30 % of the stall samples of a persistent-kernel epilogue).  `+ zero`: x + (-0) is x for every x, and ReLU's
// 0 * negative = -0 becomes the +0 that fmaxf(v, 0) gives.
struct Act {
  float slope, zero;
  __device__ __forceinline__ explicit Act(int act)
      : slope(act == ACT_NONE ? 1.f : (act == ACT_LRELU_01 ? 0.1f : (act == ACT_LRELU_001 ? 0.01f : 0.f))),
        zero(act == ACT_RELU ? 0.f : -0.f) {}
  __device__ __forceinline__ float operator()(float v) const { return fmaxf(v, slope * v) + zero; }
};
__device__ __forceinline__ float act_apply(float v, int act) { return Act(act)(v); }

template <typename T>
__device__ __forceinline__ void conv_epilogue(const ConvArgs<T>& a, int b, int t, int co, float v) {
  v += Elem<T>::to_f(a.bias[co]);
  if (a.add) v += a.add[((size_t)b * a.add_tg + (a.add_tg > 1 ? t : 0)) * a.add_ld + co];
  const size_t o = ((size_t)b * a.Tout + t) * a.o_ld + a.o_off + (a.o_rev ? a.Cout - 1 - co : co);
  if (a.res32) v = a.res32[o] + a.res_sign * v;
  if (a.acc32) {
    if (!a.acc_init) v += a.acc32[o];
    a.acc32[o] = v;
    v *= a.acc_scale;
  }
  if (a.mask) v *= Elem<T>::to_f(a.mask[(size_t)b * a.Tout + t]);
  if (a.out32) a.out32[o] = v;
  if (a.outT) a.outT[o] = Elem<T>::from_f(act_apply(v, a.act));
}

// the same epilogue for 8 consecutive output channels [co, co+8) of one row, vectorised when the output
// channel map is not reversed (co multiple of 8; every pointer 16-byte aligned at such a channel)
template <typename T>
__device__ __forceinline__ void conv_epilogue_row8(const ConvArgs<T>& a, int b, int t, int co, const float* acc) {
  if (a.o_rev) {
#pragma unroll
    for (int j = 0; j < 8; ++j) conv_epilogue<T>(a, b, t, co + j, acc[j]);
    return;
  }
  float v[8], tmp[8];
  unpack8<T>(*reinterpret_cast<const uint4*>(a.bias + co), tmp);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = acc[j] + tmp[j];
  if (a.add) {
    const float* ap = a.add + ((size_t)b * a.add_tg + (a.add_tg > 1 ? t : 0)) * a.add_ld + co;
    const float4 x = *reinterpret_cast<const float4*>(ap), y = *reinterpret_cast<const float4*>(ap + 4);
    v[0] += x.x; v[1] += x.y; v[2] += x.z; v[3] += x.w; v[4] += y.x; v[5] += y.y; v[6] += y.z; v[7] += y.w;
  }
  const size_t o = ((size_t)b * a.Tout + t) * a.o_ld + a.o_off + co;
  if (a.res32) {
    const float4 x = *reinterpret_cast<const float4*>(a.res32 + o), y = *reinterpret_cast<const float4*>(a.res32 + o + 4);
    const float rr[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = rr[j] + a.res_sign * v[j];
  }
  if (a.acc32) {
    if (!a.acc_init) {
      const float4 x = *reinterpret_cast<const float4*>(a.acc32 + o), y = *reinterpret_cast<const float4*>(a.acc32 + o + 4);
      v[0] += x.x; v[1] += x.y; v[2] += x.z; v[3] += x.w; v[4] += y.x; v[5] += y.y; v[6] += y.z; v[7] += y.w;
    }
    *reinterpret_cast<float4*>(a.acc32 + o) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(a.acc32 + o + 4) = make_float4(v[4], v[5], v[6], v[7]);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= a.acc_scale;
  }
  if (a.mask) {
    const float m = Elem<T>::to_f(a.mask[(size_t)b * a.Tout + t]);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= m;
  }
  if (a.out32) {
    *reinterpret_cast<float4*>(a.out32 + o) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(a.out32 + o + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (a.outT) {
    const Act act(a.act);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = act(v[j]);
    *reinterpret_cast<uint4*>(a.outT + o) = pack8<T>(v);
  }
}

// load 8 consecutive logical input channels [c, c+8) of row (b, t) as floats (zero outside [0,Tin))
template <typename T>
__device__ __forceinline__ void load_in8(const ConvArgs<T>& a, int b, int t, int c, float* f) {
  if (t < 0 || t >= a.Tin) {
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.f;
    return;
  }
  const T* row = a.in + ((size_t)b * a.Tin + t) * a.in_ld + a.in_off;
  if (!a.in_rev) {
    unpack8<T>(*reinterpret_cast<const uint4*>(row + c), f);
  } else {
    float r[8];
    unpack8<T>(*reinterpret_cast<const uint4*>(row + (a.Cin - 8 - c)), r);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = r[7 - j];
  }
}

// ---- "same" dilated Conv1d, channels-last, CUDA cores --------------------------------------------------
// CTA tile: TT = 4096/CO_T time steps x CO_T output channels; 256 threads, 4x4 outputs each.
template <typename T, int CO_T>
__global__ void __launch_bounds__(256) conv1d_kernel(const ConvArgs<T> a) {
  constexpr int NTX = CO_T / 4;
  constexpr int TT = (256 / NTX) * 4;
  constexpr int HALO_MAX = 64;
  __shared__ float xs[8][TT + HALO_MAX];
  __shared__ __align__(16) float ws[8][CO_T];
  const int tid = threadIdx.x, tx = tid % NTX, ty = tid / NTX;
  const int b = blockIdx.z, t0 = blockIdx.x * TT, co0 = blockIdx.y * CO_T;
  const int halo = (a.KW - 1) * a.dil, pad = halo / 2;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int c0 = 0; c0 < a.Cin; c0 += 8) {
    __syncthreads();
    for (int r = tid; r < TT + halo; r += 256) {
      float f[8];
      load_in8<T>(a, b, t0 - pad + r, c0, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) xs[j][r] = f[j];
    }
    for (int k = 0; k < a.KW; ++k) {
      __syncthreads();
      if (tid < CO_T) {
        float f[8];
        if (co0 + tid < a.Cout) unpack8<T>(ld_weight(a.w + k * a.w_tap + (size_t)(co0 + tid) * a.Cin + c0), f);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) ws[j][tid] = f[j];
      }
      __syncthreads();
      const int off = ty * 4 + k * a.dil;
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
        const float4 w4 = *reinterpret_cast<const float4*>(&ws[ci][tx * 4]);
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float x = xs[ci][off + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(x, wv[j], acc[i][j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty * 4 + i;
    if (t >= a.Tout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co < a.Cout) conv_epilogue<T>(a, b, t, co, acc[i][j]);
    }
  }
}

// ---- ConvTranspose1d(stride s, kernel KW, padding (KW-s)/2), polyphase gather form -------------------------
//   out[t][co] = sum_ci sum_{j == (t+pad) mod s, j < KW, step s} x[(t+pad-j)/s][ci] * w[j][co][ci]
template <typename T, int CO_T>
__global__ void __launch_bounds__(256) convt1d_kernel(const ConvArgs<T> a) {
  constexpr int NTX = CO_T / 4;
  constexpr int TT = (256 / NTX) * 4;     // output samples per CTA
  constexpr int NIN_MAX = TT / 2 + 24;    // input frames needed (stride >= 2, KW <= 16)
  __shared__ float xs[8][NIN_MAX];
  __shared__ __align__(16) float ws[16][8][CO_T];
  const int tid = threadIdx.x, tx = tid % NTX, ty = tid / NTX;
  const int b = blockIdx.z, t0 = blockIdx.x * TT, co0 = blockIdx.y * CO_T;
  const int s = a.stride, pad = (a.KW - s) / 2;
  // input frame range touched by outputs [t0, t0+TT)
  int tin0 = (t0 + pad - (a.KW - 1));
  tin0 = tin0 >= 0 ? tin0 / s : -((-tin0 + s - 1) / s);
  const int tin1 = (t0 + TT - 1 + pad) / s;
  const int nin = tin1 - tin0 + 1;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int c0 = 0; c0 < a.Cin; c0 += 8) {
    __syncthreads();
    for (int r = tid; r < nin; r += 256) {
      float f[8];
      load_in8<T>(a, b, tin0 + r, c0, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) xs[j][r] = f[j];
    }
    for (int i = tid; i < a.KW * CO_T; i += 256) {
      const int k = i / CO_T, co = i - k * CO_T;
      float f[8];
      if (co0 + co < a.Cout) unpack8<T>(ld_weight(a.w + k * a.w_tap + (size_t)(co0 + co) * a.Cin + c0), f);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) ws[k][j][co] = f[j];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 + ty * 4 + i;
      const int r = (t + pad) % s;
      for (int k = r; k < a.KW; k += s) {
        const int ti = (t + pad - k) / s - tin0;     // exact division; frames outside [0,Tin) were zero-filled
        if (ti < 0 || ti >= nin) continue;
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
          const float x = xs[ci][ti];
          const float4 w4 = *reinterpret_cast<const float4*>(&ws[k][ci][tx * 4]);
          acc[i][0] = fmaf(x, w4.x, acc[i][0]);
          acc[i][1] = fmaf(x, w4.y, acc[i][1]);
          acc[i][2] = fmaf(x, w4.z, acc[i][2]);
          acc[i][3] = fmaf(x, w4.w, acc[i][3]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty * 4 + i;
    if (t >= a.Tout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co < a.Cout) conv_epilogue<T>(a, b, t, co, acc[i][j]);
    }
  }
}

// ---- conv_post (C -> 1, k7, no bias) + tanh (models.py:128-130); input already leaky-ReLU(0.01)'d -------------
template <typename T>
__global__ void __launch_bounds__(256) conv_post_kernel(const T* __restrict__ in, const T* __restrict__ w, T* __restrict__ out,
                                                        int B, int Tn, int C) {
  extern __shared__ float wsm[];   // [7][C]
  for (int i = threadIdx.x; i < 7 * C; i += blockDim.x) wsm[i] = Elem<T>::to_f(w[i]);
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * Tn) return;
  const int b = (int)(idx / Tn), t = (int)(idx - (long long)b * Tn);
  float acc = 0.f;
  for (int k = 0; k < 7; ++k) {
    const int ti = t + k - 3;
    if (ti < 0 || ti >= Tn) continue;
    const T* row = in + ((size_t)b * Tn + ti) * C;
    for (int c = 0; c < C; c += 8) {
      float f[8];
      unpack8<T>(*reinterpret_cast<const uint4*>(row + c), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(f[j], wsm[k * C + c + j], acc);
    }
  }
  out[idx] = Elem<T>::from_f(tanhf(acc));
}

// ---- WaveNet gate: u = tanh(a[:H]) * sigmoid(a[H:]) (commons.py:14-21) --------------------------------------------
template <typename T>
__global__ void gate_kernel(const float* __restrict__ a32, T* __restrict__ u, long long rows, int Hc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * Hc) return;
  const long long r = i / Hc;
  const int c = (int)(i - r * Hc);
  const float ta = a32[r * 2 * Hc + c], sa = a32[r * 2 * Hc + Hc + c];
  u[i] = Elem<T>::from_f(tanhf(ta) * (1.f / (1.f + __expf(-sa))));
}

// ---- layout changes at the boundary: torch [B][C][T] <-> time-major [B][T][C] ------------------------------------------
template <typename T>
__global__ void to_time_major_kernel(const T* __restrict__ in, float* __restrict__ out32, T* __restrict__ outT, int C, int Tn) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < Tn) ? Elem<T>::to_f(in[((size_t)b * C + c) * Tn + t]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    if (t < Tn && c < C) {
      const float v = tile[threadIdx.x][i];
      const size_t o = ((size_t)b * Tn + t) * C + c;
      if (out32) out32[o] = v;
      if (outT) outT[o] = Elem<T>::from_f(v);
    }
  }
}
template <typename T>
__global__ void to_channel_major_kernel(const float* __restrict__ in32, T* __restrict__ out, int C, int Tn) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < Tn) ? in32[((size_t)b * Tn + t) * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    if (t < Tn && c < C) out[((size_t)b * C + c) * Tn + t] = Elem<T>::from_f(tile[threadIdx.x][i]);
  }
}

#include "conv_umma.cuh"

// ---- MRF combine: x = (r0 + r1 + r2) / n (models.py:121-127), then the next stage's leaky-ReLU -> 16-bit --------------
template <typename T>
__global__ void mrf_combine_kernel(const float* __restrict__ r0, const float* __restrict__ r1, const float* __restrict__ r2, int n_r,
                                   T* __restrict__ out, long long n8, int act) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const float inv = 1.f / (float)n_r;
  float v[8];
  const float4 a0 = reinterpret_cast<const float4*>(r0)[2 * i], a1 = reinterpret_cast<const float4*>(r0)[2 * i + 1];
  v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
  if (n_r > 1) {
    const float4 b0 = reinterpret_cast<const float4*>(r1)[2 * i], b1 = reinterpret_cast<const float4*>(r1)[2 * i + 1];
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
  if (n_r > 2) {
    const float4 c0 = reinterpret_cast<const float4*>(r2)[2 * i], c1 = reinterpret_cast<const float4*>(r2)[2 * i + 1];
    v[0] += c0.x; v[1] += c0.y; v[2] += c0.z; v[3] += c0.w; v[4] += c1.x; v[5] += c1.y; v[6] += c1.z; v[7] += c1.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = act_apply(v[j] * inv, act);
  reinterpret_cast<uint4*>(out)[i] = pack8<T>(v);
}

// ---- Flip folded into the weights (modules.py:504-511): a coupling layer that sees channel-reversed data is the same
//      layer with its 1x1 `pre` reversed along Cin and its `post` (weight rows and bias) reversed along Cout -------------
template <typename T>
__global__ void flip_rows_or_cols_kernel(const T* __restrict__ in, T* __restrict__ out, int rows, int cols, int flip_cols) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int r = i / cols, c = i - r * cols;
  out[i] = flip_cols ? in[r * cols + (cols - 1 - c)] : in[(rows - 1 - r) * cols + c];
}

struct Weight {
  const void* w;
  const void* b;
};

// one cached pair of tensor maps per convolution call site of a flow_dec pass (re-encoded when the signature changes)
struct MapCacheEntry {
  const void *in, *w;
  int in_ld, Tin, B, Cin, Cout, KW, bn, bk, a_rows;
  long long w_tap;
  alignas(64) CUtensorMap tm_a;
  alignas(64) CUtensorMap tm_w;
};

}  // namespace

struct gsv_voc_ctx {
  gsv_voc_dims dims;
  std::map<std::string, Weight> weights;
  void* scratch;
  size_t scratch_bytes;
  void* zero_bias;       // conv layers without bias
  void* debug_z;
  long long launches;
  std::vector<MapCacheEntry> map_cache;   // indexed by call site order within one flow_dec pass
  size_t op_index;
  int use_umma;                           // GSV_VOC_IMPL=cuda disables the tensor-core path (A/B checks)
  int use_ws;                             // GSV_VOC_WS=0 disables the weight-stationary persistent kernel (A/B checks)
  int use_fuse;                           // GSV_VOC_FUSE=0: the two convolutions of a ResBlock unit as two launches; 2: fused on small grids too (tests)
  cudaStream_t side[2];                   // the three ResBlocks of an MRF stage run as three concurrent chains
  cudaEvent_t ev_fork, ev_join[2];
  int mrf_streams;                        // GSV_VOC_MRF=serial: one chain after the other on the caller's stream
  std::vector<void*> owned;               // library-owned copies of weights (channel-reversed pre / post of odd flows)
  int num_sms;
  // CUDA graphs of the streaming shapes (B = 1, T <= 64: the 50 / 55-frame chunks of infer_stream, ~180 launches each):
  // captured on the second call of a shape over library-owned input / output buffers, replayed afterwards
  struct ChunkGraph {
    int T, Tg, calls, failed;
    cudaGraphExec_t exec;
    void *z, *mask, *ge, *out;
    long long launches;
  };
  std::vector<ChunkGraph> graphs;
  int use_graph;                          // GSV_VOC_GRAPH=0: always launch kernel by kernel
};

namespace {

// ---- tensor-core path (conv_umma.cuh) ---------------------------------------------------------------------
template <typename T>
bool umma_eligible(const ConvArgs<T>& a) {
  return !a.in_rev && a.Cin >= 16 && a.Cin % 8 == 0 && a.in_ld % 8 == 0 && a.Cout % 16 == 0 && a.KW <= 16 &&
         (a.Cin >= 64 || a.Cin == 32 || a.Cin == 16) && (a.KW - 1) * a.dil <= 120 &&
         (reinterpret_cast<uintptr_t>(a.in) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.w) & 15) == 0 && (a.w_tap % 8) == 0;
}

template <typename T, int BN, int BK>
int launch_umma_inst(const umma::Params<T>& P, dim3 grid, size_t smem, cudaStream_t st) {
  static unsigned long long attr_set = 0ull;        // one bit per device: the attribute is per device, not per process
  int dev = 0;
  GSV_CUDA(cudaGetDevice(&dev));
  if (!((attr_set >> (dev & 63)) & 1ull)) {
    GSV_CUDA(cudaFuncSetAttribute(umma::conv_umma_kernel<T, BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  umma::kSmemBudget + 2048));
    attr_set |= 1ull << (dev & 63);
  }
  static int use_pdl = -1;
  if (use_pdl < 0) { const char* e = getenv("GSV_PDL"); use_pdl = (e && e[0] == '0') ? 0 : 1; }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = dim3(umma::kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
  GSV_CUDA(cudaLaunchKernelEx(&cfg, umma::conv_umma_kernel<T, BN, BK>, P));
  return GSV_OK;
}

template <typename T>
int launch_conv_umma(std::vector<MapCacheEntry>& map_cache, int num_sms, const ConvArgs<T>& a, size_t op, cudaStream_t st) {
  const bool transposed = a.stride > 1;
  const int n_phase = transposed ? a.stride : 1;
  const int max_taps = transposed ? (a.KW + a.stride - 1) / a.stride : a.KW;
  const int halo = transposed ? max_taps - 1 : (a.KW - 1) * a.dil;
  const int m_ext = transposed ? a.Tin + (a.KW - 1) / a.stride + 1 : a.Tout;
  const int mt = (m_ext + umma::BM - 1) / umma::BM;
  const int bk = a.Cin >= 64 ? 64 : a.Cin;                 // 64 / 32 / 16 channels per chunk
  // N tile: as wide as divides Cout, narrowed while the grid would leave most SMs idle
  int bn = a.Cout % 128 == 0 ? 128 : (a.Cout % 64 == 0 ? 64 : (a.Cout % 32 == 0 ? 32 : 16));
  const int bn_min = 16;
  while (bn > bn_min && (long long)mt * (a.Cout / bn) * a.B * n_phase < num_sms) bn >>= 1;
  if (bk == 32 && bn > 32) bn = 32;
  if (bk == 16) bn = 16;
  if (a.Cout % bn != 0) { gsv_set_error("conv: Cout=%d not a multiple of the N tile %d", a.Cout, bn); return GSV_ERR_ARG; }
  const int a_rows = umma::BM + halo;
  if (op >= map_cache.size()) map_cache.resize(op + 1);
  MapCacheEntry& e = map_cache[op];
  if (e.in != a.in || e.w != a.w || e.in_ld != a.in_ld || e.Tin != a.Tin || e.B != a.B || e.Cin != a.Cin || e.Cout != a.Cout ||
      e.KW != a.KW || e.bn != bn || e.bk != bk || e.w_tap != a.w_tap || e.a_rows != a_rows) {
    const bool bf16 = std::is_same<T, __nv_bfloat16>::value;
    int rc = umma::make_map(&e.tm_a, bf16, a.in, (uint64_t)a.in_ld, (uint64_t)a.Tin, (uint64_t)a.B, (uint64_t)a.in_ld * 2,
                            (uint64_t)a.Tin * a.in_ld * 2, (uint32_t)bk, (uint32_t)a_rows);
    if (rc) return rc;
    rc = umma::make_map(&e.tm_w, bf16, a.w, (uint64_t)a.Cin, (uint64_t)a.Cout, (uint64_t)a.KW, (uint64_t)a.Cin * 2,
                        (uint64_t)a.w_tap * 2, (uint32_t)bk, (uint32_t)bn);
    if (rc) return rc;
    e.in = a.in; e.w = a.w; e.in_ld = a.in_ld; e.Tin = a.Tin; e.B = a.B; e.Cin = a.Cin; e.Cout = a.Cout; e.KW = a.KW;
    e.bn = bn; e.bk = bk; e.w_tap = a.w_tap; e.a_rows = a_rows;
  }
  umma::Params<T> P;
  P.tm_a = e.tm_a; P.tm_w = e.tm_w;
  P.kchunks = (a.Cin + bk - 1) / bk;
  P.in_off = a.in_off;
  P.KW = a.KW; P.n_phase = n_phase;
  P.dil = a.dil; P.pad = (a.KW - 1) * a.dil / 2;
  P.t_pad = transposed ? (a.KW - a.stride) / 2 : 0;
  P.m_ext = m_ext;
  P.a_rows = a_rows;
  P.a_stage_bytes = (a_rows * bk * 2 + 1023) & ~1023;
  const int w_stage = (bn * bk * 2 + 1023) & ~1023;
  // ring depths: with one tap (a linear) activations and weights advance together, so both rings get the same
  // depth; with many taps per chunk the weight ring paces the pipeline and 3 activation stages are enough
  // large grids: keep a CTA under ~100 KB so that two share an SM (one's epilogue overlaps the other's main loop)
  const long long n_ctas = (long long)mt * (a.Cout / bn) * a.B * n_phase;
  const int budget = n_ctas >= 2LL * num_sms ? 100 * 1024 : umma::kSmemBudget;
  int sa = max_taps == 1 ? budget / (P.a_stage_bytes + w_stage) : (n_ctas >= 2LL * num_sms ? 2 : 3);
  sa = sa > umma::kMaxSA ? umma::kMaxSA : sa;
  P.sa = P.kchunks < sa ? P.kchunks : sa;
  int sw = (budget - P.sa * P.a_stage_bytes) / w_stage;
  if (sw < 2) sw = 2;
  const int n_w = P.kchunks * max_taps;
  sw = sw > umma::kMaxSW ? umma::kMaxSW : sw;
  P.sw = sw > n_w ? n_w : sw;
  P.ep = a;
  const size_t smem = (size_t)P.sa * P.a_stage_bytes + (size_t)P.sw * w_stage + 1024;
  const dim3 grid(mt, a.Cout / bn, a.B * n_phase);
  P.pdl_early = (long long)grid.x * grid.y * grid.z <= num_sms ? 1 : 0;
  int rc = GSV_ERR_ARG;
  if (bk == 64) {
    if (bn == 128) rc = launch_umma_inst<T, 128, 64>(P, grid, smem, st);
    else if (bn == 64) rc = launch_umma_inst<T, 64, 64>(P, grid, smem, st);
    else if (bn == 32) rc = launch_umma_inst<T, 32, 64>(P, grid, smem, st);
    else rc = launch_umma_inst<T, 16, 64>(P, grid, smem, st);
  } else if (bk == 32) {
    if (bn == 32) rc = launch_umma_inst<T, 32, 32>(P, grid, smem, st);
    else rc = launch_umma_inst<T, 16, 32>(P, grid, smem, st);
  } else {
    rc = launch_umma_inst<T, 16, 16>(P, grid, smem, st);
  }
  if (rc) return rc;
  GSV_CHECK_LAUNCH();
  return GSV_OK;
}

// ---- weight-stationary persistent variant (conv_umma.cuh, second kernel) ---------------------------------------------
struct WsPlan {
  int bn, bk, nacc, ms, sa, n_nt, mt, resident, sw;
  size_t smem;
};
constexpr size_t kWsBudget = 222 * 1024;      // dynamic shared memory of the persistent kernel (227 KB per CTA minus barriers)
inline int ws_ms(int bn) { return bn <= 32 ? 4 : 2; }      // M tiles per accumulator hand-over
inline int ws_nacc(int bn) { return bn >= 64 ? 2 : 3; }
// Persistent kernel for a stride-1 convolution: N tile, channels per chunk, ring depths, and whether the N tile's weights
// (all taps) stay in shared memory or are streamed through a ring (conv_umma.cuh).
template <typename T>
bool conv_ws_plan(const ConvArgs<T>& a, int num_sms, bool need_large, WsPlan& pl) {
  // (the epilogue moves 8 fp32 / 16 sixteen-bit channels per access: rows and channel offsets must keep them 32-byte aligned)
  if (a.stride != 1 || a.in_rev || a.o_rev || a.Cin < 16 || a.Cin % 8 || a.in_ld % 8 || a.Cout % 8 || a.KW > 16 || a.o_ld % 8 || a.o_off % 8 ||
      (a.add && a.add_ld % 8) ||
      (a.KW - 1) * a.dil > 120 || (reinterpret_cast<uintptr_t>(a.in) & 15) || (reinterpret_cast<uintptr_t>(a.w) & 15) || (a.w_tap % 8))
    return false;
  const int mt = (a.Tout + umma::BM - 1) / umma::BM;
  const long long n_mtiles = (long long)a.B * mt;
  if (need_large && n_mtiles < 2LL * num_sms) return false;
  int bk0;
  if (a.Cin == 16) bk0 = 16;
  else if (a.Cin <= 32) bk0 = 32;
  else if (a.Cin < 64) bk0 = 64;
  else if (a.Cin % 64 == 0 || a.Cin % 64 > 32) bk0 = 64;
  else bk0 = 32;
  const int a_rows = umma::BM + (a.KW - 1) * a.dil;
  static const int cand[6] = {128, 96, 64, 48, 32, 16};
  // pass 0: weights resident with a ring of two units; pass 1: weights streamed (wide N tiles only: every N tile re-reads the
  // activations); pass 2: weights resident with whatever ring still fits
  for (int pass = 0; pass < 3; ++pass) {
    for (int ci = 0; ci < 6; ++ci) {
      const int bn = cand[ci];
      if (!(a.Cout % bn == 0 || (bn > a.Cout && bn - a.Cout <= 8))) continue;
      if (bn < 64 && bn < a.Cout) break;      // narrow N tiles re-read the activations too often: the one-tile kernel does better
      const int n_nt = (a.Cout + bn - 1) / bn;
      if (n_nt > num_sms) continue;
      const int ms = ws_ms(bn);
      const int bk = bk0;
      if ((bk == 16 && bn != 16) || (bk == 32 && (bn == 128 || bn == 64))) continue;      // instantiated pairs only
      {
        const int kchunks = (a.Cin + bk - 1) / bk;
        const int n_w = kchunks * a.KW;
        const int a_stage = (a_rows * bk * 2 + 1023) & ~1023;
        const int w_stage = (bn * bk * 2 + 1023) & ~1023;
        int sa = 2 * ms * kchunks;
        sa = sa < 3 ? 3 : sa;
        sa = sa > umma::kMaxSA ? umma::kMaxSA : sa;
        const int sa_good = 2 * ms < sa ? 2 * ms : sa, sa_min = ms + 1;
        if (pass != 1) {
          const size_t w_bytes = (size_t)n_w * w_stage;
          const int floor_sa = pass == 0 ? sa_good : sa_min;
          while (sa > floor_sa && w_bytes + (size_t)sa * a_stage + 1024 > kWsBudget) --sa;
          if (w_bytes + (size_t)sa * a_stage + 1024 > kWsBudget) continue;
          pl.resident = 1; pl.sw = n_w;
          pl.smem = w_bytes + (size_t)sa * a_stage + 1024;
        } else {
          if (n_w < 2) continue;
          sa = sa_good > 3 ? sa_good : 3;
          long long room = (long long)kWsBudget - 1024 - (long long)sa * a_stage;
          int sw = (int)(room / w_stage);
          if (sw > umma::kMaxSW) sw = umma::kMaxSW;
          if (sw > n_w) sw = n_w;
          if (sw < 3) continue;
          pl.resident = 0; pl.sw = sw;
          pl.smem = (size_t)sw * w_stage + (size_t)sa * a_stage + 1024;
        }
        pl.bn = bn; pl.bk = bk; pl.nacc = ws_nacc(bn); pl.ms = ms; pl.sa = sa; pl.n_nt = n_nt; pl.mt = mt;
        return true;
      }
    }
  }
  return false;
}

template <typename T, int BN, int BK, int NACC, int MS>
int launch_ws_inst(const umma::ParamsWS<T>& P, dim3 grid, size_t smem, cudaStream_t st) {
  static unsigned long long attr_set = 0ull;        // one bit per device
  int dev = 0;
  GSV_CUDA(cudaGetDevice(&dev));
  if (!((attr_set >> (dev & 63)) & 1ull)) {
    GSV_CUDA(cudaFuncSetAttribute(umma::conv_umma_ws_kernel<T, BN, BK, NACC, MS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kWsBudget));
    attr_set |= 1ull << (dev & 63);
  }
  static int use_pdl = -1;
  if (use_pdl < 0) { const char* e = getenv("GSV_PDL"); use_pdl = (e && e[0] == '0') ? 0 : 1; }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = dim3(64 + 128 * NACC); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
  GSV_CUDA(cudaLaunchKernelEx(&cfg, umma::conv_umma_ws_kernel<T, BN, BK, NACC, MS>, P));
  return GSV_OK;
}

template <typename T>
int launch_conv_ws(std::vector<MapCacheEntry>& map_cache, int num_sms, const ConvArgs<T>& a, const WsPlan& pl, size_t op, cudaStream_t st) {
  const int bn = pl.bn, bk = pl.bk;
  const int a_rows = umma::BM + (a.KW - 1) * a.dil;
  if (op >= map_cache.size()) map_cache.resize(op + 1);
  MapCacheEntry& e = map_cache[op];
  if (e.in != a.in || e.w != a.w || e.in_ld != a.in_ld || e.Tin != a.Tin || e.B != a.B || e.Cin != a.Cin || e.Cout != a.Cout ||
      e.KW != a.KW || e.bn != bn || e.bk != bk || e.w_tap != a.w_tap || e.a_rows != a_rows) {
    const bool bf16 = std::is_same<T, __nv_bfloat16>::value;
    int rc = umma::make_map(&e.tm_a, bf16, a.in, (uint64_t)a.in_ld, (uint64_t)a.Tin, (uint64_t)a.B, (uint64_t)a.in_ld * 2,
                            (uint64_t)a.Tin * a.in_ld * 2, (uint32_t)bk, (uint32_t)a_rows);
    if (rc) return rc;
    rc = umma::make_map(&e.tm_w, bf16, a.w, (uint64_t)a.Cin, (uint64_t)a.Cout, (uint64_t)a.KW, (uint64_t)a.Cin * 2,
                        (uint64_t)a.w_tap * 2, (uint32_t)bk, (uint32_t)bn);
    if (rc) return rc;
    e.in = a.in; e.w = a.w; e.in_ld = a.in_ld; e.Tin = a.Tin; e.B = a.B; e.Cin = a.Cin; e.Cout = a.Cout; e.KW = a.KW;
    e.bn = bn; e.bk = bk; e.w_tap = a.w_tap; e.a_rows = a_rows;
  }
  umma::ParamsWS<T> P;
  P.tm_a = e.tm_a; P.tm_w = e.tm_w;
  P.kchunks = (a.Cin + bk - 1) / bk;
  P.in_off = a.in_off;
  P.KW = a.KW; P.dil = a.dil; P.pad = (a.KW - 1) * a.dil / 2;
  P.mt = pl.mt; P.n_mtiles = a.B * pl.mt;
  P.a_rows = a_rows;
  P.a_stage_bytes = (a_rows * bk * 2 + 1023) & ~1023;
  P.sa = pl.sa;
  P.w_resident = pl.resident; P.sw = pl.sw;
  P.ep = a;
  const int n_units = (P.n_mtiles + pl.ms - 1) / pl.ms;
  int gx = num_sms / pl.n_nt;
  if (gx > n_units) gx = n_units;
  const dim3 grid(gx, pl.n_nt, 1);
  int rc = GSV_ERR_ARG;
#define GSV_WS(BN_, BK_) rc = launch_ws_inst<T, BN_, BK_, (BN_ >= 64 ? 2 : 3), (BN_ <= 32 ? 4 : 2)>(P, grid, pl.smem, st)
  if (bk == 64) {
    if (bn == 128) GSV_WS(128, 64);
    else if (bn == 96) GSV_WS(96, 64);
    else if (bn == 64) GSV_WS(64, 64);
    else if (bn == 48) GSV_WS(48, 64);
    else if (bn == 32) GSV_WS(32, 64);
    else GSV_WS(16, 64);
  } else if (bk == 32) {
    if (bn == 96) GSV_WS(96, 32);
    else if (bn == 48) GSV_WS(48, 32);
    else if (bn == 32) GSV_WS(32, 32);
    else GSV_WS(16, 32);
  } else {
    GSV_WS(16, 16);
  }
#undef GSV_WS
  if (rc) return rc;
  GSV_CHECK_LAUNCH();
  return GSV_OK;
}

// ---- one ResBlock unit (two convolutions) in one persistent kernel (conv_umma.cuh, resunit_umma_kernel) ---------------------
template <typename T, int C, int BK, int MS>
int launch_ru_inst(const umma::ParamsRU<T>& P, int grid, size_t smem, cudaStream_t st) {
  static unsigned long long attr_set = 0ull;        // one bit per device
  int dev = 0;
  GSV_CUDA(cudaGetDevice(&dev));
  if (!((attr_set >> (dev & 63)) & 1ull)) {
    GSV_CUDA(cudaFuncSetAttribute(umma::resunit_umma_kernel<T, C, BK, MS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWsBudget));
    attr_set |= 1ull << (dev & 63);
  }
  static int use_pdl = -1;
  if (use_pdl < 0) { const char* e = getenv("GSV_PDL"); use_pdl = (e && e[0] == '0') ? 0 : 1; }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64 + 128 * 3); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
  GSV_CUDA(cudaLaunchKernelEx(&cfg, umma::resunit_umma_kernel<T, C, BK, MS>, P));
  return GSV_OK;
}

// a1: xt = act(conv_d(in) + b1) -> 16-bit; a2: x = conv_1(xt) + b2 (+ residual ...), reading a1's output.  Returns GSV_OK and
// sets *done when the pair was launched as one kernel; leaves *done = false when the shapes do not qualify (the caller then
// launches the two convolutions).  a2.outT must not alias a1.in (tiles read their neighbours' input rows).
template <typename T>
int launch_resunit(gsv_voc_ctx* ctx, const ConvArgs<T>& a1, const ConvArgs<T>& a2, cudaStream_t st, bool* done) {
  *done = false;
  const int C = a1.Cin, KW = a1.KW;
  if (!ctx->use_umma || !ctx->use_fuse || !(C == 64 || C == 48 || C == 32 || C == 16) || a1.Cout != C || a2.Cin != C || a2.Cout != C ||
      a2.KW != KW || a2.dil != 1 || a1.stride != 1 || a2.stride != 1 || (KW & 1) == 0 || KW > 15 || a1.in_ld != C || a1.in_off != 0 ||
      a1.in_rev || a1.o_rev || a2.in_rev || a2.o_rev || a2.in != a1.outT || a1.add || a1.res32 || a1.acc32 || a1.mask || a1.out32 ||
      a2.mask || a2.add || a2.o_ld != C || a2.o_off != 0 || a1.Tout != a2.Tout || a1.Tin != a1.Tout || (KW - 1) * a1.dil > 120 ||
      (const void*)a2.outT == (const void*)a1.in)
    return GSV_OK;
  const int Tn = a1.Tout, R = umma::BM - (KW - 1);
  const int mt = (Tn + R - 1) / R;
  const long long n_tiles = (long long)a1.B * mt;
  if (ctx->use_fuse != 2 && n_tiles < 2LL * ctx->num_sms) return GSV_OK;
  const int bk = C == 48 ? 64 : C;          // channels per K chunk = row width of the operand tiles
  const int a_rows = umma::BM + (KW - 1) * a1.dil;
  const int a_stage = (a_rows * bk * 2 + 1023) & ~1023;
  const int w_stage = (C * bk * 2 + 1023) & ~1023;
  const int a2_bytes = ((umma::BM + 16) * bk * 2 + 1023) & ~1023;
  const int ms = C <= 32 ? 4 : 1;           // tiles per hand-over chain (resunit_umma_kernel)
  const size_t fixed = (size_t)2 * KW * w_stage + (size_t)2 * ms * a2_bytes + 1024;
  int sa = 2 * ms < 4 ? 4 : 2 * ms;
  while (sa > ms + 1 && sa > 2 && fixed + (size_t)sa * a_stage > kWsBudget) --sa;
  if (fixed + (size_t)sa * a_stage > kWsBudget) return GSV_OK;
  umma::ParamsRU<T> P;
  const bool bf16 = std::is_same<T, __nv_bfloat16>::value;
  int rc = umma::make_map(&P.tm_a, bf16, a1.in, (uint64_t)C, (uint64_t)Tn, (uint64_t)a1.B, (uint64_t)C * 2, (uint64_t)Tn * C * 2,
                          (uint32_t)bk, (uint32_t)a_rows);
  if (rc) return rc;
  if ((rc = umma::make_map(&P.tm_w1, bf16, a1.w, (uint64_t)C, (uint64_t)C, (uint64_t)KW, (uint64_t)C * 2, (uint64_t)a1.w_tap * 2,
                           (uint32_t)bk, (uint32_t)C)))
    return rc;
  if ((rc = umma::make_map(&P.tm_w2, bf16, a2.w, (uint64_t)C, (uint64_t)C, (uint64_t)KW, (uint64_t)C * 2, (uint64_t)a2.w_tap * 2,
                           (uint32_t)bk, (uint32_t)C)))
    return rc;
  P.KW = KW; P.dil = a1.dil; P.Tn = Tn; P.R = R; P.mt = mt; P.n_tiles = (int)n_tiles;
  P.a_rows = a_rows; P.a_stage_bytes = a_stage; P.sa = sa; P.a2_bytes = a2_bytes;
  P.bias1 = a1.bias; P.act1 = a1.act;
  P.ep = a2;
  const long long n_units = (n_tiles + ms - 1) / ms;
  const int grid = n_units < ctx->num_sms ? (int)n_units : ctx->num_sms;
  const size_t smem = fixed + (size_t)sa * a_stage;
  if (C == 64) rc = launch_ru_inst<T, 64, 64, 1>(P, grid, smem, st);
  else if (C == 48) rc = launch_ru_inst<T, 48, 64, 1>(P, grid, smem, st);
  else if (C == 32) rc = launch_ru_inst<T, 32, 32, 4>(P, grid, smem, st);
  else rc = launch_ru_inst<T, 16, 16, 4>(P, grid, smem, st);
  if (rc) return rc;
  GSV_CHECK_LAUNCH();
  ctx->op_index += 2;                       // the two call sites' tensor-map cache slots stay theirs
  ctx->launches += 1;
  *done = true;
  return GSV_OK;
}

template <typename T>
int launch_conv(gsv_voc_ctx* ctx, const ConvArgs<T>& a, cudaStream_t st) {
  if (a.Cin % 8 != 0 || a.in_off % 8 != 0 || a.in_ld % 8 != 0) {
    gsv_set_error("conv: channel counts must be multiples of 8 (Cin=%d off=%d ld=%d)", a.Cin, a.in_off, a.in_ld);
    return GSV_ERR_ARG;
  }
  {
    const size_t op = ctx->op_index++;
    const bool old_ok = ctx->use_umma && umma_eligible<T>(a);
    WsPlan pl;
    // large grids, and shapes the one-tile kernel has no instance for (48 / 24 / 96 channels of V2ProPlus)
    if (ctx->use_umma && ctx->use_ws && conv_ws_plan<T>(a, ctx->num_sms, old_ok && ctx->use_ws != 2, pl)) {
      ctx->launches += 1;
      return launch_conv_ws<T>(ctx->map_cache, ctx->num_sms, a, pl, op, st);
    }
    if (old_ok) {
      ctx->launches += 1;
      return launch_conv_umma<T>(ctx->map_cache, ctx->num_sms, a, op, st);
    }
  }
  if (a.stride == 1) {
    if ((a.KW - 1) * a.dil > 64) { gsv_set_error("conv: halo too large"); return GSV_ERR_ARG; }
    if (a.Cout > 32) {
      dim3 g((a.Tout + 63) / 64, (a.Cout + 63) / 64, a.B);
      conv1d_kernel<T, 64><<<g, 256, 0, st>>>(a);
    } else if (a.Cout > 16) {
      dim3 g((a.Tout + 127) / 128, 1, a.B);
      conv1d_kernel<T, 32><<<g, 256, 0, st>>>(a);
    } else {
      dim3 g((a.Tout + 255) / 256, 1, a.B);
      conv1d_kernel<T, 16><<<g, 256, 0, st>>>(a);
    }
  } else {
    if (a.KW > 16 || a.stride < 2) { gsv_set_error("convT: kernel %d stride %d unsupported", a.KW, a.stride); return GSV_ERR_ARG; }
    if (a.Cout > 32) {
      dim3 g((a.Tout + 63) / 64, (a.Cout + 63) / 64, a.B);
      convt1d_kernel<T, 64><<<g, 256, 0, st>>>(a);
    } else if (a.Cout > 16) {
      dim3 g((a.Tout + 127) / 128, 1, a.B);
      convt1d_kernel<T, 32><<<g, 256, 0, st>>>(a);
    } else {
      dim3 g((a.Tout + 255) / 256, 1, a.B);
      convt1d_kernel<T, 16><<<g, 256, 0, st>>>(a);
    }
  }
  ctx->launches += 1;
  GSV_CHECK_LAUNCH();
  return GSV_OK;
}

struct Arena {
  char* base;
  size_t off, cap;
  template <typename P> P* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    P* p = reinterpret_cast<P*>(base + off);
    off += n * sizeof(P);
    return p;
  }
};

size_t voc_samples_per_frame(const gsv_voc_ctx* ctx) {
  size_t spf = 1;
  for (int i = 0; i < ctx->dims.n_ups; ++i) spf *= ctx->dims.upsample_rates[i];
  return spf;
}

void voc_drop_graphs(gsv_voc_ctx* ctx) {
  for (auto& g : ctx->graphs) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.z) cudaFree(g.z);
    if (g.mask) cudaFree(g.mask);
    if (g.ge) cudaFree(g.ge);
    if (g.out) cudaFree(g.out);
  }
  ctx->graphs.clear();
}

template <typename T>
int flow_dec_impl(gsv_voc_ctx* ctx, const void* z_p_, const void* mask_, const void* ge_, int B, int Tn, int Tg, void* out_,
                  cudaStream_t st) {
  const gsv_voc_dims& D = ctx->dims;
  const int C = D.inter_channels, Hc = D.hidden_channels, half = C / 2, gin = D.gin_channels, C0 = D.upsample_initial_channel;
  const int NL = D.wn_layers, NF = D.n_flows, NK = D.n_resblock_kernels;
  const T* z_p = reinterpret_cast<const T*>(z_p_);
  const T* mask = reinterpret_cast<const T*>(mask_);
  const T* ge = reinterpret_cast<const T*>(ge_);
  T* out = reinterpret_cast<T*>(out_);
  auto W = [&](const std::string& name, Weight& w) -> int {
    auto it = ctx->weights.find(name);
    if (it == ctx->weights.end()) { gsv_set_error("vocoder weight '%s' was never set", name.c_str()); return GSV_ERR_STATE; }
    w = it->second;
    if (!w.b) w.b = ctx->zero_bias;
    return GSV_OK;
  };

  // ---- scratch carve-up -----------------------------------------------------------------------------
  size_t maxn = 0;   // max over generator stages of (samples per frame * channels)
  {
    size_t spf = 1; int ch = C0;
    for (int i = 0; i < D.n_ups; ++i) { spf *= D.upsample_rates[i]; ch /= 2; maxn = spf * ch > maxn ? spf * ch : maxn; }
  }
  const size_t BT = (size_t)B * Tn;
  float *z32, *h32, *a32, *o32, *condF, *condD, *X0, *XJ[4];
  T *zT, *hT, *uT, *oT, *geT, *preT, *XA0, *XAJ[4], *TA[4], *NEXT;
  auto carve = [&](Arena& ar) {
    z32 = ar.take<float>(BT * C);
    zT = ar.take<T>(BT * C);
    h32 = ar.take<float>(BT * Hc);
    hT = ar.take<T>(BT * Hc);
    a32 = ar.take<float>(BT * 2 * Hc);
    uT = ar.take<T>(BT * Hc);
    o32 = ar.take<float>(BT * Hc);
    oT = ar.take<T>(BT * Hc);
    geT = ar.take<T>((size_t)B * Tg * gin);
    condF = ar.take<float>((size_t)NF * B * Tg * 2 * Hc * NL);
    condD = ar.take<float>((size_t)B * Tg * C0);
    preT = ar.take<T>(BT * C0);
    X0 = ar.take<float>(BT * maxn);
    XA0 = ar.take<T>(BT * maxn);
    for (int j = 0; j < NK; ++j) {     // one residual stream / activated copy / temporary per ResBlock chain
      XJ[j] = ar.take<float>(BT * maxn);
      XAJ[j] = ar.take<T>(BT * maxn);
      TA[j] = ar.take<T>(BT * maxn);
    }
    NEXT = ar.take<T>(BT * maxn);
  };
  Arena dry{nullptr, 0, 0};
  carve(dry);
  const size_t need = dry.off + 256;
  if (need > ctx->scratch_bytes) {
    GSV_CUDA(cudaStreamSynchronize(st));
    voc_drop_graphs(ctx);                 // captured graphs point into the old scratch
    if (ctx->scratch) GSV_CUDA(cudaFree(ctx->scratch));
    ctx->scratch = nullptr; ctx->scratch_bytes = 0;
    GSV_CUDA(cudaMalloc(&ctx->scratch, need));
    ctx->scratch_bytes = need;
  }
  Arena ar{reinterpret_cast<char*>(ctx->scratch), 0, ctx->scratch_bytes};
  carve(ar);

  int rc;
  ctx->op_index = 0;
  const dim3 tb(32, 8);
  // ---- boundary transposes ------------------------------------------------------------------------------
  to_time_major_kernel<T><<<dim3((Tn + 31) / 32, (C + 31) / 32, B), tb, 0, st>>>(z_p, z32, zT, C, Tn);
  to_time_major_kernel<T><<<dim3((Tg + 31) / 32, (gin + 31) / 32, B), tb, 0, st>>>(ge, nullptr, geT, gin, Tg);
  ctx->launches += 2;
  GSV_CHECK_LAUNCH();

  auto base_args = [&]() {
    ConvArgs<T> a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.dil = 1; a.stride = 1; a.KW = 1; a.res_sign = 1.f; a.acc_scale = 1.f;
    return a;
  };

  // ---- conditioning: cond_layer of each flow and dec.cond are 1x1 convs of ge (modules.py:83-84, models.py:116)
  for (int f = 0; f <= NF; ++f) {
    Weight w;
    const bool dec = f == NF;
    if ((rc = W(dec ? std::string("dec.cond") : "flow.flows." + std::to_string(2 * f) + ".enc.cond_layer", w))) return rc;
    ConvArgs<T> a = base_args();
    a.in = geT; a.in_ld = gin; a.Tin = a.Tout = Tg; a.Cin = gin;
    a.Cout = dec ? C0 : 2 * Hc * NL;
    a.w = reinterpret_cast<const T*>(w.w); a.w_tap = (long long)a.Cout * gin; a.bias = reinterpret_cast<const T*>(w.b);
    a.out32 = dec ? condD : condF + (size_t)f * B * Tg * 2 * Hc * NL;
    a.o_ld = a.Cout;
    if ((rc = launch_conv<T>(ctx, a, st))) return rc;
  }

  // ---- reverse flow: for fi = 3,2,1,0: Flip, then coupling.reverse (models.py:63-64) --------------------------------
  for (int n = 0; n < NF; ++n) {
    const int f = NF - 1 - n;
    const bool odd = ((n + 1) & 1) != 0;          // flips applied so far
    // logical x0 = first half after the flips, logical x1 = second half
    const int x0_off = odd ? half : 0, x1_off = odd ? 0 : half;
    const std::string P = "flow.flows." + std::to_string(2 * f) + ".";
    Weight w;
    // pre: h = (Wpre x0 + b) * mask (modules.py:484)
    if ((rc = W(P + "pre", w))) return rc;
    {
      ConvArgs<T> a = base_args();
      a.in = zT; a.in_ld = C; a.in_off = x0_off; a.in_rev = 0; a.Tin = a.Tout = Tn; a.Cin = half; a.Cout = Hc;   // reversal folded into the weights
      a.w = reinterpret_cast<const T*>(w.w); a.w_tap = (long long)Hc * half; a.bias = reinterpret_cast<const T*>(w.b);
      a.mask = mask; a.out32 = h32; a.outT = hT; a.o_ld = Hc;
      if ((rc = launch_conv<T>(ctx, a, st))) return rc;
    }
    const float* cond = condF + (size_t)f * B * Tg * 2 * Hc * NL;
    for (int l = 0; l < NL; ++l) {
      // in_layer conv5 + cond slice -> gate (modules.py:87-94)
      if ((rc = W(P + "enc.in_layers." + std::to_string(l), w))) return rc;
      {
        ConvArgs<T> a = base_args();
        a.in = hT; a.in_ld = Hc; a.Tin = a.Tout = Tn; a.Cin = Hc; a.Cout = 2 * Hc; a.KW = D.wn_kernel;
        a.w = reinterpret_cast<const T*>(w.w); a.w_tap = (long long)2 * Hc * Hc; a.bias = reinterpret_cast<const T*>(w.b);
        a.add = cond + (size_t)l * 2 * Hc; a.add_ld = 2 * Hc * NL; a.add_tg = Tg;
        a.out32 = a32; a.o_ld = 2 * Hc;
        if ((rc = launch_conv<T>(ctx, a, st))) return rc;
      }
      {
        const long long n_el = (long long)BT * Hc;
        gate_kernel<T><<<(unsigned)((n_el + 255) / 256), 256, 0, st>>>(a32, uT, (long long)BT, Hc);
        ctx->launches += 1;
        GSV_CHECK_LAUNCH();
      }
      // res_skip (modules.py:97-103)
      if ((rc = W(P + "enc.res_skip_layers." + std::to_string(l), w))) return rc;
      const bool last = l == NL - 1;
      const int rows = last ? Hc : 2 * Hc;
      if (!last) {   // residual half: h = (h + r[:H]) * mask
        ConvArgs<T> a = base_args();
        a.in = uT; a.in_ld = Hc; a.Tin = a.Tout = Tn; a.Cin = Hc; a.Cout = Hc;
        a.w = reinterpret_cast<const T*>(w.w); a.w_tap = (long long)rows * Hc; a.bias = reinterpret_cast<const T*>(w.b);
        a.res32 = h32; a.mask = mask; a.out32 = h32; a.outT = hT; a.o_ld = Hc;
        if ((rc = launch_conv<T>(ctx, a, st))) return rc;
      }
      {              // skip half: out += r[H:] (or all of r for the last layer); out*mask feeds post
        ConvArgs<T> a = base_args();
        a.in = uT; a.in_ld = Hc; a.Tin = a.Tout = Tn; a.Cin = Hc; a.Cout = Hc;
        a.w = reinterpret_cast<const T*>(w.w) + (last ? 0 : (size_t)Hc * Hc); a.w_tap = (long long)rows * Hc;
        a.bias = reinterpret_cast<const T*>(w.b) + (last ? 0 : Hc);
        a.acc32 = o32; a.acc_init = l == 0; a.o_ld = Hc;
        if (last) { a.mask = mask; a.outT = oT; }
        if ((rc = launch_conv<T>(ctx, a, st))) return rc;
      }
    }
    // post + coupling: x1 = (x1 - (Wpost out + b) * mask) * mask (modules.py:486, 499); mask is 0/1
    if ((rc = W(P + "post", w))) return rc;
    {
      ConvArgs<T> a = base_args();
      a.in = oT; a.in_ld = Hc; a.Tin = a.Tout = Tn; a.Cin = Hc; a.Cout = half;
      a.w = reinterpret_cast<const T*>(w.w); a.w_tap = (long long)half * Hc; a.bias = reinterpret_cast<const T*>(w.b);
      a.res32 = z32; a.res_sign = -1.f; a.mask = mask; a.out32 = z32; a.outT = zT;
      a.o_ld = C; a.o_off = x1_off; a.o_rev = 0;
      if ((rc = launch_conv<T>(ctx, a, st))) return rc;
    }
  }
  if (ctx->debug_z) {
    to_channel_major_kernel<T><<<dim3((Tn + 31) / 32, (C + 31) / 32, B), tb, 0, st>>>(z32, reinterpret_cast<T*>(ctx->debug_z), C, Tn);
    ctx->launches += 1;
    GSV_CHECK_LAUNCH();
  }

  // ---- generator (models.py:113-132) ----------------------------------------------------------------------
  Weight w;
  if ((rc = W("dec.conv_pre", w))) return rc;
  {
    // dec(z * mask): re-mask z on the way in (a no-op after >= 2 flows, kept for n_flows < 2)
    ConvArgs<T> a = base_args();
    a.in = zT; a.in_ld = C; a.Tin = a.Tout = Tn; a.Cin = C; a.Cout = C0; a.KW = 7;
    a.w = reinterpret_cast<const T*>(w.w); a.w_tap = (long long)C0 * C; a.bias = reinterpret_cast<const T*>(w.b);
    a.add = condD; a.add_ld = C0; a.add_tg = Tg;
    a.outT = preT; a.act = ACT_LRELU_01; a.o_ld = C0;
    if ((rc = launch_conv<T>(ctx, a, st))) return rc;
  }
  const T* stage_in = preT;
  int ch = C0, Tcur = Tn;
  for (int i = 0; i < D.n_ups; ++i) {
    const int u = D.upsample_rates[i], ku = D.upsample_kernel_sizes[i];
    const int cin = ch, cout = ch / 2, Tnext = Tcur * u;
    if ((rc = W("dec.ups." + std::to_string(i), w))) return rc;
    {
      ConvArgs<T> a = base_args();
      a.in = stage_in; a.in_ld = cin; a.Tin = Tcur; a.Tout = Tnext; a.Cin = cin; a.Cout = cout; a.KW = ku; a.stride = u;
      a.w = reinterpret_cast<const T*>(w.w); a.w_tap = (long long)cout * cin; a.bias = reinterpret_cast<const T*>(w.b);
      a.out32 = X0; a.outT = XA0; a.act = ACT_LRELU_01; a.o_ld = cout;
      if ((rc = launch_conv<T>(ctx, a, st))) return rc;
    }
    ch = cout; Tcur = Tnext;
    // MRF: the NK ResBlocks read the same x and are independent until their mean: chain j runs on its own stream
    // (chain 0 on the caller's), so the critical path of a stage is 6 convolutions instead of 18
    const bool fork = ctx->mrf_streams && NK >= 2 && NK <= 3;
    if (fork) {
      GSV_CUDA(cudaEventRecord(ctx->ev_fork, st));
      for (int j = 1; j < NK; ++j) GSV_CUDA(cudaStreamWaitEvent(ctx->side[j - 1], ctx->ev_fork, 0));
    }
    for (int j = 0; j < NK; ++j) {
      const int k = D.resblock_kernel_sizes[j];
      const std::string R = "dec.resblocks." + std::to_string(i * NK + j) + ".";
      cudaStream_t sj = (fork && j > 0) ? ctx->side[j - 1] : st;
      T* cur = XA0;
      for (int c = 0; c < 3; ++c) {
        const int dl = D.resblock_dilations[j][c];
        Weight w1, w2;
        if ((rc = W(R + "convs1." + std::to_string(c), w1))) return rc;
        if ((rc = W(R + "convs2." + std::to_string(c), w2))) return rc;
        // The 16-bit activated copy of x lives in `cur` (the stage input XA0, then TA[j] or XAJ[j]).  A fused unit writes the
        // new copy to the other buffer (its tiles read their neighbours' input rows); the two-launch path needs a buffer for
        // the intermediate and may write the new copy back in place (its second convolution does not read `cur`).
        T* const f1 = cur == TA[j] ? XAJ[j] : TA[j];
        T* const f2 = cur == XA0 ? XAJ[j] : nullptr;                  // a second free buffer exists only while cur is the stage input
        ConvArgs<T> a1 = base_args();                                 // xt = lrelu(conv_d(lrelu(x)))
        a1.in = cur; a1.in_ld = ch; a1.Tin = a1.Tout = Tcur; a1.Cin = ch; a1.Cout = ch; a1.KW = k; a1.dil = dl;
        a1.w = reinterpret_cast<const T*>(w1.w); a1.w_tap = (long long)ch * ch; a1.bias = reinterpret_cast<const T*>(w1.b);
        a1.outT = f1; a1.act = ACT_LRELU_01; a1.o_ld = ch;
        ConvArgs<T> a2 = base_args();                                 // x = conv_1(xt) + x (the last one leaves the ResBlock's output in its fp32 stream)
        a2.in = f1; a2.in_ld = ch; a2.Tin = a2.Tout = Tcur; a2.Cin = ch; a2.Cout = ch; a2.KW = k;
        a2.w = reinterpret_cast<const T*>(w2.w); a2.w_tap = (long long)ch * ch; a2.bias = reinterpret_cast<const T*>(w2.b);
        a2.res32 = c == 0 ? X0 : XJ[j]; a2.o_ld = ch;
        a2.out32 = XJ[j];
        T* next_cur = cur;
        bool fused = false;
        {
          ConvArgs<T> f = a2;                                         // fused: intermediate in shared memory, new copy -> f1
          if (c < 2) { f.outT = f1; f.act = ACT_LRELU_01; }
          if ((rc = launch_resunit<T>(ctx, a1, f, sj, &fused))) return rc;
          if (fused) next_cur = f1;
        }
        if (!fused) {
          if (c < 2) { a2.outT = f2 ? f2 : cur; a2.act = ACT_LRELU_01; next_cur = a2.outT; }
          if ((rc = launch_conv<T>(ctx, a1, sj))) return rc;
          if ((rc = launch_conv<T>(ctx, a2, sj))) return rc;
        }
        cur = next_cur;
      }
      if (fork && j > 0) {
        GSV_CUDA(cudaEventRecord(ctx->ev_join[j - 1], sj));
        GSV_CUDA(cudaStreamWaitEvent(st, ctx->ev_join[j - 1], 0));
      }
    }
    {
      // x = mean of the ResBlock outputs (summed in order 0, 1, 2), then the next stage's leaky-ReLU (0.1) or the final
      // one (default slope 0.01, models.py:128)
      const long long n8 = (long long)BT * (Tcur / Tn) * ch / 8;
      mrf_combine_kernel<T><<<(unsigned)((n8 + 255) / 256), 256, 0, st>>>(XJ[0], NK > 1 ? XJ[1] : nullptr, NK > 2 ? XJ[2] : nullptr, NK,
                                                                         NEXT, n8, (i == D.n_ups - 1) ? ACT_LRELU_001 : ACT_LRELU_01);
      ctx->launches += 1;
      GSV_CHECK_LAUNCH();
    }
    stage_in = NEXT;
    // NEXT is consumed by the next stage's upsampling before it is written again (same stream)
  }
  if ((rc = W("dec.conv_post", w))) return rc;
  {
    const long long n_out = (long long)B * Tcur;
    conv_post_kernel<T><<<(unsigned)((n_out + 255) / 256), 256, 7 * ch * sizeof(float), st>>>(
        stage_in, reinterpret_cast<const T*>(w.w), out, B, Tcur, ch);
    ctx->launches += 1;
    GSV_CHECK_LAUNCH();
  }
  return GSV_OK;
}

}  // namespace

extern "C" int gsv_voc_create(const gsv_voc_dims* dims, gsv_voc_ctx** out) {
  GSV_ARG(dims && out);
  GSV_ARG(dims->inter_channels % 16 == 0 && dims->hidden_channels % 8 == 0 && dims->gin_channels % 8 == 0);
  GSV_ARG(dims->n_ups >= 1 && dims->n_ups <= 8 && dims->n_resblock_kernels >= 1 && dims->n_resblock_kernels <= 4);
  GSV_ARG(dims->upsample_initial_channel % (8 << dims->n_ups) == 0);
  GSV_ARG(dims->dtype == GSV_F16 || dims->dtype == GSV_BF16);
  for (int i = 0; i < dims->n_ups; ++i) {
    GSV_ARG(dims->upsample_rates[i] >= 2 && dims->upsample_kernel_sizes[i] <= 16);
    GSV_ARG((dims->upsample_kernel_sizes[i] - dims->upsample_rates[i]) % 2 == 0);
  }
  int dev = 0;
  GSV_CUDA(cudaGetDevice(&dev));
  int rc = gsv_device_check(dev);
  if (rc) return rc;
  gsv_voc_ctx* ctx = new (std::nothrow) gsv_voc_ctx();
  GSV_ARG(ctx != nullptr);
  ctx->dims = *dims;
  ctx->scratch = nullptr; ctx->scratch_bytes = 0; ctx->debug_z = nullptr; ctx->launches = 0;
  ctx->side[0] = ctx->side[1] = nullptr; ctx->ev_fork = nullptr; ctx->ev_join[0] = ctx->ev_join[1] = nullptr;
  ctx->op_index = 0;
  {
    const char* e = getenv("GSV_VOC_IMPL");
    ctx->use_umma = (e && strcmp(e, "cuda") == 0) ? 0 : 1;
    const char* eg = getenv("GSV_VOC_GRAPH");
    ctx->use_graph = (eg && eg[0] == '0') ? 0 : 1;
    const char* ef = getenv("GSV_VOC_FUSE");
    ctx->use_fuse = (ef && ef[0] == '0') ? 0 : ((ef && ef[0] == '2') ? 2 : 1);
    const char* ews = getenv("GSV_VOC_WS");
    ctx->use_ws = (ews && ews[0] == '0') ? 0 : ((ews && ews[0] == '2') ? 2 : 1);   // 2: also on small grids (tests)
  }
  GSV_CUDA(cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, dev));
  {
    const char* e = getenv("GSV_VOC_MRF");
    ctx->mrf_streams = (e && strcmp(e, "serial") == 0) ? 0 : 1;
  }
  for (int i = 0; i < 2; ++i) {
    GSV_CUDA(cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking));
    GSV_CUDA(cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming));
  }
  GSV_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  GSV_CUDA(cudaMalloc(&ctx->zero_bias, 8192));
  GSV_CUDA(cudaMemset(ctx->zero_bias, 0, 8192));
  *out = ctx;
  return GSV_OK;
}

extern "C" int gsv_voc_set_weight(gsv_voc_ctx* ctx, const char* name, const void* dev_weight, const void* dev_bias) {
  GSV_ARG(ctx && name && dev_weight);
  voc_drop_graphs(ctx);                   // captured graphs hold the old pointers
  // flows applied after an odd number of Flips (reverse order: flow f sees NF - f flips): fold the channel reversal
  // into `pre` (along Cin) and `post` (along Cout, bias too), so that no convolution reads or writes reversed channels
  int f2 = -1;
  char tail[16] = "";
  if (sscanf(name, "flow.flows.%d.%15s", &f2, tail) == 2 && (strcmp(tail, "pre") == 0 || strcmp(tail, "post") == 0)) {
    const int f = f2 / 2, NF = ctx->dims.n_flows;
    if (((NF - f) & 1) != 0) {
      const bool pre = strcmp(tail, "pre") == 0;
      const int half = ctx->dims.inter_channels / 2, Hc = ctx->dims.hidden_channels;
      const int rows = pre ? Hc : half, cols = pre ? half : Hc;          // [Cout][Cin], one tap
      const size_t esz = 2;
      void* wcopy = nullptr;
      GSV_CUDA(cudaMalloc(&wcopy, (size_t)rows * cols * esz));
      ctx->owned.push_back(wcopy);
      const int n = rows * cols;
      if (ctx->dims.dtype == GSV_F16)
        flip_rows_or_cols_kernel<__half><<<(n + 255) / 256, 256>>>(reinterpret_cast<const __half*>(dev_weight), reinterpret_cast<__half*>(wcopy), rows, cols, pre ? 1 : 0);
      else
        flip_rows_or_cols_kernel<__nv_bfloat16><<<(n + 255) / 256, 256>>>(reinterpret_cast<const __nv_bfloat16*>(dev_weight), reinterpret_cast<__nv_bfloat16*>(wcopy), rows, cols, pre ? 1 : 0);
      const void* bcopy = dev_bias;
      if (!pre && dev_bias) {
        void* b2 = nullptr;
        GSV_CUDA(cudaMalloc(&b2, (size_t)rows * esz));
        ctx->owned.push_back(b2);
        if (ctx->dims.dtype == GSV_F16)
          flip_rows_or_cols_kernel<__half><<<(rows + 255) / 256, 256>>>(reinterpret_cast<const __half*>(dev_bias), reinterpret_cast<__half*>(b2), rows, 1, 0);
        else
          flip_rows_or_cols_kernel<__nv_bfloat16><<<(rows + 255) / 256, 256>>>(reinterpret_cast<const __nv_bfloat16*>(dev_bias), reinterpret_cast<__nv_bfloat16*>(b2), rows, 1, 0);
        bcopy = b2;
      }
      GSV_CHECK_LAUNCH();
      GSV_CUDA(cudaDeviceSynchronize());
      ctx->weights[std::string(name)] = Weight{wcopy, bcopy};
      return GSV_OK;
    }
  }
  ctx->weights[std::string(name)] = Weight{dev_weight, dev_bias};
  return GSV_OK;
}

extern "C" int gsv_voc_destroy(gsv_voc_ctx* ctx) {
  if (!ctx) return GSV_OK;
  voc_drop_graphs(ctx);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->zero_bias) cudaFree(ctx->zero_bias);
  for (void* q : ctx->owned) cudaFree(q);
  for (int i = 0; i < 2; ++i) {
    if (ctx->side[i]) cudaStreamDestroy(ctx->side[i]);
    if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  delete ctx;
  return GSV_OK;
}

extern "C" int gsv_voc_set_debug_z(gsv_voc_ctx* ctx, void* dev_z) {
  GSV_ARG(ctx);
  ctx->debug_z = dev_z;
  return GSV_OK;
}

extern "C" int64_t gsv_voc_launch_count(gsv_voc_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int gsv_voc_graph_count(gsv_voc_ctx* ctx) {
  int n = 0;
  if (ctx)
    for (auto& g : ctx->graphs) n += g.exec ? 1 : 0;
  return n;
}

static int flow_dec_dispatch(gsv_voc_ctx* ctx, const void* z, const void* mask, const void* ge, int B, int T, int Tg, void* out,
                             cudaStream_t st) {
  if (ctx->dims.dtype == GSV_F16) return flow_dec_impl<__half>(ctx, z, mask, ge, B, T, Tg, out, st);
  return flow_dec_impl<__nv_bfloat16>(ctx, z, mask, ge, B, T, Tg, out, st);
}

extern "C" int gsv_voc_flow_dec(gsv_voc_ctx* ctx, const void* dev_z_p, const void* dev_mask, const void* dev_ge, int B, int T,
                                int Tg, void* dev_out, void* stream) {
  GSV_ARG(ctx && dev_z_p && dev_mask && dev_ge && dev_out);
  GSV_ARG(B >= 1 && T >= 1 && (Tg == 1 || Tg == T));
  cudaStream_t st = (cudaStream_t)stream;
  if (!(ctx->use_graph && B == 1 && T <= 64 && !ctx->debug_z)) return flow_dec_dispatch(ctx, dev_z_p, dev_mask, dev_ge, B, T, Tg, dev_out, st);
  // streaming chunk: first call of a shape runs kernel by kernel (sizes the scratch, encodes the tensor maps), the second is
  // captured over library-owned buffers, later ones copy in, replay, copy out
  gsv_voc_ctx::ChunkGraph* g = nullptr;
  for (auto& c : ctx->graphs)
    if (c.T == T && c.Tg == Tg) g = &c;
  if (!g) {
    ctx->graphs.push_back(gsv_voc_ctx::ChunkGraph{T, Tg, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, 0});
    g = &ctx->graphs.back();
  }
  g->calls += 1;
  if (g->failed || g->calls == 1) return flow_dec_dispatch(ctx, dev_z_p, dev_mask, dev_ge, B, T, Tg, dev_out, st);
  const size_t es = 2, zb = (size_t)ctx->dims.inter_channels * T * es, mb = (size_t)T * es,
               gb = (size_t)ctx->dims.gin_channels * Tg * es, ob = (size_t)T * voc_samples_per_frame(ctx) * es;
  if (!g->exec) {
    const int key_T = g->T, key_Tg = g->Tg;
    void *bz = nullptr, *bm = nullptr, *bg = nullptr, *bo = nullptr;
    if (cudaMalloc(&bz, zb) != cudaSuccess || cudaMalloc(&bm, mb) != cudaSuccess || cudaMalloc(&bg, gb) != cudaSuccess ||
        cudaMalloc(&bo, ob) != cudaSuccess) {
      cudaGetLastError();
      if (bz) cudaFree(bz);
      if (bm) cudaFree(bm);
      if (bg) cudaFree(bg);
      g->failed = 1;
      return flow_dec_dispatch(ctx, dev_z_p, dev_mask, dev_ge, B, T, Tg, dev_out, st);
    }
    g->z = bz; g->mask = bm; g->ge = bg; g->out = bo;
    const long long l0 = ctx->launches;
    cudaGraph_t graph = nullptr;
    // (the legacy default stream cannot be captured: callers on it keep the kernel-by-kernel path)
    cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    bool ok = ce == cudaSuccess;
    int rc = GSV_OK;
    if (ok) {
      rc = flow_dec_dispatch(ctx, bz, bm, bg, 1, T, Tg, bo, st);
      ce = cudaStreamEndCapture(st, &graph);
      ok = ce == cudaSuccess && rc == GSV_OK && graph != nullptr;
    }
    if (!ok && getenv("GSV_VOC_GRAPH_DEBUG"))
      fprintf(stderr, "[gsv] chunk graph T=%d not captured: %s (rc %d: %s)\n", T, cudaGetErrorString(ce), rc, rc ? gsv_last_error() : "");
    // flow_dec_impl may have dropped and re-created the table entry's neighbours only on a scratch growth, which cannot happen
    // here (the first call sized the scratch); look the entry up again all the same
    g = nullptr;
    for (auto& c : ctx->graphs)
      if (c.T == key_T && c.Tg == key_Tg) g = &c;
    if (!g) { if (graph) cudaGraphDestroy(graph); return flow_dec_dispatch(ctx, dev_z_p, dev_mask, dev_ge, B, T, Tg, dev_out, st); }
    cudaGraphExec_t exec = nullptr;
    if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
      cudaGetLastError();
      g->failed = 1;
      ctx->launches = l0;
      return flow_dec_dispatch(ctx, dev_z_p, dev_mask, dev_ge, B, T, Tg, dev_out, st);
    }
    g->exec = exec;
    g->launches = ctx->launches - l0;
    ctx->launches = l0;
  }
  GSV_CUDA(cudaMemcpyAsync(g->z, dev_z_p, zb, cudaMemcpyDeviceToDevice, st));
  GSV_CUDA(cudaMemcpyAsync(g->mask, dev_mask, mb, cudaMemcpyDeviceToDevice, st));
  GSV_CUDA(cudaMemcpyAsync(g->ge, dev_ge, gb, cudaMemcpyDeviceToDevice, st));
  GSV_CUDA(cudaGraphLaunch(g->exec, st));
  GSV_CUDA(cudaMemcpyAsync(dev_out, g->out, ob, cudaMemcpyDeviceToDevice, st));
  ctx->launches += g->launches;
  return GSV_OK;
}

// ---- nn.Linear on the same tensor-core kernel (a convolution with one tap), for the GPT prefill and the
//      batched decode step: out[r][n] = act(sum_k X[r][k] W[n][k] + bias[n]) --------------------------------------
struct gsv_umma_cache {
  std::vector<MapCacheEntry> maps;
  int num_sms;
};

gsv_umma_cache* gsv_umma_cache_create(int num_sms) {
  gsv_umma_cache* c = new (std::nothrow) gsv_umma_cache();
  if (c) c->num_sms = num_sms;
  return c;
}
void gsv_umma_cache_destroy(gsv_umma_cache* c) { delete c; }

template <typename T>
static int umma_linear_t(gsv_umma_cache* c, size_t op, const void* X, int rows, int rows_cap, int K, const void* W, const void* bias,
                         int N, void* out, int relu, cudaStream_t st) {
  ConvArgs<T> a;
  memset(&a, 0, sizeof(a));
  a.in = reinterpret_cast<const T*>(X); a.in_ld = K; a.B = 1; a.Tin = rows_cap; a.Tout = rows; a.Cin = K; a.Cout = N;
  a.KW = 1; a.dil = 1; a.stride = 1; a.res_sign = 1.f; a.acc_scale = 1.f;
  a.w = reinterpret_cast<const T*>(W); a.w_tap = (long long)N * K; a.bias = reinterpret_cast<const T*>(bias);
  a.outT = reinterpret_cast<T*>(out); a.o_ld = N; a.act = relu ? ACT_RELU : ACT_NONE;
  if (!umma_eligible<T>(a)) { gsv_set_error("umma linear: unsupported shape K=%d N=%d", K, N); return GSV_ERR_ARG; }
  return launch_conv_umma<T>(c->maps, c->num_sms, a, op, st);
}

int gsv_umma_linear(gsv_umma_cache* c, size_t op, int dtype, const void* X, int rows, int rows_cap, int K, const void* W,
                    const void* bias, int N, void* out, int relu, cudaStream_t st) {
  if (dtype == GSV_F16) return umma_linear_t<__half>(c, op, X, rows, rows_cap, K, W, bias, N, out, relu, st);
  return umma_linear_t<__nv_bfloat16>(c, op, X, rows, rows_cap, K, W, bias, N, out, relu, st);
}

// Conv1d with KW taps ('same' zero padding, odd KW) over channels-last rows: out[t][n] = act(sum_k X[t + k - KW/2] . W[k][n] + bias[n]),
// X [rows][K] T, W [KW][N][K] T.  The FFN convolutions of the prior encoder (encp.cu; attentions.py:244-271).
template <typename T>
static int umma_conv_t(gsv_umma_cache* c, size_t op, const void* X, int rows, int K, const void* W, const void* bias, int N, int KW,
                       void* out, int relu, cudaStream_t st) {
  ConvArgs<T> a;
  memset(&a, 0, sizeof(a));
  a.in = reinterpret_cast<const T*>(X); a.in_ld = K; a.B = 1; a.Tin = rows; a.Tout = rows; a.Cin = K; a.Cout = N;
  a.KW = KW; a.dil = 1; a.stride = 1; a.res_sign = 1.f; a.acc_scale = 1.f;
  a.w = reinterpret_cast<const T*>(W); a.w_tap = (long long)N * K; a.bias = reinterpret_cast<const T*>(bias);
  a.outT = reinterpret_cast<T*>(out); a.o_ld = N; a.act = relu ? ACT_RELU : ACT_NONE;
  if (!umma_eligible<T>(a) || (KW & 1) == 0) { gsv_set_error("umma conv: unsupported shape K=%d N=%d KW=%d", K, N, KW); return GSV_ERR_ARG; }
  return launch_conv_umma<T>(c->maps, c->num_sms, a, op, st);
}

int gsv_umma_conv(gsv_umma_cache* c, size_t op, int dtype, const void* X, int rows, int K, const void* W, const void* bias, int N, int KW,
                  void* out, int relu, cudaStream_t st) {
  if (dtype == GSV_F16) return umma_conv_t<__half>(c, op, X, rows, K, W, bias, N, KW, out, relu, st);
  return umma_conv_t<__nv_bfloat16>(c, op, X, rows, K, W, bias, N, KW, out, relu, st);
}
