// glue.cu -- the host glue of TTS.infer / infer_stream as device kernels.
//
// The reference runs these between and after the two models, on the critical path of every utterance / chunk:
//   * _viterbi_monotonic (gsv_tts/TTS.py:1744-1797): a Python loop over the T frames of the MRTE attention map, several
//     tiny launches and a host->device copy per frame;
//   * _find_head_threshold_offsets / _find_tail_threshold_offsets (:1630-1662): frame RMS + nonzero() + .item() syncs;
//   * _sola_algorithm (:1612-1628): two conv1d correlations, argmax().item(), slicing and a cross-fade.
// Here each is one or two launches that leave their result on the device (the caller reads one int when it needs it).
#include "gpt_internal.cuh"

namespace {

constexpr int VT = 256;

// ---- monotonic alignment ------------------------------------------------------------------------------------------------
// One CTA.  normal[t][n] = mean over the heads whose arg-max is not the null key (n = N-1), or a fixed prior when every
// head looks at the null key; dp[t][n] = normal[t][n] + max(dp[t-1][n], dp[t-1][n-1]) (ties keep n); back-pointers in
// `step` ([T][N] bytes: 0 = stay, 1 = came from n-1); frames before the first one whose normal row peaks at n = 0 get -1.
__global__ void __launch_bounds__(VT) viterbi_kernel(const float* __restrict__ attn, int H, int T, int N, unsigned char* __restrict__ step,
                                                     float* __restrict__ normal, int* __restrict__ assign) {
  extern __shared__ float sm[];
  float* dp0 = sm;             // [N]
  float* dp1 = sm + N;         // [N]
  __shared__ float red_v[VT / 32];
  __shared__ int red_i[VT / 32];
  __shared__ int s_first_zero, s_arg;
  __shared__ int s_count;
  __shared__ int s_mask[16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // first maximum of v over threads (value, index): larger value wins, ties -> smaller index
  auto block_argmax = [&](float v, int i) -> int {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
      if (v2 > v || (v2 == v && i2 < i)) { v = v2; i = i2; }
    }
    __syncthreads();
    if (lane == 0) { red_v[warp] = v; red_i[warp] = i; }
    __syncthreads();
    if (tid == 0) {
      float bv = red_v[0]; int bi = red_i[0];
      for (int w = 1; w < VT / 32; ++w) if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bi)) { bv = red_v[w]; bi = red_i[w]; }
      s_arg = bi;
    }
    __syncthreads();
    return s_arg;
  };
  if (tid == 0) s_first_zero = -1;
  // pass 1: normal rows and the first frame whose row peaks at index 0
  float dsum = 0.f;
  for (int n = 0; n < N; ++n) dsum += n == N - 1 ? 0.9f / (float)N : (n == 1 ? 1.1f / (float)N : 1.0f / (float)N);
  for (int t = 0; t < T; ++t) {
    for (int h = 0; h < H; ++h) {
      float bv = -3.0e38f; int bi = 0x7fffffff;
      for (int n = tid; n < N; n += VT) {
        const float v = attn[((size_t)h * T + t) * N + n];
        if (v > bv) { bv = v; bi = n; }
      }
      const int am = block_argmax(bv, bi);
      if (tid == 0) s_mask[h] = am != N - 1 ? 1 : 0;
    }
    __syncthreads();
    if (tid == 0) { int c = 0; for (int h = 0; h < H; ++h) c += s_mask[h]; s_count = c; }
    __syncthreads();
    const int cnt = s_count;
    float bv = -3.0e38f; int bi = 0x7fffffff;
    for (int n = tid; n < N; n += VT) {
      float v;
      if (cnt > 0) {
        float s = 0.f;
        for (int h = 0; h < H; ++h) s += attn[((size_t)h * T + t) * N + n] * (float)s_mask[h];
        v = s / ((float)cnt + 1e-9f);
      } else {
        v = (n == N - 1 ? 0.9f / (float)N : (n == 1 ? 1.1f / (float)N : 1.0f / (float)N)) / dsum;
      }
      normal[(size_t)t * N + n] = v;
      if (v > bv) { bv = v; bi = n; }
    }
    const int am = block_argmax(bv, bi);
    if (tid == 0 && am == 0 && s_first_zero < 0) s_first_zero = t;
    __syncthreads();
  }
  // pass 2: dynamic programme over the frames
  for (int n = tid; n < N; n += VT) dp0[n] = normal[n];
  __syncthreads();
  float* prev = dp0;
  float* cur = dp1;
  for (int t = 1; t < T; ++t) {
    for (int n = tid; n < N; n += VT) {
      const float a = prev[n], b = n > 0 ? prev[n - 1] : -__int_as_float(0x7f800000);
      const bool from_left = b > a;
      cur[n] = normal[(size_t)t * N + n] + (from_left ? b : a);
      step[(size_t)t * N + n] = from_left ? 1 : 0;
    }
    __syncthreads();
    float* tmp = prev; prev = cur; cur = tmp;
  }
  float bv = -3.0e38f; int bi = 0x7fffffff;
  for (int n = tid; n < N; n += VT) { const float v = prev[n]; if (v > bv) { bv = v; bi = n; } }
  const int last = block_argmax(bv, bi);
  if (tid == 0) {
    int n = last;
    assign[T - 1] = n;
    for (int t = T - 2; t >= 0; --t) { n -= step[(size_t)(t + 1) * N + n]; assign[t] = n; }
    const int fz = s_first_zero < 0 ? 0 : s_first_zero;
    for (int t = 0; t < fz; ++t) assign[t] = -1;
  }
}

// ---- silence offsets: first / last 512-sample frame (hop 256) whose RMS exceeds the threshold ------------------------------
template <typename T>
__global__ void frame_hit_kernel(const T* __restrict__ audio, int n, int frame_length, int hop, float threshold, int* __restrict__ first_last) {
  // one warp per frame; first_last[0] = min hit index (init INT_MAX), first_last[1] = max hit index (init -1)
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n_frames = n >= frame_length ? (n - frame_length) / hop + 1 : 0;
  if (warp >= n_frames) return;
  const T* f = audio + (size_t)warp * hop;
  float s = 0.f;
  for (int i = lane; i < frame_length; i += 32) { const float v = Elem<T>::to_f(f[i]); s = fmaf(v, v, s); }
  s = warp_sum(s);
  if (lane == 0 && sqrtf(s / (float)frame_length) > threshold) {
    atomicMin(first_last, warp);
    atomicMax(first_last + 1, warp);
  }
}
__global__ void offsets_kernel(const int* __restrict__ first_last, int n_search, int hop, int margin, int tail, int* __restrict__ out) {
  if (threadIdx.x != 0) return;
  if (!tail) {
    const int f = first_last[0];
    *out = f == 0x7fffffff ? n_search : max(0, f * hop - margin);          // TTS.py:1638-1643
  } else {
    const int l = first_last[1];
    *out = l < 0 ? n_search : max(1, n_search - l * hop - margin);         // TTS.py:1655-1660
  }
}

// ---- SOLA: best alignment of the new chunk's head against the previous chunk's tail, cross-fade, splice ---------------------
template <typename T>
__global__ void __launch_bounds__(256) sola_corr_kernel(const T* __restrict__ f1, const T* __restrict__ f2, int ov, float* __restrict__ norm) {
  // block k: corr[k] = sum_i f2[k+i] f1[i], energy[k] = sum_i f2[k+i]^2 + 1e-8
  const int k = blockIdx.x;
  float c = 0.f, e = 0.f;
  for (int i = threadIdx.x; i < ov; i += blockDim.x) {
    const float a = Elem<T>::to_f(f2[k + i]), q = Elem<T>::to_f(f1[i]);
    c = fmaf(a, q, c);
    e = fmaf(a, a, e);
  }
  __shared__ float rc[8], re[8];
  c = warp_sum(c); e = warp_sum(e);
  if ((threadIdx.x & 31) == 0) { rc[threadIdx.x >> 5] = c; re[threadIdx.x >> 5] = e; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float cs = 0.f, es = 0.f;
    for (int w = 0; w < 8; ++w) { cs += rc[w]; es += re[w]; }
    norm[k] = cs / sqrtf(es + 1e-8f);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) sola_splice_kernel(const T* __restrict__ f1, const T* __restrict__ f2, int n2, int ov, int n_search,
                                                          const float* __restrict__ norm, int* __restrict__ offset_out, T* __restrict__ out) {
  __shared__ int s_off;
  if (threadIdx.x == 0) {
    // every block recomputes the arg-max (n_search <= a few hundred); first maximum wins
    float bv = norm[0]; int bi = 0;
    for (int k = 1; k < n_search; ++k) if (norm[k] > bv) { bv = norm[k]; bi = k; }
    s_off = bi;
    if (blockIdx.x == 0) *offset_out = bi;
  }
  __syncthreads();
  const int off = s_off, n_out = n2 - off;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += gridDim.x * blockDim.x) {
    float v = Elem<T>::to_f(f2[off + i]);
    if (i < ov) {
      // f1 * (1 - alpha) + f2 * alpha with alpha = linspace(0, 1, ov) in the storage type, each op rounded (TTS.py:1623-1626)
      const float alpha = Elem<T>::to_f(Elem<T>::from_f(ov > 1 ? (float)i / (float)(ov - 1) : 0.f));
      const float one_m = Elem<T>::to_f(Elem<T>::from_f(1.f - alpha));
      const float a = Elem<T>::to_f(Elem<T>::from_f(Elem<T>::to_f(f1[i]) * one_m));
      const float b = Elem<T>::to_f(Elem<T>::from_f(v * alpha));
      v = a + b;
    }
    out[i] = Elem<T>::from_f(v);
  }
}

template <typename T>
int offsets_t(const void* audio, int n, float threshold, int frame_length, int hop, int search_len, int margin, int tail, int* work,
              int* out, cudaStream_t st) {
  const int ns = n < search_len ? n : search_len;
  const T* a = reinterpret_cast<const T*>(audio) + (tail ? n - ns : 0);
  const int init[2] = {0x7fffffff, -1};
  GSV_CUDA(cudaMemcpyAsync(work, init, sizeof(init), cudaMemcpyHostToDevice, st));
  const int n_frames = ns >= frame_length ? (ns - frame_length) / hop + 1 : 0;
  if (n_frames > 0) frame_hit_kernel<T><<<(n_frames + 7) / 8, 256, 0, st>>>(a, ns, frame_length, hop, threshold, work);
  offsets_kernel<<<1, 32, 0, st>>>(work, ns, hop, margin, tail, out);
  GSV_CHECK_LAUNCH();
  return GSV_OK;
}

template <typename T>
int sola_t(const void* f1, const void* f2, int n2, int ov, int search_len, float* work, int* offset_out, void* out, cudaStream_t st) {
  const int n_search = (n2 < ov + search_len ? n2 : ov + search_len) - ov + 1;       // conv1d output length (TTS.py:1614-1616)
  if (n_search < 1) { gsv_set_error("sola: the new chunk (%d samples) is shorter than the overlap (%d)", n2, ov); return GSV_ERR_ARG; }
  sola_corr_kernel<T><<<n_search, 256, 0, st>>>(reinterpret_cast<const T*>(f1), reinterpret_cast<const T*>(f2), ov, work);
  sola_splice_kernel<T><<<64, 256, 0, st>>>(reinterpret_cast<const T*>(f1), reinterpret_cast<const T*>(f2), n2, ov, n_search, work, offset_out,
                                            reinterpret_cast<T*>(out));
  GSV_CHECK_LAUNCH();
  return GSV_OK;
}

}  // namespace

extern "C" int gsv_glue_viterbi_monotonic(const float* dev_attn, int H, int T, int N, void* dev_work, int32_t* dev_assign, void* stream) {
  GSV_ARG(dev_attn && dev_work && dev_assign && H >= 1 && H <= 16 && T >= 1 && N >= 2);
  const size_t smem = (size_t)2 * N * sizeof(float);
  GSV_ARG(smem <= 48 * 1024);
  // work: [T][N] fp32 normal rows, then [T][N] bytes of back-pointers
  float* normal = reinterpret_cast<float*>(dev_work);
  unsigned char* step = reinterpret_cast<unsigned char*>(normal + (size_t)T * N);
  viterbi_kernel<<<1, VT, smem, (cudaStream_t)stream>>>(dev_attn, H, T, N, step, normal, dev_assign);
  GSV_CHECK_LAUNCH();
  return GSV_OK;
}

extern "C" int gsv_glue_silence_offset(const void* dev_audio, int n, int dtype, int tail, float threshold, int frame_length, int hop_length,
                                       int search_len, int margin, void* dev_work, int32_t* dev_offset, void* stream) {
  GSV_ARG(dev_audio && dev_work && dev_offset && n >= 1 && frame_length >= 1 && hop_length >= 1 && search_len >= 1);
  if (dtype == GSV_F16)
    return offsets_t<__half>(dev_audio, n, threshold, frame_length, hop_length, search_len, margin, tail, reinterpret_cast<int*>(dev_work),
                             dev_offset, (cudaStream_t)stream);
  return offsets_t<__nv_bfloat16>(dev_audio, n, threshold, frame_length, hop_length, search_len, margin, tail, reinterpret_cast<int*>(dev_work),
                                  dev_offset, (cudaStream_t)stream);
}

extern "C" int gsv_glue_sola(const void* dev_f1_overlap, const void* dev_f2, int n2, int overlap_len, int search_len, int dtype,
                             void* dev_work, int32_t* dev_offset, void* dev_out, void* stream) {
  GSV_ARG(dev_f1_overlap && dev_f2 && dev_work && dev_offset && dev_out && overlap_len >= 1 && search_len >= 0);
  if (dtype == GSV_F16)
    return sola_t<__half>(dev_f1_overlap, dev_f2, n2, overlap_len, search_len, reinterpret_cast<float*>(dev_work), dev_offset, dev_out,
                          (cudaStream_t)stream);
  return sola_t<__nv_bfloat16>(dev_f1_overlap, dev_f2, n2, overlap_len, search_len, reinterpret_cast<float*>(dev_work), dev_offset, dev_out,
                               (cudaStream_t)stream);
}
