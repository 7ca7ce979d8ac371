// gpt_decode_cl8.cu -- cluster decode kernel for up to EIGHT sequences per cluster on the tensor cores.
//
// Same arithmetic as the other decode kernels (reference t2s_model.py:67-105, 129-143, 442-456) and the same
// skeleton as gpt_decode_cl.cu (one CTA per attention head, push exchanges completing on the
// receiver's mbarrier).  What changes: the cluster's live sequences are the N = 8 columns of mma.sync m16n8k16 tiles
// whose M = 16 rows are weight rows, so a weight row that has been streamed from HBM is multiplied with all eight
// inputs by one instruction (a CUDA-core variant with 2 / 4 sequences per cluster paid the full dot product per sequence: 657 / 1017 us per step).
//
// Weights are re-tiled ONCE (first launch, cl8_pack_*_kernel) into the order the kernel consumes them: per (layer,
// head CTA) twelve CHUNKS of 32 rows x D columns (q, k, v rows of the head; its 32 out-proj rows; 4 x 32 MLP-up
// rows; its 32 MLP-down rows, one K-quarter per chunk), each chunk stored in mma A-FRAGMENT order
// [M-tile][k-step][lane][16 B].  Measured reasons (tools/cl8_timeline.py, tools/ubench/bulk_bw.cu): (1) an SM ingests
// 1 KB bulk copies at 20 GB/s but >= 8 KB copies at ~200 GB/s -- a chunk is ONE 32 KB copy into a 3-slot ring shared
// by the CTA; (2) legacy mma.sync issues at about one per 16 cycles per sub-partition, so tiles must use all 16 rows
// -- a row-per-warp mapping that fills 2..4 of them spends 9 us per layer in the tensor pipe; (3) fragment order makes
// the A operand one conflict-free 16-byte shared load per lane and k-step.
// Per chunk the 16 warps split (M-tile, K-eighth); partial accumulators meet in shared memory and warps 0..7 (warp n
// = sequence n, lane = row) finish bias / residual / activation.  The head is 33 more chunks (vocabulary rows padded
// to 32) dealt round-robin to the CTAs.  Activations cross CTAs in the storage type (as the reference rounds them)
// and land in the layout the B operand is read from ([sequence][k], rows padded by 8 elements); the residual stream
// (y1, y2, next input) stays fp32.  LayerNorm of sequence n is done once per CTA by warp n; attention of sequence n
// by warps 2n and 2n+1.
#include <cstdlib>
#include <type_traits>

#include "gpt_cluster_common.cuh"

namespace {

constexpr int NB8 = 8;                   // sequences per cluster
constexpr int NSLOT = 3;                 // chunk ring slots
constexpr int CHUNKS = 12;               // chunks per layer
constexpr int XPAD = 8;                  // padding elements per staged activation row
constexpr int PF = 3;                    // attention: passes (8 cached positions each) of K/V loads in flight per warp
constexpr int RW = 176;                  // floats per warp in the partial-accumulator buffer: [8 sequences][20] (+16: bank shift between M-tiles)

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

// 16 bytes into the same shared-memory location of CTA `rank`, completing that many bytes on its copy of `bar`
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
  __align__(16) float stage[NWARP][NB8][4]; // per-warp staging of the attention output (32 values in the storage type) before it is pushed
  float alive[NB8];                         // pushed by the sampler CTAs together with the next inputs
  int alive_i;
  int slot[NB8], kv[NB8];
  uint64_t cbar[NSLOT];                     // chunk ring: bytes landed
  uint64_t xbar[4];                         // inboxes: 0 inA (fp32: xin / y1 / y2), 1 att, 2 h, 3 logits (CTA n for sequence n)
};

// One thread per 16-byte fragment element: lane (g = lane / 4, t = lane % 4) of k-step ks of M-tile `tile` holds
// A[g][k0..k0+1], A[g+8][k0..], A[g][k0+8..], A[g+8][k0+8..] with k0 = 16 ks + 2 t.
template <typename T>
__global__ void cl8_pack_layers_kernel(const T* __restrict__ Wqkv, const T* __restrict__ Wo, const T* __restrict__ W1,
                                       const T* __restrict__ W2, uint4* __restrict__ out, int L, int H, int D) {
  const size_t total = (size_t)L * H * CHUNKS * 2 * (D / 16) * 32;
  const int F = 4 * D;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int lane = (int)(r & 31); r >>= 5;
    const int ks = (int)(r % (D / 16)); r /= (D / 16);
    const int tile = (int)(r & 1); r >>= 1;
    const int c = (int)(r % CHUNKS); r /= CHUNKS;
    const int rank = (int)(r % H);
    const int l = (int)(r / H);
    const int g = lane >> 2, t = lane & 3, r0 = tile * 16 + g, k0 = ks * 16 + 2 * t;
    const T* base; size_t ld;
    if (c < 3) { base = Wqkv + ((size_t)l * 3 * D + (size_t)c * D + rank * GSV_HEAD_DIM) * D; ld = D; }
    else if (c == 3) { base = Wo + ((size_t)l * D + rank * GSV_HEAD_DIM) * D; ld = D; }
    else if (c < 8) { base = W1 + ((size_t)l * F + rank * (4 * GSV_HEAD_DIM) + (c - 4) * 32) * D; ld = D; }
    else { base = W2 + ((size_t)l * D + rank * GSV_HEAD_DIM) * F + (size_t)(c - 8) * D; ld = F; }
    uint4 v;
    v.x = *reinterpret_cast<const unsigned*>(base + (size_t)r0 * ld + k0);
    v.y = *reinterpret_cast<const unsigned*>(base + (size_t)(r0 + 8) * ld + k0);
    v.z = *reinterpret_cast<const unsigned*>(base + (size_t)r0 * ld + k0 + 8);
    v.w = *reinterpret_cast<const unsigned*>(base + (size_t)(r0 + 8) * ld + k0 + 8);
    out[idx] = v;
  }
}
template <typename T>
__global__ void cl8_pack_head_kernel(const T* __restrict__ Wh, uint4* __restrict__ out, int n_chunks, int V, int D) {
  const size_t total = (size_t)n_chunks * 2 * (D / 16) * 32;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int lane = (int)(r & 31); r >>= 5;
    const int ks = (int)(r % (D / 16)); r /= (D / 16);
    const int tile = (int)(r & 1); r >>= 1;
    const int hc = (int)r;
    const int g = lane >> 2, t = lane & 3, r0 = hc * 32 + tile * 16 + g, k0 = ks * 16 + 2 * t;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r0 < V) {
      v.x = *reinterpret_cast<const unsigned*>(Wh + (size_t)r0 * D + k0);
      v.z = *reinterpret_cast<const unsigned*>(Wh + (size_t)r0 * D + k0 + 8);
    }
    if (r0 + 8 < V) {
      v.y = *reinterpret_cast<const unsigned*>(Wh + (size_t)(r0 + 8) * D + k0);
      v.w = *reinterpret_cast<const unsigned*>(Wh + (size_t)(r0 + 8) * D + k0 + 8);
    }
    out[idx] = v;
  }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(NT, 1) gpt_decode_cl8_kernel(const GptParams p, const int n_steps, const unsigned char* __restrict__ pack,
                                                                const unsigned char* __restrict__ hpack, unsigned* const resident) {
  extern __shared__ __align__(16) float smem[];
  __shared__ Cl8Shared sh;
  constexpr int D = NCH * 256, F = 4 * D;
  constexpr int LDX = D + XPAD, LDH = F + XPAD;          // element strides of the staged activation rows
  constexpr int KSW = D / 128;                            // k-steps of a chunk per warp (K split eight ways)
  constexpr unsigned CHUNK_BYTES = 64u * D;               // 2 M-tiles x D/16 k-steps x 512 B
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;                  // mma fragment coordinates
  const int H = p.H, L = p.L, V = p.V, S = p.S;
  const unsigned rank = cluster_rank();                  // = head index
  const int cid = blockIdx.x / H, ncl = gridDim.x / H;    // cluster index: serves the live sequences of rank cid, cid + ncl, ...

  // shared memory: inA[8][D] fp32 | xa[8][LDX] T | attb[8][LDX] T | hb[8][LDH] T | xin_s[D] | sampler scratch (+ logits) |
  //                red[2][16][RW] | ystage[8][32] fp32 | hstage[8][128] T | chunk ring
  float* inA = smem;
  T* xa = reinterpret_cast<T*>(inA + NB8 * D);
  T* attb = xa + NB8 * LDX;
  T* hb = attb + NB8 * LDX;
  float* xin_s = reinterpret_cast<float*>(hb + NB8 * LDH);
  float* samp = xin_s + D;
  float* red = samp + ((GSV_SAMPLE_SMEM_FLOATS + 3) & ~3);
  float* ystage = red + 2 * NWARP * RW;
  T* hstage = reinterpret_cast<T*>(ystage + NB8 * GSV_HEAD_DIM);
  unsigned char* ring = reinterpret_cast<unsigned char*>(hstage + NB8 * 4 * GSV_HEAD_DIM);
  const int mtile = warp & 1, kq = warp >> 1;             // this warp's share of a chunk

  const T* const Bqkv = reinterpret_cast<const T*>(p.b_qkv);
  const T* const Bo = reinterpret_cast<const T*>(p.b_o);
  const T* const B1 = reinterpret_cast<const T*>(p.b_1);
  const T* const B2 = reinterpret_cast<const T*>(p.b_2);
  const T* const G1 = reinterpret_cast<const T*>(p.ln1_g);
  const T* const Be1 = reinterpret_cast<const T*>(p.ln1_b);
  const T* const G2 = reinterpret_cast<const T*>(p.ln2_g);
  const T* const Be2 = reinterpret_cast<const T*>(p.ln2_b);

  // ---- which sequences: the active slots are dealt round-robin over the clusters (balanced attention / sampling work) ----
  if (tid < 32) {
    // slot table of up to 64: lane t looks at slots t and t + 32
    const int flag = tid < p.slots ? ld_cg(p.active + tid) : 0;
    const int flag2 = tid + 32 < p.slots ? ld_cg(p.active + tid + 32) : 0;
    const unsigned m = __ballot_sync(0xffffffffu, flag != 0);
    const unsigned m2 = __ballot_sync(0xffffffffu, flag2 != 0);
    const unsigned below = (1u << tid) - 1u;
    const int rk = __popc(m & below);                     // rank of a slot among the live ones
    const int rk2 = __popc(m) + __popc(m2 & below);
    const int pos = (flag && rk % ncl == cid) ? rk / ncl : -1;
    const int pos2 = (flag2 && rk2 % ncl == cid) ? rk2 / ncl : -1;
    if (tid < NB8) { sh.slot[tid] = -1; sh.kv[tid] = 0; sh.alive[tid] = 0.f; }
    __syncwarp();
    if (pos >= 0 && pos < NB8) { sh.slot[pos] = tid; sh.kv[pos] = ld_cg(p.kv_len + tid); sh.alive[pos] = 1.f; }
    if (pos2 >= 0 && pos2 < NB8) { sh.slot[pos2] = tid + 32; sh.kv[pos2] = ld_cg(p.kv_len + tid + 32); sh.alive[pos2] = 1.f; }
  }
  // zero the staged operands once: columns of sequences that are not live must hold finite values
  for (int i = tid; i < (NB8 * LDX * 2 + NB8 * LDH) / 2; i += NT) reinterpret_cast<unsigned*>(xa)[i] = 0u;
  __syncthreads();
  unsigned livemask = 0;                                  // bit n: sequence n of this cluster is live (uniform across its CTAs)
#pragma unroll
  for (int n = 0; n < NB8; ++n) livemask |= (sh.slot[n] >= 0 ? 1u : 0u) << n;
  if (livemask == 0) { if (tid == 0) atomicAdd(resident, 1u); return; }      // (counted as resident: gsv_gpt_wait_resident)
  int na = __popc(livemask);

  // ---- chunk ring: the CTA consumes, per token, CHUNKS chunks per layer and then its head chunks rank, rank + H, ... ----
  const int n_hchunks = ((V + 31) >> 5);
  const int HC = (n_hchunks - (int)rank + H - 1) / H;      // head chunks of this CTA
  int iss_l = 0, iss_c = 0;                                // next chunk to request (thread 0)
  int ck_slot = 0;
  unsigned ck_par = 0;
  auto issue_chunk = [&](int slot_i) {
    const unsigned char* src = iss_l < L ? pack + (((size_t)iss_l * H + rank) * CHUNKS + iss_c) * CHUNK_BYTES
                                         : hpack + (size_t)((int)rank + iss_c * H) * CHUNK_BYTES;
    mbar_expect_tx(&sh.cbar[slot_i], CHUNK_BYTES);
    bulk_g2s(ring + (size_t)slot_i * CHUNK_BYTES, src, CHUNK_BYTES, &sh.cbar[slot_i]);
    ++iss_c;
    if (iss_l < L) {
      if (iss_c == CHUNKS) { iss_c = 0; ++iss_l; if (iss_l == L && HC == 0) iss_l = 0; }
    } else if (iss_c == HC) { iss_c = 0; iss_l = 0; }
  };
  // the current chunk, once its bytes have landed
  auto chunk_wait = [&]() -> const unsigned char* {
    mbar_wait(&sh.cbar[ck_slot], ck_par);
    return ring + (size_t)ck_slot * CHUNK_BYTES;
  };
  // after a __syncthreads() that follows the last read of the current chunk: refill its slot, move on
  auto chunk_release = [&]() {
    if (tid == 0) issue_chunk(ck_slot);
    if (++ck_slot == NSLOT) { ck_slot = 0; ck_par ^= 1u; }
  };
  // this warp's share of a chunk: M-tile `mtile`, k-steps [kq * KSW, +KSW), against the staged operand rows xb[n][koff + k]
  auto mma_chunk = [&](const unsigned char* chunk, const T* xb, int ldb, int koff, float (&acc)[4]) {
    const uint4* ap = reinterpret_cast<const uint4*>(chunk) + ((size_t)(mtile * (D / 16) + kq * KSW) * 32 + lane);
    const T* brow = xb + (size_t)g * ldb + koff + kq * KSW * 16 + 2 * t;
#pragma unroll
    for (int ks = 0; ks < KSW; ++ks) {
      const uint4 a = ap[ks * 32];
      const unsigned b0 = *reinterpret_cast<const unsigned*>(brow + ks * 16);
      const unsigned b1 = *reinterpret_cast<const unsigned*>(brow + ks * 16 + 8);
      Mma16816<T>::run(acc, a.x, a.y, a.z, a.w, b0, b1);
    }
  };
  // partial accumulators: warp w writes rb[w][sequence][row of its M-tile]; thread (warp n < 8, lane = chunk row) sums
  int rsel = 0;
  auto red_write = [&](const float (&acc)[4]) {
    float* w = red + (size_t)(rsel * NWARP + warp) * RW;
    w[(2 * t) * 20 + g] = acc[0];
    w[(2 * t + 1) * 20 + g] = acc[1];
    w[(2 * t) * 20 + g + 8] = acc[2];
    w[(2 * t + 1) * 20 + g + 8] = acc[3];
  };
  auto red_read = [&]() -> float {
    const float* rb = red + (size_t)rsel * NWARP * RW + (size_t)(lane >> 4) * RW + warp * 20 + (lane & 15);
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) v += rb[(size_t)2 * q * RW];
    return v;
  };
  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) mbar_init(&sh.cbar[i], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&sh.xbar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < NSLOT; ++i) issue_chunk(i);
  }
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
  if (tid == 0) atomicAdd(resident, 1u);    // gsv_gpt_wait_resident: other streams' work is held back until every CTA is here

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
        const bool epi = warp < NB8 && ((livemask >> warp) & 1u);      // this warp finishes sequence `warp`, lane = chunk row
        float bq[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) bq[u] = Elem<T>::to_f(Bqkv[(size_t)l * 3 * D + u * D + rank * GSV_HEAD_DIM + lane]);
        if (l > 0) { mbar_wait(&sh.xbar[0], parA); parA ^= 1u; }      // y2 of the previous layer has arrived
        stage_ln(l > 0, &sh.xres[0][0]);
        __syncthreads();                                // xa staged by the LayerNorm warps; inA fully read
        if (l > 0 && tid == 0) mbar_expect_tx(&sh.xbar[0], (unsigned)na * D * 4u);     // re-arm inA for this layer's y1
        mark(p, 49);
        // chunks q, k, v: the 32 rows of this head
#pragma unroll 1
        for (int u = 0; u < 3; ++u) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          mma_chunk(chunk_wait(), xa, LDX, 0, acc);
          red_write(acc);
          __syncthreads();
          chunk_release();
          if (epi) {
            const int n = warp, c = lane;
            const float v = red_read() + (u == 0 ? bq[0] : (u == 1 ? bq[1] : bq[2]));
            if (u == 0) {
              sh.qs[n][c] = v * (rsqrtf((float)GSV_HEAD_DIM) * 1.4426950408889634f);
            } else {
              const T t16 = Elem<T>::from_f(v);         // the reference attends over the 16-bit cache entry it has just written
              (u == 1 ? sh.kn : sh.vn)[n][c] = Elem<T>::to_f(t16);
              T* cache = reinterpret_cast<T*>(u == 1 ? p.kc : p.vc);
              const size_t hbase = ((size_t)(l * p.slots + sh.slot[n]) * H + rank) * (size_t)S * GSV_HEAD_DIM;
              cache[hbase + (size_t)sh.kv[n] * GSV_HEAD_DIM + c] = t16;
            }
          }
          rsel ^= 1;
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
            // PF passes of K/V rows in flight (L2 latency >> the arithmetic of a pass)
            uint4 kq3[PF], vq3[PF];
#pragma unroll
            for (int i = 0; i < PF; ++i) {
              kq3[i] = make_uint4(0, 0, 0, 0); vq3[i] = kq3[i];
              const int pp = pb + i * 8 + pg;
              if (pp < pe) { kq3[i] = ld_cg16(kb + (size_t)pp * GSV_HEAD_DIM); vq3[i] = ld_cg16(vb + (size_t)pp * GSV_HEAD_DIM); }
            }
#pragma unroll 1
            for (int base = pb; base < pe; base += 8 * PF) {
#pragma unroll
              for (int i = 0; i < PF; ++i) {
                const int pos = base + i * 8 + pg;
                const bool ok = pos < pe;
                const uint4 kr = kq3[i], vr = vq3[i];
                if (pos + 8 * PF < pe) {
                  kq3[i] = ld_cg16(kb + (size_t)(pos + 8 * PF) * GSV_HEAD_DIM);
                  vq3[i] = ld_cg16(vb + (size_t)(pos + 8 * PF) * GSV_HEAD_DIM);
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
        const bool epi = warp < NB8 && ((livemask >> warp) & 1u);
        const float o_bias = Elem<T>::to_f(Bo[(size_t)l * D + rank * GSV_HEAD_DIM + lane]);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        mma_chunk(chunk_wait(), attb, LDX, 0, acc);
        red_write(acc);
        __syncthreads();                                // every warp has read att: re-arm its inbox for the next layer
        chunk_release();
        if (tid == 0 && l + 1 < L) mbar_expect_tx(&sh.xbar[1], (unsigned)na * D * 2u);
        if (epi) ystage[warp * GSV_HEAD_DIM + lane] = red_read() + o_bias + sh.xres[warp][lane];
        rsel ^= 1;
        __syncthreads();
        mark(p, 31);
        // y1 rows of this CTA to every CTA: 16 bytes per store, (target, sequence, quad)
        for (int i = tid; i < H * 64; i += NT) {
          const int tgt = i >> 6, n = (i >> 3) & 7, q4 = i & 7;
          if ((livemask >> n) & 1u)
            st_async_v4(inA + n * D + (int)rank * GSV_HEAD_DIM + q4 * 4, &sh.xbar[0], (unsigned)tgt,
                        *reinterpret_cast<const uint4*>(ystage + n * GSV_HEAD_DIM + q4 * 4));
        }
      }
      mark(p, 53);
      mbar_wait(&sh.xbar[0], parA); parA ^= 1u;          // y1 has arrived
      mark(p, 4);
      // ================= M1: x1 = LN1(y1); h = relu(x1 W1^T + b1) ==================
      {
        const bool epi = warp < NB8 && ((livemask >> warp) & 1u);
        float b1v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) b1v[u] = Elem<T>::to_f(B1[(size_t)l * F + rank * (4 * GSV_HEAD_DIM) + u * 32 + lane]);
        stage_ln(true, &sh.xres1[0][0]);
        __syncthreads();                                // xa staged; inA fully read
        if (tid == 0) mbar_expect_tx(&sh.xbar[0], (unsigned)na * D * 4u);             // re-arm inA for y2
        mark(p, 41);
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          mma_chunk(chunk_wait(), xa, LDX, 0, acc);
          red_write(acc);
          __syncthreads();
          chunk_release();
          if (epi) {
            const float bias = u == 0 ? b1v[0] : (u == 1 ? b1v[1] : (u == 2 ? b1v[2] : b1v[3]));
            hstage[warp * (4 * GSV_HEAD_DIM) + u * 32 + lane] = Elem<T>::from_f(fmaxf(red_read() + bias, 0.f));
          }
          rsel ^= 1;
        }
        __syncthreads();
        mark(p, 42);
        // h rows of this CTA (128 per sequence, storage type) to every CTA
        for (int i = tid; i < H * 128; i += NT) {
          const int tgt = i >> 7, n = (i >> 4) & 7, q8 = i & 15;
          if ((livemask >> n) & 1u)
            st_async_v4(hb + n * LDH + (int)rank * (4 * GSV_HEAD_DIM) + q8 * 8, &sh.xbar[2], (unsigned)tgt,
                        *reinterpret_cast<const uint4*>(hstage + n * (4 * GSV_HEAD_DIM) + q8 * 8));
        }
      }
      mark(p, 54);
      mbar_wait(&sh.xbar[2], parH); parH ^= 1u;          // h has arrived
      mark(p, 5);
      // ================= M2: y2 = x1 + h W2^T + b2 ==================
      {
        load_vec<T, NCH>(G2 + (size_t)l * D, lane, gv);
        load_vec<T, NCH>(Be2 + (size_t)l * D, lane, bv);
        const bool epi = warp < NB8 && ((livemask >> warp) & 1u);
        const float m_bias = Elem<T>::to_f(B2[(size_t)l * D + rank * GSV_HEAD_DIM + lane]);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int q4 = 0; q4 < 4; ++q4) {                // one K-quarter of the 32 rows per chunk
          mma_chunk(chunk_wait(), hb, LDH, q4 * D, acc);
          if (q4 == 3) red_write(acc);
          __syncthreads();
          chunk_release();
        }
        // every warp has read h: re-arm its inbox for the next layer
        if (tid == 0 && l + 1 < L) mbar_expect_tx(&sh.xbar[2], (unsigned)na * F * 2u);
        if (epi) ystage[warp * GSV_HEAD_DIM + lane] = red_read() + m_bias + sh.xres1[warp][lane];
        rsel ^= 1;
        __syncthreads();
        mark(p, 51);
        for (int i = tid; i < H * 64; i += NT) {
          const int tgt = i >> 6, n = (i >> 3) & 7, q = i & 7;
          if ((livemask >> n) & 1u)
            st_async_v4(inA + n * D + (int)rank * GSV_HEAD_DIM + q * 4, &sh.xbar[0], (unsigned)tgt,
                        *reinterpret_cast<const uint4*>(ystage + n * GSV_HEAD_DIM + q * 4));
        }
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
      const bool epi = warp < NB8 && ((livemask >> warp) & 1u);
#pragma unroll 1
      for (int j = 0; j < HC; ++j) {                     // 32 vocabulary rows per chunk
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        mma_chunk(chunk_wait(), xa, LDX, 0, acc);
        red_write(acc);
        __syncthreads();
        chunk_release();
        const int gg = ((int)rank + j * H) * 32 + lane;
        if (epi && gg < V) st_async(samp + gg, &sh.xbar[3], (unsigned)warp, red_read());
        rsel ^= 1;
      }
    }
    mark(p, 20);
    // ================= sampling: CTA n samples sequence n; next inputs and alive flags pushed to every CTA ==================
    if ((int)rank < NB8 && ((livemask >> rank) & 1u)) {
      const int n = (int)rank;
      SamplePre pre;                                     // the sampler's own global reads overlap the wait for the logits
      sample_prefetch<T>(p, sh.slot[n], sh.kv[n] + 1, pre);
      mbar_wait(&sh.xbar[3], parL); parL ^= 1u;          // all V logits of sequence n have arrived
      SampleLL io;
      io.preloaded = true;
      io.xin_ll = nullptr;
      io.status_ll = nullptr;
      io.tag = 0;
      io.kv_len = sh.kv[n] + 1;
      io.xin_smem = xin_s;
      io.alive_smem = &sh.alive_i;
      sample_slot<T>(p, sh.slot[n], samp, &io, &pre);
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
  for (int i = 0; i < NSLOT; ++i) {                       // requests that ran ahead of the last token
    mbar_wait(&sh.cbar[ck_slot], ck_par);
    if (++ck_slot == NSLOT) { ck_slot = 0; ck_par ^= 1u; }
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
  const int D = ctx->p.d, F = ctx->p.F, H = ctx->p.H, L = ctx->p.L, V = ctx->p.V;
  const size_t chunk_bytes = (size_t)64 * D;
  const int n_hchunks = (V + 31) / 32;
  if (!ctx->cl8_pack) {
    // one-time re-tiling of the block and head weights into chunk / fragment order (as large as the weights themselves)
    void *pk = nullptr, *hp = nullptr;
    GSV_CUDA(cudaMalloc(&pk, (size_t)L * H * CHUNKS * chunk_bytes));
    GSV_CUDA(cudaMalloc(&hp, (size_t)n_hchunks * chunk_bytes));
    cl8_pack_layers_kernel<T><<<ctx->num_sms * 4, 256, 0, st>>>(
        reinterpret_cast<const T*>(ctx->p.w_qkv), reinterpret_cast<const T*>(ctx->p.w_o), reinterpret_cast<const T*>(ctx->p.w_1),
        reinterpret_cast<const T*>(ctx->p.w_2), reinterpret_cast<uint4*>(pk), L, H, D);
    cl8_pack_head_kernel<T><<<ctx->num_sms, 256, 0, st>>>(reinterpret_cast<const T*>(ctx->p.w_head), reinterpret_cast<uint4*>(hp), n_hchunks, V, D);
    GSV_CUDA(cudaGetLastError());
    ctx->cl8_pack = pk;
    ctx->cl8_head_pack = hp;
    ctx->launches += 2;
  }
  const size_t bytes = (size_t)NB8 * D * 4 + (size_t)2 * NB8 * (D + XPAD) * 2 + (size_t)NB8 * (F + XPAD) * 2 + (size_t)D * 4 +
                       (size_t)((GSV_SAMPLE_SMEM_FLOATS + 3) & ~3) * 4 + (size_t)2 * NWARP * RW * 4 + (size_t)NB8 * GSV_HEAD_DIM * 4 +
                       (size_t)NB8 * 4 * GSV_HEAD_DIM * 2 + (size_t)NSLOT * chunk_bytes;
  GSV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  GSV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  // Clusters: each streams ALL weights through its GPC's port whatever it serves, so fewer sequences per cluster only shorten
  // its attention and sampling phases (5.0 us and 11 us of a 17 us layer / a token at eight sequences).  Measured us per step
  // (clusters: 1 / 2 / 4 / 6 / 7): 8 sequences 425 / 358 / 341 / 333 / 336; 16: - / 423 / 361 / 354 / 349; 32: - / - / 432 /
  // 400 / 392.  Two sequences per cluster, at most six clusters (seven are co-resident; the second stream's prompts and
  // vocoder need SMs too).
  int n_clusters = (live + 1) / 2;
  n_clusters = n_clusters > 6 ? 6 : n_clusters;
  {
    const char* e = getenv("GSV_CL8_CLUSTERS");
    if (e && atoi(e) > 0) n_clusters = atoi(e);
  }
  if (n_clusters * NB8 < live) n_clusters = (live + NB8 - 1) / NB8;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(n_clusters * H); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = bytes; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = H; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  GptParams p = ctx->p;
  int ns = n_steps;
  const unsigned char* pk = reinterpret_cast<const unsigned char*>(ctx->cl8_pack);
  const unsigned char* hp = reinterpret_cast<const unsigned char*>(ctx->cl8_head_pack);
  unsigned* resident = ctx->hx_resident;
  ctx->hx_resident_expected += (unsigned)(n_clusters * H);
  void* args[] = {&p, &ns, &pk, &hp, &resident};
  GSV_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  ctx->launches += 1;
  return GSV_OK;
}

}  // namespace

int gsv_gpt_decode_cl8_launch(gsv_gpt_ctx* ctx, int live_slots, int n_steps, cudaStream_t st) {
  if (ctx->dims.dtype == GSV_F16) return launch_cl8<__half>(ctx, live_slots, n_steps, st);
  return launch_cl8<__nv_bfloat16>(ctx, live_slots, n_steps, st);
}
