// gpt_internal.cuh -- parameter block shared by the GPT prefill and decode kernels.
#pragma once
#include "common.cuh"

#define GSV_MAX_SLOTS 64      // device slot table; the cluster decode kernels (hx, cl8) rank all 64, the others the first 32
#define GSV_SLOT_TILE 8          // slots whose activations are staged in shared memory together
#define GSV_NSPLIT_MAX 16        // split-KV factor upper bound
#define GSV_PART_STRIDE 36       // floats per attention partial: m, l, pad, pad, o[32]
#define GSV_HEAD_DIM 32
#define GSV_DECODE_THREADS 512
#define GSV_VOCAB_MAX 2048       // sampling scratch is sized for this

// Parity hooks of one slot (tests only; they do not change the arithmetic).  Device-resident so that a refill can
// re-point them between two launches without touching the kernels' parameter block.
struct GptSlotHooks {
  const float* noise;   // [noise_rows][V] Exp(1) rows consumed one per sample() call of the slot's current request
  const int* forced;    // [n_forced] teacher forcing: the i-th sample() call returns forced[i]
  float* trace;         // [trace_max][V] raw logits appended per sample() call (row 0 = prefill)
  int noise_rows, n_forced, trace_max;
  int forced_pos, trace_pos, pad;
};

struct GptParams {
  int d, H, L, F, V, eos, S, slots, n_pos, d_bert, n_phoneme;
  // weights (element type T)
  const void *w_qkv, *b_qkv, *w_o, *b_o, *w_1, *b_1, *w_2, *b_2;
  const void *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  const void *w_head, *emb_audio, *pe_audio, *emb_text, *pe_text, *w_bert, *b_bert;
  // KV cache, layout [L][slots][H][S][32] (T): one (layer, slot, head) stream is contiguous
  void *kc, *vc;
  // per-slot state
  int* kv_len;      // [slots] live positions in the cache
  int* x_len;       // [slots] Nx (phoneme count): PE index of the next token is kv_len - x_len
  int* tokens;      // [slots][S] sampled tokens, index 0 = first token after prefill (s0)
  int* n_gen;       // [slots]
  int* active;      // [slots]
  unsigned* seen;   // [slots][GSV_VOCAB_MAX/32] bitmap of prompt + sampled tokens (repetition penalty)
  gsv_gpt_sampling* samp;           // [slots]
  unsigned long long* samp_count;   // [slots] sample() calls so far (Philox counter / noise row)
  // step buffers exchanged between CTAs (fp32, read with ld.cg)
  float* xin;     // [slots][d] input of layer 0 for the next step (embedding + PE)
  float* xres;    // [slots][d] post-LN layer input, residual for the attention half
  float* q;       // [slots][d]
  float* part;    // [slots][H][GSV_NSPLIT_MAX][GSV_PART_STRIDE]
  float* y1;      // [slots][d] x + out_proj(attn) (pre-LN1)
  float* xres1;   // [slots][d] LN1 output, residual for the MLP half
  float* hbuf;    // [slots][F]
  float* y2;      // [slots][d] pre-LN2
  float* logits;  // [slots][GSV_VOCAB_MAX]
  unsigned* barrier;
  // parity hooks: [slots], or nullptr while none is set (the product path)
  GptSlotHooks* hooks;
  // in-kernel timeline (tools/decode_timeline.py): {id, clock64} records written by one CTA
  long long* prof; int prof_max; int prof_cta;
};

struct gsv_umma_cache;

struct gsv_gpt_ctx {
  gsv_gpt_dims dims;
  GptParams p;
  int device;
  int num_sms;
  int decode_grid;
  size_t decode_smem;
  long long launches;
  // prefill scratch (T unless noted), sized for max_seq rows
  void *pf_x, *pf_qkv, *pf_attn, *pf_h, *pf_tmp;   // [pf_rows][...] T: stacked rows of the prompts of one prefill pass
  int pf_rows;
  void* pf_last;                  // [slots][d] T: last prompt row of a prefill body, read by its tail
  int pf_nx[GSV_MAX_SLOTS], pf_n[GSV_MAX_SLOTS];   // text / total prompt length of the body waiting for its tail (0 = none)
  float* pf_f32;
  void* all_allocs[64];
  int n_allocs;
  // small-batch LL decode kernel (gpt_decode_ll.cu)
  void* ll_buf;
  unsigned long long ll_seq;
  int slot_live[GSV_MAX_SLOTS];   // host-side view: prefilled and not yet released
  GptSlotHooks* hooks_dev;        // [slots] device copy of the parity hooks
  GptSlotHooks hooks_host[GSV_MAX_SLOTS];
  gsv_umma_cache* umma;           // tensor-map cache of the prefill / batched-decode linears
  void *dx, *dqkv, *datt, *dh, *dtmp;   // batched decode step rows [slots][.] (T)
  void* step_graph_exec;          // cudaGraphExec_t of one batched decode step
  int force_gemm;                 // GSV_DECODE_IMPL=gemm: multi-kernel tensor-core step for any live count
  int use_cl;                     // GSV_DECODE_IMPL=cl: cluster-per-sequence kernel
  int decode_sms;                 // gsv_gpt_set_decode_sms: CTAs of the single-sequence decode kernel (0 = every SM)
  void* cl8_pack;                 // gpt_decode_cl8.cu: block weights re-tiled into chunk / mma-fragment order (first launch)
  void* cl8_head_pack;            // ... and the head rows
  int use_cl8;                    // GSV_DECODE_IMPL=cl8: the tensor-core cluster kernel for any live count
  void* hx_pack;                  // gpt_decode_hx.cu: per-(layer, CTA) weight blobs (first launch)
  void* hx_head_pack;             // ... and the per-CTA head rows
  int hx_clusters_ok;             // 0 not asked yet, 1 every cluster of the kernel is co-resident, -1 not
  int hx_cs;                      // CTAs per cluster of the head-cluster kernel (0: not chosen yet)
  unsigned* hx_resident;          // device counter: CTAs of head-cluster launches that have become resident
  unsigned hx_resident_expected;  // ... and its value once the last launch is fully resident
  int force_hx;                   // GSV_DECODE_IMPL=hx
  int use_umma_linear;            // GSV_GPT_GEMM=cuda disables the tensor-core linears (A/B checks)
  int force_barrier_kernel;       // GSV_DECODE_IMPL=barrier
  int force_ll1;                  // GSV_DECODE_IMPL=ll1: first-generation small-batch kernel (A/B checks)
};

// nn.Linear on the tcgen05 implicit-GEMM kernel (vocoder.cu / conv_umma.cuh): out[r][n] = act(X[r] . W[n] + bias[n]),
// X [rows_cap][K] T (rows >= `rows` are ignored), W [N][K] T, out [rows][N] T.  `op` names the call site (tensor maps
// are cached per call site).
gsv_umma_cache* gsv_umma_cache_create(int num_sms);
void gsv_umma_cache_destroy(gsv_umma_cache* c);
int gsv_umma_linear(gsv_umma_cache* c, size_t op, int dtype, const void* X, int rows, int rows_cap, int K, const void* W,
                    const void* bias, int N, void* out, int relu, cudaStream_t st);

// Conv1d with KW taps ('same' padding, odd KW) on the same kernel: X [rows][K] T, W [KW][N][K] T, out [rows][N] T.
int gsv_umma_conv(gsv_umma_cache* c, size_t op, int dtype, const void* X, int rows, int K, const void* W, const void* bias, int N, int KW,
                  void* out, int relu, cudaStream_t st);

// kernels / launchers implemented in gpt_decode.cu and gpt_prefill.cu
int gsv_gpt_decode_launch(gsv_gpt_ctx* ctx, int n_steps, cudaStream_t st);
int gsv_gpt_prefill_body_many(gsv_gpt_ctx* ctx, int n_prompts, const int* slots, const int64_t* const* xs, const int* nxs,
                              const int64_t* const* ys, const int* nys, const void* const* berts, cudaStream_t st);
int gsv_gpt_prefill_body(gsv_gpt_ctx* ctx, int slot, const int64_t* x, int nx, const int64_t* y, int ny, const void* bert, cudaStream_t st);
int gsv_gpt_prefill_tail(gsv_gpt_ctx* ctx, int slot, const int64_t* y, int ny, const gsv_gpt_sampling* samp, cudaStream_t st);
int gsv_gpt_decode_configure(gsv_gpt_ctx* ctx);
size_t gsv_gpt_ll_buffer_bytes(const gsv_gpt_ctx* ctx);
bool gsv_gpt_ll_supported(const gsv_gpt_ctx* ctx, int live_slots, int n_steps);
int gsv_gpt_decode_ll_launch(gsv_gpt_ctx* ctx, int live_slots, int n_steps, cudaStream_t st);
int gsv_gpt_decode_gemm_launch(gsv_gpt_ctx* ctx, int n_steps, cudaStream_t st);
bool gsv_gpt_cl_supported(const gsv_gpt_ctx* ctx, int live_slots);
int gsv_gpt_decode_cl_launch(gsv_gpt_ctx* ctx, int live_slots, int n_steps, cudaStream_t st);
int gsv_gpt_decode_cl8_launch(gsv_gpt_ctx* ctx, int live_slots, int n_steps, cudaStream_t st);
size_t gsv_gpt_hx_buffer_words(const gsv_gpt_ctx* ctx);
bool gsv_gpt_hx_supported(const gsv_gpt_ctx* ctx, int live_slots, int n_steps);
int gsv_gpt_decode_hx_launch(gsv_gpt_ctx* ctx, int n_steps, cudaStream_t st);
int gsv_gpt_hx_gate(gsv_gpt_ctx* ctx, cudaStream_t st);
