// capi_gpt.cu -- extern "C" entry points of the GPT half of libgsv_b200 (see include/gsv_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "gpt_internal.cuh"

static thread_local char g_err[512] = "";
void gsv_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* gsv_last_error(void) { return g_err; }
extern "C" int gsv_version(void) { return 100; }

extern "C" int gsv_device_check(int device) {
  cudaDeviceProp prop;
  GSV_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    gsv_set_error("device %d is sm_%d%d; libgsv_b200 is built for sm_100a only", device, prop.major, prop.minor);
    return GSV_ERR_NODEVICE;
  }
  return GSV_OK;
}

template <typename P>
static int dev_alloc(gsv_gpt_ctx* ctx, P** out, size_t bytes, bool zero) {
  void* ptr = nullptr;
  GSV_CUDA(cudaMalloc(&ptr, bytes));
  if (zero) GSV_CUDA(cudaMemset(ptr, 0, bytes));
  ctx->all_allocs[ctx->n_allocs++] = ptr;
  *out = reinterpret_cast<P*>(ptr);
  return GSV_OK;
}

extern "C" int gsv_gpt_create(const gsv_gpt_dims* dims, const gsv_gpt_weights* w, gsv_gpt_ctx** out) {
  GSV_ARG(dims && w && out);
  GSV_ARG(dims->n_head > 0 && dims->d_model == dims->n_head * GSV_HEAD_DIM);
  GSV_ARG(dims->d_model % 256 == 0 && dims->d_model <= 1024 && dims->d_ff == 4 * dims->d_model);
  GSV_ARG(dims->vocab > 1 && dims->vocab <= GSV_VOCAB_MAX && dims->eos >= 0 && dims->eos < dims->vocab);
  GSV_ARG(dims->max_slots >= 1 && dims->max_slots <= GSV_MAX_SLOTS);
  GSV_ARG(dims->max_seq >= 8 && dims->max_seq <= dims->n_pos);
  GSV_ARG(dims->dtype == GSV_F16 || dims->dtype == GSV_BF16);
  GSV_ARG(dims->d_bert % 32 == 0);
  int dev = 0;
  GSV_CUDA(cudaGetDevice(&dev));
  int rc = gsv_device_check(dev);
  if (rc) return rc;
  gsv_gpt_ctx* ctx = new (std::nothrow) gsv_gpt_ctx();
  GSV_ARG(ctx != nullptr);
  memset(ctx, 0, sizeof(*ctx));
  ctx->dims = *dims;
  ctx->device = dev;
  GSV_CUDA(cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, dev));
  GptParams& p = ctx->p;
  p.d = dims->d_model; p.H = dims->n_head; p.L = dims->n_layer; p.F = dims->d_ff; p.V = dims->vocab;
  p.eos = dims->eos; p.S = dims->max_seq; p.slots = dims->max_slots; p.n_pos = dims->n_pos;
  p.d_bert = dims->d_bert; p.n_phoneme = dims->n_phoneme;
  p.w_qkv = w->w_qkv; p.b_qkv = w->b_qkv; p.w_o = w->w_o; p.b_o = w->b_o; p.w_1 = w->w_1; p.b_1 = w->b_1;
  p.w_2 = w->w_2; p.b_2 = w->b_2; p.ln1_g = w->ln1_g; p.ln1_b = w->ln1_b; p.ln2_g = w->ln2_g; p.ln2_b = w->ln2_b;
  p.w_head = w->w_head; p.emb_audio = w->emb_audio; p.pe_audio = w->pe_audio; p.emb_text = w->emb_text;
  p.pe_text = w->pe_text; p.w_bert = w->w_bert; p.b_bert = w->b_bert;
  const size_t S = p.S, B = p.slots, d = p.d, F = p.F;
  const size_t kv_bytes = (size_t)p.L * B * S * d * 2;
#define A(field, bytes, zero) if ((rc = dev_alloc(ctx, &field, (bytes), (zero)))) { gsv_gpt_destroy(ctx); return rc; }
  A(p.kc, kv_bytes, true);
  A(p.vc, kv_bytes, true);
  A(p.kv_len, B * sizeof(int), true);
  A(p.x_len, B * sizeof(int), true);
  A(p.tokens, B * S * sizeof(int), true);
  A(p.n_gen, B * sizeof(int), true);
  A(p.active, B * sizeof(int), true);
  A(p.seen, B * (GSV_VOCAB_MAX / 32) * sizeof(unsigned), true);
  A(p.samp, B * sizeof(gsv_gpt_sampling), true);
  A(p.samp_count, B * sizeof(unsigned long long), true);
  A(p.xin, B * d * sizeof(float), true);
  A(p.xres, B * d * sizeof(float), true);
  A(p.q, B * d * sizeof(float), true);
  A(p.part, B * p.H * GSV_NSPLIT_MAX * GSV_PART_STRIDE * sizeof(float), true);
  A(p.y1, B * d * sizeof(float), true);
  A(p.xres1, B * d * sizeof(float), true);
  A(p.hbuf, B * F * sizeof(float), true);
  A(p.y2, B * d * sizeof(float), true);
  A(p.logits, B * GSV_VOCAB_MAX * sizeof(float), true);
  A(p.barrier, 64, true);
  A(ctx->hooks_dev, GSV_MAX_SLOTS * sizeof(GptSlotHooks), true);
  A(ctx->hx_resident, 64, true);
  // prefill scratch: rows of several prompts stacked (gsv_gpt_prefill_begin_many), at least one full-length prompt
  ctx->pf_rows = (int)(S > 4096 ? S : 4096);
  const size_t R = (size_t)ctx->pf_rows;
  A(ctx->pf_x, R * d * 2, false);
  A(ctx->pf_qkv, R * 3 * d * 2, false);
  A(ctx->pf_attn, R * d * 2, false);
  A(ctx->pf_h, R * F * 2, false);
  A(ctx->pf_tmp, R * d * 2, false);
  A(ctx->pf_last, B * d * 2, true);
  A(ctx->ll_buf, gsv_gpt_ll_buffer_bytes(ctx), true);
  A(ctx->dx, B * d * 2, true);
  A(ctx->dqkv, B * 3 * d * 2, true);
  A(ctx->datt, B * d * 2, true);
  A(ctx->dh, B * F * 2, true);
  A(ctx->dtmp, B * d * 2, true);
#undef A
  {
    const char* e = getenv("GSV_DECODE_IMPL");
    ctx->force_barrier_kernel = (e && strcmp(e, "barrier") == 0) ? 1 : 0;
    ctx->force_ll1 = (e && strcmp(e, "ll1") == 0) ? 1 : 0;
    ctx->force_hx = (e && strcmp(e, "hx") == 0) ? 1 : 0;
    ctx->force_gemm = (e && strcmp(e, "gemm") == 0) ? 1 : 0;
    ctx->use_cl = (e && strcmp(e, "cl") == 0) ? 1 : 0;
    ctx->use_cl8 = (e && strcmp(e, "cl8") == 0) ? 1 : 0;
    const char* g = getenv("GSV_GPT_GEMM");
    ctx->use_umma_linear = (g && strcmp(g, "cuda") == 0) ? 0 : 1;
    ctx->umma = gsv_umma_cache_create(ctx->num_sms);
  }
  if ((rc = gsv_gpt_decode_configure(ctx))) { gsv_gpt_destroy(ctx); return rc; }
  *out = ctx;
  return GSV_OK;
}

extern "C" int gsv_gpt_destroy(gsv_gpt_ctx* ctx) {
  if (!ctx) return GSV_OK;
  for (int i = 0; i < ctx->n_allocs; ++i) cudaFree(ctx->all_allocs[i]);
  if (ctx->step_graph_exec) cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(ctx->step_graph_exec));
  gsv_umma_cache_destroy(ctx->umma);
  if (ctx->hx_pack) cudaFree(ctx->hx_pack);
  if (ctx->hx_head_pack) cudaFree(ctx->hx_head_pack);
  if (ctx->cl8_pack) cudaFree(ctx->cl8_pack);
  if (ctx->cl8_head_pack) cudaFree(ctx->cl8_head_pack);
  delete ctx;
  return GSV_OK;
}

static int check_prompt(gsv_gpt_ctx* ctx, int nx, int ny) {
  if (nx + ny >= ctx->p.S) {
    // the reference does not validate this and fails with a shape error inside process_prompt
    // (t2s_model.py:49; SURVEY.md 8b "Errors"); here it is an argument error
    gsv_set_error("prompt length %d+%d does not fit the KV cache (max_seq %d)", nx, ny, ctx->p.S);
    return GSV_ERR_ARG;
  }
  return GSV_OK;
}

extern "C" int gsv_gpt_prefill(gsv_gpt_ctx* ctx, int slot, const int64_t* dev_x, int nx, const int64_t* dev_y, int ny,
                               const void* dev_bert, const gsv_gpt_sampling* samp, void* stream) {
  GSV_ARG(ctx && dev_x && dev_y && dev_bert && samp);
  GSV_ARG(slot >= 0 && slot < ctx->p.slots);
  GSV_ARG(nx >= 1 && ny >= 1);
  int rc = check_prompt(ctx, nx, ny);
  if (rc) return rc;
  if ((rc = gsv_gpt_prefill_body(ctx, slot, dev_x, nx, dev_y, ny, dev_bert, (cudaStream_t)stream))) return rc;
  ctx->slot_live[slot] = 1;
  rc = gsv_gpt_prefill_tail(ctx, slot, dev_y, ny, samp, (cudaStream_t)stream);
  ctx->pf_n[slot] = 0;
  return rc;
}

extern "C" int gsv_gpt_prefill_begin(gsv_gpt_ctx* ctx, int slot, const int64_t* dev_x, int nx, const int64_t* dev_y, int ny,
                                     const void* dev_bert, void* stream) {
  GSV_ARG(ctx && dev_x && dev_y && dev_bert);
  GSV_ARG(slot >= 0 && slot < ctx->p.slots);
  GSV_ARG(nx >= 1 && ny >= 1);
  int rc = check_prompt(ctx, nx, ny);
  if (rc) return rc;
  // caller's contract: the slot's previous sequence has finished (its `active` flag read back as 0) or was released
  return gsv_gpt_prefill_body(ctx, slot, dev_x, nx, dev_y, ny, dev_bert, (cudaStream_t)stream);
}

extern "C" int gsv_gpt_prefill_begin_many(gsv_gpt_ctx* ctx, int n_prompts, const int* slots, const int64_t* const* dev_x, const int* nx,
                                          const int64_t* const* dev_y, const int* ny, const void* const* dev_bert, void* stream) {
  GSV_ARG(ctx && slots && dev_x && nx && dev_y && ny && dev_bert);
  GSV_ARG(n_prompts >= 1 && n_prompts <= GSV_MAX_SLOTS);
  long long rows = 0;
  unsigned long long seen = 0ull;
  for (int i = 0; i < n_prompts; ++i) {
    GSV_ARG(dev_x[i] && dev_y[i] && dev_bert[i]);
    GSV_ARG(slots[i] >= 0 && slots[i] < ctx->p.slots);
    GSV_ARG(nx[i] >= 1 && ny[i] >= 1);
    if ((seen >> slots[i]) & 1ull) { gsv_set_error("gsv_gpt_prefill_begin_many: slot %d named twice", slots[i]); return GSV_ERR_ARG; }
    seen |= 1ull << slots[i];
    const int rc = check_prompt(ctx, nx[i], ny[i]);
    if (rc) return rc;
    rows += nx[i] + ny[i];
  }
  if (rows > ctx->pf_rows) {
    gsv_set_error("gsv_gpt_prefill_begin_many: %lld prompt rows exceed the pass capacity %d (gsv_gpt_prefill_capacity)", rows, ctx->pf_rows);
    return GSV_ERR_ARG;
  }
  return gsv_gpt_prefill_body_many(ctx, n_prompts, slots, dev_x, nx, dev_y, ny, dev_bert, (cudaStream_t)stream);
}

extern "C" int gsv_gpt_prefill_capacity(gsv_gpt_ctx* ctx) { return ctx ? ctx->pf_rows : 0; }

extern "C" int gsv_gpt_prefill_finish(gsv_gpt_ctx* ctx, int slot, const int64_t* dev_y, int ny, const gsv_gpt_sampling* samp,
                                      void* stream) {
  GSV_ARG(ctx && dev_y && samp);
  GSV_ARG(slot >= 0 && slot < ctx->p.slots);
  if (ctx->pf_n[slot] <= 0) { gsv_set_error("gsv_gpt_prefill_finish: no gsv_gpt_prefill_begin pending for slot %d", slot); return GSV_ERR_STATE; }
  GSV_ARG(ny >= 1 && ny < ctx->pf_n[slot]);
  ctx->slot_live[slot] = 1;
  const int rc = gsv_gpt_prefill_tail(ctx, slot, dev_y, ny, samp, (cudaStream_t)stream);
  ctx->pf_n[slot] = 0;
  return rc;
}

extern "C" int gsv_gpt_decode(gsv_gpt_ctx* ctx, int n_steps, void* stream) {
  GSV_ARG(ctx && n_steps >= 1);
  int live = 0;
  for (int i = 0; i < ctx->p.slots; ++i) live += ctx->slot_live[i];
  if (live == 0) { gsv_set_error("gsv_gpt_decode: no slot has been prefilled"); return GSV_ERR_STATE; }
  // Measured on B200 (tools/decode_speed.py / bench.py, bf16, kv 164..364), us per step (tokens per second):
  //   live   hx (head clusters)   ll (grid-wide LL)   cl (1 seq/cluster)   cl8 (<= 8/cluster, mma)   gemm (multi-kernel, tcgen05 linears)
  //    1     168 ( 5.9k)          293                 355                                            1760
  //    2      -                   381                 350 ( 5.7k)          334 ( 6.0k)
  //    4      -                   595                 352 (11.4k)          332 (12.1k)
  //    6      -                    -                  354 (17.0k)          338 (17.7k)   <- at most 7 sixteen-CTA clusters are co-resident; more run in waves
  //    8      -                    -                  712 (11.2k)          341 (23.5k)               1567 ( 5.1k)
  //   16      -                    -                   -                   354 (45.2k)
  //   32      -                    -                ~2140 (14.9k)          400 (79.9k)               1583 (20.2k)
  //   (cl8 with the live sequences dealt over min(6, ceil(live / 2)) clusters)
  // -> 1: head-cluster kernel (grid-wide LL kernel where its clusters are not co-resident); 2 and more: the tensor-core
  //    cluster kernel (one cluster per sequence for models with fewer than 8 heads); the multi-kernel step where clusters of
  //    H CTAs cannot be launched; the grid-barrier kernel for shapes none of them takes (GSV_DECODE_IMPL overrides).
  const bool explicit_impl = ctx->force_barrier_kernel || ctx->force_ll1 || ctx->force_gemm;
  // one live sequence: head-cluster kernel (2 grid-wide exchanges per layer instead of 5)
  if ((ctx->force_hx || (live == 1 && !explicit_impl && !ctx->use_cl && !ctx->use_cl8)) && gsv_gpt_hx_supported(ctx, live, n_steps)) {
    const int rc = gsv_gpt_decode_hx_launch(ctx, n_steps, (cudaStream_t)stream);
    if (rc != GSV_ERR_STATE || ctx->force_hx) return rc;      // GSV_ERR_STATE: clusters not co-resident here -> grid-wide kernel below
  }
  if ((ctx->use_cl8 || (!explicit_impl && !ctx->use_cl && live >= 2)) && gsv_gpt_cl_supported(ctx, live) && ctx->p.H >= 8)
    return gsv_gpt_decode_cl8_launch(ctx, live, n_steps, (cudaStream_t)stream);
  // (the kernels below rank the first 32 slots only; slots 32.. exist for requests that wait prefilled beside a full batch)
  bool hi = false;
  for (int i = 32; i < ctx->p.slots; ++i) hi = hi || ctx->slot_live[i];
  if (!hi && (ctx->use_cl || (!explicit_impl && live >= 2 && live <= 7)) && gsv_gpt_cl_supported(ctx, live))   // models with fewer than 8 heads
    return gsv_gpt_decode_cl_launch(ctx, live, n_steps, (cudaStream_t)stream);
  if (ctx->force_gemm || ((live > 4 || hi) && !ctx->force_barrier_kernel && ctx->use_umma_linear))
    return gsv_gpt_decode_gemm_launch(ctx, n_steps, (cudaStream_t)stream);
  if (hi) { gsv_set_error("gsv_gpt_decode: live slots beyond 32 need the cluster kernels or the multi-kernel step"); return GSV_ERR_STATE; }
  if (!ctx->force_barrier_kernel && gsv_gpt_ll_supported(ctx, live, n_steps))
    return gsv_gpt_decode_ll_launch(ctx, live, n_steps, (cudaStream_t)stream);
  return gsv_gpt_decode_launch(ctx, n_steps, (cudaStream_t)stream);
}

extern "C" int gsv_gpt_read(gsv_gpt_ctx* ctx, int32_t* host_n_gen, int32_t* host_active, int32_t* host_tokens, int first_slot,
                            int n_slots, void* stream) {
  GSV_ARG(ctx && first_slot >= 0 && n_slots >= 1 && first_slot + n_slots <= ctx->p.slots);
  cudaStream_t st = (cudaStream_t)stream;
  if (host_n_gen)
    GSV_CUDA(cudaMemcpyAsync(host_n_gen, ctx->p.n_gen + first_slot, n_slots * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (host_active)
    GSV_CUDA(cudaMemcpyAsync(host_active, ctx->p.active + first_slot, n_slots * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (host_tokens)
    GSV_CUDA(cudaMemcpyAsync(host_tokens, ctx->p.tokens + (size_t)first_slot * ctx->p.S,
                             (size_t)n_slots * ctx->p.S * sizeof(int), cudaMemcpyDeviceToHost, st));
  return GSV_OK;
}

extern "C" int gsv_gpt_state_ptrs(gsv_gpt_ctx* ctx, int32_t** dev_tokens, int32_t** dev_n_gen, int32_t** dev_active) {
  GSV_ARG(ctx);
  if (dev_tokens) *dev_tokens = ctx->p.tokens;
  if (dev_n_gen) *dev_n_gen = ctx->p.n_gen;
  if (dev_active) *dev_active = ctx->p.active;
  return GSV_OK;
}

extern "C" int gsv_gpt_release_slot(gsv_gpt_ctx* ctx, int slot, void* stream) {
  GSV_ARG(ctx && slot >= -1 && slot < ctx->p.slots);
  if (slot < 0) {                          // every slot: one memset instead of one per slot at the start of every request
    GSV_CUDA(cudaMemsetAsync(ctx->p.active, 0, sizeof(int) * (size_t)ctx->p.slots, (cudaStream_t)stream));
    for (int i = 0; i < ctx->p.slots; ++i) ctx->slot_live[i] = 0;
    return GSV_OK;
  }
  GSV_CUDA(cudaMemsetAsync(ctx->p.active + slot, 0, sizeof(int), (cudaStream_t)stream));
  ctx->slot_live[slot] = 0;
  return GSV_OK;
}

// The batched step's CUDA graph bakes the parameter block into its kernel nodes: drop it whenever a hook changes it.
static void invalidate_step_graph(gsv_gpt_ctx* ctx) {
  if (ctx->step_graph_exec) {
    cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(ctx->step_graph_exec));
    ctx->step_graph_exec = nullptr;
  }
}

// Parity hooks live in a device array indexed by slot; the kernels see it only while at least one hook is set.
static int push_slot_hooks(gsv_gpt_ctx* ctx, int slot, const GptSlotHooks& h, cudaStream_t st, bool device_sync) {
  GptSlotHooks& cur = ctx->hooks_host[slot];
  if (cur.noise == h.noise && cur.noise_rows == h.noise_rows && cur.forced == h.forced && cur.n_forced == h.n_forced &&
      cur.trace == h.trace && cur.trace_max == h.trace_max && !(h.noise || h.forced || h.trace))
    return GSV_OK;                                  // off -> off: nothing to do (the product path never pays for a hook)
  cur = h;
  if (device_sync) {
    GSV_CUDA(cudaDeviceSynchronize());
    GSV_CUDA(cudaMemcpy(ctx->hooks_dev + slot, &cur, sizeof(cur), cudaMemcpyHostToDevice));
  } else {
    GSV_CUDA(cudaMemcpyAsync(ctx->hooks_dev + slot, &cur, sizeof(cur), cudaMemcpyHostToDevice, st));   // pageable source: staged before return
  }
  bool any = false;
  for (int i = 0; i < ctx->p.slots; ++i) any = any || ctx->hooks_host[i].noise || ctx->hooks_host[i].forced || ctx->hooks_host[i].trace;
  GptSlotHooks* want = any ? ctx->hooks_dev : nullptr;
  if (want != ctx->p.hooks) { invalidate_step_graph(ctx); ctx->p.hooks = want; }
  return GSV_OK;
}

extern "C" int gsv_gpt_set_slot_hooks(gsv_gpt_ctx* ctx, int slot, const float* dev_noise, int n_noise_rows, const int32_t* dev_forced,
                                      int n_forced, float* dev_trace_rows, int max_trace_rows, void* stream) {
  GSV_ARG(ctx && slot >= 0 && slot < ctx->p.slots);
  GptSlotHooks h;
  memset(&h, 0, sizeof(h));
  h.noise = dev_noise; h.noise_rows = dev_noise ? n_noise_rows : 0;
  h.forced = dev_forced; h.n_forced = dev_forced ? n_forced : 0;
  h.trace = dev_trace_rows; h.trace_max = dev_trace_rows ? max_trace_rows : 0;
  return push_slot_hooks(ctx, slot, h, (cudaStream_t)stream, false);
}

extern "C" int gsv_gpt_set_noise(gsv_gpt_ctx* ctx, const float* dev_noise, int n_rows) {
  GSV_ARG(ctx);
  GptSlotHooks h = ctx->hooks_host[0];
  h.noise = dev_noise; h.noise_rows = dev_noise ? n_rows : 0; h.forced_pos = 0; h.trace_pos = 0;
  return push_slot_hooks(ctx, 0, h, nullptr, true);
}

extern "C" int gsv_gpt_set_forced(gsv_gpt_ctx* ctx, const int32_t* dev_forced, int n) {
  GSV_ARG(ctx);
  GptSlotHooks h = ctx->hooks_host[0];
  h.forced = dev_forced; h.n_forced = dev_forced ? n : 0; h.forced_pos = 0; h.trace_pos = 0;
  return push_slot_hooks(ctx, 0, h, nullptr, true);
}

extern "C" int gsv_gpt_set_logits_trace(gsv_gpt_ctx* ctx, float* dev_rows, int max_rows) {
  GSV_ARG(ctx);
  GptSlotHooks h = ctx->hooks_host[0];
  h.trace = dev_rows; h.trace_max = dev_rows ? max_rows : 0; h.forced_pos = 0; h.trace_pos = 0;
  return push_slot_hooks(ctx, 0, h, nullptr, true);
}

extern "C" int64_t gsv_gpt_launch_count(gsv_gpt_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int gsv_gpt_set_decode_sms(gsv_gpt_ctx* ctx, int n_sms) {
  GSV_ARG(ctx && n_sms >= 0);
  ctx->decode_sms = n_sms;
  return GSV_OK;
}

extern "C" int gsv_gpt_wait_resident(gsv_gpt_ctx* ctx, void* stream) {
  GSV_ARG(ctx);
  return gsv_gpt_hx_gate(ctx, (cudaStream_t)stream);
}

extern "C" int gsv_gpt_set_timeline(gsv_gpt_ctx* ctx, int64_t* dev_records, int max_records, int cta) {
  GSV_ARG(ctx);
  invalidate_step_graph(ctx);
  ctx->p.prof = reinterpret_cast<long long*>(dev_records);
  ctx->p.prof_max = dev_records ? max_records : 0;
  ctx->p.prof_cta = cta;
  return GSV_OK;
}
