/* gsv_b200.h -- C ABI of libgsv_b200.so: the two hot paths of GSV-TTS-Lite on B200 (sm_100a).
 *
 * The reference (chinokikiss/GSV-TTS-Lite, /root/reference) has no native layer: its seam is
 * Python duck typing at two objects built by gsv_tts/Loader.py.  Each entry point below cites
 * the reference code whose work it takes over.  Signatures use plain pointers and sizes only:
 * no torch types cross this boundary.  All pointers named dev_* are device pointers owned by
 * the caller; `stream` is a cudaStream_t passed as void*.  Every function returns GSV_OK (0)
 * or a negative gsv_status and never throws; gsv_last_error() gives the message.
 *
 * Element type of every "T" array is dims.dtype (fp16 or bf16); accumulation is fp32.
 */
#ifndef GSV_B200_H
#define GSV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  GSV_OK = 0,
  GSV_ERR_ARG = -1,      /* bad argument / unsupported shape */
  GSV_ERR_CUDA = -2,     /* CUDA runtime error (message in gsv_last_error) */
  GSV_ERR_STATE = -3,    /* call sequence error (e.g. decode on an empty slot set) */
  GSV_ERR_NODEVICE = -4  /* not an sm_100 device */
} gsv_status;

typedef enum { GSV_F16 = 0, GSV_BF16 = 1 } gsv_dtype;

const char* gsv_last_error(void);
int gsv_version(void);
/* Device check: GSV_OK only on compute capability 10.x. */
int gsv_device_check(int device);

/* ======================================================================================
 * GPT semantic-token decoder
 * replaces Text2SemanticDecoder (reference gsv_tts/GPT_SoVITS/GPT/t2s_model.py:158-734 and
 * t2s_model_flash_attn.py) and what it calls in torch / flash-attn.
 * ====================================================================================== */

typedef struct {
  int32_t d_model;     /* config["model"]["hidden_dim"]   (t2s_model.py:161) */
  int32_t n_head;      /* config["model"]["head"]; head_dim must be 32       */
  int32_t n_layer;
  int32_t d_ff;        /* 4*d_model (t2s_model.py:14 mlp_ratio)             */
  int32_t vocab;       /* 1025, EOS included                                 */
  int32_t eos;
  int32_t n_phoneme;
  int32_t d_bert;      /* 1024 (t2s_model.py:172)                            */
  int32_t n_pos;       /* rows of the positional tables (4000, :212-213)     */
  int32_t dtype;       /* gsv_dtype                                          */
  int32_t max_slots;   /* KV-cache slots: max batch size over gpt_cache + the spare slots of infer_batched (<= 64) */
  int32_t max_seq;     /* max sequence length over gpt_cache                 */
} gsv_gpt_dims;

/* Device pointers to weights in the reference's own row-major layouts (nn.Linear: [out][in]).
 * Per-layer tensors are stacked on a leading layer axis.  All of type T. */
typedef struct {
  const void* w_qkv;   /* [L][3d][d]   blocks.{i}.qkv.weight      */
  const void* b_qkv;   /* [L][3d]                                 */
  const void* w_o;     /* [L][d][d]    out_proj.weight            */
  const void* b_o;     /* [L][d]                                  */
  const void* w_1;     /* [L][F][d]    mlp.0.weight               */
  const void* b_1;     /* [L][F]                                  */
  const void* w_2;     /* [L][d][F]    mlp.2.weight               */
  const void* b_2;     /* [L][d]                                  */
  const void* ln1_g;   /* [L][d] norm1.weight                     */
  const void* ln1_b;   /* [L][d] norm1.bias                       */
  const void* ln2_g;   /* [L][d] norm2.weight                     */
  const void* ln2_b;   /* [L][d] norm2.bias                       */
  const void* w_head;  /* [V][d]  ar_predict_layer.weight (no bias) */
  const void* emb_audio; /* [V][d] ar_audio_embedding              */
  const void* pe_audio;  /* [n_pos][d] alpha_audio * pe ("pe_cache", t2s_model.py:409) */
  const void* emb_text;  /* [P][d] ar_text_embedding               */
  const void* pe_text;   /* [n_pos][d] alpha_text * pe             */
  const void* w_bert;    /* [d][d_bert] bert_proj.weight           */
  const void* b_bert;    /* [d]                                    */
} gsv_gpt_weights;

/* Per-slot sampling controls = keyword arguments of infer / infer_stream / infer_batched
 * (t2s_model.py:386-397, 556-566) and of sample() (GPT/utils.py:12-59). */
typedef struct {
  int32_t top_k;              /* <=0: disabled                                   */
  float top_p;                /* >=1: disabled                                   */
  float temperature;
  float repetition_penalty;   /* 1.0: disabled (batched mode, t2s_model.py:651)  */
  int32_t suppress_steps;     /* tokens {280,486,EOS} masked while idx < this (initial_suppression_steps; 0 batched) */
  int32_t max_new_tokens;     /* <=0: until the cache is full; else EOS is forced after this many   */
  int32_t mask_eos;           /* bench hook: never sample EOS              */
  int32_t max_kv;             /* <=0: dims.max_seq; else the slot stops when kv_len reaches this (the
                                 largest bucket length of its batch size, t2s_model.py:425, :656)     */
  int32_t suppress_first;     /* 1: the FIRST sample masks {280,486,EOS} whatever suppress_steps is (infer and
                                 infer_stream, t2s_model.py:415); 0: it does not (infer_batched, :613) */
  int32_t reserved;
  uint64_t seed;              /* Philox key for the slot's Exp(1) noise          */
} gsv_gpt_sampling;

typedef struct gsv_gpt_ctx gsv_gpt_ctx;

/* Allocates KV cache [L][slots][H][S][32] x2 and all step buffers (initialize_runtime,
 * t2s_model.py:210-298: one K root, one V root, static step I/O).  Nothing is captured here: decode is one
 * persistent launch per call.  The first decode with 8 or more live sequences re-tiles the block weights into
 * the tensor-core kernel's chunk order (one extra copy of the weights, kept until destroy). */
int gsv_gpt_create(const gsv_gpt_dims* dims, const gsv_gpt_weights* w, gsv_gpt_ctx** out);
int gsv_gpt_destroy(gsv_gpt_ctx* ctx);

/* Prefill one request into a cache slot and sample its first token:
 * process_single_data + T2STransformer.process_prompt + first sample() (t2s_model.py:351-383,
 * 114-127, 414-420; refill path :696-722).  dev_x [nx] int64 phoneme ids, dev_y [ny] int64
 * prompt tokens, dev_bert [nx][d_bert] T.  The slot becomes active. */
int gsv_gpt_prefill(gsv_gpt_ctx* ctx, int slot, const int64_t* dev_x, int nx, const int64_t* dev_y, int ny,
                    const void* dev_bert, const gsv_gpt_sampling* samp, void* stream);

/* The same prefill in two halves, for refilling a slot of a running batch without stalling the other slots (the
 * reference prefills a finished slot's successor eagerly between two decode steps, t2s_model.py:696-722, and every
 * live slot waits for it).  `begin` computes the prompt (embeddings, all layers, the slot's K/V rows) and may be
 * enqueued on ANOTHER stream while gsv_gpt_decode launches run: decode kernels never touch an inactive slot.
 * `finish` samples the first token and makes the slot live; enqueue it on the stream of the decode launches, after
 * the caller has ordered it behind `begin` (event).  The slot must be idle (its previous sequence finished or
 * released); dev_y must stay valid until `finish` has run. */
int gsv_gpt_prefill_begin(gsv_gpt_ctx* ctx, int slot, const int64_t* dev_x, int nx, const int64_t* dev_y, int ny,
                          const void* dev_bert, void* stream);
int gsv_gpt_prefill_finish(gsv_gpt_ctx* ctx, int slot, const int64_t* dev_y, int ny, const gsv_gpt_sampling* samp,
                           void* stream);

/* `begin` for several idle slots in ONE pass: the prompts' rows are stacked, so each linear of a layer is
 * one tensor-core launch over all of them and the launch count does not grow with n_prompts (the reference's refill path
 * prefills one request at a time with ~20 eager ops per layer, t2s_model.py:696-722).  Host arrays of length n_prompts:
 * slots (distinct), device pointers dev_x[i] / dev_y[i] / dev_bert[i], lengths nx[i] / ny[i]; the sum of nx[i] + ny[i] must
 * not exceed gsv_gpt_prefill_capacity(ctx).  Each slot is then made live by its own gsv_gpt_prefill_finish. */
int gsv_gpt_prefill_begin_many(gsv_gpt_ctx* ctx, int n_prompts, const int* slots, const int64_t* const* dev_x, const int* nx,
                               const int64_t* const* dev_y, const int* ny, const void* const* dev_bert, void* stream);
int gsv_gpt_prefill_capacity(gsv_gpt_ctx* ctx);

/* Run up to n_steps decode steps over every active slot: T2STransformer.decode_next_token x n
 * (t2s_model.py:129-143) + ar_predict_layer + sample + next-token embedding (:430-456, :637-653, :727-728),
 * with per-slot stop at EOS / full cache evaluated on the device (no host sync per token, cf. :426, :451-453).
 * The kernel is picked from the number of live slots: 1 -> head-cluster kernel (4 CTAs per attention head, 2 grid-wide
 * exchanges per layer), 2..7 -> one thread-block cluster per sequence, 8 and more -> clusters serving eight sequences each on
 * tensor-core tiles.
 * Environment (tuning / A-B only): GSV_DECODE_IMPL = hx | ll1 | cl | cl8 | gemm | barrier; GSV_HX_CS = 4 | 8 | 16; GSV_GPT_GEMM = cuda. */
int gsv_gpt_decode(gsv_gpt_ctx* ctx, int n_steps, void* stream);

/* Copy slot state to host memory (asynchronously on `stream`; caller synchronises):
 * n_gen[slot] = tokens sampled so far INCLUDING the first one (s0), active[slot] = still decoding,
 * tokens[slot*max_seq + i].  Any pointer may be NULL. */
int gsv_gpt_read(gsv_gpt_ctx* ctx, int32_t* host_n_gen, int32_t* host_active, int32_t* host_tokens,
                 int first_slot, int n_slots, void* stream);
/* Device pointers to the same state (for zero-copy consumers on the GPU). */
int gsv_gpt_state_ptrs(gsv_gpt_ctx* ctx, int32_t** dev_tokens, int32_t** dev_n_gen, int32_t** dev_active);
/* Make a slot idle (its sequence is dropped); slot = -1 releases every slot. */
int gsv_gpt_release_slot(gsv_gpt_ctx* ctx, int slot, void* stream);

/* ---- parity hooks (used by tests/; they do not change the arithmetic) ------------------- */
/* Exp(1) noise rows [n_rows][vocab] fp32 consumed one row per sample() call of slot 0, in call
 * order, instead of the in-kernel Philox stream (SURVEY.md 7 "sampling parity under a seed"). */
int gsv_gpt_set_noise(gsv_gpt_ctx* ctx, const float* dev_noise, int n_rows);
/* Teacher forcing for slot 0: its i-th sample() call (i = 0 is the first token after prefill)
 * returns forced[i] instead of its own draw, so decode step i+1 is fed forced[i]. */
int gsv_gpt_set_forced(gsv_gpt_ctx* ctx, const int32_t* dev_forced, int n);
/* Raw logits of slot 0 are appended as rows [vocab] fp32 (row 0 = prefill). */
int gsv_gpt_set_logits_trace(gsv_gpt_ctx* ctx, float* dev_rows, int max_rows);
/* The same three hooks for ANY slot, set together (NULL = off), for the request the slot is about to be given: the
 * row / forcing cursors restart at the slot's next prefill.  The update is enqueued on `stream`; order it before the
 * gsv_gpt_prefill / gsv_gpt_prefill_finish of that request (the three calls above are this one for slot 0 after a
 * device synchronisation).  With these the multi-sequence kernels are held to the oracle slot by slot. */
int gsv_gpt_set_slot_hooks(gsv_gpt_ctx* ctx, int slot, const float* dev_noise, int n_noise_rows, const int32_t* dev_forced,
                           int n_forced, float* dev_trace_rows, int max_trace_rows, void* stream);
/* Number of kernel launches issued by this context so far (bench "gpu_launches"). */
int64_t gsv_gpt_launch_count(gsv_gpt_ctx* ctx);

/* Number of SMs the single-sequence decode kernel occupies (0 = all).  No reference counterpart: the reference
 * runs the GPT chunk and the vocoder chunk of infer_stream back to back on one stream (TTS.py:402-470); here the
 * vocoder of chunk c runs on a second stream, on the SMs left free, while the GPT decodes chunk c+1.  The step
 * time of the decode kernel is the same from 128 SMs up. */
int gsv_gpt_set_decode_sms(gsv_gpt_ctx* ctx, int n_sms);
/* Holds `stream` back (a one-thread kernel, bounded to 3 ms) until every thread block of the last decode
 * launch (any of the cluster kernels) is resident.  Those kernels need whole thread-block clusters on the chip at once; a
 * second stream that keeps launching small vocoder / prefill kernels can otherwise keep them waiting for 8 or 16 free SMs
 * in one GPC for tens of milliseconds (measured: 46 -> 144 ms per utterance, at random).  Enqueue it on the other stream
 * after the decode launch. */
int gsv_gpt_wait_resident(gsv_gpt_ctx* ctx, void* stream);
/* Tuning hook: CTA `cta` of the decode kernel appends {marker id, SM clock} pairs (2 x int64 per
 * record, record 0 holds the count) to dev_records; NULL disables. */
int gsv_gpt_set_timeline(gsv_gpt_ctx* ctx, int64_t* dev_records, int max_records, int cta);

/* ======================================================================================
 * SoVITS reverse flow + HiFi-GAN generator
 * replaces SynthesizerTrn.flow_dec and the graph path of decode()
 * (reference gsv_tts/GPT_SoVITS/SoVITS/models.py:380-383, 406-425, 23-138;
 *  module/modules.py:80-104, 190-203, 482-511).
 * ====================================================================================== */

typedef struct {
  int32_t inter_channels;   /* 192 */
  int32_t hidden_channels;  /* 192 */
  int32_t gin_channels;     /* 512 (v2) / 1024 (v2Pro, v2ProPlus) */
  int32_t n_flows;          /* 4 coupling layers (models.py:31)    */
  int32_t wn_layers;        /* 4 (models.py:303)                    */
  int32_t wn_kernel;        /* 5                                    */
  int32_t upsample_initial_channel; /* 512 / 768                   */
  int32_t n_ups;            /* 5                                    */
  int32_t upsample_rates[8];
  int32_t upsample_kernel_sizes[8];
  int32_t n_resblock_kernels;       /* 3 */
  int32_t resblock_kernel_sizes[4]; /* 3,7,11 */
  int32_t resblock_dilations[4][3]; /* 1,3,5  */
  int32_t dtype;
} gsv_voc_dims;

/* Weights are passed by name, one tensor at a time, already weight-norm-folded (w = g*v/||v||,
 * SURVEY.md A.6) and converted to T by the host loader.  Layouts expected (time-major, so that
 * channels are the contiguous axis of every activation and weight tap):
 *   Conv1d           : [k][Cout][Cin]
 *   ConvTranspose1d  : [k][Cout][Cin]   (from torch's [Cin][Cout][k])
 *   bias             : [Cout]
 * Names are the reference state-dict names without the ".weight"/".bias" suffix, e.g.
 * "flow.flows.0.enc.in_layers.2", "dec.ups.3", "dec.resblocks.7.convs1.0", "dec.conv_post". */
typedef struct gsv_voc_ctx gsv_voc_ctx;

int gsv_voc_create(const gsv_voc_dims* dims, gsv_voc_ctx** out);
int gsv_voc_set_weight(gsv_voc_ctx* ctx, const char* name, const void* dev_weight, const void* dev_bias);
int gsv_voc_destroy(gsv_voc_ctx* ctx);

/* o = dec(flow(z_p, mask, ge, reverse) * mask, g=ge).  Every convolution with >= 16 input channels runs on the
 * tcgen05 implicit-GEMM kernel (csrc/conv_umma.cuh); GSV_VOC_IMPL=cuda selects the CUDA-core kernels and
 * GSV_VOC_MRF=serial / GSV_PDL=0 switch off stream-level and launch-level overlap (A-B only).
 * dev_z_p [B][192][T] T (torch layout, as handed to flow_dec), dev_mask [B][T] T,
 * dev_ge [B][gin][Tg] T with Tg == 1 or Tg == T (batched path, reference TTS.py:740-744),
 * dev_out [B][T*samples_per_frame] T.  Scratch is grown on demand and cached in ctx. */
int gsv_voc_flow_dec(gsv_voc_ctx* ctx, const void* dev_z_p, const void* dev_mask, const void* dev_ge,
                     int B, int T, int Tg, void* dev_out, void* stream);
/* Test hook: also return the flow output z [B][192][T] T (NULL to skip). */
int gsv_voc_set_debug_z(gsv_voc_ctx* ctx, void* dev_z);
int64_t gsv_voc_launch_count(gsv_voc_ctx* ctx);
/* Streaming shapes (B = 1, T <= 64) are captured into CUDA graphs on their second call and replayed afterwards (the reference
 * captures its SoVITS buckets at load, models.py:322-369); number of graphs currently instantiated (GSV_VOC_GRAPH=0 disables). */
int gsv_voc_graph_count(gsv_voc_ctx* ctx);

/* ======================================================================================
 * SoVITS prior encoder: semantic tokens -> z_p (the stage between the two hot paths)
 * replaces the front of SynthesizerTrn.decode (reference SoVITS/models.py:385-404): quantizer.decode
 * (module/core_vq.py:133-135, 222-226) + x2 nearest interpolation, ge_to512 (:396), TextEncoder.infer
 * (models.py:196-224; module/attentions.py:10-278; module/mrte_model.py:19-38) and the prior sample (:404).
 * ====================================================================================== */

typedef struct {
  int32_t hidden_channels;   /* 192 */
  int32_t filter_channels;   /* 768 */
  int32_t inter_channels;    /* 192: m_p / logs_p channels */
  int32_t n_heads;           /* 2   */
  int32_t n_layers;          /* 6: encoder_text; encoder_ssl and encoder2 have n_layers / 2 */
  int32_t kernel_size;       /* 3 (FFN convolutions) */
  int32_t ssl_dim;           /* 768: codebook width */
  int32_t n_codes;           /* 1024 codebook entries */
  int32_t n_symbols;         /* 732 phoneme symbols */
  int32_t mrte_channels;     /* 512 */
  int32_t mrte_heads;        /* 4   */
  int32_t gin_channels;      /* 512 (v2) / 1024 (v2Pro, v2ProPlus: ge goes through ge_to512 first) */
  int32_t dtype;
} gsv_encp_dims;

typedef struct gsv_encp_ctx gsv_encp_ctx;

int gsv_encp_create(const gsv_encp_dims* dims, gsv_encp_ctx** out);
/* Weights by name, already in T: linear / 1x1 convolution [out][in]; FFN convolution [k][out][in]; tables as stored.
 * Names: "quantizer.codebook" [n_codes][ssl_dim]; "enc_p.ssl_proj"; "enc_p.text_embedding" [n_symbols][C];
 * "enc_p.{encoder_ssl,encoder_text,encoder2}.attn_layers.{i}.qkv" (conv_q | conv_k | conv_v stacked on the output axis),
 * "...attn_layers.{i}.conv_o", "...attn_layers.{i}.emb_rel_k" / "emb_rel_v" [9][96], "...norm_layers_{1,2}.{i}" (weight =
 * gamma, bias = beta), "...ffn_layers.{i}.conv_{1,2}"; "enc_p.mrte.{c_pre,text_pre,c_post}",
 * "enc_p.mrte.cross_attention.{conv_q,kv,conv_o}" (kv = conv_k | conv_v stacked); "enc_p.proj"; optional "ge_to512". */
int gsv_encp_set_weight(gsv_encp_ctx* ctx, const char* name, const void* dev_weight, const void* dev_bias);
int gsv_encp_destroy(gsv_encp_ctx* ctx);
/* Frames of z_p a call produces: 2 * n_codes, minus valid_start in stream mode, then int(T / speed) + 1 if speed != 1. */
int gsv_encp_output_frames(gsv_encp_ctx* ctx, int n_codes, float speed, int stream_mode, int valid_start);
/* dev_codes [n_codes] int64, dev_text [n_text] int64, dev_ge [gin][Tg] T (torch layout; Tg == 1 or 2 * n_codes; NULL: none).
 * stream_mode / valid_start / overlap_len: the cross-fade with the previous chunk's tail, kept in the context
 * (enc_p.y_overlap, models.py:208-215; gsv_encp_reset_stream forgets it, TTS.py:498).  Text window of the MRTE cross
 * attention (mrte_model.py:26-32): dev_slices [n_slices][2] int32 rows (start, end), n_slices == 1 (one window for every
 * frame) or 2 * n_codes (one per frame: the concatenated utterances of infer_batched, TTS.py:740-747); the last text
 * position is always attended; NULL: every position.  dev_noise [inter][T'] fp32 stands for
 * randn_like(m_p) (NULL: in-kernel Philox normal keyed by `seed`).  Outputs: dev_z_p [inter][T'] T (what flow_dec takes);
 * optional dev_m_p / dev_logs_p [inter][T'] fp32; optional dev_attn [mrte_heads][2 n_codes][n_text] fp32 (what decode()
 * returns as attn, models.py:427-429); *out_frames = T'. */
int gsv_encp_forward(gsv_encp_ctx* ctx, const int64_t* dev_codes, int n_codes, const int64_t* dev_text, int n_text, const void* dev_ge,
                     int Tg, float speed, int stream_mode, int valid_start, int overlap_len, const int32_t* dev_slices, int n_slices,
                     const float* dev_noise, float noise_scale, uint64_t seed, void* dev_z_p, float* dev_m_p, float* dev_logs_p,
                     float* dev_attn, int* out_frames, void* stream);
int gsv_encp_reset_stream(gsv_encp_ctx* ctx);
/* on != 0: the NEXT gsv_encp_forward call on this context carries the same dev_text contents (and n_text) as the previous
 * one, so its text branch -- text_embedding, encoder_text, the MRTE's text_pre and k/v projections (models.py:199-204,
 * mrte_model.py:24-25), 45 of the ~100 launches of a call, independent of the codes -- is taken from that call instead of
 * recomputed (bit-identical: the same kernels on the same input).  The flag is consumed by the call; it is ignored when
 * n_text differs or a weight was set in between.  TTS.infer_phones_stream sets it for every chunk after the first. */
int gsv_encp_reuse_text(gsv_encp_ctx* ctx, int on);
/* Undo the last stream_mode call's update of the cross-chunk state (one level; the previous tail is kept in a second
 * buffer): TTS.infer_phones_stream computes a held-back chunk ahead of time and drops it when the stream ends before the
 * next chunk boundary (the reference never decodes that chunk: t2s_model.py:540-553 merges it into the final one). */
int gsv_encp_stream_rollback(gsv_encp_ctx* ctx);
int64_t gsv_encp_launch_count(gsv_encp_ctx* ctx);

/* ======================================================================================
 * Host glue of TTS.infer / infer_stream as device kernels: results stay on the device.
 * ====================================================================================== */

/* _viterbi_monotonic (reference gsv_tts/TTS.py:1744-1797): dev_attn [H][T][N] fp32 (what decode() returns) ->
 * dev_assign [T] int32: text index per frame, monotonic, -1 before the first frame whose averaged row peaks at index 0.
 * dev_work: T * N * 5 bytes of scratch. */
int gsv_glue_viterbi_monotonic(const float* dev_attn, int H, int T, int N, void* dev_work, int32_t* dev_assign, void* stream);
/* _find_head_threshold_offsets (tail = 0, TTS.py:1630-1645) / _find_tail_threshold_offsets (tail = 1, :1647-1662): the
 * first / last frame (frame_length samples every hop_length) of the first / last search_len samples whose RMS exceeds
 * `threshold`, turned into the number of samples to cut.  dev_audio [n] T; dev_work: 8 bytes; *dev_offset int32. */
int gsv_glue_silence_offset(const void* dev_audio, int n, int dtype, int tail, float threshold, int frame_length, int hop_length,
                            int search_len, int margin, void* dev_work, int32_t* dev_offset, void* stream);
/* _sola_algorithm (TTS.py:1612-1628): dev_f1_overlap [overlap_len] T (tail of the previous chunk), dev_f2 [n2] T (new
 * chunk) -> *dev_offset (best alignment in [0, search_len]) and dev_out [n2 - offset] T: the cross-faded overlap followed by
 * the rest of the aligned chunk.  dev_work: (search_len + 1) floats; dev_out must hold n2 elements. */
int gsv_glue_sola(const void* dev_f1_overlap, const void* dev_f2, int n2, int overlap_len, int search_len, int dtype, void* dev_work,
                  int32_t* dev_offset, void* dev_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GSV_B200_H */
