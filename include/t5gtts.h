/*
 * t5gtts.h -- C ABI of libt5gtts.so, the B200-native (sm_100a) engine for the T5Gemma-TTS
 * token-generation hot path.
 *
 * The reference (tori29umai0123/T5Gemma-TTS) is pure Python and has no FFI; the "binding" a
 * maintainer adds is a ctypes stub.  Each entry point below names the reference
 * code it replaces (paths relative to the reference root; "HF:" = transformers/models/t5gemma/
 * modeling_t5gemma.py, the third-party module the reference subclasses).
 *
 * Conventions: plain C, no C++/torch types.  Every function returns 0 on success or a negative
 * T5G_ERR_* code; t5g_last_error() returns a human-readable message for the calling thread's last
 * failure.  Nothing throws across the ABI.  One engine per device; an engine is not thread-safe.
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  The engine owns its
 * packed weights, KV pages and workspaces (one cudaMalloc arena created in t5g_create /
 * t5g_finalize_weights); it borrows caller pointers only for the duration of a call unless stated.
 */
#ifndef T5GTTS_H
#define T5GTTS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define T5G_ABI_VERSION 1
#define T5G_MAX_LAYERS 64

enum { T5G_OK = 0, T5G_ERR_INVALID = -1, T5G_ERR_CUDA = -2, T5G_ERR_OOM = -3, T5G_ERR_STATE = -4,
       T5G_ERR_UNSUPPORTED = -5 };

enum { T5G_F32 = 0, T5G_BF16 = 1, T5G_F16 = 2 };

/* Model + engine geometry.  Mirrors T5GemmaVoiceConfig (hf_export/configuration_t5gemma_voice.py:50-151)
 * and the backbone's T5GemmaModuleConfig (HF:configuration_t5gemma.py:68-99). */
typedef struct T5GConfig {
  int32_t abi_version;          /* must be T5G_ABI_VERSION */
  int32_t hidden;               /* 2304 */
  int32_t inter;                /* 9216 */
  int32_t n_enc_layers;         /* 26 */
  int32_t n_dec_layers;         /* 26 */
  int32_t n_heads;              /* 8 */
  int32_t n_kv_heads;           /* 4 */
  int32_t head_dim;             /* 256 */
  int32_t sliding_window;       /* 4096 */
  int32_t text_vocab;           /* 256000 */
  int32_t n_audio_tokens;       /* audio_vocab_size + n_special = 65541 */
  int32_t eos_token;            /* models/t5gemma.py:861-863: eos if eos>0 else eog */
  int32_t encodec_sr;           /* 50 */
  int32_t text_guard_frames_per_token; /* 0 = off (models/t5gemma.py:1026-1040) */
  float   attn_scale;           /* query_pre_attn_scalar^-0.5 */
  float   attn_softcap;         /* attn_logit_softcapping under "eager"; 0 = none ("sdpa" drops it) */
  float   rms_eps;              /* 1e-6 */
  float   rope_theta;           /* 10000 */
  float   progress_scale;       /* 2000 (PM-RoPE) */
  float   extra_cutoff;         /* 5.0 s (time-budget stop rule, models/t5gemma.py:1042-1046) */
  uint8_t enc_layer_sliding[T5G_MAX_LAYERS]; /* 1 = sliding_attention layer */
  uint8_t dec_layer_sliding[T5G_MAX_LAYERS];
  /* engine sizing */
  int32_t max_slots;            /* concurrent requests (decode batch rows) */
  int32_t max_text_len;         /* per request encoder tokens */
  int32_t max_dec_len;          /* per request decoder tokens (BOS+prompt+generated) */
  int32_t max_prefill_tokens;   /* tokens per prefill call (sum over requests, max(enc, dec)) */
  int32_t kv_page_tokens;       /* KV page size in tokens (multiple of 4; 32 = default, 16 or 32 for the TMA attention front end) */
  int32_t reserved0;
} T5GConfig;

typedef struct T5GEngine T5GEngine;

/* Sampling parameters of one request: the kwargs of inference_tts (models/t5gemma.py:835-850). */
typedef struct T5GSampling {
  int32_t top_k;                /* <=0 disables (models/utils.py:82) */
  float   top_p;                /* >=1 disables */
  float   min_p;                /* (0,1) enables and bypasses k/p (models/utils.py:72-80) */
  float   temperature;
} T5GSampling;

/* One request handed to t5g_prefill. */
typedef struct T5GRequest {
  int32_t slot;                 /* engine row this request occupies, 0 <= slot < max_slots */
  int32_t n_text;               /* x_lens */
  const int64_t* text_ids;      /* HOST pointer, n_text ids (x) */
  int32_t n_dec;                /* 1 + prompt length: BOS ++ y (models/t5gemma.py:902-908) */
  const int64_t* dec_ids;       /* HOST pointer, n_dec audio ids */
  int32_t target_total;         /* tgt_y_lens; < 0 = None: est_total from the 2 s lookahead, no time budget
                                 * (models/t5gemma.py:896-933); eos is then forced when the slot's max_dec_len is reached */
  int32_t prompt_frames;        /* kwargs["prompt_frames"] */
  int32_t max_new_tokens;       /* 0 = reference behaviour (stop rules only) */
  T5GSampling sampling;
  const int32_t* top_k_schedule;/* HOST pointer or NULL: top_k as a per-step list (models/t5gemma.py:991-994) */
  int32_t n_top_k_schedule;
  const float* uniforms;        /* DEVICE pointer, n_uniforms fp32 draws in [0,1); borrowed until the slot is released */
  int32_t n_uniforms;
  /* Teacher forcing (parity tests): HOST pointer or NULL.  When set, step i feeds forced_tokens[i] to the
   * decoder instead of the engine's own pick; the pick is still recorded (t5g_read_picks). */
  const int32_t* forced_tokens;
  int32_t n_forced;
  /* Silence-repetition penalty (models/t5gemma.py:999-1011, 1050-1054): HOST pointer to the silence token ids
   * (NULL/0 = off, what every shipped caller passes) and the stop_repetition argument of inference_tts. */
  const int32_t* silence_tokens;
  int32_t n_silence;
  int32_t stop_repetition;
} T5GRequest;

/* Host-visible per-slot state after t5g_decode / t5g_poll. */
typedef struct T5GSlotState {
  int32_t active;               /* 1 while generating */
  int32_t finished;             /* 1 once eos was emitted */
  int32_t n_generated;          /* tokens emitted so far (incl. the final eos) */
  int32_t cur_len;              /* decoder length incl. BOS+prompt */
} T5GSlotState;

/* Lifecycle.  Replaces model construction: T5GemmaVoiceModel.__init__ (models/t5gemma.py:272-418) /
 * T5GemmaVoiceForConditionalGeneration.__init__ (hf_export/modeling_t5gemma_voice.py:343-479). */
int t5g_create(const T5GConfig* cfg, int device, T5GEngine** out);
int t5g_destroy(T5GEngine* eng);

/* Weight loading.  Replaces load_state_dict (models/t5gemma.py:1131-1141, inference_commandline.py:
 * 131-156) / from_pretrained.  `name` is the reference state_dict key, e.g.
 * "backbone.model.decoder.layers.3.cross_attn.q_proj.weight", "audio_embedding.0.weight",
 * "predict_layer.0.2.bias".  `data` may be a host or device pointer (on_device flag); it is
 * converted/packed into the engine's own bf16/fp32 layout before the call returns control of
 * `data` to the caller (the copy is enqueued on the engine's internal stream and synchronised). */
int t5g_load_tensor(T5GEngine* eng, const char* name, const void* data, int dtype, int ndim,
                    const int64_t* shape, int on_device);
int t5g_finalize_weights(T5GEngine* eng);   /* fails listing the first missing tensor */

/* Prefill: encoder pass (a4), cross-attention K/V precompute with PM-RoPE (a8), decoder pass over
 * BOS+prompt (a5-a7) for n_req requests at once (varlen-packed).  Replaces models/t5gemma.py:867-963.
 * Afterwards every slot holds its last hidden state and is ready for t5g_decode. */
int t5g_prefill(T5GEngine* eng, const T5GRequest* reqs, int n_req, void* stream);

/* Decode: runs up to `max_steps` iterations of the hot loop (models/t5gemma.py:1057-1115) for all
 * active slots -- head (a13) -> sample + stop rules (S) -> embed -> 26 decoder layers -> final norm --
 * each iteration one CUDA-graph replay, no host sync in between.  Returns after enqueueing; call
 * t5g_poll to synchronise and read the slot states. */
int t5g_decode(T5GEngine* eng, int max_steps, void* stream);
int t5g_poll(T5GEngine* eng, T5GSlotState* states /* [max_slots] host */, void* stream);

/* Copies the generated tokens of a slot (int32, n_generated entries incl. final eos) to host. */
int t5g_read_tokens(T5GEngine* eng, int slot, int32_t* out, int max_tokens, int* n_out, void* stream);
/* The engine's own sampled ids per step, before teacher forcing and stop rules (== tokens when not forcing). */
int t5g_read_picks(T5GEngine* eng, int slot, int32_t* out, int max_tokens, int* n_out, void* stream);
int t5g_release_slot(T5GEngine* eng, int slot);

/* Introspection used by the parity tests (device->host copies, synchronous on `stream`). */
int t5g_read_memory(T5GEngine* eng, int slot, float* out /* [n_text, hidden] */, void* stream);
int t5g_read_last_hidden(T5GEngine* eng, int slot, float* out /* [hidden] */, void* stream);
int t5g_read_logits(T5GEngine* eng, int slot, float* out /* [n_audio_tokens], logits of the last step */, void* stream);
/* Teacher-forced logits for all n_dec positions of the last t5g_prefill of this slot
 * (decoder states -> predict_layer; the oracle is models/t5gemma.py:666-833 `forward`). */
int t5g_prefill_logits(T5GEngine* eng, int slot, float* out /* [n_dec, n_audio_tokens] */, void* stream);

/* Standalone sampling step S (models/t5gemma.py:971-1055 + models/utils.py:53-122) on caller logits:
 * DEVICE logits [n_rows, n_audio_tokens] fp32 (edited in place like the reference), per-row params. */
typedef struct T5GSampleRow {
  T5GSampling sampling;
  float   u;                    /* uniform draw in [0,1) */
  int32_t cur_num_gen;
  int32_t current_length;
  int32_t prompt_offset;        /* prompt_frames + 1 */
  int32_t target_total;
  int32_t n_text;
  int32_t prev_token;           /* silence-repetition state (models/t5gemma.py:967-968): -1 initially */
  int32_t consec_silence_count;
} T5GSampleRow;
int t5g_sample(T5GEngine* eng, float* logits, const T5GSampleRow* rows /* host */, int n_rows,
               int32_t* out_tokens /* host */, int32_t* out_argmax /* host, may be NULL */, void* stream);
/* Silence-repetition settings used by t5g_sample rows (shared by all rows of the call; NULL/0 = off). */
int t5g_sample_set_silence(T5GEngine* eng, const int32_t* silence_tokens /* host */, int n_silence, int stop_repetition);

/* Accounting for bench.py: kernel launches issued by the engine since creation, and per-step byte model. */
int64_t t5g_launch_count(const T5GEngine* eng);
int64_t t5g_weight_bytes_per_step(const T5GEngine* eng);
int64_t t5g_kv_bytes_per_token(const T5GEngine* eng);
/* Per-phase device timings of the last calls, in milliseconds (CUDA events on the caller's stream):
 * [0]=encoder prefill [1]=cross-KV [2]=decoder prefill [3]=last t5g_decode call. */
int t5g_get_timings(T5GEngine* eng, float* out_ms4);

/* Running totals since t5g_create (CUDA-event time of every t5g_prefill call and of every t5g_decode call that was
 * followed by t5g_poll): [0] prefill ms, [1] decode ms, [2] decode steps enqueued, [3] prefill calls, [4] kernel launches,
 * [5] kernels per decode step of the last call, [6..7] reserved. */
int t5g_get_counters(T5GEngine* eng, double* out8);

const char* t5g_last_error(void);
int t5g_abi_version(void);

/* Kernel-level entry points used by tests/ and the roofline micro-benchmarks. */
/* With T5G_TRACE=1 in the environment at t5g_create, every kernel of the decode step records
 * (min over CTAs of the time after its dependency wait, max over CTAs of its exit time) in globaltimer ns;
 * this returns the records of the last executed step in launch order. */
int t5g_debug_trace(T5GEngine* eng, uint64_t* begin_ns, uint64_t* end_ns, int max_entries, int* n_out);
int t5g_debug_gemm(T5GEngine* eng, const void* x_bf16 /* dev [M,K] */, const void* w_bf16 /* dev [N,K] */,
                   float* out /* dev [M,N] */, int M, int N, int K,
                   int impl /* low byte: 0 = simt, 1 = tcgen05; bits 8-15: epilogue (0 fp32 [M,N], 4 bf16 [M,N], 1 GeGLU bf16 [M,N/2]) */,
                   void* stream);
/* Prefill attention kernel alone (impl 0 = CUDA-core kernel, 1 = tcgen05 kernel) on caller device buffers: q [Tq, Hq*D],
 * k/v [Tk, Hkv*D] bf16, varlen segments (device int32 arrays: q_seg_off/k_seg_off [n_seg+1], q_seg_of [Tq]), out bf16
 * [Tq, Hq*D].  Geometry (heads, head_dim, scale) comes from the engine config. */
int t5g_debug_attn_prefill(T5GEngine* eng, const void* q, const void* k, const void* v, const int32_t* q_seg_off,
                           const int32_t* k_seg_off, const int32_t* q_seg_of, int n_seg, int Tq, int Tk, int max_lq,
                           int causal, int window, float softcap, void* out, int impl, void* stream);
/* Launches the decode step's dominant kernel (gate|up projection: post-norm + residual + pre-norm prologue, GeGLU
 * epilogue; gemv_kernel<1,P_RES_NORM,E_GEGLU>) on caller-provided interleaved weights [2*inter, hidden] (device,
 * bf16) using the engine's own decode buffers of row 0.  Used by bench.py to time that kernel live. */
int t5g_debug_gemv_gateup(T5GEngine* eng, const void* w_bf16, void* stream);
int t5g_debug_gemv(T5GEngine* eng, const float* x /* dev [B,K] */, const void* w_bf16 /* dev [N,K] */,
                   float* out /* dev [B,N] */, int B, int N, int K, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* T5GTTS_H */
