// Internal kernel launch interfaces of libt5gtts (not part of the C ABI).
#pragma once
#include "common.cuh"

// ---------------- GEMV family (gemv.cu) ----------------
enum { P_PLAIN = 0, P_NORM = 1, P_RES_NORM = 2, P_EMBED_NORM = 3 };
enum { E_STORE = 0, E_GEGLU = 1, E_BIAS_GELU = 2, E_BIAS = 3 };

struct GemvArgs {
  const bf16* W; int N; int K;       // W [N,K] row-major bf16 (GeGLU: rows interleaved gate_j, up_j)
  int B;                             // live batch rows (<= 4)
  const float* x;                    // P_PLAIN: [B,K]
  const float* h_in;                 // residual stream [B,K]
  const float* y;                    // sublayer output awaiting post-norm [B,K]
  const float* g_post;               // (1+w) post-norm gain
  const float* g_pre;                // (1+w) pre-norm gain
  float* h_out;                      // updated residual (written by CTA 0), may be null
  const bf16* emb; float emb_scale;  // P_EMBED_NORM: audio embedding table, sqrt(hidden)
  float eps;
  const float* bias;
  float* out; int out_stride;
  const SlotDev* slots; int slot0;   // optional activity gating / last_token source
  unsigned long long* trace;         // optional [2]: begin/end timestamps
};
cudaError_t launch_gemv(const GemvArgs& a, int P, int E, int num_sms, cudaStream_t st, bool pdl);

// ---------------- o_proj -> (grid barrier) -> cross q_proj in one kernel, bs<=4 decode (gemv_pair.cu) ----------------
struct GemvPairArgs {
  const bf16* W1; int N1, K1;        // phase A: y[B,N1] = W1 . x[B,K1]
  const float* x; float* y;
  const bf16* W2; int N2, K2;        // phase B: out[B,N2] = W2 . norm_pre(h_in + norm_post(y)),  K2 == N1
  const float* h_in; const float* g_post; const float* g_pre; float* h_out;
  float eps;
  float* out; int out_stride;
  int B; const SlotDev* slots; SlotDev* err_slots;
  unsigned long long* barrier;       // ticket counter of the grid-wide barrier (zeroed once at engine creation)
  unsigned long long* trace;
};
bool gemv_pair_supported(const GemvPairArgs& a);
cudaError_t launch_gemv_pair(const GemvPairArgs& a, int num_sms, cudaStream_t st, bool pdl);

// ---------------- decode attention over the paged KV pool (attention.cu) ----------------
struct KVPool {
  bf16* base;            // [layer][2][page][Hkv][page_tokens][D]
  int n_pages, page_tokens, Hkv, D;
  __host__ __device__ size_t page_elems() const { return (size_t)Hkv * page_tokens * D; }
  __host__ __device__ bf16* ptr(int layer, int kv, int page) const {
    return base + ((size_t)(layer * 2 + kv) * n_pages + page) * page_elems();
  }
};

struct AttnDecodeArgs {
  KVPool pool; int layer;
  const int* block_table; int bt_stride;   // [slot][bt_stride] page ids
  const float* q; int q_stride;            // [B, q_stride] raw (pre-RoPE) queries, Hq*D used
  int preload;                             // 1: issue the first K/V row loads before griddepcontrol.wait
  int mma;                                 // 1 (batched rows): tile kernel of attention_mma.cu (cp.async ring + mma.sync)
  const float* kv_new; int kv_stride;      // self: raw k at kv_new[b*kv_stride + 0..KD), v at +KD ; null for cross
  const SlotDev* slots;
  int B, Hq, Hkv, D;
  int n_splits;                            // split-KV CTAs per (request, kv head) = cluster size (1,2,4,8)
  int is_cross;                            // 1: length = n_text, no causal/window, no append
  int window;                              // >0: sliding window (self only)
  float scale, softcap; const float* inv_freq;   // inv_freq [D/2] fp32 (HF:143-145), host-computed
  const float* rope_cs;                    // optional [B][D]: cos[D/2] | sin[D/2] of the row's position (written by the sampler)
  float* out;                              // [B, Hq*D] final (normalised) attention output (fp32), or
  bf16* out_bf;                            // bf16 copy for the tensor-core o_proj of the batched path (either may be null)
  unsigned long long* trace;
  unsigned long long* probe;               // optional [B][11] in-kernel checkpoints of the (kv head 0, split 0) CTAs (debug)
  const int* row_order;                    // optional [B]: blockIdx.z -> row; longest rows first so they are scheduled first
  // chunked mode of the tile kernel (batched rows): blockIdx.y = chunk of `chunk_tokens` keys of the row's range, so the
  // work of a step is dealt in equal pieces whatever the spread of the rows' contexts; the chunks of a (row, kv head)
  // leave their partial (m, l, O) in global scratch and the last one to arrive merges them (arrival counter, self-resetting)
  int chunk_tokens;                        // 0: off (one CTA per split as above); else a multiple of 32, n_splits must be 1
  int max_chunks;                          // grid.y; chunks past a row's range exit at once (<= 16)
  float* part_o;                           // [B][Hkv][16][G][D] unnormalised partial outputs
  float* part_ml;                          // [B][Hkv][16][G][2] running max / sum of every chunk
  int* part_cnt;                           // [B][Hkv] arrival counters, zero between launches
  int n_layers_pool;                       // layers in the pool (the TMA front end addresses the whole pool as one tensor)
};
cudaError_t launch_attn_decode(const AttnDecodeArgs& a, cudaStream_t st, bool pdl);
bool attn_decode_mma_supported(const AttnDecodeArgs& a);
cudaError_t launch_attn_decode_mma(const AttnDecodeArgs& a, cudaStream_t st, bool pdl);   // attention_mma.cu
bool attn_decode_tma_supported(const AttnDecodeArgs& a);                                    // head_dim 64/128/256, 16-token pages
cudaError_t launch_attn_decode_tma(const AttnDecodeArgs& a, cudaStream_t st, bool pdl);   // attention_tma.cu (same math, TMA tile loads)

// ---------------- prefill-side kernels (prefill.cu) ----------------
// varlen packing: token t belongs to request seg_of[t]; seg_off[r]..seg_off[r+1] are its tokens
cudaError_t launch_embed(const bf16* table, const int* ids, float scale, float* h, int M, int d, cudaStream_t st);
// h_out = h_in + rmsnorm(y)*g_post (if y) ; xn = bf16(rmsnorm(h_out)*g_pre) (if xn) ; hf32 = fp32 normed (if xf)
cudaError_t launch_norm(const float* h_in, const float* y, const float* g_post, const float* g_pre, float* h_out,
                        bf16* xn, float* xf, int M, int d, float eps, cudaStream_t st, bool pdl = false,
                        float* zero_a = nullptr, int na = 0, float* zero_b = nullptr, int nb = 0,
                        unsigned long long* trace = nullptr);
// qkv fp32 [M, ld] -> RoPE(q,k) at pos[M]; q_out bf16 [M,Hq*D]; k_out/v_out bf16 [M,Hkv*D]; optional page append
struct RopeSplitArgs {
  const float* qkv; int ld; int q_off, k_off, v_off;   // column offsets (negative = absent)
  const float* pos; int M, Hq, Hkv, D; const float* inv_freq;
  bf16* q_out; bf16* k_out; bf16* v_out;
  // optional append into the paged pool
  KVPool pool; int layer; const int* block_table; int bt_stride; const int* tok_slot; const int* tok_idx;
};
cudaError_t launch_rope_split(const RopeSplitArgs& a, cudaStream_t st, bool pdl = false);
struct AttnPrefillArgs {
  const bf16* q; const bf16* k; const bf16* v;   // [Tq,Hq*D], [Tk,Hkv*D]
  const int* q_seg_off; const int* k_seg_off;    // [n_seg+1]
  const int* q_seg_of;                           // [Tq] segment of every query token
  int Tq, Hq, Hkv, D;
  int causal; int window;                        // window>0: |q-k|<=w (bidirectional) or k>q-w (causal)
  float scale, softcap;
  bf16* out;                                     // [Tq,Hq*D]
};
cudaError_t launch_attn_prefill(const AttnPrefillArgs& a, cudaStream_t st);
// tensor-core (tcgen05) prefill attention (attention_tc.cu): q/k/v as above; Tk = rows of k/v, max_lq = longest query segment
bool attn_prefill_tc_supported(int D);
cudaError_t launch_attn_prefill_tc(const AttnPrefillArgs& a, int Tk, int n_seg, int max_lq, cudaStream_t st, bool pdl = false);
// C[M,N] = A[M,K] * W[N,K]^T, bf16 inputs, fp32 accumulate.
enum { GE_F32 = 0, GE_GEGLU_BF16 = 1, GE_BIAS_GELU_BF16 = 2, GE_BIAS_F32 = 3, GE_BF16 = 4 };
struct GemmArgs {
  const bf16* A; const bf16* W; int M, N, K;
  int epilogue; const float* bias;
  void* out; int ldo;     // GE_GEGLU: N counts interleaved rows, out is [M, N/2]
  int out_zeroed;         // GE_F32 split-K: the caller guarantees `out` is already zero (no memset node is inserted)
  unsigned long long* trace = nullptr;   // optional in-step timestamps (T5G_TRACE)
  unsigned long long* probe = nullptr;   // optional [8] in-kernel checkpoints of one CTA (debug)
  // optional side job of the tcgen05 kernel's epilogue warps while they wait for the accumulator: zero up to two fp32
  // buffers (counts in floats, multiples of 4) that LATER split-K GEMMs accumulate into with red.global.add
  float* zero_a = nullptr; size_t zero_na = 0;
  float* zero_b = nullptr; size_t zero_nb = 0;
};
cudaError_t launch_gemm_simt(const GemmArgs& a, cudaStream_t st);
cudaError_t launch_gemm_tc(const GemmArgs& a, cudaStream_t st, int num_sms, bool pdl = false);   // tcgen05/TMEM/TMA path (gemm_tc.cu)
bool gemm_tc_supported(const GemmArgs& a);
bool gemm_tc_wants_zeroed_out(const GemmArgs& a, int num_sms);   // split-K with red.global.add: `out` must start at zero
// h[b] = table[slots[b].last_token] * scale for every row (batched decode step)
cudaError_t launch_embed_slots(const bf16* table, const SlotDev* slots, float scale, float* h, int B, int d, cudaStream_t st, bool pdl = false);
cudaError_t launch_gather_rows(const float* src, const int* rows, float* dst, int n, int d, cudaStream_t st);
cudaError_t launch_f32_to_bf16(const float* src, bf16* dst, size_t n, cudaStream_t st);

// ---------------- weights (weights.cu) ----------------
// dst[(row_off + r*row_mul) * cols + c] = convert(src[r*cols + c]) ; add_one: store 1+w (RMSNorm gains)
cudaError_t launch_pack(const void* src, int src_dtype, void* dst, int dst_is_bf16, int64_t rows, int64_t cols,
                        int64_t row_off, int64_t row_mul, int add_one, cudaStream_t st);

// ---------------- sampler (sampler.cu) ----------------
struct SamplerArgs {
  float* logits; int ld; int V;       // [rows, ld]
  SlotDev* slots;                     // per-row state (engine path) -- updated in place
  const int* topk_sched_pool;
  int eos, encodec_sr, text_guard;
  float progress_scale;
  int* tokens_out; int tokens_stride; // [slot][tokens_stride], entry n_generated (flat_tokens: entry 0)
  int flat_tokens;
  int* host_mirror;                   // optional mapped-host [rows][8]: active, finished, n_generated, cur_len, error flags
  float* rope_out; const float* inv_freq; int head_dim;   // optional: cos|sin table of the new position per row
  unsigned long long* trace;
  unsigned long long* scratch_u64;    // general path: [rows][2][V8] composites (V8 = V rounded up to 8), may be null
  float* scratch_f32;                 // [rows][2][V8]
  int* argmax_out;                    // optional [rows]
  int* picks_out;                     // optional [slot][tokens_stride]: engine's own sampled id per step
  const int* forced_pool;             // optional [slot][tokens_stride]: teacher-forced ids
  int rows;
  unsigned long long* probe;      // optional [12] in-kernel checkpoints of row 0 (debug)
};
cudaError_t launch_sampler(const SamplerArgs& a, cudaStream_t st, bool pdl);
