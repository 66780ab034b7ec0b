// libt5gtts engine: owns packed weights, the paged KV pool and all workspaces; orchestrates the
// prefill (encoder -> cross K/V -> decoder over BOS+prompt) and the CUDA-graph decode step.
// C ABI in include/t5gtts.h.  Host control flow mirrors models/t5gemma.py:835-1129 (inference_tts).
#include "../../include/t5gtts.h"
#include "kernels.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <map>
#include <set>
#include <algorithm>

// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void t5g_set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
extern "C" const char* t5g_last_error(void) { return g_err; }
extern "C" int t5g_abi_version(void) { return T5G_ABI_VERSION; }

namespace {

struct EncLayer { bf16 *wqkv, *wo, *wgu, *wd; float *g_pre_sa, *g_post_sa, *g_pre_ff, *g_post_ff; };
struct DecLayer {
  bf16 *wqkv, *wo, *wq_c, *wkv_c, *wo_c, *wgu, *wd;
  float *g_pre_sa, *g_post_sa, *g_pre_ca, *g_post_ca, *g_pre_ff, *g_post_ff;
};

struct HostSlot {
  bool in_use = false;
  std::vector<int> self_pages, cross_pages;
  int n_text = 0, n_dec = 0;
  int mem_off = -1;      // token offset of this slot's encoder output in `memory` (last prefill)
  int dec_off = -1;      // token offset of this slot's decoder states in `dec_final`
  uint64_t prefill_id = 0;
};

}  // namespace

struct T5GEngine {
  T5GConfig c;
  int device = 0, num_sms = 148;
  int d, I, Hq, Hkv, D, QD, KD, QKV, V, Vpad, PT;
  int max_self_pages, max_cross_pages, n_pages;
  bool use_pdl = true, use_graph = true;
  int gemm_impl = 1;                       // 1 tcgen05/TMEM/TMA (default), 0 = SIMT cross-check kernel
  cudaStream_t load_stream = nullptr;
  std::vector<void*> allocs;               // every cudaMalloc of this engine
  size_t bytes_allocated = 0;
  // weights
  bf16* enc_embed = nullptr; float* g_enc_final = nullptr; float* g_dec_final = nullptr;
  std::vector<EncLayer> enc; std::vector<DecLayer> dec;
  bf16* audio_emb = nullptr; bf16* head_w1 = nullptr; float* head_b1 = nullptr; bf16* head_w2 = nullptr; float* head_b2 = nullptr;
  float* inv_freq = nullptr;
  std::set<std::string> required, loaded;
  bool finalized = false;
  // kv
  KVPool pool{}; std::vector<int> free_pages;
  int* d_self_bt = nullptr; int* d_cross_bt = nullptr;
  std::vector<HostSlot> hslots;
  // slot state
  SlotDev* d_slots = nullptr; std::vector<SlotDev> h_slots;   // host shadow used when (re)initialising
  int* h_mirror = nullptr; int* d_mirror = nullptr;           // mapped pinned [max_slots][8]: active, finished, n_generated, cur_len, error flags
  int* h_tokens = nullptr; int* d_tokens = nullptr;           // mapped pinned [max_slots][max_dec_len]
  int* h_picks = nullptr; int* d_picks = nullptr;             // mapped pinned [max_slots][max_dec_len]
  int* d_forced = nullptr;                                    // [max_slots][max_dec_len]
  int* d_topk_pool = nullptr; int topk_pool_cap = 0;   // int pool [max_slots][max_dec_len]: per-slot top-k schedule + silence token list
  unsigned long long* d_samp_u64 = nullptr; float* d_samp_f32 = nullptr; int samp_scratch_rows = 0;   // sampler general path
  int* d_sample_silence = nullptr; int n_sample_silence = 0, sample_stop_repetition = 0;   // t5g_sample settings
  // prefill workspaces (T = max_prefill_tokens)
  float *p_h = nullptr, *p_y = nullptr, *p_qkv = nullptr, *p_memory = nullptr, *p_ckv = nullptr, *p_final = nullptr;
  bf16 *p_xn = nullptr, *p_q = nullptr, *p_k = nullptr, *p_v = nullptr, *p_att = nullptr, *p_act = nullptr,
       *p_mem_bf = nullptr, *p_ck = nullptr, *p_cv = nullptr;
  bool use_tc_attn = true; int attn_mma = 1, attn_tma = 1;
  float* p_logits = nullptr; int logits_chunk = 128;
  int *p_ids = nullptr, *p_seg_of = nullptr, *p_seg_off_e = nullptr, *p_seg_off_d = nullptr, *p_tok_slot = nullptr,
      *p_tok_idx = nullptr, *p_last_rows = nullptr;
  float* p_pos = nullptr;
  void* h_stage = nullptr; size_t h_stage_bytes = 0;          // pinned staging for prefill metadata
  uint64_t prefill_counter = 0;
  // decode workspaces (B = max_slots)
  float *d_hA = nullptr, *d_hB = nullptr, *d_y = nullptr, *d_qkv = nullptr, *d_qc = nullptr, *d_act = nullptr,
        *d_t1 = nullptr, *d_logits = nullptr, *d_rope = nullptr;
  float* d_sample_u = nullptr; SlotDev* d_sample_slots = nullptr;   // t5g_sample scratch
  float* d_attn = nullptr;                                          // attention output [B,QD]
  bf16 *d_xn = nullptr, *d_attn_bf = nullptr, *d_act_bf = nullptr, *d_t1_bf = nullptr;   // batched (tensor-core) decode step
  unsigned long long* d_trace = nullptr; bool use_trace = false;    // [2][T5G_TRACE_STRIDE] begin/end timestamps
  int ns_self = 8, ns_cross = 2;
  int h_end = 0;                                               // which h buffer holds the residual at step end
  // decode graphs, keyed by (steps per graph, self-attention chunks, cross-attention chunks): the chunk counts are grid
  // dimensions of the batched attention kernels and follow the longest live row
  struct StepGraph { cudaGraphExec_t exec = nullptr; int nodes = 0; };
  std::map<int, StepGraph> graphs; int graph_steps = 4;
  int attn_chunk = 384, chunks_self = 1, chunks_cross = 1;    // batched rows: keys per attention CTA, current grid.y
  float *d_part_o = nullptr, *d_part_ml = nullptr; int* d_part_cnt = nullptr;
  int last_nodes_per_step = 0;
  int *d_order_self = nullptr, *d_order_cross = nullptr;                 // batched attention: rows by descending length
  unsigned long long* d_barrier = nullptr; bool use_pair = true;         // o_proj + cross q_proj in one kernel (gemv_pair.cu)
  int64_t launches = 0;
  cudaEvent_t ev[6] = {};
  float timings[4] = {0, 0, 0, 0};
  double tot_prefill_ms = 0, tot_decode_ms = 0; int64_t tot_decode_steps = 0, tot_prefill_calls = 0;   // t5g_get_counters
  bool decode_pending = false; int pending_steps = 0;
};

namespace {

int dmalloc(T5GEngine* e, void** p, size_t bytes) {
  if (bytes == 0) bytes = 16;
  cudaError_t err = cudaMalloc(p, bytes);
  if (err != cudaSuccess) { t5g_set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(err)); return T5G_ERR_OOM; }
  e->allocs.push_back(*p);
  e->bytes_allocated += bytes;
  return T5G_OK;
}
#define DM(ptr, count) do { int _r = dmalloc(e, (void**)&(ptr), sizeof(*(ptr)) * (size_t)(count)); if (_r) return _r; } while (0)

std::string lname(const char* side, int l, const char* rest) {
  char buf[256]; snprintf(buf, sizeof(buf), "backbone.model.%s.layers.%d.%s", side, l, rest); return buf;
}

struct Dest { void* ptr; bool is_bf16; int64_t rows, cols, row_off, row_mul; bool add_one; int64_t dst_rows; };

// maps a reference state_dict key to its packed destination
bool resolve(T5GEngine* e, const std::string& name, Dest* out) {
  auto mk = [&](void* p, bool bf, int64_t r, int64_t c, int64_t off = 0, int64_t mul = 1, bool one = false) {
    *out = Dest{p, bf, r, c, off, mul, one, 0}; return true;
  };
  const int d = e->d, I = e->I, QD = e->QD, KD = e->KD;
  if (name == "backbone.model.encoder.embed_tokens.weight") return mk(e->enc_embed, true, e->c.text_vocab, d);
  if (name == "backbone.model.encoder.norm.weight") return mk(e->g_enc_final, false, 1, d, 0, 1, true);
  if (name == "backbone.model.decoder.norm.weight") return mk(e->g_dec_final, false, 1, d, 0, 1, true);
  if (name == "audio_embedding.0.weight") return mk(e->audio_emb, true, e->V, d);
  if (name == "predict_layer.0.0.weight") return mk(e->head_w1, true, d, d);
  if (name == "predict_layer.0.0.bias") return mk(e->head_b1, false, 1, d);
  if (name == "predict_layer.0.2.weight") return mk(e->head_w2, true, e->V, d);
  if (name == "predict_layer.0.2.bias") return mk(e->head_b2, false, 1, e->V);
  int l = -1; char rest[128];
  bool is_enc = sscanf(name.c_str(), "backbone.model.encoder.layers.%d.%127s", &l, rest) == 2;
  bool is_dec = !is_enc && sscanf(name.c_str(), "backbone.model.decoder.layers.%d.%127s", &l, rest) == 2;
  if (!is_enc && !is_dec) return false;
  if (l < 0 || l >= (is_enc ? e->c.n_enc_layers : e->c.n_dec_layers)) return false;
  const std::string r = rest;
  bf16 *wqkv, *wo, *wgu, *wd; float *pre_sa, *post_sa, *pre_ff, *post_ff;
  if (is_enc) { EncLayer& L = e->enc[l]; wqkv = L.wqkv; wo = L.wo; wgu = L.wgu; wd = L.wd; pre_sa = L.g_pre_sa; post_sa = L.g_post_sa; pre_ff = L.g_pre_ff; post_ff = L.g_post_ff; }
  else { DecLayer& L = e->dec[l]; wqkv = L.wqkv; wo = L.wo; wgu = L.wgu; wd = L.wd; pre_sa = L.g_pre_sa; post_sa = L.g_post_sa; pre_ff = L.g_pre_ff; post_ff = L.g_post_ff; }
  if (r == "self_attn.q_proj.weight") return mk(wqkv, true, QD, d, 0);
  if (r == "self_attn.k_proj.weight") return mk(wqkv, true, KD, d, QD);
  if (r == "self_attn.v_proj.weight") return mk(wqkv, true, KD, d, QD + KD);
  if (r == "self_attn.o_proj.weight") return mk(wo, true, d, QD);
  if (r == "mlp.gate_proj.weight") return mk(wgu, true, I, d, 0, 2);
  if (r == "mlp.up_proj.weight") return mk(wgu, true, I, d, 1, 2);
  if (r == "mlp.down_proj.weight") return mk(wd, true, d, I);
  if (r == "pre_self_attn_layernorm.weight") return mk(pre_sa, false, 1, d, 0, 1, true);
  if (r == "post_self_attn_layernorm.weight") return mk(post_sa, false, 1, d, 0, 1, true);
  if (r == "pre_feedforward_layernorm.weight") return mk(pre_ff, false, 1, d, 0, 1, true);
  if (r == "post_feedforward_layernorm.weight") return mk(post_ff, false, 1, d, 0, 1, true);
  if (is_dec) {
    DecLayer& L = e->dec[l];
    if (r == "cross_attn.q_proj.weight") return mk(L.wq_c, true, QD, d);
    if (r == "cross_attn.k_proj.weight") return mk(L.wkv_c, true, KD, d, 0);
    if (r == "cross_attn.v_proj.weight") return mk(L.wkv_c, true, KD, d, KD);
    if (r == "cross_attn.o_proj.weight") return mk(L.wo_c, true, d, QD);
    if (r == "pre_cross_attn_layernorm.weight") return mk(L.g_pre_ca, false, 1, d, 0, 1, true);
    if (r == "post_cross_attn_layernorm.weight") return mk(L.g_post_ca, false, 1, d, 0, 1, true);
  }
  return false;
}

void add_required(T5GEngine* e) {
  auto& R = e->required;
  R.insert("backbone.model.encoder.embed_tokens.weight");
  R.insert("backbone.model.encoder.norm.weight");
  R.insert("backbone.model.decoder.norm.weight");
  R.insert("audio_embedding.0.weight");
  R.insert("predict_layer.0.0.weight"); R.insert("predict_layer.0.0.bias");
  R.insert("predict_layer.0.2.weight"); R.insert("predict_layer.0.2.bias");
  const char* common[] = {"self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight",
                          "self_attn.o_proj.weight", "mlp.gate_proj.weight", "mlp.up_proj.weight", "mlp.down_proj.weight",
                          "pre_self_attn_layernorm.weight", "post_self_attn_layernorm.weight",
                          "pre_feedforward_layernorm.weight", "post_feedforward_layernorm.weight"};
  const char* cross[] = {"cross_attn.q_proj.weight", "cross_attn.k_proj.weight", "cross_attn.v_proj.weight",
                         "cross_attn.o_proj.weight", "pre_cross_attn_layernorm.weight", "post_cross_attn_layernorm.weight"};
  for (int l = 0; l < e->c.n_enc_layers; ++l) for (auto s : common) R.insert(lname("encoder", l, s));
  for (int l = 0; l < e->c.n_dec_layers; ++l) {
    for (auto s : common) R.insert(lname("decoder", l, s));
    for (auto s : cross) R.insert(lname("decoder", l, s));
  }
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- GEMM dispatch (prefill / batched path) ----------------------------------------------------
cudaError_t gemm(T5GEngine* e, const bf16* A, const bf16* W, int M, int N, int K, int epi, const float* bias,
                 void* out, int ldo, cudaStream_t st, int out_zeroed = 0) {
  GemmArgs g{A, W, M, N, K, epi, bias, out, ldo, out_zeroed};
  e->launches++;
  // programmatic dependent launch in the prefill chain as well: the GEMM streams its first weight tiles while the previous
  // kernel drains (kernels that are not PDL-aware simply complete first)
  if (e->gemm_impl == 1 && gemm_tc_supported(g)) return launch_gemm_tc(g, st, e->num_sms, e->use_pdl);
  return launch_gemm_simt(g, st);
}

// true when the fp32 GEMM [M,N] = A[M,K] W^T will run split-K with red.global.add: the kernel before it zeroes the output
// (norm_kernel's side job) so that no memset node interrupts the programmatic-dependent-launch chain of a small prefill
bool wants_zero(T5GEngine* e, int M, int N, int K) {
  GemmArgs g{e->p_xn, e->p_xn, M, N, K, GE_F32, nullptr, e->p_y, N, 0};
  return e->gemm_impl == 1 && gemm_tc_wants_zeroed_out(g, e->num_sms);
}

#define CU(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { t5g_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); return T5G_ERR_CUDA; } } while (0)

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" int t5g_create(const T5GConfig* cfg, int device, T5GEngine** out) {
  T5G_CHECK(cfg && out, T5G_ERR_INVALID, "null argument");
  T5G_CHECK(cfg->abi_version == T5G_ABI_VERSION, T5G_ERR_INVALID, "ABI version mismatch: %d vs %d", cfg->abi_version, T5G_ABI_VERSION);
  T5G_CHECK(cfg->n_enc_layers > 0 && cfg->n_enc_layers <= T5G_MAX_LAYERS && cfg->n_dec_layers > 0 && cfg->n_dec_layers <= T5G_MAX_LAYERS,
            T5G_ERR_INVALID, "layer count out of range");
  T5G_CHECK(cfg->n_heads % cfg->n_kv_heads == 0, T5G_ERR_INVALID, "n_heads must be a multiple of n_kv_heads");
  const int G = cfg->n_heads / cfg->n_kv_heads, D = cfg->head_dim;
  T5G_CHECK((G == 1 || G == 2 || G == 4) && (D == 16 || D == 32 || D == 64 || D == 128 || D == 256) && !(G == 4 && D == 256),
            T5G_ERR_UNSUPPORTED, "unsupported attention geometry G=%d D=%d", G, D);
  T5G_CHECK(cfg->hidden % 8 == 0 && cfg->inter % 8 == 0 && cfg->hidden <= 4096, T5G_ERR_UNSUPPORTED, "hidden/inter must be multiples of 8, hidden<=4096");
  T5G_CHECK(cfg->max_slots >= 1 && cfg->max_text_len >= 1 && cfg->max_dec_len >= 2 && cfg->max_prefill_tokens >= 1,
            T5G_ERR_INVALID, "bad engine sizing");
  T5G_CHECK(cfg->kv_page_tokens > 0 && cfg->kv_page_tokens % 4 == 0, T5G_ERR_INVALID, "kv_page_tokens must be a positive multiple of 4");
  T5G_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  T5G_CUDA(cudaGetDeviceProperties(&prop, device));
  T5G_CHECK(prop.major == 10, T5G_ERR_UNSUPPORTED, "libt5gtts requires an sm_100 (B200) device, found sm_%d%d", prop.major, prop.minor);

  T5GEngine* e = new T5GEngine();
  e->c = *cfg; e->device = device; e->num_sms = prop.multiProcessorCount;
  e->d = cfg->hidden; e->I = cfg->inter; e->Hq = cfg->n_heads; e->Hkv = cfg->n_kv_heads; e->D = D;
  e->QD = e->Hq * D; e->KD = e->Hkv * D; e->QKV = e->QD + 2 * e->KD;
  e->V = cfg->n_audio_tokens; e->Vpad = cdiv(e->V, 64) * 64; e->PT = cfg->kv_page_tokens;
  if (const char* s = getenv("T5G_PDL")) e->use_pdl = atoi(s) != 0;
  if (const char* s = getenv("T5G_GRAPH")) e->use_graph = atoi(s) != 0;
  if (const char* s = getenv("T5G_GRAPH_STEPS")) e->graph_steps = atoi(s);
  if (const char* s = getenv("T5G_GEMM")) e->gemm_impl = atoi(s);
  if (const char* s = getenv("T5G_TRACE")) e->use_trace = atoi(s) != 0;
  *out = e;   // so that the caller can destroy on failure
  T5G_CUDA(cudaStreamCreateWithFlags(&e->load_stream, cudaStreamNonBlocking));
  for (auto& ev : e->ev) T5G_CUDA(cudaEventCreate(&ev));

  const int d = e->d, I = e->I, QD = e->QD, KD = e->KD, QKV = e->QKV;
  // ---- weights ----
  DM(e->enc_embed, (size_t)cfg->text_vocab * d);
  DM(e->g_enc_final, d); DM(e->g_dec_final, d);
  e->enc.resize(cfg->n_enc_layers); e->dec.resize(cfg->n_dec_layers);
  for (auto& L : e->enc) {
    DM(L.wqkv, (size_t)QKV * d); DM(L.wo, (size_t)d * QD); DM(L.wgu, (size_t)2 * I * d); DM(L.wd, (size_t)d * I);
    DM(L.g_pre_sa, d); DM(L.g_post_sa, d); DM(L.g_pre_ff, d); DM(L.g_post_ff, d);
  }
  for (auto& L : e->dec) {
    DM(L.wqkv, (size_t)QKV * d); DM(L.wo, (size_t)d * QD); DM(L.wq_c, (size_t)QD * d); DM(L.wkv_c, (size_t)2 * KD * d);
    DM(L.wo_c, (size_t)d * QD); DM(L.wgu, (size_t)2 * I * d); DM(L.wd, (size_t)d * I);
    DM(L.g_pre_sa, d); DM(L.g_post_sa, d); DM(L.g_pre_ca, d); DM(L.g_post_ca, d); DM(L.g_pre_ff, d); DM(L.g_post_ff, d);
  }
  DM(e->audio_emb, (size_t)e->V * d);
  DM(e->head_w1, (size_t)d * d); DM(e->head_b1, d);
  DM(e->head_w2, (size_t)e->Vpad * d); DM(e->head_b2, e->Vpad);
  T5G_CUDA(cudaMemset(e->head_w2, 0, sizeof(bf16) * (size_t)e->Vpad * d));
  T5G_CUDA(cudaMemset(e->head_b2, 0, sizeof(float) * e->Vpad));
  DM(e->inv_freq, D / 2);
  {
    std::vector<float> f(D / 2);
    // HF:143-145: 1/(theta^(2i/D)) in fp32
    for (int i = 0; i < D / 2; ++i) f[i] = (float)(1.0 / pow((double)cfg->rope_theta, (double)(2 * i) / (double)D));
    T5G_CUDA(cudaMemcpy(e->inv_freq, f.data(), sizeof(float) * f.size(), cudaMemcpyHostToDevice));
  }
  add_required(e);

  // ---- KV pool ----
  const int B = cfg->max_slots;
  e->max_self_pages = cdiv(cfg->max_dec_len, e->PT);
  e->max_cross_pages = cdiv(cfg->max_text_len, e->PT);
  e->n_pages = B * (e->max_self_pages + e->max_cross_pages);
  e->pool.n_pages = e->n_pages; e->pool.page_tokens = e->PT; e->pool.Hkv = e->Hkv; e->pool.D = D;
  DM(e->pool.base, (size_t)cfg->n_dec_layers * 2 * e->n_pages * e->pool.page_elems());
  // the TMA front end of the batched attention loads whole pages: rows that were never written must hold finite values
  T5G_CUDA(cudaMemset(e->pool.base, 0, sizeof(bf16) * (size_t)cfg->n_dec_layers * 2 * e->n_pages * e->pool.page_elems()));
  for (int p = e->n_pages - 1; p >= 0; --p) e->free_pages.push_back(p);
  DM(e->d_self_bt, (size_t)B * e->max_self_pages); DM(e->d_cross_bt, (size_t)B * e->max_cross_pages);
  T5G_CUDA(cudaMemset(e->d_self_bt, 0, sizeof(int) * (size_t)B * e->max_self_pages));
  T5G_CUDA(cudaMemset(e->d_cross_bt, 0, sizeof(int) * (size_t)B * e->max_cross_pages));
  e->hslots.resize(B); e->h_slots.resize(B);
  memset(e->h_slots.data(), 0, sizeof(SlotDev) * B);
  DM(e->d_slots, B);
  T5G_CUDA(cudaMemset(e->d_slots, 0, sizeof(SlotDev) * B));
  T5G_CUDA(cudaHostAlloc((void**)&e->h_mirror, sizeof(int) * 8 * B, cudaHostAllocMapped));
  memset(e->h_mirror, 0, sizeof(int) * 8 * B);
  T5G_CUDA(cudaHostGetDevicePointer((void**)&e->d_mirror, e->h_mirror, 0));
  T5G_CUDA(cudaHostAlloc((void**)&e->h_tokens, sizeof(int) * (size_t)B * cfg->max_dec_len, cudaHostAllocMapped));
  T5G_CUDA(cudaHostGetDevicePointer((void**)&e->d_tokens, e->h_tokens, 0));
  T5G_CUDA(cudaHostAlloc((void**)&e->h_picks, sizeof(int) * (size_t)B * cfg->max_dec_len, cudaHostAllocMapped));
  T5G_CUDA(cudaHostGetDevicePointer((void**)&e->d_picks, e->h_picks, 0));
  DM(e->d_forced, (size_t)B * cfg->max_dec_len);
  e->topk_pool_cap = B * cfg->max_dec_len;
  DM(e->d_topk_pool, e->topk_pool_cap);

  // ---- prefill workspaces ----
  const size_t T = cfg->max_prefill_tokens;
  DM(e->p_h, T * d); DM(e->p_y, T * d); DM(e->p_qkv, T * QKV); DM(e->p_memory, T * d); DM(e->p_ckv, T * 2 * KD); DM(e->p_final, T * d);
  DM(e->p_xn, T * d); DM(e->p_q, T * QD); DM(e->p_k, T * KD); DM(e->p_v, T * KD); DM(e->p_att, T * QD); DM(e->p_act, T * I);
  DM(e->p_mem_bf, T * d); DM(e->p_ck, T * KD); DM(e->p_cv, T * KD);
  if (const char* s = getenv("T5G_ATTN_TC")) e->use_tc_attn = atoi(s) != 0;
  if (const char* s = getenv("T5G_GEMV_PAIR")) e->use_pair = atoi(s) != 0;
  if (const char* s = getenv("T5G_ATTN_MMA")) e->attn_mma = atoi(s) != 0;
  if (const char* s = getenv("T5G_ATTN_TMA")) e->attn_tma = atoi(s) != 0;
  DM(e->p_logits, (size_t)e->logits_chunk * e->Vpad);
  DM(e->p_ids, T); DM(e->p_seg_of, T); DM(e->p_seg_off_e, B + 1); DM(e->p_seg_off_d, B + 1); DM(e->p_tok_slot, T); DM(e->p_tok_idx, T);
  DM(e->p_last_rows, B); DM(e->p_pos, T);
  e->h_stage_bytes = (T * 5 + 3 * (B + 1)) * 4 + 256;
  T5G_CUDA(cudaHostAlloc(&e->h_stage, e->h_stage_bytes, cudaHostAllocDefault));

  // ---- decode workspaces ----
  e->h_end = (cfg->n_dec_layers - 1) & 1;
  {
    int want = (2 * e->num_sms) / (e->Hkv * B);
    // 16-CTA clusters (non-portable size) while every cluster of the launch can be resident at once: bs <= 2 at 4 kv heads.
    // ctx 720, bs=1: self-attention 9.6 -> 7.8 us per layer; no change at ctx 150
    e->ns_self = (want >= 16 && e->Hkv * B * 16 <= e->num_sms) ? 16 : want >= 8 ? 8 : want >= 4 ? 4 : want >= 2 ? 2 : 1;
    while (D % e->ns_self) e->ns_self >>= 1;
    e->ns_cross = std::min(want >= 8 ? 4 : 2, e->ns_self);    // 64 text keys over 4 CTAs: 5.8 -> 5.3 us per layer at bs=1
    if (const char* s = getenv("T5G_NS_CROSS")) { const int v = atoi(s); if (v >= 1 && v <= 16 && !(v & (v - 1)) && D % v == 0) e->ns_cross = v; }
    if (const char* s = getenv("T5G_NS_SELF")) { const int v = atoi(s); if (v >= 1 && v <= 16 && !(v & (v - 1)) && D % v == 0) e->ns_self = v; }
  }
  DM(e->d_hA, (size_t)B * d); DM(e->d_hB, (size_t)B * d); DM(e->d_y, (size_t)B * d); DM(e->d_qkv, (size_t)B * QKV);
  DM(e->d_qc, (size_t)B * QD); DM(e->d_act, (size_t)B * I); DM(e->d_t1, (size_t)B * d); DM(e->d_logits, (size_t)B * e->Vpad);
  DM(e->d_rope, (size_t)B * D);
  T5G_CUDA(cudaMemset(e->d_y, 0, sizeof(float) * (size_t)B * d));
  T5G_CUDA(cudaMemset(e->d_qkv, 0, sizeof(float) * (size_t)B * QKV));   // batched step: split-K outputs start every layer at zero
  T5G_CUDA(cudaMemset(e->d_qc, 0, sizeof(float) * (size_t)B * QD));
  T5G_CUDA(cudaMemset(e->d_hA, 0, sizeof(float) * (size_t)B * d));
  T5G_CUDA(cudaMemset(e->d_hB, 0, sizeof(float) * (size_t)B * d));
  DM(e->d_sample_u, 4096); DM(e->d_sample_slots, 4096); DM(e->d_sample_silence, 256);
  e->samp_scratch_rows = std::max(B, 16);
  { const size_t V8 = ((size_t)e->V + 7) & ~(size_t)7; DM(e->d_samp_u64, (size_t)e->samp_scratch_rows * 2 * V8); DM(e->d_samp_f32, (size_t)e->samp_scratch_rows * 2 * V8); }
  DM(e->d_attn, (size_t)B * QD); DM(e->d_trace, 2 * T5G_TRACE_STRIDE);
  DM(e->d_barrier, 2); T5G_CUDA(cudaMemset(e->d_barrier, 0, 2 * sizeof(unsigned long long)));
  DM(e->d_order_self, B); DM(e->d_order_cross, B);
  { std::vector<int> id(B); for (int i = 0; i < B; ++i) id[i] = i;
    T5G_CUDA(cudaMemcpy(e->d_order_self, id.data(), sizeof(int) * B, cudaMemcpyHostToDevice));
    T5G_CUDA(cudaMemcpy(e->d_order_cross, id.data(), sizeof(int) * B, cudaMemcpyHostToDevice)); }
  DM(e->d_xn, (size_t)B * d); DM(e->d_attn_bf, (size_t)B * QD); DM(e->d_act_bf, (size_t)B * I); DM(e->d_t1_bf, (size_t)B * d);
  if (B > 4) {
    // chunked attention of the batched step: <= 16 chunks of attn_chunk keys per (row, kv head)
    if (const char* s = getenv("T5G_ATTN_CHUNK")) e->attn_chunk = atoi(s);
    if (e->attn_chunk > 0) {
      e->attn_chunk = std::max(32, (e->attn_chunk + 31) / 32 * 32);
      const int longest = std::max(cfg->max_dec_len, cfg->max_text_len);
      while (cdiv(longest, e->attn_chunk) > 16) e->attn_chunk += 32;
      DM(e->d_part_o, (size_t)B * e->Hkv * 16 * (e->Hq / e->Hkv) * D); DM(e->d_part_ml, (size_t)B * e->Hkv * 16 * (e->Hq / e->Hkv) * 2);
      DM(e->d_part_cnt, (size_t)B * e->Hkv);
      T5G_CUDA(cudaMemset(e->d_part_cnt, 0, sizeof(int) * (size_t)B * e->Hkv));
    }
  }
  T5G_CUDA(cudaDeviceSynchronize());
  return T5G_OK;
}

extern "C" int t5g_destroy(T5GEngine* e) {
  if (!e) return T5G_OK;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (auto& kv : e->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  for (void* p : e->allocs) cudaFree(p);
  if (e->h_mirror) cudaFreeHost(e->h_mirror);
  if (e->h_tokens) cudaFreeHost(e->h_tokens);
  if (e->h_picks) cudaFreeHost(e->h_picks);
  if (e->h_stage) cudaFreeHost(e->h_stage);
  if (e->load_stream) cudaStreamDestroy(e->load_stream);
  for (auto& ev : e->ev) if (ev) cudaEventDestroy(ev);
  delete e;
  return T5G_OK;
}

extern "C" int t5g_load_tensor(T5GEngine* e, const char* name, const void* data, int dtype, int ndim,
                               const int64_t* shape, int on_device) {
  T5G_CHECK(e && name && data && shape, T5G_ERR_INVALID, "null argument");
  T5G_CHECK(dtype >= 0 && dtype <= 2, T5G_ERR_INVALID, "bad dtype %d", dtype);
  T5G_CUDA(cudaSetDevice(e->device));
  const std::string n = name;
  // aliases / pruned text modules / buffers of the reference module are accepted and ignored
  if (n.rfind("encoder_module.", 0) == 0 || n.rfind("decoder_module.", 0) == 0 || n.rfind("backbone.lm_head", 0) == 0 ||
      n.rfind("backbone.model.decoder.embed_tokens", 0) == 0 || n == "class_weight" || n.find("inv_freq") != std::string::npos ||
      n.rfind("accuracy_metrics.", 0) == 0)     // legacy keys the reference drops as well (models/t5gemma.py:1131-1141)
    return T5G_OK;
  Dest dst;
  T5G_CHECK(resolve(e, n, &dst), T5G_ERR_INVALID, "unknown tensor name '%s'", name);
  int64_t rows = 1, cols = 1;
  if (ndim == 1) { rows = 1; cols = shape[0]; }
  else if (ndim == 2) { rows = shape[0]; cols = shape[1]; }
  else T5G_CHECK(false, T5G_ERR_INVALID, "tensor '%s': ndim %d unsupported", name, ndim);
  T5G_CHECK(rows == dst.rows && cols == dst.cols, T5G_ERR_INVALID, "tensor '%s': shape [%lld,%lld] != expected [%lld,%lld]",
            name, (long long)rows, (long long)cols, (long long)dst.rows, (long long)dst.cols);
  const size_t esz = dtype == T5G_F32 ? 4 : 2;
  const void* src = data;
  void* tmp = nullptr;
  if (!on_device) {
    T5G_CUDA(cudaMalloc(&tmp, (size_t)rows * cols * esz));
    cudaError_t er = cudaMemcpyAsync(tmp, data, (size_t)rows * cols * esz, cudaMemcpyHostToDevice, e->load_stream);
    if (er != cudaSuccess) { cudaFree(tmp); T5G_CUDA(er); }
    src = tmp;
  } else {
    T5G_CUDA(cudaDeviceSynchronize());   // the caller's producer stream is unknown
  }
  cudaError_t er = launch_pack(src, dtype, dst.ptr, dst.is_bf16 ? 1 : 0, rows, cols, dst.row_off, dst.row_mul, dst.add_one ? 1 : 0, e->load_stream);
  e->launches++;
  if (er == cudaSuccess) er = cudaStreamSynchronize(e->load_stream);
  if (tmp) cudaFree(tmp);
  T5G_CUDA(er);
  e->loaded.insert(n);
  return T5G_OK;
}

extern "C" int t5g_finalize_weights(T5GEngine* e) {
  T5G_CHECK(e, T5G_ERR_INVALID, "null engine");
  for (const auto& r : e->required)
    T5G_CHECK(e->loaded.count(r), T5G_ERR_STATE, "missing tensor '%s' (%zu of %zu loaded)", r.c_str(), e->loaded.size(), e->required.size());
  e->finalized = true;
  return T5G_OK;
}

// ------------------------------------------------------------------------------------------------
// Prefill
// ------------------------------------------------------------------------------------------------
namespace {

int alloc_pages(T5GEngine* e, std::vector<int>& dst, int n) {
  T5G_CHECK((int)e->free_pages.size() >= n, T5G_ERR_OOM, "KV pool exhausted: need %d pages, %zu free", n, e->free_pages.size());
  for (int i = 0; i < n; ++i) { dst.push_back(e->free_pages.back()); e->free_pages.pop_back(); }
  return T5G_OK;
}
void free_slot_pages(T5GEngine* e, HostSlot& hs) {
  for (int p : hs.self_pages) e->free_pages.push_back(p);
  for (int p : hs.cross_pages) e->free_pages.push_back(p);
  hs.self_pages.clear(); hs.cross_pages.clear();
}

// one transformer stack pass over M packed tokens.  is_dec selects the decoder (self causal + cross).
struct StackIO {
  int M; const int* seg_off; const int* seg_of; const float* pos;   // this stack's tokens
  int n_seg;
};

}  // namespace

static cudaError_t prefill_attention(T5GEngine* e, const AttnPrefillArgs& a, int Tk, int n_seg, int max_lq, cudaStream_t st) {
  e->launches++;
  if (e->use_tc_attn && attn_prefill_tc_supported(a.D)) return launch_attn_prefill_tc(a, Tk, n_seg, max_lq, st, e->use_pdl);
  return launch_attn_prefill(a, st);
}

extern "C" int t5g_prefill(T5GEngine* e, const T5GRequest* reqs, int n_req, void* stream_) {
  T5G_CHECK(e && reqs && n_req > 0, T5G_ERR_INVALID, "bad arguments");
  T5G_CHECK(e->finalized, T5G_ERR_STATE, "weights not finalized");
  T5G_CHECK(n_req <= e->c.max_slots, T5G_ERR_INVALID, "n_req %d > max_slots %d", n_req, e->c.max_slots);
  T5G_CUDA(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream_;
  const int d = e->d, I = e->I, QD = e->QD, KD = e->KD, QKV = e->QKV, D = e->D, PT = e->PT;
  const T5GConfig& c = e->c;
  // ---- validate + host bookkeeping ----
  int Te = 0, Td = 0, max_text = 0, max_dec = 0;
  std::set<int> seen;
  for (int r = 0; r < n_req; ++r) {
    const T5GRequest& q = reqs[r];
    T5G_CHECK(q.slot >= 0 && q.slot < c.max_slots && !seen.count(q.slot), T5G_ERR_INVALID, "request %d: bad/duplicate slot %d", r, q.slot);
    seen.insert(q.slot);
    T5G_CHECK(q.n_text >= 1 && q.n_text <= c.max_text_len && q.text_ids, T5G_ERR_INVALID, "request %d: n_text %d out of range (max %d)", r, q.n_text, c.max_text_len);
    T5G_CHECK(q.n_dec >= 1 && q.dec_ids, T5G_ERR_INVALID, "request %d: n_dec must be >= 1 (BOS)", r);
    for (int i = 0; i < q.n_text; ++i) T5G_CHECK(q.text_ids[i] >= 0 && q.text_ids[i] < c.text_vocab, T5G_ERR_INVALID, "request %d: text id %lld out of range", r, (long long)q.text_ids[i]);
    for (int i = 0; i < q.n_dec; ++i) T5G_CHECK(q.dec_ids[i] >= 0 && q.dec_ids[i] < e->V, T5G_ERR_INVALID, "request %d: audio id %lld out of range", r, (long long)q.dec_ids[i]);
    const T5GSampling& s = q.sampling;
    const bool minp = s.min_p > 0.f && s.min_p < 1.f;
    T5G_CHECK(s.temperature > 0.f, T5G_ERR_INVALID, "request %d: temperature must be > 0", r);
    (void)minp;
    Te += q.n_text; Td += q.n_dec;
    max_text = std::max(max_text, q.n_text); max_dec = std::max(max_dec, q.n_dec);
  }
  T5G_CHECK(Te <= c.max_prefill_tokens && Td <= c.max_prefill_tokens, T5G_ERR_INVALID, "prefill tokens (%d text, %d audio) exceed max_prefill_tokens %d", Te, Td, c.max_prefill_tokens);

  // staging layout (pinned): ids_e[Te] ids_d[Td] seg_of_e[Te] seg_of_d[Td] pos_e[Te] pos_d[Td] slot_e idx_e slot_d idx_d offs
  const int Tm = c.max_prefill_tokens;
  int* hs = (int*)e->h_stage;
  int* h_ids = hs; int* h_seg_of = hs + Tm; float* h_pos = (float*)(hs + 2 * Tm); int* h_tslot = hs + 3 * Tm; int* h_tidx = hs + 4 * Tm;
  int* h_off_e = hs + 5 * Tm; int* h_off_d = h_off_e + (c.max_slots + 1); int* h_last = h_off_d + (c.max_slots + 1);

  e->prefill_counter++;
  std::vector<int> est_totals(n_req);
  for (int r = 0; r < n_req; ++r) {
    const T5GRequest& q = reqs[r];
    HostSlot& hsl = e->hslots[q.slot];
    free_slot_pages(e, hsl);
    const int prompt_offset = q.prompt_frames + 1;
    // models/t5gemma.py:925-933: est_total = target_total + 1 (BOS); without a target (tgt_y_lens=None, target_total < 0
    // here) the reference falls back to current_length + encodec_sr * progress_lookahead_secs (2.0) and has no time budget
    const bool has_target = q.target_total >= 0;
    const int est_total = std::max(has_target ? q.target_total + 1 : (int)(q.n_dec + (int)(c.encodec_sr * 2.0)), q.n_dec);
    est_totals[r] = est_total;
    int budget_limit, max_new, slot_max_new = q.max_new_tokens;
    if (has_target) {
      const double lim = (double)q.target_total - (double)prompt_offset + (double)c.encodec_sr * (double)c.extra_cutoff;
      budget_limit = (int)std::floor(lim);
      max_new = budget_limit + 2;
      if (max_new < 1) max_new = 1;
      if (q.max_new_tokens > 0) max_new = std::min(max_new, q.max_new_tokens);
    } else {
      // unbounded in the reference; here the slot's KV capacity is the bound (eos is forced at max_dec_len)
      budget_limit = 0x7fffffff;
      max_new = c.max_dec_len - q.n_dec;
      T5G_CHECK(max_new >= 1, T5G_ERR_INVALID, "request %d: no room to generate (n_dec %d, max_dec_len %d)", r, q.n_dec, c.max_dec_len);
      if (q.max_new_tokens > 0) max_new = std::min(max_new, q.max_new_tokens);
      slot_max_new = max_new;
    }
    const int max_len = q.n_dec + max_new;
    T5G_CHECK(max_len <= c.max_dec_len, T5G_ERR_INVALID, "request %d: needs %d decoder tokens > max_dec_len %d", r, max_len, c.max_dec_len);
    int rc = alloc_pages(e, hsl.self_pages, cdiv(max_len, PT)); if (rc) return rc;
    rc = alloc_pages(e, hsl.cross_pages, cdiv(q.n_text, PT)); if (rc) return rc;
    hsl.in_use = true; hsl.n_text = q.n_text; hsl.n_dec = q.n_dec; hsl.prefill_id = e->prefill_counter;
    SlotDev& sd = e->h_slots[q.slot];
    memset(&sd, 0, sizeof(sd));
    sd.active = 1; sd.finished = 0; sd.n_generated = 0; sd.cur_len = q.n_dec; sd.prompt_offset = prompt_offset;
    sd.target_total = q.target_total; sd.est_total = est_total; sd.n_text = q.n_text; sd.budget_limit = budget_limit;
    sd.max_new_tokens = slot_max_new; sd.top_k = q.sampling.top_k; sd.top_p = q.sampling.top_p; sd.min_p = q.sampling.min_p;
    sd.temperature = q.sampling.temperature; sd.uniforms = q.uniforms; sd.n_uniforms = q.n_uniforms;
    sd.topk_sched_off = -1; sd.n_topk_sched = 0; sd.last_token = 0; sd.pos = 0.f;
    sd.prev_token = -1; sd.consec_silence = 0; sd.stop_repetition = q.stop_repetition; sd.silence_off = -1; sd.n_silence = 0;
    // every slot owns a fixed region of the int pool ([slot*max_dec_len, (slot+1)*max_dec_len)): a prefill that admits new
    // requests while other slots are still decoding never touches their schedules / silence lists
    int pool_off = q.slot * c.max_dec_len;
    const int pool_end = pool_off + c.max_dec_len;
    if (q.silence_tokens && q.n_silence > 0) {
      T5G_CHECK(pool_off + q.n_silence <= pool_end, T5G_ERR_OOM, "request %d: %d silence tokens exceed the slot's int pool (%d)", r, q.n_silence, c.max_dec_len);
      CU(cudaMemcpyAsync(e->d_topk_pool + pool_off, q.silence_tokens, sizeof(int) * q.n_silence, cudaMemcpyHostToDevice, st));
      sd.silence_off = pool_off; sd.n_silence = q.n_silence;
      pool_off += q.n_silence;
    }
    if (q.top_k_schedule && q.n_top_k_schedule > 0) {
      // entries past the last step that can run are never read (models/t5gemma.py:991-994 clamps the index the other way)
      const int n_sched = std::min(q.n_top_k_schedule, std::max(1, max_new));
      T5G_CHECK(pool_off + n_sched <= pool_end, T5G_ERR_OOM, "request %d: top_k schedule + silence list exceed the slot's int pool (%d)", r, c.max_dec_len);
      CU(cudaMemcpyAsync(e->d_topk_pool + pool_off, q.top_k_schedule, sizeof(int) * n_sched, cudaMemcpyHostToDevice, st));
      sd.topk_sched_off = pool_off; sd.n_topk_sched = n_sched;
    }
    if (q.forced_tokens && q.n_forced > 0) {
      T5G_CHECK(q.n_forced <= c.max_dec_len, T5G_ERR_INVALID, "request %d: n_forced %d > max_dec_len", r, q.n_forced);
      for (int i = 0; i < q.n_forced; ++i) T5G_CHECK(q.forced_tokens[i] >= 0 && q.forced_tokens[i] < e->V, T5G_ERR_INVALID, "request %d: forced token out of range", r);
      CU(cudaMemcpyAsync(e->d_forced + (size_t)q.slot * c.max_dec_len, q.forced_tokens, sizeof(int) * q.n_forced, cudaMemcpyHostToDevice, st));
      sd.n_forced = q.n_forced;
    }
    CU(cudaMemcpyAsync(e->d_slots + q.slot, &sd, sizeof(SlotDev), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->d_self_bt + (size_t)q.slot * e->max_self_pages, hsl.self_pages.data(), sizeof(int) * hsl.self_pages.size(), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->d_cross_bt + (size_t)q.slot * e->max_cross_pages, hsl.cross_pages.data(), sizeof(int) * hsl.cross_pages.size(), cudaMemcpyHostToDevice, st));
    e->h_mirror[q.slot * 8 + 0] = 1; e->h_mirror[q.slot * 8 + 1] = 0; e->h_mirror[q.slot * 8 + 2] = 0; e->h_mirror[q.slot * 8 + 3] = q.n_dec; e->h_mirror[q.slot * 8 + 4] = 0;
  }
  // the pageable host vectors above must outlive the async copies
  CU(cudaStreamSynchronize(st));

  auto stage_and_upload = [&](bool dec) -> int {
    int off = 0;
    int* h_off = dec ? h_off_d : h_off_e;
    for (int r = 0; r < n_req; ++r) {
      const T5GRequest& q = reqs[r];
      const int n = dec ? q.n_dec : q.n_text;
      h_off[r] = off;
      if (dec) e->hslots[q.slot].dec_off = off; else e->hslots[q.slot].mem_off = off;
      for (int i = 0; i < n; ++i) {
        h_ids[off + i] = (int)(dec ? q.dec_ids[i] : q.text_ids[i]);
        h_seg_of[off + i] = r; h_tslot[off + i] = q.slot; h_tidx[off + i] = i;
        float p;
        if (dec) {            // models/t5gemma.py:945-947: arange/max(1,est_total-1)*scale in fp32
          p = (float)i / (float)std::max(1, est_totals[r] - 1);
          p = p * c.progress_scale;
        } else {              // models/t5gemma.py:609-624
          p = (float)i / ((float)std::max(n, 2) - 1.0f);
          p = p * c.progress_scale;
        }
        h_pos[off + i] = p;
      }
      off += n;
      if (dec) h_last[r] = off - 1;
    }
    h_off[n_req] = off;
    CU(cudaMemcpyAsync(e->p_ids, h_ids, sizeof(int) * off, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->p_seg_of, h_seg_of, sizeof(int) * off, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->p_pos, h_pos, sizeof(float) * off, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->p_tok_slot, h_tslot, sizeof(int) * off, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->p_tok_idx, h_tidx, sizeof(int) * off, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(dec ? e->p_seg_off_d : e->p_seg_off_e, h_off, sizeof(int) * (n_req + 1), cudaMemcpyHostToDevice, st));
    return T5G_OK;
  };

  // ================= encoder (HF:664-718, 426-452) =================
  CU(cudaEventRecord(e->ev[0], st));
  { int rc = stage_and_upload(false); if (rc) return rc; }
  CU(launch_embed(e->enc_embed, e->p_ids, sqrtf((float)d), e->p_h, Te, d, st)); e->launches++;
  const int ze_qkv = wants_zero(e, Te, QKV, d), ze_o = wants_zero(e, Te, d, QD), ze_down = wants_zero(e, Te, d, I);
  for (int l = 0; l < c.n_enc_layers; ++l) {
    const EncLayer& L = e->enc[l];
    // h += post_ff(prev y) ; xn = pre_sa(h)
    // small prefills: the norm kernels zero the fp32 outputs of the split-K GEMMs that follow them (after reading them)
    if (l == 0) CU(launch_norm(e->p_h, nullptr, nullptr, L.g_pre_sa, nullptr, e->p_xn, nullptr, Te, d, c.rms_eps, st, e->use_pdl,
                               ze_qkv ? e->p_qkv : nullptr, QKV, ze_o ? e->p_y : nullptr, d));
    else CU(launch_norm(e->p_h, e->p_y, e->enc[l - 1].g_post_ff, L.g_pre_sa, e->p_h, e->p_xn, nullptr, Te, d, c.rms_eps, st, e->use_pdl,
                        ze_qkv ? e->p_qkv : nullptr, QKV, ze_o ? e->p_y : nullptr, d));
    e->launches++;
    CU(gemm(e, e->p_xn, L.wqkv, Te, QKV, d, GE_F32, nullptr, e->p_qkv, QKV, st, ze_qkv));
    RopeSplitArgs ra{}; ra.qkv = e->p_qkv; ra.ld = QKV; ra.q_off = 0; ra.k_off = QD; ra.v_off = QD + KD; ra.pos = e->p_pos; ra.M = Te;
    ra.Hq = e->Hq; ra.Hkv = e->Hkv; ra.D = D; ra.inv_freq = e->inv_freq; ra.q_out = e->p_q; ra.k_out = e->p_k; ra.v_out = e->p_v; ra.block_table = nullptr;
    CU(launch_rope_split(ra, st, e->use_pdl)); e->launches++;
    AttnPrefillArgs aa{}; aa.q = e->p_q; aa.k = e->p_k; aa.v = e->p_v; aa.q_seg_off = e->p_seg_off_e; aa.k_seg_off = e->p_seg_off_e; aa.q_seg_of = e->p_seg_of;
    aa.Tq = Te; aa.Hq = e->Hq; aa.Hkv = e->Hkv; aa.D = D; aa.causal = 0; aa.window = c.enc_layer_sliding[l] ? c.sliding_window : 0;
    aa.scale = c.attn_scale; aa.softcap = c.attn_softcap; aa.out = e->p_att;
    CU(prefill_attention(e, aa, Te, n_req, max_text, st));
    CU(gemm(e, e->p_att, L.wo, Te, d, QD, GE_F32, nullptr, e->p_y, d, st, ze_o));
    CU(launch_norm(e->p_h, e->p_y, L.g_post_sa, L.g_pre_ff, e->p_h, e->p_xn, nullptr, Te, d, c.rms_eps, st, e->use_pdl,
                   ze_down ? e->p_y : nullptr, d)); e->launches++;
    CU(gemm(e, e->p_xn, L.wgu, Te, 2 * I, d, GE_GEGLU_BF16, nullptr, e->p_act, I, st));
    CU(gemm(e, e->p_act, L.wd, Te, d, I, GE_F32, nullptr, e->p_y, d, st, ze_down));
  }
  // memory = final_norm(h + post_ff(y))
  CU(launch_norm(e->p_h, e->p_y, e->enc[c.n_enc_layers - 1].g_post_ff, e->g_enc_final, nullptr, e->p_mem_bf, e->p_memory, Te, d, c.rms_eps, st, e->use_pdl)); e->launches++;
  CU(cudaEventRecord(e->ev[1], st));

  // ================= decoder over BOS + prompt (HF:748-828; models/t5gemma.py:183-243) =================
  // encoder-side metadata needed for the cross K/V (positions, slots) is re-derived per layer from a
  // second staging copy kept on the device: pos_e/tok_* are overwritten by the decoder upload, so keep copies.
  float* pos_e; int *tslot_e, *tidx_e, *seg_of_e;
  // copies of encoder metadata (device->device) into the (unused during prefill) p_final scratch region
  // p_final is [T,d] floats; we need 4*Te ints, d >= 8 guarantees room.
  pos_e = e->p_final; tslot_e = (int*)(e->p_final + Tm); tidx_e = (int*)(e->p_final + 2 * (size_t)Tm); seg_of_e = (int*)(e->p_final + 3 * (size_t)Tm);
  CU(cudaMemcpyAsync(pos_e, e->p_pos, sizeof(float) * Te, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(tslot_e, e->p_tok_slot, sizeof(int) * Te, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(tidx_e, e->p_tok_idx, sizeof(int) * Te, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(seg_of_e, e->p_seg_of, sizeof(int) * Te, cudaMemcpyDeviceToDevice, st));
  CU(cudaStreamSynchronize(st));     // staging buffer is reused by the decoder upload
  { int rc = stage_and_upload(true); if (rc) return rc; }
  CU(cudaMemcpyAsync(e->p_last_rows, h_last, sizeof(int) * n_req, cudaMemcpyHostToDevice, st));
  CU(launch_embed(e->audio_emb, e->p_ids, sqrtf((float)d), e->p_h, Td, d, st)); e->launches++;
  const int zd_qkv = wants_zero(e, Td, QKV, d), zd_o = wants_zero(e, Td, d, QD), zd_qc = wants_zero(e, Td, QD, d),
            zd_down = wants_zero(e, Td, d, I);
  for (int l = 0; l < c.n_dec_layers; ++l) {
    const DecLayer& L = e->dec[l];
    if (l == 0) CU(launch_norm(e->p_h, nullptr, nullptr, L.g_pre_sa, nullptr, e->p_xn, nullptr, Td, d, c.rms_eps, st, e->use_pdl,
                               zd_qkv ? e->p_qkv : nullptr, QKV, zd_o ? e->p_y : nullptr, d));
    else CU(launch_norm(e->p_h, e->p_y, e->dec[l - 1].g_post_ff, L.g_pre_sa, e->p_h, e->p_xn, nullptr, Td, d, c.rms_eps, st, e->use_pdl,
                        zd_qkv ? e->p_qkv : nullptr, QKV, zd_o ? e->p_y : nullptr, d));
    e->launches++;
    CU(gemm(e, e->p_xn, L.wqkv, Td, QKV, d, GE_F32, nullptr, e->p_qkv, QKV, st, zd_qkv));
    RopeSplitArgs ra{}; ra.qkv = e->p_qkv; ra.ld = QKV; ra.q_off = 0; ra.k_off = QD; ra.v_off = QD + KD; ra.pos = e->p_pos; ra.M = Td;
    ra.Hq = e->Hq; ra.Hkv = e->Hkv; ra.D = D; ra.inv_freq = e->inv_freq; ra.q_out = e->p_q; ra.k_out = e->p_k; ra.v_out = e->p_v;
    ra.pool = e->pool; ra.layer = l; ra.block_table = e->d_self_bt; ra.bt_stride = e->max_self_pages; ra.tok_slot = e->p_tok_slot; ra.tok_idx = e->p_tok_idx;
    CU(launch_rope_split(ra, st, e->use_pdl)); e->launches++;
    AttnPrefillArgs aa{}; aa.q = e->p_q; aa.k = e->p_k; aa.v = e->p_v; aa.q_seg_off = e->p_seg_off_d; aa.k_seg_off = e->p_seg_off_d; aa.q_seg_of = e->p_seg_of;
    aa.Tq = Td; aa.Hq = e->Hq; aa.Hkv = e->Hkv; aa.D = D; aa.causal = 1; aa.window = c.dec_layer_sliding[l] ? c.sliding_window : 0;
    aa.scale = c.attn_scale; aa.softcap = c.attn_softcap; aa.out = e->p_att;
    CU(prefill_attention(e, aa, Td, n_req, max_dec, st));
    CU(gemm(e, e->p_att, L.wo, Td, d, QD, GE_F32, nullptr, e->p_y, d, st, zd_o));
    // p_qkv is free again (RoPE split consumed it): rows of width QD for the cross q projection; p_y for o_cross
    CU(launch_norm(e->p_h, e->p_y, L.g_post_sa, L.g_pre_ca, e->p_h, e->p_xn, nullptr, Td, d, c.rms_eps, st, e->use_pdl,
                   zd_qc ? e->p_qkv : nullptr, QD, zd_o ? e->p_y : nullptr, d)); e->launches++;
    // cross attention: q = RoPE(q_proj(x), decoder pos); K/V of this layer computed once from memory
    CU(gemm(e, e->p_xn, L.wq_c, Td, QD, d, GE_F32, nullptr, e->p_qkv, QD, st, zd_qc));
    RopeSplitArgs rq{}; rq.qkv = e->p_qkv; rq.ld = QD; rq.q_off = 0; rq.k_off = -1; rq.v_off = -1; rq.pos = e->p_pos; rq.M = Td;
    rq.Hq = e->Hq; rq.Hkv = e->Hkv; rq.D = D; rq.inv_freq = e->inv_freq; rq.q_out = e->p_q; rq.block_table = nullptr;
    CU(launch_rope_split(rq, st, e->use_pdl)); e->launches++;
    CU(gemm(e, e->p_mem_bf, L.wkv_c, Te, 2 * KD, d, GE_F32, nullptr, e->p_ckv, 2 * KD, st));
    RopeSplitArgs rk{}; rk.qkv = e->p_ckv; rk.ld = 2 * KD; rk.q_off = -1; rk.k_off = 0; rk.v_off = KD; rk.pos = pos_e; rk.M = Te;
    rk.Hq = e->Hq; rk.Hkv = e->Hkv; rk.D = D; rk.inv_freq = e->inv_freq; rk.k_out = e->p_ck; rk.v_out = e->p_cv;
    rk.pool = e->pool; rk.layer = l; rk.block_table = e->d_cross_bt; rk.bt_stride = e->max_cross_pages; rk.tok_slot = tslot_e; rk.tok_idx = tidx_e;
    CU(launch_rope_split(rk, st, e->use_pdl)); e->launches++;
    AttnPrefillArgs ac{}; ac.q = e->p_q; ac.k = e->p_ck; ac.v = e->p_cv; ac.q_seg_off = e->p_seg_off_d; ac.k_seg_off = e->p_seg_off_e; ac.q_seg_of = e->p_seg_of;
    ac.Tq = Td; ac.Hq = e->Hq; ac.Hkv = e->Hkv; ac.D = D; ac.causal = 0; ac.window = 0; ac.scale = c.attn_scale; ac.softcap = c.attn_softcap; ac.out = e->p_att;
    CU(prefill_attention(e, ac, Te, n_req, max_dec, st));
    CU(gemm(e, e->p_att, L.wo_c, Td, d, QD, GE_F32, nullptr, e->p_y, d, st, zd_o));
    CU(launch_norm(e->p_h, e->p_y, L.g_post_ca, L.g_pre_ff, e->p_h, e->p_xn, nullptr, Td, d, c.rms_eps, st, e->use_pdl,
                   zd_down ? e->p_y : nullptr, d)); e->launches++;
    CU(gemm(e, e->p_xn, L.wgu, Td, 2 * I, d, GE_GEGLU_BF16, nullptr, e->p_act, I, st));
    CU(gemm(e, e->p_act, L.wd, Td, d, I, GE_F32, nullptr, e->p_y, d, st, zd_down));
  }
  // h = h + post_ff(y) (kept, pre-final-norm) ; p_qkv <- final_norm(h) fp32 for teacher-forced logits
  CU(launch_norm(e->p_h, e->p_y, e->dec[c.n_dec_layers - 1].g_post_ff, e->g_dec_final, e->p_h, nullptr, e->p_final, Td, d, c.rms_eps, st, e->use_pdl)); e->launches++;
  // hand the last token of every request to the decode buffers: h_end buffer <- h, y <- 0
  {
    for (int r = 0; r < n_req; ++r) {
      float* hdst = ((e->c.max_slots > 4 || e->h_end == 0) ? e->d_hA : e->d_hB) + (size_t)reqs[r].slot * d;
      CU(cudaMemcpyAsync(hdst, e->p_h + (size_t)h_last[r] * d, sizeof(float) * d, cudaMemcpyDeviceToDevice, st));
      CU(cudaMemsetAsync(e->d_y + (size_t)reqs[r].slot * d, 0, sizeof(float) * d, st));
    }
  }
  CU(cudaEventRecord(e->ev[2], st));
  CU(cudaStreamSynchronize(st));
  cudaEventElapsedTime(&e->timings[0], e->ev[0], e->ev[1]);
  cudaEventElapsedTime(&e->timings[2], e->ev[1], e->ev[2]);
  e->timings[1] = 0.f;
  e->tot_prefill_ms += (double)e->timings[0] + (double)e->timings[2]; e->tot_prefill_calls++;
  return T5G_OK;
}

// ------------------------------------------------------------------------------------------------
// Decode step
// ------------------------------------------------------------------------------------------------
namespace {

int enqueue_step(T5GEngine* e, cudaStream_t st, int* n_launch) {
  const T5GConfig& c = e->c;
  const int d = e->d, I = e->I, QD = e->QD, QKV = e->QKV, D = e->D;
  const int B = c.max_slots;
  const bool pdl = e->use_pdl;
  int nl = 0;
  float* hbuf[2] = {e->d_hA, e->d_hB};
  int kidx = 0;                                   // kernel sequence number inside the step (tracing)
  auto next_trace = [&]() -> unsigned long long* {
    if (!e->use_trace || kidx >= T5G_TRACE_STRIDE) return nullptr;
    return e->d_trace + (kidx++);
  };
  // batch rows are processed in groups of <= 4 by the GEMV family
  auto gemv_all = [&](GemvArgs a, int P, int E, size_t in_stride, size_t out_stride, bool overlap = true) -> cudaError_t {
    for (int b0 = 0; b0 < B; b0 += 4) {
      GemvArgs g = a;
      g.B = std::min(4, B - b0); g.slot0 = b0;
      if (g.x) g.x += (size_t)b0 * in_stride;
      if (g.h_in) g.h_in += (size_t)b0 * d;
      if (g.y) g.y += (size_t)b0 * d;
      if (g.h_out) g.h_out += (size_t)b0 * d;
      g.out += (size_t)b0 * out_stride;
      g.trace = next_trace();
      cudaError_t er = launch_gemv(g, P, E, e->num_sms, st, pdl && overlap);
      if (er != cudaSuccess) return er;
      nl++;
    }
    return cudaSuccess;
  };
  GemvArgs z{}; z.eps = c.rms_eps; z.slots = e->d_slots;
  if (e->use_trace) {
    CU(cudaMemsetAsync(e->d_trace, 0xFF, sizeof(unsigned long long) * T5G_TRACE_STRIDE, st));
    CU(cudaMemsetAsync(e->d_trace + T5G_TRACE_STRIDE, 0, sizeof(unsigned long long) * T5G_TRACE_STRIDE, st));
  }

  const DecLayer& Llast = e->dec[c.n_dec_layers - 1];
  // ---- head: t1 = gelu(W1 * final_norm(h + post_ff(y)) + b1) ; logits = W2 t1 + b2 ----
  { GemvArgs a = z; a.W = e->head_w1; a.N = d; a.K = d; a.h_in = hbuf[e->h_end]; a.y = e->d_y; a.g_post = Llast.g_post_ff; a.g_pre = e->g_dec_final;
    a.h_out = nullptr; a.bias = e->head_b1; a.out = e->d_t1; a.out_stride = d;
    CU(gemv_all(a, P_RES_NORM, E_BIAS_GELU, 0, d)); }
  { GemvArgs a = z; a.W = e->head_w2; a.N = e->Vpad; a.K = d; a.x = e->d_t1; a.bias = e->head_b2; a.out = e->d_logits; a.out_stride = e->Vpad;
    CU(gemv_all(a, P_PLAIN, E_BIAS, d, e->Vpad)); }
  // ---- sample + stop rules + state update ----
  { SamplerArgs s{}; s.logits = e->d_logits; s.ld = e->Vpad; s.V = e->V; s.slots = e->d_slots; s.topk_sched_pool = e->d_topk_pool;
    s.eos = c.eos_token; s.encodec_sr = c.encodec_sr; s.text_guard = c.text_guard_frames_per_token; s.progress_scale = c.progress_scale;
    s.tokens_out = e->d_tokens; s.tokens_stride = c.max_dec_len; s.argmax_out = nullptr; s.rows = B; s.host_mirror = e->d_mirror;
    s.picks_out = e->d_picks; s.forced_pool = e->d_forced; s.rope_out = e->d_rope; s.inv_freq = e->inv_freq; s.head_dim = D;
    s.scratch_u64 = e->d_samp_u64; s.scratch_f32 = e->d_samp_f32;
    s.trace = next_trace(); s.probe = e->use_trace ? e->d_trace + 1000 : nullptr;
    CU(launch_sampler(s, st, pdl)); nl++; }
  // ---- 26 decoder layers at q_len = 1 ----
  int t = 0;   // hbuf[t] holds the current residual
  for (int l = 0; l < c.n_dec_layers; ++l) {
    const DecLayer& L = e->dec[l];
    { GemvArgs a = z; a.W = L.wqkv; a.N = QKV; a.K = d; a.g_pre = L.g_pre_sa; a.out = e->d_qkv; a.out_stride = QKV;
      // the first kernel after the sampler starts only when the sampler has COMPLETED (no programmatic overlap): every later
      // kernel may then read the slot state / RoPE table the sampler wrote before its own griddepcontrol.wait
      if (l == 0) { a.emb = e->audio_emb; a.emb_scale = sqrtf((float)d); a.h_out = hbuf[0]; t = 0; CU(gemv_all(a, P_EMBED_NORM, E_STORE, 0, QKV, false)); }
      else { a.h_in = hbuf[t]; a.y = e->d_y; a.g_post = e->dec[l - 1].g_post_ff; a.h_out = hbuf[t ^ 1]; t ^= 1; CU(gemv_all(a, P_RES_NORM, E_STORE, 0, QKV)); } }
    { AttnDecodeArgs a{}; a.pool = e->pool; a.layer = l; a.block_table = e->d_self_bt; a.bt_stride = e->max_self_pages; a.q = e->d_qkv; a.q_stride = QKV;
      a.kv_new = e->d_qkv + QD; a.kv_stride = QKV; a.slots = e->d_slots; a.B = B; a.Hq = e->Hq; a.Hkv = e->Hkv; a.D = D; a.n_splits = e->ns_self;
      a.is_cross = 0; a.window = c.dec_layer_sliding[l] ? c.sliding_window : 0; a.scale = c.attn_scale; a.softcap = c.attn_softcap; a.inv_freq = e->inv_freq; a.rope_cs = e->d_rope;
      a.out = e->d_attn; a.preload = 1; a.trace = next_trace();
      CU(launch_attn_decode(a, st, pdl));
      nl++; }
    GemvPairArgs gp{};
    gp.W1 = L.wo; gp.N1 = d; gp.K1 = QD; gp.x = e->d_attn; gp.y = e->d_y; gp.W2 = L.wq_c; gp.N2 = QD; gp.K2 = d;
    gp.h_in = hbuf[t]; gp.g_post = L.g_post_sa; gp.g_pre = L.g_pre_ca; gp.h_out = hbuf[t ^ 1]; gp.eps = c.rms_eps;
    gp.out = e->d_qc; gp.out_stride = QD; gp.B = B; gp.slots = e->d_slots; gp.err_slots = e->d_slots; gp.barrier = e->d_barrier;
    if (e->use_pair && B <= 4 && gemv_pair_supported(gp)) {
      gp.trace = next_trace();
      CU(launch_gemv_pair(gp, e->num_sms, st, pdl)); nl++; t ^= 1;
    } else {
      { GemvArgs a = z; a.W = L.wo; a.N = d; a.K = QD; a.x = e->d_attn; a.out = e->d_y; a.out_stride = d;
        CU(gemv_all(a, P_PLAIN, E_STORE, QD, d)); }
      { GemvArgs a = z; a.W = L.wq_c; a.N = QD; a.K = d; a.h_in = hbuf[t]; a.y = e->d_y; a.g_post = L.g_post_sa; a.g_pre = L.g_pre_ca; a.h_out = hbuf[t ^ 1]; t ^= 1;
        a.out = e->d_qc; a.out_stride = QD;
        CU(gemv_all(a, P_RES_NORM, E_STORE, 0, QD)); }
    }
    {
      { AttnDecodeArgs a{}; a.pool = e->pool; a.layer = l; a.block_table = e->d_cross_bt; a.bt_stride = e->max_cross_pages; a.q = e->d_qc; a.q_stride = QD;
        a.kv_new = nullptr; a.kv_stride = 0; a.slots = e->d_slots; a.B = B; a.Hq = e->Hq; a.Hkv = e->Hkv; a.D = D; a.n_splits = e->ns_cross;
        a.is_cross = 1; a.window = 0; a.scale = c.attn_scale; a.softcap = c.attn_softcap; a.inv_freq = e->inv_freq; a.rope_cs = e->d_rope;
        a.out = e->d_attn; a.preload = 1; a.trace = next_trace();
        CU(launch_attn_decode(a, st, pdl));
      nl++; }
      { GemvArgs a = z; a.W = L.wo_c; a.N = d; a.K = QD; a.x = e->d_attn; a.out = e->d_y; a.out_stride = d;
        CU(gemv_all(a, P_PLAIN, E_STORE, QD, d)); }
    }
    { GemvArgs a = z; a.W = L.wgu; a.N = 2 * I; a.K = d; a.h_in = hbuf[t]; a.y = e->d_y; a.g_post = L.g_post_ca; a.g_pre = L.g_pre_ff; a.h_out = hbuf[t ^ 1]; t ^= 1;
      a.out = e->d_act; a.out_stride = I;
      CU(gemv_all(a, P_RES_NORM, E_GEGLU, 0, I)); }
    { GemvArgs a = z; a.W = L.wd; a.N = d; a.K = I; a.x = e->d_act; a.out = e->d_y; a.out_stride = d;
      CU(gemv_all(a, P_PLAIN, E_STORE, I, d)); }
  }
  if (t != e->h_end) { t5g_set_error("internal: residual buffer parity %d != %d", t, e->h_end); return T5G_ERR_STATE; }
  *n_launch = nl;
  return T5G_OK;
}


// Batched decode step (B > 4 rows): the projections become skinny tensor-core GEMMs with M = B (weights are
// streamed once for the whole batch); attention and sampling are the same kernels as the bs=1 path.
int enqueue_step_batched(T5GEngine* e, cudaStream_t st, int* n_launch) {
  const T5GConfig& c = e->c;
  const int d = e->d, I = e->I, QD = e->QD, QKV = e->QKV, D = e->D;
  const int B = c.max_slots;
  int nl = 0;
  const bool pdl = e->use_pdl;
  const bool pdl_norm = pdl, pdl_gemm = pdl, pdl_attn = pdl, pdl_samp = pdl;
  int kidx = 0;
  auto next_trace = [&]() -> unsigned long long* {
    if (!e->use_trace || kidx >= T5G_TRACE_STRIDE) return nullptr;
    return e->d_trace + (kidx++);
  };
  if (e->use_trace) {
    CU(cudaMemsetAsync(e->d_trace, 0xFF, sizeof(unsigned long long) * T5G_TRACE_STRIDE, st));
    CU(cudaMemsetAsync(e->d_trace + T5G_TRACE_STRIDE, 0, sizeof(unsigned long long) * T5G_TRACE_STRIDE, st));
  }
  // fp32 outputs of the split-K GEMMs are zeroed ahead of time by the idle epilogue warps of an EARLIER GEMM of the step
  // (za / zb): d_y by the qkv, q_c and gate|up GEMMs (each after the norm kernel that read it), d_qc by the qkv GEMM, d_qkv
  // of the next layer by the gate|up GEMM.  (The norm kernels did this before: 1 us of their 3.4 us at 64 rows.)
  auto G = [&](const bf16* A, const bf16* W, int N, int K, int epi, const float* bias, void* out, int ldo,
               float* za = nullptr, size_t na = 0, float* zb = nullptr, size_t nb = 0) -> cudaError_t {
    GemmArgs g{A, W, B, N, K, epi, bias, out, ldo, 1};
    g.zero_a = za; g.zero_na = na; g.zero_b = zb; g.zero_nb = nb;
    g.trace = next_trace();
    if (g.trace && kidx - 1 == 4 + 5 * 11 + 9) g.probe = e->d_trace + 1016;   // layer 5 gate|up: in-kernel checkpoints
    nl += 1;
    if (e->gemm_impl == 1 && gemm_tc_supported(g)) return launch_gemm_tc(g, st, e->num_sms, pdl_gemm);
    return launch_gemm_simt(g, st);
  };
  float* h = e->d_hA;
  const DecLayer& Llast = e->dec[c.n_dec_layers - 1];
  // head: h += post_ff(y) ; xn = final_norm(h)
  CU(launch_norm(h, e->d_y, Llast.g_post_ff, e->g_dec_final, h, e->d_xn, nullptr, B, d, c.rms_eps, st, pdl_norm, nullptr, 0, nullptr, 0, next_trace())); nl++;
  CU(G(e->d_xn, e->head_w1, d, d, GE_BIAS_GELU_BF16, e->head_b1, e->d_t1_bf, d));
  CU(G(e->d_t1_bf, e->head_w2, e->Vpad, d, GE_BIAS_F32, e->head_b2, e->d_logits, e->Vpad));
  { SamplerArgs s{}; s.logits = e->d_logits; s.ld = e->Vpad; s.V = e->V; s.slots = e->d_slots; s.topk_sched_pool = e->d_topk_pool;
    s.eos = c.eos_token; s.encodec_sr = c.encodec_sr; s.text_guard = c.text_guard_frames_per_token; s.progress_scale = c.progress_scale;
    s.tokens_out = e->d_tokens; s.tokens_stride = c.max_dec_len; s.argmax_out = nullptr; s.rows = B; s.host_mirror = e->d_mirror;
    s.picks_out = e->d_picks; s.forced_pool = e->d_forced; s.rope_out = e->d_rope; s.inv_freq = e->inv_freq; s.head_dim = D;
    s.scratch_u64 = e->d_samp_u64; s.scratch_f32 = e->d_samp_f32; s.trace = next_trace();
    CU(launch_sampler(s, st, pdl_samp)); nl++; }
  // no programmatic overlap with the sampler: later kernels read the slot state before their griddepcontrol.wait
  CU(launch_embed_slots(e->audio_emb, e->d_slots, sqrtf((float)d), h, B, d, st, false)); nl++;
  for (int l = 0; l < c.n_dec_layers; ++l) {
    const DecLayer& L = e->dec[l];
    if (l == 0) CU(launch_norm(h, nullptr, nullptr, L.g_pre_sa, nullptr, e->d_xn, nullptr, B, d, c.rms_eps, st, pdl_norm, nullptr, 0, nullptr, 0, next_trace()));
    else CU(launch_norm(h, e->d_y, e->dec[l - 1].g_post_ff, L.g_pre_sa, h, e->d_xn, nullptr, B, d, c.rms_eps, st, pdl_norm, nullptr, 0, nullptr, 0, next_trace()));
    nl++;
    CU(G(e->d_xn, L.wqkv, QKV, d, GE_F32, nullptr, e->d_qkv, QKV, e->d_y, (size_t)B * d, e->d_qc, (size_t)B * QD));
    { AttnDecodeArgs a{}; a.pool = e->pool; a.layer = l; a.block_table = e->d_self_bt; a.bt_stride = e->max_self_pages; a.q = e->d_qkv; a.q_stride = QKV;
      a.kv_new = e->d_qkv + QD; a.kv_stride = QKV; a.slots = e->d_slots; a.B = B; a.Hq = e->Hq; a.Hkv = e->Hkv; a.D = D; a.n_splits = e->ns_self;
      a.is_cross = 0; a.window = c.dec_layer_sliding[l] ? c.sliding_window : 0; a.scale = c.attn_scale; a.softcap = c.attn_softcap; a.inv_freq = e->inv_freq; a.rope_cs = e->d_rope;
      a.row_order = e->d_order_self;
      a.n_splits = 1;   // batched rows: key ranges are cut into chunks (grid.y), not into cluster splits
      if (e->attn_chunk > 0) { a.chunk_tokens = e->attn_chunk; a.max_chunks = e->chunks_self; a.part_o = e->d_part_o; a.part_ml = e->d_part_ml; a.part_cnt = e->d_part_cnt; }
      a.probe = (e->use_trace && l == 5) ? e->d_trace + 300 : nullptr;
      a.out = nullptr; a.out_bf = e->d_attn_bf; a.preload = 1; a.mma = e->attn_mma; a.trace = next_trace();
      a.n_layers_pool = c.n_dec_layers;
      if (a.mma && e->attn_tma && attn_decode_tma_supported(a)) CU(launch_attn_decode_tma(a, st, pdl_attn));
      else if (a.mma && attn_decode_mma_supported(a)) CU(launch_attn_decode_mma(a, st, pdl_attn));
      else CU(launch_attn_decode(a, st, pdl_attn));
      nl++; }
    CU(G(e->d_attn_bf, L.wo, d, QD, GE_F32, nullptr, e->d_y, d));
    CU(launch_norm(h, e->d_y, L.g_post_sa, L.g_pre_ca, h, e->d_xn, nullptr, B, d, c.rms_eps, st, pdl_norm, nullptr, 0, nullptr, 0, next_trace())); nl++;
    CU(G(e->d_xn, L.wq_c, QD, d, GE_F32, nullptr, e->d_qc, QD, e->d_y, (size_t)B * d));
    { AttnDecodeArgs a{}; a.pool = e->pool; a.layer = l; a.block_table = e->d_cross_bt; a.bt_stride = e->max_cross_pages; a.q = e->d_qc; a.q_stride = QD;
      a.kv_new = nullptr; a.kv_stride = 0; a.slots = e->d_slots; a.B = B; a.Hq = e->Hq; a.Hkv = e->Hkv; a.D = D; a.n_splits = e->ns_cross;
      a.is_cross = 1; a.window = 0; a.scale = c.attn_scale; a.softcap = c.attn_softcap; a.inv_freq = e->inv_freq; a.rope_cs = e->d_rope;
      a.row_order = e->d_order_cross;
      a.n_splits = 1;
      if (e->attn_chunk > 0) { a.chunk_tokens = e->attn_chunk; a.max_chunks = e->chunks_cross; a.part_o = e->d_part_o; a.part_ml = e->d_part_ml; a.part_cnt = e->d_part_cnt; }
      a.out = nullptr; a.out_bf = e->d_attn_bf; a.preload = 1; a.mma = e->attn_mma; a.trace = next_trace();
      a.n_layers_pool = c.n_dec_layers;
      if (a.mma && e->attn_tma && attn_decode_tma_supported(a)) CU(launch_attn_decode_tma(a, st, pdl_attn));
      else if (a.mma && attn_decode_mma_supported(a)) CU(launch_attn_decode_mma(a, st, pdl_attn));
      else CU(launch_attn_decode(a, st, pdl_attn));
      nl++; }
    CU(G(e->d_attn_bf, L.wo_c, d, QD, GE_F32, nullptr, e->d_y, d));
    CU(launch_norm(h, e->d_y, L.g_post_ca, L.g_pre_ff, h, e->d_xn, nullptr, B, d, c.rms_eps, st, pdl_norm, nullptr, 0, nullptr, 0, next_trace())); nl++;
    CU(G(e->d_xn, L.wgu, 2 * I, d, GE_GEGLU_BF16, nullptr, e->d_act_bf, I, e->d_y, (size_t)B * d, e->d_qkv, (size_t)B * QKV));
    CU(G(e->d_act_bf, L.wd, d, I, GE_F32, nullptr, e->d_y, d));
  }
  *n_launch = nl;
  return T5G_OK;
}

}  // namespace

extern "C" int t5g_decode(T5GEngine* e, int max_steps, void* stream_) {
  T5G_CHECK(e && max_steps > 0, T5G_ERR_INVALID, "bad arguments");
  T5G_CHECK(e->finalized, T5G_ERR_STATE, "weights not finalized");
  T5G_CUDA(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream_;
  CU(cudaEventRecord(e->ev[3], st));
  if (e->c.max_slots > 4) {
    // rows by descending key count (host's last known lengths: prompt + tokens generated at the last poll)
    const int B = e->c.max_slots;
    std::vector<int> os(B), oc(B), ls(B), lc(B);
    for (int s = 0; s < B; ++s) {
      os[s] = oc[s] = s;
      ls[s] = e->hslots[s].in_use ? e->hslots[s].n_dec + e->h_mirror[s * 8 + 2] : -1;
      lc[s] = e->hslots[s].in_use ? e->hslots[s].n_text : -1;
    }
    std::stable_sort(os.begin(), os.end(), [&](int x, int y) { return ls[x] > ls[y]; });
    std::stable_sort(oc.begin(), oc.end(), [&](int x, int y) { return lc[x] > lc[y]; });
    CU(cudaMemcpyAsync(e->d_order_self, os.data(), sizeof(int) * B, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->d_order_cross, oc.data(), sizeof(int) * B, cudaMemcpyHostToDevice, st));
    if (e->attn_chunk > 0) {
      // grid.y of the attention kernels: enough chunks for the longest row at the END of this call (a row that outgrows
      // the bound is still computed correctly: its last chunk takes the remainder)
      int longest = 1, longest_text = 1;
      for (int s = 0; s < B; ++s) if (e->hslots[s].in_use) {
        longest = std::max(longest, ls[s] + (e->decode_pending ? e->pending_steps : 0) + max_steps + 1);
        longest_text = std::max(longest_text, lc[s]);
      }
      longest = std::min(longest, e->c.max_dec_len);
      e->chunks_self = std::min(16, cdiv(longest, e->attn_chunk));
      e->chunks_cross = std::min(16, cdiv(longest_text, e->attn_chunk));
    }
  }
  auto enqueue = [&](cudaStream_t s_, int* nl) -> int {
    return (e->c.max_slots > 4) ? enqueue_step_batched(e, s_, nl) : enqueue_step(e, s_, nl);
  };
  if (e->use_graph) {
    // one graph = one step, plus a graph of `graph_steps` consecutive steps: inside it the head of step t+1 is a
    // programmatic dependent of the last kernel of step t, which removes the ~20 us gap between graph launches
    auto get_graph = [&](cudaGraphExec_t& graph, int& nodes, int steps) -> int {
      T5GEngine::StepGraph& sg = e->graphs[(steps * 32 + e->chunks_self) * 32 + e->chunks_cross];
      if (sg.exec) { graph = sg.exec; nodes = sg.nodes; return T5G_OK; }
      cudaStream_t cs;
      CU(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
      int nl_total = 0;
      cudaGraph_t g = nullptr;
      CU(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
      int rc = T5G_OK;
      for (int i = 0; i < steps && rc == T5G_OK; ++i) { int nl = 0; rc = enqueue(cs, &nl); nl_total += nl; }
      cudaError_t er = cudaStreamEndCapture(cs, &g);
      if (rc) { if (g) cudaGraphDestroy(g); cudaStreamDestroy(cs); return rc; }
      CU(er);
      CU(cudaGraphInstantiate(&graph, g, 0));
      cudaGraphDestroy(g);
      cudaStreamDestroy(cs);
      nodes = nl_total;
      sg.exec = graph; sg.nodes = nodes;
      return T5G_OK;
    };
    const int MS = (e->use_trace || e->graph_steps < 2) ? 1 : e->graph_steps;
    cudaGraphExec_t g1 = nullptr, gm = nullptr;
    int n1 = 0, nm = 0;
    int done = 0;
    if (MS > 1 && max_steps >= MS) {
      int rc = get_graph(gm, nm, MS);
      if (rc) return rc;
      for (; max_steps - done >= MS; done += MS) CU(cudaGraphLaunch(gm, st));
      e->launches += (int64_t)nm * (done / MS);
      e->last_nodes_per_step = nm / MS;
    }
    if (done < max_steps) {
      int rc = get_graph(g1, n1, 1);
      if (rc) return rc;
      e->launches += (int64_t)n1 * (max_steps - done);
      for (; done < max_steps; ++done) CU(cudaGraphLaunch(g1, st));
      e->last_nodes_per_step = n1;
    }
  } else {
    for (int i = 0; i < max_steps; ++i) {
      int nl = 0;
      int rc = enqueue(st, &nl);
      if (rc) return rc;
      e->launches += nl;
      e->last_nodes_per_step = nl;
    }
  }
  CU(cudaEventRecord(e->ev[4], st));
  e->decode_pending = true; e->pending_steps = max_steps;
  return T5G_OK;
}

extern "C" int t5g_poll(T5GEngine* e, T5GSlotState* states, void* stream_) {
  T5G_CHECK(e && states, T5G_ERR_INVALID, "bad arguments");
  T5G_CUDA(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize((cudaStream_t)stream_));
  if (cudaEventQuery(e->ev[4]) == cudaSuccess) {
    cudaEventElapsedTime(&e->timings[3], e->ev[3], e->ev[4]);
    if (e->decode_pending) { e->tot_decode_ms += e->timings[3]; e->tot_decode_steps += e->pending_steps; e->decode_pending = false; }
  }
  cudaGetLastError();
  for (int s = 0; s < e->c.max_slots; ++s) {
    states[s].active = e->h_mirror[s * 8 + 0]; states[s].finished = e->h_mirror[s * 8 + 1];
    states[s].n_generated = e->h_mirror[s * 8 + 2]; states[s].cur_len = e->h_mirror[s * 8 + 3];
    // device-side error flags (SlotDev.error): 2 = sampler scratch missing, 4 = grid barrier of gemv_pair timed out
    const int err = e->h_mirror[s * 8 + 4];
    T5G_CHECK((err & 6) == 0, T5G_ERR_STATE, "slot %d: device-side error flags 0x%x (2: sampler scratch missing, 4: grid barrier timeout)", s, err);
  }
  return T5G_OK;
}

extern "C" int t5g_read_tokens(T5GEngine* e, int slot, int32_t* out, int max_tokens, int* n_out, void* stream_) {
  T5G_CHECK(e && out && n_out && slot >= 0 && slot < e->c.max_slots, T5G_ERR_INVALID, "bad arguments");
  CU(cudaStreamSynchronize((cudaStream_t)stream_));
  const int n = std::min(max_tokens, e->h_mirror[slot * 8 + 2]);
  memcpy(out, e->h_tokens + (size_t)slot * e->c.max_dec_len, sizeof(int) * n);
  *n_out = n;
  return T5G_OK;
}

extern "C" int t5g_read_picks(T5GEngine* e, int slot, int32_t* out, int max_tokens, int* n_out, void* stream_) {
  T5G_CHECK(e && out && n_out && slot >= 0 && slot < e->c.max_slots, T5G_ERR_INVALID, "bad arguments");
  CU(cudaStreamSynchronize((cudaStream_t)stream_));
  const int n = std::min(max_tokens, e->h_mirror[slot * 8 + 2]);
  memcpy(out, e->h_picks + (size_t)slot * e->c.max_dec_len, sizeof(int) * n);
  *n_out = n;
  return T5G_OK;
}

extern "C" int t5g_release_slot(T5GEngine* e, int slot) {
  T5G_CHECK(e && slot >= 0 && slot < e->c.max_slots, T5G_ERR_INVALID, "bad slot");
  T5G_CUDA(cudaSetDevice(e->device));
  HostSlot& hs = e->hslots[slot];
  free_slot_pages(e, hs);
  hs.in_use = false;
  SlotDev sd; memset(&sd, 0, sizeof(sd));
  e->h_slots[slot] = sd;
  CU(cudaMemcpy(e->d_slots + slot, &sd, sizeof(sd), cudaMemcpyHostToDevice));
  for (int i = 0; i < 8; ++i) e->h_mirror[slot * 8 + i] = 0;
  return T5G_OK;
}

extern "C" int t5g_read_memory(T5GEngine* e, int slot, float* out, void* stream_) {
  T5G_CHECK(e && out && slot >= 0 && slot < e->c.max_slots, T5G_ERR_INVALID, "bad arguments");
  const HostSlot& hs = e->hslots[slot];
  T5G_CHECK(hs.in_use && hs.prefill_id == e->prefill_counter && hs.mem_off >= 0, T5G_ERR_STATE, "slot %d was not part of the last prefill", slot);
  CU(cudaStreamSynchronize((cudaStream_t)stream_));
  CU(cudaMemcpy(out, e->p_memory + (size_t)hs.mem_off * e->d, sizeof(float) * (size_t)hs.n_text * e->d, cudaMemcpyDeviceToHost));
  return T5G_OK;
}

extern "C" int t5g_read_last_hidden(T5GEngine* e, int slot, float* out, void* stream_) {
  T5G_CHECK(e && out && slot >= 0 && slot < e->c.max_slots, T5G_ERR_INVALID, "bad arguments");
  const HostSlot& hs = e->hslots[slot];
  T5G_CHECK(hs.in_use && hs.prefill_id == e->prefill_counter, T5G_ERR_STATE, "slot %d was not part of the last prefill", slot);
  CU(cudaStreamSynchronize((cudaStream_t)stream_));
  CU(cudaMemcpy(out, e->p_final + (size_t)(hs.dec_off + hs.n_dec - 1) * e->d, sizeof(float) * e->d, cudaMemcpyDeviceToHost));
  return T5G_OK;
}

extern "C" int t5g_read_logits(T5GEngine* e, int slot, float* out, void* stream_) {
  T5G_CHECK(e && out && slot >= 0 && slot < e->c.max_slots, T5G_ERR_INVALID, "bad arguments");
  CU(cudaStreamSynchronize((cudaStream_t)stream_));
  CU(cudaMemcpy(out, e->d_logits + (size_t)slot * e->Vpad, sizeof(float) * e->V, cudaMemcpyDeviceToHost));
  return T5G_OK;
}

extern "C" int t5g_prefill_logits(T5GEngine* e, int slot, float* out, void* stream_) {
  T5G_CHECK(e && out && slot >= 0 && slot < e->c.max_slots, T5G_ERR_INVALID, "bad arguments");
  const HostSlot& hs = e->hslots[slot];
  T5G_CHECK(hs.in_use && hs.prefill_id == e->prefill_counter, T5G_ERR_STATE, "slot %d was not part of the last prefill", slot);
  cudaStream_t st = (cudaStream_t)stream_;
  const int d = e->d;
  for (int t0 = 0; t0 < hs.n_dec; t0 += e->logits_chunk) {
    const int n = std::min(e->logits_chunk, hs.n_dec - t0);
    const float* src = e->p_final + (size_t)(hs.dec_off + t0) * d;
    CU(launch_f32_to_bf16(src, e->p_xn, (size_t)n * d, st)); e->launches++;
    CU(gemm(e, e->p_xn, e->head_w1, n, d, d, GE_BIAS_GELU_BF16, e->head_b1, e->p_act, d, st));
    CU(gemm(e, e->p_act, e->head_w2, n, e->Vpad, d, GE_BIAS_F32, e->head_b2, e->p_logits, e->Vpad, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaMemcpy2D(out + (size_t)t0 * e->V, sizeof(float) * e->V, e->p_logits, sizeof(float) * e->Vpad, sizeof(float) * e->V, n, cudaMemcpyDeviceToHost));
  }
  return T5G_OK;
}

extern "C" int t5g_sample(T5GEngine* e, float* logits, const T5GSampleRow* rows, int n_rows, int32_t* out_tokens,
                          int32_t* out_argmax, void* stream_) {
  T5G_CHECK(e && logits && rows && out_tokens && n_rows > 0 && n_rows <= 4096, T5G_ERR_INVALID, "bad arguments");
  T5G_CUDA(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream_;
  const T5GConfig& c = e->c;
  std::vector<SlotDev> sd(n_rows);
  std::vector<float> us(n_rows);
  for (int i = 0; i < n_rows; ++i) {
    const T5GSampleRow& r = rows[i];
    const bool minp = r.sampling.min_p > 0.f && r.sampling.min_p < 1.f;
    T5G_CHECK(r.sampling.temperature > 0.f, T5G_ERR_INVALID, "row %d: temperature must be > 0", i);
    (void)minp;
    SlotDev& s = sd[i]; memset(&s, 0, sizeof(s));
    s.active = 1; s.n_generated = r.cur_num_gen; s.cur_len = r.current_length; s.prompt_offset = r.prompt_offset; s.target_total = r.target_total;
    s.est_total = std::max(r.target_total + 1, 1); s.n_text = r.n_text;
    s.budget_limit = (int)std::floor((double)r.target_total - (double)r.prompt_offset + (double)c.encodec_sr * (double)c.extra_cutoff);
    s.top_k = r.sampling.top_k; s.top_p = r.sampling.top_p; s.min_p = r.sampling.min_p; s.temperature = r.sampling.temperature;
    s.uniforms = e->d_sample_u + i; s.n_uniforms = 1; s.topk_sched_off = -1;
    s.prev_token = r.prev_token; s.consec_silence = r.consec_silence_count; s.stop_repetition = e->sample_stop_repetition;
    s.silence_off = 0; s.n_silence = e->n_sample_silence;
    // the kernel indexes uniforms by n_generated (clamped to n_uniforms-1 = 0)
    us[i] = r.u;
  }
  int *d_tok = nullptr, *d_amax = nullptr;
  CU(cudaMalloc(&d_tok, sizeof(int) * n_rows * 2));
  d_amax = d_tok + n_rows;
  CU(cudaMemcpyAsync(e->d_sample_u, us.data(), sizeof(float) * n_rows, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(e->d_sample_slots, sd.data(), sizeof(SlotDev) * n_rows, cudaMemcpyHostToDevice, st));
  SamplerArgs s{}; s.logits = logits; s.ld = e->V; s.V = e->V; s.slots = e->d_sample_slots; s.topk_sched_pool = e->d_sample_silence;
  s.eos = c.eos_token; s.encodec_sr = c.encodec_sr; s.text_guard = c.text_guard_frames_per_token; s.progress_scale = c.progress_scale;
  s.tokens_stride = 1; s.flat_tokens = 1; s.host_mirror = nullptr;
  s.scratch_u64 = e->d_samp_u64; s.scratch_f32 = e->d_samp_f32;
  cudaError_t er = cudaSuccess;
  for (int off = 0; off < n_rows && er == cudaSuccess; off += e->samp_scratch_rows) {     // scratch covers samp_scratch_rows rows
    SamplerArgs c2 = s;
    c2.rows = std::min(e->samp_scratch_rows, n_rows - off);
    c2.logits = logits + (size_t)off * e->V; c2.slots = e->d_sample_slots + off;
    c2.tokens_out = d_tok + off; c2.argmax_out = d_amax + off;
    er = launch_sampler(c2, st, false);
    e->launches++;
  }
  if (er == cudaSuccess) er = cudaMemcpyAsync(out_tokens, d_tok, sizeof(int) * n_rows, cudaMemcpyDeviceToHost, st);
  if (er == cudaSuccess && out_argmax) er = cudaMemcpyAsync(out_argmax, d_amax, sizeof(int) * n_rows, cudaMemcpyDeviceToHost, st);
  if (er == cudaSuccess) er = cudaStreamSynchronize(st);
  cudaFree(d_tok);
  CU(er);
  return T5G_OK;
}

extern "C" int t5g_sample_set_silence(T5GEngine* e, const int32_t* toks, int n, int stop_repetition) {
  T5G_CHECK(e && n >= 0 && n <= 256 && (n == 0 || toks), T5G_ERR_INVALID, "bad arguments (at most 256 silence tokens)");
  T5G_CUDA(cudaSetDevice(e->device));
  if (n > 0) CU(cudaMemcpy(e->d_sample_silence, toks, sizeof(int) * n, cudaMemcpyHostToDevice));
  e->n_sample_silence = n; e->sample_stop_repetition = stop_repetition;
  return T5G_OK;
}

extern "C" int64_t t5g_launch_count(const T5GEngine* e) { return e ? e->launches : 0; }

extern "C" int64_t t5g_weight_bytes_per_step(const T5GEngine* e) {
  if (!e) return 0;
  const int64_t d = e->d, I = e->I, QD = e->QD, QKV = e->QKV;
  // decoder layers without the cross K/V projections, six norm gains per layer (bf16 in the
  // reference's accounting), final norm, head (weights + biases), one embedding row
  int64_t per_layer = QKV * d + d * QD + QD * d + d * QD + 2 * I * d + d * I + 6 * d;
  int64_t params = per_layer * e->c.n_dec_layers + d + (d * d + d) + ((int64_t)e->V * d + e->V) + d;
  return params * 2;
}

extern "C" int64_t t5g_kv_bytes_per_token(const T5GEngine* e) {
  if (!e) return 0;
  return (int64_t)e->c.n_dec_layers * 2 * e->KD * 2;
}

extern "C" int t5g_get_timings(T5GEngine* e, float* out) {
  T5G_CHECK(e && out, T5G_ERR_INVALID, "bad arguments");
  for (int i = 0; i < 4; ++i) out[i] = e->timings[i];
  return T5G_OK;
}

extern "C" int t5g_get_counters(T5GEngine* e, double* out) {
  T5G_CHECK(e && out, T5G_ERR_INVALID, "bad arguments");
  out[0] = e->tot_prefill_ms; out[1] = e->tot_decode_ms; out[2] = (double)e->tot_decode_steps; out[3] = (double)e->tot_prefill_calls;
  out[4] = (double)e->launches; out[5] = (double)e->last_nodes_per_step; out[6] = 0; out[7] = 0;
  return T5G_OK;
}

extern "C" int t5g_debug_trace(T5GEngine* e, uint64_t* begin_ns, uint64_t* end_ns, int max_entries, int* n_out) {
  T5G_CHECK(e && begin_ns && end_ns && n_out, T5G_ERR_INVALID, "bad arguments");
  T5G_CHECK(e->use_trace, T5G_ERR_STATE, "tracing is off (set T5G_TRACE=1 before creating the engine)");
  T5G_CUDA(cudaSetDevice(e->device));
  CU(cudaDeviceSynchronize());
  const int n = std::min(std::min(max_entries, e->last_nodes_per_step > 0 ? e->last_nodes_per_step : T5G_TRACE_STRIDE), T5G_TRACE_STRIDE);
  const int ncopy = std::min(max_entries, T5G_TRACE_STRIDE);   // entries past the kernel count hold optional in-kernel probes
  CU(cudaMemcpy(begin_ns, e->d_trace, sizeof(uint64_t) * ncopy, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(end_ns, e->d_trace + T5G_TRACE_STRIDE, sizeof(uint64_t) * ncopy, cudaMemcpyDeviceToHost));
  *n_out = n;
  return T5G_OK;
}

extern "C" int t5g_debug_gemm(T5GEngine* e, const void* x, const void* w, float* out, int M, int N, int K, int impl, void* stream_) {
  T5G_CHECK(e && x && w && out, T5G_ERR_INVALID, "bad arguments");
  T5G_CUDA(cudaSetDevice(e->device));
  // impl: low byte 1 = tcgen05 path, 0 = CUDA-core cross-check; bits 8-15 = epilogue kind (GE_*, no bias; `out` is fp32 [M,N],
  // bf16 [M,N] for GE_BF16 or bf16 [M,N/2] for GE_GEGLU_BF16)
  const int epi = (impl >> 8) & 0xff;
  impl &= 0xff;
  T5G_CHECK(epi == GE_F32 || ((epi == GE_BF16 || epi == GE_GEGLU_BF16) && impl == 1), T5G_ERR_INVALID, "unsupported debug epilogue %d", epi);
  GemmArgs g{(const bf16*)x, (const bf16*)w, M, N, K, epi, nullptr, out, epi == GE_GEGLU_BF16 ? N / 2 : N, 0};
  e->launches++;
  CU(impl == 1 ? launch_gemm_tc(g, (cudaStream_t)stream_, e->num_sms) : launch_gemm_simt(g, (cudaStream_t)stream_));
  return T5G_OK;
}

extern "C" int t5g_debug_attn_prefill(T5GEngine* e, const void* q, const void* k, const void* v, const int32_t* q_seg_off,
                                      const int32_t* k_seg_off, const int32_t* q_seg_of, int n_seg, int Tq, int Tk, int max_lq,
                                      int causal, int window, float softcap, void* out, int impl, void* stream_) {
  T5G_CHECK(e && q && k && v && out && q_seg_off && k_seg_off && q_seg_of, T5G_ERR_INVALID, "bad arguments");
  T5G_CUDA(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream_;
  AttnPrefillArgs a{}; a.q = (const bf16*)q; a.k = (const bf16*)k; a.v = (const bf16*)v; a.q_seg_off = q_seg_off; a.k_seg_off = k_seg_off;
  a.q_seg_of = q_seg_of; a.Tq = Tq; a.Hq = e->Hq; a.Hkv = e->Hkv; a.D = e->D; a.causal = causal; a.window = window;
  a.scale = e->c.attn_scale; a.softcap = softcap; a.out = (bf16*)out;
  if (impl == 1) {
    T5G_CHECK(attn_prefill_tc_supported(e->D), T5G_ERR_UNSUPPORTED, "tensor-core prefill attention needs head_dim 64/128/256");
    T5G_CHECK(n_seg <= e->c.max_slots + 64 && n_seg >= 1, T5G_ERR_INVALID, "too many segments");
    cudaError_t er = launch_attn_prefill_tc(a, Tk, n_seg, max_lq, st);
    if (er == cudaSuccess) er = cudaStreamSynchronize(st);
    CU(er);
  } else {
    CU(launch_attn_prefill(a, st));
  }
  e->launches++;
  return T5G_OK;
}

extern "C" int t5g_debug_gemv_gateup(T5GEngine* e, const void* w, void* stream_) {
  T5G_CHECK(e && w && e->finalized, T5G_ERR_INVALID, "bad arguments / weights not finalized");
  T5G_CUDA(cudaSetDevice(e->device));
  const DecLayer& L = e->dec[0];
  GemvArgs a{}; a.W = (const bf16*)w; a.N = 2 * e->I; a.K = e->d; a.B = 1; a.h_in = e->d_hA; a.y = e->d_y; a.g_post = L.g_post_ca;
  a.g_pre = L.g_pre_ff; a.h_out = nullptr; a.eps = e->c.rms_eps; a.out = e->d_act; a.out_stride = e->I; a.slots = nullptr;
  e->launches++;
  CU(launch_gemv(a, P_RES_NORM, E_GEGLU, e->num_sms, (cudaStream_t)stream_, e->use_pdl));
  return T5G_OK;
}

extern "C" int t5g_debug_gemv(T5GEngine* e, const float* x, const void* w, float* out, int B, int N, int K, void* stream_) {
  T5G_CHECK(e && x && w && out && B >= 1 && B <= 4, T5G_ERR_INVALID, "bad arguments");
  T5G_CUDA(cudaSetDevice(e->device));
  GemvArgs a{}; a.W = (const bf16*)w; a.N = N; a.K = K; a.B = B; a.x = x; a.out = out; a.out_stride = N; a.slots = nullptr;
  e->launches++;
  CU(launch_gemv(a, P_PLAIN, E_STORE, e->num_sms, (cudaStream_t)stream_, e->use_pdl));
  return T5G_OK;
}
