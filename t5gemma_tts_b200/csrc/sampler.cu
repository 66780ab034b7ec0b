// Fused sampling step S: eos edits -> temperature -> (min-p | top-k -> top-p) -> inverse-CDF draw with an
// explicit uniform -> stop rules -> per-slot state update.  One CTA per batch row, no host sync.
// Replaces sample_helper (models/t5gemma.py:971-1055) + topk_sampling/top_k_top_p_filtering
// (models/utils.py:53-122): ~15 ATen launches and 2-3 .item() syncs per token in the reference.
//
// Bit-exactness contract (oracle/sampler_oracle.py): every floating-point op that decides the outcome is
// an individually rounded fp32 op (__fadd_rn/__fmul_rn/__fdiv_rn: never contracted into fma), exp is the
// fixed polynomial det_exp(), and sums run in the oracle's stated order.  Top-k uses an exact radix
// select over order-preserving integer keys; survivors are ordered (value desc, index asc).
#include "kernels.h"

namespace {

constexpr int SAMP_THREADS = 1024;
constexpr int CAND_CAP = 2048;       // max survivors handled on chip (k + ties)
constexpr int CHUNK = 256;           // oracle's full-vocabulary summation chunk
constexpr int MAX_CHUNKS = 1024;

__device__ __forceinline__ float det_exp(float x) {
  if (x < -86.0f) return 0.f;
  const float LOG2E = __int_as_float(0x3fb8aa3b), LN2_HI = __int_as_float(0x3f318000),
              LN2_LO = __int_as_float(0xb95e8083);
  const float n = rintf(__fmul_rn(x, LOG2E));
  float r = __fsub_rn(x, __fmul_rn(n, LN2_HI));
  r = __fsub_rn(r, __fmul_rn(n, LN2_LO));
  float p = __int_as_float(0x39506967);
  p = __fadd_rn(__fmul_rn(p, r), __int_as_float(0x3ab743ce));
  p = __fadd_rn(__fmul_rn(p, r), __int_as_float(0x3c088908));
  p = __fadd_rn(__fmul_rn(p, r), __int_as_float(0x3d2aa9c1));
  p = __fadd_rn(__fmul_rn(p, r), __int_as_float(0x3e2aaaaa));
  p = __fadd_rn(__fmul_rn(p, r), __int_as_float(0x3f000000));
  float y = __fmul_rn(p, __fmul_rn(r, r));
  y = __fadd_rn(y, r);
  y = __fadd_rn(y, 1.0f);
  return __fmul_rn(y, __int_as_float(((int)n + 127) << 23));
}

__device__ __forceinline__ unsigned key_of(float z) {
  unsigned u = (z == 0.f) ? 0u : __float_as_uint(z);      // -0 == +0
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float val_of(unsigned key) {
  unsigned u = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
  return __uint_as_float(u);
}

constexpr int HI_CAP = 73728;        // vocabularies up to this size use the shared-memory fast path

struct Shared {
  unsigned short hi16[HI_CAP];       // top 16 bits of every order-preserving key (one global pass)
  unsigned tmax[SAMP_THREADS];       // per-thread maxima (lower-bound selection for 32 < k <= 1024)
  unsigned redk[32];
  unsigned lb; int n_surv;
  unsigned whist[32][16];            // full-sort path: per-warp digit counts / bases
  unsigned hist[256];
  unsigned long long cand[CAND_CAP];
  float e[CAND_CAP];
  float csum[MAX_CHUNKS];
  float redv[32]; int redi[32];
  unsigned prefix; int kth; unsigned n_cand; int overflow;
  float zmax; int amax; int n_keep; float total;
};


// ---- general path helpers (survivor sets larger than CAND_CAP: nucleus sampling without top-k, top_k > 1024,
//      pathological ties): stable LSD radix sort (4-bit digits, 8 passes) of all V composites in global scratch by
//      one CTA.  Every warp owns a contiguous block so that ballot/match ranking keeps the sort stable. ------------
__device__ __noinline__ void full_sort_desc(unsigned long long* src, unsigned long long* dst, int V, Shared& S, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  const int BL = ((V + 31) / 32 + 31) / 32 * 32;                 // block per warp, multiple of 32
  const int w_begin = min(V, warp * BL), w_end = min(V, w_begin + BL);
  for (int shift = 32; shift < 64; shift += 4) {
    if (tid < 512) (&S.whist[0][0])[tid] = 0u;
    __syncthreads();
    for (int i0 = w_begin; i0 < w_end; i0 += 32) {
      const int i = i0 + lane;
      const bool ok = i < w_end;
      const unsigned d = ok ? 15u - (unsigned)((src[i] >> shift) & 15ull) : 0u;
      const unsigned act = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const unsigned peers = __match_any_sync(act, d);
        if ((int)(__ffs(peers) - 1) == lane) S.whist[warp][d] += (unsigned)__popc(peers);
      }
      __syncwarp();
    }
    __syncthreads();
    if (tid == 0) {                                              // exclusive scan, digit-major then warp
      unsigned run = 0;
      for (int d = 0; d < 16; ++d)
        for (int w = 0; w < 32; ++w) { const unsigned c = S.whist[w][d]; S.whist[w][d] = run; run += c; }
    }
    __syncthreads();
    for (int i0 = w_begin; i0 < w_end; i0 += 32) {
      const int i = i0 + lane;
      const bool ok = i < w_end;
      const unsigned long long v = ok ? src[i] : 0ull;
      const unsigned d = ok ? 15u - (unsigned)((v >> shift) & 15ull) : 0u;
      const unsigned act = __ballot_sync(0xffffffffu, ok);
      unsigned peers = 0, base = 0;
      if (ok) {
        peers = __match_any_sync(act, d);
        base = S.whist[warp][d];
      }
      __syncwarp();
      if (ok) {
        dst[base + (unsigned)__popc(peers & ((1u << lane) - 1u))] = v;
        if ((int)(__ffs(peers) - 1) == lane) S.whist[warp][d] = base + (unsigned)__popc(peers);
      }
      __syncwarp();
    }
    __syncthreads();
    unsigned long long* t = src; src = dst; dst = t;
  }
  // 8 passes: the result is back in the original `src` buffer
}

// sequential left-to-right fp32 sum of e[0..n) by ONE thread (the oracle's order), 8 loads in flight
__device__ float seq_sum_global(const float* e, int n) {
  float s = 0.f;
  int i = 0;
  for (; i + 8 <= n; i += 8) {
    const float4 a = *reinterpret_cast<const float4*>(e + i), b = *reinterpret_cast<const float4*>(e + i + 4);
    s = __fadd_rn(s, a.x); s = __fadd_rn(s, a.y); s = __fadd_rn(s, a.z); s = __fadd_rn(s, a.w);
    s = __fadd_rn(s, b.x); s = __fadd_rn(s, b.y); s = __fadd_rn(s, b.z); s = __fadd_rn(s, b.w);
  }
  for (; i < n; ++i) s = __fadd_rn(s, e[i]);
  return s;
}

// first i in [0,n) whose running sequential sum (after adding p[i]) exceeds x, else -1; ONE thread, oracle order
__device__ int seq_first_exceed(const float* p, int n, float x) {
  float c = 0.f;
  int i = 0;
  for (; i + 8 <= n; i += 8) {
    const float4 a = *reinterpret_cast<const float4*>(p + i), b = *reinterpret_cast<const float4*>(p + i + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) { c = __fadd_rn(c, v[j]); if (c > x) return i + j; }
  }
  for (; i < n; ++i) { c = __fadd_rn(c, p[i]); if (c > x) return i; }
  return -1;
}

// General path: all V tokens sorted (value desc, index asc); top-k / min-p / top-p are prefix cuts of that order and
// the sums run sequentially exactly as in oracle/sampler_oracle.py.  Called by every thread of the CTA.
__device__ __noinline__ int sample_large(const SamplerArgs& a, Shared& S, int row, int tid, const float* lg, int V, float T, bool scaled,
                            float zmax, int top_k, float top_p, float min_p, bool minp_mode, const SlotDev& sl, int n_gen) {
  const int Vs = (V + 7) & ~7;
  unsigned long long* A = a.scratch_u64 + (size_t)row * 2 * Vs;
  unsigned long long* Bf = A + Vs;
  float* E = a.scratch_f32 + (size_t)row * 2 * Vs;
  float* P = E + Vs;
  for (int i = tid; i < V; i += SAMP_THREADS) {
    const float v = lg[i];
    const float z = scaled ? __fdiv_rn(v, T) : v;
    A[i] = ((unsigned long long)key_of(z) << 32) | (0xffffffffu - (unsigned)i);
  }
  __syncthreads();
  full_sort_desc(A, Bf, V, S, tid);
  if (tid == 0) S.n_surv = V;
  __syncthreads();
  if (minp_mode) {
    const float tot = S.total;
    for (int i = tid; i < V; i += SAMP_THREADS) {
      const bool keep = !(__fdiv_rn(det_exp(__fsub_rn(val_of((unsigned)(A[i] >> 32)), zmax)), tot) < min_p);
      const bool keep_next = (i + 1 < V) && !(__fdiv_rn(det_exp(__fsub_rn(val_of((unsigned)(A[i + 1] >> 32)), zmax)), tot) < min_p);
      if (keep && !keep_next) S.n_surv = i + 1;
    }
  } else if (top_k > 0) {
    const int k = min(max(top_k, 1), V);
    const unsigned thr = (unsigned)(A[k - 1] >> 32);
    for (int i = tid; i < V; i += SAMP_THREADS) {
      const bool in = (unsigned)(A[i] >> 32) >= thr;
      const bool next_in = (i + 1 < V) && ((unsigned)(A[i + 1] >> 32) >= thr);
      if (in && !next_in) S.n_surv = i + 1;
    }
  }
  __syncthreads();
  const int n = S.n_surv;
  const float z0 = val_of((unsigned)(A[0] >> 32));
  for (int i = tid; i < n; i += SAMP_THREADS) E[i] = det_exp(__fsub_rn(val_of((unsigned)(A[i] >> 32)), z0));
  __syncthreads();
  if (tid == 0) S.n_keep = n;
  if (top_p < 1.0f && !minp_mode) {
    if (tid == 0) S.total = seq_sum_global(E, n);
    __syncthreads();
    const float s = S.total;
    for (int i = tid; i < n; i += SAMP_THREADS) P[i] = __fdiv_rn(E[i], s);
    __syncthreads();
    if (tid == 0) {
      const int j = seq_first_exceed(P, n - 1, top_p);      // cum through i exceeds p -> token i+1 (and later) removed
      S.n_keep = (j < 0) ? n : j + 1;
    }
  }
  __syncthreads();
  const int n_keep = S.n_keep;
  if (tid == 0) S.total = seq_sum_global(E, n_keep);
  __syncthreads();
  const float s2 = S.total;
  for (int i = tid; i < n_keep; i += SAMP_THREADS) P[i] = __fdiv_rn(E[i], s2);
  __syncthreads();
  if (tid == 0) {
    const float u = sl.uniforms ? sl.uniforms[min(n_gen, max(sl.n_uniforms - 1, 0))] : 0.5f;
    const int j = seq_first_exceed(P, n_keep, u);
    const int pick = (j < 0) ? n_keep - 1 : j;
    S.n_keep = (int)(0xffffffffu - (unsigned)(A[pick] & 0xffffffffull));
  }
  __syncthreads();
  return S.n_keep;
}

__global__ void __launch_bounds__(SAMP_THREADS, 1) sampler_kernel(SamplerArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  Shared& S = *reinterpret_cast<Shared*>(smraw);
  pdl_launch_dependents();
  pdl_wait();
  trace_begin(a.trace);
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#define SP_PROBE(k) do { if (a.probe && row == 0 && tid == 0) a.probe[k] = globaltimer_ns(); } while (0)
  SP_PROBE(0);
  // the slot record is staged in shared memory (one coalesced load) and written back once at the end: the stop-rule /
  // state-update code of thread 0 otherwise pays one L2 round trip per field (loads cannot move above its stores)
  __shared__ SlotDev sl_s;
  constexpr int SLOT_WORDS = sizeof(SlotDev) / 4;
  static_assert(sizeof(SlotDev) % 4 == 0 && SLOT_WORDS <= SAMP_THREADS, "SlotDev is copied word-wise");
  if (tid < SLOT_WORDS) reinterpret_cast<int*>(&sl_s)[tid] = reinterpret_cast<const int*>(&a.slots[row])[tid];
  __syncthreads();
  SlotDev& sl = sl_s;
  if (!sl.active) return;
  SP_PROBE(1);
  float* lg = a.logits + (size_t)row * a.ld;
  const int V = a.V, eos = a.eos;

  // (1) in-place eos edits (models/t5gemma.py:986-997)
  const int n_gen = sl.n_generated, cur_len = sl.cur_len;
  const int eff_len = max(0, cur_len - sl.prompt_offset);
  float u_draw = 0.5f;                                   // this step's uniform: requested now, consumed after the filters
  if (tid == 0 && sl.uniforms) u_draw = sl.uniforms[min(n_gen, max(sl.n_uniforms - 1, 0))];
  if (tid == 0) {
    if (eff_len == 0) lg[eos] = -1e9f;
    if (n_gen <= a.encodec_sr / 5) lg[eos] = -10000.0f;
    // silence-repetition penalty (models/t5gemma.py:999-1011), in place like the reference
    if (sl.stop_repetition > 0 && sl.n_silence > 0 && sl.prev_token >= 0 && sl.consec_silence > sl.stop_repetition) {
      bool in_set = false;
      for (int i = 0; i < sl.n_silence; ++i) in_set |= (a.topk_sched_pool[sl.silence_off + i] == sl.prev_token);
      if (in_set) {
        const float f = (float)(sl.consec_silence - (sl.stop_repetition - 1));
        const float v = lg[sl.prev_token];
        lg[sl.prev_token] = (v < 0.f) ? __fmul_rn(v, f) : __fdiv_rn(v, f);
      }
    }
    S.n_cand = 0; S.overflow = 0;
  }
  __syncthreads();

  SP_PROBE(2);
  // (2) one pipelined global pass: argmax of the adjusted logits (first index on ties), and -- for the
  //     top-k fast path -- the top 16 bits of every key into shared memory + per-thread max key
  const float T = sl.temperature;
  const bool scaled = (T != 1.0f);
  auto zof = [&](int i) -> float { float v = lg[i]; return scaled ? __fdiv_rn(v, T) : v; };
  const bool use_hi = V <= HI_CAP;
  // Keys of this pass are taken from the RAW logits: dividing by T > 0 is monotone, so the order (hence every maximum
  // and the lower bound derived from them) is that of z = v / T, and the pass saves an IEEE division per element (it is
  // instruction-issue bound: 64 elements per thread).  Distinct raw values may round to EQUAL z, and ties at the k-th
  // value must survive: the candidate filter below therefore lowers its bound by 2 keys (adjacent floats differ by 1 in
  // key space; values 3 ulps apart cannot collide); candidates are re-read and compared exactly in z.
  const bool vspace = !scaled || T > 0.f;
  float bv = -INFINITY; int bi = 0x7fffffff;
  unsigned mk = 0u;
  if (vspace && (a.ld & 3) == 0 && (reinterpret_cast<uintptr_t>(lg) & 15) == 0) {
    // 16-byte loads, 8 in flight per thread; elements of a thread are visited in increasing index order, so a strict
    // `>` keeps the first index of the maximum
    const float4* lg4 = reinterpret_cast<const float4*>(lg);
    const int n4 = (V + 3) >> 2;
    constexpr int PB = 8;
    for (int g0 = tid; g0 < n4; g0 += SAMP_THREADS * PB) {
      float4 v4[PB];
#pragma unroll
      for (int j = 0; j < PB; ++j) { const int g = g0 + j * SAMP_THREADS; v4[j] = (g < n4) ? lg4[g] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY); }
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        const int g = g0 + j * SAMP_THREADS;
        if (g < n4) {
          const float e4[4] = {v4[j].x, v4[j].y, v4[j].z, v4[j].w};
          unsigned kk[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int i = g * 4 + c;
            const bool ok = i < V;                         // the padded tail of the row is not part of the vocabulary
            const float v = ok ? e4[c] : -INFINITY;
            if (v > bv) { bv = v; bi = i; }
            kk[c] = ok ? key_of(v) : 0u;
            mk = max(mk, kk[c]);
          }
          if (use_hi) {
            uint2 pk;
            pk.x = (kk[0] >> 16) | (kk[1] & 0xffff0000u);
            pk.y = (kk[2] >> 16) | (kk[3] & 0xffff0000u);
            *reinterpret_cast<uint2*>(&S.hi16[g * 4]) = pk;
          }
        }
      }
    }
  } else {
    constexpr int PB = 16;
    for (int i0 = tid; i0 < V; i0 += SAMP_THREADS * PB) {
      float v[PB];
#pragma unroll
      for (int j = 0; j < PB; ++j) { const int i = i0 + j * SAMP_THREADS; v[j] = (i < V) ? lg[i] : -INFINITY; }
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        const int i = i0 + j * SAMP_THREADS;
        if (i < V) {
          if (v[j] > bv || (v[j] == bv && i < bi)) { bv = v[j]; bi = i; }
          const unsigned key = key_of(scaled ? __fdiv_rn(v[j], T) : v[j]);
          mk = max(mk, key);
          if (use_hi) S.hi16[i] = (unsigned short)(key >> 16);
        }
      }
    }
  }
  SP_PROBE(3);
  S.tmax[tid] = mk;
  {
    unsigned wk = mk;
    for (int o = 16; o > 0; o >>= 1) wk = max(wk, __shfl_xor_sync(0xffffffffu, wk, o));
    if (lane == 0) S.redk[warp] = wk;
  }
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, bv, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) { S.redv[warp] = bv; S.redi[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    bv = S.redv[lane]; bi = S.redi[lane];
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { S.amax = bi; S.zmax = bv; }
  }
  __syncthreads();
  const float zmax = scaled ? __fdiv_rn(S.zmax, T) : S.zmax;
  SP_PROBE(4);

  int top_k = sl.top_k;
  if (sl.topk_sched_off >= 0 && sl.n_topk_sched > 0)
    top_k = a.topk_sched_pool[sl.topk_sched_off + min(sl.n_topk_sched - 1, n_gen)];
  float top_p = sl.top_p;
  const float min_p = sl.min_p;
  const int nchunk = (V + CHUNK - 1) / CHUNK;
  bool have_csum = false;
  bool filtered = false, minp_mode = false;

  // (3) min-p (models/utils.py:72-80): full softmax with the oracle's chunked sum
  if (min_p > 0.f && min_p < 1.f) {
    for (int c = tid; c < nchunk; c += SAMP_THREADS) {
      float s = 0.f;
      const int hi = min(V, (c + 1) * CHUNK);
      for (int i = c * CHUNK; i < hi; ++i) s = __fadd_rn(s, det_exp(__fsub_rn(zof(i), zmax)));
      S.csum[c] = s;
    }
    __syncthreads();
    if (tid == 0) { float s = 0.f; for (int c = 0; c < nchunk; ++c) s = __fadd_rn(s, S.csum[c]); S.total = s; }
    __syncthreads();
    have_csum = true;
    const float tot = S.total;
    for (int i = tid; i < V; i += SAMP_THREADS) {
      const float z = zof(i);
      const float pr = __fdiv_rn(det_exp(__fsub_rn(z, zmax)), tot);
      if (!(pr < min_p)) {
        unsigned slot = atomicAdd(&S.n_cand, 1u);
        if (slot < CAND_CAP) S.cand[slot] = ((unsigned long long)key_of(z) << 32) | (0xffffffffu - (unsigned)i);
        else S.overflow = 1;
      }
    }
    __syncthreads();
    if (S.n_cand > 0) { filtered = true; minp_mode = true; top_k = 0; top_p = 1.0f; }
  }

  // (4) top-k, ties kept (models/utils.py:82-86).
  // Fast path: a lower bound LB on the k-th largest key is the k-th largest of 32 warp maxima (k <= 32) or
  // of the 1024 thread maxima; every element whose top-16 key bits reach LB's is a candidate (a superset
  // of the top-k and of all ties at the k-th value).  Candidates are re-read exactly, sorted (value desc,
  // index asc) and cut at the k-th value.  Falls back to the exact radix select if the superset overflows.
  bool fast_done = false;
  bool large = !filtered && ((top_k <= 0 && top_p < 1.0f) || top_k > 1024);   // needs the full sort
  if (filtered && S.overflow) { large = true; filtered = false; }               // min-p kept more than CAND_CAP
  if (!large && !filtered && top_k > 0 && use_hi) {
    const int k = min(max(top_k, 1), V);
    if (k <= 32) {
      if (warp == 0) {
        unsigned x = S.redk[lane];
        // bitonic sort of 32 keys across the warp, descending
        for (int size = 2; size <= 32; size <<= 1)
          for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const unsigned y = __shfl_xor_sync(0xffffffffu, x, stride);
            const bool up = ((lane & size) == 0);              // this block sorts descending
            const bool lower = ((lane & stride) == 0);
            const unsigned mx = max(x, y), mn = min(x, y);
            x = (up == lower) ? mx : mn;
          }
        if (lane == k - 1) S.lb = x;
      }
    } else {
      for (int size = 2; size <= SAMP_THREADS; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          __syncthreads();
          const int j = tid ^ stride;
          if (j > tid) {
            const bool desc = ((tid & size) == 0);
            const unsigned x = S.tmax[tid], y = S.tmax[j];
            if ((x < y) == desc) { S.tmax[tid] = y; S.tmax[j] = x; }
          }
        }
      __syncthreads();
      if (tid == 0) S.lb = S.tmax[k - 1];
    }
    __syncthreads();
    SP_PROBE(5);
    unsigned thr16 = S.lb >> 16;
    // raw-logit keys: values up to 2 keys below the bound can still round to the same z = v / T (3 ulps cannot), and ties
    // at the k-th value must survive
    if (vspace && scaled) thr16 = (S.lb - min(S.lb, 2u)) >> 16;
    for (int g = tid; g * 8 < V; g += SAMP_THREADS) {        // 8 keys per 16-byte shared-memory load
      const uint4 h = *reinterpret_cast<const uint4*>(&S.hi16[g * 8]);
      const unsigned w[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const unsigned k16 = (c & 1) ? (w[c >> 1] >> 16) : (w[c >> 1] & 0xffffu);
        const int i = g * 8 + c;
        if (k16 >= thr16 && i < V) {
          const unsigned slot = atomicAdd(&S.n_cand, 1u);
          if (slot < CAND_CAP) S.cand[slot] = ((unsigned long long)key_of(zof(i)) << 32) | (0xffffffffu - (unsigned)i);
          else S.overflow = 1;
        }
      }
    }
    __syncthreads();
    SP_PROBE(6);
    if (!S.overflow) {
      const int n = (int)S.n_cand;
      if (a.probe && row == 0 && tid == 0) a.probe[11] = (unsigned long long)n;
      int np2 = 1; while (np2 < n) np2 <<= 1;
      for (int i = n + tid; i < np2; i += SAMP_THREADS) S.cand[i] = 0ull;
      __syncthreads();
      if (np2 <= 1024) {
        // a few dozen candidates: one compare-exchange per thread per bitonic stage, and the barrier spans only the
        // warps that hold pairs (a 1024-thread __syncthreads per stage cost 0.25 us x 28 stages)
        const int npairs = np2 >> 1;
        const int nthr = max(32, (npairs + 31) & ~31);
        if (tid < nthr) {
          for (int size = 2; size <= np2; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
              if (tid < npairs) {
                const int i = ((tid & ~(stride - 1)) << 1) | (tid & (stride - 1)), j = i | stride;
                const bool desc = ((i & size) == 0);
                const unsigned long long x = S.cand[i], y = S.cand[j];
                if ((x < y) == desc) { S.cand[i] = y; S.cand[j] = x; }
              }
              asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
            }
        }
        __syncthreads();
      } else {
        for (int size = 2; size <= np2; size <<= 1)
          for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < np2; i += SAMP_THREADS) {
              const int j = i ^ stride;
              if (j > i) {
                const bool desc = ((i & size) == 0);
                unsigned long long x = S.cand[i], y = S.cand[j];
                if ((x < y) == desc) { S.cand[i] = y; S.cand[j] = x; }
              }
            }
            __syncthreads();
          }
      }
      const unsigned thr_key = (unsigned)(S.cand[k - 1] >> 32);
      for (int i = tid; i < n; i += SAMP_THREADS) {
        const bool in = (unsigned)(S.cand[i] >> 32) >= thr_key;
        const bool next_in = (i + 1 < n) && ((unsigned)(S.cand[i + 1] >> 32) >= thr_key);
        if (in && !next_in) S.n_surv = i + 1;
      }
      __syncthreads();
      fast_done = true;
      filtered = true;
    } else {
      __syncthreads();
      if (tid == 0) { S.n_cand = 0; S.overflow = 0; }
      __syncthreads();
    }
  }
  if (!large && !filtered && top_k > 0) {
    int k = min(max(top_k, 1), V);
    unsigned prefix = 0, mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int i = tid; i < 256; i += SAMP_THREADS) S.hist[i] = 0;
      __syncthreads();
      for (int i0 = 0; i0 < V; i0 += SAMP_THREADS) {
        const int i = i0 + tid;
        const bool ok = i < V;
        unsigned key = ok ? key_of(zof(i)) : 0u;
        const bool match = ok && ((key & mask) == prefix);
        const unsigned digit = (key >> shift) & 0xffu;
        // warp-aggregated histogram update
        const unsigned act = __ballot_sync(0xffffffffu, match);
        if (match) {
          const unsigned peers = __match_any_sync(act, digit);
          if ((int)(__ffs(peers) - 1) == lane) atomicAdd(&S.hist[digit], (unsigned)__popc(peers));
        }
      }
      __syncthreads();
      if (tid == 0) {
        int rem = k; int d = 255;
        for (; d > 0; --d) { if ((int)S.hist[d] >= rem) break; rem -= (int)S.hist[d]; }
        S.prefix = prefix | ((unsigned)d << shift); S.kth = rem;
      }
      __syncthreads();
      prefix = S.prefix; k = S.kth; mask |= (0xffu << shift);
      __syncthreads();
    }
    const unsigned thr_key = prefix;
    for (int i = tid; i < V; i += SAMP_THREADS) {
      const unsigned key = key_of(zof(i));
      if (key >= thr_key) {
        unsigned slot = atomicAdd(&S.n_cand, 1u);
        if (slot < CAND_CAP) S.cand[slot] = ((unsigned long long)key << 32) | (0xffffffffu - (unsigned)i);
        else S.overflow = 1;
      }
    }
    __syncthreads();
    if (S.overflow) large = true; else filtered = true;       // > CAND_CAP ties at the k-th value: full sort
  }

  SP_PROBE(7);
  int token = 0;
  if (large) {
    if (!a.scratch_u64) { if (tid == 0) sl.error |= 2; token = S.amax; }
    else token = sample_large(a, S, row, tid, lg, V, T, scaled, zmax, top_k, top_p, min_p, minp_mode, sl, n_gen);
  } else if (filtered) {
    // sort survivors (value desc, index asc): bitonic on the 64-bit composite, descending
    int n;
    if (fast_done) {
      n = S.n_surv;
    } else {
      n = min((int)S.n_cand, CAND_CAP);
      int np2 = 1; while (np2 < n) np2 <<= 1;
      for (int i = n + tid; i < np2; i += SAMP_THREADS) S.cand[i] = 0ull;
      __syncthreads();
      for (int size = 2; size <= np2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          for (int i = tid; i < np2; i += SAMP_THREADS) {
            const int j = i ^ stride;
            if (j > i) {
              const bool desc = ((i & size) == 0);
              unsigned long long x = S.cand[i], y = S.cand[j];
              if ((x < y) == desc) { S.cand[i] = y; S.cand[j] = x; }
            }
          }
          __syncthreads();
        }
      }
    }
    const float z0 = val_of((unsigned)(S.cand[0] >> 32));
    for (int i = tid; i < n; i += SAMP_THREADS)
      S.e[i] = det_exp(__fsub_rn(val_of((unsigned)(S.cand[i] >> 32)), z0));
    __syncthreads();
    if (tid == 0) {
      int n_keep = n;
      if (top_p < 1.0f) {          // (5) nucleus on the top-k-filtered distribution, shift-right rule
        float s = 0.f;
        for (int i = 0; i < n; ++i) s = __fadd_rn(s, S.e[i]);
        float cum = 0.f;
        n_keep = 1;
        for (int i = 0; i + 1 < n; ++i) {
          cum = __fadd_rn(cum, __fdiv_rn(S.e[i], s));
          if (cum > top_p) break;   // token i+1 removed (cum through i exceeds p); cum is monotone
          n_keep = i + 2;
        }
      }
      // (6) inverse-CDF draw over the kept prefix
      const float u = u_draw;
      float s2 = 0.f;
      for (int i = 0; i < n_keep; ++i) s2 = __fadd_rn(s2, S.e[i]);
      float c = 0.f; int pick = n_keep - 1;
      for (int i = 0; i < n_keep; ++i) {
        c = __fadd_rn(c, __fdiv_rn(S.e[i], s2));
        if (u < c) { pick = i; break; }
      }
      S.n_keep = (int)(0xffffffffu - (unsigned)(S.cand[pick] & 0xffffffffull));
    }
    __syncthreads();
    token = S.n_keep;
  } else {
    // unfiltered: whole vocabulary in index order, chunked sums, target t = u*s
    if (!have_csum) {
      for (int c = tid; c < nchunk; c += SAMP_THREADS) {
        float s = 0.f;
        const int hi = min(V, (c + 1) * CHUNK);
        for (int i = c * CHUNK; i < hi; ++i) s = __fadd_rn(s, det_exp(__fsub_rn(zof(i), zmax)));
        S.csum[c] = s;
      }
      __syncthreads();
    }
    if (tid == 0) {
      float s = 0.f;
      for (int c = 0; c < nchunk; ++c) s = __fadd_rn(s, S.csum[c]);
      const float u = u_draw;
      const float t = __fmul_rn(u, s);
      float cc = 0.f, base = 0.f; int cidx = -1;
      for (int c = 0; c < nchunk; ++c) { base = cc; cc = __fadd_rn(cc, S.csum[c]); if (t < cc) { cidx = c; break; } }
      int pick = V - 1;
      if (cidx >= 0) {
        const int hi = min(V, (cidx + 1) * CHUNK);
        float run = base; pick = hi - 1;
        for (int i = cidx * CHUNK; i < hi; ++i) {
          run = __fadd_rn(run, det_exp(__fsub_rn(zof(i), zmax)));
          if (t < run) { pick = i; break; }
        }
      }
      S.n_keep = pick;
    }
    __syncthreads();
    token = S.n_keep;
  }

  SP_PROBE(8);
  // (7) stop rules + state update (models/t5gemma.py:1020-1048, 1075-1099)
  if (tid == 0) {
    const int amax = S.amax;
    if (a.picks_out && !a.flat_tokens) a.picks_out[(size_t)row * a.tokens_stride + n_gen] = token;
    if (a.forced_pool && n_gen < sl.n_forced) token = a.forced_pool[(size_t)row * a.tokens_stride + n_gen];
    bool force = (token == eos) || (amax == eos);
    if (a.text_guard > 0) force = force || (eff_len > max(1, sl.n_text) * a.text_guard);
    const bool budget = n_gen > sl.budget_limit;
    if (force || budget) token = eos;
    if (sl.max_new_tokens > 0 && n_gen + 1 >= sl.max_new_tokens) token = eos;
    if (a.tokens_out) a.tokens_out[(size_t)row * a.tokens_stride + (a.flat_tokens ? 0 : n_gen)] = token;
    if (a.argmax_out) a.argmax_out[row] = amax;
    if (S.overflow) sl.error |= 1;
    {   // models/t5gemma.py:1050-1054
      bool in_set = false;
      for (int i = 0; i < sl.n_silence; ++i) in_set |= (a.topk_sched_pool[sl.silence_off + i] == token);
      sl.consec_silence = (in_set && token == sl.prev_token) ? sl.consec_silence + 1 : 0;
      sl.prev_token = token;
    }
    sl.n_generated = n_gen + 1;
    const int new_len = cur_len + 1;
    sl.cur_len = new_len;
    sl.last_token = token;
    if (token == eos) { sl.active = 0; sl.finished = 1; }
    else {
      // python float64 arithmetic, then fp32 (models/t5gemma.py:1087-1094)
      double p = (double)(new_len - 1) / (double)max(1, sl.est_total - 1) * (double)a.progress_scale;
      p = fmin(p, (double)a.progress_scale);
      sl.pos = (float)p;
    }
    if (a.host_mirror) {
      volatile int* hm = a.host_mirror + row * 8;
      hm[0] = sl.active; hm[1] = sl.finished; hm[2] = n_gen + 1; hm[3] = new_len; hm[4] = sl.error;
    }
    S.total = sl.pos;
  }
  SP_PROBE(9);
  __syncthreads();
  if (tid < SLOT_WORDS) reinterpret_cast<int*>(&a.slots[row])[tid] = reinterpret_cast<const int*>(&sl_s)[tid];
  if (a.rope_out) {
    // cos/sin of the new token's PM-RoPE angle, once per step instead of once per attention kernel
    const float pos = S.total;
    const int half = a.head_dim / 2;
    for (int i = tid; i < half; i += SAMP_THREADS) {
      float sn, cs;
      sincosf(pos * a.inv_freq[i], &sn, &cs);
      a.rope_out[(size_t)row * a.head_dim + i] = cs;
      a.rope_out[(size_t)row * a.head_dim + half + i] = sn;
    }
  }
  SP_PROBE(10);
  trace_end(a.trace);
#undef SP_PROBE
}

}  // namespace

cudaError_t launch_sampler(const SamplerArgs& a, cudaStream_t st, bool pdl) {
  if (a.V > MAX_CHUNKS * CHUNK) return cudaErrorInvalidValue;
  static PerDeviceFlag attr_set;
  if (!attr_set.here()) {
    cudaError_t e = cudaFuncSetAttribute(sampler_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Shared));
    if (e != cudaSuccess) return e;
    attr_set.here() = 1;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(a.rows);
  cfg.blockDim = dim3(SAMP_THREADS);
  cfg.dynamicSmemBytes = sizeof(Shared);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, sampler_kernel, a);
}
