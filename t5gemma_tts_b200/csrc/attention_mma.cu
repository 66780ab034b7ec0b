// Decode attention for BATCHED rows (B > 4): tile kernel over the paged KV pool.
//
// The CUDA-core kernel of attention.cu spends ~360 instructions per two cached tokens (per-token shuffles, exps and
// accumulator rescaling), which at 64 rows makes the step's attention instruction-latency-bound (ncu: 7 k
// instructions per warp, 1-2 warps per scheduler, < 1 % of the time waiting on memory).  Here the CTA of a
// (request, kv head, split) walks its key range in tiles of 32 tokens:
//   * K/V rows travel with cp.async into a 2-stage shared-memory ring (first tiles issued BEFORE
//     griddepcontrol.wait; rows padded by 16 bytes so ldmatrix is conflict-free); the page lookups of the whole range
//     are resolved once before the loop;
//   * scores  S^T[token, head] = K_tile . Q^T  on mma.sync m16n8k16 (the G query heads of the group sit in the
//     n = 8 columns), scale / softcap / mask in the accumulator registers;
//   * one online-softmax update per tile and head (warp h owns head h), probabilities rounded to bf16 like the
//     reference's `attn_weights.to(dtype)` (HF:modeling_t5gemma.py:235);
//   * O^T[dim, head] += V_tile^T . P^T  on mma.sync, the D dims dealt over the 4 warps.
// ~120 instructions per warp per 32 tokens.  tcgen05 is not used: the useful M is G = 2 rows and the work is
// bandwidth / latency bound.  Same fused glue as attention.cu: PM-RoPE of q and of the new k, in-place KV
// append.
// Chunked mode (the batched step's default): blockIdx.y = chunk.  A row's key range is cut into ceil(range / chunk_tokens)
// equal pieces, so a step's attention is dealt in pieces of similar size whatever the spread of the rows' contexts (one
// CTA per row made the kernel as long as its longest row: 53.6 -> 32.6 us per layer at contexts U[0,900), +27 % on a
// ragged job); the chunks of a (row, kv head) park (m, l, O) in global scratch and the last one to arrive merges them.
// Measured and rejected for the tile loads: one producer warp issuing the 16-byte cp.async (63.6 us: a single warp's
// instruction stream is too slow) and one cp.async.bulk per 512-byte row (38.9 us: 64 small bulk copies per tile).
#include "kernels.h"

namespace {

constexpr int AM_NT = 256, AM_WARPS = 8;  // the kernel is instruction-issue bound: two warps per scheduler
constexpr int AM_TT = 32;                 // tokens per tile
constexpr int AM_NST = 2;                 // ring stages
constexpr int AM_BT_CACHE = 256;
constexpr int AM_ROWS_PRE = 512;          // key ranges up to this length resolve their page lookups once, before the tile loop
constexpr int AM_MAX_CHUNKS = 16;         // chunked mode: chunks per (row, kv head)

__device__ __forceinline__ void am_cp_async16(void* dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;          // src-size 0: zero fill (rows past the key range must not hold NaN bit patterns)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];"
               : "=r"(r[0]), "=r"(r[1]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int D> struct AmGeo {
  static constexpr int LD = D + 8;                         // padded row (elements): 16-byte skew per row
  static constexpr int KS = D / 16;                        // k-steps of the score MMA
  static constexpr int KSH = (KS + 1) / 2;                 // ... per k-half (two warps share a 16-token group)
  static constexpr int NMT = D / 16;                       // 16-dim m-tiles of the output MMA
  static constexpr int MTW = (NMT + AM_WARPS - 1) / AM_WARPS;   // m-tiles per warp
  static constexpr size_t stage_elems = (size_t)2 * AM_TT * LD; // K tile + V tile
  // q / probability operand tiles hold G real rows + one shared zero row (the MMA's n = 8 columns beyond G read it)
  static constexpr size_t ring_bytes(int G) { return (AM_NST * stage_elems + (size_t)(G + 1) * LD + (size_t)(G + 1) * (AM_TT + 8)) * sizeof(bf16); }
  static constexpr size_t dyn_bytes(int G) { return ring_bytes(G); }
};

// grid (Hkv, chunks, B), 256 threads
template <int G, int D>
__global__ void __launch_bounds__(AM_NT) attn_decode_mma_kernel(AttnDecodeArgs a) {
  using Geo = AmGeo<D>;
  constexpr int LD = Geo::LD, KS = Geo::KS, KSH = Geo::KSH, NMT = Geo::NMT, MTW = Geo::MTW;
  static_assert(AM_TT == 32, "one token per lane in the softmax phase, two 16-token groups");
  extern __shared__ __align__(16) unsigned char am_dyn[];
  bf16* ring = reinterpret_cast<bf16*>(am_dyn);                         // [NST][2][TT][LD]
  bf16* qb = ring + AM_NST * Geo::stage_elems;                          // [G+1][LD]   rotated q, row G is zero
  bf16* pb = qb + (G + 1) * LD;                                         // [G+1][TT+8] probabilities of the tile, row G is zero
  __shared__ float cs[D / 2], sn[D / 2];
  __shared__ float knew[D], vnew[D];
  __shared__ int bt_s[AM_BT_CACHE];
  __shared__ float scp[2][G][AM_TT];                                    // partial q.k of the two k-halves
  __shared__ float corr_s[8];
  __shared__ unsigned rowoff[AM_NST][AM_TT];                            // row offsets (elements) inside the layer's K plane
  __shared__ unsigned rowoff_all[AM_ROWS_PRE];                          // ... of the whole key range when it is short enough
  __shared__ float ml_s[G][2];
  __shared__ float cw_s[AM_MAX_CHUNKS][G], cl_s[AM_MAX_CHUNKS][G];      // chunked mode: merge weights / sums of the chunks
  __shared__ int last_s;

  pdl_launch_dependents();
  // rows in descending key count (host-maintained, written before the launch): CTAs are scheduled in blockIdx order, so
  // the long rows start first -- beside the producer GEMM's CTAs -- instead of defining the tail of the kernel
  const int hk = blockIdx.x, split = blockIdx.y, b = a.row_order ? a.row_order[blockIdx.z] : (int)blockIdx.z;
  unsigned long long* probe = (a.probe && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && b < 64) ? a.probe + b * 11 : nullptr;
#define AM_PROBE(k) do { if (probe) probe[k] = globaltimer_ns(); } while (0)
  AM_PROBE(0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int PT = a.pool.page_tokens;
  const int* bt = a.block_table + (size_t)b * a.bt_stride;
  // ---- before the dependency resolves: state of earlier steps only (see attention.cu / engine.cu) ----
  const SlotDev& sl = a.slots[b];
  const int active = sl.active;
  const int L = a.is_cross ? sl.n_text : sl.cur_len;
  const float pos = sl.pos;
  {
    // loads first, shared-memory stores after: one round trip for the slot, the RoPE table and the block table
    float rc = 1.f, rs = 0.f;
    const bool rope_tab = a.rope_cs != nullptr && tid < D / 2;
    if (rope_tab) { rc = a.rope_cs[(size_t)b * D + tid]; rs = a.rope_cs[(size_t)b * D + D / 2 + tid]; }
    const int btv = (tid < a.bt_stride) ? bt[tid] : 0;                  // AM_BT_CACHE == AM_NT entries
    if (rope_tab) { cs[tid] = rc; sn[tid] = rs; }
    if (!a.rope_cs) for (int i = tid; i < D / 2; i += AM_NT) { float s_, c_; sincosf(pos * a.inv_freq[i], &s_, &c_); cs[i] = c_; sn[i] = s_; }
    bt_s[tid] = btv;
  }
  static_assert(AM_BT_CACHE == AM_NT && D / 2 <= AM_NT, "prologue mapping");
  for (int i = tid; i < (G + 1) * LD; i += AM_NT) qb[i] = __float2bfloat16(0.f);
  for (int i = tid; i < (G + 1) * (AM_TT + 8); i += AM_NT) pb[i] = __float2bfloat16(0.f);
  const int lo = (!a.is_cross && a.window > 0) ? max(0, L - a.window) : 0;
  int chunk = (L - lo + 15) / 16 * 16;                       // chunking off: the whole range
  int n_chunks = 1;
  if (a.chunk_tokens > 0) {                                 // chunked mode: equal pieces of at most chunk_tokens keys
    n_chunks = min(max((L - lo + a.chunk_tokens - 1) / a.chunk_tokens, 1), a.max_chunks);
    if (!active || split >= n_chunks) return;               // nothing to do for this CTA (exited CTAs release the dependents)
    chunk = ((L - lo + n_chunks - 1) / n_chunks + AM_TT - 1) / AM_TT * AM_TT;
  }
  const int t_begin = lo + split * chunk, t_end = min(L, t_begin + chunk);
  const bool has_new = (!a.is_cross) && (t_end == L) && (t_end > t_begin);
  const int n_tiles = (t_end > t_begin) ? (t_end - t_begin + AM_TT - 1) / AM_TT : 0;
  __syncthreads();
  auto page_of = [&](int t) -> int { const int pi = t / PT; return pi < AM_BT_CACHE ? bt_s[pi] : bt[pi]; };
  auto stage_ptr = [&](int stage, int kv) -> bf16* { return ring + (size_t)stage * Geo::stage_elems + (size_t)kv * AM_TT * LD; };
  // tile ti = tokens [t_begin + ti*TT, +TT) -> ring stage ti % NST.  One thread per token resolves the page (an integer
  // division and a block-table lookup) into a row offset; the 16-byte copies then cost a handful of instructions each
  // (the naive per-chunk address computation made the ISSUE loop the critical path: 3.7 us per 32-token tile).
  const bf16* kbase = a.pool.ptr(a.layer, 0, 0);
  const size_t kv_stride = (size_t)a.pool.n_pages * a.pool.page_elems();
  // row offsets of the whole key range, resolved once (a chunk is at most AM_ROWS_PRE keys unless chunking is off)
  const bool ro_pre = n_tiles * AM_TT <= AM_ROWS_PRE;
  if (ro_pre) {
    for (int i = tid; i < n_tiles * AM_TT; i += AM_NT) {
      const int t = t_begin + i;
      const bool valid = t < t_end && !(has_new && t == L - 1);
      unsigned ro = 0xFFFFFFFFu;
      if (valid) { const int page = page_of(t), off = t % PT; ro = (unsigned)((size_t)page * a.pool.page_elems() + ((size_t)hk * PT + off) * D); }
      rowoff_all[i] = ro;
    }
    __syncthreads();
  }
  auto issue_tile = [&](int ti) {
    if (ti < n_tiles) {                                       // CTA-uniform
      const int stage = ti % AM_NST;
      const unsigned* ro_t = rowoff_all + ti * AM_TT;
      if (!ro_pre) {
        if (tid < AM_TT) {
          const int t = t_begin + ti * AM_TT + tid;
          const bool valid = t < t_end && !(has_new && t == L - 1);
          unsigned ro = 0xFFFFFFFFu;
          if (valid) { const int page = page_of(t), off = t % PT; ro = (unsigned)((size_t)page * a.pool.page_elems() + ((size_t)hk * PT + off) * D); }
          rowoff[stage][tid] = ro;
        }
        __syncthreads();
        ro_t = rowoff[stage];
      }
      constexpr int CPR = D / 8;                              // 16-byte chunks per row
      bf16* kst = stage_ptr(stage, 0);
#pragma unroll 4
      for (int c = tid; c < AM_TT * CPR * 2; c += AM_NT) {
        const int tok = c / (2 * CPR), rem = c - tok * (2 * CPR), kv = rem / CPR, col = rem - kv * CPR;
        const unsigned ro = ro_t[tok];
        const bool valid = ro != 0xFFFFFFFFu;                 // rows past the range / the new token: zero fill
        am_cp_async16(kst + (size_t)kv * AM_TT * LD + (size_t)tok * LD + col * 8,
                      kbase + (kv ? kv_stride : 0) + (valid ? ro : 0u) + col * 8, valid);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // one group per tile, empty ones included
  };
  if (active) {                                               // in flight while the producer kernel is still running
#pragma unroll
    for (int s_ = 0; s_ < AM_NST; ++s_) issue_tile(s_);
  }

  AM_PROBE(1);
  pdl_wait();
  trace_begin(a.trace);
  AM_PROBE(2);
  if (!active) return;                               // CTA-uniform
  // ---- the producer's outputs: raw q and the new k/v.  A thread loads both elements of a rotation pair (j, j + D/2),
  //      rotates in registers (PM-RoPE at the row's progress position) and stores the MMA operand directly. ----
  {
    constexpr int QP = (G * D / 2 + AM_NT - 1) / AM_NT;
    float x1[QP], x2[QP];
    float k1 = 0.f, k2 = 0.f, vv = 0.f;
    const float* qp = a.q + (size_t)b * a.q_stride + (size_t)(hk * G) * D;
#pragma unroll
    for (int u = 0; u < QP; ++u) {
      const int i = tid + u * AM_NT, g = i / (D / 2), j = i - g * (D / 2);
      const bool ok = i < G * D / 2;
      x1[u] = ok ? __ldcg(qp + g * D + j) : 0.f;
      x2[u] = ok ? __ldcg(qp + g * D + j + D / 2) : 0.f;
    }
    if (has_new) {
      const float* kp = a.kv_new + (size_t)b * a.kv_stride + (size_t)hk * D;
      const float* vp = a.kv_new + (size_t)b * a.kv_stride + (size_t)(a.Hkv + hk) * D;
      if (tid < D / 2) { k1 = __ldcg(kp + tid); k2 = __ldcg(kp + tid + D / 2); }
      if (tid < D) vv = __ldcg(vp + tid);
    }
#pragma unroll
    for (int u = 0; u < QP; ++u) {
      const int i = tid + u * AM_NT, g = i / (D / 2), j = i - g * (D / 2);
      if (i < G * D / 2) {
        qb[g * LD + j] = __float2bfloat16(x1[u] * cs[j] - x2[u] * sn[j]);
        qb[g * LD + j + D / 2] = __float2bfloat16(x2[u] * cs[j] + x1[u] * sn[j]);
      }
    }
    if (has_new) {
      if (tid < D / 2) { knew[tid] = k1 * cs[tid] - k2 * sn[tid]; knew[tid + D / 2] = k2 * cs[tid] + k1 * sn[tid]; }
      if (tid < D) vnew[tid] = vv;
    }
  }
  static_assert(D <= AM_NT, "one thread per element of the new k/v row");
  __syncthreads();
  AM_PROBE(3);
  if (has_new) {   // append to the page (K post-RoPE), visible to later steps
    const int t = L - 1, page = page_of(t), off = t % PT;
    bf16* kd = a.pool.ptr(a.layer, 0, page) + ((size_t)hk * PT + off) * D;
    bf16* vd = a.pool.ptr(a.layer, 1, page) + ((size_t)hk * PT + off) * D;
    if (tid < D) { kd[tid] = __float2bfloat16(knew[tid]); vd[tid] = __float2bfloat16(vnew[tid]); }
  }
  // q fragments (B operand of the score MMA): B[k = dim][n = head] from qb[head][dim]; this warp's k-half only
  const int tg = warp & 1, kh = (warp >> 1) & 1;      // score phase (warps 0-3): 16-token group, k-half
  uint32_t qf[KSH][2];
#pragma unroll
  for (int i = 0; i < KSH; ++i) {
    const int ks = kh * KSH + i;
    if (ks < KS) ldsm_x2(qf[i], qb + (size_t)min(lane & 7, G) * LD + ks * 16 + ((lane >> 3) & 1) * 8);
    else { qf[i][0] = 0u; qf[i][1] = 0u; }
  }
  AM_PROBE(4);

  const float inv_cap = a.softcap > 0.f ? 1.f / a.softcap : 0.f;
  float m_run = -INFINITY, l_run = 0.f;              // warp h < G: running max / sum of head h (lane-replicated)
  float acc[MTW][4];
#pragma unroll
  for (int i = 0; i < MTW; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
  const int g8 = lane >> 2, t4 = lane & 3;

  for (int ti = 0; ti < n_tiles; ++ti) {
    asm volatile("cp.async.wait_group %0;" ::"n"(AM_NST - 1) : "memory");   // tile ti has landed (groups retire in order)
    __syncthreads();
    if (ti == 0) AM_PROBE(5);
    if (ti == 1) AM_PROBE(6);
    if (ti == 2) AM_PROBE(7);
    bf16* kb = stage_ptr(ti % AM_NST, 0);
    bf16* vb = stage_ptr(ti % AM_NST, 1);
    const int tile_t0 = t_begin + ti * AM_TT;
    if (has_new && L - 1 >= tile_t0 && L - 1 < tile_t0 + AM_TT) {       // CTA-uniform: the new token's row comes from registers
      const int tok = L - 1 - tile_t0;
      if (tid < D) { kb[(size_t)tok * LD + tid] = __float2bfloat16(knew[tid]); vb[(size_t)tok * LD + tid] = __float2bfloat16(vnew[tid]); }
      __syncthreads();
    }
    // ---- partial scores: warps 0-3 = (16-token group tg) x (k-half kh); rows past the range are zero-filled ----
    if (warp < 4) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      const bf16* arow = kb + (size_t)(tg * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + (lane >> 4) * 8 + kh * KSH * 16;
#pragma unroll
      for (int i = 0; i < KSH; ++i) {
        if (kh * KSH + i < KS) {
          uint32_t af[4];
          ldsm_x4(af, arow + i * 16);
          mma_bf16_16816(c, af, qf[i]);
        }
      }
      // c0,c1: token g8, heads 2*t4, 2*t4+1 ; c2,c3: token g8+8
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int head = 2 * t4 + (r & 1), tok = tg * 16 + g8 + (r >> 1) * 8;
        if (head < G) scp[kh][head][tok] = c[r];
      }
    }
    __syncthreads();
    // ---- scale / softcap / mask + online softmax: warp h owns head h, lane = token of the tile ----
    if (warp < G) {
      float s = (scp[0][warp][lane] + scp[1][warp][lane]) * a.scale;
      if (a.softcap > 0.f) {
        const float e2 = __expf(2.f * s * inv_cap);
        s = a.softcap * (1.f - __fdividef(2.f, e2 + 1.f));
      }
      if (tile_t0 + lane >= t_end) s = -INFINITY;
      float tm = s;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, o));
      const float mn = fmaxf(m_run, tm);
      const float corr = (m_run == -INFINITY) ? 0.f : __expf(m_run - mn);
      const bf16 pq = __float2bfloat16((s == -INFINITY) ? 0.f : __expf(s - mn));
      pb[warp * (AM_TT + 8) + lane] = pq;
      const float ps = warp_sum(__bfloat162float(pq));         // the sum uses the rounded weights the MMA will see
      l_run = l_run * corr + ps;
      m_run = mn;
      if (lane == 0) corr_s[warp] = corr;
    }
    __syncthreads();
    // ---- O^T[dim, head] = corr * O^T + V^T P^T : m-tiles (16 dims) dealt over the warps ----
    {
      const float c0 = (2 * t4 < G) ? corr_s[2 * t4] : 0.f, c1 = (2 * t4 + 1 < G) ? corr_s[2 * t4 + 1] : 0.f;
      uint32_t pf[AM_TT / 16][2];
#pragma unroll
      for (int kk = 0; kk < AM_TT / 16; ++kk) ldsm_x2(pf[kk], pb + (size_t)min(lane & 7, G) * (AM_TT + 8) + kk * 16 + ((lane >> 3) & 1) * 8);
#pragma unroll
      for (int i = 0; i < MTW; ++i) {
        const int mt = warp + i * AM_WARPS;
        if (mt < NMT) {
          acc[i][0] *= c0; acc[i][1] *= c1; acc[i][2] *= c0; acc[i][3] *= c1;
#pragma unroll
          for (int kk = 0; kk < AM_TT / 16; ++kk) {
            uint32_t vf[4];
            ldsm_x4_t(vf, vb + (size_t)(kk * 16 + (lane & 7) + (lane >> 4) * 8) * LD + mt * 16 + ((lane >> 3) & 1) * 8);
            mma_bf16_16816(acc[i], vf, pf[kk]);
          }
        }
      }
    }
    __syncthreads();                                       // stage, scp and pb are free again
    issue_tile(ti + AM_NST);
  }
  AM_PROBE(8);

  if (warp < G && lane == 0) { ml_s[warp][0] = m_run; ml_s[warp][1] = l_run; }
  if (n_chunks > 1) {
    // ---- chunked mode, several chunks in this row's range: park the unnormalised partial, last arriver merges ----
    const size_t slot = ((size_t)b * a.Hkv + hk) * AM_MAX_CHUNKS;
    float* po = a.part_o + (slot + split) * G * D;
#pragma unroll
    for (int i = 0; i < MTW; ++i) {
      const int mt = warp + i * AM_WARPS;
      if (mt < NMT) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int head = 2 * t4 + (r & 1), dim = mt * 16 + g8 + (r >> 1) * 8;
          if (head < G) __stcg(po + head * D + dim, acc[i][r]);
        }
      }
    }
    if (warp < G && lane == 0) { __stcg(a.part_ml + ((slot + split) * G + warp) * 2, m_run); __stcg(a.part_ml + ((slot + split) * G + warp) * 2 + 1, l_run); }
    __syncthreads();
    // one acq_rel atomic by thread 0 orders the CTA's partial (observed through the barrier above) before the count and
    // the other chunks' partials before this CTA's reads below -- no membar.gl by all 256 threads
    if (tid == 0) {
      unsigned old;
      asm volatile("atom.add.acq_rel.gpu.u32 %0, [%1], 1;" : "=r"(old) : "l"(a.part_cnt + (size_t)b * a.Hkv + hk) : "memory");
      last_s = (old == (unsigned)(n_chunks - 1));
    }
    __syncthreads();
    if (!last_s) { trace_end(a.trace); return; }
    if (tid == 0) a.part_cnt[(size_t)b * a.Hkv + hk] = 0;   // every chunk has arrived: ready for the next launch
    // partial outputs first (independent loads, in flight while the weights are computed)
    constexpr int OPT = (G * D + AM_NT - 1) / AM_NT;
    float po_r[OPT][AM_MAX_CHUNKS];
#pragma unroll
    for (int u = 0; u < OPT; ++u)
#pragma unroll
      for (int c = 0; c < AM_MAX_CHUNKS; ++c) {
        const int i = tid + u * AM_NT;
        po_r[u][c] = (c < n_chunks && i < G * D) ? __ldcg(a.part_o + (slot + c) * G * D + i) : 0.f;
      }
    if (tid < G * AM_MAX_CHUNKS) {                            // (chunk, head) -> m and l, one load each
      const int c = tid / G, g = tid % G;
      const bool ok = c < n_chunks;
      cw_s[c][g] = ok ? __ldcg(a.part_ml + ((slot + c) * G + g) * 2) : -INFINITY;
      cl_s[c][g] = ok ? __ldcg(a.part_ml + ((slot + c) * G + g) * 2 + 1) : 0.f;
    }
    __syncthreads();
    if (tid < G) {
      float M = -INFINITY;
      for (int c = 0; c < n_chunks; ++c) M = fmaxf(M, cw_s[c][tid]);
      float den = 0.f;
      for (int c = 0; c < n_chunks; ++c) {
        const float m = cw_s[c][tid];
        const float wt = (m == -INFINITY) ? 0.f : __expf(m - M);
        den = fmaf(wt, cl_s[c][tid], den);
        cw_s[c][tid] = wt;
      }
      const float inv = den > 0.f ? 1.f / den : 0.f;
      for (int c = 0; c < n_chunks; ++c) cw_s[c][tid] *= inv;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < OPT; ++u) {
      const int i = tid + u * AM_NT;
      if (i >= G * D) break;
      const int g = i / D;
      float o = 0.f;
#pragma unroll
      for (int c = 0; c < AM_MAX_CHUNKS; ++c) if (c < n_chunks) o = fmaf(cw_s[c][g], po_r[u][c], o);
      const size_t idx = (size_t)b * a.Hq * D + (size_t)(hk * G) * D + i;
      if (a.out) a.out[idx] = o;
      if (a.out_bf) a.out_bf[idx] = __float2bfloat16(o);
    }
    AM_PROBE(10);
    trace_end(a.trace);
    return;
  }
  // ---- no split: normalise and store straight from the accumulator registers ----
  __syncthreads();
  const float i0 = (2 * t4 < G && ml_s[2 * t4][1] > 0.f) ? 1.f / ml_s[2 * t4][1] : 0.f;
  const float i1 = (2 * t4 + 1 < G && ml_s[2 * t4 + 1][1] > 0.f) ? 1.f / ml_s[2 * t4 + 1][1] : 0.f;
#pragma unroll
  for (int i = 0; i < MTW; ++i) {
    const int mt = warp + i * AM_WARPS;
    if (mt < NMT) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int head = 2 * t4 + (r & 1), dim = mt * 16 + g8 + (r >> 1) * 8;
        if (head < G) {
          const float o = acc[i][r] * ((r & 1) ? i1 : i0);
          const size_t idx = (size_t)b * a.Hq * D + (size_t)(hk * G + head) * D + dim;
          if (a.out) a.out[idx] = o;
          if (a.out_bf) a.out_bf[idx] = __float2bfloat16(o);
        }
      }
    }
  }
  AM_PROBE(10);
  trace_end(a.trace);
#undef AM_PROBE
}

template <int G, int D>
cudaError_t launch_am(const AttnDecodeArgs& a, cudaStream_t st, bool pdl) {
  auto kern = attn_decode_mma_kernel<G, D>;
  const size_t smem = AmGeo<D>::dyn_bytes(G);
  static PerDeviceFlag attr_set;
  if (smem > attr_set.here()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // batched-step kernels all ask for the maximum shared-memory carve-out: CTAs of consecutive kernels can then share an SM
    if (batched_carveout() >= 0 && (e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, batched_carveout())) != cudaSuccess) return e;
    attr_set.here() = smem;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(a.Hkv, a.chunk_tokens > 0 ? a.max_chunks : 1, a.B);
  cfg.blockDim = dim3(AM_NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, a);
}

}  // namespace

bool attn_decode_mma_supported(const AttnDecodeArgs& a) {
  const int G = a.Hkv > 0 ? a.Hq / a.Hkv : 0;
  const bool d_ok = a.D == 16 || a.D == 32 || a.D == 64 || a.D == 128 || a.D == 256;
  if (a.n_splits != 1) return false;                        // key ranges are split by chunks here, not by clusters
  if (a.chunk_tokens > 0 && (a.chunk_tokens % AM_TT || a.max_chunks < 1 || a.max_chunks > AM_MAX_CHUNKS ||
                             !a.part_o || !a.part_ml || !a.part_cnt)) return false;
  return d_ok && (G == 1 || G == 2 || G == 4);
}

cudaError_t launch_attn_decode_mma(const AttnDecodeArgs& a, cudaStream_t st, bool pdl) {
  if (!attn_decode_mma_supported(a)) return cudaErrorNotSupported;
  const int G = a.Hq / a.Hkv;
#define AM_CASE(GG, DD) if (G == GG && a.D == DD) return launch_am<GG, DD>(a, st, pdl)
  AM_CASE(1, 16); AM_CASE(1, 32); AM_CASE(1, 64); AM_CASE(1, 128); AM_CASE(1, 256);
  AM_CASE(2, 16); AM_CASE(2, 32); AM_CASE(2, 64); AM_CASE(2, 128); AM_CASE(2, 256);
  AM_CASE(4, 16); AM_CASE(4, 32); AM_CASE(4, 64); AM_CASE(4, 128); AM_CASE(4, 256);
#undef AM_CASE
  return cudaErrorNotSupported;
}
