// Decode attention for BATCHED rows, TMA front end (head_dim 64 / 128 / 256, 16- or 32-token KV pages).
//
// Same math, fused glue (PM-RoPE of q and of the new k, in-place KV append) and chunked work partition as the tile kernel
// of attention_mma.cu; what changes is how a 32-token tile reaches shared memory.  The KV pool is one 2-D tensor
// [layer * 2 * page * kv head * page tokens, head_dim]: the tokens of a (page, kv head) are consecutive rows, so ONE thread
// moves a tile with (K, V) x head_dim/64 TMA boxes of [32 tokens x 64 dims] at 32-token pages (4 KB each; two 2 KB boxes
// of 16 tokens at 16-token pages), 128-byte swizzle,
// conflict-free ldmatrix through the same XOR) instead of 2048 16-byte cp.async issued by all threads -- ncu had 37 % of
// the cp.async kernel's instructions in that issue loop and `barrier` as its top stall (five CTA barriers per tile).  The
// producer thread runs ahead of the 8 consumer warps through full / empty mbarriers (two consumer-only barriers per tile
// remain) and starts BEFORE griddepcontrol.wait: cached rows of earlier tokens are immutable.  Rows of a tile outside the
// key range (before a sliding window's start, past the last token, or a clamped second page) hold finite pool data and
// are masked to -inf before the softmax; the new token's row is overwritten in shared memory from registers.
#include "kernels.h"
#include "tc_common.cuh"

namespace {

constexpr int AM_CONS = 256, AM_WARPS = 8; // consumer threads / warps (two per scheduler)
constexpr int AM_NT = AM_CONS + 32;        // + the producer warp
constexpr int AM_TT = 32;                 // tokens per tile
constexpr int AM_NST = 2;                 // ring stages
constexpr int AM_BT_CACHE = 256;
constexpr int AM_MAX_CHUNKS = 16;         // chunked mode: chunks per (row, kv head)

__device__ __forceinline__ uint32_t am_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// byte offset of element (token `tok` of the tile, dim d) inside a K or V tile laid out as [D/64 slabs][32 tokens][128 B] with the
// TMA 128-byte swizzle (16-byte chunk index XOR token & 7; the tile base is 1024-byte aligned, tokens are the 128-byte rows)
__device__ __forceinline__ uint32_t at_off(int tok, int d) {
  return (uint32_t)((d >> 6) * (AM_TT * 128) + tok * 128 + ((((d & 63) >> 3) ^ (tok & 7)) << 4) + (d & 7) * 2);
}
__device__ __forceinline__ void am_mbar_init(uint64_t* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(am_smem(b)), "r"(count)); }
__device__ __forceinline__ void am_mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(am_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void am_mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(am_smem(b)) : "memory"); }
__device__ __forceinline__ void am_mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "AM_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra AM_DONE;\n\t"
      "bra AM_WAIT;\n\t"
      "AM_DONE:\n\t}" ::"r"(am_smem(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void am_cons_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 consumer warps only
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

template <int D> struct AtGeo {
  static constexpr int LD = D + 8;                         // padded row of the q operand tile (elements)
  static constexpr int NA = D / 64;                        // 64-dim slabs (one TMA box column each)
  static constexpr int TILE_BYTES = AM_TT * D * 2;         // K (or V) tile: [NA][32][128 B]
  static constexpr int STAGE_BYTES = 2 * TILE_BYTES;       // K tile + V tile
  static constexpr int KS = D / 16;                        // k-steps of the score MMA
  static constexpr int KSH = (KS + 3) / 4;                 // ... per k-quarter (four warps share a 16-token group)
  static constexpr int NMT = D / 16;                       // 16-dim m-tiles of the output MMA
  static constexpr int MTW = (NMT + AM_WARPS - 1) / AM_WARPS;   // m-tiles per warp
  // q / probability operand tiles hold G real rows + one shared zero row (the MMA's n = 8 columns beyond G read it).
  // 68 KB at head_dim 256 (+ 1 KB alignment slack): three CTAs per SM
  static constexpr size_t dyn_bytes(int G) { return 1024 + (size_t)AM_NST * STAGE_BYTES + ((size_t)(G + 1) * LD + (size_t)(G + 1) * (AM_TT + 8)) * sizeof(bf16); }
};

// grid (Hkv, chunks, B), 288 threads
template <int G, int D>
__global__ void __launch_bounds__(AM_NT, 3) attn_decode_tma_kernel(const __grid_constant__ CUtensorMap map_kv, AttnDecodeArgs a) {
  using Geo = AtGeo<D>;
  constexpr int LD = Geo::LD, KS = Geo::KS, KSH = Geo::KSH, NMT = Geo::NMT, MTW = Geo::MTW, NA = Geo::NA;
  static_assert(AM_TT == 32, "one token per lane in the softmax phase, two 16-token pages per tile");
  extern __shared__ __align__(16) unsigned char am_dyn[];
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(am_dyn) + 1023) & ~(uintptr_t)1023);   // [NST][K|V][NA][32][128 B]
  bf16* qb = reinterpret_cast<bf16*>(ring + AM_NST * Geo::STAGE_BYTES); // [G+1][LD]   rotated q, row G is zero
  bf16* pb = qb + (G + 1) * LD;                                         // [G+1][TT+8] probabilities of the tile, row G is zero
  __shared__ float cs[D / 2], sn[D / 2];
  __shared__ float knew[D], vnew[D];
  __shared__ int bt_s[AM_BT_CACHE];
  __shared__ float scp[4][G][AM_TT];                                    // partial q.k of the four k-quarters
  __shared__ float corr_s[8];
  __shared__ float ml_s[G][2];
  __shared__ float cw_s[AM_MAX_CHUNKS][G], cl_s[AM_MAX_CHUNKS][G];      // merge weights / sums of the chunks
  __shared__ int last_s;
  __shared__ __align__(8) uint64_t full_b[AM_NST], empty_b[AM_NST];

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
  const int lo = (!a.is_cross && a.window > 0) ? max(0, L - a.window) : 0;
  int n_chunks = 1, chunk = AM_TT * AM_MAX_CHUNKS * 4096;
  if (a.chunk_tokens > 0) {                                 // equal pieces of at most chunk_tokens keys
    n_chunks = min(max((L - lo + a.chunk_tokens - 1) / a.chunk_tokens, 1), a.max_chunks);
    chunk = ((L - lo + n_chunks - 1) / n_chunks + AM_TT - 1) / AM_TT * AM_TT;
  }
  if (!active || split >= n_chunks) return;                 // nothing to do for this CTA (exited CTAs release the dependents)
  if (tid < AM_CONS) {
    // loads first, shared-memory stores after: one round trip for the slot, the RoPE table and the block table
    float rc = 1.f, rs = 0.f;
    const bool rope_tab = a.rope_cs != nullptr && tid < D / 2;
    if (rope_tab) { rc = a.rope_cs[(size_t)b * D + tid]; rs = a.rope_cs[(size_t)b * D + D / 2 + tid]; }
    const int btv = (tid < a.bt_stride) ? bt[tid] : 0;                  // AM_BT_CACHE == AM_CONS entries
    if (rope_tab) { cs[tid] = rc; sn[tid] = rs; }
    if (!a.rope_cs) for (int i = tid; i < D / 2; i += AM_CONS) { float s_, c_; sincosf(pos * a.inv_freq[i], &s_, &c_); cs[i] = c_; sn[i] = s_; }
    bt_s[tid] = btv;
  } else if (lane == 0) {
    for (int i = 0; i < AM_NST; ++i) { am_mbar_init(&full_b[i], 1); am_mbar_init(&empty_b[i], AM_WARPS); }
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  static_assert(AM_BT_CACHE == AM_CONS && D / 2 <= AM_CONS, "prologue mapping");
  for (int i = tid; i < (G + 1) * LD; i += AM_NT) qb[i] = __float2bfloat16(0.f);
  for (int i = tid; i < (G + 1) * (AM_TT + 8); i += AM_NT) pb[i] = __float2bfloat16(0.f);
  const int t_begin = lo + split * chunk, t_end = min(L, t_begin + chunk);
  const bool has_new = (!a.is_cross) && (t_end == L) && (t_end > t_begin);
  const int t0a = t_begin & ~(PT - 1);                        // tiles start on a page boundary (PT = 16 or 32); keys before t_begin are masked
  const int n_tiles = (t_end > t_begin) ? (t_end - t0a + AM_TT - 1) / AM_TT : 0;
  const int last_page = (t_end > t_begin) ? (t_end - 1) / PT : 0;
  __syncthreads();                                          // the only CTA-wide barrier: roles split below
  auto page_of = [&](int t) -> int { const int pi = t / PT; return pi < AM_BT_CACHE ? bt_s[pi] : bt[pi]; };
  auto stage_ptr = [&](int stage, int kv) -> unsigned char* { return ring + (size_t)stage * Geo::STAGE_BYTES + (size_t)kv * Geo::TILE_BYTES; };

  if (warp == AM_WARPS) {
    // ===== producer: one thread.  Row of (layer, K|V, page, kv head, token 0) in the pool tensor =
    //       (((layer * 2 + kv) * n_pages + page) * Hkv + hk) * PT; a box is the PT tokens of that page x 64 dims: with
    //       32-token pages (the default) a tile is 8 boxes of 4 KB, with 16-token pages 16 boxes of 2 KB (measured
    //       equal within noise: 26.9 vs 27.2 us per layer at contexts U[0,900) -- the request size is not the limit). =====
    if (lane == 0) {
      const int n_pages = a.pool.n_pages, ppt = AM_TT / PT;   // pages per tile
      for (int ti = 0; ti < n_tiles; ++ti) {
        const int stage = ti % AM_NST;
        if (ti >= AM_NST) am_mbar_wait(&empty_b[stage], ((ti / AM_NST) - 1) & 1);
        am_mbar_expect_tx(&full_b[stage], (uint32_t)Geo::STAGE_BYTES);
        for (int half = 0; half < ppt; ++half) {
          const int pi = min(t0a / PT + ppt * ti + half, last_page);       // a page past the range: the last one again (finite, masked)
          const int page = page_of(pi * PT);
#pragma unroll
          for (int kv = 0; kv < 2; ++kv) {
            const int row = (((a.layer * 2 + kv) * n_pages + page) * a.Hkv + hk) * PT;
            unsigned char* dst = stage_ptr(stage, kv) + half * (PT * 128);
#pragma unroll
            for (int sl_ = 0; sl_ < NA; ++sl_) tc::tma_load_2d(dst + sl_ * (AM_TT * 128), &map_kv, sl_ * 64, row, &full_b[stage]);
          }
        }
      }
    }
    return;
  }

  // ===== consumer warps =====
  AM_PROBE(1);
  pdl_wait();
  trace_begin(a.trace);
  AM_PROBE(2);
  // ---- the producer kernel's outputs: raw q and the new k/v.  A thread loads both elements of a rotation pair
  //      (j, j + D/2), rotates in registers (PM-RoPE at the row's progress position) and stores the MMA operand. ----
  {
    constexpr int QP = (G * D / 2 + AM_CONS - 1) / AM_CONS;
    float x1[QP], x2[QP];
    float k1 = 0.f, k2 = 0.f, vv = 0.f;
    const float* qp = a.q + (size_t)b * a.q_stride + (size_t)(hk * G) * D;
#pragma unroll
    for (int u = 0; u < QP; ++u) {
      const int i = tid + u * AM_CONS, g = i / (D / 2), j = i - g * (D / 2);
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
      const int i = tid + u * AM_CONS, g = i / (D / 2), j = i - g * (D / 2);
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
  static_assert(D <= AM_CONS, "one thread per element of the new k/v row");
  am_cons_sync();
  AM_PROBE(3);
  if (has_new) {   // append to the page (K post-RoPE), visible to later steps
    const int t = L - 1, page = page_of(t), off = t % PT;
    bf16* kd = a.pool.ptr(a.layer, 0, page) + ((size_t)hk * PT + off) * D;
    bf16* vd = a.pool.ptr(a.layer, 1, page) + ((size_t)hk * PT + off) * D;
    if (tid < D) { kd[tid] = __float2bfloat16(knew[tid]); vd[tid] = __float2bfloat16(vnew[tid]); }
  }
  // q fragments (B operand of the score MMA): B[k = dim][n = head] from qb[head][dim]; this warp's k-half only
  const int tg = warp & 1, kh = warp >> 1;            // score phase (all 8 warps): 16-token group, k-quarter
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
    const int stage = ti % AM_NST;
    am_mbar_wait(&full_b[stage], (ti / AM_NST) & 1);          // the tile has landed (and the zero fill is visible)
    if (ti == 0) AM_PROBE(5);
    if (ti == 1) AM_PROBE(6);
    if (ti == 2) AM_PROBE(7);
    unsigned char* kb = stage_ptr(stage, 0);
    unsigned char* vb = stage_ptr(stage, 1);
    const int tile_t0 = t0a + ti * AM_TT;
    if (has_new && L - 1 >= tile_t0 && L - 1 < tile_t0 + AM_TT) {       // CTA-uniform: the new token's row comes from registers
      const int tok = L - 1 - tile_t0;
      if (tid < D) {
        *reinterpret_cast<bf16*>(kb + at_off(tok, tid)) = __float2bfloat16(knew[tid]);
        *reinterpret_cast<bf16*>(vb + at_off(tok, tid)) = __float2bfloat16(vnew[tid]);
      }
      am_cons_sync();
    }
    // ---- partial scores: warp = (16-token group tg) x (k-quarter kh) ----
    {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      const int atok = tg * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ad0 = (lane >> 4) * 8 + kh * KSH * 16;
#pragma unroll
      for (int i = 0; i < KSH; ++i) {
        if (kh * KSH + i < KS) {
          uint32_t af[4];
          ldsm_x4(af, kb + at_off(atok, ad0 + i * 16));
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
    am_cons_sync();
    // ---- scale / softcap / mask + online softmax: warp h owns head h, lane = token of the tile ----
    if (warp < G) {
      float s = ((scp[0][warp][lane] + scp[1][warp][lane]) + (scp[2][warp][lane] + scp[3][warp][lane])) * a.scale;
      if (a.softcap > 0.f) {
        const float e2 = __expf(2.f * s * inv_cap);
        s = a.softcap * (1.f - __fdividef(2.f, e2 + 1.f));
      }
      if (tile_t0 + lane >= t_end || tile_t0 + lane < t_begin) s = -INFINITY;
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
    am_cons_sync();
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
            ldsm_x4_t(vf, vb + at_off(kk * 16 + (lane & 7) + (lane >> 4) * 8, mt * 16 + ((lane >> 3) & 1) * 8));
            mma_bf16_16816(acc[i], vf, pf[kk]);
          }
        }
      }
    }
    // this warp is done with the stage (its K reads ended before the barriers above); scp / pb / corr_s are rewritten only
    // after the next tile's barriers, which every warp reaches after its own P.V of this tile
    __syncwarp();
    if (lane == 0) am_mbar_arrive(&empty_b[stage]);
  }
  AM_PROBE(8);

  if (warp < G && lane == 0) { ml_s[warp][0] = m_run; ml_s[warp][1] = l_run; }
  if (n_chunks > 1) {
    // ---- several chunks in this row's range: park the unnormalised partial, last arriver merges ----
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
    am_cons_sync();
    // one acq_rel atomic by thread 0 orders the CTA's partial (observed through the barrier above) before the count and
    // the other chunks' partials before this CTA's reads below -- no membar.gl by all 256 threads
    if (tid == 0) {
      unsigned old;
      asm volatile("atom.add.acq_rel.gpu.u32 %0, [%1], 1;" : "=r"(old) : "l"(a.part_cnt + (size_t)b * a.Hkv + hk) : "memory");
      last_s = (old == (unsigned)(n_chunks - 1));
    }
    am_cons_sync();
    if (!last_s) { trace_end(a.trace); return; }
    if (tid == 0) a.part_cnt[(size_t)b * a.Hkv + hk] = 0;   // every chunk has arrived: ready for the next launch
    // partial outputs first (independent loads, in flight while the weights are computed)
    constexpr int OPT = (G * D + AM_CONS - 1) / AM_CONS;
    float po_r[OPT][AM_MAX_CHUNKS];
#pragma unroll
    for (int u = 0; u < OPT; ++u)
#pragma unroll
      for (int c = 0; c < AM_MAX_CHUNKS; ++c) {
        const int i = tid + u * AM_CONS;
        po_r[u][c] = (c < n_chunks && i < G * D) ? __ldcg(a.part_o + (slot + c) * G * D + i) : 0.f;
      }
    if (tid < G * AM_MAX_CHUNKS) {                            // (chunk, head) -> m and l, one load each
      const int c = tid / G, g = tid % G;
      const bool ok = c < n_chunks;
      cw_s[c][g] = ok ? __ldcg(a.part_ml + ((slot + c) * G + g) * 2) : -INFINITY;
      cl_s[c][g] = ok ? __ldcg(a.part_ml + ((slot + c) * G + g) * 2 + 1) : 0.f;
    }
    am_cons_sync();
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
    am_cons_sync();
#pragma unroll
    for (int u = 0; u < OPT; ++u) {
      const int i = tid + u * AM_CONS;
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
  // ---- one chunk: normalise and store straight from the accumulator registers ----
  am_cons_sync();
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
cudaError_t launch_at(const AttnDecodeArgs& a, cudaStream_t st, bool pdl) {
  auto kern = attn_decode_tma_kernel<G, D>;
  const size_t smem = AtGeo<D>::dyn_bytes(G);
  static PerDeviceFlag attr_set;
  if (!attr_set.here()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // batched-step kernels all ask for the maximum shared-memory carve-out: CTAs of consecutive kernels can then share an SM
    if (batched_carveout() >= 0 && (e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, batched_carveout())) != cudaSuccess) return e;
    attr_set.here() = 1;
  }
  // the whole pool as one 2-D tensor of token rows; box = one (page, kv head) x 64 dims
  CUtensorMap map;
  const uint64_t rows = (uint64_t)a.n_layers_pool * 2 * a.pool.n_pages * a.pool.Hkv * a.pool.page_tokens;
  if (!tc::make_map_2d(&map, a.pool.base, rows, (uint64_t)D, (uint64_t)D, (uint32_t)a.pool.page_tokens)) return cudaErrorNotSupported;
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
  return cudaLaunchKernelEx(&cfg, kern, map, a);
}

}  // namespace

bool attn_decode_tma_supported(const AttnDecodeArgs& a) {
  const int G = a.Hkv > 0 ? a.Hq / a.Hkv : 0;
  if (a.n_splits != 1 || (a.pool.page_tokens != 16 && a.pool.page_tokens != 32) || a.n_layers_pool <= 0 || a.pool.D != a.D || a.pool.Hkv != a.Hkv) return false;
  if (a.chunk_tokens > 0 && (a.chunk_tokens % AM_TT || a.max_chunks < 1 || a.max_chunks > AM_MAX_CHUNKS ||
                             !a.part_o || !a.part_ml || !a.part_cnt)) return false;
  return (a.D == 64 || a.D == 128 || a.D == 256) && (G == 1 || G == 2 || G == 4) && !(G == 4 && a.D == 256);
}

cudaError_t launch_attn_decode_tma(const AttnDecodeArgs& a, cudaStream_t st, bool pdl) {
  if (!attn_decode_tma_supported(a)) return cudaErrorNotSupported;
  const int G = a.Hq / a.Hkv;
#define AT_CASE(GG, DD) if (G == GG && a.D == DD) return launch_at<GG, DD>(a, st, pdl)
  AT_CASE(1, 64); AT_CASE(1, 128); AT_CASE(1, 256);
  AT_CASE(2, 64); AT_CASE(2, 128); AT_CASE(2, 256);
  AT_CASE(4, 64); AT_CASE(4, 128);
#undef AT_CASE
  return cudaErrorNotSupported;
}
