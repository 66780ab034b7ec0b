// Prefill attention on the tensor cores (tcgen05 / TMEM / TMA), head_dim 64..256, varlen-packed requests.
// Replaces the q_len>1 uses of HF:modeling_t5gemma.py:209-240 (encoder bidirectional, decoder causal, cross) incl. the
// Gemma-2 attn-logit softcap and sliding windows.
//
// One CTA = 128 queries of one query head of one request, ONE pass over the keys in blocks of 64 (online softmax):
//   S_j = Q K_j^T          UMMA, A = Q and B = K_j from shared memory, fp32 scores in TMEM (two S buffers, ping-pong)
//   softmax                4 warps, thread = query row = TMEM lane: tcgen05.ld S_j -> scale / softcap / mask -> running max
//                          in the log2 domain -> P_j = exp2(y - m) as packed bf16 written back with tcgen05.st OVER S_j
//   O  += P_j V_j          UMMA with the A operand read straight from TMEM (no shared-memory round trip for P), B = V_j^T
// The MMA thread issues S_{j+1} before it waits for P_j, so the tensor pipe computes the next scores while the softmax
// warps work on the current ones.  The running max is only raised when it grows by more than 8 (a factor 256 in P, exact in
// bf16/fp32 range): the O accumulator in TMEM is then rescaled in place by the owning warp, which happens a handful of
// times per row instead of once per block; the 1/l normalisation is applied in the epilogue.
// Separate TMA rings for K (3 stages, freed as soon as S_j has been issued and completed) and V^T (2 stages, freed after
// O += P_j V_j) keep K two blocks ahead of the score MMAs.  Interior key blocks (no causal / window / length edge inside
// the tile) skip the per-element mask arithmetic.  Softcap uses one MUFU (tanh.approx, rel. error 2^-11 -- below the bf16
// rounding of P) so a score costs two MUFU operations (tanh, ex2).
// V is consumed as it is stored ([key][dim], dims contiguous): an MN-major B operand -- no transposed copy (round 1 and the
// first version of this kernel staged V^T through a transpose kernel and an 8-token-aligned scratch tensor).
#include "kernels.h"
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int FA_BQ = 128;       // queries per CTA (UMMA M)
constexpr int FA_BK = 64;        // keys per block (one 128-byte swizzle atom of V^T, UMMA N of the score MMA)
constexpr int FA_KST = 3;        // K ring stages
constexpr int FA_VST = 2;        // V^T ring stages
constexpr int FA_THREADS = 224;  // warp 0: Q + K TMA, warp 1: MMA + TMEM alloc, warps 2-5: softmax / epilogue, warp 6: V^T TMA

template <int D>
struct __align__(1024) FaSmem {
  static constexpr int NA = D / 64;                         // 64-element K atoms along head_dim
  unsigned char q[NA][FA_BQ * 128];                         // Q tile: NA atoms of [128 rows x 128 B]
  unsigned char k[FA_KST][NA][FA_BK * 128];                 // K block stages: NA atoms of [64 keys x 128 B]
  unsigned char v[FA_VST][NA][FA_BK * 128];                 // V block stages: NA slabs of [64 keys x 64 dims (128 B)]
  uint64_t q_full, k_full[FA_KST], k_empty[FA_KST], v_full[FA_VST], v_empty[FA_VST], s_full[2], p_ready[2], pv_done;
  uint32_t tmem_base;
};

struct FaParams {
  const int* q_seg_off; const int* k_seg_off;
  int Hq, Hkv, causal, window;
  float scale, softcap;
  bf16* out;                                               // [Tq, Hq*D]
};

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
               "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                 "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                 "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem: row m in lane m, 16 bf16 of K packed in 8 columns] * B[smem descriptor]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D>
__global__ void __launch_bounds__(FA_THREADS, 1)
attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                       const __grid_constant__ CUtensorMap map_v, FaParams p) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  FaSmem<D>& S = *reinterpret_cast<FaSmem<D>*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~(uintptr_t)1023);
  constexpr int NA = D / 64;
  constexpr uint32_t O_COL = 0, S_COL = 256, S_STRIDE = 64;   // TMEM: O at columns [0,D), S/P buffers at [256,320) and [320,384)
  constexpr int TMEM_COLS = 512;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seg = blockIdx.z, h = blockIdx.y, hk = h / (p.Hq / p.Hkv);
  pdl_launch_dependents();           // programmatic dependent in the prefill chain: nothing is read before the wait
  pdl_wait();
  const int q_beg = p.q_seg_off[seg], Lq = p.q_seg_off[seg + 1] - q_beg;
  const int k_beg = p.k_seg_off[seg], Lk = p.k_seg_off[seg + 1] - k_beg;
  const int q0 = blockIdx.x * FA_BQ;                        // first query (within the request) of this tile
  if (q0 >= Lq) return;                                     // uniform: tiles beyond this request's length
  // key range this tile can attend to (block granularity; exact masks are applied per element in the edge blocks)
  int k_hi = Lk, k_lo = 0;
  if (p.causal) { k_hi = min(Lk, q0 + FA_BQ); if (p.window > 0) k_lo = max(0, q0 - p.window + 1); }
  else if (p.window > 0) { k_lo = max(0, q0 - p.window); k_hi = min(Lk, q0 + FA_BQ + p.window); }
  const int b0 = k_lo / FA_BK, nb = (k_hi + FA_BK - 1) / FA_BK - b0;      // key blocks [b0, b0+nb)

  if (threadIdx.x == 0) {
    mbar_init(&S.q_full, 1);
    for (int i = 0; i < FA_KST; ++i) { mbar_init(&S.k_full[i], 1); mbar_init(&S.k_empty[i], 1); }
    for (int i = 0; i < FA_VST; ++i) { mbar_init(&S.v_full[i], 1); mbar_init(&S.v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&S.s_full[i], 1); mbar_init(&S.p_ready[i], 128); }
    mbar_init(&S.pv_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;

  if (warp == 0) {
    // ===== TMA producer: Q once, then the K ring =====
    if (lane == 0 && nb > 0) {
      mbar_expect_tx(&S.q_full, NA * FA_BQ * 128);
      for (int a = 0; a < NA; ++a) tma_load_2d(S.q[a], &map_q, h * D + a * 64, q_beg + q0, &S.q_full);
      int st = 0; uint32_t round = 0;
      for (int j = 0; j < nb; ++j) {
        if (round > 0) mbar_wait(&S.k_empty[st], (round - 1) & 1);
        mbar_expect_tx(&S.k_full[st], NA * FA_BK * 128);
        for (int a = 0; a < NA; ++a) tma_load_2d(S.k[st][a], &map_k, hk * D + a * 64, k_beg + (b0 + j) * FA_BK, &S.k_full[st]);
        if (++st == FA_KST) { st = 0; ++round; }
      }
    }
  } else if (warp == 6) {
    // ===== TMA producer: the V^T ring =====
    if (lane == 0) {
      for (int j = 0; j < nb; ++j) {
        const int st = j & 1;
        if (j >= FA_VST) mbar_wait(&S.v_empty[st], ((j >> 1) - 1) & 1);
        mbar_expect_tx(&S.v_full[st], D * 128);
        for (int a = 0; a < NA; ++a) tma_load_2d(S.v[st][a], &map_v, hk * D + a * 64, k_beg + (b0 + j) * FA_BK, &S.v_full[st]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0 && nb > 0) {
      constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FA_BK >> 3) << 17) | ((uint32_t)(FA_BQ >> 4) << 24);
      constexpr uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) /* B is MN-major */ | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(FA_BQ >> 4) << 24);
      mbar_wait(&S.q_full, 0);
      int kst = 0; uint32_t kround = 0;
      auto issue_scores = [&](int j) {
        mbar_wait(&S.k_full[kst], kround & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem + S_COL + (uint32_t)(j & 1) * S_STRIDE;
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          const uint32_t qa = smem_u32(S.q[a]), ka = smem_u32(S.k[kst][a]);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(d, umma_desc_sw128(qa + kk * 32), umma_desc_sw128(ka + kk * 32), idesc_s, (a | kk) ? 1u : 0u);
        }
        umma_commit(&S.k_empty[kst]);
        umma_commit(&S.s_full[j & 1]);
        if (++kst == FA_KST) { kst = 0; ++kround; }
      };
      issue_scores(0);
      for (int j = 0; j < nb; ++j) {
        if (j + 1 < nb) issue_scores(j + 1);                 // runs on the tensor pipe while the softmax warps work on S_j
        const int b = j & 1;
        mbar_wait(&S.v_full[b], (j >> 1) & 1);
        mbar_wait(&S.p_ready[b], (j >> 1) & 1);             // P_j sits in TMEM over S_j
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t pa = tmem + S_COL + (uint32_t)b * S_STRIDE, va = smem_u32(S.v[b][0]);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          {
            // MN-major B operand with the 128-byte swizzle: a 64-dim slab of the tile is 64 key rows x 128 B, i.e. 8-row
            // swizzle atoms of 1024 B along K (stride-dimension offset) and slabs 8 KB apart along N (leading-dimension
            // offset); one MMA consumes 16 keys = 2048 B of every slab
            constexpr uint64_t lbo = FA_BK * 128, sbo = 1024;
            const uint64_t db = (uint64_t)(((va + kk * 2048) & 0x3FFFF) >> 4) | ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1ull << 46) | (2ull << 61);
            umma_bf16_ts(tmem + O_COL, pa + kk * 8, db, idesc_o, (j > 0 || kk) ? 1u : 0u);
          }
        umma_commit(&S.v_empty[b]);
        umma_commit(&S.pv_done);
      }
    }
  } else {
    // ===== softmax / epilogue: thread = query row = TMEM lane =====
    const int qtr = warp & 3;
    const int row = qtr * 32 + lane;
    const int qi = q0 + row;                                  // query index inside the request
    const uint32_t lane_addr = tmem + ((uint32_t)(qtr * 32) << 16);
    constexpr float LOG2E = 1.4426950408889634f;
    const bool capped = p.softcap > 0.f;
    const float k_in = capped ? p.scale / p.softcap : p.scale * LOG2E;
    const float k_out = p.softcap * LOG2E;
    float m = -INFINITY, l = 0.f;                             // running max (log2 domain) and sum of P
    for (int j = 0; j < nb; ++j) {
      const int b = j & 1;
      const int kb = (b0 + j) * FA_BK, ke = kb + FA_BK - 1;   // first / last key of the block
      bool interior = ke < Lk;
      if (p.causal) { interior = interior && ke <= q0; if (p.window > 0) interior = interior && kb > q0 + FA_BQ - 1 - p.window; }
      else if (p.window > 0) interior = interior && kb >= q0 + FA_BQ - 1 - p.window && ke <= q0 + p.window;
      mbar_wait(&S.s_full[b], (j >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[FA_BK];
      tmem_ld32_nowait(lane_addr + S_COL + (uint32_t)b * S_STRIDE, r);
      tmem_ld32_nowait(lane_addr + S_COL + (uint32_t)b * S_STRIDE + 32, r + 32);
      tmem_ld_wait();
      float bm = -INFINITY;
      if (interior) {
#pragma unroll
        for (int c = 0; c < FA_BK; ++c) {
          float v = __uint_as_float(r[c]) * k_in;
          if (capped) v = tanh_approx(v) * k_out;
          r[c] = __float_as_uint(v);
          bm = fmaxf(bm, v);
        }
      } else {
#pragma unroll
        for (int c = 0; c < FA_BK; ++c) {
          const int kk = kb + c;
          bool ok = kk < Lk;
          if (p.causal) { ok = ok && kk <= qi; if (p.window > 0) ok = ok && kk > qi - p.window; }
          else if (p.window > 0) ok = ok && (kk >= qi - p.window) && (kk <= qi + p.window);
          float v = __uint_as_float(r[c]) * k_in;
          if (capped) v = tanh_approx(v) * k_out;
          v = ok ? v : -INFINITY;
          r[c] = __float_as_uint(v);
          bm = fmaxf(bm, v);
        }
      }
      // lazy running max: raise it only when it would grow by more than 8 (P stays <= 2^8)
      const float mn = fmaxf(m, bm);
      float alpha = 1.f;
      bool raise = false;
      if (m == -INFINITY) m = mn;                             // nothing accumulated for this row yet (its O row is 0)
      else if (mn > m + 8.f) { alpha = ex2_approx(m - mn); m = mn; raise = true; }
      if (__any_sync(0xffffffffu, raise)) {                   // rescale this warp's O rows in place (rare)
        l *= alpha;
        mbar_wait(&S.pv_done, (j - 1) & 1);                   // O += P_{j-1} V_{j-1} has landed; P_j V_j is not issued yet
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c = 0; c < D; c += 32) {
          uint32_t o[32];
          tmem_ld32_nowait(lane_addr + O_COL + (uint32_t)c, o);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t) o[t] = __float_as_uint(__uint_as_float(o[t]) * alpha);
          tmem_st32(lane_addr + O_COL + (uint32_t)c, o);
        }
        tmem_st_wait();
      }
      const float ms = (m == -INFINITY) ? 0.f : m;
      float sum = 0.f;
      uint32_t pk[FA_BK / 2];
#pragma unroll
      for (int c = 0; c < FA_BK; c += 2) {
        const float a = ex2_approx(__uint_as_float(r[c]) - ms), bb = ex2_approx(__uint_as_float(r[c + 1]) - ms);
        sum += a + bb;
        __nv_bfloat162 t2 = __floats2bfloat162_rn(a, bb);
        pk[c >> 1] = *reinterpret_cast<uint32_t*>(&t2);
      }
      l += sum;
      tmem_st32(lane_addr + S_COL + (uint32_t)b * S_STRIDE, pk);   // P_j (bf16 pairs) over the first 32 columns of S_j
      tmem_st_wait();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&S.p_ready[b])) : "memory");
    }
    // epilogue: O / l -> bf16 -> out[token][h*D + d]
    if (nb > 0) {
      mbar_wait(&S.pv_done, (nb - 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const float inv_l = l > 0.f ? 1.f / l : 0.f;
    bf16* orow = p.out + (size_t)(q_beg + min(qi, Lq - 1)) * p.Hq * D + (size_t)h * D;
#pragma unroll 1
    for (int c = 0; c < D; c += 32) {
      uint32_t o[32];
      if (nb > 0) { tmem_ld32_nowait(lane_addr + O_COL + (uint32_t)c, o); tmem_ld_wait(); }   // .sync.aligned: whole warp
      if (qi < Lq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t w[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(o[g * 8 + 2 * t]) * inv_l, __uint_as_float(o[g * 8 + 2 * t + 1]) * inv_l);
            w[t] = (nb > 0) ? *reinterpret_cast<uint32_t*>(&t2) : 0u;
          }
          reinterpret_cast<uint4*>(orow + c)[g] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

template <int D>
cudaError_t launch_fa(const AttnPrefillArgs& a, int Tk, int n_seg, int max_lq, cudaStream_t st, bool pdl) {
  CUtensorMap mq, mk, mv;
  if (!make_map_2d(&mq, a.q, a.Tq, (uint64_t)a.Hq * D, (uint64_t)a.Hq * D, FA_BQ) ||
      !make_map_2d(&mk, a.k, Tk, (uint64_t)a.Hkv * D, (uint64_t)a.Hkv * D, FA_BK) ||
      !make_map_2d(&mv, a.v, Tk, (uint64_t)a.Hkv * D, (uint64_t)a.Hkv * D, FA_BK))
    return cudaErrorNotSupported;
  auto kern = attn_prefill_tc_kernel<D>;
  const size_t smem = sizeof(FaSmem<D>) + 1024;
  static PerDeviceFlag attr_set;
  if (!attr_set.here()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set.here() = 1;
  }
  FaParams p{a.q_seg_off, a.k_seg_off, a.Hq, a.Hkv, a.causal, a.window, a.scale, a.softcap, a.out};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((max_lq + FA_BQ - 1) / FA_BQ, a.Hq, n_seg);
  cfg.blockDim = dim3(FA_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, mq, mk, mv, p);
}

}  // namespace

bool attn_prefill_tc_supported(int D) { return D == 64 || D == 128 || D == 256; }

cudaError_t launch_attn_prefill_tc(const AttnPrefillArgs& a, int Tk, int n_seg, int max_lq, cudaStream_t st, bool pdl) {
  if (a.Tq <= 0 || n_seg <= 0) return cudaSuccess;
  switch (a.D) {
    case 64: return launch_fa<64>(a, Tk, n_seg, max_lq, st, pdl);
    case 128: return launch_fa<128>(a, Tk, n_seg, max_lq, st, pdl);
    case 256: return launch_fa<256>(a, Tk, n_seg, max_lq, st, pdl);
    default: return cudaErrorNotSupported;
  }
}
