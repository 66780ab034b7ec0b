// Prefill attention on the tensor cores (tcgen05 / TMEM / TMA), head_dim 64..256, varlen-packed requests.
// Replaces the q_len>1 uses of HF:modeling_t5gemma.py:209-240 (encoder bidirectional, decoder causal, cross) incl. the
// Gemma-2 attn-logit softcap and sliding windows.
//
// One CTA = 128 queries of one query head of one request.  Keys are processed in blocks of 64 in TWO passes so that no
// accumulator rescaling is ever needed:
//   pass 1: S = Q K^T (UMMA, fp32 in TMEM) -> per-row running max / sum (one thread owns one query row = one TMEM lane)
//   pass 2: S again -> P = exp(S - m) / l as bf16 written by the softmax threads straight into the 128-byte-swizzled
//           K-major shared-memory layout of a UMMA A operand -> O += P V  (V^T tile as B operand, fp32 O in TMEM)
// The recomputed Q K^T costs 1.5x the attention flops but keeps the kernel a straight pipeline:
// TMA producer warp -> MMA issuer thread -> 4 softmax/epilogue warps, all handshakes on mbarriers / tcgen05.commit.
// V is consumed as V^T [Hkv*D, tokens] (K-major B operand); `launch_transpose_v` produces it once per layer.
#include "kernels.h"
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int FA_BQ = 128;       // queries per CTA (UMMA M)
constexpr int FA_BK = 64;        // keys per block (one 128-byte swizzle atom of P / V^T, UMMA N of the score MMA)
constexpr int FA_THREADS = 192;  // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-5 softmax/epilogue

template <int D>
struct __align__(1024) FaSmem {
  static constexpr int NA = D / 64;                         // 64-element K atoms along head_dim
  unsigned char q[NA][FA_BQ * 128];                         // Q tile: NA atoms of [128 rows x 128 B]
  unsigned char k[2][NA][FA_BK * 128];                      // K block stages: NA atoms of [64 keys x 128 B]
  unsigned char vt[2][D * 128];                             // V^T block stages: [D rows x 64 keys (128 B)]
  unsigned char p[FA_BQ * 128];                             // P block: [128 queries x 64 keys (128 B)]
  uint64_t q_full, kv_full[2], kv_empty[2], s_full, sm_done, o_full;
  uint32_t tmem_base;
};

struct FaParams {
  const int* q_seg_off; const int* k_seg_off; const int* vt_seg_off;   // vt_seg_off: 8-aligned column of each request in V^T
  int Hq, Hkv, causal, window;
  float scale, softcap;
  bf16* out;                                               // [Tq, Hq*D]
};

template <int D>
__global__ void __launch_bounds__(FA_THREADS, 1)
attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                       const __grid_constant__ CUtensorMap map_vt, FaParams p) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  FaSmem<D>& S = *reinterpret_cast<FaSmem<D>*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~(uintptr_t)1023);
  constexpr int NA = D / 64;
  constexpr int O_COL = 0, S_COL = 256;                     // TMEM: O at columns [0,D), S at [256,320) (N-aligned bases)
  constexpr int TMEM_COLS = 512;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seg = blockIdx.z, h = blockIdx.y, hk = h / (p.Hq / p.Hkv);
  const int q_beg = p.q_seg_off[seg], Lq = p.q_seg_off[seg + 1] - q_beg;
  const int k_beg = p.k_seg_off[seg], Lk = p.k_seg_off[seg + 1] - k_beg;
  const int vt_beg = p.vt_seg_off[seg];
  const int q0 = blockIdx.x * FA_BQ;                        // first query (within the request) of this tile
  if (q0 >= Lq) return;                                     // uniform: tiles beyond this request's length
  // key range this tile can attend to (block granularity; exact masks are applied per element)
  int k_hi = Lk, k_lo = 0;
  if (p.causal) { k_hi = min(Lk, q0 + FA_BQ); if (p.window > 0) k_lo = max(0, q0 - p.window + 1); }
  else if (p.window > 0) { k_lo = max(0, q0 - p.window); k_hi = min(Lk, q0 + FA_BQ + p.window); }
  const int b0 = k_lo / FA_BK, nb = (k_hi + FA_BK - 1) / FA_BK - b0;      // key blocks [b0, b0+nb)
  const int nsteps = 2 * nb;

  if (threadIdx.x == 0) {
    mbar_init(&S.q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&S.kv_full[i], 1); mbar_init(&S.kv_empty[i], 1); }
    mbar_init(&S.s_full, 1); mbar_init(&S.sm_done, 128); mbar_init(&S.o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_vt) : "memory");
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
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(&S.q_full, NA * FA_BQ * 128);
      for (int a = 0; a < NA; ++a) tma_load_2d(S.q[a], &map_q, h * D + a * 64, q_beg + q0, &S.q_full);
      for (int j = 0; j < nsteps; ++j) {
        const int st = j & 1, f = j >> 1;
        const bool pass2 = j >= nb;
        const int blk = b0 + (pass2 ? j - nb : j);
        if (f >= 1) mbar_wait(&S.kv_empty[st], (f - 1) & 1);
        mbar_expect_tx(&S.kv_full[st], NA * FA_BK * 128 + (pass2 ? D * 128 : 0));
        for (int a = 0; a < NA; ++a) tma_load_2d(S.k[st][a], &map_k, hk * D + a * 64, k_beg + blk * FA_BK, &S.kv_full[st]);
        if (pass2) tma_load_2d(S.vt[st], &map_vt, vt_beg + blk * FA_BK, hk * D, &S.kv_full[st]);   // 16-byte aligned start
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FA_BK >> 3) << 17) | ((uint32_t)(FA_BQ >> 4) << 24);
      constexpr uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(FA_BQ >> 4) << 24);
      mbar_wait(&S.q_full, 0);
      for (int j = 0; j < nsteps; ++j) {
        const int st = j & 1, f = j >> 1;
        const bool pass2 = j >= nb;
        mbar_wait(&S.kv_full[st], f & 1);
        if (j > 0) mbar_wait(&S.sm_done, (j - 1) & 1);       // softmax threads are done with the previous S
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          const uint32_t qa = smem_u32(S.q[a]), ka = smem_u32(S.k[st][a]);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem + S_COL, umma_desc_sw128(qa + kk * 32), umma_desc_sw128(ka + kk * 32), idesc_s, (a | kk) ? 1u : 0u);
        }
        umma_commit(&S.s_full);
        if (!pass2) {
          umma_commit(&S.kv_empty[st]);
        } else {
          mbar_wait(&S.sm_done, j & 1);                      // P block is in shared memory
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t pa = smem_u32(S.p), va = smem_u32(S.vt[st]);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem + O_COL, umma_desc_sw128(pa + kk * 32), umma_desc_sw128(va + kk * 32), idesc_o, (j > nb || kk) ? 1u : 0u);
          umma_commit(&S.kv_empty[st]);
        }
      }
      umma_commit(&S.o_full);
    }
  } else {
    // ===== softmax / epilogue: thread = query row = TMEM lane =====
    const int qtr = warp & 3;
    const int row = qtr * 32 + lane;
    const int qi = q0 + row;                                  // query index inside the request
    const uint32_t lane_addr = (uint32_t)(qtr * 32) << 16;
    const float inv_cap = p.softcap > 0.f ? 1.f / p.softcap : 0.f;
    float m = -INFINITY, l = 0.f, inv_l = 0.f;
    for (int j = 0; j < nsteps; ++j) {
      const bool pass2 = j >= nb;
      const int blk = b0 + (pass2 ? j - nb : j);
      mbar_wait(&S.s_full, j & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float s[FA_BK];
#pragma unroll
      for (int c = 0; c < FA_BK; c += 16) tmem_ld16(tmem + lane_addr + (uint32_t)(S_COL + c), s + c);
      float bm = -INFINITY;
#pragma unroll
      for (int c = 0; c < FA_BK; ++c) {
        const int kk = blk * FA_BK + c;
        bool ok = kk < Lk && qi < Lq;
        if (p.causal) { ok = ok && kk <= qi; if (p.window > 0) ok = ok && kk > qi - p.window; }
        else if (p.window > 0) ok = ok && (kk >= qi - p.window) && (kk <= qi + p.window);
        float v = s[c] * p.scale;
        if (p.softcap > 0.f) { const float e2 = __expf(2.f * v * inv_cap); v = p.softcap * (1.f - __fdividef(2.f, e2 + 1.f)); }
        s[c] = ok ? v : -INFINITY;
        bm = fmaxf(bm, s[c]);
      }
      if (!pass2) {
        const float mn = fmaxf(m, bm);
        if (mn > -INFINITY) {
          float acc = 0.f;
#pragma unroll
          for (int c = 0; c < FA_BK; ++c) acc += __expf(s[c] - mn);
          l = l * __expf(m - mn) + acc;
          m = mn;
        }
        if (j == nb - 1) inv_l = l > 0.f ? 1.f / l : 0.f;
      } else {
        // P row -> bf16 -> swizzled K-major smem (chunk c16 of row r lives at ((c16 ^ (r & 7)) * 16)
        unsigned char* prow = S.p + row * 128;
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16) {
          uint32_t w[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float a = (m > -INFINITY) ? __expf(s[c16 * 8 + 2 * t] - m) * inv_l : 0.f;
            const float b = (m > -INFINITY) ? __expf(s[c16 * 8 + 2 * t + 1] - m) * inv_l : 0.f;
            __nv_bfloat162 pk = __floats2bfloat162_rn(a, b);
            w[t] = *reinterpret_cast<uint32_t*>(&pk);
          }
          *reinterpret_cast<uint4*>(prow + ((c16 ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> tcgen05 (async proxy) reads
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&S.sm_done)) : "memory");
    }
    // epilogue: O (already normalised) -> bf16 -> out[token][h*D + d]
    mbar_wait(&S.o_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    bf16* orow = p.out + (size_t)(q_beg + min(qi, Lq - 1)) * p.Hq * D + (size_t)h * D;
#pragma unroll 1
    for (int c = 0; c < D; c += 16) {
      float v[16];
      tmem_ld16(tmem + lane_addr + (uint32_t)(O_COL + c), v);          // .sync.aligned: executed by the whole warp
      if (qi < Lq) {
        uint32_t w[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) { __nv_bfloat162 pk = __floats2bfloat162_rn(v[2 * t], v[2 * t + 1]); w[t] = *reinterpret_cast<uint32_t*>(&pk); }
        reinterpret_cast<uint4*>(orow + c)[0] = make_uint4(w[0], w[1], w[2], w[3]);
        reinterpret_cast<uint4*>(orow + c)[1] = make_uint4(w[4], w[5], w[6], w[7]);
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

// v [T, C] -> vt [C, ldt] (bf16), 32x32 shared-memory tiles; request `z` lands at column vt_seg_off[z] (a multiple of 8,
// so every TMA box of V^T starts 16-byte aligned) and is zero-padded up to the next request.
__global__ void transpose_v_kernel(const bf16* __restrict__ v, bf16* __restrict__ vt, const int* __restrict__ k_seg_off,
                                   const int* __restrict__ vt_seg_off, int C, int ldt) {
  __shared__ bf16 tile[32][33];
  const int seg = blockIdx.z;
  const int t_beg = k_seg_off[seg], L = k_seg_off[seg + 1] - t_beg;
  const int o_beg = vt_seg_off[seg], o_len = vt_seg_off[seg + 1] - o_beg;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  if (t0 >= o_len) return;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (t < L && c < C) ? v[(size_t)(t_beg + t) * C + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    if (c < C && t < o_len && o_beg + t < ldt) vt[(size_t)c * ldt + o_beg + t] = tile[threadIdx.x][i];
  }
}

template <int D>
cudaError_t launch_fa(const AttnPrefillArgs& a, const bf16* vt, const int* vt_seg_off, int ldt, int Tk, int n_seg, int max_lq, cudaStream_t st) {
  CUtensorMap mq, mk, mv;
  if (!make_map_2d(&mq, a.q, a.Tq, (uint64_t)a.Hq * D, (uint64_t)a.Hq * D, FA_BQ) ||
      !make_map_2d(&mk, a.k, Tk, (uint64_t)a.Hkv * D, (uint64_t)a.Hkv * D, FA_BK) ||
      !make_map_2d(&mv, vt, (uint64_t)a.Hkv * D, ldt, ldt, D))
    return cudaErrorNotSupported;
  auto kern = attn_prefill_tc_kernel<D>;
  const size_t smem = sizeof(FaSmem<D>) + 1024;
  static PerDeviceFlag attr_set;
  if (!attr_set.here()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set.here() = 1;
  }
  FaParams p{a.q_seg_off, a.k_seg_off, vt_seg_off, a.Hq, a.Hkv, a.causal, a.window, a.scale, a.softcap, a.out};
  dim3 grid((max_lq + FA_BQ - 1) / FA_BQ, a.Hq, n_seg);
  kern<<<grid, FA_THREADS, smem, st>>>(mq, mk, mv, p);
  return cudaGetLastError();
}

}  // namespace

bool attn_prefill_tc_supported(int D) { return D == 64 || D == 128 || D == 256; }

cudaError_t launch_transpose_v(const bf16* v, bf16* vt, const int* k_seg_off, const int* vt_seg_off, int n_seg, int max_lk,
                               int C, int ldt, cudaStream_t st) {
  if (n_seg <= 0 || max_lk <= 0) return cudaSuccess;
  dim3 grid((max_lk + 8 + 31) / 32, (C + 31) / 32, n_seg), block(32, 8);
  transpose_v_kernel<<<grid, block, 0, st>>>(v, vt, k_seg_off, vt_seg_off, C, ldt);
  return cudaGetLastError();
}

cudaError_t launch_attn_prefill_tc(const AttnPrefillArgs& a, const bf16* vt, const int* vt_seg_off, int ldt, int Tk, int n_seg,
                                   int max_lq, cudaStream_t st) {
  if (a.Tq <= 0 || n_seg <= 0) return cudaSuccess;
  switch (a.D) {
    case 64: return launch_fa<64>(a, vt, vt_seg_off, ldt, Tk, n_seg, max_lq, st);
    case 128: return launch_fa<128>(a, vt, vt_seg_off, ldt, Tk, n_seg, max_lq, st);
    case 256: return launch_fa<256>(a, vt, vt_seg_off, ldt, Tk, n_seg, max_lq, st);
    default: return cudaErrorNotSupported;
  }
}
