// GQA attention kernels (CUDA cores; the work is bandwidth/latency-bound at q_len=1).
//   * attn_decode_kernel: one new query per request over the PAGED KV pool, split-KV ("flash decoding")
//     with fp32 online softmax, Gemma-2 attn-logit softcap, causal + sliding window (self) or dense
//     non-causal (cross, over the per-utterance encoder K/V pages).  Fuses: PM-RoPE of q (and of the new
//     k) at the request's fractional progress position, and the in-place KV append (K stored post-RoPE,
//     like HF DynamicLayer.update).  Replaces HF:modeling_t5gemma.py:209-240,274-314 and
//     models/t5gemma.py:85-172 for q_len == 1.
//   * attn_prefill_kernel: varlen-packed prefill attention (bidirectional encoder / causal decoder /
//     cross) over contiguous K/V.
// The G = Hq/Hkv query heads of a KV group are processed together so K/V bytes are read once.
#include "kernels.h"
#include "attn_common.cuh"
#include <cooperative_groups.h>

namespace {

constexpr int ATD_MAX_NS = 16;       // 16-CTA clusters need cudaFuncAttributeNonPortableClusterSizeAllowed
constexpr int ATD_BT_CACHE = 256;      // block-table entries staged in shared memory (covers 4096 tokens at 16/page)

// grid (Hkv, NS, B), thread-block cluster (1, NS, 1): the NS split CTAs of one (request, kv head) exchange
// their partial softmax states through distributed shared memory instead of a global round trip: every CTA
// PUSHES the slice of its partial that rank r will finalise into rank r's receive buffer, one cluster barrier,
// then every rank merges locally.
template <int G, int D, int ATD_WARPS>
__global__ void __launch_bounds__(ATD_WARPS * 32) attn_decode_kernel(AttnDecodeArgs a) {
  constexpr int LPT = Geo<D>::LPT, DPL = Geo<D>::DPL, TPW = Geo<D>::TPW, NV = Geo<D>::NV;
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float qs[G][D];
  __shared__ float cs[D / 2], sn[D / 2];
  __shared__ float knew[D], vnew[D];
  __shared__ int bt_s[ATD_BT_CACHE];
  __shared__ __align__(16) float w_o[ATD_WARPS][G][D];
  __shared__ float w_ml[ATD_WARPS][G][2];
  __shared__ __align__(16) float recv_o[G * D];                  // [src rank][g][dslice = D / NS]
  __shared__ float recv_ml[ATD_MAX_NS][G][2];
  __shared__ float w_wt[ATD_WARPS][G], c_ml[G][2], f_wt[ATD_MAX_NS][G];

  pdl_launch_dependents();
  const int hk = blockIdx.x, split = blockIdx.y, b = blockIdx.z;
  const int NS = a.n_splits;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = lane / LPT, l8 = lane % LPT;
  const int PT = a.pool.page_tokens;
  const int* bt = a.block_table + (size_t)b * a.bt_stride;
  constexpr int NT = ATD_WARPS * 32;
  constexpr int QPT = (G * D + NT - 1) / NT, KPT = (D + NT - 1) / NT, RPT = (D / 2 + NT - 1) / NT;
  // ---- before the PDL dependency resolves: everything that only depends on EARLIER steps.  The slot state and
  //      the RoPE table were written by this step's sampler, which is complete because the first kernel after it
  //      is launched without programmatic overlap (engine.cu); the block table is host-written before the launch
  //      and cached K/V rows of earlier tokens are immutable.  Only q and the new k/v come from the producer. ----
  const SlotDev& sl = a.slots[b];
  const int active = sl.active;
  const int L = a.is_cross ? sl.n_text : sl.cur_len;
  const float pos = sl.pos;
  {
    float rc[RPT], rs[RPT];
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
      const int i = tid + u * NT;
      const bool ok = a.rope_cs && i < D / 2;
      rc[u] = ok ? a.rope_cs[(size_t)b * D + i] : 1.f;
      rs[u] = ok ? a.rope_cs[(size_t)b * D + D / 2 + i] : 0.f;
    }
    const int bt_v = (tid < ATD_BT_CACHE && tid < a.bt_stride) ? bt[tid] : 0;
#pragma unroll
    for (int u = 0; u < RPT; ++u) { const int i = tid + u * NT; if (i < D / 2) { cs[i] = rc[u]; sn[i] = rs[u]; } }
    if (tid < ATD_BT_CACHE) bt_s[tid] = bt_v;
    if (!a.rope_cs) {
      for (int i = tid; i < D / 2; i += blockDim.x) {
        float s_, c_;
        sincosf(pos * a.inv_freq[i], &s_, &c_);
        cs[i] = c_; sn[i] = s_;
      }
    }
  }
  const int lo = (!a.is_cross && a.window > 0) ? max(0, L - a.window) : 0;
  int chunk = (L - lo + NS - 1) / NS;
  chunk = (chunk + TPW - 1) / TPW * TPW;
  const int t_begin = lo + split * chunk, t_end = min(L, t_begin + chunk);
  const bool has_new = (!a.is_cross) && (t_end == L) && (t_end > t_begin);
  __syncthreads();
  auto page_of = [&](int t) -> int { const int pi = t / PT; return pi < ATD_BT_CACHE ? bt_s[pi] : bt[pi]; };
  // K/V rows of the first two token groups of this warp: in flight while the producer kernel is still running
  // two register buffers of two token groups each: the rows of iteration i+1 are in flight while iteration i is consumed
  // (with one buffer every iteration after the first paid a full memory round trip: 5.8 us at ctx 160, 10.3 us at ctx 720)
  uint4 ku[2][NV], vu[2][NV], ku2[2][NV], vu2[2][NV];
  auto load_groups_into = [&](uint4 (&kd)[2][NV], uint4 (&vd)[2][NV], int t0) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int t = t0 + u * ATD_WARPS * TPW + grp;
      if (t < t_end && !(has_new && t == L - 1)) {
        const int page = page_of(t), off = t % PT;
        const bf16* kp = a.pool.ptr(a.layer, 0, page) + ((size_t)hk * PT + off) * D + l8 * DPL;
        const bf16* vp = a.pool.ptr(a.layer, 1, page) + ((size_t)hk * PT + off) * D + l8 * DPL;
#pragma unroll
        for (int i = 0; i < NV; ++i) { kd[u][i] = *reinterpret_cast<const uint4*>(kp + i * 8); vd[u][i] = *reinterpret_cast<const uint4*>(vp + i * 8); }
      }
    }
  };
  auto load_groups = [&](int t0) { load_groups_into(ku, vu, t0); };
  const int t_first = t_begin + warp * TPW;
  constexpr int STEP = 2 * ATD_WARPS * TPW;               // tokens of the CTA's range covered per iteration
  bool pre2 = false;
  if (active && a.preload) {
    load_groups(t_first);
    if (t_first + STEP < t_end) { load_groups_into(ku2, vu2, t_first + STEP); pre2 = true; }
  }

  pdl_wait();
  trace_begin(a.trace);
  if (active && !a.preload) load_groups(t_first);
  // ---- the producer's outputs: raw q and the new k/v, loaded into registers before the first shared-memory store ----
  float rq[QPT], rk[KPT], rv[KPT];
#pragma unroll
  for (int u = 0; u < QPT; ++u) {
    const int i = tid + u * NT;
    rq[u] = (i < G * D) ? __ldcg(a.q + (size_t)b * a.q_stride + (size_t)(hk * G) * D + i) : 0.f;
  }
#pragma unroll
  for (int u = 0; u < KPT; ++u) {
    const int j = tid + u * NT;
    const bool ok = !a.is_cross && j < D;
    rk[u] = ok ? __ldcg(a.kv_new + (size_t)b * a.kv_stride + (size_t)hk * D + j) : 0.f;
    rv[u] = ok ? __ldcg(a.kv_new + (size_t)b * a.kv_stride + (size_t)(a.Hkv + hk) * D + j) : 0.f;
  }
  if (!active) return;                             // uniform over the whole cluster (same b)
#pragma unroll
  for (int u = 0; u < QPT; ++u) { const int i = tid + u * NT; if (i < G * D) qs[i / D][i % D] = rq[u]; }
#pragma unroll
  for (int u = 0; u < KPT; ++u) {
    const int j = tid + u * NT;
    if (j < D) { knew[j] = rk[u]; vnew[j] = __bfloat162float(__float2bfloat16(rv[u])); }
  }
  __syncthreads();
  cluster.barrier_arrive();                        // "this CTA is running": waited on before the first remote store
  // rotate in place: element pairs (j, j+D/2)
  for (int i = tid; i < G * D / 2; i += blockDim.x) {
    const int g = i / (D / 2), j = i - g * (D / 2);
    const float x1 = qs[g][j], x2 = qs[g][j + D / 2];
    qs[g][j] = x1 * cs[j] - x2 * sn[j];
    qs[g][j + D / 2] = x2 * cs[j] + x1 * sn[j];
  }
  if (has_new) {
    for (int j = tid; j < D / 2; j += blockDim.x) {
      const float x1 = knew[j], x2 = knew[j + D / 2];
      knew[j] = __bfloat162float(__float2bfloat16(x1 * cs[j] - x2 * sn[j]));
      knew[j + D / 2] = __bfloat162float(__float2bfloat16(x2 * cs[j] + x1 * sn[j]));
    }
  }
  __syncthreads();
  if (has_new) {   // append to the page (K post-RoPE), visible to later steps
    const int t = L - 1, page = page_of(t), off = t % PT;
    bf16* kd = a.pool.ptr(a.layer, 0, page) + ((size_t)hk * PT + off) * D;
    bf16* vd = a.pool.ptr(a.layer, 1, page) + ((size_t)hk * PT + off) * D;
    for (int j = tid; j < D; j += blockDim.x) { kd[j] = __float2bfloat16(knew[j]); vd[j] = __float2bfloat16(vnew[j]); }
  }

  float qreg[G][DPL];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int i = 0; i < DPL; ++i) qreg[g][i] = qs[g][l8 * DPL + i];

  GroupState<G, DPL> st;
  st.init();
  // two token groups per iteration: both K/V row loads are in flight before either is consumed
  auto consume = [&](const uint4 (&kd)[2][NV], const uint4 (&vd)[2][NV], int t0) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (t0 + u * ATD_WARPS * TPW >= t_end) break;      // warp-uniform
      const int t = t0 + u * ATD_WARPS * TPW + grp;
      const bool valid = t < t_end, fresh = valid && has_new && t == L - 1;
      float kf[DPL], vf[DPL];
      if (fresh) {
#pragma unroll
        for (int i = 0; i < DPL; ++i) { kf[i] = knew[l8 * DPL + i]; vf[i] = vnew[l8 * DPL + i]; }
      } else if (valid) {
#pragma unroll
        for (int i = 0; i < NV; ++i) { bf16x8_to_f32(kd[u][i], kf + i * 8); bf16x8_to_f32(vd[u][i], vf + i * 8); }
      } else {
#pragma unroll
        for (int i = 0; i < DPL; ++i) { kf[i] = 0.f; vf[i] = 0.f; }
      }
      // all lanes execute the shuffles; invalid groups contribute nothing
      group_update<G, D>(st, qreg, kf, vf, a.scale, a.softcap, valid);
    }
  };
  for (int t0 = t_first; t0 < t_end; t0 += 2 * STEP) {
    if (t0 + STEP < t_end && !(pre2 && t0 == t_first)) load_groups_into(ku2, vu2, t0 + STEP);
    consume(ku, vu, t0);
    if (t0 + STEP >= t_end) break;
    if (t0 + 2 * STEP < t_end) load_groups_into(ku, vu, t0 + 2 * STEP);
    consume(ku2, vu2, t0 + STEP);
  }
  warp_merge<G, D>(st);
  if (grp == 0) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (l8 == 0) { w_ml[warp][g][0] = st.m[g]; w_ml[warp][g][1] = st.l[g]; }
#pragma unroll
      for (int i = 0; i < DPL; ++i) w_o[warp][g][l8 * DPL + i] = st.acc[g][i];
    }
  }
  __syncthreads();
  // CTA-level merge of the warps, pushed straight into the receive buffer of the rank that owns each dim slice.
  // Per-warp merge weights are computed once (G*WARPS exps) instead of once per output element.
  const int dslice = D / NS;
  if (tid < G) {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < ATD_WARPS; ++w) M = fmaxf(M, w_ml[w][tid][0]);
    float den = 0.f;
#pragma unroll
    for (int w = 0; w < ATD_WARPS; ++w) {
      const float m = w_ml[w][tid][0];
      const float wt = (m == -INFINITY) ? 0.f : __expf(m - M);
      den = fmaf(wt, w_ml[w][tid][1], den);
      w_wt[w][tid] = wt;
    }
    c_ml[tid][0] = M; c_ml[tid][1] = den;
  }
  __syncthreads();
  cluster.barrier_wait();                                // every peer has started: its shared memory may be written
  for (int i = tid; i < G * D; i += blockDim.x) {
    const int g = i / D, d = i - g * D;
    float num = 0.f;
#pragma unroll
    for (int w = 0; w < ATD_WARPS; ++w) num = fmaf(w_wt[w][g], w_o[w][g][d], num);
    const int dst = d / dslice;                          // rank that finalises this dim
    float* ro = cluster.map_shared_rank(&recv_o[0], dst);
    ro[((size_t)split * G + g) * dslice + (d - dst * dslice)] = num;
  }
  if (tid < G * NS) {                                    // (m, l) of this CTA to every rank
    const int g = tid % G, dst = tid / G;
    float* rm = cluster.map_shared_rank(&recv_ml[0][0][0], dst);
    rm[(split * G + g) * 2] = c_ml[g][0]; rm[(split * G + g) * 2 + 1] = c_ml[g][1];
  }
  cluster.sync();                                        // all pushes have landed
  // ---- local final merge: rank `split` owns dims [split*dslice, (split+1)*dslice) of every head of the group ----
  if (tid < G) {
    float M = -INFINITY;
    for (int r = 0; r < NS; ++r) M = fmaxf(M, recv_ml[r][tid][0]);
    float den = 0.f;
    for (int r = 0; r < NS; ++r) {
      const float m = recv_ml[r][tid][0];
      const float wt = (m == -INFINITY) ? 0.f : __expf(m - M);
      den = fmaf(wt, recv_ml[r][tid][1], den);
      f_wt[r][tid] = wt;
    }
    const float inv = den > 0.f ? 1.f / den : 0.f;
    for (int r = 0; r < NS; ++r) f_wt[r][tid] *= inv;
  }
  __syncthreads();
  for (int i = tid; i < G * dslice; i += blockDim.x) {
    const int g = i / dslice, dd = i - g * dslice;
    float o = 0.f;
    for (int r = 0; r < NS; ++r) o = fmaf(f_wt[r][g], recv_o[(r * G + g) * dslice + dd], o);
    const int d = split * dslice + dd;
    if (a.out) a.out[(size_t)b * a.Hq * D + (size_t)(hk * G + g) * D + d] = o;
    if (a.out_bf) a.out_bf[(size_t)b * a.Hq * D + (size_t)(hk * G + g) * D + d] = __float2bfloat16(o);
  }
  trace_end(a.trace);
}

// ---- prefill: one warp per (query token, kv head) ----------------------------------------------
template <int G, int D>
__global__ void __launch_bounds__(ATT_WARPS * 32) attn_prefill_kernel(AttnPrefillArgs a) {
  constexpr int LPT = Geo<D>::LPT, DPL = Geo<D>::DPL, TPW = Geo<D>::TPW, NV = Geo<D>::NV;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tq = blockIdx.x * ATT_WARPS + warp, hk = blockIdx.y;
  if (tq >= a.Tq) return;
  const int seg = a.q_seg_of[tq];
  const int qi = tq - a.q_seg_off[seg];                    // index inside the request
  const int k0 = a.k_seg_off[seg], Lk = a.k_seg_off[seg + 1] - k0;
  int lo = 0, hi = Lk;
  if (a.causal) { hi = min(Lk, qi + 1); if (a.window > 0) lo = max(0, qi - a.window + 1); }
  else if (a.window > 0) { lo = max(0, qi - a.window); hi = min(Lk, qi + a.window + 1); }
  const int grp = lane / LPT, l8 = lane % LPT;
  float qreg[G][DPL];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const bf16* qp = a.q + (size_t)tq * a.Hq * D + (size_t)(hk * G + g) * D + l8 * DPL;
#pragma unroll
    for (int i = 0; i < NV; ++i) bf16x8_to_f32(*reinterpret_cast<const uint4*>(qp + i * 8), &qreg[g][i * 8]);
  }
  GroupState<G, DPL> st;
  st.init();
  const size_t ldk = (size_t)a.Hkv * D;
  for (int t0 = lo; t0 < hi; t0 += TPW) {
    const int t = t0 + grp;
    const bool valid = t < hi;
    float kf[DPL], vf[DPL];
    if (valid) {
      const bf16* kp = a.k + (size_t)(k0 + t) * ldk + (size_t)hk * D + l8 * DPL;
      const bf16* vp = a.v + (size_t)(k0 + t) * ldk + (size_t)hk * D + l8 * DPL;
      uint4 ku[NV], vu[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) { ku[i] = *reinterpret_cast<const uint4*>(kp + i * 8); vu[i] = *reinterpret_cast<const uint4*>(vp + i * 8); }
#pragma unroll
      for (int i = 0; i < NV; ++i) { bf16x8_to_f32(ku[i], kf + i * 8); bf16x8_to_f32(vu[i], vf + i * 8); }
    } else {
#pragma unroll
      for (int i = 0; i < DPL; ++i) { kf[i] = 0.f; vf[i] = 0.f; }
    }
    group_update<G, D>(st, qreg, kf, vf, a.scale, a.softcap, valid);
  }
  warp_merge<G, D>(st);
  if (grp == 0) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float inv = st.l[g] > 0.f ? 1.f / st.l[g] : 0.f;
      bf16* op = a.out + (size_t)tq * a.Hq * D + (size_t)(hk * G + g) * D + l8 * DPL;
#pragma unroll
      for (int i = 0; i < DPL; ++i) op[i] = __float2bfloat16(st.acc[g][i] * inv);
    }
  }
}

template <int G, int D>
cudaError_t launch_decode_gd(const AttnDecodeArgs& a, cudaStream_t st, bool pdl) {
  // 8 warps per CTA for latency (bs<=4: few CTAs, long token loops); 4 warps for batches so that 3 CTAs fit per SM
  const bool wide = a.B <= 4;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(a.Hkv, a.n_splits, a.B);
  cfg.blockDim = dim3((wide ? 8 : 4) * 32);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = a.n_splits; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  if (a.n_splits > 8) {                                  // non-portable cluster size
    static PerDeviceFlag np_set;
    if (!np_set.here()) {
      cudaError_t e = wide ? cudaFuncSetAttribute(attn_decode_kernel<G, D, 8>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)
                           : cudaFuncSetAttribute(attn_decode_kernel<G, D, 4>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (e != cudaSuccess) return e;
      np_set.here() = 1;
    }
  }
  if (wide) return cudaLaunchKernelEx(&cfg, attn_decode_kernel<G, D, 8>, a);
  return cudaLaunchKernelEx(&cfg, attn_decode_kernel<G, D, 4>, a);
}

template <int G, int D>
cudaError_t launch_prefill_gd(const AttnPrefillArgs& a, cudaStream_t st) {
  dim3 grid((a.Tq + ATT_WARPS - 1) / ATT_WARPS, a.Hkv);
  attn_prefill_kernel<G, D><<<grid, ATT_WARPS * 32, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace

#define DISPATCH_GD(G, D, FN, ...)                                                     \
  do {                                                                                 \
    if (G == 1) {                                                                      \
      if (D == 16) return FN<1, 16>(__VA_ARGS__); if (D == 32) return FN<1, 32>(__VA_ARGS__);   \
      if (D == 64) return FN<1, 64>(__VA_ARGS__); if (D == 128) return FN<1, 128>(__VA_ARGS__); \
      if (D == 256) return FN<1, 256>(__VA_ARGS__);                                    \
    } else if (G == 2) {                                                               \
      if (D == 16) return FN<2, 16>(__VA_ARGS__); if (D == 32) return FN<2, 32>(__VA_ARGS__);   \
      if (D == 64) return FN<2, 64>(__VA_ARGS__); if (D == 128) return FN<2, 128>(__VA_ARGS__); \
      if (D == 256) return FN<2, 256>(__VA_ARGS__);                                    \
    } else if (G == 4) {                                                               \
      if (D == 16) return FN<4, 16>(__VA_ARGS__); if (D == 32) return FN<4, 32>(__VA_ARGS__);   \
      if (D == 64) return FN<4, 64>(__VA_ARGS__); if (D == 128) return FN<4, 128>(__VA_ARGS__); \
    }                                                                                  \
  } while (0)

cudaError_t launch_attn_decode(const AttnDecodeArgs& a, cudaStream_t st, bool pdl) {
  const int G = a.Hq / a.Hkv;
  if (a.n_splits < 1 || a.n_splits > ATD_MAX_NS || (a.n_splits & (a.n_splits - 1)) || a.D % a.n_splits) return cudaErrorInvalidValue;
  DISPATCH_GD(G, a.D, launch_decode_gd, a, st, pdl);
  return cudaErrorInvalidValue;
}

cudaError_t launch_attn_prefill(const AttnPrefillArgs& a, cudaStream_t st) {
  const int G = a.Hq / a.Hkv;
  DISPATCH_GD(G, a.D, launch_prefill_gd, a, st);
  return cudaErrorInvalidValue;
}
