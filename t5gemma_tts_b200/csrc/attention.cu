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
#include <cooperative_groups.h>

namespace {

template <int D> struct Geo {
  static constexpr int LPT = (D / 8 < 16) ? D / 8 : 16; // lanes per token (16 for D=256: 16 dims per lane keeps
                                                        // the q/acc/k/v slices at ~100 registers)
  static constexpr int DPL = D / LPT;                   // dims per lane (multiple of 8)
  static constexpr int TPW = 32 / LPT;                  // tokens per warp iteration
  static constexpr int NV = DPL / 8;                    // 16-byte loads per lane per row
};

constexpr int ATT_WARPS = 4;

// online-softmax state of one token group (replicated over the group's LPT lanes)
template <int G, int DPL> struct GroupState {
  float m[G], l[G], acc[G][DPL];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      m[g] = -INFINITY; l[g] = 0.f;
#pragma unroll
      for (int i = 0; i < DPL; ++i) acc[g][i] = 0.f;
    }
  }
};

template <int G, int D>
__device__ __forceinline__ void group_update(GroupState<G, Geo<D>::DPL>& st, const float (*qreg)[Geo<D>::DPL],
                                             const float* kf, const float* vf, float scale, float softcap,
                                             bool valid) {
  constexpr int LPT = Geo<D>::LPT, DPL = Geo<D>::DPL;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < DPL; ++i) dot = fmaf(qreg[g][i], kf[i], dot);
#pragma unroll
    for (int o = LPT >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    float s = dot * scale;
    if (softcap > 0.f) s = softcap * tanhf(s / softcap);
    if (!valid) s = -INFINITY;
    const float mn = fmaxf(st.m[g], s);
    const float corr = (st.m[g] == -INFINITY) ? 0.f : __expf(st.m[g] - mn);
    const float p = valid ? __expf(s - mn) : 0.f;
    st.l[g] = st.l[g] * corr + p;
#pragma unroll
    for (int i = 0; i < DPL; ++i) st.acc[g][i] = fmaf(p, vf[i], st.acc[g][i] * corr);
    st.m[g] = mn;
  }
}

// merges the TPW token groups of a warp with xor shuffles over the group-index bits of the lane id
template <int G, int D>
__device__ __forceinline__ void warp_merge(GroupState<G, Geo<D>::DPL>& st) {
  constexpr int LPT = Geo<D>::LPT, DPL = Geo<D>::DPL;
#pragma unroll
  for (int o = LPT; o < 32; o <<= 1) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float mo = __shfl_xor_sync(0xffffffffu, st.m[g], o);
      const float lo_ = __shfl_xor_sync(0xffffffffu, st.l[g], o);
      const float mn = fmaxf(st.m[g], mo);
      const float w0 = (st.m[g] == -INFINITY) ? 0.f : __expf(st.m[g] - mn);
      const float w1 = (mo == -INFINITY) ? 0.f : __expf(mo - mn);
      st.l[g] = st.l[g] * w0 + lo_ * w1;
#pragma unroll
      for (int i = 0; i < DPL; ++i) {
        const float ao = __shfl_xor_sync(0xffffffffu, st.acc[g][i], o);
        st.acc[g][i] = st.acc[g][i] * w0 + ao * w1;
      }
      st.m[g] = mn;
    }
  }
}

constexpr int ATD_WARPS = 8;

// grid (Hkv, NS, B), thread-block cluster (1, NS, 1): the NS split CTAs of one (request, kv head) exchange
// their partial softmax states through distributed shared memory instead of a global round trip.
template <int G, int D>
__global__ void __launch_bounds__(ATD_WARPS * 32) attn_decode_kernel(AttnDecodeArgs a) {
  constexpr int LPT = Geo<D>::LPT, DPL = Geo<D>::DPL, TPW = Geo<D>::TPW, NV = Geo<D>::NV;
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float qs[G][D];
  __shared__ float cs[D / 2], sn[D / 2];
  __shared__ float knew[D], vnew[D];
  __shared__ __align__(16) float w_o[ATD_WARPS][G][D];
  __shared__ float w_ml[ATD_WARPS][G][2];
  __shared__ __align__(16) float c_o[G][D];      // this CTA's partial, read by the cluster peers
  __shared__ float c_ml[G][2];

  {
    const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x, n = gridDim.x * gridDim.y * gridDim.z;
    l2_prefetch_range(a.pf[0], cta, n);
    l2_prefetch_range(a.pf[1], cta, n);
  }
  pdl_launch_dependents();
  pdl_wait();
  trace_begin(a.trace);
  const int hk = blockIdx.x, split = blockIdx.y, b = blockIdx.z;
  const int NS = a.n_splits;
  const SlotDev& sl = a.slots[b];
  if (!sl.active) return;                          // uniform over the whole cluster (same b)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = a.is_cross ? sl.n_text : sl.cur_len;
  const int lo = (!a.is_cross && a.window > 0) ? max(0, L - a.window) : 0;
  int chunk = (L - lo + NS - 1) / NS;
  chunk = (chunk + TPW - 1) / TPW * TPW;
  const int t_begin = lo + split * chunk, t_end = min(L, t_begin + chunk);
  const int PT = a.pool.page_tokens;
  const int* bt = a.block_table + (size_t)b * a.bt_stride;
  const bool has_new = (!a.is_cross) && (t_end == L) && (t_end > t_begin);

  // RoPE table of this request's position (fp32 angles from a FLOAT position, HF:150-161)
  if (a.rope_cs) {
    for (int i = tid; i < D / 2; i += blockDim.x) { cs[i] = a.rope_cs[(size_t)b * D + i]; sn[i] = a.rope_cs[(size_t)b * D + D / 2 + i]; }
  } else {
    const float pos = sl.pos;
    for (int i = tid; i < D / 2; i += blockDim.x) {
      float s, c;
      sincosf(pos * a.inv_freq[i], &s, &c);
      cs[i] = c; sn[i] = s;
    }
  }
  // raw q (and the new k/v) are fetched in the same round trip as the table
  for (int i = tid; i < G * D; i += blockDim.x) {
    const int g = i / D, j = i - g * D;
    qs[g][j] = a.q[(size_t)b * a.q_stride + (size_t)(hk * G + g) * D + j];
  }
  if (has_new) {
    const float* kp = a.kv_new + (size_t)b * a.kv_stride + (size_t)hk * D;
    const float* vp = kp + (size_t)a.Hkv * D;
    for (int j = tid; j < D; j += blockDim.x) { knew[j] = kp[j]; vnew[j] = __bfloat162float(__float2bfloat16(vp[j])); }
  }
  __syncthreads();
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
    const int t = L - 1, page = bt[t / PT], off = t % PT;
    bf16* kd = a.pool.ptr(a.layer, 0, page) + ((size_t)hk * PT + off) * D;
    bf16* vd = a.pool.ptr(a.layer, 1, page) + ((size_t)hk * PT + off) * D;
    for (int j = tid; j < D; j += blockDim.x) { kd[j] = __float2bfloat16(knew[j]); vd[j] = __float2bfloat16(vnew[j]); }
  }

  const int grp = lane / LPT, l8 = lane % LPT;
  float qreg[G][DPL];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int i = 0; i < DPL; ++i) qreg[g][i] = qs[g][l8 * DPL + i];

  GroupState<G, DPL> st;
  st.init();
  for (int t0 = t_begin + warp * TPW; t0 < t_end; t0 += ATD_WARPS * TPW) {
    const int t = t0 + grp;
    const bool valid = t < t_end;
    float kf[DPL], vf[DPL];
    if (valid) {
      if (has_new && t == L - 1) {
#pragma unroll
        for (int i = 0; i < DPL; ++i) { kf[i] = knew[l8 * DPL + i]; vf[i] = vnew[l8 * DPL + i]; }
      } else {
        const int page = bt[t / PT], off = t % PT;
        const bf16* kp = a.pool.ptr(a.layer, 0, page) + ((size_t)hk * PT + off) * D + l8 * DPL;
        const bf16* vp = a.pool.ptr(a.layer, 1, page) + ((size_t)hk * PT + off) * D + l8 * DPL;
        uint4 ku[NV], vu[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) { ku[i] = *reinterpret_cast<const uint4*>(kp + i * 8); vu[i] = *reinterpret_cast<const uint4*>(vp + i * 8); }
#pragma unroll
        for (int i = 0; i < NV; ++i) { bf16x8_to_f32(ku[i], kf + i * 8); bf16x8_to_f32(vu[i], vf + i * 8); }
      }
    } else {
#pragma unroll
      for (int i = 0; i < DPL; ++i) { kf[i] = 0.f; vf[i] = 0.f; }
    }
    // all lanes execute the shuffles; invalid groups contribute nothing
    group_update<G, D>(st, qreg, kf, vf, a.scale, a.softcap, valid);
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
  // CTA-level merge of the warps -> c_o / c_ml (unnormalised, relative to the CTA max)
  for (int i = tid; i < G * D; i += blockDim.x) {
    const int g = i / D, d = i - g * D;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < ATD_WARPS; ++w) M = fmaxf(M, w_ml[w][g][0]);
    float num = 0.f, den = 0.f;
    if (M > -INFINITY) {
#pragma unroll
      for (int w = 0; w < ATD_WARPS; ++w) {
        const float m = w_ml[w][g][0];
        const float wt = (m == -INFINITY) ? 0.f : __expf(m - M);
        num = fmaf(wt, w_o[w][g][d], num);
        den = fmaf(wt, w_ml[w][g][1], den);
      }
    }
    c_o[g][d] = num;
    if (d == 0) { c_ml[g][0] = M; c_ml[g][1] = den; }
  }
  cluster.sync();
  // ---- distributed final merge over the cluster: rank `split` finalises dims [split*D/NS, (split+1)*D/NS) ----
  const int dslice = D / NS;
  for (int i = tid; i < G * dslice; i += blockDim.x) {
    const int g = i / dslice, d = split * dslice + (i - g * dslice);
    float M = -INFINITY;
    for (int r = 0; r < NS; ++r) M = fmaxf(M, cluster.map_shared_rank(&c_ml[0][0], r)[g * 2]);
    float num = 0.f, den = 0.f;
    for (int r = 0; r < NS; ++r) {
      const float* rml = cluster.map_shared_rank(&c_ml[0][0], r);
      const float m = rml[g * 2];
      if (m == -INFINITY) continue;
      const float wt = __expf(m - M);
      num = fmaf(wt, cluster.map_shared_rank(&c_o[0][0], r)[g * D + d], num);
      den = fmaf(wt, rml[g * 2 + 1], den);
    }
    const float o = den > 0.f ? num / den : 0.f;
    if (a.out) a.out[(size_t)b * a.Hq * D + (size_t)(hk * G + g) * D + d] = o;
    if (a.out_bf) a.out_bf[(size_t)b * a.Hq * D + (size_t)(hk * G + g) * D + d] = __float2bfloat16(o);
  }
  cluster.sync();                                  // peers may still be reading this CTA's shared memory
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
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(a.Hkv, a.n_splits, a.B);
  cfg.blockDim = dim3(ATD_WARPS * 32);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = a.n_splits; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, attn_decode_kernel<G, D>, a);
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
  if (a.n_splits < 1 || a.n_splits > 8 || (a.n_splits & (a.n_splits - 1)) || a.D % a.n_splits) return cudaErrorInvalidValue;
  DISPATCH_GD(G, a.D, launch_decode_gd, a, st, pdl);
  return cudaErrorInvalidValue;
}

cudaError_t launch_attn_prefill(const AttnPrefillArgs& a, cudaStream_t st) {
  const int G = a.Hq / a.Hkv;
  DISPATCH_GD(G, a.D, launch_prefill_gd, a, st);
  return cudaErrorInvalidValue;
}
