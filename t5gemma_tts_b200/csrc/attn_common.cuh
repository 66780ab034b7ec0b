// Shared pieces of the CUDA-core attention kernels (attention.cu): lane geometry of a K/V row,
// online-softmax state of a token group and its update / warp-level merge (HF:modeling_t5gemma.py:209-240 semantics:
// scores * scale, optional softcap*tanh(s/softcap), fp32 softmax).
#pragma once
#include "common.cuh"

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
  const float inv_softcap = softcap > 0.f ? 1.f / softcap : 0.f;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < DPL; ++i) dot = fmaf(qreg[g][i], kf[i], dot);
#pragma unroll
    for (int o = LPT >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    float s = dot * scale;
    if (softcap > 0.f) {                     // softcap*tanh(s/softcap), tanh(x) = 1 - 2/(exp(2x)+1)
      const float e2 = __expf(2.f * s * inv_softcap);
      s = softcap * (1.f - __fdividef(2.f, e2 + 1.f));
    }
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

