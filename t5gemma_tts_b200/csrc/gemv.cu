// Small-batch decode projections: bandwidth-bound GEMV family (bf16 weights streamed once with
// 128-bit loads, fp32 activations/accumulators), with the decoder layer's glue fused in:
//   prologues: embedding gather*sqrt(d) | post-norm + residual + pre-norm (RMSNorm sandwich,
//              models/t5gemma.py:205-243; HF:modeling_t5gemma.py:66-74) | split-KV attention combine
//   epilogues: store | GeGLU (HF:92-96) | bias + exact GELU | bias (predict_layer, models/t5gemma.py:397-406)
// One warp owns one output row (or one gate/up row pair); rows are dealt round-robin to CTAs so every
// SM streams the same number of bytes.  The first weight batch is issued BEFORE griddepcontrol.wait, so
// under programmatic dependent launch the HBM stream of kernel n+1 overlaps the tail of kernel n.
#include "kernels.h"

namespace {

constexpr int GEMV_THREADS = 512;
constexpr int GEMV_U = 4;          // 16-byte loads in flight per lane per batch
constexpr int NORM_MAXPER = 8;

template <int NB>
struct XSmem {
  // x for NB rows, split into lo/hi 16-byte halves of every 8-element chunk so that a warp's
  // LDS.128 is conflict-free: lo[b][chunk] , hi[b][chunk]
  float4* lo; float4* hi; int nchunks;
  __device__ XSmem(float* base, int K) : nchunks(K >> 3) {
    lo = reinterpret_cast<float4*>(base);
    hi = lo + NB * nchunks;
  }
  __device__ __forceinline__ void store(int b, int k, float v) {
    int c = k >> 3, j = k & 7;
    float* p = reinterpret_cast<float*>((j < 4 ? lo : hi) + b * nchunks + c) + (j & 3);
    *p = v;
  }
};

template <int NB>
__device__ __forceinline__ void fma_chunk(const uint4& w, const XSmem<NB>& xs, int c, float* acc) {
  float wf[8];
  bf16x8_to_f32(w, wf);
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    float4 a = xs.lo[b * xs.nchunks + c], h = xs.hi[b * xs.nchunks + c];
    acc[b] = fmaf(wf[0], a.x, acc[b]); acc[b] = fmaf(wf[1], a.y, acc[b]);
    acc[b] = fmaf(wf[2], a.z, acc[b]); acc[b] = fmaf(wf[3], a.w, acc[b]);
    acc[b] = fmaf(wf[4], h.x, acc[b]); acc[b] = fmaf(wf[5], h.y, acc[b]);
    acc[b] = fmaf(wf[6], h.z, acc[b]); acc[b] = fmaf(wf[7], h.w, acc[b]);
  }
}

template <int NB, int P, int E>
__global__ void __launch_bounds__(GEMV_THREADS, 1) gemv_kernel(GemvArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[32];
  XSmem<NB> xs(smem, a.K);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = GEMV_THREADS / 32;
  const int K = a.K, nchunks = K >> 3;
  constexpr int RPU = (E == E_GEGLU) ? 2 : 1;              // rows per unit
  const int n_units = a.N / RPU;
  // unit u -> CTA u % grid, warp (u / grid) % nwarps
  const int first_unit = blockIdx.x + warp * gridDim.x;
  const int unit_stride = gridDim.x * nwarps;

  // ---- weight prefetch that does not depend on the previous kernel --------------------------
  uint4 wpre[RPU][GEMV_U];
  {
    const int u = first_unit;
#pragma unroll
    for (int r = 0; r < RPU; ++r)
#pragma unroll
      for (int i = 0; i < GEMV_U; ++i) {
        int c = lane + i * 32;
        wpre[r][i] = (u < n_units && c < nchunks)
                         ? ldg_stream(a.W + (size_t)(u * RPU + r) * K + (size_t)c * 8) : make_uint4(0, 0, 0, 0);
      }
  }
  pdl_launch_dependents();
  pdl_wait();

  // ---- early exit when no row of this batch is generating -----------------------------------
  if (a.slots) {
    int any = 0;
    for (int b = 0; b < a.B; ++b) any |= a.slots[a.slot0 + b].active;
    if (!any) return;
  }

  // ---- prologue: build x[NB][K] in shared memory --------------------------------------------
  for (int b = 0; b < NB; ++b) {
    const bool valid = b < a.B;
    if (P == P_PLAIN) {
      for (int k = threadIdx.x; k < K; k += GEMV_THREADS) xs.store(b, k, valid ? a.x[(size_t)b * K + k] : 0.f);
    } else if (P == P_COMBINE) {
      // x[head*D + d] = sum_s exp(m_s-M) o_s[d] / sum_s exp(m_s-M) l_s  (split-KV attention merge)
      const int D = a.head_dim, NS = a.n_splits, H = K / D;
      for (int k = threadIdx.x; k < K; k += GEMV_THREADS) {
        float v = 0.f;
        if (valid) {
          int hd = k / D, d = k - hd * D;
          const float* ml = a.part_ml + ((size_t)(b * H + hd) * NS) * 2;
          const float* po = a.part_o + ((size_t)(b * H + hd) * NS) * D + d;
          float M = -INFINITY;
          for (int s = 0; s < NS; ++s) M = fmaxf(M, ml[2 * s]);
          float num = 0.f, den = 0.f;
          for (int s = 0; s < NS; ++s) {
            float m = ml[2 * s];
            if (m == -INFINITY) continue;
            float wgt = __expf(m - M);
            num = fmaf(wgt, po[(size_t)s * D], num);
            den = fmaf(wgt, ml[2 * s + 1], den);
          }
          v = den > 0.f ? num / den : 0.f;
        }
        xs.store(b, k, v);
      }
    } else {
      // P_NORM / P_RES_NORM / P_EMBED_NORM: RMSNorm sandwich in fp32
      constexpr int per = NORM_MAXPER;                              // K <= 512*8 for norm prologues (host-checked)
      float hreg[per];
      float ss = 0.f;
      if (P == P_RES_NORM) {
        float ys = 0.f;
        _Pragma("unroll") for (int i = 0; i < per; ++i) {
          int k = threadIdx.x + i * GEMV_THREADS;
          float yv = (valid && k < K) ? a.y[(size_t)b * K + k] : 0.f;
          ys = fmaf(yv, yv, ys);
        }
        ys = block_sum(ys, red);
        const float rinv = rsqrtf(ys / (float)K + a.eps);
        _Pragma("unroll") for (int i = 0; i < per; ++i) {
          int k = threadIdx.x + i * GEMV_THREADS;
          float hv = 0.f;
          if (valid && k < K) {
            hv = a.h_in[(size_t)b * K + k] + a.y[(size_t)b * K + k] * rinv * a.g_post[k];
            if (a.h_out && blockIdx.x == 0) a.h_out[(size_t)b * K + k] = hv;
          }
          hreg[i] = hv;
          ss = fmaf(hv, hv, ss);
        }
      } else if (P == P_EMBED_NORM) {
        const int tok = valid ? a.slots[a.slot0 + b].last_token : 0;
        _Pragma("unroll") for (int i = 0; i < per; ++i) {
          int k = threadIdx.x + i * GEMV_THREADS;
          float hv = 0.f;
          if (valid && k < K) {
            hv = __bfloat162float(a.emb[(size_t)tok * K + k]) * a.emb_scale;
            if (a.h_out && blockIdx.x == 0) a.h_out[(size_t)b * K + k] = hv;
          }
          hreg[i] = hv;
          ss = fmaf(hv, hv, ss);
        }
      } else {  // P_NORM
        _Pragma("unroll") for (int i = 0; i < per; ++i) {
          int k = threadIdx.x + i * GEMV_THREADS;
          float hv = (valid && k < K) ? a.h_in[(size_t)b * K + k] : 0.f;
          hreg[i] = hv;
          ss = fmaf(hv, hv, ss);
        }
      }
      ss = block_sum(ss, red);
      const float rinv = rsqrtf(ss / (float)K + a.eps);
      _Pragma("unroll") for (int i = 0; i < per; ++i) {
        int k = threadIdx.x + i * GEMV_THREADS;
        if (k < K) xs.store(b, k, hreg[i] * rinv * a.g_pre[k]);
      }
    }
  }
  __syncthreads();

  // ---- stream the rows ------------------------------------------------------------------------
  bool first = true;
  for (int u = first_unit; u < n_units; u += unit_stride) {
    float acc[RPU][NB];
#pragma unroll
    for (int r = 0; r < RPU; ++r)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[r][b] = 0.f;
    const bf16* wrow = a.W + (size_t)u * RPU * K;
    for (int c0 = 0; c0 < nchunks; c0 += 32 * GEMV_U) {
      uint4 w[RPU][GEMV_U];
      if (first && c0 == 0) {
#pragma unroll
        for (int r = 0; r < RPU; ++r)
#pragma unroll
          for (int i = 0; i < GEMV_U; ++i) w[r][i] = wpre[r][i];
      } else {
#pragma unroll
        for (int r = 0; r < RPU; ++r)
#pragma unroll
          for (int i = 0; i < GEMV_U; ++i) {
            int c = c0 + lane + i * 32;
            w[r][i] = (c < nchunks) ? ldg_stream(wrow + (size_t)r * K + (size_t)c * 8) : make_uint4(0, 0, 0, 0);
          }
      }
#pragma unroll
      for (int i = 0; i < GEMV_U; ++i) {
        int c = c0 + lane + i * 32;
        if (c < nchunks) {
#pragma unroll
          for (int r = 0; r < RPU; ++r) fma_chunk<NB>(w[r][i], xs, c, acc[r]);
        }
      }
    }
    first = false;
#pragma unroll
    for (int r = 0; r < RPU; ++r)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[r][b] = warp_sum(acc[r][b]);
    if (lane == 0) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        if (b >= a.B) break;
        if (E == E_GEGLU) {
          a.out[(size_t)b * a.out_stride + u] = gelu_tanh_f(acc[0][b]) * acc[RPU - 1][b];
        } else if (E == E_BIAS_GELU) {
          a.out[(size_t)b * a.out_stride + u] = gelu_erf_f(acc[0][b] + a.bias[u]);
        } else if (E == E_BIAS) {
          a.out[(size_t)b * a.out_stride + u] = acc[0][b] + a.bias[u];
        } else {
          a.out[(size_t)b * a.out_stride + u] = acc[0][b];
        }
      }
    }
  }
}

template <int NB, int P, int E>
cudaError_t launch_one(const GemvArgs& a, int grid, cudaStream_t st, bool pdl) {
  size_t smem = (size_t)NB * a.K * sizeof(float);
  auto kern = gemv_kernel<NB, P, E>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMV_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, a);
}

template <int NB>
cudaError_t launch_nb(const GemvArgs& a, int P, int E, int grid, cudaStream_t st, bool pdl) {
#define CASE(PP, EE) if (P == PP && E == EE) return launch_one<NB, PP, EE>(a, grid, st, pdl)
  CASE(P_PLAIN, E_STORE);
  CASE(P_COMBINE, E_STORE);
  CASE(P_NORM, E_STORE);
  CASE(P_RES_NORM, E_STORE);
  CASE(P_EMBED_NORM, E_STORE);
  CASE(P_RES_NORM, E_GEGLU);
  CASE(P_NORM, E_GEGLU);
  CASE(P_RES_NORM, E_BIAS_GELU);
  CASE(P_NORM, E_BIAS_GELU);
  CASE(P_PLAIN, E_BIAS);
#undef CASE
  return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t launch_gemv(const GemvArgs& a, int P, int E, int num_sms, cudaStream_t st, bool pdl) {
  if (a.K % 8 != 0) return cudaErrorInvalidValue;
  if (P != P_PLAIN && P != P_COMBINE && a.K > GEMV_THREADS * NORM_MAXPER) return cudaErrorInvalidValue;
  int NB = a.B <= 1 ? 1 : (a.B <= 2 ? 2 : 4);
  if (a.B > 4) return cudaErrorInvalidValue;
  if ((size_t)NB * a.K * 4 > 200 * 1024) return cudaErrorInvalidValue;
  int grid = num_sms;
  switch (NB) {
    case 1: return launch_nb<1>(a, P, E, grid, st, pdl);
    case 2: return launch_nb<2>(a, P, E, grid, st, pdl);
    default: return launch_nb<4>(a, P, E, grid, st, pdl);
  }
}
