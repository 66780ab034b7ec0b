// Small-batch decode projections: the bandwidth-bound GEMV family of the bs=1..4 decode step.
//
// Design (B200; measurements in profiles/r1_gemv_design_experiments.md): bf16 weights are streamed once with
// 128-bit `ld.global.nc.L1::no_allocate` loads, 9 independent loads in flight per lane and 32 warps per SM
// (~147 KB in flight per SM, 96 % of the measured copy bandwidth on the vocabulary projection).  The work
// unit is one 2304-element K-segment of one weight row, so every warp can issue ALL loads of a unit at once:
// K=9216 rows are split into 4 segments (split-K inside the CTA, combined through shared memory), and the
// gate/up rows of a GeGLU pair are two units of the same output.  The first unit of every warp is loaded
// BEFORE griddepcontrol.wait: under programmatic dependent launch the HBM stream of kernel n+1 starts while
// kernel n is still finishing.  Activations stay fp32 (shared memory, conflict-free split layout).
// Fused glue:
//   prologues: embedding gather*sqrt(d) | post-norm + residual + pre-norm (RMSNorm sandwich,
//              models/t5gemma.py:205-243; HF:modeling_t5gemma.py:66-74)
//   epilogues: store | GeGLU (HF:92-96) | bias + exact GELU | bias (predict_layer, models/t5gemma.py:397-406)
#include "kernels.h"

namespace {

constexpr int GV_THREADS = 512;
constexpr int GV_WARPS = GV_THREADS / 32;
constexpr int GV_U = 9;                        // 16-byte loads in flight per lane per unit
constexpr int SEG_CHUNKS = 32 * GV_U;          // 288 chunks = 2304 elements per K-segment
constexpr int NORM_MAXPER = 8;                 // norm prologues: K <= 512*8
constexpr int MAX_PARTS_PER_CTA = 2048;        // shared-memory partials

template <int NB>
struct XSmem {
  // x for NB rows, split into lo/hi 16-byte halves of every 8-element chunk so that a warp's LDS.128 is
  // conflict-free: lo[b][chunk], hi[b][chunk]
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

// unit -> weights: local unit lu of this CTA = (local output ol, part p); p = rowsel*KSEG + seg
struct UnitMap {
  int K, nchunks, kseg, parts, n_out, unit_rows;
  __device__ __forceinline__ const bf16* row_ptr(const bf16* W, int out, int p, int& seg) const {
    const int rowsel = p / kseg;
    seg = p - rowsel * kseg;
    return W + (size_t)(out * unit_rows + rowsel) * K;
  }
};

template <int NB, int P, int E, int NP>
__global__ void __launch_bounds__(GV_THREADS, (NB == 1) ? 2 : 1) gemv_kernel(GemvArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = a.K, nchunks = K >> 3;
  XSmem<NB> xs(smem, K);
  float* part = smem + (size_t)NB * K;                     // [outs_per_cta][parts][NB]
  UnitMap um;
  um.K = K; um.nchunks = nchunks; um.kseg = (nchunks + SEG_CHUNKS - 1) / SEG_CHUNKS;
  um.unit_rows = (E == E_GEGLU) ? 2 : 1;
  um.parts = um.kseg * um.unit_rows;
  um.n_out = a.N / um.unit_rows;
  // outputs are dealt round-robin to CTAs: local output ol <-> out = blockIdx.x + ol*gridDim.x
  const int outs_here = (um.n_out > (int)blockIdx.x) ? (um.n_out - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int units_here = outs_here * um.parts;

  // ---- first unit of this warp: issue its weight loads before the dependency resolves --------------
  const uint64_t pol = l2_evict_first_policy();
  uint4 w[GV_U];
  int lu = warp;
  {
    if (lu < units_here) {
      const int ol = lu / um.parts, p = lu - ol * um.parts;
      int seg;
      const bf16* wr = um.row_ptr(a.W, blockIdx.x + ol * gridDim.x, p, seg);
#pragma unroll
      for (int i = 0; i < GV_U; ++i) {
        const int c = seg * SEG_CHUNKS + lane + 32 * i;
        w[i] = (c < nchunks) ? ldg_stream(wr + (size_t)c * 8, pol) : make_uint4(0, 0, 0, 0);
      }
    }
  }
  // RMSNorm gains are weights: fetch them before the dependency resolves as well
  float gpre[NP], gpost[NP];
  if (P == P_NORM || P == P_RES_NORM || P == P_EMBED_NORM) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int k = threadIdx.x + i * GV_THREADS;
      gpre[i] = (k < K) ? a.g_pre[k] : 0.f;
      gpost[i] = (P == P_RES_NORM && k < K) ? a.g_post[k] : 0.f;
    }
  }
  // slot activity was settled by the sampler, which completed before any kernel that can overlap this one was
  // launched (engine.cu: the first kernel after the sampler has no programmatic overlap): read it before the wait
  int any = 1;
  if (a.slots) {
    any = 0;
    for (int b = 0; b < a.B; ++b) any |= a.slots[a.slot0 + b].active;
  }
  pdl_launch_dependents();
  pdl_wait();
  trace_begin(a.trace);
  if (!any) return;

  // ---- prologue: build x[NB][K] in shared memory --------------------------------------------------------
  for (int b = 0; b < NB; ++b) {
    const bool valid = b < a.B;
    if (P == P_PLAIN) {
      // 16-byte loads, a batch of them in registers before the first shared-memory store (a load->store loop
      // would pay one L2 round trip per iteration); float4 #i is the lo (i even) / hi (i odd) half of chunk i/2
      const float4* xv = reinterpret_cast<const float4*>(a.x + (size_t)b * K);
      const int n4 = K >> 2;
      constexpr int XB = 3;
      for (int base = 0; base < n4; base += GV_THREADS * XB) {
        float4 r[XB];
#pragma unroll
        for (int u = 0; u < XB; ++u) {
          const int i = base + threadIdx.x + u * GV_THREADS;
          r[u] = (valid && i < n4) ? __ldcg(xv + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < XB; ++u) {
          const int i = base + threadIdx.x + u * GV_THREADS;
          if (i < n4) ((i & 1) ? xs.hi : xs.lo)[b * xs.nchunks + (i >> 1)] = r[u];
        }
      }
    } else {
      // RMSNorm sandwich in fp32: h = h_in + rmsnorm(y)*g_post ; x = rmsnorm(h)*g_pre
      constexpr int per = NP;
      float hreg[per];
      float ss = 0.f;
      bool have_ss = false;
      if (P == P_RES_NORM) {
        // one reduction pass: with r = rsqrt(mean(y^2)+eps),  sum(h + y r g)^2 = S2 + 2 r S3 + r^2 S4
        float yg[per];
        float s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
        _Pragma("unroll") for (int i = 0; i < per; ++i) {
          const int k = threadIdx.x + i * GV_THREADS;
          const bool ok = valid && k < K;
          const float yv = ok ? a.y[(size_t)b * K + k] : 0.f;
          hreg[i] = ok ? a.h_in[(size_t)b * K + k] : 0.f;
          yg[i] = yv * gpost[i];
          s1 = fmaf(yv, yv, s1); s2 = fmaf(hreg[i], hreg[i], s2);
          s3 = fmaf(hreg[i], yg[i], s3); s4 = fmaf(yg[i], yg[i], s4);
        }
        block_sum4(s1, s2, s3, s4, red);
        const float rinv = rsqrtf(s1 / (float)K + a.eps);
        _Pragma("unroll") for (int i = 0; i < per; ++i) {
          const int k = threadIdx.x + i * GV_THREADS;
          if (valid && k < K) {
            hreg[i] = fmaf(yg[i], rinv, hreg[i]);
            if (a.h_out && blockIdx.x == 0) a.h_out[(size_t)b * K + k] = hreg[i];
          }
        }
        ss = s2 + 2.f * rinv * s3 + rinv * rinv * s4;
        have_ss = true;
      } else if (P == P_EMBED_NORM) {
        const int tok = valid ? a.slots[a.slot0 + b].last_token : 0;
        _Pragma("unroll") for (int i = 0; i < per; ++i) {
          const int k = threadIdx.x + i * GV_THREADS;
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
          const int k = threadIdx.x + i * GV_THREADS;
          const float hv = (valid && k < K) ? a.h_in[(size_t)b * K + k] : 0.f;
          hreg[i] = hv;
          ss = fmaf(hv, hv, ss);
        }
      }
      if (!have_ss) ss = block_sum(ss, red);
      const float rinv = rsqrtf(ss / (float)K + a.eps);
      _Pragma("unroll") for (int i = 0; i < per; ++i) {
        const int k = threadIdx.x + i * GV_THREADS;
        if (k < K) xs.store(b, k, hreg[i] * rinv * gpre[i]);
      }
    }
  }
  __syncthreads();

  // ---- stream this CTA's units (first one is already in registers) -----------------------------------------
  const bool direct = (um.parts == 1);
  for (; lu < units_here; lu += GV_WARPS) {
    const int ol = lu / um.parts, p = lu - ol * um.parts;
    const int out = blockIdx.x + ol * gridDim.x;
    int seg;
    const bf16* wr = um.row_ptr(a.W, out, p, seg);
    if (lu != warp) {      // whole unit issued back to back: 9 x 512 B contiguous per warp keeps DRAM pages open
#pragma unroll
      for (int i = 0; i < GV_U; ++i) {
        const int c = seg * SEG_CHUNKS + lane + 32 * i;
        w[i] = (c < nchunks) ? ldg_stream(wr + (size_t)c * 8, pol) : make_uint4(0, 0, 0, 0);
      }
    }
    float acc[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = 0.f;
#pragma unroll
    for (int i = 0; i < GV_U; ++i) {
      const int c = seg * SEG_CHUNKS + lane + 32 * i;
      if (c < nchunks) fma_chunk<NB>(w[i], xs, c, acc);
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = warp_sum(acc[b]);
    if (lane == 0) {
      if (direct) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          if (b >= a.B) break;
          float o = acc[b];
          if (E == E_BIAS_GELU) o = gelu_erf_f(o + a.bias[out]);
          else if (E == E_BIAS) o = o + a.bias[out];
          a.out[(size_t)b * a.out_stride + out] = o;
        }
      } else {
#pragma unroll
        for (int b = 0; b < NB; ++b) part[(size_t)lu * NB + b] = acc[b];
      }
    }
  }
  if (direct) { trace_end(a.trace); return; }
  __syncthreads();
  // ---- combine the K-segments (and the gate/up pair) of every output of this CTA ---------------------------
  for (int i = threadIdx.x; i < outs_here * NB; i += GV_THREADS) {
    const int ol = i / NB, b = i - ol * NB;
    if (b >= a.B) continue;
    const int out = blockIdx.x + ol * gridDim.x;
    float v0 = 0.f, v1 = 0.f;
    for (int s = 0; s < um.kseg; ++s) {
      v0 += part[(size_t)(ol * um.parts + s) * NB + b];
      if (E == E_GEGLU) v1 += part[(size_t)(ol * um.parts + um.kseg + s) * NB + b];
    }
    float o;
    if (E == E_GEGLU) o = gelu_tanh_f(v0) * v1;
    else if (E == E_BIAS_GELU) o = gelu_erf_f(v0 + a.bias[out]);
    else if (E == E_BIAS) o = v0 + a.bias[out];
    else o = v0;
    a.out[(size_t)b * a.out_stride + out] = o;
  }
  trace_end(a.trace);
}

template <int NB, int P, int E, int NP>
cudaError_t launch_one(const GemvArgs& a, int grid, size_t smem, cudaStream_t st, bool pdl) {
  auto kern = gemv_kernel<NB, P, E, NP>;
  static PerDeviceFlag attr_set;
  if (!attr_set.here()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr_set.here() = 1;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GV_THREADS);
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
cudaError_t launch_nb(const GemvArgs& a, int P, int E, int grid, size_t smem, cudaStream_t st, bool pdl) {
  const bool small = a.K <= GV_THREADS * 5;
#define CASE(PP, EE) if (P == PP && E == EE) return launch_one<NB, PP, EE, 1>(a, grid, smem, st, pdl)
#define CASEN(PP, EE) if (P == PP && E == EE) return small ? launch_one<NB, PP, EE, 5>(a, grid, smem, st, pdl) \
                                                            : launch_one<NB, PP, EE, NORM_MAXPER>(a, grid, smem, st, pdl)
  CASE(P_PLAIN, E_STORE);
  CASEN(P_NORM, E_STORE);
  CASEN(P_RES_NORM, E_STORE);
  CASEN(P_EMBED_NORM, E_STORE);
  CASEN(P_RES_NORM, E_GEGLU);
  CASEN(P_RES_NORM, E_BIAS_GELU);
  CASE(P_PLAIN, E_BIAS);
#undef CASE
#undef CASEN
  return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t launch_gemv(const GemvArgs& a, int P, int E, int num_sms, cudaStream_t st, bool pdl) {
  if (a.K % 8 != 0 || a.B < 1 || a.B > 4) return cudaErrorInvalidValue;
  if (P != P_PLAIN && a.K > GV_THREADS * NORM_MAXPER) return cudaErrorInvalidValue;
  const int NB = a.B <= 1 ? 1 : (a.B <= 2 ? 2 : 4);
  const int unit_rows = (E == E_GEGLU) ? 2 : 1;
  if (a.N % unit_rows) return cudaErrorInvalidValue;
  const int nchunks = a.K / 8, kseg = (nchunks + SEG_CHUNKS - 1) / SEG_CHUNKS, parts = kseg * unit_rows;
  const int n_out = a.N / unit_rows;
  // two CTAs per SM for single-row launches: 147 KB of weight loads in flight per SM (one CTA per SM was measured at
  // 22.4 instead of 15.6 us on the gate|up projection)
  const int grid = (NB == 1 ? 2 : 1) * num_sms;
  const int outs_per_cta = (n_out + grid - 1) / grid;
  const size_t part_floats = (parts == 1) ? 0 : (size_t)outs_per_cta * parts * NB;
  if (part_floats > MAX_PARTS_PER_CTA * 4) return cudaErrorInvalidValue;
  const size_t smem = ((size_t)NB * a.K + part_floats) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  switch (NB) {
    case 1: return launch_nb<1>(a, P, E, grid, smem, st, pdl);
    case 2: return launch_nb<2>(a, P, E, grid, smem, st, pdl);
    default: return launch_nb<4>(a, P, E, grid, smem, st, pdl);
  }
}
