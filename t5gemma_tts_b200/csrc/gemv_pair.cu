// Two dependent small projections of the bs<=4 decode step in ONE kernel:
//   phase A:  y   = W1 . x                                   (self-attention o_proj,  plain store)
//   -- grid-wide barrier --
//   phase B:  out = W2 . rmsnorm_pre(h + rmsnorm_post(y))     (cross-attention q_proj, RMSNorm sandwich prologue)
// One CTA per SM, 16 warps; every warp owns one row of W1 (loaded before griddepcontrol.wait) and one row of W2
// (requested right after it, in flight during phase A), so after the producer (the attention kernel) finishes, the two projections cost two L2
// round trips, one block reduction and one grid barrier instead of two kernels with their drain / launch / ramp
// (o_proj 2.6+0.5 us and q_cross 4.5+0.65 us as separate kernels).  Same arithmetic, in the same order, as
// gemv_kernel<NB, P_PLAIN, E_STORE> followed by gemv_kernel<NB, P_RES_NORM, E_STORE> (gemv.cu).
// The barrier is a ticket counter in global memory: the CTAs of launch n draw tickets [n*G, (n+1)*G) and wait for
// the counter to reach (n+1)*G; all CTAs are co-resident (grid <= number of SMs, one CTA per SM), and the dependent
// kernel of the PDL chain cannot take their place because it is only scheduled after every CTA of this grid has
// started.
#include "kernels.h"

namespace {

constexpr int GP_THREADS = 512, GP_WARPS = 16, GP_U = 9, GP_NP = 5;    // rows of up to 2304 elements

template <int NB>
struct GpX {      // activation vector(s) in shared memory, split lo/hi halves of every 8-element chunk (gemv.cu XSmem)
  float4* lo; float4* hi; int nchunks;
  __device__ GpX(float* base, int K) : nchunks(K >> 3) { lo = reinterpret_cast<float4*>(base); hi = lo + NB * nchunks; }
  __device__ __forceinline__ void store(int b, int k, float v) {
    const int c = k >> 3, j = k & 7;
    reinterpret_cast<float*>((j < 4 ? lo : hi) + b * nchunks + c)[j & 3] = v;
  }
};

template <int NB>
__device__ __forceinline__ void gp_dot(const uint4 (&w)[GP_U], const GpX<NB>& xs, int lane, float* acc) {
#pragma unroll
  for (int b = 0; b < NB; ++b) acc[b] = 0.f;
#pragma unroll
  for (int i = 0; i < GP_U; ++i) {
    const int c = lane + 32 * i;
    if (c < xs.nchunks) {
      float wf[8];
      bf16x8_to_f32(w[i], wf);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float4 a = xs.lo[b * xs.nchunks + c], h = xs.hi[b * xs.nchunks + c];
        acc[b] = fmaf(wf[0], a.x, acc[b]); acc[b] = fmaf(wf[1], a.y, acc[b]);
        acc[b] = fmaf(wf[2], a.z, acc[b]); acc[b] = fmaf(wf[3], a.w, acc[b]);
        acc[b] = fmaf(wf[4], h.x, acc[b]); acc[b] = fmaf(wf[5], h.y, acc[b]);
        acc[b] = fmaf(wf[6], h.z, acc[b]); acc[b] = fmaf(wf[7], h.w, acc[b]);
      }
    }
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) acc[b] = warp_sum(acc[b]);
}

__device__ __forceinline__ void gp_load_row(uint4 (&w)[GP_U], const bf16* row, int nchunks, int lane, uint64_t pol) {
#pragma unroll
  for (int i = 0; i < GP_U; ++i) {
    const int c = lane + 32 * i;
    w[i] = (c < nchunks) ? ldg_stream(row + (size_t)c * 8, pol) : make_uint4(0, 0, 0, 0);
  }
}

template <int NB>
__global__ void __launch_bounds__(GP_THREADS, 1) gemv_pair_kernel(GemvPairArgs a) {
  extern __shared__ __align__(16) float gp_smem[];
  __shared__ float red[128];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K1 = a.K1, K2 = a.K2;
  const int total_warps = gridDim.x * GP_WARPS, gw = blockIdx.x * GP_WARPS + warp;

  // ---- before the dependency resolves: both weight rows of this warp, the norm gains, slot activity ----
  const uint64_t pol = l2_evict_first_policy();
  uint4 w1[GP_U], w2[GP_U];
  if (gw < a.N1) gp_load_row(w1, a.W1 + (size_t)gw * K1, K1 >> 3, lane, pol);
  float gpre[GP_NP], gpost[GP_NP];
#pragma unroll
  for (int i = 0; i < GP_NP; ++i) {
    const int k = tid + i * GP_THREADS;
    gpre[i] = (k < K2) ? a.g_pre[k] : 0.f;
    gpost[i] = (k < K2) ? a.g_post[k] : 0.f;
  }
  int any = 1;
  if (a.slots) { any = 0; for (int b = 0; b < a.B; ++b) any |= a.slots[b].active; }
  pdl_launch_dependents();
  pdl_wait();
  trace_begin(a.trace);
  if (!any) { trace_end(a.trace); return; }                // uniform over the grid: nobody reaches the barrier
  // phase B's weights are requested now (not before the wait: 19 MB of early traffic slowed the attention kernel this
  // grid overlaps with); they have phase A and the barrier to arrive
  if (gw < a.N2) gp_load_row(w2, a.W2 + (size_t)gw * K2, K2 >> 3, lane, pol);

  // ---- phase A: y = W1 . x ----
  {
    GpX<NB> xs(gp_smem, K1);
    const int n4 = K1 >> 2;
    for (int b = 0; b < NB; ++b) {
      const float4* xv = reinterpret_cast<const float4*>(a.x + (size_t)b * K1);
      for (int base = 0; base < n4; base += GP_THREADS * 2) {
        float4 r[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int i = base + tid + u * GP_THREADS;
          r[u] = (b < a.B && i < n4) ? __ldcg(xv + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int i = base + tid + u * GP_THREADS;
          if (i < n4) ((i & 1) ? xs.hi : xs.lo)[b * xs.nchunks + (i >> 1)] = r[u];
        }
      }
    }
    __syncthreads();
    for (int row = gw; row < a.N1; row += total_warps) {
      if (row != gw) gp_load_row(w1, a.W1 + (size_t)row * K1, K1 >> 3, lane, pol);
      float acc[NB];
      gp_dot<NB>(w1, xs, lane, acc);
      if (lane == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b) if (b < a.B) a.y[(size_t)b * a.N1 + row] = acc[b];
      }
    }
  }
  // the residual stream does not depend on phase A: request it before the barrier
  float hpre[NB][GP_NP];
#pragma unroll
  for (int b = 0; b < NB; ++b)
#pragma unroll
    for (int i = 0; i < GP_NP; ++i) {
      const int k = tid + i * GP_THREADS;
      hpre[b][i] = (b < a.B && k < K2) ? __ldcg(a.h_in + (size_t)b * K2 + k) : 0.f;
    }
  // ---- grid-wide barrier (ticket counter).  The launch is cooperative, so the driver guarantees that all gridDim.x CTAs
  //      are co-resident (the launch fails otherwise).  The wait is still bounded by time: after one second of
  //      %globaltimer the step raises the slot's error flag and skips phase B instead of publishing a partial result ----
  __shared__ int gp_timed_out;
  __syncthreads();                                         // this CTA's y stores are issued; xs is free
  if (tid == 0) {
    __threadfence();                                       // ... and visible before the arrival
    const unsigned long long ticket = atomicAdd(a.barrier, 1ULL);
    const unsigned long long target = (ticket / gridDim.x + 1ULL) * gridDim.x;
    unsigned spins = 0;
    unsigned long long t0 = 0;
    int bad = 0;
    while (*reinterpret_cast<volatile unsigned long long*>(a.barrier) < target) {
      if ((++spins & 0x3ff) == 0) {
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 1000000000ULL) { if (a.err_slots) atomicOr(&a.err_slots[0].error, 4); bad = 1; break; }
      }
    }
    gp_timed_out = bad;
    __threadfence();
  }
  __syncthreads();
  if (gp_timed_out) { trace_end(a.trace); return; }

  // ---- phase B: RMSNorm sandwich (h = h_in + rmsnorm(y) g_post ; x = rmsnorm(h) g_pre), then out = W2 . x ----
  {
    GpX<NB> xs(gp_smem, K2);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const bool valid = b < a.B;
      float hreg[GP_NP], yg[GP_NP];
      float s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
#pragma unroll
      for (int i = 0; i < GP_NP; ++i) {
        const int k = tid + i * GP_THREADS;
        const bool ok = valid && k < K2;
        const float yv = ok ? __ldcg(a.y + (size_t)b * K2 + k) : 0.f;
        hreg[i] = hpre[b][i];
        yg[i] = yv * gpost[i];
        s1 = fmaf(yv, yv, s1); s2 = fmaf(hreg[i], hreg[i], s2);
        s3 = fmaf(hreg[i], yg[i], s3); s4 = fmaf(yg[i], yg[i], s4);
      }
      block_sum4(s1, s2, s3, s4, red);
      const float ry = rsqrtf(s1 / (float)K2 + a.eps);
#pragma unroll
      for (int i = 0; i < GP_NP; ++i) {
        const int k = tid + i * GP_THREADS;
        if (valid && k < K2) {
          hreg[i] = fmaf(yg[i], ry, hreg[i]);
          if (a.h_out && blockIdx.x == 0) a.h_out[(size_t)b * K2 + k] = hreg[i];
        }
      }
      const float ss = s2 + 2.f * ry * s3 + ry * ry * s4;
      const float rinv = rsqrtf(ss / (float)K2 + a.eps);
#pragma unroll
      for (int i = 0; i < GP_NP; ++i) {
        const int k = tid + i * GP_THREADS;
        if (k < K2) xs.store(b, k, hreg[i] * rinv * gpre[i]);
      }
    }
    __syncthreads();
    for (int row = gw; row < a.N2; row += total_warps) {
      if (row != gw) gp_load_row(w2, a.W2 + (size_t)row * K2, K2 >> 3, lane, pol);
      float acc[NB];
      gp_dot<NB>(w2, xs, lane, acc);
      if (lane == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b) if (b < a.B) a.out[(size_t)b * a.out_stride + row] = acc[b];
      }
    }
  }
  trace_end(a.trace);
}

template <int NB>
cudaError_t launch_gp(const GemvPairArgs& a, int num_sms, cudaStream_t st, bool pdl) {
  auto kern = gemv_pair_kernel<NB>;
  const size_t smem = (size_t)NB * (a.K1 > a.K2 ? a.K1 : a.K2) * sizeof(float);
  static PerDeviceFlag attr_set;
  if (smem > attr_set.here()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set.here() = smem;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(num_sms);
  cfg.blockDim = dim3(GP_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeCooperative;             // co-residency of the grid is requested, not assumed
  attr[0].val.cooperative = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, a);
}

}  // namespace

bool gemv_pair_supported(const GemvPairArgs& a) {
  return a.B >= 1 && a.B <= 4 && a.K1 % 8 == 0 && a.K2 % 8 == 0 && a.K1 <= GP_U * 256 && a.K2 <= GP_U * 256 &&
         a.K2 <= GP_THREADS * GP_NP && a.N1 == a.K2 && a.barrier != nullptr;
}

cudaError_t launch_gemv_pair(const GemvPairArgs& a, int num_sms, cudaStream_t st, bool pdl) {
  if (!gemv_pair_supported(a)) return cudaErrorInvalidValue;
  if (a.B <= 1) return launch_gp<1>(a, num_sms, st, pdl);
  if (a.B <= 2) return launch_gp<2>(a, num_sms, st, pdl);
  return launch_gp<4>(a, num_sms, st, pdl);
}
