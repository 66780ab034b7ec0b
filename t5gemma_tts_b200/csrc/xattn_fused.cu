// Cross-attention fused into its output projection for the bs<=4 decode step (short texts).
//
// Replaces two dependent kernels of the step (attn_decode_kernel over the encoder K/V + the o_proj GEMV,
// models/t5gemma.py:85-172 at q_len == 1) by one: the grid is made of thread-block clusters of Hq CTAs; CTA
// rank h of EVERY cluster computes head h of the cross-attention (redundantly across clusters: the encoder K/V
// of a short text is a few hundred KB and L2-resident), pushes its D outputs into the activation buffer of all
// CTAs of its cluster through distributed shared memory, and after one cluster barrier every CTA runs its slice
// of the o_proj GEMV with the weights it already holds in registers.
// Everything that does not depend on the producer kernel (the q projection) is issued BEFORE
// griddepcontrol.wait: o_proj weights (registers), the head's K/V rows (cp.async into shared memory), slot
// state, block table and the RoPE table.  After the wait only q is loaded.
#include "kernels.h"
#include <cooperative_groups.h>

namespace {

constexpr int XF_THREADS = 1024;               // 32 warps: 9 clusters x 8 CTAs x 32 warps = one o_proj row per warp (2b-2b),
                                               // one CTA per SM, every cluster resident in the first wave (a GPC hosts two)
constexpr int XF_WARPS = XF_THREADS / 32;
constexpr int XF_U = 9;                        // 16-byte weight loads per lane: rows of up to 2304 elements
constexpr int XF_MAX_PAGES = 64;               // block-table entries cached per row

__device__ __forceinline__ void cp_async16_xf(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

struct XfSmem {      // carve-up of the dynamic shared memory (all offsets 16-byte aligned)
  float4* x_lo; float4* x_hi;     // [NB][K/8] split activation layout of the GEMV (gemv.cu XSmem)
  bf16* kbuf; bf16* vbuf;         // [cap][D]
  float* sc;                      // [cap] scores -> probabilities of the concatenated rows
  float* part;                    // [nparts][NB][D] P.V partials
  float* q_s;                     // [NB][D] rotated queries
  float* ml;                      // [NB] 1/sum
  int* bt_s;                      // [NB][XF_MAX_PAGES]
};

template <int NB>
__global__ void __launch_bounds__(XF_THREADS, 1) xattn_oproj_kernel(XAttnOprojArgs a) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char xf_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = a.D, K = a.K, nchunks = K >> 3, cap = a.max_keys;
  const int h = (int)cluster.block_rank(), hk = h / (a.Hq / a.Hkv);
  XfSmem S;
  {
    unsigned char* p = xf_raw;
    S.x_lo = reinterpret_cast<float4*>(p); S.x_hi = S.x_lo + NB * nchunks; p += (size_t)NB * K * 4;
    S.kbuf = reinterpret_cast<bf16*>(p); p += (size_t)cap * D * 2;
    S.vbuf = reinterpret_cast<bf16*>(p); p += (size_t)cap * D * 2;
    S.sc = reinterpret_cast<float*>(p); p += (size_t)((cap + 3) & ~3) * 4;
    S.part = reinterpret_cast<float*>(p); p += (size_t)(XF_THREADS / (D / 2)) * NB * D * 4;
    S.q_s = reinterpret_cast<float*>(p); p += (size_t)NB * D * 4;
    S.ml = reinterpret_cast<float*>(p); p += 16;
    S.bt_s = reinterpret_cast<int*>(p);
  }

  // ---------------- before the dependency resolves ----------------
  // (1) o_proj weights: row `gw` of this warp (rows are dealt round-robin over all warps of the grid)
  const uint64_t pol = l2_evict_first_policy();
  const int total_warps = gridDim.x * XF_WARPS;
  const int gw = blockIdx.x * XF_WARPS + warp;
  uint4 w[XF_U];
  if (gw < a.N) {
    const bf16* wr = a.W + (size_t)gw * K;
#pragma unroll
    for (int i = 0; i < XF_U; ++i) {
      const int c = lane + 32 * i;
      w[i] = (c < nchunks) ? ldg_stream(wr + (size_t)c * 8, pol) : make_uint4(0, 0, 0, 0);
    }
  }
  // (2) slot state (settled by the sampler, see engine.cu), key counts of the rows, block tables
  int nk[NB], koff[NB + 1];
  int any = 0;
  koff[0] = 0;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    int n = 0;
    if (b < a.B) { const SlotDev& sl = a.slots[a.slot0 + b]; if (sl.active) { n = sl.n_text; any = 1; } }
    if (koff[b] + n > cap) n = max(0, cap - koff[b]);          // never overrun shared memory (host guarantees the fit)
    nk[b] = n; koff[b + 1] = koff[b] + n;
  }
  const int PT = a.pool.page_tokens;
  for (int i = tid; i < NB * XF_MAX_PAGES; i += XF_THREADS) {
    const int b = i / XF_MAX_PAGES, pi = i - b * XF_MAX_PAGES;
    // unconditional (no dependency on the slot loads above: both are in flight together); entries past the row's text are unused
    S.bt_s[i] = (b < a.B && pi < a.bt_stride) ? a.block_table[(size_t)(a.slot0 + b) * a.bt_stride + pi] : 0;
  }
  // (3) RoPE table entries of the (row, pair) this thread will rotate
  const int half = D >> 1;
  float rc[NB], rs[NB];           // thread t rotates pair j = t % half of rows b = t / half + u * (XF_THREADS / half)
  const int rows_per_pass = XF_THREADS / half;
#pragma unroll
  for (int u = 0; u < NB; ++u) {
    const int b = tid / half + u * rows_per_pass, j = tid % half;
    const bool ok = b < NB && b < a.B;
    rc[u] = ok ? a.rope_cs[(size_t)(a.slot0 + b) * D + j] : 1.f;
    rs[u] = ok ? a.rope_cs[(size_t)(a.slot0 + b) * D + half + j] : 0.f;
  }
  pdl_launch_dependents();
  __syncthreads();                                   // block-table cache visible
  // (4) K/V rows of head hk for every live row: cp.async straight into shared memory
  {
    const int cpr = D >> 3;                          // 16-byte chunks per K/V row
    const int total = koff[NB] * cpr;
    for (int c = tid; c < total; c += XF_THREADS) {
      const int tok = c / cpr, col = c - tok * cpr;
      int b = 0;
#pragma unroll
      for (int bb = 1; bb < NB; ++bb) if (tok >= koff[bb]) b = bb;
      const int t = tok - koff[b];
      const int page = S.bt_s[b * XF_MAX_PAGES + t / PT], off = t % PT;
      const size_t src = ((size_t)hk * PT + off) * D + col * 8;
      cp_async16_xf(S.kbuf + (size_t)tok * D + col * 8, a.pool.ptr(a.layer, 0, page) + src);
      cp_async16_xf(S.vbuf + (size_t)tok * D + col * 8, a.pool.ptr(a.layer, 1, page) + src);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  pdl_wait();
  trace_begin(a.trace);
  if (!any) return;                                  // uniform over the grid
  cluster.barrier_arrive();                          // "this CTA is running": waited on before the first remote store

  // ---------------- q of head h: load, rotate (PM-RoPE at the row's progress position) ----------------
  {
    float x1[NB], x2[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int b = tid / half + u * rows_per_pass, j = tid % half;
      const bool ok = b < NB && b < a.B;
      const float* qp = a.q + (size_t)b * a.q_stride + (size_t)h * D;
      x1[u] = ok ? __ldcg(qp + j) : 0.f;
      x2[u] = ok ? __ldcg(qp + j + half) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int b = tid / half + u * rows_per_pass, j = tid % half;
      if (b < NB) {
        S.q_s[b * D + j] = x1[u] * rc[u] - x2[u] * rs[u];
        S.q_s[b * D + j + half] = x2[u] * rc[u] + x1[u] * rs[u];
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---------------- scores: LPK lanes per key, one 16-byte chunk of the key row per lane ----------------
  {
    const int LPK = min(32, D >> 3), KPW = 32 / LPK;       // lanes per key, keys per warp pass
    const int chunk = lane % LPK, sub = lane / LPK;
    const float inv_cap = a.softcap > 0.f ? 1.f / a.softcap : 0.f;
    for (int k0 = warp * KPW; k0 < koff[NB]; k0 += XF_WARPS * KPW) {
      const int k = k0 + sub;
      const bool valid = k < koff[NB];
      int b = 0;
#pragma unroll
      for (int bb = 1; bb < NB; ++bb) if (k >= koff[bb]) b = bb;
      float dot = 0.f;
      for (int c = chunk; c < (D >> 3); c += LPK) {        // one pass for D <= 256
        float kf[8];
        const uint4 kv = valid ? *reinterpret_cast<const uint4*>(S.kbuf + (size_t)k * D + c * 8) : make_uint4(0, 0, 0, 0);
        bf16x8_to_f32(kv, kf);
        const float4 q0 = *reinterpret_cast<const float4*>(S.q_s + b * D + c * 8);
        const float4 q1 = *reinterpret_cast<const float4*>(S.q_s + b * D + c * 8 + 4);
        dot = fmaf(q0.x, kf[0], dot); dot = fmaf(q0.y, kf[1], dot); dot = fmaf(q0.z, kf[2], dot); dot = fmaf(q0.w, kf[3], dot);
        dot = fmaf(q1.x, kf[4], dot); dot = fmaf(q1.y, kf[5], dot); dot = fmaf(q1.z, kf[6], dot); dot = fmaf(q1.w, kf[7], dot);
      }
      for (int o = LPK >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      float s = dot * a.scale;
      if (a.softcap > 0.f) {                               // softcap*tanh(s/softcap), same form as attention.cu
        const float e2 = __expf(2.f * s * inv_cap);
        s = a.softcap * (1.f - __fdividef(2.f, e2 + 1.f));
      }
      if (valid && chunk == 0) S.sc[k] = s;
    }
  }
  __syncthreads();
  // ---------------- softmax per row (warp b owns row b) ----------------
  if (warp < NB) {
    const int b = warp, n = nk[b], o0 = koff[b];
    float m = -INFINITY;
    for (int k = lane; k < n; k += 32) m = fmaxf(m, S.sc[o0 + k]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float l = 0.f;
    for (int k = lane; k < n; k += 32) { const float p = __expf(S.sc[o0 + k] - m); S.sc[o0 + k] = p; l += p; }
    l = warp_sum(l);
    if (lane == 0) S.ml[b] = l > 0.f ? 1.f / l : 0.f;
  }
  __syncthreads();
  // ---------------- P.V: thread = (dim pair, key partition) ----------------
  const int nparts = XF_THREADS / half;
  {
    const int dp = tid % half, part = tid / half;
    float acc[NB][2];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      acc[b][0] = 0.f; acc[b][1] = 0.f;
      for (int k = koff[b] + part; k < koff[b + 1]; k += nparts) {
        const float p = S.sc[k];
        const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(S.vbuf + (size_t)k * D + 2 * dp);
        acc[b][0] = fmaf(p, __low2float(v2), acc[b][0]);
        acc[b][1] = fmaf(p, __high2float(v2), acc[b][1]);
      }
      *reinterpret_cast<float2*>(S.part + ((size_t)part * NB + b) * D + 2 * dp) = make_float2(acc[b][0], acc[b][1]);
    }
  }
  __syncthreads();
  // ---------------- head output -> activation buffer of every CTA of the cluster (DSMEM) ----------------
  cluster.barrier_wait();                                  // every peer has started: its shared memory may be written
  {
    const int q4n = D >> 2, per_dst = NB * q4n;
    for (int i = tid; i < a.Hq * per_dst; i += XF_THREADS) {
      const int dst = i / per_dst, r = i - dst * per_dst, b = r / q4n, q4 = r - b * q4n;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p = 0; p < nparts; ++p) {
        const float4 v = *reinterpret_cast<const float4*>(S.part + ((size_t)p * NB + b) * D + 4 * q4);
        o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
      }
      const float inv = S.ml[b];
      o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
      const int kx = h * D + 4 * q4;                       // index into the o_proj input vector
      float4* base = ((kx >> 2) & 1) ? S.x_hi : S.x_lo;
      float4* remote = cluster.map_shared_rank(base, dst);
      remote[b * nchunks + (kx >> 3)] = o;
    }
  }
  cluster.sync();                                          // all heads have landed everywhere

  // ---------------- o_proj GEMV: weights are in registers ----------------
  for (int row = gw; row < a.N; row += total_warps) {
    if (row != gw) {
      const bf16* wr = a.W + (size_t)row * K;
#pragma unroll
      for (int i = 0; i < XF_U; ++i) {
        const int c = lane + 32 * i;
        w[i] = (c < nchunks) ? ldg_stream(wr + (size_t)c * 8, pol) : make_uint4(0, 0, 0, 0);
      }
    }
    float acc[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = 0.f;
#pragma unroll
    for (int i = 0; i < XF_U; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        float wf[8];
        bf16x8_to_f32(w[i], wf);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const float4 lo = S.x_lo[b * nchunks + c], hi = S.x_hi[b * nchunks + c];
          acc[b] = fmaf(wf[0], lo.x, acc[b]); acc[b] = fmaf(wf[1], lo.y, acc[b]);
          acc[b] = fmaf(wf[2], lo.z, acc[b]); acc[b] = fmaf(wf[3], lo.w, acc[b]);
          acc[b] = fmaf(wf[4], hi.x, acc[b]); acc[b] = fmaf(wf[5], hi.y, acc[b]);
          acc[b] = fmaf(wf[6], hi.z, acc[b]); acc[b] = fmaf(wf[7], hi.w, acc[b]);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = warp_sum(acc[b]);
    if (lane == 0) {
#pragma unroll
      for (int b = 0; b < NB; ++b) if (b < a.B) a.out[(size_t)b * a.out_stride + row] = acc[b];
    }
  }
  trace_end(a.trace);
}

size_t xf_smem_bytes(int NB, int K, int D, int cap) {
  return (size_t)NB * K * 4 + (size_t)cap * D * 4 + (size_t)((cap + 3) & ~3) * 4 + (size_t)(XF_THREADS / (D / 2)) * NB * D * 4 +
         (size_t)NB * D * 4 + 16 + (size_t)NB * XF_MAX_PAGES * 4;
}

template <int NB>
cudaError_t launch_xf(const XAttnOprojArgs& a, int num_sms, cudaStream_t st, bool pdl) {
  auto kern = xattn_oproj_kernel<NB>;
  const size_t smem = xf_smem_bytes(NB, a.K, a.D, a.max_keys);
  static size_t attr_set = 0;
  if (smem > attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if ((e = step_carveout(kern)) != cudaSuccess) return e;
    attr_set = smem;
  }
  cudaLaunchConfig_t cfg{};
  const int clusters = std::max(1, std::min(num_sms / a.Hq, (a.N + a.Hq * XF_WARPS - 1) / (a.Hq * XF_WARPS)));
  cfg.gridDim = dim3(clusters * a.Hq);
  cfg.blockDim = dim3(XF_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = a.Hq; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, a);
}

}  // namespace

// largest number of concatenated encoder keys (over the live rows) the fused kernel can stage; 0 = unsupported shape
int xattn_oproj_max_keys(int B, int Hq, int Hkv, int D, int K, int page_tokens, int num_sms) {
  if (B < 1 || B > 4 || Hq < 1 || Hq > 8 || Hkv < 1 || Hq % Hkv || num_sms < Hq) return 0;
  if (D < 16 || D > 256 || (D & (D - 1)) || K != Hq * D || K > XF_U * 32 * 8 || (K & 7)) return 0;
  const int NB = B <= 1 ? 1 : (B <= 2 ? 2 : 4);
  // stay inside the 100 KB shared-memory configuration: the GEMVs that follow lose bandwidth when the SM is left with a
  // small L1 (in-flight loads are tracked there): gate|up 15.5 -> 19.8 us at the maximum carve-out
  const size_t budget = 99 * 1024;
  const size_t fixed = xf_smem_bytes(NB, K, D, 0);
  if (fixed >= budget) return 0;
  int cap = (int)((budget - fixed) / ((size_t)D * 4 + 4));
  cap = std::min(cap, XF_MAX_PAGES * page_tokens);
  return cap & ~3;
}

cudaError_t launch_xattn_oproj(const XAttnOprojArgs& a, int num_sms, cudaStream_t st, bool pdl) {
  if (a.max_keys <= 0 || !a.rope_cs) return cudaErrorInvalidValue;
  if (a.B <= 1) return launch_xf<1>(a, num_sms, st, pdl);
  if (a.B <= 2) return launch_xf<2>(a, num_sms, st, pdl);
  return launch_xf<4>(a, num_sms, st, pdl);
}
