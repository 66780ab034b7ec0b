// Prefill-side glue kernels: embedding gather, RMSNorm sandwich, RoPE split + paged KV append,
// baseline SIMT GEMM (bring-up / cross-check for the tcgen05 GEMM), weight packing.
#include "kernels.h"

namespace {

__global__ void embed_kernel(const bf16* __restrict__ table, const int* __restrict__ ids, float scale,
                             float* __restrict__ h, int M, int d) {
  const int t = blockIdx.x;
  const bf16* row = table + (size_t)ids[t] * d;
  for (int k = threadIdx.x; k < d; k += blockDim.x) h[(size_t)t * d + k] = __bfloat162float(row[k]) * scale;
}

// (slots are written by the sampler of the same step: no __restrict__/read-only path under PDL)
__global__ void embed_slots_kernel(const bf16* __restrict__ table, const SlotDev* slots, float scale,
                                   float* h, int d) {
  const int b = blockIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (!slots[b].active) return;
  const bf16* row = table + (size_t)slots[b].last_token * d;
  for (int k = threadIdx.x; k < d; k += blockDim.x) h[(size_t)b * d + k] = __bfloat162float(row[k]) * scale;
}

// one CTA per token: h_out = h_in + rmsnorm(y)*g_post ; xn/xf = rmsnorm(h_out)*g_pre  (HF:66-74 in fp32).
// Single reduction pass: with r = rsqrt(mean(y^2)+eps), sum(h + y r g)^2 = S2 + 2 r S3 + r^2 S4; all loads are
// 128-bit and issued before the reduction; the gains are fetched before griddepcontrol.wait (PDL).
// Two shapes: 512 threads x 2 float4 for the batched decode step (few rows: the row's latency is what counts) and
// 192 threads x 4 float4 for prefill (thousands of rows: d = 2304 is exactly 3 float4 per thread, ten CTAs per SM).
// (activations are read with ld.global.cg: under PDL this kernel is launched while its producer is still running, so
// they must not come from the non-coherent read-only path)
template <int NK_THREADS, int NK_MAXV>
__global__ void __launch_bounds__(NK_THREADS) norm_kernel(const float* h_in, const float* y,
                                                          const float* __restrict__ g_post, const float* __restrict__ g_pre,
                                                          float* h_out, bf16* xn, float* xf, int d, float eps,
                                                          float* zero_a, int na, float* zero_b, int nb,
                                                          unsigned long long* trace) {
  __shared__ float red[128];
  const size_t base = (size_t)blockIdx.x * d;
  const int nv = d >> 2;
  float4 gp[NK_MAXV], gq[NK_MAXV];
#pragma unroll
  for (int u = 0; u < NK_MAXV; ++u) {
    const int i = threadIdx.x + u * NK_THREADS;
    gp[u] = (g_pre && i < nv) ? reinterpret_cast<const float4*>(g_pre)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    gq[u] = (y && i < nv) ? reinterpret_cast<const float4*>(g_post)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  pdl_launch_dependents();
  pdl_wait();
  trace_begin(trace);
  float4 hv[NK_MAXV], yv[NK_MAXV];
  float s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
#pragma unroll
  for (int u = 0; u < NK_MAXV; ++u) {
    const int i = threadIdx.x + u * NK_THREADS;
    hv[u] = (i < nv) ? __ldcg(reinterpret_cast<const float4*>(h_in + base) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    yv[u] = (y && i < nv) ? __ldcg(reinterpret_cast<const float4*>(y + base) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int u = 0; u < NK_MAXV; ++u) {
    const float hh[4] = {hv[u].x, hv[u].y, hv[u].z, hv[u].w};
    const float yy[4] = {yv[u].x, yv[u].y, yv[u].z, yv[u].w};
    const float gg[4] = {gq[u].x, gq[u].y, gq[u].z, gq[u].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float yg = yy[j] * gg[j];
      s1 = fmaf(yy[j], yy[j], s1); s2 = fmaf(hh[j], hh[j], s2); s3 = fmaf(hh[j], yg, s3); s4 = fmaf(yg, yg, s4);
    }
    yv[u] = make_float4(yy[0] * gg[0], yy[1] * gg[1], yy[2] * gg[2], yy[3] * gg[3]);     // y*g_post
  }
  // optional side job: zero this row of up to two buffers that the following split-K GEMMs accumulate into with
  // red.global.add (after the loads above, so `y` itself may be one of them)
  if (zero_a) for (int i = threadIdx.x; i < (na >> 2); i += NK_THREADS) reinterpret_cast<float4*>(zero_a + (size_t)blockIdx.x * na)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (zero_b) for (int i = threadIdx.x; i < (nb >> 2); i += NK_THREADS) reinterpret_cast<float4*>(zero_b + (size_t)blockIdx.x * nb)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  block_sum4(s1, s2, s3, s4, red);
  const float ry = y ? rsqrtf(s1 / (float)d + eps) : 0.f;
  const float ss = s2 + 2.f * ry * s3 + ry * ry * s4;
  const float rinv = rsqrtf(ss / (float)d + eps);
#pragma unroll
  for (int u = 0; u < NK_MAXV; ++u) {
    const int i = threadIdx.x + u * NK_THREADS;
    if (i >= nv) continue;
    float4 hn = make_float4(fmaf(yv[u].x, ry, hv[u].x), fmaf(yv[u].y, ry, hv[u].y), fmaf(yv[u].z, ry, hv[u].z), fmaf(yv[u].w, ry, hv[u].w));
    if (h_out) reinterpret_cast<float4*>(h_out + base)[i] = hn;
    if (g_pre) {
      const float4 o = make_float4(hn.x * rinv * gp[u].x, hn.y * rinv * gp[u].y, hn.z * rinv * gp[u].z, hn.w * rinv * gp[u].w);
      if (xf) reinterpret_cast<float4*>(xf + base)[i] = o;
      if (xn) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
        uint2 pk; pk.x = *reinterpret_cast<unsigned*>(&lo); pk.y = *reinterpret_cast<unsigned*>(&hi);
        reinterpret_cast<uint2*>(xn + base)[i] = pk;
      }
    }
  }
  trace_end(trace);
}

// one CTA per token; threads over (head, j<D/2)
__global__ void rope_split_kernel(RopeSplitArgs a) {
  pdl_launch_dependents();           // launched as a programmatic dependent in the prefill chain: nothing is read before the wait
  pdl_wait();
  const int t = blockIdx.x;
  const int D = a.D, half = D / 2;
  const float pos = a.pos ? a.pos[t] : 0.f;
  const float* row = a.qkv + (size_t)t * a.ld;
  int slot = -1, idx = 0, page = 0, off = 0;
  if (a.block_table) {
    slot = a.tok_slot[t]; idx = a.tok_idx[t];
    page = a.block_table[(size_t)slot * a.bt_stride + idx / a.pool.page_tokens];
    off = idx % a.pool.page_tokens;
  }
  // thread = (two adjacent rotation pairs j, j+1; head group): the angles are computed once per thread, the loads of a batch
  // of heads are in flight together, and every access is 8 bytes in / 4 bytes out per lane (the per-element version paid a
  // sincosf and an L2 round trip per head; the one-pair version moved 57 us per 8192-token layer at 3.5 TB/s)
  const int nqh = a.q_off >= 0 ? a.Hq : 0, nkh = a.k_off >= 0 ? a.Hkv : 0, HT = nqh + nkh;
  const int hp = half / 2;                                   // pairs of pairs per head
  const int j = 2 * (threadIdx.x % hp), grp = threadIdx.x / hp, ngrp = blockDim.x / hp;
  float s0 = 0.f, c0 = 1.f, s1 = 0.f, c1 = 1.f;
  if (a.pos) { sincosf(pos * a.inv_freq[j], &s0, &c0); sincosf(pos * a.inv_freq[j + 1], &s1, &c1); }
  constexpr int HB = 4;
  for (int h0 = grp; h0 < HT; h0 += ngrp * HB) {
    float2 x1[HB], x2[HB];
#pragma unroll
    for (int u = 0; u < HB; ++u) {
      const int h = h0 + u * ngrp;
      if (h < HT) {
        const float* src = row + (h < nqh ? a.q_off + h * D : a.k_off + (h - nqh) * D);
        x1[u] = *reinterpret_cast<const float2*>(src + j); x2[u] = *reinterpret_cast<const float2*>(src + j + half);
      }
    }
#pragma unroll
    for (int u = 0; u < HB; ++u) {
      const int h = h0 + u * ngrp;
      if (h >= HT) continue;
      const __nv_bfloat162 r1 = __floats2bfloat162_rn(x1[u].x * c0 - x2[u].x * s0, x1[u].y * c1 - x2[u].y * s1);
      const __nv_bfloat162 r2 = __floats2bfloat162_rn(x2[u].x * c0 + x1[u].x * s0, x2[u].y * c1 + x1[u].y * s1);
      if (h < nqh) {
        bf16* dst = a.q_out + (size_t)t * a.Hq * D + h * D;
        *reinterpret_cast<__nv_bfloat162*>(dst + j) = r1; *reinterpret_cast<__nv_bfloat162*>(dst + j + half) = r2;
      } else {
        const int hd = h - nqh;
        if (a.k_out) {
          bf16* dst = a.k_out + (size_t)t * a.Hkv * D + hd * D;
          *reinterpret_cast<__nv_bfloat162*>(dst + j) = r1; *reinterpret_cast<__nv_bfloat162*>(dst + j + half) = r2;
        }
        if (a.block_table) {
          bf16* dst = a.pool.ptr(a.layer, 0, page) + ((size_t)hd * a.pool.page_tokens + off) * D;
          *reinterpret_cast<__nv_bfloat162*>(dst + j) = r1; *reinterpret_cast<__nv_bfloat162*>(dst + j + half) = r2;
        }
      }
    }
  }
  if (a.v_off >= 0) {
    const int nv2 = a.Hkv * D / 2;                           // pairs of V elements
    for (int i0 = threadIdx.x; i0 < nv2; i0 += blockDim.x * HB) {
      float2 v[HB];
#pragma unroll
      for (int u = 0; u < HB; ++u) { const int i = i0 + u * blockDim.x; v[u] = (i < nv2) ? *reinterpret_cast<const float2*>(row + a.v_off + 2 * i) : make_float2(0.f, 0.f); }
#pragma unroll
      for (int u = 0; u < HB; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i >= nv2) continue;
        const int e0 = 2 * i, hd = e0 / D, jj = e0 - hd * D;
        const __nv_bfloat162 vb = __floats2bfloat162_rn(v[u].x, v[u].y);
        if (a.v_out) *reinterpret_cast<__nv_bfloat162*>(a.v_out + (size_t)t * a.Hkv * D + e0) = vb;
        if (a.block_table) *reinterpret_cast<__nv_bfloat162*>(a.pool.ptr(a.layer, 1, page) + ((size_t)hd * a.pool.page_tokens + off) * D + jj) = vb;
      }
    }
  }
}

// ---- baseline GEMM: 64x64 tile, BK=16, 256 threads, 4x4 micro-tile, fp32 FMA ---------------------
constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;
__device__ __forceinline__ void gemm_store(const GemmArgs& a, int m, int n, float v, float v_pair) {
  switch (a.epilogue) {
    case GE_F32: reinterpret_cast<float*>(a.out)[(size_t)m * a.ldo + n] = v; break;
    case GE_BF16: reinterpret_cast<bf16*>(a.out)[(size_t)m * a.ldo + n] = __float2bfloat16(v); break;
    case GE_BIAS_F32: reinterpret_cast<float*>(a.out)[(size_t)m * a.ldo + n] = v + a.bias[n]; break;
    case GE_BIAS_GELU_BF16:
      reinterpret_cast<bf16*>(a.out)[(size_t)m * a.ldo + n] = __float2bfloat16(gelu_erf_f(v + a.bias[n])); break;
    case GE_GEGLU_BF16:   // n even = gate row, v_pair = up row
      reinterpret_cast<bf16*>(a.out)[(size_t)m * a.ldo + (n >> 1)] = __float2bfloat16(gelu_tanh_f(v) * v_pair); break;
  }
}

__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs a) {
  __shared__ float As[SG_BK][SG_BM + 1];
  __shared__ float Ws[SG_BK][SG_BN + 1];
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16x16 threads, each 4x4
  float acc[4][4] = {};
  for (int k0 = 0; k0 < a.K; k0 += SG_BK) {
    for (int i = threadIdx.x; i < SG_BM * SG_BK; i += 256) {
      int r = i / SG_BK, c = i % SG_BK;
      int m = m0 + r, k = k0 + c;
      As[c][r] = (m < a.M && k < a.K) ? __bfloat162float(a.A[(size_t)m * a.K + k]) : 0.f;
      int n = n0 + r;
      Ws[c][r] = (n < a.N && k < a.K) ? __bfloat162float(a.W[(size_t)n * a.K + k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float av[4], wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = As[k][ty * 4 + i]; wv[i] = Ws[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
    if (a.epilogue == GE_GEGLU_BF16) {
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        const int n = n0 + tx * 4 + j;
        if (n + 1 < a.N) gemm_store(a, m, n, acc[i][j], acc[i][j + 1]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n < a.N) gemm_store(a, m, n, acc[i][j], 0.f);
      }
    }
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ rows,
                                   float* __restrict__ dst, int d) {
  const size_t s = (size_t)rows[blockIdx.x] * d, o = (size_t)blockIdx.x * d;
  for (int k = threadIdx.x; k < d; k += blockDim.x) dst[o + k] = src[s + k];
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

__device__ __forceinline__ float load_as_f32(const void* p, int dtype, size_t i) {
  if (dtype == 0) return reinterpret_cast<const float*>(p)[i];
  if (dtype == 1) return __bfloat162float(reinterpret_cast<const bf16*>(p)[i]);
  return __half2float(reinterpret_cast<const __half*>(p)[i]);
}

__global__ void pack_kernel(const void* __restrict__ src, int src_dtype, void* __restrict__ dst, int dst_is_bf16,
                            long long rows, long long cols, long long row_off, long long row_mul, int add_one) {
  const long long n = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols, c = i - r * cols;
    float v = load_as_f32(src, src_dtype, (size_t)i);
    if (add_one) v += 1.0f;
    const size_t o = (size_t)(row_off + r * row_mul) * cols + c;
    if (dst_is_bf16) reinterpret_cast<bf16*>(dst)[o] = __float2bfloat16(v);
    else reinterpret_cast<float*>(dst)[o] = v;
  }
}

}  // namespace

cudaError_t launch_embed(const bf16* table, const int* ids, float scale, float* h, int M, int d, cudaStream_t st) {
  if (M <= 0) return cudaSuccess;
  embed_kernel<<<M, 256, 0, st>>>(table, ids, scale, h, M, d);
  return cudaGetLastError();
}

cudaError_t launch_embed_slots(const bf16* table, const SlotDev* slots, float scale, float* h, int B, int d, cudaStream_t st, bool pdl) {
  if (B <= 0) return cudaSuccess;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(B);
  cfg.blockDim = dim3(256);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, embed_slots_kernel, table, slots, scale, h, d);
}

cudaError_t launch_norm(const float* h_in, const float* y, const float* g_post, const float* g_pre, float* h_out,
                        bf16* xn, float* xf, int M, int d, float eps, cudaStream_t st, bool pdl,
                        float* zero_a, int na, float* zero_b, int nb, unsigned long long* trace) {
  if (M <= 0) return cudaSuccess;
  if (d % 4 != 0 || d > 512 * 4 * 2 || (na & 3) || (nb & 3)) return cudaErrorInvalidValue;
  const bool many = M >= 1024 && d <= 192 * 4 * 4;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(M);
  cfg.blockDim = dim3(many ? 192 : 512);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  if (many) return cudaLaunchKernelEx(&cfg, norm_kernel<192, 4>, h_in, y, g_post, g_pre, h_out, xn, xf, d, eps, zero_a, na, zero_b, nb, trace);
  return cudaLaunchKernelEx(&cfg, norm_kernel<512, 2>, h_in, y, g_post, g_pre, h_out, xn, xf, d, eps, zero_a, na, zero_b, nb, trace);
}

cudaError_t launch_rope_split(const RopeSplitArgs& a, cudaStream_t st, bool pdl) {
  if (a.M <= 0) return cudaSuccess;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(a.M);
  cfg.blockDim = dim3(256);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, rope_split_kernel, a);
}

cudaError_t launch_gemm_simt(const GemmArgs& a, cudaStream_t st) {
  if (a.M <= 0) return cudaSuccess;
  dim3 grid((a.N + SG_BN - 1) / SG_BN, (a.M + SG_BM - 1) / SG_BM);
  gemm_simt_kernel<<<grid, 256, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_gather_rows(const float* src, const int* rows, float* dst, int n, int d, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  gather_rows_kernel<<<n, 256, 0, st>>>(src, rows, dst, d);
  return cudaGetLastError();
}

cudaError_t launch_f32_to_bf16(const float* src, bf16* dst, size_t n, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  int grid = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  f32_to_bf16_kernel<<<grid, 256, 0, st>>>(src, dst, n);
  return cudaGetLastError();
}

cudaError_t launch_pack(const void* src, int src_dtype, void* dst, int dst_is_bf16, int64_t rows, int64_t cols,
                        int64_t row_off, int64_t row_mul, int add_one, cudaStream_t st) {
  const int64_t n = rows * cols;
  if (n == 0) return cudaSuccess;
  int grid = (int)((n + 255) / 256 < 8192 ? (n + 255) / 256 : 8192);
  pack_kernel<<<grid, 256, 0, st>>>(src, src_dtype, dst, dst_is_bf16, rows, cols, row_off, row_mul, add_one);
  return cudaGetLastError();
}
