// Dense projections on the 5th-generation tensor cores: C[M,N] = A[M,K] * W[N,K]^T, bf16 in, fp32 accumulate.
// Used for the encoder/decoder prefill GEMMs (M = packed tokens) and for batched decode (M = batch rows).
//
// Formulation ("swap-AB"): the WEIGHT tile is the tcgen05 A operand (UMMA M = 128 output features = the 128
// TMEM lanes) and the ACTIVATION tile is the B operand (UMMA N = 16..256 tokens = TMEM columns), so one
// kernel covers 8-row batched decode up to 8192-token prefill without padding the token dimension to 128.
// Both operands are K-major in global memory exactly as stored ([N,K] weights, [M,K] activations): TMA
// (cp.async.bulk.tensor.2d, 128-byte swizzle) stages 64-element K-slabs through an mbarrier ring, one elected
// thread issues tcgen05.mma (kind::f16, fp32 accumulators in TMEM), tcgen05.commit releases ring slots and
// signals the epilogue warps, which read the accumulators with tcgen05.ld (32 lanes x 32 bit, 16 columns per
// instruction) and apply the fused epilogue (bias / exact GELU / GeGLU / bf16 cast).  When the output grid is
// smaller than the machine (weight-streaming regime) K is split across the CTAs of a thread-block cluster
// (1,1,split<=8): every CTA parks its fp32 partial tile in its own shared memory and the cluster reduces it through
// distributed shared memory (no atomics, no zero-fill, every fused epilogue stays available).
#include "kernels.h"
#include "tc_common.cuh"

#include <cooperative_groups.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

namespace {

constexpr int TC_BM = 128;          // output features per tile (UMMA M)
constexpr int TC_BK = 64;           // K elements per stage (one 128-byte swizzle atom of bf16)
constexpr int TC_THREADS = 320;     // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-9: epilogue (two per TMEM lane quarter)

using namespace tc;

struct TcParams {
  int M, N, K;
  int epilogue; const float* bias;
  void* out; int ldo;
  int k_blocks_per_split;     // K-slabs handled by one CTA (blockIdx.z selects the split)
  int split_k;
  int atomic;                 // split-K partials reduced with red.global.add.f32 into a zeroed fp32 output (GE_F32)
  unsigned long long* trace;
  unsigned long long* probe;
  float* zero_a; size_t zero_na;
  float* zero_b; size_t zero_nb;
};

// GeGLU in the bf16 epilogues: tanh.approx.f32 (MUFU, rel. error ~2^-11, far inside bf16's 2^-8) -- the batched
// epilogue is instruction-issue bound (64 values per thread), the libm tanhf path costs ~40 instructions per value
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(k0 * (x + k1 * x * x * x)));
  return 0.5f * x * (1.f + t);
}

__device__ __forceinline__ void epilogue_store(const TcParams& p, int t, int f, float val, float up, float bias, bool even) {
  if (t >= p.M || f >= p.N) return;
  switch (p.epilogue) {
    case GE_F32: reinterpret_cast<float*>(p.out)[(size_t)t * p.ldo + f] = val; break;
    case GE_BF16: reinterpret_cast<bf16*>(p.out)[(size_t)t * p.ldo + f] = __float2bfloat16(val); break;
    case GE_BIAS_F32: reinterpret_cast<float*>(p.out)[(size_t)t * p.ldo + f] = val + bias; break;
    case GE_BIAS_GELU_BF16:
      reinterpret_cast<bf16*>(p.out)[(size_t)t * p.ldo + f] = __float2bfloat16(gelu_erf_f(val + bias)); break;
    case GE_GEGLU_BF16:     // even feature = gate row, `up` = the odd neighbour
      if (even) reinterpret_cast<bf16*>(p.out)[(size_t)t * p.ldo + (f >> 1)] = __float2bfloat16(gelu_tanh_fast(val) * up);
      break;
  }
}

// one 16-token chunk of a TMEM lane (= output feature f) -> global, epilogue kind resolved at compile time
template <int EPI>
__device__ __forceinline__ void store_chunk(const TcParams& p, const float (&v)[16], int tbase, int f, float bias, bool even) {
  const bool fok = f < p.N;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int t = tbase + j;
    if (t >= p.M) break;                                   // warp-uniform
    if (EPI == GE_GEGLU_BF16) {
      const float up = __shfl_xor_sync(0xffffffffu, v[j], 1);
      if (even && fok) reinterpret_cast<bf16*>(p.out)[(size_t)t * p.ldo + (f >> 1)] = __float2bfloat16(gelu_tanh_fast(v[j]) * up);
    } else if (fok) {
      if (EPI == GE_F32) reinterpret_cast<float*>(p.out)[(size_t)t * p.ldo + f] = v[j];
      else if (EPI == GE_BF16) reinterpret_cast<bf16*>(p.out)[(size_t)t * p.ldo + f] = __float2bfloat16(v[j]);
      else if (EPI == GE_BIAS_F32) reinterpret_cast<float*>(p.out)[(size_t)t * p.ldo + f] = v[j] + bias;
      else if (EPI == GE_BIAS_GELU_BF16) reinterpret_cast<bf16*>(p.out)[(size_t)t * p.ldo + f] = __float2bfloat16(gelu_erf_f(v[j] + bias));
    }
  }
}

template <int TOKT, int NSTAGE>
struct __align__(1024) TcSmem {
  unsigned char w[NSTAGE][TC_BM * 128];       // weight tiles: 128 rows x 128 B (swizzled by TMA)
  unsigned char x[NSTAGE][TOKT * 128];        // activation tiles: TOKT rows x 128 B
  uint64_t full[NSTAGE], empty[NSTAGE], acc_full;
  uint32_t tmem_base;
};

template <int TOKT, int NSTAGE>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, TcParams p) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  // dynamic shared memory is only guaranteed 16-byte aligned: realign to the 1024 B the swizzle atom needs
  TcSmem<TOKT, NSTAGE>& S = *reinterpret_cast<TcSmem<TOKT, NSTAGE>*>(
      (reinterpret_cast<uintptr_t>(smraw) + 1023) & ~(uintptr_t)1023);
  constexpr int TMEM_COLS = TOKT < 32 ? 32 : TOKT == 192 ? 256 : TOKT;     // power of two
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f0 = blockIdx.x * TC_BM, t0 = blockIdx.y * TOKT;
  const int kb_total = p.K / TC_BK;
  const int kb0 = blockIdx.z * p.k_blocks_per_split;
  const int nkb = min(p.k_blocks_per_split, kb_total - kb0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&S.full[i], 1); mbar_init(&S.empty[i], 1); }
    mbar_init(&S.acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = S.tmem_base;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      constexpr uint32_t STAGE_BYTES = TC_BM * 128 + TOKT * 128;
      // weights are immutable: fill the ring with the weight tiles before the dependency on the previous kernel
      // resolves (programmatic dependent launch); the activation tiles follow after griddepcontrol.wait
      const int pre = min(nkb, NSTAGE);
      for (int i = 0; i < pre; ++i) {
        mbar_expect_tx(&S.full[i], STAGE_BYTES);
        tma_load_2d(S.w[i], &map_w, (kb0 + i) * TC_BK, f0, &S.full[i]);
      }
      pdl_launch_dependents();
      pdl_wait();
      trace_begin(p.trace);
      for (int i = 0; i < pre; ++i) tma_load_2d(S.x[i], &map_x, (kb0 + i) * TC_BK, t0, &S.full[i]);
      for (int i = NSTAGE; i < nkb; ++i) {
        const int s = i % NSTAGE;
        mbar_wait(&S.empty[s], ((i / NSTAGE) - 1) & 1);
        mbar_expect_tx(&S.full[s], STAGE_BYTES);
        const int kc = (kb0 + i) * TC_BK;
        tma_load_2d(S.w[s], &map_w, kc, f0, &S.full[s]);
        tma_load_2d(S.x[s], &map_x, kc, t0, &S.full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      // instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, K-major A/B, N>>3 [17,23), M>>4 [24,29)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TOKT >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % NSTAGE;
        mbar_wait(&S.full[s], (i / NSTAGE) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t wa = smem_u32(S.w[s]), xa = smem_u32(S.x[s]);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {       // UMMA K = 16 bf16 = 32 bytes inside the swizzle atom
          const uint64_t da = umma_desc_sw128(wa + k * 32), db = umma_desc_sw128(xa + k * 32);
          umma_bf16(tmem_d, da, db, idesc, (i | k) ? 1u : 0u);
        }
        umma_commit(&S.empty[s]);                     // frees the slot once these MMAs have read it
      }
      umma_commit(&S.acc_full);                       // accumulator complete
    }
  } else {
    // ===== epilogue warps: TMEM -> registers -> global (or -> shared memory for the cluster split-K reduce).
    //       Warps 2-5 take the first half of the token columns, warps 6-9 the second half; the epilogue kind is
    //       resolved once (a per-element switch + libm tanhf made the epilogue as long as the main loop). =====
    const int q = warp & 3;                           // TMEM lane quarter this warp may access
    const int fl = q * 32 + lane;                     // feature inside the tile
    const int f = f0 + fl;
    constexpr int NCH = TOKT / 16;
    const int hsel = (warp - 2) >> 2;
    const int ch0 = (NCH >= 2) ? hsel * (NCH / 2) : 0;
    const int ch1 = (NCH >= 2) ? (hsel + 1) * (NCH / 2) : (hsel == 0 ? NCH : 0);
    pdl_wait();                                       // the output buffer may still be read by the previous kernel
    if (p.zero_a) {                                   // side job while the main loop runs (see GemmArgs)
      const size_t nthr = (size_t)gridDim.x * gridDim.y * gridDim.z * (TC_THREADS - 64);
      const size_t gtid = ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * (TC_THREADS - 64) + (threadIdx.x - 64);
      for (size_t i = gtid; i < p.zero_na / 4; i += nthr) reinterpret_cast<float4*>(p.zero_a)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.zero_b) for (size_t i = gtid; i < p.zero_nb / 4; i += nthr) reinterpret_cast<float4*>(p.zero_b)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    mbar_wait(&S.acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* red = reinterpret_cast<float*>(&S.w[0][0]);   // ring storage is free once acc_full has fired
    const float bias = (p.bias && f < p.N) ? p.bias[f] : 0.f;
    const bool even = (lane & 1) == 0;
#pragma unroll 1
    for (int ch = ch0; ch < ch1; ++ch) {
      const int c = ch * 16;
      if (t0 + c >= p.M) break;                       // warp-uniform
      float v[16];
      tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      if (p.atomic) {
        float* o = reinterpret_cast<float*>(p.out) + (size_t)(t0 + c) * p.ldo + f;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (t0 + c + j < p.M && f < p.N) atomicAdd(o + (size_t)j * p.ldo, v[j]);
      } else if (p.split_k > 1) {
#pragma unroll
        for (int j = 0; j < 16; ++j) red[(c + j) * TC_BM + fl] = v[j];
      } else {
        switch (p.epilogue) {
          case GE_F32: store_chunk<GE_F32>(p, v, t0 + c, f, bias, even); break;
          case GE_BF16: store_chunk<GE_BF16>(p, v, t0 + c, f, bias, even); break;
          case GE_BIAS_F32: store_chunk<GE_BIAS_F32>(p, v, t0 + c, f, bias, even); break;
          case GE_BIAS_GELU_BF16: store_chunk<GE_BIAS_GELU_BF16>(p, v, t0 + c, f, bias, even); break;
          default: store_chunk<GE_GEGLU_BF16>(p, v, t0 + c, f, bias, even); break;
        }
      }
    }
  }
  if (p.split_k > 1 && !p.atomic) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();                                   // all partial tiles are parked in shared memory
    // all 8 epilogue warps reduce: warp pair (w, w+4) shares a lane quarter's 32 features and alternates over this rank's
    // tokens; the remote reads of up to 4 tokens x 8 ranks are issued before anything is summed (a dependent chain of
    // distributed-shared-memory round trips made this reduction 4x slower than the atomic one: fc1 of the batched head 16 us)
    if (warp >= 2) {
      const int rank = (int)cluster.block_rank(), nr = p.split_k;
      const int fl = (warp & 3) * 32 + lane, f = f0 + fl;
      const int hsel = (warp - 2) >> 2;
      const float bias = (p.bias && f < p.N) ? p.bias[f] : 0.f;
      float* red = reinterpret_cast<float*>(&S.w[0][0]);
      const float* rr[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) rr[r] = cluster.map_shared_rank(red, r < nr ? r : 0);
      const bool geglu = p.epilogue == GE_GEGLU_BF16;
      constexpr int UT = 4;                                   // tokens in flight per thread
      for (int j0 = rank + hsel * nr; j0 < TOKT && t0 + j0 < p.M; j0 += 2 * UT * nr) {
        float v[UT][8], vu[UT][8];
#pragma unroll
        for (int u = 0; u < UT; ++u) {
          const int j = j0 + u * 2 * nr;
          const bool ok = j < TOKT && t0 + j < p.M;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            v[u][r] = (ok && r < nr) ? rr[r][j * TC_BM + fl] : 0.f;
            vu[u][r] = (ok && r < nr && geglu) ? rr[r][j * TC_BM + (fl | 1)] : 0.f;
          }
        }
#pragma unroll
        for (int u = 0; u < UT; ++u) {
          const int j = j0 + u * 2 * nr;
          if (j < TOKT && t0 + j < p.M) {
            float acc = 0.f, acc_up = 0.f;
#pragma unroll
            for (int r = 0; r < 8; ++r) { acc += v[u][r]; acc_up += vu[u][r]; }   // rank order: same sum as before
            epilogue_store(p, t0 + j, f, acc, acc_up, bias, (fl & 1) == 0);
          }
        }
      }
    }
    cluster.sync();                                   // peers may still be reading this CTA's shared memory
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  trace_end(p.trace);
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TMEM_COLS) : "memory");
  }
}

template <int TOKT, int NSTAGE>
cudaError_t launch_tc(const CUtensorMap& mw, const CUtensorMap& mx, const TcParams& p, dim3 grid, cudaStream_t st, bool pdl) {
  auto kern = gemm_tc_kernel<TOKT, NSTAGE>;
  const size_t smem = sizeof(TcSmem<TOKT, NSTAGE>) + 1024;
  static PerDeviceFlag attr_set;
  if (!attr_set.here()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // batched-step kernels all ask for the maximum shared-memory carve-out: CTAs of consecutive kernels can then share an SM
    if (batched_carveout() >= 0 && (e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, batched_carveout())) != cudaSuccess) return e;
    attr_set.here() = 1;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = p.atomic ? 1 : p.split_k;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, mw, mx, p);
}

// ------------------------------------------------------------------------------------------------------------------
// Prefill GEMM (tokens >> 128): persistent CTA PAIRS, tcgen05.mma.cta_group::2, two TMEM accumulators.
//
// What the one-tile-per-CTA kernel above left on the table at 8192 tokens (measured, B200): 9 us of fixed cost per 128x256
// tile (launch, ring fill, a TMEM -> global epilogue that nothing overlaps) against 9.4 us of MMAs at K = 2304, and a main
// loop that ran at 64 % of the MMA rate: a single SM reads 12 KB of operands per 128-cycle MMA (96 B/clk) while TMA writes
// 48 KB per K-slab into the same shared memory (96 B/clk) -- together 1.5x the 128 B/clk the SM's shared memory moves.
// (Multicasting the activation tile over a cluster did not help: it removes L2 reads, not shared-memory traffic.)
//
// Here a cluster of two CTAs (the two SMs of a TPC) owns a 256-feature x 256-token tile: each CTA stages ITS 128 weight
// rows and ITS 128 token rows (32 KB per K-slab instead of 48), the leader's single thread issues 256x256x16 MMAs that
// read both CTAs' shared memory, and each CTA's TMEM receives its 128 features x 256 tokens.  Pairs are persistent: they
// walk a rasterised tile list (groups of 8 feature tiles x all token tiles, so a wave re-reads ~2k weight rows and ~2.4k
// token rows from L2), the 6-stage TMA ring runs ahead across tile boundaries, and the accumulator ping-pongs between
// TMEM columns [0,256) and [256,512) so the 8 epilogue warps drain tile i while the MMAs of tile i+1 are in flight.
// Barriers: full[s] lives in the leader (both CTAs' TMA bytes complete there), empty[s] / acc_full[a] are signalled in
// both CTAs by multicast tcgen05.commit, acc_empty[a] lives in the leader and collects the 16 epilogue warps of the pair.
// ------------------------------------------------------------------------------------------------------------------
constexpr int P2_TOK = 256;         // tokens per tile (UMMA N)
constexpr int P2_STAGES = 6;
constexpr int P2_GROUP = 8;         // feature tiles (of 256) per raster group

struct __align__(1024) Tc2Smem {
  unsigned char w[P2_STAGES][TC_BM * 128];     // this CTA's 128 weight rows of the K-slab
  unsigned char x[P2_STAGES][128 * 128];       // this CTA's 128 token rows of the K-slab
  uint64_t full[P2_STAGES], empty[P2_STAGES], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void tc2_tile(int id, int nfp, int mt, int& fp, int& tt) {
  const int per_group = P2_GROUP * mt;
  const int g = id / per_group, r = id - g * per_group;
  const int gw = min(P2_GROUP, nfp - g * P2_GROUP);
  tt = r / gw;
  fp = g * P2_GROUP + (r - tt * gw);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, TcParams p) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  Tc2Smem& S = *reinterpret_cast<Tc2Smem*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nfp = (p.N + 255) / 256, mt = (p.M + P2_TOK - 1) / P2_TOK;
  const int ntiles = nfp * mt;
  const int nkb = p.K / TC_BK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < P2_STAGES; ++i) { mbar_init(&S.full[i], 1); mbar_init(&S.empty[i], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&S.acc_full[a], 1); mbar_init(&S.acc_empty[a], 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
  }
  if (warp == 1) {      // the same warp of both CTAs allocates all 512 columns in both TMEMs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                                  // the peer's barriers exist before anything is signalled on them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = S.tmem_base;

  if (warp == 0) {
    // ===== TMA producer (one thread in each CTA) =====
    if (lane == 0) {
      pdl_launch_dependents();
      pdl_wait();
      trace_begin(p.trace);
      const uint32_t full_leader = mapa_u32(smem_u32(&S.full[0]), 0);
      int s = 0; uint32_t round = 0;
      for (int tile = pair; tile < ntiles; tile += npairs) {
        int fp, tt;
        tc2_tile(tile, nfp, mt, fp, tt);
        const int f0 = fp * 256 + (int)crank * 128, t0 = tt * P2_TOK + (int)crank * 128;
        for (int kb = 0; kb < nkb; ++kb) {
          if (round > 0) mbar_wait(&S.empty[s], (round - 1) & 1);
          if (crank == 0) mbar_expect_tx(&S.full[s], 2u * (TC_BM * 128 + 128 * 128));
          tma_load_2d_cg2(S.w[s], &map_w, kb * TC_BK, f0, full_leader + 8u * s);
          tma_load_2d_cg2(S.x[s], &map_x, kb * TC_BK, t0, full_leader + 8u * s);
          if (++s == P2_STAGES) { s = 0; ++round; }
        }
      }
      // drain: the leader's multicast commits of the last ring-full still arrive on THIS CTA's empty barriers
      for (int i = 0; i < P2_STAGES; ++i) {
        if (round > 0 || i < s) {                      // slot s was last filled in round (i < s ? round : round - 1)
          const uint32_t r = (i < s) ? round : round - 1;
          mbar_wait(&S.empty[i], r & 1);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread of the leader CTA =====
    if (lane == 0 && crank == 0) {
      // instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, K-major A/B, N>>3 [17,23), M>>4 [24,29)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P2_TOK >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int s = 0; uint32_t round = 0, tl = 0;
      for (int tile = pair; tile < ntiles; tile += npairs, ++tl) {
        const uint32_t a = tl & 1;
        if (tl >= 2) {                                 // both CTAs' epilogue warps have drained this accumulator
          mbar_wait(&S.acc_empty[a], ((tl >> 1) - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t tmem_d = tmem_base + a * P2_TOK;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&S.full[s], round & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t wa = smem_u32(S.w[s]), xa = smem_u32(S.x[s]);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_bf16_cg2(tmem_d, umma_desc_sw128(wa + k * 32), umma_desc_sw128(xa + k * 32), idesc, (kb | k) ? 1u : 0u);
          umma_commit_cg2(&S.empty[s], 3);             // slot free in both CTAs once these MMAs have read it
          if (++s == P2_STAGES) { s = 0; ++round; }
        }
        umma_commit_cg2(&S.acc_full[a], 3);            // accumulator complete: wakes both CTAs' epilogue warps
      }
    }
  } else {
    // ===== epilogue warps (both CTAs): this CTA's 128 features x 256 tokens, TMEM -> registers -> global =====
    const int q = warp & 3;
    const int fl = q * 32 + lane;
    const int hsel = (warp - 2) >> 2;                  // warps 2-5: tokens [0,128), warps 6-9: tokens [128,256)
    const bool even = (lane & 1) == 0;
    const uint32_t acc_empty_leader = mapa_u32(smem_u32(&S.acc_empty[0]), 0);
    pdl_wait();                                        // the output buffer may still be read by the previous kernel
    if (p.zero_a) {                                    // side job while the first tile's main loop runs (see GemmArgs)
      const size_t nthr = (size_t)gridDim.x * (TC_THREADS - 64);
      const size_t gtid = (size_t)blockIdx.x * (TC_THREADS - 64) + (threadIdx.x - 64);
      for (size_t i = gtid; i < p.zero_na / 4; i += nthr) reinterpret_cast<float4*>(p.zero_a)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.zero_b) for (size_t i = gtid; i < p.zero_nb / 4; i += nthr) reinterpret_cast<float4*>(p.zero_b)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint32_t tl = 0;
    for (int tile = pair; tile < ntiles; tile += npairs, ++tl) {
      int fp, tt;
      tc2_tile(tile, nfp, mt, fp, tt);
      const int f = fp * 256 + (int)crank * 128 + fl, t0 = tt * P2_TOK;
      const uint32_t a = tl & 1;
      const float bias = (p.bias && f < p.N) ? p.bias[f] : 0.f;
      mbar_wait(&S.acc_full[a], (tl >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int ch = hsel * 8; ch < hsel * 8 + 8; ++ch) {
        const int c = ch * 16;
        if (t0 + c >= p.M) break;                      // warp-uniform
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + a * P2_TOK + (uint32_t)c, v);
        switch (p.epilogue) {
          case GE_F32: store_chunk<GE_F32>(p, v, t0 + c, f, bias, even); break;
          case GE_BF16: store_chunk<GE_BF16>(p, v, t0 + c, f, bias, even); break;
          case GE_BIAS_F32: store_chunk<GE_BIAS_F32>(p, v, t0 + c, f, bias, even); break;
          case GE_BIAS_GELU_BF16: store_chunk<GE_BIAS_GELU_BF16>(p, v, t0 + c, f, bias, even); break;
          default: store_chunk<GE_GEGLU_BF16>(p, v, t0 + c, f, bias, even); break;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc_empty_leader + 8u * a);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                                  // no MMA of the pair still reads this CTA's shared memory / TMEM
  trace_end(p.trace);
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// T5G_GEMM_PAIR=0 keeps the one-tile-per-CTA kernel for prefill shapes (A/B measurements)
bool use_tc2() {
  static const bool on = [] { const char* v = getenv("T5G_GEMM_PAIR"); return !v || atoi(v) != 0; }();
  return on;
}

cudaError_t launch_tc2(const CUtensorMap& mw, const CUtensorMap& mx, const TcParams& p, int num_sms, cudaStream_t st, bool pdl) {
  const size_t smem = sizeof(Tc2Smem) + 1024;
  static PerDeviceFlag attr_set;
  if (!attr_set.here()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set.here() = 1;
  }
  const int ntiles = ((p.N + 255) / 256) * ((p.M + P2_TOK - 1) / P2_TOK);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * (unsigned)std::min(num_sms / 2, ntiles));
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, gemm_tc2_kernel, mw, mx, p);
}

}  // namespace

bool gemm_tc_supported(const GemmArgs& a) {
  return a.K % TC_BK == 0 && a.K >= TC_BK && a.M >= 1 && a.N >= 1 && (reinterpret_cast<uintptr_t>(a.A) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(a.W) & 15) == 0 && (a.epilogue != GE_GEGLU_BF16 || a.N % 2 == 0);
}

// token tile (UMMA N) of a launch.  192 serves the voice-prompt prefill of one utterance (152 tokens): a 5-stage ring with
// 24 KB activation tiles instead of 4 stages with 32 KB tiles, a third of which would be padding
static int tc_tokt(int M) { return M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : M <= 128 ? 128 : M <= 192 ? 192 : 256; }

// split-K factor of a launch: K is cut over CTAs while the tile grid is smaller than the machine (weight-streaming regime)
static int tc_split(int M, int N, int K, int num_sms, int* kbps_out) {
  const int tokt = tc_tokt(M);
  const int kb_total = K / TC_BK;
  const int tiles = ((N + TC_BM - 1) / TC_BM) * ((M + tokt - 1) / tokt);
  int split = 1;
  while (split < 8 && tiles * split * 2 <= num_sms + num_sms / 4 && kb_total / (split * 2) >= 3) split *= 2;
  int kbps = (kb_total + split - 1) / split;
  while (split > 1 && (split - 1) * kbps >= kb_total) { split >>= 1; kbps = (kb_total + split - 1) / split; }
  if (kbps_out) *kbps_out = kbps;
  return split;
}

// true when this launch accumulates split-K partials with red.global.add into `out`: the caller may zero `out` itself
// (out_zeroed = 1, e.g. in the kernel that runs before) instead of paying a memset node that also breaks the PDL chain
bool gemm_tc_wants_zeroed_out(const GemmArgs& a, int num_sms) {
  return a.M > 0 && gemm_tc_supported(a) && a.epilogue == GE_F32 && tc_split(a.M, a.N, a.K, num_sms, nullptr) > 1;
}

cudaError_t launch_gemm_tc(const GemmArgs& a, cudaStream_t st, int num_sms, bool pdl) {
  if (a.M <= 0) return cudaSuccess;
  if (!gemm_tc_supported(a)) return cudaErrorNotSupported;
  const int tokt = tc_tokt(a.M);
  CUtensorMap mw, mx;
  if (!make_map_2d(&mw, a.W, a.N, a.K, a.K, TC_BM) || !make_map_2d(&mx, a.A, a.M, a.K, a.K, tokt)) return cudaErrorNotSupported;
  const int kb_total = a.K / TC_BK;
  const int tiles = ((a.N + TC_BM - 1) / TC_BM) * ((a.M + tokt - 1) / tokt);
  int kbps = 0;
  int split = tc_split(a.M, a.N, a.K, num_sms, &kbps);
  // plain fp32 outputs reduce fastest with red.global.add (measured: 17 vs 21 us at 64x2304x2304); the fused
  // epilogues (bias / GELU / GeGLU / bf16) need the full sum and use the cluster/DSMEM reduction instead
  const int atomic = (split > 1 && a.epilogue == GE_F32) ? 1 : 0;
  // cluster reduction: at one CTA per SM a GPC hosts two 8-CTA clusters, i.e. 16 on the chip -- more 8-CTA clusters than
  // that run as a second wave (fc1 of the batched head, 18 tiles x 8: 13.7 us); clusters of 4 fit in one wave
  if (!atomic && split == 8 && tiles > 16) { split = 4; kbps = (kb_total + split - 1) / split; }
  if (atomic && !a.out_zeroed) {
    cudaError_t e = cudaMemsetAsync(a.out, 0, sizeof(float) * ((size_t)(a.M - 1) * a.ldo + a.N), st);
    if (e != cudaSuccess) return e;
    pdl = false;
  }
  TcParams p{a.M, a.N, a.K, a.epilogue, a.bias, a.out, a.ldo, kbps, split, atomic, a.trace, a.probe, a.zero_a, a.zero_na, a.zero_b, a.zero_nb};
  dim3 grid((a.N + TC_BM - 1) / TC_BM, (a.M + tokt - 1) / tokt, split);
  if (tokt == 256 && split == 1 && use_tc2()) {            // prefill: persistent CTA pairs
    CUtensorMap mxh;
    if (!make_map_2d(&mxh, a.A, a.M, a.K, a.K, 128)) return cudaErrorNotSupported;
    return launch_tc2(mw, mxh, p, num_sms, st, pdl);
  }
  switch (tokt) {
    case 16: return launch_tc<16, 8>(mw, mx, p, grid, st, pdl);
    case 32: return launch_tc<32, 8>(mw, mx, p, grid, st, pdl);
    case 64: return launch_tc<64, 6>(mw, mx, p, grid, st, pdl);
    case 128: return launch_tc<128, 5>(mw, mx, p, grid, st, pdl);
    case 192: return launch_tc<192, 5>(mw, mx, p, grid, st, pdl);
    default: return launch_tc<256, 4>(mw, mx, p, grid, st, pdl);
  }
}

namespace tc {
EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}
bool make_map_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace tc
