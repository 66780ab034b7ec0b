// Persistent decode-layers kernel of the single-row (bs = 1) decode step.
//
// The 26 decoder layers of one step (models/t5gemma.py:183-243 PMDecoderLayer at q_len = 1: self-attention,
// cross-attention, GeGLU MLP inside the six-RMSNorm sandwich) run in ONE cooperative launch with one CTA per SM
// instead of 182 dependent kernels.  Why: at bs = 1 a layer streams 175 MB of bf16 weights (27 us at the measured
// 6.47 TB/s) through 8 dependent phases; as separate kernels every phase boundary drained the machine (a streaming
// kernel fills the register file, so the next kernel's CTAs -- and their weight prefetch -- could only start when it
// exited) and the step ran at 0.47 of the HBM roofline.  Here
//   * every warp walks a STATIC stream of weight units (one unit = a 2304-element K-segment of one weight row,
//     4.6 KB) that covers all six projections of all layers, and keeps a private ring of RING units in shared
//     memory filled with 16-byte cp.async copies (L2 evict-first).  The ring runs ahead of the math across phase
//     boundaries: while the CTA waits on a grid barrier or computes attention, the weights of the following phases
//     keep streaming (147 KB per SM, 21.8 MB chip-wide, in flight or landed);
//   * phases are separated by a ticket-counter grid barrier (release/acquire on one L2 word, ~1.2 us); the launch is
//     cooperative, so co-residency of the grid is guaranteed by the driver, not assumed;
//   * the residual stream h lives in registers, replicated in every CTA; the RMSNorm sandwich (HF:66-74; post-norm of
//     the previous sub-layer + residual + pre-norm of the next, one reduction pass) is recomputed redundantly by every
//     CTA from the 9 KB sub-layer output, so no extra phase exists for it;
//   * attention (HF:209-240, 274-314; models/t5gemma.py:85-172) runs on Hkv * ns CTAs as split-KV partials over the
//     paged pool (PM-RoPE of q and of the new k, softcap, window, in-place append), and the partials are merged by the
//     prologue of the following o-projection -- no separate merge phase.
// Data layout, rounding points (new K/V rounded to bf16 before use, K stored post-RoPE) and the split-K order inside a
// row are those of the multi-kernel path (gemv.cu / attention.cu), which stays in the library for 2-4 rows.
#include "kernels.h"
#include "attn_common.cuh"

namespace {

constexpr int DP_THREADS = 512;
constexpr int DP_WARPS = DP_THREADS / 32;
constexpr int DP_RING = 2;                       // units per warp kept in flight / landed
constexpr int DP_UNIT_CHUNKS = 288;              // 16-byte chunks per unit (9 per lane) = 2304 bf16
constexpr int DP_UNIT_BYTES = DP_UNIT_CHUNKS * 16;
constexpr int DP_NP = 8;                         // residual elements per thread: hidden <= 4096
constexpr int DP_BT = 256;                       // block-table entries cached per table
constexpr int DP_MAX_NS = 8;
constexpr int N_PHASE = 6;                       // weight phases per layer: qkv, o, q_cross, o_cross, gate|up, down

struct PhaseGeo { int n_tasks, rows_per_task, kseg, K; };

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, uint64_t pol) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Grid-wide barrier on a monotonically increasing ticket counter (never reset: every launch adds a multiple of the grid
// size).  `target` lives in thread 0.  A bounded wait (4 s of %globaltimer) turns a lost CTA into an error flag instead of
// a hung device.
__device__ __forceinline__ void grid_barrier(unsigned long long* counter, unsigned long long& target, int* err) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (target == 0) {
      const unsigned long long ticket = atomicAdd(counter, 1ULL);
      target = (ticket / gridDim.x + 1ULL) * gridDim.x;
    } else {
      red_release_add_u64(counter, 1ULL);
      target += gridDim.x;
    }
    unsigned spins = 0;
    unsigned long long t0 = 0;
    while (ld_acquire_u64(counter) < target) {
      if ((++spins & 0xfff) == 0) {
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000ULL) { if (err) atomicOr(err, 4); break; }
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// x vector in shared memory, split into lo/hi 16-byte halves of every 8-element chunk (conflict-free LDS.128)
struct XBuf {
  float4* lo; float4* hi;
  __device__ __forceinline__ XBuf(float* base, int K) { lo = reinterpret_cast<float4*>(base); hi = lo + (K >> 3); }
  __device__ __forceinline__ void store(int k, float v) {
    const int c = k >> 3, j = k & 7;
    reinterpret_cast<float*>((j < 4 ? lo : hi) + c)[j & 3] = v;
  }
};

// position of a warp in its static unit stream
struct Cursor {
  int l, p, j, u;
};

template <int G, int D>
__global__ void __launch_bounds__(DP_THREADS, 1) decode_layers_kernel(PersistArgs a) {
  constexpr int LPT = Geo<D>::LPT, DPL = Geo<D>::DPL, TPW = Geo<D>::TPW, NV = Geo<D>::NV;
  extern __shared__ __align__(16) unsigned char dp_smem[];
  unsigned char* ring = dp_smem;                                                   // [DP_WARPS][DP_RING][DP_UNIT_BYTES]
  float* xbase = reinterpret_cast<float*>(dp_smem + (size_t)DP_WARPS * DP_RING * DP_UNIT_BYTES);   // a.xbuf_floats floats
  __shared__ PersistLayer lay[T5G_PERSIST_MAX_LAYERS];
  __shared__ PhaseGeo geo[N_PHASE];
  __shared__ float red[128];
  __shared__ int bt_self[DP_BT], bt_cross[DP_BT];
  __shared__ float cs[D / 2], sn[D / 2];
  __shared__ float qs[G][D], knew[D], vnew[D];
  __shared__ float w_ml[DP_WARPS][G][2], w_wt[DP_WARPS][G], c_ml[G][2];
  __shared__ float m_wt[T5G_PERSIST_MAX_HEADS][DP_MAX_NS];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x, n_cta = gridDim.x;
  const int d = a.d, I = a.I, QD = a.QD, KD = a.KD, QKV = a.QKV;
  const uint64_t pol = l2_evict_first_policy();

  // ---- static tables (host-written before the launch: safe to read before the dependency resolves) ----
  for (int i = tid; i < a.n_layers * (int)(sizeof(PersistLayer) / 4); i += DP_THREADS)
    reinterpret_cast<int*>(lay)[i] = reinterpret_cast<const int*>(a.layers)[i];
  if (tid == 0) {
    auto ks = [](int K) { return ((K >> 3) + DP_UNIT_CHUNKS - 1) / DP_UNIT_CHUNKS; };
    geo[0] = PhaseGeo{QKV, 1, ks(d), d};
    geo[1] = PhaseGeo{d, 1, ks(QD), QD};
    geo[2] = PhaseGeo{QD, 1, ks(d), d};
    geo[3] = PhaseGeo{d, 1, ks(QD), QD};
    geo[4] = PhaseGeo{I, 2, ks(d), d};
    geo[5] = PhaseGeo{d, 1, ks(I), I};
  }
  for (int i = tid; i < DP_BT; i += DP_THREADS) {
    bt_self[i] = i < a.self_bt_stride ? a.self_bt[i] : 0;
    bt_cross[i] = i < a.cross_bt_stride ? a.cross_bt[i] : 0;
  }
  __syncthreads();

  // ---- the warp's unit stream ----
  auto task_of = [&](int j) { return cta + n_cta * (warp + DP_WARPS * j); };
  auto normalize = [&](Cursor& c) {          // move to the next existing (layer, phase, task) at or after c
    while (c.l < a.n_layers) {
      if (c.p < N_PHASE && task_of(c.j) < geo[c.p].n_tasks) return;
      c.j = 0; c.u = 0;
      if (++c.p >= N_PHASE) { c.p = 0; ++c.l; }
    }
  };
  auto advance = [&](Cursor& c) {
    const PhaseGeo g = geo[c.p];
    if (++c.u >= g.rows_per_task * g.kseg) { c.u = 0; ++c.j; }
    normalize(c);
  };
  auto weight_of = [&](int l, int p) -> const bf16* {
    const PersistLayer& L = lay[l];
    return p == 0 ? L.wqkv : p == 1 ? L.wo : p == 2 ? L.wq_c : p == 3 ? L.wo_c : p == 4 ? L.wgu : L.wd;
  };
  unsigned char* my_ring = ring + (size_t)warp * DP_RING * DP_UNIT_BYTES;
  auto issue = [&](const Cursor& c, int slot) {       // one commit group per unit (empty past the end of the stream)
    if (c.l < a.n_layers) {
      const PhaseGeo g = geo[c.p];
      const int row = task_of(c.j) * g.rows_per_task + c.u / g.kseg, seg = c.u % g.kseg;
      const int nvalid = min(DP_UNIT_CHUNKS, (g.K >> 3) - seg * DP_UNIT_CHUNKS);
      const bf16* src = weight_of(c.l, c.p) + (size_t)row * g.K + (size_t)seg * DP_UNIT_CHUNKS * 8;
      unsigned char* dst = my_ring + (size_t)slot * DP_UNIT_BYTES;
#pragma unroll
      for (int i = 0; i < DP_UNIT_CHUNKS / 32; ++i) {
        const int ch = lane + 32 * i;
        if (ch < nvalid) cp_async16(dst + ch * 16, src + ch * 8, pol);
      }
    }
    cp_async_commit();
  };
  Cursor pre{0, 0, 0, 0};
  normalize(pre);
  int n_issued = 0, n_consumed = 0;
#pragma unroll
  for (int r = 0; r < DP_RING; ++r) { issue(pre, n_issued % DP_RING); ++n_issued; if (pre.l < a.n_layers) advance(pre); }

  pdl_wait();
  // dependents (the head of the next step) may only become resident once the sampler of THIS step has completed: they
  // read the slot state before their own griddepcontrol.wait
  pdl_launch_dependents();
  trace_begin(a.trace);
  // ---- state written by this step's sampler ----
  const SlotDev& sl = a.slots[0];
  const int active = sl.active;
  if (!active) { cp_async_wait<0>(); return; }     // uniform over the grid: no barrier has been touched
  const int L_self = sl.cur_len, L_cross = sl.n_text, last_token = sl.last_token;
  for (int i = tid; i < D / 2; i += DP_THREADS) { cs[i] = a.rope_cs[i]; sn[i] = a.rope_cs[D / 2 + i]; }
  unsigned long long bar_target = 0;
  unsigned long long* probe = (a.probe && cta == 0 && tid == 0) ? a.probe : nullptr;
  int probe_i = 0;
#define DP_PROBE(l_) do { if (probe && (l_) == a.probe_layer && probe_i < 32) probe[probe_i++] = globaltimer_ns(); } while (0)

  // consume all tasks of phase p of layer l that belong to this warp; EPI: 0 store, 1 GeGLU
  auto run_phase = [&](int l, int p, float* out) {
    const PhaseGeo g = geo[p];
    const XBuf xs(xbase, g.K);
    const int upt = g.rows_per_task * g.kseg;
    for (int j = 0; task_of(j) < g.n_tasks; ++j) {
      const int task = task_of(j);
      float acc0 = 0.f, acc1 = 0.f;
      for (int u = 0; u < upt; ++u) {
        cp_async_wait<DP_RING - 1>();
        __syncwarp();
        const uint4* slot = reinterpret_cast<const uint4*>(my_ring + (size_t)(n_consumed % DP_RING) * DP_UNIT_BYTES);
        const int seg = u % g.kseg, cbase = seg * DP_UNIT_CHUNKS;
        const int nvalid = min(DP_UNIT_CHUNKS, (g.K >> 3) - cbase);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < DP_UNIT_CHUNKS / 32; ++i) {
          const int ch = lane + 32 * i;
          if (ch < nvalid) {
            float wf[8];
            bf16x8_to_f32(slot[ch], wf);
            const float4 x0 = xs.lo[cbase + ch], x1 = xs.hi[cbase + ch];
            acc = fmaf(wf[0], x0.x, acc); acc = fmaf(wf[1], x0.y, acc); acc = fmaf(wf[2], x0.z, acc); acc = fmaf(wf[3], x0.w, acc);
            acc = fmaf(wf[4], x1.x, acc); acc = fmaf(wf[5], x1.y, acc); acc = fmaf(wf[6], x1.z, acc); acc = fmaf(wf[7], x1.w, acc);
          }
        }
        if (u < g.kseg) acc0 += acc; else acc1 += acc;
        __syncwarp();                                  // every lane is done with the slot before it is refilled
        issue(pre, n_consumed % DP_RING);
        ++n_issued; ++n_consumed;
        if (pre.l < a.n_layers) advance(pre);
      }
      acc0 = warp_sum(acc0);
      if (g.rows_per_task == 2) {
        acc1 = warp_sum(acc1);
        if (lane == 0) out[task] = gelu_tanh_f(acc0) * acc1;
      } else if (lane == 0) {
        out[task] = acc0;
      }
    }
  };

  // residual stream, replicated per CTA: thread t holds h[t + 512 i]
  float hreg[DP_NP];
  // h += rmsnorm(y) * g_post (when y) ; xbuf = rmsnorm(h) * g_pre
  auto sandwich = [&](const float* y, const float* g_post, const float* g_pre) {
    XBuf xs(xbase, d);
    float gp[DP_NP], gq[DP_NP], yv[DP_NP];
#pragma unroll
    for (int i = 0; i < DP_NP; ++i) {
      const int k = tid + i * DP_THREADS;
      gp[i] = (k < d) ? g_pre[k] : 0.f;
      gq[i] = (y && k < d) ? g_post[k] : 0.f;
      yv[i] = (y && k < d) ? __ldcg(y + k) : 0.f;
    }
    float s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
#pragma unroll
    for (int i = 0; i < DP_NP; ++i) {
      const float yg = yv[i] * gq[i];
      s1 = fmaf(yv[i], yv[i], s1); s2 = fmaf(hreg[i], hreg[i], s2);
      s3 = fmaf(hreg[i], yg, s3); s4 = fmaf(yg, yg, s4);
      yv[i] = yg;
    }
    block_sum4(s1, s2, s3, s4, red);
    float ss = s2;
    if (y) {
      const float ry = rsqrtf(s1 / (float)d + a.eps);
#pragma unroll
      for (int i = 0; i < DP_NP; ++i) hreg[i] = fmaf(yv[i], ry, hreg[i]);
      ss = s2 + 2.f * ry * s3 + ry * ry * s4;
    }
    const float rinv = rsqrtf(ss / (float)d + a.eps);
#pragma unroll
    for (int i = 0; i < DP_NP; ++i) {
      const int k = tid + i * DP_THREADS;
      if (k < d) xs.store(k, hreg[i] * rinv * gp[i]);
    }
    __syncthreads();
  };

  // split-KV attention partial of (kv head hk, split) on this CTA; result to part_o / part_ml
  auto attention = [&](int l, bool is_cross, const float* q, const float* kv_new, int ns) {
    const int hk = cta % a.Hkv, split = cta / a.Hkv;
    const int* bt = is_cross ? bt_cross : bt_self;
    const int* bt_g = is_cross ? a.cross_bt : a.self_bt;
    const int PT = a.pool.page_tokens;
    const int Lk = is_cross ? L_cross : L_self;
    const int window = (!is_cross && lay[l].sliding) ? a.window : 0;
    const int lo = window > 0 ? max(0, Lk - window) : 0;
    int chunk = (Lk - lo + ns - 1) / ns;
    chunk = (chunk + TPW - 1) / TPW * TPW;
    const int t_begin = lo + split * chunk, t_end = min(Lk, t_begin + chunk);
    const bool has_new = (!is_cross) && (t_end == Lk) && (t_end > t_begin);
    const int grp = lane / LPT, l8 = lane % LPT;
    auto page_of = [&](int t) -> int { const int pi = t / PT; return pi < DP_BT ? bt[pi] : bt_g[pi]; };
    // producer outputs: raw q of the group's heads and the new k/v
    for (int i = tid; i < G * D; i += DP_THREADS) qs[i / D][i % D] = __ldcg(q + (size_t)(hk * G) * D + i);
    if (has_new)
      for (int j = tid; j < D; j += DP_THREADS) {
        knew[j] = __ldcg(kv_new + (size_t)hk * D + j);
        vnew[j] = __bfloat162float(__float2bfloat16(__ldcg(kv_new + (size_t)(a.Hkv + hk) * D + j)));
      }
    __syncthreads();
    for (int i = tid; i < G * D / 2; i += DP_THREADS) {           // PM-RoPE, half-split pairs (j, j + D/2)
      const int g = i / (D / 2), j = i - g * (D / 2);
      const float x1 = qs[g][j], x2 = qs[g][j + D / 2];
      qs[g][j] = x1 * cs[j] - x2 * sn[j];
      qs[g][j + D / 2] = x2 * cs[j] + x1 * sn[j];
    }
    if (has_new)
      for (int j = tid; j < D / 2; j += DP_THREADS) {
        const float x1 = knew[j], x2 = knew[j + D / 2];
        knew[j] = __bfloat162float(__float2bfloat16(x1 * cs[j] - x2 * sn[j]));
        knew[j + D / 2] = __bfloat162float(__float2bfloat16(x2 * cs[j] + x1 * sn[j]));
      }
    __syncthreads();
    if (has_new) {                                                  // append (K post-RoPE), visible to later steps
      const int t = Lk - 1, page = page_of(t), off = t % PT;
      bf16* kd = a.pool.ptr(l, 0, page) + ((size_t)hk * PT + off) * D;
      bf16* vd = a.pool.ptr(l, 1, page) + ((size_t)hk * PT + off) * D;
      for (int j = tid; j < D; j += DP_THREADS) { kd[j] = __float2bfloat16(knew[j]); vd[j] = __float2bfloat16(vnew[j]); }
    }
    float qreg[G][DPL];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int i = 0; i < DPL; ++i) qreg[g][i] = qs[g][l8 * DPL + i];
    GroupState<G, DPL> st;
    st.init();
    for (int t0 = t_begin + warp * TPW; t0 < t_end; t0 += 2 * DP_WARPS * TPW) {
      uint4 ku[2][NV], vu[2][NV];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int t = t0 + u * DP_WARPS * TPW + grp;
        if (t < t_end && !(has_new && t == Lk - 1)) {
          const int page = page_of(t), off = t % PT;
          const bf16* kp = a.pool.ptr(l, 0, page) + ((size_t)hk * PT + off) * D + l8 * DPL;
          const bf16* vp = a.pool.ptr(l, 1, page) + ((size_t)hk * PT + off) * D + l8 * DPL;
#pragma unroll
          for (int i = 0; i < NV; ++i) { ku[u][i] = *reinterpret_cast<const uint4*>(kp + i * 8); vu[u][i] = *reinterpret_cast<const uint4*>(vp + i * 8); }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (t0 + u * DP_WARPS * TPW >= t_end) break;               // warp-uniform
        const int t = t0 + u * DP_WARPS * TPW + grp;
        const bool valid = t < t_end, fresh = valid && has_new && t == Lk - 1;
        float kf[DPL], vf[DPL];
        if (fresh) {
#pragma unroll
          for (int i = 0; i < DPL; ++i) { kf[i] = knew[l8 * DPL + i]; vf[i] = vnew[l8 * DPL + i]; }
        } else if (valid) {
#pragma unroll
          for (int i = 0; i < NV; ++i) { bf16x8_to_f32(ku[u][i], kf + i * 8); bf16x8_to_f32(vu[u][i], vf + i * 8); }
        } else {
#pragma unroll
          for (int i = 0; i < DPL; ++i) { kf[i] = 0.f; vf[i] = 0.f; }
        }
        group_update<G, D>(st, qreg, kf, vf, a.scale, a.softcap, valid);
      }
    }
    warp_merge<G, D>(st);
    float* w_o = xbase;                                             // [DP_WARPS][G][D] (the x buffer is dead here)
    if (grp == 0) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (l8 == 0) { w_ml[warp][g][0] = st.m[g]; w_ml[warp][g][1] = st.l[g]; }
#pragma unroll
        for (int i = 0; i < DPL; ++i) w_o[((size_t)warp * G + g) * D + l8 * DPL + i] = st.acc[g][i];
      }
    }
    __syncthreads();
    if (tid < G) {
      float M = -INFINITY;
#pragma unroll
      for (int w = 0; w < DP_WARPS; ++w) M = fmaxf(M, w_ml[w][tid][0]);
      float den = 0.f;
#pragma unroll
      for (int w = 0; w < DP_WARPS; ++w) {
        const float m = w_ml[w][tid][0];
        const float wt = (m == -INFINITY) ? 0.f : __expf(m - M);
        den = fmaf(wt, w_ml[w][tid][1], den);
        w_wt[w][tid] = wt;
      }
      c_ml[tid][0] = M; c_ml[tid][1] = den;
    }
    __syncthreads();
    for (int i = tid; i < G * D; i += DP_THREADS) {
      const int g = i / D, dd = i - g * D;
      float num = 0.f;
#pragma unroll
      for (int w = 0; w < DP_WARPS; ++w) num = fmaf(w_wt[w][g], w_o[((size_t)w * G + g) * D + dd], num);
      a.part_o[((size_t)(hk * G + g) * DP_MAX_NS + split) * D + dd] = num;
    }
    if (tid < G) {
      a.part_ml[((hk * G + tid) * DP_MAX_NS + split) * 2] = c_ml[tid][0];
      a.part_ml[((hk * G + tid) * DP_MAX_NS + split) * 2 + 1] = c_ml[tid][1];
    }
  };

  // xbuf[0..QD) = merged attention output (every CTA, redundantly)
  auto merge_partials = [&](int ns) {
    XBuf xs(xbase, QD);
    if (tid < a.Hq) {
      float M = -INFINITY, m[DP_MAX_NS], lv[DP_MAX_NS];
#pragma unroll
      for (int s = 0; s < DP_MAX_NS; ++s) {
        m[s] = s < ns ? __ldcg(a.part_ml + (tid * DP_MAX_NS + s) * 2) : -INFINITY;
        lv[s] = s < ns ? __ldcg(a.part_ml + (tid * DP_MAX_NS + s) * 2 + 1) : 0.f;
        M = fmaxf(M, m[s]);
      }
      float den = 0.f, wt[DP_MAX_NS];
#pragma unroll
      for (int s = 0; s < DP_MAX_NS; ++s) { wt[s] = (m[s] == -INFINITY) ? 0.f : __expf(m[s] - M); den = fmaf(wt[s], lv[s], den); }
      const float inv = den > 0.f ? 1.f / den : 0.f;
#pragma unroll
      for (int s = 0; s < DP_MAX_NS; ++s) m_wt[tid][s] = wt[s] * inv;
    }
    __syncthreads();
    for (int i = tid; i < QD; i += DP_THREADS) {
      const int h = i / D, dd = i - h * D;
      float v[DP_MAX_NS];
#pragma unroll
      for (int s = 0; s < DP_MAX_NS; ++s) v[s] = s < ns ? __ldcg(a.part_o + ((size_t)h * DP_MAX_NS + s) * D + dd) : 0.f;
      float o = 0.f;
#pragma unroll
      for (int s = 0; s < DP_MAX_NS; ++s) o = fmaf(m_wt[h][s], v[s], o);
      xs.store(i, o);
    }
    __syncthreads();
  };

  auto n_splits = [&](int keys, int ns_max) { return max(1, min(ns_max, (keys + a.keys_per_split - 1) / a.keys_per_split)); };
  const int ns_cross = n_splits(L_cross, a.ns_max);

  // ---- layer 0 input: audio embedding * sqrt(d) (models/t5gemma.py:1083; HF:769) ----
#pragma unroll
  for (int i = 0; i < DP_NP; ++i) {
    const int k = tid + i * DP_THREADS;
    hreg[i] = (k < d) ? __bfloat162float(a.emb[(size_t)last_token * d + k]) * a.emb_scale : 0.f;
  }

  for (int l = 0; l < a.n_layers; ++l) {
    const PersistLayer& Ly = lay[l];
    DP_PROBE(l);
    // P0: qkv = Wqkv . pre_sa(h + post_ff(y of the previous layer))
    if (l == 0) sandwich(nullptr, nullptr, Ly.g_pre_sa);
    else sandwich(a.y, lay[l - 1].g_post_ff, Ly.g_pre_sa);
    DP_PROBE(l);
    run_phase(l, 0, a.qkv);
    DP_PROBE(l);
    grid_barrier(a.barrier, bar_target, a.err);
    DP_PROBE(l);
    // P1: self-attention partials
    const int win = Ly.sliding ? a.window : 0;
    const int keys_self = win > 0 ? min(L_self, win) : L_self;
    const int ns_self = n_splits(keys_self, a.ns_max);
    if (cta < a.Hkv * ns_self) attention(l, false, a.qkv, a.qkv + QD, ns_self);
    DP_PROBE(l);
    grid_barrier(a.barrier, bar_target, a.err);
    DP_PROBE(l);
    // P2: y = Wo . attn
    merge_partials(ns_self);
    run_phase(l, 1, a.y);
    DP_PROBE(l);
    grid_barrier(a.barrier, bar_target, a.err);
    DP_PROBE(l);
    // P3: qc = Wq_c . pre_ca(h + post_sa(y))
    sandwich(a.y, Ly.g_post_sa, Ly.g_pre_ca);
    run_phase(l, 2, a.qc);
    DP_PROBE(l);
    grid_barrier(a.barrier, bar_target, a.err);
    DP_PROBE(l);
    // P4: cross-attention partials over the encoder pages
    if (cta < a.Hkv * ns_cross) attention(l, true, a.qc, nullptr, ns_cross);
    DP_PROBE(l);
    grid_barrier(a.barrier, bar_target, a.err);
    DP_PROBE(l);
    // P5: y = Wo_c . attn
    merge_partials(ns_cross);
    run_phase(l, 3, a.y);
    DP_PROBE(l);
    grid_barrier(a.barrier, bar_target, a.err);
    DP_PROBE(l);
    // P6: act = GeGLU(Wgu . pre_ff(h + post_ca(y)))
    sandwich(a.y, Ly.g_post_ca, Ly.g_pre_ff);
    run_phase(l, 4, a.act);
    DP_PROBE(l);
    grid_barrier(a.barrier, bar_target, a.err);
    DP_PROBE(l);
    // P7: y = Wd . act
    {
      XBuf xs(xbase, I);
      const float4* av = reinterpret_cast<const float4*>(a.act);
      for (int i = tid; i < (I >> 2); i += DP_THREADS) ((i & 1) ? xs.hi : xs.lo)[i >> 1] = __ldcg(av + i);
      __syncthreads();
    }
    run_phase(l, 5, a.y);
    DP_PROBE(l);
    if (l + 1 < a.n_layers) grid_barrier(a.barrier, bar_target, a.err);
    DP_PROBE(l);
  }
  cp_async_wait<0>();
  // the head kernel applies post_ff of the last layer: hand it h (CTA 0) and y (already in a.y)
  if (cta == 0) {
#pragma unroll
    for (int i = 0; i < DP_NP; ++i) {
      const int k = tid + i * DP_THREADS;
      if (k < d) a.h_out[k] = hreg[i];
    }
  }
  trace_end(a.trace);
#undef DP_PROBE
}

template <int G, int D>
cudaError_t launch_gd(const PersistArgs& a, int num_sms, cudaStream_t st, bool pdl, size_t smem) {
  auto kern = decode_layers_kernel<G, D>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(num_sms);
  cfg.blockDim = dim3(DP_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, a);
}

size_t smem_bytes(const PersistArgs& a) {
  return (size_t)DP_WARPS * DP_RING * DP_UNIT_BYTES + (size_t)a.xbuf_floats * sizeof(float);
}

}  // namespace

int decode_persist_xbuf_floats(int d, int I, int QD, int G, int D) {
  int m = d > I ? d : I;
  if (QD > m) m = QD;
  const int attn = DP_WARPS * G * D;          // per-warp attention partials alias the x buffer
  if (attn > m) m = attn;
  return (m + 7) & ~7;
}

bool decode_persist_supported(int d, int I, int Hq, int Hkv, int D, int n_layers, int num_sms) {
  if (Hkv <= 0 || Hq % Hkv) return false;
  const int G = Hq / Hkv;
  if (G != 2 || !(D == 16 || D == 32 || D == 64 || D == 128 || D == 256)) return false;
  if (d % 8 || I % 8 || d > DP_THREADS * DP_NP) return false;
  if (n_layers > T5G_PERSIST_MAX_LAYERS || Hq > T5G_PERSIST_MAX_HEADS || Hkv > num_sms) return false;
  const size_t smem = (size_t)DP_WARPS * DP_RING * DP_UNIT_BYTES + (size_t)decode_persist_xbuf_floats(d, I, Hq * D, G, D) * 4;
  return smem + 16 * 1024 <= 227 * 1024;
}

cudaError_t launch_decode_persist(const PersistArgs& a, int num_sms, cudaStream_t st, bool pdl) {
  const int G = a.Hq / a.Hkv;
  const size_t smem = smem_bytes(a);
  if (G == 2) {
    switch (a.D) {
      case 16: return launch_gd<2, 16>(a, num_sms, st, pdl, smem);
      case 32: return launch_gd<2, 32>(a, num_sms, st, pdl, smem);
      case 64: return launch_gd<2, 64>(a, num_sms, st, pdl, smem);
      case 128: return launch_gd<2, 128>(a, num_sms, st, pdl, smem);
      case 256: return launch_gd<2, 256>(a, num_sms, st, pdl, smem);
    }
  }
  return cudaErrorInvalidValue;
}
