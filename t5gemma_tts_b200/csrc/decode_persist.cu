// Persistent decode-layers kernel of the single-row (bs = 1) decode step.
//
// The 26 decoder layers of one step (models/t5gemma.py:183-243 PMDecoderLayer at q_len = 1: self-attention,
// cross-attention, GeGLU MLP inside the six-RMSNorm sandwich) run in ONE cooperative launch with one CTA per SM
// instead of 182 dependent kernels.  Why: at bs = 1 a layer streams 175 MB of bf16 weights (27 us at the measured
// 6.47 TB/s) through 8 dependent phases; as separate kernels every phase boundary drained the machine (a streaming
// kernel fills the register file, so the next kernel's CTAs -- and their weight prefetch -- could only start when it
// exited) and the step ran at 0.47 of the HBM roofline.  Here
//   * every CTA walks a STATIC stream of weight units (one unit = a 2304-element K-segment of one weight row,
//     4.6 KB) that covers all six projections of all layers.  A dedicated producer warp copies the stream into a
//     ~36-slot shared-memory ring with 16-byte cp.async (L2 evict-first, completion on per-slot mbarriers) and runs
//     ahead of the math across phase boundaries: while the consumers wait on a grid barrier or compute attention,
//     the weights of the following phases keep arriving (~165 KB per SM, 24 MB chip-wide).  The number of units IN
//     FLIGHT is capped separately (Little: bytes in flight beyond bandwidth x latency only queue in the memory
//     system and inflate the latency of the phase-critical loads and of the barrier word);
//   * the 16 consumer warps keep their slice of the activation vector in registers for a whole phase, so a unit
//     costs one shared-memory read of the weights and nothing else (x re-reads from shared memory capped the first
//     version at the shared-memory bandwidth);
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

constexpr int DP_CONS_WARPS = 12;                // consumer warps: prologues, dot products, attention, barriers (512 threads
                                                 // in all, so every thread may use 128 registers: 72 of them hold x)
constexpr int DP_PROD_WARPS = 4;                 // producer warps: only issue cp.async weight copies (one warp alone issues
                                                 // ~15 GB/s: 2.2 TB/s chip-wide, measured)
constexpr int DP_CONS = DP_CONS_WARPS * 32;
constexpr int DP_THREADS = DP_CONS + 32 * DP_PROD_WARPS;
constexpr int DP_UNIT_CHUNKS = 288;              // 16-byte chunks per unit (9 per lane) = 2304 bf16
constexpr int DP_UPL = DP_UNIT_CHUNKS / 32;
constexpr int DP_UNIT_BYTES = DP_UNIT_CHUNKS * 16;
constexpr int DP_MAX_SLOTS = 64;
constexpr int DP_NP = 6;                         // residual elements per thread: hidden <= 2304 (wider models use the multi-kernel step)
constexpr int DP_BT = 256;                       // block-table entries cached per table
constexpr int DP_MAX_NS = 8;
constexpr int DP_PART = 1024;                    // split-K partials per CTA
constexpr int N_PHASE = 6;                       // weight phases per layer: qkv, o, q_cross, o_cross, gate|up, down

// per-phase constants of this CTA, computed once (integer divisions are ~40 dependent instructions each: none may sit on
// the per-unit path of a warp that has nothing else to hide latency with)
struct PhaseGeo {
  int n_tasks, rows_per_task, kseg, K;
  int nt_cta;            // tasks of this CTA
  int ng;                // warp groups (DP_CONS_WARPS / kseg)
  int units;             // units of this CTA in the phase = nt_cta * rows_per_task * kseg
  int umod, upar;        // units % n_slots, (units / n_slots) & 1
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {     // arrives once this thread's earlier copies have landed
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "DP_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DP_DONE;\n\t"
      "bra DP_WAIT;\n\t"
      "DP_DONE:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
// bounded wait: a lost arrival turns into an error code (site) instead of a hung device
__device__ __forceinline__ void mbar_wait_checked(uint64_t* b, uint32_t parity, int* err, int site) {
  for (unsigned spins = 0;; ++spins) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (done) return;
    if (spins > (1u << 20)) { if (err) atomicCAS(err, 0, 4 | (site << 20)); return; }
  }
}
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(DP_CONS) : "memory"); }

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- flag-in-data exchange between CTAs -------------------------------------------------------------------------------
// Every value a phase hands to the next one travels as an 8-byte (value, tag) pair written with ONE 64-bit store (single-
// copy atomic); the tag encodes (launch epoch, layer, stage).  A reader polls the pair itself until the tag matches, so an
// exchange costs one L2 round trip after the producer's store -- no fence, no atomic, no grid barrier (a ticket-counter
// barrier cost 1.2-2.4 us per phase here: membar + atomic + poll round trip + two CTA barriers, 8 times per layer).
// Buffer reuse is safe without barriers because every stage's output depends on the previous stage's outputs of ALL CTAs.
__device__ __forceinline__ void st_tag(uint2* p, float v, unsigned tag) {
  asm volatile("st.global.cg.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 ld_tag(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.cg.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  return r;
}
// B tagged elements per thread (element k = k0 + i * stride, valid while k < n): all loads of a round are in flight together.
// Waiting must not flood L2: 57 K spinning threads saturate its request bandwidth and starve the weight stream and the
// very stores they wait for (measured: phases slowed down from microseconds to tens of milliseconds).  So only lane 0 of a
// warp polls (its own first element, with a short sleep between polls) until that shows up; then every lane reads its
// elements and re-reads, again with a sleep, only if some tag is still missing.
template <int B>
__device__ __forceinline__ void ld_tag_vec(const uint2* src, int k0, int stride, int n, unsigned tag, float* out, int* err, int site = 0) {
  if ((threadIdx.x & 31) == 0 && k0 < n) {
    // every CTA polls a different element: 148 CTAs spinning on the same dozen addresses load single L2 slices enough to
    // slow the weight stream that is striped over all of them
    const int kp = (stride == DP_CONS) ? (int)((blockIdx.x * 41u + (threadIdx.x >> 5) * 32u) % (unsigned)n) : k0;   // (dense vectors only)
    unsigned spins = 0;
    while (ld_tag(src + kp).y != tag) {
      if (++spins > (1u << 22)) { if (err) atomicCAS(err, 0, 4 | ((tag & 0x3ff) << 8) | (site << 20)); break; }
    }
  }
  __syncwarp();
  uint2 r[B];
  unsigned miss = 0;
#pragma unroll
  for (int i = 0; i < B; ++i) {
    const int k = k0 + i * stride;
    r[i] = (k < n) ? ld_tag(src + k) : make_uint2(0u, tag);
  }
#pragma unroll
  for (int i = 0; i < B; ++i) miss |= (r[i].y != tag) ? (1u << i) : 0u;
  unsigned spins = 0;
  while (miss) {                                   // only the elements whose producer is late are read again
#pragma unroll
    for (int i = 0; i < B; ++i)
      if (miss & (1u << i)) {
        r[i] = ld_tag(src + k0 + i * stride);
        if (r[i].y == tag) miss &= ~(1u << i);
      }
    if (++spins > (1u << 22)) { if (err) atomicCAS(err, 0, 4 | ((tag & 0x3ff) << 8) | (site << 20) | (1 << 24)); break; }
  }
#pragma unroll
  for (int i = 0; i < B; ++i) out[i] = __uint_as_float(r[i].x);
}

// Grid-wide barrier of the consumer threads on a monotonically increasing ticket counter (never reset: every launch adds
// a multiple of the grid size).  Release/acquire at gpu scope on one word; the data exchanged between phases is read
// with ld.global.cg, so no L1 invalidation is needed.  `target` lives in thread 0.  A bounded wait (4 s of %globaltimer)
// turns a lost CTA into an error flag instead of a hung device.
__device__ __forceinline__ void grid_barrier(unsigned long long* counter, unsigned long long& target, int* err) {
  cons_sync();
  if (threadIdx.x == 0) {
    if (target == 0) {
      __threadfence();
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
  }
  cons_sync();
}

// sum of four values over the consumer threads
__device__ __forceinline__ void cons_sum4(float& a, float& b, float& c, float& d, float* red /* >= 128 floats */) {
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c); d = warp_sum(d);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  cons_sync();
  if (lane == 0) { red[w] = a; red[32 + w] = b; red[64 + w] = c; red[96 + w] = d; }
  cons_sync();
  a = warp_sum(lane < DP_CONS_WARPS ? red[lane] : 0.f); b = warp_sum(lane < DP_CONS_WARPS ? red[32 + lane] : 0.f);
  c = warp_sum(lane < DP_CONS_WARPS ? red[64 + lane] : 0.f); d = warp_sum(lane < DP_CONS_WARPS ? red[96 + lane] : 0.f);
}

template <int G, int D>
__global__ void __launch_bounds__(DP_THREADS, 1) decode_layers_kernel(PersistArgs a) {
  constexpr int LPT = Geo<D>::LPT, DPL = Geo<D>::DPL, TPW = Geo<D>::TPW, NV = Geo<D>::NV;
  extern __shared__ __align__(16) unsigned char dp_smem[];
  unsigned char* ring = dp_smem;                                                   // [n_slots][DP_UNIT_BYTES]
  float* xbuf = reinterpret_cast<float*>(dp_smem + (size_t)a.n_slots * DP_UNIT_BYTES);   // a.xbuf_floats floats
  __shared__ PersistLayer lay[T5G_PERSIST_MAX_LAYERS];
  __shared__ PhaseGeo geo[N_PHASE];
  __shared__ __align__(8) uint64_t full_bar[DP_MAX_SLOTS], empty_bar[DP_MAX_SLOTS];
  __shared__ __align__(16) float red[128], part[DP_PART];
  __shared__ int bt_self[DP_BT], bt_cross[DP_BT];
  __shared__ __align__(16) float cs[D / 2], sn[D / 2];
  __shared__ __align__(16) float qs[G][D], knew[D], vnew[D];      // read with 16-byte loads
  __shared__ float w_ml[DP_CONS_WARPS][G][2], w_wt[DP_CONS_WARPS][G], c_ml[G][2];
  __shared__ float h_park[DP_NP * DP_CONS];            // the residual registers are parked here while a CTA runs attention

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x, n_cta = gridDim.x;
  const int d = a.d, I = a.I, QD = a.QD, QKV = a.QKV;
  const int NS = a.n_slots;

  // ---- static tables (host-written before the launch: safe to read before the dependency resolves) ----
  for (int i = tid; i < a.n_layers * (int)(sizeof(PersistLayer) / 4); i += DP_THREADS)
    reinterpret_cast<int*>(lay)[i] = reinterpret_cast<const int*>(a.layers)[i];
  if (tid == 0) {
    auto ks = [](int K) { return ((K >> 3) + DP_UNIT_CHUNKS - 1) / DP_UNIT_CHUNKS; };
    geo[0] = PhaseGeo{QKV, 1, ks(d), d, 0, 0, 0, 0, 0};
    geo[1] = PhaseGeo{d, 1, ks(QD), QD, 0, 0, 0, 0, 0};
    geo[2] = PhaseGeo{QD, 1, ks(d), d, 0, 0, 0, 0, 0};
    geo[3] = PhaseGeo{d, 1, ks(QD), QD, 0, 0, 0, 0, 0};
    geo[4] = PhaseGeo{I, 2, ks(d), d, 0, 0, 0, 0, 0};
    geo[5] = PhaseGeo{d, 1, ks(I), I, 0, 0, 0, 0, 0};
    for (int p = 0; p < N_PHASE; ++p) {
      PhaseGeo& g = geo[p];
      g.nt_cta = g.n_tasks > cta ? (g.n_tasks - cta + n_cta - 1) / n_cta : 0;
      g.ng = DP_CONS_WARPS / g.kseg;
      g.units = g.nt_cta * g.rows_per_task * g.kseg;
      g.umod = g.units % NS; g.upar = (g.units / NS) & 1;
    }
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 32); mbar_init(&empty_bar[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < DP_BT; i += DP_THREADS) {
    bt_self[i] = i < a.self_bt_stride ? a.self_bt[i] : 0;
    bt_cross[i] = i < a.cross_bt_stride ? a.cross_bt[i] : 0;
  }
  __syncthreads();
  auto weight_of = [&](int l, int p) -> const bf16* {
    const PersistLayer& L = lay[l];
    return p == 0 ? L.wqkv : p == 1 ? L.wo : p == 2 ? L.wq_c : p == 3 ? L.wo_c : p == 4 ? L.wgu : L.wd;
  };

  // =====================================================================================================================
  // producer warp: walks the CTA's static unit stream (layer, phase, task, row of the task, K-segment) and copies every
  // unit into the next ring slot; at most a.max_inflight units are in flight (more would only queue in the memory system
  // and inflate the latency of the phase-critical loads and barrier words), the rest of the ring holds landed units
  // =====================================================================================================================
  const bool dbg_nomath = a.dbg & 1, dbg_nostream = a.dbg & 2;
  if (warp >= DP_CONS_WARPS) {
    const int pw = warp - DP_CONS_WARPS;               // units n = pw (mod DP_PROD_WARPS) are this warp's
    if (dbg_nostream) return;
    const uint64_t pol = l2_evict_first_policy();
    const uint32_t ring_u32 = smem_u32(ring);
    // this warp's units are n = pw, pw + 4, ...; slot / parity of the unit and of the in-flight reference advance without
    // divisions (n_slots is a multiple of 4, so both stay congruent to pw)
    int n = 0;
    bool waited = false;
    const int F = max(1, a.max_inflight / DP_PROD_WARPS) * DP_PROD_WARPS;   // units between this warp's issue and its own earlier unit
    int slot = pw, par = 0;                    // of unit n_mine (next unit of this warp)
    int fslot = pw, fpar = 0, n_mine = pw;     // of unit n_mine - F once n_mine >= F
#pragma unroll 1
    for (int l = 0; l < a.n_layers; ++l)
#pragma unroll 1
      for (int p = 0; p < N_PHASE; ++p) {
        const PhaseGeo g = geo[p];
        const bf16* W = weight_of(l, p);
        const int nch = g.K >> 3, upt = g.rows_per_task * g.kseg;
        // unit index inside the phase: q = i * upt + r * kseg + seg ; walk q = first unit of this warp, += 4
        int q = (pw - n % DP_PROD_WARPS + DP_PROD_WARPS) % DP_PROD_WARPS;
        int i = q / upt, rem = q - i * upt;
#pragma unroll 1
        for (; q < g.units; q += DP_PROD_WARPS) {
          if (!waited && n_mine >= NS) {
            // the first ring-full of weights needs nothing from this step; everything later needs a free slot, i.e. a
            // running consumer, so the slot state (written by this step's sampler) is checked first
            pdl_wait();
            waited = true;
            if (!a.slots[0].active) { cp_async_wait_all(); return; }
          }
          const int r = rem / g.kseg, seg = rem - r * g.kseg;      // kseg, upt are tiny: these stay cheap only because
                                                                   // the producer warps are off the critical path
          mbar_wait_checked(&empty_bar[slot], par ^ 1, a.err, 10);
          if (n_mine >= F) {
            mbar_wait_checked(&full_bar[fslot], fpar, a.err, 11);
            fslot += DP_PROD_WARPS; if (fslot >= NS) { fslot -= NS; fpar ^= 1; }
          }
          const bf16* src = W + ((size_t)(cta + n_cta * i) * g.rows_per_task + r) * g.K + (size_t)seg * DP_UNIT_CHUNKS * 8;
          const int nvalid = min(DP_UNIT_CHUNKS, nch - seg * DP_UNIT_CHUNKS);
          const uint32_t dst = ring_u32 + (uint32_t)slot * DP_UNIT_BYTES;
#pragma unroll
          for (int u = 0; u < DP_UPL; ++u) {
            const int ch = lane + 32 * u;
            if (ch < nvalid) cp_async16(dst + ch * 16, src + ch * 8, pol);
          }
          cp_async_arrive(&full_bar[slot]);
          slot += DP_PROD_WARPS; if (slot >= NS) { slot -= NS; par ^= 1; }
          n_mine += DP_PROD_WARPS;
          rem += DP_PROD_WARPS;
          while (rem >= upt) { rem -= upt; ++i; }
        }
        n += g.units;
      }
    if (!waited) pdl_wait();
    cp_async_wait_all();
    return;
  }

  // =====================================================================================================================
  // consumer warps
  // =====================================================================================================================
  pdl_wait();
  // dependents (the head of the next step) may only become resident once the sampler of THIS step has completed: they
  // read the slot state before their own griddepcontrol.wait
  pdl_launch_dependents();
  trace_begin(a.trace);
  // ---- state written by this step's sampler ----
  const SlotDev& sl = a.slots[0];
  if (!sl.active) return;                              // uniform over the grid: no barrier has been touched
  // step-invariant scalars are parked in shared memory rather than in (spilled) registers
  __shared__ int s_Lself, s_Lcross, s_ns_cross;
  const int last_token = sl.last_token;
  if (tid == 0) { s_Lself = sl.cur_len; s_Lcross = sl.n_text; }
  for (int i = tid; i < D / 2; i += DP_CONS) { cs[i] = a.rope_cs[i]; sn[i] = a.rope_cs[D / 2 + i]; }
  // thread-0-only bookkeeping lives in shared memory: every register counts (a spill is an L2 round trip here)
  __shared__ unsigned long long bar_target;
  __shared__ int probe_i, fine_i, cur_layer;
  if (tid == 0) { bar_target = 0; probe_i = 0; fine_i = 32; cur_layer = 0; }
  const bool probing = a.probe && cta == 0 && tid == 0;
#define DP_PROBE(l_) do { if (probing && (l_) == a.probe_layer && probe_i < 32) a.probe[probe_i++] = globaltimer_ns(); } while (0)
  // finer checkpoints of the same layer: entries [32, 96)
#define DP_FINE() do { if (probing && cur_layer == a.probe_layer && fine_i < 96) a.probe[fine_i++] = globaltimer_ns(); } while (0)

  // ring slot / parity of the first unit of the current phase: uniform values kept in shared memory, double-buffered by
  // phase parity (phases alternate 0,1,0,1,...; a warp cannot enter phase k+1 before every warp has left the CTA barrier
  // that follows their read in phase k)
  __shared__ int s_base[2][2];
  if (tid == 0) { s_base[0][0] = 0; s_base[0][1] = 0; }
  int rb_slot = 0, rb_par = 0;                         // (dbg & 4: per-thread copies instead, for bisecting)
  // all tasks of phase p that belong to this warp; the x vector of the phase is in xbuf[0, K)
  auto run_phase = [&](int p, uint2* out, unsigned tag, float* plain) {
    const PhaseGeo g = geo[p];
    const int KS = g.kseg, rpt = g.rows_per_task, ng = g.ng;
    int seg = warp, gi = 0;
    while (seg >= KS) { seg -= KS; ++gi; }             // warp = gi * KS + seg (KS is 1, 2 or 4 in practice)
    const int nt = (gi < ng) ? g.nt_cta : 0;           // DP_CONS_WARPS % KS warps sit the phase out
    const int cbase = seg * DP_UNIT_CHUNKS, nvalid = min(DP_UNIT_CHUNKS, (g.K >> 3) - cbase);
    // this warp's slice of x stays in registers for the whole phase
    float xr[DP_UPL][8];
#pragma unroll
    for (int u = 0; u < DP_UPL; ++u) {
      const int ch = lane + 32 * u;
      if (ch < nvalid) {
        const float4 x0 = *reinterpret_cast<const float4*>(xbuf + (size_t)(cbase + ch) * 8);
        const float4 x1 = *reinterpret_cast<const float4*>(xbuf + (size_t)(cbase + ch) * 8 + 4);
        xr[u][0] = x0.x; xr[u][1] = x0.y; xr[u][2] = x0.z; xr[u][3] = x0.w;
        xr[u][4] = x1.x; xr[u][5] = x1.y; xr[u][6] = x1.z; xr[u][7] = x1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) xr[u][j] = 0.f;
      }
    }
    // first unit of this warp: n = base + gi * rpt * KS + seg ; next row of the task: + KS ; next task: + ng * rpt * KS
    int base_slot = (a.dbg & 4) ? rb_slot : s_base[p & 1][0], base_par = (a.dbg & 4) ? rb_par : s_base[p & 1][1];
    int slot = base_slot + gi * rpt * KS + seg, par = base_par;
    while (slot >= NS) { slot -= NS; par ^= 1; }
    const int task_stride = ng * rpt * KS;
    const bool full_unit = (nvalid == DP_UNIT_CHUNKS);
    DP_FINE();
#pragma unroll 1
    for (int i = gi; i < nt; i += ng) {
      float acc0 = 0.f, acc1 = 0.f;       // (no indexed local array: with ~200 KB of shared memory carved out, L1 is too
                                          //  small for local memory and every spill / stack access costs an L2 round trip)
      int s2 = slot, p2 = par;
#pragma unroll 1
      for (int r = 0; r < rpt; ++r) {
        if (!dbg_nostream) mbar_wait_checked(&full_bar[s2], p2, a.err, 9);
        const uint4* w = reinterpret_cast<const uint4*>(ring + (size_t)s2 * DP_UNIT_BYTES) + lane;
        float ae = 0.f, ao = 0.f, be = 0.f, bo = 0.f;   // four independent FMA chains
        if (!dbg_nomath && !dbg_nostream) {
          if (full_unit) {
#pragma unroll
            for (int u = 0; u < DP_UPL; u += 3) {
              const uint4 w0 = w[32 * u], w1 = w[32 * (u + 1)], w2 = w[32 * (u + 2)];
              float f0[8], f1[8], f2[8];
              bf16x8_to_f32(w0, f0); bf16x8_to_f32(w1, f1); bf16x8_to_f32(w2, f2);
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                ae = fmaf(f0[j], xr[u][j], ae); ao = fmaf(f0[j + 1], xr[u][j + 1], ao);
                be = fmaf(f1[j], xr[u + 1][j], be); bo = fmaf(f1[j + 1], xr[u + 1][j + 1], bo);
                ae = fmaf(f2[j], xr[u + 2][j], ae); ao = fmaf(f2[j + 1], xr[u + 2][j + 1], ao);
              }
            }
          } else {
#pragma unroll
            for (int u = 0; u < DP_UPL; ++u) {
              if (lane + 32 * u < nvalid) {
                float f0[8];
                bf16x8_to_f32(w[32 * u], f0);
#pragma unroll
                for (int j = 0; j < 8; j += 2) { ae = fmaf(f0[j], xr[u][j], ae); ao = fmaf(f0[j + 1], xr[u][j + 1], ao); }
              }
            }
          }
        }
        const float acc = (ae + ao) + (be + bo);
        if (r == 0) acc0 = acc; else acc1 = acc;
        __syncwarp();
        if (lane == 0 && !dbg_nostream) mbar_arrive(&empty_bar[s2]);
        s2 += KS; if (s2 >= NS) { s2 -= NS; p2 ^= 1; }
      }
      const float v0 = warp_sum(acc0);
      float v1 = 0.f;
      if (rpt == 2) v1 = warp_sum(acc1);
      if (lane == 0) {
        const float o = (rpt == 2) ? gelu_tanh_f(v0) * v1 : v0;
        if (KS == 1) { st_tag(out + cta + n_cta * i, o, tag); if (plain) plain[cta + n_cta * i] = o; }
        else part[i * KS + seg] = o;
      }
      slot += task_stride;
      while (slot >= NS) { slot -= NS; par ^= 1; }
    }
    DP_FINE();
    base_slot += g.umod; if (base_slot >= NS) { base_slot -= NS; base_par ^= 1; }
    base_par ^= g.upar;
    rb_slot = base_slot; rb_par = base_par;
    if (tid == 0) { s_base[(p & 1) ^ 1][0] = base_slot; s_base[(p & 1) ^ 1][1] = base_par; }
    if (KS > 1) {                                      // split-K inside the CTA: combine the K-segments of every row
      cons_sync();
      for (int i = tid; i < g.nt_cta; i += DP_CONS) {
        float v = 0.f;
        for (int k2 = 0; k2 < KS; ++k2) v += part[i * KS + k2];
        st_tag(out + cta + n_cta * i, v, tag);
        if (plain) plain[cta + n_cta * i] = v;
      }
    }
  };

  // residual stream, replicated per CTA: thread t holds h[t + 512 i]
  float hreg[DP_NP];
  // h += rmsnorm(y) * g_post (when y) ; xbuf = rmsnorm(h) * g_pre
  auto sandwich = [&](const uint2* y, unsigned tag, const float* g_post, const float* g_pre) {
    float gp[DP_NP], gq[DP_NP], yv[DP_NP];
#pragma unroll
    for (int i = 0; i < DP_NP; ++i) {
      const int k = tid + i * DP_CONS;
      gp[i] = (k < d) ? g_pre[k] : 0.f;
      gq[i] = (y && k < d) ? g_post[k] : 0.f;
      yv[i] = 0.f;
    }
    if (y) ld_tag_vec<DP_NP>(y, tid, DP_CONS, d, tag, yv, a.err, 1);
    float s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
    if (yv[0] == 12345.678f) s1 = 1.f;                 // (keeps the probe below after the loads have returned)
    DP_FINE();
#pragma unroll
    for (int i = 0; i < DP_NP; ++i) {
      const float yg = yv[i] * gq[i];
      s1 = fmaf(yv[i], yv[i], s1); s2 = fmaf(hreg[i], hreg[i], s2);
      s3 = fmaf(hreg[i], yg, s3); s4 = fmaf(yg, yg, s4);
      yv[i] = yg;
    }
    cons_sum4(s1, s2, s3, s4, red);
    DP_FINE();
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
      const int k = tid + i * DP_CONS;
      if (k < d) xbuf[k] = hreg[i] * rinv * gp[i];
    }
    cons_sync();
  };

  // split-KV attention of (kv head hk, split) on this CTA: partial -> exchange among the ns CTAs of the head group -> every
  // CTA finalises a D/ns slice of the group's heads and publishes it in attn_out
  auto attention = [&](int l, bool is_cross, const uint2* q, const uint2* kv_new, unsigned tag_in, int ns,
                       uint2* part_o, uint2* part_ml, unsigned tag_part, uint2* attn_out, unsigned tag_out) {
    const int hk = cta % a.Hkv, split = cta / a.Hkv;
    const int* bt = is_cross ? bt_cross : bt_self;
    const int* bt_g = is_cross ? a.cross_bt : a.self_bt;
    const int PT = a.pool.page_tokens;
    const int Lk = is_cross ? s_Lcross : s_Lself;
    const int window = (!is_cross && lay[l].sliding) ? a.window : 0;
    const int lo = window > 0 ? max(0, Lk - window) : 0;
    int chunk = (Lk - lo + ns - 1) / ns;
    chunk = (chunk + TPW - 1) / TPW * TPW;
    const int t_begin = lo + split * chunk, t_end = min(Lk, t_begin + chunk);
    const bool has_new = (!is_cross) && (t_end == Lk) && (t_end > t_begin);
    const int grp = lane / LPT, l8 = lane % LPT;
    auto page_of = [&](int t) -> int { const int pi = t / PT; return pi < DP_BT ? bt[pi] : bt_g[pi]; };
    // K/V rows of earlier tokens are immutable: this warp's first tile is requested before anything that depends on
    // the producer phase, and every later tile one iteration ahead of its use
    uint4 ku[NV], vu[NV];
    auto load_tile = [&](int t0, uint4* kd, uint4* vd) {
      const int t = t0 + grp;
      if (t < t_end && !(has_new && t == Lk - 1)) {
        const int page = page_of(t), off = t % PT;
        const bf16* kp = a.pool.ptr(l, 0, page) + ((size_t)hk * PT + off) * D + l8 * DPL;
        const bf16* vp = a.pool.ptr(l, 1, page) + ((size_t)hk * PT + off) * D + l8 * DPL;
#pragma unroll
        for (int i = 0; i < NV; ++i) { kd[i] = *reinterpret_cast<const uint4*>(kp + i * 8); vd[i] = *reinterpret_cast<const uint4*>(vp + i * 8); }
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) { kd[i] = make_uint4(0, 0, 0, 0); vd[i] = make_uint4(0, 0, 0, 0); }
      }
    };
    load_tile(t_begin + warp * TPW, ku, vu);
    // producer outputs: raw q of the group's heads and the new k/v
    {
      constexpr int QB = (G * D + DP_CONS - 1) / DP_CONS, KB = (D + DP_CONS - 1) / DP_CONS;
      float qv[QB];
      ld_tag_vec<QB>(q + (size_t)(hk * G) * D, tid, DP_CONS, G * D, tag_in, qv, a.err, 2);
#pragma unroll
      for (int i = 0; i < QB; ++i) { const int k = tid + i * DP_CONS; if (k < G * D) qs[k / D][k % D] = qv[i]; }
      if (has_new) {
        float kv[KB], vv[KB];
        ld_tag_vec<KB>(kv_new + (size_t)hk * D, tid, DP_CONS, D, tag_in, kv, a.err, 3);
        ld_tag_vec<KB>(kv_new + (size_t)(a.Hkv + hk) * D, tid, DP_CONS, D, tag_in, vv, a.err, 4);
#pragma unroll
        for (int i = 0; i < KB; ++i) {
          const int k = tid + i * DP_CONS;
          if (k < D) { knew[k] = kv[i]; vnew[k] = __bfloat162float(__float2bfloat16(vv[i])); }
        }
      }
    }
    cons_sync();
    for (int i = tid; i < G * D / 2; i += DP_CONS) {             // PM-RoPE, half-split pairs (j, j + D/2)
      const int g = i / (D / 2), j = i - g * (D / 2);
      const float x1 = qs[g][j], x2 = qs[g][j + D / 2];
      qs[g][j] = x1 * cs[j] - x2 * sn[j];
      qs[g][j + D / 2] = x2 * cs[j] + x1 * sn[j];
    }
    if (has_new)
      for (int j = tid; j < D / 2; j += DP_CONS) {
        const float x1 = knew[j], x2 = knew[j + D / 2];
        knew[j] = __bfloat162float(__float2bfloat16(x1 * cs[j] - x2 * sn[j]));
        knew[j + D / 2] = __bfloat162float(__float2bfloat16(x2 * cs[j] + x1 * sn[j]));
      }
    cons_sync();
    if (has_new) {                                                  // append (K post-RoPE), visible to later steps
      const int t = Lk - 1, page = page_of(t), off = t % PT;
      bf16* kd = a.pool.ptr(l, 0, page) + ((size_t)hk * PT + off) * D;
      bf16* vd = a.pool.ptr(l, 1, page) + ((size_t)hk * PT + off) * D;
      for (int j = tid; j < D; j += DP_CONS) { kd[j] = __float2bfloat16(knew[j]); vd[j] = __float2bfloat16(vnew[j]); }
    }
    GroupState<G, DPL> st;
    st.init();
    const float inv_softcap = a.softcap > 0.f ? 1.f / a.softcap : 0.f;
#pragma unroll 1
    for (int t0 = t_begin + warp * TPW; t0 < t_end; t0 += DP_CONS_WARPS * TPW) {
      const int t = t0 + grp;
      const bool valid = t < t_end, fresh = valid && has_new && t == Lk - 1;
      uint4 kn[NV], vn[NV];
      load_tile(t0 + DP_CONS_WARPS * TPW, kn, vn);       // next tile in flight while this one is consumed
      // scores of the group's G query heads (q stays in shared memory: registers are the scarce resource here)
      float dot[G];
#pragma unroll
      for (int g = 0; g < G; ++g) dot[g] = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float kf[8];
        if (fresh) {
#pragma unroll
          for (int j = 0; j < 8; ++j) kf[j] = knew[l8 * DPL + i * 8 + j];
        } else {
          bf16x8_to_f32(ku[i], kf);
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float4 q0 = *reinterpret_cast<const float4*>(&qs[g][l8 * DPL + i * 8]);
          const float4 q1 = *reinterpret_cast<const float4*>(&qs[g][l8 * DPL + i * 8 + 4]);
          dot[g] = fmaf(q0.x, kf[0], dot[g]); dot[g] = fmaf(q0.y, kf[1], dot[g]); dot[g] = fmaf(q0.z, kf[2], dot[g]); dot[g] = fmaf(q0.w, kf[3], dot[g]);
          dot[g] = fmaf(q1.x, kf[4], dot[g]); dot[g] = fmaf(q1.y, kf[5], dot[g]); dot[g] = fmaf(q1.z, kf[6], dot[g]); dot[g] = fmaf(q1.w, kf[7], dot[g]);
        }
      }
      float pw[G], corr[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float dd = dot[g];
#pragma unroll
        for (int o = LPT >> 1; o > 0; o >>= 1) dd += __shfl_xor_sync(0xffffffffu, dd, o);
        float sc = dd * a.scale;
        if (a.softcap > 0.f) {                   // softcap*tanh(s/softcap), tanh(x) = 1 - 2/(exp(2x)+1)
          const float e2 = __expf(2.f * sc * inv_softcap);
          sc = a.softcap * (1.f - __fdividef(2.f, e2 + 1.f));
        }
        if (!valid) sc = -INFINITY;
        const float mn = fmaxf(st.m[g], sc);
        corr[g] = (st.m[g] == -INFINITY) ? 0.f : __expf(st.m[g] - mn);
        pw[g] = valid ? __expf(sc - mn) : 0.f;
        st.l[g] = st.l[g] * corr[g] + pw[g];
        st.m[g] = mn;
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float vf[8];
        if (fresh) {
#pragma unroll
          for (int j = 0; j < 8; ++j) vf[j] = vnew[l8 * DPL + i * 8 + j];
        } else {
          bf16x8_to_f32(vu[i], vf);
        }
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
          for (int j = 0; j < 8; ++j) st.acc[g][i * 8 + j] = fmaf(pw[g], vf[j], st.acc[g][i * 8 + j] * corr[g]);
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) { ku[i] = kn[i]; vu[i] = vn[i]; }
    }
    warp_merge<G, D>(st);
    float* w_o = xbuf;                                              // [DP_CONS_WARPS][G][D] (the x buffer is dead here)
    if (grp == 0) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (l8 == 0) { w_ml[warp][g][0] = st.m[g]; w_ml[warp][g][1] = st.l[g]; }
#pragma unroll
        for (int i = 0; i < DPL; ++i) w_o[((size_t)warp * G + g) * D + l8 * DPL + i] = st.acc[g][i];
      }
    }
    cons_sync();
    if (tid < G) {
      float M = -INFINITY;
#pragma unroll
      for (int w = 0; w < DP_CONS_WARPS; ++w) M = fmaxf(M, w_ml[w][tid][0]);
      float den = 0.f;
#pragma unroll
      for (int w = 0; w < DP_CONS_WARPS; ++w) {
        const float m = w_ml[w][tid][0];
        const float wt = (m == -INFINITY) ? 0.f : __expf(m - M);
        den = fmaf(wt, w_ml[w][tid][1], den);
        w_wt[w][tid] = wt;
      }
      c_ml[tid][0] = M; c_ml[tid][1] = den;
    }
    cons_sync();
    for (int i = tid; i < G * D; i += DP_CONS) {
      const int g = i / D, dd = i - g * D;
      float num = 0.f;
#pragma unroll
      for (int w = 0; w < DP_CONS_WARPS; ++w) num = fmaf(w_wt[w][g], w_o[((size_t)w * G + g) * D + dd], num);
      st_tag(part_o + ((size_t)(hk * G + g) * DP_MAX_NS + split) * D + dd, num, tag_part);
    }
    if (tid < G) {
      st_tag(part_ml + ((hk * G + tid) * DP_MAX_NS + split) * 2, c_ml[tid][0], tag_part);
      st_tag(part_ml + ((hk * G + tid) * DP_MAX_NS + split) * 2 + 1, c_ml[tid][1], tag_part);
    }
    cons_sync();      // w_o aliases the x buffer, which the next stage's prologue overwrites: every warp is done reading it
    // second hop, among the ns CTAs of this head group only: dims [split * D/ns, (split+1) * D/ns) of the G heads
    const int dsl = D / ns;
    for (int ob = warp * 32; ob < G * dsl; ob += DP_CONS) {        // warp-uniform trip count (ld_tag_vec syncs the warp)
      const int o = ob + lane;
      const bool live = o < G * dsl;
      const int g = live ? o / dsl : 0, dd = split * dsl + (live ? o - g * dsl : 0), h = hk * G + g;
      float v[DP_MAX_NS], ml[2 * DP_MAX_NS];
      ld_tag_vec<DP_MAX_NS>(part_o + (size_t)h * DP_MAX_NS * D + dd, 0, D, live ? ns * D : 0, tag_part, v, a.err, 5);
      ld_tag_vec<2 * DP_MAX_NS>(part_ml + h * DP_MAX_NS * 2, 0, 1, live ? 2 * ns : 0, tag_part, ml, a.err, 6);
      float M = -INFINITY;
#pragma unroll
      for (int s2 = 0; s2 < DP_MAX_NS; ++s2) if (s2 < ns) M = fmaxf(M, ml[2 * s2]);
      float den = 0.f, num = 0.f;
#pragma unroll
      for (int s2 = 0; s2 < DP_MAX_NS; ++s2) {
        if (s2 < ns) {
          const float wt = (ml[2 * s2] == -INFINITY) ? 0.f : __expf(ml[2 * s2] - M);
          den = fmaf(wt, ml[2 * s2 + 1], den);
          num = fmaf(wt, v[s2], num);
        }
      }
      if (live) st_tag(attn_out + (size_t)h * D + dd, den > 0.f ? num / den : 0.f, tag_out);
    }
  };

  // xbuf[0, n) = a tagged vector published by the previous stage (attention output, GeGLU activations)
  auto load_x = [&](const uint2* src, int n, unsigned tag) {
    constexpr int XB = 24;
#pragma unroll 1
    for (int base = 0; base < n; base += XB * DP_CONS) {
      float v[XB];
      ld_tag_vec<XB>(src + base, tid, DP_CONS, n - base, tag, v, a.err, 7);
#pragma unroll
      for (int i = 0; i < XB; ++i) { const int k = tid + i * DP_CONS; if (k < n - base) xbuf[base + k] = v[i]; }
    }
    cons_sync();
  };

  auto n_splits = [&](int keys, int ns_max) {          // power of two (the second hop deals D/ns dims to every split CTA)
    const int want = max(1, min(ns_max, (keys + a.keys_per_split - 1) / a.keys_per_split));
    int ns = 1;
    while (ns * 2 <= want && D % (ns * 2) == 0) ns *= 2;
    return ns;
  };
  cons_sync();
  if (tid == 0) s_ns_cross = n_splits(s_Lcross, a.ns_max);

  // ---- layer 0 input: audio embedding * sqrt(d) (models/t5gemma.py:1083; HF:769) ----
#pragma unroll
  for (int i = 0; i < DP_NP; ++i) {
    const int k = tid + i * DP_CONS;
    hreg[i] = (k < d) ? __bfloat162float(a.emb[(size_t)last_token * d + k]) * a.emb_scale : 0.f;
  }

  // exchange buffers (tagged pairs): offsets live in shared memory (a dozen live 64-bit pointers would spill)
  enum { X_QKV = 0, X_PS_O, X_PS_ML, X_AS, X_YO, X_QC, X_PC_O, X_PC_ML, X_AC, X_YOC, X_ACT, X_YDN, X_N };
  __shared__ int xoff[X_N];
  if (tid == 0) {
    const int sizes[X_N] = {QKV, a.Hq * DP_MAX_NS * D, a.Hq * DP_MAX_NS * 2, QD, d, QD, a.Hq * DP_MAX_NS * D, a.Hq * DP_MAX_NS * 2, QD, d, I, d};
    int o = 0;
    for (int i = 0; i < X_N; ++i) { xoff[i] = o; o += sizes[i]; }
  }
  cons_sync();
#define XB(k) (a.xchg + xoff[k])
  const unsigned ep_tag = (*reinterpret_cast<const volatile unsigned*>(a.epoch)) << 10;
  auto TAG = [&](int l, int sid) -> unsigned { return ep_tag + (unsigned)(l * 16 + sid + 1); };
  enum { S_QKV = 0, S_PS, S_AS, S_YO, S_QC, S_PC, S_AC, S_YOC, S_ACT, S_YDN };

  // One layer = 8 stages.  The stage loop has ONE call site per building block (prologue kinds, attention, projection):
  // the first version inlined them per stage and grew to 270 KB of SASS, which made the whole kernel instruction-fetch
  // bound (69 us per layer with the weight stream switched off).
#pragma unroll 1
  for (int l = 0; l < a.n_layers; ++l) {
    if (tid == 0) cur_layer = l;
    const int win = lay[l].sliding ? a.window : 0;
    const int ns_self = n_splits(win > 0 ? min(s_Lself, win) : s_Lself, a.ns_max);
#pragma unroll 1
    for (int stage = 0; stage < 8; ++stage) {
      DP_PROBE(l);
      const PersistLayer& Ly = lay[l];
      const bool is_attn = (stage == 1 || stage == 4);
      if (stage == 0 || stage == 3 || stage == 6) {
        // qkv / cross q / gate|up consume pre_norm(h + post_norm(previous sub-layer output))
        const uint2* y = stage == 0 ? (l == 0 ? nullptr : XB(X_YDN)) : stage == 3 ? XB(X_YO) : XB(X_YOC);
        const unsigned tg = stage == 0 ? TAG(l - 1, S_YDN) : stage == 3 ? TAG(l, S_YO) : TAG(l, S_YOC);
        const float* g_post = stage == 0 ? (l > 0 ? lay[l - 1].g_post_ff : nullptr) : stage == 3 ? Ly.g_post_sa : Ly.g_post_ca;
        const float* g_pre = stage == 0 ? Ly.g_pre_sa : stage == 3 ? Ly.g_pre_ca : Ly.g_pre_ff;
        sandwich(y, tg, g_post, g_pre);
      } else if (stage == 2) {
        load_x(XB(X_AS), QD, TAG(l, S_AS));                              // o projections consume the attention output
      } else if (stage == 5) {
        load_x(XB(X_AC), QD, TAG(l, S_AC));
      } else if (stage == 7) {
        load_x(XB(X_ACT), I, TAG(l, S_ACT));                             // down projection consumes the GeGLU activations
      }
      if (is_attn) {
        const bool cross = stage == 4;
        const int ns = cross ? s_ns_cross : ns_self;
        if (cta < a.Hkv * ns) {
#pragma unroll
          for (int i = 0; i < DP_NP; ++i) h_park[i * DP_CONS + tid] = hreg[i];
          if (cross) attention(l, true, XB(X_QC), nullptr, TAG(l, S_QC), ns, XB(X_PC_O), XB(X_PC_ML), TAG(l, S_PC), XB(X_AC), TAG(l, S_AC));
          else attention(l, false, XB(X_QKV), XB(X_QKV) + QD, TAG(l, S_QKV), ns, XB(X_PS_O), XB(X_PS_ML), TAG(l, S_PS), XB(X_AS), TAG(l, S_AS));
#pragma unroll
          for (int i = 0; i < DP_NP; ++i) hreg[i] = h_park[i * DP_CONS + tid];
        }
      } else {
        const int p = stage == 0 ? 0 : stage == 2 ? 1 : stage == 3 ? 2 : stage == 5 ? 3 : stage == 6 ? 4 : 5;
        uint2* out = XB(stage == 0 ? X_QKV : stage == 2 ? X_YO : stage == 3 ? X_QC : stage == 5 ? X_YOC : stage == 6 ? X_ACT : X_YDN);
        const int sid = stage == 0 ? S_QKV : stage == 2 ? S_YO : stage == 3 ? S_QC : stage == 5 ? S_YOC : stage == 6 ? S_ACT : S_YDN;
        // the head kernel of the next step reads the last layer's MLP output as plain floats
        run_phase(p, out, TAG(l, sid), (stage == 7 && l + 1 == a.n_layers) ? a.y : nullptr);
      }
      DP_PROBE(l);
    }
  }
  if (cta == 0 && tid == 0) atomicAdd(a.epoch, 1u);     // every CTA read the epoch before it could hand CTA 0 anything
  // the head kernel applies post_ff of the last layer: hand it h (CTA 0) and y (already in a.y)
  if (cta == 0) {
#pragma unroll
    for (int i = 0; i < DP_NP; ++i) {
      const int k = tid + i * DP_CONS;
      if (k < d) a.h_out[k] = hreg[i];
    }
  }
  trace_end(a.trace);
#undef DP_PROBE
#undef XB
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

}  // namespace

int decode_persist_xbuf_floats(int d, int I, int QD, int G, int D) {
  int m = d > I ? d : I;
  if (QD > m) m = QD;
  const int attn = DP_CONS_WARPS * G * D;     // per-warp attention partials alias the x buffer
  if (attn > m) m = attn;
  return (m + 7) & ~7;
}

size_t decode_persist_xchg_entries(int d, int I, int Hq, int Hkv, int D) {
  const size_t QD = (size_t)Hq * D, QKV = QD + 2 * (size_t)Hkv * D, part = (size_t)Hq * DP_MAX_NS * (D + 2);
  return QKV + 2 * part + 3 * QD + 3 * (size_t)d + I + 64;
}

// ring slots that fit next to the x buffer and ~30 KB of static shared memory
int decode_persist_slots(int xbuf_floats) {
  const long budget = 227L * 1024 - 32L * 1024 - (long)xbuf_floats * 4;
  long n = budget / DP_UNIT_BYTES;
  if (n > DP_MAX_SLOTS) n = DP_MAX_SLOTS;
  return (int)(n / DP_PROD_WARPS * DP_PROD_WARPS);
}

bool decode_persist_supported(int d, int I, int Hq, int Hkv, int D, int n_layers, int num_sms) {
  if (Hkv <= 0 || Hq % Hkv) return false;
  const int G = Hq / Hkv;
  if (G != 2 || !(D == 16 || D == 32 || D == 64 || D == 128 || D == 256)) return false;
  if (d % 8 || I % 8 || d > DP_CONS * DP_NP) return false;
  if (n_layers > T5G_PERSIST_MAX_LAYERS || Hq > T5G_PERSIST_MAX_HEADS || Hkv > num_sms) return false;
  auto ks = [](int K) { return ((K >> 3) + DP_UNIT_CHUNKS - 1) / DP_UNIT_CHUNKS; };
  const int Ks[3] = {d, Hq * D, I}, Ns[3] = {I > Hq * D + 2 * Hkv * D ? I : Hq * D + 2 * Hkv * D, d, d};
  for (int i = 0; i < 3; ++i) {
    const int k = ks(Ks[i]);
    if (k > DP_CONS_WARPS) return false;                                         // K-segments are dealt to warp groups
    if (k > 1 && ((Ns[i] + num_sms - 1) / num_sms) * k > DP_PART) return false;
  }
  return decode_persist_slots(decode_persist_xbuf_floats(d, I, Hq * D, G, D)) >= 8;
}

cudaError_t launch_decode_persist(const PersistArgs& a, int num_sms, cudaStream_t st, bool pdl) {
  const int G = a.Hq / a.Hkv;
  if (a.n_slots < 2 * DP_PROD_WARPS || a.n_slots > DP_MAX_SLOTS || a.n_slots % DP_PROD_WARPS || a.max_inflight < 1)
    return cudaErrorInvalidValue;
  const size_t smem = (size_t)a.n_slots * DP_UNIT_BYTES + (size_t)a.xbuf_floats * sizeof(float);
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
