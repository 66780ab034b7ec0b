// Shared device/host helpers for libt5gtts (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libt5gtts is written for sm_100a (B200) only"
#endif

typedef __nv_bfloat16 bf16;

#define T5G_WARP 32

// ---- host error plumbing -------------------------------------------------------------------
void t5g_set_error(const char* fmt, ...);
#define T5G_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      t5g_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));   \
      return T5G_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)
#define T5G_CHECK(cond, code, ...)                                                           \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      t5g_set_error(__VA_ARGS__);                                                            \
      return (code);                                                                         \
    }                                                                                        \
  } while (0)

// ---- device helpers ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum; `red` is >= 32 floats of shared memory; all threads must call.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();                 // protect `red` against a previous use
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

// block-wide sum of four values at once (one barrier pair)
__device__ __forceinline__ void block_sum4(float& a, float& b, float& c, float& d, float* red /* >= 128 floats */) {
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c); d = warp_sum(d);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) { red[w] = a; red[32 + w] = b; red[64 + w] = c; red[96 + w] = d; }
  __syncthreads();
  a = warp_sum(lane < nw ? red[lane] : 0.f); b = warp_sum(lane < nw ? red[32 + lane] : 0.f);
  c = warp_sum(lane < nw ? red[64 + lane] : 0.f); d = warp_sum(lane < nw ? red[96 + lane] : 0.f);
}

// 8 bf16 packed in a uint4 -> 8 floats (bf16 -> fp32 is a 16-bit shift)
__device__ __forceinline__ void bf16x8_to_f32(const uint4& u, float* f) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}

// streaming 128-bit weight load: read-only path, no L1 allocation, L2 evict-first (the 4.85 GB/step weight
// stream must not evict the KV pages and activations that the rest of the step re-reads from L2)
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint4 ldg_stream(const void* p, uint64_t pol) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}

#define T5G_TRACE_STRIDE 1024
// optional in-step tracing (T5G_TRACE=1): per kernel, min over CTAs of the time after griddepcontrol.wait and
// max over CTAs of the exit time, in %globaltimer nanoseconds
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_begin(unsigned long long* tr) {
  if (tr && threadIdx.x == 0) atomicMin(tr, globaltimer_ns());
}
__device__ __forceinline__ void trace_end(unsigned long long* tr) {
  if (tr && threadIdx.x == 0) atomicMax(tr + T5G_TRACE_STRIDE, globaltimer_ns());
}

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // gelu_pytorch_tanh: 0.5x(1+tanh(sqrt(2/pi)(x+0.044715x^3)))
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.f + tanhf(inner));
}
__device__ __forceinline__ float gelu_erf_f(float x) {
  return 0.5f * x * (1.f + erff(x * 0.7071067811865476f));
}

// Programmatic dependent launch (PDL) controls
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- per-slot request state, resident on the device (one per engine row) ---------------------
struct SlotDev {
  int active;          // 1 while generating
  int finished;        // eos emitted
  int n_generated;     // cur_num_gen
  int cur_len;         // current_length (BOS + prompt + generated)
  int prompt_offset;   // prompt_frames + 1
  int target_total;
  int est_total;
  int n_text;
  int budget_limit;    // floor(target_total - prompt_offset + encodec_sr*extra_cutoff)
  int max_new_tokens;  // 0 = unlimited
  int top_k;
  float top_p, min_p, temperature;
  const float* uniforms; int n_uniforms;
  int topk_sched_off;  // offset into the engine's schedule pool, -1 = none
  int n_topk_sched;
  int last_token;      // token sampled this step (embedding input of the decoder step)
  float pos;           // PM-RoPE position of that token
  int error;           // sampler error flags
  int n_forced;        // teacher-forced prefix length (tests)
  int prev_token;      // silence-repetition state (models/t5gemma.py:967-968)
  int consec_silence;
  int stop_repetition;
  int silence_off;     // offset into the engine's int pool, -1 = none
  int n_silence;
  int pad_;
};

// Function attributes (dynamic shared memory limit, carve-out) are per DEVICE: the "already set" state of a kernel is
// tracked per device ordinal so that engines on several GPUs of one process all get them (setting one twice is harmless).
#define T5G_MAX_DEVICES 64
struct PerDeviceFlag {
  size_t v[T5G_MAX_DEVICES] = {};
  size_t& here() { int d = 0; cudaGetDevice(&d); return v[(d >= 0 && d < T5G_MAX_DEVICES) ? d : 0]; }
};

// Kernels of the batched (tensor-core) decode step all ask for the maximum shared-memory carve-out: a uniform
// configuration lets CTAs of consecutive kernels co-reside on an SM under programmatic dependent launch.  (The single-row
// step keeps its kernels under 100 KB instead: in-flight global loads are tracked in L1, and a large carve-out left behind
// by one kernel slowed the following weight-streaming GEMVs -- gate|up 15.5 -> 19.8 us, measured.)
inline int batched_carveout() { return 100; }
