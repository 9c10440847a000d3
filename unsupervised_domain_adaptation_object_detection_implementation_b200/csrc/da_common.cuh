// Shared device/host helpers of libda_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/da_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libda_b200 is written for sm_100a (B200) only"
#endif

namespace da {

// ---- error + launch accounting (host) --------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
// ---- per-device state (indexed by the CURRENT device of the calling thread; one process may drive several GPUs) -----
// seed_counter: device-resident dropout step counter (da_set_dropout_counter): kernels add *counter to their seed, so a
//   CUDA-graph replay of a captured step still draws fresh masks when the host bumps the counter on device.
// sm_limit: SM budget of the persistent kernels (da_set_sm_limit).
constexpr int kMaxDevices = 64;
struct DeviceState {
  const unsigned long long* seed_counter;
  int sm_limit;
  int sm_count;   // cached cudaDevAttrMultiProcessorCount (0 = not queried yet)
};
extern DeviceState g_dev[kMaxDevices];
inline int cur_dev() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
inline DeviceState& dev_state() {
  int dev = 0;
  cudaGetDevice(&dev);
  return g_dev[(dev >= 0 && dev < kMaxDevices) ? dev : 0];
}
#define g_seed_counter (da::dev_state().seed_counter)

// ---- debug / test options: read ONCE from the environment when the library is loaded (never in a launch path);
// da_set_option changes one at run time (test hook: e.g. force the CUDA-core RoIAlign path).
struct Options {
  int roi_no_tc;      // DA_ROI_NO_TC      : bf16 RoIAlign on the CUDA-core kernels instead of tcgen05
  int umma_no_bn64;   // DA_UMMA_NO_BN64   : no 64-wide GEMM tiles
  int umma_no_2sm;    // DA_UMMA_NO_2SM    : no cta_group::2 pair MMA
  int no_pdl;         // DA_NO_PDL         : no programmatic dependent launch
  int umma_dbg;       // DA_UMMA_DBG       : timing experiments (results are wrong when set)
  int umma_no_bn512;  // DA_UMMA_NO_BN512  : no 512-wide pair tiles for long-K forward GEMMs
  int chain_no_bn128; // DA_CHAIN_NO_BN128  : instance-head chain kernel with 64-wide tiles only (A/B)
  int roi_bwd_dbg;    // DA_ROI_BWD_DBG    : timing experiments (results are wrong when set)
  int roi_fwd_dbg;    // DA_ROI_FWD_DBG    : timing experiments (bit 0: no MMAs, bit 1: no output stores; results are wrong when set)
  unsigned long long chain_trace;     // device address of a u64 buffer: per-group globaltimer stamps of the chain kernel (tools/trace_chain.py)
  unsigned long long roi_bwd_trace;   // DA_ROI_BWD_TRACE: device address of a [ctas][8] u64 trace buffer (tools/trace_roi_bwd.py)
};
extern Options g_opt;
__device__ __forceinline__ unsigned long long effective_seed(unsigned long long seed, const unsigned long long* ctr) {
  return ctr ? seed + *ctr : seed;
}

#define DA_REQUIRE(cond, code, ...)            \
  do {                                         \
    if (!(cond)) {                             \
      da::set_error(__VA_ARGS__);              \
      return (code);                           \
    }                                          \
  } while (0)

#define DA_CUDA_OK(expr)                                                            \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      da::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),         \
                    __FILE__, __LINE__);                                            \
      return DA_ERR_CUDA;                                                           \
    }                                                                               \
  } while (0)

#define DA_LAUNCH_CHECK()                                                           \
  do {                                                                              \
    da::count_launch();                                                             \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      da::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),     \
                    __FILE__, __LINE__);                                            \
      return DA_ERR_CUDA;                                                           \
    }                                                                               \
  } while (0)

// SM budget of the persistent kernels (da_set_sm_limit): while a NCCL all-reduce holds some SMs, a persistent grid of
// one CTA per PHYSICAL SM would run in two waves; the caller lowers the budget for the kernels it overlaps.
inline int num_sms_physical() {
  DeviceState& d = dev_state();
  if (d.sm_count == 0) {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    d.sm_count = n > 0 ? n : 148;
  }
  return d.sm_count;
}
inline int num_sms() {
  const int n = num_sms_physical();
  const int lim = dev_state().sm_limit;
  return (lim > 0 && lim < n) ? lim : n;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- device helpers --------------------------------------------------------------------
// Programmatic dependent launch: a kernel launched with the programmatic-serialization attribute may start (and run
// its prologue: barrier init, TMEM allocation, descriptor prefetch) while its predecessor drains; it must not touch
// global memory before pdl_wait().  pdl_launch_dependents() lets the successor start early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
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

// Block-wide sum; result valid in thread 0 (and broadcast to all when kBroadcast).
// `red` must hold >= 33 floats of shared memory.
template <bool kBroadcast>
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` against a previous use
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = (lane < nw) ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  if (kBroadcast) {
    __syncthreads();
    return red[32];
  }
  return (threadIdx.x == 0) ? red[32] : 0.f;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// Stateless dropout keep decision: splitmix64-style hash of (seed, element index).
// keep probability = 1 - p.  Shared by the conv epilogues, act-backward and da_dropout_mask
// so forward, backward and the exported mask agree bit for bit.
__host__ __device__ __forceinline__ uint32_t drop_hash(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  // keep iff hash >= threshold; threshold = p * 2^32
  double t = (double)p * 4294967296.0;
  if (t <= 0.0) return 0u;
  if (t >= 4294967295.0) return 0xFFFFFFFFu;
  return (uint32_t)t;
}

}  // namespace da
