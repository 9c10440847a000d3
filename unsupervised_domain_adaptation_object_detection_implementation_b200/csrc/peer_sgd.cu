// Gradient mean over the ranks of one NVSwitch box FUSED with the SGD step and with the broadcast of the refreshed
// bf16 operand copy — one kernel over NVLink peer memory instead of  all-reduce -> SGD  (sm_100a).
//
// The reference wires MMDistributedDataParallel (mmdet/apis/train.py:113-121): a bucketed NCCL all-reduce of every
// gradient, then torch.optim.SGD on every rank (train.py:127).  For the one tensor that dominates this path (FC1 of
// the shared bbox head, 103 M parameters = 411 MB of fp32 gradient) that is 2*(N-1)/N*411 MB over the links plus a
// 2.26 GB/rank optimizer pass that every rank repeats.  Here every rank OWNS a contiguous 1/N slice of the tensor:
//
//   barrier-in   ready[r] flags: every rank's weight-gradient kernel has finished (stream order on the sender)
//   reduce       g[i] = (1/N) * sum_r grad_r[i]   for i in the own slice: N-1 streams of 16-byte P2P loads,
//                summed in rank order 0..N-1 (one owner per element => every rank sees the same bits)
//   update       torch.optim.SGD rule on the own slice of the fp32 master + momentum (momentum is stored sharded)
//   publish      bf16(w) of the slice is stored into EVERY rank's operand copy (P2P stores; optionally the fp32 master too)
//   barrier-out  the last CTA of a rank raises done[rank] on every peer; da_peer_wait (a one-warp kernel behind this
//                one) holds the stream until every peer's done flag arrived, i.e. until nobody reads this rank's
//                gradient or writes its operand copy any more.
//
// Link traffic per rank: (N-1)/N * 4 B in + (N-1)/N * 2 B out per parameter (NCCL ring all-reduce: 2*(N-1)/N*4 B each
// way) and the optimizer pass shrinks by N.  Flags live in cudaMalloc'ed blocks exchanged through CUDA IPC
// (da_peer_alloc / da_peer_export / da_peer_open); they carry a monotonically increasing epoch kept on the device, so a
// captured CUDA graph replays correctly.  Every spin has a time-out (error flag, never a hang).
#include "da_common.cuh"
#include <string.h>

namespace da {

constexpr int PEER_THREADS = 512;
constexpr unsigned long long PEER_TIMEOUT_NS = 4000000000ull;   // 4 s

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_cv4(const float* p) {
  float4 v;
  // plain weak load, not ld.cv: measured 14.6 GB/s per SM over NVLink against 7.7 for ld.cv / ld.relaxed.sys (tools/p2p_bw.cu).
  // Safe: a gradient word is read once per kernel, after the acquire of its owner's ready flag, and L1 starts a kernel empty.
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// true when flag >= epoch arrived before the time-out
__device__ __forceinline__ bool spin_until(const int* flag, int epoch) {
  if (ld_acquire_sys(flag) - epoch >= 0) return true;
  const unsigned long long t0 = globaltimer_ns();
  while (ld_acquire_sys(flag) - epoch < 0) {
    __nanosleep(200);
    if (globaltimer_ns() - t0 > PEER_TIMEOUT_NS) return false;
  }
  return true;
}

// local_state: [0] epoch of the last completed call, [1] CTA ticket, [2] error (1 = barrier-in timed out, 2 = barrier-out)
// flags block of a rank: ready[DA_MAX_PEERS] | done[DA_MAX_PEERS]; slot r is written by rank r only.
template <int W>
__global__ void __launch_bounds__(PEER_THREADS, 1)
sgd_step_peer_kernel(const da_peer_sgd_args a, float lr, float mu, float wd, int first) {
  __shared__ int s_epoch;
  const int world = W > 0 ? W : a.world;
  if (threadIdx.x == 0) s_epoch = a.local_state[0] + 1;
  __syncthreads();
  const int epoch = s_epoch;
  // ---- barrier-in: tell every peer that this rank's gradient is complete, wait for theirs
  if (blockIdx.x == 0 && threadIdx.x < world && threadIdx.x != a.rank)
    st_release_sys(a.flags[threadIdx.x] + a.rank, epoch);
  if (threadIdx.x < world && threadIdx.x != a.rank) {
    if (!spin_until(a.flags[a.rank] + threadIdx.x, epoch)) atomicExch(a.local_state + 2, 1);
  }
  __syncthreads();

  // ---- own slice [lo, hi): multiples of 1024 elements, the last rank takes the remainder
  const int64_t per = ((a.n + world - 1) / world + 1023) / 1024 * 1024;
  const int64_t lo = per * a.rank < a.n ? per * a.rank : a.n;
  const int64_t hi = lo + per < a.n ? lo + per : a.n;
  const float inv = 1.f / (float)world;
  const int64_t nvec = (hi - lo) >> 3;   // 8 elements per thread and iteration
  for (int64_t v = (int64_t)blockIdx.x * PEER_THREADS + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * PEER_THREADS) {
    const int64_t i = lo + (v << 3);
    float4 g0[W > 0 ? W : 1], g1[W > 0 ? W : 1];
    float4 s0, s1;
    if (W > 0) {
#pragma unroll
      for (int r = 0; r < W; ++r) { g0[r] = ld_cv4(a.grad[r] + i); g1[r] = ld_cv4(a.grad[r] + i + 4); }
    }
    float4 w0 = *reinterpret_cast<const float4*>(a.w + i), w1 = *reinterpret_cast<const float4*>(a.w + i + 4);
    float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
    if (!first) {
      b0 = *reinterpret_cast<const float4*>(a.momentum_shard + (i - lo));
      b1 = *reinterpret_cast<const float4*>(a.momentum_shard + (i - lo) + 4);
    }
    if (W > 0) {
      s0 = g0[0]; s1 = g1[0];
#pragma unroll
      for (int r = 1; r < W; ++r) {
        s0.x += g0[r].x; s0.y += g0[r].y; s0.z += g0[r].z; s0.w += g0[r].w;
        s1.x += g1[r].x; s1.y += g1[r].y; s1.z += g1[r].z; s1.w += g1[r].w;
      }
    } else {
      s0 = ld_cv4(a.grad[0] + i); s1 = ld_cv4(a.grad[0] + i + 4);
      for (int r = 1; r < world; ++r) {
        const float4 t0 = ld_cv4(a.grad[r] + i), t1 = ld_cv4(a.grad[r] + i + 4);
        s0.x += t0.x; s0.y += t0.y; s0.z += t0.z; s0.w += t0.w;
        s1.x += t1.x; s1.y += t1.y; s1.z += t1.z; s1.w += t1.w;
      }
    }
    float* wp0 = &w0.x; float* wp1 = &w1.x; float* bp0 = &b0.x; float* bp1 = &b1.x;
    const float* gp0 = &s0.x; const float* gp1 = &s1.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d0 = fmaf(wd, wp0[j], gp0[j] * inv), d1 = fmaf(wd, wp1[j], gp1[j] * inv);
      bp0[j] = first ? d0 : fmaf(mu, bp0[j], d0);
      bp1[j] = first ? d1 : fmaf(mu, bp1[j], d1);
      wp0[j] = fmaf(-lr, bp0[j], wp0[j]);
      wp1[j] = fmaf(-lr, bp1[j], wp1[j]);
    }
    *reinterpret_cast<float4*>(a.w + i) = w0;
    *reinterpret_cast<float4*>(a.w + i + 4) = w1;
    *reinterpret_cast<float4*>(a.momentum_shard + (i - lo)) = b0;
    *reinterpret_cast<float4*>(a.momentum_shard + (i - lo) + 4) = b1;
    __nv_bfloat162 p0 = __floats2bfloat162_rn(w0.x, w0.y), p1 = __floats2bfloat162_rn(w0.z, w0.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(w1.x, w1.y), p3 = __floats2bfloat162_rn(w1.z, w1.w);
    const uint4 u = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                               *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
    // publish, own rank first, then the peers starting at rank+1 (spreads the instantaneous load over the ports)
#pragma unroll
    for (int k = 0; k < (W > 0 ? W : DA_MAX_PEERS); ++k) {
      if (k >= world) break;
      int r = a.rank + k;
      if (r >= world) r -= world;
      if (a.w_bf16[r]) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.w_bf16[r]) + i) = u;
      if (k > 0 && a.w_f32[r]) {
        *reinterpret_cast<float4*>(a.w_f32[r] + i) = w0;
        *reinterpret_cast<float4*>(a.w_f32[r] + i + 4) = w1;
      }
    }
  }
  // tail of the slice (< 8 elements; only when n is not a multiple of 8)
  if (blockIdx.x == 0) {
    const int64_t i = lo + (nvec << 3) + threadIdx.x;
    if (i < hi) {
      float g = 0.f;
      for (int r = 0; r < world; ++r) g += __ldcv(a.grad[r] + i);
      const float d = fmaf(wd, a.w[i], g * inv);
      const float b = first ? d : fmaf(mu, a.momentum_shard[i - lo], d);
      a.momentum_shard[i - lo] = b;
      const float wn = fmaf(-lr, b, a.w[i]);
      a.w[i] = wn;
      for (int r = 0; r < world; ++r) {
        if (a.w_bf16[r]) reinterpret_cast<__nv_bfloat16*>(a.w_bf16[r])[i] = __float2bfloat16_rn(wn);
        if (r != a.rank && a.w_f32[r]) a.w_f32[r][i] = wn;
      }
    }
  }
  // ---- barrier-out, sender side: all stores of this rank are visible system-wide before done[rank] is raised
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ticket = atomicAdd(a.local_state + 1, 1);
    s_epoch = (ticket == (int)gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_epoch) {   // last CTA of this rank
    if (threadIdx.x < world) st_release_sys(a.flags[threadIdx.x] + DA_MAX_PEERS + a.rank, epoch);
    if (threadIdx.x == 0) {
      a.local_state[1] = 0;
      __threadfence();
      a.local_state[0] = epoch;
    }
  }
}

__global__ void peer_wait_kernel(const int* flags_local, int* local_state, int world) {
  const int epoch = local_state[0];   // written by the kernel in front of this one (stream order)
  if (threadIdx.x < world) {
    if (!spin_until(flags_local + DA_MAX_PEERS + threadIdx.x, epoch)) atomicExch(local_state + 2, 2);
  }
}

}  // namespace da

using namespace da;

extern "C" int da_peer_alloc(size_t bytes, void** out) {
  DA_REQUIRE(out && bytes > 0, DA_ERR_INVALID_ARG, "peer_alloc: bad args");
  DA_CUDA_OK(cudaMalloc(out, bytes));
  DA_CUDA_OK(cudaMemset(*out, 0, bytes));
  DA_CUDA_OK(cudaDeviceSynchronize());
  return DA_OK;
}
extern "C" int da_peer_free(void* p) {
  if (p) DA_CUDA_OK(cudaFree(p));
  return DA_OK;
}
extern "C" int da_peer_export(const void* p, unsigned char* handle64) {
  DA_REQUIRE(p && handle64, DA_ERR_INVALID_ARG, "peer_export: bad args");
  static_assert(sizeof(cudaIpcMemHandle_t) == DA_PEER_HANDLE_BYTES, "handle size");
  cudaIpcMemHandle_t h;
  DA_CUDA_OK(cudaIpcGetMemHandle(&h, const_cast<void*>(p)));
  memcpy(handle64, &h, sizeof(h));
  return DA_OK;
}
extern "C" int da_peer_open(const unsigned char* handle64, void** out) {
  DA_REQUIRE(handle64 && out, DA_ERR_INVALID_ARG, "peer_open: bad args");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  DA_CUDA_OK(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return DA_OK;
}
extern "C" int da_peer_close(void* p) {
  if (p) DA_CUDA_OK(cudaIpcCloseMemHandle(p));
  return DA_OK;
}

extern "C" int da_sgd_step_peer(const da_peer_sgd_args* a, float lr, float momentum, float weight_decay, int first_step,
                                int max_ctas, da_stream_t stream) {
  DA_REQUIRE(a && a->w && a->momentum_shard && a->local_state && a->n > 0, DA_ERR_INVALID_ARG, "sgd_step_peer: bad args");
  DA_REQUIRE(a->world >= 1 && a->world <= DA_MAX_PEERS && a->rank >= 0 && a->rank < a->world, DA_ERR_INVALID_ARG,
             "sgd_step_peer: world %d / rank %d out of range (max %d peers)", a->world, a->rank, DA_MAX_PEERS);
  uintptr_t bits = (uintptr_t)a->w | (uintptr_t)a->momentum_shard;
  for (int r = 0; r < a->world; ++r) {
    DA_REQUIRE(a->grad[r] && a->flags[r], DA_ERR_INVALID_ARG, "sgd_step_peer: rank %d has no gradient / flag block", r);
    bits |= (uintptr_t)a->grad[r] | (uintptr_t)a->w_bf16[r] | (uintptr_t)a->w_f32[r];
  }
  DA_REQUIRE((bits & 15) == 0, DA_ERR_INVALID_ARG, "sgd_step_peer: pointers must be 16-byte aligned");
  const int64_t per = ((a->n + a->world - 1) / a->world + 1023) / 1024 * 1024;
  int ctas = (int)((per / 8 + PEER_THREADS - 1) / PEER_THREADS);
  if (max_ctas <= 0) max_ctas = num_sms_physical();
  if (ctas > max_ctas) ctas = max_ctas;
  if (ctas < 1) ctas = 1;
  cudaStream_t st = (cudaStream_t)stream;
  switch (a->world) {
    case 2: sgd_step_peer_kernel<2><<<ctas, PEER_THREADS, 0, st>>>(*a, lr, momentum, weight_decay, first_step); break;
    case 4: sgd_step_peer_kernel<4><<<ctas, PEER_THREADS, 0, st>>>(*a, lr, momentum, weight_decay, first_step); break;
    case 8: sgd_step_peer_kernel<8><<<ctas, PEER_THREADS, 0, st>>>(*a, lr, momentum, weight_decay, first_step); break;
    default: sgd_step_peer_kernel<0><<<ctas, PEER_THREADS, 0, st>>>(*a, lr, momentum, weight_decay, first_step); break;
  }
  DA_LAUNCH_CHECK();
  peer_wait_kernel<<<1, 32, 0, st>>>(a->flags[a->rank], a->local_state, a->world);
  DA_LAUNCH_CHECK();
  return DA_OK;
}
