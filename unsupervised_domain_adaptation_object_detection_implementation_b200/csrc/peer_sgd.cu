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
// Two transports (publish_mode):
//   DA_PEER_PUBLISH_STORES     everything above in this one kernel (P2P loads + stores issued by the SMs).
//   DA_PEER_PUBLISH_BY_CALLER  the COPY ENGINES move the data and the kernel only touches local memory: before the call
//                              the caller pushes slice r of its gradient into rank r's staging area (da_peer_copy),
//                              grad[r] of the call points at the local staging slots, w_bf16[r != rank] is NULL; after
//                              the call it pushes the refreshed bf16 slice to every rank and raises done through
//                              da_peer_publish_done.  Measured on B200 NVLink (tools/p2p_bw.cu, profiles/): SM-issued
//                              peer loads reach 14.6 GB/s per SM and collapse to ~370 GB/s per GPU when the same kernel
//                              also pushes, SM pushes 690 GB/s, the copy engines 800 GB/s with no SM at all - so this is
//                              the mode the train step uses.
// Link traffic per rank: (N-1)/N * 4 B in + (N-1)/N * 2 B out per parameter (NCCL ring all-reduce: 2*(N-1)/N*4 B each
// way) and the optimizer pass shrinks by N.  Flags live in cudaMalloc'ed blocks exchanged through CUDA IPC
// (da_peer_alloc / da_peer_export / da_peer_open); they carry a monotonically increasing epoch kept on the device, so a
// captured CUDA graph replays correctly.  Every spin has a time-out (error flag, never a hang).
#include "da_common.cuh"
#include <string.h>

namespace da {

constexpr int PEER_THREADS = 128;
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
__device__ __forceinline__ float4 ld_stream4(const float* p) {
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
// W = world size (0: run-time), V = float4 per thread, rank and iteration.  Small CTAs (128 threads, <= 96 registers) on
// purpose: they fit NEXT to the resident CTA of the tensor-core kernels this call overlaps (RoIAlign backward: 576
// threads x 87 registers; the persistent GEMMs: 320 x 112), so the transfer takes no SM away from them.  NVLink loads are
// limited per SM (tools/p2p_bw.cu: 14.6 GB/s per SM whatever the unroll), so the grid covers every SM twice.
template <int W, int V>
__global__ void __launch_bounds__(PEER_THREADS, 5)
sgd_step_peer_kernel(const da_peer_sgd_args a, float lr, float mu, float wd, int first, int signal_done) {
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
  // caller-published mode: this rank's previous bf16 slice must have left (the copies read what this kernel rewrites);
  // local_state[3] counts the publishes completed on this rank (peer_signal_done_kernel, behind the copies in stream order)
  if (!signal_done && world > 1 && threadIdx.x == world) {
    if (!spin_until(a.local_state + 3, epoch - 1)) atomicExch(a.local_state + 2, 3);
  }
  __syncthreads();

  // ---- own slice [lo, hi): multiples of 1024 elements, the last rank takes the remainder
  const int64_t per = ((a.n + world - 1) / world + 1023) / 1024 * 1024;
  const int64_t lo = per * a.rank < a.n ? per * a.rank : a.n;
  const int64_t hi = lo + per < a.n ? lo + per : a.n;
  const float inv = 1.f / (float)world;
  constexpr int TILE = PEER_THREADS * 4 * V;             // elements per CTA and iteration; thread t owns float4 t + k*128
  const int64_t ntile = (hi - lo) / TILE;
  for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int64_t i0 = lo + tile * TILE + 4 * threadIdx.x;
    float4 s[V], wv[V], bv[V];
    if (W > 0) {
      float4 g[W > 0 ? W : 1][V];
#pragma unroll
      for (int r = 0; r < W; ++r)
#pragma unroll
        for (int k = 0; k < V; ++k) g[r][k] = ld_stream4(a.grad[r] + i0 + k * (4 * PEER_THREADS));
#pragma unroll
      for (int k = 0; k < V; ++k) {
        wv[k] = *reinterpret_cast<const float4*>(a.w + i0 + k * (4 * PEER_THREADS));
        bv[k] = first ? make_float4(0.f, 0.f, 0.f, 0.f)
                      : *reinterpret_cast<const float4*>(a.momentum_shard + (i0 - lo) + k * (4 * PEER_THREADS));
      }
#pragma unroll
      for (int k = 0; k < V; ++k) {
        s[k] = g[0][k];
#pragma unroll
        for (int r = 1; r < W; ++r) { s[k].x += g[r][k].x; s[k].y += g[r][k].y; s[k].z += g[r][k].z; s[k].w += g[r][k].w; }
      }
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        wv[k] = *reinterpret_cast<const float4*>(a.w + i0 + k * (4 * PEER_THREADS));
        bv[k] = first ? make_float4(0.f, 0.f, 0.f, 0.f)
                      : *reinterpret_cast<const float4*>(a.momentum_shard + (i0 - lo) + k * (4 * PEER_THREADS));
        s[k] = ld_stream4(a.grad[0] + i0 + k * (4 * PEER_THREADS));
      }
      for (int r = 1; r < world; ++r) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          const float4 t = ld_stream4(a.grad[r] + i0 + k * (4 * PEER_THREADS));
          s[k].x += t.x; s[k].y += t.y; s[k].z += t.z; s[k].w += t.w;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float* wp = &wv[k].x; float* bp = &bv[k].x; const float* gp = &s[k].x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d = fmaf(wd, wp[j], gp[j] * inv);
        bp[j] = first ? d : fmaf(mu, bp[j], d);
        wp[j] = fmaf(-lr, bp[j], wp[j]);
      }
      const int64_t i = i0 + k * (4 * PEER_THREADS);
      *reinterpret_cast<float4*>(a.w + i) = wv[k];
      *reinterpret_cast<float4*>(a.momentum_shard + (i - lo)) = bv[k];
    }
    // publish, own rank first, then the peers starting at rank+1 (spreads the instantaneous load over the ports)
#pragma unroll
    for (int q = 0; q < (W > 0 ? W : DA_MAX_PEERS); ++q) {
      if (q >= world) break;
      int r = a.rank + q;
      if (r >= world) r -= world;
      __nv_bfloat16* sh = reinterpret_cast<__nv_bfloat16*>(a.w_bf16[r]);
      float* mf = q > 0 ? a.w_f32[r] : nullptr;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const int64_t i = i0 + k * (4 * PEER_THREADS);
        if (sh) {
          __nv_bfloat162 p0 = __floats2bfloat162_rn(wv[k].x, wv[k].y), p1 = __floats2bfloat162_rn(wv[k].z, wv[k].w);
          *reinterpret_cast<uint2*>(sh + i) = make_uint2(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1));
        }
        if (mf) *reinterpret_cast<float4*>(mf + i) = wv[k];
      }
    }
  }
  // tail of the slice (< one tile): one element per thread
  for (int64_t i = lo + ntile * TILE + (int64_t)blockIdx.x * PEER_THREADS + threadIdx.x; i < hi; i += (int64_t)gridDim.x * PEER_THREADS) {
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
  // ---- barrier-out, sender side: all stores of this rank are visible system-wide before done[rank] is raised
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ticket = atomicAdd(a.local_state + 1, 1);
    s_epoch = (ticket == (int)gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_epoch) {   // last CTA of this rank
    if (signal_done && threadIdx.x < world) st_release_sys(a.flags[threadIdx.x] + DA_MAX_PEERS + a.rank, epoch);
    if (threadIdx.x == 0) {
      a.local_state[1] = 0;
      __threadfence();
      a.local_state[0] = epoch;
    }
  }
}

// done[rank] := number of publishes this rank has completed, on every rank (the caller's copies are in front of this kernel
// in stream order).  The count is kept in local_state[3], NOT derived from the update epoch: a deferred publish may run
// while the next step is already under way.
__global__ void peer_signal_done_kernel(const da_peer_sgd_args a) {
  const int epoch = a.local_state[3] + 1;
  if (threadIdx.x < a.world) st_release_sys(a.flags[threadIdx.x] + DA_MAX_PEERS + a.rank, epoch);
  __syncwarp();
  if (threadIdx.x == 0) {
    __threadfence();
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.local_state + 3), "r"(epoch) : "memory");
  }
}

__global__ void peer_wait_kernel(const int* flags_local, int* local_state, int world) {
  const int epoch = local_state[0];   // written by the kernel in front of this one (stream order)
  if (threadIdx.x < world) {
    if (!spin_until(flags_local + DA_MAX_PEERS + threadIdx.x, epoch)) atomicExch(local_state + 2, 2);
  }
}

// CUDA loads kernels lazily, and loading one can wait for the kernels that are running: a spinning peer kernel in front
// of the FIRST launch of another peer kernel would stall the host until the time-out.  Load them all up front.
static void preload_peer_kernels() {
  static bool done = false;
  if (done) return;
  done = true;
  cudaFuncAttributes fa;
  (void)cudaFuncGetAttributes(&fa, sgd_step_peer_kernel<1, 4>);
  (void)cudaFuncGetAttributes(&fa, sgd_step_peer_kernel<2, 4>);
  (void)cudaFuncGetAttributes(&fa, sgd_step_peer_kernel<4, 2>);
  (void)cudaFuncGetAttributes(&fa, sgd_step_peer_kernel<8, 1>);
  (void)cudaFuncGetAttributes(&fa, sgd_step_peer_kernel<0, 2>);
  (void)cudaFuncGetAttributes(&fa, sgd_step_peer_kernel<0, 1>);
  (void)cudaFuncGetAttributes(&fa, peer_signal_done_kernel);
  (void)cudaFuncGetAttributes(&fa, peer_wait_kernel);
  (void)cudaGetLastError();
}

}  // namespace da

using namespace da;

extern "C" int da_peer_alloc(size_t bytes, void** out) {
  DA_REQUIRE(out && bytes > 0, DA_ERR_INVALID_ARG, "peer_alloc: bad args");
  preload_peer_kernels();
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

extern "C" int da_peer_copy(void* dst, const void* src, size_t bytes, da_stream_t stream) {
  if (bytes == 0) return DA_OK;
  DA_REQUIRE(dst && src, DA_ERR_INVALID_ARG, "peer_copy: null pointer");
  DA_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return DA_OK;
}

extern "C" int da_peer_signal_done(const da_peer_sgd_args* a, da_stream_t stream) {
  DA_REQUIRE(a && a->local_state && a->world >= 1 && a->world <= DA_MAX_PEERS && a->rank >= 0 && a->rank < a->world,
             DA_ERR_INVALID_ARG, "peer_signal_done: bad args");
  preload_peer_kernels();
  peer_signal_done_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*a);
  DA_LAUNCH_CHECK();
  return DA_OK;
}

extern "C" int da_peer_wait_done(const da_peer_sgd_args* a, da_stream_t stream) {
  DA_REQUIRE(a && a->local_state && a->world >= 1 && a->world <= DA_MAX_PEERS && a->rank >= 0 && a->rank < a->world,
             DA_ERR_INVALID_ARG, "peer_wait_done: bad args");
  preload_peer_kernels();
  peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a->flags[a->rank], a->local_state, a->world);
  DA_LAUNCH_CHECK();
  return DA_OK;
}

extern "C" int da_peer_publish_done(const da_peer_sgd_args* a, da_stream_t stream) {
  DA_REQUIRE(a && a->local_state && a->world >= 1 && a->world <= DA_MAX_PEERS && a->rank >= 0 && a->rank < a->world,
             DA_ERR_INVALID_ARG, "peer_publish_done: bad args");
  peer_signal_done_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*a);
  DA_LAUNCH_CHECK();
  peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a->flags[a->rank], a->local_state, a->world);
  DA_LAUNCH_CHECK();
  return DA_OK;
}

extern "C" int da_sgd_step_peer(const da_peer_sgd_args* a, float lr, float momentum, float weight_decay, int first_step,
                                int max_ctas, int publish_mode, da_stream_t stream) {
  DA_REQUIRE(publish_mode == DA_PEER_PUBLISH_STORES || publish_mode == DA_PEER_PUBLISH_BY_CALLER, DA_ERR_INVALID_ARG,
             "sgd_step_peer: publish_mode %d", publish_mode);
  const int sd = publish_mode == DA_PEER_PUBLISH_STORES ? 1 : 0;
  preload_peer_kernels();
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
  const int v = a->world <= 2 ? 4 : (a->world <= 4 ? 2 : 1);
  int ctas = (int)((per + PEER_THREADS * 4 * v - 1) / (PEER_THREADS * 4 * v));
  if (max_ctas <= 0) max_ctas = (sd ? 2 : 16) * num_sms_physical();   // all-local update: enough CTAs to stream HBM
  if (ctas > max_ctas) ctas = max_ctas;
  if (ctas < 1) ctas = 1;
  cudaStream_t st = (cudaStream_t)stream;
  switch (a->world) {
    case 1: sgd_step_peer_kernel<1, 4><<<ctas, PEER_THREADS, 0, st>>>(*a, lr, momentum, weight_decay, first_step, sd); break;
    case 2: sgd_step_peer_kernel<2, 4><<<ctas, PEER_THREADS, 0, st>>>(*a, lr, momentum, weight_decay, first_step, sd); break;
    case 4: sgd_step_peer_kernel<4, 2><<<ctas, PEER_THREADS, 0, st>>>(*a, lr, momentum, weight_decay, first_step, sd); break;
    case 8: sgd_step_peer_kernel<8, 1><<<ctas, PEER_THREADS, 0, st>>>(*a, lr, momentum, weight_decay, first_step, sd); break;
    default:
      if (v == 2) sgd_step_peer_kernel<0, 2><<<ctas, PEER_THREADS, 0, st>>>(*a, lr, momentum, weight_decay, first_step, sd);
      else sgd_step_peer_kernel<0, 1><<<ctas, PEER_THREADS, 0, st>>>(*a, lr, momentum, weight_decay, first_step, sd);
      break;
  }
  DA_LAUNCH_CHECK();
  if (sd) {
    peer_wait_kernel<<<1, 32, 0, st>>>(a->flags[a->rank], a->local_state, a->world);
    DA_LAUNCH_CHECK();
  }
  return DA_OK;
}
