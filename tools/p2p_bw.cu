// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/p2p_bw.cu -o gpurun_out/p2p_bw && gpurun_out/p2p_bw
// NVLink peer-memory bandwidth of plain kernels on one box (one process, two devices): which load flavour, how many
// CTAs and how many bytes in flight per thread a pull (remote read) or a push (remote write) needs.  Evidence for the
// design of csrc/peer_sgd.cu; not part of the product.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE> __device__ __forceinline__ float4 ld(const float4* p) {
  float4 v;
  if (MODE == 0) v = *p;
  else if (MODE == 1) asm volatile("ld.global.cv.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  else if (MODE == 2) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  else asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// pull: read `src` (remote), write `dst` (local).  U float4 in flight per thread.
template <int MODE, int U>
__global__ void __launch_bounds__(512) pull(const float4* __restrict__ src, float4* __restrict__ dst, long n4) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i + (U - 1) * stride < n4; i += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ld<MODE>(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) dst[i + u * stride] = v[u];
  }
}
// push: read `src` (local), write `dst` (remote)
template <int U>
__global__ void __launch_bounds__(512) push(const float4* __restrict__ src, float4* __restrict__ dst, long n4) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i + (U - 1) * stride < n4; i += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = src[i + u * stride];
#pragma unroll
    for (int u = 0; u < U; ++u) dst[i + u * stride] = v[u];
  }
}

template <typename F> static float time_ms(F launch) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  int nd = 0;
  CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
  const long bytes = 256l << 20, n4 = bytes / 16;
  float4 *loc, *rem, *loc2;
  CK(cudaSetDevice(1)); CK(cudaMalloc(&rem, bytes)); CK(cudaMemset(rem, 1, bytes)); CK(cudaDeviceSynchronize());
  CK(cudaSetDevice(0)); CK(cudaMalloc(&loc, bytes)); CK(cudaMalloc(&loc2, bytes)); CK(cudaMemset(loc, 2, bytes));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  const int ctas[] = {16, 32, 48, 64, 96, 148, 296};
  for (int c : ctas) {
    float t;
    t = time_ms([&] { pull<0, 2><<<c, 512>>>(rem, loc, n4); });  printf("pull ld      U=2 ctas=%3d  %7.1f GB/s\n", c, bytes / t / 1e6);
    t = time_ms([&] { pull<1, 2><<<c, 512>>>(rem, loc, n4); });  printf("pull ld.cv   U=2 ctas=%3d  %7.1f GB/s\n", c, bytes / t / 1e6);
    t = time_ms([&] { pull<2, 2><<<c, 512>>>(rem, loc, n4); });  printf("pull ld.nc   U=2 ctas=%3d  %7.1f GB/s\n", c, bytes / t / 1e6);
    t = time_ms([&] { pull<3, 2><<<c, 512>>>(rem, loc, n4); });  printf("pull ld.sys  U=2 ctas=%3d  %7.1f GB/s\n", c, bytes / t / 1e6);
    t = time_ms([&] { pull<0, 8><<<c, 512>>>(rem, loc, n4); });  printf("pull ld      U=8 ctas=%3d  %7.1f GB/s\n", c, bytes / t / 1e6);
    t = time_ms([&] { pull<2, 8><<<c, 512>>>(rem, loc, n4); });  printf("pull ld.nc   U=8 ctas=%3d  %7.1f GB/s\n", c, bytes / t / 1e6);
    t = time_ms([&] { push<2><<<c, 512>>>(loc, rem, n4); });     printf("push         U=2 ctas=%3d  %7.1f GB/s\n", c, bytes / t / 1e6);
    t = time_ms([&] { push<8><<<c, 512>>>(loc, rem, n4); });     printf("push         U=8 ctas=%3d  %7.1f GB/s\n", c, bytes / t / 1e6);
    t = time_ms([&] { pull<0, 8><<<c, 512>>>(loc2, loc, n4); }); printf("local copy   U=8 ctas=%3d  %7.1f GB/s\n", c, bytes / t / 1e6);
  }
  // both directions at once: pull on stream A, push on stream B
  cudaStream_t sa, sb;
  CK(cudaStreamCreate(&sa)); CK(cudaStreamCreate(&sb));
  float4* rem2;
  CK(cudaSetDevice(1)); CK(cudaMalloc(&rem2, bytes)); CK(cudaSetDevice(0));
  for (int c : {32, 64}) {
    float t = time_ms([&] { pull<2, 8><<<c, 512, 0, sa>>>(rem, loc, n4); push<8><<<c, 512, 0, sb>>>(loc2, rem2, n4); CK(cudaDeviceSynchronize()); });
    printf("pull+push concurrently ctas=%d each: %7.1f GB/s per direction (host-timed incl. sync)\n", c, bytes / t / 1e6);
  }
  CK(cudaMemcpyPeer(loc, 0, rem, 1, bytes));
  float t = time_ms([&] { CK(cudaMemcpyPeerAsync(loc, 0, rem, 1, bytes, 0)); });
  printf("cudaMemcpyPeer 1->0           %7.1f GB/s\n", bytes / t / 1e6);
  return 0;
}
