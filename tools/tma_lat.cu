// Latency / concurrency microbenchmark of cp.async.bulk (global -> shared, mbarrier completion) and ld.global on B200.
// One CTA per SM; thread 0 keeps DEPTH copies of BYTES bytes in flight from pseudo-random chunk-aligned offsets of a buffer of
// FOOT bytes and reports mean SM cycles per copy (DEPTH = 1: the latency).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void expect_tx(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void wait(uint32_t b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W;\n}" ::"r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t n, uint32_t b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(n), "r"(b) : "memory");
}
template <int DEPTH>
__global__ void k_bulk(const uint8_t* buf, size_t chunks, uint32_t chunk_bytes, uint32_t bytes, int iters, long long* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t bars[DEPTH];
  if (threadIdx.x == 0) {
    for (int i = 0; i < DEPTH; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    uint64_t x = 0x9E3779B97F4A7C15ull * (blockIdx.x + 1);
    auto next = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (size_t)(x % chunks); };
    for (int i = 0; i < DEPTH; ++i) { expect_tx(smem_u32(&bars[i]), bytes); bulk(smem_u32(sm + (size_t)i * 32768), buf + next() * chunk_bytes, bytes, smem_u32(&bars[i])); }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int s = it % DEPTH;
      wait(smem_u32(&bars[s]), (it / DEPTH) & 1);
      expect_tx(smem_u32(&bars[s]), bytes);
      bulk(smem_u32(sm + (size_t)s * 32768), buf + next() * chunk_bytes, bytes, smem_u32(&bars[s]));
    }
    long long t1 = clock64();
    for (int i = 0; i < DEPTH; ++i) wait(smem_u32(&bars[(iters + i) % DEPTH]), ((iters + i) / DEPTH) & 1);
    out[blockIdx.x] = (t1 - t0) / iters;
  }
}
// one warp: DEPTH dependent-free 16-byte loads per lane per round from a random chunk, consumed before the next round
__global__ void k_ldg(const uint4* buf, size_t chunks, uint32_t chunk_bytes, int per_lane, int iters, long long* out) {
  uint64_t x = 0x9E3779B97F4A7C15ull * (blockIdx.x + 1);
  auto next = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (size_t)(x % chunks); };
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint4* p = buf + next() * (chunk_bytes / 16) + threadIdx.x;
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) if (k < per_lane) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w) : "l"(p + k * blockDim.x));
#pragma unroll
    for (int k = 0; k < 8; ++k) if (k < per_lane) acc += v[k].x;
    x += acc & 1;   // serialise rounds on the loaded data
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters + (acc == 0x12345678u);
}
int main() {
  const int sms = 148, iters = 2000;
  long long* out; cudaMallocManaged(&out, sms * sizeof(long long));
  auto mean = [&]() { double s = 0; for (int i = 0; i < sms; ++i) s += out[i]; return s / sms; };
  for (size_t foot : {(size_t)32 << 20, (size_t)1024 << 20}) {
    uint8_t* buf; cudaMalloc(&buf, foot); cudaMemset(buf, 1, foot);
    const uint32_t chunk = 32768; const size_t chunks = foot / chunk;
    for (uint32_t bytes : {16u, 4096u, 25088u, 32768u}) {
#define RUN(D) { cudaFuncSetAttribute(k_bulk<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, D * 32768); \
        k_bulk<D><<<sms, 32, D * 32768>>>(buf, chunks, chunk, bytes, iters, out); cudaError_t e = cudaDeviceSynchronize(); \
        printf("bulk foot %4zu MB bytes %5u depth %d: %.0f cycles/copy%s\n", foot >> 20, bytes, D, mean(), e ? cudaGetErrorString(e) : ""); }
      RUN(1) RUN(2) RUN(3) RUN(6)
    }
    for (int threads : {32, 512}) for (int per_lane : {1, 4}) {
      k_ldg<<<sms, threads>>>((const uint4*)buf, chunks, chunk, per_lane, iters, out); cudaDeviceSynchronize();
      printf("ldg  foot %4zu MB threads %3d x %d x16B: %.0f cycles/round\n", foot >> 20, threads, per_lane, mean());
    }
    cudaFree(buf);
  }
  return 0;
}
