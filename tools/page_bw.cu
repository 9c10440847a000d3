// HBM bandwidth of the fused-SGD access pattern as a function of the contiguous bytes touched per row visit.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/page_bw.bin tools/page_bw.cu && tools/page_bw.bin
// Two fp32 arrays [ROWS][COLS] are read and written (master, momentum) + a bf16 array written: 18 B per element, like
// da_conv_backward_weight_sgd.  A CTA owns a tile of 128 rows x SEG bytes; tiles are dealt round-robin over a persistent grid
// (rows fastest, like the kernel: neighbouring CTAs take neighbouring row blocks of the same column block).
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int SEG>   // bytes per row per tile (128 .. 4096), 256 threads
__global__ void tile_update(float* __restrict__ w, float* __restrict__ m, __nv_bfloat16* __restrict__ s, int rows, int cols, long long tiles) {
  constexpr int F4 = SEG / 16;                 // float4 per row segment
  const int rb_count = rows / 128;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int rb = (int)(t % rb_count);
    const long long cb = t / rb_count;
    // 128 rows x F4 float4: thread -> (row, f4) with f4 fastest
    for (int i = threadIdx.x; i < 128 * F4; i += 256) {
      const int r = i / F4, f = i % F4;
      const size_t off = ((size_t)(rb * 128 + r) * cols + (size_t)cb * (SEG / 4)) + (size_t)f * 4;
      float4 a = *reinterpret_cast<const float4*>(w + off);
      float4 b = *reinterpret_cast<const float4*>(m + off);
      b.x = 0.9f * b.x + a.x * 1e-4f; b.y = 0.9f * b.y + a.y * 1e-4f; b.z = 0.9f * b.z + a.z * 1e-4f; b.w = 0.9f * b.w + a.w * 1e-4f;
      a.x -= 0.01f * b.x; a.y -= 0.01f * b.y; a.z -= 0.01f * b.z; a.w -= 0.01f * b.w;
      *reinterpret_cast<float4*>(w + off) = a;
      *reinterpret_cast<float4*>(m + off) = b;
      __nv_bfloat162 lo = __floats2bfloat162_rn(a.x, a.y), hi = __floats2bfloat162_rn(a.z, a.w);
      *reinterpret_cast<uint2*>(s + off) = make_uint2(*reinterpret_cast<unsigned*>(&lo), *reinterpret_cast<unsigned*>(&hi));
    }
  }
}

// Same tile (ROWS_T rows x SEG bytes) but visited as SEG/128 PASSES of 128 B per row: pass p touches bytes [128p, 128p+128) of
// every row of the tile before pass p+1 starts (what a sequence of 128-byte-wide TMA boxes over the same rows does).
template <int SEG, int ROWS_T>
__global__ void tile_update_passes(float* __restrict__ w, float* __restrict__ m, __nv_bfloat16* __restrict__ s, int rows, int cols, long long tiles) {
  const int rb_count = rows / ROWS_T;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int rb = (int)(t % rb_count);
    const long long cb = t / rb_count;
    for (int p = 0; p < SEG / 128; ++p)
      for (int i = threadIdx.x; i < ROWS_T * 8; i += 256) {
        const int r = i / 8, f = i % 8;
        const size_t off = ((size_t)(rb * ROWS_T + r) * cols + (size_t)cb * (SEG / 4)) + (size_t)p * 32 + (size_t)f * 4;
        float4 a = *reinterpret_cast<const float4*>(w + off);
        float4 b = *reinterpret_cast<const float4*>(m + off);
        b.x = 0.9f * b.x + a.x * 1e-4f; b.y = 0.9f * b.y + a.y * 1e-4f; b.z = 0.9f * b.z + a.z * 1e-4f; b.w = 0.9f * b.w + a.w * 1e-4f;
        a.x -= 0.01f * b.x; a.y -= 0.01f * b.y; a.z -= 0.01f * b.z; a.w -= 0.01f * b.w;
        *reinterpret_cast<float4*>(w + off) = a;
        *reinterpret_cast<float4*>(m + off) = b;
        __nv_bfloat162 lo = __floats2bfloat162_rn(a.x, a.y), hi = __floats2bfloat162_rn(a.z, a.w);
        *reinterpret_cast<uint2*>(s + off) = make_uint2(*reinterpret_cast<unsigned*>(&lo), *reinterpret_cast<unsigned*>(&hi));
      }
  }
}
template <int SEG, int ROWS_T>
static int run_passes(float* w, float* m, __nv_bfloat16* s, int rows, int cols, int ctas_per_sm) {
  const long long tiles = (long long)(rows / ROWS_T) * ((long long)cols * 4 / SEG);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * ctas_per_sm;
  tile_update_passes<SEG, ROWS_T><<<grid, 256>>>(w, m, s, rows, cols, tiles);
  CK(cudaDeviceSynchronize());
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    cudaEventRecord(e0);
    tile_update_passes<SEG, ROWS_T><<<grid, 256>>>(w, m, s, rows, cols, tiles);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  printf("passes: tile %3d rows x %4d B as %d x 128 B  ctas/SM %d : %.3f ms  %.0f GB/s\n", ROWS_T, SEG, SEG / 128, ctas_per_sm, best,
         (double)rows * cols * 18.0 / best / 1e6);
  return 0;
}

template <int SEG>
static int run(float* w, float* m, __nv_bfloat16* s, int rows, int cols, int ctas_per_sm) {
  const long long tiles = (long long)(rows / 128) * ((long long)cols * 4 / SEG);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * ctas_per_sm;
  tile_update<SEG><<<grid, 256>>>(w, m, s, rows, cols, tiles);
  CK(cudaDeviceSynchronize());
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    cudaEventRecord(e0);
    tile_update<SEG><<<grid, 256>>>(w, m, s, rows, cols, tiles);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double bytes = (double)rows * cols * 18.0;
  printf("seg %5d B/row  ctas/SM %d : %.3f ms  %.0f GB/s\n", SEG, ctas_per_sm, best, bytes / best / 1e6);
  return 0;
}

int main() {
  const int rows = 1024, cols = 100352;
  float *w, *m; __nv_bfloat16* s;
  CK(cudaMalloc(&w, (size_t)rows * cols * 4)); CK(cudaMalloc(&m, (size_t)rows * cols * 4)); CK(cudaMalloc(&s, (size_t)rows * cols * 2));
  CK(cudaMemset(w, 0, (size_t)rows * cols * 4)); CK(cudaMemset(m, 0, (size_t)rows * cols * 4));
  for (int c : {2, 4}) {
    if (run_passes<512, 32>(w, m, s, rows, cols, c)) return 1;
    if (run_passes<512, 128>(w, m, s, rows, cols, c)) return 1;
    if (run_passes<1024, 32>(w, m, s, rows, cols, c)) return 1;
    if (run_passes<256, 128>(w, m, s, rows, cols, c)) return 1;
  }
  for (int c : {4}) {
    if (run<128>(w, m, s, rows, cols, c)) return 1;
    if (run<256>(w, m, s, rows, cols, c)) return 1;
    if (run<512>(w, m, s, rows, cols, c)) return 1;
    if (run<1024>(w, m, s, rows, cols, c)) return 1;
    if (run<4096>(w, m, s, rows, cols, c)) return 1;
  }
  return 0;
}
