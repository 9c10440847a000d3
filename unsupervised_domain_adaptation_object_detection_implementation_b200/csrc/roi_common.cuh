// Shared definitions of the RoIAlign kernels (roi_align.cu, roi_align_tc.cu).
#pragma once
#include "da_common.cuh"

namespace da {

constexpr int P = 7;          // pooled size supported by the register-tiled kernels
constexpr int PP = P * P;     // 49
constexpr int WROW = 8;       // floats per weight-table row (7 bins + pad)

struct __align__(16) RoiMeta {
  int b;        // batch index, -1 when invalid / out of range
  int gh, gw;   // sampling grid (roi_bin_grid_h / _w)
  int y_lo, ny; // first footprint row, number of rows (0 = empty)
  int x_lo, nx;
  int count;    // max(gh*gw,1)
};

// workspace layout: int hdr[4] = {error flag, work-queue counter, -, -} | RoiMeta[R] | float tables[R][(H+W)*8] | int order[R]
// (hdr is zeroed before every prep launch; order = RoIs by decreasing footprint, the tensor-core forward's work queue)
__host__ __device__ inline size_t ws_meta_off() { return 16; }
__host__ __device__ inline size_t ws_table_off(int R) {
  return 16 + ((size_t)R * sizeof(RoiMeta) + 255) / 256 * 256;
}
__host__ __device__ inline size_t ws_order_off(int R, int H, int W) {
  return (ws_table_off(R) + (size_t)R * (size_t)(H + W) * WROW * sizeof(float) + 255) / 256 * 256;
}


// tensor-core forward (roi_align_tc.cu) handles every RoI with a non-empty footprint
__host__ __device__ inline bool roi_tc_eligible(const RoiMeta& m) { return m.ny > 0 && m.nx > 0; }

int roi_align_bwd_tc(const void* grad_out, int layout, int N, int C, int H, int W, int R, const void* ws, void* grad_in, int grad_in_dtype, cudaStream_t st);
int roi_align_fwd_tc(const void* feat, int N, int C, int H, int W, int R, const void* ws, void* out, int out_dtype, int layout,
                     cudaStream_t st);

}  // namespace da
