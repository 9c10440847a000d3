// RPN proposal stage of one image and one feature level — sm_100a.
//
// Reference: RPNHeadDA._get_bboxes_single / _bbox_post_process (mmdet/models/dense_heads/rpn_head_da.py:170-303):
//   scores = sigmoid(cls.permute(1,2,0).reshape(-1)); sort descending, keep nms_pre; anchors / deltas gathered;
//   DeltaXYWHBBoxCoder.decode = delta2bbox (mmdet/core/bbox/coder/delta_xywh_bbox_coder.py:224-259) with clipping to the image;
//   boxes with w or h <= min_bbox_size dropped; batched_nms (mmcv.ops.nms, one level => plain NMS, IoU > thr suppresses);
//   first max_per_img survivors as [x1,y1,x2,y2,score].
// The reference does this with ~15 ATen launches, two gathers of [n,4] tensors, an NMS whose result is read back by the host, and
// a Python loop over images.  Here:
//   rpn_scores_kernel   sigmoid + the (A,H,W) -> (H,W,A) reorder in one pass (the order the reference ranks in)
//   [the ranking itself is torch.sort(stable) on the device — library code, stated in DESIGN.md]
//   rpn_decode_kernel   anchors are never materialised: anchor(idx) = base[idx % A] + stride * (cell % W, cell / W); deltas are read
//                       from the conv output where it lies ([4A,H,W]); delta2bbox with the reference's fp32 operation order
//                       (separate multiply / add roundings, no FMA contraction); validity flag instead of a compaction
//   nms_mask_kernel     64 x 64 tiles of the upper triangle of the suppression matrix, one bit per (i, j > i) pair
//   nms_scan_kernel     ONE CTA walks the boxes in rank order, 64 at a time: a single thread resolves the 64 x 64 diagonal
//                       tile, all threads OR the kept rows into the running "removed" bitmap; writes the first max_out kept
//                       boxes and their count.  No host round trip anywhere: the count stays on the device.
// Index work (ranking order, keep set) is bit-exact against the oracle given the same boxes; decoded coordinates are fp32 with
// the reference's rounding points (expf is the only library call).
#include "da_common.cuh"

namespace da {

__global__ void rpn_scores_kernel(const float* __restrict__ cls, int A, int HW, float* __restrict__ scores) {
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // output index = cell * A + a
  if (i >= (long long)A * HW) return;
  const int cell = (int)(i / A), a = (int)(i - (long long)cell * A);
  const float x = cls[(size_t)a * HW + cell];
  scores[i] = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));
}

struct DecodeParams {
  float mean[4], stdv[4];
  float max_ratio, img_h, img_w, min_size, stride;
  int A, H, W;
};

__global__ void rpn_decode_kernel(const float* __restrict__ reg, const float* __restrict__ base, const long long* __restrict__ top_idx,
                                  int n, DecodeParams P, float4* __restrict__ boxes, unsigned char* __restrict__ valid) {
  pdl_launch_dependents();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long idx = top_idx[i];
  const int HW = P.H * P.W;
  const int cell = (int)(idx / P.A), a = (int)(idx - (long long)cell * P.A);
  const int cy = cell / P.W, cx = cell - cy * P.W;
  const float sx = __fmul_rn((float)cx, P.stride), sy = __fmul_rn((float)cy, P.stride);
  // grid_anchors: shifts + base (one fp32 add per coordinate)
  const float ax1 = __fadd_rn(sx, base[a * 4 + 0]), ay1 = __fadd_rn(sy, base[a * 4 + 1]);
  const float ax2 = __fadd_rn(sx, base[a * 4 + 2]), ay2 = __fadd_rn(sy, base[a * 4 + 3]);
  float d[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) d[j] = __fadd_rn(__fmul_rn(reg[(size_t)(a * 4 + j) * HW + cell], P.stdv[j]), P.mean[j]);
  const float px = __fmul_rn(__fadd_rn(ax1, ax2), 0.5f), py = __fmul_rn(__fadd_rn(ay1, ay2), 0.5f);
  const float pw = __fsub_rn(ax2, ax1), ph = __fsub_rn(ay2, ay1);
  const float dw = fminf(fmaxf(d[2], -P.max_ratio), P.max_ratio), dh = fminf(fmaxf(d[3], -P.max_ratio), P.max_ratio);
  const float gx = __fadd_rn(px, __fmul_rn(pw, d[0])), gy = __fadd_rn(py, __fmul_rn(ph, d[1]));
  const float gw = __fmul_rn(pw, expf(dw)), gh = __fmul_rn(ph, expf(dh));
  const float hw = __fmul_rn(gw, 0.5f), hh = __fmul_rn(gh, 0.5f);
  float x1 = __fsub_rn(gx, hw), y1 = __fsub_rn(gy, hh), x2 = __fadd_rn(gx, hw), y2 = __fadd_rn(gy, hh);
  x1 = fminf(fmaxf(x1, 0.f), P.img_w); x2 = fminf(fmaxf(x2, 0.f), P.img_w);
  y1 = fminf(fmaxf(y1, 0.f), P.img_h); y2 = fminf(fmaxf(y2, 0.f), P.img_h);
  boxes[i] = make_float4(x1, y1, x2, y2);
  valid[i] = (P.min_size < 0.f) || (__fsub_rn(x2, x1) > P.min_size && __fsub_rn(y2, y1) > P.min_size);
}

__device__ __forceinline__ bool iou_over(const float4 a, const float4 b, float thr) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z), top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float w = fmaxf(__fsub_rn(right, left), 0.f), h = fmaxf(__fsub_rn(bottom, top), 0.f);
  if (thr >= 0.f && !(w > 0.f && h > 0.f)) return false;     // inter = 0: 0 / x > thr is false for every x (0/0 = NaN included); skips the division
                                               // for the disjoint pairs, i.e. nearly all of them (ncu: math_pipe_throttle 20 %)
  const float inter = __fmul_rn(w, h);
  const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y)), sb = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter)) > thr;
}

// grid (words, words), 64 threads: block (cb, rb) with cb >= rb compares rows [64 rb, +64) with columns [64 cb, +64)
__global__ void __launch_bounds__(64)
nms_mask_kernel(const float4* __restrict__ boxes, int n, float thr, int words, unsigned long long* __restrict__ mask) {
  pdl_launch_dependents();
  const int cb = blockIdx.x, rb = blockIdx.y;
  if (cb < rb) return;
  __shared__ float4 cbox[64];
  const int t = threadIdx.x;
  const int cj = cb * 64 + t;
  if (cj < n) cbox[t] = boxes[cj];
  __syncthreads();
  const int ri = rb * 64 + t;
  if (ri >= n) return;
  const float4 me = boxes[ri];
  const int ncol = min(64, n - cb * 64);
  unsigned long long bits = 0ull;
  const int start = (cb == rb) ? t + 1 : 0;
  for (int j = start; j < ncol; ++j)
    if (iou_over(me, cbox[j], thr)) bits |= 1ull << j;
  mask[(size_t)ri * words + cb] = bits;
}

// 32 warps: the OR phase hands one kept row to each warp, and with dense survivors (up to 64 kept rows per block) its L2 round
// trips are what a block costs (8 warps: 4 us per block under ncu; a single-warp variant with the bitmap in registers was slower still)
constexpr int SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS)
nms_scan_kernel(const unsigned long long* __restrict__ mask, const unsigned char* __restrict__ valid, const float4* __restrict__ boxes,
                const float* __restrict__ scores, int n, int words, int max_out, float* __restrict__ dets, int* __restrict__ count,
                int* __restrict__ keep_idx) {
  pdl_launch_dependents();
  extern __shared__ unsigned long long remv[];          // [words]
  __shared__ unsigned long long diag[64];
  __shared__ unsigned long long s_kept;
  __shared__ unsigned s_valid[2];
  __shared__ int s_count;
  const int t = threadIdx.x;
  for (int w = t; w < words; w += SCAN_THREADS) remv[w] = 0ull;
  if (t == 0) s_count = 0;
  __syncthreads();
  const int warp = t >> 5, lane = t & 31;
  for (int b = 0; b < words; ++b) {
    if (s_count >= max_out) break;                       // uniform: s_count is only written between barriers
    const int i0 = b * 64, nb = min(64, n - i0);
    if (t < 64) {
      diag[t] = (t < nb) ? mask[(size_t)(i0 + t) * words + b] : 0ull;
      const unsigned vb = __ballot_sync(0xffffffffu, t < nb && valid[i0 + t] != 0);      // validity of the 64 boxes as two words
      if (lane == 0) s_valid[warp] = vb;
    }
    __syncthreads();
    if (t == 0) {
      unsigned long long cur = remv[b], kept = 0ull;
      const unsigned long long vmask = (unsigned long long)s_valid[0] | ((unsigned long long)s_valid[1] << 32);
      unsigned long long cand = vmask & ~cur;            // boxes of this block still alive; a kept box may clear later ones
      while (cand) {
        const int j = __ffsll((long long)cand) - 1;
        kept |= 1ull << j;
        cur |= diag[j];
        cand &= ~cur;
        cand &= ~((2ull << j) - 1ull);                   // strictly after j
      }
      s_kept = kept;
    }
    __syncthreads();
    const unsigned long long kept = s_kept;
    const int base_count = s_count;
    // kept boxes of this block go out in rank order
    if (t < 64 && ((kept >> t) & 1ull)) {
      const int pos = base_count + __popcll(kept & ((1ull << t) - 1ull));
      if (pos < max_out) {
        const float4 bx = boxes[i0 + t];
        float* o = dets + (size_t)pos * 5;
        o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w; o[4] = scores[i0 + t];
        if (keep_idx) keep_idx[pos] = i0 + t;
      }
    }
    // every later word: OR the rows of the kept boxes into the removed bitmap.  One warp per kept row (coalesced row reads, all of a
    // row's loads in flight at once), shared-memory atomicOr: OR commutes, so the result does not depend on the order.
    {
      unsigned long long k = kept;
      int r = 0;
      while (k) {
        const int j = __ffsll((long long)k) - 1;
        k &= k - 1;
        if ((r++ & (SCAN_THREADS / 32 - 1)) != warp) continue;
        const unsigned long long* row = mask + (size_t)(i0 + j) * words;
        for (int w = b + 1 + lane; w < words; w += 32) {
          const unsigned long long v = row[w];
          if (v) atomicOr(&remv[w], v);
        }
      }
    }
    __syncthreads();
    if (t == 0) s_count = base_count + __popcll(kept);
    __syncthreads();
  }
  const int total = min(s_count, max_out);
  if (t == 0) *count = total;
  for (int i = total * 5 + t; i < max_out * 5; i += SCAN_THREADS) dets[i] = 0.f;   // zero padding behind the survivors
  if (keep_idx) for (int i = total + t; i < max_out; i += SCAN_THREADS) keep_idx[i] = -1;
}

}  // namespace da

using namespace da;

extern "C" int da_rpn_scores(const float* cls, int A, int HW, float* scores, da_stream_t stream) {
  DA_REQUIRE(cls && scores && A > 0 && HW > 0, DA_ERR_INVALID_ARG, "rpn_scores: bad args (A=%d HW=%d)", A, HW);
  const long long n = (long long)A * HW;
  rpn_scores_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(cls, A, HW, scores);
  DA_LAUNCH_CHECK();
  return DA_OK;
}

static size_t rpn_ws_layout(int n, size_t* off_valid, size_t* off_mask) {
  const size_t words = (size_t)(n + 63) / 64;
  size_t o = align_up((size_t)n * sizeof(float4), 256);
  *off_valid = o;
  o = align_up(o + (size_t)n, 256);
  *off_mask = o;
  return align_up(o + (size_t)n * words * sizeof(unsigned long long), 256);
}

extern "C" size_t da_rpn_proposals_workspace_bytes(int n) {
  if (n <= 0) return 256;
  size_t a, b;
  return rpn_ws_layout(n, &a, &b);
}

extern "C" int da_rpn_proposals(const float* reg, int A, int H, int W, const float* base_anchors, float stride,
                                const int64_t* top_idx, const float* top_scores, int n,
                                const float* means4, const float* stds4, float max_ratio, float img_h, float img_w, float min_size,
                                float iou_thr, int max_out, float* dets, int32_t* count, int32_t* keep_idx,
                                void* workspace, size_t workspace_bytes, da_stream_t stream) {
  DA_REQUIRE(n >= 0 && max_out > 0 && dets && count, DA_ERR_INVALID_ARG, "rpn_proposals: bad args (n=%d max_out=%d)", n, max_out);
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {      // proposals.new_zeros(0, 5) in the reference
    DA_CUDA_OK(cudaMemsetAsync(dets, 0, (size_t)max_out * 5 * sizeof(float), st));
    DA_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
    if (keep_idx) DA_CUDA_OK(cudaMemsetAsync(keep_idx, 0xff, (size_t)max_out * sizeof(int32_t), st));
    return DA_OK;
  }
  DA_REQUIRE(reg && base_anchors && top_idx && top_scores && means4 && stds4 && A > 0 && H > 0 && W > 0, DA_ERR_INVALID_ARG,
             "rpn_proposals: null tensor or empty map");
  DA_REQUIRE(n <= A * H * W && n <= (1 << 20), DA_ERR_UNSUPPORTED, "rpn_proposals: n=%d out of range", n);
  size_t off_valid, off_mask;
  const size_t need = rpn_ws_layout(n, &off_valid, &off_mask);
  DA_REQUIRE(workspace && workspace_bytes >= need, DA_ERR_INVALID_ARG, "rpn_proposals: workspace too small (%zu < %zu)", workspace_bytes, need);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  float4* boxes = reinterpret_cast<float4*>(ws);
  unsigned char* valid = ws + off_valid;
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(ws + off_mask);
  const int words = (n + 63) / 64;
  DA_REQUIRE((size_t)words * 8 <= 200 * 1024, DA_ERR_UNSUPPORTED, "rpn_proposals: n too large for the scan kernel's bitmap");
  DecodeParams P;
  for (int j = 0; j < 4; ++j) { P.mean[j] = means4[j]; P.stdv[j] = stds4[j]; }
  P.max_ratio = max_ratio; P.img_h = img_h; P.img_w = img_w; P.min_size = min_size; P.stride = stride;
  P.A = A; P.H = H; P.W = W;
  rpn_decode_kernel<<<(n + 255) / 256, 256, 0, st>>>(reg, base_anchors, reinterpret_cast<const long long*>(top_idx), n, P, boxes, valid);
  DA_LAUNCH_CHECK();
  nms_mask_kernel<<<dim3((unsigned)words, (unsigned)words), 64, 0, st>>>(boxes, n, iou_thr, words, mask);
  DA_LAUNCH_CHECK();
  const size_t smem = (size_t)words * sizeof(unsigned long long);
  if (smem > 48 * 1024) DA_CUDA_OK(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_scan_kernel<<<1, SCAN_THREADS, smem, st>>>(mask, valid, boxes, top_scores, n, words, max_out, dets, count, keep_idx);
  DA_LAUNCH_CHECK();
  return DA_OK;
}

// decoded boxes + validity of the last da_rpn_proposals call on this workspace (tests / callers that want the pre-NMS set)
extern "C" int da_rpn_proposals_peek(const void* workspace, int n, float* boxes_out, unsigned char* valid_out, da_stream_t stream) {
  DA_REQUIRE(workspace && n > 0 && boxes_out, DA_ERR_INVALID_ARG, "rpn_proposals_peek: bad args");
  size_t off_valid, off_mask;
  rpn_ws_layout(n, &off_valid, &off_mask);
  const unsigned char* ws = static_cast<const unsigned char*>(workspace);
  DA_CUDA_OK(cudaMemcpyAsync(boxes_out, ws, (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  if (valid_out) DA_CUDA_OK(cudaMemcpyAsync(valid_out, ws + off_valid, (size_t)n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return DA_OK;
}
