// Domain-loss tails of the DA hot path and the 1-channel
// head tail.  All are HBM/launch-latency bound reductions: one pass over the logits,
// warp-shuffle + shared-memory tree, deterministic (no floating-point atomics).
#include "da_common.cuh"
#include <float.h>

namespace da {

// ---------------------------------------------------------------------------------------
// L1 / L2 pixel loss  (resnet_da_daf_org.py:816-822, resnet_da_cbam.py:971-979)
// ---------------------------------------------------------------------------------------
constexpr int PL_THREADS = 256;
constexpr int PL_MAX_BLOCKS = 64;  // per image

__host__ __device__ inline int pl_blocks(int64_t L) {
  int64_t b = (L + PL_THREADS * 4 - 1) / (PL_THREADS * 4);
  return (int)(b < 1 ? 1 : (b > PL_MAX_BLOCKS ? PL_MAX_BLOCKS : b));
}

// partial[(n*nb + blk)*2 + {0,1}] = sum sigmoid(p)^2, sum sigmoid(1-p)^2 over the block's slice
__global__ void pixel_loss_partial_kernel(const float* __restrict__ logits, int64_t L, int nb,
                                          float* __restrict__ partial) {
  __shared__ float red[33];
  const int n = blockIdx.y, blk = blockIdx.x;
  const float* p = logits + (size_t)n * L;
  float s0 = 0.f, s1 = 0.f;
  for (int64_t i = (int64_t)blk * PL_THREADS + threadIdx.x; i < L; i += (int64_t)nb * PL_THREADS) {
    const float v = p[i];
    const float a = sigmoidf_(v), b = sigmoidf_(1.f - v);
    s0 = fmaf(a, a, s0);
    s1 = fmaf(b, b, s1);
  }
  s0 = block_sum<false>(s0, red);
  s1 = block_sum<false>(s1, red);
  if (threadIdx.x == 0) {
    partial[((size_t)n * nb + blk) * 2 + 0] = s0;
    partial[((size_t)n * nb + blk) * 2 + 1] = s1;
  }
}

__global__ void pixel_loss_final_kernel(const float* __restrict__ partial, int N, int nb, int64_t L,
                                        const int32_t* __restrict__ domain, int whole_batch,
                                        float* __restrict__ loss_out) {
  // one warp; N and nb are small
  const int lane = threadIdx.x;
  float total = 0.f;
  if (whole_batch) {
    float s0 = 0.f, s1 = 0.f;
    for (int i = lane; i < N * nb; i += 32) { s0 += partial[2 * i]; s1 += partial[2 * i + 1]; }
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    int n_src = 0, n_tgt = 0;
    for (int i = 0; i < N; ++i) { n_src += (domain[i] == 0); n_tgt += (domain[i] == 1); }
    const float denom = (float)((double)N * (double)L);
    total = 0.5f * ((float)n_src * (s0 / denom) + (float)n_tgt * (s1 / denom));
  } else {
    for (int n = 0; n < N; ++n) {
      const int d = domain[n];
      float s = 0.f;
      for (int i = lane; i < nb; i += 32) s += partial[((size_t)n * nb + i) * 2 + (d == 1 ? 1 : 0)];
      s = warp_sum(s);
      if (d == 0 || d == 1) total += 0.5f * (s / (float)L);
    }
  }
  if (lane == 0) loss_out[0] = total;
}

__global__ void pixel_loss_bwd_kernel(const float* __restrict__ logits, int N, int64_t L,
                                      const int32_t* __restrict__ domain, int whole_batch,
                                      const float* __restrict__ grad_loss, float scale,
                                      float* __restrict__ dlogits) {
  const int n = blockIdx.y;
  const float g = (grad_loss ? grad_loss[0] : 1.f) * scale;
  float ca, cb;  // coefficients of sigma(p)^2(1-sigma(p)) and -sigma(1-p)^2(1-sigma(1-p))
  if (whole_batch) {
    int n_src = 0, n_tgt = 0;
    for (int i = 0; i < N; ++i) { n_src += (domain[i] == 0); n_tgt += (domain[i] == 1); }
    const float denom = (float)((double)N * (double)L);
    ca = g * (float)n_src / denom;
    cb = g * (float)n_tgt / denom;
  } else {
    const int d = domain[n];
    ca = (d == 0) ? g / (float)L : 0.f;
    cb = (d == 1) ? g / (float)L : 0.f;
  }
  const float* p = logits + (size_t)n * L;
  float* o = dlogits + (size_t)n * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = p[i];
    const float a = sigmoidf_(v), b = sigmoidf_(1.f - v);
    o[i] = ca * a * a * (1.f - a) - cb * b * b * (1.f - b);
  }
}

// ---------------------------------------------------------------------------------------
// L3 / L4 two-class cross entropy (optionally on sigmoid outputs, Q4)
// ---------------------------------------------------------------------------------------
__global__ void ce2_fwd_kernel(const float* __restrict__ z, const int32_t* __restrict__ labels, int R,
                               int on_sigmoid, float* __restrict__ pred_out, float* __restrict__ loss_out) {
  __shared__ float red[33];
  float acc = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    float u0 = z[2 * r], u1 = z[2 * r + 1];
    if (on_sigmoid) { u0 = sigmoidf_(u0); u1 = sigmoidf_(u1); }
    if (pred_out) { pred_out[2 * r] = u0; pred_out[2 * r + 1] = u1; }
    const int l = labels[r];
    if (l == 0 || l == 1) {
      const float mx = fmaxf(u0, u1);
      const float lse = mx + logf(expf(u0 - mx) + expf(u1 - mx));
      acc += lse - (l ? u1 : u0);
    }
  }
  acc = block_sum<false>(acc, red);
  if (threadIdx.x == 0) loss_out[0] = acc / (float)R;
}

__global__ void ce2_bwd_kernel(const float* __restrict__ z, const int32_t* __restrict__ labels, int R,
                               int on_sigmoid, const float* __restrict__ grad_loss, float scale,
                               const float* __restrict__ grad_pred, float* __restrict__ dz) {
  pdl_launch_dependents();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float g = (grad_loss ? grad_loss[0] : 1.f) * scale / (float)R;
  float u0 = z[2 * r], u1 = z[2 * r + 1];
  if (on_sigmoid) { u0 = sigmoidf_(u0); u1 = sigmoidf_(u1); }
  const int l = labels[r];
  float d0 = 0.f, d1 = 0.f;
  if (l == 0 || l == 1) {
    const float mx = fmaxf(u0, u1);
    const float e0 = expf(u0 - mx), e1 = expf(u1 - mx);
    const float inv = 1.f / (e0 + e1);
    d0 = g * (e0 * inv - (l == 0 ? 1.f : 0.f));
    d1 = g * (e1 * inv - (l == 1 ? 1.f : 0.f));
  }
  if (grad_pred) { d0 += grad_pred[2 * r]; d1 += grad_pred[2 * r + 1]; }
  if (on_sigmoid) { d0 *= u0 * (1.f - u0); d1 *= u1 * (1.f - u1); }
  dz[2 * r] = d0;
  dz[2 * r + 1] = d1;
}

// ---------------------------------------------------------------------------------------
// L6 sigmoid focal loss (mmcv sigmoid_focal_loss semantics, FLT_MIN clamp)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float focal_elem(float u, bool pos, float gamma, float alpha) {
  const float p = sigmoidf_(u);
  if (pos) return -alpha * powf(1.f - p, gamma) * logf(fmaxf(p, FLT_MIN));
  return -(1.f - alpha) * powf(p, gamma) * logf(fmaxf(1.f - p, FLT_MIN));
}
__device__ __forceinline__ float focal_grad(float u, bool pos, float gamma, float alpha) {
  const float p = sigmoidf_(u);
  if (pos) return -alpha * powf(1.f - p, gamma) * (1.f - p - gamma * p * logf(fmaxf(p, FLT_MIN)));
  return (1.f - alpha) * powf(p, gamma) * (p - gamma * (1.f - p) * logf(fmaxf(1.f - p, FLT_MIN)));
}

__global__ void focal2_fwd_kernel(const float* __restrict__ u, const int32_t* __restrict__ labels, int k,
                                  float gamma, float alpha, float* __restrict__ loss_out) {
  __shared__ float red[33];
  float acc = 0.f;
  for (int i = threadIdx.x; i < 2 * k; i += blockDim.x) {
    const int r = i >> 1, c = i & 1;
    acc += focal_elem(u[i], labels[r] == c, gamma, alpha);
  }
  acc = block_sum<false>(acc, red);
  if (threadIdx.x == 0) loss_out[0] = acc / (float)(2 * k);
}

__global__ void focal2_bwd_kernel(const float* __restrict__ u, const int32_t* __restrict__ labels, int k,
                                  float gamma, float alpha, const float* __restrict__ grad_loss,
                                  float scale, float* __restrict__ du) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * k) return;
  const float g = (grad_loss ? grad_loss[0] : 1.f) * scale / (float)(2 * k);
  du[i] = g * focal_grad(u[i], labels[i >> 1] == (i & 1), gamma, alpha);
}

// ---------------------------------------------------------------------------------------
// L7 consistency regulariser (DAFaster_rcnn_Orig.py:161-175, closed form)
// ---------------------------------------------------------------------------------------
__global__ void consistency_fwd_kernel(const float* __restrict__ img_logits, int64_t n_img,
                                       const float* __restrict__ ins_pred, const int32_t* __restrict__ labels,
                                       int R, float* __restrict__ mean_out, float* __restrict__ loss_out) {
  __shared__ float red[33];
  float s = 0.f;
  if ((n_img & 3) == 0 && (reinterpret_cast<uintptr_t>(img_logits) & 15) == 0) {
    // one block on purpose (deterministic, no second launch); 16-byte loads, four in flight per thread: the scalar loop was
    // a chain of L2 round trips (12 us for 16 K logits)
    const float4* p = reinterpret_cast<const float4*>(img_logits);
    const int64_t n4 = n_img >> 2;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = p[i];
      s0 += sigmoidf_(v.x); s1 += sigmoidf_(v.y); s2 += sigmoidf_(v.z); s3 += sigmoidf_(v.w);
    }
    s = (s0 + s1) + (s2 + s3);
  } else {
    for (int64_t i = threadIdx.x; i < n_img; i += blockDim.x) s += sigmoidf_(img_logits[i]);
  }
  const float m = block_sum<true>(s, red) / (float)n_img;
  float acc = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const int l = labels[r];
    if (l == 0 || l == 1) acc += fabsf(m - sigmoidf_(ins_pred[2 * r + l]));
  }
  acc = block_sum<false>(acc, red);
  if (threadIdx.x == 0) { mean_out[0] = m; loss_out[0] = acc; }
}

__global__ void consistency_bwd_kernel(const float* __restrict__ img_logits, int64_t n_img,
                                       const float* __restrict__ ins_pred, const int32_t* __restrict__ labels,
                                       int R, const float* __restrict__ mean_in,
                                       const float* __restrict__ grad_loss, float scale,
                                       float* __restrict__ d_img, float* __restrict__ d_pred) {
  pdl_launch_dependents();
  __shared__ float red[33];
  const float g = (grad_loss ? grad_loss[0] : 1.f) * scale;
  const float m = mean_in[0];
  float sgn_sum = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const int l = labels[r];
    float d0 = 0.f, d1 = 0.f;
    if (l == 0 || l == 1) {
      const float s = sigmoidf_(ins_pred[2 * r + l]);
      const float diff = m - s;
      const float sg = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
      sgn_sum += sg;
      const float d = -g * sg * s * (1.f - s);
      if (l == 0) d0 = d; else d1 = d;
    }
    if (d_pred) { d_pred[2 * r] = d0; d_pred[2 * r + 1] = d1; }
  }
  const float S = block_sum<true>(sgn_sum, red);
  if (d_img) {
    const float c = g * S / (float)n_img;
    if ((n_img & 3) == 0 && ((reinterpret_cast<uintptr_t>(img_logits) | reinterpret_cast<uintptr_t>(d_img)) & 15) == 0) {
      const float4* p = reinterpret_cast<const float4*>(img_logits);
      float4* q = reinterpret_cast<float4*>(d_img);
      const int64_t n4 = n_img >> 2;
#pragma unroll 4
      for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
        const float4 v = p[i];
        const float a0 = sigmoidf_(v.x), a1 = sigmoidf_(v.y), a2 = sigmoidf_(v.z), a3 = sigmoidf_(v.w);
        q[i] = make_float4(c * a0 * (1.f - a0), c * a1 * (1.f - a1), c * a2 * (1.f - a2), c * a3 * (1.f - a3));
      }
    } else {
      for (int64_t i = threadIdx.x; i < n_img; i += blockDim.x) {
        const float a = sigmoidf_(img_logits[i]);
        d_img[i] = c * a * (1.f - a);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// 1-channel head tail: logits[m] = act(x[m,:] . w + bias)
// ---------------------------------------------------------------------------------------
constexpr int PH_WARPS = 8;

template <typename T>
__device__ __forceinline__ float dot_row(const T* __restrict__ x, const float* __restrict__ w_s, int K, int lane);
template <>
__device__ __forceinline__ float dot_row<float>(const float* __restrict__ x, const float* __restrict__ w_s, int K, int lane) {
  float acc = 0.f;
  if ((K & 3) == 0) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* w4 = reinterpret_cast<const float4*>(w_s);
    for (int i = lane; i < (K >> 2); i += 32) {
      const float4 a = __ldg(x4 + i), b = w4[i];
      acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    }
  } else {
    for (int i = lane; i < K; i += 32) acc = fmaf(x[i], w_s[i], acc);
  }
  return warp_sum(acc);
}
template <>
__device__ __forceinline__ float dot_row<__nv_bfloat16>(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w_s, int K, int lane) {
  float acc = 0.f;
  if ((K & 7) == 0) {
    const uint4* x8 = reinterpret_cast<const uint4*>(x);
    for (int i = lane; i < (K >> 3); i += 32) {
      const uint4 a = __ldg(x8 + i);
      const float* w = w_s + i * 8;
      const unsigned v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc = fmaf(__uint_as_float(v[j] << 16), w[2 * j], acc);
        acc = fmaf(__uint_as_float(v[j] & 0xffff0000u), w[2 * j + 1], acc);
      }
    }
  } else {
    for (int i = lane; i < K; i += 32) acc = fmaf(__bfloat162float(x[i]), w_s[i], acc);
  }
  return warp_sum(acc);
}

template <typename T>
__global__ void pixel_head_fwd_kernel(const T* __restrict__ x, int64_t M, int K, const float* __restrict__ w,
                                      const float* __restrict__ bias, int relu, float* __restrict__ logits) {
  extern __shared__ __align__(16) float w_s[];
  for (int i = threadIdx.x; i < K; i += blockDim.x) w_s[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float bv = bias ? bias[0] : 0.f;
  for (int64_t m = (int64_t)blockIdx.x * PH_WARPS + wid; m < M; m += (int64_t)gridDim.x * PH_WARPS) {
    float v = dot_row<T>(x + (size_t)m * K, w_s, K, lane) + bv;
    if (relu) v = fmaxf(v, 0.f);
    if (lane == 0) logits[m] = v;
  }
}

// dx[m,k] = dl[m]*w[k]; partial dw over the block's rows -> workspace[blk][K]; dbias partial at [nb*K + blk]
template <typename T, typename TDx>
__global__ void pixel_head_bwd_kernel(const T* __restrict__ x, int64_t M, int K, const float* __restrict__ w,
                                      const float* __restrict__ dlogits, const float* __restrict__ post_relu,
                                      TDx* __restrict__ dx, float* __restrict__ partial, int rows_per_block) {
  const int64_t m0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t m1 = (m0 + rows_per_block < M) ? m0 + rows_per_block : M;
  float db = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float wk = w[k];
    float acc = 0.f;
    for (int64_t m = m0; m < m1; ++m) {
      float dl = dlogits[m];
      if (post_relu && !(post_relu[m] > 0.f)) dl = 0.f;
      acc = fmaf(dl, to_f32<T>(x[(size_t)m * K + k]), acc);
      if (dx) dx[(size_t)m * K + k] = from_f32<TDx>(dl * wk);
    }
    partial[(size_t)blockIdx.x * K + k] = acc;
  }
  if (threadIdx.x == 0) {
    for (int64_t m = m0; m < m1; ++m) {
      float dl = dlogits[m];
      if (post_relu && !(post_relu[m] > 0.f)) dl = 0.f;
      db += dl;
    }
    partial[(size_t)gridDim.x * K + blockIdx.x] = db;
  }
}

// bf16 x and dx, K in {64 .. 2048} with 256 % (K/8) == 0: 16-byte accesses, a thread owns 8 consecutive channels and every
// SUB-th row; partial dw sums meet in shared memory (the scalar kernel: 25 us for [16384, 512]).
__global__ void __launch_bounds__(256)
pixel_head_bwd_vec8_kernel(const __nv_bfloat16* __restrict__ x, int64_t M, int K, const float* __restrict__ w,
                           const float* __restrict__ dlogits, const float* __restrict__ post_relu,
                           __nv_bfloat16* __restrict__ dx, float* __restrict__ partial, int rows_per_block) {
  extern __shared__ float ph_red[];   // [SUB][K] when SUB > 1
  const int64_t m0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t m1 = (m0 + rows_per_block < M) ? m0 + rows_per_block : M;
  const int G = K >> 3, SUB = 256 / G;
  const int cg = threadIdx.x % G, sub = threadIdx.x / G, k0 = cg << 3;
  float wk[8], acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { wk[j] = w[k0 + j]; acc[j] = 0.f; }
  for (int64_t m = m0 + sub; m < m1; m += SUB) {
    float dl = dlogits[m];
    if (post_relu && !(post_relu[m] > 0.f)) dl = 0.f;
    const uint4 xv = *reinterpret_cast<const uint4*>(x + (size_t)m * K + k0);
    const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(&xv);
    uint4 ov;
    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(&ov);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j] = fmaf(dl, __bfloat162float(xp[j]), acc[j]);
      op[j] = __float2bfloat16_rn(dl * wk[j]);
    }
    if (dx) *reinterpret_cast<uint4*>(dx + (size_t)m * K + k0) = ov;
  }
  if (SUB > 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) ph_red[(size_t)sub * K + k0 + j] = acc[j];
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += 256) {
      float a = 0.f;
      for (int q = 0; q < SUB; ++q) a += ph_red[(size_t)q * K + k];
      partial[(size_t)blockIdx.x * K + k] = a;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) partial[(size_t)blockIdx.x * K + k0 + j] = acc[j];
  }
  if (threadIdx.x == 0) {
    float db = 0.f;
    for (int64_t m = m0; m < m1; ++m) {
      float dl = dlogits[m];
      if (post_relu && !(post_relu[m] > 0.f)) dl = 0.f;
      db += dl;
    }
    partial[(size_t)gridDim.x * K + blockIdx.x] = db;
  }
}

__global__ void pixel_head_bwd_final_kernel(const float* __restrict__ partial, int nb, int K,
                                            float* __restrict__ dw, float* __restrict__ dbias) {
  // block (32 columns x 32 row lanes); the last block column (k == K) reduces the bias partials
  __shared__ float red[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int k = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (k < K) {
    for (int b = ty; b < nb; b += 32) acc += partial[(size_t)b * K + k];
  } else if (k == K) {
    for (int b = ty; b < nb; b += 32) acc += partial[(size_t)nb * K + b];
  }
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) t += red[j][tx];
    if (k < K && dw) dw[k] = t;
    if (k == K && dbias) dbias[0] = t;
  }
}

// ---------------------------------------------------------------------------------------
// global average pool over pixels: x [N,HW,C] -> y [N,C]
// ---------------------------------------------------------------------------------------
constexpr int AP_SPLIT = 32;

template <typename T>
__global__ void avgpool_partial_kernel(const T* __restrict__ x, int HW, int C, float* __restrict__ partial) {
  // grid (ceil(C/128), AP_SPLIT, N), block 128: thread = channel, loops its pixel slice
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = blockIdx.y, n = blockIdx.z;
  if (c >= C) return;
  const int per = (HW + AP_SPLIT - 1) / AP_SPLIT;
  const int p0 = s * per, p1 = min(p0 + per, HW);
  const T* px = x + ((size_t)n * HW) * C + c;
  float acc = 0.f;
#pragma unroll 4
  for (int p = p0; p < p1; ++p) acc += to_f32<T>(px[(size_t)p * C]);
  partial[((size_t)n * AP_SPLIT + s) * C + c] = acc;
}
// bf16, C % 8 == 0: thread = 8 consecutive channels, 16-byte loads (the scalar kernel reads 2 bytes per thread and pixel:
// 90 us for the [2, 70x134, 4608] SRM activation of C5, 173 MB)
__global__ void avgpool_partial_vec8_kernel(const __nv_bfloat16* __restrict__ x, int HW, int C, float* __restrict__ partial) {
  const int cg = blockIdx.x * blockDim.x + threadIdx.x;   // group of 8 channels
  const int s = blockIdx.y, n = blockIdx.z;
  if (cg * 8 >= C) return;
  const int per = (HW + AP_SPLIT - 1) / AP_SPLIT;
  const int p0 = s * per, p1 = min(p0 + per, HW);
  const __nv_bfloat16* px = x + ((size_t)n * HW) * C + (size_t)cg * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll 4
  for (int p = p0; p < p1; ++p) {
    const uint4 v = *reinterpret_cast<const uint4*>(px + (size_t)p * C);
    const __nv_bfloat16* vp = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += __bfloat162float(vp[j]);
  }
  float* o = partial + ((size_t)n * AP_SPLIT + s) * C + (size_t)cg * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = acc[j];
}
__global__ void avgpool_final_kernel(const float* __restrict__ partial, int HW, int C, int N, float* __restrict__ y) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (c >= C) return;
  float acc = 0.f;
  for (int s = 0; s < AP_SPLIT; ++s) acc += partial[((size_t)n * AP_SPLIT + s) * C + c];
  y[(size_t)n * C + c] = acc / (float)HW;
}
template <typename T>
__global__ void avgpool_bwd_kernel(const float* __restrict__ dy, int HW, int C, T* __restrict__ dx, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t n = i / ((int64_t)HW * C);
    dx[i] = from_f32<T>(dy[n * C + c] / (float)HW);
  }
}

// bf16, C % 8 == 0: one 16-byte store per thread and 8 channels (the scalar kernel: 200-300 us per SRM head)
__global__ void avgpool_bwd_vec8_kernel(const float* __restrict__ dy, int HW, int C, __nv_bfloat16* __restrict__ dx, int64_t total8) {
  for (int64_t i8 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i8 < total8; i8 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i8 * 8;
    const int c = (int)(i % C);
    const int64_t n = i / ((int64_t)HW * C);
    const float* d = dy + n * C + c;
    uint4 ov;
    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(&ov);
#pragma unroll
    for (int j = 0; j < 8; ++j) op[j] = __float2bfloat16_rn(d[j] / (float)HW);
    *reinterpret_cast<uint4*>(dx + i) = ov;
  }
}

// ---------------------------------------------------------------------------------------
// softmax over the query axis (dim 0 of s[q,k]); block (32,32): 32 columns x 32 row lanes
// ---------------------------------------------------------------------------------------
__global__ void softmax_dim0_fwd_kernel(const float* __restrict__ s, int T, int ld, float* __restrict__ p) {
  pdl_launch_dependents();
  __shared__ float red[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int k = blockIdx.x * 32 + tx;
  float mx = -INFINITY;
  if (k < T) for (int q = ty; q < T; q += 32) mx = fmaxf(mx, s[(size_t)q * ld + k]);
  red[ty][tx] = mx;
  __syncthreads();
  if (ty == 0) { float m = red[0][tx]; for (int j = 1; j < 32; ++j) m = fmaxf(m, red[j][tx]); red[0][tx] = m; }
  __syncthreads();
  mx = red[0][tx];
  __syncthreads();
  float sum = 0.f;
  if (k < T) for (int q = ty; q < T; q += 32) sum += expf(s[(size_t)q * ld + k] - mx);
  red[ty][tx] = sum;
  __syncthreads();
  if (ty == 0) { float m = 0.f; for (int j = 0; j < 32; ++j) m += red[j][tx]; red[0][tx] = m; }
  __syncthreads();
  const float inv = 1.f / red[0][tx];
  if (k < T) for (int q = ty; q < T; q += 32) p[(size_t)q * ld + k] = expf(s[(size_t)q * ld + k] - mx) * inv;
}
__global__ void softmax_dim0_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp, int T, int ld,
                                        float* __restrict__ ds) {
  pdl_launch_dependents();
  __shared__ float red[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int k = blockIdx.x * 32 + tx;
  float dot = 0.f;
  if (k < T) for (int q = ty; q < T; q += 32) dot = fmaf(dp[(size_t)q * ld + k], p[(size_t)q * ld + k], dot);
  red[ty][tx] = dot;
  __syncthreads();
  if (ty == 0) { float m = 0.f; for (int j = 0; j < 32; ++j) m += red[j][tx]; red[0][tx] = m; }
  __syncthreads();
  dot = red[0][tx];
  if (k < T) for (int q = ty; q < T; q += 32) {
    const size_t i = (size_t)q * ld + k;
    ds[i] = p[i] * (dp[i] - dot);
  }
}

}  // namespace da

using namespace da;

// ---------------------------------------------------------------------------------------
// W1: lambda-weighting of the DA losses and their total (DAFaster_rcnn_Orig.py:143-157, base.py:176-219) in one launch
// ---------------------------------------------------------------------------------------
struct ScalarPtrs { const float* p[DA_MAX_WEIGHTED]; float w[DA_MAX_WEIGHTED]; };
__global__ void weighted_sum_fwd_kernel(ScalarPtrs a, int n, float* __restrict__ scaled, float* __restrict__ total) {
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < n; ++i) { const float v = a.w[i] * a.p[i][0]; scaled[i] = v; t += v; }    // fixed order
    total[0] = t;
  }
}
// d loss_i = w_i * (g_total + g_scaled[i])
__global__ void weighted_sum_bwd_kernel(ScalarPtrs a, int n, const float* __restrict__ g_total, const float* __restrict__ g_scaled,
                                        float* __restrict__ d_in) {
  pdl_launch_dependents();
  const int i = threadIdx.x;
  if (i < n) d_in[i] = a.w[i] * ((g_total ? g_total[0] : 0.f) + (g_scaled ? g_scaled[i] : 0.f));
}

extern "C" size_t da_pixel_loss_workspace_bytes(int N, int64_t L) {
  return (size_t)(N > 0 ? N : 0) * pl_blocks(L) * 2 * sizeof(float) + 64;
}

extern "C" int da_pixel_domain_loss_forward(const float* logits, int N, int64_t L, const int32_t* domain,
                                            int whole_batch, float* loss_out, void* workspace,
                                            size_t workspace_bytes, da_stream_t stream) {
  DA_REQUIRE(N > 0 && L > 0 && logits && domain && loss_out, DA_ERR_INVALID_ARG, "pixel_domain_loss_forward: bad args");
  DA_REQUIRE(workspace && workspace_bytes >= da_pixel_loss_workspace_bytes(N, L), DA_ERR_WORKSPACE, "pixel_domain_loss_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = pl_blocks(L);
  pixel_loss_partial_kernel<<<dim3(nb, N), PL_THREADS, 0, st>>>(logits, L, nb, (float*)workspace);
  DA_LAUNCH_CHECK();
  pixel_loss_final_kernel<<<1, 32, 0, st>>>((const float*)workspace, N, nb, L, domain, whole_batch, loss_out);
  DA_LAUNCH_CHECK();
  return DA_OK;
}

extern "C" int da_pixel_domain_loss_backward(const float* logits, int N, int64_t L, const int32_t* domain,
                                             int whole_batch, const float* grad_loss, float scale,
                                             float* dlogits, da_stream_t stream) {
  DA_REQUIRE(N > 0 && L > 0 && logits && domain && dlogits, DA_ERR_INVALID_ARG, "pixel_domain_loss_backward: bad args");
  const int nb = pl_blocks(L);
  pixel_loss_bwd_kernel<<<dim3(nb, N), PL_THREADS, 0, (cudaStream_t)stream>>>(logits, N, L, domain, whole_batch, grad_loss, scale, dlogits);
  DA_LAUNCH_CHECK();
  return DA_OK;
}

extern "C" int da_ce2_forward(const float* z, const int32_t* labels, int R, int on_sigmoid,
                              float* pred_out, float* loss_out, da_stream_t stream) {
  DA_REQUIRE(R > 0 && z && labels && loss_out, DA_ERR_INVALID_ARG, "ce2_forward: bad args (R=%d)", R);
  ce2_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(z, labels, R, on_sigmoid, pred_out, loss_out);
  DA_LAUNCH_CHECK();
  return DA_OK;
}
extern "C" int da_ce2_backward(const float* z, const int32_t* labels, int R, int on_sigmoid,
                               const float* grad_loss, float scale, const float* grad_pred,
                               float* dz, da_stream_t stream) {
  DA_REQUIRE(R > 0 && z && labels && dz, DA_ERR_INVALID_ARG, "ce2_backward: bad args (R=%d)", R);
  ce2_bwd_kernel<<<(R + 255) / 256, 256, 0, (cudaStream_t)stream>>>(z, labels, R, on_sigmoid, grad_loss, scale, grad_pred, dz);
  DA_LAUNCH_CHECK();
  return DA_OK;
}

extern "C" int da_focal2_forward(const float* u, const int32_t* labels, int k, float gamma, float alpha,
                                 float* loss_out, da_stream_t stream) {
  DA_REQUIRE(k > 0 && u && labels && loss_out, DA_ERR_INVALID_ARG, "focal2_forward: bad args (k=%d)", k);
  focal2_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(u, labels, k, gamma, alpha, loss_out);
  DA_LAUNCH_CHECK();
  return DA_OK;
}
extern "C" int da_focal2_backward(const float* u, const int32_t* labels, int k, float gamma, float alpha,
                                  const float* grad_loss, float scale, float* du, da_stream_t stream) {
  DA_REQUIRE(k > 0 && u && labels && du, DA_ERR_INVALID_ARG, "focal2_backward: bad args (k=%d)", k);
  focal2_bwd_kernel<<<(2 * k + 255) / 256, 256, 0, (cudaStream_t)stream>>>(u, labels, k, gamma, alpha, grad_loss, scale, du);
  DA_LAUNCH_CHECK();
  return DA_OK;
}

extern "C" int da_consistency_forward(const float* img_logits, int64_t n_img, const float* ins_pred,
                                      const int32_t* labels, int R, float* mean_out, float* loss_out,
                                      da_stream_t stream) {
  DA_REQUIRE(n_img > 0 && R >= 0 && img_logits && mean_out && loss_out, DA_ERR_INVALID_ARG, "consistency_forward: bad args");
  DA_REQUIRE(R == 0 || (ins_pred && labels), DA_ERR_INVALID_ARG, "consistency_forward: null instance inputs");
  consistency_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(img_logits, n_img, ins_pred, labels, R, mean_out, loss_out);
  DA_LAUNCH_CHECK();
  return DA_OK;
}
extern "C" int da_consistency_backward(const float* img_logits, int64_t n_img, const float* ins_pred,
                                       const int32_t* labels, int R, const float* mean_in,
                                       const float* grad_loss, float scale,
                                       float* d_img_logits, float* d_ins_pred, da_stream_t stream) {
  DA_REQUIRE(n_img > 0 && R >= 0 && img_logits && mean_in, DA_ERR_INVALID_ARG, "consistency_backward: bad args");
  consistency_bwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(img_logits, n_img, ins_pred, labels, R, mean_in, grad_loss, scale, d_img_logits, d_ins_pred);
  DA_LAUNCH_CHECK();
  return DA_OK;
}

extern "C" int da_pixel_head_forward(const void* x, int x_dtype, int64_t M, int K, const float* w,
                                     const float* bias, int relu, float* logits, da_stream_t stream) {
  DA_REQUIRE(M > 0 && K > 0 && x && w && logits, DA_ERR_INVALID_ARG, "pixel_head_forward: bad args");
  DA_REQUIRE((size_t)K * 4 <= 200 * 1024, DA_ERR_UNSUPPORTED, "pixel_head_forward: K=%d too large", K);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = (M + PH_WARPS - 1) / PH_WARPS;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  const size_t smem = (size_t)K * sizeof(float);
  if (x_dtype == DA_F32) {
    if (smem > 48 * 1024) DA_CUDA_OK(cudaFuncSetAttribute(pixel_head_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pixel_head_fwd_kernel<float><<<(int)blocks, PH_WARPS * 32, smem, st>>>((const float*)x, M, K, w, bias, relu, logits);
  } else if (x_dtype == DA_BF16) {
    if (smem > 48 * 1024) DA_CUDA_OK(cudaFuncSetAttribute(pixel_head_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pixel_head_fwd_kernel<__nv_bfloat16><<<(int)blocks, PH_WARPS * 32, smem, st>>>((const __nv_bfloat16*)x, M, K, w, bias, relu, logits);
  } else {
    DA_REQUIRE(false, DA_ERR_INVALID_ARG, "pixel_head_forward: bad dtype %d", x_dtype);
  }
  DA_LAUNCH_CHECK();
  return DA_OK;
}

static inline int ph_rows_per_block(int64_t M) {
  int64_t target_blocks = (int64_t)num_sms() * 4;
  int64_t rpb = (M + target_blocks - 1) / target_blocks;
  if (rpb < 8) rpb = 8;
  return (int)rpb;
}
extern "C" size_t da_pixel_head_workspace_bytes(int64_t M, int K) {
  if (M <= 0 || K <= 0) return 64;
  const int rpb = ph_rows_per_block(M);
  const int64_t nb = (M + rpb - 1) / rpb;
  return (size_t)nb * (K + 1) * sizeof(float) + 64;
}
extern "C" int da_pixel_head_backward(const void* x, int x_dtype, int64_t M, int K, const float* w,
                                      const float* dlogits, const float* post_relu_logits,
                                      void* dx, int dx_dtype, float* dw, float* dbias,
                                      void* workspace, size_t workspace_bytes, da_stream_t stream) {
  DA_REQUIRE(M > 0 && K > 0 && x && w && dlogits, DA_ERR_INVALID_ARG, "pixel_head_backward: bad args");
  DA_REQUIRE(workspace && workspace_bytes >= da_pixel_head_workspace_bytes(M, K), DA_ERR_WORKSPACE, "pixel_head_backward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int rpb = ph_rows_per_block(M);
  const int nb = (int)((M + rpb - 1) / rpb);
  float* partial = (float*)workspace;
#define DA_PH_BWD(T, TDX)                                                                           \
  pixel_head_bwd_kernel<T, TDX><<<nb, 256, 0, st>>>((const T*)x, M, K, w, dlogits, post_relu_logits, \
                                                    (TDX*)dx, partial, rpb)
  if (x_dtype == DA_F32 && dx_dtype == DA_F32) DA_PH_BWD(float, float);
  else if (x_dtype == DA_F32 && dx_dtype == DA_BF16) DA_PH_BWD(float, __nv_bfloat16);
  else if (x_dtype == DA_BF16 && dx_dtype == DA_F32) DA_PH_BWD(__nv_bfloat16, float);
  else if (x_dtype == DA_BF16 && dx_dtype == DA_BF16 && (K & 7) == 0 && (K >> 3) <= 256 && 256 % (K >> 3) == 0 &&
           ((((uintptr_t)x) | ((uintptr_t)dx)) & 15) == 0) {
    const int sub = 256 / (K >> 3);
    pixel_head_bwd_vec8_kernel<<<nb, 256, sub > 1 ? (size_t)sub * K * sizeof(float) : 0, st>>>(
        (const __nv_bfloat16*)x, M, K, w, dlogits, post_relu_logits, (__nv_bfloat16*)dx, partial, rpb);
  } else if (x_dtype == DA_BF16 && dx_dtype == DA_BF16) DA_PH_BWD(__nv_bfloat16, __nv_bfloat16);
  else DA_REQUIRE(false, DA_ERR_INVALID_ARG, "pixel_head_backward: bad dtype");
#undef DA_PH_BWD
  DA_LAUNCH_CHECK();
  pixel_head_bwd_final_kernel<<<(K + 1 + 31) / 32, dim3(32, 32), 0, st>>>(partial, nb, K, dw, dbias);
  DA_LAUNCH_CHECK();
  return DA_OK;
}

extern "C" size_t da_global_avgpool_workspace_bytes(int N, int C) {
  return (size_t)(N > 0 ? N : 0) * AP_SPLIT * (C > 0 ? C : 0) * sizeof(float) + 64;
}
extern "C" int da_global_avgpool_forward(const void* x, int x_dtype, int N, int HW, int C, float* y,
                                         void* workspace, size_t workspace_bytes, da_stream_t stream) {
  DA_REQUIRE(N > 0 && HW > 0 && C > 0 && x && y, DA_ERR_INVALID_ARG, "global_avgpool_forward: bad args");
  DA_REQUIRE(workspace && workspace_bytes >= da_global_avgpool_workspace_bytes(N, C), DA_ERR_WORKSPACE, "global_avgpool_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((C + 127) / 128, AP_SPLIT, N);
  if (x_dtype == DA_F32) avgpool_partial_kernel<float><<<grid, 128, 0, st>>>((const float*)x, HW, C, (float*)workspace);
  else if (x_dtype == DA_BF16 && (C & 7) == 0 && (((uintptr_t)x) & 15) == 0)
    avgpool_partial_vec8_kernel<<<dim3((C / 8 + 127) / 128, AP_SPLIT, N), 128, 0, st>>>((const __nv_bfloat16*)x, HW, C, (float*)workspace);
  else if (x_dtype == DA_BF16) avgpool_partial_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((const __nv_bfloat16*)x, HW, C, (float*)workspace);
  else DA_REQUIRE(false, DA_ERR_INVALID_ARG, "global_avgpool_forward: bad dtype");
  DA_LAUNCH_CHECK();
  avgpool_final_kernel<<<dim3((C + 127) / 128, N), 128, 0, st>>>((const float*)workspace, HW, C, N, y);
  DA_LAUNCH_CHECK();
  return DA_OK;
}
extern "C" int da_global_avgpool_backward(const float* dy, int N, int HW, int C, void* dx, int dx_dtype,
                                          da_stream_t stream) {
  DA_REQUIRE(N > 0 && HW > 0 && C > 0 && dy && dx, DA_ERR_INVALID_ARG, "global_avgpool_backward: bad args");
  const int64_t total = (int64_t)N * HW * C;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (dx_dtype == DA_F32) avgpool_bwd_kernel<float><<<(int)blocks, 256, 0, st>>>(dy, HW, C, (float*)dx, total);
  else if (dx_dtype == DA_BF16 && (C & 7) == 0 && (((uintptr_t)dx) & 15) == 0) {
    int64_t b8 = (total / 8 + 255) / 256;
    if (b8 > (int64_t)num_sms() * 16) b8 = (int64_t)num_sms() * 16;
    if (b8 < 1) b8 = 1;
    avgpool_bwd_vec8_kernel<<<(int)b8, 256, 0, st>>>(dy, HW, C, (__nv_bfloat16*)dx, total / 8);
  } else if (dx_dtype == DA_BF16) avgpool_bwd_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>(dy, HW, C, (__nv_bfloat16*)dx, total);
  else DA_REQUIRE(false, DA_ERR_INVALID_ARG, "global_avgpool_backward: bad dtype");
  DA_LAUNCH_CHECK();
  return DA_OK;
}

extern "C" int da_softmax_dim0_forward(const float* s, int T, int ldk, float* p, da_stream_t stream) {
  DA_REQUIRE(T > 0 && ldk >= T && s && p, DA_ERR_INVALID_ARG, "softmax_dim0_forward: bad args");
  softmax_dim0_fwd_kernel<<<(T + 31) / 32, dim3(32, 32), 0, (cudaStream_t)stream>>>(s, T, ldk, p);
  DA_LAUNCH_CHECK();
  return DA_OK;
}
extern "C" int da_softmax_dim0_backward(const float* p, const float* dp, int T, int ldk, float* ds, da_stream_t stream) {
  DA_REQUIRE(T > 0 && ldk >= T && p && dp && ds, DA_ERR_INVALID_ARG, "softmax_dim0_backward: bad args");
  softmax_dim0_bwd_kernel<<<(T + 31) / 32, dim3(32, 32), 0, (cudaStream_t)stream>>>(p, dp, T, ldk, ds);
  DA_LAUNCH_CHECK();
  return DA_OK;
}

extern "C" int da_weighted_sum_forward(const float* const* losses_host, const float* weights_host, int n, float* scaled, float* total,
                                       da_stream_t stream) {
  DA_REQUIRE(n > 0 && n <= DA_MAX_WEIGHTED && losses_host && weights_host && scaled && total, DA_ERR_INVALID_ARG, "weighted_sum_forward: bad args (n=%d)", n);
  ScalarPtrs a;
  for (int i = 0; i < DA_MAX_WEIGHTED; ++i) { a.p[i] = i < n ? losses_host[i] : nullptr; a.w[i] = i < n ? weights_host[i] : 0.f; }
  weighted_sum_fwd_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a, n, scaled, total);
  DA_LAUNCH_CHECK();
  return DA_OK;
}
extern "C" int da_weighted_sum_backward(const float* weights_host, int n, const float* grad_total, const float* grad_scaled, float* d_losses,
                                        da_stream_t stream) {
  DA_REQUIRE(n > 0 && n <= DA_MAX_WEIGHTED && weights_host && d_losses, DA_ERR_INVALID_ARG, "weighted_sum_backward: bad args (n=%d)", n);
  ScalarPtrs a;
  for (int i = 0; i < DA_MAX_WEIGHTED; ++i) { a.p[i] = nullptr; a.w[i] = i < n ? weights_host[i] : 0.f; }
  weighted_sum_bwd_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a, n, grad_total, grad_scaled, d_losses);
  DA_LAUNCH_CHECK();
  return DA_OK;
}
