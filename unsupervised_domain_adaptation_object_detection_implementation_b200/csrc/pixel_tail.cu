// north_star kernel (1), terminal part: the 1-channel conv that ends a pixel-level domain classifier (ImgAlignmentHead conv2,
// resnet_da_daf_org.py:125,131; LocalAlignmentHead conv3, resnet_da_cbam.py:87,112) fused with its activation, the per-pixel
// domain loss and the mean -- ONE kernel forward -- and, backward, fused with the loss derivative, the terminal conv's data /
// weight / bias gradients AND the activation derivative (ReLU, dropout, folded BN scale) of the producing conv layer, emitting
// the gradient of that layer's accumulator directly -- ONE kernel + one small deterministic reduction.  The producing
// conv itself is the tcgen05 implicit GEMM of umma_conv.cu; da_grl_conv_loss_forward/backward are the composite entry points
// (conv -> tail, tail -> weight gradient -> data gradient with the GRL weight folded into its epilogue).
//
// Per-pixel loss modes:
//   DAF_SQ_BATCH  L1  0.5*mean(sigmoid(p)^2) per source slot + 0.5*mean(sigmoid(1-p)^2) per target slot, means over the WHOLE
//                     batch (resnet_da_daf_org.py:816-822, Q5)
//   DAF_SQ_IMAGE  L2  the same integrands, mean per image (resnet_da_cbam.py:971-979)
//   BCE               F.binary_cross_entropy_with_logits(p, domain of the image), mean over all pixels ("per-level mean")
//   FOCAL             py_sigmoid_focal_loss(p, domain, gamma, alpha), mean over all pixels (losses/focal_loss.py:12-57)
// Reductions are deterministic: per-block partials, the last block to finish (ticket) sums them in block order.
#include "da_common.cuh"
#include <float.h>

namespace da {
size_t umma_workspace_bytes(const da_conv_desc* d);
size_t simt_workspace_bytes(const da_conv_desc* d);

constexpr int PT_THREADS = 256;
constexpr int PT_WARPS = PT_THREADS / 32;

struct TailArgs {
  int N;
  long long L;        // pixels per image
  int K;
  int relu, mode;
  float gamma, alpha;
  const float* w;
  const float* bias;
  const int32_t* domain;
};

__device__ __forceinline__ float softplus_neg_abs(float p) { return log1pf(expf(-fabsf(p))); }

// value of the per-pixel loss term (before the mean) and its derivative w.r.t. the logit, target t in {0,1}
__device__ __forceinline__ float bce_term(float p, float t) { return fmaxf(p, 0.f) - p * t + softplus_neg_abs(p); }
__device__ __forceinline__ float focal_term(float p, float t, float gamma, float alpha) {
  const float s = sigmoidf_(p);
  const float pt = (1.f - s) * t + s * (1.f - t);
  return bce_term(p, t) * (alpha * t + (1.f - alpha) * (1.f - t)) * powf(pt, gamma);
}
__device__ __forceinline__ float focal_grad(float p, float t, float gamma, float alpha) {
  const float s = sigmoidf_(p);
  const float pt = (1.f - s) * t + s * (1.f - t);
  const float fw = (alpha * t + (1.f - alpha) * (1.f - t));
  const float dpt = (1.f - 2.f * t) * s * (1.f - s);                       // d pt / d p
  const float ptg = powf(pt, gamma);
  const float dptg = (pt > 0.f) ? gamma * ptg / pt * dpt : 0.f;
  return fw * ((s - t) * ptg + bce_term(p, t) * dptg);
}

template <typename T>
__device__ __forceinline__ float tail_dot(const T* __restrict__ x, const float* __restrict__ w_s, int K, int lane);
template <>
__device__ __forceinline__ float tail_dot<float>(const float* __restrict__ x, const float* __restrict__ w_s, int K, int lane) {
  float acc = 0.f;
  for (int i = lane; i < K; i += 32) acc = fmaf(x[i], w_s[i], acc);
  return warp_sum(acc);
}
template <>
__device__ __forceinline__ float tail_dot<__nv_bfloat16>(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w_s, int K, int lane) {
  float acc = 0.f;
  if ((K & 7) == 0) {
    for (int i = lane * 8; i < K; i += 256) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + i);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        acc = fmaf(f.x, w_s[i + 2 * e], acc);
        acc = fmaf(f.y, w_s[i + 2 * e + 1], acc);
      }
    }
  } else {
    for (int i = lane; i < K; i += 32) acc = fmaf(__bfloat162float(x[i]), w_s[i], acc);
  }
  return warp_sum(acc);
}

// two rows at once: both rows' loads are in flight before the first FMA (the per-pixel chain load -> dot -> shuffle tree is
// latency bound)
__device__ __forceinline__ void tail_dot2_bf16(const __nv_bfloat16* __restrict__ x0, const __nv_bfloat16* __restrict__ x1,
                                               const float* __restrict__ w_s, int K, int lane, float& p0, float& p1) {
  float a0 = 0.f, a1 = 0.f;
#pragma unroll 2
  for (int i = lane * 8; i < K; i += 256) {
    const uint4 u0 = *reinterpret_cast<const uint4*>(x0 + i);
    const uint4 u1 = *reinterpret_cast<const uint4*>(x1 + i);
    const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&u0);
    const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&u1);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f0 = __bfloat1622float2(h0[e]), f1 = __bfloat1622float2(h1[e]);
      const float w0 = w_s[i + 2 * e], w1 = w_s[i + 2 * e + 1];
      a0 = fmaf(f0.x, w0, a0); a0 = fmaf(f0.y, w1, a0);
      a1 = fmaf(f1.x, w0, a1); a1 = fmaf(f1.y, w1, a1);
    }
  }
  p0 = warp_sum(a0);
  p1 = warp_sum(a1);
}

// grid (blocks per image, N).  partial[(n*nb + blk)*2 + {0,1}]; ticket: last block reduces.
template <typename T>
__global__ void __launch_bounds__(PT_THREADS)
pixel_tail_fwd_kernel(const T* __restrict__ h, TailArgs a, float* __restrict__ logits, float* __restrict__ partial,
                      unsigned int* __restrict__ ticket, float* __restrict__ loss_out) {
  extern __shared__ __align__(16) float w_s[];
  __shared__ float red[33];
  __shared__ bool last;
  pdl_wait();
  for (int i = threadIdx.x; i < a.K; i += PT_THREADS) w_s[i] = a.w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n = blockIdx.y, nb = gridDim.x;
  const float bv = a.bias ? a.bias[0] : 0.f;
  const int d = a.domain[n];
  const float t = d == 1 ? 1.f : 0.f;
  float s0 = 0.f, s1 = 0.f;
  const long long stride = (long long)nb * PT_WARPS;
  auto account = [&](long long m, float p) {
    p += bv;
    if (a.relu) p = fmaxf(p, 0.f);
    if (lane == 0) {
      logits[m] = p;
      if (a.mode <= 1) {
        const float x0 = sigmoidf_(p), x1 = sigmoidf_(1.f - p);
        s0 = fmaf(x0, x0, s0);
        s1 = fmaf(x1, x1, s1);
      } else if (d == 0 || d == 1) {
        s0 += a.mode == 2 ? bce_term(p, t) : focal_term(p, t, a.gamma, a.alpha);
      }
    }
  };
  long long i = (long long)blockIdx.x * PT_WARPS + wid;
  if (sizeof(T) == 2 && (a.K & 7) == 0) {
    for (; i + stride < a.L; i += 2 * stride) {      // same per-pixel order of the partial sums as the one-row loop
      const long long m0 = (long long)n * a.L + i, m1 = m0 + stride;
      float p0, p1;
      tail_dot2_bf16(reinterpret_cast<const __nv_bfloat16*>(h) + (size_t)m0 * a.K, reinterpret_cast<const __nv_bfloat16*>(h) + (size_t)m1 * a.K,
                     w_s, a.K, lane, p0, p1);
      account(m0, p0);
      account(m1, p1);
    }
  }
  for (; i < a.L; i += stride) {
    const long long m = (long long)n * a.L + i;
    account(m, tail_dot<T>(h + (size_t)m * a.K, w_s, a.K, lane));
  }
  s0 = block_sum<false>(s0, red);
  s1 = block_sum<false>(s1, red);
  if (threadIdx.x == 0) {
    partial[((size_t)n * nb + blockIdx.x) * 2 + 0] = s0;
    partial[((size_t)n * nb + blockIdx.x) * 2 + 1] = s1;
    __threadfence();
    last = atomicAdd(ticket, 1u) == (unsigned)(nb * a.N) - 1u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the last block: fixed-order sum over (image, block)
  if (threadIdx.x < 32) {
    float total = 0.f;
    if (a.mode == 0) {
      float t0 = 0.f, t1 = 0.f;
      for (int i = lane; i < a.N * nb; i += 32) { t0 += partial[2 * i]; t1 += partial[2 * i + 1]; }
      t0 = warp_sum(t0); t1 = warp_sum(t1);
      int n_src = 0, n_tgt = 0;
      for (int i = 0; i < a.N; ++i) { n_src += (a.domain[i] == 0); n_tgt += (a.domain[i] == 1); }
      const float denom = (float)((double)a.N * (double)a.L);
      total = 0.5f * ((float)n_src * (t0 / denom) + (float)n_tgt * (t1 / denom));
    } else if (a.mode == 1) {
      for (int img = 0; img < a.N; ++img) {
        const int dd = a.domain[img];
        float s = 0.f;
        for (int i = lane; i < nb; i += 32) s += partial[((size_t)img * nb + i) * 2 + (dd == 1 ? 1 : 0)];
        s = warp_sum(s);
        if (dd == 0 || dd == 1) total += 0.5f * (s / (float)a.L);
      }
    } else {
      float s = 0.f;
      for (int i = lane; i < a.N * nb; i += 32) s += partial[2 * i];
      total = warp_sum(s) / (float)((double)a.N * (double)a.L);
    }
    if (lane == 0) { loss_out[0] = total; *ticket = 0u; }     // ticket re-armed for the next launch
  }
}

// d loss / d logit of pixel (n, .), including the mean's divisor and the upstream scalar g
__device__ __forceinline__ float tail_dlogit(const TailArgs& a, float p, int d, float g, float ca, float cb) {
  if (a.mode <= 1) {
    const float x0 = sigmoidf_(p), x1 = sigmoidf_(1.f - p);
    return ca * x0 * x0 * (1.f - x0) - cb * x1 * x1 * (1.f - x1);
  }
  if (d != 0 && d != 1) return 0.f;
  const float t = d == 1 ? 1.f : 0.f;
  const float inv = g / (float)((double)a.N * (double)a.L);
  return inv * (a.mode == 2 ? (sigmoidf_(p) - t) : focal_grad(p, t, a.gamma, a.alpha));
}

// Backward: a thread owns 8 consecutive channels and every SUB-th row of the block's row range (bf16 y, K % 8 == 0,
// 256 % (K/8) == 0 or K/8 > 256 handled by the scalar kernel).  Per element:
//   dy = dl * w[k];  dv = y > 0 ? dy * keep_scale : 0   (ReLU + dropout derivative from the stored post-activation);
//   dz = dv * scale[k];  dw_tail[k] += dl * y;  dshift[k] += dv;  dvdot[k] += dv * y / keep_scale.
// partial layout per block: [4][K] (dw_tail, dshift, dvdot, -) then the bias partial at [nb*4K + blk].
template <typename T>
__global__ void __launch_bounds__(PT_THREADS)
pixel_tail_bwd_kernel(const T* __restrict__ y, TailArgs a, const float* __restrict__ logits, const float* __restrict__ grad_loss,
                      float loss_scale, const float* __restrict__ grad_logits, const float* __restrict__ scale, int act_relu,
                      float keep_scale, T* __restrict__ dz, float* __restrict__ partial, int rows_per_block, int want_stats) {
  extern __shared__ float pt_red[];   // dl[rows_per_block] | [3][SUB][K] when SUB > 1
  pdl_wait();
  const long long M = (long long)a.N * a.L;
  const long long m0 = (long long)blockIdx.x * rows_per_block;
  const long long m1 = (m0 + rows_per_block < M) ? m0 + rows_per_block : M;
  const int K = a.K;
  const int G = (K + 7) >> 3;
  const int SUB = G <= PT_THREADS ? PT_THREADS / G : 1;
  const float g = (grad_loss ? grad_loss[0] : 1.f) * loss_scale;
  int n_src = 0, n_tgt = 0;
  if (a.mode == 0) for (int i = 0; i < a.N; ++i) { n_src += (a.domain[i] == 0); n_tgt += (a.domain[i] == 1); }
  // d loss / d logit of the block's rows, ONCE per row (the transcendental part), with the ReLU mask of the logit
  float* dl_s = pt_red;
  float* red_s = pt_red + rows_per_block;
  for (long long m = m0 + threadIdx.x; m < m1; m += PT_THREADS) {
    const int n = (int)(m / a.L);
    const int d = a.domain[n];
    float ca = 0.f, cb = 0.f;
    if (a.mode == 0) {
      const float denom = (float)((double)a.N * (double)a.L);
      ca = g * (float)n_src / denom; cb = g * (float)n_tgt / denom;
    } else if (a.mode == 1) {
      ca = d == 0 ? g / (float)a.L : 0.f; cb = d == 1 ? g / (float)a.L : 0.f;
    }
    const float p = logits[m];
    float dl = tail_dlogit(a, p, d, g, ca, cb);
    if (grad_logits) dl += grad_logits[m];
    if (a.relu && !(p > 0.f)) dl = 0.f;
    dl_s[m - m0] = dl;
  }
  __syncthreads();
  const float inv_keep = 1.f / keep_scale;
  float dbias = 0.f;
  for (int cg = threadIdx.x % G; cg < G; cg += (G <= PT_THREADS ? G : PT_THREADS)) {
    const int sub = G <= PT_THREADS ? threadIdx.x / G : 0;
    if (sub >= SUB) break;
    const int k0 = cg << 3;
    const int nk = min(8, K - k0);
    float wk[8], sc[8], aw[8], as[8], ad[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      wk[j] = j < nk ? a.w[k0 + j] : 0.f;
      sc[j] = (scale && j < nk) ? scale[k0 + j] : 1.f;
      aw[j] = as[j] = ad[j] = 0.f;
    }
#pragma unroll 2
    for (long long m = m0 + sub; m < m1; m += SUB) {
      const float dl = dl_s[m - m0];
      if (cg == 0) dbias += dl;
      float yv[8], ov[8];
      if (nk == 8 && sizeof(T) == 2) {
        const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(y) + (size_t)m * K + k0);
        const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(&u);
#pragma unroll
        for (int j = 0; j < 8; ++j) yv[j] = __bfloat162float(hp[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) yv[j] = j < nk ? to_f32<T>(y[(size_t)m * K + k0 + j]) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dy = dl * wk[j];
        const float dv = (!act_relu || yv[j] > 0.f) ? dy * keep_scale : 0.f;
        aw[j] = fmaf(dl, yv[j], aw[j]);
        as[j] += dv;
        ad[j] = fmaf(dv, yv[j] * inv_keep, ad[j]);
        ov[j] = dv * sc[j];
      }
      if (nk == 8 && sizeof(T) == 2) {
        uint4 u;
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(&u);
#pragma unroll
        for (int j = 0; j < 8; ++j) op[j] = __float2bfloat16_rn(ov[j]);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(dz) + (size_t)m * K + k0) = u;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < nk) dz[(size_t)m * K + k0 + j] = from_f32<T>(ov[j]);
      }
    }
    float* out = partial + (size_t)blockIdx.x * 4 * K;
    if (SUB > 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j < nk) {
        red_s[((size_t)0 * SUB + sub) * K + k0 + j] = aw[j];
        red_s[((size_t)1 * SUB + sub) * K + k0 + j] = as[j];
        red_s[((size_t)2 * SUB + sub) * K + k0 + j] = ad[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j < nk) { out[k0 + j] = aw[j]; out[K + k0 + j] = as[j]; out[2 * K + k0 + j] = ad[j]; }
    }
  }
  if (SUB > 1) {
    __syncthreads();
    float* out = partial + (size_t)blockIdx.x * 4 * K;
    for (int i = threadIdx.x; i < 3 * K; i += PT_THREADS) {
      const int which = i / K, k = i - which * K;
      float s = 0.f;
      for (int q = 0; q < SUB; ++q) s += red_s[((size_t)which * SUB + q) * K + k];
      out[which * K + k] = s;
    }
  }
  // bias partial: the threads of channel group 0 hold disjoint rows; fixed-order sum through shared memory
  __shared__ float db_s[PT_THREADS];
  db_s[threadIdx.x] = dbias;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < PT_THREADS; ++i) s += db_s[i];
    partial[(size_t)gridDim.x * 4 * K + blockIdx.x] = s;
  }
  (void)want_stats;
}

// out[0..3K) = sum over blocks of partial[blk][0..3K); out[3K] = sum of the bias partials.  Block = 32 columns x 8 block
// lanes; the 8 lane sums meet in shared memory in a fixed order (deterministic).
constexpr int PT_FINAL_LANES = 32;      // block lanes per column (1024 threads): ~9 partials per thread, four loads in flight
__global__ void __launch_bounds__(32 * PT_FINAL_LANES)
pixel_tail_final_kernel(const float* __restrict__ partial, int nb, int K, float* __restrict__ dw_tail,
                        float* __restrict__ dshift, float* __restrict__ dvdot, float* __restrict__ dbias) {
  __shared__ float red[PT_FINAL_LANES][33];
  pdl_wait();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (i < 3 * K) {
    const int which = i / K, k = i - which * K;
#pragma unroll 4
    for (int b = ty; b < nb; b += PT_FINAL_LANES) s += partial[(size_t)b * 4 * K + which * K + k];
  } else if (i == 3 * K) {
#pragma unroll 4
    for (int b = ty; b < nb; b += PT_FINAL_LANES) s += partial[(size_t)nb * 4 * K + b];
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < PT_FINAL_LANES; ++q) t += red[q][tx];
    if (i < 3 * K) {
      const int which = i / K, k = i - which * K;
      float* dst = which == 0 ? dw_tail : (which == 1 ? dshift : dvdot);
      if (dst) dst[k] = t;
    } else if (i == 3 * K && dbias) {
      dbias[0] = t;
    }
  }
}

static int tail_blocks_fwd(long long L) {
  long long b = (L + PT_WARPS * 2 - 1) / (PT_WARPS * 2);     // two pixels per warp: the per-pixel chain (load, shuffle tree) is latency bound
  const long long cap = 4 * (long long)num_sms();
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
static int tail_rows_per_block(long long M) {
  long long nb = 2 * (long long)num_sms();
  long long rows = (M + nb - 1) / nb;
  if (rows < 16) rows = 16;
  return (int)rows;
}

template <typename F>
static int launch_pdl(F kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, void** args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_opt.no_pdl ? 0 : 1;
  DA_CUDA_OK(cudaLaunchKernelExC(&cfg, (const void*)kern, args));
  DA_LAUNCH_CHECK();
  return DA_OK;
}

}  // namespace da

using namespace da;

static int tail_args(const da_conv_desc* d, const da_pixel_tail* t, TailArgs* a, const char* who) {
  DA_REQUIRE(d && t && t->w && t->domain, DA_ERR_INVALID_ARG, "%s: null argument", who);
  DA_REQUIRE(t->mode >= DA_PIXEL_LOSS_DAF_SQ_BATCH && t->mode <= DA_PIXEL_LOSS_FOCAL, DA_ERR_INVALID_ARG, "%s: unknown loss mode %d", who, t->mode);
  const int OH = (d->H + 2 * d->pad - d->KH) / d->stride + 1, OW = (d->W + 2 * d->pad - d->KW) / d->stride + 1;
  a->N = d->N; a->L = (long long)OH * OW; a->K = d->Cout; a->relu = t->relu; a->mode = t->mode; a->gamma = t->gamma; a->alpha = t->alpha;
  a->w = t->w; a->bias = t->bias; a->domain = t->domain;
  DA_REQUIRE((size_t)a->K * 4 <= 160 * 1024, DA_ERR_UNSUPPORTED, "%s: Cout=%d too large for the tail kernel", who, a->K);
  return DA_OK;
}

static size_t tail_ws_bytes(const da_conv_desc* d) {
  const int OH = (d->H + 2 * d->pad - d->KH) / d->stride + 1, OW = (d->W + 2 * d->pad - d->KW) / d->stride + 1;
  const long long M = (long long)d->N * OH * OW;
  const long long nbf = (long long)tail_blocks_fwd((long long)OH * OW) * d->N;
  const int rows = tail_rows_per_block(M);
  const long long nbb = (M + rows - 1) / rows;
  const size_t fwd = 256 + (size_t)nbf * 2 * sizeof(float);
  const size_t bwd = 256 + (size_t)nbb * (4 * (size_t)d->Cout + 1) * sizeof(float);
  return align_up(fwd > bwd ? fwd : bwd, 256);
}

extern "C" size_t da_grl_conv_loss_workspace_bytes(const da_conv_desc* d) {
  if (!d) return 0;
  return align_up(da_conv_workspace_bytes(d), 256) + tail_ws_bytes(d);
}

extern "C" int da_pixel_tail_forward(const da_conv_desc* d, const void* y, const da_pixel_tail* tail, float* logits, float* loss,
                                     void* workspace, size_t workspace_bytes, da_stream_t stream) {
  TailArgs a;
  int rc = tail_args(d, tail, &a, "pixel_tail_forward");
  if (rc) return rc;
  DA_REQUIRE(y && logits && loss, DA_ERR_INVALID_ARG, "pixel_tail_forward: null tensor");
  DA_REQUIRE(workspace && workspace_bytes >= tail_ws_bytes(d), DA_ERR_WORKSPACE, "pixel_tail_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* ticket = (unsigned int*)workspace;
  float* partial = (float*)((uint8_t*)workspace + 256);
  static bool zeroed_dev[kMaxDevices] = {};
  (void)zeroed_dev;
  DA_CUDA_OK(cudaMemsetAsync(ticket, 0, 16, st));
  const dim3 grid(tail_blocks_fwd(a.L), a.N);
  const size_t smem = (size_t)a.K * sizeof(float);
  void* args[] = {(void*)&y, (void*)&a, (void*)&logits, (void*)&partial, (void*)&ticket, (void*)&loss};
  if (d->y_dtype == DA_BF16) {
    auto k = pixel_tail_fwd_kernel<__nv_bfloat16>;
    if (smem > 48 * 1024) DA_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return launch_pdl(k, grid, dim3(PT_THREADS), smem, st, args);
  }
  auto k = pixel_tail_fwd_kernel<float>;
  if (smem > 48 * 1024) DA_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return launch_pdl(k, grid, dim3(PT_THREADS), smem, st, args);
}

extern "C" int da_pixel_tail_backward(const da_conv_desc* d, const void* y, const da_pixel_tail* tail, const float* logits,
                                      const float* grad_loss, float loss_scale, const float* grad_logits, const float* scale,
                                      int act_relu, float drop_p, void* dz, float* dw_tail, float* dbias_tail, float* dshift,
                                      float* dvdot, void* workspace, size_t workspace_bytes, da_stream_t stream) {
  TailArgs a;
  int rc = tail_args(d, tail, &a, "pixel_tail_backward");
  if (rc) return rc;
  DA_REQUIRE(y && logits && dz && dw_tail, DA_ERR_INVALID_ARG, "pixel_tail_backward: null tensor");
  DA_REQUIRE(workspace && workspace_bytes >= tail_ws_bytes(d), DA_ERR_WORKSPACE, "pixel_tail_backward: workspace too small");
  DA_REQUIRE(drop_p >= 0.f && drop_p < 1.f, DA_ERR_INVALID_ARG, "pixel_tail_backward: drop_p out of [0,1)");
  DA_REQUIRE(drop_p == 0.f || act_relu, DA_ERR_UNSUPPORTED,
             "pixel_tail_backward: dropout without ReLU (the mask is taken from the stored activation: a dropped unit must store exactly 0)");
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)((uint8_t*)workspace + 256);
  const long long M = (long long)a.N * a.L;
  int rows = tail_rows_per_block(M);
  const int nb = (int)((M + rows - 1) / rows);
  const int G = (a.K + 7) >> 3;
  const int SUB = G <= PT_THREADS ? PT_THREADS / G : 1;
  const size_t smem = ((size_t)rows + (SUB > 1 ? (size_t)3 * SUB * a.K : 0)) * sizeof(float);
  const float keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  int want = 1;
  void* args[] = {(void*)&y, (void*)&a, (void*)&logits, (void*)&grad_loss, (void*)&loss_scale, (void*)&grad_logits, (void*)&scale,
                  (void*)&act_relu, (void*)&keep, (void*)&dz, (void*)&partial, (void*)&rows, (void*)&want};
  if (d->y_dtype == DA_BF16) {
    auto k = pixel_tail_bwd_kernel<__nv_bfloat16>;
    if (smem > 48 * 1024) DA_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rc = launch_pdl(k, dim3(nb), dim3(PT_THREADS), smem, st, args);
  } else {
    auto k = pixel_tail_bwd_kernel<float>;
    if (smem > 48 * 1024) DA_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rc = launch_pdl(k, dim3(nb), dim3(PT_THREADS), smem, st, args);
  }
  if (rc) return rc;
  const float* cpartial = partial;
  int K = a.K, nbv = nb;
  void* fargs[] = {(void*)&cpartial, (void*)&nbv, (void*)&K, (void*)&dw_tail, (void*)&dshift, (void*)&dvdot, (void*)&dbias_tail};
  return launch_pdl(pixel_tail_final_kernel, dim3((3 * a.K + 1 + 31) / 32), dim3(32 * PT_FINAL_LANES), 0, st, fargs);
}

// Composite entry points: producing conv (tcgen05 implicit GEMM, fused BN/bias + ReLU + dropout epilogue) -> tail.
extern "C" int da_grl_conv_loss_forward(const da_conv_desc* d, const void* x, const void* w, const float* scale, const float* shift,
                                        int relu, float drop_p, uint64_t drop_seed, void* y, const da_pixel_tail* tail, float* logits,
                                        float* loss, void* workspace, size_t workspace_bytes, da_stream_t stream) {
  DA_REQUIRE(d != nullptr, DA_ERR_INVALID_ARG, "grl_conv_loss_forward: null descriptor");
  DA_REQUIRE(workspace && workspace_bytes >= da_grl_conv_loss_workspace_bytes(d), DA_ERR_WORKSPACE, "grl_conv_loss_forward: workspace too small");
  const size_t cw = align_up(da_conv_workspace_bytes(d), 256);
  int rc = da_conv_forward(d, x, w, scale, shift, relu, drop_p, drop_seed, y, workspace, cw, stream);
  if (rc) return rc;
  return da_pixel_tail_forward(d, y, tail, logits, loss, (uint8_t*)workspace + cw, workspace_bytes - cw, stream);
}

extern "C" int da_grl_conv_loss_backward(const da_conv_desc* d, const void* x, const void* w, const float* scale, int relu,
                                         float drop_p, const void* y, const da_pixel_tail* tail, const float* logits,
                                         const float* grad_loss, float loss_scale, const float* grad_logits, float grl, void* dx,
                                         float* dw, float* dshift, float* dvdot, float* dw_tail, float* dbias_tail, void* dz_scratch,
                                         void* workspace, size_t workspace_bytes, da_stream_t stream) {
  DA_REQUIRE(d != nullptr, DA_ERR_INVALID_ARG, "grl_conv_loss_backward: null descriptor");
  DA_REQUIRE(d->x_dtype == d->y_dtype, DA_ERR_UNSUPPORTED, "grl_conv_loss_backward: the layer's input and output dtypes must agree");
  DA_REQUIRE(workspace && workspace_bytes >= da_grl_conv_loss_workspace_bytes(d), DA_ERR_WORKSPACE, "grl_conv_loss_backward: workspace too small");
  const size_t cw = align_up(da_conv_workspace_bytes(d), 256);
  int rc = da_pixel_tail_backward(d, y, tail, logits, grad_loss, loss_scale, grad_logits, scale, relu, drop_p, dz_scratch, dw_tail,
                                  dbias_tail, dshift, dvdot, (uint8_t*)workspace + cw, workspace_bytes - cw, stream);
  if (rc) return rc;
  if (dw) {
    rc = da_conv_backward_weight(d, x, dz_scratch, dw, workspace, cw, stream);
    if (rc) return rc;
  }
  if (dx) rc = da_conv_backward_data(d, dz_scratch, w, grl, dx, workspace, cw, stream);
  return rc;
}
