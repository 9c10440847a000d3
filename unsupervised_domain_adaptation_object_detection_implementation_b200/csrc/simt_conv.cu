// CUDA-core fp32 implicit-GEMM engine (DA_ENGINE_SIMT_F32) for the domain-classifier convs
// and FC stacks, plus the engine-independent activation backward.
//
// This is the fp32 PARITY engine: plain FMA accumulation in fp32, so losses/gradients match
// the reference's fp32 PyTorch math to ~1e-6.  The throughput engine is the tcgen05 one in
// umma_conv.cu; both sit behind da_conv_forward / da_conv_backward_{data,weight}.
//
// GEMM views (NHWC activations, OHWI weights):
//   forward : Y[m,co]  = sum_{tap,ci} X[pix(m,tap),ci] * W[co,tap,ci]      M=N*OH*OW
//   dgrad   : dX[p,ci] = sum_{tap,co} dZ[opix(p,tap),co] * W[co,tap,ci]    M=N*H*W
//   wgrad   : dW[co,tap,ci] = sum_m dZ[m,co] * X[pix(m,tap),ci]            K=N*OH*OW
#include "da_common.cuh"

namespace da {

constexpr int TM = 64, TN = 64, TK = 16;
constexpr int LDS_ = TM + 4;

struct ConvGeom {
  int N, H, W, Cin, Cout, KH, KW, stride, pad, OH, OW;
};

template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                     __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
}

__device__ __forceinline__ void mma_tile(const float (*As)[LDS_], const float (*Bs)[LDS_], int ty, int tx,
                                         float acc[4][4]) {
#pragma unroll
  for (int k = 0; k < TK; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
    const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// ---- forward ---------------------------------------------------------------------------
template <typename T, typename TY>
__global__ void __launch_bounds__(256)
simt_conv_fwd_kernel(ConvGeom g, const T* __restrict__ x, const T* __restrict__ w,
                     const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                     float drop_p, uint64_t seed0, const unsigned long long* seed_ctr, TY* __restrict__ y) {
  const uint64_t seed = effective_seed(seed0, seed_ctr);
  __shared__ __align__(16) float As[TK][LDS_];
  __shared__ __align__(16) float Bs[TK][LDS_];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t M = (int64_t)g.N * g.OH * g.OW;
  const int64_t m0 = (int64_t)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  const int taps = g.KH * g.KW;
  const int Ktot = taps * g.Cin;

  // loader role: row lr (0..63), k-quad lq (0..3)
  const int lr = tid >> 2, lq = tid & 3;
  const int64_t am = m0 + lr;
  int an = 0, aoh = 0, aow = 0;
  const bool arow_ok = am < M;
  if (arow_ok) { an = (int)(am / (g.OH * g.OW)); const int rem = (int)(am % (g.OH * g.OW)); aoh = rem / g.OW; aow = rem % g.OW; }
  const int bn = n0 + lr;
  const bool brow_ok = bn < g.Cout;

  const bool vec = (g.Cin % TK) == 0;  // fast path: a 16-wide K chunk never straddles taps
  float acc[4][4] = {};
  for (int k0 = 0; k0 < Ktot; k0 += TK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
    if (vec) {
      const int tap = k0 / g.Cin, ci0 = k0 % g.Cin;
      const int kh = tap / g.KW, kw = tap % g.KW;
      if (arow_ok) {
        const int ih = aoh * g.stride + kh - g.pad, iw = aow * g.stride + kw - g.pad;
        if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
          av = load4<T>(x + (((size_t)an * g.H + ih) * g.W + iw) * g.Cin + ci0 + lq * 4);
      }
      if (brow_ok) bv = load4<T>(w + (size_t)bn * Ktot + k0 + lq * 4);
    } else {
      float ae[4] = {0.f, 0.f, 0.f, 0.f}, be[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + lq * 4 + e;
        if (k >= Ktot) continue;
        const int tap = k / g.Cin, ci = k % g.Cin;
        const int kh = tap / g.KW, kw = tap % g.KW;
        if (arow_ok) {
          const int ih = aoh * g.stride + kh - g.pad, iw = aow * g.stride + kw - g.pad;
          if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
            ae[e] = to_f32<T>(x[(((size_t)an * g.H + ih) * g.W + iw) * g.Cin + ci]);
        }
        if (brow_ok) be[e] = to_f32<T>(w[(size_t)bn * Ktot + k]);
      }
      av = make_float4(ae[0], ae[1], ae[2], ae[3]);
      bv = make_float4(be[0], be[1], be[2], be[3]);
    }
    __syncthreads();
    As[lq * 4 + 0][lr] = av.x; As[lq * 4 + 1][lr] = av.y; As[lq * 4 + 2][lr] = av.z; As[lq * 4 + 3][lr] = av.w;
    Bs[lq * 4 + 0][lr] = bv.x; Bs[lq * 4 + 1][lr] = bv.y; Bs[lq * 4 + 2][lr] = bv.z; Bs[lq * 4 + 3][lr] = bv.w;
    __syncthreads();
    mma_tile(As, Bs, ty, tx, acc);
  }
  const uint32_t thr = drop_threshold(drop_p);
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= g.Cout) continue;
      float v = acc[i][j];
      if (scale) v *= scale[c];
      if (shift) v += shift[c];
      if (relu) v = fmaxf(v, 0.f);
      if (drop_p > 0.f) v = (drop_hash(seed, (uint64_t)m * g.Cout + c) >= thr) ? v * keep_scale : 0.f;
      y[(size_t)m * g.Cout + c] = from_f32<TY>(v);
    }
  }
}

// ---- dgrad -----------------------------------------------------------------------------
template <typename T, typename TX>
__global__ void __launch_bounds__(256)
simt_conv_dgrad_kernel(ConvGeom g, const T* __restrict__ dz, const T* __restrict__ w, float out_scale,
                       TX* __restrict__ dx) {
  __shared__ __align__(16) float As[TK][LDS_];
  __shared__ __align__(16) float Bs[TK][LDS_];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t M = (int64_t)g.N * g.H * g.W;
  const int64_t m0 = (int64_t)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;  // over Cin
  const int taps = g.KH * g.KW;
  const int Ktot = taps * g.Cout;

  const int lr = tid >> 2, lq = tid & 3;  // A loader: row lr, k-quad lq
  const int64_t am = m0 + lr;
  int an = 0, ah = 0, aw = 0;
  const bool arow_ok = am < M;
  if (arow_ok) { an = (int)(am / (g.H * g.W)); const int rem = (int)(am % (g.H * g.W)); ah = rem / g.W; aw = rem % g.W; }
  // B loader: k row bk (0..15), column quad bq (0..15)
  const int bk = tid >> 4, bq = tid & 15;

  const bool vec = (g.Cout % TK) == 0 && (g.Cin % 4) == 0;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < Ktot; k0 += TK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
    if (vec) {
      const int tap = k0 / g.Cout, co0 = k0 % g.Cout;
      const int kh = tap / g.KW, kw = tap % g.KW;
      if (arow_ok) {
        const int th = ah + g.pad - kh, tw = aw + g.pad - kw;
        if (th >= 0 && tw >= 0 && th % g.stride == 0 && tw % g.stride == 0) {
          const int oh = th / g.stride, ow = tw / g.stride;
          if (oh < g.OH && ow < g.OW)
            av = load4<T>(dz + (((size_t)an * g.OH + oh) * g.OW + ow) * g.Cout + co0 + lq * 4);
        }
      }
      const int ci = n0 + bq * 4;
      if (ci < g.Cin) bv = load4<T>(w + ((size_t)(co0 + bk) * taps + tap) * g.Cin + ci);
    } else {
      float ae[4] = {0.f, 0.f, 0.f, 0.f}, be[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + lq * 4 + e;
        if (k < Ktot && arow_ok) {
          const int tap = k / g.Cout, co = k % g.Cout;
          const int kh = tap / g.KW, kw = tap % g.KW;
          const int th = ah + g.pad - kh, tw = aw + g.pad - kw;
          if (th >= 0 && tw >= 0 && th % g.stride == 0 && tw % g.stride == 0) {
            const int oh = th / g.stride, ow = tw / g.stride;
            if (oh < g.OH && ow < g.OW) ae[e] = to_f32<T>(dz[(((size_t)an * g.OH + oh) * g.OW + ow) * g.Cout + co]);
          }
        }
        const int kb = k0 + bk, ci = n0 + bq * 4 + e;
        if (kb < Ktot && ci < g.Cin) {
          const int tap = kb / g.Cout, co = kb % g.Cout;
          be[e] = to_f32<T>(w[((size_t)co * taps + tap) * g.Cin + ci]);
        }
      }
      av = make_float4(ae[0], ae[1], ae[2], ae[3]);
      bv = make_float4(be[0], be[1], be[2], be[3]);
    }
    __syncthreads();
    As[lq * 4 + 0][lr] = av.x; As[lq * 4 + 1][lr] = av.y; As[lq * 4 + 2][lr] = av.z; As[lq * 4 + 3][lr] = av.w;
    *reinterpret_cast<float4*>(&Bs[bk][bq * 4]) = bv;
    __syncthreads();
    mma_tile(As, Bs, ty, tx, acc);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c < g.Cin) dx[(size_t)m * g.Cin + c] = from_f32<TX>(acc[i][j] * out_scale);
    }
  }
}

// ---- wgrad: grid (Cout/64, Cin/64, taps * ksplit); partial sums into dw_part[ks] ----------
template <typename T>
__global__ void __launch_bounds__(256)
simt_conv_wgrad_kernel(ConvGeom g, const T* __restrict__ x, const T* __restrict__ dz, float* __restrict__ dw_part,
                       int ksplit) {
  __shared__ __align__(16) float As[TK][LDS_];
  __shared__ __align__(16) float Bs[TK][LDS_];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int taps = g.KH * g.KW;
  const int tap = blockIdx.z % taps, ks = blockIdx.z / taps;
  const int kh = tap / g.KW, kw = tap % g.KW;
  const int co0 = blockIdx.x * TM, ci0 = blockIdx.y * TN;
  const int64_t M = (int64_t)g.N * g.OH * g.OW;
  const int64_t per = ((M + ksplit - 1) / ksplit + TK - 1) / TK * TK;
  const int64_t mbeg = (int64_t)ks * per, mend = (mbeg + per < M) ? mbeg + per : M;

  const int lk = tid >> 4, lq = tid & 15;  // pixel row lk (0..15), column quad lq
  const bool vec = (g.Cout % 4) == 0 && (g.Cin % 4) == 0;
  float acc[4][4] = {};
  for (int64_t mb = mbeg; mb < mend; mb += TK) {
    const int64_t m = mb + lk;
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
    if (m < mend) {
      const int n = (int)(m / (g.OH * g.OW)); const int rem = (int)(m % (g.OH * g.OW));
      const int oh = rem / g.OW, ow = rem % g.OW;
      const int co = co0 + lq * 4;
      const int ih = oh * g.stride + kh - g.pad, iw = ow * g.stride + kw - g.pad;
      const int ci = ci0 + lq * 4;
      const bool in_ok = ih >= 0 && ih < g.H && iw >= 0 && iw < g.W;
      if (vec) {
        if (co < g.Cout) av = load4<T>(dz + (size_t)m * g.Cout + co);
        if (ci < g.Cin && in_ok) bv = load4<T>(x + (((size_t)n * g.H + ih) * g.W + iw) * g.Cin + ci);
      } else {
        float ae[4] = {0.f, 0.f, 0.f, 0.f}, be[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (co + e < g.Cout) ae[e] = to_f32<T>(dz[(size_t)m * g.Cout + co + e]);
          if (ci + e < g.Cin && in_ok) be[e] = to_f32<T>(x[(((size_t)n * g.H + ih) * g.W + iw) * g.Cin + ci + e]);
        }
        av = make_float4(ae[0], ae[1], ae[2], ae[3]);
        bv = make_float4(be[0], be[1], be[2], be[3]);
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[lk][lq * 4]) = av;
    *reinterpret_cast<float4*>(&Bs[lk][lq * 4]) = bv;
    __syncthreads();
    mma_tile(As, Bs, ty, tx, acc);
  }
  float* out = dw_part + (size_t)ks * g.Cout * taps * g.Cin;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= g.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci < g.Cin) out[((size_t)co * taps + tap) * g.Cin + ci] = acc[i][j];
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ part, int ksplit, int64_t n, float* __restrict__ out) {
  pdl_launch_dependents();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < ksplit; ++s) acc += part[(size_t)s * n + i];
    out[i] = acc;
  }
}

// ---- thin outputs (1x1 conv / FC with Cout <= 4: the 2-logit domain classifiers) -------------
// The 64x64 tile kernels leave 63 of 64 columns empty there (46 us for the [1024,512]x[512,2] logits of the
// instance head); these are plain reductions.  fp32 accumulation like the tile kernels.
constexpr int THIN_MAX = 4;

template <typename T, typename TY>
__global__ void __launch_bounds__(256)
thin_fwd_kernel(int64_t M, int K, int Cout, const T* __restrict__ x, const T* __restrict__ w,
                const float* __restrict__ scale, const float* __restrict__ shift, int relu, float drop_p, uint64_t seed0,
                const unsigned long long* seed_ctr, TY* __restrict__ y) {
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // one warp per row
  if (m >= M) return;
  float acc[THIN_MAX] = {0.f, 0.f, 0.f, 0.f};
  for (int k = lane; k < K; k += 32) {
    const float xv = to_f32<T>(x[(size_t)m * K + k]);
#pragma unroll
    for (int n = 0; n < THIN_MAX; ++n)
      if (n < Cout) acc[n] = fmaf(xv, to_f32<T>(w[(size_t)n * K + k]), acc[n]);
  }
#pragma unroll
  for (int n = 0; n < THIN_MAX; ++n) acc[n] = warp_sum(acc[n]);
  if (lane == 0) {
    const uint64_t seed = effective_seed(seed0, seed_ctr);
    const uint32_t thr = drop_threshold(drop_p);
    const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
#pragma unroll
    for (int n = 0; n < THIN_MAX; ++n) {
      if (n >= Cout) break;
      float v = acc[n];
      if (scale) v *= scale[n];
      if (shift) v += shift[n];
      if (relu) v = fmaxf(v, 0.f);
      if (drop_p > 0.f) v = (drop_hash(seed, (uint64_t)m * Cout + n) >= thr) ? v * keep_scale : 0.f;
      y[(size_t)m * Cout + n] = from_f32<TY>(v);
    }
  }
}

template <typename T, typename TX>
__global__ void __launch_bounds__(256)
thin_dgrad_kernel(int64_t M, int K, int Cout, const T* __restrict__ dz, const T* __restrict__ w, float out_scale,
                  TX* __restrict__ dx) {
  pdl_launch_dependents();
  const int64_t total = M * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / K;
    const int k = (int)(i - m * K);
    float acc = 0.f;
    for (int n = 0; n < Cout; ++n) acc = fmaf(to_f32<T>(dz[(size_t)m * Cout + n]), to_f32<T>(w[(size_t)n * K + k]), acc);
    dx[i] = from_f32<TX>(acc * out_scale);
  }
}

// dW[n,k] = sum_m dz[m,n] x[m,k]: block = 32 k-columns, 8 warps stride the rows of one M-split, fixed-order reduce
template <typename T>
__global__ void __launch_bounds__(256)
thin_wgrad_kernel(int64_t M, int K, int Cout, const T* __restrict__ x, const T* __restrict__ dz, float* __restrict__ part,
                  int msplit) {
  pdl_launch_dependents();
  __shared__ float red[8][THIN_MAX][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + lane;
  const int64_t per = (M + msplit - 1) / msplit;
  const int64_t m_lo = (int64_t)blockIdx.y * per, m_hi = m_lo + per < M ? m_lo + per : M;
  float acc[THIN_MAX] = {0.f, 0.f, 0.f, 0.f};
  if (k < K) {
    for (int64_t m = m_lo + warp; m < m_hi; m += 8) {
      const float xv = to_f32<T>(x[(size_t)m * K + k]);
#pragma unroll
      for (int n = 0; n < THIN_MAX; ++n)
        if (n < Cout) acc[n] = fmaf(to_f32<T>(dz[(size_t)m * Cout + n]), xv, acc[n]);
    }
  }
#pragma unroll
  for (int n = 0; n < THIN_MAX; ++n) red[warp][n][lane] = acc[n];
  __syncthreads();
  if (warp < Cout && k < K) {
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) v += red[wv][warp][lane];
    part[(size_t)blockIdx.y * Cout * K + (size_t)warp * K + k] = v;
  }
}

static inline bool thin_ok(const ConvGeom& g) {
  return g.KH == 1 && g.KW == 1 && g.stride == 1 && g.pad == 0 && g.Cout <= THIN_MAX;
}

// ---- activation backward (engine independent) --------------------------------------------
// dz[m,c] = dy[m,c] * scale[c] * relu'(y) * keep/(1-p);  colsum partial of (dy*mask) per block
// rows per block: enough blocks to fill the machine even for the [1024, C] FC activations
__host__ __device__ inline int ab_rows(int64_t M) {
  int64_t r = (M + 591) / 592;
  return (int)(r < 1 ? 1 : (r > 64 ? 64 : r));
}
template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, int64_t M, int C,
                               const float* __restrict__ scale, int relu, float drop_p, uint64_t seed0,
                               const unsigned long long* seed_ctr,
                               T* __restrict__ dz, float* __restrict__ partial, float* __restrict__ partial2) {
  pdl_launch_dependents();
  const uint64_t seed = effective_seed(seed0, seed_ctr);
  const int AB_ROWS = ab_rows(M);
  const int64_t m0 = (int64_t)blockIdx.x * AB_ROWS;
  const int64_t m1 = (m0 + AB_ROWS < M) ? m0 + AB_ROWS : M;
  const uint32_t thr = drop_threshold(drop_p);
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float sc = scale ? scale[c] : 1.f;
    float colsum = 0.f, colsum2 = 0.f;
    for (int64_t m = m0; m < m1; ++m) {
      const size_t i = (size_t)m * C + c;
      float d = to_f32<T>(dy[i]);
      const float yv = y ? to_f32<T>(y[i]) : 0.f;
      bool on = true;
      if (relu) on = yv > 0.f;
      if (drop_p > 0.f) { on = on && (drop_hash(seed, (uint64_t)i) >= thr); d *= keep_scale; }
      d = on ? d : 0.f;
      colsum += d;
      colsum2 = fmaf(d, yv / keep_scale, colsum2);  // v = acc*scale+shift where the unit is on
      dz[i] = from_f32<T>(d * sc);
    }
    if (partial) partial[(size_t)blockIdx.x * C + c] = colsum;
    if (partial2) partial2[(size_t)blockIdx.x * C + c] = colsum2;
  }
}
// bf16 activations, C % 8 == 0: a thread owns groups of 8 consecutive channels (one 16-byte load of dy and of y, one
// 16-byte store of dz per row).  With G = C/8 <= 128 column groups the 256 threads also split the block's rows SUB = 256/G
// ways and the SUB partial column sums meet in shared memory; with more groups a thread walks the rows alone and takes every
// 256th group.  The scalar kernel above moves 2 bytes per thread and access (ncu: 35 us for the [16384, 512] activation of the
// image head, 480-510 us for the [18760, 4608] activation of the C5 SRM head).
__global__ void __launch_bounds__(256)
act_bwd_vec8_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y, int64_t M, int C,
                    const float* __restrict__ scale, int relu, float drop_p, uint64_t seed0, const unsigned long long* seed_ctr,
                    __nv_bfloat16* __restrict__ dz, float* __restrict__ partial, float* __restrict__ partial2) {
  pdl_launch_dependents();
  extern __shared__ float ab_red[];   // [2][SUB][C] when SUB > 1
  const uint64_t seed = effective_seed(seed0, seed_ctr);
  const int AB_ROWS = ab_rows(M);
  const int64_t m0 = (int64_t)blockIdx.x * AB_ROWS;
  const int64_t m1 = (m0 + AB_ROWS < M) ? m0 + AB_ROWS : M;
  const uint32_t thr = drop_threshold(drop_p);
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const float inv_keep = 1.f / keep_scale;
  const int G = C >> 3;
  const int SUB = G <= 128 ? 256 / G : 1;                 // row-parallel ways (threads beyond G*SUB idle in the main loop)
  const int sub = G <= 128 ? threadIdx.x / G : 0;
  const int cg_first = G <= 128 ? threadIdx.x % G : threadIdx.x;
  const int cg_step = G <= 128 ? G : 256;
  const bool active = sub < SUB;
  for (int cg = cg_first; cg < G; cg += cg_step) {
    const int c0 = cg << 3;
    float sc[8], cs[8], cs2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = scale ? scale[c0 + j] : 1.f; cs[j] = 0.f; cs2[j] = 0.f; }
    if (active) {
      for (int64_t m = m0 + sub; m < m1; m += SUB) {
        const size_t i = (size_t)m * C + c0;
        const uint4 dv = *reinterpret_cast<const uint4*>(dy + i);
        uint4 yv4 = make_uint4(0u, 0u, 0u, 0u);
        if (y) yv4 = *reinterpret_cast<const uint4*>(y + i);
        const __nv_bfloat16* dp = reinterpret_cast<const __nv_bfloat16*>(&dv);
        const __nv_bfloat16* yp = reinterpret_cast<const __nv_bfloat16*>(&yv4);
        uint4 ov;
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(&ov);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float d = __bfloat162float(dp[j]);
          const float yv = y ? __bfloat162float(yp[j]) : 0.f;
          bool on = true;
          if (relu) on = yv > 0.f;
          if (drop_p > 0.f) { on = on && (drop_hash(seed, (uint64_t)(i + j)) >= thr); d *= keep_scale; }
          d = on ? d : 0.f;
          cs[j] += d;
          cs2[j] = fmaf(d, yv * inv_keep, cs2[j]);
          op[j] = __float2bfloat16_rn(d * sc[j]);
        }
        *reinterpret_cast<uint4*>(dz + i) = ov;
      }
    }
    if (!partial && !partial2) continue;
    if (SUB > 1) {                    // G <= 128: exactly one pass of the cg loop, every thread reaches the barrier
      if (active) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          ab_red[(size_t)sub * C + c0 + j] = cs[j];
          ab_red[(size_t)(SUB + sub) * C + c0 + j] = cs2[j];
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (partial) partial[(size_t)blockIdx.x * C + c0 + j] = cs[j];
        if (partial2) partial2[(size_t)blockIdx.x * C + c0 + j] = cs2[j];
      }
    }
  }
  if (SUB > 1 && (partial || partial2)) {
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
      float a = 0.f, b = 0.f;
      for (int k = 0; k < SUB; ++k) { a += ab_red[(size_t)k * C + c]; b += ab_red[(size_t)(SUB + k) * C + c]; }
      if (partial) partial[(size_t)blockIdx.x * C + c] = a;
      if (partial2) partial2[(size_t)blockIdx.x * C + c] = b;
    }
  }
}
static inline bool act_vec8_ok(int C, const void* a, const void* b, const void* c) {
  return (C & 7) == 0 && C >= 8 && ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)c)) & 15) == 0;
}
// out[c] = sum_b partial[b][c]; block (32 columns x 32 row lanes), coalesced rows, smem tree over the lanes
__global__ void colsum_final_kernel(const float* __restrict__ partial, int nb, int C, float* __restrict__ out) {
  pdl_launch_dependents();
  __shared__ float red[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (c < C)
    for (int b = ty; b < nb; b += 32) acc += partial[(size_t)b * C + c];
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) t += red[j][tx];
    out[c] = t;
  }
}

static inline ConvGeom make_geom(const da_conv_desc* d) {
  ConvGeom g;
  g.N = d->N; g.H = d->H; g.W = d->W; g.Cin = d->Cin; g.Cout = d->Cout; g.KH = d->KH; g.KW = d->KW;
  g.stride = d->stride; g.pad = d->pad;
  g.OH = (d->H + 2 * d->pad - d->KH) / d->stride + 1;
  g.OW = (d->W + 2 * d->pad - d->KW) / d->stride + 1;
  return g;
}

int simt_wgrad_ksplit(const ConvGeom& g) {
  const int64_t M = (int64_t)g.N * g.OH * g.OW;
  const int64_t tiles = (int64_t)((g.Cout + TM - 1) / TM) * ((g.Cin + TN - 1) / TN) * g.KH * g.KW;
  int ks = (int)((num_sms() * 4 + tiles - 1) / tiles);
  const int64_t maxks = (M + 255) / 256;
  if (ks > maxks) ks = (int)maxks;
  if (ks < 1) ks = 1;
  if (ks > 64) ks = 64;
  return ks;
}

size_t simt_workspace_bytes(const da_conv_desc* d) {
  const ConvGeom g = make_geom(d);
  const size_t wsz = (size_t)g.Cout * g.KH * g.KW * g.Cin * sizeof(float);
  const int64_t M = (int64_t)g.N * g.OH * g.OW;
  const size_t act = 2 * (size_t)((M + ab_rows(M) - 1) / ab_rows(M)) * g.Cout * sizeof(float);
  const size_t a = wsz * simt_wgrad_ksplit(g);
  return (a > act ? a : act) + 256;
}

int simt_conv_forward(const da_conv_desc* d, const void* x, const void* w, const float* scale,
                      const float* shift, int relu, float drop_p, uint64_t seed, void* y, cudaStream_t st) {
  const ConvGeom g = make_geom(d);
  const int64_t M = (int64_t)g.N * g.OH * g.OW;
  if (thin_ok(g)) {
    const unsigned blocks = (unsigned)((M + 7) / 8);
#define DA_THIN(T, TY) thin_fwd_kernel<T, TY><<<blocks, 256, 0, st>>>(M, g.Cin, g.Cout, (const T*)x, (const T*)w, scale, shift, relu, drop_p, seed, g_seed_counter, (TY*)y)
    if (d->x_dtype == DA_F32 && d->y_dtype == DA_F32) DA_THIN(float, float);
    else if (d->x_dtype == DA_F32 && d->y_dtype == DA_BF16) DA_THIN(float, __nv_bfloat16);
    else if (d->x_dtype == DA_BF16 && d->y_dtype == DA_F32) DA_THIN(__nv_bfloat16, float);
    else if (d->x_dtype == DA_BF16 && d->y_dtype == DA_BF16) DA_THIN(__nv_bfloat16, __nv_bfloat16);
    else DA_REQUIRE(false, DA_ERR_INVALID_ARG, "simt conv forward: bad dtype");
#undef DA_THIN
    DA_LAUNCH_CHECK();
    return DA_OK;
  }
  dim3 grid((unsigned)((M + TM - 1) / TM), (g.Cout + TN - 1) / TN);
#define DA_FWD(T, TY) simt_conv_fwd_kernel<T, TY><<<grid, 256, 0, st>>>(g, (const T*)x, (const T*)w, scale, shift, relu, drop_p, seed, g_seed_counter, (TY*)y)
  if (d->x_dtype == DA_F32 && d->y_dtype == DA_F32) DA_FWD(float, float);
  else if (d->x_dtype == DA_F32 && d->y_dtype == DA_BF16) DA_FWD(float, __nv_bfloat16);
  else if (d->x_dtype == DA_BF16 && d->y_dtype == DA_F32) DA_FWD(__nv_bfloat16, float);
  else if (d->x_dtype == DA_BF16 && d->y_dtype == DA_BF16) DA_FWD(__nv_bfloat16, __nv_bfloat16);
  else DA_REQUIRE(false, DA_ERR_INVALID_ARG, "simt conv forward: bad dtype");
#undef DA_FWD
  DA_LAUNCH_CHECK();
  return DA_OK;
}

int simt_conv_backward_data(const da_conv_desc* d, const void* dz, const void* w, float out_scale, void* dx,
                            cudaStream_t st) {
  const ConvGeom g = make_geom(d);
  const int64_t M = (int64_t)g.N * g.H * g.W;
  if (thin_ok(g)) {
    int64_t blocks = (M * g.Cin + 255) / 256;
    if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
#define DA_THIN(T, TX) thin_dgrad_kernel<T, TX><<<(unsigned)blocks, 256, 0, st>>>(M, g.Cin, g.Cout, (const T*)dz, (const T*)w, out_scale, (TX*)dx)
    if (d->x_dtype == DA_F32 && d->y_dtype == DA_F32) DA_THIN(float, float);
    else if (d->x_dtype == DA_F32 && d->y_dtype == DA_BF16) DA_THIN(float, __nv_bfloat16);
    else if (d->x_dtype == DA_BF16 && d->y_dtype == DA_F32) DA_THIN(__nv_bfloat16, float);
    else if (d->x_dtype == DA_BF16 && d->y_dtype == DA_BF16) DA_THIN(__nv_bfloat16, __nv_bfloat16);
    else DA_REQUIRE(false, DA_ERR_INVALID_ARG, "simt conv dgrad: bad dtype");
#undef DA_THIN
    DA_LAUNCH_CHECK();
    return DA_OK;
  }
  dim3 grid((unsigned)((M + TM - 1) / TM), (g.Cin + TN - 1) / TN);
#define DA_DG(T, TX) simt_conv_dgrad_kernel<T, TX><<<grid, 256, 0, st>>>(g, (const T*)dz, (const T*)w, out_scale, (TX*)dx)
  if (d->x_dtype == DA_F32 && d->y_dtype == DA_F32) DA_DG(float, float);
  else if (d->x_dtype == DA_F32 && d->y_dtype == DA_BF16) DA_DG(float, __nv_bfloat16);
  else if (d->x_dtype == DA_BF16 && d->y_dtype == DA_F32) DA_DG(__nv_bfloat16, float);
  else if (d->x_dtype == DA_BF16 && d->y_dtype == DA_BF16) DA_DG(__nv_bfloat16, __nv_bfloat16);
  else DA_REQUIRE(false, DA_ERR_INVALID_ARG, "simt conv dgrad: bad dtype");
#undef DA_DG
  DA_LAUNCH_CHECK();
  return DA_OK;
}

int simt_conv_backward_weight(const da_conv_desc* d, const void* x, const void* dz, float* dw, void* ws,
                              size_t ws_bytes, cudaStream_t st) {
  const ConvGeom g = make_geom(d);
  const int ks = simt_wgrad_ksplit(g);
  const int taps = g.KH * g.KW;
  const int64_t wn = (int64_t)g.Cout * taps * g.Cin;
  DA_REQUIRE(ks == 1 || (ws && ws_bytes >= (size_t)wn * ks * sizeof(float)), DA_ERR_WORKSPACE, "simt conv wgrad: workspace too small");
  float* part = ks == 1 ? dw : (float*)ws;
  if (thin_ok(g)) {
    const int64_t M = (int64_t)g.N * g.OH * g.OW;
    dim3 tgrid((g.Cin + 31) / 32, ks);
    if (d->x_dtype == DA_F32) thin_wgrad_kernel<float><<<tgrid, 256, 0, st>>>(M, g.Cin, g.Cout, (const float*)x, (const float*)dz, part, ks);
    else if (d->x_dtype == DA_BF16) thin_wgrad_kernel<__nv_bfloat16><<<tgrid, 256, 0, st>>>(M, g.Cin, g.Cout, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dz, part, ks);
    else DA_REQUIRE(false, DA_ERR_INVALID_ARG, "simt conv wgrad: bad dtype");
    DA_LAUNCH_CHECK();
    if (ks > 1) {
      splitk_reduce_kernel<<<(int)((wn + 255) / 256), 256, 0, st>>>(part, ks, wn, dw);
      DA_LAUNCH_CHECK();
    }
    return DA_OK;
  }
  dim3 grid((g.Cout + TM - 1) / TM, (g.Cin + TN - 1) / TN, taps * ks);
  DA_REQUIRE(grid.y <= 65535 && grid.z <= 65535, DA_ERR_UNSUPPORTED, "simt conv wgrad: grid too large");
  if (d->x_dtype == DA_F32) simt_conv_wgrad_kernel<float><<<grid, 256, 0, st>>>(g, (const float*)x, (const float*)dz, part, ks);
  else if (d->x_dtype == DA_BF16) simt_conv_wgrad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(g, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dz, part, ks);
  else DA_REQUIRE(false, DA_ERR_INVALID_ARG, "simt conv wgrad: bad dtype");
  DA_LAUNCH_CHECK();
  if (ks > 1) {
    int64_t blocks = (wn + 255) / 256;
    if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
    splitk_reduce_kernel<<<(int)blocks, 256, 0, st>>>(part, ks, wn, dw);
    DA_LAUNCH_CHECK();
  }
  return DA_OK;
}

}  // namespace da

using namespace da;

extern "C" int da_conv_act_backward(const da_conv_desc* d, const void* dy, const void* y,
                                    const float* scale, int relu, float drop_p, uint64_t drop_seed,
                                    void* dz, float* dshift, float* dvdot, void* workspace,
                                    size_t workspace_bytes, da_stream_t stream) {
  DA_REQUIRE(d && dy && dz, DA_ERR_INVALID_ARG, "conv_act_backward: null argument");
  DA_REQUIRE(!relu || y, DA_ERR_INVALID_ARG, "conv_act_backward: relu needs the forward output y");
  const ConvGeom g = make_geom(d);
  const int64_t M = (int64_t)g.N * g.OH * g.OW;
  const int nb = (int)((M + ab_rows(M) - 1) / ab_rows(M));
  float* partial = nullptr;
  float* partial2 = nullptr;
  DA_REQUIRE(!dvdot || y, DA_ERR_INVALID_ARG, "conv_act_backward: dvdot needs the forward output y");
  if (dshift || dvdot) {
    DA_REQUIRE(workspace && workspace_bytes >= 2 * (size_t)nb * g.Cout * sizeof(float), DA_ERR_WORKSPACE, "conv_act_backward: workspace too small");
    partial = (float*)workspace;
    partial2 = partial + (size_t)nb * g.Cout;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (d->x_dtype == DA_F32) act_bwd_kernel<float><<<nb, 256, 0, st>>>((const float*)dy, (const float*)y, M, g.Cout, scale, relu, drop_p, drop_seed, g_seed_counter, (float*)dz, partial, dvdot ? partial2 : nullptr);
  else if (d->x_dtype == DA_BF16 && act_vec8_ok(g.Cout, dy, y, dz)) {
    const int grp = g.Cout >> 3;
    const int sub = grp <= 128 ? 256 / grp : 1;
    const size_t smem = sub > 1 ? 2 * (size_t)sub * g.Cout * sizeof(float) : 0;     // <= 2 * 256 * 8 * 4 = 16 KB
    act_bwd_vec8_kernel<<<nb, 256, smem, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, M, g.Cout, scale, relu, drop_p, drop_seed,
                                               g_seed_counter, (__nv_bfloat16*)dz, partial, dvdot ? partial2 : nullptr);
  } else if (d->x_dtype == DA_BF16) act_bwd_kernel<__nv_bfloat16><<<nb, 256, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, M, g.Cout, scale, relu, drop_p, drop_seed, g_seed_counter, (__nv_bfloat16*)dz, partial, dvdot ? partial2 : nullptr);
  else DA_REQUIRE(false, DA_ERR_INVALID_ARG, "conv_act_backward: bad dtype");
  DA_LAUNCH_CHECK();
  if (dshift) {
    colsum_final_kernel<<<(g.Cout + 31) / 32, dim3(32, 32), 0, st>>>(partial, nb, g.Cout, dshift);
    DA_LAUNCH_CHECK();
  }
  if (dvdot) {
    colsum_final_kernel<<<(g.Cout + 31) / 32, dim3(32, 32), 0, st>>>(partial2, nb, g.Cout, dvdot);
    DA_LAUNCH_CHECK();
  }
  return DA_OK;
}
