// RoIAlign forward / backward for sm_100a — replaces mmcv.ops.RoIAlign on the DA hot path
// (reference call sites: mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:54-60,
//  single_level_roi_extractor.py:79,103).  Algorithm contract: SURVEY.md Appendix A
// (mmcv-full 1.3.17 roi_align_cuda_kernel.cuh == torchvision roi_align lineage), avg mode.
//
// B200 design (HBM-bound op):
//   * features are NHWC so that a warp reads 32 consecutive channels of one pixel
//     (128 B, coalesced); every footprint pixel of an RoI is read ONCE per channel;
//   * bilinear sampling is separable: out[ph,pw] = 1/count * sum_y sum_x Wy[y,ph] Wx[x,pw] f[y,x]
//     where Wy/Wx are the per-axis sums of the bilinear tap weights of all samples of a
//     bin.  A tiny prep kernel builds (Wy,Wx) and the sampling grid (gh,gw) per RoI with the
//     reference's exact fp32 operation order (no FMA contraction), the pooling kernels
//     stage them in shared memory;
//   * forward: CTA = (RoI, 256-channel block); 49 accumulators per channel live in
//     registers; the [256][49] result tile is transposed through shared memory so the
//     reference layout [R,C,7,7] is written as one contiguous 50 KB run;
//   * backward: gather, no atomics: CTA = (image, 16x32-pixel tile, 32-channel block) keeps
//     its slice of grad_input in shared memory, walks the RoIs overlapping the tile in
//     index order (deterministic), and writes every grad_input element exactly once.
#include "da_common.cuh"
#include "da_ptx.cuh"
#include "roi_common.cuh"
#include <stdlib.h>

namespace da {

// One axis of the reference's sample enumeration.  Thread `bin` (0..6) walks its g samples.
// Mode 0: return (min low index, max high index) of valid samples through lo/hi.
// Mode 1: accumulate tap weights into tab[(idx - base)*8 + bin].
__device__ __forceinline__ void axis_samples(float start, float bin_size, int g, int bin,
                                             int extent, int mode, int base, float* tab,
                                             int& lo, int& hi) {
  const float fext = (float)extent;
  for (int i = 0; i < g; ++i) {
    // reference: start + p*bin + (i + .5f) * bin / g   (left-to-right, no contraction)
    float v = __fadd_rn(__fadd_rn(start, __fmul_rn((float)bin, bin_size)),
                        __fdiv_rn(__fmul_rn((float)i + .5f, bin_size), (float)g));
    if (v < -1.0f || v > fext) continue;  // sample contributes 0
    if (v <= 0.f) v = 0.f;
    int l = (int)v, h;
    if (l >= extent - 1) { h = l = extent - 1; v = (float)l; } else { h = l + 1; }
    const float lw = v - (float)l, hw = 1.f - lw;
    if (mode == 0) {
      lo = min(lo, l);
      hi = max(hi, h);
    } else {
      tab[(l - base) * WROW + bin] += hw;
      tab[(h - base) * WROW + bin] += lw;
    }
  }
}

// grid = ceil(R / PREP_WARPS) blocks of PREP_WARPS warps: one warp per RoI.  want_order: the last block to finish also sorts the
// RoIs by decreasing footprint (work queue of the tensor-core forward).
constexpr int PREP_WARPS = 8;
__global__ void __launch_bounds__(32 * PREP_WARPS)
roi_prep_kernel(const float* __restrict__ rois, int R, int N, int H, int W,
                float spatial_scale, int sampling_ratio, int aligned,
                unsigned char* __restrict__ ws, int32_t* __restrict__ grid_out, int want_order) {
  const int r = blockIdx.x * PREP_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  __shared__ int s_last;
  __shared__ int hist[256];
  if (r < R) {
  int* err = reinterpret_cast<int*>(ws);
  RoiMeta* metas = reinterpret_cast<RoiMeta*>(ws + ws_meta_off());
  float* tab = reinterpret_cast<float*>(ws + ws_table_off(R)) + (size_t)r * (H + W) * WROW;

  const float* roi = rois + (size_t)r * 5;
  const float fb = roi[0];
  int b = (int)fb;
  const float offset = aligned ? 0.5f : 0.f;
  const float x1 = __fsub_rn(__fmul_rn(roi[1], spatial_scale), offset);
  const float y1 = __fsub_rn(__fmul_rn(roi[2], spatial_scale), offset);
  const float x2 = __fsub_rn(__fmul_rn(roi[3], spatial_scale), offset);
  const float y2 = __fsub_rn(__fmul_rn(roi[4], spatial_scale), offset);
  float rw = __fsub_rn(x2, x1), rh = __fsub_rn(y2, y1);
  if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
  const float bh = __fdiv_rn(rh, (float)P), bw = __fdiv_rn(rw, (float)P);
  int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)P));
  int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)P));
  if (!(fb >= 0.f) || b < 0 || b >= N) {  // also catches NaN
    if (lane == 0) atomicExch(&err[0], 1);
    b = -1;
  }
  const int gh_s = max(gh, 0), gw_s = max(gw, 0);  // negative grid == empty loops in the reference

  int lo = 0x7fffffff, hi = -1;
  if (lane < P) axis_samples(x1, bw, gw_s, lane, W, 0, 0, nullptr, lo, hi);
  else if (lane >= 8 && lane < 8 + P) axis_samples(y1, bh, gh_s, lane - 8, H, 0, 0, nullptr, lo, hi);
  const int bin_lo = lo, bin_hi = hi;     // lanes 8..14: first / last map row bin (lane - 8) touches (0x7fffffff / -1: none)
  // segmented (8-lane) min / max
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  const int x_lo = __shfl_sync(0xffffffffu, lo, 0), x_hi = __shfl_sync(0xffffffffu, hi, 0);
  const int y_lo = __shfl_sync(0xffffffffu, lo, 8), y_hi = __shfl_sync(0xffffffffu, hi, 8);
  int nx = (x_hi >= x_lo) ? x_hi - x_lo + 1 : 0;
  int ny = (y_hi >= y_lo) ? y_hi - y_lo + 1 : 0;
  if (nx == 0 || ny == 0 || b < 0) { nx = 0; ny = 0; }

  // tables: wy rows [0,ny) then wx rows at offset H*8
  float* wy = tab;
  float* wx = tab + (size_t)H * WROW;
  for (int i = lane; i < ny * WROW; i += 32) wy[i] = 0.f;
  for (int i = lane; i < nx * WROW; i += 32) wx[i] = 0.f;
  __syncwarp();
  if (nx > 0) {
    int d0, d1;
    if (lane < P) axis_samples(x1, bw, gw_s, lane, W, 1, x_lo, wx, d0, d1);
    else if (lane >= 8 && lane < 8 + P) axis_samples(y1, bh, gh_s, lane - 8, H, 1, y_lo, wy, d0, d1);
  }
  {
    // Pad column of the Wy rows: for map row y, (first bin row reaching y or beyond) | (1 + last bin row starting at y or
    // before) << 8.  A pixel tile covering rows [ya, yb) then needs only the bin rows [first(ya), last(yb - 1)] of this RoI
    // (a superset of the intersecting ones whatever the orientation): the tensor-core backward for [R,7,7,C] gradients
    // fetches and multiplies only those.  No other kernel reads the pad.
    int l7[P], h7[P];
#pragma unroll
    for (int ph = 0; ph < P; ++ph) {
      l7[ph] = __shfl_sync(0xffffffffu, bin_lo, 8 + ph);
      h7[ph] = __shfl_sync(0xffffffffu, bin_hi, 8 + ph);
    }
    for (int i = lane; i < ny; i += 32) {
      const int y = y_lo + i;
      int first = P, last = -1;
#pragma unroll
      for (int ph = P - 1; ph >= 0; --ph) if (h7[ph] >= y) first = ph;
#pragma unroll
      for (int ph = 0; ph < P; ++ph) if (l7[ph] <= y) last = ph;
      wy[i * WROW + P] = __int_as_float(first | ((last + 1) << 8));
    }
  }
  if (lane == 0) {
    RoiMeta m;
    m.b = b; m.gh = gh; m.gw = gw;
    m.y_lo = ny ? y_lo : 0; m.ny = ny; m.x_lo = nx ? x_lo : 0; m.nx = nx;
    m.count = max(gh * gw, 1);
    metas[r] = m;
    if (grid_out) { grid_out[2 * r] = gh; grid_out[2 * r + 1] = gw; }
  }
  }   // r < R
  if (!want_order) return;
  // The LAST block orders the RoIs by decreasing footprint: counting sort over 256 buckets of quad counts (a separate
  // single-block kernel cost a launch and 8 us for this).  The order inside a bucket depends on atomics, the results do not
  // (every RoI's output is independent of when it is computed).
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(reinterpret_cast<int*>(ws) + 2, 1) == (int)gridDim.x - 1;   // hdr[2]: blocks done (zeroed with the header)
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const RoiMeta* all = reinterpret_cast<const RoiMeta*>(ws + ws_meta_off());
  int* order = reinterpret_cast<int*>(ws + ws_order_off(R, H, W));
  const int t = threadIdx.x, nt = blockDim.x;
  for (int k = t; k < 256; k += nt) hist[k] = 0;
  __syncthreads();
  for (int i = t; i < R; i += nt) {
    const RoiMeta m = all[i];
    const int nq = ((m.ny + 3) >> 2) * ((m.nx + 3) >> 2);
    atomicAdd(&hist[255 - min(255, nq)], 1);
  }
  __syncthreads();
  if (t < 32) {   // exclusive prefix: lane l owns buckets [8l, 8l+8)
    int loc[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { loc[j] = sum; sum += hist[8 * t + j]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (t >= o) incl += v;
    }
    const int base = incl - sum;
#pragma unroll
    for (int j = 0; j < 8; ++j) hist[8 * t + j] = base + loc[j];
  }
  __syncthreads();
  for (int i = t; i < R; i += nt) {
    const RoiMeta m = all[i];
    const int nq = ((m.ny + 3) >> 2) * ((m.nx + 3) >> 2);
    order[atomicAdd(&hist[255 - min(255, nq)], 1)] = i;
  }
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
constexpr int FWD_THREADS = 128;
constexpr int FWD_CB = 256;  // channels per CTA (2 per thread)

template <typename TIn> struct FwdMap;
template <> struct FwdMap<float> {  // channels (t, t+128): two coalesced 128 B warp loads per pixel
  __device__ static int chan(int t, int j) { return t + 128 * j; }
  __device__ static int stage(int cl, int k) { return cl * PP + k; }
  __device__ static int stage_linear(int i) { return i; }
};
template <> struct FwdMap<__nv_bfloat16> {  // channels (2t, 2t+1): one packed 4 B load per pixel
  __device__ static int chan(int t, int j) { return 2 * t + j; }
  __device__ static int stage(int cl, int k) { return cl * PP + k + (cl >> 5); }
  __device__ static int stage_linear(int i) { return i + ((i / PP) >> 5); }
};

template <typename TIn>
__device__ __forceinline__ void load2(const TIn* p, bool v0, bool v1, float& f0, float& f1);
template <>
__device__ __forceinline__ void load2<float>(const float* p, bool v0, bool v1, float& f0, float& f1) {
  f0 = v0 ? __ldg(p) : 0.f;
  f1 = v1 ? __ldg(p + 128) : 0.f;
}
template <>
__device__ __forceinline__ void load2<__nv_bfloat16>(const __nv_bfloat16* p, bool v0, bool v1,
                                                     float& f0, float& f1) {
  if (v1) {  // both valid (C even): one 32-bit load
    const unsigned u = __ldg(reinterpret_cast<const unsigned*>(p));
    f0 = __uint_as_float(u << 16);
    f1 = __uint_as_float(u & 0xffff0000u);
  } else {
    f0 = v0 ? __bfloat162float(p[0]) : 0.f;
    f1 = 0.f;
  }
}

template <typename TIn, typename TOut, int kLayout>
__global__ void __launch_bounds__(FWD_THREADS, 3)
roi_align_fwd_kernel(const TIn* __restrict__ feat, int C, int H, int W, int R,
                     const unsigned char* __restrict__ ws, TOut* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  const int r = blockIdx.y;
  const int c0 = blockIdx.x * FWD_CB;
  const int t = threadIdx.x;
  const RoiMeta m = reinterpret_cast<const RoiMeta*>(ws + ws_meta_off())[r];
  const int nch = min(FWD_CB, C - c0);

  // smem: [wy: ny*8][wx: nx*8][stage: 256*49 + 8]
  float* wy_s = smem;
  float* wx_s = smem + (size_t)H * WROW;
  float* stage = smem + (size_t)(H + W) * WROW;
  {
    const float* tab = reinterpret_cast<const float*>(ws + ws_table_off(R)) + (size_t)r * (H + W) * WROW;
    const float4* gy = reinterpret_cast<const float4*>(tab);
    const float4* gx = reinterpret_cast<const float4*>(tab + (size_t)H * WROW);
    float4* sy = reinterpret_cast<float4*>(wy_s);
    float4* sx = reinterpret_cast<float4*>(wx_s);
    for (int i = t; i < m.ny * 2; i += FWD_THREADS) sy[i] = gy[i];
    for (int i = t; i < m.nx * 2; i += FWD_THREADS) sx[i] = gx[i];
  }
  __syncthreads();

  float acc0[PP], acc1[PP];
#pragma unroll
  for (int k = 0; k < PP; ++k) { acc0[k] = 0.f; acc1[k] = 0.f; }

  const int ca = FwdMap<TIn>::chan(t, 0), cb = FwdMap<TIn>::chan(t, 1);
  const bool va = ca < nch, vb = cb < nch;

  if (m.ny > 0 && (va || vb)) {
    const TIn* base = feat + (((size_t)m.b * H + m.y_lo) * W + m.x_lo) * C + c0 + ca;
    for (int ry = 0; ry < m.ny; ++ry) {
      float T0[P], T1[P];
#pragma unroll
      for (int q = 0; q < P; ++q) { T0[q] = 0.f; T1[q] = 0.f; }
      const TIn* prow = base + (size_t)ry * W * C;
#pragma unroll 4
      for (int rx = 0; rx < m.nx; ++rx) {
        float f0, f1;
        load2<TIn>(prow + (size_t)rx * C, va, vb, f0, f1);
        const float4 wa = reinterpret_cast<const float4*>(wx_s)[rx * 2];
        const float4 wb = reinterpret_cast<const float4*>(wx_s)[rx * 2 + 1];
        T0[0] = fmaf(wa.x, f0, T0[0]); T1[0] = fmaf(wa.x, f1, T1[0]);
        T0[1] = fmaf(wa.y, f0, T0[1]); T1[1] = fmaf(wa.y, f1, T1[1]);
        T0[2] = fmaf(wa.z, f0, T0[2]); T1[2] = fmaf(wa.z, f1, T1[2]);
        T0[3] = fmaf(wa.w, f0, T0[3]); T1[3] = fmaf(wa.w, f1, T1[3]);
        T0[4] = fmaf(wb.x, f0, T0[4]); T1[4] = fmaf(wb.x, f1, T1[4]);
        T0[5] = fmaf(wb.y, f0, T0[5]); T1[5] = fmaf(wb.y, f1, T1[5]);
        T0[6] = fmaf(wb.z, f0, T0[6]); T1[6] = fmaf(wb.z, f1, T1[6]);
      }
      const float4 ya = reinterpret_cast<const float4*>(wy_s)[ry * 2];
      const float4 yb = reinterpret_cast<const float4*>(wy_s)[ry * 2 + 1];
      const float wyv[P] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z};
#pragma unroll
      for (int ph = 0; ph < P; ++ph) {
        if (wyv[ph] != 0.f) {  // CTA-uniform: all threads share the RoI
#pragma unroll
          for (int pw = 0; pw < P; ++pw) {
            acc0[ph * P + pw] = fmaf(wyv[ph], T0[pw], acc0[ph * P + pw]);
            acc1[ph * P + pw] = fmaf(wyv[ph], T1[pw], acc1[ph * P + pw]);
          }
        }
      }
    }
  }
  const float cnt = (float)m.count;

  if (kLayout == DA_ROI_OUT_RHWC) {
    // [R,7,7,C]: lanes already map to consecutive channels
    TOut* o = out + (size_t)r * PP * C + c0;
#pragma unroll
    for (int k = 0; k < PP; ++k) {
      if (va) o[(size_t)k * C + ca] = from_f32<TOut>(acc0[k] / cnt);
      if (vb) o[(size_t)k * C + cb] = from_f32<TOut>(acc1[k] / cnt);
    }
  } else {
    // [R,C,7,7]: transpose through smem, then write the contiguous nch*49 run coalesced
#pragma unroll
    for (int k = 0; k < PP; ++k) {
      stage[FwdMap<TIn>::stage(ca, k)] = acc0[k] / cnt;
      stage[FwdMap<TIn>::stage(cb, k)] = acc1[k] / cnt;
    }
    __syncthreads();
    TOut* o = out + ((size_t)r * C + c0) * PP;
    const int total = nch * PP;
    for (int i = t; i < total; i += FWD_THREADS)
      o[i] = from_f32<TOut>(stage[FwdMap<TIn>::stage_linear(i)]);
  }
}

// ---------------------------------------------------------------------------------------
// backward (gather, atomics-free)
// ---------------------------------------------------------------------------------------
constexpr int BWD_THREADS = 256;
constexpr int BWD_TY = 16, BWD_TX = 32;  // pixel tile
constexpr int BWD_CB = 32;               // channels per CTA (one per lane)
constexpr int BWD_LIST = 1024;           // RoI list chunk

template <typename TG, int kLayout>
__global__ void __launch_bounds__(BWD_THREADS, 2)
roi_align_bwd_kernel(const TG* __restrict__ grad_out, int C, int H, int W, int R,
                     const unsigned char* __restrict__ ws, float* __restrict__ grad_in,
                     int tiles_x) {
  extern __shared__ __align__(16) float smem[];
  // smem: acc[TY*TX*32] | g[32*49] | wy[TY*8] | wx[TX*8] | list[BWD_LIST] | misc
  float* acc = smem;
  float* g_s = acc + BWD_TY * BWD_TX * BWD_CB;
  float* wy_s = g_s + BWD_CB * PP + 16;
  float* wx_s = wy_s + BWD_TY * WROW;
  int* list = reinterpret_cast<int*>(wx_s + BWD_TX * WROW);
  __shared__ int s_count;
  __shared__ int s_wcount[BWD_THREADS / 32];

  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * BWD_CB;
  const int ty0 = (blockIdx.x / tiles_x) * BWD_TY, tx0 = (blockIdx.x % tiles_x) * BWD_TX;
  const int ty1 = min(ty0 + BWD_TY, H), tx1 = min(tx0 + BWD_TX, W);
  const RoiMeta* metas = reinterpret_cast<const RoiMeta*>(ws + ws_meta_off());
  const float* tables = reinterpret_cast<const float*>(ws + ws_table_off(R));
  const bool cvalid = (c0 + lane) < C;

  for (int i = t; i < BWD_TY * BWD_TX * BWD_CB; i += BWD_THREADS) acc[i] = 0.f;

  for (int rbase = 0; rbase < R; rbase += BWD_LIST) {
    // ---- ordered compaction of the RoIs of image b that touch this tile
    if (t == 0) s_count = 0;
    __syncthreads();
    const int rend = min(rbase + BWD_LIST, R);
    for (int r0 = rbase; r0 < rend; r0 += BWD_THREADS) {
      const int r = r0 + t;
      bool hit = false;
      if (r < rend) {
        const RoiMeta m = metas[r];
        hit = (m.b == b) && m.ny > 0 && m.y_lo < ty1 && m.y_lo + m.ny > ty0 &&
              m.x_lo < tx1 && m.x_lo + m.nx > tx0;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) s_wcount[wid] = __popc(bal);
      __syncthreads();
      int off = s_count;
      for (int w = 0; w < wid; ++w) off += s_wcount[w];
      if (hit) list[off + __popc(bal & ((1u << lane) - 1u))] = r;
      __syncthreads();
      if (t == 0) {
        int tot = 0;
        for (int w = 0; w < BWD_THREADS / 32; ++w) tot += s_wcount[w];
        s_count += tot;
      }
      __syncthreads();
    }
    const int n_list = s_count;

    // ---- walk the list in RoI order
    for (int li = 0; li < n_list; ++li) {
      const int r = list[li];
      const RoiMeta m = metas[r];
      const int ya = max(m.y_lo, ty0), yb = min(m.y_lo + m.ny, ty1);  // rows [ya,yb)
      const int xa = max(m.x_lo, tx0), xb = min(m.x_lo + m.nx, tx1);
      // stage grad_out[r, c0:c0+32, 49] and the table slices
      if (kLayout == DA_ROI_OUT_RCHW) {
        const TG* g = grad_out + ((size_t)r * C + c0) * PP;
        const int total = min(BWD_CB, C - c0) * PP;
        for (int i = t; i < BWD_CB * PP; i += BWD_THREADS) g_s[i] = (i < total) ? to_f32<TG>(g[i]) : 0.f;
      } else {
        const TG* g = grad_out + (size_t)r * PP * C + c0;
        for (int i = t; i < BWD_CB * PP; i += BWD_THREADS) {
          const int k = i >> 5, cl = i & 31;
          g_s[cl * PP + k] = (c0 + cl < C) ? to_f32<TG>(g[(size_t)k * C + cl]) : 0.f;
        }
      }
      const float* tab = tables + (size_t)r * (H + W) * WROW;
      const float inv_count = 1.f / (float)m.count;
      for (int i = t; i < (yb - ya) * WROW; i += BWD_THREADS)
        wy_s[i] = tab[(size_t)(ya - m.y_lo) * WROW + i] * inv_count;
      for (int i = t; i < (xb - xa) * WROW; i += BWD_THREADS)
        wx_s[i] = tab[(size_t)H * WROW + (size_t)(xa - m.x_lo) * WROW + i];
      __syncthreads();

      // warp `wid` takes rows ya+wid, ya+wid+8, ...; lane = channel
      for (int y = ya + wid; y < yb; y += BWD_THREADS / 32) {
        float U[P];
#pragma unroll
        for (int q = 0; q < P; ++q) U[q] = 0.f;
        const float4 wa = reinterpret_cast<const float4*>(wy_s)[(y - ya) * 2];
        const float4 wb = reinterpret_cast<const float4*>(wy_s)[(y - ya) * 2 + 1];
        const float wyv[P] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z};
#pragma unroll
        for (int ph = 0; ph < P; ++ph) {
          if (wyv[ph] != 0.f) {
#pragma unroll
            for (int pw = 0; pw < P; ++pw)
              U[pw] = fmaf(wyv[ph], g_s[lane * PP + ph * P + pw], U[pw]);
          }
        }
        float* arow = acc + ((size_t)(y - ty0) * BWD_TX + (xa - tx0)) * BWD_CB + lane;
        for (int x = xa; x < xb; ++x) {
          const float4 xa4 = reinterpret_cast<const float4*>(wx_s)[(x - xa) * 2];
          const float4 xb4 = reinterpret_cast<const float4*>(wx_s)[(x - xa) * 2 + 1];
          float v = xa4.x * U[0];
          v = fmaf(xa4.y, U[1], v); v = fmaf(xa4.z, U[2], v); v = fmaf(xa4.w, U[3], v);
          v = fmaf(xb4.x, U[4], v); v = fmaf(xb4.y, U[5], v); v = fmaf(xb4.z, U[6], v);
          arow[(x - xa) * BWD_CB] += v;
        }
      }
      __syncthreads();
    }
  }

  // ---- write the tile once (NHWC: 32 channels = 128 B per pixel)
  for (int i = t; i < BWD_TY * BWD_TX * BWD_CB; i += BWD_THREADS) {
    const int cl = i & 31, p = i >> 5;
    const int y = ty0 + p / BWD_TX, x = tx0 + p % BWD_TX;
    if (y < H && x < W && c0 + cl < C)
      grad_in[(((size_t)b * H + y) * W + x) * C + c0 + cl] = acc[i];
  }
  (void)cvalid;
}


// =======================================================================================
// Bulk-async (TMA engine) pipelined variants — the production path.
//
// The synchronous kernels above keep at most a few loads per thread in flight and measured
// ~5% of HBM peak on B200 (latency bound).  Here one producer warp streams the RoI's
// footprint (forward) / the per-RoI gradient chunks (backward) into a shared-memory ring with
// cp.async.bulk + mbarrier complete_tx, so tens of KB per SM are in flight independent of the
// consumers' register budget, and the forward result tile leaves through one bulk store.
// =======================================================================================
constexpr int FA_CONSUMERS = 128;
constexpr int FA_THREADS = FA_CONSUMERS + 32;
constexpr int FA_SEG = 8;  // footprint pixels per ring slot
template <typename TIn> __host__ __device__ constexpr int fa_slots() { return sizeof(TIn) == 2 ? 8 : 6; }

template <typename TIn, typename TOut>
__host__ __device__ constexpr size_t fa_smem_bytes(int H, int W) {
  return 128 + (size_t)fa_slots<TIn>() * FA_SEG * FWD_CB * sizeof(TIn) + (size_t)FWD_CB * PP * sizeof(TOut) + 64 +
         (size_t)(H + W) * WROW * sizeof(float) + 2 * 8 * 8;
}

template <typename TIn, typename TOut, int kLayout>
__global__ void __launch_bounds__(FA_THREADS, 2)
roi_align_fwd_async_kernel(const TIn* __restrict__ feat, int C, int H, int W, int R,
                           const unsigned char* __restrict__ ws, TOut* __restrict__ out, int skip_tc) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int NSLOT = fa_slots<TIn>();
  constexpr int SLOT_BYTES = FA_SEG * FWD_CB * (int)sizeof(TIn);
  uint8_t* sm = smem_raw + ((128 - (smem_u32(smem_raw) & 127)) & 127);
  TIn* ring = reinterpret_cast<TIn*>(sm);
  TOut* stage = reinterpret_cast<TOut*>(sm + NSLOT * SLOT_BYTES);
  float* wy_s = reinterpret_cast<float*>(sm + NSLOT * SLOT_BYTES + ((FWD_CB * PP * sizeof(TOut) + 63) / 64) * 64);
  float* wx_s = wy_s + (size_t)H * WROW;
  const uint32_t bars = smem_u32(wx_s + (size_t)W * WROW);
  const uint32_t full0 = bars, empty0 = bars + 8 * NSLOT;

  const int r = blockIdx.y;
  const int c0 = blockIdx.x * FWD_CB;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const RoiMeta m = reinterpret_cast<const RoiMeta*>(ws + ws_meta_off())[r];
  const int nch = min(FWD_CB, C - c0);
  if (skip_tc) {   // this RoI is produced by the tensor-core kernel (roi_align_tc.cu)
    if (roi_tc_eligible(m)) return;
  }

  if (t == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, FA_CONSUMERS / 32); }
    fence_barrier_init();
  }
  {
    const float* tab = reinterpret_cast<const float*>(ws + ws_table_off(R)) + (size_t)r * (H + W) * WROW;
    const float4* gy = reinterpret_cast<const float4*>(tab);
    const float4* gx = reinterpret_cast<const float4*>(tab + (size_t)H * WROW);
    for (int i = t; i < m.ny * 2; i += FA_THREADS) reinterpret_cast<float4*>(wy_s)[i] = gy[i];
    for (int i = t; i < m.nx * 2; i += FA_THREADS) reinterpret_cast<float4*>(wx_s)[i] = gx[i];
  }
  __syncthreads();

  const int nseg_row = (m.nx + FA_SEG - 1) / FA_SEG;

  if (warp == FA_CONSUMERS / 32) {
    // ------------------------------ producer warp
    const uint32_t row_bytes = (uint32_t)nch * sizeof(TIn);
    int s = 0;
    for (int ry = 0; ry < m.ny; ++ry) {
      const TIn* prow = feat + (((size_t)m.b * H + m.y_lo + ry) * W + m.x_lo) * C + c0;
      for (int sx = 0; sx < nseg_row; ++sx, ++s) {
        const int slot = s % NSLOT;
        const uint32_t par = (uint32_t)(s / NSLOT) & 1u;
        mbar_wait(empty0 + 8 * slot, par ^ 1u);
        const int xs = sx * FA_SEG;
        const int len = min(FA_SEG, m.nx - xs);
        if (lane == 0) mbar_expect_tx(full0 + 8 * slot, (uint32_t)len * row_bytes);
        __syncwarp();
        if (lane < len)
          bulk_g2s(smem_u32(ring) + slot * SLOT_BYTES + lane * FWD_CB * (int)sizeof(TIn),
                   prow + (size_t)(xs + lane) * C, row_bytes, full0 + 8 * slot);
      }
    }
    return;
  }

  // -------------------------------- consumer warps: channels t and t+128
  float acc0[PP], acc1[PP];
#pragma unroll
  for (int k = 0; k < PP; ++k) { acc0[k] = 0.f; acc1[k] = 0.f; }
  int s = 0;
  for (int ry = 0; ry < m.ny; ++ry) {
    float T0[P], T1[P];
#pragma unroll
    for (int q = 0; q < P; ++q) { T0[q] = 0.f; T1[q] = 0.f; }
    for (int sx = 0; sx < nseg_row; ++sx, ++s) {
      const int slot = s % NSLOT;
      const uint32_t par = (uint32_t)(s / NSLOT) & 1u;
      mbar_wait(full0 + 8 * slot, par);
      const int xs = sx * FA_SEG;
      const int len = min(FA_SEG, m.nx - xs);
      const TIn* px = ring + (size_t)slot * FA_SEG * FWD_CB + t;
#pragma unroll 4
      for (int i = 0; i < len; ++i) {
        const float f0 = to_f32<TIn>(px[i * FWD_CB]);
        const float f1 = to_f32<TIn>(px[i * FWD_CB + 128]);
        const float4 wa = reinterpret_cast<const float4*>(wx_s)[(xs + i) * 2];
        const float4 wb = reinterpret_cast<const float4*>(wx_s)[(xs + i) * 2 + 1];
        T0[0] = fmaf(wa.x, f0, T0[0]); T1[0] = fmaf(wa.x, f1, T1[0]);
        T0[1] = fmaf(wa.y, f0, T0[1]); T1[1] = fmaf(wa.y, f1, T1[1]);
        T0[2] = fmaf(wa.z, f0, T0[2]); T1[2] = fmaf(wa.z, f1, T1[2]);
        T0[3] = fmaf(wa.w, f0, T0[3]); T1[3] = fmaf(wa.w, f1, T1[3]);
        T0[4] = fmaf(wb.x, f0, T0[4]); T1[4] = fmaf(wb.x, f1, T1[4]);
        T0[5] = fmaf(wb.y, f0, T0[5]); T1[5] = fmaf(wb.y, f1, T1[5]);
        T0[6] = fmaf(wb.z, f0, T0[6]); T1[6] = fmaf(wb.z, f1, T1[6]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * slot);
    }
    const float4 ya = reinterpret_cast<const float4*>(wy_s)[ry * 2];
    const float4 yb = reinterpret_cast<const float4*>(wy_s)[ry * 2 + 1];
    const float wyv[P] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z};
#pragma unroll
    for (int ph = 0; ph < P; ++ph) {
      if (wyv[ph] != 0.f) {
#pragma unroll
        for (int pw = 0; pw < P; ++pw) {
          acc0[ph * P + pw] = fmaf(wyv[ph], T0[pw], acc0[ph * P + pw]);
          acc1[ph * P + pw] = fmaf(wyv[ph], T1[pw], acc1[ph * P + pw]);
        }
      }
    }
  }
  const float cnt = (float)m.count;
  const bool va = t < nch, vb = t + 128 < nch;
  if (kLayout == DA_ROI_OUT_RHWC) {
    TOut* o = out + (size_t)r * PP * C + c0;
#pragma unroll
    for (int k = 0; k < PP; ++k) {
      if (va) o[(size_t)k * C + t] = from_f32<TOut>(acc0[k] / cnt);
      if (vb) o[(size_t)k * C + t + 128] = from_f32<TOut>(acc1[k] / cnt);
    }
  } else {
#pragma unroll
    for (int k = 0; k < PP; ++k) {
      stage[t * PP + k] = from_f32<TOut>(acc0[k] / cnt);
      stage[(t + 128) * PP + k] = from_f32<TOut>(acc1[k] / cnt);
    }
    fence_proxy_async();                 // generic-proxy smem writes -> visible to the bulk engine
    named_bar_sync(1, FA_CONSUMERS);
    if (t == 0) {
      bulk_s2g(out + ((size_t)r * C + c0) * PP, smem_u32(stage), (uint32_t)(nch * PP * sizeof(TOut)));
      bulk_commit();
      bulk_wait_read0();
    }
  }
}

// ---------------------------------------------------------------------------------------
constexpr int BA_CWARPS = 8;
constexpr int BA_THREADS = (BA_CWARPS + 1) * 32;
constexpr int BA_SLOTS = 4;
constexpr int BA_HDR = 32;  // bytes: ya, yb, xa, xb, inv_count
template <typename TG> __host__ __device__ constexpr int ba_slot_bytes() {
  return BA_HDR + ((BWD_CB * PP * (int)sizeof(TG) + 127) / 128) * 128 + (BWD_TY + BWD_TX) * WROW * 4;
}
template <typename TG> __host__ __device__ constexpr size_t ba_smem_bytes() {
  return 128 + (size_t)BWD_TY * BWD_TX * BWD_CB * 4 + (size_t)BA_SLOTS * ba_slot_bytes<TG>() + BWD_LIST * 4 + 2 * 8 * BA_SLOTS + 64;
}

template <typename TG, int kLayout>
__global__ void __launch_bounds__(BA_THREADS, 2)
roi_align_bwd_async_kernel(const TG* __restrict__ grad_out, int C, int H, int W, int R,
                           const unsigned char* __restrict__ ws, float* __restrict__ grad_in, int tiles_x) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int SLOT = ba_slot_bytes<TG>();
  constexpr int G_BYTES = ((BWD_CB * PP * (int)sizeof(TG) + 127) / 128) * 128;
  uint8_t* sm = smem_raw + ((128 - (smem_u32(smem_raw) & 127)) & 127);
  float* acc = reinterpret_cast<float*>(sm);
  uint8_t* slots = sm + (size_t)BWD_TY * BWD_TX * BWD_CB * 4;
  int* list = reinterpret_cast<int*>(slots + BA_SLOTS * SLOT);
  const uint32_t bars = smem_u32(list + BWD_LIST);
  const uint32_t full0 = bars, empty0 = bars + 8 * BA_SLOTS;
  __shared__ int s_count;
  __shared__ int s_wcount[BA_THREADS / 32];

  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * BWD_CB;
  const int nch = min(BWD_CB, C - c0);
  const int ty0 = (blockIdx.x / tiles_x) * BWD_TY, tx0 = (blockIdx.x % tiles_x) * BWD_TX;
  const int ty1 = min(ty0 + BWD_TY, H), tx1 = min(tx0 + BWD_TX, W);
  const RoiMeta* metas = reinterpret_cast<const RoiMeta*>(ws + ws_meta_off());
  const float* tables = reinterpret_cast<const float*>(ws + ws_table_off(R));

  if (t == 0) {
    for (int i = 0; i < BA_SLOTS; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, BA_CWARPS); }
    fence_barrier_init();
  }
  for (int i = t; i < BWD_TY * BWD_TX * BWD_CB; i += BA_THREADS) acc[i] = 0.f;

  int seq = 0;  // running slot sequence number (same in every thread)
  for (int rbase = 0; rbase < R; rbase += BWD_LIST) {
    if (t == 0) s_count = 0;
    __syncthreads();
    const int rend = min(rbase + BWD_LIST, R);
    for (int r0 = rbase; r0 < rend; r0 += BA_THREADS) {
      const int r = r0 + t;
      bool hit = false;
      if (r < rend) {
        const RoiMeta m = metas[r];
        hit = (m.b == b) && m.ny > 0 && m.y_lo < ty1 && m.y_lo + m.ny > ty0 && m.x_lo < tx1 && m.x_lo + m.nx > tx0;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) s_wcount[wid] = __popc(bal);
      __syncthreads();
      int off = s_count;
      for (int w = 0; w < wid; ++w) off += s_wcount[w];
      if (hit) list[off + __popc(bal & ((1u << lane) - 1u))] = r;
      __syncthreads();
      if (t == 0) {
        int tot = 0;
        for (int w = 0; w < BA_THREADS / 32; ++w) tot += s_wcount[w];
        s_count += tot;
      }
      __syncthreads();
    }
    const int n_list = s_count;

    if (wid == BA_CWARPS) {
      // ------------------------------ producer warp
      for (int li = 0; li < n_list; ++li) {
        const int sq = seq + li, slot = sq % BA_SLOTS;
        const uint32_t par = (uint32_t)(sq / BA_SLOTS) & 1u;
        mbar_wait(empty0 + 8 * slot, par ^ 1u);
        const int r = list[li];
        const RoiMeta m = metas[r];
        const int ya = max(m.y_lo, ty0), yb = min(m.y_lo + m.ny, ty1);
        const int xa = max(m.x_lo, tx0), xb = min(m.x_lo + m.nx, tx1);
        uint8_t* sl = slots + slot * SLOT;
        const uint32_t g_bytes = (uint32_t)(nch * PP * sizeof(TG));
        if (lane == 0) {
          int* hdr = reinterpret_cast<int*>(sl);
          hdr[0] = ya; hdr[1] = yb; hdr[2] = xa; hdr[3] = xb;
          reinterpret_cast<float*>(sl)[4] = 1.f / (float)m.count;
          mbar_expect_tx(full0 + 8 * slot, g_bytes + (uint32_t)((yb - ya) + (xb - xa)) * WROW * 4);
        }
        __syncwarp();
        const float* tab = tables + (size_t)r * (H + W) * WROW;
        const uint32_t sbase = smem_u32(sl) + BA_HDR;
        if (kLayout == DA_ROI_OUT_RCHW) {
          if (lane == 0) bulk_g2s(sbase, grad_out + ((size_t)r * C + c0) * PP, g_bytes, full0 + 8 * slot);
        } else {
          const uint32_t seg = (uint32_t)(nch * sizeof(TG));
          for (int k = lane; k < PP; k += 32)
            bulk_g2s(sbase + k * BWD_CB * (int)sizeof(TG), grad_out + ((size_t)r * PP + k) * C + c0, seg, full0 + 8 * slot);
        }
        if (lane == 1) bulk_g2s(sbase + G_BYTES, tab + (size_t)(ya - m.y_lo) * WROW, (uint32_t)(yb - ya) * WROW * 4, full0 + 8 * slot);
        if (lane == 2) bulk_g2s(sbase + G_BYTES + BWD_TY * WROW * 4, tab + (size_t)H * WROW + (size_t)(xa - m.x_lo) * WROW,
                                (uint32_t)(xb - xa) * WROW * 4, full0 + 8 * slot);
      }
    } else {
      // ------------------------------ consumer warps: static pixel ownership
      // warp w owns rows of parity (w>>2) and the 8-column strip (w&3) of the tile -> no two warps ever
      // touch the same accumulator, RoIs are applied in list order (deterministic)
      const int rpar = wid >> 2, strip = wid & 3;
      const int sx0 = tx0 + strip * 8, sx1 = sx0 + 8;
      for (int li = 0; li < n_list; ++li) {
        const int sq = seq + li, slot = sq % BA_SLOTS;
        const uint32_t par = (uint32_t)(sq / BA_SLOTS) & 1u;
        mbar_wait(full0 + 8 * slot, par);
        const uint8_t* sl = slots + slot * SLOT;
        const int* hdr = reinterpret_cast<const int*>(sl);
        const int ya = hdr[0], yb = hdr[1], xa = hdr[2], xb = hdr[3];
        const float inv_count = reinterpret_cast<const float*>(sl)[4];
        const TG* g_s = reinterpret_cast<const TG*>(sl + BA_HDR);
        const float* wy_s = reinterpret_cast<const float*>(sl + BA_HDR + G_BYTES);
        const float* wx_s = wy_s + BWD_TY * WROW;
        const int cxa = max(xa, sx0), cxb = min(xb, sx1);
        if (cxa < cxb) {
          int y = ya + (((ya - ty0) & 1) != rpar ? 1 : 0);
          for (; y < yb; y += 2) {
            const float4 wa = reinterpret_cast<const float4*>(wy_s)[(y - ya) * 2];
            const float4 wb = reinterpret_cast<const float4*>(wy_s)[(y - ya) * 2 + 1];
            const float wyv[P] = {wa.x * inv_count, wa.y * inv_count, wa.z * inv_count, wa.w * inv_count,
                                  wb.x * inv_count, wb.y * inv_count, wb.z * inv_count};
            float U[P];
#pragma unroll
            for (int q = 0; q < P; ++q) U[q] = 0.f;
#pragma unroll
            for (int ph = 0; ph < P; ++ph) {
              if (wyv[ph] != 0.f) {
#pragma unroll
                for (int pw = 0; pw < P; ++pw) {
                  const float gv = (kLayout == DA_ROI_OUT_RCHW) ? to_f32<TG>(g_s[lane * PP + ph * P + pw])
                                                                : to_f32<TG>(g_s[(ph * P + pw) * BWD_CB + lane]);
                  U[pw] = fmaf(wyv[ph], gv, U[pw]);
                }
              }
            }
            float* arow = acc + ((size_t)(y - ty0) * BWD_TX + (cxa - tx0)) * BWD_CB + lane;
            for (int x = cxa; x < cxb; ++x) {
              const float4 xa4 = reinterpret_cast<const float4*>(wx_s)[(x - xa) * 2];
              const float4 xb4 = reinterpret_cast<const float4*>(wx_s)[(x - xa) * 2 + 1];
              float v = xa4.x * U[0];
              v = fmaf(xa4.y, U[1], v); v = fmaf(xa4.z, U[2], v); v = fmaf(xa4.w, U[3], v);
              v = fmaf(xb4.x, U[4], v); v = fmaf(xb4.y, U[5], v); v = fmaf(xb4.z, U[6], v);
              arow[(x - cxa) * BWD_CB] += v;
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * slot);
      }
    }
    seq += n_list;
  }
  __syncthreads();
  for (int i = t; i < BWD_TY * BWD_TX * BWD_CB; i += BA_THREADS) {
    const int cl = i & 31, p = i >> 5;
    const int y = ty0 + p / BWD_TX, x = tx0 + p % BWD_TX;
    if (y < H && x < W && cl < nch) grad_in[(((size_t)b * H + y) * W + x) * C + c0 + cl] = acc[i];
  }
}

__global__ void map_roi_levels_kernel(const float* __restrict__ rois, int R, int num_levels,
                                      float finest_scale, int32_t* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float* q = rois + (size_t)r * 5;
  // scale = sqrt((x2-x1)*(y2-y1)); lvl = floor(log2(scale/finest + 1e-6)) clamped
  const float s = sqrtf(__fmul_rn(__fsub_rn(q[3], q[1]), __fsub_rn(q[4], q[2])));
  float l = floorf(log2f(__fadd_rn(__fdiv_rn(s, finest_scale), 1e-6f)));
  // NaN (negative area) -> clamp(min) in torch yields NaN -> .long() is implementation
  // defined; we map it to level 0.
  int lv = (l != l) ? 0 : (int)fminf(fmaxf(l, 0.f), (float)(num_levels - 1));
  out[r] = lv;
}

}  // namespace da

using namespace da;

extern "C" size_t da_roi_align_workspace_bytes(int R, int H, int W) {
  if (R < 0) R = 0;
  return ws_order_off(R, H, W) + (size_t)R * sizeof(int) + 256;
}

static int check_common(int N, int C, int H, int W, int R, int ph, int pw, const void* rois,
                        const void* ws, size_t ws_bytes) {
  DA_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, DA_ERR_INVALID_ARG, "roi_align: bad feature shape [%d,%d,%d,%d]", N, C, H, W);
  DA_REQUIRE(R >= 0, DA_ERR_INVALID_ARG, "roi_align: negative RoI count %d", R);
  DA_REQUIRE(ph == P && pw == P, DA_ERR_UNSUPPORTED, "roi_align: only output_size=7 is built (got %dx%d)", ph, pw);
  DA_REQUIRE(R == 0 || rois != nullptr, DA_ERR_INVALID_ARG, "roi_align: rois is null");
  DA_REQUIRE(R == 0 || ws != nullptr, DA_ERR_WORKSPACE, "roi_align: workspace is null");
  DA_REQUIRE(ws_bytes >= da_roi_align_workspace_bytes(R, H, W), DA_ERR_WORKSPACE,
             "roi_align: workspace too small (%zu < %zu)", ws_bytes, da_roi_align_workspace_bytes(R, H, W));
  DA_REQUIRE((size_t)(H + W) * WROW * 4 + (FWD_CB * PP + 8) * 4 <= 220 * 1024, DA_ERR_UNSUPPORTED,
             "roi_align: H+W=%d too large for the shared-memory weight tables", H + W);
  return DA_OK;
}

static int run_prep(const float* rois, int R, int N, int H, int W, float scale, int sr, int aligned,
                    void* ws, int32_t* grid_out, cudaStream_t st, int want_order = 0) {
  DA_CUDA_OK(cudaMemsetAsync(ws, 0, 16, st));
  roi_prep_kernel<<<(R + PREP_WARPS - 1) / PREP_WARPS, 32 * PREP_WARPS, 0, st>>>(rois, R, N, H, W, scale, sr, aligned, (unsigned char*)ws,
                                                                                grid_out, want_order);
  DA_LAUNCH_CHECK();
  return DA_OK;
}

template <typename TIn, typename TOut>
static int launch_fwd(const void* feat, int N, int C, int H, int W, int R, const void* ws, void* out,
                      int layout, cudaStream_t st) {
  dim3 grid((C + FWD_CB - 1) / FWD_CB, R);
  // tensor-core path: bf16 features, reference layout, 16-byte granular channel runs
  if (sizeof(TIn) == 2 && C % 64 == 0 && ((uintptr_t)feat & 15) == 0 &&
      ((uintptr_t)out & 15) == 0 && !g_opt.roi_no_tc)
    return roi_align_fwd_tc(feat, N, C, H, W, R, ws, out, sizeof(TOut) == 2 ? DA_BF16 : DA_F32, layout, st);
  const int skip_tc = 0;
  // bulk-async path: every per-pixel channel run and the result tile must be 16-byte granular
  const bool async_ok = ((size_t)C * sizeof(TIn)) % 16 == 0 && ((uintptr_t)feat & 15) == 0 &&
                        (layout == DA_ROI_OUT_RHWC || (((size_t)C * sizeof(TOut)) % 16 == 0 && ((uintptr_t)out & 15) == 0)) &&
                        fa_smem_bytes<TIn, TOut>(H, W) <= 110 * 1024;
  if (async_ok) {
    const size_t smem = fa_smem_bytes<TIn, TOut>(H, W);
    if (layout == DA_ROI_OUT_RCHW) {
      auto k = roi_align_fwd_async_kernel<TIn, TOut, DA_ROI_OUT_RCHW>;
      DA_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, FA_THREADS, smem, st>>>((const TIn*)feat, C, H, W, R, (const unsigned char*)ws, (TOut*)out, skip_tc);
    } else {
      auto k = roi_align_fwd_async_kernel<TIn, TOut, DA_ROI_OUT_RHWC>;
      DA_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, FA_THREADS, smem, st>>>((const TIn*)feat, C, H, W, R, (const unsigned char*)ws, (TOut*)out, 0);
    }
    DA_LAUNCH_CHECK();
    return DA_OK;
  }
  const size_t smem = ((size_t)(H + W) * WROW + FWD_CB * PP + 8) * sizeof(float);
  if (layout == DA_ROI_OUT_RCHW) {
    auto k = roi_align_fwd_kernel<TIn, TOut, DA_ROI_OUT_RCHW>;
    DA_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, FWD_THREADS, smem, st>>>((const TIn*)feat, C, H, W, R, (const unsigned char*)ws, (TOut*)out);
  } else {
    auto k = roi_align_fwd_kernel<TIn, TOut, DA_ROI_OUT_RHWC>;
    DA_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, FWD_THREADS, smem, st>>>((const TIn*)feat, C, H, W, R, (const unsigned char*)ws, (TOut*)out);
  }
  DA_LAUNCH_CHECK();
  return DA_OK;
}

extern "C" int da_roi_align_forward(const void* feat, int feat_dtype, int N, int C, int H, int W,
                                    const float* rois, int R, int pooled_h, int pooled_w,
                                    float spatial_scale, int sampling_ratio, int aligned,
                                    void* out, int out_dtype, int out_layout, int32_t* grid_out,
                                    void* workspace, size_t workspace_bytes, da_stream_t stream) {
  int rc = check_common(N, C, H, W, R, pooled_h, pooled_w, rois, workspace, workspace_bytes);
  if (rc) return rc;
  DA_REQUIRE(out_layout == DA_ROI_OUT_RCHW || out_layout == DA_ROI_OUT_RHWC, DA_ERR_INVALID_ARG, "roi_align: bad out_layout %d", out_layout);
  if (R == 0) return DA_OK;  // single_level_roi_extractor.py:77-78: empty RoI set -> empty output
  DA_REQUIRE(feat && out, DA_ERR_INVALID_ARG, "roi_align: null feature/output pointer");
  DA_REQUIRE(feat_dtype != DA_BF16 || (C % 2 == 0), DA_ERR_UNSUPPORTED, "roi_align: bf16 features need even C");
  cudaStream_t st = (cudaStream_t)stream;
  // the footprint order is only read by the tensor-core forward (launch_fwd's first branch)
  const int want_order = feat_dtype == DA_BF16 && C % 64 == 0 && !g_opt.roi_no_tc;
  rc = run_prep(rois, R, N, H, W, spatial_scale, sampling_ratio, aligned, workspace, grid_out, st, want_order);
  if (rc) return rc;
  if (feat_dtype == DA_F32 && out_dtype == DA_F32) return launch_fwd<float, float>(feat, N, C, H, W, R, workspace, out, out_layout, st);
  if (feat_dtype == DA_F32 && out_dtype == DA_BF16) return launch_fwd<float, __nv_bfloat16>(feat, N, C, H, W, R, workspace, out, out_layout, st);
  if (feat_dtype == DA_BF16 && out_dtype == DA_F32) return launch_fwd<__nv_bfloat16, float>(feat, N, C, H, W, R, workspace, out, out_layout, st);
  if (feat_dtype == DA_BF16 && out_dtype == DA_BF16) return launch_fwd<__nv_bfloat16, __nv_bfloat16>(feat, N, C, H, W, R, workspace, out, out_layout, st);
  DA_REQUIRE(false, DA_ERR_INVALID_ARG, "roi_align: bad dtype %d/%d", feat_dtype, out_dtype);
}

template <typename TG>
static int launch_bwd(const void* g, int layout, int N, int C, int H, int W, int R, const void* ws,
                      void* gin_v, int gin_dtype, cudaStream_t st) {
  const int tiles_x = (W + BWD_TX - 1) / BWD_TX, tiles_y = (H + BWD_TY - 1) / BWD_TY;
  dim3 grid(tiles_x * tiles_y, (C + BWD_CB - 1) / BWD_CB, N);
  DA_REQUIRE(grid.y <= 65535 && grid.z <= 65535, DA_ERR_UNSUPPORTED, "roi_align_backward: grid too large");
  // tensor-core path: bf16 gradients in the reference layout, 16-byte granular channel runs
  // ([R,7,7,C] gradients: the operand is fetched by TMA as it lies in memory, C % 64 == 0)
  if (sizeof(TG) == 2 && C % (layout == DA_ROI_OUT_RCHW ? 8 : 64) == 0 && ((uintptr_t)g & 15) == 0 && !g_opt.roi_no_tc)
    return roi_align_bwd_tc(g, layout, N, C, H, W, R, ws, gin_v, gin_dtype, st);
  DA_REQUIRE(gin_dtype == DA_F32, DA_ERR_UNSUPPORTED,
             "roi_align_backward: bf16 grad_input needs the tensor-core path (bf16 [R,C,7,7] gradients, C %% 8 == 0)");
  float* gin = static_cast<float*>(gin_v);
  const bool async_ok = ((size_t)C * sizeof(TG)) % 16 == 0 && ((uintptr_t)g & 15) == 0;
  if (async_ok) {
    const size_t smem = ba_smem_bytes<TG>();
    if (layout == DA_ROI_OUT_RCHW) {
      auto k = roi_align_bwd_async_kernel<TG, DA_ROI_OUT_RCHW>;
      DA_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, BA_THREADS, smem, st>>>((const TG*)g, C, H, W, R, (const unsigned char*)ws, gin, tiles_x);
    } else {
      auto k = roi_align_bwd_async_kernel<TG, DA_ROI_OUT_RHWC>;
      DA_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, BA_THREADS, smem, st>>>((const TG*)g, C, H, W, R, (const unsigned char*)ws, gin, tiles_x);
    }
    DA_LAUNCH_CHECK();
    return DA_OK;
  }
  const size_t smem = ((size_t)BWD_TY * BWD_TX * BWD_CB + BWD_CB * PP + 16 + (BWD_TY + BWD_TX) * WROW) * sizeof(float) +
                      BWD_LIST * sizeof(int);
  if (layout == DA_ROI_OUT_RCHW) {
    auto k = roi_align_bwd_kernel<TG, DA_ROI_OUT_RCHW>;
    DA_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, BWD_THREADS, smem, st>>>((const TG*)g, C, H, W, R, (const unsigned char*)ws, gin, tiles_x);
  } else {
    auto k = roi_align_bwd_kernel<TG, DA_ROI_OUT_RHWC>;
    DA_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, BWD_THREADS, smem, st>>>((const TG*)g, C, H, W, R, (const unsigned char*)ws, gin, tiles_x);
  }
  DA_LAUNCH_CHECK();
  return DA_OK;
}

static int roi_backward_impl(bool prepared, const void* grad_out, int grad_dtype, int out_layout,
                             const float* rois, int R, int pooled_h, int pooled_w,
                             float spatial_scale, int sampling_ratio, int aligned,
                             void* grad_in, int grad_in_dtype, int N, int C, int H, int W,
                             void* workspace, size_t workspace_bytes, da_stream_t stream) {
  int rc = check_common(N, C, H, W, R, pooled_h, pooled_w, rois, workspace, workspace_bytes);
  if (rc) return rc;
  DA_REQUIRE(grad_in != nullptr, DA_ERR_INVALID_ARG, "roi_align_backward: grad_in is null");
  DA_REQUIRE(grad_in_dtype == DA_F32 || grad_in_dtype == DA_BF16, DA_ERR_INVALID_ARG, "roi_align_backward: bad grad_in dtype %d", grad_in_dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (R == 0) {
    DA_CUDA_OK(cudaMemsetAsync(grad_in, 0, (size_t)N * C * H * W * (grad_in_dtype == DA_BF16 ? 2 : 4), st));
    return DA_OK;
  }
  DA_REQUIRE(grad_out != nullptr, DA_ERR_INVALID_ARG, "roi_align_backward: grad_out is null");
  if (!prepared) {
    rc = run_prep(rois, R, N, H, W, spatial_scale, sampling_ratio, aligned, workspace, nullptr, st);
    if (rc) return rc;
  }
  if (grad_dtype == DA_F32) return launch_bwd<float>(grad_out, out_layout, N, C, H, W, R, workspace, grad_in, grad_in_dtype, st);
  if (grad_dtype == DA_BF16) return launch_bwd<__nv_bfloat16>(grad_out, out_layout, N, C, H, W, R, workspace, grad_in, grad_in_dtype, st);
  DA_REQUIRE(false, DA_ERR_INVALID_ARG, "roi_align_backward: bad dtype %d", grad_dtype);
}

extern "C" int da_roi_align_backward(const void* grad_out, int grad_dtype, int out_layout,
                                     const float* rois, int R, int pooled_h, int pooled_w,
                                     float spatial_scale, int sampling_ratio, int aligned,
                                     void* grad_in, int grad_in_dtype, int N, int C, int H, int W,
                                     void* workspace, size_t workspace_bytes, da_stream_t stream) {
  return roi_backward_impl(false, grad_out, grad_dtype, out_layout, rois, R, pooled_h, pooled_w, spatial_scale, sampling_ratio, aligned,
                           grad_in, grad_in_dtype, N, C, H, W, workspace, workspace_bytes, stream);
}
extern "C" int da_roi_align_backward_prepared(const void* grad_out, int grad_dtype, int out_layout,
                                              const float* rois, int R, int pooled_h, int pooled_w,
                                              float spatial_scale, int sampling_ratio, int aligned,
                                              void* grad_in, int grad_in_dtype, int N, int C, int H, int W,
                                              void* workspace, size_t workspace_bytes, da_stream_t stream) {
  return roi_backward_impl(true, grad_out, grad_dtype, out_layout, rois, R, pooled_h, pooled_w, spatial_scale, sampling_ratio, aligned,
                           grad_in, grad_in_dtype, N, C, H, W, workspace, workspace_bytes, stream);
}

extern "C" int da_map_roi_levels(const float* rois, int R, int num_levels, float finest_scale,
                                 int32_t* levels_out, da_stream_t stream) {
  DA_REQUIRE(R >= 0 && num_levels > 0, DA_ERR_INVALID_ARG, "map_roi_levels: bad args");
  if (R == 0) return DA_OK;
  map_roi_levels_kernel<<<(R + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rois, R, num_levels, finest_scale, levels_out);
  DA_LAUNCH_CHECK();
  return DA_OK;
}
