// Query-axis softmax of one KEY BLOCK of a NonLocalBlock's score matrix — sm_100a.
//
// The reference block (mmdet/models/backbones/resnet_da_deep.py:402-445, roi_heads/instance_da.py:150-192) applies
// nn.Softmax(dim=1) to the [b, T, T] scores theta^T.phi, i.e. it normalises over the QUERY axis (Q11): every key column k has
// its own max and its own sum over the T queries.  Key columns are therefore independent, and the T x T matrix (4.3 GB per image
// in fp32 at T = 32768, a 1024x2048 input on C3) never has to exist: functional.nonlocal_attention_blocked walks the keys in
// blocks of Tk columns; per block S_b = theta.phi_b^T is [Tq x Tk] and these kernels normalise it in two passes:
//   pass 1  per (row chunk, column): running max m and sum of exp(s - m), rescaled once per 4 rows    (colstats_partial)
//           per column: merge the chunks -> stats = (max, 1 / sum)                               (colstats_combine)
//   pass 2  p = exp(s - max) / sum, written once in the GEMM operand dtype (bf16 or fp32)         (col_apply)
// The backward of the block, ds = p * (dp - sum_q p*dp), has the same shape: per-chunk column dots, merge, apply.
// HBM-bound element-wise work: coalesced rows (a warp reads 64 consecutive columns as float2 = 256 B), 4 rows in flight per
// thread, deterministic fixed-order merges, no atomics.  Algorithmic bytes per element of S_b: forward 4 + 4 read + sizeof(p)
// written; backward sizeof(p) + 4 read twice + sizeof(ds) written.
#include "da_common.cuh"

namespace da {

constexpr int CS_TY = 8;          // row lanes per CTA
constexpr int CS_UNROLL = 4;      // rows in flight per thread

struct ColTiling { int strips, nchunk, rows_per_chunk, vec; };
static ColTiling col_tiling(int Tq, int Tk, int vec) {
  ColTiling t;
  t.vec = vec;
  t.strips = (Tk + 32 * vec - 1) / (32 * vec);
  int want = (4 * num_sms_physical() + t.strips - 1) / t.strips;     // >= 4 CTAs per SM over the whole grid
  const int max_chunks = (Tq + CS_TY * CS_UNROLL - 1) / (CS_TY * CS_UNROLL);
  if (want > max_chunks) want = max_chunks;
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  t.rows_per_chunk = ((Tq + want - 1) / want + CS_TY - 1) / CS_TY * CS_TY;
  t.nchunk = (Tq + t.rows_per_chunk - 1) / t.rows_per_chunk;
  return t;
}

template <typename T> struct Ld2;
template <> struct Ld2<float> {
  static __device__ __forceinline__ float2 ld(const float* p) { return *reinterpret_cast<const float2*>(p); }
  static __device__ __forceinline__ void st(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
};
template <> struct Ld2<__nv_bfloat16> {
  static __device__ __forceinline__ float2 ld(const __nv_bfloat16* p) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p)); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float a, float b) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b); }
};

__device__ __forceinline__ void merge(float m2, float z2, float& m, float& z) {
  if (z2 == 0.f) return;
  if (z == 0.f) { m = m2; z = z2; return; }
  const float mm = fmaxf(m, m2);
  z = z * expf(m - mm) + z2 * expf(m2 - mm);
  m = mm;
}

// grid (strips, nchunk), block (32, CS_TY).  VEC = 2: lane owns columns 2*tx, 2*tx+1 of a 64-column strip.
template <int VEC>
__global__ void __launch_bounds__(32 * CS_TY)
colstats_partial_kernel(const float* __restrict__ s, int Tq, int Tk, int ld, int rows_per_chunk, float2* __restrict__ partial) {
  pdl_launch_dependents();
  __shared__ float sm_m[CS_TY][32 * VEC], sm_z[CS_TY][32 * VEC];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int k0 = (blockIdx.x * 32 + tx) * VEC;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(r0 + rows_per_chunk, Tq);
  float m[VEC], z[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { m[v] = -INFINITY; z[v] = 0.f; }
  if (k0 < Tk) {
    for (int r = r0 + ty; r < r1; r += CS_TY * CS_UNROLL) {
      float x[CS_UNROLL][VEC];
#pragma unroll
      for (int u = 0; u < CS_UNROLL; ++u) {
        const int rr = r + u * CS_TY;
        if (rr < r1) {
          if (VEC == 2) { const float2 t = Ld2<float>::ld(s + (size_t)rr * ld + k0); x[u][0] = t.x; x[u][VEC - 1] = t.y; }
          else x[u][0] = s[(size_t)rr * ld + k0];
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) x[u][v] = -INFINITY;
        }
      }
      // one rescale per CS_UNROLL rows and no data-dependent branch: ncu of the per-element form (rescale only when the max
      // moves) showed the kernel issue-bound at 40 % of DRAM -- both sides of the divergent branch ran, two exps per element
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float mx = x[0][v];
#pragma unroll
        for (int u = 1; u < CS_UNROLL; ++u) mx = fmaxf(mx, x[u][v]);
        const float mn = fmaxf(m[v], mx);                 // finite: row u = 0 of the group is always in range
        float acc = z[v] * expf(m[v] - mn);               // first group: 0 * exp(-inf) = 0
#pragma unroll
        for (int u = 0; u < CS_UNROLL; ++u) acc += expf(x[u][v] - mn);    // masked rows: exp(-inf) = 0
        z[v] = acc;
        m[v] = mn;
      }
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) { sm_m[ty][tx * VEC + v] = m[v]; sm_z[ty][tx * VEC + v] = z[v]; }
  __syncthreads();
  if (ty == 0 && k0 < Tk) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float mm = sm_m[0][tx * VEC + v], zz = sm_z[0][tx * VEC + v];
      for (int j = 1; j < CS_TY; ++j) merge(sm_m[j][tx * VEC + v], sm_z[j][tx * VEC + v], mm, zz);     // fixed order
      if (k0 + v < Tk) partial[(size_t)blockIdx.y * Tk + k0 + v] = make_float2(mm, zz);
    }
  }
}

__global__ void colstats_combine_kernel(const float2* __restrict__ partial, int nchunk, int Tk, float* __restrict__ stats) {
  pdl_launch_dependents();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Tk) return;
  float m = -INFINITY, z = 0.f;
  for (int c = 0; c < nchunk; ++c) {
    const float2 t = partial[(size_t)c * Tk + k];
    merge(t.x, t.y, m, z);
  }
  stats[k] = m;
  stats[Tk + k] = 1.f / z;
}

template <int VEC, typename TP>
__global__ void __launch_bounds__(32 * CS_TY)
col_apply_kernel(const float* __restrict__ s, int Tq, int Tk, int ld, int rows_per_chunk, const float* __restrict__ stats,
                 TP* __restrict__ p) {
  pdl_launch_dependents();
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int k0 = (blockIdx.x * 32 + tx) * VEC;
  if (k0 >= Tk) return;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(r0 + rows_per_chunk, Tq);
  float m[VEC], iz[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { m[v] = stats[min(k0 + v, Tk - 1)]; iz[v] = stats[Tk + min(k0 + v, Tk - 1)]; }
  for (int r = r0 + ty; r < r1; r += CS_TY * CS_UNROLL) {
    float x[CS_UNROLL][VEC];
#pragma unroll
    for (int u = 0; u < CS_UNROLL; ++u) {
      const int rr = r + u * CS_TY;
      if (rr < r1) {
        if (VEC == 2) { const float2 t = Ld2<float>::ld(s + (size_t)rr * ld + k0); x[u][0] = t.x; x[u][VEC - 1] = t.y; }
        else x[u][0] = s[(size_t)rr * ld + k0];
      }
    }
#pragma unroll
    for (int u = 0; u < CS_UNROLL; ++u) {
      const int rr = r + u * CS_TY;
      if (rr < r1) {
        if (VEC == 2) Ld2<TP>::st(p + (size_t)rr * ld + k0, expf(x[u][0] - m[0]) * iz[0], expf(x[u][VEC - 1] - m[VEC - 1]) * iz[VEC - 1]);
        else p[(size_t)rr * ld + k0] = from_f32<TP>(expf(x[u][0] - m[0]) * iz[0]);
      }
    }
  }
}

// backward pass 1: per (row chunk, column) sum_q p*dp
template <int VEC, typename TP>
__global__ void __launch_bounds__(32 * CS_TY)
coldot_partial_kernel(const TP* __restrict__ p, const float* __restrict__ dp, int Tq, int Tk, int ld, int rows_per_chunk,
                      float* __restrict__ partial) {
  pdl_launch_dependents();
  __shared__ float sm_d[CS_TY][32 * VEC];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int k0 = (blockIdx.x * 32 + tx) * VEC;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(r0 + rows_per_chunk, Tq);
  float d[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) d[v] = 0.f;
  if (k0 < Tk) {
    for (int r = r0 + ty; r < r1; r += CS_TY * CS_UNROLL) {
      float a[CS_UNROLL][VEC], b[CS_UNROLL][VEC];
#pragma unroll
      for (int u = 0; u < CS_UNROLL; ++u) {
        const int rr = r + u * CS_TY;
#pragma unroll
        for (int v = 0; v < VEC; ++v) { a[u][v] = 0.f; b[u][v] = 0.f; }
        if (rr < r1) {
          const size_t i = (size_t)rr * ld + k0;
          if (VEC == 2) {
            const float2 ta = Ld2<TP>::ld(p + i), tb = Ld2<float>::ld(dp + i);
            a[u][0] = ta.x; a[u][VEC - 1] = ta.y; b[u][0] = tb.x; b[u][VEC - 1] = tb.y;
          } else { a[u][0] = to_f32<TP>(p[i]); b[u][0] = dp[i]; }
        }
      }
#pragma unroll
      for (int u = 0; u < CS_UNROLL; ++u)
#pragma unroll
        for (int v = 0; v < VEC; ++v) d[v] = fmaf(a[u][v], b[u][v], d[v]);
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) sm_d[ty][tx * VEC + v] = d[v];
  __syncthreads();
  if (ty == 0 && k0 < Tk) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float t = sm_d[0][tx * VEC + v];
      for (int j = 1; j < CS_TY; ++j) t += sm_d[j][tx * VEC + v];
      if (k0 + v < Tk) partial[(size_t)blockIdx.y * Tk + k0 + v] = t;
    }
  }
}

__global__ void coldot_combine_kernel(const float* __restrict__ partial, int nchunk, int Tk, float* __restrict__ dot) {
  pdl_launch_dependents();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Tk) return;
  float t = 0.f;
  for (int c = 0; c < nchunk; ++c) t += partial[(size_t)c * Tk + k];
  dot[k] = t;
}

template <int VEC, typename TP, typename TD>
__global__ void __launch_bounds__(32 * CS_TY)
col_bwd_apply_kernel(const TP* __restrict__ p, const float* __restrict__ dp, int Tq, int Tk, int ld, int rows_per_chunk,
                     const float* __restrict__ dot, TD* __restrict__ ds) {
  pdl_launch_dependents();
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int k0 = (blockIdx.x * 32 + tx) * VEC;
  if (k0 >= Tk) return;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(r0 + rows_per_chunk, Tq);
  float dk[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) dk[v] = dot[min(k0 + v, Tk - 1)];
  for (int r = r0 + ty; r < r1; r += CS_TY * CS_UNROLL) {
    float a[CS_UNROLL][VEC], b[CS_UNROLL][VEC];
#pragma unroll
    for (int u = 0; u < CS_UNROLL; ++u) {
      const int rr = r + u * CS_TY;
      if (rr < r1) {
        const size_t i = (size_t)rr * ld + k0;
        if (VEC == 2) {
          const float2 ta = Ld2<TP>::ld(p + i), tb = Ld2<float>::ld(dp + i);
          a[u][0] = ta.x; a[u][VEC - 1] = ta.y; b[u][0] = tb.x; b[u][VEC - 1] = tb.y;
        } else { a[u][0] = to_f32<TP>(p[i]); b[u][0] = dp[i]; }
      }
    }
#pragma unroll
    for (int u = 0; u < CS_UNROLL; ++u) {
      const int rr = r + u * CS_TY;
      if (rr < r1) {
        const size_t i = (size_t)rr * ld + k0;
        if (VEC == 2) Ld2<TD>::st(ds + i, a[u][0] * (b[u][0] - dk[0]), a[u][VEC - 1] * (b[u][VEC - 1] - dk[VEC - 1]));
        else ds[i] = from_f32<TD>(a[u][0] * (b[u][0] - dk[0]));
      }
    }
  }
}

static bool aligned_for(const void* q, int dtype) {      // a 2-element access of that dtype
  return (reinterpret_cast<uintptr_t>(q) & (dtype == DA_F32 ? 7u : 3u)) == 0;
}

}  // namespace da

using namespace da;

extern "C" size_t da_colsoftmax_workspace_bytes(int Tq, int Tk) {
  if (Tq <= 0 || Tk <= 0) return 0;
  const int nchunk = max(col_tiling(Tq, Tk, 1).nchunk, col_tiling(Tq, Tk, 2).nchunk);
  return align_up((size_t)nchunk * Tk * sizeof(float2) + (size_t)Tk * sizeof(float), 256);
}

#define CS_LAUNCH(kern, grid, ...)                                  \
  do {                                                              \
    kern<<<grid, dim3(32, CS_TY), 0, (cudaStream_t)stream>>>(__VA_ARGS__); \
    DA_LAUNCH_CHECK();                                              \
  } while (0)

extern "C" int da_colsoftmax_forward(const float* s, int Tq, int Tk, int ld, void* p, int p_dtype, float* stats, int have_stats,
                                     void* workspace, size_t workspace_bytes, da_stream_t stream) {
  DA_REQUIRE(Tq > 0 && Tk > 0 && ld >= Tk && s && p && stats, DA_ERR_INVALID_ARG, "colsoftmax_forward: bad args (Tq=%d Tk=%d ld=%d)", Tq, Tk, ld);
  DA_REQUIRE(p_dtype == DA_F32 || p_dtype == DA_BF16, DA_ERR_UNSUPPORTED, "colsoftmax_forward: p must be fp32 or bf16");
  DA_REQUIRE(have_stats || (workspace && workspace_bytes >= da_colsoftmax_workspace_bytes(Tq, Tk)), DA_ERR_INVALID_ARG,
             "colsoftmax_forward: workspace too small");
  const int vec = (Tk % 2 == 0 && ld % 2 == 0 && aligned_for(s, DA_F32) && aligned_for(p, p_dtype)) ? 2 : 1;
  const ColTiling t = col_tiling(Tq, Tk, vec);
  const dim3 grid((unsigned)t.strips, (unsigned)t.nchunk);
  if (!have_stats) {
    float2* partial = static_cast<float2*>(workspace);
    if (vec == 2) CS_LAUNCH(colstats_partial_kernel<2>, grid, s, Tq, Tk, ld, t.rows_per_chunk, partial);
    else CS_LAUNCH(colstats_partial_kernel<1>, grid, s, Tq, Tk, ld, t.rows_per_chunk, partial);
    colstats_combine_kernel<<<(Tk + 255) / 256, 256, 0, (cudaStream_t)stream>>>(partial, t.nchunk, Tk, stats);
    DA_LAUNCH_CHECK();
  }
  if (p_dtype == DA_BF16) {
    if (vec == 2) CS_LAUNCH((col_apply_kernel<2, __nv_bfloat16>), grid, s, Tq, Tk, ld, t.rows_per_chunk, stats, static_cast<__nv_bfloat16*>(p));
    else CS_LAUNCH((col_apply_kernel<1, __nv_bfloat16>), grid, s, Tq, Tk, ld, t.rows_per_chunk, stats, static_cast<__nv_bfloat16*>(p));
  } else {
    if (vec == 2) CS_LAUNCH((col_apply_kernel<2, float>), grid, s, Tq, Tk, ld, t.rows_per_chunk, stats, static_cast<float*>(p));
    else CS_LAUNCH((col_apply_kernel<1, float>), grid, s, Tq, Tk, ld, t.rows_per_chunk, stats, static_cast<float*>(p));
  }
  return DA_OK;
}

template <int VEC, typename TP>
static int colsoftmax_bwd_t(const TP* p, const float* dp, int Tq, int Tk, int ld, void* ds, int ds_dtype, void* workspace,
                            const ColTiling& t, da_stream_t stream) {
  float* partial = static_cast<float*>(workspace);
  float* dot = partial + (size_t)t.nchunk * Tk;
  const dim3 grid((unsigned)t.strips, (unsigned)t.nchunk);
  CS_LAUNCH((coldot_partial_kernel<VEC, TP>), grid, p, dp, Tq, Tk, ld, t.rows_per_chunk, partial);
  coldot_combine_kernel<<<(Tk + 255) / 256, 256, 0, (cudaStream_t)stream>>>(partial, t.nchunk, Tk, dot);
  DA_LAUNCH_CHECK();
  if (ds_dtype == DA_BF16) CS_LAUNCH((col_bwd_apply_kernel<VEC, TP, __nv_bfloat16>), grid, p, dp, Tq, Tk, ld, t.rows_per_chunk, dot, static_cast<__nv_bfloat16*>(ds));
  else CS_LAUNCH((col_bwd_apply_kernel<VEC, TP, float>), grid, p, dp, Tq, Tk, ld, t.rows_per_chunk, dot, static_cast<float*>(ds));
  return DA_OK;
}

extern "C" int da_colsoftmax_backward(const void* p, int p_dtype, const float* dp, int Tq, int Tk, int ld, void* ds, int ds_dtype,
                                      void* workspace, size_t workspace_bytes, da_stream_t stream) {
  DA_REQUIRE(Tq > 0 && Tk > 0 && ld >= Tk && p && dp && ds, DA_ERR_INVALID_ARG, "colsoftmax_backward: bad args (Tq=%d Tk=%d ld=%d)", Tq, Tk, ld);
  DA_REQUIRE((p_dtype == DA_F32 || p_dtype == DA_BF16) && (ds_dtype == DA_F32 || ds_dtype == DA_BF16), DA_ERR_UNSUPPORTED,
             "colsoftmax_backward: p / ds must be fp32 or bf16");
  DA_REQUIRE(workspace && workspace_bytes >= da_colsoftmax_workspace_bytes(Tq, Tk), DA_ERR_INVALID_ARG, "colsoftmax_backward: workspace too small");
  const int vec = (Tk % 2 == 0 && ld % 2 == 0 && aligned_for(dp, DA_F32) && aligned_for(p, p_dtype) && aligned_for(ds, ds_dtype)) ? 2 : 1;
  const ColTiling t = col_tiling(Tq, Tk, vec);
  if (p_dtype == DA_BF16)
    return vec == 2 ? colsoftmax_bwd_t<2>(static_cast<const __nv_bfloat16*>(p), dp, Tq, Tk, ld, ds, ds_dtype, workspace, t, stream)
                    : colsoftmax_bwd_t<1>(static_cast<const __nv_bfloat16*>(p), dp, Tq, Tk, ld, ds, ds_dtype, workspace, t, stream);
  return vec == 2 ? colsoftmax_bwd_t<2>(static_cast<const float*>(p), dp, Tq, Tk, ld, ds, ds_dtype, workspace, t, stream)
                  : colsoftmax_bwd_t<1>(static_cast<const float*>(p), dp, Tq, Tk, ld, ds, ds_dtype, workspace, t, stream);
}
