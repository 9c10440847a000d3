// Library plumbing (error string, launch counter) + layout / elementwise kernels:
// NCHW<->NHWC transposes (the reference's tensors are NCHW contiguous),
// dtype casts, the bf16 hi/lo split used by the BF16X3 engine, the standalone GRL
// backward (instance_da.py:20-23) and the exported dropout keep-mask.
#include "da_common.cuh"
#include <string.h>
#include <stdlib.h>

namespace da {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
DeviceState g_dev[kMaxDevices] = {};

static int env_int(const char* name) {
  const char* e = getenv(name);
  return e ? (e[0] ? atoi(e) ? atoi(e) : 1 : 1) : 0;     // set (even to "" or a non-number) = 1
}
static Options read_options() {
  Options o;
  memset(&o, 0, sizeof(o));
  o.roi_no_tc = env_int("DA_ROI_NO_TC");
  o.umma_no_bn64 = env_int("DA_UMMA_NO_BN64");
  o.umma_no_2sm = env_int("DA_UMMA_NO_2SM");
  o.no_pdl = env_int("DA_NO_PDL");
  { const char* e = getenv("DA_UMMA_DBG"); o.umma_dbg = e ? atoi(e) : 0; }
  o.umma_no_bn512 = getenv("DA_UMMA_NO_BN512") != nullptr;
  o.chain_no_bn128 = getenv("DA_CHAIN_NO_BN128") != nullptr;
  { const char* e = getenv("DA_ROI_BWD_DBG"); o.roi_bwd_dbg = e ? atoi(e) : 0; }
  { const char* e = getenv("DA_ROI_FWD_DBG"); o.roi_fwd_dbg = e ? atoi(e) : 0; }
  { const char* e = getenv("DA_ROI_BWD_TRACE"); o.roi_bwd_trace = e ? strtoull(e, nullptr, 0) : 0ull; }
  return o;
}
Options g_opt = read_options();     // library load time: the launch paths never call getenv

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// [N][C][HW] -> [N][HW][C] (and the reverse), 32x32 tiles through padded smem
template <typename TS, typename TD>
__global__ void transpose_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int rows, int cols) {
  pdl_launch_dependents();
  // src is [rows][cols] per batch (blockIdx.z), dst is [cols][rows]
  __shared__ float tile[32][33];
  const size_t boff = (size_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = to_f32<TS>(src[boff + (size_t)r * cols + c]);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[boff + (size_t)c * rows + r] = from_f32<TD>(tile[threadIdx.x][j]);
  }
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int64_t n, float mul) {
  pdl_launch_dependents();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = from_f32<TD>(to_f32<TS>(src[i]) * mul);
}

__global__ void split_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, int64_t n) {
  pdl_launch_dependents();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = src[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

__global__ void dropout_mask_kernel(uint64_t seed0, const unsigned long long* ctr, int64_t n, uint32_t thr, uint8_t* __restrict__ keep) {
  const uint64_t seed = effective_seed(seed0, ctr);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    keep[i] = drop_hash(seed, (uint64_t)i) >= thr ? 1 : 0;
}

// torch.optim.SGD (momentum, weight decay, dampening 0, no nesterov) fused with the refresh of the
// bf16 shadow copy the tensor-core engine reads: one pass over (w, grad, buf) instead of
// torch's multi-tensor passes plus a separate fp32->bf16 cast of every weight each step.
__global__ void sgd_step_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ buf,
                                int64_t n, float lr, float mu, float wd, int first, __nv_bfloat16* __restrict__ shadow) {
  pdl_launch_dependents();
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 wv = reinterpret_cast<float4*>(w)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 bv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(buf)[i];
    float* wp = &wv.x; const float* gp = &gv.x; float* bp = &bv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d = fmaf(wd, wp[j], gp[j]);
      bp[j] = first ? d : fmaf(mu, bp[j], d);
      wp[j] = fmaf(-lr, bp[j], wp[j]);
    }
    reinterpret_cast<float4*>(w)[i] = wv;
    reinterpret_cast<float4*>(buf)[i] = bv;
    if (shadow) {
      __nv_bfloat162 a = __floats2bfloat162_rn(wv.x, wv.y), b = __floats2bfloat162_rn(wv.z, wv.w);
      uint2 u;
      u.x = *reinterpret_cast<unsigned*>(&a); u.y = *reinterpret_cast<unsigned*>(&b);
      reinterpret_cast<uint2*>(shadow)[i] = u;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {  // tail
    const int64_t i = (n4 << 2) + threadIdx.x;
    const float d = fmaf(wd, w[i], g[i]);
    const float b = first ? d : fmaf(mu, buf[i], d);
    buf[i] = b;
    w[i] = fmaf(-lr, b, w[i]);
    if (shadow) shadow[i] = __float2bfloat16_rn(w[i]);
  }
}

static inline int ew_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

template <typename TS, typename TD>
static int launch_transpose(const void* src, void* dst, int batch, int rows, int cols, cudaStream_t st) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch);
  DA_REQUIRE(grid.y <= 65535 && grid.z <= 65535, DA_ERR_UNSUPPORTED, "transpose: grid too large");
  transpose_kernel<TS, TD><<<grid, dim3(32, 8), 0, st>>>((const TS*)src, (TD*)dst, rows, cols);
  DA_LAUNCH_CHECK();
  return DA_OK;
}
static int transpose_dispatch(const void* src, int sd, void* dst, int dd, int batch, int rows, int cols, cudaStream_t st) {
  if (sd == DA_F32 && dd == DA_F32) return launch_transpose<float, float>(src, dst, batch, rows, cols, st);
  if (sd == DA_F32 && dd == DA_BF16) return launch_transpose<float, __nv_bfloat16>(src, dst, batch, rows, cols, st);
  if (sd == DA_BF16 && dd == DA_F32) return launch_transpose<__nv_bfloat16, float>(src, dst, batch, rows, cols, st);
  if (sd == DA_BF16 && dd == DA_BF16) return launch_transpose<__nv_bfloat16, __nv_bfloat16>(src, dst, batch, rows, cols, st);
  DA_REQUIRE(false, DA_ERR_INVALID_ARG, "transpose: bad dtype %d/%d", sd, dd);
}

// one block per (tensor, DA_SGD_CHUNK-element chunk) (8 K: the ~5 M parameters of the heads give 600 blocks; at 64 K they
// were 76 blocks on 148 SMs and the pass took 45 us instead of ~20): the small tensors ride along with the big ones
__global__ void __launch_bounds__(256)
sgd_step_multi_kernel(const da_sgd_entry* __restrict__ entries, const int32_t* __restrict__ chunks, float lr, float mu, float wd) {
  pdl_launch_dependents();
  const da_sgd_entry e = entries[chunks[2 * blockIdx.x]];
  const int64_t lo = (int64_t)chunks[2 * blockIdx.x + 1] * DA_SGD_CHUNK;
  const int64_t hi = lo + DA_SGD_CHUNK < e.n ? lo + DA_SGD_CHUNK : e.n;
  const bool first = e.first_step != 0;
  __nv_bfloat16* shadow = reinterpret_cast<__nv_bfloat16*>(e.w_bf16);
  const int64_t hi4 = lo + ((hi - lo) & ~(int64_t)3);
  for (int64_t i = lo + 4 * (int64_t)threadIdx.x; i < hi4; i += 4 * 256) {
    float4 wv = *reinterpret_cast<float4*>(e.w + i);
    const float4 gv = *reinterpret_cast<const float4*>(e.grad + i);
    float4 bv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<float4*>(e.momentum_buf + i);
    float* wp = &wv.x; const float* gp = &gv.x; float* bp = &bv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d = fmaf(wd, wp[j], gp[j]);
      bp[j] = first ? d : fmaf(mu, bp[j], d);
      wp[j] = fmaf(-lr, bp[j], wp[j]);
    }
    *reinterpret_cast<float4*>(e.w + i) = wv;
    *reinterpret_cast<float4*>(e.momentum_buf + i) = bv;
    if (shadow) {
      __nv_bfloat162 a = __floats2bfloat162_rn(wv.x, wv.y), b = __floats2bfloat162_rn(wv.z, wv.w);
      uint2 u;
      u.x = *reinterpret_cast<unsigned*>(&a); u.y = *reinterpret_cast<unsigned*>(&b);
      *reinterpret_cast<uint2*>(shadow + i) = u;
    }
  }
  const int64_t i = hi4 + threadIdx.x;   // tail of the tensor (< 4 elements)
  if (i < hi) {
    const float d = fmaf(wd, e.w[i], e.grad[i]);
    const float b = first ? d : fmaf(mu, e.momentum_buf[i], d);
    e.momentum_buf[i] = b;
    const float wn = fmaf(-lr, b, e.w[i]);
    e.w[i] = wn;
    if (shadow) shadow[i] = __float2bfloat16_rn(wn);
  }
}

}  // namespace da

using namespace da;

extern "C" int da_version(void) { return 100; }
extern "C" const char* da_last_error(void) { return g_err; }
extern "C" int64_t da_launch_count(void) { return g_launches.load(); }
extern "C" void da_launch_count_reset(void) { g_launches.store(0); }

extern "C" int da_nchw_to_nhwc(const void* src, int src_dtype, void* dst, int dst_dtype,
                               int N, int C, int H, int W, da_stream_t stream) {
  DA_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0, DA_ERR_INVALID_ARG, "nchw_to_nhwc: bad args");
  return transpose_dispatch(src, src_dtype, dst, dst_dtype, N, C, H * W, (cudaStream_t)stream);
}
extern "C" int da_nhwc_to_nchw(const void* src, int src_dtype, void* dst, int dst_dtype,
                               int N, int C, int H, int W, da_stream_t stream) {
  DA_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0, DA_ERR_INVALID_ARG, "nhwc_to_nchw: bad args");
  return transpose_dispatch(src, src_dtype, dst, dst_dtype, N, H * W, C, (cudaStream_t)stream);
}

static int cast_scaled(const void* src, int sd, void* dst, int dd, int64_t n, float mul, cudaStream_t st) {
  if (n == 0) return DA_OK;
  DA_REQUIRE(src && dst && n > 0, DA_ERR_INVALID_ARG, "cast: bad args");
  const int b = ew_blocks(n);
  if (sd == DA_F32 && dd == DA_F32) cast_kernel<float, float><<<b, 256, 0, st>>>((const float*)src, (float*)dst, n, mul);
  else if (sd == DA_F32 && dd == DA_BF16) cast_kernel<float, __nv_bfloat16><<<b, 256, 0, st>>>((const float*)src, (__nv_bfloat16*)dst, n, mul);
  else if (sd == DA_BF16 && dd == DA_F32) cast_kernel<__nv_bfloat16, float><<<b, 256, 0, st>>>((const __nv_bfloat16*)src, (float*)dst, n, mul);
  else if (sd == DA_BF16 && dd == DA_BF16) cast_kernel<__nv_bfloat16, __nv_bfloat16><<<b, 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n, mul);
  else DA_REQUIRE(false, DA_ERR_INVALID_ARG, "cast: bad dtype %d/%d", sd, dd);
  DA_LAUNCH_CHECK();
  return DA_OK;
}
extern "C" int da_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, da_stream_t stream) {
  return cast_scaled(src, src_dtype, dst, dst_dtype, n, 1.f, (cudaStream_t)stream);
}
extern "C" int da_grl_backward(const void* grad_out, void* grad_in, int dtype, int64_t n, float weight,
                               da_stream_t stream) {
  return cast_scaled(grad_out, dtype, grad_in, dtype, n, weight, (cudaStream_t)stream);
}
extern "C" int da_split_bf16(const float* src, void* hi, void* lo, int64_t n, da_stream_t stream) {
  if (n == 0) return DA_OK;
  DA_REQUIRE(src && hi && lo && n > 0, DA_ERR_INVALID_ARG, "split_bf16: bad args");
  split_bf16_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n);
  DA_LAUNCH_CHECK();
  return DA_OK;
}
extern "C" int da_dropout_mask(uint64_t seed, int64_t n, float drop_p, uint8_t* keep_out, da_stream_t stream) {
  if (n == 0) return DA_OK;
  DA_REQUIRE(keep_out && n > 0 && drop_p >= 0.f && drop_p < 1.f, DA_ERR_INVALID_ARG, "dropout_mask: bad args");
  dropout_mask_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(seed, g_seed_counter, n, drop_threshold(drop_p), keep_out);
  DA_LAUNCH_CHECK();
  return DA_OK;
}

extern "C" int da_sgd_step(float* w, const float* grad, float* momentum_buf, int64_t n, float lr, float momentum,
                           float weight_decay, int first_step, void* w_bf16, da_stream_t stream) {
  if (n == 0) return DA_OK;
  DA_REQUIRE(w && grad && momentum_buf && n > 0, DA_ERR_INVALID_ARG, "sgd_step: bad args");
  DA_REQUIRE((((uintptr_t)w | (uintptr_t)grad | (uintptr_t)momentum_buf) & 15) == 0 && (((uintptr_t)w_bf16) & 7) == 0,
             DA_ERR_INVALID_ARG, "sgd_step: pointers must be 16-byte aligned");
  sgd_step_kernel<<<ew_blocks((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(w, grad, momentum_buf, n, lr, momentum, weight_decay,
                                                                           first_step, (__nv_bfloat16*)w_bf16);
  DA_LAUNCH_CHECK();
  return DA_OK;
}

extern "C" int da_sgd_step_multi(const da_sgd_entry* entries, int n_entries, const int32_t* chunks, int n_chunks,
                                 float lr, float momentum, float weight_decay, da_stream_t stream) {
  DA_REQUIRE(n_entries >= 0 && n_chunks >= 0, DA_ERR_INVALID_ARG, "sgd_step_multi: negative counts");
  if (n_entries == 0 || n_chunks == 0) return DA_OK;
  DA_REQUIRE(entries && chunks, DA_ERR_INVALID_ARG, "sgd_step_multi: null table");
  sgd_step_multi_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(entries, chunks, lr, momentum, weight_decay);
  DA_LAUNCH_CHECK();
  return DA_OK;
}

extern "C" int da_set_sm_limit(int n) {        // applies to the calling thread's CURRENT device
  dev_state().sm_limit = n > 0 ? n : 0;
  return DA_OK;
}

extern "C" int da_set_dropout_counter(const void* counter_dev) {   // per device, like the counter itself
  dev_state().seed_counter = (const unsigned long long*)counter_dev;
  return DA_OK;
}

extern "C" int da_set_option(const char* name, long long value) {
  DA_REQUIRE(name != nullptr, DA_ERR_INVALID_ARG, "set_option: null name");
  if (!strcmp(name, "roi_no_tc")) g_opt.roi_no_tc = (int)value;
  else if (!strcmp(name, "umma_no_bn64")) g_opt.umma_no_bn64 = (int)value;
  else if (!strcmp(name, "umma_no_2sm")) g_opt.umma_no_2sm = (int)value;
  else if (!strcmp(name, "umma_no_bn512")) g_opt.umma_no_bn512 = (int)value;
  else if (!strcmp(name, "chain_no_bn128")) g_opt.chain_no_bn128 = (int)value;
  else if (!strcmp(name, "no_pdl")) g_opt.no_pdl = (int)value;
  else if (!strcmp(name, "umma_dbg")) g_opt.umma_dbg = (int)value;
  else if (!strcmp(name, "roi_bwd_dbg")) g_opt.roi_bwd_dbg = (int)value;
  else if (!strcmp(name, "roi_fwd_dbg")) g_opt.roi_fwd_dbg = (int)value;
  else if (!strcmp(name, "roi_bwd_trace")) g_opt.roi_bwd_trace = (unsigned long long)value;
  else if (!strcmp(name, "chain_trace")) g_opt.chain_trace = (unsigned long long)value;
  else DA_REQUIRE(false, DA_ERR_INVALID_ARG, "set_option: unknown option '%s'", name);
  return DA_OK;
}
