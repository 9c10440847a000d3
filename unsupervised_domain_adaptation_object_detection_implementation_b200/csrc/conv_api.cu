// C-ABI dispatch of the dense contractions (da_conv_*) onto the two engines:
//   DA_ENGINE_SIMT_F32            -> simt_conv.cu  (fp32 FMA parity engine)
//   DA_ENGINE_UMMA_BF16 / _BF16X3 / _BF16X6 -> umma_conv.cu  (tcgen05 + TMEM + TMA implicit GEMM)
#include "da_common.cuh"

namespace da {
size_t simt_workspace_bytes(const da_conv_desc* d);
int simt_conv_forward(const da_conv_desc* d, const void* x, const void* w, const float* scale,
                      const float* shift, int relu, float drop_p, uint64_t seed, void* y, cudaStream_t st);
int simt_conv_backward_data(const da_conv_desc* d, const void* dz, const void* w, float out_scale, void* dx,
                            cudaStream_t st);
int simt_conv_backward_weight(const da_conv_desc* d, const void* x, const void* dz, float* dw, void* ws,
                              size_t ws_bytes, cudaStream_t st);

size_t umma_workspace_bytes(const da_conv_desc* d);
int umma_conv_forward(const da_conv_desc* d, const void* x, const void* w, const float* scale,
                      const float* shift, int relu, float drop_p, uint64_t seed, void* y, void* ws,
                      size_t ws_bytes, cudaStream_t st);
int umma_conv_backward_data(const da_conv_desc* d, const void* dz, const void* w, float out_scale, void* dx,
                            void* ws, size_t ws_bytes, cudaStream_t st);
int umma_conv_backward_weight(const da_conv_desc* d, const void* x, const void* dz, float* dw, void* ws,
                              size_t ws_bytes, cudaStream_t st, const da_sgd_fuse* sgd);
}  // namespace da

using namespace da;

static int check_desc(const da_conv_desc* d, const char* who) {
  DA_REQUIRE(d != nullptr, DA_ERR_INVALID_ARG, "%s: null descriptor", who);
  DA_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, DA_ERR_INVALID_ARG,
             "%s: bad shape N=%d H=%d W=%d Cin=%d Cout=%d", who, d->N, d->H, d->W, d->Cin, d->Cout);
  DA_REQUIRE(d->KH > 0 && d->KW > 0 && d->stride > 0 && d->pad >= 0, DA_ERR_INVALID_ARG,
             "%s: bad filter %dx%d stride %d pad %d", who, d->KH, d->KW, d->stride, d->pad);
  DA_REQUIRE(d->H + 2 * d->pad >= d->KH && d->W + 2 * d->pad >= d->KW, DA_ERR_INVALID_ARG,
             "%s: filter larger than padded input", who);
  DA_REQUIRE(d->engine >= DA_ENGINE_SIMT_F32 && d->engine <= DA_ENGINE_UMMA_BF16X6, DA_ERR_INVALID_ARG,
             "%s: unknown engine %d", who, d->engine);
  return DA_OK;
}

extern "C" size_t da_conv_workspace_bytes(const da_conv_desc* d) {
  if (!d) return 0;
  const size_t a = simt_workspace_bytes(d);
  const size_t b = umma_workspace_bytes(d);
  return a > b ? a : b;
}

extern "C" int da_conv_forward(const da_conv_desc* d, const void* x, const void* w,
                               const float* scale, const float* shift, int relu,
                               float drop_p, uint64_t drop_seed, void* y, void* workspace,
                               size_t workspace_bytes, da_stream_t stream) {
  int rc = check_desc(d, "conv_forward");
  if (rc) return rc;
  DA_REQUIRE(x && w && y, DA_ERR_INVALID_ARG, "conv_forward: null tensor");
  DA_REQUIRE(drop_p >= 0.f && drop_p < 1.f, DA_ERR_INVALID_ARG, "conv_forward: drop_p=%f out of [0,1)", drop_p);
  cudaStream_t st = (cudaStream_t)stream;
  if (d->engine == DA_ENGINE_SIMT_F32) return simt_conv_forward(d, x, w, scale, shift, relu, drop_p, drop_seed, y, st);
  return umma_conv_forward(d, x, w, scale, shift, relu, drop_p, drop_seed, y, workspace, workspace_bytes, st);
}

extern "C" int da_conv_backward_data(const da_conv_desc* d, const void* dz, const void* w,
                                     float out_scale, void* dx, void* workspace, size_t workspace_bytes,
                                     da_stream_t stream) {
  int rc = check_desc(d, "conv_backward_data");
  if (rc) return rc;
  DA_REQUIRE(dz && w && dx, DA_ERR_INVALID_ARG, "conv_backward_data: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->engine == DA_ENGINE_SIMT_F32) return simt_conv_backward_data(d, dz, w, out_scale, dx, st);
  return umma_conv_backward_data(d, dz, w, out_scale, dx, workspace, workspace_bytes, st);
}

extern "C" int da_conv_backward_weight(const da_conv_desc* d, const void* x, const void* dz,
                                       float* dw, void* workspace, size_t workspace_bytes,
                                       da_stream_t stream) {
  int rc = check_desc(d, "conv_backward_weight");
  if (rc) return rc;
  DA_REQUIRE(x && dz && dw, DA_ERR_INVALID_ARG, "conv_backward_weight: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->engine == DA_ENGINE_SIMT_F32) return simt_conv_backward_weight(d, x, dz, dw, workspace, workspace_bytes, st);
  return umma_conv_backward_weight(d, x, dz, dw, workspace, workspace_bytes, st, nullptr);
}

extern "C" int da_conv_backward_weight_sgd(const da_conv_desc* d, const void* x, const void* dz, const da_sgd_fuse* sgd,
                                           void* workspace, size_t workspace_bytes, da_stream_t stream) {
  int rc = check_desc(d, "conv_backward_weight_sgd");
  if (rc) return rc;
  DA_REQUIRE(x && dz && sgd, DA_ERR_INVALID_ARG, "conv_backward_weight_sgd: null argument");
  DA_REQUIRE(d->engine != DA_ENGINE_SIMT_F32, DA_ERR_UNSUPPORTED,
             "conv_backward_weight_sgd: the fused update lives in the tcgen05 weight-gradient kernel (engine umma_*)");
  return umma_conv_backward_weight(d, x, dz, nullptr, workspace, workspace_bytes, (cudaStream_t)stream, sgd);
}
