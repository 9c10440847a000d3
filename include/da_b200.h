/*
 * da_b200.h — C-ABI of the B200-native domain-adaptation hot path.
 *
 * One shared library (libda_b200.so, sm_100a only) replaces every native op the
 * reference reaches on its DA path.  The reference itself ships no native code
 * (/root/reference/setup.py:220, ext_modules=[]); what it binds is mmcv-full's
 * `ext_module` (mmcv.ops.RoIAlign, sigmoid_focal_loss) plus ATen/cuDNN/cuBLAS
 * through torch.nn.  Each entry point below cites the reference call site it
 * replaces (path:line under /root/reference).
 *
 * Conventions (same as mmcv's ext_module):
 *   - the CALLER owns every buffer, including workspaces (query *_workspace_bytes);
 *     the library never allocates device memory;
 *   - all pointers are device pointers unless the name ends in _host;
 *   - kernels are enqueued on `stream` and never synchronise;
 *   - return 0 on success, a DA_ERR_* code otherwise; da_last_error() gives the
 *     message for the calling thread;
 *   - no torch types, no C++ types: plain pointers, ints, floats.
 *
 * Layouts: activation tensors are NHWC ("channels-last", the physical layout of a
 * torch channels_last tensor).  The reference's NCHW tensors enter through
 * da_nchw_to_nhwc.  RoI features keep the reference layout [R,C,ph,pw]
 * (DA_ROI_OUT_RCHW) by default because the bbox head flattens them as (c,ph,pw)
 * (mmdet/models/roi_heads/bbox_heads/convfc_bbox_head.py:208).  DA_ROI_OUT_RHWC
 * ([R,ph,pw,C], "bin-major") is the faster layout for bf16 tensors with C % 64 == 0
 * (both tensor-core kernels serve it; the backward then fetches its operand by TMA as it
 * lies in memory); a caller using it holds the first shared FC's weight with its input
 * columns in (ph,pw,c) order ("bin-major RoI features").
 */
#ifndef DA_B200_H_
#define DA_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* da_stream_t; /* cudaStream_t */

enum da_status {
  DA_OK = 0,
  DA_ERR_INVALID_ARG = 1,
  DA_ERR_UNSUPPORTED = 2,
  DA_ERR_CUDA = 3,
  DA_ERR_WORKSPACE = 4,
  DA_ERR_ROI_BATCH_INDEX = 5 /* reported asynchronously through the error flag of the workspace */
};

enum da_dtype { DA_F32 = 0, DA_BF16 = 1 };

enum da_roi_out_layout { DA_ROI_OUT_RCHW = 0, DA_ROI_OUT_RHWC = 1 };

/* GEMM engine of the dense contractions (conv / FC). */
enum da_engine {
  DA_ENGINE_SIMT_F32 = 0,   /* CUDA-core fp32 FMA: bit-faithful fp32 parity mode          */
  DA_ENGINE_UMMA_BF16 = 1,  /* tcgen05.mma kind::f16 (bf16 in, fp32 TMEM accumulate)      */
  DA_ENGINE_UMMA_BF16X3 = 2, /* 3-term split bf16 (hi*hi + hi*lo + lo*hi), ~2^-16 relative */
  DA_ENGINE_UMMA_BF16X6 = 3  /* exact 3-way bf16 split of fp32 operands, 6 product terms on tcgen05:
                                fp32-class (<= 1e-5 parity bar), 1/6 of the bf16 tensor rate (= 3xTF32) */
};

/* ---- library ------------------------------------------------------------ */
int da_version(void);
const char* da_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t da_launch_count(void);
void da_launch_count_reset(void);
/* Debug / test options.  The library reads its DA_* environment variables ONCE, when it is loaded (no getenv on any
 * launch path); this call changes one afterwards.  Names: "roi_no_tc" (bf16 RoIAlign on the CUDA-core kernels instead of
 * the tcgen05 ones), "umma_no_bn64", "umma_no_2sm", "umma_no_bn512", "chain_no_bn128", "no_pdl", "umma_dbg", "roi_bwd_dbg", "roi_fwd_dbg", "roi_bwd_trace", "chain_trace".  Process-wide
 * and not meant to be flipped while other threads launch; defaults (all 0) are what production runs.
 * Per-DEVICE state (da_set_sm_limit, da_set_dropout_counter) belongs to the calling thread's current CUDA device. */
int da_set_option(const char* name, long long value);

/* ---- layout + GRL --------------------------------------------------------
 * GRL: mmdet/models/roi_heads/instance_da.py:14-40 (_GradientScalarLayer): forward is the
 * identity, backward is weight * grad.  Normally folded into a dgrad epilogue
 * (out_scale of da_conv_backward_data); this standalone form serves callers that
 * wrap an arbitrary torch sub-graph. */
int da_grl_backward(const void* grad_out, void* grad_in, int dtype, int64_t n, float weight,
                    da_stream_t stream);
int da_nchw_to_nhwc(const void* src, int src_dtype, void* dst, int dst_dtype,
                    int N, int C, int H, int W, da_stream_t stream);
int da_nhwc_to_nchw(const void* src, int src_dtype, void* dst, int dst_dtype,
                    int N, int C, int H, int W, da_stream_t stream);
/* 3-term bf16 split of an fp32 tensor: hi=bf16(x), lo=bf16(x-hi). */
int da_split_bf16(const float* src, void* hi, void* lo, int64_t n, da_stream_t stream);
int da_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, da_stream_t stream);

/* SGD step of the reference recipe (torch.optim.SGD built from da_configs/faster_rcnn/
 * faster_rcnn_r50_daf_c2f.py:8 by mmdet/apis/train.py:127): momentum, weight decay, dampening 0,
 * no nesterov:  d = grad + wd*w;  buf = first_step ? d : mu*buf + d;  w -= lr*buf.
 * w_bf16 (nullable): bf16 shadow of w refreshed in the same pass (operand of the tcgen05 engine). */
/* SM budget of the persistent kernels (0 = all SMs).  Lower it for launches that overlap a collective holding SMs
 * (dist.OverlappedGradAllReduce): a persistent grid larger than the free SMs runs in two waves. */
int da_set_sm_limit(int n);
int da_sgd_step(float* w, const float* grad, float* momentum_buf, int64_t n, float lr, float momentum,
                float weight_decay, int first_step, void* w_bf16, da_stream_t stream);
/* The same update for MANY tensors in one launch (torch.optim.SGD's foreach path, mmdet/apis/train.py:127).
 * `entries` is a DEVICE array of n_entries records; `chunks` a DEVICE array of n_chunks (entry, chunk) pairs that
 * tiles every tensor in pieces of DA_SGD_CHUNK elements (built once by the caller: sizes do not change). */
#define DA_SGD_CHUNK 8192
typedef struct da_sgd_entry {
  float* w;              /* fp32 weights, 16-byte aligned */
  const float* grad;     /* fp32 gradient, same memory order */
  float* momentum_buf;   /* fp32 momentum buffer */
  void* w_bf16;          /* optional bf16 shadow (NULL = none) */
  int64_t n;             /* elements */
  int32_t first_step;    /* 1 = the buffer is uninitialised (torch's first-step rule: buf = d) */
  int32_t _pad;
} da_sgd_entry;
int da_sgd_step_multi(const da_sgd_entry* entries, int n_entries, const int32_t* chunks, int n_chunks,
                      float lr, float momentum, float weight_decay, da_stream_t stream);

/* ---- gradient mean + SGD + operand broadcast over NVLink peer memory (N GPUs of one box) ------------------
 * Replaces, for one large tensor, the pair  MMDistributedDataParallel bucket all-reduce (mmdet/apis/train.py:113-121)
 * -> torch.optim.SGD.step (train.py:127): rank r owns the r-th 1/world slice (slices are multiples of 1024 elements),
 * reads that slice of EVERY rank's gradient through P2P loads, averages in rank order, applies the SGD rule above to
 * its slice of the fp32 master (momentum is stored sharded: momentum_shard holds the slice only) and stores the bf16
 * result into every rank's operand copy (and, when w_f32[r] is given, the fp32 result into every rank's master; with
 * w_f32[r] == NULL the master of a rank is current on its own slice only, ZeRO-1 style).
 * Cross-GPU ordering uses flag words in peer memory: the call may start as soon as THIS rank's gradient is complete in
 * stream order; on return of the second (wait) kernel nobody reads this rank's gradient or writes its operand copy
 * any more.  Every rank must make the same sequence of calls.  A barrier that does not complete within 4 s sets
 * local_state[2] (1 = in, 2 = out) instead of hanging.
 * Memory for grad / w_bf16 / flags must be reachable by every peer: da_peer_alloc (cudaMalloc, zero-filled) +
 * da_peer_export (64-byte CUDA IPC handle, exchanged by the caller) + da_peer_open on the other ranks. */
#define DA_MAX_PEERS 8
#define DA_PEER_HANDLE_BYTES 64
#define DA_PEER_FLAG_INTS (2 * DA_MAX_PEERS)   /* ready[DA_MAX_PEERS] | done[DA_MAX_PEERS], zero-initialised */
typedef struct da_peer_sgd_args {
  float* w;                          /* local fp32 master, full tensor */
  float* momentum_shard;             /* local fp32 momentum of the own slice (slice-relative index) */
  const float* grad[DA_MAX_PEERS];   /* gradient buffer of every rank (own entry = local pointer) */
  void* w_bf16[DA_MAX_PEERS];        /* bf16 operand copy of every rank (NULL = skip) */
  float* w_f32[DA_MAX_PEERS];        /* fp32 master of every rank (NULL = master stays sharded) */
  int* flags[DA_MAX_PEERS];          /* flag block (DA_PEER_FLAG_INTS ints) of every rank */
  int* local_state;                  /* 4 local ints, zero-initialised: epoch, ticket, error, pad */
  int64_t n;                         /* elements of the tensor */
  int32_t world, rank;
} da_peer_sgd_args;
 /* publish_mode DA_PEER_PUBLISH_STORES: the kernel does all of the above with SM-issued P2P loads and stores, followed by
 * the wait kernel.  DA_PEER_PUBLISH_BY_CALLER: the copy engines carry the data - before the call the caller pushes slice r
 * of its gradient into rank r's staging area with da_peer_copy (so grad[r] here are LOCAL pointers biased such that
 * grad[r] + i addresses element i of the own slice), after the call it pushes the refreshed bf16 slice (w_bf16[rank] is
 * the local copy, the other entries NULL) to every rank and ends with da_peer_publish_done (done flags + wait). */
#define DA_PEER_PUBLISH_STORES 0
#define DA_PEER_PUBLISH_BY_CALLER 1
int da_sgd_step_peer(const da_peer_sgd_args* args, float lr, float momentum, float weight_decay, int first_step,
                     int max_ctas, int publish_mode, da_stream_t stream);
/* Stream-ordered copy between any two device pointers this process can address (local or da_peer_open'ed): runs on a
 * copy engine, no SM involved. */
int da_peer_copy(void* dst, const void* src, size_t bytes, da_stream_t stream);
int da_peer_publish_done(const da_peer_sgd_args* args, da_stream_t stream);   /* = da_peer_signal_done + da_peer_wait_done */
/* The two halves, for a DEFERRED publish: the pushes of step t and da_peer_signal_done may be enqueued after the step (on a
 * side stream, outside a captured graph) and run under the beginning of step t+1; da_peer_wait_done then goes in front of
 * the first kernel of step t+1 that reads the operand copy.  local_state[3] counts this rank's completed publishes; the
 * update kernel of step t+1 waits for publish t before it rewrites the slice the pushes read (error 3 on time-out). */
int da_peer_signal_done(const da_peer_sgd_args* args, da_stream_t stream);
int da_peer_wait_done(const da_peer_sgd_args* args, da_stream_t stream);
/* SETUP-TIME ONLY (never on the step path): the one place the library allocates, because a CUDA IPC handle needs a
 * dedicated cudaMalloc block (a sub-allocation of a caching allocator cannot be exported); zero-fills and synchronises. */
int da_peer_alloc(size_t bytes, void** out);
int da_peer_free(void* p);
int da_peer_export(const void* p, unsigned char* handle64);
int da_peer_open(const unsigned char* handle64, void** out);
int da_peer_close(void* p);

/* ---- RoIAlign ------------------------------------------------------------
 * Replaces mmcv.ops.RoIAlign (ext_module.roi_align_forward / roi_align_backward,
 * mmcv-full 1.3.17) as built at mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:54-60
 * and called at single_level_roi_extractor.py:79,103.  pool_mode='avg'.
 * rois: [R,5] fp32 (batch_ind, x1, y1, x2, y2) in image pixels.
 *
 * Workspace: da_roi_align_workspace_bytes(R,H,W) bytes.  Layout: int32 err_flag[4], one meta
 * record per RoI, then per RoI the separable tap-weight tables ((H+W) rows of 8 floats).  After the stream is synchronised err_flag[0] != 0 means some
 * RoI carried a batch index outside [0,N); such RoIs produce zeros.
 * grid_out (nullable): int32 [R,2] = (roi_bin_grid_h, roi_bin_grid_w), the sampling grid. */
size_t da_roi_align_workspace_bytes(int R, int H, int W);
int da_roi_align_forward(const void* feat_nhwc, int feat_dtype, int N, int C, int H, int W,
                         const float* rois, int R, int pooled_h, int pooled_w,
                         float spatial_scale, int sampling_ratio, int aligned,
                         void* out, int out_dtype, int out_layout,
                         int32_t* grid_out, void* workspace, size_t workspace_bytes,
                         da_stream_t stream);
/* grad_in_nhwc [N,H,W,C] fp32 is written exactly once per element (no atomics, no
 * pre-zeroing required). */
int da_roi_align_backward(const void* grad_out, int grad_dtype, int out_layout,
                          const float* rois, int R, int pooled_h, int pooled_w,
                          float spatial_scale, int sampling_ratio, int aligned,
                          void* grad_in_nhwc, int grad_in_dtype, int N, int C, int H, int W,
                          void* workspace, size_t workspace_bytes, da_stream_t stream);
/* Same, for a workspace that still holds the preparation of THIS RoI set: the last call that wrote `workspace` was
 * da_roi_align_forward (or da_roi_align_backward) with the same rois, R, N, H, W, spatial_scale, sampling_ratio and aligned,
 * on the same stream.  Skips the preparation launch (tap tables, footprints) -- the caller keeps track (functional.py does). */
int da_roi_align_backward_prepared(const void* grad_out, int grad_dtype, int out_layout,
                          const float* rois, int R, int pooled_h, int pooled_w,
                          float spatial_scale, int sampling_ratio, int aligned,
                          void* grad_in_nhwc, int grad_in_dtype, int N, int C, int H, int W,
                          void* workspace, size_t workspace_bytes, da_stream_t stream);
/* FPN level mapping, single_level_roi_extractor.py:36-55. levels_out int32 [R]. */
int da_map_roi_levels(const float* rois, int R, int num_levels, float finest_scale,
                      int32_t* levels_out, da_stream_t stream);

/* ---- RPN proposal stage ------------------------------
 * Replaces RPNHeadDA._get_bboxes_single + _bbox_post_process (mmdet/models/dense_heads/rpn_head_da.py:170-303) for one image
 * and one level: sigmoid scores in the reference's (H,W,A) order, DeltaXYWHBBoxCoder.decode = delta2bbox
 * (mmdet/core/bbox/coder/delta_xywh_bbox_coder.py:224-259), the min_bbox_size filter, mmcv.ops.batched_nms (single level =
 * plain NMS: a box is dropped when its IoU with a kept, higher-ranked box is > iou_thr) and the first max_out survivors.
 *   da_rpn_scores    cls [A,H*W] fp32 logits (the conv output as it lies) -> scores [H*W*A], index = cell*A + a.
 *   (ranking: the caller sorts scores descending, stable, and passes the first n = min(nms_pre, H*W*A) indices + scores)
 *   da_rpn_proposals reg [4A,H,W] fp32 deltas as they lie; base_anchors [A,4] (AnchorGenerator.base_anchors of the level),
 *                    anchor(idx) = base[idx % A] + stride * (cell % W, cell / W, ...) is computed, never stored;
 *                    means4 / stds4 / max_ratio = |log(wh_ratio_clip)| / (img_h, img_w) = max_shape / min_size < 0 disables
 *                    the size filter.  dets [max_out,5] = (x1,y1,x2,y2,score) in rank order, zero padded; *count (device
 *                    int32) = number of valid rows; keep_idx (nullable, int32 [max_out]) = rank of each kept box, -1 padded.
 * Nothing is read back by the host.  Workspace: da_rpn_proposals_workspace_bytes(n) (decoded boxes, validity flags and the
 * n x ceil(n/64) suppression bitmap); da_rpn_proposals_peek copies the decoded boxes / flags of the last call out of it. */
int da_rpn_scores(const float* cls, int A, int HW, float* scores, da_stream_t stream);
size_t da_rpn_proposals_workspace_bytes(int n);
int da_rpn_proposals(const float* reg, int A, int H, int W, const float* base_anchors, float stride,
                     const int64_t* top_idx, const float* top_scores, int n,
                     const float* means4, const float* stds4, float max_ratio, float img_h, float img_w, float min_size,
                     float iou_thr, int max_out, float* dets, int32_t* count, int32_t* keep_idx,
                     void* workspace, size_t workspace_bytes, da_stream_t stream);
int da_rpn_proposals_peek(const void* workspace, int n, float* boxes_out, unsigned char* valid_out, da_stream_t stream);

/* ---- domain losses -------------------------------
 * Every forward writes fp32 scalars on the device; every backward takes the upstream
 * gradient as a DEVICE scalar pointer (nullable = 1) times a host scalar `scale`
 * (lambda weights, and -1 of the GRL when the caller folds it here). */

/* L1/L2 pixel loss: resnet_da_daf_org.py:816-822 (whole_batch=1) and
 * resnet_da_cbam.py:971-979 (whole_batch=0).  logits [N,L] fp32, domain int32 [N].
 * partial: workspace of da_pixel_loss_workspace_bytes(N,L). */
size_t da_pixel_loss_workspace_bytes(int N, int64_t L);
int da_pixel_domain_loss_forward(const float* logits, int N, int64_t L, const int32_t* domain,
                                 int whole_batch, float* loss_out, void* workspace,
                                 size_t workspace_bytes, da_stream_t stream);
int da_pixel_domain_loss_backward(const float* logits, int N, int64_t L, const int32_t* domain,
                                  int whole_batch, const float* grad_loss, float scale,
                                  float* dlogits, da_stream_t stream);

/* L3/L4 cross-entropy over 2 classes, mean over rows: nn.CrossEntropyLoss at
 * resnet_da_cbam.py:966-968 (raw logits, on_sigmoid=0) and resnet_da.py:846-848 /
 * DAFaster_rcnn_Orig.py:177-188 (applied to sigmoid outputs, on_sigmoid=1, Q4).
 * z [R,2] raw logits; pred_out (nullable) receives sigmoid(z) when on_sigmoid. */
int da_ce2_forward(const float* z, const int32_t* labels, int R, int on_sigmoid,
                   float* pred_out, float* loss_out, da_stream_t stream);
int da_ce2_backward(const float* z, const int32_t* labels, int R, int on_sigmoid,
                    const float* grad_loss, float scale, const float* grad_pred,
                    float* dz, da_stream_t stream);

/* L6 sigmoid focal loss, mean over k*2: mmdet/models/losses/focal_loss.py:12-57,60-103
 * (mmcv sigmoid_focal_loss_forward/backward). u [k,2], labels in {0,1}. */
int da_focal2_forward(const float* u, const int32_t* labels, int k, float gamma, float alpha,
                      float* loss_out, da_stream_t stream);
int da_focal2_backward(const float* u, const int32_t* labels, int k, float gamma, float alpha,
                       const float* grad_loss, float scale, float* du, da_stream_t stream);

/* L7 consistency regulariser: DAFaster_rcnn_Orig.py:161-175.
 * m = mean(sigmoid(img_logits[0..n_img))); loss = sum_r |m - sigmoid(ins_pred[r, label_r])|.
 * mean_out: device scalar (kept for backward). */
int da_consistency_forward(const float* img_logits, int64_t n_img, const float* ins_pred,
                           const int32_t* labels, int R, float* mean_out, float* loss_out,
                           da_stream_t stream);
int da_consistency_backward(const float* img_logits, int64_t n_img, const float* ins_pred,
                            const int32_t* labels, int R, const float* mean_in,
                            const float* grad_loss, float scale,
                            float* d_img_logits, float* d_ins_pred, da_stream_t stream);

/* ---- 1-channel head tail (terminal conv of ImgAlignmentHead / LocalAlignmentHead) ------
 * resnet_da_daf_org.py:125,131 (conv2 512->1 + bias + ReLU) and resnet_da_cbam.py:87,112
 * (conv3 C->1, no bias).  x [M,K] NHWC rows (f32 or bf16), w [K] f32.
 * logits[m] = act(dot(x[m,:], w) + bias). */
int da_pixel_head_forward(const void* x, int x_dtype, int64_t M, int K, const float* w,
                          const float* bias, int relu, float* logits, da_stream_t stream);
/* dl'[m] = dlogits[m] masked by (post_relu_logits[m] > 0) when post_relu_logits != NULL;
 * dx[m,k] = dl'[m] * w[k] (nullable), dw[k] = sum_m dl'[m]*x[m,k], dbias = sum_m dl'[m].
 * workspace: da_pixel_head_workspace_bytes(M,K). */
size_t da_pixel_head_workspace_bytes(int64_t M, int K);
int da_pixel_head_backward(const void* x, int x_dtype, int64_t M, int K, const float* w,
                           const float* dlogits, const float* post_relu_logits,
                           void* dx, int dx_dtype, float* dw, float* dbias,
                           void* workspace, size_t workspace_bytes, da_stream_t stream);

/* ---- dense contractions: domain-classifier convs and FC stacks ------------------------
 * One implicit-GEMM for every conv of the DA heads (nn.Conv2d at resnet_da_daf_org.py:124-125,
 * resnet_da_cbam.py:83-87,123-146, resnet_da.py:89-91, roi_heads/local_da.py:56-61) and every
 * nn.Linear (instance_da.py:52-56,111-113; H=W=KH=KW=1).
 *   x  [N,H,W,Cin]  NHWC, dtype x_dtype
 *   w  [Cout,KH,KW,Cin] ("OHWI" = torch channels_last weight), dtype x_dtype
 *   y  [N,OH,OW,Cout], OH=(H+2*pad-KH)/stride+1
 * Epilogue: v = acc*scale[c] + shift[c] (nullable: 1 / 0; bias or folded eval-mode BN, Q9);
 *           if relu: v=max(v,0); if drop_p>0: v = keep(seed,idx) ? v/(1-drop_p) : 0.
 * The dropout keep-mask is a stateless counter hash of (seed, element index): backward
 * regenerates it, da_dropout_mask exports it for the oracle. */
typedef struct da_conv_desc {
  int N, H, W, Cin;
  int Cout, KH, KW;
  int stride, pad;
  int engine;   /* enum da_engine */
  int x_dtype;  /* dtype of x, w (and of dy in backward) */
  int y_dtype;  /* dtype of y (forward) / dx (backward data) */
} da_conv_desc;

size_t da_conv_workspace_bytes(const da_conv_desc* d);
int da_conv_forward(const da_conv_desc* d, const void* x, const void* w,
                    const float* scale, const float* shift, int relu,
                    float drop_p, uint64_t drop_seed,
                    void* y, void* workspace, size_t workspace_bytes, da_stream_t stream);
/* dz = dy * act'(y): fuses the activation derivative of THIS layer's forward epilogue
 * (relu mask from y, dropout mask regenerated, BN scale) into one pass.  With dv the
 * gradient w.r.t. v = acc*scale+shift it also reduces (both nullable)
 *   dshift[c] = sum_m dv[m,c]            (bias / BN beta gradient)
 *   dvdot[c]  = sum_m dv[m,c] * v[m,c]   (BN gamma gradient = (dvdot - beta*dshift)/gamma)
 * dy, y and dz share dtype y_dtype == x_dtype.  Workspace: da_conv_workspace_bytes(d). */
int da_conv_act_backward(const da_conv_desc* d, const void* dy, const void* y,
                         const float* scale, int relu, float drop_p, uint64_t drop_seed,
                         void* dz, float* dshift, float* dvdot, void* workspace,
                         size_t workspace_bytes, da_stream_t stream);
/* dx = out_scale * conv_transpose(dz, w).  out_scale carries the GRL weight (-lambda) when
 * x is the head input, so the reversed gradient is emitted in the same pass. */
int da_conv_backward_data(const da_conv_desc* d, const void* dz, const void* w,
                          float out_scale, void* dx, void* workspace, size_t workspace_bytes,
                          da_stream_t stream);
/* dw [Cout,KH,KW,Cin] fp32 = sum over pixels dz^T x */
int da_conv_backward_weight(const da_conv_desc* d, const void* x, const void* dz,
                            float* dw, void* workspace, size_t workspace_bytes,
                            da_stream_t stream);
/* Weight gradient FUSED with the optimizer step of that weight (one GPU: no gradient exchange in between): the
 * epilogue of the tcgen05 weight-gradient kernel applies the SGD rule of da_sgd_step (same operation order, bit-identical
 * results) to master / momentum / bf16 operand copy at each tile's addresses; the gradient itself is never written.
 * Replaces  layer backward (torch autograd) -> torch.optim.SGD.step (mmdet/apis/train.py:127) for one weight: 18 B per
 * parameter of HBM traffic instead of 4 + 22.  The tensors use the memory order of dw ([Cout,KH,KW,Cin]); Cin % 32 == 0;
 * call it AFTER the layer's data gradient (the operand copy is rewritten). */
typedef struct da_sgd_fuse {
  float* w;              /* fp32 master */
  float* momentum_buf;   /* fp32 momentum */
  void* w_bf16;          /* bf16 operand copy (NULL = none) */
  float lr, momentum, weight_decay;
  int32_t first_step;    /* 1 = momentum buffer uninitialised */
} da_sgd_fuse;
int da_conv_backward_weight_sgd(const da_conv_desc* d, const void* x, const void* dz, const da_sgd_fuse* sgd,
                                void* workspace, size_t workspace_bytes, da_stream_t stream);
int da_dropout_mask(uint64_t seed, int64_t n, float drop_p, uint8_t* keep_out, da_stream_t stream);
/* Optional device-resident uint64 step counter added to every dropout seed at kernel run time (NULL to
 * disable): lets a CUDA-graph replay of a captured train step draw fresh masks. */
int da_set_dropout_counter(const void* counter_dev);

/* global average pool over H*W (F.avg_pool2d(x,(H,W)) at resnet_da_cbam.py:186, resnet_da.py:101):
 * x [N,H,W,C] -> y [N,C] fp32; backward broadcasts dy/(H*W). */
size_t da_global_avgpool_workspace_bytes(int N, int C);
int da_global_avgpool_forward(const void* x, int x_dtype, int N, int HW, int C, float* y,
                              void* workspace, size_t workspace_bytes, da_stream_t stream);
int da_global_avgpool_backward(const float* dy, int N, int HW, int C, void* dx, int dx_dtype,
                               da_stream_t stream);

/* NonLocalBlock attention core (instance_da.py:150-192, resnet_da_deep.py:402-445):
 * softmax over the QUERY axis (nn.Softmax(dim=1) on [b,q,k], Q11).  s [T,T] fp32 row-major
 * with s[q,k]; p[q,k] = exp(s[q,k]) / sum_q' exp(s[q',k]).  backward: ds from dp and p. */
int da_softmax_dim0_forward(const float* s, int T, int ldk, float* p, da_stream_t stream);
int da_softmax_dim0_backward(const float* p, const float* dp, int T, int ldk, float* ds,
                             da_stream_t stream);

/* Blocked form of the same normalisation (NonLocalAlignmentHead at 1024x2048, T = 32768 tokens,
 * resnet_da_deep.py:122-164,402-445): the softmax runs over the QUERY axis, so every key column is independent and the
 * T x T matrix is processed one KEY BLOCK at a time.  s [Tq,Tk] fp32 row-major (row stride ld) = theta . phi_block^T.
 *   forward : stats[0:Tk] = column max, stats[Tk:2Tk] = 1 / column sum of exp(s - max) (computed unless have_stats != 0, in
 *             which case they are read: the backward re-creates p from a recomputed s without a second statistics pass);
 *             p[q,k] = exp(s[q,k] - max_k) / sum_k  in p_dtype (DA_F32 | DA_BF16 = the next GEMM's operand dtype), same ld.
 *   backward: ds[q,k] = p[q,k] * (dp[q,k] - sum_q' p[q',k] dp[q',k]),  dp fp32, ds in ds_dtype.
 * Two passes over the block each (per-row-chunk partials merged in a fixed order: deterministic, no atomics).
 * Workspace: da_colsoftmax_workspace_bytes(Tq,Tk), caller-owned. */
size_t da_colsoftmax_workspace_bytes(int Tq, int Tk);
int da_colsoftmax_forward(const float* s, int Tq, int Tk, int ld, void* p, int p_dtype, float* stats, int have_stats,
                          void* workspace, size_t workspace_bytes, da_stream_t stream);
int da_colsoftmax_backward(const void* p, int p_dtype, const float* dp, int Tq, int Tk, int ld, void* ds, int ds_dtype,
                           void* workspace, size_t workspace_bytes, da_stream_t stream);

/* W1: the lambda-weighted DA loss entries AND their total in one launch (detectors/DAFaster_rcnn_Orig.py:143-157 weights,
 * detectors/base.py:176-219 sum): scaled[i] = w[i] * *losses[i], total = sum_i scaled[i] (fixed order).  `losses_host` is a HOST
 * array of n device pointers, `weights_host` a host array of n floats (n <= DA_MAX_WEIGHTED).  Backward: d losses[i] =
 * w[i] * (grad_total + grad_scaled[i]) (either gradient nullable). */
#define DA_MAX_WEIGHTED 8
int da_weighted_sum_forward(const float* const* losses_host, const float* weights_host, int n, float* scaled, float* total,
                            da_stream_t stream);
int da_weighted_sum_backward(const float* weights_host, int n, const float* grad_total, const float* grad_scaled, float* d_losses,
                             da_stream_t stream);

/* ---- pixel-level domain classifier: producing conv -> terminal 1-channel conv -> per-pixel loss -> mean (north_star kernel 1) ----
 * ImgAlignmentHead (resnet_da_daf_org.py:120-146) + L1 (:816-822); LocalAlignmentHead (resnet_da_cbam.py:77-115) + L2 (:971-979);
 * plus the plain per-pixel sigmoid-BCE / focal modes north_star names (parity: F.binary_cross_entropy_with_logits,
 * py_sigmoid_focal_loss losses/focal_loss.py:12-57).  The producing conv is the implicit GEMM of da_conv_forward (GRL, bias /
 * folded BN, ReLU, dropout in its epilogue); the TAIL is one kernel: logit[m] = act(y[m,:].w + b), per-pixel loss term by
 * `mode`, deterministic mean (per-block partials, the last block sums them in block order) -> loss scalar on the device.
 * Backward tail = one kernel + one small reduction: d loss / d logit (times grad_loss * loss_scale, plus grad_logits from
 * other consumers of the logits, e.g. the consistency loss), ReLU mask of the logit, terminal conv's weight / bias gradient,
 * AND the activation derivative of the producing layer (ReLU / dropout taken from the stored y, folded-BN scale): it writes
 * dz = d loss / d accumulator of the producing conv, ready for da_conv_backward_weight / _data (GRL weight = out_scale). */
enum da_pixel_loss_mode {
  DA_PIXEL_LOSS_DAF_SQ_BATCH = 0,  /* L1: 0.5*mean(sigmoid(p)^2) | 0.5*mean(sigmoid(1-p)^2), means over the WHOLE batch per slot (Q5) */
  DA_PIXEL_LOSS_DAF_SQ_IMAGE = 1,  /* L2: the same integrands, mean per image */
  DA_PIXEL_LOSS_BCE = 2,           /* sigmoid BCE against the image's domain label, mean over all pixels */
  DA_PIXEL_LOSS_FOCAL = 3          /* sigmoid focal loss (gamma, alpha) against the domain label, mean over all pixels */
};
typedef struct da_pixel_tail {
  const float* w;         /* [K] fp32 weights of the terminal 1-channel conv (K = Cout of the producing conv) */
  const float* bias;      /* [1] or NULL */
  int relu;               /* ReLU on the logit (ImgAlignmentHead: resnet_da_daf_org.py:131) */
  int mode;               /* enum da_pixel_loss_mode */
  float gamma, alpha;     /* focal */
  const int32_t* domain;  /* [N] 0 = source, 1 = target */
} da_pixel_tail;
/* `d` describes the PRODUCING conv (x -> y); logits [N,OH,OW] fp32; loss fp32 scalar. */
size_t da_grl_conv_loss_workspace_bytes(const da_conv_desc* d);
int da_pixel_tail_forward(const da_conv_desc* d, const void* y, const da_pixel_tail* tail, float* logits, float* loss,
                          void* workspace, size_t workspace_bytes, da_stream_t stream);
int da_pixel_tail_backward(const da_conv_desc* d, const void* y, const da_pixel_tail* tail, const float* logits,
                           const float* grad_loss, float loss_scale, const float* grad_logits, const float* scale, int act_relu,
                           float drop_p, void* dz, float* dw_tail, float* dbias_tail, float* dshift, float* dvdot,
                           void* workspace, size_t workspace_bytes, da_stream_t stream);
/* Composite: conv (+ fused epilogue) -> tail; tail -> weight gradient -> data gradient * grl.  dx / dw nullable. */
int da_grl_conv_loss_forward(const da_conv_desc* d, const void* x, const void* w, const float* scale, const float* shift,
                             int relu, float drop_p, uint64_t drop_seed, void* y, const da_pixel_tail* tail, float* logits,
                             float* loss, void* workspace, size_t workspace_bytes, da_stream_t stream);
int da_grl_conv_loss_backward(const da_conv_desc* d, const void* x, const void* w, const float* scale, int relu, float drop_p,
                              const void* y, const da_pixel_tail* tail, const float* logits, const float* grad_loss,
                              float loss_scale, const float* grad_logits, float grl, void* dx, float* dw, float* dshift,
                              float* dvdot, float* dw_tail, float* dbias_tail, void* dz_scratch, void* workspace,
                              size_t workspace_bytes, da_stream_t stream);

/* ---- instance-level domain classifier + its loss as ONE persistent kernel (forward) / ONE (backward) ----------------------
 * InstanceAlignmentHead (nlb = 1: GRL -> NonLocalBlock over the R RoIs -> FC C-H1-H2-2 -> sigmoid, mmdet/models/roi_heads/
 * instance_da.py:42-101,150-192) or InstanceAlignmentHead_DAF (nlb = 0, :103-148) fused with the CE-on-sigmoid instance loss
 * (detectors/DAFaster_rcnn_Orig.py:177-188): north_star kernel (3).  bf16 operands (the tcgen05 engine), fp32 accumulation,
 * fp32 logits / predictions / loss / weight gradients.  The kernel is a program of GEMM tiles and elementwise ops separated
 * by grid-wide barriers (csrc/chain.cu); every tensor, saved activation and scratch buffer belongs to the caller.
 * Dropout (p after fc1 and fc2) uses the same counter hash as da_conv_forward (keep(seed, row*width + col)); backward takes
 * the ReLU/dropout derivative from the stored post-activation (a clamped or dropped unit stored exactly 0). */
typedef struct da_instance_fc_desc {
  int R, C, I, H1, H2;   /* RoIs; feature width (1024); NonLocalBlock inner width (C/2, ignored when !nlb); FC widths */
  int nlb;               /* 1: NonLocalBlock in front of the FC stack */
  float drop_p;          /* 0 = eval */
  uint64_t seed1, seed2; /* dropout seeds of fc1 / fc2 */
  float grl;             /* gradient-reversal weight folded into dx (backward), instance_da.py:20-23 */
  /* Optional feeding layer in the same kernel (the last shared FC of the bbox head, convfc_bbox_head.py:229-237, which produces
   * the features the head reads): x = relu(xin * w0^T + b0).  C0 = width of xin, 0 = no such layer (x is an input).
   * gate_in = 1: xin is itself a post-ReLU activation of the caller's previous layer; backward then emits dxin already
   * multiplied by that layer's ReLU derivative (xin > 0) and its bias gradient db_in = colsum(dxin). */
  int C0, gate_in;
} da_instance_fc_desc;
typedef struct da_instance_fc_tensors {
  const void* x;         /* bf16 [R,C]; an OUTPUT of forward (saved activation) when desc.C0 > 0 */
  const void* w_proj;    /* bf16 [3I,C]: conv_theta | conv_phi | conv_g weights in ONE buffer (nlb) */
  const void* w_mask;    /* bf16 [C,I] (nlb) */
  const void* w1; const float* b1;   /* bf16 [H1,C], f32 [H1] */
  const void* w2; const float* b2;   /* bf16 [H2,H1], f32 [H2] */
  const void* w3; const float* b3;   /* bf16 [2,H2], f32 [2] */
  const int32_t* labels; /* [R], 0 = source, 1 = target */
  /* saved activations: written by forward, read by backward */
  void* proj;            /* bf16 [R,3I]  theta | phi | g (nlb) */
  void* attn;            /* bf16 [R, (R+7)&~7] softmax over the QUERY axis (nlb) */
  void* y;               /* bf16 [R,I] (nlb) */
  void* t;               /* bf16 [R,C] NonLocalBlock output (nlb) */
  void* h1; void* h2;    /* bf16 [R,H1], [R,H2], after ReLU and dropout */
  float* z;              /* f32 [R,2] raw logits of fc3 */
  float* pred;           /* f32 [R,2] sigmoid(z): what the reference head returns */
  float* loss;           /* f32 scalar: mean CE(pred, labels) */
  const void* xin; const void* w0; const float* b0;   /* feeding layer (desc.C0 > 0): bf16 [R,C0], bf16 [C,C0], f32 [C] */
} da_instance_fc_tensors;
typedef struct da_instance_fc_grads {
  const float* grad_loss;   /* device scalar (NULL = 1) */
  float loss_scale;         /* host scalar multiplied in (lambda) */
  const float* grad_pred;   /* f32 [R,2] gradient w.r.t. pred from other consumers (the consistency loss), or NULL */
  void* dx;                 /* bf16 [R,C], already multiplied by desc.grl */
  float* dw_proj;           /* f32 [3I,C] (nlb) */
  float* dw_mask;           /* f32 [C,I] (nlb) */
  float* dw1; float* db1; float* dw2; float* db2; float* dw3; float* db3;
  void* dz2; void* dz1;     /* scratch bf16 [R,H2], [R,H1] */
  void* dt; void* dy; void* dproj;   /* scratch bf16 [R,C], [R,I], [R,3I] (nlb) */
  /* feeding layer (desc.C0 > 0): dx then holds the gradient w.r.t. the layer's PRE-activation (dx * (x > 0)) */
  void* dxin;               /* bf16 [R,C0] */
  float* dw0; float* db0;   /* f32 [C,C0], [C] */
  float* db_in;             /* f32 [C0] (gate_in) */
} da_instance_fc_grads;
size_t da_instance_fc_workspace_bytes(int R);
int da_instance_fc_forward(const da_instance_fc_desc* d, const da_instance_fc_tensors* t, void* workspace, size_t workspace_bytes,
                           da_stream_t stream);
int da_instance_fc_backward(const da_instance_fc_desc* d, const da_instance_fc_tensors* t, const da_instance_fc_grads* g,
                            void* workspace, size_t workspace_bytes, da_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DA_B200_H_ */
