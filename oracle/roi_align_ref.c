/*
 * ORACLE — test infrastructure only.  Nothing under oracle/ is imported, linked or executed by
 * the product path (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it).
 *
 * CPU restatement of RoIAlign (avg mode) as the DA path calls it:
 *   /root/reference/mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:54-60 builds
 *   mmcv.ops.RoIAlign(spatial_scale=1/16, output_size=7, sampling_ratio=0) (aligned=True, avg),
 *   single_level_roi_extractor.py:79 calls it.
 * The arithmetic lives in the un-vendored dependency mmcv-full==1.3.17
 * (mmcv/ops/csrc/pytorch/cpu/roi_align.cpp, same Detectron2 lineage as torchvision's
 * roi_align_kernel.cpp); it is restated here from the published algorithm (SURVEY.md
 * Appendix A) and PINNED against torchvision's C++ CPU operator
 * torch.ops.torchvision.roi_align (tests/test_oracle_cpu.py, tests/golden/roi_align_*.pt).
 * Build with -ffp-contract=off: the pinned CPU operator evaluates the sample coordinates
 * without FMA contraction.
 *
 * Layout: feat [N,C,H,W], rois [R,5] (batch_ind,x1,y1,x2,y2), out [R,C,ph,pw].
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct { int ok; int yl, yh, xl, xh; float w1, w2, w3, w4; } tap_t;

static tap_t bilinear_taps(float y, float x, int H, int W) {
  tap_t t; memset(&t, 0, sizeof(t));
  if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) return t; /* contributes 0 */
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else { yh = yl + 1; }
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else { xh = xl + 1; }
  float ly = y - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
  t.ok = 1; t.yl = yl; t.yh = yh; t.xl = xl; t.xh = xh;
  t.w1 = hy * hx; t.w2 = hy * lx; t.w3 = ly * hx; t.w4 = ly * lx;
  return t;
}

typedef struct { float x1, y1, bin_w, bin_h; int gh, gw, b; float count; } roi_geom_t;

static roi_geom_t roi_geom(const float* roi, int ph, int pw, float scale, int sr, int aligned) {
  roi_geom_t g;
  float off = aligned ? 0.5f : 0.f;
  g.b = (int)roi[0];
  g.x1 = roi[1] * scale - off; g.y1 = roi[2] * scale - off;
  float x2 = roi[3] * scale - off, y2 = roi[4] * scale - off;
  float rw = x2 - g.x1, rh = y2 - g.y1;
  if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
  g.bin_h = rh / (float)ph; g.bin_w = rw / (float)pw;
  g.gh = sr > 0 ? sr : (int)ceilf(rh / (float)ph);
  g.gw = sr > 0 ? sr : (int)ceilf(rw / (float)pw);
  int c = g.gh * g.gw; g.count = (float)(c > 1 ? c : 1);
  return g;
}

/* grid (nullable): int32 [R,2] = (roi_bin_grid_h, roi_bin_grid_w); bidx (nullable): int32 [R] */
/* [r0,r1): RoI range, so that a host thread pool can split the work (ctypes drops the GIL). */
void roi_align_forward_ref_range(const float* feat, int N, int C, int H, int W, const float* rois,
                                 int r0, int r1, int ph, int pw, float scale, int sr, int aligned,
                                 float* out, int32_t* grid, int32_t* bidx) {
  for (int r = r0; r < r1; ++r) {
    roi_geom_t g = roi_geom(rois + 5 * r, ph, pw, scale, sr, aligned);
    if (grid) { grid[2 * r] = g.gh; grid[2 * r + 1] = g.gw; }
    if (bidx) bidx[r] = g.b;
    for (int c = 0; c < C; ++c) {
      float* o = out + ((size_t)r * C + c) * ph * pw;
      if (g.b < 0 || g.b >= N) { memset(o, 0, sizeof(float) * ph * pw); continue; } /* Q1: bounds-checked */
      const float* f = feat + ((size_t)g.b * C + c) * H * W;
      for (int i = 0; i < ph; ++i)
        for (int j = 0; j < pw; ++j) {
          float acc = 0.f;
          for (int iy = 0; iy < g.gh; ++iy) {
            const float y = g.y1 + i * g.bin_h + (iy + .5f) * g.bin_h / (float)g.gh;
            for (int ix = 0; ix < g.gw; ++ix) {
              const float x = g.x1 + j * g.bin_w + (ix + .5f) * g.bin_w / (float)g.gw;
              tap_t t = bilinear_taps(y, x, H, W);
              if (!t.ok) continue;
              acc += t.w1 * f[t.yl * W + t.xl] + t.w2 * f[t.yl * W + t.xh] +
                     t.w3 * f[t.yh * W + t.xl] + t.w4 * f[t.yh * W + t.xh];
            }
          }
          o[i * pw + j] = acc / g.count;
        }
    }
  }
}

void roi_align_forward_ref(const float* feat, int N, int C, int H, int W, const float* rois, int R,
                           int ph, int pw, float scale, int sr, int aligned, float* out,
                           int32_t* grid, int32_t* bidx) {
  roi_align_forward_ref_range(feat, N, C, H, W, rois, 0, R, ph, pw, scale, sr, aligned, out, grid, bidx);
}

/* gin [N,C,H,W]: accumulation in double then rounded (the reference scatters with float atomics in an
 * unspecified order, so its own result is only defined to rounding).  [c0,c1): channel range, so that a host
 * thread pool can split the work without atomics (every thread owns whole channel planes); the caller zeroes gin. */
void roi_align_backward_ref_range(const float* gout, const float* rois, int R, int ph, int pw, float scale,
                                  int sr, int aligned, double* gin, int N, int C, int H, int W, int c0, int c1) {
  for (int r = 0; r < R; ++r) {
    roi_geom_t g = roi_geom(rois + 5 * r, ph, pw, scale, sr, aligned);
    if (g.b < 0 || g.b >= N) continue;
    for (int c = c0; c < c1; ++c) {
      double* gi = gin + ((size_t)g.b * C + c) * H * W;
      const float* go = gout + ((size_t)r * C + c) * ph * pw;
      for (int i = 0; i < ph; ++i)
        for (int j = 0; j < pw; ++j) {
          const float gv = go[i * pw + j];
          for (int iy = 0; iy < g.gh; ++iy) {
            const float y = g.y1 + i * g.bin_h + (iy + .5f) * g.bin_h / (float)g.gh;
            for (int ix = 0; ix < g.gw; ++ix) {
              const float x = g.x1 + j * g.bin_w + (ix + .5f) * g.bin_w / (float)g.gw;
              tap_t t = bilinear_taps(y, x, H, W);
              if (!t.ok) continue;
              gi[t.yl * W + t.xl] += (double)(gv * t.w1 / g.count);
              gi[t.yl * W + t.xh] += (double)(gv * t.w2 / g.count);
              gi[t.yh * W + t.xl] += (double)(gv * t.w3 / g.count);
              gi[t.yh * W + t.xh] += (double)(gv * t.w4 / g.count);
            }
          }
        }
    }
  }
}

void roi_align_backward_ref(const float* gout, const float* rois, int R, int ph, int pw, float scale,
                            int sr, int aligned, double* gin, int N, int C, int H, int W) {
  memset(gin, 0, sizeof(double) * (size_t)N * C * H * W);
  roi_align_backward_ref_range(gout, rois, R, ph, pw, scale, sr, aligned, gin, N, C, H, W, 0, C);
}

/* The same scatter with fp32 accumulation (what the reference's CPU / CUDA kernels do: `+=` / atomicAdd on float):
 * the arithmetic bench.py's CPU baseline times.  Channel range as above; the caller zeroes gin. */
void roi_align_backward_f32_range(const float* gout, const float* rois, int R, int ph, int pw, float scale,
                                  int sr, int aligned, float* gin, int N, int C, int H, int W, int c0, int c1) {
  for (int r = 0; r < R; ++r) {
    roi_geom_t g = roi_geom(rois + 5 * r, ph, pw, scale, sr, aligned);
    if (g.b < 0 || g.b >= N) continue;
    for (int c = c0; c < c1; ++c) {
      float* gi = gin + ((size_t)g.b * C + c) * H * W;
      const float* go = gout + ((size_t)r * C + c) * ph * pw;
      for (int i = 0; i < ph; ++i)
        for (int j = 0; j < pw; ++j) {
          const float gv = go[i * pw + j];
          for (int iy = 0; iy < g.gh; ++iy) {
            const float y = g.y1 + i * g.bin_h + (iy + .5f) * g.bin_h / (float)g.gh;
            for (int ix = 0; ix < g.gw; ++ix) {
              const float x = g.x1 + j * g.bin_w + (ix + .5f) * g.bin_w / (float)g.gw;
              tap_t t = bilinear_taps(y, x, H, W);
              if (!t.ok) continue;
              gi[t.yl * W + t.xl] += gv * t.w1 / g.count;
              gi[t.yl * W + t.xh] += gv * t.w2 / g.count;
              gi[t.yh * W + t.xl] += gv * t.w3 / g.count;
              gi[t.yh * W + t.xh] += gv * t.w4 / g.count;
            }
          }
        }
    }
  }
}

/* FPN level mapping, single_level_roi_extractor.py:36-55 (fp32). */
void map_roi_levels_ref(const float* rois, int R, int num_levels, float finest_scale, int32_t* out) {
  for (int r = 0; r < R; ++r) {
    const float* q = rois + 5 * r;
    float s = sqrtf((q[3] - q[1]) * (q[4] - q[2]));
    float l = floorf(log2f(s / finest_scale + 1e-6f));
    int lv = (l != l) ? 0 : (int)fminf(fmaxf(l, 0.f), (float)(num_levels - 1));
    out[r] = lv;
  }
}
