/*
 * abr_oracle.c -- CPU restatement of the ABR_IOD RoI hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker.  The product
 * (abr_iod_b200/) never falls back to it.
 *
 * Every function restates, in plain scalar C and in the reference's own
 * operation order, one function of /root/reference (paths below are relative
 * to maskrcnn_benchmark/).  All arithmetic is IEEE fp32 without contraction
 * (build with -ffp-contract=off, no -ffast-math), which is what the
 * reference's CPU build computes.
 *
 * Parity pins (tests/test_oracle_pins.py, tests/golden/):
 *   - orc_roi_align_fwd   == compiled csrc/cpu/ROIAlign_cpu.cpp (oracle/_ref), bit for bit
 *   - orc_nms (ge=1)      == compiled csrc/cpu/nms_cpu.cpp (oracle/_ref), index for index
 *   - orc_roi_align_bwd, orc_roi_pool_*: no CPU reference exists
 *     (csrc/ROIAlign.h:44, csrc/ROIPool.h:23,44); restated from the CUDA
 *     sources and cross-checked against torchvision CPU ops (third party).
 *   - orc_ard             vs. the imported reference Python
 *     distillation/distillation.py:86-130 and its autograd (golden fixtures)
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* ROIAlign                                                           */
/* ------------------------------------------------------------------ */

/* One bilinear sample: indices and weights.
 * Follows csrc/cuda/ROIAlign_cuda.cu:125-174 (bilinear_interpolate_gradient)
 * which is the same case analysis as csrc/cpu/ROIAlign_cpu.cpp:46-92. */
typedef struct {
  int y_low, y_high, x_low, x_high; /* -1 => sample is outside, contributes 0 */
  float w1, w2, w3, w4;
} orc_tap_t;

static void orc_bilinear_taps(int height, int width, float y, float x, orc_tap_t* t) {
  if (y < -1.0 || y > height || x < -1.0 || x > width) {
    t->w1 = t->w2 = t->w3 = t->w4 = 0.f;
    t->x_low = t->x_high = t->y_low = t->y_high = -1;
    return;
  }
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int y_low = (int)y, x_low = (int)x, y_high, x_high;
  if (y_low >= height - 1) {
    y_high = y_low = height - 1;
    y = (float)y_low;
  } else {
    y_high = y_low + 1;
  }
  if (x_low >= width - 1) {
    x_high = x_low = width - 1;
    x = (float)x_low;
  } else {
    x_high = x_low + 1;
  }
  float ly = y - y_low;
  float lx = x - x_low;
  float hy = (float)(1. - ly), hx = (float)(1. - lx);
  t->w1 = hy * hx; t->w2 = hy * lx; t->w3 = ly * hx; t->w4 = ly * lx;
  t->y_low = y_low; t->y_high = y_high; t->x_low = x_low; t->x_high = x_high;
}

/* RoI geometry shared by forward and backward.
 * csrc/cuda/ROIAlign_cuda.cu:78-104 == csrc/cpu/ROIAlign_cpu.cpp:139-170. */
typedef struct {
  int batch;
  float start_w, start_h, bin_h, bin_w;
  int grid_h, grid_w;
  float count;
} orc_roi_geom_t;

static void orc_roi_geom(const float* roi, float scale, int PH, int PW, int ratio, orc_roi_geom_t* g) {
  g->batch = (int)roi[0];
  float roi_start_w = roi[1] * scale;
  float roi_start_h = roi[2] * scale;
  float roi_end_w = roi[3] * scale;
  float roi_end_h = roi[4] * scale;
  float roi_width = fmaxf(roi_end_w - roi_start_w, 1.f);
  float roi_height = fmaxf(roi_end_h - roi_start_h, 1.f);
  g->bin_h = roi_height / (float)PH;
  g->bin_w = roi_width / (float)PW;
  g->grid_h = (ratio > 0) ? ratio : (int)ceilf(roi_height / PH);
  g->grid_w = (ratio > 0) ? ratio : (int)ceilf(roi_width / PW);
  g->count = (float)(g->grid_h * g->grid_w);
  g->start_w = roi_start_w;
  g->start_h = roi_start_h;
}

/* Forward.  in [B,C,H,W], rois [R,5], out [R,C,PH,PW], all contiguous fp32.
 * csrc/cpu/ROIAlign_cpu.cpp:114-219 (per-RoI precomputed taps, then channels). */
ORC_API void orc_roi_align_fwd(const float* in, const float* rois, float* out,
                               int B, int C, int H, int W, int R, int PH, int PW,
                               float scale, int ratio) {
  (void)B;
  for (int n = 0; n < R; n++) {
    orc_roi_geom_t g;
    orc_roi_geom(rois + 5 * n, scale, PH, PW, ratio, &g);
    size_t ntap = (size_t)g.grid_h * g.grid_w * PH * PW;
    orc_tap_t* taps = (orc_tap_t*)malloc(sizeof(orc_tap_t) * (ntap ? ntap : 1));
    size_t k = 0;
    for (int ph = 0; ph < PH; ph++)
      for (int pw = 0; pw < PW; pw++)
        for (int iy = 0; iy < g.grid_h; iy++) {
          const float yy = g.start_h + ph * g.bin_h + (float)(iy + .5f) * g.bin_h / (float)g.grid_h;
          for (int ix = 0; ix < g.grid_w; ix++) {
            const float xx = g.start_w + pw * g.bin_w + (float)(ix + .5f) * g.bin_w / (float)g.grid_w;
            orc_bilinear_taps(H, W, yy, xx, &taps[k++]);
          }
        }
    for (int c = 0; c < C; c++) {
      const float* plane = in + ((size_t)g.batch * C + c) * H * W;
      float* o = out + ((size_t)n * C + c) * PH * PW;
      k = 0;
      for (int ph = 0; ph < PH; ph++)
        for (int pw = 0; pw < PW; pw++) {
          float acc = 0.f;
          for (int s = 0; s < g.grid_h * g.grid_w; s++) {
            const orc_tap_t* t = &taps[k++];
            if (t->y_low < 0) { /* outside: reference adds 0*plane[0]; identical unless plane[0] is inf/nan */
              continue;
            }
            acc += t->w1 * plane[t->y_low * W + t->x_low] + t->w2 * plane[t->y_low * W + t->x_high] +
                   t->w3 * plane[t->y_high * W + t->x_low] + t->w4 * plane[t->y_high * W + t->x_high];
          }
          acc /= g.count;
          o[ph * PW + pw] = acc;
        }
    }
    free(taps);
  }
}

/* Backward.  gout [R,C,PH,PW] -> gin [B,C,H,W] (zero-filled here, like
 * at::zeros at csrc/cuda/ROIAlign_cuda.cu:316).  Follows the kernel
 * csrc/cuda/ROIAlign_cuda.cu:177-254 element by element in index order; the
 * reference's atomicAdd order is unspecified, so fp32 sums agree with it only
 * to rounding. */
ORC_API void orc_roi_align_bwd(const float* gout, const float* rois, float* gin,
                               int B, int C, int H, int W, int R, int PH, int PW,
                               float scale, int ratio) {
  memset(gin, 0, sizeof(float) * (size_t)B * C * H * W);
  for (int n = 0; n < R; n++) {
    orc_roi_geom_t g;
    orc_roi_geom(rois + 5 * n, scale, PH, PW, ratio, &g);
    for (int c = 0; c < C; c++) {
      float* plane = gin + ((size_t)g.batch * C + c) * H * W;
      const float* go = gout + ((size_t)n * C + c) * PH * PW;
      for (int ph = 0; ph < PH; ph++)
        for (int pw = 0; pw < PW; pw++) {
          const float top = go[ph * PW + pw];
          for (int iy = 0; iy < g.grid_h; iy++) {
            const float y = g.start_h + ph * g.bin_h + (float)(iy + .5f) * g.bin_h / (float)g.grid_h;
            for (int ix = 0; ix < g.grid_w; ix++) {
              const float x = g.start_w + pw * g.bin_w + (float)(ix + .5f) * g.bin_w / (float)g.grid_w;
              orc_tap_t t;
              orc_bilinear_taps(H, W, y, x, &t);
              float g1 = top * t.w1 / g.count;
              float g2 = top * t.w2 / g.count;
              float g3 = top * t.w3 / g.count;
              float g4 = top * t.w4 / g.count;
              if (t.x_low >= 0 && t.x_high >= 0 && t.y_low >= 0 && t.y_high >= 0) {
                plane[t.y_low * W + t.x_low] += g1;
                plane[t.y_low * W + t.x_high] += g2;
                plane[t.y_high * W + t.x_low] += g3;
                plane[t.y_high * W + t.x_high] += g4;
              }
            }
          }
        }
    }
  }
}

/* ------------------------------------------------------------------ */
/* ROIPool                                                            */
/* ------------------------------------------------------------------ */

/* csrc/cuda/ROIPool_cuda.cu:16-77.  argmax is int32, flat h*W+w in the plane. */
ORC_API void orc_roi_pool_fwd(const float* in, const float* rois, float* out, int32_t* argmax,
                              int B, int C, int H, int W, int R, int PH, int PW, float scale) {
  (void)B;
  for (int n = 0; n < R; n++) {
    const float* roi = rois + 5 * n;
    int b = (int)roi[0];
    int roi_start_w = (int)roundf(roi[1] * scale);
    int roi_start_h = (int)roundf(roi[2] * scale);
    int roi_end_w = (int)roundf(roi[3] * scale);
    int roi_end_h = (int)roundf(roi[4] * scale);
    int roi_width = roi_end_w - roi_start_w + 1; if (roi_width < 1) roi_width = 1;
    int roi_height = roi_end_h - roi_start_h + 1; if (roi_height < 1) roi_height = 1;
    float bin_size_h = (float)roi_height / (float)PH;
    float bin_size_w = (float)roi_width / (float)PW;
    for (int c = 0; c < C; c++) {
      const float* plane = in + ((size_t)b * C + c) * H * W;
      for (int ph = 0; ph < PH; ph++)
        for (int pw = 0; pw < PW; pw++) {
          int hstart = (int)floorf((float)ph * bin_size_h);
          int wstart = (int)floorf((float)pw * bin_size_w);
          int hend = (int)ceilf((float)(ph + 1) * bin_size_h);
          int wend = (int)ceilf((float)(pw + 1) * bin_size_w);
          hstart = hstart + roi_start_h; if (hstart < 0) hstart = 0; if (hstart > H) hstart = H;
          hend = hend + roi_start_h; if (hend < 0) hend = 0; if (hend > H) hend = H;
          wstart = wstart + roi_start_w; if (wstart < 0) wstart = 0; if (wstart > W) wstart = W;
          wend = wend + roi_start_w; if (wend < 0) wend = 0; if (wend > W) wend = W;
          int is_empty = (hend <= hstart) || (wend <= wstart);
          float maxval = is_empty ? 0.f : -FLT_MAX;
          int maxidx = -1;
          for (int h = hstart; h < hend; ++h)
            for (int w = wstart; w < wend; ++w) {
              int idx = h * W + w;
              if (plane[idx] > maxval) { maxval = plane[idx]; maxidx = idx; }
            }
          size_t o = (((size_t)n * C + c) * PH + ph) * PW + pw;
          out[o] = maxval;
          argmax[o] = maxidx;
        }
    }
  }
}

/* csrc/cuda/ROIPool_cuda.cu:79-108, sequential index order. */
ORC_API void orc_roi_pool_bwd(const float* gout, const int32_t* argmax, const float* rois, float* gin,
                              int B, int C, int H, int W, int R, int PH, int PW) {
  memset(gin, 0, sizeof(float) * (size_t)B * C * H * W);
  for (int n = 0; n < R; n++) {
    int b = (int)rois[5 * n];
    for (int c = 0; c < C; c++) {
      float* plane = gin + ((size_t)b * C + c) * H * W;
      size_t top = ((size_t)n * C + c) * PH * PW;
      for (int i = 0; i < PH * PW; i++) {
        int a = argmax[top + i];
        if (a != -1) plane[a] += gout[top + i];
      }
    }
  }
}

/* ------------------------------------------------------------------ */
/* NMS                                                                */
/* ------------------------------------------------------------------ */

typedef struct { float s; int64_t i; } orc_sv_t;

/* scores.sort(0, descending=True) (csrc/cuda/nms.cu:74, csrc/cpu/nms_cpu.cpp:24):
 * ties keep ascending index (stable); NaN sorts first, as torch does. */
static int orc_sv_cmp(const void* pa, const void* pb) {
  const orc_sv_t* a = (const orc_sv_t*)pa;
  const orc_sv_t* b = (const orc_sv_t*)pb;
  int an = a->s != a->s, bn = b->s != b->s;
  if (an != bn) return an ? -1 : 1;
  if (!an) {
    if (a->s > b->s) return -1;
    if (a->s < b->s) return 1;
  }
  return (a->i < b->i) ? -1 : (a->i > b->i);
}

/* Greedy NMS with the +1 pixel convention.
 *   ge == 0: CUDA flavour -- suppress when IoU >  thr (csrc/cuda/nms.cu:13-21,60,105-123)
 *   ge == 1: CPU  flavour -- suppress when IoU >= thr (csrc/cpu/nms_cpu.cpp:38-62)
 * keep receives the surviving ORIGINAL indices in ascending order
 * (csrc/cuda/nms.cu:127-130, csrc/cpu/nms_cpu.cpp:64).  Returns their count. */
ORC_API int64_t orc_nms(const float* boxes, const float* scores, int64_t n, float thr, int ge,
                        int64_t* keep) {
  if (n <= 0) return 0;
  orc_sv_t* sv = (orc_sv_t*)malloc(sizeof(orc_sv_t) * n);
  for (int64_t i = 0; i < n; i++) { sv[i].s = scores[i]; sv[i].i = i; }
  qsort(sv, n, sizeof(orc_sv_t), orc_sv_cmp);
  uint8_t* sup = (uint8_t*)calloc(n, 1);
  float* area = (float*)malloc(sizeof(float) * n);
  for (int64_t i = 0; i < n; i++) {
    const float* b = boxes + 4 * i;
    area[i] = (b[2] - b[0] + 1) * (b[3] - b[1] + 1);
  }
  for (int64_t _i = 0; _i < n; _i++) {
    int64_t i = sv[_i].i;
    if (sup[i]) continue;
    const float* a = boxes + 4 * i;
    for (int64_t _j = _i + 1; _j < n; _j++) {
      int64_t j = sv[_j].i;
      if (sup[j]) continue;
      const float* b = boxes + 4 * j;
      float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
      float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
      float w = fmaxf(right - left + 1, 0.f), h = fmaxf(bottom - top + 1, 0.f);
      float inter = w * h;
      float ovr = inter / (area[i] + area[j] - inter);
      if (ge ? (ovr >= thr) : (ovr > thr)) sup[j] = 1;
    }
  }
  int64_t k = 0;
  for (int64_t i = 0; i < n; i++)
    if (!sup[i]) keep[k++] = i;
  free(sv); free(sup); free(area);
  return k;
}

/* ------------------------------------------------------------------ */
/* Attentive RoI Distillation                                         */
/* ------------------------------------------------------------------ */

/* distillation/distillation.py:86-130 with the call-site argument order of
 * tools/train_incremental.py:115:  Fo = argument 0 (old model / teacher, no
 * grad), Fn = argument 1 (new model / student).
 *   m_x[n,hw] = mean_c Fx^2                       (:124-126)
 *   A_x       = HW * softmax_hw(m_x)              (:128)
 *   L_pad     = mean_{n,hw} |A_n - A_o|           (:114-118)
 *   L_afd     = mean_{n,c,hw} (Fo*sqrt(A_o) - Fn*sqrt(A_o))^2   (:103-111)
 *   L         = L_afd + gamma * L_pad             (:99)
 * Accumulates in double (a tighter answer than either fp32 implementation;
 * tests compare with rtol 1e-5).  loss3 = {L, L_afd, L_pad}.  dFn (may be
 * NULL) receives dL/dFn in fp32 -- the closed form of the reference's autograd:
 *   dFn = 2 A_o (Fn-Fo)/(N C HW) + (2 Fn / C) * HW * s_i (g_i - sum_j g_j s_j),
 *   s = softmax(m_n), g = gamma * sign(A_n - A_o) / (N HW).
 * Layout [N,C,HW] contiguous. */
ORC_API void orc_ard(const float* Fo, const float* Fn, int N, int C, int HW, double gamma,
                     double* loss3, float* dFn) {
  double* mo = (double*)malloc(sizeof(double) * HW);
  double* mn = (double*)malloc(sizeof(double) * HW);
  double* Ao = (double*)malloc(sizeof(double) * HW);
  double* An = (double*)malloc(sizeof(double) * HW);
  double* sn = (double*)malloc(sizeof(double) * HW);
  double* kk = (double*)malloc(sizeof(double) * HW);
  double afd = 0.0, pad = 0.0;
  const double inv_all = 1.0 / ((double)N * C * HW);
  for (int n = 0; n < N; n++) {
    const float* fo = Fo + (size_t)n * C * HW;
    const float* fn = Fn + (size_t)n * C * HW;
    for (int p = 0; p < HW; p++) { mo[p] = 0; mn[p] = 0; }
    for (int c = 0; c < C; c++)
      for (int p = 0; p < HW; p++) {
        double a = fo[(size_t)c * HW + p], b = fn[(size_t)c * HW + p];
        mo[p] += a * a; mn[p] += b * b;
      }
    double maxo = -INFINITY, maxn = -INFINITY;
    for (int p = 0; p < HW; p++) {
      mo[p] /= C; mn[p] /= C;
      if (mo[p] > maxo) maxo = mo[p];
      if (mn[p] > maxn) maxn = mn[p];
    }
    double so = 0, ssn = 0;
    for (int p = 0; p < HW; p++) {
      Ao[p] = exp(mo[p] - maxo); so += Ao[p];
      An[p] = exp(mn[p] - maxn); ssn += An[p];
    }
    double gs = 0;
    for (int p = 0; p < HW; p++) {
      Ao[p] = HW * (Ao[p] / so);
      sn[p] = An[p] / ssn;
      An[p] = HW * sn[p];
      double d = An[p] - Ao[p];
      pad += fabs(d);
      double g = gamma * ((d > 0) - (d < 0)) / ((double)N * HW);
      kk[p] = g;
      gs += g * sn[p];
    }
    for (int p = 0; p < HW; p++) kk[p] = (2.0 / C) * HW * sn[p] * (kk[p] - gs);
    for (int c = 0; c < C; c++)
      for (int p = 0; p < HW; p++) {
        size_t i = (size_t)c * HW + p;
        double d = (double)fn[i] - (double)fo[i];
        afd += Ao[p] * d * d;
        if (dFn) dFn[(size_t)n * C * HW + i] = (float)(2.0 * Ao[p] * d * inv_all + (double)fn[i] * kk[p]);
      }
  }
  afd *= inv_all;
  pad /= ((double)N * HW);
  loss3[1] = afd; loss3[2] = pad; loss3[0] = afd + gamma * pad;
  free(mo); free(mn); free(Ao); free(An); free(sn); free(kk);
}

/* ------------------------------------------------------------------ */
/* ABR paste (pixel part; coordinates are planned on the host)        */
/* ------------------------------------------------------------------ */

/* Mixup blend of one prototype crop into an image region, in place.
 * data/datasets/voc_abr.py:659-678:  region = Lambda*region + (1-Lambda)*crop
 * evaluated in float64 (numpy promotes uint8*python-float), two roundings for
 * the products and one for the sum, then the unsafe float64->uint8 cast
 * (truncation) of the numpy slice assignment.  img is HWC uint8 [H,W,3];
 * src is HWC uint8 [sh,sw,3]; the region is img[y0:y1, x0:x1] and the crop
 * starts at (sy0, sx0) inside src. */
ORC_API void orc_paste_mixup(uint8_t* img, int H, int W, const uint8_t* src, int sh, int sw,
                             int y0, int x0, int y1, int x1, int sy0, int sx0, double lambda) {
  (void)H; (void)sh;
  const double one_minus = 1 - lambda;
  for (int y = y0; y < y1; y++)
    for (int x = x0; x < x1; x++)
      for (int ch = 0; ch < 3; ch++) {
        size_t di = ((size_t)y * W + x) * 3 + ch;
        size_t si = ((size_t)(sy0 + y - y0) * sw + (sx0 + x - x0)) * 3 + ch;
        double a = lambda * (double)img[di];
        double b = one_minus * (double)src[si];
        double v = a + b;
        img[di] = (uint8_t)v; /* v in [0,255]: C truncation == numpy's cast */
      }
}

/* Mosaic copy of one prototype crop into the canvas.
 * data/datasets/voc_abr.py:744,763,804: canvas float32 filled with 114, region
 * overwritten with uint8 pixels, final np.uint8() -- i.e. a plain byte copy. */
ORC_API void orc_paste_copy(uint8_t* img, int H, int W, const uint8_t* src, int sh, int sw,
                            int y0, int x0, int y1, int x1, int sy0, int sx0) {
  (void)H; (void)sh;
  for (int y = y0; y < y1; y++)
    memcpy(img + ((size_t)y * W + x0) * 3, src + ((size_t)(sy0 + y - y0) * sw + sx0) * 3,
           (size_t)(x1 - x0) * 3);
}

ORC_API int orc_version(void) { return 1; }
