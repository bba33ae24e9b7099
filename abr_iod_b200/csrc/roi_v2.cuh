// roi_v2.cuh -- gather-form ROIAlign ("v2") device logic, shared by the kernels in roi_v2.cu and by the host emulation
// in tools/emu/roi_v2_emu.cpp (which compiles this very file with g++ and checks it against the CPU oracle before any
// GPU time is spent; every function below is warp-uniform in its control flow, so running it lane by lane on the host
// is equivalent to the warp's execution).
//
// Semantics: maskrcnn_benchmark/csrc/cuda/ROIAlign_cuda.cu:64-254 of the reference.  Design (not a port):
//   * the bilinear weight of a sample is separable and so is the "outside the map => 0" rule, hence
//        out[ph][pw] = (1/count) * sum_y Wy[ph][y] * sum_x Wx[pw][x] * V[y][x]            (exactly)
//     with per-axis tables Wy / Wx that depend on the RoI only.  plan kernel: one 16-word record per bin row, per bin
//     column and per footprint pixel column:  { first index | count << 16, up to 15 weights }.
//   * FORWARD, a warp owns (RoI, bin column pw, 32*V channels).  It works in STRIPS of up to kV2Rows map rows: phase 1
//     computes  T[y] = sum_x Wx[pw][x] * V[y][x]  for the strip's rows -- four rows at a time, all of their loads
//     (nx coalesced 512-byte requests per row) issued before the first is used, so a warp keeps up to 16 requests in
//     flight instead of one dependent gather after another -- into a lane-private strip of shared memory (every lane
//     reads back only what it wrote: no barrier); phase 2 emits every bin whose rows lie inside the strip,
//     out[ph] = sum_y Wy[ph][y] * T[y].  Sample rows grow monotonically with ph, so the next strip starts at the first
//     row of the first bin not yet emitted; each footprint pixel of the column is loaded once (plus a few rows where
//     strips overlap), for thin bins (several bins inside one map pixel, the P=14 case) as well as for fat ones.
//   * BACKWARD is the transposed gather.  A CTA owns (RoI, 32*V channels): phase 1 streams the RoI's pooled-gradient
//     tile [PH*PW][32*V] into shared memory with independent coalesced loads (FUSED: forms it on the fly, see below);
//     phase 2: a warp owns a footprint PIXEL column x; for every bin row it forms
//     G = sum_{pw covering x} Wx[pw][x] * g[ph][pw]  from the tile, adds Wy[ph][y] * G into a two-row cache of pixel
//     sums and, when a row leaves the cache, issues ONE vector reduction for that pixel: every footprint pixel of a RoI
//     is reduced exactly once per channel slice (the reference issues 4*g*g scalar atomicAdds per output element;
//     round 1 issued one reduction per pixel per covering column).
//   * NT = 2 forward pools the teacher and the student map with the same plan in one pass and emits the three
//     per-position channel sums the ARD loss needs (sum f_old^2, sum f_new^2, sum (f_new - f_old)^2); the FUSED
//     backward reads both pooled tensors, forms dL/df_new = ka*(f_new - f_old) + kb*f_new on the fly from per-position
//     coefficients and scatters it, so the ARD gradient tensor is never materialised.
#pragma once

#ifndef ABR_EMU
#include "common.cuh"
#define ABR_DEV __device__ __forceinline__
#define ABR_DEVM __device__ __forceinline__
#define ABR_HD __device__ __forceinline__
#define ABR_HOSTDEV __host__ __device__ __forceinline__
#define ABR_LDG4I(p) __ldg(reinterpret_cast<const int4*>(p))
#define ABR_LDGI(p) __ldg(p)
#define ABR_LDG2F(p) __ldg(p)
#endif

namespace abr {

struct LevelTable {
  void* ptr[ABR_MAX_LEVELS];
  int H[ABR_MAX_LEVELS];
  int W[ABR_MAX_LEVELS];
  float scale[ABR_MAX_LEVELS];
};

struct RoiGeom {
  int batch, level;
  float start_h, start_w, bin_h, bin_w;
  int grid_h, grid_w;
  float count;
};

// ROIAlign_cuda.cu:78-104.  No rounding of the scaled corners; RoI size floor is 1 feature pixel.
ABR_HD RoiGeom roi_geometry(const float* __restrict__ rois, const int32_t* __restrict__ levels, const LevelTable& lv, int r,
                            int PH, int PW, int ratio) {
  RoiGeom g;
  const float* roi = rois + 5 * (size_t)r;
  g.level = levels ? levels[r] : 0;
  const float scale = lv.scale[g.level];
  g.batch = (int)roi[0];
  g.start_w = __fmul_rn(roi[1], scale);
  g.start_h = __fmul_rn(roi[2], scale);
  float end_w = __fmul_rn(roi[3], scale);
  float end_h = __fmul_rn(roi[4], scale);
  float roi_w = fmaxf(__fsub_rn(end_w, g.start_w), 1.f);
  float roi_h = fmaxf(__fsub_rn(end_h, g.start_h), 1.f);
  g.bin_h = __fdiv_rn(roi_h, (float)PH);
  g.bin_w = __fdiv_rn(roi_w, (float)PW);
  g.grid_h = ratio > 0 ? ratio : (int)ceilf(__fdiv_rn(roi_h, (float)PH));
  g.grid_w = ratio > 0 ? ratio : (int)ceilf(__fdiv_rn(roi_w, (float)PW));
  g.count = (float)(g.grid_h * g.grid_w);
  return g;
}

// Sample coordinate of ROIAlign_cuda.cu:109,112 in the reference's operation order (no contraction).
ABR_HD float v2_sample_coord(float start, float bin, int p, int i, int grid) {
  return __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)), __fdiv_rn(__fmul_rn((float)i + .5f, bin), (float)grid));
}

// ------------------------------------------------------------------------------------------------ plan (v2)
constexpr int kV2Rec = 16;      // words per record: { lo | n << 16, w[0..14] }
constexpr int kV2Sup = 15;      // widest support (map pixels of one bin / bins over one pixel) a record holds
constexpr int kV2MaxFW = 64;    // widest footprint (map pixels) with pixel-column records
constexpr int kV2Hdr = 16;
enum V2Mode { V2_EMPTY = 0, V2_PLAN = 1, V2_GENERIC = 3 };
// hdr: [0] mode [1] batch [2] level [3] H [4] W [5] 1/count [6] X0 [7] FW [8] Y0 [9] Y1

ABR_HOSTDEV size_t v2_plan_words(int PH, int PW) { return (size_t)kV2Hdr + (size_t)(PH + PW + kV2MaxFW) * kV2Rec; }
ABR_HD const int* v2_col_rec(const int* plan, int pw) { return plan + kV2Hdr + pw * kV2Rec; }
ABR_HD const int* v2_bin_rec(const int* plan, int PW, int ph) { return plan + kV2Hdr + (PW + ph) * kV2Rec; }
ABR_HD const int* v2_pix_rec(const int* plan, int PH, int PW, int k) { return plan + kV2Hdr + (PW + PH + k) * kV2Rec; }

// One bin of one axis (ROIAlign_cuda.cu:22-47 along one axis): the map indices its samples touch and the summed
// bilinear weights.  Returns the support size n (0: no sample inside the map) or -1 when it exceeds kV2Sup.
ABR_HD int v2_axis_record(int* rec, int p, int S, float start, float bin, int grid) {
  float w[kV2Sup];
#pragma unroll
  for (int i = 0; i < kV2Sup; i++) w[i] = 0.f;
  int lo = -1, hi = -1;
  const float fS = (float)S;
  for (int i = 0; i < grid; i++) {
    float c = v2_sample_coord(start, bin, p, i, grid);
    if (c < -1.0f || c > fS) continue;
    if (c <= 0.f) c = 0.f;
    int low = (int)c, high;
    if (low >= S - 1) {
      high = low = S - 1;
      c = (float)low;
    } else {
      high = low + 1;
    }
    const float l = c - (float)low, h = 1.f - l;
    if (lo < 0) lo = low;
    if (low < lo || high - lo >= kV2Sup) return -1;
    w[low - lo] += h;
    w[high - lo] += l;
    if (high > hi) hi = high;
  }
  const int n = lo < 0 ? 0 : hi - lo + 1;
  rec[0] = (lo < 0 ? 0 : lo) | (n << 16);
  for (int i = 0; i < kV2Sup; i++) rec[1 + i] = __float_as_int(w[i]);
  return n;
}

// Phase 1 (threads tid, tid + nth, ...): the PW column records and the PH bin-row records.
ABR_HD void v2_plan_axes(int* plan, const RoiGeom& g, int H, int W, int PH, int PW, int tid, int nth) {
  for (int i = tid; i < PW + PH; i += nth) {
    int* rec = plan + kV2Hdr + i * kV2Rec;
    const int n = i < PW ? v2_axis_record(rec, i, W, g.start_w, g.bin_w, g.grid_w)
                         : v2_axis_record(rec, i - PW, H, g.start_h, g.bin_h, g.grid_h);
    if (n < 0) rec[0] = -1;  // support too wide: the RoI is left to the per-sample path
  }
}

// Phase 2 (one thread, after phase 1 is visible): header.
ABR_HD void v2_plan_header(int* plan, const RoiGeom& g, int H, int W, int PH, int PW) {
  int X0 = W, X1 = -1, Y0 = H, Y1 = -1;
  bool generic = false;
  for (int i = 0; i < PW + PH; i++) {
    const int w0 = plan[kV2Hdr + i * kV2Rec];
    if (w0 < 0) { generic = true; continue; }
    const int lo = w0 & 0xffff, n = w0 >> 16;
    if (n == 0) continue;
    if (i < PW) { X0 = lo < X0 ? lo : X0; X1 = lo + n - 1 > X1 ? lo + n - 1 : X1; }
    else { Y0 = lo < Y0 ? lo : Y0; Y1 = lo + n - 1 > Y1 ? lo + n - 1 : Y1; }
  }
  int mode = V2_PLAN;
  if (generic) mode = V2_GENERIC;
  else if (X1 < X0 || Y1 < Y0) mode = V2_EMPTY;
  else if (X1 - X0 + 1 > kV2MaxFW) mode = V2_GENERIC;
  plan[0] = mode; plan[1] = g.batch; plan[2] = g.level; plan[3] = H; plan[4] = W;
  plan[5] = __float_as_int(1.f / g.count);
  plan[6] = X1 < X0 ? 0 : X0; plan[7] = X1 < X0 ? 0 : X1 - X0 + 1;
  plan[8] = Y1 < Y0 ? 0 : Y0; plan[9] = Y1 < Y0 ? 0 : Y1;
  for (int i = 10; i < kV2Hdr; i++) plan[i] = 0;
}

// Phase 3 (threads tid, tid + nth, ..., after phase 2 is visible): one record per footprint pixel column x = X0 + k --
// the contiguous range of bin columns whose support contains x and their weights Wx[pw][x] (the transpose of the column
// records; supports start and end monotonically in pw, so the covering columns are contiguous).
ABR_HD void v2_plan_pix(int* plan, int PH, int PW, int tid, int nth) {
  if (plan[0] != V2_PLAN) return;
  const int X0 = plan[6], FW = plan[7];
  for (int k = tid; k < FW; k += nth) {
    const int x = X0 + k;
    int* rec = plan + kV2Hdr + (PW + PH + k) * kV2Rec;
    int q0 = -1, nq = 0;
    for (int i = 0; i < kV2Sup; i++) rec[1 + i] = 0;
    for (int q = 0; q < PW; q++) {
      const int* cr = plan + kV2Hdr + q * kV2Rec;
      const int lo = cr[0] & 0xffff, n = cr[0] >> 16;
      if (n == 0 || x < lo || x > lo + n - 1) continue;
      if (q0 < 0) q0 = q;
      if (q - q0 < kV2Sup) rec[1 + (q - q0)] = cr[1 + (x - lo)];
      nq = q - q0 + 1;
    }
    // a pixel under more than kV2Sup bin columns (a one-pixel RoI pooled to PW > 15) does not fit a record: the RoI goes
    // to the per-sample path (several threads may store the same value)
    if (nq > kV2Sup) plan[0] = V2_GENERIC;
    rec[0] = (q0 < 0 ? 0 : q0) | ((nq > kV2Sup ? kV2Sup : nq) << 16);
  }
}

// ------------------------------------------------------------------------------------------------ record access
struct V2Rec {
  int lo, n;
  float w0, w1, w2;
  const int* p;
};
ABR_DEV V2Rec v2_load_rec(const int* __restrict__ rec) {
  const int4 a = ABR_LDG4I(rec);
  V2Rec r;
  r.lo = a.x & 0xffff;
  r.n = a.x >> 16;
  r.w0 = __int_as_float(a.y);
  r.w1 = __int_as_float(a.z);
  r.w2 = __int_as_float(a.w);
  r.p = rec;
  return r;
}
ABR_DEV float v2_rec_w(const V2Rec& r, int i) {  // warp-uniform i
  return i == 0 ? r.w0 : i == 1 ? r.w1 : i == 2 ? r.w2 : __int_as_float(ABR_LDGI(r.p + 1 + i));
}

// Three warp totals with six shuffles (reduce-scatter): on return lane 0 holds sum(a), lane 16 sum(b), lane 8 sum(c).
ABR_DEV float v2_reduce3(float a, float b, float c, int lane) {
#ifndef ABR_EMU
  const bool hi = (lane & 16) != 0;
  float keep = (hi ? b : a) + __shfl_xor_sync(0xffffffffu, hi ? a : b, 16);
  c += __shfl_xor_sync(0xffffffffu, c, 16);
  const bool h8 = (lane & 8) != 0;
  float v = (h8 ? c : keep) + __shfl_xor_sync(0xffffffffu, h8 ? keep : c, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
#else
  (void)a; (void)b; (void)c; (void)lane;
  return 0.f;
#endif
}

// ------------------------------------------------------------------------------------------------ forward
// Emits one output bin of NT tensors: scale by 1/count, store, and (NT == 2) the ARD channel sums of this slice.
// sums_bin -> the three floats of (RoI, slice, bin).
template <typename T, int V, int NT>
ABR_DEV void v2_emit_bin(float (&acc)[NT][V], float inv_count, T* const (&o)[NT], bool active, float* sums_bin, int lane) {
#pragma unroll
  for (int t = 0; t < NT; t++) {
#pragma unroll
    for (int k = 0; k < V; k++) acc[t][k] *= inv_count;
    if (active) VecIO<T, V>::store(o[t], acc[t]);
  }
  if (NT == 2) {
    float so = 0.f, sn = 0.f, sd = 0.f;
    if (active) {
#pragma unroll
      for (int k = 0; k < V; k++) {
        const float a = acc[0][k], b = acc[NT - 1][k], d = b - a;
        so = fmaf(a, a, so);
        sn = fmaf(b, b, sn);
        sd = fmaf(d, d, sd);
      }
    }
#ifndef ABR_EMU
    const float v = v2_reduce3(so, sn, sd, lane);
    if ((lane & 7) == 0) {
      if (lane == 0) sums_bin[0] = v;
      else if (lane == 16) sums_bin[1] = v;
      else if (lane == 8) sums_bin[2] = v;
    }
#else
    (void)lane;
    sums_bin[0] += so; sums_bin[1] += sn; sums_bin[2] += sd;  // the harness zeroes the buffer and runs the lanes in turn
#endif
  }
}

// Lane-private / CTA-shared fp32 staging in shared memory: V consecutive floats of one lane, 16-byte accesses when V % 4 == 0.
template <int V>
ABR_DEV void v2_sm_store(float* p, const float (&v)[V]) {
#ifndef ABR_EMU
  if (V % 4 == 0) {
#pragma unroll
    for (int i = 0; i < V / 4; i++) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    return;
  }
#endif
#pragma unroll
  for (int i = 0; i < V; i++) p[i] = v[i];
}
template <int V>
ABR_DEV void v2_sm_load(const float* p, float (&v)[V]) {
#ifndef ABR_EMU
  if (V % 4 == 0) {
#pragma unroll
    for (int i = 0; i < V / 4; i++) {
      const float4 t = reinterpret_cast<const float4*>(p)[i];
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
    return;
  }
#endif
#pragma unroll
  for (int i = 0; i < V; i++) v[i] = p[i];
}

constexpr int kV2Rows = 16;  // map rows per strip (>= kV2Sup + 1, so the widest bin always fits one strip)

// Floats of shared memory one forward warp needs: [NT][kV2Rows][32 lanes][V].
ABR_HOSTDEV size_t v2_strip_floats(int V, int NT) { return (size_t)NT * kV2Rows * 32 * V; }

// One bin column of one RoI from its plan.  maps[t]: the level's map of tensor t ([B][H][W][C]); outs[t]: pooled tensor
// ([R][PH][PW][C]); c: first channel of this lane; sums_rs: &sums[(r * nslices + slice) * PH*PW * 3] (NT == 2);
// strip: this WARP's v2_strip_floats(V, NT) floats of shared memory.
template <typename T, int V, int NT>
ABR_DEV void v2_fwd_column(const int* __restrict__ plan, const T* const (&maps)[NT], T* const (&outs)[NT], float* sums_rs,
                           float* strip, int r, int pw, int c, bool active, int C, int PH, int PW, int lane) {
  const int4 h0 = ABR_LDG4I(plan), h1 = ABR_LDG4I(plan + 4), h2 = ABR_LDG4I(plan + 8);
  const int mode = h0.x, batch = h0.y, H = h0.w, W = h1.x, Y1 = h2.y;
  const float inv_count = __int_as_float(h1.y);
  const size_t pix = (size_t)C, binstride = (size_t)PW * C;
  T* o[NT];
#pragma unroll
  for (int t = 0; t < NT; t++) o[t] = outs[t] + ((size_t)r * PH * PW + pw) * C + c;
  float* sb = NT == 2 ? sums_rs + (size_t)pw * 3 : nullptr;
  const V2Rec col = v2_load_rec(v2_col_rec(plan, pw));
  const int nx = mode == V2_PLAN ? col.n : 0;
  const size_t rowstride = (size_t)W * C;
  const T* base[NT];
#pragma unroll
  for (int t = 0; t < NT; t++) base[t] = maps[t] + ((size_t)batch * H * W + col.lo) * C + c;
  float* mine = strip + (size_t)lane * V;  // [t][row] at (t * kV2Rows + row) * 32 * V
  constexpr int RB = V >= 8 ? 2 : 4;       // rows whose loads are in flight together (register budget)

  int ph = 0;
  while (ph < PH) {
    V2Rec bin = v2_load_rec(v2_bin_rec(plan, PW, ph));
    if (nx == 0 || bin.n == 0) {  // no sample of this bin (column, RoI) falls inside the map
      float z[NT][V];
#pragma unroll
      for (int t = 0; t < NT; t++)
#pragma unroll
        for (int k = 0; k < V; k++) z[t][k] = 0.f;
      v2_emit_bin<T, V, NT>(z, 0.f, o, active, sb, lane);
#pragma unroll
      for (int t = 0; t < NT; t++) o[t] += binstride;
      if (NT == 2) sb += (size_t)PW * 3;
      ph++;
      continue;
    }
    // ---- phase 1: T of the rows ystart .. ystart + nrows - 1
    const int ystart = bin.lo;
    const int nrows = Y1 - ystart + 1 < kV2Rows ? Y1 - ystart + 1 : kV2Rows;
#pragma unroll
    for (int t = 0; t < NT; t++) {
      const T* colbase = base[t] + (size_t)ystart * rowstride;
      for (int i0 = 0; i0 < nrows; i0 += RB) {
        float Tacc[RB][V];
#pragma unroll
        for (int j = 0; j < RB; j++)
#pragma unroll
          for (int k = 0; k < V; k++) Tacc[j][k] = 0.f;
        for (int kg = 0; kg < nx; kg += 4) {  // warp-uniform; one trip unless the column is wider than four map pixels
          float v[RB][4][V];
#pragma unroll
          for (int j = 0; j < RB; j++) {
            if (i0 + j < nrows) {
              const T* p = colbase + (size_t)(i0 + j) * rowstride + (size_t)kg * pix;
              VecIO<T, V>::load(p, v[j][0]);
              if (kg + 1 < nx) VecIO<T, V>::load(p + pix, v[j][1]);
              if (kg + 2 < nx) VecIO<T, V>::load(p + 2 * pix, v[j][2]);
              if (kg + 3 < nx) VecIO<T, V>::load(p + 3 * pix, v[j][3]);
            }
          }
          const float w0 = v2_rec_w(col, kg);
          const float w1 = kg + 1 < nx ? v2_rec_w(col, kg + 1) : 0.f;
          const float w2 = kg + 2 < nx ? v2_rec_w(col, kg + 2) : 0.f;
          const float w3 = kg + 3 < nx ? v2_rec_w(col, kg + 3) : 0.f;
#pragma unroll
          for (int j = 0; j < RB; j++) {
            if (i0 + j < nrows) {
#pragma unroll
              for (int k = 0; k < V; k++) {
                float s = fmaf(w0, v[j][0][k], Tacc[j][k]);
                if (kg + 1 < nx) s = fmaf(w1, v[j][1][k], s);
                if (kg + 2 < nx) s = fmaf(w2, v[j][2][k], s);
                if (kg + 3 < nx) s = fmaf(w3, v[j][3][k], s);
                Tacc[j][k] = s;
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < RB; j++)
          if (i0 + j < nrows) v2_sm_store<V>(mine + (size_t)(t * kV2Rows + i0 + j) * 32 * V, Tacc[j]);
      }
    }
    // ---- phase 2: every bin whose rows lie inside the strip (at least the one that started it: n <= kV2Sup < kV2Rows)
    while (true) {
      float acc[NT][V];
#pragma unroll
      for (int t = 0; t < NT; t++)
#pragma unroll
        for (int k = 0; k < V; k++) acc[t][k] = 0.f;
      const int off = bin.lo - ystart;
      for (int i = 0; i < bin.n; i++) {
        const float wy = v2_rec_w(bin, i);
#pragma unroll
        for (int t = 0; t < NT; t++) {
          float x[V];
          v2_sm_load<V>(mine + (size_t)(t * kV2Rows + off + i) * 32 * V, x);
#pragma unroll
          for (int k = 0; k < V; k++) acc[t][k] = fmaf(wy, x[k], acc[t][k]);
        }
      }
      v2_emit_bin<T, V, NT>(acc, inv_count, o, active, sb, lane);
#pragma unroll
      for (int t = 0; t < NT; t++) o[t] += binstride;
      if (NT == 2) sb += (size_t)PW * 3;
      if (++ph >= PH) break;
      bin = v2_load_rec(v2_bin_rec(plan, PW, ph));
      if (bin.n == 0 || bin.lo + bin.n > ystart + nrows) break;  // an empty bin or one that needs a new strip: outer loop
    }
  }
}

// Per-sample evaluation of one bin column, the reference's loop nest (ROIAlign_cuda.cu:64-122): for the rare RoIs the
// plan marks GENERIC (a bin wider than kV2Sup map pixels, a footprint wider than kV2MaxFW).
template <typename T, int V, int NT>
ABR_DEV void v2_generic_fwd_column(const RoiGeom& g, int H, int W, const T* const (&maps)[NT], T* const (&outs)[NT],
                                   float* sums_rs, int r, int pw, int c, bool active, int C, int PH, int PW, int lane) {
  const size_t binstride = (size_t)PW * C;
  T* o[NT];
#pragma unroll
  for (int t = 0; t < NT; t++) o[t] = outs[t] + ((size_t)r * PH * PW + pw) * C + c;
  float* sb = NT == 2 ? sums_rs + (size_t)pw * 3 : nullptr;
  const float fH = (float)H, fW = (float)W;
  const float inv_count = 1.f / g.count;
  for (int ph = 0; ph < PH; ph++) {
    float acc[NT][V];
#pragma unroll
    for (int t = 0; t < NT; t++)
#pragma unroll
      for (int k = 0; k < V; k++) acc[t][k] = 0.f;
    for (int iy = 0; iy < g.grid_h; iy++) {
      float y = v2_sample_coord(g.start_h, g.bin_h, ph, iy, g.grid_h);
      if (y < -1.0f || y > fH) continue;
      if (y <= 0.f) y = 0.f;
      int yl = (int)y, yh;
      if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
      const float ly = y - (float)yl, hy = 1.f - ly;
      for (int ix = 0; ix < g.grid_w; ix++) {
        float x = v2_sample_coord(g.start_w, g.bin_w, pw, ix, g.grid_w);
        if (x < -1.0f || x > fW) continue;
        if (x <= 0.f) x = 0.f;
        int xl = (int)x, xh;
        if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
        const float lx = x - (float)xl, hx = 1.f - lx;
        const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
#pragma unroll
        for (int t = 0; t < NT; t++) {
          const T* img = maps[t] + (size_t)g.batch * H * W * C + c;
          float v1[V], v2[V], v3[V], v4[V];
          VecIO<T, V>::load(img + ((size_t)yl * W + xl) * C, v1);
          VecIO<T, V>::load(img + ((size_t)yl * W + xh) * C, v2);
          VecIO<T, V>::load(img + ((size_t)yh * W + xl) * C, v3);
          VecIO<T, V>::load(img + ((size_t)yh * W + xh) * C, v4);
#pragma unroll
          for (int k = 0; k < V; k++) acc[t][k] += w1 * v1[k] + w2 * v2[k] + w3 * v3[k] + w4 * v4[k];
        }
      }
    }
    v2_emit_bin<T, V, NT>(acc, inv_count, o, active, sb, lane);
#pragma unroll
    for (int t = 0; t < NT; t++) o[t] += binstride;
    if (NT == 2) sb += (size_t)PW * 3;
  }
}

// ------------------------------------------------------------------------------------------------ backward
// Source of the pooled gradient g[bin] (V channels of this lane): the upstream gradient tensor, or (FUSED) the ARD
// gradient ka*(f_new - f_old) + kb*f_new formed from the two pooled tensors and the per-position coefficients.
template <typename T, int V, bool FUSED>
struct V2Grad {
  const T* a;          // gout, or f_old (FUSED), at [r][0][0][c]
  const T* b;          // f_new (FUSED)
  const float2* coef;  // [PH*PW] of this RoI (FUSED)
  ABR_DEVM void load(int bin, size_t C, float (&g)[V]) const {
    if (FUSED) {
      const float2 kc = ABR_LDG2F(coef + bin);
      float fo[V], fn[V];
      VecIO<T, V>::load(a + (size_t)bin * C, fo);
      VecIO<T, V>::load(b + (size_t)bin * C, fn);
#pragma unroll
      for (int k = 0; k < V; k++) g[k] = fmaf(kc.x, fn[k] - fo[k], kc.y * fn[k]);
    } else {
      VecIO<T, V>::load(a + (size_t)bin * C, g);
    }
  }
};

// Phase 1 of the backward: warp `warp` of `nw` brings the bins warp, warp + nw, ... of this (RoI, slice) gradient tile into
// shared memory -- tile[(bin * 32 + lane) * V .. + V) -- four bins' loads in flight at a time.  Lanes of a ragged last
// slice (inactive) store zeros.
template <typename T, int V, bool FUSED>
ABR_DEV void v2_bwd_fill_tile(float* tile, const V2Grad<T, V, FUSED>& src, int nbin, int C, int warp, int nw, int lane, bool active) {
  float* mine = tile + (size_t)lane * V;
  for (int b0 = warp; b0 < nbin; b0 += 4 * nw) {
    float g[4][V];
#pragma unroll
    for (int j = 0; j < 4; j++) {
#pragma unroll
      for (int k = 0; k < V; k++) g[j][k] = 0.f;
      if (active && b0 + j * nw < nbin) src.load(b0 + j * nw, (size_t)C, g[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (b0 + j * nw < nbin) v2_sm_store<V>(mine + (size_t)(b0 + j * nw) * 32 * V, g[j]);
  }
}

// Phase 2: one footprint pixel column x = X0 + k of one RoI.  gmap: the level's gradient map [B][H][W][C]; tile: the
// (RoI, slice) gradient tile in shared memory.
template <typename T, int V>
ABR_DEV void v2_bwd_pixcol(const int* __restrict__ plan, T* gmap, const float* tile, int k, int c, int C, int PH, int PW, int lane) {
  const int4 h0 = ABR_LDG4I(plan), h1 = ABR_LDG4I(plan + 4);
  const int batch = h0.y, H = h0.w, W = h1.x, X0 = h1.z;
  const float inv_count = __int_as_float(h1.y);
  const V2Rec px = v2_load_rec(v2_pix_rec(plan, PH, PW, k));
  const int nq = px.n;
  if (nq == 0) return;  // a map column between two bin columns' supports (sparse fixed-ratio sampling)
  const float wq3 = nq > 3 ? v2_rec_w(px, 3) : 0.f;
  const size_t rowstride = (size_t)W * C;
  T* gin = gmap + ((size_t)batch * H * W + (X0 + k)) * C + c;
  const float* mine = tile + (size_t)lane * V;
  float Sa[V], Sb[V];
  int ya = -1, yb = -1;
#pragma unroll
  for (int i = 0; i < V; i++) Sa[i] = Sb[i] = 0.f;
  for (int ph = 0; ph < PH; ph++) {
    const V2Rec bin = v2_load_rec(v2_bin_rec(plan, PW, ph));
    if (bin.n == 0) continue;
    // G = sum over the bin columns covering x of Wx[pw][x] * g[ph][pw]
    float G[V];
    {
      float g0[V], g1[V], g2[V], g3[V];
      const float* t0 = mine + (size_t)(ph * PW + px.lo) * 32 * V;
      v2_sm_load<V>(t0, g0);
      if (nq > 1) v2_sm_load<V>(t0 + 32 * V, g1);
      if (nq > 2) v2_sm_load<V>(t0 + 2 * 32 * V, g2);
      if (nq > 3) v2_sm_load<V>(t0 + 3 * 32 * V, g3);
#pragma unroll
      for (int i = 0; i < V; i++) {
        float s = px.w0 * g0[i];
        if (nq > 1) s = fmaf(px.w1, g1[i], s);
        if (nq > 2) s = fmaf(px.w2, g2[i], s);
        if (nq > 3) s = fmaf(wq3, g3[i], s);
        G[i] = s;
      }
      for (int j = 4; j < nq; j++) {
        const float wj = v2_rec_w(px, j);
        float gj[V];
        v2_sm_load<V>(t0 + (size_t)j * 32 * V, gj);
#pragma unroll
        for (int i = 0; i < V; i++) G[i] = fmaf(wj, gj[i], G[i]);
      }
    }
    for (int i = 0; i < bin.n; i++) {
      const float wy = v2_rec_w(bin, i) * inv_count;
      if (wy == 0.f) continue;
      const int y = bin.lo + i;
      if (y != yb && y != ya) {  // a new row: the older cached row is complete, reduce it into the map
        if (ya >= 0) VecIO<T, V>::red_add(gin + (size_t)ya * rowstride, Sa);
#pragma unroll
        for (int q = 0; q < V; q++) { Sa[q] = Sb[q]; Sb[q] = 0.f; }
        ya = yb;
        yb = y;
      }
      if (y == yb) {
#pragma unroll
        for (int q = 0; q < V; q++) Sb[q] = fmaf(wy, G[q], Sb[q]);
      } else {
#pragma unroll
        for (int q = 0; q < V; q++) Sa[q] = fmaf(wy, G[q], Sa[q]);
      }
    }
  }
  if (ya >= 0) VecIO<T, V>::red_add(gin + (size_t)ya * rowstride, Sa);
  if (yb >= 0) VecIO<T, V>::red_add(gin + (size_t)yb * rowstride, Sb);
}

// Per-sample backward of one bin column (ROIAlign_cuda.cu:177-254) for GENERIC RoIs, gradients from the tile.
template <typename T, int V>
ABR_DEV void v2_generic_bwd_column(const RoiGeom& g, int H, int W, T* gmap, const float* tile, int pw, int c, int C, int PH, int PW,
                                   int lane) {
  const float fH = (float)H, fW = (float)W;
  T* img = gmap + (size_t)g.batch * H * W * C + c;
  for (int ph = 0; ph < PH; ph++) {
    float top[V];
    v2_sm_load<V>(tile + ((size_t)(ph * PW + pw) * 32 + lane) * V, top);
    for (int iy = 0; iy < g.grid_h; iy++) {
      float y = v2_sample_coord(g.start_h, g.bin_h, ph, iy, g.grid_h);
      if (y < -1.0f || y > fH) continue;
      if (y <= 0.f) y = 0.f;
      int yl = (int)y, yh;
      if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
      const float ly = y - (float)yl, hy = 1.f - ly;
      for (int ix = 0; ix < g.grid_w; ix++) {
        float x = v2_sample_coord(g.start_w, g.bin_w, pw, ix, g.grid_w);
        if (x < -1.0f || x > fW) continue;
        if (x <= 0.f) x = 0.f;
        int xl = (int)x, xh;
        if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
        const float lx = x - (float)xl, hx = 1.f - lx;
        const float w[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
        const size_t off[4] = {((size_t)yl * W + xl) * C, ((size_t)yl * W + xh) * C, ((size_t)yh * W + xl) * C,
                               ((size_t)yh * W + xh) * C};
#pragma unroll
        for (int q = 0; q < 4; q++) {
          float v[V];
#pragma unroll
          for (int i = 0; i < V; i++) v[i] = top[i] * w[q] / g.count;
          VecIO<T, V>::red_add(img + off[q], v);
        }
      }
    }
  }
}

}  // namespace abr
