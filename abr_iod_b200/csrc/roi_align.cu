// roi_align.cu -- ROIAlign forward / backward for sm_100a, single- and multi-level in one launch.
//
// Semantics: maskrcnn_benchmark/csrc/cuda/ROIAlign_cuda.cu:64-254 of the reference (== csrc/cpu/ROIAlign_cpu.cpp).
// Design (not a port): the reference gives every output scalar its own thread, which re-derives the RoI
// geometry and issues 4*g*g scattered 4-byte gathers.  Here
//   1. the two SEPARABLE interpolation tables of every RoI are built once (plan_kernel, or in shared memory by the
//      self-contained kernels):
//        Wy[ph][y] = sum over the bin's sample rows of the bilinear row weight landing on map row y
//        Wx[pw][x] = same along x
//      (the reference's weight w1..w4 of a sample is hy*hx, hy*lx, ly*hx, ly*lx and its "outside the
//      map => 0" rule is the AND of a y-test and an x-test, so  out = Wy * V * Wx^T / count  exactly);
//   2. NHWC path: a warp owns (RoI, bin column, 32*V channels) with V consecutive channels per lane (16-byte
//      vectors: 4 x fp32 or 8 x bf16), so each map pixel is one coalesced 512 B warp load, every distinct pixel of
//      the column's footprint is read once per RoI instead of once per sample tap, and all weights are warp-uniform;
//      NCHW path (the reference's contiguous layout): threads walk the flat (c,ph,pw) output so stores are
//      fully coalesced, with the same tables in shared memory.
//   3. backward mirrors the forward sweep: one vector reduction per footprint pixel of a column
//      (red.global.add.v4.f32 / v4.bf16x2, a contiguous 512 B request per warp) instead of 4*g*g scalar atomicAdds
//      per output element.
// Sample coordinates are evaluated with explicitly rounded fp32 intrinsics in the reference's operation order
// so that the in/out-of-map decisions (ROIAlign_cuda.cu:22-25) are identical to the reference's.
#include <cuda.h>

#include <cstdlib>

#include <type_traits>

#include "common.cuh"
#include "roi_v2.cuh"
#include "roi_v2.h"

namespace abr {

// Builds one axis table:  Wt[p*stride + i] (zero elsewhere) and the closed support range [lo[p], hi[p]]
// (lo > hi when bin p has no sample inside the map).  Every thread of the CTA must call this.
// Coordinates follow ROIAlign_cuda.cu:109,112 and the case analysis of :22-47 along one axis.
__device__ __forceinline__ void build_axis_table(float* Wt, int* lo, int* hi, int P, int S, int stride, float start,
                                                 float bin, int grid) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < P * stride; i += nt) Wt[i] = 0.f;
  for (int i = tid; i < P; i += nt) { lo[i] = S; hi[i] = -1; }
  __syncthreads();
  const long long total = (long long)P * grid;
  const float fgrid = (float)grid, fS = (float)S;
  for (long long s = tid; s < total; s += nt) {
    const int p = (int)(s / grid), i = (int)(s - (long long)p * grid);
    float c = __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)), __fdiv_rn(__fmul_rn((float)i + .5f, bin), fgrid));
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
    atomicAdd(&Wt[p * stride + low], h);
    atomicAdd(&Wt[p * stride + high], l);
    atomicMin(&lo[p], low);
    atomicMax(&hi[p], high);
  }
  __syncthreads();
}

struct Tables {
  float *Wy, *Wx;
  int *ylo, *yhi, *xlo, *xhi;
  int *aux;  // backward only: per-row / per-column bin ranges
};

__host__ __device__ inline size_t tables_floats(int PH, int PW, int Hs, int Ws) {
  return (size_t)PH * Hs + (size_t)PW * Ws;
}
// shared memory: [Wy PH*Hs][Wx PW*Ws][ylo PH][yhi PH][xlo PW][xhi PW][aux 2*Hs + 2*Ws (+4)]
static size_t tables_bytes(int PH, int PW, int Hs, int Ws, bool backward) {
  size_t b = tables_floats(PH, PW, Hs, Ws) * 4 + (size_t)(2 * PH + 2 * PW) * 4;
  if (backward) b += (size_t)(2 * Hs + 2 * Ws + 8) * 4;
  return b;
}
__device__ __forceinline__ Tables carve(float* smem, int PH, int PW, int Hs, int Ws) {
  Tables t;
  t.Wy = smem;
  t.Wx = t.Wy + (size_t)PH * Hs;
  t.ylo = reinterpret_cast<int*>(t.Wx + (size_t)PW * Ws);
  t.yhi = t.ylo + PH;
  t.xlo = t.yhi + PH;
  t.xhi = t.xlo + PW;
  t.aux = t.xhi + PW;
  return t;
}

// ------------------------------------------------------------------------------------------ backward helpers
// After the two axis tables exist: overall footprint [Y0,Y1]x[X0,X1] and, for every map row / column inside it,
// the (contiguous) range of bins whose support contains it.  aux = [plo Hs][phi Hs][qlo Ws][qhi Ws][Y0,Y1,X0,X1,span]
// where span = max over rows of (number of bins containing the row) - 1 (xspan: the same over columns).
__device__ __forceinline__ void build_inverse_ranges(const Tables& t, int PH, int PW, int H, int W, int Hs, int Ws) {
  int* plo = t.aux;
  int* phi = plo + Hs;
  int* qlo = phi + Hs;
  int* qhi = qlo + Ws;
  int* fp = qhi + Ws;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int y = tid; y < H; y += nt) {
    int lo = PH, hi = -1;
    for (int p = 0; p < PH; p++)
      if (t.ylo[p] <= y && y <= t.yhi[p]) { lo = min(lo, p); hi = p; }
    plo[y] = lo; phi[y] = hi;
  }
  for (int x = tid; x < W; x += nt) {
    int lo = PW, hi = -1;
    for (int p = 0; p < PW; p++)
      if (t.xlo[p] <= x && x <= t.xhi[p]) { lo = min(lo, p); hi = p; }
    qlo[x] = lo; qhi[x] = hi;
  }
  __syncthreads();
  if (tid == 0) {
    int a = H, b = -1, c = W, d = -1, span = 0;
    for (int p = 0; p < PH; p++)
      if (t.ylo[p] <= t.yhi[p]) { a = min(a, t.ylo[p]); b = max(b, t.yhi[p]); }
    for (int p = 0; p < PW; p++)
      if (t.xlo[p] <= t.xhi[p]) { c = min(c, t.xlo[p]); d = max(d, t.xhi[p]); }
    int xspan = 0;
    for (int y = a; y <= b; y++) span = max(span, phi[y] - plo[y]);
    for (int x = c; x <= d; x++) xspan = max(xspan, qhi[x] - qlo[x]);
    fp[0] = a; fp[1] = b; fp[2] = c; fp[3] = d; fp[4] = span; fp[5] = xspan;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------ RoI plans
// A "plan" is the compact form of one RoI's two interpolation tables, written once per call by plan_kernel (a small
// CTA per RoI) into caller-provided workspace and then read, warp-uniformly and straight out of L1/L2, by the sweep
// kernels, which therefore need no shared memory, no barriers and no per-CTA setup:
//   hdr  [16 words]         mode, batch index, level, Y0, Y1, 1/count, H, W, X0, X1, backward scheme (1 column pairs,
//                           2 pixel records, 0 neither), (5 spare)
//   col  [PW][4 + kPlanNx]  x0, nx, -, -, then the nx column weights Wx[pw][x0 ..] zero-padded to kPlanNx
//   row  [Y1-Y0+1][8 | 16]  ROLLING: first bin a holding the row (-1: none), Wy[a][y], Wy[a+1][y] (0 if not shared)
//                           THIN:    Wy[0..PH-1][y]   (8 words per row when PH <= 8, else 16)
// ROLLING = every map row feeds at most two vertically adjacent bins (bins at least ~1 map pixel tall);
// THIN    = thinner bins and PH <= 16;  GENERIC = anything else (columns wider than kPlanNx pixels, thin bins with
// PH > 16): those RoIs are left to the table-in-shared-memory kernels below;  EMPTY = no sample inside the map.
constexpr int kPlanNx = 16;
constexpr int kTileMaxPx = 64;  // widest footprint row the TMA-staged kernel holds in one ring slot
constexpr int kPlanHdr = 16;
constexpr int kPlanCol = 4 + kPlanNx;
constexpr int kPlanRow = 8;      // words of a row record when PH <= 8 (the TMA-staged kernels' format)
constexpr int kThinBins = 8;     // bins a THIN row record holds in that format
constexpr int kThinBinsMax = 16; // PH <= 16 (the framework's default 14x14 pooling): 16-word row records, sweep kernels only
__host__ __device__ inline int plan_row_words(int PH) { return PH <= kThinBins ? kPlanRow : kThinBinsMax; }
// Backward only, narrow RoIs (bins thinner than a map pixel in x, so a pixel lies in three or more bin columns and the
// column-pair scheme below does not apply): one record per footprint pixel -- first covering column, number of covering
// columns (<= kPixCols), their weights -- so that the TMA-staged backward reduces such a pixel once instead of once per
// covering column.  Footprints of up to kPixMax pixels; the area sits after the row records.
constexpr int kPixMax = 16, kPixCols = 7, kPixRec = 8, kPlanPix = kPixMax * kPixRec;  // record: first column | count << 8, 7 weights
enum PlanMode { PLAN_EMPTY = 0, PLAN_ROLLING = 1, PLAN_THIN = 2, PLAN_GENERIC = 3 };

__host__ __device__ inline size_t plan_stride_words(int PW, int Hs, int PH) {
  return (size_t)kPlanHdr + (size_t)PW * kPlanCol + (size_t)Hs * plan_row_words(PH) + kPlanPix;
}

__global__ void __launch_bounds__(128) plan_kernel(LevelTable lv, const float* __restrict__ rois,
                                                  const int32_t* __restrict__ levels, int* __restrict__ plans,
                                                  size_t stride, int PH, int PW, int ratio, int Hs, int Ws) {
  extern __shared__ float smem[];
  __shared__ int s_mode;
  const int r = blockIdx.x;
  const RoiGeom g = roi_geometry(rois, levels, lv, r, PH, PW, ratio);
  const int H = lv.H[g.level], W = lv.W[g.level];
  Tables t = carve(smem, PH, PW, Hs, Ws);
  build_axis_table(t.Wy, t.ylo, t.yhi, PH, H, Hs, g.start_h, g.bin_h, g.grid_h);
  build_axis_table(t.Wx, t.xlo, t.xhi, PW, W, Ws, g.start_w, g.bin_w, g.grid_w);
  build_inverse_ranges(t, PH, PW, H, W, Hs, Ws);
  const int* plo = t.aux;
  const int* phi = plo + Hs;
  const int* fp = phi + Hs + 2 * Ws;
  const int Y0 = fp[0], Y1 = fp[1];
  int* plan = plans + (size_t)r * stride;
  if (threadIdx.x == 0) {
    int mode;
    if (Y0 > Y1 || fp[2] > fp[3]) {
      mode = PLAN_EMPTY;
    } else {
      bool wide = fp[3] - fp[2] + 1 > kTileMaxPx;  // footprint wider than the staged row tile
      for (int p = 0; p < PW; p++) wide |= (t.xhi[p] - t.xlo[p] + 1 > kPlanNx);
      mode = wide ? PLAN_GENERIC : (fp[4] <= 1 ? PLAN_ROLLING : (PH <= kThinBinsMax ? PLAN_THIN : PLAN_GENERIC));
    }
    s_mode = mode;
    plan[0] = mode; plan[1] = g.batch; plan[2] = g.level; plan[3] = Y0; plan[4] = Y1;
    plan[5] = __float_as_int(1.f / g.count); plan[6] = H; plan[7] = W;
    plan[8] = fp[2]; plan[9] = fp[3];
    int scheme = 0;
    if (mode == PLAN_ROLLING && fp[5] <= 1) scheme = 1;
    else if (mode == PLAN_ROLLING && fp[5] < kPixCols && fp[3] - fp[2] + 1 <= kPixMax && PH <= kThinBins) scheme = 2;
    plan[10] = scheme;
    for (int i = 11; i < kPlanHdr; i++) plan[i] = 0;
  }
  __syncthreads();
  const int mode = s_mode;
  if (mode == PLAN_EMPTY || mode == PLAN_GENERIC) return;
  int* col = plan + kPlanHdr;
  for (int i = threadIdx.x; i < PW * kPlanCol; i += blockDim.x) {
    const int p = i / kPlanCol, k = i - p * kPlanCol;
    const int x0 = t.xlo[p], x1 = t.xhi[p];
    const bool has = x0 <= x1;
    // backward only: a map pixel shared by two adjacent columns is reduced ONCE, by the later column, which adds the
    // earlier column's term; [2] = trailing pixels this column leaves to the next one, [3] = leading pixels for which it
    // also carries the previous column.  Valid when no pixel lies in three columns and rows are ROLLING.
    const bool pair = mode == PLAN_ROLLING && fp[5] <= 1;
    int v = 0;
    if (k == 0) v = has ? x0 : 0;
    else if (k == 1) v = has ? x1 - x0 + 1 : 0;
    else if (k == 2) v = (pair && has && p + 1 < PW && t.xlo[p + 1] <= t.xhi[p + 1]) ? max(0, x1 - t.xlo[p + 1] + 1) : 0;
    else if (k == 3) v = (pair && has && p > 0 && t.xlo[p - 1] <= t.xhi[p - 1]) ? max(0, t.xhi[p - 1] - x0 + 1) : 0;
    else if (k >= 4) v = (has && x0 + (k - 4) <= x1) ? __float_as_int(t.Wx[(size_t)p * Ws + x0 + (k - 4)]) : 0;
    col[i] = v;
  }
  int* row = col + PW * kPlanCol;
  const int nrows = Y1 - Y0 + 1;
  const int rw = plan_row_words(PH);
  for (int i = threadIdx.x; i < nrows * rw; i += blockDim.x) {
    const int y = Y0 + i / rw, k = i % rw;
    int v = 0;
    if (mode == PLAN_ROLLING) {
      const int p0 = plo[y], p1 = phi[y];
      if (k == 0) v = p0 <= p1 ? p0 : -1;
      else if (k == 1) v = p0 <= p1 ? __float_as_int(t.Wy[(size_t)p0 * Hs + y]) : 0;
      else if (k == 2) v = p1 > p0 ? __float_as_int(t.Wy[(size_t)(p0 + 1) * Hs + y]) : 0;
    } else {
      v = k < PH ? __float_as_int(t.Wy[(size_t)k * Hs + y]) : 0;
    }
    row[i] = v;
  }
  if (mode == PLAN_ROLLING && fp[5] > 1 && fp[5] < kPixCols && fp[3] - fp[2] + 1 <= kPixMax && PH <= kThinBins) {
    const int* qlo = phi + Hs;
    const int* qhi = qlo + Ws;
    int* pixrec = plan + kPlanHdr + PW * kPlanCol + (size_t)Hs * rw;
    const int X0 = fp[2], npx = fp[3] - fp[2] + 1;
    for (int i = threadIdx.x; i < npx * kPixRec; i += blockDim.x) {
      const int x = X0 + i / kPixRec, k = i % kPixRec;
      const int lo = qlo[x], cnt = qhi[x] - lo + 1;
      int v = 0;
      if (k == 0) v = cnt > 0 ? (lo | (cnt << 8)) : 0;
      else if (k - 1 < cnt) v = __float_as_int(t.Wx[(size_t)(lo + k - 1) * Ws + x]);
      pixrec[i] = v;
    }
  }
}

struct SweepTask {
  int mode, batch, level, Y0, nrows, H, W, x0, nx, pw, c;
  float inv_count;
  bool active;
  const int* col;
  const int4* rows;
  int rw4;  // int4 units per row record
};

// Decodes the warp's task = (RoI, bin column pw, slice of 32*V channels).  Returns false when there is nothing to do.
template <int V>
__device__ __forceinline__ bool sweep_task(SweepTask& k, const int* __restrict__ plans, size_t stride, int C, int PH, int PW,
                                           int nslices, long long ntasks, int& r) {
  const long long task = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (task >= ntasks) return false;
  const int per_roi = PW * nslices;
  r = (int)(task / per_roi);
  const int tt = (int)(task - (long long)r * per_roi);
  const int slice = tt / PW;
  k.pw = tt - slice * PW;
  const int* plan = plans + (size_t)r * stride;
  const int4 h0 = __ldg(reinterpret_cast<const int4*>(plan));
  const int4 h1 = __ldg(reinterpret_cast<const int4*>(plan) + 1);
  k.mode = h0.x; k.batch = h0.y; k.level = h0.z; k.Y0 = h0.w;
  k.nrows = h1.x - h0.w + 1; k.inv_count = __int_as_float(h1.y); k.H = h1.z; k.W = h1.w;
  if (k.mode == PLAN_GENERIC) return false;
  k.c = (slice * 32 + (threadIdx.x & 31)) * V;
  k.active = k.c < C;
  if (!k.active) k.c = 0;  // idle lanes of a ragged last slice shadow channel 0 and never store
  k.col = plan + kPlanHdr + k.pw * kPlanCol;
  k.rows = reinterpret_cast<const int4*>(plan + kPlanHdr + PW * kPlanCol);
  k.rw4 = plan_row_words(PH) / 4;
  k.x0 = k.col[0];
  k.nx = k.mode == PLAN_EMPTY ? 0 : k.col[1];
  return true;
}

// t = sum over the column's pixels of Wx * v for one map row; q0..q3 point at the row's first four column pixels
// (clamped onto the last one when the column is narrower; their weights are then 0).
template <typename T, int V>
__device__ __forceinline__ void row_dot(float (&tr)[V], const T* q0, const T* q1, const T* q2, const T* q3,
                                        const float4 w, int nx, size_t pix, const int* __restrict__ col) {
  float v0[V], v1[V], v2[V], v3[V];
  VecIO<T, V>::load(q0, v0);
  VecIO<T, V>::load(q1, v1);
  VecIO<T, V>::load(q2, v2);
  VecIO<T, V>::load(q3, v3);
#pragma unroll
  for (int k = 0; k < V; k++) tr[k] = fmaf(w.w, v3[k], fmaf(w.z, v2[k], fmaf(w.y, v1[k], w.x * v0[k])));
  if (nx > 4) {  // warp-uniform, rare: columns of 5..kPlanNx pixels
    const T* q = q0 + 4 * pix;
    for (int j = 4; j < nx; j++, q += pix) {
      float v[V];
      VecIO<T, V>::load(q, v);
      const float b = __int_as_float(__ldg(col + 4 + j));
#pragma unroll
      for (int k = 0; k < V; k++) tr[k] = fmaf(b, v[k], tr[k]);
    }
  }
}

// ------------------------------------------------------------------------------------------ forward, NHWC, plan-driven
// A WARP owns (RoI, output column pw, slice of 32*V channels): its lanes read the same map pixel at consecutive
// channels (one coalesced 512 B request), all weights are warp-uniform, and it sweeps the column's footprint rows
// ONCE: for every row y it forms  t = sum_x Wx[pw][x] * v[y][x]  and adds  Wy[ph][y] * t  to the (at most two)
// vertically adjacent bins the row belongs to, held in two rolling register accumulators.  Every distinct pixel of
// the column's footprint is loaded once per RoI, not once per bin or per sample tap.
template <typename T, int V, int NB>
__global__ void __launch_bounds__(256) roi_align_fwd_sweep_kernel(LevelTable lv, const int* __restrict__ plans,
                                                                 size_t stride, T* __restrict__ out, int C, int PH,
                                                                 int PW, int nslices, long long ntasks) {
  SweepTask k;
  int r;
  if (!sweep_task<V>(k, plans, stride, C, PH, PW, nslices, ntasks, r)) return;
  const size_t pix = (size_t)C, binstride = (size_t)PW * C;
  T* o = out + ((size_t)r * PH * PW + k.pw) * C + k.c;
  if (k.nx == 0) {  // no sample of this column (or of the whole RoI) falls inside the map
    float z[V];
#pragma unroll
    for (int i = 0; i < V; i++) z[i] = 0.f;
    if (k.active)
      for (int ph = 0; ph < PH; ph++) VecIO<T, V>::store_stream(o + (size_t)ph * binstride, z);
    return;
  }
  const size_t rowstride = (size_t)k.W * C;
  const T* q0 = static_cast<const T*>(lv.ptr[k.level]) + ((size_t)k.batch * k.H + k.Y0) * rowstride + (size_t)k.x0 * pix + k.c;
  const T* q1 = q0 + (size_t)min(1, k.nx - 1) * pix;
  const T* q2 = q0 + (size_t)min(2, k.nx - 1) * pix;
  const T* q3 = q0 + (size_t)min(3, k.nx - 1) * pix;
  const float4 w = __ldg(reinterpret_cast<const float4*>(k.col + 4));
  const int4* rr = k.rows;
  const float inv_count = k.inv_count;
  if (k.mode == PLAN_ROLLING) {
    float accA[V], accB[V];
#pragma unroll
    for (int i = 0; i < V; i++) accA[i] = accB[i] = 0.f;
    int a = 0;
    int4 info_next = __ldg(rr);
    for (int n = k.nrows; n > 0; n--, q0 += rowstride, q1 += rowstride, q2 += rowstride, q3 += rowstride) {
      const int4 info = info_next;
      rr += k.rw4;
      if (n > 1) info_next = __ldg(rr);  // row records are fetched one row ahead
      if (info.x < 0) continue;
      float tr[V];
      row_dot<T, V>(tr, q0, q1, q2, q3, w, k.nx, pix, k.col);
      if (a != info.x) {  // bins a .. info.x-1 are complete: emit them and roll the two-bin window
        do {
#pragma unroll
          for (int i = 0; i < V; i++) accA[i] *= inv_count;
          if (k.active) VecIO<T, V>::store_stream(o + (size_t)a * binstride, accA);
#pragma unroll
          for (int i = 0; i < V; i++) { accA[i] = accB[i]; accB[i] = 0.f; }
        } while (++a < info.x);
      }
      const float wa = __int_as_float(info.y), wb = __int_as_float(info.z);
#pragma unroll
      for (int i = 0; i < V; i++) {
        accA[i] = fmaf(wa, tr[i], accA[i]);
        accB[i] = fmaf(wb, tr[i], accB[i]);
      }
    }
    for (; a < PH; a++) {
#pragma unroll
      for (int i = 0; i < V; i++) accA[i] *= inv_count;
      if (k.active) VecIO<T, V>::store_stream(o + (size_t)a * binstride, accA);
#pragma unroll
      for (int i = 0; i < V; i++) { accA[i] = accB[i]; accB[i] = 0.f; }
    }
  } else {  // PLAN_THIN: one statically indexed accumulator per bin; rows carry all PH weights (NB = 8 or 16 of them)
    float acc[NB][V];
#pragma unroll
    for (int p = 0; p < NB; p++)
#pragma unroll
      for (int i = 0; i < V; i++) acc[p][i] = 0.f;
    for (int n = k.nrows; n > 0; n--, rr += NB / 4, q0 += rowstride, q1 += rowstride, q2 += rowstride, q3 += rowstride) {
      int4 wv[NB / 4];
#pragma unroll
      for (int i = 0; i < NB / 4; i++) wv[i] = __ldg(rr + i);
      float tr[V];
      row_dot<T, V>(tr, q0, q1, q2, q3, w, k.nx, pix, k.col);
#pragma unroll
      for (int p = 0; p < NB; p++) {
        const int4 q4 = wv[p / 4];
        const float wy = __int_as_float((p & 3) == 0 ? q4.x : (p & 3) == 1 ? q4.y : (p & 3) == 2 ? q4.z : q4.w);
#pragma unroll
        for (int i = 0; i < V; i++) acc[p][i] = fmaf(wy, tr[i], acc[p][i]);
      }
    }
#pragma unroll
    for (int p = 0; p < NB; p++) {
      if (p < PH && k.active) {
#pragma unroll
        for (int i = 0; i < V; i++) acc[p][i] *= inv_count;
        VecIO<T, V>::store_stream(o + (size_t)p * binstride, acc[p]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ backward, NHWC, plan-driven
// Same ownership.  For every footprint row the warp folds the bins of its column that hold the row into
// s = sum_p Wy[p][y] * g[p][pw] / count  (g kept in registers) and issues ONE vector reduction
// gin[y][x] += Wx[pw][x] * s  per footprint pixel of the column: a warp-wide contiguous 512 B
// red.global.add.v4.f32 (v4.bf16x2 for bf16) instead of the reference's 4*g*g scalar atomicAdds per output element.
// gin[y][x0 + j] += w_j * s for the column's pixels of one row (zero weights skipped: the address may be a clamp).
template <typename T, int V>
__device__ __forceinline__ void row_scatter(T* q0, const float (&s)[V], const float4 w, int nx, size_t pix,
                                            const int* __restrict__ col) {
  float v[V];
  if (w.x != 0.f) {
#pragma unroll
    for (int i = 0; i < V; i++) v[i] = w.x * s[i];
    VecIO<T, V>::red_add(q0, v);
  }
  if (nx > 1 && w.y != 0.f) {
#pragma unroll
    for (int i = 0; i < V; i++) v[i] = w.y * s[i];
    VecIO<T, V>::red_add(q0 + pix, v);
  }
  if (nx > 2 && w.z != 0.f) {
#pragma unroll
    for (int i = 0; i < V; i++) v[i] = w.z * s[i];
    VecIO<T, V>::red_add(q0 + 2 * pix, v);
  }
  if (nx > 3 && w.w != 0.f) {
#pragma unroll
    for (int i = 0; i < V; i++) v[i] = w.w * s[i];
    VecIO<T, V>::red_add(q0 + 3 * pix, v);
  }
  if (nx > 4) {
    T* q = q0 + 4 * pix;
    for (int j = 4; j < nx; j++, q += pix) {
      const float b = __int_as_float(__ldg(col + 4 + j));
      if (b != 0.f) {
#pragma unroll
        for (int i = 0; i < V; i++) v[i] = b * s[i];
        VecIO<T, V>::red_add(q, v);
      }
    }
  }
}

template <typename T, int V, int NB>
__global__ void __launch_bounds__(256) roi_align_bwd_sweep_kernel(LevelTable lv, const int* __restrict__ plans,
                                                                 size_t stride, const T* __restrict__ gout, int C, int PH,
                                                                 int PW, int nslices, long long ntasks) {
  SweepTask k;
  int r;
  if (!sweep_task<V>(k, plans, stride, C, PH, PW, nslices, ntasks, r)) return;
  if (k.nx == 0) return;  // warp-uniform
  const bool active = k.active;  // idle lanes of a ragged last slice shadow channel 0 and never reduce
  const size_t pix = (size_t)C, binstride = (size_t)PW * C;
  const T* __restrict__ go = gout + ((size_t)r * PH * PW + k.pw) * C + k.c;
  const size_t rowstride = (size_t)k.W * C;
  T* q0 = static_cast<T*>(lv.ptr[k.level]) + ((size_t)k.batch * k.H + k.Y0) * rowstride + (size_t)k.x0 * pix + k.c;
  const float4 w = __ldg(reinterpret_cast<const float4*>(k.col + 4));
  const int4* rr = k.rows;
  const float inv_count = k.inv_count;
  if (k.mode == PLAN_ROLLING) {
    // Pixels shared with the next column are left to it; for the leading pixels shared with the previous column this
    // warp adds that column's term too, so every footprint pixel of the RoI receives exactly one reduction.
    const int n_emit = k.nx - k.col[2], n_prev = k.col[3];
    const int* pcol = k.col - kPlanCol;  // previous column's record (only read when n_prev > 0)
    const int poff = n_prev > 0 ? k.x0 - pcol[0] : 0;
    float4 wp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n_prev > 0) {
      wp.x = __int_as_float(pcol[4 + poff]);
      if (n_prev > 1) wp.y = __int_as_float(pcol[4 + poff + 1]);
      if (n_prev > 2) wp.z = __int_as_float(pcol[4 + poff + 2]);
      if (n_prev > 3) wp.w = __int_as_float(pcol[4 + poff + 3]);
    }
    const T* __restrict__ gp = go - C;  // the previous column's bins
    // the column's PH gradient vectors come from DRAM exactly once: start all of them now so that the rolling window
    // below finds them in L1/L2 instead of paying one exposed DRAM latency per bin
    for (int p = 0; p < PH; p++) {
      asm volatile("prefetch.global.L1 [%0];" ::"l"(go + (size_t)p * binstride));
      if (n_prev > 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(gp + (size_t)p * binstride));
    }
    float gA[V], gB[V], pA[V], pB[V];  // gradients of bins a and a+1 of this column and of the previous one
#pragma unroll
    for (int i = 0; i < V; i++) pA[i] = pB[i] = 0.f;
    int a = -2;  // no bin cached yet
    const int lane = threadIdx.x & 31;
    for (int base = 0; base < k.nrows; base += 32) {  // row records: one coalesced fetch per 32 rows, then shuffles
      const int cnt = min(32, k.nrows - base);
      const int4 mine = lane < cnt ? __ldg(rr + k.rw4 * (base + lane)) : make_int4(-1, 0, 0, 0);
      for (int j = 0; j < cnt; j++, q0 += rowstride) {
        int4 info;
        info.x = __shfl_sync(0xffffffffu, mine.x, j);
        info.y = __shfl_sync(0xffffffffu, mine.y, j);
        info.z = __shfl_sync(0xffffffffu, mine.z, j);
        if (info.x < 0) continue;
        if (a != info.x) {
          if (a + 1 == info.x) {
#pragma unroll
            for (int i = 0; i < V; i++) { gA[i] = gB[i]; pA[i] = pB[i]; }
          } else {
            VecIO<T, V>::load(go + (size_t)info.x * binstride, gA);
            if (n_prev > 0) VecIO<T, V>::load(gp + (size_t)info.x * binstride, pA);
          }
          a = info.x;
          if (a + 1 < PH) {
            VecIO<T, V>::load(go + (size_t)(a + 1) * binstride, gB);
            if (n_prev > 0) VecIO<T, V>::load(gp + (size_t)(a + 1) * binstride, pB);
          } else {
#pragma unroll
            for (int i = 0; i < V; i++) gB[i] = pB[i] = 0.f;
          }
        }
        const float wa = __int_as_float(info.y) * inv_count, wb = __int_as_float(info.z) * inv_count;
        float s[V], sp[V];
#pragma unroll
        for (int i = 0; i < V; i++) {
          s[i] = fmaf(wb, gB[i], wa * gA[i]);
          sp[i] = fmaf(wb, pB[i], wa * pA[i]);
        }
        float v[V];
        if (n_emit > 0 && (w.x != 0.f || wp.x != 0.f)) {
#pragma unroll
          for (int i = 0; i < V; i++) v[i] = fmaf(wp.x, sp[i], w.x * s[i]);
          if (active) VecIO<T, V>::red_add(q0, v);
        }
        if (n_emit > 1 && (w.y != 0.f || wp.y != 0.f)) {
#pragma unroll
          for (int i = 0; i < V; i++) v[i] = fmaf(wp.y, sp[i], w.y * s[i]);
          if (active) VecIO<T, V>::red_add(q0 + pix, v);
        }
        if (n_emit > 2 && (w.z != 0.f || wp.z != 0.f)) {
#pragma unroll
          for (int i = 0; i < V; i++) v[i] = fmaf(wp.z, sp[i], w.z * s[i]);
          if (active) VecIO<T, V>::red_add(q0 + 2 * pix, v);
        }
        if (n_emit > 3 && (w.w != 0.f || wp.w != 0.f)) {
#pragma unroll
          for (int i = 0; i < V; i++) v[i] = fmaf(wp.w, sp[i], w.w * s[i]);
          if (active) VecIO<T, V>::red_add(q0 + 3 * pix, v);
        }
        if (n_emit > 4) {  // warp-uniform, rare: columns of 5..kPlanNx pixels
          T* q = q0 + 4 * pix;
          for (int jj = 4; jj < n_emit; jj++, q += pix) {
            const float bw = __int_as_float(__ldg(k.col + 4 + jj));
            const float bp = jj < n_prev ? __int_as_float(__ldg(pcol + 4 + poff + jj)) : 0.f;
            if (bw != 0.f || bp != 0.f) {
#pragma unroll
              for (int i = 0; i < V; i++) v[i] = fmaf(bp, sp[i], bw * s[i]);
              if (active) VecIO<T, V>::red_add(q, v);
            }
          }
        }
      }
    }
  } else {  // PLAN_THIN
    float g[NB][V];
#pragma unroll
    for (int p = 0; p < NB; p++) {
      if (p < PH) {
        VecIO<T, V>::load(go + (size_t)p * binstride, g[p]);
      } else {
#pragma unroll
        for (int i = 0; i < V; i++) g[p][i] = 0.f;
      }
    }
    for (int n = k.nrows; n > 0; n--, rr += NB / 4, q0 += rowstride) {
      int4 wv[NB / 4];
#pragma unroll
      for (int i = 0; i < NB / 4; i++) wv[i] = __ldg(rr + i);
      float s[V];
#pragma unroll
      for (int i = 0; i < V; i++) s[i] = 0.f;
#pragma unroll
      for (int p = 0; p < NB; p++) {
        const int4 q4 = wv[p / 4];
        const float wy = __int_as_float((p & 3) == 0 ? q4.x : (p & 3) == 1 ? q4.y : (p & 3) == 2 ? q4.z : q4.w);
#pragma unroll
        for (int i = 0; i < V; i++) s[i] = fmaf(wy, g[p][i], s[i]);
      }
#pragma unroll
      for (int i = 0; i < V; i++) s[i] *= inv_count;
      if (active) row_scatter<T, V>(q0, s, w, k.nx, pix, k.col);
    }
  }
}

// ------------------------------------------------------------------------------------------ forward, NHWC, TMA-staged
// Warp-specialised, persistent.  A CTA owns (RoI, slice of 32*V channels = 512 bytes per pixel) at a time:
//   * warp 0 (producer) streams the RoI's footprint into a ring of shared-memory slots, ONE TMA tensor op per slot
//     (cp.async.bulk.tensor.3d: a box of bw pixels x bh rows x 512 bytes of the map, bw the narrowest power of two that
//     covers the footprint width, bw * bh = 64; completion counted on the slot's mbarrier), and the RoI's plan (column
//     records + per-row bin weights, cp.async.bulk) into a double-buffered plan area;
//   * warps 1..PW (consumers, one per bin column) wait for a slot, take their column's pixels of each staged row out of
//     shared memory, form t = sum_x Wx*v and add Wy[p][y]*t into two rolling register accumulators (or one per bin for
//     thin bins), then hand the slot back.
// Every distinct footprint pixel leaves L2 once per (RoI, slice) -- the columns share the staged rows -- the loads are
// asynchronous and no thread ever waits on a dependent global load.  PH <= 8, PW <= 7.
constexpr int kRing = 2;  // 32 KB slots: one being filled while the other is consumed
constexpr int kMaxBins = 8;

__device__ __forceinline__ unsigned s_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(b)) : "memory");
}
__device__ __forceinline__ void mb_wait(unsigned long long* b, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s_u32(b)), "r"(parity) : "memory");
}
// Producer-side wait: the producer warp is never on the critical path, so it backs off between polls instead of
// taking issue slots from the consumer warps of its SM sub-partition.
__device__ __forceinline__ void mb_wait_backoff(unsigned long long* b, unsigned parity) {
  const unsigned a = s_u32(b);
  unsigned done;
  for (;;) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(100);
  }
}
__device__ __forceinline__ void mb_wait_a(unsigned addr, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void mb_arrive_a(unsigned addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ unsigned lds32(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128(unsigned addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tma_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)),
               "l"(src), "r"(bytes), "r"(s_u32(b))
               : "memory");
}

// One TMA tensor op: a box of bh rows x bw pixels x 512 bytes of the [B*H rows][W pixels][C channels] view of a map
// (bw * bh = kTileMaxPx, so every box is one 32 KB ring slot).
__device__ __forceinline__ void tma_box_g2s(void* dst, const CUtensorMap* map, int c0, int x, int row, unsigned long long* b) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s_u32(dst)),
               "l"(map), "r"(c0), "r"(x), "r"(row), "r"(s_u32(b))
               : "memory");
}
constexpr int kTmaLevels = 4;  // feature levels the tensor-map path carries (kernel parameter space)
constexpr int kTmaBoxes = 6;   // boxes of 2x32, 4x16, 8x8, 16x4, 32x2, 64x1 (pixels x rows)
struct TmaMaps {
  CUtensorMap m[kTmaLevels][kTmaBoxes];
};

template <typename T, int V>
__device__ __forceinline__ void unpack16(const uint4 raw, float (&v)[V]);
template <>
__device__ __forceinline__ void unpack16<float, 4>(const uint4 raw, float (&v)[4]) {
  v[0] = __uint_as_float(raw.x); v[1] = __uint_as_float(raw.y); v[2] = __uint_as_float(raw.z); v[3] = __uint_as_float(raw.w);
}
template <>
__device__ __forceinline__ void unpack16<__nv_bfloat16, 8>(const uint4 raw, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256, 3) roi_align_fwd_tma_kernel(const __grid_constant__ TmaMaps maps,
                                                               const int* __restrict__ plans, size_t stride,
                                                               T* __restrict__ out, int C, int PH, int PW, int R,
                                                               int nslices, int Hs) {
  static_assert(V * sizeof(T) == 16, "one lane moves 16 bytes");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // [ring kRing][kTileMaxPx][32 lanes] uint4 | [plan 2][hdr+cols | wrow Hs*8] ints | barriers
  uint4* ring = reinterpret_cast<uint4*>(smem_raw);
  const int headw = kPlanHdr + PW * kPlanCol;   // words of header + column records
  const int planw = headw + Hs * 8;             // ... + per-row weights
  int* planbuf = reinterpret_cast<int*>(ring + (size_t)kRing * kTileMaxPx * 32);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(planbuf + 2 * (size_t)planw);
  unsigned long long* full = bars;              // [kRing] producer -> consumers (tx bytes)
  unsigned long long* empty = bars + kRing;     // [kRing] consumers -> producer (PW arrivals)
  unsigned long long* pfull = bars + 2 * kRing; // [2]
  unsigned long long* pempty = pfull + 2;       // [2]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; i++) { mb_init(&full[i], 1); mb_init(&empty[i], PW * 32); }
    for (int i = 0; i < 2; i++) { mb_init(&pfull[i], 1); mb_init(&pempty[i], PW * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int dr = (int)(gridDim.x / (unsigned)nslices), dslice = (int)(gridDim.x % (unsigned)nslices);
  unsigned it = 0;  // rows streamed so far by this CTA (ring position), identical in every warp
  int ti = 0;       // tasks done so far (plan buffer position)

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    // task = r * nslices + slice walks blockIdx.x, blockIdx.x + gridDim.x, ...: (r, slice) advance without a division
    for (int r = (int)(blockIdx.x / (unsigned)nslices), slice = (int)(blockIdx.x % (unsigned)nslices); r < R; ti++, slice += dslice, r += dr) {
      if (slice >= nslices) { slice -= nslices; r++; if (r >= R) break; }
      const int* plan = plans + (size_t)r * stride;
      const int4 h0 = __ldg(reinterpret_cast<const int4*>(plan));
      const int4 h1 = __ldg(reinterpret_cast<const int4*>(plan) + 1);
      const int4 h2 = __ldg(reinterpret_cast<const int4*>(plan) + 2);
      const int mode = h0.x, batch = h0.y, level = h0.z, Y0 = h0.w, nrows = h1.x - h0.w + 1, H = h1.z, W = h1.w;
      const int X0 = h2.x, wf = h2.y - h2.x + 1;
      const int pb = ti & 1;
      mb_wait_backoff(&pempty[pb], ((ti >> 1) & 1) ^ 1);
      int* dst = planbuf + (size_t)pb * planw;
      const bool live = mode == PLAN_ROLLING || mode == PLAN_THIN;
      if (lane == 0) {
        const unsigned rowbytes = live ? (unsigned)nrows * 32u : 0u;
        mb_expect_tx(&pfull[pb], (unsigned)headw * 4u + rowbytes);
        tma_g2s(dst, plan, (unsigned)headw * 4u, &pfull[pb]);
        if (rowbytes) tma_g2s(dst + headw, plan + headw, rowbytes, &pfull[pb]);
      }
      if (!live) continue;
      // one tensor op per ring slot: the narrowest box (2..32 pixels wide) that covers the footprint width, as many
      // rows tall as fit in the slot; channels beyond C and pixels beyond the map's width are zero-filled by the TMA unit
      int bi = 0;
      while ((2 << bi) < wf) bi++;
      const int bh = kTileMaxPx >> (bi + 1);
      const CUtensorMap* map = &maps.m[level][bi];
      const int c0 = slice * 32 * V;
      int grow = batch * H + Y0;
      for (int row = 0; row < nrows; row += bh, it++, grow += bh) {
        const int slot = it % kRing;
        mb_wait_backoff(&empty[slot], ((it / kRing) & 1) ^ 1);
        if (lane == 0) {
          mb_expect_tx(&full[slot], (unsigned)kTileMaxPx * 512u);
          tma_box_g2s(ring + (size_t)slot * kTileMaxPx * 32, map, c0, X0, grow, &full[slot]);
        }
      }
    }
  } else if (warp <= PW) {
    // ------------------------------------------------------------------ consumers: warp w owns bin column w-1
    const int pw = warp - 1;
    const unsigned ring_a = s_u32(ring) + lane * 16, full_a = s_u32(full), empty_a = s_u32(empty);
    // task = r * nslices + slice walks blockIdx.x, blockIdx.x + gridDim.x, ...: (r, slice) advance without a division
    for (int r = (int)(blockIdx.x / (unsigned)nslices), slice = (int)(blockIdx.x % (unsigned)nslices); r < R; ti++, slice += dslice, r += dr) {
      if (slice >= nslices) { slice -= nslices; r++; if (r >= R) break; }
      const int pb = ti & 1;
      mb_wait(&pfull[pb], (ti >> 1) & 1);
      const int* pl = planbuf + (size_t)pb * planw;
      const int mode = pl[0], nrows = pl[4] - pl[3] + 1, X0 = pl[8];
      const float inv_count = __int_as_float(pl[5]);
      const int c = (slice * 32 + lane) * V;
      const bool active = c < C;
      T* o = out + ((size_t)r * PH * PW + pw) * C + c;
      const size_t binstride = (size_t)PW * C;
      if (mode == PLAN_GENERIC) {  // served by the self-contained kernel
        mb_arrive(&pempty[pb]);
        continue;
      }
      const int* col = pl + kPlanHdr + pw * kPlanCol;
      const int nx = mode == PLAN_EMPTY ? 0 : col[1];
      if (nx == 0) {  // no sample of this column (or of the whole RoI) falls inside the map: zeros, but keep the ring moving
        if (mode != PLAN_EMPTY) {
          const int wf = pl[9] - X0 + 1;
          int bi = 0;
          while ((2 << bi) < wf) bi++;
          const int bh = kTileMaxPx >> (bi + 1);
          for (int row0 = 0; row0 < nrows; row0 += bh, it++) {
            const int slot = it % kRing;
            mb_wait_a(full_a + slot * 8, (it / kRing) & 1);
            mb_arrive_a(empty_a + slot * 8);
          }
        }
        float z[V];
#pragma unroll
        for (int i = 0; i < V; i++) z[i] = 0.f;
        if (active)
          for (int p = 0; p < PH; p++) VecIO<T, V>::store_stream(o + (size_t)p * binstride, z);
        mb_arrive(&pempty[pb]);
        continue;
      }
      const float w0 = __int_as_float(col[4]), w1 = __int_as_float(col[5]), w2 = __int_as_float(col[6]), w3 = __int_as_float(col[7]);
      const int off = col[0] - X0;
      const int wf = pl[9] - X0 + 1;
      int bi = 0;
      while ((2 << bi) < wf) bi++;
      const int bw = 2 << bi, bh = kTileMaxPx >> (bi + 1);
      // byte offsets of the column's first four pixels inside a staged row (clamped onto the last one, weight 0)
      const unsigned p0 = (unsigned)off * 512u, p1 = p0 + (unsigned)min(1, nx - 1) * 512u, p2 = p0 + (unsigned)min(2, nx - 1) * 512u,
                     p3 = p0 + (unsigned)min(3, nx - 1) * 512u;
      const unsigned rowbytes = (unsigned)bw * 512u;
      const unsigned rec_a = s_u32(pl + headw);
      // t = sum_x Wx * v of the staged row at shared address `ra`
      auto row_dot = [&](unsigned ra, float (&tr)[V]) {
        float v0[V], v1[V], v2[V], v3[V];
        unpack16<T, V>(lds128(ra + p0), v0);
        unpack16<T, V>(lds128(ra + p1), v1);
        unpack16<T, V>(lds128(ra + p2), v2);
        unpack16<T, V>(lds128(ra + p3), v3);
#pragma unroll
        for (int i = 0; i < V; i++) tr[i] = fmaf(w3, v3[i], fmaf(w2, v2[i], fmaf(w1, v1[i], w0 * v0[i])));
        for (int j = 4; j < nx; j++) {  // warp-uniform, rare: columns of 5..kPlanNx pixels
          float x[V];
          unpack16<T, V>(lds128(ra + p0 + j * 512u), x);
          const float bw_ = __int_as_float(col[4 + j]);
#pragma unroll
          for (int i = 0; i < V; i++) tr[i] = fmaf(bw_, x[i], tr[i]);
        }
      };
      if (mode == PLAN_ROLLING) {
        float accA[V], accB[V];
#pragma unroll
        for (int i = 0; i < V; i++) accA[i] = accB[i] = 0.f;
        int a = 0;
        T* oa = o;  // output address of bin a of this column
        for (int row0 = 0; row0 < nrows; row0 += bh, it++) {
          const int slot = it % kRing;
          mb_wait_a(full_a + slot * 8, (it / kRing) & 1);
          const int rows_here = min(bh, nrows - row0);
          unsigned ra = ring_a + (unsigned)slot * (kTileMaxPx * 512u);
          unsigned rec = rec_a + (unsigned)row0 * 32u;
          for (int rr = 0; rr < rows_here; rr++, ra += rowbytes, rec += 32u) {
            const uint4 info = lds128(rec);
            const int ia = (int)info.x;
            if (ia < 0) continue;  // a map row between two bins' supports (sparse fixed-ratio sampling)
            float tr[V];
            row_dot(ra, tr);
            if (a != ia) {  // bins a .. ia-1 are complete: emit them and roll the two-bin window
              do {
#pragma unroll
                for (int i = 0; i < V; i++) accA[i] *= inv_count;
                if (active) VecIO<T, V>::store_stream(oa, accA);
                oa += binstride;
#pragma unroll
                for (int i = 0; i < V; i++) { accA[i] = accB[i]; accB[i] = 0.f; }
              } while (++a < ia);
            }
            const float wa = __uint_as_float(info.y), wb = __uint_as_float(info.z);
#pragma unroll
            for (int i = 0; i < V; i++) {
              accA[i] = fmaf(wa, tr[i], accA[i]);
              accB[i] = fmaf(wb, tr[i], accB[i]);
            }
          }
          mb_arrive_a(empty_a + slot * 8);
        }
        for (; a < PH; a++, oa += binstride) {
#pragma unroll
          for (int i = 0; i < V; i++) accA[i] *= inv_count;
          if (active) VecIO<T, V>::store_stream(oa, accA);
#pragma unroll
          for (int i = 0; i < V; i++) { accA[i] = accB[i]; accB[i] = 0.f; }
        }
      } else {  // PLAN_THIN: one accumulator per bin; a row record holds all PH weights
        float acc[kMaxBins][V];
#pragma unroll
        for (int p = 0; p < kMaxBins; p++)
#pragma unroll
          for (int i = 0; i < V; i++) acc[p][i] = 0.f;
        for (int row0 = 0; row0 < nrows; row0 += bh, it++) {
          const int slot = it % kRing;
          mb_wait_a(full_a + slot * 8, (it / kRing) & 1);
          const int rows_here = min(bh, nrows - row0);
          unsigned ra = ring_a + (unsigned)slot * (kTileMaxPx * 512u);
          for (int rr = 0; rr < rows_here; rr++, ra += rowbytes) {
            float tr[V];
            row_dot(ra, tr);
            const uint4 wl = lds128(rec_a + (unsigned)(row0 + rr) * 32u), wh = lds128(rec_a + (unsigned)(row0 + rr) * 32u + 16u);
            const float wy[kMaxBins] = {__uint_as_float(wl.x), __uint_as_float(wl.y), __uint_as_float(wl.z), __uint_as_float(wl.w),
                                        __uint_as_float(wh.x), __uint_as_float(wh.y), __uint_as_float(wh.z), __uint_as_float(wh.w)};
#pragma unroll
            for (int p = 0; p < kMaxBins; p++)
#pragma unroll
              for (int i = 0; i < V; i++) acc[p][i] = fmaf(wy[p], tr[i], acc[p][i]);
          }
          mb_arrive_a(empty_a + slot * 8);
        }
#pragma unroll
        for (int p = 0; p < kMaxBins; p++) {
          if (p < PH && active) {
#pragma unroll
            for (int i = 0; i < V; i++) acc[p][i] *= inv_count;
            VecIO<T, V>::store_stream(o + (size_t)p * binstride, acc[p]);
          }
        }
      }
      mb_arrive(&pempty[pb]);
    }
  }
}

// ------------------------------------------------------------------------------------------ backward, NHWC, TMA-staged
// Same warp specialisation as the forward.  Per (RoI, slice of 32*V channels) the producer warp brings, with ONE TMA
// tensor op, the RoI's whole [PH*PW bins][32*V channels] gradient tile (25 KB for P=7) plus the plan into a
// double-buffered stage; the consumer warps (one per bin column) then sweep the footprint rows reading bin gradients
// out of shared memory -- no dependent global load anywhere -- and issue one vector reduction per distinct footprint
// pixel (pixels shared by two columns are reduced once, by the later column, see plan_kernel).
template <typename T, int V>
__global__ void __launch_bounds__(256, 4) roi_align_bwd_tma_kernel(const __grid_constant__ CUtensorMap gmap, LevelTable lv,
                                                               const int* __restrict__ plans, size_t stride, int C, int PH,
                                                               int PW, int R, int nslices, int Hs) {
  static_assert(V * sizeof(T) == 16, "one lane moves 16 bytes");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nbin = PH * PW;
  const unsigned gbytes = (unsigned)nbin * 512u;             // one stage of bin gradients
  const unsigned gstage = (gbytes + 127u) & ~127u;
  const int headw = kPlanHdr + PW * kPlanCol;
  const int rowsw = Hs * kPlanRow;
  const int planw = headw + rowsw + kPlanPix;  // header + columns | row records | pixel records
  uint4* gtile = reinterpret_cast<uint4*>(smem_raw);         // [2][nbin][32 lanes]
  int* planbuf = reinterpret_cast<int*>(smem_raw + 2 * (size_t)gstage);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(planbuf + 2 * (size_t)planw);
  unsigned long long* pfull = bars;        // [2] producer -> consumers (tx bytes)
  unsigned long long* pempty = bars + 2;   // [2] consumers -> producer
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; i++) { mb_init(&pfull[i], 1); mb_init(&pempty[i], PW * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int dr = (int)(gridDim.x / (unsigned)nslices), dslice = (int)(gridDim.x % (unsigned)nslices);
  int ti = 0;
  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    // task = r * nslices + slice walks blockIdx.x, blockIdx.x + gridDim.x, ...: (r, slice) advance without a division
    for (int r = (int)(blockIdx.x / (unsigned)nslices), slice = (int)(blockIdx.x % (unsigned)nslices); r < R; ti++, slice += dslice, r += dr) {
      if (slice >= nslices) { slice -= nslices; r++; if (r >= R) break; }
      const int* plan = plans + (size_t)r * stride;
      const int4 h0 = __ldg(reinterpret_cast<const int4*>(plan));
      const int4 h1 = __ldg(reinterpret_cast<const int4*>(plan) + 1);
      const int4 h2 = __ldg(reinterpret_cast<const int4*>(plan) + 2);
      const int mode = h0.x, nrows = h1.x - h0.w + 1;
      const bool pixels = h2.z == 2;  // narrow RoI: per-pixel records follow the row records
      const int pb = ti & 1;
      mb_wait_backoff(&pempty[pb], ((ti >> 1) & 1) ^ 1);
      const bool live = mode == PLAN_ROLLING || mode == PLAN_THIN;
      if (lane == 0) {
        const unsigned rowbytes = live ? (unsigned)nrows * (kPlanRow * 4u) : 0u;
        const unsigned pixbytes = (live && pixels) ? (unsigned)(h2.y - h2.x + 1) * (kPixRec * 4u) : 0u;
        mb_expect_tx(&pfull[pb], (unsigned)headw * 4u + rowbytes + pixbytes + (live ? gbytes : 0u));
        int* dst = planbuf + (size_t)pb * planw;
        tma_g2s(dst, plan, (unsigned)headw * 4u, &pfull[pb]);
        if (pixbytes) tma_g2s(dst + headw + rowsw, plan + headw + rowsw, pixbytes, &pfull[pb]);
        if (live) {
          tma_g2s(dst + headw, plan + headw, rowbytes, &pfull[pb]);
          tma_box_g2s(smem_raw + (size_t)pb * gstage, &gmap, slice * 32 * V, 0, r, &pfull[pb]);
        }
      }
    }
  } else if (warp <= PW) {
    // ------------------------------------------------------------------ consumers: warp w owns bin column w-1
    const int pw = warp - 1;
    // task = r * nslices + slice walks blockIdx.x, blockIdx.x + gridDim.x, ...: (r, slice) advance without a division
    for (int r = (int)(blockIdx.x / (unsigned)nslices), slice = (int)(blockIdx.x % (unsigned)nslices); r < R; ti++, slice += dslice, r += dr) {
      if (slice >= nslices) { slice -= nslices; r++; if (r >= R) break; }
      (void)r;
      const int pb = ti & 1;
      mb_wait(&pfull[pb], (ti >> 1) & 1);
      const int* pl = planbuf + (size_t)pb * planw;
      const int mode = pl[0];
      const int* col = pl + kPlanHdr + pw * kPlanCol;
      const int c = (slice * 32 + lane) * V;
      if (mode == PLAN_ROLLING && pl[10] == 2) {
        // Narrow RoI, pixel by pixel: this warp owns the footprint pixels X0 + pw, X0 + pw + PW, ... and reduces each of
        // them ONCE per row, adding up the (<= kPixCols) bin columns that cover it from the pixel's record.
        if (c < C) {
          const int X0 = pl[8], X1 = pl[9], Y0p = pl[3], nrows_p = pl[4] - pl[3] + 1;
          const float inv_cnt = __int_as_float(pl[5]);
          const size_t rowstride_p = (size_t)pl[7] * C;
          const unsigned g0 = s_u32(smem_raw + (size_t)pb * gstage) + lane * 16;  // bin (0, 0)
          const unsigned binrow_p = (unsigned)PW * 512u;
          const unsigned rec_p = s_u32(pl + headw);
          const unsigned pix_p = s_u32(pl + headw + rowsw);
          T* qrow = static_cast<T*>(lv.ptr[pl[2]]) + ((size_t)pl[1] * pl[6] + Y0p) * rowstride_p + c;
          for (int row = 0; row < nrows_p; row++, qrow += rowstride_p) {
            const uint4 info = lds128(rec_p + (unsigned)row * 32u);
            const int a = (int)info.x;
            if (a < 0) continue;
            const float wa = __uint_as_float(info.y) * inv_cnt, wb = (a + 1 < PH) ? __uint_as_float(info.z) * inv_cnt : 0.f;
            const unsigned ga = g0 + (unsigned)a * binrow_p, gb = ga + ((a + 1 < PH) ? binrow_p : 0u);
            for (int x = X0 + pw; x <= X1; x += PW) {
              const unsigned ra = pix_p + (unsigned)(x - X0) * (kPixRec * 4u);
              const uint4 r0 = lds128(ra), r1 = lds128(ra + 16u);
              const int clo = (int)(r0.x & 255u), cnt = (int)(r0.x >> 8);
              const float wx[kPixCols] = {__uint_as_float(r0.y), __uint_as_float(r0.z), __uint_as_float(r0.w), __uint_as_float(r1.x),
                                          __uint_as_float(r1.y), __uint_as_float(r1.z), __uint_as_float(r1.w)};
              float v[V];
#pragma unroll
              for (int i = 0; i < V; i++) v[i] = 0.f;
#pragma unroll
              for (int j = 0; j < kPixCols; j++) {
                if (j < cnt && wx[j] != 0.f) {
                  const unsigned off = (unsigned)(clo + j) * 512u;
                  float gA[V], gB[V];
                  unpack16<T, V>(lds128(ga + off), gA);
                  unpack16<T, V>(lds128(gb + off), gB);
                  const float ka = wx[j] * wa, kb = wx[j] * wb;
#pragma unroll
                  for (int i = 0; i < V; i++) v[i] = fmaf(kb, gB[i], fmaf(ka, gA[i], v[i]));
                }
              }
              if (cnt > 0) VecIO<T, V>::red_add(qrow + (size_t)x * C, v);
            }
          }
        }
        mb_arrive(&pempty[pb]);
        continue;
      }
      const int nx = (mode == PLAN_ROLLING || mode == PLAN_THIN) ? col[1] : 0;
      if (nx == 0 || c >= C) {  // nothing to reduce (EMPTY / GENERIC RoI, column outside the map, idle lane of a ragged slice)
        mb_arrive(&pempty[pb]);
        continue;
      }
      const int batch = pl[1], level = pl[2], Y0 = pl[3], nrows = pl[4] - pl[3] + 1, H = pl[6], W = pl[7];
      const float inv_count = __int_as_float(pl[5]);
      const size_t pix = (size_t)C, rowstride = (size_t)W * C;
      T* q0 = static_cast<T*>(lv.ptr[level]) + (((size_t)batch * H + Y0) * W + col[0]) * C + c;
      const float w0 = __int_as_float(col[4]), w1 = __int_as_float(col[5]), w2 = __int_as_float(col[6]), w3 = __int_as_float(col[7]);
      // pixels shared with the next column are left to it; for the leading pixels shared with the previous column this
      // warp adds that column's term too: every footprint pixel of the RoI receives exactly one reduction
      const int n_emit = nx - col[2], n_prev = col[3];
      const int* pcol = col - kPlanCol;
      const int poff = n_prev > 0 ? col[0] - pcol[0] : 0;
      const float wp0 = n_prev > 0 ? __int_as_float(pcol[4 + poff]) : 0.f, wp1 = n_prev > 1 ? __int_as_float(pcol[5 + poff]) : 0.f,
                  wp2 = n_prev > 2 ? __int_as_float(pcol[6 + poff]) : 0.f, wp3 = n_prev > 3 ? __int_as_float(pcol[7 + poff]) : 0.f;
      const unsigned g_a = s_u32(smem_raw + (size_t)pb * gstage) + lane * 16 + (unsigned)pw * 512u;  // bin (0, pw)
      const unsigned binrow = (unsigned)PW * 512u;                                                    // next bin row
      const unsigned rec_a = s_u32(pl + headw);
      for (int row = 0; row < nrows; row++, q0 += rowstride) {
        float s[V], sp[V];
        if (mode == PLAN_ROLLING) {
          const uint4 info = lds128(rec_a + (unsigned)row * 32u);
          const int a = (int)info.x;
          if (a < 0) continue;
          const float wa = __uint_as_float(info.y) * inv_count, wb = __uint_as_float(info.z) * inv_count;
          float gA[V], gB[V];
          unpack16<T, V>(lds128(g_a + (unsigned)a * binrow), gA);
          if (a + 1 < PH) {
            unpack16<T, V>(lds128(g_a + (unsigned)(a + 1) * binrow), gB);
          } else {
#pragma unroll
            for (int i = 0; i < V; i++) gB[i] = 0.f;
          }
#pragma unroll
          for (int i = 0; i < V; i++) s[i] = fmaf(wb, gB[i], wa * gA[i]);
          if (n_prev > 0) {
            unpack16<T, V>(lds128(g_a - 512u + (unsigned)a * binrow), gA);
            if (a + 1 < PH) unpack16<T, V>(lds128(g_a - 512u + (unsigned)(a + 1) * binrow), gB);
#pragma unroll
            for (int i = 0; i < V; i++) sp[i] = fmaf(wb, gB[i], wa * gA[i]);
          } else {
#pragma unroll
            for (int i = 0; i < V; i++) sp[i] = 0.f;
          }
        } else {  // PLAN_THIN: the row record holds all PH weights (the pair scheme is off: n_prev == 0)
#pragma unroll
          for (int i = 0; i < V; i++) s[i] = sp[i] = 0.f;
          for (int p = 0; p < PH; p++) {
            const float wy = __int_as_float(pl[headw + row * kPlanRow + p]) * inv_count;
            if (wy == 0.f) continue;
            float g[V];
            unpack16<T, V>(lds128(g_a + (unsigned)p * binrow), g);
#pragma unroll
            for (int i = 0; i < V; i++) s[i] = fmaf(wy, g[i], s[i]);
          }
        }
        float v[V];
        if (n_emit > 0 && (w0 != 0.f || wp0 != 0.f)) {
#pragma unroll
          for (int i = 0; i < V; i++) v[i] = fmaf(wp0, sp[i], w0 * s[i]);
          VecIO<T, V>::red_add(q0, v);
        }
        if (n_emit > 1 && (w1 != 0.f || wp1 != 0.f)) {
#pragma unroll
          for (int i = 0; i < V; i++) v[i] = fmaf(wp1, sp[i], w1 * s[i]);
          VecIO<T, V>::red_add(q0 + pix, v);
        }
        if (n_emit > 2 && (w2 != 0.f || wp2 != 0.f)) {
#pragma unroll
          for (int i = 0; i < V; i++) v[i] = fmaf(wp2, sp[i], w2 * s[i]);
          VecIO<T, V>::red_add(q0 + 2 * pix, v);
        }
        if (n_emit > 3 && (w3 != 0.f || wp3 != 0.f)) {
#pragma unroll
          for (int i = 0; i < V; i++) v[i] = fmaf(wp3, sp[i], w3 * s[i]);
          VecIO<T, V>::red_add(q0 + 3 * pix, v);
        }
        if (n_emit > 4) {  // warp-uniform, rare: columns of 5..kPlanNx pixels
          T* q = q0 + 4 * pix;
          for (int j = 4; j < n_emit; j++, q += pix) {
            const float bw = __int_as_float(col[4 + j]);
            const float bp = j < n_prev ? __int_as_float(pcol[4 + poff + j]) : 0.f;
            if (bw != 0.f || bp != 0.f) {
#pragma unroll
              for (int i = 0; i < V; i++) v[i] = fmaf(bp, sp[i], bw * s[i]);
              VecIO<T, V>::red_add(q, v);
            }
          }
        }
      }
      mb_arrive(&pempty[pb]);
    }
  }
}

// ------------------------------------------------------------------------------------------ forward, NHWC, self-contained
// Table-in-shared-memory kernel: a CTA owns one RoI, every thread V consecutive channels, plain per-bin loops over the
// separable tables.  Used when the caller passes no workspace, and for the RoIs a plan marks GENERIC.
template <typename T, int V>
__global__ void __launch_bounds__(256) roi_align_fwd_nhwc_kernel(LevelTable lv, const float* __restrict__ rois,
                                                                const int32_t* __restrict__ levels,
                                                                T* __restrict__ out, int C, int PH, int PW, int ratio,
                                                                int Hs, int Ws, const int* __restrict__ plans,
                                                                size_t plan_stride) {
  extern __shared__ float smem[];
  const int r = blockIdx.x;
  if (plans && plans[(size_t)r * plan_stride] != PLAN_GENERIC) return;
  const RoiGeom g = roi_geometry(rois, levels, lv, r, PH, PW, ratio);
  const int H = lv.H[g.level], W = lv.W[g.level];
  Tables t = carve(smem, PH, PW, Hs, Ws);
  build_axis_table(t.Wy, t.ylo, t.yhi, PH, H, Hs, g.start_h, g.bin_h, g.grid_h);
  build_axis_table(t.Wx, t.xlo, t.xhi, PW, W, Ws, g.start_w, g.bin_w, g.grid_w);

  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv * V >= C) return;
  const T* __restrict__ img = static_cast<const T*>(lv.ptr[g.level]) + (size_t)g.batch * H * W * C + (size_t)cv * V;
  T* o = out + (size_t)r * PH * PW * C + (size_t)cv * V;
  for (int ph = 0; ph < PH; ph++) {
    const int y0 = t.ylo[ph], y1 = t.yhi[ph];
    const float* wy = t.Wy + (size_t)ph * Hs;
    for (int pw = 0; pw < PW; pw++) {
      const int x0 = t.xlo[pw], x1 = t.xhi[pw];
      const float* wx = t.Wx + (size_t)pw * Ws;
      float acc[V];
#pragma unroll
      for (int k = 0; k < V; k++) acc[k] = 0.f;
      for (int y = y0; y <= y1; y++) {
        const float a = wy[y];
        const T* row = img + (size_t)y * W * C;
        float racc[V];
#pragma unroll
        for (int k = 0; k < V; k++) racc[k] = 0.f;
        for (int x = x0; x <= x1; x++) {
          float v[V];
          VecIO<T, V>::load(row + (size_t)x * C, v);
          const float b = wx[x];
#pragma unroll
          for (int k = 0; k < V; k++) racc[k] = fmaf(b, v[k], racc[k]);
        }
#pragma unroll
        for (int k = 0; k < V; k++) acc[k] = fmaf(a, racc[k], acc[k]);
      }
#pragma unroll
      for (int k = 0; k < V; k++) acc[k] = acc[k] / g.count;
      VecIO<T, V>::store_stream(o + ((size_t)ph * PW + pw) * C, acc);
    }
  }
}

// ------------------------------------------------------------------------------------------ forward, NCHW
template <typename T>
__global__ void __launch_bounds__(256) roi_align_fwd_nchw_kernel(LevelTable lv, const float* __restrict__ rois,
                                                                const int32_t* __restrict__ levels,
                                                                T* __restrict__ out, int C, int PH, int PW, int ratio,
                                                                int Hs, int Ws, int cchunk) {
  extern __shared__ float smem[];
  const int r = blockIdx.x;
  const RoiGeom g = roi_geometry(rois, levels, lv, r, PH, PW, ratio);
  const int H = lv.H[g.level], W = lv.W[g.level];
  Tables t = carve(smem, PH, PW, Hs, Ws);
  build_axis_table(t.Wy, t.ylo, t.yhi, PH, H, Hs, g.start_h, g.bin_h, g.grid_h);
  build_axis_table(t.Wx, t.xlo, t.xhi, PW, W, Ws, g.start_w, g.bin_w, g.grid_w);

  const int c0 = blockIdx.y * cchunk;
  const int nC = min(cchunk, C - c0);
  const int nbin = PH * PW;
  const T* __restrict__ img = static_cast<const T*>(lv.ptr[g.level]) + ((size_t)g.batch * C + c0) * H * W;
  T* o = out + ((size_t)r * C + c0) * nbin;
  for (int e = threadIdx.x; e < nC * nbin; e += blockDim.x) {
    const int c = e / nbin, bin = e - c * nbin;
    const int ph = bin / PW, pw = bin - ph * PW;
    const int y0 = t.ylo[ph], y1 = t.yhi[ph], x0 = t.xlo[pw], x1 = t.xhi[pw];
    const float* wy = t.Wy + (size_t)ph * Hs;
    const float* wx = t.Wx + (size_t)pw * Ws;
    const T* plane = img + (size_t)c * H * W;
    float acc = 0.f;
    for (int y = y0; y <= y1; y++) {
      const T* row = plane + (size_t)y * W;
      float racc = 0.f;
      for (int x = x0; x <= x1; x++) {
        float v[1];
        VecIO<T, 1>::load(row + x, v);
        racc = fmaf(wx[x], v[0], racc);
      }
      acc = fmaf(wy[y], racc, acc);
    }
    float res[1] = {acc / g.count};
    VecIO<T, 1>::store(o + e, res);
  }
}

// ------------------------------------------------------------------------------------------ backward, NHWC, self-contained
// Gather over the RoI's footprint: each map pixel collects its contributing bins and is updated by one vector
// reduction per RoI.  Used when the caller passes no workspace, and for the RoIs a plan marks GENERIC.
template <typename T, int V>
__global__ void __launch_bounds__(256) roi_align_bwd_nhwc_kernel(LevelTable lv, const float* __restrict__ rois,
                                                                const int32_t* __restrict__ levels,
                                                                const T* __restrict__ gout, int C, int PH, int PW,
                                                                int ratio, int Hs, int Ws, const int* __restrict__ plans,
                                                                size_t plan_stride) {
  extern __shared__ float smem[];
  const int r = blockIdx.x;
  if (plans && plans[(size_t)r * plan_stride] != PLAN_GENERIC) return;
  const RoiGeom g = roi_geometry(rois, levels, lv, r, PH, PW, ratio);
  const int H = lv.H[g.level], W = lv.W[g.level];
  Tables t = carve(smem, PH, PW, Hs, Ws);
  build_axis_table(t.Wy, t.ylo, t.yhi, PH, H, Hs, g.start_h, g.bin_h, g.grid_h);
  build_axis_table(t.Wx, t.xlo, t.xhi, PW, W, Ws, g.start_w, g.bin_w, g.grid_w);
  build_inverse_ranges(t, PH, PW, H, W, Hs, Ws);
  const int* plo = t.aux;
  const int* phi = plo + Hs;
  const int* qlo = phi + Hs;
  const int* qhi = qlo + Ws;
  const int* fp = qhi + Ws;

  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv * V >= C) return;
  T* gin = static_cast<T*>(lv.ptr[g.level]) + (size_t)g.batch * H * W * C + (size_t)cv * V;
  const T* __restrict__ go = gout + (size_t)r * PH * PW * C + (size_t)cv * V;
  const float inv = 1.f / g.count;
  for (int y = fp[0]; y <= fp[1]; y++) {
    const int p0 = plo[y], p1 = phi[y];
    for (int x = fp[2]; x <= fp[3]; x++) {
      const int q0 = qlo[x], q1 = qhi[x];
      float acc[V];
#pragma unroll
      for (int k = 0; k < V; k++) acc[k] = 0.f;
      float wsum = 0.f;
      for (int p = p0; p <= p1; p++) {
        const float a = t.Wy[(size_t)p * Hs + y];
        for (int q = q0; q <= q1; q++) {
          const float w = a * t.Wx[(size_t)q * Ws + x];
          float v[V];
          VecIO<T, V>::load(go + ((size_t)p * PW + q) * C, v);
#pragma unroll
          for (int k = 0; k < V; k++) acc[k] = fmaf(w, v[k], acc[k]);
          wsum += w;
        }
      }
      if (wsum != 0.f) {  // warp-uniform: the weights do not depend on the channel
#pragma unroll
        for (int k = 0; k < V; k++) acc[k] *= inv;
        VecIO<T, V>::red_add(gin + ((size_t)y * W + x) * C, acc);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ backward, NCHW
template <typename T>
__global__ void __launch_bounds__(256) roi_align_bwd_nchw_kernel(LevelTable lv, const float* __restrict__ rois,
                                                                const int32_t* __restrict__ levels,
                                                                const T* __restrict__ gout, int C, int PH, int PW,
                                                                int ratio, int Hs, int Ws, int cchunk) {
  extern __shared__ float smem[];
  const int r = blockIdx.x;
  const RoiGeom g = roi_geometry(rois, levels, lv, r, PH, PW, ratio);
  const int H = lv.H[g.level], W = lv.W[g.level];
  Tables t = carve(smem, PH, PW, Hs, Ws);
  build_axis_table(t.Wy, t.ylo, t.yhi, PH, H, Hs, g.start_h, g.bin_h, g.grid_h);
  build_axis_table(t.Wx, t.xlo, t.xhi, PW, W, Ws, g.start_w, g.bin_w, g.grid_w);
  build_inverse_ranges(t, PH, PW, H, W, Hs, Ws);
  const int* plo = t.aux;
  const int* phi = plo + Hs;
  const int* qlo = phi + Hs;
  const int* qhi = qlo + Ws;
  const int* fp = qhi + Ws;
  const int fh = fp[1] - fp[0] + 1, fw = fp[3] - fp[2] + 1;
  if (fh <= 0 || fw <= 0) return;

  const int c0 = blockIdx.y * cchunk;
  const int nC = min(cchunk, C - c0);
  const int nbin = PH * PW, npix = fh * fw;
  T* gin = static_cast<T*>(lv.ptr[g.level]) + ((size_t)g.batch * C + c0) * H * W;
  const T* __restrict__ go = gout + ((size_t)r * C + c0) * nbin;
  const float inv = 1.f / g.count;
  for (int e = threadIdx.x; e < nC * npix; e += blockDim.x) {
    const int c = e / npix, pix = e - c * npix;
    const int y = fp[0] + pix / fw, x = fp[2] + pix % fw;
    const T* gc = go + (size_t)c * nbin;
    float acc = 0.f, wsum = 0.f;
    for (int p = plo[y]; p <= phi[y]; p++) {
      const float a = t.Wy[(size_t)p * Hs + y];
      for (int q = qlo[x]; q <= qhi[x]; q++) {
        const float w = a * t.Wx[(size_t)q * Ws + x];
        float v[1];
        VecIO<T, 1>::load(gc + p * PW + q, v);
        acc = fmaf(w, v[0], acc);
        wsum += w;
      }
    }
    if (wsum != 0.f) {
      float res[1] = {acc * inv};
      VecIO<T, 1>::red_add(gin + ((size_t)c * H + y) * W + x, res);
    }
  }
}

// ------------------------------------------------------------------------------------------ FPN level mapper
__global__ void fpn_map_levels_kernel(const float* __restrict__ rois, int32_t* __restrict__ levels, int R, float k_min,
                                      float k_max, float s0, float lvl0, float eps) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float* b = rois + 5 * (size_t)r + 1;
  // BoxList.area(), +1 convention (structures/bounding_box.py:227-231); modeling/poolers.py:37-42
  const float area = __fmul_rn(__fadd_rn(__fsub_rn(b[2], b[0]), 1.f), __fadd_rn(__fsub_rn(b[3], b[1]), 1.f));
  const float s = sqrtf(area);
  float lvl = floorf(__fadd_rn(lvl0, log2f(__fadd_rn(__fdiv_rn(s, s0), eps))));
  lvl = fminf(fmaxf(lvl, k_min), k_max);
  levels[r] = (int32_t)((long long)lvl - (long long)k_min);
}

// ------------------------------------------------------------------------------------------ NCHW staging
// src [batch][rows][cols] -> dst [batch][cols][rows] (32x32 tiles through padded shared memory, coalesced both ways).
// With `add` the transposed values are accumulated into dst.  Used to run NCHW callers through the channels-last
// kernels: [B][C][HW] <-> [B][HW][C] for the maps, [R][C][P*P] <-> [R][P*P][C] for the pooled tensors.
template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ src, T* __restrict__ dst, int rows, int cols,
                                                       long long batch, int add) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (long long b = blockIdx.z; b < batch; b += gridDim.z) {
    const T* s = src + (size_t)b * rows * cols;
    T* d = dst + (size_t)b * rows * cols;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      const int r = r0 + ty + i, c = c0 + tx;
      if (r < rows && c < cols) {
        float v[1];
        VecIO<T, 1>::load(s + (size_t)r * cols + c, v);
        tile[ty + i][tx] = v[0];
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      const int c = c0 + ty + i, r = r0 + tx;
      if (r < rows && c < cols) {
        float v[1] = {tile[tx][ty + i]};
        T* p = d + (size_t)c * rows + r;
        if (add) {
          float o[1];
          VecIO<T, 1>::load(p, o);
          v[0] += o[0];
        }
        VecIO<T, 1>::store(p, v);
      }
    }
    __syncthreads();
  }
}

// Pooled tensors ([R][C][HW] <-> [R][HW][C], HW = 49 or 196): a CTA moves one RoI's chunk of CH channels (as many as
// ~100 KB of shared memory hold, two CTAs per SM).  The NCHW side of the chunk is one contiguous run (CH*HW elements),
// the NHWC side is HW runs of CH channels (2 KB each for CH = 512); the chunk sits in shared memory as [c][HW | 1]
// (odd pitch: conflict-free in both directions).
constexpr int kPoolThreads = 512;
template <typename T, int CH>
__global__ void __launch_bounds__(kPoolThreads) transpose_pooled_kernel(const T* __restrict__ src, T* __restrict__ dst, int C, int HW, int to_nhwc) {
  extern __shared__ float chunk[];
  const int r = blockIdx.y, c0 = blockIdx.x * CH;
  const int nC = min(CH, C - c0), pitch = HW | 1, total = nC * HW;
  const size_t nchw = ((size_t)r * C + c0) * HW, nhwc = (size_t)r * HW * C + c0;
  if constexpr (std::is_same<T, float>::value) {
    // full fp32 chunks move as 16-byte vectors on both global sides (the shared-memory side stays scalar)
    if (nC == CH && C % 4 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
      const int total4 = total / 4;
      if (to_nhwc) {
        const float4* s4 = reinterpret_cast<const float4*>(src + nchw);
        for (int i4 = threadIdx.x; i4 < total4; i4 += kPoolThreads) {
          const float4 v = __ldg(s4 + i4);
          const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int i = i4 * 4 + k, c = i / HW;
            chunk[c * pitch + (i - c * HW)] = e[k];
          }
        }
        __syncthreads();
        for (int i4 = threadIdx.x; i4 < total4; i4 += kPoolThreads) {
          const int p2 = i4 / (CH / 4), c = (i4 - p2 * (CH / 4)) * 4;
          const float4 v = make_float4(chunk[c * pitch + p2], chunk[(c + 1) * pitch + p2], chunk[(c + 2) * pitch + p2], chunk[(c + 3) * pitch + p2]);
          *reinterpret_cast<float4*>(dst + nhwc + (size_t)p2 * C + c) = v;
        }
      } else {
        for (int i4 = threadIdx.x; i4 < total4; i4 += kPoolThreads) {
          const int p2 = i4 / (CH / 4), c = (i4 - p2 * (CH / 4)) * 4;
          const float4 v = __ldg(reinterpret_cast<const float4*>(src + nhwc + (size_t)p2 * C + c));
          chunk[c * pitch + p2] = v.x; chunk[(c + 1) * pitch + p2] = v.y; chunk[(c + 2) * pitch + p2] = v.z; chunk[(c + 3) * pitch + p2] = v.w;
        }
        __syncthreads();
        float4* d4 = reinterpret_cast<float4*>(dst + nchw);
        for (int i4 = threadIdx.x; i4 < total4; i4 += kPoolThreads) {
          float e[4];
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int i = i4 * 4 + k, c = i / HW;
            e[k] = chunk[c * pitch + (i - c * HW)];
          }
          d4[i4] = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
      return;
    }
  }
  if (to_nhwc) {
    for (int i = threadIdx.x; i < total; i += kPoolThreads) {
      const int c = i / HW, p2 = i - c * HW;
      float v[1];
      VecIO<T, 1>::load(src + nchw + i, v);
      chunk[c * pitch + p2] = v[0];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += kPoolThreads) {
      const int p2 = i / nC, c = i - p2 * nC;
      const float v[1] = {chunk[c * pitch + p2]};
      VecIO<T, 1>::store(dst + nhwc + (size_t)p2 * C + c, v);
    }
  } else {
    for (int i = threadIdx.x; i < total; i += kPoolThreads) {
      const int p2 = i / nC, c = i - p2 * nC;
      float v[1];
      VecIO<T, 1>::load(src + nhwc + (size_t)p2 * C + c, v);
      chunk[c * pitch + p2] = v[0];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += kPoolThreads) {
      const int c = i / HW, p2 = i - c * HW;
      const float v[1] = {chunk[c * pitch + p2]};
      VecIO<T, 1>::store(dst + nchw + i, v);
    }
  }
}

constexpr size_t kPoolSmem = 101 * 1024;  // per CTA: two CTAs per SM
static bool pooled_transpose_fits(int HW, int R) { return (size_t)64 * (HW | 1) * sizeof(float) <= kPoolSmem && R <= 65535; }

template <typename T, int CH>
static int launch_transpose_pooled(const void* src, void* dst, int R, int C, int HW, int to_nhwc, cudaStream_t st) {
  const size_t smem = (size_t)CH * (HW | 1) * sizeof(float);
  auto kern = transpose_pooled_kernel<T, CH>;
  if (smem > 48 * 1024) ABR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3(ceil_div(C, CH), R), kPoolThreads, smem, st>>>(static_cast<const T*>(src), static_cast<T*>(dst), C, HW, to_nhwc);
  ABR_CHECK_LAUNCH("roi_align_transpose_pooled");
  return ABR_OK;
}

// [R][C][HW] -> [R][HW][C] (to_nhwc) or back
template <typename T>
static int transpose_pooled_t(const void* src, void* dst, int R, int C, int HW, int to_nhwc, cudaStream_t st) {
  const size_t per_ch = (size_t)(HW | 1) * sizeof(float);
  if (C >= 512 && 512 * per_ch <= kPoolSmem) return launch_transpose_pooled<T, 512>(src, dst, R, C, HW, to_nhwc, st);
  if (C >= 128 && 128 * per_ch <= kPoolSmem) return launch_transpose_pooled<T, 128>(src, dst, R, C, HW, to_nhwc, st);
  return launch_transpose_pooled<T, 64>(src, dst, R, C, HW, to_nhwc, st);
}
static int transpose_pooled(const void* src, void* dst, int R, int C, int HW, int to_nhwc, int dtype, cudaStream_t st) {
  if (R <= 0) return ABR_OK;
  ABR_REQUIRE(pooled_transpose_fits(HW, R), ABR_ERR_UNSUPPORTED, "roi_align: pooled transpose of %d RoIs x %d positions", R, HW);
  return dtype == ABR_F32 ? transpose_pooled_t<float>(src, dst, R, C, HW, to_nhwc, st)
                          : transpose_pooled_t<__nv_bfloat16>(src, dst, R, C, HW, to_nhwc, st);
}

template <typename T>
static int launch_transpose(const void* src, void* dst, int rows, int cols, long long batch, int add, cudaStream_t st) {
  if (rows <= 0 || cols <= 0 || batch <= 0) return ABR_OK;
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32), (unsigned)(batch < 65535 ? batch : 65535));
  transpose_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T*>(src), static_cast<T*>(dst), rows, cols, batch, add);
  ABR_CHECK_LAUNCH("roi_align_transpose");
  return ABR_OK;
}
static int transpose_any(const void* src, void* dst, int rows, int cols, long long batch, int add, int dtype, cudaStream_t st) {
  return dtype == ABR_F32 ? launch_transpose<float>(src, dst, rows, cols, batch, add, st)
                          : launch_transpose<__nv_bfloat16>(src, dst, rows, cols, batch, add, st);
}

// ------------------------------------------------------------------------------------------ host side
static int check_common(const void* a, const float* rois, const void* b, int B, int C, int R, int PH, int PW,
                        int dtype, int layout) {
  ABR_REQUIRE(B >= 0 && C > 0 && R >= 0 && PH > 0 && PW > 0, ABR_ERR_BAD_ARG,
              "roi_align: bad sizes B=%d C=%d R=%d PH=%d PW=%d", B, C, R, PH, PW);
  ABR_REQUIRE(dtype == ABR_F32 || dtype == ABR_BF16, ABR_ERR_UNSUPPORTED, "roi_align: dtype %d not supported", dtype);
  ABR_REQUIRE(layout == ABR_NCHW || layout == ABR_NHWC || layout == ABR_NCHW_MAPS_NHWC_POOLED, ABR_ERR_UNSUPPORTED,
              "roi_align: layout %d not supported", layout);
  if (R > 0) ABR_REQUIRE(a && rois && b, ABR_ERR_BAD_ARG, "roi_align: null pointer");
  return ABR_OK;
}

static int fill_levels(LevelTable& lv, void* const* ptrs, const int* hs, const int* ws, const float* scales, int L,
                       int& Hs, int& Ws) {
  ABR_REQUIRE(L >= 1 && L <= ABR_MAX_LEVELS, ABR_ERR_BAD_ARG, "roi_align: %d levels (max %d)", L, ABR_MAX_LEVELS);
  Hs = Ws = 0;
  for (int l = 0; l < L; l++) {
    ABR_REQUIRE(ptrs[l] && hs[l] > 0 && ws[l] > 0, ABR_ERR_BAD_ARG, "roi_align: level %d: null map or empty size", l);
    lv.ptr[l] = ptrs[l];
    lv.H[l] = hs[l];
    lv.W[l] = ws[l];
    lv.scale[l] = scales[l];
    Hs = hs[l] > Hs ? hs[l] : Hs;
    Ws = ws[l] > Ws ? ws[l] : Ws;
  }
  for (int l = L; l < ABR_MAX_LEVELS; l++) { lv.ptr[l] = nullptr; lv.H[l] = lv.W[l] = 0; lv.scale[l] = 0.f; }
  return ABR_OK;
}

template <typename K>
static int set_smem(K kernel, size_t bytes, const char* name) {
  ABR_REQUIRE(bytes <= 227 * 1024, ABR_ERR_UNSUPPORTED, "%s: interpolation tables need %zu B of shared memory (> 227 KB)",
              name, bytes);
  if (bytes > 48 * 1024) ABR_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return ABR_OK;
}

static inline int nchw_channel_chunk(int R, int C) {
  // enough CTAs to fill 148 SMs several times over, while amortising the table build over many channels
  int chunk = 64;
  while (chunk > 8 && (long long)R * ceil_div(C, chunk) < 4LL * num_sms()) chunk >>= 1;
  return chunk < C ? chunk : C;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

struct Call {
  int B, L;
  LevelTable lv;
  const float* rois;
  const int32_t* levels;
  int C, R, PH, PW, ratio, Hs, Ws, layout;
  int* plans;  // null: no workspace, self-contained kernels only
  bool plan_ready;  // the workspace already holds the plans of exactly these RoIs / levels / geometry
  cudaStream_t st;
};

// Which plan format / kernel family serves an NHWC call: the gather-form v2 kernels (roi_v2.cu: any output size up to
// 16x16) or the TMA-staged / sweep kernels of this file (the staged ones need PH <= 8, PW <= 7).  Depends only on the output
// size, so a forward and the backward that reuses its plans always agree.  abr_set_option("roi_v2", 0 / 1) overrides (measurements, tests).
static bool use_v2(int PH, int PW) {
  const int forced = options().roi_v2;
  if (!v2_supported(PH, PW) || forced == 0) return false;
  if (forced == 1) return true;
  return !(PH <= kMaxBins && PW <= 7);
}

static size_t workspace_need(int R, int PW, int Hs, int PH) {
  const size_t v1 = (size_t)R * plan_stride_words(PW, Hs, PH) * sizeof(int);
  const size_t v2 = v2_supported(PH, PW) ? v2_workspace_bytes(R, PH, PW) : 0;
  return v1 > v2 ? v1 : v2;
}

// Tensor maps of every level's [B*H rows][W pixels][C channels] view, one per box shape.  False when the driver entry point is
// missing or a map cannot be encoded (the caller then uses the sweep kernel).
// Encoding a tensor map costs a few microseconds of host time (six per level and call: 23-41 us per call in round 1, which
// bounds the small per-image calls of the real step), and a map depends only on (pointer, C, W, rows, element size): a small
// per-thread cache of the most recent (level view -> its six box-shape maps) makes repeated calls on the same feature-map
// buffers (the teacher / student pair, the backward after the forward, a caching allocator handing the same block back)
// skip the driver.
struct TmaCacheEntry {
  const void* ptr;
  cuuint64_t c, w, rows;
  int esize;
  unsigned long long stamp;
  CUtensorMap m[kTmaBoxes];
};
constexpr int kTmaCacheSize = 32;

template <typename T>
static bool encode_level(EncodeTiledFn enc, const void* ptr, cuuint64_t C, cuuint64_t W, cuuint64_t rows, CUtensorMap* out) {
  static thread_local TmaCacheEntry cache[kTmaCacheSize];
  static thread_local unsigned long long clock = 0;
  int victim = 0;
  for (int i = 0; i < kTmaCacheSize; i++) {
    TmaCacheEntry& e = cache[i];
    if (e.stamp && e.ptr == ptr && e.c == C && e.w == W && e.rows == rows && e.esize == (int)sizeof(T)) {
      e.stamp = ++clock;
      for (int b = 0; b < kTmaBoxes; b++) out[b] = e.m[b];
      return true;
    }
    if (e.stamp < cache[victim].stamp) victim = i;
  }
  const cuuint64_t dims[3] = {C, W, rows};
  const cuuint64_t strides[2] = {C * sizeof(T), W * C * sizeof(T)};
  if (strides[0] % 16 != 0 || (reinterpret_cast<uintptr_t>(ptr) & 15)) return false;
  TmaCacheEntry e;
  for (int b = 0; b < kTmaBoxes; b++) {
    const cuuint32_t box[3] = {(cuuint32_t)(512 / sizeof(T)), (cuuint32_t)(2 << b), (cuuint32_t)(kTileMaxPx >> (b + 1))};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult rc = enc(&e.m[b], sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                            const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return false;
  }
  e.ptr = ptr; e.c = C; e.w = W; e.rows = rows; e.esize = (int)sizeof(T); e.stamp = ++clock;
  cache[victim] = e;
  for (int b = 0; b < kTmaBoxes; b++) out[b] = e.m[b];
  return true;
}

template <typename T>
static bool build_tma_maps(const Call& c, TmaMaps& maps) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc || c.L > kTmaLevels) return false;
  for (int l = 0; l < kTmaLevels; l++) {
    const int ll = l < c.L ? l : 0;
    if (!encode_level<T>(enc, c.lv.ptr[ll], (cuuint64_t)c.C, (cuuint64_t)c.lv.W[ll], (cuuint64_t)c.B * c.lv.H[ll], maps.m[l])) return false;
  }
  return true;
}

constexpr int kPlanThreads = 64;  // per RoI: the table builds are barrier-bound, two warps measured best (32: 0.2258, 64: 0.2255, 128: 0.2280 ms per teacher forward)
static int run_plan(const Call& c) {
  if (c.plan_ready) return ABR_OK;
  const size_t smem = tables_bytes(c.PH, c.PW, c.Hs, c.Ws, true);
  int rc = set_smem(plan_kernel, smem, "roi_align plan");
  if (rc) return rc;
  plan_kernel<<<c.R, kPlanThreads, smem, c.st>>>(c.lv, c.rois, c.levels, c.plans, plan_stride_words(c.PW, c.Hs, c.PH), c.PH, c.PW, c.ratio, c.Hs, c.Ws);
  ABR_CHECK_LAUNCH("roi_align_plan");
  return ABR_OK;
}

template <typename T, int V>
static int launch_fwd(const Call& c, void* out) {
  const size_t smem = tables_bytes(c.PH, c.PW, c.Hs, c.Ws, false);
  if (c.layout == ABR_NHWC) {
    const size_t stride = plan_stride_words(c.PW, c.Hs, c.PH);
    if (c.plans && use_v2(c.PH, c.PW)) {
      if (!c.plan_ready) {
        int rc = v2_plan(c.lv, c.rois, c.levels, c.plans, c.R, c.PH, c.PW, c.ratio, c.st);
        if (rc) return rc;
      }
      return v2_forward(c.lv, c.plans, c.rois, c.levels, out, c.C, c.R, c.PH, c.PW, c.ratio,
                        std::is_same<T, float>::value ? ABR_F32 : ABR_BF16, c.st);
    }
    if (c.plans) {
      int rc = run_plan(c);
      if (rc) return rc;
      const int nslices = ceil_div(c.C, 32 * V);
      bool staged = false;
      if constexpr (V * sizeof(T) == 16) {
        const bool use_tma = options().fwd_tma != 0;
        const size_t planw = (size_t)kPlanHdr + (size_t)c.PW * kPlanCol + (size_t)c.Hs * 8;
        const size_t tma_smem = (size_t)kRing * kTileMaxPx * 32 * 16 + 2 * planw * 4 + (2 * kRing + 4) * 8;
        TmaMaps maps;
        if (use_tma && c.PH <= kMaxBins && c.PW <= 7 && tma_smem <= 200 * 1024 && build_tma_maps<T>(c, maps)) {
          auto kern = roi_align_fwd_tma_kernel<T, V>;
          ABR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem));
          int per_sm = 0;
          if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, tma_smem) != cudaSuccess || per_sm < 1) per_sm = 1;
          long long blocks = (long long)per_sm * num_sms();
          if (blocks > (long long)c.R * nslices) blocks = (long long)c.R * nslices;
          kern<<<(unsigned)blocks, 256, tma_smem, c.st>>>(maps, c.plans, stride, static_cast<T*>(out), c.C, c.PH, c.PW, c.R, nslices, c.Hs);
          ABR_CHECK_LAUNCH("roi_align_forward_tma");
          staged = true;
        }
      }
      if (!staged) {
        const long long ntasks = (long long)c.R * c.PW * nslices;
        const long long blocks = ceil_div<long long>(ntasks, 8);
        ABR_REQUIRE(blocks <= 0x7fffffffLL, ABR_ERR_UNSUPPORTED, "roi_align_forward: too many tasks");
        if (c.PH <= kThinBins)
          roi_align_fwd_sweep_kernel<T, V, kThinBins><<<(unsigned)blocks, 256, 0, c.st>>>(c.lv, c.plans, stride, static_cast<T*>(out), c.C, c.PH, c.PW, nslices, ntasks);
        else
          roi_align_fwd_sweep_kernel<T, V, kThinBinsMax><<<(unsigned)blocks, 256, 0, c.st>>>(c.lv, c.plans, stride, static_cast<T*>(out), c.C, c.PH, c.PW, nslices, ntasks);
        ABR_CHECK_LAUNCH("roi_align_forward_sweep");
      }
    }
    const int nvec = ceil_div(c.C, V);
    const int threads = min(256, ceil_div(nvec, 32) * 32);
    dim3 grid(c.R, ceil_div(nvec, threads));
    int rc = set_smem(roi_align_fwd_nhwc_kernel<T, V>, smem, "roi_align_forward");
    if (rc) return rc;
    roi_align_fwd_nhwc_kernel<T, V><<<grid, threads, smem, c.st>>>(c.lv, c.rois, c.levels, static_cast<T*>(out), c.C, c.PH, c.PW, c.ratio, c.Hs, c.Ws, c.plans, stride);
  } else {
    const int chunk = nchw_channel_chunk(c.R, c.C);
    dim3 grid(c.R, ceil_div(c.C, chunk));
    int rc = set_smem(roi_align_fwd_nchw_kernel<T>, smem, "roi_align_forward");
    if (rc) return rc;
    roi_align_fwd_nchw_kernel<T><<<grid, 256, smem, c.st>>>(c.lv, c.rois, c.levels, static_cast<T*>(out), c.C, c.PH, c.PW, c.ratio, c.Hs, c.Ws, chunk);
  }
  ABR_CHECK_LAUNCH("roi_align_forward");
  return ABR_OK;
}

template <typename T, int V>
static int launch_bwd(const Call& c, const void* gout) {
  const size_t smem = tables_bytes(c.PH, c.PW, c.Hs, c.Ws, true);
  if (c.layout == ABR_NHWC) {
    const size_t stride = plan_stride_words(c.PW, c.Hs, c.PH);
    if (c.plans && use_v2(c.PH, c.PW)) {
      if (!c.plan_ready) {
        int rc = v2_plan(c.lv, c.rois, c.levels, c.plans, c.R, c.PH, c.PW, c.ratio, c.st);
        if (rc) return rc;
      }
      return v2_backward(c.lv, c.plans, c.rois, c.levels, gout, c.C, c.R, c.PH, c.PW, c.ratio,
                         std::is_same<T, float>::value ? ABR_F32 : ABR_BF16, c.st);
    }
    if (c.plans) {
      int rc = run_plan(c);
      if (rc) return rc;
      const int nslices = ceil_div(c.C, 32 * V);
      bool staged = false;
      if constexpr (V * sizeof(T) == 16) {
        const bool use_tma = options().bwd_tma != 0;
        const int nbin = c.PH * c.PW;
        const size_t gstage = ((size_t)nbin * 512 + 127) & ~(size_t)127;
        const size_t planw = (size_t)kPlanHdr + (size_t)c.PW * kPlanCol + (size_t)c.Hs * kPlanRow + kPlanPix;
        const size_t tma_smem = 2 * gstage + 2 * planw * 4 + 4 * 8;
        EncodeTiledFn enc = encode_tiled_fn();
        CUtensorMap gmap;
        bool ok = use_tma && enc && nbin <= 256 && c.PW <= 7 && tma_smem <= 200 * 1024 && (reinterpret_cast<uintptr_t>(gout) & 15) == 0 &&
                  ((size_t)c.C * sizeof(T)) % 16 == 0;
        if (ok) {
          const cuuint64_t dims[3] = {(cuuint64_t)c.C, (cuuint64_t)nbin, (cuuint64_t)c.R};
          const cuuint64_t strides[2] = {(cuuint64_t)c.C * sizeof(T), (cuuint64_t)nbin * c.C * sizeof(T)};
          const cuuint32_t box[3] = {(cuuint32_t)(512 / sizeof(T)), (cuuint32_t)nbin, 1};
          const cuuint32_t estr[3] = {1, 1, 1};
          ok = enc(&gmap, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(gout),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
        if (ok) {
          auto kern = roi_align_bwd_tma_kernel<T, V>;
          ABR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem));
          int per_sm = 0;
          if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, tma_smem) != cudaSuccess || per_sm < 1) per_sm = 1;
          long long blocks = (long long)per_sm * num_sms();
          if (blocks > (long long)c.R * nslices) blocks = (long long)c.R * nslices;
          kern<<<(unsigned)blocks, 256, tma_smem, c.st>>>(gmap, c.lv, c.plans, stride, c.C, c.PH, c.PW, c.R, nslices, c.Hs);
          ABR_CHECK_LAUNCH("roi_align_backward_tma");
          staged = true;
        }
      }
      if (!staged) {
        const long long ntasks = (long long)c.R * c.PW * nslices;
        const long long blocks = ceil_div<long long>(ntasks, 8);
        ABR_REQUIRE(blocks <= 0x7fffffffLL, ABR_ERR_UNSUPPORTED, "roi_align_backward: too many tasks");
        if (c.PH <= kThinBins)
          roi_align_bwd_sweep_kernel<T, V, kThinBins><<<(unsigned)blocks, 256, 0, c.st>>>(c.lv, c.plans, stride, static_cast<const T*>(gout), c.C, c.PH, c.PW, nslices, ntasks);
        else
          roi_align_bwd_sweep_kernel<T, V, kThinBinsMax><<<(unsigned)blocks, 256, 0, c.st>>>(c.lv, c.plans, stride, static_cast<const T*>(gout), c.C, c.PH, c.PW, nslices, ntasks);
        ABR_CHECK_LAUNCH("roi_align_backward_sweep");
      }
    }
    const int nvec = ceil_div(c.C, V);
    const int threads = min(256, ceil_div(nvec, 32) * 32);
    dim3 grid(c.R, ceil_div(nvec, threads));
    int rc = set_smem(roi_align_bwd_nhwc_kernel<T, V>, smem, "roi_align_backward");
    if (rc) return rc;
    roi_align_bwd_nhwc_kernel<T, V><<<grid, threads, smem, c.st>>>(c.lv, c.rois, c.levels, static_cast<const T*>(gout), c.C, c.PH, c.PW, c.ratio, c.Hs, c.Ws, c.plans, stride);
  } else {
    const int chunk = nchw_channel_chunk(c.R, c.C);
    dim3 grid(c.R, ceil_div(c.C, chunk));
    int rc = set_smem(roi_align_bwd_nchw_kernel<T>, smem, "roi_align_backward");
    if (rc) return rc;
    roi_align_bwd_nchw_kernel<T><<<grid, 256, smem, c.st>>>(c.lv, c.rois, c.levels, static_cast<const T*>(gout), c.C, c.PH, c.PW, c.ratio, c.Hs, c.Ws, chunk);
  }
  ABR_CHECK_LAUNCH("roi_align_backward");
  return ABR_OK;
}

static int dispatch_fwd(const Call& c, void* out, int dtype) {
  if (dtype == ABR_F32) {
    if (c.layout == ABR_NHWC && c.C % 4 == 0) return launch_fwd<float, 4>(c, out);
    return launch_fwd<float, 1>(c, out);
  }
  if (c.layout == ABR_NHWC && c.C % 8 == 0) return launch_fwd<__nv_bfloat16, 8>(c, out);
  return launch_fwd<__nv_bfloat16, 1>(c, out);
}

static int dispatch_bwd(const Call& c, const void* gout, int dtype) {
  if (dtype == ABR_F32) {
    if (c.layout == ABR_NHWC && c.C % 4 == 0) return launch_bwd<float, 4>(c, gout);
    return launch_bwd<float, 1>(c, gout);
  }
  if (c.layout == ABR_NHWC && c.C % 8 == 0) return launch_bwd<__nv_bfloat16, 8>(c, gout);
  return launch_bwd<__nv_bfloat16, 1>(c, gout);
}

static size_t elem_size(int dtype) { return dtype == ABR_F32 ? 4 : 2; }

// NCHW callers can be run through the channels-last kernels when the workspace also has room for a channels-last copy
// of every level's map and of the pooled tensor (after the plans, 256-byte aligned).
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static size_t staging_need(int B, int C, long long sum_hw, int R, int PH, int PW, int dtype, int layout = ABR_NCHW) {
  const size_t maps = align256((size_t)B * C * sum_hw * (dtype == ABR_F32 ? 4 : 2));
  if (layout == ABR_NCHW_MAPS_NHWC_POOLED) return maps;  // the pooled tensor already is channels-last
  return maps + align256((size_t)R * C * PH * PW * (dtype == ABR_F32 ? 4 : 2));
}

// The plans are only used by the NHWC kernels; a workspace that is absent or too small selects the self-contained path.
static int* usable_workspace(void* ws, size_t bytes, int R, int PH, int PW, int Hs, int layout) {
  if (!ws || layout != ABR_NHWC || (reinterpret_cast<uintptr_t>(ws) & 15)) return nullptr;
  return bytes >= workspace_need(R, PW, Hs, PH) ? static_cast<int*>(ws) : nullptr;
}

}  // namespace abr

using namespace abr;

extern "C" {

size_t abr_roi_align_workspace_bytes(int R, int PH, int PW, int max_h) {
  if (R <= 0 || PW <= 0 || max_h <= 0) return 0;
  return workspace_need(R, PW, max_h, PH);
}

size_t abr_roi_align_workspace_bytes_nchw(int R, int PH, int PW, int max_h, int B, int C, long long sum_hw, int dtype) {
  if (R <= 0 || PW <= 0 || PH <= 0 || max_h <= 0 || B <= 0 || C <= 0 || sum_hw <= 0) return 0;
  return align256(workspace_need(R, PW, max_h, PH)) + staging_need(B, C, sum_hw, R, PH, PW, dtype);
}

size_t abr_roi_align_workspace_bytes_layout(int R, int PH, int PW, int max_h, int B, int C, long long sum_hw, int dtype, int layout) {
  if (layout == ABR_NHWC) return abr_roi_align_workspace_bytes(R, PH, PW, max_h);
  if (R <= 0 || PW <= 0 || PH <= 0 || max_h <= 0 || B <= 0 || C <= 0 || sum_hw <= 0) return 0;
  return align256(workspace_need(R, PW, max_h, PH)) + staging_need(B, C, sum_hw, R, PH, PW, dtype, layout);
}

int abr_roi_align_multilevel_forward(const void* const* inputs_host, const int* hs_host, const int* ws_host,
                                     const float* scales_host, int L, const float* rois, const int32_t* levels,
                                     void* output, int B, int C, int R, int PH, int PW, int sampling_ratio, int dtype,
                                     int layout, void* workspace, size_t workspace_bytes, int workspace_has_plan,
                                     abr_stream_t stream) {
  ABR_REQUIRE(inputs_host && hs_host && ws_host && scales_host, ABR_ERR_BAD_ARG, "roi_align: null level arrays");
  int rc = check_common(inputs_host, rois, output, B, C, R, PH, PW, dtype, layout);
  if (rc) return rc;
  if (R == 0) return ABR_OK;  // ROIAlign_cuda.cu:278-281
  ABR_REQUIRE(L == 1 || levels, ABR_ERR_BAD_ARG, "roi_align: %d levels but no per-RoI level array", L);
  Call c;
  rc = fill_levels(c.lv, const_cast<void* const*>(reinterpret_cast<const void* const*>(inputs_host)), hs_host, ws_host,
                   scales_host, L, c.Hs, c.Ws);
  if (rc) return rc;
  c.rois = rois; c.levels = L == 1 ? nullptr : levels;
  c.B = B; c.L = L;
  c.C = C; c.R = R; c.PH = PH; c.PW = PW; c.ratio = sampling_ratio; c.layout = layout;
  c.plan_ready = false;  // set below once the plans' location is known
  c.st = static_cast<cudaStream_t>(stream);
  long long sum_hw = 0;
  for (int l = 0; l < L; l++) sum_hw += (long long)hs_host[l] * ws_host[l];
  const size_t plan_bytes = align256(workspace_need(R, PW, c.Hs, PH));
  const bool mixed = layout == ABR_NCHW_MAPS_NHWC_POOLED;
  const bool staged_ok = workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 &&
                         workspace_bytes >= plan_bytes + staging_need(B, C, sum_hw, R, PH, PW, dtype, layout);
  ABR_REQUIRE(!mixed || staged_ok, ABR_ERR_WORKSPACE, "roi_align: layout %d needs a 256-byte aligned workspace of %zu B", layout,
              plan_bytes + staging_need(B, C, sum_hw, R, PH, PW, dtype, layout));
  if (layout != ABR_NHWC && staged_ok) {
    // NCHW maps with staging room: channels-last copies of the maps, channels-last kernels, transposed result
    const size_t es = elem_size(dtype);
    char* stage = static_cast<char*>(workspace) + plan_bytes;
    for (int l = 0; l < L; l++) {
      const int hw = hs_host[l] * ws_host[l];
      rc = transpose_any(inputs_host[l], stage, C, hw, B, 0, dtype, c.st);
      if (rc) return rc;
      c.lv.ptr[l] = stage;
      stage += (size_t)B * C * hw * es;
    }
    void* pooled = static_cast<char*>(workspace) + plan_bytes + align256((size_t)B * C * sum_hw * es);
    c.layout = ABR_NHWC;
    c.plans = static_cast<int*>(workspace);
    c.plan_ready = workspace_has_plan != 0;
    if (mixed) return dispatch_fwd(c, output, dtype);  // the caller's pooled tensor is channels-last already
    rc = dispatch_fwd(c, pooled, dtype);
    if (rc) return rc;
    if (pooled_transpose_fits(PH * PW, R)) return transpose_pooled(pooled, output, R, C, PH * PW, 0, dtype, c.st);
    return transpose_any(pooled, output, PH * PW, C, R, 0, dtype, c.st);
  }
  c.plans = usable_workspace(workspace, workspace_bytes, R, PH, PW, c.Hs, layout);
  c.plan_ready = c.plans != nullptr && workspace_has_plan != 0;
  return dispatch_fwd(c, output, dtype);
}

int abr_roi_align_multilevel_backward(const void* grad_output, const float* rois, const int32_t* levels,
                                      void* const* grad_inputs_host, const int* hs_host, const int* ws_host,
                                      const float* scales_host, int L, int B, int C, int R, int PH, int PW,
                                      int sampling_ratio, int dtype, int layout, int zero_init, void* workspace,
                                      size_t workspace_bytes, int workspace_has_plan, abr_stream_t stream) {
  ABR_REQUIRE(grad_inputs_host && hs_host && ws_host && scales_host, ABR_ERR_BAD_ARG, "roi_align: null level arrays");
  int rc = check_common(grad_inputs_host, rois, grad_output, B, C, R, PH, PW, dtype, layout);
  if (rc) return rc;
  Call c;
  c.st = static_cast<cudaStream_t>(stream);
  rc = fill_levels(c.lv, grad_inputs_host, hs_host, ws_host, scales_host, L, c.Hs, c.Ws);
  if (rc) return rc;
  if (zero_init)
    for (int l = 0; l < L; l++)
      ABR_CUDA_OK(cudaMemsetAsync(c.lv.ptr[l], 0, (size_t)B * C * c.lv.H[l] * c.lv.W[l] * elem_size(dtype), c.st));
  if (R == 0) return ABR_OK;  // ROIAlign_cuda.cu:323-326
  ABR_REQUIRE(L == 1 || levels, ABR_ERR_BAD_ARG, "roi_align: %d levels but no per-RoI level array", L);
  c.rois = rois; c.levels = L == 1 ? nullptr : levels;
  c.B = B; c.L = L;
  c.C = C; c.R = R; c.PH = PH; c.PW = PW; c.ratio = sampling_ratio; c.layout = layout;
  long long sum_hw = 0;
  for (int l = 0; l < L; l++) sum_hw += (long long)hs_host[l] * ws_host[l];
  const size_t plan_bytes = align256(workspace_need(R, PW, c.Hs, PH));
  const bool mixed = layout == ABR_NCHW_MAPS_NHWC_POOLED;
  const bool staged_ok = workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 &&
                         workspace_bytes >= plan_bytes + staging_need(B, C, sum_hw, R, PH, PW, dtype, layout);
  ABR_REQUIRE(!mixed || staged_ok, ABR_ERR_WORKSPACE, "roi_align: layout %d needs a 256-byte aligned workspace of %zu B", layout,
              plan_bytes + staging_need(B, C, sum_hw, R, PH, PW, dtype, layout));
  if (layout != ABR_NHWC && staged_ok) {
    // NCHW maps with staging room: channels-last copy of the upstream gradient, channels-last kernels into zeroed
    // channels-last maps, then transposed (added when zero_init == 0) into the caller's gradient maps
    const size_t es = elem_size(dtype);
    char* stage0 = static_cast<char*>(workspace) + plan_bytes;
    const void* pooled = grad_output;  // mixed layout: already channels-last
    if (!mixed) {
      void* staged = stage0 + align256((size_t)B * C * sum_hw * es);
      if (pooled_transpose_fits(PH * PW, R))
        rc = transpose_pooled(grad_output, staged, R, C, PH * PW, 1, dtype, c.st);
      else
        rc = transpose_any(grad_output, staged, C, PH * PW, R, 0, dtype, c.st);
      if (rc) return rc;
      pooled = staged;
    }
    ABR_CUDA_OK(cudaMemsetAsync(stage0, 0, (size_t)B * C * sum_hw * es, c.st));
    void* user[ABR_MAX_LEVELS];
    char* stage = stage0;
    for (int l = 0; l < L; l++) {
      user[l] = c.lv.ptr[l];
      c.lv.ptr[l] = stage;
      stage += (size_t)B * C * hs_host[l] * ws_host[l] * es;
    }
    c.layout = ABR_NHWC;
    c.plans = static_cast<int*>(workspace);
    c.plan_ready = workspace_has_plan != 0;
    rc = dispatch_bwd(c, pooled, dtype);
    if (rc) return rc;
    stage = stage0;
    for (int l = 0; l < L; l++) {
      const int hw = hs_host[l] * ws_host[l];
      // the zero_init memset of the caller's maps above makes "add" and "store" equivalent; skip the read when zeroed
      rc = transpose_any(stage, user[l], hw, C, B, zero_init ? 0 : 1, dtype, c.st);
      if (rc) return rc;
      stage += (size_t)B * C * hw * es;
    }
    return ABR_OK;
  }
  c.plans = usable_workspace(workspace, workspace_bytes, R, PH, PW, c.Hs, layout);
  c.plan_ready = c.plans != nullptr && workspace_has_plan != 0;
  return dispatch_bwd(c, grad_output, dtype);
}

int abr_roi_align_forward(const void* input, const float* rois, void* output, int B, int C, int H, int W, int R, int PH,
                          int PW, float spatial_scale, int sampling_ratio, int dtype, int layout, void* workspace,
                          size_t workspace_bytes, int workspace_has_plan, abr_stream_t stream) {
  if (R > 0) ABR_REQUIRE(input && H > 0 && W > 0, ABR_ERR_BAD_ARG, "roi_align_forward: null or empty input");
  if (R == 0) return check_common(&input, rois, output, B, C, R, PH, PW, dtype, layout);
  const void* ptrs[1] = {input};
  return abr_roi_align_multilevel_forward(ptrs, &H, &W, &spatial_scale, 1, rois, nullptr, output, B, C, R, PH, PW,
                                          sampling_ratio, dtype, layout, workspace, workspace_bytes, workspace_has_plan, stream);
}

int abr_roi_align_backward(const void* grad_output, const float* rois, void* grad_input, int B, int C, int H, int W,
                           int R, int PH, int PW, float spatial_scale, int sampling_ratio, int dtype, int layout,
                           int zero_init, void* workspace, size_t workspace_bytes, int workspace_has_plan,
                           abr_stream_t stream) {
  ABR_REQUIRE(grad_input || (size_t)B * C * H * W == 0, ABR_ERR_BAD_ARG, "roi_align_backward: null grad_input");
  if ((size_t)B * C * H * W == 0) return ABR_OK;
  void* ptrs[1] = {grad_input};
  return abr_roi_align_multilevel_backward(grad_output, rois, nullptr, ptrs, &H, &W, &spatial_scale, 1, B, C, R, PH, PW,
                                           sampling_ratio, dtype, layout, zero_init, workspace, workspace_bytes, workspace_has_plan,
                                           stream);
}

int abr_fpn_map_levels(const float* rois, int32_t* levels, int R, float k_min, float k_max, float canonical_scale,
                       float canonical_level, float eps, abr_stream_t stream) {
  ABR_REQUIRE(R >= 0, ABR_ERR_BAD_ARG, "fpn_map_levels: R=%d", R);
  if (R == 0) return ABR_OK;
  ABR_REQUIRE(rois && levels, ABR_ERR_BAD_ARG, "fpn_map_levels: null pointer");
  fpn_map_levels_kernel<<<ceil_div(R, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(rois, levels, R, k_min, k_max,
                                                                                         canonical_scale, canonical_level, eps);
  ABR_CHECK_LAUNCH("fpn_map_levels");
  return ABR_OK;
}

}  // extern "C"
