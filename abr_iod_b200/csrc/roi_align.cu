// roi_align.cu -- ROIAlign forward / backward for sm_100a, single- and multi-level in one launch.
//
// Semantics: maskrcnn_benchmark/csrc/cuda/ROIAlign_cuda.cu:64-254 of the reference (== csrc/cpu/ROIAlign_cpu.cpp).
// Design (not a port): the reference gives every output scalar its own thread, which re-derives the RoI
// geometry and issues 4*g*g scattered 4-byte gathers.  Here a CTA owns one RoI and
//   1. builds, once, the two SEPARABLE interpolation tables of that RoI in shared memory:
//        Wy[ph][y] = sum over the bin's sample rows of the bilinear row weight landing on map row y
//        Wx[pw][x] = same along x
//      (the reference's weight w1..w4 of a sample is hy*hx, hy*lx, ly*hx, ly*lx and its "outside the
//      map => 0" rule is the AND of a y-test and an x-test, so  out = Wy * V * Wx^T / count  exactly);
//   2. NHWC path: every thread owns V consecutive channels (16-byte vectors: 4 x fp32 or 8 x bf16), so each
//      map pixel is one coalesced 512 B warp load shared by all its channels, every distinct pixel of a bin's
//      footprint is read once per bin instead of once per sample tap, and weights are warp-uniform broadcasts;
//      NCHW path (the reference's contiguous layout): threads walk the flat (c,ph,pw) output so stores are
//      fully coalesced, with the same tables.
//   3. backward is a GATHER over the RoI's footprint: each map pixel collects its (typically 2x2) contributing
//      bins and is updated by ONE vector reduction per RoI (red.global.add.v4.f32 / v4.bf16x2, a contiguous
//      512 B request per warp) instead of 4*g*g scalar atomicAdds per output element.
// Sample coordinates are evaluated with explicitly rounded fp32 intrinsics in the reference's operation order
// so that the in/out-of-map decisions (ROIAlign_cuda.cu:22-25) are identical to the reference's.
#include "common.cuh"

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
__device__ __forceinline__ RoiGeom roi_geometry(const float* __restrict__ rois, const int32_t* __restrict__ levels,
                                                const LevelTable& lv, int r, int PH, int PW, int ratio) {
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
  if (backward) b += (size_t)(2 * Hs + 2 * Ws + 8) * 4 + (size_t)Hs * 16 + 16;
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
// where span = max over rows of (number of bins containing the row) - 1.
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
    for (int y = a; y <= b; y++) span = max(span, phi[y] - plo[y]);
    fp[0] = a; fp[1] = b; fp[2] = c; fp[3] = d; fp[4] = span;
  }
  __syncthreads();
}

constexpr int kThinBins = 8;  // pooled heights up to this use the statically indexed thin-bin path

// Per-row record for the row-sweep kernels (valid when every row feeds at most two bins): the first bin `a` holding
// the row, its weight Wy[a][y] and the weight Wy[a+1][y] of the next bin (0 when the row is not shared).
__device__ __forceinline__ void build_row_info(const Tables& t, float4* rinfo, int PH, int H, int Hs) {
  const int* plo = t.aux;
  const int* phi = plo + Hs;
  for (int y = threadIdx.x; y < H; y += blockDim.x) {
    const int p0 = plo[y], p1 = phi[y];
    float4 v = make_float4(__int_as_float(-1), 0.f, 0.f, 0.f);
    if (p0 <= p1) {
      v.x = __int_as_float(p0);
      v.y = t.Wy[(size_t)p0 * Hs + y];
      v.z = (p1 > p0) ? t.Wy[(size_t)(p0 + 1) * Hs + y] : 0.f;
    }
    rinfo[y] = v;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------ forward, NHWC
// A WARP owns (RoI, output column pw, slice of 32*V channels): its lanes read the same map pixel at consecutive
// channels (one coalesced 512 B request), all weights are warp-uniform shared-memory broadcasts, and it sweeps the
// RoI's footprint rows ONCE: for every row y it forms  t = sum_x Wx[pw][x] * v[y][x]  and adds  Wy[ph][y] * t  to the
// (at most two) vertically adjacent bins that row belongs to, held in two rolling register accumulators.  Every
// distinct pixel of the column's footprint is therefore loaded once per RoI, not once per bin or per sample tap.
// RoIs whose rows feed more than two bins (bins thinner than one map pixel) take the plain per-bin loop.
template <typename T, int V>
__global__ void __launch_bounds__(896) roi_align_fwd_nhwc_kernel(LevelTable lv, const float* __restrict__ rois,
                                                                const int32_t* __restrict__ levels,
                                                                T* __restrict__ out, int C, int PH, int PW, int ratio,
                                                                int Hs, int Ws, int slices_per_cta) {
  extern __shared__ float smem[];
  const int r = blockIdx.x;
  const RoiGeom g = roi_geometry(rois, levels, lv, r, PH, PW, ratio);
  const int H = lv.H[g.level], W = lv.W[g.level];
  Tables t = carve(smem, PH, PW, Hs, Ws);
  build_axis_table(t.Wy, t.ylo, t.yhi, PH, H, Hs, g.start_h, g.bin_h, g.grid_h);
  build_axis_table(t.Wx, t.xlo, t.xhi, PW, W, Ws, g.start_w, g.bin_w, g.grid_w);
  build_inverse_ranges(t, PH, PW, H, W, Hs, Ws);
  const int* fp = t.aux + 2 * Hs + 2 * Ws;
  float4* rinfo = reinterpret_cast<float4*>((reinterpret_cast<uintptr_t>(t.aux + 2 * Hs + 2 * Ws + 8) + 15) & ~uintptr_t(15));
  build_row_info(t, rinfo, PH, H, Hs);
  const int Y0 = fp[0], Y1 = fp[1];
  const bool rolling = fp[4] <= 1;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const size_t pix = (size_t)C;                // elements between horizontally adjacent pixels
  const size_t rowstride = (size_t)W * C;      // ... and between vertically adjacent ones
  const size_t binstride = (size_t)PW * C;     // output elements between vertically adjacent bins
  const T* __restrict__ img0 = static_cast<const T*>(lv.ptr[g.level]) + (size_t)g.batch * H * W * C;
  T* out0 = out + (size_t)r * PH * PW * C;
  const float inv_count = 1.f / g.count;

  for (int task = warp; task < PW * slices_per_cta; task += nwarp) {
    const int pw = task % PW;
    const int c = ((blockIdx.y * slices_per_cta + task / PW) * 32 + lane) * V;
    if (c >= C) continue;
    const int x0 = t.xlo[pw], x1 = t.xhi[pw];
    const float* wx = t.Wx + (size_t)pw * Ws;
    const T* __restrict__ img = img0 + c;
    T* o = out0 + (size_t)pw * C + c;
    if (x0 > x1) {  // the whole column lies outside the map: every sample contributes 0
      float z[V];
#pragma unroll
      for (int k = 0; k < V; k++) z[k] = 0.f;
      for (int ph = 0; ph < PH; ph++) VecIO<T, V>::store(o + (size_t)ph * binstride, z);
      continue;
    }
    // the first four column weights live in registers (columns are rarely wider); loads beyond the column are
    // clamped onto its last pixel and carry weight 0, which keeps the row loop free of predicates
    const int nx = x1 - x0 + 1;
    const float w0 = wx[x0];
    const float w1 = nx > 1 ? wx[x0 + 1] : 0.f;
    const float w2 = nx > 2 ? wx[x0 + 2] : 0.f;
    const float w3 = nx > 3 ? wx[x0 + 3] : 0.f;
    const size_t o1 = (size_t)min(1, nx - 1) * pix, o2 = (size_t)min(2, nx - 1) * pix, o3 = (size_t)min(3, nx - 1) * pix;
    const T* row = img + (size_t)Y0 * rowstride + (size_t)x0 * pix;
    if (rolling) {
      float accA[V], accB[V];
#pragma unroll
      for (int k = 0; k < V; k++) accA[k] = accB[k] = 0.f;
      int a = 0;
      for (int y = Y0; y <= Y1; y++, row += rowstride) {
        const float4 info = rinfo[y];
        const int ra = __float_as_int(info.x);
        if (ra < 0) continue;
        float v0[V], v1[V], v2[V], v3[V];
        VecIO<T, V>::load(row, v0);
        VecIO<T, V>::load(row + o1, v1);
        VecIO<T, V>::load(row + o2, v2);
        VecIO<T, V>::load(row + o3, v3);
        if (a != ra) {  // bins a .. ra-1 are complete: emit them and roll the two-bin window
          do {
#pragma unroll
            for (int k = 0; k < V; k++) accA[k] *= inv_count;
            VecIO<T, V>::store(o + (size_t)a * binstride, accA);
#pragma unroll
            for (int k = 0; k < V; k++) { accA[k] = accB[k]; accB[k] = 0.f; }
          } while (++a < ra);
        }
        float tr[V];
#pragma unroll
        for (int k = 0; k < V; k++) tr[k] = fmaf(w3, v3[k], fmaf(w2, v2[k], fmaf(w1, v1[k], w0 * v0[k])));
        if (nx > 4) {
          const T* q = row + 4 * pix;
          for (int x = x0 + 4; x <= x1; x++, q += pix) {
            float v[V];
            VecIO<T, V>::load(q, v);
            const float b = wx[x];
#pragma unroll
            for (int k = 0; k < V; k++) tr[k] = fmaf(b, v[k], tr[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < V; k++) {
          accA[k] = fmaf(info.y, tr[k], accA[k]);
          accB[k] = fmaf(info.z, tr[k], accB[k]);
        }
      }
      for (; a < PH; a++) {
#pragma unroll
        for (int k = 0; k < V; k++) accA[k] *= inv_count;
        VecIO<T, V>::store(o + (size_t)a * binstride, accA);
#pragma unroll
        for (int k = 0; k < V; k++) { accA[k] = accB[k]; accB[k] = 0.f; }
      }
    } else if (PH <= kThinBins) {
      // thin bins (a row feeds three or more bins; the RoI is only a few rows tall): one accumulator per bin,
      // statically indexed, every row added to every bin with its table weight (0 outside the bin's support)
      float acc[kThinBins][V];
#pragma unroll
      for (int p = 0; p < kThinBins; p++)
#pragma unroll
        for (int k = 0; k < V; k++) acc[p][k] = 0.f;
      for (int y = Y0; y <= Y1; y++, row += rowstride) {
        float v0[V], v1[V], v2[V], v3[V];
        VecIO<T, V>::load(row, v0);
        VecIO<T, V>::load(row + o1, v1);
        VecIO<T, V>::load(row + o2, v2);
        VecIO<T, V>::load(row + o3, v3);
        float tr[V];
#pragma unroll
        for (int k = 0; k < V; k++) tr[k] = fmaf(w3, v3[k], fmaf(w2, v2[k], fmaf(w1, v1[k], w0 * v0[k])));
        if (nx > 4) {
          const T* q = row + 4 * pix;
          for (int x = x0 + 4; x <= x1; x++, q += pix) {
            float v[V];
            VecIO<T, V>::load(q, v);
            const float b = wx[x];
#pragma unroll
            for (int k = 0; k < V; k++) tr[k] = fmaf(b, v[k], tr[k]);
          }
        }
#pragma unroll
        for (int p = 0; p < kThinBins; p++) {
          const float wa = p < PH ? t.Wy[(size_t)p * Hs + y] : 0.f;
#pragma unroll
          for (int k = 0; k < V; k++) acc[p][k] = fmaf(wa, tr[k], acc[p][k]);
        }
      }
#pragma unroll
      for (int p = 0; p < kThinBins; p++) {
        if (p < PH) {
#pragma unroll
          for (int k = 0; k < V; k++) acc[p][k] *= inv_count;
          VecIO<T, V>::store(o + (size_t)p * binstride, acc[p]);
        }
      }
    } else {
      for (int ph = 0; ph < PH; ph++) {
        const float* wy = t.Wy + (size_t)ph * Hs;
        float acc[V];
#pragma unroll
        for (int k = 0; k < V; k++) acc[k] = 0.f;
        for (int y = t.ylo[ph]; y <= t.yhi[ph]; y++) {
          float tr[V];
#pragma unroll
          for (int k = 0; k < V; k++) tr[k] = 0.f;
          const T* q = img + ((size_t)y * W + x0) * pix;
          for (int x = x0; x <= x1; x++, q += pix) {
            float v[V];
            VecIO<T, V>::load(q, v);
            const float b = wx[x];
#pragma unroll
            for (int k = 0; k < V; k++) tr[k] = fmaf(b, v[k], tr[k]);
          }
          const float wa = wy[y];
#pragma unroll
          for (int k = 0; k < V; k++) acc[k] = fmaf(wa, tr[k], acc[k]);
        }
#pragma unroll
        for (int k = 0; k < V; k++) acc[k] *= inv_count;
        VecIO<T, V>::store(o + (size_t)ph * binstride, acc);
      }
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

// ------------------------------------------------------------------------------------------ backward, NHWC
// Same ownership as the forward: a warp = (RoI, column pw, 32*V channels).  For every footprint row y it folds the
// (usually two) bins of its column that contain the row into  s = sum_p Wy[p][y] * g[p][pw]  and issues one vector
// reduction  gin[y][x] += Wx[pw][x] * s / count  per footprint pixel of the column: a warp-wide, contiguous 512 B
// red.global.add.v4.f32 instead of the reference's 4*g*g scalar atomicAdds per output element.
template <typename T, int V>
__global__ void __launch_bounds__(896) roi_align_bwd_nhwc_kernel(LevelTable lv, const float* __restrict__ rois,
                                                                const int32_t* __restrict__ levels,
                                                                const T* __restrict__ gout, int C, int PH, int PW,
                                                                int ratio, int Hs, int Ws, int slices_per_cta) {
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
  const int* fp = phi + Hs + 2 * Ws;
  const int Y0 = fp[0], Y1 = fp[1];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const size_t pix = (size_t)C;
  T* gin0 = static_cast<T*>(lv.ptr[g.level]) + (size_t)g.batch * H * W * C;
  const T* __restrict__ go0 = gout + (size_t)r * PH * PW * C;
  const float inv_count = 1.f / g.count;

  for (int task = warp; task < PW * slices_per_cta; task += nwarp) {
    const int pw = task % PW;
    const int c = ((blockIdx.y * slices_per_cta + task / PW) * 32 + lane) * V;
    if (c >= C) continue;
    const int x0 = t.xlo[pw], x1 = t.xhi[pw];
    if (x0 > x1) continue;
    const float* wx = t.Wx + (size_t)pw * Ws;
    const T* __restrict__ go = go0 + (size_t)pw * C + c;
    T* gin = gin0 + c;
    int a = -2;  // no bin cached yet (a + 1 must not match any real bin)
    float gA[V], gB[V];  // grad of bins a and a+1 of this column (rolling)
#pragma unroll
    for (int k = 0; k < V; k++) gA[k] = gB[k] = 0.f;
    for (int y = Y0; y <= Y1; y++) {
      const int p0 = plo[y], p1 = phi[y];
      if (p0 > p1) continue;
      float sacc[V];
      if (p1 - p0 <= 1) {
        if (a != p0) {
          if (a + 1 == p0) {
#pragma unroll
            for (int k = 0; k < V; k++) gA[k] = gB[k];
          } else {
            VecIO<T, V>::load(go + (size_t)p0 * PW * C, gA);
          }
          a = p0;
          if (a + 1 < PH) {
            VecIO<T, V>::load(go + (size_t)(a + 1) * PW * C, gB);
          } else {
#pragma unroll
            for (int k = 0; k < V; k++) gB[k] = 0.f;
          }
        }
        const float wa = t.Wy[(size_t)a * Hs + y];
        const float wb = p1 > a ? t.Wy[(size_t)(a + 1) * Hs + y] : 0.f;
#pragma unroll
        for (int k = 0; k < V; k++) sacc[k] = fmaf(wb, gB[k], wa * gA[k]);
      } else {  // thin bins: several bins share the row
#pragma unroll
        for (int k = 0; k < V; k++) sacc[k] = 0.f;
        for (int p = p0; p <= p1; p++) {
          float v[V];
          VecIO<T, V>::load(go + (size_t)p * PW * C, v);
          const float w = t.Wy[(size_t)p * Hs + y];
#pragma unroll
          for (int k = 0; k < V; k++) sacc[k] = fmaf(w, v[k], sacc[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < V; k++) sacc[k] *= inv_count;
      T* row = gin + ((size_t)y * W + x0) * pix;
      for (int x = x0; x <= x1; x++, row += pix) {
        const float b = wx[x];
        if (b != 0.f) {
          float v[V];
#pragma unroll
          for (int k = 0; k < V; k++) v[k] = b * sacc[k];
          VecIO<T, V>::red_add(row, v);
        }
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

// ------------------------------------------------------------------------------------------ host side
static int check_common(const void* a, const float* rois, const void* b, int B, int C, int R, int PH, int PW,
                        int dtype, int layout) {
  ABR_REQUIRE(B >= 0 && C > 0 && R >= 0 && PH > 0 && PW > 0, ABR_ERR_BAD_ARG,
              "roi_align: bad sizes B=%d C=%d R=%d PH=%d PW=%d", B, C, R, PH, PW);
  ABR_REQUIRE(dtype == ABR_F32 || dtype == ABR_BF16, ABR_ERR_UNSUPPORTED, "roi_align: dtype %d not supported", dtype);
  ABR_REQUIRE(layout == ABR_NCHW || layout == ABR_NHWC, ABR_ERR_UNSUPPORTED, "roi_align: layout %d not supported", layout);
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

// NHWC kernels: one warp per (pw, slice of 32*V channels); a CTA carries `spc` slices, i.e. PW*spc warps (<= 28).
static inline void nhwc_launch_shape(int C, int V, int PW, int& spc, int& threads, int& gy) {
  const int nslices = ceil_div(C, 32 * V);
  spc = 28 / PW;
  if (spc < 1) spc = 1;
  if (spc > nslices) spc = nslices;
  int warps = PW * spc;
  if (warps > 28) warps = 28;  // wide poolers: warps loop over their tasks
  threads = warps * 32;
  gy = ceil_div(nslices, spc);
}

template <typename T, int V>
static int launch_fwd(const LevelTable& lv, const float* rois, const int32_t* levels, void* out, int C, int R, int PH,
                      int PW, int ratio, int Hs, int Ws, int layout, cudaStream_t st) {
  const size_t smem = tables_bytes(PH, PW, Hs, Ws, true);
  if (layout == ABR_NHWC) {
    int spc, threads, gy;
    nhwc_launch_shape(C, V, PW, spc, threads, gy);
    dim3 grid(R, gy);
    int rc = set_smem(roi_align_fwd_nhwc_kernel<T, V>, smem, "roi_align_forward");
    if (rc) return rc;
    roi_align_fwd_nhwc_kernel<T, V><<<grid, threads, smem, st>>>(lv, rois, levels, static_cast<T*>(out), C, PH, PW, ratio, Hs, Ws, spc);
  } else {
    const int chunk = nchw_channel_chunk(R, C);
    dim3 grid(R, ceil_div(C, chunk));
    int rc = set_smem(roi_align_fwd_nchw_kernel<T>, smem, "roi_align_forward");
    if (rc) return rc;
    roi_align_fwd_nchw_kernel<T><<<grid, 256, smem, st>>>(lv, rois, levels, static_cast<T*>(out), C, PH, PW, ratio, Hs, Ws, chunk);
  }
  ABR_CHECK_LAUNCH("roi_align_forward");
  return ABR_OK;
}

template <typename T, int V>
static int launch_bwd(const LevelTable& lv, const float* rois, const int32_t* levels, const void* gout, int C, int R,
                      int PH, int PW, int ratio, int Hs, int Ws, int layout, cudaStream_t st) {
  const size_t smem = tables_bytes(PH, PW, Hs, Ws, true);
  if (layout == ABR_NHWC) {
    int spc, threads, gy;
    nhwc_launch_shape(C, V, PW, spc, threads, gy);
    dim3 grid(R, gy);
    int rc = set_smem(roi_align_bwd_nhwc_kernel<T, V>, smem, "roi_align_backward");
    if (rc) return rc;
    roi_align_bwd_nhwc_kernel<T, V><<<grid, threads, smem, st>>>(lv, rois, levels, static_cast<const T*>(gout), C, PH, PW, ratio, Hs, Ws, spc);
  } else {
    const int chunk = nchw_channel_chunk(R, C);
    dim3 grid(R, ceil_div(C, chunk));
    int rc = set_smem(roi_align_bwd_nchw_kernel<T>, smem, "roi_align_backward");
    if (rc) return rc;
    roi_align_bwd_nchw_kernel<T><<<grid, 256, smem, st>>>(lv, rois, levels, static_cast<const T*>(gout), C, PH, PW, ratio, Hs, Ws, chunk);
  }
  ABR_CHECK_LAUNCH("roi_align_backward");
  return ABR_OK;
}

static int dispatch_fwd(const LevelTable& lv, const float* rois, const int32_t* levels, void* out, int C, int R, int PH,
                        int PW, int ratio, int Hs, int Ws, int dtype, int layout, cudaStream_t st) {
  if (dtype == ABR_F32) {
    if (layout == ABR_NHWC && C % 4 == 0) return launch_fwd<float, 4>(lv, rois, levels, out, C, R, PH, PW, ratio, Hs, Ws, layout, st);
    return launch_fwd<float, 1>(lv, rois, levels, out, C, R, PH, PW, ratio, Hs, Ws, layout, st);
  }
  if (layout == ABR_NHWC && C % 8 == 0)
    return launch_fwd<__nv_bfloat16, 8>(lv, rois, levels, out, C, R, PH, PW, ratio, Hs, Ws, layout, st);
  return launch_fwd<__nv_bfloat16, 1>(lv, rois, levels, out, C, R, PH, PW, ratio, Hs, Ws, layout, st);
}

static int dispatch_bwd(const LevelTable& lv, const float* rois, const int32_t* levels, const void* gout, int C, int R,
                        int PH, int PW, int ratio, int Hs, int Ws, int dtype, int layout, cudaStream_t st) {
  if (dtype == ABR_F32) {
    if (layout == ABR_NHWC && C % 4 == 0) return launch_bwd<float, 4>(lv, rois, levels, gout, C, R, PH, PW, ratio, Hs, Ws, layout, st);
    return launch_bwd<float, 1>(lv, rois, levels, gout, C, R, PH, PW, ratio, Hs, Ws, layout, st);
  }
  if (layout == ABR_NHWC && C % 8 == 0)
    return launch_bwd<__nv_bfloat16, 8>(lv, rois, levels, gout, C, R, PH, PW, ratio, Hs, Ws, layout, st);
  return launch_bwd<__nv_bfloat16, 1>(lv, rois, levels, gout, C, R, PH, PW, ratio, Hs, Ws, layout, st);
}

static size_t elem_size(int dtype) { return dtype == ABR_F32 ? 4 : 2; }

}  // namespace abr

using namespace abr;

extern "C" {

int abr_roi_align_multilevel_forward(const void* const* inputs_host, const int* hs_host, const int* ws_host,
                                     const float* scales_host, int L, const float* rois, const int32_t* levels,
                                     void* output, int B, int C, int R, int PH, int PW, int sampling_ratio, int dtype,
                                     int layout, abr_stream_t stream) {
  ABR_REQUIRE(inputs_host && hs_host && ws_host && scales_host, ABR_ERR_BAD_ARG, "roi_align: null level arrays");
  int rc = check_common(inputs_host, rois, output, B, C, R, PH, PW, dtype, layout);
  if (rc) return rc;
  if (R == 0) return ABR_OK;  // ROIAlign_cuda.cu:278-281
  ABR_REQUIRE(L == 1 || levels, ABR_ERR_BAD_ARG, "roi_align: %d levels but no per-RoI level array", L);
  LevelTable lv;
  int Hs, Ws;
  rc = fill_levels(lv, const_cast<void* const*>(reinterpret_cast<const void* const*>(inputs_host)), hs_host, ws_host,
                   scales_host, L, Hs, Ws);
  if (rc) return rc;
  return dispatch_fwd(lv, rois, L == 1 ? nullptr : levels, output, C, R, PH, PW, sampling_ratio, Hs, Ws, dtype, layout,
                      static_cast<cudaStream_t>(stream));
}

int abr_roi_align_multilevel_backward(const void* grad_output, const float* rois, const int32_t* levels,
                                      void* const* grad_inputs_host, const int* hs_host, const int* ws_host,
                                      const float* scales_host, int L, int B, int C, int R, int PH, int PW,
                                      int sampling_ratio, int dtype, int layout, int zero_init, abr_stream_t stream) {
  ABR_REQUIRE(grad_inputs_host && hs_host && ws_host && scales_host, ABR_ERR_BAD_ARG, "roi_align: null level arrays");
  int rc = check_common(grad_inputs_host, rois, grad_output, B, C, R, PH, PW, dtype, layout);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LevelTable lv;
  int Hs, Ws;
  rc = fill_levels(lv, grad_inputs_host, hs_host, ws_host, scales_host, L, Hs, Ws);
  if (rc) return rc;
  if (zero_init)
    for (int l = 0; l < L; l++)
      ABR_CUDA_OK(cudaMemsetAsync(lv.ptr[l], 0, (size_t)B * C * lv.H[l] * lv.W[l] * elem_size(dtype), st));
  if (R == 0) return ABR_OK;  // ROIAlign_cuda.cu:323-326
  ABR_REQUIRE(L == 1 || levels, ABR_ERR_BAD_ARG, "roi_align: %d levels but no per-RoI level array", L);
  return dispatch_bwd(lv, rois, L == 1 ? nullptr : levels, grad_output, C, R, PH, PW, sampling_ratio, Hs, Ws, dtype,
                      layout, st);
}

int abr_roi_align_forward(const void* input, const float* rois, void* output, int B, int C, int H, int W, int R, int PH,
                          int PW, float spatial_scale, int sampling_ratio, int dtype, int layout, abr_stream_t stream) {
  if (R > 0) ABR_REQUIRE(input && H > 0 && W > 0, ABR_ERR_BAD_ARG, "roi_align_forward: null or empty input");
  if (R == 0) return check_common(&input, rois, output, B, C, R, PH, PW, dtype, layout);
  const void* ptrs[1] = {input};
  return abr_roi_align_multilevel_forward(ptrs, &H, &W, &spatial_scale, 1, rois, nullptr, output, B, C, R, PH, PW,
                                          sampling_ratio, dtype, layout, stream);
}

int abr_roi_align_backward(const void* grad_output, const float* rois, void* grad_input, int B, int C, int H, int W,
                           int R, int PH, int PW, float spatial_scale, int sampling_ratio, int dtype, int layout,
                           int zero_init, abr_stream_t stream) {
  ABR_REQUIRE(grad_input || (size_t)B * C * H * W == 0, ABR_ERR_BAD_ARG, "roi_align_backward: null grad_input");
  if ((size_t)B * C * H * W == 0) return ABR_OK;
  void* ptrs[1] = {grad_input};
  return abr_roi_align_multilevel_backward(grad_output, rois, nullptr, ptrs, &H, &W, &spatial_scale, 1, B, C, R, PH, PW,
                                           sampling_ratio, dtype, layout, zero_init, stream);
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
