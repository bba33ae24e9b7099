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
  if (backward) b += (size_t)(2 * Hs + 2 * Ws + 4) * 4;
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

// ------------------------------------------------------------------------------------------ forward, NHWC
template <typename T, int V>
__global__ void __launch_bounds__(256) roi_align_fwd_nhwc_kernel(LevelTable lv, const float* __restrict__ rois,
                                                                const int32_t* __restrict__ levels,
                                                                T* __restrict__ out, int C, int PH, int PW, int ratio,
                                                                int Hs, int Ws) {
  extern __shared__ float smem[];
  const int r = blockIdx.x;
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
#pragma unroll 4
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
      VecIO<T, V>::store(o + ((size_t)ph * PW + pw) * C, acc);
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

// ------------------------------------------------------------------------------------------ backward helpers
// After the two axis tables exist: overall footprint [Y0,Y1]x[X0,X1] and, for every map row / column inside it,
// the (contiguous) range of bins whose support contains it.  aux = [plo Hs][phi Hs][qlo Ws][qhi Ws][Y0,Y1,X0,X1].
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
  if (tid == 0) {
    int a = H, b = -1, c = W, d = -1;
    for (int p = 0; p < PH; p++)
      if (t.ylo[p] <= t.yhi[p]) { a = min(a, t.ylo[p]); b = max(b, t.yhi[p]); }
    for (int p = 0; p < PW; p++)
      if (t.xlo[p] <= t.xhi[p]) { c = min(c, t.xlo[p]); d = max(d, t.xhi[p]); }
    fp[0] = a; fp[1] = b; fp[2] = c; fp[3] = d;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------ backward, NHWC
template <typename T, int V>
__global__ void __launch_bounds__(256) roi_align_bwd_nhwc_kernel(LevelTable lv, const float* __restrict__ rois,
                                                                const int32_t* __restrict__ levels,
                                                                const T* __restrict__ gout, int C, int PH, int PW,
                                                                int ratio, int Hs, int Ws) {
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

template <typename T, int V>
static int launch_fwd(const LevelTable& lv, const float* rois, const int32_t* levels, void* out, int C, int R, int PH,
                      int PW, int ratio, int Hs, int Ws, int layout, cudaStream_t st) {
  const size_t smem = tables_bytes(PH, PW, Hs, Ws, false);
  if (layout == ABR_NHWC) {
    const int nvec = ceil_div(C, V);
    const int threads = min(256, ceil_div(nvec, 32) * 32);
    dim3 grid(R, ceil_div(nvec, threads));
    int rc = set_smem(roi_align_fwd_nhwc_kernel<T, V>, smem, "roi_align_forward");
    if (rc) return rc;
    roi_align_fwd_nhwc_kernel<T, V><<<grid, threads, smem, st>>>(lv, rois, levels, static_cast<T*>(out), C, PH, PW, ratio, Hs, Ws);
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
    const int nvec = ceil_div(C, V);
    const int threads = min(256, ceil_div(nvec, 32) * 32);
    dim3 grid(R, ceil_div(nvec, threads));
    int rc = set_smem(roi_align_bwd_nhwc_kernel<T, V>, smem, "roi_align_backward");
    if (rc) return rc;
    roi_align_bwd_nhwc_kernel<T, V><<<grid, threads, smem, st>>>(lv, rois, levels, static_cast<const T*>(gout), C, PH, PW, ratio, Hs, Ws);
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
