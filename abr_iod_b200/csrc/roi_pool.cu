// roi_pool.cu -- ROIPool (max pooling) forward / backward for sm_100a.
//
// Semantics: maskrcnn_benchmark/csrc/cuda/ROIPool_cuda.cu:16-108 of the reference: RoI corners rounded half away
// from zero, +1 on width/height, bin [floor(p*bin), ceil((p+1)*bin)) + start clipped to the map, empty bin -> 0 with
// argmax -1, strict '>' so the first (row-major) maximum wins, argmax = h*W+w as int32.
// One thread per output element; the element order follows the storage order of `output` so stores (and, for NHWC,
// every map read: consecutive threads are consecutive channels of the same pixel) are coalesced.
#include <cfloat>

#include "common.cuh"

namespace abr {

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void stf(T* p, float v);
template <>
__device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T, bool NHWC>
__global__ void __launch_bounds__(256) roi_pool_fwd_kernel(const T* __restrict__ in, const float* __restrict__ rois,
                                                          T* __restrict__ out, int32_t* __restrict__ argmax, long long n,
                                                          int C, int H, int W, int PH, int PW, float scale) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    int r, c, ph, pw;
    if (NHWC) {
      c = (int)(idx % C);
      pw = (int)((idx / C) % PW);
      ph = (int)((idx / C / PW) % PH);
      r = (int)(idx / C / PW / PH);
    } else {
      pw = (int)(idx % PW);
      ph = (int)((idx / PW) % PH);
      c = (int)((idx / PW / PH) % C);
      r = (int)(idx / PW / PH / C);
    }
    const float* roi = rois + 5 * (size_t)r;
    const int b = (int)roi[0];
    const int sw = (int)roundf(__fmul_rn(roi[1], scale)), sh = (int)roundf(__fmul_rn(roi[2], scale));
    const int ew = (int)roundf(__fmul_rn(roi[3], scale)), eh = (int)roundf(__fmul_rn(roi[4], scale));
    const int rw = max(ew - sw + 1, 1), rh = max(eh - sh + 1, 1);
    const float bin_h = __fdiv_rn((float)rh, (float)PH), bin_w = __fdiv_rn((float)rw, (float)PW);
    int hs = (int)floorf(__fmul_rn((float)ph, bin_h)), ws = (int)floorf(__fmul_rn((float)pw, bin_w));
    int he = (int)ceilf(__fmul_rn((float)(ph + 1), bin_h)), we = (int)ceilf(__fmul_rn((float)(pw + 1), bin_w));
    hs = min(max(hs + sh, 0), H);
    he = min(max(he + sh, 0), H);
    ws = min(max(ws + sw, 0), W);
    we = min(max(we + sw, 0), W);
    const bool empty = (he <= hs) || (we <= ws);
    float best = empty ? 0.f : -FLT_MAX;
    int besti = -1;
    const T* base = NHWC ? in + (size_t)b * H * W * C + c : in + ((size_t)b * C + c) * H * W;
    const size_t pix_stride = NHWC ? (size_t)C : 1;
    for (int h = hs; h < he; ++h)
      for (int w = ws; w < we; ++w) {
        const int i = h * W + w;
        const float v = ldf<T>(base + (size_t)i * pix_stride);
        if (v > best) { best = v; besti = i; }
      }
    stf<T>(out + idx, best);
    argmax[idx] = besti;
  }
}

template <typename T, bool NHWC>
__global__ void __launch_bounds__(256) roi_pool_bwd_kernel(const T* __restrict__ gout, const int32_t* __restrict__ argmax,
                                                          const float* __restrict__ rois, T* __restrict__ gin, long long n,
                                                          int C, int H, int W, int PH, int PW) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int a = argmax[idx];
    if (a == -1) continue;
    int r, c;
    if (NHWC) {
      c = (int)(idx % C);
      r = (int)(idx / C / PW / PH);
    } else {
      c = (int)((idx / PW / PH) % C);
      r = (int)(idx / PW / PH / C);
    }
    const int b = (int)rois[5 * (size_t)r];
    T* dst = NHWC ? gin + ((size_t)b * H * W + a) * C + c : gin + ((size_t)b * C + c) * H * W + a;
    float v[1] = {ldf<T>(gout + idx)};
    VecIO<T, 1>::red_add(dst, v);
  }
}

static int pool_check(const void* a, const float* rois, const void* b, int B, int C, int H, int W, int R, int PH, int PW,
                      int dtype, int layout) {
  ABR_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0 && R >= 0 && PH > 0 && PW > 0, ABR_ERR_BAD_ARG,
              "roi_pool: bad sizes B=%d C=%d H=%d W=%d R=%d PH=%d PW=%d", B, C, H, W, R, PH, PW);
  ABR_REQUIRE(dtype == ABR_F32 || dtype == ABR_BF16, ABR_ERR_UNSUPPORTED, "roi_pool: dtype %d not supported", dtype);
  ABR_REQUIRE(layout == ABR_NCHW || layout == ABR_NHWC, ABR_ERR_UNSUPPORTED, "roi_pool: layout %d not supported", layout);
  if (R > 0) ABR_REQUIRE(a && rois && b, ABR_ERR_BAD_ARG, "roi_pool: null pointer");
  return ABR_OK;
}

}  // namespace abr

using namespace abr;

extern "C" {

int abr_roi_pool_forward(const void* input, const float* rois, void* output, int32_t* argmax, int B, int C, int H, int W,
                         int R, int PH, int PW, float spatial_scale, int dtype, int layout, abr_stream_t stream) {
  int rc = pool_check(input, rois, output, B, C, H, W, R, PH, PW, dtype, layout);
  if (rc) return rc;
  if (R == 0) return ABR_OK;
  ABR_REQUIRE(argmax, ABR_ERR_BAD_ARG, "roi_pool_forward: null argmax");
  const long long n = (long long)R * C * PH * PW;
  const int blocks = (int)std::min<long long>(ceil_div<long long>(n, 256), (long long)num_sms() * 32);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define ABR_POOL_FWD(T, L) \
  roi_pool_fwd_kernel<T, L><<<blocks, 256, 0, st>>>(static_cast<const T*>(input), rois, static_cast<T*>(output), argmax, n, C, H, W, PH, PW, spatial_scale)
  if (dtype == ABR_F32) {
    if (layout == ABR_NHWC) ABR_POOL_FWD(float, true); else ABR_POOL_FWD(float, false);
  } else {
    if (layout == ABR_NHWC) ABR_POOL_FWD(__nv_bfloat16, true); else ABR_POOL_FWD(__nv_bfloat16, false);
  }
#undef ABR_POOL_FWD
  ABR_CHECK_LAUNCH("roi_pool_forward");
  return ABR_OK;
}

int abr_roi_pool_backward(const void* grad_output, const int32_t* argmax, const float* rois, void* grad_input, int B,
                          int C, int H, int W, int R, int PH, int PW, int dtype, int layout, int zero_init,
                          abr_stream_t stream) {
  int rc = pool_check(grad_output, rois, grad_input, B, C, H, W, R, PH, PW, dtype, layout);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t map_elems = (size_t)B * C * H * W;
  if (map_elems == 0) return ABR_OK;
  ABR_REQUIRE(grad_input, ABR_ERR_BAD_ARG, "roi_pool_backward: null grad_input");
  if (zero_init) ABR_CUDA_OK(cudaMemsetAsync(grad_input, 0, map_elems * (dtype == ABR_F32 ? 4 : 2), st));
  if (R == 0) return ABR_OK;
  ABR_REQUIRE(argmax, ABR_ERR_BAD_ARG, "roi_pool_backward: null argmax");
  const long long n = (long long)R * C * PH * PW;
  const int blocks = (int)std::min<long long>(ceil_div<long long>(n, 256), (long long)num_sms() * 32);
#define ABR_POOL_BWD(T, L) \
  roi_pool_bwd_kernel<T, L><<<blocks, 256, 0, st>>>(static_cast<const T*>(grad_output), argmax, rois, static_cast<T*>(grad_input), n, C, H, W, PH, PW)
  if (dtype == ABR_F32) {
    if (layout == ABR_NHWC) ABR_POOL_BWD(float, true); else ABR_POOL_BWD(float, false);
  } else {
    if (layout == ABR_NHWC) ABR_POOL_BWD(__nv_bfloat16, true); else ABR_POOL_BWD(__nv_bfloat16, false);
  }
#undef ABR_POOL_BWD
  ABR_CHECK_LAUNCH("roi_pool_backward");
  return ABR_OK;
}

}  // extern "C"
