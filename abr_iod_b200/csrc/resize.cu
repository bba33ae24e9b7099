// resize.cu -- bicubic resize of Box-Rehearsal prototypes on the device, bit-exact with Pillow (sm_100a).
//
// The reference rescales a prototype before pasting it (data/datasets/voc_abr.py:538-548: PIL `resize((int(s*w),
// int(s*h)))`, default filter = BICUBIC for RGB) on the host, inside the DataLoader workers.  This file restates
// Pillow's 8-bit resampling (third-party arithmetic the reference calls: Pillow src/libImaging/Resample.c,
// ImagingResampleHorizontal_8bpc / Vertical_8bpc) so that the crops never leave the GPU:
//   * per axis, a table of fixed-point taps (22 fractional bits) and a [first, count] window per output coordinate --
//     built on the host in float64 exactly as Pillow's precompute_coeffs / normalize_coeffs_8bpc do
//     (abr_iod_b200/data/resample.py), uploaded with the batch;
//   * horizontal pass into a uint8 intermediate [src_h][dst_w] (rounded and clipped, as Pillow does), then the vertical
//     pass; a pass whose size does not change is skipped, like Pillow's need_horizontal / need_vertical.  (Pillow 12
//     resamples columns first when a source is more than 100x taller than wide; the caller keeps such slivers on PIL.)
// One thread per output pixel (3 channels), one launch per pass for ALL crops of a batch.  Byte work, bound by the
// (tiny) traffic; bit-exactness against PIL is what the tests check.
#include "common.cuh"

namespace abr {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Pillow: PRECISION_BITS

__device__ __forceinline__ uint8_t clip8(int v) {  // Pillow: clip8_lookups[v >> PRECISION_BITS]
  v >>= kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// pass 0: horizontal (src -> tmp, or src -> dst when the height does not change); pass 1: vertical.
__global__ void __launch_bounds__(256) resize_pass_kernel(const uint8_t* __restrict__ pool, uint8_t* __restrict__ out,
                                                         const abr_resize_job_t* __restrict__ jobs,
                                                         const int32_t* __restrict__ taps, int pass) {
  const abr_resize_job_t j = jobs[blockIdx.y];
  const bool need_h = j.dst_w != j.src_w, need_v = j.dst_h != j.src_h;
  if (pass == 0 ? !need_h : !need_v) return;
  const uint8_t* src;
  uint8_t* dst;
  int in_w, out_h, out_w;
  if (pass == 0) {
    src = pool + j.src_offset;
    dst = out + (need_v ? j.tmp_offset : j.dst_offset);
    in_w = j.src_w; out_h = j.src_h; out_w = j.dst_w;
  } else {
    src = need_h ? out + j.tmp_offset : pool + j.src_offset;
    dst = out + j.dst_offset;
    in_w = j.dst_w; out_h = j.dst_h; out_w = j.dst_w;
  }
  const int32_t* tab = taps + (pass == 0 ? j.x_taps : j.y_taps);  // per output coordinate: first, count, ksize taps
  const int stride = 2 + (pass == 0 ? j.x_ksize : j.y_ksize);
  const int npix = out_h * out_w;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
    const int y = p / out_w, x = p - y * out_w;
    const int32_t* t = tab + (size_t)(pass == 0 ? x : y) * stride;
    const int first = t[0], count = t[1];
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    if (pass == 0) {
      const uint8_t* q = src + ((size_t)y * in_w + first) * 3;
      for (int k = 0; k < count; k++, q += 3) {
        const int w = t[2 + k];
        s0 += q[0] * w; s1 += q[1] * w; s2 += q[2] * w;
      }
    } else {
      const uint8_t* q = src + ((size_t)first * in_w + x) * 3;
      for (int k = 0; k < count; k++, q += (size_t)in_w * 3) {
        const int w = t[2 + k];
        s0 += q[0] * w; s1 += q[1] * w; s2 += q[2] * w;
      }
    }
    uint8_t* o = dst + (size_t)p * 3;
    o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
  }
}

}  // namespace abr

using namespace abr;

extern "C" int abr_resize_bicubic_batch(const uint8_t* pool, uint8_t* out, const abr_resize_job_t* jobs, int n_jobs,
                                        const int32_t* taps, int max_pixels, abr_stream_t stream) {
  ABR_REQUIRE(n_jobs >= 0 && max_pixels >= 0, ABR_ERR_BAD_ARG, "resize: negative size");
  if (n_jobs == 0 || max_pixels == 0) return ABR_OK;
  ABR_REQUIRE(pool && out && jobs && taps, ABR_ERR_BAD_ARG, "resize: null pointer");
  ABR_REQUIRE(n_jobs <= 65535, ABR_ERR_UNSUPPORTED, "resize: %d crops in one call (max 65535)", n_jobs);
  int bx = ceil_div(max_pixels, 256);
  const int cap = ceil_div(num_sms() * 8, n_jobs);
  if (bx > cap) bx = cap > 0 ? cap : 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int pass = 0; pass < 2; pass++) {
    resize_pass_kernel<<<dim3(bx, n_jobs), 256, 0, st>>>(pool, out, jobs, taps, pass);
    ABR_CHECK_LAUNCH(pass == 0 ? "resize_bicubic (horizontal)" : "resize_bicubic (vertical)");
  }
  return ABR_OK;
}
