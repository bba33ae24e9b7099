// paste.cu -- ABR mixup / mosaic paste of Box-Rehearsal prototypes into a batch of images, one launch (sm_100a).
//
// Semantics: the pixel arithmetic of PascalVOCDataset_ABR._start_mixup (data/datasets/voc_abr.py:659-678) and
// _start_boxes_mosaic (:744-763,804) of the reference.  The random draws and the integer rectangle arithmetic stay
// on the host (abr_iod_b200/data/abr_paste.py reproduces the reference's draw order); the host hands this kernel a
// table of rectangles.  Blend is evaluated exactly like numpy does it: lambda*dst and (1-lambda)*src are two
// individually rounded float64 products, their sum is rounded once more, and the uint8 store truncates
// (__dmul_rn / __dadd_rn: no FMA contraction).  HBM-bound byte work: every destination pixel is visited by exactly
// one thread, which applies the image's ops in order (a later mixup box must see the earlier blend).
#include "common.cuh"

namespace abr {

__global__ void __launch_bounds__(256) paste_batch_kernel(uint8_t* __restrict__ canvas,
                                                         const abr_paste_image_t* __restrict__ images,
                                                         const abr_paste_op_t* __restrict__ ops,
                                                         const uint8_t* __restrict__ pool) {
  const abr_paste_image_t im = images[blockIdx.y];
  const int npix = im.height * im.width;
  uint8_t* dst = canvas + im.offset;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
    const int y = pix / im.width, x = pix - y * im.width;
    bool touched = false;
    double v[3] = {0.0, 0.0, 0.0};
    for (int k = 0; k < im.n_ops; k++) {
      const abr_paste_op_t& op = ops[im.first_op + k];
      if (y < op.y0 || y >= op.y1 || x < op.x0 || x >= op.x1) continue;
      if (!touched && op.kind == ABR_PASTE_BLEND) {
#pragma unroll
        for (int c = 0; c < 3; c++) v[c] = (double)dst[(size_t)pix * 3 + c];
      }
      touched = true;
      if (op.kind == ABR_PASTE_FILL) {
        v[0] = v[1] = v[2] = (double)op.fill;
      } else {
        const uint8_t* s = pool + op.src_offset + ((size_t)(op.sy0 + y - op.y0) * op.src_width + (op.sx0 + x - op.x0)) * 3;
        if (op.kind == ABR_PASTE_COPY) {
#pragma unroll
          for (int c = 0; c < 3; c++) v[c] = (double)s[c];
        } else {
          const double lam = op.lambda, one_minus = __dsub_rn(1.0, lam);
#pragma unroll
          for (int c = 0; c < 3; c++) {
            const double r = __dadd_rn(__dmul_rn(lam, v[c]), __dmul_rn(one_minus, (double)s[c]));
            v[c] = (double)(uint8_t)(int)r;  // numpy's float64 -> uint8 cast truncates (values are within 0..255)
          }
        }
      }
    }
    if (touched) {
#pragma unroll
      for (int c = 0; c < 3; c++) dst[(size_t)pix * 3 + c] = (uint8_t)(int)v[c];
    }
  }
}

}  // namespace abr

using namespace abr;

extern "C" int abr_paste_batch(uint8_t* canvas, const abr_paste_image_t* images, int n_images, const abr_paste_op_t* ops,
                               int n_ops, const uint8_t* pool, int max_pixels_per_image, abr_stream_t stream) {
  ABR_REQUIRE(n_images >= 0 && n_ops >= 0 && max_pixels_per_image >= 0, ABR_ERR_BAD_ARG, "paste: negative size");
  if (n_images == 0 || n_ops == 0 || max_pixels_per_image == 0) return ABR_OK;
  ABR_REQUIRE(canvas && images && ops, ABR_ERR_BAD_ARG, "paste: null pointer");
  ABR_REQUIRE(n_images <= 65535, ABR_ERR_UNSUPPORTED, "paste: %d images in one call (max 65535)", n_images);
  int bx = ceil_div(max_pixels_per_image, 256);
  const int cap = ceil_div(num_sms() * 8, n_images);
  if (bx > cap) bx = cap > 0 ? cap : 1;
  paste_batch_kernel<<<dim3(bx, n_images), 256, 0, static_cast<cudaStream_t>(stream)>>>(canvas, images, ops, pool);
  ABR_CHECK_LAUNCH("paste_batch");
  return ABR_OK;
}
