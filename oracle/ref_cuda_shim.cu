// ref_cuda_shim.cu -- compiles the REFERENCE's own CUDA kernels, in place, into oracle/_ref/libabr_ref_cuda.so.
//
// TEST INFRASTRUCTURE ONLY.  No reference source is copied into this repository: the three translation units below
// (csrc/cuda/ROIAlign_cuda.cu, ROIPool_cuda.cu, nms.cu -- SURVEY 2.2 "the reference kernels recompiled as-is") are
// #included from where they lie under /root/reference.  They are the same-box GPU comparator of bench.py
// (`reference_cuda`) and the third parity opinion of tests/test_gpu_ref_cuda.py for the rows whose reference has no CPU
// implementation (ROIAlign backward, ROIPool, the '>' NMS).  The product never links this library.
//
// What keeps the unmodified sources compiling against torch 2.11:
//   * <THC/...> headers: three small stand-ins under oracle/thc_stub (THCudaMalloc -> the caching allocator, ...);
//   * AT_DISPATCH_FLOATING_TYPES(tensor.type(), ...): re-pointed at an overload taking either spelling (as in
//     oracle/ref_shim.cpp);
//   * `THCState* state = at::globalContext().lazyInitCUDA();` (nms.cu:83): lazyInitCUDA is re-pointed at a member-free
//     helper through a macro that is defined only after every torch header has been included.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/types.h>

#include <THC/THC.h>

namespace abr_ref_shim {
inline at::ScalarType scalar_type_of(at::ScalarType t) { return t; }
inline at::ScalarType scalar_type_of(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
struct Ctx {
  THCState* lazyInitCUDA() const {
    static THCState s;
    return &s;
  }
};
inline Ctx ctx() { return Ctx(); }
}  // namespace abr_ref_shim
namespace at {
inline ::abr_ref_shim::Ctx abr_ref_ctx() { return ::abr_ref_shim::Ctx(); }
}  // namespace at

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
  AT_DISPATCH_SWITCH(::abr_ref_shim::scalar_type_of(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))
#define globalContext abr_ref_ctx

#include "cuda/ROIAlign_cuda.cu"
#undef CUDA_1D_KERNEL_LOOP
#include "cuda/ROIPool_cuda.cu"
#include "cuda/nms.cu"
#undef globalContext

#include <cstdint>

namespace {
at::Tensor dev_f32(const void* p, at::IntArrayRef shape) {
  int dev = 0;
  cudaGetDevice(&dev);
  return at::from_blob(const_cast<void*>(p), shape, at::TensorOptions().dtype(at::kFloat).device(at::kCUDA, dev));
}
}  // namespace

#define REF_API extern "C" __attribute__((visibility("default")))

// All pointers are DEVICE pointers of contiguous fp32 NCHW tensors on the current device; kernels run on torch's current
// stream, results are copied into the caller's buffers on the same stream.
REF_API void ref_cuda_roi_align_forward(const float* in, const float* rois, float* out, int B, int C, int H, int W, int R,
                                        int PH, int PW, float scale, int ratio) {
  at::Tensor o = ROIAlign_forward_cuda(dev_f32(in, {B, C, H, W}), dev_f32(rois, {R, 5}), scale, PH, PW, ratio);
  dev_f32(out, {R, C, PH, PW}).copy_(o);
}
REF_API void ref_cuda_roi_align_backward(const float* grad, const float* rois, float* gin, int B, int C, int H, int W, int R,
                                         int PH, int PW, float scale, int ratio) {
  at::Tensor g = ROIAlign_backward_cuda(dev_f32(grad, {R, C, PH, PW}), dev_f32(rois, {R, 5}), scale, PH, PW, B, C, H, W, ratio);
  dev_f32(gin, {B, C, H, W}).copy_(g);
}
// The kernels alone, writing into reference-allocated tensors (no copy): what bench.py times.
REF_API void ref_cuda_roi_align_forward_nocopy(const float* in, const float* rois, int B, int C, int H, int W, int R, int PH,
                                               int PW, float scale, int ratio) {
  ROIAlign_forward_cuda(dev_f32(in, {B, C, H, W}), dev_f32(rois, {R, 5}), scale, PH, PW, ratio);
}
REF_API void ref_cuda_roi_align_backward_nocopy(const float* grad, const float* rois, int B, int C, int H, int W, int R, int PH,
                                                int PW, float scale, int ratio) {
  ROIAlign_backward_cuda(dev_f32(grad, {R, C, PH, PW}), dev_f32(rois, {R, 5}), scale, PH, PW, B, C, H, W, ratio);
}
REF_API void ref_cuda_roi_pool_forward(const float* in, const float* rois, float* out, int32_t* argmax, int B, int C, int H,
                                       int W, int R, int PH, int PW, float scale) {
  auto r = ROIPool_forward_cuda(dev_f32(in, {B, C, H, W}), dev_f32(rois, {R, 5}), scale, PH, PW);
  dev_f32(out, {R, C, PH, PW}).copy_(std::get<0>(r));
  int dev = 0;
  cudaGetDevice(&dev);
  at::from_blob(argmax, {R, C, PH, PW}, at::TensorOptions().dtype(at::kInt).device(at::kCUDA, dev)).copy_(std::get<1>(r));
}
REF_API void ref_cuda_roi_pool_backward(const float* grad, const float* in, const float* rois, const int32_t* argmax,
                                        float* gin, int B, int C, int H, int W, int R, int PH, int PW, float scale) {
  int dev = 0;
  cudaGetDevice(&dev);
  at::Tensor am = at::from_blob(const_cast<int32_t*>(argmax), {R, C, PH, PW}, at::TensorOptions().dtype(at::kInt).device(at::kCUDA, dev));
  at::Tensor g = ROIPool_backward_cuda(dev_f32(grad, {R, C, PH, PW}), dev_f32(in, {B, C, H, W}), dev_f32(rois, {R, 5}), am, scale,
                                       PH, PW, B, C, H, W);
  dev_f32(gin, {B, C, H, W}).copy_(g);
}
// boxes_scores [N,5] (x1,y1,x2,y2,score) as the reference's nms.h:20 builds it; keep (device, int64, room for N)
// receives ascending original indices; returns the count (the reference's nms_cuda synchronises the device itself).
REF_API int64_t ref_cuda_nms(const float* boxes_scores, int64_t n, float thr, int64_t* keep) {
  at::Tensor k = nms_cuda(dev_f32(boxes_scores, {n, 5}), thr);
  int dev = 0;
  cudaGetDevice(&dev);
  at::from_blob(keep, {k.numel()}, at::TensorOptions().dtype(at::kLong).device(at::kCUDA, dev)).copy_(k);
  return k.numel();
}
REF_API int ref_cuda_shim_version() { return 1; }
