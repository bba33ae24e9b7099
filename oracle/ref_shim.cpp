// ref_shim.cpp -- compiles the REFERENCE's own CPU ops, in place, into oracle/_ref/.
//
// TEST INFRASTRUCTURE ONLY (see oracle/abr_oracle.c header).  No reference
// source is copied into this repository: the two translation units below are
// #included from where they lie under /root/reference (include path given by
// oracle/Makefile).  They are used to pin oracle/abr_oracle.c and, on the GPU
// box, as the "reference" CPU baseline of bench.py.
//
// The reference passes `tensor.type()` (DeprecatedTypeProperties) to
// AT_DISPATCH_FLOATING_TYPES (csrc/cpu/ROIAlign_cpu.cpp:242,
// csrc/cpu/nms_cpu.cpp:71), which torch 2.11 no longer accepts.  Instead of
// patching a copy, the dispatch macro is re-pointed at an overload that takes
// either spelling; the reference code is compiled unmodified.
#include <torch/extension.h>

namespace abr_ref_shim {
inline at::ScalarType scalar_type_of(at::ScalarType t) { return t; }
inline at::ScalarType scalar_type_of(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace abr_ref_shim

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
  AT_DISPATCH_SWITCH(::abr_ref_shim::scalar_type_of(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))

#include "cpu/ROIAlign_cpu.cpp"  // ROIAlign_forward_cpu
#include "cpu/nms_cpu.cpp"       // nms_cpu

#include <cstdint>
#include <cstring>

extern "C" {

// in [B,C,H,W], rois [R,5], out [R,C,PH,PW]; contiguous fp32 host buffers.
__attribute__((visibility("default"))) void ref_roi_align_forward_cpu(
    const float* in, const float* rois, float* out, int B, int C, int H, int W, int R, int PH, int PW,
    float scale, int ratio) {
  auto opt = at::TensorOptions().dtype(at::kFloat);
  at::Tensor tin = at::from_blob(const_cast<float*>(in), {B, C, H, W}, opt);
  at::Tensor troi = at::from_blob(const_cast<float*>(rois), {R, 5}, opt);
  at::Tensor tout = ROIAlign_forward_cpu(tin, troi, scale, PH, PW, ratio);
  std::memcpy(out, tout.data_ptr<float>(), sizeof(float) * (size_t)tout.numel());
}

// boxes [N,4], scores [N]; keep receives ascending original indices; returns the count.
__attribute__((visibility("default"))) int64_t ref_nms_cpu(const float* boxes, const float* scores, int64_t n,
                                                            float thr, int64_t* keep) {
  auto opt = at::TensorOptions().dtype(at::kFloat);
  at::Tensor tb = at::from_blob(const_cast<float*>(boxes), {n, 4}, opt);
  at::Tensor ts = at::from_blob(const_cast<float*>(scores), {n}, opt);
  at::Tensor k = nms_cpu(tb, ts, thr).contiguous();
  std::memcpy(keep, k.data_ptr<int64_t>(), sizeof(int64_t) * (size_t)k.numel());
  return k.numel();
}

__attribute__((visibility("default"))) int ref_shim_version() { return 1; }
}
