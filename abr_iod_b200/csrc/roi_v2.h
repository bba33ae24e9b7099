// roi_v2.h -- internal (library-private) interface between roi_align.cu, roi_v2.cu and ard.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace abr {

struct LevelTable;

// Gather-form ROIAlign (roi_v2.cu): any output size up to 16x16, any number of levels.
bool v2_supported(int PH, int PW);
size_t v2_workspace_bytes(int R, int PH, int PW);
int v2_plan(const LevelTable& lv, const float* rois, const int32_t* levels, int* plans, int R, int PH, int PW, int ratio,
            cudaStream_t st);
int v2_forward(const LevelTable& lv, const int* plans, const float* rois, const int32_t* levels, void* out, int C, int R, int PH,
               int PW, int ratio, int dtype, cudaStream_t st);
int v2_backward(const LevelTable& lv, const int* plans, const float* rois, const int32_t* levels, const void* gout, int C, int R,
                int PH, int PW, int ratio, int dtype, cudaStream_t st);

// ARD from per-slice channel sums (ard.cu): sums [N][nslices][HW][3] -> coef [N][HW] (ka, kb) and loss3; `ws` holds
// ard_coeff_workspace_bytes(N) bytes.
size_t ard_coeff_workspace_bytes(int N);
// counter_is_clear: the caller has already zeroed the completion counter at the start of `ws` on this stream;
// zero_fill / zero_bytes: a buffer (16-byte aligned, a 16-byte multiple long) the kernel also fills with zeros, or null.
int ard_coeff_run(const float* sums, int nslices, float2* coef, float* loss3, int N, int C, int HW, float gamma, float grad_scale,
                  void* ws, cudaStream_t st, bool counter_is_clear = false, void* zero_fill = nullptr, size_t zero_bytes = 0);

}  // namespace abr
