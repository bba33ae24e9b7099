// roi_v2.cu -- kernels and launchers of the gather-form ROIAlign (device logic and design notes: roi_v2.cuh), and the
// fused Attentive-RoI-Distillation step built on them (abr_roi_ard_fused):
//     v2_plan_kernel      one warp per RoI -> per-RoI records
//     v2_fwd_kernel<NT>   CTA = (RoI, slice of 32*V channels), one warp per bin column (two for outputs wider than 8), each
//                         with a private strip of shared memory; NT = 2 pools the teacher and the student map in one
//                         pass and emits the ARD channel sums of every position
//     ard_coeff_kernel    (ard.cu) softmaxes + per-position gradient coefficients + the loss, one small CTA per RoI
//     v2_bwd_kernel<F>    CTA = (RoI, slice): the pooled-gradient tile goes to shared memory (F = fused: formed on the fly
//                         from the two pooled tensors and the coefficients), then warps walk the footprint's pixel columns
// Reference semantics: csrc/cuda/ROIAlign_cuda.cu:64-346, distillation/distillation.py:86-130,
// tools/train_incremental.py:84-115 (teacher pooling, student pooling, ARD loss, backward into the student's map).
#include <cstdlib>

#include "roi_v2.cuh"
#include "roi_v2.h"

namespace abr {

// Header + axis records of one RoI (PH, PW <= 16): the planning warp's working copy in shared memory
constexpr int kV2PlanStage = kV2Hdr + 32 * kV2Rec;

__global__ void __launch_bounds__(128) v2_plan_kernel(LevelTable lv, const float* __restrict__ rois,
                                                     const int32_t* __restrict__ levels, int* __restrict__ plans, size_t stride,
                                                     int R, int PH, int PW, int ratio) {
  // The header and the transposed records are derived from the axis records: those are built in shared memory (the
  // dependent reads of the later phases cost ~30 cycles there instead of an L2 round trip each) and copied out at the end.
  __shared__ __align__(16) int stage[4][kV2PlanStage];
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= R) return;
  const RoiGeom g = roi_geometry(rois, levels, lv, r, PH, PW, ratio);
  const int H = lv.H[g.level], W = lv.W[g.level];
  int* plan = plans + (size_t)r * stride;
  int* mine = stage[threadIdx.x >> 5];
  v2_plan_axes(mine, g, H, W, PH, PW, lane, 32);
  __syncwarp();
  if (lane == 0) v2_plan_header(mine, g, H, W, PH, PW);
  __syncwarp();
  v2_plan_transposed(mine, plan, PH, PW, lane, 32);
  __syncwarp();
  const int4* src = reinterpret_cast<const int4*>(mine);
  int4* dst = reinterpret_cast<int4*>(plan);
  for (int i = lane; i < (kV2Hdr + (PW + PH) * kV2Rec) / 4; i += 32) dst[i] = src[i];
}

template <typename T, int NT>
struct V2FwdArgs {
  LevelTable lv[NT];
  const int* plans;
  size_t stride;
  const float* rois;
  const int32_t* levels;
  T* out[NT];
  float* sums;  // [R][nslices][PH*PW][3] (NT == 2)
  unsigned int* clear_word;  // set to 0 by the first thread (the completion counter of the coefficient kernel that follows), or null
  int C, PH, PW, ratio, nslices, plan_smem;
};

static size_t align128(size_t x) { return (x + 127) & ~(size_t)127; }

template <typename T, int V, int NT>
__global__ void __launch_bounds__(256, NT == 2 ? 2 : 3) v2_fwd_kernel(const __grid_constant__ V2FwdArgs<T, NT> a) {
  extern __shared__ __align__(128) unsigned char v2_smem[];
  const int r = blockIdx.x / a.nslices, slice = blockIdx.x - r * a.nslices;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int* plan = a.plans + (size_t)r * a.stride;
  // the plan is staged before its header is looked at (one L2 round trip instead of two; a GENERIC RoI does not use it)
  const v2_sptr plan_s = v2_sptr_of(v2_smem);
  v2_stage_plan(plan, plan_s, a.PW + a.PH, threadIdx.x, blockDim.x);
  __syncthreads();
  const int4 h0 = v2_lds4i(plan_s);
  const int mode = h0.x, level = h0.z;
  if (a.clear_word && blockIdx.x == 0 && threadIdx.x == 0) *a.clear_word = 0u;
  int c = (slice * 32 + lane) * V;
  const bool active = c < a.C;
  if (!active) c = 0;  // idle lanes of a ragged last slice shadow channel 0 and never store
  const T* maps[NT];
  T* outs[NT];
#pragma unroll
  for (int t = 0; t < NT; t++) {
    maps[t] = static_cast<const T*>(a.lv[t].ptr[level]);
    outs[t] = a.out[t];
  }
  float* srs = NT == 2 ? a.sums + ((size_t)r * a.nslices + slice) * a.PH * a.PW * 3 : nullptr;
  bool generic = mode == V2_GENERIC;  // (the whole CTA)
  // a bin taller than the strip (two-tensor kernel: 13..15 map rows, i.e. a RoI more than ~12 * PH map rows high)
  // sends the RoI down the per-sample path
  if (!generic && v2_strip_rows_for(NT) <= kV2Sup) generic = v2_lds4i(plan_s + 32).z > v2_strip_rows_for(NT);  // hdr[10]
  if (generic) {
    const RoiGeom g = roi_geometry(a.rois, a.levels, a.lv[0], r, a.PH, a.PW, a.ratio);
    for (int pw = warp; pw < a.PW; pw += nw)
      v2_generic_fwd_column<T, V, NT>(g, a.lv[0].H[g.level], a.lv[0].W[g.level], maps, outs, srs, r, pw, c, active, a.C, a.PH,
                                      a.PW, lane);
    return;
  }
  const v2_sptr strip = plan_s + (uint32_t)a.plan_smem + warp * (uint32_t)(v2_strip_bytes(V, NT) + v2_sums_bytes(NT));
  const v2_sptr sums_buf = strip + (uint32_t)v2_strip_bytes(V, NT);
  for (int pw = warp; pw < a.PW; pw += nw)
    v2_fwd_column<T, V, NT>(plan_s, maps, outs, srs, strip, sums_buf, r, pw, c, active, a.C, a.PH, a.PW, lane);
}

template <typename T>
struct V2BwdArgs {
  LevelTable lv;  // gradient maps
  const int* plans;
  size_t stride;
  const float* rois;
  const int32_t* levels;
  const T* a;          // upstream gradient [R][PH][PW][C], or the teacher's pooled tensor (fused)
  const T* b;          // the student's pooled tensor (fused)
  const float2* coef;  // [R][PH*PW] (fused)
  int C, PH, PW, ratio, nslices, plan_smem;
};

template <typename T, int V, bool FUSED, int NTHR>
__global__ void __launch_bounds__(NTHR, 1024 / NTHR) v2_bwd_kernel(const __grid_constant__ V2BwdArgs<T> a) {
  extern __shared__ __align__(128) unsigned char v2_smem[];
  const int r = blockIdx.x / a.nslices, slice = blockIdx.x - r * a.nslices;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int nw = NTHR / 32;
  const int* plan = a.plans + (size_t)r * a.stride;
  const int nbin = a.PH * a.PW;
  int c = (slice * 32 + lane) * V;
  const bool active = c < a.C;
  if (!active) c = 0;
  const v2_sptr plan_s = v2_sptr_of(v2_smem), tile = plan_s + (uint32_t)a.plan_smem;
  V2Grad<T, V, FUSED> src;
  src.a = a.a + (size_t)r * nbin * a.C + c;
  src.b = FUSED ? a.b + (size_t)r * nbin * a.C + c : nullptr;
  src.coef = FUSED ? a.coef + (size_t)r * nbin : nullptr;
  // the plain form puts the tile's traffic in flight first: it does not depend on the plan, whose header and records cost
  // two L2 round trips (3 % on the kernel); for the fused form, with twice the requests per warp, the old order measured
  // 0.3 % better
  if (!FUSED) v2_bwd_fill_issue<T, V, FUSED>(tile, src, nbin, a.C, warp, nw, lane);
  const int4 h0 = __ldg(reinterpret_cast<const int4*>(plan)), h1 = __ldg(reinterpret_cast<const int4*>(plan + 4)),
             h2 = __ldg(reinterpret_cast<const int4*>(plan + 8));
  const int mode = h0.x, level = h0.z;
  const int FW = mode == V2_PLAN ? h1.w : 0, FH = mode == V2_PLAN ? h2.y - h2.x + 1 : 0;
  const int na = 4 * (1 + a.PW + a.PH + FW), nb = 4 * FH;  // 16-byte pieces of the plan's two staged ranges
  if (FUSED && mode == V2_PLAN && na <= NTHR && nb <= NTHR) {
    // one piece of each range per thread: the loads are issued, then the tile's traffic, then the pieces are parked -- the
    // plan's round trip and the tile's overlap
    const int tid = threadIdx.x;
    int4 pa = make_int4(0, 0, 0, 0), pb = pa;
    const int* rows = plan + kV2Hdr + (a.PW + a.PH + kV2MaxFW) * kV2Rec;
    if (tid < na) pa = __ldg(reinterpret_cast<const int4*>(plan) + tid);
    if (tid < nb) pb = __ldg(reinterpret_cast<const int4*>(rows) + tid);
    v2_bwd_fill_issue<T, V, FUSED>(tile, src, nbin, a.C, warp, nw, lane);
    if (tid < na) v2_sts4i(plan_s + 16 * tid, pa);
    if (tid < nb) v2_sts4i(plan_s + 64 * (1 + a.PW + a.PH + kV2MaxFW) + 16 * tid, pb);
  } else {
    if (mode == V2_PLAN) {
      v2_stage_plan(plan, plan_s, a.PW + a.PH + FW, threadIdx.x, NTHR);
      v2_stage_records(plan, plan_s, a.PW + a.PH + kV2MaxFW, FH, threadIdx.x, NTHR);
    }
    if (FUSED) v2_bwd_fill_issue<T, V, FUSED>(tile, src, nbin, a.C, warp, nw, lane);
  }
  v2_bwd_fill_combine<T, V, FUSED>(tile, src, nbin, a.C, warp, nw, lane);  // (also drains the asynchronous copies)
  if (mode == V2_EMPTY) return;  // (the whole CTA) no sample of the RoI falls inside the map
  __syncthreads();
  if (!active) return;  // idle lane of a ragged last slice (no warp-level primitive below)
  T* gmap = static_cast<T*>(a.lv.ptr[level]);
  if (mode == V2_GENERIC) {
    const RoiGeom g = roi_geometry(a.rois, a.levels, a.lv, r, a.PH, a.PW, a.ratio);
    for (int pw = warp; pw < a.PW; pw += nw)
      v2_generic_bwd_column<T, V>(g, a.lv.H[g.level], a.lv.W[g.level], gmap, tile, pw, c, a.C, a.PH, a.PW, lane);
    return;
  }
  if (NTHR >= 512) {
    const int S = v2_bwd_row_chunks(FW, nw), chunk = (FH + S - 1) / S;
    for (int task = warp; task < FW * S; task += nw) {
      const int k = task / S, j0 = (task - k * S) * chunk, j1 = j0 + chunk < FH ? j0 + chunk : FH;
      if (j0 < j1) v2_bwd_pixcol<T, V>(plan_s, gmap, tile, k, j0, j1, c, a.C, a.PH, a.PW, lane);
    }
  } else {
    for (int k = warp; k < FW; k += nw) v2_bwd_pixcol<T, V>(plan_s, gmap, tile, k, 0, FH, c, a.C, a.PH, a.PW, lane);
  }
}

// ------------------------------------------------------------------------------------------------ host side
bool v2_supported(int PH, int PW) { return PH >= 1 && PW >= 1 && PH <= 16 && PW <= 16; }

size_t v2_workspace_bytes(int R, int PH, int PW) { return (size_t)R * v2_plan_words(PH, PW) * sizeof(int); }

int v2_plan(const LevelTable& lv, const float* rois, const int32_t* levels, int* plans, int R, int PH, int PW, int ratio,
            cudaStream_t st) {
  v2_plan_kernel<<<ceil_div(R, 4), 128, 0, st>>>(lv, rois, levels, plans, v2_plan_words(PH, PW), R, PH, PW, ratio);
  ABR_CHECK_LAUNCH("roi_align_plan (v2)");
  return ABR_OK;
}

static int fwd_warps(int PW) { return PW <= 8 ? PW : ceil_div(PW, ceil_div(PW, 8)); }

template <typename T, int V, int NT>
static int launch_fwd2(const LevelTable* lv, const int* plans, const float* rois, const int32_t* levels, void* const* outs,
                       float* sums, int C, int R, int PH, int PW, int ratio, cudaStream_t st, unsigned int* clear_word = nullptr) {
  V2FwdArgs<T, NT> a;
  for (int t = 0; t < NT; t++) { a.lv[t] = lv[t]; a.out[t] = static_cast<T*>(outs[t]); }
  a.plans = plans; a.stride = v2_plan_words(PH, PW); a.rois = rois; a.levels = levels; a.sums = sums; a.clear_word = clear_word;
  a.C = C; a.PH = PH; a.PW = PW; a.ratio = ratio; a.nslices = ceil_div(C, 32 * V);
  const long long blocks = (long long)R * a.nslices;
  ABR_REQUIRE(blocks <= 0x7fffffffLL, ABR_ERR_UNSUPPORTED, "roi_align_forward: too many (RoI, slice) tasks");
  const int nw = fwd_warps(PW);
  a.plan_smem = (int)align128(v2_plan_smem_bytes(PW + PH));
  const size_t smem = a.plan_smem + (size_t)nw * (v2_strip_bytes(V, NT) + v2_sums_bytes(NT));
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    ABR_CUDA_OK(cudaFuncSetAttribute(v2_fwd_kernel<T, V, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(align128(v2_plan_smem_bytes(32)) + 8 * (v2_strip_bytes(V, NT) + v2_sums_bytes(NT)))));
    attr_set = true;
  }
  v2_fwd_kernel<T, V, NT><<<(unsigned)blocks, 32 * nw, smem, st>>>(a);
  ABR_CHECK_LAUNCH(NT == 2 ? "roi_align_forward (v2, teacher+student)" : "roi_align_forward (v2)");
  return ABR_OK;
}

int v2_forward(const LevelTable& lv, const int* plans, const float* rois, const int32_t* levels, void* out, int C, int R, int PH,
               int PW, int ratio, int dtype, cudaStream_t st) {
  void* outs[1] = {out};
  if (dtype == ABR_F32) {
    if (C % 4 == 0) return launch_fwd2<float, 4, 1>(&lv, plans, rois, levels, outs, nullptr, C, R, PH, PW, ratio, st);
    return launch_fwd2<float, 1, 1>(&lv, plans, rois, levels, outs, nullptr, C, R, PH, PW, ratio, st);
  }
  if (C % 8 == 0) return launch_fwd2<__nv_bfloat16, 8, 1>(&lv, plans, rois, levels, outs, nullptr, C, R, PH, PW, ratio, st);
  return launch_fwd2<__nv_bfloat16, 1, 1>(&lv, plans, rois, levels, outs, nullptr, C, R, PH, PW, ratio, st);
}

template <typename T, int V, bool FUSED>
static int launch_bwd2(const LevelTable& lv, const int* plans, const float* rois, const int32_t* levels, const void* a_, const void* b_,
                       const float2* coef, int C, int R, int PH, int PW, int ratio, cudaStream_t st) {
  V2BwdArgs<T> a;
  a.lv = lv; a.plans = plans; a.stride = v2_plan_words(PH, PW); a.rois = rois; a.levels = levels;
  a.a = static_cast<const T*>(a_); a.b = static_cast<const T*>(b_); a.coef = coef;
  a.C = C; a.PH = PH; a.PW = PW; a.ratio = ratio; a.nslices = ceil_div(C, 32 * V);
  const long long blocks = (long long)R * a.nslices;
  ABR_REQUIRE(blocks <= 0x7fffffffLL, ABR_ERR_UNSUPPORTED, "roi_align_backward: too many (RoI, slice) tasks");
  a.plan_smem = (int)align128(v2_plan_smem_bytes(PW + PH + kV2MaxFW + kV2MaxFH));
  const size_t smem = a.plan_smem + (size_t)PH * PW * 32 * V * sizeof(float);  // the plan + the (RoI, slice) gradient tile
  ABR_REQUIRE(smem <= 227 * 1024, ABR_ERR_UNSUPPORTED, "roi_align_backward: %dx%d bins of %d channels do not fit shared memory", PH, PW, 32 * V);
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    ABR_CUDA_OK(cudaFuncSetAttribute(v2_bwd_kernel<T, V, FUSED, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    ABR_CUDA_OK(cudaFuncSetAttribute(v2_bwd_kernel<T, V, FUSED, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  // large tiles leave room for two CTAs per SM at most: give those 16 warps each
  if (PH * PW >= 100) v2_bwd_kernel<T, V, FUSED, 512><<<(unsigned)blocks, 512, smem, st>>>(a);
  else v2_bwd_kernel<T, V, FUSED, 256><<<(unsigned)blocks, 256, smem, st>>>(a);
  ABR_CHECK_LAUNCH(FUSED ? "roi_align_backward (v2, fused ARD gradient)" : "roi_align_backward (v2)");
  return ABR_OK;
}

int v2_backward(const LevelTable& lv, const int* plans, const float* rois, const int32_t* levels, const void* gout, int C, int R,
                int PH, int PW, int ratio, int dtype, cudaStream_t st) {
  if (dtype == ABR_F32) {
    if (C % 4 == 0) return launch_bwd2<float, 4, false>(lv, plans, rois, levels, gout, nullptr, nullptr, C, R, PH, PW, ratio, st);
    return launch_bwd2<float, 1, false>(lv, plans, rois, levels, gout, nullptr, nullptr, C, R, PH, PW, ratio, st);
  }
  if (C % 8 == 0 && (size_t)PH * PW * 32 * 8 * sizeof(float) <= 220 * 1024)  // the tile of 8-channel lanes must fit shared memory
    return launch_bwd2<__nv_bfloat16, 8, false>(lv, plans, rois, levels, gout, nullptr, nullptr, C, R, PH, PW, ratio, st);
  return launch_bwd2<__nv_bfloat16, 1, false>(lv, plans, rois, levels, gout, nullptr, nullptr, C, R, PH, PW, ratio, st);
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct FusedLayout {
  size_t plans, sums, coef, ard, total;
};
static FusedLayout fused_layout(int R, int C, int PH, int PW) {
  const int V = C % 4 == 0 ? 4 : 1;
  const size_t nslices = ceil_div(C, 32 * V), HW = (size_t)PH * PW;
  FusedLayout f;
  f.plans = 0;
  f.sums = align256(v2_workspace_bytes(R, PH, PW));
  f.coef = f.sums + align256((size_t)R * nslices * HW * 3 * sizeof(float));
  f.ard = f.coef + align256((size_t)R * HW * sizeof(float2));
  f.total = f.ard + align256(ard_coeff_workspace_bytes(R));
  return f;
}

}  // namespace abr

using namespace abr;

extern "C" {

size_t abr_roi_ard_fused_workspace_bytes(int R, int C, int PH, int PW) {
  if (R <= 0 || C <= 0 || !v2_supported(PH, PW)) return 0;
  return fused_layout(R, C, PH, PW).total;
}

int abr_roi_ard_fused(const void* teacher_map, const void* student_map, const float* rois, void* pooled_old, void* pooled_new,
                      void* grad_student_map, float* loss3, int B, int C, int H, int W, int R, int PH, int PW,
                      float spatial_scale, int sampling_ratio, float gamma, float grad_scale, int dtype, int layout,
                      int zero_init, void* workspace, size_t workspace_bytes, int workspace_has_plan, abr_stream_t stream) {
  ABR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && R > 0, ABR_ERR_BAD_ARG,
              "roi_ard_fused: bad sizes B=%d C=%d H=%d W=%d R=%d (the reference's mean over zero RoIs is NaN)", B, C, H, W, R);
  ABR_REQUIRE(teacher_map && student_map && rois && pooled_old && pooled_new && loss3, ABR_ERR_BAD_ARG, "roi_ard_fused: null pointer");
  ABR_REQUIRE(dtype == ABR_F32 && layout == ABR_NHWC, ABR_ERR_UNSUPPORTED,
              "roi_ard_fused: fp32 channels-last only (dtype %d, layout %d); use the separate ops otherwise", dtype, layout);
  ABR_REQUIRE(v2_supported(PH, PW) && PH * PW <= 1024, ABR_ERR_UNSUPPORTED, "roi_ard_fused: output size %dx%d (max 16x16)", PH, PW);
  const FusedLayout f = fused_layout(R, C, PH, PW);
  ABR_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && workspace_bytes >= f.total, ABR_ERR_WORKSPACE,
              "roi_ard_fused: needs a 256-byte aligned workspace of %zu B (got %zu)", f.total, workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  int* plans = reinterpret_cast<int*>(ws + f.plans);
  float* sums = reinterpret_cast<float*>(ws + f.sums);
  float2* coef = reinterpret_cast<float2*>(ws + f.coef);
  LevelTable lv[2];
  for (int t = 0; t < 2; t++) {
    for (int l = 0; l < ABR_MAX_LEVELS; l++) { lv[t].ptr[l] = nullptr; lv[t].H[l] = lv[t].W[l] = 0; lv[t].scale[l] = 0.f; }
    lv[t].ptr[0] = const_cast<void*>(t == 0 ? teacher_map : student_map);
    lv[t].H[0] = H; lv[t].W[0] = W; lv[t].scale[0] = spatial_scale;
  }
  int rc;
  stage_mark(st, 0);
  if (!workspace_has_plan) {
    rc = v2_plan(lv[0], rois, nullptr, plans, R, PH, PW, sampling_ratio, st);
    if (rc) return rc;
  }
  stage_mark(st, 1);
  void* outs[2] = {pooled_old, pooled_new};
  // the pooling kernel also clears the completion counter of the coefficient kernel, and the coefficient kernel -- one
  // small latency-bound CTA per RoI -- also zero-fills the gradient map (when its size is a 16-byte multiple): no memset
  // node between the four kernels of the step
  unsigned int* counter = reinterpret_cast<unsigned int*>(ws + f.ard);
  if (C % 4 == 0) rc = launch_fwd2<float, 4, 2>(lv, plans, rois, nullptr, outs, sums, C, R, PH, PW, sampling_ratio, st, counter);
  else rc = launch_fwd2<float, 1, 2>(lv, plans, rois, nullptr, outs, sums, C, R, PH, PW, sampling_ratio, st, counter);
  if (rc) return rc;
  stage_mark(st, 2);
  const size_t gmap_bytes = (size_t)B * C * H * W * sizeof(float);
  const bool fill_in_kernel = grad_student_map && zero_init && R >= 256 /* enough CTAs to stream the fill */ && gmap_bytes % 16 == 0 &&
                              (reinterpret_cast<uintptr_t>(grad_student_map) & 15) == 0;
  rc = ard_coeff_run(sums, ceil_div(C, 32 * (C % 4 == 0 ? 4 : 1)), coef, loss3, R, C, PH * PW, gamma, grad_scale, ws + f.ard, st,
                     /*counter_is_clear=*/true, fill_in_kernel ? grad_student_map : nullptr, fill_in_kernel ? gmap_bytes : 0);
  if (rc) return rc;
  stage_mark(st, 3);
  if (grad_student_map) {
    if (zero_init && !fill_in_kernel) ABR_CUDA_OK(cudaMemsetAsync(grad_student_map, 0, gmap_bytes, st));
    LevelTable gl = lv[1];
    gl.ptr[0] = grad_student_map;
    if (C % 4 == 0) rc = launch_bwd2<float, 4, true>(gl, plans, rois, nullptr, pooled_old, pooled_new, coef, C, R, PH, PW, sampling_ratio, st);
    else rc = launch_bwd2<float, 1, true>(gl, plans, rois, nullptr, pooled_old, pooled_new, coef, C, R, PH, PW, sampling_ratio, st);
    if (rc) return rc;
  }
  stage_mark(st, 4);
  return ABR_OK;
}

}  // extern "C"
