// roi_v2_emu.cpp -- host emulation of the v2 ROIAlign kernels: compiles abr_iod_b200/csrc/roi_v2.cuh (the device logic
// itself) with g++ and runs it lane by lane, exactly in the task decomposition of roi_v2.cu's kernels.  TEST
// INFRASTRUCTURE: tests/test_v2_emulation.py drives it with numpy inputs and compares with the CPU oracle.
//   g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -o tools/emu/_build/libroi_v2_emu.so tools/emu/roi_v2_emu.cpp
#include "emu_shim.h"

#include "../../abr_iod_b200/csrc/roi_v2.cuh"

using namespace abr;

static LevelTable one_level(const void* p, int H, int W, float scale) {
  LevelTable lv;
  memset(&lv, 0, sizeof(lv));
  lv.ptr[0] = const_cast<void*>(p);
  lv.H[0] = H;
  lv.W[0] = W;
  lv.scale[0] = scale;
  return lv;
}

extern "C" {

__attribute__((visibility("default"))) int emu_plan_words(int PH, int PW) { return (int)v2_plan_words(PH, PW); }

// plan_kernel of roi_v2.cu: `nth` emulated threads per RoI, the three phases in order
__attribute__((visibility("default"))) void emu_plan(const float* rois, int R, int H, int W, float scale, int PH, int PW, int ratio,
                                                     int* plans, int nth) {
  const LevelTable lv = one_level(nullptr, H, W, scale);
  const size_t stride = v2_plan_words(PH, PW);
  for (int r = 0; r < R; r++) {
    int* plan = plans + (size_t)r * stride;
    const RoiGeom g = roi_geometry(rois, nullptr, lv, r, PH, PW, ratio);
    for (int t = 0; t < nth; t++) v2_plan_axes(plan, g, H, W, PH, PW, t, nth);
    v2_plan_header(plan, g, H, W, PH, PW);
    for (int t = 0; t < nth; t++) v2_plan_transposed(plan, plan, PH, PW, t, nth);
  }
}

}  // extern "C"

template <int V, int NT>
static void fwd_t(const int* plans, const float* rois, const float* m0, const float* m1, float* o0, float* o1, float* sums, int R,
                  int C, int H, int W, int PH, int PW, float scale, int ratio) {
  const LevelTable lv = one_level(nullptr, H, W, scale);
  const size_t stride = v2_plan_words(PH, PW);
  const int nslices = (C + 32 * V - 1) / (32 * V);
  const float* maps[NT];
  float* outs[NT];
  maps[0] = m0; outs[0] = o0;
  if (NT == 2) { maps[NT - 1] = m1; outs[NT - 1] = o1; }
  char* strip = new char[v2_strip_bytes(V, NT)];  // one warp's strip (lanes run one after the other, each in its own columns)
  char* plan_s = new char[v2_plan_smem_bytes(PW + PH)];
  for (int r = 0; r < R; r++)
    for (int slice = 0; slice < nslices; slice++)
      for (int pw = 0; pw < PW; pw++)
        for (int lane = 0; lane < 32; lane++) {
          const int* plan = plans + (size_t)r * stride;
          int c = (slice * 32 + lane) * V;
          const bool active = c < C;
          if (!active) c = 0;
          float* srs = NT == 2 ? sums + ((size_t)r * nslices + slice) * PH * PW * 3 : nullptr;
          for (int t = 0; t < 7; t++) v2_stage_plan(plan, plan_s, PW + PH, t, 7);  // the CTA's copy, then __syncthreads()
          if (plan[0] == V2_GENERIC || (plan[0] == V2_PLAN && plan[10] > v2_strip_rows_for(NT))) {
            const RoiGeom g = roi_geometry(rois, nullptr, lv, r, PH, PW, ratio);
            v2_generic_fwd_column<float, V, NT>(g, H, W, maps, outs, srs, r, pw, c, active, C, PH, PW, lane);
          } else {
            v2_fwd_column<float, V, NT>(plan_s, maps, outs, srs, strip, nullptr, r, pw, c, active, C, PH, PW, lane);
          }
        }
  delete[] strip;
  delete[] plan_s;
}

extern "C" __attribute__((visibility("default"))) void emu_fwd(const int* plans, const float* rois, const float* m0, const float* m1, float* o0,
                                                    float* o1, float* sums, int R, int C, int H, int W, int PH, int PW, float scale,
                                                    int ratio, int V) {
  if (m1) {
    const int nslices = (C + 32 * V - 1) / (32 * V);
    memset(sums, 0, sizeof(float) * (size_t)R * nslices * PH * PW * 3);
    if (V == 4) fwd_t<4, 2>(plans, rois, m0, m1, o0, o1, sums, R, C, H, W, PH, PW, scale, ratio);
    else fwd_t<1, 2>(plans, rois, m0, m1, o0, o1, sums, R, C, H, W, PH, PW, scale, ratio);
  } else {
    if (V == 4) fwd_t<4, 1>(plans, rois, m0, nullptr, o0, nullptr, nullptr, R, C, H, W, PH, PW, scale, ratio);
    else fwd_t<1, 1>(plans, rois, m0, nullptr, o0, nullptr, nullptr, R, C, H, W, PH, PW, scale, ratio);
  }
}

template <int V, bool FUSED>
static void bwd_t(const int* plans, const float* rois, float* gmap, const float* a, const float* b, const float* coef, int R, int C,
                  int H, int W, int PH, int PW, float scale, int ratio) {
  const LevelTable lv = one_level(nullptr, H, W, scale);
  const size_t stride = v2_plan_words(PH, PW);
  const int nslices = (C + 32 * V - 1) / (32 * V);
  const int nwarps = 8, nbin = PH * PW;
  char* tile = new char[(size_t)nbin * 32 * V * 4];
  char* plan_s = new char[v2_plan_smem_bytes(PW + PH + kV2MaxFW + kV2MaxFH)];
  for (int r = 0; r < R; r++)
    for (int slice = 0; slice < nslices; slice++) {
      const int* plan = plans + (size_t)r * stride;
      const int mode = plan[0];
      if (mode == V2_EMPTY) continue;
      if (mode == V2_PLAN)
        for (int t = 0; t < 5; t++) {
          v2_stage_plan(plan, plan_s, PW + PH + plan[7], t, 5);
          v2_stage_records(plan, plan_s, PW + PH + kV2MaxFW, plan[9] - plan[8] + 1, t, 5);
        }
      for (int phase = 0; phase < 2; phase++)  // __syncthreads() between filling the tile and walking the pixel columns
        for (int warp = 0; warp < nwarps; warp++)
          for (int lane = 0; lane < 32; lane++) {
            int c = (slice * 32 + lane) * V;
            const bool active = c < C;
            if (!active) c = 0;
            V2Grad<float, V, FUSED> src;
            src.a = a + (size_t)r * nbin * C + c;
            src.b = FUSED ? b + (size_t)r * nbin * C + c : nullptr;
            src.coef = FUSED ? reinterpret_cast<const float2*>(coef) + (size_t)r * nbin : nullptr;
            if (phase == 0) {
              v2_bwd_fill_tile<float, V, FUSED>(tile, src, nbin, C, warp, nwarps, lane);
              continue;
            }
            if (!active) continue;
            if (mode == V2_GENERIC) {
              const RoiGeom g = roi_geometry(rois, nullptr, lv, r, PH, PW, ratio);
              for (int pw = warp; pw < PW; pw += nwarps) v2_generic_bwd_column<float, V>(g, H, W, gmap, tile, pw, c, C, PH, PW, lane);
            } else {
              const int FW = plan[7];
              const int FH = plan[9] - plan[8] + 1, S = v2_bwd_row_chunks(FW, nwarps), chunk = (FH + S - 1) / S;
              for (int task = warp; task < FW * S; task += nwarps) {
                const int k = task / S, j0 = (task - k * S) * chunk, j1 = j0 + chunk < FH ? j0 + chunk : FH;
                if (j0 < j1) v2_bwd_pixcol<float, V>(plan_s, gmap, tile, k, j0, j1, c, C, PH, PW, lane);
              }
            }
          }
    }
  delete[] tile;
  delete[] plan_s;
}

extern "C" __attribute__((visibility("default"))) void emu_bwd(const int* plans, const float* rois, float* gmap, const float* a, const float* b,
                                                    const float* coef, int R, int C, int H, int W, int PH, int PW, float scale,
                                                    int ratio, int V, int fused) {
  if (fused) {
    if (V == 4) bwd_t<4, true>(plans, rois, gmap, a, b, coef, R, C, H, W, PH, PW, scale, ratio);
    else bwd_t<1, true>(plans, rois, gmap, a, b, coef, R, C, H, W, PH, PW, scale, ratio);
  } else {
    if (V == 4) bwd_t<4, false>(plans, rois, gmap, a, b, coef, R, C, H, W, PH, PW, scale, ratio);
    else bwd_t<1, false>(plans, rois, gmap, a, b, coef, R, C, H, W, PH, PW, scale, ratio);
  }
}

