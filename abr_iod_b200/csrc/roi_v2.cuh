// roi_v2.cuh -- gather-form ROIAlign ("v2") device logic, shared by the kernels in roi_v2.cu and by the host emulation
// in tools/emu/roi_v2_emu.cpp (which compiles this very file with g++ and checks it against the CPU oracle before any
// GPU time is spent; every function below is warp-uniform in its control flow, so running it lane by lane on the host
// is equivalent to the warp's execution).
//
// Semantics: maskrcnn_benchmark/csrc/cuda/ROIAlign_cuda.cu:64-254 of the reference.  Design (not a port):
//   * the bilinear weight of a sample is separable and so is the "outside the map => 0" rule, hence
//        out[ph][pw] = (1/count) * sum_y Wy[ph][y] * sum_x Wx[pw][x] * V[y][x]            (exactly)
//     with per-axis tables Wy / Wx that depend on the RoI only.  plan kernel: one 16-word record per bin row, per bin
//     column and per footprint pixel column:  { first index | count << 16, up to 15 weights }.
//   * FORWARD, a warp owns (RoI, bin column pw, 32*V channels).  It works in STRIPS of up to 16 map rows: phase 1
//     computes  T[y] = sum_x Wx[pw][x] * V[y][x]  for the strip's rows -- four rows at a time, all of their loads
//     (nx coalesced 512-byte requests per row) issued before the first is used, so a warp keeps up to 16 requests in
//     flight instead of one dependent gather after another -- into a lane-private strip of shared memory (every lane
//     reads back only what it wrote: no barrier); phase 2 emits every bin whose rows lie inside the strip,
//     out[ph] = sum_y Wy[ph][y] * T[y].  Sample rows grow monotonically with ph, so the next strip starts at the first
//     row of the first bin not yet emitted; each footprint pixel of the column is loaded once (plus a few rows where
//     strips overlap), for thin bins (several bins inside one map pixel, the P=14 case) as well as for fat ones.
//   * BACKWARD is the transposed gather.  A CTA owns (RoI, 32*V channels): phase 1 streams the RoI's pooled-gradient
//     tile [PH*PW][32*V] into shared memory with independent coalesced loads (FUSED: forms it on the fly, see below);
//     phase 2: a warp owns a footprint PIXEL column x; for every bin row it forms
//     G = sum_{pw covering x} Wx[pw][x] * g[ph][pw]  from the tile, adds Wy[ph][y] * G into a two-row cache of pixel
//     sums and, when a row leaves the cache, issues ONE vector reduction for that pixel: every footprint pixel of a RoI
//     is reduced exactly once per channel slice (the reference issues 4*g*g scalar atomicAdds per output element;
//     round 1 issued one reduction per pixel per covering column).
//   * NT = 2 forward pools the teacher and the student map with the same plan in one pass and emits the three
//     per-position channel sums the ARD loss needs (sum f_old^2, sum f_new^2, sum (f_new - f_old)^2); the FUSED
//     backward reads both pooled tensors, forms dL/df_new = ka*(f_new - f_old) + kb*f_new on the fly from per-position
//     coefficients and scatters it, so the ARD gradient tensor is never materialised.
#pragma once

#ifndef ABR_EMU
#include "common.cuh"
#define ABR_DEV __device__ __forceinline__
#define ABR_DEV_COLD __device__ __noinline__  // the rare per-sample paths: kept out of the hot paths' register budget
#define ABR_DEVM __device__ __forceinline__
#define ABR_HD __device__ __forceinline__
#define ABR_HOSTDEV __host__ __device__ __forceinline__
#define ABR_LDG4I(p) __ldg(reinterpret_cast<const int4*>(p))
#define ABR_LDGI(p) __ldg(p)
#define ABR_LDG2F(p) __ldg(p)
#endif

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
ABR_HD RoiGeom roi_geometry(const float* __restrict__ rois, const int32_t* __restrict__ levels, const LevelTable& lv, int r,
                            int PH, int PW, int ratio) {
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

// Sample coordinate of ROIAlign_cuda.cu:109,112 in the reference's operation order (no contraction).
ABR_HD float v2_sample_coord(float start, float bin, int p, int i, int grid) {
  return __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)), __fdiv_rn(__fmul_rn((float)i + .5f, bin), (float)grid));
}

// ------------------------------------------------------------------------------------------------ plan (v2)
constexpr int kV2Rec = 16;      // words per record: { lo | n << 16, w[0..14] }
constexpr int kV2Sup = 15;      // widest support (map pixels of one bin / bins over one pixel) a record holds
constexpr int kV2MaxFW = 64;    // widest footprint (map pixels) with pixel-column records
constexpr int kV2MaxFH = 64;    // tallest footprint (map rows) with row records
constexpr int kV2Hdr = 16;
enum V2Mode { V2_EMPTY = 0, V2_PLAN = 1, V2_GENERIC = 3 };
// hdr: [0] mode [1] batch [2] level [3] H [4] W [5] 1/count [6] X0 [7] FW [8] Y0 [9] Y1 [10] tallest bin (map rows)

// records after the header: PW bin columns, PH bin rows, kV2MaxFW footprint pixel columns, kV2MaxFH footprint rows
ABR_HOSTDEV size_t v2_plan_words(int PH, int PW) { return (size_t)kV2Hdr + (size_t)(PH + PW + kV2MaxFW + kV2MaxFH) * kV2Rec; }

// One bin of one axis (ROIAlign_cuda.cu:22-47 along one axis): the map indices its samples touch and the summed
// bilinear weights.  Returns the support size n (0: no sample inside the map) or -1 when it exceeds kV2Sup.
ABR_HD int v2_axis_record(int* rec, int p, int S, float start, float bin, int grid) {
  float w[kV2Sup];
#pragma unroll
  for (int i = 0; i < kV2Sup; i++) w[i] = 0.f;
  int lo = -1, hi = -1;
  const float fS = (float)S;
  for (int i = 0; i < grid; i++) {
    float c = v2_sample_coord(start, bin, p, i, grid);
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
    if (lo < 0) lo = low;
    if (low < lo || high - lo >= kV2Sup) return -1;
    w[low - lo] += h;
    w[high - lo] += l;
    if (high > hi) hi = high;
  }
  const int n = lo < 0 ? 0 : hi - lo + 1;
  rec[0] = (lo < 0 ? 0 : lo) | (n << 16);
  for (int i = 0; i < kV2Sup; i++) rec[1 + i] = __float_as_int(w[i]);
  return n;
}

// Phase 1 (threads tid, tid + nth, ...): the PW column records and the PH bin-row records.
ABR_HD void v2_plan_axes(int* plan, const RoiGeom& g, int H, int W, int PH, int PW, int tid, int nth) {
  for (int i = tid; i < PW + PH; i += nth) {
    int* rec = plan + kV2Hdr + i * kV2Rec;
    const int n = i < PW ? v2_axis_record(rec, i, W, g.start_w, g.bin_w, g.grid_w)
                         : v2_axis_record(rec, i - PW, H, g.start_h, g.bin_h, g.grid_h);
    if (n < 0) rec[0] = -1;  // support too wide: the RoI is left to the per-sample path
  }
}

// Phase 2 (one thread, after phase 1 is visible): header.
ABR_HD void v2_plan_header(int* plan, const RoiGeom& g, int H, int W, int PH, int PW) {
  int X0 = W, X1 = -1, Y0 = H, Y1 = -1, tallest = 0;
  bool generic = false;
  for (int i = 0; i < PW + PH; i++) {
    const int w0 = plan[kV2Hdr + i * kV2Rec];
    if (w0 < 0) { generic = true; continue; }
    const int lo = w0 & 0xffff, n = w0 >> 16;
    if (n == 0) continue;
    if (i < PW) { X0 = lo < X0 ? lo : X0; X1 = lo + n - 1 > X1 ? lo + n - 1 : X1; }
    else { Y0 = lo < Y0 ? lo : Y0; Y1 = lo + n - 1 > Y1 ? lo + n - 1 : Y1; tallest = n > tallest ? n : tallest; }
  }
  int mode = V2_PLAN;
  if (generic) mode = V2_GENERIC;
  else if (X1 < X0 || Y1 < Y0) mode = V2_EMPTY;
  else if (X1 - X0 + 1 > kV2MaxFW || Y1 - Y0 + 1 > kV2MaxFH) mode = V2_GENERIC;
  plan[0] = mode; plan[1] = g.batch; plan[2] = g.level; plan[3] = H; plan[4] = W;
  plan[5] = __float_as_int(1.f / g.count);
  plan[6] = X1 < X0 ? 0 : X0; plan[7] = X1 < X0 ? 0 : X1 - X0 + 1;
  plan[8] = Y1 < Y0 ? 0 : Y0; plan[9] = Y1 < Y0 ? 0 : Y1;
  plan[10] = tallest;  // the forward sends RoIs with a bin taller than its strip down the per-sample path
  for (int i = 11; i < kV2Hdr; i++) plan[i] = 0;
}

// Phase 3 (threads tid, tid + nth, ..., after phase 2 is visible): the transposed records the backward walks.
//   * one per footprint pixel column x = X0 + k -- the contiguous range of bin columns whose support contains x and their
//     weights Wx[pw][x] (supports start and end monotonically in pw, so the covering columns are contiguous);
//   * one per footprint row y = Y0 + j -- likewise the bin rows whose support contains y with a non-zero weight, and
//     Wy[ph][y].
// Record k of the pixel columns sits after the PW + PH axis records, row record j after kV2MaxFW pixel-column records.
ABR_HD void v2_plan_transposed(int* axes, int* plan, int PH, int PW, int tid, int nth) {
  // `axes`: where the header and the PW + PH axis records are read (the plan itself, or the planning warp's copy in
  // shared memory, which is written to the plan afterwards); the transposed records go to `plan`.
  if (axes[0] != V2_PLAN) return;
  const int X0 = axes[6], FW = axes[7], Y0 = axes[8], FH = axes[9] - axes[8] + 1;
  for (int k = tid; k < FW + FH; k += nth) {
    const bool is_row = k >= FW;
    const int at = is_row ? Y0 + (k - FW) : X0 + k;                    // map column / row
    const int first = is_row ? PW : 0, count = is_row ? PH : PW;       // the axis records to transpose
    int* rec = plan + kV2Hdr + (PW + PH + (is_row ? kV2MaxFW + (k - FW) : k)) * kV2Rec;
    int q0 = -1, nq = 0;
    int w[kV2Sup];
#pragma unroll
    for (int i = 0; i < kV2Sup; i++) w[i] = 0;
    for (int q = 0; q < count; q++) {
      const int* cr = axes + kV2Hdr + (first + q) * kV2Rec;
      const int lo = cr[0] & 0xffff, n = cr[0] >> 16;
      if (n == 0 || at < lo || at > lo + n - 1) continue;
      const int wq = cr[1 + (at - lo)];
      if (is_row && __int_as_float(wq) == 0.f) continue;  // e.g. the upper tap of a sample that sits exactly on a map row
      if (q0 < 0) q0 = q;
      if (q - q0 < kV2Sup) {
#pragma unroll
        for (int i = 0; i < kV2Sup; i++)
          if (i == q - q0) w[i] = wq;
      }
      nq = q - q0 + 1;
    }
    // more than kV2Sup bins over one map pixel (a one-pixel RoI pooled to 16 bins) do not fit a record: the RoI goes to
    // the per-sample path (several threads may store the same value)
    if (nq > kV2Sup) axes[0] = V2_GENERIC;
    rec[0] = (q0 < 0 ? 0 : q0) | ((nq > kV2Sup ? kV2Sup : nq) << 16);
#pragma unroll
    for (int i = 0; i < kV2Sup; i++) rec[1 + i] = w[i];
  }
}

// ------------------------------------------------------------------------------------------------ shared memory
// The kernels keep the RoI's plan, the forward's T strips and the backward's gradient tile in shared memory and address
// them with 32-bit shared-window addresses (one cvta per kernel; no generic-pointer arithmetic in the inner loops).  The
// host emulation maps the same helpers onto plain memory.
#ifndef ABR_EMU
typedef uint32_t v2_sptr;
ABR_DEV v2_sptr v2_sptr_of(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
ABR_DEV int4 v2_lds4i(v2_sptr a) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
ABR_DEV void v2_sts4i(v2_sptr a, int4 v) {
  asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
ABR_DEV float v2_ldsf(v2_sptr a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
ABR_DEV void v2_stsf(v2_sptr a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
#else
typedef char* v2_sptr;
ABR_DEV v2_sptr v2_sptr_of(const void* p) { return (char*)p; }
ABR_DEV int4 v2_lds4i(v2_sptr a) { int4 v; memcpy(&v, a, 16); return v; }
ABR_DEV void v2_sts4i(v2_sptr a, int4 v) { memcpy(a, &v, 16); }
ABR_DEV float v2_ldsf(v2_sptr a) { float v; memcpy(&v, a, 4); return v; }
ABR_DEV void v2_stsf(v2_sptr a, float v) { memcpy(a, &v, 4); }
#endif
// V consecutive floats of one lane (16-byte accesses when V % 4 == 0)
template <int V>
ABR_DEV void v2_sm_store(v2_sptr a, const float (&v)[V]) {
  if (V % 4 == 0) {
#pragma unroll
    for (int i = 0; i < V / 4; i++) {
      int4 t;
      t.x = __float_as_int(v[4 * i]); t.y = __float_as_int(v[4 * i + 1]); t.z = __float_as_int(v[4 * i + 2]); t.w = __float_as_int(v[4 * i + 3]);
      v2_sts4i(a + 16 * i, t);
    }
  } else {
#pragma unroll
    for (int i = 0; i < V; i++) v2_stsf(a + 4 * i, v[i]);
  }
}
template <int V>
ABR_DEV void v2_sm_load(v2_sptr a, float (&v)[V]) {
  if (V % 4 == 0) {
#pragma unroll
    for (int i = 0; i < V / 4; i++) {
      const int4 t = v2_lds4i(a + 16 * i);
      v[4 * i] = __int_as_float(t.x); v[4 * i + 1] = __int_as_float(t.y); v[4 * i + 2] = __int_as_float(t.z); v[4 * i + 3] = __int_as_float(t.w);
    }
  } else {
#pragma unroll
    for (int i = 0; i < V; i++) v[i] = v2_ldsf(a + 4 * i);
  }
}

ABR_DEV v2_sptr v2_srec(v2_sptr plan_s, int k) { return plan_s + 64 * (1 + k); }
ABR_DEV float v2_srec_w(v2_sptr rec, int i) { return v2_ldsf(rec + 4 + 4 * i); }

// acc += w * x over V channels: with an even V two channels go through one packed fp32 FMA (fma.rn.f32x2, sm_100: the
// same IEEE result per element, half the issue slots; worth 0.5-2 % here -- the kernels are latency-, not issue-bound).
template <int V>
ABR_DEV void v2_axpy(float w, const float (&x)[V], float (&acc)[V]) {
#if !defined(ABR_V2_NO_FFMA2) && !defined(ABR_EMU)
  if (V % 2 == 0) {
    const float2 ww = make_float2(w, w);
#pragma unroll
    for (int i = 0; i < V / 2; i++) {
      const float2 r = __ffma2_rn(ww, make_float2(x[2 * i], x[2 * i + 1]), make_float2(acc[2 * i], acc[2 * i + 1]));
      acc[2 * i] = r.x;
      acc[2 * i + 1] = r.y;
    }
    return;
  }
#endif
#pragma unroll
  for (int i = 0; i < V; i++) acc[i] = fmaf(w, x[i], acc[i]);
}

// The RoI's plan, copied into shared memory by the whole CTA (a __syncthreads() follows in the kernel): `nrec` records
// after the header.  Record k of the shared copy sits at plan_s + 64 * (1 + k).
ABR_DEV void v2_stage_plan(const int* __restrict__ plan, v2_sptr plan_s, int nrec, int tid, int nth) {
  for (int i = tid; i < 4 * (1 + nrec); i += nth) v2_sts4i(plan_s + 16 * i, ABR_LDG4I(plan + 4 * i));
}
// ... `nrec` records starting at record `first` (the backward's row records sit after a fixed-size gap)
ABR_DEV void v2_stage_records(const int* __restrict__ plan, v2_sptr plan_s, int first, int nrec, int tid, int nth) {
  for (int i = tid; i < 4 * nrec; i += nth) v2_sts4i(plan_s + 64 * (1 + first) + 16 * i, ABR_LDG4I(plan + kV2Hdr + first * kV2Rec + 4 * i));
}
ABR_HOSTDEV size_t v2_plan_smem_bytes(int nrec) { return (size_t)64 * (1 + nrec); }
// Three warp totals with six shuffles (reduce-scatter): on return lane 0 holds sum(a), lane 16 sum(b), lane 8 sum(c).
ABR_DEV float v2_reduce3(float a, float b, float c, int lane) {
#ifndef ABR_EMU
  const bool hi = (lane & 16) != 0;
  float keep = (hi ? b : a) + __shfl_xor_sync(0xffffffffu, hi ? a : b, 16);
  c += __shfl_xor_sync(0xffffffffu, c, 16);
  const bool h8 = (lane & 8) != 0;
  float v = (h8 ? c : keep) + __shfl_xor_sync(0xffffffffu, h8 ? keep : c, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
#else
  (void)a; (void)b; (void)c; (void)lane;
  return 0.f;
#endif
}

// ------------------------------------------------------------------------------------------------ forward
// ARD channel sums of the two-tensor forward.  Per output bin every lane has three partial sums over its V channels
// (sum f_old^2, sum f_new^2, sum (f_new - f_old)^2); the warp totals go to sums[(RoI, slice)][bin][3].  Reducing each bin
// with shuffles costs a dependent chain of five shuffle/add pairs per bin; instead the partials of up to kV2SumBins bins
// are parked in a warp-private shared-memory table ([value][lane], rows skewed by one word so that both the lane-major
// writes and the value-major reads are conflict-free) and folded by 3 * bins lanes at once, 32 sequential adds each.
constexpr int kV2SumBins = 8;
ABR_HOSTDEV size_t v2_sums_bytes(int NT) { return NT == 2 ? (size_t)(kV2SumBins * 3 * 33 * 4 + 127) / 128 * 128 : 0; }

struct V2Sums {
  float* rs;    // &sums[(r * nslices + slice) * PH * PW * 3]
  v2_sptr buf;  // the warp's table (device only)
  int pw, PW, first_ph, n;
};
ABR_DEV void v2_sums_flush(V2Sums& q, int lane) {
#ifndef ABR_EMU
  __syncwarp();
  if (lane < q.n * 3) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 32; k++) s += v2_ldsf(q.buf + (lane * 33 + k) * 4);
    const int b = lane / 3;
    q.rs[((size_t)(q.first_ph + b) * q.PW + q.pw) * 3 + (lane - 3 * b)] = s;
  }
  __syncwarp();
#else
  (void)lane;
#endif
  q.first_ph += q.n;
  q.n = 0;
}
ABR_DEV void v2_sums_add(V2Sums& q, float so, float sn, float sd, int lane) {
#ifndef ABR_EMU
  const v2_sptr at = q.buf + (q.n * 3 * 33 + lane) * 4;
  v2_stsf(at, so);
  v2_stsf(at + 33 * 4, sn);
  v2_stsf(at + 2 * 33 * 4, sd);
#else
  float* o = q.rs + ((size_t)(q.first_ph + q.n) * q.PW + q.pw) * 3;  // the harness zeroes the buffer and runs the lanes in turn
  o[0] += so; o[1] += sn; o[2] += sd;
  (void)lane;
#endif
  if (++q.n == kV2SumBins) v2_sums_flush(q, lane);
}

// 16-byte loads with an L2 policy (fp32, V = 4; other types take the plain form).  The backward reads the pooled tensors
// once (evict-first: 0.5 % on the fused step); an evict-last policy on the forward's map loads was 8 % SLOWER than plain
// ld.global.nc and is not used.  `pol`: createpolicy word of the calling thread.
#ifndef ABR_EMU
ABR_DEV uint64_t v2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
#else
ABR_DEV uint64_t v2_policy_evict_first() { return 0; }
#endif
template <typename T, int V>
ABR_DEV void v2_load_hint(const T* p, float (&v)[V], uint64_t pol) {
#ifndef ABR_EMU
  if (sizeof(T) == 4 && V == 4) {
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p), "l"(pol));
    return;
  }
#endif
  (void)pol;
  VecIO<T, V>::load(p, v);
}

// Pooled outputs are written once and not read again before ~2 GB of other traffic has passed: the stores carry the
// evict-first hint so that the (re-read) feature maps keep their L2 lines (two-tensor forward at configs[0]: 0.60 -> 0.565 ms).
template <typename T, int V>
ABR_DEV void v2_store_out(T* p, const float (&v)[V]) {
#ifndef ABR_EMU
  VecIO<T, V>::store_stream(p, v);
#else
  VecIO<T, V>::store(p, v);
#endif
}

// The lane's three ARD partial sums over its V channels of one bin (sum a^2, sum b^2, sum (b - a)^2).  With an even V the
// channels go two at a time through packed fp32 FMAs (even channels in one half, odd ones in the other; d = b - a as one
// fused a * -1 + b: the same rounding) -- 0.9 % of the fused step; the scalar form adds in the same order.
template <int V>
ABR_DEV void v2_ard_partials(const float (&a)[V], const float (&b)[V], float& so, float& sn, float& sd) {
#if !defined(ABR_V2_NO_FFMA2) && !defined(ABR_EMU)
  if (V % 2 == 0) {
    float2 o2 = make_float2(0.f, 0.f), n2 = o2, d2 = o2;
    const float2 m1 = make_float2(-1.f, -1.f);
#pragma unroll
    for (int i = 0; i < V / 2; i++) {
      const float2 av = make_float2(a[2 * i], a[2 * i + 1]), bv = make_float2(b[2 * i], b[2 * i + 1]);
      const float2 dv = __ffma2_rn(av, m1, bv);
      o2 = __ffma2_rn(av, av, o2);
      n2 = __ffma2_rn(bv, bv, n2);
      d2 = __ffma2_rn(dv, dv, d2);
    }
    so = o2.x + o2.y; sn = n2.x + n2.y; sd = d2.x + d2.y;
    return;
  }
#endif
  if (V % 2 == 0) {
    float o2[2] = {0.f, 0.f}, n2[2] = {0.f, 0.f}, d2[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < V; k++) {
      const float d = b[k] - a[k];
      o2[k & 1] = fmaf(a[k], a[k], o2[k & 1]);
      n2[k & 1] = fmaf(b[k], b[k], n2[k & 1]);
      d2[k & 1] = fmaf(d, d, d2[k & 1]);
    }
    so = o2[0] + o2[1]; sn = n2[0] + n2[1]; sd = d2[0] + d2[1];
    return;
  }
  so = sn = sd = 0.f;
#pragma unroll
  for (int k = 0; k < V; k++) {
    const float d = b[k] - a[k];
    so = fmaf(a[k], a[k], so);
    sn = fmaf(b[k], b[k], sn);
    sd = fmaf(d, d, sd);
  }
}

// Emits one output bin of NT tensors: scale by 1/count, store, and (NT == 2) the lane's ARD partial sums.
template <typename T, int V, int NT>
ABR_DEV void v2_emit_bin(float (&acc)[NT][V], float inv_count, T* const (&o)[NT], bool active, V2Sums& sums, int lane) {
#pragma unroll
  for (int t = 0; t < NT; t++) {
#pragma unroll
    for (int k = 0; k < V; k++) acc[t][k] *= inv_count;
    if (active) v2_store_out<T, V>(o[t], acc[t]);
  }
  if (NT == 2) {
    float so, sn, sd;
    v2_ard_partials<V>(acc[0], acc[NT - 1], so, sn, sd);
    v2_sums_add(sums, so, sn, sd, lane);  // (an idle lane of a ragged slice has inv_count = 0: zero sums)
  }
}
// ... the same with the warp totals formed by shuffles and stored at once (the per-sample path)
template <typename T, int V, int NT>
ABR_DEV void v2_emit_bin_now(float (&acc)[NT][V], float inv_count, T* const (&o)[NT], bool active, float* sums_bin, int lane) {
#pragma unroll
  for (int t = 0; t < NT; t++) {
#pragma unroll
    for (int k = 0; k < V; k++) acc[t][k] *= inv_count;
    if (active) v2_store_out<T, V>(o[t], acc[t]);
  }
  if (NT == 2) {
    float so = 0.f, sn = 0.f, sd = 0.f;
    if (active) {
#pragma unroll
      for (int k = 0; k < V; k++) {
        const float a = acc[0][k], b = acc[NT - 1][k], d = b - a;
        so = fmaf(a, a, so);
        sn = fmaf(b, b, sn);
        sd = fmaf(d, d, sd);
      }
    }
#ifndef ABR_EMU
    const float v = v2_reduce3(so, sn, sd, lane);
    if (lane == 0) sums_bin[0] = v;
    if (lane == 16) sums_bin[1] = v;
    if (lane == 8) sums_bin[2] = v;
#else
    (void)lane;
    sums_bin[0] += so; sums_bin[1] += sn; sums_bin[2] += sd;  // the harness zeroes the buffer and runs the lanes in turn
#endif
  }
}

// Map rows per strip: 16 holds the tallest bin a record can describe (kV2Sup = 15 rows); the two-tensor kernel takes 12 so
// that two 7-warp CTAs fit an SM (8 rows and three CTAs measured the same: shorter strips overlap more and recompute more
// rows) and sends the rare RoIs with a taller bin down the per-sample path.
ABR_HOSTDEV constexpr int v2_strip_rows_for(int NT) { return NT == 2 ? 12 : 16; }

// Bytes of shared memory one forward warp's strips take: [NT][ROWS][32 lanes][V] floats.
ABR_HOSTDEV size_t v2_strip_bytes(int V, int NT) { return (size_t)NT * v2_strip_rows_for(NT) * 32 * V * sizeof(float); }

// Phase 1 of a strip for a column NX map pixels wide (compile-time): T[row] = sum_k w[k] * V[row][k] for `nrows` rows,
// RB rows -- RB * NX independent 16-byte loads per lane -- in flight at a time.  The last batch repeats the last row
// instead of predicating (the strip has room for all ROWS rows -- a multiple of RB --; rows past nrows are never read).
template <typename T, int V, int NX, int RB>
ABR_DEV void v2_strip_rows(const T* colbase, size_t rowstride, size_t pix, int nrows, v2_sptr colrec, v2_sptr dst) {
  float w[NX];
#pragma unroll
  for (int k = 0; k < NX; k++) w[k] = v2_srec_w(colrec, k);
  const T* p = colbase;  // row i0 + j, stepping one map row per load group and stopping at the last row
  for (int i0 = 0; i0 < nrows; i0 += RB) {
    float v[RB][NX][V];
#pragma unroll
    for (int j = 0; j < RB; j++) {
#pragma unroll
      for (int k = 0; k < NX; k++) VecIO<T, V>::load(p + (size_t)k * pix, v[j][k]);
      if (i0 + j + 1 < nrows) p += rowstride;
    }
#pragma unroll
    for (int j = 0; j < RB; j++) {
      float t[V];
#pragma unroll
      for (int q = 0; q < V; q++) t[q] = w[0] * v[j][0][q];
#pragma unroll
      for (int k = 1; k < NX; k++) v2_axpy<V>(w[k], v[j][k], t);
      v2_sm_store<V>(dst + (i0 + j) * (32 * V * 4), t);
    }
  }
}
// The same for NT tensors at once (same rows, same weights): the loads of all tensors are issued before the first is used,
// which halves the number of exposed gather latencies of the two-tensor kernel (0.66 -> 0.60 ms at configs[0]).
template <typename T, int V, int NT, int NX, int RB>
ABR_DEV void v2_strip_rows_nt(const T* const (&colbase)[NT], size_t rowstride, size_t pix, int nrows, v2_sptr colrec, v2_sptr dst0,
                              int dst_stride) {
  float w[NX];
#pragma unroll
  for (int k = 0; k < NX; k++) w[k] = v2_srec_w(colrec, k);
  size_t off = 0;  // row i0 + j, stepping one map row per load group and stopping at the last row
  for (int i0 = 0; i0 < nrows; i0 += RB) {
    float v[NT][RB][NX][V];
#pragma unroll
    for (int j = 0; j < RB; j++) {
#pragma unroll
      for (int t = 0; t < NT; t++)
#pragma unroll
        for (int k = 0; k < NX; k++) VecIO<T, V>::load(colbase[t] + off + (size_t)k * pix, v[t][j][k]);
      if (i0 + j + 1 < nrows) off += rowstride;
    }
#pragma unroll
    for (int t = 0; t < NT; t++)
#pragma unroll
      for (int j = 0; j < RB; j++) {
        float r[V];
#pragma unroll
        for (int q = 0; q < V; q++) r[q] = w[0] * v[t][j][0][q];
#pragma unroll
        for (int k = 1; k < NX; k++) v2_axpy<V>(w[k], v[t][j][k], r);
        v2_sm_store<V>(dst0 + t * dst_stride + (i0 + j) * (32 * V * 4), r);
      }
  }
}
// ... and for columns wider than four map pixels (fat bins of large RoIs): four rows at a time, pixel by pixel.
template <typename T, int V>
ABR_DEV void v2_strip_rows_wide(const T* colbase, size_t rowstride, size_t pix, int nrows, int nx, v2_sptr colrec, v2_sptr dst) {
  constexpr int RB = V >= 8 ? 2 : 4;
  for (int i0 = 0; i0 < nrows; i0 += RB) {
    float t[RB][V];
    const T* p[RB];
#pragma unroll
    for (int j = 0; j < RB; j++) {
      p[j] = colbase + (size_t)(i0 + j < nrows ? i0 + j : nrows - 1) * rowstride;
#pragma unroll
      for (int q = 0; q < V; q++) t[j][q] = 0.f;
    }
    for (int k = 0; k < nx; k++) {
      const float w = v2_srec_w(colrec, k);
      float v[RB][V];
#pragma unroll
      for (int j = 0; j < RB; j++) VecIO<T, V>::load(p[j] + (size_t)k * pix, v[j]);
#pragma unroll
      for (int j = 0; j < RB; j++)
#pragma unroll
        for (int q = 0; q < V; q++) t[j][q] = fmaf(w, v[j][q], t[j][q]);
    }
#pragma unroll
    for (int j = 0; j < RB; j++) v2_sm_store<V>(dst + (i0 + j) * (32 * V * 4), t[j]);
  }
}

// One bin column of one RoI.  plan_s: the plan in shared memory (header, PW column records, PH bin-row records);
// maps[t]: the level's map of tensor t ([B][H][W][C]); outs[t]: pooled tensor ([R][PH][PW][C]); c: first channel of this
// lane; sums_rs: &sums[(r * nslices + slice) * PH*PW * 3] (NT == 2); strip: this WARP's v2_strip_bytes(V, NT) of shared
// memory.
template <typename T, int V, int NT>
ABR_DEV void v2_fwd_column(v2_sptr plan_s, const T* const (&maps)[NT], T* const (&outs)[NT], float* sums_rs, v2_sptr strip,
                           v2_sptr sums_buf, int r, int pw, int c, bool active, int C, int PH, int PW, int lane) {
  constexpr int ROWB = 32 * V * 4;  // bytes of one strip row
  constexpr int kV2Rows = v2_strip_rows_for(NT);
  constexpr int RB2 = V >= 8 ? 1 : 2, RB4 = V >= 8 ? 2 : 4;
  const int4 h0 = v2_lds4i(plan_s), h1 = v2_lds4i(plan_s + 16), h2 = v2_lds4i(plan_s + 32);
  const int mode = h0.x, batch = h0.y, H = h0.w, W = h1.x, Y1 = h2.y;
  const float inv_count = active ? __int_as_float(h1.y) : 0.f;  // idle lanes of a ragged slice emit zeros (and park zero sums)
  const size_t pix = (size_t)C, binstride = (size_t)PW * C;
  T* o[NT];
#pragma unroll
  for (int t = 0; t < NT; t++) o[t] = outs[t] + ((size_t)r * PH * PW + pw) * C + c;
  V2Sums sums;
  sums.rs = sums_rs; sums.buf = sums_buf; sums.pw = pw; sums.PW = PW; sums.first_ph = 0; sums.n = 0;
  const v2_sptr colrec = v2_srec(plan_s, pw);
  const int col0 = v2_lds4i(colrec).x;
  const int nx = mode == V2_PLAN ? col0 >> 16 : 0;
  const size_t rowstride = (size_t)W * C;
  const T* base[NT];
#pragma unroll
  for (int t = 0; t < NT; t++) base[t] = maps[t] + ((size_t)batch * H * W + (col0 & 0xffff)) * C + c;
  const v2_sptr mine = strip + lane * (V * 4);  // [t][row] at (t * kV2Rows + row) * ROWB

  int ph = 0;
  while (ph < PH) {
    v2_sptr binrec = v2_srec(plan_s, PW + ph);
    int4 b = v2_lds4i(binrec);
    if (nx == 0 || (b.x >> 16) == 0) {  // no sample of this bin (column, RoI) falls inside the map
      float z[NT][V];
#pragma unroll
      for (int t = 0; t < NT; t++)
#pragma unroll
        for (int k = 0; k < V; k++) z[t][k] = 0.f;
      v2_emit_bin<T, V, NT>(z, 0.f, o, active, sums, lane);
#pragma unroll
      for (int t = 0; t < NT; t++) o[t] += binstride;
      ph++;
      continue;
    }
    const int ystart = b.x & 0xffff;
    // ---- phase 1: T of the rows ystart .. ystart + nrows - 1
    const int nrows = Y1 - ystart + 1 < kV2Rows ? Y1 - ystart + 1 : kV2Rows;
    if (NT == 2 && nx <= 4) {  // both tensors' loads of a row batch in flight together
      const T* cb[NT];
#pragma unroll
      for (int t = 0; t < NT; t++) cb[t] = base[t] + (size_t)ystart * rowstride;
      switch (nx) {
        case 1: v2_strip_rows_nt<T, V, NT, 1, RB4>(cb, rowstride, pix, nrows, colrec, mine, kV2Rows * ROWB); break;
        case 2: v2_strip_rows_nt<T, V, NT, 2, RB2>(cb, rowstride, pix, nrows, colrec, mine, kV2Rows * ROWB); break;
        case 3: v2_strip_rows_nt<T, V, NT, 3, RB2>(cb, rowstride, pix, nrows, colrec, mine, kV2Rows * ROWB); break;
        default: v2_strip_rows_nt<T, V, NT, 4, RB2 / 2 ? RB2 / 2 : 1>(cb, rowstride, pix, nrows, colrec, mine, kV2Rows * ROWB); break;
      }
    } else
#pragma unroll
    for (int t = 0; t < NT; t++) {
      const T* colbase = base[t] + (size_t)ystart * rowstride;
      const v2_sptr dst = mine + t * (kV2Rows * ROWB);
      switch (nx) {
        case 1: v2_strip_rows<T, V, 1, RB4>(colbase, rowstride, pix, nrows, colrec, dst); break;
        case 2: v2_strip_rows<T, V, 2, RB4>(colbase, rowstride, pix, nrows, colrec, dst); break;
        case 3: v2_strip_rows<T, V, 3, RB2>(colbase, rowstride, pix, nrows, colrec, dst); break;
        case 4: v2_strip_rows<T, V, 4, RB2>(colbase, rowstride, pix, nrows, colrec, dst); break;
        default: v2_strip_rows_wide<T, V>(colbase, rowstride, pix, nrows, nx, colrec, dst); break;
      }
    }
    // ---- phase 2: every bin whose rows lie inside the strip (at least the one that started it: n <= kV2Rows)
    while (true) {
      const int n = b.x >> 16;
      const v2_sptr src = mine + ((b.x & 0xffff) - ystart) * ROWB;
      float acc[NT][V];
#pragma unroll
      for (int t = 0; t < NT; t++) {
        float x[V];
        v2_sm_load<V>(src + t * (kV2Rows * ROWB), x);
        const float w0 = __int_as_float(b.y);
#pragma unroll
        for (int k = 0; k < V; k++) acc[t][k] = w0 * x[k];
      }
      if (n > 1) {
        const float w1 = __int_as_float(b.z);
#pragma unroll
        for (int t = 0; t < NT; t++) {
          float x[V];
          v2_sm_load<V>(src + t * (kV2Rows * ROWB) + ROWB, x);
          v2_axpy<V>(w1, x, acc[t]);
        }
      }
      if (n > 2) {
        const float w2 = __int_as_float(b.w);
#pragma unroll
        for (int t = 0; t < NT; t++) {
          float x[V];
          v2_sm_load<V>(src + t * (kV2Rows * ROWB) + 2 * ROWB, x);
          v2_axpy<V>(w2, x, acc[t]);
        }
      }
      for (int i = 3; i < n; i++) {
        const float wy = v2_srec_w(binrec, i);
#pragma unroll
        for (int t = 0; t < NT; t++) {
          float x[V];
          v2_sm_load<V>(src + t * (kV2Rows * ROWB) + i * ROWB, x);
          v2_axpy<V>(wy, x, acc[t]);
        }
      }
      v2_emit_bin<T, V, NT>(acc, inv_count, o, active, sums, lane);
#pragma unroll
      for (int t = 0; t < NT; t++) o[t] += binstride;
      if (++ph >= PH) break;
      binrec += 64;
      b = v2_lds4i(binrec);
      const int n2 = b.x >> 16;
      if (n2 == 0 || (b.x & 0xffff) + n2 > ystart + nrows) break;  // an empty bin, or one that needs a new strip: outer loop
    }
  }
  if (NT == 2 && sums.n > 0) v2_sums_flush(sums, lane);
}

// Per-sample evaluation of one bin column, the reference's loop nest (ROIAlign_cuda.cu:64-122): for the rare RoIs the
// plan marks GENERIC (a bin wider than kV2Sup map pixels, a footprint wider than kV2MaxFW).
template <typename T, int V, int NT>
ABR_DEV_COLD void v2_generic_fwd_column(const RoiGeom& g, int H, int W, const T* const (&maps)[NT], T* const (&outs)[NT],
                                   float* sums_rs, int r, int pw, int c, bool active, int C, int PH, int PW, int lane) {
  const size_t binstride = (size_t)PW * C;
  T* o[NT];
#pragma unroll
  for (int t = 0; t < NT; t++) o[t] = outs[t] + ((size_t)r * PH * PW + pw) * C + c;
  float* sb = NT == 2 ? sums_rs + (size_t)pw * 3 : nullptr;
  const float fH = (float)H, fW = (float)W;
  const float inv_count = 1.f / g.count;
  for (int ph = 0; ph < PH; ph++) {
    float acc[NT][V];
#pragma unroll
    for (int t = 0; t < NT; t++)
#pragma unroll
      for (int k = 0; k < V; k++) acc[t][k] = 0.f;
    for (int iy = 0; iy < g.grid_h; iy++) {
      float y = v2_sample_coord(g.start_h, g.bin_h, ph, iy, g.grid_h);
      if (y < -1.0f || y > fH) continue;
      if (y <= 0.f) y = 0.f;
      int yl = (int)y, yh;
      if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
      const float ly = y - (float)yl, hy = 1.f - ly;
      for (int ix = 0; ix < g.grid_w; ix++) {
        float x = v2_sample_coord(g.start_w, g.bin_w, pw, ix, g.grid_w);
        if (x < -1.0f || x > fW) continue;
        if (x <= 0.f) x = 0.f;
        int xl = (int)x, xh;
        if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
        const float lx = x - (float)xl, hx = 1.f - lx;
        const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
#pragma unroll
        for (int t = 0; t < NT; t++) {
          const T* img = maps[t] + (size_t)g.batch * H * W * C + c;
          float v1[V], v2[V], v3[V], v4[V];
          VecIO<T, V>::load(img + ((size_t)yl * W + xl) * C, v1);
          VecIO<T, V>::load(img + ((size_t)yl * W + xh) * C, v2);
          VecIO<T, V>::load(img + ((size_t)yh * W + xl) * C, v3);
          VecIO<T, V>::load(img + ((size_t)yh * W + xh) * C, v4);
#pragma unroll
          for (int k = 0; k < V; k++) acc[t][k] += w1 * v1[k] + w2 * v2[k] + w3 * v3[k] + w4 * v4[k];
        }
      }
    }
    v2_emit_bin_now<T, V, NT>(acc, inv_count, o, active, sb, lane);
#pragma unroll
    for (int t = 0; t < NT; t++) o[t] += binstride;
    if (NT == 2) sb += (size_t)PW * 3;
  }
}

// ------------------------------------------------------------------------------------------------ backward
// Source of the pooled gradient g[bin] (V channels of this lane): the upstream gradient tensor, or (FUSED) the ARD
// gradient ka*(f_new - f_old) + kb*f_new formed from the two pooled tensors and the per-position coefficients.
template <typename T, int V, bool FUSED>
struct V2Grad {
  const T* a;          // gout, or f_old (FUSED), at [r][0][0][c]
  const T* b;          // f_new (FUSED)
  const float2* coef;  // [PH*PW] of this RoI (FUSED)
  // V channels of bin `bin`; off = bin * C (elements)
  ABR_DEVM void load(int bin, size_t off, float (&g)[V]) const {
    if (FUSED) {
      const float2 kc = ABR_LDG2F(coef + bin);
      float fo[V], fn[V];
      VecIO<T, V>::load(a + off, fo);
      VecIO<T, V>::load(b + off, fn);
#pragma unroll
      for (int k = 0; k < V; k++) g[k] = fmaf(kc.x, fn[k] - fo[k], kc.y * fn[k]);
    } else {
      VecIO<T, V>::load(a + off, g);
    }
  }
};

// Phase 1 of the backward: warp `warp` of `nw` brings the bins warp, warp + nw, ... of this (RoI, slice) gradient tile into
// shared memory -- bin b of lane l at tile + (b * 32 + l) * V * 4 bytes.  Lanes of a ragged last slice shadow channel 0;
// they never read the tile.  (Generic form: U bins' loads in flight at a time through registers, the last batch repeating
// the last bin instead of predicating so that every load is issued before the first is used.)
#ifndef ABR_EMU
// 16-byte asynchronous copy with an L2 policy (createpolicy) attached: the pooled tensors are read once
ABR_DEV void v2_cp_async16_hint(v2_sptr dst, const void* src, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
ABR_DEV void v2_cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
#endif

// Phase 1 in two parts, so that the kernel can put the tile's traffic in flight BEFORE it reads the RoI's plan (two
// dependent L2 round trips) and combine afterwards.  v2_bwd_fill_issue: the asynchronous part (fp32, 16 bytes per lane:
// the first operand by cp.async straight into the tile, every line of the second requested from L2); nothing for the
// other types.  v2_bwd_fill_tile = issue + the rest.
template <typename T, int V, bool FUSED>
ABR_DEV void v2_bwd_fill_issue(v2_sptr tile, const V2Grad<T, V, FUSED>& src, int nbin, int C, int warp, int nw, int lane) {
#ifndef ABR_EMU
  if (sizeof(T) == 4 && V == 4) {
    constexpr int BINB = 32 * V * 4;
    const v2_sptr mine = tile + lane * (V * 4);
    const uint64_t pol1 = v2_policy_evict_first();
    for (int b = warp; b < nbin; b += nw) v2_cp_async16_hint(mine + b * BINB, src.a + (size_t)b * C, pol1);
    if (FUSED && (lane & 7) == 0)  // one request per 128-byte line
      for (int b = warp; b < nbin; b += nw) asm volatile("prefetch.global.L2 [%0];" ::"l"(src.b + (size_t)b * C));
  }
#else
  (void)tile; (void)src; (void)nbin; (void)C; (void)warp; (void)nw; (void)lane;
#endif
}

template <typename T, int V, bool FUSED>
ABR_DEV void v2_bwd_fill_combine(v2_sptr tile, const V2Grad<T, V, FUSED>& src, int nbin, int C, int warp, int nw, int lane) {
  constexpr int BINB = 32 * V * 4;
  const v2_sptr mine = tile + lane * (V * 4);
#ifndef ABR_EMU
  if (sizeof(T) == 4 && V == 4) {
    if (!FUSED) {
      v2_cp_async_wait_all();
      return;
    }
    // The second operand goes through registers one bin at a time.  (Batches of loads held in registers were slower -- 8
    // bins per batch: +5 % on the fused step at configs[0] -- the prefetches already put all of the tile's traffic in flight.)
    const uint64_t pol1 = v2_policy_evict_first();
    bool landed = false;
    for (int b = warp; b < nbin; b += nw) {
      float fn[V], fo[V], g[V];
      v2_load_hint<T, V>(src.b + (size_t)b * C, fn, pol1);
      const float2 kc = __ldg(src.coef + b);
      if (!landed) {  // the first operand's copies (all of this warp's bins) have to have landed
        v2_cp_async_wait_all();
        landed = true;
      }
      v2_sm_load<V>(mine + b * BINB, fo);
#pragma unroll
      for (int q = 0; q < V; q++) g[q] = fmaf(kc.x, fn[q] - fo[q], kc.y * fn[q]);
      v2_sm_store<V>(mine + b * BINB, g);
    }
    return;
  }
#endif
  constexpr int U = (FUSED || V >= 8) ? 4 : 8;  // bins in flight per warp
  const size_t step = (size_t)nw * C, last = (size_t)(nbin - 1) * C;
  size_t off = (size_t)warp * C;
  for (int b0 = warp; b0 < nbin; b0 += U * nw, off += U * step) {
    float g[U][V];
#pragma unroll
    for (int j = 0; j < U; j++) {
      const bool in = b0 + j * nw < nbin;
      src.load(in ? b0 + j * nw : nbin - 1, in ? off + j * step : last, g[j]);
    }
#pragma unroll
    for (int j = 0; j < U; j++)
      if (b0 + j * nw < nbin) v2_sm_store<V>(mine + (b0 + j * nw) * BINB, g[j]);
  }
}

template <typename T, int V, bool FUSED>
ABR_DEV void v2_bwd_fill_tile(v2_sptr tile, const V2Grad<T, V, FUSED>& src, int nbin, int C, int warp, int nw, int lane) {
  v2_bwd_fill_issue<T, V, FUSED>(tile, src, nbin, C, warp, nw, lane);
  v2_bwd_fill_combine<T, V, FUSED>(tile, src, nbin, C, warp, nw, lane);
}

// The rows of one footprint pixel column covered by NQ bin columns (compile-time; NQ = 0: nq at run time).  For every
// footprint row y:  dV[y][x] = (1/count) * sum_{ph over y} sum_{q < nq} Wy[ph][y] * Wx[q0 + q][x] * g[ph][q0 + q],
// straight from the tile (no state carried between rows), then ONE vector reduction into the map.
template <typename T, int V, int NQ>
ABR_DEV void v2_bwd_rows(v2_sptr rowrec, int FH, v2_sptr tcol, int pwb, v2_sptr pxrec, int nq, float inv_count, T* gin, size_t rowstride) {
  constexpr int BINB = 32 * V * 4;
  float wq[NQ > 0 ? NQ : 1];
#pragma unroll
  for (int q = 0; q < NQ; q++) wq[q] = v2_srec_w(pxrec, q);
  for (int j = 0; j < FH; j++, rowrec += 64, gin += rowstride) {
    const int r0 = v2_lds4i(rowrec).x;
    const int m = r0 >> 16;
    if (m == 0) continue;
    float S[V];
#pragma unroll
    for (int i = 0; i < V; i++) S[i] = 0.f;
    v2_sptr p = tcol + (r0 & 0xffff) * pwb;
    for (int i = 0; i < m; i++, p += pwb) {
      const float wy = v2_srec_w(rowrec, i);
      if (NQ > 0) {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
          float g[V];
          v2_sm_load<V>(p + q * BINB, g);
          v2_axpy<V>(wy * wq[q], g, S);
        }
      } else {
        for (int q = 0; q < nq; q++) {
          float g[V];
          v2_sm_load<V>(p + q * BINB, g);
          const float w = wy * v2_srec_w(pxrec, q);
#pragma unroll
          for (int t = 0; t < V; t++) S[t] = fmaf(w, g[t], S[t]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < V; q++) S[q] *= inv_count;
    VecIO<T, V>::red_add(gin, S);
  }
}

// How phase 2 is cut into warp tasks in the 16-warp kernel: a footprint of FW pixel columns gives FW tasks, fewer than the
// CTA has warps for most RoIs; every column is therefore split into up to four row chunks (the rows of a column are
// independent) so that every warp has a task.  Returns the chunks per column.  (The 8-warp kernel of the small outputs
// keeps whole columns: the split costs it 1 %.)
ABR_HD int v2_bwd_row_chunks(int FW, int nw) {
  int s = FW > 0 ? (nw + FW - 1) / FW : 1;
  return s < 1 ? 1 : (s > 4 ? 4 : s);
}

// Phase 2: rows j0 .. j1 - 1 of one footprint pixel column x = X0 + k of one RoI.  plan_s: the plan in shared memory (header, PW + PH axis
// records, pixel-column records, row records); gmap: the level's gradient map [B][H][W][C]; tile: the (RoI, slice)
// gradient tile in shared memory.
template <typename T, int V>
ABR_DEV void v2_bwd_pixcol(v2_sptr plan_s, T* gmap, v2_sptr tile, int k, int j0, int j1, int c, int C, int PH, int PW, int lane) {
  constexpr int BINB = 32 * V * 4;
  const int4 h0 = v2_lds4i(plan_s), h1 = v2_lds4i(plan_s + 16), h2 = v2_lds4i(plan_s + 32);
  const int batch = h0.y, H = h0.w, W = h1.x, X0 = h1.z, Y0 = h2.x + j0, FH = j1 - j0;  // footprint rows j0 .. j1 - 1 of the column
  const float inv_count = __int_as_float(h1.y);
  const v2_sptr pxrec = v2_srec(plan_s, PW + PH + k);
  const int px0 = v2_lds4i(pxrec).x;
  const int nq = px0 >> 16;
  if (nq == 0) return;  // a map column between two bin columns' supports (sparse fixed-ratio sampling)
  const size_t rowstride = (size_t)W * C;
  T* gin = gmap + ((size_t)batch * H * W + (X0 + k)) * C + c + (size_t)Y0 * rowstride;
  const v2_sptr tcol = tile + lane * (V * 4) + (px0 & 0xffff) * BINB;  // bins [.][q0 ..] of this lane
  const v2_sptr rowrec = v2_srec(plan_s, PW + PH + kV2MaxFW + j0);
  const int pwb = PW * BINB;
  switch (nq) {
    case 1: v2_bwd_rows<T, V, 1>(rowrec, FH, tcol, pwb, pxrec, nq, inv_count, gin, rowstride); break;
    case 2: v2_bwd_rows<T, V, 2>(rowrec, FH, tcol, pwb, pxrec, nq, inv_count, gin, rowstride); break;
    case 3: v2_bwd_rows<T, V, 3>(rowrec, FH, tcol, pwb, pxrec, nq, inv_count, gin, rowstride); break;
    case 4: v2_bwd_rows<T, V, 4>(rowrec, FH, tcol, pwb, pxrec, nq, inv_count, gin, rowstride); break;
    default: v2_bwd_rows<T, V, 0>(rowrec, FH, tcol, pwb, pxrec, nq, inv_count, gin, rowstride); break;
  }
}

// Per-sample backward of one bin column (ROIAlign_cuda.cu:177-254) for GENERIC RoIs, gradients from the tile.
template <typename T, int V>
ABR_DEV_COLD void v2_generic_bwd_column(const RoiGeom& g, int H, int W, T* gmap, v2_sptr tile, int pw, int c, int C, int PH, int PW,
                                   int lane) {
  const float fH = (float)H, fW = (float)W;
  T* img = gmap + (size_t)g.batch * H * W * C + c;
  for (int ph = 0; ph < PH; ph++) {
    float top[V];
    v2_sm_load<V>(tile + ((ph * PW + pw) * 32 + lane) * (V * 4), top);
    for (int iy = 0; iy < g.grid_h; iy++) {
      float y = v2_sample_coord(g.start_h, g.bin_h, ph, iy, g.grid_h);
      if (y < -1.0f || y > fH) continue;
      if (y <= 0.f) y = 0.f;
      int yl = (int)y, yh;
      if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
      const float ly = y - (float)yl, hy = 1.f - ly;
      for (int ix = 0; ix < g.grid_w; ix++) {
        float x = v2_sample_coord(g.start_w, g.bin_w, pw, ix, g.grid_w);
        if (x < -1.0f || x > fW) continue;
        if (x <= 0.f) x = 0.f;
        int xl = (int)x, xh;
        if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
        const float lx = x - (float)xl, hx = 1.f - lx;
        const float w[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
        const size_t off[4] = {((size_t)yl * W + xl) * C, ((size_t)yl * W + xh) * C, ((size_t)yh * W + xl) * C,
                               ((size_t)yh * W + xh) * C};
#pragma unroll
        for (int q = 0; q < 4; q++) {
          float v[V];
#pragma unroll
          for (int i = 0; i < V; i++) v[i] = top[i] * w[q] / g.count;
          VecIO<T, V>::red_add(img + off[q], v);
        }
      }
    }
  }
}

}  // namespace abr
