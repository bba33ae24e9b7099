// match.cu -- proposal <-> ground-truth matching for the box head's RoI sampling, one launch per batch (sm_100a).
//
// Semantics: FastRCNNLossComputation.match_targets_to_proposals + prepare_targets (modeling/roi_heads/box_head/loss.py:
// 43-84 of the reference) = boxlist_iou (structures/boxlist_ops.py:53-88) -> Matcher without low-quality matches
// (modeling/matcher.py:52-81) -> labels (0 = background, -1 = ignored) -> BoxCoder.encode (modeling/box_coder.py:22-50).
// The reference materialises the [G, n] IoU matrix and runs ~35 tensor kernels per image from a Python loop; here one
// thread owns one proposal, walks its image's ground-truth boxes (a handful, staged in shared memory) and writes the
// matched index, the label and the regression target.  IoU uses individually rounded fp32 operations in the reference's
// order, so the threshold decisions and the arg-max (first maximum wins, like torch.max) are exact.
#include "common.cuh"

namespace abr {

constexpr int kMatchImages = 64;
constexpr int kMatchMaxGt = 1024;  // ground-truth boxes of one image staged in shared memory

struct MatchBatch {
  int n_images;
  int row_off[kMatchImages], n[kMatchImages];  // proposals
  int gt_off[kMatchImages], g[kMatchImages];   // ground truth
};

__device__ __forceinline__ float area_plus_one(const float4 b) {
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
}

// boxlist_ops.py:74-87 for one pair (box1 = ground truth, box2 = proposal)
__device__ __forceinline__ float pair_iou(const float4 a, float area_a, const float4 b, float area_b) {
  const float w = fmaxf(__fadd_rn(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 1.f), 0.f);
  const float h = fmaxf(__fadd_rn(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 1.f), 0.f);
  const float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}

__global__ void __launch_bounds__(256) match_kernel(MatchBatch mb, const float4* __restrict__ proposals,
                                                    const float4* __restrict__ gt_boxes, const long long* __restrict__ gt_labels,
                                                    float high, float low, float wx, float wy, float ww, float wh,
                                                    long long* __restrict__ matched, long long* __restrict__ labels,
                                                    float4* __restrict__ targets) {
  __shared__ float4 sbox[kMatchMaxGt];
  __shared__ float sarea[kMatchMaxGt];
  const int img = blockIdx.y;
  const int G = mb.g[img], n = mb.n[img];
  for (int i = threadIdx.x; i < G; i += blockDim.x) {
    const float4 b = gt_boxes[mb.gt_off[img] + i];
    sbox[i] = b;
    sarea[i] = area_plus_one(b);
  }
  __syncthreads();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float4 p = proposals[mb.row_off[img] + r];
  const float pa = area_plus_one(p);
  float best = -INFINITY;
  int arg = 0;
  for (int i = 0; i < G; i++) {
    const float v = pair_iou(sbox[i], sarea[i], p, pa);
    if (v > best || (v != v && best == best)) { best = v; arg = i; }  // first maximum wins; NaN counts as the maximum
  }
  // matcher.py:72-78
  long long m = arg;
  if (best < low) m = -1;                       // BELOW_LOW_THRESHOLD
  else if (best >= low && best < high) m = -2;  // BETWEEN_THRESHOLDS
  const int gi = mb.gt_off[img] + (m < 0 ? 0 : arg);  // matched_idxs.clamp(min=0)
  long long lab = gt_labels[gi];
  if (m == -1) lab = 0;
  if (m == -2) lab = -1;
  // box_coder.py:33-50 with reference box = matched ground truth
  const float4 ref = sbox[m < 0 ? 0 : arg];
  const float ex_w = __fadd_rn(__fsub_rn(p.z, p.x), 1.f), ex_h = __fadd_rn(__fsub_rn(p.w, p.y), 1.f);
  const float ex_cx = __fadd_rn(p.x, __fmul_rn(0.5f, ex_w)), ex_cy = __fadd_rn(p.y, __fmul_rn(0.5f, ex_h));
  const float gt_w = __fadd_rn(__fsub_rn(ref.z, ref.x), 1.f), gt_h = __fadd_rn(__fsub_rn(ref.w, ref.y), 1.f);
  const float gt_cx = __fadd_rn(ref.x, __fmul_rn(0.5f, gt_w)), gt_cy = __fadd_rn(ref.y, __fmul_rn(0.5f, gt_h));
  float4 t;
  t.x = __fdiv_rn(__fmul_rn(wx, __fsub_rn(gt_cx, ex_cx)), ex_w);
  t.y = __fdiv_rn(__fmul_rn(wy, __fsub_rn(gt_cy, ex_cy)), ex_h);
  t.z = __fmul_rn(ww, logf(__fdiv_rn(gt_w, ex_w)));
  t.w = __fmul_rn(wh, logf(__fdiv_rn(gt_h, ex_h)));
  const int o = mb.row_off[img] + r;
  matched[o] = m;
  labels[o] = lab;
  targets[o] = t;
}

// boxlist_iou: out[i][j] = IoU(boxes1[i], boxes2[j])
__global__ void box_iou_kernel(const float4* __restrict__ b1, int N, const float4* __restrict__ b2, int M, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * M) return;
  const int i = (int)(t / M), j = (int)(t - (long long)i * M);
  const float4 a = b1[i], b = b2[j];
  out[t] = pair_iou(a, area_plus_one(a), b, area_plus_one(b));
}

}  // namespace abr

using namespace abr;

extern "C" {

int abr_match_proposals(const float* proposals, const int* boxes_per_image_host, const float* gt_boxes, const int64_t* gt_labels,
                        const int* gt_per_image_host, int n_images, float high_threshold, float low_threshold,
                        const float* weights4_host, int64_t* matched_idxs, int64_t* labels, float* regression_targets,
                        abr_stream_t stream) {
  ABR_REQUIRE(n_images >= 0, ABR_ERR_BAD_ARG, "match: n_images=%d", n_images);
  if (n_images == 0) return ABR_OK;
  ABR_REQUIRE(boxes_per_image_host && gt_per_image_host && weights4_host, ABR_ERR_BAD_ARG, "match: null host array");
  ABR_REQUIRE(proposals && gt_boxes && gt_labels && matched_idxs && labels && regression_targets, ABR_ERR_BAD_ARG, "match: null pointer");
  ABR_REQUIRE(((reinterpret_cast<uintptr_t>(proposals) | reinterpret_cast<uintptr_t>(gt_boxes) | reinterpret_cast<uintptr_t>(regression_targets)) & 15) == 0,
              ABR_ERR_BAD_ARG, "match: box tensors must be 16-byte aligned");
  ABR_REQUIRE(low_threshold <= high_threshold, ABR_ERR_BAD_ARG, "match: low_threshold > high_threshold");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int row = 0, gt = 0;
  for (int base = 0; base < n_images; base += kMatchImages) {
    MatchBatch mb;
    mb.n_images = n_images - base < kMatchImages ? n_images - base : kMatchImages;
    int nmax = 0;
    for (int i = 0; i < mb.n_images; i++) {
      const int n = boxes_per_image_host[base + i], g = gt_per_image_host[base + i];
      // matcher.py:53-62: empty targets or proposals are not supported during training
      ABR_REQUIRE(g > 0, ABR_ERR_BAD_ARG, "match: no ground-truth boxes for image %d", base + i);
      ABR_REQUIRE(n > 0, ABR_ERR_BAD_ARG, "match: no proposal boxes for image %d", base + i);
      ABR_REQUIRE(g <= kMatchMaxGt, ABR_ERR_UNSUPPORTED, "match: %d ground-truth boxes in image %d (max %d)", g, base + i, kMatchMaxGt);
      mb.row_off[i] = row; mb.n[i] = n; mb.gt_off[i] = gt; mb.g[i] = g;
      row += n; gt += g;
      nmax = n > nmax ? n : nmax;
    }
    match_kernel<<<dim3(ceil_div(nmax, 256), mb.n_images), 256, 0, st>>>(
        mb, reinterpret_cast<const float4*>(proposals), reinterpret_cast<const float4*>(gt_boxes),
        reinterpret_cast<const long long*>(gt_labels), high_threshold, low_threshold, weights4_host[0], weights4_host[1],
        weights4_host[2], weights4_host[3], reinterpret_cast<long long*>(matched_idxs), reinterpret_cast<long long*>(labels),
        reinterpret_cast<float4*>(regression_targets));
    ABR_CHECK_LAUNCH("match_proposals");
  }
  return ABR_OK;
}

int abr_box_iou(const float* boxes1, int N, const float* boxes2, int M, float* iou, abr_stream_t stream) {
  ABR_REQUIRE(N >= 0 && M >= 0, ABR_ERR_BAD_ARG, "box_iou: N=%d M=%d", N, M);
  if (N == 0 || M == 0) return ABR_OK;
  ABR_REQUIRE(boxes1 && boxes2 && iou, ABR_ERR_BAD_ARG, "box_iou: null pointer");
  ABR_REQUIRE(((reinterpret_cast<uintptr_t>(boxes1) | reinterpret_cast<uintptr_t>(boxes2)) & 15) == 0, ABR_ERR_BAD_ARG,
              "box_iou: boxes must be 16-byte aligned");
  const long long total = (long long)N * M;
  box_iou_kernel<<<(unsigned)ceil_div<long long>(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(boxes1), N, reinterpret_cast<const float4*>(boxes2), M, iou);
  ABR_CHECK_LAUNCH("box_iou");
  return ABR_OK;
}

}  // extern "C"
