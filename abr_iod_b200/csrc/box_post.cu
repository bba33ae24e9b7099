// box_post.cu -- box-head post-processing (softmax, per-class decode, score threshold, class-batched NMS, per-image
// detection cut) for a whole batch on the device (sm_100a).
//
// Semantics: PostProcessor.forward / filter_results (modeling/roi_heads/box_head/inference.py:42-151 of the reference).
// Design (not a port).  The reference decodes every class, then loops over images and classes in Python: a `nonzero`,
// two fancy-index gathers, an NMS round trip to the host and a `torch.full` per (image, class), a `cat`, and a
// `kthvalue` on the CPU per image -- ~21 host synchronisations per image.  Here:
//   1. bp_candidates_kernel -- one CTA per (class, image): softmax probability of the class for every proposal of the
//                              image, score threshold, ORDERED compaction (block scan) of the surviving rows, decode of
//                              exactly those boxes (BoxCoder.decode + clip) -- only what passes the threshold is decoded;
//   2. nms_run              -- ONE batched NMS over all (image, class) segments with device-side box counts;
//   3. bp_assemble_kernel   -- one CTA per image: concatenation of the foreground classes in class order, the
//                              detections_per_img cut (radix select of the k-th largest score in shared memory; ties
//                              survive, as with the reference's `scores >= kthvalue`), labels, and class 0 apart.
// No host synchronisation; the caller reads the per-image counts once for the batch.
#include <vector>

#include "common.cuh"
#include "decode.cuh"

namespace abr {

int nms_run(const float* boxes, const float* scores, const int* offsets_host, const int* counts_dev,
            const unsigned long long* invalid, int n_images, float thresh, int ge, int max_keep, int64_t* keep,
            int keep_stride, int32_t* n_keep, void* workspace, size_t workspace_bytes, cudaStream_t st);

constexpr int kBpMaxImages = 64;
constexpr int kBpThreads = 256;

struct BpBatch {
  int n_images, first_image, C;
  int row_off[kBpMaxImages];  // first proposal row of the image
  int n[kBpMaxImages];        // proposals of the image
  int im_w[kBpMaxImages], im_h[kBpMaxImages];
};

// Segment (image, class j) owns candidate slots [C*row_off[img] + j*n[img], +n[img]).
__device__ __forceinline__ long long seg_base(const BpBatch& b, int img, int j) {
  return (long long)b.C * b.row_off[img] + (long long)j * b.n[img];
}

__global__ void __launch_bounds__(kBpThreads) bp_candidates_kernel(BpBatch b, const float* __restrict__ logits,
                                                                  const float* __restrict__ regression, int reg_stride,
                                                                  int cls_agnostic, const float* __restrict__ proposals,
                                                                  BoxCoderParams coder, float score_thresh,
                                                                  float4* __restrict__ boxes_c, float* __restrict__ scores_c,
                                                                  int* __restrict__ rows_c, int* __restrict__ seg_count) {
  __shared__ int warp_tot[kBpThreads / 32];
  const int j = blockIdx.x, img = blockIdx.y, C = b.C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = b.n[img], row0 = b.row_off[img];
  const long long base = seg_base(b, img, j);
  const float xmax = (float)(b.im_w[img] - 1), ymax = (float)(b.im_h[img] - 1);
  int running = 0;
  for (int r0 = 0; r0 < n; r0 += kBpThreads) {
    const int r = r0 + tid;
    bool ok = false;
    float p = 0.f;
    if (r < n) {
      // softmax over the row's C logits (inference.py:56): exp(x - max) / sum
      const float* l = logits + (size_t)(row0 + r) * C;
      float m = -INFINITY;
      for (int c = 0; c < C; c++) m = fmaxf(m, __ldg(l + c));
      float sum = 0.f;
      for (int c = 0; c < C; c++) sum = __fadd_rn(sum, expf(__fsub_rn(__ldg(l + c), m)));
      p = __fdiv_rn(expf(__fsub_rn(__ldg(l + j), m)), sum);
      ok = p > score_thresh;  // inference.py:117
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) warp_tot[warp] = __popc(ballot);
    __syncthreads();
    int pos = running + __popc(ballot & ((1u << lane) - 1u));
    int chunk = 0;
    for (int w2 = 0; w2 < kBpThreads / 32; w2++) {
      if (w2 < warp) pos += warp_tot[w2];
      chunk += warp_tot[w2];
    }
    if (ok) {
      const float* rg = regression + (size_t)(row0 + r) * reg_stride + (cls_agnostic ? reg_stride - 4 : 4 * j);
      const float4 an = __ldg(reinterpret_cast<const float4*>(proposals) + row0 + r);
      boxes_c[base + pos] = decode_and_clip(an, __ldg(rg), __ldg(rg + 1), __ldg(rg + 2), __ldg(rg + 3), coder, xmax, ymax);
      scores_c[base + pos] = p;
      rows_c[base + pos] = r;
    }
    running += chunk;
    __syncthreads();
  }
  if (tid == 0) seg_count[(b.first_image + img) * C + j] = running;
}

// One CTA per image.  Temp arrays t_* hold the concatenation of the kept foreground boxes (capacity (C-1)*n).
__global__ void __launch_bounds__(kBpThreads) bp_assemble_kernel(
    BpBatch b, const float4* __restrict__ boxes_c, const float* __restrict__ scores_c, const int* __restrict__ rows_c,
    const int* __restrict__ seg_count, const long long* __restrict__ keep, int keep_stride, const int* __restrict__ n_keep,
    int detections_per_img, float4* __restrict__ t_boxes, float* __restrict__ t_scores, int* __restrict__ t_rows,
    int* __restrict__ t_labels, float4* __restrict__ det_boxes, float* __restrict__ det_scores,
    long long* __restrict__ det_labels, int* __restrict__ det_rows, int* __restrict__ n_det, int det_stride,
    float4* __restrict__ bg_boxes, float* __restrict__ bg_scores, int* __restrict__ n_bg, int bg_stride) {
  extern __shared__ int class_off[];  // [C + 1]: exclusive scan of the kept counts of classes 1..C-1
  __shared__ unsigned hist[256];
  __shared__ unsigned sel_prefix, sel_mask;
  __shared__ int sel_remaining, running_sh;
  __shared__ int warp_tot[kBpThreads / 32];
  const int img = blockIdx.x, g = b.first_image + img, C = b.C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // kept count of a segment: the NMS result, or (no NMS) every candidate
  auto kept_of = [&](int j) { return keep ? n_keep[g * C + j] : seg_count[g * C + j]; };
  auto src_of = [&](int j, int i) { return keep ? (int)keep[(size_t)(g * C + j) * keep_stride + i] : i; };
  if (tid == 0) {
    int acc = 0;
    class_off[0] = 0;
    for (int j = 1; j < C; j++) { class_off[j] = acc; acc += kept_of(j); }
    class_off[C] = acc;
  }
  __syncthreads();
  const int D = class_off[C];
  const long long tbase = seg_base(b, img, 0);  // temp capacity C*n per image, same addressing as the candidates
  // concatenate the foreground classes in class order (inference.py:133-139)
  for (int j = 1; j < C; j++) {
    const int cnt = class_off[j + 1 <= C - 1 ? j + 1 : C] - class_off[j];
    const long long sb = seg_base(b, img, j);
    for (int i = tid; i < cnt; i += kBpThreads) {
      const int src = src_of(j, i);
      const long long o = tbase + class_off[j] + i;
      t_boxes[o] = boxes_c[sb + src];
      t_scores[o] = scores_c[sb + src];
      t_rows[o] = rows_c[sb + src];
      t_labels[o] = j;
    }
  }
  // class 0 apart (inference.py:137-138)
  {
    const int cnt = kept_of(0);
    const long long sb = seg_base(b, img, 0);
    for (int i = tid; i < bg_stride; i += kBpThreads) {
      float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
      float sc = 0.f;
      if (i < cnt) {
        const int src = src_of(0, i);
        bx = boxes_c[sb + src];
        sc = scores_c[sb + src];
      }
      bg_boxes[(size_t)g * bg_stride + i] = bx;
      bg_scores[(size_t)g * bg_stride + i] = sc;
    }
    if (tid == 0) n_bg[g] = cnt;
  }
  __syncthreads();  // (global writes of this CTA are visible to this CTA after the barrier)
  // detections_per_img cut (inference.py:142-149): threshold = the det-th largest score, keep scores >= threshold
  unsigned thr_key = 0;  // keep everything
  if (detections_per_img > 0 && D > detections_per_img) {
    if (tid == 0) { sel_prefix = 0; sel_mask = 0; sel_remaining = detections_per_img; }
    __syncthreads();
    for (int pass = 0; pass < 4; pass++) {
      const int shift = 24 - 8 * pass;
      for (int i = tid; i < 256; i += kBpThreads) hist[i] = 0;
      __syncthreads();
      const unsigned prefix = sel_prefix, mask = sel_mask;
      for (int i = tid; i < D; i += kBpThreads) {
        const unsigned key = ordered_bits(t_scores[tbase + i]);
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        int above = 0, rem = sel_remaining;
        for (int bin = 255; bin >= 0; bin--) {
          const int c = (int)hist[bin];
          if (above + c >= rem) {
            sel_prefix = prefix | ((unsigned)bin << shift);
            sel_mask = mask | (255u << shift);
            sel_remaining = rem - above;
            break;
          }
          above += c;
        }
      }
      __syncthreads();
    }
    thr_key = sel_prefix;
  }
  // ordered compaction of the survivors into the outputs
  if (tid == 0) running_sh = 0;
  __syncthreads();
  float4* ob = det_boxes + (size_t)g * det_stride;
  float* os = det_scores + (size_t)g * det_stride;
  long long* ol = det_labels + (size_t)g * det_stride;
  int* orow = det_rows ? det_rows + (size_t)g * det_stride : nullptr;
  for (int i0 = 0; i0 < D; i0 += kBpThreads) {
    const int i = i0 + tid;
    const bool ok = i < D && ordered_bits(t_scores[tbase + i]) >= thr_key;
    const unsigned ballot = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) warp_tot[warp] = __popc(ballot);
    __syncthreads();
    int pos = running_sh + __popc(ballot & ((1u << lane) - 1u));
    int chunk = 0;
    for (int w2 = 0; w2 < kBpThreads / 32; w2++) {
      if (w2 < warp) pos += warp_tot[w2];
      chunk += warp_tot[w2];
    }
    if (ok && pos < det_stride) {
      ob[pos] = t_boxes[tbase + i];
      os[pos] = t_scores[tbase + i];
      ol[pos] = t_labels[tbase + i];
      if (orow) orow[pos] = t_rows[tbase + i];
    }
    __syncthreads();
    if (tid == 0) running_sh += chunk;
    __syncthreads();
  }
  const int total = running_sh;
  for (int i = min(total, det_stride) + tid; i < det_stride; i += kBpThreads) {
    ob[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    os[i] = 0.f;
    ol[i] = -1;
    if (orow) orow[i] = -1;
  }
  if (tid == 0) n_det[g] = total;  // may exceed det_stride when equal scores tie at the cut: the caller re-runs wider
}

struct BpLayout {
  size_t seg_count, n_keep, boxes, scores, rows, t_boxes, t_scores, t_rows, t_labels, keep, nms, total, nms_bytes;
  int max_n, segments;
  long long slots;
};

static size_t bp_a256(size_t x) { return (x + 255) & ~(size_t)255; }

static BpLayout bp_layout(const int* boxes_per_image, int n_images, int C) {
  BpLayout l;
  long long R = 0;
  l.max_n = 0;
  for (int i = 0; i < n_images; i++) {
    R += boxes_per_image[i];
    l.max_n = boxes_per_image[i] > l.max_n ? boxes_per_image[i] : l.max_n;
  }
  l.slots = R * C;
  l.segments = n_images * C;
  size_t o = 0;
  l.seg_count = o; o += bp_a256((size_t)l.segments * 4);
  l.n_keep = o; o += bp_a256((size_t)l.segments * 4);
  l.boxes = o; o += bp_a256((size_t)l.slots * 16);
  l.scores = o; o += bp_a256((size_t)l.slots * 4);
  l.rows = o; o += bp_a256((size_t)l.slots * 4);
  l.t_boxes = o; o += bp_a256((size_t)l.slots * 16);
  l.t_scores = o; o += bp_a256((size_t)l.slots * 4);
  l.t_rows = o; o += bp_a256((size_t)l.slots * 4);
  l.t_labels = o; o += bp_a256((size_t)l.slots * 4);
  l.keep = o; o += bp_a256((size_t)l.segments * (l.max_n > 0 ? l.max_n : 1) * 8);
  std::vector<int> offsets((size_t)l.segments + 1);
  offsets[0] = 0;
  for (int i = 0; i < n_images; i++)
    for (int j = 0; j < C; j++) offsets[(size_t)i * C + j + 1] = offsets[(size_t)i * C + j] + boxes_per_image[i];
  l.nms_bytes = bp_a256(abr_nms_workspace_bytes(offsets.data(), l.segments));
  l.nms = o; o += l.nms_bytes;
  l.total = o;
  return l;
}

}  // namespace abr

using namespace abr;

extern "C" {

size_t abr_box_postprocess_workspace_bytes(const int* boxes_per_image_host, int n_images, int num_classes) {
  if (!boxes_per_image_host || n_images <= 0 || num_classes <= 0) return 0;
  return bp_layout(boxes_per_image_host, n_images, num_classes).total;
}

int abr_box_postprocess(const float* class_logits, const float* box_regression, int reg_row_stride, int cls_agnostic,
                        const float* proposals, const int* boxes_per_image_host, const int* image_sizes_host, int n_images,
                        int num_classes, float score_thresh, float nms_thresh, int ge, int detections_per_img,
                        const float* weights4_host, float bbox_xform_clip, float* det_boxes, float* det_scores,
                        int64_t* det_labels, int32_t* det_rows, int32_t* n_det, int det_stride, float* bg_boxes,
                        float* bg_scores, int32_t* n_bg, int bg_stride, void* workspace, size_t workspace_bytes,
                        abr_stream_t stream) {
  ABR_REQUIRE(n_images >= 0, ABR_ERR_BAD_ARG, "box_post: n_images=%d", n_images);
  if (n_images == 0) return ABR_OK;
  const int C = num_classes;
  ABR_REQUIRE(C >= 1 && C <= 4096, ABR_ERR_BAD_ARG, "box_post: num_classes=%d", C);
  ABR_REQUIRE(boxes_per_image_host && image_sizes_host && weights4_host && det_boxes && det_scores && det_labels && n_det &&
                  bg_boxes && bg_scores && n_bg,
              ABR_ERR_BAD_ARG, "box_post: null pointer");
  ABR_REQUIRE(reg_row_stride >= (cls_agnostic ? 4 : 4 * C), ABR_ERR_BAD_ARG, "box_post: box_regression rows of %d floats, need %d",
              reg_row_stride, cls_agnostic ? 4 : 4 * C);
  ABR_REQUIRE(det_stride >= 0 && bg_stride >= 0, ABR_ERR_BAD_ARG, "box_post: negative stride");
  long long R = 0;
  for (int i = 0; i < n_images; i++) {
    ABR_REQUIRE(boxes_per_image_host[i] >= 0, ABR_ERR_BAD_ARG, "box_post: negative box count");
    ABR_REQUIRE(bg_stride >= boxes_per_image_host[i], ABR_ERR_BAD_ARG, "box_post: bg_stride %d < %d", bg_stride, boxes_per_image_host[i]);
    R += boxes_per_image_host[i];
  }
  ABR_REQUIRE(R * C < (1ll << 31), ABR_ERR_UNSUPPORTED, "box_post: %lld candidate slots", R * C);
  if (R > 0) ABR_REQUIRE(class_logits && box_regression && proposals, ABR_ERR_BAD_ARG, "box_post: null inputs");
  ABR_REQUIRE(((reinterpret_cast<uintptr_t>(proposals) | reinterpret_cast<uintptr_t>(det_boxes) | reinterpret_cast<uintptr_t>(bg_boxes)) & 15) == 0,
              ABR_ERR_BAD_ARG, "box_post: proposals / box outputs must be 16-byte aligned");
  const BpLayout lay = bp_layout(boxes_per_image_host, n_images, C);
  ABR_REQUIRE(workspace && workspace_bytes >= lay.total && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, ABR_ERR_WORKSPACE,
              "box_post: workspace %zu B < %zu B (or not 256-byte aligned)", workspace_bytes, lay.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  int* seg_count = reinterpret_cast<int*>(ws + lay.seg_count);
  int* n_keep = reinterpret_cast<int*>(ws + lay.n_keep);
  float4* boxes_c = reinterpret_cast<float4*>(ws + lay.boxes);
  float* scores_c = reinterpret_cast<float*>(ws + lay.scores);
  int* rows_c = reinterpret_cast<int*>(ws + lay.rows);
  long long* keep = reinterpret_cast<long long*>(ws + lay.keep);
  BoxCoderParams coder;
  coder.wx = weights4_host[0]; coder.wy = weights4_host[1]; coder.ww = weights4_host[2]; coder.wh = weights4_host[3];
  coder.clip = bbox_xform_clip;

  std::vector<BpBatch> groups;
  int row = 0;
  for (int base = 0; base < n_images; base += kBpMaxImages) {
    BpBatch b;
    b.first_image = base; b.C = C;
    b.n_images = n_images - base < kBpMaxImages ? n_images - base : kBpMaxImages;
    for (int i = 0; i < b.n_images; i++) {
      b.row_off[i] = row;
      b.n[i] = boxes_per_image_host[base + i];
      b.im_w[i] = image_sizes_host[2 * (base + i)];
      b.im_h[i] = image_sizes_host[2 * (base + i) + 1];
      row += b.n[i];
    }
    groups.push_back(b);
  }
  for (const BpBatch& b : groups) {
    bp_candidates_kernel<<<dim3(C, b.n_images), kBpThreads, 0, st>>>(b, class_logits, box_regression, reg_row_stride, cls_agnostic,
                                                                    proposals, coder, score_thresh, boxes_c, scores_c, rows_c, seg_count);
    ABR_CHECK_LAUNCH("box_post_candidates");
  }
  const bool run_nms = nms_thresh > 0.f;  // structures/boxlist_ops.py:22-23
  const int keep_stride = lay.max_n > 0 ? lay.max_n : 1;
  if (run_nms) {
    std::vector<int> offsets((size_t)lay.segments + 1);
    offsets[0] = 0;
    for (int i = 0; i < n_images; i++)
      for (int j = 0; j < C; j++) offsets[(size_t)i * C + j + 1] = offsets[(size_t)i * C + j] + boxes_per_image_host[i];
    int rc = nms_run(reinterpret_cast<const float*>(boxes_c), scores_c, offsets.data(), seg_count, nullptr, lay.segments, nms_thresh,
                     ge, -1, reinterpret_cast<int64_t*>(keep), keep_stride, n_keep, ws + lay.nms, lay.nms_bytes, st);
    if (rc) return rc;
  }
  const size_t smem = (size_t)(C + 1) * sizeof(int);
  ABR_REQUIRE(smem <= 40 * 1024, ABR_ERR_UNSUPPORTED, "box_post: too many classes");
  for (const BpBatch& b : groups) {
    bp_assemble_kernel<<<b.n_images, kBpThreads, smem, st>>>(
        b, boxes_c, scores_c, rows_c, seg_count, run_nms ? keep : nullptr, keep_stride, n_keep, detections_per_img,
        reinterpret_cast<float4*>(ws + lay.t_boxes), reinterpret_cast<float*>(ws + lay.t_scores),
        reinterpret_cast<int*>(ws + lay.t_rows), reinterpret_cast<int*>(ws + lay.t_labels),
        reinterpret_cast<float4*>(det_boxes), det_scores, reinterpret_cast<long long*>(det_labels), det_rows, n_det, det_stride,
        reinterpret_cast<float4*>(bg_boxes), bg_scores, n_bg, bg_stride);
    ABR_CHECK_LAUNCH("box_post_assemble");
  }
  return ABR_OK;
}

}  // extern "C"
