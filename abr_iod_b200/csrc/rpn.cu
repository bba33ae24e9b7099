// rpn.cu -- RPN proposal selection around NMS for a whole batch, entirely on the device (sm_100a).
//
// Semantics: RPNPostProcessor.forward_for_single_feature_map (modeling/rpn/inference.py:76-118 of the reference):
// sigmoid -> top pre_nms_top_n per image -> gather regression + anchors -> BoxCoder.decode (modeling/box_coder.py:52-95)
// -> clip_to_image (structures/bounding_box.py:214-219) -> remove_small_boxes (structures/boxlist_ops.py:34-48) ->
// boxlist_nms with max_proposals = post_nms_top_n (structures/boxlist_ops.py:9-31).
// Design (not a port).  The reference permutes and copies both head outputs, runs a library top-k, three fancy-index
// gathers, ~25 elementwise kernels and then loops over the images in Python with a `nonzero` and an NMS round trip per
// image.  Here the head outputs are read in place (NCHW or channels-last) and the batch is:
//   1. rpn_hist_kernel x5   -- MSB-first radix SELECT of the pre_nms_top_n-th key on 54-bit keys
//                              (order-preserving logit bits : 22 bits of reversed anchor index), 11 bits per pass, many
//                              CTAs per image; a pass whose bin holds exactly the remaining count ends the search, so the
//                              index digits only run when equal logits straddle the cut.  Ranking by logit instead of
//                              by sigmoid(logit) is the same order (monotonic) without a transcendental in the key;
//   2. rpn_collect_kernel   -- the selected keys, unordered, into a per-image candidate list;
//   3. rpn_chunk_sort_kernel -- bitonic sort of 2048-key chunks in shared memory, one CTA per chunk;
//   4. rpn_rank_decode_kernel -- one thread per candidate: rank = position in its chunk + binary searches in the other
//                              chunks (a merge by counting), then decode + clip, written AT its rank; boxes failing the
//                              size filter are only flagged in a bitmap (no compaction pass);
//   5. nms_run              -- the batched NMS of nms.cu, started with the flagged boxes already suppressed (the inputs
//                              are sorted, so it takes its O(N) presort path and the prefix pass);
//   6. rpn_gather_kernel    -- kept boxes / scores / anchor indices into the padded outputs + per-image counts.
// No host synchronisation, no allocation; the caller reads n_out once for the batch.
#include <cfloat>
#include <vector>

#include "common.cuh"
#include "decode.cuh"

namespace abr {

int nms_run(const float* boxes, const float* scores, const int* offsets_host, const int* counts_dev,
            const unsigned long long* invalid, int n_images, float thresh, int ge, int max_keep, int64_t* keep,
            int keep_stride, int32_t* n_keep, void* workspace, size_t workspace_bytes, cudaStream_t st);

constexpr int kRpnBins = 2048;        // 11-bit digits
constexpr int kRpnPasses = 5;         // 11 + 11 + 10 bits of logit, 11 + 11 bits of index
constexpr int kRpnIdxBits = 22;       // anchors per image < 2^22
constexpr int kRpnMaxCand = 16384;    // candidates one CTA sorts in shared memory
constexpr int kRpnChunk = 2048;       // elements per CTA in the histogram / collect kernels
constexpr int kRpnMaxImages = 64;     // images per launch group (image sizes travel as kernel arguments)

__device__ __forceinline__ int rpn_shift(int pass) {
  // digit positions inside the 54-bit key: [53:43] [42:32] [31:22] [21:11] [10:0]
  return pass == 0 ? 43 : pass == 1 ? 32 : pass == 2 ? 22 : pass == 3 ? 11 : 0;
}
__device__ __forceinline__ unsigned rpn_digit_mask(int pass) { return pass == 2 ? 1023u : 2047u; }

__device__ __forceinline__ unsigned long long rpn_key(float logit, int anchor) {
  return ((unsigned long long)ordered_bits(logit) << kRpnIdxBits) | (unsigned long long)(((1u << kRpnIdxBits) - 1u) - (unsigned)anchor);
}

struct RpnShape {
  int N, A, H, W, layout, k;  // k = min(pre_nms_top_n, A*H*W)
};

// storage index of an objectness element -> anchor index in the reference's flattened (h, w, a) order
__device__ __forceinline__ int anchor_of_storage(const RpnShape& s, int e) {
  if (s.layout == ABR_NHWC) return e;
  const int hw = s.H * s.W;
  const int a = e / hw;
  return (e - a * hw) * s.A + a;
}

struct SelectState {
  unsigned long long prefix, mask;
  int remaining;
  int done;  // a bin held exactly the remaining count: everything matching the prefix is taken
};

// Replays passes [0, upto) from their finished histograms.  Block-wide; every CTA of an image derives the same state.
__device__ void derive_state(const unsigned* __restrict__ hist_img, int upto, int k, SelectState& out_state) {
  __shared__ SelectState st;
  __shared__ int warp_tot[32];
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  if (tid == 0) { st.prefix = 0; st.mask = 0; st.remaining = k; st.done = 0; }
  __syncthreads();
  const int per = kRpnBins / nthr;  // bins per thread, walked from the top bin down
  for (int q = 0; q < upto; q++) {
    if (st.done) break;
    const unsigned* h = hist_img + q * kRpnBins;
    const int top = kRpnBins - 1 - tid * per;  // this thread owns bins top, top-1, ..., top-per+1
    int mine = 0;
    for (int b = 0; b < per; b++) mine += (int)h[top - b];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int before = 0;
    for (int w2 = 0; w2 < warp; w2++) before += warp_tot[w2];
    (void)nwarp;
    incl += before;
    const int excl = incl - mine;
    const int remaining = st.remaining;
    __syncthreads();  // everyone has read st.remaining and warp_tot
    if (excl < remaining && remaining <= incl) {  // the crossing bin is one of mine (exactly one thread gets here)
      int above = excl;
      for (int b = 0; b < per; b++) {
        const int c = (int)h[top - b];
        if (above + c >= remaining) {
          const int sh = rpn_shift(q);
          st.prefix |= (unsigned long long)(top - b) << sh;
          st.mask |= (unsigned long long)rpn_digit_mask(q) << sh;
          st.remaining = remaining - above;
          st.done = (c == remaining - above);
          break;
        }
        above += c;
      }
    }
    __syncthreads();
  }
  out_state = st;
  __syncthreads();
}

// One radix-select pass: histogram of digit `pass` over the elements that match the prefix found so far.
__global__ void __launch_bounds__(256) rpn_hist_kernel(RpnShape s, const float* __restrict__ objectness,
                                                       unsigned* __restrict__ hist, int pass) {
  __shared__ unsigned sh[kRpnBins];
  const int img = blockIdx.y;
  const int M = s.A * s.H * s.W;
  unsigned* hist_img = hist + (size_t)img * kRpnPasses * kRpnBins;
  SelectState st;
  derive_state(hist_img, pass, s.k, st);
  if (st.done) return;
  for (int b = threadIdx.x; b < kRpnBins; b += blockDim.x) sh[b] = 0;
  __syncthreads();
  const float* obj = objectness + (size_t)img * M;
  const int shift = rpn_shift(pass);
  const unsigned dmask = rpn_digit_mask(pass);
  const int e0 = blockIdx.x * kRpnChunk;
#pragma unroll
  for (int u = 0; u < kRpnChunk / 256; u++) {
    const int e = e0 + u * 256 + threadIdx.x;
    if (e < M) {
      const unsigned long long key = rpn_key(__ldg(obj + e), anchor_of_storage(s, e));
      if ((key & st.mask) == st.prefix) atomicAdd(&sh[(unsigned)(key >> shift) & dmask], 1u);
    }
  }
  __syncthreads();
  unsigned* g = hist_img + pass * kRpnBins;
  for (int b = threadIdx.x; b < kRpnBins; b += blockDim.x)
    if (sh[b]) atomicAdd(&g[b], sh[b]);
}

// The selected keys (exactly k per image) into cand[img][0..k), unordered.
__global__ void __launch_bounds__(256) rpn_collect_kernel(RpnShape s, const float* __restrict__ objectness,
                                                          const unsigned* __restrict__ hist,
                                                          unsigned long long* __restrict__ cand, int* __restrict__ cand_count) {
  __shared__ unsigned long long list[kRpnChunk];
  __shared__ int n_local, base;
  const int img = blockIdx.y;
  const int M = s.A * s.H * s.W;
  SelectState st;
  derive_state(hist + (size_t)img * kRpnPasses * kRpnBins, kRpnPasses, s.k, st);
  if (threadIdx.x == 0) n_local = 0;
  __syncthreads();
  const float* obj = objectness + (size_t)img * M;
  const int e0 = blockIdx.x * kRpnChunk;
#pragma unroll
  for (int u = 0; u < kRpnChunk / 256; u++) {
    const int e = e0 + u * 256 + threadIdx.x;
    if (e < M) {
      const unsigned long long key = rpn_key(__ldg(obj + e), anchor_of_storage(s, e));
      if ((key & st.mask) >= st.prefix) list[atomicAdd(&n_local, 1)] = key;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) base = n_local ? atomicAdd(&cand_count[img], n_local) : 0;
  __syncthreads();
  unsigned long long* out = cand + (size_t)img * s.k;
  for (int i = threadIdx.x; i < n_local; i += blockDim.x)
    if (base + i < s.k) out[base + i] = list[i];
}

struct RpnDecode {
  float wx, wy, ww, wh, clip, min_size;
  long long anchor_image_stride;  // floats between the anchor sets of consecutive images (0 = shared)
  int im_w[kRpnMaxImages], im_h[kRpnMaxImages];
};

constexpr int kSortChunk = 2048;   // keys one CTA sorts (two per thread)
constexpr int kSortThreads = 1024;
constexpr int kMaxChunks = kRpnMaxCand / kSortChunk;

// Stage 1 of the sort: every CTA sorts one 2048-key chunk of an image's candidates in shared memory (bitonic, descending,
// one compare-exchange per thread and stage), in place.
__global__ void __launch_bounds__(kSortThreads) rpn_chunk_sort_kernel(int k, unsigned long long* __restrict__ cand) {
  __shared__ unsigned long long keys[kSortChunk];
  const int img = blockIdx.y, tid = threadIdx.x;
  const int c0 = blockIdx.x * kSortChunk;
  unsigned long long* io = cand + (size_t)img * k;
  keys[tid] = c0 + tid < k ? io[c0 + tid] : 0ull;  // real keys are > 0: the padding sorts last
  keys[tid + kSortThreads] = c0 + tid + kSortThreads < k ? io[c0 + tid + kSortThreads] : 0ull;
  __syncthreads();
  for (int size = 2; size <= kSortChunk; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int i = 2 * tid - (tid & (stride - 1));
      const int j = i + stride;
      const unsigned long long a = keys[i], b = keys[j];
      const bool desc = (i & size) == 0;
      if ((a < b) == desc) { keys[i] = b; keys[j] = a; }
      __syncthreads();
    }
  }
  if (c0 + tid < k) io[c0 + tid] = keys[tid];
  if (c0 + tid + kSortThreads < k) io[c0 + tid + kSortThreads] = keys[tid + kSortThreads];
}

// Stage 2: the rank of a candidate = its position in its own sorted chunk + the number of larger keys in every other
// chunk (one binary search per chunk, all in flight together; keys are unique).  The candidate is decoded and written at
// its rank, so the outputs are in (logit descending, anchor ascending) order without a merge pass.  Boxes that fail the
// size filter stay in place and are flagged in `invalid` (one bit per rank) -- the NMS starts with them suppressed.
__global__ void __launch_bounds__(256) rpn_rank_decode_kernel(RpnShape s, RpnDecode d, int first_image,
                                                              const unsigned long long* __restrict__ cand,
                                                              const float* __restrict__ box_regression,
                                                              const float* __restrict__ anchors,
                                                              float4* __restrict__ boxes_c, float* __restrict__ scores_c,
                                                              int* __restrict__ anchor_c,
                                                              unsigned long long* __restrict__ invalid, int invalid_words) {
  const int img = blockIdx.y;
  const int k = s.k;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= k) return;
  const unsigned long long* in = cand + (size_t)img * k;
  const unsigned long long key = in[e];
  const int own = e / kSortChunk;
  const int nchunks = ceil_div(k, kSortChunk);
  int lo[kMaxChunks], hi[kMaxChunks];
#pragma unroll
  for (int c = 0; c < kMaxChunks; c++) {
    lo[c] = 0;
    hi[c] = (c < nchunks && c != own) ? min(kSortChunk, k - c * kSortChunk) : 0;
  }
#pragma unroll 1
  for (int step = 0; step < 12; step++) {  // 2^11 = chunk size; one more step to settle lo == hi
#pragma unroll
    for (int c = 0; c < kMaxChunks; c++) {
      if (lo[c] < hi[c]) {
        const int mid = (lo[c] + hi[c]) >> 1;
        if (in[c * kSortChunk + mid] > key) lo[c] = mid + 1;
        else hi[c] = mid;
      }
    }
  }
  int rank = e - own * kSortChunk;
#pragma unroll
  for (int c = 0; c < kMaxChunks; c++) rank += lo[c];

  const int anchor = (int)(((1u << kRpnIdxBits) - 1u) - (unsigned)(key & ((1u << kRpnIdxBits) - 1u)));
  const float logit = from_ordered_bits((unsigned)(key >> kRpnIdxBits));
  const int hw = s.H * s.W;
  const int M = s.A * hw;
  const float* reg = box_regression + (size_t)img * 4 * M;
  const float* anc = anchors + (size_t)(first_image + img) * d.anchor_image_stride;
  float r0, r1, r2, r3;
  if (s.layout == ABR_NHWC) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(reg) + anchor);
    r0 = r.x; r1 = r.y; r2 = r.z; r3 = r.w;
  } else {
    const int p = anchor / s.A, a = anchor - p * s.A;
    const float* rp = reg + (size_t)(a * 4) * hw + p;
    r0 = __ldg(rp); r1 = __ldg(rp + hw); r2 = __ldg(rp + 2 * hw); r3 = __ldg(rp + 3 * hw);
  }
  const float4 an = __ldg(reinterpret_cast<const float4*>(anc) + anchor);
  BoxCoderParams bc;
  bc.wx = d.wx; bc.wy = d.wy; bc.ww = d.ww; bc.wh = d.wh; bc.clip = d.clip;
  const float4 box = decode_and_clip(an, r0, r1, r2, r3, bc, (float)(d.im_w[img] - 1), (float)(d.im_h[img] - 1));
  const float x1 = box.x, y1 = box.y, x2 = box.z, y2 = box.w;
  // remove_small_boxes on the xywh sides (+1 convention)
  const float bw = __fadd_rn(__fsub_rn(x2, x1), 1.f), bh = __fadd_rn(__fsub_rn(y2, y1), 1.f);
  const bool ok = bw >= d.min_size && bh >= d.min_size;
  const size_t o = (size_t)img * k + rank;
  boxes_c[o] = box;
  scores_c[o] = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-logit)));
  anchor_c[o] = anchor;
  if (!ok) atomicOr(invalid + (size_t)(first_image + img) * invalid_words + (rank >> 6), 1ull << (rank & 63));
}

// Survivors -> padded outputs.  With keep: the NMS result (indices into the ranked candidates, ascending = the
// reference's order).  Without (nms_thresh <= 0, the reference returns the list untouched): every valid candidate in
// rank order (ordered compaction over the validity bitmap).
__global__ void __launch_bounds__(256) rpn_gather_kernel(int k, int first_image, const float4* __restrict__ boxes_c,
                                                         const float* __restrict__ scores_c, const int* __restrict__ anchor_c,
                                                         const unsigned long long* __restrict__ invalid, int invalid_words,
                                                         const long long* __restrict__ keep, int keep_stride,
                                                         const int* __restrict__ n_keep, float4* __restrict__ proposals,
                                                         float* __restrict__ scores, int* __restrict__ anchor_index,
                                                         int* __restrict__ n_out, int out_stride) {
  __shared__ int warp_tot[8];
  __shared__ int total;
  const int img = blockIdx.x, g = first_image + img, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float4* b = boxes_c + (size_t)img * k;
  const float* sc = scores_c + (size_t)img * k;
  const int* an = anchor_c + (size_t)img * k;
  float4* op = proposals + (size_t)g * out_stride;
  float* os = scores + (size_t)g * out_stride;
  int* oa = anchor_index ? anchor_index + (size_t)g * out_stride : nullptr;
  int n;
  if (keep) {
    n = n_keep[g];
    for (int j = tid; j < n; j += blockDim.x) {
      const int src = (int)keep[(size_t)g * keep_stride + j];
      op[j] = b[src]; os[j] = sc[src];
      if (oa) oa[j] = an[src];
    }
  } else {
    const unsigned long long* inv = invalid + (size_t)g * invalid_words;
    int running = 0;
    for (int w0 = 0; w0 < invalid_words; w0 += 256) {  // one 64-candidate word per thread
      const int w = w0 + tid;
      unsigned long long bits = 0;
      if (w < invalid_words) {
        const int rem = k - w * 64;
        const unsigned long long live = rem >= 64 ? ~0ull : (rem > 0 ? (1ull << rem) - 1ull : 0ull);
        bits = ~inv[w] & live;
      }
      const int cnt = __popcll(bits);
      int incl = cnt;
#pragma unroll
      for (int o2 = 1; o2 < 32; o2 <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o2);
        if (lane >= o2) incl += t;
      }
      if (lane == 31) warp_tot[warp] = incl;
      __syncthreads();
      int pos = running + incl - cnt;
      for (int w2 = 0; w2 < warp; w2++) pos += warp_tot[w2];
      while (bits) {
        const int bit = __ffsll((long long)bits) - 1;
        bits &= bits - 1;
        const int src = w * 64 + bit;
        if (pos < out_stride) {
          op[pos] = b[src]; os[pos] = sc[src];
          if (oa) oa[pos] = an[src];
        }
        pos++;
      }
      for (int w2 = 0; w2 < 8; w2++) running += warp_tot[w2];
      __syncthreads();
    }
    if (tid == 0) total = min(running, out_stride);
    __syncthreads();
    n = total;
  }
  for (int j = n + tid; j < out_stride; j += blockDim.x) {
    op[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    os[j] = 0.f;
    if (oa) oa[j] = -1;
  }
  if (tid == 0) n_out[g] = n;
}

struct RpnLayout {
  size_t hist, cand_count, invalid, zero_end, n_keep, cand, boxes, scores, anchor, keep, nms, total;
  size_t nms_bytes;
  int k, keep_stride, invalid_words;
};

static size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }

static RpnLayout rpn_layout(int N, int A, int H, int W, int pre_nms_top_n, int post_nms_top_n) {
  RpnLayout l;
  const long long M = (long long)A * H * W;
  l.k = (int)(pre_nms_top_n < M ? pre_nms_top_n : M);
  l.keep_stride = post_nms_top_n > 0 && post_nms_top_n < l.k ? post_nms_top_n : l.k;
  size_t o = 0;
  l.hist = o; o += a256((size_t)N * kRpnPasses * kRpnBins * sizeof(unsigned));
  l.cand_count = o; o += a256((size_t)N * sizeof(int));
  l.invalid_words = ceil_div(l.k, 64);
  l.invalid = o; o += a256((size_t)N * l.invalid_words * 8);
  l.zero_end = o;  // [hist, zero_end) is cleared by one memset per call
  l.n_keep = o; o += a256((size_t)N * sizeof(int));
  l.cand = o; o += a256((size_t)N * l.k * 8);
  l.boxes = o; o += a256((size_t)N * l.k * 16);
  l.scores = o; o += a256((size_t)N * l.k * 4);
  l.anchor = o; o += a256((size_t)N * l.k * 4);
  l.keep = o; o += a256((size_t)N * l.keep_stride * 8);
  // NMS scratch for N images of capacity k
  std::vector<int> offsets((size_t)N + 1);
  for (int i = 0; i <= N; i++) offsets[i] = i * l.k;
  l.nms_bytes = a256(abr_nms_workspace_bytes(offsets.data(), N));
  l.nms = o; o += l.nms_bytes;
  l.total = o;
  return l;
}

}  // namespace abr

using namespace abr;

extern "C" {

size_t abr_rpn_proposals_workspace_bytes(int N, int A, int H, int W, int pre_nms_top_n, int post_nms_top_n) {
  if (N <= 0 || A <= 0 || H <= 0 || W <= 0 || pre_nms_top_n <= 0) return 0;
  return rpn_layout(N, A, H, W, pre_nms_top_n, post_nms_top_n).total;
}

int abr_rpn_proposals(const float* objectness, const float* box_regression, const float* anchors,
                      long long anchor_image_stride, const int* image_sizes_host, int N, int A, int H, int W, int layout,
                      int pre_nms_top_n, int post_nms_top_n, float nms_thresh, int ge, float min_size,
                      const float* weights4_host, float bbox_xform_clip, float* proposals, float* scores,
                      int32_t* anchor_index, int32_t* n_out, int out_stride, void* workspace, size_t workspace_bytes,
                      abr_stream_t stream) {
  ABR_REQUIRE(N >= 0, ABR_ERR_BAD_ARG, "rpn: N=%d", N);
  if (N == 0) return ABR_OK;
  ABR_REQUIRE(A > 0 && H > 0 && W > 0 && pre_nms_top_n > 0, ABR_ERR_BAD_ARG, "rpn: bad sizes A=%d H=%d W=%d pre_nms_top_n=%d", A, H, W, pre_nms_top_n);
  ABR_REQUIRE((long long)A * H * W < (1ll << kRpnIdxBits), ABR_ERR_UNSUPPORTED, "rpn: %lld anchors per image (max %d)", (long long)A * H * W, (1 << kRpnIdxBits) - 1);
  ABR_REQUIRE(layout == ABR_NCHW || layout == ABR_NHWC, ABR_ERR_UNSUPPORTED, "rpn: layout %d not supported", layout);
  ABR_REQUIRE(objectness && box_regression && anchors && image_sizes_host && weights4_host && proposals && scores && n_out,
              ABR_ERR_BAD_ARG, "rpn: null pointer");
  ABR_REQUIRE(((reinterpret_cast<uintptr_t>(anchors) | reinterpret_cast<uintptr_t>(proposals)) & 15) == 0 && anchor_image_stride % 4 == 0,
              ABR_ERR_BAD_ARG, "rpn: anchors / proposals must be 16-byte aligned");
  if (layout == ABR_NHWC)
    ABR_REQUIRE((reinterpret_cast<uintptr_t>(box_regression) & 15) == 0, ABR_ERR_BAD_ARG, "rpn: channels-last box_regression must be 16-byte aligned");
  const RpnLayout lay = rpn_layout(N, A, H, W, pre_nms_top_n, post_nms_top_n);
  ABR_REQUIRE(lay.k <= kRpnMaxCand, ABR_ERR_UNSUPPORTED, "rpn: pre_nms_top_n %d > %d", lay.k, kRpnMaxCand);
  const int need_stride = nms_thresh > 0.f ? lay.keep_stride : lay.k;
  ABR_REQUIRE(out_stride >= need_stride, ABR_ERR_BAD_ARG, "rpn: out_stride %d < %d", out_stride, need_stride);
  ABR_REQUIRE(workspace && workspace_bytes >= lay.total && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, ABR_ERR_WORKSPACE,
              "rpn: workspace %zu B < %zu B (or not 256-byte aligned)", workspace_bytes, lay.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  unsigned* hist = reinterpret_cast<unsigned*>(ws + lay.hist);
  int* cand_count = reinterpret_cast<int*>(ws + lay.cand_count);
  unsigned long long* invalid = reinterpret_cast<unsigned long long*>(ws + lay.invalid);
  int* n_keep = reinterpret_cast<int*>(ws + lay.n_keep);
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(ws + lay.cand);
  float4* boxes_c = reinterpret_cast<float4*>(ws + lay.boxes);
  float* scores_c = reinterpret_cast<float*>(ws + lay.scores);
  int* anchor_c = reinterpret_cast<int*>(ws + lay.anchor);
  long long* keep = reinterpret_cast<long long*>(ws + lay.keep);
  ABR_CUDA_OK(cudaMemsetAsync(ws + lay.hist, 0, lay.zero_end - lay.hist, st));  // histograms, counters, validity bitmap

  const int M = A * H * W;
  const int k = lay.k;
  const int chunks = ceil_div(M, kRpnChunk);
  for (int base = 0; base < N; base += kRpnMaxImages) {
    const int n = N - base < kRpnMaxImages ? N - base : kRpnMaxImages;
    RpnShape s;
    s.N = n; s.A = A; s.H = H; s.W = W; s.layout = layout; s.k = k;
    RpnDecode d;
    d.wx = weights4_host[0]; d.wy = weights4_host[1]; d.ww = weights4_host[2]; d.wh = weights4_host[3];
    d.clip = bbox_xform_clip; d.min_size = min_size; d.anchor_image_stride = anchor_image_stride;
    for (int i = 0; i < n; i++) {
      d.im_w[i] = image_sizes_host[2 * (base + i)];
      d.im_h[i] = image_sizes_host[2 * (base + i) + 1];
    }
    const float* obj = objectness + (size_t)base * M;
    const float* reg = box_regression + (size_t)base * 4 * M;
    unsigned* h = hist + (size_t)base * kRpnPasses * kRpnBins;
    for (int pass = 0; pass < kRpnPasses; pass++) {
      rpn_hist_kernel<<<dim3(chunks, n), 256, 0, st>>>(s, obj, h, pass);
      ABR_CHECK_LAUNCH("rpn_hist");
    }
    rpn_collect_kernel<<<dim3(chunks, n), 256, 0, st>>>(s, obj, h, cand + (size_t)base * k, cand_count + base);
    ABR_CHECK_LAUNCH("rpn_collect");
    rpn_chunk_sort_kernel<<<dim3(ceil_div(k, kSortChunk), n), kSortThreads, 0, st>>>(k, cand + (size_t)base * k);
    ABR_CHECK_LAUNCH("rpn_chunk_sort");
    rpn_rank_decode_kernel<<<dim3(ceil_div(k, 256), n), 256, 0, st>>>(s, d, base, cand + (size_t)base * k, reg, anchors,
                                                                     boxes_c + (size_t)base * k, scores_c + (size_t)base * k,
                                                                     anchor_c + (size_t)base * k, invalid, lay.invalid_words);
    ABR_CHECK_LAUNCH("rpn_rank_decode");
  }
  const bool run_nms = nms_thresh > 0.f;  // structures/boxlist_ops.py:22-23: otherwise the list is returned untouched
  if (run_nms) {
    std::vector<int> offsets((size_t)N + 1);
    for (int i = 0; i <= N; i++) offsets[i] = i * k;
    int rc = nms_run(reinterpret_cast<const float*>(boxes_c), scores_c, offsets.data(), nullptr, invalid, N, nms_thresh, ge,
                     post_nms_top_n > 0 ? post_nms_top_n : -1, reinterpret_cast<int64_t*>(keep), lay.keep_stride, n_keep,
                     ws + lay.nms, lay.nms_bytes, st);
    if (rc) return rc;
  }
  for (int base = 0; base < N; base += kRpnMaxImages) {
    const int n = N - base < kRpnMaxImages ? N - base : kRpnMaxImages;
    rpn_gather_kernel<<<n, 256, 0, st>>>(k, base, boxes_c + (size_t)base * k, scores_c + (size_t)base * k,
                                         anchor_c + (size_t)base * k, invalid, lay.invalid_words, run_nms ? keep : nullptr,
                                         lay.keep_stride, n_keep,
                                         reinterpret_cast<float4*>(proposals), scores, anchor_index, n_out, out_stride);
    ABR_CHECK_LAUNCH("rpn_gather");
  }
  return ABR_OK;
}

}  // extern "C"
