// nms.cu -- batched greedy IoU NMS for sm_100a, entirely on the device (no D2H, no host sweep).
//
// Semantics: maskrcnn_benchmark/csrc/cuda/nms.cu:13-131 of the reference (+1 pixel convention, IoU > thr, greedy in
// score order, result = surviving ORIGINAL indices ascending) -- `ge` selects csrc/cpu/nms_cpu.cpp:60 (IoU >= thr).
// Design (not a port).  The reference sorts with a library call, fills the full N x N/64 bitmask (the lower triangle
// is never read), copies 4.5-18 MB to the host with a blocking cudaMemcpy and sweeps it serially on the CPU, once per
// image from a Python loop.  Here a whole batch of images is three launches on the caller's stream:
//   1. rank_sort_kernel   -- order = stable descending sort of the scores by all-pairs rank counting on 64-bit
//                            (ordered score bits, ~index) keys: O(N^2) like the IoU stage, but embarrassingly
//                            parallel, deterministic, and no temporary storage or library call;
//   2. iou_mask_kernel    -- 64x64 tiles of the UPPER triangle only; IoU in explicitly rounded fp32
//                            (__fmul_rn/__fadd_rn/__fdiv_rn: bit-exact with the reference's source-order arithmetic,
//                            immune to FMA contraction of Sa+Sb-w*h);
//   3. sweep_kernel       -- one CTA per image: warp 0 resolves each 64-box diagonal tile with a register-resident
//                            suppression word and warp shuffles, then the whole CTA ORs the mask rows of the boxes
//                            kept in that tile into the shared-memory `removed` words (coalesced 8-byte loads);
//                            finally the kept set is compacted to ascending original indices, cut to max_keep and
//                            written with its count -- the host never sees the mask.
#include "common.cuh"

namespace abr {

constexpr int kTile = 64;          // boxes per mask word
constexpr int kImagesPerLaunch = 32;

struct NmsBatch {
  int n_images;
  int box_off[kImagesPerLaunch];        // first box of the image inside boxes / scores
  int n[kImagesPerLaunch];              // boxes in the image
  long long mask_off[kImagesPerLaunch]; // first mask word of the image (u64 units)
};

// Ordered key: larger key = earlier in torch.sort(descending=True, stable=True).  NaN sorts first (torch treats NaN as
// the largest value), -0.0 ties with +0.0, equal scores keep ascending index.
__device__ __forceinline__ unsigned long long sort_key(float s, int idx) {
  unsigned int u = __float_as_uint(s);
  if (s != s) u = 0xFFFFFFFFu;
  else {
    if (u == 0x80000000u) u = 0u;
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  }
  return ((unsigned long long)u << 32) | (unsigned int)(~(unsigned int)idx);
}

constexpr int kRankThreads = 256;
constexpr int kRankPerThread = 4;

__global__ void __launch_bounds__(kRankThreads) rank_sort_kernel(NmsBatch nb, const float* __restrict__ boxes,
                                                                const float* __restrict__ scores,
                                                                float4* __restrict__ sorted_boxes,
                                                                int* __restrict__ order) {
  const int img = blockIdx.y;
  const int n = nb.n[img], off = nb.box_off[img];
  const int base = blockIdx.x * (kRankThreads * kRankPerThread);
  if (base >= n) return;
  __shared__ unsigned long long tile[kRankThreads * 2];
  unsigned long long mine[kRankPerThread];
  int rank[kRankPerThread];
#pragma unroll
  for (int k = 0; k < kRankPerThread; k++) {
    const int i = base + k * kRankThreads + threadIdx.x;
    mine[k] = i < n ? sort_key(scores[off + i], i) : 0ull;
    rank[k] = 0;
  }
  for (int j0 = 0; j0 < n; j0 += kRankThreads * 2) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const int j = j0 + k * kRankThreads + threadIdx.x;
      // key 0 is smaller than every real key (a real key has a nonzero low word unless idx == 0xFFFFFFFF)
      tile[k * kRankThreads + threadIdx.x] = j < n ? sort_key(scores[off + j], j) : 0ull;
    }
    __syncthreads();
    const int lim = min(kRankThreads * 2, n - j0);
#pragma unroll 8
    for (int j = 0; j < lim; j++) {
      const unsigned long long kj = tile[j];
#pragma unroll
      for (int k = 0; k < kRankPerThread; k++) rank[k] += (kj > mine[k]) ? 1 : 0;
    }
  }
#pragma unroll
  for (int k = 0; k < kRankPerThread; k++) {
    const int i = base + k * kRankThreads + threadIdx.x;
    if (i < n) {
      order[off + rank[k]] = i;
      sorted_boxes[off + rank[k]] = __ldg(reinterpret_cast<const float4*>(boxes) + off + i);
    }
  }
}

// devIoU of csrc/cuda/nms.cu:13-21 with every operation individually rounded (no contraction).
__device__ __forceinline__ float iou_plus_one(const float4 a, const float4 b) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float width = fmaxf(__fadd_rn(__fsub_rn(right, left), 1.f), 0.f);
  const float height = fmaxf(__fadd_rn(__fsub_rn(bottom, top), 1.f), 0.f);
  const float inter = __fmul_rn(width, height);
  const float sa = __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.f), __fadd_rn(__fsub_rn(a.w, a.y), 1.f));
  const float sb = __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter));
}

// One CTA of 64 threads per (column tile, row tile) with column >= row.  Thread t owns row box 64*row+t and emits the
// 64-bit word of column boxes it suppresses (only boxes AFTER it in score order: csrc/cuda/nms.cu:53-65).
__global__ void __launch_bounds__(kTile) iou_mask_kernel(NmsBatch nb, const float4* __restrict__ sorted_boxes,
                                                        unsigned long long* __restrict__ mask, float thresh, int ge) {
  const int img = blockIdx.z;
  const int n = nb.n[img];
  const int cb = ceil_div(n, kTile);
  const int row = blockIdx.y, col = blockIdx.x;
  if (row >= cb || col >= cb || col < row) return;
  const float4* bx = sorted_boxes + nb.box_off[img];
  __shared__ float4 cbox[kTile];
  const int col_size = min(n - col * kTile, kTile), row_size = min(n - row * kTile, kTile);
  if ((int)threadIdx.x < col_size) cbox[threadIdx.x] = bx[col * kTile + threadIdx.x];
  __syncthreads();
  if ((int)threadIdx.x < row_size) {
    const int cur = row * kTile + threadIdx.x;
    const float4 a = bx[cur];
    unsigned long long t = 0;
    const int start = (row == col) ? threadIdx.x + 1 : 0;
    for (int i = start; i < col_size; i++) {
      const float v = iou_plus_one(a, cbox[i]);
      if (ge ? (v >= thresh) : (v > thresh)) t |= 1ull << i;
    }
    mask[nb.mask_off[img] + (long long)cur * cb + col] = t;
  }
}

constexpr int kSweepThreads = 512;

// One CTA per image.  Dynamic shared memory: removed[cb], kept_sorted[cb], kept_orig[cb] (u64 each) + scan scratch.
__global__ void __launch_bounds__(kSweepThreads) sweep_kernel(NmsBatch nb, const unsigned long long* __restrict__ mask,
                                                             const int* __restrict__ order, int max_keep,
                                                             long long* __restrict__ keep, int keep_stride,
                                                             int* __restrict__ n_keep, int image_base) {
  extern __shared__ unsigned long long sm[];
  const int img = blockIdx.x;
  const int n = nb.n[img];
  const int cb = ceil_div(n, kTile);
  unsigned long long* removed = sm;
  unsigned long long* kept_sorted = removed + cb;
  unsigned long long* kept_orig = kept_sorted + cb;
  int* scan = reinterpret_cast<int*>(kept_orig + cb);  // [kSweepThreads/32 + 1]
  __shared__ unsigned long long kept_now;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  long long* out = keep + (long long)(image_base + img) * keep_stride;

  for (int i = tid; i < cb; i += kSweepThreads) { removed[i] = 0; kept_orig[i] = 0; kept_sorted[i] = 0; }
  __syncthreads();
  const unsigned long long* m = mask + nb.mask_off[img];
  const int* ord = order + nb.box_off[img];

  for (int k = 0; k < cb; k++) {
    if (warp == 0) {
      const int first = k * kTile;
      const int size = min(n - first, kTile);
      // diagonal tile: lane l holds the words of boxes first+l and first+l+32
      const unsigned long long d0 = lane < size ? m[(long long)(first + lane) * cb + k] : 0ull;
      const unsigned long long d1 = lane + 32 < size ? m[(long long)(first + lane + 32) * cb + k] : 0ull;
      const unsigned long long valid = size == kTile ? ~0ull : ((1ull << size) - 1ull);
      unsigned long long alive = ~removed[k] & valid;
      unsigned long long kept = 0;
      while (alive) {  // greedy chain inside the tile (csrc/cuda/nms.cu:112-123)
        const int b = __ffsll((long long)alive) - 1;
        kept |= 1ull << b;
        const unsigned long long lo = __shfl_sync(0xffffffffu, d0, b & 31);
        const unsigned long long hi = __shfl_sync(0xffffffffu, d1, b & 31);
        const unsigned long long d = (b < 32) ? lo : hi;
        alive &= ~d;
        alive &= ~(1ull << b);
      }
      if (lane == 0) { kept_now = kept; kept_sorted[k] = kept; }
    }
    __syncthreads();
    const unsigned long long kept = kept_now;
    // OR the rows of the kept boxes into the later `removed` words; thread t owns word k+1+t, k+1+t+T, ...
    for (int j = k + 1 + tid; j < cb; j += kSweepThreads) {
      unsigned long long acc = 0;
      unsigned long long bits = kept;
      while (bits) {
        const int b = __ffsll((long long)bits) - 1;
        bits &= bits - 1;
        acc |= m[(long long)(k * kTile + b) * cb + j];
      }
      removed[j] |= acc;
    }
    __syncthreads();
  }

  // kept (sorted positions) -> bitmap over original indices
  for (int w = tid; w < cb; w += kSweepThreads) {
    unsigned long long bits = kept_sorted[w];
    while (bits) {
      const int b = __ffsll((long long)bits) - 1;
      bits &= bits - 1;
      const int orig = ord[w * kTile + b];
      atomicOr(&kept_orig[orig >> 6], 1ull << (orig & 63));
    }
  }
  __syncthreads();
  // ascending original indices: block-wide exclusive scan of per-word popcounts, chunk by chunk
  int running = 0;
  for (int w0 = 0; w0 < cb; w0 += kSweepThreads) {
    const int w = w0 + tid;
    const unsigned long long bits = w < cb ? kept_orig[w] : 0ull;
    const int cnt = __popcll(bits);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) scan[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int v = lane < kSweepThreads / 32 ? scan[lane] : 0;
      int s = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
      }
      if (lane < kSweepThreads / 32) scan[lane] = s - v;
      if (lane == kSweepThreads / 32 - 1) scan[kSweepThreads / 32] = s;
    }
    __syncthreads();
    int pos = running + scan[warp] + incl - cnt;
    unsigned long long b2 = bits;
    while (b2) {
      const int b = __ffsll((long long)b2) - 1;
      b2 &= b2 - 1;
      if (pos < keep_stride && (max_keep <= 0 || pos < max_keep)) out[pos] = (long long)(w * kTile + b);
      pos++;
    }
    running += scan[kSweepThreads / 32];
    __syncthreads();
  }
  int total = running;
  if (max_keep > 0) total = min(total, max_keep);
  total = min(total, keep_stride);
  for (int i = total + tid; i < keep_stride; i += kSweepThreads) out[i] = -1;
  if (tid == 0) n_keep[image_base + img] = total;
}

__global__ void nms_fill_empty_kernel(long long* keep, int keep_stride, int* n_keep, int image) {
  for (int i = threadIdx.x; i < keep_stride; i += blockDim.x) keep[(long long)image * keep_stride + i] = -1;
  if (threadIdx.x == 0) n_keep[image] = 0;
}

struct NmsLayout {
  size_t sorted_boxes, order, mask, total;
};
static NmsLayout nms_layout(const int* offsets, int n_images) {
  NmsLayout l;
  const size_t total_boxes = (size_t)offsets[n_images];
  size_t words = 0;
  for (int i = 0; i < n_images; i++) {
    const size_t n = (size_t)(offsets[i + 1] - offsets[i]);
    words += n * ceil_div<size_t>(n, kTile);
  }
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  l.sorted_boxes = 0;
  l.order = align(total_boxes * sizeof(float4));
  l.mask = l.order + align(total_boxes * sizeof(int));
  l.total = l.mask + align(words * 8);
  return l;
}

}  // namespace abr

using namespace abr;

extern "C" {

size_t abr_nms_workspace_bytes(const int* offsets_host, int n_images) {
  if (!offsets_host || n_images <= 0) return 0;
  return nms_layout(offsets_host, n_images).total;
}

int abr_nms_batched(const float* boxes, const float* scores, const int* offsets_host, int n_images, float thresh, int ge,
                    int max_keep, int64_t* keep, int keep_stride, int32_t* n_keep, void* workspace,
                    size_t workspace_bytes, abr_stream_t stream) {
  ABR_REQUIRE(n_images >= 0, ABR_ERR_BAD_ARG, "nms: n_images=%d", n_images);
  if (n_images == 0) return ABR_OK;
  ABR_REQUIRE(offsets_host && keep && n_keep && keep_stride >= 0, ABR_ERR_BAD_ARG, "nms: null pointer or negative stride");
  ABR_REQUIRE(offsets_host[0] == 0, ABR_ERR_BAD_ARG, "nms: offsets must start at 0");
  for (int i = 0; i < n_images; i++) {
    const int n = offsets_host[i + 1] - offsets_host[i];
    ABR_REQUIRE(n >= 0, ABR_ERR_BAD_ARG, "nms: offsets must be non-decreasing");
    const int need = max_keep > 0 ? (n < max_keep ? n : max_keep) : n;
    ABR_REQUIRE(keep_stride >= need, ABR_ERR_BAD_ARG, "nms: keep_stride %d < %d needed by image %d", keep_stride, need, i);
  }
  const int total = offsets_host[n_images];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (total > 0) ABR_REQUIRE(boxes && scores, ABR_ERR_BAD_ARG, "nms: null boxes/scores");
  if (total > 0 && (reinterpret_cast<uintptr_t>(boxes) & 15))
    ABR_REQUIRE(false, ABR_ERR_BAD_ARG, "nms: boxes must be 16-byte aligned");
  const NmsLayout lay = nms_layout(offsets_host, n_images);
  if (total > 0)
    ABR_REQUIRE(workspace && workspace_bytes >= lay.total, ABR_ERR_WORKSPACE, "nms: workspace %zu B < %zu B", workspace_bytes, lay.total);
  char* ws = static_cast<char*>(workspace);
  float4* sorted_boxes = reinterpret_cast<float4*>(ws + lay.sorted_boxes);
  int* order = reinterpret_cast<int*>(ws + lay.order);
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(ws + lay.mask);

  long long mask_cursor = 0;
  for (int base = 0; base < n_images; base += kImagesPerLaunch) {
    NmsBatch nb;
    nb.n_images = 0;
    int nmax = 0;
    const int lim = n_images - base < kImagesPerLaunch ? n_images - base : kImagesPerLaunch;
    for (int i = 0; i < lim; i++) {
      const int n = offsets_host[base + i + 1] - offsets_host[base + i];
      nb.box_off[i] = offsets_host[base + i];
      nb.n[i] = n;
      nb.mask_off[i] = mask_cursor;
      mask_cursor += (long long)n * ceil_div(n, kTile);
      nmax = n > nmax ? n : nmax;
    }
    nb.n_images = lim;
    if (nmax == 0) {
      for (int i = 0; i < lim; i++) {
        nms_fill_empty_kernel<<<1, 128, 0, st>>>(reinterpret_cast<long long*>(keep), keep_stride, n_keep, base + i);
        ABR_CHECK_LAUNCH("nms_fill_empty");
      }
      continue;
    }
    const int cbmax = ceil_div(nmax, kTile);
    ABR_REQUIRE(cbmax <= 65535, ABR_ERR_UNSUPPORTED, "nms: %d boxes in one image (max %d)", nmax, 65535 * kTile);
    rank_sort_kernel<<<dim3(ceil_div(nmax, kRankThreads * kRankPerThread), lim), kRankThreads, 0, st>>>(nb, boxes, scores, sorted_boxes, order);
    ABR_CHECK_LAUNCH("nms_rank_sort");
    iou_mask_kernel<<<dim3(cbmax, cbmax, lim), kTile, 0, st>>>(nb, sorted_boxes, mask, thresh, ge);
    ABR_CHECK_LAUNCH("nms_iou_mask");
    const size_t smem = (size_t)cbmax * 3 * 8 + (kSweepThreads / 32 + 1) * sizeof(int);
    ABR_REQUIRE(smem <= 200 * 1024, ABR_ERR_UNSUPPORTED, "nms: %d boxes in one image need %zu B of shared memory", nmax, smem);
    if (smem > 48 * 1024) ABR_CUDA_OK(cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sweep_kernel<<<lim, kSweepThreads, smem, st>>>(nb, mask, order, max_keep, reinterpret_cast<long long*>(keep), keep_stride, n_keep, base);
    ABR_CHECK_LAUNCH("nms_sweep");
  }
  return ABR_OK;
}

}  // extern "C"
