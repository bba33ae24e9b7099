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
constexpr int kImagesPerLaunch = 64;

struct NmsBatch {
  int n_images;
  int first_image;                      // index of the launch group's first image in the whole call
  int box_off[kImagesPerLaunch];        // first box of the image inside boxes / scores
  int n[kImagesPerLaunch];              // boxes of the image this pass looks at (a prefix in the prefix pass)
  int n_full[kImagesPerLaunch];         // boxes in the image
  long long mask_off[kImagesPerLaunch]; // first mask word of the image (u64 units)
  const int* counts;                    // optional, device: boxes actually present per image (<= the host-side size)
  const unsigned long long* invalid;    // optional, device: bit i of image's words = box i (INPUT order) never takes part
  long long invalid_off[kImagesPerLaunch];
};

// boxes of image `img` a kernel looks at: the host-side figure, cut to the device-side count when there is one
__device__ __forceinline__ int live_boxes(const NmsBatch& nb, int img, int host_n) {
  return nb.counts ? min(host_n, __ldg(nb.counts + nb.first_image + img)) : host_n;
}

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

// O(N) first step: assume the scores already are in stable descending order (true for the RPN, which feeds NMS the
// output of a sorted top-k: modeling/rpn/inference.py:94-95) -- write the identity order and a copy of the boxes, and
// raise unsorted[img] on the first adjacent pair that contradicts it.  rank_sort_kernel then only runs for unsorted
// images.
__global__ void presort_kernel(NmsBatch nb, const float* __restrict__ boxes, const float* __restrict__ scores,
                               float4* __restrict__ sorted_boxes, int* __restrict__ order, int* __restrict__ unsorted) {
  const int img = blockIdx.y;
  const int n = live_boxes(nb, img, nb.n_full[img]), off = nb.box_off[img];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  order[off + i] = i;
  sorted_boxes[off + i] = __ldg(reinterpret_cast<const float4*>(boxes) + off + i);
  if (i + 1 < n && !(sort_key(scores[off + i], i) > sort_key(scores[off + i + 1], i + 1))) unsorted[nb.first_image + img] = 1;
}

constexpr int kRankThreads = 128;  // candidates per CTA: small, so that even one image fills the machine

// Stable descending sort by all-pairs rank counting (only for images presort_kernel flagged as unsorted).
__global__ void __launch_bounds__(kRankThreads) rank_sort_kernel(NmsBatch nb, const float* __restrict__ boxes,
                                                                const float* __restrict__ scores,
                                                                float4* __restrict__ sorted_boxes,
                                                                int* __restrict__ order, int* __restrict__ unsorted) {
  const int img = blockIdx.y;
  const int n = live_boxes(nb, img, nb.n_full[img]), off = nb.box_off[img];
  const int base = blockIdx.x * kRankThreads;
  if (base >= n || unsorted[nb.first_image + img] == 0) return;  // presort_kernel found the image already sorted
  __shared__ unsigned long long tile[512];
  const int i = base + threadIdx.x;
  const unsigned long long mine = i < n ? sort_key(scores[off + i], i) : 0ull;
  int rank = 0;
  for (int j0 = 0; j0 < n; j0 += 512) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 512 / kRankThreads; k++) {
      const int j = j0 + k * kRankThreads + threadIdx.x;
      // key 0 is smaller than every real key (a real key's high word is at least 0x007fffff)
      tile[k * kRankThreads + threadIdx.x] = j < n ? sort_key(scores[off + j], j) : 0ull;
    }
    __syncthreads();
    const int lim = min(512, n - j0);
    int r0 = 0, r1 = 0, r2 = 0, r3 = 0;
    int j = 0;
    for (; j + 3 < lim; j += 4) {
      r0 += tile[j] > mine;
      r1 += tile[j + 1] > mine;
      r2 += tile[j + 2] > mine;
      r3 += tile[j + 3] > mine;
    }
    for (; j < lim; j++) r0 += tile[j] > mine;
    rank += r0 + r1 + r2 + r3;
  }
  if (i < n) {
    order[off + rank] = i;
    sorted_boxes[off + rank] = __ldg(reinterpret_cast<const float4*>(boxes) + off + i);
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

__device__ __forceinline__ float box_area_plus_one(const float4 b) {
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
}

// Upper-triangular tile pairs only, enumerated row-major: tile t of an image with cb tile rows is (row, col >= row).
// A 256-thread CTA takes kMaskTilesPerCta consecutive pairs, one per 64-thread group; thread i of a group owns row box
// 64*row+i and emits the 64-bit word of column boxes it suppresses (only boxes AFTER it in score order:
// csrc/cuda/nms.cu:53-65).
constexpr int kMaskTilesPerCta = 4;

__global__ void __launch_bounds__(kTile * kMaskTilesPerCta) iou_mask_kernel(NmsBatch nb, const float4* __restrict__ sorted_boxes,
                                                                          unsigned long long* __restrict__ mask,
                                                                          float thresh, int ge, const int* __restrict__ done) {
  const int img = blockIdx.y;
  if (done && done[nb.first_image + img]) return;  // full pass: settled by the prefix pass; prefix pass: image not sorted
  const int n = live_boxes(nb, img, nb.n[img]);
  const int cb = ceil_div(n, kTile);
  const long long ntiles = (long long)cb * (cb + 1) / 2;
  const int grp = threadIdx.x / kTile, lane64 = threadIdx.x % kTile;
  const long long t = (long long)blockIdx.x * kMaskTilesPerCta + grp;
  __shared__ float4 cbox[kMaskTilesPerCta][kTile];
  __shared__ float carea[kMaskTilesPerCta][kTile];
  int row = 0, col = 0;
  const bool live = t < ntiles;
  if (live) {
    // row r starts at tile index r*cb - r*(r-1)/2; invert with a float guess and fix up
    const double fc = (double)cb + 0.5;
    row = (int)(fc - sqrt(fc * fc - 2.0 * (double)t));
    if (row < 0) row = 0;
    if (row >= cb) row = cb - 1;
    while (row > 0 && (long long)row * cb - (long long)row * (row - 1) / 2 > t) row--;
    while ((long long)(row + 1) * cb - (long long)(row + 1) * row / 2 <= t) row++;
    col = row + (int)(t - ((long long)row * cb - (long long)row * (row - 1) / 2));
  }
  const float4* bx = sorted_boxes + nb.box_off[img];
  const int col_size = live ? min(n - col * kTile, kTile) : 0, row_size = live ? min(n - row * kTile, kTile) : 0;
  if (lane64 < col_size) {
    const float4 b = bx[col * kTile + lane64];
    cbox[grp][lane64] = b;
    carea[grp][lane64] = box_area_plus_one(b);
  }
  __syncthreads();
  if (lane64 < row_size) {
    const int cur = row * kTile + lane64;
    const float4 a = bx[cur];
    const float sa = box_area_plus_one(a);
    unsigned long long w = 0;
    const int start = (row == col) ? lane64 + 1 : 0;
    // The decision is fl(inter / union) > thresh exactly as the reference rounds it, but the division only runs for
    // the few pairs within 2^-20 of the threshold: pairs without intersection and pairs clearly on one side are
    // settled by a product.  (thresh <= 0 or non-finite values take the exact path for every pair.)
    const bool fast = thresh >= 1e-6f;  // keeps thresh * union a normal number for every union the fast path accepts
    const float hi_k = 1.f + 9.5367431640625e-07f, lo_k = 1.f - 9.5367431640625e-07f;  // 1 +- 2^-20
#pragma unroll 4
    for (int i = start; i < col_size; i++) {
      const float4 b = cbox[grp][i];
      const float width = fmaxf(__fadd_rn(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 1.f), 0.f);
      const float height = fmaxf(__fadd_rn(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 1.f), 0.f);
      const float inter = __fmul_rn(width, height);
      const float uni = __fsub_rn(__fadd_rn(sa, carea[grp][i]), inter);
      const float p = __fmul_rn(thresh, uni);
      // branch-free for all but the pairs within 2^-20 of the threshold (a pair without intersection has inter = 0 < p)
      const bool usable = fast && uni > 1e-10f;
      bool hit = usable && inter > __fmul_rn(p, hi_k);
      const bool settled = usable && (hit || inter < __fmul_rn(p, lo_k));
      if (!settled) {
        const float v = __fdiv_rn(inter, uni);
        hit = ge ? (v >= thresh) : (v > thresh);
      }
      if (hit) w |= 1ull << i;
    }
    mask[nb.mask_off[img] + (long long)cur * cb + col] = w;
  }
}

constexpr int kSweepThreads = 512;

// OR into a 64-bit shared-memory word as two native 32-bit atomics (the 64-bit form compiles to a CAS spin loop)
__device__ __forceinline__ void or_shared_u64(unsigned long long* word, unsigned long long v) {
  unsigned* h = reinterpret_cast<unsigned*>(word);
  if ((unsigned)v) atomicOr(h, (unsigned)v);
  if ((unsigned)(v >> 32)) atomicOr(h + 1, (unsigned)(v >> 32));
}

// One CTA per image.  Per 64-box tile: warp 0 resolves the greedy chain inside the tile from the diagonal mask words
// (suppression word and chain in registers, diagonal and first off-diagonal words prefetched a tile ahead) and publishes
// the kept boxes; warps 1..15 OR the remaining mask words of the kept rows into the shared-memory `removed` words --
// (kept row, word) pairs dealt out to all helper threads, four independent 8-byte loads in flight per thread, merged
// with shared-memory atomicOr -- while warp 0 already runs the chain of the next tile (named barriers, see below).
// Sorted inputs stop as soon as max_keep boxes are kept.
// Dynamic shared memory: removed[cb], kept_sorted[cb], kept_orig[cb] (u64 each) + scan scratch.
__global__ void __launch_bounds__(kSweepThreads) sweep_kernel(NmsBatch nb, const unsigned long long* __restrict__ mask,
                                                             const int* __restrict__ order,
                                                             const int* __restrict__ unsorted, int max_keep,
                                                             long long* __restrict__ keep, int keep_stride,
                                                             int* __restrict__ n_keep, int* __restrict__ done, int prefix_pass) {
  extern __shared__ unsigned long long sm[];
  const int img = blockIdx.x;
  if (!prefix_pass && done && done[nb.first_image + img]) return;  // settled by the prefix pass
  if (prefix_pass && unsorted[nb.first_image + img]) return;       // a prefix only answers for sorted images (done stays 0)
  const int n = live_boxes(nb, img, nb.n[img]);
  const int cb = ceil_div(n, kTile);
  unsigned long long* removed = sm;
  unsigned long long* kept_sorted = removed + cb;
  unsigned long long* kept_orig = kept_sorted + cb;
  int* scan = reinterpret_cast<int*>(kept_orig + cb);  // [kSweepThreads/32 + 1]
  __shared__ int kept_rows[2][kTile];
  __shared__ int kept_count[2], stop_flag[2], kept_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  long long* out = keep + (long long)(nb.first_image + img) * keep_stride;

  // boxes flagged invalid (bitmap in INPUT order) start out suppressed
  const unsigned long long* inv = nb.invalid ? nb.invalid + nb.invalid_off[img] : nullptr;
  const bool reordered = unsorted[nb.first_image + img] != 0;
  for (int i = tid; i < cb; i += kSweepThreads) {
    unsigned long long r = 0;
    if (inv && !reordered) r = inv[i];
    else if (inv) {
      const int* ord0 = order + nb.box_off[img];
      for (int b = 0; b < kTile && i * kTile + b < n; b++) {
        const int orig = ord0[i * kTile + b];
        r |= ((inv[orig >> 6] >> (orig & 63)) & 1ull) << b;
      }
    }
    removed[i] = r; kept_orig[i] = 0; kept_sorted[i] = 0;
  }
  if (tid == 0) kept_total = 0;
  __syncthreads();
  const unsigned long long* m = mask + nb.mask_off[img];
  const int* ord = order + nb.box_off[img];
  const bool may_stop = max_keep > 0 && unsorted[nb.first_image + img] == 0;

  // Two roles, pipelined with named barriers (ids 1,2 = "kept rows of tile t published", 3,4 = "helpers finished tile
  // t"; parity t & 1).  Warp 0 runs the serial chain: for tile t it needs removed[t] = (helper ORs of tiles <= t-2) |
  // (word t of the rows kept in tile t-1); the latter comes from words it prefetched itself a tile ahead, so the chain
  // of tile t+1 starts right after the chain of tile t while warps 1..15 are still ORing the far words (columns
  // >= t+2) of tile t's kept rows into shared memory.
  if (warp == 0) {
    // lane l holds, for boxes l and l+32 of the current tile, the diagonal word and the first off-diagonal word
    unsigned long long d0 = 0, d1 = 0, f0 = 0, f1 = 0;
    if (cb > 0) {
      const int size = min(n, kTile);
      if (lane < size) { d0 = m[(long long)lane * cb]; f0 = cb > 1 ? m[(long long)lane * cb + 1] : 0ull; }
      if (lane + 32 < size) { d1 = m[(long long)(lane + 32) * cb]; f1 = cb > 1 ? m[(long long)(lane + 32) * cb + 1] : 0ull; }
    }
    unsigned long long carry = 0;  // word t of the rows kept in tile t-1
    int total = 0, published = 0;
    for (int k = 0; k < cb; k++) {
      const int first = k * kTile;
      const int size = min(n - first, kTile);
      const unsigned long long c0 = d0, c1 = d1, g0 = f0, g1 = f1;
      if (k + 1 < cb) {  // prefetch the next tile's words while this tile's chain runs
        const int nfirst = first + kTile, nsize = min(n - nfirst, kTile);
        const bool more = k + 2 < cb;
        d0 = d1 = f0 = f1 = 0ull;
        if (lane < nsize) {
          const unsigned long long* row = m + (long long)(nfirst + lane) * cb + k + 1;
          d0 = row[0];
          if (more) f0 = row[1];
        }
        if (lane + 32 < nsize) {
          const unsigned long long* row = m + (long long)(nfirst + lane + 32) * cb + k + 1;
          d1 = row[0];
          if (more) f1 = row[1];
        }
      }
      if (k >= 2) asm volatile("bar.sync %0, %1;" ::"r"(3 + (k & 1)), "r"(kSweepThreads) : "memory");  // helpers done with tile k-2
      const unsigned long long valid = size == kTile ? ~0ull : ((1ull << size) - 1ull);
      unsigned long long alive = ~(removed[k] | carry) & valid;
      // Greedy chain inside the tile (csrc/cuda/nms.cu:112-123), visiting all 64 boxes in order: box b is kept iff its
      // bit is still set when visited; its row then clears later boxes.  Rows only hold bits > b, so a kept bit stays set
      // and the final `alive` IS the kept set.  The row broadcasts do not depend on the chain (constant source lanes), so
      // the dependent path per box is a bit test and a predicated AND.
      unsigned alo = (unsigned)alive, ahi = (unsigned)(alive >> 32);
#pragma unroll
      for (int b = 0; b < 32; b++) {
        const unsigned rlo = __shfl_sync(0xffffffffu, (unsigned)c0, b);
        const unsigned rhi = __shfl_sync(0xffffffffu, (unsigned)(c0 >> 32), b);
        if ((alo >> b) & 1u) { alo &= ~rlo; ahi &= ~rhi; }
      }
#pragma unroll
      for (int b = 0; b < 32; b++) {
        const unsigned rhi = __shfl_sync(0xffffffffu, (unsigned)(c1 >> 32), b);  // rows 32..63 only reach columns > 32
        if ((ahi >> b) & 1u) ahi &= ~rhi;
      }
      const unsigned long long kept = ((unsigned long long)ahi << 32) | alo;
      const int cnt = __popcll(kept);
      int* rows = kept_rows[k & 1];
      if ((alo >> lane) & 1u) rows[__popc(alo & ((1u << lane) - 1u))] = first + lane;
      if ((ahi >> lane) & 1u) rows[__popc(alo) + __popc(ahi & ((1u << lane) - 1u))] = first + 32 + lane;
      total += cnt;
      const bool last = (k + 1 == cb) || (may_stop && total >= max_keep);  // later boxes have larger indices than the cut
      if (lane == 0) { kept_sorted[k] = kept; kept_count[k & 1] = cnt; stop_flag[k & 1] = last ? 1 : 0; }
      __syncwarp();
      asm volatile("bar.arrive %0, %1;" ::"r"(1 + (k & 1)), "r"(kSweepThreads) : "memory");  // tile k published
      published = k + 1;
      if (last) break;
      // word k+1 of the rows just kept: OR over the kept lanes' prefetched words
      unsigned long long mine = (((kept >> lane) & 1ull) ? g0 : 0ull) | (((kept >> (lane + 32)) & 1ull) ? g1 : 0ull);
      const unsigned lo32 = __reduce_or_sync(0xffffffffu, (unsigned)mine);
      const unsigned hi32 = __reduce_or_sync(0xffffffffu, (unsigned)(mine >> 32));
      carry = ((unsigned long long)hi32 << 32) | lo32;
    }
    if (lane == 0) kept_total = total;
    // helpers arrive once per tile they worked on (every published tile but the last); the loop waited for all of those
    // except the one before the last
    if (published >= 2) asm volatile("bar.sync %0, %1;" ::"r"(3 + ((published - 2) & 1)), "r"(kSweepThreads) : "memory");
  } else {
    const int htid = tid - 32;
    constexpr int nhelp = kSweepThreads - 32;           // 480 helper threads
    // Speculative window: the 60 columns after k+1.  Helper (c, g) = (htid / 8, htid % 8) owns column k+2+c for rows
    // 8g..8g+7 of the tile; the eight row groups of a column sit in adjacent lanes, so a column's OR is three shuffles
    // and one pair of 32-bit atomics by its g == 0 lane (a 64-bit shared-memory atomicOr is a CAS spin loop).
    constexpr int kNear = nhelp / 8, kSpec = kTile / 8;
    const int hc = htid >> 3, hg = htid & 7;
    for (int k = 0; k < cb; k++) {
      // fetched for ALL rows of tile k while warp 0 still runs the chain of tile k: once the kept set is published only
      // register work and one shared-memory word per column remain
      unsigned long long v[kSpec];
      const int first = k * kTile;
      const int col = k + 2 + hc;
#pragma unroll
      for (int u = 0; u < kSpec; u++) {
        const int r = first + hg * kSpec + u;
        v[u] = (r < n && col < cb) ? m[(long long)r * cb + col] : 0ull;
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + (k & 1)), "r"(kSweepThreads) : "memory");  // kept rows of tile k are published
      if (stop_flag[k & 1]) break;  // tile k was the last one: nothing further depends on its far words
      const unsigned kept8 = (unsigned)(kept_sorted[k] >> (hg * kSpec)) & 0xffu;
      unsigned long long acc = 0;
#pragma unroll
      for (int u = 0; u < kSpec; u++)
        if ((kept8 >> u) & 1u) acc |= v[u];
      acc |= __shfl_xor_sync(0xffffffffu, acc, 1);
      acc |= __shfl_xor_sync(0xffffffffu, acc, 2);
      acc |= __shfl_xor_sync(0xffffffffu, acc, 4);
      if (hg == 0 && col < cb && acc) or_shared_u64(&removed[col], acc);  // helpers of adjacent tiles may overlap here
      // Far part (columns beyond the speculative window; only long full passes have any): kept rows only, (row, word)
      // pairs dealt out to all helper threads, word index fastest, four independent loads in flight.
      const int cnt = kept_count[k & 1];
      const int* rows = kept_rows[k & 1];
      const int far0 = k + 2 + kNear;
      const int nwords = cb - far0;
      const int total = nwords > 0 ? cnt * nwords : 0;
      for (int t = htid; t < total; t += 4 * nhelp) {
        unsigned long long w[4];
        int j[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int tt = t + u * nhelp;
          if (tt < total) {
            const int ri = tt / nwords;
            j[u] = far0 + (tt - ri * nwords);
            w[u] = m[(long long)rows[ri] * cb + j[u]];
          } else {
            j[u] = -1;
            w[u] = 0;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (j[u] >= 0 && w[u]) or_shared_u64(&removed[j[u]], w[u]);
      }
      asm volatile("bar.arrive %0, %1;" ::"r"(3 + (k & 1)), "r"(kSweepThreads) : "memory");  // done with tile k
    }
  }
  __syncthreads();
  if (prefix_pass) {
    // Only a prefix of a SORTED image was examined: the answer is final iff it already holds max_keep boxes (later
    // boxes could only add indices beyond the cut); otherwise the full pass redoes the image.
    // (or when the prefix happened to cover the whole image)
    const bool final_answer = (may_stop && kept_total >= max_keep) || n >= live_boxes(nb, img, nb.n_full[img]);
    if (tid == 0) done[nb.first_image + img] = final_answer ? 1 : 0;
    if (!final_answer) return;
  }

  // kept (sorted positions) -> bitmap over original indices.  Sorted inputs: the order is the identity.  Otherwise one
  // thread per POSITION (coalesced reads of the order, independent of each other), not one per word.
  if (!reordered) {
    for (int w = tid; w < cb; w += kSweepThreads) kept_orig[w] = kept_sorted[w];
  } else {
    for (int p = tid; p < n; p += kSweepThreads) {
      if ((kept_sorted[p >> 6] >> (p & 63)) & 1ull) {
        const int orig = ord[p];
        atomicOr(reinterpret_cast<unsigned*>(kept_orig) + (orig >> 5), 1u << (orig & 31));
      }
    }
  }
  __syncthreads();
  // ascending original indices: block-wide exclusive scan of per-word popcounts (chunk by chunk) into word offsets ...
  int* word_pos = reinterpret_cast<int*>(removed);  // the suppression words are no longer needed
  int running = 0;
  for (int w0 = 0; w0 < cb; w0 += kSweepThreads) {
    const int w = w0 + tid;
    const int cnt = w < cb ? __popcll(kept_orig[w]) : 0;
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
      int s2 = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, s2, o);
        if (lane >= o) s2 += t;
      }
      if (lane < kSweepThreads / 32) scan[lane] = s2 - v;
      if (lane == kSweepThreads / 32 - 1) scan[kSweepThreads / 32] = s2;
    }
    __syncthreads();
    if (w < cb) word_pos[w] = running + scan[warp] + incl - cnt;
    running += scan[kSweepThreads / 32];
    __syncthreads();
  }
  // ... then one thread per box index writes its own slot
  const int limit = max_keep > 0 ? min(max_keep, keep_stride) : keep_stride;
  for (int p = tid; p < cb * kTile; p += kSweepThreads) {
    const unsigned long long bits = kept_orig[p >> 6];
    if ((bits >> (p & 63)) & 1ull) {
      const int pos = word_pos[p >> 6] + __popcll(bits & ((1ull << (p & 63)) - 1ull));
      if (pos < limit) out[pos] = (long long)p;
    }
  }
  int total = running;
  if (max_keep > 0) total = min(total, max_keep);
  total = min(total, keep_stride);
  for (int i = total + tid; i < keep_stride; i += kSweepThreads) out[i] = -1;
  if (tid == 0) n_keep[nb.first_image + img] = total;
}

__global__ void nms_fill_empty_kernel(long long* keep, int keep_stride, int* n_keep, int image) {
  for (int i = threadIdx.x; i < keep_stride; i += blockDim.x) keep[(long long)image * keep_stride + i] = -1;
  if (threadIdx.x == 0) n_keep[image] = 0;
}

struct NmsLayout {
  size_t sorted_boxes, order, unsorted, done, mask, total;
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
  l.unsorted = l.order + align(total_boxes * sizeof(int));
  l.done = l.unsorted + align((size_t)n_images * sizeof(int));
  l.mask = l.done + align((size_t)n_images * sizeof(int));
  l.total = l.mask + align(words * 8);
  return l;
}

// abr_nms_batched with an optional device-side box count per image (used by the RPN proposal path, where the number of
// boxes that survive the small-box filter is only known on the device): offsets_host then describes capacities.
int nms_run(const float* boxes, const float* scores, const int* offsets_host, const int* counts_dev,
            const unsigned long long* invalid, int n_images, float thresh, int ge, int max_keep, int64_t* keep,
            int keep_stride, int32_t* n_keep, void* workspace, size_t workspace_bytes, cudaStream_t st);

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
  return nms_run(boxes, scores, offsets_host, nullptr, nullptr, n_images, thresh, ge, max_keep, keep, keep_stride, n_keep, workspace,
                 workspace_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

namespace abr {

int nms_run(const float* boxes, const float* scores, const int* offsets_host, const int* counts_dev,
            const unsigned long long* invalid, int n_images, float thresh, int ge, int max_keep, int64_t* keep,
            int keep_stride, int32_t* n_keep, void* workspace, size_t workspace_bytes, cudaStream_t st) {
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
  if (total > 0) ABR_REQUIRE(boxes && scores, ABR_ERR_BAD_ARG, "nms: null boxes/scores");
  if (total > 0 && (reinterpret_cast<uintptr_t>(boxes) & 15))
    ABR_REQUIRE(false, ABR_ERR_BAD_ARG, "nms: boxes must be 16-byte aligned");
  const NmsLayout lay = nms_layout(offsets_host, n_images);
  if (total > 0)
    ABR_REQUIRE(workspace && workspace_bytes >= lay.total, ABR_ERR_WORKSPACE, "nms: workspace %zu B < %zu B", workspace_bytes, lay.total);
  char* ws = static_cast<char*>(workspace);
  float4* sorted_boxes = reinterpret_cast<float4*>(ws + lay.sorted_boxes);
  int* order = reinterpret_cast<int*>(ws + lay.order);
  int* unsorted = reinterpret_cast<int*>(ws + lay.unsorted);
  int* done = reinterpret_cast<int*>(ws + lay.done);
  // unsorted and done are adjacent regions: one memset clears both
  if (total > 0) ABR_CUDA_OK(cudaMemsetAsync(unsorted, 0, lay.mask - lay.unsorted, st));
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(ws + lay.mask);

  long long mask_cursor = 0, invalid_cursor = 0;
  for (int base = 0; base < n_images; base += kImagesPerLaunch) {
    NmsBatch nb;
    nb.n_images = 0;
    nb.first_image = base;
    nb.counts = counts_dev;
    nb.invalid = invalid;
    int nmax = 0;
    const int lim = n_images - base < kImagesPerLaunch ? n_images - base : kImagesPerLaunch;
    for (int i = 0; i < lim; i++) {
      const int n = offsets_host[base + i + 1] - offsets_host[base + i];
      nb.box_off[i] = offsets_host[base + i];
      nb.n[i] = nb.n_full[i] = n;
      nb.mask_off[i] = mask_cursor;
      mask_cursor += (long long)n * ceil_div(n, kTile);
      nb.invalid_off[i] = invalid_cursor;
      invalid_cursor += ceil_div(n, kTile);
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
    const size_t smem = (size_t)cbmax * 3 * 8 + (kSweepThreads / 32 + 1) * sizeof(int);
    ABR_REQUIRE(smem <= 200 * 1024, ABR_ERR_UNSUPPORTED, "nms: %d boxes in one image need %zu B of shared memory", nmax, smem);
    if (smem > 48 * 1024) ABR_CUDA_OK(cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    presort_kernel<<<dim3(ceil_div(nmax, 256), lim), 256, 0, st>>>(nb, boxes, scores, sorted_boxes, order, unsorted);
    ABR_CHECK_LAUNCH("nms_presort");
    rank_sort_kernel<<<dim3(ceil_div(nmax, kRankThreads), lim), kRankThreads, 0, st>>>(nb, boxes, scores, sorted_boxes, order, unsorted);
    ABR_CHECK_LAUNCH("nms_rank_sort");

    // Prefix pass (only with a max_keep cut): sorted images rarely need more than the first ~2*max_keep boxes to collect
    // max_keep survivors, and both the mask and the sweep are quadratic / linear in what they look at.  Images the prefix
    // settles (sorted and max_keep reached) are skipped by the full pass; the others are redone in full.
    const int prefix = max_keep > 0 ? ceil_div(2 * max_keep > 512 ? 2 * max_keep : 512, kTile) * kTile : 0;
    const int* done_in = nullptr;
    if (prefix > 0 && prefix < nmax) {
      NmsBatch pb = nb;
      int pmax = 0;
      for (int i = 0; i < lim; i++) {
        pb.n[i] = nb.n_full[i] < prefix ? nb.n_full[i] : prefix;
        pmax = pb.n[i] > pmax ? pb.n[i] : pmax;
      }
      const int pcb = ceil_div(pmax, kTile);
      const long long ptiles = (long long)pcb * (pcb + 1) / 2;
      iou_mask_kernel<<<dim3((unsigned)ceil_div<long long>(ptiles, kMaskTilesPerCta), lim), kTile * kMaskTilesPerCta, 0, st>>>(pb, sorted_boxes, mask, thresh, ge, unsorted);
      ABR_CHECK_LAUNCH("nms_iou_mask_prefix");
      sweep_kernel<<<lim, kSweepThreads, smem, st>>>(pb, mask, order, unsorted, max_keep, reinterpret_cast<long long*>(keep), keep_stride, n_keep, done, 1);
      ABR_CHECK_LAUNCH("nms_sweep_prefix");
      done_in = done;
    }
    const long long tiles = (long long)cbmax * (cbmax + 1) / 2;
    iou_mask_kernel<<<dim3((unsigned)ceil_div<long long>(tiles, kMaskTilesPerCta), lim), kTile * kMaskTilesPerCta, 0, st>>>(nb, sorted_boxes, mask, thresh, ge, done_in);
    ABR_CHECK_LAUNCH("nms_iou_mask");
    sweep_kernel<<<lim, kSweepThreads, smem, st>>>(nb, mask, order, unsorted, max_keep, reinterpret_cast<long long*>(keep), keep_stride, n_keep, done, 0);
    ABR_CHECK_LAUNCH("nms_sweep");
  }
  return ABR_OK;
}

}  // namespace abr
