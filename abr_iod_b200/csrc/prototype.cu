// prototype.cu -- Prototype Box Selection on the device (sm_100a).
//
//  * abr_channel_mean        : the per-RoI descriptor of tools/prototype_box_selection.py:96-101 of the reference
//                              (torch.mean(roi_align_features.cpu(), dim=1): [R,C,7,7] -> [R,7,7]) without the 411 MB
//                              device-to-host copy it is computed after -- one streaming pass, HBM-bound.
//  * abr_prototype_distances : the scoring of Mem.mean_feature_sampling (tools/extract_memory.py:111-147): class mean of
//                              the descriptors, normalised; descriptors divided by the Frobenius norm of ALL of them;
//                              Euclidean distance of each to the mean -- in float64 like the numpy code.  The caller sorts
//                              the n distances (ascending; the closest num_bbox_per_cls boxes become the prototypes).
//  * abr_prototype_herding   : the greedy loop of Mem.herding_feature_sampling (tools/extract_memory.py:163-211): k times,
//                              the box whose inclusion brings the running centre closest to the class mean -- float64,
//                              numpy's operation order, first index on ties (argmin).
#include "common.cuh"

namespace abr {

// NHWC: one warp per (RoI, position) row of C contiguous channels.
template <typename T, int V>
__global__ void __launch_bounds__(256) channel_mean_nhwc_kernel(const T* __restrict__ x, long long rows, int C, float* __restrict__ out) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const T* p = x + (size_t)row * C;
  float acc = 0.f;
  for (int c = lane * V; c < C; c += 32 * V) {
    float v[V];
    VecIO<T, V>::load(p + c, v);
#pragma unroll
    for (int i = 0; i < V; i++) acc += v[i];
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc / (float)C;
}

// NCHW: one CTA per RoI; thread (g, p) owns position p and walks channels g, g+G, ... (consecutive threads touch
// consecutive elements), partial sums combined through shared memory.
template <typename T>
__global__ void __launch_bounds__(1024) channel_mean_nchw_kernel(const T* __restrict__ x, int C, int HW, int G, float* __restrict__ out) {
  extern __shared__ float part[];  // [G][HW]
  const int r = blockIdx.x, tid = threadIdx.x;
  const int g = tid / HW, p = tid - g * HW;
  const T* base = x + (size_t)r * C * HW;
  if (g < G) {
    float acc = 0.f;
    for (int c = g; c < C; c += G) {
      float v[1];
      VecIO<T, 1>::load(base + (size_t)c * HW + p, v);
      acc += v[0];
    }
    part[g * HW + p] = acc;
  }
  __syncthreads();
  if (tid < HW) {
    float acc = 0.f;
    for (int k = 0; k < G; k++) acc += part[k * HW + tid];
    out[(size_t)r * HW + tid] = acc / (float)C;
  }
}

__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nwarp; i++) t += scratch[i];  // same order in every thread
  return t;
}

// One CTA: mu = mean_i f_i, mu /= |mu|;  phi_i = f_i / |F|_Frobenius;  dist_i = |mu - phi_i|  (extract_memory.py:125-141)
__global__ void __launch_bounds__(256) prototype_distance_kernel(const float* __restrict__ f, int n, int F, double* __restrict__ mu,
                                                                double* __restrict__ dist) {
  __shared__ double scratch[8];
  const int tid = threadIdx.x;
  double sq_all = 0.0;
  for (long long i = tid; i < (long long)n * F; i += blockDim.x) { const double v = (double)f[i]; sq_all += v * v; }
  const double fro = sqrt(block_sum(sq_all, scratch));
  double sq_mu = 0.0;
  for (int k = tid; k < F; k += blockDim.x) {
    double acc = 0.0;
    for (int i = 0; i < n; i++) acc += (double)f[(size_t)i * F + k];
    acc /= (double)n;
    mu[k] = acc;
    sq_mu += acc * acc;
  }
  const double nmu = sqrt(block_sum(sq_mu, scratch));
  for (int k = tid; k < F; k += blockDim.x) mu[k] /= nmu;
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k < F; k++) { const double d = mu[k] - (double)f[(size_t)i * F + k] / fro; acc += d * d; }
    dist[i] = sqrt(acc);
  }
}

// numpy's pairwise summation of a contiguous run (add.reduce over the last axis): n < 8 sequential; n <= 128 eight strided
// accumulators folded as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail; longer runs split in halves.
struct HerdTerm {
  const float* x;      // the candidate's descriptor
  const double* c;     // running centre
  const double* mu;    // class mean
  double f, f1;        // f and f + 1
  __device__ __forceinline__ double operator()(int j) const {
    const double cand = __dadd_rn(__ddiv_rn(__dmul_rn(c[j], f), f1), __ddiv_rn((double)x[j], f1));  // centre*f/(f+1) + x/(f+1)
    const double d = __dsub_rn(cand, mu[j]);
    return __dmul_rn(d, d);                                                                        // pow(., 2)
  }
};
__device__ double numpy_pairwise(const HerdTerm& t, int lo, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; i++) res = __dadd_rn(res, t(lo + i));
    return res;
  }
  if (n <= 128) {
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = t(lo + j);
    int i = 8;
    for (; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; j++) r[j] = __dadd_rn(r[j], t(lo + i + j));
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; i++) res = __dadd_rn(res, t(lo + i));
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __dadd_rn(numpy_pairwise(t, lo, n2), numpy_pairwise(t, lo + n2, n - n2));
}

// One CTA.  selected[0..k): the herding order.  taken: n bytes of scratch; centre: F doubles of scratch (global).
__global__ void __launch_bounds__(256) prototype_herding_kernel(const float* __restrict__ f, int n, int F, int k, const double* __restrict__ mu,
                                                               int64_t* __restrict__ selected, unsigned char* __restrict__ taken,
                                                               double* __restrict__ centre) {
  __shared__ double best_d[256];
  __shared__ int best_i[256];
  const int tid = threadIdx.x;
  for (int i = tid; i < n; i += blockDim.x) taken[i] = 0;
  for (int j = tid; j < F; j += blockDim.x) centre[j] = 0.0;
  __syncthreads();
  for (int step = 0; step < k; step++) {
    HerdTerm t;
    t.c = centre; t.mu = mu; t.f = (double)step; t.f1 = (double)(step + 1);
    double bd = INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < n; i += blockDim.x) {
      t.x = f + (size_t)i * F;
      const double d = taken[i] ? INFINITY : numpy_pairwise(t, 0, F);
      if (d < bd || (d == bd && i < bi)) { bd = d; bi = i; }
    }
    best_d[tid] = bd; best_i[tid] = bi;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {  // argmin: the smallest distance, the first index among equals
      if (tid < s) {
        const double od = best_d[tid + s];
        const int oi = best_i[tid + s];
        if (od < best_d[tid] || (od == best_d[tid] && oi < best_i[tid])) { best_d[tid] = od; best_i[tid] = oi; }
      }
      __syncthreads();
    }
    const int pick = best_i[0];
    __syncthreads();
    if (tid == 0) { selected[step] = pick; taken[pick] = 1; }
    for (int j = tid; j < F; j += blockDim.x)
      centre[j] = __dadd_rn(__ddiv_rn(__dmul_rn(centre[j], (double)step), (double)(step + 1)), __ddiv_rn((double)f[(size_t)pick * F + j], (double)(step + 1)));
    __syncthreads();
  }
}

// ---- balanced fg / bg sampling of RoIs (modeling/balanced_positive_negative_sampler.py:19-68) for a whole batch.
// One CTA per image: positives (matched >= 1) and negatives (matched == 0) are ranked among their own kind by a
// caller-supplied random key (ties by index); the num_pos / num_neg smallest-key ones are marked.
__global__ void __launch_bounds__(256) sample_fg_bg_kernel(const int64_t* __restrict__ matched, const float* __restrict__ keys,
                                                          const int* __restrict__ offsets, int per_image, int want_pos,
                                                          unsigned char* __restrict__ pos_mask, unsigned char* __restrict__ neg_mask,
                                                          int* __restrict__ counts) {
  __shared__ int n_pos_s, n_neg_s;
  const int img = blockIdx.x, lo = offsets[img], n = offsets[img + 1] - lo;
  if (threadIdx.x == 0) { n_pos_s = 0; n_neg_s = 0; }
  __syncthreads();
  int lp = 0, ln = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t m = matched[lo + i];
    lp += m >= 1;
    ln += m == 0;
  }
  atomicAdd(&n_pos_s, lp);
  atomicAdd(&n_neg_s, ln);
  __syncthreads();
  // want_pos = int(batch_size_per_image * positive_fraction), evaluated by the caller in double like the reference
  const int num_pos = n_pos_s < want_pos ? n_pos_s : want_pos;          // min(positive.numel(), num_pos)
  const int want_neg = per_image - num_pos;
  const int num_neg = n_neg_s < want_neg ? n_neg_s : want_neg;          // min(negative.numel(), num_neg)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t m = matched[lo + i];
    const bool is_pos = m >= 1, is_neg = m == 0;
    int rank = 0;
    if (is_pos || is_neg) {
      const float ki = keys[lo + i];
      for (int j = 0; j < n; j++) {
        const int64_t mj = matched[lo + j];
        if ((is_pos ? mj >= 1 : mj == 0)) {
          const float kj = keys[lo + j];
          rank += (kj < ki) || (kj == ki && j < i);
        }
      }
    }
    pos_mask[lo + i] = is_pos && rank < num_pos;
    neg_mask[lo + i] = is_neg && rank < num_neg;
  }
  if (threadIdx.x == 0) { counts[2 * img] = num_pos; counts[2 * img + 1] = num_neg; }
}

}  // namespace abr

using namespace abr;

extern "C" {

int abr_prototype_herding(const float* features, int n, int F, int k, const double* mean, int64_t* selected, void* workspace,
                          size_t workspace_bytes, abr_stream_t stream) {
  ABR_REQUIRE(n > 0 && F > 0 && k >= 0 && k <= n, ABR_ERR_BAD_ARG, "prototype_herding: n=%d F=%d k=%d", n, F, k);
  if (k == 0) return ABR_OK;
  ABR_REQUIRE(features && mean && selected, ABR_ERR_BAD_ARG, "prototype_herding: null pointer");
  const size_t need = (size_t)F * sizeof(double) + (size_t)n;
  ABR_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0 && workspace_bytes >= need, ABR_ERR_WORKSPACE,
              "prototype_herding: needs an 8-byte aligned workspace of %zu B (got %zu)", need, workspace_bytes);
  double* centre = static_cast<double*>(workspace);
  unsigned char* taken = reinterpret_cast<unsigned char*>(centre + F);
  prototype_herding_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(features, n, F, k, mean, selected, taken, centre);
  ABR_CHECK_LAUNCH("prototype_herding");
  return ABR_OK;
}

int abr_sample_fg_bg(const int64_t* matched_idxs, const float* keys, const int* offsets_dev, int n_images, int batch_size_per_image,
                     int max_positives, uint8_t* pos_mask, uint8_t* neg_mask, int* counts, abr_stream_t stream) {
  ABR_REQUIRE(n_images >= 0 && batch_size_per_image >= 0 && max_positives >= 0 && max_positives <= batch_size_per_image, ABR_ERR_BAD_ARG,
              "sample_fg_bg: n_images=%d batch_size_per_image=%d max_positives=%d", n_images, batch_size_per_image, max_positives);
  if (n_images == 0) return ABR_OK;
  ABR_REQUIRE(matched_idxs && keys && offsets_dev && pos_mask && neg_mask && counts, ABR_ERR_BAD_ARG, "sample_fg_bg: null pointer");
  sample_fg_bg_kernel<<<n_images, 256, 0, static_cast<cudaStream_t>(stream)>>>(matched_idxs, keys, offsets_dev, batch_size_per_image,
                                                                                max_positives, pos_mask, neg_mask, counts);
  ABR_CHECK_LAUNCH("sample_fg_bg");
  return ABR_OK;
}

int abr_channel_mean(const void* pooled, int R, int C, int HW, int dtype, int layout, float* out, abr_stream_t stream) {
  ABR_REQUIRE(R >= 0 && C > 0 && HW > 0, ABR_ERR_BAD_ARG, "channel_mean: R=%d C=%d HW=%d", R, C, HW);
  if (R == 0) return ABR_OK;
  ABR_REQUIRE(pooled && out, ABR_ERR_BAD_ARG, "channel_mean: null pointer");
  ABR_REQUIRE(dtype == ABR_F32 || dtype == ABR_BF16, ABR_ERR_UNSUPPORTED, "channel_mean: dtype %d not supported", dtype);
  ABR_REQUIRE(layout == ABR_NCHW || layout == ABR_NHWC, ABR_ERR_UNSUPPORTED, "channel_mean: layout %d not supported", layout);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (layout == ABR_NHWC) {
    const long long rows = (long long)R * HW;
    const unsigned blocks = (unsigned)ceil_div<long long>(rows, 8);
    const bool aligned = (reinterpret_cast<uintptr_t>(pooled) & 15) == 0;
    if (dtype == ABR_F32) {
      if (aligned && C % 4 == 0) channel_mean_nhwc_kernel<float, 4><<<blocks, 256, 0, st>>>(static_cast<const float*>(pooled), rows, C, out);
      else channel_mean_nhwc_kernel<float, 1><<<blocks, 256, 0, st>>>(static_cast<const float*>(pooled), rows, C, out);
    } else {
      if (aligned && C % 8 == 0) channel_mean_nhwc_kernel<__nv_bfloat16, 8><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(pooled), rows, C, out);
      else channel_mean_nhwc_kernel<__nv_bfloat16, 1><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(pooled), rows, C, out);
    }
  } else {
    ABR_REQUIRE(HW <= 1024, ABR_ERR_UNSUPPORTED, "channel_mean: %d positions per RoI (max 1024)", HW);
    int G = 1024 / HW;
    if (G > C) G = C;
    const int threads = ceil_div(G * HW, 32) * 32;
    const size_t smem = (size_t)G * HW * sizeof(float);
    if (dtype == ABR_F32) channel_mean_nchw_kernel<float><<<R, threads, smem, st>>>(static_cast<const float*>(pooled), C, HW, G, out);
    else channel_mean_nchw_kernel<__nv_bfloat16><<<R, threads, smem, st>>>(static_cast<const __nv_bfloat16*>(pooled), C, HW, G, out);
  }
  ABR_CHECK_LAUNCH("channel_mean");
  return ABR_OK;
}

int abr_prototype_distances(const float* features, int n, int F, double* mean_out, double* dist, abr_stream_t stream) {
  ABR_REQUIRE(n > 0 && F > 0, ABR_ERR_BAD_ARG, "prototype_distances: n=%d F=%d", n, F);
  ABR_REQUIRE(features && mean_out && dist, ABR_ERR_BAD_ARG, "prototype_distances: null pointer");
  prototype_distance_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(features, n, F, mean_out, dist);
  ABR_CHECK_LAUNCH("prototype_distances");
  return ABR_OK;
}

}  // extern "C"
