// ard.cu -- Attentive RoI Distillation loss, forward + backward in ONE kernel (sm_100a).
//
// Semantics: distillation/distillation.py:86-130 of the reference with the call-site argument order of
// tools/train_incremental.py:115 (f_old = teacher / old model, no grad; f_new = student):
//     m_x[n,p] = mean_c f_x^2          A_x = HW * softmax_p(m_x)
//     L = mean_{n,c,p} A_old (f_old - f_new)^2  +  gamma * mean_{n,p} |A_new - A_old|
//     dL/df_new = 2 A_old (f_new - f_old)/(N C HW) + (2 f_new / C) HW s_p (g_p - sum_q g_q s_q),
//         s = softmax(m_new),  g = gamma sign(A_new - A_old)/(N HW)
// The reference runs ~16 full-tensor PyTorch kernels forward plus autograd's backward and keeps ~10 [N,C,H,W]
// temporaries.  Here one CTA owns one RoI: pass 1 streams f_old / f_new once (coalesced, vectorised) and reduces
// sum_c f_old^2, sum_c f_new^2 and sum_c (f_new - f_old)^2 per position; one warp then does both softmaxes and
// every per-position coefficient in shared memory; pass 2 re-reads the RoI (it was just read by this SM, 0.4-1.6 MB,
// so it is served from L2) and writes the gradient.  HBM traffic = read 2 tensors + write 1.  The per-RoI loss
// partials are reduced in fixed order by the last CTA to finish (deterministic, no float atomics).
#include <cstdlib>

#include "common.cuh"
#include "roi_v2.h"

namespace abr {

struct ArdParams {
  int N, C, HW;
  float gamma, grad_scale;
  float* partials;        // [N][2]: sum_p A_old[p] * dd[p],  sum_p |A_new - A_old|
  unsigned int* counter;  // zeroed before launch
  float* loss3;
};

// Shared memory (floats): m_old[HW] m_new[HW] dd[HW] a_old[HW] kk[HW]
__device__ __forceinline__ void ard_position_phase(const ArdParams& p, float* m_old, float* m_new, const float* dd,
                                                   float* a_old, float* kk, int n, bool write_partials = true) {
  // executed by warp 0 only; m_* hold sum_c f^2 on entry
  const int lane = threadIdx.x & 31;
  const int HW = p.HW;
  const float invC = 1.f / (float)p.C;
  float mxo = -INFINITY, mxn = -INFINITY;
  for (int i = lane; i < HW; i += 32) {
    const float a = m_old[i] * invC, b = m_new[i] * invC;
    m_old[i] = a; m_new[i] = b;
    mxo = fmaxf(mxo, a); mxn = fmaxf(mxn, b);
  }
  mxo = warp_max(mxo); mxn = warp_max(mxn);
  float so = 0.f, sn = 0.f;
  for (int i = lane; i < HW; i += 32) {
    const float eo = expf(m_old[i] - mxo), en = expf(m_new[i] - mxn);
    m_old[i] = eo; m_new[i] = en;
    so += eo; sn += en;
  }
  so = warp_sum(so); sn = warp_sum(sn);
  const float fHW = (float)HW;
  const float gmag = p.gamma / ((float)p.N * fHW);
  float pad = 0.f, gs = 0.f, afd = 0.f;
  for (int i = lane; i < HW; i += 32) {
    const float ao = fHW * (m_old[i] / so);
    const float s = m_new[i] / sn;
    const float d = __fsub_rn(__fmul_rn(fHW, s), ao);  // no FMA: identical inputs must give exactly 0, as in the reference
    pad += fabsf(d);
    const float g = d > 0.f ? gmag : (d < 0.f ? -gmag : 0.f);
    gs = fmaf(g, s, gs);
    afd = fmaf(ao, dd[i], afd);
    a_old[i] = ao;
    m_new[i] = s;  // keep the softmax
    kk[i] = g;
  }
  pad = warp_sum(pad); gs = warp_sum(gs); afd = warp_sum(afd);
  const float inv_all = 1.f / ((float)p.N * (float)p.C * fHW);
  const float ka = 2.f * inv_all * p.grad_scale;
  const float kb = 2.f * invC * fHW * p.grad_scale;
  for (int i = lane; i < HW; i += 32) {
    kk[i] = kb * m_new[i] * (kk[i] - gs);
    a_old[i] *= ka;  // pass 2 needs only ka * A_old
  }
  if (lane == 0 && write_partials) {
    p.partials[2 * n] = afd;
    p.partials[2 * n + 1] = pad;
  }
}

// Last CTA: fixed-order reduction of the per-RoI partials (double accumulation).
__device__ __forceinline__ void ard_finish(const ArdParams& p) {
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(p.counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < 32) {
    double afd = 0.0, pad = 0.0;
    const volatile float* part = p.partials;
    for (int i = threadIdx.x; i < p.N; i += 32) { afd += (double)part[2 * i]; pad += (double)part[2 * i + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      afd += __shfl_xor_sync(0xffffffffu, afd, o);
      pad += __shfl_xor_sync(0xffffffffu, pad, o);
    }
    if (threadIdx.x == 0) {
      const double l_afd = afd / ((double)p.N * p.C * p.HW);
      const double l_pad = pad / ((double)p.N * p.HW);
      p.loss3[0] = (float)(l_afd + (double)p.gamma * l_pad);
      p.loss3[1] = (float)l_afd;
      p.loss3[2] = (float)l_pad;
    }
  }
}

// ------------------------------------------------------------------------------------------ NHWC: [N][HW][C]
// A warp owns whole position rows (C contiguous elements); lanes read 16-byte vectors.
template <typename T, int V, bool GRAD>
__global__ void __launch_bounds__(1024) ard_nhwc_kernel(ArdParams p, const T* __restrict__ f_old,
                                                       const T* __restrict__ f_new, T* __restrict__ grad) {
  extern __shared__ float sm[];
  const int HW = p.HW, C = p.C;
  float* m_old = sm;
  float* m_new = m_old + HW;
  float* dd = m_new + HW;
  float* a_old = dd + HW;
  float* kk = a_old + HW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  // persistent: one CTA per SM walks the RoIs, so the RoIs in flight (148 x 2*C*HW elements) stay L2-resident
  // between pass 1 and pass 2
  for (int n = blockIdx.x; n < p.N; n += gridDim.x) {
  const size_t base = (size_t)n * HW * C;
  const T* __restrict__ fo = f_old + base;
  const T* __restrict__ fn = f_new + base;

  for (int pos = warp; pos < HW; pos += nwarp) {
    const T* ro = fo + (size_t)pos * C;
    const T* rn = fn + (size_t)pos * C;
    float so = 0.f, sn = 0.f, sd = 0.f;
#pragma unroll 4
    for (int c = lane * V; c < C; c += 32 * V) {
      float a[V], b[V];
      VecIO<T, V>::load(ro + c, a);
      VecIO<T, V>::load(rn + c, b);
#pragma unroll
      for (int k = 0; k < V; k++) {
        so = fmaf(a[k], a[k], so);
        sn = fmaf(b[k], b[k], sn);
        const float d = b[k] - a[k];
        sd = fmaf(d, d, sd);
      }
    }
    so = warp_sum(so); sn = warp_sum(sn); sd = warp_sum(sd);
    if (lane == 0) { m_old[pos] = so; m_new[pos] = sn; dd[pos] = sd; }
  }
  __syncthreads();
  if (warp == 0) ard_position_phase(p, m_old, m_new, dd, a_old, kk, n);
  if (GRAD) {
    __syncthreads();
    T* g = grad + base;
    for (int pos = warp; pos < HW; pos += nwarp) {
      const T* ro = fo + (size_t)pos * C;
      const T* rn = fn + (size_t)pos * C;
      T* rg = g + (size_t)pos * C;
      const float ka = a_old[pos], kb = kk[pos];
#pragma unroll 4
      for (int c = lane * V; c < C; c += 32 * V) {
        float a[V], b[V], o[V];
        VecIO<T, V>::load(ro + c, a);
        VecIO<T, V>::load(rn + c, b);
#pragma unroll
        for (int k = 0; k < V; k++) o[k] = fmaf(ka, b[k] - a[k], kb * b[k]);
        VecIO<T, V>::store_stream(rg + c, o);
      }
    }
  }
  __syncthreads();  // the per-position tables are reused by the next RoI
  }
  ard_finish(p);
}

// ------------------------------------------------------------------------------------------ NCHW: [N][C][HW]
// blockDim.x = G*HW: thread t owns position t % HW and channels t / HW, t / HW + G, ... so that the threads of a
// CTA always touch blockDim.x consecutive elements (coalesced) and accumulate their position in a register.
template <typename T, bool GRAD>
__global__ void __launch_bounds__(1024) ard_nchw_kernel(ArdParams p, const T* __restrict__ f_old,
                                                       const T* __restrict__ f_new, T* __restrict__ grad, int G) {
  extern __shared__ float sm[];
  const int HW = p.HW, C = p.C;
  float* m_old = sm;
  float* m_new = m_old + HW;
  float* dd = m_new + HW;
  float* a_old = dd + HW;
  float* kk = a_old + HW;
  float* red = kk + HW;  // [3][G*HW]
  const int T_ = G * HW;  // blockDim.x is T_ rounded up to whole warps; the surplus threads only hit the barriers
  const int tid = threadIdx.x;
  const bool active = tid < T_;
  const int grp = active ? tid / HW : C, pos = active ? tid - grp * HW : 0;  // grp == C: loops are empty
  for (int n = blockIdx.x; n < p.N; n += gridDim.x) {  // persistent, see ard_nhwc_kernel
  const size_t base = (size_t)n * HW * C;
  const T* __restrict__ fo = f_old + base;
  const T* __restrict__ fn = f_new + base;

  // fp32 runs of 16 channels folded into double totals: as accurate as the reference's pairwise fp32 sums or better,
  // also for the hundreds of channels a thread walks when HW is large
  double tso = 0.0, tsn = 0.0, tsd = 0.0;
  for (int c0 = grp; c0 < C; c0 += 16 * G) {
    float so = 0.f, sn = 0.f, sd = 0.f;
#pragma unroll 4
    for (int c = c0; c < min(C, c0 + 16 * G); c += G) {
      float a[1], b[1];
      VecIO<T, 1>::load(fo + (size_t)c * HW + pos, a);
      VecIO<T, 1>::load(fn + (size_t)c * HW + pos, b);
      so = fmaf(a[0], a[0], so);
      sn = fmaf(b[0], b[0], sn);
      const float d = b[0] - a[0];
      sd = fmaf(d, d, sd);
    }
    tso += (double)so; tsn += (double)sn; tsd += (double)sd;
  }
  if (active) { red[tid] = (float)tso; red[T_ + tid] = (float)tsn; red[2 * T_ + tid] = (float)tsd; }
  __syncthreads();
  if (tid < HW) {
    double x = 0.0, y = 0.0, z = 0.0;
    for (int g = 0; g < G; g++) { x += (double)red[g * HW + tid]; y += (double)red[T_ + g * HW + tid]; z += (double)red[2 * T_ + g * HW + tid]; }
    m_old[tid] = (float)x; m_new[tid] = (float)y; dd[tid] = (float)z;
  }
  __syncthreads();
  if (tid < 32) ard_position_phase(p, m_old, m_new, dd, a_old, kk, n);
  if (GRAD) {
    __syncthreads();
    T* g = grad + base;
    const float ka = a_old[pos], kb = kk[pos];
#pragma unroll 4
    for (int c = grp; c < C; c += G) {
      float a[1], b[1], o[1];
      VecIO<T, 1>::load(fo + (size_t)c * HW + pos, a);
      VecIO<T, 1>::load(fn + (size_t)c * HW + pos, b);
      o[0] = fmaf(ka, b[0] - a[0], kb * b[0]);
      VecIO<T, 1>::store_stream(g + (size_t)c * HW + pos, o);
    }
  }
  __syncthreads();  // red[] and the per-position tables are reused by the next RoI
  }
  ard_finish(p);
}

// ------------------------------------------------------------------------------------------ NHWC, shared-memory resident
// One thread-block CLUSTER owns one RoI at a time and keeps BOTH tensors of it in shared memory, so HBM is touched
// exactly once per element (read f_old, read f_new, write the gradient): CTA k of the cluster holds the position
// rows [k*rows_per_cta, ...) -- for C=1024, P=7 that is 25 rows x 4 KB x 2 tensors = 200 KB in each of 2 CTAs.
// The rows arrive by TMA bulk copies (cp.async.bulk, a few rows per mbarrier so that pass 1 starts while the rest is
// still in flight); the per-position sums are exchanged through distributed shared memory (every CTA writes its rows'
// sums into all peers' tables, one cluster barrier), every CTA then evaluates the softmaxes redundantly and runs pass 2
// out of its own shared memory.
constexpr int kArdChunks = 8;
constexpr int kArdClusterThreads = 512;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> this CTA's shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// store one float into the same shared-memory variable of cluster CTA `rank`
__device__ __forceinline__ void dsmem_store(float* local_addr, unsigned rank, float v) {
  unsigned remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_addr)), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}

template <bool GRAD>
__global__ void __launch_bounds__(kArdClusterThreads, 1) ard_nhwc_cluster_kernel(ArdParams p, const float* __restrict__ f_old,
                                                                              const float* __restrict__ f_new,
                                                                              float* __restrict__ grad, int rows_per_cta) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = p.HW, C = p.C;
  float* t_old = reinterpret_cast<float*>(smem_raw);                 // [rows_per_cta][C]
  float* t_new = t_old + (size_t)rows_per_cta * C;                   // [rows_per_cta][C]
  float* ex = t_new + (size_t)rows_per_cta * C;                      // [2 parities][3][HW]: sum f_old^2, sum f_new^2, sum diff^2
  float* a_old = ex + 6 * HW;
  float* kk = a_old + HW;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(kk + HW);  // 2*rows*C + 8*HW floats: 8-byte aligned
  const unsigned rank = cluster_ctarank(), csize = cluster_nctarank();
  const int cluster_id = blockIdx.x / csize, nclusters = gridDim.x / csize;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = kArdClusterThreads / 32;
  const int row0 = rank * rows_per_cta;
  const int nrows = max(0, min(rows_per_cta, HW - row0));
  const int chunk_rows = max(1, ceil_div(rows_per_cta, kArdChunks));
  const int nchunks = ceil_div(nrows, chunk_rows);
  if (tid == 0) {
    for (int i = 0; i < kArdChunks; i++) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();  // barriers visible; every peer's shared memory is live before the first remote store

  int iter = 0;
  for (int n = cluster_id; n < p.N; n += nclusters, iter++) {
    const unsigned parity = iter & 1;
    const size_t base = ((size_t)n * HW + row0) * C;
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads of the previous RoI are done
      for (int ch = 0; ch < nchunks; ch++) {
        const int r0 = ch * chunk_rows, rn = min(chunk_rows, nrows - r0);
        const unsigned bytes = (unsigned)rn * C * sizeof(float);
        mbar_expect_tx(&bars[ch], 2 * bytes);
        tma_load_1d(t_old + (size_t)r0 * C, f_old + base + (size_t)r0 * C, bytes, &bars[ch]);
        tma_load_1d(t_new + (size_t)r0 * C, f_new + base + (size_t)r0 * C, bytes, &bars[ch]);
      }
    }
    float* ex_old = ex + parity * 3 * HW;
    float* ex_new = ex_old + HW;
    float* ex_dd = ex_new + HW;
    // pass 1 out of shared memory, chunk by chunk as the copies land
    for (int row = warp; row < nrows; row += nwarp) {
      mbar_wait(&bars[row / chunk_rows], parity);
      const float4* ro = reinterpret_cast<const float4*>(t_old + (size_t)row * C);
      const float4* rn = reinterpret_cast<const float4*>(t_new + (size_t)row * C);
      float so = 0.f, sn = 0.f, sd = 0.f;
#pragma unroll 4
      for (int c = lane; c < C / 4; c += 32) {
        const float4 a = ro[c], b = rn[c];
        so = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, so))));
        sn = fmaf(b.x, b.x, fmaf(b.y, b.y, fmaf(b.z, b.z, fmaf(b.w, b.w, sn))));
        const float d0 = b.x - a.x, d1 = b.y - a.y, d2 = b.z - a.z, d3 = b.w - a.w;
        sd = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, sd))));
      }
      so = warp_sum(so); sn = warp_sum(sn); sd = warp_sum(sd);
      if (lane < (int)csize) {  // lane r publishes this row's sums in CTA r's table
        dsmem_store(ex_old + row0 + row, lane, so);
        dsmem_store(ex_new + row0 + row, lane, sn);
        dsmem_store(ex_dd + row0 + row, lane, sd);
      }
    }
    cluster_sync_all();  // all HW sums of this RoI are in every CTA's table
    if (warp == 0) ard_position_phase(p, ex_old, ex_new, ex_dd, a_old, kk, n, rank == 0);
    __syncthreads();
    if (GRAD) {
      float* g = grad + base;
      for (int row = warp; row < nrows; row += nwarp) {
        const float4* ro = reinterpret_cast<const float4*>(t_old + (size_t)row * C);
        const float4* rn = reinterpret_cast<const float4*>(t_new + (size_t)row * C);
        float4* rg = reinterpret_cast<float4*>(g + (size_t)row * C);
        const float ka = a_old[row0 + row], kb = kk[row0 + row];
#pragma unroll 4
        for (int c = lane; c < C / 4; c += 32) {
          const float4 a = ro[c], b = rn[c];
          float4 o;
          o.x = fmaf(ka, b.x - a.x, kb * b.x);
          o.y = fmaf(ka, b.y - a.y, kb * b.y);
          o.z = fmaf(ka, b.z - a.z, kb * b.z);
          o.w = fmaf(ka, b.w - a.w, kb * b.w);
          rg[c] = o;
        }
      }
    }
    __syncthreads();  // the tiles and a_old / kk are free for the next RoI
  }
  cluster_sync_all();  // no CTA exits while a peer may still store into its shared memory
  ard_finish(p);
}

// cluster size and rows per CTA for the shared-memory resident kernel; 0 = does not fit (use the two-pass kernel)
static int ard_cluster_size(int C, int HW, size_t& smem_bytes, int& rows_per_cta) {
  for (int cs = 1; cs <= 8; cs *= 2) {
    if (cs > HW) break;
    rows_per_cta = ceil_div(HW, cs);
    smem_bytes = (size_t)2 * rows_per_cta * C * sizeof(float) + (size_t)(8 * HW + 2) * sizeof(float) + kArdChunks * 8 + 128;
    if (smem_bytes <= 227 * 1024) return cs;
  }
  return 0;
}

// Clusters of `cs` CTAs that can be resident at once (a cluster must fit inside one GPC, so this can be fewer than
// SMs / cs): a persistent grid larger than that would leave late clusters to run after the others have finished.
template <typename K>
static int resident_clusters(K kern, int cs, size_t smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(num_sms() / cs * cs);
  cfg.blockDim = dim3(kArdClusterThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = num_sms() / cs;
  }
  return n < num_sms() / cs ? n : num_sms() / cs;
}

static int launch_nhwc_cluster(const ArdParams& p, const void* fo, const void* fn, void* g, cudaStream_t st, int cs,
                               size_t smem, int rows_per_cta) {
  auto kern = g ? ard_nhwc_cluster_kernel<true> : ard_nhwc_cluster_kernel<false>;
  ABR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (cs > 8) ABR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  int clusters = resident_clusters(kern, cs, smem);
  if (clusters > p.N) clusters = p.N;
  if (clusters < 1) clusters = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * cs);
  cfg.blockDim = dim3(kArdClusterThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ABR_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p, static_cast<const float*>(fo), static_cast<const float*>(fn), static_cast<float*>(g), rows_per_cta));
  ABR_CHECK_LAUNCH("ard_forward_backward (cluster)");
  return ABR_OK;
}

// ------------------------------------------------------------------------------------------ NCHW cluster kernel
// [N][C][HW]: a RoI is the same contiguous C*HW*4 bytes as in NHWC, so the CTAs of a cluster take contiguous channel
// ranges with 1-D bulk copies.  The first T = HW*G threads (G = 512 / HW) each own the four consecutive floats
// 4t .. 4t+3 of every sweep of S = 4*T floats through the tile: S is a multiple of HW, so a thread's four floats
// always belong to the same four positions (4t+j) mod HW -- 16-byte shared-memory loads and global stores, per-thread
// accumulators for its four positions.  Per-position sums are partial per CTA: every CTA stores its three partial rows
// into every peer's table and each CTA adds them up in rank order (deterministic).
template <bool GRAD>
__global__ void __launch_bounds__(kArdClusterThreads, 1) ard_nchw_cluster_kernel(ArdParams p, const float* __restrict__ f_old,
                                                                              const float* __restrict__ f_new,
                                                                              float* __restrict__ grad, int ch_per_cta, int G) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = p.HW, C = p.C;
  const unsigned rank = cluster_ctarank(), csize = cluster_nctarank();
  const int T = HW * G, S = 4 * T;                                  // owning threads, floats per sweep
  float* t_old = reinterpret_cast<float*>(smem_raw);                 // [ch_per_cta][HW]
  float* t_new = t_old + (size_t)ch_per_cta * HW;
  float* ex = t_new + (size_t)ch_per_cta * HW;                       // [2 parities][3][csize][HW]
  float* part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ex + (size_t)6 * csize * HW) + 15) & ~(uintptr_t)15);  // [3][S], 16-byte aligned:
                                                                     // entry e = group (e / HW), position (e % HW)
  float* tot = part + (size_t)3 * S;                                 // [3][HW]
  float* a_old = tot + 3 * HW;
  float* kk = a_old + HW;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(
      (reinterpret_cast<uintptr_t>(kk + HW) + 7) & ~(uintptr_t)7);
  const int cluster_id = blockIdx.x / csize, nclusters = gridDim.x / csize;
  const int tid = threadIdx.x;
  const int c0 = rank * ch_per_cta;
  const int nch = max(0, min(ch_per_cta, C - c0));
  const int nfl = nch * HW;                                          // floats of this CTA's tile (a multiple of 4)
  const int chunk_ch = max(1, ceil_div(ch_per_cta, kArdChunks));
  const int chunk_fl = chunk_ch * HW;                                // floats per chunk (a multiple of 4)
  const int nchunks = ceil_div(nch, chunk_ch);
  const bool active = tid < T;
  int pos[4];
#pragma unroll
  for (int j2 = 0; j2 < 4; j2++) pos[j2] = (4 * tid + j2) % HW;
  if (tid == 0) {
    for (int i = 0; i < kArdChunks; i++) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();

  int iter = 0;
  for (int n = cluster_id; n < p.N; n += nclusters, iter++) {
    const unsigned parity = iter & 1;
    const size_t base = ((size_t)n * C + c0) * HW;
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      for (int ch = 0; ch < nchunks; ch++) {
        const int r0 = ch * chunk_ch, rn = min(chunk_ch, nch - r0);
        const unsigned bytes = (unsigned)rn * HW * sizeof(float);
        mbar_expect_tx(&bars[ch], 2 * bytes);
        tma_load_1d(t_old + (size_t)r0 * HW, f_old + base + (size_t)r0 * HW, bytes, &bars[ch]);
        tma_load_1d(t_new + (size_t)r0 * HW, f_new + base + (size_t)r0 * HW, bytes, &bars[ch]);
      }
    }
    float so[4] = {0.f, 0.f, 0.f, 0.f}, sn[4] = {0.f, 0.f, 0.f, 0.f}, sd[4] = {0.f, 0.f, 0.f, 0.f};
    int ready = -1;  // chunks 0..ready have landed
    if (active) {
      for (int e = 4 * tid; e < nfl; e += S) {
        const int ch = e / chunk_fl;
        while (ready < ch) mbar_wait(&bars[++ready], parity);
        const float4 a = *reinterpret_cast<const float4*>(t_old + e), b = *reinterpret_cast<const float4*>(t_new + e);
        so[0] = fmaf(a.x, a.x, so[0]); so[1] = fmaf(a.y, a.y, so[1]); so[2] = fmaf(a.z, a.z, so[2]); so[3] = fmaf(a.w, a.w, so[3]);
        sn[0] = fmaf(b.x, b.x, sn[0]); sn[1] = fmaf(b.y, b.y, sn[1]); sn[2] = fmaf(b.z, b.z, sn[2]); sn[3] = fmaf(b.w, b.w, sn[3]);
        const float d0 = b.x - a.x, d1 = b.y - a.y, d2 = b.z - a.z, d3 = b.w - a.w;
        sd[0] = fmaf(d0, d0, sd[0]); sd[1] = fmaf(d1, d1, sd[1]); sd[2] = fmaf(d2, d2, sd[2]); sd[3] = fmaf(d3, d3, sd[3]);
      }
    }
    while (ready < nchunks - 1) mbar_wait(&bars[++ready], parity);  // every thread observes every chunk's phase
    if (active) {
      *reinterpret_cast<float4*>(part + 4 * tid) = make_float4(so[0], so[1], so[2], so[3]);
      *reinterpret_cast<float4*>(part + S + 4 * tid) = make_float4(sn[0], sn[1], sn[2], sn[3]);
      *reinterpret_cast<float4*>(part + 2 * S + 4 * tid) = make_float4(sd[0], sd[1], sd[2], sd[3]);
    }
    __syncthreads();
    float* ex_par = ex + (size_t)parity * 3 * csize * HW;
    for (int t = tid; t < 3 * HW; t += kArdClusterThreads) {
      const int q = t / HW, pp = t - q * HW;
      double acc = 0.0;  // up to 4*G (hundreds of) partial sums: accumulated in double, like a pairwise fp32 sum or better
      for (int gg = 0; gg < 4 * G; gg++) acc += (double)part[q * S + gg * HW + pp];
      const float v = (float)acc;
      for (unsigned r = 0; r < csize; r++) dsmem_store(ex_par + ((size_t)q * csize + rank) * HW + pp, r, v);
    }
    cluster_sync_all();  // every CTA's partial rows are in every CTA's table
    for (int t = tid; t < 3 * HW; t += kArdClusterThreads) {
      const int q = t / HW, pp = t - q * HW;
      float v = 0.f;
      for (unsigned r = 0; r < csize; r++) v += ex_par[((size_t)q * csize + r) * HW + pp];
      tot[t] = v;
    }
    __syncthreads();
    if (tid < 32) ard_position_phase(p, tot, tot + HW, tot + 2 * HW, a_old, kk, n, rank == 0);
    __syncthreads();
    if (GRAD && active) {
      float* gp = grad + base;
      const float ka0 = a_old[pos[0]], ka1 = a_old[pos[1]], ka2 = a_old[pos[2]], ka3 = a_old[pos[3]];
      const float kb0 = kk[pos[0]], kb1 = kk[pos[1]], kb2 = kk[pos[2]], kb3 = kk[pos[3]];
#pragma unroll 2
      for (int e = 4 * tid; e < nfl; e += S) {
        const float4 a = *reinterpret_cast<const float4*>(t_old + e), b = *reinterpret_cast<const float4*>(t_new + e);
        float4 o;
        o.x = fmaf(ka0, b.x - a.x, kb0 * b.x);
        o.y = fmaf(ka1, b.y - a.y, kb1 * b.y);
        o.z = fmaf(ka2, b.z - a.z, kb2 * b.z);
        o.w = fmaf(ka3, b.w - a.w, kb3 * b.w);
        *reinterpret_cast<float4*>(gp + e) = o;
      }
    }
    __syncthreads();  // tiles, part, tot, a_old / kk are free for the next RoI
  }
  cluster_sync_all();
  ard_finish(p);
}

// cluster size, channels per CTA and thread groups for the NCHW shared-memory resident kernel; 0 = does not fit
static int ard_nchw_cluster_size(int C, int HW, size_t& smem_bytes, int& ch_per_cta, int& G) {
  if (HW > kArdClusterThreads || ((size_t)C * HW) % 4 != 0) return 0;
  G = kArdClusterThreads / HW;
  for (int cs = 1; cs <= 8; cs *= 2) {
    if (cs > C) break;
    ch_per_cta = ceil_div(C, cs);
    if (((size_t)ch_per_cta * HW) % 4 != 0) continue;  // 16-byte bulk copies and vector accesses
    // every chunk of channels must be a 16-byte multiple as well
    const int chunk_ch = ceil_div(ch_per_cta, kArdChunks) > 0 ? ceil_div(ch_per_cta, kArdChunks) : 1;
    if (((size_t)chunk_ch * HW) % 4 != 0) continue;
    smem_bytes = ((size_t)2 * ch_per_cta * HW + (size_t)6 * cs * HW + (size_t)12 * G * HW + 5 * HW) * sizeof(float) + 16 + 8 +
                 kArdChunks * 8 + 128;
    if (smem_bytes <= 227 * 1024) return cs;
  }
  return 0;
}

static int launch_nchw_cluster(const ArdParams& p, const void* fo, const void* fn, void* g, cudaStream_t st, int cs,
                               size_t smem, int ch_per_cta, int G) {
  auto kern = g ? ard_nchw_cluster_kernel<true> : ard_nchw_cluster_kernel<false>;
  ABR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (cs > 8) ABR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  int clusters = resident_clusters(kern, cs, smem);
  if (clusters > p.N) clusters = p.N;
  if (clusters < 1) clusters = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * cs);
  cfg.blockDim = dim3(kArdClusterThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ABR_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p, static_cast<const float*>(fo), static_cast<const float*>(fn), static_cast<float*>(g), ch_per_cta, G));
  ABR_CHECK_LAUNCH("ard_forward_backward (nchw cluster)");
  return ABR_OK;
}

// ------------------------------------------------------------------------------------------ ARD from channel sums
// The pooling kernel of the fused step (roi_v2.cu, v2_fwd_kernel<NT = 2>) leaves, per RoI and channel slice, the three
// channel sums of every position.  One small CTA per RoI folds the slices in fixed order (double), runs the same
// position phase as the kernels above and writes the two per-position coefficients of
//     dL/df_new = ka * (f_new - f_old) + kb * f_new
// which the fused backward applies on the fly; the loss partials are reduced by the last CTA (ard_finish).
__global__ void __launch_bounds__(256, 8) ard_coeff_kernel(ArdParams p, const float* __restrict__ sums, int nslices,
                                                       float2* __restrict__ coef, float4* __restrict__ zero_fill, size_t zero_n) {
  extern __shared__ float sm[];
  const int HW = p.HW, n = blockIdx.x;
  // the gradient map the next kernel accumulates into: zero-filled here (stores retire behind the loads below)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < zero_n; i += (size_t)gridDim.x * blockDim.x)
    zero_fill[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float* m_old = sm;
  float* m_new = m_old + HW;
  float* dd = m_new + HW;
  float* a_old = dd + HW;
  float* kk = a_old + HW;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    double so = 0.0, sn = 0.0, sd = 0.0;
    for (int s = 0; s < nslices; s++) {
      const float* q = sums + (((size_t)n * nslices + s) * HW + i) * 3;
      so += (double)__ldg(q); sn += (double)__ldg(q + 1); sd += (double)__ldg(q + 2);
    }
    m_old[i] = (float)so; m_new[i] = (float)sn; dd[i] = (float)sd;
  }
  __syncthreads();
  if (threadIdx.x < 32) ard_position_phase(p, m_old, m_new, dd, a_old, kk, n);
  __syncthreads();
  for (int i = threadIdx.x; i < HW; i += blockDim.x) coef[(size_t)n * HW + i] = make_float2(a_old[i], kk[i]);
  ard_finish(p);
}

size_t ard_coeff_workspace_bytes(int N) { return 256 + (size_t)(N > 0 ? N : 0) * 2 * sizeof(float); }

int ard_coeff_run(const float* sums, int nslices, float2* coef, float* loss3, int N, int C, int HW, float gamma, float grad_scale,
                  void* ws, cudaStream_t st, bool counter_is_clear, void* zero_fill, size_t zero_bytes) {
  ArdParams p;
  p.N = N; p.C = C; p.HW = HW; p.gamma = gamma; p.grad_scale = grad_scale;
  p.counter = static_cast<unsigned int*>(ws);
  p.partials = reinterpret_cast<float*>(static_cast<char*>(ws) + 256);
  p.loss3 = loss3;
  if (!counter_is_clear) ABR_CUDA_OK(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), st));
  const int threads = HW >= 192 ? 256 : (HW >= 96 ? 128 : 64);
  ard_coeff_kernel<<<N, threads, (size_t)5 * HW * sizeof(float), st>>>(p, sums, nslices, coef, static_cast<float4*>(zero_fill),
                                                                       zero_fill ? zero_bytes / 16 : 0);
  ABR_CHECK_LAUNCH("ard_coefficients");
  return ABR_OK;
}

template <typename T>
__global__ void scale_if_needed_kernel(T* data, size_t n, const float* __restrict__ scale, float expected) {
  const float s = *scale;
  if (s == expected) return;
  const float f = s / expected;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v[1];
    v[0] = (float)data[i] * f;
    VecIO<T, 1>::store(data + i, v);
  }
}

template <typename T, int V>
static int launch_nhwc(const ArdParams& p, const void* fo, const void* fn, void* g, cudaStream_t st) {
  const size_t smem = (size_t)5 * p.HW * sizeof(float);
  const int threads = 1024;  // one CTA per SM: the RoIs in flight stay L2-resident between the two passes
  const int per_sm = 2048 / threads;
  const int grid = p.N < num_sms() * per_sm ? p.N : num_sms() * per_sm;
  if (g) ard_nhwc_kernel<T, V, true><<<grid, threads, smem, st>>>(p, static_cast<const T*>(fo), static_cast<const T*>(fn), static_cast<T*>(g));
  else ard_nhwc_kernel<T, V, false><<<grid, threads, smem, st>>>(p, static_cast<const T*>(fo), static_cast<const T*>(fn), nullptr);
  ABR_CHECK_LAUNCH("ard_forward_backward");
  return ABR_OK;
}

template <typename T>
static int launch_nchw(const ArdParams& p, const void* fo, const void* fn, void* g, cudaStream_t st) {
  int G = 1024 / p.HW;
  if (G < 1) G = 1;
  if (G > p.C) G = p.C;
  const int threads = ceil_div(G * p.HW, 32) * 32;
  const size_t smem = (size_t)(5 * p.HW + 3 * G * p.HW) * sizeof(float);
  const int per_sm = 2048 / threads > 0 ? 2048 / threads : 1;  // resident CTAs per SM at this block size
  const int grid = p.N < num_sms() * per_sm ? p.N : num_sms() * per_sm;
  if (g) ard_nchw_kernel<T, true><<<grid, threads, smem, st>>>(p, static_cast<const T*>(fo), static_cast<const T*>(fn), static_cast<T*>(g), G);
  else ard_nchw_kernel<T, false><<<grid, threads, smem, st>>>(p, static_cast<const T*>(fo), static_cast<const T*>(fn), nullptr, G);
  ABR_CHECK_LAUNCH("ard_forward_backward");
  return ABR_OK;
}

}  // namespace abr

using namespace abr;

extern "C" {

size_t abr_ard_workspace_bytes(int N, int C, int HW) {
  (void)C; (void)HW;
  if (N < 0) return 0;
  return 256 + (size_t)N * 2 * sizeof(float);
}

int abr_ard_forward_backward(const void* f_old, const void* f_new, void* grad_new, float* loss3, int N, int C, int HW,
                             float gamma, float grad_scale, int dtype, int layout, void* workspace,
                             size_t workspace_bytes, abr_stream_t stream) {
  ABR_REQUIRE(N > 0 && C > 0 && HW > 0, ABR_ERR_BAD_ARG, "ard: bad sizes N=%d C=%d HW=%d (the reference's mean over an empty tensor is NaN)", N, C, HW);
  ABR_REQUIRE(f_old && f_new && loss3, ABR_ERR_BAD_ARG, "ard: null pointer");
  ABR_REQUIRE(dtype == ABR_F32 || dtype == ABR_BF16, ABR_ERR_UNSUPPORTED, "ard: dtype %d not supported", dtype);
  ABR_REQUIRE(layout == ABR_NCHW || layout == ABR_NHWC, ABR_ERR_UNSUPPORTED, "ard: layout %d not supported", layout);
  ABR_REQUIRE(HW <= 1024, ABR_ERR_UNSUPPORTED, "ard: %d positions per RoI (max 1024)", HW);
  ABR_REQUIRE(workspace && workspace_bytes >= abr_ard_workspace_bytes(N, C, HW), ABR_ERR_WORKSPACE,
              "ard: workspace %zu B < %zu B", workspace_bytes, abr_ard_workspace_bytes(N, C, HW));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ArdParams p;
  p.N = N; p.C = C; p.HW = HW; p.gamma = gamma; p.grad_scale = grad_scale;
  p.counter = static_cast<unsigned int*>(workspace);
  p.partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  p.loss3 = loss3;
  ABR_CUDA_OK(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), st));
  if (layout == ABR_NHWC) {
    if (dtype == ABR_F32 && C % 4 == 0) {
      const bool use_cluster = options().ard_cluster != 0;
      size_t smem = 0;
      int rows = 0;
      const int cs = use_cluster ? ard_cluster_size(C, HW, smem, rows) : 0;
      if (cs > 0) return launch_nhwc_cluster(p, f_old, f_new, grad_new, st, cs, smem, rows);
      return launch_nhwc<float, 4>(p, f_old, f_new, grad_new, st);
    }
    if (dtype == ABR_F32) return launch_nhwc<float, 1>(p, f_old, f_new, grad_new, st);
    return (C % 8 == 0) ? launch_nhwc<__nv_bfloat16, 8>(p, f_old, f_new, grad_new, st) : launch_nhwc<__nv_bfloat16, 1>(p, f_old, f_new, grad_new, st);
  }
  if (dtype == ABR_F32) {
    const bool use_cluster = options().ard_cluster != 0;
    size_t smem = 0;
    int cpc = 0, G = 0;
    const bool aligned = ((reinterpret_cast<uintptr_t>(f_old) | reinterpret_cast<uintptr_t>(f_new)) & 15) == 0;
    const int cs = use_cluster && aligned ? ard_nchw_cluster_size(C, HW, smem, cpc, G) : 0;
    if (cs > 0) return launch_nchw_cluster(p, f_old, f_new, grad_new, st, cs, smem, cpc, G);
    return launch_nchw<float>(p, f_old, f_new, grad_new, st);
  }
  return launch_nchw<__nv_bfloat16>(p, f_old, f_new, grad_new, st);
}

int abr_scale_if_needed(void* data, size_t n, const float* scale_dev, float expected, int dtype, abr_stream_t stream) {
  ABR_REQUIRE(dtype == ABR_F32 || dtype == ABR_BF16, ABR_ERR_UNSUPPORTED, "scale: dtype %d not supported", dtype);
  ABR_REQUIRE(expected != 0.f, ABR_ERR_BAD_ARG, "scale: expected scale must be non-zero");
  if (n == 0) return ABR_OK;
  ABR_REQUIRE(data && scale_dev, ABR_ERR_BAD_ARG, "scale: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = (int)std::min<size_t>(ceil_div<size_t>(n, 256), (size_t)num_sms() * 16);
  if (dtype == ABR_F32) scale_if_needed_kernel<float><<<blocks, 256, 0, st>>>(static_cast<float*>(data), n, scale_dev, expected);
  else scale_if_needed_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<__nv_bfloat16*>(data), n, scale_dev, expected);
  ABR_CHECK_LAUNCH("scale_if_needed");
  return ABR_OK;
}

}  // extern "C"
