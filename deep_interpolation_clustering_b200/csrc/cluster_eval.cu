// Device reductions behind the cluster-validity metrics that the gap loop evaluates for every K
// (p2_clustering_optK.py:401-405 -> internal_eval.py:15-147).
//
//   dic_cluster_scatter : per-cluster  sum_i ||x_i - c_label(i)||  and  sum_i ||x_i - c_label(i)||^2  in ONE pass
//                         over X - the within-cluster dispersion of Calinski-Harabasz (internal_eval.py:125-135,
//                         sklearn.metrics.calinski_harabasz_score) and the mean centroid distance s_k of
//                         Davies-Bouldin (internal_eval.py:138-147, sklearn.metrics.davies_bouldin_score).  The
//                         centroids come from the Lloyd pass's own sums / counts (dic_kmeans_assign with
//                         DIC_KM_KEEP_LABELS); everything else of the two scores is K-sized.
//   dic_dunn_minmax     : the (K, K) table of nearest distances between the points of two clusters and the largest
//                         distance inside one cluster - the O(N^2) part of the Dunn index (internal_eval.py:15-109:
//                         "nearest" inter-cluster distances, "farthest" diameter) - from 64 x 64 distance tiles that
//                         never leave the registers; only the upper triangle of the tile grid is visited.
//
// Both are HBM / FP32-throughput bound CUDA-core kernels (direct (x_i - c)^2 and (x_i - x_j)^2 forms in the data's
// own precision: extrema and dispersions have no use for a Gram-form tensor-core contraction, whose cancellation error
// is largest exactly on the nearest pairs the Dunn index looks for).
#include "common.cuh"

namespace dic {
namespace {

constexpr int kScatterWarps = 8;
constexpr int kScatterMaxK = 64;

// one warp per row; per-warp private accumulators in shared memory (no atomics, fixed order => deterministic)
template <typename T>
__global__ void __launch_bounds__(kScatterWarps * 32)
cluster_scatter_kernel(const T* __restrict__ X, const int32_t* __restrict__ labels, const T* __restrict__ centers,
                       double* __restrict__ partial, int64_t N, int D, int K) {
  __shared__ double acc[kScatterWarps][kScatterMaxK][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kScatterWarps * kScatterMaxK * 2; i += blockDim.x) (&acc[0][0][0])[i] = 0.0;
  __syncthreads();
  const int64_t wstride = (int64_t)gridDim.x * kScatterWarps;
  for (int64_t row = (int64_t)blockIdx.x * kScatterWarps + warp; row < N; row += wstride) {
    const int k = labels[row];
    const T* x = X + row * D;
    const T* c = centers + (int64_t)k * D;
    T s = 0;
    for (int d = lane; d < D; d += 32) {
      const T df = x[d] - c[d];
      s += df * df;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && (unsigned)k < (unsigned)K) {
      acc[warp][k][0] += sqrt((double)s);
      acc[warp][k][1] += (double)s;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * 2; i += blockDim.x) {
    double t = 0.0;
    for (int w = 0; w < kScatterWarps; ++w) t += (&acc[w][0][0])[i];
    partial[(int64_t)blockIdx.x * K * 2 + i] = t;
  }
}

__global__ void cluster_scatter_finish_kernel(const double* __restrict__ partial, double* __restrict__ out, int nb,
                                              int K2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K2) return;
  double t = 0.0;
  for (int b = 0; b < nb; ++b) t += partial[(int64_t)b * K2 + i];
  out[i] = t;
}

// ---- Dunn extrema ----------------------------------------------------------------------------------------------
constexpr int kTile = 64, kDc = 16;
constexpr int kDunnMaxK = 32;

__device__ __forceinline__ unsigned long long dbits(double v) { return (unsigned long long)__double_as_longlong(v); }
constexpr unsigned long long kInfBits = 0x7ff0000000000000ULL;

// out: [K*K] nearest distance between the points of cluster a and cluster b (+inf: no pair seen; diagonal unused) | [1]
// largest distance between two points of one cluster
__global__ void dunn_init_kernel(double* out, int KK) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < KK) out[i] = __longlong_as_double((long long)kInfBits);
  if (i == KK) out[i] = 0.0;
}

template <typename T>
__global__ void __launch_bounds__(256)
dunn_minmax_kernel(const T* __restrict__ X, const int32_t* __restrict__ labels, double* __restrict__ out, int64_t N,
                   int D, int K, int64_t ntiles, int64_t npairs) {
  __shared__ T xi[kTile][kDc + 1], xj[kTile][kDc + 1];
  __shared__ int li[kTile], lj[kTile];
  __shared__ unsigned long long tab[kDunnMaxK * kDunnMaxK];     // per-block minima, as bit patterns of doubles >= 0
  __shared__ double red[8];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;      // thread = 4 x 4 block of the 64 x 64 tile
  for (int i = tid; i < K * K; i += 256) tab[i] = kInfBits;
  double bmax = 0.0;
  for (int64_t p = blockIdx.x; p < npairs; p += gridDim.x) {
    // p -> (ti, tj), ti <= tj, row-major over the upper triangle
    int64_t ti = (int64_t)((2.0 * ntiles + 1.0 - sqrt((2.0 * ntiles + 1.0) * (2.0 * ntiles + 1.0) - 8.0 * (double)p)) * 0.5);
    if (ti < 0) ti = 0;
    if (ti >= ntiles) ti = ntiles - 1;
    while (ti > 0 && ti * ntiles - ti * (ti - 1) / 2 > p) --ti;
    while ((ti + 1) * ntiles - (ti + 1) * ti / 2 <= p) ++ti;
    const int64_t tj = ti + (p - (ti * ntiles - ti * (ti - 1) / 2));
    const int64_t i0 = ti * kTile, j0 = tj * kTile;
    T acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0;
    __syncthreads();
    if (tid < kTile) {
      li[tid] = i0 + tid < N ? labels[i0 + tid] : -1;
      lj[tid] = j0 + tid < N ? labels[j0 + tid] : -1;
    }
    for (int d0 = 0; d0 < D; d0 += kDc) {
      __syncthreads();
      for (int e = tid; e < kTile * kDc; e += 256) {
        const int r = e / kDc, d = e - r * kDc;
        xi[r][d] = (i0 + r < N && d0 + d < D) ? X[(i0 + r) * D + d0 + d] : (T)0;
        xj[r][d] = (j0 + r < N && d0 + d < D) ? X[(j0 + r) * D + d0 + d] : (T)0;
      }
      __syncthreads();
#pragma unroll
      for (int d = 0; d < kDc; ++d) {
        T a4[4], b4[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) a4[a] = xi[ty * 4 + a][d];
#pragma unroll
        for (int b = 0; b < 4; ++b) b4[b] = xj[tx * 4 + b][d];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const T df = a4[a] - b4[b];
            acc[a][b] += df * df;
          }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int64_t gi = i0 + ty * 4 + a, gj = j0 + tx * 4 + b;
        if (gi < N && gj < N && gi < gj) {                 // each unordered pair once (the diagonal tile sees both orders)
          const double dist = sqrt((double)acc[a][b]);
          const int la = li[ty * 4 + a], lb = lj[tx * 4 + b];
          if (la == lb) bmax = fmax(bmax, dist);
          else {
            const int lo = min(la, lb), hi = max(la, lb);
            unsigned long long* cell = &tab[lo * K + hi];
            if (dbits(dist) < *reinterpret_cast<volatile unsigned long long*>(cell)) atomicMin(cell, dbits(dist));
          }
        }
      }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bmax = fmax(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
  if ((tid & 31) == 0) red[tid >> 5] = bmax;
  __syncthreads();
  // non-negative doubles order like their bit patterns; min / max commute, so the result is deterministic
  for (int i = tid; i < K * K; i += 256)
    if (tab[i] != kInfBits) atomicMin(reinterpret_cast<unsigned long long*>(out) + i, tab[i]);
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) bmax = fmax(bmax, red[w]);
    atomicMax(reinterpret_cast<unsigned long long*>(out + K * K), dbits(bmax));
  }
}

}  // namespace
}  // namespace dic

using namespace dic;

extern "C" size_t dic_cluster_scatter_workspace_bytes(int K) {
  return K > 0 ? (size_t)kColsumBlocks * (size_t)K * 2 * sizeof(double) : 0;
}

extern "C" int dic_cluster_scatter(const void* X, const int32_t* labels, const void* centers, double* out,
                                   void* workspace, int64_t N, int D, int K, int dtype, dic_stream_t stream) {
  DIC_REQUIRE(out && centers && ((X && labels && workspace) || N == 0), DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE(N >= 0 && D > 0 && K > 0 && K <= kScatterMaxK, DIC_ERR_INVALID_ARGUMENT,
              "bad sizes N=%lld D=%d K=%d (K <= %d)", (long long)N, D, K, kScatterMaxK);
  DIC_REQUIRE(dtype == 0 || dtype == 1, DIC_ERR_INVALID_ARGUMENT, "dtype must be 0 (float32) or 1 (float64)");
  cudaStream_t st = as_stream(stream);
  if (N == 0) {
    DIC_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * 2 * K, st));
    return DIC_OK;
  }
  int nb = (int)((N + kScatterWarps - 1) / kScatterWarps);
  if (nb > kColsumBlocks) nb = kColsumBlocks;
  double* partial = static_cast<double*>(workspace);
  if (dtype == 0)
    cluster_scatter_kernel<float><<<nb, kScatterWarps * 32, 0, st>>>(static_cast<const float*>(X), labels,
                                                                     static_cast<const float*>(centers), partial, N, D, K);
  else
    cluster_scatter_kernel<double><<<nb, kScatterWarps * 32, 0, st>>>(static_cast<const double*>(X), labels,
                                                                      static_cast<const double*>(centers), partial, N, D, K);
  DIC_LAUNCH_CHECK("cluster_scatter_kernel");
  cluster_scatter_finish_kernel<<<(2 * K + 127) / 128, 128, 0, st>>>(partial, out, nb, 2 * K);
  DIC_LAUNCH_CHECK("cluster_scatter_finish_kernel");
  return DIC_OK;
}

extern "C" int dic_dunn_minmax(const void* X, const int32_t* labels, double* out, int64_t N, int D, int K, int dtype,
                               dic_stream_t stream) {
  DIC_REQUIRE(out && ((X && labels) || N == 0), DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE(N >= 0 && D > 0 && K > 0, DIC_ERR_INVALID_ARGUMENT, "bad sizes N=%lld D=%d K=%d", (long long)N, D, K);
  DIC_REQUIRE(K <= kDunnMaxK, DIC_ERR_UNSUPPORTED, "K=%d exceeds the %d clusters the Dunn kernel covers", K, kDunnMaxK);
  DIC_REQUIRE(dtype == 0 || dtype == 1, DIC_ERR_INVALID_ARGUMENT, "dtype must be 0 (float32) or 1 (float64)");
  cudaStream_t st = as_stream(stream);
  dunn_init_kernel<<<(K * K + 1 + 255) / 256, 256, 0, st>>>(out, K * K);
  DIC_LAUNCH_CHECK("dunn_init_kernel");
  if (N < 2) return DIC_OK;
  const int64_t ntiles = (N + kTile - 1) / kTile;
  const int64_t npairs = ntiles * (ntiles + 1) / 2;
  const int nb = (int)(npairs < 148 * 4 ? npairs : 148 * 4);
  if (dtype == 0)
    dunn_minmax_kernel<float><<<nb, 256, 0, st>>>(static_cast<const float*>(X), labels, out, N, D, K, ntiles, npairs);
  else
    dunn_minmax_kernel<double><<<nb, 256, 0, st>>>(static_cast<const double*>(X), labels, out, N, D, K, ntiles, npairs);
  DIC_LAUNCH_CHECK("dunn_minmax_kernel");
  return DIC_OK;
}
