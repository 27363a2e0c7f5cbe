// k-means (Lloyd) / k-means++ / pairwise-distance "inertia" kernels for sm_100a.
//
// The reference delegates to scikit-learn (unpinned; 1.9.0 in the container):
//   KMeans.fit / fit_predict / predict      p2_clustering_optK.py:260,284,372,377,
//                                           clustering_trainer.py:75-82
//   pairwise_distances(X[a == c]) -> mean   p2_clustering_optK.py:334-351
// These kernels implement the same arithmetic (sklearn/cluster/_k_means_lloyd.pyx:23-218,
// _k_means_common.pyx:96-124, _kmeans.py:224-281) in the data's own dtype (float32 data,
// float64 reference draws) with float64 cross-block reductions.
//
// Lloyd pass = one read of X.  A CTA streams tiles of TILE rows through shared memory
// (coalesced global reads, odd 16-byte row stride => conflict-free 128-bit reads), phase 1
// gives each thread one row against all K centres (centres broadcast from shared memory, 4
// centres register-blocked), phase 2 re-maps threads to columns and adds the tile into
// per-CTA [K][D] sums (no atomics on the data path).  HBM-bound: N*D*sizeof(T) bytes/pass.

#include "common.cuh"

namespace dic {
// tensor-core path (pairwise_tc.cu)
size_t pairwise_tc_workspace_bytes(int64_t n, int D);
bool pairwise_tc_supported(const void* X, int D);
int launch_pairwise_tc(const float* X, double* out, void* workspace, int64_t n, int D, cudaStream_t st, int part,
                       int n_parts);
int launch_cluster_rowsums_tc(const float* X, const int32_t* perm, const int32_t* tile_cluster, double* rowsum,
                              void* workspace, int64_t n_pad, int D, int K, cudaStream_t st);

// kmeans_tc.cu
bool kmeans_tc_available();
bool kmeans_tc_covers(const void* X, int D, int K, int flags);
int launch_kmeans_assign_tc(const float* X, const float* centers, int32_t* labels, double* ws, int64_t N, int D, int K,
                            int flags, int want_sums, const double* done, int max_blocks, int* nb_out, cudaStream_t st);
bool kmeans_tc64_covers(const void* X, int D, int K, int flags);
int launch_kmeans_assign_tc64(const double* X, const double* centers, int32_t* labels, double* ws, int64_t N, int K,
                              int flags, int want_sums, const double* done, int max_blocks, int* nb_out, cudaStream_t st);

namespace {

template <typename T> struct Vec16;
template <> struct Vec16<float> { using type = float4; static constexpr int n = 4; };
template <> struct Vec16<double> { using type = double2; static constexpr int n = 2; };

__device__ __forceinline__ float dot16(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); return fmaf(a.w, b.w, acc);
}
__device__ __forceinline__ double dot16(const double2& a, const double2& b, double acc) {
  acc = fma(a.x, b.x, acc); return fma(a.y, b.y, acc);
}
__device__ __forceinline__ float sqd16(const float4& a, const float4& b, float acc) {
  const float p = a.x - b.x, q = a.y - b.y, r = a.z - b.z, s = a.w - b.w;
  acc = fmaf(p, p, acc); acc = fmaf(q, q, acc); acc = fmaf(r, r, acc); return fmaf(s, s, acc);
}
__device__ __forceinline__ double sqd16(const double2& a, const double2& b, double acc) {
  const double p = a.x - b.x, q = a.y - b.y;
  acc = fma(p, p, acc); return fma(q, q, acc);
}

// Row stride in 16-byte units: covers D elements, forced odd (conflict-free 128-bit reads).
template <typename T> static inline int row_stride16(int D) {
  const int per = 16 / (int)sizeof(T);
  return ((D + per - 1) / per) | 1;
}

struct KmLayout {
  size_t off_c, off_cn, off_acc, off_cnt, off_lab, off_x, off_x2, off_bar, total;
};
template <typename T> static KmLayout km_layout(int K, int D, int tile) {
  const int s16 = row_stride16<T>(D);
  KmLayout L;
  size_t o = 0;
  L.off_c = o;   o += (size_t)K * s16 * 16;
  {   // tile buffer, later reused to fold the per-group register sums: [tile/D groups][K][D]
    const size_t tile_bytes = (size_t)tile * s16 * 16, fold_bytes = (size_t)tile * K * sizeof(T);
    L.off_x = o;
    o += tile_bytes > fold_bytes ? tile_bytes : fold_bytes;
    o = (o + 15) & ~(size_t)15;
  }
  L.off_acc = o; o += (size_t)K * D * sizeof(T);
  o = (o + 15) & ~(size_t)15;
  L.off_cn = o;  o += (size_t)K * sizeof(T);
  o = (o + 15) & ~(size_t)15;
  L.off_cnt = o; o += (size_t)K * sizeof(int);
  L.off_lab = o; o += (size_t)tile * sizeof(int);
  o = (o + 15) & ~(size_t)15;
  L.off_bar = o; o += 16;                                   // two mbarriers
  L.off_x2 = 0;                                             // second tile buffer: see km_add_buffer
  L.total = (o + 15) & ~(size_t)15;
  return L;
}
template <typename T> static void km_add_buffer(KmLayout& L, int D, int tile) {
  L.off_x2 = L.total;
  L.total += (size_t)tile * row_stride16<T>(D) * 16;
}

template <typename T>
__global__ void kmeans_assign_kernel(const T* __restrict__ X, const T* __restrict__ centers,
                                     int32_t* __restrict__ labels, double* __restrict__ ws, int64_t N, int D,
                                     int K, int s16, int flags, KmLayout L, int want_sums,
                                     const double* __restrict__ done) {
  if (done && *done != 0.0) return;     // a batched Lloyd run has stopped: nothing left to do (see dic_kmeans_lloyd_run)
  using V = typename Vec16<T>::type;
  constexpr int PER = Vec16<T>::n;
  extern __shared__ __align__(16) unsigned char smem[];
  V* sc = reinterpret_cast<V*>(smem + L.off_c);
  V* sx = reinterpret_cast<V*>(smem + L.off_x);
  T* sacc = reinterpret_cast<T*>(smem + L.off_acc);
  T* scn = reinterpret_cast<T*>(smem + L.off_cn);
  int* scnt = reinterpret_cast<int*>(smem + L.off_cnt);
  int* slab = reinterpret_cast<int*>(smem + L.off_lab);
  const int tile = blockDim.x, tid = threadIdx.x;
  const int Dp = s16 * PER;

  // centres (zero padded to Dp), their squared norms, zeroed accumulators
  for (int i = tid; i < K * Dp; i += tile) {
    const int k = i / Dp, d = i - k * Dp;
    reinterpret_cast<T*>(sc)[i] = d < D ? centers[(int64_t)k * D + d] : T(0);
  }
  for (int i = tid; i < K * D; i += tile) sacc[i] = T(0);
  for (int i = tid; i < K; i += tile) scnt[i] = 0;
  __syncthreads();
  for (int k = tid; k < K; k += tile) {
    T s = T(0);
    for (int d = 0; d < D; ++d) {
      const T c = reinterpret_cast<T*>(sc)[k * Dp + d];
      s += c * c;
    }
    scn[k] = s;
  }
  __syncthreads();

  double inertia = 0.0, dist_sum = 0.0;
  int changed = 0;
  const int64_t ntiles = (N + tile - 1) / tile;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t row0 = t * tile;
    const int rows = (int)min((int64_t)tile, N - row0);
    // coalesced load of rows*D contiguous elements into padded shared rows
    const T* src = X + row0 * D;
    for (int i = tid; i < rows * D; i += tile) {
      const int r = i / D, d = i - r * D;
      reinterpret_cast<T*>(sx)[r * Dp + d] = src[i];
    }
    if (Dp != D)
      for (int i = tid; i < rows * (Dp - D); i += tile) {
        const int r = i / (Dp - D), d = D + (i - r * (Dp - D));
        reinterpret_cast<T*>(sx)[r * Dp + d] = T(0);
      }
    __syncthreads();

    if (tid < rows) {
      const V* xr = sx + tid * s16;
      int best = 0;
      if (flags & DIC_KM_KEEP_LABELS) {
        best = labels[row0 + tid];
      } else {
        T bestd = T(0);
        for (int kb = 0; kb < K; kb += 4) {
          const V* c0 = sc + min(kb + 0, K - 1) * s16;
          const V* c1 = sc + min(kb + 1, K - 1) * s16;
          const V* c2 = sc + min(kb + 2, K - 1) * s16;
          const V* c3 = sc + min(kb + 3, K - 1) * s16;
          T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
          for (int j = 0; j < s16; ++j) {
            const V xv = xr[j];
            a0 = dot16(xv, c0[j], a0);
            a1 = dot16(xv, c1[j], a1);
            a2 = dot16(xv, c2[j], a2);
            a3 = dot16(xv, c3[j], a3);
          }
          const T acc[4] = {a0, a1, a2, a3};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int k = kb + i;
            if (k < K) {
              const T dk = scn[k] - T(2) * acc[i];      // ||c||^2 - 2 x.c  (||x||^2 omitted)
              if (k == 0 || dk < bestd) {               // strict '<': lowest index wins ties
                bestd = dk;
                best = k;
              }
            }
          }
        }
        if (flags & DIC_KM_COUNT_CHANGES) changed += (labels[row0 + tid] != best);
        labels[row0 + tid] = best;
      }
      // direct squared distance to the chosen centre
      const V* cb = sc + best * s16;
      T d2 = T(0);
      for (int j = 0; j < s16; ++j) d2 = sqd16(xr[j], cb[j], d2);
      inertia += (double)d2;
      dist_sum += sqrt((double)d2);
      slab[tid] = best;
      if (want_sums) atomicAdd(&scnt[best], 1);
    }
    __syncthreads();
    if (want_sums) {
      for (int d = tid; d < D; d += tile) {
        for (int r = 0; r < rows; ++r) {
          const int k = slab[r];
          sacc[k * D + d] += reinterpret_cast<const T*>(sx)[r * Dp + d];
        }
      }
    }
    __syncthreads();
  }

  // per-block partials: [K*D] sums | [K] counts | inertia | changed | dist_sum | 0
  double* out = ws + (int64_t)blockIdx.x * ((int64_t)K * D + K + 4);
  if (want_sums) {
    for (int i = tid; i < K * D; i += tile) out[i] = (double)sacc[i];
    for (int i = tid; i < K; i += tile) out[(int64_t)K * D + i] = (double)scnt[i];
  }
  // block reduction of the three scalars (reuse the tile buffer)
  double* red = reinterpret_cast<double*>(sx);
  inertia = warp_sum(inertia);
  dist_sum = warp_sum(dist_sum);
  double ch = warp_sum((double)changed);
  const int warp = tid >> 5, lane = tid & 31, nwarps = (tile + 31) >> 5;
  if (lane == 0) {
    red[warp * 3 + 0] = inertia;
    red[warp * 3 + 1] = ch;
    red[warp * 3 + 2] = dist_sum;
  }
  __syncthreads();
  if (tid < 3) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += red[w * 3 + tid];
    out[(int64_t)K * D + K + tid] = s;
  }
  if (tid == 3) out[(int64_t)K * D + K + 3] = 0.0;
}

// Fast Lloyd pass for K <= 16 (the reference sweeps K = 2..10 and trains with K = 4): the per-CTA
// cluster sums live in REGISTERS for the whole kernel (thread = one column of one row group, a
// warp-uniform switch on the row's label picks the accumulator), tiles are loaded with 128-bit
// accesses and no per-element division.  One read of X, ~1.1k issue slots per 128-row tile.
constexpr int kKmTile = 128;

template <typename T, int KR, int NC>
__global__ void __launch_bounds__(kKmTile)
kmeans_assign_fast_kernel(const T* __restrict__ X, const T* __restrict__ centers, int32_t* __restrict__ labels,
                          double* __restrict__ ws, int64_t N, int D, int K, int s16, int flags,
                          KmLayout L, int want_sums, int vec_ok, const double* __restrict__ done) {
  if (done && *done != 0.0) return;
  using V = typename Vec16<T>::type;
  constexpr int PER = Vec16<T>::n;
  extern __shared__ __align__(16) unsigned char smem[];
  V* sc = reinterpret_cast<V*>(smem + L.off_c);
  V* sx0 = reinterpret_cast<V*>(smem + L.off_x);
  T* scn = reinterpret_cast<T*>(smem + L.off_cn);
  int* scnt = reinterpret_cast<int*>(smem + L.off_cnt);
  int* slab = reinterpret_cast<int*>(smem + L.off_lab);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  const int tid = threadIdx.x;
  const int Dp = s16 * PER;
  // Tile loads.  Fast path: every thread issues ONE 1-D TMA bulk copy (its row, D*sizeof(T) bytes)
  // into its padded shared row, completion on an mbarrier; with a second buffer the next tile
  // streams in while this one is processed.  Fallback: element-wise loads without division.
  const bool vec = vec_ok != 0;
  const bool two = vec && L.off_x2 != 0;
  V* sx1 = two ? reinterpret_cast<V*>(smem + L.off_x2) : sx0;

  for (int i = tid; i < K * Dp; i += kKmTile) {
    const int k = i / Dp, d = i - k * Dp;
    reinterpret_cast<T*>(sc)[i] = d < D ? centers[(int64_t)k * D + d] : T(0);
  }
  for (int i = tid; i < kKmTile * Dp; i += kKmTile) {          // pad columns stay 0 for good
    reinterpret_cast<T*>(sx0)[i] = T(0);
    if (two) reinterpret_cast<T*>(sx1)[i] = T(0);
  }
  for (int i = tid; i < K; i += kKmTile) scnt[i] = 0;
  if (vec && tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
  }
  __syncthreads();
  if (vec) fence_proxy_async();          // generic-proxy zero fill / barrier init before async-proxy writes
  for (int k = tid; k < K; k += kKmTile) {
    T s = T(0);
    for (int d = 0; d < D; ++d) {
      const T c = reinterpret_cast<T*>(sc)[k * Dp + d];
      s += c * c;
    }
    scn[k] = s;
  }
  __syncthreads();

  const uint32_t row_bytes = (uint32_t)D * (uint32_t)sizeof(T);
  const int64_t ntiles = (N + kKmTile - 1) / kKmTile;
  auto issue = [&](int64_t t, int buf) {
    const int64_t r0 = t * kKmTile;
    const int nr = (int)min((int64_t)kKmTile, N - r0);
    V* dst = buf ? sx1 : sx0;
    if (tid == 0) mbar_expect_tx(&bars[buf], (uint32_t)nr * row_bytes);
    if (tid < nr) bulk_g2s(dst + tid * s16, X + (r0 + tid) * D, row_bytes, &bars[buf]);
  };
  const int u_r0 = tid / D, u_j0 = tid - u_r0 * D;
  const int u_dr = kKmTile / D, u_dj = kKmTile - u_dr * D;
  // phase-2 ownership: column d of row group g (G groups), or NC columns when D > tile
  const int G = D < kKmTile ? kKmTile / D : 1;
  const int g = D < kKmTile ? tid / D : 0;
  const int dcol = D < kKmTile ? tid - g * D : tid;
  const bool owner = D < kKmTile ? (g < G) : true;
  T acc[KR][NC];
#pragma unroll
  for (int k = 0; k < KR; ++k)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[k][c] = T(0);

  double inertia = 0.0, dist_sum = 0.0;
  int changed = 0;
  if (two && (int64_t)blockIdx.x < ntiles) issue(blockIdx.x, 0);
  int iter = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++iter) {
    const int64_t row0 = t * kKmTile;
    const int rows = (int)min((int64_t)kKmTile, N - row0);
    const int buf = two ? (iter & 1) : 0;
    V* const sx = buf ? sx1 : sx0;
    if (vec) {
      if (!two) issue(t, 0);
      else if (t + gridDim.x < ntiles) issue(t + gridDim.x, buf ^ 1);     // prefetch the next tile
      mbar_wait(&bars[buf], two ? (uint32_t)((iter >> 1) & 1) : (uint32_t)(iter & 1));
    } else {
      int r = u_r0, j = u_j0;
      const T* src = X + row0 * D;
      for (int i = tid; i < rows * D; i += kKmTile) {
        reinterpret_cast<T*>(sx)[r * Dp + j] = src[i];
        r += u_dr; j += u_dj;
        if (j >= D) { j -= D; ++r; }
      }
      __syncthreads();
    }

    if (tid < rows) {
      const V* xr = sx + tid * s16;
      int best = 0;
      if (flags & DIC_KM_KEEP_LABELS) {
        best = labels[row0 + tid];
      } else {
        T bestd = T(0);
        for (int kb = 0; kb < K; kb += 4) {
          const V* c0 = sc + min(kb + 0, K - 1) * s16;
          const V* c1 = sc + min(kb + 1, K - 1) * s16;
          const V* c2 = sc + min(kb + 2, K - 1) * s16;
          const V* c3 = sc + min(kb + 3, K - 1) * s16;
          T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
          for (int j = 0; j < s16; ++j) {
            const V xv = xr[j];
            a0 = dot16(xv, c0[j], a0);
            a1 = dot16(xv, c1[j], a1);
            a2 = dot16(xv, c2[j], a2);
            a3 = dot16(xv, c3[j], a3);
          }
          const T av[4] = {a0, a1, a2, a3};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int k = kb + i;
            if (k < K) {
              const T dk = scn[k] - T(2) * av[i];       // ||c||^2 - 2 x.c  (||x||^2 omitted)
              if (k == 0 || dk < bestd) {               // strict '<': lowest index wins ties
                bestd = dk;
                best = k;
              }
            }
          }
        }
        if (flags & DIC_KM_COUNT_CHANGES) changed += (labels[row0 + tid] != best);
        labels[row0 + tid] = best;
      }
      const V* cb = sc + best * s16;
      T d2 = T(0);
      for (int j = 0; j < s16; ++j) d2 = sqd16(xr[j], cb[j], d2);
      inertia += (double)d2;
      dist_sum += sqrt((double)d2);
      slab[tid] = best;
      if (want_sums) atomicAdd(&scnt[best], 1);
    }
    __syncthreads();
    if (want_sums && owner) {
      for (int r = g; r < rows; r += G) {
        const int lab = slab[r];
        T xv[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int d = dcol + c * kKmTile;
          xv[c] = d < D ? reinterpret_cast<const T*>(sx)[r * Dp + d] : T(0);
        }
#define DIC_KM_CASE(k_)                                        \
  case k_:                                                     \
    if (KR > k_) {                                             \
      _Pragma("unroll") for (int c = 0; c < NC; ++c) acc[k_ < KR ? k_ : 0][c] += xv[c]; \
    }                                                          \
    break;
        switch (lab) {
          DIC_KM_CASE(0) DIC_KM_CASE(1) DIC_KM_CASE(2) DIC_KM_CASE(3) DIC_KM_CASE(4) DIC_KM_CASE(5)
          DIC_KM_CASE(6) DIC_KM_CASE(7) DIC_KM_CASE(8) DIC_KM_CASE(9) DIC_KM_CASE(10) DIC_KM_CASE(11)
          DIC_KM_CASE(12) DIC_KM_CASE(13) DIC_KM_CASE(14) DIC_KM_CASE(15)
          default: break;
        }
#undef DIC_KM_CASE
      }
    }
    __syncthreads();     // everyone is done with this buffer (and slab) before it is refilled
  }

  // per-block partials: [K*D] sums | [K] counts | inertia | changed | dist_sum | 0
  double* out = ws + (int64_t)blockIdx.x * ((int64_t)K * D + K + 4);
  if (want_sums) {
    // fold the G row groups through shared memory (the tile buffer is free now)
    T* fold = reinterpret_cast<T*>(sx0);                 // [G][K][D]
    if (owner) {
#pragma unroll
      for (int k = 0; k < KR; ++k)
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int d = dcol + c * kKmTile;
          if (k < K && d < D) fold[(g * K + k) * D + d] = acc[k][c];
        }
    }
    __syncthreads();
    for (int i = tid; i < K * D; i += kKmTile) {
      double sum = 0.0;
      for (int gg = 0; gg < G; ++gg) sum += (double)fold[gg * K * D + i];
      out[i] = sum;
    }
    for (int i = tid; i < K; i += kKmTile) out[(int64_t)K * D + i] = (double)scnt[i];
    __syncthreads();
  }
  double* red = reinterpret_cast<double*>(sx0);
  inertia = warp_sum(inertia);
  dist_sum = warp_sum(dist_sum);
  double ch = warp_sum((double)changed);
  const int warp = tid >> 5, lane = tid & 31;
  if (lane == 0) {
    red[warp * 3 + 0] = inertia;
    red[warp * 3 + 1] = ch;
    red[warp * 3 + 2] = dist_sum;
  }
  __syncthreads();
  if (tid < 3) {
    double sum = 0.0;
    for (int w = 0; w < kKmTile / 32; ++w) sum += red[w * 3 + tid];
    out[(int64_t)K * D + K + tid] = sum;
  }
  if (tid == 3) out[(int64_t)K * D + K + 3] = 0.0;
}

// One warp per output element: lanes stride over the per-block partials (independent loads in
// flight), then a fixed-order shuffle reduction -> deterministic and latency-tolerant.
__global__ void kmeans_finish_kernel(const double* __restrict__ ws, double* __restrict__ sums,
                                     double* __restrict__ counts, double* __restrict__ stats, int nblocks,
                                     int K, int D, const double* __restrict__ done) {
  if (done && *done != 0.0) return;
  const int64_t stride = (int64_t)K * D + K + 4;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // element of [sums | counts | stats]
  if (i >= K * D + K + 4) return;
  if (!sums && i < K * D + K) return;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += ws[b * stride + i];
  s = warp_sum(s);
  if (lane == 0) {
    if (i < K * D) sums[i] = s;
    else if (i < K * D + K) counts[i - K * D] = s;
    else stats[i - K * D - K] = s;
  }
}

// M-step on device (sklearn/cluster/_k_means_common.pyx:236-260 + _kmeans.py:703-733): new centre =
// sum / count (an empty cluster keeps its old centre; relocation is the host's rare path),
// squared centre shift, and a 4-double status [changed labels, sum shift^2, empty clusters,
// inertia] so the host reads ONE small buffer per iteration.
// run != 0 (dic_kmeans_lloyd_run): status has 8 doubles; status[4] is the stop flag every kernel of the batch
// checks first (1 = converged, 2 = an empty cluster needs the host), status[5] counts the iterations done,
// status[6] = 1 for strict convergence (no label changed), and the stopping rule of _kmeans.py:700-733 is applied
// here with tolerance `tol`.
template <typename T>
__global__ void kmeans_update_kernel(const double* __restrict__ sums, const double* __restrict__ counts,
                                     const double* __restrict__ stats, T* __restrict__ centers,
                                     double* __restrict__ status, int K, int D, int run, double tol) {
  if (run && status[4] != 0.0) return;
  __shared__ double red[32];
  __shared__ int s_empty;
  if (threadIdx.x == 0) {
    int e = 0;
    for (int k = 0; k < K; ++k) e += (counts[k] == 0.0);
    s_empty = e;
  }
  __syncthreads();
  // an empty cluster is the host's (rare) relocation path: leave the centres exactly as they are
  const bool frozen = s_empty > 0;
  double shift2 = 0.0;
  for (int i = threadIdx.x; i < K * D && !frozen; i += blockDim.x) {
    const int k = i / D;
    const double cnt = counts[k];
    if (cnt > 0.0) {
      const T nc = (T)(sums[i] / cnt);
      const double df = (double)nc - (double)centers[i];
      shift2 += df * df;
      centers[i] = nc;
    }
  }
  shift2 = warp_sum(shift2);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = shift2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    status[0] = stats[1];
    status[1] = s;
    status[2] = (double)s_empty;
    status[3] = stats[0];
    if (run) {
      status[5] += 1.0;
      if (s_empty > 0) status[4] = 2.0;
      else if (stats[1] == 0.0) { status[4] = 1.0; status[6] = 1.0; }
      else if (s <= tol) status[4] = 1.0;
    }
  }
}

// ---- k-means++ potentials -----------------------------------------------------------------
constexpr int kMaxCands = 16;
constexpr int kPotThreads = 256;
constexpr int kPotBlocks = 148 * 4;

template <typename T>
__global__ void __launch_bounds__(kPotThreads)
kmeans_min_d2_kernel(const T* __restrict__ X, const T* __restrict__ cands, const T* min_d2,
                     T* min_d2_out /* may alias min_d2: in place */, double* __restrict__ ws, int64_t N, int D, int L) {
  extern __shared__ __align__(16) unsigned char smem[];
  T* sc = reinterpret_cast<T*>(smem);                  // [L][D]
  for (int i = threadIdx.x; i < L * D; i += blockDim.x) sc[i] = cands[i];
  __syncthreads();
  // one warp per row: lanes stride over D (coalesced), shuffle-reduce the L distances
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  double pot[kMaxCands];
#pragma unroll
  for (int l = 0; l < kMaxCands; ++l) pot[l] = 0.0;
  for (int64_t row = (int64_t)blockIdx.x * wpb + warp; row < N; row += (int64_t)gridDim.x * wpb) {
    const T* xr = X + row * D;
    T part[kMaxCands];
#pragma unroll
    for (int l = 0; l < kMaxCands; ++l) part[l] = T(0);
    for (int d = lane; d < D; d += 32) {
      const T xv = xr[d];
#pragma unroll
      for (int l = 0; l < kMaxCands; ++l) {
        if (l < L) {
          const T df = xv - sc[l * D + d];
          part[l] += df * df;
        }
      }
    }
    const T prev = min_d2 ? min_d2[row] : T(0);
#pragma unroll
    for (int l = 0; l < kMaxCands; ++l) {
      if (l < L) {
        T v = part[l];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (min_d2 && prev < v) v = prev;
        if (lane == 0) {
          pot[l] += (double)v;
          if (min_d2_out && l == 0) min_d2_out[row] = v;
        }
      }
    }
  }
  __shared__ double red[kPotThreads / 32][kMaxCands];
  if (lane == 0)
    for (int l = 0; l < kMaxCands; ++l) red[warp][l] = pot[l];
  __syncthreads();
  if (threadIdx.x < L) {
    double s = 0.0;
    for (int w = 0; w < wpb; ++w) s += red[w][threadIdx.x];
    ws[(int64_t)blockIdx.x * L + threadIdx.x] = s;
  }
}

__global__ void sum_blocks_kernel(const double* __restrict__ ws, double* __restrict__ out, int nblocks, int cols) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // one warp per column
  if (c >= cols) return;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += ws[(int64_t)b * cols + c];
  s = warp_sum(s);
  if (lane == 0) out[c] = s;
}

// ---- pairwise Euclidean distance sum ----------------------------------------------------------
// 64 x 64 tile of the n x n distance matrix per step, upper triangle only (off-diagonal tiles
// count twice), 256 threads x (4 x 4) register tile, direct (x_i - x_j)^2 accumulation in the
// data dtype, sqrt, float64 running sum per thread.
constexpr int kPwTile = 64;
constexpr int kPwThreads = 256;
constexpr int kPwChunk = 32;        // feature columns staged per step
constexpr int kPwBlocks = 148 * 4;

template <typename T>
__global__ void __launch_bounds__(kPwThreads)
pairwise_sum_kernel(const T* __restrict__ X, double* __restrict__ ws, int64_t n, int D, int part, int n_parts) {
  __shared__ T sa[kPwChunk][kPwTile + 1];   // [d][row]: rows of the i-tile, transposed
  __shared__ T sb[kPwChunk][kPwTile + 1];
  const int tid = threadIdx.x;
  const int ti = tid / 16, tj = tid % 16;          // 16 x 16 threads, each a 4 x 4 sub-tile
  const int64_t nb = (n + kPwTile - 1) / kPwTile;
  const int64_t ntiles = nb * (nb + 1) / 2;
  double total = 0.0;
  // stripe `part` of `n_parts` (multi-GPU): tiles part, part + n_parts, ... of this CTA's sequence
  for (int64_t t = (int64_t)blockIdx.x * n_parts + part; t < ntiles; t += (int64_t)gridDim.x * n_parts) {
    // unrank t -> (bi <= bj) in the upper triangle, row-major
    int64_t bi = (int64_t)((2.0 * nb + 1.0 - sqrt((2.0 * nb + 1.0) * (2.0 * nb + 1.0) - 8.0 * (double)t)) * 0.5);
    while (bi * nb - bi * (bi - 1) / 2 > t) --bi;
    while ((bi + 1) * nb - (bi + 1) * bi / 2 <= t) ++bi;
    const int64_t bj = bi + (t - (bi * nb - bi * (bi - 1) / 2));
    const int64_t i0 = bi * kPwTile, j0 = bj * kPwTile;
    T acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = T(0);
    for (int d0 = 0; d0 < D; d0 += kPwChunk) {
      __syncthreads();
      for (int idx = tid; idx < kPwTile * kPwChunk; idx += kPwThreads) {
        const int r = idx / kPwChunk, d = idx % kPwChunk;
        const int64_t gi = i0 + r, gj = j0 + r;
        sa[d][r] = (gi < n && d0 + d < D) ? X[gi * D + d0 + d] : T(0);
        sb[d][r] = (gj < n && d0 + d < D) ? X[gj * D + d0 + d] : T(0);
      }
      __syncthreads();
#pragma unroll 8
      for (int d = 0; d < kPwChunk; ++d) {
        T av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) av[a] = sa[d][ti * 4 + a];
#pragma unroll
        for (int b = 0; b < 4; ++b) bv[b] = sb[d][tj * 4 + b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const T df = av[a] - bv[b];
            acc[a][b] += df * df;
          }
      }
    }
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int64_t gi = i0 + ti * 4 + a, gj = j0 + tj * 4 + b;
        if (gi < n && gj < n && gi != gj) s += sqrt((double)acc[a][b]);
      }
    total += (bi == bj) ? s : 2.0 * s;
  }
  __shared__ double red[kPwThreads / 32];
  total = warp_sum(total);
  if ((tid & 31) == 0) red[tid >> 5] = total;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < kPwThreads / 32; ++w) s += red[w];
    ws[blockIdx.x] = s;
  }
}

// ---- specialised tile kernel: rows of exactly S16 sixteen-byte vectors, K <= 16 --------------------------
// Same structure as kmeans_assign_fast_kernel (128-row tiles, thread = row for the E-step) with everything
// the compiler needs to unroll known at compile time, and a branch-free M-step:
//   * rows arrive by per-row 1-D TMA bulk copies into an odd-stride tile (conflict-free 128-bit reads);
//   * E-step: the row sits in registers (float32) and 4 centres at a time are read as shared-memory
//     broadcasts: K*D FMAs + K*S16 broadcast loads per row, nothing else in the loop;
//   * M-step: thread = (row group g, vector column): x += into sacc[g][label][column] in SHARED memory - the
//     label is an address, not a branch (the register-resident sums of the older kernel cost a 16-way switch
//     per row); each thread owns its slots, so there are no atomics and the order is fixed (deterministic);
//   * previous labels are fetched before the tile wait, inertia / distance sums only without NO_INERTIA.
// Instruction budget at K = 10, float32, D = 64: ~32 issue slots per row (the older kernel: 76).
constexpr int kT2Rows = 128;

template <typename T, int S16>
__global__ void __launch_bounds__(kT2Rows)
kmeans_assign_tile2_kernel(const T* __restrict__ X, const T* __restrict__ centers, int32_t* __restrict__ labels,
                           double* __restrict__ ws, int64_t N, int K, int flags, int want_sums,
                           const double* __restrict__ done) {
  if (done && *done != 0.0) return;
  using V = typename Vec16<T>::type;
  constexpr int PER = Vec16<T>::n;
  constexpr int D = S16 * PER;
  constexpr int RS = S16 + 1;                 // odd row stride (in vectors)
  constexpr int G = kT2Rows / S16;            // row groups of the M-step
  constexpr bool XREG = sizeof(T) == 4 && S16 <= 16;     // keep the row in registers during the E-step
  extern __shared__ __align__(16) unsigned char smem[];
  V* sx = reinterpret_cast<V*>(smem);                         // [128][RS]
  V* sc = sx + kT2Rows * RS;                                  // [K][S16]
  V* sacc = sc + K * S16;                                     // [G][K][S16]
  T* scn = reinterpret_cast<T*>(sacc + (want_sums ? G * K * S16 : 0));   // [16]
  int* scnt = reinterpret_cast<int*>(scn + 16);               // [G][16]
  int* slab = scnt + G * 16;                                  // [128]
  uint64_t* bar = reinterpret_cast<uint64_t*>(slab + kT2Rows);
  const int tid = threadIdx.x;

  for (int i = tid; i < K * D; i += kT2Rows) reinterpret_cast<T*>(sc)[i] = centers[i];
  if (want_sums)
    for (int i = tid; i < G * K * D; i += kT2Rows) reinterpret_cast<T*>(sacc)[i] = T(0);
  for (int i = tid; i < G * 16; i += kT2Rows) scnt[i] = 0;
  for (int i = tid; i < kT2Rows * RS * PER; i += kT2Rows) reinterpret_cast<T*>(sx)[i] = T(0);
  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  fence_proxy_async();
  if (tid < 16) {
    T sum = T(0);
    if (tid < K)
      for (int d = 0; d < D; ++d) {
        const T c = reinterpret_cast<T*>(sc)[tid * D + d];
        sum += c * c;
      }
    scn[tid] = sum;
  }
  __syncthreads();

  const bool keep = (flags & DIC_KM_KEEP_LABELS) != 0;
  const bool count_changes = (flags & DIC_KM_COUNT_CHANGES) != 0;
  const bool want_d2 = (flags & DIC_KM_NO_INERTIA) == 0;
  const int g = tid / S16, col = tid - g * S16;
  const int64_t ntiles = (N + kT2Rows - 1) / kT2Rows;
  constexpr uint32_t row_bytes = (uint32_t)(D * sizeof(T));
  double inertia = 0.0, dist_sum = 0.0;
  int changed = 0, iter = 0;

  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++iter) {
    const int64_t row0 = t * kT2Rows;
    const int rows = (int)min((int64_t)kT2Rows, N - row0);
    if (tid == 0) mbar_expect_tx(bar, (uint32_t)rows * row_bytes);
    __syncthreads();                          // expect_tx is armed; the previous M-step is done with sx / slab
    if (tid < rows) bulk_g2s(sx + tid * RS, X + (row0 + tid) * D, row_bytes, bar);
    int oldl = 0;
    if ((keep || count_changes) && tid < rows) oldl = labels[row0 + tid];
    mbar_wait(bar, (uint32_t)(iter & 1));

    if (tid < rows) {
      const V* xr = sx + tid * RS;
      V xreg[XREG ? S16 : 1];
      if (XREG) {
#pragma unroll
        for (int j = 0; j < S16; ++j) xreg[j] = xr[j];
      }
      int best = oldl;
      if (!keep) {
        T bestd = T(0);
        best = 0;
        for (int kb = 0; kb < K; kb += 4) {
          const V* c0 = sc + min(kb + 0, K - 1) * S16;
          const V* c1 = sc + min(kb + 1, K - 1) * S16;
          const V* c2 = sc + min(kb + 2, K - 1) * S16;
          const V* c3 = sc + min(kb + 3, K - 1) * S16;
          T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
#pragma unroll
          for (int j = 0; j < S16; ++j) {
            const V xv = XREG ? xreg[j] : xr[j];
            a0 = dot16(xv, c0[j], a0);
            a1 = dot16(xv, c1[j], a1);
            a2 = dot16(xv, c2[j], a2);
            a3 = dot16(xv, c3[j], a3);
          }
          const T av[4] = {a0, a1, a2, a3};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int k = kb + i;
            if (k < K) {
              const T dk = scn[k] - T(2) * av[i];       // ||c||^2 - 2 x.c  (||x||^2 omitted)
              if (k == 0 || dk < bestd) {               // strict '<': lowest index wins ties
                bestd = dk;
                best = k;
              }
            }
          }
        }
        if (count_changes) changed += (oldl != best);
        labels[row0 + tid] = best;
      }
      if (want_d2) {                          // direct squared distance to the chosen centre
        const V* cb = sc + best * S16;
        T d2 = T(0);
#pragma unroll
        for (int j = 0; j < S16; ++j) d2 = sqd16(XREG ? xreg[j] : xr[j], cb[j], d2);
        inertia += (double)d2;
        dist_sum += sqrt((double)d2);
      }
      slab[tid] = best;
    }
    if (want_sums) {
      __syncthreads();
      for (int r = g; r < rows; r += G) {
        const int lab = slab[r];
        const V xv = sx[r * RS + col];
        V* dst = sacc + (g * K + lab) * S16 + col;
        V a = *dst;
        if constexpr (sizeof(T) == 4) {
          a.x += xv.x; a.y += xv.y; a.z += xv.z; a.w += xv.w;
        } else {
          a.x += xv.x; a.y += xv.y;
        }
        *dst = a;
        if (col == 0) scnt[g * 16 + lab] += 1;
      }
    }
  }
  __syncthreads();

  // per-block partials: [K*D] sums | [K] counts | inertia | changed | dist_sum | 0
  double* out = ws + (int64_t)blockIdx.x * ((int64_t)K * D + K + 4);
  if (want_sums) {
    for (int i = tid; i < K * D; i += kT2Rows) {
      double sum = 0.0;
      for (int gg = 0; gg < G; ++gg) sum += (double)reinterpret_cast<const T*>(sacc)[gg * K * D + i];
      out[i] = sum;
    }
    for (int i = tid; i < K; i += kT2Rows) {
      int c = 0;
      for (int gg = 0; gg < G; ++gg) c += scnt[gg * 16 + i];
      out[(int64_t)K * D + i] = (double)c;
    }
  }
  __syncthreads();
  double* red = reinterpret_cast<double*>(sx);
  inertia = warp_sum(inertia);
  dist_sum = warp_sum(dist_sum);
  const double ch = warp_sum((double)changed);
  const int warp = tid >> 5, lane = tid & 31;
  if (lane == 0) {
    red[warp * 3 + 0] = inertia;
    red[warp * 3 + 1] = ch;
    red[warp * 3 + 2] = dist_sum;
  }
  __syncthreads();
  if (tid < 3) {
    double sum = 0.0;
    for (int w = 0; w < kT2Rows / 32; ++w) sum += red[w * 3 + tid];
    out[(int64_t)K * D + K + tid] = sum;
  }
  if (tid == 3) out[(int64_t)K * D + K + 3] = 0.0;
}

template <typename T>
size_t tile2_smem_bytes(int K, int S16, bool want_sums) {
  const size_t G = kT2Rows / S16;
  return 16 * ((size_t)kT2Rows * (S16 + 1) + (size_t)K * S16 + (want_sums ? G * K * S16 : 0)) + 16 * sizeof(T) +
         sizeof(int) * (G * 16 + kT2Rows) + 16;
}

// ---- row-per-half-warp Lloyd pass ------------------------------------------------------------------
// The streaming form of the E-step + accumulation: NO staging of X in shared memory.  Sixteen lanes own
// one row (lane l holds elements [l*E, l*E+E) as 16-byte vectors), so one warp-wide 128-bit load covers
// two adjacent rows and is perfectly coalesced; U = 4 such loads are in flight per warp.  The K dot
// products are reduced over the 16 lanes with a transposed butterfly (KP/2 + KP/4 + ... shuffles, lane l
// ends up owning centre l >> (4 - log2 KP)), the arg-min runs over the lanes (strict '<', lowest index on
// ties, like _k_means_lloyd.pyx), and every half-warp adds its row into ITS OWN [K][D] accumulator in
// shared memory (no atomics, fixed order => deterministic), folded in float64 at the end.
//   flags & DIC_KM_NO_INERTIA skips the direct ||x - c||^2 / ||x - c|| sums (not needed inside the loop).
constexpr int kRwThreads = 256;
constexpr int kRwHalves = kRwThreads / 16;
constexpr int kRwUnroll = 4;

template <typename T> struct RwVec;
template <> struct RwVec<float> { using type = float4; static constexpr int n = 4; };
template <> struct RwVec<double> { using type = double2; static constexpr int n = 2; };

__device__ __forceinline__ void rw_unpack(const float4& v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ void rw_unpack(const double2& v, double* o) { o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ float4 rw_pack(const float* o) { return make_float4(o[0], o[1], o[2], o[3]); }
__device__ __forceinline__ double2 rw_pack(const double* o) { return make_double2(o[0], o[1]); }
__device__ __forceinline__ float rw_inf(float) { return __int_as_float(0x7f800000); }
__device__ __forceinline__ double rw_inf(double) { return __longlong_as_double(0x7ff0000000000000LL); }

template <typename T, int E, int KP>
__global__ void __launch_bounds__(kRwThreads)
kmeans_assign_rw_kernel(const T* __restrict__ X, const T* __restrict__ centers, int32_t* __restrict__ labels,
                        double* __restrict__ ws, int64_t N, int D, int K, int flags, int want_sums,
                        const double* __restrict__ done) {
  if (done && *done != 0.0) return;
  using V = typename RwVec<T>::type;
  constexpr int PER = RwVec<T>::n;          // elements per 16-byte vector
  constexpr int NV = E / PER;               // vectors per lane
  constexpr int DP = 16 * E;                // padded row length
  constexpr int LOGKP = KP == 16 ? 4 : (KP == 8 ? 3 : 2);
  extern __shared__ __align__(16) unsigned char smem[];
  T* sc = reinterpret_cast<T*>(smem);                       // [K][DP] centres, zero beyond D
  T* sacc = sc + K * DP;                                    // [halves][K][DP] per-half-warp cluster sums
  T* scn = sacc + (want_sums ? kRwHalves * K * DP : 0);     // [KP] ||c||^2
  int* scnt = reinterpret_cast<int*>(scn + KP);             // [halves][KP] per-half-warp counts
  double* red = reinterpret_cast<double*>(smem);            // reused at the very end
  const int tid = threadIdx.x, lane = tid & 31, hl = lane & 15, half = lane >> 4;
  const int hw = tid >> 4;                                  // half-warp of this CTA

  for (int i = tid; i < K * DP; i += kRwThreads) {
    const int k = i / DP, d = i - k * DP;
    sc[i] = d < D ? centers[(int64_t)k * D + d] : T(0);
  }
  if (want_sums)
    for (int i = tid; i < kRwHalves * K * DP; i += kRwThreads) sacc[i] = T(0);
  for (int i = tid; i < kRwHalves * KP; i += kRwThreads) scnt[i] = 0;
  __syncthreads();
  for (int k = tid; k < KP; k += kRwThreads) {
    T sum = T(0);
    if (k < K)
      for (int d = 0; d < D; ++d) sum += sc[k * DP + d] * sc[k * DP + d];
    scn[k] = sum;
  }
  __syncthreads();

  const int k_own0 = hl >> (4 - LOGKP);
  const T cn_own = k_own0 < K ? scn[k_own0] : rw_inf(T(0));
  const bool lane_live = hl * E < D;                        // this lane's elements exist (D % PER == 0)
  const V* scv = reinterpret_cast<const V*>(sc) + hl * NV;
  V* saccv = reinterpret_cast<V*>(sacc + (int64_t)hw * K * DP) + hl * NV;
  int* mycnt = scnt + hw * KP;
  const bool keep = (flags & DIC_KM_KEEP_LABELS) != 0;
  const bool count_changes = (flags & DIC_KM_COUNT_CHANGES) != 0;
  const bool want_d2 = (flags & DIC_KM_NO_INERTIA) == 0;

  double inertia = 0.0, dist_sum = 0.0;
  int changed = 0;
  const int64_t warps_total = (int64_t)gridDim.x * (kRwThreads / 32);
  const int64_t warp_g = (int64_t)blockIdx.x * (kRwThreads / 32) + (tid >> 5);
  const int64_t npairs = (N + 1) / 2;                       // a warp step handles rows 2*pair, 2*pair + 1

  for (int64_t pair0 = warp_g; pair0 < npairs; pair0 += warps_total * kRwUnroll) {
    V xv[kRwUnroll][NV];
    int64_t rows[kRwUnroll];
    int oldl[kRwUnroll];
#pragma unroll
    for (int u = 0; u < kRwUnroll; ++u) {
      const int64_t pair = pair0 + (int64_t)u * warps_total;
      rows[u] = 2 * pair + half;
      // the previous label travels with the row (a dependent load per row would serialise the warp)
      oldl[u] = ((keep || count_changes) && pair < npairs && rows[u] < N) ? labels[rows[u]] : 0;
      const bool ok = pair < npairs && rows[u] < N && lane_live;
      const V* src = reinterpret_cast<const V*>(X + (ok ? rows[u] : 0) * D) + hl * NV;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        T z[PER];
#pragma unroll
        for (int e = 0; e < PER; ++e) z[e] = T(0);
        // a lane's vectors beyond D (ragged D) are zero; D % PER == 0 is guaranteed by the launcher
        xv[u][v] = (ok && (hl * E + v * PER) < D) ? __ldg(src + v) : rw_pack(z);
      }
    }
#pragma unroll
    for (int u = 0; u < kRwUnroll; ++u) {
      const int64_t pair = pair0 + (int64_t)u * warps_total;
      if (pair >= npairs) break;                            // warp-uniform
      const int64_t row = rows[u];
      const bool row_ok = row < N;                          // uniform per half-warp
      T x[E];
#pragma unroll
      for (int v = 0; v < NV; ++v) rw_unpack(xv[u][v], x + v * PER);
      int label;
      if (keep) {
        label = oldl[u];
      } else {
        T dot[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          dot[k] = T(0);
          if (k < K) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              T c[PER];
              rw_unpack(scv[k * (DP / PER) + v], c);
#pragma unroll
              for (int e = 0; e < PER; ++e) dot[k] += x[v * PER + e] * c[e];
            }
          }
        }
        // transposed butterfly over the 16 lanes of the row
        int nred = KP;
#pragma unroll
        for (int off = 8; off >= 1; off >>= 1) {
          if (nred > 1) {
            nred >>= 1;
            const bool up = (hl & off) != 0;
#pragma unroll
            for (int i = 0; i < KP / 2; ++i) {
              if (i < nred) {
                const T send = up ? dot[i] : dot[i + nred];
                const T keepv = up ? dot[i + nred] : dot[i];
                dot[i] = keepv + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
          } else {
            dot[0] += __shfl_xor_sync(0xffffffffu, dot[0], off);
          }
        }
        T dk = cn_own - T(2) * dot[0];                      // ||c||^2 - 2 x.c  (||x||^2 omitted); inf for k >= K
        int kb = k_own0;
#pragma unroll
        for (int off = 8; off >= (16 >> LOGKP); off >>= 1) {
          const T od = __shfl_xor_sync(0xffffffffu, dk, off);
          const int ok2 = __shfl_xor_sync(0xffffffffu, kb, off);
          if (od < dk || (od == dk && ok2 < kb)) {          // lowest index wins ties
            dk = od;
            kb = ok2;
          }
        }
        label = kb;
        if (hl == 0 && row_ok) {
          if (count_changes) changed += (oldl[u] != label);
          labels[row] = label;
        }
      }
      if (want_d2) {                                        // warp-uniform: the shuffles need all 32 lanes
        T d2 = T(0);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          T c[PER];
          rw_unpack(scv[label * (DP / PER) + v], c);
#pragma unroll
          for (int e = 0; e < PER; ++e) {
            const T df = x[v * PER + e] - c[e];
            d2 += df * df;
          }
        }
#pragma unroll
        for (int off = 8; off >= 1; off >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, off);
        if (hl == 0 && row_ok) {
          inertia += (double)d2;
          dist_sum += sqrt((double)d2);
        }
      }
      if (row_ok) {
        if (want_sums) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            T a[PER];
            rw_unpack(saccv[label * (DP / PER) + v], a);
#pragma unroll
            for (int e = 0; e < PER; ++e) a[e] += x[v * PER + e];
            saccv[label * (DP / PER) + v] = rw_pack(a);
          }
          if (hl == 0) mycnt[label] += 1;
        }
      }
    }
  }
  __syncthreads();

  // per-block partials: [K*D] sums | [K] counts | inertia | changed | dist_sum | 0
  double* out = ws + (int64_t)blockIdx.x * ((int64_t)K * D + K + 4);
  if (want_sums) {
    for (int i = tid; i < K * D; i += kRwThreads) {
      const int k = i / D, d = i - k * D;
      double sum = 0.0;
      for (int h = 0; h < kRwHalves; ++h) sum += (double)sacc[((int64_t)h * K + k) * DP + d];
      out[i] = sum;
    }
    for (int i = tid; i < K; i += kRwThreads) {
      int c = 0;
      for (int h = 0; h < kRwHalves; ++h) c += scnt[h * KP + i];
      out[(int64_t)K * D + i] = (double)c;
    }
  }
  __syncthreads();
  inertia = warp_sum(inertia);
  dist_sum = warp_sum(dist_sum);
  const double ch = warp_sum((double)changed);
  if (lane == 0) {
    red[(tid >> 5) * 3 + 0] = inertia;
    red[(tid >> 5) * 3 + 1] = ch;
    red[(tid >> 5) * 3 + 2] = dist_sum;
  }
  __syncthreads();
  if (tid < 3) {
    double sum = 0.0;
    for (int w = 0; w < kRwThreads / 32; ++w) sum += red[w * 3 + tid];
    out[(int64_t)K * D + K + tid] = sum;
  }
  if (tid == 3) out[(int64_t)K * D + K + 3] = 0.0;
}

// k-means++ potentials, streaming form (same row layout as kmeans_assign_rw_kernel): 16 lanes own a row, the
// L candidate distances are reduced with the transposed butterfly, lane hl ends up owning candidate
// hl >> (4 - log2 LP) and accumulates its potential in float64.  One read of X (+ min_d2), no staging.
template <typename T, int E, int LP>
__global__ void __launch_bounds__(kRwThreads)
kmeans_min_d2_rw_kernel(const T* __restrict__ X, const T* __restrict__ cands, const T* min_d2,
                        T* min_d2_out /* may alias min_d2: in place */, double* __restrict__ ws, int64_t N, int D, int L) {
  using V = typename RwVec<T>::type;
  constexpr int PER = RwVec<T>::n;
  constexpr int NV = E / PER;
  constexpr int DP = 16 * E;
  constexpr int LOGLP = LP == 16 ? 4 : (LP == 4 ? 2 : 0);
  extern __shared__ __align__(16) unsigned char smem[];
  T* sc = reinterpret_cast<T*>(smem);                       // [L][DP] candidates, zero beyond D
  __shared__ double red[kRwThreads / 32][16];
  const int tid = threadIdx.x, lane = tid & 31, hl = lane & 15, half = lane >> 4;
  for (int i = tid; i < L * DP; i += kRwThreads) {
    const int l = i / DP, d = i - l * DP;
    sc[i] = d < D ? cands[(int64_t)l * D + d] : T(0);
  }
  __syncthreads();
  const int l_own = hl >> (4 - LOGLP);
  const bool owner = (hl & ((16 >> LOGLP) - 1)) == 0 && l_own < L;    // one lane per (row, candidate)
  const bool lane_live = hl * E < D;
  const V* scv = reinterpret_cast<const V*>(sc) + hl * NV;
  double pot = 0.0;
  const int64_t warps_total = (int64_t)gridDim.x * (kRwThreads / 32);
  const int64_t warp_g = (int64_t)blockIdx.x * (kRwThreads / 32) + (tid >> 5);
  const int64_t npairs = (N + 1) / 2;
  for (int64_t pair0 = warp_g; pair0 < npairs; pair0 += warps_total * kRwUnroll) {
    V xv[kRwUnroll][NV];
    T prev[kRwUnroll];
    int64_t rows[kRwUnroll];
#pragma unroll
    for (int u = 0; u < kRwUnroll; ++u) {
      const int64_t pair = pair0 + (int64_t)u * warps_total;
      rows[u] = 2 * pair + half;
      const bool row_ok = pair < npairs && rows[u] < N;
      prev[u] = (min_d2 && row_ok) ? min_d2[rows[u]] : rw_inf(T(0));
      const V* src = reinterpret_cast<const V*>(X + (row_ok ? rows[u] : 0) * D) + hl * NV;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        T z[PER];
#pragma unroll
        for (int e = 0; e < PER; ++e) z[e] = T(0);
        xv[u][v] = (row_ok && lane_live && (hl * E + v * PER) < D) ? __ldg(src + v) : rw_pack(z);
      }
    }
#pragma unroll
    for (int u = 0; u < kRwUnroll; ++u) {
      const int64_t pair = pair0 + (int64_t)u * warps_total;
      if (pair >= npairs) break;                            // warp-uniform
      T x[E];
#pragma unroll
      for (int v = 0; v < NV; ++v) rw_unpack(xv[u][v], x + v * PER);
      T part[LP];
#pragma unroll
      for (int l = 0; l < LP; ++l) {
        part[l] = T(0);
        if (l < L) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            T c[PER];
            rw_unpack(scv[l * (DP / PER) + v], c);
#pragma unroll
            for (int e = 0; e < PER; ++e) {
              const T df = x[v * PER + e] - c[e];
              part[l] += df * df;
            }
          }
        }
      }
      int nred = LP;
#pragma unroll
      for (int off = 8; off >= 1; off >>= 1) {
        if (nred > 1) {
          nred >>= 1;
          const bool up = (hl & off) != 0;
#pragma unroll
          for (int i = 0; i < (LP > 1 ? LP / 2 : 1); ++i) {
            if (i < nred) {
              const T send = up ? part[i] : part[i + nred];
              const T keepv = up ? part[i + nred] : part[i];
              part[i] = keepv + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
        } else {
          part[0] += __shfl_xor_sync(0xffffffffu, part[0], off);
        }
      }
      T v = part[0];
      if (prev[u] < v) v = prev[u];
      if (owner && rows[u] < N) {
        pot += (double)v;
        if (min_d2_out && l_own == 0) min_d2_out[rows[u]] = v;
      }
    }
  }
  // lanes that own candidate l: fixed-order sum over the block
  if (lane < 16) {
#pragma unroll
    for (int l = 0; l < 16; ++l) red[tid >> 5][l] = 0.0;
  }
  __syncwarp();
  const double both = pot + __shfl_xor_sync(0xffffffffu, pot, 16);      // the two rows of the warp step
  if (half == 0 && owner) red[tid >> 5][l_own] = both;
  __syncthreads();
  if (tid < L) {
    double sum = 0.0;
    for (int w = 0; w < kRwThreads / 32; ++w) sum += red[w][tid];
    ws[(int64_t)blockIdx.x * L + tid] = sum;
  }
}

template <typename T>
size_t rw_smem_bytes(int K, int E, int KP, bool want_sums) {
  const size_t dp = 16 * (size_t)E;
  return sizeof(T) * (K * dp + (want_sums ? (size_t)kRwHalves * K * dp : 0) + KP) + sizeof(int) * kRwHalves * KP + 64;
}

int km_blocks(int K, int D) {
  int64_t per = (int64_t)K * D + K + 4;
  int64_t b = (8LL << 20) / per;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 148) b = 148;
  return (int)b;
}

// Where the tensor-core pass beats the CUDA-core kernels (measured per pass under a CUDA graph, benchmarks/_km_small.py
// and _km_pass.py: 80 against 96-148 us at 1M x 64, 136 against 370-810 us at 500k x 256; below ~20k rows both run at
// their launch + ramp floor of 8-10 us).
static bool kmeans_tc_wins(int64_t N, int D, int K) {
  if (D == 256) return N >= 4096;
  if (D == 128) return N >= 50000 || (K >= 8 && N >= 8192);
  if (D == 64) return N >= 100000 || (K >= 10 && N >= 20000);
  return false;
}

// float64 rows (D = 64): 108-157 us against 166-290 us of the CUDA-core kernel at 1M rows; at the launch floor below ~20k.
static bool kmeans_tc64_wins(int64_t N, int K) { return N >= 50000 || (K >= 10 && N >= 16384); }

template <typename T>
int launch_assign(const void* X, const void* centers, int32_t* labels, double* sums, double* counts,
                  double* stats, void* workspace, int64_t N, int D, int K, int flags, cudaStream_t st,
                  const double* done = nullptr) {
  // bits 8..11 of flags: 0 = pick by shape and measured cost; 1..4 = that kernel or DIC_ERR_UNSUPPORTED (parity tests
  // and benchmarks address every kernel through the ABI; nothing here reads the environment)
  const int which = (flags >> 8) & 15;
  DIC_REQUIRE(which <= 6, DIC_ERR_INVALID_ARGUMENT, "DIC_KM_KERNEL(%d): unknown kernel", which);
  // tensor-core E-step (kmeans_tc.cu): float32 rows of 64 / 128 / 256 elements in the Lloyd-loop form of the pass
  if constexpr (sizeof(T) == 4) {
    const bool tc_ok = kmeans_tc_covers(X, D, K, flags);
    DIC_REQUIRE(tc_ok || which != 5, DIC_ERR_UNSUPPORTED, "DIC_KM_KERNEL(5): the tensor-core pass covers float32, D in "
                "{64,128,256}, K <= 16 with DIC_KM_NO_INERTIA and without DIC_KM_KEEP_LABELS (got D=%d K=%d flags=%d)", D, K,
                flags & 255);
    if (tc_ok && (which == 5 || (which == 0 && kmeans_tc_wins(N, D, K) && kmeans_tc_available()))) {
      double* wsd = static_cast<double*>(workspace);
      int nb = 0;
      int rc = launch_kmeans_assign_tc(static_cast<const float*>(X), static_cast<const float*>(centers), labels, wsd, N, D, K,
                                       flags, sums != nullptr, done, km_blocks(K, D), &nb, st);
      if (rc) return rc;
      const int nn = K * D + K + 4;
      kmeans_finish_kernel<<<(nn + 7) / 8, 256, 0, st>>>(wsd, sums, counts, stats, nb, K, D, done);
      DIC_LAUNCH_CHECK("kmeans_finish_kernel");
      return DIC_OK;
    }
    DIC_REQUIRE(which != 6, DIC_ERR_UNSUPPORTED, "DIC_KM_KERNEL(6): the float64 tensor-core pass is float64 only");
  } else {
    DIC_REQUIRE(which != 5, DIC_ERR_UNSUPPORTED, "DIC_KM_KERNEL(5): the tensor-core pass is float32 only");
    // float64 rows of 64 elements (the gap statistic's reference sets): tensor-core screen + exact rows (kmeans_tc.cu)
    const bool tc_ok = kmeans_tc64_covers(X, D, K, flags);
    DIC_REQUIRE(tc_ok || which != 6, DIC_ERR_UNSUPPORTED, "DIC_KM_KERNEL(6): the float64 tensor-core pass covers D = 64, "
                "K <= 16 with DIC_KM_NO_INERTIA and without DIC_KM_KEEP_LABELS (got D=%d K=%d flags=%d)", D, K, flags & 255);
    if (tc_ok && (which == 6 || (which == 0 && kmeans_tc64_wins(N, K) && kmeans_tc_available()))) {
      double* wsd = static_cast<double*>(workspace);
      int nb = 0;
      int rc = launch_kmeans_assign_tc64(static_cast<const double*>(X), static_cast<const double*>(centers), labels, wsd, N,
                                         K, flags, sums != nullptr, done, km_blocks(K, D), &nb, st);
      if (rc) return rc;
      const int nn = K * D + K + 4;
      kmeans_finish_kernel<<<(nn + 7) / 8, 256, 0, st>>>(wsd, sums, counts, stats, nb, K, D, done);
      DIC_LAUNCH_CHECK("kmeans_finish_kernel");
      return DIC_OK;
    }
  }
  // specialised tile kernel: rows of exactly 16, 32 or 64 sixteen-byte vectors (D = 64 / 128 / 256 float32,
  // 32 / 64 / 128 float64).  64 vectors (the reference's latent width, D = 256) leave room for ONE resident CTA per
  // SM only, still 9x the general kernel that served this shape before (3.6 ms per pass at 500k x 256, K = 10).
  {
    constexpr int PER = 16 / (int)sizeof(T);
    const int s16x = D / PER;
    const bool shape_ok = D % PER == 0 && (s16x == 16 || s16x == 32 || s16x == 64) && K <= 16 && aligned16(X);
    const size_t smem = shape_ok ? tile2_smem_bytes<T>(K, s16x, sums != nullptr) : 0;
    const bool fits = shape_ok && smem <= (s16x == 64 ? 200 : 110) * 1024;
    DIC_REQUIRE(fits || which != 1, DIC_ERR_UNSUPPORTED, "DIC_KM_KERNEL(1): the specialised tile kernel does not cover "
                "K=%d D=%d", K, D);
    if (fits && (which == 0 || which == 1)) {
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      int per_sm = (int)((size_t)220 * 1024 / (smem + 1024));
      per_sm = per_sm > 6 ? 6 : (per_sm < 1 ? 1 : per_sm);
      int nb = sms * per_sm;
      const int cap = km_blocks(K, D);
      if (nb > cap) nb = cap;
      const int64_t nt = (N + kT2Rows - 1) / kT2Rows;
      if (nt < nb) nb = (int)nt;
      if (nb < 1) nb = 1;
      double* wsd = static_cast<double*>(workspace);
#define DIC_T2_LAUNCH(S16_)                                                                                   \
  {                                                                                                           \
    auto kf = kmeans_assign_tile2_kernel<T, S16_>;                                                            \
    if (smem > 48 * 1024)                                                                                     \
      DIC_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
    kf<<<nb, kT2Rows, smem, st>>>(static_cast<const T*>(X), static_cast<const T*>(centers), labels, wsd, N, K, \
                                  flags, sums != nullptr, done);                                              \
  }
      if (s16x == 16) DIC_T2_LAUNCH(16) else if (s16x == 32) DIC_T2_LAUNCH(32) else DIC_T2_LAUNCH(64)
#undef DIC_T2_LAUNCH
      DIC_LAUNCH_CHECK("kmeans_assign_tile2_kernel");
      const int nn = K * D + K + 4;
      kmeans_finish_kernel<<<(nn + 7) / 8, 256, 0, st>>>(wsd, sums, counts, stats, nb, K, D, done);
      DIC_LAUNCH_CHECK("kmeans_finish_kernel");
      return DIC_OK;
    }
  }
  // streaming row-per-half-warp kernel: K <= 16, rows of whole 16-byte vectors, <= 16 elements per lane
  {
    constexpr int PER = 16 / (int)sizeof(T);
    const int e_need = (D + 15) / 16;
    const int E = e_need <= 4 ? 4 : (e_need <= 8 ? 8 : 16);
    const int KP = K <= 4 ? 4 : (K <= 8 ? 8 : 16);
    const size_t smem = rw_smem_bytes<T>(K, E, KP, sums != nullptr);
    // measured on B200 at 1M x 64 (benchmarks/_km_pass.py): the streaming kernel wins for K <= 8 in float32
    // (0.15-0.19 ms vs 0.21-0.23 ms per pass) and K <= 4 in float64; beyond that its per-row shuffle
    // reductions cost more issue slots than the tile kernel's shared-memory staging
    const bool rw_wins = sizeof(T) == 4 ? K <= 8 : K <= 4;
    const bool rw_ok = K <= 16 && D <= 256 && D % PER == 0 && aligned16(X) && smem <= 100 * 1024;
    DIC_REQUIRE(rw_ok || which != 2, DIC_ERR_UNSUPPORTED, "DIC_KM_KERNEL(2): the streaming kernel does not cover K=%d D=%d",
                K, D);
    if (rw_ok && ((which == 0 && rw_wins) || which == 2)) {
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const int per_sm = smem <= 36 * 1024 ? 4 : (smem <= 56 * 1024 ? 3 : 2);      // resident CTAs (8 warps each)
      int nb = sms * per_sm;
      const int cap = km_blocks(K, D);                 // the workspace holds this many per-block partials
      if (nb > cap) nb = cap;
      const int64_t want = (N + 2 * (kRwThreads / 32) * kRwUnroll - 1) / (2 * (kRwThreads / 32) * kRwUnroll);
      if (want < nb) nb = (int)want;
      if (nb < 1) nb = 1;
      double* wsd = static_cast<double*>(workspace);
#define DIC_RW_LAUNCH(E_, KP_)                                                                               \
  {                                                                                                          \
    auto kf = kmeans_assign_rw_kernel<T, E_, KP_>;                                                           \
    if (smem > 48 * 1024)                                                                                    \
      DIC_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
    kf<<<nb, kRwThreads, smem, st>>>(static_cast<const T*>(X), static_cast<const T*>(centers), labels, wsd,  \
                                     N, D, K, flags, sums != nullptr, done);                                 \
  }
#define DIC_RW_KP(E_)                                                                                        \
  if (KP == 4) DIC_RW_LAUNCH(E_, 4) else if (KP == 8) DIC_RW_LAUNCH(E_, 8) else DIC_RW_LAUNCH(E_, 16)
      if (E == 4) { DIC_RW_KP(4) } else if (E == 8) { DIC_RW_KP(8) } else { DIC_RW_KP(16) }
#undef DIC_RW_KP
#undef DIC_RW_LAUNCH
      DIC_LAUNCH_CHECK("kmeans_assign_rw_kernel");
      const int nn = K * D + K + 4;
      kmeans_finish_kernel<<<(nn + 7) / 8, 256, 0, st>>>(wsd, sums, counts, stats, nb, K, D, done);
      DIC_LAUNCH_CHECK("kmeans_finish_kernel");
      return DIC_OK;
    }
  }
  const int s16 = row_stride16<T>(D);
  int tile = 128;
  KmLayout L = km_layout<T>(K, D, tile);
  const bool fast_ok = K <= 16 && D <= 4 * kKmTile && L.total <= 100 * 1024;
  DIC_REQUIRE(fast_ok || which != 3, DIC_ERR_UNSUPPORTED, "DIC_KM_KERNEL(3): the generic tile kernel does not cover "
              "K=%d D=%d", K, D);
  if (fast_ok && which != 4) {
    const int64_t nt = (N + kKmTile - 1) / kKmTile;
    int nb = km_blocks(K, D);
    if (nt < nb) nb = (int)nt;
    if (nb < 1) nb = 1;
    double* wsd = static_cast<double*>(workspace);
    {
      KmLayout L2 = L;
      km_add_buffer<T>(L2, D, kKmTile);
      if (L2.total <= 72 * 1024) L = L2;          // double-buffer when >= 3 CTAs per SM still fit
    }
    const int kr = K <= 4 ? 4 : (K <= 8 ? 8 : 16);
    const int nc = D <= kKmTile ? 1 : (D <= 2 * kKmTile ? 2 : 4);
#define DIC_KM_LAUNCH(KR_, NC_)                                                                          \
  {                                                                                                      \
    auto kf = kmeans_assign_fast_kernel<T, KR_, NC_>;                                                    \
    if (L.total > 48 * 1024)                                                                             \
      DIC_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));     \
    kf<<<nb, kKmTile, L.total, st>>>(static_cast<const T*>(X), static_cast<const T*>(centers), labels,   \
                                     wsd, N, D, K, s16, flags, L, sums != nullptr,                       \
                                     (int)(aligned16(X) && D % (16 / (int)sizeof(T)) == 0), done);       \
  }
#define DIC_KM_NC(KR_)                                                                \
  if (nc == 1) DIC_KM_LAUNCH(KR_, 1) else if (nc == 2) DIC_KM_LAUNCH(KR_, 2) else DIC_KM_LAUNCH(KR_, 4)
    if (kr == 4) { DIC_KM_NC(4) } else if (kr == 8) { DIC_KM_NC(8) } else { DIC_KM_NC(16) }
#undef DIC_KM_NC
#undef DIC_KM_LAUNCH
    DIC_LAUNCH_CHECK("kmeans_assign_fast_kernel");
    const int nn = K * D + K + 4;
    kmeans_finish_kernel<<<(nn + 7) / 8, 256, 0, st>>>(wsd, sums, counts, stats, nb, K, D, done);
    DIC_LAUNCH_CHECK("kmeans_finish_kernel");
    return DIC_OK;
  }
  while (tile > 32 && L.total > 100 * 1024) {
    tile >>= 1;
    L = km_layout<T>(K, D, tile);
  }
  DIC_REQUIRE(L.total <= (size_t)kMaxSmemBytes, DIC_ERR_UNSUPPORTED,
              "K=%d D=%d needs %zu bytes of shared memory (limit %d)", K, D, L.total, kMaxSmemBytes);
  auto kern = kmeans_assign_kernel<T>;
  if (L.total > 48 * 1024)
    DIC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  const int64_t ntiles = (N + tile - 1) / tile;
  int blocks = km_blocks(K, D);
  if (ntiles < blocks) blocks = (int)ntiles;
  if (blocks < 1) blocks = 1;
  double* ws = static_cast<double*>(workspace);
  kern<<<blocks, tile, L.total, st>>>(static_cast<const T*>(X), static_cast<const T*>(centers), labels, ws,
                                      N, D, K, s16, flags, L, sums != nullptr, done);
  DIC_LAUNCH_CHECK("kmeans_assign_kernel");
  const int n = K * D + K + 4;
  kmeans_finish_kernel<<<(n + 7) / 8, 256, 0, st>>>(ws, sums, counts, stats, blocks, K, D, done);
  DIC_LAUNCH_CHECK("kmeans_finish_kernel");
  return DIC_OK;
}

template <typename T>
int launch_min_d2(const void* X, const void* cands, const void* min_d2, void* min_d2_out, double* pots,
                  void* workspace, int64_t N, int D, int L, cudaStream_t st) {
  {   // streaming half-warp-per-row kernel: rows of whole 16-byte vectors, <= 16 elements per lane
    constexpr int PER = 16 / (int)sizeof(T);
    const int e_need = (D + 15) / 16;
    const int E = e_need <= 4 ? 4 : (e_need <= 8 ? 8 : 16);
    if (D <= 256 && D % PER == 0 && aligned16(X) && L <= 16) {
      const size_t smem_rw = (size_t)L * 16 * E * sizeof(T);
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      int blocks = sms * 4;
      if (blocks > kPotBlocks) blocks = kPotBlocks;
      const int64_t want = (N + 2 * (kRwThreads / 32) * kRwUnroll - 1) / (2 * (kRwThreads / 32) * kRwUnroll);
      if (want < blocks) blocks = (int)want;
      if (blocks < 1) blocks = 1;
      double* wsd = static_cast<double*>(workspace);
      const int LP = L == 1 ? 1 : (L <= 4 ? 4 : 16);
#define DIC_MD_LAUNCH(E_, LP_)                                                                                 \
  kmeans_min_d2_rw_kernel<T, E_, LP_><<<blocks, kRwThreads, smem_rw, st>>>(                                    \
      static_cast<const T*>(X), static_cast<const T*>(cands), static_cast<const T*>(min_d2),                   \
      static_cast<T*>(min_d2_out), wsd, N, D, L);
#define DIC_MD_LP(E_)                                                                                          \
  if (LP == 1) { DIC_MD_LAUNCH(E_, 1) } else if (LP == 4) { DIC_MD_LAUNCH(E_, 4) } else { DIC_MD_LAUNCH(E_, 16) }
      if (E == 4) { DIC_MD_LP(4) } else if (E == 8) { DIC_MD_LP(8) } else { DIC_MD_LP(16) }
#undef DIC_MD_LP
#undef DIC_MD_LAUNCH
      DIC_LAUNCH_CHECK("kmeans_min_d2_rw_kernel");
      sum_blocks_kernel<<<(L + 7) / 8, 256, 0, st>>>(wsd, pots, blocks, L);
      DIC_LAUNCH_CHECK("sum_blocks_kernel");
      return DIC_OK;
    }
  }
  const size_t smem = (size_t)L * D * sizeof(T);
  auto kern = kmeans_min_d2_kernel<T>;
  if (smem > 48 * 1024)
    DIC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t want = (N + kPotThreads / 32 - 1) / (kPotThreads / 32);
  int blocks = (int)(want < kPotBlocks ? want : kPotBlocks);
  if (blocks < 1) blocks = 1;
  double* ws = static_cast<double*>(workspace);
  kern<<<blocks, kPotThreads, smem, st>>>(static_cast<const T*>(X), static_cast<const T*>(cands),
                                          static_cast<const T*>(min_d2), static_cast<T*>(min_d2_out), ws, N,
                                          D, L);
  DIC_LAUNCH_CHECK("kmeans_min_d2_kernel");
  sum_blocks_kernel<<<(L + 7) / 8, 256, 0, st>>>(ws, pots, blocks, L);
  DIC_LAUNCH_CHECK("sum_blocks_kernel");
  return DIC_OK;
}

template <typename T>
int launch_pairwise(const void* X, double* out, void* workspace, int64_t n, int D, cudaStream_t st, int part,
                    int n_parts) {
  const int64_t nb = (n + kPwTile - 1) / kPwTile;
  const int64_t ntiles = nb * (nb + 1) / 2;
  int blocks = (int)(ntiles < kPwBlocks ? ntiles : kPwBlocks);
  if (blocks < 1) blocks = 1;
  double* ws = static_cast<double*>(workspace);
  pairwise_sum_kernel<T><<<blocks, kPwThreads, 0, st>>>(static_cast<const T*>(X), ws, n, D, part, n_parts);
  DIC_LAUNCH_CHECK("pairwise_sum_kernel");
  sum_blocks_kernel<<<1, 256, 0, st>>>(ws, out, blocks, 1);
  DIC_LAUNCH_CHECK("sum_blocks_kernel");
  return DIC_OK;
}

}  // namespace
}  // namespace dic

using namespace dic;

extern "C" size_t dic_kmeans_workspace_bytes(int K, int D) {
  if (K <= 0 || D <= 0) return 0;
  // the block count shrinks as K grows, so the product is not monotone in K: a buffer sized for K must also serve
  // every smaller K' (a gap sweep reuses one buffer for K = 2..k_max)
  size_t a = 0;
  for (int k = 1; k <= K; ++k) {
    const size_t ak = (size_t)km_blocks(k, D) * ((size_t)k * D + k + 4) * sizeof(double);
    a = ak > a ? ak : a;
  }
  size_t b = (size_t)kPotBlocks * kMaxCands * sizeof(double);
  return (a > b ? a : b) + 256;
}

extern "C" int dic_kmeans_assign(const void* X, const void* centers, int32_t* labels, double* sums,
                                 double* counts, double* stats, void* workspace, int64_t N, int D, int K,
                                 int dtype, int flags, dic_stream_t stream) {
  DIC_REQUIRE(X && centers && labels && stats && workspace, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE((sums == nullptr) == (counts == nullptr), DIC_ERR_INVALID_ARGUMENT,
              "sums and counts must be given together");
  DIC_REQUIRE(N >= 0 && D > 0 && K > 0, DIC_ERR_INVALID_ARGUMENT, "bad sizes N=%lld D=%d K=%d", (long long)N, D, K);
  DIC_REQUIRE(K <= 64 && D <= 512, DIC_ERR_UNSUPPORTED, "k-means supports K <= 64, D <= 512 (got K=%d D=%d)", K, D);
  DIC_REQUIRE(dtype == 0 || dtype == 1, DIC_ERR_INVALID_ARGUMENT, "dtype must be 0 (float32) or 1 (float64)");
  cudaStream_t st = as_stream(stream);
  if (N == 0) {
    if (sums) {
      DIC_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * K * D, st));
      DIC_CUDA(cudaMemsetAsync(counts, 0, sizeof(double) * K, st));
    }
    DIC_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 4, st));
    return DIC_OK;
  }
  return dtype == 0 ? launch_assign<float>(X, centers, labels, sums, counts, stats, workspace, N, D, K, flags, st)
                    : launch_assign<double>(X, centers, labels, sums, counts, stats, workspace, N, D, K, flags, st);
}

extern "C" int dic_kmeans_update(const double* sums, const double* counts, const double* stats, void* centers,
                                 double* status, int D, int K, int dtype, dic_stream_t stream) {
  DIC_REQUIRE(sums && counts && stats && centers && status, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE(D > 0 && K > 0, DIC_ERR_INVALID_ARGUMENT, "bad sizes D=%d K=%d", D, K);
  DIC_REQUIRE(dtype == 0 || dtype == 1, DIC_ERR_INVALID_ARGUMENT, "dtype must be 0 (float32) or 1 (float64)");
  cudaStream_t st = as_stream(stream);
  if (dtype == 0)
    kmeans_update_kernel<float><<<1, 256, 0, st>>>(sums, counts, stats, static_cast<float*>(centers), status, K, D, 0, 0.0);
  else
    kmeans_update_kernel<double><<<1, 256, 0, st>>>(sums, counts, stats, static_cast<double*>(centers), status, K, D, 0, 0.0);
  DIC_LAUNCH_CHECK("kmeans_update_kernel");
  return DIC_OK;
}

extern "C" int dic_kmeans_lloyd_step(const void* X, void* centers, int32_t* labels, double* sums, double* counts,
                                     double* stats, double* status, void* workspace, int64_t N, int D, int K,
                                     int dtype, int flags, dic_stream_t stream) {
  DIC_REQUIRE(sums && counts && status, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  int rc = dic_kmeans_assign(X, centers, labels, sums, counts, stats, workspace, N, D, K, dtype, flags, stream);
  if (rc) return rc;
  return dic_kmeans_update(sums, counts, stats, centers, status, D, K, dtype, stream);
}

extern "C" int dic_kmeans_lloyd_run(const void* X, void* centers, int32_t* labels, double* sums, double* counts,
                                    double* stats, double* status8, void* workspace, int64_t N, int D, int K,
                                    int dtype, int flags, int n_steps, double tol, dic_stream_t stream) {
  DIC_REQUIRE(X && centers && labels && sums && counts && stats && status8 && workspace, DIC_ERR_INVALID_ARGUMENT,
              "null pointer argument");
  DIC_REQUIRE(N > 0 && D > 0 && K > 0 && n_steps > 0, DIC_ERR_INVALID_ARGUMENT, "bad sizes N=%lld D=%d K=%d steps=%d",
              (long long)N, D, K, n_steps);
  DIC_REQUIRE(K <= 64 && D <= 512, DIC_ERR_UNSUPPORTED, "k-means supports K <= 64, D <= 512 (got K=%d D=%d)", K, D);
  DIC_REQUIRE(dtype == 0 || dtype == 1, DIC_ERR_INVALID_ARGUMENT, "dtype must be 0 (float32) or 1 (float64)");
  cudaStream_t st = as_stream(stream);
  const double* done = status8 + 4;
  for (int i = 0; i < n_steps; ++i) {
    int rc = dtype == 0 ? launch_assign<float>(X, centers, labels, sums, counts, stats, workspace, N, D, K, flags, st, done)
                        : launch_assign<double>(X, centers, labels, sums, counts, stats, workspace, N, D, K, flags, st, done);
    if (rc) return rc;
    if (dtype == 0)
      kmeans_update_kernel<float><<<1, 256, 0, st>>>(sums, counts, stats, static_cast<float*>(centers), status8, K, D, 1, tol);
    else
      kmeans_update_kernel<double><<<1, 256, 0, st>>>(sums, counts, stats, static_cast<double*>(centers), status8, K, D, 1, tol);
    DIC_LAUNCH_CHECK("kmeans_update_kernel");
  }
  return DIC_OK;
}

extern "C" int dic_kmeans_min_d2(const void* X, const void* cands, const void* min_d2, void* min_d2_out,
                                 double* pots, void* workspace, int64_t N, int D, int L, int dtype,
                                 dic_stream_t stream) {
  DIC_REQUIRE(X && cands && pots && workspace, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE(N > 0 && D > 0 && L > 0, DIC_ERR_INVALID_ARGUMENT, "bad sizes N=%lld D=%d L=%d", (long long)N, D, L);
  DIC_REQUIRE(L <= kMaxCands, DIC_ERR_UNSUPPORTED, "at most %d candidates per call (got %d)", kMaxCands, L);
  DIC_REQUIRE(!min_d2_out || L == 1, DIC_ERR_INVALID_ARGUMENT, "min_d2_out requires exactly one candidate");
  DIC_REQUIRE(dtype == 0 || dtype == 1, DIC_ERR_INVALID_ARGUMENT, "dtype must be 0 (float32) or 1 (float64)");
  cudaStream_t st = as_stream(stream);
  return dtype == 0 ? launch_min_d2<float>(X, cands, min_d2, min_d2_out, pots, workspace, N, D, L, st)
                    : launch_min_d2<double>(X, cands, min_d2, min_d2_out, pots, workspace, N, D, L, st);
}

extern "C" size_t dic_pairwise_workspace_bytes(int64_t n, int D) {
  const size_t a = (size_t)kPwBlocks * sizeof(double) + 256;
  const size_t b = pairwise_tc_workspace_bytes(n < 0 ? 0 : n, D);
  return a > b ? a : b;
}

extern "C" int dic_cluster_rowsums(const float* X, const int32_t* perm, const int32_t* tile_cluster,
                                   double* rowsum, void* workspace, int64_t n_pad, int D, int K,
                                   dic_stream_t stream) {
  DIC_REQUIRE(X && perm && tile_cluster && rowsum && workspace, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE(n_pad > 0 && n_pad % 128 == 0, DIC_ERR_INVALID_ARGUMENT,
              "n_pad=%lld must be a positive multiple of 128 (clusters padded to whole tiles)", (long long)n_pad);
  DIC_REQUIRE(K > 0, DIC_ERR_INVALID_ARGUMENT, "K=%d", K);
  DIC_REQUIRE(D > 0 && D <= 256 && D % 4 == 0 && aligned16(X), DIC_ERR_UNSUPPORTED,
              "the tensor-core row-sum kernel needs D <= 256, D %% 4 == 0 and 16-byte aligned rows (got D=%d)", D);
  return launch_cluster_rowsums_tc(X, perm, tile_cluster, rowsum, workspace, n_pad, D, K, as_stream(stream));
}

static int pairwise_dist_sum_impl(const void* Xc, double* out, void* workspace, int64_t n, int D, int dtype, int part,
                                  int n_parts, dic_stream_t stream) {
  DIC_REQUIRE(Xc && out && workspace, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE(n >= 0 && D > 0, DIC_ERR_INVALID_ARGUMENT, "bad sizes n=%lld D=%d", (long long)n, D);
  const bool force_exact = (dtype & DIC_PAIRWISE_EXACT) != 0;
  dtype &= ~DIC_PAIRWISE_EXACT;
  DIC_REQUIRE(dtype == 0 || dtype == 1, DIC_ERR_INVALID_ARGUMENT, "dtype must be 0 (float32) or 1 (float64)");
  DIC_REQUIRE(n_parts >= 1 && n_parts <= 1024 && part >= 0 && part < n_parts, DIC_ERR_INVALID_ARGUMENT,
              "bad stripe %d of %d", part, n_parts);
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    DIC_CUDA(cudaMemsetAsync(out, 0, sizeof(double), st));
    return DIC_OK;
  }
  // float32 clusters of useful size go to the tensor-core kernel (split float16 / TF32 operands, float32-grade dots; the
  // caller centres the cluster); dtype | DIC_PAIRWISE_EXACT forces the direct (x_i - x_j)^2 CUDA-core kernel in the
  // data's own precision, float64 always uses it.
  if (dtype == 0 && !force_exact && n >= 512 && pairwise_tc_supported(Xc, D))
    return launch_pairwise_tc(static_cast<const float*>(Xc), out, workspace, n, D, st, part, n_parts);
  return dtype == 0 ? launch_pairwise<float>(Xc, out, workspace, n, D, st, part, n_parts)
                    : launch_pairwise<double>(Xc, out, workspace, n, D, st, part, n_parts);
}

extern "C" int dic_pairwise_dist_sum(const void* Xc, double* out, void* workspace, int64_t n, int D, int dtype,
                                     dic_stream_t stream) {
  return pairwise_dist_sum_impl(Xc, out, workspace, n, D, dtype, 0, 1, stream);
}

extern "C" int dic_pairwise_dist_sum_part(const void* Xc, double* out, void* workspace, int64_t n, int D, int dtype,
                                          int part, int n_parts, dic_stream_t stream) {
  return pairwise_dist_sum_impl(Xc, out, workspace, n, D, dtype, part, n_parts, stream);
}
