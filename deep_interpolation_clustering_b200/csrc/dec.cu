// DEC soft assignment / target distribution / KL for sm_100a.
//
// Reference: dec.py:49-63 (ClusterAssignment.forward), dec.py:66-76
// (target_distribution), clustering_interp.py:205-207 (kl_loss); closed-form gradients in
// SURVEY.md Appendix A.4.
//     d2_ij = sum_d (z_id - mu_jd)^2            (direct form, exactly as dec.py:56)
//     nu_ij = 1 / (1 + d2_ij / alpha),  n_ij = nu^((alpha+1)/2),  q_ij = n_ij / sum_j n_ij
//     f_j = sum_i q_ij,  p_ij = (q_ij^2 / f_j) / sum_j' (q_ij'^2 / f_j')
//
// HBM-bound (D*4 bytes per row in, K*4 out): one warp per latent row, persistent grid.
// Each lane owns float4 slices of the row, the K centres sit in shared memory, the K
// partial distances are reduced across the warp with a transposed butterfly (log2 K
// halving steps) so lane L ends up owning cluster L >> (5 - log2 KP).  K <= 16 and D <= 256
// never make a dense contraction worth a tensor-core GEMM (north_star); the
// ||z||^2 - 2 z.mu form would also lose the 1e-6 parity on q to cancellation.
#include <type_traits>

#include "common.cuh"

namespace dic {
namespace {

constexpr int kDecWarps = 8;
constexpr int kDecThreads = kDecWarps * 32;

template <int KP> struct Log2 { static constexpr int v = 1 + Log2<KP / 2>::v; };
template <> struct Log2<1> { static constexpr int v = 0; };

// sum / max over the KP lanes that own distinct clusters
template <int KP>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int off = 16; off >= 32 / KP && off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

struct RowCtx {
  float nu;     // 1 / (1 + d2/alpha) for this lane's cluster (0 for padded clusters)
  float q;      // soft assignment for this lane's cluster
  int k;        // this lane's cluster
  bool valid;   // k < K
};

// Loads one row (float4 slices into z4) and computes nu/q for the lane's cluster.
template <int KP, int NV>
__device__ __forceinline__ RowCtx dec_row(const float* __restrict__ zrow, const float4* __restrict__ smu4,
                                          float4 (&z4)[NV], int D4, int K, float alpha, int lane) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int l = lane + 32 * i;
    z4[i] = l < D4 ? __ldg(reinterpret_cast<const float4*>(zrow) + l) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float part[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    float s = 0.f;
    if (k < K) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int l = lane + 32 * i;
        if (l < D4) {
          const float4 m = smu4[k * D4 + l];
          const float a = z4[i].x - m.x, b = z4[i].y - m.y, c = z4[i].z - m.z, d = z4[i].w - m.w;
          s = fmaf(a, a, s);
          s = fmaf(b, b, s);
          s = fmaf(c, c, s);
          s = fmaf(d, d, s);
        }
      }
    }
    part[k] = s;
  }
  const float d2 = warp_reduce_multi<KP>(part, lane);
  RowCtx r;
  r.k = lane >> (5 - Log2<KP>::v);
  r.valid = r.k < K;
  r.nu = r.valid ? 1.0f / (1.0f + d2 / alpha) : 0.f;                      // dec.py:57
  const float power = (alpha + 1.0f) * 0.5f;                               // dec.py:58
  const float num = (alpha == 1.0f) ? r.nu : (r.valid ? powf(r.nu, power) : 0.f);   // dec.py:59-60
  r.q = num / group_sum<KP>(num);                                          // dec.py:61
  return r;
}

__device__ __forceinline__ void load_centres(float4* smu4, const float* mu, int K, int D4) {
  for (int i = threadIdx.x; i < K * D4; i += blockDim.x) smu4[i] = __ldg(reinterpret_cast<const float4*>(mu) + i);
}

template <int KP, int NV>
__global__ void __launch_bounds__(kDecThreads)
dec_q_fwd_kernel(const float* __restrict__ z, const float* __restrict__ mu, float* __restrict__ q,
                 int32_t* __restrict__ labels, double* __restrict__ ws_colsum, int64_t B, int D, int K,
                 float alpha) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* smu4 = reinterpret_cast<float4*>(smem_raw);
  __shared__ double s_col[kDecWarps][32];
  const int D4 = D >> 2;
  load_centres(smu4, mu, K, D4);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t wstride = (int64_t)gridDim.x * kDecWarps;
  constexpr int kShift = 5 - Log2<KP>::v;
  const bool writer = (lane & ((1 << kShift) - 1)) == 0;
  double colacc = 0.0;
  for (int64_t row = (int64_t)blockIdx.x * kDecWarps + warp; row < B; row += wstride) {
    float4 z4[NV];
    const RowCtx r = dec_row<KP, NV>(z + row * D, smu4, z4, D4, K, alpha, lane);
    if (writer && r.valid) q[row * K + r.k] = r.q;
    colacc += (double)r.q;
    if (labels) {
      // argmax_j q_ij, lowest index on ties
      float best = r.valid ? r.q : -1.f;
      int bi = r.k;
#pragma unroll
      for (int off = 16; off >= 32 / KP && off >= 1; off >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ob > best || (ob == best && oi < bi)) {
          best = ob;
          bi = oi;
        }
      }
      if (lane == 0) labels[row] = bi;
    }
  }
  if (ws_colsum) {
    s_col[warp][lane] = (writer && (lane >> kShift) < K) ? colacc : 0.0;
    __syncthreads();
    if (threadIdx.x < K) {
      double t = 0.0;
      for (int w = 0; w < kDecWarps; ++w) t += s_col[w][threadIdx.x << kShift];
      ws_colsum[(int64_t)blockIdx.x * K + threadIdx.x] = t;
    }
  }
}

__global__ void sum_blocks_f64_kernel(const double* __restrict__ ws, double* __restrict__ out, int nblocks,
                                      int cols) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // one warp per column
  if (c >= cols) return;
  double t = 0.0;
  for (int i = lane; i < nblocks; i += 32) t += ws[(int64_t)i * cols + c];
  t = warp_sum(t);
  if (lane == 0) out[c] = t;
}

// p from q and the (global) column sum: one thread per row.
__global__ void dec_p_kernel(const float* __restrict__ q, const double* __restrict__ colsum,
                             float* __restrict__ p, int64_t B, int K) {
  __shared__ float sinv[64];
  if (threadIdx.x < K) sinv[threadIdx.x] = (float)colsum[threadIdx.x];
  __syncthreads();
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= B) return;
  float s = 0.f;
  for (int k = 0; k < K; ++k) {
    const float qq = q[row * K + k];
    s += qq * qq / sinv[k];                   // dec.py:73
  }
  for (int k = 0; k < K; ++k) {
    const float qq = q[row * K + k];
    p[row * K + k] = (qq * qq / sinv[k]) / s;   // dec.py:74
  }
}

// Shared backward body.  MODE 0: upstream grad_q given.  MODE 1: KL(p||q) with p from colsum.
//   c_ij (MODE 0) = -((a+1)/a) q nu (g - <g,q>)
//   c_ij (MODE 1) =  ((a+1)/a) scale nu (p - q)
//   grad_z_i = (sum_j c_ij) z_i - sum_j c_ij mu_j
//   grad_mu_j = -(sum_i c_ij z_i) + (sum_i c_ij) mu_j
// Per-warp private accumulators for sum_i c_ij z_i live in shared memory (no atomics).
template <int KP, int NV, int MODE>
__global__ void __launch_bounds__(kDecThreads)
dec_bwd_kernel(const float* __restrict__ z, const float* __restrict__ mu, const float* __restrict__ grad_q,
               const double* __restrict__ colsum, float* __restrict__ p_out, float* __restrict__ grad_z,
               float* __restrict__ ws_dmu /*[grid][K*D + K + 1]*/, int64_t B, int D, int K, float alpha,
               float scale, int warps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int D4 = D >> 2;
  float4* smu4 = reinterpret_cast<float4*>(smem_raw);
  float4* sacc4 = smu4 + K * D4;                              // [warps][K][D4]
  float* scs = reinterpret_cast<float*>(sacc4 + (size_t)warps * K * D4);   // [warps][KP] sum_i c_ij
  double* skl = reinterpret_cast<double*>(scs + warps * 32);  // [warps]
  __shared__ float s_invf[32];
  load_centres(smu4, mu, K, D4);
  for (int i = threadIdx.x; i < warps * K * D4; i += blockDim.x) sacc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (MODE == 1 && threadIdx.x < 32) s_invf[threadIdx.x] = threadIdx.x < K ? (float)colsum[threadIdx.x] : 1.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t wstride = (int64_t)gridDim.x * warps;
  constexpr int kShift = 5 - Log2<KP>::v;
  const bool writer = (lane & ((1 << kShift) - 1)) == 0;
  const float coef = (alpha + 1.0f) / alpha;
  float4* my_acc = sacc4 + (size_t)warp * K * D4;
  float csum_acc = 0.f;   // sum_i c_ij for this lane's cluster
  double kl_acc = 0.0;
  for (int64_t row = (int64_t)blockIdx.x * warps + warp; row < B; row += wstride) {
    float4 z4[NV];
    const RowCtx r = dec_row<KP, NV>(z + row * D, smu4, z4, D4, K, alpha, lane);
    float c;
    if (MODE == 0) {
      const float g = r.valid ? __ldg(grad_q + row * K + r.k) : 0.f;
      const float gq = group_sum<KP>(g * r.q);
      c = -coef * r.q * r.nu * (g - gq);
    } else {
      const float w = r.valid ? r.q * r.q / s_invf[r.k] : 0.f;
      const float p = w / group_sum<KP>(w);
      if (p_out && writer && r.valid) p_out[row * K + r.k] = p;
      if (writer && r.valid && p > 0.f) kl_acc += (double)(p * (logf(p) - logf(r.q)));
      c = coef * scale * r.nu * (p - r.q);
    }
    if (!r.valid) c = 0.f;
    if (writer) csum_acc += c;
    const float ctot = group_sum<KP>(c);
    float4 dz[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i)
      dz[i] = make_float4(ctot * z4[i].x, ctot * z4[i].y, ctot * z4[i].z, ctot * z4[i].w);
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      if (k < K) {
        const float ck = __shfl_sync(0xffffffffu, c, k << kShift);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int l = lane + 32 * i;
          if (l < D4) {
            const float4 m = smu4[k * D4 + l];
            dz[i].x = fmaf(-ck, m.x, dz[i].x);
            dz[i].y = fmaf(-ck, m.y, dz[i].y);
            dz[i].z = fmaf(-ck, m.z, dz[i].z);
            dz[i].w = fmaf(-ck, m.w, dz[i].w);
            float4 a = my_acc[k * D4 + l];
            a.x = fmaf(ck, z4[i].x, a.x);
            a.y = fmaf(ck, z4[i].y, a.y);
            a.z = fmaf(ck, z4[i].z, a.z);
            a.w = fmaf(ck, z4[i].w, a.w);
            my_acc[k * D4 + l] = a;
          }
        }
      }
    }
    if (grad_z) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int l = lane + 32 * i;
        if (l < D4) reinterpret_cast<float4*>(grad_z + row * D)[l] = dz[i];
      }
    }
  }
  scs[warp * 32 + lane] = (writer && (lane >> kShift) < K) ? csum_acc : 0.f;
  kl_acc = warp_sum(kl_acc);
  if (lane == 0) skl[warp] = kl_acc;
  __syncthreads();
  // per-block partials: [K*D] sum_i c_ij z_i | [K] sum_i c_ij | [1] kl (as float pair hi/lo)
  float* out = ws_dmu + (int64_t)blockIdx.x * (K * D + K + 2);
  for (int i = threadIdx.x; i < K * D; i += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < warps; ++w) t += reinterpret_cast<const float*>(sacc4)[(size_t)w * K * D + i];
    out[i] = t;
  }
  if (threadIdx.x < K) {
    float t = 0.f;
    for (int w = 0; w < warps; ++w) t += scs[w * 32 + (threadIdx.x << kShift)];
    out[K * D + threadIdx.x] = t;
  }
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < warps; ++w) t += skl[w];
    const float hi = (float)t;
    out[K * D + K] = hi;
    out[K * D + K + 1] = (float)(t - (double)hi);
  }
}

// grad_mu[j,d] = -(sum_blocks S[j,d]) + (sum_blocks cs[j]) mu[j,d];  kl = sum_blocks
__global__ void dec_bwd_finish_kernel(const float* __restrict__ ws, const float* __restrict__ mu,
                                      float* __restrict__ grad_mu, double* __restrict__ kl_sum, int nblocks,
                                      int K, int D) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // one warp per element (+1 for kl)
  const int stride = K * D + K + 2;
  if (i < K * D) {
    const int j = i / D;
    double s = 0.0, cs = 0.0;
    for (int b = lane; b < nblocks; b += 32) {
      s += (double)ws[(int64_t)b * stride + i];
      cs += (double)ws[(int64_t)b * stride + K * D + j];
    }
    s = warp_sum(s);
    cs = warp_sum(cs);
    if (lane == 0) grad_mu[i] = (float)(-s + cs * (double)mu[i]);
  } else if (i == K * D && kl_sum) {
    double t = 0.0;
    for (int b = lane; b < nblocks; b += 32)
      t += (double)ws[(int64_t)b * stride + K * D + K] + (double)ws[(int64_t)b * stride + K * D + K + 1];
    t = warp_sum(t);
    if (lane == 0) *kl_sum = t;
  }
}

int g_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

constexpr int kDecMaxBlocks = 148 * 8;

int check(const void* z, const void* mu, int64_t B, int D, int K, float alpha) {
  DIC_REQUIRE((z || B == 0) && mu, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");   // empty batch: z may be NULL
  DIC_REQUIRE(B >= 0 && D > 0 && K > 0, DIC_ERR_INVALID_ARGUMENT, "bad sizes B=%lld D=%d K=%d", (long long)B, D, K);
  DIC_REQUIRE(alpha > 0.f, DIC_ERR_INVALID_ARGUMENT, "alpha must be positive (got %g)", (double)alpha);
  DIC_REQUIRE(K <= 32, DIC_ERR_UNSUPPORTED, "cluster_number <= 32 supported (got %d)", K);
  DIC_REQUIRE(D % 4 == 0 && D <= 1024, DIC_ERR_UNSUPPORTED,
              "embedding_dimension must be a multiple of 4 and <= 1024 (got %d)", D);
  DIC_REQUIRE(aligned16(z) && aligned16(mu), DIC_ERR_INVALID_ARGUMENT, "z and mu must be 16-byte aligned");
  return DIC_OK;
}

int pow2_at_least(int k) {
  int p = 1;
  while (p < k) p <<= 1;
  return p;
}

// Dispatch over (KP, NV): KP = K rounded up to a power of two, NV = ceil(D/128) in {1,2,4,8}.
template <int V> using IntC = std::integral_constant<int, V>;

template <int KP, typename F>
int dispatch_nv(int nv, F&& f) {
  if (nv <= 1) return f(IntC<KP>{}, IntC<1>{});
  if (nv <= 2) return f(IntC<KP>{}, IntC<2>{});
  if (nv <= 4) return f(IntC<KP>{}, IntC<4>{});
  return f(IntC<KP>{}, IntC<8>{});
}

template <typename F>
int dispatch_kp_nv(int kp, int nv, F&& f) {
  switch (kp) {
    case 1: return dispatch_nv<1>(nv, f);
    case 2: return dispatch_nv<2>(nv, f);
    case 4: return dispatch_nv<4>(nv, f);
    case 8: return dispatch_nv<8>(nv, f);
    case 16: return dispatch_nv<16>(nv, f);
    default: return dispatch_nv<32>(nv, f);
  }
}

int bwd_warps(int K, int D) {
  // mu + warps * K*D accumulators must fit in ~200 KB
  const size_t kd = (size_t)K * D * sizeof(float);
  int w = (int)((200 * 1024 - kd) / kd);
  if (w > kDecWarps) w = kDecWarps;
  return w;
}

size_t bwd_smem(int K, int D, int warps) {
  return (size_t)K * D * 4 * (1 + warps) + (size_t)warps * 32 * 4 + (size_t)warps * 8 + 16;
}

int launch_bwd(int mode, const float* z, const float* mu, const float* grad_q, const double* colsum,
               float* p_out, double* kl_sum, float* grad_z, float* grad_mu, void* workspace, int64_t B,
               int D, int K, float alpha, float scale, cudaStream_t st) {
  const int warps = bwd_warps(K, D);
  DIC_REQUIRE(warps >= 1, DIC_ERR_UNSUPPORTED, "K*D = %d too large for the DEC backward kernel", K * D);
  const size_t smem = bwd_smem(K, D, warps);
  int64_t want = (B + warps - 1) / warps;
  int blocks = (int)(want < (int64_t)g_sm_count() * 2 ? want : (int64_t)g_sm_count() * 2);
  if (blocks < 1) blocks = 1;
  if (blocks > kDecMaxBlocks) blocks = kDecMaxBlocks;
  float* ws = static_cast<float*>(workspace);
  const int kp = pow2_at_least(K), nv = (D + 127) / 128;
  int rc = dispatch_kp_nv(kp, nv, [&](auto kpc, auto nvc) -> int {
    constexpr int KP = decltype(kpc)::value, NV = decltype(nvc)::value;
    if (mode == 0) {
      auto kern = dec_bwd_kernel<KP, NV, 0>;
      if (smem > 48 * 1024)
        DIC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<blocks, warps * 32, smem, st>>>(z, mu, grad_q, colsum, p_out, grad_z, ws, B, D, K, alpha,
                                             scale, warps);
    } else {
      auto kern = dec_bwd_kernel<KP, NV, 1>;
      if (smem > 48 * 1024)
        DIC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<blocks, warps * 32, smem, st>>>(z, mu, grad_q, colsum, p_out, grad_z, ws, B, D, K, alpha,
                                             scale, warps);
    }
    return DIC_OK;
  });
  if (rc) return rc;
  DIC_LAUNCH_CHECK("dec_bwd_kernel");
  dec_bwd_finish_kernel<<<(K * D + 1 + 7) / 8, 256, 0, st>>>(ws, mu, grad_mu, kl_sum, blocks, K, D);
  DIC_LAUNCH_CHECK("dec_bwd_finish_kernel");
  return DIC_OK;
}

}  // namespace
}  // namespace dic

using namespace dic;

extern "C" size_t dic_dec_workspace_bytes(int K, int D) {
  if (K <= 0 || D <= 0) return 0;
  size_t a = (size_t)kDecMaxBlocks * K * sizeof(double);
  size_t b = (size_t)kDecMaxBlocks * ((size_t)K * D + K + 2) * sizeof(float);
  return (a > b ? a : b) + 256;
}

extern "C" int dic_dec_q_fwd(const float* z, const float* mu, float* q, int32_t* labels, double* colsum,
                             void* workspace, int64_t B, int D, int K, float alpha, dic_stream_t stream) {
  int rc = check(z, mu, B, D, K, alpha);
  if (rc) return rc;
  DIC_REQUIRE(q || B == 0, DIC_ERR_INVALID_ARGUMENT, "null output pointer");
  DIC_REQUIRE(!colsum || workspace, DIC_ERR_INVALID_ARGUMENT, "colsum requested without a workspace");
  cudaStream_t st = as_stream(stream);
  if (B == 0) {
    if (colsum) DIC_CUDA(cudaMemsetAsync(colsum, 0, sizeof(double) * K, st));
    return DIC_OK;
  }
  const size_t smem = (size_t)K * D * sizeof(float);
  int64_t want = (B + kDecWarps - 1) / kDecWarps;
  int blocks = (int)(want < (int64_t)g_sm_count() * 8 ? want : (int64_t)g_sm_count() * 8);
  if (blocks > kDecMaxBlocks) blocks = kDecMaxBlocks;
  double* ws = colsum ? static_cast<double*>(workspace) : nullptr;
  const int kp = pow2_at_least(K), nv = (D + 127) / 128;
  rc = dispatch_kp_nv(kp, nv, [&](auto kpc, auto nvc) -> int {
    constexpr int KP = decltype(kpc)::value, NV = decltype(nvc)::value;
    auto kern = dec_q_fwd_kernel<KP, NV>;
    if (smem > 48 * 1024)
      DIC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, kDecThreads, smem, st>>>(z, mu, q, labels, ws, B, D, K, alpha);
    return DIC_OK;
  });
  if (rc) return rc;
  DIC_LAUNCH_CHECK("dec_q_fwd_kernel");
  if (colsum) {
    sum_blocks_f64_kernel<<<(K + 7) / 8, 256, 0, st>>>(ws, colsum, blocks, K);
    DIC_LAUNCH_CHECK("sum_blocks_f64_kernel");
  }
  return DIC_OK;
}

extern "C" int dic_dec_p(const float* q, const double* colsum, float* p, int64_t B, int K,
                         dic_stream_t stream) {
  DIC_REQUIRE(((q && p) || B == 0) && colsum, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE(B >= 0 && K > 0, DIC_ERR_INVALID_ARGUMENT, "bad sizes B=%lld K=%d", (long long)B, K);
  DIC_REQUIRE(K <= 64, DIC_ERR_UNSUPPORTED, "cluster_number <= 64 supported (got %d)", K);
  if (B == 0) return DIC_OK;
  dec_p_kernel<<<(unsigned)((B + 255) / 256), 256, 0, as_stream(stream)>>>(q, colsum, p, B, K);
  DIC_LAUNCH_CHECK("dec_p_kernel");
  return DIC_OK;
}

extern "C" int dic_dec_q_bwd(const float* z, const float* mu, const float* grad_q, float* grad_z,
                             float* grad_mu, void* workspace, int64_t B, int D, int K, float alpha,
                             dic_stream_t stream) {
  int rc = check(z, mu, B, D, K, alpha);
  if (rc) return rc;
  DIC_REQUIRE((grad_q || B == 0) && grad_mu && workspace, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE(!grad_z || aligned16(grad_z), DIC_ERR_INVALID_ARGUMENT, "grad_z must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  if (B == 0) {
    DIC_CUDA(cudaMemsetAsync(grad_mu, 0, sizeof(float) * K * D, st));
    return DIC_OK;
  }
  return launch_bwd(0, z, mu, grad_q, nullptr, nullptr, nullptr, grad_z, grad_mu, workspace, B, D, K, alpha,
                    1.0f, st);
}

extern "C" int dic_dec_kl_fwd_bwd(const float* z, const float* mu, const double* colsum, float* p,
                                  double* kl_sum, float* grad_z, float* grad_mu, void* workspace, int64_t B,
                                  int D, int K, float alpha, float scale, dic_stream_t stream) {
  int rc = check(z, mu, B, D, K, alpha);
  if (rc) return rc;
  DIC_REQUIRE(colsum && grad_mu && workspace, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE(!grad_z || aligned16(grad_z), DIC_ERR_INVALID_ARGUMENT, "grad_z must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  if (B == 0) {
    DIC_CUDA(cudaMemsetAsync(grad_mu, 0, sizeof(float) * K * D, st));
    if (kl_sum) DIC_CUDA(cudaMemsetAsync(kl_sum, 0, sizeof(double), st));
    return DIC_OK;
  }
  return launch_bwd(1, z, mu, nullptr, colsum, p, kl_sum, grad_z, grad_mu, workspace, B, D, K, alpha, scale,
                    st);
}
