// CrossChannelInterp forward / backward for sm_100a.
//
// Reference: interpolation_layer.py:99-127; closed-form backward in SURVEY.md Appendix A.2.
// Per encounter (planar rows y, w, y' of length R for each of C vitals):
//     what[c,r] = softmax over c of w[.,r]
//     ybar[c]   = mean over r of y[c,.]
//     z[c',r]   = sum_c what[c,r] (y[c,r] - ybar[c]) K[c,c'] + ybar[c']
//     out       = [ z | exp(w) | y' - z ]
// HBM-bound (reads and writes 12 C R bytes per encounter): one CTA per encounter, one
// thread per reference point, the C values of a point live in registers, every global
// access is coalesced along r, the only cross-thread step is the mean over r.
#include "common.cuh"

namespace dic {
namespace {

constexpr int kCciThreads = 128;

// Block-wide sums of MAXC per-thread values; result broadcast to all threads via smem.
template <int MAXC>
__device__ __forceinline__ void block_sum_vec(float (&v)[MAXC], int C, float* red /*[warps][MAXC]*/,
                                              float* bcast /*[MAXC]*/) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    if (c < C) {
      const float s = warp_sum(v[c]);
      if (lane == 0) red[warp * MAXC + c] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x < C) {
    float s = 0.f;
    for (int w = 0; w < nwarps; ++w) s += red[w * MAXC + threadIdx.x];
    bcast[threadIdx.x] = s;
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
    if (c < C) v[c] = bcast[c];
  __syncthreads();
}

template <int MAXC>
__global__ void __launch_bounds__(kCciThreads)
cci_fwd_kernel(const float* __restrict__ u, const float* __restrict__ kernel, float* __restrict__ out,
               int C, int R) {
  __shared__ float sK[MAXC * MAXC];
  __shared__ float red[(kCciThreads / 32) * MAXC];
  __shared__ float bcast[MAXC];
  const int64_t b = blockIdx.x;
  const float* ub = u + b * (int64_t)(3 * C) * R;
  float* ob = out + b * (int64_t)(3 * C) * R;
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) sK[i] = __ldg(kernel + i);

  // pass 1: ybar[c] = mean_r y[c, r]
  float ybar[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) ybar[c] = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < C) ybar[c] += ub[c * R + r];
  }
  block_sum_vec<MAXC>(ybar, C, red, bcast);
  const float invR = 1.0f / (float)R;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) ybar[c] *= invR;

  // pass 2 (rows are L1/L2 hot)
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    float a[MAXC], w[MAXC];
    float wmax = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < C) {
        w[c] = ub[(C + c) * R + r];
        wmax = fmaxf(wmax, w[c]);
      }
    }
    const float shift = (wmax == -INFINITY) ? 0.f : wmax;   // torch.logsumexp guards an all -inf row
    float den = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < C) {
        a[c] = expf(w[c] - shift);
        den += a[c];
      }
    }
    const float inv_den = 1.0f / den;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < C) {
        a[c] = a[c] * inv_den * (ub[c * R + r] - ybar[c]);
        ob[(C + c) * R + r] = expf(w[c]);                    // intensity, :104
      }
    }
#pragma unroll
    for (int cp = 0; cp < MAXC; ++cp) {
      if (cp < C) {
        float z = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
          if (c < C) z = fmaf(a[c], sK[c * C + cp], z);
        z += ybar[cp];
        ob[cp * R + r] = z;
        ob[(2 * C + cp) * R + r] = ub[(2 * C + cp) * R + r] - z;   // transient minus smooth, :122-123
      }
    }
  }
}

// Backward (Appendix A.2).  Per encounter:
//   gzt = gz - gT;  uu[c,r] = sum_c' K[c,c'] gzt[c',r]
//   dK[c,c'] += what[c,r] (y[c,r]-ybar[c]) gzt[c',r]
//   dy[c,r]  = what uu - mean_r(what uu) + mean_r(gzt[c,.])
//   dw[c,r]  = what[c,r] ( (y-ybar) uu - sum_c'' what (y-ybar) uu ) + exp(w) gI
//   dy'[c,r] = gT
template <int MAXC>
__global__ void __launch_bounds__(kCciThreads)
cci_bwd_kernel(const float* __restrict__ u, const float* __restrict__ kernel,
               const float* __restrict__ grad_out, float* __restrict__ grad_u,
               float* __restrict__ partial /*(B, C*C)*/, int C, int R) {
  __shared__ float sK[MAXC * MAXC];
  __shared__ float red[(kCciThreads / 32) * MAXC];
  __shared__ float bcast[MAXC];
  __shared__ float sdK[(kCciThreads / 32) * MAXC * MAXC];   // one slice per warp: no atomics
  const int64_t b = blockIdx.x;
  const float* ub = u + b * (int64_t)(3 * C) * R;
  const float* gb = grad_out + b * (int64_t)(3 * C) * R;
  float* dub = grad_u + b * (int64_t)(3 * C) * R;
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) sK[i] = __ldg(kernel + i);
  for (int i = threadIdx.x; i < (kCciThreads / 32) * MAXC * MAXC; i += blockDim.x) sdK[i] = 0.f;
  float* my_dK = sdK + (threadIdx.x >> 5) * MAXC * MAXC;

  float ybar[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) ybar[c] = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < C) ybar[c] += ub[c * R + r];
  }
  block_sum_vec<MAXC>(ybar, C, red, bcast);   // also orders the sK/sdK initialisation
  const float invR = 1.0f / (float)R;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) ybar[c] *= invR;

  // pass A: the two means over r that dy needs, and dK
  float m_wu[MAXC], m_g[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) m_wu[c] = m_g[c] = 0.f;
  for (int r0 = 0; r0 < R; r0 += blockDim.x) {
    const int r = r0 + threadIdx.x;
    const bool live = r < R;
    float what[MAXC], gzt[MAXC];
    if (live) {
      float wmax = -INFINITY;
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (c < C) {
          what[c] = ub[(C + c) * R + r];
          wmax = fmaxf(wmax, what[c]);
        }
      const float shift = (wmax == -INFINITY) ? 0.f : wmax;
      float den = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (c < C) {
          what[c] = expf(what[c] - shift);
          den += what[c];
        }
      const float inv_den = 1.0f / den;
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (c < C) {
          what[c] *= inv_den;
          gzt[c] = gb[c * R + r] - gb[(2 * C + c) * R + r];
          m_g[c] += gzt[c];
        }
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (c < C) {
          float uu = 0.f;
#pragma unroll
          for (int cp = 0; cp < MAXC; ++cp)
            if (cp < C) uu = fmaf(sK[c * C + cp], gzt[cp], uu);
          m_wu[c] += what[c] * uu;
        }
    }
    // dK: warp-reduce each of the C*C products into this warp's private slice
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < C) {
        const float ac = live ? what[c] * (ub[c * R + r] - ybar[c]) : 0.f;
#pragma unroll
        for (int cp = 0; cp < MAXC; ++cp) {
          if (cp < C) {
            const float s = warp_sum(live ? ac * gzt[cp] : 0.f);
            if ((threadIdx.x & 31) == 0) my_dK[c * C + cp] += s;
          }
        }
      }
    }
  }
  block_sum_vec<MAXC>(m_wu, C, red, bcast);
  block_sum_vec<MAXC>(m_g, C, red, bcast);
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) {   // fixed warp order: deterministic
    float t = 0.f;
    for (int w = 0; w < kCciThreads / 32; ++w) t += sdK[w * MAXC * MAXC + i];
    partial[b * (int64_t)(C * C) + i] = t;
  }

  // pass B: the gradients themselves
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    float what[MAXC], w[MAXC], gzt[MAXC], t[MAXC], uu[MAXC];
    float wmax = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < C) {
        w[c] = ub[(C + c) * R + r];
        wmax = fmaxf(wmax, w[c]);
      }
    const float shift = (wmax == -INFINITY) ? 0.f : wmax;
    float den = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < C) {
        what[c] = expf(w[c] - shift);
        den += what[c];
      }
    const float inv_den = 1.0f / den;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < C) {
        what[c] *= inv_den;
        gzt[c] = gb[c * R + r] - gb[(2 * C + c) * R + r];
      }
    float tw = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < C) {
        float s = 0.f;
#pragma unroll
        for (int cp = 0; cp < MAXC; ++cp)
          if (cp < C) s = fmaf(sK[c * C + cp], gzt[cp], s);
        uu[c] = s;
        t[c] = (ub[c * R + r] - ybar[c]) * s;
        tw = fmaf(what[c], t[c], tw);
      }
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < C) {
        dub[c * R + r] = what[c] * uu[c] - m_wu[c] * invR + m_g[c] * invR;
        dub[(C + c) * R + r] = what[c] * (t[c] - tw) + expf(w[c]) * gb[(C + c) * R + r];
        dub[(2 * C + c) * R + r] = gb[(2 * C + c) * R + r];
      }
  }
}

// ---- warp-per-encounter variants (C <= 8: the reference's C is 6) ---------------------------
// One warp owns one encounter, lane l owns reference points l, l+32, ...; every cross-thread
// step is a shuffle reduction, so there is no block barrier and no shared-memory traffic on the
// data path.  HBM-bound: ~1.1k warp instructions per encounter against 20.7 KB of traffic.
constexpr int kCciWarps = 4;

// CT > 0: the channel count is a compile-time constant (the reference's 6 vitals): the padded-to-8 loops and
// their `c < C` predicates fold away (36 instead of 64 MACs per 6x6 product, ...).  CT = 0: runtime C <= 8.
template <int CT>
__device__ __forceinline__ void softmax_c8(const float* __restrict__ wrow /*ub + C*R + r*/, int C_rt, int R,
                                           float (&w)[8], float (&what)[8]) {
  const int C = CT ? CT : C_rt;
  float wmax = -INFINITY;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (c < C) {
      w[c] = wrow[c * R];
      wmax = fmaxf(wmax, w[c]);
    }
  const float shift = (wmax == -INFINITY) ? 0.f : wmax;   // torch.logsumexp guards an all -inf row
  float den = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (c < C) {
      what[c] = expf(w[c] - shift);
      den += what[c];
    }
  const float inv_den = 1.0f / den;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (c < C) what[c] *= inv_den;
}

// Stages one encounter's (3C x R) tile(s) into this warp's slice of shared memory with 1-D TMA
// bulk copies on a per-warp mbarrier: the loads of all warps of all resident CTAs are in flight
// at once (>= 100 KB per SM), which is what an HBM-bound kernel with this little work needs.
__device__ __forceinline__ const float* warp_tile_load(unsigned char* smem, int warp, int lane, int ntiles,
                                                       const float* g0, const float* g1, uint32_t bytes,
                                                       const float* g2 = nullptr) {
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem) + warp;
  float* tile = reinterpret_cast<float*>(smem + 64 + (size_t)warp * ntiles * bytes);
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_proxy_async();
  }
  __syncwarp();
  if (lane == 0) {
    mbar_expect_tx(bar, bytes * ntiles);
    bulk_g2s(tile, g0, bytes, bar);
    if (ntiles > 1) bulk_g2s(reinterpret_cast<unsigned char*>(tile) + bytes, g1, bytes, bar);
    if (ntiles > 2) bulk_g2s(reinterpret_cast<unsigned char*>(tile) + 2 * bytes, g2, bytes, bar);
  }
  mbar_wait(bar, 0);
  return tile;
}

template <bool TILE, int CT>
__global__ void __launch_bounds__(kCciWarps * 32)
cci_fwd_warp_kernel(const float* __restrict__ u, const float* __restrict__ kernel, float* __restrict__ out,
                    int64_t B, int C_rt, int R) {
  const int C = CT ? CT : C_rt;
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ float sK[64];
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) sK[i] = __ldg(kernel + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * kCciWarps + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* ub = u + b * (int64_t)(3 * C) * R;
  if (TILE) ub = warp_tile_load(dyn_smem, threadIdx.x >> 5, lane, 1, ub, nullptr, (uint32_t)(3 * C * R) * 4u);
  float* ob = out + b * (int64_t)(3 * C) * R;
  float ybar[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) ybar[c] = 0.f;
  for (int r = lane; r < R; r += 32) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < C) ybar[c] += ub[c * R + r];
  }
  const float invR = 1.0f / (float)R;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (c < C) ybar[c] = warp_sum(ybar[c]) * invR;
  for (int r = lane; r < R; r += 32) {
    float w[8], a[8];
    softmax_c8<CT>(ub + C * R + r, C, R, w, a);
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < C) {
        a[c] *= ub[c * R + r] - ybar[c];
        ob[(C + c) * R + r] = expf(w[c]);                          // intensity, :104
      }
#pragma unroll
    for (int cp = 0; cp < 8; ++cp)
      if (cp < C) {
        float z = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < C) z = fmaf(a[c], sK[c * C + cp], z);
        z += ybar[cp];
        ob[cp * R + r] = z;
        ob[(2 * C + cp) * R + r] = ub[(2 * C + cp) * R + r] - z;   // transient minus smooth, :122-123
      }
  }
}

// FUSE: the SCI backward folded in (dic_cci_sci_bwd).  The gradient with respect to the SCI output is consumed where it
// is produced - d alpha_c = - sum_r [ gy U1 + gw U0 + gy' U1' ] against the moment rows `stats` the forward pass saved
// (interp_sci.cu) - instead of making a round trip through HBM as grad_u (13.8 KB per encounter written, then read).
template <bool TILE, int CT, bool FUSE>
__global__ void __launch_bounds__(kCciWarps * 32)
cci_bwd_warp_kernel(const float* __restrict__ u, const float* __restrict__ kernel,
                    const float* __restrict__ grad_out, float* __restrict__ grad_u,
                    float* __restrict__ partial /*(B, C*C)*/, const float* __restrict__ stats,
                    float* __restrict__ sci_partial /*(B, C)*/, int64_t B, int C_rt, int R) {
  const int C = CT ? CT : C_rt;
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ float sK[64];
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) sK[i] = __ldg(kernel + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * kCciWarps + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* ub = u + b * (int64_t)(3 * C) * R;
  const float* gb = grad_out + b * (int64_t)(3 * C) * R;
  const float* sb = FUSE ? stats + b * (int64_t)(3 * C) * R : nullptr;
  if (TILE) {     // u | grad_out (| stats): all tiles of this warp in flight on one mbarrier
    ub = warp_tile_load(dyn_smem, threadIdx.x >> 5, lane, FUSE ? 3 : 2, ub, gb, (uint32_t)(3 * C * R) * 4u, sb);
    gb = ub + 3 * C * R;
    if (FUSE) sb = gb + 3 * C * R;
  }
  float* dub = FUSE ? nullptr : grad_u + b * (int64_t)(3 * C) * R;
  const float invR = 1.0f / (float)R;

  float ybar[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) ybar[c] = 0.f;
  for (int r = lane; r < R; r += 32) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < C) ybar[c] += ub[c * R + r];
  }
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (c < C) ybar[c] = warp_sum(ybar[c]) * invR;

  // pass A: dK products and the two means over r that dy needs
  float acc0[32], acc1[32], means[16];        // dK rows 0-3 | rows 4-7 | [mean(what uu) | mean(gzt)]
#pragma unroll
  for (int i = 0; i < 32; ++i) acc0[i] = acc1[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) means[i] = 0.f;
  for (int r = lane; r < R; r += 32) {
    float w[8], what[8], gzt[8];
    softmax_c8<CT>(ub + C * R + r, C, R, w, what);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      gzt[c] = 0.f;
      if (c < C) {
        gzt[c] = gb[c * R + r] - gb[(2 * C + c) * R + r];
        means[8 + c] += gzt[c];
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < C) {
        float uu = 0.f;
#pragma unroll
        for (int cp = 0; cp < 8; ++cp)
          if (cp < C) uu = fmaf(sK[c * C + cp], gzt[cp], uu);
        means[c] = fmaf(what[c], uu, means[c]);
        const float ac = what[c] * (ub[c * R + r] - ybar[c]);
#pragma unroll
        for (int cp = 0; cp < 8; ++cp) {
          if (c < 4) acc0[c * 8 + cp] = fmaf(ac, gzt[cp], acc0[c * 8 + cp]);
          else acc1[(c - 4) * 8 + cp] = fmaf(ac, gzt[cp], acc1[(c - 4) * 8 + cp]);
        }
      }
  }
  {
    // lane l ends with dK[(l >> 3) (+4), l & 7]: one coalesced store per half
    const float v0 = warp_reduce_multi<32>(acc0, lane);
    const float v1 = warp_reduce_multi<32>(acc1, lane);
    const int c0 = lane >> 3, cp = lane & 7;
    float* pb = partial + b * (int64_t)(C * C);
    if (c0 < C && cp < C) pb[c0 * C + cp] = v0;
    if (c0 + 4 < C && cp < C) pb[(c0 + 4) * C + cp] = v1;
  }
  float m_wu[8], m_g[8];
  {
    const float v = warp_reduce_multi<16>(means, lane) * invR;    // lane l holds value l >> 1
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      m_wu[c] = __shfl_sync(0xffffffffu, v, c << 1);
      m_g[c] = __shfl_sync(0xffffffffu, v, (8 + c) << 1);
    }
  }

  // pass B: the gradients (rows are L1/L2 hot)
  float dal[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) dal[c] = 0.f;
  for (int r = lane; r < R; r += 32) {
    float w[8], what[8], gzt[8], t[8], uu[8];
    softmax_c8<CT>(ub + C * R + r, C, R, w, what);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      gzt[c] = 0.f;
      if (c < C) gzt[c] = gb[c * R + r] - gb[(2 * C + c) * R + r];
    }
    float tw = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < C) {
        float s = 0.f;
#pragma unroll
        for (int cp = 0; cp < 8; ++cp)
          if (cp < C) s = fmaf(sK[c * C + cp], gzt[cp], s);
        uu[c] = s;
        t[c] = (ub[c * R + r] - ybar[c]) * s;
        tw = fmaf(what[c], t[c], tw);
      }
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < C) {
        const float gy = what[c] * uu[c] - m_wu[c] + m_g[c];
        const float gw = what[c] * (t[c] - tw) + expf(w[c]) * gb[(C + c) * R + r];
        const float gyp = gb[(2 * C + c) * R + r];
        if (FUSE) {
          const float u0 = sb[(C + c) * R + r];
          if (u0 >= 0.f)        // U0 < 0 marks an all-masked vital: no contribution (its upstream gradients may be NaN)
            dal[c] = fmaf(gy, sb[c * R + r], fmaf(gw, u0, fmaf(gyp, sb[(2 * C + c) * R + r], dal[c])));
        } else {
          dub[c * R + r] = gy;
          dub[(C + c) * R + r] = gw;
          dub[(2 * C + c) * R + r] = gyp;
        }
      }
  }
  if (FUSE) {
    const float v = warp_reduce_multi<8>(dal, lane);          // lane l holds vital l >> 2
    if ((lane & 3) == 0 && (lane >> 2) < C) sci_partial[b * C + (lane >> 2)] = -v;
  }
}

__global__ void sigmoid_vec_kernel(const float* __restrict__ kernel, float* __restrict__ out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = sigmoid_ref(kernel[c]);
}

int check(const void* u, const void* kernel, int64_t B, int C, int R) {
  DIC_REQUIRE((u || B == 0) && kernel, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE(B >= 0 && C > 0 && R > 0, DIC_ERR_INVALID_ARGUMENT, "bad sizes B=%lld C=%d R=%d",
              (long long)B, C, R);
  DIC_REQUIRE(C <= 16, DIC_ERR_UNSUPPORTED, "CrossChannelInterp supports d_dim <= 16 (got %d)", C);
  DIC_REQUIRE(B <= 2147483647LL, DIC_ERR_UNSUPPORTED, "B=%lld exceeds the grid limit", (long long)B);
  return DIC_OK;
}

}  // namespace
}  // namespace dic

using namespace dic;

extern "C" int dic_cci_fwd(const float* u, const float* kernel, float* out, int64_t B, int C, int R,
                           dic_stream_t stream) {
  int rc = check(u, kernel, B, C, R);
  if (rc) return rc;
  DIC_REQUIRE(out || B == 0, DIC_ERR_INVALID_ARGUMENT, "null output pointer");
  if (B == 0) return DIC_OK;
  cudaStream_t st = as_stream(stream);
  if (C <= 8) {
    const unsigned grid = (unsigned)((B + kCciWarps - 1) / kCciWarps);
    const size_t tile = (size_t)3 * C * R * 4;
    const size_t smem = 64 + kCciWarps * tile;
    if (tile % 16 == 0 && aligned16(u) && smem <= (size_t)kMaxSmemBytes) {
      auto kf = C == 6 ? cci_fwd_warp_kernel<true, 6> : cci_fwd_warp_kernel<true, 0>;
      if (smem > 48 * 1024)
        DIC_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kf<<<grid, kCciWarps * 32, smem, st>>>(u, kernel, out, B, C, R);
    } else
      cci_fwd_warp_kernel<false, 0><<<grid, kCciWarps * 32, 0, st>>>(u, kernel, out, B, C, R);
  } else cci_fwd_kernel<16><<<(unsigned)B, kCciThreads, 0, st>>>(u, kernel, out, C, R);
  DIC_LAUNCH_CHECK("cci_fwd_kernel");
  return DIC_OK;
}

extern "C" size_t dic_cci_bwd_workspace_bytes(int64_t B, int C) {
  if (B < 0 || C <= 0) return 0;
  size_t part = ((size_t)B * C * C * sizeof(float) + 255) / 256 * 256;
  return part + (size_t)kColsumBlocks * C * C * sizeof(double) + 256;
}

extern "C" int dic_cci_bwd(const float* u, const float* kernel, const float* grad_out, float* grad_u,
                           float* d_kernel, void* workspace, int64_t B, int C, int R,
                           dic_stream_t stream) {
  int rc = check(u, kernel, B, C, R);
  if (rc) return rc;
  DIC_REQUIRE(((grad_out && grad_u && workspace) || B == 0) && d_kernel, DIC_ERR_INVALID_ARGUMENT,
              "null pointer argument");
  cudaStream_t st = as_stream(stream);
  if (B == 0) {
    DIC_CUDA(cudaMemsetAsync(d_kernel, 0, sizeof(float) * C * C, st));
    return DIC_OK;
  }
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  float* partial = reinterpret_cast<float*>(ws);
  double* red = reinterpret_cast<double*>(ws + ((size_t)B * C * C * sizeof(float) + 255) / 256 * 256);
  if (C <= 8) {
    const unsigned grid = (unsigned)((B + kCciWarps - 1) / kCciWarps);
    const size_t tile = (size_t)3 * C * R * 4, smem = 64 + kCciWarps * 2 * tile;
    if (tile % 16 == 0 && aligned16(u) && aligned16(grad_out) && smem <= (size_t)kMaxSmemBytes) {
      auto kf = C == 6 ? cci_bwd_warp_kernel<true, 6, false> : cci_bwd_warp_kernel<true, 0, false>;
      if (smem > 48 * 1024)
        DIC_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kf<<<grid, kCciWarps * 32, smem, st>>>(u, kernel, grad_out, grad_u, partial, nullptr, nullptr, B, C, R);
    } else {
      cci_bwd_warp_kernel<false, 0, false><<<grid, kCciWarps * 32, 0, st>>>(u, kernel, grad_out, grad_u, partial, nullptr,
                                                                          nullptr, B, C, R);
    }
  } else
    cci_bwd_kernel<16><<<(unsigned)B, kCciThreads, 0, st>>>(u, kernel, grad_out, grad_u, partial, C, R);
  DIC_LAUNCH_CHECK("cci_bwd_kernel");
  return colsum_f32_launch(partial, nullptr, d_kernel, nullptr, red, B, C * C, st);
}

extern "C" size_t dic_cci_sci_bwd_workspace_bytes(int64_t B, int C) {
  if (B < 0 || C <= 0) return 0;
  const size_t p_cci = ((size_t)B * C * C * sizeof(float) + 255) / 256 * 256;
  const size_t p_sci = ((size_t)B * C * sizeof(float) + 255) / 256 * 256;
  return p_cci + p_sci + (size_t)kColsumBlocks * C * C * sizeof(double) + (size_t)C * sizeof(float) + 512;
}

extern "C" int dic_cci_sci_bwd(const float* u, const float* cci_kernel, const float* sci_kernel, const float* stats,
                               const float* grad_out, float* d_cci_kernel, float* d_sci_kernel, void* workspace,
                               int64_t B, int C, int R, dic_stream_t stream) {
  int rc = check(u, cci_kernel, B, C, R);
  if (rc) return rc;
  DIC_REQUIRE(C <= 8, DIC_ERR_UNSUPPORTED, "the fused CCI + SCI backward covers d_dim <= 8 (got %d): call dic_cci_bwd and "
              "dic_sci_bwd", C);
  DIC_REQUIRE(((grad_out && stats && workspace) || B == 0) && d_cci_kernel && d_sci_kernel && sci_kernel,
              DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  cudaStream_t st = as_stream(stream);
  if (B == 0) {
    DIC_CUDA(cudaMemsetAsync(d_cci_kernel, 0, sizeof(float) * C * C, st));
    DIC_CUDA(cudaMemsetAsync(d_sci_kernel, 0, sizeof(float) * C, st));
    return DIC_OK;
  }
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  const size_t p_cci = ((size_t)B * C * C * sizeof(float) + 255) / 256 * 256;
  const size_t p_sci = ((size_t)B * C * sizeof(float) + 255) / 256 * 256;
  float* partial = reinterpret_cast<float*>(ws);
  float* sci_partial = reinterpret_cast<float*>(ws + p_cci);
  double* red = reinterpret_cast<double*>(ws + p_cci + p_sci);
  float* sig = reinterpret_cast<float*>(ws + p_cci + p_sci + (size_t)kColsumBlocks * C * C * sizeof(double));
  const unsigned grid = (unsigned)((B + kCciWarps - 1) / kCciWarps);
  const size_t tile = (size_t)3 * C * R * 4, smem = 64 + kCciWarps * 3 * tile;
  if (tile % 16 == 0 && aligned16(u) && aligned16(grad_out) && aligned16(stats) && smem <= (size_t)kMaxSmemBytes) {
    auto kf = C == 6 ? cci_bwd_warp_kernel<true, 6, true> : cci_bwd_warp_kernel<true, 0, true>;
    if (smem > 48 * 1024)
      DIC_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kf<<<grid, kCciWarps * 32, smem, st>>>(u, cci_kernel, grad_out, nullptr, partial, stats, sci_partial, B, C, R);
  } else {
    cci_bwd_warp_kernel<false, 0, true><<<grid, kCciWarps * 32, 0, st>>>(u, cci_kernel, grad_out, nullptr, partial, stats,
                                                                       sci_partial, B, C, R);
  }
  DIC_LAUNCH_CHECK("cci_bwd_warp_kernel<fused>");
  rc = colsum_f32_launch(partial, nullptr, d_cci_kernel, nullptr, red, B, C * C, st);
  if (rc) return rc;
  sigmoid_vec_kernel<<<(C + 127) / 128, 128, 0, st>>>(sci_kernel, sig, C);
  DIC_LAUNCH_CHECK("sigmoid_vec_kernel");
  return colsum_f32_launch(sci_partial, nullptr, d_sci_kernel, sig, red, B, C, st);
}
