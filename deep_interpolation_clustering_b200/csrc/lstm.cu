// Bidirectional LSTM (hidden 128) for the encoder / decoder either side of the interpolation network
// (pretrain_interp.py:14-41,108-112,138-142: nn.LSTM(18 -> 128) and nn.LSTM(256 -> 128), one layer, bidirectional).
// SURVEY 8(f) rank 2.  float32 semantics of torch.nn.LSTM: gates i, f, g, o (that row order in weight_ih / weight_hh),
//     a_t = W_ih x_t + b_ih + b_hh + W_hh h_(t-1),  i,f,o = sigmoid, g = tanh,  c_t = f c_(t-1) + i g,  h_t = o tanh(c_t).
//
// Forward = ONE persistent kernel over all time steps (dic_lstm_fwd).  The input projection W_ih x_t + b for all steps is
// a plain library GEMM done by the caller ("pre"); what cannot be a library call is the recurrence: R dependent steps of
// a (B x 128) x (128 x 512) product with a non-linear epilogue in between.
//   * a CLUSTER of 4 CTAs owns 128 encounters of one direction for all R steps; CTA q owns hidden units [32q, 32q + 32)
//     = 128 gate columns, and keeps its slice of W_hh resident in shared memory for the whole sequence as the B operand
//     of tcgen05.mma (fp16 hi + lo halves of the power-of-two-scaled weights: 2 x 32 KB);
//   * h_(t-1) (128 x 128) is the A operand, also as fp16 hi + lo (|h| < 1: the two halves carry 22 bits); three
//     MMAs per K step (hi.hi + hi.lo + lo.hi, kind::f16, M = N = 128, K = 16) accumulate float32-grade products in TMEM;
//   * epilogue: 8 warps, thread = (encounter, 16 units): tcgen05.ld of its 64 gate pre-activations, + pre (prefetched from
//     global before the MMA wait), MUFU.EX2 / MUFU.RCP activations, the cell state lives in registers across all steps;
//     h_t goes out as float32 (the layer output), as saved state for the backward pass, and as fp16 hi / lo halves
//     straight into the A buffers of ALL FOUR CTAs of the cluster through distributed shared memory
//     (st.shared::cluster) - the exchange of the recurrence never touches global memory;
//   * one barrier.cluster per step (release / acquire) orders the DSMEM writes before the next step's MMAs; the A
//     operand is double buffered, so a fast CTA never overwrites what a slow one still multiplies.
// Backward (dic_lstm_bwd_step): the gate-gradient algebra of one step fused in one kernel (reads the saved gates and
// cell states, the upstream and the recurrent gradient; writes d a_t (B, 512) and the running d c); the recurrent
// product d h_(t-1) = d a_t W_hh and the weight gradients are library GEMMs on the saved tensors (host side).
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

namespace dic {
namespace {

using namespace tc;

constexpr int kH = 128;                 // hidden size
constexpr int kRows = 128;              // encounters per cluster tile
constexpr int kCl = 4;                  // CTAs per cluster = slices of the gate dimension
constexpr int kThreads = 256;
constexpr int kLbo = kRows * 16;        // bytes between consecutive 16-byte K chunks (128 rows x 16 B)
constexpr int kSbo = 128;               // bytes between 8-row groups
constexpr int kTileB = 16 * kLbo;       // one fp16 operand tile, 128 x 128: 32 KB
constexpr uint32_t kIdescF16 = make_idesc(0u, 128u, 128u);

struct LstmSmem {
  static constexpr size_t w = 0;                          // [hi | lo] B operand: W_hh slice, 64 KB
  static constexpr size_t a = w + 2 * (size_t)kTileB;     // [buf 0: hi | lo][buf 1: hi | lo] A operand: h, 128 KB
  static constexpr size_t bars = a + 4 * (size_t)kTileB;
  static constexpr size_t total = bars + 64;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
// generic-proxy writes to (distributed) shared memory -> visible to the tensor core's async-proxy reads.  The
// unqualified fence.proxy.async also orders GLOBAL memory and costs microseconds per step here.
__device__ __forceinline__ void fence_proxy_async_cluster() {
  asm volatile("fence.proxy.async.shared::cluster;" ::: "memory");
}

// tcgen05.ld 32 lanes x 32 bit x 8 columns, issue only (tcgen05.wait::ld by the caller once all pieces are in flight)
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_wait4x8(uint32_t (&a)[8], uint32_t (&b)[8], uint32_t (&c)[8], uint32_t (&d)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                 "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]),
                 "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7]),
                 "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]), "+r"(d[4]), "+r"(d[5]), "+r"(d[6]), "+r"(d[7])
               :
               : "memory");
}

// sigmoid / tanh with float32-grade accuracy from MUFU.EX2 + MUFU.RCP (each ~1 ulp; tanh.approx has only 2^-11),
// branch free: the epilogue is MUFU bound (10 per hidden unit and step), so everything else has to stay off the
// critical path - an if / else per activation cost 25 % of the kernel's instructions as BSSY / BRA / BSYNC.
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_acc(float x) { return rcp_approx(1.0f + ex2_approx(-kLog2e * x)); }
__device__ __forceinline__ float tanh_acc(float x) {
  const float ax = fabsf(x);
  const float x2 = x * x;
  const float small = x * fmaf(x2, fmaf(x2, 0.13333334f, -0.33333334f), 1.0f);   // |x| < 0.04: error < 2e-9 (no cancellation)
  const float t = ex2_approx(-2.0f * kLog2e * ax);
  const float big = copysignf((1.0f - t) * rcp_approx(1.0f + t), x);
  return ax < 0.04f ? small : big;
}

// 8 floats -> 8 fp16 "hi" and 8 fp16 "lo" (residual), one 16-byte K chunk each
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  __half2 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
    const float2 back = __half22float2(h[i]);
    l[i] = __floats2half2_rn(x[2 * i] - back.x, x[2 * i + 1] - back.y);
  }
  hi = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]),
                  *reinterpret_cast<uint32_t*>(&h[2]), *reinterpret_cast<uint32_t*>(&h[3]));
  lo = make_uint4(*reinterpret_cast<uint32_t*>(&l[0]), *reinterpret_cast<uint32_t*>(&l[1]),
                  *reinterpret_cast<uint32_t*>(&l[2]), *reinterpret_cast<uint32_t*>(&l[3]));
}

// Column n of CTA q's gate slice <-> (gate g, hidden unit j): n = 64 half + 16 g + u, j = 32 q + 16 half + u.
// W_hh (4H x H, rows i|f|g|o) -> per (direction, q): the UMMA B operand (N = 128 gate columns x K = 128 hidden), canonical
// no-swizzle K-major fp16 tiles [hi | lo] of s W, s = the power of two that puts max |W| into [2^12, 2^13).
__global__ void lstm_pack_whh_kernel(const float* __restrict__ w_hh, const float* __restrict__ w_hh_rev,
                                     unsigned char* __restrict__ packed, float* __restrict__ inv_scale) {
  const int dir = blockIdx.y, q = blockIdx.x;
  const float* W = dir ? w_hh_rev : w_hh;
  __shared__ float red[32];
  float mx = 0.f;
  for (int i = threadIdx.x; i < 4 * kH * kH; i += blockDim.x) mx = fmaxf(mx, fabsf(W[i]));
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  int e = 0;
  if (mx > 0.f && isfinite(mx)) frexpf(mx, &e);             // mx = m 2^e, m in [0.5, 1)
  const float s = ldexpf(1.0f, 13 - e);                    // s mx in [2^12, 2^13)
  if (q == 0 && threadIdx.x == 0) inv_scale[dir] = 1.0f / s;
  unsigned char* dst = packed + (size_t)(dir * kCl + q) * 2 * kTileB;
  for (int idx = threadIdx.x; idx < 128 * 16; idx += blockDim.x) {      // (n, K chunk c)
    const int n = idx >> 4, c = idx & 15;
    const int half = n >> 6, g = (n >> 4) & 3, u = n & 15;
    const int row = g * kH + 32 * q + 16 * half + u;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = s * W[(size_t)row * kH + c * 8 + i];
    uint4 hi, lo;
    split8(x, hi, lo);
    const size_t off = (size_t)c * kLbo + (size_t)n * 16;
    *reinterpret_cast<uint4*>(dst + off) = hi;
    *reinterpret_cast<uint4*>(dst + kTileB + off) = lo;
  }
}

// pre   (R, B, 2 * 512) float32: W_ih x_t + b_ih + b_hh, columns ordered [dir][q][n] (n as above)
// h0/c0 (2, B, 128) float32 or NULL (zeros)
// out   (R, B, 256): [forward h_t | reverse h_t];  hn / cn (2, B, 128)
// save  (2, R, B, 5, 128) float32 or NULL: i, f, g, o, c_t per step (what the backward pass reads)
__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kThreads, 1)
lstm_fwd_kernel(const float* __restrict__ pre, const unsigned char* __restrict__ packed, const float* __restrict__ inv_scale,
                const float* __restrict__ h0, const float* __restrict__ c0, float* __restrict__ out,
                float* __restrict__ hn, float* __restrict__ cn, float* __restrict__ save, int R, int64_t B) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sw = smem + LstmSmem::w;
  unsigned char* sa = smem + LstmSmem::a;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + LstmSmem::bars);
  uint64_t* bar_acc = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t q = cluster_ctarank();
  const int64_t cluster_id = blockIdx.x / kCl;
  const int dir = (int)(cluster_id & 1);
  const int64_t b0 = (cluster_id >> 1) * kRows;
  const int row = 32 * (warp & 3) + lane;          // TMEM lane quarter of a warp = warp % 4
  const int half = warp >> 2;                      // which 64 of the 128 gate columns (16 units x 4 gates)
  const int64_t b = b0 + row;
  const bool live = b < B;
  const int j0 = 32 * (int)q + 16 * half;          // first hidden unit of this thread

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_acc, 1);
    fence_proxy_async();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  __syncthreads();
  if (tid == 0) {                                   // the weight slice: one 64 KB bulk copy, resident for all R steps
    mbar_expect_tx(bar_w, 2u * kTileB);
    bulk_g2s(sw, packed + (size_t)(dir * kCl + q) * 2 * kTileB, 2u * kTileB, bar_w);
  }
  // h_(-1): every CTA fills its own A buffer 0 completely (thread = row x 64 hidden units); cell state -> registers
  {
    const int k0 = 64 * half;
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = (h0 && live) ? __ldg(h0 + ((int64_t)dir * B + b) * kH + k0 + cc * 8 + i) : 0.f;
      uint4 hi, lo;
      split8(x, hi, lo);
      const size_t off = (size_t)(k0 / 8 + cc) * kLbo + (size_t)row * 16;
      *reinterpret_cast<uint4*>(sa + off) = hi;
      *reinterpret_cast<uint4*>(sa + kTileB + off) = lo;
    }
  }
  float c[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) c[u] = (c0 && live) ? __ldg(c0 + ((int64_t)dir * B + b) * kH + j0 + u) : 0.f;
  const float inv_s = __ldg(inv_scale + dir);
  fence_proxy_async_cluster();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  mbar_wait(bar_w, 0);
  cluster_arrive();                                  // every CTA of the cluster is running and initialised
  cluster_wait();

  const uint32_t sa_u = smem_u32(sa), sw_u = smem_u32(sw);
  uint32_t remote_a[kCl];
#pragma unroll
  for (int p = 0; p < kCl; ++p) remote_a[p] = map_to_cta(sa_u, (uint32_t)p);
  const uint32_t my_chunk_off = (uint32_t)((4 * q + 2 * half) * kLbo + row * 16);     // first of my two K chunks of h

  float hh[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) hh[u] = 0.f;

  for (int step = 0; step < R; ++step) {
    const int t = dir ? R - 1 - step : step;
    const int buf = step & 1;
    // ---- prefetch this step's input projection (64 floats of my row) while the tensor core works ----
    float4 pv[16];
    {
      const float4* src = reinterpret_cast<const float4*>(pre + ((int64_t)t * B + (live ? b : 0)) * (2 * 4 * kH) +
                                                          dir * 4 * kH + (int)q * 128 + 64 * half);
#pragma unroll
      for (int i = 0; i < 16; ++i) pv[i] = live ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // ---- a_t (recurrent part) = h_(t-1) W_hh^T on the tensor core ----
    if (warp == 0) {
      fence_proxy_async();                          // the peers' h slices (acquired by the cluster barrier) -> async proxy
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_hi = sa_u + (uint32_t)buf * 2u * kTileB, a_lo = a_hi + kTileB;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {            // one MMA consumes K = 16 halves = two 16-byte chunks
          const uint64_t dah = make_desc_kmajor(a_hi + ks * 2 * kLbo, kLbo, kSbo);
          const uint64_t dal = make_desc_kmajor(a_lo + ks * 2 * kLbo, kLbo, kSbo);
          const uint64_t dbh = make_desc_kmajor(sw_u + ks * 2 * kLbo, kLbo, kSbo);
          const uint64_t dbl = make_desc_kmajor(sw_u + kTileB + ks * 2 * kLbo, kLbo, kSbo);
          umma_f16(tmem, dah, dbh, kIdescF16, ks > 0 ? 1u : 0u);
          umma_f16(tmem, dah, dbl, kIdescF16, 1u);
          umma_f16(tmem, dal, dbh, kIdescF16, 1u);
        }
        umma_commit(bar_acc);
      }
      __syncwarp();
    }
    mbar_wait(bar_acc, (uint32_t)step & 1u);
    tc_fence_after();
    // ---- gates, cell, hidden: two passes of 8 units (4 x 8 accumulator columns each) keep the register count down ----
    const float* pf = reinterpret_cast<const float*>(pv);
    float* sp = (save && live) ? save + (((int64_t)dir * R + t) * B + b) * (5 * kH) + j0 : nullptr;
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {
      uint32_t ai[8], af[8], ag[8], ao[8];
      const uint32_t tcol = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(64 * half + 8 * ps);
      tmem_ld8_issue(tcol, ai);
      tmem_ld8_issue(tcol + 16u, af);
      tmem_ld8_issue(tcol + 32u, ag);
      tmem_ld8_issue(tcol + 48u, ao);
      tmem_wait4x8(ai, af, ag, ao);
      float gi[8], gf[8], gg[8], go[8];
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const int u = 8 * ps + v;
        gi[v] = sigmoid_acc(fmaf(__uint_as_float(ai[v]), inv_s, pf[u]));
        gf[v] = sigmoid_acc(fmaf(__uint_as_float(af[v]), inv_s, pf[16 + u]));
        gg[v] = tanh_acc(fmaf(__uint_as_float(ag[v]), inv_s, pf[32 + u]));
        go[v] = sigmoid_acc(fmaf(__uint_as_float(ao[v]), inv_s, pf[48 + u]));
        c[u] = fmaf(gf[v], c[u], gi[v] * gg[v]);
        hh[u] = go[v] * tanh_acc(c[u]);
      }
      if (sp) {
        float4* s4 = reinterpret_cast<float4*>(sp + 8 * ps);
        s4[0] = make_float4(gi[0], gi[1], gi[2], gi[3]);
        s4[1] = make_float4(gi[4], gi[5], gi[6], gi[7]);
        s4 = reinterpret_cast<float4*>(sp + kH + 8 * ps);
        s4[0] = make_float4(gf[0], gf[1], gf[2], gf[3]);
        s4[1] = make_float4(gf[4], gf[5], gf[6], gf[7]);
        s4 = reinterpret_cast<float4*>(sp + 2 * kH + 8 * ps);
        s4[0] = make_float4(gg[0], gg[1], gg[2], gg[3]);
        s4[1] = make_float4(gg[4], gg[5], gg[6], gg[7]);
        s4 = reinterpret_cast<float4*>(sp + 3 * kH + 8 * ps);
        s4[0] = make_float4(go[0], go[1], go[2], go[3]);
        s4[1] = make_float4(go[4], go[5], go[6], go[7]);
        s4 = reinterpret_cast<float4*>(sp + 4 * kH + 8 * ps);
        s4[0] = make_float4(c[8 * ps], c[8 * ps + 1], c[8 * ps + 2], c[8 * ps + 3]);
        s4[1] = make_float4(c[8 * ps + 4], c[8 * ps + 5], c[8 * ps + 6], c[8 * ps + 7]);
      }
    }
    tc_fence_before();
    // ---- h_t -> the A operand (next buffer) of all four CTAs, through distributed shared memory ----
    {
      uint4 hi0, lo0, hi1, lo1;
      const float(&x0)[8] = *reinterpret_cast<const float(*)[8]>(&hh[0]);
      const float(&x1)[8] = *reinterpret_cast<const float(*)[8]>(&hh[8]);
      split8(x0, hi0, lo0);
      split8(x1, hi1, lo1);
      const uint32_t nb = (uint32_t)(buf ^ 1) * 2u * kTileB + my_chunk_off;
#pragma unroll
      for (int p = 0; p < kCl; ++p) {
        const uint32_t base = remote_a[p] + nb;
        st_cluster_v4(base, hi0);
        st_cluster_v4(base + kLbo, hi1);
        st_cluster_v4(base + kTileB, lo0);
        st_cluster_v4(base + kTileB + kLbo, lo1);
      }
    }
    fence_proxy_async_cluster();
    cluster_arrive();
    // ---- the layer output (overlaps the barrier latency) ----
    if (live) {
      float4* o4 = reinterpret_cast<float4*>(out + ((int64_t)t * B + b) * (2 * kH) + dir * kH + j0);
#pragma unroll
      for (int i = 0; i < 4; ++i) o4[i] = make_float4(hh[4 * i], hh[4 * i + 1], hh[4 * i + 2], hh[4 * i + 3]);
    }
    __syncwarp();
    cluster_wait();                                  // all four slices of h_t have landed in every A buffer
  }
  if (live) {
    float* hp = hn + ((int64_t)dir * B + b) * kH + j0;
    float* cp = cn + ((int64_t)dir * B + b) * kH + j0;
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      hp[u] = hh[u];
      cp[u] = c[u];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

// The same recurrence with TWO 128-encounter tiles per cluster, interleaved (dic_lstm_fwd takes this kernel when
// B > 128).  The single-tile kernel above is a chain barrier -> MMA -> TMEM load -> MUFU-bound epilogue -> DSMEM stores ->
// barrier with nothing to overlap (ncu: tensor pipe 5 % busy, issue slots 14 %, a cluster step 13.7 k clk against ~4 k
// of epilogue work).  Here every half-step issues the MMAs of one tile and then runs the epilogue of the OTHER tile
// underneath them:
//     half-step hs (tile = hs & 1):  issue MMA(other, next step)  |  epilogue(tile): gates, h -> A[tile] of all 4 CTAs
//                                    wait until MMA(other) has completed  ->  barrier.cluster
//   * A[tile] is single buffered (2 x 64 KB: the shared-memory budget of the double-buffered single tile): its next
//     writers are the epilogues of half-step hs, and every CTA has waited for ITS MMA on that tile before the barrier
//     that ended half-step hs - 1, so no CTA's tensor core still reads what a fast peer overwrites;
//   * the h slices written in half-step hs are consumed by the MMA issued in half-step hs + 1, after the barrier;
//   * two 128-column TMEM accumulators, one mbarrier each;
//   * pre of the next half-step is pulled into L2 with prefetch.global.L2 one half-step ahead (no registers), the
//     register loads at the start of an epilogue then cost an L2 hit.
struct Lstm2Smem {
  static constexpr size_t w = 0;                          // [hi | lo] W_hh slice, 64 KB
  static constexpr size_t a = w + 2 * (size_t)kTileB;     // [tile 0: hi | lo][tile 1: hi | lo], 128 KB
  static constexpr size_t bars = a + 4 * (size_t)kTileB;
  static constexpr size_t total = bars + 64;
};

__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kThreads, 1)
lstm_fwd2_kernel(const float* __restrict__ pre, const unsigned char* __restrict__ packed, const float* __restrict__ inv_scale,
                 const float* __restrict__ h0, const float* __restrict__ c0, float* __restrict__ out,
                 float* __restrict__ hn, float* __restrict__ cn, float* __restrict__ save, int R, int64_t B) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sw = smem + Lstm2Smem::w;
  unsigned char* sa = smem + Lstm2Smem::a;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + Lstm2Smem::bars);
  uint64_t* bar_acc = bar_w + 1;                   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 3);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t q = cluster_ctarank();
  const int64_t cluster_id = blockIdx.x / kCl;
  const int dir = (int)(cluster_id & 1);
  const int64_t b0 = (cluster_id >> 1) * (2 * kRows);
  const int row = 32 * (warp & 3) + lane;
  const int half = warp >> 2;
  const int j0 = 32 * (int)q + 16 * half;
  const int64_t bt[2] = {b0 + row, b0 + kRows + row};
  const bool live[2] = {bt[0] < B, bt[1] < B};

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_acc, 1);
    mbar_init(bar_acc + 1, 1);
    fence_proxy_async();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar_w, 2u * kTileB);
    bulk_g2s(sw, packed + (size_t)(dir * kCl + q) * 2 * kTileB, 2u * kTileB, bar_w);
  }
  float c[2][16];
#pragma unroll
  for (int tl = 0; tl < 2; ++tl) {
    const int k0 = 64 * half;
    unsigned char* at = sa + (size_t)tl * 2 * kTileB;
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        x[i] = (h0 && live[tl]) ? __ldg(h0 + ((int64_t)dir * B + bt[tl]) * kH + k0 + cc * 8 + i) : 0.f;
      uint4 hi, lo;
      split8(x, hi, lo);
      const size_t off = (size_t)(k0 / 8 + cc) * kLbo + (size_t)row * 16;
      *reinterpret_cast<uint4*>(at + off) = hi;
      *reinterpret_cast<uint4*>(at + kTileB + off) = lo;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u)
      c[tl][u] = (c0 && live[tl]) ? __ldg(c0 + ((int64_t)dir * B + bt[tl]) * kH + j0 + u) : 0.f;
  }
  const float inv_s = __ldg(inv_scale + dir);
  fence_proxy_async_cluster();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  mbar_wait(bar_w, 0);
  cluster_arrive();
  cluster_wait();

  const uint32_t sa_u = smem_u32(sa), sw_u = smem_u32(sw);
  uint32_t remote_a[kCl];
#pragma unroll
  for (int p = 0; p < kCl; ++p) remote_a[p] = map_to_cta(sa_u, (uint32_t)p);
  const uint32_t my_chunk_off = (uint32_t)((4 * q + 2 * half) * kLbo + row * 16);

  auto issue_mma = [&](int tl) {                     // h_(t-1) W_hh^T of tile tl -> accumulator tl (whole warp 0 calls)
    fence_proxy_async();
    tc_fence_after();
    if (elect_one()) {
      const uint32_t a_hi = sa_u + (uint32_t)tl * 2u * kTileB, a_lo = a_hi + kTileB;
      const uint32_t d_tmem = tmem + (uint32_t)tl * 128u;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t dah = make_desc_kmajor(a_hi + ks * 2 * kLbo, kLbo, kSbo);
        const uint64_t dal = make_desc_kmajor(a_lo + ks * 2 * kLbo, kLbo, kSbo);
        const uint64_t dbh = make_desc_kmajor(sw_u + ks * 2 * kLbo, kLbo, kSbo);
        const uint64_t dbl = make_desc_kmajor(sw_u + kTileB + ks * 2 * kLbo, kLbo, kSbo);
        umma_f16(d_tmem, dah, dbh, kIdescF16, ks > 0 ? 1u : 0u);
        umma_f16(d_tmem, dah, dbl, kIdescF16, 1u);
        umma_f16(d_tmem, dal, dbh, kIdescF16, 1u);
      }
      umma_commit(bar_acc + tl);
    }
    __syncwarp();
  };
  auto pre_ptr = [&](int tl, int step) {
    const int t = dir ? R - 1 - step : step;
    return pre + ((int64_t)t * B + (live[tl] ? bt[tl] : 0)) * (2 * 4 * kH) + dir * 4 * kH + (int)q * 128 + 64 * half;
  };

  uint32_t phase[2] = {0u, 0u};
  if (warp == 0) issue_mma(0);
  mbar_wait(bar_acc, phase[0]);
  phase[0] ^= 1u;
  cluster_arrive();                                  // every CTA's first MMAs have read A[0] before its first overwrite
  cluster_wait();
  const int nhs = 2 * R;
  auto half_step = [&](auto tile_c, int hs) {          // the tile is a compile-time constant: its state stays in registers
    constexpr int tl = decltype(tile_c)::value, ot = tl ^ 1;
    const int step = hs >> 1;
    const int t = dir ? R - 1 - step : step;
    float(&ct)[16] = c[tl];
    // (a) the other tile's next MMAs run underneath this epilogue
    const bool more = hs + 1 < nhs;
    if (more && warp == 0) issue_mma(ot);
    // this epilogue's input projection (L2 hit: prefetched one half-step ago), and the next one's prefetch
    float4 pv[16];
    {
      const float4* src = reinterpret_cast<const float4*>(pre_ptr(tl, step));
#pragma unroll
      for (int i = 0; i < 16; ++i) pv[i] = live[tl] ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (more) {
        const float* nx = pre_ptr(ot, (hs + 1) >> 1);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 32));
      }
    }
    // (b) epilogue of tile tl (its accumulator was awaited at the end of the previous half-step)
    tc_fence_after();
    const float* pf = reinterpret_cast<const float*>(pv);
    float* sp = (save && live[tl]) ? save + (((int64_t)dir * R + t) * B + bt[tl]) * (5 * kH) + j0 : nullptr;
    float hh[16];
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {
      uint32_t ai[8], af[8], ag[8], ao[8];
      const uint32_t tcol = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(128 * tl + 64 * half + 8 * ps);
      tmem_ld8_issue(tcol, ai);
      tmem_ld8_issue(tcol + 16u, af);
      tmem_ld8_issue(tcol + 32u, ag);
      tmem_ld8_issue(tcol + 48u, ao);
      tmem_wait4x8(ai, af, ag, ao);
      float gi[8], gf[8], gg[8], go[8];
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const int u = 8 * ps + v;
        gi[v] = sigmoid_acc(fmaf(__uint_as_float(ai[v]), inv_s, pf[u]));
        gf[v] = sigmoid_acc(fmaf(__uint_as_float(af[v]), inv_s, pf[16 + u]));
        gg[v] = tanh_acc(fmaf(__uint_as_float(ag[v]), inv_s, pf[32 + u]));
        go[v] = sigmoid_acc(fmaf(__uint_as_float(ao[v]), inv_s, pf[48 + u]));
        ct[u] = fmaf(gf[v], ct[u], gi[v] * gg[v]);
        hh[u] = go[v] * tanh_acc(ct[u]);
      }
      if (sp) {
        float4* s4 = reinterpret_cast<float4*>(sp + 8 * ps);
        s4[0] = make_float4(gi[0], gi[1], gi[2], gi[3]);
        s4[1] = make_float4(gi[4], gi[5], gi[6], gi[7]);
        s4 = reinterpret_cast<float4*>(sp + kH + 8 * ps);
        s4[0] = make_float4(gf[0], gf[1], gf[2], gf[3]);
        s4[1] = make_float4(gf[4], gf[5], gf[6], gf[7]);
        s4 = reinterpret_cast<float4*>(sp + 2 * kH + 8 * ps);
        s4[0] = make_float4(gg[0], gg[1], gg[2], gg[3]);
        s4[1] = make_float4(gg[4], gg[5], gg[6], gg[7]);
        s4 = reinterpret_cast<float4*>(sp + 3 * kH + 8 * ps);
        s4[0] = make_float4(go[0], go[1], go[2], go[3]);
        s4[1] = make_float4(go[4], go[5], go[6], go[7]);
        s4 = reinterpret_cast<float4*>(sp + 4 * kH + 8 * ps);
        s4[0] = make_float4(ct[8 * ps], ct[8 * ps + 1], ct[8 * ps + 2], ct[8 * ps + 3]);
        s4[1] = make_float4(ct[8 * ps + 4], ct[8 * ps + 5], ct[8 * ps + 6], ct[8 * ps + 7]);
      }
    }
    tc_fence_before();
    {   // h_t -> A[tile] of all four CTAs
      uint4 hi0, lo0, hi1, lo1;
      const float(&x0)[8] = *reinterpret_cast<const float(*)[8]>(&hh[0]);
      const float(&x1)[8] = *reinterpret_cast<const float(*)[8]>(&hh[8]);
      split8(x0, hi0, lo0);
      split8(x1, hi1, lo1);
      const uint32_t nb = (uint32_t)tl * 2u * kTileB + my_chunk_off;
#pragma unroll
      for (int p = 0; p < kCl; ++p) {
        const uint32_t base = remote_a[p] + nb;
        st_cluster_v4(base, hi0);
        st_cluster_v4(base + kLbo, hi1);
        st_cluster_v4(base + kTileB, lo0);
        st_cluster_v4(base + kTileB + kLbo, lo1);
      }
    }
    // (c) my MMAs on the other tile are complete before anybody may overwrite its A operand (next half-step)
    if (more) {
      mbar_wait(bar_acc + ot, phase[ot]);
      phase[ot] ^= 1u;
    }
    fence_proxy_async_cluster();
    cluster_arrive();
    if (live[tl]) {
      float4* o4 = reinterpret_cast<float4*>(out + ((int64_t)t * B + bt[tl]) * (2 * kH) + dir * kH + j0);
#pragma unroll
      for (int i = 0; i < 4; ++i) o4[i] = make_float4(hh[4 * i], hh[4 * i + 1], hh[4 * i + 2], hh[4 * i + 3]);
      if (step == R - 1) {
        float* hp = hn + ((int64_t)dir * B + bt[tl]) * kH + j0;
        float* cp = cn + ((int64_t)dir * B + bt[tl]) * kH + j0;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          hp[u] = hh[u];
          cp[u] = ct[u];
        }
      }
    }
    __syncwarp();
    cluster_wait();
  };
  for (int step = 0; step < R; ++step) {
    half_step(std::integral_constant<int, 0>{}, 2 * step);
    half_step(std::integral_constant<int, 1>{}, 2 * step + 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// One step of the backward recurrence, all gate algebra fused (thread = (encounter, hidden unit)):
//   d h = gh_out[t] + dh_rec;  d o = d h tanh(c);  d c = d h o (1 - tanh(c)^2) + dc_rec
//   d a = [ d c g i (1 - i) | d c c_prev f (1 - f) | d c i (1 - g^2) | d o o (1 - o) ];  dc_rec <- d c f
// save_t: (B, 5, 128) i f g o c of this step; c_prev: (B, stride) cell state of the previous step (c0 or save of it).
__global__ void lstm_bwd_step_kernel(const float* __restrict__ save_t, const float* __restrict__ c_prev, int64_t c_prev_stride,
                                     const float* __restrict__ gh_out, int64_t gh_stride, const float* __restrict__ dh_rec,
                                     float* __restrict__ dc_rec, float* __restrict__ da, int64_t B) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * kH) return;
  const int64_t b = idx / kH;
  const int j = (int)(idx - b * kH);
  const float* s = save_t + b * (5 * kH) + j;
  const float i = s[0], f = s[kH], g = s[2 * kH], o = s[3 * kH], c = s[4 * kH];
  const float cp = c_prev ? c_prev[b * c_prev_stride + j] : 0.f;
  const float tc_ = tanh_acc(c);
  const float dh = (gh_out ? gh_out[b * gh_stride + j] : 0.f) + dh_rec[idx];
  const float dc = fmaf(dh * o, 1.0f - tc_ * tc_, dc_rec[idx]);
  float* d = da + b * (4 * kH) + j;
  d[0] = dc * g * i * (1.0f - i);
  d[kH] = dc * cp * f * (1.0f - f);
  d[2 * kH] = dc * i * (1.0f - g * g);
  d[3 * kH] = dh * tc_ * o * (1.0f - o);
  dc_rec[idx] = dc * f;
}

// ---- input projection on the tensor cores --------------------------------------------------------------------
// pre (M, 1024) = A (M, K) Wp^T + bias for all R * B rows at once (dic_lstm_project): the float32 SIMT library GEMM that
// fed the recurrence cost 2-4x the recurrence itself (24 / 48 ms against 10 ms at B = 32,768), and the library's TF32
// tensor-core GEMMs are outside the 1e-5 parity.  Same arithmetic as the recurrence instead: operands split into fp16
// hi + lo halves (A on the fly, while it is staged; Wp once, power-of-two scaled), three tcgen05 MMAs per K step
// (hi.hi + hi.lo + lo.hi, M = 128, N = 256, K = 16), float32 accumulation in TMEM over K <= 256.
//   CTA = 128 rows of A x 256 of the 1024 output columns at a time; K in chunks of 64 (zero padded);
//   two CTAs per SM (96 KB of shared memory, 256 TMEM columns each) overlap one CTA's operand staging with the
//   other's MMAs; the epilogue adds the bias and stores whole 128-byte lines per thread.
constexpr int kPN = 256;                         // output columns per accumulator
constexpr int kPK = 64;                          // K chunk
constexpr int kPLboA = kRows * 16;               // A: 128 rows x 16 B per K chunk of 8 halves
constexpr int kPLboB = kPN * 16;                 // B: 256 columns x 16 B
constexpr int kPTileA = (kPK / 8) * kPLboA;      // 16 KB (hi or lo)
constexpr int kPTileB = (kPK / 8) * kPLboB;      // 32 KB (hi or lo)
constexpr uint32_t kIdescProj = make_idesc(0u, 128u, 256u);

struct ProjSmem {
  static constexpr size_t a = 0;                                   // [hi | lo]
  static constexpr size_t b = a + 2 * (size_t)kPTileA;             // [hi | lo]
  static constexpr size_t bars = b + 2 * (size_t)kPTileB;
  static constexpr size_t total = bars + 64;
};

// Wp (1024, K) row-major float32 -> per (column block nb, K chunk kc): the UMMA B operand [hi | lo] (fp16, scaled)
__global__ void lstm_pack_wih_kernel(const float* __restrict__ wp, unsigned char* __restrict__ packed,
                                     float* __restrict__ inv_scale, int K, int nkc) {
  __shared__ float red[32];
  float mx = 0.f;
  for (int i = threadIdx.x; i < 1024 * K; i += blockDim.x) mx = fmaxf(mx, fabsf(wp[i]));
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  int e = 0;
  if (mx > 0.f && isfinite(mx)) frexpf(mx, &e);
  const float s = ldexpf(1.0f, 13 - e);
  if (blockIdx.x == 0 && threadIdx.x == 0) inv_scale[0] = 1.0f / s;
  const int nb = blockIdx.x / nkc, kc = blockIdx.x % nkc;
  unsigned char* dst = packed + (size_t)blockIdx.x * 2 * kPTileB;
  for (int idx = threadIdx.x; idx < kPN * (kPK / 8); idx += blockDim.x) {
    const int n = idx / (kPK / 8), c = idx % (kPK / 8);
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kc * kPK + c * 8 + i;
      x[i] = k < K ? s * wp[(size_t)(nb * kPN + n) * K + k] : 0.f;
    }
    uint4 hi, lo;
    split8(x, hi, lo);
    const size_t off = (size_t)c * kPLboB + (size_t)n * 16;
    *reinterpret_cast<uint4*>(dst + off) = hi;
    *reinterpret_cast<uint4*>(dst + kPTileB + off) = lo;
  }
}

__global__ void __launch_bounds__(kThreads, 2)
lstm_project_kernel(const float* __restrict__ A, int64_t lda, const unsigned char* __restrict__ packed,
                    const float* __restrict__ inv_scale, const float* __restrict__ bias, float* __restrict__ out,
                    int64_t M, int K, int nkc, int relu) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sa = smem + ProjSmem::a;
  unsigned char* sb = smem + ProjSmem::b;
  uint64_t* bar_b = reinterpret_cast<uint64_t*>(smem + ProjSmem::bars);
  uint64_t* bar_acc = bar_b + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_b + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(bar_b, 1);
    mbar_init(bar_acc, 1);
    fence_proxy_async();
  }
  if (warp == 0) tmem_alloc(tmem_slot, kPN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float inv_s = __ldg(inv_scale);
  const uint32_t sa_u = smem_u32(sa), sb_u = smem_u32(sb);
  const int64_t mtiles = (M + kRows - 1) / kRows;
  const bool vec4 = (K % 4 == 0) && (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15u) == 0);
  uint32_t ph_b = 0, ph_acc = 0;
  for (int64_t job = blockIdx.x; job < mtiles * 4; job += gridDim.x) {
    const int64_t mt = job >> 2;
    const int nb = (int)(job & 3);
    const int64_t r0 = mt * kRows;
    for (int kc = 0; kc < nkc; ++kc) {
      if (tid == 0) {                                // the weight chunk: one 64 KB bulk copy (L2 resident after the first tile)
        mbar_expect_tx(bar_b, 2u * kPTileB);
        bulk_g2s(sb, packed + (size_t)(nb * nkc + kc) * 2 * kPTileB, 2u * kPTileB, bar_b);
      }
      // A chunk: 128 rows x 64 floats -> fp16 hi / lo in the UMMA layout; item = (row, 8-float K chunk)
      for (int item = tid; item < kRows * (kPK / 8); item += kThreads) {
        const int c = item & 7, row = item >> 3;     // 8 consecutive lanes read 256 contiguous bytes of one row
        const int64_t gr = r0 + row;
        const int k0 = kc * kPK + c * 8;
        float x[8];
        if (gr < M && vec4 && k0 + 8 <= K) {
          const float4 v0 = __ldg(reinterpret_cast<const float4*>(A + gr * lda + k0));
          const float4 v1 = __ldg(reinterpret_cast<const float4*>(A + gr * lda + k0) + 1);
          x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = (gr < M && k0 + i < K) ? __ldg(A + gr * lda + k0 + i) : 0.f;
        }
        if (relu) {
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = fmaxf(x[i], 0.f);
        }
        uint4 hi, lo;
        split8(x, hi, lo);
        const size_t off = (size_t)c * kPLboA + (size_t)row * 16;
        *reinterpret_cast<uint4*>(sa + off) = hi;
        *reinterpret_cast<uint4*>(sa + kPTileA + off) = lo;
      }
      fence_proxy_async();
      __syncthreads();
      if (warp == 0) {
        mbar_wait(bar_b, ph_b);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < kPK / 16; ++ks) {
            const uint64_t dah = make_desc_kmajor(sa_u + ks * 2 * kPLboA, kPLboA, kSbo);
            const uint64_t dal = make_desc_kmajor(sa_u + kPTileA + ks * 2 * kPLboA, kPLboA, kSbo);
            const uint64_t dbh = make_desc_kmajor(sb_u + ks * 2 * kPLboB, kPLboB, kSbo);
            const uint64_t dbl = make_desc_kmajor(sb_u + kPTileB + ks * 2 * kPLboB, kPLboB, kSbo);
            umma_f16(tmem, dah, dbh, kIdescProj, (kc > 0 || ks > 0) ? 1u : 0u);
            umma_f16(tmem, dah, dbl, kIdescProj, 1u);
            umma_f16(tmem, dal, dbh, kIdescProj, 1u);
          }
          umma_commit(bar_acc);
        }
        __syncwarp();
      }
      ph_b ^= 1u;
      mbar_wait(bar_acc, ph_acc);                    // the MMAs have read both operand buffers: they may be refilled
      ph_acc ^= 1u;
    }
    // epilogue: thread = (row, 128 of the 256 columns): + bias, whole 128-byte lines per thread
    tc_fence_after();
    {
      const int row = 32 * (warp & 3) + lane, half = warp >> 2;
      const int64_t gr = r0 + row;
      float* orow = out + gr * 1024 + nb * kPN + half * 128;
      const float* brow = bias + nb * kPN + half * 128;
#pragma unroll 1
      for (int cb = 0; cb < 4; ++cb) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(half * 128 + cb * 32), v);
        if (gr < M) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(brow + cb * 32) + i);
            float4 o;
            o.x = fmaf(__uint_as_float(v[4 * i]), inv_s, b4.x);
            o.y = fmaf(__uint_as_float(v[4 * i + 1]), inv_s, b4.y);
            o.z = fmaf(__uint_as_float(v[4 * i + 2]), inv_s, b4.z);
            o.w = fmaf(__uint_as_float(v[4 * i + 3]), inv_s, b4.w);
            reinterpret_cast<float4*>(orow + cb * 32)[i] = o;
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();                                 // the accumulator is drained before the next job overwrites it
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem, kPN);
}

}  // namespace
}  // namespace dic

using namespace dic;

extern "C" size_t dic_lstm_packed_bytes(void) { return (size_t)2 * kCl * 2 * kTileB + 64; }

extern "C" int dic_lstm_pack_whh(const float* w_hh, const float* w_hh_reverse, void* packed, dic_stream_t stream) {
  DIC_REQUIRE(w_hh && w_hh_reverse && packed, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE(aligned16(packed), DIC_ERR_INVALID_ARGUMENT, "packed must be 16-byte aligned");
  unsigned char* p = static_cast<unsigned char*>(packed);
  float* inv_scale = reinterpret_cast<float*>(p + (size_t)2 * kCl * 2 * kTileB);
  lstm_pack_whh_kernel<<<dim3(kCl, 2), 256, 0, as_stream(stream)>>>(w_hh, w_hh_reverse, p, inv_scale);
  DIC_LAUNCH_CHECK("lstm_pack_whh_kernel");
  return DIC_OK;
}

extern "C" int dic_lstm_fwd(const float* pre, const void* packed, const float* h0, const float* c0, float* out, float* hn,
                            float* cn, float* save, int R, int64_t B, int hidden, dic_stream_t stream) {
  DIC_REQUIRE(hidden == kH, DIC_ERR_UNSUPPORTED, "hidden size %d: the persistent kernel covers hidden = %d "
              "(pretrain_interp.py:96)", hidden, kH);
  DIC_REQUIRE(R > 0 && B >= 0, DIC_ERR_INVALID_ARGUMENT, "bad sizes R=%d B=%lld", R, (long long)B);
  DIC_REQUIRE(packed && ((pre && out && hn && cn) || B == 0), DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE(aligned16(pre) && aligned16(out) && aligned16(packed) && (!save || aligned16(save)), DIC_ERR_INVALID_ARGUMENT,
              "pre / out / save / packed must be 16-byte aligned");
  if (B == 0) return DIC_OK;
  const int64_t tiles = (B + kRows - 1) / kRows;
  DIC_REQUIRE(tiles * 2 * kCl <= 2147483647LL, DIC_ERR_UNSUPPORTED, "B=%lld exceeds the grid limit", (long long)B);
  const unsigned char* p = static_cast<const unsigned char*>(packed);
  const float* inv_scale = reinterpret_cast<const float*>(p + (size_t)2 * kCl * 2 * kTileB);
  if (B > kRows) {       // two interleaved 128-encounter tiles per cluster
    const int64_t pairs = (B + 2 * kRows - 1) / (2 * kRows);
    DIC_CUDA(cudaFuncSetAttribute(lstm_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lstm2Smem::total));
    lstm_fwd2_kernel<<<(unsigned)(pairs * 2 * kCl), kThreads, Lstm2Smem::total, as_stream(stream)>>>(
        pre, p, inv_scale, h0, c0, out, hn, cn, save, R, B);
    DIC_LAUNCH_CHECK("lstm_fwd2_kernel");
    return DIC_OK;
  }
  DIC_CUDA(cudaFuncSetAttribute(lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LstmSmem::total));
  lstm_fwd_kernel<<<(unsigned)(tiles * 2 * kCl), kThreads, LstmSmem::total, as_stream(stream)>>>(
      pre, p, inv_scale, h0, c0, out, hn, cn, save, R, B);
  DIC_LAUNCH_CHECK("lstm_fwd_kernel");
  return DIC_OK;
}

extern "C" int dic_lstm_bwd_step(const float* save_t, const float* c_prev, int64_t c_prev_stride, const float* gh_out,
                                 int64_t gh_stride, const float* dh_rec, float* dc_rec, float* da, int64_t B, int hidden,
                                 dic_stream_t stream) {
  DIC_REQUIRE(hidden == kH, DIC_ERR_UNSUPPORTED, "hidden size %d (covered: %d)", hidden, kH);
  DIC_REQUIRE(B >= 0 && ((save_t && dh_rec && dc_rec && da) || B == 0), DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  if (B == 0) return DIC_OK;
  const int64_t n = B * kH;
  lstm_bwd_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(save_t, c_prev, c_prev_stride, gh_out,
                                                                                  gh_stride, dh_rec, dc_rec, da, B);
  DIC_LAUNCH_CHECK("lstm_bwd_step_kernel");
  return DIC_OK;
}

extern "C" size_t dic_lstm_project_packed_bytes(int K) {
  if (K <= 0) return 0;
  const int nkc = (K + kPK - 1) / kPK;
  return (size_t)4 * nkc * 2 * kPTileB + 64;
}

extern "C" int dic_lstm_pack_wih(const float* wp, void* packed, int K, dic_stream_t stream) {
  DIC_REQUIRE(wp && packed && K > 0 && K <= 1024, DIC_ERR_INVALID_ARGUMENT, "bad arguments (K=%d)", K);
  DIC_REQUIRE(aligned16(packed), DIC_ERR_INVALID_ARGUMENT, "packed must be 16-byte aligned");
  const int nkc = (K + kPK - 1) / kPK;
  unsigned char* p = static_cast<unsigned char*>(packed);
  float* inv_scale = reinterpret_cast<float*>(p + (size_t)4 * nkc * 2 * kPTileB);
  lstm_pack_wih_kernel<<<4 * nkc, 256, 0, as_stream(stream)>>>(wp, p, inv_scale, K, nkc);
  DIC_LAUNCH_CHECK("lstm_pack_wih_kernel");
  return DIC_OK;
}

extern "C" int dic_lstm_project(const float* A, int64_t lda, const void* packed, const float* bias, float* out, int64_t M,
                                int K, int relu, dic_stream_t stream) {
  DIC_REQUIRE(M >= 0 && K > 0 && K <= 1024 && lda >= K, DIC_ERR_INVALID_ARGUMENT, "bad sizes M=%lld K=%d lda=%lld",
              (long long)M, K, (long long)lda);
  DIC_REQUIRE(packed && bias && ((A && out) || M == 0), DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE(aligned16(out) && aligned16(packed) && aligned16(bias), DIC_ERR_INVALID_ARGUMENT,
              "out / packed / bias must be 16-byte aligned");
  if (M == 0) return DIC_OK;
  const int nkc = (K + kPK - 1) / kPK;
  const unsigned char* p = static_cast<const unsigned char*>(packed);
  const float* inv_scale = reinterpret_cast<const float*>(p + (size_t)4 * nkc * 2 * kPTileB);
  int dev = 0, sms = 148;
  DIC_CUDA(cudaGetDevice(&dev));
  DIC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t jobs = ((M + kRows - 1) / kRows) * 4;
  const int grid = (int)(jobs < 2LL * sms ? jobs : 2LL * sms);
  DIC_CUDA(cudaFuncSetAttribute(lstm_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ProjSmem::total));
  lstm_project_kernel<<<grid, kThreads, ProjSmem::total, as_stream(stream)>>>(A, lda, p, inv_scale, bias, out, M, K, nkc,
                                                                             relu);
  DIC_LAUNCH_CHECK("lstm_project_kernel");
  return DIC_OK;
}
