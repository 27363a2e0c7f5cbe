// SingleChannelInterp forward / backward for sm_100a.
//
// Reference: interpolation_layer.py:31-86 (forward); the backward is what autograd
// derives from it (closed form in SURVEY.md Appendix A.1).
//
// Per (encounter b, vital c, reference point r) with alpha = softplus(kernel[c]):
//     s_t  = -alpha (d_t - r)^2 + log m_t
//     w    = logsumexp_t s_t                       (log-intensity)
//     y    = sum_t softmax_t(s) x_t                (low-pass)
//     y'   = the same with 10 alpha                (high-pass, kappa = 10)
//
// Kernel design (one CTA = one encounter, one warp-task = one vital x 32*RPT grid points)
//   * rows are staged with one TMA bulk copy and canonicalised (interp_stage.cuh): a vectorised
//     check accepts the pipeline's left-packed time-ordered rows as they are, anything else is
//     compacted and sorted;
//   * the softmax shift is known without a pass over the data: the maximum of s_t is at the
//     observation nearest to r, found by binary search in the sorted times (d*);
//   * the exponent -(alpha log2 e) ((d-r)^2 - (d*-r)^2) is formed as fma(delta, delta, -hi) - lo
//     with (d*-r)^2 = hi + lo carried as an exact two-float: the difference of squares is
//     rounded once, so the high-pass weight e^10 stays ~30x closer to the float64 truth than the
//     reference's own float32 evaluation (which subtracts numbers in the hundreds);
//   * both filters share the shift, so the high-pass exponent is 10x the low-pass one: one MUFU.EX2 per
//     (t, r) in the outer window, a second one (2^(10 arg)) only inside the narrow high-pass window;
//   * sliding windows: a Gaussian weight below 2^-kCut of the largest one is dropped, so a lane
//     only walks the observations within +-sqrt(nmin + kCut/a) hours of its grid points for the
//     low-pass sums, and in a second, short loop those within +-sqrt(nmin + kCut/(10a)) for the
//     high-pass sums (make_window2: one warp-uniform trip count per loop);
//   * accumulators stay in registers, each lane owns RPT ADJACENT grid points, observations are
//     read as 128-bit shared loads; vitals are dealt to warps heaviest-first in snake order so
//     the warps of a CTA finish together.
#include <type_traits>

#include "interp_stage.cuh"

namespace dic {
namespace {

constexpr int kMaxWarps = 8;
constexpr float kExactAbove = 12.0f;   // ak (d* - r)^2 above which a filter evaluates its exponents in float64 (sci_fwd_task)
constexpr float kCut = 24.0f;      // log2 of the dropped weight ratio: the dropped tail is < 2e-8 of S and < 5e-7 of the moment sums

struct SciSmem {
  // dynamic shared memory layout: bar | null chunks | rows[3][C][Tp] | n_valid[C] | order[C] | part[] | vpar | lbtab
  uint64_t* bar;
  float* nullc;   // [0..3] = kPadTime (a chunk of observations that weigh exactly 0), [4..7] = 0
  float* rows;
  int* n_valid;   // > 0: binary mask; < 0: weighted (|n| entries); 0: all masked
  int* order;     // vitals sorted by descending observation count
  float* part;
  float* vpar;    // [C][4] per-vital constants: alpha, a = alpha log2 e, kCut / a, kCut / (10 a)
  int* lbtab;     // [C][R + 1]: lbtab[c][j] = #{t : d_t < r_j} (j < R), lbtab[c][R] = n  (forward kernel only)
};

__device__ __forceinline__ SciSmem sci_carve(unsigned char* base, int C, int Tp, int R) {
  SciSmem s;
  s.bar = reinterpret_cast<uint64_t*>(base);
  s.nullc = reinterpret_cast<float*>(base + 16);
  s.rows = reinterpret_cast<float*>(base + 48);
  s.n_valid = reinterpret_cast<int*>(s.rows + 3 * C * Tp);
  s.order = s.n_valid + C;
  s.part = reinterpret_cast<float*>(s.order + C);
  s.vpar = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(s.part + C * ((R + 31) / 32)) + 15) & ~(uintptr_t)15);
  s.lbtab = reinterpret_cast<int*>(s.vpar + 4 * C);
  return s;
}

static size_t sci_smem_bytes(int C, int Tp, int R) {
  return 48 + sizeof(float) * (3 * (size_t)C * Tp) + 2 * sizeof(int) * C +
         sizeof(float) * (size_t)C * ((R + 31) / 32) + 16 + sizeof(float) * 4 * C + sizeof(int) * (size_t)C * (R + 1);
}

// Stage + canonicalise one encounter.  After return (CTA-synchronised):
//   rows[0][c] = m*x, rows[1][c] = m, rows[2][c] = d, compacted, sorted by time and padded to a multiple of 4 with
//   null entries; n_valid[c]; order[]; vpar[c] (per-vital constants, computed ONCE per encounter while the bulk copy
//   is in flight); lbtab[c][j] = #{t : d_t < r_j} for every grid point (the binary searches of the whole encounter,
//   done by the warp that canonicalised the row) - the tasks read the nearest observation AND their window bounds
//   from this table.  Returns true when the grid is uniform (ref_t[j] = r0 + j h within 1 % of h).
template <int RPT>
__device__ __forceinline__ bool sci_stage(const SciSmem& s, const float* xb, const float* __restrict__ kernel,
                                          const float* __restrict__ ref_t, int C, int T, int Tp, int R, bool use_tma) {
  if (threadIdx.x < 8) s.nullc[threadIdx.x] = threadIdx.x < 4 ? kPadTime : 0.f;
  stage_rows_issue(s.rows, xb, 3 * C, T, Tp, s.bar, use_tma);
  // ---- overlapped with the copy: per-vital constants, regularity of the grid ----
  if (threadIdx.x < C) {
    const float alpha = softplus_ref(__ldg(kernel + threadIdx.x));
    const float a = alpha * kLog2e;
    reinterpret_cast<float4*>(s.vpar)[threadIdx.x] = make_float4(alpha, a, kCut / a, kCut / (10.f * a));
  }
  const float r0 = __ldg(ref_t), rl = __ldg(ref_t + R - 1);
  const float h = R > 1 ? (rl - r0) / (float)(R - 1) : 1.0f;
  int irregular = !(h > 0.f);
  for (int j = threadIdx.x; j < R; j += blockDim.x)
    irregular |= fabsf(__ldg(ref_t + j) - (r0 + h * (float)j)) > 0.01f * h;
  stage_rows_wait(s.bar, use_tma);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int chunks = (R + 32 * RPT - 1) / (32 * RPT);
  for (int c = warp; c < C; c += nwarps) {
    float* sx = s.rows + (0 * C + c) * Tp;
    float* sm = s.rows + (1 * C + c) * Tp;
    float* sd = s.rows + (2 * C + c) * Tp;
    int n = warp_canonical_count(sm, sd, Tp, lane);
    int weighted = 0;
    if (n < 0) {   // general rows: drop masked entries, sort by time, detect fractional weights
      n = warp_compact3(sx, sm, sd, T, lane, /*fold_mask=*/true);
      warp_sort3(sd, sx, sm, n, lane);
      for (int t = lane; t < n; t += 32) weighted |= (sm[t] != 1.0f);
      weighted = __any_sync(0xffffffffu, weighted);
    }
    warp_pad4_far(sd, sx, sm, n, lane);
    if (lane == 0) s.n_valid[c] = weighted ? -n : n;
    int* tab = s.lbtab + c * (R + 1);
    if (n > 0) {
      for (int ch = 0; ch < chunks; ++ch) {
        float rr[RPT];
        bool upk[RPT];
        int lb[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          rr[k] = __ldg(ref_t + min((ch * 32 + lane) * RPT + k, R - 1));
          upk[k] = false;
        }
        multi_bound<RPT>(sd, n, rr, upk, lb);
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          const int j = (ch * 32 + lane) * RPT + k;
          if (j < R) tab[j] = lb[k];
        }
      }
    }
    if (lane == 0) tab[R] = n;
  }
  irregular = __syncthreads_or(irregular);
  if (threadIdx.x == 0) {   // insertion sort of C indices by descending count (C is ~6)
    for (int c = 0; c < C; ++c) {
      const int key = abs(s.n_valid[c]);
      int j = c;
      while (j > 0 && abs(s.n_valid[s.order[j - 1]]) < key) {
        s.order[j] = s.order[j - 1];
        --j;
      }
      s.order[j] = c;
    }
  }
  __syncthreads();
  return !irregular;
}

// Task k of `ntasks` (heaviest vital first) in snake order over the warps.
__device__ __forceinline__ int snake_task(int round, int warp, int nwarps) {
  return round * nwarps + ((round & 1) ? (nwarps - 1 - warp) : warp);
}

// Window bounds from the per-vital table of lower bounds (uniform grid): the observations within +-w of the lane's
// grid points [r_first, r_last] lie between the grid points that bracket r_first - w and r_last + w, so four table
// reads replace four 8-step binary searches; the bracket is at most one grid step (~1 observation) wider per side,
// which the 4-entry chunk rounding mostly absorbs.
__device__ __forceinline__ void table_range(const int* __restrict__ tab, int n, int R, float lo, float hi, float r0,
                                            float inv_h, int& start, int& end) {
  const float f0 = (lo - r0) * inv_h - 0.02f, f1 = (hi - r0) * inv_h + 0.02f;     // 0.02: the grid's 1 % tolerance
  start = f0 < 0.f ? 0 : tab[min(__float2int_rd(f0), R)];               // #{d < r_j}, r_j <= lo
  const int j1 = __float2int_ru(f1) + 1;                                // #{d < r_(j+1)} >= #{d <= hi}
  end = j1 >= R ? n : tab[max(j1, 0)];
}

// Forward task.  Per filter it accumulates S = sum e and SY = sum e x (y = SY / S) and, when the backward pass
// will run (MOM), the two moments that carry the whole kernel gradient:
//     Q0 = sum t e,  Q1 = sum t e x,   t = (d - r)^2 - (d* - r)^2 >= 0   (the shifted squared distance)
// With n_t = t + (d* - r)^2:
//     sum_t n_t e_t (x_t - y) = Q1 - y Q0 + (d* - r)^2 (SY - y S) = Q1 - y Q0        (SY - y S == 0 by definition)
//     sum_t n_t e_t           = Q0 + (d* - r)^2 S
// so d alpha needs no second sweep over the observations (Appendix A.1):
//     d alpha_c = - sum_r [ gy U1 + gw U0 + gy' U1' ],  U1 = (Q1 - y Q0)/S,  U0 = Q0/S + (d*-r)^2,
//                                                       U1' = 10 (Q1' - y' Q0')/S'
// The three U rows are what `stats` holds.  The term that made a direct sum of n e (x A - y A) cancel
// catastrophically when one observation dominates - (d*-r)^2 sum e (x - y) - is the one that vanishes
// analytically here and is never formed; the dominant observation itself has t = 0.
template <int RPT, bool WEIGHTED, bool MOM>
__device__ __forceinline__ void sci_fwd_task(const float* __restrict__ sx, const float* __restrict__ sm,
                                             const float* __restrict__ sd, int n, const float4 par, int chunk,
                                             int lane, int c, int C, int R, const float* __restrict__ ref_t,
                                             float* __restrict__ ub, float* __restrict__ sb,
                                             const float* __restrict__ nullc, const int* __restrict__ tab,
                                             bool regular, float r0, float inv_h) {
  const float alpha = par.x, a = par.y, na = -a;
  int ridx[RPT];
  float rr[RPT], nhi[RPT], dstar[RPT], dst[RPT];
  float nmax = 0.f;
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    ridx[k] = (chunk * 32 + lane) * RPT + k;
    const int rc = min(ridx[k], R - 1);
    rr[k] = __ldg(ref_t + rc);
    const int is = nearest_index(sd, n, tab[rc], rr[k]);               // tab[rc] = first observation at or after r
    dstar[k] = sd[is];
    dst[k] = dstar[k] - rr[k];                                       // delta* = d* - r
    nhi[k] = dst[k] * dst[k];
    nmax = fmaxf(nmax, nhi[k]);
  }
  Window2 w;
  const float wo = sqrtf(nmax + par.z), wi = sqrtf(nmax + par.w);
  if (WEIGHTED) {
    w.ob = w.ib = 0;
    w.ot = w.it = (n + 3) & ~3;
  } else if (regular) {
    int so, eo, si, ei;
    table_range(tab, n, R, rr[0] - wo, rr[RPT - 1] + wo, r0, inv_h, so, eo);
    table_range(tab, n, R, rr[0] - wi, rr[RPT - 1] + wi, r0, inv_h, si, ei);
    w.ob = so & ~3;
    w.ib = si & ~3;
    w.ot = (warp_max_i(eo - w.ob) + 3) & ~3;
    w.it = (warp_max_i(ei - w.ib) + 3) & ~3;
  } else {
    w = make_window2(sd, n, rr[0], rr[RPT - 1], wo, wi, false);
  }
  const int n4 = (n + 3) & ~3;

  // One filter: HIGH = false walks the outer window with exponent -a t, HIGH = true the inner window with -10 a t
  // (same shift).  One MUFU.EX2 per pair.
  // The shifted squared distance t = (d - r)^2 - (d* - r)^2 is formed as a PRODUCT, t = (d - d*) ((d - d*) + 2 (d* - r)):
  // d - d* is exact or rounded relative to itself, so t carries ~2e-7 relative error wherever the grid point lies.  (The
  // difference of squares fma(d - r, d - r, -(d* - r)^2) inherits the absolute rounding of d - r times 2 |d - r|: for a
  // sparse vital whose nearest observations are 8 h from the grid point that is 1.5e-5 on t and, times 10 alpha, 5e-5 on
  // the high-pass output - found on the 4,096-encounter c2 slice, tests/test_gpu_interp.py.)  The exponent comes out in
  // the same three packed instructions: dd = d - d*, q = fma(dd, -ak, -ak 2 (d* - r)), arg = dd q = -ak t.
  // Packed float32x2 arithmetic over pairs of consecutive observations (two aligned register pairs per 128-bit
  // load): the loop is issue bound, and FFMA2 / FADD2 / FMUL2 halve the issue slots of its arithmetic.
  // The moments are accumulated on arg = -ak t (Q0 = sum arg e, Q1 = sum arg e x) and divided by -ak in the finish.
  auto filter = [&](auto high, float (&S)[RPT], float (&SC)[RPT], float (&Q0)[RPT], float (&Q1)[RPT]) {
    constexpr bool HIGH = decltype(high)::value;
    const float nak = HIGH ? 10.f * na : na;
    const f2_t nak2 = pack2(nak, nak);
    f2_t nds2[RPT], nck2[RPT], S2[RPT], SC2[RPT], Q02[RPT], Q12[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const float nc = nak * (2.f * dst[k]);
      nds2[k] = pack2(-dstar[k], -dstar[k]);
      nck2[k] = pack2(nc, nc);
      S2[k] = SC2[k] = Q02[k] = Q12[k] = pack2(0.f, 0.f);
    }
    const int base = HIGH ? w.ib : w.ob, trip = HIGH ? w.it : w.ot;
#pragma unroll 2
    for (int t0 = 0; t0 < trip; t0 += 4) {
      const int t = base + t0;
      const bool in_row = (unsigned)t < (unsigned)n4;           // chunks off the row weigh nothing
      const ulonglong2 d4 = *reinterpret_cast<const ulonglong2*>(in_row ? sd + t : nullc);
      const ulonglong2 x4 = *reinterpret_cast<const ulonglong2*>(in_row ? sx + t : nullc + 4);
      const f2_t dd[2] = {d4.x, d4.y}, xx[2] = {x4.x, x4.y};
      f2_t mm[2] = {0, 0};
      if (WEIGHTED) {
        const ulonglong2 m4 = *reinterpret_cast<const ulonglong2*>(sm + t);   // full range: always in the row
        mm[0] = m4.x; mm[1] = m4.y;
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          const f2_t dl = add2(dd[j], nds2[k]);                  // d - d*
          const f2_t q = fma2(dl, nak2, nck2[k]);                // -ak ((d - d*) + 2 (d* - r))
          const f2_t arg = mul2(dl, q);                          // -ak ((d - r)^2 - (d* - r)^2)
          float a0, a1;
          unpack2(arg, a0, a1);
          const f2_t e = pack2(ex2_approx(a0), ex2_approx(a1));
          // rows with fractional weights hold x m, and the weight of the pair is m e
          S2[k] = WEIGHTED ? fma2(mm[j], e, S2[k]) : add2(S2[k], e);
          SC2[k] = fma2(e, xx[j], SC2[k]);
          if (MOM) {
            const f2_t te = mul2(arg, e);
            Q02[k] = WEIGHTED ? fma2(mm[j], te, Q02[k]) : add2(Q02[k], te);
            Q12[k] = fma2(te, xx[j], Q12[k]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      S[k] = sum2(S2[k]);
      SC[k] = sum2(SC2[k]);
      Q0[k] = MOM ? sum2(Q02[k]) : 0.f;
      Q1[k] = MOM ? sum2(Q12[k]) : 0.f;
    }
  };

  // The same sums with the shifted squared distance evaluated in float64, for lanes whose NEAREST observation is far
  // from the grid point (sparse vitals, gaps of hours).  There float32 cannot hold t: an observation at the same distance
  // on the other side of the grid point has weight ~1 and t = (d - d*) (d + d* - 2 r) with a factor that cancels to the
  // rounding of d - r itself, ~ulp(d* - r) |d - d*| = 1e-6 on t at 3 h, 5e-6 at 8 h, times 10 alpha on the high-pass
  // exponent (measured: 4e-5 / 5e-5 on y' against the float64 reference, c1 and the 4,096-encounter c2 slice).  The float32
  // error of the exponent is bounded by 2.4e-7 ak max (d* - r)^2, so a filter keeps the packed float32 loop while
  // ak max (d* - r)^2 <= 12 (3e-6 on the exponent; ak = a or 10 a): for the low-pass sums that is a nearest observation
  // within ~2.9 h, for the high-pass sums within ~0.9 h.  A warp without a far lane never enters this loop, and the rows
  // that do are the short ones (measured cost at c2: +3 % of the kernel).
  const bool far_low = !WEIGHTED && a * nmax > kExactAbove, far_high = !WEIGHTED && 10.f * a * nmax > kExactAbove;
  const bool any_far_low = __any_sync(0xffffffffu, far_low), any_far_high = __any_sync(0xffffffffu, far_high);
  auto filter_exact = [&](auto high, float (&S)[RPT], float (&SC)[RPT], float (&Q0)[RPT], float (&Q1)[RPT]) {
    constexpr bool HIGH = decltype(high)::value;
    const bool far = HIGH ? far_high : far_low;
    const double nakd = HIGH ? 10.0 * (double)na : (double)na;
    double v2[RPT], rd[RPT];
    float s_[RPT], sc_[RPT], q0_[RPT], q1_[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      rd[k] = (double)rr[k];
      const double v = (double)dstar[k] - rd[k];
      v2[k] = v * v;
      s_[k] = sc_[k] = q0_[k] = q1_[k] = 0.f;
    }
    const int base = HIGH ? w.ib : w.ob, trip = HIGH ? w.it : w.ot;
    if (far) {
      for (int t = max(base, 0); t < min(base + trip, n); ++t) {
        const double d = (double)sd[t];
        const float x = sx[t];
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          const double u = d - rd[k];
          const float arg = (float)(nakd * fma(u, u, -v2[k]));
          const float e = ex2_approx(arg);
          s_[k] += e;
          sc_[k] = fmaf(e, x, sc_[k]);
          if (MOM) {
            const float te = arg * e;
            q0_[k] += te;
            q1_[k] = fmaf(te, x, q1_[k]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        S[k] = s_[k]; SC[k] = sc_[k]; Q0[k] = q0_[k]; Q1[k] = q1_[k];
      }
    }
  };

  float S[RPT], SC[RPT], Q0[RPT], Q1[RPT];
  const float inv_na = __frcp_rn(na);
  filter(std::false_type{}, S, SC, Q0, Q1);
  if (any_far_low) filter_exact(std::false_type{}, S, SC, Q0, Q1);
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    if (ridx[k] < R) {
      const float inv = __frcp_rn(S[k]), yc = SC[k] * inv;
      ub[(0 * C + c) * R + ridx[k]] = yc;
      // log S through MUFU.LG2 (absolute error ~2e-7 on log S <= ~6; w itself is -alpha n_min + log S)
      ub[(1 * C + c) * R + ridx[k]] = fmaf(__log2f(S[k]), kLn2, -alpha * nhi[k]);
      if (MOM) {          // the sums are over arg = -a t: back to t with 1 / (-a)
        const float invt = inv * inv_na;
        sb[(0 * C + c) * R + ridx[k]] = fmaf(-yc, Q0[k], Q1[k]) * invt;         // U1
        sb[(1 * C + c) * R + ridx[k]] = fmaxf(fmaf(Q0[k], invt, nhi[k]), 0.f);  // U0 >= 0
      }
    }
  }
  filter(std::true_type{}, S, SC, Q0, Q1);
  if (any_far_high) filter_exact(std::true_type{}, S, SC, Q0, Q1);
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    if (ridx[k] < R) {
      const float inv = __frcp_rn(S[k]), yc = SC[k] * inv;
      ub[(2 * C + c) * R + ridx[k]] = yc;
      if (MOM) sb[(2 * C + c) * R + ridx[k]] = fmaf(-yc, Q0[k], Q1[k]) * inv * inv_na;   // U1' = 10 (..) / (-10 a S')
    }
  }
}

// MOM: also produce the three gradient-moment rows (stats (B, 3C, R) = [U1 | U0 | U1']) for dic_sci_bwd.
template <int RPT, bool MOM>
__global__ void __launch_bounds__(kMaxWarps * 32)          // (a 64-register cap spills and measures slower)
sci_fwd_kernel(const float* __restrict__ x, const float* __restrict__ kernel,
               const float* __restrict__ ref_t, float* __restrict__ u, float* __restrict__ stats,
               int C, int T, int Tp, int R, int use_tma, int64_t x_stride) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const SciSmem s = sci_carve(smem_raw, C, Tp, R);
  const int64_t b = blockIdx.x;
  const bool regular = sci_stage<RPT>(s, x + b * x_stride, kernel, ref_t, C, T, Tp, R, use_tma != 0);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int chunks = (R + 32 * RPT - 1) / (32 * RPT);
  const int ntasks = C * chunks;
  float* ub = u + b * (int64_t)(3 * C) * R;
  float* sb = MOM ? stats + b * (int64_t)(3 * C) * R : nullptr;
  const float r0 = __ldg(ref_t);
  const float inv_h = R > 1 ? (float)(R - 1) / (__ldg(ref_t + R - 1) - r0) : 1.0f;

  for (int round = 0;; ++round) {
    const int task = snake_task(round, warp, nwarps);
    if (round * nwarps >= ntasks) break;
    if (task >= ntasks) continue;
    const int vi = chunks == 1 ? task : task / chunks, chunk = chunks == 1 ? 0 : task - vi * chunks;
    const int c = s.order[vi];
    const float* sx = s.rows + (0 * C + c) * Tp;
    const float* sm = s.rows + (1 * C + c) * Tp;
    const float* sd = s.rows + (2 * C + c) * Tp;
    const int nv = s.n_valid[c];
    const float4 par = reinterpret_cast<const float4*>(s.vpar)[c];
    const int* tab = s.lbtab + c * (R + 1);
    if (nv > 0) {
      sci_fwd_task<RPT, false, MOM>(sx, sm, sd, nv, par, chunk, lane, c, C, R, ref_t, ub, sb, s.nullc, tab, regular,
                                    r0, inv_h);
    } else if (nv < 0) {
      sci_fwd_task<RPT, true, MOM>(sx, sm, sd, -nv, par, chunk, lane, c, C, R, ref_t, ub, sb, s.nullc, tab, regular,
                                   r0, inv_h);
    } else {
      // all-masked vital: the reference yields w = -inf, y = y' = NaN (logsumexp of -inf)
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        const int r = (chunk * 32 + lane) * RPT + k;
        if (r < R) {
          ub[(0 * C + c) * R + r] = __int_as_float(0x7fc00000);
          ub[(1 * C + c) * R + r] = -INFINITY;
          ub[(2 * C + c) * R + r] = __int_as_float(0x7fc00000);
          if (MOM) {        // U0 < 0 marks "no contribution" (the reference's gradient is NaN there; U0 >= 0 otherwise)
            sb[(0 * C + c) * R + r] = 0.f;
            sb[(1 * C + c) * R + r] = -1.f;
            sb[(2 * C + c) * R + r] = 0.f;
          }
        }
      }
    }
  }
}

// Backward: d alpha_c = - sum_r [ gy U1 + gw U0 + gy' U1' ] from the moment rows the forward pass saved
// (see sci_fwd_task): one read of stats and grad_u, no second sweep over the observations.  One warp per
// encounter, lanes stride over the grid points of each vital (coalesced rows), fixed-order shuffle reduction.
constexpr int kSciBwdWarps = 8;

__global__ void __launch_bounds__(kSciBwdWarps * 32)
sci_bwd_kernel(const float* __restrict__ stats, const float* __restrict__ grad_u, float* __restrict__ partial,
               int64_t B, int C, int R, int vec4) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * kSciBwdWarps + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* sb = stats + b * (int64_t)(3 * C) * R;
  const float* gb = grad_u + b * (int64_t)(3 * C) * R;
  if (vec4) {      // R % 4 == 0 and 16-byte aligned bases: 128-bit loads, all six rows of a vital in flight at once
    const int R4 = R >> 2;
    for (int c = 0; c < C; ++c) {
      float acc = 0.f;
      for (int i = lane; i < R4; i += 32) {
        const float4 u1 = __ldg(reinterpret_cast<const float4*>(sb + (0 * C + c) * R) + i);
        const float4 u0 = __ldg(reinterpret_cast<const float4*>(sb + (1 * C + c) * R) + i);
        const float4 u2 = __ldg(reinterpret_cast<const float4*>(sb + (2 * C + c) * R) + i);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gb + (0 * C + c) * R) + i);
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gb + (1 * C + c) * R) + i);
        const float4 g2 = __ldg(reinterpret_cast<const float4*>(gb + (2 * C + c) * R) + i);
        if (u0.x >= 0.f) acc = fmaf(g0.x, u1.x, fmaf(g1.x, u0.x, fmaf(g2.x, u2.x, acc)));
        if (u0.y >= 0.f) acc = fmaf(g0.y, u1.y, fmaf(g1.y, u0.y, fmaf(g2.y, u2.y, acc)));
        if (u0.z >= 0.f) acc = fmaf(g0.z, u1.z, fmaf(g1.z, u0.z, fmaf(g2.z, u2.z, acc)));
        if (u0.w >= 0.f) acc = fmaf(g0.w, u1.w, fmaf(g1.w, u0.w, fmaf(g2.w, u2.w, acc)));
      }
      acc = warp_sum(acc);
      if (lane == 0) partial[b * C + c] = -acc;
    }
    return;
  }
  for (int c = 0; c < C; ++c) {
    float acc = 0.f;
    for (int r = lane; r < R; r += 32) {
      const float u0 = __ldg(sb + (1 * C + c) * R + r);
      if (u0 >= 0.f) {      // an all-masked vital contributes nothing (its upstream gradients may be NaN)
        acc = fmaf(__ldg(gb + (0 * C + c) * R + r), __ldg(sb + (0 * C + c) * R + r), acc);
        acc = fmaf(__ldg(gb + (1 * C + c) * R + r), u0, acc);
        acc = fmaf(__ldg(gb + (2 * C + c) * R + r), __ldg(sb + (2 * C + c) * R + r), acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) partial[b * C + c] = -acc;
  }
}

// d_kernel[c] = sigmoid(kernel[c]) * sum_b partial[b][c]
__global__ void sigmoid_vec_kernel(const float* __restrict__ kernel, float* __restrict__ out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = sigmoid_ref(kernel[c]);
}

int pick_rpt(int R) { return R <= 32 ? 1 : (R <= 64 ? 2 : 3); }

int pick_warps(int C, int R, int rpt) {
  const int chunks = (R + 32 * rpt - 1) / (32 * rpt);
  int w = (C * chunks + 1) / 2;          // two tasks per warp, paired heavy + light (snake order)
  if (w > kMaxWarps) w = kMaxWarps;
  if (w < 2) w = 2;
  return w;
}

template <typename K>
int prepare(K kern, size_t smem) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
  }
  return DIC_OK;
}

int check_common(const void* x, const void* kernel, const void* ref_t, int64_t B, int C, int T, int R,
                 int64_t& x_stride) {
  // per-batch buffers of an EMPTY batch may be NULL (torch hands out data_ptr() == 0 for empty tensors)
  DIC_REQUIRE((x || B == 0) && kernel && ref_t, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  if (x_stride == 0) x_stride = (int64_t)4 * C * T;
  DIC_REQUIRE(x_stride >= (int64_t)3 * C * T, DIC_ERR_INVALID_ARGUMENT,
              "x_stride=%lld is smaller than the 3*C*T live planes of an encounter", (long long)x_stride);
  DIC_REQUIRE(B >= 0 && C > 0 && T > 0 && R > 0, DIC_ERR_INVALID_ARGUMENT,
              "bad sizes B=%lld C=%d T=%d R=%d", (long long)B, C, T, R);
  DIC_REQUIRE(B <= 2147483647LL, DIC_ERR_UNSUPPORTED, "B=%lld exceeds the grid limit; split the batch",
              (long long)B);
  DIC_REQUIRE(sci_smem_bytes(C, round_up(T, 4), R) <= (size_t)kMaxSmemBytes, DIC_ERR_UNSUPPORTED,
              "C=%d T=%d needs %zu bytes of shared memory per encounter (limit %d)", C, T,
              sci_smem_bytes(C, round_up(T, 4), R), kMaxSmemBytes);
  return DIC_OK;
}

}  // namespace
}  // namespace dic

using namespace dic;

extern "C" int dic_sci_fwd(const float* x, const float* kernel, const float* ref_t, float* u,
                           float* stats, int64_t B, int C, int T, int R, int64_t x_stride,
                           dic_stream_t stream) {
  int rc = check_common(x, kernel, ref_t, B, C, T, R, x_stride);
  if (rc) return rc;
  DIC_REQUIRE(u || B == 0, DIC_ERR_INVALID_ARGUMENT, "null output pointer");
  if (B == 0) return DIC_OK;
  const int Tp = round_up(T, 4);
  const size_t smem = sci_smem_bytes(C, Tp, R);
  const int use_tma = (Tp == T) && aligned16(x) && ((3LL * C * T * 4) % 16 == 0) &&
                      ((x_stride * 4) % 16 == 0);
  const int rpt = pick_rpt(R);
  const int threads = 32 * pick_warps(C, R, rpt);
  cudaStream_t st = as_stream(stream);
#define DIC_SCI_FWD_(RPT_, MOM_)                                                                        \
  {                                                                                                     \
    rc = prepare(sci_fwd_kernel<RPT_, MOM_>, smem);                                                     \
    if (rc) return rc;                                                                                  \
    sci_fwd_kernel<RPT_, MOM_><<<(unsigned)B, threads, smem, st>>>(x, kernel, ref_t, u, stats, C, T, Tp, \
                                                                    R, use_tma, x_stride);              \
  }
#define DIC_SCI_FWD(RPT_) \
  if (stats) DIC_SCI_FWD_(RPT_, true) else DIC_SCI_FWD_(RPT_, false)
  if (rpt == 1) { DIC_SCI_FWD(1) } else if (rpt == 2) { DIC_SCI_FWD(2) } else { DIC_SCI_FWD(3) }
#undef DIC_SCI_FWD_
#undef DIC_SCI_FWD
  DIC_LAUNCH_CHECK("sci_fwd_kernel");
  return DIC_OK;
}

extern "C" size_t dic_interp_bwd_workspace_bytes(int64_t B, int C) {
  if (B < 0 || C <= 0) return 0;
  size_t part = ((size_t)B * C * sizeof(float) + 255) / 256 * 256;
  return part + (size_t)kColsumBlocks * C * sizeof(double) + (size_t)C * sizeof(float) + 256;
}

extern "C" int dic_sci_bwd(const float* x, const float* kernel, const float* ref_t, const float* u,
                           const float* stats, const float* grad_u, float* d_kernel,
                           void* workspace, int64_t B, int C, int T, int R, int64_t x_stride,
                           dic_stream_t stream) {
  // x, ref_t and u are part of the signature for symmetry with dic_sci_fwd; the gradient is a function of the
  // saved moment rows and the upstream gradient alone, so they are not read (and may be NULL)
  (void)x; (void)ref_t; (void)u; (void)x_stride;
  DIC_REQUIRE(kernel && d_kernel, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE(B >= 0 && C > 0 && T > 0 && R > 0, DIC_ERR_INVALID_ARGUMENT, "bad sizes B=%lld C=%d T=%d R=%d",
              (long long)B, C, T, R);
  DIC_REQUIRE((stats && grad_u && workspace) || B == 0, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  cudaStream_t st = as_stream(stream);
  if (B == 0) {
    DIC_CUDA(cudaMemsetAsync(d_kernel, 0, sizeof(float) * C, st));
    return DIC_OK;
  }
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  float* partial = reinterpret_cast<float*>(ws);
  size_t off = ((size_t)B * C * sizeof(float) + 255) / 256 * 256;
  double* red = reinterpret_cast<double*>(ws + off);
  float* sig = reinterpret_cast<float*>(ws + off + (size_t)kColsumBlocks * C * sizeof(double));
  const int vec4 = (R % 4 == 0) && aligned16(stats) && aligned16(grad_u);
  sci_bwd_kernel<<<(unsigned)((B + kSciBwdWarps - 1) / kSciBwdWarps), kSciBwdWarps * 32, 0, st>>>(stats, grad_u, partial,
                                                                                                  B, C, R, vec4);
  DIC_LAUNCH_CHECK("sci_bwd_kernel");
  sigmoid_vec_kernel<<<(C + 127) / 128, 128, 0, st>>>(kernel, sig, C);
  DIC_LAUNCH_CHECK("sigmoid_vec_kernel");
  return colsum_f32_launch(partial, nullptr, d_kernel, sig, red, B, C, st);
}
