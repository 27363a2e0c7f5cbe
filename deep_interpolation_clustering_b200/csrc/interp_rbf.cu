// RBF read-out (grid -> observation times) forward / backward for sm_100a.
//
// Reference: rbf.py:57-108 with the gaussian basis rbf.py:129-131, after compress_fc
// (v = compress_fc output, (B, C, R)); closed-form backward in SURVEY.md Appendix A.3.
//     beta = softplus(kernel[c])
//     phi[t,r] = m_t exp(-beta (d_t - r_r)^2),  N_t = sum_r phi
//     rec_t = m_t * sum_r phi[t,r] v[c,r] / (N_t + 1e-10)
// The reduction runs over the reference grid (the transpose of SCI).  There is no
// softmax shift in the reference (plain exp, absolute epsilon), so none is applied here.
//
// Forward: one CTA per encounter, warp tasks of 32 consecutive observations of one vital; each lane walks a
// uniform-length window of the vital's grid row (s r_j, v_j) in shared memory; one MUFU.EX2 per (t, r).
// Backward: the reduction for grad_v runs over observations, so the encounter is staged
// like SCI (interp_stage.cuh) and each lane owns RPT grid points.
#include "interp_stage.cuh"

namespace dic {
namespace {

constexpr int kMaxWarps = 8;

// Window cut-off: a (t, r) pair whose basis value is below 2^-kRbfCut (6e-8) is skipped.  The
// normaliser N_t is >= ~1 whenever the observation lies on the grid's span; summed over the grid the dropped
// tail is < 2e-8 of N_t and < 5e-7 of sum n phi (the d beta moment) for any kernel width - two orders below
// the 1e-5 parity bound (2^-30 measured the same parity and 5 % more pairs).  Observations more than one hour
// outside the grid span, and non-uniform grids, take the full range.
constexpr float kRbfCut = 24.0f;

// ---- forward, warp-task form -------------------------------------------------------------------------
// The encounter's mask + time planes and its grid values arrive by TMA bulk copies while the parameters are
// computed, and the work is cut into TASKS of 32 consecutive observations of one
// vital: within a task the vital, its window length and its (s r_j, v_j) row are warp-uniform, every lane
// reads its observation from shared memory and the 32 results leave as one coalesced store.  Tasks are dealt
// round-robin to the warps (about 27 tasks for 4 warps at c2).
constexpr int kRbfFwd2Warps = 4;

struct RbfFwd2Smem {
  uint64_t* bar;
  float* planes;    // [2][C][Tp]: mask | time
  float* sv;        // [C][R] grid values
  float2* srv;      // [C][Rp] (-r_j, v_cj), pad (-huge, 0)
  float* ssc;       // [C] s_c = sqrt(beta_c log2 e)
  int* strip;       // [C] grid points per window (even)
  float* swin;      // [C] window half-width in hours, sqrt(kRbfCut / (beta log2 e))
  int* nval;        // [C] observation slots to visit: the prefix length, or T for a general mask
  int* tbase;       // [C + 1] prefix sums of the task counts
};
__host__ __device__ inline size_t rbf_fwd2_offsets(int C, int Tp, int R, int Rp, size_t (&off)[8]) {
  size_t o = 16;
  off[0] = o; o += sizeof(float) * 2 * (size_t)C * Tp;
  off[1] = o; o += (sizeof(float) * (size_t)C * R + 15) / 16 * 16;
  off[2] = o; o += sizeof(float2) * (size_t)C * Rp;
  off[3] = o; o += sizeof(float) * C;
  off[4] = o; o += sizeof(int) * C;
  off[5] = o; o += sizeof(int) * C;
  off[6] = o; o += sizeof(int) * (C + 1);
  off[7] = o; o += sizeof(float) * C;
  return (o + 15) / 16 * 16;
}
__device__ __forceinline__ RbfFwd2Smem rbf_fwd2_carve(unsigned char* base, int C, int Tp, int R, int Rp) {
  size_t off[8];
  rbf_fwd2_offsets(C, Tp, R, Rp, off);
  RbfFwd2Smem s;
  s.bar = reinterpret_cast<uint64_t*>(base);
  s.planes = reinterpret_cast<float*>(base + off[0]);
  s.sv = reinterpret_cast<float*>(base + off[1]);
  s.srv = reinterpret_cast<float2*>(base + off[2]);
  s.ssc = reinterpret_cast<float*>(base + off[3]);
  s.strip = reinterpret_cast<int*>(base + off[4]);
  s.nval = reinterpret_cast<int*>(base + off[5]);
  s.tbase = reinterpret_cast<int*>(base + off[6]);
  s.swin = reinterpret_cast<float*>(base + off[7]);
  return s;
}

__global__ void __launch_bounds__(kRbfFwd2Warps * 32)
rbf_fwd2_kernel(const float* __restrict__ v, const float* __restrict__ x, const float* __restrict__ kernel,
                const float* __restrict__ ref_t, float* __restrict__ rec, float* __restrict__ inv_norm, int C,
                int T, int Tp, int R, int Rp, int64_t x_stride, int use_tma) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const RbfFwd2Smem s = rbf_fwd2_carve(smem_raw, C, Tp, R, Rp);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b = blockIdx.x;
  const float* mb = x + b * x_stride + (int64_t)C * T;         // mask plane, followed by the time plane
  const float* vb = v + b * (int64_t)C * R;
  float* rb = rec + b * (int64_t)C * T;
  float* nb = inv_norm ? inv_norm + b * (int64_t)C * T : nullptr;

  if (use_tma && tid == 0) {                                   // both copies are in flight during the set-up below
    mbar_init(s.bar, 1);
    fence_proxy_async();
    const uint32_t pb = (uint32_t)(2 * C * T) * 4u, vbytes = (uint32_t)(C * R) * 4u;
    mbar_expect_tx(s.bar, pb + vbytes);
    bulk_g2s(s.planes, mb, pb, s.bar);
    bulk_g2s(s.sv, vb, vbytes, s.bar);
  }
  const float r0 = __ldg(ref_t), rl = __ldg(ref_t + R - 1);
  const float h = R > 1 ? (rl - r0) / (float)(R - 1) : 1.0f;
  int irregular = !(h > 0.f);
  for (int c = tid; c < C; c += blockDim.x) {
    const float b2 = softplus_ref(__ldg(kernel + c)) * kLog2e;
    s.ssc[c] = sqrtf(b2);
    // the grid points within +-w of an observation: at most floor(2 w / h) + 1, plus one for the even start
    const float w = sqrtf(kRbfCut / b2);
    s.swin[c] = w;
    s.strip[c] = ((int)floorf(2.f * w / h + 0.02f) + 3) & ~1;
  }
  for (int j = tid; j < R; j += blockDim.x) irregular |= fabsf(__ldg(ref_t + j) - (r0 + h * (float)j)) > 0.01f * h;
  if (!use_tma) {
    for (int i = tid; i < 2 * C * Tp; i += blockDim.x) {
      const int row = i / Tp, t = i - row * Tp;
      s.planes[i] = t < T ? __ldg(mb + row * T + t) : 0.f;
    }
    for (int i = tid; i < C * R; i += blockDim.x) s.sv[i] = __ldg(vb + i);
  }
  irregular = __syncthreads_or(irregular);                     // also publishes ssc / strip / the mbarrier
  if (use_tma) mbar_wait(s.bar, 0);

  const float* smask = s.planes;
  const float* stime = s.planes + C * Tp;
  // rows of (-r_j, -r_j+1, v_cj, v_cj+1) per pair of grid points: two aligned register pairs per 128-bit
  // load for the packed float32x2 loop below; a pad point sits at -1e12 with value 0 => weight exactly 0.
  // The coordinates stay UNSCALED: d - r is formed first (its rounding error is relative to the small
  // difference) and then scaled by s_c.  Pre-scaled coordinates s d - s r carry the rounding of numbers up to
  // s H, which a narrow kernel (beta = 8: s H = 81) turns into 1e-5 relative errors of the basis values.
  for (int i = tid; i < C * (Rp / 2); i += blockDim.x) {
    const int c = i / (Rp / 2), j = 2 * (i - c * (Rp / 2));
    const float ra = -__ldg(ref_t + j), rb2 = j + 1 < R ? -__ldg(ref_t + j + 1) : -1.0e12f;
    reinterpret_cast<float4*>(s.srv)[i] = make_float4(ra, rb2, s.sv[c * R + j], j + 1 < R ? s.sv[c * R + j + 1] : 0.f);
  }
  for (int c = warp; c < C; c += kRbfFwd2Warps) {              // prefix masks: visit [0, n) and zero the tail
    const float* mrow = smask + c * Tp;
    int ok = 1, cnt = 0;
    for (int t = lane; t < T; t += 32) {
      const float m = mrow[t];
      const float mn = t + 1 < T ? mrow[t + 1] : 0.f;
      ok &= ((m == 0.f) | (m == 1.f)) & (m >= mn);
      cnt += (m == 1.f);
    }
    ok = __all_sync(0xffffffffu, ok);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    const int n = ok ? cnt : T;                                // general mask: every slot is visited
    for (int t = n + lane; t < T; t += 32) {
      rb[c * T + t] = 0.f;
      if (nb) nb[c * T + t] = 0.f;
    }
    if (lane == 0) s.nval[c] = n;
  }
  __syncthreads();
  if (tid == 0) {
    s.tbase[0] = 0;
    for (int c = 0; c < C; ++c) s.tbase[c + 1] = s.tbase[c] + (s.nval[c] + 31) / 32;
  }
  __syncthreads();

  const float inv_h = 1.0f / h;
  const int ntask = s.tbase[C];
  int c = 0;
  for (int k = warp; k < ntask; k += kRbfFwd2Warps) {
    while (k >= s.tbase[c + 1]) ++c;                           // warp-uniform, monotone in k
    const int t = (k - s.tbase[c]) * 32 + lane;
    const int n = s.nval[c];
    const bool live = t < n;
    const int tc = live ? t : n - 1;                           // idle lanes shadow the last observation
    const float m = smask[c * Tp + tc];
    const float d = stime[c * Tp + tc];
    const float sc = s.ssc[c];
    // every lane walks the same number of grid points from its own even offset
    int trip = min(Rp, s.strip[c]);
    const bool wide = irregular || !(d >= r0 - 1.0f && d <= rl + 1.0f);
    if (__any_sync(0xffffffffu, wide && live && m != 0.f)) trip = Rp;      // rare: the whole row for this task
    const int jlo = trip == Rp ? 0
                               : min(max(0, __float2int_ru((d - s.swin[c] - r0) * inv_h - 0.01f)) & ~1, Rp - trip);
    f2_t N2 = pack2(0.f, 0.f), S2 = N2;
    const f2_t d2 = pack2(d, d), sc2 = pack2(sc, sc);
    const ulonglong2* rw = reinterpret_cast<const ulonglong2*>(s.srv) + c * (Rp / 2) + (jlo >> 1);
#pragma unroll 4
    for (int j = 0; j < (trip >> 1); ++j) {
      const ulonglong2 p = rw[j];                                  // (-r0, -r1) | (v0, v1)
      const f2_t dl = mul2(add2(d2, p.x), sc2);                    // s (d - r)
      const f2_t n2 = mul2(dl, dl);
      float n0, n1;
      unpack2(n2, n0, n1);
      const f2_t e = pack2(ex2_approx(-n0), ex2_approx(-n1));
      N2 = add2(N2, e);
      S2 = fma2(e, p.y, S2);
    }
    const float N = sum2(N2), S = sum2(S2);
    if (live) {
      // phi = m e  =>  N_ref = m N, sum phi v = m S; a masked slot of a general row reconstructs to 0
      const float inv = m != 0.f ? __frcp_rn(fmaf(m, N, 1e-10f)) : 0.f;
      rb[c * T + t] = (m * S) * inv * m;                        // rbf.py:106-107
      if (nb) nb[c * T + t] = inv;
    }
  }
}

struct RbfSmem {
  uint64_t* bar;
  float* rows;     // [3][C][Tp]: mask, then a' = m^2 g invN | s d (scaled time) | a'S = m g invN rec
  int* n_valid;    // [C]
  int* order;      // [C] vitals by descending count
  float* part;     // [C * ceil(R/32)]
};

__device__ __forceinline__ RbfSmem rbf_carve(unsigned char* base, int C, int Tp) {
  RbfSmem s;
  s.bar = reinterpret_cast<uint64_t*>(base);
  s.rows = reinterpret_cast<float*>(base + 16);
  s.n_valid = reinterpret_cast<int*>(s.rows + 3 * C * Tp);
  s.order = s.n_valid + C;
  s.part = reinterpret_cast<float*>(s.order + C);
  return s;
}

static size_t rbf_bwd_smem_bytes(int C, int Tp, int R) {
  return 16 + sizeof(float) * (3 * (size_t)C * Tp) + 2 * sizeof(int) * C +
         sizeof(float) * (size_t)C * ((R + 31) / 32);
}

// grad_v[c,r] = sum_t a'_t e_tr,  d beta_c = -sum_{t,r} n_tr e_tr (a'_t v_r - a'S_t)
// with e_tr = exp(-beta n_tr), a'_t = m_t^2 g_t invN_t, a'S_t = m_t g_t invN_t rec_t.
// Same structure as the SCI kernels: rows sorted by time, each lane owns RPT adjacent grid
// points and walks only the observations within +-sqrt(kRbfCut / (beta log2 e)) hours of them.
template <int RPT>
__global__ void __launch_bounds__(kMaxWarps * 32)
rbf_bwd_kernel(const float* __restrict__ v, const float* __restrict__ x,
               const float* __restrict__ kernel, const float* __restrict__ ref_t,
               const float* __restrict__ rec, const float* __restrict__ inv_norm,
               const float* __restrict__ grad_rec, float* __restrict__ grad_v,
               float* __restrict__ partial, int C, int T, int Tp, int R, int64_t x_stride, int use_tma) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const RbfSmem s = rbf_carve(smem_raw, C, Tp);
  const int64_t b = blockIdx.x;
  const float* mb = x + b * x_stride + (int64_t)C * T;
  const float* rb = rec + b * (int64_t)C * T;
  const float* nb = inv_norm + b * (int64_t)C * T;
  const float* gb = grad_rec + b * (int64_t)C * T;
  float* sa = s.rows;                 // mask first, then a'
  float* sd = s.rows + C * Tp;        // times (unscaled: the loop scales the difference d - r, see rbf_fwd2_kernel)
  float* sas = s.rows + 2 * C * Tp;   // a'S
  // the mask and time planes of one encounter are adjacent: ONE bulk copy of 2*C*T floats
  stage_rows(s.rows, mb, 2 * C, T, Tp, s.bar, use_tma != 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int c = warp; c < C; c += nwarps) {
    float* rd = sd + c * Tp;
    float* ra = sa + c * Tp;
    float* ras = sas + c * Tp;
    int n = warp_canonical_count(ra, rd, Tp, lane);
    __syncwarp();
    for (int t = lane; t < T; t += 32) {        // mask -> a', a'S (each lane rewrites its own slots)
      const float m = ra[t];
      const float gi = __ldg(gb + c * T + t) * __ldg(nb + c * T + t) * m;
      ra[t] = gi * m;
      ras[t] = -gi * __ldg(rb + c * T + t);              // stored NEGATED: the loop forms a' v - a'S with one FMA
    }
    __syncwarp();
    if (n < 0) {   // general rows: keep entries that carry gradient (a' != 0), sort by time
      n = warp_compact3(ras, ra, rd, T, lane, false);
      warp_sort3(rd, ra, ras, n, lane);
    }
    warp_pad4_far(rd, ra, ras, n, lane);
    // pad entries at 1e12 h instead of kPadTime: (s 1e12)^2 stays finite for any kernel the float32 softplus can
    // produce, so the moment term e * n is 0 * finite, never 0 * inf
    if (lane < ((n + 3) & ~3) - n) rd[n + lane] = 1.0e12f;
    __syncwarp();
    if (lane == 0) s.n_valid[c] = n;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int c = 0; c < C; ++c) {
      const int key = s.n_valid[c];
      int j = c;
      while (j > 0 && s.n_valid[s.order[j - 1]] < key) {
        s.order[j] = s.order[j - 1];
        --j;
      }
      s.order[j] = c;
    }
  }
  __syncthreads();

  const int chunks = (R + 32 * RPT - 1) / (32 * RPT);
  const int ntasks = C * chunks;
  const float* vb = v + b * (int64_t)C * R;
  float* gvb = grad_v + b * (int64_t)C * R;
  for (int round = 0;; ++round) {
    if (round * nwarps >= ntasks) break;
    const int task = round * nwarps + ((round & 1) ? (nwarps - 1 - warp) : warp);
    if (task >= ntasks) continue;
    const int c = s.order[task / chunks], chunk = task % chunks;
    const float* rd = sd + c * Tp;
    const float* ra = sa + c * Tp;
    const float* ras = sas + c * Tp;
    const int n = s.n_valid[c];
    const float b2 = softplus_ref(__ldg(kernel + c)) * kLog2e;
    const float sc = sqrtf(b2);
    int ridx[RPT];
    float rr[RPT], vv[RPT], dv[RPT], acc[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      ridx[k] = (chunk * 32 + lane) * RPT + k;
      const int rc = min(ridx[k], R - 1);
      rr[k] = __ldg(ref_t + rc);
      vv[k] = __ldg(vb + c * R + rc);
      dv[k] = acc[k] = 0.f;
    }
    // one window [r_first - w, r_last + w] (w = sqrt(cut / b2) hours), two searches, uniform trip
    const int n4 = (n + 3) & ~3;
    int wbase = 0, wtrip = n4;
    if (n > 0) {
      const float wcut = sqrtf(kRbfCut / b2);
      const float tv[2] = {rr[0] - wcut, rr[RPT - 1] + wcut};
      const bool up[2] = {false, true};
      int pos[2];
      multi_bound<2>(rd, n, tv, up, pos);
      wbase = pos[0] & ~3;
      wtrip = (warp_max_i(pos[1] - wbase) + 3) & ~3;
    }
    // packed float32x2 over pairs of consecutive observations (two aligned register pairs per 128-bit load)
    f2_t nrr2[RPT], vv2[RPT], dv2[RPT], acc2[RPT];
    const f2_t sc2 = pack2(sc, sc);
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      nrr2[k] = pack2(-rr[k], -rr[k]);
      vv2[k] = pack2(vv[k], vv[k]);
      dv2[k] = acc2[k] = pack2(0.f, 0.f);
    }
    for (int t0 = 0; t0 < wtrip; t0 += 4) {             // (unrolling by 2 measured slower: 14.7 -> 16.4 ms)
      const int t = wbase + t0;
      if ((unsigned)t >= (unsigned)n4) continue;        // chunk off the row (lane at an end of the record)
      const ulonglong2 d4 = *reinterpret_cast<const ulonglong2*>(rd + t);
      const ulonglong2 a4 = *reinterpret_cast<const ulonglong2*>(ra + t);
      const ulonglong2 s4 = *reinterpret_cast<const ulonglong2*>(ras + t);      // -a'S
      const f2_t dd[2] = {d4.x, d4.y}, aa[2] = {a4.x, a4.y}, ns[2] = {s4.x, s4.y};
#pragma unroll
      for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          const f2_t dl = mul2(add2(dd[j], nrr2[k]), sc2);       // s (d - r): the difference first, then the scale
          const f2_t n2 = mul2(dl, dl);                          // beta log2(e) (d - r)^2
          float n0, n1;
          unpack2(n2, n0, n1);
          const f2_t e = pack2(ex2_approx(-n0), ex2_approx(-n1));
          dv2[k] = fma2(e, aa[j], dv2[k]);
          const f2_t tt = fma2(aa[j], vv2[k], ns[j]);
          acc2[k] = fma2(mul2(e, n2), tt, acc2[k]);              // sum of +b2 n e (...): rescaled below
        }
      }
    }
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      dv[k] = sum2(dv2[k]);
      acc[k] = sum2(acc2[k]);
    }
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      if (ridx[k] < R) {
        gvb[c * R + ridx[k]] = dv[k];
        tot += acc[k];
      }
    }
    tot = warp_sum(tot) * (1.0f / b2);                         // back to sum n e (a' v - a'S)
    if (lane == 0) s.part[c * chunks + chunk] = tot;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
    for (int k = 0; k < chunks; ++k) t += s.part[c * chunks + k];
    partial[b * C + c] = -t;
  }
}

__global__ void sigmoid_vec_kernel(const float* __restrict__ kernel, float* __restrict__ out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = sigmoid_ref(kernel[c]);
}

int check(const void* v, const void* x, const void* kernel, const void* ref_t, int64_t B, int C,
          int T, int R, int64_t& x_stride) {
  DIC_REQUIRE(((v && x) || B == 0) && kernel && ref_t, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  if (x_stride == 0) x_stride = (int64_t)4 * C * T;
  DIC_REQUIRE(x_stride >= (int64_t)3 * C * T, DIC_ERR_INVALID_ARGUMENT,
              "x_stride=%lld is smaller than the 3*C*T live planes of an encounter", (long long)x_stride);
  DIC_REQUIRE(B >= 0 && C > 0 && T > 0 && R > 0, DIC_ERR_INVALID_ARGUMENT,
              "bad sizes B=%lld C=%d T=%d R=%d", (long long)B, C, T, R);
  DIC_REQUIRE(B <= 2147483647LL, DIC_ERR_UNSUPPORTED, "B=%lld exceeds the grid limit", (long long)B);
  return DIC_OK;
}

}  // namespace
}  // namespace dic

using namespace dic;

extern "C" int dic_rbf_fwd(const float* v, const float* x, const float* kernel, const float* ref_t,
                           float* rec, float* inv_norm, int64_t B, int C, int T, int R,
                           int64_t x_stride, dic_stream_t stream) {
  int rc = check(v, x, kernel, ref_t, B, C, T, R, x_stride);
  if (rc) return rc;
  DIC_REQUIRE(rec || B == 0, DIC_ERR_INVALID_ARGUMENT, "null output pointer");
  if (B == 0) return DIC_OK;
  const int Rp = round_up(R, 2);
  const int Tp = round_up(T, 4);
  size_t off[8];
  const size_t smem2 = rbf_fwd2_offsets(C, Tp, R, Rp, off);
  DIC_REQUIRE(smem2 <= (size_t)kMaxSmemBytes, DIC_ERR_UNSUPPORTED,
              "C=%d T=%d R=%d needs %zu bytes of shared memory per encounter (limit %d)", C, T, R, smem2, kMaxSmemBytes);
  const int use_tma = (Tp == T) && aligned16(x) && aligned16(v) && ((x_stride * 4) % 16 == 0) &&
                      (((int64_t)C * T * 4) % 16 == 0) && (((int64_t)C * R * 4) % 16 == 0);
  if (smem2 > 48 * 1024)
    DIC_CUDA(cudaFuncSetAttribute(rbf_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  rbf_fwd2_kernel<<<(unsigned)B, kRbfFwd2Warps * 32, smem2, as_stream(stream)>>>(
      v, x, kernel, ref_t, rec, inv_norm, C, T, Tp, R, Rp, x_stride, use_tma);
  DIC_LAUNCH_CHECK("rbf_fwd2_kernel");
  return DIC_OK;
}

extern "C" int dic_rbf_bwd(const float* v, const float* x, const float* kernel, const float* ref_t,
                           const float* rec, const float* inv_norm, const float* grad_rec,
                           float* grad_v, float* d_kernel, void* workspace, int64_t B, int C, int T,
                           int R, int64_t x_stride, dic_stream_t stream) {
  int rc = check(v, x, kernel, ref_t, B, C, T, R, x_stride);
  if (rc) return rc;
  DIC_REQUIRE(((rec && inv_norm && grad_rec && grad_v && workspace) || B == 0) && d_kernel,
              DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  cudaStream_t st = as_stream(stream);
  if (B == 0) {
    DIC_CUDA(cudaMemsetAsync(d_kernel, 0, sizeof(float) * C, st));
    return DIC_OK;
  }
  const int Tp = round_up(T, 4);
  const size_t smem = rbf_bwd_smem_bytes(C, Tp, R);
  DIC_REQUIRE(smem <= (size_t)kMaxSmemBytes, DIC_ERR_UNSUPPORTED,
              "C=%d T=%d needs %zu bytes of shared memory per encounter (limit %d)", C, T, smem,
              kMaxSmemBytes);
  const int use_tma = (Tp == T) && aligned16(x) && ((x_stride * 4) % 16 == 0) && (((int64_t)C * T * 4) % 16 == 0);
  const int rpt = R <= 32 ? 1 : (R <= 64 ? 2 : 3);
  const int chunks = (R + 32 * rpt - 1) / (32 * rpt);
  int warps = (C * chunks + 1) / 2;      // two tasks per warp, paired heavy + light (snake order)
  warps = warps > kMaxWarps ? kMaxWarps : (warps < 2 ? 2 : warps);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  float* partial = reinterpret_cast<float*>(ws);
  size_t off = ((size_t)B * C * sizeof(float) + 255) / 256 * 256;
  double* red = reinterpret_cast<double*>(ws + off);
  float* sig = reinterpret_cast<float*>(ws + off + (size_t)kColsumBlocks * C * sizeof(double));
#define DIC_RBF_BWD(RPT_)                                                                        \
  {                                                                                              \
    if (smem > 48 * 1024)                                                                        \
      DIC_CUDA(cudaFuncSetAttribute(rbf_bwd_kernel<RPT_>,                                        \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    rbf_bwd_kernel<RPT_><<<(unsigned)B, warps * 32, smem, st>>>(                                 \
        v, x, kernel, ref_t, rec, inv_norm, grad_rec, grad_v, partial, C, T, Tp, R, x_stride,    \
        use_tma);                                                                                \
  }
  if (rpt == 1) DIC_RBF_BWD(1) else if (rpt == 2) DIC_RBF_BWD(2) else DIC_RBF_BWD(3)
#undef DIC_RBF_BWD
  DIC_LAUNCH_CHECK("rbf_bwd_kernel");
  sigmoid_vec_kernel<<<(C + 127) / 128, 128, 0, st>>>(kernel, sig, C);
  DIC_LAUNCH_CHECK("sigmoid_vec_kernel");
  return colsum_f32_launch(partial, nullptr, d_kernel, sig, red, B, C, st);
}
