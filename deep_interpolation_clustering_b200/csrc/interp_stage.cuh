// Per-encounter staging shared by the SCI forward/backward and RBF backward kernels.
//
// One CTA owns one encounter.  Its observation rows are brought into shared memory
// (one 1-D TMA bulk copy when the span is 16-byte friendly), then every channel is
// CANONICALISED by one warp: observations with a zero mask are dropped (stable
// compaction), the rest are sorted by time (odd-even transposition with early exit:
// free for the left-packed, time-ordered rows the reference pipeline produces,
// p0_data_process.py:44-67; correct for the jittered or shuffled rows of
// dataloader.py:207-208), and the tail is padded to a multiple of four with
// zero-weight entries so the inner loops can use 128-bit shared loads.
#pragma once

#include "common.cuh"

namespace dic {

// Stable in-place compaction of three parallel rows by `keep(m)`; returns the count.
// r0 <- f0(r0, m) lets the caller fold the mask into the value row on the fly.
// Executed by ONE warp.
__device__ __forceinline__ int warp_compact3(float* r0, float* rm, float* r2, int T, int lane,
                                             bool fold_mask_into_r0) {
  int n = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    float a = 0.f, m = 0.f, d = 0.f;
    if (t < T) {
      a = r0[t];
      m = rm[t];
      d = r2[t];
    }
    const bool keep = (t < T) && (m != 0.f);
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int pos = n + __popc(bal & ((1u << lane) - 1u));
    __syncwarp();
    if (keep) {
      r0[pos] = fold_mask_into_r0 ? a * m : a;
      rm[pos] = m;
      r2[pos] = d;
    }
    n += __popc(bal);
    __syncwarp();
  }
  return n;
}

// Odd-even transposition sort of rows (key, a, b) by key, ascending; one warp.
// Terminates after the first full pass without a swap.
__device__ __forceinline__ void warp_sort3(float* key, float* a, float* b, int n, int lane) {
  if (n < 2) return;
  for (int pass = 0; pass < n; ++pass) {
    int swapped = 0;
#pragma unroll
    for (int phase = 0; phase < 2; ++phase) {
      for (int i = phase + 2 * lane; i + 1 < n; i += 64) {
        const float k0 = key[i], k1 = key[i + 1];
        if (k0 > k1) {
          key[i] = k1;
          key[i + 1] = k0;
          float t = a[i];
          a[i] = a[i + 1];
          a[i + 1] = t;
          t = b[i];
          b[i] = b[i + 1];
          b[i + 1] = t;
          swapped = 1;
        }
      }
      __syncwarp();
    }
    if (!__any_sync(0xffffffffu, swapped)) break;
  }
}

// Pads rows to a multiple of 4 entries with zero-weight copies of the last key.
__device__ __forceinline__ void warp_pad4(float* key, float* a, float* b, int n, int lane) {
  const int n4 = (n + 3) & ~3;
  if (lane < n4 - n) {
    key[n + lane] = n > 0 ? key[n - 1] : 0.f;
    a[n + lane] = 0.f;
    b[n + lane] = 0.f;
  }
  __syncwarp();
}

// Fast path of the canonicalisation.  Returns the observation count if the row already is what
// the reference pipeline produces - a 0/1 mask that is a left-packed prefix and non-decreasing
// times on that prefix (p0_data_process.py:44-67) - and -1 otherwise.  One vectorised pass,
// nothing is moved.  Rows hold Tp (multiple of 4) entries with zero mask beyond T.  One warp.
__device__ __forceinline__ int warp_canonical_count(const float* rm, const float* rd, int Tp, int lane) {
  int ok = 1, cnt = 0;
  for (int t = 4 * lane; t < Tp; t += 128) {
    const float4 m = *reinterpret_cast<const float4*>(rm + t);
    const float4 d = *reinterpret_cast<const float4*>(rd + t);
    const bool last = t + 4 >= Tp;
    const float mn = last ? 0.f : rm[t + 4];
    const float dn = last ? 0.f : rd[t + 4];
    const float mm[5] = {m.x, m.y, m.z, m.w, mn};
    const float dd[5] = {d.x, d.y, d.z, d.w, dn};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ok &= (mm[i] == 0.f) | (mm[i] == 1.f);
      ok &= (mm[i] >= mm[i + 1]);                                   // prefix: never 0 -> 1
      ok &= (mm[i + 1] == 0.f) | (dd[i] <= dd[i + 1]);              // sorted where both valid
      cnt += (mm[i] == 1.f);
    }
  }
  ok = __all_sync(0xffffffffu, ok);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  return ok ? cnt : -1;
}

constexpr float kPadTime = 1.0e12f;   // a (kPadTime - d*)^2 stays finite for any kernel (no 0 * inf in the moments) and 2^(-that) == 0

// Pads rows [n, round_up(n,4)) with entries that contribute exactly nothing.
__device__ __forceinline__ void warp_pad4_far(float* key, float* a, float* b, int n, int lane) {
  const int n4 = (n + 3) & ~3;
  if (lane < n4 - n) {
    key[n + lane] = kPadTime;
    a[n + lane] = 0.f;
    b[n + lane] = 0.f;
  }
  __syncwarp();
}

// First index in sorted key[0..n) with key[i] >= v (lower) / key[i] > v (upper).
__device__ __forceinline__ int lower_bound_sorted(const float* key, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (key[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int upper_bound_sorted(const float* key, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (key[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Branch-free binary searches for NV targets at once over sorted key[0..n), n >= 1:
//   pos[k] = #{ i : key[i] < v[k] }   (lower bound)      when !upper[k]
//   pos[k] = #{ i : key[i] <= v[k] }  (upper bound)      when  upper[k]
// The trip count depends only on n (warp-uniform) and the NV probes of a step are independent,
// so their shared-memory latencies overlap.
template <int NV>
__device__ __forceinline__ void multi_bound(const float* key, int n, const float (&v)[NV],
                                            const bool (&upper)[NV], int (&pos)[NV]) {
#pragma unroll
  for (int k = 0; k < NV; ++k) pos[k] = 0;
  for (int step = 1 << (31 - __clz(n)); step > 0; step >>= 1) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int p = pos[k] + step;
      const float kv = key[min(p, n) - 1];
      const bool go = (p <= n) && (upper[k] ? (kv <= v[k]) : (kv < v[k]));
      pos[k] = go ? p : pos[k];
    }
  }
}

// Per-lane sliding window over the sorted observations of one vital.
//   A lane owns RPT adjacent grid points [r_first, r_last]; only observations within +-w_out of
//   them matter (weights below 2^-cut are dropped), and only those within +-w_in feed the
//   high-pass sums.  Every lane of the warp walks the SAME number of 4-entry chunks (`trip`)
//   from its own 4-aligned base, and the inner window sits at the SAME iteration range
//   [in0, in1) for every lane, so there is neither divergence nor a union-of-windows penalty.
//   A lane near the ends of the record gets a base before the row (or runs past it): those
//   chunks are redirected to a null chunk (time = kPadTime => weight exactly 0).
struct Window {
  int base, trip, in0, in1;
};
__device__ __forceinline__ Window make_window(const float* key, int n, float r_first, float r_last, float w_out,
                                              float w_in, bool full) {
  const int n4 = (n + 3) & ~3;
  Window w;
  if (full || n == 0) {
    w.base = 0; w.trip = n4; w.in0 = 0; w.in1 = n4;
    return w;
  }
  const float tv[4] = {r_first - w_out, r_first - w_in, r_last + w_in, r_last + w_out};
  const bool up[4] = {false, false, true, true};
  int pos[4];
  multi_bound<4>(key, n, tv, up, pos);                 // lo <= ilo <= ihi <= hi
  const int ilo4 = pos[1] & ~3;
  w.in0 = (warp_max_i(pos[1] - pos[0]) + 3) & ~3;      // entries between outer and inner start
  const int lin = (warp_max_i(pos[2] - ilo4) + 3) & ~3;
  const int tail = (warp_max_i(pos[3] - pos[2]) + 3) & ~3;
  w.in1 = w.in0 + lin;
  w.trip = w.in1 + tail;
  w.base = ilo4 - w.in0;
  return w;
}

// Two independent windows for the two filters (interp_sci.cu): the low-pass sums walk [ob, ob + ot) and the
// high-pass sums walk [ib, ib + it) in a second, much shorter loop.  Each window costs ONE warp maximum
// (uniform trip from per-lane 4-aligned bases); with a shared loop the three segments outer-left | inner |
// outer-right each paid their own maximum over the lanes plus their own round-up to whole chunks.
struct Window2 {
  int ob, ot, ib, it;
};
__device__ __forceinline__ Window2 make_window2(const float* key, int n, float r_first, float r_last, float w_out,
                                                float w_in, bool full) {
  const int n4 = (n + 3) & ~3;
  Window2 w;
  if (full || n == 0) {
    w.ob = 0; w.ot = n4; w.ib = 0; w.it = n4;
    return w;
  }
  const float tv[4] = {r_first - w_out, r_first - w_in, r_last + w_in, r_last + w_out};
  const bool up[4] = {false, false, true, true};
  int pos[4];
  multi_bound<4>(key, n, tv, up, pos);                 // lo <= ilo <= ihi <= hi
  w.ob = pos[0] & ~3;
  w.ib = pos[1] & ~3;
  w.ot = (warp_max_i(pos[3] - w.ob) + 3) & ~3;
  w.it = (warp_max_i(pos[2] - w.ib) + 3) & ~3;
  return w;
}

// d* - r for the observation nearest to r, given lb = #{i : key[i] < r} in sorted key[0..n), n >= 1.
__device__ __forceinline__ float nearest_delta(const float* key, int n, int lb, float r) {
  const float below = key[max(lb, 1) - 1] - r;        // <= 0 when lb >= 1
  const float above = key[min(lb, n - 1)] - r;        // >= 0 when lb <  n
  if (lb == 0) return above;
  if (lb == n) return below;
  return (-below <= above) ? below : above;
}

// Index of the observation nearest to r, given lb = #{i : key[i] < r} in sorted key[0..n), n >= 1 (ties: the earlier).
__device__ __forceinline__ int nearest_index(const float* key, int n, int lb, float r) {
  const int ib = max(lb, 1) - 1, ia = min(lb, n - 1);
  if (lb == 0) return ia;
  if (lb == n) return ib;
  return (r - key[ib] <= key[ia] - r) ? ib : ia;
}

// Index of the key nearest to r in sorted key[0..n), n >= 1.
__device__ __forceinline__ int nearest_sorted(const float* key, int n, float r) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (key[mid] < r) lo = mid + 1; else hi = mid;
  }
  if (lo == n) return n - 1;
  if (lo == 0) return 0;
  return (r - key[lo - 1] <= key[lo] - r) ? lo - 1 : lo;
}

// Brings `nrows` consecutive global rows of length T (row stride T) into shared rows of
// stride Tp.  Uses one TMA bulk copy when Tp == T and the span is 16-byte aligned/sized;
// otherwise plain coalesced loads.  Must be called by all threads; ends with the data
// visible to the whole CTA.
__device__ __forceinline__ void stage_rows(float* smem_rows, const float* gsrc, int nrows, int T,
                                           int Tp, uint64_t* bar, bool use_tma) {
  if (use_tma) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      fence_proxy_async();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t bytes = static_cast<uint32_t>(nrows) * T * 4u;
      mbar_expect_tx(bar, bytes);
      bulk_g2s(smem_rows, gsrc, bytes, bar);
    }
    mbar_wait(bar, 0);
  } else {
    const int total = nrows * Tp;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int row = i / Tp, t = i - row * Tp;
      smem_rows[i] = t < T ? __ldg(gsrc + row * T + t) : 0.f;     // pad columns: mask 0
    }
    __syncthreads();
  }
}

// The same in two halves, so that per-encounter set-up work can run while the bulk copy is in flight:
// stage_rows_issue (all threads; returns at once on the TMA path) ... independent work ... stage_rows_wait.
__device__ __forceinline__ void stage_rows_issue(float* smem_rows, const float* gsrc, int nrows, int T, int Tp,
                                                 uint64_t* bar, bool use_tma) {
  if (use_tma) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      fence_proxy_async();
      const uint32_t bytes = static_cast<uint32_t>(nrows) * T * 4u;
      mbar_expect_tx(bar, bytes);
      bulk_g2s(smem_rows, gsrc, bytes, bar);
    }
  } else {
    const int total = nrows * Tp;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int row = i / Tp, t = i - row * Tp;
      smem_rows[i] = t < T ? __ldg(gsrc + row * T + t) : 0.f;     // pad columns: mask 0
    }
  }
}
__device__ __forceinline__ void stage_rows_wait(uint64_t* bar, bool use_tma) {
  __syncthreads();                 // publishes the mbarrier's initialisation / the plain loads
  if (use_tma) mbar_wait(bar, 0);
}

}  // namespace dic
