// One Lloyd pass with the E-step on the tensor cores (tcgen05 + TMEM), fed by 2-D tensor-map TMA: float32 rows of
// 64 / 128 / 256 elements and float64 rows of 64, K <= 16.  sklearn/cluster/_k_means_lloyd.pyx:23-218 (E-step: argmin_j ||c_j||^2 - 2 x.c_j,
// lowest index on ties; accumulation of the per-cluster sums and counts for the M-step).
//
// Why: the CUDA-core kernels (kmeans.cu) pay one shared-memory broadcast operand per FMA - K*D operands per row - and
// sit at 0.1-0.4 of the HBM roofline (ncu: bound by their instruction count).  D x K >= 64 x 8 is a genuine dense
// contraction (north_star), so the x.c products move to tcgen05.mma and the CUDA cores keep only what is not a
// contraction: the low-order operand halves, the arg-min and the per-cluster accumulation.
//
// Version 2 (round 2).  Version 1 converted every tile to split-tf32 operand tiles on the CUDA cores (8 instructions per
// element) into ONE operand buffer and issued the MMAs from a single thread: conversion, MMA issue and accumulation ran
// in series, 0.145 ms per pass at 1M x 64 against 0.097-0.148 ms of the CUDA-core kernel.  Here
//   * the rows arrive by cp.async.bulk.tensor.2d (SASS UTMALDG) through a tensor map with 128-byte swizzle, 128 rows x
//     128 bytes (16 KB) per copy: row-wise 128-bit reads of such a unit are bank-conflict free;
//   * the raw float32 bits ARE the high operand (kind::tf32 ignores the 13 low significand bits: hi = trunc_tf32(x), no
//     instruction); the low operand is lo = x - trunc_tf32(x) (exact), rounded to tf32 by an integer add of half an ulp:
//     three instructions per element; both halves are written to TENSOR MEMORY (tcgen05.st, thread = row) and the MMAs
//     take their A operand from there (no shared-memory operand traffic for A);
//   * x_hi.[c_hi; c_lo] (N = 32) and x_lo.c_hi (N = 16) per K step of 8 accumulate float32-grade dot products in 32 TMEM
//     columns (error of a product < 2^-21 |x||c|, unbiased); the MMA warp runs in uniform control flow with only the
//     tcgen05 instructions elected;
//   * arg-min (4 warps, thread = row, tcgen05.ld) ends with a stable counting sort of the tile's rows by label, and the
//     M-step (8 warps, half-warp = 8 consecutive sorted rows, lane = 16 bytes of the row) sums runs of equal labels in
//     registers and flushes a run into its private accumulator row when the label changes: the label is an address, no
//     atomics, fixed order, deterministic; the rows are read from the same resident raw units (no second read of X).
// One persistent CTA per SM, 22 warps: TMA producer | MMA issuer | 2 x 4 operand-half warps (alternate units) | 4 arg-min
// warps | 8 M-step warps, linked by mbarriers only.  The float64 form (reference sets of the gap statistic) follows the
// float32 kernel below.  Only the Lloyd-loop form of the pass (DIC_KM_NO_INERTIA, labels re-assigned) takes these
// kernels; the final labelling pass with its direct ||x - c||^2 sums stays on the CUDA-core kernels.
// Measured: DESIGN.md section 4.5, profiles/r02_kmeans_*; phase probes: benchmarks/_ktc_probe.sh (-DDIC_KTC_SKIP).
#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

#ifndef DIC_KTC_SKIP
#define DIC_KTC_SKIP 0      // benchmark builds: phases removed at compile time (benchmarks/_ktc_probe.sh)
#endif

namespace dic {
namespace {

using namespace tc;

constexpr int kRows = 128;                     // rows per tile (M of the MMA)
constexpr int kUnitB = kRows * 128;            // one ring unit: 128 rows x 128 bytes, 128-byte swizzle (16 KB)
constexpr int kMaxUnits = 12;
constexpr int kLboB = 32 * 16;                 // centres: 32 rows (16 hi | 16 lo) x 16 bytes per K chunk (no swizzle)
constexpr int kBTileB = 16 * kLboB;            // stacked centre tile of one 64-wide K chunk (8 KB)
constexpr int kSbo = 128;                      // bytes between 8-row groups of the centre tiles
constexpr int kNLoGroups = 2;                  // the low-half / split role: groups of 4 warps taking alternate units
constexpr int kNLo = 4 * kNLoGroups, kNArg = 4, kNAcc = 8;  // warps per role
constexpr int kTcThreads = 32 * (2 + kNLo + kNArg + kNAcc);
constexpr uint32_t kIdescTf32N16 = make_idesc(2u, 128u, 16u);
constexpr uint32_t kIdescTf32N32 = make_idesc(2u, 128u, 32u);
constexpr uint32_t kTmemCols = 256;            // [0,64): two accumulators of 32 columns; [64,128): two units of high
constexpr uint32_t kTmemHi = 64, kTmemLo = 128;   // operand halves (the raw float32 bits); [128,192): their low halves

struct TcBars {
  uint64_t full[kMaxUnits], slot_free[kMaxUnits];
  uint64_t lo_full[2], lo_free[2], acc_full[2], acc_free[2], lab_full[2], lab_free[2];
  uint32_t tmem_base;
  int timeout;
};

// Shared-memory plan (offsets from the 1024-byte aligned base); `nu` ring units.
struct TcPlan {
  int nu;
  size_t bt, sacc, scn, scnt, ssort, wcnt, bars, total;
};
__host__ __device__ inline TcPlan tc_plan(int NCH, int K, int nu) {
  TcPlan p;
  p.nu = nu;
  p.bt = (size_t)nu * kUnitB;
  p.sacc = p.bt + (size_t)NCH * kBTileB;                      // [G][K][D] floats, G * NCH = 8
  p.scn = p.sacc + (size_t)8 * K * 64 * 4;
  p.scnt = p.scn + 64;                                        // [8][16] ints
  p.ssort = p.scnt + 8 * 16 * 4;                              // [2][128] ints: (label << 8 | row), sorted by label
  p.wcnt = p.ssort + 2 * kRows * 4;                           // [2][4][32] ints: per-warp label counts of a tile
  p.bars = p.wcnt + 2 * 4 * 32 * 4;
  p.total = p.bars + sizeof(TcBars) + 1024;                   // + alignment slack
  return p;
}

__device__ __forceinline__ bool wait_bar(TcBars* B, uint64_t* bar, uint32_t phase) {
  if (*reinterpret_cast<volatile int*>(&B->timeout)) return false;
  if (bar_wait_bounded(bar, phase)) return true;
  *reinterpret_cast<volatile int*>(&B->timeout) = 1;      // a stalled hand-off ends the pass with NaN statistics, not a hang
  return false;
}

__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ float lo_tf32(float x) {
  // x - trunc_tf32(x) is exact; + 0x1000 on the bit pattern = round to nearest (ties away) once the tensor core drops
  // the 13 low bits of the operand
  const float lo = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  return __uint_as_float(__float_as_uint(lo) + 0x1000u);
}

__device__ __forceinline__ void split_tf32(const float4& x, float4& h, float4& l) {
  h.x = to_tf32(x.x); h.y = to_tf32(x.y); h.z = to_tf32(x.z); h.w = to_tf32(x.w);
  l.x = to_tf32(x.x - h.x); l.y = to_tf32(x.y - h.y); l.z = to_tf32(x.z - h.z); l.w = to_tf32(x.w - h.w);
}

// Ring position of the i-th unit this CTA handles.
struct RingPos {
  int slot;
  uint32_t phase;
  __device__ __forceinline__ void advance(int nu) {
    if (++slot == nu) {
      slot = 0;
      phase ^= 1u;
    }
  }
};

// Ring position `off` (< nu) units after `base` (no division: the per-tile positions are tracked incrementally).
__device__ __forceinline__ RingPos ring_at(const RingPos& base, int off, int nu) {
  RingPos r{base.slot + off, base.phase};
  if (r.slot >= nu) {
    r.slot -= nu;
    r.phase ^= 1u;
  }
  return r;
}

// (An M-step that re-reads the tile from L2 instead of holding its units - D = 256: one tile is 8 of the <= 12 units - was
// measured: 0.135 ms per pass while the E-step was slower, 0.18 against 0.14-0.16 ms of the resident form once the
// front of the pipeline ran at HBM speed, because the re-read then misses L2.  Not kept.)
template <int NCH>
__global__ void __launch_bounds__(kTcThreads, 1)
kmeans_assign_tc_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ centers, int32_t* labels,
                        double* __restrict__ ws, int64_t N, int K, int flags, int want_sums, int nu,
                        const double* __restrict__ done) {
  if (done && *done != 0.0) return;
  constexpr int D = 64 * NCH, G = 8 / NCH, UPT = 2 * NCH;     // units per tile
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const TcPlan P = tc_plan(NCH, K, nu);
  unsigned char* units = smem;
  unsigned char* bt = smem + P.bt;
  float* sacc = reinterpret_cast<float*>(smem + P.sacc);
  float* scn = reinterpret_cast<float*>(smem + P.scn);
  int* scnt = reinterpret_cast<int*>(smem + P.scnt);
  int* ssort = reinterpret_cast<int*>(smem + P.ssort);
  int* wcnt = reinterpret_cast<int*>(smem + P.wcnt);
  TcBars* B = reinterpret_cast<TcBars*>(smem + P.bars);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < kMaxUnits; ++s) {
      mbar_init(&B->full[s], 1);
      mbar_init(&B->slot_free[s], 4 + (want_sums ? G : 0));   // operand halves taken (+ the M-step warps)
    }
    for (int k = 0; k < 2; ++k) {
      mbar_init(&B->lo_full[k], 4);
      mbar_init(&B->lo_free[k], 1);
      mbar_init(&B->acc_full[k], 1);
      mbar_init(&B->acc_free[k], kNArg);
      mbar_init(&B->lab_full[k], kNArg);
      mbar_init(&B->lab_free[k], kNAcc);
    }
    B->timeout = 0;
    fence_proxy_async();
  }
  if (warp == 0) tmem_alloc(&B->tmem_base, kTmemCols);
  // centres -> stacked split operand tiles [chunk][K chunk][16 hi rows | 16 lo rows][16 bytes] (rows >= K are zero),
  // norms, zeroed accumulators
  for (int idx = tid; idx < NCH * 256; idx += kTcThreads) {
    const int n = idx & 15, c16 = (idx >> 4) & 15, ch = idx >> 8;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f), h, l;
    if (n < K) x = __ldg(reinterpret_cast<const float4*>(centers + (size_t)n * D + ch * 64 + c16 * 4));
    split_tf32(x, h, l);
    const int off = ch * kBTileB + c16 * kLboB + n * 16;
    *reinterpret_cast<float4*>(bt + off) = h;
    *reinterpret_cast<float4*>(bt + off + 256) = l;
  }
  if (want_sums)
    for (int i = tid; i < 8 * K * 64; i += kTcThreads) sacc[i] = 0.f;
  for (int i = tid; i < 8 * 16; i += kTcThreads) scnt[i] = 0;
  for (int i = tid; i < 2 * 4 * 32; i += kTcThreads) wcnt[i] = 0;
  if (tid < 256) {                               // ||c_k||^2: 16 lanes per centre, fixed summation order
    const int k = tid >> 4, l16 = tid & 15;
    float sum = 0.f;
    if (k < K) {
#pragma unroll
      for (int d = 0; d < D / 16; ++d) {
        const float c = __ldg(centers + (size_t)k * D + d * 16 + l16);
        sum = fmaf(c, c, sum);
      }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (l16 == 0) scn[k] = sum;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = B->tmem_base;
  const bool count_changes = (flags & DIC_KM_COUNT_CHANGES) != 0;
  const int64_t ntiles = (N + kRows - 1) / kRows;
  int changed = 0;

  if (warp == 0) {
    // ================= producer: one tensor-map copy per unit =================
    RingPos rp{0, 0u};
    bool ok = true;
    for (int64_t t = blockIdx.x; t < ntiles && ok; t += gridDim.x) {
      for (int u = 0; u < UPT; ++u) {
        ok = __all_sync(0xffffffffu, wait_bar(B, &B->slot_free[rp.slot], rp.phase ^ 1u));
        if (!ok) break;
        if (elect_one()) {
          mbar_expect_tx(&B->full[rp.slot], (uint32_t)kUnitB);
          tma_load_2d(units + (size_t)rp.slot * kUnitB, &tmap, u * 32, (int)(t * kRows), &B->full[rp.slot]);
        }
        rp.advance(nu);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp in uniform control flow, tcgen05 under elect) =================
    const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem);
    const uint32_t bt_u = smem_u32(bt);
    uint32_t v = 0;                                 // unit counter (operand ring in tensor memory)
    int64_t tl = 0;
    bool ok = true;
    for (int64_t t = blockIdx.x; t < ntiles && ok; t += gridDim.x, ++tl) {
      const uint32_t buf = (uint32_t)(tl & 1);
      const uint32_t d_tmem = tmem_u + buf * 32u;
      for (int u = 0; u < UPT && ok; ++u, ++v) {
        const uint32_t lb = v & 1u;
        // (lo_full implies the unit has landed; the raw unit itself may already be recycled when nothing else holds it)
        ok = __all_sync(0xffffffffu, wait_bar(B, &B->lo_full[lb], (v >> 1) & 1u));
        if (u == 0) ok = ok && __all_sync(0xffffffffu, wait_bar(B, &B->acc_free[buf], ((uint32_t)(tl >> 1) & 1u) ^ 1u));
        if (!ok) break;
        tc_fence_after();
        const uint32_t a_hi = tmem_u + kTmemHi + lb * 32u, a_lo = tmem_u + kTmemLo + lb * 32u;
        const uint32_t bo = bt_u + (uint32_t)((u >> 1) * kBTileB + (u & 1) * 8 * kLboB);
        uint64_t db[4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) db[ks] = make_desc_kmajor(bo + ks * 2 * kLboB, kLboB, kSbo);   // K = 8 per MMA
        if (elect_one()) {
          if (!(DIC_KTC_SKIP & 2)) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              // columns [0,16): x_hi.c_hi + x_lo.c_hi, columns [16,32): x_hi.c_lo
              umma_tf32_ts(d_tmem, a_hi + ks * 8, db[ks], kIdescTf32N32, (u > 0 || ks > 0) ? 1u : 0u);
              umma_tf32_ts(d_tmem, a_lo + ks * 8, db[ks], kIdescTf32N16, 1u);
            }
          }
          umma_commit(&B->lo_free[lb]);                         // the unit's operand columns (TMEM) are reusable
          if (u == UPT - 1) umma_commit(&B->acc_full[buf]);     // the tile's dots are complete
        }
        __syncwarp();
      }
    }
  } else if (warp < 2 + kNLo) {
    // ================= operand halves: thread = row, raw bits | lo = rn_tf32(x - trunc_tf32(x)) -> TMEM (A operand) =========
    const int q = warp & 3, row = 32 * q + lane;
    const uint32_t grp = (uint32_t)(warp - 2) >> 2;          // this group takes the units with (unit & 1) == grp
    const uint32_t sw = (uint32_t)(row & 7);
    const size_t roff = (size_t)(row >> 3) * 1024 + (size_t)sw * 128;
    RingPos rp{0, 0u};
    uint32_t v = 0;
    bool ok = true;
    for (int64_t t = blockIdx.x; t < ntiles && ok; t += gridDim.x) {
      for (int u = 0; u < UPT; ++u, ++v) {
        const uint32_t lb = v & 1u;
        if (lb != grp) {
          rp.advance(nu);
          continue;
        }
        ok = wait_bar(B, &B->full[rp.slot], rp.phase);
        ok = ok && wait_bar(B, &B->lo_free[lb], ((v >> 1) & 1u) ^ 1u);
        ok = __all_sync(0xffffffffu, ok);
        if (!ok) break;
        tc_fence_after();
        const unsigned char* src = units + (size_t)rp.slot * kUnitB + roff;
        const uint32_t t_hi = tmem + ((uint32_t)(32 * q) << 16) + kTmemHi + lb * 32u;
        const uint32_t t_lo = tmem + ((uint32_t)(32 * q) << 16) + kTmemLo + lb * 32u;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {             // 16 elements at a time: raw bits = high operand, lo_tf32 = low
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!(DIC_KTC_SKIP & 1)) x = *reinterpret_cast<const float4*>(src + (((uint32_t)(4 * hh + c) ^ sw) << 4));
            hi[4 * c + 0] = __float_as_uint(x.x); lo[4 * c + 0] = __float_as_uint(lo_tf32(x.x));
            hi[4 * c + 1] = __float_as_uint(x.y); lo[4 * c + 1] = __float_as_uint(lo_tf32(x.y));
            hi[4 * c + 2] = __float_as_uint(x.z); lo[4 * c + 2] = __float_as_uint(lo_tf32(x.z));
            hi[4 * c + 3] = __float_as_uint(x.w); lo[4 * c + 3] = __float_as_uint(lo_tf32(x.w));
          }
          tmem_st16(t_hi + 16 * hh, hi);
          tmem_st16(t_lo + 16 * hh, lo);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&B->lo_full[lb]);
          mbar_arrive(&B->slot_free[rp.slot]);
        }
        rp.advance(nu);
      }
    }
  } else if (warp < 2 + kNLo + kNArg) {
    // ================= arg-min: thread = row; then the tile's rows sorted by label for the M-step =================
    const int q = warp & 3;                           // TMEM lane quarter this warp may read
    const int row = 32 * q + lane;
    const uint32_t lt_mask = (1u << lane) - 1u;
    int64_t tl = 0;
    int mycount = 0;
    bool ok = true;
    int oldl_next = -1;
    if (count_changes && (int64_t)blockIdx.x * kRows + row < N) oldl_next = labels[(int64_t)blockIdx.x * kRows + row];
    for (int64_t t = blockIdx.x; t < ntiles && ok; t += gridDim.x, ++tl) {
      const uint32_t buf = (uint32_t)(tl & 1);
      const uint32_t ph = (uint32_t)(tl >> 1) & 1u;
      const int64_t row0 = t * kRows;
      const int rows = (int)min((int64_t)kRows, N - row0);
      const int oldl = oldl_next;                      // previous labels: fetched one tile ahead
      {
        const int64_t rn = row0 + (int64_t)gridDim.x * kRows + row;
        oldl_next = (count_changes && rn < N) ? labels[rn] : -1;
      }
      ok = __all_sync(0xffffffffu, wait_bar(B, &B->acc_full[buf], ph));
      if (!ok) break;
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + buf * 32u, v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&B->acc_free[buf]);           // this warp has its dots in registers
      int best = 0;
      float bestd = fmaf(-2.f, __uint_as_float(v[0]) + __uint_as_float(v[16]), scn[0]);
#pragma unroll
      for (int k = 1; k < 16; ++k) {
        if (k >= K) break;                                              // (uniform)
        const float dk = fmaf(-2.f, __uint_as_float(v[k]) + __uint_as_float(v[16 + k]), scn[k]);   // ||c||^2 - 2 x.c
        if (dk < bestd) {                                               // strict '<': lowest index wins ties
          bestd = dk;
          best = k;
        }
      }
      if (row < rows) {
        if (count_changes) changed += (oldl != best);
        labels[row0 + row] = best;
      }
      if (want_sums) {
        // stable counting sort of the tile's 128 rows by label (bin 16 = rows past N): position = rows of smaller
        // bins + rows of the same bin in earlier warps + rank inside the warp.  Every warp owns one row of the count
        // table; the row of the OTHER buffer is cleared after the barrier (its readers passed the previous barrier).
        const int bin = row < rows ? best : 16;
        const uint32_t peers = __match_any_sync(0xffffffffu, bin);
        const int myrank = __popc(peers & lt_mask);
        int* wc = wcnt + (int)buf * 128;
        if (myrank == 0) wc[q * 32 + bin] = __popc(peers);
        named_bar_sync(1, 32 * kNArg);
        const int c0 = wc[lane], c1 = wc[32 + lane], c2 = wc[64 + lane], c3 = wc[96 + lane];
        wcnt[(int)(buf ^ 1u) * 128 + q * 32 + lane] = 0;
        const int tot = c0 + c1 + c2 + c3;
        if (lane < 16) mycount += q == 0 ? c0 : (q == 1 ? c1 : (q == 2 ? c2 : c3));
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        const int base = incl - tot + (q > 0 ? c0 : 0) + (q > 1 ? c1 : 0) + (q > 2 ? c2 : 0);
        const int pos = __shfl_sync(0xffffffffu, base, bin) + myrank;
        ok = __all_sync(0xffffffffu, wait_bar(B, &B->lab_free[buf], ph ^ 1u));
        if (!ok) break;
        ssort[buf * kRows + pos] = (bin << 8) | row;
        __syncwarp();
        if (lane == 0) mbar_arrive(&B->lab_full[buf]);
      }
    }
    if (lane < 16) scnt[q * 16 + lane] = mycount;       // rows per cluster seen by this warp
  } else if (want_sums) {
    // ================= M-step accumulation: warp = (part g of the sorted rows, chunk ch) =================
    // A half-warp owns 8 consecutive sorted rows per step and a lane 16 bytes of the 256-byte chunk row: runs of equal
    // labels are summed in registers and flushed into the warp's private sacc[g][label] row when the label changes
    // (the label is an address: no atomics, fixed order, deterministic).  The two half-warps of a warp walk
    // consecutive ranges, so their flush targets differ except for the last run of the first against the second's:
    // that one is handed over by shuffle.  Rows past N are zero (TMA fill) and carry bin 16: never flushed.
    const int w = warp - (2 + kNLo + kNArg);
    const int g = w % G, ch = w / G;
    const int h = lane >> 4, cc = lane & 15;
    constexpr int NPOS = kRows / G;                   // sorted rows per warp: 16 / 32 / 64
    float4* acc0 = reinterpret_cast<float4*>(sacc + (size_t)g * K * D + ch * 64) + cc;
    int64_t tl = 0;
    RingPos tp{0, 0u};                                // first unit of the current tile
    bool ok = true;
    for (int64_t t = blockIdx.x; t < ntiles && ok; t += gridDim.x, ++tl) {
      const uint32_t buf = (uint32_t)(tl & 1);
      // ring positions of this warp's two units of the tile (halves of chunk ch)
      const RingPos r0 = ring_at(tp, 2 * ch, nu), r1 = ring_at(tp, 2 * ch + 1, nu);
      tp = ring_at(tp, UPT, nu);
      const int s0 = r0.slot, s1 = r1.slot;
      const uint32_t p0 = r0.phase, p1 = r1.phase;
      ok = wait_bar(B, &B->lab_full[buf], (uint32_t)(tl >> 1) & 1u);
      ok = ok && wait_bar(B, &B->full[s0], p0) && wait_bar(B, &B->full[s1], p1);   // (long complete: acquire only)
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      const unsigned char* ub = units + (size_t)((cc & 8) ? s1 : s0) * kUnitB;
      const int4* srt = reinterpret_cast<const int4*>(ssort + buf * kRows + g * NPOS + h * (NPOS / 2));
      auto flush = [&](int lab, const float4& a) {
        if (lab < 16) {
          float4* dst = acc0 + (size_t)lab * (D / 4);
          float4 o = *dst;
          o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
          *dst = o;
        }
      };
      int cur = 0;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!(DIC_KTC_SKIP & 4)) {
#pragma unroll 1
        for (int st = 0; st < NPOS / 16; ++st) {
          const int4 ea = srt[2 * st], eb = srt[2 * st + 1];
          const int e[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
          float4 x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = e[j] & 255;
            x[j] = *reinterpret_cast<const float4*>(ub + ((((r << 3) | ((r ^ cc) & 7))) << 4));
          }
          if (st == 0) cur = e[0] >> 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int lab = e[j] >> 8;
            if (lab != cur) {
              flush(cur, acc);
              cur = lab;
              acc = x[j];
            } else {
              acc.x += x[j].x; acc.y += x[j].y; acc.z += x[j].z; acc.w += x[j].w;
            }
          }
        }
        // Last runs.  In sorted order the flush targets of the two half-warps differ (first half < second half) unless
        // the whole second half continues the first half's last label: then the first half's sum is handed over.
        __syncwarp();
        const int last1 = __shfl_sync(0xffffffffu, cur, 0), last2 = __shfl_sync(0xffffffffu, cur, 16);
        if (last1 == last2) {
          float4 o;
          o.x = __shfl_xor_sync(0xffffffffu, acc.x, 16);
          o.y = __shfl_xor_sync(0xffffffffu, acc.y, 16);
          o.z = __shfl_xor_sync(0xffffffffu, acc.z, 16);
          o.w = __shfl_xor_sync(0xffffffffu, acc.w, 16);
          if (h == 1) {
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
            flush(cur, acc);
          }
        } else {
          flush(cur, acc);
        }
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&B->slot_free[s0]);
        mbar_arrive(&B->slot_free[s1]);
        mbar_arrive(&B->lab_free[buf]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();

  // per-block partials: [K*D] sums | [K] counts | inertia | changed | dist_sum | 0   (layout of kmeans_finish_kernel)
  double* out = ws + (int64_t)blockIdx.x * ((int64_t)K * D + K + 4);
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const bool dead = B->timeout != 0;
  if (want_sums) {
    for (int i = tid; i < K * D; i += kTcThreads) {
      double sum = 0.0;
      for (int gg = 0; gg < G; ++gg) sum += (double)sacc[(size_t)gg * K * D + i];
      out[i] = dead ? nan : sum;
    }
    for (int i = tid; i < K; i += kTcThreads) {
      int cc = 0;
      for (int gg = 0; gg < 8; ++gg) cc += scnt[gg * 16 + i];
      out[(int64_t)K * D + i] = (double)cc;
    }
  }
  double* red = reinterpret_cast<double*>(units);
  const double chs = warp_sum((double)changed);
  if (lane == 0) red[warp] = chs;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int wq = 0; wq < kTcThreads / 32; ++wq) s += red[wq];
    out[(int64_t)K * D + K + 0] = dead ? nan : 0.0;      // inertia: not computed in this form of the pass
    out[(int64_t)K * D + K + 1] = dead ? nan : s;
    out[(int64_t)K * D + K + 2] = 0.0;
    out[(int64_t)K * D + K + 3] = 0.0;
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// float64 rows (the reference sets of the gap statistic are numpy float64 draws, p2_clustering_optK.py:353-410), D = 64.
// The tensor core SCREENS: every double is split into two tf32 halves (hi = trunc_tf32(float(x)), lo = rn_tf32(x - hi)),
// both written to tensor memory as A operands (TS MMAs: no shared-memory operand traffic at all), and the float32-grade
// distances decide the label wherever the margin between the two nearest centres exceeds their error bound
// 2^-17 (|x|^2 + max|c|^2) (the products carry < 2^-21 |x||c| each).  Rows inside the bound - a few per million - are
// re-evaluated exactly, direct (x - c)^2 in float64 from the resident tile, by the whole warp.  Sums are float64.
// Units are 128 rows x 16 doubles (128 bytes, SWIZZLE_128B: conflict-free row reads), 4 per tile.
struct Tc64Plan {
  int nu;
  size_t bt, cen, sacc, scn32, scn64, xn, ssort, wcnt, bars, total;
};
__host__ __device__ inline Tc64Plan tc64_plan(int K, int nu) {
  Tc64Plan p;
  p.nu = nu;
  p.bt = (size_t)nu * kUnitB;
  p.cen = p.bt + kBTileB;                                     // [16][64] doubles (exact re-evaluation)
  p.sacc = p.cen + 16 * 64 * 8;                               // [8][K][64] doubles
  p.scn32 = p.sacc + (size_t)8 * K * 64 * 8;
  p.scn64 = p.scn32 + 64;
  p.xn = p.scn64 + 128;                                       // [2][2][128] floats: |x|^2 of a tile's rows, per group
  p.ssort = p.xn + 4 * kRows * 4;
  p.wcnt = p.ssort + 2 * kRows * 4;
  p.bars = p.wcnt + 2 * 4 * 32 * 4;
  p.total = p.bars + sizeof(TcBars) + 1024;
  return p;
}

__global__ void __launch_bounds__(kTcThreads, 1)
kmeans_assign_tc64_kernel(const __grid_constant__ CUtensorMap tmap, const double* __restrict__ centers, int32_t* labels,
                          double* __restrict__ ws, int64_t N, int K, int flags, int want_sums, int nu,
                          const double* __restrict__ done) {
  if (done && *done != 0.0) return;
  constexpr int D = 64, UPT = 4;
  constexpr uint32_t kHi = 64, kLo = 96;                       // TMEM columns: [0,64) accumulators, 2 x 16 hi, 2 x 16 lo
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const Tc64Plan P = tc64_plan(K, nu);
  unsigned char* units = smem;
  unsigned char* bt = smem + P.bt;
  double* cen = reinterpret_cast<double*>(smem + P.cen);
  double* sacc = reinterpret_cast<double*>(smem + P.sacc);
  float* scn32 = reinterpret_cast<float*>(smem + P.scn32);
  double* scn64 = reinterpret_cast<double*>(smem + P.scn64);
  float* xnorm = reinterpret_cast<float*>(smem + P.xn);
  int* ssort = reinterpret_cast<int*>(smem + P.ssort);
  int* wcnt = reinterpret_cast<int*>(smem + P.wcnt);
  TcBars* B = reinterpret_cast<TcBars*>(smem + P.bars);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < kMaxUnits; ++s) {
      mbar_init(&B->full[s], 1);
      mbar_init(&B->slot_free[s], 4 + kNArg + (want_sums ? kNAcc : 0));   // split, arg-min (exact rows), M-step
    }
    for (int k = 0; k < 2; ++k) {
      mbar_init(&B->lo_full[k], 4);
      mbar_init(&B->lo_free[k], 1);
      mbar_init(&B->acc_full[k], 1);
      mbar_init(&B->acc_free[k], kNArg);
      mbar_init(&B->lab_full[k], kNArg);
      mbar_init(&B->lab_free[k], kNAcc);
    }
    B->timeout = 0;
    fence_proxy_async();
  }
  if (warp == 0) tmem_alloc(&B->tmem_base, 128);
  // centres: float64 copy, stacked split tf32 operand tile [K chunk][16 hi rows | 16 lo rows][16 bytes], norms
  for (int idx = tid; idx < 256; idx += kTcThreads) {
    const int n = idx & 15, c16 = idx >> 4;
    float h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double c = n < K ? centers[(size_t)n * D + c16 * 4 + i] : 0.0;
      cen[n * D + c16 * 4 + i] = c;
      h[i] = to_tf32((float)c);
      l[i] = to_tf32((float)(c - (double)h[i]));
    }
    const int off = c16 * kLboB + n * 16;
    *reinterpret_cast<float4*>(bt + off) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(bt + off + 256) = make_float4(l[0], l[1], l[2], l[3]);
  }
  if (want_sums)
    for (int i = tid; i < 8 * K * 64; i += kTcThreads) sacc[i] = 0.0;
  for (int i = tid; i < 2 * 4 * 32; i += kTcThreads) wcnt[i] = 0;
  if (tid < 256) {                               // ||c_k||^2: 16 lanes per centre
    const int k = tid >> 4, l16 = tid & 15;
    double sum = 0.0;
    if (k < K) {
#pragma unroll
      for (int d = 0; d < D / 16; ++d) {
        const double c = centers[(size_t)k * D + d * 16 + l16];
        sum = fma(c, c, sum);
      }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (l16 == 0) {
      scn64[k] = sum;
      scn32[k] = (float)sum;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = B->tmem_base;
  const bool count_changes = (flags & DIC_KM_COUNT_CHANGES) != 0;
  const int64_t ntiles = (N + kRows - 1) / kRows;
  int changed = 0;
  int mycount = 0;

  if (warp == 0) {
    // ================= producer =================
    RingPos rp{0, 0u};
    bool ok = true;
    for (int64_t t = blockIdx.x; t < ntiles && ok; t += gridDim.x) {
      for (int u = 0; u < UPT; ++u) {
        ok = __all_sync(0xffffffffu, wait_bar(B, &B->slot_free[rp.slot], rp.phase ^ 1u));
        if (!ok) break;
        if (elect_one()) {
          mbar_expect_tx(&B->full[rp.slot], (uint32_t)kUnitB);
          tma_load_2d(units + (size_t)rp.slot * kUnitB, &tmap, u * 16, (int)(t * kRows), &B->full[rp.slot]);
        }
        rp.advance(nu);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: both operand halves come from tensor memory =================
    const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem);
    const uint32_t bt_u = smem_u32(bt);
    uint32_t v = 0;
    int64_t tl = 0;
    bool ok = true;
    for (int64_t t = blockIdx.x; t < ntiles && ok; t += gridDim.x, ++tl) {
      const uint32_t buf = (uint32_t)(tl & 1);
      const uint32_t d_tmem = tmem_u + buf * 32u;
      for (int u = 0; u < UPT && ok; ++u, ++v) {
        const uint32_t lb = v & 1u;
        ok = __all_sync(0xffffffffu, wait_bar(B, &B->lo_full[lb], (v >> 1) & 1u));
        if (u == 0) ok = ok && __all_sync(0xffffffffu, wait_bar(B, &B->acc_free[buf], ((uint32_t)(tl >> 1) & 1u) ^ 1u));
        if (!ok) break;
        tc_fence_after();
        const uint32_t a_hi = tmem_u + kHi + lb * 16u, a_lo = tmem_u + kLo + lb * 16u;
        uint64_t db[2];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) db[ks] = make_desc_kmajor(bt_u + (u * 4 + ks * 2) * kLboB, kLboB, kSbo);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            // columns [0,16): x_hi.c_hi + x_lo.c_hi, columns [16,32): x_hi.c_lo
            umma_tf32_ts(d_tmem, a_hi + ks * 8, db[ks], kIdescTf32N32, (u > 0 || ks > 0) ? 1u : 0u);
            umma_tf32_ts(d_tmem, a_lo + ks * 8, db[ks], kIdescTf32N16, 1u);
          }
          umma_commit(&B->lo_free[lb]);
          if (u == UPT - 1) umma_commit(&B->acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 2 + kNLo) {
    // ================= split: thread = row; 16 doubles of a unit -> hi | lo tf32 columns in tensor memory =================
    const int q = warp & 3, row = 32 * q + lane;
    const uint32_t grp = (uint32_t)(warp - 2) >> 2;          // this group takes the units with (unit & 1) == grp
    const uint32_t sw = (uint32_t)(row & 7);
    const size_t roff = (size_t)row * 128;
    RingPos rp{0, 0u};
    uint32_t v = 0;
    int64_t tl = 0;
    bool ok = true;
    for (int64_t t = blockIdx.x; t < ntiles && ok; t += gridDim.x, ++tl) {
      float xn = 0.f;
      for (int u = 0; u < UPT; ++u, ++v) {
        const uint32_t lb = v & 1u;
        if (lb != grp) {
          rp.advance(nu);
          continue;
        }
        ok = wait_bar(B, &B->full[rp.slot], rp.phase) && wait_bar(B, &B->lo_free[lb], ((v >> 1) & 1u) ^ 1u);
        ok = __all_sync(0xffffffffu, ok);
        if (!ok) break;
        tc_fence_after();
        const unsigned char* src = units + (size_t)rp.slot * kUnitB + roff;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const double2 x = *reinterpret_cast<const double2*>(src + (((uint32_t)c ^ sw) << 4));
          const double xs[2] = {x.x, x.y};
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float f = (float)xs[i];
            const uint32_t hb = __float_as_uint(f) & 0xFFFFE000u;
            const float l = (float)(xs[i] - (double)__uint_as_float(hb));
            hi[2 * c + i] = hb;
            lo[2 * c + i] = __float_as_uint(l) + 0x1000u;
            xn = fmaf(f, f, xn);
          }
        }
        tmem_st16(tmem + ((uint32_t)(32 * q) << 16) + kHi + lb * 16u, hi);
        tmem_st16(tmem + ((uint32_t)(32 * q) << 16) + kLo + lb * 16u, lo);
        tmem_st_wait();
        if (u >= UPT - 2) xnorm[((tl & 1) * 2 + grp) * kRows + row] = xn;   // this group's half of |x|^2 (arg-min warps)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&B->lo_full[lb]);
          mbar_arrive(&B->slot_free[rp.slot]);
        }
        rp.advance(nu);
      }
    }
  } else if (warp < 2 + kNLo + kNArg) {
    // ================= arg-min (float32 screen + exact float64 rows), counting sort =================
    const int q = warp & 3;
    const int row = 32 * q + lane;
    const uint32_t lt_mask = (1u << lane) - 1u;
    float cmax = 0.f;
    for (int k = 0; k < K; ++k) cmax = fmaxf(cmax, scn32[k]);
    int64_t tl = 0;
    RingPos tp{0, 0u};                                // first unit of the current tile
    bool ok = true;
    int oldl_next = -1;
    if (count_changes && (int64_t)blockIdx.x * kRows + row < N) oldl_next = labels[(int64_t)blockIdx.x * kRows + row];
    for (int64_t t = blockIdx.x; t < ntiles && ok; t += gridDim.x, ++tl) {
      const uint32_t buf = (uint32_t)(tl & 1);
      const uint32_t ph = (uint32_t)(tl >> 1) & 1u;
      const int64_t row0 = t * kRows;
      const int rows = (int)min((int64_t)kRows, N - row0);
      const int oldl = oldl_next;
      {
        const int64_t rn = row0 + (int64_t)gridDim.x * kRows + row;
        oldl_next = (count_changes && rn < N) ? labels[rn] : -1;
      }
      ok = __all_sync(0xffffffffu, wait_bar(B, &B->acc_full[buf], ph));
      if (!ok) break;
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + buf * 32u, v);
      const float xn_r = xnorm[buf * 2 * kRows + row] + xnorm[(buf * 2 + 1) * kRows + row];   // (before acc_free)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&B->acc_free[buf]);
      int best = 0;
      float b1 = fmaf(-2.f, __uint_as_float(v[0]) + __uint_as_float(v[16]), scn32[0]), b2 = 3.0e38f;
#pragma unroll
      for (int k = 1; k < 16; ++k) {
        if (k >= K) break;
        const float dk = fmaf(-2.f, __uint_as_float(v[k]) + __uint_as_float(v[16 + k]), scn32[k]);
        if (dk < b1) {
          b2 = b1;
          b1 = dk;
          best = k;
        } else {
          b2 = fminf(b2, dk);
        }
      }
      // rows whose two nearest centres are closer than the error bound of the screen: exact float64, whole warp per row
      const float tau = 1.52587890625e-05f * (xn_r + cmax) * 0.5f;
      uint32_t unc = __ballot_sync(0xffffffffu, row < rows && K > 1 && (b2 - b1) <= tau);
      if (unc) {
        const int un = lane >> 3, chunk = lane & 7;
        const RingPos ru = ring_at(tp, un, nu);
        const int slot = ru.slot;
        wait_bar(B, &B->full[slot], ru.phase);      // (long complete: acquire only)
        while (unc) {
          const int src = __ffs(unc) - 1;
          unc &= unc - 1;
          const int r = 32 * q + src;
          const double2 xv = *reinterpret_cast<const double2*>(units + (size_t)slot * kUnitB + (size_t)r * 128 +
                                                                ((chunk ^ (r & 7)) << 4));
          int eb = 0;
          double ed = 0.0;
          for (int k = 0; k < K; ++k) {
            const double2 cv = *reinterpret_cast<const double2*>(cen + k * D + 2 * lane);
            const double a = xv.x - cv.x, b = xv.y - cv.y;
            double part = fma(a, a, b * b);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (k == 0 || part < ed) {
              ed = part;
              eb = k;
            }
          }
          if (lane == src) best = eb;
        }
      }
      __syncwarp();
      if (lane == 0)
        for (int u = 0; u < UPT; ++u) mbar_arrive(&B->slot_free[ring_at(tp, u, nu).slot]);
      tp = ring_at(tp, UPT, nu);
      if (row < rows) {
        if (count_changes) changed += (oldl != best);
        labels[row0 + row] = best;
      }
      if (want_sums) {
        const int bin = row < rows ? best : 16;
        const uint32_t peers = __match_any_sync(0xffffffffu, bin);
        const int myrank = __popc(peers & lt_mask);
        int* wc = wcnt + (int)buf * 128;
        if (myrank == 0) wc[q * 32 + bin] = __popc(peers);
        named_bar_sync(1, 32 * kNArg);
        const int c0 = wc[lane], c1 = wc[32 + lane], c2 = wc[64 + lane], c3 = wc[96 + lane];
        wcnt[(int)(buf ^ 1u) * 128 + q * 32 + lane] = 0;
        const int tot = c0 + c1 + c2 + c3;
        if (lane < 16) mycount += q == 0 ? c0 : (q == 1 ? c1 : (q == 2 ? c2 : c3));
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        const int base = incl - tot + (q > 0 ? c0 : 0) + (q > 1 ? c1 : 0) + (q > 2 ? c2 : 0);
        const int pos = __shfl_sync(0xffffffffu, base, bin) + myrank;
        ok = __all_sync(0xffffffffu, wait_bar(B, &B->lab_free[buf], ph ^ 1u));
        if (!ok) break;
        ssort[buf * kRows + pos] = (bin << 8) | row;
        __syncwarp();
        if (lane == 0) mbar_arrive(&B->lab_full[buf]);
      }
    }
  } else if (want_sums) {
    // ================= M-step: warp = 16 sorted rows, lane = 2 doubles of the 512-byte row; float64 run sums =================
    const int g = warp - (2 + kNLo + kNArg);
    const int un = lane >> 3, chunk = lane & 7;
    double2* acc0 = reinterpret_cast<double2*>(sacc + (size_t)g * K * D) + lane;
    int64_t tl = 0;
    RingPos tp{0, 0u};                                // first unit of the current tile
    bool ok = true;
    for (int64_t t = blockIdx.x; t < ntiles && ok; t += gridDim.x, ++tl) {
      const uint32_t buf = (uint32_t)(tl & 1);
      const RingPos ru = ring_at(tp, un, nu);
      const int slot = ru.slot;
      ok = wait_bar(B, &B->lab_full[buf], (uint32_t)(tl >> 1) & 1u) &&
           wait_bar(B, &B->full[slot], ru.phase);   // (long complete: acquire only)
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      const unsigned char* ub = units + (size_t)slot * kUnitB;
      const int4* srt = reinterpret_cast<const int4*>(ssort + buf * kRows + g * 16);
      auto flush = [&](int lab, const double2& a) {
        if (lab < 16) {
          double2* dst = acc0 + (size_t)lab * (D / 2);
          double2 o = *dst;
          o.x += a.x;
          o.y += a.y;
          *dst = o;
        }
      };
      int cur = 0;
      double2 acc = make_double2(0.0, 0.0);
#pragma unroll 1
      for (int st = 0; st < 2; ++st) {
        const int4 ea = srt[2 * st], eb = srt[2 * st + 1];
        const int e[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
        double2 x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = e[j] & 255;
          x[j] = *reinterpret_cast<const double2*>(ub + (((r << 3) | ((r ^ chunk) & 7)) << 4));
        }
        if (st == 0) cur = e[0] >> 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int lab = e[j] >> 8;
          if (lab != cur) {
            flush(cur, acc);
            cur = lab;
            acc = x[j];
          } else {
            acc.x += x[j].x;
            acc.y += x[j].y;
          }
        }
      }
      flush(cur, acc);
      __syncwarp();
      if (lane == 0) {
        for (int u = 0; u < UPT; ++u) mbar_arrive(&B->slot_free[ring_at(tp, u, nu).slot]);
        mbar_arrive(&B->lab_free[buf]);
      }
      tp = ring_at(tp, UPT, nu);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp >= 2 + kNLo && warp < 2 + kNLo + kNArg && lane < 16) wcnt[(warp & 3) * 16 + lane] = mycount;
  __syncthreads();

  double* out = ws + (int64_t)blockIdx.x * ((int64_t)K * D + K + 4);
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const bool dead = B->timeout != 0;
  if (want_sums) {
    for (int i = tid; i < K * D; i += kTcThreads) {
      double sum = 0.0;
      for (int gg = 0; gg < 8; ++gg) sum += sacc[(size_t)gg * K * D + i];
      out[i] = dead ? nan : sum;
    }
    for (int i = tid; i < K; i += kTcThreads)
      out[(int64_t)K * D + i] = (double)(wcnt[i] + wcnt[16 + i] + wcnt[32 + i] + wcnt[48 + i]);
  }
  double* red = reinterpret_cast<double*>(units);
  const double chs = warp_sum((double)changed);
  if (lane == 0) red[warp] = chs;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int wq = 0; wq < kTcThreads / 32; ++wq) s += red[wq];
    out[(int64_t)K * D + K + 0] = dead ? nan : 0.0;
    out[(int64_t)K * D + K + 1] = dead ? nan : s;
    out[(int64_t)K * D + K + 2] = 0.0;
    out[(int64_t)K * D + K + 3] = 0.0;
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

}  // namespace

// The passes need cuTensorMapEncodeTiled from the driver; without it the dispatcher keeps the CUDA-core kernels.
bool kmeans_tc_available() { return encode_tiled_fn() != nullptr; }

// Shapes and flags the tensor-core pass covers (see the header comment).
bool kmeans_tc_covers(const void* X, int D, int K, int flags) {
  return (D == 64 || D == 128 || D == 256) && K >= 1 && K <= 16 && aligned16(X) &&
         (flags & DIC_KM_NO_INERTIA) != 0 && (flags & DIC_KM_KEEP_LABELS) == 0;
}

// Launches the pass; *nb_out = number of per-block partials written to `ws` (<= max_blocks).
int launch_kmeans_assign_tc(const float* X, const float* centers, int32_t* labels, double* ws, int64_t N, int D, int K,
                            int flags, int want_sums, const double* done, int max_blocks, int* nb_out,
                            cudaStream_t st) {
  int dev = 0, sms = 148;
  DIC_CUDA(cudaGetDevice(&dev));
  DIC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t ntiles = (N + kRows - 1) / kRows;
  int nb = sms < max_blocks ? sms : max_blocks;
  if (ntiles < nb) nb = (int)ntiles;
  if (nb < 1) nb = 1;
  *nb_out = nb;

  // rows as a 2-D tensor (D, N) of float32, boxes of 32 floats x 128 rows, 128-byte swizzle, rows past N read as zero
  EncodeTiledFn encode = encode_tiled_fn();
  DIC_REQUIRE(encode != nullptr, DIC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)N};
  const cuuint64_t gstride[1] = {(cuuint64_t)D * sizeof(float)};
  const cuuint32_t box[2] = {32u, (cuuint32_t)kRows};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2u, const_cast<float*>(X), gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DIC_REQUIRE(cr == CUDA_SUCCESS, DIC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for N=%lld D=%d", (int)cr,
              (long long)N, D);

  const int NCH = D / 64;
  int nu = kMaxUnits;
  while (nu > 2 * NCH && tc_plan(NCH, K, nu).total > (size_t)kMaxSmemBytes) --nu;
  const size_t smem = tc_plan(NCH, K, nu).total;
  DIC_REQUIRE(smem <= (size_t)kMaxSmemBytes, DIC_ERR_UNSUPPORTED, "tensor-core Lloyd pass: %zu bytes of shared memory",
              smem);
#define DIC_KTC(NCH_)                                                                                          \
  {                                                                                                            \
    auto kf = kmeans_assign_tc_kernel<NCH_>;                                                                   \
    DIC_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                \
    kf<<<nb, kTcThreads, smem, st>>>(tmap, centers, labels, ws, N, K, flags, want_sums, nu, done);             \
  }
  if (D == 64) DIC_KTC(1) else if (D == 128) DIC_KTC(2) else DIC_KTC(4)
#undef DIC_KTC
  DIC_LAUNCH_CHECK("kmeans_assign_tc_kernel");
  return DIC_OK;
}

bool kmeans_tc64_covers(const void* X, int D, int K, int flags) {
  return D == 64 && K >= 1 && K <= 16 && aligned16(X) && (flags & DIC_KM_NO_INERTIA) != 0 &&
         (flags & DIC_KM_KEEP_LABELS) == 0;
}

// float64 rows of 64 elements; same contract as launch_kmeans_assign_tc.
int launch_kmeans_assign_tc64(const double* X, const double* centers, int32_t* labels, double* ws, int64_t N, int K,
                              int flags, int want_sums, const double* done, int max_blocks, int* nb_out,
                              cudaStream_t st) {
  int dev = 0, sms = 148;
  DIC_CUDA(cudaGetDevice(&dev));
  DIC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t ntiles = (N + kRows - 1) / kRows;
  int nb = sms < max_blocks ? sms : max_blocks;
  if (ntiles < nb) nb = (int)ntiles;
  if (nb < 1) nb = 1;
  *nb_out = nb;
  EncodeTiledFn encode = encode_tiled_fn();
  DIC_REQUIRE(encode != nullptr, DIC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {64u, (cuuint64_t)N};
  const cuuint64_t gstride[1] = {64u * sizeof(double)};
  const cuuint32_t box[2] = {16u, (cuuint32_t)kRows};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2u, const_cast<double*>(X), gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DIC_REQUIRE(cr == CUDA_SUCCESS, DIC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for N=%lld (float64)", (int)cr,
              (long long)N);
  int nu = kMaxUnits;
  while (nu > 4 && tc64_plan(K, nu).total > (size_t)kMaxSmemBytes) --nu;
  const size_t smem = tc64_plan(K, nu).total;
  DIC_REQUIRE(smem <= (size_t)kMaxSmemBytes, DIC_ERR_UNSUPPORTED, "tensor-core Lloyd pass: %zu bytes of shared memory",
              smem);
  auto kf = kmeans_assign_tc64_kernel;
  DIC_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kf<<<nb, kTcThreads, smem, st>>>(tmap, centers, labels, ws, N, K, flags, want_sums, nu, done);
  DIC_LAUNCH_CHECK("kmeans_assign_tc64_kernel");
  return DIC_OK;
}

}  // namespace dic
