// Pairwise Euclidean distance sum on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// This is the one genuinely dense contraction on the hot path: the "inertia" of
// KM.compute_inertia_v1 / computer_intertia_v2 (p2_clustering_optK.py:334-351) sums
// ||x_i - x_j|| over the full n_c x n_c matrix of every cluster.  The reference materialises that
// matrix (sklearn pairwise_distances: ||x||^2 + ||y||^2 - 2 x.y, clamp, sqrt); here 128 x 128 tiles of
// the Gram matrix X X^T come out of tcgen05.mma (accumulator in TMEM), are read back with tcgen05.ld and
// reduced on the fly, so nothing is ever written to memory.  Two kernels:
//
//   * D <= 256: tc64::pairwise_tc64_kernel (second half of this file) - a pre-pass packs the rows ONCE as split
//     float16 operand tiles (kind::f16), TMA bulk copies feed them, one warp issues, eight warps drain; also
//     the silhouette's per-row, per-cluster sums (ROWSUMS).
//   * D > 256: pairwise_tc_kernel (first half) - the earlier register-staged design on split TF32
//     (kind::tf32): 4 loader warps convert and store the next column block while the tensor core multiplies
//     the current one into one of two TMEM accumulators and 4 epilogue warps drain the other.
//
// Common to both:
//   * split operands: each element is hi + lo with 11-bit significands and the tile accumulates
//     hi.hi + hi.lo + lo.hi (+ lo.lo in the TF32 kernel), which restores float32-grade dot products; the
//     caller centres the cluster first, so ||x||^2 stays small against the distances.
//   * operands sit in the canonical no-swizzle K-major UMMA layout (8-row x 16-byte core matrices).
//   * upper triangle only: off-diagonal tiles count twice; the diagonal is excluded explicitly.
//   * epilogue: ni + nj - 2 dot, clamp, sqrt.approx, float32 two-sum per thread, float64 across threads;
//     mbarriers carry the stage-free / accumulator-full / accumulator-empty hand-offs, all waits bounded.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

// Per-role cycle counters (clock64 around every mbarrier wait of CTA 0) exist only in the benchmark build:
//   NVCC_EXTRA=-DDIC_TC_PROFILE python -m deep_interpolation_clustering_b200.build --force   (benchmarks/README.md)
// The product library carries none of it: no counters in the pipeline loops, no allocation, sync or output in a call.
#ifdef DIC_TC_PROFILE
#define DIC_PROF(...) __VA_ARGS__
#else
#define DIC_PROF(...)
#endif

namespace dic {
namespace {

using namespace tc;

constexpr int kTile = 128;           // tile rows (i) and columns (j)
constexpr int kKC = 64;              // K chunk (float32 elements) held in shared memory at once
constexpr int kThreads = 256;
constexpr int kTileBytes = kTile * kKC * 4;          // 32 KB per operand tile
constexpr int kChunkStride = kTile * 16;             // bytes between consecutive 16-byte K chunks (LBO)
constexpr int kGroupStride = 128;                    // bytes between 8-row groups (SBO)
constexpr uint32_t kTmemCols = 256;       // two 128-column accumulators

struct __align__(16) TcSmem {
  // operand tiles first (16-byte aligned): the row block A (kept across a tile row) and a
  // double-buffered column block B; then the pipeline state
  unsigned char a_hi[kTileBytes], a_lo[kTileBytes];
  unsigned char b_hi[2][kTileBytes], b_lo[2][kTileBytes];
  float nj[kTile];        // column norms of the tile in the epilogue (owned by the epilogue warps)
  uint64_t se[2];         // smem stage empty  (tensor core -> loaders,  tcgen05.commit)
  uint64_t tf[2];         // TMEM accumulator full (tensor core -> epilogue, tcgen05.commit)
  uint64_t te[2];         // TMEM accumulator empty (epilogue -> MMA issuer, 128 arrivals)
  uint32_t tmem_base;
  int timeout;
};

// Shared-memory matrix descriptor, no swizzle, K-major (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start >> 4 | [16,30) leading-dim byte offset >> 4 (between the two 16-byte K chunks of
//   one MMA) | [32,46) stride-dim byte offset >> 4 (between 8-row groups) | [46,48) version = 1.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((kChunkStride >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((kGroupStride >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6), a/b format TF32 (2)
// at [7,10)/[10,13), both K-major, N >> 3 at [17,23), M >> 4 at [24,29).
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTile >> 3) << 17) |
                            ((uint32_t)(kTile >> 4) << 24);

// The same with a/b format F16 (0): kind::f16, K = 16 per instruction.
constexpr uint32_t kIdescF16 = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(kTile >> 3) << 17) |
                               ((uint32_t)(kTile >> 4) << 24);

__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// rows r0.. of X (n x D), K elements [k0, k0 + 64) -> hi / lo operand tiles in the UMMA layout, in
// two halves so that the global loads of the NEXT tile are in flight while the current one is
// converted and multiplied.  128 loader threads, 16 float4 each: a warp step covers 8 rows x 4
// chunks, so every touched 32-byte sector is fully used and the four 128-byte groups of the
// shared store are contiguous (4 wavefronts, the minimum for 512 bytes).
struct TileRegs {
  float4 v[16];
};
__device__ __forceinline__ void tile_fetch(TileRegs& t, const float* __restrict__ X, int64_t n, int D, int64_t r0,
                                           int k0, int lt) {
  const int lane = lt & 31, w = lt >> 5;
  const int rl = lane & 7, cl = lane >> 3;
  const bool interior = (r0 + kTile <= n) && (k0 + kKC <= D);
  const float* base = X + (r0 + rl) * D + k0 + 4 * cl;
  if (interior) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int step = w + 4 * u;            // step = rg * 4 + cq
      t.v[u] = __ldg(reinterpret_cast<const float4*>(base + (int64_t)(step >> 2) * 8 * D + (step & 3) * 16));
    }
  } else {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int step = w + 4 * u;
      const int64_t gr = r0 + (step >> 2) * 8 + rl;
      const int k = k0 + (step & 3) * 16 + 4 * cl;
      t.v[u] = (gr < n && k < D)
                   ? __ldg(reinterpret_cast<const float4*>(base + (int64_t)(step >> 2) * 8 * D + (step & 3) * 16))
                   : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}
__device__ __forceinline__ void tile_store(const TileRegs& t, unsigned char* hi, unsigned char* lo, int lt) {
  const int lane = lt & 31, w = lt >> 5;
  const int rl = lane & 7, cl = lane >> 3;
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int step = w + 4 * u;
    const int rg = step >> 2, c = (step & 3) * 4 + cl;
    float4 h, l;     // hi = rn_tf32(x), lo = rn_tf32(x - hi): round-to-nearest keeps the split unbiased
    h.x = to_tf32(t.v[u].x); h.y = to_tf32(t.v[u].y); h.z = to_tf32(t.v[u].z); h.w = to_tf32(t.v[u].w);
    l.x = to_tf32(t.v[u].x - h.x); l.y = to_tf32(t.v[u].y - h.y);
    l.z = to_tf32(t.v[u].z - h.z); l.w = to_tf32(t.v[u].w - h.w);
    const int off = c * kChunkStride + rg * kGroupStride + rl * 16;
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }
}

__global__ void row_norms_kernel(const float* __restrict__ X, float* __restrict__ norms, int64_t n, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = X[row * D + d];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (lane == 0) norms[row] = s;
}

__device__ __forceinline__ void unrank_tile(int64_t t, int64_t nb, int64_t& bi, int64_t& bj) {
  // row-major enumeration of the upper triangle (bi <= bj)
  bi = (int64_t)((2.0 * nb + 1.0 - sqrt((2.0 * nb + 1.0) * (2.0 * nb + 1.0) - 8.0 * (double)t)) * 0.5);
  while (bi * nb - bi * (bi - 1) / 2 > t) --bi;
  while ((bi + 1) * nb - (bi + 1) * bi / 2 <= t) ++bi;
  bj = bi + (t - (bi * nb - bi * (bi - 1) / 2));
}

// Warp-specialised pipeline (one persistent CTA per SM, 8 warps):
//   warps 4-7  loaders: stage the column block of tile i+1 (split into hi/lo) while the tensor core
//              works on tile i; thread 128 issues the MMAs of a tile and commits them to two
//              mbarriers (stage free, accumulator full)
//   warps 0-3  epilogue: tcgen05.ld the accumulator of tile i (TMEM lanes 32 w .. 32 w + 31), form
//              the distances and sum them while tile i+1 is being multiplied into the other
//              accumulator
__global__ void __launch_bounds__(kThreads, 1)
pairwise_tc_kernel(const float* __restrict__ X, const float* __restrict__ norms, double* __restrict__ partial,
                   int64_t n, int D, DIC_PROF(long long* __restrict__ dbg,) int part, int n_parts) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  TcSmem& S = *reinterpret_cast<TcSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int k = 0; k < 2; ++k) {
      mbar_init(&S.se[k], 1);
      mbar_init(&S.tf[k], 1);
      mbar_init(&S.te[k], kTile);
    }
    S.timeout = 0;
    fence_proxy_async();
  }
  if (warp == 0) tmem_alloc(&S.tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;

  const int64_t nb = (n + kTile - 1) / kTile;
  const int64_t ntiles = nb * (nb + 1) / 2;
  const int kchunks = (D + kKC - 1) / kKC;
  double total = 0.0;
  // stripe `part` of `n_parts` (multi-GPU): this CTA acts as CTA vcta of a grid of vgrid CTAs
  const int64_t vcta = (int64_t)blockIdx.x * n_parts + part, vgrid = (int64_t)gridDim.x * n_parts;

  if (warp >= 4) {
    // ================= loaders + MMA issuer =================
    const int lt = tid - 128;
    const uint32_t a_hi = smem_u32(S.a_hi), a_lo = smem_u32(S.a_lo);
    const int64_t nstages = vcta < ntiles ? ((ntiles - vcta + vgrid - 1) / vgrid) * kchunks : 0;   // this CTA's stages
    int64_t cur_bi = -1;
    TileRegs breg;                 // the column block of the NEXT stage, already on its way from L2
    if (nstages > 0) {
      int64_t bi, bj;
      unrank_tile(vcta, nb, bi, bj);
      tile_fetch(breg, X, n, D, bj * kTile, 0, lt);
    }
    for (int64_t stage = 0; stage < nstages; ++stage) {
      const int64_t i = stage / kchunks;                       // running tile count of this CTA
      const int kc = (int)(stage - i * kchunks);
      const int64_t t = vcta + i * vgrid;
      int64_t bi, bj;
      unrank_tile(t, nb, bi, bj);
      const int acc = (int)(i & 1), b = (int)(stage & 1);
      const int u = (int)(stage >> 1);
      DIC_PROF(long long c0 = clock64();)
      if (!bar_wait_bounded(&S.se[b], (uint32_t)((u & 1) ^ 1))) S.timeout = 1;   // stage b free again
      DIC_PROF(long long c1 = clock64();)
      if (kchunks > 1 || bi != cur_bi) {
        // the row block is about to change: the MMAs of the previous stage must be done with it
        if (stage > 0 && !bar_wait_bounded(&S.se[b ^ 1], (uint32_t)(((stage - 1) >> 1) & 1))) S.timeout = 1;
        TileRegs areg;
        tile_fetch(areg, X, n, D, bi * kTile, kc * kKC, lt);
        tile_store(areg, S.a_hi, S.a_lo, lt);
        cur_bi = bi;
      }
      DIC_PROF(long long c2 = clock64();)
      tile_store(breg, S.b_hi[b], S.b_lo[b], lt);
      DIC_PROF(long long c3 = clock64();)
      if (stage + 1 < nstages) {                               // prefetch the next stage's column block
        const int64_t i2 = (stage + 1) / kchunks;
        const int kc2 = (int)(stage + 1 - i2 * kchunks);
        int64_t bi2, bj2;
        unrank_tile(vcta + i2 * vgrid, nb, bi2, bj2);
        tile_fetch(breg, X, n, D, bj2 * kTile, kc2 * kKC, lt);
      }
      DIC_PROF(long long c4 = clock64();)
      fence_proxy_async();         // generic-proxy stores -> visible to the tensor core (async proxy)
      named_bar_sync(1, 128);
      DIC_PROF(long long c5 = clock64();)
#ifdef DIC_TC_PROFILE
      if (dbg && lt == 32 && blockIdx.x == 0) {
        atomicAdd((unsigned long long*)&dbg[0], (unsigned long long)(c1 - c0));   // wait stage free
        atomicAdd((unsigned long long*)&dbg[1], (unsigned long long)(c2 - c1));   // row block reload
        atomicAdd((unsigned long long*)&dbg[2], (unsigned long long)(c3 - c2));   // convert + store
        atomicAdd((unsigned long long*)&dbg[3], (unsigned long long)(c4 - c3));   // prefetch issue
        atomicAdd((unsigned long long*)&dbg[4], (unsigned long long)(c5 - c4));   // loader barrier
        atomicAdd((unsigned long long*)&dbg[5], 1ull);
      }
#endif
      if (lt == 0) {
        DIC_PROF(long long d0 = clock64();)
        if (kc == 0) {             // accumulator drained by the epilogue of tile i - 2
          if (!bar_wait_bounded(&S.te[acc], (uint32_t)(((i >> 1) & 1) ^ 1))) S.timeout = 1;
        }
#ifdef DIC_TC_PROFILE
        if (dbg && blockIdx.x == 0) atomicAdd((unsigned long long*)&dbg[6], (unsigned long long)(clock64() - d0));
#endif
        tc_fence_after();
        const uint32_t d_tmem = tmem + (uint32_t)acc * kTile;
        const uint32_t b_hi = smem_u32(S.b_hi[b]), b_lo = smem_u32(S.b_lo[b]);
#pragma unroll
        for (int s = 0; s < kKC / 8; ++s) {        // one MMA consumes K = 8 tf32 = two 16-byte chunks
          const uint32_t koff = (uint32_t)s * 2u * kChunkStride;
          const uint64_t dah = make_desc(a_hi + koff), dal = make_desc(a_lo + koff);
          const uint64_t dbh = make_desc(b_hi + koff), dbl = make_desc(b_lo + koff);
          umma_tf32(d_tmem, dah, dbh, kIdesc, (kc > 0 || s > 0) ? 1u : 0u);
          umma_tf32(d_tmem, dah, dbl, kIdesc, 1u);
          umma_tf32(d_tmem, dal, dbh, kIdesc, 1u);
          umma_tf32(d_tmem, dal, dbl, kIdesc, 1u);
        }
        umma_commit(&S.se[b]);                      // stage b reusable once these MMAs have read it
        if (kc + 1 == kchunks) umma_commit(&S.tf[acc]);   // accumulator complete -> epilogue
#ifdef DIC_TC_PROFILE
        if (dbg && blockIdx.x == 0) atomicAdd((unsigned long long*)&dbg[7], (unsigned long long)(clock64() - d0));
#endif
      }
    }
  } else {
    // ================= epilogue =================
    int i = 0;
    float acc_hi = 0.f, acc_lo = 0.f;
    for (int64_t t = vcta; t < ntiles; t += vgrid, ++i) {
      int64_t bi, bj;
      unrank_tile(t, nb, bi, bj);
      const int acc = i & 1;
      const int64_t i0 = bi * kTile, j0 = bj * kTile;
      named_bar_sync(2, 128);                       // previous tile's reads of nj[] are finished
      S.nj[tid] = (j0 + tid < n) ? __ldg(norms + j0 + tid) : 0.f;
      const int64_t gi = i0 + tid;                  // TMEM lane = tile row = thread
      const float ni = gi < n ? __ldg(norms + gi) : 0.f;
      named_bar_sync(2, 128);
      DIC_PROF(long long e0 = clock64();)
      if (!bar_wait_bounded(&S.tf[acc], (uint32_t)((i >> 1) & 1))) S.timeout = 1;
      DIC_PROF(long long e1 = clock64();)
      tc_fence_after();
      float tile_sum = 0.f;
      const bool plain = (bi != bj) && (i0 + kTile <= n) && (j0 + kTile <= n);   // no diagonal, no ragged edge
#pragma unroll 1
      for (int h = 0; h < 4; ++h) {
        const int col0 = 32 * h;
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(acc * kTile + col0), v);
        if (plain) {
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const float4 nj4 = *reinterpret_cast<const float4*>(&S.nj[col0 + c]);
            tile_sum += sqrt_approx(fmaxf(fmaf(-2.f, __uint_as_float(v[c + 0]), ni + nj4.x), 0.f));
            tile_sum += sqrt_approx(fmaxf(fmaf(-2.f, __uint_as_float(v[c + 1]), ni + nj4.y), 0.f));
            tile_sum += sqrt_approx(fmaxf(fmaf(-2.f, __uint_as_float(v[c + 2]), ni + nj4.z), 0.f));
            tile_sum += sqrt_approx(fmaxf(fmaf(-2.f, __uint_as_float(v[c + 3]), ni + nj4.w), 0.f));
          }
        } else if (gi < n) {
          const int jmax = (int)min((int64_t)32, n - j0 - col0);       // valid columns in this chunk
          const int jdiag = (int)(gi - j0 - col0);                      // the diagonal, if inside
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float d2 = fmaf(-2.f, __uint_as_float(v[c]), ni + S.nj[col0 + c]);
            const float d = sqrt_approx(fmaxf(d2, 0.f));
            tile_sum += (c < jmax && c != jdiag) ? d : 0.f;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&S.te[acc]);                      // this thread is done with accumulator `acc`
#ifdef DIC_TC_PROFILE
      if (dbg && tid == 0 && blockIdx.x == 0) {
        atomicAdd((unsigned long long*)&dbg[8], (unsigned long long)(e1 - e0));          // wait accumulator
        atomicAdd((unsigned long long*)&dbg[9], (unsigned long long)(clock64() - e1));   // drain + distances
      }
#endif
      {                                             // float32 two-sum: no float64 arithmetic per tile
        const float v = (bi == bj) ? tile_sum : 2.f * tile_sum;
        const float sum = acc_hi + v;
        const float bb = sum - acc_hi;
        acc_lo += (acc_hi - (sum - bb)) + (v - bb);
        acc_hi = sum;
      }
    }
    total = (double)acc_hi + (double)acc_lo;
  }

  __shared__ double red[kThreads / 32];
  total = warp_sum(total);
  if (lane == 0) red[warp] = total;
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) s += red[w];
    partial[blockIdx.x] = S.timeout ? __longlong_as_double(0x7ff8000000000000LL) : s;   // NaN = pipeline stalled
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, kTmemCols);
  }
}

__global__ void sum_partials_kernel(const double* __restrict__ ws, double* __restrict__ out, int nblocks) {
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += ws[b];
  s = warp_sum(s);
  if (lane == 0) *out = s;
}


// =================================================================================================
// D <= 64: TMA-fed, A-stationary variant.
//
// A one-off pre-pass (absmax_kernel + pack_split_kernel) writes every 128-row block of X ONCE as split
// float16 operand tiles (x scaled by a power of two so that max |x| lies in [2^13, 2^14); hi = rn_f16,
// lo = rn_f16(x - hi): the same 2 x 11 significand bits as split TF32 at half the bytes and twice the K per
// MMA) in the exact shared-memory image the tensor core reads (hi tile | lo tile, 32 KB per block, K padded
// to 64) plus the row norms.  The main kernel then never touches a float again on its way to the tensor
// core: one elected thread streams tiles with 1-D TMA bulk copies, one thread issues tcgen05.mma
// (kind::f16), eight warps drain TMEM.
//
//   * supertile = 256 rows (two 128-row blocks: the "row pair", resident in shared memory for a
//     whole work item, 64 KB) x 128 columns; every column tile that arrives (32 KB) is multiplied
//     against BOTH row blocks, which halves the L2 -> SM traffic per distance;
//   * column tiles arrive whole (32 KB stages, 4 in flight): 64 + 128 KB of shared memory;
//   * TMEM holds 2 (double buffer) x 2 (row blocks) accumulators of 128 columns = all 512 columns;
//   * work items = (row pair p, segment of <= L column tiles), enumerated segment-major so that the
//     CTAs running at the same time read the same column tiles out of L2, dealt cyclically.
// =================================================================================================
namespace tc64 {

constexpr int kBlk = 128;
constexpr int kTileB = kBlk * 64 * 2;        // 16 KB: one hi or lo operand tile (128 rows x 64 halves of K)
constexpr int kChunkB = 2 * kTileB;          // 32 KB: hi | lo of one 64-wide K chunk of a 128-row block
constexpr int kKSteps = 4;                   // MMAs (K = 16 halves = two 16-byte chunks) per 64-wide K chunk
constexpr int kEpiWarps = 8;
constexpr int kThreads64 = 32 * (2 + kEpiWarps);

// NK = 64-wide K chunks per row (D <= 64 NK), NH = 128-row blocks resident per work item.  <1,2>: the layout
// described above.  <2,2> (D <= 128): 128 KB resident + 3 stages.  <4,1> (D <= 256): ONE row block resident
// (128 KB), 3 stages, the eight epilogue warps split the 128 columns of its accumulator in halves.
template <int NK, int NH>
struct Cfg {
  static constexpr int kStages = NK == 1 ? 4 : 3;          // 32 KB stages: one K chunk of a column tile each (5 measured: no gain)
  static constexpr int kBlockB = NK * kChunkB;             // bytes of one packed 128-row block
};

template <int NK, int NH>
struct __align__(128) Smem64 {
  static constexpr int kStages = Cfg<NK, NH>::kStages;
  unsigned char a[NH][NK][2][kTileB];        // [row block of the item][K chunk][hi | lo]
  unsigned char b[kStages][2][kTileB];       // [stage][hi | lo], one K chunk of a column tile
  uint64_t a_full, a_empty, b_full[kStages], b_empty[kStages], acc_full[2], acc_empty[2];
  double red[kEpiWarps];
  uint32_t tmem_base;
  int timeout;
};

// The sequence of work items of CTA `cta`: segment s covers column tiles [sL, (s+1)L); row group p (nh row
// blocks) takes part in it when nh p < (s+1)L.  Item (s, p) has global index w(s) + p and belongs to CTA
// (w(s) + p) % G.
struct ItemIter {
  int64_t nb, P, S, s, w, p;
  int L, G, cta, nh;
  bool full;     // true: the whole nb x nb grid of tiles (row sums); false: the upper triangle (symmetric sum)
  __device__ void init(int64_t nb_, int L_, int G_, int cta_, bool full_, int nh_) {
    nb = nb_; L = L_; G = G_; cta = cta_; full = full_; nh = nh_;
    P = (nb + nh - 1) / nh;
    S = (nb + L - 1) / L;
    s = 0; w = 0;
    p = first();
  }
  __device__ int64_t first() const { return (int64_t)((((cta - w) % G) + G) % G); }
  __device__ int64_t pmax() const { return full ? P - 1 : min(P - 1, ((s + 1) * L - 1) / nh); }
  __device__ bool next(int64_t& op, int64_t& j0, int64_t& j1) {
    while (s < S) {
      if (p <= pmax()) {
        op = p;
        j0 = full ? s * L : max(s * L, (int64_t)nh * p);
        j1 = min((s + 1) * L, nb);
        p += G;
        return true;
      }
      w += pmax() + 1;
      ++s;
      p = first();
    }
    return false;
  }
};

// Power-of-two scale that puts max |x| into [2^13, 2^14): the two-term float16 split below then keeps 22 bits of
// every element whose magnitude is within 2^-11 of the largest (smaller ones lose nothing that matters: the error
// of a dot product is relative to |x||y|).  absmax_bits = bit pattern of max |x| (non-negative floats order like
// unsigned integers).
__device__ __forceinline__ float f16_scale(unsigned bits, float* inv) {
  int e = bits ? (int)((bits >> 23) & 0xffu) - 127 : 13;       // floor(log2 max|x|); all-zero data: scale 1
  e = max(-100, min(100, e));
  *inv = __uint_as_float((uint32_t)(e - 13 + 127) << 23);
  return __uint_as_float((uint32_t)(13 - e + 127) << 23);
}

// thread -> (row, 16-byte chunk of 8 halves) of a 128-row block, shared by the two pre-pass kernels
__device__ __forceinline__ void load_row_chunk(const float* __restrict__ X, int64_t src, int D, int c, float (&v)[8]) {
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  if (src >= 0 && 8 * c < D) a = __ldg(reinterpret_cast<const float4*>(X + src * D + 8 * c));
  if (src >= 0 && 8 * c + 4 < D) b = __ldg(reinterpret_cast<const float4*>(X + src * D + 8 * c + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

template <int NK>
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ X, unsigned* __restrict__ absmax_bits, int64_t n, int D,
              const int32_t* __restrict__ perm) {
  const int64_t blk = blockIdx.x;
  float m = 0.f;
#pragma unroll
  for (int it = 0; it < 4 * NK; ++it) {
    const int idx = it * 256 + threadIdx.x;
    const int row = idx / (8 * NK), c = idx % (8 * NK);
    const int64_t gr = blk * kBlk + row;
    const int64_t src = gr < n ? (perm ? (int64_t)__ldg(perm + gr) : gr) : -1;
    float v[8];
    load_row_chunk(X, src, D, c, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) m = fmaxf(m, fabsf(v[k]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(absmax_bits, __float_as_uint(m));
}

// X (n x D, D <= 64 NK, D % 4 == 0) -> packed split-float16 operand tiles (hi = rn_f16(s x), lo = rn_f16(s x - hi);
// per 128-row block NK images [hi | lo] of 64-wide K chunks) + row norms of s x; rows >= n and K >= D are zero.
// One CTA per 128-row block, thread -> (row, 16-byte K chunk); the 8 NK chunks of a row sit in adjacent lanes.
template <int NK>
__global__ void __launch_bounds__(256)
pack_split_kernel(const float* __restrict__ X, unsigned char* __restrict__ packed, float* __restrict__ norms,
                  int64_t n, int D, const int32_t* __restrict__ perm, float pad_norm,
                  const unsigned* __restrict__ absmax_bits) {
  const int64_t blk = blockIdx.x;
  unsigned char* out = packed + blk * (int64_t)(NK * kChunkB);
  float inv;
  const float scale = f16_scale(__ldg(absmax_bits), &inv);
#pragma unroll
  for (int it = 0; it < 4 * NK; ++it) {
    const int idx = it * 256 + threadIdx.x;
    const int row = idx / (8 * NK), c = idx % (8 * NK);
    const int64_t gr = blk * kBlk + row;
    // perm (row-sums mode): packed row gr holds X[perm[gr]], perm < 0 = padding between clusters
    const int64_t src = gr < n ? (perm ? (int64_t)__ldg(perm + gr) : gr) : -1;
    float v[8];
    load_row_chunk(X, src, D, c, v);
    __half2 h[4], l[4];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float x0 = v[2 * k] * scale, x1 = v[2 * k + 1] * scale;          // power of two: exact
      const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
      h[k] = __halves2half2(h0, h1);
      l[k] = __halves2half2(__float2half_rn(x0 - __half2float(h0)), __float2half_rn(x1 - __half2float(h1)));
      s = fmaf(x0, x0, fmaf(x1, x1, s));
    }
    const int off = (c >> 3) * kChunkB + (c & 7) * kChunkStride + (row >> 3) * kGroupStride + (row & 7) * 16;
    *reinterpret_cast<uint4*>(out + off) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(out + kTileB + off) = *reinterpret_cast<const uint4*>(l);
#pragma unroll
    for (int o = 4 * NK; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    // a padding row gets pad_norm: -inf makes every distance to it exactly 0 (max(-inf, 0) under the sqrt)
    if (c == 0) norms[gr] = src >= 0 ? s : pad_norm;
  }
}

// ROWSUMS = false: partial[cta] = this CTA's share of sum_{i,j} ||x_i - x_j|| (upper triangle, doubled).
// ROWSUMS = true : rows are sorted by cluster and every cluster is padded to whole 128-row tiles (padding
//                  rows carry norm = -inf => distance 0), tile_cluster[bj] names the cluster of column tile
//                  bj, and rowsum[i][k] += sum_{j in tile, cluster k} ||x_i - x_j|| over the FULL grid: what
//                  the silhouette needs (sklearn.metrics.silhouette_samples) without the n x n matrix.
template <bool ROWSUMS, int NK, int NH>
__global__ void __launch_bounds__(kThreads64, 1)
pairwise_tc64_kernel(const unsigned char* __restrict__ packed, const float* __restrict__ norms,
                     double* __restrict__ partial, int64_t n, int L, DIC_PROF(long long* __restrict__ dbg,)
                     const int32_t* __restrict__ tile_cluster, double* __restrict__ rowsum, int K, int part,
                     int n_parts, const unsigned* __restrict__ absmax_bits) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  using SmemT = Smem64<NK, NH>;
  constexpr int kStages = SmemT::kStages;
  constexpr int kBlockB = NK * kChunkB;
  SmemT& S = *reinterpret_cast<SmemT*>(smem_raw);
  // the warp index through a shuffle: the compiler then knows that the role branches below are warp-uniform
  // (cutlass::canonical_warp_idx_sync), which is what lets the MMA operands live in uniform registers
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int64_t nb = (n + kBlk - 1) / kBlk;

  if (tid == 0) {
    mbar_init(&S.a_full, 1);
    mbar_init(&S.a_empty, 1);
    for (int k = 0; k < kStages; ++k) {
      mbar_init(&S.b_full[k], 1);
      mbar_init(&S.b_empty[k], 1);
    }
    for (int k = 0; k < 2; ++k) {
      mbar_init(&S.acc_full[k], 1);
      mbar_init(&S.acc_empty[k], kEpiWarps);
    }
    S.timeout = 0;
    fence_proxy_async();
  }
  if (warp == 2) tmem_alloc(&S.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  volatile int* timeout = &S.timeout;
  float inv_scale;                        // distances come out in units of the packed (scaled) data
  f16_scale(__ldg(absmax_bits), &inv_scale);

  ItemIter it;
  // stripe `part` of `n_parts` (multi-GPU): this CTA takes the items of CTA blockIdx.x * n_parts + part of a grid
  // n_parts times as large, so the stripes of all parts tile the item list exactly once
  it.init(nb, L, (int)gridDim.x * n_parts, (int)blockIdx.x * n_parts + part, ROWSUMS, NH);
  int64_t p, j0, j1;
  double total = 0.0;

  if (warp == 0) {
    // ================= producer: TMA bulk copies =================
    if (lane == 0) {
      int64_t item = 0, g = 0;
      DIC_PROF(long long pw_a = 0, pw_b = 0;)
      while (it.next(p, j0, j1) && !*timeout) {
        DIC_PROF(long long c0 = clock64();)
        if (!bar_wait_bounded(&S.a_empty, (uint32_t)((item & 1) ^ 1))) { *timeout = 1; break; }
        DIC_PROF(pw_a += clock64() - c0;)
        mbar_expect_tx(&S.a_full, (uint32_t)(NH * kBlockB));
        const unsigned char* arow = packed + NH * p * (int64_t)kBlockB;    // the NH blocks of the item are adjacent
#pragma unroll
        for (int q = 0; q < 2 * NH * NK; ++q)
          bulk_g2s(&S.a[0][0][0][0] + q * kTileB, arow + q * kTileB, kTileB, &S.a_full);
        for (int64_t bj = j0; bj < j1 && !*timeout; ++bj) {
          const unsigned char* bcol = packed + bj * (int64_t)kBlockB;
          for (int kc = 0; kc < NK; ++kc, ++g) {
            const int slot = (int)(g % kStages);
            DIC_PROF(c0 = clock64();)
            if (!bar_wait_bounded(&S.b_empty[slot], (uint32_t)(((g / kStages) & 1) ^ 1))) { *timeout = 1; break; }
            DIC_PROF(pw_b += clock64() - c0;)
            mbar_expect_tx(&S.b_full[slot], (uint32_t)kChunkB);
            bulk_g2s(S.b[slot][0], bcol + kc * kChunkB, kTileB, &S.b_full[slot]);
            bulk_g2s(S.b[slot][1], bcol + kc * kChunkB + kTileB, kTileB, &S.b_full[slot]);
          }
        }
        ++item;
      }
#ifdef DIC_TC_PROFILE
      if (dbg && blockIdx.x == 0) { dbg[10] = pw_a; dbg[11] = pw_b; }
#endif
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The WHOLE warp runs the loop in uniform control flow and only the tcgen05 instructions themselves sit under
    // `if (elect_one())`: descriptors and addresses are then computed on the uniform datapath and the three MMAs
    // of a K step issue back to back.  (With the loop inside `if (lane == 0)` every UTCHMMA operand went through
    // an ELECT / R2UR.BROADCAST waterfall: 96 clk per MMA issued against ~53 clk of tensor-pipe time.)
    {
      DIC_PROF(const bool leader = lane == 0;)
      // redux.sync results are uniform by construction: the compiler cannot know that a shared-memory load is
      const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem);
      auto stalled = [&]() { return __reduce_or_sync(0xffffffffu, (unsigned)*timeout) != 0u; };
      int64_t item = 0, g = 0, t = 0;
      DIC_PROF(long long w_a = 0, w_acc = 0, w_b = 0;)
      DIC_PROF(const long long m0 = clock64();)
      const uint32_t a_base = smem_u32(&S.a[0][0][0][0]);
      while (it.next(p, j0, j1) && !stalled()) {
        DIC_PROF(long long c0 = clock64();)
        if (!__all_sync(0xffffffffu, bar_wait_bounded(&S.a_full, (uint32_t)(item & 1)))) { *timeout = 1; break; }
        DIC_PROF(w_a += clock64() - c0;)
        for (int64_t bj = j0; bj < j1 && !stalled(); ++bj, ++t) {
          const int buf = (int)__reduce_or_sync(0xffffffffu, (unsigned)(t & 1));          // uniform (see tmem_u)
          DIC_PROF(c0 = clock64();)
          if (!__all_sync(0xffffffffu, bar_wait_bounded(&S.acc_empty[buf], (uint32_t)(((t >> 1) & 1) ^ 1)))) {
            *timeout = 1;
            break;
          }
          DIC_PROF(w_acc += clock64() - c0;)
          for (int kc = 0; kc < NK; ++kc, ++g) {
            const int slot = (int)__reduce_or_sync(0xffffffffu, (unsigned)(g % kStages));   // uniform
            DIC_PROF(c0 = clock64();)
            if (!__all_sync(0xffffffffu, bar_wait_bounded(&S.b_full[slot], (uint32_t)((g / kStages) & 1)))) {
              *timeout = 1;
              break;
            }
            DIC_PROF(w_b += clock64() - c0;)
            tc_fence_after();
            const uint32_t b_hi = smem_u32(S.b[slot][0]), b_lo = smem_u32(S.b[slot][1]);
#pragma unroll
            for (int h = 0; h < NH; ++h) {
              const uint32_t d_tmem = tmem_u + (uint32_t)((buf * NH + h) * kBlk);
              const uint32_t a_hi = a_base + (uint32_t)(h * kBlockB) + (uint32_t)kc * (uint32_t)kChunkB;
              const uint32_t a_lo = a_hi + kTileB;
#pragma unroll
              for (int ks = 0; ks < kKSteps; ++ks) {      // one MMA consumes K = 16 halves = two 16-byte chunks
                const uint32_t kk = (uint32_t)(ks * 2 * kChunkStride);
                const uint64_t dah = make_desc(a_hi + kk), dal = make_desc(a_lo + kk);
                const uint64_t dbh = make_desc(b_hi + kk), dbl = make_desc(b_lo + kk);
                // split float16 (same 11-bit significands as TF32, twice the K per instruction and half the
                // operand bytes): hi.hi + hi.lo + lo.hi; the dropped lo.lo term is < 2^-22 |x||y|, below the
                // float32 rounding of the norms it is added to
                if (elect_one()) {
                  umma_f16(d_tmem, dah, dbh, kIdescF16, (kc > 0 || ks > 0) ? 1u : 0u);
                  umma_f16(d_tmem, dah, dbl, kIdescF16, 1u);
                  umma_f16(d_tmem, dal, dbh, kIdescF16, 1u);
                }
              }
            }
            if (elect_one()) umma_commit(&S.b_empty[slot]);   // stage reusable once these MMAs have read it
          }
          if (elect_one()) umma_commit(&S.acc_full[buf]);     // the NH accumulators of the supertile are complete
        }
        if (elect_one()) umma_commit(&S.a_empty);             // the resident row blocks may be replaced
        ++item;
      }
#ifdef DIC_TC_PROFILE
      if (dbg && blockIdx.x == 0 && leader) {
        dbg[0] = clock64() - m0; dbg[1] = w_a; dbg[2] = w_acc; dbg[3] = w_b; dbg[4] = t; dbg[5] = item;
      }
#endif
    }
  } else {
    // ================= epilogue: 8 warps, (row block h | column half, TMEM lane quarter q) =================
    // NH = 2: warps 0-3 drain row block 0, warps 4-7 row block 1, all 128 columns (CC = 4 chunks of 32).
    // NH = 1: both sets drain the one row block, columns [0, 64) and [64, 128) (CC = 2 chunks from chunk cbeg).
    constexpr int CC = NH == 2 ? 4 : 2;
    const int ew = warp - 2, q = warp & 3;
    const int h = NH == 2 ? ew >> 2 : 0, cbeg = NH == 2 ? 0 : 2 * (ew >> 2);
    int64_t t = 0;
    bool dead = false;
    DIC_PROF(long long w_full = 0, w_work = 0;)
    DIC_PROF(const long long ep0 = clock64();)
    while (!dead && it.next(p, j0, j1)) {
      const int64_t bi = NH * p + h;
      const int64_t i0 = bi * kBlk, gi = i0 + 32 * q + lane;
      const float ni = gi < n ? __ldg(norms + gi) : 0.f;
      float4 njr[8];                       // norms of the next 32 columns to be processed, prefetched
#pragma unroll
      for (int k = 0; k < 8; ++k) njr[k] = __ldg(reinterpret_cast<const float4*>(norms + j0 * kBlk + 32 * cbeg) + k);
      int kcur = -1;                       // ROWSUMS: cluster of the column tiles summed into racc so far
      // Tile sums are collected in a float32 two-sum (value + rounding error, exact to ~2^-48) and turned into
      // float64 once per item / cluster run: a DADD per tile cost 16 % of this loop's stall samples (ncu).
      float acc_hi = 0.f, acc_lo = 0.f;
      auto acc_add = [&](float v) {
        const float sum = acc_hi + v;
        const float bb = sum - acc_hi;
        acc_lo += (acc_hi - (sum - bb)) + (v - bb);
        acc_hi = sum;
      };
      for (int64_t bj = j0; bj < j1; ++bj, ++t) {
        const int buf = (int)(t & 1);
        const int64_t c0 = bj * kBlk;
        if (ROWSUMS) {
          const int kc = __ldg(tile_cluster + bj);
          if (kc != kcur) {
            if (kcur >= 0 && gi < n)
              atomicAdd(rowsum + gi * K + kcur, ((double)acc_hi + (double)acc_lo) * (double)inv_scale);
            acc_hi = 0.f;
            acc_lo = 0.f;
            kcur = kc;
          }
        }
        DIC_PROF(const long long e0 = clock64();)
        if (!bar_wait_bounded(&S.acc_full[buf], (uint32_t)((t >> 1) & 1))) { *timeout = 1; dead = true; break; }
        DIC_PROF(const long long e1 = clock64();)
        DIC_PROF(w_full += e1 - e0;)
        tc_fence_after();
        float tile_sum = 0.f;
        // ROWSUMS: every tile counts, padding is neutralised by its -inf norm, only the diagonal needs care
        const bool active = ROWSUMS ? (i0 < n) : ((bj >= bi) && (i0 < n));
        const bool plain = ROWSUMS ? (active && bj != bi)
                                   : (active && (bj > bi) && (i0 + kBlk <= n) && (c0 + kBlk <= n));   // no diagonal / ragged edge
        const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)((buf * NH + h) * kBlk + 32 * cbeg);
        bool released = false;
        if (plain) {
          // Software pipeline over the CC 32-column chunks: the TMEM load of chunk c+1 and the norms of
          // chunk c+1 (the first chunk of the NEXT column tile after the last one) are in flight while
          // chunk c is turned into distances.
          uint32_t va[32], vb[32];
          tmem_ld32_issue(taddr, va);
#pragma unroll
          for (int cc = 0; cc < CC; ++cc) {
            uint32_t(&cur)[32] = (cc & 1) ? vb : va;
            uint32_t(&nxt)[32] = (cc & 1) ? va : vb;
            tmem_ld_wait(cur);
            if (cc < CC - 1) {
              tmem_ld32_issue(taddr + 32 * (cc + 1), nxt);
            } else {                     // everything is in registers: hand the accumulators back early
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&S.acc_empty[buf]);
              released = true;
            }
            float4 njc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) njc[k] = njr[k];
            // the chunk after the last one is this warp's first chunk of tile bj + 1 (norms is padded)
            const float* nxp = norms + c0 + 32 * cbeg + (cc + 1 < CC ? 32 * (cc + 1) : kBlk);
#pragma unroll
            for (int k = 0; k < 8; ++k) njr[k] = __ldg(reinterpret_cast<const float4*>(nxp) + k);
            float d[32];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              d[4 * k + 0] = fmaxf(fmaf(-2.f, __uint_as_float(cur[4 * k + 0]), ni + njc[k].x), 0.f);
              d[4 * k + 1] = fmaxf(fmaf(-2.f, __uint_as_float(cur[4 * k + 1]), ni + njc[k].y), 0.f);
              d[4 * k + 2] = fmaxf(fmaf(-2.f, __uint_as_float(cur[4 * k + 2]), ni + njc[k].z), 0.f);
              d[4 * k + 3] = fmaxf(fmaf(-2.f, __uint_as_float(cur[4 * k + 3]), ni + njc[k].w), 0.f);
            }
#pragma unroll
            // (x * rsqrt(x) instead of sqrt.approx was measured: no change - the SFU is not what bounds this loop)
            for (int c = 0; c < 32; ++c) d[c] = sqrt_approx(d[c]);
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              s0 += d[c];
              s1 += d[c + 1];
              s2 += d[c + 2];
              s3 += d[c + 3];
            }
            tile_sum += (s0 + s1) + (s2 + s3);
          }
        } else {
          if (active) {                    // warp-uniform: tcgen05.ld is .sync.aligned (all 32 lanes take part)
#pragma unroll 1
            for (int cc = 0; cc < CC; ++cc) {
              uint32_t v[32];
              tmem_ld32(taddr + 32 * cc, v);
              const int64_t col0 = c0 + 32 * (cbeg + cc);
              const float* njp = norms + col0;
              const int jmax = gi < n ? (ROWSUMS ? 32 : (int)min((int64_t)32, n - col0)) : 0;   // valid columns
              const int jdiag = (int)(gi - col0);                             // the diagonal, if inside
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                const float d2 = fmaf(-2.f, __uint_as_float(v[c]), ni + __ldg(njp + c));   // norms is padded
                const float d = sqrt_approx(fmaxf(d2, 0.f));
                tile_sum += (c < jmax && c != jdiag) ? d : 0.f;
              }
            }
          }
#pragma unroll
          for (int k = 0; k < 8; ++k)      // keep the prefetch invariant: njr = chunk 0 of the next column tile
            njr[k] = __ldg(reinterpret_cast<const float4*>(norms + c0 + kBlk + 32 * cbeg) + k);
        }
        if (!released) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&S.acc_empty[buf]);       // this warp is done with the accumulators
        }
        DIC_PROF(w_work += clock64() - e1;)
        acc_add((ROWSUMS || bj == bi) ? tile_sum : 2.f * tile_sum);
      }
      if (ROWSUMS) {
        if (kcur >= 0 && gi < n)
          atomicAdd(rowsum + gi * K + kcur, ((double)acc_hi + (double)acc_lo) * (double)inv_scale);
      } else {
        total += (double)acc_hi + (double)acc_lo;
      }
    }
    total = warp_sum(total);
    if (lane == 0) S.red[ew] = total;
#ifdef DIC_TC_PROFILE
    if (dbg && blockIdx.x == 0 && lane == 0 && (ew == 0 || ew == 5)) {
      dbg[6 + 2 * (ew != 0)] = w_full; dbg[7 + 2 * (ew != 0)] = w_work;
      if (ew == 0) { dbg[12] = clock64() - ep0; dbg[13] = t; }
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < kEpiWarps; ++w) s += S.red[w];
    partial[blockIdx.x] = S.timeout ? __longlong_as_double(0x7ff8000000000000LL) : s * (double)inv_scale;   // NaN = stalled
    if (ROWSUMS && S.timeout) rowsum[0] = __longlong_as_double(0x7ff8000000000000LL);
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace tc64

}  // namespace

static int64_t tc64_blocks(int64_t n) { return ((n + tc64::kBlk - 1) / tc64::kBlk + 1) / 2 * 2; }   // even
static int tc64_nk(int D) { return D <= 64 ? 1 : (D <= 128 ? 2 : 4); }   // 64-wide K chunks per row (K zero-padded)
constexpr int kTc64MaxD = 256;

size_t pairwise_tc_workspace_bytes(int64_t n, int D) {
  const size_t v1 = ((size_t)n * sizeof(float) + 255) / 256 * 256 + 1024 * sizeof(double);
  if (D > kTc64MaxD) return v1;
  // packed operand tiles | norms (padded) | partials | scale
  const size_t v2 = (size_t)tc64_blocks(n) * ((size_t)tc64_nk(D) * tc64::kChunkB + tc64::kBlk * sizeof(float)) +
                    tc64::kBlk * sizeof(float) + 1024 * sizeof(double) + 256 + 1024;
  return v2 > v1 ? v2 : v1;
}

// Shared launcher.  rowsum == nullptr: pairwise sum of the n rows of X -> out.  Otherwise: X is read
// through perm (n = padded row count), rowsum (n, K) is zeroed and filled.
template <int NK, int NH>
static int launch_tc64_t(const float* X, double* out, void* workspace, int64_t n, int D, const int32_t* perm,
                         const int32_t* tile_cluster, double* rowsum, int K, cudaStream_t st, int part, int n_parts) {
  using namespace tc64;
  constexpr int kBlockB = NK * kChunkB;
  const bool rows_mode = rowsum != nullptr;
  const int64_t nblk = tc64_blocks(n), nb = (n + kBlk - 1) / kBlk;
  unsigned char* packed = static_cast<unsigned char*>(workspace);          // cudaMalloc alignment (>= 256)
  float* norms = reinterpret_cast<float*>(packed + (size_t)nblk * kBlockB);
  // norms holds one block more than the packed tiles: the epilogue prefetches one column tile ahead
  double* partial = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(norms) +
                                              ((size_t)(nblk + 1) * kBlk * sizeof(float) + 255) / 256 * 256);
  unsigned* absmax = reinterpret_cast<unsigned*>(partial + 1024);
  DIC_CUDA(cudaMemsetAsync(norms + nblk * kBlk, 0, kBlk * sizeof(float), st));
  DIC_CUDA(cudaMemsetAsync(absmax, 0, sizeof(unsigned), st));
  if (rows_mode) DIC_CUDA(cudaMemsetAsync(rowsum, 0, (size_t)n * K * sizeof(double), st));
  const float pad_norm = rows_mode ? -INFINITY : 0.f;
  absmax_kernel<NK><<<(unsigned)nblk, 256, 0, st>>>(X, absmax, n, D, perm);
  DIC_LAUNCH_CHECK("absmax_kernel");
  pack_split_kernel<NK><<<(unsigned)nblk, 256, 0, st>>>(X, packed, norms, n, D, perm, pad_norm, absmax);
  DIC_LAUNCH_CHECK("pack_split_kernel");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t P = (nb + NH - 1) / NH;
  const int64_t supertiles = rows_mode ? P * nb : P * nb - NH * P * (P - 1) / 2;   // full grid | sum_p (nb - NH p)
  int64_t L = supertiles / (8 * (int64_t)sms);
  L = L < 2 ? 2 : (L > 64 ? 64 : L);
  int64_t items = 0;
  for (int64_t s = 0; s * L < nb; ++s)
    items += rows_mode ? P : (((s + 1) * L - 1) / NH < P - 1 ? ((s + 1) * L - 1) / NH + 1 : P);
  const int64_t my_items = items / n_parts > 0 ? items / n_parts : 1;
  int blocks = (int)(my_items < sms ? my_items : sms);
  if (blocks > 1024) blocks = 1024;
  const size_t smem = sizeof(Smem64<NK, NH>) + 1024;
#ifdef DIC_TC_PROFILE      // benchmark build only: per-role cycle counters of CTA 0, printed after the run
  long long* dbg = nullptr;
    cudaMalloc(&dbg, 16 * sizeof(long long));
    cudaMemsetAsync(dbg, 0, 16 * sizeof(long long), st);
#endif
  if (rows_mode) {
    auto kern = pairwise_tc64_kernel<true, NK, NH>;
    DIC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, kThreads64, smem, st>>>(packed, norms, partial, n, (int)L, DIC_PROF(dbg,) tile_cluster, rowsum, K, part,
                                           n_parts, absmax);
  } else {
    auto kern = pairwise_tc64_kernel<false, NK, NH>;
    DIC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, kThreads64, smem, st>>>(packed, norms, partial, n, (int)L, DIC_PROF(dbg,) nullptr, nullptr, 0, part, n_parts,
                                           absmax);
  }
#ifdef DIC_TC_PROFILE
  if (dbg) {
    long long h[16];
    cudaMemcpyAsync(h, dbg, sizeof(h), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    const double T = (double)(h[4] ? h[4] : 1);
    fprintf(stderr, "[tc64 profile <NK=%d,NH=%d>, CTA 0: %lld supertiles, %lld items, L=%lld] MMA thread: total %.0f | "
            "wait A %.0f acc_empty %.0f B %.0f || epilogue warp 0: wait %.0f work %.0f | warp 5: wait %.0f work %.0f || "
            "producer: wait a_empty %.0f b_empty %.0f | epilogue warp 0 loop total %.0f over %lld tiles "
            "(cycles/supertile)\n",
            NK, NH, h[4], h[5], (long long)L, h[0] / T, h[1] / T, h[2] / T, h[3] / T, h[6] / T, h[7] / T, h[8] / T,
            h[9] / T, h[10] / T, h[11] / T, h[12] / T, h[13]);
    cudaFree(dbg);
  }
#endif
  DIC_LAUNCH_CHECK("pairwise_tc64_kernel");
  if (!rows_mode) {
    sum_partials_kernel<<<1, 32, 0, st>>>(partial, out, blocks);
    DIC_LAUNCH_CHECK("sum_partials_kernel");
  }
  return DIC_OK;
}

static int launch_tc64(const float* X, double* out, void* workspace, int64_t n, int D, const int32_t* perm,
                       const int32_t* tile_cluster, double* rowsum, int K, cudaStream_t st, int part = 0,
                       int n_parts = 1) {
  switch (tc64_nk(D)) {
    case 1: return launch_tc64_t<1, 2>(X, out, workspace, n, D, perm, tile_cluster, rowsum, K, st, part, n_parts);
    case 2: return launch_tc64_t<2, 2>(X, out, workspace, n, D, perm, tile_cluster, rowsum, K, st, part, n_parts);
    default: return launch_tc64_t<4, 1>(X, out, workspace, n, D, perm, tile_cluster, rowsum, K, st, part, n_parts);
  }
}

static int launch_pairwise_tc64(const float* X, double* out, void* workspace, int64_t n, int D, cudaStream_t st,
                                int part, int n_parts) {
  return launch_tc64(X, out, workspace, n, D, nullptr, nullptr, nullptr, 0, st, part, n_parts);
}

// Row sums by cluster (silhouette): see pairwise_tc64_kernel<true>.  n_pad % 128 == 0.
int launch_cluster_rowsums_tc(const float* X, const int32_t* perm, const int32_t* tile_cluster, double* rowsum,
                              void* workspace, int64_t n_pad, int D, int K, cudaStream_t st) {
  return launch_tc64(X, nullptr, workspace, n_pad, D, perm, tile_cluster, rowsum, K, st);
}

bool pairwise_tc_supported(const void* X, int D) { return D % 4 == 0 && D >= 4 && aligned16(X); }

int launch_pairwise_tc(const float* X, double* out, void* workspace, int64_t n, int D, cudaStream_t st, int part,
                       int n_parts) {
  if (D <= kTc64MaxD) return launch_pairwise_tc64(X, out, workspace, n, D, st, part, n_parts);
  float* norms = static_cast<float*>(workspace);
  double* partial = reinterpret_cast<double*>(static_cast<unsigned char*>(workspace) +
                                              ((size_t)n * sizeof(float) + 255) / 256 * 256);
  row_norms_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(X, norms, n, D);
  DIC_LAUNCH_CHECK("row_norms_kernel");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t nb = (n + kTile - 1) / kTile;
  const int64_t ntiles = nb * (nb + 1) / 2;
  const int64_t my_tiles = ntiles / n_parts > 0 ? ntiles / n_parts : 1;
  int blocks = (int)(my_tiles < sms ? my_tiles : sms);
  if (blocks > 1024) blocks = 1024;
  const size_t smem = sizeof(TcSmem) + 1024;
  DIC_CUDA(cudaFuncSetAttribute(pairwise_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#ifdef DIC_TC_PROFILE      // benchmark build only: per-role cycle counters of CTA 0, printed after the run
  long long* dbg = nullptr;
    cudaMalloc(&dbg, 16 * sizeof(long long));
    cudaMemsetAsync(dbg, 0, 16 * sizeof(long long), st);
#endif
  pairwise_tc_kernel<<<blocks, kThreads, smem, st>>>(X, norms, partial, n, D, DIC_PROF(dbg,) part, n_parts);
#ifdef DIC_TC_PROFILE
  if (dbg) {
    long long h[16];
    cudaMemcpyAsync(h, dbg, sizeof(h), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    fprintf(stderr, "[tc profile, CTA 0, %lld stages] loader: wait_stage %lld reload_A %lld store %lld prefetch %lld "
            "barrier %lld | issuer: wait_acc %lld total %lld | epilogue: wait_full %lld drain %lld (cycles/stage)\n",
            h[5], h[0] / (h[5] + 1), h[1] / (h[5] + 1), h[2] / (h[5] + 1), h[3] / (h[5] + 1), h[4] / (h[5] + 1),
            h[6] / (h[5] + 1), h[7] / (h[5] + 1), h[8] / (h[5] + 1), h[9] / (h[5] + 1));
    cudaFree(dbg);
  }
#endif
  DIC_LAUNCH_CHECK("pairwise_tc_kernel");
  sum_partials_kernel<<<1, 32, 0, st>>>(partial, out, blocks);
  DIC_LAUNCH_CHECK("sum_partials_kernel");
  return DIC_OK;
}

}  // namespace dic
