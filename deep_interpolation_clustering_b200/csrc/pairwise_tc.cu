// Pairwise Euclidean distance sum on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// This is the one genuinely dense contraction on the hot path: the "inertia" of
// KM.compute_inertia_v1 / computer_intertia_v2 (p2_clustering_optK.py:334-351) sums
// ||x_i - x_j|| over the full n_c x n_c matrix of every cluster.  The reference materialises that
// matrix (sklearn pairwise_distances: ||x||^2 + ||y||^2 - 2 x.y, clamp, sqrt); here a 128 x 128
// tile of the Gram matrix X X^T is produced by tcgen05.mma (kind::tf32, accumulator in TMEM), read
// back with tcgen05.ld and reduced on the fly, so nothing is ever written to memory.
//
//   * split TF32: each operand is split in shared memory into hi = rn_tf32(x) and lo = rn_tf32(x - hi)
//     and the tile accumulates hi.hi + hi.lo + lo.hi + lo.lo, which restores float32-grade dot
//     products (the MMAs are ~3% of the tile time, the epilogue's sqrt is the bound; the caller
//     centres the cluster first, so ||x||^2 stays small against the distances).
//   * operands sit in the canonical no-swizzle K-major UMMA layout (8-row x 16-byte core matrices);
//     the row block of a tile row is kept in shared memory and reused across the tiles of that row.
//   * upper triangle only: off-diagonal tiles count twice; the diagonal is excluded explicitly.
//   * the epilogue is the bound (one sqrt per pair): 8 warps read TMEM (warp w -> lanes 32 (w & 3),
//     columns 64 (w >> 2)), form ni + nj - 2 dot, clamp, sqrt.approx, and keep float64 sums.
#include "common.cuh"

namespace dic {
namespace {

constexpr int kTile = 128;           // tile rows (i) and columns (j)
constexpr int kKC = 64;              // K chunk (float32 elements) held in shared memory at once
constexpr int kThreads = 256;
constexpr int kTileBytes = kTile * kKC * 4;          // 32 KB per operand tile
constexpr int kChunkStride = kTile * 16;             // bytes between consecutive 16-byte K chunks (LBO)
constexpr int kGroupStride = 128;                    // bytes between 8-row groups (SBO)
constexpr uint32_t kTmemCols = 128;

struct __align__(16) TcSmem {
  // operand tiles first (16-byte aligned), then scalars
  unsigned char a_hi[kTileBytes], a_lo[kTileBytes], b_hi[kTileBytes], b_lo[kTileBytes];
  float nj[kTile];
  uint64_t bar;
  uint32_t tmem_base;
  int timeout;
};

// ---- raw tcgen05 / TMEM PTX (forms as emitted by CUTLASS's sm100 headers) ------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)),
               "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, M = 128, N = 128, K = 8 (tf32), single CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31},"
      "[%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
        "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
        "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

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

__device__ __forceinline__ bool bar_wait_bounded(uint64_t* bar, uint32_t phase) {
  for (int spin = 0; spin < (1 << 22); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    if (ok) return true;
  }
  return false;
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
  return __uint_as_float(y);
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// rows r0.. of X (n x D), K elements [k0, k0 + 64) -> hi / lo operand tiles in the UMMA layout.
// Consecutive threads take consecutive rows of one 16-byte chunk: shared stores are contiguous
// (conflict-free); the strided global reads come from L1/L2 (a row block is reused many times).
__device__ __forceinline__ void load_split_tile(const float* __restrict__ X, int64_t n, int D, int64_t r0, int k0,
                                                unsigned char* hi, unsigned char* lo) {
  for (int idx = threadIdx.x; idx < kTile * (kKC / 4); idx += kThreads) {
    const int c = idx >> 7, r = idx & (kTile - 1);
    const int64_t gr = r0 + r;
    const int k = k0 + 4 * c;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr < n && k < D) v = __ldg(reinterpret_cast<const float4*>(X + gr * D + k));   // D % 4 == 0
    float4 h, l;     // hi = rn_tf32(x), lo = rn_tf32(x - hi): round-to-nearest keeps the split unbiased
    h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
    l.x = to_tf32(v.x - h.x); l.y = to_tf32(v.y - h.y); l.z = to_tf32(v.z - h.z); l.w = to_tf32(v.w - h.w);
    const int off = c * kChunkStride + (r >> 3) * kGroupStride + (r & 7) * 16;
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

__global__ void __launch_bounds__(kThreads, 1)
pairwise_tc_kernel(const float* __restrict__ X, const float* __restrict__ norms, double* __restrict__ partial,
                   int64_t n, int D) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  TcSmem& S = *reinterpret_cast<TcSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&S.bar, 1);
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
  const bool keep_a = kchunks == 1;               // the row block stays resident across a tile row
  int64_t cur_bi = -1;
  uint32_t phase = 0;
  double total = 0.0;
  const uint32_t a_hi = smem_u32(S.a_hi), a_lo = smem_u32(S.a_lo), b_hi = smem_u32(S.b_hi), b_lo = smem_u32(S.b_lo);

  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    // unrank t -> (bi <= bj), row-major over the upper triangle
    int64_t bi = (int64_t)((2.0 * nb + 1.0 - sqrt((2.0 * nb + 1.0) * (2.0 * nb + 1.0) - 8.0 * (double)t)) * 0.5);
    while (bi * nb - bi * (bi - 1) / 2 > t) --bi;
    while ((bi + 1) * nb - (bi + 1) * bi / 2 <= t) ++bi;
    const int64_t bj = bi + (t - (bi * nb - bi * (bi - 1) / 2));
    const int64_t i0 = bi * kTile, j0 = bj * kTile;

    for (int kc = 0; kc < kchunks; ++kc) {
      if (!(keep_a && bi == cur_bi)) load_split_tile(X, n, D, i0, kc * kKC, S.a_hi, S.a_lo);
      load_split_tile(X, n, D, j0, kc * kKC, S.b_hi, S.b_lo);
      if (kc == 0 && tid < kTile) S.nj[tid] = (j0 + tid < n) ? __ldg(norms + j0 + tid) : 0.f;
      fence_proxy_async();            // generic-proxy stores -> visible to the tensor core (async proxy)
      tc_fence_before();              // previous tile's tcgen05.ld ordered before this tile's MMAs
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < kKC / 8; ++s) {        // one MMA consumes K = 8 tf32 = two 16-byte chunks
          const uint32_t koff = (uint32_t)s * 2u * kChunkStride;
          const uint64_t dah = make_desc(a_hi + koff), dal = make_desc(a_lo + koff);
          const uint64_t dbh = make_desc(b_hi + koff), dbl = make_desc(b_lo + koff);
          umma_tf32(tmem, dah, dbh, kIdesc, (kc > 0 || s > 0) ? 1u : 0u);
          umma_tf32(tmem, dah, dbl, kIdesc, 1u);
          umma_tf32(tmem, dal, dbh, kIdesc, 1u);
          umma_tf32(tmem, dal, dbl, kIdesc, 1u);
        }
        umma_commit(&S.bar);          // arrives when every MMA above has completed
      }
      if (!bar_wait_bounded(&S.bar, phase)) S.timeout = 1;
      phase ^= 1;
      tc_fence_after();
      if (kc + 1 < kchunks) __syncthreads();   // operand tiles are free again
    }
    cur_bi = bi;

    // epilogue: warp w reads TMEM lanes 32 (w & 3), columns 64 (w >> 2) .. +63
    const int row = 32 * (warp & 3) + lane;
    const int64_t gi = i0 + row;
    const float ni = gi < n ? __ldg(norms + gi) : 0.f;
    float tile_sum = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col0 = 64 * (warp >> 2) + 32 * h;
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)col0, v);
      if (gi < n) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int64_t gj = j0 + col0 + c;
          const float d2 = fmaf(-2.f, __uint_as_float(v[c]), ni + S.nj[col0 + c]);
          const float d = sqrt_approx(fmaxf(d2, 0.f));
          tile_sum += (gj < n && gj != gi) ? d : 0.f;
        }
      }
    }
    total += (bi == bj) ? (double)tile_sum : 2.0 * (double)tile_sum;
    // every warp is done with TMEM and with nj[] before the next tile overwrites them
    tc_fence_before();
    __syncthreads();
  }

  __shared__ double red[kThreads / 32];
  total = warp_sum(total);
  if (lane == 0) red[warp] = total;
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) s += red[w];
    partial[blockIdx.x] = S.timeout ? __longlong_as_double(0x7ff8000000000000LL) : s;   // NaN = MMA never completed
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

}  // namespace

size_t pairwise_tc_workspace_bytes(int64_t n) {
  return ((size_t)n * sizeof(float) + 255) / 256 * 256 + 1024 * sizeof(double);
}

bool pairwise_tc_supported(const void* X, int D) { return D % 4 == 0 && D >= 4 && aligned16(X); }

int launch_pairwise_tc(const float* X, double* out, void* workspace, int64_t n, int D, cudaStream_t st) {
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
  int blocks = (int)(ntiles < sms ? ntiles : sms);
  if (blocks > 1024) blocks = 1024;
  const size_t smem = sizeof(TcSmem) + 1024;
  DIC_CUDA(cudaFuncSetAttribute(pairwise_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pairwise_tc_kernel<<<blocks, kThreads, smem, st>>>(X, norms, partial, n, D);
  DIC_LAUNCH_CHECK("pairwise_tc_kernel");
  sum_partials_kernel<<<1, 32, 0, st>>>(partial, out, blocks);
  DIC_LAUNCH_CHECK("sum_partials_kernel");
  return DIC_OK;
}

}  // namespace dic
