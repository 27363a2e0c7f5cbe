// One Lloyd pass with the E-step on the tensor cores (tcgen05 + TMEM): float32 rows of 64 / 128 / 256 elements,
// K <= 16.  sklearn/cluster/_k_means_lloyd.pyx:23-218 (E-step: argmin_j ||c_j||^2 - 2 x.c_j, lowest index on ties;
// accumulation of the per-cluster sums and counts for the M-step).
//
// Why: the CUDA-core kernels (kmeans.cu) pay one shared-memory broadcast operand per FMA - K*D operands per row -
// and sit at 0.1-0.4 of the HBM roofline (ncu: issue / LSU bound).  D x K >= 64 x 8 is a genuine dense contraction
// (north_star), so the x.c products move to tcgen05.mma and the CUDA cores keep only what is not a contraction:
// operand conversion, the arg-min and the per-cluster accumulation.
//
// Structure (one persistent CTA per SM: TMA producer warp | 4 converter warps | 4 epilogue warps):
//   * work unit ("job") = 128 rows x 64 floats (32 KB), brought in by the producer warp with TMA bulk copies (ONE
//     copy per job when D = 64: the rows of a tile are contiguous; 128 row copies otherwise) into a slot ring;
//   * conversion: every float is split into hi = rna_tf32(x), lo = rna_tf32(x - hi) and written as the two operand
//     tiles in the canonical no-swizzle K-major UMMA layout; the K-chunk stride is 128 x 16 + 16 bytes, which makes
//     both the row-major reads and the chunk-major writes of a warp bank-conflict free;
//   * 3 MMAs per K step (kind::tf32, M = 128, N = 16, K = 8): hi.hi + hi.lo + lo.hi accumulate float32-grade dot
//     products in 16 TMEM columns (the dropped lo.lo term is < 2^-22 |x||c|);
//   * epilogue: thread = row reads its 16 dots with tcgen05.ld, d_j = ||c_j||^2 - 2 x.c_j, strict '<' arg-min;
//   * accumulation (M-step sums): thread = (row group, 16-byte column) adds x into sacc[copy][label][column] in
//     shared memory - the label is an address, no atomics, fixed order (deterministic); wide rows are re-read from
//     L2 once the labels are known, with 8 / (D / 64) private accumulator copies used in rounds;
//   * the epilogue of tile t overlaps the conversion and the MMAs of tile t + 1 (two TMEM accumulators).
// Only the Lloyd-loop form of the pass (DIC_KM_NO_INERTIA, labels re-assigned) takes this kernel; the final
// labelling pass with its direct ||x - c||^2 sums stays on the CUDA-core kernels.
//
// Measured (B200, 1M x 64, profiles/r02_kmeans_tc_*.txt): 0.145 ms per pass for every K <= 16 (the contraction itself
// is free) against 0.097 (K = 4) ... 0.148 ms (K = 16) of the specialised CUDA-core kernel, i.e. no win yet.  Phase
// probes: an empty pipeline (TMA + hand-offs + arg-min) runs at 0.053 ms = 4.8 TB/s; the 24 MMAs of a tile issued by
// one thread cost ~3.2 k clk (an operand-descriptor waterfall per UTCHMMA - the MMA warp must run in uniform control
// flow as in pairwise_tc.cu), the tf32 split ~3.4 k clk (cvt.rna.tf32 is 4 ALU instructions; 4 warps) and the
// accumulation ~2.1 k clk per tile, against 1.5 k clk of HBM time per tile.  The kernel is therefore selectable
// (DIC_KM_KERNEL(5), parity-tested) but not dispatched by default.
#include "common.cuh"
#include "tc_common.cuh"

namespace dic {
namespace {

using namespace tc;

constexpr int kRows = 128;
constexpr int kSlotB = kRows * 256;            // one job: 128 rows x 64 floats
constexpr int kLboA = kRows * 16 + 16;         // bytes between consecutive 16-byte K chunks of an A operand tile
constexpr int kATileB = 16 * kLboA;            // hi or lo tile of one job (16 K chunks)
constexpr int kLboB = 16 * 16;                 // centres: 16 rows x 16 bytes per K chunk
constexpr int kBTileB = 16 * kLboB;            // hi or lo centre tile of one 64-wide K chunk
constexpr int kSbo = 128;                      // bytes between 8-row groups
constexpr int kRole = 128;                     // threads per role (converters, epilogue)
constexpr int kTcThreads = 32 + 2 * kRole;     // producer warp | 4 converter warps | 4 epilogue warps
constexpr uint32_t kIdescTf32N16 = make_idesc(2u, 128u, 16u);

struct TcBars {
  uint64_t full[3], slot_free[3], a_free, acc_full[2], acc_free[2];
  uint32_t tmem_base;
  int timeout;
};

template <int NCH>
struct TcLayout {
  static constexpr int D = 64 * NCH;
  static constexpr int G = 8 / NCH;             // private accumulator copies
  static constexpr int NSLOT = NCH == 1 ? 3 : 2;
  static constexpr size_t slots = 0;
  static constexpr size_t a_hi = slots + NSLOT * (size_t)kSlotB;
  static constexpr size_t a_lo = a_hi + kATileB;
  static constexpr size_t b_hi = a_lo + kATileB;
  static constexpr size_t b_lo = b_hi + (size_t)NCH * kBTileB;
  static constexpr size_t sacc = b_lo + (size_t)NCH * kBTileB;          // [G][16][D] floats (K <= 16)
  static constexpr size_t scn = sacc + (size_t)G * 16 * D * 4;          // [16] floats
  static constexpr size_t scnt = scn + 64;                              // [8][16] ints
  static constexpr size_t slab = scnt + 8 * 16 * 4;                     // [128] ints
  static constexpr size_t bars = slab + kRows * 4;
  static constexpr size_t total = bars + sizeof(TcBars) + 16;
};

__device__ __forceinline__ bool wait_bar(TcBars* B, uint64_t* bar, uint32_t phase) {
  if (*reinterpret_cast<volatile int*>(&B->timeout)) return false;
  if (bar_wait_bounded(bar, phase)) return true;
  *reinterpret_cast<volatile int*>(&B->timeout) = 1;      // a stalled hand-off ends the pass with NaN statistics, not a hang
  return false;
}

__device__ __forceinline__ void split_tf32(const float4& x, float4& h, float4& l) {
  h.x = to_tf32(x.x); h.y = to_tf32(x.y); h.z = to_tf32(x.z); h.w = to_tf32(x.w);
  l.x = to_tf32(x.x - h.x); l.y = to_tf32(x.y - h.y); l.z = to_tf32(x.z - h.z); l.w = to_tf32(x.w - h.w);
}

// Warp-specialised pipeline, one persistent CTA per SM:
//   warp 0      producer: TMA bulk copies of the jobs (128 rows x 64 floats) into the slot ring
//   warps 1-4   converters: slot -> split tf32 operand tiles; their thread 0 issues the 24 MMAs of the job
//   warps 5-8   epilogue: tcgen05.ld of tile t's 16 dots per row, arg-min, labels, per-cluster accumulation - while
//               the converters and the tensor core work on tile t + 1 (two TMEM accumulators of 16 columns)
template <int NCH>
__global__ void __launch_bounds__(kTcThreads, 1)
kmeans_assign_tc_kernel(const float* __restrict__ X, const float* __restrict__ centers, int32_t* labels,
                        double* __restrict__ ws, int64_t N, int K, int flags, int want_sums,
                        const double* __restrict__ done) {
  if (done && *done != 0.0) return;
  using L = TcLayout<NCH>;
  constexpr int D = L::D, G = L::G, ROUNDS = NCH, NSLOT = L::NSLOT;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* slots = smem + L::slots;
  unsigned char* a_hi = smem + L::a_hi;
  unsigned char* a_lo = smem + L::a_lo;
  unsigned char* b_hi = smem + L::b_hi;
  unsigned char* b_lo = smem + L::b_lo;
  float* sacc = reinterpret_cast<float*>(smem + L::sacc);
  float* scn = reinterpret_cast<float*>(smem + L::scn);
  int* scnt = reinterpret_cast<int*>(smem + L::scnt);
  int* slab = reinterpret_cast<int*>(smem + L::slab);
  TcBars* B = reinterpret_cast<TcBars*>(smem + L::bars);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool m_from_slot = NCH == 1 && want_sums;      // the epilogue reads the rows of a one-job tile from its slot

  if (tid == 0) {
    for (int s = 0; s < 3; ++s) {
      mbar_init(&B->full[s], 1);
      mbar_init(&B->slot_free[s], m_from_slot ? 2 : 1);      // converters (+ epilogue) are done with the slot
    }
    mbar_init(&B->a_free, 1);
    for (int k = 0; k < 2; ++k) {
      mbar_init(&B->acc_full[k], 1);
      mbar_init(&B->acc_free[k], 4);                           // one arrival per epilogue warp
    }
    B->timeout = 0;
    fence_proxy_async();
  }
  if (warp == 0) tmem_alloc(&B->tmem_base, 32);
  // centres -> split operand tiles [chunk][K chunk][16 rows][16 bytes] (rows >= K are zero), norms, zeroed accumulators
  for (int idx = tid; idx < NCH * 256; idx += kTcThreads) {
    const int n = idx & 15, c16 = (idx >> 4) & 15, ch = idx >> 8;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f), h, l;
    if (n < K) x = __ldg(reinterpret_cast<const float4*>(centers + (size_t)n * D + ch * 64 + c16 * 4));
    split_tf32(x, h, l);
    const int off = ch * kBTileB + c16 * kLboB + n * 16;
    *reinterpret_cast<float4*>(b_hi + off) = h;
    *reinterpret_cast<float4*>(b_lo + off) = l;
  }
  if (want_sums)
    for (int i = tid; i < G * 16 * D; i += kTcThreads) sacc[i] = 0.f;
  for (int i = tid; i < 8 * 16; i += kTcThreads) scnt[i] = 0;
  if (tid < 16) {
    float sum = 0.f;
    if (tid < K)
      for (int d = 0; d < D; ++d) {
        const float c = __ldg(centers + (size_t)tid * D + d);
        sum += c * c;
      }
    scn[tid] = sum;
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
    // ================= producer =================
    int64_t j = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int rows = (int)min((int64_t)kRows, N - t * kRows);
      for (int ch = 0; ch < NCH; ++ch, ++j) {
        const int slot = (int)(j % NSLOT);
        const uint32_t use = (uint32_t)(j / NSLOT);
        const bool ok = wait_bar(B, &B->slot_free[slot], (use & 1u) ^ 1u);
        if (!__all_sync(0xffffffffu, ok)) break;
        unsigned char* dst = slots + (size_t)slot * kSlotB;
        if (lane == 0) mbar_expect_tx(&B->full[slot], (uint32_t)rows * 256u);
        __syncwarp();
        if (NCH == 1) {
          if (lane == 0) bulk_g2s(dst, X + (size_t)t * kRows * D, (uint32_t)rows * 256u, &B->full[slot]);
        } else {
          for (int r = lane; r < rows; r += 32)
            bulk_g2s(dst + r * 256, X + ((size_t)t * kRows + r) * D + ch * 64, 256u, &B->full[slot]);
        }
      }
    }
  } else if (warp <= 4) {
    // ================= converters + MMA issue =================
    const int ct = tid - 32;
    const uint32_t a_hi_u = smem_u32(a_hi), a_lo_u = smem_u32(a_lo), b_hi_u = smem_u32(b_hi), b_lo_u = smem_u32(b_lo);
    int64_t j = 0, tl = 0;                     // job and tile counters of this CTA
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++tl) {
      const int buf = (int)(tl & 1);
      for (int ch = 0; ch < NCH; ++ch, ++j) {
        const int slot = (int)(j % NSLOT);
        wait_bar(B, &B->full[slot], (uint32_t)(j / NSLOT) & 1u);
        if (j > 0) wait_bar(B, &B->a_free, (uint32_t)(j - 1) & 1u);      // the MMAs of the previous job have read A
        const float* sx = reinterpret_cast<const float*>(slots + (size_t)slot * kSlotB);
#pragma unroll 4
        for (int it = 0; it < 16; ++it) {
          const int item = it * kRole + ct;
          const int row = item >> 4, c = item & 15;        // lanes 0-15: one row's 16 chunks; 16-31: the next row
          const float4 x = *reinterpret_cast<const float4*>(sx + row * 64 + c * 4);
          float4 h, l;
          split_tf32(x, h, l);
          const int off = c * kLboA + row * 16;
          *reinterpret_cast<float4*>(a_hi + off) = h;
          *reinterpret_cast<float4*>(a_lo + off) = l;
        }
        fence_proxy_async();               // generic-proxy stores -> visible to the tensor core (async proxy)
        named_bar_sync(1, kRole);
        if (ct == 0) {
          mbar_arrive(&B->slot_free[slot]);                                 // the converters are done with the slot
          if (ch == 0 && tl >= 2) wait_bar(B, &B->acc_free[buf], (uint32_t)((tl >> 1) - 1) & 1u);   // accumulator drained
          tc_fence_after();
          const uint32_t d_tmem = tmem + (uint32_t)buf * 16u;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {       // one MMA consumes K = 8 tf32 = two 16-byte chunks
            const uint64_t dah = make_desc_kmajor(a_hi_u + ks * 2 * kLboA, kLboA, kSbo);
            const uint64_t dal = make_desc_kmajor(a_lo_u + ks * 2 * kLboA, kLboA, kSbo);
            const uint64_t dbh = make_desc_kmajor(b_hi_u + ch * kBTileB + ks * 2 * kLboB, kLboB, kSbo);
            const uint64_t dbl = make_desc_kmajor(b_lo_u + ch * kBTileB + ks * 2 * kLboB, kLboB, kSbo);
            umma_tf32(d_tmem, dah, dbh, kIdescTf32N16, (ch > 0 || ks > 0) ? 1u : 0u);
            umma_tf32(d_tmem, dah, dbl, kIdescTf32N16, 1u);
            umma_tf32(d_tmem, dal, dbh, kIdescTf32N16, 1u);
          }
          umma_commit(&B->a_free);                           // operand tiles reusable once these MMAs have read them
          if (ch == NCH - 1) umma_commit(&B->acc_full[buf]); // the tile's 16 dots per row are complete
        }
      }
    }
  } else {
    // ================= epilogue: arg-min, labels, per-cluster accumulation =================
    const int et = tid - 32 - kRole;                  // 0..127
    const int row = 32 * (warp & 3) + lane;           // TMEM lane quarter of this warp = warp % 4
    const int g8 = et >> 4, col = et & 15;
    int64_t tl = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++tl) {
      const int buf = (int)(tl & 1);
      const int64_t row0 = t * kRows;
      const int rows = (int)min((int64_t)kRows, N - row0);
      int oldl = -1;
      if (count_changes && row < rows) oldl = labels[row0 + row];
      wait_bar(B, &B->acc_full[buf], (uint32_t)(tl >> 1) & 1u);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld16(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)buf * 16u, v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&B->acc_free[buf]);           // this warp has its dots in registers
      int best = 0;
      float bestd = fmaf(-2.f, __uint_as_float(v[0]), scn[0]);
#pragma unroll
      for (int k = 1; k < 16; ++k) {
        const float dk = fmaf(-2.f, __uint_as_float(v[k]), scn[k]);     // ||c||^2 - 2 x.c  (||x||^2 omitted)
        if (k < K && dk < bestd) {                                      // strict '<': lowest index wins ties
          bestd = dk;
          best = k;
        }
      }
      if (row < rows) {
        if (count_changes) changed += (oldl != best);
        labels[row0 + row] = best;
      }
      if (want_sums) {
        slab[row] = best;
        named_bar_sync(2, kRole);                              // the labels of the tile are visible to the epilogue warps
        const int slot = (int)(tl % NSLOT);                    // NCH == 1: tile t of this CTA is job t
        if (NCH == 1) wait_bar(B, &B->full[slot], (uint32_t)(tl / NSLOT) & 1u);      // (long complete: acquire only)
        const float4* sx4 = reinterpret_cast<const float4*>(slots + (size_t)slot * kSlotB);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
          for (int round = 0; round < ROUNDS; ++round) {
            if ((g8 % ROUNDS) == round) {
              const int copy = g8 / ROUNDS;
              for (int r = g8; r < rows; r += 8) {
                const int lab = slab[r];
                // wide rows: the tile went through L2 a moment ago (E jobs); plain coalesced re-reads, no staging
                const float4 xv = NCH == 1 ? sx4[r * 16 + col]
                                           : __ldg(reinterpret_cast<const float4*>(X + (size_t)(row0 + r) * D + ch * 64) + col);
                float4* dst = reinterpret_cast<float4*>(sacc) + ((size_t)(copy * 16 + lab) * NCH + ch) * 16 + col;
                float4 a = *dst;
                a.x += xv.x; a.y += xv.y; a.z += xv.z; a.w += xv.w;
                *dst = a;
                if (col == 0 && ch == 0) scnt[g8 * 16 + lab] += 1;
              }
            }
            if (ROUNDS > 1) named_bar_sync(2, kRole);
          }
        }
        named_bar_sync(2, kRole);                              // slab and the slot are no longer read
        if (NCH == 1 && et == 0) mbar_arrive(&B->slot_free[slot]);
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
      const int k = i / D, d = i - k * D;
      double sum = 0.0;
      for (int c = 0; c < G; ++c) sum += (double)sacc[(size_t)(c * 16 + k) * D + d];
      out[i] = dead ? nan : sum;
    }
    for (int i = tid; i < K; i += kTcThreads) {
      int c = 0;
      for (int gg = 0; gg < 8; ++gg) c += scnt[gg * 16 + i];
      out[(int64_t)K * D + i] = (double)c;
    }
  }
  double* red = reinterpret_cast<double*>(slots);
  const double chs = warp_sum((double)changed);
  if (lane == 0) red[warp] = chs;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < kTcThreads / 32; ++w) s += red[w];
    out[(int64_t)K * D + K + 0] = dead ? nan : 0.0;      // inertia: not computed in this form of the pass
    out[(int64_t)K * D + K + 1] = dead ? nan : s;
    out[(int64_t)K * D + K + 2] = 0.0;
    out[(int64_t)K * D + K + 3] = 0.0;
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 32);
  }
}

}  // namespace

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
#define DIC_KTC(NCH_)                                                                                          \
  {                                                                                                            \
    auto kf = kmeans_assign_tc_kernel<NCH_>;                                                                   \
    const size_t smem = TcLayout<NCH_>::total;                                                                 \
    DIC_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                \
    kf<<<nb, kTcThreads, smem, st>>>(X, centers, labels, ws, N, K, flags, want_sums, done);                    \
  }
  if (D == 64) DIC_KTC(1) else if (D == 128) DIC_KTC(2) else DIC_KTC(4)
#undef DIC_KTC
  DIC_LAUNCH_CHECK("kmeans_assign_tc_kernel");
  return DIC_OK;
}

}  // namespace dic
