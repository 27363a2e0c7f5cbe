// Shared helpers for the sm_100a kernels of libdic_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "dic_b200.h"

namespace dic {

// ---- host side: thread-local error text, argument checks, launch checks -------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define DIC_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      ::dic::set_error(__VA_ARGS__);      \
      return (code);                      \
    }                                     \
  } while (0)

#define DIC_CUDA(call)                                        \
  do {                                                        \
    cudaError_t _e = (call);                                  \
    if (_e != cudaSuccess) return ::dic::cuda_fail(_e, #call); \
  } while (0)

#define DIC_LAUNCH_CHECK(name)                                        \
  do {                                                                \
    cudaError_t _e = cudaGetLastError();                              \
    if (_e != cudaSuccess) return ::dic::cuda_fail(_e, "launch " name); \
  } while (0)

constexpr int kMaxSmemBytes = 227 * 1024;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

static inline cudaStream_t as_stream(dic_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Deterministic column sums of a (rows, cols) fp32 matrix -> fp64, two stages.
// workspace: kColsumBlocks * cols doubles.  scale_by may be NULL (else out[c] *= scale_by[c]).
constexpr int kColsumBlocks = 592;  // 4 x 148 SMs
int colsum_f32_launch(const float* a, double* out_f64, float* out_f32, const float* scale_by,
                      double* workspace, int64_t rows, int cols, cudaStream_t stream);

#ifdef __CUDACC__
// ---- device side ------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- packed float32 x 2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2) ---------------------------------------
// One instruction issues two float32 operations on an aligned register pair.  The FMA pipe does not get wider
// (measured: 32 TFMA/s either way, benchmarks/probes/ffma2_probe.cu) but the sweep loops are ISSUE bound with
// the FMA pipe ~50 % busy, so halving the issue slots of their arithmetic is what moves them.
typedef unsigned long long f2_t;     // .x in the low word, .y in the high word
__device__ __forceinline__ f2_t pack2(float lo, float hi) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) {
  f2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f2_t add2(f2_t a, f2_t b) {
  f2_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b) {
  f2_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float sum2(f2_t v) {
  float lo, hi;
  unpack2(v, lo, hi);
  return lo + hi;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Transposed butterfly: reduces KP per-lane values across the warp in log2(KP) halving steps
// (KP/2 + KP/4 + ... shuffles) followed by plain xor steps; afterwards lane L holds the total of
// value index L >> (5 - log2 KP).  KP is a power of two <= 32.
template <int KP>
__device__ __forceinline__ float warp_reduce_multi(float (&v)[KP], int lane) {
  int n = KP;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (n > 1) {
      n >>= 1;
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < (KP > 1 ? KP / 2 : 1); ++i) {
        if (i < n) {
          const float send = up ? v[i] : v[i + n];
          const float keep = up ? v[i + n] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    }
  }
  return v[0];
}

__device__ __forceinline__ float softplus_ref(float k) {
  // log(1 + exp(k)) with the plain formula of interpolation_layer.py:51 / rbf.py:78
  return logf(1.0f + expf(k));
}
__device__ __forceinline__ float sigmoid_ref(float k) { return 1.0f / (1.0f + expf(-k)); }

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS: UBLKCP) ----------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
#endif  // __CUDACC__

}  // namespace dic
