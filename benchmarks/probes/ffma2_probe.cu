// Throughput of packed float32x2 FMA (FFMA2) against scalar FFMA on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
// B200: 32.0 TFMA/s scalar, 31.9 TFMA/s packed (half the instructions): the FMA pipe does not get wider, issue slots are saved.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ffma(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f, a4 = a0 + .4f, a5 = a0 + .5f, a6 = a0 + .6f, a7 = a0 + .7f;
  const float m = 0.999f, c = 1e-4f;
  for (int i = 0; i < iters; ++i) {
    a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
    a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
  }
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 123456.f) out[0] = a0;
}
__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b, const float2& c) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
}
__global__ void k_ffma2(float* out, int iters) {
  float2 a0 = make_float2(threadIdx.x * 1e-3f, 0.05f), a1 = make_float2(a0.x + 0.1f, .15f), a2 = make_float2(a0.x + 0.2f, .25f), a3 = make_float2(a0.x + .3f, .35f);
  const float2 m = make_float2(0.999f, 0.998f), c = make_float2(1e-4f, 2e-4f);
  for (int i = 0; i < iters; ++i) {
    ffma2(a0, a0, m, c); ffma2(a1, a1, m, c); ffma2(a2, a2, m, c); ffma2(a3, a3, m, c);
  }
  if (a0.x + a1.x + a2.x + a3.x + a0.y + a1.y + a2.y + a3.y == 123456.f) out[0] = a0.x;
}
// mixed: MUFU + FFMA2 stream resembling the sweep loops
int main() {
  float* out; cudaMalloc(&out, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, blocks = 148 * 8, threads = 256;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k_ffma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("FFMA : %.3f ms  %.1f Gfma/s\n", ms, 8.0 * iters * blocks * threads / ms / 1e6);
    cudaEventRecord(e0); k_ffma2<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("FFMA2: %.3f ms  %.1f Gfma/s (8 fma per iteration in 4 instructions)\n", ms, 8.0 * iters * blocks * threads / ms / 1e6);
  }
  return 0;
}
