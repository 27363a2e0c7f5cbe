// Host -> device paths for a batch of encounters.
//
// The trainer's `ob / padding_mask / timestamp .to(device)` (pretrain_trainer.py:132-136) ships dense
// (B, C, T) planes although the pipeline's rows are left-packed (p0_data_process.py:44-67): on average half
// of every plane is padding and the mask plane of a left-packed row is one integer.  Two paths:
//   * dic_upload_encounters          one strided DMA of the three live planes (any mask);
//   * dic_pack_encounters_host +     the RAGGED form: per (encounter, vital) only the valid prefix of the
//     dic_upload_encounters_packed   value and time rows plus one count; ~3x fewer PCIe bytes at c2.  The
//     + dic_expand_encounters        interpolation kernels stage straight from it (dic_*_packed entry points,
//                                    interp_*.cu), or it is expanded on the device to the dense planes.
#include <string.h>

#include "common.cuh"
#include "packed.cuh"

namespace dic {
namespace {

// One CTA per encounter: rebuilds the three dense planes [value | mask | time] of x (B, dev_planes, T) from the
// packed rows.  Every thread writes 4 consecutive slots of one row with one 128-bit store when T % 4 == 0.
__global__ void __launch_bounds__(256)
expand_encounters_kernel(const float* __restrict__ packed, const int32_t* __restrict__ n_obs,
                         const int64_t* __restrict__ enc_off, float* __restrict__ x, int64_t B, int C, int T,
                         int64_t x_stride, int vec4) {
  __shared__ int s_off[kPackedMaxC + 1];
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const int32_t* nb = n_obs + b * C;
    if (threadIdx.x == 0) {
      int o = 0;
      for (int c = 0; c < C; ++c) {
        s_off[c] = o;
        o += 2 * packed_round4(nb[c]);
      }
    }
    __syncthreads();
    const float* pb = packed + (enc_off[b] - enc_off[0]);
    float* xb = x + b * x_stride;
    if (vec4) {
      const int T4 = T >> 2;
      for (int i = threadIdx.x; i < C * T4; i += blockDim.x) {
        const int c = i / T4, t = (i - c * T4) << 2;
        const int n4 = packed_round4(nb[c]), n = min(nb[c], T);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f), d = v, m = v;
        if (t < n) {      // the packed rows are 16-byte aligned and padded to whole groups of 4
          v = *reinterpret_cast<const float4*>(pb + s_off[c] + t);
          d = *reinterpret_cast<const float4*>(pb + s_off[c] + n4 + t);
          m = make_float4(1.f, t + 1 < n ? 1.f : 0.f, t + 2 < n ? 1.f : 0.f, t + 3 < n ? 1.f : 0.f);
          d.y *= m.y; d.z *= m.z; d.w *= m.w;      // pad slots hold kPadTime in the packed row, 0 in the dense one
          v.y *= m.y; v.z *= m.z; v.w *= m.w;
        }
        *reinterpret_cast<float4*>(xb + (int64_t)(0 * C + c) * T + t) = v;
        *reinterpret_cast<float4*>(xb + (int64_t)(1 * C + c) * T + t) = m;
        *reinterpret_cast<float4*>(xb + (int64_t)(2 * C + c) * T + t) = d;
      }
    } else {
      for (int i = threadIdx.x; i < C * T; i += blockDim.x) {
        const int c = i / T, t = i - c * T;
        const int n4 = packed_round4(nb[c]), n = min(nb[c], T);
        const bool live = t < n;
        xb[(int64_t)(0 * C + c) * T + t] = live ? pb[s_off[c] + t] : 0.f;
        xb[(int64_t)(1 * C + c) * T + t] = live ? 1.f : 0.f;
        xb[(int64_t)(2 * C + c) * T + t] = live ? pb[s_off[c] + n4 + t] : 0.f;
      }
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace dic

using namespace dic;

extern "C" int dic_upload_encounters(float* x_dev, const float* x_host, int64_t B, int C, int T,
                                     int host_planes, int dev_planes, dic_stream_t stream) {
  DIC_REQUIRE(x_dev && x_host, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE(B >= 0 && C > 0 && T > 0, DIC_ERR_INVALID_ARGUMENT, "bad sizes B=%lld C=%d T=%d", (long long)B, C, T);
  DIC_REQUIRE(host_planes >= 3 * C && dev_planes >= 3 * C, DIC_ERR_INVALID_ARGUMENT,
              "an encounter has 3*C live planes: host_planes=%d dev_planes=%d C=%d", host_planes, dev_planes, C);
  if (B == 0) return DIC_OK;
  // one strided DMA: B rows of 3*C*T floats (value | mask | time); the hold-out plane stays on the host
  DIC_CUDA(cudaMemcpy2DAsync(x_dev, (size_t)dev_planes * T * sizeof(float), x_host,
                             (size_t)host_planes * T * sizeof(float), (size_t)3 * C * T * sizeof(float),
                             (size_t)B, cudaMemcpyHostToDevice, as_stream(stream)));
  return DIC_OK;
}

extern "C" int64_t dic_pack_encounters_host(const float* x_host, int64_t B, int C, int T, int host_planes,
                                            int32_t* n_obs_host, int64_t* enc_off_host, float* packed_host,
                                            int64_t capacity, int* all_sorted) {
  DIC_REQUIRE(x_host || B == 0, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE((n_obs_host || B == 0) && enc_off_host, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE(B >= 0 && C > 0 && C <= kPackedMaxC && T > 0, DIC_ERR_INVALID_ARGUMENT,
              "bad sizes B=%lld C=%d (<= %d) T=%d", (long long)B, C, kPackedMaxC, T);
  DIC_REQUIRE(host_planes >= 3 * C, DIC_ERR_INVALID_ARGUMENT, "host_planes=%d < 3*C", host_planes);
  const size_t es = (size_t)host_planes * T;
  int64_t off = 0;
  int sorted = 1;
  for (int64_t b = 0; b < B; ++b) {
    const float* xb = x_host + (size_t)b * es;
    enc_off_host[b] = off;
    for (int c = 0; c < C; ++c) {
      const float* val = xb + (size_t)(0 * C + c) * T;
      const float* msk = xb + (size_t)(1 * C + c) * T;
      const float* tim = xb + (size_t)(2 * C + c) * T;
      int n = 0;
      while (n < T && msk[n] == 1.0f) ++n;
      for (int t = n; t < T; ++t)
        DIC_REQUIRE(msk[t] == 0.0f, DIC_ERR_UNSUPPORTED,
                    "encounter %lld vital %d: the mask is not a left-packed 0/1 prefix (slot %d = %g); use the dense "
                    "upload for general masks", (long long)b, c, t, (double)msk[t]);
      for (int t = 1; t < n; ++t) sorted &= (tim[t - 1] <= tim[t]);
      const int n4 = packed_round4(n);
      n_obs_host[b * C + c] = n;
      if (packed_host) {
        DIC_REQUIRE(off + 2 * n4 <= capacity, DIC_ERR_INVALID_ARGUMENT,
                    "packed buffer too small (capacity %lld floats)", (long long)capacity);
        float* pv = packed_host + off;
        float* pd = pv + n4;
        memcpy(pv, val, sizeof(float) * n);
        memcpy(pd, tim, sizeof(float) * n);
        for (int t = n; t < n4; ++t) {
          pv[t] = 0.0f;
          pd[t] = kPackedPadTime;
        }
      }
      off += 2 * n4;
    }
  }
  enc_off_host[B] = off;
  if (all_sorted) *all_sorted = sorted;
  return off;
}

extern "C" int dic_upload_encounters_packed(const float* packed_host, const int32_t* n_obs_host,
                                            const int64_t* enc_off_host, float* packed_dev, int32_t* n_obs_dev,
                                            int64_t* enc_off_dev, int64_t B, int C, dic_stream_t stream) {
  DIC_REQUIRE(B >= 0 && C > 0, DIC_ERR_INVALID_ARGUMENT, "bad sizes B=%lld C=%d", (long long)B, C);
  if (B == 0) return DIC_OK;
  DIC_REQUIRE(packed_host && n_obs_host && enc_off_host && packed_dev && n_obs_dev && enc_off_dev,
              DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE(aligned16(packed_dev), DIC_ERR_INVALID_ARGUMENT, "packed_dev must be 16-byte aligned");
  const int64_t floats = enc_off_host[B] - enc_off_host[0];
  DIC_REQUIRE(floats >= 0, DIC_ERR_INVALID_ARGUMENT, "enc_off_host is not increasing");
  cudaStream_t st = as_stream(stream);
  // three contiguous copies; offsets stay absolute (the kernels subtract enc_off[0] of the chunk)
  DIC_CUDA(cudaMemcpyAsync(enc_off_dev, enc_off_host, sizeof(int64_t) * (size_t)(B + 1), cudaMemcpyHostToDevice, st));
  DIC_CUDA(cudaMemcpyAsync(n_obs_dev, n_obs_host, sizeof(int32_t) * (size_t)B * C, cudaMemcpyHostToDevice, st));
  if (floats > 0)
    DIC_CUDA(cudaMemcpyAsync(packed_dev, packed_host + enc_off_host[0], sizeof(float) * (size_t)floats,
                             cudaMemcpyHostToDevice, st));
  return DIC_OK;
}

extern "C" int dic_expand_encounters(const float* packed_dev, const int32_t* n_obs_dev, const int64_t* enc_off_dev,
                                     float* x_dev, int64_t B, int C, int T, int dev_planes, dic_stream_t stream) {
  DIC_REQUIRE(B >= 0 && C > 0 && C <= kPackedMaxC && T > 0, DIC_ERR_INVALID_ARGUMENT,
              "bad sizes B=%lld C=%d (<= %d) T=%d", (long long)B, C, kPackedMaxC, T);
  DIC_REQUIRE(dev_planes >= 3 * C, DIC_ERR_INVALID_ARGUMENT, "dev_planes=%d < 3*C", dev_planes);
  if (B == 0) return DIC_OK;
  DIC_REQUIRE(packed_dev && n_obs_dev && enc_off_dev && x_dev, DIC_ERR_INVALID_ARGUMENT, "null pointer argument");
  DIC_REQUIRE(aligned16(packed_dev), DIC_ERR_INVALID_ARGUMENT, "packed_dev must be 16-byte aligned");
  const int64_t stride = (int64_t)dev_planes * T;
  const int vec4 = (T % 4 == 0) && aligned16(x_dev) && (stride % 4 == 0);
  int dev = 0, sms = 0;
  DIC_CUDA(cudaGetDevice(&dev));
  DIC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t grid = B < (int64_t)sms * 16 ? B : (int64_t)sms * 16;
  expand_encounters_kernel<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(packed_dev, n_obs_dev, enc_off_dev, x_dev,
                                                                         B, C, T, stride, vec4);
  DIC_LAUNCH_CHECK("expand_encounters_kernel");
  return DIC_OK;
}
