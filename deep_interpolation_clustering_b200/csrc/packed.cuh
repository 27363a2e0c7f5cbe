// The RAGGED (packed) encounter format shared by the upload path and the interpolation kernels.
//
//   n_obs   (B, C) int32    valid observations of vital c of encounter b (the left-packed prefix length,
//                           p0_data_process.py:44-67)
//   enc_off (B + 1) int64   float offset of encounter b in `packed`; a multiple of 4 (16-byte rows for TMA)
//   packed                  for b, for c: [ value[0..n4) | time[0..n4) ],  n4 = round_up(n_obs[b,c], 4);
//                           pad slots hold value 0 and time kPackedPadTime, i.e. observations whose Gaussian
//                           weight is exactly 0 - the same null entries the dense staging pads rows with
//                           (interp_stage.cuh: warp_pad4_far), so a packed row IS a canonical shared-memory row.
#pragma once

#include <stdint.h>

namespace dic {

constexpr int kPackedMaxC = 16;
constexpr float kPackedPadTime = 3.0e18f;   // a far-away time; the expansion writes 0 in the dense planes, whose kernels pad on their own

__host__ __device__ inline int packed_round4(int n) { return (n + 3) & ~3; }

}  // namespace dic
