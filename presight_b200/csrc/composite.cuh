// Device functions for volumetric compositing (shared with the fused per-level kernels).
#pragma once
#include "common.cuh"

namespace ps {

// One chunk of 32 consecutive samples of a ray, one sample per lane.
//   dd    : this lane's delta*sigma (0 for padding lanes)
//   carry : sum of dd over all earlier chunks (fp64), updated to include this chunk
// Outputs w = nan_to_num(alpha * T) and T = exp(-sum_{k<i} dd_k)  (cameras/rays.py:138-148).
// Returns whether the raw product was finite (nan_to_num passes gradient only there).
__device__ __forceinline__ bool weight_step(float dd, int lane, double& carry, float& w, float& T) {
    const double incl = warp_scan_incl((double)dd, lane) + carry;
    const double prev = __shfl_up_sync(0xffffffffu, incl, 1);
    const double excl = lane == 0 ? carry : prev;
    carry = __shfl_sync(0xffffffffu, incl, 31);
    T = expf(-(float)excl);
    const float alpha = __fsub_rn(1.f, expf(-dd));
    const float raw = __fmul_rn(alpha, T);
    w = nan_to_num(raw);
    return isfinite(raw);
}

}  // namespace ps
