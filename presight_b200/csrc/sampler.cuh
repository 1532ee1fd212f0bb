// Device functions for the samplers (shared with the fused per-level kernels).
#pragma once
#include "common.cuh"

namespace ps {

// PreSight's piecewise spacing (models/PreSight/nerfacto_nusc_ms.py:312-317):
//   fn(x)  = x < thr ? x / (2 thr) : 1 - 1 / (2 x / thr)
//   inv(x) = x < 0.5 ? x * (2 thr) : thr / (2 - 2 x)
__device__ __forceinline__ float spacing_fn(float x, float thr) {
    const float two_thr = __fmul_rn(2.f, thr);
    return x < thr ? __fdiv_rn(x, two_thr) : __fsub_rn(1.f, __fdiv_rn(1.f, __fdiv_rn(__fmul_rn(2.f, x), thr)));
}
__device__ __forceinline__ float spacing_inv(float x, float thr) {
    const float two_thr = __fmul_rn(2.f, thr);
    return x < 0.5f ? __fmul_rn(x, two_thr) : __fdiv_rn(thr, __fsub_rn(2.f, __fmul_rn(2.f, x)));
}
// spacing_to_euclidean_fn (model_components/ray_samplers.py:113-114): inv(x*s_far + (1-x)*s_near)
__device__ __forceinline__ float spacing_to_euclidean(float x, float s_near, float s_far, float thr) {
    return spacing_inv(__fadd_rn(__fmul_rn(x, s_far), __fmul_rn(__fsub_rn(1.f, x), s_near)), thr);
}

// torch.searchsorted(side="right"): number of elements <= v in an ascending array
__device__ __forceinline__ int upper_bound(const float* a, int n, float v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

}  // namespace ps
