// Depth-supervision terms of the loss dict as ONE kernel (loss sums + gradients): expected-depth loss (LiDAR or
// mono-depth) and the line-of-sight loss on the final level's weights.
// Reference: model_components/PreSight/losses.py:24-103, call sites models/PreSight/nerfacto_nusc_ms.py:577-629.
// One warp per ray, lanes over the samples; the per-ray arithmetic lives in depth_loss_core.h (shared with the host
// harness).  The means are over the rays that pass the depth mask, whose number is only known at the end of the grid:
// the kernel accumulates {count, sum of squared errors, sum of line-of-sight terms} and writes un-normalised gradients;
// the caller divides (on the device).
#include "common.cuh"

#define PS_HD __device__
#define PS_EXPF(a) expf(a)
#define PS_LOGF(a) logf(a)
#include "depth_loss_core.h"

namespace ps {

constexpr int kDepthWarps = 8;

__global__ void __launch_bounds__(kDepthWarps * 32) depth_losses_kernel(
    const float* __restrict__ weights, const float* __restrict__ eu_bins, const float* __restrict__ steps_m,
    const float* __restrict__ expected, const float* __restrict__ target, const float* __restrict__ sky, int64_t N, int S,
    float pose_scale, const float* __restrict__ pose_scale_dev, float sigma, float upper_bound, int mode,
    float* __restrict__ sums, float* __restrict__ g_expected, float* __restrict__ g_weights) {
    __shared__ float part[kDepthWarps][3];
    if (pose_scale_dev) pose_scale = __ldg(pose_scale_dev);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kDepthWarps + warp;
    float cnt = 0.f, se = 0.f, los = 0.f;
    if (n < N) {
        const float t = __ldg(target + n);
        const bool on = depthloss::ray_supervised(t, upper_bound, sky, n);
        if (expected) {
            float d_pred = 0.f;
            const float pred = __fdiv_rn(__ldg(expected + n), pose_scale);
            const float v = depthloss::expected_depth_term(t, pred, upper_bound, mode, d_pred);
            if (lane == 0) {
                if (on) se = v;
                if (g_expected) g_expected[n] = on ? d_pred / pose_scale : 0.f;
            }
        }
        if (lane == 0 && on) cnt = 1.f;
        if (weights) {
            const depthloss::LosConsts c = depthloss::los_consts(sigma);
            for (int s = lane; s < S; s += 32) {
                float step;
                if (steps_m) {
                    step = __ldg(steps_m + n * S + s);
                } else {
                    const float* b = eu_bins + n * (S + 1) + s;
                    step = __fdiv_rn(__fdiv_rn(__fadd_rn(__ldg(b), __ldg(b + 1)), 2.f), pose_scale);
                }
                float d_w = 0.f;
                const float v = depthloss::los_term(__ldg(weights + n * S + s), step, t, c, d_w);
                if (on) los += v;
                if (g_weights) g_weights[n * S + s] = on ? d_w : 0.f;
            }
        }
    }
    los = warp_sum(los);
    if (lane == 0) {
        part[warp][0] = cnt;
        part[warp][1] = se;
        part[warp][2] = los;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kDepthWarps; ++w) s += part[w][threadIdx.x];
        if (s != 0.f) atomicAdd(sums + threadIdx.x, s);
    }
}

}  // namespace ps

using namespace ps;

extern "C" int ps_depth_losses(const float* weights, const float* eu_bins, const float* steps_m, const float* expected_depth,
                               const float* target_depth_m, const float* sky_mask, int64_t N, int S, float pose_scale,
                               const float* pose_scale_dev, float sigma, float upper_bound, int mode, float* sums,
                               float* g_expected, float* g_weights, void* stream) {
    if (N == 0) return 0;
    PS_REQUIRE(target_depth_m && sums, "depth_losses: null pointer");
    PS_REQUIRE(weights == nullptr || ((eu_bins != nullptr) != (steps_m != nullptr)),
               "depth_losses: the line-of-sight term needs exactly one of eu_bins / steps_m");
    PS_REQUIRE(weights != nullptr || g_weights == nullptr, "depth_losses: g_weights without weights");
    PS_REQUIRE(expected_depth != nullptr || g_expected == nullptr, "depth_losses: g_expected without expected_depth");
    PS_REQUIRE(weights == nullptr || S >= 1, "depth_losses: samples per ray %d < 1", S);
    PS_REQUIRE((pose_scale_dev != nullptr || pose_scale > 0.f) && upper_bound > 0.f,
               "depth_losses: pose scale and upper bound must be positive");
    PS_REQUIRE(weights == nullptr || sigma > 0.f, "depth_losses: sigma must be positive");
    PS_REQUIRE(mode == 0 || mode == 1, "depth_losses: mode %d not in {0 normalised, 1 inverse}", mode);
    depth_losses_kernel<<<(unsigned)cdiv(N, kDepthWarps), kDepthWarps * 32, 0, (cudaStream_t)stream>>>(
        weights, eu_bins, steps_m, expected_depth, target_depth_m, sky_mask, N, S, pose_scale, pose_scale_dev, sigma,
        upper_bound, mode, sums, g_expected, g_weights);
    return check_launch("depth_losses");
}
