// Host harness around presight_b200/csrc/depth_loss_core.h (the per-ray code of ps_depth_losses) for
// tests/test_depth_loss_host.py: same loop structure as the kernel (csrc/depth_loss.cu), one ray at a time.
#include <cstdint>

#include "../../presight_b200/csrc/depth_loss_core.h"

extern "C" void depth_losses_host(const float* weights, const float* steps_m, const float* expected, const float* target,
                                  const float* sky, int64_t N, int S, float pose_scale, float sigma, float upper_bound,
                                  int mode, double* sums, float* g_expected, float* g_weights) {
    using namespace ps::depthloss;
    const LosConsts c = los_consts(sigma);
    for (int64_t n = 0; n < N; ++n) {
        const float t = target[n];
        const bool on = ray_supervised(t, upper_bound, sky, n);
        if (on) sums[0] += 1.0;
        if (expected) {
            float d_pred = 0.f;
            const float v = expected_depth_term(t, expected[n] / pose_scale, upper_bound, mode, d_pred);
            if (on) sums[1] += v;
            if (g_expected) g_expected[n] = on ? d_pred / pose_scale : 0.f;
        }
        if (weights) {
            for (int s = 0; s < S; ++s) {
                float d_w = 0.f;
                const float v = los_term(weights[n * S + s], steps_m[n * S + s], t, c, d_w);
                if (on) sums[2] += v;
                if (g_weights) g_weights[n * S + s] = on ? d_w : 0.f;
            }
        }
    }
}
