// Per-ray core of the depth-supervision terms of the loss dict (the rows after the proposal / distortion /
// rendered-output losses).  Reference: model_components/PreSight/losses.py
//   :24-25   normalize_depth            clip(d / upper_bound, 0, 1)
//   :28-65   line_of_sight_loss         Urban-Radiance-Fields term on the final level's weights
//   :67-81   expected_depth_loss        LiDAR supervision of the rendered expected depth
//   :83-103  expected_monodepth_loss    mono-depth supervision (optionally on 1 / (d + 5))
// and their call sites models/PreSight/nerfacto_nusc_ms.py:577-629 (sample mid-points and the rendered depth divided by
// the pose scale factor, i.e. brought to metres).  All three are means over the rays that pass the depth mask
//   target > 1 m  and  target < upper_bound  [and  sky == 0];
// this core produces one ray's summands and gradients, the caller divides by the number of masked rays.
// Plain C++ shared by the CUDA kernel (csrc/depth_loss.cu) and the host harness of tests/test_depth_loss_host.py.
#pragma once

#ifndef PS_HD
#define PS_HD
#endif
#ifndef PS_EXPF
#include <cmath>
#define PS_EXPF(a) (std::exp(a))
#define PS_LOGF(a) (std::log(a))
#endif

namespace ps {
namespace depthloss {

enum Mode { kNormalized = 0, kInverse = 1 };

PS_HD inline bool ray_supervised(float target_m, float upper_bound, const float* sky, long long n) {
    return target_m > 1.f && target_m < upper_bound && (sky == nullptr || sky[n] == 0.f);
}

PS_HD inline float clip01(float v) { return v < 0.f ? 0.f : (v > 1.f ? 1.f : v); }

// squared error of the (mapped) expected depth and its derivative w.r.t. the predicted depth in metres
PS_HD inline float expected_depth_term(float target_m, float pred_m, float upper_bound, int mode, float& d_pred) {
    if (mode == kInverse) {
        const float t = 1.f / (target_m + 5.f), p = 1.f / (pred_m + 5.f);
        const float err = t - p;
        d_pred = 2.f * err * p * p;                        // d/d pred of (t - 1/(pred+5))^2
        return err * err;
    }
    const float q = pred_m / upper_bound;
    const float err = clip01(target_m / upper_bound) - clip01(q);
    d_pred = (q >= 0.f && q <= 1.f) ? -2.f * err / upper_bound : 0.f;   // clamp passes the gradient on [min, max]
    return err * err;
}

// constants of the target distribution N(0, sigma / 3) (URF_SIGMA_SCALE_FACTOR = 3)
struct LosConsts {
    float sigma, two_var, log_norm;
};
PS_HD inline LosConsts los_consts(float sigma) {
    LosConsts c;
    const float std_ = sigma / 3.f;
    c.sigma = sigma;
    c.two_var = 2.f * (std_ * std_);
    c.log_norm = PS_LOGF(std_) + 0.91893853320467274178f;      // log(std) + log(sqrt(2 pi))
    return c;
}

// one sample's line-of-sight summand and its derivative w.r.t. the sample's weight
PS_HD inline float los_term(float w, float step_m, float target_m, const LosConsts& c, float& d_w) {
    const float lo = target_m - c.sigma, hi = target_m + c.sigma;
    float v = 0.f;
    d_w = 0.f;
    if (step_m <= hi && step_m >= lo) {
        const float x = step_m - target_m;
        const float g = PS_EXPF(-(x * x) / c.two_var - c.log_norm);
        const float e = w - g;
        v = e * e;
        d_w = 2.f * e;
    }
    if (step_m < lo) {
        v += w * w;
        d_w += 2.f * w;
    }
    return v;
}

}  // namespace depthloss
}  // namespace ps
