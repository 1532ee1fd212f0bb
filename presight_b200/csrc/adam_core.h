// Per-element core of the Adam update the reference trains with: torch.optim.Adam(lr=1e-2, eps=1e-15, weight_decay=1e-5)
// (configs/method_configs.py:115,121; engine/optimizers.py:133-140), i.e. torch's `_single_tensor_adam` without amsgrad:
//   g   = grad + weight_decay * p
//   m   = m + (g - m) * (1 - beta1)                     (Tensor.lerp_)
//   v   = v * beta2 + (1 - beta2) * g * g               (mul_ + addcmul_)
//   p   = p - step_size * m / (sqrt(v) / sqrt(1 - beta2^t) + eps),   step_size = lr / (1 - beta1^t)
// The scalars are computed by the host in double precision, as torch does, and passed as fp32.  Plain C++ shared by the
// CUDA kernel (csrc/adam.cu) and the host harness of tests/test_adam_host.py.
#pragma once

#ifndef PS_HD
#define PS_HD
#endif
#ifndef PS_SQRTF
#include <cmath>
#define PS_SQRTF(a) (std::sqrt(a))
#endif

namespace ps {
namespace adam {

struct Scalars {
    float weight_decay, one_minus_beta1, beta2, one_minus_beta2, step_size, bias2_sqrt, eps;
};

PS_HD inline void update(float& p, float grad, float& m, float& v, const Scalars& s) {
    const float g = s.weight_decay != 0.f ? grad + s.weight_decay * p : grad;
    m = m + (g - m) * s.one_minus_beta1;
    v = v * s.beta2 + s.one_minus_beta2 * g * g;
    const float denom = PS_SQRTF(v) / s.bias2_sqrt + s.eps;
    p = p - s.step_size * (m / denom);
}

}  // namespace adam
}  // namespace ps
