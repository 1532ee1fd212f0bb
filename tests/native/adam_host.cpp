// Host harness around presight_b200/csrc/adam_core.h (the per-element code of ps_adam_step) for tests/test_adam_host.py.
#include <cmath>
#include <cstdint>

#include "../../presight_b200/csrc/adam_core.h"

extern "C" void adam_host(float* p, const float* g, float* m, float* v, int64_t n, double lr, double b1, double b2,
                          double eps, double weight_decay, int64_t step) {
    ps::adam::Scalars s;       // same derivation as ps_adam_step (csrc/adam.cu)
    s.weight_decay = (float)weight_decay;
    s.one_minus_beta1 = (float)(1.0 - b1);
    s.beta2 = (float)b2;
    s.one_minus_beta2 = (float)(1.0 - b2);
    s.step_size = (float)(lr / (1.0 - std::pow(b1, (double)step)));
    s.bias2_sqrt = (float)std::sqrt(1.0 - std::pow(b2, (double)step));
    s.eps = (float)eps;
    for (int64_t i = 0; i < n; ++i) ps::adam::update(p[i], g[i], m[i], v[i], s);
}
