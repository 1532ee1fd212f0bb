// Host harness around presight_b200/csrc/zaa_core.h: the per-ray code of the z-anti-aliased interlevel loss compiled
// with g++ so that tests/test_zaa_host.py can check it against the live reference's fixture without a GPU.
// (Test infrastructure; the product reaches the same code only through the CUDA kernel in csrc/losses.cu.)
#include <cstdint>

#include "../../presight_b200/csrc/zaa_core.h"

extern "C" int zaa_loss_host(const float* c, const float* w, int64_t N, int S, const float* cp, const float* wp, int Sp,
                             double r, double* loss_sum, float* grad_wp) {
    if (S < 1 || S > ps::zaa::kMaxS || Sp < 1) return 1;
    double total = 0.0;
    for (int64_t n = 0; n < N; ++n)
        total += (double)ps::zaa::ray_loss(c + n * (S + 1), w + n * S, S, cp + n * (Sp + 1), wp + n * Sp, Sp, r,
                                           grad_wp ? grad_wp + n * Sp : nullptr);
    *loss_sum = total;
    return 0;
}
