// Host harness around presight_b200/csrc/raygen_core.h (the per-ray code of ps_generate_rays) for tests/test_raygen_host.py.
// Compiled with -ffp-contract=off so that the plain operators are the un-fused fp32 operations the device macros name.
#include <cstdint>

#include "../../presight_b200/csrc/raygen_core.h"

extern "C" int raygen_host(const float* c2w, const float* fx, const float* fy, const float* cx, const float* cy, int C,
                           const int64_t* idx, int64_t N, float pixel_offset, float* origins, float* directions,
                           float* pixel_area, float* norm) {
    for (int64_t n = 0; n < N; ++n) {
        const int64_t cam = idx[3 * n];
        if (cam < 0 || cam >= C) return 1;
        ps::raygen::pinhole_ray(c2w + cam * 12, fx[cam], fy[cam], cx[cam], cy[cam], idx[3 * n + 1], idx[3 * n + 2],
                                pixel_offset, origins + 3 * n, directions + 3 * n, pixel_area + n, norm + n);
    }
    return 0;
}
