// Per-ray core of pinhole ray generation (no lens distortion): RayGenerator.forward
// (model_components/ray_generators.py:43-61) -> Cameras._generate_rays_from_coords (cameras/cameras.py:497-880,
// PERSPECTIVE branch :773-779).  Plain C++ shared by the CUDA kernel (csrc/raygen.cu) and the host harness of
// tests/test_raygen_host.py, which checks it against the live reference's fixture.
//
//   (y, x)   = (row, col) + pixel_offset
//   d0       = R [ (x - cx) / fx, -(y - cy) / fy, -1 ],  d1: x + 1,  d2: y + 1      (R = c2w[:, :3], rows dotted with d)
//   each normalised by max(|d|, 4 eps_f64);  pixel_area = |d0 - d1| * |d0 - d2|
// PS_MUL / PS_ADD / PS_SUB / PS_DIV / PS_SQRT are the un-fused fp32 operations (`__fmul_rn` ... on the device): torch
// evaluates every product and sum separately, and pixel_area is a difference of nearly equal unit vectors, so a fused
// multiply-add in the rotation would show up in its fourth digit.
#pragma once
#include <cstdint>

#ifndef PS_HD
#define PS_HD
#endif
#ifndef PS_MUL
#include <cmath>
#define PS_MUL(a, b) ((a) * (b))
#define PS_ADD(a, b) ((a) + (b))
#define PS_SUB(a, b) ((a) - (b))
#define PS_DIV(a, b) ((a) / (b))
#define PS_SQRT(a) (std::sqrt(a))
#endif

namespace ps {
namespace raygen {

PS_HD inline void unit_dir(const float* rot34, float u, float v, float (&d)[3], float& norm) {
    // camera-space direction (u, v, -1) rotated into the world: d[i] = R[i][0] u + R[i][1] v + R[i][2] (-1)
    float q[3];
    for (int i = 0; i < 3; ++i)
        q[i] = PS_ADD(PS_ADD(PS_MUL(u, rot34[4 * i]), PS_MUL(v, rot34[4 * i + 1])), PS_MUL(-1.f, rot34[4 * i + 2]));
    const float n2 = PS_ADD(PS_ADD(PS_MUL(q[0], q[0]), PS_MUL(q[1], q[1])), PS_MUL(q[2], q[2]));
    float n = PS_SQRT(n2);
    const float tiny = 8.8817841970012523e-16f;          // 4 * eps(float64), camera_utils.py:30,299
    n = n > tiny ? n : tiny;
    for (int i = 0; i < 3; ++i) d[i] = PS_DIV(q[i], n);
    norm = n;
}

// c2w [3][4] row-major of this ray's camera; row / col: integer pixel; -> origin[3], dir[3], pixel_area, dir_norm
PS_HD inline void pinhole_ray(const float* c2w, float fx, float fy, float cx, float cy, int64_t row, int64_t col,
                              float pixel_offset, float* origin, float* dir, float* pixel_area, float* dir_norm) {
    const float y = PS_ADD((float)row, pixel_offset), x = PS_ADD((float)col, pixel_offset);
    const float xc = PS_SUB(x, cx), yc = PS_SUB(y, cy);
    const float u0 = PS_DIV(xc, fx), v0 = -PS_DIV(yc, fy);
    const float u1 = PS_DIV(PS_ADD(xc, 1.f), fx), v1 = -PS_DIV(PS_ADD(yc, 1.f), fy);
    float d0[3], d1[3], d2[3], n0, n1, n2;
    unit_dir(c2w, u0, v0, d0, n0);
    unit_dir(c2w, u1, v0, d1, n1);
    unit_dir(c2w, u0, v1, d2, n2);
    float sx = 0.f, sy = 0.f;
    for (int i = 0; i < 3; ++i) {
        const float ex = PS_SUB(d0[i], d1[i]), ey = PS_SUB(d0[i], d2[i]);
        sx = PS_ADD(sx, PS_MUL(ex, ex));
        sy = PS_ADD(sy, PS_MUL(ey, ey));
    }
    for (int i = 0; i < 3; ++i) {
        origin[i] = c2w[4 * i + 3];
        dir[i] = d0[i];
    }
    *pixel_area = PS_MUL(PS_SQRT(sx), PS_SQRT(sy));
    if (dir_norm) *dir_norm = n0;
}

}  // namespace raygen
}  // namespace ps
