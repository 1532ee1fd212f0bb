// Device functions for the position prologue (shared with the fused per-level kernels).
#pragma once
#include "common.cuh"

namespace ps {

struct Aabb {
    float lo[3];
    float hi[3];
};

// World position -> unit-cube position in place; returns the selector (all 0 < p < 1).
// Masked points are moved to the origin exactly like `positions * selector[..., None]`
// (fields/PreSight/ingp_field.py:169-177).  Contraction: spatial_distortions.py:66-69, order = inf.
__device__ __forceinline__ bool normalize_point(float (&v)[3], const Aabb& box, bool contract) {
    if (contract) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float t = __fdiv_rn(__fsub_rn(v[k], box.lo[k]), __fsub_rn(box.hi[k], box.lo[k]));  // 0..1
            v[k] = __fsub_rn(__fmul_rn(t, 2.f), 1.f);                                                 // -1..1
        }
        const float mag = fmaxf(fabsf(v[0]), fmaxf(fabsf(v[1]), fabsf(v[2])));
        if (!(mag < 1.f)) {
            const float k = __fsub_rn(2.f, __fdiv_rn(1.f, mag));
#pragma unroll
            for (int q = 0; q < 3; ++q) v[q] = __fmul_rn(k, __fdiv_rn(v[q], mag));
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) v[q] = __fdiv_rn(__fadd_rn(v[q], 2.f), 4.f);
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = __fdiv_rn(__fsub_rn(v[k], box.lo[k]), __fsub_rn(box.hi[k], box.lo[k]));
    }
    const bool inside = v[0] > 0.f && v[0] < 1.f && v[1] > 0.f && v[1] < 1.f && v[2] > 0.f && v[2] < 1.f;
    if (!inside) {
        // p * 0 keeps the sign of zero and NaN; the hash only sees floor/ceil of it
        v[0] = __fmul_rn(v[0], 0.f);
        v[1] = __fmul_rn(v[1], 0.f);
        v[2] = __fmul_rn(v[2], 0.f);
    }
    return inside;
}

// cameras/rays.py:56: origins + directions * (starts + ends) / 2, evaluated left to right
__device__ __forceinline__ void frustum_midpoint(const float* o, const float* d, float start, float end,
                                                 float (&out)[3]) {
    const float se = __fadd_rn(start, end);
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] = __fadd_rn(o[k], __fdiv_rn(__fmul_rn(d[k], se), 2.f));
}

// utils/math.py:27-74 on the (d+1)/2-mapped direction (fields/base_field.py:136-142)
__device__ __forceinline__ void sh4_of_mapped(float x, float y, float z, float (&c)[16]) {
    const float xx = x * x, yy = y * y, zz = z * z;
    c[0] = 0.28209479177387814f;
    c[1] = 0.4886025119029199f * y;
    c[2] = 0.4886025119029199f * z;
    c[3] = 0.4886025119029199f * x;
    c[4] = 1.0925484305920792f * x * y;
    c[5] = 1.0925484305920792f * y * z;
    c[6] = 0.9461746957575601f * zz - 0.31539156525251999f;
    c[7] = 1.0925484305920792f * x * z;
    c[8] = 0.5462742152960396f * (xx - yy);
    c[9] = 0.5900435899266435f * y * (3.f * xx - yy);
    c[10] = 2.890611442640554f * x * y * z;
    c[11] = 0.4570457994644658f * y * (5.f * zz - 1.f);
    c[12] = 0.3731763325901154f * z * (5.f * zz - 3.f);
    c[13] = 0.4570457994644658f * x * (5.f * zz - 1.f);
    c[14] = 1.445305721320277f * z * (xx - yy);
    c[15] = 0.5900435899266435f * x * (xx - 3.f * yy);
}
__device__ __forceinline__ void sh4_of_direction(float dx, float dy, float dz, float (&c)[16]) {
    sh4_of_mapped((dx + 1.f) / 2.f, (dy + 1.f) / 2.f, (dz + 1.f) / 2.f, c);
}

}  // namespace ps
