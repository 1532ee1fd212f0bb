// Device-side pieces of the multiresolution hash encoding, shared by the stand-alone
// kernels (hash_grid.cu) and the fused per-level kernels.
//
// Arithmetic contract (reference: field_components/encodings.py:343-384, oracle.hash_encode):
//   scaled = x * scale_l            (fp32 multiply, never contracted into an FMA)
//   c = ceil(scaled), f = floor(scaled) taken independently (c == f on exact integers)
//   o = scaled - f
//   row = ((ix*1) ^ (iy*2654435761) ^ (iz*805459861)) mod 2^k + l*2^k   in uint32 wrap-around
//   corners h0..h7 = ccc, cfc, ffc, fcc, ccf, cff, fff, fcf
//   lerp along x (03,12,56,47), then y, then z, each `hi*o + lo*(1-o)` with separately rounded products.
#pragma once
#include "common.cuh"

namespace ps {

struct HashParams {
    float scale[PS_MAX_LEVELS];
    int L;
    int log2_T;
};

constexpr uint32_t kPrimeY = 2654435761u;
constexpr uint32_t kPrimeZ = 805459861u;

struct Corner8 {
    uint32_t row[8];   // row within the level (0..T-1), reference corner order
    float ox, oy, oz;  // fractional offsets
};

__device__ __forceinline__ Corner8 hash_corners(float px, float py, float pz, float scale, uint32_t mask) {
    Corner8 c;
    const float sx = __fmul_rn(px, scale), sy = __fmul_rn(py, scale), sz = __fmul_rn(pz, scale);
    const float fx = floorf(sx), fy = floorf(sy), fz = floorf(sz);
    const float cx = ceilf(sx), cy = ceilf(sy), cz = ceilf(sz);
    c.ox = __fsub_rn(sx, fx);
    c.oy = __fsub_rn(sy, fy);
    c.oz = __fsub_rn(sz, fz);
    // float -> int32 -> uint32: two's complement low bits match the reference's int64 arithmetic mod 2^k
    const uint32_t xf = (uint32_t)(int)fx, xc = (uint32_t)(int)cx;
    const uint32_t yf = (uint32_t)(int)fy * kPrimeY, yc = (uint32_t)(int)cy * kPrimeY;
    const uint32_t zf = (uint32_t)(int)fz * kPrimeZ, zc = (uint32_t)(int)cz * kPrimeZ;
    c.row[0] = (xc ^ yc ^ zc) & mask;
    c.row[1] = (xc ^ yf ^ zc) & mask;
    c.row[2] = (xf ^ yf ^ zc) & mask;
    c.row[3] = (xf ^ yc ^ zc) & mask;
    c.row[4] = (xc ^ yc ^ zf) & mask;
    c.row[5] = (xc ^ yf ^ zf) & mask;
    c.row[6] = (xf ^ yf ^ zf) & mask;
    c.row[7] = (xf ^ yc ^ zf) & mask;
    return c;
}

// hi*o + lo*(1-o) exactly as torch evaluates it (two rounded products, one rounded add)
__device__ __forceinline__ float lerp_ref(float hi, float lo, float o, float one_minus_o) {
    return __fadd_rn(__fmul_rn(hi, o), __fmul_rn(lo, one_minus_o));
}

// trilinear blend of the 8 corner values of one feature channel in reference order
__device__ __forceinline__ float trilerp_ref(const float (&t)[8], float ox, float oy, float oz) {
    const float mx = __fsub_rn(1.f, ox), my = __fsub_rn(1.f, oy), mz = __fsub_rn(1.f, oz);
    const float f03 = lerp_ref(t[0], t[3], ox, mx);
    const float f12 = lerp_ref(t[1], t[2], ox, mx);
    const float f56 = lerp_ref(t[5], t[6], ox, mx);
    const float f47 = lerp_ref(t[4], t[7], ox, mx);
    const float f0312 = lerp_ref(f03, f12, oy, my);
    const float f4756 = lerp_ref(f47, f56, oy, my);
    return lerp_ref(f0312, f4756, oz, mz);
}

// gradient weights of the 8 corners for an upstream gradient of 1 (autograd order: z, then y, then x)
__device__ __forceinline__ void corner_weights(float ox, float oy, float oz, float (&w)[8]) {
    const float mx = 1.f - ox, my = 1.f - oy, mz = 1.f - oz;
    const float z1y1 = oz * oy, z1y0 = oz * my, z0y1 = mz * oy, z0y0 = mz * my;
    w[0] = z1y1 * ox;  // f03 <- f0312 <- out
    w[3] = z1y1 * mx;
    w[1] = z1y0 * ox;  // f12
    w[2] = z1y0 * mx;
    w[4] = z0y1 * ox;  // f47 <- f4756
    w[7] = z0y1 * mx;
    w[5] = z0y0 * ox;  // f56
    w[6] = z0y0 * mx;
}

template <int F>
struct FeatVec;
template <>
struct FeatVec<1> {
    using T = float;
};
template <>
struct FeatVec<2> {
    using T = float2;
};
template <>
struct FeatVec<4> {
    using T = float4;
};

// gather one corner's F features (vectorised: 4/8/16-byte loads through the read-only path)
template <int F>
__device__ __forceinline__ void gather_row(const float* __restrict__ level_table, uint32_t row, float (&v)[F]) {
    if constexpr (F == 1) {
        v[0] = __ldg(level_table + row);
    } else if constexpr (F == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(level_table) + row);
        v[0] = t.x;
        v[1] = t.y;
    } else if constexpr (F == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(level_table) + row);
        v[0] = t.x;
        v[1] = t.y;
        v[2] = t.z;
        v[3] = t.w;
    } else {
        static_assert(F == 8, "F must be 1, 2, 4 or 8");
        const float4 a = __ldg(reinterpret_cast<const float4*>(level_table) + 2 * (size_t)row);
        const float4 b = __ldg(reinterpret_cast<const float4*>(level_table) + 2 * (size_t)row + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
}

// scatter-add one corner's F gradient values with the widest reduction available
template <int F>
__device__ __forceinline__ void scatter_row(float* __restrict__ level_grad, uint32_t row, const float (&g)[F], float w) {
    float* dst = level_grad + (size_t)row * F;
    if constexpr (F == 1) {
        red_add(dst, g[0] * w);
    } else if constexpr (F == 2) {
        red_add_v2(dst, g[0] * w, g[1] * w);
    } else if constexpr (F == 4) {
        red_add_v4(dst, g[0] * w, g[1] * w, g[2] * w, g[3] * w);
    } else {
        red_add_v4(dst, g[0] * w, g[1] * w, g[2] * w, g[3] * w);
        red_add_v4(dst + 4, g[4] * w, g[5] * w, g[6] * w, g[7] * w);
    }
}

}  // namespace ps
