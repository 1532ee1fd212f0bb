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

// scatter the (already weighted) gradient of the two x-corners of one (y,z) pair.  The hash multiplies x by 1, so for an
// even floor coordinate the two corners are rows r and r^1 — one aligned slot of 2F floats, ONE vector reduction.
// (Corners whose interpolation weight is exactly zero — ceil == floor — contribute an exact 0.0, which leaves the table
// value unchanged, so they are not special-cased.)
template <int F>
__device__ __forceinline__ void scatter_xpair(float* __restrict__ lg, uint32_t r_hi, uint32_t r_lo,
                                              const float (&v_hi)[F], const float (&v_lo)[F]) {
    if constexpr (F <= 2) {
        if ((r_hi ^ r_lo) == 1u) {
            const uint32_t base = r_hi & ~1u;
            const bool hi_first = !(r_hi & 1u);
            if constexpr (F == 1)
                red_add_v2(lg + base, hi_first ? v_hi[0] : v_lo[0], hi_first ? v_lo[0] : v_hi[0]);
            else
                red_add_v4(lg + (size_t)base * 2, hi_first ? v_hi[0] : v_lo[0], hi_first ? v_hi[1] : v_lo[1],
                           hi_first ? v_lo[0] : v_hi[0], hi_first ? v_lo[1] : v_hi[1]);
            return;
        }
    }
    scatter_row<F>(lg, r_hi, v_hi, 1.f);
    scatter_row<F>(lg, r_lo, v_lo, 1.f);
}

// Scatter-add one level's feature gradient `g[F]` of this lane's point into the level's gradient table `lg`
// (all 32 lanes must call; `valid` = this lane carries a real point).
//  * warp pre-aggregation: consecutive samples of a ray that fall into the same grid cell hit the same 8 rows; they are
//    summed inside the warp and the first lane of each run issues the atomics.  The cell is identified by its floor
//    coordinates plus the three "exact integer" flags (ceil = floor + !exact);
//  * x-pair merge (see scatter_xpair).
template <int F>
__device__ __forceinline__ void scatter_level_preagg(float* __restrict__ lg, const Corner8& c, float px, float py,
                                                     float pz, float scale, const float (&g)[F], bool valid, int lane) {
    float w[8];
    corner_weights(c.ox, c.oy, c.oz, w);
    float v[8][F];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int f = 0; f < F; ++f) v[k][f] = valid ? g[f] * w[k] : 0.f;
    const int kx = (int)floorf(__fmul_rn(px, scale)), ky = (int)floorf(__fmul_rn(py, scale)),
              kz = (int)floorf(__fmul_rn(pz, scale));
    const int kf = (c.ox == 0.f ? 1 : 0) | (c.oy == 0.f ? 2 : 0) | (c.oz == 0.f ? 4 : 0) | (valid ? 0 : 8);
    // (shuffles executed unconditionally by all 32 lanes, compared afterwards)
    const int nx = __shfl_up_sync(0xffffffffu, kx, 1), ny = __shfl_up_sync(0xffffffffu, ky, 1),
              nz = __shfl_up_sync(0xffffffffu, kz, 1), nf = __shfl_up_sync(0xffffffffu, kf, 1);
    const bool same_prev = lane > 0 && nx == kx && ny == ky && nz == kz && nf == kf;
    const uint32_t heads = __ballot_sync(0xffffffffu, !same_prev);
    bool issue = valid;
    if (heads != 0xffffffffu) {   // warp-uniform: at least one run of length > 1
        // Segmented reduction into each run's first lane (all 8F values per step).  Only the butterfly
        // steps the warp's longest run needs are executed: beyond it no lane has a same-run partner, so the skipped
        // steps would add nothing (fine levels mostly have runs of 2: one step instead of five).
        const int head_pos = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
        const int max_pos = __reduce_max_sync(0xffffffffu, lane - head_pos);
        for (int o = 1; o <= max_pos; o <<= 1) {
            const bool same = (lane + o < 32) && ((((heads >> lane) >> 1) & ((1u << o) - 1u)) == 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    const float other = __shfl_down_sync(0xffffffffu, v[k][f], o);
                    if (same) v[k][f] += other;
                }
        }
        issue = valid && !same_prev;
    }
    if (issue) {
        // (y,z) corner pairs in reference order: {x-ceil, x-floor} = {h0,h3}, {h1,h2}, {h4,h7}, {h5,h6}
        scatter_xpair<F>(lg, c.row[0], c.row[3], v[0], v[3]);
        scatter_xpair<F>(lg, c.row[1], c.row[2], v[1], v[2]);
        scatter_xpair<F>(lg, c.row[4], c.row[7], v[4], v[7]);
        scatter_xpair<F>(lg, c.row[5], c.row[6], v[5], v[6]);
    }
}

}  // namespace ps
