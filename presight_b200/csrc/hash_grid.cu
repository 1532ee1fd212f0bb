// Kernel #1: stand-alone multiresolution hash encoding, forward / backward / index probe.
// Thread mapping: one thread per (point, level), level fastest — the [P, L*F] output row of a
// point is written by L consecutive lanes, so stores (fwd) and dout loads (bwd) are fully
// coalesced for any L, while each lane keeps 8 independent vector gathers in flight.
#include "hash_grid.cuh"

namespace ps {

template <int F>
__global__ void __launch_bounds__(256) hash_fwd_kernel(const float* __restrict__ x, int64_t P,
                                                       const float* __restrict__ table, HashParams hp,
                                                       float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * hp.L) return;
    const int64_t p = i / hp.L;
    const int l = (int)(i - p * hp.L);
    const float px = __ldg(x + 3 * p), py = __ldg(x + 3 * p + 1), pz = __ldg(x + 3 * p + 2);
    const uint32_t mask = (1u << hp.log2_T) - 1u;
    const Corner8 c = hash_corners(px, py, pz, hp.scale[l], mask);
    const float* lt = table + ((size_t)l << hp.log2_T) * F;
    float v[8][F];
#pragma unroll
    for (int k = 0; k < 8; ++k) gather_row<F>(lt, c.row[k], v[k]);
    float o[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
        const float t[8] = {v[0][f], v[1][f], v[2][f], v[3][f], v[4][f], v[5][f], v[6][f], v[7][f]};
        o[f] = trilerp_ref(t, c.ox, c.oy, c.oz);
    }
    float* dst = out + i * F;
    if constexpr (F == 1) {
        dst[0] = o[0];
    } else if constexpr (F == 2) {
        *reinterpret_cast<float2*>(dst) = make_float2(o[0], o[1]);
    } else if constexpr (F == 4) {
        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
}

template <int F, bool WITH_DX>
__global__ void __launch_bounds__(256) hash_bwd_kernel(const float* __restrict__ x, int64_t P,
                                                       const float* __restrict__ table, HashParams hp,
                                                       const float* __restrict__ dout, float* __restrict__ dtable,
                                                       float* __restrict__ dx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * hp.L) return;
    const int64_t p = i / hp.L;
    const int l = (int)(i - p * hp.L);
    const float px = __ldg(x + 3 * p), py = __ldg(x + 3 * p + 1), pz = __ldg(x + 3 * p + 2);
    const uint32_t mask = (1u << hp.log2_T) - 1u;
    const float scale = hp.scale[l];
    const Corner8 c = hash_corners(px, py, pz, scale, mask);
    float g[F];
    const float* src = dout + i * F;
    if constexpr (F == 1) {
        g[0] = __ldg(src);
    } else if constexpr (F == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(src));
        g[0] = t.x; g[1] = t.y;
    } else {
#pragma unroll
        for (int q = 0; q < F / 4; ++q) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(src) + q);
            g[4 * q] = t.x; g[4 * q + 1] = t.y; g[4 * q + 2] = t.z; g[4 * q + 3] = t.w;
        }
    }
    float w[8];
    corner_weights(c.ox, c.oy, c.oz, w);
    float* lg = dtable + ((size_t)l << hp.log2_T) * F;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        // a zero weight (exact-integer coordinate: ceil == floor) contributes nothing; skip the atomic
        if (w[k] != 0.f) scatter_row<F>(lg, c.row[k], g, w[k]);
    }
    if constexpr (WITH_DX) {
        // d out / d offset, then d offset / d x = scale (floor/ceil have zero gradient)
        const float* lt = table + ((size_t)l << hp.log2_T) * F;
        float v[8][F];
#pragma unroll
        for (int k = 0; k < 8; ++k) gather_row<F>(lt, c.row[k], v[k]);
        const float ox = c.ox, oy = c.oy, oz = c.oz, mx = 1.f - ox, my = 1.f - oy, mz = 1.f - oz;
        float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const float f03 = v[0][f] * ox + v[3][f] * mx, f12 = v[1][f] * ox + v[2][f] * mx;
            const float f56 = v[5][f] * ox + v[6][f] * mx, f47 = v[4][f] * ox + v[7][f] * mx;
            const float f0312 = f03 * oy + f12 * my, f4756 = f47 * oy + f56 * my;
            gz += g[f] * (f0312 - f4756);
            gy += g[f] * (oz * (f03 - f12) + mz * (f47 - f56));
            gx += g[f] * (oz * (oy * (v[0][f] - v[3][f]) + my * (v[1][f] - v[2][f])) +
                          mz * (oy * (v[4][f] - v[7][f]) + my * (v[5][f] - v[6][f])));
        }
        atomicAdd(dx + 3 * p, gx * scale);
        atomicAdd(dx + 3 * p + 1, gy * scale);
        atomicAdd(dx + 3 * p + 2, gz * scale);
    }
}

__global__ void __launch_bounds__(256) hash_indices_kernel(const float* __restrict__ x, int64_t P, HashParams hp,
                                                           int64_t* __restrict__ idx, float* __restrict__ offset) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * hp.L) return;
    const int64_t p = i / hp.L;
    const int l = (int)(i - p * hp.L);
    const uint32_t mask = (1u << hp.log2_T) - 1u;
    const Corner8 c = hash_corners(x[3 * p], x[3 * p + 1], x[3 * p + 2], hp.scale[l], mask);
#pragma unroll
    for (int k = 0; k < 8; ++k) idx[i * 8 + k] = (int64_t)c.row[k] + ((int64_t)l << hp.log2_T);
    if (offset) {
        offset[i * 3] = c.ox;
        offset[i * 3 + 1] = c.oy;
        offset[i * 3 + 2] = c.oz;
    }
}

static int fill_params(HashParams& hp, const float* scalings_host, int L, int log2_T) {
    PS_REQUIRE(L >= 1 && L <= PS_MAX_LEVELS, "hash: num_levels %d out of range [1,%d]", L, PS_MAX_LEVELS);
    PS_REQUIRE(log2_T >= 1 && log2_T <= 31, "hash: log2_hashmap_size %d out of range", log2_T);
    PS_REQUIRE(scalings_host != nullptr, "hash: scalings_host is null");
    for (int l = 0; l < L; ++l) hp.scale[l] = scalings_host[l];
    hp.L = L;
    hp.log2_T = log2_T;
    return 0;
}

}  // namespace ps

using namespace ps;

extern "C" int ps_hash_fwd(const float* x01, int64_t P, const float* table, const float* scalings_host, int L, int F,
                           int log2_T, float* out, void* stream) {
    HashParams hp;
    if (int e = fill_params(hp, scalings_host, L, log2_T)) return e;
    PS_REQUIRE(P >= 0, "hash_fwd: negative P");
    if (P == 0) return 0;
    PS_REQUIRE(x01 && table && out, "hash_fwd: null pointer");
    const int threads = 256;
    const int64_t blocks = cdiv(P * L, threads);
    PS_REQUIRE(blocks < (1ll << 31), "hash_fwd: too many points");
    cudaStream_t s = (cudaStream_t)stream;
    switch (F) {
        case 1: hash_fwd_kernel<1><<<(unsigned)blocks, threads, 0, s>>>(x01, P, table, hp, out); break;
        case 2: hash_fwd_kernel<2><<<(unsigned)blocks, threads, 0, s>>>(x01, P, table, hp, out); break;
        case 4: hash_fwd_kernel<4><<<(unsigned)blocks, threads, 0, s>>>(x01, P, table, hp, out); break;
        case 8: hash_fwd_kernel<8><<<(unsigned)blocks, threads, 0, s>>>(x01, P, table, hp, out); break;
        default: PS_REQUIRE(false, "hash_fwd: features_per_level %d not in {1,2,4,8}", F);
    }
    return check_launch("hash_fwd");
}

extern "C" int ps_hash_bwd(const float* x01, int64_t P, const float* table, const float* scalings_host, int L, int F,
                           int log2_T, const float* dout, float* dtable, float* dx, void* stream) {
    HashParams hp;
    if (int e = fill_params(hp, scalings_host, L, log2_T)) return e;
    PS_REQUIRE(P >= 0, "hash_bwd: negative P");
    if (P == 0) return 0;
    PS_REQUIRE(x01 && dout && dtable, "hash_bwd: null pointer");
    PS_REQUIRE(dx == nullptr || table != nullptr, "hash_bwd: dx requested but table is null");
    const int threads = 256;
    const int64_t blocks = cdiv(P * L, threads);
    PS_REQUIRE(blocks < (1ll << 31), "hash_bwd: too many points");
    cudaStream_t s = (cudaStream_t)stream;
#define PS_LAUNCH_BWD(FF)                                                                                     \
    if (dx)                                                                                                   \
        hash_bwd_kernel<FF, true><<<(unsigned)blocks, threads, 0, s>>>(x01, P, table, hp, dout, dtable, dx);  \
    else                                                                                                      \
        hash_bwd_kernel<FF, false><<<(unsigned)blocks, threads, 0, s>>>(x01, P, table, hp, dout, dtable, dx);
    switch (F) {
        case 1: PS_LAUNCH_BWD(1) break;
        case 2: PS_LAUNCH_BWD(2) break;
        case 4: PS_LAUNCH_BWD(4) break;
        case 8: PS_LAUNCH_BWD(8) break;
        default: PS_REQUIRE(false, "hash_bwd: features_per_level %d not in {1,2,4,8}", F);
    }
#undef PS_LAUNCH_BWD
    return check_launch("hash_bwd");
}

extern "C" int ps_hash_indices(const float* x01, int64_t P, const float* scalings_host, int L, int log2_T,
                               int64_t* idx, float* offset, void* stream) {
    HashParams hp;
    if (int e = fill_params(hp, scalings_host, L, log2_T)) return e;
    if (P == 0) return 0;
    PS_REQUIRE(x01 && idx, "hash_indices: null pointer");
    const int threads = 256;
    hash_indices_kernel<<<(unsigned)cdiv(P * L, threads), threads, 0, (cudaStream_t)stream>>>(x01, P, hp, idx, offset);
    return check_launch("hash_indices");
}
