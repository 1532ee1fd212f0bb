// Kernel #1: stand-alone multiresolution hash encoding, forward / backward / index probe.
//
// Launch shape: grid = (point blocks, level groups), level group = blockIdx.y.  CTAs are dispatched x-fastest,
// so the machine works through the table level group by level group: the live working set is LPT levels
// (chosen <= ~40 MiB) instead of the whole table, which keeps the random gathers / scatter-adds resident in the
// 126 MB L2 — DRAM sees each level's table roughly once per pass instead of one sector per corner.
// Within a CTA consecutive threads are consecutive points (= consecutive samples along a ray).
//
// x-neighbour merge: the hash multiplies x by 1, so for an even floor coordinate the two x-corners of a (y,z)
// pair are rows r and r^1 — one aligned 2F-float slot.  Backward then issues ONE vector reduction
// (red.global.add.v4.f32 for F=2, .v2 for F=1) for both corners.
#include <stdlib.h>

#include "hash_grid.cuh"

namespace ps {

// Sub-field mode (ms_route.cu): rows are grouped by sub-field in tiles of 128; tile -> sub-field -> that sub-field's table
// (forward) or gradient table (backward).  All null = the ordinary single-table call.
struct MsTables {
    const float* const* tables;
    float* const* dtables;
    const uint8_t* tile_sf;
    const int32_t* perm;
};

template <int F, int LPT>
__global__ void __launch_bounds__(256) hash_fwd_kernel(const float* __restrict__ x, int64_t P,
                                                       const float* __restrict__ table, HashParams hp,
                                                       float* __restrict__ out, int level_major, MsTables ms) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    if (ms.tile_sf) {                      // sub-field-homogeneous tiles of 128 rows: this tile's table (255 = unused tile)
        const int sf = ms.tile_sf[p >> 7];
        if (sf == 255) return;
        table = ms.tables[sf];
    }
    const int l0 = blockIdx.y * LPT;
    const float px = __ldg(x + 3 * p), py = __ldg(x + 3 * p + 1), pz = __ldg(x + 3 * p + 2);
    const uint32_t mask = (1u << hp.log2_T) - 1u;
    float o[LPT][F];
#pragma unroll
    for (int i = 0; i < LPT; ++i) {
        const int l = l0 + i;
        const Corner8 c = hash_corners(px, py, pz, hp.scale[l], mask);
        const float* lt = table + ((size_t)l << hp.log2_T) * F;
        float v[8][F];
#pragma unroll
        for (int k = 0; k < 8; ++k) gather_row<F>(lt, c.row[k], v[k]);
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const float t[8] = {v[0][f], v[1][f], v[2][f], v[3][f], v[4][f], v[5][f], v[6][f], v[7][f]};
            o[i][f] = trilerp_ref(t, c.ox, c.oy, c.oz);
        }
    }
    if (level_major) {
        // features stored [L][P][F]: consecutive threads write consecutive F-vectors of one level (coalesced)
#pragma unroll
        for (int i = 0; i < LPT; ++i) {
            float* d = out + ((size_t)(l0 + i) * P + p) * F;
            if constexpr (F == 1) {
                d[0] = o[i][0];
            } else if constexpr (F == 2) {
                *reinterpret_cast<float2*>(d) = make_float2(o[i][0], o[i][1]);
            } else {
#pragma unroll
                for (int q = 0; q < F / 4; ++q)
                    reinterpret_cast<float4*>(d)[q] = make_float4(o[i][4 * q], o[i][4 * q + 1], o[i][4 * q + 2], o[i][4 * q + 3]);
            }
        }
        return;
    }
    // LPT*F consecutive floats of this point's output row
    float* dst = out + (p * hp.L + l0) * F;
    constexpr int NV = LPT * F;
    const float* src = &o[0][0];
    if constexpr (NV % 4 == 0) {
#pragma unroll
        for (int q = 0; q < NV / 4; ++q)
            reinterpret_cast<float4*>(dst)[q] = make_float4(src[4 * q], src[4 * q + 1], src[4 * q + 2], src[4 * q + 3]);
    } else if constexpr (NV % 2 == 0) {
#pragma unroll
        for (int q = 0; q < NV / 2; ++q) reinterpret_cast<float2*>(dst)[q] = make_float2(src[2 * q], src[2 * q + 1]);
    } else {
#pragma unroll
        for (int q = 0; q < NV; ++q) dst[q] = src[q];
    }
}

template <int F, int LPT, bool WITH_DX>
__global__ void __launch_bounds__(256) hash_bwd_kernel(const float* __restrict__ x, int64_t P,
                                                       const float* __restrict__ table, HashParams hp,
                                                       const float* __restrict__ dout, float* __restrict__ dtable,
                                                       float* __restrict__ dx, int level_major, MsTables ms) {
    // every lane stays alive (full-mask shuffles below); out-of-range lanes carry zero gradient
    const int64_t pi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = pi < P;
    if (ms.tile_sf) {                      // (warp-uniform: a warp lies inside one 128-row tile)
        const int sf = ms.tile_sf[(valid ? pi : P - 1) >> 7];
        if (sf == 255) return;
        dtable = ms.dtables[sf];
        valid = valid && ms.perm[pi] >= 0; // padding rows of a sub-field's segment carry no point
    }
    const int64_t p = valid ? pi : P - 1;
    const int lane = threadIdx.x & 31;
    const int l0 = blockIdx.y * LPT;
    const float px = __ldg(x + 3 * p), py = __ldg(x + 3 * p + 1), pz = __ldg(x + 3 * p + 2);
    const uint32_t mask = (1u << hp.log2_T) - 1u;
    // LPT*F consecutive gradient values of this point
    constexpr int NV = LPT * F;
    float gin[NV];
    const float* src = dout + (p * hp.L + l0) * F;
    if (level_major) {
#pragma unroll
        for (int i = 0; i < LPT; ++i) {
            const float* d = dout + ((size_t)(l0 + i) * P + p) * F;
            if constexpr (F == 1) {
                gin[i] = __ldg(d);
            } else if constexpr (F == 2) {
                const float2 t = __ldg(reinterpret_cast<const float2*>(d));
                gin[2 * i] = t.x; gin[2 * i + 1] = t.y;
            } else {
#pragma unroll
                for (int q = 0; q < F / 4; ++q) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(d) + q);
                    gin[i * F + 4 * q] = t.x; gin[i * F + 4 * q + 1] = t.y; gin[i * F + 4 * q + 2] = t.z; gin[i * F + 4 * q + 3] = t.w;
                }
            }
        }
    } else if constexpr (NV % 4 == 0) {
#pragma unroll
        for (int q = 0; q < NV / 4; ++q) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(src) + q);
            gin[4 * q] = t.x; gin[4 * q + 1] = t.y; gin[4 * q + 2] = t.z; gin[4 * q + 3] = t.w;
        }
    } else if constexpr (NV % 2 == 0) {
#pragma unroll
        for (int q = 0; q < NV / 2; ++q) {
            const float2 t = __ldg(reinterpret_cast<const float2*>(src) + q);
            gin[2 * q] = t.x; gin[2 * q + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int q = 0; q < NV; ++q) gin[q] = __ldg(src + q);
    }
    if (!valid) {
#pragma unroll
        for (int q = 0; q < NV; ++q) gin[q] = 0.f;
    }
    float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
    for (int i = 0; i < LPT; ++i) {
        const int l = l0 + i;
        const float scale = hp.scale[l];
        const Corner8 c = hash_corners(px, py, pz, scale, mask);
        float gl[F];
#pragma unroll
        for (int f = 0; f < F; ++f) gl[f] = gin[i * F + f];
        scatter_level_preagg<F>(dtable + ((size_t)l << hp.log2_T) * F, c, px, py, pz, scale, gl, valid, lane);
        if constexpr (WITH_DX) {
            // d out / d offset, then d offset / d x = scale (floor/ceil have zero gradient)
            const float* lt = table + ((size_t)l << hp.log2_T) * F;
            float t[8][F];
#pragma unroll
            for (int k = 0; k < 8; ++k) gather_row<F>(lt, c.row[k], t[k]);
            const float ox = c.ox, oy = c.oy, oz = c.oz, mx = 1.f - ox, my = 1.f - oy, mz = 1.f - oz;
            float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const float g = gin[i * F + f];
                const float f03 = t[0][f] * ox + t[3][f] * mx, f12 = t[1][f] * ox + t[2][f] * mx;
                const float f56 = t[5][f] * ox + t[6][f] * mx, f47 = t[4][f] * ox + t[7][f] * mx;
                const float f0312 = f03 * oy + f12 * my, f4756 = f47 * oy + f56 * my;
                az += g * (f0312 - f4756);
                ay += g * (oz * (f03 - f12) + mz * (f47 - f56));
                ax += g * (oz * (oy * (t[0][f] - t[3][f]) + my * (t[1][f] - t[2][f])) +
                           mz * (oy * (t[4][f] - t[7][f]) + my * (t[5][f] - t[6][f])));
            }
            gx += ax * scale; gy += ay * scale; gz += az * scale;
        }
    }
    if constexpr (WITH_DX) {
        if (valid) {
            atomicAdd(dx + 3 * p, gx);
            atomicAdd(dx + 3 * p + 1, gy);
            atomicAdd(dx + 3 * p + 2, gz);
        }
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

// levels per thread: the largest of {8,4,2,1} dividing L whose tables (LPT levels) stay well inside L2
static int pick_lpt(int L, int F, int log2_T) {
    const double level_bytes = (double)(1ull << log2_T) * F * 4.0;
    const double budget = 40.0 * 1024 * 1024;
    for (int lpt : {8, 4, 2}) {
        if (lpt * F > 16) continue;  // register budget
        if (L % lpt == 0 && lpt * level_bytes <= budget) return lpt;
    }
    return 1;
}

}  // namespace ps

using namespace ps;

#define PS_DISPATCH_F_LPT(F, LPT, MACRO)                                   \
    switch (F) {                                                           \
        case 1:                                                            \
            switch (LPT) {                                                 \
                case 8: MACRO(1, 8) break;                                 \
                case 4: MACRO(1, 4) break;                                 \
                case 2: MACRO(1, 2) break;                                 \
                default: MACRO(1, 1) break;                                \
            }                                                              \
            break;                                                         \
        case 2:                                                            \
            switch (LPT) {                                                 \
                case 8: MACRO(2, 8) break;                                 \
                case 4: MACRO(2, 4) break;                                 \
                case 2: MACRO(2, 2) break;                                 \
                default: MACRO(2, 1) break;                                \
            }                                                              \
            break;                                                         \
        case 4:                                                            \
            switch (LPT) {                                                 \
                case 4: MACRO(4, 4) break;                                 \
                case 2: MACRO(4, 2) break;                                 \
                default: MACRO(4, 1) break;                                \
            }                                                              \
            break;                                                         \
        case 8:                                                            \
            switch (LPT) {                                                 \
                case 2: MACRO(8, 2) break;                                 \
                default: MACRO(8, 1) break;                                \
            }                                                              \
            break;                                                         \
    }

static int hash_fwd_impl(const float* x01, int64_t P, const float* table, const float* scalings_host, int L, int F,
                         int log2_T, float* out, int level_major, void* stream, MsTables ms = MsTables{}) {
    HashParams hp;
    if (int e = fill_params(hp, scalings_host, L, log2_T)) return e;
    PS_REQUIRE(F == 1 || F == 2 || F == 4 || F == 8, "hash_fwd: features_per_level %d not in {1,2,4,8}", F);
    PS_REQUIRE(P >= 0, "hash_fwd: negative P");
    if (P == 0) return 0;
    PS_REQUIRE(x01 && (table || ms.tables) && out, "hash_fwd: null pointer");
    const int threads = 256;
    const int64_t blocks = cdiv(P, threads);
    PS_REQUIRE(blocks < (1ll << 31), "hash_fwd: too many points");
    cudaStream_t s = (cudaStream_t)stream;
    const int lpt = pick_lpt(L, F, log2_T);
    const dim3 grid((unsigned)blocks, (unsigned)(L / lpt));
#define PS_FWD(FF, LL) hash_fwd_kernel<FF, LL><<<grid, threads, 0, s>>>(x01, P, table, hp, out, level_major, ms);
    PS_DISPATCH_F_LPT(F, lpt, PS_FWD)
#undef PS_FWD
    return check_launch("hash_fwd");
}

static int hash_bwd_impl(const float* x01, int64_t P, const float* table, const float* scalings_host, int L, int F,
                         int log2_T, const float* dout, float* dtable, float* dx, int level_major, void* stream,
                         MsTables ms = MsTables{}) {
    HashParams hp;
    if (int e = fill_params(hp, scalings_host, L, log2_T)) return e;
    PS_REQUIRE(F == 1 || F == 2 || F == 4 || F == 8, "hash_bwd: features_per_level %d not in {1,2,4,8}", F);
    PS_REQUIRE(P >= 0, "hash_bwd: negative P");
    if (P == 0) return 0;
    PS_REQUIRE(x01 && dout && (dtable || ms.dtables), "hash_bwd: null pointer");
    PS_REQUIRE(dx == nullptr || table != nullptr, "hash_bwd: dx requested but table is null");
    const int threads = 256;
    const int64_t blocks = cdiv(P, threads);
    PS_REQUIRE(blocks < (1ll << 31), "hash_bwd: too many points");
    cudaStream_t s = (cudaStream_t)stream;
    const int lpt = pick_lpt(L, F, log2_T);
    const dim3 grid((unsigned)blocks, (unsigned)(L / lpt));
#define PS_BWD(FF, LL)                                                                                  \
    if (dx)                                                                                             \
        hash_bwd_kernel<FF, LL, true><<<grid, threads, 0, s>>>(x01, P, table, hp, dout, dtable, dx, level_major, ms); \
    else                                                                                                \
        hash_bwd_kernel<FF, LL, false><<<grid, threads, 0, s>>>(x01, P, table, hp, dout, dtable, dx, level_major, ms);
    PS_DISPATCH_F_LPT(F, lpt, PS_BWD)
#undef PS_BWD
    return check_launch("hash_bwd");
}

extern "C" int ps_hash_fwd(const float* x01, int64_t P, const float* table, const float* scalings_host, int L, int F,
                           int log2_T, float* out, void* stream) {
    return hash_fwd_impl(x01, P, table, scalings_host, L, F, log2_T, out, 0, stream);
}
extern "C" int ps_hash_bwd(const float* x01, int64_t P, const float* table, const float* scalings_host, int L, int F,
                           int log2_T, const float* dout, float* dtable, float* dx, void* stream) {
    return hash_bwd_impl(x01, P, table, scalings_host, L, F, log2_T, dout, dtable, dx, 0, stream);
}
// level-major feature layout [L][P][F] (fused path: coalesced per-level stores / loads)
extern "C" int ps_hash_fwd_lm(const float* x01, int64_t P, const float* table, const float* scalings_host, int L,
                              int F, int log2_T, float* out, void* stream) {
    return hash_fwd_impl(x01, P, table, scalings_host, L, F, log2_T, out, 1, stream);
}
extern "C" int ps_hash_bwd_lm(const float* x01, int64_t P, const float* table, const float* scalings_host, int L,
                              int F, int log2_T, const float* dout, float* dtable, float* dx, void* stream) {
    return hash_bwd_impl(x01, P, table, scalings_host, L, F, log2_T, dout, dtable, dx, 1, stream);
}

// sub-field mode (level-major features [L][rows][F] over the rows of ms_route.cu's segments; rows a multiple of 128):
// tables / dtables = device arrays of one pointer per sub-field
extern "C" int ps_hash_fwd_ms(const float* x01_sorted, int64_t rows, const float* const* tables, const uint8_t* tile_sf,
                              const float* scalings_host, int L, int F, int log2_T, float* out, void* stream) {
    PS_REQUIRE(tables && tile_sf && rows % 128 == 0, "hash_fwd_ms: null pointer or rows not a multiple of 128");
    MsTables ms{tables, nullptr, tile_sf, nullptr};
    return hash_fwd_impl(x01_sorted, rows, nullptr, scalings_host, L, F, log2_T, out, 1, stream, ms);
}
extern "C" int ps_hash_bwd_ms(const float* x01_sorted, int64_t rows, float* const* dtables, const uint8_t* tile_sf,
                              const int32_t* perm, const float* scalings_host, int L, int F, int log2_T, const float* dout,
                              void* stream) {
    PS_REQUIRE(dtables && tile_sf && perm && rows % 128 == 0, "hash_bwd_ms: null pointer or rows not a multiple of 128");
    MsTables ms{nullptr, dtables, tile_sf, perm};
    return hash_bwd_impl(x01_sorted, rows, nullptr, scalings_host, L, F, log2_T, dout, nullptr, nullptr, 1, stream, ms);
}

// how many consecutive levels one thread handles for this grid shape (callers use it to choose the feature layout:
// level-major pays off when a thread's LPT*F floats do not fill a 32-byte sector)
extern "C" int ps_hash_levels_per_thread(int L, int F, int log2_T) { return pick_lpt(L, F, log2_T); }

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
