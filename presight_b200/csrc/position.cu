// Position prologue / small per-point kernels: aabb normalisation + L-inf contraction + selector,
// frustum mid-points, degree-4 SH, nearest-centroid routing, trunc_exp.
#include <cuda_fp16.h>

#include "position.cuh"

namespace ps {

__global__ void __launch_bounds__(256) normalize_kernel(const float* __restrict__ pos, int64_t P, Aabb box,
                                                        int contract, float* __restrict__ x01,
                                                        uint8_t* __restrict__ sel) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float v[3] = {pos[3 * p], pos[3 * p + 1], pos[3 * p + 2]};
    const bool inside = normalize_point(v, box, contract != 0);
    x01[3 * p] = v[0];
    x01[3 * p + 1] = v[1];
    x01[3 * p + 2] = v[2];
    if (sel) sel[p] = inside ? 1 : 0;
}

__global__ void __launch_bounds__(256) sample_positions_kernel(const float* __restrict__ o,
                                                               const float* __restrict__ d,
                                                               const float* __restrict__ bins, int64_t N, int S,
                                                               float* __restrict__ pos) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * S) return;
    const int64_t n = i / S;
    const int s = (int)(i - n * S);
    const float a = bins[n * (S + 1) + s], b = bins[n * (S + 1) + s + 1];
    float out[3];
    frustum_midpoint(o + 3 * n, d + 3 * n, a, b, out);
    pos[3 * i] = out[0];
    pos[3 * i + 1] = out[1];
    pos[3 * i + 2] = out[2];
}

// frustum mid-point + aabb normalisation + contraction + selector in one pass (no world-space positions in HBM)
__global__ void __launch_bounds__(256) ray_points_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                                         const float* __restrict__ bins, int64_t N, int S, Aabb box,
                                                         int contract, float* __restrict__ x01,
                                                         uint8_t* __restrict__ sel) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * S) return;
    const int64_t n = i / S;
    const int s = (int)(i - n * S);
    float v[3];
    frustum_midpoint(o + 3 * n, d + 3 * n, __ldg(bins + n * (S + 1) + s), __ldg(bins + n * (S + 1) + s + 1), v);
    const bool inside = normalize_point(v, box, contract != 0);
    x01[3 * i] = v[0];
    x01[3 * i + 1] = v[1];
    x01[3 * i + 2] = v[2];
    sel[i] = inside ? 1 : 0;
}

__global__ void __launch_bounds__(256) sh4_kernel(const float* __restrict__ dirs, int64_t P, int mapped,
                                                  float* __restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float c[16];
    if (mapped)
        sh4_of_mapped(dirs[3 * p], dirs[3 * p + 1], dirs[3 * p + 2], c);
    else
        sh4_of_direction(dirs[3 * p], dirs[3 * p + 1], dirs[3 * p + 2], c);
    float4* dst = reinterpret_cast<float4*>(out + 16 * p);
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = make_float4(c[4 * q], c[4 * q + 1], c[4 * q + 2], c[4 * q + 3]);
}

struct Centroids {
    float c[PS_MAX_FIELDS][3];
    int n;
};

__global__ void __launch_bounds__(256) nearest_centroid_kernel(const float* __restrict__ pos, int64_t P,
                                                               const float* __restrict__ cent, int nf,
                                                               int32_t* __restrict__ assign) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float x = pos[3 * p], y = pos[3 * p + 1], z = pos[3 * p + 2];
    float best = INFINITY;
    int arg = 0;
    for (int j = 0; j < nf; ++j) {
        const float dx = x - __ldg(cent + 3 * j), dy = y - __ldg(cent + 3 * j + 1), dz = z - __ldg(cent + 3 * j + 2);
        // torch.cdist (p=2) returns sqrt of the squared distance; argmin keeps the first minimum
        const float d2 = sqrtf(dx * dx + dy * dy + dz * dz);
        if (d2 < best) {
            best = d2;
            arg = j;
        }
    }
    assign[p] = arg;
}

__global__ void __launch_bounds__(256) trunc_exp_fwd_kernel(const float* __restrict__ x,
                                                            const uint8_t* __restrict__ sel, int64_t P,
                                                            int64_t xs, float* __restrict__ y) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float e = expf(x[p * xs]);
    y[p] = sel ? e * (float)sel[p] : e;
}

__global__ void __launch_bounds__(256) trunc_exp_bwd_kernel(const float* __restrict__ x,
                                                            const uint8_t* __restrict__ sel,
                                                            const float* __restrict__ dy, int64_t P, int64_t xs,
                                                            float* __restrict__ dx, int64_t dxs) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float g = dy[p];
    if (sel) g *= (float)sel[p];
    dx[p * dxs] = g * expf(fminf(fmaxf(x[p * xs], -15.f), 15.f));
}

struct DensityPtrs {
    const float* p[8];
    int k;
};

// scripts/extract_priors.py:137-138: mean over the k density estimates, features clipped to [0,1] and cast to fp16
__global__ void __launch_bounds__(256) prior_finalize_kernel(DensityPtrs d, const float* __restrict__ sem, int64_t M,
                                                             int C, float* __restrict__ mean,
                                                             __half* __restrict__ feats) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M && mean) {
        float s = 0.f;
        for (int j = 0; j < d.k; ++j) s += __ldg(d.p[j] + i);
        mean[i] = s / (float)d.k;
    }
    if (feats) {
        // C is a multiple of 2: one half2 per thread-iteration, grid-stride over M*C/2 pairs
        const int64_t pairs = M * C / 2;
        for (int64_t q = i; q < pairs; q += (int64_t)gridDim.x * blockDim.x) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(sem) + q);
            reinterpret_cast<__half2*>(feats)[q] =
                __floats2half2_rn(fminf(fmaxf(v.x, 0.f), 1.f), fminf(fmaxf(v.y, 0.f), 1.f));
        }
    }
}

}  // namespace ps

using namespace ps;

extern "C" int ps_prior_finalize(const float* const* densities_host, int k, const float* sem, int64_t M, int C,
                                 float* mean, void* feats_half, void* stream) {
    if (M == 0) return 0;
    PS_REQUIRE(k >= 1 && k <= 8 && densities_host, "prior_finalize: need 1..8 density arrays");
    PS_REQUIRE(feats_half == nullptr || (sem != nullptr && C % 2 == 0), "prior_finalize: C must be even");
    DensityPtrs d;
    d.k = k;
    for (int j = 0; j < k; ++j) {
        PS_REQUIRE(densities_host[j] != nullptr, "prior_finalize: density %d is null", j);
        d.p[j] = densities_host[j];
    }
    prior_finalize_kernel<<<(unsigned)cdiv(M, 256), 256, 0, (cudaStream_t)stream>>>(d, sem, M, C, mean,
                                                                                    (__half*)feats_half);
    return check_launch("prior_finalize");
}

extern "C" int ps_normalize_positions(const float* pos, int64_t P, const float* aabb_host, int contract, float* x01,
                                      uint8_t* selector, void* stream) {
    if (P == 0) return 0;
    PS_REQUIRE(pos && aabb_host && x01, "normalize_positions: null pointer");
    Aabb box;
    for (int k = 0; k < 3; ++k) {
        box.lo[k] = aabb_host[k];
        box.hi[k] = aabb_host[3 + k];
    }
    normalize_kernel<<<(unsigned)cdiv(P, 256), 256, 0, (cudaStream_t)stream>>>(pos, P, box, contract, x01, selector);
    return check_launch("normalize_positions");
}

extern "C" int ps_sample_positions(const float* origins, const float* dirs, const float* eu_bins, int64_t N, int S,
                                   float* pos, void* stream) {
    if (N == 0 || S == 0) return 0;
    PS_REQUIRE(origins && dirs && eu_bins && pos, "sample_positions: null pointer");
    sample_positions_kernel<<<(unsigned)cdiv(N * S, 256), 256, 0, (cudaStream_t)stream>>>(origins, dirs, eu_bins, N, S,
                                                                                          pos);
    return check_launch("sample_positions");
}

extern "C" int ps_ray_points(const float* origins, const float* dirs, const float* eu_bins, int64_t N, int S,
                             const float* aabb_host, int contract, float* x01, uint8_t* selector, void* stream) {
    if (N == 0 || S == 0) return 0;
    PS_REQUIRE(origins && dirs && eu_bins && aabb_host && x01 && selector, "ray_points: null pointer");
    Aabb box;
    for (int k = 0; k < 3; ++k) {
        box.lo[k] = aabb_host[k];
        box.hi[k] = aabb_host[3 + k];
    }
    ray_points_kernel<<<(unsigned)cdiv(N * S, 256), 256, 0, (cudaStream_t)stream>>>(origins, dirs, eu_bins, N, S, box,
                                                                                    contract, x01, selector);
    return check_launch("ray_points");
}

extern "C" int ps_sh4(const float* dirs, int64_t P, int mapped, float* out, void* stream) {
    if (P == 0) return 0;
    PS_REQUIRE(dirs && out, "sh4: null pointer");
    sh4_kernel<<<(unsigned)cdiv(P, 256), 256, 0, (cudaStream_t)stream>>>(dirs, P, mapped, out);
    return check_launch("sh4");
}

extern "C" int ps_nearest_centroid(const float* pos, int64_t P, const float* centroids, int nf, int32_t* assign,
                                   void* stream) {
    if (P == 0) return 0;
    PS_REQUIRE(pos && centroids && assign, "nearest_centroid: null pointer");
    PS_REQUIRE(nf >= 1, "nearest_centroid: need at least one centroid");
    nearest_centroid_kernel<<<(unsigned)cdiv(P, 256), 256, 0, (cudaStream_t)stream>>>(pos, P, centroids, nf, assign);
    return check_launch("nearest_centroid");
}

extern "C" int ps_trunc_exp_fwd(const float* x, const uint8_t* sel, int64_t P, int64_t x_stride, float* y,
                                void* stream) {
    if (P == 0) return 0;
    PS_REQUIRE(x && y, "trunc_exp_fwd: null pointer");
    trunc_exp_fwd_kernel<<<(unsigned)cdiv(P, 256), 256, 0, (cudaStream_t)stream>>>(x, sel, P, x_stride, y);
    return check_launch("trunc_exp_fwd");
}

extern "C" int ps_trunc_exp_bwd(const float* x, const uint8_t* sel, const float* dy, int64_t P, int64_t x_stride,
                                float* dx, int64_t dx_stride, void* stream) {
    if (P == 0) return 0;
    PS_REQUIRE(x && dy && dx, "trunc_exp_bwd: null pointer");
    trunc_exp_bwd_kernel<<<(unsigned)cdiv(P, 256), 256, 0, (cudaStream_t)stream>>>(x, sel, dy, P, x_stride, dx,
                                                                                   dx_stride);
    return check_launch("trunc_exp_bwd");
}
