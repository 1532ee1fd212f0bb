// Pinhole ray generation on the device (the caller on the input side of the hot path): one thread per ray
// runs raygen_core.h.  Replaces RayGenerator.forward + Cameras.generate_rays (model_components/ray_generators.py:43-61,
// cameras/cameras.py:497-880) for PERSPECTIVE cameras without distortion parameters — PreSight's nuScenes cameras.
#include "common.cuh"

#define PS_HD __device__
#define PS_MUL(a, b) __fmul_rn((a), (b))
#define PS_ADD(a, b) __fadd_rn((a), (b))
#define PS_SUB(a, b) __fsub_rn((a), (b))
#define PS_DIV(a, b) __fdiv_rn((a), (b))
#define PS_SQRT(a) __fsqrt_rn(a)
#include "raygen_core.h"

namespace ps {

__global__ void __launch_bounds__(256) generate_rays_kernel(const float* __restrict__ c2w, const float* __restrict__ fx,
                                                            const float* __restrict__ fy, const float* __restrict__ cx,
                                                            const float* __restrict__ cy,
                                                            const int64_t* __restrict__ ray_indices, int64_t N, int C,
                                                            float pixel_offset, float* __restrict__ origins,
                                                            float* __restrict__ directions, float* __restrict__ pixel_area,
                                                            float* __restrict__ directions_norm) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    int64_t cam = ray_indices[3 * n];
    // the caller checks the camera indices (RayGenerator.forward: a device-side assert, where the reference's gather
    // raises); the clamp only keeps a bad index from reading out of range before that assert is observed
    cam = cam < 0 ? 0 : (cam >= C ? C - 1 : cam);
    float m[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) m[i] = __ldg(c2w + cam * 12 + i);
    float o[3], d[3], area, norm;
    raygen::pinhole_ray(m, __ldg(fx + cam), __ldg(fy + cam), __ldg(cx + cam), __ldg(cy + cam), ray_indices[3 * n + 1],
                        ray_indices[3 * n + 2], pixel_offset, o, d, &area, &norm);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        origins[3 * n + i] = o[i];
        directions[3 * n + i] = d[i];
    }
    pixel_area[n] = area;
    if (directions_norm) directions_norm[n] = norm;
}

}  // namespace ps

using namespace ps;

extern "C" int ps_generate_rays(const float* c2w, const float* fx, const float* fy, const float* cx, const float* cy,
                                int C, const int64_t* ray_indices, int64_t N, float pixel_offset, float* origins,
                                float* directions, float* pixel_area, float* directions_norm, void* stream) {
    if (N == 0) return 0;
    PS_REQUIRE(C >= 1, "generate_rays: no cameras");
    PS_REQUIRE(c2w && fx && fy && cx && cy && ray_indices && origins && directions && pixel_area,
               "generate_rays: null pointer");
    generate_rays_kernel<<<(unsigned)cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(
        c2w, fx, fy, cx, cy, ray_indices, N, C, pixel_offset, origins, directions, pixel_area, directions_norm);
    return check_launch("generate_rays");
}
