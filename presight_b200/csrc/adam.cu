// Fused Adam step over a dense fp32 parameter: one pass over p / grad / m / v instead of the ~10 elementwise
// passes of the unfused optimiser — for the 512 MiB main hash table that is 7 x 512 MiB of HBM traffic per step.
// Reference: torch.optim.Adam as configured by configs/method_configs.py:115 (lr 1e-2, eps 1e-15, weight_decay 1e-5).
#include "common.cuh"

#define PS_HD __device__
#define PS_SQRTF(a) __fsqrt_rn(a)
#include "adam_core.h"

namespace ps {

template <bool VEC>
__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                        adam::Scalars s) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = VEC ? (n >> 2) : 0;          // VEC = false: buffers not 16-byte aligned (views at odd offsets)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i],
               vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
        adam::update(pp.x, gg.x, mm.x, vv.x, s);
        adam::update(pp.y, gg.y, mm.y, vv.y, s);
        adam::update(pp.z, gg.z, mm.z, vv.z, s);
        adam::update(pp.w, gg.w, mm.w, vv.w, s);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        adam::update(p[i], g[i], m[i], v[i], s);
}

}  // namespace ps

using namespace ps;

extern "C" int ps_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                            double beta1, double beta2, double eps, double weight_decay, int64_t step, void* stream) {
    if (n == 0) return 0;
    PS_REQUIRE(param && grad && exp_avg && exp_avg_sq, "adam_step: null pointer");
    PS_REQUIRE(step >= 1, "adam_step: step must be >= 1 (it counts the update being made)");
    const bool vec = ((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0;
    // the scalars in double precision, as torch computes them (optim/adam.py, _single_tensor_adam)
    // (hyper-parameters arrive as doubles — torch holds them as Python floats; 1 - float(0.999) is off by 1.3e-5)
    const double bias1 = 1.0 - pow(beta1, (double)step), bias2 = 1.0 - pow(beta2, (double)step);
    adam::Scalars s;
    s.weight_decay = (float)weight_decay;
    s.one_minus_beta1 = (float)(1.0 - beta1);
    s.beta2 = (float)beta2;
    s.one_minus_beta2 = (float)(1.0 - beta2);
    s.step_size = (float)(lr / bias1);
    s.bias2_sqrt = (float)sqrt(bias2);
    s.eps = (float)eps;
    const int64_t work = (n + 3) / 4;
    int64_t blocks = cdiv(work, 256);
    const int64_t cap = (int64_t)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    if (vec) adam_step_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, s);
    else adam_step_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, s);
    return check_launch("adam_step");
}
