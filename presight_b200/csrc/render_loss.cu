// Per-ray tail of the train step: the model epilogue (sky blending of rgb / semantics) and the rendered-output loss terms,
// each as ONE kernel per direction instead of ~90 elementwise / reduction launches between compositing-forward and
// compositing-backward.  One warp per ray everywhere.
//
// References:
//   epilogue  models/PreSight/nerfacto_nusc_ms.py:512-532
//               accumulation = clamp(acc_raw, 0, 1); rgb = rgb + (1 - accumulation) * sky_rgb;
//               semantics = semantics + (1 - accumulation) * sky_semantics          (eval: rgb clamped to [0,1] first)
//   rgb loss  models/PreSight/nerfacto_nusc_ms.py:560-567  (MSELoss, mean over N*3)
//   sky loss  model_components/PreSight/losses.py:106-115  (BCE on clip(acc, eps, 1-eps) against 1 - sky_mask, mean;
//             torch's BCE clamps both logs at -100)
//   semantic  model_components/PreSight/losses.py:117-125  (MSE against clip(target, 0, 1), mean over N*C)
#include "common.cuh"

namespace ps {

constexpr int kTailWarps = 8;

__global__ void __launch_bounds__(kTailWarps * 32) sky_blend_fwd_kernel(
    const float* __restrict__ rgb_f, const float* __restrict__ acc_raw, const float* __restrict__ sem_f,
    const float* __restrict__ sky_rgb, const float* __restrict__ sky_sem, int64_t N, int C, int clamp_rgb,
    float* __restrict__ rgb, float* __restrict__ acc, float* __restrict__ sem) {
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kTailWarps + (threadIdx.x >> 5);
    if (n >= N) return;
    const float a = fminf(fmaxf(__ldg(acc_raw + n), 0.f), 1.f);
    const float bg = 1.f - a;
    if (lane == 3) acc[n] = a;
    if (lane < 3) {
        float v = __ldg(rgb_f + n * 3 + lane);
        if (clamp_rgb) v = fminf(fmaxf(v, 0.f), 1.f);
        if (sky_rgb) v = __fadd_rn(v, __fmul_rn(bg, __ldg(sky_rgb + n * 3 + lane)));   // un-fused, as torch evaluates it
        rgb[n * 3 + lane] = v;
    }
    if (sem) {
        for (int c = lane; c < C; c += 32) {
            float v = __ldg(sem_f + n * C + c);
            if (sky_sem) v = __fadd_rn(v, __fmul_rn(bg, __ldg(sky_sem + n * C + c)));
            sem[n * C + c] = v;
        }
    }
}

// d_rgb_f / d_sem_f equal the incoming gradients (the host wrapper passes them through) unless rgb was clamped.
__global__ void __launch_bounds__(kTailWarps * 32) sky_blend_bwd_kernel(
    const float* __restrict__ rgb_f, const float* __restrict__ acc_raw, const float* __restrict__ sky_rgb,
    const float* __restrict__ sky_sem, const float* __restrict__ d_rgb, const float* __restrict__ d_acc,
    const float* __restrict__ d_sem, int64_t N, int C, int clamp_rgb, float* __restrict__ d_rgb_f,
    float* __restrict__ d_acc_raw, float* __restrict__ d_sky_rgb, float* __restrict__ d_sky_sem) {
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kTailWarps + (threadIdx.x >> 5);
    if (n >= N) return;
    const float ar = __ldg(acc_raw + n);
    const float a = fminf(fmaxf(ar, 0.f), 1.f);
    const float bg = 1.f - a;
    float dot = 0.f;   // sum of (upstream gradient x sky value): -d/d accumulation of the blended outputs
    if (lane < 3 && d_rgb) {
        const float g = __ldg(d_rgb + n * 3 + lane);
        if (sky_rgb) {
            dot += g * __ldg(sky_rgb + n * 3 + lane);
            if (d_sky_rgb) d_sky_rgb[n * 3 + lane] = bg * g;
        }
        if (d_rgb_f) {
            const float v = __ldg(rgb_f + n * 3 + lane);
            d_rgb_f[n * 3 + lane] = (!clamp_rgb || (v >= 0.f && v <= 1.f)) ? g : 0.f;
        }
    } else if (lane < 3) {
        if (d_sky_rgb) d_sky_rgb[n * 3 + lane] = 0.f;
        if (d_rgb_f) d_rgb_f[n * 3 + lane] = 0.f;
    }
    if (sky_sem) {
        for (int c = lane; c < C; c += 32) {
            const float g = d_sem ? __ldg(d_sem + n * C + c) : 0.f;
            dot += g * __ldg(sky_sem + n * C + c);
            if (d_sky_sem) d_sky_sem[n * C + c] = bg * g;
        }
    }
    dot = warp_sum(dot);
    if (lane == 0) {
        const float g = (d_acc ? __ldg(d_acc + n) : 0.f) - dot;
        d_acc_raw[n] = (ar >= 0.f && ar <= 1.f) ? g : 0.f;     // torch.clamp passes the gradient on [min, max]
    }
}

// losses[0..2] += mean terms (rgb MSE, sky BCE, semantic MSE); g_* = d losses[i] / d input (already divided by the
// element counts).  Terms whose inputs are NULL are skipped.
__global__ void __launch_bounds__(kTailWarps * 32) render_losses_kernel(
    const float* __restrict__ rgb, const float* __restrict__ gt_rgb, const float* __restrict__ acc,
    const float* __restrict__ sky_mask, const float* __restrict__ sem, const float* __restrict__ gt_sem, int64_t N,
    int C, float eps, float* __restrict__ losses, float* __restrict__ g_rgb, float* __restrict__ g_acc,
    float* __restrict__ g_sem) {
    __shared__ float part[kTailWarps][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n = (int64_t)blockIdx.x * kTailWarps + warp;
    float l_rgb = 0.f, l_sky = 0.f, l_sem = 0.f;
    if (n < N) {
        if (rgb && lane < 3) {
            const float d = __ldg(rgb + n * 3 + lane) - __ldg(gt_rgb + n * 3 + lane);
            l_rgb = d * d;
            if (g_rgb) g_rgb[n * 3 + lane] = 2.f * d / (float)(3 * N);
        }
        if (acc && lane == 3) {
            const float a = __ldg(acc + n);
            const float x = fminf(fmaxf(a, eps), 1.f - eps);
            const float t = 1.f - __ldg(sky_mask + n);
            l_sky = -(t * fmaxf(logf(x), -100.f) + (1.f - t) * fmaxf(logf(1.f - x), -100.f));
            if (g_acc) {
                const float g = (x - t) / fmaxf((1.f - x) * x, 1e-12f);
                g_acc[n] = (a >= eps && a <= 1.f - eps) ? g / (float)N : 0.f;
            }
        }
        if (sem) {
            const float inv = 1.f / ((float)N * (float)C);
            for (int c = lane; c < C; c += 32) {
                const float t = fminf(fmaxf(__ldg(gt_sem + n * C + c), 0.f), 1.f);
                const float d = __ldg(sem + n * C + c) - t;
                l_sem += d * d;
                if (g_sem) g_sem[n * C + c] = 2.f * d * inv;
            }
        }
    }
    l_rgb = warp_sum(l_rgb);
    l_sky = warp_sum(l_sky);
    l_sem = warp_sum(l_sem);
    if (lane == 0) {
        part[warp][0] = l_rgb;
        part[warp][1] = l_sky;
        part[warp][2] = l_sem;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kTailWarps; ++w) s += part[w][threadIdx.x];
        const float scale = threadIdx.x == 0 ? 1.f / (float)(3 * N) : (threadIdx.x == 1 ? 1.f / (float)N : 1.f / ((float)N * (float)C));
        const bool on = threadIdx.x == 0 ? rgb != nullptr : (threadIdx.x == 1 ? acc != nullptr : sem != nullptr);
        if (on) atomicAdd(losses + threadIdx.x, s * scale);
    }
}

}  // namespace ps

using namespace ps;

extern "C" int ps_sky_blend_fwd(const float* rgb_f, const float* acc_raw, const float* sem_f, const float* sky_rgb,
                                const float* sky_sem, int64_t N, int C, int clamp_rgb, float* rgb, float* acc, float* sem,
                                void* stream) {
    if (N == 0) return 0;
    PS_REQUIRE(rgb_f && acc_raw && rgb && acc, "sky_blend_fwd: null pointer");
    PS_REQUIRE((sem_f == nullptr) == (sem == nullptr), "sky_blend_fwd: sem_f and sem must both be given or both be null");
    PS_REQUIRE(sky_sem == nullptr || sem_f != nullptr, "sky_blend_fwd: sky_sem without sem_f");
    PS_REQUIRE(C >= 0 && (sem_f == nullptr || C >= 1), "sky_blend_fwd: bad channel count %d", C);
    sky_blend_fwd_kernel<<<(unsigned)cdiv(N, kTailWarps), kTailWarps * 32, 0, (cudaStream_t)stream>>>(
        rgb_f, acc_raw, sem_f, sky_rgb, sky_sem, N, C, clamp_rgb, rgb, acc, sem);
    return check_launch("sky_blend_fwd");
}

extern "C" int ps_sky_blend_bwd(const float* rgb_f, const float* acc_raw, const float* sky_rgb, const float* sky_sem,
                                const float* d_rgb, const float* d_acc, const float* d_sem, int64_t N, int C,
                                int clamp_rgb, float* d_rgb_f, float* d_acc_raw, float* d_sky_rgb, float* d_sky_sem,
                                void* stream) {
    if (N == 0) return 0;
    PS_REQUIRE(acc_raw && d_acc_raw, "sky_blend_bwd: null pointer");
    PS_REQUIRE(d_rgb_f == nullptr || rgb_f != nullptr, "sky_blend_bwd: d_rgb_f needs rgb_f");
    PS_REQUIRE(C >= 0, "sky_blend_bwd: bad channel count %d", C);
    sky_blend_bwd_kernel<<<(unsigned)cdiv(N, kTailWarps), kTailWarps * 32, 0, (cudaStream_t)stream>>>(
        rgb_f, acc_raw, sky_rgb, sky_sem, d_rgb, d_acc, d_sem, N, C, clamp_rgb, d_rgb_f, d_acc_raw, d_sky_rgb, d_sky_sem);
    return check_launch("sky_blend_bwd");
}

extern "C" int ps_render_losses(const float* rgb, const float* gt_rgb, const float* acc, const float* sky_mask,
                                const float* sem, const float* gt_sem, int64_t N, int C, float eps, float* losses,
                                float* g_rgb, float* g_acc, float* g_sem, void* stream) {
    if (N == 0) return 0;
    PS_REQUIRE(losses != nullptr, "render_losses: losses is null");
    PS_REQUIRE((rgb == nullptr) == (gt_rgb == nullptr), "render_losses: rgb and gt_rgb go together");
    PS_REQUIRE((acc == nullptr) == (sky_mask == nullptr), "render_losses: acc and sky_mask go together");
    PS_REQUIRE((sem == nullptr) == (gt_sem == nullptr), "render_losses: sem and gt_sem go together");
    PS_REQUIRE(sem == nullptr || C >= 1, "render_losses: bad channel count %d", C);
    render_losses_kernel<<<(unsigned)cdiv(N, kTailWarps), kTailWarps * 32, 0, (cudaStream_t)stream>>>(
        rgb, gt_rgb, acc, sky_mask, sem, gt_sem, N, C, eps, losses, g_rgb, g_acc, g_sem);
    return check_launch("render_losses");
}
