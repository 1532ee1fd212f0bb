// Kernel #4: volumetric compositing, one warp per ray, samples striped across lanes in
// chunks of 32 with a running carry (scan by warp shuffles, fp64 accumulation like torch's CPU cumsum).
#include "composite.cuh"

namespace ps {

constexpr int kRayWarps = 8;  // warps (rays) per CTA

// ---- RaySamples.get_weights (cameras/rays.py:138-150) -----------------------------------
__global__ void __launch_bounds__(kRayWarps * 32) weights_fwd_kernel(const float* __restrict__ deltas,
                                                                     const float* __restrict__ density, int64_t N,
                                                                     int S, float* __restrict__ weights) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kRayWarps + warp;
    if (n >= N) return;
    double carry = 0.0;
    for (int base = 0; base < S; base += 32) {
        const int j = base + lane;
        const float dd = j < S ? __fmul_rn(__ldg(deltas + n * S + j), __ldg(density + n * S + j)) : 0.f;
        float w, T;
        weight_step(dd, lane, carry, w, T);
        if (j < S) weights[n * S + j] = w;
    }
}

__global__ void __launch_bounds__(kRayWarps * 32) weights_bwd_kernel(const float* __restrict__ deltas,
                                                                     const float* __restrict__ density,
                                                                     const float* __restrict__ dweights, int64_t N,
                                                                     int S, float* __restrict__ ddensity) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kRayWarps + warp;
    if (n >= N) return;
    // pass 1: G = sum_k g_k w_k
    double carry = 0.0, G = 0.0;
    for (int base = 0; base < S; base += 32) {
        const int j = base + lane;
        const float dd = j < S ? __fmul_rn(__ldg(deltas + n * S + j), __ldg(density + n * S + j)) : 0.f;
        float w, T;
        const bool finite = weight_step(dd, lane, carry, w, T);
        if (j < S && finite) G += (double)__ldg(dweights + n * S + j) * (double)w;
    }
    G = warp_sum(G);
    // pass 2: d dd_i = g_i T_{i+1} - (G - P_i),  P_i inclusive prefix of g_k w_k
    carry = 0.0;
    double pcarry = 0.0;
    for (int base = 0; base < S; base += 32) {
        const int j = base + lane;
        const float dl = j < S ? __ldg(deltas + n * S + j) : 0.f;
        const float dd = j < S ? __fmul_rn(dl, __ldg(density + n * S + j)) : 0.f;
        float w, T;
        const bool finite = weight_step(dd, lane, carry, w, T);
        const float g = (j < S && finite) ? __ldg(dweights + n * S + j) : 0.f;
        const double P = warp_scan_incl((double)g * (double)w, lane) + pcarry;
        pcarry = __shfl_sync(0xffffffffu, P, 31);
        if (j < S) {
            const float Tnext = T * expf(-dd);
            ddensity[n * S + j] = dl * (float)((double)g * (double)Tnext - (G - P));
        }
    }
}

// ---- weighted sums along the ray --------------------------------------------------------
// lanes over samples (small C)
template <int C>
__global__ void __launch_bounds__(kRayWarps * 32) render_fwd_small_kernel(const float* __restrict__ weights,
                                                                          const float* __restrict__ values,
                                                                          int64_t N, int S, float* __restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kRayWarps + warp;
    if (n >= N) return;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    for (int j = lane; j < S; j += 32) {
        const float w = __ldg(weights + n * S + j);
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] += values ? w * __ldg(values + (n * S + j) * C + c) : w;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float t = warp_sum(acc[c]);
        if (lane == 0) out[n * C + c] = t;
    }
}

// lanes over channels (C >= 32 or generic)
__global__ void __launch_bounds__(kRayWarps * 32) render_fwd_wide_kernel(const float* __restrict__ weights,
                                                                         const float* __restrict__ values, int64_t N,
                                                                         int S, int C, float* __restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kRayWarps + warp;
    if (n >= N) return;
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        float acc = 0.f;
        if (c < C)
            for (int j = 0; j < S; ++j) acc += __ldg(weights + n * S + j) * __ldg(values + (n * S + j) * C + c);
        if (c < C) out[n * C + c] = acc;
    }
}

__global__ void __launch_bounds__(256) render_bwd_kernel(const float* __restrict__ weights,
                                                         const float* __restrict__ values,
                                                         const float* __restrict__ dout, int64_t N, int S, int C,
                                                         float* __restrict__ dweights, float* __restrict__ dvalues) {
    // one thread per (ray, sample): dw += <v, dout>, dv = w * dout
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * S) return;
    const int64_t n = i / S;
    const float w = __ldg(weights + i);
    float dot = 0.f;
    for (int c = 0; c < C; ++c) {
        const float g = __ldg(dout + n * C + c);
        if (values) {
            dot += g * __ldg(values + i * C + c);
            if (dvalues) dvalues[i * C + c] = w * g;
        } else {
            dot += g;
        }
    }
    dweights[i] += dot;
}

// ---- DepthRenderer "threshold" (renderers.py:352-362) -------------------------------------
__global__ void __launch_bounds__(kRayWarps * 32) depth_threshold_kernel(const float* __restrict__ weights,
                                                                         const float* __restrict__ bins, int64_t N,
                                                                         int S, float thr, float* __restrict__ depth,
                                                                         int64_t* __restrict__ index) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kRayWarps + warp;
    if (n >= N) return;
    double carry = 0.0;
    int found = S;
    for (int base = 0; base < S && found == S; base += 32) {
        const int j = base + lane;
        const float w = j < S ? __ldg(weights + n * S + j) : 0.f;
        const double incl = warp_scan_incl((double)w, lane) + carry;
        carry = __shfl_sync(0xffffffffu, incl, 31);
        const unsigned hit = __ballot_sync(0xffffffffu, j < S && (float)incl >= thr);  // searchsorted(side="left")
        if (hit) found = base + __ffs(hit) - 1;
    }
    found = min(found, S - 1);
    if (lane == 0) {
        const float a = __ldg(bins + n * (S + 1) + found), b = __ldg(bins + n * (S + 1) + found + 1);
        depth[n] = __fdiv_rn(__fadd_rn(a, b), 2.f);
        if (index) index[n] = found;
    }
}

// ---- one-pass compositing for the model fast path ------------------------------------------
__global__ void __launch_bounds__(kRayWarps * 32) composite_fwd_kernel(
    const float* __restrict__ bins, const float* __restrict__ density, const float* __restrict__ rgb,
    const float* __restrict__ sem, int64_t N, int S, int C, float thr, float* __restrict__ weights,
    float* __restrict__ rgb_out, float* __restrict__ acc_out, float* __restrict__ dexp_out,
    float* __restrict__ dthr_out, float* __restrict__ sem_out, float* __restrict__ tminmax) {
    extern __shared__ float smem[];  // per warp: w[S]
    __shared__ float s_min[kRayWarps], s_max[kRayWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kRayWarps + warp;
    float tmin = INFINITY, tmax = -INFINITY;
    if (n < N) {
        float* wbuf = smem + (size_t)warp * S;
        const float* b = bins + n * (S + 1);
        double carry = 0.0, wcarry = 0.0;
        float r = 0.f, g = 0.f, bl = 0.f, acc = 0.f, dnum = 0.f;
        int found = S;
        for (int base = 0; base < S; base += 32) {
            const int j = base + lane;
            float t0 = 0.f, t1 = 0.f, sigma = 0.f;
            if (j < S) {
                t0 = __ldg(b + j);
                t1 = __ldg(b + j + 1);
                sigma = __ldg(density + n * S + j);
            }
            const float dd = __fmul_rn(__fsub_rn(t1, t0), sigma);
            float w, T;
            weight_step(dd, lane, carry, w, T);
            if (j >= S) w = 0.f;
            const float tm = __fdiv_rn(__fadd_rn(t0, t1), 2.f);
            if (j < S) {
                wbuf[j] = w;
                if (weights) weights[n * S + j] = w;
                tmin = fminf(tmin, tm);
                tmax = fmaxf(tmax, tm);
                acc += w;
                dnum += w * tm;
                if (rgb) {
                    const float* c = rgb + (n * S + j) * 3;
                    r += w * __ldg(c);
                    g += w * __ldg(c + 1);
                    bl += w * __ldg(c + 2);
                }
            }
            const double incl = warp_scan_incl((double)w, lane) + wcarry;
            wcarry = __shfl_sync(0xffffffffu, incl, 31);
            const unsigned hit = __ballot_sync(0xffffffffu, j < S && (float)incl >= thr);
            if (hit && found == S) found = base + __ffs(hit) - 1;
        }
        r = warp_sum(r); g = warp_sum(g); bl = warp_sum(bl); acc = warp_sum(acc); dnum = warp_sum(dnum);
        found = min(found, S - 1);
        if (lane == 0) {
            if (rgb_out) { rgb_out[3 * n] = r; rgb_out[3 * n + 1] = g; rgb_out[3 * n + 2] = bl; }
            if (acc_out) acc_out[n] = acc;
            if (dexp_out) dexp_out[n] = dnum / (acc + 1e-10f);
            if (dthr_out) dthr_out[n] = __fdiv_rn(__fadd_rn(__ldg(b + found), __ldg(b + found + 1)), 2.f);
        }
        if (sem && sem_out) {
            __syncwarp();
            for (int c0 = 0; c0 < C; c0 += 32) {
                const int c = c0 + lane;
                if (c < C) {
                    float a = 0.f;
                    for (int j = 0; j < S; ++j) a += wbuf[j] * __ldg(sem + (n * S + j) * C + c);
                    sem_out[n * C + c] = a;
                }
            }
        }
    }
    if (tminmax) {
        tmin = warp_min(tmin);
        tmax = warp_max(tmax);
        if (lane == 0) { s_min[warp] = tmin; s_max[warp] = tmax; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < kRayWarps; ++k) { tmin = fminf(tmin, s_min[k]); tmax = fmaxf(tmax, s_max[k]); }
            if (tmin <= tmax) { atomic_min_float(tminmax, tmin); atomic_max_float(tminmax + 1, tmax); }
        }
    }
}

__global__ void __launch_bounds__(kRayWarps * 32) composite_bwd_kernel(
    const float* __restrict__ bins, const float* __restrict__ density, const float* __restrict__ rgb,
    const float* __restrict__ sem, const float* __restrict__ acc_in, const float* __restrict__ dexp_in, int64_t N,
    int S, int C, const float* __restrict__ d_w_in, const float* __restrict__ d_rgb_out,
    const float* __restrict__ d_acc, const float* __restrict__ d_dexp, const float* __restrict__ d_sem_out,
    float* __restrict__ d_density, float* __restrict__ d_rgb, float* __restrict__ d_sem) {
    extern __shared__ float smem[];  // per warp: gw[S] | gsem[C]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kRayWarps + warp;
    if (n >= N) return;
    const int Spad = (S + 3) & ~3;  // keeps gsem 16-byte aligned for the float4 reads
    float* gwbuf = smem + (size_t)warp * (Spad + C);
    float* gsem = gwbuf + Spad;
    const float* b = bins + n * (S + 1);
    const bool has_sem = sem && d_sem_out;
    if (has_sem)
        for (int c = lane; c < C; c += 32) gsem[c] = __ldg(d_sem_out + n * C + c);
    float gr = 0.f, gg = 0.f, gb = 0.f;
    if (rgb && d_rgb_out) { gr = __ldg(d_rgb_out + 3 * n); gg = __ldg(d_rgb_out + 3 * n + 1); gb = __ldg(d_rgb_out + 3 * n + 2); }
    const float gacc = d_acc ? __ldg(d_acc + n) : 0.f;
    float gdep = 0.f, depth = 0.f, inv_den = 0.f;
    if (d_dexp) {
        gdep = __ldg(d_dexp + n);
        depth = __ldg(dexp_in + n);
        inv_den = 1.f / (__ldg(acc_in + n) + 1e-10f);
    }
    __syncwarp();
    // pass 1: total upstream gradient per weight, G = sum g_k w_k, and the value gradients
    double carry = 0.0, G = 0.0;
    for (int base = 0; base < S; base += 32) {
        const int j = base + lane;
        float t0 = 0.f, t1 = 0.f, sigma = 0.f;
        if (j < S) { t0 = __ldg(b + j); t1 = __ldg(b + j + 1); sigma = __ldg(density + n * S + j); }
        const float dd = __fmul_rn(__fsub_rn(t1, t0), sigma);
        float w, T;
        const bool finite = weight_step(dd, lane, carry, w, T);
        if (j < S) {
            const float tm = __fdiv_rn(__fadd_rn(t0, t1), 2.f);
            float g = d_w_in ? __ldg(d_w_in + n * S + j) : 0.f;
            g += gacc + gdep * (tm - depth) * inv_den;
            if (rgb && d_rgb_out) {
                const float* c = rgb + (n * S + j) * 3;
                g += gr * __ldg(c) + gg * __ldg(c + 1) + gb * __ldg(c + 2);
                if (d_rgb) {
                    float* dc = d_rgb + (n * S + j) * 3;
                    dc[0] = w * gr; dc[1] = w * gg; dc[2] = w * gb;
                }
            }
            if (has_sem) {
                const float4* srow = reinterpret_cast<const float4*>(sem + (n * S + j) * C);
                float4* drow = d_sem ? reinterpret_cast<float4*>(d_sem + (n * S + j) * C) : nullptr;
                float dot = 0.f;
                for (int q = 0; q < C / 4; ++q) {
                    const float4 v = __ldg(srow + q);
                    const float4 gs = *reinterpret_cast<const float4*>(gsem + 4 * q);
                    dot += v.x * gs.x + v.y * gs.y + v.z * gs.z + v.w * gs.w;
                    if (drow) drow[q] = make_float4(w * gs.x, w * gs.y, w * gs.z, w * gs.w);
                }
                g += dot;
            }
            if (!finite) g = 0.f;
            gwbuf[j] = g;
            G += (double)g * (double)w;
        }
    }
    G = warp_sum(G);
    __syncwarp();
    // pass 2: d sigma_i = delta_i * (g_i T_{i+1} - sum_{k>i} g_k w_k)
    carry = 0.0;
    double pcarry = 0.0;
    for (int base = 0; base < S; base += 32) {
        const int j = base + lane;
        float t0 = 0.f, t1 = 0.f, sigma = 0.f;
        if (j < S) { t0 = __ldg(b + j); t1 = __ldg(b + j + 1); sigma = __ldg(density + n * S + j); }
        const float dl = __fsub_rn(t1, t0);
        const float dd = __fmul_rn(dl, sigma);
        float w, T;
        weight_step(dd, lane, carry, w, T);
        const float g = j < S ? gwbuf[j] : 0.f;
        const double P = warp_scan_incl((double)g * (double)w, lane) + pcarry;
        pcarry = __shfl_sync(0xffffffffu, P, 31);
        if (j < S) d_density[n * S + j] = dl * (float)((double)g * (double)(T * expf(-dd)) - (G - P));
    }
}

}  // namespace ps

using namespace ps;

static inline unsigned ray_blocks(int64_t N) { return (unsigned)cdiv(N, kRayWarps); }

extern "C" int ps_weights_fwd(const float* deltas, const float* density, int64_t N, int S, float* weights,
                              void* stream) {
    if (N == 0 || S == 0) return 0;
    PS_REQUIRE(deltas && density && weights, "weights_fwd: null pointer");
    weights_fwd_kernel<<<ray_blocks(N), kRayWarps * 32, 0, (cudaStream_t)stream>>>(deltas, density, N, S, weights);
    return check_launch("weights_fwd");
}

extern "C" int ps_weights_bwd(const float* deltas, const float* density, const float* dweights, int64_t N, int S,
                              float* ddensity, void* stream) {
    if (N == 0 || S == 0) return 0;
    PS_REQUIRE(deltas && density && dweights && ddensity, "weights_bwd: null pointer");
    weights_bwd_kernel<<<ray_blocks(N), kRayWarps * 32, 0, (cudaStream_t)stream>>>(deltas, density, dweights, N, S,
                                                                                  ddensity);
    return check_launch("weights_bwd");
}

extern "C" int ps_render_fwd(const float* weights, const float* values, int64_t N, int S, int C, float* out,
                             void* stream) {
    if (N == 0) return 0;
    PS_REQUIRE(weights && out, "render_fwd: null pointer");
    PS_REQUIRE(values != nullptr || C == 1, "render_fwd: values may be null only with C == 1");
    cudaStream_t s = (cudaStream_t)stream;
    if (C == 1)
        render_fwd_small_kernel<1><<<ray_blocks(N), kRayWarps * 32, 0, s>>>(weights, values, N, S, out);
    else if (C == 3)
        render_fwd_small_kernel<3><<<ray_blocks(N), kRayWarps * 32, 0, s>>>(weights, values, N, S, out);
    else
        render_fwd_wide_kernel<<<ray_blocks(N), kRayWarps * 32, 0, s>>>(weights, values, N, S, C, out);
    return check_launch("render_fwd");
}

extern "C" int ps_render_bwd(const float* weights, const float* values, const float* dout, int64_t N, int S, int C,
                             float* dweights, float* dvalues, void* stream) {
    if (N == 0 || S == 0) return 0;
    PS_REQUIRE(weights && dout && dweights, "render_bwd: null pointer");
    render_bwd_kernel<<<(unsigned)cdiv(N * S, 256), 256, 0, (cudaStream_t)stream>>>(weights, values, dout, N, S, C,
                                                                                    dweights, dvalues);
    return check_launch("render_bwd");
}

extern "C" int ps_depth_threshold(const float* weights, const float* eu_bins, int64_t N, int S, float threshold,
                                  float* depth, int64_t* index, void* stream) {
    if (N == 0) return 0;
    PS_REQUIRE(S >= 1 && weights && eu_bins && depth, "depth_threshold: bad arguments");
    depth_threshold_kernel<<<ray_blocks(N), kRayWarps * 32, 0, (cudaStream_t)stream>>>(weights, eu_bins, N, S, threshold,
                                                                                      depth, index);
    return check_launch("depth_threshold");
}

extern "C" int ps_composite_fwd(const float* eu_bins, const float* density, const float* rgb, const float* sem,
                                int64_t N, int S, int C, float threshold, float* weights, float* rgb_out, float* acc,
                                float* depth_exp, float* depth_thr, float* sem_out, float* tminmax, void* stream) {
    if (N == 0) return 0;
    PS_REQUIRE(S >= 1 && S <= 1024, "composite_fwd: S %d out of range [1,1024]", S);
    PS_REQUIRE(eu_bins && density, "composite_fwd: null pointer");
    PS_REQUIRE(sem == nullptr || (C >= 1 && C <= 128), "composite_fwd: C %d out of range", C);
    const size_t smem = (size_t)kRayWarps * S * sizeof(float);
    composite_fwd_kernel<<<ray_blocks(N), kRayWarps * 32, smem, (cudaStream_t)stream>>>(
        eu_bins, density, rgb, sem, N, S, C, threshold, weights, rgb_out, acc, depth_exp, depth_thr, sem_out, tminmax);
    return check_launch("composite_fwd");
}

extern "C" int ps_composite_bwd(const float* eu_bins, const float* density, const float* rgb, const float* sem,
                                const float* weights, const float* acc, const float* depth_exp, int64_t N, int S,
                                int C, const float* d_weights_in, const float* d_rgb_out, const float* d_acc,
                                const float* d_depth_exp, const float* d_sem_out, float* d_density, float* d_rgb,
                                float* d_sem, void* stream) {
    (void)weights;
    if (N == 0) return 0;
    PS_REQUIRE(S >= 1 && S <= 1024, "composite_bwd: S %d out of range [1,1024]", S);
    PS_REQUIRE(eu_bins && density && d_density, "composite_bwd: null pointer");
    PS_REQUIRE(d_depth_exp == nullptr || (acc && depth_exp), "composite_bwd: d_depth_exp needs acc and depth_exp");
    PS_REQUIRE(sem == nullptr || (C >= 4 && C <= 128 && C % 4 == 0), "composite_bwd: C %d must be a multiple of 4 <= 128",
               C);
    const int Cs = sem ? C : 0;
    const size_t smem = (size_t)kRayWarps * (((S + 3) & ~3) + Cs) * sizeof(float);
    composite_bwd_kernel<<<ray_blocks(N), kRayWarps * 32, smem, (cudaStream_t)stream>>>(
        eu_bins, density, rgb, sem, acc, depth_exp, N, S, Cs, d_weights_in, d_rgb_out, d_acc, d_depth_exp, d_sem_out,
        d_density, d_rgb, d_sem);
    return check_launch("composite_bwd");
}
