// Proposal ("interlevel") loss of mip-NeRF 360 as ONE kernel: loss and its gradient w.r.t. the proposal weights.
// Reference: model_components/losses.py:48-126 (outer / lossfun_outer / interlevel_loss).  One warp per ray.
//
//   w_outer_i = cy[idx_hi_i + 1] - cy[idx_lo_i],   cy = [0, cumsum(w_env)]
//   idx_lo_i  = clamp(searchsorted(t_env[:-1], c_i,   right) - 1, 0, Sp-1)
//   idx_hi_i  = clamp(searchsorted(t_env[1:],  c_i+1, right),     0, Sp-1)
//   loss_i    = max(w_i - w_outer_i, 0)^2 / (w_i + 1e-7)
// d loss_i / d w_env[k] = g_i * ([k <= idx_hi_i] - [k < idx_lo_i]),  g_i = -2 max(w_i - w_outer_i, 0) / (w_i + 1e-7),
// accumulated through two per-ray histograms (A over idx_hi, B over idx_lo) and suffix sums.
#include <cstdlib>

#include "sampler.cuh"
#define PS_HD __host__ __device__
#include "zaa_core.h"

namespace ps {

constexpr int kLossWarps = 4;

__global__ void __launch_bounds__(kLossWarps * 32) interlevel_kernel(const float* __restrict__ c,
                                                                     const float* __restrict__ w,
                                                                     const float* __restrict__ t_env,
                                                                     const float* __restrict__ w_env, int64_t N, int S,
                                                                     int Sp, float* __restrict__ loss_sum,
                                                                     float* __restrict__ grad_w_env) {
    extern __shared__ float smem[];  // per warp: te[Sp+1] | cy[Sp+1] | A[Sp] | B[Sp]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kLossWarps + warp;
    if (n >= N) return;
    float* te = smem + (size_t)warp * (4 * Sp + 2);
    float* cy = te + Sp + 1;
    float* A = cy + Sp + 1;
    float* B = A + Sp;
    for (int k = lane; k <= Sp; k += 32) te[k] = __ldg(t_env + n * (Sp + 1) + k);
    for (int k = lane; k < Sp; k += 32) { A[k] = 0.f; B[k] = 0.f; }
    // cy = [0, cumsum(w_env)] with fp64 accumulation (torch's CPU cumsum)
    double carry = 0.0;
    for (int base = 0; base < Sp; base += 32) {
        const int k = base + lane;
        const float v = k < Sp ? __ldg(w_env + n * Sp + k) : 0.f;
        const double incl = warp_scan_incl((double)v, lane) + carry;
        if (k < Sp) cy[k + 1] = (float)incl;
        carry = __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) cy[0] = 0.f;
    __syncwarp();
    float loss = 0.f;
    for (int i = lane; i < S; i += 32) {
        const float t0s = __ldg(c + n * (S + 1) + i), t0e = __ldg(c + n * (S + 1) + i + 1);
        const float wi = __ldg(w + n * S + i);
        int lo = upper_bound(te, Sp, t0s) - 1;          // starts = te[0..Sp-1]
        lo = min(max(lo, 0), Sp - 1);
        int hi = upper_bound(te + 1, Sp, t0e);          // ends   = te[1..Sp]
        hi = min(max(hi, 0), Sp - 1);
        const float w_outer = cy[hi + 1] - cy[lo];
        const float r = fmaxf(wi - w_outer, 0.f);
        const float inv = 1.f / (wi + 1.0e-7f);
        loss += r * r * inv;
        if (grad_w_env && r > 0.f) {
            const float g = -2.f * r * inv;
            atomicAdd(A + hi, g);
            atomicAdd(B + lo, g);
        }
    }
    loss = warp_sum(loss);
    if (lane == 0) atomicAdd(loss_sum, loss);
    if (!grad_w_env) return;
    __syncwarp();
    // grad[k] = sum_{m >= k} A[m] - sum_{m > k} B[m]   (reverse scans, 32 entries at a time from the top)
    float carryA = 0.f, carryB = 0.f;
    for (int top = ((Sp + 31) / 32) * 32; top > 0; top -= 32) {
        const int k = top - 1 - lane;                   // lane 0 handles the highest index of the chunk
        const float a = k < Sp ? A[k] : 0.f, b = k < Sp ? B[k] : 0.f;
        const float sa = warp_scan_incl(a, lane) + carryA;      // inclusive suffix sum of A at k
        const float sb = warp_scan_incl(b, lane) + carryB;      // inclusive suffix sum of B at k
        if (k < Sp && k >= 0) grad_w_env[n * Sp + k] = sa - (sb - b);
        carryA = __shfl_sync(0xffffffffu, sa, 31);
        carryB = __shfl_sync(0xffffffffu, sb, 31);
    }
}

// Distortion loss of mip-NeRF 360 with its gradient w.r.t. the weights, ONE kernel (model_components/losses.py:130-149):
//   u = bin mid-points;  loss_n = sum_i w_i sum_j w_j |u_i - u_j| + sum_i w_i^2 (t_{i+1} - t_i) / 3
//   d loss_n / d w_k = 2 sum_j w_j |u_k - u_j| + 2 w_k (t_{k+1} - t_k) / 3
// One warp per ray; the O(S^2) pair sum is evaluated as written (no sortedness assumption) from shared memory:
// S = 64 costs 4 096 multiply-adds per ray — 0.3 GFLOP for a 65 536-ray batch.
__global__ void __launch_bounds__(kLossWarps * 32) distortion_kernel(const float* __restrict__ c,
                                                                     const float* __restrict__ w, int64_t N, int S,
                                                                     float* __restrict__ loss_sum,
                                                                     float* __restrict__ grad_w) {
    extern __shared__ float smem[];  // per warp: u[S] | w[S]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kLossWarps + warp;
    if (n >= N) return;
    float* us = smem + (size_t)warp * 2 * S;
    float* ws = us + S;
    for (int k = lane; k < S; k += 32) {
        us[k] = __fdiv_rn(__fadd_rn(__ldg(c + n * (S + 1) + k + 1), __ldg(c + n * (S + 1) + k)), 2.f);
        ws[k] = __ldg(w + n * S + k);
    }
    __syncwarp();
    float loss = 0.f;
    for (int i = lane; i < S; i += 32) {
        const float ui = us[i], wi = ws[i];
        float a = 0.f;
        for (int j = 0; j < S; ++j) a = fmaf(ws[j], fabsf(ui - us[j]), a);
        const float dt = __fsub_rn(__ldg(c + n * (S + 1) + i + 1), __ldg(c + n * (S + 1) + i));
        loss += wi * a + wi * wi * dt / 3.f;
        if (grad_w) grad_w[n * S + i] = 2.f * a + 2.f * wi * dt / 3.f;
    }
    loss = warp_sum(loss);
    if (lane == 0) atomicAdd(loss_sum, loss);
}

// z-anti-aliased (zip-NeRF) interlevel loss, one proposal level per launch: one THREAD per ray runs zaa_core.h's
// sequential per-ray code (merge of the two shifted edge lists, the nested fp64-carried cumsums, a forward sweep over
// the sorted query edges).  First version (PS_ZAA_WARP=0): correctness against the reference's tie semantics first.
__global__ void __launch_bounds__(128) zaa_interlevel_kernel(const float* __restrict__ c, const float* __restrict__ w,
                                                             int64_t N, int S, const float* __restrict__ cp,
                                                             const float* __restrict__ wp, int Sp, double r,
                                                             float* __restrict__ loss_sum, float* __restrict__ grad_wp) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f;
    if (n < N)
        l = zaa::ray_loss(c + n * (S + 1), w + n * S, S, cp + n * (Sp + 1), wp + n * Sp, Sp, r,
                          grad_wp ? grad_wp + n * Sp : nullptr);
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0 && l != 0.f) atomicAdd(loss_sum, l);
}

// The default kernel (B200, 65 536 rays, S = 64, Sp = 128: 0.24 ms per level against 0.91 ms for the thread-per-ray
// kernel above, gpurun_out/r2_zaa_exp.txt; both pass the reference fixture): the same loss with one WARP per ray, following the loop-free formulation that tools/zaa_parallel_prototype.py checks
// against the reference's fixture — merge by rank ("c - r first on ties"; ranks found by short walks from the knot's own
// index), fp64 running sums over the knots (blocked per lane, one warp scan per pass), interval lookup by a branch-free
// binary search per query, flat-run lookup only where the integral is flat.
// Shared memory per warp (floats): c[S+1] | wn[S+1] | xr[K] | y2[K] | yr[K] | cdf[K] | ret[Sp+1],  K = 2S + 2.
constexpr int kZaaWarps = 4;

__device__ __forceinline__ int zaa_count_le(const float* a, int n, float v) {   // #{i : a[i] <= v}, a sorted
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int zaa_count_lt(const float* a, int n, float v) {   // #{i : a[i] < v}, a sorted
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kZaaWarps * 32, 12) zaa_interlevel_warp_kernel(
    const float* __restrict__ c, const float* __restrict__ w, int64_t N, int S, const float* __restrict__ cp,
    const float* __restrict__ wp, int Sp, double r, float* __restrict__ loss_sum, float* __restrict__ grad_wp) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kZaaWarps + warp;
    if (n >= N) return;                                      // warp-uniform; only __syncwarp below
    const int K = 2 * S + 2;
    float* cs = smem + (size_t)warp * (2 * (S + 1) + 4 * K + (Sp + 1));
    float* wn = cs + (S + 1);
    float* xr = wn + (S + 1);
    float* y2 = xr + K;
    float* yr = y2 + K;
    float* cdf = yr + K;
    float* ret = cdf + K;
    const float rf = (float)r, two_r = (float)(2.0 * r);
    for (int k = lane; k <= S; k += 32) cs[k] = __ldg(c + n * (S + 1) + k);
    __syncwarp();
    for (int k = lane; k <= S; k += 32) wn[k] = k < S ? __fdiv_rn(__ldg(w + n * S + k), __fsub_rn(cs[k + 1], cs[k])) : 0.f;
    __syncwarp();
    // ---- merge by rank ----------------------------------------------------------------------------------------------
    // knot a_e = c[e] - r goes behind the a-knots before it and the b-knots c[k] + r < a_e — all of which have k < e, and
    // (c sorted) form a prefix: walk down from e - 1 instead of searching (the pulse is narrower than most bins: 0-2 steps).
    // Likewise b_e = c[e] + r goes behind all a-knots c[k] - r <= b_e: every k <= e, and a run of k > e.
    for (int e = lane; e <= S; e += 32) {
        const float a = __fsub_rn(cs[e], rf), b = __fadd_rn(cs[e], rf);
        int k = e - 1;
        while (k >= 0 && !(__fadd_rn(cs[k], rf) < a)) --k;
        const int pos_a = e + (k + 1);
        k = e + 1;
        while (k <= S && __fsub_rn(cs[k], rf) <= b) ++k;
        const int pos_b = e + k;
        const float y1 = __fdiv_rn(__fsub_rn(wn[e], e > 0 ? wn[e - 1] : 0.f), two_r);     // wn[S] = 0 is the right pad
        xr[pos_a] = a; y2[pos_a] = y1;
        xr[pos_b] = b; y2[pos_b] = -y1;
    }
    __syncwarp();
    // ---- the two nested running sums (fp64 carry, fp32 outputs), then the running integral -------------------------------
    // Blocked: lane l owns the `per` consecutive intervals [l * per, (l + 1) * per) (per odd: conflict-free strides), sums
    // them serially and ONE warp scan per pass turns the lane totals into offsets (3 scans instead of 3 per chunk of 32).
    int per = (K - 1 + 31) / 32;
    per |= 1;
    const int k0 = lane * per, k1 = min(k0 + per, K - 1);
    if (lane == 0) { yr[0] = 0.f; cdf[0] = 0.f; }
    {
        double t = 0.0;
        for (int k = k0; k < k1; ++k) t += (double)y2[k];
        double run = warp_scan_incl(t, lane) - t;            // exclusive offset of this lane's block
        double t2 = 0.0;
        for (int k = k0; k < k1; ++k) {
            run += (double)y2[k];
            const float prod = __fmul_rn(__fsub_rn(xr[k + 1], xr[k]), (float)run);
            cdf[k + 1] = prod;                               // parked here until the third pass overwrites it
            t2 += (double)prod;
        }
        double run2 = warp_scan_incl(t2, lane) - t2;
        for (int k = k0; k < k1; ++k) {
            run2 += (double)cdf[k + 1];
            yr[k + 1] = fmaxf((float)run2, 0.f);
        }
    }
    __syncwarp();
    {
        double t = 0.0;
        for (int k = k0; k < k1; ++k)
            t += (double)__fmul_rn(__fmul_rn(0.5f, __fadd_rn(yr[k + 1], yr[k])), __fsub_rn(xr[k + 1], xr[k]));
        double run = warp_scan_incl(t, lane) - t;
        for (int k = k0; k < k1; ++k) {
            run += (double)__fmul_rn(__fmul_rn(0.5f, __fadd_rn(yr[k + 1], yr[k])), __fsub_rn(xr[k + 1], xr[k]));
            cdf[k + 1] = (float)run;
        }
    }
    __syncwarp();
    // ---- sorted_interp_quad at the proposal bin edges -------------------------------------------------------------------
    const float cdf_last = cdf[K - 1], yr_first = yr[0];
    const int top = 1 << (31 - __clz(K));
    for (int m = lane; m <= Sp; m += 32) {
        const float x = __ldg(cp + n * (Sp + 1) + m);
        int j = 0;                                           // #{knots <= x} by a branch-free descent, then - 1
        for (int step = top; step >= 1; step >>= 1)
            if (j + step <= K && xr[j + step - 1] <= x) j += step;
        j -= 1;                                              // last knot <= x
        float v = 0.f;
        if (j >= 0) {
            const float x0 = xr[j], c0 = cdf[j];
            // first knot of the flat run of the integral (argmax on ties): the knot itself unless the integral is flat here
            const int jf = (j == 0 || cdf[j - 1] < c0) ? j : zaa_count_lt(cdf, K, c0);
            const float f0 = yr[jf];
            float f1, o;
            if (j + 1 < K) {
                f1 = cdf_last == cdf[j + 1] ? yr_first : yr[j + 1];
                o = __fdiv_rn(__fsub_rn(x, x0), __fsub_rn(xr[j + 1], x0));
                o = o != o ? 0.f : fminf(fmaxf(o, 0.f), 1.f);
            } else {
                f1 = yr_first;
                o = x > x0 ? 1.f : 0.f;
            }
            const float inner = __fadd_rn(__fadd_rn(f0, __fmul_rn(f1, o)), __fmul_rn(f0, __fsub_rn(1.f, o)));
            v = __fadd_rn(c0, __fdiv_rn(__fmul_rn(__fsub_rn(x, x0), inner), 2.f));
        }
        ret[m] = v;
    }
    __syncwarp();
    float loss = 0.f;
    for (int m = lane; m < Sp; m += 32) {
        const float q = __ldg(wp + n * Sp + m);
        const float d = __fsub_rn(__fsub_rn(ret[m + 1], ret[m]), q);
        const float rr = d > 0.f ? d : 0.f;
        const float den = q + 1e-5f;
        loss += rr * rr / den;
        if (grad_wp) grad_wp[n * Sp + m] = -2.f * rr / den - rr * rr / (den * den);
    }
    loss = warp_sum(loss);
    if (lane == 0 && loss != 0.f) atomicAdd(loss_sum, loss);
}

}  // namespace ps

using namespace ps;

extern "C" int ps_zaa_interlevel_loss(const float* c, const float* w, int64_t N, int S, const float* t_env,
                                      const float* w_env, int Sp, double pulse_width, float* loss_sum, float* grad_w_env,
                                      void* stream) {
    if (N == 0) return 0;
    PS_REQUIRE(c && w && t_env && w_env && loss_sum, "zaa_interlevel_loss: null pointer");
    PS_REQUIRE(S >= 1 && S <= zaa::kMaxS, "zaa_interlevel_loss: final-level samples per ray %d out of range [1,%d]", S,
               zaa::kMaxS);
    PS_REQUIRE(Sp >= 1, "zaa_interlevel_loss: proposal samples per ray %d < 1", Sp);
    PS_REQUIRE(pulse_width > 0.0, "zaa_interlevel_loss: pulse width must be positive");
    // PS_ZAA_WARP=0 selects the first, thread-per-ray kernel (kept as a second implementation of the same arithmetic)
    static const bool warp_per_ray = getenv("PS_ZAA_WARP") == nullptr || atoi(getenv("PS_ZAA_WARP")) != 0;
    if (warp_per_ray) {
        const size_t smem = (size_t)kZaaWarps * (2 * (S + 1) + 4 * (2 * S + 2) + (Sp + 1)) * sizeof(float);
        PS_REQUIRE(smem <= 200 * 1024, "zaa_interlevel_loss: %zu bytes of shared memory", smem);
        static bool configured = false;
        if (!configured) {
            cudaFuncSetAttribute(zaa_interlevel_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            configured = true;
        }
        zaa_interlevel_warp_kernel<<<(unsigned)cdiv(N, kZaaWarps), kZaaWarps * 32, smem, (cudaStream_t)stream>>>(
            c, w, N, S, t_env, w_env, Sp, pulse_width, loss_sum, grad_w_env);
        return check_launch("zaa_interlevel_loss(warp)");
    }
    zaa_interlevel_kernel<<<(unsigned)cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>(c, w, N, S, t_env, w_env, Sp,
                                                                                  pulse_width, loss_sum, grad_w_env);
    return check_launch("zaa_interlevel_loss");
}

extern "C" int ps_distortion_loss(const float* c, const float* w, int64_t N, int S, float* loss_sum, float* grad_w,
                                  void* stream) {
    if (N == 0) return 0;
    PS_REQUIRE(c && w && loss_sum, "distortion_loss: null pointer");
    PS_REQUIRE(S >= 1 && S <= 1024, "distortion_loss: samples per ray %d out of range [1,1024]", S);
    const size_t smem = (size_t)kLossWarps * 2 * S * sizeof(float);
    distortion_kernel<<<(unsigned)cdiv(N, kLossWarps), kLossWarps * 32, smem, (cudaStream_t)stream>>>(c, w, N, S, loss_sum,
                                                                                                   grad_w);
    return check_launch("distortion_loss");
}

extern "C" int ps_interlevel_loss(const float* c, const float* w, const float* t_env, const float* w_env, int64_t N,
                                  int S, int Sp, float* loss_sum, float* grad_w_env, void* stream) {
    if (N == 0) return 0;
    PS_REQUIRE(c && w && t_env && w_env && loss_sum, "interlevel_loss: null pointer");
    PS_REQUIRE(S >= 1 && Sp >= 1 && Sp <= 2048, "interlevel_loss: sample counts out of range");
    const size_t smem = (size_t)kLossWarps * (4 * Sp + 2) * sizeof(float);
    interlevel_kernel<<<(unsigned)cdiv(N, kLossWarps), kLossWarps * 32, smem, (cudaStream_t)stream>>>(
        c, w, t_env, w_env, N, S, Sp, loss_sum, grad_w_env);
    return check_launch("interlevel_loss");
}
