// Per-ray core of the z-anti-aliased (zip-NeRF) interlevel loss — the reference's default proposal loss
// (enable_z_anti_aliasing, models/PreSight/nerfacto_nusc_ms.py:129,293-295).  Plain C++ so that the SAME code is
// compiled into the CUDA kernel (csrc/losses.cu) and into a host harness the CPU test-suite checks against the live
// reference's fixture (tests/test_zaa_host.py).
//
// Reference arithmetic (model_components/PreSight/losses.py:127-206), for one ray and one proposal level:
//   wn      = w / (c[1:] - c[:-1])                                   final level's histogram as a density
//   xr, yr  = blur_stepfun(c, wn, r)                                 box blur of half-width r -> piecewise linear:
//               xr = sort(c - r, c + r);  y2 = +-(wn_pad[k] - wn_pad[k-1]) / (2 r) in that order, last one dropped;
//               yr = [0, clamp_min(cumsum((xr[1:] - xr[:-1]) * cumsum(y2)), 0)]
//   cdf     = [0, cumsum(0.5 (yr[1:] + yr[:-1]) (xr[1:] - xr[:-1]))]
//   W(x)    = cdf[j] + (x - xr[j]) (yr[j] + yr[j+1] o + yr[j] (1 - o)) / 2,  j = last knot <= x, o = (x - xr[j]) / (xr[j+1] - xr[j])
//   w_s     = diff(W(cp));   loss terms = max(w_s - wp, 0)^2 / (wp + 1e-5)
// torch's CPU cumsum carries a double accumulator and rounds every output to fp32; so does this code.
#pragma once

#ifndef PS_HD
#define PS_HD
#endif

namespace ps {
namespace zaa {

constexpr int kMaxS = 128;                 // samples per ray of the final level
constexpr int kMaxKnots = 2 * kMaxS + 2;

// c [S+1], w [S]: final level (constants); cp [Sp+1], wp [Sp]: proposal level; r: pulse width.
// Returns the ray's sum of loss terms; grad_wp [Sp] (nullable) receives d(sum)/d wp.
PS_HD inline float ray_loss(const float* c, const float* w, int S, const float* cp, const float* wp, int Sp, double r,
                            float* grad_wp) {
    float xr[kMaxKnots], yr[kMaxKnots], cdf[kMaxKnots], wn[kMaxS];
    const float rf = (float)r, two_r = (float)(2.0 * r);
    const int K = 2 * S + 2;
    for (int k = 0; k < S; ++k) wn[k] = w[k] / (c[k + 1] - c[k]);        // each bin's density once (every edge is visited
                                                                          // twice, and needs the bins on both sides)
    // ---- blur_stepfun: merge the two sorted edge lists (c - r first on ties), run the two nested cumsums ------------
    int ia = 0, ib = 0;
    double s1 = 0.0, s2 = 0.0;
    float x_prev = 0.f;
    yr[0] = 0.f;
    for (int k = 0; k < K; ++k) {
        const float a = ia <= S ? c[ia] - rf : 0.f, b = ib <= S ? c[ib] + rf : 0.f;
        const bool take_a = ia <= S && (ib > S || a <= b);
        const int e = take_a ? ia : ib;                                   // edge index of this knot
        const float x = take_a ? a : b;
        const float hi = e < S ? wn[e] : 0.f;                             // wn_pad[e]
        const float lo = e > 0 ? wn[e - 1] : 0.f;                         // wn_pad[e - 1]
        const float y1 = (hi - lo) / two_r;
        if (take_a) ++ia; else ++ib;
        xr[k] = x;
        if (k > 0) {
            // s1 holds cumsum(y2)[k - 1] (the knot BEFORE this one contributes the slope of this segment)
            const float term = (x - x_prev) * (float)s1;
            s2 += (double)term;
            const float v = (float)s2;
            yr[k] = v > 0.f ? v : 0.f;
        }
        s1 += (double)(take_a ? y1 : -y1);
        x_prev = x;
    }
    // ---- running integral of the piecewise-linear density ----------------------------------------------------------
    double s3 = 0.0;
    cdf[0] = 0.f;
    for (int k = 0; k + 1 < K; ++k) {
        const float area = (0.5f * (yr[k + 1] + yr[k])) * (xr[k + 1] - xr[k]);
        s3 += (double)area;
        cdf[k + 1] = (float)s3;
    }
    // ---- sorted_interp_quad at the proposal bin edges, differences, loss ------------------------------------------------
    int j = -1;                      // last knot <= x (queries are sorted, so j only moves forward)
    float prev = 0.f, total = 0.f;
    for (int m = 0; m <= Sp; ++m) {
        const float x = cp[m];
        while (j + 1 < K && xr[j + 1] <= x) ++j;
        float ret;
        if (j < 0) {
            ret = 0.f;               // left of the first knot: cdf[0] = yr[0] = 0
        } else {
            // The reference picks the two density values by the ARGMAX / ARGMIN of the masked running integral
            // (`find_interval(fcdf, return_idx=True)`, first index on ties), not by the bracketing knots: where the fp32
            // integral is flat over several knots (tiny areas absorbed by rounding) the left value comes from the FIRST
            // knot of the flat run, and where it is flat up to the end the right value is yr[0] = 0.  Knot positions and
            // the integral itself are taken by value (max / min), i.e. from the bracketing knots.
            const float x0 = xr[j], c0 = cdf[j];
            int i0 = j;
            while (i0 > 0 && cdf[i0 - 1] == c0) --i0;
            const float f0 = yr[i0];
            float f1, o;
            if (j + 1 < K) {
                f1 = cdf[K - 1] == cdf[j + 1] ? yr[0] : yr[j + 1];
                o = (x - x0) / (xr[j + 1] - x0);
                o = o != o ? 0.f : (o < 0.f ? 0.f : (o > 1.f ? 1.f : o));
            } else {                 // right of the last knot: the "next" knot degenerates to (x0, yr[0])
                f1 = yr[0];
                o = x > x0 ? 1.f : 0.f;
            }
            ret = c0 + (x - x0) * ((f0 + f1 * o) + f0 * (1.f - o)) / 2.f;
        }
        if (m > 0) {
            const float q = wp[m - 1];
            const float d = (ret - prev) - q;
            const float rr = d > 0.f ? d : 0.f;
            const float den = q + 1e-5f;
            total += rr * rr / den;
            if (grad_wp) grad_wp[m - 1] = -2.f * rr / den - rr * rr / (den * den);
        }
        prev = ret;
    }
    return total;
}

}  // namespace zaa
}  // namespace ps
