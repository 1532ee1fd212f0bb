// Kernel #3: initial spaced sampler and inverse-CDF (PDF) resampling, one warp per ray.
#include "sampler.cuh"

namespace ps {

// ---- SpacedSampler (ray_samplers.py:98-128) ---------------------------------------------
__global__ void __launch_bounds__(256) spaced_bins_kernel(const float* __restrict__ nears,
                                                          const float* __restrict__ fars,
                                                          const float* __restrict__ lin, const float* __restrict__ t_rand,
                                                          int64_t N, int S, float thr, float* __restrict__ sp,
                                                          float* __restrict__ eu) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int nb = S + 1;
    if (i >= N * nb) return;
    const int64_t n = i / nb;
    const int j = (int)(i - n * nb);
    float b = __ldg(lin + j);
    if (t_rand) {
        // stratified single jitter: bins = lower + (upper - lower) * t, with lower/upper the bin centres
        const float lower = j == 0 ? b : __fdiv_rn(__fadd_rn(b, __ldg(lin + j - 1)), 2.f);
        const float upper = j == S ? b : __fdiv_rn(__fadd_rn(__ldg(lin + j + 1), b), 2.f);
        b = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), __ldg(t_rand + n)));
    }
    const float s_near = spacing_fn(__ldg(nears + n), thr), s_far = spacing_fn(__ldg(fars + n), thr);
    sp[i] = b;
    eu[i] = spacing_to_euclidean(b, s_near, s_far, thr);
}

// ---- PDFSampler (ray_samplers.py:305-362) -----------------------------------------------
// smem per warp: cdf[S_in+1] | bins_in[S_in+1]
constexpr int kPdfWarps = 4;

__global__ void __launch_bounds__(kPdfWarps * 32) pdf_resample_kernel(
    const float* __restrict__ weights, const float* __restrict__ sp_in, const float* __restrict__ u_base,
    const float* __restrict__ rand, const float* __restrict__ nears, const float* __restrict__ fars, int64_t N,
    int S_in, int S_out, float padding, float eps, float anneal, float thr, float* __restrict__ sp_out,
    float* __restrict__ eu_out, int64_t* __restrict__ inds_out, float* __restrict__ cdf_out,
    float* __restrict__ u_out) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * kPdfWarps + warp;
    if (n >= N) return;
    const int nc = S_in + 1;
    float* cdf = smem + (size_t)warp * 2 * nc;
    float* bins = cdf + nc;
    const float* w_row = weights + n * S_in;

    // weights + padding, their sum (RS:305-308).  fp64 accumulation = correctly rounded fp32 sum.
    double part = 0.0;
    for (int j = lane; j < S_in; j += 32) {
        float w = __ldg(w_row + j);
        if (anneal != 1.f) w = powf(w, anneal);
        w = __fadd_rn(w, padding);
        cdf[j + 1] = w;  // stash
        part += (double)w;
    }
    for (int j = lane; j < nc; j += 32) bins[j] = __ldg(sp_in + n * nc + j);
    float wsum = (float)warp_sum(part);
    const float pad = fmaxf(__fsub_rn(eps, wsum), 0.f);         // relu(eps - sum)      (RS:309)
    const float pad_each = __fdiv_rn(pad, (float)S_in);         // padding / S          (RS:310)
    wsum = __fadd_rn(wsum, pad);                                // (RS:311)
    __syncwarp();
    // pdf = w / sum; cdf = min(1, cumsum(pdf)) with fp64 accumulation like torch's CPU cumsum (RS:313-315)
    double carry = 0.0;
    for (int base = 0; base < S_in; base += 32) {
        const int j = base + lane;
        float pdf = 0.f;
        if (j < S_in) pdf = __fdiv_rn(__fadd_rn(cdf[j + 1], pad_each), wsum);
        const double incl = warp_scan_incl((double)pdf, lane) + carry;
        if (j < S_in) cdf[j + 1] = fminf(1.f, (float)incl);
        carry = __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) cdf[0] = 0.f;
    __syncwarp();
    if (cdf_out)
        for (int j = lane; j < nc; j += 32) cdf_out[n * nc + j] = cdf[j];

    const int nb = S_out + 1;
    const float s_near = spacing_fn(__ldg(nears + n), thr), s_far = spacing_fn(__ldg(fars + n), thr);
    // train: u = linspace + rand/nb ; eval: u = linspace + 1/(2 nb)   (RS:317-331)
    const float shift = rand ? __fdiv_rn(__ldg(rand + n), (float)nb) : (float)(1.0 / (2.0 * (double)nb));
    for (int j = lane; j < nb; j += 32) {
        const float u = __fadd_rn(__ldg(u_base + j), shift);
        const int ind = upper_bound(cdf, nc, u);                 // searchsorted(side="right")   (RS:345)
        const int below = min(max(ind - 1, 0), S_in), above = min(max(ind, 0), S_in);
        const float c0 = cdf[below], c1 = cdf[above], b0 = bins[below], b1 = bins[above];
        float t = __fdiv_rn(__fsub_rn(u, c0), __fsub_rn(c1, c0));
        t = fminf(fmaxf(nan_to_num(t), 0.f), 1.f);               // clip(nan_to_num(.,0),0,1)    (RS:353)
        const float b = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
        sp_out[n * nb + j] = b;
        eu_out[n * nb + j] = spacing_to_euclidean(b, s_near, s_far, thr);
        if (inds_out) inds_out[n * nb + j] = ind;
        if (u_out) u_out[n * nb + j] = u;
    }
}

__global__ void __launch_bounds__(256) searchsorted_right_kernel(const float* __restrict__ cdf,
                                                                 const float* __restrict__ u, int64_t N, int nc,
                                                                 int nu, int64_t* __restrict__ inds) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * nu) return;
    const int64_t n = i / nu;
    inds[i] = upper_bound(cdf + n * nc, nc, u[i]);
}

}  // namespace ps

using namespace ps;

extern "C" int ps_spaced_bins(const float* nears, const float* fars, const float* lin_bins, const float* t_rand,
                              int64_t N, int S, float thr, float* sp_bins, float* eu_bins, void* stream) {
    if (N == 0) return 0;
    PS_REQUIRE(S >= 1, "spaced_bins: num_samples must be >= 1");
    PS_REQUIRE(nears && fars && lin_bins && sp_bins && eu_bins, "spaced_bins: null pointer");
    spaced_bins_kernel<<<(unsigned)cdiv(N * (S + 1), 256), 256, 0, (cudaStream_t)stream>>>(nears, fars, lin_bins, t_rand,
                                                                                           N, S, thr, sp_bins, eu_bins);
    return check_launch("spaced_bins");
}

extern "C" int ps_pdf_resample(const float* weights, const float* sp_in, const float* u_base, const float* rand,
                               const float* nears, const float* fars, int64_t N, int S_in, int S_out, float padding,
                               float eps, float anneal, float thr, float* sp_out, float* eu_out, int64_t* inds,
                               float* cdf, float* u, void* stream) {
    if (N == 0) return 0;
    PS_REQUIRE(S_in >= 1 && S_in <= 1024, "pdf_resample: S_in %d out of range [1,1024]", S_in);
    PS_REQUIRE(S_out >= 1, "pdf_resample: num_samples must be >= 1");
    PS_REQUIRE(weights && sp_in && u_base && nears && fars && sp_out && eu_out, "pdf_resample: null pointer");
    const size_t smem = (size_t)kPdfWarps * 2 * (S_in + 1) * sizeof(float);
    pdf_resample_kernel<<<(unsigned)cdiv(N, kPdfWarps), kPdfWarps * 32, smem, (cudaStream_t)stream>>>(
        weights, sp_in, u_base, rand, nears, fars, N, S_in, S_out, padding, eps, anneal, thr, sp_out, eu_out, inds, cdf,
        u);
    return check_launch("pdf_resample");
}

extern "C" int ps_searchsorted_right(const float* cdf, const float* u, int64_t N, int n_cdf, int n_u, int64_t* inds,
                                     void* stream) {
    if (N == 0 || n_u == 0) return 0;
    PS_REQUIRE(cdf && u && inds, "searchsorted_right: null pointer");
    searchsorted_right_kernel<<<(unsigned)cdiv(N * n_u, 256), 256, 0, (cudaStream_t)stream>>>(cdf, u, N, n_cdf, n_u,
                                                                                              inds);
    return check_launch("searchsorted_right");
}
