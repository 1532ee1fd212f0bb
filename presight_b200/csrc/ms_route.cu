// Sub-field routing on the device: the nearest-centroid routers of
//   fields/PreSight/ingp_field_ms.py:80-126 (cdist(points, centroids).argmin, one boolean-mask gather + field call +
//   masked scatter per sub-field, a `torch.any` host sync each), prop_density_field_ms.py:86-105
// as three small kernels with no host synchronisation, feeding the sub-field-homogeneous point tiles of the fused level
// kernels (prop_tc5_ms.cu / field_tc5_ms.cu):
//   1. route    every sample point -> nearest centroid (uint8) + a histogram per block of 256 points
//   2. plan     block histograms -> (one CTA per sub-field) the first row of every block inside its sub-field's segment and
//               the sub-field totals; then segment starts padded to whole tiles and the tile -> sub-field table
//   3. scatter  every point -> its row in its sub-field's segment: perm[row] = point, unit-cube position normalised
//               with THAT sub-field's aabb (fields/PreSight/utils.py:6-10 + contraction) and the in-box selector
// The sort is STABLE (a counting sort with per-block offsets and in-block ranks): inside a sub-field's segment the
// points keep their original order, so consecutive samples of a ray stay neighbours — that is what the warp
// pre-aggregation of the hash scatter-add and the gather's cache locality live on — and the result is deterministic.
// A level's points are then P_pad <= P + nf * pad rows in sub-field order; row -> point through perm (-1 = padding).
#include "position.cuh"

namespace ps {

constexpr int kMaxSub = PS_MAX_FIELDS;

__device__ __forceinline__ void point_of(const float* __restrict__ positions, const float* __restrict__ origins,
                                         const float* __restrict__ dirs, const float* __restrict__ eu, int64_t p, int S,
                                         float (&x)[3]) {
    if (positions) {
        x[0] = __ldg(positions + 3 * p); x[1] = __ldg(positions + 3 * p + 1); x[2] = __ldg(positions + 3 * p + 2);
        return;
    }
    const int64_t ray = p / S;
    const int s = (int)(p - ray * S);
    const float o[3] = {__ldg(origins + 3 * ray), __ldg(origins + 3 * ray + 1), __ldg(origins + 3 * ray + 2)};
    const float d[3] = {__ldg(dirs + 3 * ray), __ldg(dirs + 3 * ray + 1), __ldg(dirs + 3 * ray + 2)};
    frustum_midpoint(o, d, __ldg(eu + ray * (S + 1) + s), __ldg(eu + ray * (S + 1) + s + 1), x);
}

constexpr int kRouteThreads = 256;

// rank of this thread among the threads of its block with the same key, in thread order (all threads must call);
// also returns the block's count of that key through `hist` (shared, [kMaxSub], zeroed by the caller before the call)
__device__ __forceinline__ int block_rank(int key, bool active, int32_t (*warp_hist)[kMaxSub], int nf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned peers = __match_any_sync(0xffffffffu, active ? key : -1 - lane);
    const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    for (int k = lane; k < nf; k += 32) warp_hist[warp][k] = 0;
    __syncwarp();
    if (active && rank_in_warp == 0) warp_hist[warp][key] = __popc(peers);
    __syncthreads();
    int before = 0;
    if (active)
        for (int w = 0; w < warp; ++w) before += warp_hist[w][key];
    return before + rank_in_warp;
}

__global__ void __launch_bounds__(kRouteThreads) ms_route_kernel(const float* __restrict__ positions,
                                                                 const float* __restrict__ origins,
                                                                 const float* __restrict__ dirs, const float* __restrict__ eu,
                                                                 int64_t P, int S, const float* __restrict__ cent, int nf,
                                                                 uint8_t* __restrict__ sf_out, int32_t* __restrict__ block_hist) {
    __shared__ float c[kMaxSub][3];
    __shared__ int32_t warp_hist[kRouteThreads / 32][kMaxSub];
    for (int i = threadIdx.x; i < nf * 3; i += blockDim.x) c[i / 3][i % 3] = cent[i];
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = p < P;
    int arg = 0;
    if (on) {
        float x[3];
        point_of(positions, origins, dirs, eu, p, S, x);
        float best = INFINITY;
        for (int j = 0; j < nf; ++j) {
            const float dx = x[0] - c[j][0], dy = x[1] - c[j][1], dz = x[2] - c[j][2];
            const float d = sqrtf(dx * dx + dy * dy + dz * dz);          // same arithmetic as ps_nearest_centroid
            if (d < best) { best = d; arg = j; }
        }
        sf_out[p] = (uint8_t)arg;
    }
    block_rank(arg, on, warp_hist, nf);
    for (int k = threadIdx.x; k < nf; k += blockDim.x) {
        int32_t tot = 0;
        for (int w = 0; w < kRouteThreads / 32; ++w) tot += warp_hist[w][k];
        block_hist[(int64_t)k * gridDim.x + blockIdx.x] = tot;          // [nf][nblocks]: a sub-field's column is contiguous
    }
}

// CTA k: exclusive scan of sub-field k's block counts (block_hist[k][*] is overwritten by each block's first row RELATIVE to
// the segment start), total -> totals[k]
__global__ void __launch_bounds__(1024) ms_scan_kernel(int32_t* __restrict__ block_hist, int64_t nblocks,
                                                       int32_t* __restrict__ totals) {
    __shared__ int32_t part[1024];
    const int tid = threadIdx.x;
    int32_t* col = block_hist + (int64_t)blockIdx.x * nblocks;
    const int64_t chunk = (nblocks + blockDim.x - 1) / blockDim.x;
    const int64_t b0 = (int64_t)tid * chunk, b1 = b0 + chunk < nblocks ? b0 + chunk : nblocks;
    int32_t sum = 0;
    for (int64_t b = b0; b < b1; ++b) sum += col[b];
    part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {                       // inclusive scan of the per-thread partial sums
        const int32_t v = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int32_t run = part[tid] - sum;
    for (int64_t b = b0; b < b1; ++b) {
        const int32_t cnt = col[b];
        col[b] = run;
        run += cnt;
    }
    if (tid == 1023) totals[blockIdx.x] = part[1023];
}

// one CTA: padded segment starts, tile -> sub-field table (255 = beyond the last segment)
__global__ void __launch_bounds__(1024) ms_plan_kernel(const int32_t* __restrict__ totals, int nf, int pad, int tile_rows,
                                                       int64_t max_rows, int32_t* __restrict__ seg_start,
                                                       uint8_t* __restrict__ tile_sf) {
    __shared__ int32_t start[kMaxSub + 1];
    if (threadIdx.x == 0) {
        int32_t acc = 0;
        for (int k = 0; k < nf; ++k) {
            start[k] = acc;
            acc += (totals[k] + pad - 1) / pad * pad;
        }
        start[nf] = acc;
        for (int k = 0; k <= nf; ++k) seg_start[k] = start[k];
    }
    __syncthreads();
    const int64_t ntiles = max_rows / tile_rows;
    for (int64_t t = threadIdx.x; t < ntiles; t += blockDim.x) {
        const int64_t row = t * tile_rows;
        int sf = 255;
        for (int k = 0; k < nf; ++k)
            if (row >= start[k] && row < start[k + 1]) sf = k;
        tile_sf[t] = (uint8_t)sf;
    }
}

struct SubBoxes {
    Aabb box[kMaxSub];
};

__global__ void __launch_bounds__(kRouteThreads) ms_scatter_kernel(const float* __restrict__ positions, const float* __restrict__ origins,
                                                         const float* __restrict__ dirs, const float* __restrict__ eu,
                                                         int64_t P, int S, const uint8_t* __restrict__ sf_in,
                                                         const float* __restrict__ aabbs, int nf, int contract,
                                                         const int32_t* __restrict__ block_off,
                                                         const int32_t* __restrict__ seg_start, int32_t* __restrict__ perm,
                                                         float* __restrict__ x01s, uint8_t* __restrict__ sels) {
    __shared__ Aabb boxes[kMaxSub];
    __shared__ int32_t warp_hist[kRouteThreads / 32][kMaxSub];
    for (int i = threadIdx.x; i < nf * 6; i += blockDim.x) {
        const int k = i / 6, j = i % 6;
        if (j < 3) boxes[k].lo[j] = aabbs[i]; else boxes[k].hi[j - 3] = aabbs[i];
    }
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = p < P;
    const int sf = on ? sf_in[p] : 0;
    const int rank = block_rank(sf, on, warp_hist, nf);
    if (!on) return;
    const int slot = seg_start[sf] + block_off[(int64_t)sf * gridDim.x + blockIdx.x] + rank;
    float x[3];
    point_of(positions, origins, dirs, eu, p, S, x);
    const bool inside = normalize_point(x, boxes[sf], contract != 0);
    perm[slot] = (int32_t)p;
    x01s[3 * (int64_t)slot] = x[0];
    x01s[3 * (int64_t)slot + 1] = x[1];
    x01s[3 * (int64_t)slot + 2] = x[2];
    sels[slot] = inside ? 1 : 0;
}

}  // namespace ps

using namespace ps;

extern "C" int ps_ms_route(const float* positions, const float* origins, const float* dirs, const float* eu_bins, int64_t P,
                           int S, const float* centroids, int nf, uint8_t* sf_out, int32_t* block_hist, void* stream) {
    if (P == 0) return 0;
    PS_REQUIRE(nf >= 1 && nf <= kMaxSub, "ms_route: %d sub-fields outside [1, %d]", nf, kMaxSub);
    PS_REQUIRE(centroids && sf_out && block_hist, "ms_route: null pointer");
    PS_REQUIRE(positions != nullptr || (origins && dirs && eu_bins && S >= 1), "ms_route: give positions or rays + bins");
    PS_REQUIRE(P < (1ll << 31), "ms_route: too many points");
    ms_route_kernel<<<(unsigned)cdiv(P, kRouteThreads), kRouteThreads, 0, (cudaStream_t)stream>>>(
        positions, origins, dirs, eu_bins, P, S, centroids, nf, sf_out, block_hist);
    return check_launch("ms_route");
}

extern "C" int ps_ms_plan(int32_t* block_hist, int64_t P, int nf, int pad, int tile_rows, int64_t max_rows,
                          int32_t* seg_start, uint8_t* tile_sf, void* stream) {
    PS_REQUIRE(block_hist && seg_start && tile_sf, "ms_plan: null pointer");
    PS_REQUIRE(nf >= 1 && nf <= kMaxSub, "ms_plan: %d sub-fields outside [1, %d]", nf, kMaxSub);
    PS_REQUIRE(tile_rows >= 1 && pad >= tile_rows && pad % tile_rows == 0 && max_rows % tile_rows == 0,
               "ms_plan: pad %d must be a multiple of the tile (%d rows) and max_rows a multiple of the tile", pad, tile_rows);
    PS_REQUIRE(max_rows >= (P + pad - 1) / pad * pad + (int64_t)nf * pad, "ms_plan: max_rows %lld too small for %lld points",
               (long long)max_rows, (long long)P);
    // totals live behind the nf + 1 segment starts (seg_start has room for 2 * nf + 1 ints)
    int32_t* totals = seg_start + nf + 1;
    ms_scan_kernel<<<nf, 1024, 0, (cudaStream_t)stream>>>(block_hist, cdiv(P, kRouteThreads), totals);
    if (int e = check_launch("ms_plan(scan)")) return e;
    ms_plan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(totals, nf, pad, tile_rows, max_rows, seg_start, tile_sf);
    return check_launch("ms_plan");
}

extern "C" int ps_ms_scatter(const float* positions, const float* origins, const float* dirs, const float* eu_bins, int64_t P,
                             int S, const uint8_t* sf, const float* aabbs, int nf, int contract, const int32_t* block_off,
                             const int32_t* seg_start, int32_t* perm, float* x01_sorted, uint8_t* sel_sorted, void* stream) {
    if (P == 0) return 0;
    PS_REQUIRE(sf && aabbs && block_off && seg_start && perm && x01_sorted && sel_sorted, "ms_scatter: null pointer");
    PS_REQUIRE(nf >= 1 && nf <= kMaxSub, "ms_scatter: %d sub-fields outside [1, %d]", nf, kMaxSub);
    PS_REQUIRE(positions != nullptr || (origins && dirs && eu_bins && S >= 1), "ms_scatter: give positions or rays + bins");
    ms_scatter_kernel<<<(unsigned)cdiv(P, kRouteThreads), kRouteThreads, 0, (cudaStream_t)stream>>>(
        positions, origins, dirs, eu_bins, P, S, sf, aabbs, nf, contract, block_off, seg_start, perm, x01_sorted, sel_sorted);
    return check_launch("ms_scatter");
}
