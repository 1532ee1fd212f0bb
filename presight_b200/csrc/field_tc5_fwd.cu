// Fused field level, forward: hash features -> base MLP -> density -> semantic head -> colour head -> compositing,
// one kernel, nothing but the per-ray results and the weights written to HBM.
// Reference: fields/PreSight/ingp_field.py:163-251, cameras/rays.py:128-150, model_components/renderers.py:70-117,
// 286-314, 332-383, models/PreSight/nerfacto_nusc_ms.py:497-530.
//
// CTA = 256 threads = two independent groups of 128; a group owns one 128-point tile at a time (128 / S rays) with
// one thread per point (row).  All eight layers run on tcgen05.mma (bf16 x bf16 -> fp32 in TMEM, M = 128) in five
// GEMM -> epilogue phases per tile (base 0, base 1, then the colour and the semantic head side by side); the
// epilogues (ReLU, bf16 re-pack into the next layer's A tile; the bias is one more K step of the GEMM) are thread-per-row, so everything that is "per
// sample" — density, weights, the dot products of compositing — is plain per-thread code.  While one group waits for
// its MMA the other runs its epilogue; the weights (54 KB bf16) are staged once per CTA and shared by both groups.
// (Prefetching each group's next tile of features with bulk copies, as the backward kernel does, was measured and changes
// nothing here — 0.495 vs 0.485 ms for 32 768 rays: the other group's work already covers a group's load latency.)
#include "field_tc5.cuh"

// phase clocks of the debug build: see field_tc5_bwd.cu / tools/phase_clocks.py
#ifdef PS_PHASE_CLOCKS
__device__ long long* g_phase_buf_fwd = nullptr;
#define PS_STAMP(code)                                     \
    do {                                                   \
        if (stamp_on && nstamp < 250) {                    \
            g_phase_buf_fwd[2 * nstamp] = (code);          \
            g_phase_buf_fwd[2 * nstamp + 1] = clock64();   \
            ++nstamp;                                      \
        }                                                  \
    } while (0)
#else
#define PS_STAMP(code)
#endif

namespace ps {
namespace ftc5 {

constexpr int kFwdThreads = 256;
constexpr int kGroups = 2;

template <int K0>
struct FwdSmem {
    using WL = WLayout<K0>;
    // per group
    static constexpr uint32_t h = 0;                                    // [128 x 80]
    static constexpr uint32_t shapp = h + cm_bytes(kRows, kBaseOut);    // [128 x 32]
    static constexpr uint32_t bufa = shapp + cm_bytes(kRows, 32);       // [128 x 64]  X0 / S1 / R1
    static constexpr uint32_t bufb = bufa + cm_bytes(kRows, 64);        // [128 x 64]  H1 / S2 / R2
    static constexpr uint32_t red = bufb + cm_bytes(kRows, 64);         // float [4 warps][72]
    static constexpr uint32_t tails = red + 4 * 72 * 4;                 // double [4][2]
    static constexpr uint32_t found = tails + 4 * 2 * 8;                // int [4]
    static constexpr uint32_t group_bytes = ((found + 16 + 127) / 128) * 128;
    static constexpr uint32_t groups = ((WL::end + 127) / 128) * 128;
    static constexpr uint32_t bars = groups + kGroups * group_bytes;    // 2 mbarriers + tmem slot
    static constexpr uint32_t total = bars + 32;
};

// named barrier of thread group g (immediate ids: tc5.cuh)
__device__ __forceinline__ void group_sync(int g) {
    if (g == 0) bar_sync<1, 128>();
    else bar_sync<2, 128>();
}

// epilogue of a hidden layer: accumulator row (bias already added by the GEMM) -> ReLU -> bf16 -> columns [0, 64)
__device__ __forceinline__ void hidden_epilogue64(uint32_t trow, unsigned char* tile, int r) {
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
        float v[32];
        tmem_ld32_nowait(trow + c, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; i += 8) store_chunk_relu(tile, kRows, r, c + i, v + i);
    }
}

// sum v[0..n) over the 32 lanes with a butterfly that halves the live values each step; afterwards lane l holds the
// warp totals of channels  l * n / 32 + j  in v[j], j < n / 32  (n = 64 -> two channels per lane)
template <int NV>
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[NV], int lane) {
    static_assert(NV == 64 || NV == 32, "NV");
#pragma unroll
    for (int step = 0, half = NV / 2; step < 5; ++step, half >>= 1) {
        const int bit = 16 >> step;
        const bool upper = (lane & bit) != 0;
#pragma unroll
        for (int j = 0; j < NV / 2; ++j) {
            if (j < half) {
                // keep the half selected by this lane's bit, send the other half to the partner
                const float keep = upper ? v[j + half] : v[j];
                const float send = upper ? v[j] : v[j + half];
                v[j] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
            }
        }
    }
}

template <int K0>
__global__ void __launch_bounds__(kFwdThreads, 1) field_fwd_kernel(FieldArgs a) {
    using WL = WLayout<K0>;
    using SM = FwdSmem<K0>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, g = tid >> 7, t = tid & 127, warp = t >> 5, lane = tid & 31;
    unsigned char* wbase = smem;
    unsigned char* gs = smem + SM::groups + g * SM::group_bytes;
    unsigned char* Ht = gs + SM::h;
    unsigned char* SHAPPt = gs + SM::shapp;
    unsigned char* BufA = gs + SM::bufa;
    unsigned char* BufB = gs + SM::bufb;
    float* red = reinterpret_cast<float*>(gs + SM::red);
    double* tails = reinterpret_cast<double*>(gs + SM::tails);
    int* foundw = reinterpret_cast<int*>(gs + SM::found);
    uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(smem + SM::bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::bars + 16);

    load_all_weights<K0>(a.net, wbase, tid, kFwdThreads);
    if (tid < 32) tmem_alloc(tmem_slot, 256);
    if (tid == 0) {
        mbar_init(smem_u32(bar_ptr), 1);
        mbar_init(smem_u32(bar_ptr + 1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot + g * 128;                        // this group's 80-column accumulator
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t bar = smem_u32(bar_ptr + g);
    const uint32_t wb = smem_u32(wbase), ones = wb + WL::ones;
    const uint32_t aH = smem_u32(Ht), aSH = smem_u32(SHAPPt), aA = smem_u32(BufA), aB = smem_u32(BufB);
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);      // provably warp-uniform copy (issue branch)
    uint32_t phase = 0;

    const int S = a.S;
    const int rpt = kRows / S;                 // rays per tile (S is a multiple of 32, <= 128)
    const int rows_used = rpt * S;
    const int wpr = S / 32;                    // warps per ray
    const int64_t P = a.N * S;
    const int64_t ntiles = (a.N + rpt - 1) / rpt;
    float tmin = INFINITY, tmax = -INFINITY;

#define FT_SYNC_ISSUE(...)                 \
    PS_STAMP(0);                           \
    fence_async_smem();                    \
    fence_before();                        \
    group_sync(g);                  \
    PS_STAMP(1);                           \
    if (warp_u == 0) {                     \
        if (elect_one()) {                 \
            fence_after();                 \
            __VA_ARGS__;                   \
            umma_commit(bar);              \
        }                                  \
        __syncwarp();                      \
    }                                      \
    PS_STAMP(2);
#define FT_WAIT()           \
    PS_STAMP(5);            \
    mbar_wait(bar, phase);  \
    phase ^= 1;             \
    fence_after();          \
    PS_STAMP(3);

#ifdef PS_PHASE_CLOCKS
    int nstamp = 0, tile_iter = 0;
#endif
    for (int64_t tile = (int64_t)blockIdx.x * kGroups + g; tile < ntiles; tile += (int64_t)gridDim.x * kGroups) {
#ifdef PS_PHASE_CLOCKS
        const bool stamp_on = g_phase_buf_fwd != nullptr && blockIdx.x == 0 && tid == 0 && (tile_iter == 3 || tile_iter == 4);
        ++tile_iter;
        PS_STAMP(9);
#endif
        const int q = t / S;                                   // ray within the tile
        const int s = t - q * S;
        const int64_t ray = tile * rpt + q;
        const bool valid = t < rows_used && ray < a.N;
        const int64_t p = ray * S + s;
        // ---- stage inputs (all global loads of the row in one batch) -----------------------------------
        float t0, t1, selv;
        {
            RowInputs<K0> in;
            load_row_inputs<K0>(a, P, ray, s, valid, true, true, in);
            stage_features<K0>(in, BufA, t);
            stage_shapp<K0>(in, valid, SHAPPt, t);
            t0 = in.t0; t1 = in.t1; selv = in.selv;
        }
        PS_STAMP(8);
        // ---- base network ------------------------------------------------------------------------------
        FT_SYNC_ISSUE(gemm_bias(tmem, ones, wb + WL::bt(B0), wb + WL::btz, kHid);
                      gemm_kk(tmem, aA, kRows, wb + WL::b0, kHid, kHid, K0, true))
        FT_WAIT()
        hidden_epilogue64(trow, BufB, t);
        FT_SYNC_ISSUE(gemm_bias(tmem, ones, wb + WL::bt(B1), wb + WL::btz, kBaseOut);
                      gemm_kk(tmem, aB, kRows, wb + WL::b1, kBaseOut, kBaseOut, kHid, true))
        FT_WAIT()
        float raw;
        {
            float v[32];
            tmem_ld32_nowait(trow, v);
            tmem_wait_ld();
            raw = v[0];
#pragma unroll
            for (int i = 0; i < 32; i += 8) store_chunk(Ht, kRows, t, i, v + i);
            tmem_ld32_nowait(trow + 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; i += 8) store_chunk(Ht, kRows, t, 32 + i, v + i);
            float u[16];
            tmem_ld16_nowait(trow + 64, u);
            tmem_wait_ld();
            store_chunk(Ht, kRows, t, 64, u);
            store_chunk(Ht, kRows, t, 72, u + 8);
        }
        // ---- both heads, layer by layer in the same phases (independent chains, two accumulators): the colour head
        // [sh | h[0:16] | app] -> columns 0..63, the semantic head h[16:80] (chunks 2..9 of the H tile) -> columns 64..127
        FT_SYNC_ISSUE(gemm_bias(tmem, ones, wb + WL::bt(R0), wb + WL::btz, kHid);
                      gemm_kk(tmem, aSH, kRows, wb + WL::r0, kHid, kHid, 16, true);
                      gemm_kk(tmem, aH, kRows, wb + WL::r0 + 2 * kHid * 16, kHid, kHid, 16, true);
                      gemm_kk(tmem, aSH + 2 * kRows * 16, kRows, wb + WL::r0 + 4 * kHid * 16, kHid, kHid, 16, true);
                      gemm_bias(tmem + 64, ones, wb + WL::bt(S0), wb + WL::btz, kHid);
                      gemm_kk(tmem + 64, aH + 2 * kRows * 16, kRows, wb + WL::s0, kHid, kHid, kSem, true))
        // ---- weights of this ray (overlaps the MMA): rays.py:138-148 -------------------------------------
        const float density = valid ? expf(raw) * selv : 0.f;
        const float dd = __fmul_rn(__fsub_rn(t1, t0), density);
        const double dd_incl = warp_scan_incl((double)dd, lane);
        if (lane == 31) tails[warp * 2] = dd_incl;
        group_sync(g);
        const int w_first = (warp / wpr) * wpr;                // first warp of this warp's ray
        double carry = 0.0;
        for (int k = w_first; k < warp; ++k) carry += tails[k * 2];
        float w, T;
        {
            const double incl = dd_incl + carry;
            const double prev = __shfl_up_sync(0xffffffffu, incl, 1);
            const double excl = lane == 0 ? carry : prev;
            T = expf(-(float)excl);
            const float alpha = __fsub_rn(1.f, expf(-dd));
            w = nan_to_num(__fmul_rn(alpha, T));
            if (!valid) w = 0.f;
        }
        const float tm = __fdiv_rn(__fadd_rn(t0, t1), 2.f);
        if (valid) {
            a.weights[p] = w;
            tmin = fminf(tmin, tm);
            tmax = fmaxf(tmax, tm);
        }
        // threshold depth (renderers.py:352-362): first sample whose inclusive cumsum of weights reaches thr
        const double w_incl = warp_scan_incl((double)w, lane);
        if (lane == 31) tails[warp * 2 + 1] = w_incl;
        const float acc_w = warp_sum(w), dnum_w = warp_sum(w * tm);
        if (lane == 0) { red[warp * 72 + 64] = acc_w; red[warp * 72 + 65] = dnum_w; }
        FT_WAIT()
        hidden_epilogue64(trow, BufA, t);            // colour hidden 1   (X0 is dead: its GEMM completed long ago)
        hidden_epilogue64(trow + 64, BufB, t);       // semantic hidden 1 (H1 likewise)
        FT_SYNC_ISSUE(gemm_bias(tmem, ones, wb + WL::bt(R1), wb + WL::btz, kHid);
                      gemm_kk(tmem, aA, kRows, wb + WL::r1, kHid, kHid, kHid, true);
                      gemm_bias(tmem + 64, ones, wb + WL::bt(S1), wb + WL::btz, kHid);
                      gemm_kk(tmem + 64, aB, kRows, wb + WL::s1, kHid, kHid, kHid, true))
        {   // (after the barrier inside FT_SYNC_ISSUE the weight tails of all warps are visible)
            double wc = 0.0;
            for (int k = w_first; k < warp; ++k) wc += tails[k * 2 + 1];
            const unsigned hit = __ballot_sync(0xffffffffu, valid && (float)(w_incl + wc) >= a.thr);
            if (lane == 0) foundw[warp] = hit ? (warp - w_first) * 32 + __ffs(hit) - 1 : 0x7fffffff;
        }
        FT_WAIT()
        hidden_epilogue64(trow, BufA, t);            // hidden 2 of both heads, in place: the GEMMs that read the tiles
        hidden_epilogue64(trow + 64, BufB, t);       // have completed
        FT_SYNC_ISSUE(gemm_bias(tmem, ones, wb + WL::bt(R2), wb + WL::btz, kRgbOut);
                      gemm_kk(tmem, aA, kRows, wb + WL::r2, kRgbOut, kRgbOut, kHid, true);
                      gemm_bias(tmem + 64, ones, wb + WL::bt(S2), wb + WL::btz, kSem);
                      gemm_kk(tmem + 64, aB, kRows, wb + WL::s2, kSem, kSem, kHid, true))
        FT_WAIT()
        {
            float v[64];
            tmem_ld32_nowait(trow + 64, *reinterpret_cast<float(*)[32]>(v));
            tmem_ld32_nowait(trow + 96, *reinterpret_cast<float(*)[32]>(v + 32));
            float u[16];
            tmem_ld16_nowait(trow, u);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] *= w;
            warp_transpose_reduce<64>(v, lane);
            red[warp * 72 + 2 * lane] = v[0];
            red[warp * 72 + 2 * lane + 1] = v[1];
            float c3[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) c3[i] = warp_sum(w * sigmoid_f(u[i]));
            if (lane == 0) { red[warp * 72 + 66] = c3[0]; red[warp * 72 + 67] = c3[1]; red[warp * 72 + 68] = c3[2]; }
        }
        fence_before();
        group_sync(g);          // red[] / foundw[] complete; every thread is done with the accumulators and tiles
        // per-ray semantics / colour / accumulation / depths
        for (int i = t; i < rpt * 64; i += 128) {
            const int qq = i >> 6, c = i & 63;
            const int64_t rr = tile * rpt + qq;
            if (rr < a.N) {
                float sum = 0.f;
                for (int k = 0; k < wpr; ++k) sum += red[(qq * wpr + k) * 72 + c];
                a.sem_out[rr * kSem + c] = sum;
                if (c < 3) {
                    float col = 0.f;
                    for (int k = 0; k < wpr; ++k) col += red[(qq * wpr + k) * 72 + 66 + c];
                    a.rgb_out[rr * 3 + c] = col;
                }
                if (c == 3) {
                    float accv = 0.f, dn = 0.f;
                    int fnd = 0x7fffffff;
                    for (int k = 0; k < wpr; ++k) {
                        accv += red[(qq * wpr + k) * 72 + 64];
                        dn += red[(qq * wpr + k) * 72 + 65];
                        fnd = min(fnd, foundw[qq * wpr + k]);
                    }
                    fnd = min(fnd, S - 1);
                    a.acc[rr] = accv;
                    a.dexp[rr] = dn / (accv + 1e-10f);
                    const float* b = a.eu + rr * (S + 1);
                    a.dthr[rr] = __fdiv_rn(__fadd_rn(__ldg(b + fnd), __ldg(b + fnd + 1)), 2.f);
                }
            }
        }
        group_sync(g);          // red[] consumed before the next tile overwrites it
        PS_STAMP(7);
    }
#undef FT_SYNC_ISSUE
#undef FT_WAIT
    if (a.tminmax) {
        tmin = warp_min(tmin);
        tmax = warp_max(tmax);
        if (lane == 0 && tmin <= tmax) {
            atomic_min_float(a.tminmax, tmin);
            atomic_max_float(a.tminmax + 1, tmax);
        }
    }
    fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(*tmem_slot, 256);
}

template <int K0>
static int launch_field_fwd(const FieldArgs& a, cudaStream_t stream) {
    constexpr size_t smem = FwdSmem<K0>::total;
    static_assert(smem <= 227 * 1024, "field_fwd: shared memory");
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(field_fwd_kernel<K0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess) {
            set_error("field_level_fwd: cannot reserve %zu bytes of shared memory", smem);
            return 2;
        }
        configured = true;
    }
    const int rpt = kRows / a.S;
    const int64_t ntiles = (a.N + rpt - 1) / rpt;
    const int64_t pairs = (ntiles + kGroups - 1) / kGroups;
    const int grid = (int)(pairs < kNumSMs ? pairs : kNumSMs);
    field_fwd_kernel<K0><<<grid, kFwdThreads, smem, stream>>>(a);
    return check_launch("field_level_fwd");
}


// ------------------------------------------------------------------------------------------------------------------
// Sub-field mode forward (see FieldMsArgs in field_tc5.cuh): the same five GEMM -> epilogue phases per 128-row tile, two
// tiles (one per thread group) in lock step so that both always belong to the same sub-field; CTAs take tile pairs
// round-robin and restage the 54 KB of weights when the sub-field changes.  Outputs per point.
template <int K0>
__global__ void __launch_bounds__(kFwdThreads, 1) field_fwd_ms_kernel(FieldMsArgs a) {
    using WL = WLayout<K0>;
    using SM = FwdSmem<K0>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, g = tid >> 7, t = tid & 127, warp = t >> 5;
    unsigned char* wbase = smem;
    unsigned char* gs = smem + SM::groups + g * SM::group_bytes;
    unsigned char* Ht = gs + SM::h;
    unsigned char* SHAPPt = gs + SM::shapp;
    unsigned char* BufA = gs + SM::bufa;
    unsigned char* BufB = gs + SM::bufb;
    uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(smem + SM::bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::bars + 16);
    if (tid < 32) tmem_alloc(tmem_slot, 256);
    if (tid == 0) {
        mbar_init(smem_u32(bar_ptr), 1);
        mbar_init(smem_u32(bar_ptr + 1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot + g * 128;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t bar = smem_u32(bar_ptr + g);
    const uint32_t wb = smem_u32(wbase), ones = wb + WL::ones;
    const uint32_t aH = smem_u32(Ht), aSH = smem_u32(SHAPPt), aA = smem_u32(BufA), aB = smem_u32(BufB);
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    uint32_t phase = 0;
    const int64_t npairs = a.rows / (2 * kRows);       // taken round-robin: all CTAs stay on (mostly) the same sub-field
    int cur = -1;
    FieldNet net{};

#define FT_SYNC_ISSUE(...)                 \
    fence_async_smem();                    \
    fence_before();                        \
    group_sync(g);                  \
    if (warp_u == 0) {                     \
        if (elect_one()) {                 \
            fence_after();                 \
            __VA_ARGS__;                   \
            umma_commit(bar);              \
        }                                  \
        __syncwarp();                      \
    }
#define FT_WAIT()           \
    mbar_wait(bar, phase);  \
    phase ^= 1;             \
    fence_after();

    for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        const int sf = a.tile_sf[2 * pair];
        if (sf == 255) break;                                   // the used tiles are a prefix
        if (sf != cur) {
            __syncthreads();                                    // both groups are done with the previous weights
            net = a.nets[sf];
            load_all_weights<K0>(net, wbase, tid, kFwdThreads);
            fence_async_smem();
            __syncthreads();
            cur = sf;
        }
        const int64_t i = (2 * pair + g) * kRows + t;
        const int32_t p = a.perm[i];
        const bool valid = p >= 0;
        float selv;
        {
            RowInputs<K0> in;
            load_row_inputs_ms<K0>(a, net, i, p, true, true, in);
            stage_features<K0>(in, BufA, t);
            stage_shapp<K0>(in, valid, SHAPPt, t);
            selv = in.selv;
        }
        FT_SYNC_ISSUE(gemm_bias(tmem, ones, wb + WL::bt(B0), wb + WL::btz, kHid);
                      gemm_kk(tmem, aA, kRows, wb + WL::b0, kHid, kHid, K0, true))
        FT_WAIT()
        hidden_epilogue64(trow, BufB, t);
        FT_SYNC_ISSUE(gemm_bias(tmem, ones, wb + WL::bt(B1), wb + WL::btz, kBaseOut);
                      gemm_kk(tmem, aB, kRows, wb + WL::b1, kBaseOut, kBaseOut, kHid, true))
        FT_WAIT()
        float raw;
        {
            float v[32];
            tmem_ld32_nowait(trow, v);
            tmem_wait_ld();
            raw = v[0];
#pragma unroll
            for (int k = 0; k < 32; k += 8) store_chunk(Ht, kRows, t, k, v + k);
            tmem_ld32_nowait(trow + 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; k += 8) store_chunk(Ht, kRows, t, 32 + k, v + k);
            float u[16];
            tmem_ld16_nowait(trow + 64, u);
            tmem_wait_ld();
            store_chunk(Ht, kRows, t, 64, u);
            store_chunk(Ht, kRows, t, 72, u + 8);
        }
        FT_SYNC_ISSUE(gemm_bias(tmem, ones, wb + WL::bt(R0), wb + WL::btz, kHid);
                      gemm_kk(tmem, aSH, kRows, wb + WL::r0, kHid, kHid, 16, true);
                      gemm_kk(tmem, aH, kRows, wb + WL::r0 + 2 * kHid * 16, kHid, kHid, 16, true);
                      gemm_kk(tmem, aSH + 2 * kRows * 16, kRows, wb + WL::r0 + 4 * kHid * 16, kHid, kHid, 16, true);
                      gemm_bias(tmem + 64, ones, wb + WL::bt(S0), wb + WL::btz, kHid);
                      gemm_kk(tmem + 64, aH + 2 * kRows * 16, kRows, wb + WL::s0, kHid, kHid, kSem, true))
        if (valid) a.density[p] = expf(raw) * selv;               // ingp_field.py:185-190
        FT_WAIT()
        hidden_epilogue64(trow, BufA, t);
        hidden_epilogue64(trow + 64, BufB, t);
        FT_SYNC_ISSUE(gemm_bias(tmem, ones, wb + WL::bt(R1), wb + WL::btz, kHid);
                      gemm_kk(tmem, aA, kRows, wb + WL::r1, kHid, kHid, kHid, true);
                      gemm_bias(tmem + 64, ones, wb + WL::bt(S1), wb + WL::btz, kHid);
                      gemm_kk(tmem + 64, aB, kRows, wb + WL::s1, kHid, kHid, kHid, true))
        FT_WAIT()
        hidden_epilogue64(trow, BufA, t);
        hidden_epilogue64(trow + 64, BufB, t);
        FT_SYNC_ISSUE(gemm_bias(tmem, ones, wb + WL::bt(R2), wb + WL::btz, kRgbOut);
                      gemm_kk(tmem, aA, kRows, wb + WL::r2, kRgbOut, kRgbOut, kHid, true);
                      gemm_bias(tmem + 64, ones, wb + WL::bt(S2), wb + WL::btz, kSem);
                      gemm_kk(tmem + 64, aB, kRows, wb + WL::s2, kSem, kSem, kHid, true))
        FT_WAIT()
        {
            float v[64];
            tmem_ld32_nowait(trow + 64, *reinterpret_cast<float(*)[32]>(v));
            tmem_ld32_nowait(trow + 96, *reinterpret_cast<float(*)[32]>(v + 32));
            float u[16];
            tmem_ld16_nowait(trow, u);
            tmem_wait_ld();
            if (valid) {
                float4* dst = reinterpret_cast<float4*>(a.sem + (int64_t)p * kSem);
#pragma unroll
                for (int k = 0; k < 16; ++k) dst[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
#pragma unroll
                for (int k = 0; k < 3; ++k) a.rgb[(int64_t)p * 3 + k] = sigmoid_f(u[k]);
            }
        }
        fence_before();
        group_sync(g);          // every thread of the group is done with the accumulators and tiles
    }
#undef FT_SYNC_ISSUE
#undef FT_WAIT
    fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(*tmem_slot, 256);
}

template <int K0>
static int launch_field_fwd_ms(const FieldMsArgs& a, cudaStream_t stream) {
    constexpr size_t smem = FwdSmem<K0>::total;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(field_fwd_ms_kernel<K0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess) {
            set_error("field_level_fwd_ms: cannot reserve %zu bytes of shared memory", smem);
            return 2;
        }
        configured = true;
    }
    const int64_t npairs = a.rows / (2 * kRows);
    const int grid = (int)(npairs < kNumSMs ? npairs : kNumSMs);
    field_fwd_ms_kernel<K0><<<grid, kFwdThreads, smem, stream>>>(a);
    return check_launch("field_level_fwd_ms");
}

}  // namespace ftc5
}  // namespace ps

using namespace ps;
using namespace ps::ftc5;

#ifdef PS_PHASE_CLOCKS
/* tools only (debug build) */
extern "C" int ps_debug_phase_buf_fwd(long long* buf) {
    return cudaMemcpyToSymbol(g_phase_buf_fwd, &buf, sizeof(buf)) == cudaSuccess ? 0 : 2;
}
#endif

int ps_field_check_common(const ps_field_net* net, int L, int F, int64_t N, int S, const char* what);

extern "C" int ps_field_level_fwd(const ps_field_net* net, const float* feat_lm, int L, int F, const uint8_t* sel,
                                  const float* eu_bins, const float* dirs, const float* app, int64_t N, int S,
                                  float threshold, float* weights, float* rgb_out, float* acc, float* depth_exp,
                                  float* depth_thr, float* sem_out, float* tminmax, void* stream) {
    if (N == 0) return 0;
    if (int e = ps_field_check_common(net, L, F, N, S, "field_level_fwd")) return e;
    PS_REQUIRE(feat_lm && eu_bins && dirs && weights && rgb_out && acc && depth_exp && depth_thr && sem_out,
               "field_level_fwd: null pointer");
    PS_REQUIRE(net->app_dim == 0 || app != nullptr, "field_level_fwd: appearance is null");
    FieldArgs a{};
    for (int l = 0; l < kLayers; ++l) { a.net.W[l] = net->W[l]; a.net.B[l] = net->B[l]; }
    a.net.in_dim = L * F;
    a.net.app_dim = net->app_dim;
    a.feat = feat_lm; a.L = L; a.F = F; a.sel = sel; a.eu = eu_bins; a.dirs = dirs; a.app = app; a.N = N; a.S = S;
    a.thr = threshold;
    a.weights = weights; a.rgb_out = rgb_out; a.acc = acc; a.dexp = depth_exp; a.dthr = depth_thr; a.sem_out = sem_out;
    a.tminmax = tminmax;
    if (L * F <= 32) return launch_field_fwd<32>(a, (cudaStream_t)stream);
    return launch_field_fwd<48>(a, (cudaStream_t)stream);
}

int ps_field_check_common(const ps_field_net* net, int L, int F, int64_t N, int S, const char* what) {
    PS_REQUIRE(net != nullptr, "%s: net is null", what);
    PS_REQUIRE(F == 2 || F == 4, "%s: features_per_level %d not in {2,4}", what, F);
    PS_REQUIRE(L >= 1 && L * F <= 48, "%s: L*F = %d exceeds 48", what, L * F);
    PS_REQUIRE(S >= 32 && S <= 128 && S % 32 == 0, "%s: samples per ray %d must be 32, 64, 96 or 128", what, S);
    PS_REQUIRE(N > 0 && N * (int64_t)S < (1ll << 31), "%s: too many points", what);
    PS_REQUIRE(net->app_dim >= 0 && net->app_dim <= 16, "%s: appearance dim %d exceeds 16", what, net->app_dim);
    for (int l = 0; l < kLayers; ++l) PS_REQUIRE(net->W[l] != nullptr, "%s: weight %d is null", what, l);
    return 0;
}

int ps_field_ms_check(const ps_field_net_dev* nets_dev, int L, int F, int64_t rows, int S, int app_dim, const char* what) {
    PS_REQUIRE(nets_dev != nullptr, "%s: nets is null", what);
    PS_REQUIRE(F == 2 || F == 4, "%s: features_per_level %d not in {2,4}", what, F);
    PS_REQUIRE(L >= 1 && L * F <= 48, "%s: L*F = %d exceeds 48", what, L * F);
    PS_REQUIRE(rows > 0 && rows % 256 == 0 && rows < (1ll << 31), "%s: rows must be a positive multiple of 256", what);
    PS_REQUIRE(S >= 1, "%s: samples per ray %d < 1", what, S);
    PS_REQUIRE(app_dim >= 0 && app_dim <= 16, "%s: appearance dim %d exceeds 16", what, app_dim);
    return 0;
}

static_assert(sizeof(ps_field_net_dev) == sizeof(ps::ftc5::FieldNet), "ps_field_net_dev layout");

extern "C" int ps_field_level_fwd_ms(const ps_field_net_dev* nets_dev, int app_dim, const float* feat_lm_sorted, int L, int F,
                                     const uint8_t* sel_sorted, const int32_t* perm, const uint8_t* tile_sf, int64_t rows,
                                     int S, const float* dirs, const float* app, float* density, float* rgb, float* sem,
                                     void* stream) {
    if (int e = ps_field_ms_check(nets_dev, L, F, rows, S, app_dim, "field_level_fwd_ms")) return e;
    PS_REQUIRE(feat_lm_sorted && sel_sorted && perm && tile_sf && dirs && density && rgb && sem,
               "field_level_fwd_ms: null pointer");
    PS_REQUIRE(app_dim == 0 || app != nullptr, "field_level_fwd_ms: appearance is null");
    FieldMsArgs a{};
    a.nets = reinterpret_cast<const FieldNet*>(nets_dev);
    a.feat = feat_lm_sorted; a.L = L; a.F = F; a.sels = sel_sorted; a.perm = perm; a.tile_sf = tile_sf; a.rows = rows;
    a.S = S; a.dirs = dirs; a.app = app; a.density = density; a.rgb = rgb; a.sem = sem;
    if (L * F <= 32) return launch_field_fwd_ms<32>(a, (cudaStream_t)stream);
    return launch_field_fwd_ms<48>(a, (cudaStream_t)stream);
}
