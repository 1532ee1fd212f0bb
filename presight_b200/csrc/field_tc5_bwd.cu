// Fused field level, backward: recomputes the forward of field_tc5_fwd.cu on chip and produces the gradient of the
// hash features, of the appearance embedding and of every weight / bias — one kernel.
// Reference for the arithmetic being differentiated: fields/PreSight/ingp_field.py:163-251 (field),
// field_components/activations.py:28-41 (trunc_exp), cameras/rays.py:128-150 (weights),
// model_components/renderers.py:70-117, 286-314, 332-383 and nerfacto_nusc_ms.py:497-530 (compositing).
//
// CTA = 256 threads, one 128-point tile at a time: thread (r = tid % 128, half = tid / 128) owns row r and one half of
// the columns of every epilogue.  Per layer three GEMMs run on tcgen05.mma:
//   forward        Z  = A W^T            (K-major A, K-major B)
//   input gradient dA = dZ W             (K-major A = dZ tile, MN-major B = the forward weight tile)
//   weight gradient dW += dZ^T A          (MN-major A = dZ tile, MN-major B = activation tile; reduction over the
//                                          128 points; accumulators stay in TMEM for the whole kernel — 416..432 of
//                                          the 512 columns — and are flushed once per CTA)
// The activation / gradient tiles are written once (thread-per-row, bank-conflict free) in the chunk-major layout of
// tc5.cuh and serve all three forms without a transpose.  The weight-gradient GEMM of a layer is issued behind its
// input-gradient GEMM and is not waited for: it overlaps the next epilogue; the two gradient tiles alternate so a
// tile is rewritten only after the GEMMs reading it have completed (in-order tensor pipe).
// Biases: the forward adds them as one more K step against a constant operand; their gradients dB_l = dZ_l^T 1 are one
// more reduction over the points on the tensor core, against a one-hot constant operand (field_tc5.cuh:gemm_bias_grad),
// all layers into the 16-column accumulator that also holds the colour head's last weight gradient.  (Round 1 summed
// them in registers with a transposing shuffle butterfly per epilogue because each tcgen05.mma then cost ~100 issue
// cycles; with the elected-lane issue an MMA is a handful of instructions and the butterflies — 30 % of the kernel's
// executed instructions, profiles/r1_z — are gone.)
#include <stdlib.h>

#include "field_tc5.cuh"

// Phase clocks (tools/phase_clocks.py, debug build only: PS_NVCC_DEFS=-DPS_PHASE_CLOCKS): thread 0 of CTA 0 stamps
// clock64() at every barrier / issue / wait of one steady-state tile into a global buffer as (code, cycles) pairs.
#ifdef PS_PHASE_CLOCKS
__device__ long long* g_phase_buf_bwd = nullptr;
#define PS_STAMP(code)                                     \
    do {                                                   \
        if (stamp_on && nstamp < 250) {                    \
            g_phase_buf_bwd[2 * nstamp] = (code);          \
            g_phase_buf_bwd[2 * nstamp + 1] = clock64();   \
            ++nstamp;                                      \
        }                                                  \
    } while (0)
#else
#define PS_STAMP(code)
#endif

namespace ps {
namespace ftc5 {

constexpr int kBwdThreads = 256;

template <int K0>
struct BwdSmem {
    using WL = WLayout<K0>;
    // order matters: tiles used as the M = 128 "X" operand of a weight-gradient GEMM (DZb, DZa, A2) are followed by
    // at least 32 KB of further shared memory (their padding rows read past the tile)
    static constexpr uint32_t dzb = ((WL::end + 127) / 128) * 128;       // [128 x 80]
    static constexpr uint32_t dza = dzb + cm_bytes(kRows, 80);           // [128 x 80]
    static constexpr uint32_t a2 = dza + cm_bytes(kRows, 80);            // [128 x 64]  R2 / S2
    static constexpr uint32_t a1 = a2 + cm_bytes(kRows, 64);             // [128 x 64]  R1 / S1
    static constexpr uint32_t x0 = a1 + cm_bytes(kRows, 64);             // [128 x K0]
    static constexpr uint32_t h1 = x0 + cm_bytes(kRows, K0);             // [128 x 64]
    static constexpr uint32_t h = h1 + cm_bytes(kRows, 64);              // [128 x 80]
    static constexpr uint32_t shapp = h + cm_bytes(kRows, 80);           // [128 x 32]
    static constexpr uint32_t rayc = shapp + cm_bytes(kRows, 32);        // float [4 rays][72]
    static constexpr uint32_t raw = rayc + 4 * 72 * 4;                   // float [128]
    static constexpr uint32_t dots = raw + 128 * 4;                      // float [3][128]: sem half 0, sem half 1, rgb
    static constexpr uint32_t tails = dots + 3 * 128 * 4;                // double [2 halves][4 warps][2]
    static constexpr uint32_t bars = tails + 2 * 4 * 2 * 8;              // mbarrier + tmem slot
    static constexpr uint32_t total = bars + 48;
};

// TMEM columns: working accumulator first, then the weight-gradient accumulators
template <int K0>
struct BwdTmem {
    static constexpr int acc = 0;          // 80
    static constexpr int b0 = 80;          // [64 x K0]
    static constexpr int b1 = b0 + K0;     // [80 x 64]
    static constexpr int s0 = b1 + 64;
    static constexpr int s1 = s0 + 64;
    static constexpr int s2 = s1 + 64;
    static constexpr int r0 = s2 + 64;     // [64 x 48]
    static constexpr int r1 = r0 + 48;
    static constexpr int r2 = r1 + 64;     // transposed: [64 (k) x 16 (n)]
    static constexpr int end = r2 + 16;
    static_assert(end <= 512, "TMEM budget");
};

// 32 accumulator columns (bias included by the GEMM) -> ReLU -> bf16 -> tile columns [c0, c0 + 32)
__device__ __forceinline__ void relu_epilogue32(uint32_t trow, int c0, unsigned char* tile, int r) {
    float v[32];
    tmem_ld32_nowait(trow + c0, v);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; i += 8) store_chunk_relu(tile, kRows, r, c0 + i, v + i);
}

// warp column sums of v[0..32): lane l receives the total of column l
__device__ __forceinline__ float column_sums32(const float (&v)[32], int lane) {
    float t[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) t[i] = v[i];
#pragma unroll
    for (int step = 0, half = 16; step < 5; ++step, half >>= 1) {
        const int bit = 16 >> step;
        const bool upper = (lane & bit) != 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (j < half) {
                const float keep = upper ? t[j + half] : t[j];
                const float send = upper ? t[j] : t[j + half];
                t[j] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
            }
        }
    }
    return t[0];
}

// input-gradient epilogue of a hidden layer: 32 accumulator columns, gated by the sign of the layer's forward
// activation (read back from its tile) -> bf16 dZ
__device__ __forceinline__ void dgrad_epilogue32(uint32_t trow, int c0, const unsigned char* act_tile, unsigned char* dz_tile,
                                                 int r) {
    float v[32];
    tmem_ld32_nowait(trow + c0, v);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; i += 8) store_chunk_relu_grad(dz_tile, act_tile, kRows, r, c0 + i, v + i);
}

template <int K0>
__global__ void __launch_bounds__(kBwdThreads, 1) field_bwd_kernel(FieldArgs a) {
    using WL = WLayout<K0>;
    using SM = BwdSmem<K0>;
    using TM = BwdTmem<K0>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, half = tid >> 7, r = tid & 127, warp = r >> 5, lane = tid & 31;
    unsigned char* wbase = smem;
    unsigned char* DZb = smem + SM::dzb;
    unsigned char* DZa = smem + SM::dza;
    unsigned char* A2 = smem + SM::a2;
    unsigned char* A1 = smem + SM::a1;
    unsigned char* X0 = smem + SM::x0;
    unsigned char* H1 = smem + SM::h1;
    unsigned char* Ht = smem + SM::h;
    unsigned char* SHAPPt = smem + SM::shapp;
    float* rayc = reinterpret_cast<float*>(smem + SM::rayc);
    float* raws = reinterpret_cast<float*>(smem + SM::raw);
    float* dots = reinterpret_cast<float*>(smem + SM::dots);
    double* tails = reinterpret_cast<double*>(smem + SM::tails) + half * 8;
    uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(smem + SM::bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::bars + 24);

    load_all_weights<K0>(a.net, wbase, tid, kBwdThreads);
    // The weight-gradient GEMMs read their dZ operand with M = 128: rows past a layer's width come from whatever lies
    // behind its tile and land in accumulator rows nobody reads.  Tiles start out zeroed all the same.
    for (uint32_t i = SM::dzb + tid * 16; i < SM::rayc; i += kBwdThreads * 16)
        *reinterpret_cast<uint4*>(smem + i) = make_uint4(0u, 0u, 0u, 0u);
    if (tid < 32) tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        mbar_init(smem_u32(bar_ptr), 1);
        mbar_init(smem_u32(bar_ptr + 1), 1);
        mbar_init(smem_u32(bar_ptr + 2), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t bar = smem_u32(bar_ptr), barB0 = smem_u32(bar_ptr + 1), barB1 = smem_u32(bar_ptr + 2);
    const uint32_t wb = smem_u32(wbase), ones = wb + WL::ones, onehot = wb + WL::onehot;
    const uint32_t aDZa = smem_u32(DZa), aDZb = smem_u32(DZb), aA1 = smem_u32(A1), aA2 = smem_u32(A2),
                   aX0 = smem_u32(X0), aH1 = smem_u32(H1), aH = smem_u32(Ht), aSH = smem_u32(SHAPPt);
    constexpr uint32_t CH = kRows * 16;     // bytes per 8-column chunk of a 128-row tile
    uint32_t phase = 0, phaseB0 = 0, phaseB1 = 0;
    // warp index of the CTA as a provably warp-uniform value: the MMA-issuing branches below are taken by whole warps
    // and one elected lane issues (see tc5::elect_one)
    const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);

    const int S = a.S;
    const int rpt = kRows / S, rows_used = rpt * S, wpr = S / 32;
    const int64_t P = a.N * S;
    const int64_t ntiles = (a.N + rpt - 1) / rpt;
    const int A = a.net.app_dim;
    bool first = true;
    // Bias gradients dB_l = dZ_l^T 1 ride on the tensor core behind each layer's weight-gradient GEMM (gemm_bias_grad) and
    // share the 16-column accumulator TM::r2; only the colour head's output layer (3 columns) is summed in registers.
    float db_r2 = 0.f;

#define FB_SYNC_ISSUE(...)     \
    PS_STAMP(0);               \
    fence_async_smem();        \
    fence_before();            \
    __syncthreads();           \
    PS_STAMP(1);               \
    if (warp_u == 0) {         \
        if (elect_one()) {     \
            fence_after();     \
            __VA_ARGS__;       \
        }                      \
        __syncwarp();          \
    }                          \
    PS_STAMP(2);
#define FB_WAIT()           \
    PS_STAMP(5);            \
    mbar_wait(bar, phase);  \
    phase ^= 1;             \
    fence_after();          \
    PS_STAMP(3);
// Two issuing threads.  A single thread pays ~100 cycles per tcgen05.mma it issues, and a tile needs ~220 of them: the
// forward / input-gradient GEMMs (whose results the epilogues wait for) are issued by thread 0, the weight- and
// bias-gradient GEMMs (results needed only at the end of the kernel) by thread 128, each with its own commit barrier.
// The B groups are waited one phase late, just before the tiles they read can be rewritten; they alternate between two
// mbarriers (KB = 0, 1) so that a barrier never completes two phases before every thread has observed the first.
#define FB_SYNC_ISSUE2(A_LIST, B_LIST, KB) \
    PS_STAMP(0);                           \
    fence_async_smem();                    \
    fence_before();                        \
    __syncthreads();                       \
    PS_STAMP(1);                           \
    if (warp_u == 0) {                     \
        if (elect_one()) {                 \
            fence_after();                 \
            A_LIST;                        \
            umma_commit(bar);              \
        }                                  \
        __syncwarp();                      \
    } else if (warp_u == 4) {              \
        if (elect_one()) {                 \
            fence_after();                 \
            B_LIST;                        \
            umma_commit(KB ? barB1 : barB0); \
        }                                  \
        __syncwarp();                      \
    }                                      \
    PS_STAMP(2);
#define FB_WAIT_B(KB)                  \
    if (KB) {                          \
        mbar_wait(barB1, phaseB1);     \
        phaseB1 ^= 1;                  \
    } else {                           \
        mbar_wait(barB0, phaseB0);     \
        phaseB0 ^= 1;                  \
    }                                  \
    PS_STAMP(4);

#ifdef PS_PHASE_CLOCKS
    int nstamp = 0, tile_iter = 0;
#endif
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#ifdef PS_PHASE_CLOCKS
        const bool stamp_on = g_phase_buf_bwd != nullptr && blockIdx.x == 0 && tid == 0 && (tile_iter == 3 || tile_iter == 4);
        ++tile_iter;
        PS_STAMP(9);
#endif
        const bool acc_dw = !first;
        first = false;
        const int q = r / S, s = r - q * S;
        const int64_t ray = tile * rpt + q;
        const bool valid = r < rows_used && ray < a.N;
        const int64_t p = ray * S + s;
        // ---- stage inputs ----------------------------------------------------------------------------------
        float t0, t1, selv, gw_in;
        {
            RowInputs<K0> in;
            load_row_inputs<K0>(a, P, ray, s, valid, half == 0, half == 1, in);
            if (half == 0) stage_features<K0>(in, X0, r);
            else stage_shapp<K0>(in, valid, SHAPPt, r);
            t0 = in.t0; t1 = in.t1; selv = in.selv; gw_in = in.gw;
        }
        for (int i = tid; i < rpt * 72; i += kBwdThreads) {
            const int qq = i / 72, c = i - qq * 72;
            const int64_t rr = tile * rpt + qq;
            float v = 0.f;
            if (rr < a.N) {
                if (c < 64) v = a.d_sem ? __ldg(a.d_sem + rr * kSem + c) : 0.f;
                else if (c < 67) v = a.d_rgb ? __ldg(a.d_rgb + rr * 3 + (c - 64)) : 0.f;
                else if (c == 67) v = a.d_acc ? __ldg(a.d_acc + rr) : 0.f;
                else if (c == 68) v = a.d_dexp ? __ldg(a.d_dexp + rr) : 0.f;
                else if (c == 69) v = a.d_dexp ? __ldg(a.dexp + rr) : 0.f;
                else if (c == 70) v = a.d_dexp ? 1.f / (__ldg(a.acc + rr) + 1e-10f) : 0.f;
            }
            rayc[i] = v;
        }
        const float* rc = rayc + (q < rpt ? q : 0) * 72;
        PS_STAMP(8);
        // ---- base network, forward --------------------------------------------------------------------------
        FB_SYNC_ISSUE(gemm_bias(tmem + TM::acc, ones, wb + WL::bt(B0), kHid, kHid);
                      gemm_kk(tmem + TM::acc, aX0, kRows, wb + WL::b0, kHid, kHid, K0, true); umma_commit(bar))
        FB_WAIT()
        relu_epilogue32(trow + TM::acc, 32 * half, H1, r);
        FB_SYNC_ISSUE(gemm_bias(tmem + TM::acc, ones, wb + WL::bt(B1), kBaseOut, kBaseOut);
                      gemm_kk(tmem + TM::acc, aH1, kRows, wb + WL::b1, kBaseOut, kBaseOut, kHid, true); umma_commit(bar))
        FB_WAIT()
        {
            // half 0: columns 0..47 (raw density, geo, first 32 semantic inputs); half 1: columns 48..79
            float v[32];
            const int c0 = half == 0 ? 0 : 48;
            tmem_ld32_nowait(trow + TM::acc + c0, v);
            tmem_wait_ld();
            if (half == 0) raws[r] = v[0];
#pragma unroll
            for (int i = 0; i < 32; i += 8) store_chunk(Ht, kRows, r, c0 + i, v + i);
            if (half == 0) {
                float u[16];
                tmem_ld16_nowait(trow + TM::acc + 32, u);
                tmem_wait_ld();
                store_chunk(Ht, kRows, r, 32, u);
                store_chunk(Ht, kRows, r, 40, u + 8);
            }
        }
        // ---- colour head, forward ---------------------------------------------------------------------------
        FB_SYNC_ISSUE(gemm_bias(tmem + TM::acc, ones, wb + WL::bt(R0), kHid, kHid);
                      gemm_kk(tmem + TM::acc, aSH, kRows, wb + WL::r0, kHid, kHid, 16, true);
                      gemm_kk(tmem + TM::acc, aH, kRows, wb + WL::r0 + 2 * kHid * 16, kHid, kHid, 16, true);
                      gemm_kk(tmem + TM::acc, aSH + 2 * CH, kRows, wb + WL::r0 + 4 * kHid * 16, kHid, kHid, 16, true);
                      umma_commit(bar))
        // weights of the ray (both halves compute them; rays.py:138-148)
        const float raw = raws[r];
        const float density = valid ? expf(raw) * selv : 0.f;
        const float dl = __fsub_rn(t1, t0);
        const float dd = __fmul_rn(dl, density);
        const double dd_incl = warp_scan_incl((double)dd, lane);
        if (lane == 31) tails[warp * 2] = dd_incl;
        FB_WAIT()
        relu_epilogue32(trow + TM::acc, 32 * half, A1, r);
        FB_SYNC_ISSUE(gemm_bias(tmem + TM::acc, ones, wb + WL::bt(R1), kHid, kHid);
                      gemm_kk(tmem + TM::acc, aA1, kRows, wb + WL::r1, kHid, kHid, kHid, true); umma_commit(bar))
        const int w_first = (warp / wpr) * wpr;
        float w, T;
        bool finite;
        {
            double carry = 0.0;
            for (int k = w_first; k < warp; ++k) carry += tails[k * 2];
            const double incl = dd_incl + carry;
            const double prev = __shfl_up_sync(0xffffffffu, incl, 1);
            const double excl = lane == 0 ? carry : prev;
            T = expf(-(float)excl);
            const float alpha = __fsub_rn(1.f, expf(-dd));
            const float rawp = __fmul_rn(alpha, T);
            w = nan_to_num(rawp);
            finite = isfinite(rawp) && valid;
            if (!valid) w = 0.f;
        }
        const float tm = __fdiv_rn(__fadd_rn(t0, t1), 2.f);
        FB_WAIT()
        relu_epilogue32(trow + TM::acc, 32 * half, A2, r);
        FB_SYNC_ISSUE(gemm_bias(tmem + TM::acc, ones, wb + WL::bt(R2), kRgbOut, kRgbOut);
                      gemm_kk(tmem + TM::acc, aA2, kRows, wb + WL::r2, kRgbOut, kRgbOut, kHid, true); umma_commit(bar))
        FB_WAIT()
        // ---- colour head, backward --------------------------------------------------------------------------
        float dz3[3] = {0.f, 0.f, 0.f};
        if (half == 0) {
            float u[16];
            tmem_ld16_nowait(trow + TM::acc, u);
            tmem_wait_ld();
            float dz[32];
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) dz[i] = 0.f;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float y = sigmoid_f(u[i]);
                dot += y * rc[64 + i];
                dz[i] = w * rc[64 + i] * y * (1.f - y);
            }
            dots[2 * 128 + r] = dot;
            store_chunk(DZa, kRows, r, 0, dz);
            store_chunk(DZa, kRows, r, 8, dz + 8);
            dz3[0] = dz[0]; dz3[1] = dz[1]; dz3[2] = dz[2];
        }
        FB_SYNC_ISSUE2(gemm_dgrad(tmem + TM::acc, aDZa, kRows, wb + WL::r2, kRgbOut, kHid, 16, false),
                       gemm_wgrad(tmem + TM::r2, aA2, aDZa, 16, acc_dw), 0)
        if (half == 0) {
            const float s0 = warp_sum(dz3[0]), s1 = warp_sum(dz3[1]), s2 = warp_sum(dz3[2]);
            db_r2 += lane == 0 ? s0 : (lane == 1 ? s1 : s2);
        }
        FB_WAIT()
        dgrad_epilogue32(trow + TM::acc, 32 * half, A2, DZb, r);
        FB_SYNC_ISSUE2(gemm_dgrad(tmem + TM::acc, aDZb, kRows, wb + WL::r1, kHid, kHid, kHid, false),
                       gemm_wgrad(tmem + TM::r1, aDZb, aA1, kHid, acc_dw);
                       gemm_bias_grad(tmem + TM::r2, aDZb, onehot + R1 * 256), 1)
        FB_WAIT()
        FB_WAIT_B(0)      // the r2 group (read DZa, A2) is complete: DZa may be rewritten
        dgrad_epilogue32(trow + TM::acc, 32 * half, A1, DZa, r);
        FB_SYNC_ISSUE2(gemm_dgrad(tmem + TM::acc, aDZa, kRows, wb + WL::r0, kHid, kRgbIn, kHid, false),
                       gemm_wgrad(tmem + TM::r0, aDZa, aSH, 16, acc_dw);
                       gemm_wgrad(tmem + TM::r0 + 16, aDZa, aH, 16, acc_dw);
                       gemm_wgrad(tmem + TM::r0 + 32, aDZa, aSH + 2 * CH, 16, acc_dw);
                       gemm_bias_grad(tmem + TM::r2, aDZa, onehot + R0 * 256), 0)
        FB_WAIT()
        FB_WAIT_B(1)      // the r1 group (read DZb, A1) is complete
        float d_h01[16];   // half 0: gradient of h[0:16] from the colour head (column 0 is zero by construction)
        {
            float u[16];
            tmem_ld16_nowait(trow + TM::acc + (half == 0 ? 16 : 32), u);
            tmem_wait_ld();
            if (half == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) d_h01[i] = u[i];
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) d_h01[i] = 0.f;
                if (a.dapp && A > 0) {
                    // appearance gradient: sum over the ray's samples (a warp lies inside one ray)
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = i < 16 ? u[i] : 0.f;
                    const float tot = column_sums32(v, lane);
                    const int64_t wray = tile * rpt + (warp * 32) / S;
                    if (lane < A && warp * 32 < rows_used && wray < a.N) atomicAdd(a.dapp + wray * A + lane, tot);
                }
            }
        }
        // ---- semantic head, forward -------------------------------------------------------------------------
        FB_SYNC_ISSUE(gemm_bias(tmem + TM::acc, ones, wb + WL::bt(S0), kHid, kHid);
                      gemm_kk(tmem + TM::acc, aH + 2 * CH, kRows, wb + WL::s0, kHid, kHid, kSem, true); umma_commit(bar))
        FB_WAIT()
        FB_WAIT_B(0)      // the r0 group (read DZa, SHAPP, H) is complete
        relu_epilogue32(trow + TM::acc, 32 * half, A1, r);
        FB_SYNC_ISSUE(gemm_bias(tmem + TM::acc, ones, wb + WL::bt(S1), kHid, kHid);
                      gemm_kk(tmem + TM::acc, aA1, kRows, wb + WL::s1, kHid, kHid, kHid, true); umma_commit(bar))
        FB_WAIT()
        relu_epilogue32(trow + TM::acc, 32 * half, A2, r);
        FB_SYNC_ISSUE(gemm_bias(tmem + TM::acc, ones, wb + WL::bt(S2), kSem, kSem);
                      gemm_kk(tmem + TM::acc, aA2, kRows, wb + WL::s2, kSem, kSem, kHid, true); umma_commit(bar))
        FB_WAIT()
        // ---- semantic head, backward ------------------------------------------------------------------------
        {
            float v[32];
            tmem_ld32_nowait(trow + TM::acc + 32 * half, v);
            tmem_wait_ld();
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float gs = rc[32 * half + i];
                dot += v[i] * gs;
                v[i] = w * gs;
            }
            dots[half * 128 + r] = dot;
#pragma unroll
            for (int i = 0; i < 32; i += 8) store_chunk(DZb, kRows, r, 32 * half + i, v + i);
        }
        FB_SYNC_ISSUE2(gemm_dgrad(tmem + TM::acc, aDZb, kRows, wb + WL::s2, kSem, kHid, kSem, false),
                       gemm_wgrad(tmem + TM::s2, aDZb, aA2, kHid, acc_dw);
                       gemm_bias_grad(tmem + TM::r2, aDZb, onehot + S2 * 256), 1)
        // compositing backward, pass 1 (the barrier above published dots[]): total gradient on this weight
        float g = gw_in + rc[67] + rc[68] * (tm - rc[69]) * rc[70] + dots[r] + dots[128 + r] + dots[256 + r];
        if (!finite) g = 0.f;
        const double gw_incl = warp_scan_incl((double)g * (double)w, lane);
        if (lane == 31) tails[warp * 2 + 1] = gw_incl;
        FB_WAIT()
        dgrad_epilogue32(trow + TM::acc, 32 * half, A2, DZa, r);
        FB_SYNC_ISSUE2(gemm_dgrad(tmem + TM::acc, aDZa, kRows, wb + WL::s1, kHid, kHid, kHid, false),
                       gemm_wgrad(tmem + TM::s1, aDZa, aA1, kHid, acc_dw);
                       gemm_bias_grad(tmem + TM::r2, aDZa, onehot + S1 * 256), 0)
        // pass 2: d sigma_i = delta_i * (g_i T_{i+1} - sum_{k>i} g_k w_k); d raw = d sigma * sel * exp(clamp(raw))
        float d_raw;
        {
            double pc = 0.0, G = 0.0;
            for (int k = w_first; k < w_first + wpr; ++k) {
                if (k < warp) pc += tails[k * 2 + 1];
                G += tails[k * 2 + 1];
            }
            const double Pi = gw_incl + pc;
            const float d_sigma = dl * (float)((double)g * (double)(T * expf(-dd)) - (G - Pi));
            d_raw = valid ? d_sigma * selv * expf(fminf(fmaxf(raw, -15.f), 15.f)) : 0.f;
        }
        FB_WAIT()
        FB_WAIT_B(1)      // the s2 group (read DZb, A2) is complete: DZb may be rewritten
        dgrad_epilogue32(trow + TM::acc, 32 * half, A1, DZb, r);
        FB_SYNC_ISSUE2(gemm_dgrad(tmem + TM::acc, aDZb, kRows, wb + WL::s0, kHid, kSem, kHid, false),
                       gemm_wgrad(tmem + TM::s0, aDZb, aH + 2 * CH, kSem, acc_dw);
                       gemm_bias_grad(tmem + TM::r2, aDZb, onehot + S0 * 256), 1)
        FB_WAIT()
        FB_WAIT_B(0)      // the s1 group (read DZa, A1) is complete: DZa may take dH
        // ---- base network, backward: dH = [d raw | colour head (15) | semantic head (64)] ----------------------
        {
            float v[32];
            tmem_ld32_nowait(trow + TM::acc + 32 * half, v);      // gradient of h[16 + 32*half ...]
            tmem_wait_ld();
            if (!valid) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 32; i += 8) store_chunk(DZa, kRows, r, 16 + 32 * half + i, v + i);
            if (half == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) d_h01[i] = valid ? d_h01[i] : 0.f;
                d_h01[0] = d_raw;
                store_chunk(DZa, kRows, r, 0, d_h01);
                store_chunk(DZa, kRows, r, 8, d_h01 + 8);
            }
        }
        FB_SYNC_ISSUE2(gemm_dgrad(tmem + TM::acc, aDZa, kRows, wb + WL::b1, kBaseOut, kHid, kBaseOut, false),
                       gemm_wgrad(tmem + TM::b1, aDZa, aH1, kHid, acc_dw);
                       gemm_bias_grad(tmem + TM::r2, aDZa, onehot + B1 * 256), 0)
        FB_WAIT()
        FB_WAIT_B(1)      // the s0 group (read DZb, H) is complete
        dgrad_epilogue32(trow + TM::acc, 32 * half, H1, DZb, r);
        FB_SYNC_ISSUE2(gemm_dgrad(tmem + TM::acc, aDZb, kRows, wb + WL::b0, kHid, K0, kHid, false),
                       gemm_wgrad(tmem + TM::b0, aDZb, aX0, K0, acc_dw);
                       gemm_bias_grad(tmem + TM::r2, aDZb, onehot + B0 * 256), 1)
        FB_WAIT()
        FB_WAIT_B(0)      // the b1 group (read DZa, H1) ...
        FB_WAIT_B(1)      // ... and the b0 group (read DZb, X0) are complete: every tile may be restaged
        // ---- hash-feature gradient (level-major [L][P][F]) ---------------------------------------------------
        if (a.dfeat) {
#pragma unroll
            for (int blk = 0; blk < K0 / 16; ++blk) {
                if ((blk & 1) != half) continue;
                float u[16];
                tmem_ld16_nowait(trow + TM::acc + 16 * blk, u);
                tmem_wait_ld();
                if (valid) {
                    if (a.F == 2) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int l = 8 * blk + i;
                            if (l < a.L)
                                *reinterpret_cast<float2*>(a.dfeat + ((int64_t)l * P + p) * 2) = make_float2(u[2 * i], u[2 * i + 1]);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int l = 4 * blk + i;
                            if (l < a.L)
                                *reinterpret_cast<float4*>(a.dfeat + ((int64_t)l * P + p) * 4) =
                                    make_float4(u[4 * i], u[4 * i + 1], u[4 * i + 2], u[4 * i + 3]);
                        }
                    }
                }
            }
        }
        // the staging of the next tile is ordered behind these TMEM reads by the next FB_SYNC_ISSUE barrier; the tiles
        // it overwrites (X0, SHAPP, rayc) were last read by GEMMs / epilogues that completed before the wait above
        PS_STAMP(7);
        __syncthreads();
    }
#undef FB_SYNC_ISSUE
#undef FB_SYNC_ISSUE2
#undef FB_WAIT
#undef FB_WAIT_B

    // ---- flush: weight and bias gradients (TMEM) -> global atomics ---------------------------------------------------
    if (!first) {
        fence_after();
        // region: TMEM column offset, rows (out features), 16-column blocks, global pointer, real row length, column map
        auto flush = [&](int col0, int n_real, int ncols, float* dW, int k_real, int kind) {
            const int n = warp * 32 + lane;
            if (warp * 32 >= n_real) return;           // warp-uniform
            for (int blk = 0; blk < ncols / 16; ++blk) {
                if ((blk & 1) != half) continue;
                float u[16];
                tmem_ld16_nowait(trow + col0 + 16 * blk, u);
                tmem_wait_ld();
                if (n < n_real) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        int k = 16 * blk + i;
                        if (kind == 1) {               // colour head layer 0: staged column -> reference column
                            if (k < 16) {}
                            else if (k == 16) k = -1;
                            else if (k < 32) k -= 1;
                            else k = (k - 32 < A) ? k - 1 : -1;
                        }
                        if (kind == 2) {               // transposed accumulator [k][n]: dW[n_out][k]
                            if (k < 3) atomicAdd(dW + (size_t)k * k_real + n, u[i]);
                            continue;
                        }
                        if (k >= 0 && k < k_real && u[i] != 0.f) atomicAdd(dW + (size_t)n * k_real + k, u[i]);
                    }
                }
            }
        };
        flush(TM::b0, kHid, K0, a.net.dW[B0], a.net.in_dim, 0);
        flush(TM::b1, kBaseOut, kHid, a.net.dW[B1], kHid, 0);
        flush(TM::s0, kHid, kSem, a.net.dW[S0], kSem, 0);
        flush(TM::s1, kHid, kHid, a.net.dW[S1], kHid, 0);
        flush(TM::s2, kSem, kHid, a.net.dW[S2], kHid, 0);
        flush(TM::r0, kHid, kRgbIn, a.net.dW[R0], 16 + kGeo + A, 1);
        flush(TM::r1, kHid, kHid, a.net.dW[R1], kHid, 0);
        flush(TM::r2, kHid, kRgbOut, a.net.dW[R2], kHid, 2);
        // bias gradients: column kBiasCol0 + l of the shared 16-column accumulator, row = out feature
        if (half == 0) {
            float u[16];
            tmem_ld16_nowait(trow + TM::r2, u);
            tmem_wait_ld();
            const int n = warp * 32 + lane;
#pragma unroll
            for (int l = 0; l < kLayers; ++l) {
                if (l == R2) continue;
                if (n < WL::rows(l) && a.net.dB[l]) atomicAdd(a.net.dB[l] + n, u[kBiasCol0 + l]);
            }
            if (lane < 3 && a.net.dB[R2]) atomicAdd(a.net.dB[R2] + lane, db_r2);
        }
    }
    fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 512);
}

template <int K0>
static int launch_field_bwd(const FieldArgs& a, cudaStream_t stream) {
    constexpr size_t smem = BwdSmem<K0>::total;
    static_assert(smem <= 227 * 1024, "field_bwd: shared memory");
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(field_bwd_kernel<K0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess) {
            set_error("field_level_bwd: cannot reserve %zu bytes of shared memory", smem);
            return 2;
        }
        configured = true;
    }
    const int rpt = kRows / a.S;
    const int64_t ntiles = (a.N + rpt - 1) / rpt;
    const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
    field_bwd_kernel<K0><<<grid, kBwdThreads, smem, stream>>>(a);
    return check_launch("field_level_bwd");
}


// ------------------------------------------------------------------------------------------------------------------
// Sub-field mode backward (see FieldMsArgs in field_tc5.cuh): the same recompute + dgrad + wgrad chain per 128-row tile,
// driven by per-point gradients of density / rgb / semantics (the compositing backward runs in ps_composite_bwd).  CTAs take
// tiles round-robin; when the sub-field changes it flushes the TMEM-resident weight / bias gradient
// accumulators into the finished sub-field's buffers and restages the next sub-field's weights.
template <int K0>
__global__ void __launch_bounds__(kBwdThreads, 1) field_bwd_ms_kernel(FieldMsArgs a) {
    using WL = WLayout<K0>;
    using SM = BwdSmem<K0>;
    using TM = BwdTmem<K0>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, half = tid >> 7, r = tid & 127, warp = r >> 5, lane = tid & 31;
    unsigned char* wbase = smem;
    unsigned char* DZb = smem + SM::dzb;
    unsigned char* DZa = smem + SM::dza;
    unsigned char* A2 = smem + SM::a2;
    unsigned char* A1 = smem + SM::a1;
    unsigned char* X0 = smem + SM::x0;
    unsigned char* H1 = smem + SM::h1;
    unsigned char* Ht = smem + SM::h;
    unsigned char* SHAPPt = smem + SM::shapp;
    float* raws = reinterpret_cast<float*>(smem + SM::raw);
    uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(smem + SM::bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::bars + 24);

    for (uint32_t k = SM::dzb + tid * 16; k < SM::rayc; k += kBwdThreads * 16)
        *reinterpret_cast<uint4*>(smem + k) = make_uint4(0u, 0u, 0u, 0u);
    if (tid < 32) tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        mbar_init(smem_u32(bar_ptr), 1);
        mbar_init(smem_u32(bar_ptr + 1), 1);
        mbar_init(smem_u32(bar_ptr + 2), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t bar = smem_u32(bar_ptr), barB0 = smem_u32(bar_ptr + 1), barB1 = smem_u32(bar_ptr + 2);
    const uint32_t wb = smem_u32(wbase), ones = wb + WL::ones, onehot = wb + WL::onehot;
    const uint32_t aDZa = smem_u32(DZa), aDZb = smem_u32(DZb), aA1 = smem_u32(A1), aA2 = smem_u32(A2),
                   aX0 = smem_u32(X0), aH1 = smem_u32(H1), aH = smem_u32(Ht), aSH = smem_u32(SHAPPt);
    constexpr uint32_t CH = kRows * 16;
    uint32_t phase = 0, phaseB0 = 0, phaseB1 = 0;
    const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
    // Tiles are taken round-robin, so at any moment all CTAs work on neighbouring tiles, i.e. (mostly) on the SAME sub-field:
    // the live hash-table working set is one sub-field's tables, which fit the 126 MB L2, instead of all of them.
    const int64_t ntiles = a.rows / kRows;
    bool first = true;
    float db_r2 = 0.f;
    int cur = -1;
    FieldNet net{};

    // weight and bias gradients of the sub-field just finished: TMEM accumulators -> its gradient buffers (atomics)
    auto flush_all = [&]() {
        const int A = net.app_dim;
        if (!first) {
            fence_after();
            auto flush = [&](int col0, int n_real, int ncols, float* dW, int k_real, int kind) {
                const int n = warp * 32 + lane;
                if (warp * 32 >= n_real) return;
                for (int blk = 0; blk < ncols / 16; ++blk) {
                    if ((blk & 1) != half) continue;
                    float u[16];
                    tmem_ld16_nowait(trow + col0 + 16 * blk, u);
                    tmem_wait_ld();
                    if (n < n_real) {
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            int k = 16 * blk + q;
                            if (kind == 1) {
                                if (k < 16) {}
                                else if (k == 16) k = -1;
                                else if (k < 32) k -= 1;
                                else k = (k - 32 < A) ? k - 1 : -1;
                            }
                            if (kind == 2) {
                                if (k < 3) atomicAdd(dW + (size_t)k * k_real + n, u[q]);
                                continue;
                            }
                            if (k >= 0 && k < k_real && u[q] != 0.f) atomicAdd(dW + (size_t)n * k_real + k, u[q]);
                        }
                    }
                }
            };
            flush(TM::b0, kHid, K0, net.dW[B0], net.in_dim, 0);
            flush(TM::b1, kBaseOut, kHid, net.dW[B1], kHid, 0);
            flush(TM::s0, kHid, kSem, net.dW[S0], kSem, 0);
            flush(TM::s1, kHid, kHid, net.dW[S1], kHid, 0);
            flush(TM::s2, kSem, kHid, net.dW[S2], kHid, 0);
            flush(TM::r0, kHid, kRgbIn, net.dW[R0], 16 + kGeo + A, 1);
            flush(TM::r1, kHid, kHid, net.dW[R1], kHid, 0);
            flush(TM::r2, kHid, kRgbOut, net.dW[R2], kHid, 2);
            if (half == 0) {
                float u[16];
                tmem_ld16_nowait(trow + TM::r2, u);
                tmem_wait_ld();
                const int n = warp * 32 + lane;
#pragma unroll
                for (int l = 0; l < kLayers; ++l) {
                    if (l == R2) continue;
                    if (n < WL::rows(l) && net.dB[l]) atomicAdd(net.dB[l] + n, u[kBiasCol0 + l]);
                }
                if (lane < 3 && net.dB[R2]) atomicAdd(net.dB[R2] + lane, db_r2);
            }
            fence_before();
        }
        db_r2 = 0.f;
        first = true;
    };

#define FB_SYNC_ISSUE(...)     \
    fence_async_smem();        \
    fence_before();            \
    __syncthreads();           \
    if (warp_u == 0) {         \
        if (elect_one()) {     \
            fence_after();     \
            __VA_ARGS__;       \
        }                      \
        __syncwarp();          \
    }
#define FB_WAIT()           \
    mbar_wait(bar, phase);  \
    phase ^= 1;             \
    fence_after();
#define FB_SYNC_ISSUE2(A_LIST, B_LIST, KB) \
    fence_async_smem();                    \
    fence_before();                        \
    __syncthreads();                       \
    if (warp_u == 0) {                     \
        if (elect_one()) {                 \
            fence_after();                 \
            A_LIST;                        \
            umma_commit(bar);              \
        }                                  \
        __syncwarp();                      \
    } else if (warp_u == 4) {              \
        if (elect_one()) {                 \
            fence_after();                 \
            B_LIST;                        \
            umma_commit(KB ? barB1 : barB0); \
        }                                  \
        __syncwarp();                      \
    }
#define FB_WAIT_B(KB)                  \
    if (KB) {                          \
        mbar_wait(barB1, phaseB1);     \
        phaseB1 ^= 1;                  \
    } else {                           \
        mbar_wait(barB0, phaseB0);     \
        phaseB0 ^= 1;                  \
    }

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int sf = a.tile_sf[tile];
        if (sf == 255) break;
        if (sf != cur) {
            if (cur >= 0) flush_all();
            __syncthreads();
            net = a.nets[sf];
            load_all_weights<K0>(net, wbase, tid, kBwdThreads);
            fence_async_smem();
            __syncthreads();
            cur = sf;
        }
        const int A = net.app_dim;
        const bool acc_dw = !first;
        first = false;
        const int64_t i = tile * kRows + r;
        const int32_t p = a.perm[i];
        const bool valid = p >= 0;
        const int64_t ray = valid ? p / a.S : 0;
        float selv;
        {
            RowInputs<K0> in;
            load_row_inputs_ms<K0>(a, net, i, p, half == 0, half == 1, in);
            if (half == 0) stage_features<K0>(in, X0, r);
            else stage_shapp<K0>(in, valid, SHAPPt, r);
            selv = in.selv;
        }
        // per-point upstream gradients (this thread's share): density, rgb (half 0), 32 semantic channels
        const float g_den = valid ? __ldg(a.d_density + p) : 0.f;
        float g_rgb[3] = {0.f, 0.f, 0.f};
        if (valid && half == 0) {
            g_rgb[0] = __ldg(a.d_rgb + (int64_t)p * 3);
            g_rgb[1] = __ldg(a.d_rgb + (int64_t)p * 3 + 1);
            g_rgb[2] = __ldg(a.d_rgb + (int64_t)p * 3 + 2);
        }
        // ---- base network, forward --------------------------------------------------------------------------
        FB_SYNC_ISSUE(gemm_bias(tmem + TM::acc, ones, wb + WL::bt(B0), kHid, kHid);
                      gemm_kk(tmem + TM::acc, aX0, kRows, wb + WL::b0, kHid, kHid, K0, true); umma_commit(bar))
        FB_WAIT()
        relu_epilogue32(trow + TM::acc, 32 * half, H1, r);
        FB_SYNC_ISSUE(gemm_bias(tmem + TM::acc, ones, wb + WL::bt(B1), kBaseOut, kBaseOut);
                      gemm_kk(tmem + TM::acc, aH1, kRows, wb + WL::b1, kBaseOut, kBaseOut, kHid, true); umma_commit(bar))
        FB_WAIT()
        {
            float v[32];
            const int c0 = half == 0 ? 0 : 48;
            tmem_ld32_nowait(trow + TM::acc + c0, v);
            tmem_wait_ld();
            if (half == 0) raws[r] = v[0];
#pragma unroll
            for (int k = 0; k < 32; k += 8) store_chunk(Ht, kRows, r, c0 + k, v + k);
            if (half == 0) {
                float u[16];
                tmem_ld16_nowait(trow + TM::acc + 32, u);
                tmem_wait_ld();
                store_chunk(Ht, kRows, r, 32, u);
                store_chunk(Ht, kRows, r, 40, u + 8);
            }
        }
        // ---- colour head, forward ---------------------------------------------------------------------------
        FB_SYNC_ISSUE(gemm_bias(tmem + TM::acc, ones, wb + WL::bt(R0), kHid, kHid);
                      gemm_kk(tmem + TM::acc, aSH, kRows, wb + WL::r0, kHid, kHid, 16, true);
                      gemm_kk(tmem + TM::acc, aH, kRows, wb + WL::r0 + 2 * kHid * 16, kHid, kHid, 16, true);
                      gemm_kk(tmem + TM::acc, aSH + 2 * CH, kRows, wb + WL::r0 + 4 * kHid * 16, kHid, kHid, 16, true);
                      umma_commit(bar))
        const float raw = raws[r];
        // density = exp(raw) * sel; gradient through the clamped exponential (activations.py:28-41)
        const float d_raw = valid ? g_den * selv * expf(fminf(fmaxf(raw, -15.f), 15.f)) : 0.f;
        FB_WAIT()
        relu_epilogue32(trow + TM::acc, 32 * half, A1, r);
        FB_SYNC_ISSUE(gemm_bias(tmem + TM::acc, ones, wb + WL::bt(R1), kHid, kHid);
                      gemm_kk(tmem + TM::acc, aA1, kRows, wb + WL::r1, kHid, kHid, kHid, true); umma_commit(bar))
        FB_WAIT()
        relu_epilogue32(trow + TM::acc, 32 * half, A2, r);
        FB_SYNC_ISSUE(gemm_bias(tmem + TM::acc, ones, wb + WL::bt(R2), kRgbOut, kRgbOut);
                      gemm_kk(tmem + TM::acc, aA2, kRows, wb + WL::r2, kRgbOut, kRgbOut, kHid, true); umma_commit(bar))
        FB_WAIT()
        // ---- colour head, backward --------------------------------------------------------------------------
        float dz3[3] = {0.f, 0.f, 0.f};
        if (half == 0) {
            float u[16];
            tmem_ld16_nowait(trow + TM::acc, u);
            tmem_wait_ld();
            float dz[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) dz[k] = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float y = sigmoid_f(u[k]);
                dz[k] = g_rgb[k] * y * (1.f - y);
            }
            store_chunk(DZa, kRows, r, 0, dz);
            store_chunk(DZa, kRows, r, 8, dz + 8);
            dz3[0] = dz[0]; dz3[1] = dz[1]; dz3[2] = dz[2];
        }
        FB_SYNC_ISSUE2(gemm_dgrad(tmem + TM::acc, aDZa, kRows, wb + WL::r2, kRgbOut, kHid, 16, false),
                       gemm_wgrad(tmem + TM::r2, aA2, aDZa, 16, acc_dw), 0)
        if (half == 0) {
            const float s0 = warp_sum(dz3[0]), s1 = warp_sum(dz3[1]), s2 = warp_sum(dz3[2]);
            db_r2 += lane == 0 ? s0 : (lane == 1 ? s1 : s2);
        }
        FB_WAIT()
        dgrad_epilogue32(trow + TM::acc, 32 * half, A2, DZb, r);
        FB_SYNC_ISSUE2(gemm_dgrad(tmem + TM::acc, aDZb, kRows, wb + WL::r1, kHid, kHid, kHid, false),
                       gemm_wgrad(tmem + TM::r1, aDZb, aA1, kHid, acc_dw);
                       gemm_bias_grad(tmem + TM::r2, aDZb, onehot + R1 * 256), 1)
        FB_WAIT()
        FB_WAIT_B(0)
        dgrad_epilogue32(trow + TM::acc, 32 * half, A1, DZa, r);
        FB_SYNC_ISSUE2(gemm_dgrad(tmem + TM::acc, aDZa, kRows, wb + WL::r0, kHid, kRgbIn, kHid, false),
                       gemm_wgrad(tmem + TM::r0, aDZa, aSH, 16, acc_dw);
                       gemm_wgrad(tmem + TM::r0 + 16, aDZa, aH, 16, acc_dw);
                       gemm_wgrad(tmem + TM::r0 + 32, aDZa, aSH + 2 * CH, 16, acc_dw);
                       gemm_bias_grad(tmem + TM::r2, aDZa, onehot + R0 * 256), 0)
        FB_WAIT()
        FB_WAIT_B(1)
        float d_h01[16];
        {
            float u[16];
            tmem_ld16_nowait(trow + TM::acc + (half == 0 ? 16 : 32), u);
            tmem_wait_ld();
            if (half == 0) {
#pragma unroll
                for (int k = 0; k < 16; ++k) d_h01[k] = u[k];
            } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) d_h01[k] = 0.f;
                // appearance gradient: rows of a tile belong to different rays here -> one atomic per (row, channel)
                if (a.dapp && A > 0 && valid)
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if (k < A) atomicAdd(a.dapp + ray * A + k, u[k]);
            }
        }
        // ---- semantic head, forward -------------------------------------------------------------------------
        FB_SYNC_ISSUE(gemm_bias(tmem + TM::acc, ones, wb + WL::bt(S0), kHid, kHid);
                      gemm_kk(tmem + TM::acc, aH + 2 * CH, kRows, wb + WL::s0, kHid, kHid, kSem, true); umma_commit(bar))
        FB_WAIT()
        FB_WAIT_B(0)
        relu_epilogue32(trow + TM::acc, 32 * half, A1, r);
        FB_SYNC_ISSUE(gemm_bias(tmem + TM::acc, ones, wb + WL::bt(S1), kHid, kHid);
                      gemm_kk(tmem + TM::acc, aA1, kRows, wb + WL::s1, kHid, kHid, kHid, true); umma_commit(bar))
        FB_WAIT()
        relu_epilogue32(trow + TM::acc, 32 * half, A2, r);
        // ---- semantic head, backward: dZ of the (linear) output layer = the per-point gradient itself ----------------
        {
            float v[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = 0.f;
            if (valid) {
                const float4* src = reinterpret_cast<const float4*>(a.d_sem + (int64_t)p * kSem + 32 * half);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 q = __ldg(src + k);
                    v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
                }
            }
#pragma unroll
            for (int k = 0; k < 32; k += 8) store_chunk(DZb, kRows, r, 32 * half + k, v + k);
        }
        FB_SYNC_ISSUE2(gemm_dgrad(tmem + TM::acc, aDZb, kRows, wb + WL::s2, kSem, kHid, kSem, false),
                       gemm_wgrad(tmem + TM::s2, aDZb, aA2, kHid, acc_dw);
                       gemm_bias_grad(tmem + TM::r2, aDZb, onehot + S2 * 256), 1)
        FB_WAIT()
        dgrad_epilogue32(trow + TM::acc, 32 * half, A2, DZa, r);
        FB_SYNC_ISSUE2(gemm_dgrad(tmem + TM::acc, aDZa, kRows, wb + WL::s1, kHid, kHid, kHid, false),
                       gemm_wgrad(tmem + TM::s1, aDZa, aA1, kHid, acc_dw);
                       gemm_bias_grad(tmem + TM::r2, aDZa, onehot + S1 * 256), 0)
        FB_WAIT()
        FB_WAIT_B(1)
        dgrad_epilogue32(trow + TM::acc, 32 * half, A1, DZb, r);
        FB_SYNC_ISSUE2(gemm_dgrad(tmem + TM::acc, aDZb, kRows, wb + WL::s0, kHid, kSem, kHid, false),
                       gemm_wgrad(tmem + TM::s0, aDZb, aH + 2 * CH, kSem, acc_dw);
                       gemm_bias_grad(tmem + TM::r2, aDZb, onehot + S0 * 256), 1)
        FB_WAIT()
        FB_WAIT_B(0)
        // ---- base network, backward ---------------------------------------------------------------------------
        {
            float v[32];
            tmem_ld32_nowait(trow + TM::acc + 32 * half, v);
            tmem_wait_ld();
            if (!valid) {
#pragma unroll
                for (int k = 0; k < 32; ++k) v[k] = 0.f;
            }
#pragma unroll
            for (int k = 0; k < 32; k += 8) store_chunk(DZa, kRows, r, 16 + 32 * half + k, v + k);
            if (half == 0) {
#pragma unroll
                for (int k = 0; k < 16; ++k) d_h01[k] = valid ? d_h01[k] : 0.f;
                d_h01[0] = d_raw;
                store_chunk(DZa, kRows, r, 0, d_h01);
                store_chunk(DZa, kRows, r, 8, d_h01 + 8);
            }
        }
        FB_SYNC_ISSUE2(gemm_dgrad(tmem + TM::acc, aDZa, kRows, wb + WL::b1, kBaseOut, kHid, kBaseOut, false),
                       gemm_wgrad(tmem + TM::b1, aDZa, aH1, kHid, acc_dw);
                       gemm_bias_grad(tmem + TM::r2, aDZa, onehot + B1 * 256), 0)
        FB_WAIT()
        FB_WAIT_B(1)
        dgrad_epilogue32(trow + TM::acc, 32 * half, H1, DZb, r);
        FB_SYNC_ISSUE2(gemm_dgrad(tmem + TM::acc, aDZb, kRows, wb + WL::b0, kHid, K0, kHid, false),
                       gemm_wgrad(tmem + TM::b0, aDZb, aX0, K0, acc_dw);
                       gemm_bias_grad(tmem + TM::r2, aDZb, onehot + B0 * 256), 1)
        FB_WAIT()
        FB_WAIT_B(0)
        FB_WAIT_B(1)
        // ---- hash-feature gradient (level-major [L][rows][F], row order) ---------------------------------------
        if (a.dfeat) {
#pragma unroll
            for (int blk = 0; blk < K0 / 16; ++blk) {
                if ((blk & 1) != half) continue;
                float u[16];
                tmem_ld16_nowait(trow + TM::acc + 16 * blk, u);
                tmem_wait_ld();
                if (a.F == 2) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int l = 8 * blk + k;
                        if (l < a.L)
                            *reinterpret_cast<float2*>(a.dfeat + ((int64_t)l * a.rows + i) * 2) =
                                valid ? make_float2(u[2 * k], u[2 * k + 1]) : make_float2(0.f, 0.f);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int l = 4 * blk + k;
                        if (l < a.L)
                            *reinterpret_cast<float4*>(a.dfeat + ((int64_t)l * a.rows + i) * 4) =
                                valid ? make_float4(u[4 * k], u[4 * k + 1], u[4 * k + 2], u[4 * k + 3])
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
        }
        __syncthreads();
    }
#undef FB_SYNC_ISSUE
#undef FB_SYNC_ISSUE2
#undef FB_WAIT
#undef FB_WAIT_B
    if (cur >= 0) flush_all();
    fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 512);
}

template <int K0>
static int launch_field_bwd_ms(const FieldMsArgs& a, cudaStream_t stream) {
    constexpr size_t smem = BwdSmem<K0>::total;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(field_bwd_ms_kernel<K0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess) {
            set_error("field_level_bwd_ms: cannot reserve %zu bytes of shared memory", smem);
            return 2;
        }
        configured = true;
    }
    const int64_t ntiles = a.rows / kRows;
    const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
    field_bwd_ms_kernel<K0><<<grid, kBwdThreads, smem, stream>>>(a);
    return check_launch("field_level_bwd_ms");
}

}  // namespace ftc5
}  // namespace ps

using namespace ps;
using namespace ps::ftc5;

#ifdef PS_PHASE_CLOCKS
/* tools only (debug build) */
extern "C" int ps_debug_phase_buf_bwd(long long* buf) {
    return cudaMemcpyToSymbol(g_phase_buf_bwd, &buf, sizeof(buf)) == cudaSuccess ? 0 : 2;
}
#endif

int ps_field_check_common(const ps_field_net* net, int L, int F, int64_t N, int S, const char* what);

namespace ps {
namespace ftc5 {
int launch_field_bwd2(const FieldArgs& a, cudaStream_t stream);          // field_tc5_bwd2.cu
int launch_field_bwd2_ms(const FieldMsArgs& a, cudaStream_t stream);
}
}
// PS_FIELD_BWD_V1=1 selects the first (single-chain) backward kernels for A/B timing
static bool use_v1() {
    static const bool v = [] { const char* e = getenv("PS_FIELD_BWD_V1"); return e && e[0] == '1'; }();
    return v;
}

extern "C" int ps_field_level_bwd(const ps_field_net* net, const float* feat_lm, int L, int F, const uint8_t* sel,
                                  const float* eu_bins, const float* dirs, const float* app, int64_t N, int S,
                                  const float* acc, const float* depth_exp, const float* d_weights,
                                  const float* d_rgb_out, const float* d_acc, const float* d_depth_exp,
                                  const float* d_sem_out, float* dfeat_lm, float* dapp, void* stream) {
    if (N == 0) return 0;
    if (int e = ps_field_check_common(net, L, F, N, S, "field_level_bwd")) return e;
    PS_REQUIRE(feat_lm && eu_bins && dirs, "field_level_bwd: null pointer");
    PS_REQUIRE(net->app_dim == 0 || app != nullptr, "field_level_bwd: appearance is null");
    PS_REQUIRE(d_depth_exp == nullptr || (acc && depth_exp), "field_level_bwd: d_depth_exp needs acc and depth_exp");
    for (int l = 0; l < kLayers; ++l)
        PS_REQUIRE(net->dW[l] != nullptr && net->dB[l] != nullptr, "field_level_bwd: gradient buffer %d is null", l);
    FieldArgs a{};
    for (int l = 0; l < kLayers; ++l) {
        a.net.W[l] = net->W[l]; a.net.B[l] = net->B[l]; a.net.dW[l] = net->dW[l]; a.net.dB[l] = net->dB[l];
    }
    a.net.in_dim = L * F;
    a.net.app_dim = net->app_dim;
    a.feat = feat_lm; a.dfeat = dfeat_lm; a.L = L; a.F = F; a.sel = sel; a.eu = eu_bins; a.dirs = dirs; a.app = app;
    a.dapp = dapp; a.N = N; a.S = S;
    a.acc = const_cast<float*>(acc); a.dexp = const_cast<float*>(depth_exp);
    a.d_w = d_weights; a.d_rgb = d_rgb_out; a.d_acc = d_acc; a.d_dexp = d_depth_exp; a.d_sem = d_sem_out;
    if (!use_v1()) return launch_field_bwd2(a, (cudaStream_t)stream);
    if (L * F <= 32) return launch_field_bwd<32>(a, (cudaStream_t)stream);
    return launch_field_bwd<48>(a, (cudaStream_t)stream);
}

int ps_field_ms_check(const ps_field_net_dev* nets_dev, int L, int F, int64_t rows, int S, int app_dim, const char* what);

extern "C" int ps_field_level_bwd_ms(const ps_field_net_dev* nets_dev, int app_dim, const float* feat_lm_sorted, int L,
                                     int F, const uint8_t* sel_sorted, const int32_t* perm, const uint8_t* tile_sf,
                                     int64_t rows, int S, const float* dirs, const float* app, const float* d_density,
                                     const float* d_rgb, const float* d_sem, float* dfeat_lm_sorted, float* dapp,
                                     void* stream) {
    if (int e = ps_field_ms_check(nets_dev, L, F, rows, S, app_dim, "field_level_bwd_ms")) return e;
    PS_REQUIRE(feat_lm_sorted && sel_sorted && perm && tile_sf && dirs && d_density && d_rgb && d_sem && dfeat_lm_sorted,
               "field_level_bwd_ms: null pointer");
    PS_REQUIRE(app_dim == 0 || app != nullptr, "field_level_bwd_ms: appearance is null");
    FieldMsArgs a{};
    a.nets = reinterpret_cast<const FieldNet*>(nets_dev);
    a.feat = feat_lm_sorted; a.dfeat = dfeat_lm_sorted; a.L = L; a.F = F; a.sels = sel_sorted; a.perm = perm;
    a.tile_sf = tile_sf; a.rows = rows; a.S = S; a.dirs = dirs; a.app = app; a.dapp = dapp;
    a.d_density = d_density; a.d_rgb = d_rgb; a.d_sem = d_sem;
    if (!use_v1()) return launch_field_bwd2_ms(a, (cudaStream_t)stream);
    if (L * F <= 32) return launch_field_bwd_ms<32>(a, (cudaStream_t)stream);
    return launch_field_bwd_ms<48>(a, (cudaStream_t)stream);
}
