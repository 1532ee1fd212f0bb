// Fused field level, backward: recomputes the forward of field_tc5_fwd.cu on chip and produces the gradient of the hash
// features, of the appearance embedding and of every weight / bias — one kernel (recompute, then every input / weight / bias
// gradient; reference: fields/PreSight/ingp_field.py:163-251, field_components/activations.py:28-41, cameras/rays.py:128-150,
// model_components/renderers.py:70-117, 286-314, 332-383, nerfacto_nusc_ms.py:497-530).  Structured around what the phase
// clocks of the first version showed (tools/phase_clocks.py, profiles/r2_b_phase_clocks.txt): a tile was a serial chain of
// 16 GEMM -> epilogue phases of ~1.7 k cycles each with the tensor pipe idle during every epilogue.
//
// Here the colour head and the semantic head — independent once the base network's output H exists — run as two
// concurrent chains on the same 128-point tile: warps 0-3 (group R, one thread per point) drive the colour head and the
// compositing, warps 4-7 (group S) the semantic head.  Each chain has its own issuing warp, commit barriers, named
// barrier and 64-column accumulator, so one head's epilogue runs under the other head's GEMMs: 2 + 6 + 2 dependent phases
// per tile instead of 16.  What makes both heads' tiles fit in 227 KB of shared memory:
//   * gradients are written IN PLACE: dZ_{k-1} = dA_{k-1} * relu'(A_{k-1}) overwrites the activation tile A_{k-1} (row-local);
//     the weight-gradient GEMM that still reads A_{k-1} is committed to a second barrier that is waited before the store,
//     while the TMEM load and the packing run under it;
//   * the dH tile doubles as the dZ tile of both heads' output layers (columns 0-15 colour, 16-79 semantics).
// The GEMMs are issued by two extra warps (8: base network + colour chain, 9: semantic chain) that do nothing else: a
// tcgen05.mma costs its issuing thread 40-80 cycles, a backward phase has ~20 of them, and on an epilogue warp that time sat
// on the chain's critical path (first two-chain version: no faster than the single chain).  Epilogue threads only ARRIVE on
// the "operands ready" named barrier and go straight to the commit barrier of the result they need.
// What makes the accumulators fit in 512 TMEM columns: weight gradients of the 64-row layers are M = 64 GEMMs, two
// accumulators per column range where both are fed by the same issuing thread (tc5.cuh:kLaneHi) — 352 columns for all
// weight / bias gradients beside the two 64-column working accumulators (B2Tmem) — which also cost 25 % less tensor time.
#include <type_traits>

#include "field_tc5.cuh"

// Phase clocks of the debug build (PS_NVCC_DEFS=-DPS_PHASE_CLOCKS, tools/phase_clocks.py): threads 0 (group R), 128 (group S),
// 256 / 288 (issuing warps) of CTA 0 stamp (code, clock64) pairs of one steady-state tile, 128 pairs per role.
#ifdef PS_PHASE_CLOCKS
__device__ long long* g_phase_buf_bwd2 = nullptr;
#define PS_STAMP(code)                                                        \
    do {                                                                      \
        if (stamp_on && nstamp < 128) {                                       \
            g_phase_buf_bwd2[(stamp_role * 128 + nstamp) * 2] = (code);       \
            g_phase_buf_bwd2[(stamp_role * 128 + nstamp) * 2 + 1] = clock64(); \
            ++nstamp;                                                         \
        }                                                                     \
    } while (0)
#else
#define PS_STAMP(code)
#endif

namespace ps {
namespace ftc5 {

constexpr int kB2Epi = 256;                 // epilogue threads: group R (0-127), group S (128-255)
constexpr int kB2Threads = kB2Epi + 64;     // + the two issuing warps

template <int K0>
struct B2Smem {
    using WL = WLayout<K0>;
    // DH is the only M = 128 "X" operand (80 real columns): its 48 padding columns are read from the tile behind it
    static constexpr uint32_t dh = ((WL::end + 127) / 128) * 128;        // [128 x 80]  dH; dZ of R2 (0-15) / S2 (16-79)
    static constexpr uint32_t h1 = dh + cm_bytes(kRows, 80);             // [128 x 64]  base hidden, later dZ_B0
    static constexpr uint32_t h = h1 + cm_bytes(kRows, 64);              // [128 x 80]
    static constexpr uint32_t x0 = h + cm_bytes(kRows, 80);              // [128 x K0]
    static constexpr uint32_t shapp = x0 + cm_bytes(kRows, K0);          // [128 x 32]
    static constexpr uint32_t a1r = shapp + cm_bytes(kRows, 32);         // [128 x 64]  colour hidden 1, later its dZ
    static constexpr uint32_t a2r = a1r + cm_bytes(kRows, 64);
    static constexpr uint32_t a1s = a2r + cm_bytes(kRows, 64);           // semantic hidden 1 / 2
    static constexpr uint32_t a2s = a1s + cm_bytes(kRows, 64);
    static constexpr uint32_t tiles_end = a2s + cm_bytes(kRows, 64);
    static constexpr uint32_t rayc = tiles_end;                          // float [4 rays][72]
    static constexpr uint32_t wbuf = rayc + 4 * 72 * 4;                  // float [128]: weights, group R -> group S
    static constexpr uint32_t dsem = wbuf + 128 * 4;                     // float [128]: <sem, d_sem>, group S -> group R
    static constexpr uint32_t tails = dsem + 128 * 4;                    // double [2][4]
    static constexpr uint32_t bars = tails + 2 * 4 * 8;                  // 7 mbarriers + tmem slot
    // The next tile's hash features (fp32, level-major [L][128][F] as they lie in global memory: L * F * 512 bytes) are copied
    // by bulk copies (TMA) into the semantic chain's two activation tiles, which are dead from the join to the next tile's
    // semantic forward: no shared memory of their own, for any K0.
    static constexpr uint32_t stage = a1s;
    static constexpr uint32_t stage_cap = 2 * cm_bytes(kRows, 64);
    static constexpr uint32_t total = bars + 64;
};

// TMEM columns.  Working accumulators first; the base network uses accR .. accR + 80.
template <int K0>
struct B2Tmem {
    // Accumulators that share columns (or bias-gradient columns) are always fed by the same issuing thread.
    static constexpr int accR = 0, accS = 64;
    static constexpr int p1 = 128;          // s0 (lanes lo) | s1 (lanes hi), [64 x 64] each       (semantic chain)
    static constexpr int s2 = p1 + 64;      // [64 x 64]                                            (semantic chain)
    static constexpr int bs = s2 + 64;      // 16 columns: kBiasCol0 + l = bias gradient of S0 / S1 / S2 (semantic chain)
    static constexpr int p2 = bs + 16;      // r1 [64 x 64] (lo) | b0 [64 x K0] (hi)                (warp 0)
    static constexpr int r0 = p2 + 64;      // [64 x 48]
    static constexpr int b1 = r0 + 48;      // [80 x 64], M = 128
    static constexpr int r2 = b1 + 64;      // [64 (k) x 16]: columns 0-2 = dW_r2^T, kBiasCol0 + l = bias gradient of R0 / R1 / B0
    static constexpr int bb1 = r2 + 16;     // 16 columns, M = 128: column kBiasCol0 + B1 = bias gradient of B1
    static constexpr int end = bb1 + 16;
    static_assert(K0 <= 64 && end <= 512, "TMEM budget");
};

// 64 accumulator columns -> ReLU -> bf16 -> tile columns [0, 64)
__device__ __forceinline__ void relu_epilogue64(uint32_t tacc, unsigned char* tile, int r) {
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
        float v[32];
        tmem_ld32_nowait(tacc + c, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; i += 8) store_chunk_relu(tile, kRows, r, c + i, v + i);
    }
}

// 32 input-gradient columns gated by the forward activation they belong to -> packed bf16 (4 x 16 bytes, not stored yet)
__device__ __forceinline__ void dgrad_pack32(uint32_t tacc, const unsigned char* tile, int r, int c0, uint4 (&q)[4]) {
    float v[32];
    tmem_ld32_nowait(tacc, v);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint4 a = *reinterpret_cast<const uint4*>(tile + cm_off(kRows, r, c0 + 8 * i));
        q[i].x = relu_grad_pack_bf16x2(v[8 * i], v[8 * i + 1], a.x);
        q[i].y = relu_grad_pack_bf16x2(v[8 * i + 2], v[8 * i + 3], a.y);
        q[i].z = relu_grad_pack_bf16x2(v[8 * i + 4], v[8 * i + 5], a.z);
        q[i].w = relu_grad_pack_bf16x2(v[8 * i + 6], v[8 * i + 7], a.w);
    }
}
__device__ __forceinline__ void store_packed32(unsigned char* tile, int r, int c0, const uint4 (&q)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(tile + cm_off(kRows, r, c0 + 8 * i)) = q[i];
}

// bias gradient of one layer on the tensor core (field_tc5.cuh:gemm_bias_grad), M = 64 or 128
__device__ __forceinline__ void gemm_bias_grad_m(uint32_t tmem_d, uint32_t dz_addr, uint32_t onehot_addr, int M,
                                                 bool accumulate) {
    const uint32_t idesc = make_idesc_m(M, 16, 1, 1);
#pragma unroll
    for (int kk = 0; kk < kRows / 16; ++kk)
        umma_bf16(tmem_d, make_desc(dz_addr + kk * 256, 128, kRows * 16), make_desc(onehot_addr, 0, 128), idesc,
                  (accumulate || kk > 0) ? 1u : 0u);
}

// warp column sums of 16 values per lane: afterwards lane l < 16 holds the total of column l
__device__ __forceinline__ float column_sums16(const float (&u)[16], int lane) {
    float t[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = u[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] += __shfl_xor_sync(0xffffffffu, t[i], 16);
#pragma unroll
    for (int step = 0, half = 8; step < 4; ++step, half >>= 1) {
        const int bit = 8 >> step;
        const bool upper = (lane & bit) != 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < half) {
                const float keep = upper ? t[j + half] : t[j];
                const float send = upper ? t[j] : t[j + half];
                t[j] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
            }
        }
    }
    return t[0];
}

template <int K0, typename Args>
__global__ void __maxnreg__(128) field_bwd2_kernel(Args a) {
    constexpr bool MS = std::is_same<Args, FieldMsArgs>::value;
    using WL = WLayout<K0>;
    using SM = B2Smem<K0>;
    using TM = B2Tmem<K0>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, g = tid >> 7, r = tid & 127, warp = r >> 5, lane = tid & 31;
    unsigned char* wbase = smem;
    unsigned char* DH = smem + SM::dh;
    unsigned char* H1 = smem + SM::h1;
    unsigned char* Ht = smem + SM::h;
    unsigned char* X0 = smem + SM::x0;
    unsigned char* SHAPPt = smem + SM::shapp;
    unsigned char* A1r = smem + SM::a1r;
    unsigned char* A2r = smem + SM::a2r;
    unsigned char* A1s = smem + SM::a1s;
    unsigned char* A2s = smem + SM::a2s;
    float* rayc = reinterpret_cast<float*>(smem + SM::rayc);
    float* wbuf = reinterpret_cast<float*>(smem + SM::wbuf);
    float* dsem = reinterpret_cast<float*>(smem + SM::dsem);
    double* tails = reinterpret_cast<double*>(smem + SM::tails);
    uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(smem + SM::bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::bars + 56);

    // all tiles start out as zeros (padding columns / rows are read by GEMMs and must be finite)
    for (uint32_t i = SM::dh + tid * 16; i < SM::tiles_end; i += kB2Threads * 16)
        *reinterpret_cast<uint4*>(smem + i) = make_uint4(0u, 0u, 0u, 0u);
    if (tid < 32) tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        for (int i = 0; i < 7; ++i) mbar_init(smem_u32(bar_ptr + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    FieldNet net{};
    if constexpr (!MS) {
        net = a.net;
        load_all_weights<K0>(net, wbase, tid, kB2Threads);
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    // commit barriers: base network (d = result the epilogue waits for, w = weight-gradient group), colour chain, semantic chain
    const uint32_t barBd = smem_u32(bar_ptr), barBw = smem_u32(bar_ptr + 1), barRd = smem_u32(bar_ptr + 2),
                   barRw = smem_u32(bar_ptr + 3), barSd = smem_u32(bar_ptr + 4), barSw = smem_u32(bar_ptr + 5),
                   barF = smem_u32(bar_ptr + 6);           // next tile's features have landed in the staging buffer
    const uint32_t wb = smem_u32(wbase), ones = wb + WL::ones, onehot = wb + WL::onehot;
    const uint32_t aDH = smem_u32(DH), aH1 = smem_u32(H1), aH = smem_u32(Ht), aX0 = smem_u32(X0), aSH = smem_u32(SHAPPt),
                   aA1r = smem_u32(A1r), aA2r = smem_u32(A2r), aA1s = smem_u32(A1s), aA2s = smem_u32(A2s);
    constexpr uint32_t CH = kRows * 16;     // bytes per 8-column chunk of a 128-row tile
    uint32_t phBd = 0, phBw = 0, phRd = 0, phRw = 0, phSd = 0, phSw = 0, phF = 0;
    const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);      // provably warp-uniform warp index (issue branches)

    const int S = a.S;
    int rpt = 1, rows_used = kRows, wpr = 1;
    int64_t P = 0, ntiles;
    if constexpr (MS) {
        ntiles = a.rows / kRows;
    } else {
        rpt = kRows / S;
        rows_used = rpt * S;
        wpr = S / 32;
        P = a.N * S;
        ntiles = (a.N + rpt - 1) / rpt;
    }
    bool first = true;
    float db_r2 = 0.f;      // bias gradient of the colour head's output layer (3 values: lanes 0-2 of every group-R warp)
    int cur = -1;           // MS: sub-field whose weights are staged

    // ---- feature prefetch: one lane of warp 9 copies tile t's rows of every level (contiguous in the level-major layout)
    // into the staging area (B2Smem::stage) during the base-network backward of tile t - gridDim.x ------------------------
    int64_t feat_rows;      // rows per level of the feature array
    if constexpr (MS) feat_rows = a.rows;
    else feat_rows = P;
    const bool use_pref = (uint32_t)(a.L * a.F) * kRows * 4 <= SM::stage_cap && (reinterpret_cast<uintptr_t>(a.feat) & 15) == 0 &&
                          ((feat_rows * a.F * 4) & 15) == 0;
    auto prefetch_tile = [&](int64_t t) {        // called by ONE thread
        if (t >= ntiles) return;
        int64_t p0;
        int npts;
        if constexpr (MS) {
            if (a.tile_sf[t] == 255) return;
            p0 = t * kRows;
            npts = kRows;
        } else {
            const int64_t ray0 = t * rpt;
            const int64_t nr = a.N - ray0 < rpt ? a.N - ray0 : rpt;
            p0 = ray0 * S;
            npts = (int)nr * S;
        }
        const uint32_t bytes_l = (uint32_t)npts * a.F * 4;
        mbar_expect_tx(barF, bytes_l * a.L);
        for (int l = 0; l < a.L; ++l)
            bulk_g2s(smem_u32(smem + SM::stage) + (uint32_t)l * kRows * a.F * 4, a.feat + ((int64_t)l * feat_rows + p0) * a.F, bytes_l,
                     barF);
    };
    if (use_pref && warp_u == 9) {
        if (elect_one()) prefetch_tile(blockIdx.x);
        __syncwarp();
    }

// epilogue side: this thread's operands are written (generic proxy -> async proxy) and its TMEM reads are done; it does
// NOT wait here — the issuing warp does (bar_sync on the same id) — but goes on to the commit barrier of the result it needs.
// Barrier ids (immediates, see tc5.cuh): 1 group R internal, 2 / 3 weights / <sem, d_sem> hand-offs between the groups,
// 4 end of tile (epilogue threads), 5 base network operands and fork (256 + both issuing warps), 6 colour chain (128 + warp 8),
// 7 semantic chain (128 + warp 9)
constexpr int kBarR = 1, kBarW = 2, kBarD = 3, kBarEnd = 4, kBarBase = 5, kBarCR = 6, kBarCS = 7;
#define B2_READY(ID, COUNT) \
    PS_STAMP(0);            \
    fence_async_smem();     \
    fence_before();         \
    bar_arrive<ID, COUNT>();
#define B2_READY_BASE() B2_READY(kBarBase, kB2Epi + 64)
#define B2_READY_R() B2_READY(kBarCR, 128 + 32)
#define B2_READY_S() B2_READY(kBarCS, 128 + 32)
// issuing side: wait for the operands, then one elected lane issues
#define B2_ISSUE(ID, COUNT, ...) \
    bar_sync<ID, COUNT>();       \
    PS_STAMP(1);                 \
    if (elect_one()) {           \
        fence_after();           \
        __VA_ARGS__;             \
    }                            \
    __syncwarp();                \
    PS_STAMP(2);
#define B2_WAIT(BAR, PH)  \
    PS_STAMP(5);          \
    mbar_wait(BAR, PH);   \
    PH ^= 1;              \
    fence_after();        \
    PS_STAMP(3);

    // weight and bias gradients: TMEM accumulators -> global atomics (end of the kernel / of a sub-field)
    auto flush_all = [&]() {
        if (first) return;
        const int A = net.app_dim;
        fence_after();
        // an M = 64 accumulator: row m of the layer sits in lane 32 * (m / 16) + m % 16 (+ 16 for the second one of a pair)
        auto flush64 = [&](int col0, bool hi, int ncols, float* dW, int k_real, int kind) {
            const int n = warp * 16 + (lane & 15);
            const bool mine = ((lane >> 4) != 0) == hi;
            for (int blk = 0; blk < ncols / 16; ++blk) {
                if ((blk & 1) != g) continue;
                float u[16];
                tmem_ld16_nowait(trow + col0 + 16 * blk, u);
                tmem_wait_ld();
                if (!mine) continue;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    int k = 16 * blk + i;
                    if (kind == 1) {               // colour head layer 0: staged column -> reference column
                        if (k < 16) {}
                        else if (k == 16) k = -1;
                        else if (k < 32) k -= 1;
                        else k = (k - 32 < A) ? k - 1 : -1;
                    }
                    if (k >= 0 && k < k_real && u[i] != 0.f) atomicAdd(dW + (size_t)n * k_real + k, u[i]);
                }
            }
        };
        flush64(TM::p1, false, kSem, net.dW[S0], kSem, 0);
        flush64(TM::p1, true, kHid, net.dW[S1], kHid, 0);
        flush64(TM::s2, false, kHid, net.dW[S2], kHid, 0);
        flush64(TM::p2, false, kHid, net.dW[R1], kHid, 0);
        flush64(TM::p2, true, K0, net.dW[B0], net.in_dim, 0);
        flush64(TM::r0, false, kRgbIn, net.dW[R0], 16 + kGeo + A, 1);
        // base layer 1: M = 128, row = lane
        for (int blk = 0; blk < kHid / 16; ++blk) {
            if ((blk & 1) != g) continue;
            float u[16];
            tmem_ld16_nowait(trow + TM::b1 + 16 * blk, u);
            tmem_wait_ld();
            const int n = warp * 32 + lane;
            if (n < kBaseOut) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (u[i] != 0.f) atomicAdd(net.dW[B1] + (size_t)n * kHid + 16 * blk + i, u[i]);
            }
        }
        {
            // g == 0: colour head / base layer 0 bias columns and dW_r2^T;  g == 1: semantic head bias columns, base layer 1
            float u[16], ub[16];
            tmem_ld16_nowait(trow + (g == 0 ? TM::r2 : TM::bs), u);
            tmem_ld16_nowait(trow + TM::bb1, ub);
            tmem_wait_ld();
            const int m64 = warp * 16 + (lane & 15), m128 = warp * 32 + lane;
            if (lane < 16) {
                if (g == 0) {
                    // transposed accumulator [k][n]: dW_r2[n][k]
#pragma unroll
                    for (int n = 0; n < 3; ++n) atomicAdd(net.dW[R2] + (size_t)n * kHid + m64, u[n]);
                }
#pragma unroll
                for (int l = 0; l < kLayers; ++l) {
                    const bool sem_layer = l == S0 || l == S1 || l == S2;
                    if (l == R2 || l == B1 || sem_layer != (g == 1)) continue;
                    if (net.dB[l]) atomicAdd(net.dB[l] + m64, u[kBiasCol0 + l]);
                }
            }
            if (g == 1 && m128 < kBaseOut && net.dB[B1]) atomicAdd(net.dB[B1] + m128, ub[kBiasCol0 + B1]);
            if (g == 0 && lane < 3 && net.dB[R2]) atomicAdd(net.dB[R2] + lane, db_r2);
        }
        fence_before();
        db_r2 = 0.f;
        first = true;
    };

#ifdef PS_PHASE_CLOCKS
    int nstamp = 0, tile_iter = 0;
#endif
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if constexpr (MS) {
            const int sf = a.tile_sf[tile];
            if (sf == 255) break;
            if (sf != cur) {
                // the finished sub-field's gradients leave TMEM, the next sub-field's weights are staged (all 320 threads)
                if (cur >= 0) {
                    if (tid < kB2Epi) flush_all();
                    else first = true;
                }
                __syncthreads();
                net = a.nets[sf];
                load_all_weights<K0>(net, wbase, tid, kB2Threads);
                fence_async_smem();
                __syncthreads();
                cur = sf;
            }
        }
        const int A = net.app_dim;
        const bool acc_dw = !first;
        first = false;
#ifdef PS_PHASE_CLOCKS
        const int stamp_role = tid >> 7 == 2 ? 2 + ((tid >> 5) & 1) : tid >> 7;
        const bool stamp_on = g_phase_buf_bwd2 != nullptr && blockIdx.x == 0 && (tid & 127 & ~32) == 0 && (tid < 256 || tid == 256 || tid == 288) &&
                              (tid == 0 || tid == 128 || tid == 256 || tid == 288) && tile_iter == 3;
        ++tile_iter;
        PS_STAMP(9);
#endif
        if (tid >= kB2Epi) {
            // =========================== the two issuing warps =====================================================
            if (warp_u == 8) {
                // base network, forward
                B2_ISSUE(kBarBase, kB2Epi + 64, gemm_bias(tmem + TM::accR, ones, wb + WL::bt(B0), wb + WL::btz, kHid);
                         gemm_kk(tmem + TM::accR, aX0, kRows, wb + WL::b0, kHid, kHid, K0, true); umma_commit(barBd))
                B2_ISSUE(kBarBase, kB2Epi + 64, gemm_bias(tmem + TM::accR, ones, wb + WL::bt(B1), wb + WL::btz, kBaseOut);
                         gemm_kk(tmem + TM::accR, aH1, kRows, wb + WL::b1, kBaseOut, kBaseOut, kHid, true); umma_commit(barBd))
                // colour head, forward
                B2_ISSUE(kBarBase, kB2Epi + 64, gemm_bias(tmem + TM::accR, ones, wb + WL::bt(R0), wb + WL::btz, kHid);
                         gemm_kk(tmem + TM::accR, aSH, kRows, wb + WL::r0, kHid, kHid, 16, true);
                         gemm_kk(tmem + TM::accR, aH, kRows, wb + WL::r0 + 2 * kHid * 16, kHid, kHid, 16, true);
                         gemm_kk(tmem + TM::accR, aSH + 2 * CH, kRows, wb + WL::r0 + 4 * kHid * 16, kHid, kHid, 16, true);
                         umma_commit(barRd))
                B2_ISSUE(kBarCR, 128 + 32, gemm_bias(tmem + TM::accR, ones, wb + WL::bt(R1), wb + WL::btz, kHid);
                         gemm_kk(tmem + TM::accR, aA1r, kRows, wb + WL::r1, kHid, kHid, kHid, true); umma_commit(barRd))
                B2_ISSUE(kBarCR, 128 + 32, gemm_bias(tmem + TM::accR, ones, wb + WL::bt(R2), wb + WL::btz, kRgbOut);
                         gemm_kk(tmem + TM::accR, aA2r, kRows, wb + WL::r2, kRgbOut, kRgbOut, kHid, true); umma_commit(barRd))
                // colour head, backward: input gradient first (its epilogue starts under the weight / bias gradient GEMMs)
                B2_ISSUE(kBarCR, 128 + 32, gemm_dgrad(tmem + TM::accR, aDH, kRows, wb + WL::r2, kRgbOut, kHid, 16, false);
                         umma_commit(barRd);
                         gemm_wgrad64(tmem + TM::r2, aA2r, aDH, 16, acc_dw); umma_commit(barRw))
                B2_ISSUE(kBarCR, 128 + 32, gemm_dgrad(tmem + TM::accR, aA2r, kRows, wb + WL::r1, kHid, kHid, kHid, false);
                         umma_commit(barRd);
                         gemm_wgrad64(tmem + TM::p2, aA2r, aA1r, kHid, acc_dw);
                         gemm_bias_grad_m(tmem + TM::r2, aA2r, onehot + R1 * 256, 64, true); umma_commit(barRw))
                B2_ISSUE(kBarCR, 128 + 32, gemm_dgrad(tmem + TM::accR, aA1r, kRows, wb + WL::r0, kHid, kRgbIn, kHid, false);
                         umma_commit(barRd);
                         gemm_wgrad64(tmem + TM::r0, aA1r, aSH, 16, acc_dw);
                         gemm_wgrad64(tmem + TM::r0 + 16, aA1r, aH, 16, acc_dw);
                         gemm_wgrad64(tmem + TM::r0 + 32, aA1r, aSH + 2 * CH, 16, acc_dw);
                         gemm_bias_grad_m(tmem + TM::r2, aA1r, onehot + R0 * 256, 64, true); umma_commit(barRw))
                // base network, backward
                B2_ISSUE(kBarBase, kB2Epi + 64, gemm_dgrad(tmem + TM::accR, aDH, kRows, wb + WL::b1, kBaseOut, kHid, kBaseOut, false);
                         umma_commit(barBd);
                         gemm_wgrad(tmem + TM::b1, aDH, aH1, kHid, acc_dw);
                         gemm_bias_grad_m(tmem + TM::bb1, aDH, onehot + B1 * 256, 128, acc_dw); umma_commit(barBw))
                B2_ISSUE(kBarBase, kB2Epi + 64, gemm_dgrad(tmem + TM::accR, aH1, kRows, wb + WL::b0, kHid, K0, kHid, false);
                         umma_commit(barBd);
                         gemm_wgrad64(tmem + TM::p2 + kLaneHi, aH1, aX0, K0, acc_dw);
                         gemm_bias_grad_m(tmem + TM::r2, aH1, onehot + B0 * 256, 64, true); umma_commit(barBw))
            } else {
                // (the base network's operand barrier counts both issuing warps: pass its two forward generations)
                bar_sync<kBarBase, kB2Epi + 64>();
                if constexpr (!MS) {
                    // every epilogue thread has staged this tile's inputs: the next tile's small per-ray / per-sample inputs
                    // are pulled into L2 (one line per lane)
                    const int64_t nt = tile + gridDim.x;
                    if (nt < ntiles) {
                        const int64_t ray0 = nt * rpt, p0 = ray0 * S;
                        const char* ptrs[6] = {reinterpret_cast<const char*>(a.eu + ray0 * (S + 1)),
                                               reinterpret_cast<const char*>(a.d_w ? a.d_w + p0 : nullptr),
                                               reinterpret_cast<const char*>(a.d_sem ? a.d_sem + ray0 * kSem : nullptr),
                                               reinterpret_cast<const char*>(a.sel ? a.sel + p0 : nullptr),
                                               reinterpret_cast<const char*>(a.dirs + ray0 * 3),
                                               reinterpret_cast<const char*>(a.app ? a.app + ray0 * A : nullptr)};
                        const int bytes[6] = {rpt * (S + 1) * 4, kRows * 4, rpt * kSem * 4, kRows, rpt * 12, rpt * A * 4};
                        int k = lane;
#pragma unroll
                        for (int j = 0; j < 6; ++j) {
                            const int lines = (bytes[j] + 127) / 128 + 1;
                            if (ptrs[j] && k >= 0 && k < lines) prefetch_l2(ptrs[j] + k * 128);
                            k -= lines;
                        }
                    }
                }
                bar_sync<kBarBase, kB2Epi + 64>();
                // semantic head, forward (sub-field mode: the output layer is linear and its dZ is an input, no recompute)
                B2_ISSUE(kBarBase, kB2Epi + 64, gemm_bias(tmem + TM::accS, ones, wb + WL::bt(S0), wb + WL::btz, kHid);
                         gemm_kk(tmem + TM::accS, aH + 2 * CH, kRows, wb + WL::s0, kHid, kHid, kSem, true); umma_commit(barSd))
                B2_ISSUE(kBarCS, 128 + 32, gemm_bias(tmem + TM::accS, ones, wb + WL::bt(S1), wb + WL::btz, kHid);
                         gemm_kk(tmem + TM::accS, aA1s, kRows, wb + WL::s1, kHid, kHid, kHid, true); umma_commit(barSd))
                if constexpr (!MS) {
                    B2_ISSUE(kBarCS, 128 + 32, gemm_bias(tmem + TM::accS, ones, wb + WL::bt(S2), wb + WL::btz, kSem);
                             gemm_kk(tmem + TM::accS, aA2s, kRows, wb + WL::s2, kSem, kSem, kHid, true); umma_commit(barSd))
                }
                // semantic head, backward
                B2_ISSUE(kBarCS, 128 + 32, gemm_dgrad(tmem + TM::accS, aDH + 2 * CH, kRows, wb + WL::s2, kSem, kHid, kSem, false);
                         umma_commit(barSd);
                         gemm_wgrad64(tmem + TM::s2, aDH + 2 * CH, aA2s, kHid, acc_dw);
                         gemm_bias_grad_m(tmem + TM::bs, aDH + 2 * CH, onehot + S2 * 256, 64, acc_dw); umma_commit(barSw))
                B2_ISSUE(kBarCS, 128 + 32, gemm_dgrad(tmem + TM::accS, aA2s, kRows, wb + WL::s1, kHid, kHid, kHid, false);
                         umma_commit(barSd);
                         gemm_wgrad64(tmem + TM::p1 + kLaneHi, aA2s, aA1s, kHid, acc_dw);
                         gemm_bias_grad_m(tmem + TM::bs, aA2s, onehot + S1 * 256, 64, true); umma_commit(barSw))
                B2_ISSUE(kBarCS, 128 + 32, gemm_dgrad(tmem + TM::accS, aA1s, kRows, wb + WL::s0, kHid, kSem, kHid, false);
                         umma_commit(barSd);
                         gemm_wgrad64(tmem + TM::p1, aA1s, aH + 2 * CH, kSem, acc_dw);
                         gemm_bias_grad_m(tmem + TM::bs, aA1s, onehot + S0 * 256, 64, true); umma_commit(barSw))
                bar_sync<kBarBase, kB2Epi + 64>();      // ... and the two backward generations of the base network
                // (join passed: every epilogue thread has seen the semantic chain's last commit, A1s / A2s are dead — the next
                // tile's features land there while the base network's backward runs)
                if (use_pref) {
                    if (elect_one()) prefetch_tile(tile + gridDim.x);
                    __syncwarp();
                }
                bar_sync<kBarBase, kB2Epi + 64>();
            }
            continue;
        }
        // ---- this thread's row ---------------------------------------------------------------------------------
        bool valid;
        int64_t ray, p, row_i = 0;
        int s = 0, q = 0;
        if constexpr (MS) {
            row_i = tile * kRows + r;
            const int32_t pp = a.perm[row_i];
            valid = pp >= 0;
            p = pp;
            ray = valid ? p / S : 0;
        } else {
            q = r / S;
            s = r - q * S;
            ray = tile * rpt + q;
            valid = r < rows_used && ray < a.N;
            p = ray * S + s;
        }
        // ---- stage inputs: group R the hash features, group S [sh | appearance] -----------------------------------
        // (every global load of the tile start is issued before the first one is consumed: one memory latency, not two)
        float t0 = 0.f, t1 = 0.f, selv, gw_in = 0.f;
        float g_den = 0.f, g_rgb[3] = {0.f, 0.f, 0.f};
        float rcv[2] = {0.f, 0.f};         // this thread's share of the per-ray upstream gradients (rayc)
        const float* rc = rayc;
        if constexpr (MS) {
            if (g == 0 && valid) {
                g_den = __ldg(a.d_density + p);
                g_rgb[0] = __ldg(a.d_rgb + p * 3);
                g_rgb[1] = __ldg(a.d_rgb + p * 3 + 1);
                g_rgb[2] = __ldg(a.d_rgb + p * 3 + 2);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int i = tid + j * kB2Epi;
                if (i < rpt * 72) {
                    const int qq = i / 72, c = i - qq * 72;
                    const int64_t rr = tile * rpt + qq;
                    float v = 0.f;
                    if (rr < a.N) {
                        if (c < 64) v = a.d_sem ? __ldg(a.d_sem + rr * kSem + c) : 0.f;
                        else if (c < 67) v = a.d_rgb ? __ldg(a.d_rgb + rr * 3 + (c - 64)) : 0.f;
                        else if (c == 67) v = a.d_acc ? __ldg(a.d_acc + rr) : 0.f;
                        else if (c == 68) v = a.d_dexp ? __ldg(a.d_dexp + rr) : 0.f;
                        else if (c == 69) v = a.d_dexp ? __ldg(a.dexp + rr) : 0.f;
                        else if (c == 70) v = a.d_dexp ? __ldg(a.acc + rr) : 0.f;
                    }
                    rcv[j] = v;
                }
            }
            rc = rayc + (q < rpt ? q : 0) * 72;
        }
        {
            RowInputs<K0> in;
            const bool direct = g == 0 && !use_pref;
            if constexpr (MS) load_row_inputs_ms<K0>(a, net, row_i, (int32_t)(valid ? p : -1), direct, g == 1, in);
            else load_row_inputs<K0>(a, P, ray, s, valid, direct, g == 1, in);
            if (g == 0) {
                if (use_pref) {
                    // this tile's features were copied into the staging buffer while the previous tile was processed
                    mbar_wait(barF, phF);
                    phF ^= 1;
                    if (valid) {
                        const float* stg = reinterpret_cast<const float*>(smem + SM::stage);
                        if (a.F == 2) {
#pragma unroll
                            for (int l = 0; l < K0 / 2; ++l)
                                if (l < a.L) {
                                    const float2 qv = *reinterpret_cast<const float2*>(stg + (l * kRows + r) * 2);
                                    in.feat[2 * l] = qv.x;
                                    in.feat[2 * l + 1] = qv.y;
                                }
                        } else {
#pragma unroll
                            for (int l = 0; l < K0 / 4; ++l)
                                if (l < a.L) {
                                    const float4 qv = *reinterpret_cast<const float4*>(stg + (l * kRows + r) * 4);
                                    in.feat[4 * l] = qv.x; in.feat[4 * l + 1] = qv.y; in.feat[4 * l + 2] = qv.z; in.feat[4 * l + 3] = qv.w;
                                }
                        }
                    }
                }
                stage_features<K0>(in, X0, r);
            } else {
                stage_shapp<K0>(in, valid, SHAPPt, r);
            }
            t0 = in.t0; t1 = in.t1; selv = in.selv; gw_in = in.gw;
        }
        if constexpr (!MS) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int i = tid + j * kB2Epi;
                if (i < rpt * 72) {
                    const int c = i % 72;
                    rayc[i] = (c == 70 && a.d_dexp) ? 1.f / (rcv[j] + 1e-10f) : rcv[j];
                }
            }
        }
        PS_STAMP(8);
        // ---- base network, forward (both groups: half of the columns each) ----------------------------------------
        B2_READY_BASE()
        B2_WAIT(barBd, phBd)
        {
            float v[32];
            tmem_ld32_nowait(trow + TM::accR + 32 * g, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; i += 8) store_chunk_relu(H1, kRows, r, 32 * g + i, v + i);
        }
        B2_READY_BASE()
        B2_WAIT(barBd, phBd)
        float raw = 0.f;
        {
            // group R: columns 0..47 (raw density, geo, first 32 semantic inputs); group S: columns 48..79
            float v[32];
            const int c0 = g == 0 ? 0 : 48;
            tmem_ld32_nowait(trow + TM::accR + c0, v);
            tmem_wait_ld();
            raw = v[0];
#pragma unroll
            for (int i = 0; i < 32; i += 8) store_chunk(Ht, kRows, r, c0 + i, v + i);
            if (g == 0) {
                float u[16];
                tmem_ld16_nowait(trow + TM::accR + 32, u);
                tmem_wait_ld();
                store_chunk(Ht, kRows, r, 32, u);
                store_chunk(Ht, kRows, r, 40, u + 8);
            }
        }
        B2_READY_BASE()          // H complete: the two chains fork
        if (g == 0) {
            // =========================== colour chain + compositing (group R) ====================================
            // weights of the ray (rays.py:138-148), under the GEMM
            float w = 0.f, T = 0.f, dd = 0.f, dl = 0.f, tm = 0.f;
            bool finite = false;
            int w_first = 0;
            if constexpr (!MS) {
                const float density = valid ? expf(raw) * selv : 0.f;
                dl = __fsub_rn(t1, t0);
                dd = __fmul_rn(dl, density);
                const double dd_incl = warp_scan_incl((double)dd, lane);
                if (lane == 31) tails[warp] = dd_incl;
                bar_sync<kBarR, 128>();
                w_first = (warp / wpr) * wpr;
                double carry = 0.0;
                for (int k = w_first; k < warp; ++k) carry += tails[k];
                const double incl = dd_incl + carry;
                const double prev = __shfl_up_sync(0xffffffffu, incl, 1);
                const double excl = lane == 0 ? carry : prev;
                T = expf(-(float)excl);
                const float alpha = __fsub_rn(1.f, expf(-dd));
                const float rawp = __fmul_rn(alpha, T);
                w = nan_to_num(rawp);
                finite = isfinite(rawp) && valid;
                if (!valid) w = 0.f;
                tm = __fdiv_rn(__fadd_rn(t0, t1), 2.f);
                wbuf[r] = w;
                __threadfence_block();
                bar_arrive<kBarW, kB2Epi>();         // group S waits for the weights before its output layer's epilogue
            }
            B2_WAIT(barRd, phRd)
            relu_epilogue64(trow + TM::accR, A1r, r);
            B2_READY_R()
            B2_WAIT(barRd, phRd)
            relu_epilogue64(trow + TM::accR, A2r, r);
            B2_READY_R()
            B2_WAIT(barRd, phRd)
            // ---- colour head, backward ---------------------------------------------------------------------------
            float dot_rgb = 0.f;
            {
                float u[16];
                tmem_ld16_nowait(trow + TM::accR, u);
                tmem_wait_ld();
                float dz[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) dz[i] = 0.f;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float y = sigmoid_f(u[i]);
                    if constexpr (MS) {
                        dz[i] = g_rgb[i] * y * (1.f - y);
                    } else {
                        dot_rgb += y * rc[64 + i];
                        dz[i] = w * rc[64 + i] * y * (1.f - y);
                    }
                }
                store_chunk(DH, kRows, r, 0, dz);
                store_chunk(DH, kRows, r, 8, dz + 8);
                const float s0 = warp_sum(dz[0]), s1 = warp_sum(dz[1]), s2 = warp_sum(dz[2]);
                db_r2 += lane == 0 ? s0 : (lane == 1 ? s1 : s2);
            }
            B2_READY_R()
            {
                uint4 qa[4], qb[4];
                B2_WAIT(barRd, phRd)
                dgrad_pack32(trow + TM::accR, A2r, r, 0, qa);
                dgrad_pack32(trow + TM::accR + 32, A2r, r, 32, qb);
                B2_WAIT(barRw, phRw)          // the weight-gradient GEMM has read A2r: overwrite it with dZ
                store_packed32(A2r, r, 0, qa);
                store_packed32(A2r, r, 32, qb);
            }
            B2_READY_R()
            {
                uint4 qa[4], qb[4];
                B2_WAIT(barRd, phRd)
                dgrad_pack32(trow + TM::accR, A1r, r, 0, qa);
                dgrad_pack32(trow + TM::accR + 32, A1r, r, 32, qb);
                B2_WAIT(barRw, phRw)
                store_packed32(A1r, r, 0, qa);
                store_packed32(A1r, r, 32, qb);
            }
            B2_READY_R()
            B2_WAIT(barRd, phRd)
            float d_h01[16];      // gradient of h[0:16] from the colour head (column 0 is zero by construction)
            {
                tmem_ld16_nowait(trow + TM::accR + 16, d_h01);
                float u[16];
                tmem_ld16_nowait(trow + TM::accR + 32, u);
                tmem_wait_ld();
                if (a.dapp && A > 0) {
                    if constexpr (MS) {
                        // rows of a tile belong to different rays: one atomic per (row, channel)
                        if (valid)
#pragma unroll
                            for (int k = 0; k < 16; ++k)
                                if (k < A) atomicAdd(a.dapp + ray * A + k, u[k]);
                    } else {
                        // sum over the ray's samples (a warp lies inside one ray)
                        const float tot = column_sums16(u, lane);
                        const int64_t wray = tile * rpt + (warp * 32) / S;
                        if (lane < A && warp * 32 < rows_used && wray < a.N) atomicAdd(a.dapp + wray * A + lane, tot);
                    }
                }
            }
            float d_raw;
            if constexpr (MS) {
                d_raw = valid ? g_den * selv * expf(fminf(fmaxf(raw, -15.f), 15.f)) : 0.f;
            } else {
                // compositing backward.  pass 1: total gradient on this weight
                bar_sync<kBarD, kB2Epi>();           // group S has published <sem, d_sem>
                float gt = gw_in + rc[67] + rc[68] * (tm - rc[69]) * rc[70] + dsem[r] + dot_rgb;
                if (!finite) gt = 0.f;
                const double gw_incl = warp_scan_incl((double)gt * (double)w, lane);
                if (lane == 31) tails[4 + warp] = gw_incl;
                bar_sync<kBarR, 128>();
                // pass 2: d sigma_i = delta_i * (g_i T_{i+1} - sum_{k>i} g_k w_k); d raw = d sigma * sel * exp(clamp(raw))
                double pc = 0.0, G = 0.0;
                for (int k = w_first; k < w_first + wpr; ++k) {
                    if (k < warp) pc += tails[4 + k];
                    G += tails[4 + k];
                }
                const double Pi = gw_incl + pc;
                const float d_sigma = dl * (float)((double)gt * (double)(T * expf(-dd)) - (G - Pi));
                d_raw = valid ? d_sigma * selv * expf(fminf(fmaxf(raw, -15.f), 15.f)) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) d_h01[i] = valid ? d_h01[i] : 0.f;
            d_h01[0] = d_raw;
            // (the dZ of the output layer that lived in these columns was last read by GEMMs waited for above)
            store_chunk(DH, kRows, r, 0, d_h01);
            store_chunk(DH, kRows, r, 8, d_h01 + 8);
            B2_WAIT(barRw, phRw)              // SHAPP / H / A1r reads complete
        } else {
            // =========================== semantic chain (group S) ================================================
            if constexpr (MS) {
                // dZ of the (linear) output layer = the per-point gradient itself: straight into the dH tile, under the GEMM
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    float v[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k) v[k] = 0.f;
                    if (valid) {
                        const float4* src = reinterpret_cast<const float4*>(a.d_sem + p * kSem + 32 * hh);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float4 qv = __ldg(src + k);
                            v[4 * k] = qv.x; v[4 * k + 1] = qv.y; v[4 * k + 2] = qv.z; v[4 * k + 3] = qv.w;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 32; k += 8) store_chunk(DH, kRows, r, 16 + 32 * hh + k, v + k);
                }
            }
            B2_WAIT(barSd, phSd)
            relu_epilogue64(trow + TM::accS, A1s, r);
            B2_READY_S()
            B2_WAIT(barSd, phSd)
            relu_epilogue64(trow + TM::accS, A2s, r);
            if constexpr (!MS) {
                B2_READY_S()
                bar_sync<kBarW, kB2Epi>();           // group R has published the weights
                const float w = wbuf[r];
                B2_WAIT(barSd, phSd)
                float dot = 0.f;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    float v[32];
                    tmem_ld32_nowait(trow + TM::accS + 32 * hh, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float gs = rc[32 * hh + i];
                        dot += v[i] * gs;
                        v[i] = w * gs;
                    }
#pragma unroll
                    for (int i = 0; i < 32; i += 8) store_chunk(DH, kRows, r, 16 + 32 * hh + i, v + i);
                }
                dsem[r] = dot;
                __threadfence_block();
                bar_arrive<kBarD, kB2Epi>();
            }
            B2_READY_S()
            {
                uint4 qa[4], qb[4];
                B2_WAIT(barSd, phSd)
                dgrad_pack32(trow + TM::accS, A2s, r, 0, qa);
                dgrad_pack32(trow + TM::accS + 32, A2s, r, 32, qb);
                B2_WAIT(barSw, phSw)
                store_packed32(A2s, r, 0, qa);
                store_packed32(A2s, r, 32, qb);
            }
            B2_READY_S()
            {
                uint4 qa[4], qb[4];
                B2_WAIT(barSd, phSd)
                dgrad_pack32(trow + TM::accS, A1s, r, 0, qa);
                dgrad_pack32(trow + TM::accS + 32, A1s, r, 32, qb);
                B2_WAIT(barSw, phSw)
                store_packed32(A1s, r, 0, qa);
                store_packed32(A1s, r, 32, qb);
            }
            B2_READY_S()
            B2_WAIT(barSd, phSd)
            // gradient of h[16:80] (the dZ of the output layer that lived in these columns was last read by GEMMs waited above)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float v[32];
                tmem_ld32_nowait(trow + TM::accS + 32 * hh, v);
                tmem_wait_ld();
                if (!valid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < 32; i += 8) store_chunk(DH, kRows, r, 16 + 32 * hh + i, v + i);
            }
            B2_WAIT(barSw, phSw)              // H / A1s reads complete
        }
        // ---- join; base network, backward: dH = [d raw | colour head (15) | semantic head (64)] -------------------------
        B2_READY_BASE()
        {
            uint4 qa[4];
            B2_WAIT(barBd, phBd)
            dgrad_pack32(trow + TM::accR + 32 * g, H1, r, 32 * g, qa);
            B2_WAIT(barBw, phBw)
            store_packed32(H1, r, 32 * g, qa);
        }
        B2_READY_BASE()
        B2_WAIT(barBd, phBd)
        // ---- hash-feature gradient (level-major) -------------------------------------------------------------------
        if (a.dfeat) {
            int64_t rows_total, idx;
            bool wr;
            if constexpr (MS) { rows_total = a.rows; idx = row_i; wr = true; }
            else { rows_total = P; idx = p; wr = valid; }
#pragma unroll
            for (int blk = 0; blk < K0 / 16; ++blk) {
                if ((blk & 1) != g) continue;
                float u[16];
                tmem_ld16_nowait(trow + TM::accR + 16 * blk, u);
                tmem_wait_ld();
                if (!valid) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) u[i] = 0.f;
                }
                if (wr) {
                    if (a.F == 2) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int l = 8 * blk + i;
                            if (l < a.L)
                                *reinterpret_cast<float2*>(a.dfeat + ((int64_t)l * rows_total + idx) * 2) = make_float2(u[2 * i], u[2 * i + 1]);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int l = 4 * blk + i;
                            if (l < a.L)
                                *reinterpret_cast<float4*>(a.dfeat + ((int64_t)l * rows_total + idx) * 4) =
                                    make_float4(u[4 * i], u[4 * i + 1], u[4 * i + 2], u[4 * i + 3]);
                        }
                    }
                }
            }
        }
        B2_WAIT(barBw, phBw)          // X0 / H1 reads complete: every tile may be restaged
        PS_STAMP(7);
        fence_before();
        bar_sync<kBarEnd, kB2Epi>();              // ... and every thread is done with the accumulators
    }
#undef B2_READY
#undef B2_READY_BASE
#undef B2_READY_R
#undef B2_READY_S
#undef B2_ISSUE
#undef B2_WAIT
    if (tid < kB2Epi) flush_all();
    fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem, 512);
}

template <int K0, typename Args>
static int launch_b2(const Args& a, int64_t ntiles, const char* what, cudaStream_t stream) {
    constexpr size_t smem = B2Smem<K0>::total;
    static_assert(smem <= 227 * 1024, "field_bwd2: shared memory");
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(field_bwd2_kernel<K0, Args>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess) {
            set_error("%s: cannot reserve %zu bytes of shared memory", what, smem);
            return 2;
        }
        configured = true;
    }
    const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
    field_bwd2_kernel<K0, Args><<<grid, kB2Threads, smem, stream>>>(a);
    return check_launch(what);
}

static int launch_field_bwd2(const FieldArgs& a, cudaStream_t stream) {
    const int rpt = kRows / a.S;
    const int64_t ntiles = (a.N + rpt - 1) / rpt;
    if (a.L * a.F <= 32) return launch_b2<32, FieldArgs>(a, ntiles, "field_level_bwd", stream);
    return launch_b2<48, FieldArgs>(a, ntiles, "field_level_bwd", stream);
}

static int launch_field_bwd2_ms(const FieldMsArgs& a, cudaStream_t stream) {
    const int64_t ntiles = a.rows / kRows;
    if (a.L * a.F <= 32) return launch_b2<32, FieldMsArgs>(a, ntiles, "field_level_bwd_ms", stream);
    return launch_b2<48, FieldMsArgs>(a, ntiles, "field_level_bwd_ms", stream);
}

}  // namespace ftc5
}  // namespace ps

using namespace ps;
using namespace ps::ftc5;

int ps_field_check_common(const ps_field_net* net, int L, int F, int64_t N, int S, const char* what);


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
    return launch_field_bwd2(a, (cudaStream_t)stream);
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
    return launch_field_bwd2_ms(a, (cudaStream_t)stream);
}

#ifdef PS_PHASE_CLOCKS
/* tools only (debug build) */
extern "C" int ps_debug_phase_buf_bwd2(long long* buf) {
    return cudaMemcpyToSymbol(g_phase_buf_bwd2, &buf, sizeof(buf)) == cudaSuccess ? 0 : 2;
}
#endif
