// Self-test of the tcgen05 operand conventions of tc5.cuh (K-major / MN-major descriptors over one chunk-major
// tile, M = 128 padding rows, accumulation across GEMM calls).  Test-only entry point, one CTA.
#include "tc5.cuh"

namespace ps {
namespace tc5 {

// X [128 x 64], Y [128 x 64], W [64 x 64] fp32 -> C1 = X W^T, C2 = X W, C3 = 2 X^T Y  (all [.. x 64] fp32),
// C4 [128 x 16]: column 3 = column sums of X (rows 0..63), other columns zero
__global__ void __launch_bounds__(128) tc5_probe_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                                        const float* __restrict__ W, float* __restrict__ C1,
                                                        float* __restrict__ C2, float* __restrict__ C3,
                                                        float* __restrict__ C4) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* Xs = smem;                       // 16 KB
    unsigned char* Ys = Xs + cm_bytes(128, 64);     // 16 KB (also the "garbage" behind X for the M = 128 wgrad)
    unsigned char* Ws = Ys + cm_bytes(128, 64);     // 8 KB
    unsigned char* Oh = Ws + cm_bytes(64, 64);     // 256 B: one-hot block (column 3) + zero block
    uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(Oh + 256);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar_ptr + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int c0 = 0; c0 < 64; c0 += 8) {
        float v[8], u[8];
        for (int i = 0; i < 8; ++i) { v[i] = X[tid * 64 + c0 + i]; u[i] = Y[tid * 64 + c0 + i]; }
        store_chunk(Xs, 128, tid, c0, v);
        store_chunk(Ys, 128, tid, c0, u);
    }
    load_weight_cm(W, 64, 64, 64, 64, Ws, nullptr, tid, 128);
    if (tid < 128) reinterpret_cast<__nv_bfloat16*>(Oh)[tid] = __float2bfloat16_rn((tid < 64 && (tid & 7) == 3) ? 1.f : 0.f);
    const uint32_t bar = smem_u32(bar_ptr);
    if (warp == 0) tmem_alloc(slot, 256);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *slot;
    if (tid == 0) {
        gemm_kk(tmem, smem_u32(Xs), 128, smem_u32(Ws), 64, 64, 64, false);
        gemm_dgrad(tmem + 64, smem_u32(Xs), 128, smem_u32(Ws), 64, 64, 64, false);
        gemm_wgrad(tmem + 128, smem_u32(Xs), smem_u32(Ys), 64, false);
        gemm_wgrad(tmem + 128, smem_u32(Xs), smem_u32(Ys), 64, true);
        // column sums of X through a one-hot B operand whose K stride (LBO) is zero: every 8-row group of the
        // reduction re-reads the same 128-byte block {0,0,0,1,0,0,0,0} x 8; the second 8-column chunk is a zero block
        for (int kk = 0; kk < 8; ++kk)
            umma_bf16(tmem + 192, make_desc(smem_u32(Xs) + kk * 256, 128, 128 * 16), make_desc(smem_u32(Oh), 0, 128),
                      make_idesc(16, 1, 1), kk > 0 ? 1u : 0u);
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    fence_after();
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    {
        float v[16];
        tmem_ld16_nowait(trow + 192, v);
        tmem_wait_ld();
        for (int i = 0; i < 16; ++i) C4[tid * 16 + i] = v[i];
    }
    float* outs[3] = {C1, C2, C3};
    for (int m = 0; m < 3; ++m)
        for (int c = 0; c < 64; c += 32) {
            float v[32];
            tmem_ld32_nowait(trow + 64 * m + c, v);
            tmem_wait_ld();
            for (int i = 0; i < 32; ++i) outs[m][tid * 64 + c + i] = v[i];
        }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace tc5
}  // namespace ps

extern "C" int ps_tc5_probe(const float* X, const float* Y, const float* W, float* C1, float* C2, float* C3,
                            float* C4, void* stream) {
    using namespace ps;
    PS_REQUIRE(X && Y && W && C1 && C2 && C3 && C4, "tc5_probe: null pointer");
    const size_t smem = tc5::cm_bytes(128, 64) * 2 + tc5::cm_bytes(64, 64) + 256 + 64;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(tc5::tc5_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess) {
            set_error("tc5_probe: cannot reserve %zu bytes of shared memory", smem);
            return 2;
        }
        configured = true;
    }
    tc5::tc5_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(X, Y, W, C1, C2, C3, C4);
    return check_launch("tc5_probe");
}

// ---- timing probe: cycles per tcgen05.mma as a function of N, operand majorness, number of independent accumulators
// and number of issuing threads (tools/mma_cost.py) ----------------------------------------------------------------
namespace ps {
namespace tc5 {
// n_acc independent accumulators used round-robin (n_acc * N <= 512 columns); n_issuers threads (thread 0 of the
// first n_issuers warps) issue n_mma instructions each, each issuer on its own accumulator set when n_acc >= n_issuers.
// out[0] = cycles until issuer 0 has issued its last instruction, out[1] = until every issuer's commit has arrived.
__global__ void __launch_bounds__(128) mma_cost_kernel(int N, int mn_major, int n_mma, int n_acc, int M, int n_issuers,
                                                       long long* out) {
    extern __shared__ __align__(128) unsigned char smem[];      // 64 KB of zeros: operands
    uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(smem + 65536);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar_ptr + 4);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 0) tmem_alloc(slot, 512);
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(bar_ptr + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *slot;
    __shared__ long long ts[2];
    long long t0 = 0, t1 = 0;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    if (warp_u < n_issuers) {
      if (elect_one()) {
        const uint32_t a = smem_u32(smem), b = smem_u32(smem) + 32768;
        // make_idesc encodes M = 128; patch the M field (bits 24..28 = M >> 4) for the M = 64 variant
        const uint32_t idesc = (make_idesc(N, mn_major, mn_major) & ~(0x1Fu << 24)) | ((uint32_t)(M >> 4) << 24);
        const int per = n_acc >= n_issuers ? n_acc / n_issuers : 1;       // accumulators per issuer
        const int base = n_acc >= n_issuers ? warp * per : 0;
        t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            const uint32_t d = tmem + (uint32_t)((base + i % per) * N);
            if (mn_major)
                umma_bf16(d, make_desc(a + (i & 7) * 256, 128, kRows * 16), make_desc(b + (i & 7) * 256, 128, kRows * 16), idesc, 1u);
            else
                umma_bf16(d, make_desc(a + (i & 3) * 4096, kRows * 16, 128), make_desc(b + (i & 3) * 4096, kRows * 16, 128), idesc, 1u);
        }
        t1 = clock64();
        if (warp == 0) { ts[0] = t0; ts[1] = t1; }
        umma_commit(smem_u32(bar_ptr + warp));
      }
      __syncwarp();
    }
    if (tid == 0) {
        for (int i = 0; i < n_issuers; ++i) mbar_wait(smem_u32(bar_ptr + i), 0);
        const long long t2 = clock64();
        out[0] = ts[1] - ts[0];
        out[1] = t2 - ts[0];
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}
}  // namespace tc5
}  // namespace ps

/* tools only (not declared in include/presight_b200.h) */
extern "C" int ps_tc5_mma_cost(int N, int mn_major, int n_mma, int n_acc, int M, int n_issuers, long long* out,
                               void* stream) {
    using namespace ps;
    PS_REQUIRE(n_acc >= 1 && n_acc * N <= 512 && n_issuers >= 1 && n_issuers <= 4 && (!mn_major || N <= 128),
               "tc5_mma_cost: bad shape");
    const size_t smem = 65536 + 64;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(tc5::mma_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    tc5::mma_cost_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(N, mn_major, n_mma, n_acc, M, n_issuers, out);
    return check_launch("tc5_mma_cost");
}

// ---- latency probe (tools/mma_cost.py --lat): the fixed costs of one GEMM -> epilogue phase of the fused kernels, measured
// by thread 0 of a single 128-thread CTA with clock64 ------------------------------------------------------------
namespace ps {
namespace tc5 {
__global__ void __launch_bounds__(128) phase_lat_kernel(long long* out) {
    extern __shared__ __align__(128) unsigned char smem[];      // 64 KB of zeros: operands
    uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(smem + 65536);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar_ptr + 4);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 0) tmem_alloc(slot, 512);
    if (tid == 0) {
        mbar_init(smem_u32(bar_ptr), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t bar = smem_u32(bar_ptr);
    const uint32_t a = smem_u32(smem), b = a + 32768;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    uint32_t phase = 0;
    long long t[24];
    for (int i = 0; i < 24; ++i) t[i] = 0;
    const uint32_t idesc64 = make_idesc(64, 0, 0), idesc16 = make_idesc(16, 0, 0), idesc64mn = make_idesc(64, 1, 1);
    // n MMAs (K-major, N = 64) + commit -> wait: round trip as seen by the issuing thread
#define RT(NM, IDESC, SLOT)                                                                      \
    __syncthreads();                                                                             \
    t[SLOT] = clock64();                                                                         \
    if (warp_u == 0) {                                                                           \
        if (elect_one()) {                                                                       \
            _Pragma("unroll") for (int i = 0; i < NM; ++i)                                       \
                umma_bf16(tmem, make_desc(a + (i & 3) * 4096, kRows * 16, 128),                  \
                          make_desc(b + (i & 3) * 4096, kRows * 16, 128), IDESC, 1u);            \
            t[SLOT + 1] = clock64();                                                             \
            umma_commit(bar);                                                                    \
        }                                                                                        \
        __syncwarp();                                                                            \
    }                                                                                            \
    mbar_wait(bar, phase);                                                                       \
    phase ^= 1;                                                                                  \
    fence_after();                                                                               \
    t[SLOT + 2] = clock64();
    RT(1, idesc64, 0)
    RT(4, idesc64, 3)
    RT(16, idesc64, 6)
    RT(16, idesc16, 9)
    // MN-major (weight-gradient form)
    __syncthreads();
    t[12] = clock64();
    if (warp_u == 0) {
        if (elect_one()) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                umma_bf16(tmem + 64, make_desc(a + (i & 7) * 256, 128, kRows * 16), make_desc(b + (i & 7) * 256, 128, kRows * 16),
                          idesc64mn, 1u);
            t[13] = clock64();
            umma_commit(bar);
        }
        __syncwarp();
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    fence_after();
    t[14] = clock64();
    // tcgen05.ld of 32 columns + wait
    float v[32];
    __syncthreads();
    t[15] = clock64();
    tmem_ld32_nowait(trow, v);
    tmem_wait_ld();
    t[16] = clock64();
    // pack + 4 x 16-byte shared stores + proxy fence + tcgen05 fence + barrier
#pragma unroll
    for (int i = 0; i < 32; i += 8) store_chunk_relu(smem, kRows, tid, i, v + i);
    t[17] = clock64();
    fence_async_smem();
    t[18] = clock64();
    fence_before();
    __syncthreads();
    t[19] = clock64();
    // two back-to-back tcgen05.ld x32 (64 columns)
    float u[32];
    tmem_ld32_nowait(trow + 32, v);
    tmem_ld32_nowait(trow + 64, u);
    tmem_wait_ld();
    t[20] = clock64();
    float acc = 0.f;
    for (int i = 0; i < 32; ++i) acc += v[i] + u[i];
    if (acc == 123.f) out[31] = 1;
    if (tid == 0)
        for (int i = 0; i < 24; ++i) out[i] = t[i];
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
#undef RT
}
}  // namespace tc5
}  // namespace ps

/* tools only (not declared in include/presight_b200.h) */
extern "C" int ps_tc5_phase_lat(long long* out, void* stream) {
    using namespace ps;
    const size_t smem = 65536 + 64;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(tc5::phase_lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    tc5::phase_lat_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(out);
    return check_launch("tc5_phase_lat");
}

// ---- M = 64 probe (tools/m64_probe.py): where do the 64 rows of a cta_group::1, M = 64 accumulator live in TMEM, and does
// a lane offset of 16 in the D address select the other half of each 32-lane quadrant? ---------------------------------
namespace ps {
namespace tc5 {
// X, Y [128 x 64] fp32.  dump[0] = all 128 lanes x 64 columns after D(lane 0) = X^T Y with M = 64;
// dump[1] = the same columns after a second product D(lane 16) = 2 X^T Y with M = 64 issued on top (lane_off = 16 in the
// D address); cyc[0..1] = issue->complete cycles of 16 MN-major N = 64 MMAs with M = 64 / M = 128.
__global__ void __launch_bounds__(128) m64_probe_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                                        float* __restrict__ dump, long long* cyc, int lane_off) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* Xs = smem;                       // 16 KB
    unsigned char* Ys = Xs + cm_bytes(128, 64);     // 16 KB
    uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(Ys + cm_bytes(128, 64) + 32768);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar_ptr + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int c0 = 0; c0 < 64; c0 += 8) {
        float v[8], u[8];
        for (int i = 0; i < 8; ++i) { v[i] = X[tid * 64 + c0 + i]; u[i] = Y[tid * 64 + c0 + i]; }
        store_chunk(Xs, 128, tid, c0, v);
        store_chunk(Ys, 128, tid, c0, u);
    }
    const uint32_t bar = smem_u32(bar_ptr);
    if (warp == 0) tmem_alloc(slot, 256);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t idesc64 = (make_idesc(64, 1, 1) & ~(0x1Fu << 24)) | ((uint32_t)(64 >> 4) << 24);
    const uint32_t idesc128 = make_idesc(64, 1, 1);
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    uint32_t phase = 0;
    // zero the 64 columns through a product with accumulate = 0 at M = 128 of zeros?  simpler: first product overwrites
    // with M = 128 (all lanes defined), value 4 X^T Y
    if (warp_u == 0) {
        if (elect_one()) {
            for (int kk = 0; kk < 8; ++kk)
                umma_bf16(tmem, make_desc(smem_u32(Xs) + kk * 256, 128, 128 * 16), make_desc(smem_u32(Ys) + kk * 256, 128, 128 * 16),
                          idesc128, kk > 0 ? 1u : 0u);
            umma_commit(bar);
        }
        __syncwarp();
    }
    mbar_wait(bar, phase); phase ^= 1; fence_after();
    // M = 64 product, accumulate = 0, D at lane 0
    if (warp_u == 0) {
        if (elect_one()) {
            for (int kk = 0; kk < 8; ++kk)
                umma_bf16(tmem, make_desc(smem_u32(Xs) + kk * 256, 128, 128 * 16), make_desc(smem_u32(Ys) + kk * 256, 128, 128 * 16),
                          idesc64, kk > 0 ? 1u : 0u);
            umma_commit(bar);
        }
        __syncwarp();
    }
    mbar_wait(bar, phase); phase ^= 1; fence_after();
    for (int c = 0; c < 64; c += 32) {
        float v[32];
        tmem_ld32_nowait(trow + c, v);
        tmem_wait_ld();
        for (int i = 0; i < 32; ++i) dump[tid * 64 + c + i] = v[i];
    }
    fence_before();
    __syncthreads();
    fence_after();
    if (lane_off >= 0) {
        // second M = 64 product: two passes over K (= 2 X^T Y), D address with a lane offset
        if (warp_u == 0) {
            if (elect_one()) {
                for (int rep = 0; rep < 2; ++rep)
                    for (int kk = 0; kk < 8; ++kk)
                        umma_bf16(tmem + ((uint32_t)lane_off << 16), make_desc(smem_u32(Xs) + kk * 256, 128, 128 * 16),
                                  make_desc(smem_u32(Ys) + kk * 256, 128, 128 * 16), idesc64, (rep > 0 || kk > 0) ? 1u : 0u);
                umma_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, phase); phase ^= 1; fence_after();
    }
    for (int c = 0; c < 64; c += 32) {
        float v[32];
        tmem_ld32_nowait(trow + c, v);
        tmem_wait_ld();
        for (int i = 0; i < 32; ++i) dump[128 * 64 + tid * 64 + c + i] = v[i];
    }
    // timing: 16 MN-major MMAs, M = 64 vs M = 128
    for (int m = 0; m < 2; ++m) {
        fence_before();
        __syncthreads();
        fence_after();
        const long long t0 = clock64();
        if (warp_u == 0) {
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    umma_bf16(tmem + 128, make_desc(smem_u32(Xs) + (i & 7) * 256, 128, 128 * 16),
                              make_desc(smem_u32(Ys) + (i & 7) * 256, 128, 128 * 16), m == 0 ? idesc64 : idesc128, 1u);
                umma_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, phase); phase ^= 1; fence_after();
        if (tid == 0) cyc[m] = clock64() - t0;
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}
}  // namespace tc5
}  // namespace ps

/* tools only (not declared in include/presight_b200.h) */
extern "C" int ps_tc5_m64_probe(const float* X, const float* Y, float* dump, long long* cyc, int lane_off, void* stream) {
    using namespace ps;
    const size_t smem = 65536 + 64;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(tc5::m64_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    tc5::m64_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(X, Y, dump, cyc, lane_off);
    return check_launch("tc5_m64_probe");
}
