// Blackwell tensor-core plumbing shared by the tcgen05 kernels: UMMA descriptors, issue / commit / wait,
// TMEM allocation and loads, and the shared-memory operand layout.
//
// Operand layout ("chunk-major"): a [ROWS x COLS] bf16 tile is stored as COLS/8 chunks of ROWS x 16 bytes,
//     byte offset of (row r, col c) = (c / 8) * ROWS * 16 + r * 16 + (c % 8) * 2.
// One 8-row x 16-byte block is a UMMA core matrix (128 contiguous bytes), so the SAME bytes are
//   * a K-major operand  [rows = M/N index, cols = K index]:  LBO (next K chunk) = ROWS*16, SBO (next 8 rows) = 128;
//   * an MN-major operand [cols = M/N index, rows = K index]: SBO (next MN chunk) = ROWS*16, LBO (next 8 K rows) = 128
// (canonical no-swizzle layouts of cute::UMMA::make_umma_desc).  A thread that owns row r writes 16-byte vectors at
// stride ROWS*16 — consecutive threads hit consecutive 16-byte slots, so the stores are bank-conflict free — and the
// activation / gradient tiles written once serve the forward GEMM (K-major A), the input-gradient GEMM (K-major A)
// and the weight-gradient GEMM (MN-major A and B, reduction over the rows = points) without any transpose.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace ps {
namespace tc5 {

constexpr int kRows = 128;  // UMMA_M and rows per tile

__host__ __device__ constexpr uint32_t cm_off(int rows, int r, int c) {
    return (uint32_t)((c >> 3) * rows * 16 + r * 16 + (c & 7) * 2);
}
__host__ __device__ constexpr uint32_t cm_bytes(int rows, int cols) { return (uint32_t)(rows * cols * 2); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 64-bit shared-memory matrix descriptor: start address, LBO, SBO (all >> 4), version 1 (Blackwell), no swizzle
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
    const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14);
    return ((uint64_t)hi << 32) | lo;
}

// instruction descriptor: D = F32, A = B = BF16, M = 128, N; *_mn = 1 selects an MN-major operand
__host__ __device__ constexpr uint32_t make_idesc(int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
}

// the same with an explicit M (64 or 128).  Measured on B200 (tools/m64_probe.py): an M = 64 accumulator of cta_group::1 keeps
// row m in TMEM lane 32 * (m / 16) + m % 16, i.e. the lower 16 lanes of every 32-lane quadrant; a lane offset of 16 in the
// D address selects the upper 16 lanes, so two M = 64 accumulators share one set of columns.  16 MN-major N = 64
// instructions take 785 cycles at M = 64 against 1040 at M = 128 (the A operand read from shared memory halves).
__host__ __device__ constexpr uint32_t make_idesc_m(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
constexpr uint32_t kLaneHi = 16u << 16;   // D-address offset of the second M = 64 accumulator of a column pair

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// ---- bulk asynchronous copies (TMA engine, 1-D): global -> shared, completion counted in bytes on an mbarrier -------------
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// dst / src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// one lane of a fully converged warp (the caller's branch must be warp-uniform): the compiler then knows exactly one
// thread runs the guarded block and keeps the MMA descriptors in uniform registers (no per-instruction waterfall loop)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Named barriers over COUNT threads (a multiple of 32); id 0 is __syncthreads' barrier.  The id is an IMMEDIATE: with a
// register operand ptxas reserves all 16 hardware barriers for the CTA ("used 16 barriers"), and since the SM has 16 in
// all, no other CTA — not even a kernel that uses none — can then share the SM (measured: the fused field kernels and
// the hash kernels stopped overlapping, tools/overlap_probe.py).
template <int ID, int COUNT>
__device__ __forceinline__ void bar_sync() {
    static_assert(ID >= 1 && ID < 16 && COUNT % 32 == 0, "named barrier");
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}

// producer side of a named barrier shared with bar_sync callers: does not wait (count = arriving + waiting threads)
template <int ID, int COUNT>
__device__ __forceinline__ void bar_arrive() {
    static_assert(ID >= 1 && ID < 16 && COUNT % 32 == 0, "named barrier");
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}

// fp32 accumulator columns [col, col + 16 / 32) of this thread's row (TMEM lane = 32 * (warp % 4) + lane).
// The loads are asynchronous: call tmem_wait_ld() before reading the registers.
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float (&v)[16]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
        "%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8(const float* v) {
    uint4 q;
    q.x = pack_bf16x2(v[0], v[1]);
    q.y = pack_bf16x2(v[2], v[3]);
    q.z = pack_bf16x2(v[4], v[5]);
    q.w = pack_bf16x2(v[6], v[7]);
    return q;
}
// packed epilogue arithmetic (two columns per instruction): ReLU on the bf16 pair, and "gradient through a ReLU" as a
// multiply with hgt2(activation, 0) = {1.0, 0.0} — the activation pair is read back from the forward tile, so no mask
// registers are carried from the forward pass.  Both give bit-identical results to the fp32 form followed by rounding.
__device__ __forceinline__ uint32_t relu_pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __hmax2(__floats2bfloat162_rn(lo, hi), __floats2bfloat162_rn(0.f, 0.f));
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t relu_grad_pack_bf16x2(float lo, float hi, uint32_t act) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&act);
    __nv_bfloat162 v = __hmul2(__floats2bfloat162_rn(lo, hi), __hgt2(a, __floats2bfloat162_rn(0.f, 0.f)));
    return *reinterpret_cast<uint32_t*>(&v);
}
// ReLU + store of 8 consecutive columns
__device__ __forceinline__ void store_chunk_relu(unsigned char* tile, int rows, int r, int c0, const float* v) {
    uint4 q;
    q.x = relu_pack_bf16x2(v[0], v[1]);
    q.y = relu_pack_bf16x2(v[2], v[3]);
    q.z = relu_pack_bf16x2(v[4], v[5]);
    q.w = relu_pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(tile + cm_off(rows, r, c0)) = q;
}
// dz[c0..c0+8) = da * (act > 0), act read from the same position of the forward activation tile
__device__ __forceinline__ void store_chunk_relu_grad(unsigned char* dz_tile, const unsigned char* act_tile, int rows, int r,
                                                      int c0, const float* da) {
    const uint4 a = *reinterpret_cast<const uint4*>(act_tile + cm_off(rows, r, c0));
    uint4 q;
    q.x = relu_grad_pack_bf16x2(da[0], da[1], a.x);
    q.y = relu_grad_pack_bf16x2(da[2], da[3], a.y);
    q.z = relu_grad_pack_bf16x2(da[4], da[5], a.z);
    q.w = relu_grad_pack_bf16x2(da[6], da[7], a.w);
    *reinterpret_cast<uint4*>(dz_tile + cm_off(rows, r, c0)) = q;
}
// store 8 consecutive columns [c0, c0+8) of row r of a chunk-major tile (c0 a multiple of 8)
__device__ __forceinline__ void store_chunk(unsigned char* tile, int rows, int r, int c0, const float* v) {
    *reinterpret_cast<uint4*>(tile + cm_off(rows, r, c0)) = pack8(v);
}

// ---- GEMM issue helpers (called by ONE thread) ---------------------------------------------------------------
// D[128 x N] (+)= A[128 x K] * B[N x K]^T.  A: K-major tile with a_rows rows at a_addr (first K chunk of the GEMM);
// B: weight tile [b_rows x K] K-major at b_addr.  K a multiple of 16.
__device__ __forceinline__ void gemm_kk(uint32_t tmem_d, uint32_t a_addr, int a_rows, uint32_t b_addr, int b_rows, int N,
                                        int K, bool accumulate) {
    const uint32_t idesc = make_idesc(N, 0, 0);
    for (int kk = 0; kk < K / 16; ++kk) {
        const uint64_t da = make_desc(a_addr + kk * 2 * a_rows * 16, a_rows * 16, 128);
        const uint64_t db = make_desc(b_addr + kk * 2 * b_rows * 16, b_rows * 16, 128);
        umma_bf16(tmem_d, da, db, idesc, (accumulate || kk > 0) ? 1u : 0u);
    }
}
// Input gradient: D[128 x N] = dZ[128 x K] * W[K x N]  (W stored as the forward weight tile [w_rows = K_out x N_in],
// i.e. an MN-major B operand: MN = in-feature chunks at stride w_rows*16, K = out-feature rows).
// w_addr points at the first in-feature chunk wanted.
__device__ __forceinline__ void gemm_dgrad(uint32_t tmem_d, uint32_t dz_addr, int dz_rows, uint32_t w_addr, int w_rows,
                                           int N, int K, bool accumulate) {
    const uint32_t idesc = make_idesc(N, 0, 1);
    for (int kk = 0; kk < K / 16; ++kk) {
        const uint64_t da = make_desc(dz_addr + kk * 2 * dz_rows * 16, dz_rows * 16, 128);
        const uint64_t db = make_desc(w_addr + kk * 256, 128, w_rows * 16);
        umma_bf16(tmem_d, da, db, idesc, (accumulate || kk > 0) ? 1u : 0u);
    }
}
// Weight gradient: D[m][n] (+)= sum_p X[p][m] * Y[p][n], reduction over the 128 rows (points) of two chunk-major
// tiles with 128 rows.  x_addr / y_addr point at the first column chunk wanted; M is always 128 (the rows of D beyond
// X's real column count are garbage and must be ignored), N = number of Y columns.
__device__ __forceinline__ void gemm_wgrad(uint32_t tmem_d, uint32_t x_addr, uint32_t y_addr, int N, bool accumulate) {
    const uint32_t idesc = make_idesc(N, 1, 1);
    for (int kk = 0; kk < kRows / 16; ++kk) {
        const uint64_t da = make_desc(x_addr + kk * 256, 128, kRows * 16);
        const uint64_t db = make_desc(y_addr + kk * 256, 128, kRows * 16);
        umma_bf16(tmem_d, da, db, idesc, (accumulate || kk > 0) ? 1u : 0u);
    }
}

// Weight gradient with M = 64 (out features <= 64): D[m][n] (+)= sum_p X[p][m] * Y[p][n] over the 128 rows of two
// chunk-major tiles; reads exactly 64 columns of X.  tmem_d may carry kLaneHi.
__device__ __forceinline__ void gemm_wgrad64(uint32_t tmem_d, uint32_t x_addr, uint32_t y_addr, int N, bool accumulate) {
    const uint32_t idesc = make_idesc_m(64, N, 1, 1);
    for (int kk = 0; kk < kRows / 16; ++kk) {
        const uint64_t da = make_desc(x_addr + kk * 256, 128, kRows * 16);
        const uint64_t db = make_desc(y_addr + kk * 256, 128, kRows * 16);
        umma_bf16(tmem_d, da, db, idesc, (accumulate || kk > 0) ? 1u : 0u);
    }
}

// nn.Linear weight [n_real][k_real] fp32 (global) -> bf16 chunk-major tile [N rows][K cols], zero padded.
// kmap (nullable): destination column -> source column (or -1 for a zero column).
__device__ __forceinline__ void load_weight_cm(const float* __restrict__ Wg, int n_real, int k_real, int N, int K,
                                               unsigned char* Ws, const int* kmap, int tid, int nthreads) {
    for (int i = tid; i < N * K; i += nthreads) {
        const int n = i / K, k = i - n * K;
        const int ks = kmap ? kmap[k] : (k < k_real ? k : -1);
        const float v = (n < n_real && ks >= 0) ? __ldg(Wg + (size_t)n * k_real + ks) : 0.f;
        *reinterpret_cast<__nv_bfloat16*>(Ws + cm_off(N, n, k)) = __float2bfloat16_rn(v);
    }
}

}  // namespace tc5
}  // namespace ps
