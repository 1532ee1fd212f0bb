// Warp-level building blocks of the fused MLPs (kernel #2), shared by the stand-alone MLP kernels
// and the fused per-level kernels.
//
// Each warp owns 16 points (rows) and ALL feature columns of those rows, so consecutive layers chain
// through registers: the fp32 accumulator fragment of layer l (m16n8 "C layout": lane (g,t) holds
// rows g, g+8 and columns 2t, 2t+1 of every 8-column block) is re-packed in registers as the A operand
// of layer l+1 — activations never touch shared or global memory between layers.
//
// Two arithmetic modes (PREC):
//   kBF16   : mma.sync m16n8k16 bf16 x bf16 -> fp32            (1e-2 parity class)
//   kTF32x3 : mma.sync m16n8k8 tf32, error-compensated 3-term split (a_hi*b_hi + a_hi*b_lo + a_lo*b_hi)
//             -> fp32-grade results on the tensor pipe                     (1e-3 parity class)
// For tf32 the k index inside each 8-block is permuted (slot t <-> column 2t, slot t+4 <-> column 2t+1)
// consistently on A and B, which makes the C layout directly usable as A operand.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace ps {
namespace mma {

constexpr int kBF16 = 1;
constexpr int kTF32x3 = 0;

template <int PREC>
struct Elem {
    using type = float;
};
template <>
struct Elem<kBF16> {
    using type = __nv_bfloat16;
};

// row stride (in elements) of a [rows][K] shared-memory operand, chosen so that fragment loads/stores and
// ldmatrix rows are bank-conflict free:  bf16: K+8 ;  fp32: == 8 (mod 32)
template <int PREC>
__host__ __device__ constexpr int stride_of(int K) {
    return PREC == kBF16 ? K + 8 : ((K + 31) / 32) * 32 + 8;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// error-compensated product: small terms first
__device__ __forceinline__ void mma_tf32x3(float (&d)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4],
                                           uint32_t bhi0, uint32_t bhi1, uint32_t blo0, uint32_t blo1) {
    mma_tf32(d, alo, bhi0, bhi1);
    mma_tf32(d, ahi, blo0, blo1);
    mma_tf32(d, ahi, bhi0, bhi1);
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}

// A operand (k16 step kk) of a bf16 MMA from a C-layout fp32 activation [K/8][4]
template <int KB>
__device__ __forceinline__ void a_from_c_bf16(const float (&c)[KB][4], int kk, uint32_t (&a)[4]) {
    a[0] = pack_bf16(c[2 * kk][0], c[2 * kk][1]);
    a[1] = pack_bf16(c[2 * kk][2], c[2 * kk][3]);
    a[2] = pack_bf16(c[2 * kk + 1][0], c[2 * kk + 1][1]);
    a[3] = pack_bf16(c[2 * kk + 1][2], c[2 * kk + 1][3]);
}
// A operand (k8 step j, permuted slots) of a tf32 MMA from a C-layout activation
template <int KB>
__device__ __forceinline__ void a_from_c_tf32(const float (&c)[KB][4], int j, uint32_t (&hi)[4], uint32_t (&lo)[4]) {
    split_tf32(c[j][0], hi[0], lo[0]);  // (row g,   slot t   = col 2t)
    split_tf32(c[j][2], hi[1], lo[1]);  // (row g+8, slot t)
    split_tf32(c[j][1], hi[2], lo[2]);  // (row g,   slot t+4 = col 2t+1)
    split_tf32(c[j][3], hi[3], lo[3]);  // (row g+8, slot t+4)
}

// ------------------------------------------------------------------------------------------------
// out[16 x N] = in[16 x K] * W^T + bias.   W in shared memory as [N][stride_of(K)] (nn.Linear layout).
template <int K, int N, int PREC>
__device__ __forceinline__ void layer_forward(const float (&in)[K / 8][4], const typename Elem<PREC>::type* W,
                                              const float* bias, float (&out)[N / 8][4], int lane) {
    constexpr int SW = stride_of<PREC>(K);
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int j = 0; j < N / 8; ++j) {
        const float b0 = bias[8 * j + 2 * t], b1 = bias[8 * j + 2 * t + 1];
        out[j][0] = b0; out[j][1] = b1; out[j][2] = b0; out[j][3] = b1;
    }
    if constexpr (PREC == kBF16) {
#pragma unroll
        for (int kk = 0; kk < K / 16; ++kk) {
            uint32_t a[4];
            a_from_c_bf16<K / 8>(in, kk, a);
#pragma unroll
            for (int j = 0; j < N / 8; ++j) {
                const __nv_bfloat16* w = W + (8 * j + g) * SW + 16 * kk + 2 * t;
                mma_bf16(out[j], a, *reinterpret_cast<const uint32_t*>(w), *reinterpret_cast<const uint32_t*>(w + 8));
            }
        }
    } else {
#pragma unroll
        for (int kj = 0; kj < K / 8; ++kj) {
            uint32_t ahi[4], alo[4];
            a_from_c_tf32<K / 8>(in, kj, ahi, alo);
#pragma unroll
            for (int j = 0; j < N / 8; ++j) {
                const float2 w = *reinterpret_cast<const float2*>(W + (8 * j + g) * SW + 8 * kj + 2 * t);
                uint32_t h0, l0, h1, l1;
                split_tf32(w.x, h0, l0);
                split_tf32(w.y, h1, l1);
                mma_tf32x3(out[j], ahi, alo, h0, h1, l0, l1);
            }
        }
    }
}

// da[16 x K] = dz[16 x N] * W   (input gradient; W is the same forward-layout [N][stride_of(K)] buffer)
template <int K, int N, int PREC>
__device__ __forceinline__ void layer_backward_input(const float (&dz)[N / 8][4], const typename Elem<PREC>::type* W,
                                                     float (&da)[K / 8][4], int lane) {
    constexpr int SW = stride_of<PREC>(K);
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int j = 0; j < K / 8; ++j) da[j][0] = da[j][1] = da[j][2] = da[j][3] = 0.f;
    if constexpr (PREC == kBF16) {
        // B[k = n][n' = k_l] = W[n][k_l] is row-major [K][N] storage -> ldmatrix.trans yields the col-major fragment.
        // x4: matrices {rows n0..n0+7, n0+8..n0+15} x {cols k0..k0+7, k0+8..k0+15}
        const int mrow = (lane & 7) + ((lane >> 3) & 1) * 8;
        const int mcol = (lane >> 4) * 8;
#pragma unroll
        for (int kk = 0; kk < N / 16; ++kk) {
            uint32_t a[4];
            a_from_c_bf16<N / 8>(dz, kk, a);
#pragma unroll
            for (int jo = 0; jo < K / 16; ++jo) {
                uint32_t b[4];
                ldmatrix_x4_trans(b, W + (16 * kk + mrow) * SW + 16 * jo + mcol);
                mma_bf16(da[2 * jo], a, b[0], b[1]);
                mma_bf16(da[2 * jo + 1], a, b[2], b[3]);
            }
        }
    } else {
#pragma unroll
        for (int kj = 0; kj < N / 8; ++kj) {
            uint32_t ahi[4], alo[4];
            a_from_c_tf32<N / 8>(dz, kj, ahi, alo);
#pragma unroll
            for (int jo = 0; jo < K / 8; ++jo) {
                const float w0 = W[(8 * kj + 2 * t) * SW + 8 * jo + g];
                const float w1 = W[(8 * kj + 2 * t + 1) * SW + 8 * jo + g];
                uint32_t h0, l0, h1, l1;
                split_tf32(w0, h0, l0);
                split_tf32(w1, h1, l1);
                mma_tf32x3(da[jo], ahi, alo, h0, h1, l0, l1);
            }
        }
    }
}

// Stage a C-layout [16 x K] fragment into a point-major shared tile [TILE][stride_of(K)] (rows row0+g, row0+g+8).
template <int K, int PREC>
__device__ __forceinline__ void stage_rows(const float (&c)[K / 8][4], typename Elem<PREC>::type* tile, int row0,
                                           int lane) {
    constexpr int S = stride_of<PREC>(K);
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int j = 0; j < K / 8; ++j) {
        if constexpr (PREC == kBF16) {
            *reinterpret_cast<uint32_t*>(tile + (row0 + g) * S + 8 * j + 2 * t) = pack_bf16(c[j][0], c[j][1]);
            *reinterpret_cast<uint32_t*>(tile + (row0 + g + 8) * S + 8 * j + 2 * t) = pack_bf16(c[j][2], c[j][3]);
        } else {
            *reinterpret_cast<float2*>(tile + (row0 + g) * S + 8 * j + 2 * t) = make_float2(c[j][0], c[j][1]);
            *reinterpret_cast<float2*>(tile + (row0 + g + 8) * S + 8 * j + 2 * t) = make_float2(c[j][2], c[j][3]);
        }
    }
}

// Weight-gradient block: acc[16 (n0..) x 16 (f0..)] += sum over the TILE points of dZ[p][n0+m] * Act[p][f0+n'].
// dZs: [TILE][SZ], Acts: [TILE][SA] point-major shared tiles.  acc is two C-layout n8 blocks.
template <int TILE, int PREC>
__device__ __forceinline__ void dw_block(const typename Elem<PREC>::type* dZs, int SZ, int n0,
                                         const typename Elem<PREC>::type* Acts, int SA, int f0, float (&acc)[2][4],
                                         int lane) {
    if constexpr (PREC == kBF16) {
        // A[m = n][k = p] and B[k = p][n' = f]: both stored [k][.] row-major -> ldmatrix.trans for both.
        const int r8 = lane & 7;
        const int a_row = r8 + (lane >> 4) * 8, a_col = ((lane >> 3) & 1) * 8;  // m0: p0..7/n0..7  m1: p0..7/n8..15  m2: p8..15/n0..7  m3: p8..15/n8..15
        const int b_row = r8 + ((lane >> 3) & 1) * 8, b_col = (lane >> 4) * 8;  // m0: p0..7/f0..7  m1: p8..15/f0..7  m2: p0..7/f8..15  m3: p8..15/f8..15
#pragma unroll
        for (int ks = 0; ks < TILE / 16; ++ks) {
            uint32_t a[4], b[4];
            ldmatrix_x4_trans(a, dZs + (16 * ks + a_row) * SZ + n0 + a_col);
            ldmatrix_x4_trans(b, Acts + (16 * ks + b_row) * SA + f0 + b_col);
            mma_bf16(acc[0], a, b[0], b[1]);
            mma_bf16(acc[1], a, b[2], b[3]);
        }
    } else {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll 4
        for (int ks = 0; ks < TILE / 8; ++ks) {
            const float* z0 = dZs + (8 * ks + t) * SZ + n0 + g;
            const float* z1 = dZs + (8 * ks + t + 4) * SZ + n0 + g;
            uint32_t ahi[4], alo[4];
            split_tf32(z0[0], ahi[0], alo[0]);
            split_tf32(z0[8], ahi[1], alo[1]);
            split_tf32(z1[0], ahi[2], alo[2]);
            split_tf32(z1[8], ahi[3], alo[3]);
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) {
                uint32_t h0, l0, h1, l1;
                split_tf32(Acts[(8 * ks + t) * SA + f0 + 8 * nb + g], h0, l0);
                split_tf32(Acts[(8 * ks + t + 4) * SA + f0 + 8 * nb + g], h1, l1);
                mma_tf32x3(acc[nb], ahi, alo, h0, h1, l0, l1);
            }
        }
    }
}

// Cooperative load of an nn.Linear weight [n_real][k_real] (fp32, global) into the padded shared layout
// [N][stride_of(K)], zero-filling the padding, and of the bias into bias_s[N].
template <int K, int N, int PREC>
__device__ __forceinline__ void load_weights(const float* __restrict__ Wg, const float* __restrict__ bg, int n_real,
                                             int k_real, typename Elem<PREC>::type* Ws, float* bias_s, int tid,
                                             int nthreads) {
    constexpr int SW = stride_of<PREC>(K);
    for (int i = tid; i < N * SW; i += nthreads) {
        const int n = i / SW, k = i - n * SW;
        const float v = (n < n_real && k < k_real) ? __ldg(Wg + (size_t)n * k_real + k) : 0.f;
        if constexpr (PREC == kBF16)
            Ws[i] = __float2bfloat16_rn(v);
        else
            Ws[i] = v;
    }
    for (int i = tid; i < N; i += nthreads) bias_s[i] = (i < n_real && bg) ? __ldg(bg + i) : 0.f;
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

}  // namespace mma
}  // namespace ps
