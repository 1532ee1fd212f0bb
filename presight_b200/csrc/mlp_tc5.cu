// Kernel #2, Blackwell-native forward: fused MLP on tcgen05.mma with accumulators in tensor memory (TMEM).
//
// One CTA = 128 threads = one 128-point tile (UMMA M = 128, cta_group::1); thread t owns row t of the tile for the
// row I/O and for the epilogue (TMEM lane t).  Per layer:
//   1. the layer's input tile sits in shared memory as bf16 in the canonical K-major, no-swizzle UMMA layout
//      (8-row x 16-byte core matrices; SBO = distance between 8-row groups, LBO = distance between the two 16-byte
//      K chunks of one K=16 instruction), the weights [N][K] (nn.Linear layout = K-major B operand) likewise;
//   2. one elected thread issues K/16 tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM) and commits them to an
//      mbarrier;
//   3. every thread waits on the mbarrier, reads its row of the accumulator with tcgen05.ld (32x32b), adds the bias,
//      applies ReLU and writes the bf16 row back into the shared A tile as the next layer's input — or, for the last
//      layer, applies the output activation / density epilogue and stores the row to global memory.
// Hidden activations never leave the SM.  Shared/TMEM use is small (<= ~60 KB, <= 128 columns), so several CTAs are
// resident per SM and overlap each other's MMA, epilogue and global I/O phases.
#include <cuda_bf16.h>

#include "mlp_dispatch.cuh"

namespace ps {
namespace tc5 {

using mma::MlpArgs;
using mma::RowSeg;

constexpr int kRows = 128;     // UMMA_M
constexpr int kThreadsTc5 = 128;

// byte offset of element (row r, column k) of a [rows][K] bf16 operand in the canonical K-major no-swizzle layout
__host__ __device__ constexpr uint32_t core_offset(int r, int k, int K) {
    return (uint32_t)(((r >> 3) * (K >> 3) + (k >> 3)) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, LBO, SBO (all >> 4),
// version = 1 (Blackwell), base offset 0, layout type 0 (no swizzle)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
    const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14);
    return ((uint64_t)hi << 32) | lo;
}

// 32-bit instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major, M = 128, N
__host__ __device__ constexpr uint32_t make_idesc(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16 consecutive fp32 accumulator columns of this thread's row (TMEM lane)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint4 pack8(const float* v) {
    uint4 q;
    q.x = mma::pack_bf16(v[0], v[1]);
    q.y = mma::pack_bf16(v[2], v[3]);
    q.z = mma::pack_bf16(v[4], v[5]);
    q.w = mma::pack_bf16(v[6], v[7]);
    return q;
}

template <class S>
struct Layout {
    static constexpr int KA = S::K0 > S::H ? S::K0 : S::H;                        // widest A operand
    static constexpr int NMAX = S::NOUT > S::N0 ? S::NOUT : S::N0;                // widest accumulator
    static constexpr int TMEM_COLS = NMAX <= 32 ? 32 : (NMAX <= 64 ? 64 : (NMAX <= 128 ? 128 : 256));
    static constexpr size_t a_bytes = (size_t)kRows * KA * 2;
    static constexpr size_t w0_bytes = (size_t)S::N0 * S::K0 * 2;
    static constexpr size_t wmid_bytes = (size_t)S::H * S::H * 2;
    static constexpr size_t wlast_bytes = S::NHID > 0 ? (size_t)S::NOUT * S::H * 2 : 0;
    static constexpr size_t bias_floats = S::N0 + S::NMID * S::H + (S::NHID > 0 ? S::NOUT : 0);
    static constexpr size_t off_w0 = a_bytes;
    static constexpr size_t off_wmid = off_w0 + w0_bytes;
    static constexpr size_t off_wlast = off_wmid + S::NMID * wmid_bytes;
    static constexpr size_t off_bias = off_wlast + wlast_bytes;
    static constexpr size_t off_bar = ((off_bias + bias_floats * 4 + 15) / 16) * 16;
    static constexpr size_t total = off_bar + 16;
};

// nn.Linear weight [n_real][k_real] fp32 -> bf16 [N][K] core-matrix layout (zero padded); bias -> fp32
template <int K, int N>
__device__ __forceinline__ void load_weight_core(const float* __restrict__ Wg, const float* __restrict__ bg, int n_real,
                                                 int k_real, unsigned char* Ws, float* bias_s, int tid) {
    for (int i = tid; i < N * K; i += kThreadsTc5) {
        const int n = i / K, k = i - n * K;
        const float v = (n < n_real && k < k_real) ? __ldg(Wg + (size_t)n * k_real + k) : 0.f;
        *reinterpret_cast<__nv_bfloat16*>(Ws + core_offset(n, k, K)) = __float2bfloat16_rn(v);
    }
    for (int i = tid; i < N; i += kThreadsTc5) bias_s[i] = (i < n_real && bg) ? __ldg(bg + i) : 0.f;
}

// issue the K/16 MMAs of one layer: D[128 x N] (+)= A[128 x K] * W[N x K]^T
template <int K, int N>
__device__ __forceinline__ void issue_layer(uint32_t a_addr, uint32_t w_addr, uint32_t tmem_d, uint32_t bar) {
    constexpr uint32_t idesc = make_idesc(N);
    constexpr uint32_t sbo = (K / 8) * 128;   // next 8-row group
    constexpr uint32_t lbo = 128;             // next 16-byte K chunk
#pragma unroll
    for (int kk = 0; kk < K / 16; ++kk) {
        const uint64_t da = make_desc(a_addr + kk * 256, lbo, sbo);
        const uint64_t db = make_desc(w_addr + kk * 256, lbo, sbo);
        umma_bf16(tmem_d, da, db, idesc, kk > 0 ? 1u : 0u);
    }
    umma_commit(bar);
}

// epilogue of a hidden layer: TMEM row -> +bias, ReLU -> bf16 -> next layer's A tile (K_next = N)
template <int N>
__device__ __forceinline__ void hidden_epilogue(uint32_t taddr_row, const float* bias, unsigned char* A, int row) {
#pragma unroll
    for (int c = 0; c < N / 16; ++c) {
        float v[16];
        tmem_ld16(taddr_row + 16 * c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i] + bias[16 * c + i], 0.f);
        *reinterpret_cast<uint4*>(A + core_offset(row, 16 * c, N)) = pack8(v);
        *reinterpret_cast<uint4*>(A + core_offset(row, 16 * c + 8, N)) = pack8(v + 8);
    }
}

template <class S>
__global__ void __launch_bounds__(kThreadsTc5) mlp_fwd_tc5_kernel(MlpArgs a) {
    using L = Layout<S>;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* A = smem;
    unsigned char* W0 = smem + L::off_w0;
    unsigned char* Wmid = smem + L::off_wmid;
    unsigned char* Wlast = smem + L::off_wlast;
    float* bias0 = reinterpret_cast<float*>(smem + L::off_bias);
    float* biasmid = bias0 + S::N0;
    float* biaslast = biasmid + S::NMID * S::H;
    uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(smem + L::off_bar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::off_bar + 8);
    const int tid = threadIdx.x, warp = tid >> 5;

    if constexpr (S::NHID == 0) {
        load_weight_core<S::K0, S::NOUT>(a.W[0], a.b[0], a.out_dim, a.in_dim, W0, bias0, tid);
    } else {
        load_weight_core<S::K0, S::H>(a.W[0], a.b[0], S::H, a.in_dim, W0, bias0, tid);
#pragma unroll
        for (int m = 0; m < S::NMID; ++m)
            load_weight_core<S::H, S::H>(a.W[1 + m], a.b[1 + m], S::H, S::H, Wmid + m * L::wmid_bytes, biasmid + m * S::H,
                                         tid);
        load_weight_core<S::H, S::NOUT>(a.W[S::NHID], a.b[S::NHID], a.out_dim, S::H, Wlast, biaslast, tid);
    }
    const uint32_t bar = smem_u32(bar_ptr);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)L::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();   // weights were written with generic stores; the tensor core reads them via the async proxy
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t taddr_row = tmem_base + ((uint32_t)(warp * 32) << 16);   // this warp's 32 TMEM lanes
    const uint32_t a_addr = smem_u32(A), w0_addr = smem_u32(W0), wmid_addr = smem_u32(Wmid), wlast_addr = smem_u32(Wlast);
    uint32_t phase = 0;

    const int64_t ntiles = (a.P + kRows - 1) / kRows;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t r = tile * kRows + tid;
        const bool valid = r < a.P;
        // ---- input tile -> bf16 -> A tile (K = K0) ----------------------------------------------------------
        if (a.nseg == 1 && a.seg[0].group == 1 && (a.in_dim & 3) == 0 && ((a.seg[0].stride | a.seg[0].col0) & 3) == 0 &&
            (reinterpret_cast<uintptr_t>(a.seg[0].src) & 15) == 0) {
            // one per-point source: consecutive threads read consecutive 16-byte chunks of a row (coalesced) and
            // scatter them into the core-matrix layout
            const int q_per_row = a.in_dim >> 2;                       // float4 chunks per row
            const float* src = a.seg[0].src + a.seg[0].col0;
            const int64_t row_base = tile * kRows;
            for (int i = tid; i < kRows * (S::K0 / 4); i += kThreadsTc5) {
                const int rr = i / (S::K0 / 4), kq = i - rr * (S::K0 / 4);
                float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row_base + rr < a.P && kq < q_per_row)
                    p = __ldg(reinterpret_cast<const float4*>(src + (row_base + rr) * a.seg[0].stride) + kq);
                uint2 packed;
                packed.x = mma::pack_bf16(p.x, p.y);
                packed.y = mma::pack_bf16(p.z, p.w);
                *reinterpret_cast<uint2*>(A + core_offset(rr, 4 * kq, S::K0)) = packed;
            }
        } else {
            const float* base[PS_MLP_MAX_SEGMENTS];
#pragma unroll
            for (int s = 0; s < PS_MLP_MAX_SEGMENTS; ++s) {
                if (s < a.nseg && valid) {
                    const RowSeg& sg = a.seg[s];
                    const int64_t q = sg.group == 1 ? r : (int64_t)((uint32_t)r / (uint32_t)sg.group);
                    base[s] = sg.src + q * sg.stride + sg.col0 - sg.begin;
                } else {
                    base[s] = nullptr;
                }
            }
#pragma unroll
            for (int kc = 0; kc < S::K0 / 8; ++kc) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int col = 8 * kc + i;
                    float x = 0.f;
                    if (valid && col < a.in_dim) {
                        const int s = mma::seg_of(a, col);
                        const float* p = s == 0 ? base[0] : (s == 1 ? base[1] : base[2]);
                        x = __ldg(p + col);
                    }
                    v[i] = x;
                }
                *reinterpret_cast<uint4*>(A + core_offset(tid, 8 * kc, S::K0)) = pack8(v);
            }
        }
        fence_async_smem();
        fence_before();
        __syncthreads();   // A tile complete; every warp has finished reading the previous tile's accumulator
        // ---- layers ------------------------------------------------------------------------------------
        if constexpr (S::NHID > 0) {
            if (tid == 0) {
                fence_after();
                issue_layer<S::K0, S::H>(a_addr, w0_addr, tmem_base, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
            fence_after();
            hidden_epilogue<S::H>(taddr_row, bias0, A, tid);
#pragma unroll
            for (int m = 0; m < S::NMID; ++m) {
                fence_async_smem();
                fence_before();
                __syncthreads();
                if (tid == 0) {
                    fence_after();
                    issue_layer<S::H, S::H>(a_addr, wmid_addr + m * (uint32_t)L::wmid_bytes, tmem_base, bar);
                }
                mbar_wait(bar, phase);
                phase ^= 1;
                fence_after();
                hidden_epilogue<S::H>(taddr_row, biasmid + m * S::H, A, tid);
            }
            fence_async_smem();
            fence_before();
            __syncthreads();
            if (tid == 0) {
                fence_after();
                issue_layer<S::H, S::NOUT>(a_addr, wlast_addr, tmem_base, bar);
            }
        } else {
            if (tid == 0) {
                fence_after();
                issue_layer<S::K0, S::NOUT>(a_addr, w0_addr, tmem_base, bar);
            }
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        fence_after();
        // ---- output epilogue: bias, activation, density, store the row -----------------------------------
        const float* bl = S::NHID > 0 ? biaslast : bias0;
#pragma unroll
        for (int c = 0; c < S::NOUT / 16; ++c) {
            float v[16];
            tmem_ld16(taddr_row + 16 * c, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += bl[16 * c + i];
            if (c == 0 && a.density_out && valid) a.density_out[r] = expf(v[0]) * (a.sel ? (float)a.sel[r] : 1.f);
            if (a.out_act == PS_ACT_SIGMOID) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = mma::sigmoidf(v[i]);
            } else if (a.out_act == PS_ACT_RELU) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if (a.y && valid) {
                float* dst = a.y + r * a.out_dim + 16 * c;
                if (16 * c + 16 <= a.out_dim && (a.out_dim & 3) == 0 && (reinterpret_cast<uintptr_t>(a.y) & 15) == 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        reinterpret_cast<float4*>(dst)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (16 * c + i < a.out_dim) dst[i] = v[i];
                }
            }
        }
        // the next tile's A-tile barrier orders these TMEM reads before the next MMA overwrites the accumulator
    }
    fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)L::TMEM_COLS)
                     : "memory");
    }
}

template <class S>
int launch_fwd_tc5(const MlpArgs& a, cudaStream_t stream) {
    constexpr size_t smem = Layout<S>::total;
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        if (cudaFuncSetAttribute(mlp_fwd_tc5_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess) {
            set_error("mlp_fwd_tc5: cannot reserve %zu bytes of shared memory", smem);
            return 2;
        }
        // resident CTAs per SM: bounded by shared memory (227 KB, +1 KB reserved per CTA), by tensor memory
        // (512 columns) and by a cap of 8 (registers are not a limit: ~42 per thread)
        int by_smem = (int)((227 * 1024) / (smem + 1024));
        int by_tmem = 512 / Layout<S>::TMEM_COLS;
        ctas_per_sm = by_smem < by_tmem ? by_smem : by_tmem;
        if (ctas_per_sm > 8) ctas_per_sm = 8;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        int api = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&api, mlp_fwd_tc5_kernel<S>, kThreadsTc5, smem) == cudaSuccess &&
            api >= 1 && api < ctas_per_sm && api > 1)
            ctas_per_sm = api;
    }
    const int64_t ntiles = (a.P + kRows - 1) / kRows;
    const int grid = (int)(ntiles < (int64_t)kNumSMs * ctas_per_sm ? ntiles : (int64_t)kNumSMs * ctas_per_sm);
    mlp_fwd_tc5_kernel<S><<<grid, kThreadsTc5, smem, stream>>>(a);
    return check_launch("mlp_fwd_tc5");
}

}  // namespace tc5

namespace mma {

#define PS_TC5_CASE(k0, h, nhid, nout) \
    if (K0 == k0 && H == h && NHID == nhid && NOUT == nout) return tc5::launch_fwd_tc5<Shape<k0, h, nhid, nout>>(a, s);

// returns -1 when the shape is not instantiated
int dispatch_tc5_fwd(int K0, int H, int NHID, int NOUT, const MlpArgs& a, cudaStream_t s) {
    PS_MLP_GROUP0(PS_TC5_CASE)
    PS_MLP_GROUP1(PS_TC5_CASE)
    PS_MLP_GROUP2(PS_TC5_CASE)
    return -1;
}

}  // namespace mma
}  // namespace ps
