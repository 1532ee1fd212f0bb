// Kernel #2: stand-alone fused MLP forward / backward built from the warp-level blocks of mlp_mma.cuh.
// One CTA = 8 warps = 128 points per tile, persistent over tiles; weights live in shared memory for the
// whole kernel; hidden activations never leave the SM (backward recomputes the forward).
#pragma once
#include "mlp_mma.cuh"

namespace ps {
namespace mma {

constexpr int kTile = 128;
constexpr int kWarps = kTile / 16;
constexpr int kThreads = kWarps * 32;

// One source of input columns.  The logical input row of point r is the concatenation of up to 3 segments;
// segment s supplies columns [begin, end) from src[(r / group) * stride + col0 + (c - begin)].  group > 1 means
// the segment is shared by `group` consecutive points (a per-ray vector broadcast to the ray's samples).
// On backward, dst (nullable) receives the input gradient with the same addressing; for group > 1 the gradients
// of the group's points are summed (warp reduction + atomicAdd, dst must be zero-filled by the caller).
struct RowSeg {
    const float* src;
    float* dst;
    int64_t stride;
    int col0, begin, end, group;
};

struct MlpArgs {
    RowSeg seg[PS_MLP_MAX_SEGMENTS];
    int nseg;
    int any_group_dst;  // some segment has group > 1 and dst != null
    int want_dx;        // some segment has dst != null
    int vec2_x;         // single per-point segment whose src/dst rows can be accessed as aligned float2 pairs
    int lm_F;           // > 0: the single segment is level-major hash features [L][P][F] (see ps_row_segment)
    int vec2_y;         // y / dy rows can be accessed as aligned float2 pairs
    float* y;           // [P, out_dim] (nullable when only density_out is wanted)
    const float* dy;    // [P, out_dim] (nullable: zero)
    int64_t P;
    const float* W[PS_MAX_MLP_LAYERS];
    const float* b[PS_MAX_MLP_LAYERS];
    float* dW[PS_MAX_MLP_LAYERS];
    float* db[PS_MAX_MLP_LAYERS];
    int in_dim, out_dim;  // real (unpadded) sizes; hidden width is exact
    int out_act;
    // density epilogue (ingp_field.py:189-190 / activations.py:28-41): density = exp(y[:,0]) * sel
    const uint8_t* sel;        // [P] nullable (treated as 1)
    float* density_out;        // [P] nullable
    const float* d_density;    // [P] nullable; when set, the gradient of column 0 is d_density*sel*exp(clamp(raw))
};

// Network archetype: K0 -> H -> ... (NHID hidden layers) ... -> NOUT, all padded to multiples of 16.
template <int K0_, int H_, int NHID_, int NOUT_>
struct Shape {
    static constexpr int K0 = K0_, H = H_, NHID = NHID_, NOUT = NOUT_;
    static constexpr int NL = NHID + 1;
    static constexpr int N0 = NHID == 0 ? NOUT : H;  // width after the first layer
    static constexpr int NMID = NHID > 1 ? NHID - 1 : 0;
    static_assert(K0 % 16 == 0 && H % 16 == 0 && NOUT % 16 == 0, "dims must be padded to multiples of 16");
    static_assert(H <= 64, "ReLU masks are kept in one 32-bit word per layer");
};

template <int KB>
__device__ __forceinline__ void load_rows(const float* __restrict__ x, int64_t P, int dim, int64_t row0, int lane,
                                          float (&c)[KB][4]) {
    const int g = lane >> 2, t = lane & 3;
    const int64_t r0 = row0 + g, r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < KB; ++j) {
        const int col = 8 * j + 2 * t;
        c[j][0] = (r0 < P && col < dim) ? __ldg(x + r0 * dim + col) : 0.f;
        c[j][1] = (r0 < P && col + 1 < dim) ? __ldg(x + r0 * dim + col + 1) : 0.f;
        c[j][2] = (r1 < P && col < dim) ? __ldg(x + r1 * dim + col) : 0.f;
        c[j][3] = (r1 < P && col + 1 < dim) ? __ldg(x + r1 * dim + col + 1) : 0.f;
    }
}

template <int KB>
__device__ __forceinline__ void load_rows_v2(const float* __restrict__ x, int64_t P, int dim, int64_t row0, int lane,
                                             float (&c)[KB][4]) {
    const int g = lane >> 2, t = lane & 3;
    const int64_t r0 = row0 + g, r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < KB; ++j) {
        const int col = 8 * j + 2 * t;
        const float2 a = (r0 < P && col < dim) ? __ldg(reinterpret_cast<const float2*>(x + r0 * dim + col)) : make_float2(0.f, 0.f);
        const float2 b = (r1 < P && col < dim) ? __ldg(reinterpret_cast<const float2*>(x + r1 * dim + col)) : make_float2(0.f, 0.f);
        c[j][0] = a.x; c[j][1] = a.y; c[j][2] = b.x; c[j][3] = b.y;
    }
}

template <int KB>
__device__ __forceinline__ void store_rows_v2(float* __restrict__ y, int64_t P, int dim, int64_t row0, int lane,
                                              const float (&c)[KB][4]) {
    const int g = lane >> 2, t = lane & 3;
    const int64_t r0 = row0 + g, r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < KB; ++j) {
        const int col = 8 * j + 2 * t;
        if (r0 < P && col < dim) *reinterpret_cast<float2*>(y + r0 * dim + col) = make_float2(c[j][0], c[j][1]);
        if (r1 < P && col < dim) *reinterpret_cast<float2*>(y + r1 * dim + col) = make_float2(c[j][2], c[j][3]);
    }
}

template <int KB>
__device__ __forceinline__ void store_rows(float* __restrict__ y, int64_t P, int dim, int64_t row0, int lane,
                                           const float (&c)[KB][4]) {
    const int g = lane >> 2, t = lane & 3;
    const int64_t r0 = row0 + g, r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < KB; ++j) {
        const int col = 8 * j + 2 * t;
        if (r0 < P && col < dim) y[r0 * dim + col] = c[j][0];
        if (r0 < P && col + 1 < dim) y[r0 * dim + col + 1] = c[j][1];
        if (r1 < P && col < dim) y[r1 * dim + col] = c[j][2];
        if (r1 < P && col + 1 < dim) y[r1 * dim + col + 1] = c[j][3];
    }
}

// ---- segmented input rows -----------------------------------------------------------------------------
__device__ __forceinline__ int seg_of(const MlpArgs& a, int col) {
    int s = 0;
    if (a.nseg > 1 && col >= a.seg[1].begin) s = 1;
    if (a.nseg > 2 && col >= a.seg[2].begin) s = 2;
    return s;
}

template <int KB>
__device__ __forceinline__ void load_rows_seg(const MlpArgs& a, int64_t row0, int lane, float (&c)[KB][4]) {
    const int g = lane >> 2, t = lane & 3;
    const int64_t r0 = row0 + g, r1 = r0 + 8;
    const bool v0 = r0 < a.P, v1 = r1 < a.P;
    if (a.lm_F > 0) {
        // level-major hash features [L][P][F]: element (r, c) at ((c / F) * P + r) * F + c % F.  For a fixed level
        // the 8 rows of a lane group are contiguous, so every request moves full sectors.
        const int F = a.lm_F;
        const float* src = a.seg[0].src;
#pragma unroll
        for (int j = 0; j < KB; ++j) {
            const int col = 8 * j + 2 * t;
            float x00 = 0.f, x01 = 0.f, x10 = 0.f, x11 = 0.f;
            if (col < a.in_dim) {
                if (F >= 2) {       // the pair (col, col+1) lies inside one level's F-vector
                    const int64_t lv = (int64_t)(col / F) * a.P;
                    const int f = col % F;
                    if (v0) { const float2 q = __ldg(reinterpret_cast<const float2*>(src + (lv + r0) * F + f)); x00 = q.x; x01 = q.y; }
                    if (v1) { const float2 q = __ldg(reinterpret_cast<const float2*>(src + (lv + r1) * F + f)); x10 = q.x; x11 = q.y; }
                } else {
                    const bool second = col + 1 < a.in_dim;
                    if (v0) { x00 = __ldg(src + (int64_t)col * a.P + r0); if (second) x01 = __ldg(src + (int64_t)(col + 1) * a.P + r0); }
                    if (v1) { x10 = __ldg(src + (int64_t)col * a.P + r1); if (second) x11 = __ldg(src + (int64_t)(col + 1) * a.P + r1); }
                }
            }
            c[j][0] = x00; c[j][1] = x01; c[j][2] = x10; c[j][3] = x11;
        }
        return;
    }
    if (a.nseg == 1 && a.seg[0].group == 1) {
        // common case: one per-point source (possibly a strided column window)
        const float* p0 = a.seg[0].src + r0 * a.seg[0].stride + a.seg[0].col0;
        const float* p1 = a.seg[0].src + r1 * a.seg[0].stride + a.seg[0].col0;
        if (a.vec2_x) {
#pragma unroll
            for (int j = 0; j < KB; ++j) {
                const int col = 8 * j + 2 * t;          // in_dim is even: the pair is live or dead as a whole
                const bool live = col < a.in_dim;
                const float2 x0 = (live && v0) ? __ldg(reinterpret_cast<const float2*>(p0 + col)) : make_float2(0.f, 0.f);
                const float2 x1 = (live && v1) ? __ldg(reinterpret_cast<const float2*>(p1 + col)) : make_float2(0.f, 0.f);
                c[j][0] = x0.x; c[j][1] = x0.y; c[j][2] = x1.x; c[j][3] = x1.y;
            }
            return;
        }
#pragma unroll
        for (int j = 0; j < KB; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = 8 * j + 2 * t + h;
                const bool live = col < a.in_dim;
                c[j][h] = (live && v0) ? __ldg(p0 + col) : 0.f;
                c[j][2 + h] = (live && v1) ? __ldg(p1 + col) : 0.f;
            }
        }
        return;
    }
    // per-segment base addresses of the two rows.  A warp's 16 rows (row0 is a multiple of 16, groups are multiples
    // of 16) always share one group index, so the division is done once per warp in 32 bits.
    const float* b0[PS_MLP_MAX_SEGMENTS];
    const float* b1[PS_MLP_MAX_SEGMENTS];
#pragma unroll
    for (int s = 0; s < PS_MLP_MAX_SEGMENTS; ++s) {
        if (s < a.nseg) {
            const RowSeg& sg = a.seg[s];
            if (sg.group == 1) {
                b0[s] = sg.src + (v0 ? r0 : 0) * sg.stride + sg.col0 - sg.begin;
                b1[s] = sg.src + (v1 ? r1 : 0) * sg.stride + sg.col0 - sg.begin;
            } else {
                const int64_t q = (int64_t)((uint32_t)row0 / (uint32_t)sg.group);
                b0[s] = b1[s] = sg.src + q * sg.stride + sg.col0 - sg.begin;
            }
        } else {
            b0[s] = b1[s] = nullptr;
        }
    }
#pragma unroll
    for (int j = 0; j < KB; ++j) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int col = 8 * j + 2 * t + h;
            float x0 = 0.f, x1 = 0.f;
            if (col < a.in_dim) {
                const int s = seg_of(a, col);
                const float* p0 = s == 0 ? b0[0] : (s == 1 ? b0[1] : b0[2]);
                const float* p1 = s == 0 ? b1[0] : (s == 1 ? b1[1] : b1[2]);
                if (v0) x0 = __ldg(p0 + col);
                if (v1) x1 = __ldg(p1 + col);
            }
            c[j][h] = x0;
            c[j][2 + h] = x1;
        }
    }
}

// input-gradient rows: plain strided store for per-point segments, warp-reduced atomicAdd for per-group segments
template <int KB>
__device__ __forceinline__ void store_dx_seg(const MlpArgs& a, int64_t row0, int lane, const float (&d)[KB][4]) {
    const int g = lane >> 2, t = lane & 3;
    const int64_t r0 = row0 + g, r1 = r0 + 8;
    const bool v0 = r0 < a.P, v1 = r1 < a.P;
    if (a.lm_F > 0) {
        const int F = a.lm_F;
        float* dst = a.seg[0].dst;
#pragma unroll
        for (int j = 0; j < KB; ++j) {
            const int col = 8 * j + 2 * t;
            if (col < a.in_dim) {
                if (F >= 2) {
                    const int64_t lv = (int64_t)(col / F) * a.P;
                    const int f = col % F;
                    if (v0) *reinterpret_cast<float2*>(dst + (lv + r0) * F + f) = make_float2(d[j][0], d[j][1]);
                    if (v1) *reinterpret_cast<float2*>(dst + (lv + r1) * F + f) = make_float2(d[j][2], d[j][3]);
                } else {
                    const bool second = col + 1 < a.in_dim;
                    if (v0) { dst[(int64_t)col * a.P + r0] = d[j][0]; if (second) dst[(int64_t)(col + 1) * a.P + r0] = d[j][1]; }
                    if (v1) { dst[(int64_t)col * a.P + r1] = d[j][2]; if (second) dst[(int64_t)(col + 1) * a.P + r1] = d[j][3]; }
                }
            }
        }
        return;
    }
    if (a.nseg == 1 && a.seg[0].group == 1) {
        float* p0 = a.seg[0].dst + r0 * a.seg[0].stride + a.seg[0].col0;
        float* p1 = a.seg[0].dst + r1 * a.seg[0].stride + a.seg[0].col0;
        if (a.vec2_x) {
#pragma unroll
            for (int j = 0; j < KB; ++j) {
                const int col = 8 * j + 2 * t;
                if (col < a.in_dim) {
                    if (v0) *reinterpret_cast<float2*>(p0 + col) = make_float2(d[j][0], d[j][1]);
                    if (v1) *reinterpret_cast<float2*>(p1 + col) = make_float2(d[j][2], d[j][3]);
                }
            }
            return;
        }
#pragma unroll
        for (int j = 0; j < KB; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = 8 * j + 2 * t + h;
                if (col < a.in_dim) {
                    if (v0) p0[col] = d[j][h];
                    if (v1) p1[col] = d[j][2 + h];
                }
            }
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < KB; ++j) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int col = 8 * j + 2 * t + h;
            const bool live = col < a.in_dim;
            const int s = live ? seg_of(a, col) : 0;
            const RowSeg& sg = a.seg[s];
            if (a.any_group_dst) {
                // column sum over the warp's 16 rows (all lanes take part; only group segments use it)
                float sum = (v0 ? d[j][h] : 0.f) + (v1 ? d[j][2 + h] : 0.f);
                sum += __shfl_xor_sync(0xffffffffu, sum, 4);
                sum += __shfl_xor_sync(0xffffffffu, sum, 8);
                sum += __shfl_xor_sync(0xffffffffu, sum, 16);
                if (live && sg.group > 1 && sg.dst && g == 0 && row0 < a.P)
                    atomicAdd(sg.dst + (int64_t)((uint32_t)row0 / (uint32_t)sg.group) * sg.stride + sg.col0 + (col - sg.begin), sum);
            }
            if (live && sg.group == 1 && sg.dst) {
                float* base = sg.dst + sg.col0 + (col - sg.begin);
                if (v0) base[r0 * sg.stride] = d[j][h];
                if (v1) base[r1 * sg.stride] = d[j][2 + h];
            }
        }
    }
}

template <int KB>
__device__ __forceinline__ uint32_t relu_inplace(float (&c)[KB][4]) {
    uint32_t mask = 0;
#pragma unroll
    for (int j = 0; j < KB; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (c[j][q] > 0.f)
                mask |= 1u << (4 * j + q);
            else
                c[j][q] = 0.f;
        }
    return mask;
}
template <int KB>
__device__ __forceinline__ void apply_mask(float (&c)[KB][4], uint32_t mask) {
#pragma unroll
    for (int j = 0; j < KB; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (!((mask >> (4 * j + q)) & 1u)) c[j][q] = 0.f;
}

// shared-memory carve-up shared by forward and backward
template <class S, int PREC>
struct Smem {
    using E = typename Elem<PREC>::type;
    static constexpr size_t w0 = (size_t)S::N0 * stride_of<PREC>(S::K0);
    static constexpr size_t wmid = (size_t)S::H * stride_of<PREC>(S::H);
    static constexpr size_t wlast = S::NHID > 0 ? (size_t)S::NOUT * stride_of<PREC>(S::H) : 0;
    static constexpr size_t w_elems = w0 + S::NMID * wmid + wlast;
    static constexpr size_t bias_floats = S::N0 + S::NMID * S::H + (S::NHID > 0 ? S::NOUT : 0);
    static constexpr size_t fwd_bytes = ((w_elems * sizeof(E) + 15) / 16) * 16 + bias_floats * sizeof(float);
    // backward adds the staged activations (inputs of every layer) and one dZ tile
    static constexpr size_t act0 = (size_t)kTile * stride_of<PREC>(S::K0);
    static constexpr size_t acth = (size_t)kTile * stride_of<PREC>(S::H);
    static constexpr int maxN = S::NOUT > S::H ? S::NOUT : S::H;
    static constexpr size_t dz = (size_t)kTile * stride_of<PREC>(maxN);
    static constexpr size_t bwd_bytes = ((fwd_bytes + 15) / 16) * 16 + (act0 + S::NHID * acth + dz) * sizeof(E);
};

template <class S, int PREC>
struct Weights {
    using E = typename Elem<PREC>::type;
    E* w0;
    E* wmid;   // NMID consecutive [H][stride(H)] blocks
    E* wlast;
    float* b0;
    float* bmid;
    float* blast;
    unsigned char* end;

    __device__ __forceinline__ void carve(unsigned char* base) {
        using SM = Smem<S, PREC>;
        w0 = reinterpret_cast<E*>(base);
        wmid = w0 + SM::w0;
        wlast = wmid + S::NMID * SM::wmid;
        float* bias = reinterpret_cast<float*>(base + ((SM::w_elems * sizeof(E) + 15) / 16) * 16);
        b0 = bias;
        bmid = b0 + S::N0;
        blast = bmid + S::NMID * S::H;
        end = base + ((SM::fwd_bytes + 15) / 16) * 16;
    }
    __device__ __forceinline__ void load(const MlpArgs& a, int tid) {
        if constexpr (S::NHID == 0) {
            load_weights<S::K0, S::NOUT, PREC>(a.W[0], a.b[0], a.out_dim, a.in_dim, w0, b0, tid, kThreads);
        } else {
            load_weights<S::K0, S::H, PREC>(a.W[0], a.b[0], S::H, a.in_dim, w0, b0, tid, kThreads);
#pragma unroll
            for (int m = 0; m < S::NMID; ++m)
                load_weights<S::H, S::H, PREC>(a.W[1 + m], a.b[1 + m], S::H, S::H, wmid + m * Smem<S, PREC>::wmid,
                                               bmid + m * S::H, tid, kThreads);
            load_weights<S::H, S::NOUT, PREC>(a.W[S::NHID], a.b[S::NHID], a.out_dim, S::H, wlast, blast, tid, kThreads);
        }
    }
};

// ------------------------------------------------------------------------------------------------
template <class S, int PREC>
__global__ void __launch_bounds__(kThreads) mlp_fwd_kernel(MlpArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Weights<S, PREC> w;
    w.carve(smem_raw);
    w.load(a, threadIdx.x);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ntiles = (a.P + kTile - 1) / kTile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * kTile + warp * 16;
        if (row0 >= a.P) continue;
        float in[S::K0 / 8][4];
        load_rows_seg<S::K0 / 8>(a, row0, lane, in);
        float out[S::NOUT / 8][4];
        if constexpr (S::NHID == 0) {
            layer_forward<S::K0, S::NOUT, PREC>(in, w.w0, w.b0, out, lane);
        } else {
            float h[S::H / 8][4];
            layer_forward<S::K0, S::H, PREC>(in, w.w0, w.b0, h, lane);
            relu_inplace<S::H / 8>(h);
#pragma unroll
            for (int m = 0; m < S::NMID; ++m) {
                float h2[S::H / 8][4];
                layer_forward<S::H, S::H, PREC>(h, w.wmid + m * Smem<S, PREC>::wmid, w.bmid + m * S::H, h2, lane);
                relu_inplace<S::H / 8>(h2);
#pragma unroll
                for (int j = 0; j < S::H / 8; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) h[j][q] = h2[j][q];
            }
            layer_forward<S::H, S::NOUT, PREC>(h, w.wlast, w.blast, out, lane);
        }
        if (a.out_act == PS_ACT_SIGMOID) {
#pragma unroll
            for (int j = 0; j < S::NOUT / 8; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) out[j][q] = sigmoidf(out[j][q]);
        } else if (a.out_act == PS_ACT_RELU) {
            relu_inplace<S::NOUT / 8>(out);
        }
        if (a.density_out && (lane & 3) == 0) {
            // column 0 lives in block 0 of the lanes with t == 0: q = 0 (row g) and q = 2 (row g + 8)
            const int64_t r0 = row0 + (lane >> 2), r1 = r0 + 8;
            if (r0 < a.P) a.density_out[r0] = expf(out[0][0]) * (a.sel ? (float)a.sel[r0] : 1.f);
            if (r1 < a.P) a.density_out[r1] = expf(out[0][2]) * (a.sel ? (float)a.sel[r1] : 1.f);
        }
        if (a.y) {
            if (a.vec2_y)
                store_rows_v2<S::NOUT / 8>(a.y, a.P, a.out_dim, row0, lane, out);
            else
                store_rows<S::NOUT / 8>(a.y, a.P, a.out_dim, row0, lane, out);
        }
    }
}

// Accumulate one layer's weight/bias gradient from the staged tiles.
//   dZs [kTile][stride(N)] (this layer's pre-activation gradient), Acts [kTile][stride(K)] (this layer's input)
template <int K, int N, int PREC, int MAXB>
__device__ __forceinline__ void accumulate_dw(const typename Elem<PREC>::type* dZs, const typename Elem<PREC>::type* Acts,
                                              float (&acc)[MAXB][2][4], float& db, int warp, int lane, int tid) {
    constexpr int KBLK = K / 16, NBLK = (N / 16) * KBLK;
#pragma unroll
    for (int i = 0; i < MAXB; ++i) {
        const int b = warp + i * kWarps;
        if (b < NBLK) {
            const int n0 = (b / KBLK) * 16, f0 = (b % KBLK) * 16;
            dw_block<kTile, PREC>(dZs, stride_of<PREC>(N), n0, Acts, stride_of<PREC>(K), f0, acc[i], lane);
        }
    }
    // bias gradient = column sums of dZ: the tile's rows are split over R = kThreads / N thread groups so the
    // dependent-add chain is kTile / R long; partial sums stay in one persistent register per thread
    {
        constexpr int SZ = stride_of<PREC>(N);
        constexpr int R = kThreads / N, ROWS = (kTile + R - 1) / R;
        if (tid < R * N) {
            const int col = tid % N, r0 = (tid / N) * ROWS;
            float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
            for (int p = 0; p < ROWS; p += 2) {
                const int pa = r0 + p, pb = r0 + p + 1;
                if constexpr (PREC == kBF16) {
                    if (pa < kTile) s0 += __bfloat162float(dZs[pa * SZ + col]);
                    if (pb < kTile && p + 1 < ROWS) s1 += __bfloat162float(dZs[pb * SZ + col]);
                } else {
                    if (pa < kTile) s0 += dZs[pa * SZ + col];
                    if (pb < kTile && p + 1 < ROWS) s1 += dZs[pb * SZ + col];
                }
            }
            db += s0 + s1;
        }
    }
}

template <int K, int N, int MAXB>
__device__ __forceinline__ void flush_dw(float* __restrict__ dW, float* __restrict__ dbg, int n_real, int k_real,
                                         const float (&acc)[MAXB][2][4], float db, int warp, int lane, int tid) {
    constexpr int KBLK = K / 16, NBLK = (N / 16) * KBLK;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < MAXB; ++i) {
        const int b = warp + i * kWarps;
        if (b < NBLK) {
            const int n0 = (b / KBLK) * 16, f0 = (b % KBLK) * 16;
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int n = n0 + g + (q >> 1) * 8, k = f0 + 8 * nb + 2 * t + (q & 1);
                    if (n < n_real && k < k_real && acc[i][nb][q] != 0.f) atomicAdd(dW + (size_t)n * k_real + k, acc[i][nb][q]);
                }
        }
    }
    if (dbg && tid < (kThreads / N) * N && (tid % N) < n_real) atomicAdd(dbg + (tid % N), db);
}

template <int K, int N>
constexpr int max_blocks() {
    return ((N / 16) * (K / 16) + kWarps - 1) / kWarps;
}

template <class S, int PREC>
__global__ void __launch_bounds__(kThreads, PREC == kBF16 ? 2 : 1) mlp_bwd_kernel(MlpArgs a) {
    using E = typename Elem<PREC>::type;
    using SM = Smem<S, PREC>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Weights<S, PREC> w;
    w.carve(smem_raw);
    E* acts0 = reinterpret_cast<E*>(w.end);
    E* actsh = acts0 + SM::act0;              // NHID tiles [kTile][stride(H)]
    E* dZs = actsh + S::NHID * SM::acth;
    w.load(a, threadIdx.x);
    __syncthreads();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    constexpr int MB0 = max_blocks<S::K0, S::N0>();
    constexpr int MBM = max_blocks<S::H, S::H>();
    constexpr int MBL = S::NHID > 0 ? max_blocks<S::H, S::NOUT>() : 1;
    float acc0[MB0][2][4] = {};
    float accm[S::NMID > 0 ? S::NMID : 1][MBM][2][4] = {};
    float accl[MBL][2][4] = {};
    float db0 = 0.f, dbl = 0.f, dbm[S::NMID > 0 ? S::NMID : 1] = {};

    const int64_t ntiles = (a.P + kTile - 1) / kTile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * kTile + warp * 16;
        const int srow = warp * 16;
        // ---- recompute the forward, staging every layer's input --------------------------------
        float in[S::K0 / 8][4];
        load_rows_seg<S::K0 / 8>(a, row0, lane, in);
        stage_rows<S::K0, PREC>(in, acts0, srow, lane);
        float dz[S::NOUT / 8][4];
        uint32_t mask[S::NHID > 0 ? S::NHID : 1] = {};
        {
            float out[S::NOUT / 8][4];
            if constexpr (S::NHID == 0) {
                layer_forward<S::K0, S::NOUT, PREC>(in, w.w0, w.b0, out, lane);
            } else {
                float h[S::H / 8][4];
                layer_forward<S::K0, S::H, PREC>(in, w.w0, w.b0, h, lane);
                mask[0] = relu_inplace<S::H / 8>(h);
                stage_rows<S::H, PREC>(h, actsh, srow, lane);
#pragma unroll
                for (int m = 0; m < S::NMID; ++m) {
                    float h2[S::H / 8][4];
                    layer_forward<S::H, S::H, PREC>(h, w.wmid + m * SM::wmid, w.bmid + m * S::H, h2, lane);
                    mask[m + 1] = relu_inplace<S::H / 8>(h2);
                    stage_rows<S::H, PREC>(h2, actsh + (m + 1) * SM::acth, srow, lane);
#pragma unroll
                    for (int j = 0; j < S::H / 8; ++j)
#pragma unroll
                        for (int q = 0; q < 4; ++q) h[j][q] = h2[j][q];
                }
                layer_forward<S::H, S::NOUT, PREC>(h, w.wlast, w.blast, out, lane);
            }
            // ---- gradient w.r.t. the last pre-activation ----------------------------------------
            if (a.dy) {
                if (a.vec2_y)
                    load_rows_v2<S::NOUT / 8>(a.dy, a.P, a.out_dim, row0, lane, dz);
                else
                    load_rows<S::NOUT / 8>(a.dy, a.P, a.out_dim, row0, lane, dz);
            } else {
#pragma unroll
                for (int j = 0; j < S::NOUT / 8; ++j) dz[j][0] = dz[j][1] = dz[j][2] = dz[j][3] = 0.f;
            }
            if (a.d_density && (lane & 3) == 0) {
                // gradient of column 0 comes from the density: d raw = d density * sel * exp(clamp(raw, -15, 15))
                const int64_t r0 = row0 + (lane >> 2), r1 = r0 + 8;
                dz[0][0] = r0 < a.P ? a.d_density[r0] * (a.sel ? (float)a.sel[r0] : 1.f) *
                                          expf(fminf(fmaxf(out[0][0], -15.f), 15.f)) : 0.f;
                dz[0][2] = r1 < a.P ? a.d_density[r1] * (a.sel ? (float)a.sel[r1] : 1.f) *
                                          expf(fminf(fmaxf(out[0][2], -15.f), 15.f)) : 0.f;
            }
            if (a.out_act == PS_ACT_SIGMOID) {
#pragma unroll
                for (int j = 0; j < S::NOUT / 8; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float y = sigmoidf(out[j][q]);
                        dz[j][q] *= y * (1.f - y);
                    }
            } else if (a.out_act == PS_ACT_RELU) {
#pragma unroll
                for (int j = 0; j < S::NOUT / 8; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (!(out[j][q] > 0.f)) dz[j][q] = 0.f;
            }
        }
        // ---- walk the layers backwards ----------------------------------------------------------
        if constexpr (S::NHID == 0) {
            stage_rows<S::NOUT, PREC>(dz, dZs, srow, lane);
            __syncthreads();
            accumulate_dw<S::K0, S::NOUT, PREC, MB0>(dZs, acts0, acc0, db0, warp, lane, tid);
            if (a.want_dx) {
                float da[S::K0 / 8][4];
                layer_backward_input<S::K0, S::NOUT, PREC>(dz, w.w0, da, lane);
                store_dx_seg<S::K0 / 8>(a, row0, lane, da);
            }
            __syncthreads();
        } else {
            float dzh[S::H / 8][4];
            stage_rows<S::NOUT, PREC>(dz, dZs, srow, lane);
            __syncthreads();
            accumulate_dw<S::H, S::NOUT, PREC, MBL>(dZs, actsh + (S::NHID - 1) * SM::acth, accl, dbl, warp, lane, tid);
            layer_backward_input<S::H, S::NOUT, PREC>(dz, w.wlast, dzh, lane);
            apply_mask<S::H / 8>(dzh, mask[S::NHID - 1]);
            __syncthreads();
#pragma unroll
            for (int m = S::NMID - 1; m >= 0; --m) {
                stage_rows<S::H, PREC>(dzh, dZs, srow, lane);
                __syncthreads();
                accumulate_dw<S::H, S::H, PREC, MBM>(dZs, actsh + m * SM::acth, accm[m], dbm[m], warp, lane, tid);
                float da[S::H / 8][4];
                layer_backward_input<S::H, S::H, PREC>(dzh, w.wmid + m * SM::wmid, da, lane);
                apply_mask<S::H / 8>(da, mask[m]);
#pragma unroll
                for (int j = 0; j < S::H / 8; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) dzh[j][q] = da[j][q];
                __syncthreads();
            }
            stage_rows<S::H, PREC>(dzh, dZs, srow, lane);
            __syncthreads();
            accumulate_dw<S::K0, S::H, PREC, MB0>(dZs, acts0, acc0, db0, warp, lane, tid);
            if (a.want_dx) {
                float da[S::K0 / 8][4];
                layer_backward_input<S::K0, S::H, PREC>(dzh, w.w0, da, lane);
                store_dx_seg<S::K0 / 8>(a, row0, lane, da);
            }
            __syncthreads();
        }
    }
    // ---- flush the per-CTA partial weight gradients ------------------------------------------------
    if constexpr (S::NHID == 0) {
        flush_dw<S::K0, S::NOUT, MB0>(a.dW[0], a.db[0], a.out_dim, a.in_dim, acc0, db0, warp, lane, tid);
    } else {
        flush_dw<S::K0, S::H, MB0>(a.dW[0], a.db[0], S::H, a.in_dim, acc0, db0, warp, lane, tid);
#pragma unroll
        for (int m = 0; m < S::NMID; ++m)
            flush_dw<S::H, S::H, MBM>(a.dW[1 + m], a.db[1 + m], S::H, S::H, accm[m], dbm[m], warp, lane, tid);
        flush_dw<S::H, S::NOUT, MBL>(a.dW[S::NHID], a.db[S::NHID], a.out_dim, S::H, accl, dbl, warp, lane, tid);
    }
}

// host-side launchers -------------------------------------------------------------------------------
template <class S, int PREC>
int launch_fwd(const MlpArgs& a, cudaStream_t stream) {
    constexpr size_t smem = Smem<S, PREC>::fwd_bytes;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(mlp_fwd_kernel<S, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess) {
            set_error("mlp_fwd: cannot reserve %zu bytes of shared memory", smem);
            return 2;
        }
        configured = true;
    }
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, mlp_fwd_kernel<S, PREC>, kThreads, smem) !=
                cudaSuccess || ctas_per_sm < 1)
            ctas_per_sm = 1;
    }
    const int64_t ntiles = (a.P + kTile - 1) / kTile;
    const int grid = (int)(ntiles < (int64_t)kNumSMs * ctas_per_sm ? ntiles : (int64_t)kNumSMs * ctas_per_sm);
    mlp_fwd_kernel<S, PREC><<<grid, kThreads, smem, stream>>>(a);
    return check_launch("mlp_fwd");
}

template <class S, int PREC>
int launch_bwd(const MlpArgs& a, cudaStream_t stream) {
    constexpr size_t smem = Smem<S, PREC>::bwd_bytes;
    static_assert(smem <= 227 * 1024, "backward tile does not fit in shared memory");
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(mlp_bwd_kernel<S, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess) {
            set_error("mlp_bwd: cannot reserve %zu bytes of shared memory", smem);
            return 2;
        }
        configured = true;
    }
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, mlp_bwd_kernel<S, PREC>, kThreads, smem) !=
                cudaSuccess || ctas_per_sm < 1)
            ctas_per_sm = 1;
    }
    const int64_t ntiles = (a.P + kTile - 1) / kTile;
    const int grid = (int)(ntiles < (int64_t)kNumSMs * ctas_per_sm ? ntiles : (int64_t)kNumSMs * ctas_per_sm);
    mlp_bwd_kernel<S, PREC><<<grid, kThreads, smem, stream>>>(a);
    return check_launch("mlp_bwd");
}

}  // namespace mma
}  // namespace ps
