// Shared definitions of the fused field-level kernels (field_tc5_fwd.cu / field_tc5_bwd.cu):
// the PreSight field of fields/PreSight/ingp_field.py:118-161 with semantics, evaluated per 128-point tile on
// tcgen05 tensor cores with one thread per point.
//
//   base  : feat[K0] -> 64 relu -> 80 = [raw density | geo 15 | semantic input 64]           (ingp_field.py:130-138)
//   sem   : h[16:80] -> 64 relu -> 64 relu -> 64                                              (:142-151)
//   rgb   : [sh16(dir) | geo 15 | app A<=16] -> 64 relu -> 64 relu -> 3 sigmoid               (:153-161)
//
// Column bookkeeping of the colour head: the kernel feeds [sh 16 | h[0:16] | app 16] (48 columns) and the staged
// weight has a zero column where h[0] (the raw density) sits, so no shifted copy of h is ever made:
//   staged column 0..15 = reference column 0..15 (sh), 16 = zero, 17..31 = reference 16..30 (geo),
//   32..32+A-1 = reference 31..31+A-1 (appearance), rest zero.
#pragma once
#include "composite.cuh"
#include "position.cuh"
#include "tc5.cuh"

namespace ps {
namespace ftc5 {

using namespace tc5;

constexpr int kHid = 64;      // hidden width of all three networks
constexpr int kBaseOut = 80;  // 1 + 15 + 64
constexpr int kGeo = 15;
constexpr int kSem = 64;      // semantic input and output width
constexpr int kRgbIn = 48;    // staged colour-head input width
constexpr int kRgbOut = 16;   // padded (3 real)

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// layer indices into FieldNet::W / B / dW / dB
enum { B0 = 0, B1 = 1, S0 = 2, S1 = 3, S2 = 4, R0 = 5, R1 = 6, R2 = 7, kLayers = 8 };

struct FieldNet {
    const float* W[kLayers];
    const float* B[kLayers];
    float* dW[kLayers];   // backward only (accumulated)
    float* dB[kLayers];
    int in_dim;           // L * F (real)
    int app_dim;          // A <= 16
};

struct FieldArgs {
    FieldNet net;
    const float* feat;     // level-major hash features [L][P][F]
    float* dfeat;          // backward: gradient, same layout (written)
    int L, F;
    const uint8_t* sel;    // [P]
    const float* eu;       // [N, S+1] euclidean bin edges
    const float* dirs;     // [N, 3]
    const float* app;      // [N, A] (nullable when A == 0)
    float* dapp;           // backward: [N, A] accumulated (nullable)
    int64_t N;
    int S;
    float thr;
    // forward outputs
    float* weights;        // [N, S]
    float* rgb_out;        // [N, 3]
    float* acc;            // [N]
    float* dexp;           // [N]
    float* dthr;           // [N]
    float* sem_out;        // [N, 64]
    float* tminmax;        // [2]
    // backward inputs (acc / dexp above are then inputs)
    const float* d_w;      // [N, S] nullable
    const float* d_rgb;    // [N, 3] nullable
    const float* d_acc;    // [N] nullable
    const float* d_dexp;   // [N] nullable
    const float* d_sem;    // [N, 64] nullable
};

// shared-memory weight tiles (bf16 chunk-major, rows = out features).  Every layer also has a 16-column "bias tile"
// [N x 16]: column 0 = bf16(b), column 1 = bf16(b - bf16(b)) — one extra K step of the forward GEMM against a constant
// A operand whose first two columns are 1 adds the bias (to ~2^-17 relative) on the tensor core, so no epilogue touches
// it.  Only the tile's first 8-column chunk is stored per layer; the second (all zeros) is ONE chunk shared by all layers —
// the descriptor's distance between K chunks simply points at it (6.4 KB less shared memory).  `ones`: that constant A operand as one 128-byte core-matrix block {1,1,0,..} x 8 rows + a zero block (the
// descriptor's row-group stride is 0, so all 128 rows alias the block).
template <int K0>
struct WLayout {
    static constexpr uint32_t b0 = 0;
    static constexpr uint32_t b1 = b0 + cm_bytes(kHid, K0);
    static constexpr uint32_t s0 = b1 + cm_bytes(kBaseOut, kHid);
    static constexpr uint32_t s1 = s0 + cm_bytes(kHid, kSem);
    static constexpr uint32_t s2 = s1 + cm_bytes(kHid, kHid);
    static constexpr uint32_t r0 = s2 + cm_bytes(kSem, kHid);
    static constexpr uint32_t r1 = r0 + cm_bytes(kHid, kRgbIn);
    static constexpr uint32_t r2 = r1 + cm_bytes(kHid, kHid);
    static constexpr uint32_t bt0 = r2 + cm_bytes(kRgbOut, kHid);            // bias chunks [N x 8], layer order B0..R2
    static constexpr uint32_t btz = bt0 + 480 * 16;                           // the zero chunk all bias tiles share (80 rows)
    static constexpr uint32_t ones = btz + kBaseOut * 16;                     // 256 B
    static constexpr uint32_t onehot = ones + 256;                            // kLayers x 256 B (backward: bias gradients)
    static constexpr uint32_t end = onehot + kLayers * 256;
    __host__ __device__ static constexpr int rows(int l) {
        return l == B1 ? kBaseOut : (l == R2 ? kRgbOut : kHid);
    }
    __host__ __device__ static constexpr uint32_t bt(int l) {
        uint32_t off = bt0;
        for (int i = 0; i < l; ++i) off += rows(i) * 16;
        return off;
    }
};

// Bias gradients on the tensor core (backward): dB_l = dZ_l^T 1 is one more reduction over the tile's 128 points, against
// a constant B operand [K = 128 points x N = 16] that is 1 in column kBiasCol0 + l and 0 elsewhere — stored once as a
// 128-byte core-matrix block pair (8 K-rows x 16 columns) that every 8-row group of the reduction re-reads (the
// descriptor's K stride is 0).  All layers accumulate into ONE 16-column accumulator, column kBiasCol0 + l = layer l
// (rows = out features); the colour head's last layer (3 real outputs) keeps its bias gradient in warp sums.
constexpr int kBiasCol0 = 3;
static_assert(kBiasCol0 + kLayers - 1 <= 16, "one-hot columns");

// D[dz column][kBiasCol0 + l] += sum over the 128 points of dZ[p][column]   (dz_addr: first column chunk of the dZ tile)
__device__ __forceinline__ void gemm_bias_grad(uint32_t tmem_d, uint32_t dz_addr, uint32_t onehot_addr) {
    const uint32_t idesc = make_idesc(16, 1, 1);
#pragma unroll
    for (int kk = 0; kk < kRows / 16; ++kk)
        umma_bf16(tmem_d, make_desc(dz_addr + kk * 256, 128, kRows * 16), make_desc(onehot_addr, 0, 128), idesc, 1u);
}

// D[128 x N] = 1 * bias^T : the first K step of every forward GEMM (accumulate = false)
__device__ __forceinline__ void gemm_bias(uint32_t tmem_d, uint32_t ones_addr, uint32_t bias_tile, uint32_t zero_chunk, int N) {
    umma_bf16(tmem_d, make_desc(ones_addr, 128, 0), make_desc(bias_tile, zero_chunk - bias_tile, 128), make_idesc(N, 0, 0), 0u);
}

template <int K0>
__device__ __forceinline__ void load_all_weights(const FieldNet& net, unsigned char* wbase, int tid, int nthreads) {
    using WL = WLayout<K0>;
    __shared__ int kmap[kRgbIn];
    if (tid < kRgbIn) {
        int src = -1;
        if (tid < 16) src = tid;
        else if (tid >= 17 && tid < 32) src = tid - 1;
        else if (tid >= 32 && tid - 32 < net.app_dim) src = tid - 1;
        kmap[tid] = src;
    }
    __syncthreads();
    load_weight_cm(net.W[B0], kHid, net.in_dim, kHid, K0, wbase + WL::b0, nullptr, tid, nthreads);
    load_weight_cm(net.W[B1], kBaseOut, kHid, kBaseOut, kHid, wbase + WL::b1, nullptr, tid, nthreads);
    load_weight_cm(net.W[S0], kHid, kSem, kHid, kSem, wbase + WL::s0, nullptr, tid, nthreads);
    load_weight_cm(net.W[S1], kHid, kHid, kHid, kHid, wbase + WL::s1, nullptr, tid, nthreads);
    load_weight_cm(net.W[S2], kSem, kHid, kSem, kHid, wbase + WL::s2, nullptr, tid, nthreads);
    load_weight_cm(net.W[R0], kHid, 16 + kGeo + net.app_dim, kHid, kRgbIn, wbase + WL::r0, kmap, tid, nthreads);
    load_weight_cm(net.W[R1], kHid, kHid, kHid, kHid, wbase + WL::r1, nullptr, tid, nthreads);
    load_weight_cm(net.W[R2], 3, kHid, kRgbOut, kHid, wbase + WL::r2, nullptr, tid, nthreads);
    const int nreal[kLayers] = {kHid, kBaseOut, kHid, kHid, kSem, kHid, kHid, 3};
    for (int l = 0; l < kLayers; ++l) {
        const int N = WL::rows(l);
        unsigned char* bt = wbase + WL::bt(l);
        for (int i = tid; i < N * 8; i += nthreads) {
            const int n = i >> 3, k = i & 7;
            float v = 0.f;
            if (k < 2 && n < nreal[l] && net.B[l]) {
                const float b = __ldg(net.B[l] + n);
                const float hi = __bfloat162float(__float2bfloat16_rn(b));
                v = k == 0 ? hi : b - hi;
            }
            *reinterpret_cast<__nv_bfloat16*>(bt + cm_off(N, n, k)) = __float2bfloat16_rn(v);
        }
    }
    for (int i = tid; i < kBaseOut * 8; i += nthreads) reinterpret_cast<__nv_bfloat16*>(wbase + WL::btz)[i] = __float2bfloat16_rn(0.f);
    // constant blocks: element e of a 128-byte block = row e / 8, column e % 8
    for (int i = tid; i < 128; i += nthreads) {
        reinterpret_cast<__nv_bfloat16*>(wbase + WL::ones)[i] = __float2bfloat16_rn((i < 64 && (i & 7) < 2) ? 1.f : 0.f);
    }
    // one-hot blocks of the bias-gradient GEMMs: layer l has ones in column kBiasCol0 + l of its [8 x 16] block pair
    for (int i = tid; i < kLayers * 128; i += nthreads) {
        const int l = i >> 7, e = i & 127;
        const int col = (e >> 6) * 8 + (e & 7);
        reinterpret_cast<__nv_bfloat16*>(wbase + WL::onehot)[i] = __float2bfloat16_rn(col == kBiasCol0 + l ? 1.f : 0.f);
    }
}

// Per-point / per-ray inputs of one row, loaded into registers in ONE batch (all global loads are issued before the
// first is consumed: staged separately, each group of loads would expose its own global-memory latency).
template <int K0>
struct RowInputs {
    float feat[K0];
    float dir[3];
    float app[16];
    float t0, t1, selv, gw;
};

template <int K0>
__device__ __forceinline__ void load_row_inputs(const FieldArgs& a, int64_t P, int64_t ray, int s, bool valid,
                                                bool want_feat, bool want_ray, RowInputs<K0>& in) {
#pragma unroll
    for (int c = 0; c < K0; ++c) in.feat[c] = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) in.app[c] = 0.f;
    in.dir[0] = in.dir[1] = in.dir[2] = 0.f;
    in.t0 = in.t1 = in.selv = in.gw = 0.f;
    if (!valid) return;
    const int64_t p = ray * a.S + s;
    if (want_feat) {
        if (a.F == 2) {
#pragma unroll
            for (int l = 0; l < K0 / 2; ++l)
                if (l < a.L) {
                    const float2 q = __ldg(reinterpret_cast<const float2*>(a.feat + ((int64_t)l * P + p) * 2));
                    in.feat[2 * l] = q.x;
                    in.feat[2 * l + 1] = q.y;
                }
        } else {  // F == 4
#pragma unroll
            for (int l = 0; l < K0 / 4; ++l)
                if (l < a.L) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(a.feat + ((int64_t)l * P + p) * 4));
                    in.feat[4 * l] = q.x; in.feat[4 * l + 1] = q.y; in.feat[4 * l + 2] = q.z; in.feat[4 * l + 3] = q.w;
                }
        }
    }
    if (want_ray) {
        in.dir[0] = __ldg(a.dirs + 3 * ray);
        in.dir[1] = __ldg(a.dirs + 3 * ray + 1);
        in.dir[2] = __ldg(a.dirs + 3 * ray + 2);
        if (a.app)
#pragma unroll
            for (int c = 0; c < 16; ++c)
                if (c < a.net.app_dim) in.app[c] = __ldg(a.app + ray * a.net.app_dim + c);
    }
    in.t0 = __ldg(a.eu + ray * (a.S + 1) + s);
    in.t1 = __ldg(a.eu + ray * (a.S + 1) + s + 1);
    in.selv = a.sel ? (float)a.sel[p] : 1.f;
    if (a.d_w) in.gw = __ldg(a.d_w + p);
}

// ---- sub-field mode (router: fields/PreSight/ingp_field_ms.py:80-126) -------------------------------------
// The level's points arrive grouped by sub-field (ms_route.cu): row i is point perm[i] (-1 = padding), tiles of 128 rows
// are sub-field-homogeneous (segments padded to 256 rows), hash features were gathered from each tile's own table in
// row order.  Rays are not contiguous in a tile any more, so these kernels stop at the per-point field outputs
// (density, rgb, semantics) and start from their gradients; compositing runs in ps_composite_fwd / _bwd.
struct FieldMsArgs {
    const FieldNet* nets;        // one per sub-field, array in device memory (dW / dB used by the backward)
    const float* feat;           // level-major hash features [L][rows][F] in row order
    float* dfeat;                // backward: same layout (written)
    int L, F;
    const uint8_t* sels;         // [rows]
    const int32_t* perm;         // [rows]
    const uint8_t* tile_sf;      // [rows / 128], 255 = unused
    int64_t rows;
    int S;                       // samples per ray: ray of point p = p / S
    const float* dirs;           // [N, 3]
    const float* app;            // [N, A] (nullable)
    float* dapp;                 // backward: [N, A] accumulated (nullable)
    // forward outputs, written at the point's own index
    float* density;              // [P]
    float* rgb;                  // [P, 3]
    float* sem;                  // [P, 64]
    // backward inputs
    const float* d_density;      // [P]
    const float* d_rgb;          // [P, 3]
    const float* d_sem;          // [P, 64]
};

template <int K0>
__device__ __forceinline__ void load_row_inputs_ms(const FieldMsArgs& a, const FieldNet& net, int64_t i, int32_t p,
                                                   bool want_feat, bool want_ray, RowInputs<K0>& in) {
#pragma unroll
    for (int c = 0; c < K0; ++c) in.feat[c] = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) in.app[c] = 0.f;
    in.dir[0] = in.dir[1] = in.dir[2] = 0.f;
    in.t0 = in.t1 = in.selv = in.gw = 0.f;
    if (p < 0) return;
    if (want_feat) {
        if (a.F == 2) {
#pragma unroll
            for (int l = 0; l < K0 / 2; ++l)
                if (l < a.L) {
                    const float2 q = __ldg(reinterpret_cast<const float2*>(a.feat + ((int64_t)l * a.rows + i) * 2));
                    in.feat[2 * l] = q.x;
                    in.feat[2 * l + 1] = q.y;
                }
        } else {  // F == 4
#pragma unroll
            for (int l = 0; l < K0 / 4; ++l)
                if (l < a.L) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(a.feat + ((int64_t)l * a.rows + i) * 4));
                    in.feat[4 * l] = q.x; in.feat[4 * l + 1] = q.y; in.feat[4 * l + 2] = q.z; in.feat[4 * l + 3] = q.w;
                }
        }
    }
    if (want_ray) {
        const int64_t ray = p / a.S;
        in.dir[0] = __ldg(a.dirs + 3 * ray);
        in.dir[1] = __ldg(a.dirs + 3 * ray + 1);
        in.dir[2] = __ldg(a.dirs + 3 * ray + 2);
        if (a.app)
#pragma unroll
            for (int c = 0; c < 16; ++c)
                if (c < net.app_dim) in.app[c] = __ldg(a.app + ray * net.app_dim + c);
    }
    in.selv = (float)a.sels[i];
}

// per-ray inputs only (view direction, appearance embedding)
template <int K0>
__device__ __forceinline__ void load_ray_inputs(const FieldArgs& a, int64_t ray, bool valid, RowInputs<K0>& in) {
    if (!valid) return;
    in.dir[0] = __ldg(a.dirs + 3 * ray);
    in.dir[1] = __ldg(a.dirs + 3 * ray + 1);
    in.dir[2] = __ldg(a.dirs + 3 * ray + 2);
    if (a.app)
#pragma unroll
        for (int c = 0; c < 16; ++c)
            if (c < a.net.app_dim) in.app[c] = __ldg(a.app + ray * a.net.app_dim + c);
}

// features -> bf16 -> the first K0 columns of `tile`
template <int K0>
__device__ __forceinline__ void stage_features(const RowInputs<K0>& in, unsigned char* tile, int r) {
#pragma unroll
    for (int c = 0; c < K0; c += 8) store_chunk(tile, kRows, r, c, in.feat + c);
}

// [sh16(direction) | appearance (zero padded to 16)] of this row's ray -> the 32-column SHAPP tile
template <int K0>
__device__ __forceinline__ void stage_shapp(const RowInputs<K0>& in, bool valid, unsigned char* tile, int r) {
    float v[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = 0.f;
    if (valid) {
        float sh[16];
        sh4_of_direction(in.dir[0], in.dir[1], in.dir[2], sh);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            v[c] = sh[c];
            v[16 + c] = in.app[c];
        }
    }
#pragma unroll
    for (int c = 0; c < 32; c += 8) store_chunk(tile, kRows, r, c, v + c);
}

}  // namespace ftc5
}  // namespace ps
