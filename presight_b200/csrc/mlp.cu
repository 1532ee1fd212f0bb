// C-ABI of kernel #2 (fused MLP): argument validation and shape dispatch.
#include "mlp_dispatch.cuh"

using namespace ps;
using namespace ps::mma;

static int pad16(int v) { return (v + 15) / 16 * 16; }

static int resolve(const int* dims, int n_layers, int& K0, int& H, int& NHID, int& NOUT) {
    PS_REQUIRE(dims != nullptr, "mlp: dims_host is null");
    PS_REQUIRE(n_layers >= 1 && n_layers <= 4, "mlp: %d layers unsupported (1..4)", n_layers);
    for (int i = 0; i <= n_layers; ++i) PS_REQUIRE(dims[i] >= 1, "mlp: dims[%d] = %d", i, dims[i]);
    K0 = pad16(dims[0]);
    NOUT = pad16(dims[n_layers]);
    NHID = n_layers - 1;
    H = NHID > 0 ? dims[1] : 16;
    for (int i = 1; i < n_layers; ++i)
        PS_REQUIRE(dims[i] == H, "mlp: hidden widths must be equal (got %d and %d)", H, dims[i]);
    PS_REQUIRE(H % 16 == 0, "mlp: hidden width %d must be a multiple of 16", H);
    return 0;
}

static int run(int K0, int H, int NHID, int NOUT, int prec, bool bwd, const MlpArgs& a, cudaStream_t s,
               const int* dims, int n_layers) {
    PS_REQUIRE(prec == 0 || prec == 1, "mlp: precision %d (0 = tf32x3 fp32-grade, 1 = bf16)", prec);
    int r = dispatch_group0(K0, H, NHID, NOUT, prec, bwd, a, s);
    if (r < 0) r = dispatch_group1(K0, H, NHID, NOUT, prec, bwd, a, s);
    if (r < 0) r = dispatch_group2(K0, H, NHID, NOUT, prec, bwd, a, s);
    if (r < 0) {
        set_error("mlp: no kernel instantiated for %d -> %d x%d -> %d (padded %d/%d/%d); add it to mlp_dispatch.cuh",
                  dims[0], H, NHID, dims[n_layers], K0, H, NOUT);
        return 3;
    }
    return r;
}

extern "C" int ps_mlp_fwd(const float* x, int64_t P, const float* const* W_host, const float* const* b_host,
                          const int* dims_host, int n_layers, int out_act, int precision, float* y, void* stream) {
    int K0, H, NHID, NOUT;
    if (int e = resolve(dims_host, n_layers, K0, H, NHID, NOUT)) return e;
    if (P == 0) return 0;
    PS_REQUIRE(x && y && W_host && b_host, "mlp_fwd: null pointer");
    MlpArgs a{};
    a.x = x; a.y = y; a.P = P; a.in_dim = dims_host[0]; a.out_dim = dims_host[n_layers]; a.out_act = out_act;
    for (int i = 0; i < n_layers; ++i) {
        PS_REQUIRE(W_host[i] != nullptr, "mlp_fwd: weight %d is null", i);
        a.W[i] = W_host[i];
        a.b[i] = b_host[i];
    }
    return run(K0, H, NHID, NOUT, precision, false, a, (cudaStream_t)stream, dims_host, n_layers);
}

extern "C" int ps_mlp_bwd(const float* x, const float* y, const float* dy, int64_t P, const float* const* W_host,
                          const float* const* b_host, const int* dims_host, int n_layers, int out_act, int precision,
                          float* dx, float* const* dW_host, float* const* db_host, void* stream) {
    (void)y;  // the forward is recomputed on chip
    int K0, H, NHID, NOUT;
    if (int e = resolve(dims_host, n_layers, K0, H, NHID, NOUT)) return e;
    if (P == 0) return 0;
    PS_REQUIRE(x && dy && W_host && b_host && dW_host && db_host, "mlp_bwd: null pointer");
    MlpArgs a{};
    a.x = x; a.dy = dy; a.dx = dx; a.P = P; a.in_dim = dims_host[0]; a.out_dim = dims_host[n_layers];
    a.out_act = out_act;
    for (int i = 0; i < n_layers; ++i) {
        PS_REQUIRE(W_host[i] != nullptr && dW_host[i] != nullptr, "mlp_bwd: weight/grad %d is null", i);
        a.W[i] = W_host[i];
        a.b[i] = b_host[i];
        a.dW[i] = dW_host[i];
        a.db[i] = db_host[i];
    }
    return run(K0, H, NHID, NOUT, precision, true, a, (cudaStream_t)stream, dims_host, n_layers);
}
