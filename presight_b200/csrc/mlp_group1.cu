#include "mlp_dispatch.cuh"
namespace ps {
namespace mma {
int dispatch_group1(int K0, int H, int NHID, int NOUT, int prec, bool bwd, const MlpArgs& a, cudaStream_t s) {
    PS_MLP_GROUP1(PS_MLP_CASE)
    return -1;
}
}  // namespace mma
}  // namespace ps
