// K1/K3 (stage form): FFMA kernels for NARROW MLP vector fields (every width <= 128, <= 4 Linear layers).
//
//   tdyn_mlp_ffma_forward : out = sign * f(y)                                   (one launch instead of ~3 ATen launches)
//   tdyn_mlp_ffma_adjoint : dx   = sign * f(x)
//                           dlam = sign * (-lam^T df/dx)                        (reference numerics/sensitivity.py:62-66)
//                           dmu  = sign * (-sum_batch lam^T df/dtheta)          (parameter order of vf.parameters())
//                           (one launch instead of the ~40 launches of unflatten + forward + autograd.grad + cat)
//
// Narrow nets are latency / issue bound, not tensor-core work (north_star: FFMA below width 256): all weights sit in
// shared memory ([out][in+1] padded -> conflict-free for both the forward (lanes over out) and the dgrad (lanes over
// in) access), a CTA walks 32-row tiles, activations are kept feature-major ([feature][32 rows + pad]) so one LDS.128
// feeds 4 rows, every thread owns (4 rows x 1 neuron).  The parameter-gradient contraction (split-K over the batch)
// is accumulated per CTA in shared memory and reduced across CTAs in a FIXED order by the last CTA (deterministic).
#include "tdyn_mlp_ffma.cuh"


using namespace tdyn;

extern "C" {

int tdyn_mlp_ffma_supported(const tdyn_mlp_t* m) {
    ffma::Net net;
    if (!m || ffma::make_net(m, net, "tdyn_mlp_ffma_supported") != 0) return 0;
    if (!(m->act == TDYN_ACT_TANH || m->act == TDYN_ACT_SOFTPLUS || m->act == TDYN_ACT_RELU)) return 0;
    return ffma::adj_smem(net) <= ffma::kSmemLimit ? 1 : 0;
}

int64_t tdyn_mlp_ffma_workspace_bytes(const tdyn_mlp_t* m) {
    ffma::Net net;
    if (!m || ffma::make_net(m, net, "tdyn_mlp_ffma_workspace_bytes") != 0) return 0;
    return (int64_t)(128 + sizeof(float) * (size_t)ffma::kMaxGrid * net.n_params);
}

#define TDYN_ACT_SWITCH(KERNEL, ...)                                                                              \
    switch (net.act) {                                                                                            \
        case TDYN_ACT_TANH:                                                                                       \
            TDYN_CHECK_CUDA(cudaFuncSetAttribute(KERNEL<TDYN_ACT_TANH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            KERNEL<TDYN_ACT_TANH><<<grid, ffma::kThreads, smem, st>>>(__VA_ARGS__); break;                         \
        case TDYN_ACT_SOFTPLUS:                                                                                   \
            TDYN_CHECK_CUDA(cudaFuncSetAttribute(KERNEL<TDYN_ACT_SOFTPLUS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            KERNEL<TDYN_ACT_SOFTPLUS><<<grid, ffma::kThreads, smem, st>>>(__VA_ARGS__); break;                     \
        default:                                                                                                  \
            TDYN_CHECK_CUDA(cudaFuncSetAttribute(KERNEL<TDYN_ACT_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            KERNEL<TDYN_ACT_RELU><<<grid, ffma::kThreads, smem, st>>>(__VA_ARGS__); break;                         \
    }

int tdyn_mlp_ffma_forward(const tdyn_mlp_t* m, const float* y, float* out, int64_t rows, float sign, void* stream) {
    ffma::Net net;
    if (int rc = ffma::make_net(m, net, "tdyn_mlp_ffma_forward")) return rc;
    TDYN_REQUIRE(y && out, "tdyn_mlp_ffma_forward: null argument");
    if (rows <= 0) return 0;
    const size_t smem = ffma::fwd_smem(net);
    TDYN_REQUIRE(smem <= ffma::kSmemLimit, "tdyn_mlp_ffma_forward: network too large for shared memory (%zu B)", smem);
    const int64_t tiles = (rows + ffma::kRows - 1) / ffma::kRows;
    const int grid = (int)(tiles < 148 * 4 ? tiles : 148 * 4);
    cudaStream_t st = (cudaStream_t)stream;
    TDYN_ACT_SWITCH(ffma::mlp_forward_kernel, net, y, out, rows, sign)
    TDYN_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int tdyn_mlp_ffma_adjoint(const tdyn_mlp_t* m, const float* x, const float* lam, float* dx, float* dlam, float* dmu,
                          void* workspace, int64_t rows, float sign, void* stream) {
    ffma::Net net;
    if (int rc = ffma::make_net(m, net, "tdyn_mlp_ffma_adjoint")) return rc;
    TDYN_REQUIRE(x && lam && dx && dlam && dmu && workspace, "tdyn_mlp_ffma_adjoint: null argument");
    TDYN_REQUIRE(rows > 0, "tdyn_mlp_ffma_adjoint: rows must be positive");
    const size_t smem = ffma::adj_smem(net);
    TDYN_REQUIRE(smem <= ffma::kSmemLimit, "tdyn_mlp_ffma_adjoint: network too large for shared memory (%zu B)", smem);
    const int64_t tiles = (rows + ffma::kRows - 1) / ffma::kRows;
    const int grid = (int)(tiles < ffma::kMaxGrid ? tiles : ffma::kMaxGrid);
    cudaStream_t st = (cudaStream_t)stream;
    TDYN_ACT_SWITCH(ffma::mlp_adjoint_kernel, net, x, lam, dx, dlam, dmu, (ffma::AdjWs*)workspace, rows, sign)
    TDYN_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
