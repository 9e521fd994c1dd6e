// Fused continuous-normalizing-flow vector field with Hutchinson's trace estimator, for narrow MLPs.
//
// Reference: torchdyn/models/cnf.py:26-31 (hutch_trace), :55-65 (CNF.forward): for the state s = [logp, x],
//     F(s) = [ -eps^T (df/dx) eps ,  f(x) ] (+ 0*s),
// where the reference obtains eps^T (df/dx) with one autograd VJP and the adjoint (numerics/sensitivity.py:57-75)
// differentiates THROUGH that VJP (second-order autograd, create_graph=True).  Here both are written out for
// f = W_L act(... act(W_1 x + b_1) ...) + b_L:
//   forward  : z_l = W_l h_{l-1} + b_l, h_l = act(z_l)
//   VJP      : a_L = eps; g_l = W_l^T a_l; a_{l-1} = g_l * act'(z_{l-1});  tr = eps . g_1
//   adjoint  : cotangents c = lam_logp on tr and -lam_x on f.  Reverse of the VJP (increasing l):
//                gbar_1 = c*eps; abar_l = W_l gbar_l; Wbar_l += a_l (x) gbar_l;
//                gbar_{l+1} = abar_l * act'(z_l); zbar_l = abar_l * g_{l+1} * act''(z_l)
//              then the ordinary backward of the forward pass with delta_L = -lam_x and zbar_l added to the
//              pre-activation cotangents; xbar = W_1^T delta_1.
//   dlam = [0, xbar],  dmu = sum over the batch of (Wbar, bbar),  ds = [-tr, f].
// One launch replaces the reference's forward + VJP (+ double backward) autograd graphs.  Same tiling as the MLP FFMA
// kernels (tdyn_mlp_ffma.cuh): 32-row tiles, weights in shared memory, feature-major padded activations.
#include "tdyn_mlp_ffma.cuh"

namespace tdyn {
namespace cnf {

using namespace ffma;

// dst[k][r] = sum_n W[n][k] * src[n][r]  (W^T src), optionally times act'(h[k][r])
template <int ACT>
__device__ __forceinline__ void matvec_t(const Net& net, const float* w_s, int l, const float* src, float* dst, const float* h_for_deriv) {
    const int K = net.dims[l], N = net.dims[l + 1], ld = w_ld(K);
    if ((K & 3) == 0) {
        gemm_kn_4x4(N, K, w_s + net.w_off[l], ld, src, dst, [&](float v, int k, int r) {
            return h_for_deriv ? v * act_bwd<ACT>(h_for_deriv[k * kLd + r]) : v;
        });
        return;
    }
    for (int i = threadIdx.x; i < kRG * K; i += kThreads) {
        const int g = i / K, k = i - g * K;
        const float* wc = w_s + net.w_off[l] + k;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int nn = 0; nn < N; ++nn) {
            const float w = wc[nn * ld];
            const float4 sv = *reinterpret_cast<const float4*>(src + nn * kLd + 4 * g);
            acc.x = fmaf(w, sv.x, acc.x); acc.y = fmaf(w, sv.y, acc.y); acc.z = fmaf(w, sv.z, acc.z); acc.w = fmaf(w, sv.w, acc.w);
        }
        if (h_for_deriv) {
            const float4 hv = *reinterpret_cast<const float4*>(h_for_deriv + k * kLd + 4 * g);
            acc.x *= act_bwd<ACT>(hv.x); acc.y *= act_bwd<ACT>(hv.y); acc.z *= act_bwd<ACT>(hv.z); acc.w *= act_bwd<ACT>(hv.w);
        }
        *reinterpret_cast<float4*>(dst + k * kLd + 4 * g) = acc;
    }
}

// dst[n][r] = sum_k W[n][k] * src[k][r]  (W src, no bias)
__device__ __forceinline__ void matvec(const Net& net, const float* w_s, int l, const float* src, float* dst) {
    const int K = net.dims[l], N = net.dims[l + 1], ld = w_ld(K);
    if (((N | K) & 3) == 0) {
        // (threads 256..: the outer product of the same phase runs on threads 0..255)
        gemm_nk_4x4(N, K, w_s + net.w_off[l], ld, nullptr, src, dst, [](float v, int, int) { return v; }, 256);
        return;
    }
    for (int i = threadIdx.x; i < kRG * N; i += kThreads) {
        const int g = i / N, n = i - g * N;
        const float* wr = w_s + net.w_off[l] + n * ld;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        // (unrolled: eight iterations' shared-memory loads are in flight before their FFMAs; the accumulation order is unchanged)
#pragma unroll 8
        for (int k = 0; k < K; ++k) {
            const float w = wr[k];
            const float4 sv = *reinterpret_cast<const float4*>(src + k * kLd + 4 * g);
            acc.x = fmaf(w, sv.x, acc.x); acc.y = fmaf(w, sv.y, acc.y); acc.z = fmaf(w, sv.z, acc.z); acc.w = fmaf(w, sv.w, acc.w);
        }
        *reinterpret_cast<float4*>(dst + n * kLd + 4 * g) = acc;
    }
}

// tile loaders for the [rows, D+1] state layout: column 0 = logp, columns 1..D = x
__device__ __forceinline__ void load_x(float* dst, const float* __restrict__ s, int D, int64_t row0, int64_t rows, float scale) {
    for (int i = threadIdx.x; i < kRows * D; i += kThreads) {
        const int r = i / D, d = i - r * D;
        const int64_t gr = row0 + r;
        dst[d * kLd + r] = gr < rows ? scale * __ldg(s + gr * (D + 1) + 1 + d) : 0.f;
    }
}

// forward pass + VJP with cotangent eps; leaves h_l in h_s, a_l (l = 1..L) in a_s, g_l (l = 1..L, stored at h_off[l-1]) in
// q_s and the per-row trace estimates in tr_s[kRows]
template <int ACT>
__device__ __forceinline__ void forward_and_trace(const Net& net, const float* w_s, float* h_s, float* a_s, float* q_s, float* tr_s,
                                                  const float* __restrict__ eps, int64_t row0, int64_t rows) {
    const int D = net.dims[0], L = net.L;
    forward_tile<ACT>(net, w_s, h_s);
    load_tile(a_s + net.h_off[L], eps, D, row0, rows, 1.f);          // a_L = eps
    __syncthreads();
    for (int l = L - 1; l >= 0; --l) {
        // g = W_l^T a_{l+1} (code layers are 0-based: layer l maps h_l -> z_{l+1})
        matvec_t<ACT>(net, w_s, l, a_s + net.h_off[l + 1], q_s + net.h_off[l], nullptr);
        __syncthreads();
        if (l > 0) {
            const int K = net.dims[l];
            for (int i = threadIdx.x; i < K * kRows; i += kThreads) {
                const int k = i / kRows, r = i - k * kRows;
                const int o = net.h_off[l] + k * kLd + r;
                a_s[o] = q_s[o] * act_bwd<ACT>(h_s[o]);
            }
            __syncthreads();
        }
    }
    if (threadIdx.x < kRows) {
        float t = 0.f;
        for (int d = 0; d < D; ++d) t = fmaf(a_s[net.h_off[L] + d * kLd + threadIdx.x], q_s[net.h_off[0] + d * kLd + threadIdx.x], t);
        tr_s[threadIdx.x] = t;
    }
    __syncthreads();
}

template <int ACT>
__global__ void __launch_bounds__(kThreads) cnf_forward_kernel(const Net net, const float* __restrict__ s, const float* __restrict__ eps,
                                                               float* __restrict__ out, int64_t rows, float sign) {
    extern __shared__ __align__(16) float sm[];
    float* w_s = sm;
    float* h_s = w_s + net.w_floats;
    float* a_s = h_s + net.h_floats;
    float* q_s = a_s + net.h_floats;
    float* tr_s = q_s + net.h_floats;
    load_weights(net, w_s);
    const int D = net.dims[0];
    const int64_t tiles = (rows + kRows - 1) / kRows;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        __syncthreads();
        const int64_t row0 = t * kRows;
        load_x(h_s + net.h_off[0], s, D, row0, rows, 1.f);
        __syncthreads();
        forward_and_trace<ACT>(net, w_s, h_s, a_s, q_s, tr_s, eps, row0, rows);
        for (int i = threadIdx.x; i < kRows * (D + 1); i += kThreads) {
            const int r = i / (D + 1), c = i - r * (D + 1);
            const int64_t gr = row0 + r;
            if (gr < rows) out[gr * (D + 1) + c] = sign * (c == 0 ? -tr_s[r] : h_s[net.h_off[net.L] + (c - 1) * kLd + r]);
        }
    }
}

template <int ACT>
__global__ void __launch_bounds__(kThreads) cnf_adjoint_kernel(const Net net, const float* __restrict__ s, const float* __restrict__ lam,
                                                               const float* __restrict__ eps, float* __restrict__ ds,
                                                               float* __restrict__ dlam, float* __restrict__ dmu, AdjWs* ws,
                                                               int64_t rows, float sign) {
    extern __shared__ __align__(16) float sm[];
    float* w_s = sm;
    float* h_s = w_s + net.w_floats;
    float* a_s = h_s + net.h_floats;
    float* q_s = a_s + net.h_floats;
    float* z_s = q_s + net.h_floats;                 // zbar_l (second-order terms), activation-area layout
    float* tr_s = z_s + net.h_floats;                // [kRows] trace, then [kRows] c = lam_logp
    float* d_s = tr_s + 2 * kRows;                   // two scratch tiles [maxdim][kLd]
    float* g_s = d_s + 2 * net.maxdim * kLd;         // parameter-gradient accumulator [n_params]
    __shared__ bool is_last;
    load_weights(net, w_s);
    for (int i = threadIdx.x; i < net.n_params; i += kThreads) g_s[i] = 0.f;
    const int D = net.dims[0], L = net.L;
    const int64_t tiles = (rows + kRows - 1) / kRows;
    auto gacc = [&](int idx, float v) { g_s[idx] += v; };
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        __syncthreads();
        const int64_t row0 = t * kRows;
        load_x(h_s + net.h_off[0], s, D, row0, rows, 1.f);
        if (threadIdx.x < kRows) {
            const int64_t gr = row0 + threadIdx.x;
            tr_s[kRows + threadIdx.x] = gr < rows ? __ldg(lam + gr * (D + 1)) : 0.f;      // c = lam_logp
        }
        __syncthreads();
        forward_and_trace<ACT>(net, w_s, h_s, a_s, q_s, tr_s, eps, row0, rows);
        for (int i = threadIdx.x; i < kRows * (D + 1); i += kThreads) {
            const int r = i / (D + 1), c = i - r * (D + 1);
            const int64_t gr = row0 + r;
            if (gr < rows) ds[gr * (D + 1) + c] = sign * (c == 0 ? -tr_s[r] : h_s[net.h_off[L] + (c - 1) * kLd + r]);
        }
        // ---- reverse of the VJP: gbar_1 = c * eps
        float* gbar = d_s;
        float* abar = d_s + net.maxdim * kLd;
        for (int i = threadIdx.x; i < D * kRows; i += kThreads) {
            const int d = i / kRows, r = i - d * kRows;
            gbar[d * kLd + r] = tr_s[kRows + r] * a_s[net.h_off[L] + d * kLd + r];
        }
        __syncthreads();
        for (int l = 0; l < L; ++l) {
            const int K = net.dims[l], N = net.dims[l + 1];
            matvec(net, w_s, l, gbar, abar);                                            // abar_{l+1} = W_l gbar_l
            outer_tile(N, K, a_s + net.h_off[l + 1], gbar, net.p_off[l], gacc);         // Wbar_l += a_{l+1} (x) gbar_l
            __syncthreads();
            if (l + 1 < L) {
                for (int i = threadIdx.x; i < N * kRows; i += kThreads) {
                    const int n = i / kRows, r = i - n * kRows;
                    const int o = net.h_off[l + 1] + n * kLd + r;
                    const float ab = abar[n * kLd + r], hv = h_s[o];
                    gbar[n * kLd + r] = ab * act_bwd<ACT>(hv);                          // gbar_{l+1}
                    z_s[o] = ab * q_s[o] * act_bwd2<ACT>(hv);                           // zbar = abar * g_{l+2} * act''
                }
                __syncthreads();
            }
        }
        // ---- ordinary backward with delta_L = -lam_x and the zbar terms
        load_x(d_s, lam, D, row0, rows, -1.f);
        __syncthreads();
        float* dfin = backward_tile<ACT>(net, w_s, h_s, d_s, d_s + net.maxdim * kLd, gacc, z_s);
        for (int i = threadIdx.x; i < kRows * (D + 1); i += kThreads) {
            const int r = i / (D + 1), c = i - r * (D + 1);
            const int64_t gr = row0 + r;
            if (gr < rows) dlam[gr * (D + 1) + c] = c == 0 ? 0.f : sign * dfin[(c - 1) * kLd + r];
        }
    }
    // ---- deterministic cross-CTA reduction of the parameter gradient
    __syncthreads();
    float* mine = ws->partials + (size_t)blockIdx.x * net.n_params;
    for (int i = threadIdx.x; i < net.n_params; i += kThreads) mine[i] = g_s[i];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int tk = atomicAdd(&ws->ticket, 1u);
        is_last = tk == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int i = threadIdx.x; i < net.n_params; i += kThreads) {
        float sv = 0.f;
        for (unsigned int b = 0; b < gridDim.x; ++b) sv += __ldcg(ws->partials + (size_t)b * net.n_params + i);
        dmu[i] = sign * sv;
    }
    if (threadIdx.x == 0) ws->ticket = 0u;
}

static size_t fwd_smem(const Net& n) { return sizeof(float) * (size_t)(n.w_floats + 3 * n.h_floats + kRows); }
static size_t adj_smem(const Net& n) {
    return sizeof(float) * (size_t)(n.w_floats + 4 * n.h_floats + 2 * kRows + 2 * n.maxdim * kLd + n.n_params);
}

}  // namespace cnf
}  // namespace tdyn

using namespace tdyn;

extern "C" {

int tdyn_cnf_ffma_supported(const tdyn_mlp_t* m) {
    ffma::Net net;
    if (!m || ffma::make_net(m, net, "tdyn_cnf_ffma_supported") != 0) return 0;
    if (!(m->act == TDYN_ACT_TANH || m->act == TDYN_ACT_SOFTPLUS || m->act == TDYN_ACT_RELU)) return 0;
    return cnf::adj_smem(net) <= ffma::kSmemLimit ? 1 : 0;
}

#define TDYN_CNF_SWITCH(KERNEL, ...)                                                                              \
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

int tdyn_cnf_ffma_forward(const tdyn_mlp_t* m, const float* s, const float* eps, float* out, int64_t rows, float sign,
                          void* stream) {
    ffma::Net net;
    if (int rc = ffma::make_net(m, net, "tdyn_cnf_ffma_forward")) return rc;
    TDYN_REQUIRE(s && eps && out, "tdyn_cnf_ffma_forward: null argument");
    if (rows <= 0) return 0;
    const size_t smem = cnf::fwd_smem(net);
    TDYN_REQUIRE(smem <= ffma::kSmemLimit, "tdyn_cnf_ffma_forward: network too large for shared memory (%zu B)", smem);
    const int64_t tiles = (rows + ffma::kRows - 1) / ffma::kRows;
    const int grid = (int)(tiles < 148 * 4 ? tiles : 148 * 4);
    cudaStream_t st = (cudaStream_t)stream;
    TDYN_CNF_SWITCH(cnf::cnf_forward_kernel, net, s, eps, out, rows, sign)
    TDYN_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int tdyn_cnf_ffma_adjoint(const tdyn_mlp_t* m, const float* s, const float* lam, const float* eps, float* ds, float* dlam,
                          float* dmu, void* workspace, int64_t rows, float sign, void* stream) {
    ffma::Net net;
    if (int rc = ffma::make_net(m, net, "tdyn_cnf_ffma_adjoint")) return rc;
    TDYN_REQUIRE(s && lam && eps && ds && dlam && dmu && workspace, "tdyn_cnf_ffma_adjoint: null argument");
    TDYN_REQUIRE(rows > 0, "tdyn_cnf_ffma_adjoint: rows must be positive");
    const size_t smem = cnf::adj_smem(net);
    TDYN_REQUIRE(smem <= ffma::kSmemLimit, "tdyn_cnf_ffma_adjoint: network too large for shared memory (%zu B)", smem);
    const int64_t tiles = (rows + ffma::kRows - 1) / ffma::kRows;
    const int grid = (int)(tiles < ffma::kMaxGrid ? tiles : ffma::kMaxGrid);
    cudaStream_t st = (cudaStream_t)stream;
    TDYN_CNF_SWITCH(cnf::cnf_adjoint_kernel, net, s, lam, eps, ds, dlam, dmu, (ffma::AdjWs*)workspace, rows, sign)
    TDYN_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
