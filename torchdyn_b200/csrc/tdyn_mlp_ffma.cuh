// Device code shared by the narrow-MLP FFMA kernels (stage form: tdyn_mlp_ffma.cu; persistent solver: tdyn_mlp_solve.cu).
#pragma once
#include "tdyn_common.cuh"

namespace tdyn {
namespace ffma {

constexpr int kThreads = 512;
constexpr int kRows = 32;            // rows per tile
constexpr int kRG = kRows / 4;       // row groups of 4
constexpr int kLd = kRows + 4;       // row stride of the feature-major tiles: 36 floats = 9 x 16 B (odd), so that 8 lanes
                                     // reading consecutive features with LDS.128 touch 8 distinct 16-byte bank groups
                                     // (an unpadded stride of 8 x 16 B would be 8-way conflicted)
constexpr int kMaxWidth = 128;
constexpr int kMaxGrid = 148 * 2;
#ifndef TDYN_FFMA_FEAT_TILE
#define TDYN_FFMA_FEAT_TILE 2
#endif
constexpr int kFeatTile = TDYN_FFMA_FEAT_TILE;      // features per thread in the register-tiled GEMM pieces (x 4 rows)

// Row stride of a weight matrix [N][K] in shared memory: a multiple of 4 floats (rows are 16-byte aligned, so four
// consecutive k of one neuron are one LDS.128) and = 4 mod 8, so that 8 consecutive rows start in 8 distinct 16-byte bank
// groups (lanes that read the same k-quad of consecutive neurons do not conflict).
__host__ __device__ __forceinline__ int w_ld(int K) { return ((K + 3) & ~3) + 4; }

struct Net {
    int L;
    int dims[TDYN_MAX_LAYERS + 1];
    int act;
    const float* w[TDYN_MAX_LAYERS];
    const float* b[TDYN_MAX_LAYERS];
    int w_off[TDYN_MAX_LAYERS];      // float offsets into the shared weight area ([N][w_ld(K)])
    int b_off[TDYN_MAX_LAYERS];
    int p_off[TDYN_MAX_LAYERS];      // float offsets of layer l's (weight, bias) block in the flat parameter vector
    int h_off[TDYN_MAX_LAYERS + 1];  // float offsets of h_l ([dims[l]][kLd]) in the activation area
    int n_params, w_floats, h_floats, maxdim;
};

template <int ACT>
__device__ __forceinline__ float act_fwd(float z) {
    if (ACT == TDYN_ACT_RELU) return fmaxf(z, 0.f);
    if (ACT == TDYN_ACT_SOFTPLUS) return z > 20.f ? z : log1pf(expf(z));
    return tanhf(z);
}
// derivative of the activation expressed through its OUTPUT h
template <int ACT>
__device__ __forceinline__ float act_bwd(float h) {
    if (ACT == TDYN_ACT_RELU) return h > 0.f ? 1.f : 0.f;
    if (ACT == TDYN_ACT_SOFTPLUS) return 1.f - expf(-h);      // sigmoid(z) = 1 - exp(-softplus(z))
    return 1.f - h * h;
}

// second derivative of the activation through its output h
template <int ACT>
__device__ __forceinline__ float act_bwd2(float h) {
    if (ACT == TDYN_ACT_RELU) return 0.f;
    if (ACT == TDYN_ACT_SOFTPLUS) { const float s = 1.f - expf(-h); return s * (1.f - s); }
    return -2.f * h * (1.f - h * h);
}

__device__ __forceinline__ void load_weights(const Net& net, float* w_s) {
    for (int l = 0; l < net.L; ++l) {
        const int K = net.dims[l], N = net.dims[l + 1];
        const int ld = w_ld(K);
        for (int i = threadIdx.x; i < N * K; i += kThreads) w_s[net.w_off[l] + (i / K) * ld + (i % K)] = __ldg(net.w[l] + i);
        for (int i = threadIdx.x; i < N; i += kThreads) w_s[net.b_off[l] + i] = __ldg(net.b[l] + i);
    }
}

// h_0[d][r] <- src[(row0 + r) * D + d]   (zero beyond `rows`)
__device__ __forceinline__ void load_tile(float* dst, const float* __restrict__ src, int D, int64_t row0, int64_t rows, float scale) {
    for (int i = threadIdx.x; i < kRows * D; i += kThreads) {
        const int r = i / D, d = i % D;
        const int64_t gr = row0 + r;
        dst[d * kLd + r] = gr < rows ? scale * __ldg(src + gr * D + d) : 0.f;
    }
}
__device__ __forceinline__ void store_tile(float* __restrict__ dst, const float* src, int D, int64_t row0, int64_t rows, float sign) {
    for (int i = threadIdx.x; i < kRows * D; i += kThreads) {
        const int r = i / D, d = i % D;
        const int64_t gr = row0 + r;
        if (gr < rows) dst[gr * D + d] = sign * src[d * kLd + r];
    }
}

// `t0`: the thread that takes item 0.  Independent pieces of one phase (parameter gradient / input gradient / bias
// gradient) start at different threads, so that they run on different warps at the same time instead of one after the
// other on the first warps.
// ---- register-tiled GEMM pieces.  A thread owns a block of 4 rows of the tile x kFeatTile features and steps the
// contraction four at a time: kFeatTile + 4 LDS.128 feed 16 * kFeatTile FFMA (the first version owned 4 x 1: two LDS per 4
// FFMA, shared-memory bound at ~1/10 of the FP32 rate).  Measured on config 4 / the 2-64-64-64-2 persistent solve:
// kFeatTile = 4 (128 busy threads per 64-wide layer) 14.7 / 9.05 ms, kFeatTile = 2 (256 busy threads) 14.3 / 7.9 ms.  Feature ownership is interleaved (feature = group + i * n_groups) where the lanes of a warp
// read different weight rows, consecutive where they read one row: all weight / activation reads are conflict-free
// (w_ld, kLd).  Summation order per output element is unchanged (ascending contraction index): same bits as before.

// out[n][r] = epi(bias[n] + sum_k W[n][k] * in[k][r]) for N % 4 == 0, K % 4 == 0
template <int NT = kFeatTile, class Epi>
__device__ __forceinline__ void gemm_nk_4x4(int N, int K, const float* __restrict__ W, int ld, const float* __restrict__ bias,
                                            const float* __restrict__ in, float* __restrict__ out, Epi&& epi, int t0 = 0) {
    const int NQ = N / NT;
    for (int item = (threadIdx.x + kThreads - t0) % kThreads; item < kRG * NQ; item += kThreads) {
        const int ng = item % NQ, rg = item / NQ;
        float acc[NT][4];
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            const float bv = bias ? bias[ng + NQ * i] : 0.f;
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[i][r] = bv;
        }
        const float* w0 = W + ng * ld;
        const float* h0 = in + 4 * rg;
        for (int k = 0; k < K; k += 4) {
            float4 w[NT], h[4];
#pragma unroll
            for (int i = 0; i < NT; ++i) w[i] = *reinterpret_cast<const float4*>(w0 + (NQ * i) * ld + k);
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = *reinterpret_cast<const float4*>(h0 + (k + j) * kLd);
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                const float wv[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][0] = fmaf(wv[j], h[j].x, acc[i][0]); acc[i][1] = fmaf(wv[j], h[j].y, acc[i][1]);
                    acc[i][2] = fmaf(wv[j], h[j].z, acc[i][2]); acc[i][3] = fmaf(wv[j], h[j].w, acc[i][3]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            const int n = ng + NQ * i;
            *reinterpret_cast<float4*>(out + n * kLd + 4 * rg) =
                make_float4(epi(acc[i][0], n, 4 * rg), epi(acc[i][1], n, 4 * rg + 1), epi(acc[i][2], n, 4 * rg + 2), epi(acc[i][3], n, 4 * rg + 3));
        }
    }
}

// out[k][r] = epi(sum_n W[n][k] * in[n][r]) for K % 4 == 0 (any N)
template <int KT = kFeatTile, class Epi>
__device__ __forceinline__ void gemm_kn_4x4(int N, int K, const float* __restrict__ W, int ld, const float* __restrict__ in,
                                            float* __restrict__ out, Epi&& epi, int t0 = 0) {
    const int KQ = K / KT;
    for (int item = (threadIdx.x + kThreads - t0) % kThreads; item < kRG * KQ; item += kThreads) {
        const int kg = item % KQ, rg = item / KQ;
        float acc[KT][4];
#pragma unroll
        for (int j = 0; j < KT; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[j][r] = 0.f;
        const float* w0 = W + KT * kg;
        const float* d0 = in + 4 * rg;
#pragma unroll 4
        for (int n = 0; n < N; ++n) {
            float wv[KT];
            if (KT == 4) { const float4 w = *reinterpret_cast<const float4*>(w0 + n * ld); wv[0] = w.x; wv[1] = w.y; wv[KT - 2] = w.z; wv[KT - 1] = w.w; }
            else { const float2 w = *reinterpret_cast<const float2*>(w0 + n * ld); wv[0] = w.x; wv[1] = w.y; }
            const float4 d = *reinterpret_cast<const float4*>(d0 + n * kLd);
#pragma unroll
            for (int j = 0; j < KT; ++j) {
                acc[j][0] = fmaf(wv[j], d.x, acc[j][0]); acc[j][1] = fmaf(wv[j], d.y, acc[j][1]);
                acc[j][2] = fmaf(wv[j], d.z, acc[j][2]); acc[j][3] = fmaf(wv[j], d.w, acc[j][3]);
            }
        }
#pragma unroll
        for (int j = 0; j < KT; ++j) {
            const int k = KT * kg + j;
            *reinterpret_cast<float4*>(out + k * kLd + 4 * rg) =
                make_float4(epi(acc[j][0], k, 4 * rg), epi(acc[j][1], k, 4 * rg + 1), epi(acc[j][2], k, 4 * rg + 2), epi(acc[j][3], k, 4 * rg + 3));
        }
    }
}

// forward through all layers for the tile in h_s[h_off[0]]; leaves every h_l in shared memory
template <int ACT>
__device__ __forceinline__ void forward_tile(const Net& net, const float* w_s, float* h_s) {
    for (int l = 0; l < net.L; ++l) {
        const int K = net.dims[l], N = net.dims[l + 1], ld = w_ld(K);
        const float* hin = h_s + net.h_off[l];
        float* hout = h_s + net.h_off[l + 1];
        const bool last = l == net.L - 1;
        if (((N | K) & 3) == 0) {
            gemm_nk_4x4(N, K, w_s + net.w_off[l], ld, w_s + net.b_off[l], hin, hout,
                        [&](float v, int, int) { return last ? v : act_fwd<ACT>(v); });
        } else {
            for (int i = threadIdx.x; i < kRG * N; i += kThreads) {
                const int g = i / N, n = i % N;
                const float* wr = w_s + net.w_off[l] + n * ld;
                const float bias = w_s[net.b_off[l] + n];
                float4 acc = make_float4(bias, bias, bias, bias);
                // (unrolled: eight iterations' shared-memory loads are in flight before their FFMAs; the accumulation order is unchanged)
#pragma unroll 8
                for (int k = 0; k < K; ++k) {
                    const float w = wr[k];
                    const float4 hv = *reinterpret_cast<const float4*>(hin + k * kLd + 4 * g);
                    acc.x = fmaf(w, hv.x, acc.x); acc.y = fmaf(w, hv.y, acc.y);
                    acc.z = fmaf(w, hv.z, acc.z); acc.w = fmaf(w, hv.w, acc.w);
                }
                if (!last) { acc.x = act_fwd<ACT>(acc.x); acc.y = act_fwd<ACT>(acc.y); acc.z = act_fwd<ACT>(acc.z); acc.w = act_fwd<ACT>(acc.w); }
                *reinterpret_cast<float4*>(hout + n * kLd + 4 * g) = acc;
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float s) {
    s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); return fmaf(a.w, b.w, s);
}

// gacc(base + n*K + k, sum_r A[n][r] * B[k][r]) for an [N][kLd] tile A and a [K][kLd] tile B (rows beyond the batch are 0).
// A thread owns 4 consecutive n x 1 k, lanes run over k: the B loads are conflict-free (stride kLd), the A loads are
// warp broadcasts -- 5 LDS.128 per 16 FFMA per row group.
template <class GAcc>
__device__ __forceinline__ void outer_tile(int N, int K, const float* A, const float* B, int base, GAcc&& gacc) {
    if (((N | K) & 3) == 0) {
        // 4 n x 4 k per thread, both interleaved (n = ng + i*N/4, k = kg + j*K/4), lanes over kg: 8 LDS.128 per 64 FFMA
        const int NQ = N >> 2, KQ = K >> 2;
        for (int item = threadIdx.x; item < NQ * KQ; item += kThreads) {
            const int kg = item % KQ, ng = item / KQ;
            float sacc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) sacc[i][j] = 0.f;
#pragma unroll 2
            for (int q = 0; q < kRG; ++q) {
                float4 a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(A + (ng + NQ * i) * kLd + 4 * q);
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(B + (kg + KQ * j) * kLd + 4 * q);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) sacc[i][j] = dot4(a[i], b[j], sacc[i][j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) gacc(base + (ng + NQ * i) * K + kg + KQ * j, sacc[i][j]);
        }
    } else if ((N & 3) == 0) {
        for (int i = threadIdx.x; i < (N >> 2) * K; i += kThreads) {
            const int nb = i / K, k = i - nb * K;
            const float* a0 = A + (4 * nb) * kLd;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int q = 0; q < kRG; ++q) {
                const float4 bv = *reinterpret_cast<const float4*>(B + k * kLd + 4 * q);
                s0 = dot4(*reinterpret_cast<const float4*>(a0 + 4 * q), bv, s0);
                s1 = dot4(*reinterpret_cast<const float4*>(a0 + kLd + 4 * q), bv, s1);
                s2 = dot4(*reinterpret_cast<const float4*>(a0 + 2 * kLd + 4 * q), bv, s2);
                s3 = dot4(*reinterpret_cast<const float4*>(a0 + 3 * kLd + 4 * q), bv, s3);
            }
            const int o = base + (4 * nb) * K + k;
            gacc(o, s0); gacc(o + K, s1); gacc(o + 2 * K, s2); gacc(o + 3 * K, s3);
        }
    } else {
        for (int i = threadIdx.x; i < N * K; i += kThreads) {
            const int nn = i / K, k = i - nn * K;
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < kRG; ++q)
                s = dot4(*reinterpret_cast<const float4*>(A + nn * kLd + 4 * q), *reinterpret_cast<const float4*>(B + k * kLd + 4 * q), s);
            gacc(base + i, s);
        }
    }
}

// VJP of the tile whose activations h_l are in h_s and whose output cotangent delta_L is in dcur ([dims[L]][kLd]):
// for every layer, the tile's parameter-gradient contribution (weight [N][K] then bias [N], flat index p_off[l] + ...)
// is handed to `gacc(flat_index, value)` (each index is produced by exactly one thread), and delta is pulled back
// through W_l and the activation derivative.  Returns the buffer holding delta_0 = (delta_L)^T df/dx.
// wgrad mapping: a thread owns 4 consecutive neurons x 1 input feature, lanes run over the input feature -- the
// h loads are conflict-free (stride kLd), the delta loads are warp broadcasts: 5 LDS.128 per 16 FFMA per row group.
// `zbar` (optional, activation-area layout): extra cotangent added to the pre-activation of hidden layer l
// (zbar + h_off[l]) -- the second-order terms of a CNF's trace (tdyn_cnf_ffma.cu); nullptr for a plain MLP.
template <int ACT, class GAcc>
__device__ __forceinline__ float* backward_tile(const Net& net, const float* w_s, const float* h_s, float* dcur, float* dnxt,
                                                GAcc&& gacc, const float* zbar = nullptr) {
    for (int l = net.L - 1; l >= 0; --l) {
        const int K = net.dims[l], N = net.dims[l + 1];
        const float* hin = h_s + net.h_off[l];
        const int po = net.p_off[l];
        outer_tile(N, K, dcur, hin, po, gacc);
        for (int nn = (threadIdx.x + kThreads - 384) % kThreads; nn < N; nn += kThreads) {
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < kRows; ++r) s += dcur[nn * kLd + r];
            gacc(po + N * K + nn, s);
        }
        // dgrad: delta_prev[k][r] = (sum_n delta[n][r] * W[n][k]) * act'(h_l[k][r])   (no act' into the input layer)
        const int ld = w_ld(K);
        if ((K & 3) == 0) {
            const float* zb = zbar ? zbar + net.h_off[l] : nullptr;
            gemm_kn_4x4(N, K, w_s + net.w_off[l], ld, dcur, dnxt, [&](float v, int k, int r) {
                if (l > 0) {
                    v *= act_bwd<ACT>(hin[k * kLd + r]);
                    if (zb) v += zb[k * kLd + r];
                }
                return v;
            }, 256);       // (outer_tile occupies threads 0..255 for a 64 x 64 layer, the bias gradient 384..)
        } else
        for (int i = threadIdx.x; i < kRG * K; i += kThreads) {
            const int g = i / K, k = i - g * K;
            const float* wc = w_s + net.w_off[l] + k;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
            for (int nn = 0; nn < N; ++nn) {
                const float w = wc[nn * ld];
                const float4 dv = *reinterpret_cast<const float4*>(dcur + nn * kLd + 4 * g);
                acc.x = fmaf(w, dv.x, acc.x); acc.y = fmaf(w, dv.y, acc.y);
                acc.z = fmaf(w, dv.z, acc.z); acc.w = fmaf(w, dv.w, acc.w);
            }
            if (l > 0) {
                const float4 hv = *reinterpret_cast<const float4*>(hin + k * kLd + 4 * g);
                acc.x *= act_bwd<ACT>(hv.x); acc.y *= act_bwd<ACT>(hv.y); acc.z *= act_bwd<ACT>(hv.z); acc.w *= act_bwd<ACT>(hv.w);
                if (zbar) {
                    const float4 zb = *reinterpret_cast<const float4*>(zbar + net.h_off[l] + k * kLd + 4 * g);
                    acc.x += zb.x; acc.y += zb.y; acc.z += zb.z; acc.w += zb.w;
                }
            }
            *reinterpret_cast<float4*>(dnxt + k * kLd + 4 * g) = acc;
        }
        __syncthreads();
        float* tmp = dcur; dcur = dnxt; dnxt = tmp;
    }
    return dcur;
}

template <int ACT>
__global__ void __launch_bounds__(kThreads) mlp_forward_kernel(const Net net, const float* __restrict__ y, float* __restrict__ out,
                                                               int64_t rows, float sign) {
    extern __shared__ __align__(16) float sm[];
    float* w_s = sm;
    float* h_s = sm + net.w_floats;
    load_weights(net, w_s);
    const int D = net.dims[0];
    const int64_t tiles = (rows + kRows - 1) / kRows;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        __syncthreads();
        load_tile(h_s + net.h_off[0], y, D, t * kRows, rows, 1.f);
        __syncthreads();
        forward_tile<ACT>(net, w_s, h_s);
        store_tile(out, h_s + net.h_off[net.L], D, t * kRows, rows, sign);
    }
}

struct AdjWs {
    unsigned int ticket;
    unsigned int pad[31];
    float partials[1];       // [gridDim.x][n_params]
};

template <int ACT>
__global__ void __launch_bounds__(kThreads) mlp_adjoint_kernel(const Net net, const float* __restrict__ x, const float* __restrict__ lam,
                                                               float* __restrict__ dx, float* __restrict__ dlam,
                                                               float* __restrict__ dmu, AdjWs* ws, int64_t rows, float sign) {
    extern __shared__ __align__(16) float sm[];
    float* w_s = sm;
    float* h_s = w_s + net.w_floats;
    float* d_s = h_s + net.h_floats;                 // two delta buffers [maxdim][kLd]
    float* g_s = d_s + 2 * net.maxdim * kLd;         // parameter-gradient accumulator [n_params]
    __shared__ bool is_last;
    load_weights(net, w_s);
    for (int i = threadIdx.x; i < net.n_params; i += kThreads) g_s[i] = 0.f;
    const int D = net.dims[0];
    const int64_t tiles = (rows + kRows - 1) / kRows;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        __syncthreads();
        load_tile(h_s + net.h_off[0], x, D, t * kRows, rows, 1.f);
        load_tile(d_s, lam, D, t * kRows, rows, -1.f);        // delta_L = -lam  (rows beyond the batch contribute 0)
        __syncthreads();
        forward_tile<ACT>(net, w_s, h_s);
        store_tile(dx, h_s + net.h_off[net.L], D, t * kRows, rows, sign);
        float* dfin = backward_tile<ACT>(net, w_s, h_s, d_s, d_s + net.maxdim * kLd, [&](int idx, float v) { g_s[idx] += v; });
        store_tile(dlam, dfin, D, t * kRows, rows, sign);
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
        float s = 0.f;
        for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(ws->partials + (size_t)b * net.n_params + i);
        dmu[i] = sign * s;
    }
    if (threadIdx.x == 0) ws->ticket = 0u;
}

inline int make_net(const tdyn_mlp_t* m, Net& net, const char* who) {
    TDYN_REQUIRE(m && m->n_layers >= 1 && m->n_layers <= TDYN_MAX_LAYERS, "%s: n_layers out of range", who);
    net.L = m->n_layers;
    net.act = m->act;
    int woff = 0, hoff = 0, poff = 0, maxdim = 0;
    for (int l = 0; l <= net.L; ++l) {
        net.dims[l] = m->dims[l];
        TDYN_REQUIRE(net.dims[l] >= 1 && net.dims[l] <= kMaxWidth, "%s: width %d outside 1..%d", who, net.dims[l], kMaxWidth);
        net.h_off[l] = hoff;
        hoff += net.dims[l] * kLd;
        if (net.dims[l] > maxdim) maxdim = net.dims[l];
    }
    TDYN_REQUIRE(net.dims[0] == net.dims[net.L], "%s: not an autonomous vector field (in %d != out %d)", who, net.dims[0], net.dims[net.L]);
    for (int l = 0; l < net.L; ++l) {
        TDYN_REQUIRE(m->w[l] && m->b[l], "%s: null weight/bias", who);
        net.w[l] = m->w[l]; net.b[l] = m->b[l];
        const int K = net.dims[l], N = net.dims[l + 1];
        net.w_off[l] = woff; woff += N * w_ld(K);
        net.b_off[l] = woff; woff += N;
        net.p_off[l] = poff; poff += N * K + N;
    }
    net.w_floats = (woff + 3) & ~3;
    net.h_floats = hoff;
    net.n_params = poff;
    net.maxdim = maxdim;
    return 0;
}

inline size_t fwd_smem(const Net& n) { return sizeof(float) * (size_t)(n.w_floats + n.h_floats); }
inline size_t adj_smem(const Net& n) { return sizeof(float) * (size_t)(n.w_floats + n.h_floats + 2 * n.maxdim * kLd + n.n_params); }
constexpr size_t kSmemLimit = 220 * 1024;

}  // namespace ffma
}  // namespace tdyn
