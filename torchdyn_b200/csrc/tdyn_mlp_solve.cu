// K1 / K3 (persistent form): a whole `odeint` call for a NARROW MLP vector field in ONE cooperative kernel.
//
//   forward  (adjoint = 0): state x [rows, D]                  -- reference numerics/odeint.py:29-91, :326-408, :411-445
//   backsolve (adjoint = 1): flat A = [x, lam, mu, lam_t]      -- reference numerics/sensitivity.py:55-84 (one interval)
//
// Every CTA owns a fixed set of 32-row tiles for the whole solve.  RK stages never cross rows, so they need no
// grid-wide synchronisation: per stage a CTA forms the stage input of its tiles (reference rounding order), runs the
// MLP (and, for the adjoint, its VJP) from shared-memory weights and writes k_i.  The only global coupling is the
// reference's single step controller: per attempted step ONE grid barrier carries the fp64 partial sums of the
// scaled squared error; every CTA then takes the same accept/reject + dt decision redundantly (fp32 scalar
// arithmetic of numerics/utils.py:57-63 and odeint.py:351-407, incl. the checkpoint clamp / `dt_old - dt` quirk and
// exact float equality for hitting output times).  init_step (utils.py:37-54) runs in-kernel too.
// Adjoint: mu never feeds back into the dynamics and is excluded from the error norm by the reference's seminorm, so
// each CTA integrates ITS batch-partial of mu locally (accepted steps only) and the partials are summed across CTAs
// in a fixed order once at the end; the full-batch d(mu) is only formed for the two init_step norms.
#include "tdyn_mlp_ffma.cuh"

namespace tdyn {
namespace solve {

using ffma::Net;
using ffma::kRows;
using ffma::kRG;
using ffma::kLd;
constexpr int kThreads = ffma::kThreads;
constexpr int kMaxExtra = 6;

struct Plan {
    int n_extra;                       // stages after k1: 6 (dopri5/tsit5), 3 (rk4), 0 (euler)
    int mode[kMaxExtra];
    int nk[kMaxExtra];
    float a[kMaxExtra][TDYN_MAX_STAGES];
    float c[kMaxExtra];
    int n_sol; float bsol[TDYN_MAX_STAGES];
    int n_err; float berr[TDYN_MAX_STAGES];    // n_err == 0 -> fixed step
    int order;
    float safety, min_factor, max_factor;
};

constexpr int kXchgSlots = 64;
constexpr int kVecMax = 16384;         // largest parameter count whose d(mu) can be exchanged in-kernel
constexpr size_t kXchgScalarBytes = sizeof(double) * 4 * kXchgSlots * TDYN_MAX_PEERS;      // XchgSlot = 32 B
constexpr size_t kXchgVecBytes = sizeof(float) * 2 * TDYN_MAX_PEERS * kVecMax;
constexpr size_t kXchgFlagBytes = sizeof(unsigned long long) * 2 * TDYN_MAX_PEERS;
struct XchgSlot {
    double v[3];
    unsigned long long tag;
};

struct Args {
    Net net;
    Plan plan;
    int adjoint;
    float sign;                        // -1: integrate -f(-t, .) (the host already negated t_span)
    float atol, rtol;
    const float* t_span; int n_t;
    int return_all;
    float* Y[2];                       // state ping-pong (flat, reference layout)
    float* K[TDYN_MAX_STAGES];         // stage derivatives (x / lam blocks only)
    float* out; float* t_out; int out_cap;
    double* partials;                  // [3][grid]
    float* mu_partial;                 // [grid][P]
    unsigned int* barrier;
    int* status;                       // n_out, n_attempt, n_accept, nfe, error, final_state_index
    float* trace; int trace_cap;       // [0] = dt0, then (t, dt, ratio) per attempt
    int64_t rows;
    int max_steps;
    // ---- batch sharded over GPUs: per-reduction exchange of the fp64 partial sums by direct peer stores (NVLink)
    int world, rank;
    unsigned int epoch;                // distinguishes launches: slots carry (epoch << 32 | sequence number)
    XchgSlot* peer[TDYN_MAX_PEERS];    // every rank's exchange buffer (peer-mapped); [kXchgSlots][TDYN_MAX_PEERS]
    float* vec_peer[TDYN_MAX_PEERS];   // every rank's vector exchange area: [2 regions][TDYN_MAX_PEERS][kVecMax] floats
    unsigned long long* vflag_peer[TDYN_MAX_PEERS];   // [2 regions][TDYN_MAX_PEERS] tags
    double* gtot;                      // [3] device-local publication of the global totals
};

enum { ST_NOUT = 0, ST_ATTEMPT, ST_ACCEPT, ST_NFE, ST_ERROR, ST_FINAL, ST_COUNT };

__device__ __forceinline__ void grid_sync(unsigned int* bar) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int n = gridDim.x;
        const unsigned int old = atomicAdd(bar, 1u);
        const unsigned int target = (old / n + 1u) * n;
        while (true) {
            unsigned int v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if (v >= target) break;
            __nanosleep(32);
        }
        __threadfence();
    }
    __syncthreads();
}

// sum `v` over the block; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* sm) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kThreads / 32; ++w) s += sm[w];
    return s;
}

// stage input for flat element i: reference association order (ode.py:148-155, Appendix A.2 of SURVEY.md)
__device__ __forceinline__ float combine_at(const float* y, float* const* K, const int* kidx, int nk, int mode,
                                            const float* coef, float dt, int64_t i) {
    const float x = y[i];
    if (nk == 0) return x;
    if (mode == TDYN_MODE_CHAIN) {
        float v = x;
        for (int j = 0; j < nk; ++j) v = add_rn(v, mul_rn(mul_rn(dt, coef[j]), K[kidx[j]][i]));
        return v;
    }
    float s = mul_rn(coef[0], K[kidx[0]][i]);
    for (int j = 1; j < nk; ++j) s = add_rn(s, mul_rn(coef[j], K[kidx[j]][i]));
    return add_rn(x, mul_rn(dt, s));
}

template <int ACT>
__global__ void __launch_bounds__(kThreads, 1) ode_solve_kernel(const Args A) {
    extern __shared__ __align__(16) float sm[];
    const Net& net = A.net;
    const Plan& pl = A.plan;
    float* w_s = sm;
    float* h_s = w_s + net.w_floats;
    float* d_s = h_s + net.h_floats;                                   // adjoint only: 2 x [maxdim][kLd]
    float* g_k1 = d_s + (A.adjoint ? 2 * net.maxdim * kLd : 0);        // adjoint only: 4 x [P]
    float* g_kl = g_k1 + net.n_params;
    float* g_step = g_kl + net.n_params;
    float* mu_acc = g_step + net.n_params;
    __shared__ double red_sm[kThreads / 32];

    ffma::load_weights(net, w_s);
    const int D = net.dims[0];
    const int P = net.n_params;
    const int64_t n = A.rows * D;                       // elements of one state block
    const int64_t tiles = (A.rows + kRows - 1) / kRows;
    const int n_norm_blocks = A.adjoint ? 2 : 1;        // blocks covered by the error norm (seminorm: x and lam)
    if (A.adjoint)
        for (int i = threadIdx.x; i < 4 * P; i += kThreads) g_k1[i] = 0.f;
    __syncthreads();

    int cur = 0;
    int kidx[TDYN_MAX_STAGES] = {0, 1, 2, 3, 4, 5, 6};

    // Element-wise passes touch ONLY the rows of this CTA's own tiles (x block and, for the adjoint, the lam block):
    // a CTA then never reads what another CTA wrote, so stages and updates need no grid-wide synchronisation.
    const int tile_elems = kRows * D;
    const int64_t my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto for_own = [&](auto&& body) {
        for (int64_t q = threadIdx.x; q < my_tiles * tile_elems; q += kThreads) {
            const int64_t m = q / tile_elems;
            const int64_t e = ((int64_t)blockIdx.x + m * gridDim.x) * tile_elems + (q % tile_elems);
            if (e < n) {
                body(e);
                if (A.adjoint) body(e + n);
            }
        }
    };

    // ---- evaluate the vector field of stage `so` (output K[kidx[so]]) at the stage input combine(Y[cur], ...)
    // gmode: 0 = none, 1 = into g_k1, 2 = into g_kl (last stage), 3 = weighted into g_step
    auto eval_stage = [&](int so, int nk, int mode, const float* coef, float dt, int gmode, float gweight) {
        float* gdst = gmode == 1 ? g_k1 : (gmode == 2 ? g_kl : g_step);
        if (gmode != 3) gweight = 1.f;
        if (A.adjoint && (gmode == 1 || gmode == 2))
            for (int i = threadIdx.x; i < P; i += kThreads) gdst[i] = 0.f;
        const float* y = A.Y[cur];
        float* kout = A.K[kidx[so]];
        for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
            __syncthreads();
            const int64_t row0 = t * kRows;
            for (int i = threadIdx.x; i < kRows * D; i += kThreads) {
                const int r = i / D, d = i % D;
                const int64_t gr = row0 + r;
                h_s[net.h_off[0] + d * kLd + r] = gr < A.rows ? combine_at(y, A.K, kidx, nk, mode, coef, dt, gr * D + d) : 0.f;
            }
            if (A.adjoint) {
                // lam block of the stage input: K buffers hold [x-block | lam-block], so offset every pointer by n
                for (int i = threadIdx.x; i < kRows * D; i += kThreads) {
                    const int r = i / D, d = i % D;
                    const int64_t gr = row0 + r;
                    float lv = 0.f;
                    if (gr < A.rows) {
                        const int64_t e = n + gr * D + d;
                        lv = -combine_at(y, A.K, kidx, nk, mode, coef, dt, e);
                    }
                    d_s[d * kLd + r] = lv;        // delta_L = -lam
                }
            }
            __syncthreads();
            ffma::forward_tile<ACT>(net, w_s, h_s);
            ffma::store_tile(kout, h_s + net.h_off[net.L], D, row0, A.rows, A.sign);
            if (A.adjoint) {
                float* dfin = ffma::backward_tile<ACT>(net, w_s, h_s, d_s, d_s + net.maxdim * kLd,
                                                       [&](int idx, float v) { gdst[idx] = fmaf(gweight, v, gdst[idx]); });
                ffma::store_tile(kout + n, dfin, D, row0, A.rows, A.sign);
            }
        }
        __syncthreads();
    };

    // ---- grid-wide sum of up to 3 per-thread fp64 values (fixed order) -> every thread gets the totals.
    // Sharded over GPUs (world > 1): CTA 0 then stores this GPU's totals into EVERY rank's exchange buffer through the
    // peer mapping (NVLink stores, release at system scope), spins until all ranks' slots of this sequence number have
    // arrived in its own buffer, and adds them in rank order -- the reference's single global controller, without
    // leaving the kernel and without a host round trip.
    __shared__ double tot_sm[3];
    unsigned int xseq = 0;
    auto grid_sum = [&](double v0, double v1, double v2, int nv) {
        const double s0 = block_sum(v0, red_sm);
        const double s1 = nv > 1 ? block_sum(v1, red_sm) : 0.0;
        const double s2 = nv > 2 ? block_sum(v2, red_sm) : 0.0;
        if (threadIdx.x == 0) {
            A.partials[0 * gridDim.x + blockIdx.x] = s0;
            A.partials[1 * gridDim.x + blockIdx.x] = s1;
            A.partials[2 * gridDim.x + blockIdx.x] = s2;
        }
        grid_sync(A.barrier);
        if (threadIdx.x < 3) {
            double s = 0.0;
            for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(A.partials + threadIdx.x * gridDim.x + b);
            tot_sm[threadIdx.x] = s;
        }
        if (A.world > 1) {
            ++xseq;
            __syncthreads();
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                const unsigned long long tag = ((unsigned long long)A.epoch << 32) | xseq;
                const int slot = xseq % kXchgSlots;
                for (int r = 0; r < A.world; ++r) {
                    XchgSlot* dst = A.peer[r] + slot * TDYN_MAX_PEERS + A.rank;
                    dst->v[0] = tot_sm[0]; dst->v[1] = tot_sm[1]; dst->v[2] = tot_sm[2];
                    __threadfence_system();
                    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&dst->tag), "l"(tag) : "memory");
                }
                double g0 = 0.0, g1 = 0.0, g2 = 0.0;
                for (int q = 0; q < A.world; ++q) {
                    const XchgSlot* src = A.peer[A.rank] + slot * TDYN_MAX_PEERS + q;
                    unsigned long long seen = 0;
                    for (long long spin = 0; spin < (1ll << 27); ++spin) {
                        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(&src->tag) : "memory");
                        if (seen == tag) break;
                        __nanosleep(64);
                    }
                    if (seen != tag) { g0 = g1 = g2 = __longlong_as_double(0x7ff8000000000000ll); break; }   // NaN -> loop exits with an error
                    g0 += *(volatile const double*)&src->v[0]; g1 += *(volatile const double*)&src->v[1];
                    g2 += *(volatile const double*)&src->v[2];
                }
                A.gtot[0] = g0; A.gtot[1] = g1; A.gtot[2] = g2;
                __threadfence();
            }
            grid_sync(A.barrier);
            if (threadIdx.x < 3) tot_sm[threadIdx.x] = __ldcg(A.gtot + threadIdx.x);
        }
        grid_sync(A.barrier);          // partials may be overwritten by the next reduction only after everyone has read
    };

    // ---- adjoint init_step only: per-CTA d(mu) partials to global region `reg` (0: k1, 1: trial point)
    auto publish_g = [&](const float* g, int reg, float scale) {
        float* dst = A.mu_partial + ((size_t)reg * gridDim.x + blockIdx.x) * P;
        for (int i = threadIdx.x; i < P; i += kThreads) dst[i] = scale * g[i];
        grid_sync(A.barrier);
    };
    auto full_g = [&](int reg, int i) {
        float s = 0.f;
        const float* src = A.mu_partial + (size_t)reg * gridDim.x * P + i;
        for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(src + (size_t)b * P);
        return s;
    };
    // full-batch d(mu) of region `reg` for index i: the cross-CTA sum and, when sharded, the cross-GPU sum.
    // exchange_g() must have been called for the region (collective over the grid and over the ranks); afterwards
    // gvec(reg, i) is valid for the indices i of THIS thread's slice (i == blockIdx*kThreads + threadIdx mod grid*kThreads).
    float* gvec_store = A.mu_partial + (size_t)2 * gridDim.x * P;        // [2][P] device-local results
    unsigned int vseq = 0;
    auto exchange_g = [&](int reg) {
        if (A.world == 1) {
            for (int i = blockIdx.x * kThreads + threadIdx.x; i < P; i += gridDim.x * kThreads) gvec_store[reg * P + i] = full_g(reg, i);
            return;
        }
        ++vseq;
        const int vr = vseq & 1;
        for (int i = blockIdx.x * kThreads + threadIdx.x; i < P; i += gridDim.x * kThreads) {
            const float sv = full_g(reg, i);
            for (int r = 0; r < A.world; ++r) A.vec_peer[r][((size_t)vr * TDYN_MAX_PEERS + A.rank) * kVecMax + i] = sv;
        }
        __threadfence_system();
        grid_sync(A.barrier);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            const unsigned long long tag = ((unsigned long long)A.epoch << 32) | vseq;
            for (int r = 0; r < A.world; ++r)
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(A.vflag_peer[r] + vr * TDYN_MAX_PEERS + A.rank), "l"(tag) : "memory");
            for (int q = 0; q < A.world; ++q) {
                unsigned long long seen = 0;
                for (long long spin = 0; spin < (1ll << 27); ++spin) {
                    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(A.vflag_peer[A.rank] + vr * TDYN_MAX_PEERS + q) : "memory");
                    if (seen == tag) break;
                    __nanosleep(64);
                }
                if (seen != tag) A.gtot[3] = __longlong_as_double(0x7ff8000000000000ll);     // poison: reported as an error
            }
            __threadfence();
        }
        grid_sync(A.barrier);
        const float* own = A.vec_peer[A.rank] + (size_t)vr * TDYN_MAX_PEERS * kVecMax;
        for (int i = blockIdx.x * kThreads + threadIdx.x; i < P; i += gridDim.x * kThreads) {
            float sv = 0.f;
            for (int q = 0; q < A.world; ++q) sv += *(volatile const float*)(own + (size_t)q * kVecMax + i);
            gvec_store[reg * P + i] = sv;
        }
    };
    auto gvec = [&](int reg, int i) { return gvec_store[reg * P + i]; };

    const int nblk = A.adjoint ? 2 : 1;
    const int64_t n_state = nblk * n;                          // elements of the x (and lam) blocks
    int64_t count_norm = n_norm_blocks * n;                    // error-norm count (seminorm)
    int64_t count_all = n_state + (A.adjoint ? P + 1 : 0);
    const int64_t save_stride = count_all;                     // local size of one saved state
    if (A.world > 1) {                                         // global element counts (shards may be unequal)
        const bool one = blockIdx.x == 0 && threadIdx.x == 0;
        const int64_t mine_all = n_state + ((A.adjoint && A.rank == 0) ? P + 1 : 0);     // replicated mu, lam_t: once
        grid_sum(one ? (double)count_norm : 0.0, one ? (double)mine_all : 0.0, 0.0, 2);
        count_norm = (int64_t)(tot_sm[0] + 0.5);
        count_all = (int64_t)(tot_sm[1] + 0.5);
    }
    const float* mu0 = A.adjoint ? A.Y[0] + 2 * n : nullptr;   // incoming mu (full) and lam_t
    const float t0 = __ldg(A.t_span);
    const float T = __ldg(A.t_span + A.n_t - 1);
    int nfe = 0;

    // k1 = f_(t0, y0)
    eval_stage(0, 0, 0, nullptr, 0.f, 1, 1.f);
    ++nfe;

    float dt;
    if (pl.n_err > 0) {
        // ================= init_step (utils.py:37-54); the un-negated f is used for the trial point (reference quirk):
        // f(t0 + h0, y0 + h0*f0_) with f0_ = k1 = sign*f: the trial STATE uses k1, the trial FIELD is +f (sign dropped)
        double a0 = 0.0, a1 = 0.0;
        for_own([&](int64_t e) {
            const float xv = A.Y[cur][e], fv = A.K[kidx[0]][e];
            const float scale = add_rn(A.atol, mul_rn(fabsf(xv), A.rtol));
            const float q0 = div_rn(xv, scale), q1 = div_rn(fv, scale);
            a0 += (double)mul_rn(q0, q0); a1 += (double)mul_rn(q1, q1);
        });
        if (A.adjoint) {
            publish_g(g_k1, 0, A.sign);                  // mu part of f0 = f_(t0, A) = sign * (-lam^T df/dtheta)
            exchange_g(0);
            // the mu block is replicated across ranks: it enters the global norm once (rank 0)
            for (int i = blockIdx.x * kThreads + threadIdx.x; i < (A.rank == 0 ? P : 0); i += gridDim.x * kThreads) {
                const float mv = mu0[i], fv = gvec(0, i);
                const float scale = add_rn(A.atol, mul_rn(fabsf(mv), A.rtol));
                const float q0 = div_rn(mv, scale), q1 = div_rn(fv, scale);
                a0 += (double)mul_rn(q0, q0); a1 += (double)mul_rn(q1, q1);
            }
            if (blockIdx.x == 0 && threadIdx.x == 0 && A.rank == 0) {
                const float lt = mu0[P];
                const float q0 = div_rn(lt, add_rn(A.atol, mul_rn(fabsf(lt), A.rtol)));
                a0 += (double)mul_rn(q0, q0);          // d(lam_t)/dt = 0 for an autonomous field
            }
        }
        grid_sum(a0, a1, 0.0, 2);
        const float d0 = sqrtf((float)(tot_sm[0] / (double)count_all)), d1 = sqrtf((float)(tot_sm[1] / (double)count_all));
        float h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : div_rn(mul_rn(0.01f, d0), d1);
        // trial evaluation: state y0 + h0*k1 (CHAIN with coefficient 1), field WITHOUT the sign -> K[kidx[1]] = sign*f;
        // the reference needs +f there, i.e. sign * K.
        const float one = 1.f;
        eval_stage(1, 1, TDYN_MODE_CHAIN, &one, h0, 2, 1.f);
        ++nfe;
        double a2 = 0.0;
        for_own([&](int64_t e) {
            const float xv = A.Y[cur][e], f0 = A.K[kidx[0]][e], f1 = mul_rn(A.sign, A.K[kidx[1]][e]);
            const float scale = add_rn(A.atol, mul_rn(fabsf(xv), A.rtol));
            const float q = div_rn(sub_rn(f1, f0), scale);
            a2 += (double)mul_rn(q, q);
        });
        if (A.adjoint) {
            publish_g(g_kl, 1, 1.f);                      // un-negated field at the trial point
            exchange_g(1);
            for (int i = blockIdx.x * kThreads + threadIdx.x; i < (A.rank == 0 ? P : 0); i += gridDim.x * kThreads) {
                const float mv = mu0[i];
                const float scale = add_rn(A.atol, mul_rn(fabsf(mv), A.rtol));
                const float q = div_rn(sub_rn(gvec(1, i), gvec(0, i)), scale);
                a2 += (double)mul_rn(q, q);
            }
            // lam_t: f1 - f0 = 0
        }
        grid_sum(a2, 0.0, 0.0, 1);
        const float d2 = div_rn(sqrtf((float)(tot_sm[0] / (double)count_all)), h0);
        float h1;
        if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, mul_rn(h0, 1e-3f));
        else {
            const float base = mul_rn(div_rn(1.f, fmaxf(d1, d2)), 0.01f);
            h1 = (float)pow((double)base, 1.0 / (double)(pl.order + 1));
        }
        dt = fminf(mul_rn(100.f, h0), h1);
        if (blockIdx.x == 0 && threadIdx.x == 0 && A.trace_cap >= 8) {     // init_step internals, for diagnostics
            A.trace[A.trace_cap - 4] = d0; A.trace[A.trace_cap - 3] = d1; A.trace[A.trace_cap - 2] = d2; A.trace[A.trace_cap - 1] = h0;
        }
    } else {
        dt = sub_rn(__ldg(A.t_span + 1), t0);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && A.trace_cap > 0) A.trace[0] = dt;

    // ================= main loop
    float t = t0;
    int ck = 0, n_out = 0, n_attempt = 0, n_accept = 0, err_flag = 0;
    const int n_eval = A.n_t - 1;
    auto save = [&](const float* src, float tv) {
        if (A.out == nullptr) {
        } else if (n_out < A.out_cap) {
            float* dst = A.out + (size_t)n_out * (size_t)save_stride;
            for_own([&](int64_t e) { dst[e] = src[e]; });
            if (blockIdx.x == 0 && threadIdx.x == 0) A.t_out[n_out] = tv;
        } else {
            err_flag = 2;
        }
        ++n_out;
    };
    save(A.Y[cur], t);
    int fixed_step = 1;

    while (true) {
        if (pl.n_err > 0) { if (!(t < T)) break; }
        else { if (fixed_step > n_eval) break; }
        if (n_attempt >= A.max_steps || dt != dt) { err_flag = 1; break; }
        ++n_attempt;
        float dt_keep = 0.f;
        bool ck_flag = false;
        if (pl.n_err > 0) {
            if (add_rn(t, dt) > T) dt = sub_rn(T, t);
            if (ck < n_eval && add_rn(t, dt) > __ldg(A.t_span + 1 + ck)) {
                dt_keep = dt; ck_flag = true;
                dt = sub_rn(__ldg(A.t_span + 1 + ck), t);
            }
        }
        // ---- stages 2..s (k1 is carried: FSAL for the embedded pairs; recomputed for fixed-step methods)
        if (pl.n_err == 0 && n_attempt > 1) { eval_stage(0, 0, 0, nullptr, 0.f, 1, 1.f); ++nfe; }
        if (A.adjoint)
            for (int i = threadIdx.x; i < P; i += kThreads) g_step[i] = 0.f;
        for (int s = 0; s < pl.n_extra; ++s) {
            const int so = s + 1;
            const bool last = so == pl.n_extra;
            const float bw = so < pl.n_sol ? pl.bsol[so] : 0.f;
            // last stage of an FSAL pair: its d(mu) becomes next step's k1 part (b = 0 there), others go to g_step
            eval_stage(so, pl.nk[s], pl.mode[s], pl.a[s], dt, (last && pl.n_err > 0) ? 2 : 3, bw);
            ++nfe;
        }
        // ---- x_new, err, scaled squared error
        float* ynew = A.Y[cur ^ 1];
        double acc = 0.0;
        for_own([&](int64_t e) {
            const float xv = A.Y[cur][e];
            float s = mul_rn(pl.bsol[0], A.K[kidx[0]][e]);
            for (int j = 1; j < pl.n_sol; ++j) s = add_rn(s, mul_rn(pl.bsol[j], A.K[kidx[j]][e]));
            const float xn = add_rn(xv, mul_rn(dt, s));
            ynew[e] = xn;
            if (pl.n_err > 0) {
                float ee = mul_rn(pl.berr[0], A.K[kidx[0]][e]);
                for (int j = 1; j < pl.n_err; ++j) ee = add_rn(ee, mul_rn(pl.berr[j], A.K[kidx[j]][e]));
                const float er = mul_rn(dt, ee);
                const float den = add_rn(A.atol, mul_rn(A.rtol, fmaxf(fabsf(xv), fabsf(xn))));
                const float r = div_rn(er, den);
                if (e < (int64_t)n_norm_blocks * n) acc += (double)mul_rn(r, r);
            }
        });
        float ratio = 0.f;
        bool accept = true;
        if (pl.n_err > 0) {
            grid_sum(acc, 0.0, 0.0, 1);
            ratio = sqrtf((float)(tot_sm[0] / (double)count_norm));
            accept = ratio <= 1.f;
        } else {
            __syncthreads();
        }
        if (blockIdx.x == 0 && threadIdx.x == 0 && 3 * n_attempt < A.trace_cap) {
            A.trace[3 * n_attempt - 2] = t; A.trace[3 * n_attempt - 1] = dt; A.trace[3 * n_attempt] = ratio;
        }
        if (accept) {
            ++n_accept;
            const float tn = add_rn(t, dt);
            if (A.adjoint)
                for (int i = threadIdx.x; i < P; i += kThreads)
                    mu_acc[i] = fmaf(dt, fmaf(pl.bsol[0], g_k1[i], g_step[i]), mu_acc[i]);
            if (pl.n_err > 0) {
                const bool hit = ck < n_eval && tn == __ldg(A.t_span + 1 + ck);
                if (hit || A.return_all) { save(ynew, tn); if (hit) ++ck; }
                // FSAL: k1 <- k_last (and its d(mu) partial)
                const int tmp = kidx[0]; kidx[0] = kidx[pl.n_extra]; kidx[pl.n_extra] = tmp;
                if (A.adjoint) {
                    __syncthreads();
                    for (int i = threadIdx.x; i < P; i += kThreads) g_k1[i] = g_kl[i];
                }
            } else {
                // fixed step: every grid point is an output candidate; the host filters by save_at
                save(ynew, tn);
            }
            t = tn;
            cur ^= 1;
        }
        if (pl.n_err > 0) {
            if (ck_flag) dt = sub_rn(dt_keep, dt);
            // adapt_step (utils.py:57-63)
            if (ratio == 0.f) dt = mul_rn(dt, pl.max_factor);
            else {
                const float lo = ratio < 1.f ? 1.f : pl.min_factor;
                const float pw = (float)pow((double)ratio, (double)div_rn(1.f, (float)pl.order));
                const float cand = div_rn(pl.safety, pw);
                const float fac = (cand != cand) ? cand : fminf(pl.max_factor, fmaxf(cand, lo));
                dt = mul_rn(dt, fac);
            }
        } else {
            ++fixed_step;
            if (fixed_step <= n_eval) dt = sub_rn(__ldg(A.t_span + fixed_step), t);
        }
        __syncthreads();
    }

    // ================= epilogue: final state index, mu reduction, status
    if (A.adjoint) {
        __syncthreads();
        for (int i = threadIdx.x; i < P; i += kThreads) A.mu_partial[(size_t)blockIdx.x * P + i] = mu_acc[i];
        grid_sync(A.barrier);
        // every CTA reduces a slice; result = incoming mu + sign * sum (d(mu) carries the sign of the reversed field)
        float* yfin = A.Y[cur];
        for (int i = blockIdx.x * kThreads + threadIdx.x; i < P; i += gridDim.x * kThreads) {
            float s = 0.f;
            for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(A.mu_partial + (size_t)b * P + i);
            yfin[2 * n + i] = add_rn(mu0[i], mul_rn(A.sign, s));
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) yfin[2 * n + P] = mu0[P];
    }
    if (A.world > 1 && A.gtot[3] != A.gtot[3]) err_flag = 3;      // a peer never showed up in a vector exchange
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        A.status[ST_NOUT] = n_out; A.status[ST_ATTEMPT] = n_attempt; A.status[ST_ACCEPT] = n_accept;
        A.status[ST_NFE] = nfe; A.status[ST_ERROR] = err_flag; A.status[ST_FINAL] = cur;
    }
}


// =====================================================================================================================
// K4: one interval of the INTERPOLATED adjoint (reference numerics/sensitivity.py:130-152) in one cooperative kernel.
// State A = [lam (rows*D), mu (P)]; x(t) comes from the natural cubic spline of the forward trajectory (sol, kd, two_c,
// three_d as built by tdyn_spline_build; piece = clamp(#knots < t - 1, 0, M-2), like CubicSpline.evaluate).  Unlike the
// backsolve adjoint there is NO seminorm: mu is inside the error norm, so every attempted step reduces the CTAs'
// batch-partials of sum_j b_j dmu_j and sum_j berr_j dmu_j (two P-vectors) across the grid before the norm.
struct InterpArgs {
    Net net;
    Plan plan;
    float atol, rtol;
    float t_start, t_end;              // negated span: integrate from -t_hi up to -t_lo
    float* Y[2];                       // [lam | mu]
    float* K[TDYN_MAX_STAGES];         // lam blocks
    const float* sol; const float* kd; const float* two_c; const float* three_d; const float* knots; int n_knots;
    double* partials; float* mu_partial; unsigned int* barrier;
    int* status; float* trace; int trace_cap;
    int64_t rows; int max_steps;
};

template <int ACT>
__global__ void __launch_bounds__(kThreads, 1) interp_adjoint_kernel(const InterpArgs A) {
    extern __shared__ __align__(16) float sm[];
    const Net& net = A.net;
    const Plan& pl = A.plan;
    float* w_s = sm;
    float* h_s = w_s + net.w_floats;
    float* d_s = h_s + net.h_floats;
    float* g_k1 = d_s + 2 * net.maxdim * kLd;          // d(mu) batch-partial of the current k1 (raw: delta_L = -lam)
    float* g_kl = g_k1 + net.n_params;                 // ... of the last stage (next step's k1 under FSAL)
    float* s_b = g_kl + net.n_params;                  // sum_j>=1 bsol_j * dmu_j partial
    float* s_e = s_b + net.n_params;                   // sum_j>=1 berr_j * dmu_j partial
    __shared__ double red_sm[kThreads / 32];
    __shared__ double tot_sm[3];
    constexpr float sign = -1.f;                       // reversed time: f_(t, A) = -dyn(-t, A)

    ffma::load_weights(net, w_s);
    const int D = net.dims[0];
    const int P = net.n_params;
    const int64_t n = A.rows * D;
    const int64_t tiles = (A.rows + kRows - 1) / kRows;
    for (int i = threadIdx.x; i < 4 * P; i += kThreads) g_k1[i] = 0.f;
    __syncthreads();
    int cur = 0;
    int kidx[TDYN_MAX_STAGES] = {0, 1, 2, 3, 4, 5, 6};
    const int tile_elems = kRows * D;
    const int64_t my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto for_own = [&](auto&& body) {
        for (int64_t q = threadIdx.x; q < my_tiles * tile_elems; q += kThreads) {
            const int64_t m = q / tile_elems;
            const int64_t e = ((int64_t)blockIdx.x + m * gridDim.x) * tile_elems + (q % tile_elems);
            if (e < n) body(e);
        }
    };
    // mu indices owned by this thread (norm / update slices)
    auto for_mu = [&](auto&& body) {
        for (int i = blockIdx.x * kThreads + threadIdx.x; i < P; i += gridDim.x * kThreads) body(i);
    };

    // gmode: 1 -> g_k1, 2 -> g_kl (both raw, weight 1), 3 -> s_b += wb * g, s_e += we * g
    auto eval_stage = [&](int so, int nk, int mode, const float* coef, float dt, float tau, int gmode, float wb, float we) {
        float* gdst = gmode == 1 ? g_k1 : g_kl;
        if (gmode != 3)
            for (int i = threadIdx.x; i < P; i += kThreads) gdst[i] = 0.f;
        // spline piece for the ACTUAL time tau (CubicSpline.evaluate: bucketize(t) - 1, clamped)
        int cnt = 0;
        for (int m = 0; m < A.n_knots; ++m) cnt += (__ldg(A.knots + m) < tau) ? 1 : 0;
        int idx = cnt - 1;
        idx = idx < 0 ? 0 : (idx > A.n_knots - 2 ? A.n_knots - 2 : idx);
        const float sfrac = sub_rn(tau, __ldg(A.knots + idx));
        const float* pa = A.sol + (size_t)idx * n;
        const float* pb = A.kd + (size_t)idx * n;
        const float* pc = A.two_c + (size_t)idx * n;
        const float* pd = A.three_d + (size_t)idx * n;
        const float* y = A.Y[cur];
        float* kout = A.K[kidx[so]];
        for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
            __syncthreads();
            const int64_t row0 = t * kRows;
            for (int i = threadIdx.x; i < kRows * D; i += kThreads) {
                const int r = i / D, d = i % D;
                const int64_t gr = row0 + r;
                float xv = 0.f, lv = 0.f;
                if (gr < A.rows) {
                    const int64_t e = gr * D + d;
                    float inner = add_rn(mul_rn(0.5f, __ldg(pc + e)), div_rn(mul_rn(__ldg(pd + e), sfrac), 3.f));
                    inner = add_rn(__ldg(pb + e), mul_rn(inner, sfrac));
                    xv = add_rn(__ldg(pa + e), mul_rn(inner, sfrac));
                    lv = -combine_at(y, A.K, kidx, nk, mode, coef, dt, e);
                }
                h_s[net.h_off[0] + d * kLd + r] = xv;
                d_s[d * kLd + r] = lv;                    // delta_L = -lam
            }
            __syncthreads();
            ffma::forward_tile<ACT>(net, w_s, h_s);
            float* dfin = ffma::backward_tile<ACT>(net, w_s, h_s, d_s, d_s + net.maxdim * kLd, [&](int idx, float v) {
                if (gmode == 3) { s_b[idx] = fmaf(wb, v, s_b[idx]); s_e[idx] = fmaf(we, v, s_e[idx]); }
                else gdst[idx] += v;
            });
            ffma::store_tile(kout, dfin, D, row0, A.rows, sign);
        }
        __syncthreads();
    };

    auto grid_sum = [&](double v0, double v1, double v2, int nv) {
        const double s0 = block_sum(v0, red_sm);
        const double s1 = nv > 1 ? block_sum(v1, red_sm) : 0.0;
        const double s2 = nv > 2 ? block_sum(v2, red_sm) : 0.0;
        if (threadIdx.x == 0) {
            A.partials[0 * gridDim.x + blockIdx.x] = s0;
            A.partials[1 * gridDim.x + blockIdx.x] = s1;
            A.partials[2 * gridDim.x + blockIdx.x] = s2;
        }
        grid_sync(A.barrier);
        if (threadIdx.x < 3) {
            double s = 0.0;
            for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(A.partials + threadIdx.x * gridDim.x + b);
            tot_sm[threadIdx.x] = s;
        }
        grid_sync(A.barrier);
    };
    // publish two per-CTA P-vectors (region r0 <- scale0*v0, region r1 <- scale1*v1), then grid barrier
    auto publish2 = [&](const float* v0, float sc0, const float* v1, float sc1) {
        float* d0 = A.mu_partial + ((size_t)0 * gridDim.x + blockIdx.x) * P;
        float* d1 = A.mu_partial + ((size_t)1 * gridDim.x + blockIdx.x) * P;
        for (int i = threadIdx.x; i < P; i += kThreads) { d0[i] = sc0 * v0[i]; d1[i] = v1 ? sc1 * v1[i] : 0.f; }
        grid_sync(A.barrier);
    };
    auto full_g = [&](int reg, int i) {
        float sv = 0.f;
        const float* src = A.mu_partial + (size_t)reg * gridDim.x * P + i;
        for (unsigned int b = 0; b < gridDim.x; ++b) sv += __ldcg(src + (size_t)b * P);
        return sv;
    };
    float* keep = A.mu_partial + (size_t)2 * gridDim.x * P;       // [P] device-local: full signed d(mu) of k1 at t0

    const int64_t count_all = n + P;
    const float t0 = A.t_start, T = A.t_end;
    int nfe = 0;
    const float* mu_in = A.Y[0] + n;

    eval_stage(0, 0, 0, nullptr, 0.f, -t0, 1, 0.f, 0.f);            // k1 = f_(t0, A) = -dyn(-t0, A)
    ++nfe;
    // ---- init_step (utils.py:37-54); trial point uses the UN-negated dynamics at time (t0 + h0): reference quirk
    float dt;
    {
        double a0 = 0.0, a1 = 0.0;
        for_own([&](int64_t e) {
            const float xv = A.Y[cur][e], fv = A.K[kidx[0]][e];
            const float scale = add_rn(A.atol, mul_rn(fabsf(xv), A.rtol));
            const float q0 = div_rn(xv, scale), q1 = div_rn(fv, scale);
            a0 += (double)mul_rn(q0, q0); a1 += (double)mul_rn(q1, q1);
        });
        publish2(g_k1, sign, nullptr, 0.f);
        for_mu([&](int i) {
            const float mv = mu_in[i], fv = full_g(0, i);
            keep[i] = fv;
            const float scale = add_rn(A.atol, mul_rn(fabsf(mv), A.rtol));
            const float q0 = div_rn(mv, scale), q1 = div_rn(fv, scale);
            a0 += (double)mul_rn(q0, q0); a1 += (double)mul_rn(q1, q1);
        });
        grid_sum(a0, a1, 0.0, 2);
        const float d0 = sqrtf((float)(tot_sm[0] / (double)count_all)), d1 = sqrtf((float)(tot_sm[1] / (double)count_all));
        const float h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : div_rn(mul_rn(0.01f, d0), d1);
        const float one = 1.f;
        eval_stage(1, 1, TDYN_MODE_CHAIN, &one, h0, add_rn(t0, h0), 2, 0.f, 0.f);
        ++nfe;
        double a2 = 0.0;
        for_own([&](int64_t e) {
            const float xv = A.Y[cur][e], f0 = A.K[kidx[0]][e], f1 = mul_rn(sign, A.K[kidx[1]][e]);
            const float scale = add_rn(A.atol, mul_rn(fabsf(xv), A.rtol));
            const float q = div_rn(sub_rn(f1, f0), scale);
            a2 += (double)mul_rn(q, q);
        });
        publish2(g_kl, 1.f, nullptr, 0.f);
        for_mu([&](int i) {
            const float mv = mu_in[i];
            const float scale = add_rn(A.atol, mul_rn(fabsf(mv), A.rtol));
            const float q = div_rn(sub_rn(full_g(0, i), keep[i]), scale);
            a2 += (double)mul_rn(q, q);
        });
        grid_sum(a2, 0.0, 0.0, 1);
        const float d2 = div_rn(sqrtf((float)(tot_sm[0] / (double)count_all)), h0);
        float h1;
        if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, mul_rn(h0, 1e-3f));
        else h1 = (float)pow((double)mul_rn(div_rn(1.f, fmaxf(d1, d2)), 0.01f), 1.0 / (double)(pl.order + 1));
        dt = fminf(mul_rn(100.f, h0), h1);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && A.trace_cap > 0) A.trace[0] = dt;

    float t = t0;
    int ck = 0, n_attempt = 0, n_accept = 0, err_flag = 0;
    while (t < T) {
        if (n_attempt >= A.max_steps || dt != dt) { err_flag = 1; break; }
        ++n_attempt;
        float dt_keep = 0.f;
        bool ck_flag = false;
        if (add_rn(t, dt) > T) dt = sub_rn(T, t);
        if (ck < 1 && add_rn(t, dt) > T) { dt_keep = dt; ck_flag = true; dt = sub_rn(T, t); }
        // weighted d(mu) sums of this attempt start from k1's contribution
        for (int i = threadIdx.x; i < P; i += kThreads) { s_b[i] = pl.bsol[0] * g_k1[i]; s_e[i] = pl.berr[0] * g_k1[i]; }
        for (int s = 0; s < pl.n_extra; ++s) {
            const int so = s + 1;
            const bool last = so == pl.n_extra;
            const float wb = so < pl.n_sol ? pl.bsol[so] : 0.f;
            const float we = so < pl.n_err ? pl.berr[so] : 0.f;
            const float ts = add_rn(t, mul_rn(pl.c[s], dt));
            eval_stage(so, pl.nk[s], pl.mode[s], pl.a[s], dt, -ts, last ? 2 : 3, wb, we);
            ++nfe;
            if (last)     // FSAL stage: raw partial kept in g_kl for the next step; its weights (b = 0, berr_7) applied here
                for (int i = threadIdx.x; i < P; i += kThreads) { s_b[i] = fmaf(wb, g_kl[i], s_b[i]); s_e[i] = fmaf(we, g_kl[i], s_e[i]); }
        }
        float* ynew = A.Y[cur ^ 1];
        double acc = 0.0;
        for_own([&](int64_t e) {
            const float xv = A.Y[cur][e];
            float sv = mul_rn(pl.bsol[0], A.K[kidx[0]][e]);
            for (int j = 1; j < pl.n_sol; ++j) sv = add_rn(sv, mul_rn(pl.bsol[j], A.K[kidx[j]][e]));
            const float xn = add_rn(xv, mul_rn(dt, sv));
            ynew[e] = xn;
            float ee = mul_rn(pl.berr[0], A.K[kidx[0]][e]);
            for (int j = 1; j < pl.n_err; ++j) ee = add_rn(ee, mul_rn(pl.berr[j], A.K[kidx[j]][e]));
            const float r = div_rn(mul_rn(dt, ee), add_rn(A.atol, mul_rn(A.rtol, fmaxf(fabsf(xv), fabsf(xn)))));
            acc += (double)mul_rn(r, r);
        });
        // mu block: full-batch weighted sums (signed: reversed field) -> mu_new, err, scaled error
        __syncthreads();
        publish2(s_b, sign, s_e, sign);
        for_mu([&](int i) {
            const float mv = A.Y[cur][n + i];
            const float mn = add_rn(mv, mul_rn(dt, full_g(0, i)));
            ynew[n + i] = mn;
            const float r = div_rn(mul_rn(dt, full_g(1, i)), add_rn(A.atol, mul_rn(A.rtol, fmaxf(fabsf(mv), fabsf(mn)))));
            acc += (double)mul_rn(r, r);
        });
        grid_sum(acc, 0.0, 0.0, 1);
        const float ratio = sqrtf((float)(tot_sm[0] / (double)count_all));
        const bool accept = ratio <= 1.f;
        if (blockIdx.x == 0 && threadIdx.x == 0 && 3 * n_attempt < A.trace_cap) {
            A.trace[3 * n_attempt - 2] = t; A.trace[3 * n_attempt - 1] = dt; A.trace[3 * n_attempt] = ratio;
        }
        if (accept) {
            ++n_accept;
            const float tn = add_rn(t, dt);
            if (ck < 1 && tn == T) ++ck;
            const int tmp = kidx[0]; kidx[0] = kidx[pl.n_extra]; kidx[pl.n_extra] = tmp;
            for (int i = threadIdx.x; i < P; i += kThreads) g_k1[i] = g_kl[i];
            t = tn;
            cur ^= 1;
        }
        if (ck_flag) dt = sub_rn(dt_keep, dt);
        if (ratio == 0.f) dt = mul_rn(dt, pl.max_factor);
        else {
            const float lo = ratio < 1.f ? 1.f : pl.min_factor;
            const float pw = (float)pow((double)ratio, (double)div_rn(1.f, (float)pl.order));
            const float cand = div_rn(pl.safety, pw);
            const float fac = (cand != cand) ? cand : fminf(pl.max_factor, fmaxf(cand, lo));
            dt = mul_rn(dt, fac);
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        A.status[ST_NOUT] = 0; A.status[ST_ATTEMPT] = n_attempt; A.status[ST_ACCEPT] = n_accept;
        A.status[ST_NFE] = nfe; A.status[ST_ERROR] = err_flag; A.status[ST_FINAL] = cur;
    }
}

}  // namespace solve
}  // namespace tdyn

using namespace tdyn;

extern "C" {

int64_t tdyn_mlp_solve_workspace_bytes(const tdyn_mlp_t* m) {
    ffma::Net net;
    if (!m || ffma::make_net(m, net, "tdyn_mlp_solve_workspace_bytes") != 0) return 0;
    // barrier (256 B) | partials 3 x 1024 doubles | mu partials 2 x grid x P floats | t_span copy (4 KB)
    return (int64_t)(256 + 3 * 1024 * sizeof(double) + sizeof(float) * 2 * (size_t)1024 * net.n_params + 4096);
}

int64_t tdyn_comm_buffer_bytes(void) {
    static_assert(sizeof(solve::XchgSlot) == 32, "exchange slot layout");
    return (int64_t)(solve::kXchgScalarBytes + solve::kXchgVecBytes + solve::kXchgFlagBytes);
}

int tdyn_mlp_solve_supported(const tdyn_mlp_t* m, int adjoint) {
    ffma::Net net;
    if (!m || ffma::make_net(m, net, "tdyn_mlp_solve_supported") != 0) return 0;
    if (!(m->act == TDYN_ACT_TANH || m->act == TDYN_ACT_SOFTPLUS || m->act == TDYN_ACT_RELU)) return 0;
    size_t smem = sizeof(float) * (size_t)(net.w_floats + net.h_floats);
    if (adjoint) smem += sizeof(float) * (size_t)(2 * net.maxdim * ffma::kLd + 4 * net.n_params);
    return smem <= ffma::kSmemLimit ? 1 : 0;
}

int tdyn_mlp_solve(const tdyn_mlp_t* m, const tdyn_rk_plan_t* plan, int adjoint, float sign, float atol, float rtol,
                   const float* t_span, int n_t, int return_all, float* y0, float* y1, float* k, float* out, float* t_out,
                   int out_cap, void* workspace, int* status, float* trace, int trace_cap, int64_t rows, int max_steps,
                   const tdyn_comm_t* comm, void* stream) {
    solve::Args A{};
    if (int rc = ffma::make_net(m, A.net, "tdyn_mlp_solve")) return rc;
    TDYN_REQUIRE(tdyn_mlp_solve_supported(m, adjoint), "tdyn_mlp_solve: network not supported (shared memory)");
    TDYN_REQUIRE(plan && t_span && y0 && y1 && k && workspace && status, "tdyn_mlp_solve: null argument");
    TDYN_REQUIRE(n_t >= 2 && n_t <= 1024 && rows > 0, "tdyn_mlp_solve: need 2 <= n_t <= 1024 and rows > 0");
    TDYN_REQUIRE(plan->n_extra >= 0 && plan->n_extra <= solve::kMaxExtra && plan->n_sol >= 1 && plan->n_sol <= TDYN_MAX_STAGES &&
                 plan->n_err >= 0 && plan->n_err <= TDYN_MAX_STAGES, "tdyn_mlp_solve: bad plan");
    solve::Plan& pl = A.plan;
    pl.n_extra = plan->n_extra;
    for (int s = 0; s < solve::kMaxExtra; ++s) {
        pl.mode[s] = plan->mode[s]; pl.nk[s] = plan->nk[s]; pl.c[s] = plan->c[s];
        for (int j = 0; j < TDYN_MAX_STAGES; ++j) pl.a[s][j] = plan->a[s][j];
    }
    pl.n_sol = plan->n_sol; pl.n_err = plan->n_err; pl.order = plan->order;
    for (int j = 0; j < TDYN_MAX_STAGES; ++j) { pl.bsol[j] = plan->bsol[j]; pl.berr[j] = plan->berr[j]; }
    pl.safety = plan->safety; pl.min_factor = plan->min_factor; pl.max_factor = plan->max_factor;
    A.adjoint = adjoint ? 1 : 0; A.sign = sign; A.atol = atol; A.rtol = rtol;
    A.n_t = n_t; A.return_all = return_all;
    const int64_t n = rows * A.net.dims[0];
    const int64_t stride = (adjoint ? 2 : 1) * n;
    A.Y[0] = y0; A.Y[1] = y1;
    for (int j = 0; j < TDYN_MAX_STAGES; ++j) A.K[j] = k + (size_t)j * stride;
    A.out = out; A.t_out = t_out; A.out_cap = out_cap;
    uint8_t* ws = (uint8_t*)workspace;
    A.barrier = (unsigned int*)ws;
    A.partials = (double*)(ws + 256);
    A.mu_partial = (float*)(ws + 256 + 3 * 1024 * sizeof(double));
    float* t_dev = (float*)(ws + 256 + 3 * 1024 * sizeof(double) + sizeof(float) * 2 * (size_t)1024 * A.net.n_params);
    A.t_span = t_dev;                      // `t_span` is a HOST array: staged into the workspace below
    A.status = status; A.trace = trace; A.trace_cap = trace ? trace_cap : 0;
    A.rows = rows; A.max_steps = max_steps;
    A.world = 1; A.rank = 0; A.epoch = 0;
    A.gtot = (double*)(ws + 256 + 3 * 1024 * sizeof(double)) - 4;      // last 4 doubles of the partials area (grid <= 1020)
    if (comm && comm->world > 1) {
        TDYN_REQUIRE(!adjoint || A.net.n_params <= solve::kVecMax,
                     "tdyn_mlp_solve: %d parameters exceed the in-kernel exchange capacity (%d)", A.net.n_params, solve::kVecMax);
        TDYN_REQUIRE(comm->world <= TDYN_MAX_PEERS && comm->rank >= 0 && comm->rank < comm->world, "tdyn_mlp_solve: bad comm");
        A.world = comm->world; A.rank = comm->rank; A.epoch = comm->epoch;
        for (int r = 0; r < comm->world; ++r) {
            TDYN_REQUIRE(comm->peer[r], "tdyn_mlp_solve: comm->peer[%d] is null", r);
            A.peer[r] = (solve::XchgSlot*)comm->peer[r];
            A.vec_peer[r] = (float*)((uint8_t*)comm->peer[r] + solve::kXchgScalarBytes);
            A.vflag_peer[r] = (unsigned long long*)((uint8_t*)comm->peer[r] + solve::kXchgScalarBytes + solve::kXchgVecBytes);
        }
    }

    size_t smem = sizeof(float) * (size_t)(A.net.w_floats + A.net.h_floats);
    if (adjoint) smem += sizeof(float) * (size_t)(2 * A.net.maxdim * ffma::kLd + 4 * A.net.n_params);
    const void* fn = A.net.act == TDYN_ACT_TANH ? (const void*)solve::ode_solve_kernel<TDYN_ACT_TANH>
                   : A.net.act == TDYN_ACT_SOFTPLUS ? (const void*)solve::ode_solve_kernel<TDYN_ACT_SOFTPLUS>
                                                    : (const void*)solve::ode_solve_kernel<TDYN_ACT_RELU>;
    TDYN_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    TDYN_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, solve::kThreads, smem));
    TDYN_REQUIRE(per_sm >= 1, "tdyn_mlp_solve: kernel does not fit on an SM");
    int dev = 0, sms = 0;
    TDYN_CHECK_CUDA(cudaGetDevice(&dev));
    TDYN_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t tiles = (rows + ffma::kRows - 1) / ffma::kRows;
    int64_t grid = (int64_t)per_sm * sms;
    if (grid > tiles) grid = tiles;
    if (grid > 1020) grid = 1020;
    cudaStream_t st = (cudaStream_t)stream;
    TDYN_CHECK_CUDA(cudaMemsetAsync(ws, 0, 256, st));          // barrier counter
    TDYN_CHECK_CUDA(cudaMemsetAsync(A.gtot, 0, 4 * sizeof(double), st));
    TDYN_CHECK_CUDA(cudaMemcpyAsync(t_dev, t_span, sizeof(float) * (size_t)n_t, cudaMemcpyHostToDevice, st));
    void* args[] = {(void*)&A};
    TDYN_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(solve::kThreads), args, smem, st));
    return 0;
}

int tdyn_mlp_solve_interp(const tdyn_mlp_t* m, const tdyn_rk_plan_t* plan, float atol, float rtol, float t_start, float t_end,
                          float* y0, float* y1, float* k, const float* sol, const float* kd, const float* two_c,
                          const float* three_d, const float* knots, int n_knots, void* workspace, int* status, float* trace,
                          int trace_cap, int64_t rows, int max_steps, void* stream) {
    solve::InterpArgs A{};
    if (int rc = ffma::make_net(m, A.net, "tdyn_mlp_solve_interp")) return rc;
    TDYN_REQUIRE(tdyn_mlp_solve_supported(m, 1), "tdyn_mlp_solve_interp: network not supported (shared memory)");
    TDYN_REQUIRE(plan && y0 && y1 && k && sol && kd && two_c && three_d && knots && workspace && status, "tdyn_mlp_solve_interp: null argument");
    TDYN_REQUIRE(n_knots >= 2 && rows > 0 && plan->n_err > 0, "tdyn_mlp_solve_interp: need >= 2 knots, rows > 0 and an embedded pair");
    solve::Plan& pl = A.plan;
    pl.n_extra = plan->n_extra;
    for (int s = 0; s < solve::kMaxExtra; ++s) {
        pl.mode[s] = plan->mode[s]; pl.nk[s] = plan->nk[s]; pl.c[s] = plan->c[s];
        for (int j = 0; j < TDYN_MAX_STAGES; ++j) pl.a[s][j] = plan->a[s][j];
    }
    pl.n_sol = plan->n_sol; pl.n_err = plan->n_err; pl.order = plan->order;
    for (int j = 0; j < TDYN_MAX_STAGES; ++j) { pl.bsol[j] = plan->bsol[j]; pl.berr[j] = plan->berr[j]; }
    pl.safety = plan->safety; pl.min_factor = plan->min_factor; pl.max_factor = plan->max_factor;
    A.atol = atol; A.rtol = rtol; A.t_start = t_start; A.t_end = t_end;
    const int64_t n = rows * A.net.dims[0];
    A.Y[0] = y0; A.Y[1] = y1;
    for (int j = 0; j < TDYN_MAX_STAGES; ++j) A.K[j] = k + (size_t)j * n;
    A.sol = sol; A.kd = kd; A.two_c = two_c; A.three_d = three_d; A.knots = knots; A.n_knots = n_knots;
    uint8_t* ws = (uint8_t*)workspace;
    A.barrier = (unsigned int*)ws;
    A.partials = (double*)(ws + 256);
    A.mu_partial = (float*)(ws + 256 + 3 * 1024 * sizeof(double));
    A.status = status; A.trace = trace; A.trace_cap = trace ? trace_cap : 0;
    A.rows = rows; A.max_steps = max_steps;
    const size_t smem = sizeof(float) * (size_t)(A.net.w_floats + A.net.h_floats + 2 * A.net.maxdim * ffma::kLd + 4 * A.net.n_params);
    const void* fn = A.net.act == TDYN_ACT_TANH ? (const void*)solve::interp_adjoint_kernel<TDYN_ACT_TANH>
                   : A.net.act == TDYN_ACT_SOFTPLUS ? (const void*)solve::interp_adjoint_kernel<TDYN_ACT_SOFTPLUS>
                                                    : (const void*)solve::interp_adjoint_kernel<TDYN_ACT_RELU>;
    TDYN_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    TDYN_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, solve::kThreads, smem));
    TDYN_REQUIRE(per_sm >= 1, "tdyn_mlp_solve_interp: kernel does not fit on an SM");
    int dev = 0, sms = 0;
    TDYN_CHECK_CUDA(cudaGetDevice(&dev));
    TDYN_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t tiles = (rows + ffma::kRows - 1) / ffma::kRows;
    int64_t grid = (int64_t)per_sm * sms;
    if (grid > tiles) grid = tiles;
    if (grid > 1020) grid = 1020;
    cudaStream_t st = (cudaStream_t)stream;
    TDYN_CHECK_CUDA(cudaMemsetAsync(ws, 0, 256, st));
    void* args[] = {(void*)&A};
    TDYN_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(solve::kThreads), args, smem, st));
    return 0;
}

}  // extern "C"
