/*
 * tdyn_b200.h -- C ABI of libtdyn_b200.so, the sm_100a kernel library behind the
 * torchdyn-compatible ODE integration path (torchdyn_b200/).
 *
 * The reference (DiffEqML/torchdyn 1.0.6) has NO native boundary: the hot path is Python over
 * eager ATen ops.  Each entry point below therefore cites the reference *Python* expression(s)
 * it replaces (paths relative to /root/reference/torchdyn/).  INTEGRATION.md shows the ctypes
 * binding a maintainer would add on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is available from
 *     tdyn_last_error() (thread-local).  No exception crosses the ABI.
 *   - all device buffers are BORROWED for the duration of the stream-ordered call; the library
 *     allocates nothing.  Pointer *tables* (k[]) and coefficient arrays are HOST arrays and are
 *     copied into kernel parameters by value.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - kernels run on the CUDA device current on the calling thread (callers in other threads,
 *     e.g. the autograd engine's device thread, must have it set -- PyTorch does).
 *   - arithmetic is fp32 with one IEEE rounding per reference ATen op (no FMA contraction), so
 *     the element-wise results are bit-identical to the reference's; reductions accumulate the
 *     fp32 squares in fp64 with a fixed summation order (deterministic).
 */
#ifndef TDYN_B200_H
#define TDYN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TDYN_API __attribute__((visibility("default")))
#else
#define TDYN_API
#endif

#define TDYN_ABI_VERSION 15
#define TDYN_MAX_STAGES 7
#define TDYN_MAX_PEERS 8

/* combination shapes of the reference steppers (numerics/solvers/ode.py) */
#define TDYN_MODE_CHAIN 0 /* y = x + (dt*a0)*k0 + (dt*a1)*k1 + ...   left to right (ode.py:148,150-153 / :170,172-175) */
#define TDYN_MODE_INNER 1 /* y = x + dt*(a0*k0 + a1*k1 + ...)        (ode.py:68-71,100-103,149,154); x==NULL -> y = dt*(...) (:155) */

TDYN_API int tdyn_version(void);
TDYN_API const char* tdyn_last_error(void);

/* Measurement aid (bench.py): runs `iters` x 64 dependent-chain FP32 FMAs per thread with register operands on a full grid
 * and reports the FLOPs issued; timed by the caller with CUDA events, it gives the FP32 FMA rate this GPU sustains -- the
 * denominator of the FFMA kernels' roofline.  `out`: one device float (never written). */
TDYN_API int tdyn_probe_ffma(float* out, int iters, double* flops_out, void* stream);

/* bytes of zero-initialised device scratch every reducing entry point needs (partials + ticket). */
TDYN_API int64_t tdyn_reduce_workspace_bytes(void);

/* K5. One RK stage input.  Replaces the per-stage expressions of Euler/RungeKutta4/DormandPrince45/
 * Tsitouras45.step (numerics/solvers/ode.py:68-71, :97-104, :145-156, :167-178). */
TDYN_API int tdyn_rk_combine(float* y, const float* x, const float* const* k, const float* coef, int n_k,
                    int mode, float dt, int64_t n, void* stream);

/* K6. Step finish: x_new = x + dt*(sum bsol_j k_j)  (ode.py:154), err = dt*(sum berr_j k_j) (:155, written
 * only when err_out != NULL), scaled error e = err/(atol + rtol*max(|x|,|x_new|)) and sum(e^2) over the first
 * `norm_n` elements (numerics/odeint.py:365-372, numerics/utils.py:33-34).  *sumsq_out (device, fp64) receives
 * the sum; error_ratio = sqrt(sumsq/norm_n) is formed by the caller (after the cross-GPU sum when sharded). */
TDYN_API int tdyn_rk_finish(float* x_new, float* err_out, const float* x, const float* const* k,
                   const float* bsol, int n_sol, const float* berr, int n_err,
                   float dt, float atol, float rtol, int64_t n, int64_t norm_n,
                   double* sumsq_out, void* workspace, void* stream);

/* K9. init_step norms (numerics/utils.py:37-54): with scale = atol + |x0|*rtol writes
 * out[0] = sum((x0/scale)^2), out[1] = sum((f0/scale)^2), out[2] = sum(((f1-f0)/scale)^2) (f1 may be NULL). */
TDYN_API int tdyn_init_norms(const float* x0, const float* f0, const float* f1, float atol, float rtol,
                    int64_t n, double* out3, void* workspace, void* stream);

/* K7. dopri5 4th-order dense output (numerics/odeint.py:380-384, numerics/interpolators.py:29-37,57-63):
 * x_mid = x0 + dt*sum(bmid_i k_i); fit; evaluate at theta = (t - t0)/(t1 - t0) with theta^2..theta^4 given
 * as the running fp32 powers th[0..3] = theta, theta^2, theta^3, theta^4. */
TDYN_API int tdyn_dense_eval(float* out, const float* x0, const float* x1, const float* const* k,
                    const float* bmid, float dt, const float* th, int64_t n, void* stream);

/* K8. Natural cubic spline through the M saved states sol[M, n] of the interpolated adjoint (torchcde 0.2.x
 * algorithm; reference call sites numerics/sensitivity.py:126-127 build, :132 evaluate).  `aux` is a device
 * array of 4*M floats computed by the host from the knot times alone (rows: r=1/h, r^2, Thomas multipliers w,
 * pivots d'); outputs are the knot derivatives kd[M, n] and the per-piece two_c, three_d [M-1, n] (a = sol).
 * With 2 knots the spline degenerates to the chord (aux unused, t1_minus_t0 used). */
TDYN_API int tdyn_spline_build(const float* sol, const float* aux, float* kd, float* two_c, float* three_d, int n_knots,
                      int64_t n, float t1_minus_t0, void* stream);
/* x(t) = a + s*(b + s*(two_c/2 + three_d*s/3)) for the piece the host selected (s = t - t_piece). */
TDYN_API int tdyn_spline_eval(float* out, const float* a, const float* b, const float* two_c, const float* three_d, float s,
                     int64_t n, void* stream);

/* ---- fused MLP vector fields ------------------------------------------------------------------
 * f(x) = W_L act(... act(W_1 x + b_1) ...) + b_L, weights in nn.Linear layout ([out,in] row-major). */
#define TDYN_ACT_TANH 0
#define TDYN_ACT_SOFTPLUS 1 /* torch.nn.Softplus(beta=1, threshold=20) */
#define TDYN_ACT_RELU 2
#define TDYN_MAX_LAYERS 4

typedef struct {
    int n_layers;                      /* number of Linear layers, 1..4 */
    int dims[TDYN_MAX_LAYERS + 1];     /* dims[0] = state dim = dims[n_layers] */
    int act;                           /* TDYN_ACT_* applied between Linear layers */
    const float* w[TDYN_MAX_LAYERS];   /* device, [dims[l+1], dims[l]] */
    const float* b[TDYN_MAX_LAYERS];   /* device, [dims[l+1]] */
} tdyn_mlp_t;

/* K1/K3 (stage form). FFMA kernels for NARROW MLP vector fields (every width <= 128, <= 4 Linear layers; weights in
 * shared memory).  `sign` (+1 / -1) multiplies every output: reversed-time solves integrate -f(-t, .)
 * (numerics/odeint.py:54-58) without an extra negation pass.
 *   forward : out = sign * f(y)                                             replaces f(t, y) (core/defunc.py:30-35)
 *   adjoint : dx = sign*f(x); dlam = sign*(-lam^T df/dx); dmu = sign*(-sum_batch lam^T df/dtheta), i.e. the body of
 *             adjoint_dynamics (numerics/sensitivity.py:57-75 and :130-147): forward + VJP + the batch-reduced
 *             parameter-gradient contraction in ONE launch; dmu is in vf.parameters() order (weight [out,in]
 *             row-major, then bias, layer by layer).  `workspace`: tdyn_mlp_ffma_workspace_bytes() zero-initialised
 *             bytes (per-CTA partials + ticket; deterministic fixed-order cross-CTA sum). */
TDYN_API int tdyn_mlp_ffma_supported(const tdyn_mlp_t* mlp);
TDYN_API int64_t tdyn_mlp_ffma_workspace_bytes(const tdyn_mlp_t* mlp);
TDYN_API int tdyn_mlp_ffma_forward(const tdyn_mlp_t* mlp, const float* y, float* out, int64_t rows, float sign, void* stream);
TDYN_API int tdyn_mlp_ffma_adjoint(const tdyn_mlp_t* mlp, const float* x, const float* lam, float* dx, float* dlam,
                                   float* dmu, void* workspace, int64_t rows, float sign, void* stream);

/* Fused CNF vector field with Hutchinson's trace estimator (models/cnf.py:26-31, :55-65) for a narrow MLP `mlp`:
 * state s = [logp, x] ([rows, D+1]), eps [rows, D] the noise sampled once per forward (core/neuralde.py:112-116).
 *   forward : out = sign * [ -eps^T (df/dx) eps , f(x) ]                       (MLP forward + one in-kernel VJP)
 *   adjoint : ds as above, dlam = sign * (-lam^T dF/ds), dmu = sign * (-sum_batch lam^T dF/dtheta) -- the body of
 *             adjoint_dynamics (numerics/sensitivity.py:57-75) including the SECOND-ORDER terms the reference gets from
 *             autograd's create_graph=True, written out analytically.  workspace as for tdyn_mlp_ffma_adjoint. */
TDYN_API int tdyn_cnf_ffma_supported(const tdyn_mlp_t* mlp);
TDYN_API int tdyn_cnf_ffma_forward(const tdyn_mlp_t* mlp, const float* s, const float* eps, float* out, int64_t rows, float sign,
                                   void* stream);
TDYN_API int tdyn_cnf_ffma_adjoint(const tdyn_mlp_t* mlp, const float* s, const float* lam, const float* eps, float* ds,
                                   float* dlam, float* dmu, void* workspace, int64_t rows, float sign, void* stream);

/* K1/K3 (persistent form). A whole `odeint` call (numerics/odeint.py:29-91 + :326-408: k1, init_step, the adaptive loop
 * with checkpoint handling, FSAL, the step controller) for a narrow MLP vector field in ONE cooperative kernel; one
 * grid barrier per attempted step carries the global error norm.  adjoint = 1 integrates one interval of the backsolve
 * adjoint (numerics/sensitivity.py:55-84): y0/y1 are the flat A = [x, lam, mu, lam_t]; mu is integrated per CTA on
 * its batch-partial and reduced once at the end.  `plan` restates the stepper (numerics/solvers/ode.py:145-178).
 * status (device, 6 ints): n_out, n_attempt, n_accept, nfe, error (0 ok, 1 max_steps/NaN, 2 out_cap), index (0/1) of
 * the y buffer holding the final state.  trace (device, optional): [0] = dt0, then (t, dt, error_ratio) per attempt.
 * `k`: 7 stage buffers, contiguous, each (adjoint ? 2 : 1) * rows * D floats.  `out`: [out_cap][state elems].
 * `t_span` is a HOST array of n_t <= 1024 times (already negated for reversed spans); everything else is device memory. */
typedef struct {
    int n_extra;                        /* stages after k1: 6 (dopri5/tsit5), 3 (rk4), 0 (euler) */
    int mode[6];                        /* TDYN_MODE_* of each stage input */
    int nk[6];
    float a[6][TDYN_MAX_STAGES];
    float c[6];
    int n_sol; float bsol[TDYN_MAX_STAGES];
    int n_err; float berr[TDYN_MAX_STAGES];   /* n_err == 0: fixed-step method */
    int order;
    float safety, min_factor, max_factor;
} tdyn_rk_plan_t;
/* Batch sharding over the GPUs of one box (one process per GPU).  `peer[r]` is rank r's exchange buffer
 * (tdyn_comm_buffer_bytes() bytes, zero-initialised once) as mapped into THIS process (symmetric / peer memory over
 * NVLink); `epoch` must increase with every collective launch and be equal on all ranks.  Each global reduction of the
 * solve (init_step norms, the error norm of every attempted step) is then an in-kernel all-gather by direct peer
 * stores + fixed-order sum -- the reference's single global controller without leaving the kernel.  comm == NULL or
 * world == 1: single GPU.  For the adjoint the full-batch d(mu) needed by init_step is exchanged the same way (vectors of
 * up to 16384 parameters); mu itself stays a per-rank partial sum (all-reduced by the caller, like a gradient). */
typedef struct {
    int world, rank;
    unsigned int epoch;
    void* peer[TDYN_MAX_PEERS];
} tdyn_comm_t;
TDYN_API int64_t tdyn_comm_buffer_bytes(void);
TDYN_API int tdyn_mlp_solve_supported(const tdyn_mlp_t* mlp, int adjoint);
TDYN_API int64_t tdyn_mlp_solve_workspace_bytes(const tdyn_mlp_t* mlp);
TDYN_API int tdyn_mlp_solve(const tdyn_mlp_t* mlp, const tdyn_rk_plan_t* plan, int adjoint, float sign, float atol, float rtol,
                            const float* t_span, int n_t, int return_all, float* y0, float* y1, float* k, float* out,
                            float* t_out, int out_cap, void* workspace, int* status, float* trace, int trace_cap,
                            int64_t rows, int max_steps, const tdyn_comm_t* comm, void* stream);

/* K4. One interval of the interpolated adjoint (numerics/sensitivity.py:130-152) for a narrow MLP in one cooperative
 * kernel: state y = [lam (rows*D), mu (P)], x(t) from the natural cubic spline built by tdyn_spline_build (sol, kd,
 * two_c, three_d, knots = device array of n_knots times), no seminorm (mu is inside the error norm: its batch-partial
 * sums are reduced across the grid every attempted step).  t_start/t_end: the NEGATED span (-t_hi, -t_lo).
 * `k`: 7 stage buffers of rows*D floats.  status/trace/workspace as for tdyn_mlp_solve. */
TDYN_API int tdyn_mlp_solve_interp(const tdyn_mlp_t* mlp, const tdyn_rk_plan_t* plan, float atol, float rtol, float t_start,
                                   float t_end, float* y0, float* y1, float* k, const float* sol, const float* kd,
                                   const float* two_c, const float* three_d, const float* knots, int n_knots,
                                   void* workspace, int* status, float* trace, int trace_cap, int64_t rows, int max_steps,
                                   void* stream);

/* K2. Tensor-core (tcgen05, split precision: TF32 main term + bf16 correction chain) stage kernel for the
 * 32-256-256-32 MLP class of BASELINE configs[1] (tdyn_mlp_tc_supported tells; other wide shapes: tdyn_lin_tc_* below):
 *   y = combine(x, k[], coef, mode, dt)   (same semantics as tdyn_rk_combine; n_k == 0 -> y = x)
 *   k_out = f(y)                          [rows, dims[0]]
 * and, when x_new != NULL (last stage of a fixed-step method), additionally
 *   x_new = x + dt*(sum_j bsol_j k_j) with k_out as the last term          (ode.py:104)
 * Replaces `f(t + c*dt, x + dt*(...))` of ode.py:100-103 for nn.Sequential MLP vector fields.
 * `wsplit` is a device buffer of tdyn_mlp_tc_prepared_bytes() holding the pre-split weights written by
 * tdyn_mlp_tc_prepare (once per parameter update). */
TDYN_API int64_t tdyn_mlp_tc_prepared_bytes(const tdyn_mlp_t* mlp);
TDYN_API int tdyn_mlp_tc_supported(const tdyn_mlp_t* mlp);
TDYN_API int tdyn_mlp_tc_prepare(const tdyn_mlp_t* mlp, void* wsplit, void* stream);
TDYN_API int tdyn_mlp_tc_stage(const tdyn_mlp_t* mlp, const void* wsplit, float* k_out, float* y_out,
                      const float* x, const float* const* k, const float* coef, int n_k, int mode, float dt,
                      float* x_new, const float* bsol, int n_sol, int64_t rows, void* stream);

/* K2, forward with the hidden activations kept: out = scale * f(x), a1 / a2 = the two hidden activations [rows, 256] --
 * the forward half of the adjoint dynamics (numerics/sensitivity.py:62-66: `dx = vf(t, x)` under autograd keeps the
 * activations for the VJP); `scale` folds the negation of the reversed-time field. */
TDYN_API int tdyn_mlp_tc_forward_save(const tdyn_mlp_t* mlp, const void* wsplit, float* out, const float* x, float scale,
                                      float* a1, float* a2, int64_t rows, void* stream);

/* K2, input-gradient chain of the same MLP class in one launch (the input half of the VJP `grad(dx, x, -lambda)`,
 * numerics/sensitivity.py:62-66): with the TRANSPOSED network `mlp_t` (w[0] = W3^T [256,32], w[1] = W2^T [256,256],
 * w[2] = W1^T [32,256], zero biases; `wsplit_t` = its prepared image)
 *   d2 = (g W3) * act'(a2),  d1 = (d2 W2) * act'(a1),  out = scale * (d1 W1)
 * where a1 / a2 are the hidden activations saved by tdyn_mlp_tc_forward_save; d2 / d1 ([rows, 256]) are written out for
 * the parameter gradients (tdyn_wgrad_tc). */
TDYN_API int tdyn_mlp_tc_dgrad(const tdyn_mlp_t* mlp_t, const void* wsplit_t, float* out, const float* g, float scale,
                               const float* a2, const float* a1, float* d2, float* d1, int64_t rows, void* stream);

/* K2 (persistent form). A whole FIXED-STEP solve of the same MLP class -- the loop of numerics/odeint.py:411-445 around
 * RungeKutta4.step / Euler.step / Midpoint.step (numerics/solvers/ode.py:61-104) -- in ONE launch: every CTA walks
 * n_steps x n_stages passes over its own row tiles (rows are independent and each thread re-reads only what it wrote, so
 * there is no grid-wide synchronisation); the solution update x + dt*(sum_j bsol_j k_j) of a step is folded into the
 * first stage input of the next one, terms whose tableau coefficient is exactly 0 are not loaded (they add +-0).
 * `plan`: the stepper (n_err == 0).  `dts[n_steps]`, `save_slot[n_steps + 1]`: HOST arrays -- the step sizes (fp32, from the
 * accumulated t as in odeint.py:431-434) and, per time point t_s, the row of `out` the state is saved to or -1 (slot of
 * t_0 is ignored: the caller owns x; the final state must be saved).  `xw` [rows*32] and `k` [n_stages*rows*32] are
 * scratch, `workspace` = tdyn_mlp_tc_solve_workspace_bytes(n_steps, n_stages) bytes of device memory. */
TDYN_API int tdyn_mlp_tc_solve_supported(const tdyn_mlp_t* mlp, int64_t rows);
TDYN_API int64_t tdyn_mlp_tc_solve_workspace_bytes(int n_steps, int n_stages);
TDYN_API int tdyn_mlp_tc_solve_fixed(const tdyn_mlp_t* mlp, const void* wsplit, const tdyn_rk_plan_t* plan, const float* x_in,
                                     float* xw, float* k, float* out, const float* dts, const int* save_slot, int n_steps,
                                     int64_t rows, void* workspace, void* stream);

/* K2 (persistent ADAPTIVE form).  A whole adaptive `odeint` call of the same MLP class -- numerics/odeint.py:29-91 + :326-408
 * around DormandPrince45.step / Tsitouras45.step (numerics/solvers/ode.py:145-178) -- in ONE cooperative launch of the
 * tensor-core stage kernel: k1 = f(x0) and init_step (numerics/utils.py:37-54; skipped when dt0 > 0: then k[0] must hold
 * k1 = sign*f(x0) and dt0 is the first step size), then per attempted step the stages 2..7 on the tensor cores (stage
 * combination in the prologue), x_new = x + dt*(sum bsol_j k_j), err = dt*(sum berr_j k_j), the scaled sum of squares
 * (odeint.py:365-372) reduced across the grid by ONE grid barrier, and the fp32 step controller with the checkpoint
 * handling of odeint.py:352-401 (accept / reject, FSAL by swapping the roles of two stage buffers, adapt_step of
 * numerics/utils.py:57-63), computed identically by one thread of every CTA.  Rows are independent and every thread
 * re-reads only what it wrote itself: the error norm is the only grid-wide exchange.  Any number of rows >= 1: one CTA per
 * 128-row tile up to the SM count; a CTA that owns a single tile chains its stages serially.
 * `t_span`: HOST array of 2..1024 increasing times (already negated for a reversed span; `sign` = -1 then: every stage
 * derivative is sign*f, except init_step's trial evaluation, as in the reference).  x0 [rows*32] holds the initial state
 * on entry; x0 / x1 ping-pong, status[5] tells which one holds the final state.  `k`: 7 stage buffers of rows*32 floats.
 * out [out_cap][rows*32], t_out [out_cap]: slot 0 = the initial state.  status (device, 6 ints), trace (device,
 * optional): as for tdyn_mlp_solve.  workspace: tdyn_mlp_tc_adaptive_workspace_bytes() bytes of device memory. */
TDYN_API int64_t tdyn_mlp_tc_adaptive_workspace_bytes(void);
TDYN_API int tdyn_mlp_tc_solve_adaptive(const tdyn_mlp_t* mlp, const void* wsplit, const tdyn_rk_plan_t* plan, float sign,
                                        float atol, float rtol, const float* t_span, int n_t, int return_all, float dt0,
                                        float* x0, float* x1, float* k, float* out, float* t_out, int out_cap, void* workspace,
                                        int* status, float* trace, int trace_cap, int64_t rows, int max_steps, void* stream);

/* ---- Layer-wise tensor-core kernels for wide MLP vector fields (any widths that are multiples of 32; output widths
 * <= 128 or multiples of 128, <= 2048): BASELINE configs[2] (128-1024-128) and the adjoint of the 32-256-256-32 class.
 * Replaces, per layer, the nn.Linear / activation calls of the vector field (reference core/defunc.py:30-35) and the
 * autograd VJP `grad(dx, (x, t) + params, -lambda)` of the adjoint dynamics (numerics/sensitivity.py:62-66, 134-138).
 *
 * tdyn_lin_tc_prepare: image of the B operand of one layer (TF32 hi part + bf16 correction part, swizzled), `transpose` = 0 for the forward
 *   (B = W [w_rows = out, w_cols = in]) and 1 for the input gradient (B = W^T).  Rebuild after a parameter update.
 * tdyn_lin_tc_apply:  out[r, n] = scale * epi(sum_k in[r, k] B[n, k] + bias[n]);  mode 0: epi = identity,
 *   1: epi = act, 2: epi(v) = v * act'(z) with act' evaluated from the saved activation OUTPUT saved[r, n].
 *   in [rows, k_in], out / saved [rows, n_out], contiguous fp32, 16-byte aligned; bias may be NULL.
 * tdyn_wgrad_tc: dw[n, k] = scale * sum_r g[r, n] in[r, k], db[n] = scale * sum_r g[r, n]  (g [rows, n_out],
 *   in [rows, k_in]); deterministic (fixed-order reduction of per-CTA partials held in `ws`). */
TDYN_API int tdyn_lin_tc_supported(int n_out, int k_in);
TDYN_API int64_t tdyn_lin_tc_image_bytes(int n_out, int k_in);
TDYN_API int tdyn_lin_tc_prepare(const float* w, int w_rows, int w_cols, int transpose, void* img, void* stream);
TDYN_API int tdyn_lin_tc_apply(const void* img, int n_out, int k_in, const float* in, const float* bias, const float* saved,
                      float* out, int act, int mode, float scale, int64_t rows, void* stream);
TDYN_API int64_t tdyn_wgrad_tc_workspace_bytes(int n_out, int k_in, int64_t rows);
TDYN_API int tdyn_wgrad_tc(const float* g, int n_out, const float* in, int k_in, float* dw, float* db, float scale,
                  void* ws, int64_t ws_bytes, int64_t rows, void* stream);

/* ---- Device-side adaptive loop for an OPAQUE vector field.  The accept/reject loop of the reference's
 * _adaptive_odeint (numerics/odeint.py:351-407: clamp to T / next output time, attempt, error ratio, accept, output
 * bookkeeping, adapt_step of numerics/utils.py:57-63) as a CUDA-graph conditional WHILE node: the body -- one attempted
 * step -- is captured by the caller (tdyn_loop_begin, [tdyn_rk_combine_dev, f] x stages, tdyn_rk_finish_dev,
 * tdyn_loop_end, tdyn_loop_commit on a capturing stream; f's own kernels are captured by PyTorch) and handed to
 * tdyn_graph_while_create, which clones it into `while (*keep_going) body;`, instantiates it and returns the executable
 * graph: one tdyn_graph_launch per odeint call, no host synchronisation per step.
 * Controller state: `cf` (tdyn_loop_ctrl_sizes floats) and `ci` (ints), device arrays the caller initialises:
 *   cf: [0] t, [1] dt, [2] T, [3] dt_keep, [4] dt of the current attempt (what *_dev kernels read), [5] error ratio,
 *       [6] safety, [7] min_factor, [8] max_factor, [10..15] c_i, [16..21] stage times t + c_i*dt (1-element views for f)
 *   ci: [0] ck (next output time), [1] ck_flag, [2] n_eval, [3] n_attempt, [4] n_accept, [5] n_out, [6] out_cap,
 *       [7] return_all_eval, [8] keep_going, [9] accept, [10] output slot of this step or -1, [11] status (0 ok, 1 max
 *       steps / NaN dt, 2 output buffer full), [12] max_steps, [13] order, [14] number of stage times, [15] trace_cap
 * `t_eval`: device array of the n_eval output times after t0; `t_out` / `out`: [out_cap] times and states;
 * `trace` (optional): [1 + 3*attempts] = dt0, then (t, dt, ratio) per attempt. */
TDYN_API int tdyn_loop_ctrl_sizes(int* n_float, int* n_int);
TDYN_API int tdyn_rk_combine_dev(float* y, const float* x, const float* const* k, const float* coef, int n_k, int mode,
                                 const float* dt_dev, int64_t n, void* stream);
TDYN_API int tdyn_rk_finish_dev(float* x_new, float* err_out, const float* x, const float* const* k, const float* bsol,
                                int n_sol, const float* berr, int n_err, const float* dt_dev, float atol, float rtol, int64_t n,
                                int64_t norm_n, double* sumsq_out, void* workspace, void* stream);
TDYN_API int tdyn_loop_begin(float* cf, int* ci, const float* t_eval, void* stream);
TDYN_API int tdyn_loop_end(float* cf, int* ci, const double* sumsq, int64_t norm_count, const float* t_eval, float* t_out,
                           float* trace, void* stream);
TDYN_API int tdyn_loop_commit(const int* ci, float* x, const float* x_new, float* k1, const float* k_last, float* out,
                              int64_t n, void* stream);
TDYN_API int tdyn_graph_while_create(void* body_graph, const int* keep_going, void** exec_out);
TDYN_API int tdyn_graph_launch(void* exec, void* stream);
TDYN_API int tdyn_graph_destroy(void* exec);

/* ---- In-kernel all-reduce(SUM) over the GPUs of one box through peer-mapped (symmetric) memory (NVLink / NVSwitch):
 * the exchange step of the sharded solves -- the d(mu)/dt block of every stage of the sharded interpolated adjoint
 * (numerics/sensitivity.py:124,152 puts mu inside the error norm) and the partial sums of squares of the host-driven
 * adaptive loops (numerics/utils.py:33-34 over a sharded batch).  `comm->peer[r]` = rank r's exchange buffer of
 * tdyn_peer_allreduce_buffer_bytes(slot_bytes) bytes (zero-initialised once), `comm->epoch` = a call counter that is equal
 * on all ranks and increases by one per call.  One launch per GPU: all-gather by direct peer stores, per-rank release
 * flags, fixed-order sum (deterministic).  `vec` (n floats or doubles) is reduced in place. */
TDYN_API int64_t tdyn_peer_allreduce_buffer_bytes(int64_t slot_bytes);
TDYN_API int tdyn_peer_allreduce(const tdyn_comm_t* comm, void* vec, int64_t n, int is_double, int64_t slot_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TDYN_B200_H */
