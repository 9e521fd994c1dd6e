// Device-side adaptive loop for an OPAQUE vector field: the accept/reject loop of the reference's
// _adaptive_odeint (numerics/odeint.py:351-407) runs as a CUDA-graph conditional WHILE node whose body is one attempted
// step -- [controller begin, (stage combine, f) x 6, finish, controller end, commit] -- so a whole odeint call is ONE
// graph launch with zero host synchronisations per step.  f stays a torch call: its kernels are captured by PyTorch into
// the body graph (torch.cuda.graph), this file supplies the controller kernels, the device-dt variants of the
// stage-combine / finish kernels' arguments, and the graph plumbing (conditional handle, WHILE node, condition kernel).
//
// Controller state lives in two small device arrays (`cf` floats, `ci` ints; indices below) owned by the caller.
// Arithmetic: IEEE fp32 with the reference's per-op rounding, identical to the host controller
// (torchdyn_b200/numerics/odeint.py, numerics/utils.py) and to the in-kernel controller of tdyn_mlp_solve.cu.
#include "tdyn_common.cuh"

namespace tdyn {
namespace gl {

// ---- cf (float) slots
enum { F_T = 0, F_DT, F_TEND, F_DT_KEEP, F_DT_CUR, F_RATIO, F_SAFETY, F_MINF, F_MAXF, F_C0 = 10 /* c[0..5] */,
       F_TS0 = 16 /* stage times t + c_i*dt, handed to f as 1-element tensors */, F_COUNT = 24 };
// ---- ci (int) slots
enum { I_CK = 0, I_CK_FLAG, I_N_EVAL, I_N_ATTEMPT, I_N_ACCEPT, I_N_OUT, I_OUT_CAP, I_RETURN_ALL, I_KEEP_GOING, I_ACCEPT,
       I_COPY_SLOT, I_STATUS /* 0 ok, 1 max_steps / NaN, 2 out_cap */, I_MAX_STEPS, I_ORDER, I_N_STAGE_T, I_TRACE_CAP, I_COUNT = 16 };

// start of an attempt (odeint.py:352-361): clamp to T, clamp to the next output time, publish dt and the stage times
__global__ void loop_begin_kernel(float* cf, int* ci, const float* t_eval) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float t = cf[F_T], T = cf[F_TEND];
    float dt = cf[F_DT];
    ci[I_CK_FLAG] = 0;
    if (add_rn(t, dt) > T) dt = sub_rn(T, t);
    const int ck = ci[I_CK];
    if (ck < ci[I_N_EVAL] && add_rn(t, dt) > t_eval[ck]) {
        cf[F_DT_KEEP] = dt;
        ci[I_CK_FLAG] = 1;
        dt = sub_rn(t_eval[ck], t);
    }
    cf[F_DT_CUR] = dt;
    const int ns = ci[I_N_STAGE_T];
    for (int i = 0; i < ns; ++i) cf[F_TS0 + i] = add_rn(t, mul_rn(cf[F_C0 + i], dt));       // t + c_i*dt (ode.py:148-153)
}

// end of an attempt (odeint.py:371-407): error ratio, accept / reject, output bookkeeping, step-size controller
__global__ void loop_end_kernel(float* cf, int* ci, const double* sumsq, double count, const float* t_eval, float* t_out,
                                float* trace) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float t = cf[F_T], T = cf[F_TEND];
    float dt = cf[F_DT_CUR];
    const float ratio = sqrtf((float)(sumsq[0] / count));
    cf[F_RATIO] = ratio;
    const int n_attempt = ++ci[I_N_ATTEMPT];
    if (trace && 3 * n_attempt < ci[I_TRACE_CAP]) { trace[3 * n_attempt - 2] = t; trace[3 * n_attempt - 1] = dt; trace[3 * n_attempt] = ratio; }
    const bool accept = ratio <= 1.f;
    ci[I_ACCEPT] = accept ? 1 : 0;
    ci[I_COPY_SLOT] = -1;
    float t_now = t;
    if (accept) {
        ++ci[I_N_ACCEPT];
        const float tn = add_rn(t, dt);
        const int ck = ci[I_CK];
        const bool hit = ck < ci[I_N_EVAL] && tn == t_eval[ck];          // exact float equality, like the reference
        if (hit || ci[I_RETURN_ALL]) {
            const int slot = ci[I_N_OUT];
            if (slot < ci[I_OUT_CAP]) { ci[I_COPY_SLOT] = slot; t_out[slot] = tn; ci[I_N_OUT] = slot + 1; }
            else ci[I_STATUS] = 2;
            if (hit) ci[I_CK] = ck + 1;
        }
        cf[F_T] = tn;
        t_now = tn;
    }
    if (ci[I_CK_FLAG]) dt = sub_rn(cf[F_DT_KEEP], dt);                  // the remainder of the intended step (:399-401)
    // adapt_step (utils.py:57-63)
    if (ratio == 0.f) dt = mul_rn(dt, cf[F_MAXF]);
    else {
        const float lo = ratio < 1.f ? 1.f : cf[F_MINF];
        const float pw = (float)pow((double)ratio, (double)div_rn(1.f, (float)ci[I_ORDER]));
        const float cand = div_rn(cf[F_SAFETY], pw);
        const float fac = (cand != cand) ? cand : fminf(cf[F_MAXF], fmaxf(cand, lo));
        dt = mul_rn(dt, fac);
    }
    cf[F_DT] = dt;
    int go = t_now < T;
    if (go && (n_attempt >= ci[I_MAX_STEPS] || dt != dt)) { ci[I_STATUS] = 1; go = 0; }
    if (ci[I_STATUS] == 2) go = 0;
    ci[I_KEEP_GOING] = go;
}

// accepted step: x <- x_new, k1 <- k_last (FSAL), and the new state into its output slot
__global__ void __launch_bounds__(256) loop_commit_kernel(const int* ci, float* __restrict__ x, const float* __restrict__ x_new,
                                                          float* __restrict__ k1, const float* __restrict__ k_last,
                                                          float* __restrict__ out, int64_t n, int vec) {
    if (!ci[I_ACCEPT]) return;
    const int slot = ci[I_COPY_SLOT];
    float* dst = slot >= 0 ? out + (size_t)slot * (size_t)n : nullptr;
    const int64_t tid = (int64_t)blockIdx.x * 256 + threadIdx.x, stride = (int64_t)gridDim.x * 256;
    int64_t done = 0;
    if (vec) {
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += stride) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(x_new) + i);
            reinterpret_cast<float4*>(x)[i] = v;
            if (dst) reinterpret_cast<float4*>(dst)[i] = v;
            reinterpret_cast<float4*>(k1)[i] = __ldg(reinterpret_cast<const float4*>(k_last) + i);
        }
        done = n4 << 2;
    }
    for (int64_t i = done + tid; i < n; i += stride) {
        const float v = x_new[i];
        x[i] = v;
        if (dst) dst[i] = v;
        k1[i] = k_last[i];
    }
}

__global__ void set_condition_kernel(cudaGraphConditionalHandle handle, const int* keep_going) {
    if (threadIdx.x == 0 && blockIdx.x == 0) cudaGraphSetConditional(handle, (unsigned int)(*keep_going != 0));
}

}  // namespace gl
}  // namespace tdyn

using namespace tdyn;

extern "C" {

int tdyn_loop_ctrl_sizes(int* n_float, int* n_int) {
    if (n_float) *n_float = gl::F_COUNT;
    if (n_int) *n_int = gl::I_COUNT;
    return 0;
}

int tdyn_loop_begin(float* cf, int* ci, const float* t_eval, void* stream) {
    TDYN_REQUIRE(cf && ci && t_eval, "tdyn_loop_begin: null argument");
    gl::loop_begin_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(cf, ci, t_eval);
    TDYN_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int tdyn_loop_end(float* cf, int* ci, const double* sumsq, int64_t norm_count, const float* t_eval, float* t_out, float* trace,
                  void* stream) {
    TDYN_REQUIRE(cf && ci && sumsq && t_eval && t_out && norm_count > 0, "tdyn_loop_end: bad argument");
    gl::loop_end_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(cf, ci, sumsq, (double)norm_count, t_eval, t_out, trace);
    TDYN_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int tdyn_loop_commit(const int* ci, float* x, const float* x_new, float* k1, const float* k_last, float* out, int64_t n,
                     void* stream) {
    TDYN_REQUIRE(ci && x && x_new && k1 && k_last && out && n > 0, "tdyn_loop_commit: bad argument");
    const int vec = ((((uintptr_t)x | (uintptr_t)x_new | (uintptr_t)k1 | (uintptr_t)k_last | (uintptr_t)out) & 15) == 0) && (n % 4 == 0);
    int64_t blocks = ((vec ? n / 4 : n) + 255) / 256;
    if (blocks > kMaxReduceBlocks) blocks = kMaxReduceBlocks;
    if (blocks < 1) blocks = 1;
    gl::loop_commit_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(ci, x, x_new, k1, k_last, out, n, vec);
    TDYN_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Wrap a captured body graph (one attempted step) into `while (*keep_going) body;` and instantiate it.  The body is cloned
// (child-graph node inside the conditional node's body), followed by the kernel that feeds *keep_going to the condition.
int tdyn_graph_while_create(void* body_graph, const int* keep_going, void** exec_out) {
    TDYN_REQUIRE(body_graph && keep_going && exec_out, "tdyn_graph_while_create: null argument");
    cudaGraph_t g = nullptr;
    TDYN_CHECK_CUDA(cudaGraphCreate(&g, 0));
    cudaGraphConditionalHandle h;
    cudaError_t e = cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
    if (e != cudaSuccess) { cudaGraphDestroy(g); set_error("cudaGraphConditionalHandleCreate: %s", cudaGetErrorString(e)); return 1; }
    cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = h;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t cond = nullptr;
    e = cudaGraphAddNode(&cond, g, nullptr, 0, &np);
    if (e != cudaSuccess) { cudaGraphDestroy(g); set_error("cudaGraphAddNode(conditional WHILE): %s", cudaGetErrorString(e)); return 1; }
    cudaGraph_t body = np.conditional.phGraph_out[0];
    cudaGraphNode_t child = nullptr;
    e = cudaGraphAddChildGraphNode(&child, body, nullptr, 0, (cudaGraph_t)body_graph);
    if (e != cudaSuccess) { cudaGraphDestroy(g); set_error("cudaGraphAddChildGraphNode(body): %s", cudaGetErrorString(e)); return 1; }
    cudaKernelNodeParams kp;
    memset(&kp, 0, sizeof(kp));
    void* args[2] = {(void*)&h, (void*)&keep_going};
    kp.func = (void*)gl::set_condition_kernel;
    kp.gridDim = dim3(1, 1, 1);
    kp.blockDim = dim3(32, 1, 1);
    kp.kernelParams = args;
    cudaGraphNode_t setc = nullptr;
    e = cudaGraphAddKernelNode(&setc, body, &child, 1, &kp);
    if (e != cudaSuccess) { cudaGraphDestroy(g); set_error("cudaGraphAddKernelNode(set condition): %s", cudaGetErrorString(e)); return 1; }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { set_error("cudaGraphInstantiate(while graph): %s", cudaGetErrorString(e)); return 1; }
    *exec_out = (void*)exec;
    return 0;
}

int tdyn_graph_launch(void* exec, void* stream) {
    TDYN_REQUIRE(exec, "tdyn_graph_launch: null graph");
    TDYN_CHECK_CUDA(cudaGraphLaunch((cudaGraphExec_t)exec, (cudaStream_t)stream));
    return 0;
}

int tdyn_graph_destroy(void* exec) {
    if (exec) TDYN_CHECK_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)exec));
    return 0;
}

}  // extern "C"
