// K2: tensor-core (tcgen05 / TMEM, split precision: TF32 + bf16 corrections) stage kernel for wide MLP vector fields on sm_100a.
//
//   y     = combine(x, k[], coef, mode, dt)                  (reference numerics/solvers/ode.py:100-103,148-153)
//   k_out = W3 act(W2 act(W1 y + b1) + b2) + b3              (the nn.Sequential vector field, f(t, y))
//   x_new = x + dt*(sum_j bsol_j k_j), k_out last            (ode.py:104, optional, last stage of a fixed-step method)
//
// Shape class (BASELINE.json configs[1]): state dim D = 32, two hidden layers of width 256, three Linear layers.
// One persistent CTA per SM walks 128-row tiles of the batch.  Per tile:
//   workers (16 warps) : load x/k rows, combine in the reference's rounding order, split y = hi + lo (hi = TF32) and
//                        store the TF32 tile and the bf16 correction tile [bf16(hi) | bf16(lo)] K-major / 128B-swizzled in
//                        shared memory (the UMMA canonical layout);
//   MMA warp (1 thread): layer 1 D1[128x256] (TMEM cols 0..255 = X) = A_hi*W_hi as tcgen05.mma kind::tf32 plus
//                        A_hi*W_lo + A_lo*W_hi as ONE kind::f16 (bf16) chain over the concatenated K (see store_corr8):
//                        fp32-accurate products, fp32 accumulate, 2/3 of the tensor time of plain 3xTF32;
//   workers            : X -> registers (tcgen05.ld, next chunk prefetched), +bias (shared memory), activation, split:
//                        hi goes back into the SAME TMEM columns (tcgen05.st) and is consumed by A-from-TMEM MMAs, the
//                        correction tile goes through a 3-stage shared-memory ring; one 32-wide K-block at a time, so layer 2's MMAs
//                        (TMEM cols 256..511 = Y) start while later K-blocks are still being produced; same again for
//                        layer 3 (N = 32, A_hi = Y in place, accumulator X[0,32));
//   producer (1 thread): streams the pre-split, pre-swizzled weight images (tdyn_mlp_tc_prepare) from L2 into a
//                        2 x 64 KB ring with cp.async.bulk + mbarrier complete_tx (weights total 640 KB per tile pass:
//                        they cannot stay resident in 227 KB of shared memory).
// The next tile's stage input is produced during layer 2; its layer 1 is issued once this tile's output left TMEM.
// Algorithmic work: 163 840 FLOP per sample-NFE (issued as 1 TF32 + 2 bf16 FLOPs = 4 bf16-equivalents).  HBM traffic is negligible (128-256 B per
// sample-NFE), the bound is the tensor pipe with the activation epilogue (512 tanh per sample-NFE) next.
#include <cuda_bf16.h>
#include <vector>
#include "tdyn_tc_ptx.cuh"

namespace tdyn {
namespace tc {

using namespace tcx;
constexpr int kD = 32;          // state dim
constexpr int kH = 256;         // hidden width
constexpr int kWorkers = 512;   // 16 epilogue/prologue warps (4 per TMEM lane quarter)
constexpr int kThreads = 64 + kWorkers;
constexpr int kKB = 32;         // K-block: 32 fp32 = 128 B = one swizzle row
constexpr uint32_t kATile = kTileM * 128;           // 16 KB: one [128 x 32] operand tile
constexpr uint32_t kAStage = 2 * kATile;            // hi + lo
constexpr uint32_t kWStage = 65536;                 // one weight stage image
constexpr int kWStagesPerTile = 10;                 // W1 | W2 k-blocks 0..7 | W3 (all k-blocks)
constexpr int kLoStages = 3;                        // ring of [128 x 32] TF32 "lo" operand tiles (the "hi" parts live in TMEM)
constexpr uint32_t kSmemA1 = 0;                                     // layer-1 operand: hi tile | lo tile
constexpr uint32_t kSmemLoRing = kSmemA1 + kAStage;                 // kLoStages x 16 KB
constexpr uint32_t kSmemWRing = kSmemLoRing + kLoStages * kATile;   // 2 stages x 64 KB
constexpr uint32_t kSmemW1B = kSmemWRing + 2 * kWStage;             // W1 rows 224..255 (hi 4 KB | lo 4 KB), see layer1B
constexpr uint32_t kW1BBytes = 8192;
constexpr uint32_t kSmemBias = kSmemW1B + kW1BBytes;                // b1[256] | b2[256] | b3[32] (fp32)
constexpr uint32_t kSmemBars = kSmemBias + (2 * kH + kD) * 4;
constexpr int kNSplit = 224;                        // layer 1 is issued as N = 224 (early) + N = 32 (once D3 left TMEM)
constexpr uint32_t kSmemTotal = kSmemBars + 256 + 1024;   // + slack for the manual 1024-byte alignment
constexpr uint32_t kTmemCols = 512;

enum Bar { W_FULL0 = 0, W_FULL1, W_EMPTY0, W_EMPTY1, A_FULL0, A_FULL1, A_FULL2, A_EMPTY0, A_EMPTY1, A_EMPTY2,
           A1_FULL, D1_FULL, D1B_FULL, D2_FULL, D3_FULL, D_FREE, W1B_FULL, W1B_EMPTY, NUM_BARS };

// One PASS = one sweep over all row tiles: build a stage input y = combine(x, k[], coef, mode, dt), optionally store it
// (the new state / an output time), evaluate the MLP on it and write the stage derivative.  tdyn_mlp_tc_stage launches ONE
// pass; tdyn_mlp_tc_solve_fixed launches a whole fixed-step solve = n_steps x n_stage passes (rows are independent and
// every thread re-reads only what it wrote itself, so no grid-wide synchronisation is needed between passes).
// Everything that describes a pass lives in the kernel parameters (constant bank: no load latency in the prologue); only
// the per-step dt and output slot come from two small device arrays.
struct StagePlan {                      // one stage input: which stage derivatives, which coefficients
    int n_k; int mode;
    int kidx[TDYN_MAX_STAGES];          // indices into Params::kp (terms with a zero coefficient dropped in solve mode)
    float coef[TDYN_MAX_STAGES];
};
constexpr int kUpd = TDYN_MAX_STAGES;   // st[kUpd]: the solution update x + dt*(sum_j bsol_j k_j)

struct Params {
    const uint8_t* wimg;       // 10 x 64 KB stage images
    const float* b1; const float* b2; const float* b3;
    StagePlan st[TDYN_MAX_STAGES + 2];        // st[g]: input of stage g; st[kUpd]: update; st[kUpd + 1]: empty (n_k = 0)
    const float* kp[TDYN_MAX_STAGES];         // stage-derivative buffers (stage mode: the caller's k pointers)
    const float* x_in;         // the caller's state (stage mode: x of this launch; solve mode: read by the very first pass)
    float* xw;                 // solve mode: working copy of the state (written by the first stage of every step)
    float* k_out0;             // stage mode: output
    float* out;                // solve mode: [n_saved][rows*32]
    const float* dts;          // solve mode: dt of every step (device)
    const int* slots;          // solve mode: output slot of every time point t_s, s = 0..n_steps (device; -1 = not saved)
    unsigned long long n_state;
    float dt0;                 // stage mode: dt
    int n_stage, n_steps;      // stage mode: 1, 1
    int solve;                 // 0: stage mode, 1: solve mode
    float* x_new; Coefs bsol; int n_sol;      // stage mode only: x + dt*(sum_j bsol_j k_j) in the output epilogue
    float* act_out[2];         // stage mode only: the hidden activations a1, a2 ([rows][256]) are also written here (the
                               // adjoint's input / parameter gradients need them), or NULL
    float out_scale;           // k_out = out_scale * f(y)  (1 except for the reversed-time field of an adjoint)
    const float* dact_in[2];   // dgrad variant: saved forward activations whose act' multiplies epilogue 0 (a2) / 1 (a1)
    // ---- adaptive variant (tdyn_mlp_tc_solve_adaptive): a whole dopri5 / tsit5 solve in one cooperative launch
    float* xbuf[2];            // state ping-pong (xbuf[0] holds x0 on entry)
    Coefs berr; int n_err;     // embedded error weights (st[kUpd] holds bsol, literal: no term dropped)
    const float* t_span;       // device copy of the (already negated, if reversed) output times, n_t of them
    int n_t, return_all, out_cap, max_steps, order, trace_cap;
    float t_dt0;               // > 0: first step size given and kp[0] = k1 given; <= 0: k1 and init_step run in the kernel
    float atol, rtol, safety, min_factor, max_factor;
    float* t_out; int* status; float* trace;
    unsigned int* gbar;        // grid barrier counter (zero on entry)
    double* partials;          // [2 parities][2 sums][gridDim.x] per-CTA partial sums of squares
    int64_t rows; int num_tiles; int act;
};

// what a work item needs beyond the constant-bank plan
struct Item {
    int g;           // stage index (k_out = kp[g]); the plan is st[g], or st[kUpd] for g == 0 (st[kUpd + 1], empty, for the
                     // very first pass, which reads Params::x_in); g == 0 also stores the stage input into Params::xw
    int slot;        // ... and into this output slot (-1: none)
    float dt;
};

// act(acc + bias); for tanh `bias` is the pre-scaled value (see the bias staging in the kernel)
template <int ACT>
__device__ __forceinline__ float activate_biased(float acc, float bias) {
    if (ACT != TDYN_ACT_TANH) return activate<ACT>(acc + bias);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(acc, kTanhScale, bias)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    return fmaf(-2.f, r, 1.f);
}

// Split precision: a = hi + lo with hi = rna_tf32(a) (exactly representable in TF32) and lo = a - hi (exact in fp32,
// |lo| <= 2^-12 |a|).  D = A_hi*W_hi runs as kind::tf32; the two correction terms A_hi*W_lo + A_lo*W_hi are 2^-11 of
// the result and only need ~10 significant bits, so they run as ONE kind::f16 (bf16) chain over a concatenated K:
// A' = [bf16(a_hi) | bf16(a_lo)] (one 128-byte row per batch row = 64 bf16), W' = [bf16(w_lo) ; bf16(w_hi)].  A bf16 MMA
// covers K = 16 in the cycles a TF32 MMA needs for K = 8: a K-block of 32 costs 4 TF32 + 4 BF16 MMAs instead of the 12
// TF32 MMAs of plain 3xTF32 (2/3 of the tensor-pipe time; the kernel is power-limited on the tensor pipe).  The bf16
// rounding of the correction operands adds ~2^-21 relative per product, below the accumulator's truncation noise.
__device__ __forceinline__ uint32_t pack_bf16(float lo_elem, float hi_elem) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo_elem, hi_elem);     // .x (low 16 bits) = element at the lower address
    return *reinterpret_cast<const uint32_t*>(&p);
}
// the correction tile row: 16-byte chunk `sub` holds bf16(hi) of k = 8*sub..8*sub+7, chunk 4+sub holds bf16(lo)
__device__ __forceinline__ void store_corr8(uint8_t* tile, int row, int sub, const float (&hi)[8], const float (&lo)[8]) {
    *reinterpret_cast<uint4*>(tile + swz((uint32_t)row, (uint32_t)sub)) =
        make_uint4(pack_bf16(hi[0], hi[1]), pack_bf16(hi[2], hi[3]), pack_bf16(hi[4], hi[5]), pack_bf16(hi[6], hi[7]));
    *reinterpret_cast<uint4*>(tile + swz((uint32_t)row, (uint32_t)(4 + sub))) =
        make_uint4(pack_bf16(lo[0], lo[1]), pack_bf16(lo[2], lo[3]), pack_bf16(lo[4], lo[5]), pack_bf16(lo[6], lo[7]));
}
// layer-1 operand: TF32 hi tile + bf16 correction tile, both in shared memory (K-major, 128B swizzle)
__device__ __forceinline__ void store_split8(uint8_t* stage, int row, int sub, const float (&a)[8]) {
    float hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { hi[i] = tf32_rna_fast(a[i]); lo[i] = a[i] - hi[i]; }
#pragma unroll
    for (int i = 0; i < 2; ++i)
        *reinterpret_cast<float4*>(stage + swz((uint32_t)row, (uint32_t)(sub * 2 + i))) =
            make_float4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
    store_corr8(stage + kATile, row, sub, hi, lo);
}
// hidden-layer operand: TF32 hi back into TMEM IN PLACE of the accumulator columns it came from (A-from-TMEM MMAs read
// it without touching shared memory), the bf16 correction tile into the shared-memory ring
__device__ __forceinline__ void store_split8_tmem(uint32_t taddr, uint8_t* corr_tile, int row, int sub, const float (&a)[8]) {
    uint32_t hiu[8];
    float hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        hi[i] = tf32_rna_fast(a[i]);
        hiu[i] = __float_as_uint(hi[i]);
        lo[i] = a[i] - hi[i];
    }
    tmem_st8(taddr, hiu);
    store_corr8(corr_tile, row, sub, hi, lo);
    tmem_wait_st();
}

// one K-block (32 wide) with all operands in shared memory: D (+)= A'*W' (4 bf16 MMAs, K = 16) + A_hi*W_hi (4 TF32 MMAs)
__device__ __forceinline__ void issue_kblock(uint32_t d_tmem, uint32_t a_hi, uint32_t a_corr, uint32_t b_hi, uint32_t b_corr,
                                             int n, bool first) {
    const uint32_t id32 = umma_idesc(n), id16 = umma_idesc_bf16(n);
#pragma unroll
    for (int k = 0; k < 4; ++k) tc_mma_bf16(d_tmem, umma_desc(a_corr + k * 32), umma_desc(b_corr + k * 32), id16, (first && k == 0) ? 0u : 1u);
#pragma unroll
    for (int k = 0; k < 4; ++k) tc_mma_tf32(d_tmem, umma_desc(a_hi + k * 32), umma_desc(b_hi + k * 32), id32, 1u);
}

// hidden-layer K-block: A_hi in TMEM (32 columns at a_hi_tmem), the correction tile in shared memory
__device__ __forceinline__ void issue_kblock_ts(uint32_t d_tmem, uint32_t a_hi_tmem, uint32_t a_corr, uint32_t b_hi, uint32_t b_corr,
                                                int n, bool first) {
    const uint32_t id32 = umma_idesc(n), id16 = umma_idesc_bf16(n);
#pragma unroll
    for (int k = 0; k < 4; ++k) tc_mma_bf16(d_tmem, umma_desc(a_corr + k * 32), umma_desc(b_corr + k * 32), id16, (first && k == 0) ? 0u : 1u);
#pragma unroll
    for (int k = 0; k < 4; ++k) tc_mma_tf32_ts(d_tmem, a_hi_tmem + k * 8, umma_desc(b_hi + k * 32), id32, 1u);
}

struct WorkerCtx {
    int row, sub;
    uint32_t lane_addr;
};

// 16-byte global load that is coherent with this thread's own earlier stores (multi-pass mode re-reads what the same
// thread wrote in an earlier pass): L2-cached, never the non-coherent texture path
// (every re-read is separated from the store it depends on by mbarrier waits, which order it for the compiler as well)
__device__ __forceinline__ float4 ld_f4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// y = combine(x, k, coef) for 8 elements of this thread's row, in the reference's rounding order.  The k's are fetched in
// groups of four with every load of a group issued before the first use (one exposed L2 round trip per group).
// A flag that is the same in every thread, handed to the compiler as a warp-uniform value (a vote result): what is indexed
// with it stays a constant-bank operand.  (__activemask: also called where a ragged last tile has switched lanes off.)
__device__ __forceinline__ bool uniform_flag(bool p) { return __all_sync(__activemask(), p); }

// `swapped` (adaptive variant only): FSAL without a copy -- after an accepted step the first and the last stage buffer swap
// roles (k1 <- k_last).  Always handed over as a warp-UNIFORM value (a vote result), so that the pointer stays an indexed
// constant-bank operand.
__device__ __forceinline__ const float* kbuf(const Params& P, bool swapped, int j) {
    const int last = P.n_stage - 1;
    return P.kp[swapped ? (j == 0 ? last : (j == last ? 0 : j)) : j];
}

__device__ __forceinline__ void combine_row8(const Params& P, const StagePlan& sp, const float* x, float dt, size_t goff, bool valid,
                                             float (&y)[8], bool kmap = false) {
    float xv[8];
    const int n_k = valid ? sp.n_k : 0;
    float4 q[4][2];
    if (valid) {
        const float4 a = ld_f4(x + goff), b = ld_f4(x + goff + 4);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
            if (jj < n_k) { const float* kp = kbuf(P, kmap, sp.kidx[jj]) + goff; q[jj][0] = ld_f4(kp); q[jj][1] = ld_f4(kp + 4); }
        xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w; xv[4] = b.x; xv[5] = b.y; xv[6] = b.z; xv[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) xv[i] = 0.f;
    }
    if (n_k == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = xv[i];
        return;
    }
    const bool chain = sp.mode == TDYN_MODE_CHAIN;
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = xv[i];         // CHAIN: running y; INNER: overwritten by the first product
    for (int j0 = 0; j0 < n_k; j0 += 4) {
        if (j0 > 0) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
                if (j0 + jj < n_k) { const float* kp = kbuf(P, kmap, sp.kidx[j0 + jj]) + goff; q[jj][0] = ld_f4(kp); q[jj][1] = ld_f4(kp + 4); }
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            if (j0 + jj < n_k) {
                const float c = chain ? mul_rn(dt, sp.coef[j0 + jj]) : sp.coef[j0 + jj];
                const float kv[8] = {q[jj][0].x, q[jj][0].y, q[jj][0].z, q[jj][0].w, q[jj][1].x, q[jj][1].y, q[jj][1].z, q[jj][1].w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float m = mul_rn(c, kv[i]);
                    s[i] = (!chain && j0 + jj == 0) ? m : add_rn(s[i], m);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = chain ? s[i] : add_rn(xv[i], mul_rn(dt, s[i]));
}

__device__ __forceinline__ void store_row8(float* dst, size_t goff, const float (&y)[8]) {
    reinterpret_cast<float4*>(dst + goff)[0] = make_float4(y[0], y[1], y[2], y[3]);
    reinterpret_cast<float4*>(dst + goff)[1] = make_float4(y[4], y[5], y[6], y[7]);
}

// the work item of pass `pass` (solve mode: step = pass / n_stage, stage = pass % n_stage)
__device__ __forceinline__ Item make_item(const Params& P, int pass) {
    Item it;
    it.slot = -1;
    if (!P.solve) { it.g = 0; it.dt = P.dt0; return it; }
    const int s = pass / P.n_stage, g = pass - s * P.n_stage;
    it.g = g;
    if (g == 0) {
        it.dt = s > 0 ? __ldg(P.dts + s - 1) : 0.f;
        if (s > 0) it.slot = __ldg(P.slots + s);
    } else {
        it.dt = __ldg(P.dts + s);
    }
    return it;
}
__device__ __forceinline__ int plan_index(const Params& P, const Item& it, bool first_pass) {
    return !P.solve ? 0 : (it.g > 0 ? it.g : (first_pass ? kUpd + 1 : kUpd));
}
// Adaptive variant.  A batch is one of three PHASES: 0 = k1 = f(x0) (the empty plan st[kUpd + 1]); 1 = init_step's trial
// evaluation f(x0 + h0*k1) (plan st[0], written to the last stage buffer); 2 = one attempted step, stages 2..s (pass p
// evaluates stage p + 1 from plan st[p + 1]).  `slot` carries the plan index here.
__device__ __forceinline__ Item make_item_adapt(const Params& P, int phase, int pass, float dt) {
    Item it;
    it.dt = dt;
    if (phase == 2) { it.g = pass + 1; it.slot = pass + 1; }
    else if (phase == 0) { it.g = 0; it.slot = kUpd + 1; }
    else { it.g = P.n_stage - 1; it.slot = 0; }
    return it;
}

// stage input of one tile -> A1 (hi/lo, swizzled); solve mode also stores it as the new state / an output
__device__ __forceinline__ void prologue(const Params& P, const Item& it, bool first_pass, uint8_t* smem, const WorkerCtx& w,
                                         int64_t tile, const float* x_adaptive = nullptr, bool kmap = false) {
    const int64_t grow = tile * kTileM + w.row;
    const bool valid = grow < P.rows;
    const size_t goff = (size_t)grow * kD + w.sub * 8;
    float y[8];
    // (the very first pass of a solve copies the caller's state: no stage derivatives exist yet)
    const float* xsrc = x_adaptive ? x_adaptive : ((first_pass || !P.solve) ? P.x_in : P.xw);
    combine_row8(P, P.st[x_adaptive ? it.slot : plan_index(P, it, first_pass)], xsrc, it.dt, goff, valid, y, kmap);
    if (valid && P.solve && !x_adaptive && it.g == 0) {
        store_row8(P.xw, goff, y);
        if (it.slot >= 0) store_row8(P.out + (size_t)it.slot * P.n_state, goff, y);
    }
    store_split8(smem + kSmemA1, w.row, w.sub, y);
    fence_proxy_async();
}

// adaptive variant, end of an attempted step for 8 elements of one row: x_new = x + dt*(sum bsol_j k_j), err = dt*(sum
// berr_j k_j) in the reference's rounding (ode.py:154-155), returns the sum of the squared scaled errors
__device__ __forceinline__ double finish_row8(const Params& P, bool kmap, const float* x, float* x_new, float dt, size_t goff) {
    const StagePlan& sol = P.st[kUpd];               // bsol (no zero-dropping here: literal expression)
    float xv[8], ss[8], se[8];
    const float4 a = ld_f4(x + goff), b = ld_f4(x + goff + 4);
    xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w; xv[4] = b.x; xv[5] = b.y; xv[6] = b.z; xv[7] = b.w;
    for (int j0 = 0; j0 < P.n_err; j0 += 4) {
        float4 q[4][2];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
            if (j0 + jj < P.n_err) { const float* kp = kbuf(P, kmap, j0 + jj) + goff; q[jj][0] = ld_f4(kp); q[jj][1] = ld_f4(kp + 4); }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = j0 + jj;
            if (j < P.n_err) {
                const float kv[8] = {q[jj][0].x, q[jj][0].y, q[jj][0].z, q[jj][0].w, q[jj][1].x, q[jj][1].y, q[jj][1].z, q[jj][1].w};
                const float be = P.berr.c[j];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float m = mul_rn(be, kv[i]);
                    se[i] = j == 0 ? m : add_rn(se[i], m);
                }
                if (j < sol.n_k) {
                    const float bs = sol.coef[j];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float m = mul_rn(bs, kv[i]);
                        ss[i] = j == 0 ? m : add_rn(ss[i], m);
                    }
                }
            }
        }
    }
    double acc = 0.0;
    float xn[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        xn[i] = add_rn(xv[i], mul_rn(dt, ss[i]));
        const float er = mul_rn(dt, se[i]);
        const float den = add_rn(P.atol, mul_rn(P.rtol, fmaxf(fabsf(xv[i]), fabsf(xn[i]))));
        const float r = div_rn(er, den);
        acc += (double)mul_rn(r, r);
    }
    store_row8(x_new, goff, xn);
    return acc;
}

// adaptive variant, init_step (numerics/utils.py:37-54) for 8 elements of one row.  phase 0: sums of (x0/scale)^2 and
// (f0/scale)^2; phase 1: sum of ((f1 - f0)/scale)^2 with f1 = f(x0 + h0*f0) in the last stage buffer
__device__ __forceinline__ void init_row8(const Params& P, int phase, const float* x, size_t goff, double& a0, double& a1) {
    const float4 xa = ld_f4(x + goff), xb = ld_f4(x + goff + 4);
    const float4 fa = ld_f4(P.kp[0] + goff), fb = ld_f4(P.kp[0] + goff + 4);
    const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
    const float f0[8] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w};
    if (phase == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float scale = add_rn(P.atol, mul_rn(fabsf(xv[i]), P.rtol));
            const float q0 = div_rn(xv[i], scale), q1 = div_rn(f0[i], scale);
            a0 += (double)mul_rn(q0, q0); a1 += (double)mul_rn(q1, q1);
        }
    } else {
        const float* k1p = P.kp[P.n_stage - 1];
        const float4 ga = ld_f4(k1p + goff), gb = ld_f4(k1p + goff + 4);
        const float f1[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float scale = add_rn(P.atol, mul_rn(fabsf(xv[i]), P.rtol));
            const float q = div_rn(sub_rn(f1[i], f0[i]), scale);
            a0 += (double)mul_rn(q, q);
        }
    }
}

// adaptive variant: the step controller.  One worker thread per CTA runs it; every CTA computes the same values from the
// same grid-wide sums, so no decision is ever communicated between CTAs.
struct Ctrl {
    float t, dt, dt_keep, T, h0, d1;
    int ck, ck_flag, nout, attempt, accept, err, nfe;
    unsigned int nbar;
    // what the CTA's threads read at the start of the next batch / after the decision
    int phase; float dt_attempt; int xcur; int swapped; int save_slot;
};
// start of an attempt (odeint.py:352-361): clamp to T, clamp to the next output time
__device__ __forceinline__ void ctrl_begin_attempt(const Params& P, Ctrl& c) {
    c.ck_flag = 0;
    if (add_rn(c.t, c.dt) > c.T) c.dt = sub_rn(c.T, c.t);
    if (c.ck < P.n_t - 1 && add_rn(c.t, c.dt) > __ldg(P.t_span + 1 + c.ck)) {
        c.dt_keep = c.dt; c.ck_flag = 1;
        c.dt = sub_rn(__ldg(P.t_span + 1 + c.ck), c.t);
    }
    c.dt_attempt = c.dt;
}
// after the grid-wide sums of a batch: returns the number of work items of the next batch (0: the solve is over)
__device__ __noinline__ int ctrl_end_batch(const Params& P, Ctrl& c, double tot0, double tot1, int my_tiles, bool lead) {
    const double count = (double)P.rows * kD;
    c.save_slot = -1;
    bool go = true;
    if (c.phase == 0) {
        // init_step, first half (utils.py:40-47)
        const float d0 = sqrtf((float)(tot0 / count)), d1 = sqrtf((float)(tot1 / count));
        c.h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : div_rn(mul_rn(0.01f, d0), d1);
        c.d1 = d1;
        c.nfe = 1;
        c.phase = 1;
        c.dt_attempt = c.h0;
        return my_tiles;
    }
    if (c.phase == 1) {
        // init_step, second half (utils.py:48-54)
        const float d2 = div_rn(sqrtf((float)(tot0 / count)), c.h0);
        float h1;
        if (c.d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, mul_rn(c.h0, 1e-3f));
        else {
            const float base = mul_rn(div_rn(1.f, fmaxf(c.d1, d2)), 0.01f);
            h1 = (float)pow((double)base, 1.0 / (double)(P.order + 1));
        }
        c.dt = fminf(mul_rn(100.f, c.h0), h1);
        c.nfe = 2;
        c.phase = 2;
        if (lead && P.trace_cap > 0) P.trace[0] = c.dt;
        go = c.t < c.T;
    } else {
        // error ratio, accept / reject, output bookkeeping (odeint.py:365-401), adapt_step (utils.py:57-63)
        const float ratio = sqrtf((float)(tot0 / count));
        const float dt = c.dt;
        ++c.attempt;
        c.nfe += P.n_stage - 1;
        if (lead && 3 * c.attempt < P.trace_cap) {
            P.trace[3 * c.attempt - 2] = c.t; P.trace[3 * c.attempt - 1] = dt; P.trace[3 * c.attempt] = ratio;
        }
        if (ratio <= 1.f) {
            ++c.accept;
            const float tn = add_rn(c.t, dt);
            const bool hit = c.ck < P.n_t - 1 && tn == __ldg(P.t_span + 1 + c.ck);
            if (hit || P.return_all) {
                if (c.nout < P.out_cap) {
                    c.save_slot = c.nout;
                    if (lead) P.t_out[c.nout] = tn;
                } else {
                    c.err = 2;
                }
                ++c.nout;
                if (hit) ++c.ck;
            }
            c.t = tn;
            c.xcur ^= 1;
            // FSAL: k1 <- k_last by swapping the two buffers' roles
            c.swapped ^= 1;
        }
        if (c.ck_flag) c.dt = sub_rn(c.dt_keep, dt);             // the remainder of the intended step (:399-401)
        if (ratio == 0.f) c.dt = mul_rn(c.dt, P.max_factor);
        else {
            const float lo = ratio < 1.f ? 1.f : P.min_factor;
            const float pw = (float)pow((double)ratio, (double)div_rn(1.f, (float)P.order));
            const float cand = div_rn(P.safety, pw);
            const float fac = (cand != cand) ? cand : fminf(P.max_factor, fmaxf(cand, lo));
            c.dt = mul_rn(c.dt, fac);
        }
        go = c.t < c.T && c.err == 0;
        if (go && (c.attempt >= P.max_steps || c.dt != c.dt)) { c.err = 1; go = false; }
    }
    if (go) {
        ctrl_begin_attempt(P, c);
        return my_tiles * (P.n_stage - 1);
    }
    if (lead) {
        P.status[0] = c.nout; P.status[1] = c.attempt; P.status[2] = c.accept;
        P.status[3] = c.nfe; P.status[4] = c.err; P.status[5] = c.xcur;
    }
    return 0;
}

// adaptive variant, end of a batch (all worker threads): phase 2: x_new, err and the sum of squared scaled errors
// (odeint.py:365-372); phases 0 / 1: init_step's norms; then ONE grid barrier carries the global sums, the controller
// decides, and an accepted output step is copied out.  A separate (not inlined) function: nothing of it may cost the
// item loop a register.
__device__ __noinline__ void adaptive_batch_end(const Params& P, Ctrl& ctrl, int& batch_n, double* red_sm, int my_tiles) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = (warp - 2) >> 2, row = (warp & 3) * 32 + lane;
    const int phase = ctrl.phase, xcur = ctrl.xcur;
    const bool kmap = uniform_flag(ctrl.swapped != 0);
    const float dt_att = ctrl.dt_attempt;
    const bool second = uniform_flag(xcur != 0);
    const float* xc = second ? P.xbuf[1] : P.xbuf[0];
    float* xn = second ? P.xbuf[0] : P.xbuf[1];
    double a0 = 0.0, a1 = 0.0;
    for (int i = 0; i < my_tiles; ++i) {
        const int64_t grow = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * kTileM + row;
        if (grow < P.rows) {
            const size_t goff = (size_t)grow * kD + sub * 8;
            if (phase == 2) a0 += finish_row8(P, kmap, xc, xn, dt_att, goff);
            else init_row8(P, phase, xc, goff, a0, a1);
        }
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1);
    if (lane == 0) { red_sm[2 * (warp - 2)] = a0; red_sm[2 * (warp - 2) + 1] = a1; }
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (threadIdx.x == 64) {
        double s0 = 0.0, s1 = 0.0;
        for (int q = 0; q < kWorkers / 32; ++q) { s0 += red_sm[2 * q]; s1 += red_sm[2 * q + 1]; }
        // all CTAs are co-resident (cooperative launch)
        double* part = P.partials + (size_t)(ctrl.nbar & 1u) * 2 * gridDim.x;
        part[blockIdx.x] = s0; part[gridDim.x + blockIdx.x] = s1;
        __threadfence();
        atomicAdd(P.gbar, 1u);
        const unsigned int target = (++ctrl.nbar) * gridDim.x;
        for (;;) {
            unsigned int v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(P.gbar) : "memory");
            if (v >= target) break;
            __nanosleep(32);
        }
        double t0 = 0.0, t1 = 0.0;
        for (unsigned int bq = 0; bq < gridDim.x; ++bq) { t0 += __ldcg(part + bq); t1 += __ldcg(part + gridDim.x + bq); }
        // the controller: every CTA computes the same decision from the same totals
        batch_n = ctrl_end_batch(P, ctrl, t0, t1, my_tiles, blockIdx.x == 0);
    }
    asm volatile("bar.sync 1, 512;" ::: "memory");
    const int save_slot = ctrl.save_slot;
    if (save_slot >= 0) {
        // an accepted step that is an output: this thread copies what it wrote itself
        float* dst = P.out + (size_t)save_slot * P.n_state;
        for (int i = 0; i < my_tiles; ++i) {
            const int64_t grow = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * kTileM + row;
            if (grow < P.rows) {
                const size_t goff = (size_t)grow * kD + sub * 8;
                reinterpret_cast<float4*>(dst + goff)[0] = ld_f4(xn + goff);
                reinterpret_cast<float4*>(dst + goff)[1] = ld_f4(xn + goff + 4);
            }
        }
    }
}

// VAR 1 (save): also write the hidden activations (and scale the output) -- the adjoint's forward half.
// VAR 2 (dgrad): the INPUT-GRADIENT chain of the same MLP, which has the forward's shape with transposed weights:
//   delta2 = (g W3) * act'(a2),  delta1 = (delta2 W2) * act'(a1),  out = scale * (delta1 W1)      (32 -> 256 -> 256 -> 32)
// -- the "activation" of a hidden layer is the multiplication with act' of the saved forward activation (read from global
// memory), no biases; delta2 / delta1 are written out for the parameter gradients.  Separate instantiations, so that the
// integration kernels carry none of it (registers are at the limit of 576 threads per SM).
constexpr int kVarPlain = 0, kVarSave = 1, kVarDgrad = 2, kVarAdaptive = 3;
template <int ACT>
__device__ __forceinline__ float dact_from_out(float a) {
    if (ACT == TDYN_ACT_RELU) return a > 0.f ? 1.f : 0.f;
    if (ACT == TDYN_ACT_SOFTPLUS) return -expm1f(-a);       // sigmoid(z) = 1 - exp(-softplus(z))
    return fmaf(-a, a, 1.f);                                  // 1 - tanh^2
}
template <int ACT, int VAR>
__global__ void __launch_bounds__(kThreads, 1) mlp_tc_stage_kernel(const __grid_constant__ Params P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t tmem_base_slot;
    // operand tiles need 1024-byte alignment (128B swizzle atoms): align the dynamic window by hand
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + kSmemBars;
    auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < NUM_BARS; ++i) {
            uint32_t cnt = 1;
            if ((i >= A_FULL0 && i <= A_FULL2) || i == A1_FULL || i == D_FREE) cnt = kWorkers;
            mbar_init(bar(i), cnt);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // biases -> shared memory (the L1 is almost entirely carved out as shared memory: global re-reads would go to L2)
        float* sb = reinterpret_cast<float*>(smem + kSmemBias);
        // tanh: the hidden-layer biases are stored pre-multiplied by 2*log2(e) so that bias add and argument scaling of
        // tanh(z) = 1 - 2/(1 + 2^(2z*log2e)) are one FMA in the epilogue
        const float bscale = ACT == TDYN_ACT_TANH ? kTanhScale : 1.f;
        for (int i = threadIdx.x; i < 2 * kH + kD; i += kThreads)
            sb[i] = i < kH ? bscale * __ldg(P.b1 + i) : (i < 2 * kH ? bscale * __ldg(P.b2 + i - kH) : __ldg(P.b3 + i - 2 * kH));
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    const int my_tiles = (P.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    constexpr bool ADAPT = VAR == kVarAdaptive;
    // Work items of this CTA: pass-major, its tiles within a pass (the launchers guarantee my_tiles >= 2 when there is more
    // than one pass: the stage input of item i + 1 is built while item i is still in flight, so the two must be different
    // tiles).  Items come in BATCHES: one batch = everything whose inputs are known up front -- the whole launch for a
    // stage / a fixed-step solve; ONE ATTEMPTED STEP (stages 2..s) for the adaptive variant, where the next batch exists only
    // after the grid-wide error norm and the step controller have run.  The three roles leave and re-enter the software
    // pipeline at batch boundaries; every mbarrier completes once per item, so their parities follow the global item count.
    __shared__ int batch_n;              // items of the next batch (0: stop); written by the controller thread
    __shared__ double red_sm[2 * kWorkers / 32];
    __shared__ Ctrl ctrl;
    if (ADAPT && threadIdx.x == 0) {
        ctrl.t = __ldg(P.t_span); ctrl.T = __ldg(P.t_span + P.n_t - 1);
        ctrl.ck = 0; ctrl.ck_flag = 0; ctrl.nout = 1; ctrl.attempt = 0; ctrl.accept = 0; ctrl.err = 0; ctrl.nfe = 0;
        ctrl.nbar = 0; ctrl.xcur = 0; ctrl.swapped = 0; ctrl.save_slot = -1; ctrl.dt_attempt = 0.f;
        if (P.t_dt0 > 0.f) {             // k1 and the first step size come from the caller
            ctrl.dt = P.t_dt0; ctrl.phase = 2;
            if (blockIdx.x == 0 && P.trace_cap > 0) P.trace[0] = P.t_dt0;
            ctrl_begin_attempt(P, ctrl);
            batch_n = my_tiles * (P.n_stage - 1);
        } else {
            ctrl.phase = 0;
            batch_n = my_tiles;
        }
    }
    // TMEM map: X = cols [0,256): D1, converted in place to layer 2's A_hi; Y = cols [256,512): D2 -> layer 3's A_hi;
    // D3 = cols [224,256) (layer 2's A_hi is dead once layer 2 has been issued: the tensor pipe executes in order).
    // The NEXT tile's layer 1 is split: columns [0,224) are issued right after this tile's layer 2 (they run under the
    // layer-2 -> 3 epilogue), columns [224,256) once this tile's D3 has been read out.
    constexpr uint32_t kColD1 = 0, kColD2 = 256, kColD3 = kNSplit;

    // ring counters of the producer / MMA / worker roles: functions of the number of items done so far (10 weight stages
    // and 16 operand tiles per item), so that only `gbase` lives across a batch boundary
    WorkerCtx w;
    {
        const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
        w.sub = (warp - 2) >> 2;                       // which 8 of every 32 columns (0..3)
        w.row = quarter * 32 + lane;                   // row inside the tile == TMEM lane
        w.lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    }
    // adaptive variant: what the threads need of the controller state (phase, attempted dt, current state buffer, stage
    // buffer map) is re-read from shared memory at every use -- it is constant during a batch, and the kernel has no
    // registers to spare
    const volatile Ctrl* cv = &ctrl;
    auto item_at = [&](int pss, const Item& fixed) -> Item {
        if constexpr (ADAPT) {
            // the phase decides which plan (Params::st entry, constant bank) the stage input uses: hand it to the compiler
            // as a warp-UNIFORM value (vote results are), so that the plan stays an indexed constant-bank operand -- with
            // a per-thread index the parameter array would be copied to local memory
            const int ph = cv->phase;
            const bool p2 = uniform_flag(ph == 2), p1 = uniform_flag(ph == 1);
            return make_item_adapt(P, p2 ? 2 : (p1 ? 1 : 0), pss, cv->dt_attempt);
        }
        else return fixed;
    };
    auto x_adapt = [&]() -> const float* { return ADAPT ? (uniform_flag(cv->xcur != 0) ? P.xbuf[1] : P.xbuf[0]) : nullptr; };
    auto kmap_now = [&]() -> bool { return ADAPT ? uniform_flag(cv->swapped != 0) : false; };
    // SERIAL mode (adaptive variant, a CTA that owns ONE row tile): consecutive work items are the same tile, so the stage input
    // of item t + 1 needs the output of item t -- it is built after the output instead of under the first epilogue, the next
    // item's layer 1 is issued after this item's layer 3, and the weight stages are streamed in that order (W2, W3, W1).
    // Small batches then run one tile per SM instead of two tiles on half as many SMs.
    const bool serial = ADAPT && my_tiles == 1;
    int gbase = 0;
    for (;;) {
        int n_items;
        if constexpr (ADAPT) {
            __syncthreads();
            n_items = batch_n;
            if (n_items == 0) break;
        } else {
            n_items = my_tiles * P.n_stage * P.n_steps;          // one batch: no back edge in these variants
        }
        if (warp == 0) {
            // ================= weight producer: W1(0) | per item: W2[0..7], W1 + W1B(next item), W3 =================
            if (lane == 0) {
                uint32_t p_it = 10u * (uint32_t)gbase;
                auto load = [&](int s) {
                    const uint32_t slot = p_it & 1, use = p_it >> 1;
                    mbar_wait(bar(W_EMPTY0 + slot), (use & 1) ^ 1);
                    mbar_expect_tx(bar(W_FULL0 + slot), kWStage);
                    const uint32_t dst = sbase + kSmemWRing + slot * kWStage;
                    const uint8_t* src = P.wimg + (size_t)s * kWStage;
                    bulk_g2s(dst, src, kWStage / 2, bar(W_FULL0 + slot));
                    bulk_g2s(dst + kWStage / 2, src + kWStage / 2, kWStage / 2, bar(W_FULL0 + slot));
                    ++p_it;
                };
                auto load_w1b = [&](int gt) {
                    mbar_wait(bar(W1B_EMPTY), (gt & 1) ^ 1);
                    mbar_expect_tx(bar(W1B_FULL), kW1BBytes);
                    bulk_g2s(sbase + kSmemW1B, P.wimg + (size_t)kWStagesPerTile * kWStage, kW1BBytes, bar(W1B_FULL));
                };
                load(0); load_w1b(gbase);
                for (int t = 0; t < n_items; ++t) {
                    for (int s = 1; s <= 8; ++s) load(s);
                    if (!serial && t + 1 < n_items) { load(0); load_w1b(gbase + t + 1); }
                    load(9);
                    if (serial && t + 1 < n_items) { load(0); load_w1b(gbase + t + 1); }
                }
            }
        } else if (warp == 1) {
            // ================= MMA issuer =================
            if (lane == 0) {
                uint32_t m_wit = 10u * (uint32_t)gbase, m_ait = 16u * (uint32_t)gbase;
                const uint32_t a1 = sbase + kSmemA1;
                auto layer1a = [&](int gt) {        // [128 x 32] x [224 x 32]^T -> D1 cols [0,224)
                    mbar_wait(bar(A1_FULL), gt & 1);
                    const uint32_t slot = m_wit & 1, use = m_wit >> 1;
                    mbar_wait(bar(W_FULL0 + slot), use & 1);
                    tc_fence_after();
                    const uint32_t wb = sbase + kSmemWRing + slot * kWStage;
                    issue_kblock(tmem + kColD1, a1, a1 + kATile, wb, wb + kWStage / 2, kNSplit, true);
                    tc_commit(bar(W_EMPTY0 + slot));
                    tc_commit(bar(D1_FULL));
                    ++m_wit;
                };
                auto layer1b = [&](int gt) {        // remaining 32 output columns (W1 rows 224..255 from their own buffer)
                    mbar_wait(bar(W1B_FULL), gt & 1);
                    tc_fence_after();
                    const uint32_t wb = sbase + kSmemW1B;
                    issue_kblock(tmem + kColD1 + kNSplit, a1, a1 + kATile, wb, wb + 4096u, kH - kNSplit, true);
                    tc_commit(bar(W1B_EMPTY));
                    tc_commit(bar(D1B_FULL));
                };
                layer1a(gbase); layer1b(gbase);
                for (int t = 0; t < n_items; ++t) {
                    // ---- layer 2: 8 K-blocks, A_hi = X (in TMEM), A_lo ring -> D2 = Y
                    for (int j = 0; j < kH / kKB; ++j, ++m_wit, ++m_ait) {
                        const uint32_t ws = m_wit & 1, wu = m_wit >> 1, as = m_ait % kLoStages, au = m_ait / kLoStages;
                        mbar_wait(bar(A_FULL0 + as), au & 1);
                        mbar_wait(bar(W_FULL0 + ws), wu & 1);
                        tc_fence_after();
                        const uint32_t wb = sbase + kSmemWRing + ws * kWStage;
                        issue_kblock_ts(tmem + kColD2, tmem + kColD1 + (uint32_t)(j * kKB), sbase + kSmemLoRing + as * kATile,
                                        wb, wb + kWStage / 2, kH, j == 0);
                        tc_commit(bar(W_EMPTY0 + ws));
                        tc_commit(bar(A_EMPTY0 + as));
                    }
                    tc_commit(bar(D2_FULL));
                    // ---- first 224 columns of the next item's layer 1: tensor work for the layer-2 -> 3 epilogue phase
                    if (!serial && t + 1 < n_items) layer1a(gbase + t + 1);
                    // ---- layer 3: 8 K-blocks, N = 32, A_hi = Y -> D3 = X[224,256); W3 stage = 8 x [hi 4 KB | lo 4 KB]
                    {
                        const uint32_t ws = m_wit & 1, wu = m_wit >> 1;
                        mbar_wait(bar(W_FULL0 + ws), wu & 1);
                        const uint32_t wb = sbase + kSmemWRing + ws * kWStage;
                        for (int j = 0; j < kH / kKB; ++j, ++m_ait) {
                            const uint32_t as = m_ait % kLoStages, au = m_ait / kLoStages;
                            mbar_wait(bar(A_FULL0 + as), au & 1);
                            tc_fence_after();
                            const uint32_t wj = wb + (uint32_t)j * 8192u;
                            issue_kblock_ts(tmem + kColD3, tmem + kColD2 + (uint32_t)(j * kKB), sbase + kSmemLoRing + as * kATile,
                                            wj, wj + 4096u, kD, j == 0);
                            tc_commit(bar(A_EMPTY0 + as));
                        }
                        tc_commit(bar(W_EMPTY0 + ws));
                        tc_commit(bar(D3_FULL));
                        ++m_wit;
                    }
                    // ---- the last 32 columns of the next item's layer 1 overwrite D3: wait until it has been read out
                    if (t + 1 < n_items) {
                        mbar_wait(bar(D_FREE), (gbase + t) & 1);
                        if (serial) layer1a(gbase + t + 1);          // (its stage input exists only after this item's output)
                        layer1b(gbase + t + 1);
                    }
                }
            }
        } else {
            // ================= workers: prologue / activation epilogues / output =================
            uint32_t ait = 16u * (uint32_t)gbase;
            int pass = 0, tl = 0;                          // (pass, local tile index) of the current item within the batch
            Item cur = ADAPT ? Item{} : make_item(P, 0);
            prologue(P, item_at(0, cur), !ADAPT && gbase == 0, smem, w, (int64_t)blockIdx.x, x_adapt(), kmap_now());
            mbar_arrive(bar(A1_FULL));
            for (int t = 0; t < n_items; ++t) {
                const uint32_t tpar = (uint32_t)(gbase + t) & 1u;
                const int64_t tile = (int64_t)blockIdx.x + (int64_t)tl * gridDim.x;
                const int64_t grow_cur = (VAR == kVarSave || VAR == kVarDgrad) ? tile * kTileM + w.row : 0;
                const int npass = tl + 1 < my_tiles ? pass : pass + 1, ntl = tl + 1 < my_tiles ? tl + 1 : 0;   // the next item
                const int64_t ntile = (int64_t)blockIdx.x + (int64_t)ntl * gridDim.x;
                const Item nxt = (!ADAPT && t + 1 < n_items && npass != pass) ? make_item(P, npass) : cur;
                if (t + 1 < n_items && w.sub == 0) {
                    // the next item's stage-input rows (x and the k's, 128 B each) are consumed by prologue() half an item
                    // from now: pull them into L2 so that those loads do not pay the DRAM latency
                    const int64_t nrow = ntile * kTileM + w.row;
                    if (nrow < P.rows) {
                        const bool first = !ADAPT && gbase == 0 && npass == 0;
                        const Item ni = item_at(npass, nxt);
                        const StagePlan& np = P.st[ADAPT ? ni.slot : plan_index(P, ni, first)];
                        const float* xs = ADAPT ? x_adapt() : ((first || !P.solve) ? P.x_in : P.xw);
                        const bool kmap = kmap_now();
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(xs + (size_t)nrow * kD));
                        for (int j = 0; j < np.n_k; ++j) asm volatile("prefetch.global.L2 [%0];" ::"l"(kbuf(P, kmap, np.kidx[j]) + (size_t)nrow * kD));
                    }
                }
                // ---- two activation epilogues: D1 -> layer-2 operand, D2 -> layer-3 operand
#pragma unroll 1
                for (int layer = 0; layer < 2; ++layer) {
                    mbar_wait(bar(layer == 0 ? D1_FULL : D2_FULL), tpar);
                    tc_fence_after();
                    const float* bias = reinterpret_cast<const float*>(smem + kSmemBias) + layer * kH + w.sub * 8;
                    const uint32_t src = w.lane_addr + (layer == 0 ? kColD1 : kColD2) + (uint32_t)(w.sub * 8);
                    // (the adaptive variant has no registers for a second buffer: it re-issues the load of the next chunk into
                    // the SAME registers as soon as the activation has consumed them, so that the TMEM read overlaps the
                    // split / store half of the chunk)
                    constexpr bool kPrefetch = !ADAPT;
                    uint32_t ra[8], rb[kPrefetch ? 8 : 1];
                    tmem_ld8(src, ra);
                    if (kPrefetch) tmem_wait8(ra);
#pragma unroll
                    for (int j = 0; j < kH / kKB; ++j, ++ait) {
                        uint32_t (&cur)[8] = (kPrefetch && (j & 1)) ? reinterpret_cast<uint32_t (&)[8]>(rb) : ra;
                        uint32_t (&nxt)[8] = (kPrefetch && !(j & 1)) ? reinterpret_cast<uint32_t (&)[8]>(rb) : ra;
                        if (kPrefetch && j + 1 < kH / kKB) {
                            if (layer == 0 && (j + 1) * kKB == kNSplit) {       // last chunk of D1 is produced by layer1b
                                mbar_wait(bar(D1B_FULL), tpar);
                                tc_fence_after();
                            }
                            tmem_ld8(src + (uint32_t)((j + 1) * kKB), nxt);     // prefetch next chunk
                        }
                        if (!kPrefetch) tmem_wait8(ra);
                        float v[8];
                        if (VAR == kVarDgrad) {
                            // delta = acc * act'(saved activation): epilogue 0 pairs with a2, epilogue 1 with a1
                            float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
                            if (grow_cur < P.rows) {
                                const float4* sp = reinterpret_cast<const float4*>(P.dact_in[layer] + (size_t)grow_cur * kH + j * kKB + w.sub * 8);
                                s0 = __ldg(sp); s1 = __ldg(sp + 1);
                            }
                            v[0] = __uint_as_float(cur[0]) * dact_from_out<ACT>(s0.x); v[1] = __uint_as_float(cur[1]) * dact_from_out<ACT>(s0.y);
                            v[2] = __uint_as_float(cur[2]) * dact_from_out<ACT>(s0.z); v[3] = __uint_as_float(cur[3]) * dact_from_out<ACT>(s0.w);
                            v[4] = __uint_as_float(cur[4]) * dact_from_out<ACT>(s1.x); v[5] = __uint_as_float(cur[5]) * dact_from_out<ACT>(s1.y);
                            v[6] = __uint_as_float(cur[6]) * dact_from_out<ACT>(s1.z); v[7] = __uint_as_float(cur[7]) * dact_from_out<ACT>(s1.w);
                        } else {
                            const float4 b0 = *reinterpret_cast<const float4*>(bias + j * kKB);
                            const float4 b1 = *reinterpret_cast<const float4*>(bias + j * kKB + 4);
                            v[0] = activate_biased<ACT>(__uint_as_float(cur[0]), b0.x); v[1] = activate_biased<ACT>(__uint_as_float(cur[1]), b0.y);
                            v[2] = activate_biased<ACT>(__uint_as_float(cur[2]), b0.z); v[3] = activate_biased<ACT>(__uint_as_float(cur[3]), b0.w);
                            v[4] = activate_biased<ACT>(__uint_as_float(cur[4]), b1.x); v[5] = activate_biased<ACT>(__uint_as_float(cur[5]), b1.y);
                            v[6] = activate_biased<ACT>(__uint_as_float(cur[6]), b1.z); v[7] = activate_biased<ACT>(__uint_as_float(cur[7]), b1.w);
                        }
                        if (!kPrefetch && j + 1 < kH / kKB) {
                            if (layer == 0 && (j + 1) * kKB == kNSplit) {
                                mbar_wait(bar(D1B_FULL), tpar);
                                tc_fence_after();
                            }
                            tmem_ld8(src + (uint32_t)((j + 1) * kKB), ra);      // the activation has consumed ra
                        }
                        if ((VAR == kVarSave || VAR == kVarDgrad) && grow_cur < P.rows) {
                            float4* dst = reinterpret_cast<float4*>(P.act_out[layer] + (size_t)grow_cur * kH + j * kKB + w.sub * 8);
                            dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                            dst[1] = make_float4(v[4], v[5], v[6], v[7]);
                        }
                        const uint32_t as = ait % kLoStages, au = ait / kLoStages;
                        mbar_wait(bar(A_EMPTY0 + as), (au & 1) ^ 1);
                        store_split8_tmem(src + (uint32_t)(j * kKB), smem + kSmemLoRing + as * kATile, w.row, w.sub, v);
                        fence_proxy_async();
                        tc_fence_before();
                        mbar_arrive(bar(A_FULL0 + as));
                        if (kPrefetch && j + 1 < kH / kKB) tmem_wait8(nxt);
                    }
                    if (layer == 0 && t + 1 < n_items && !serial) {
                        // next item's stage input (A1 is free: this item's layer 1 has completed)
                        prologue(P, item_at(npass, nxt), !ADAPT && gbase == 0 && npass == 0, smem, w, ntile, x_adapt(), kmap_now());
                        mbar_arrive(bar(A1_FULL));
                    }
                }
                // ---- output: k_out = D3 + b3 (and x_new for the last stage of a fixed-step method)
                mbar_wait(bar(D3_FULL), tpar);
                tc_fence_after();
                uint32_t ro[8];
                tmem_ld8(w.lane_addr + kColD3 + (uint32_t)(w.sub * 8), ro);
                tmem_wait8(ro);
                tc_fence_before();
                mbar_arrive(bar(D_FREE));
                const int64_t grow = tile * kTileM + w.row;
                if (grow < P.rows) {
                    const size_t goff = (size_t)grow * kD + w.sub * 8;
                    float o[8];
                    float* k_out = ADAPT ? const_cast<float*>(kbuf(P, kmap_now(), item_at(pass, cur).g))
                                         : (P.solve ? const_cast<float*>(P.kp[cur.g]) : P.k_out0);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float4 bq = *reinterpret_cast<const float4*>(smem + kSmemBias + (2 * kH + w.sub * 8 + 4 * i) * 4);
                        // (init_step's trial evaluation is the UN-negated field also on reversed spans: odeint.py:66 hands f, not f_)
                        const float sc = VAR == kVarPlain ? 1.f : ((ADAPT && cv->phase == 1) ? 1.f : P.out_scale);
                        o[4 * i] = sc * (__uint_as_float(ro[4 * i]) + bq.x); o[4 * i + 1] = sc * (__uint_as_float(ro[4 * i + 1]) + bq.y);
                        o[4 * i + 2] = sc * (__uint_as_float(ro[4 * i + 2]) + bq.z); o[4 * i + 3] = sc * (__uint_as_float(ro[4 * i + 3]) + bq.w);
                        reinterpret_cast<float4*>(k_out + goff)[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
                    }
                    if (P.x_new) {
                        // x + dt*(b0*k0 + ... + b_last*k_out)
                        float s[8];
                        for (int j = 0; j < P.n_sol; ++j) {
                            const float b = P.bsol.c[j];
                            const bool last = j == P.n_sol - 1;
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                float4 q;
                                if (last) q = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
                                else q = __ldg(reinterpret_cast<const float4*>(P.kp[j] + goff) + i);
                                const float m0 = mul_rn(b, q.x), m1 = mul_rn(b, q.y), m2 = mul_rn(b, q.z), m3 = mul_rn(b, q.w);
                                s[4 * i] = j == 0 ? m0 : add_rn(s[4 * i], m0);
                                s[4 * i + 1] = j == 0 ? m1 : add_rn(s[4 * i + 1], m1);
                                s[4 * i + 2] = j == 0 ? m2 : add_rn(s[4 * i + 2], m2);
                                s[4 * i + 3] = j == 0 ? m3 : add_rn(s[4 * i + 3], m3);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const float4 xq = __ldg(reinterpret_cast<const float4*>(P.x_in + goff) + i);
                            reinterpret_cast<float4*>(P.x_new + goff)[i] =
                                make_float4(add_rn(xq.x, mul_rn(P.dt0, s[4 * i])), add_rn(xq.y, mul_rn(P.dt0, s[4 * i + 1])),
                                            add_rn(xq.z, mul_rn(P.dt0, s[4 * i + 2])), add_rn(xq.w, mul_rn(P.dt0, s[4 * i + 3])));
                        }
                    }
                }
                if (serial && t + 1 < n_items) {
                    // serial mode: the next item is the next stage of the SAME tile -- its input needs the k just written
                    prologue(P, item_at(npass, nxt), false, smem, w, ntile, x_adapt(), kmap_now());
                    mbar_arrive(bar(A1_FULL));
                }
                pass = npass; tl = ntl; cur = nxt;
            }
            // ================= end of the batch
            if constexpr (ADAPT) adaptive_batch_end(P, ctrl, batch_n, red_sm, my_tiles);
        }
        if constexpr (!ADAPT) break;
        gbase += n_items;
    }
    if (warp >= 2 && !ADAPT && P.solve) {
        // solve mode: the closing combination x_final = x + dt*(sum_j bsol_j k_j) of the last step (no evaluation)
        for (int i = 0; i < my_tiles; ++i) {
            const int64_t grow = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * kTileM + w.row;
            if (grow < P.rows) {
                const size_t goff = (size_t)grow * kD + w.sub * 8;
                float y[8];
                combine_row8(P, P.st[kUpd], P.xw, __ldg(P.dts + P.n_steps - 1), goff, true, y);
                store_row8(P.out + (size_t)__ldg(P.slots + P.n_steps) * P.n_state, goff, y);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
    }
}

// ------------------------------------------------------------------------------------------- weight images
// stage s of the image: s = 0: W1 [256 x 32]; s = 1..8: W2[:, 32(s-1) .. 32 s); s = 9: W3 [32 x 256] as 8 K-blocks.
__global__ void prepare_weights_kernel(const float* __restrict__ w1, const float* __restrict__ w2,
                                       const float* __restrict__ w3, uint8_t* __restrict__ img) {
    const int total = kD * kH + kH * kH + kH * kD;
    // per K-block slab: TF32 part [n][32 k] fp32 (row = 128 B) at `hi`, correction part W' [n][64 k'] bf16 (row = 128 B) at
    // `corr`: k' = kk holds bf16(w_lo) (pairs with bf16(a_hi)), k' = 32 + kk holds bf16(w_hi) (pairs with bf16(a_lo))
    auto put = [&](size_t hi, size_t corr, int n, int kk, float w) {
        const float h = tf32_rna(w);
        *reinterpret_cast<float*>(img + hi + swz(n, kk >> 2) + (kk & 3) * 4) = h;
        *reinterpret_cast<__nv_bfloat16*>(img + corr + swz(n, kk >> 3) + (kk & 7) * 2) = __float2bfloat16_rn(w - h);
        *reinterpret_cast<__nv_bfloat16*>(img + corr + swz(n, 4 + (kk >> 3)) + (kk & 7) * 2) = __float2bfloat16_rn(h);
    };
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        if (idx < kD * kH) {                       // W1[n][k], n < 256, k < 32
            const int n = idx / kD, k = idx % kD;
            put(0, kWStage / 2, n, k, w1[idx]);
            if (n >= kNSplit) {                    // duplicate rows 224..255 into the W1B slab (stage index 10)
                const size_t o = (size_t)kWStagesPerTile * kWStage;
                put(o, o + 4096, n - kNSplit, k, w1[idx]);
            }
        } else if (idx < kD * kH + kH * kH) {      // W2[n][k]
            const int e = idx - kD * kH, n = e / kH, k = e % kH, kb = k / kKB, kk = k % kKB;
            const size_t o = (size_t)(1 + kb) * kWStage;
            put(o, o + kWStage / 2, n, kk, w2[e]);
        } else {                                   // W3[n][k], n < 32, k < 256
            const int e = idx - kD * kH - kH * kH, n = e / kH, k = e % kH, kb = k / kKB, kk = k % kKB;
            const size_t o = (size_t)9 * kWStage + (size_t)kb * 8192;
            put(o, o + 4096, n, kk, w3[e]);
        }
    }
}

}  // namespace tc
}  // namespace tdyn

using namespace tdyn;

template <int VAR>
static int launch_tc_t(const tdyn_mlp_t* m, tc::Params& P, int grid, cudaStream_t st) {
    static thread_local unsigned long long seen = 0;
    if (first_use_on_device(seen)) {
        TDYN_CHECK_CUDA(cudaFuncSetAttribute(tc::mlp_tc_stage_kernel<TDYN_ACT_TANH, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemTotal));
        TDYN_CHECK_CUDA(cudaFuncSetAttribute(tc::mlp_tc_stage_kernel<TDYN_ACT_SOFTPLUS, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemTotal));
        TDYN_CHECK_CUDA(cudaFuncSetAttribute(tc::mlp_tc_stage_kernel<TDYN_ACT_RELU, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemTotal));
    }
    switch (m->act) {
        case TDYN_ACT_TANH: tc::mlp_tc_stage_kernel<TDYN_ACT_TANH, VAR><<<grid, tc::kThreads, tc::kSmemTotal, st>>>(P); break;
        case TDYN_ACT_SOFTPLUS: tc::mlp_tc_stage_kernel<TDYN_ACT_SOFTPLUS, VAR><<<grid, tc::kThreads, tc::kSmemTotal, st>>>(P); break;
        default: tc::mlp_tc_stage_kernel<TDYN_ACT_RELU, VAR><<<grid, tc::kThreads, tc::kSmemTotal, st>>>(P); break;
    }
    TDYN_CHECK_CUDA(cudaGetLastError());
    return 0;
}
static int launch_tc(const tdyn_mlp_t* m, tc::Params& P, int grid, cudaStream_t st) {
    if (P.dact_in[0]) return launch_tc_t<tc::kVarDgrad>(m, P, grid, st);
    return P.act_out[0] ? launch_tc_t<tc::kVarSave>(m, P, grid, st) : launch_tc_t<tc::kVarPlain>(m, P, grid, st);
}

extern "C" {

int tdyn_mlp_tc_supported(const tdyn_mlp_t* m) {
    if (!m) return 0;
    return m->n_layers == 3 && m->dims[0] == tc::kD && m->dims[1] == tc::kH && m->dims[2] == tc::kH && m->dims[3] == tc::kD &&
           (m->act == TDYN_ACT_TANH || m->act == TDYN_ACT_SOFTPLUS || m->act == TDYN_ACT_RELU);
}

int64_t tdyn_mlp_tc_prepared_bytes(const tdyn_mlp_t* m) {
    return tdyn_mlp_tc_supported(m) ? (int64_t)tc::kWStagesPerTile * tc::kWStage + tc::kW1BBytes : 0;
}

int tdyn_mlp_tc_prepare(const tdyn_mlp_t* m, void* wsplit, void* stream) {
    TDYN_REQUIRE(tdyn_mlp_tc_supported(m), "tdyn_mlp_tc_prepare: unsupported MLP shape (need 32-256-256-32)");
    TDYN_REQUIRE(wsplit && ((uintptr_t)wsplit & 15) == 0, "tdyn_mlp_tc_prepare: wsplit must be 16-byte aligned");
    tc::prepare_weights_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(m->w[0], m->w[1], m->w[2], (uint8_t*)wsplit);
    TDYN_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int tdyn_mlp_tc_stage(const tdyn_mlp_t* m, const void* wsplit, float* k_out, float* y_out, const float* x,
                      const float* const* k, const float* coef, int n_k, int mode, float dt, float* x_new,
                      const float* bsol, int n_sol, int64_t rows, void* stream) {
    TDYN_REQUIRE(tdyn_mlp_tc_supported(m), "tdyn_mlp_tc_stage: unsupported MLP shape (need 32-256-256-32)");
    TDYN_REQUIRE(wsplit && k_out && x, "tdyn_mlp_tc_stage: null argument");
    TDYN_REQUIRE(y_out == nullptr, "tdyn_mlp_tc_stage: y_out is not produced by this kernel (pass NULL)");
    TDYN_REQUIRE(n_k >= 0 && n_k <= TDYN_MAX_STAGES && n_sol >= 0 && n_sol <= TDYN_MAX_STAGES, "tdyn_mlp_tc_stage: bad n_k / n_sol");
    TDYN_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)k_out & 15) == 0 && (!x_new || ((uintptr_t)x_new & 15) == 0),
                 "tdyn_mlp_tc_stage: state pointers must be 16-byte aligned");
    if (rows <= 0) return 0;
    tc::Params P{};
    P.wimg = (const uint8_t*)wsplit;
    P.b1 = m->b[0]; P.b2 = m->b[1]; P.b3 = m->b[2];
    P.k_out0 = k_out; P.x_in = x; P.dt0 = dt;
    P.st[0].n_k = n_k; P.st[0].mode = mode;
    const int need_k = x_new ? (n_sol - 1 > n_k ? n_sol - 1 : n_k) : n_k;
    for (int j = 0; j < need_k; ++j) {
        TDYN_REQUIRE(k && k[j] && ((uintptr_t)k[j] & 15) == 0, "tdyn_mlp_tc_stage: k[%d] null or misaligned", j);
        P.kp[j] = k[j];
    }
    for (int j = 0; j < n_k; ++j) { P.st[0].coef[j] = coef[j]; P.st[0].kidx[j] = j; }
    P.n_stage = 1; P.n_steps = 1; P.solve = 0; P.out_scale = 1.f;
    P.x_new = x_new; P.n_sol = x_new ? n_sol : 0;
    for (int j = 0; j < P.n_sol; ++j) P.bsol.c[j] = bsol[j];
    P.rows = rows;
    P.num_tiles = (int)((rows + tc::kTileM - 1) / tc::kTileM);
    P.act = m->act;
    const int grid = P.num_tiles < kNumSMs ? P.num_tiles : kNumSMs;
    return launch_tc(m, P, grid, (cudaStream_t)stream);
}

int tdyn_mlp_tc_forward_save(const tdyn_mlp_t* m, const void* wsplit, float* out, const float* x, float scale, float* a1,
                             float* a2, int64_t rows, void* stream) {
    TDYN_REQUIRE(tdyn_mlp_tc_supported(m), "tdyn_mlp_tc_forward_save: unsupported MLP shape (need 32-256-256-32)");
    TDYN_REQUIRE(wsplit && out && x && a1 && a2, "tdyn_mlp_tc_forward_save: null argument");
    TDYN_REQUIRE((((uintptr_t)x | (uintptr_t)out | (uintptr_t)a1 | (uintptr_t)a2) & 15) == 0,
                 "tdyn_mlp_tc_forward_save: pointers must be 16-byte aligned");
    if (rows <= 0) return 0;
    tc::Params P{};
    P.wimg = (const uint8_t*)wsplit;
    P.b1 = m->b[0]; P.b2 = m->b[1]; P.b3 = m->b[2];
    P.k_out0 = out; P.x_in = x; P.dt0 = 0.f;
    P.st[0].n_k = 0; P.st[0].mode = TDYN_MODE_CHAIN;
    P.n_stage = 1; P.n_steps = 1; P.solve = 0; P.out_scale = scale;
    P.act_out[0] = a1; P.act_out[1] = a2;
    P.rows = rows;
    P.num_tiles = (int)((rows + tc::kTileM - 1) / tc::kTileM);
    P.act = m->act;
    const int grid = P.num_tiles < kNumSMs ? P.num_tiles : kNumSMs;
    return launch_tc(m, P, grid, (cudaStream_t)stream);
}

int tdyn_mlp_tc_dgrad(const tdyn_mlp_t* mt, const void* wsplit_t, float* out, const float* g, float scale, const float* a2,
                      const float* a1, float* d2, float* d1, int64_t rows, void* stream) {
    TDYN_REQUIRE(tdyn_mlp_tc_supported(mt), "tdyn_mlp_tc_dgrad: unsupported MLP shape (need 32-256-256-32)");
    TDYN_REQUIRE(wsplit_t && out && g && a1 && a2 && d1 && d2, "tdyn_mlp_tc_dgrad: null argument");
    TDYN_REQUIRE((((uintptr_t)g | (uintptr_t)out | (uintptr_t)a1 | (uintptr_t)a2 | (uintptr_t)d1 | (uintptr_t)d2) & 15) == 0,
                 "tdyn_mlp_tc_dgrad: pointers must be 16-byte aligned");
    if (rows <= 0) return 0;
    tc::Params P{};
    P.wimg = (const uint8_t*)wsplit_t;
    P.b1 = mt->b[0]; P.b2 = mt->b[1]; P.b3 = mt->b[2];        // zero vectors (the chain has no biases)
    P.k_out0 = out; P.x_in = g; P.dt0 = 0.f;
    P.st[0].n_k = 0; P.st[0].mode = TDYN_MODE_CHAIN;
    P.n_stage = 1; P.n_steps = 1; P.solve = 0; P.out_scale = scale;
    P.act_out[0] = d2; P.act_out[1] = d1;
    P.dact_in[0] = a2; P.dact_in[1] = a1;
    P.rows = rows;
    P.num_tiles = (int)((rows + tc::kTileM - 1) / tc::kTileM);
    P.act = mt->act;
    const int grid = P.num_tiles < kNumSMs ? P.num_tiles : kNumSMs;
    return launch_tc(mt, P, grid, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------- whole fixed-step solve
int tdyn_mlp_tc_solve_supported(const tdyn_mlp_t* m, int64_t rows) {
    // every CTA must own at least two row tiles (the stage input of the next work item is built while the current one is
    // in flight: they must be different tiles)
    return tdyn_mlp_tc_supported(m) && rows > 2 * tc::kTileM;
}

int64_t tdyn_mlp_tc_solve_workspace_bytes(int n_steps, int n_stages) {
    if (n_steps <= 0 || n_stages <= 0 || n_stages > TDYN_MAX_STAGES) return 0;
    return (int64_t)(((size_t)n_steps * 4 + 15) / 16 * 16 + ((size_t)n_steps + 1) * 4);       // dts | slots
}

int tdyn_mlp_tc_solve_fixed(const tdyn_mlp_t* m, const void* wsplit, const tdyn_rk_plan_t* plan, const float* x_in, float* xw,
                            float* k, float* out, const float* dts, const int* save_slot, int n_steps, int64_t rows,
                            void* workspace, void* stream) {
    TDYN_REQUIRE(m && plan && wsplit && x_in && xw && k && out && dts && save_slot && workspace, "tdyn_mlp_tc_solve_fixed: null argument");
    TDYN_REQUIRE(tdyn_mlp_tc_solve_supported(m, rows), "tdyn_mlp_tc_solve_fixed: unsupported MLP shape or too few rows");
    TDYN_REQUIRE(plan->n_err == 0 && plan->n_extra >= 0 && plan->n_extra <= 6 && plan->n_sol >= 1 && plan->n_sol <= plan->n_extra + 1,
                 "tdyn_mlp_tc_solve_fixed: need a fixed-step plan");
    TDYN_REQUIRE(n_steps >= 1, "tdyn_mlp_tc_solve_fixed: n_steps must be positive");
    TDYN_REQUIRE(save_slot[n_steps] >= 0, "tdyn_mlp_tc_solve_fixed: the final state must have an output slot");
    TDYN_REQUIRE((((uintptr_t)x_in | (uintptr_t)xw | (uintptr_t)k | (uintptr_t)out | (uintptr_t)workspace) & 15) == 0,
                 "tdyn_mlp_tc_solve_fixed: pointers must be 16-byte aligned");
    const int n_stage = plan->n_extra + 1;
    const size_t n_state = (size_t)rows * tc::kD;
    tc::Params P{};
    // terms with an exactly zero coefficient are dropped (they add +-0); the solution update x + dt*(sum_j bsol_j k_j) of
    // step s is the first stage input of step s + 1 (and the closing combination after the last step)
    auto fill = [&](tc::StagePlan& sp, int mode, int n, const float* c) {
        sp.n_k = 0; sp.mode = mode;
        for (int j = 0; j < n; ++j)
            if (c[j] != 0.f) { sp.kidx[sp.n_k] = j; sp.coef[sp.n_k] = c[j]; ++sp.n_k; }
    };
    for (int g = 1; g < n_stage; ++g) {
        TDYN_REQUIRE(plan->nk[g - 1] >= 0 && plan->nk[g - 1] <= g, "tdyn_mlp_tc_solve_fixed: stage %d combines later stages", g);
        fill(P.st[g], plan->mode[g - 1], plan->nk[g - 1], plan->a[g - 1]);
    }
    fill(P.st[tc::kUpd], plan->n_sol > 1 ? TDYN_MODE_INNER : TDYN_MODE_CHAIN, plan->n_sol, plan->bsol);
    for (int g = 0; g < n_stage; ++g) P.kp[g] = k + (size_t)g * n_state;
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* ws = (uint8_t*)workspace;
    const size_t off_slots = ((size_t)n_steps * 4 + 15) / 16 * 16;
    // (pageable sources: the copies are staged before the calls return)
    TDYN_CHECK_CUDA(cudaMemcpyAsync(ws, dts, sizeof(float) * (size_t)n_steps, cudaMemcpyHostToDevice, st));
    TDYN_CHECK_CUDA(cudaMemcpyAsync(ws + off_slots, save_slot, sizeof(int) * ((size_t)n_steps + 1), cudaMemcpyHostToDevice, st));
    P.wimg = (const uint8_t*)wsplit;
    P.b1 = m->b[0]; P.b2 = m->b[1]; P.b3 = m->b[2];
    P.x_in = x_in; P.xw = xw; P.out = out;
    P.dts = (const float*)ws; P.slots = (const int*)(ws + off_slots);
    P.n_state = n_state; P.n_stage = n_stage; P.n_steps = n_steps; P.solve = 1; P.out_scale = 1.f;
    P.rows = rows;
    P.num_tiles = (int)((rows + tc::kTileM - 1) / tc::kTileM);
    P.act = m->act;
    int grid = P.num_tiles / 2;
    if (grid > kNumSMs) grid = kNumSMs;
    return launch_tc(m, P, grid, st);
}

// ------------------------------------------------------------------------------------------- whole adaptive solve
int64_t tdyn_mlp_tc_adaptive_workspace_bytes(void) {
    return 4096 /* t_span */ + 256 /* grid barrier */ + (int64_t)sizeof(double) * 2 * 2 * kNumSMs;
}

int tdyn_mlp_tc_solve_adaptive(const tdyn_mlp_t* m, const void* wsplit, const tdyn_rk_plan_t* plan, float sign, float atol,
                               float rtol, const float* t_span, int n_t, int return_all, float dt0, float* x0, float* x1,
                               float* k, float* out, float* t_out, int out_cap, void* workspace, int* status, float* trace,
                               int trace_cap, int64_t rows, int max_steps, void* stream) {
    TDYN_REQUIRE(m && plan && wsplit && t_span && x0 && x1 && k && out && t_out && workspace && status,
                 "tdyn_mlp_tc_solve_adaptive: null argument");
    TDYN_REQUIRE(tdyn_mlp_tc_supported(m) && rows >= 1, "tdyn_mlp_tc_solve_adaptive: unsupported MLP shape or no rows");
    TDYN_REQUIRE(plan->n_extra >= 1 && plan->n_extra <= TDYN_MAX_STAGES - 1 && plan->n_err == plan->n_extra + 1 &&
                     plan->n_sol >= 1 && plan->n_sol <= plan->n_extra + 1,
                 "tdyn_mlp_tc_solve_adaptive: need an embedded FSAL pair (dopri5 / tsit5)");
    TDYN_REQUIRE(n_t >= 2 && n_t <= 1024 && t_span[n_t - 1] > t_span[0], "tdyn_mlp_tc_solve_adaptive: need 2..1024 increasing times");
    TDYN_REQUIRE(out_cap >= 1 && max_steps >= 1, "tdyn_mlp_tc_solve_adaptive: out_cap and max_steps must be positive");
    TDYN_REQUIRE((((uintptr_t)x0 | (uintptr_t)x1 | (uintptr_t)k | (uintptr_t)out | (uintptr_t)workspace) & 15) == 0,
                 "tdyn_mlp_tc_solve_adaptive: pointers must be 16-byte aligned");
    const int n_stage = plan->n_extra + 1;
    const size_t n_state = (size_t)rows * tc::kD;
    tc::Params P{};
    auto fill = [&](tc::StagePlan& sp, int mode, int n, const float* c, bool drop_zeros) {
        sp.n_k = 0; sp.mode = mode;
        for (int j = 0; j < n; ++j)
            if (!drop_zeros || c[j] != 0.f) { sp.kidx[sp.n_k] = j; sp.coef[sp.n_k] = c[j]; ++sp.n_k; }
    };
    for (int g = 1; g < n_stage; ++g) {
        TDYN_REQUIRE(plan->nk[g - 1] >= 0 && plan->nk[g - 1] <= g, "tdyn_mlp_tc_solve_adaptive: stage %d combines later stages", g);
        fill(P.st[g], plan->mode[g - 1], plan->nk[g - 1], plan->a[g - 1], true);
    }
    const float one = 1.f;
    fill(P.st[0], TDYN_MODE_CHAIN, 1, &one, false);                       // init_step's trial point x0 + h0*k1
    fill(P.st[tc::kUpd], TDYN_MODE_INNER, plan->n_sol, plan->bsol, false);
    for (int g = 0; g < n_stage; ++g) P.kp[g] = k + (size_t)g * n_state;
    P.n_err = plan->n_err;
    for (int j = 0; j < plan->n_err; ++j) P.berr.c[j] = plan->berr[j];
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* ws = (uint8_t*)workspace;
    TDYN_CHECK_CUDA(cudaMemcpyAsync(ws, t_span, sizeof(float) * (size_t)n_t, cudaMemcpyHostToDevice, st));
    TDYN_CHECK_CUDA(cudaMemsetAsync(ws + 4096, 0, 256, st));
    TDYN_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int) * 6, st));
    // the first output is the initial state at t_span[0]
    TDYN_CHECK_CUDA(cudaMemcpyAsync(out, x0, sizeof(float) * n_state, cudaMemcpyDeviceToDevice, st));
    TDYN_CHECK_CUDA(cudaMemcpyAsync(t_out, ws, sizeof(float), cudaMemcpyDeviceToDevice, st));
    P.wimg = (const uint8_t*)wsplit;
    P.b1 = m->b[0]; P.b2 = m->b[1]; P.b3 = m->b[2];
    P.xbuf[0] = x0; P.xbuf[1] = x1; P.out = out; P.t_out = t_out; P.status = status; P.trace = trace;
    P.trace_cap = trace ? trace_cap : 0;
    P.t_span = (const float*)ws; P.n_t = n_t; P.return_all = return_all; P.out_cap = out_cap; P.max_steps = max_steps;
    P.order = plan->order; P.t_dt0 = dt0; P.atol = atol; P.rtol = rtol;
    P.safety = plan->safety; P.min_factor = plan->min_factor; P.max_factor = plan->max_factor;
    P.gbar = (unsigned int*)(ws + 4096); P.partials = (double*)(ws + 4096 + 256);
    P.n_state = n_state; P.n_stage = n_stage; P.n_steps = 1; P.solve = 1; P.out_scale = sign;
    P.rows = rows;
    P.num_tiles = (int)((rows + tc::kTileM - 1) / tc::kTileM);
    P.act = m->act;
    // one CTA per row tile up to the SM count (a CTA with a single tile runs in serial mode, see the kernel)
    int grid = P.num_tiles < kNumSMs ? P.num_tiles : kNumSMs;
    static thread_local unsigned long long seen = 0;
    if (first_use_on_device(seen)) {
        TDYN_CHECK_CUDA(cudaFuncSetAttribute(tc::mlp_tc_stage_kernel<TDYN_ACT_TANH, tc::kVarAdaptive>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemTotal));
        TDYN_CHECK_CUDA(cudaFuncSetAttribute(tc::mlp_tc_stage_kernel<TDYN_ACT_SOFTPLUS, tc::kVarAdaptive>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemTotal));
        TDYN_CHECK_CUDA(cudaFuncSetAttribute(tc::mlp_tc_stage_kernel<TDYN_ACT_RELU, tc::kVarAdaptive>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemTotal));
    }
    const void* fn = m->act == TDYN_ACT_TANH ? (const void*)tc::mlp_tc_stage_kernel<TDYN_ACT_TANH, tc::kVarAdaptive>
                   : m->act == TDYN_ACT_SOFTPLUS ? (const void*)tc::mlp_tc_stage_kernel<TDYN_ACT_SOFTPLUS, tc::kVarAdaptive>
                                                 : (const void*)tc::mlp_tc_stage_kernel<TDYN_ACT_RELU, tc::kVarAdaptive>;
    void* args[] = {(void*)&P};
    // cooperative: the in-kernel grid barrier needs every CTA resident (one per SM)
    TDYN_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(tc::kThreads), args, tc::kSmemTotal, st));
    return 0;
}

}  // extern "C"
