// K2: tensor-core (tcgen05 / TMEM, 3xTF32) stage kernel for wide MLP vector fields on sm_100a.
//
//   y     = combine(x, k[], coef, mode, dt)                  (reference numerics/solvers/ode.py:100-103,148-153)
//   k_out = W3 act(W2 act(W1 y + b1) + b2) + b3              (the nn.Sequential vector field, f(t, y))
//   x_new = x + dt*(sum_j bsol_j k_j), k_out last            (ode.py:104, optional, last stage of a fixed-step method)
//
// Shape class (BASELINE.json configs[1]): state dim D = 32, two hidden layers of width 256, three Linear layers.
// One persistent CTA per SM walks 128-row tiles of the batch.  Per tile:
//   workers (8 warps)  : load x/k rows, combine in the reference's rounding order, split y into TF32 hi/lo and
//                        store it K-major / 128B-swizzled in shared memory (the UMMA canonical layout);
//   MMA warp (1 thread): layer 1 as tcgen05.mma kind::tf32  D1[128x256] (TMEM cols 0..255)
//                        = A_hi*B_hi + A_lo*B_hi + A_hi*B_lo  (3xTF32: fp32-accurate products, fp32 accumulate);
//   workers            : D1 -> registers (tcgen05.ld), +bias, activation, hi/lo split -> next layer's A operand,
//                        one 32-wide K-block at a time through a 2-stage ring, so layer 2's MMAs (TMEM cols
//                        256..511) start while later K-blocks are still being produced; same again for layer 3
//                        (N = 32, TMEM cols 0..31, D1 is dead by then);
//   producer (1 thread): streams the pre-split, pre-swizzled weight images (tdyn_mlp_tc_prepare) from L2 into a
//                        2 x 64 KB ring with cp.async.bulk + mbarrier complete_tx (weights total 640 KB per tile pass:
//                        they cannot stay resident in 227 KB of shared memory).
// Algorithmic work: 163 840 FLOP per sample-NFE (x3 issued as TF32).  HBM traffic is negligible (128-256 B per
// sample-NFE), the bound is the tensor pipe with the activation epilogue (512 tanh per sample-NFE) next.
#include "tdyn_common.cuh"

namespace tdyn {
namespace tc {

constexpr int kD = 32;          // state dim
constexpr int kH = 256;         // hidden width
constexpr int kTileM = 128;     // rows per tile = TMEM lanes
constexpr int kWorkers = 256;   // 8 epilogue/prologue warps
constexpr int kThreads = 64 + kWorkers;
constexpr int kKB = 32;         // K-block: 32 fp32 = 128 B = one swizzle row
constexpr uint32_t kATile = kTileM * 128;           // 16 KB: one [128 x 32] operand tile
constexpr uint32_t kAStage = 2 * kATile;            // hi + lo
constexpr uint32_t kWStage = 65536;                 // one weight stage image
constexpr int kWStagesPerTile = 10;                 // W1 | W2 k-blocks 0..7 | W3 (all k-blocks)
constexpr uint32_t kSmemA1 = 0;
constexpr uint32_t kSmemARing = kSmemA1 + kAStage;                  // 2 stages
constexpr uint32_t kSmemWRing = kSmemARing + 2 * kAStage;           // 2 stages
constexpr uint32_t kSmemBars = kSmemWRing + 2 * kWStage;
constexpr uint32_t kSmemTotal = kSmemBars + 256 + 1024;   // + slack for the manual 1024-byte alignment
constexpr uint32_t kTmemCols = 512;

enum Bar { W_FULL0 = 0, W_FULL1, W_EMPTY0, W_EMPTY1, A_FULL0, A_FULL1, A_EMPTY0, A_EMPTY1, A1_FULL, D1_FULL, D2_FULL,
           D3_FULL, D_FREE, NUM_BARS };

struct Params {
    const uint8_t* wimg;       // 10 x 64 KB stage images
    const float* b1; const float* b2; const float* b3;
    float* k_out;
    const float* x;
    KPtrs k; Coefs coef; int n_k; int mode; float dt;
    float* x_new; Coefs bsol; int n_sol;
    int64_t rows; int num_tiles; int act;
};

// ---------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// 16 consecutive fp32 columns of this warp's 32 TMEM lanes -> 16 registers per thread (thread = lane = row)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor: K-major operand tile, 128-byte swizzle, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address  [0,14)
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused with swizzle) [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell) [46,48)
    d |= (uint64_t)2 << 61;                             // layout: SWIZZLE_128B [61,64)
    return d;
}
// instruction descriptor: D fp32, A/B tf32, both K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

__device__ __forceinline__ float tf32_rna(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// byte offset of (row r, 16-byte chunk c) inside a [rows x 128 B] swizzled operand tile
__host__ __device__ __forceinline__ uint32_t swz(uint32_t r, uint32_t c) {
    return (r >> 3) * 1024u + (r & 7u) * 128u + ((c ^ (r & 7u)) << 4);
}

template <int ACT>
__device__ __forceinline__ float activate(float z) {
    if (ACT == TDYN_ACT_RELU) return fmaxf(z, 0.f);
    if (ACT == TDYN_ACT_SOFTPLUS) return z > 20.f ? z : log1pf(__expf(z));
    // tanh with ~2e-7 relative error everywhere: odd polynomial near 0, 1 - 2/(1+e^{2z}) elsewhere (2 MUFU).
    const float z2 = z * z;
    float p = fmaf(z2, 2.1869488e-2f, -5.3968254e-2f);       // 62/2835, -17/315
    p = fmaf(z2, p, 1.3333333e-1f);                          // 2/15
    p = fmaf(z2, p, -3.3333334e-1f);                         // -1/3
    const float small = fmaf(z * z2, p, z);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * 2.8853900817779268f));
    const float big = 1.f - __fdividef(2.f, 1.f + e);
    return fabsf(z) < 0.25f ? small : big;
}

// 16 activations of one row -> hi/lo TF32 pieces into the [128 x 32] operand tiles at `stage` (hi) / +16 KB (lo)
__device__ __forceinline__ void store_split16(uint8_t* stage, int row, int half, const float (&a)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 hi, lo;
        hi.x = tf32_rna(a[4 * i + 0]); lo.x = tf32_rna(a[4 * i + 0] - hi.x);
        hi.y = tf32_rna(a[4 * i + 1]); lo.y = tf32_rna(a[4 * i + 1] - hi.y);
        hi.z = tf32_rna(a[4 * i + 2]); lo.z = tf32_rna(a[4 * i + 2] - hi.z);
        hi.w = tf32_rna(a[4 * i + 3]); lo.w = tf32_rna(a[4 * i + 3] - hi.w);
        const uint32_t off = swz((uint32_t)row, (uint32_t)(half * 4 + i));
        *reinterpret_cast<float4*>(stage + off) = hi;
        *reinterpret_cast<float4*>(stage + kATile + off) = lo;
    }
}

// one 3xTF32 K-block (32 wide): D (+)= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi, 4 MMAs of K = 8 per term
__device__ __forceinline__ void issue_kblock(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                             uint32_t idesc, bool first) {
    const uint32_t A[3] = {a_lo, a_hi, a_hi};
    const uint32_t B[3] = {b_hi, b_lo, b_hi};
#pragma unroll
    for (int t = 0; t < 3; ++t) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            tc_mma_tf32(d_tmem, umma_desc(A[t] + k * 32), umma_desc(B[t] + k * 32), idesc, (first && t == 0 && k == 0) ? 0u : 1u);
        }
    }
}

template <int ACT>
__global__ void __launch_bounds__(kThreads, 1) mlp_tc_stage_kernel(const Params P) {
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
            if (i == A_FULL0 || i == A_FULL1 || i == A1_FULL || i == D_FREE) cnt = kWorkers;
            mbar_init(bar(i), cnt);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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

    if (warp == 0) {
        // ================= weight producer =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = 0; t < my_tiles; ++t) {
                for (int s = 0; s < kWStagesPerTile; ++s, ++it) {
                    const uint32_t slot = it & 1, use = it >> 1;
                    mbar_wait(bar(W_EMPTY0 + slot), (use & 1) ^ 1);
                    mbar_expect_tx(bar(W_FULL0 + slot), kWStage);
                    const uint32_t dst = sbase + kSmemWRing + slot * kWStage;
                    const uint8_t* src = P.wimg + (size_t)s * kWStage;
                    bulk_g2s(dst, src, kWStage / 2, bar(W_FULL0 + slot));
                    bulk_g2s(dst + kWStage / 2, src + kWStage / 2, kWStage / 2, bar(W_FULL0 + slot));
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t wit = 0, ait = 0;
            const uint32_t a1 = sbase + kSmemA1;
            for (int t = 0; t < my_tiles; ++t) {
                const uint32_t tpar = t & 1;
                if (t > 0) mbar_wait(bar(D_FREE), (t - 1) & 1);       // previous tile's TMEM fully read
                // ---- layer 1: [128 x 32] x [256 x 32]^T -> D1 (cols 0..255)
                mbar_wait(bar(A1_FULL), tpar);
                {
                    const uint32_t slot = wit & 1, use = wit >> 1;
                    mbar_wait(bar(W_FULL0 + slot), use & 1);
                    tc_fence_after();
                    const uint32_t w = sbase + kSmemWRing + slot * kWStage;
                    issue_kblock(tmem + 0, a1, a1 + kATile, w, w + kWStage / 2, umma_idesc(kH), true);
                    tc_commit(bar(W_EMPTY0 + slot));
                    tc_commit(bar(D1_FULL));
                    ++wit;
                }
                // ---- layer 2: 8 K-blocks -> D2 (cols 256..511)
                for (int j = 0; j < kH / kKB; ++j, ++wit, ++ait) {
                    const uint32_t ws = wit & 1, wu = wit >> 1, as = ait & 1, au = ait >> 1;
                    mbar_wait(bar(A_FULL0 + as), au & 1);
                    mbar_wait(bar(W_FULL0 + ws), wu & 1);
                    tc_fence_after();
                    const uint32_t a = sbase + kSmemARing + as * kAStage;
                    const uint32_t w = sbase + kSmemWRing + ws * kWStage;
                    issue_kblock(tmem + 256, a, a + kATile, w, w + kWStage / 2, umma_idesc(kH), j == 0);
                    tc_commit(bar(W_EMPTY0 + ws));
                    tc_commit(bar(A_EMPTY0 + as));
                }
                tc_commit(bar(D2_FULL));
                // ---- layer 3: 8 K-blocks, N = 32 -> D3 (cols 0..31); W3 stage = 8 x [hi 4 KB | lo 4 KB]
                {
                    const uint32_t ws = wit & 1, wu = wit >> 1;
                    mbar_wait(bar(W_FULL0 + ws), wu & 1);
                    const uint32_t w = sbase + kSmemWRing + ws * kWStage;
                    for (int j = 0; j < kH / kKB; ++j, ++ait) {
                        const uint32_t as = ait & 1, au = ait >> 1;
                        mbar_wait(bar(A_FULL0 + as), au & 1);
                        tc_fence_after();
                        const uint32_t a = sbase + kSmemARing + as * kAStage;
                        const uint32_t wj = w + (uint32_t)j * 8192u;
                        issue_kblock(tmem + 0, a, a + kATile, wj, wj + 4096u, umma_idesc(kD), j == 0);
                        tc_commit(bar(A_EMPTY0 + as));
                    }
                    tc_commit(bar(W_EMPTY0 + ws));
                    tc_commit(bar(D3_FULL));
                    ++wit;
                }
            }
        }
    } else {
        // ================= workers: prologue / activation epilogues / output =================
        const int wk = warp - 2;                       // 0..7
        const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
        const int half = wk >> 2;                      // which 16 of every 32 columns
        const int row = quarter * 32 + lane;           // row inside the tile == TMEM lane
        const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
        uint32_t ait = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const uint32_t tpar = t & 1;
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
            const int64_t grow = tile * kTileM + row;
            const bool valid = grow < P.rows;
            const size_t goff = (size_t)grow * kD + half * 16;
            // ---- prologue: y = combine(x, k, coef) for 16 elements of this row; A1 was released by the previous L1
            float xv[16], y[16];
            if (valid) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(P.x + goff) + i);
                    xv[4 * i] = q.x; xv[4 * i + 1] = q.y; xv[4 * i + 2] = q.z; xv[4 * i + 3] = q.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) xv[i] = 0.f;
            }
            if (P.n_k == 0 || !valid) {
#pragma unroll
                for (int i = 0; i < 16; ++i) y[i] = xv[i];
            } else if (P.mode == TDYN_MODE_CHAIN) {
#pragma unroll
                for (int i = 0; i < 16; ++i) y[i] = xv[i];
                for (int j = 0; j < P.n_k; ++j) {
                    const float da = mul_rn(P.dt, P.coef.c[j]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 q = __ldg(reinterpret_cast<const float4*>(P.k.p[j] + goff) + i);
                        y[4 * i] = add_rn(y[4 * i], mul_rn(da, q.x));
                        y[4 * i + 1] = add_rn(y[4 * i + 1], mul_rn(da, q.y));
                        y[4 * i + 2] = add_rn(y[4 * i + 2], mul_rn(da, q.z));
                        y[4 * i + 3] = add_rn(y[4 * i + 3], mul_rn(da, q.w));
                    }
                }
            } else {
                float s[16];
                for (int j = 0; j < P.n_k; ++j) {
                    const float a = P.coef.c[j];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 q = __ldg(reinterpret_cast<const float4*>(P.k.p[j] + goff) + i);
                        const float m0 = mul_rn(a, q.x), m1 = mul_rn(a, q.y), m2 = mul_rn(a, q.z), m3 = mul_rn(a, q.w);
                        s[4 * i] = j == 0 ? m0 : add_rn(s[4 * i], m0);
                        s[4 * i + 1] = j == 0 ? m1 : add_rn(s[4 * i + 1], m1);
                        s[4 * i + 2] = j == 0 ? m2 : add_rn(s[4 * i + 2], m2);
                        s[4 * i + 3] = j == 0 ? m3 : add_rn(s[4 * i + 3], m3);
                    }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) y[i] = add_rn(xv[i], mul_rn(P.dt, s[i]));
            }
            store_split16(smem + kSmemA1, row, half, y);
            fence_proxy_async();
            mbar_arrive(bar(A1_FULL));

            // ---- two activation epilogues: D1 -> layer-2 operand, D2 -> layer-3 operand
#pragma unroll 1
            for (int layer = 0; layer < 2; ++layer) {
                mbar_wait(bar(layer == 0 ? D1_FULL : D2_FULL), tpar);
                tc_fence_after();
                const float* bias = layer == 0 ? P.b1 : P.b2;
                const uint32_t dcol = layer == 0 ? 0u : 256u;
#pragma unroll 1
                for (int j = 0; j < kH / kKB; ++j, ++ait) {
                    const uint32_t as = ait & 1, au = ait >> 1;
                    float v[16];
                    tmem_ld16(lane_addr + dcol + (uint32_t)(j * kKB + half * 16), v);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 bq = __ldg(reinterpret_cast<const float4*>(bias + j * kKB + half * 16) + i);
                        v[4 * i] = activate<ACT>(v[4 * i] + bq.x);
                        v[4 * i + 1] = activate<ACT>(v[4 * i + 1] + bq.y);
                        v[4 * i + 2] = activate<ACT>(v[4 * i + 2] + bq.z);
                        v[4 * i + 3] = activate<ACT>(v[4 * i + 3] + bq.w);
                    }
                    mbar_wait(bar(A_EMPTY0 + as), (au & 1) ^ 1);
                    store_split16(smem + kSmemARing + as * kAStage, row, half, v);
                    fence_proxy_async();
                    mbar_arrive(bar(A_FULL0 + as));
                }
            }
            // ---- output: k_out = D3 + b3 (and x_new for the last stage of a fixed-step method)
            mbar_wait(bar(D3_FULL), tpar);
            tc_fence_after();
            float o[16];
            tmem_ld16(lane_addr + (uint32_t)(half * 16), o);
            tc_fence_before();
            mbar_arrive(bar(D_FREE));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 bq = __ldg(reinterpret_cast<const float4*>(P.b3 + half * 16) + i);
                o[4 * i] += bq.x; o[4 * i + 1] += bq.y; o[4 * i + 2] += bq.z; o[4 * i + 3] += bq.w;
            }
            if (valid) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    reinterpret_cast<float4*>(P.k_out + goff)[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
                if (P.x_new) {
                    // x + dt*(b0*k0 + ... + b_last*k_out)
                    float s[16];
                    for (int j = 0; j < P.n_sol; ++j) {
                        const float b = P.bsol.c[j];
                        const bool last = j == P.n_sol - 1;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float4 q;
                            if (last) q = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
                            else q = __ldg(reinterpret_cast<const float4*>(P.k.p[j] + goff) + i);
                            const float m0 = mul_rn(b, q.x), m1 = mul_rn(b, q.y), m2 = mul_rn(b, q.z), m3 = mul_rn(b, q.w);
                            s[4 * i] = j == 0 ? m0 : add_rn(s[4 * i], m0);
                            s[4 * i + 1] = j == 0 ? m1 : add_rn(s[4 * i + 1], m1);
                            s[4 * i + 2] = j == 0 ? m2 : add_rn(s[4 * i + 2], m2);
                            s[4 * i + 3] = j == 0 ? m3 : add_rn(s[4 * i + 3], m3);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        reinterpret_cast<float4*>(P.x_new + goff)[i] =
                            make_float4(add_rn(xv[4 * i], mul_rn(P.dt, s[4 * i])), add_rn(xv[4 * i + 1], mul_rn(P.dt, s[4 * i + 1])),
                                        add_rn(xv[4 * i + 2], mul_rn(P.dt, s[4 * i + 2])), add_rn(xv[4 * i + 3], mul_rn(P.dt, s[4 * i + 3])));
                }
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
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        float w;
        size_t hi_off, lo_off;
        if (idx < kD * kH) {                       // W1[n][k], n < 256, k < 32
            const int n = idx / kD, k = idx % kD;
            w = w1[idx];
            hi_off = swz(n, k >> 2) + (k & 3) * 4;
            lo_off = hi_off + kWStage / 2;
        } else if (idx < kD * kH + kH * kH) {      // W2[n][k]
            const int e = idx - kD * kH, n = e / kH, k = e % kH, kb = k / kKB, kk = k % kKB;
            w = w2[e];
            hi_off = (size_t)(1 + kb) * kWStage + swz(n, kk >> 2) + (kk & 3) * 4;
            lo_off = hi_off + kWStage / 2;
        } else {                                   // W3[n][k], n < 32, k < 256
            const int e = idx - kD * kH - kH * kH, n = e / kH, k = e % kH, kb = k / kKB, kk = k % kKB;
            w = w3[e];
            hi_off = (size_t)9 * kWStage + (size_t)kb * 8192 + swz(n, kk >> 2) + (kk & 3) * 4;
            lo_off = hi_off + 4096;
        }
        const float hi = tf32_rna(w);
        *reinterpret_cast<float*>(img + hi_off) = hi;
        *reinterpret_cast<float*>(img + lo_off) = tf32_rna(w - hi);
    }
}

}  // namespace tc
}  // namespace tdyn

using namespace tdyn;

extern "C" {

int tdyn_mlp_tc_supported(const tdyn_mlp_t* m) {
    if (!m) return 0;
    return m->n_layers == 3 && m->dims[0] == tc::kD && m->dims[1] == tc::kH && m->dims[2] == tc::kH && m->dims[3] == tc::kD &&
           (m->act == TDYN_ACT_TANH || m->act == TDYN_ACT_SOFTPLUS || m->act == TDYN_ACT_RELU);
}

int64_t tdyn_mlp_tc_prepared_bytes(const tdyn_mlp_t* m) {
    return tdyn_mlp_tc_supported(m) ? (int64_t)tc::kWStagesPerTile * tc::kWStage : 0;
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
    P.k_out = k_out; P.x = x; P.n_k = n_k; P.mode = mode; P.dt = dt;
    const int need_k = x_new ? (n_sol - 1 > n_k ? n_sol - 1 : n_k) : n_k;
    for (int j = 0; j < need_k; ++j) {
        TDYN_REQUIRE(k && k[j] && ((uintptr_t)k[j] & 15) == 0, "tdyn_mlp_tc_stage: k[%d] null or misaligned", j);
        P.k.p[j] = k[j];
    }
    for (int j = 0; j < n_k; ++j) P.coef.c[j] = coef[j];
    P.x_new = x_new; P.n_sol = x_new ? n_sol : 0;
    for (int j = 0; j < P.n_sol; ++j) P.bsol.c[j] = bsol[j];
    P.rows = rows;
    P.num_tiles = (int)((rows + tc::kTileM - 1) / tc::kTileM);
    P.act = m->act;
    const int grid = P.num_tiles < kNumSMs ? P.num_tiles : kNumSMs;
    cudaStream_t st = (cudaStream_t)stream;
    static thread_local bool attr_set = false;
    if (!attr_set) {
        TDYN_CHECK_CUDA(cudaFuncSetAttribute(tc::mlp_tc_stage_kernel<TDYN_ACT_TANH>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemTotal));
        TDYN_CHECK_CUDA(cudaFuncSetAttribute(tc::mlp_tc_stage_kernel<TDYN_ACT_SOFTPLUS>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemTotal));
        TDYN_CHECK_CUDA(cudaFuncSetAttribute(tc::mlp_tc_stage_kernel<TDYN_ACT_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemTotal));
        attr_set = true;
    }
    switch (m->act) {
        case TDYN_ACT_TANH: tc::mlp_tc_stage_kernel<TDYN_ACT_TANH><<<grid, tc::kThreads, tc::kSmemTotal, st>>>(P); break;
        case TDYN_ACT_SOFTPLUS: tc::mlp_tc_stage_kernel<TDYN_ACT_SOFTPLUS><<<grid, tc::kThreads, tc::kSmemTotal, st>>>(P); break;
        default: tc::mlp_tc_stage_kernel<TDYN_ACT_RELU><<<grid, tc::kThreads, tc::kSmemTotal, st>>>(P); break;
    }
    TDYN_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
