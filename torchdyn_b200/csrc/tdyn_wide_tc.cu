// K10/K11: layer-wise tensor-core (tcgen05 / TMEM, split precision: TF32 + bf16 corrections) kernels for WIDE MLP vector fields of any shape whose layer
// widths are multiples of 32 (BASELINE.json configs[2]: 128-1024-128) and for the adjoint of the 32-256-256-32 class.
//
//   K10 tdyn_lin_tc_apply : OUT[r, n] = scale * epi( sum_k IN[r, k] * B[n, k] (+ bias[n]) )
//        forward layer  : B = W            epi = act(.)            (hidden)   or identity (last layer)
//        dgrad          : B = W^T          epi = (.) * act'(a[r,n]) (hidden)  or identity (first layer)
//        = the Linear / activation calls of the vector field and the input half of its VJP
//          (reference core/defunc.py:30-35, numerics/sensitivity.py:62-66, 134-138).
//   K11 tdyn_wgrad_tc     : dW[n, k] = scale * sum_r G[r, n] * IN[r, k],  db[n] = scale * sum_r G[r, n]
//        = the parameter half of the VJP, -lambda^T df/dtheta, contracted over the batch (sensitivity.py:66, 138).
//
// Both use the split a = hi + lo (hi = rna_tf32(a)): hi*hi as kind::tf32, hi*lo + lo*hi as one bf16 chain, fp32
// accumulation in TMEM (see split8_to_tmem).
// K10: M = 128 batch rows per tile, the A operand (activations) is split by producer warps straight from global memory
// into a TMEM ring (both halves: every MMA is the A-from-TMEM form, shared memory only carries B), the pre-split /
// pre-swizzled weight image is streamed from L2 with bulk copies, accumulators are double-buffered [128 x <=128] TMEM
// tiles drained by separate epilogue warps.  K11: M = 128 output features (TMEM lanes), N = <=128 input features,
// K = batch rows; G^T goes to TMEM as it is loaded (lane = feature: no transposition), IN^T is transposed into the
// K-major swizzled layout by the store pattern; the accumulator is flushed into fp32 registers every 256 rows so the
// tensor core's truncating accumulation never runs over more than 96 partial sums; per-CTA partials are reduced in a
// fixed order by a second kernel (deterministic).
#include <cuda_bf16.h>
#include "tdyn_tc_ptx.cuh"

namespace tdyn {
namespace wtc {

using namespace tcx;

constexpr int kKB = 32;                 // K-block: 32 fp32 = 128 B = one swizzle row
constexpr int kNcMax = 128;             // accumulator tile width
constexpr int kThreads = 576;           // warp 0 weight producer, warp 1 MMA, warps 2..17 workers
constexpr int kASlots = 4;              // TMEM ring of A K-blocks: slot = 32 hi + 32 lo columns (K10; K <= 128: the whole
                                        // row tile stays resident and is reused by every output chunk)
constexpr uint32_t kColAcc0 = 0, kColAcc1 = 128, kColA = 256;
constexpr uint32_t kTmemCols = 512;
constexpr int kMaxWidth = 2048;

// ---------------------------------------------------------------------------------------------------- K10
constexpr int kBStages = 4;
constexpr int kSegKB = 8;                                       // accumulators are flushed every 8 K-blocks (K = 256): the tensor
                                                                // core adds with truncation, 96 partial sums keep it at ~3e-6
constexpr uint32_t kBStage = kNcMax * 256;                      // [Nc x 32] hi | lo  (32 KB at Nc = 128)
constexpr uint32_t kLSmemB = 0;
constexpr uint32_t kLSmemPart = kLSmemB + kBStages * kBStage;   // fp32 partial sums [warp][col][lane] when K > 256
constexpr uint32_t kLSmemStage = kLSmemPart + kNcMax * kTileM * 4;  // 8 epilogue warps x [32 rows][16 cols] output staging
constexpr uint32_t kLSmemBias = kLSmemStage + 8 * 2048;             // up to kMaxWidth floats
constexpr uint32_t kLSmemBars = kLSmemBias + kMaxWidth * 4;
constexpr uint32_t kLSmemTotal = kLSmemBars + 256 + 1024;
enum { MODE_LIN = 0, MODE_ACT = 1, MODE_DACT = 2 };
enum LBar { LB_FULL0 = 0, LB_EMPTY0 = LB_FULL0 + kBStages, LA_FULL0 = LB_EMPTY0 + kBStages, LA_EMPTY0 = LA_FULL0 + kASlots,
            LACC_FULL0 = LA_EMPTY0 + kASlots, LACC_EMPTY0 = LACC_FULL0 + 2, L_NUM_BARS = LACC_EMPTY0 + 2 };

struct LinParams {
    const uint8_t* img;        // slabs (chunk c, K-block kb): [Nc x 32] hi then lo, 128B-swizzled
    const float* in;           // [rows][K]
    const float* bias;         // [N] or null
    const float* saved;        // [rows][N] activation outputs (MODE_DACT) or null
    float* out;                // [rows][N]
    int N, K, Nc;
    float scale;
    int64_t rows;
    int num_tiles;
};

// derivative of the activation expressed through its OUTPUT a = act(z)
template <int ACT>
__device__ __forceinline__ float dact_from_out(float a) {
    if (ACT == TDYN_ACT_RELU) return a > 0.f ? 1.f : 0.f;
    if (ACT == TDYN_ACT_SOFTPLUS) return -expm1f(-a);       // sigmoid(z) = 1 - exp(-softplus(z))
    return fmaf(-a, a, 1.f);                                  // 1 - tanh^2
}

// Split precision (same scheme as the fused stage kernel, tdyn_mlp_tc.cu): a = hi + lo, hi = rna_tf32(a).  The main term
// A_hi*W_hi runs as kind::tf32, both correction terms as ONE kind::f16 (bf16) chain over a concatenated K:
// A' = [bf16(a_hi) | bf16(a_lo)], W' = [bf16(w_lo) ; bf16(w_hi)].  A TMEM operand slot (64 columns) of one 32-wide K-block:
// columns [0,32) TF32 hi, [32,48) bf16(hi) packed two per column, [48,64) bf16(lo) packed.
__device__ __forceinline__ uint32_t pack_bf16(float lo_elem, float hi_elem) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo_elem, hi_elem);     // .x (low 16 bits) = element with the lower K index
    return *reinterpret_cast<const uint32_t*>(&p);
}

// 8 consecutive K values of this thread's row: `k8` = their index / 8 inside the K-block (0..3)
__device__ __forceinline__ void split8_to_tmem(uint32_t slot, int k8, const float (&a)[8]) {
    uint32_t hi[8], hb[4], lb[4];
    float h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        h[i] = tf32_rna_fast(a[i]);
        hi[i] = __float_as_uint(h[i]);
        l[i] = a[i] - h[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        hb[i] = pack_bf16(h[2 * i], h[2 * i + 1]);
        lb[i] = pack_bf16(l[2 * i], l[2 * i + 1]);
    }
    tmem_st8(slot + (uint32_t)(k8 * 8), hi);
    tmem_st4(slot + 32u + (uint32_t)(k8 * 4), hb);
    tmem_st4(slot + 48u + (uint32_t)(k8 * 4), lb);
}

// one K-block with the A operand in TMEM: D (+)= A'*W' (4 bf16 MMAs, K = 16) + A_hi*W_hi (4 TF32 MMAs, K = 8)
__device__ __forceinline__ void issue_kblock_tt(uint32_t d_tmem, uint32_t a_slot, uint32_t b_hi, uint32_t b_corr, int n, bool first) {
    const uint32_t id32 = umma_idesc(n), id16 = umma_idesc_bf16(n);
#pragma unroll
    for (int k = 0; k < 4; ++k) tc_mma_bf16_ts(d_tmem, a_slot + 32u + k * 8, umma_desc(b_corr + k * 32), id16, (first && k == 0) ? 0u : 1u);
#pragma unroll
    for (int k = 0; k < 4; ++k) tc_mma_tf32_ts(d_tmem, a_slot + k * 8, umma_desc(b_hi + k * 32), id32, 1u);
}

template <int ACT, int MODE>
__global__ void __launch_bounds__(kThreads, 1) lin_tc_kernel(const LinParams P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t tmem_base_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + kLSmemBars;
    auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < L_NUM_BARS; ++i) {
            uint32_t cnt = 1;
            if (i >= LA_FULL0 && i < LA_FULL0 + kASlots) cnt = 256;
            if (i >= LACC_EMPTY0 && i < LACC_EMPTY0 + 2) cnt = 256;
            mbar_init(bar(i), cnt);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (P.bias) {
        float* sb = reinterpret_cast<float*>(smem + kLSmemBias);
        for (int i = threadIdx.x; i < P.N; i += kThreads) sb[i] = __ldg(P.bias + i);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    const int nchunks = P.N / P.Nc;
    const int nkb = P.K / kKB;
    // Work items = (row tile, output chunk), chunk the fast index; a CTA takes a CONTIGUOUS range [g0, g1) of the global
    // item list, so that the CTAs' loads differ by at most one item (whole tiles per CTA lose up to half of the machine
    // when there are fewer than ~3 tiles per SM: 256 tiles on 148 SMs).  A tile whose chunks are split between two CTAs
    // has its A operand produced by both.
    const long long total_items = (long long)P.num_tiles * nchunks;
    const int g0 = (int)(total_items * blockIdx.x / gridDim.x), g1 = (int)(total_items * (blockIdx.x + 1) / gridDim.x);
    const int items = g1 - g0;
    const int tile_first = g0 / nchunks;
    const int my_tiles = items > 0 ? (g1 - 1) / nchunks - tile_first + 1 : 0;       // distinct row tiles in the range
    const bool resident = nkb <= kASlots;           // A of a row tile fits the TMEM ring: produced once per tile
    const uint32_t aring = resident ? (uint32_t)nkb : (uint32_t)kASlots;
    const uint32_t slab = (uint32_t)P.Nc * 256u;

    if (warp == 0) {
        // ================= weight producer =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int item = 0; item < items; ++item) {
                const int c = (g0 + item) % nchunks;
                const uint8_t* src = P.img + (size_t)c * nkb * slab;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t slot = it % kBStages, use = it / kBStages;
                    mbar_wait(bar(LB_EMPTY0 + slot), (use & 1) ^ 1);
                    mbar_expect_tx(bar(LB_FULL0 + slot), slab);
                    bulk_g2s(sbase + kLSmemB + slot * kBStage, src + (size_t)kb * slab, slab, bar(LB_FULL0 + slot));
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t ait = 0, bit = 0, sc = 0;
            for (int item = 0; item < items; ++item) {
                for (int kb0 = 0; kb0 < nkb; kb0 += kSegKB, ++sc) {           // one accumulator buffer per K segment
                    const uint32_t buf = sc & 1, u = sc >> 1;
                    mbar_wait(bar(LACC_EMPTY0 + buf), (u & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d = tmem + (buf ? kColAcc1 : kColAcc0);
                    const int kb1 = kb0 + kSegKB < nkb ? kb0 + kSegKB : nkb;
                    const int c = (g0 + item) % nchunks;
                    const bool tile_begin = c == 0 || item == 0, tile_end = c == nchunks - 1 || item == items - 1;   // ... of MY share of the tile
                    for (int kb = kb0; kb < kb1; ++kb, ++bit) {
                        const uint32_t as = ait % aring, au = ait / aring, bs = bit % kBStages, bu = bit / kBStages;
                        if (!resident || tile_begin) mbar_wait(bar(LA_FULL0 + as), au & 1);
                        mbar_wait(bar(LB_FULL0 + bs), bu & 1);
                        tc_fence_after();
                        const uint32_t a_hi = tmem + kColA + as * 64u;
                        const uint32_t b_hi = sbase + kLSmemB + bs * kBStage;
                        issue_kblock_tt(d, a_hi, b_hi, b_hi + slab / 2, P.Nc, kb == kb0);
                        if (!resident || tile_end) tc_commit(bar(LA_EMPTY0 + as));
                        tc_commit(bar(LB_EMPTY0 + bs));
                        // resident: the ring position runs over (tile, kb) only and is rewound for every further chunk
                        ++ait;
                    }
                    if (resident && kb1 == nkb && !tile_end) ait -= (uint32_t)nkb;
                    tc_commit(bar(LACC_FULL0 + buf));
                }
            }
        }
    } else if (warp < 10) {
        // ================= A producers: IN rows -> TF32 hi/lo -> TMEM ring =================
        const int quarter = warp & 3;
        const int cg = (warp - 2) >> 2;                           // which 16 of the 32 K values
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16) + kColA;
        const int total = (resident ? my_tiles : items) * nkb;
        auto load = [&](float4 (&r)[4], int n) {
            const int item = n / nkb, kb = n - item * nkb;        // resident: item = tile index inside my range
            const int64_t tile = resident ? tile_first + item : (g0 + item) / nchunks;
            const int64_t grow = tile * kTileM + row;
            if (grow < P.rows) {
                const float4* p = reinterpret_cast<const float4*>(P.in + (size_t)grow * P.K + kb * kKB + cg * 16);
#pragma unroll
                for (int i = 0; i < 4; ++i) r[i] = __ldg(p + i);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto emit = [&](const float4 (&r)[4], int n) {
            const uint32_t as = (uint32_t)n % aring, au = (uint32_t)n / aring;
            mbar_wait(bar(LA_EMPTY0 + as), (au & 1) ^ 1);
            tc_fence_after();
            const uint32_t slot = lane_addr + as * 64u;
            const float a0[8] = {r[0].x, r[0].y, r[0].z, r[0].w, r[1].x, r[1].y, r[1].z, r[1].w};
            const float a1[8] = {r[2].x, r[2].y, r[2].z, r[2].w, r[3].x, r[3].y, r[3].z, r[3].w};
            split8_to_tmem(slot, cg * 2, a0);
            split8_to_tmem(slot, cg * 2 + 1, a1);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(bar(LA_FULL0 + as));
        };
        float4 b0[4], b1[4], b2[4], b3[4];
        if (total > 0) load(b0, 0);
        if (total > 1) load(b1, 1);
        if (total > 2) load(b2, 2);
        int n = 0;
        while (n < total) {                                       // four register buffers in rotation: three loads in flight
            if (n + 3 < total) load(b3, n + 3);
            emit(b0, n); if (++n >= total) break;
            if (n + 3 < total) load(b0, n + 3);
            emit(b1, n); if (++n >= total) break;
            if (n + 3 < total) load(b1, n + 3);
            emit(b2, n); if (++n >= total) break;
            if (n + 3 < total) load(b2, n + 3);
            emit(b3, n); ++n;
        }
    } else {
        // ================= epilogue: TMEM -> (+bias, activation | * act'(saved)) * scale -> OUT =================
        // Each thread owns one row (TMEM lane) and ncol columns of the chunk.  Results go 16 columns at a time through a
        // per-warp shared-memory tile ([32 rows][16 cols], 16-byte chunks XOR-swizzled by row) so that global stores
        // (and the loads of the saved activations, prefetched one pass ahead) are 64-byte row segments, 8 rows per
        // instruction, instead of 16 bytes per row.
        const int quarter = warp & 3;
        const int ew = warp - 10;
        const int half = ew >> 2;
        const int ncol = P.Nc >> 1;                               // columns of the chunk this thread drains: 16, 32 or 64
        const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * ncol);
        const float* sb = reinterpret_cast<const float*>(smem + kLSmemBias);
        float* part = reinterpret_cast<float*>(smem + kLSmemPart) + (size_t)ew * 64 * 32 + lane;       // [col][lane]
        uint8_t* stage = smem + kLSmemStage + (size_t)ew * 2048;
        const int lr = lane >> 2, lch = lane & 3;                 // coalesced phase: row inside a group of 8, chunk
        auto st_off = [](int r, int ch) { return (uint32_t)(r * 64 + ((ch ^ ((r >> 1) & 3)) << 4)); };
        const int nseg = (nkb + kSegKB - 1) / kSegKB;
        const int npass = ncol >> 4;
        uint32_t sc = 0;
        float4 sv[4];                                             // saved activations of the NEXT pass (MODE_DACT)
        auto fetch_saved = [&](int item, int pass) {
            const int64_t tile = (g0 + item) / nchunks;
            const int64_t row0 = tile * kTileM + quarter * 32;
            const int n0 = ((g0 + item) % nchunks) * P.Nc + half * ncol + pass * 16;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t r = row0 + i * 8 + lr;
                sv[i] = r < P.rows ? __ldg(reinterpret_cast<const float4*>(P.saved + (size_t)r * P.N + n0 + lch * 4))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        if (MODE == MODE_DACT && items > 0) fetch_saved(0, 0);
        for (int item = 0; item < items; ++item) {
            const int c = (g0 + item) % nchunks;
            const int64_t tile = (g0 + item) / nchunks;
            const int64_t row0 = tile * kTileM + quarter * 32;    // first row of this warp
            const int n0 = c * P.Nc + half * ncol;
            for (int seg = 0; seg < nseg; ++seg, ++sc) {
                const uint32_t buf = sc & 1, u = sc >> 1;
                const bool first = seg == 0, last = seg == nseg - 1;
                mbar_wait(bar(LACC_FULL0 + buf), u & 1);
                tc_fence_after();
                const uint32_t src = lane_addr + (buf ? kColAcc1 : kColAcc0);
                uint32_t ra[8], rb8[8];
                tmem_ld8(src, ra);
                for (int pass = 0; pass < npass; ++pass) {
                    const int p0 = pass * 16;
                    tmem_ld8(src + (uint32_t)(p0 + 8), rb8);
                    if (MODE == MODE_DACT && last) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(stage + st_off(i * 8 + lr, lch)) = sv[i];
                        __syncwarp();
                        if (pass + 1 < npass) fetch_saved(item, pass + 1);
                        else if (item + 1 < items) fetch_saved(item + 1, 0);
                    }
                    auto proc = [&](const uint32_t (&acc)[8], int g) {      // 8 columns starting at p0 + g (g = 0 or 8)
                        float v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(acc[i]);
                        if (!first) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] += part[(p0 + g + i) * 32];
                        }
                        if (!last) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) part[(p0 + g + i) * 32] = v[i];
                            return;
                        }
                        if (P.bias) {
                            const float4 q0 = *reinterpret_cast<const float4*>(sb + n0 + p0 + g);
                            const float4 q1 = *reinterpret_cast<const float4*>(sb + n0 + p0 + g + 4);
                            v[0] += q0.x; v[1] += q0.y; v[2] += q0.z; v[3] += q0.w;
                            v[4] += q1.x; v[5] += q1.y; v[6] += q1.z; v[7] += q1.w;
                        }
                        float4* st0 = reinterpret_cast<float4*>(stage + st_off(lane, g >> 2));
                        float4* st1 = reinterpret_cast<float4*>(stage + st_off(lane, (g >> 2) + 1));
                        if (MODE == MODE_ACT) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = activate<ACT>(v[i]);
                        } else if (MODE == MODE_DACT) {
                            const float4 s0 = *st0, s1 = *st1;
                            const float a[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] *= dact_from_out<ACT>(a[i]);
                        }
                        if (P.scale != 1.f) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] *= P.scale;
                        }
                        *st0 = make_float4(v[0], v[1], v[2], v[3]);
                        *st1 = make_float4(v[4], v[5], v[6], v[7]);
                    };
                    tmem_wait8(ra);                                   // (also completes rb8: the wait covers all loads)
                    tmem_wait8(rb8);
                    uint32_t ca[8], cb[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { ca[i] = ra[i]; cb[i] = rb8[i]; }
                    if (pass + 1 < npass) tmem_ld8(src + (uint32_t)(p0 + 16), ra);    // next pass in flight
                    proc(ca, 0);
                    proc(cb, 8);
                    if (last) {
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = i * 8 + lr;
                            if (row0 + r < P.rows)
                                *reinterpret_cast<float4*>(P.out + (size_t)(row0 + r) * P.N + n0 + p0 + lch * 4) =
                                    *reinterpret_cast<const float4*>(stage + st_off(r, lch));
                        }
                        __syncwarp();
                    }
                }
                tc_fence_before();
                mbar_arrive(bar(LACC_EMPTY0 + buf));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
    }
}

// image of operand B[n'][k']: transpose = 0: B = W ([n_out x k_in] row-major); transpose = 1: B[n'][k'] = W[k'][n']
__global__ void lin_prepare_kernel(const float* __restrict__ w, int w_rows, int w_cols, int transpose, int Nc,
                                   uint8_t* __restrict__ img) {
    const int Np = transpose ? w_cols : w_rows, Kp = transpose ? w_rows : w_cols;
    const int nkb = Kp / kKB;
    const size_t slab = (size_t)Nc * 256;
    const int total = w_rows * w_cols;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int r = idx / w_cols, cidx = idx - r * w_cols;
        const int n = transpose ? cidx : r, k = transpose ? r : cidx;
        const int c = n / Nc, nl = n - c * Nc, kb = k / kKB, kk = k - kb * kKB;
        const size_t off = ((size_t)c * nkb + kb) * slab + swz((uint32_t)nl, (uint32_t)(kk >> 2)) + (kk & 3) * 4;
        const float v = w[idx];
        const float hi = tf32_rna(v);
        *reinterpret_cast<float*>(img + off) = hi;
        // correction part W' [Nc][64 k'] bf16: k' = kk -> bf16(w_lo) (pairs with bf16(a_hi)), k' = 32 + kk -> bf16(w_hi)
        const size_t corr = ((size_t)c * nkb + kb) * slab + slab / 2;
        *reinterpret_cast<__nv_bfloat16*>(img + corr + swz((uint32_t)nl, (uint32_t)(kk >> 3)) + (kk & 7) * 2) = __float2bfloat16_rn(v - hi);
        *reinterpret_cast<__nv_bfloat16*>(img + corr + swz((uint32_t)nl, (uint32_t)(4 + (kk >> 3))) + (kk & 7) * 2) = __float2bfloat16_rn(hi);
    }
    (void)Np;
}

// ---------------------------------------------------------------------------------------------------- K11
constexpr int kWStages = 3;                         // ring index shared by the raw tiles, the A TMEM slot and the B operand stage
constexpr int kDrainLag = 2;                        // segment s is drained after K-block 8(s+1)+2 has been produced
constexpr uint32_t kRawTile = kKB * kNcMax * 4;     // 16 KB: [32 rows][128 features] fp32 as it lies in global memory
constexpr uint32_t kWSmemOp = 0;                                     // B operand stages (hi | lo), kBStage each
constexpr uint32_t kWSmemRawG = kWSmemOp + kWStages * kBStage;       // raw fp32 data: [stage][16 warps][2 KB] (G rows | IN rows)
constexpr uint32_t kWSmemBias = kWSmemRawG + kWStages * 2 * kRawTile;    // double[4][128]
constexpr uint32_t kWSmemBars = kWSmemBias + 4 * 128 * 8;
constexpr uint32_t kWSmemTotal = kWSmemBars + 256 + 1024;
enum WBar { W_FULL0 = 0, W_EMPTY0 = W_FULL0 + kWStages, WACC_FULL0 = W_EMPTY0 + kWStages, WACC_EMPTY0 = WACC_FULL0 + 2,
            W_NUM_BARS = WACC_EMPTY0 + 2 };

struct WgradParams {
    const float* g;            // [rows][No]
    const float* in;           // [rows][Ki]
    float* part_w;             // [splits][No][Ki]
    double* part_b;            // [splits][No]
    int No, Ki, Nk;            // Nk = min(128, Ki): output-tile width
    int n_tiles, k_tiles, splits;
    int64_t rows, rows_per_split;
};

__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const WgradParams P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t tmem_base_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + kWSmemBars;
    auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < W_NUM_BARS; ++i) {
            uint32_t cnt = 1;
            if (i >= W_FULL0 && i < W_FULL0 + kWStages) cnt = 512;
            if (i >= WACC_EMPTY0 && i < WACC_EMPTY0 + 2) cnt = 512;
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

    const int split = (int)blockIdx.x / (P.n_tiles * P.k_tiles);
    const int tile = (int)blockIdx.x - split * (P.n_tiles * P.k_tiles);
    const int nt = tile / P.k_tiles, kt = tile - nt * P.k_tiles;
    const int n0 = nt * kTileM, k0 = kt * P.Nk;
    const int gw = P.No - n0 < kTileM ? P.No - n0 : kTileM;          // valid output features of this tile
    const int64_t r_begin = (int64_t)split * P.rows_per_split;
    const int64_t r_end = r_begin + P.rows_per_split < P.rows ? r_begin + P.rows_per_split : P.rows;
    const int nkb = r_end > r_begin ? (int)((r_end - r_begin + kKB - 1) / kKB) : 0;
    const int nseg = (nkb + kSegKB - 1) / kSegKB;
    const uint32_t bhalf = (uint32_t)P.Nk * 128u;

    if (warp == 0) {
        // idle: every worker warp fetches its own raw sub-tiles with cp.async (64 bulk copies of 512 B per K-block
        // issued from here were 4x slower than the tensor pipe; a CTA-wide cp.async stage still 2x slower)
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            for (int j = 0; j < nkb; ++j) {
                const int seg = j / kSegKB;
                const uint32_t buf = seg & 1;
                if (j % kSegKB == 0) {
                    mbar_wait(bar(WACC_EMPTY0 + buf), (((uint32_t)seg >> 1) & 1) ^ 1);
                    tc_fence_after();
                }
                const uint32_t s = (uint32_t)j % kWStages, use = (uint32_t)j / kWStages;
                mbar_wait(bar(W_FULL0 + s), use & 1);
                tc_fence_after();
                const uint32_t a_hi = tmem + kColA + s * 64u;
                const uint32_t b_hi = sbase + kWSmemOp + s * kBStage;
                issue_kblock_tt(tmem + (buf ? kColAcc1 : kColAcc0), a_hi, b_hi, b_hi + bhalf, P.Nk, j % kSegKB == 0);
                tc_commit(bar(W_EMPTY0 + s));
                if (j % kSegKB == kSegKB - 1 || j == nkb - 1) tc_commit(bar(WACC_FULL0 + buf));
            }
        }
    } else {
        // ================= transform (raw sub-tiles -> G^T in TMEM, IN^T in shared memory) + accumulator drains =================
        // Every warp fetches exactly the raw data it transforms (cp.async into a warp-private 2 KB slice of each raw
        // stage, two K-blocks ahead, completion by cp.async.wait_group + __syncwarp): no CTA-wide barrier on the raw
        // tiles, the only cross-warp synchronisation left is the operand ring towards the MMA warp.
        const int quarter = warp & 3;
        const int w = warp - 2;                                   // 0..15
        const int rg = w >> 2;                                    // 0..3: which 8 of the 32 rows go to TMEM from this warp
        const int nl = quarter * 32 + lane;                       // output feature inside the tile = TMEM lane
        const bool n_ok = quarter * 32 < gw;                      // warp-uniform (gw is a multiple of 32)
        const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
        const int colb = (w * 32) % P.Nk;                         // B units of this warp: columns colb..colb+31,
        const int ch0 = (w * 32) / P.Nk, ch1 = ch0 + 4;           // rows 4*ch0..+3 (unit 0) and 4*ch1..+3 (unit 1)
        const bool has0 = w * 32 < P.Nk * 8, has1 = P.Nk == kNcMax;
        const int cols_per = P.Nk >> 2;                           // accumulator columns this thread drains (multiple of 8)
        const int nld = cols_per >> 3;
        const bool want_bias = kt == 0;
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        double bsum = 0.0;
        float bs32 = 0.f;                                         // bias partial of the current segment (<= 64 values)

        auto drain = [&](int seg) {
            const uint32_t buf = seg & 1;
            mbar_wait(bar(WACC_FULL0 + buf), ((uint32_t)seg >> 1) & 1);
            tc_fence_after();
            const uint32_t src = lane_base + (buf ? kColAcc1 : kColAcc0) + (uint32_t)(rg * cols_per);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (q < nld) {
                    uint32_t r[8];
                    tmem_ld8(src + (uint32_t)(q * 8), r);
                    tmem_wait8(r);
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[q * 8 + i] += __uint_as_float(r[i]);
                }
            }
            tc_fence_before();
            mbar_arrive(bar(WACC_EMPTY0 + buf));
            bsum += (double)bs32;
            bs32 = 0.f;
        };

        // ---- warp-private raw slices: [stage][warp][A: 8 rows x 32 | B0: 4 rows x 32 | B1: 4 rows x 32] fp32
        const int crow = lane >> 3, cc = lane & 7;                // this lane's 16-byte chunk: row, column chunk
        const float* gsrc = P.g + (size_t)(r_begin + rg * 8 + crow) * P.No + n0 + quarter * 32 + cc * 4;     // rows crow, crow+4
        const float* isrc = P.in + (size_t)(r_begin + ch0 * 4 + crow) * P.Ki + k0 + colb + cc * 4;          // unit 0; unit 1: +16 rows
        const uint32_t raw_w = sbase + kWSmemRawG + (uint32_t)w * 2048u;
        const uint32_t cdst = raw_w + (uint32_t)(crow * 128 + cc * 16);
        const int nfull = (int)((r_end - r_begin) / kKB);         // K-blocks whose 32 rows all exist
        auto issue = [&](int jf, uint32_t stg) {                  // raw data of K-block jf -> stage stg (then one commit group)
            const uint32_t d = cdst + stg * (2u * kRawTile);
            if (jf < nfull) {
                if (n_ok) { cp_async16(d, gsrc); cp_async16(d + 512u, gsrc + (size_t)4 * P.No); }
                if (has0) cp_async16(d + 1024u, isrc);
                if (has1) cp_async16(d + 1536u, isrc + (size_t)16 * P.Ki);
            } else if (jf < nkb) {                                // tail K-block: rows past the end are clamped (masked when read)
                const int nv = (int)(r_end - r_begin) - jf * kKB;
                auto back = [&](int rr) { return (size_t)(rr >= nv ? rr - nv + 1 : 0); };
                if (n_ok) {
                    cp_async16(d, gsrc - back(rg * 8 + crow) * P.No);
                    cp_async16(d + 512u, gsrc + ((size_t)4 - back(rg * 8 + crow + 4)) * P.No);
                }
                if (has0) cp_async16(d + 1024u, isrc - back(ch0 * 4 + crow) * P.Ki);
                if (has1) cp_async16(d + 1536u, isrc + ((size_t)16 - back(ch1 * 4 + crow)) * P.Ki);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            gsrc += (size_t)kKB * P.No;
            isrc += (size_t)kKB * P.Ki;
        };
        const float* rd = reinterpret_cast<const float*>(smem + kWSmemRawG) + w * 512 + lane;       // + stage * 2 * kRawTile / 4
        // B operand of a stage: TF32 tile [Nk][32 r] fp32 | correction tile B' [Nk][64 r'] bf16 with r' = r -> bf16(in_lo)
        // (pairs with bf16(g_hi)), r' = 32 + r -> bf16(in_hi); a unit = 4 consecutive rows r = 4*ch.. of one feature
        uint8_t* b_wr0 = smem + kWSmemOp + swz((uint32_t)(colb + lane), (uint32_t)ch0);
        uint8_t* b_wr1 = smem + kWSmemOp + swz((uint32_t)(colb + lane), (uint32_t)ch1);
        uint8_t* c_wr0 = smem + kWSmemOp + bhalf + swz((uint32_t)(colb + lane), (uint32_t)(ch0 >> 1)) + (ch0 & 1) * 8;
        uint8_t* c_wr1 = smem + kWSmemOp + bhalf + swz((uint32_t)(colb + lane), (uint32_t)(ch1 >> 1)) + (ch1 & 1) * 8;
        const uint32_t c_hi0 = swz((uint32_t)(colb + lane), (uint32_t)(4 + (ch0 >> 1))) - swz((uint32_t)(colb + lane), (uint32_t)(ch0 >> 1));
        const uint32_t c_hi1 = swz((uint32_t)(colb + lane), (uint32_t)(4 + (ch1 >> 1))) - swz((uint32_t)(colb + lane), (uint32_t)(ch1 >> 1));
        const uint32_t t_a = lane_base + kColA;
        const uint32_t bar0 = bar(0);
        auto split4 = [&](const float (&v)[4], uint8_t* dst, uint8_t* cdst, uint32_t c_hi) {
            float4 hi;
            float lo[4];
            hi.x = tf32_rna_fast(v[0]); lo[0] = v[0] - hi.x; hi.y = tf32_rna_fast(v[1]); lo[1] = v[1] - hi.y;
            hi.z = tf32_rna_fast(v[2]); lo[2] = v[2] - hi.z; hi.w = tf32_rna_fast(v[3]); lo[3] = v[3] - hi.w;
            *reinterpret_cast<float4*>(dst) = hi;
            *reinterpret_cast<uint2*>(cdst) = make_uint2(pack_bf16(lo[0], lo[1]), pack_bf16(lo[2], lo[3]));
            *reinterpret_cast<uint2*>(cdst + c_hi) = make_uint2(pack_bf16(hi.x, hi.y), pack_bf16(hi.z, hi.w));
        };

        issue(0, 0);
        issue(1, 1);
        uint32_t st = 0, ph = 0, st2 = 2;                         // ring stage / phase parity of K-block j; stage of K-block j + 2
        int seg_pos = 0;                                          // j % kSegKB
        for (int j = 0; j < nkb; ++j) {
            issue(j + 2, st2);                                    // (stage st2 was read in the previous iteration)
            asm volatile("cp.async.wait_group 2;" ::: "memory");
            __syncwarp();
            const float* r = rd + st * (2 * kRawTile / 4);
            float a[8], b0[4], b1[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = n_ok ? r[i * 32] : 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                b0[i] = has0 ? r[256 + i * 32] : 0.f;
                b1[i] = has1 ? r[384 + i * 32] : 0.f;
            }
            if (j >= nfull) {                                     // tail K-block: mask the rows that do not exist
                const int nvalid = (int)(r_end - r_begin) - j * kKB;
#pragma unroll
                for (int i = 0; i < 8; ++i) if (rg * 8 + i >= nvalid) a[i] = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (ch0 * 4 + i >= nvalid) b0[i] = 0.f;
                    if (ch1 * 4 + i >= nvalid) b1[i] = 0.f;
                }
            }
            __syncwarp();                                         // all lanes have read this stage before it is refilled
            mbar_wait(bar0 + 8u * (W_EMPTY0 + st), ph ^ 1);
            tc_fence_after();
            split8_to_tmem(t_a + st * 64u, rg, a);                // rows rg*8.. of the K-block = K values rg*8..
            if (want_bias) bs32 += ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
            if (has0) split4(b0, b_wr0 + st * kBStage, c_wr0 + st * kBStage, c_hi0);
            if (has1) split4(b1, b_wr1 + st * kBStage, c_wr1 + st * kBStage, c_hi1);
            tmem_wait_st();
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(bar0 + 8u * (W_FULL0 + st));
            // flush the previous segment's accumulator once its MMAs have had time to retire
            if (seg_pos == kDrainLag && j >= kSegKB) drain(j / kSegKB - 1);
            if (++seg_pos == kSegKB) seg_pos = 0;
            st2 = st;
            if (++st == kWStages) { st = 0; ph ^= 1; }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        // segments whose drain point lies beyond the last K-block
        for (int seg = 0; seg < nseg; ++seg)
            if ((seg + 1) * kSegKB + kDrainLag >= nkb) drain(seg);
        bsum += (double)bs32;

        // ---- partial results: part_w[split][n][k0 + rg*cols_per ...], part_b[split][n] (k-tile 0 only)
        if (n_ok) {
            float* dst = P.part_w + ((size_t)split * P.No + n0 + nl) * P.Ki + k0 + rg * cols_per;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (q < nld) {
                    reinterpret_cast<float4*>(dst + q * 8)[0] = make_float4(acc[q * 8], acc[q * 8 + 1], acc[q * 8 + 2], acc[q * 8 + 3]);
                    reinterpret_cast<float4*>(dst + q * 8)[1] = make_float4(acc[q * 8 + 4], acc[q * 8 + 5], acc[q * 8 + 6], acc[q * 8 + 7]);
                }
            }
        }
        double* sbias = reinterpret_cast<double*>(smem + kWSmemBias);
        sbias[rg * 128 + nl] = bsum;
        asm volatile("bar.sync 1, 512;" ::: "memory");             // the 16 worker warps only
        if (kt == 0 && rg == 0 && n_ok)
            P.part_b[(size_t)split * P.No + n0 + nl] = ((sbias[nl] + sbias[128 + nl]) + sbias[256 + nl]) + sbias[384 + nl];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
    }
}

// dW[i] = scale * sum_s part_w[s][i] (fp64, fixed order), db[n] likewise
__global__ void wgrad_reduce_kernel(const float* __restrict__ part_w, const double* __restrict__ part_b, float* __restrict__ dw,
                                    float* __restrict__ db, int n_w, int n_b, int splits, float scale) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_w + n_b; i += gridDim.x * blockDim.x) {
        double s = 0.0;
        if (i < n_w) {
            for (int k = 0; k < splits; ++k) s += (double)part_w[(size_t)k * n_w + i];
            dw[i] = scale * (float)s;
        } else {
            const int n = i - n_w;
            for (int k = 0; k < splits; ++k) s += part_b[(size_t)k * n_b + n];
            db[n] = scale * (float)s;
        }
    }
}

static bool width_ok(int n) { return n <= kMaxWidth && (n == 32 || n == 64 || (n >= kNcMax && n % kNcMax == 0)); }
static int chunk_of(int n) { return n < kNcMax ? n : kNcMax; }

struct WgradPlan { int n_tiles, k_tiles, splits, Nk; int64_t rows_per_split; size_t part_w_bytes, part_b_bytes; };
static WgradPlan wgrad_plan(int No, int Ki, int64_t rows) {
    WgradPlan p;
    p.Nk = chunk_of(Ki);
    p.n_tiles = (No + kTileM - 1) / kTileM;
    p.k_tiles = Ki / p.Nk;
    const int tiles = p.n_tiles * p.k_tiles;
    int splits = kNumSMs / tiles;
    if (splits < 1) splits = 1;
    const int64_t seg_rows = (int64_t)kSegKB * kKB;
    const int64_t max_splits = (rows + seg_rows - 1) / seg_rows;
    if (splits > max_splits) splits = (int)(max_splits > 0 ? max_splits : 1);
    int64_t rps = (rows + splits - 1) / splits;
    rps = (rps + kKB - 1) / kKB * kKB;
    p.rows_per_split = rps;
    p.splits = (int)((rows + rps - 1) / rps);
    if (p.splits < 1) p.splits = 1;
    p.part_w_bytes = (size_t)p.splits * No * Ki * sizeof(float);
    p.part_b_bytes = (size_t)p.splits * No * sizeof(double);
    return p;
}

}  // namespace wtc
}  // namespace tdyn

using namespace tdyn;

template <int ACT, int MODE>
static int launch_lin(const wtc::LinParams& P, int grid, cudaStream_t st) {
    static thread_local unsigned long long seen = 0;
    if (first_use_on_device(seen)) {
        TDYN_CHECK_CUDA(cudaFuncSetAttribute(wtc::lin_tc_kernel<ACT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, wtc::kLSmemTotal));
    }
    wtc::lin_tc_kernel<ACT, MODE><<<grid, wtc::kThreads, wtc::kLSmemTotal, st>>>(P);
    TDYN_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" {

int tdyn_lin_tc_supported(int n_out, int k_in) {
    return wtc::width_ok(n_out) && k_in >= 32 && k_in <= wtc::kMaxWidth && k_in % 32 == 0;
}

int64_t tdyn_lin_tc_image_bytes(int n_out, int k_in) {
    return tdyn_lin_tc_supported(n_out, k_in) ? (int64_t)n_out * k_in * 8 : 0;
}

int tdyn_lin_tc_prepare(const float* w, int w_rows, int w_cols, int transpose, void* img, void* stream) {
    const int n = transpose ? w_cols : w_rows, k = transpose ? w_rows : w_cols;
    TDYN_REQUIRE(w && img, "tdyn_lin_tc_prepare: null argument");
    TDYN_REQUIRE(tdyn_lin_tc_supported(n, k), "tdyn_lin_tc_prepare: unsupported layer shape %d x %d", n, k);
    TDYN_REQUIRE(((uintptr_t)img & 15) == 0, "tdyn_lin_tc_prepare: image must be 16-byte aligned");
    wtc::lin_prepare_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(w, w_rows, w_cols, transpose, wtc::chunk_of(n), (uint8_t*)img);
    TDYN_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int tdyn_lin_tc_apply(const void* img, int n_out, int k_in, const float* in, const float* bias, const float* saved,
                      float* out, int act, int mode, float scale, int64_t rows, void* stream) {
    TDYN_REQUIRE(tdyn_lin_tc_supported(n_out, k_in), "tdyn_lin_tc_apply: unsupported layer shape %d x %d", n_out, k_in);
    TDYN_REQUIRE(img && in && out, "tdyn_lin_tc_apply: null argument");
    TDYN_REQUIRE(mode >= 0 && mode <= 2, "tdyn_lin_tc_apply: mode must be 0 (linear), 1 (activation) or 2 (times act')");
    TDYN_REQUIRE(mode == wtc::MODE_LIN || act == TDYN_ACT_TANH || act == TDYN_ACT_SOFTPLUS || act == TDYN_ACT_RELU,
                 "tdyn_lin_tc_apply: unknown activation");
    TDYN_REQUIRE(mode != wtc::MODE_DACT || saved, "tdyn_lin_tc_apply: mode 2 needs the saved activations");
    TDYN_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)img & 15) == 0 &&
                 (!saved || ((uintptr_t)saved & 15) == 0), "tdyn_lin_tc_apply: pointers must be 16-byte aligned");
    if (rows <= 0) return 0;
    wtc::LinParams P{};
    P.img = (const uint8_t*)img; P.in = in; P.bias = bias; P.saved = saved; P.out = out;
    P.N = n_out; P.K = k_in; P.Nc = wtc::chunk_of(n_out); P.scale = scale; P.rows = rows;
    P.num_tiles = (int)((rows + tcx::kTileM - 1) / tcx::kTileM);
    const long long total_items = (long long)P.num_tiles * (n_out / P.Nc);       // (row tile, output chunk) items, spread evenly
    const int grid = total_items < kNumSMs ? (int)total_items : kNumSMs;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == wtc::MODE_LIN) return launch_lin<TDYN_ACT_TANH, wtc::MODE_LIN>(P, grid, st);
    if (mode == wtc::MODE_ACT) {
        if (act == TDYN_ACT_TANH) return launch_lin<TDYN_ACT_TANH, wtc::MODE_ACT>(P, grid, st);
        if (act == TDYN_ACT_SOFTPLUS) return launch_lin<TDYN_ACT_SOFTPLUS, wtc::MODE_ACT>(P, grid, st);
        return launch_lin<TDYN_ACT_RELU, wtc::MODE_ACT>(P, grid, st);
    }
    if (act == TDYN_ACT_TANH) return launch_lin<TDYN_ACT_TANH, wtc::MODE_DACT>(P, grid, st);
    if (act == TDYN_ACT_SOFTPLUS) return launch_lin<TDYN_ACT_SOFTPLUS, wtc::MODE_DACT>(P, grid, st);
    return launch_lin<TDYN_ACT_RELU, wtc::MODE_DACT>(P, grid, st);
}

int64_t tdyn_wgrad_tc_workspace_bytes(int n_out, int k_in, int64_t rows) {
    if (!(n_out >= 32 && n_out % 32 == 0 && n_out <= wtc::kMaxWidth && wtc::width_ok(k_in)) || rows <= 0) return 0;
    const wtc::WgradPlan p = wtc::wgrad_plan(n_out, k_in, rows);
    return (int64_t)(p.part_w_bytes + p.part_b_bytes + 256);
}

int tdyn_wgrad_tc(const float* g, int n_out, const float* in, int k_in, float* dw, float* db, float scale, void* ws,
                  int64_t ws_bytes, int64_t rows, void* stream) {
    TDYN_REQUIRE(n_out >= 32 && n_out % 32 == 0 && n_out <= wtc::kMaxWidth && wtc::width_ok(k_in),
                 "tdyn_wgrad_tc: unsupported layer shape %d x %d", n_out, k_in);
    TDYN_REQUIRE(g && in && dw && db && ws, "tdyn_wgrad_tc: null argument");
    TDYN_REQUIRE(rows > 0, "tdyn_wgrad_tc: rows must be positive");
    TDYN_REQUIRE(((uintptr_t)ws & 255) == 0, "tdyn_wgrad_tc: workspace must be 256-byte aligned");
    TDYN_REQUIRE(ws_bytes >= tdyn_wgrad_tc_workspace_bytes(n_out, k_in, rows), "tdyn_wgrad_tc: workspace too small");
    const wtc::WgradPlan p = wtc::wgrad_plan(n_out, k_in, rows);
    wtc::WgradParams P{};
    P.g = g; P.in = in;
    P.part_w = (float*)ws;
    P.part_b = (double*)((uint8_t*)ws + ((p.part_w_bytes + 255) & ~(size_t)255));
    P.No = n_out; P.Ki = k_in; P.Nk = p.Nk; P.n_tiles = p.n_tiles; P.k_tiles = p.k_tiles; P.splits = p.splits;
    P.rows = rows; P.rows_per_split = p.rows_per_split;
    cudaStream_t st = (cudaStream_t)stream;
    static thread_local unsigned long long seen = 0;
    if (first_use_on_device(seen)) {
        TDYN_CHECK_CUDA(cudaFuncSetAttribute(wtc::wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wtc::kWSmemTotal));
    }
    wtc::wgrad_tc_kernel<<<p.n_tiles * p.k_tiles * p.splits, wtc::kThreads, wtc::kWSmemTotal, st>>>(P);
    TDYN_CHECK_CUDA(cudaGetLastError());
    const int n_w = n_out * k_in;
    wtc::wgrad_reduce_kernel<<<(n_w + n_out + 255) / 256, 256, 0, st>>>(P.part_w, P.part_b, dw, db, n_w, n_out, p.splits, scale);
    TDYN_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
