// All-reduce(SUM) of a device vector over the GPUs of one box INSIDE one kernel per GPU, through peer-mapped (symmetric)
// memory over NVLink / NVSwitch -- the exchange step of the sharded solves:
//   * the d(mu)/dt block of every stage of the sharded interpolated adjoint (P floats; mu sits inside that adjoint's error
//     norm, reference numerics/sensitivity.py:124,152, so every rank needs the global sum at every stage),
//   * the per-attempt partial sums of squares of the host-driven adaptive loops (a few doubles).
// Every GPU stores its vector into slot [rank] of EVERY peer's exchange buffer (all-gather by direct stores), releases a
// per-rank flag at system scope, waits for the flags of all ranks in its own buffer and adds the slots in rank order:
// deterministic (fixed order), no NCCL call, no extra launch for the reduction.  Two buffer parities alternate, so a call
// may start writing while a slow peer still reads the previous call's slots (a rank releases call s only after it has
// finished reading call s - 1, and nobody starts call s + 1 before having seen every rank's flag of call s).
#include "tdyn_common.cuh"

namespace tdyn {
namespace peer {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 64;          // all blocks must be co-resident: they wait for one another's peers
constexpr size_t kHeader = 1024;        // flags | tickets | error

struct Header {
    unsigned int flag[2][TDYN_MAX_PEERS];       // one-shot: data of rank q has landed; two-shot: rank q's input is published
    unsigned int flag2[2][TDYN_MAX_PEERS];      // two-shot: rank q has written its chunk of the result everywhere
    unsigned int ticket;
    unsigned int ticket2;
    unsigned int error;
};

__device__ __forceinline__ bool wait_flags(const unsigned int* flags, int world, unsigned int seq) {
    for (int q = 0; q < world; ++q) {
        unsigned int seen = 0;
        for (long long spin = 0; spin < (1ll << 26); ++spin) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags + q) : "memory");
            if (seen == seq) break;
            __nanosleep(32);
        }
        if (seen != seq) return false;
    }
    return true;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) allreduce_kernel(tdyn_comm_t comm, T* __restrict__ vec, int64_t n, int64_t cap_bytes) {
    const unsigned int seq = comm.epoch;
    const int par = seq & 1;
    const int world = comm.world, rank = comm.rank;
    __shared__ bool last;
    __shared__ bool ok;
    const size_t slot_bytes = (size_t)cap_bytes;                    // one rank's slot
    auto slot = [&](int owner, int r) {                              // slot of rank r inside owner's buffer, this parity
        return reinterpret_cast<T*>((uint8_t*)comm.peer[owner] + kHeader + ((size_t)par * TDYN_MAX_PEERS + r) * slot_bytes);
    };
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x, stride = (int64_t)gridDim.x * kThreads;
    // ---- 1. all-gather by direct peer stores
    for (int64_t i = tid; i < n; i += stride) {
        const T v = vec[i];
        for (int r = 0; r < world; ++r) slot(r, rank)[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    Header* mine = reinterpret_cast<Header*>(comm.peer[rank]);
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&mine->ticket, 1u);
        last = t == gridDim.x - 1;
    }
    __syncthreads();
    // ---- 2. the last block of this GPU to finish its stores releases this rank's flag on every peer
    if (last && threadIdx.x == 0) {
        mine->ticket = 0u;
        __threadfence_system();
        for (int r = 0; r < world; ++r)
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&reinterpret_cast<Header*>(comm.peer[r])->flag[par][rank]), "r"(seq) : "memory");
    }
    // ---- 3. wait for every rank's flag in the own buffer
    if (threadIdx.x == 0) {
        bool good = true;
        for (int q = 0; q < world && good; ++q) {
            unsigned int seen = 0;
            for (long long spin = 0; spin < (1ll << 26); ++spin) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(&mine->flag[par][q]) : "memory");
                if (seen == seq) break;
                __nanosleep(32);
            }
            good = seen == seq;
        }
        if (!good) mine->error = 1u;       // a peer never arrived: reported by the caller (no hang)
        ok = good;
    }
    __syncthreads();
    if (!ok) {      // poison the vector: the step controller then stops with an error instead of using partial sums
        for (int64_t i = tid; i < n; i += stride) vec[i] = (T)__int_as_float(0x7fc00000);
        return;
    }
    // ---- 4. fixed-order sum of the slots
    for (int64_t i = tid; i < n; i += stride) {
        T s = *(volatile const T*)(slot(rank, 0) + i);
        for (int q = 1; q < world; ++q) s += *(volatile const T*)(slot(rank, q) + i);
        vec[i] = s;
    }
}

__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 ld_vol4(const float4* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// Two-shot variant for larger vectors (reduce-scatter by peer LOADS + all-gather by peer STORES): every rank publishes
// its vector in its own buffer, reduces ONE chunk of the index range over all ranks' published vectors (fixed rank order)
// and stores that chunk of the result into every rank's result area: 2 x (world-1)/world of the payload crosses NVLink per
// GPU instead of (world-1) x.  Buffer: header | in[2 parities][slot] | res[2 parities][slot].
template <typename T>
__global__ void __launch_bounds__(kThreads) allreduce2_kernel(tdyn_comm_t comm, T* __restrict__ vec, int64_t n, int64_t slot_bytes) {
    const unsigned int seq = comm.epoch;
    const int par = seq & 1;
    const int world = comm.world, rank = comm.rank;
    __shared__ bool last, ok;
    auto in_of = [&](int r) { return reinterpret_cast<T*>((uint8_t*)comm.peer[r] + kHeader + (size_t)par * slot_bytes); };
    auto res_of = [&](int r) { return reinterpret_cast<T*>((uint8_t*)comm.peer[r] + kHeader + (size_t)(2 + par) * slot_bytes); };
    Header* mine = reinterpret_cast<Header*>(comm.peer[rank]);
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x, stride = (int64_t)gridDim.x * kThreads;
    // (fp32 vectors whose address is 16-byte aligned move as float4: 16-byte NVLink transactions)
    const bool v4 = sizeof(T) == 4 && (((uintptr_t)vec) & 15) == 0;
    const int64_t n4 = v4 ? n >> 2 : 0;
    // ---- 1. publish the own vector (local copy), then tell every peer
    T* my_in = in_of(rank);
    for (int64_t i = tid; i < n4; i += stride) reinterpret_cast<float4*>(my_in)[i] = reinterpret_cast<const float4*>(vec)[i];
    for (int64_t i = 4 * n4 + tid; i < n; i += stride) my_in[i] = vec[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&mine->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        mine->ticket = 0u;
        __threadfence_system();
        for (int r = 0; r < world; ++r)
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&reinterpret_cast<Header*>(comm.peer[r])->flag[par][rank]), "r"(seq) : "memory");
    }
    if (threadIdx.x == 0) ok = wait_flags(mine->flag[par], world, seq);
    __syncthreads();
    if (ok) {
        // ---- 2. reduce the own chunk over all ranks (peer loads, rank order), scatter the result to every rank
        const int64_t chunk = ((n + world - 1) / world + 3) & ~(int64_t)3;
        const int64_t lo = chunk * rank, hi = lo + chunk < n ? lo + chunk : n;
        const int64_t hi4 = v4 ? (hi > lo ? lo + ((hi - lo) & ~(int64_t)3) : lo) : lo;       // chunk starts are multiples of 4
        for (int64_t i = (lo >> 2) + tid; i < (hi4 >> 2); i += stride) {
            float4 v[TDYN_MAX_PEERS];
#pragma unroll
            for (int q = 0; q < TDYN_MAX_PEERS; ++q)
                if (q < world) v[q] = ld_vol4(reinterpret_cast<const float4*>(in_of(q)) + i);
            float4 sv = v[0];
#pragma unroll
            for (int q = 1; q < TDYN_MAX_PEERS; ++q)
                if (q < world) sv = add4(sv, v[q]);
            for (int r = 0; r < world; ++r) reinterpret_cast<float4*>(res_of(r))[i] = sv;
        }
        for (int64_t i = hi4 + tid; i < hi; i += stride) {
            T v[TDYN_MAX_PEERS];
#pragma unroll
            for (int q = 0; q < TDYN_MAX_PEERS; ++q)          // all peer loads in flight before the first add
                if (q < world) v[q] = *(volatile const T*)(in_of(q) + i);
            T sv = v[0];
#pragma unroll
            for (int q = 1; q < TDYN_MAX_PEERS; ++q)
                if (q < world) sv += v[q];
            for (int r = 0; r < world; ++r) res_of(r)[i] = sv;
        }
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&mine->ticket2, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        mine->ticket2 = 0u;
        __threadfence_system();
        for (int r = 0; r < world; ++r)
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&reinterpret_cast<Header*>(comm.peer[r])->flag2[par][rank]), "r"(seq) : "memory");
    }
    // ---- 3. wait for every rank's chunk, copy the result out
    if (threadIdx.x == 0) ok = ok && wait_flags(mine->flag2[par], world, seq);
    __syncthreads();
    if (!ok) {
        if (threadIdx.x == 0) mine->error = 1u;
        for (int64_t i = tid; i < n; i += stride) vec[i] = (T)__int_as_float(0x7fc00000);
        return;
    }
    const T* my_res = res_of(rank);
    for (int64_t i = tid; i < n4; i += stride) reinterpret_cast<float4*>(vec)[i] = ld_vol4(reinterpret_cast<const float4*>(my_res) + i);
    for (int64_t i = 4 * n4 + tid; i < n; i += stride) vec[i] = *(volatile const T*)(my_res + i);
}

}  // namespace peer
}  // namespace tdyn

using namespace tdyn;

extern "C" {

int64_t tdyn_peer_allreduce_buffer_bytes(int64_t slot_bytes) {
    if (slot_bytes <= 0) return 0;
    // one-shot layout (2 parities x TDYN_MAX_PEERS slots) covers the two-shot layout (2 in + 2 result areas) as well
    return (int64_t)peer::kHeader + 2 * (int64_t)TDYN_MAX_PEERS * ((slot_bytes + 15) / 16 * 16);
}

int tdyn_peer_allreduce(const tdyn_comm_t* comm, void* vec, int64_t n, int is_double, int64_t slot_bytes, void* stream) {
    TDYN_REQUIRE(comm && vec && n > 0, "tdyn_peer_allreduce: bad argument");
    TDYN_REQUIRE(comm->world >= 1 && comm->world <= TDYN_MAX_PEERS && comm->rank >= 0 && comm->rank < comm->world,
                 "tdyn_peer_allreduce: bad world / rank");
    slot_bytes = (slot_bytes + 15) / 16 * 16;
    TDYN_REQUIRE(n * (is_double ? 8 : 4) <= slot_bytes, "tdyn_peer_allreduce: vector larger than the exchange slot");
    for (int r = 0; r < comm->world; ++r) TDYN_REQUIRE(comm->peer[r], "tdyn_peer_allreduce: peer[%d] is null", r);
    int64_t blocks = (n + peer::kThreads - 1) / peer::kThreads;
    if (blocks > peer::kMaxBlocks) blocks = peer::kMaxBlocks;
    cudaStream_t st = (cudaStream_t)stream;
    const bool two_shot = comm->world > 2 && n * (is_double ? 8 : 4) > 32768;      // large payloads: 2(w-1)/w instead of (w-1) x the bytes
    if (two_shot) {
        if (is_double) peer::allreduce2_kernel<double><<<(int)blocks, peer::kThreads, 0, st>>>(*comm, (double*)vec, n, slot_bytes);
        else peer::allreduce2_kernel<float><<<(int)blocks, peer::kThreads, 0, st>>>(*comm, (float*)vec, n, slot_bytes);
    } else if (is_double) peer::allreduce_kernel<double><<<(int)blocks, peer::kThreads, 0, st>>>(*comm, (double*)vec, n, slot_bytes);
    else peer::allreduce_kernel<float><<<(int)blocks, peer::kThreads, 0, st>>>(*comm, (float*)vec, n, slot_bytes);
    TDYN_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
