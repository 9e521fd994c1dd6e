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
constexpr size_t kHeader = 1024;        // flags[2][TDYN_MAX_PEERS] (u32) | ticket | error

struct Header {
    unsigned int flag[2][TDYN_MAX_PEERS];
    unsigned int ticket;
    unsigned int error;
};

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

}  // namespace peer
}  // namespace tdyn

using namespace tdyn;

extern "C" {

int64_t tdyn_peer_allreduce_buffer_bytes(int64_t slot_bytes) {
    if (slot_bytes <= 0) return 0;
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
    if (is_double) peer::allreduce_kernel<double><<<(int)blocks, peer::kThreads, 0, st>>>(*comm, (double*)vec, n, slot_bytes);
    else peer::allreduce_kernel<float><<<(int)blocks, peer::kThreads, 0, st>>>(*comm, (float*)vec, n, slot_bytes);
    TDYN_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
