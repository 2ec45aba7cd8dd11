// Peer-memory exchange of the sharded path (one process per GPU, NVLink 5 / NVSwitch P2P mappings).
//
// The only data-path exchange of the KSG path is an all-gather: every rank computes the pointwise values
// (or uploads the columns) of ITS slice and every rank needs the whole vector.  For the 1 ms headline step
// a library collective costs more host time than the kernels it sits between (two NCCL calls per step:
// ~0.2 ms of a 0.74 ms step on 8 GPUs, profiles/r2u_*), so the exchange is done by the producing kernel
// itself: it stores its results straight into every peer's buffer (the peers' device memory is mapped into
// this process: CUDA VMM handles exchanged once by the host layer) and, when its last block is done,
// releases one flag per peer.  The consumer side is a 32-thread kernel that acquires the G flags.
//
//   psi_push_kernel   psi program over this rank's count rows -> values[q_begin + i] on ALL ranks
//                     (+ this rank's unsettled-query counter into every peer's aux slot) -> flags
//   p2p_push_kernel   a contiguous block of bytes (uploaded columns) -> every peer -> flags
//   p2p_wait_kernel   spin until all G flags have reached `epoch`; sum the aux slots
//
// Flags only ever grow (the epoch of the exchange), so nothing is reset and a late reader cannot miss one;
// payload buffers are double-buffered by the caller (parity of the epoch): a rank can only be one exchange
// ahead of a peer, because its NEXT push is stream-ordered after its own wait, which needs the peer's flag.
// Ordering: every block fences its peer stores system-wide before it takes its ticket; the last block
// fences again and stores the flags with release semantics at system scope; the waiter loads them with
// acquire semantics at system scope, and the kernel that reads the payload is a later launch.
#pragma once

#include "common.cuh"
#include "epilogue.cuh"

namespace ksg {

constexpr int P2P_MAX_RANKS = 16;

struct PeerList {
    void* payload[P2P_MAX_RANKS];               // each rank's payload buffer (this exchange's parity)
    unsigned long long* flags[P2P_MAX_RANKS];   // each rank's flag array [world]
    int32_t* aux[P2P_MAX_RANKS];                // each rank's aux array [world] (this parity), or null
    int world, rank;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// the block has issued its peer stores: fence, take a ticket; the LAST block publishes aux + flags
__device__ __forceinline__ void p2p_block_done(const PeerList& peers, unsigned long long epoch, const int32_t* nflag,
                                               int32_t* ticket) {
    __shared__ int s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if ((int)threadIdx.x < peers.world) {
        const int r = threadIdx.x;
        if (peers.aux[r]) {
            peers.aux[r][peers.rank] = nflag ? __ldcg(nflag) : 0;
            __threadfence_system();
        }
        st_release_sys(&peers.flags[r][peers.rank], epoch);
    }
    if (threadIdx.x == 0) *ticket = 0;
}

static __global__ void __launch_bounds__(256)
psi_push_kernel(const int32_t* __restrict__ counts, int64_t nq, const __grid_constant__ PsiProgram prog,
                const double* __restrict__ table, int64_t entries, const __grid_constant__ PeerList peers,
                int64_t q_begin, const int32_t* __restrict__ nflag, unsigned long long epoch,
                int32_t* __restrict__ ticket) {
    // (grid-stride: few fat blocks -- every block ends with a system-wide fence, a round trip over NVLink)
    for (int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; qi < nq; qi += (int64_t)gridDim.x * blockDim.x) {
        const double v = psi_program_value(counts, nq, prog, table, entries, qi);
#pragma unroll 4
        for (int r = 0; r < peers.world; ++r) static_cast<double*>(peers.payload[r])[q_begin + qi] = v;
    }
    p2p_block_done(peers, epoch, nflag, ticket);
}

// src: n8 8-byte words; written at byte offset dst_offset (a multiple of 8) of every peer's payload
static __global__ void __launch_bounds__(256)
p2p_push_kernel(const uint2* __restrict__ src, int64_t n8, const __grid_constant__ PeerList peers, int64_t dst_offset,
                unsigned long long epoch, int32_t* __restrict__ ticket) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const uint2 v = src[i];
#pragma unroll 4
        for (int r = 0; r < peers.world; ++r)
            reinterpret_cast<uint2*>(static_cast<char*>(peers.payload[r]) + dst_offset)[i] = v;
    }
    p2p_block_done(peers, epoch, nullptr, ticket);
}

// flags: this rank's flag array; aux: this rank's aux array of this parity (or null).  *aux_sum = sum of the
// ranks' aux words (every rank computes the same number: all take the same branch afterwards).
static __global__ void p2p_wait_kernel(const unsigned long long* __restrict__ flags, int world, unsigned long long epoch,
                                       const int32_t* __restrict__ aux, int32_t* __restrict__ aux_sum) {
    const int r = threadIdx.x;
    if (r < world) {
        while (ld_acquire_sys(&flags[r]) < epoch) __nanosleep(64);
    }
    __syncwarp();
    int v = 0;
    if (aux && r < world) v = __ldcv(&aux[r]);
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (aux_sum && r == 0) *aux_sum = v;
}

}  // namespace ksg
