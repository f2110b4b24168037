// Peer-to-peer exchange over NVLink 5 / NVSwitch for the agent-per-GPU partition: the BEV pyramid all-gather and the
// head-row all-gather as plain stores into the peers' copies of a SYMMETRIC buffer, plus a flag barrier -- no NCCL call on the data
// path.  (The reference has no counterpart: it stacks all agents on one GPU, intermediate_heter_fusion_dataset.py:619,662.)
//
//   heal_p2p_push         every rank copies a slice of ITS chunk to the same offset of every peer's buffer: the slice is read
//                         once from local HBM (16-byte loads) and stored world-1 times through the peer mappings (NVLink egress);
//                         launched on a side stream right after the producing conv, so the transfer of pyramid level l overlaps
//                         the convolutions of level l+1.
//   heal_p2p_signal_wait  bumps a local sequence number, publishes it with system-scope release stores into flags[self] of every
//                         peer, then spins (acquire loads) until all world flags of its own copy reached the sequence number:
//                         "everybody's pushes of this frame have landed in my buffer".  Sequence numbers only grow: no reset, no
//                         ABA, CUDA-graph replays keep counting.
#include "common.cuh"
#include "../../include/heal_b200.h"

namespace {

constexpr int P2P_MAX = 16;
struct PeerPtrs { void* p[P2P_MAX]; };

__global__ void __launch_bounds__(512)
k_p2p_push(const uint4* __restrict__ src, PeerPtrs dst, int world, int self, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = __ldg(src + i);
#pragma unroll
        for (int r = 0; r < P2P_MAX; ++r)
            if (r < world && r != self) reinterpret_cast<uint4*>(dst.p[r])[i] = v;
    }
}

__global__ void k_p2p_signal_wait(PeerPtrs flags, int world, int self, unsigned* __restrict__ seq_dev) {
    __shared__ unsigned s_seq;
    if (threadIdx.x == 0) s_seq = ++seq_dev[0];
    __syncthreads();
    const unsigned seq = s_seq;
    const int r = threadIdx.x;
    if (r < world) {
        __threadfence_system();                                      // this GPU's earlier peer stores are ordered before the flag
        unsigned* f = reinterpret_cast<unsigned*>(flags.p[r]) + self;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(seq) : "memory");
        const unsigned* mine = reinterpret_cast<const unsigned*>(flags.p[self]) + r;
        unsigned v;
        long long t0 = clock64();
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if (clock64() - t0 > 20000000000LL) __trap();            // ~10 s: a peer died; fail loudly instead of hanging the box
        } while ((int)(v - seq) < 0);
    }
    __syncthreads();
    __threadfence_system();
}

}  // namespace

extern "C" int heal_p2p_push(const void* src_local, void* const* peer_dst_host, int world, int self, size_t bytes, void* stream_) {
    if (!src_local || !peer_dst_host || world < 1 || world > P2P_MAX || self < 0 || self >= world) return HEAL_ERR_ARG;
    if (bytes == 0 || world == 1) return HEAL_OK;
    if ((bytes & 15) || ((uintptr_t)src_local & 15)) return HEAL_ERR_UNSUPPORTED;
    PeerPtrs d;
    for (int r = 0; r < P2P_MAX; ++r) d.p[r] = r < world ? peer_dst_host[r] : nullptr;
    for (int r = 0; r < world; ++r) if (r != self && (!d.p[r] || ((uintptr_t)d.p[r] & 15))) return HEAL_ERR_ARG;
    const size_t n16 = bytes / 16;
    size_t blocks = (n16 + 511) / 512;
    if (blocks > (size_t)HEAL_NUM_SMS * 2) blocks = (size_t)HEAL_NUM_SMS * 2;   // copy kernel shares the GPU with the next level's convs
    k_p2p_push<<<(unsigned)blocks, 512, 0, (cudaStream_t)stream_>>>((const uint4*)src_local, d, world, self, n16);
    return heal_check_launch();
}

extern "C" int heal_p2p_signal_wait(void* const* peer_flags_host, int world, int self, unsigned* seq_dev, void* stream_) {
    if (!peer_flags_host || !seq_dev || world < 1 || world > P2P_MAX || self < 0 || self >= world) return HEAL_ERR_ARG;
    PeerPtrs f;
    for (int r = 0; r < P2P_MAX; ++r) f.p[r] = r < world ? peer_flags_host[r] : nullptr;
    k_p2p_signal_wait<<<1, 32, 0, (cudaStream_t)stream_>>>(f, world, self, seq_dev);
    return heal_check_launch();
}
