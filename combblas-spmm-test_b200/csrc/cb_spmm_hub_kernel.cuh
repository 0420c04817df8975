// K2H: the hub variant of K2 for matrices whose columns are very unevenly used (R-MAT / power-law inputs).
//
// Why.  K2 on an R-MAT tile is bounded by L2 -> SM traffic: every nonzero pulls one panel row out of L2 (hit rate
// 78 %, 11.4 TB/s on C2 = the measured throughput cap of the L2 slices), although a few thousand columns receive a
// third of all nonzeros (profiles/r01_hub_analysis.md).  K2H keeps the panel rows of the tile's
// most frequent columns ("hub rows") resident on the SMs for the whole multiply:
//   * one persistent CTA per SM; the CTAs of a thread-block cluster pool their shared memory: hub rank r lives in CTA
//     r mod C of every cluster, slot r div C, so a cluster of C CTAs holds C x (~200 KB / row bytes) hub rows;
//   * a nonzero whose column is a hub reads its row with ld.shared::cluster (own or a neighbour SM's shared memory,
//     traffic that does not cross the L2 slices), every other nonzero gathers from global memory exactly as in K2;
//   * chunks are handed out dynamically (one atomic per warp and chunk group), because a persistent grid cannot rely
//     on the block scheduler for balance.
// The walk over a chunk is K2's own (cb_spmm_walk), so the fold order - and therefore every result bit - is K2's.
//
// Status: validated on the CPU warp/cluster emulator (tests/emul) only; it has not run on hardware yet and is opt-in
// (cb_spmm_hub_config / CB_SPMM_HUB=1).  Replaces nothing in the reference; it is a faster K2.
#pragma once
#include "cb_spmm_kernel.cuh"

namespace cbk {

#ifndef CB_CLUSTER_INTRINSICS     // the CPU emulator supplies host versions
__device__ __forceinline__ uint32_t cb_cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cb_cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// all threads of all CTAs of the cluster; release/acquire so shared-memory writes before it are visible cluster-wide after it
__device__ __forceinline__ void cb_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ char* cb_dyn_smem() {
    extern __shared__ __align__(16) char cb_hub_smem[];
    return cb_hub_smem;
}
__device__ __forceinline__ uint32_t cb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#endif

struct HubArgs {
    const uint16_t* __restrict__ hubslot;   // [nnz] hub rank of each nonzero's column, 0xffff = not a hub
    const int32_t* __restrict__ hubcols;    // [>= nhub] rank -> column of the tile
    int nhub;                               // ranks staged by this launch: nhub <= slots per CTA x cluster size
    unsigned* __restrict__ counter;         // [gridDim.y] next chunk of each column slab; zero before the launch
};

template <class Op, int VW, int R, int U, int BT, bool FULL>
__global__ void __launch_bounds__(BT, 1)
cb_spmm_hub_kernel(const SpmmArgs a, const HubArgs h) {
    constexpr int NV = 32 / VW;
    const uint32_t cs = cb_cluster_nctarank(), me = cb_cluster_ctarank();      // cs is a power of two (launch_hub)
    uint32_t shift = 0;
    while ((1u << shift) < cs) ++shift;
    char* const smem = cb_dyn_smem();

    // stage my share of the hub rows (this column slab of them): local slot j <- rank j*cs + me
    const int slab_off = blockIdx.y * a.slab_bytes;
    const int vecs = min(a.row_bytes, a.total_row_bytes - slab_off) >> 4;
    const int nlocal = (h.nhub - (int)me + (int)cs - 1) / (int)cs;
    for (int i = threadIdx.x; i < nlocal * vecs; i += blockDim.x) {
        const int j = i / vecs, v = i - j * vecs;
        const int32_t col = h.hubcols[j * (int)cs + (int)me];
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(a.X + (int64_t)col * a.ldx_bytes + slab_off + v * 16));
        *reinterpret_cast<uint4*>(smem + (size_t)j * a.slab_bytes + v * 16) = x;
    }
    cb_cluster_sync();

    HubSrc hub;
    hub.hubslot = h.hubslot;
    hub.nhub = h.nhub;
    hub.smem = cb_smem_u32(smem) + ((threadIdx.x & 31) & (VW - 1)) * 16;
    hub.slot_bytes = (uint32_t)a.slab_bytes;
    hub.cta_mask = cs - 1;
    hub.cta_shift = shift;
    unsigned* const counter = h.counter + blockIdx.y;
    for (;;) {
        unsigned base = 0;
        if ((threadIdx.x & 31) == 0) base = atomicAdd(counter, (unsigned)NV);
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((int64_t)base >= a.nchunks) break;                                 // warp-uniform
        cb_spmm_walk<Op, VW, R, U, FULL, true>(a, (int64_t)base + (threadIdx.x & 31) / VW, hub);
    }
    cb_cluster_sync();      // nobody leaves while a neighbour may still be reading its shared memory
}

}  // namespace cbk
