// K2H: the hub variant of K2 for matrices whose columns are very unevenly used (R-MAT / power-law inputs).
//
// Why.  On an R-MAT tile every nonzero of K2 pulls one panel row out of L2 (hit rate 78 %, 10 TB/s of L2 -> SM traffic
// on C2) and waits several hundred cycles for it, although a few thousand columns receive a third of all nonzeros
// (profiles/r01_hub_analysis.md).  K2H keeps the panel rows of the tile's most frequent columns
// ("hub rows") resident on the SMs for the whole multiply:
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
// 16 bytes global -> shared without passing through registers (LDGSTS, L2 only); completion is per thread, by groups
__device__ __forceinline__ void cb_cp_async16(uint32_t smem_addr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cb_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cb_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 cb_lds16(uint32_t smem_addr) {
    uint4 u;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(smem_addr));
    return u;
}
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

// ------------------------------------------------------------------------------------------------------------------
// K2R: the same persistent, hub-aware kernel with the row gathers PIPELINED THROUGH A SHARED-MEMORY RING.
//
// K2 keeps U gathered vectors per lane in registers and folds them before it issues the next U: the bytes an SM has in
// flight are capped by its register file (64-80 KB at best, less in practice because issue and fold alternate), and an
// L2-hit gather takes several hundred cycles.  Here every lane copies its 16 bytes of the row of nonzero t+D into slot
// (t+D) mod D of its virtual warp's ring with cp.async (LDGSTS: no destination register), and folds nonzero t out of the
// ring.  A lane only ever reads back the bytes it copied itself, so the per-thread completion of cp.async.wait_group is
// all the synchronisation there is - no barrier, no mbarrier.  In flight per SM: threads x D x 16 B = 128 KB at D = 8.
// Entries (column, value, hub rank) of the next step are loaded one step early so the prefetch can run D <= VW nonzeros
// ahead across step boundaries.  Fold order is unchanged: nonzeros of a row in ascending column order.
template <class Op, int VW, int D, bool FULL>
__device__ __forceinline__ void cb_spmm_walk_ring(const SpmmArgs& a, const int64_t chunk, const HubSrc& hub, const uint32_t ring /* my lane's 16 bytes of slot 0 */) {
    static_assert(D >= 2 && D <= VW && (VW % D) == 0, "ring depth must divide the virtual warp width");
    typedef typename Op::T T;
    typedef typename Op::TA TA;
    constexpr int EPL = 16 / sizeof(T);
    constexpr bool HASVAL = Op::akind != A_PATTERN;
    const int lane = threadIdx.x & 31;
    const int vl = lane & (VW - 1);
    const bool live = chunk < a.nchunks;
    const int slab_off = blockIdx.y * a.slab_bytes;
    const int slab_row_bytes = min(a.row_bytes, a.total_row_bytes - slab_off);
    const bool lane_on = FULL || vl * 16 < slab_row_bytes;
    const char* const xbase = a.X + slab_off + vl * 16;
    const uint32_t ldx = (uint32_t)a.ldx_bytes;

    int s = 0, e = 0, ridx = 0;
    bool head_open = false;
    if (live) {
        s = a.chunk_start[chunk];
        e = a.chunk_start[chunk + 1];
        const int cr = a.chunk_row[chunk];
        ridx = cr & 0x7fffffff;
        head_open = cr < 0;
    }
    const int len = e - s;
    const int maxlen = __reduce_max_sync(0xffffffffu, len);       // every lane of the warp runs the same number of steps
    const TA* __restrict__ vals = reinterpret_cast<const TA*>(a.vals);
    Vec16<T> acc;
#pragma unroll
    for (int q = 0; q < EPL; ++q) acc.v[q] = Op::id();
    bool first = true;
    int row = live ? a.nzrows[ridx] : 0;
    char* const carry_head = a.carry + (2 * chunk) * a.carry_stride + slab_off;

    const bool has_hubs = hub.nhub > 0;
    // entries of one step: lane vl holds nonzero base + vl of the chunk
    struct Entry { int cf; TA av; int hs; };
    auto load_entries = [&](int base) {
        Entry en;
        en.cf = 0; en.av = TA(); en.hs = 0xffff;
        if (base + vl < len) {
            en.cf = ld_stream(a.colflag + s + base + vl);
            if (HASVAL) en.av = ld_stream_val<TA>(vals + s + base + vl);
            if (has_hubs) en.hs = (int)__ldcs(hub.hubslot + s + base + vl);         // no hub data without resident hubs
        }
        return en;
    };
    // start the copy of the row of nonzero t (entry j of `en`) into its ring slot; always closes a group
    auto issue = [&](const Entry& en, int j, int t) {
        const uint32_t c = (uint32_t)__shfl_sync(0xffffffffu, en.cf, j, VW) & 0x7fffffffu;
        const int h = has_hubs ? __shfl_sync(0xffffffffu, en.hs, j, VW) : 0xffff;       // has_hubs is uniform over the grid
        if (t < len && lane_on && !(h < hub.nhub)) cb_cp_async16(ring + (uint32_t)(t & (D - 1)) * hub.slot_bytes, xbase + (uint64_t)c * ldx);
        cb_cp_async_commit();
    };

    Entry cur = load_entries(0), nxt = load_entries(VW);
#pragma unroll
    for (int j = 0; j < D; ++j) issue(cur, j, j);
    for (int base = 0; base < maxlen; base += VW) {
        if (base > 0) { cur = nxt; nxt = load_entries(base + VW); }
#pragma unroll
        for (int j = 0; j < VW; ++j) {
            const int t = base + j;
            cb_cp_async_wait<D - 1>();                       // my copies for nonzero t have landed
            const int cf = __shfl_sync(0xffffffffu, cur.cf, j, VW);
            const int h = has_hubs ? __shfl_sync(0xffffffffu, cur.hs, j, VW) : 0xffff;
            const TA av = HASVAL ? (TA)__shfl_sync(0xffffffffu, cur.av, j, VW) : TA();
            if (t < len) {
                if (lane_on) {
                    Vec16<T> x;
                    if (h < hub.nhub) x = ld_hub16<T>(hub, (uint32_t)h, 0);
                    else *reinterpret_cast<uint4*>(&x) = cb_lds16(ring + (uint32_t)(t & (D - 1)) * hub.slot_bytes);
#pragma unroll
                    for (int q = 0; q < EPL; ++q) {
                        const T prod = Op::mul(av, x.v[q]);
                        acc.v[q] = (Op::first_touch && first) ? prod : Op::add(prod, acc.v[q]);
                    }
                }
                first = false;
                if (cf < 0) {                                // last nonzero of its row: write the row out
                    char* dst;
                    bool rmw = false;
                    if (head_open) { dst = carry_head; head_open = false; }
                    else { dst = a.Y + (int64_t)row * a.ldy_bytes + slab_off; rmw = a.accumulate != 0; }
                    if (lane_on) {
                        char* d = dst + vl * 16;
                        if (rmw) {
                            const Vec16<T> y = ld16<T>(d);
#pragma unroll
                            for (int q = 0; q < EPL; ++q) acc.v[q] = Op::add(y.v[q], acc.v[q]);
                        }
                        st16_stream<T>(d, acc);
#pragma unroll
                        for (int q = 0; q < EPL; ++q) acc.v[q] = Op::id();
                    }
                    first = true;
                    ++ridx;
                    if (t + 1 < len) row = a.nzrows[ridx];
                }
            }
            // refill the slot just consumed with the row of nonzero t + D (the fold above has used the old content)
            if (j + D < VW) issue(cur, j + D, t + D);
            else issue(nxt, j + D - VW, t + D);
        }
    }
    cb_cp_async_wait<0>();                                   // the ring is reused by the next chunk
    if (live && !first) {                                    // open row: park the piece for the fix-up
        char* dst = carry_head + (head_open ? 0 : a.carry_stride);
        if (lane_on) st16<T>(dst + vl * 16, acc);
    }
}

template <class Op, int VW, int D, int BT, bool FULL>
__global__ void __launch_bounds__(BT, 1)
cb_spmm_ring_kernel(const SpmmArgs a, const HubArgs h) {
    constexpr int NV = 32 / VW;
    const uint32_t cs = cb_cluster_nctarank(), me = cb_cluster_ctarank();
    uint32_t shift = 0;
    while ((1u << shift) < cs) ++shift;
    char* const smem = cb_dyn_smem();
    // shared memory: [hub slots of this CTA][rings: one of D slots per virtual warp]
    const int slab_off = blockIdx.y * a.slab_bytes;
    const int vecs = min(a.row_bytes, a.total_row_bytes - slab_off) >> 4;
    const int nlocal = (h.nhub - (int)me + (int)cs - 1) / (int)cs;
    const int nslots_max = (h.nhub + (int)cs - 1) / (int)cs;            // the same on every CTA: the rings start at the same offset
    for (int i = threadIdx.x; i < nlocal * vecs; i += blockDim.x) {
        const int j = i / vecs, v = i - j * vecs;
        const int32_t col = h.hubcols[j * (int)cs + (int)me];
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(a.X + (int64_t)col * a.ldx_bytes + slab_off + v * 16));
        *reinterpret_cast<uint4*>(smem + (size_t)j * a.slab_bytes + v * 16) = x;
    }
    cb_cluster_sync();

    const int lane = threadIdx.x & 31;
    HubSrc hub;
    hub.hubslot = h.hubslot;
    hub.nhub = h.nhub;
    hub.smem = cb_smem_u32(smem) + (lane & (VW - 1)) * 16;
    hub.slot_bytes = (uint32_t)a.slab_bytes;
    hub.cta_mask = cs - 1;
    hub.cta_shift = shift;
    const int vw_in_cta = (threadIdx.x >> 5) * NV + lane / VW;
    const uint32_t ring = cb_smem_u32(smem) + (uint32_t)(nslots_max + vw_in_cta * D) * (uint32_t)a.slab_bytes + (uint32_t)(lane & (VW - 1)) * 16;
    unsigned* const counter = h.counter + blockIdx.y;
    for (;;) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned)NV);
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((int64_t)base >= a.nchunks) break;
        cb_spmm_walk_ring<Op, VW, D, FULL>(a, (int64_t)base + lane / VW, hub, ring);
    }
    cb_cluster_sync();
}

}  // namespace cbk
