// Host-side launch of K2 (+ fix-up) for one semiring functor: picks the virtual-warp layout from the row width.
#pragma once
#include "cb_hub.cuh"
#include "cb_spmm_hub_kernel.cuh"
#include "cb_spmm_tma_kernel.cuh"

#ifndef CB_WIDE_U
#define CB_WIDE_U 4
#define CB_WIDE_B 5
#endif
#ifndef CB_WIDE_B64
#define CB_WIDE_B64 4
#endif
#ifndef CB_DEEP_B64
#define CB_DEEP_B64 3
#endif
#ifndef CB_DEEP_U
#define CB_DEEP_U 8
#define CB_DEEP_B 4
#endif
// hub variant: one persistent CTA of CB_HUB_BT threads per SM, CB_HUB_U row gathers in flight per lane.  1024 threads cap
// the kernel at 64 registers, which U=4 fits without spills for every element type (U=8 spills 16-200 bytes on 4-byte types)
#ifndef CB_PIPE_DEFAULT
#define CB_PIPE_DEFAULT 0            // ring depth of K2P used when nothing else is asked for (0 = the round-1 walk)
#endif
#ifndef CB_HUB_U
#define CB_HUB_U 4
#endif

namespace cbk {

struct LaunchParams {
    cb_ctx* ctx;
    cudaStream_t stream;
    const cb_tile* t;
    const void* X;
    int64_t ldx_bytes;
    void* Y;
    int64_t ldy_bytes;
    int total_row_bytes;      // padded to 16
    int accumulate;
    const HubPlan* hub = nullptr;   // non-null: run the hub variant (K2H) with this shape
    int slab_bytes = 0;             // > 0: column slabs of this width for plain K2 (0 = one slab as wide as the layout allows)
    int point = -1;                 // operating point of plain K2: 0 deep, 1 wide, -1 choose by footprint
    const uint8_t* hubcls = nullptr; // non-null: K2P gathers rows of columns of class <= cls_max with evict_last, others evict_first
    int cls_max = -1;
    int pipe = -1;                  // register-ring depth of the pipelined walk K2P: 4 or 8; 0 = the round-1 walk; -1 = default
    unsigned* persist_counter = nullptr;    // non-null: K2 with persistent warps (chunk counter, zeroed on the stream before the launch)
    const int32_t* win_colflag = nullptr;   // non-null: K2W - this column stream (hub columns as bit 30 + rank) and the packed hub panel at X + win_delta
    int64_t win_delta = 0;
};

enum { CB_HUB_FALLBACK = -77 };     // internal: the hub launch is not possible here, run plain K2

template <class Op, int VW, int R, int U, int MINB, bool FULL, int PIPE = 0, bool POL = false>      // PIPE: 0 K2, 1 K2P (ring), 2 K2 with prefetch, 3 K2W (hub panel), 4 K2 with persistent warps
static int launch_layout_f(const LaunchParams& p) {
    const cb_tile* t = p.t;
    SpmmArgs a;
    a.colflag = t->colflag;
    a.vals = t->vals;
    a.nzrows = t->nzrows;
    a.chunk_start = t->chunk_start;
    a.chunk_row = t->chunk_row;
    a.nchunks = t->nchunks;
    a.X = (const char*)p.X;
    a.Y = (char*)p.Y;
    a.ldx_bytes = p.ldx_bytes;
    a.ldy_bytes = p.ldy_bytes;
    a.slab_bytes = VW * R * 16;
    a.row_bytes = a.slab_bytes;
    a.total_row_bytes = p.total_row_bytes;
    a.carry = (char*)t->carry;
    a.carry_stride = p.total_row_bytes;
    a.accumulate = p.accumulate;
    a.hubcls = p.hubcls;
    a.cls_max = p.cls_max;
    a.win_delta = 0;
    if constexpr (PIPE == 3) { a.colflag = p.win_colflag; a.win_delta = p.win_delta; }
    constexpr int NV = 32 / VW;
    const int64_t vws_per_block = 8 * NV;
    dim3 grid((unsigned)((t->nchunks + vws_per_block - 1) / vws_per_block), (unsigned)((p.total_row_bytes + a.slab_bytes - 1) / a.slab_bytes));
    {
        cb_prof_scope prof(p.ctx, p.stream, CB_PROF_SPMM);
        if constexpr (PIPE == 1) cb_spmm_pipe_kernel<Op, VW, R, U, MINB, FULL, POL><<<grid, 256, 0, p.stream>>>(a);      // U = ring depth
        else if constexpr (PIPE == 2) cb_spmm_kernel<Op, VW, R, U, MINB, FULL, true><<<grid, 256, 0, p.stream>>>(a);
        else if constexpr (PIPE == 3) cb_spmm_kernel<Op, VW, R, U, MINB, FULL, false, true><<<grid, 256, 0, p.stream>>>(a);
        else if constexpr (PIPE == 4) {
            // one launch fills the chip: MINB CTAs per SM (the kernel's own occupancy bound), never more CTAs than there are chunk groups
            const unsigned full_chip = (unsigned)p.ctx->sm_count * (unsigned)MINB;
            cb_spmm_persist_kernel<Op, VW, R, U, MINB, FULL><<<dim3(grid.x < full_chip ? grid.x : full_chip, 1), 256, 0, p.stream>>>(a, p.persist_counter);
        }
        else cb_spmm_kernel<Op, VW, R, U, MINB, FULL><<<grid, 256, 0, p.stream>>>(a);
    }
    CB_LAUNCHED(p.ctx);
    CB_CUDA(p.ctx, cudaGetLastError());
    return CB_OK;
}

// rows that fill the layout exactly (k*sizeof(T) == VW*R*16, the power-of-two panels) take the predicate-free kernel
template <class Op, int VW, int R, int U, int MINB>
static int launch_layout(const LaunchParams& p) {
    if (p.total_row_bytes % (VW * R * 16) == 0) return launch_layout_f<Op, VW, R, U, MINB, true>(p);
    return launch_layout_f<Op, VW, R, U, MINB, false>(p);
}
// K2P: D row gathers in flight per lane in a register ring
template <class Op, int VW, int R, int D, int MINB>
static int launch_pipe(const LaunchParams& p) {
    const bool full = p.total_row_bytes % (VW * R * 16) == 0;
#ifdef CB_BUILD_L2HINT                 // measured in round 2 (profiles/r02_sweep_b_k2p_l2hints.jsonl): -9 % DRAM traffic, +24 % instructions, slower
    if constexpr (D * R == 8) {        // the L2-hint variant exists for the deep ring (the point used when the gathers go to DRAM)
        if (p.hubcls && p.t->n < (1LL << 30))
            return full ? launch_layout_f<Op, VW, R, D, MINB, true, 1, true>(p) : launch_layout_f<Op, VW, R, D, MINB, false, 1, true>(p);
    }
#endif
    return full ? launch_layout_f<Op, VW, R, D, MINB, true, 1>(p) : launch_layout_f<Op, VW, R, D, MINB, false, 1>(p);
}
// K2 with the entry prefetch
template <class Op, int VW, int R, int U, int MINB>
static int launch_pf(const LaunchParams& p) {
    if (p.total_row_bytes % (VW * R * 16) == 0) return launch_layout_f<Op, VW, R, U, MINB, true, 2>(p);
    return launch_layout_f<Op, VW, R, U, MINB, false, 2>(p);
}

// K2W: K2 reading the rows of the most used columns from a packed panel under a persisting L2 window (fp32 / fp64 PlusTimes only)
template <class Op, int VW, int R, int U, int MINB>
static int launch_win(const LaunchParams& p) {
    typedef typename Op::T T;
    constexpr bool built = Op::akind == A_SAME && (std::is_same<T, float>::value || std::is_same<T, double>::value);
    if constexpr (!built) return CB_HUB_FALLBACK;
    else {
        if (p.total_row_bytes % (VW * R * 16) != 0) return CB_HUB_FALLBACK;
        return launch_layout_f<Op, VW, R, U, MINB, true, 3>(p);
    }
}

// K2 with persistent warps (single column slab only; the operand types of the benchmark workloads)
template <class Op, int VW, int R, int U, int MINB>
static int launch_persist(const LaunchParams& p) {
    typedef typename Op::T T;
    constexpr bool built = Op::akind == A_SAME && (std::is_same<T, float>::value || std::is_same<T, double>::value || std::is_same<T, int32_t>::value);
    if constexpr (!built) return CB_HUB_FALLBACK;
    else {
        if (p.total_row_bytes > VW * R * 16) return CB_HUB_FALLBACK;          // more than one column slab: plain K2
        if (p.total_row_bytes % (VW * R * 16) == 0) return launch_layout_f<Op, VW, R, U, MINB, true, 4>(p);
        return launch_layout_f<Op, VW, R, U, MINB, false, 4>(p);
    }
}

// K2T: row gathers as bulk asynchronous copies into a per-warp shared-memory ring (cb_spmm_tma_kernel.cuh).  Only for the operand
// types the benchmark workloads use (every instantiation costs compile time); anything else reports CB_HUB_FALLBACK -> K2.
template <class Op, int VW>
static int launch_tma(const LaunchParams& p) {
    typedef typename Op::T T;
    constexpr bool built = Op::akind == A_SAME && (std::is_same<T, float>::value || std::is_same<T, int32_t>::value);
    if constexpr (!built) return CB_HUB_FALLBACK;
    else {
        if (p.total_row_bytes != VW * 16) return CB_HUB_FALLBACK;
        const int stages_env = getenv("CB_TMA_STAGES") ? atoi(getenv("CB_TMA_STAGES")) : 0;
        const int warps_env = getenv("CB_TMA_WARPS") ? atoi(getenv("CB_TMA_WARPS")) : 0;
        const cb_tile* t = p.t;
        SpmmArgs a;
        a.colflag = t->colflag;
        a.vals = t->vals;
        a.nzrows = t->nzrows;
        a.chunk_start = t->chunk_start;
        a.chunk_row = t->chunk_row;
        a.nchunks = t->nchunks;
        a.X = (const char*)p.X;
        a.Y = (char*)p.Y;
        a.ldx_bytes = p.ldx_bytes;
        a.ldy_bytes = p.ldy_bytes;
        a.slab_bytes = VW * 16;
        a.row_bytes = a.slab_bytes;
        a.total_row_bytes = p.total_row_bytes;
        a.carry = (char*)t->carry;
        a.carry_stride = p.total_row_bytes;
        a.accumulate = p.accumulate;
        a.hubcls = nullptr;
        a.cls_max = -1;
        const int stages = stages_env == 2 || stages_env == 4 ? stages_env : (VW == 32 ? 2 : 4);
        const size_t per_warp = (size_t)stages * 32 * VW * 16 + (size_t)stages * 8;
        int warps = warps_env > 0 ? warps_env : (int)std::min<size_t>(12, (200 * 1024) / per_warp);
        warps = std::max(1, std::min(warps, 12));
        const size_t smem = per_warp * (size_t)warps;
        auto kernel = stages == 2 ? cb_spmm_tma_kernel<Op, VW, 2, 8> : cb_spmm_tma_kernel<Op, VW, 4, 8>;
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return CB_HUB_FALLBACK; }
        const int64_t per_block = (int64_t)warps * (32 / VW);
        dim3 grid((unsigned)((t->nchunks + per_block - 1) / per_block));
        {
            cb_prof_scope prof(p.ctx, p.stream, CB_PROF_SPMM);
            kernel<<<grid, warps * 32, smem, p.stream>>>(a);
        }
        CB_LAUNCHED(p.ctx);
        CB_CUDA(p.ctx, cudaGetLastError());
        return CB_OK;
    }
}

// K2H / K2R: persistent CTAs (one per SM) in clusters that pool their shared memory for the hub rows; dynamic chunks
template <class Op, int VW, bool FULL, bool RING>
static int launch_hub_layout(const LaunchParams& p) {
    constexpr int BT = CB_HUB_BT;                                      // one CTA owns the SM
    constexpr int U = CB_HUB_U;
    const cb_tile* t = p.t;
    const HubPlan& hp = *p.hub;
    SpmmArgs a;
    a.colflag = t->colflag;
    a.vals = t->vals;
    a.nzrows = t->nzrows;
    a.chunk_start = t->chunk_start;
    a.chunk_row = t->chunk_row;
    a.nchunks = t->nchunks;
    a.X = (const char*)p.X;
    a.Y = (char*)p.Y;
    a.ldx_bytes = p.ldx_bytes;
    a.ldy_bytes = p.ldy_bytes;
    a.slab_bytes = VW * 16;
    a.row_bytes = a.slab_bytes;
    a.total_row_bytes = p.total_row_bytes;
    a.carry = (char*)t->carry;
    a.carry_stride = p.total_row_bytes;
    a.accumulate = p.accumulate;
    HubArgs h;
    h.hubslot = hp.hubslot;
    h.hubcols = hp.hubcols;
    h.nhub = hp.nhub;
    h.counter = hp.counters;
    auto kernel = RING ? cb_spmm_ring_kernel<Op, VW, (CB_RING_D <= VW ? CB_RING_D : VW), BT, FULL> : cb_spmm_hub_kernel<Op, VW, 1, U, BT, FULL>;
    const unsigned nslabs = (unsigned)((p.total_row_bytes + a.slab_bytes - 1) / a.slab_bytes);
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hp.smem_bytes) != cudaSuccess) { cudaGetLastError(); return CB_HUB_FALLBACK; }
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)hp.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)hp.cluster, nslabs, 1);
    cfg.blockDim = dim3(BT, 1, 1);
    cfg.dynamicSmemBytes = hp.smem_bytes;
    cfg.stream = p.stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg) != cudaSuccess || nclusters < 1) { cudaGetLastError(); return CB_HUB_FALLBACK; }
    // no more CTAs than there is work for: one warp takes 32/VW chunks at a time
    const int64_t warps_needed = (t->nchunks + (32 / VW) - 1) / (32 / VW);
    const int64_t clusters_needed = (warps_needed + (int64_t)hp.cluster * (BT / 32) - 1) / ((int64_t)hp.cluster * (BT / 32));
    if ((int64_t)nclusters > clusters_needed) nclusters = (int)clusters_needed;
    cfg.gridDim.x = (unsigned)(nclusters * hp.cluster);
    CB_CUDA(p.ctx, cudaMemsetAsync(hp.counters, 0, nslabs * sizeof(unsigned), p.stream));
    {
        cb_prof_scope prof(p.ctx, p.stream, CB_PROF_SPMM);
        CB_CUDA(p.ctx, cudaLaunchKernelEx(&cfg, kernel, a, h));
    }
    CB_LAUNCHED(p.ctx);
    CB_CUDA(p.ctx, cudaGetLastError());
    return CB_OK;
}

template <class Op, int VW>
static int launch_hub_vw(const LaunchParams& p) {
    const bool full = p.total_row_bytes % (VW * 16) == 0;
    if (p.hub->ring) return full ? launch_hub_layout<Op, VW, true, true>(p) : launch_hub_layout<Op, VW, false, true>(p);
    return full ? launch_hub_layout<Op, VW, true, false>(p) : launch_hub_layout<Op, VW, false, false>(p);
}

template <class Op>
static int launch_hub(const LaunchParams& p) {
    switch (p.hub->slab_bytes) {
        case 128: return launch_hub_vw<Op, 8>(p);
        case 256: return launch_hub_vw<Op, 16>(p);
        case 512: return launch_hub_vw<Op, 32>(p);
    }
    return CB_HUB_FALLBACK;
}

template <class Op>
static int launch_op(const LaunchParams& p) {
    const cb_tile* t = p.t;
    if (t->nnz > 0) {
        // CB_K2_SLAB=<bytes>: run panels wider than this as sequential column slabs of that width (gridDim.y, scheduled slab
        // after slab), so the X rows one pass touches are narrower and more of them stay in L2 (experiments / cb_k2_plan)
        static const int slab_env = getenv("CB_K2_SLAB") ? atoi(getenv("CB_K2_SLAB")) : 0;
        const int slab_force = p.slab_bytes > 0 ? p.slab_bytes : slab_env;
        int nvec = p.total_row_bytes / 16;
        if (slab_force >= 16 && nvec > slab_force / 16) nvec = slab_force / 16;
        int s = CB_HUB_FALLBACK;
        if (p.hub) s = launch_hub<Op>(p);
        if (s == CB_HUB_FALLBACK) {
            // X rows this tile touches: mostly L2-resident (R-MAT scale <= 22 class) or streaming from DRAM?
            static const int force_env = getenv("CB_K2_POINT") ? atoi(getenv("CB_K2_POINT")) : -1;  // 0 deep, 1 wide (experiments)
            const int force = p.point >= 0 ? p.point : force_env;
            const bool wide = force >= 0 ? force == 1 : (nvec <= 16 && (double)t->nzc * (double)p.total_row_bytes < 1.0e9);
            // 64-bit element types need more registers per gathered vector's arithmetic: one CTA per SM fewer
            constexpr int WB = sizeof(typename Op::T) == 8 ? CB_WIDE_B64 : CB_WIDE_B;
            constexpr int DB = sizeof(typename Op::T) == 8 ? CB_DEEP_B64 : CB_DEEP_B;
            // narrow panels (rows of 16 or 32 bytes: SpMV, k <= 8 fp32, boolean k <= 32): with the 4-lane layout 3 or 2 of the 4
            // lanes of a virtual warp carry no columns.  CB_K2_NARROW=1 runs them on 1- and 2-lane virtual warps (32 / 16 row
            // walkers per warp) - the same walker, other template arguments; opt-in until it has been timed on hardware
            static const bool narrow = getenv("CB_K2_NARROW") && atoi(getenv("CB_K2_NARROW")) != 0;
            static const int pipe_env = getenv("CB_K2_PIPE") ? atoi(getenv("CB_K2_PIPE")) : -1;
            const int pipe = p.pipe >= 0 ? p.pipe : (pipe_env >= 0 ? pipe_env : CB_PIPE_DEFAULT);
            constexpr bool W64 = sizeof(typename Op::T) == 8;
            (void)W64;
            if (p.win_colflag && slab_force < 16 && nvec >= 16) {
                if (nvec == 16) s = launch_win<Op, 16, 1, CB_DEEP_U, DB>(p);
                else if (nvec == 32) s = launch_win<Op, 32, 1, CB_DEEP_U, DB>(p);
                else if (nvec == 64) s = launch_win<Op, 32, 2, CB_DEEP_U / 2, DB>(p);
            } else if (pipe == 16 && slab_force < 16 && (nvec == 8 || nvec == 16 || nvec == 32)) {
                s = nvec == 8 ? launch_tma<Op, 8>(p) : nvec == 16 ? launch_tma<Op, 16>(p) : launch_tma<Op, 32>(p);
            }
            if (s == CB_HUB_FALLBACK && p.persist_counter && slab_force < 16 && nvec > 4) {
                if (nvec <= 8) s = wide ? launch_persist<Op, 8, 1, CB_WIDE_U, WB>(p) : launch_persist<Op, 8, 1, CB_DEEP_U, DB>(p);
                else if (nvec <= 16) s = wide ? launch_persist<Op, 16, 1, CB_WIDE_U, WB>(p) : launch_persist<Op, 16, 1, CB_DEEP_U, DB>(p);
                else if (nvec <= 32) s = wide ? launch_persist<Op, 32, 1, CB_WIDE_U, WB>(p) : launch_persist<Op, 32, 1, CB_DEEP_U, DB>(p);
                else if (nvec <= 64) s = launch_persist<Op, 32, 2, CB_DEEP_U / 2, DB>(p);
            }
            if (s != CB_HUB_FALLBACK) {
            } else if (pipe == 1 && nvec > 4) {
                if (nvec <= 8) s = wide ? launch_pf<Op, 8, 1, CB_WIDE_U, WB>(p) : launch_pf<Op, 8, 1, CB_DEEP_U, DB>(p);
                else if (nvec <= 16) s = wide ? launch_pf<Op, 16, 1, CB_WIDE_U, WB>(p) : launch_pf<Op, 16, 1, CB_DEEP_U, DB>(p);
                else if (nvec <= 32) s = wide ? launch_pf<Op, 32, 1, CB_WIDE_U, WB>(p) : launch_pf<Op, 32, 1, CB_DEEP_U, DB>(p);
                else s = launch_pf<Op, 32, 2, CB_DEEP_U / 2, DB>(p);
            }
#ifdef CB_BUILD_K2P                    // the register-ring walk: measured slower than K2 on every workload (profiles/r02_sweep_b_k2p_l2hints.jsonl)
            else if (pipe == 4 && nvec > 4) {
                if (nvec <= 8) s = launch_pipe<Op, 8, 1, 4, (W64 ? 3 : 4)>(p);
                else if (nvec <= 16) s = launch_pipe<Op, 16, 1, 4, (W64 ? 3 : 4)>(p);
                else if (nvec <= 32) s = launch_pipe<Op, 32, 1, 4, (W64 ? 3 : 4)>(p);
                else s = launch_pipe<Op, 32, 2, 4, (W64 ? 2 : 3)>(p);
            } else if (pipe == 8 && nvec > 4) {
                if (nvec <= 8) s = launch_pipe<Op, 8, 1, 8, (W64 ? 2 : 3)>(p);
                else if (nvec <= 16) s = launch_pipe<Op, 16, 1, 8, (W64 ? 2 : 3)>(p);
                else if (nvec <= 32) s = launch_pipe<Op, 32, 1, 8, (W64 ? 2 : 3)>(p);
                else s = launch_pipe<Op, 32, 2, 4, (W64 ? 2 : 3)>(p);
            }
#endif
            else if (narrow && nvec == 1) s = launch_layout<Op, 1, 1, 1, 4>(p);
            else if (narrow && nvec <= 2) s = launch_layout<Op, 2, 1, 2, 4>(p);
            else if (nvec <= 4) s = launch_layout<Op, 4, 1, 4, 4>(p);
            else if (nvec <= 8) s = wide ? launch_layout<Op, 8, 1, CB_WIDE_U, WB>(p) : launch_layout<Op, 8, 1, CB_DEEP_U, DB>(p);
            else if (nvec <= 16) s = wide ? launch_layout<Op, 16, 1, CB_WIDE_U, WB>(p) : launch_layout<Op, 16, 1, CB_DEEP_U, DB>(p);
            else if (nvec <= 32) s = wide ? launch_layout<Op, 32, 1, CB_WIDE_U, WB>(p) : launch_layout<Op, 32, 1, CB_DEEP_U, DB>(p);
            else s = launch_layout<Op, 32, 2, CB_DEEP_U / 2, DB>(p);
        }
        if (s != CB_OK) return s;
        if (t->nsplit > 0) {
            FixupArgs f;
            f.split_row = t->split_row;
            f.nsplit = t->nsplit;
            f.nzrows = t->nzrows;
            f.rowptr = t->rowptr;
            f.chunk_len = t->chunk_len;
            f.carry = (const char*)t->carry;
            f.carry_stride = p.total_row_bytes;
            f.Y = (char*)p.Y;
            f.ldy_bytes = p.ldy_bytes;
            f.total_row_bytes = p.total_row_bytes;
            f.accumulate = p.accumulate;
            {
                cb_prof_scope prof(p.ctx, p.stream, CB_PROF_FIXUP);
                cb_fixup_kernel<Op><<<(unsigned)((t->nsplit + 7) / 8), 256, 0, p.stream>>>(f);
            }
            CB_LAUNCHED(p.ctx);
            CB_CUDA(p.ctx, cudaGetLastError());
        }
    }
    return CB_OK;
}

}  // namespace cbk

// per-family entry points (one translation unit each so they compile in parallel)
int cb_launch_plus_times_f(int dtype, int akind, const cbk::LaunchParams& p);   // f32, f64
int cb_launch_plus_times_i(int dtype, int akind, const cbk::LaunchParams& p);   // i32, i64
int cb_launch_min_plus(int dtype, const cbk::LaunchParams& p);
int cb_launch_select_max(int dtype, const cbk::LaunchParams& p);
int cb_launch_or_and(int akind, const cbk::LaunchParams& p);
