// Runs the product's K2 (cb_spmm_kernel), its persistent variants K2H (hub rows in cluster shared memory) and K2R (gathers
// through a cp.async ring) and the fix-up kernel - compiled UNMODIFIED from csrc/cb_spmm_kernel.cuh and
// csrc/cb_spmm_hub_kernel.cuh - on the lock-step warp / cluster emulator, on small random tiles with hub rows, empty rows,
// ragged panel widths, column slabs and the accumulate mode; compares with a scalar loop over the same semiring functors
// and, for the variants, bit for bit with K2.  Built with -fsanitize=address,undefined: every device buffer and every CTA's
// shared memory is an exactly-sized heap vector, so an out-of-bounds access of a kernel is an ASan report.
#include "cuda_emul.h"
#include <cstdio>
#include <random>
#include "cb_spmm_hub_kernel.cuh"

using namespace cbk;

struct HostTile {                       // host restatement of the tile layout (csrc/cb_tile.cu), for the emulator only
    int64_t m, n, nnz, nzr, nchunks, nsplit;
    int32_t L;
    std::vector<int32_t> colflag, nzrows, rowptr, emptyrows, chunk_start, chunk_row, split_row;
};

static HostTile build(int64_t m, int64_t n, const std::vector<std::vector<int32_t>>& rows, int32_t L) {
    HostTile t{};
    t.m = m; t.n = n; t.L = L;
    for (int64_t r = 0; r < m; ++r) {
        if (rows[r].empty()) { t.emptyrows.push_back((int32_t)r); continue; }
        t.nzrows.push_back((int32_t)r);
        t.rowptr.push_back((int32_t)t.colflag.size());
        for (size_t q = 0; q < rows[r].size(); ++q)
            t.colflag.push_back((int32_t)((uint32_t)rows[r][q] | (q + 1 == rows[r].size() ? 0x80000000u : 0u)));
    }
    t.nnz = (int64_t)t.colflag.size();
    t.rowptr.push_back((int32_t)t.nnz);
    t.nzr = (int64_t)t.nzrows.size();
    t.nchunks = (t.nnz + L - 1) / L;
    for (int64_t g = 0; g < t.nchunks; ++g) {                       // chunk_kernel of cb_tile.cu
        const int64_t pos = g * L;
        int64_t lo = 0, hi = t.nzr;
        while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (t.rowptr[mid] <= pos) lo = mid; else hi = mid; }
        const int32_t rs = t.rowptr[lo], re = t.rowptr[lo + 1];
        if (re - rs > L) { t.chunk_start.push_back((int32_t)pos); t.chunk_row.push_back((int32_t)((uint32_t)lo | (pos > rs ? 0x80000000u : 0u))); }
        else { t.chunk_start.push_back(rs); t.chunk_row.push_back((int32_t)lo); }
    }
    t.chunk_start.push_back((int32_t)t.nnz);
    for (int64_t i = 0; i < t.nzr; ++i) if (t.rowptr[i + 1] - t.rowptr[i] > L) t.split_row.push_back((int32_t)i);
    t.nsplit = (int64_t)t.split_row.size();
    return t;
}

template <class T> static T rnd_val(std::mt19937& g) { return (T)(1 + g() % 9); }
template <> float rnd_val<float>(std::mt19937& g) { return (float)(1 + g() % 64) / 8.0f; }
template <> double rnd_val<double>(std::mt19937& g) { return (double)(1 + g() % 1024) / 32.0; }

// one multiply on the emulator; returns the number of mismatching elements.  hub_cs > 0 runs the hub variant (K2H) in
// clusters of hub_cs CTAs with the `nhub` most frequent columns resident in (distributed) shared memory and also
// requires its result to be bit-identical to plain K2's.
template <class Op, int VW, int R, int U, bool FULL>
static long run_case(const char* name, int64_t m, int64_t n, int k_elems /* elements of Op::T per panel row */, int32_t L, int hub, unsigned seed,
                     bool accumulate, int hub_cs = 0, int nhub = 0, int ring = 0, int pipe = 0) {
    typedef typename Op::T T;
    typedef typename Op::TA TA;
    std::mt19937 g(seed);
    std::vector<std::vector<int32_t>> rows((size_t)m);
    for (int64_t r = 0; r < m; ++r) {
        int deg = (r % 5 == 3) ? 0 : (int)(g() % 7);                // empty rows and short rows
        if (r == 2 || r == m - 1) deg = hub;                         // hub rows: split over several chunks, last row too
        if (r == 7) deg = L;                                         // exactly one chunk long
        if (r == 9) deg = L + 1;                                     // one more than a chunk
        std::vector<char> used((size_t)n, 0);
        for (int q = 0; q < deg && q < n; ++q) { int c; do { c = (int)(g() % n); } while (used[c]); used[c] = 1; }
        for (int c = 0; c < n; ++c) if (used[c]) rows[r].push_back(c);
    }
    HostTile t = build(m, n, rows, L);
    std::vector<TA> vals((size_t)t.nnz);
    for (auto& v : vals) v = Op::akind == A_BOOL ? (TA)(g() % 4 != 0) : rnd_val<TA>(g);
    const int row_bytes = (int)((k_elems * sizeof(T) + 15) / 16 * 16);
    const int ld_elems = row_bytes / (int)sizeof(T);
    std::vector<T> X((size_t)n * ld_elems), Y((size_t)m * ld_elems), Yref((size_t)m * ld_elems);
    for (auto& x : X) x = rnd_val<T>(g);
    for (auto& y : Y) y = rnd_val<T>(g);                            // stale content: must be overwritten or folded in
    // scalar reference with the same functors
    for (int64_t r = 0; r < m; ++r)
        for (int c = 0; c < ld_elems; ++c) {
            bool first = true;
            T acc = Op::id();
            int64_t i = std::lower_bound(t.nzrows.begin(), t.nzrows.end(), (int32_t)r) - t.nzrows.begin();
            if (i < t.nzr && t.nzrows[i] == r)
                for (int32_t p = t.rowptr[i]; p < t.rowptr[i + 1]; ++p) {
                    const T prod = Op::mul(Op::akind != A_PATTERN ? vals[p] : TA(), X[(size_t)(t.colflag[p] & 0x7fffffff) * ld_elems + c]);
                    acc = (Op::first_touch && first) ? prod : Op::add(prod, acc);
                    first = false;
                }
            T& out = Yref[(size_t)r * ld_elems + c];
            if (accumulate) out = first ? Y[(size_t)r * ld_elems + c] : Op::add(Y[(size_t)r * ld_elems + c], acc);
            else out = acc;
        }
    // K3: identity into the rows K2 will not write
    if (!accumulate)
        for (int32_t r : t.emptyrows) for (int c = 0; c < ld_elems; ++c) Y[(size_t)r * ld_elems + c] = Op::id();
    std::vector<char> carry((size_t)2 * t.nchunks * row_bytes);
    SpmmArgs a{};
    a.colflag = t.colflag.data(); a.vals = vals.data(); a.nzrows = t.nzrows.data();
    a.chunk_start = t.chunk_start.data(); a.chunk_row = t.chunk_row.data(); a.nchunks = t.nchunks;
    a.X = (const char*)X.data(); a.Y = (char*)Y.data();
    a.ldx_bytes = row_bytes; a.ldy_bytes = row_bytes;
    a.slab_bytes = VW * R * 16; a.row_bytes = a.slab_bytes; a.total_row_bytes = row_bytes;
    a.carry = carry.data(); a.carry_stride = row_bytes; a.accumulate = accumulate ? 1 : 0;
    constexpr int NV = 32 / VW;
    dim3 grid((unsigned)((t.nchunks + 8 * NV - 1) / (8 * NV)), (unsigned)((row_bytes + a.slab_bytes - 1) / a.slab_bytes));
    long long syncs = 0;
    std::vector<T> Yplain;
    if (pipe) {
        // K2P (register-ring walk, U = ring depth) after plain K2 on a copy: must reproduce it bit for bit
        std::vector<T> Ysave = Y;
        emul::launch(grid, dim3(256), [&] { cb_spmm_kernel<Op, VW, R, (U < 4 ? U : 4), 3, FULL>(a); });
        Yplain = Y;
        Y = Ysave;
        a.Y = (char*)Y.data();
        std::fill(carry.begin(), carry.end(), (char)0x77);
        std::vector<uint8_t> hubcls;
        std::vector<int32_t> wcolflag;
        std::vector<T> wpanel;
        if (pipe == 3) {     // the round-1 walk with the entry prefetch
            syncs = emul::launch(grid, dim3(256), [&] { cb_spmm_kernel<Op, VW, R, (U < 4 ? U : 4), 3, FULL, true>(a); });
        } else if (pipe == 6) {     // K2 with persistent warps: two CTAs take chunk groups from a counter until none is left
            unsigned counter = 0;
            syncs = emul::launch(dim3(2, 1), dim3(256), [&] { cb_spmm_persist_kernel<Op, VW, R, (U < 4 ? U : 4), 3, FULL>(a, &counter); });
            if ((int64_t)counter < t.nchunks) { std::printf("%-28s persistent warps left chunks behind\n", name); ++syncs; Y[0] = T(); }
        } else if (pipe == 5) {     // K2W: the nhub most used columns as bit 30 + rank in the column stream, their X rows packed into a panel
            std::vector<int32_t> cnt((size_t)n, 0), order((size_t)n);
            for (int32_t cf : t.colflag) ++cnt[cf & 0x7fffffff];
            for (int64_t c = 0; c < n; ++c) order[c] = (int32_t)c;
            std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return cnt[x] > cnt[y]; });
            int h = 0;
            while (h < nhub && h < n && cnt[order[h]] >= 2) ++h;
            std::vector<int32_t> rank_of((size_t)n, -1);
            for (int r = 0; r < h; ++r) rank_of[order[r]] = r;
            wcolflag.resize((size_t)t.nnz);
            for (int64_t p = 0; p < t.nnz; ++p) {
                const int32_t cf = t.colflag[p], r = rank_of[cf & 0x7fffffff];
                wcolflag[p] = r >= 0 ? (int32_t)(((uint32_t)cf & 0x80000000u) | 0x40000000u | (uint32_t)r) : cf;
            }
            wpanel.resize((size_t)std::max(h, 1) * ld_elems);                          // exactly h rows: a rank >= h read is an ASan report
            for (int r = 0; r < h; ++r) std::memcpy(&wpanel[(size_t)r * ld_elems], &X[(size_t)order[r] * ld_elems], row_bytes);
            a.colflag = wcolflag.data();
            a.win_delta = (const char*)wpanel.data() - (const char*)X.data();
            syncs = emul::launch(grid, dim3(256), [&] { cb_spmm_kernel<Op, VW, R, (U < 4 ? U : 4), 3, FULL, false, true>(a); });
        } else if (pipe == 2) {     // with L2 residency hints: per-nonzero use class of its column (cb_hub.cu), classes <= 2 marked
            std::vector<int32_t> cnt((size_t)n, 0), order((size_t)n);
            for (int32_t cf : t.colflag) ++cnt[cf & 0x7fffffff];
            for (int64_t c = 0; c < n; ++c) order[c] = (int32_t)c;
            std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return cnt[x] > cnt[y]; });
            std::vector<uint8_t> cls_of((size_t)n, 255);
            for (int64_t r = 0; r < n; ++r) if (cnt[order[r]] >= 2) cls_of[order[r]] = (uint8_t)(63 - __builtin_clzll((unsigned long long)(r + 1)));
            hubcls.resize((size_t)t.nnz);                                  // exactly nnz bytes: a read past the chunk is an ASan report
            for (int64_t p = 0; p < t.nnz; ++p) hubcls[p] = cls_of[t.colflag[p] & 0x7fffffff];
            a.hubcls = hubcls.data();
            a.cls_max = 2;
            if constexpr (U <= VW) syncs = emul::launch(grid, dim3(256), [&] { cb_spmm_pipe_kernel<Op, VW, R, U, 3, FULL, true>(a); });
        } else if constexpr (U <= VW) syncs = emul::launch(grid, dim3(256), [&] { cb_spmm_pipe_kernel<Op, VW, R, U, 3, FULL>(a); });
    } else if (hub_cs > 0) {
        // plain K2 on a copy first: K2H must reproduce it bit for bit (same walk, same fold order)
        std::vector<T> Ysave = Y;
        emul::launch(grid, dim3(256), [&] { cb_spmm_kernel<Op, VW, R, U, 3, FULL>(a); });
        Yplain = Y;
        Y = Ysave;
        a.Y = (char*)Y.data();
        std::fill(carry.begin(), carry.end(), (char)0x77);
        // hub ranks: columns by descending count, ties by ascending column (cb_hub_select_host)
        std::vector<int32_t> cnt((size_t)n, 0), order((size_t)n);
        for (int32_t cf : t.colflag) ++cnt[cf & 0x7fffffff];
        for (int64_t c = 0; c < n; ++c) order[c] = (int32_t)c;
        std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return cnt[x] > cnt[y]; });
        int have = 0;
        while (have < n && cnt[order[have]] > 0) ++have;
        nhub = std::min(nhub, have);
        std::vector<int32_t> hubcols(order.begin(), order.begin() + nhub);               // exactly nhub entries: reading rank >= nhub is an ASan report
        std::vector<uint16_t> rank_of((size_t)n, 0xffff), hubslot((size_t)t.nnz);
        for (int r = 0; r < have && r < 0xffff; ++r) rank_of[order[r]] = (uint16_t)r;    // ranks beyond nhub exist too: the kernel must ignore them
        for (int64_t p = 0; p < t.nnz; ++p) hubslot[p] = rank_of[t.colflag[p] & 0x7fffffff];
        std::vector<unsigned> counters(grid.y, 0u);
        HubArgs h{};
        h.hubslot = hubslot.data(); h.hubcols = hubcols.data(); h.nhub = nhub; h.counter = counters.data();
        const size_t smem = (size_t)((nhub + hub_cs - 1) / hub_cs) * a.slab_bytes;
        if (ring == 0)
            emul::launch_cluster(dim3((unsigned)(2 * hub_cs), grid.y), dim3(64), (unsigned)hub_cs, smem ? smem : 16,
                                 [&] { cb_spmm_hub_kernel<Op, VW, R, U, 64, FULL>(a, h); });
        else if constexpr (VW >= 2 && U >= 2) {
            // K2R: U doubles as the ring depth D here; shared memory = hub slots + one ring of D slots per virtual warp, exactly
            constexpr int D = U <= VW ? U : VW;
            if (nhub == 0) { h.hubslot = nullptr; h.hubcols = nullptr; }             // ring without hub data must not touch it
            const size_t ring_bytes = (size_t)(64 / 32) * NV * D * a.slab_bytes;
            emul::launch_cluster(dim3((unsigned)(2 * hub_cs), grid.y), dim3(64), (unsigned)hub_cs, smem + ring_bytes,
                                 [&] { cb_spmm_ring_kernel<Op, VW, D, 64, FULL>(a, h); });
        }
    } else {
        syncs = emul::launch(grid, dim3(256), [&] { cb_spmm_kernel<Op, VW, R, U, 3, FULL>(a); });
    }
    if (t.nsplit) {
        FixupArgs f{};
        f.split_row = t.split_row.data(); f.nsplit = t.nsplit; f.nzrows = t.nzrows.data(); f.rowptr = t.rowptr.data();
        f.chunk_len = t.L; f.carry = carry.data(); f.carry_stride = row_bytes; f.Y = (char*)Y.data(); f.ldy_bytes = row_bytes;
        f.total_row_bytes = row_bytes; f.accumulate = accumulate ? 1 : 0;
        emul::launch(dim3((unsigned)((t.nsplit + 7) / 8)), dim3(256), [&] { cb_fixup_kernel<Op>(f); });
    }
    long bad = 0;
    if (hub_cs > 0 || pipe) {
        if (t.nsplit) {      // fix-up of the plain run, so the two results are comparable row for row
            FixupArgs f{};
            f.split_row = t.split_row.data(); f.nsplit = t.nsplit; f.nzrows = t.nzrows.data(); f.rowptr = t.rowptr.data();
            f.chunk_len = t.L; f.carry = carry.data(); f.carry_stride = row_bytes; f.Y = (char*)Yplain.data(); f.ldy_bytes = row_bytes;
            f.total_row_bytes = row_bytes; f.accumulate = accumulate ? 1 : 0;
            emul::launch(dim3((unsigned)((t.nsplit + 7) / 8)), dim3(256), [&] { cb_fixup_kernel<Op>(f); });
        }
        if (std::memcmp(Yplain.data(), Y.data(), Y.size() * sizeof(T)) != 0) { ++bad; std::printf("%-28s K2H differs from K2\n", name); }
    }
    for (size_t q = 0; q < Y.size(); ++q) {
        const double x = (double)Y[q], y = (double)Yref[q];
        if (std::is_floating_point<T>::value ? std::abs(x - y) > 1e-5 * std::max(1.0, std::abs(y)) : Y[q] != Yref[q]) ++bad;
    }
    std::printf("%-28s m=%lld nnz=%lld chunks=%lld split=%lld slabs=%u warps=%u syncs=%lld mismatches=%ld\n", name, (long long)m, (long long)t.nnz,
                (long long)t.nchunks, (long long)t.nsplit, grid.y, grid.x * 8, syncs, bad);
    return bad;
}

int main(int argc, char** argv) {
    long bad = 0;
    const int rounds = argc > 1 ? std::atoi(argv[1]) : 1;
    for (int it = 0; it < rounds; ++it) {
    const unsigned sd = 100u * (unsigned)it;
    // k=64 fp32: half-warp layout, exact fill (FULL); wide and deep operating points
    bad += run_case<PlusTimes<float, A_SAME>, 16, 1, 4, true>("pt_f32 VW16 U4 full", 61, 97, 64, 32, 150, 1 + sd, false);
    bad += run_case<PlusTimes<float, A_SAME>, 16, 1, 8, true>("pt_f32 VW16 U8 accumulate", 61, 97, 64, 32, 150, 2 + sd, true);
    // ragged width: 25 vectors in a 32-lane layout (FULL = false), and a panel wider than one slab
    bad += run_case<PlusTimes<float, A_SAME>, 32, 1, 8, false>("pt_f32 VW32 k=100", 40, 200, 100, 32, 90, 3 + sd, false);
    bad += run_case<PlusTimes<float, A_PATTERN>, 32, 1, 8, false>("pt_f32 pattern 3 slabs", 33, 60, 300, 32, 55, 4 + sd, false);
    // fp64 with two vectors per lane (k=128 doubles = 1 KB rows)
    bad += run_case<PlusTimes<double, A_SAME>, 32, 2, 4, true>("pt_f64 VW32 R2", 30, 80, 128, 32, 70, 5 + sd, false);
    bad += run_case<PlusTimes<double, A_BOOL>, 32, 2, 4, false>("pt_f64 boolA R2 ragged acc", 30, 80, 100, 32, 70, 6 + sd, true);
    // integer semirings: quarter-warps, first-touch select, packed booleans
    bad += run_case<MinPlus<int32_t>, 8, 1, 4, true>("minplus_i32 VW8", 70, 64, 32, 32, 60, 7 + sd, false);
    bad += run_case<SelectMax<int64_t>, 8, 1, 8, false>("selectmax_i64 VW8 k=13", 45, 50, 13, 32, 45, 8 + sd, false);
    bad += run_case<OrAnd<A_PATTERN>, 4, 1, 4, false>("or_and VW4 k=32 bytes", 50, 40, 8, 32, 38, 9 + sd, false);      // 8 uint32 = 32 bool columns
    bad += run_case<OrAnd<A_BOOL>, 4, 1, 4, true>("or_and boolA VW4 acc", 50, 40, 16, 32, 38, 10 + sd, true);
    // a larger tile: several blocks, chunk length 64, hub rows of ~11 chunks
    bad += run_case<PlusTimes<float, A_SAME>, 16, 1, 4, true>("pt_f32 VW16 larger", 400, 900, 64, 64, 700, 11 + sd, false);
    bad += run_case<MinPlus<int64_t>, 16, 1, 8, true>("minplus_i64 VW16 larger acc", 300, 500, 32, 64, 400, 12 + sd, true);
    // narrow panels on 2- and 1-lane virtual warps (CB_K2_NARROW=1): 32- and 16-byte rows
    bad += run_case<PlusTimes<float, A_SAME>, 2, 1, 2, true>("narrow pt_f32 VW2 k=8", 61, 97, 8, 32, 90, 41 + sd, false);
    bad += run_case<PlusTimes<float, A_SAME>, 2, 1, 2, false>("narrow pt_f32 VW2 k=5 acc", 61, 97, 5, 32, 90, 42 + sd, true);
    bad += run_case<MinPlus<int64_t>, 1, 1, 1, false>("narrow minplus_i64 VW1 k=1", 70, 64, 1, 32, 60, 43 + sd, false);
    bad += run_case<PlusTimes<double, A_BOOL>, 1, 1, 1, true>("narrow pt_f64 boolA VW1 k=2", 45, 50, 2, 32, 45, 44 + sd, true);
    bad += run_case<OrAnd<A_PATTERN>, 2, 1, 2, true>("narrow or_and VW2 k=32 bytes", 50, 40, 8, 32, 38, 45 + sd, false);
    bad += run_case<SelectMax<int32_t>, 1, 1, 1, false>("narrow selectmax_i32 VW1 k=3", 45, 50, 3, 32, 45, 46 + sd, false);
    // K2P, the pipelined walk (register ring of depth D = the U argument): every layout, row ends in every position,
    // chunks shorter than a step, shorter than the ring, one step and two steps long, ragged widths, accumulate mode
    bad += run_case<PlusTimes<float, A_SAME>, 16, 1, 4, true>("pipe pt_f32 VW16 D4", 61, 97, 64, 32, 150, 51 + sd, false, 0, 0, 0, 1);
    bad += run_case<PlusTimes<float, A_SAME>, 16, 1, 8, true>("pipe pt_f32 VW16 D8 acc", 61, 97, 64, 32, 150, 52 + sd, true, 0, 0, 0, 1);
    bad += run_case<PlusTimes<float, A_SAME>, 32, 1, 8, false>("pipe pt_f32 VW32 k=100 D8", 40, 200, 100, 32, 90, 53 + sd, false, 0, 0, 0, 1);
    bad += run_case<PlusTimes<float, A_PATTERN>, 32, 1, 4, false>("pipe pt_f32 pat 3 slabs D4", 33, 60, 300, 32, 55, 54 + sd, false, 0, 0, 0, 1);
    bad += run_case<PlusTimes<double, A_SAME>, 32, 2, 4, true>("pipe pt_f64 VW32 R2 D4", 30, 80, 128, 32, 70, 55 + sd, false, 0, 0, 0, 1);
    bad += run_case<PlusTimes<double, A_BOOL>, 32, 2, 4, false>("pipe pt_f64 boolA R2 ragged acc", 30, 80, 100, 32, 70, 56 + sd, true, 0, 0, 0, 1);
    bad += run_case<MinPlus<int32_t>, 8, 1, 4, true>("pipe minplus_i32 VW8 D4", 70, 64, 32, 32, 60, 57 + sd, false, 0, 0, 0, 1);
    bad += run_case<MinPlus<int32_t>, 8, 1, 8, true>("pipe minplus_i32 VW8 D8 acc", 70, 64, 32, 32, 60, 58 + sd, true, 0, 0, 0, 1);
    bad += run_case<SelectMax<int64_t>, 8, 1, 8, false>("pipe selectmax_i64 VW8 k=13", 45, 50, 13, 32, 45, 59 + sd, false, 0, 0, 0, 1);
    bad += run_case<OrAnd<A_PATTERN>, 8, 1, 4, false>("pipe or_and VW8 k=96 bytes", 50, 40, 24, 32, 38, 60 + sd, false, 0, 0, 0, 1);
    bad += run_case<PlusTimes<float, A_SAME>, 16, 1, 4, true>("pipe pt_f32 VW16 larger D4", 400, 900, 64, 64, 700, 61 + sd, false, 0, 0, 0, 1);
    bad += run_case<MinPlus<int64_t>, 16, 1, 8, true>("pipe minplus_i64 larger acc D8", 300, 500, 32, 64, 400, 62 + sd, true, 0, 0, 0, 1);
    bad += run_case<PlusTimes<float, A_SAME>, 32, 1, 8, true>("pipe pt_f32 VW32 L=200 D8", 300, 700, 128, 200, 900, 63 + sd, false, 0, 0, 0, 1);
    bad += run_case<PlusTimes<float, A_BOOL>, 8, 1, 8, true>("pipe pt_f32 boolA VW8 L=40 D8", 200, 300, 32, 40, 170, 64 + sd, true, 0, 0, 0, 1);
    bad += run_case<PlusTimes<float, A_SAME>, 16, 1, 4, true>("prefetch pt_f32 VW16 U4", 61, 97, 64, 32, 150, 71 + sd, false, 0, 0, 0, 3);
    bad += run_case<PlusTimes<float, A_SAME>, 32, 1, 4, false>("prefetch pt_f32 VW32 k=100 acc", 40, 200, 100, 32, 90, 72 + sd, true, 0, 0, 0, 3);
    bad += run_case<PlusTimes<double, A_BOOL>, 32, 2, 4, false>("prefetch pt_f64 boolA R2 ragged", 30, 80, 100, 32, 70, 73 + sd, false, 0, 0, 0, 3);
    bad += run_case<MinPlus<int32_t>, 8, 1, 4, true>("prefetch minplus_i32 VW8", 70, 64, 32, 32, 60, 74 + sd, false, 0, 0, 0, 3);
    bad += run_case<SelectMax<int64_t>, 8, 1, 4, false>("prefetch selectmax_i64 k=13", 45, 50, 13, 32, 45, 75 + sd, false, 0, 0, 0, 3);
    bad += run_case<PlusTimes<float, A_SAME>, 32, 1, 4, true>("prefetch pt_f32 VW32 L=1200", 300, 3000, 128, 1200, 2900, 76 + sd, false, 0, 0, 0, 3);
    bad += run_case<OrAnd<A_PATTERN>, 8, 1, 4, false>("prefetch or_and VW8 L=700", 200, 2000, 24, 700, 1800, 77 + sd, false, 0, 0, 0, 3);
    // K2 with persistent warps (single column slab)
    bad += run_case<PlusTimes<float, A_SAME>, 16, 1, 4, true>("persist pt_f32 VW16", 61, 97, 64, 32, 150, 91 + sd, false, 0, 0, 0, 6);
    bad += run_case<PlusTimes<float, A_SAME>, 32, 1, 4, false>("persist pt_f32 VW32 k=100 acc", 40, 200, 100, 32, 90, 92 + sd, true, 0, 0, 0, 6);
    bad += run_case<PlusTimes<double, A_SAME>, 32, 2, 4, true>("persist pt_f64 R2 larger", 300, 700, 128, 64, 900, 93 + sd, false, 0, 0, 0, 6);
    bad += run_case<MinPlus<int32_t>, 8, 1, 4, true>("persist minplus_i32 VW8 larger", 400, 900, 32, 64, 700, 94 + sd, false, 0, 0, 0, 6);
    // K2W, the hub panel behind a window: hub entries flagged in the column stream and gathered from the packed panel
    bad += run_case<PlusTimes<float, A_SAME>, 32, 1, 4, true>("window pt_f32 VW32 20 hubs", 300, 700, 128, 200, 900, 81 + sd, false, 0, 20, 0, 5);
    bad += run_case<PlusTimes<float, A_SAME>, 16, 1, 4, true>("window pt_f32 VW16 all hubs acc", 61, 97, 64, 32, 150, 82 + sd, true, 0, 1000, 0, 5);
    bad += run_case<PlusTimes<double, A_SAME>, 32, 2, 4, true>("window pt_f64 R2 7 hubs", 30, 80, 128, 32, 70, 83 + sd, false, 0, 7, 0, 5);
    bad += run_case<PlusTimes<double, A_SAME>, 32, 1, 4, true>("window pt_f64 VW32 no hubs", 45, 50, 64, 32, 45, 84 + sd, false, 0, 0, 0, 5);
    bad += run_case<PlusTimes<float, A_SAME>, 32, 1, 8, true>("pipe+l2 pt_f32 VW32 L=200 D8", 300, 700, 128, 200, 900, 65 + sd, false, 0, 0, 0, 2);
    bad += run_case<PlusTimes<double, A_SAME>, 32, 2, 4, false>("pipe+l2 pt_f64 R2 ragged acc", 30, 80, 100, 32, 70, 66 + sd, true, 0, 0, 0, 2);
    bad += run_case<MinPlus<int32_t>, 8, 1, 8, true>("pipe+l2 minplus_i32 VW8 D8", 70, 64, 32, 32, 60, 67 + sd, false, 0, 0, 0, 2);
    bad += run_case<OrAnd<A_PATTERN>, 16, 1, 8, false>("pipe+l2 or_and VW16 D8", 90, 70, 50, 64, 60, 68 + sd, false, 0, 0, 0, 2);
    // K2H, the hub variant: persistent CTAs, dynamic chunks, hub rows in the shared memory of a 1/2/4-CTA cluster
    bad += run_case<PlusTimes<float, A_SAME>, 16, 1, 8, true>("hub pt_f32 VW16 cs1", 61, 97, 64, 32, 150, 21 + sd, false, 1, 20);
    bad += run_case<PlusTimes<float, A_SAME>, 16, 1, 8, true>("hub pt_f32 VW16 cs2 acc", 61, 97, 64, 32, 150, 22 + sd, true, 2, 33);
    bad += run_case<PlusTimes<float, A_PATTERN>, 32, 1, 8, false>("hub pt_f32 pat 3 slabs cs4", 33, 60, 300, 32, 55, 23 + sd, false, 4, 60);   // every column a hub
    bad += run_case<PlusTimes<double, A_BOOL>, 32, 1, 8, false>("hub pt_f64 boolA ragged cs2", 30, 80, 50, 32, 70, 24 + sd, true, 2, 7);
    bad += run_case<MinPlus<int32_t>, 8, 1, 8, true>("hub minplus_i32 VW8 cs4", 70, 64, 32, 32, 60, 25 + sd, false, 4, 30);
    bad += run_case<SelectMax<int64_t>, 8, 1, 8, false>("hub selectmax_i64 k=13 cs1", 45, 50, 13, 32, 45, 26 + sd, false, 1, 1);
    bad += run_case<OrAnd<A_PATTERN>, 8, 1, 8, false>("hub or_and VW8 cs2 nhub=0", 50, 40, 24, 32, 38, 27 + sd, false, 2, 0);
    bad += run_case<PlusTimes<float, A_SAME>, 16, 1, 8, true>("hub pt_f32 larger cs4", 400, 900, 64, 64, 700, 28 + sd, false, 4, 200);
    // K2R, the ring-pipelined variant (cp.async into a per-virtual-warp ring), with and without resident hub rows
    bad += run_case<PlusTimes<float, A_SAME>, 16, 1, 8, true>("ring pt_f32 VW16 D8 cs1", 61, 97, 64, 32, 150, 31 + sd, false, 1, 0, 1);
    bad += run_case<PlusTimes<float, A_SAME>, 16, 1, 4, true>("ring pt_f32 VW16 D4 hubs cs2 acc", 61, 97, 64, 32, 150, 32 + sd, true, 2, 33, 1);
    bad += run_case<PlusTimes<float, A_PATTERN>, 32, 1, 8, false>("ring pt_f32 pat 3 slabs cs4", 33, 60, 300, 32, 55, 33 + sd, false, 4, 60, 1);
    bad += run_case<PlusTimes<double, A_BOOL>, 32, 1, 8, false>("ring pt_f64 boolA ragged cs2", 30, 80, 50, 32, 70, 34 + sd, true, 2, 7, 1);
    bad += run_case<MinPlus<int32_t>, 8, 1, 8, true>("ring minplus_i32 VW8 D8 cs4", 70, 64, 32, 32, 60, 35 + sd, false, 4, 30, 1);
    bad += run_case<SelectMax<int64_t>, 8, 1, 2, false>("ring selectmax_i64 k=13 D2", 45, 50, 13, 32, 45, 36 + sd, false, 1, 1, 1);
    bad += run_case<OrAnd<A_PATTERN>, 8, 1, 4, false>("ring or_and VW8 D4 nhub=0", 50, 40, 24, 32, 38, 37 + sd, false, 2, 0, 1);
    bad += run_case<PlusTimes<float, A_SAME>, 16, 1, 16, true>("ring pt_f32 larger D16 cs4", 400, 900, 64, 64, 700, 38 + sd, false, 4, 200, 1);
    bad += run_case<MinPlus<int64_t>, 16, 1, 8, true>("ring minplus_i64 larger acc", 300, 500, 32, 64, 400, 39 + sd, true, 1, 0, 1);
    }
    std::printf(bad ? "EMULATION FAILED\n" : "emulation ok\n");
    return bad ? 1 : 0;
}
