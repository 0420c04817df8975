// SpMM<SR>(A, X): Y = A (x).(+) X for a 2D-distributed sparse A and a distributed dense X.
//
// The signature mirrors the reference's dense SpMV (include/CombBLAS/ParFriends.h:1924-1926) with the vector replaced
// by a DenseParMat; the algorithm is the SUMMA stage loop of Mult_AnXBn_Synch / _Overlap (ParFriends.h:1004-1235)
// run on the GPUs by cb_spmm_summa.  Compliance checks and abort codes are those of CheckSpGEMMCompliance /
// CheckSpMVCompliance (ParFriends.h:160-181, :1350-1366): GRIDMISMATCH, DIMMISMATCH.
// Rows of A without nonzeros give SR::id() in Y - the dense-output convention of the reference's dense SpMV
// (std::fill_n(localy, ysize, SR::id()), ParFriends.h:1960-1963).
#ifndef CB_PARFRIENDS_H
#define CB_PARFRIENDS_H

#include <cstdlib>
#include "DenseParMat.h"
#include "FullyDistVec.h"
#include "Semirings.h"
#include "SpParMat.h"

namespace combblas {

template <typename IU, typename NUM, typename NUV, typename UDER>
bool CheckSpMMCompliance(const SpParMat<IU, NUM, UDER>& A, const DenseParMat<IU, NUV>& X) {
    if (*(A.getcommgrid()) != *(X.getcommgrid())) {
        SpParHelper::Print("Grids are not comparable for SpMM\n");
        MPI_Abort(MPI_COMM_WORLD, GRIDMISMATCH);
        return false;
    }
    const IU ncol = A.getncol(), xrows = X.grows();
    if (ncol != xrows) {
        std::ostringstream outs;
        outs << "Can not multiply, dimensions does not match" << std::endl << ncol << " != " << xrows << std::endl;
        SpParHelper::Print(outs.str());
        MPI_Abort(MPI_COMM_WORLD, DIMMISMATCH);
        return false;
    }
    return true;
}

template <typename SR, typename IU, typename NUM, typename NUV, typename UDER>
DenseParMat<IU, typename promote_trait<NUM, NUV>::T_promote> SpMM(const SpParMat<IU, NUM, UDER>& A, const DenseParMat<IU, NUV>& X) {
    typedef typename promote_trait<NUM, NUV>::T_promote T_promote;
    static_assert(semiring_traits<SR>::supported,
                  "this semiring / type combination is not implemented by the B200 SpMM (PlusTimes, MinPlus, SelectMax<bool,T>, bool OR-AND)");
    static_assert(std::is_same<T_promote, NUV>::value, "the dense operand must already have the promoted type");
    CheckSpMMCompliance(A, X);
    std::shared_ptr<CommGrid> grid = A.getcommgrid();
    cb_ctx* ctx = grid->GetContext();
    const IU gm = A.getnrow(), gn = A.getncol(), gk = X.gcols();
    const IU lm = A.getlocalrows(), kl = X.getlocalcols();
    DenseParMat<IU, T_promote> Y(SR::id(), grid, lm, kl);
    cb_tile* tile = A.DeviceTile();
    // host panels in, host panel out: column slabs pipelined over H2D / stage loop / D2H on the device side
    cb_check(cb_spmm_summa_host(ctx, tile, X.data(), kl, Y.data(), kl, semiring_traits<SR>::op, gm, gn, gk, cb_dtype_of<NUV>::value), ctx,
             "cb_spmm_summa_host");
    return Y;
}

// ---------------------------------------------------------------------------------------------------------------
// Dense SpMV: y = A (x).(+) x with a FullyDistVec operand and result, the reference's SpMV<SR>(A, x)
// (include/CombBLAS/ParFriends.h:1924-1996).  The reference moves x to the transposed process (TransposeVector), gathers
// it along the processor column, runs dcsc_gespmv on an id()-filled local y (SR::axpy, Friends.h:63-78) and reduces y along
// the processor row.  Same algorithm here: every process runs the local multiply on ITS tile on its GPU (cb_spmm_local with
// a one-column panel), the partial vectors are folded along the processor row and cut into FullyDistVec pieces; the
// exchanges of vector pieces are host-side.  Starting from id() matters for one semiring: SelectMax<bool,T> yields
// max(-1, x) here, not x, for x < -1.
template <typename IU, typename NUM, typename NUV, typename UDER>
bool CheckSpMVCompliance(const SpParMat<IU, NUM, UDER>& A, const FullyDistVec<IU, NUV>& x) {         // ParFriends.h:1350-1366
    if (*(A.getcommgrid()) != *(x.getcommgrid())) {
        SpParHelper::Print("Grids are not comparable for SpMV\n");
        MPI_Abort(MPI_COMM_WORLD, GRIDMISMATCH);
        return false;
    }
    const IU ncol = A.getncol(), len = x.TotalLength();
    if (ncol != len) {
        std::ostringstream outs;
        outs << "Can not multiply, dimensions does not match" << std::endl << ncol << " != " << len << std::endl;
        SpParHelper::Print(outs.str());
        MPI_Abort(MPI_COMM_WORLD, DIMMISMATCH);
        return false;
    }
    return true;
}

template <typename SR, typename IU, typename NUM, typename NUV, typename UDER>
FullyDistVec<IU, typename promote_trait<NUM, NUV>::T_promote> SpMV(const SpParMat<IU, NUM, UDER>& A, const FullyDistVec<IU, NUV>& x) {
    typedef typename promote_trait<NUM, NUV>::T_promote T_promote;
    typedef typename cb_storage<T_promote>::type ST;
    static_assert(semiring_traits<SR>::supported, "this semiring / type combination is not implemented by the B200 engine");
    static_assert(std::is_same<T_promote, NUV>::value, "the vector must already have the promoted type");
    CheckSpMVCompliance(A, x);
    std::shared_ptr<CommGrid> grid = A.getcommgrid();
    cb_ctx* ctx = grid->GetContext();
    const int pr = grid->GetGridRows(), pc = grid->GetGridCols(), myrow = grid->GetRankInProcCol(), mycol = grid->GetRankInProcRow();
    const IU gm = A.getnrow(), gn = A.getncol();
    // the reference's own 2D algorithm (ParFriends.h:1924-1996) with the vector exchange on the devices (cb_spmv_grid): my
    // piece of x goes up, the pieces are gathered between the GPUs, every process multiplies ITS tile with the part of x that
    // matches its columns, the partial results are combined along the processor row with SR's reduction, my piece of y comes down
    (void)pr; (void)pc; (void)myrow; (void)mycol;
    FullyDistVec<IU, T_promote> y(grid, gm, SR::id());
    static_assert(sizeof(ST) == sizeof(T_promote) || std::is_same<T_promote, bool>::value, "vector elements are stored as they travel");
    std::vector<ST> xin((size_t)x.MyLocLength()), yout((size_t)y.MyLocLength());
    for (size_t q = 0; q < xin.size(); ++q) xin[q] = (ST)x.GetLocArr()[q];
    cb_check(cb_spmv_grid(ctx, A.DeviceTile(), xin.data(), x.LengthUntil(), x.MyLocLength(), yout.data(), y.LengthUntil(), y.MyLocLength(),
                          semiring_traits<SR>::op, cb_dtype_of<NUV>::value, gm, gn), ctx, "cb_spmv_grid");
    for (size_t q = 0; q < yout.size(); ++q) y.SetLocalElement((IU)q, (T_promote)yout[q]);
    return y;
}

// ---------------------------------------------------------------------------------------------------------------
// Mult_AnXBn_Synch / PSpGEMM with a SPARSE tall-skinny right-hand side - the call Applications/SpMMError.cpp:83 and
// ReleaseTests/MultTest.cpp:162 make (reference include/CombBLAS/ParFriends.h:1004-1108, SpParMat.h:454-467).
// Lowered onto the dense engine: B's tiles are expanded into dense panels, the values come from SpMM under SR and the
// nonzero STRUCTURE of C from a second SpMM of the two patterns under the boolean semiring, so C has an entry exactly
// where the reference's sparse accumulator would create one (also when the folded value equals SR::id()).
// Meant for tall-skinny B (the dense panels are m x k per block-row); cost is nnz(A)*k regardless of nnz(B).
template <typename IU, typename NU1, typename NU2, typename UDERA, typename UDERB>
bool CheckSpGEMMCompliance(const SpParMat<IU, NU1, UDERA>& A, const SpParMat<IU, NU2, UDERB>& B) {     // ParFriends.h:160-181
    if (A.getncol() != B.getnrow()) {
        std::ostringstream outs;
        outs << "Can not multiply, dimensions does not match" << std::endl << A.getncol() << " != " << B.getnrow() << std::endl;
        SpParHelper::Print(outs.str());
        MPI_Abort(MPI_COMM_WORLD, DIMMISMATCH);
        return false;
    }
    if ((void*)&A == (void*)&B) {
        SpParHelper::Print("Can not multiply, inputs alias (make a temporary copy of one of them first)\n");
        MPI_Abort(MPI_COMM_WORLD, MATRIXALIAS);
        return false;
    }
    return true;
}

// The product on the device (cb_spgemm_summa, csrc/cb_spgemm.cu): per stage the tile pair is multiplied by expansion, all
// stages' partial products are merged by one stable sort + reduce-by-key - the device counterpart of LocalHybridSpGEMM
// (mtSpGEMM.h:213-460) + MultiwayMerge (MultiwayMerge.h:411-526).  Cost is what the product touches; an entry of C exists
// exactly where the reference creates one; products are folded in ascending inner index (mtSpGEMM.h:395-423).
// CB_SPGEMM_DENSE=1 selects the round-1 lowering onto dense panels instead (Mult_AnXBn_DensePanels below).
template <typename SR, typename NUO, typename UDERO, typename IU, typename NU1, typename NU2, typename UDERA, typename UDERB>
SpParMat<IU, NUO, UDERO> Mult_AnXBn_DensePanels(SpParMat<IU, NU1, UDERA>& A, SpParMat<IU, NU2, UDERB>& B, bool clearA = false, bool clearB = false);

template <typename SR, typename NUO, typename UDERO, typename IU, typename NU1, typename NU2, typename UDERA, typename UDERB>
SpParMat<IU, NUO, UDERO> Mult_AnXBn_Synch(SpParMat<IU, NU1, UDERA>& A, SpParMat<IU, NU2, UDERB>& B, bool clearA = false, bool clearB = false) {
    typedef typename promote_trait<NU1, NU2>::T_promote T_promote;
    static_assert(semiring_traits<SR>::supported, "this semiring / type combination is not implemented by the B200 engine");
    static_assert(std::is_same<T_promote, NUO>::value, "the output value type must be the promoted type of the operands");
    static const bool dense_panels = std::getenv("CB_SPGEMM_DENSE") && std::atoi(std::getenv("CB_SPGEMM_DENSE")) != 0;
    if (dense_panels) return Mult_AnXBn_DensePanels<SR, NUO, UDERO>(A, B, clearA, clearB);
    typedef typename UDERB::LocalIT LIT;
    typedef typename UDERO::LocalIT OIT;
    typedef typename cb_storage<T_promote>::type ST;
    if (!CheckSpGEMMCompliance(A, B)) return SpParMat<IU, NUO, UDERO>();
    int stages, dummy;
    std::shared_ptr<CommGrid> grid = ProductGrid(A.getcommgrid().get(), B.getcommgrid().get(), stages, dummy, dummy);
    cb_ctx* ctx = grid->GetContext();
    const IU gm = A.getnrow(), gn = A.getncol(), gk = B.getncol();
    const int dt = cb_dtype_of<T_promote>::value;
    // B on the device with values of the product's type: its own resident tile when the types agree, a converted copy otherwise
    cb_tile* tileB = nullptr;
    bool own_b = false;
    if (std::is_same<NU2, T_promote>::value) {
        tileB = B.DeviceTile();
    } else {
        const UDERB& bt = B.seq();
        SpTuples<LIT, NU2> t = TilesToTuples(bt);
        const int64_t nz = t.getnnz();
        std::vector<int64_t> rows((size_t)nz), cols((size_t)nz);
        std::vector<ST> vals((size_t)nz);
        for (int64_t p = 0; p < nz; ++p) { rows[(size_t)p] = t.rowindex(p); cols[(size_t)p] = t.colindex(p); vals[(size_t)p] = (ST)(T_promote)t.numvalue(p); }
        cb_check(cb_tile_upload_coo(ctx, bt.getnrow(), bt.getncol(), nz, rows.data(), cols.data(), vals.data(), CB_I64, dt, &tileB), ctx, "cb_tile_upload_coo");
        own_b = true;
    }
    cb_coo* C = nullptr;
    cb_check(cb_spgemm_summa(ctx, A.DeviceTile(), tileB, semiring_traits<SR>::op, dt, gm, gn, gk, &C), ctx, "cb_spgemm_summa");
    int64_t nnzc = 0, lm = 0, kl = 0;
    cb_check(cb_coo_info(C, &nnzc, &lm, &kl, nullptr), ctx, "cb_coo_info");
    std::vector<int64_t> ci((size_t)nnzc), cj((size_t)nnzc);
    std::vector<ST> cv((size_t)nnzc);
    cb_check(cb_coo_download(C, ci.data(), cj.data(), cv.data()), ctx, "cb_coo_download");
    cb_coo_free(C);
    if (own_b) cb_tile_free(tileB);
    // already sorted by column, then row: the order of every SpTuples the reference's local multiply returns
    SpTuples<OIT, NUO> ct(0, (OIT)lm, (OIT)kl);
    ct.tuples.reserve((size_t)nnzc);
    for (int64_t p = 0; p < nnzc; ++p) ct.tuples.emplace_back((OIT)ci[(size_t)p], (OIT)cj[(size_t)p], (NUO)cv[(size_t)p]);
    if (clearA) A.FreeDeviceTile();
    if (clearB) B.FreeDeviceTile();
    return SpParMat<IU, NUO, UDERO>(new UDERO(ct, false), grid, gm, gk);
}

// Round-1 formulation, kept for comparison (CB_SPGEMM_DENSE=1): lowered onto the dense engine.
template <typename SR, typename NUO, typename UDERO, typename IU, typename NU1, typename NU2, typename UDERA, typename UDERB>
SpParMat<IU, NUO, UDERO> Mult_AnXBn_DensePanels(SpParMat<IU, NU1, UDERA>& A, SpParMat<IU, NU2, UDERB>& B, bool clearA, bool clearB) {
    typedef typename promote_trait<NU1, NU2>::T_promote T_promote;
    static_assert(semiring_traits<SR>::supported, "this semiring / type combination is not implemented by the B200 engine");
    static_assert(std::is_same<T_promote, NUO>::value, "the output value type must be the promoted type of the operands");
    typedef typename UDERB::LocalIT LIT;
    if (!CheckSpGEMMCompliance(A, B)) return SpParMat<IU, NUO, UDERO>();
    int stages, dummy;
    std::shared_ptr<CommGrid> grid = ProductGrid(A.getcommgrid().get(), B.getcommgrid().get(), stages, dummy, dummy);
    const IU gm = A.getnrow(), gn = A.getncol(), gk = B.getncol();
    // dense images of my tile of B: values (promoted type) and structure
    const UDERB& bt = B.seq();
    const IU xl = bt.getnrow(), kl = bt.getncol(), lm = A.getlocalrows();
    {
        // the lowering holds two dense lm x kl and two xl x kl panels per process: say so instead of dying in an allocation
        static const double limit = std::getenv("CB_SPGEMM_DENSE_LIMIT_GB") ? std::atof(std::getenv("CB_SPGEMM_DENSE_LIMIT_GB")) * 1e9 : 32e9;
        const double need = ((double)lm + (double)xl) * (double)kl * (double)(sizeof(T_promote) + 1);
        if (need > limit) {
            std::ostringstream outs;
            outs << "Mult_AnXBn_Synch / PSpGEMM of this build multiplies through dense panels and is meant for a tall-skinny right-hand side: "
                 << "local panels of " << lm << " x " << kl << " and " << xl << " x " << kl << " would take " << need / 1e9
                 << " GB (limit " << limit / 1e9 << " GB, CB_SPGEMM_DENSE_LIMIT_GB)" << std::endl;
            SpParHelper::Print(outs.str());
            MPI_Abort(MPI_COMM_WORLD, INVALIDPARAMS);
        }
    }
    // absent entries of B hold SR::id(): it annihilates under every supported multiply (0*a, inf_plus(a, max), a AND false,
    // select2nd -> id), so they cannot change a folded value; which entries of C exist is decided by the structure product
    DenseParMat<IU, T_promote> Xv(SR::id(), grid, xl, kl);
    DenseParMat<IU, bool> Xs(false, grid, xl, kl);
    {
        SpTuples<LIT, NU2> bt_tuples = TilesToTuples(bt);
        for (int64_t p = 0; p < bt_tuples.getnnz(); ++p) {
            Xv(bt_tuples.rowindex(p), bt_tuples.colindex(p)) = (typename cb_storage<T_promote>::type)bt_tuples.numvalue(p);
            Xs(bt_tuples.rowindex(p), bt_tuples.colindex(p)) = 1;
        }
    }
    cb_ctx* ctx = grid->GetContext();
    cb_dense *dX = nullptr, *dY = nullptr, *dS = nullptr, *dM = nullptr;
    const int dt = cb_dtype_of<T_promote>::value;
    cb_check(cb_dense_alloc(ctx, xl, kl, dt, &dX), ctx, "cb_dense_alloc");
    cb_check(cb_dense_alloc(ctx, lm, kl, dt, &dY), ctx, "cb_dense_alloc");
    cb_check(cb_dense_alloc(ctx, xl, kl, CB_U8, &dS), ctx, "cb_dense_alloc");
    cb_check(cb_dense_alloc(ctx, lm, kl, CB_U8, &dM), ctx, "cb_dense_alloc");
    std::vector<typename cb_storage<T_promote>::type> Y((size_t)lm * (size_t)kl);
    std::vector<uint8_t> M((size_t)lm * (size_t)kl);
    if (kl > 0) {
        cb_check(cb_dense_upload(dX, Xv.data(), kl), ctx, "cb_dense_upload");
        cb_check(cb_dense_upload(dS, Xs.data(), kl), ctx, "cb_dense_upload");
    }
    // Sparse-aware option (CB_SPGEMM_FILTER=1): a column of A that meets only empty rows of B cannot contribute, and the
    // reference's column-by-column SpGEMM never touches it (mtSpGEMM.h:292-441).  When fewer than half of B's rows hold
    // anything, multiply with the tile that keeps only the other columns: the engine then costs what the product touches.
    cb_tile *tileA = A.DeviceTile(), *tileP = nullptr, *filtered = nullptr, *filtered_pattern = nullptr;
    static const bool want_filter = std::getenv("CB_SPGEMM_FILTER") && std::atoi(std::getenv("CB_SPGEMM_FILTER")) != 0;
    if (want_filter) {
        std::vector<uint8_t> mine((size_t)xl, 0);                  // rows of my B tile that hold anything
        for (IU i = 0; i < xl; ++i)
            for (IU j = 0; j < kl && !mine[(size_t)i]; ++j) mine[(size_t)i] = Xs(i, j) ? 1 : 0;
        std::vector<std::vector<char>> all;
        cb_host_allgatherv(mine.data(), mine.size(), all);
        std::vector<uint8_t> active((size_t)gn, 0);                // B's row block i lives on the processes of processor row i
        for (int i = 0; i < grid->GetGridRows(); ++i) {
            IU r0, rl;
            DenseParMat<IU, bool>::Block(gn, grid->GetGridRows(), i, r0, rl);
            for (int j = 0; j < grid->GetGridCols(); ++j) {
                const std::vector<char>& b = all[(size_t)grid->GetRank(i, j)];
                for (size_t q = 0; q < b.size(); ++q) active[(size_t)r0 + q] |= (uint8_t)b[q];
            }
        }
        size_t nactive = 0;
        for (uint8_t a : active) nactive += a;
        if (2 * nactive < (size_t)gn) {                            // the same decision on every process
            IU c0, cl;
            DenseParMat<IU, bool>::Block(gn, grid->GetGridCols(), grid->GetRankInProcRow(), c0, cl);
            cb_check(cb_tile_filter_columns(ctx, tileA, active.data() + c0, &filtered), ctx, "cb_tile_filter_columns");
            cb_check(cb_tile_pattern_view(filtered, &filtered_pattern), ctx, "cb_tile_pattern_view");
            tileA = filtered;
            tileP = filtered_pattern;
        }
    }
    if (!tileP) tileP = A.DevicePatternTile();
    cb_check(cb_spmm_summa(ctx, tileA, dX, dY, semiring_traits<SR>::op, gm, gn, gk), ctx, "cb_spmm_summa");
    cb_check(cb_spmm_summa(ctx, tileP, dS, dM, CB_OR_AND, gm, gn, gk), ctx, "cb_spmm_summa (structure)");
    if (kl > 0) {
        cb_check(cb_dense_download(dY, Y.data(), kl), ctx, "cb_dense_download");
        cb_check(cb_dense_download(dM, M.data(), kl), ctx, "cb_dense_download");
    }
    cb_check(cb_ctx_sync(ctx), ctx, "cb_ctx_sync");
    cb_dense_free(dX); cb_dense_free(dY); cb_dense_free(dS); cb_dense_free(dM);
    if (filtered_pattern) cb_tile_free(filtered_pattern);
    if (filtered) cb_tile_free(filtered);
    // C tile: one entry per structural hit, column-major like every SpTuples the reference's local multiply returns
    typedef typename UDERO::LocalIT OIT;
    SpTuples<OIT, NUO> ct(0, (OIT)lm, (OIT)kl);
    for (IU j = 0; j < kl; ++j)
        for (IU i = 0; i < lm; ++i)
            if (M[(size_t)i * (size_t)kl + (size_t)j]) ct.tuples.emplace_back((OIT)i, (OIT)j, (NUO)Y[(size_t)i * (size_t)kl + (size_t)j]);
    if (clearA) A.FreeDeviceTile();
    if (clearB) B.FreeDeviceTile();
    return SpParMat<IU, NUO, UDERO>(new UDERO(ct, false), grid, gm, gk);
}

// Mult_AnXBn_DoubleBuff (ParFriends.h:798-997) and Mult_AnXBn_Overlap (:1110-1235) differ from _Synch only in how the
// reference hides its broadcasts; here double buffering and overlap live inside cb_spmm_summa, so they are the same call.
template <typename SR, typename NUO, typename UDERO, typename IU, typename NU1, typename NU2, typename UDERA, typename UDERB>
SpParMat<IU, NUO, UDERO> Mult_AnXBn_DoubleBuff(SpParMat<IU, NU1, UDERA>& A, SpParMat<IU, NU2, UDERB>& B, bool clearA = false, bool clearB = false) {
    return Mult_AnXBn_Synch<SR, NUO, UDERO>(A, B, clearA, clearB);
}
template <typename SR, typename NUO, typename UDERO, typename IU, typename NU1, typename NU2, typename UDERA, typename UDERB>
SpParMat<IU, NUO, UDERO> Mult_AnXBn_Overlap(SpParMat<IU, NU1, UDERA>& A, SpParMat<IU, NU2, UDERB>& B, bool clearA = false, bool clearB = false) {
    return Mult_AnXBn_Synch<SR, NUO, UDERO>(A, B, clearA, clearB);
}

template <typename SR, typename IU, typename NU1, typename NU2, typename UDERA, typename UDERB>
SpParMat<IU, typename promote_trait<NU1, NU2>::T_promote, typename create_trait<UDERA, typename UDERA::LocalIT, typename promote_trait<NU1, NU2>::T_promote>::T_inferred>
PSpGEMM(SpParMat<IU, NU1, UDERA>& A, SpParMat<IU, NU2, UDERB>& B, bool clearA = false, bool clearB = false) {        // SpParMat.h:454-467
    typedef typename promote_trait<NU1, NU2>::T_promote N_promote;
    typedef typename create_trait<UDERA, typename UDERA::LocalIT, N_promote>::T_inferred DER_promote;
    return Mult_AnXBn_Synch<SR, N_promote, DER_promote>(A, B, clearA, clearB);
}

}  // namespace combblas
#endif
