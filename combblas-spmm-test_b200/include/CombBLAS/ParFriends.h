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

#include "DenseParMat.h"
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
    cb_dense *dX = nullptr, *dY = nullptr;
    const int dt = cb_dtype_of<NUV>::value;
    cb_check(cb_dense_alloc(ctx, X.getlocalrows(), kl, dt, &dX), ctx, "cb_dense_alloc");
    cb_check(cb_dense_alloc(ctx, lm, kl, dt, &dY), ctx, "cb_dense_alloc");
    if (kl > 0) cb_check(cb_dense_upload(dX, X.data(), kl), ctx, "cb_dense_upload");
    cb_check(cb_spmm_summa(ctx, tile, dX, dY, semiring_traits<SR>::op, gm, gn, gk), ctx, "cb_spmm_summa");
    if (kl > 0) cb_check(cb_dense_download(dY, Y.data(), kl), ctx, "cb_dense_download");
    cb_check(cb_ctx_sync(ctx), ctx, "cb_ctx_sync");
    cb_dense_free(dX);
    cb_dense_free(dY);
    return Y;
}

}  // namespace combblas
#endif
