// CommGrid: the pr x pc logical process grid, one process per B200.
//
// Same public surface as the reference class (include/CombBLAS/CommGrid.h:44-166, src/CommGrid.cpp:37-180):
// row-major rank -> (myprocrow = rank / grcols, myproccol = rank % grcols); GetRankInProcRow() is the COLUMN
// index and GetRankInProcCol() the ROW index (CommGrid.h:109-110); nrowproc == ncolproc == 0 asks for a square
// grid and aborts with NOTSQUARE when the rank count is not a perfect square (src/CommGrid.cpp:44-53).
// Underneath, instead of MPI_Comm_split communicators, the grid owns a cb_ctx: one CUDA device with its compute /
// communication streams and the NCCL world / row / column communicators.
#ifndef CB_COMMGRID_H
#define CB_COMMGRID_H

#include <cmath>
#include <memory>
#include "SpDefs.h"

namespace combblas {

class CommGrid {
public:
    CommGrid(MPI_Comm world, int nrowproc, int ncolproc) : commWorld(world) {
        int nproc = 1;
        MPI_Comm_rank(world, &myrank);
        MPI_Comm_size(world, &nproc);
        if (nrowproc == 0 && ncolproc == 0) {
            nrowproc = ncolproc = (int)std::lround(std::sqrt((double)nproc));
            if (nrowproc * ncolproc != nproc) {
                SpParHelper::Print("This version of the Combinatorial BLAS only works on a square logical processor grid\n");
                MPI_Abort(world, NOTSQUARE);
            }
        }
        if (nrowproc * ncolproc != nproc) {
            SpParHelper::Print("COMBBLAS: the processor grid does not match the number of processes\n");
            MPI_Abort(world, INVALIDPARAMS);
        }
        grrows = nrowproc;
        grcols = ncolproc;
#ifndef CB_HAVE_MPI
        cb_rt::grid_cols() = grcols;
#endif
        myproccol = myrank % grcols;
        myprocrow = myrank / grcols;
        unsigned char uid[128] = {0};
#ifndef CB_HAVE_MPI
        if (nproc > 1) {
            if (myrank == 0) cb_check(cb_comm_unique_id(uid), nullptr, "cb_comm_unique_id");
            cb_rt::bcast_bytes(uid, sizeof uid, 0);
        }
        const int device = cb_rt::local_rank();
#else
        if (nproc > 1) {
            if (myrank == 0) cb_check(cb_comm_unique_id(uid), nullptr, "cb_comm_unique_id");
            MPI_Bcast(uid, (int)sizeof uid, MPI_BYTE, 0, world);
        }
        int ndev = 1;
        cb_device_count(&ndev);
        const int device = myrank % (ndev > 0 ? ndev : 1);
        // processor-row and processor-column communicators exactly as the reference builds them (src/CommGrid.cpp:66-67):
        // user code reduces / broadcasts over GetRowWorld() / GetColWorld()
        MPI_Comm row = MPI_COMM_NULL, col = MPI_COMM_NULL;
        MPI_Comm_split(world, myprocrow, myrank, &row);
        MPI_Comm_split(world, myproccol, myrank, &col);
        rowWorld.reset(new MPI_Comm(row), [](MPI_Comm* c) { int fin = 0; MPI_Finalized(&fin); if (!fin && *c != MPI_COMM_NULL) MPI_Comm_free(c); delete c; });
        colWorld.reset(new MPI_Comm(col), [](MPI_Comm* c) { int fin = 0; MPI_Finalized(&fin); if (!fin && *c != MPI_COMM_NULL) MPI_Comm_free(c); delete c; });
#endif
        cb_ctx* c = nullptr;
        cb_check(cb_ctx_create_grid(device, myrank, nproc, grrows, grcols, nproc > 1 ? uid : nullptr, &c), nullptr, "cb_ctx_create_grid");
        ctx.reset(c, [](cb_ctx* p) { cb_ctx_destroy(p); });
    }
    // copies share the device context (the reference duplicates its communicators, CommGrid.h:60-71)
    CommGrid(const CommGrid&) = default;
    CommGrid& operator=(const CommGrid&) = default;

    bool operator==(const CommGrid& rhs) const {              // src/CommGrid.cpp:139-150
        return grrows == rhs.grrows && grcols == rhs.grcols && myprocrow == rhs.myprocrow && myproccol == rhs.myproccol &&
               myrank == rhs.myrank;
    }
    bool operator!=(const CommGrid& rhs) const { return !(*this == rhs); }
    bool OnSameProcCol(int rhsrank) const { return myproccol == rhsrank % grcols; }
    bool OnSameProcRow(int rhsrank) const { return myprocrow == rhsrank / grcols; }

    int GetRank(int rowrank, int colrank) const { return rowrank * grcols + colrank; }
    int GetRank(int diagrank) const { return diagrank * grcols + diagrank; }
    int GetRank() const { return myrank; }
    int GetRankInProcRow() const { return myproccol; }
    int GetRankInProcCol() const { return myprocrow; }
    int GetRankInProcRow(int wholerank) const { return wholerank % grcols; }
    int GetRankInProcCol(int wholerank) const { return wholerank / grcols; }
    int GetComplementRank() const { return grcols * myproccol + myprocrow; }
    int GetGridRows() const { return grrows; }
    int GetGridCols() const { return grcols; }
    int GetSize() const { return grrows * grcols; }
    MPI_Comm GetWorld() const { return commWorld; }
#ifndef CB_HAVE_MPI
    MPI_Comm GetRowWorld() const { return 1000 + myprocrow; }     // handles of cb_mpi.h: processes with the same myprocrow /
    MPI_Comm GetColWorld() const { return 2000 + myproccol; }     // myproccol; the device collectives live inside GetContext()
#else
    MPI_Comm GetRowWorld() const { return *rowWorld; }            // MPI_Comm_split(world, myprocrow, rank), src/CommGrid.cpp:66
    MPI_Comm GetColWorld() const { return *colWorld; }            // MPI_Comm_split(world, myproccol, rank), src/CommGrid.cpp:67
#endif

    cb_ctx* GetContext() const { return ctx.get(); }

    // sums over the world / my processor row / my processor column (MPI_Allreduce in the reference)
    int64_t SumWorld(int64_t v) const { cb_check(cb_comm_allreduce_i64(ctx.get(), 0, 0, &v, 1), ctx.get(), "allreduce"); return v; }
    int64_t SumRow(int64_t v) const { cb_check(cb_comm_allreduce_i64(ctx.get(), 1, 0, &v, 1), ctx.get(), "allreduce"); return v; }
    int64_t SumCol(int64_t v) const { cb_check(cb_comm_allreduce_i64(ctx.get(), 2, 0, &v, 1), ctx.get(), "allreduce"); return v; }
    int64_t MaxWorld(int64_t v) const { cb_check(cb_comm_allreduce_i64(ctx.get(), 0, 1, &v, 1), ctx.get(), "allreduce"); return v; }
    int64_t MinWorld(int64_t v) const { cb_check(cb_comm_allreduce_i64(ctx.get(), 0, 2, &v, 1), ctx.get(), "allreduce"); return v; }

private:
    MPI_Comm commWorld;
#ifdef CB_HAVE_MPI
    std::shared_ptr<MPI_Comm> rowWorld, colWorld;                 // shared by the copies of a grid, freed with the last one
#endif
    int grrows = 1, grcols = 1, myprocrow = 0, myproccol = 0, myrank = 0;
    std::shared_ptr<cb_ctx> ctx;
};

// ProductGrid (src/CommGrid.cpp:164-180): the grids must be the same; unlike the reference any pr x pc is accepted
inline std::shared_ptr<CommGrid> ProductGrid(CommGrid* gridA, CommGrid* gridB, int& innerdim, int& Aoffset, int& Boffset) {
    if (*gridA != *gridB) {
        SpParHelper::Print("Grids don't confirm for multiplication\n");
        MPI_Abort(MPI_COMM_WORLD, GRIDMISMATCH);
    }
    innerdim = gridA->GetGridCols();
    Aoffset = Boffset = 0;
    return std::make_shared<CommGrid>(*gridA);
}

}  // namespace combblas
#endif
