// TEST ONLY: the host layer built with -DCB_HAVE_MPI (a real <mpi.h> instead of the launcher-environment shim) on a process
// grid.  The MPI here is oracle/mpi_multi (process-per-rank stand-in), the C ABI is tests/mock_abi.  Checks that
// GetRowWorld() / GetColWorld() are real sub-communicators of a processor row / column (src/CommGrid.cpp:66-67 of the
// reference): reductions over them must see only that row / column.
#include <mpi.h>
#include <cstdio>
#include <memory>
#include "CombBLAS/CombBLAS.h"
using namespace combblas;

int main(int argc, char** argv) {
    MPI_Init(&argc, &argv);
    int rank = 0, np = 1;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &np);
    const int pr = argc > 1 ? atoi(argv[1]) : 0, pc = argc > 2 ? atoi(argv[2]) : 0;
    int bad = 0;
    {
        std::shared_ptr<CommGrid> g(new CommGrid(MPI_COMM_WORLD, pr, pc));
        const int R = g->GetGridRows(), C = g->GetGridCols(), i = g->GetRankInProcCol(), j = g->GetRankInProcRow();
        if (R * C != np || rank != i * C + j) ++bad;
        int rs = 0, cs = 0, rn = 0, cn = 0, rr = -1, cr = -1;
        MPI_Comm_size(g->GetRowWorld(), &rn);
        MPI_Comm_size(g->GetColWorld(), &cn);
        MPI_Comm_rank(g->GetRowWorld(), &rr);
        MPI_Comm_rank(g->GetColWorld(), &cr);
        if (rn != C || cn != R || rr != j || cr != i) ++bad;                     // a row world has one member per grid column
        int me = rank + 1;
        MPI_Allreduce(&me, &rs, 1, MPI_INT, MPI_SUM, g->GetRowWorld());
        MPI_Allreduce(&me, &cs, 1, MPI_INT, MPI_SUM, g->GetColWorld());
        int want_r = 0, want_c = 0;
        for (int q = 0; q < np; ++q) { if (q / C == i) want_r += q + 1; if (q % C == j) want_c += q + 1; }
        if (rs != want_r || cs != want_c) ++bad;
        int root_val = rank;                                                   // broadcast along my processor row from its first member
        MPI_Bcast(&root_val, 1, MPI_INT, 0, g->GetRowWorld());
        if (root_val != i * C) ++bad;
        CommGrid copy(*g);                                                     // copies share the communicators
        int rs2 = 0;
        MPI_Allreduce(&me, &rs2, 1, MPI_INT, MPI_SUM, copy.GetRowWorld());
        if (rs2 != want_r) ++bad;
        // a distributed vector on this grid exercises the host-side allgather of the CB_HAVE_MPI branch
        FullyDistVec<int64_t, int64_t> v(g, 37, 0);
        v.iota(37, 1);
        const int64_t total = v.Reduce(std::plus<int64_t>(), (int64_t)0);
        if (total != 37 * 38 / 2) ++bad;
    }
    int all_bad = 0;
    MPI_Allreduce(&bad, &all_bad, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
    if (rank == 0) printf(all_bad ? "CB_HAVE_MPI grid FAILED (%d)\n" : "CB_HAVE_MPI grid working correctly\n", all_bad);
    MPI_Finalize();
    return all_bad ? 1 : 0;
}
