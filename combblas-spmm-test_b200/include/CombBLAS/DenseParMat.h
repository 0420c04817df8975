// DenseParMat<IT,NT>: the dense operand / result of SpMM, distributed like the sparse matrix.
//
// Keeps the reference's constructor shape DenseParMat(NT value, grid, local rows, local cols)
// (include/CombBLAS/DenseParMat.h:49-128) but stores the local block as ONE contiguous row-major array instead of
// NT** with a new[] per row (SpHelper.h:241-247), and copies in the right direction (the reference's copy
// constructor and operator= copy from the uninitialised destination, DenseParMat.h:78, DenseParMat.cpp:163).
// Global shape: rows follow the block-row rule of the grid rows, columns the block rule of the grid columns -
// i.e. X(n x k) lives on rank (i,j) as rows block i of n, columns block j of k.
#ifndef CB_DENSEPARMAT_H
#define CB_DENSEPARMAT_H

#include <memory>
#include <vector>
#include "CommGrid.h"
#include "promote.h"

namespace combblas {

template <class IT, class NT>
class DenseParMat {
public:
    typedef typename cb_storage<NT>::type ST;
    DenseParMat() : commGrid(new CommGrid(MPI_COMM_WORLD, 0, 0)), m(0), n(0) {}
    DenseParMat(NT value, std::shared_ptr<CommGrid> grid, IT rows, IT cols)
        : commGrid(grid), m(rows), n(cols), array((size_t)rows * (size_t)cols, (ST)value) {}
    // global-shape constructor: every rank gets its block of a (grows x gcols) matrix filled with `value`
    static DenseParMat Global(NT value, std::shared_ptr<CommGrid> grid, IT grows, IT gcols) {
        IT s, lr, lc;
        Block(grows, grid->GetGridRows(), grid->GetRankInProcCol(), s, lr);
        Block(gcols, grid->GetGridCols(), grid->GetRankInProcRow(), s, lc);
        return DenseParMat(value, grid, lr, lc);
    }
    std::shared_ptr<CommGrid> getcommgrid() const { return commGrid; }
    IT grows() const { return (IT)commGrid->SumCol(m); }        // DenseParMat.h:101-106
    IT gcols() const { return (IT)commGrid->SumRow(n); }        // DenseParMat.h:107-113
    IT getlocalrows() const { return m; }
    IT getlocalcols() const { return n; }
    ST& operator()(IT i, IT j) { return array[(size_t)i * (size_t)n + (size_t)j]; }
    const ST& operator()(IT i, IT j) const { return array[(size_t)i * (size_t)n + (size_t)j]; }
    ST* data() { return array.data(); }
    const ST* data() const { return array.data(); }
    // first global row / column of the local block
    void GetPlaceInGlobalGrid(IT grows_, IT gcols_, IT& rowOffset, IT& colOffset) const {
        IT l;
        Block(grows_, commGrid->GetGridRows(), commGrid->GetRankInProcCol(), rowOffset, l);
        Block(gcols_, commGrid->GetGridCols(), commGrid->GetRankInProcRow(), colOffset, l);
    }
    static void Block(IT total, int nb, int b, IT& start, IT& len) {
        const IT per = total / nb;
        start = (IT)b * per;
        len = (b == nb - 1) ? total - start : per;
    }

private:
    std::shared_ptr<CommGrid> commGrid;
    IT m, n;                  // local rows and columns
    std::vector<ST> array;    // row-major, leading dimension n
};

}  // namespace combblas
#endif
