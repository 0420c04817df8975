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
#include "FullyDistVec.h"
#include "promote.h"

namespace combblas {

template <class IU, class NU, class DER>
class SpParMat;

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
    // Fold along a dimension into a distributed vector (reference DenseParMat.cpp:36-130): dim == Row folds every row over
    // its columns (result of length grows()), dim == Column folds every column over its rows (length gcols()).  Local
    // fold first, then the partial vectors of the processes that share the rows (columns) in grid order, as the
    // reference's MPI_Reduce_scatter does; the result is laid out as a FullyDistVec.  Host-side, like the reference.
    template <typename _BinaryOperation>
    FullyDistVec<IT, NT> Reduce(Dim dim, _BinaryOperation op, NT identity) const {
        const int pr = commGrid->GetGridRows(), pc = commGrid->GetGridCols();
        std::vector<ST> part(dim == Row ? (size_t)m : (size_t)n, (ST)identity);
        for (IT i = 0; i < m; ++i)
            for (IT j = 0; j < n; ++j) {
                ST& dst = part[dim == Row ? (size_t)i : (size_t)j];
                dst = (ST)op((NT)dst, (NT)(*this)(i, j));
            }
        std::vector<std::vector<char>> all;
        cb_host_allgatherv(part.data(), part.size() * sizeof(ST), all);
        const IT glen = dim == Row ? grows() : gcols();
        std::vector<NT> whole((size_t)glen, identity);
        IT off = 0;
        for (int b = 0; b < (dim == Row ? pr : pc); ++b) {              // block b of the result = fold over the processes holding it
            IT blen = 0;
            for (int q = 0; q < (dim == Row ? pc : pr); ++q) {
                const std::vector<char>& buf = all[(size_t)(dim == Row ? commGrid->GetRank(b, q) : commGrid->GetRank(q, b))];
                blen = (IT)(buf.size() / sizeof(ST));
                const ST* v = reinterpret_cast<const ST*>(buf.data());
                for (IT i = 0; i < blen; ++i) whole[(size_t)(off + i)] = op(whole[(size_t)(off + i)], (NT)v[i]);
            }
            off += blen;
        }
        FullyDistVec<IT, NT> out(commGrid);
        out.Scatter(whole);
        return out;
    }
    // add a sparse matrix with the same distribution (reference DenseParMat.cpp:132-146, Dcsc::UpdateDense)
    template <typename DER>
    DenseParMat<IT, NT>& operator+=(const SpParMat<IT, NT, DER>& rhs);

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
