// FullyDistVec<IT,NT>: a dense vector distributed over ALL processes of the grid - the operand and result type of the
// reference's dense SpMV<SR>(A, x) (include/CombBLAS/ParFriends.h:1924-1996) and of DenseParMat::Reduce
// (DenseParMat.cpp:36-130).  SURVEY.md section 8 row f3.
//
// Distribution is the reference's two-level rule (include/CombBLAS/FullyDist.h:107-260): the global length is cut over the
// processor rows (floor division, the last row takes the remainder), each row's share is cut over its processor columns
// the same way; process (i,j) holds piece j of row-share i, so the pieces of processor row i together are exactly the
// row block i of a matrix dimension of that length.  FullyDistLayout holds that arithmetic as plain functions (tested on
// CPU against the reference's own multi-process runs); the vector itself lives in host memory like the reference's
// std::vector<NT> arr, and element-wise operations run on the host as they do there.  The multiply does not: SpMV<SR>
// (ParFriends.h of this layer) runs on the GPUs.
#ifndef CB_FULLYDISTVEC_H
#define CB_FULLYDISTVEC_H

#include <algorithm>
#include <cstring>
#include <fstream>
#include <memory>
#include <numeric>
#include <vector>
#include "CommGrid.h"
#include "promote.h"

namespace combblas {

template <class IT>
struct FullyDistLayout {
    IT glen;
    int procrows, proccols;
    FullyDistLayout(IT len, int pr, int pc) : glen(len), procrows(pr), proccols(pc) {}
    IT PerProcRow() const { return glen / procrows; }
    IT RowLength(int procrow) const { return procrow == procrows - 1 ? glen - PerProcRow() * (procrows - 1) : PerProcRow(); }     // FullyDist.h:240-249
    IT RowLenUntil(int procrow) const { return PerProcRow() * procrow; }
    IT LocLength(int procrow, int proccol) const {                                                                                  // FullyDist.h:240-260
        const IT n_thisrow = RowLength(procrow), n_perproc = n_thisrow / proccols;
        return proccol == proccols - 1 ? n_thisrow - n_perproc * (proccols - 1) : n_perproc;
    }
    IT LengthUntil(int procrow, int proccol) const { return RowLenUntil(procrow) + (RowLength(procrow) / proccols) * proccol; }     // FullyDist.h:181-200
    // owner of global index gind as (procrow, proccol) and its local index (FullyDist.h:107-150)
    void Owner(IT gind, int& procrow, int& proccol, IT& lind) const {
        const IT n_perprocrow = PerProcRow();
        procrow = n_perprocrow != 0 ? std::min((int)(gind / n_perprocrow), procrows - 1) : procrows - 1;
        const IT ind_withinrow = gind - (IT)procrow * n_perprocrow;
        const IT n_perproc = RowLength(procrow) / proccols;
        proccol = n_perproc != 0 ? std::min((int)(ind_withinrow / n_perproc), proccols - 1) : proccols - 1;
        lind = ind_withinrow - (IT)proccol * n_perproc;
    }
};

// every process contributes a byte string; afterwards everyone holds all of them in rank order (host-side exchange of
// vector pieces; MPI_Allgatherv in the reference)
inline void cb_host_allgatherv(const void* mine, size_t bytes, std::vector<std::vector<char>>& all) {
    int p = 1, r = 0;
    MPI_Comm_size(MPI_COMM_WORLD, &p);
    MPI_Comm_rank(MPI_COMM_WORLD, &r);
    all.assign((size_t)p, std::vector<char>());
    if (p == 1) { if (bytes) all[0].assign((const char*)mine, (const char*)mine + bytes); return; }
#ifdef CB_HAVE_MPI
    std::vector<long long> sizes((size_t)p);
    long long mysz = (long long)bytes;
    MPI_Allgather(&mysz, 1, MPI_LONG_LONG, sizes.data(), 1, MPI_LONG_LONG, MPI_COMM_WORLD);
    std::vector<int> cnt((size_t)p), dsp((size_t)p);
    long long tot = 0;
    for (int q = 0; q < p; ++q) { cnt[q] = (int)sizes[q]; dsp[q] = (int)tot; tot += sizes[q]; }
    std::vector<char> buf((size_t)tot);
    MPI_Allgatherv(mine, (int)bytes, MPI_BYTE, buf.data(), cnt.data(), dsp.data(), MPI_BYTE, MPI_COMM_WORLD);
    for (int q = 0; q < p; ++q) all[q].assign(buf.begin() + dsp[q], buf.begin() + dsp[q] + cnt[q]);
#else
    uint64_t mysz = bytes;
    std::vector<char> sizes;
    cb_rt::allgather_bytes(&mysz, sizeof mysz, sizes);
    uint64_t mx = 0;
    std::vector<uint64_t> sz((size_t)p);
    for (int q = 0; q < p; ++q) { std::memcpy(&sz[q], sizes.data() + sizeof(uint64_t) * (size_t)q, sizeof(uint64_t)); mx = std::max(mx, sz[q]); }
    std::vector<char> padded((size_t)mx, 0), buf;
    if (bytes) std::memcpy(padded.data(), mine, bytes);
    if (mx == 0) return;                                   // nobody has anything
    cb_rt::allgather_bytes(padded.data(), (size_t)mx, buf);
    for (int q = 0; q < p; ++q) all[q].assign(buf.begin() + (size_t)mx * (size_t)q, buf.begin() + (size_t)mx * (size_t)q + (size_t)sz[q]);
#endif
}

template <class IT, class NT>
class FullyDistVec {
public:
    static_assert(!std::is_same<NT, bool>::value, "FullyDistVec<IT,bool> does not exist in the reference either (FullyDistVec.h:61)");
    FullyDistVec() : commGrid(new CommGrid(MPI_COMM_WORLD, 0, 0)), glen(0) {}
    explicit FullyDistVec(std::shared_ptr<CommGrid> grid) : commGrid(grid), glen(0) {}
    FullyDistVec(IT globallen, NT initval) : commGrid(new CommGrid(MPI_COMM_WORLD, 0, 0)), glen(globallen) { arr.assign((size_t)MyLocLength(), initval); }
    FullyDistVec(std::shared_ptr<CommGrid> grid, IT globallen, NT initval) : commGrid(grid), glen(globallen) { arr.assign((size_t)MyLocLength(), initval); }

    std::shared_ptr<CommGrid> getcommgrid() const { return commGrid; }
    FullyDistLayout<IT> Layout() const { return FullyDistLayout<IT>(glen, commGrid->GetGridRows(), commGrid->GetGridCols()); }
    IT TotalLength() const { return glen; }
    IT MyLocLength() const { return Layout().LocLength(commGrid->GetRankInProcCol(), commGrid->GetRankInProcRow()); }
    IT LengthUntil() const { return Layout().LengthUntil(commGrid->GetRankInProcCol(), commGrid->GetRankInProcRow()); }
    IT MyRowLength() const { return Layout().RowLength(commGrid->GetRankInProcCol()); }
    IT RowLenUntil() const { return Layout().RowLenUntil(commGrid->GetRankInProcCol()); }
    int Owner(IT gind, IT& lind) const {
        int pr, pc;
        Layout().Owner(gind, pr, pc, lind);
        return commGrid->GetRank(pr, pc);
    }
    IT LocArrSize() const { return (IT)arr.size(); }
    const NT* GetLocArr() const { return arr.data(); }
    NT* GetLocArr() { return arr.data(); }
    void SetLocalElement(IT index, NT value) { arr[(size_t)index] = value; }
    NT GetLocalElement(IT index) const { return arr[(size_t)index]; }

    // element-wise assignment: only the owner stores (FullyDistVec.cpp SetElement); every process may call it
    void SetElement(IT indx, NT numx) {
        IT lind;
        if (Owner(indx, lind) == commGrid->GetRank()) arr[(size_t)lind] = numx;
    }
    // collective read of one element: the owner's value reaches everyone (MPI_Bcast in the reference)
    NT GetElement(IT indx) const {
        IT lind;
        const int owner = Owner(indx, lind);
        NT v = NT();
        if (owner == commGrid->GetRank()) v = arr[(size_t)lind];
        std::vector<std::vector<char>> all;
        cb_host_allgatherv(&v, sizeof(NT), all);
        std::memcpy(&v, all[(size_t)owner].data(), sizeof(NT));
        return v;
    }
    void iota(IT globalsize, NT first) {                                   // FullyDistVec.cpp iota: first, first+1, ...
        glen = globalsize;
        const IT len = MyLocLength(), off = LengthUntil();
        arr.resize((size_t)len);
        for (IT i = 0; i < len; ++i) arr[(size_t)i] = first + (NT)(off + i);
    }
    template <typename _UnaryOperation>
    void Apply(_UnaryOperation f) { std::transform(arr.begin(), arr.end(), arr.begin(), f); }
    template <typename _BinaryOperation>
    void EWise(const FullyDistVec<IT, NT>& rhs, _BinaryOperation op) {     // FullyDistVec.cpp:313-322
        CheckSame(rhs);
        std::transform(arr.begin(), arr.end(), rhs.arr.begin(), arr.begin(), op);
    }
    FullyDistVec<IT, NT>& operator+=(const FullyDistVec<IT, NT>& rhs) { EWise(rhs, std::plus<NT>()); return *this; }
    FullyDistVec<IT, NT>& operator-=(const FullyDistVec<IT, NT>& rhs) { EWise(rhs, std::minus<NT>()); return *this; }
    // global reduction (FullyDistVec.cpp:157-167): local fold, then the pieces in rank order
    template <typename _BinaryOperation>
    NT Reduce(_BinaryOperation op, NT identity) const {
        NT loc = std::accumulate(arr.begin(), arr.end(), identity, op);
        std::vector<std::vector<char>> all;
        cb_host_allgatherv(&loc, sizeof(NT), all);
        NT tot = identity;
        for (const std::vector<char>& b : all) { NT v; std::memcpy(&v, b.data(), sizeof(NT)); tot = op(tot, v); }
        return tot;
    }
    template <typename _Predicate>
    IT Count(_Predicate pred) const {
        const IT loc = (IT)std::count_if(arr.begin(), arr.end(), pred);
        return (IT)commGrid->SumWorld((int64_t)loc);
    }
    bool operator==(const FullyDistVec<IT, NT>& rhs) const {               // FullyDistVec.cpp:369-377 with ErrorTolerantEqual (Compare.h:46-65)
        bool same = glen == rhs.glen && *commGrid == *rhs.commGrid && arr.size() == rhs.arr.size();
        for (size_t i = 0; same && i < arr.size(); ++i) {
            if (std::is_floating_point<NT>::value) {
                const double a = (double)arr[i], b = (double)rhs.arr[i], d = a > b ? a - b : b - a;
                same = d < EPSILON || d < EPSILON * std::max(a < 0 ? -a : a, b < 0 ? -b : b);
            } else {
                same = arr[i] == rhs.arr[i];
            }
        }
        return commGrid->MinWorld(same ? 1 : 0) == 1;
    }
    // "rows cols nnz" then one "row col value" line per entry, one-based; the vector index is the row for a column vector
    // (cols == 1) and the column otherwise; absent entries are NT() (reference FullyDistVec.cpp:495-502 through
    // FullyDistSpVec::ReadDistribute, FullyDistSpVec.cpp:1397-1500).  Every process reads the stream and keeps its piece.
    std::ifstream& ReadDistribute(std::ifstream& infile, int master) {
        (void)master;
        long long numrows = 0, numcols = 0, total_nnz = 0;
        std::vector<NT> whole;
        if (infile.is_open()) {
            infile.clear();
            infile.seekg(0);
            infile >> numrows >> numcols >> total_nnz;
            whole.assign((size_t)(numcols == 1 ? numrows : numcols), NT());
            long long r, c;
            double v;
            for (long long q = 0; q < total_nnz && (infile >> r >> c >> v); ++q) {
                const long long ind = (numcols == 1 ? r : c) - 1;
                if (ind >= 0 && ind < (long long)whole.size()) whole[(size_t)ind] = (NT)v;
            }
        }
        Scatter(whole);
        return infile;
    }
    // the whole vector as text written by `master`, in the format ReadDistribute reads: "length 1 length" and one
    // "index 1 value" line per element, one-based (reference FullyDistVec.cpp:504-510 through FullyDistSpVec::SaveGathered)
    void SaveGathered(std::ofstream& outfile, int master) const {
        const std::vector<NT> whole = Gather();
        if (commGrid->GetRank() == master && outfile.is_open()) {
            outfile.precision(17);
            outfile << glen << "\t" << 1 << "\t" << glen << "\n";
            for (size_t i = 0; i < whole.size(); ++i) outfile << i + 1 << "\t" << 1 << "\t" << +whole[i] << "\n";
            outfile.flush();
        }
    }
    // the whole vector on every process (the concatenation of the pieces in owner order)
    std::vector<NT> Gather() const {
        std::vector<std::vector<char>> all;
        cb_host_allgatherv(arr.data(), arr.size() * sizeof(NT), all);
        std::vector<NT> whole((size_t)glen);
        const FullyDistLayout<IT> L = Layout();
        for (int i = 0; i < L.procrows; ++i)
            for (int j = 0; j < L.proccols; ++j) {
                const std::vector<char>& b = all[(size_t)commGrid->GetRank(i, j)];
                if (!b.empty()) std::memcpy(whole.data() + L.LengthUntil(i, j), b.data(), b.size());
            }
        return whole;
    }
    // keep my piece of a vector every process holds in full
    void Scatter(const std::vector<NT>& whole) {
        glen = (IT)whole.size();
        const IT len = MyLocLength(), off = LengthUntil();
        arr.assign(whole.begin() + off, whole.begin() + off + len);
    }

private:
    void CheckSame(const FullyDistVec<IT, NT>& rhs) const {
        if (glen != rhs.glen || *commGrid != *rhs.commGrid) {
            SpParHelper::Print("Grids or lengths are not comparable for element-wise vector operations\n");
            MPI_Abort(MPI_COMM_WORLD, glen != rhs.glen ? DIMMISMATCH : GRIDMISMATCH);
        }
    }
    std::shared_ptr<CommGrid> commGrid;
    IT glen;
    std::vector<NT> arr;
};

}  // namespace combblas
#endif
