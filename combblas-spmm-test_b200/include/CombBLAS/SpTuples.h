// Local tile formats of the host layer: SpTuples (COO), SpDCCols (doubly compressed sparse columns), SpCCols (CSC).
//
// They keep the surface the distributed layer needs from a tile type DER (reference include/CombBLAS/SpMat.h:54-174):
// typedef LocalIT/LocalNT, static esscount, GetEssentials(), Create(ess), GetArrays(), getnrow/getncol/getnnz/isZero.
// Arrays are the reference's: DCSC cp[nzc+1], jc[nzc], ir[nz], numx[nz] (dcsc.h:124-131, essentials {nnz,m,n,nzc},
// SpDCCols.cpp:787-795); CSC jc[n+1], ir[nz], num[nz] (csc.h:71-75, essentials {nnz,m,n}).
// Storage is std::vector; bool values are one byte each.
#ifndef CB_SPTUPLES_H
#define CB_SPTUPLES_H

#include <algorithm>
#include <numeric>
#include <tuple>
#include <vector>
#include "SpDefs.h"
#include "promote.h"

namespace combblas {

// {address, count} pair and the bundle of them a tile exposes: the wire format (LocArr.h:36-60)
template <class T>
struct LocArr {
    LocArr() : addr(nullptr), count(0) {}
    LocArr(T* a, size_t c) : addr(a), count(c) {}
    T* addr;
    size_t count;
};
template <class IT, class NT>
struct Arr {
    Arr(size_t nind, size_t nnum) : indarrs(nind), numarrs(nnum) {}
    std::vector<LocArr<IT>> indarrs;
    std::vector<LocArr<typename cb_storage<NT>::type>> numarrs;
    size_t totalsize() const { return indarrs.size() + numarrs.size(); }
};

template <class IT, class NT>
class SpTuples {
public:
    typedef IT LocalIT;
    typedef NT LocalNT;
    typedef typename cb_storage<NT>::type ST;
    SpTuples() : m(0), n(0) {}
    SpTuples(int64_t size, IT nRow, IT nCol) : tuples((size_t)size), m(nRow), n(nCol) {}
    SpTuples(int64_t size, IT nRow, IT nCol, std::tuple<IT, IT, NT>* mytuples, bool sorted = false)
        : tuples(mytuples, mytuples + size), m(nRow), n(nCol) { if (!sorted) SortColBased(); }
    SpTuples(IT nRow, IT nCol, const std::vector<IT>& rows, const std::vector<IT>& cols, const std::vector<NT>& vals)
        : m(nRow), n(nCol) {
        tuples.reserve(rows.size());
        for (size_t i = 0; i < rows.size(); ++i) tuples.emplace_back(rows[i], cols[i], vals[i]);
        SortColBased();
    }
    IT& rowindex(IT i) { return std::get<0>(tuples[(size_t)i]); }
    IT& colindex(IT i) { return std::get<1>(tuples[(size_t)i]); }
    NT& numvalue(IT i) { return std::get<2>(tuples[(size_t)i]); }
    IT rowindex(IT i) const { return std::get<0>(tuples[(size_t)i]); }
    IT colindex(IT i) const { return std::get<1>(tuples[(size_t)i]); }
    NT numvalue(IT i) const { return std::get<2>(tuples[(size_t)i]); }
    IT getnrow() const { return m; }
    IT getncol() const { return n; }
    int64_t getnnz() const { return (int64_t)tuples.size(); }
    bool isZero() const { return tuples.empty(); }
    void SortColBased() {                                        // column-major, rows ascending inside a column
        std::sort(tuples.begin(), tuples.end(), [](const std::tuple<IT, IT, NT>& a, const std::tuple<IT, IT, NT>& b) {
            return std::get<1>(a) != std::get<1>(b) ? std::get<1>(a) < std::get<1>(b) : std::get<0>(a) < std::get<0>(b);
        });
    }
    template <typename BINFUNC>
    void RemoveDuplicates(BINFUNC BinOp) {                       // SpTuples.cpp:271: equal (row,col) merged with BinOp
        SortColBased();
        size_t w = 0;
        for (size_t r = 0; r < tuples.size(); ++r) {
            if (w && std::get<0>(tuples[w - 1]) == std::get<0>(tuples[r]) && std::get<1>(tuples[w - 1]) == std::get<1>(tuples[r]))
                std::get<2>(tuples[w - 1]) = BinOp(std::get<2>(tuples[w - 1]), std::get<2>(tuples[r]));
            else
                tuples[w++] = tuples[r];
        }
        tuples.resize(w);
    }
    std::vector<std::tuple<IT, IT, NT>> tuples;

private:
    IT m, n;
};

// the raw view SpDCCols::GetDCSC() hands out (dcsc.h:124-131)
template <class IT, class NT>
struct Dcsc {
    typedef typename cb_storage<NT>::type ST;
    IT* cp; IT* jc; IT* ir; ST* numx;
    IT nz, nzc;
};

template <class IT, class NT>
class SpDCCols {
public:
    typedef IT LocalIT;
    typedef NT LocalNT;
    typedef typename cb_storage<NT>::type ST;
    static const IT esscount = 4;

    SpDCCols() : m(0), n(0) { cp.assign(1, 0); }
    SpDCCols(IT size, IT nRow, IT nCol, IT nzc_) : cp((size_t)nzc_ + 1, 0), jc((size_t)nzc_), ir((size_t)size), numx((size_t)size), m(nRow), n(nCol) {}
    SpDCCols(const SpTuples<IT, NT>& rhs, bool transpose) : m(transpose ? rhs.getncol() : rhs.getnrow()), n(transpose ? rhs.getnrow() : rhs.getncol()) {
        std::vector<std::tuple<IT, IT, NT>> t = rhs.tuples;
        if (transpose) for (auto& e : t) std::swap(std::get<0>(e), std::get<1>(e));
        build(t);
    }
    // column-sorted (or, with transpose, row-sorted) tuple array, as SpDCCols.cpp:197-304 takes it
    SpDCCols(IT nRow, IT nCol, IT nnz1, const std::tuple<IT, IT, NT>* rhs, bool transpose) : m(nRow), n(nCol) {
        std::vector<std::tuple<IT, IT, NT>> t(rhs, rhs + nnz1);
        if (transpose) for (auto& e : t) std::swap(std::get<0>(e), std::get<1>(e));
        build(t);
    }
    IT getnrow() const { return m; }
    IT getncol() const { return n; }
    IT getnnz() const { return (IT)ir.size(); }
    IT getnzc() const { return (IT)jc.size(); }
    bool isZero() const { return ir.empty(); }

    std::vector<IT> GetEssentials() const { return {getnnz(), m, n, getnzc()}; }
    void Create(const std::vector<IT>& ess) {                    // SpDCCols.cpp:734-745: allocate for the given essentials
        m = ess[1]; n = ess[2];
        ir.assign((size_t)ess[0], 0); numx.assign((size_t)ess[0], ST());
        jc.assign((size_t)ess[3], 0); cp.assign((size_t)ess[3] + 1, 0);
    }
    // build from a tuple array (reference SpDCCols.cpp Create(size, nRow, nCol, mytuples)); the array is consumed like there
    void Create(IT size, IT nRow, IT nCol, std::tuple<IT, IT, NT>* mytuples) {
        m = nRow; n = nCol;
        std::vector<std::tuple<IT, IT, NT>> t;
        if (mytuples && size > 0) t.assign(mytuples, mytuples + size);
        delete[] mytuples;
        build(t);
    }
    // A(ri, ci): the rows ri and columns ci of the tile, in that order; an empty index vector means "all"
    // (reference SpDCCols.cpp operator()(ri, ci), used as A.seq()(empty, single) and by SubsRefCol)
    SpDCCols operator()(const std::vector<IT>& ri, const std::vector<IT>& ci) const {
        std::vector<std::vector<IT>> newrow, newcol;             // old index -> new positions (an index may be listed twice)
        if (!ri.empty()) { newrow.resize((size_t)m); for (size_t q = 0; q < ri.size(); ++q) newrow[(size_t)ri[q]].push_back((IT)q); }
        if (!ci.empty()) { newcol.resize((size_t)n); for (size_t q = 0; q < ci.size(); ++q) newcol[(size_t)ci[q]].push_back((IT)q); }
        std::vector<std::tuple<IT, IT, NT>> t;
        for (size_t c = 0; c < jc.size(); ++c)
            for (IT p = cp[c]; p < cp[c + 1]; ++p) {
                const IT r = ir[(size_t)p], col = jc[c];
                const std::vector<IT> one_r{r}, one_c{col};
                const std::vector<IT>& rs = ri.empty() ? one_r : newrow[(size_t)r];
                const std::vector<IT>& cs = ci.empty() ? one_c : newcol[(size_t)col];
                for (IT nr : rs) for (IT nc : cs) t.emplace_back(nr, nc, (NT)numx[(size_t)p]);
            }
        SpDCCols out;
        out.m = ri.empty() ? m : (IT)ri.size();
        out.n = ci.empty() ? n : (IT)ci.size();
        out.build(t);
        return out;
    }
    Arr<IT, NT> GetArrays() const {                              // SpDCCols.cpp:826-846: {cp, jc, ir | numx}
        Arr<IT, NT> a(3, 1);
        SpDCCols* self = const_cast<SpDCCols*>(this);
        a.indarrs[0] = LocArr<IT>(self->cp.data(), cp.size());
        a.indarrs[1] = LocArr<IT>(self->jc.data(), jc.size());
        a.indarrs[2] = LocArr<IT>(self->ir.data(), ir.size());
        a.numarrs[0] = LocArr<ST>(self->numx.data(), numx.size());
        return a;
    }
    Dcsc<IT, NT>* GetDCSC() const {
        if (ir.empty()) return nullptr;
        SpDCCols* self = const_cast<SpDCCols*>(this);
        self->view.cp = self->cp.data(); self->view.jc = self->jc.data(); self->view.ir = self->ir.data(); self->view.numx = self->numx.data();
        self->view.nz = getnnz(); self->view.nzc = getnzc();
        return &self->view;
    }
    bool operator==(const SpDCCols& rhs) const {                 // dcsc.cpp:473-507 with ErrorTolerantEqual (Compare.h:46-65)
        if (ir.empty() && rhs.ir.empty()) return true;
        if (m != rhs.m || n != rhs.n || cp != rhs.cp || jc != rhs.jc || ir != rhs.ir) return false;
        for (size_t i = 0; i < numx.size(); ++i) {
            const double a = (double)numx[i], b = (double)rhs.numx[i], d = a > b ? a - b : b - a;
            if (std::is_floating_point<NT>::value ? !(d < EPSILON || d < EPSILON * std::max(std::abs(a), std::abs(b))) : numx[i] != rhs.numx[i]) return false;
        }
        return true;
    }
    std::vector<IT> cp, jc, ir;
    std::vector<ST> numx;

private:
    void build(std::vector<std::tuple<IT, IT, NT>>& t) {
        std::sort(t.begin(), t.end(), [](const std::tuple<IT, IT, NT>& a, const std::tuple<IT, IT, NT>& b) {
            return std::get<1>(a) != std::get<1>(b) ? std::get<1>(a) < std::get<1>(b) : std::get<0>(a) < std::get<0>(b);
        });
        cp.clear(); jc.clear(); ir.clear(); numx.clear();
        ir.reserve(t.size()); numx.reserve(t.size());
        for (size_t p = 0; p < t.size(); ++p) {
            if (p == 0 || std::get<1>(t[p]) != std::get<1>(t[p - 1])) { jc.push_back(std::get<1>(t[p])); cp.push_back((IT)p); }
            ir.push_back(std::get<0>(t[p]));
            numx.push_back((ST)std::get<2>(t[p]));
        }
        cp.push_back((IT)t.size());
    }
    IT m, n;
    Dcsc<IT, NT> view;
};

template <class IT, class NT>
class SpCCols {
public:
    typedef IT LocalIT;
    typedef NT LocalNT;
    typedef typename cb_storage<NT>::type ST;
    static const IT esscount = 3;

    SpCCols() : m(0), n(0) { jc.assign(1, 0); }
    SpCCols(IT size, IT nRow, IT nCol) : jc((size_t)nCol + 1, 0), ir((size_t)size), num((size_t)size), m(nRow), n(nCol) {}
    SpCCols(const SpTuples<IT, NT>& rhs, bool transpose) : m(transpose ? rhs.getncol() : rhs.getnrow()), n(transpose ? rhs.getnrow() : rhs.getncol()) {
        std::vector<std::tuple<IT, IT, NT>> t = rhs.tuples;
        if (transpose) for (auto& e : t) std::swap(std::get<0>(e), std::get<1>(e));
        std::sort(t.begin(), t.end(), [](const std::tuple<IT, IT, NT>& a, const std::tuple<IT, IT, NT>& b) {
            return std::get<1>(a) != std::get<1>(b) ? std::get<1>(a) < std::get<1>(b) : std::get<0>(a) < std::get<0>(b);
        });
        jc.assign((size_t)n + 1, 0);
        for (auto& e : t) jc[(size_t)std::get<1>(e) + 1]++;
        std::partial_sum(jc.begin(), jc.end(), jc.begin());
        for (auto& e : t) { ir.push_back(std::get<0>(e)); num.push_back((ST)std::get<2>(e)); }
    }
    IT getnrow() const { return m; }
    IT getncol() const { return n; }
    IT getnnz() const { return (IT)ir.size(); }
    bool isZero() const { return ir.empty(); }
    std::vector<IT> GetEssentials() const { return {getnnz(), m, n}; }
    void Create(const std::vector<IT>& ess) {
        m = ess[1]; n = ess[2];
        ir.assign((size_t)ess[0], 0); num.assign((size_t)ess[0], ST()); jc.assign((size_t)n + 1, 0);
    }
    Arr<IT, NT> GetArrays() const {                              // SpCCols.cpp:414-436: {jc, ir | num}
        Arr<IT, NT> a(2, 1);
        SpCCols* self = const_cast<SpCCols*>(this);
        a.indarrs[0] = LocArr<IT>(self->jc.data(), jc.size());
        a.indarrs[1] = LocArr<IT>(self->ir.data(), ir.size());
        a.numarrs[0] = LocArr<ST>(self->num.data(), num.size());
        return a;
    }
    std::vector<IT> jc, ir;
    std::vector<ST> num;

private:
    IT m, n;
};

// tile -> tuples (what SpDCCols::operator SpTuples does in the reference)
template <class IT, class NT>
SpTuples<IT, NT> TilesToTuples(const SpDCCols<IT, NT>& t) {
    SpTuples<IT, NT> out(0, t.getnrow(), t.getncol());
    out.tuples.reserve((size_t)t.getnnz());
    for (size_t c = 0; c < t.jc.size(); ++c)
        for (IT p = t.cp[c]; p < t.cp[c + 1]; ++p) out.tuples.emplace_back(t.ir[(size_t)p], t.jc[c], (NT)t.numx[(size_t)p]);
    return out;
}
template <class IT, class NT>
SpTuples<IT, NT> TilesToTuples(const SpCCols<IT, NT>& t) {
    SpTuples<IT, NT> out(0, t.getnrow(), t.getncol());
    out.tuples.reserve((size_t)t.getnnz());
    for (IT c = 0; c < t.getncol(); ++c)
        for (IT p = t.jc[(size_t)c]; p < t.jc[(size_t)c + 1]; ++p) out.tuples.emplace_back(t.ir[(size_t)p], c, (NT)t.num[(size_t)p]);
    return out;
}

// infer the tile type of a result from the tile type of an operand (SpDCCols.h:454-475 of the reference)
template <class DER, class NIT, class NNT> struct create_trait {};
template <class NIT, class NNT, class OIT, class ONT> struct create_trait<SpDCCols<OIT, ONT>, NIT, NNT> { typedef SpDCCols<NIT, NNT> T_inferred; };
template <class NIT, class NNT, class OIT, class ONT> struct create_trait<SpCCols<OIT, ONT>, NIT, NNT> { typedef SpCCols<NIT, NNT> T_inferred; };

}  // namespace combblas
#endif
