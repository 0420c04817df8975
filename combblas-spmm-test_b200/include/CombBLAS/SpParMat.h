// SpParMat<IT,NT,DER>: 2D-distributed sparse matrix = CommGrid + one local tile, with a device-resident mirror.
//
// Surface and distribution follow the reference (include/CombBLAS/SpParMat.h:67-452):
//   * block distribution with floor division, last grid row / column takes the remainder; Owner() and
//     GetPlaceInGlobalGrid() as in SpParMat.cpp:5066-5115;
//   * ctors (grid), (DER*, grid) [takes ownership, SpParMat.cpp:73-77], from global triples, ParallelReadMM
//     (Matrix Market with symmetric expansion, SpHelper.h:75-91), getnrow/getncol/getnnz (SpParMat.cpp:773-797),
//     seq()/seqptr(), PrintInfo (:2853-2875), LoadImbalance (:762-770), operator== (:2877-2884).
// B200-first differences, all behind the same calls:
//   * the tile is uploaded once (lazily) into HBM as doubly compressed rows and stays there (DeviceTile());
//   * ParallelReadMM splits the file by byte ranges like the reference; the triples are routed to their owners and merged on
//     the devices (cb_tile_from_distributed_coo).  The other ingestion paths (global triples, ReadDistribute) are replicated:
//     every process sees the whole input and keeps what it owns;
//   * GenGraph500() builds the tile directly on the GPU and never materialises it on the host.
#ifndef CB_SPPARMAT_H
#define CB_SPPARMAT_H

#include <fstream>
#include <functional>
#include <memory>
#include <sstream>
#include "CommGrid.h"
#include "DenseParMat.h"
#include "FullyDistVec.h"
#include "SpTuples.h"

namespace combblas {

template <class T>
struct maximum { T operator()(const T& x, const T& y) const { return x < y ? y : x; } };
template <class T>
struct cb_sum { T operator()(const T& x, const T& y) const { return x + y; } };

// DistEdgeList<IT>: the reference's distributed Graph500 edge list (include/CombBLAS/DistEdgeList.h:90-140).  Here it is a
// recipe, not a container: GenGraph500Data records scale / edge factor / initiator, and SpParMat(const DistEdgeList&, bool)
// has the device generator (csrc/cb_gen.cu) build the tiles directly in HBM - the edges never exist on the host.
// packed = true (what ReleaseTests/GenWriteMatrix.cpp asks for) is the reference's own stream: the Graph500 2.1 generator of
// RefGen21.h, restated for the device (cb_gen_graph500_tile) and pinned to the reference's generator bit for bit, seeded like the
// reference's -DDETERMINISTIC build; SpParMat(DEL, removeloops) then keeps or drops self loops as asked and sums duplicate edges
// (values are multiplicities), like SpParMat.cpp:3138-3254.  The initiator argument is ignored on that path, as the reference
// ignores it (RefGen21.h:69-76 fixes .57 / .19 / .19 / .05).
// packed = false: the reference seeds that generator by rank and wall clock (DistEdgeList.cpp:239-251), so its matrix depends on
// the process count; it is NOT reproduced.  The edges then come from this library's counter-based generator (same recipe, the
// given initiator, vertex ids always scrambled, self loops dropped, hashed weights) and rank 0 says so once on stderr.
template <class IT>
class DistEdgeList {
public:
    DistEdgeList() : commGrid(new CommGrid(MPI_COMM_WORLD, 0, 0)) {}
    explicit DistEdgeList(std::shared_ptr<CommGrid> grid) : commGrid(grid) {}
    void GenGraph500Data(double initiator[4], int log_numverts, int edgefactor, bool scramble = false, bool packed = false) {   // DistEdgeList.cpp:223-
        this->packed = packed;
        if (packed && !scramble) SpParHelper::Print("WARNING: Packed version does always generate scrambled vertex identifiers\n");   // DistEdgeList.cpp:225-228
        if (!packed) {
            static bool told = false;
            if (!told && commGrid->GetRank() == 0) {
                std::fprintf(stderr, "GenGraph500Data(packed = false): the reference's rank- and clock-seeded edge stream is not reproduced; edges come from "
                                     "the counter-based generator of this library (scrambled ids%s, self loops dropped, hashed weights)\n",
                             scramble ? "" : " although scramble = false was asked for");
                told = true;
            }
        }
        for (int i = 0; i < 4; ++i) init[i] = initiator[i];
        scale = log_numverts;
        ef = edgefactor;
        globalV = (IT)1 << log_numverts;
    }
    IT getGlobalV() const { return globalV; }
    std::shared_ptr<CommGrid> getcommgrid() const { return commGrid; }
    std::shared_ptr<CommGrid> commGrid;
    double init[4] = {0.57, 0.19, 0.19, 0.05};
    int scale = 0, ef = 16;
    bool packed = false;
    IT globalV = 0;
};

template <class IT, class NT, class DER>
class SpParMat {
public:
    typedef typename DER::LocalIT LocalIT;
    typedef typename DER::LocalNT LocalNT;
    typedef typename cb_storage<NT>::type ST;

    SpParMat() : commGrid(new CommGrid(MPI_COMM_WORLD, 0, 0)), spSeq(new DER()) {}
    explicit SpParMat(std::shared_ptr<CommGrid> grid) : commGrid(grid), spSeq(new DER()) {}
    SpParMat(DER* myseq, std::shared_ptr<CommGrid> grid) : commGrid(grid), spSeq(myseq) {}
    SpParMat(DER* myseq, std::shared_ptr<CommGrid> grid, IT total_m, IT total_n) : commGrid(grid), spSeq(myseq), gm(total_m), gn(total_n) {}
    // "matlab sparse" form (SpParMat.cpp:3066-3099) on replicated global triples
    SpParMat(IT total_m, IT total_n, const std::vector<IT>& rows, const std::vector<IT>& cols, const std::vector<NT>& vals,
             std::shared_ptr<CommGrid> grid, bool SumDuplicates = false)
        : commGrid(grid), spSeq(nullptr) {
        FromGlobalTriples(total_m, total_n, rows, cols, vals, SumDuplicates);
    }
    // conversion from a distributed edge list (reference SpParMat.cpp:3138-3254): generated on the device, see DistEdgeList
    template <class DELIT>
    SpParMat(const DistEdgeList<DELIT>& rhs, bool removeloops = true) : commGrid(rhs.commGrid), spSeq(nullptr) {
        if (rhs.packed) {
            // the reference's own stream and its own conversion: duplicates summed, loops kept unless asked otherwise
            gm = gn = (IT)1 << rhs.scale;
            IT r0, rl, c0, cl;
            BlockRange(gm, commGrid->GetGridRows(), commGrid->GetRankInProcCol(), r0, rl);
            BlockRange(gn, commGrid->GetGridCols(), commGrid->GetRankInProcRow(), c0, cl);
            const int vd = std::is_same<NT, bool>::value ? CB_PATTERN : cb_dtype_of<NT>::value;
            cb_ctx* ctx = commGrid->GetContext();
            cb_check(cb_gen_graph500_tile(ctx, rhs.scale, rhs.ef, 0, 0, 0, removeloops ? 1 : 0, r0, rl, c0, cl, vd, 0, &dtile), ctx, "cb_gen_graph500_tile");
            int64_t info[8];
            cb_tile_info(dtile, info);
            dnnz = info[0];
            devvals = vd != CB_PATTERN;
        } else {
            GenGraph500(rhs.scale, rhs.ef, false, 0, !std::is_same<NT, bool>::value, 1, rhs.init);      // never emits self loops
        }
    }
    // value-type conversion (reference SpParMat.h converting constructor / operator SpParMat<IT,NNT,NDER>()): same structure,
    // every value cast to NT (a nonzero becomes true for bool)
    template <class NNT, class NDER>
    SpParMat(const SpParMat<IT, NNT, NDER>& rhs) : commGrid(rhs.getcommgrid()), spSeq(nullptr), gm(rhs.getnrow()), gn(rhs.getncol()) {
        auto src = TilesToTuples(rhs.seq());
        SpTuples<LocalIT, NT> t(0, (LocalIT)rhs.seq().getnrow(), (LocalIT)rhs.seq().getncol());
        t.tuples.reserve((size_t)src.getnnz());
        for (int64_t p = 0; p < src.getnnz(); ++p) t.tuples.emplace_back((LocalIT)src.rowindex(p), (LocalIT)src.colindex(p), (NT)src.numvalue(p));
        spSeq = new DER(t, false);
    }
    SpParMat(const SpParMat& rhs) : commGrid(rhs.commGrid), spSeq(new DER(rhs.seq())), gm(rhs.gm), gn(rhs.gn) {}
    SpParMat& operator=(const SpParMat& rhs) {
        if (this != &rhs) { DER* copy = new DER(rhs.seq()); Release(); commGrid = rhs.commGrid; spSeq = copy; gm = rhs.gm; gn = rhs.gn; }
        return *this;
    }
    SpParMat(SpParMat&& rhs) noexcept
        : commGrid(rhs.commGrid), spSeq(rhs.spSeq), gm(rhs.gm), gn(rhs.gn), dtile(rhs.dtile), dpattern(rhs.dpattern), dnnz(rhs.dnnz),
          devvals(rhs.devvals) {
        rhs.spSeq = nullptr; rhs.dtile = nullptr; rhs.dpattern = nullptr;
    }
    ~SpParMat() { Release(); }

    // ---- construction helpers
    // Matrix Market coordinate file -> 0-based triples, with the transpose of every off-diagonal entry added for
    // symmetric files and value 1 for pattern files (ProcessLines / push_to_vectors, SpHelper.h:75-91,147-183).
    // Pure host code, no grid needed.  Returns false if the file cannot be opened.
    static bool ReadMMTriples(const std::string& filename, bool onebased, IT& tm, IT& tn, std::vector<IT>& rows,
                              std::vector<IT>& cols, std::vector<NT>& vals) {
        std::ifstream in(filename);
        if (!in) return false;
        std::string line;
        std::getline(in, line);
        std::string banner = line;
        for (auto& ch : banner) ch = (char)std::tolower(ch);
        const bool hasbanner = banner.compare(0, 14, "%%matrixmarket") == 0;
        const bool pattern = hasbanner && banner.find("pattern") != std::string::npos;
        const bool symmetric = hasbanner && (banner.find("symmetric") != std::string::npos || banner.find("hermitian") != std::string::npos);
        if (hasbanner) std::getline(in, line);
        while (!line.empty() && line[0] == '%' && std::getline(in, line)) {}          // comment lines before the size line
        long long m_ = 0, n_ = 0, nz_ = 0;
        std::istringstream(line) >> m_ >> n_ >> nz_;
        tm = (IT)m_; tn = (IT)n_;
        long long ii, jj;
        double vv = 1;
        while (std::getline(in, line)) {
            if (line.empty()) continue;
            std::istringstream ls(line);
            if (!(ls >> ii >> jj)) continue;
            if (!pattern) ls >> vv;
            if (onebased) { --ii; --jj; }
            rows.push_back((IT)ii); cols.push_back((IT)jj); vals.push_back((NT)vv);
            if (symmetric && ii != jj) { rows.push_back((IT)jj); cols.push_back((IT)ii); vals.push_back((NT)vv); }   // SpHelper.h:85-90
        }
        return true;
    }
    // this process's share of a Matrix Market file: the text of the lines that START inside its byte range of the data section
    // (the reference's split, SpParMat.cpp:4010-4080: every process reads fsize / nprocs bytes and finishes its last line)
    static bool ReadMMShareText(const std::string& filename, int rank, int nranks, IT& tm, IT& tn, bool& pattern, bool& symmetric, std::string& text) {
        std::ifstream in(filename, std::ios::binary);
        if (!in) return false;
        std::string line;
        std::getline(in, line);
        std::string banner = line;
        for (auto& ch : banner) ch = (char)std::tolower(ch);
        const bool hasbanner = banner.compare(0, 14, "%%matrixmarket") == 0;
        pattern = hasbanner && banner.find("pattern") != std::string::npos;
        symmetric = hasbanner && (banner.find("symmetric") != std::string::npos || banner.find("hermitian") != std::string::npos);
        if (hasbanner) std::getline(in, line);
        while (!line.empty() && line[0] == '%' && std::getline(in, line)) {}
        long long m_ = 0, n_ = 0, nz_ = 0;
        std::istringstream(line) >> m_ >> n_ >> nz_;
        tm = (IT)m_; tn = (IT)n_;
        const std::streamoff data0 = in.tellg();
        in.seekg(0, std::ios::end);
        const std::streamoff fsize = in.tellg();
        const std::streamoff len = fsize - data0;
        std::streamoff lo = data0 + len * rank / nranks, hi = data0 + len * (rank + 1) / nranks;
        in.clear();
        // a line that started before a boundary belongs to the process on its left: move both ends to the next line start
        auto next_line_start = [&](std::streamoff pos) -> std::streamoff {
            if (pos <= data0) return data0;
            if (pos >= fsize) return fsize;
            in.clear();
            in.seekg(pos - 1);
            char c;
            in.get(c);
            if (c == '\n') return pos;
            std::string skipped;
            std::getline(in, skipped);
            return in.eof() ? fsize : (std::streamoff)in.tellg();
        };
        lo = next_line_start(lo);
        hi = next_line_start(hi);
        text.assign((size_t)std::max<std::streamoff>(hi - lo, 0), '\0');
        if (hi > lo) {
            in.clear();
            in.seekg(lo);
            in.read(&text[0], hi - lo);
        }
        return true;
    }
    // the same share as 0-based triples parsed on the host (lines without two integers are skipped, like the reference's sscanf loop)
    static bool ReadMMShare(const std::string& filename, bool onebased, int rank, int nranks, IT& tm, IT& tn, std::vector<int64_t>& rows,
                            std::vector<int64_t>& cols, std::vector<ST>& vals) {
        bool pattern = false, symmetric = false;
        std::string text;
        if (!ReadMMShareText(filename, rank, nranks, tm, tn, pattern, symmetric, text)) return false;
        std::istringstream in(text);
        std::string line;
        long long ii, jj;
        double vv = 1;
        while (std::getline(in, line)) {
            if (line.empty()) continue;
            std::istringstream ls(line);
            if (!(ls >> ii >> jj)) continue;
            if (!pattern) ls >> vv;
            if (onebased) { --ii; --jj; }
            rows.push_back(ii); cols.push_back(jj); vals.push_back((ST)(NT)vv);
            if (symmetric && ii != jj) { rows.push_back(jj); cols.push_back(ii); vals.push_back((ST)(NT)vv); }   // SpHelper.h:85-90
        }
        return true;
    }
    template <typename BinOp> struct dup_rule { static const int value = -1; };
    template <typename T> struct dup_rule<maximum<T>> { static const int value = 2; };
    template <typename T> struct dup_rule<cb_sum<T>> { static const int value = 1; };
    template <typename T> struct dup_rule<std::plus<T>> { static const int value = 1; };

    // ParallelReadMM (SpParMat.cpp:3978-4115): every process reads its own byte range of the file; the text is cut into lines
    // and parsed on its GPU, the triples travel to their owners between the GPUs and are merged and turned into the tile there
    // (cb_tile_from_mm_text) - the host tile is only materialised when somebody asks for seq().  A file with numbers the device
    // parser does not reproduce exactly (it says so on every rank), a boolean matrix read from a file with values, or
    // CB_READMM_HOSTPARSE=1 parse on the host and hand the triples to the same routing (cb_tile_from_distributed_coo).  A merge
    // rule the device does not know (anything but maximum / plus) takes the replicated host path: every process reads the
    // whole file and keeps what it owns.
    template <typename BinOp = maximum<NT>>
    void ParallelReadMM(const std::string& filename, bool onebased, BinOp binop = BinOp()) {
        static const bool replicated = std::getenv("CB_READMM_REPLICATED") && std::atoi(std::getenv("CB_READMM_REPLICATED")) != 0;
        static const bool hostparse = std::getenv("CB_READMM_HOSTPARSE") && std::atoi(std::getenv("CB_READMM_HOSTPARSE")) != 0;
        if (dup_rule<BinOp>::value >= 0 && !replicated) {
            IT tm = 0, tn = 0;
            bool pattern = false, symmetric = false;
            std::string text;
            const bool found = ReadMMShareText(filename, commGrid->GetRank(), commGrid->GetSize(), tm, tn, pattern, symmetric, text);
            if (commGrid->MinWorld(found ? 1 : 0) == 0) {
                SpParHelper::Print("COMBBLAS: Matrix-market file " + filename + " can not be found\n");
                MPI_Abort(MPI_COMM_WORLD, NOFILE);
            }
            Release();
            spSeq = nullptr;
            gm = tm; gn = tn;
            cb_ctx* ctx = commGrid->GetContext();
            int vd = cb_dtype_of<NT>::value;
            bool done = false;
            if (!hostparse && (!std::is_same<NT, bool>::value || pattern)) {
                if (std::is_same<NT, bool>::value) vd = CB_PATTERN;
                const int flags = (onebased ? 1 : 0) | (pattern ? 2 : 0) | (symmetric ? 4 : 0);
                const int status = cb_tile_from_mm_text(ctx, (int64_t)gm, (int64_t)gn, text.data(), (int64_t)text.size(), flags, vd,
                                                        dup_rule<BinOp>::value, &dtile);
                if (status == CB_OK) done = true;
                else if (status != CB_ERR_UNSUPPORTED) cb_check(status, ctx, "cb_tile_from_mm_text");
                else SpParHelper::Print("COMBBLAS: " + filename + " is not for the device parser (numbers outside its exact range, or a share of 1 GiB or more), parsing it on the host\n");
            }
            if (!done) {
                std::vector<int64_t> rows, cols;
                std::vector<ST> vals;
                ReadMMShare(filename, onebased, commGrid->GetRank(), commGrid->GetSize(), tm, tn, rows, cols, vals);
                vd = cb_dtype_of<NT>::value;
                if (std::is_same<NT, bool>::value) {             // a boolean matrix whose stored entries are all true is a pattern
                    const bool alltrue = std::all_of(vals.begin(), vals.end(), [](ST v) { return v != 0; });
                    if (commGrid->MinWorld(alltrue ? 1 : 0) == 1) vd = CB_PATTERN;
                }
                cb_check(cb_tile_from_distributed_coo(ctx, (int64_t)gm, (int64_t)gn, (int64_t)rows.size(), rows.data(), cols.data(),
                                                      vd == CB_PATTERN ? nullptr : (const void*)vals.data(), vd, dup_rule<BinOp>::value, &dtile),
                         ctx, "cb_tile_from_distributed_coo");
            }
            int64_t info[8];
            cb_tile_info(dtile, info);
            dnnz = info[0];
            devvals = vd != CB_PATTERN;
            return;
        }
        IT tm = 0, tn = 0;
        std::vector<IT> rows, cols;
        std::vector<NT> vals;
        if (!ReadMMTriples(filename, onebased, tm, tn, rows, cols, vals)) {
            SpParHelper::Print("COMBBLAS: Matrix-market file " + filename + " can not be found\n");
            MPI_Abort(MPI_COMM_WORLD, NOFILE);
        }
        FromGlobalTriples(tm, tn, rows, cols, vals, true, binop);
    }

    // Triples file "m n nnz" + one "i j v" line per entry, one-based (reference SpParMat.cpp:4213-4400: the master reads and
    // scatters; here every process reads the file and keeps what it owns).  nonum: no value column, every entry is 1.
    // A missing file leaves an empty 0 x 0 matrix after the reference's message (SpParMat.cpp:4274-4279).
    void ReadDistribute(const std::string& filename, int master, bool nonum = false, bool pario = false) {
        (void)master; (void)pario;
        std::ifstream in(filename);
        if (!in) {
            SpParHelper::Print("COMBBLAS: Input file doesn't exist\n");
            FromGlobalTriples(0, 0, std::vector<IT>(), std::vector<IT>(), std::vector<NT>(), false);
            return;
        }
        std::string line;
        do { std::getline(in, line); } while (in && !line.empty() && line[0] == '%');
        long long tm = 0, tn = 0, tnz = 0;
        std::istringstream(line) >> tm >> tn >> tnz;
        std::vector<IT> rows, cols;
        std::vector<NT> vals;
        long long ii, jj;
        double vv = 1;
        while (std::getline(in, line)) {
            std::istringstream ls(line);
            if (!(ls >> ii >> jj)) continue;
            if (!nonum) ls >> vv;
            rows.push_back((IT)(ii - 1)); cols.push_back((IT)(jj - 1)); vals.push_back((NT)vv);
        }
        FromGlobalTriples((IT)tm, (IT)tn, rows, cols, vals, false);
    }

    // Matrix Market coordinate file of the whole matrix, every process contributing its tile in rank order, column by
    // column inside a tile - the text the reference's ParallelWriteMM produces (SpParMat.cpp:4118-4210: header by rank 0,
    // "row<TAB>col<TAB>value" lines).  Round-trips through ParallelReadMM; used for archived results.
    void ParallelWriteMM(const std::string& filename, bool onebased) const {
        const IT totalm = getnrow(), totaln = getncol(), totnnz = getnnz();
        IT roffset = 0, coffset = 0, len = 0;
        BlockRange(totalm, commGrid->GetGridRows(), commGrid->GetRankInProcCol(), roffset, len);
        BlockRange(totaln, commGrid->GetGridCols(), commGrid->GetRankInProcRow(), coffset, len);
        if (onebased) { roffset += 1; coffset += 1; }
        std::ostringstream ss;
        ss.precision(17);
        if (commGrid->GetRank() == 0) ss << "%%MatrixMarket matrix coordinate real general\n" << totalm << " " << totaln << " " << totnnz << "\n";
        SpTuples<LocalIT, NT> tup = TilesToTuples(seq());
        for (int64_t p = 0; p < tup.getnnz(); ++p)
            ss << (IT)tup.rowindex(p) + roffset << '\t' << (IT)tup.colindex(p) + coffset << '\t' << +tup.numvalue(p) << '\n';
        const std::string text = ss.str();
        for (int q = 0; q < commGrid->GetSize(); ++q) {                        // one writer at a time, in rank order
            if (q == commGrid->GetRank()) {
                std::ofstream out(filename, q == 0 ? std::ios::trunc : std::ios::app);
                out << text;
            }
            if (commGrid->GetSize() > 1) MPI_Barrier(commGrid->GetWorld());
        }
    }
    void SaveGathered(const std::string& filename) const { ParallelWriteMM(filename, true); }

    // Graph500-style Kronecker matrix generated on the device (GenWriteMatrix.cpp:96-131 recipe).  The local tile never
    // exists on the host; seq() downloads it on first use.
    void GenGraph500(int scale, int edgefactor, bool symmetric = true, uint64_t seed = 0, bool values = false, uint64_t val_seed = 1,
                     const double* initiator = nullptr) {
        Release();
        spSeq = nullptr;
        gm = gn = (IT)1 << scale;
        IT r0, rl, c0, cl;
        BlockRange(gm, commGrid->GetGridRows(), commGrid->GetRankInProcCol(), r0, rl);
        BlockRange(gn, commGrid->GetGridCols(), commGrid->GetRankInProcRow(), c0, cl);
        const double dflt[4] = {0.57, 0.19, 0.19, 0.05};
        const int vd = (values && !std::is_same<NT, bool>::value) ? cb_dtype_of<NT>::value : CB_PATTERN;
        cb_ctx* ctx = commGrid->GetContext();
        cb_check(cb_gen_rmat_tile(ctx, scale, edgefactor, seed, initiator ? initiator : dflt, symmetric ? 1 : 0, r0, rl, c0, cl, vd, val_seed, &dtile),
                 ctx, "cb_gen_rmat_tile");
        int64_t info[8];
        cb_tile_info(dtile, info);
        dnnz = info[0];
        devvals = vd != CB_PATTERN;
    }

    // ---- queries
    IT getnrow() const { return gm >= 0 ? gm : (IT)commGrid->SumCol(getlocalrows()); }
    IT getncol() const { return gn >= 0 ? gn : (IT)commGrid->SumRow(getlocalcols()); }
    IT getnnz() const { return (IT)commGrid->SumWorld(getlocalnnz()); }
    LocalIT getlocalrows() const { return spSeq ? spSeq->getnrow() : LocalRows(); }
    LocalIT getlocalcols() const { return spSeq ? spSeq->getncol() : LocalCols(); }
    LocalIT getlocalnnz() const { return spSeq ? spSeq->getnnz() : (LocalIT)dnnz; }
    std::shared_ptr<CommGrid> getcommgrid() const { return commGrid; }
    DER& seq() const { const_cast<SpParMat*>(this)->MaterialiseHost(); return *spSeq; }
    DER* seqptr() const { const_cast<SpParMat*>(this)->MaterialiseHost(); return spSeq; }

    float LoadImbalance() const {
        const int64_t tot = commGrid->SumWorld(getlocalnnz()), mx = commGrid->MaxWorld(getlocalnnz());
        return tot ? (float)mx / ((float)tot / commGrid->GetSize()) : 1.0f;
    }
    void PrintInfo() const {
        const IT mm = getnrow(), nn = getncol(), nz = getnnz();
        if (commGrid->GetRank() == 0)
            std::cout << "As a whole: " << mm << " rows and " << nn << " columns and " << nz << " nonzeros" << std::endl;
    }
    bool operator==(const SpParMat& rhs) const {
        const int local = (seq() == rhs.seq()) ? 1 : 0;
        return commGrid->MinWorld(local) == 1;
    }

    // scale every stored entry by the dense entry at its position (reference SpParMat.cpp:2817-2850, Dcsc::EWiseScale);
    // the device mirror is rebuilt at the next multiply
    void EWiseScale(const DenseParMat<IT, NT>& rhs) {
        if (*commGrid != *rhs.getcommgrid()) {
            SpParHelper::Print("Grids are not comparable elementwise multiplication\n");
            MPI_Abort(MPI_COMM_WORLD, GRIDMISMATCH);
        }
        DER& t = seq();
        if ((IT)t.getnrow() != rhs.getlocalrows() || (IT)t.getncol() != rhs.getlocalcols()) {
            SpParHelper::Print("Local dimensions do not match for EWiseScale\n");
            MPI_Abort(MPI_COMM_WORLD, DIMMISMATCH);
        }
        SpTuples<LocalIT, NT> tup = TilesToTuples(t);
        for (int64_t p = 0; p < tup.getnnz(); ++p)
            std::get<2>(tup.tuples[(size_t)p]) = (NT)(tup.numvalue(p) * rhs(tup.rowindex(p), tup.colindex(p)));
        DER* scaled = new DER(tup, false);
        const IT keep_m = gm, keep_n = gn;
        Release();
        spSeq = scaled; gm = keep_m; gn = keep_n;
    }

    // drop the entries on the global diagonal; returns how many there were (reference SpParMat.cpp RemoveLoops)
    IT RemoveLoops() {
        IT roff = 0, coff = 0;
        GetPlaceInGlobalGrid(roff, coff);
        SpTuples<LocalIT, NT> tup = TilesToTuples(seq()), kept(0, seq().getnrow(), seq().getncol());
        for (int64_t p = 0; p < tup.getnnz(); ++p)
            if ((IT)tup.rowindex(p) + roff != (IT)tup.colindex(p) + coff) kept.tuples.push_back(tup.tuples[(size_t)p]);
        const int64_t removed = tup.getnnz() - kept.getnnz();
        if (removed > 0) {
            DER* fresh = new DER(kept, false);
            const IT keep_m = gm, keep_n = gn;
            Release();
            spSeq = fresh; gm = keep_m; gn = keep_n;
        }
        return (IT)commGrid->SumWorld(removed);
    }

    // value(i,j) <- op(value(i,j), x[j]) for dim == Column, x[i] for dim == Row (reference SpParMat.cpp:801-880)
    template <typename _BinaryOperation>
    void DimApply(Dim dim, const FullyDistVec<IT, NT>& x, _BinaryOperation op) {
        if (*x.getcommgrid() != *commGrid) {
            SpParHelper::Print("Grids are not comparable, SpParMat::DimApply() fails!\n");
            MPI_Abort(MPI_COMM_WORLD, GRIDMISMATCH);
        }
        if (x.TotalLength() != (dim == Column ? getncol() : getnrow())) {
            SpParHelper::Print("Vector length does not match the matrix dimension in DimApply\n");
            MPI_Abort(MPI_COMM_WORLD, DIMMISMATCH);
        }
        const std::vector<NT> xw = x.Gather();
        IT roff = 0, coff = 0;
        GetPlaceInGlobalGrid(roff, coff);
        SpTuples<LocalIT, NT> tup = TilesToTuples(seq());
        for (auto& e : tup.tuples) {
            const IT g = dim == Column ? (IT)std::get<1>(e) + coff : (IT)std::get<0>(e) + roff;
            std::get<2>(e) = (NT)op(std::get<2>(e), xw[(size_t)g]);
        }
        DER* fresh = new DER(tup, false);
        const IT keep_m = gm, keep_n = gn;
        Release();
        spSeq = fresh; gm = keep_m; gn = keep_n;
    }
    // columns ci (LOCAL indices, the same list on every process) of every tile: a matrix with |ci| columns per processor column
    // (reference SpParMat.cpp:2012-2017)
    SpParMat SubsRefCol(const std::vector<IT>& ci) const {
        std::vector<LocalIT> ri, lci(ci.begin(), ci.end());
        return SpParMat(new DER(seq()(ri, lci)), commGrid);
    }
    // apply a unary function to every stored value (reference SpParMat.h Apply)
    template <typename _UnaryOperation>
    void Apply(_UnaryOperation f) {
        SpTuples<LocalIT, NT> tup = TilesToTuples(seq());
        for (auto& e : tup.tuples) std::get<2>(e) = (NT)f(std::get<2>(e));
        DER* fresh = new DER(tup, false);
        const IT keep_m = gm, keep_n = gn;
        Release();
        spSeq = fresh; gm = keep_m; gn = keep_n;
    }

    // Fold the matrix along a dimension into a distributed vector (reference SpParMat.cpp:929-1100): dim == Row folds every
    // row over its nonzeros (result of length getnrow()), dim == Column every column (length getncol()); the fold starts from
    // `id`, applies `uop` to every value and combines with `op`; rows / columns without nonzeros hold `id`.  Local fold first,
    // then the partial vectors of the processes that share the rows (columns), in grid order.  Host-side, as in the reference.
    template <typename GIT, typename VT, typename _BinaryOperation, typename _UnaryOperation>
    void Reduce(FullyDistVec<GIT, VT>& rvec, Dim dim, _BinaryOperation op, VT id, _UnaryOperation uop) const {
        if (*rvec.getcommgrid() != *commGrid) {
            SpParHelper::Print("Grids are not comparable, SpParMat::Reduce() fails!\n");
            MPI_Abort(MPI_COMM_WORLD, GRIDMISMATCH);
        }
        const int pr = commGrid->GetGridRows(), pc = commGrid->GetGridCols();
        SpTuples<LocalIT, NT> tup = TilesToTuples(seq());
        std::vector<VT> part(dim == Row ? (size_t)seq().getnrow() : (size_t)seq().getncol(), id);
        for (int64_t p = 0; p < tup.getnnz(); ++p) {
            VT& dst = part[dim == Row ? (size_t)tup.rowindex(p) : (size_t)tup.colindex(p)];
            dst = op(dst, (VT)uop(tup.numvalue(p)));
        }
        std::vector<std::vector<char>> all;
        cb_host_allgatherv(part.data(), part.size() * sizeof(VT), all);
        std::vector<VT> whole((size_t)(dim == Row ? getnrow() : getncol()), id);
        size_t off = 0;
        for (int b = 0; b < (dim == Row ? pr : pc); ++b) {
            size_t blen = 0;
            for (int q = 0; q < (dim == Row ? pc : pr); ++q) {
                const std::vector<char>& buf = all[(size_t)(dim == Row ? commGrid->GetRank(b, q) : commGrid->GetRank(q, b))];
                blen = buf.size() / sizeof(VT);
                const VT* v = reinterpret_cast<const VT*>(buf.data());
                for (size_t i = 0; i < blen; ++i) whole[off + i] = op(whole[off + i], v[i]);
            }
            off += blen;
        }
        rvec.Scatter(whole);
    }
    template <typename GIT, typename VT, typename _BinaryOperation>
    void Reduce(FullyDistVec<GIT, VT>& rvec, Dim dim, _BinaryOperation op, VT id) const {
        Reduce(rvec, dim, op, id, [](NT v) { return v; });
    }

    // A <- A^T (reference SpParMat.cpp Transpose: tiles swap between processes (i,j) and (j,i)).  Here every process
    // publishes its triples in global coordinates and keeps the transposed ones it owns - host-side, like all ingestion.
    void Transpose() {
        const IT tm = getnrow(), tn = getncol();
        IT roff = 0, coff = 0;
        GetPlaceInGlobalGrid(roff, coff);
        SpTuples<LocalIT, NT> tup = TilesToTuples(seq());
        struct Trip { IT r, c; ST v; };
        std::vector<Trip> mine((size_t)tup.getnnz());
        for (int64_t p = 0; p < tup.getnnz(); ++p) mine[(size_t)p] = Trip{(IT)tup.rowindex(p) + roff, (IT)tup.colindex(p) + coff, (ST)tup.numvalue(p)};
        std::vector<std::vector<char>> all;
        cb_host_allgatherv(mine.data(), mine.size() * sizeof(Trip), all);
        std::vector<IT> rows, cols;
        std::vector<NT> vals;
        const int me = commGrid->GetRank();
        for (const std::vector<char>& b : all) {
            const Trip* t = reinterpret_cast<const Trip*>(b.data());
            for (size_t q = 0; q < b.size() / sizeof(Trip); ++q) {
                LocalIT lr, lc;
                if (Owner(tn, tm, t[q].c, t[q].r, lr, lc) == me) { rows.push_back(t[q].c); cols.push_back(t[q].r); vals.push_back((NT)t[q].v); }
            }
        }
        FromGlobalTriplesOwned(tn, tm, rows, cols, vals);
    }
    // element-wise A += B for matrices with the same distribution (reference SpParMat.cpp operator+=: local tile addition)
    SpParMat& operator+=(const SpParMat& rhs) {
        if (*commGrid != *rhs.commGrid) {
            SpParHelper::Print("Grids are not comparable for parallel addition (A+B)\n");
            MPI_Abort(MPI_COMM_WORLD, GRIDMISMATCH);
        }
        if (getnrow() != rhs.getnrow() || getncol() != rhs.getncol()) {
            SpParHelper::Print("Dimensions do not match for parallel addition (A+B)\n");
            MPI_Abort(MPI_COMM_WORLD, DIMMISMATCH);
        }
        SpTuples<LocalIT, NT> a = TilesToTuples(seq()), b = TilesToTuples(rhs.seq());
        a.tuples.insert(a.tuples.end(), b.tuples.begin(), b.tuples.end());
        a.RemoveDuplicates(cb_sum<NT>());
        DER* sum = new DER(a, false);
        const IT keep_m = getnrow(), keep_n = getncol();
        Release();
        spSeq = sum; gm = keep_m; gn = keep_n;
        return *this;
    }

    // Owner of global entry (grow, gcol) and its local indices (SpParMat.cpp:5066-5096)
    template <typename LIT>
    int Owner(IT total_m, IT total_n, IT grow, IT gcol, LIT& lrow, LIT& lcol) const {
        const int procrows = commGrid->GetGridRows(), proccols = commGrid->GetGridCols();
        const IT m_perproc = total_m / procrows, n_perproc = total_n / proccols;
        const int own_procrow = m_perproc ? std::min((int)(grow / m_perproc), procrows - 1) : procrows - 1;
        const int own_proccol = n_perproc ? std::min((int)(gcol / n_perproc), proccols - 1) : proccols - 1;
        lrow = (LIT)(grow - own_procrow * m_perproc);
        lcol = (LIT)(gcol - own_proccol * n_perproc);
        return commGrid->GetRank(own_procrow, own_proccol);
    }
    void GetPlaceInGlobalGrid(IT& rowOffset, IT& colOffset) const {
        rowOffset = commGrid->GetRankInProcCol() * (getnrow() / commGrid->GetGridRows());
        colOffset = commGrid->GetRankInProcRow() * (getncol() / commGrid->GetGridCols());
    }
    // the same rule as Owner() for an explicit grid shape (usable without a CommGrid)
    template <typename LIT>
    static int OwnerOnGrid(int procrows, int proccols, IT total_m, IT total_n, IT grow, IT gcol, LIT& lrow, LIT& lcol) {
        const IT m_perproc = total_m / procrows, n_perproc = total_n / proccols;
        const int own_procrow = m_perproc ? std::min((int)(grow / m_perproc), procrows - 1) : procrows - 1;
        const int own_proccol = n_perproc ? std::min((int)(gcol / n_perproc), proccols - 1) : proccols - 1;
        lrow = (LIT)(grow - own_procrow * m_perproc);
        lcol = (LIT)(gcol - own_proccol * n_perproc);
        return own_procrow * proccols + own_proccol;
    }
    static void BlockRange(IT total, int nb, int b, IT& start, IT& len) {
        const IT per = total / nb;
        start = (IT)b * per;
        len = (b == nb - 1) ? total - start : per;
    }

    // ---- the device mirror: uploaded once, reused by every multiply
    cb_tile* DeviceTile() const {
        SpParMat* self = const_cast<SpParMat*>(this);
        if (!self->dtile) {
            cb_ctx* ctx = commGrid->GetContext();
            self->Upload(ctx, *spSeq);
            int64_t info[8];
            cb_tile_info(self->dtile, info);
            self->dnnz = info[0];
        }
        return self->dtile;
    }
    // the same tile without its values (structure only), for structural products under the boolean semiring
    cb_tile* DevicePatternTile() const {
        SpParMat* self = const_cast<SpParMat*>(this);
        if (!self->dpattern) cb_check(cb_tile_pattern_view(DeviceTile(), &self->dpattern), commGrid->GetContext(), "cb_tile_pattern_view");
        return self->dpattern;
    }
    void FreeDeviceTile() {
        if (dpattern) { cb_tile_free(dpattern); dpattern = nullptr; }
        if (dtile) { cb_tile_free(dtile); dtile = nullptr; }
    }

private:
    template <typename BinOp = maximum<NT>>
    void FromGlobalTriples(IT total_m, IT total_n, const std::vector<IT>& rows, const std::vector<IT>& cols, const std::vector<NT>& vals,
                           bool mergedups, BinOp binop = BinOp()) {
        Release();
        gm = total_m; gn = total_n;
        IT r0, rl, c0, cl;
        BlockRange(gm, commGrid->GetGridRows(), commGrid->GetRankInProcCol(), r0, rl);
        BlockRange(gn, commGrid->GetGridCols(), commGrid->GetRankInProcRow(), c0, cl);
        SpTuples<LocalIT, NT> mine(0, (LocalIT)rl, (LocalIT)cl);
        const int me = commGrid->GetRank();
        for (size_t i = 0; i < rows.size(); ++i) {
            LocalIT lr, lc;
            if (Owner(gm, gn, rows[i], cols[i], lr, lc) == me) mine.tuples.emplace_back(lr, lc, vals[i]);
        }
        if (mergedups) mine.RemoveDuplicates(binop);     // SparseCommon, SpParMat.cpp:2962-2967
        else mine.SortColBased();
        spSeq = new DER(mine, false);
    }
    // triples that are already known to belong to this process (global coordinates)
    void FromGlobalTriplesOwned(IT total_m, IT total_n, const std::vector<IT>& rows, const std::vector<IT>& cols, const std::vector<NT>& vals) {
        Release();
        gm = total_m; gn = total_n;
        IT r0, rl, c0, cl;
        BlockRange(gm, commGrid->GetGridRows(), commGrid->GetRankInProcCol(), r0, rl);
        BlockRange(gn, commGrid->GetGridCols(), commGrid->GetRankInProcRow(), c0, cl);
        SpTuples<LocalIT, NT> mine(0, (LocalIT)rl, (LocalIT)cl);
        for (size_t i = 0; i < rows.size(); ++i) mine.tuples.emplace_back((LocalIT)(rows[i] - r0), (LocalIT)(cols[i] - c0), vals[i]);
        mine.SortColBased();
        spSeq = new DER(mine, false);
    }
    void Upload(cb_ctx* ctx, const SpDCCols<LocalIT, NT>& t) {
        const int idt = sizeof(LocalIT) == 4 ? CB_I32 : CB_I64;
        const bool pattern = std::is_same<NT, bool>::value && std::all_of(t.numx.begin(), t.numx.end(), [](ST v) { return v != 0; });
        cb_check(cb_tile_upload_csc(ctx, t.getnrow(), t.getncol(), t.getnnz(), t.getnzc(), t.cp.data(), t.jc.data(), t.ir.data(),
                                    pattern ? nullptr : (const void*)t.numx.data(), idt, pattern ? CB_PATTERN : cb_dtype_of<NT>::value, &dtile),
                 ctx, "cb_tile_upload_csc");
    }
    void Upload(cb_ctx* ctx, const SpCCols<LocalIT, NT>& t) {
        const int idt = sizeof(LocalIT) == 4 ? CB_I32 : CB_I64;
        const bool pattern = std::is_same<NT, bool>::value && std::all_of(t.num.begin(), t.num.end(), [](ST v) { return v != 0; });
        cb_check(cb_tile_upload_csc(ctx, t.getnrow(), t.getncol(), t.getnnz(), 0, t.jc.data(), nullptr, t.ir.data(),
                                    pattern ? nullptr : (const void*)t.num.data(), idt, pattern ? CB_PATTERN : cb_dtype_of<NT>::value, &dtile),
                 ctx, "cb_tile_upload_csc");
    }
    LocalIT LocalRows() const { IT s, l; BlockRange(gm, commGrid->GetGridRows(), commGrid->GetRankInProcCol(), s, l); return (LocalIT)l; }
    LocalIT LocalCols() const { IT s, l; BlockRange(gn, commGrid->GetGridCols(), commGrid->GetRankInProcRow(), s, l); return (LocalIT)l; }
    void MaterialiseHost() {
        if (spSeq || !dtile) { if (!spSeq) spSeq = new DER(); return; }
        // device tile (generated on the GPU) -> host tile, through CSR triples
        const LocalIT lm = LocalRows(), ln = LocalCols();
        std::vector<int64_t> rowptr((size_t)lm + 1), col((size_t)dnnz);
        std::vector<ST> vals;
        int64_t info[8];
        cb_tile_info(dtile, info);
        const bool hasvals = !std::is_same<NT, bool>::value && cb_tile_has_values();
        if (hasvals) vals.resize((size_t)dnnz);
        cb_check(cb_tile_download_csr(dtile, rowptr.data(), col.data(), hasvals ? vals.data() : nullptr), commGrid->GetContext(), "cb_tile_download_csr");
        SpTuples<LocalIT, NT> t(0, lm, ln);
        t.tuples.reserve((size_t)dnnz);
        for (LocalIT r = 0; r < lm; ++r)
            for (int64_t p = rowptr[(size_t)r]; p < rowptr[(size_t)r + 1]; ++p)
                t.tuples.emplace_back(r, (LocalIT)col[(size_t)p], hasvals ? (NT)vals[(size_t)p] : NT(1));
        t.SortColBased();
        spSeq = new DER(t, false);
    }
    bool cb_tile_has_values() const { return devvals; }
    void Release() {
        FreeDeviceTile();
        delete spSeq;
        spSeq = nullptr;
    }

    std::shared_ptr<CommGrid> commGrid;
    DER* spSeq;
    IT gm = -1, gn = -1;            // global dimensions when known without communication
    cb_tile* dtile = nullptr;
    cb_tile* dpattern = nullptr;
    int64_t dnnz = 0;
    bool devvals = false;
};

// Element-wise product of two sparse matrices with the same distribution (reference ParFriends.h:2174-2200, sequential
// EWiseMult of Friends.h): exclude == false keeps the entries present in both, value A(i,j) * B(i,j) in the promoted type;
// exclude == true keeps the entries of A that are NOT in B, with A's values.
template <typename IU, typename NU1, typename NU2, typename UDERA, typename UDERB>
SpParMat<IU, typename promote_trait<NU1, NU2>::T_promote, SpDCCols<typename UDERA::LocalIT, typename promote_trait<NU1, NU2>::T_promote>>
EWiseMult(const SpParMat<IU, NU1, UDERA>& A, const SpParMat<IU, NU2, UDERB>& B, bool exclude) {
    typedef typename promote_trait<NU1, NU2>::T_promote N_promote;
    typedef typename UDERA::LocalIT LIT;
    typedef SpDCCols<LIT, N_promote> DER_promote;
    if (*A.getcommgrid() != *B.getcommgrid()) {
        SpParHelper::Print("Grids are not comparable elementwise multiplication\n");
        MPI_Abort(MPI_COMM_WORLD, GRIDMISMATCH);
    }
    auto ta = TilesToTuples(A.seq());
    auto tb = TilesToTuples(B.seq());                             // both column-major sorted
    SpTuples<LIT, N_promote> out(0, (LIT)A.seq().getnrow(), (LIT)A.seq().getncol());
    int64_t q = 0;
    for (int64_t p = 0; p < ta.getnnz(); ++p) {
        const auto ka = std::make_pair(ta.colindex(p), ta.rowindex(p));
        while (q < tb.getnnz() && std::make_pair((decltype(ka.first))tb.colindex(q), (decltype(ka.second))tb.rowindex(q)) < ka) ++q;
        const bool both = q < tb.getnnz() && (decltype(ka.first))tb.colindex(q) == ka.first && (decltype(ka.second))tb.rowindex(q) == ka.second;
        if (exclude && !both) out.tuples.emplace_back(ta.rowindex(p), ta.colindex(p), (N_promote)ta.numvalue(p));
        if (!exclude && both) out.tuples.emplace_back(ta.rowindex(p), ta.colindex(p), (N_promote)((N_promote)ta.numvalue(p) * (N_promote)tb.numvalue(q)));
    }
    return SpParMat<IU, N_promote, DER_promote>(new DER_promote(out, false), A.getcommgrid(), A.getnrow(), A.getncol());
}

template <class IT, class NT>
template <typename DER>
DenseParMat<IT, NT>& DenseParMat<IT, NT>::operator+=(const SpParMat<IT, NT, DER>& rhs) {
    if (*commGrid != *rhs.getcommgrid()) {
        SpParHelper::Print("Grids are not comparable elementwise addition\n");
        MPI_Abort(MPI_COMM_WORLD, GRIDMISMATCH);
    }
    const DER& t = rhs.seq();
    if ((IT)t.getnrow() != m || (IT)t.getncol() != n) {
        SpParHelper::Print("Local dimensions do not match for DenseParMat += SpParMat\n");
        MPI_Abort(MPI_COMM_WORLD, DIMMISMATCH);
    }
    auto tup = TilesToTuples(t);
    for (int64_t p = 0; p < tup.getnnz(); ++p) (*this)(tup.rowindex(p), tup.colindex(p)) += (ST)tup.numvalue(p);
    return *this;
}

}  // namespace combblas
#endif
