// TEST INFRASTRUCTURE ONLY: a host-memory stand-in for libcombblas_b200.so, so that the C++ HOST LAYER
// (combblas-spmm-test_b200/include/CombBLAS/*.h: SpParMat, DenseParMat, FullyDistVec, SpMM, SpMV, Mult_AnXBn_Synch, ...)
// and its driver can be exercised by the CPU test suite on a machine without a GPU.  tests/test_host_mock_cpu.py compiles
// the driver against THIS library in a temporary directory; nothing in the product links, loads or ships it, and the real
// library still has no CPU path (tests/test_abi_cpu.py).  Multiplies are naive loops over
// the semiring definitions of include/CombBLAS/Semirings.h - the driver's own replays and the GPU suite are the checkers of
// arithmetic, this only has to be a faithful enough ABI for the host logic to run.
// Process grids: with RANK / WORLD_SIZE set (one OS process per rank, as under torch.distributed.run) the mock exchanges tiles
// and panels through files in CB_RENDEZVOUS_DIR, so the host layer's multi-process paths (distribution, FullyDistVec pieces,
// SpMV, Reduce, ParallelWriteMM ordering) run on CPU too; cb_spmm_summa gathers the whole product's operands on every rank and
// computes the caller's block.
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <cmath>
#include <map>
#include <string>
#include <thread>
#include <vector>
#include "combblas_b200.h"
#include "../../combblas-spmm-test_b200/csrc/cb_gen500.cuh"       // host side of the Graph500 stream: the mock makes the same matrix as the device

struct cb_ctx { std::string err; int64_t launches = 0; int rank = 0, nranks = 1, pr = 1, pc = 1; };

// every rank contributes a byte string; afterwards everyone holds all of them (files "mock<seq>.r<rank>" in the rendezvous directory)
static void mock_allgatherv(const cb_ctx* c, const std::vector<unsigned char>& mine, std::vector<std::vector<unsigned char>>& all) {
    all.assign((size_t)c->nranks, std::vector<unsigned char>());
    if (c->nranks == 1) { all[0] = mine; return; }
    static long seq = 0;
    const char* d = std::getenv("CB_RENDEZVOUS_DIR");
    const std::string dir = d ? d : "/tmp/cb_mock_rdv";
    mkdir(dir.c_str(), 0700);
    const std::string base = dir + "/mock" + std::to_string(seq++) + ".r";
    {
        const std::string tmp = base + std::to_string(c->rank) + ".tmp", fin = base + std::to_string(c->rank);
        FILE* f = std::fopen(tmp.c_str(), "wb");
        const uint64_t n = mine.size();
        std::fwrite(&n, sizeof n, 1, f);
        if (n) std::fwrite(mine.data(), 1, mine.size(), f);
        std::fclose(f);
        std::rename(tmp.c_str(), fin.c_str());
    }
    for (int q = 0; q < c->nranks; ++q)
        for (int tries = 0;; ++tries) {
            FILE* f = std::fopen((base + std::to_string(q)).c_str(), "rb");
            if (f) {
                uint64_t n = 0;
                const bool ok = std::fread(&n, sizeof n, 1, f) == 1;
                all[(size_t)q].resize((size_t)n);
                const bool ok2 = ok && (n == 0 || std::fread(all[(size_t)q].data(), 1, (size_t)n, f) == n);
                std::fclose(f);
                if (ok2) break;
            }
            if (tries > 300000) { std::fprintf(stderr, "mock ABI: rank %d timed out waiting for rank %d\n", c->rank, q); std::exit(1); }
            std::this_thread::sleep_for(std::chrono::microseconds(200));
        }
}
template <class T>
static void put(std::vector<unsigned char>& b, const T& v) { const unsigned char* p = (const unsigned char*)&v; b.insert(b.end(), p, p + sizeof(T)); }
template <class T>
static T take(const std::vector<unsigned char>& b, size_t& off) { T v; std::memcpy(&v, b.data() + off, sizeof(T)); off += sizeof(T); return v; }
static void block_range(int64_t total, int nb, int b, int64_t* start, int64_t* len) {
    const int64_t per = total / nb;
    *start = per * b;
    *len = b == nb - 1 ? total - *start : per;
}
struct cb_tile {
    int64_t m = 0, n = 0;
    int val_dtype = CB_PATTERN;
    std::vector<int64_t> rowptr, col;
    std::vector<unsigned char> vals;       // nnz elements of val_dtype
    bool view = false;
};
struct cb_dense { int64_t rows = 0, cols = 0; int dtype = CB_F32; std::vector<unsigned char> data; };
struct cb_coo { int64_t m = 0, k = 0; int dtype = CB_F32; std::vector<int64_t> rows, cols; std::vector<unsigned char> vals; };

static std::string g_err;
static size_t esize(int dt) { return dt == CB_F32 || dt == CB_I32 ? 4 : dt == CB_F64 || dt == CB_I64 ? 8 : dt == CB_U8 ? 1 : 0; }
static int fail(cb_ctx* c, int st, const std::string& msg) { g_err = msg; if (c) c->err = msg; return st; }

static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

template <class T>
static void put_value(std::vector<unsigned char>& out, uint64_t h) {
    T v;
    if (std::is_floating_point<T>::value) v = (T)((double)((h >> 11) | 1) / 9007199254740992.0);
    else v = (T)(1 + (h >> 8) % 100);
    const unsigned char* p = reinterpret_cast<const unsigned char*>(&v);
    out.insert(out.end(), p, p + sizeof(T));
}

// rows of triples (r, c, value bytes) -> CSR with ascending columns
static void build_csr(cb_tile* t, std::vector<std::pair<std::pair<int64_t, int64_t>, std::vector<unsigned char>>>& trip) {
    std::sort(trip.begin(), trip.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    t->rowptr.assign((size_t)t->m + 1, 0);
    for (auto& e : trip) {
        ++t->rowptr[(size_t)e.first.first + 1];
        t->col.push_back(e.first.second);
        t->vals.insert(t->vals.end(), e.second.begin(), e.second.end());
    }
    for (int64_t r = 0; r < t->m; ++r) t->rowptr[(size_t)r + 1] += t->rowptr[(size_t)r];
}

template <class T> static T sr_id(int sr) { return sr == CB_MIN_PLUS ? std::numeric_limits<T>::max() : sr == CB_MAX_SEL2ND ? (T)-1 : (T)0; }
template <class T>
static T sr_add(int sr, T a, T b) {
    switch (sr) {
        case CB_MIN_PLUS: return std::min(a, b);
        case CB_MAX_SEL2ND: return std::max(a, b);
        case CB_OR_AND: return (T)((a != 0) || (b != 0));
        default: return (T)(a + b);
    }
}
template <class T>
static T sr_mul(int sr, bool has_a, T a, bool a_true, T x) {
    switch (sr) {
        case CB_MIN_PLUS: { const T inf = std::numeric_limits<T>::max(); return (a == inf || x == inf) ? inf : (T)(a + x); }
        case CB_MAX_SEL2ND: return x;
        case CB_OR_AND: return (T)(a_true && x != 0);
        default: return has_a ? (T)(a * x) : (a_true ? x : (T)0);
    }
}

template <class T>
static void multiply(const cb_tile* t, const cb_dense* X, cb_dense* Y, int sr, bool accumulate) {
    const T* x = reinterpret_cast<const T*>(X->data.data());
    T* y = reinterpret_cast<T*>(Y->data.data());
    const int64_t k = X->cols;
    const bool same = t->val_dtype == X->dtype && t->val_dtype != CB_U8;
    for (int64_t r = 0; r < t->m; ++r)
        for (int64_t j = 0; j < k; ++j) {
            bool first = true;
            T acc = sr_id<T>(sr);
            for (int64_t p = t->rowptr[(size_t)r]; p < t->rowptr[(size_t)r + 1]; ++p) {
                T a = T();
                bool a_true = true;
                if (same) std::memcpy(&a, t->vals.data() + (size_t)p * sizeof(T), sizeof(T));
                else if (t->val_dtype == CB_U8) a_true = t->vals[(size_t)p] != 0;
                const T prod = sr_mul<T>(sr, same, a, a_true, x[(size_t)t->col[(size_t)p] * (size_t)k + (size_t)j]);
                acc = first ? prod : sr_add<T>(sr, prod, acc);       // the first product is stored (mtSpGEMM.h:403-414)
                first = false;
            }
            T& out = y[(size_t)r * (size_t)k + (size_t)j];
            if (accumulate) { if (!first) out = sr_add<T>(sr, out, acc); }
            else out = acc;
        }
}

extern "C" {

int cb_abi_version(void) { return CB_ABI_VERSION; }
int cb_device_count(int* c) { *c = 1; return CB_OK; }
int cb_comm_unique_id(void* id) { std::memset(id, 0, 128); return CB_OK; }
int cb_ctx_create_grid(int, int rank, int nranks, int pr, int pc, const void*, cb_ctx** ctx) {
    if (pr * pc != nranks || rank < 0 || rank >= nranks) return fail(nullptr, CB_ERR_INVALIDPARAMS, "mock ABI: grid does not match the rank count");
    *ctx = new cb_ctx();
    (*ctx)->rank = rank; (*ctx)->nranks = nranks; (*ctx)->pr = pr; (*ctx)->pc = pc;
    return CB_OK;
}
int cb_ctx_create(int d, cb_ctx** ctx) { return cb_ctx_create_grid(d, 0, 1, 1, 1, nullptr, ctx); }
int cb_ctx_destroy(cb_ctx* c) { delete c; return CB_OK; }
int cb_ctx_grid(const cb_ctx* x, int* rank, int* pr, int* pc, int* r, int* c) {
    *rank = x->rank; *pr = x->pr; *pc = x->pc; *r = x->rank / x->pc; *c = x->rank % x->pc;
    return CB_OK;
}
int cb_ctx_sync(cb_ctx*) { return CB_OK; }
void* cb_ctx_stream(cb_ctx*) { return nullptr; }
const char* cb_last_error(const cb_ctx* c) { return c ? c->err.c_str() : g_err.c_str(); }
const char* cb_status_string(int s) { return s == CB_OK ? "ok" : "mock ABI error"; }
int cb_comm_allreduce_i64(cb_ctx* c, int which, int op, int64_t* inout, int count) {
    // which: 0 world, 1 my processor row, 2 my processor column; op: 0 sum, 1 max, 2 min
    std::vector<unsigned char> mine((const unsigned char*)inout, (const unsigned char*)(inout + count));
    std::vector<std::vector<unsigned char>> all;
    mock_allgatherv(c, mine, all);
    const int myrow = c->rank / c->pc, mycol = c->rank % c->pc;
    bool first = true;
    for (int q = 0; q < c->nranks; ++q) {
        if ((which == 1 && q / c->pc != myrow) || (which == 2 && q % c->pc != mycol)) continue;
        const int64_t* v = (const int64_t*)all[(size_t)q].data();
        for (int i = 0; i < count; ++i) inout[i] = first ? v[i] : op == 0 ? inout[i] + v[i] : op == 1 ? std::max(inout[i], v[i]) : std::min(inout[i], v[i]);
        first = false;
    }
    return CB_OK;
}
int64_t cb_launch_count(const cb_ctx* c) { return c->launches; }

int cb_tile_upload_csc(cb_ctx* ctx, int64_t m, int64_t n, int64_t nz, int64_t nzc, const void* cp, const void* jc, const void* ir,
                       const void* numx, int idt, int vdt, cb_tile** tile) {
    auto idx = [&](const void* a, int64_t i) { return idt == CB_I32 ? (int64_t)((const int32_t*)a)[i] : ((const int64_t*)a)[i]; };
    cb_tile* t = new cb_tile();
    t->m = m; t->n = n; t->val_dtype = vdt;
    const size_t es = esize(vdt == CB_PATTERN ? CB_U8 : vdt);
    std::vector<std::pair<std::pair<int64_t, int64_t>, std::vector<unsigned char>>> trip;
    const int64_t ncols = jc ? nzc : n;
    for (int64_t c = 0; c < ncols; ++c)
        for (int64_t p = idx(cp, c); p < idx(cp, c + 1); ++p) {
            std::vector<unsigned char> v;
            if (vdt != CB_PATTERN) v.assign((const unsigned char*)numx + (size_t)p * es, (const unsigned char*)numx + (size_t)(p + 1) * es);
            trip.push_back({{idx(ir, p), jc ? idx(jc, c) : c}, v});
        }
    if ((int64_t)trip.size() != nz) { delete t; return fail(ctx, CB_ERR_INVALIDPARAMS, "mock ABI: nz does not match the column pointers"); }
    build_csr(t, trip);
    *tile = t;
    return CB_OK;
}
int cb_tile_upload_coo(cb_ctx* ctx, int64_t m, int64_t n, int64_t nz, const void* rows, const void* cols, const void* vals, int idt, int vdt, cb_tile** tile) {
    auto idx = [&](const void* a, int64_t i) { return idt == CB_I32 ? (int64_t)((const int32_t*)a)[i] : ((const int64_t*)a)[i]; };
    cb_tile* t = new cb_tile();
    t->m = m; t->n = n; t->val_dtype = vdt;
    const size_t es = esize(vdt == CB_PATTERN ? CB_U8 : vdt);
    std::vector<std::pair<std::pair<int64_t, int64_t>, std::vector<unsigned char>>> trip;
    for (int64_t p = 0; p < nz; ++p) {
        if (idx(rows, p) < 0 || idx(rows, p) >= m || idx(cols, p) < 0 || idx(cols, p) >= n) { delete t; return fail(ctx, CB_ERR_INVALIDPARAMS, "mock ABI: index outside the tile"); }
        std::vector<unsigned char> v;
        if (vdt != CB_PATTERN) v.assign((const unsigned char*)vals + (size_t)p * es, (const unsigned char*)vals + (size_t)(p + 1) * es);
        trip.push_back({{idx(rows, p), idx(cols, p)}, v});
    }
    build_csr(t, trip);
    *tile = t;
    return CB_OK;
}
int cb_tile_from_distributed_coo(cb_ctx* ctx, int64_t gm, int64_t gn, int64_t nz, const int64_t* rows, const int64_t* cols, const void* vals,
                                 int vdt, int dup_op, cb_tile** tile) {
    // everyone publishes its triples; every rank keeps what it owns, merges duplicates in (rank, input) order and builds its tile
    const int pr = ctx->pr, pc = ctx->pc, myrow = ctx->rank / pc, mycol = ctx->rank % pc;
    const size_t es = vdt == CB_PATTERN ? 0 : esize(vdt);
    std::vector<unsigned char> mine;
    put<int64_t>(mine, nz);
    for (int64_t p = 0; p < nz; ++p) { put<int64_t>(mine, rows[p]); put<int64_t>(mine, cols[p]); }
    if (es) mine.insert(mine.end(), (const unsigned char*)vals, (const unsigned char*)vals + (size_t)nz * es);
    std::vector<std::vector<unsigned char>> all;
    mock_allgatherv(ctx, mine, all);
    int64_t r0, rl, c0, cl;
    block_range(gm, pr, myrow, &r0, &rl); block_range(gn, pc, mycol, &c0, &cl);
    std::map<std::pair<int64_t, int64_t>, std::vector<unsigned char>> acc;
    auto merge = [&](std::vector<unsigned char>& into, const unsigned char* v) {
        auto f = [&](auto tag) {
            typedef decltype(tag) T;
            T a, b;
            std::memcpy(&a, into.data(), sizeof(T)); std::memcpy(&b, v, sizeof(T));
            const T r = dup_op == 1 ? (T)(a + b) : dup_op == 2 ? std::max(a, b) : dup_op == 3 ? std::min(a, b) : a;
            std::memcpy(into.data(), &r, sizeof(T));
        };
        switch (vdt) { case CB_F32: f(float()); break; case CB_F64: f(double()); break; case CB_I32: f(int32_t()); break; case CB_I64: f(int64_t()); break; default: f(uint8_t()); }
    };
    for (int q = 0; q < ctx->nranks; ++q) {
        size_t off = 0;
        const std::vector<unsigned char>& b = all[(size_t)q];
        const int64_t n = take<int64_t>(b, off);
        const size_t voff = off + (size_t)n * 16;
        for (int64_t p = 0; p < n; ++p) {
            const int64_t r = take<int64_t>(b, off), c = take<int64_t>(b, off);
            if (r < 0 || r >= gm || c < 0 || c >= gn) return fail(ctx, CB_ERR_INVALIDPARAMS, "mock ABI: triple outside the matrix");
            if (r < r0 || r >= r0 + rl || c < c0 || c >= c0 + cl) continue;
            const unsigned char* v = es ? b.data() + voff + (size_t)p * es : nullptr;
            auto it = acc.find({r - r0, c - c0});
            if (it == acc.end()) acc[{r - r0, c - c0}] = es ? std::vector<unsigned char>(v, v + es) : std::vector<unsigned char>();
            else if (es) merge(it->second, v);
        }
    }
    cb_tile* t = new cb_tile();
    t->m = rl; t->n = cl; t->val_dtype = vdt;
    std::vector<std::pair<std::pair<int64_t, int64_t>, std::vector<unsigned char>>> trip(acc.begin(), acc.end());
    build_csr(t, trip);
    *tile = t;
    return CB_OK;
}
int cb_tile_from_mm_text(cb_ctx* ctx, int64_t gm, int64_t gn, const char* text, int64_t nbytes, int flags, int vdt, int dup_op, cb_tile** tile) {
    // host restatement of the device parser's contract: lines -> triples (strtoll / strtod), numbers with more than 19 significant
    // digits refused on every rank, then the distributed ingestion above
    std::vector<int64_t> rows, cols;
    std::vector<double> vv;
    int64_t hard = 0;
    const char* p = text;
    const char* end = text + nbytes;
    while (p < end) {
        const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
        const std::string line(p, nl ? nl : end);
        p = nl ? nl + 1 : end;
        char* q = nullptr;
        const char* c = line.c_str();
        const long long ii = std::strtoll(c, &q, 10);
        if (q == c) continue;
        const char* c2 = q;
        const long long jj = std::strtoll(c2, &q, 10);
        if (q == c2) continue;
        double v = 1.0;
        if (!(flags & 2)) {
            const char* c3 = q;
            v = std::strtod(c3, &q);
            if (q == c3) { ++hard; continue; }
            int sig = 0;
            bool nz = false;
            for (const char* d = c3; d < q && *d != 'e' && *d != 'E'; ++d)
                if (*d >= '0' && *d <= '9' && (nz || *d != '0')) { nz = true; ++sig; }
            if (sig > 19 || std::isinf(v) || std::isnan(v)) { ++hard; continue; }
        }
        const int64_t r = ii - ((flags & 1) ? 1 : 0), cc = jj - ((flags & 1) ? 1 : 0);
        rows.push_back(r); cols.push_back(cc); vv.push_back(v);
        if ((flags & 4) && r != cc) { rows.push_back(cc); cols.push_back(r); vv.push_back(v); }
    }
    std::vector<unsigned char> mine;
    put<int64_t>(mine, hard);
    std::vector<std::vector<unsigned char>> all;
    mock_allgatherv(ctx, mine, all);
    for (const auto& b : all) { size_t off = 0; if (take<int64_t>(b, off) != 0) return fail(ctx, CB_ERR_UNSUPPORTED, "mock ABI: numbers outside the device parser's exact range"); }
    const size_t es = vdt == CB_PATTERN ? 0 : esize(vdt);
    std::vector<unsigned char> vals(rows.size() * es);
    for (size_t k = 0; k < rows.size() && es; ++k) {
        switch (vdt) {
            case CB_F32: { float x = (float)vv[k]; std::memcpy(&vals[k * es], &x, es); break; }
            case CB_F64: { double x = vv[k]; std::memcpy(&vals[k * es], &x, es); break; }
            case CB_I32: { int32_t x = (int32_t)vv[k]; std::memcpy(&vals[k * es], &x, es); break; }
            case CB_I64: { int64_t x = (int64_t)vv[k]; std::memcpy(&vals[k * es], &x, es); break; }
            default: { uint8_t x = vv[k] != 0.0; std::memcpy(&vals[k * es], &x, es); break; }
        }
    }
    return cb_tile_from_distributed_coo(ctx, gm, gn, (int64_t)rows.size(), rows.data(), cols.data(), es ? (const void*)vals.data() : nullptr, vdt, dup_op, tile);
}
int cb_tile_free(cb_tile* t) { delete t; return CB_OK; }
int cb_tile_info(const cb_tile* t, int64_t info[8]) {
    std::memset(info, 0, 8 * sizeof(int64_t));
    info[0] = (int64_t)t->col.size(); info[1] = t->m; info[2] = t->n;
    return CB_OK;
}
int cb_tile_pattern_view(const cb_tile* t, cb_tile** view) {
    cb_tile* v = new cb_tile(*t);
    v->val_dtype = CB_PATTERN; v->vals.clear(); v->view = true;
    *view = v;
    return CB_OK;
}
int cb_tile_filter_columns(cb_ctx*, const cb_tile* t, const uint8_t* keep, cb_tile** out) {
    cb_tile* f = new cb_tile();
    f->m = t->m; f->n = t->n; f->val_dtype = t->val_dtype;
    const size_t es = esize(t->val_dtype == CB_PATTERN ? CB_U8 : t->val_dtype);
    std::vector<std::pair<std::pair<int64_t, int64_t>, std::vector<unsigned char>>> trip;
    for (int64_t r = 0; r < t->m; ++r)
        for (int64_t p = t->rowptr[(size_t)r]; p < t->rowptr[(size_t)r + 1]; ++p)
            if (keep[t->col[(size_t)p]]) {
                std::vector<unsigned char> v;
                if (t->val_dtype != CB_PATTERN) v.assign(t->vals.begin() + (size_t)p * es, t->vals.begin() + (size_t)(p + 1) * es);
                trip.push_back({{r, t->col[(size_t)p]}, v});
            }
    build_csr(f, trip);
    *out = f;
    return CB_OK;
}
int cb_tile_download_csr(cb_tile* t, int64_t* rowptr, int64_t* colidx, void* vals) {
    if (rowptr) std::copy(t->rowptr.begin(), t->rowptr.end(), rowptr);
    if (colidx) std::copy(t->col.begin(), t->col.end(), colidx);
    if (vals && !t->vals.empty()) std::memcpy(vals, t->vals.data(), t->vals.size());
    return CB_OK;
}
int cb_gen_rmat_tile(cb_ctx*, int scale, int edgefactor, uint64_t seed, const double*, int symmetric, int64_t row0, int64_t m, int64_t col0,
                     int64_t n, int val_dtype, uint64_t val_seed, cb_tile** tile) {
    // NOT the product's generator: any skewed random graph will do for host-logic tests
    const int64_t N = (int64_t)1 << scale;
    std::map<std::pair<int64_t, int64_t>, int> seen;
    for (int64_t e = 0; e < (int64_t)edgefactor * N; ++e) {
        const uint64_t h = splitmix64(seed * 0x100000001B3ULL ^ (uint64_t)e);
        int64_t i = (int64_t)((h & 0xffffffffu) % (uint64_t)N), j = (int64_t)((h >> 32) % (uint64_t)N);
        if (h & 1) i = i % std::max<int64_t>(1, N / 8);      // a few heavy rows
        if (i == j) continue;
        seen[{i, j}] = 1;
        if (symmetric) seen[{j, i}] = 1;
    }
    cb_tile* t = new cb_tile();
    t->m = m; t->n = n; t->val_dtype = val_dtype;
    std::vector<std::pair<std::pair<int64_t, int64_t>, std::vector<unsigned char>>> trip;
    for (auto& kv : seen) {
        const int64_t i = kv.first.first, j = kv.first.second;
        if (i < row0 || i >= row0 + m || j < col0 || j >= col0 + n) continue;
        std::vector<unsigned char> v;
        const uint64_t h = splitmix64(val_seed * 0x100000001B3ULL ^ (uint64_t)(i * N + j));
        switch (val_dtype) {
            case CB_F32: put_value<float>(v, h); break;
            case CB_F64: put_value<double>(v, h); break;
            case CB_I32: put_value<int32_t>(v, h); break;
            case CB_I64: put_value<int64_t>(v, h); break;
            case CB_U8: v.push_back(1); break;
            default: break;
        }
        trip.push_back({{i - row0, j - col0}, v});
    }
    build_csr(t, trip);
    *tile = t;
    return CB_OK;
}

int cb_gen_graph500_edges(cb_ctx*, int lgN, uint64_t u1, uint64_t u2, int64_t first, int64_t count, int64_t* src, int64_t* dst) {
    static g500::Tables t;
    g500::build_tables(u1, u2, &t);
    for (int64_t q = 0; q < count; ++q) {
        const uint64_t ei = (uint64_t)(first + q);
        g500::State s = t.seed;
        for (int b = 0; b < 4; ++b) { const unsigned v = (unsigned)((ei >> (8 * b)) & 0xFF); if (v) s = g500::apply(t.edge[b][v], s); }
        uint64_t a, c;
        g500::one_edge(s, lgN, t.val0, t.val1, &a, &c);
        src[q] = (int64_t)a; dst[q] = (int64_t)c;
    }
    return CB_OK;
}
int cb_gen_graph500_tile(cb_ctx* ctx, int scale, int edgefactor, uint64_t u1, uint64_t u2, int symmetric, int remove_loops, int64_t row0, int64_t m,
                         int64_t col0, int64_t n, int val_dtype, uint64_t val_seed, cb_tile** tile) {
    const int64_t nedges = (int64_t)edgefactor << scale, N = (int64_t)1 << scale;
    std::vector<int64_t> src((size_t)nedges), dst((size_t)nedges);
    cb_gen_graph500_edges(ctx, scale, u1, u2, 0, nedges, src.data(), dst.data());
    std::map<std::pair<int64_t, int64_t>, int64_t> mult;
    for (int64_t e = 0; e < nedges; ++e) {
        if (remove_loops && src[(size_t)e] == dst[(size_t)e]) continue;
        ++mult[{src[(size_t)e], dst[(size_t)e]}];
        if (symmetric) ++mult[{dst[(size_t)e], src[(size_t)e]}];
    }
    cb_tile* t = new cb_tile();
    t->m = m; t->n = n; t->val_dtype = val_dtype;
    std::vector<std::pair<std::pair<int64_t, int64_t>, std::vector<unsigned char>>> trip;
    for (auto& kv : mult) {
        const int64_t i = kv.first.first, j = kv.first.second;
        if (i < row0 || i >= row0 + m || j < col0 || j >= col0 + n) continue;
        std::vector<unsigned char> v;
        auto putv = [&](auto x) { const unsigned char* p = (const unsigned char*)&x; v.insert(v.end(), p, p + sizeof x); };
        if (val_seed == 0) {
            switch (val_dtype) {
                case CB_F32: putv((float)kv.second); break;
                case CB_F64: putv((double)kv.second); break;
                case CB_I32: putv((int32_t)kv.second); break;
                case CB_I64: putv((int64_t)kv.second); break;
                case CB_U8: putv((uint8_t)kv.second); break;
                default: break;
            }
        } else {
            const uint64_t h = splitmix64(val_seed * 0x100000001B3ULL ^ (uint64_t)(i * N + j));
            switch (val_dtype) {
                case CB_F32: put_value<float>(v, h); break;
                case CB_F64: put_value<double>(v, h); break;
                case CB_I32: put_value<int32_t>(v, h); break;
                case CB_I64: put_value<int64_t>(v, h); break;
                case CB_U8: v.push_back(1); break;
                default: break;
            }
        }
        trip.push_back({{i - row0, j - col0}, v});
    }
    build_csr(t, trip);
    *tile = t;
    return CB_OK;
}

int cb_dense_alloc(cb_ctx*, int64_t rows, int64_t cols, int dtype, cb_dense** d) {
    cb_dense* x = new cb_dense();
    x->rows = rows; x->cols = cols; x->dtype = dtype;
    x->data.assign((size_t)rows * (size_t)cols * esize(dtype), 0);
    *d = x;
    return CB_OK;
}
int cb_dense_free(cb_dense* d) { delete d; return CB_OK; }
int cb_dense_upload(cb_dense* d, const void* host, int64_t ld) {
    const size_t es = esize(d->dtype);
    for (int64_t r = 0; r < d->rows; ++r) std::memcpy(d->data.data() + (size_t)r * (size_t)d->cols * es, (const char*)host + (size_t)r * (size_t)ld * es, (size_t)d->cols * es);
    return CB_OK;
}
int cb_dense_download(cb_dense* d, void* host, int64_t ld) {
    const size_t es = esize(d->dtype);
    for (int64_t r = 0; r < d->rows; ++r) std::memcpy((char*)host + (size_t)r * (size_t)ld * es, d->data.data() + (size_t)r * (size_t)d->cols * es, (size_t)d->cols * es);
    return CB_OK;
}

int cb_spmm_local(cb_ctx* ctx, const cb_tile* t, const cb_dense* X, cb_dense* Y, int sr, int accumulate) {
    if (X->rows != t->n || Y->rows != t->m || X->cols != Y->cols) return fail(ctx, CB_ERR_DIMMISMATCH, "mock ABI: dimension mismatch");
    if (sr == CB_PLUS_TIMES && X->dtype == CB_U8) sr = CB_OR_AND;
    ++ctx->launches;
    switch (X->dtype) {
        case CB_F32: multiply<float>(t, X, Y, sr, accumulate != 0); break;
        case CB_F64: multiply<double>(t, X, Y, sr, accumulate != 0); break;
        case CB_I32: multiply<int32_t>(t, X, Y, sr, accumulate != 0); break;
        case CB_I64: multiply<int64_t>(t, X, Y, sr, accumulate != 0); break;
        case CB_U8: multiply<uint8_t>(t, X, Y, sr, accumulate != 0); break;
        default: return fail(ctx, CB_ERR_UNSUPPORTED, "mock ABI: dtype");
    }
    return CB_OK;
}
int cb_spmm_summa(cb_ctx* ctx, const cb_tile* t, const cb_dense* X, cb_dense* Y, int sr, int64_t gm, int64_t gn, int64_t gk) {
    if (ctx->nranks == 1) return cb_spmm_local(ctx, t, X, Y, sr, 0);
    // every rank publishes its A tile (global coordinates) and its X tile; everyone rebuilds block-row myprocrow of A and
    // block-column myproccol of X and multiplies them - the result the SUMMA stage loop accumulates
    const int pr = ctx->pr, pc = ctx->pc, myrow = ctx->rank / pc, mycol = ctx->rank % pc;
    int64_t r0, rl, c0, cl, x0, xl, k0, kl;
    block_range(gm, pr, myrow, &r0, &rl); block_range(gn, pc, mycol, &c0, &cl);
    block_range(gn, pr, myrow, &x0, &xl); block_range(gk, pc, mycol, &k0, &kl);
    if (t->m != rl || t->n != cl || X->rows != xl || X->cols != kl || Y->rows != rl || Y->cols != kl) return fail(ctx, CB_ERR_DIMMISMATCH, "mock ABI: summa block mismatch");
    const size_t ves = esize(t->val_dtype == CB_PATTERN ? CB_U8 : t->val_dtype), xes = esize(X->dtype);
    std::vector<unsigned char> mine;
    put<int64_t>(mine, (int64_t)t->col.size());
    for (int64_t r = 0; r < t->m; ++r)
        for (int64_t p = t->rowptr[(size_t)r]; p < t->rowptr[(size_t)r + 1]; ++p) { put<int64_t>(mine, r0 + r); put<int64_t>(mine, c0 + t->col[(size_t)p]); }
    if (t->val_dtype != CB_PATTERN) mine.insert(mine.end(), t->vals.begin(), t->vals.end());
    put<int64_t>(mine, X->rows); put<int64_t>(mine, X->cols);
    mine.insert(mine.end(), X->data.begin(), X->data.end());
    std::vector<std::vector<unsigned char>> all;
    mock_allgatherv(ctx, mine, all);
    cb_tile arow;                     // block-row myrow of A, all columns
    arow.m = rl; arow.n = gn; arow.val_dtype = t->val_dtype;
    cb_dense xcol;                    // all rows of X, block-column mycol
    xcol.rows = gn; xcol.cols = kl; xcol.dtype = X->dtype;
    xcol.data.assign((size_t)gn * (size_t)kl * xes, 0);
    std::vector<std::pair<std::pair<int64_t, int64_t>, std::vector<unsigned char>>> trip;
    for (int q = 0; q < ctx->nranks; ++q) {
        const std::vector<unsigned char>& b = all[(size_t)q];
        size_t off = 0;
        const int64_t nz = take<int64_t>(b, off);
        std::vector<std::pair<int64_t, int64_t>> rc((size_t)nz);
        for (int64_t p = 0; p < nz; ++p) { rc[(size_t)p].first = take<int64_t>(b, off); rc[(size_t)p].second = take<int64_t>(b, off); }
        for (int64_t p = 0; p < nz; ++p) {
            std::vector<unsigned char> v;
            if (t->val_dtype != CB_PATTERN) { v.assign(b.begin() + off, b.begin() + off + ves); off += ves; }
            if (q / pc == myrow) trip.push_back({{rc[(size_t)p].first - r0, rc[(size_t)p].second}, v});
        }
        const int64_t qr = take<int64_t>(b, off), qc = take<int64_t>(b, off);
        if (q % pc == mycol) {
            int64_t q0, ql;
            block_range(gn, pr, q / pc, &q0, &ql);
            if (qr != ql || qc != kl) return fail(ctx, CB_ERR_DIMMISMATCH, "mock ABI: X tile of a peer has the wrong shape");
            if (qr > 0 && qc > 0) std::memcpy(xcol.data.data() + (size_t)q0 * (size_t)kl * xes, b.data() + off, (size_t)qr * (size_t)qc * xes);
        }
    }
    build_csr(&arow, trip);
    return cb_spmm_local(ctx, &arow, &xcol, Y, sr, 0);
}
}  // extern "C"
// sparse x sparse: C(i,j) = (+)_kk A(i,kk) (x) B(kk,j), products folded in ascending kk, first product stored; column-major result
template <class T>
static void spgemm_typed(const cb_tile* A, const cb_tile* B, int sr, cb_coo* C) {
    std::map<std::pair<int64_t, int64_t>, T> acc;            // (column, row) -> value: iteration order is column-major
    const bool same = A->val_dtype == C->dtype && !(A->val_dtype == CB_U8 && C->dtype != CB_U8);
    for (int64_t i = 0; i < A->m; ++i)
        for (int64_t p = A->rowptr[(size_t)i]; p < A->rowptr[(size_t)i + 1]; ++p) {
            const int64_t kk = A->col[(size_t)p];
            T a = T();
            bool a_true = true;
            if (same) std::memcpy(&a, A->vals.data() + (size_t)p * sizeof(T), sizeof(T));
            else if (A->val_dtype == CB_U8) a_true = A->vals[(size_t)p] != 0;
            if (sr == CB_OR_AND && same) a_true = a != 0;
            for (int64_t q = B->rowptr[(size_t)kk]; q < B->rowptr[(size_t)kk + 1]; ++q) {
                T b = (T)1;
                if (B->val_dtype != CB_PATTERN) std::memcpy(&b, B->vals.data() + (size_t)q * sizeof(T), sizeof(T));
                const T prod = sr_mul<T>(sr, same, a, a_true, b);
                auto it = acc.find({B->col[(size_t)q], i});
                if (it == acc.end()) acc[{B->col[(size_t)q], i}] = prod;
                else it->second = sr_add<T>(sr, prod, it->second);
            }
        }
    for (auto& kv : acc) {
        C->cols.push_back(kv.first.first);
        C->rows.push_back(kv.first.second);
        const unsigned char* v = (const unsigned char*)&kv.second;
        C->vals.insert(C->vals.end(), v, v + sizeof(T));
    }
}
extern "C" {
int cb_spgemm_local(cb_ctx* ctx, const cb_tile* A, const cb_tile* B, int sr, int dtype, cb_coo** out) {
    if (A->n != B->m) return fail(ctx, CB_ERR_DIMMISMATCH, "mock ABI: spgemm dimension mismatch");
    if (B->val_dtype != dtype && B->val_dtype != CB_PATTERN) return fail(ctx, CB_ERR_UNSUPPORTED, "mock ABI: B must hold the product's type");
    if (sr == CB_PLUS_TIMES && dtype == CB_U8) sr = CB_OR_AND;
    ++ctx->launches;
    cb_coo* C = new cb_coo();
    C->m = A->m; C->k = B->n; C->dtype = dtype;
    switch (dtype) {
        case CB_F32: spgemm_typed<float>(A, B, sr, C); break;
        case CB_F64: spgemm_typed<double>(A, B, sr, C); break;
        case CB_I32: spgemm_typed<int32_t>(A, B, sr, C); break;
        case CB_I64: spgemm_typed<int64_t>(A, B, sr, C); break;
        case CB_U8: spgemm_typed<uint8_t>(A, B, sr, C); break;
        default: delete C; return fail(ctx, CB_ERR_UNSUPPORTED, "mock ABI: dtype");
    }
    *out = C;
    return CB_OK;
}
// a tile's triples in global coordinates, appended to a byte string / read back
static void pack_tile(std::vector<unsigned char>& b, const cb_tile* t, int64_t r0, int64_t c0) {
    put<int64_t>(b, (int64_t)t->col.size());
    put<int64_t>(b, (int64_t)t->val_dtype);
    for (int64_t r = 0; r < t->m; ++r)
        for (int64_t p = t->rowptr[(size_t)r]; p < t->rowptr[(size_t)r + 1]; ++p) { put<int64_t>(b, r0 + r); put<int64_t>(b, c0 + t->col[(size_t)p]); }
    b.insert(b.end(), t->vals.begin(), t->vals.end());
}
typedef std::vector<std::pair<std::pair<int64_t, int64_t>, std::vector<unsigned char>>> Trips;
static void unpack_tile(const std::vector<unsigned char>& b, size_t& off, bool keep, int64_t roff, int64_t coff, Trips& trip) {
    const int64_t nz = take<int64_t>(b, off);
    const int vdt = (int)take<int64_t>(b, off);
    const size_t ves = vdt == CB_PATTERN ? 0 : esize(vdt);
    std::vector<std::pair<int64_t, int64_t>> rc((size_t)nz);
    for (int64_t p = 0; p < nz; ++p) { rc[(size_t)p].first = take<int64_t>(b, off); rc[(size_t)p].second = take<int64_t>(b, off); }
    for (int64_t p = 0; p < nz; ++p) {
        std::vector<unsigned char> v(b.begin() + off, b.begin() + off + ves);
        off += ves;
        if (keep) trip.push_back({{rc[(size_t)p].first - roff, rc[(size_t)p].second - coff}, v});
    }
}
int cb_spgemm_summa(cb_ctx* ctx, const cb_tile* A, const cb_tile* B, int sr, int dtype, int64_t gm, int64_t gn, int64_t gk, cb_coo** out) {
    if (ctx->nranks == 1) return cb_spgemm_local(ctx, A, B, sr, dtype, out);
    const int pr = ctx->pr, pc = ctx->pc, myrow = ctx->rank / pc, mycol = ctx->rank % pc;
    int64_t r0, rl, c0, cl, x0, xl, k0, kl;
    block_range(gm, pr, myrow, &r0, &rl); block_range(gn, pc, mycol, &c0, &cl);
    block_range(gn, pr, myrow, &x0, &xl); block_range(gk, pc, mycol, &k0, &kl);
    if (A->m != rl || A->n != cl || B->m != xl || B->n != kl) return fail(ctx, CB_ERR_DIMMISMATCH, "mock ABI: spgemm summa block mismatch");
    std::vector<unsigned char> mine;
    pack_tile(mine, A, r0, c0);
    pack_tile(mine, B, x0, k0);
    std::vector<std::vector<unsigned char>> all;
    mock_allgatherv(ctx, mine, all);
    cb_tile arow, bcol;             // block-row myrow of A (all columns); all rows of B, block-column mycol
    arow.m = rl; arow.n = gn; arow.val_dtype = A->val_dtype;
    bcol.m = gn; bcol.n = kl; bcol.val_dtype = B->val_dtype;
    Trips ta, tb;
    for (int q = 0; q < ctx->nranks; ++q) {
        size_t off = 0;
        unpack_tile(all[(size_t)q], off, q / pc == myrow, r0, 0, ta);
        unpack_tile(all[(size_t)q], off, q % pc == mycol, 0, k0, tb);
    }
    build_csr(&arow, ta);
    build_csr(&bcol, tb);
    return cb_spgemm_local(ctx, &arow, &bcol, sr, dtype, out);
}
int cb_coo_info(const cb_coo* c, int64_t* nnz, int64_t* m, int64_t* k, int* dtype) {
    if (nnz) *nnz = (int64_t)c->rows.size();
    if (m) *m = c->m;
    if (k) *k = c->k;
    if (dtype) *dtype = c->dtype;
    return CB_OK;
}
int cb_coo_download(cb_coo* c, int64_t* rows, int64_t* cols, void* vals) {
    if (rows) std::copy(c->rows.begin(), c->rows.end(), rows);
    if (cols) std::copy(c->cols.begin(), c->cols.end(), cols);
    if (vals && !c->vals.empty()) std::memcpy(vals, c->vals.data(), c->vals.size());
    return CB_OK;
}
int cb_coo_free(cb_coo* c) { delete c; return CB_OK; }
int cb_spmv_grid(cb_ctx* ctx, const cb_tile* t, const void* x_piece, int64_t x_off, int64_t x_len, void* y_piece, int64_t y_off, int64_t y_len,
                 int sr, int dtype, int64_t gm, int64_t gn) {
    // every rank publishes its piece of x; everyone assembles x, multiplies its tile with its column block, publishes the
    // partial result and folds the partials of its processor row (starting from SR::id() like the reference's y)
    const int pr = ctx->pr, pc = ctx->pc, myrow = ctx->rank / pc, mycol = ctx->rank % pc;
    const size_t es = esize(dtype);
    int64_t r0, rl, c0, cl;
    block_range(gm, pr, myrow, &r0, &rl); block_range(gn, pc, mycol, &c0, &cl);
    if (t->m != rl || t->n != cl) return fail(ctx, CB_ERR_DIMMISMATCH, "mock ABI: spmv block mismatch");
    std::vector<unsigned char> mine;
    put<int64_t>(mine, x_off); put<int64_t>(mine, x_len);
    mine.insert(mine.end(), (const unsigned char*)x_piece, (const unsigned char*)x_piece + (size_t)x_len * es);
    std::vector<std::vector<unsigned char>> all;
    mock_allgatherv(ctx, mine, all);
    std::vector<unsigned char> xfull((size_t)gn * es, 0);
    for (int q = 0; q < ctx->nranks; ++q) {
        size_t off = 0;
        const int64_t o = take<int64_t>(all[(size_t)q], off), l = take<int64_t>(all[(size_t)q], off);
        if (l > 0) std::memcpy(xfull.data() + (size_t)o * es, all[(size_t)q].data() + off, (size_t)l * es);
    }
    cb_dense *X = nullptr, *Y = nullptr;
    cb_dense_alloc(ctx, cl, 1, dtype, &X);
    cb_dense_alloc(ctx, rl, 1, dtype, &Y);
    if (cl > 0) std::memcpy(X->data.data(), xfull.data() + (size_t)c0 * es, (size_t)cl * es);
    int st = cb_spmm_local(ctx, t, X, Y, sr, 0);
    if (st != CB_OK) { cb_dense_free(X); cb_dense_free(Y); return st; }
    std::vector<std::vector<unsigned char>> parts;
    mock_allgatherv(ctx, Y->data, parts);
    if (sr == CB_PLUS_TIMES && dtype == CB_U8) sr = CB_OR_AND;
    auto fold = [&](auto tag) {
        typedef decltype(tag) T;
        std::vector<T> acc((size_t)rl, sr_id<T>(sr));
        for (int q = 0; q < ctx->nranks; ++q) {
            if (q / pc != myrow) continue;
            const T* v = (const T*)parts[(size_t)q].data();
            for (int64_t i = 0; i < rl; ++i) acc[(size_t)i] = sr_add<T>(sr, acc[(size_t)i], v[i]);
        }
        if (y_len > 0) std::memcpy(y_piece, acc.data() + (y_off - r0), (size_t)y_len * sizeof(T));
    };
    switch (dtype) {
        case CB_F32: fold(float()); break;
        case CB_F64: fold(double()); break;
        case CB_I32: fold(int32_t()); break;
        case CB_I64: fold(int64_t()); break;
        default: fold(uint8_t()); break;
    }
    cb_dense_free(X); cb_dense_free(Y);
    return CB_OK;
}
int cb_spmm_summa_host(cb_ctx* ctx, const cb_tile* t, const void* X_host, int64_t ldx, void* Y_host, int64_t ldy, int sr,
                       int64_t gm, int64_t gn, int64_t gk, int dtype) {
    const int pr = ctx->pr, pc = ctx->pc, myrow = ctx->rank / pc, mycol = ctx->rank % pc;
    int64_t r0, rl, x0, xl, k0, kl;
    block_range(gm, pr, myrow, &r0, &rl); block_range(gn, pr, myrow, &x0, &xl); block_range(gk, pc, mycol, &k0, &kl);
    cb_dense *X = nullptr, *Y = nullptr;
    cb_dense_alloc(ctx, ctx->nranks == 1 ? gn : xl, kl, dtype, &X);
    cb_dense_alloc(ctx, rl, kl, dtype, &Y);
    if (kl > 0 && X->rows > 0) cb_dense_upload(X, X_host, ldx);
    const int st = cb_spmm_summa(ctx, t, X, Y, sr, gm, gn, gk);
    if (st == CB_OK && kl > 0 && rl > 0) cb_dense_download(Y, Y_host, ldy);
    cb_dense_free(X); cb_dense_free(Y);
    return st;
}

}  // extern "C"
