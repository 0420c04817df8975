// Process-per-rank MPI stand-in (see mpi.h in this directory).  TEST INFRASTRUCTURE ONLY.
//
// Launcher: `main` below maps one shared region, forks CBMPI_NP ranks and calls the program's own main, which the
// build renames with -Dmain=cb_rank_main, in each child.  A rank that exits non-zero, aborts or dies takes the job down.
// Environment: CBMPI_NP (ranks, default 1), CBMPI_ARENA_MB (staging bytes per rank, default 256),
// CBMPI_HEAP_MB (eager point-to-point bytes per rank, default 64), CBMPI_TIMEOUT (seconds, default 900).
#ifdef main
#undef main                                    // this file holds the launcher's real main
#endif
#include "mpi.h"
#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <signal.h>
#include <stdarg.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>
#include <algorithm>
#include <vector>

int cb_rank_main(int argc, char** argv);        // the program's main, renamed by -Dmain=cb_rank_main

namespace {

const int MAXP = 64, MAXC = 16384, QLEN = 256;

struct Pub { volatile long long off, a, b; };
struct Comm {
    volatile int size;
    int members[MAXP];                           // world ranks, in communicator order
    volatile int bar_count, bar_gen;
    Pub pub[MAXP];
};
struct Msg { volatile int state; int comm, tag; long long bytes, off; };   // state: 0 free, 1 posted, 2 consumed out of order
struct Box { Msg q[QLEN]; volatile unsigned head, tail; };                   // head: sender only, tail: receiver only
struct Shared {
    int np;
    volatile int abort_flag, abort_code, next_comm;
    size_t arena_bytes, heap_bytes;
    volatile long long heap_outstanding[MAXP];
    Comm comms[MAXC];
    Box box[MAXP][MAXP];                         // [source][destination]
};

Shared* S = nullptr;
char* g_base = nullptr;
int g_rank = 0;
int g_finalized = 0;
Comm g_self;                                      // MPI_COMM_SELF
long long g_heap_top = 0;                         // bump pointer of this rank's eager heap
std::vector<std::vector<int>> g_groups(1);        // group handle -> world ranks (handle 0 = null)
std::vector<MPI_User_function*> g_ops;
struct Req { int kind; void* buf; long long bytes; int src, tag, comm; };    // kind 0 = complete, 1 = pending receive
std::vector<Req> g_reqs(1);

[[noreturn]] void die(const char* what) {
    fprintf(stderr, "cbmpi[rank %d]: %s\n", g_rank, what);
    if (S) { S->abort_code = 86; __atomic_store_n(&S->abort_flag, 1, __ATOMIC_SEQ_CST); }
    _exit(86);
}
inline void relax(unsigned& spins) {
    if (++spins < 64) { __builtin_ia32_pause(); return; }
    sched_yield();
    if ((spins & 255) == 0) {
        if (__atomic_load_n(&S->abort_flag, __ATOMIC_SEQ_CST)) _exit(S->abort_code ? S->abort_code : 1);
        if (getppid() == 1) _exit(1);             // launcher is gone
    }
    if (spins > 4096) usleep(100);
}
inline char* arena(int wr) { return g_base + (size_t)wr * (S->arena_bytes + S->heap_bytes); }
inline char* heap(int wr) { return arena(wr) + S->arena_bytes; }

struct Ctx { Comm* c; int n, me, handle; };
Ctx ctx(MPI_Comm h) {
    if (h == MPI_COMM_SELF) return Ctx{&g_self, 1, 0, h};
    if (h <= 0 || h >= MAXC || S->comms[h].size == 0) die("invalid communicator");
    Comm* c = &S->comms[h];
    for (int i = 0; i < c->size; ++i)
        if (c->members[i] == g_rank) return Ctx{c, c->size, i, h};
    die("calling rank is not a member of the communicator");
}
void barrier(const Ctx& x) {
    if (x.n == 1) return;
    Comm* c = x.c;
    const int gen = __atomic_load_n(&c->bar_gen, __ATOMIC_SEQ_CST);
    if (__atomic_add_fetch(&c->bar_count, 1, __ATOMIC_SEQ_CST) == x.n) {
        __atomic_store_n(&c->bar_count, 0, __ATOMIC_SEQ_CST);
        __atomic_add_fetch(&c->bar_gen, 1, __ATOMIC_SEQ_CST);
    } else {
        unsigned spins = 0;
        while (__atomic_load_n(&c->bar_gen, __ATOMIC_SEQ_CST) == gen) relax(spins);
    }
}
void stage(const void* src, size_t bytes, size_t at = 0) {
    if (at + bytes > S->arena_bytes) die("collective payload exceeds CBMPI_ARENA_MB");
    if (bytes) memcpy(arena(g_rank) + at, src, bytes);
}
inline const char* staged(const Ctx& x, int idx, size_t at = 0) { return arena(x.c->members[idx]) + at; }

template <class T>
void fold_typed(int op, const T* in, T* io, long long n) {
    for (long long i = 0; i < n; ++i) {
        const T a = in[i], b = io[i];
        switch (op) {
            case MPI_SUM: io[i] = (T)(a + b); break;
            case MPI_PROD: io[i] = (T)(a * b); break;
            case MPI_MAX: io[i] = a > b ? a : b; break;
            case MPI_MIN: io[i] = a < b ? a : b; break;
            case MPI_LAND: io[i] = (T)((a != 0) && (b != 0)); break;
            case MPI_LOR: io[i] = (T)((a != 0) || (b != 0)); break;
            case MPI_LXOR: io[i] = (T)((a != 0) != (b != 0)); break;
            default: die("operation not defined for this datatype");
        }
    }
}
template <class T>
void fold_bits(int op, const T* in, T* io, long long n) {
    for (long long i = 0; i < n; ++i)
        io[i] = op == MPI_BAND ? (T)(in[i] & io[i]) : op == MPI_BOR ? (T)(in[i] | io[i]) : (T)(in[i] ^ io[i]);
}
// io = in (op) io, element-wise: the contract of an MPI_User_function
void fold(MPI_Op op, MPI_Datatype t, const void* in, void* io, int count) {
    if (op >= CBMPI_FIRST_USER_OP) {
        const size_t k = (size_t)(op - CBMPI_FIRST_USER_OP);
        if (k >= g_ops.size() || !g_ops[k]) die("unknown user operation");
        g_ops[k](const_cast<void*>(in), io, &count, &t);
        return;
    }
    const int kind = CBMPI_KIND(t), sz = CBMPI_SIZE(t);
    const bool bits = op == MPI_BAND || op == MPI_BOR || op == MPI_BXOR;
#define CB_FOLD(T) do { if (bits) fold_bits<T>(op, (const T*)in, (T*)io, count); else fold_typed<T>(op, (const T*)in, (T*)io, count); return; } while (0)
    if (kind == CBMPI_SINT) { if (sz == 1) CB_FOLD(int8_t); if (sz == 2) CB_FOLD(int16_t); if (sz == 4) CB_FOLD(int32_t); if (sz == 8) CB_FOLD(int64_t); }
    if (kind == CBMPI_UINT) { if (sz == 1) CB_FOLD(uint8_t); if (sz == 2) CB_FOLD(uint16_t); if (sz == 4) CB_FOLD(uint32_t); if (sz == 8) CB_FOLD(uint64_t); }
#undef CB_FOLD
    if (kind == CBMPI_FLOAT && !bits) {
        if (sz == 4) { fold_typed<float>(op, (const float*)in, (float*)io, count); return; }
        if (sz == 8) { fold_typed<double>(op, (const double*)in, (double*)io, count); return; }
        if (sz == 16) { fold_typed<long double>(op, (const long double*)in, (long double*)io, count); return; }
    }
    die("built-in reduction on a derived datatype");
}
// out = v[0] op v[1] op ... op v[hi-1] in communicator-rank order, every v staged at arena offset 0
void reduce_ranks(const Ctx& x, int hi, size_t at, size_t bytes, int count, MPI_Datatype t, MPI_Op op, void* out) {
    std::vector<char> tmp(bytes ? bytes : 1);
    memcpy(out, staged(x, 0, at), bytes);
    for (int r = 1; r < hi; ++r) {
        memcpy(tmp.data(), staged(x, r, at), bytes);
        fold(op, t, out, tmp.data(), count);      // tmp = out op v[r]
        memcpy(out, tmp.data(), bytes);
    }
}

// ---- eager point-to-point ----
void post(const Ctx& x, const void* buf, long long bytes, int dest, int tag) {
    if (dest < 0 || dest >= x.n) die("send: destination out of range");
    const long long need = (bytes + 63) & ~63LL;
    unsigned spins = 0;
    for (;;) {
        if (__atomic_load_n(&S->heap_outstanding[g_rank], __ATOMIC_SEQ_CST) == 0) g_heap_top = 0;
        if (g_heap_top + need <= (long long)S->heap_bytes) break;
        if (need > (long long)S->heap_bytes) die("message exceeds CBMPI_HEAP_MB");
        relax(spins);
    }
    const long long off = g_heap_top;
    g_heap_top += need;
    if (bytes) memcpy(heap(g_rank) + off, buf, (size_t)bytes);
    __atomic_add_fetch(&S->heap_outstanding[g_rank], 1, __ATOMIC_SEQ_CST);
    Box& b = S->box[g_rank][x.c->members[dest]];
    Msg& m = b.q[b.head % QLEN];
    spins = 0;
    while (__atomic_load_n(&m.state, __ATOMIC_SEQ_CST) != 0) relax(spins);
    m.comm = x.handle == MPI_COMM_SELF ? -1 - g_rank : x.handle;
    m.tag = tag; m.bytes = bytes; m.off = off;
    __atomic_store_n(&m.state, 1, __ATOMIC_SEQ_CST);
    b.head = b.head + 1;
}
bool try_match(const Ctx& x, void* buf, long long cap, int src, int tag, MPI_Status* st) {
    const int want_comm = x.handle == MPI_COMM_SELF ? -1 - g_rank : x.handle;
    const int lo = src == MPI_ANY_SOURCE ? 0 : src, hi = src == MPI_ANY_SOURCE ? x.n : src + 1;
    if (lo < 0 || hi > x.n) die("receive: source out of range");
    for (int s = lo; s < hi; ++s) {
        const int ws = x.c->members[s];
        Box& b = S->box[ws][g_rank];
        for (unsigned p = b.tail; p != b.tail + QLEN; ++p) {
            Msg& m = b.q[p % QLEN];
            const int stt = __atomic_load_n(&m.state, __ATOMIC_SEQ_CST);
            if (stt == 0) break;
            if (stt == 2) continue;
            if (m.comm != want_comm || (tag != MPI_ANY_TAG && m.tag != tag)) continue;
            if (m.bytes > cap) die("receive buffer too small for the matched message");
            if (m.bytes) memcpy(buf, heap(ws) + m.off, (size_t)m.bytes);
            if (st) { st->MPI_SOURCE = s; st->MPI_TAG = m.tag; st->MPI_ERROR = 0; st->cb_bytes = m.bytes; }
            __atomic_sub_fetch(&S->heap_outstanding[ws], 1, __ATOMIC_SEQ_CST);
            __atomic_store_n(&m.state, 2, __ATOMIC_SEQ_CST);
            while (b.q[b.tail % QLEN].state == 2) {      // retire the consumed prefix
                __atomic_store_n(&b.q[b.tail % QLEN].state, 0, __ATOMIC_SEQ_CST);
                b.tail = b.tail + 1;
            }
            return true;
        }
    }
    return false;
}
void recv_blocking(const Ctx& x, void* buf, long long cap, int src, int tag, MPI_Status* st) {
    unsigned spins = 0;
    while (!try_match(x, buf, cap, src, tag, st)) relax(spins);
}

int split_impl(MPI_Comm comm, int color, int key, MPI_Comm* out) {
    Ctx x = ctx(comm);
    if (comm == MPI_COMM_SELF) { *out = color == MPI_UNDEFINED ? MPI_COMM_NULL : MPI_COMM_SELF; return 0; }
    x.c->pub[x.me].a = color;
    x.c->pub[x.me].b = key;
    barrier(x);
    std::vector<std::pair<long long, int>> grp;      // (key, index in parent) of my colour
    if (color != MPI_UNDEFINED)
        for (int i = 0; i < x.n; ++i)
            if (x.c->pub[i].a == color) grp.push_back(std::make_pair(x.c->pub[i].b, i));
    std::sort(grp.begin(), grp.end());
    if (!grp.empty() && grp[0].second == x.me) {
        const int slot = __atomic_fetch_add(&S->next_comm, 1, __ATOMIC_SEQ_CST);
        if (slot >= MAXC) die("out of communicator slots");
        Comm* nc = &S->comms[slot];
        for (size_t i = 0; i < grp.size(); ++i) nc->members[i] = x.c->members[grp[i].second];
        nc->bar_count = 0; nc->bar_gen = 0;
        __atomic_store_n(&nc->size, (int)grp.size(), __ATOMIC_SEQ_CST);
        x.c->pub[x.me].off = slot;
    }
    barrier(x);
    *out = grp.empty() ? MPI_COMM_NULL : (MPI_Comm)x.c->pub[grp[0].second].off;
    barrier(x);
    return 0;
}

}  // namespace

extern "C" {

int MPI_Init(int*, char***) { return 0; }
int MPI_Init_thread(int*, char***, int required, int* provided) { if (provided) *provided = required; return 0; }
int MPI_Is_thread_main(int* f) { *f = 1; return 0; }
int MPI_Query_thread(int* p) { *p = MPI_THREAD_FUNNELED; return 0; }
int MPI_Finalize(void) { barrier(ctx(MPI_COMM_WORLD)); g_finalized = 1; fflush(stdout); return 0; }
int MPI_Finalized(int* f) { *f = g_finalized; return 0; }
int MPI_Initialized(int* f) { *f = 1; return 0; }
int MPI_Abort(MPI_Comm, int code) {
    fflush(stdout);
    fprintf(stderr, "MPI_Abort(%d) on rank %d\n", code, g_rank);
    S->abort_code = (code & 0xff) ? (code & 0xff) : 1;
    __atomic_store_n(&S->abort_flag, 1, __ATOMIC_SEQ_CST);
    _exit(S->abort_code);
}
double MPI_Wtime(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
int MPI_Barrier(MPI_Comm c) { barrier(ctx(c)); return 0; }
int MPI_Pcontrol(int, ...) { return 0; }
int MPI_Error_string(int, char* s, int* l) { strcpy(s, "cbmpi"); *l = 5; return 0; }

int MPI_Comm_rank(MPI_Comm c, int* r) { *r = ctx(c).me; return 0; }
int MPI_Comm_size(MPI_Comm c, int* s) { *s = ctx(c).n; return 0; }
int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* n) { return split_impl(c, color, key, n); }
int MPI_Comm_dup(MPI_Comm c, MPI_Comm* n) { return split_impl(c, 0, ctx(c).me, n); }
int MPI_Comm_free(MPI_Comm* c) { *c = MPI_COMM_NULL; return 0; }                 // slots are not recycled
int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int* r) {
    if (a == b) { *r = MPI_IDENT; return 0; }
    Ctx x = ctx(a), y = ctx(b);
    if (x.n != y.n) { *r = MPI_UNEQUAL; return 0; }
    bool same_order = true;
    for (int i = 0; i < x.n; ++i) same_order = same_order && x.c->members[i] == y.c->members[i];
    if (same_order) { *r = MPI_CONGRUENT; return 0; }
    std::vector<int> p(x.c->members, x.c->members + x.n), q(y.c->members, y.c->members + y.n);
    std::sort(p.begin(), p.end()); std::sort(q.begin(), q.end());
    *r = p == q ? MPI_SIMILAR : MPI_UNEQUAL;
    return 0;
}
int MPI_Comm_group(MPI_Comm c, MPI_Group* g) {
    Ctx x = ctx(c);
    g_groups.push_back(std::vector<int>(x.c->members, x.c->members + x.n));
    *g = (MPI_Group)g_groups.size() - 1;
    return 0;
}
int MPI_Group_incl(MPI_Group g, int n, const int* ranks, MPI_Group* o) {
    std::vector<int> v;
    for (int i = 0; i < n; ++i) v.push_back(g_groups.at((size_t)g).at((size_t)ranks[i]));
    g_groups.push_back(v);
    *o = (MPI_Group)g_groups.size() - 1;
    return 0;
}
int MPI_Group_excl(MPI_Group g, int n, const int* ranks, MPI_Group* o) {
    std::vector<int> v;
    const std::vector<int>& src = g_groups.at((size_t)g);
    for (size_t i = 0; i < src.size(); ++i)
        if (std::find(ranks, ranks + n, (int)i) == ranks + n) v.push_back(src[i]);
    g_groups.push_back(v);
    *o = (MPI_Group)g_groups.size() - 1;
    return 0;
}
int MPI_Group_free(MPI_Group* g) { *g = MPI_GROUP_NULL; return 0; }
int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm* n) {
    const std::vector<int>& v = g_groups.at((size_t)g);
    const std::vector<int>::const_iterator it = std::find(v.begin(), v.end(), g_rank);
    return split_impl(c, it == v.end() ? MPI_UNDEFINED : 0, it == v.end() ? 0 : (int)(it - v.begin()), n);
}

int MPI_Type_contiguous(int n, MPI_Datatype t, MPI_Datatype* o) {
    const long long sz = (long long)n * CBMPI_SIZE(t);
    if (sz > 0xffffff) die("derived datatype larger than 16 MB");
    *o = CBMPI_TYPE(CBMPI_OPAQUE, (int)sz);
    return 0;
}
int MPI_Type_commit(MPI_Datatype*) { return 0; }
int MPI_Type_free(MPI_Datatype* t) { *t = MPI_DATATYPE_NULL; return 0; }
int MPI_Type_size(MPI_Datatype t, int* s) { *s = CBMPI_SIZE(t); return 0; }
int MPI_Type_create_struct(int n, const int* bl, const MPI_Aint* d, const MPI_Datatype* ts, MPI_Datatype* o) {
    long end = 0;
    for (int i = 0; i < n; ++i) end = std::max(end, d[i] + (long)bl[i] * CBMPI_SIZE(ts[i]));
    *o = CBMPI_TYPE(CBMPI_OPAQUE, (int)end);
    return 0;
}
int MPI_Op_create(MPI_User_function* f, int, MPI_Op* op) { g_ops.push_back(f); *op = CBMPI_FIRST_USER_OP + (int)g_ops.size() - 1; return 0; }
int MPI_Op_free(MPI_Op* op) { *op = MPI_OP_NULL; return 0; }

int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c) {
    Ctx x = ctx(c);
    const size_t bytes = (size_t)n * CBMPI_SIZE(t);
    if (x.n == 1) return 0;
    if (x.me == root) stage(b, bytes);
    barrier(x);
    if (x.me != root && bytes) memcpy(b, staged(x, root), bytes);
    barrier(x);
    return 0;
}
int MPI_Ibcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c, MPI_Request* r) { *r = MPI_REQUEST_NULL; return MPI_Bcast(b, n, t, root, c); }
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
    Ctx x = ctx(c);
    const size_t bytes = (size_t)n * CBMPI_SIZE(t);
    stage(s == MPI_IN_PLACE ? r : s, bytes);
    barrier(x);
    reduce_ranks(x, x.n, 0, bytes, n, t, op, r);
    barrier(x);
    return 0;
}
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) {
    Ctx x = ctx(c);
    const size_t bytes = (size_t)n * CBMPI_SIZE(t);
    stage(s == MPI_IN_PLACE ? r : s, bytes);
    barrier(x);
    if (x.me == root) reduce_ranks(x, x.n, 0, bytes, n, t, op, r);
    barrier(x);
    return 0;
}
int MPI_Reduce_scatter(const void* s, void* r, const int* counts, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
    Ctx x = ctx(c);
    long long total = 0, before = 0;
    for (int i = 0; i < x.n; ++i) { if (i < x.me) before += counts[i]; total += counts[i]; }
    if (s == MPI_IN_PLACE) die("MPI_Reduce_scatter in place is not provided");
    stage(s, (size_t)total * CBMPI_SIZE(t));
    barrier(x);
    reduce_ranks(x, x.n, (size_t)before * CBMPI_SIZE(t), (size_t)counts[x.me] * CBMPI_SIZE(t), counts[x.me], t, op, r);
    barrier(x);
    return 0;
}
int MPI_Scan(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
    Ctx x = ctx(c);
    const size_t bytes = (size_t)n * CBMPI_SIZE(t);
    stage(s == MPI_IN_PLACE ? r : s, bytes);
    barrier(x);
    reduce_ranks(x, x.me + 1, 0, bytes, n, t, op, r);
    barrier(x);
    return 0;
}
int MPI_Exscan(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
    Ctx x = ctx(c);
    const size_t bytes = (size_t)n * CBMPI_SIZE(t);
    stage(s == MPI_IN_PLACE ? r : s, bytes);
    barrier(x);
    if (x.me > 0) reduce_ranks(x, x.me, 0, bytes, n, t, op, r);      // rank 0's buffer is undefined by the standard
    barrier(x);
    return 0;
}
int MPI_Allgatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* rc, const int* dp, MPI_Datatype rt, MPI_Comm c) {
    Ctx x = ctx(c);
    const size_t rsz = (size_t)CBMPI_SIZE(rt);
    if (s == MPI_IN_PLACE) stage((char*)r + (size_t)dp[x.me] * rsz, (size_t)rc[x.me] * rsz);
    else stage(s, (size_t)sn * CBMPI_SIZE(st));
    barrier(x);
    for (int i = 0; i < x.n; ++i)
        if (!(s == MPI_IN_PLACE && i == x.me) && rc[i]) memcpy((char*)r + (size_t)dp[i] * rsz, staged(x, i), (size_t)rc[i] * rsz);
    barrier(x);
    return 0;
}
int MPI_Allgather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c) {
    Ctx x = ctx(c);
    std::vector<int> rc((size_t)x.n, rn), dp((size_t)x.n);
    for (int i = 0; i < x.n; ++i) dp[i] = i * rn;
    return MPI_Allgatherv(s, sn, st, r, rc.data(), dp.data(), rt, c);
}
int MPI_Gatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* rc, const int* dp, MPI_Datatype rt, int root, MPI_Comm c) {
    Ctx x = ctx(c);
    const size_t rsz = (size_t)CBMPI_SIZE(rt);
    if (s != MPI_IN_PLACE) stage(s, (size_t)sn * CBMPI_SIZE(st));
    barrier(x);
    if (x.me == root)
        for (int i = 0; i < x.n; ++i)
            if (!(s == MPI_IN_PLACE && i == x.me) && rc[i]) memcpy((char*)r + (size_t)dp[i] * rsz, staged(x, i), (size_t)rc[i] * rsz);
    barrier(x);
    return 0;
}
int MPI_Gather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
    Ctx x = ctx(c);
    std::vector<int> rc((size_t)x.n, rn), dp((size_t)x.n);
    for (int i = 0; i < x.n; ++i) dp[i] = i * rn;
    return MPI_Gatherv(s, sn, st, r, rc.data(), dp.data(), rt, root, c);
}
int MPI_Scatterv(const void* s, const int* sc, const int* dp, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
    Ctx x = ctx(c);
    const size_t ssz = (size_t)CBMPI_SIZE(st);
    if (x.me == root) {
        long long end = 0;
        for (int i = 0; i < x.n; ++i) end = std::max(end, (long long)dp[i] + sc[i]);
        stage(s, (size_t)end * ssz);
        for (int i = 0; i < x.n; ++i) x.c->pub[i].off = (long long)dp[i] * (long long)ssz;
    }
    barrier(x);
    if (r != MPI_IN_PLACE && rn) memcpy(r, staged(x, root, (size_t)x.c->pub[x.me].off), (size_t)rn * CBMPI_SIZE(rt));
    barrier(x);
    return 0;
}
int MPI_Scatter(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
    Ctx x = ctx(c);
    std::vector<int> sc((size_t)x.n, sn), dp((size_t)x.n);
    for (int i = 0; i < x.n; ++i) dp[i] = i * sn;
    return MPI_Scatterv(s, sc.data(), dp.data(), st, r, rn, rt, root, c);
}
int MPI_Alltoallv(const void* s, const int* sc, const int* sd, MPI_Datatype st, void* r, const int* rc, const int* rd, MPI_Datatype rt, MPI_Comm c) {
    Ctx x = ctx(c);
    if (s == MPI_IN_PLACE) die("MPI_Alltoallv in place is not provided");
    const size_t ssz = (size_t)CBMPI_SIZE(st), rsz = (size_t)CBMPI_SIZE(rt);
    // staged layout: [n x (count, displacement) as long long][64-byte aligned payload]
    std::vector<long long> hdr((size_t)2 * x.n);
    long long end = 0;
    for (int i = 0; i < x.n; ++i) { hdr[2 * i] = sc[i]; hdr[2 * i + 1] = sd[i]; end = std::max(end, (long long)sd[i] + sc[i]); }
    const size_t pay = ((size_t)2 * x.n * sizeof(long long) + 63) & ~(size_t)63;
    stage(hdr.data(), hdr.size() * sizeof(long long));
    stage(s, (size_t)end * ssz, pay);
    barrier(x);
    for (int i = 0; i < x.n; ++i) {
        const long long* h = (const long long*)staged(x, i);
        const long long cnt = h[2 * x.me], dsp = h[2 * x.me + 1];
        if ((size_t)cnt * ssz > (size_t)rc[i] * rsz) die("MPI_Alltoallv: receive count smaller than what was sent");
        if (cnt) memcpy((char*)r + (size_t)rd[i] * rsz, staged(x, i, pay + (size_t)dsp * ssz), (size_t)cnt * ssz);
    }
    barrier(x);
    return 0;
}
int MPI_Alltoall(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c) {
    Ctx x = ctx(c);
    std::vector<int> sc((size_t)x.n, sn), sd((size_t)x.n), rc((size_t)x.n, rn), rd((size_t)x.n);
    for (int i = 0; i < x.n; ++i) { sd[i] = i * sn; rd[i] = i * rn; }
    return MPI_Alltoallv(s, sc.data(), sd.data(), st, r, rc.data(), rd.data(), rt, c);
}

int MPI_Send(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) { post(ctx(c), b, (long long)n * CBMPI_SIZE(t), d, tag); return 0; }
int MPI_Isend(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request* r) { *r = MPI_REQUEST_NULL; return MPI_Send(b, n, t, d, tag, c); }
int MPI_Issend(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request* r) { *r = MPI_REQUEST_NULL; return MPI_Send(b, n, t, d, tag, c); }
int MPI_Recv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status* st) { recv_blocking(ctx(c), b, (long long)n * CBMPI_SIZE(t), s, tag, st); return 0; }
int MPI_Irecv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Request* r) {
    g_reqs.push_back(Req{1, b, (long long)n * CBMPI_SIZE(t), s, tag, c});
    *r = (MPI_Request)g_reqs.size() - 1;
    return 0;
}
int MPI_Sendrecv(const void* s, int sn, MPI_Datatype st, int dest, int stag, void* r, int rn, MPI_Datatype rt, int src, int rtag, MPI_Comm c, MPI_Status* status) {
    Ctx x = ctx(c);
    post(x, s, (long long)sn * CBMPI_SIZE(st), dest, stag);
    recv_blocking(x, r, (long long)rn * CBMPI_SIZE(rt), src, rtag, status);
    return 0;
}
int MPI_Wait(MPI_Request* r, MPI_Status* st) {
    if (*r != MPI_REQUEST_NULL) {
        Req& q = g_reqs.at((size_t)*r);
        if (q.kind == 1) { recv_blocking(ctx(q.comm), q.buf, q.bytes, q.src, q.tag, st); q.kind = 0; }
        *r = MPI_REQUEST_NULL;
    }
    return 0;
}
int MPI_Waitall(int n, MPI_Request* r, MPI_Status* st) {
    for (int i = 0; i < n; ++i) MPI_Wait(&r[i], st ? &st[i] : MPI_STATUS_IGNORE);
    return 0;
}
int MPI_Test(MPI_Request* r, int* flag, MPI_Status* st) {
    *flag = 1;
    if (*r != MPI_REQUEST_NULL) {
        Req& q = g_reqs.at((size_t)*r);
        if (q.kind == 1) {
            if (try_match(ctx(q.comm), q.buf, q.bytes, q.src, q.tag, st)) q.kind = 0; else *flag = 0;
        }
        if (*flag) *r = MPI_REQUEST_NULL;
    }
    return 0;
}
int MPI_Get_count(const MPI_Status* s, MPI_Datatype t, int* n) { *n = (int)(s->cb_bytes / (CBMPI_SIZE(t) ? CBMPI_SIZE(t) : 1)); return 0; }

struct cbmpi_file { int fd; long long view, pos; };
int MPI_File_open(MPI_Comm, const char* name, int mode, MPI_Info, MPI_File* fh) {
    int flags = (mode & MPI_MODE_WRONLY) ? O_WRONLY : O_RDONLY;
    if (mode & MPI_MODE_CREATE) flags |= O_CREAT;
    const int fd = open(name, flags, 0644);
    if (fd < 0) { *fh = NULL; return 1; }
    *fh = (MPI_File)malloc(sizeof(cbmpi_file));
    (*fh)->fd = fd; (*fh)->view = 0; (*fh)->pos = 0;
    return 0;
}
int MPI_File_close(MPI_File* fh) { if (*fh) { close((*fh)->fd); free(*fh); *fh = NULL; } return 0; }
int MPI_File_read_at(MPI_File fh, MPI_Offset off, void* buf, int n, MPI_Datatype t, MPI_Status* st) {
    const size_t want = (size_t)n * CBMPI_SIZE(t);
    size_t got = 0;
    while (got < want) {
        const ssize_t k = pread(fh->fd, (char*)buf + got, want - got, (off_t)(fh->view + off) + (off_t)got);
        if (k <= 0) break;
        got += (size_t)k;
    }
    if (st) { st->MPI_SOURCE = 0; st->MPI_TAG = 0; st->MPI_ERROR = 0; st->cb_bytes = (long long)got; }
    return 0;
}
int MPI_File_set_view(MPI_File fh, MPI_Offset disp, MPI_Datatype, MPI_Datatype, const char*, MPI_Info) { fh->view = disp; fh->pos = 0; return 0; }
int MPI_File_write(MPI_File fh, const void* buf, int n, MPI_Datatype t, MPI_Status*) {
    const size_t want = (size_t)n * CBMPI_SIZE(t);
    size_t put = 0;
    while (put < want) {
        const ssize_t k = pwrite(fh->fd, (const char*)buf + put, want - put, (off_t)(fh->view + fh->pos) + (off_t)put);
        if (k <= 0) die("MPI_File_write failed");
        put += (size_t)k;
    }
    fh->pos += (long long)want;
    return 0;
}
int MPI_File_write_all(MPI_File fh, const void* buf, int n, MPI_Datatype t, MPI_Status* st) { return MPI_File_write(fh, buf, n, t, st); }
int MPI_Info_create(MPI_Info* i) { *i = 0; return 0; }
int MPI_Info_set(MPI_Info, const char*, const char*) { return 0; }
int MPI_Info_free(MPI_Info* i) { *i = 0; return 0; }

int MPI_Win_create(void*, MPI_Aint, int, MPI_Info, MPI_Comm, MPI_Win*) { die("one-sided communication is not provided"); }
int MPI_Win_free(MPI_Win*) { return 0; }
int MPI_Win_fence(int, MPI_Win) { die("one-sided communication is not provided"); }
int MPI_Win_lock(int, int, int, MPI_Win) { die("one-sided communication is not provided"); }
int MPI_Win_unlock(int, MPI_Win) { die("one-sided communication is not provided"); }
int MPI_Win_start(MPI_Group, int, MPI_Win) { die("one-sided communication is not provided"); }
int MPI_Win_post(MPI_Group, int, MPI_Win) { die("one-sided communication is not provided"); }
int MPI_Win_wait(MPI_Win) { die("one-sided communication is not provided"); }
int MPI_Win_complete(MPI_Win) { die("one-sided communication is not provided"); }
int MPI_Put(const void*, int, MPI_Datatype, int, MPI_Aint, int, MPI_Datatype, MPI_Win) { die("one-sided communication is not provided"); }
int MPI_Get(void*, int, MPI_Datatype, int, MPI_Aint, int, MPI_Datatype, MPI_Win) { die("one-sided communication is not provided"); }

}  // extern "C"

static long env_long(const char* name, long dflt) { const char* v = getenv(name); return v && *v ? atol(v) : dflt; }

int main(int argc, char** argv) {
    const int np = (int)env_long("CBMPI_NP", 1);
    if (np < 1 || np > MAXP) { fprintf(stderr, "cbmpi: CBMPI_NP must be 1..%d\n", MAXP); return 2; }
    const size_t arena_bytes = (size_t)env_long("CBMPI_ARENA_MB", 256) << 20, heap_bytes = (size_t)env_long("CBMPI_HEAP_MB", 64) << 20;
    const size_t head = (sizeof(Shared) + 4095) & ~(size_t)4095;
    const size_t total = head + (size_t)np * (arena_bytes + heap_bytes);
    void* mem = mmap(NULL, total, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (mem == MAP_FAILED) { perror("cbmpi: mmap"); return 2; }
    S = (Shared*)mem;                              // fresh anonymous pages are zero
    g_base = (char*)mem + head;
    S->np = np; S->arena_bytes = arena_bytes; S->heap_bytes = heap_bytes;
    S->next_comm = 3;                              // 0 null, 1 world, 2 self
    for (int i = 0; i < np; ++i) S->comms[MPI_COMM_WORLD].members[i] = i;
    S->comms[MPI_COMM_WORLD].size = np;
    fflush(stdout); fflush(stderr);
    std::vector<pid_t> kids;
    for (int r = 0; r < np; ++r) {
        const pid_t pid = fork();
        if (pid < 0) { perror("cbmpi: fork"); for (pid_t k : kids) kill(k, SIGKILL); return 2; }
        if (pid == 0) {
            g_rank = r;
            g_self.size = 1; g_self.members[0] = r;
            const int rc = cb_rank_main(argc, argv);
            fflush(stdout); fflush(stderr);
            _exit(rc);
        }
        kids.push_back(pid);
    }
    const double deadline = MPI_Wtime() + (double)env_long("CBMPI_TIMEOUT", 900);
    int alive = np, code = 0;
    while (alive > 0) {
        int st = 0;
        const pid_t pid = waitpid(-1, &st, WNOHANG);
        if (pid == 0) {
            if (MPI_Wtime() > deadline) { fprintf(stderr, "cbmpi: timeout\n"); code = 124; break; }
            usleep(2000);
            continue;
        }
        if (pid < 0) break;
        --alive;
        const int rc = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + WTERMSIG(st);
        if (rc != 0 && code == 0) { code = S->abort_flag && S->abort_code ? S->abort_code : rc; break; }
    }
    if (code != 0) {
        __atomic_store_n(&S->abort_flag, 1, __ATOMIC_SEQ_CST);
        usleep(20000);
        for (pid_t k : kids) kill(k, SIGKILL);
        while (waitpid(-1, NULL, 0) > 0) {}
    }
    return code;
}
