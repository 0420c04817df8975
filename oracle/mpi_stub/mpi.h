/* Single-rank MPI stand-in used ONLY to compile the unmodified reference
 * (/root/reference) into oracle/_ref/ as a parity checker.  TEST INFRASTRUCTURE:
 * never included by the product (combblas-spmm-test_b200/).
 *
 * One rank, so every collective degenerates to "copy send buffer to receive
 * buffer unless MPI_IN_PLACE".  Datatype handles are plain ints holding the
 * element size in bytes, so MPI_Type_contiguous(n, T) is n*size(T) — this is
 * exactly how the reference builds its derived types
 * (include/CombBLAS/MPIType.h:95-110).  Names inside templates of the reference
 * that are never instantiated still have to be declared; those are the
 * abort-on-call entries at the bottom.
 */
#ifndef CB_ORACLE_MPI_STUB_H
#define CB_ORACLE_MPI_STUB_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;   /* value = size in bytes; 0 = MPI_DATATYPE_NULL */
typedef int MPI_Op;
typedef int MPI_Win;
typedef int MPI_Request;
typedef int MPI_Group;
typedef int MPI_Info;
typedef long MPI_Aint;
typedef long long MPI_Offset;
typedef struct { FILE* fp; } *MPI_File;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; long long cb_bytes; } MPI_Status;
typedef void(MPI_User_function)(void*, void*, int*, MPI_Datatype*);

#define MPI_COMM_WORLD 1
#define MPI_COMM_NULL 0
#define MPI_COMM_SELF 2
#define MPI_IN_PLACE ((void*)1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_REQUEST_NULL 0
#define MPI_INFO_NULL 0
#define MPI_DATATYPE_NULL 0
#define MPI_OP_NULL 0
#define MPI_SUCCESS 0
#define MPI_IDENT 0
#define MPI_CONGRUENT 1
#define MPI_SIMILAR 2
#define MPI_UNEQUAL 3
#define MPI_MAX_ERROR_STRING 256
#define MPI_LOCK_SHARED 1
#define MPI_LOCK_EXCLUSIVE 2
#define MPI_MODE_NOPUT 1
#define MPI_MODE_NOSUCCEED 2
#define MPI_MODE_NOSTORE 4
#define MPI_MODE_NOPRECEDE 8
#define MPI_MODE_NOCHECK 16
#define MPI_MODE_RDONLY 32
#define MPI_MODE_WRONLY 64
#define MPI_MODE_CREATE 128
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_UNDEFINED (-32766)

#define MPI_CHAR 1
#define MPI_BYTE 1
#define MPI_SIGNED_CHAR 1
#define MPI_UNSIGNED_CHAR 1
#define MPI_SHORT 2
#define MPI_UNSIGNED_SHORT 2
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_FLOAT 4
#define MPI_LONG 8
#define MPI_UNSIGNED_LONG 8
#define MPI_LONG_LONG 8
#define MPI_LONG_LONG_INT 8
#define MPI_UNSIGNED_LONG_LONG 8
#define MPI_DOUBLE 8
#define MPI_LONG_DOUBLE 16
#define MPI_2INT 8
#define MPI_SHORT_INT 8
#define MPI_LONG_INT 16
#define MPI_FLOAT_INT 8
#define MPI_DOUBLE_INT 16
#define MPI_LONG_DOUBLE_INT 32

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_PROD 4
#define MPI_LAND 5
#define MPI_LOR 6
#define MPI_LXOR 7
#define MPI_BAND 8
#define MPI_BOR 9
#define MPI_BXOR 10

static inline void cb_stub_unsupported(const char* what) {
    fprintf(stderr, "mpi_stub: %s is not available in the single-rank oracle\n", what);
    abort();
}
static inline void cb_stub_copy(const void* s, void* r, long long count, MPI_Datatype t) {
    if (s != MPI_IN_PLACE && s != r && count > 0) memcpy(r, s, (size_t)count * (size_t)t);
}

/* ---- environment ---- */
static inline int MPI_Init(int* a, char*** b) { (void)a; (void)b; return 0; }
static inline int MPI_Init_thread(int* a, char*** b, int req, int* prov) { (void)a; (void)b; if (prov) *prov = req; return 0; }
static inline int MPI_Is_thread_main(int* f) { *f = 1; return 0; }
static inline int MPI_Query_thread(int* p) { *p = MPI_THREAD_FUNNELED; return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Finalized(int* f) { *f = 0; return 0; }
static inline int MPI_Initialized(int* f) { *f = 1; return 0; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; fprintf(stderr, "MPI_Abort(%d)\n", code); exit(code & 0xff ? code & 0xff : 1); return 0; }
static inline double MPI_Wtime(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
static inline int MPI_Pcontrol(int l, ...) { (void)l; return 0; }
static inline int MPI_Error_string(int e, char* s, int* l) { (void)e; strcpy(s, "mpi_stub"); *l = 8; return 0; }

/* ---- communicators / groups: identity ---- */
static inline int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int* s) { (void)c; *s = 1; return 0; }
static inline int MPI_Comm_dup(MPI_Comm c, MPI_Comm* n) { *n = c; return 0; }
static inline int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* n) { (void)color; (void)key; *n = c; return 0; }
static inline int MPI_Comm_free(MPI_Comm* c) { *c = MPI_COMM_NULL; return 0; }
static inline int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int* r) { (void)a; (void)b; *r = MPI_CONGRUENT; return 0; }
static inline int MPI_Comm_group(MPI_Comm c, MPI_Group* g) { *g = c; return 0; }
static inline int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm* n) { (void)g; *n = c; return 0; }
static inline int MPI_Group_incl(MPI_Group g, int n, const int* r, MPI_Group* o) { (void)n; (void)r; *o = g; return 0; }
static inline int MPI_Group_excl(MPI_Group g, int n, const int* r, MPI_Group* o) { (void)n; (void)r; *o = g; return 0; }
static inline int MPI_Group_free(MPI_Group* g) { *g = 0; return 0; }

/* ---- datatypes / ops ---- */
static inline int MPI_Type_contiguous(int n, MPI_Datatype t, MPI_Datatype* o) { *o = n * t; return 0; }
static inline int MPI_Type_commit(MPI_Datatype* t) { (void)t; return 0; }
static inline int MPI_Type_free(MPI_Datatype* t) { *t = MPI_DATATYPE_NULL; return 0; }
static inline int MPI_Type_size(MPI_Datatype t, int* s) { *s = t; return 0; }
static inline int MPI_Type_create_struct(int n, const int* bl, const MPI_Aint* d, const MPI_Datatype* ts, MPI_Datatype* o) {
    long end = 0; for (int i = 0; i < n; ++i) { long e = d[i] + (long)bl[i] * ts[i]; if (e > end) end = e; } *o = (int)end; return 0; }
static inline int MPI_Op_create(MPI_User_function* f, int commute, MPI_Op* op) { (void)f; (void)commute; *op = 100; return 0; }
static inline int MPI_Op_free(MPI_Op* op) { *op = MPI_OP_NULL; return 0; }

/* ---- collectives with real single-rank semantics ---- */
static inline int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b; (void)n; (void)t; (void)root; (void)c; return 0; }
static inline int MPI_Ibcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c, MPI_Request* r) { (void)b; (void)n; (void)t; (void)root; (void)c; *r = 0; return 0; }
static inline int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) { (void)op; (void)c; cb_stub_copy(s, r, n, t); return 0; }
static inline int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) { (void)op; (void)root; (void)c; cb_stub_copy(s, r, n, t); return 0; }
static inline int MPI_Reduce_scatter(const void* s, void* r, const int* cnt, MPI_Datatype t, MPI_Op op, MPI_Comm c) { (void)op; (void)c; cb_stub_copy(s, r, cnt[0], t); return 0; }
static inline int MPI_Scan(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) { (void)op; (void)c; cb_stub_copy(s, r, n, t); return 0; }
/* exclusive scan: rank 0's receive buffer is undefined by the standard; leave it untouched */
static inline int MPI_Exscan(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) { (void)s; (void)r; (void)n; (void)t; (void)op; (void)c; return 0; }
static inline int MPI_Allgather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c) { (void)rn; (void)rt; (void)c; cb_stub_copy(s, r, sn, st); return 0; }
static inline int MPI_Allgatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* rc, const int* dp, MPI_Datatype rt, MPI_Comm c) {
    (void)rc; (void)c; if (s != MPI_IN_PLACE) cb_stub_copy(s, (char*)r + (size_t)dp[0] * rt, sn, st); return 0; }
static inline int MPI_Gather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) { (void)rn; (void)rt; (void)root; (void)c; cb_stub_copy(s, r, sn, st); return 0; }
static inline int MPI_Gatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* rc, const int* dp, MPI_Datatype rt, int root, MPI_Comm c) {
    (void)rc; (void)root; (void)c; if (s != MPI_IN_PLACE) cb_stub_copy(s, (char*)r + (size_t)dp[0] * rt, sn, st); return 0; }
static inline int MPI_Scatter(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) { (void)sn; (void)st; (void)root; (void)c; if (r != MPI_IN_PLACE) cb_stub_copy(s, r, rn, rt); return 0; }
static inline int MPI_Scatterv(const void* s, const int* sc, const int* dp, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
    (void)sc; (void)root; (void)c; if (r != MPI_IN_PLACE) cb_stub_copy((const char*)s + (size_t)dp[0] * st, r, rn, rt); return 0; }
static inline int MPI_Alltoall(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c) { (void)rn; (void)rt; (void)c; cb_stub_copy(s, r, sn, st); return 0; }
static inline int MPI_Alltoallv(const void* s, const int* sc, const int* sd, MPI_Datatype st, void* r, const int* rc, const int* rd, MPI_Datatype rt, MPI_Comm c) {
    (void)rc; (void)c; if (s != MPI_IN_PLACE) cb_stub_copy((const char*)s + (size_t)sd[0] * st, (char*)r + (size_t)rd[0] * rt, sc[0], st); return 0; }
static inline int MPI_Sendrecv(const void* s, int sn, MPI_Datatype st, int dest, int stag, void* r, int rn, MPI_Datatype rt, int src, int rtag, MPI_Comm c, MPI_Status* status) {
    (void)dest; (void)stag; (void)rn; (void)src; (void)rtag; (void)c;
    cb_stub_copy(s, r, sn, st);
    if (status) { status->MPI_SOURCE = 0; status->MPI_TAG = 0; status->MPI_ERROR = 0; status->cb_bytes = (long long)sn * st; }
    (void)rt; return 0; }
static inline int MPI_Get_count(const MPI_Status* s, MPI_Datatype t, int* n) { *n = (int)(s->cb_bytes / (t ? t : 1)); return 0; }

/* ---- MPI-IO over stdio (ParallelReadMM, SpParMat.cpp:3978-4115) ---- */
static inline int MPI_File_open(MPI_Comm c, const char* name, int mode, MPI_Info info, MPI_File* fh) {
    (void)c; (void)info;
    FILE* fp = fopen(name, (mode & MPI_MODE_WRONLY) ? "wb" : "rb");
    if (!fp) { *fh = NULL; return 1; }
    *fh = (MPI_File)malloc(sizeof(**fh)); (*fh)->fp = fp; return 0; }
static inline int MPI_File_close(MPI_File* fh) { if (*fh) { fclose((*fh)->fp); free(*fh); *fh = NULL; } return 0; }
static inline int MPI_File_read_at(MPI_File fh, MPI_Offset off, void* buf, int n, MPI_Datatype t, MPI_Status* st) {
    fseeko(fh->fp, (off_t)off, SEEK_SET);
    size_t got = fread(buf, 1, (size_t)n * t, fh->fp);
    if (st) { st->MPI_SOURCE = 0; st->MPI_TAG = 0; st->MPI_ERROR = 0; st->cb_bytes = (long long)got; }
    return 0; }
static inline int MPI_File_set_view(MPI_File fh, MPI_Offset disp, MPI_Datatype e, MPI_Datatype f, const char* rep, MPI_Info info) {
    (void)e; (void)f; (void)rep; (void)info; fseeko(fh->fp, (off_t)disp, SEEK_SET); return 0; }
static inline int MPI_File_write(MPI_File fh, const void* buf, int n, MPI_Datatype t, MPI_Status* st) { (void)st; fwrite(buf, 1, (size_t)n * t, fh->fp); return 0; }
static inline int MPI_File_write_all(MPI_File fh, const void* buf, int n, MPI_Datatype t, MPI_Status* st) { return MPI_File_write(fh, buf, n, t, st); }
static inline int MPI_Info_create(MPI_Info* i) { *i = 0; return 0; }
static inline int MPI_Info_set(MPI_Info i, const char* k, const char* v) { (void)i; (void)k; (void)v; return 0; }
static inline int MPI_Info_free(MPI_Info* i) { *i = 0; return 0; }

/* ---- declared for name lookup only; a single rank never reaches them ---- */
static inline int MPI_Send(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) { (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; cb_stub_unsupported("MPI_Send"); return 1; }
static inline int MPI_Recv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status* st) { (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)st; cb_stub_unsupported("MPI_Recv"); return 1; }
static inline int MPI_Isend(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request* r) { (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; (void)r; cb_stub_unsupported("MPI_Isend"); return 1; }
static inline int MPI_Issend(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request* r) { (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; (void)r; cb_stub_unsupported("MPI_Issend"); return 1; }
static inline int MPI_Irecv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Request* r) { (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)r; cb_stub_unsupported("MPI_Irecv"); return 1; }
static inline int MPI_Wait(MPI_Request* r, MPI_Status* s) { (void)r; (void)s; return 0; }
static inline int MPI_Waitall(int n, MPI_Request* r, MPI_Status* s) { (void)n; (void)r; (void)s; return 0; }
static inline int MPI_Test(MPI_Request* r, int* flag, MPI_Status* s) { (void)r; (void)s; *flag = 1; return 0; }
static inline int MPI_Win_create(void* b, MPI_Aint sz, int du, MPI_Info i, MPI_Comm c, MPI_Win* w) { (void)b; (void)sz; (void)du; (void)i; (void)c; (void)w; cb_stub_unsupported("MPI_Win_create"); return 1; }
static inline int MPI_Win_free(MPI_Win* w) { (void)w; return 0; }
static inline int MPI_Win_fence(int a, MPI_Win w) { (void)a; (void)w; return 0; }
static inline int MPI_Win_lock(int t, int r, int a, MPI_Win w) { (void)t; (void)r; (void)a; (void)w; return 0; }
static inline int MPI_Win_unlock(int r, MPI_Win w) { (void)r; (void)w; return 0; }
static inline int MPI_Win_start(MPI_Group g, int a, MPI_Win w) { (void)g; (void)a; (void)w; return 0; }
static inline int MPI_Win_post(MPI_Group g, int a, MPI_Win w) { (void)g; (void)a; (void)w; return 0; }
static inline int MPI_Win_wait(MPI_Win w) { (void)w; return 0; }
static inline int MPI_Win_complete(MPI_Win w) { (void)w; return 0; }
static inline int MPI_Put(const void* o, int on, MPI_Datatype ot, int tr, MPI_Aint td, int tn, MPI_Datatype tt, MPI_Win w) { (void)o; (void)on; (void)ot; (void)tr; (void)td; (void)tn; (void)tt; (void)w; cb_stub_unsupported("MPI_Put"); return 1; }
static inline int MPI_Get(void* o, int on, MPI_Datatype ot, int tr, MPI_Aint td, int tn, MPI_Datatype tt, MPI_Win w) { (void)o; (void)on; (void)ot; (void)tr; (void)td; (void)tn; (void)tt; (void)w; cb_stub_unsupported("MPI_Get"); return 1; }

#ifdef __cplusplus
}
#endif
#endif
