/* Process-per-rank MPI stand-in used ONLY to run the unmodified reference (/root/reference) on a
 * pr x pc process grid inside one container that has no MPI.
 * TEST INFRASTRUCTURE: never included by the product (combblas-spmm-test_b200/).
 *
 * cbmpi.cpp holds the implementation and the launcher: `main` forks CBMPI_NP ranks that share one
 * anonymous MAP_SHARED region (communicator table, per-rank staging arenas, eager point-to-point
 * mailboxes) and then runs the program's own main (compiled with -Dmain=cb_rank_main) in each.
 * Every collective is "stage my contribution in my arena, barrier, read what I need, barrier";
 * reductions fold in communicator-rank order, so results are deterministic.
 *
 * Datatype handles are ints: low 24 bits = element size in bytes, bits 24-27 = kind (opaque, signed,
 * unsigned, floating) so built-in reductions know their arithmetic; MPI_Type_contiguous(n, T) is an
 * opaque type of n*size(T) bytes, which is how the reference builds its derived types
 * (include/CombBLAS/MPIType.h:95-110).  One-sided communication is declared but aborts when called
 * (only the reference's legacy RMA multiplies use it; they are off the path).
 */
#ifndef CB_ORACLE_MPI_MULTI_H
#define CB_ORACLE_MPI_MULTI_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Win;
typedef int MPI_Request;
typedef int MPI_Group;
typedef int MPI_Info;
typedef long MPI_Aint;
typedef long long MPI_Offset;
typedef struct cbmpi_file* MPI_File;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; long long cb_bytes; } MPI_Status;
typedef void(MPI_User_function)(void*, void*, int*, MPI_Datatype*);

#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_IN_PLACE ((void*)1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_REQUEST_NULL 0
#define MPI_INFO_NULL 0
#define MPI_DATATYPE_NULL 0
#define MPI_OP_NULL 0
#define MPI_GROUP_NULL 0
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

#define CBMPI_OPAQUE 0
#define CBMPI_SINT 1
#define CBMPI_UINT 2
#define CBMPI_FLOAT 3
#define CBMPI_TYPE(kind, size) (((kind) << 24) | (size))
#define CBMPI_SIZE(t) ((t) & 0xffffff)
#define CBMPI_KIND(t) (((t) >> 24) & 0xf)

#define MPI_CHAR CBMPI_TYPE(CBMPI_SINT, 1)
#define MPI_SIGNED_CHAR CBMPI_TYPE(CBMPI_SINT, 1)
#define MPI_BYTE CBMPI_TYPE(CBMPI_UINT, 1)
#define MPI_UNSIGNED_CHAR CBMPI_TYPE(CBMPI_UINT, 1)
#define MPI_SHORT CBMPI_TYPE(CBMPI_SINT, 2)
#define MPI_UNSIGNED_SHORT CBMPI_TYPE(CBMPI_UINT, 2)
#define MPI_INT CBMPI_TYPE(CBMPI_SINT, 4)
#define MPI_UNSIGNED CBMPI_TYPE(CBMPI_UINT, 4)
#define MPI_LONG CBMPI_TYPE(CBMPI_SINT, 8)
#define MPI_UNSIGNED_LONG CBMPI_TYPE(CBMPI_UINT, 8)
#define MPI_LONG_LONG CBMPI_TYPE(CBMPI_SINT, 8)
#define MPI_LONG_LONG_INT CBMPI_TYPE(CBMPI_SINT, 8)
#define MPI_UNSIGNED_LONG_LONG CBMPI_TYPE(CBMPI_UINT, 8)
#define MPI_FLOAT CBMPI_TYPE(CBMPI_FLOAT, 4)
#define MPI_DOUBLE CBMPI_TYPE(CBMPI_FLOAT, 8)
#define MPI_LONG_DOUBLE CBMPI_TYPE(CBMPI_FLOAT, 16)
#define MPI_2INT CBMPI_TYPE(CBMPI_OPAQUE, 8)
#define MPI_SHORT_INT CBMPI_TYPE(CBMPI_OPAQUE, 8)
#define MPI_LONG_INT CBMPI_TYPE(CBMPI_OPAQUE, 16)
#define MPI_FLOAT_INT CBMPI_TYPE(CBMPI_OPAQUE, 8)
#define MPI_DOUBLE_INT CBMPI_TYPE(CBMPI_OPAQUE, 16)
#define MPI_LONG_DOUBLE_INT CBMPI_TYPE(CBMPI_OPAQUE, 32)

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
#define CBMPI_FIRST_USER_OP 100

/* environment */
int MPI_Init(int*, char***);
int MPI_Init_thread(int*, char***, int required, int* provided);
int MPI_Is_thread_main(int*);
int MPI_Query_thread(int*);
int MPI_Finalize(void);
int MPI_Finalized(int*);
int MPI_Initialized(int*);
int MPI_Abort(MPI_Comm, int);
double MPI_Wtime(void);
int MPI_Barrier(MPI_Comm);
int MPI_Pcontrol(int, ...);
int MPI_Error_string(int, char*, int*);

/* communicators and groups */
int MPI_Comm_rank(MPI_Comm, int*);
int MPI_Comm_size(MPI_Comm, int*);
int MPI_Comm_dup(MPI_Comm, MPI_Comm*);
int MPI_Comm_split(MPI_Comm, int color, int key, MPI_Comm*);
int MPI_Comm_free(MPI_Comm*);
int MPI_Comm_compare(MPI_Comm, MPI_Comm, int*);
int MPI_Comm_group(MPI_Comm, MPI_Group*);
int MPI_Comm_create(MPI_Comm, MPI_Group, MPI_Comm*);
int MPI_Group_incl(MPI_Group, int, const int*, MPI_Group*);
int MPI_Group_excl(MPI_Group, int, const int*, MPI_Group*);
int MPI_Group_free(MPI_Group*);

/* datatypes and operations */
int MPI_Type_contiguous(int, MPI_Datatype, MPI_Datatype*);
int MPI_Type_commit(MPI_Datatype*);
int MPI_Type_free(MPI_Datatype*);
int MPI_Type_size(MPI_Datatype, int*);
int MPI_Type_create_struct(int, const int*, const MPI_Aint*, const MPI_Datatype*, MPI_Datatype*);
int MPI_Op_create(MPI_User_function*, int commute, MPI_Op*);
int MPI_Op_free(MPI_Op*);

/* collectives */
int MPI_Bcast(void*, int, MPI_Datatype, int root, MPI_Comm);
int MPI_Ibcast(void*, int, MPI_Datatype, int root, MPI_Comm, MPI_Request*);
int MPI_Allreduce(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Reduce(const void*, void*, int, MPI_Datatype, MPI_Op, int root, MPI_Comm);
int MPI_Reduce_scatter(const void*, void*, const int* counts, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Scan(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Exscan(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Allgather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm);
int MPI_Allgatherv(const void*, int, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, MPI_Comm);
int MPI_Gather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, int root, MPI_Comm);
int MPI_Gatherv(const void*, int, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, int root, MPI_Comm);
int MPI_Scatter(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, int root, MPI_Comm);
int MPI_Scatterv(const void*, const int*, const int*, MPI_Datatype, void*, int, MPI_Datatype, int root, MPI_Comm);
int MPI_Alltoall(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm);
int MPI_Alltoallv(const void*, const int*, const int*, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, MPI_Comm);

/* point to point (eager, buffered in the shared region) */
int MPI_Send(const void*, int, MPI_Datatype, int dest, int tag, MPI_Comm);
int MPI_Recv(void*, int, MPI_Datatype, int src, int tag, MPI_Comm, MPI_Status*);
int MPI_Isend(const void*, int, MPI_Datatype, int dest, int tag, MPI_Comm, MPI_Request*);
int MPI_Issend(const void*, int, MPI_Datatype, int dest, int tag, MPI_Comm, MPI_Request*);
int MPI_Irecv(void*, int, MPI_Datatype, int src, int tag, MPI_Comm, MPI_Request*);
int MPI_Sendrecv(const void*, int, MPI_Datatype, int dest, int stag, void*, int, MPI_Datatype, int src, int rtag, MPI_Comm, MPI_Status*);
int MPI_Wait(MPI_Request*, MPI_Status*);
int MPI_Waitall(int, MPI_Request*, MPI_Status*);
int MPI_Test(MPI_Request*, int* flag, MPI_Status*);
int MPI_Get_count(const MPI_Status*, MPI_Datatype, int*);

/* MPI-IO over POSIX files */
int MPI_File_open(MPI_Comm, const char*, int mode, MPI_Info, MPI_File*);
int MPI_File_close(MPI_File*);
int MPI_File_read_at(MPI_File, MPI_Offset, void*, int, MPI_Datatype, MPI_Status*);
int MPI_File_set_view(MPI_File, MPI_Offset, MPI_Datatype, MPI_Datatype, const char*, MPI_Info);
int MPI_File_write(MPI_File, const void*, int, MPI_Datatype, MPI_Status*);
int MPI_File_write_all(MPI_File, const void*, int, MPI_Datatype, MPI_Status*);
int MPI_Info_create(MPI_Info*);
int MPI_Info_set(MPI_Info, const char*, const char*);
int MPI_Info_free(MPI_Info*);

/* one-sided: declared for name lookup, abort when called */
int MPI_Win_create(void*, MPI_Aint, int, MPI_Info, MPI_Comm, MPI_Win*);
int MPI_Win_free(MPI_Win*);
int MPI_Win_fence(int, MPI_Win);
int MPI_Win_lock(int, int, int, MPI_Win);
int MPI_Win_unlock(int, MPI_Win);
int MPI_Win_start(MPI_Group, int, MPI_Win);
int MPI_Win_post(MPI_Group, int, MPI_Win);
int MPI_Win_wait(MPI_Win);
int MPI_Win_complete(MPI_Win);
int MPI_Put(const void*, int, MPI_Datatype, int, MPI_Aint, int, MPI_Datatype, MPI_Win);
int MPI_Get(void*, int, MPI_Datatype, int, MPI_Aint, int, MPI_Datatype, MPI_Win);

#ifdef __cplusplus
}
#endif
#endif
