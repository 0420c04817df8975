/* combblas_b200.h - C ABI of the B200-native SpMM path (libcombblas_b200.so).
 *
 * This is the drop-in boundary for ONE path of Combinatorial BLAS: the 2D-distributed
 * sparse x tall-skinny-dense multiply  Y = A (x).(+) X  under a semiring.  The templated
 * C++ host layer (combblas-spmm-test_b200/include/CombBLAS/) keeps the reference's
 * SpParMat / CommGrid / semiring surface and lowers onto these calls; nothing but plain
 * pointers, sizes and enums crosses this line.  Every entry point names the reference
 * interface it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - every function returns 0 (CB_OK) or a cb_status code; cb_last_error() gives the text.
 *     Codes 3001..3007 are the reference's MPI_Abort codes (include/CombBLAS/SpDefs.h:72-78).
 *   - a ctx owns one CUDA device, one compute stream and one communication stream; it is used
 *     by one host thread at a time (the reference funnels all MPI calls through the master
 *     thread, ReleaseTests/GenWriteMatrix.cpp:65-77).
 *   - host arrays are borrowed for the duration of the call only; every *_upload/_alloc handle
 *     is released by the matching *_free (the reference's raw new/delete ownership,
 *     include/CombBLAS/SpParMat.cpp:110-113).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     CB_ERR_NO_DEVICE.
 */
#ifndef COMBBLAS_B200_H
#define COMBBLAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_ABI_VERSION 1

typedef struct cb_ctx cb_ctx;     /* device + streams + (optional) pr x pc process grid */
typedef struct cb_tile cb_tile;   /* device-resident local sparse tile (doubly compressed rows) */
typedef struct cb_dense cb_dense; /* device-resident dense panel, row-major */
typedef struct cb_coo cb_coo;     /* device-resident result of a sparse x sparse product: merged triples, column-major */

/* element types.  CB_U8 carries C++ bool (one byte, 0/1).  CB_PATTERN as a tile value type means
 * "no value array, every stored entry is true/1" (what SpDCCols<IT,bool> holds after a pattern read). */
typedef enum { CB_F32 = 0, CB_F64 = 1, CB_I32 = 2, CB_I64 = 3, CB_U8 = 4, CB_PATTERN = 255 } cb_dtype;

/* semirings of include/CombBLAS/Semirings.h that cross the ABI as opcodes:
 *   CB_PLUS_TIMES  PlusTimesSRing<T1,T2>  :212-232   id 0, add +, multiply (T)a*(T)b
 *   CB_MIN_PLUS    MinPlusSRing<T,T>      :235-255   id numeric max, add min, multiply inf_plus (:40-47)
 *   CB_MAX_SEL2ND  SelectMaxSRing<bool,T> :191-210   id -1, add max, multiply(a,x)=x
 *   CB_OR_AND      PlusTimesSRing<bool,bool> (promote.h:78): add OR, multiply AND, id false */
typedef enum { CB_PLUS_TIMES = 0, CB_MIN_PLUS = 1, CB_MAX_SEL2ND = 2, CB_OR_AND = 3 } cb_semiring;

typedef enum {
    CB_OK = 0,
    CB_ERR_NO_DEVICE = 1,     /* no CUDA device / driver: the product has no CPU path */
    CB_ERR_CUDA = 2,
    CB_ERR_NCCL = 3,
    CB_ERR_UNSUPPORTED = 4,   /* semiring/dtype combination outside the ABI (compile-time error in the C++ layer) */
    CB_ERR_ALLOC = 5,
    CB_ERR_TOO_LARGE = 6,     /* a tile array has >= 2^31 elements (the reference's MPI int counts, SpParHelper.cpp:595) */
    CB_ERR_GRIDMISMATCH = 3001,
    CB_ERR_DIMMISMATCH = 3002,
    CB_ERR_NOTSQUARE = 3003,
    CB_ERR_NOFILE = 3004,
    CB_ERR_MATRIXALIAS = 3005,
    CB_ERR_INVALIDPARAMS = 3007
} cb_status;

/* ------------------------------------------------------------------ context / process grid
 * Replaces CommGrid(MPI_Comm, nrowproc, ncolproc), include/CombBLAS/CommGrid.h:47 and
 * src/CommGrid.cpp:37-75: rank -> (myprocrow = rank / pc, myproccol = rank % pc), a row
 * communicator (same myprocrow) and a column communicator (same myproccol).  One process per GPU;
 * the 128-byte NCCL unique id is produced by rank 0 with cb_comm_unique_id() and handed to the
 * other ranks by whatever launcher the host has (torch.distributed store, files, MPI). */
int cb_abi_version(void);
int cb_device_count(int* count);
int cb_comm_unique_id(void* id128);
int cb_ctx_create(int device, cb_ctx** ctx);                                   /* 1 x 1 grid */
int cb_ctx_create_grid(int device, int rank, int nranks, int pr, int pc, const void* id128, cb_ctx** ctx);
int cb_ctx_destroy(cb_ctx* ctx);
int cb_ctx_grid(const cb_ctx* ctx, int* rank, int* pr, int* pc, int* myprocrow, int* myproccol);
int cb_ctx_sync(cb_ctx* ctx);                                                  /* wait for both streams */
void* cb_ctx_stream(cb_ctx* ctx);                                              /* cudaStream_t of the compute stream */
const char* cb_last_error(const cb_ctx* ctx);                                  /* ctx may be NULL: last error of this thread */
const char* cb_status_string(int status);

/* device timers on the compute stream (replaces the MPI_Wtime brackets of ReleaseTests/MultTiming.cpp:58-92) */
int cb_timer_start(cb_ctx* ctx);
int cb_timer_stop(cb_ctx* ctx, float* milliseconds);    /* synchronises on the stop event */

/* ------------------------------------------------------------------ local sparse tile
 * Replaces the tile wire format Arr<IT,NT> = {cp, jc, ir | numx} with essentials {nnz, m, n, nzc}
 * (include/CombBLAS/SpDCCols.cpp:787-795,826-846; dcsc.h:124-131; CSC twin csc.h:71-75) and the
 * column->row transposition the reference would do with SpDCCols(const SpTuples&, bool)
 * (SpDCCols.cpp:108-184).  Input: column compressed, rows ascending inside a column.
 *   jc == NULL : plain CSC, cp has n+1 entries and nzc is ignored
 *   idx_dtype  : CB_I32 or CB_I64 (type of cp/jc/ir)
 *   val_dtype  : type of numx, or CB_PATTERN with numx == NULL
 * The tile is converted on the device into doubly compressed rows with a work partition. */
int cb_tile_upload_csc(cb_ctx* ctx, int64_t m, int64_t n, int64_t nz, int64_t nzc, const void* cp, const void* jc,
                       const void* ir, const void* numx, int idx_dtype, int val_dtype, cb_tile** tile);
/* same from unsorted local COO triples (what SpTuples holds, include/CombBLAS/SpTuples.h:64-);
 * duplicates must already be merged. */
int cb_tile_upload_coo(cb_ctx* ctx, int64_t m, int64_t n, int64_t nz, const void* rows, const void* cols,
                       const void* vals, int idx_dtype, int val_dtype, cb_tile** tile);
/* same from COO triples that already live on the device (int64 indices); used by the generators */
int cb_tile_from_device_coo(cb_ctx* ctx, int64_t m, int64_t n, int64_t nz, const int64_t* d_rows, const int64_t* d_cols,
                            const void* d_vals, int val_dtype, cb_tile** tile);
/* distributed ingestion.  Replaces the distribution half of SpParMat::ParallelReadMM and SparseCommon
 * (include/CombBLAS/SpParMat.cpp:3978-4115, :2891-2968: every process parses its byte range of the file, MPI_Alltoallv sends the
 * triples to their owners, SpTuples::RemoveDuplicates(BinOp) merges, the local tile is built): every rank hands in the triples IT
 * parsed (host arrays, global 0-based coordinates, int64; owned by anybody), they are routed to their owners between the GPUs
 * (one grouped ncclSend / ncclRecv exchange), duplicates are merged with dup_op (0 keep the first, 1 sum, 2 max, 3 min) and this
 * rank's tile is built on the device.  Collective over the grid. */
int cb_tile_from_distributed_coo(cb_ctx* ctx, int64_t gm, int64_t gn, int64_t nz, const int64_t* rows, const int64_t* cols, const void* vals,
                                 int val_dtype, int dup_op, cb_tile** tile);
/* the same with the PARSING on the device too.  Replaces SpParMat::ParallelReadMM's per-process text handling
 * (include/CombBLAS/SpParMat.cpp:4010-4095 + SpHelper.h:75-91,147-183): `text` is this rank's share of the data section of a
 * Matrix Market coordinate file - the bytes of the lines that start inside its byte range, host memory - and is cut into lines
 * and parsed on the GPU, one thread per line.  flags: 1 = the indices are one-based, 2 = pattern file (no value column, every
 * entry 1), 4 = symmetric / hermitian (the transpose of every off-diagonal entry is added).  Blank lines and lines that do not
 * start with two integers are skipped, like the reference's sscanf loop does.  Values are converted to val_dtype with a C cast
 * (CB_U8: value != 0).  Returns CB_ERR_UNSUPPORTED on every rank, before anything is exchanged, when any share holds a number the
 * device parser does not convert (more than 19 significant digits, a subnormal or overflowing value, inf / nan, hexadecimal) or
 * is 1 GiB or larger (32-bit positions);
 * everything else is the correctly rounded double strtod gives.  The caller then parses that file on the host and uses
 * cb_tile_from_distributed_coo.  Collective. */
int cb_tile_from_mm_text(cb_ctx* ctx, int64_t gm, int64_t gn, const char* text, int64_t nbytes, int flags, int val_dtype, int dup_op,
                         cb_tile** tile);
int cb_tile_free(cb_tile* tile);
/* {nnz, m, n, nonempty rows, nonempty columns, work chunks, split rows, bytes resident} */
int cb_tile_info(const cb_tile* tile, int64_t info[8]);
/* a second handle onto the same device arrays that ignores the values: the structure of the tile as a CB_PATTERN
 * matrix (used to compute the nonzero structure of a product with the boolean semiring).  The view borrows the
 * original's memory: free it before the original; freeing the view does not free the arrays. */
int cb_tile_pattern_view(const cb_tile* tile, cb_tile** view);
/* a new tile of the same shape holding only the nonzeros whose column c has keep_cols[c] != 0 (host array, one byte per
 * column of the tile).  For products with a SPARSE right-hand side B (Mult_AnXBn_Synch with a sparse B,
 * include/CombBLAS/ParFriends.h:1004-1108): columns of A that meet only empty rows of B cannot contribute, and the reference's
 * column-by-column SpGEMM never touches them (mtSpGEMM.h:292-441); dropping them makes the dense-panel engine cost what the
 * product touches instead of nnz(A) x k.  Opt-in in the host layer (CB_SPGEMM_FILTER=1). */
int cb_tile_filter_columns(cb_ctx* ctx, const cb_tile* tile, const uint8_t* keep_cols, cb_tile** out);
/* copy the device tile back as CSR (rowptr[m+1], colidx[nnz], vals) for inspection/tests; any pointer may be NULL */
int cb_tile_download_csr(cb_tile* tile, int64_t* rowptr, int64_t* colidx, void* vals);

/* row-level inspection for parity checks at sizes where copying a whole operand back is too much (bench.py's sampled-row
 * check, tests at BASELINE.json's full sizes) - the role of SpParMat::operator() / SpRef on a handful of rows
 * (include/CombBLAS/SpParMat.cpp:1743-): lengths of all m rows; then the (column, value) pairs of selected rows, concatenated
 * in the order asked for, columns ascending inside a row (vals may be NULL). */
int cb_tile_row_lengths(cb_tile* tile, int64_t* len);
int cb_tile_download_rows(cb_tile* tile, int64_t nrows, const int64_t* rows, int64_t* cols, void* vals);

/* ------------------------------------------------------------------ dense panel
 * Replaces DenseParMat<IT,NT>'s local block (include/CombBLAS/DenseParMat.h:49-128; the reference
 * allocates NT** with one new[] per row, SpHelper.h:241-247) by one contiguous row-major device
 * allocation whose leading dimension is padded to 16 bytes. */
int cb_dense_alloc(cb_ctx* ctx, int64_t rows, int64_t cols, int dtype, cb_dense** d);
int cb_dense_wrap(cb_ctx* ctx, void* device_ptr, int64_t rows, int64_t cols, int64_t ld, int dtype, cb_dense** d);
int cb_dense_free(cb_dense* d);
int cb_dense_upload(cb_dense* d, const void* host, int64_t ld_host);     /* async on the compute stream if host is pinned */
int cb_dense_download(cb_dense* d, void* host, int64_t ld_host);         /* returns after the copy completed */
int cb_dense_download_rows(cb_dense* d, int64_t nrows, const int64_t* rows, void* host);   /* host[i, 0:cols] = d[rows[i], 0:cols], packed */
int cb_dense_fill(cb_dense* d, const void* scalar);                      /* std::fill_n(localy, ysize, SR::id()), ParFriends.h:1960-1963 */
int cb_dense_info(const cb_dense* d, int64_t* rows, int64_t* cols, int64_t* ld, int* dtype, void** device_ptr);
int cb_semiring_id(int semiring, int dtype, void* scalar_out);           /* SR::id() as a value of dtype */

/* ------------------------------------------------------------------ the multiply
 * cb_spmm_local: Y (+)= tile (x).(+) X on this GPU, enqueued on the compute stream.
 *   accumulate == 0 : Y = product; rows of the tile without nonzeros receive SR::id()
 *   accumulate != 0 : Y = SR::add(Y, product) on rows with nonzeros (later SUMMA stages)
 * Replaces LocalHybridSpGEMM<SR,NTO> restricted to a dense right-hand side
 * (include/CombBLAS/mtSpGEMM.h:213-460), dcsc_gespmv generalised to k columns
 * (include/CombBLAS/Friends.h:63-78) and, through `accumulate`, MultiwayMerge
 * (include/CombBLAS/MultiwayMerge.h:411-526).  X.rows == tile.n, Y.rows == tile.m, X.cols == Y.cols,
 * dtype(Y) == dtype(X) == promote_trait<NT_A, NT_X> (promote.h:37-91). */
int cb_spmm_local(cb_ctx* ctx, const cb_tile* tile, const cb_dense* X, cb_dense* Y, int semiring, int accumulate);

/* cb_spmm_summa: the whole 2D SUMMA stage loop, collective over the ctx's grid.
 * Replaces Mult_AnXBn_Synch / _Overlap (include/CombBLAS/ParFriends.h:1004-1108, :1110-1235) with
 * its SpParHelper::BCastMatrix / GetSetSizes transport (SpParHelper.cpp:582-601, :797-809):
 * per stage the owner's A sub-tile travels along the process row and the owner's X panel along the
 * process column (NCCL on the communication stream, double buffered), overlapped with
 * cb_spmm_local of the previous stage; the Y tile is stationary.
 *   tile : this rank's A tile, rows of block-row myprocrow x columns of block-column myproccol
 *   X    : this rank's X tile, rows of block-row myprocrow of n  x  columns of block myproccol of k
 *   Y    : this rank's Y tile, rows of block-row myprocrow of m  x  columns of block myproccol of k
 *   gm, gn, gk : global dimensions (Owner rule of SpParMat.cpp:5066-5096 gives every block range)
 * Works on any pr x pc (the reference only on square grids, src/CommGrid.cpp:164-180). */
int cb_spmm_summa(cb_ctx* ctx, const cb_tile* tile, const cb_dense* X, cb_dense* Y, int semiring,
                  int64_t gm, int64_t gn, int64_t gk);
/* Small host-side reductions over the grid's communicators, for the metadata the reference gets with MPI_Allreduce
 * (SpParMat::getnnz/getnrow/getncol, include/CombBLAS/SpParMat.cpp:773-797).  which: 0 = world, 1 = processor row
 * ("RowWorld"), 2 = processor column ("ColWorld"); op: 0 = sum, 1 = max, 2 = min.  Collective. */
int cb_comm_allreduce_i64(cb_ctx* ctx, int which, int op, int64_t* inout, int count);
/* A-part caching (default on).  Tiles are immutable, so the parts of A a rank receives from its row neighbours
 * during the first cb_spmm_summa with a tile are kept in its HBM (one block-row of A per GPU: a few GB of 180 GB)
 * and later multiplies with the same tile only move the dense panels.  The reference re-broadcasts A on every call
 * (ParFriends.h:1036-1052); cb_summa_cache_a(ctx, 0) restores that behaviour (double-buffered receive slots). */
int cb_summa_cache_a(cb_ctx* ctx, int on);
/* device times of the last cb_spmm_summa on this rank: {whole stage loop on the compute stream (ms), time the
 * communication stream spent in the stage broadcasts (ms, overlaps the kernels), reserved, number of stages} */
int cb_summa_times(cb_ctx* ctx, float ms[4]);
/* The stage plan cb_spmm_summa uses, as pure host arithmetic (no device needed): [0, gn) is cut at every boundary
 * of A's pc column blocks and of X's pr row blocks (floor rule of SpParMat.cpp:5066-5096).  seg needs pr+pc+1
 * entries, the owner arrays pr+pc.  Stage s covers inner indices [seg[s], seg[s+1]); its A part is broadcast along
 * each processor row from grid column a_owner_col[s], its X panel along each processor column from grid row
 * x_owner_row[s].  On a square grid with pr | gn this is the reference's `stages = grcols` loop (ParFriends.h:1036). */
int cb_summa_plan(int pr, int pc, int64_t gn, int64_t* seg, int* a_owner_col, int* x_owner_row, int* nstages);

/* one-shot convenience with HOST operands: upload X, multiply with a resident tile, download Y.
 * This is what SpMM<SR>(A, X) of the C++ layer calls when the panels live in host memory. */
int cb_spmm_host(cb_ctx* ctx, const cb_tile* tile, const void* X_host, int64_t ldx, void* Y_host, int64_t ldy,
                 int64_t k, int dtype, int semiring);

/* cb_spmm_summa with HOST panels: this rank's X tile is read from host memory and its Y tile written to host memory (row-major,
 * leading dimensions in elements); dtype is the panels' element type.  The k-block is cut into column slabs that flow up,
 * through the stage loop and down on three streams, so both PCIe directions overlap the multiply.  Collective over the grid.
 * This is what SpMM<SR>(A, X) of the C++ layer calls: the distributed counterpart of cb_spmm_host. */
int cb_spmm_summa_host(cb_ctx* ctx, const cb_tile* tile, const void* X_host, int64_t ldx, void* Y_host, int64_t ldy, int semiring,
                       int64_t gm, int64_t gn, int64_t gk, int dtype);

/* y = A (x).(+) x for a dense vector distributed over the grid, with the vector exchange on the device.  Replaces
 * SpMV<SR>(SpParMat, FullyDistVec) (include/CombBLAS/ParFriends.h:1924-1996): TransposeVector + MPI_Allgatherv on the processor
 * column, dcsc_gespmv on an id()-filled y (Friends.h:63-78), MPI_Reduce with SR::mpi_op on the processor row.  The pieces of x
 * (host, at global offset x_off, x_len elements; the pieces of all ranks tile [0, gn) in any order) go up once, travel between
 * the GPUs over NCCL, the local multiply is the SpMM kernel on a one-column panel, the partial results are combined along the
 * processor row with ncclAllReduce (sum / min / max) and this rank's piece of y (inside its row block of gm) comes down.
 * Collective over the grid. */
int cb_spmv_grid(cb_ctx* ctx, const cb_tile* tile, const void* x_piece, int64_t x_off, int64_t x_len, void* y_piece, int64_t y_off, int64_t y_len,
                 int semiring, int dtype, int64_t gm, int64_t gn);

/* ------------------------------------------------------------------ sparse x SPARSE (tall-skinny right-hand side)
 * The literal Mult_AnXBn_Synch<SR,NUO,UDERO>(A, B) / PSpGEMM<SR>(A, B) of the reference (include/CombBLAS/ParFriends.h:1004-1108,
 * SpParMat.h:454-467) - the call of Applications/SpMMError.cpp:83 and Applications/BetwCent.cpp:185,204 - on the device.
 * Replaces LocalHybridSpGEMM's per-column hash / heap accumulator (mtSpGEMM.h:213-460) and MultiwayMerge
 * (MultiwayMerge.h:411-526) by expand - stable sort - reduce-by-key (csrc/cb_spgemm.cu): cost is what the product touches, the
 * products of one output entry are folded in ascending inner index like the reference's hash branch (mtSpGEMM.h:395-423), and an
 * entry exists exactly where the reference creates one.
 *   B     : a tile (cb_tile_upload_csc / _coo) whose values already have the product's type `dtype` (promote_trait), or a pattern
 *   dtype : element type of the product; A's values must be of that type, CB_U8 (bool) or a pattern
 * cb_spgemm_local multiplies two tiles on this GPU; cb_spgemm_summa is the stage loop over the ctx's grid (any pr x pc): per
 * stage the owner's A part travels along the processor row and the owner's B tile along the processor column as one message
 * each, the partial products of all stages are merged once at the end.  The result is this rank's block of C as triples with
 * local indices, sorted by column then row - the order SpDCCols' tuple constructor wants (SpDCCols.cpp:186-195). */
int cb_spgemm_local(cb_ctx* ctx, const cb_tile* A, const cb_tile* B, int semiring, int dtype, cb_coo** C);
int cb_spgemm_summa(cb_ctx* ctx, const cb_tile* A, const cb_tile* B, int semiring, int dtype, int64_t gm, int64_t gn, int64_t gk, cb_coo** C);
int cb_coo_info(const cb_coo* c, int64_t* nnz, int64_t* m, int64_t* k, int* dtype);
int cb_coo_download(cb_coo* c, int64_t* rows, int64_t* cols, void* vals);        /* any pointer may be NULL */
int cb_coo_free(cb_coo* c);

/* Hub variant of the local multiply (K2H, csrc/cb_spmm_hub_kernel.cuh) - OPT-IN, off by default.
 * For tiles whose columns are very unevenly used (R-MAT / power-law inputs) the panel rows of the most frequent columns
 * are kept in the shared memory of thread-block clusters for the whole multiply, so that share of the row gathers no
 * longer crosses the L2 slices.  Results are bit-identical to the default kernel (same walk, same fold order).  It has no
 * counterpart in the reference: it is a faster LocalHybridSpGEMM-with-dense-rhs (include/CombBLAS/mtSpGEMM.h:213-460).
 *   enable     : 1 on, 0 off, -1 follow the environment (CB_SPMM_HUB=1)
 *   cluster    : CTAs pooling their shared memory: 1, 2, 4, 8, or 0 for the default (CB_SPMM_HUB_CLUSTER, 4)
 *   slab_bytes : bytes of a panel row handled per column slab: 128, 256, 512, or 0 to choose by row width
 * cb_spmm_hub_info: {hub data built, hub columns known, hub rows resident in the last multiply, share of the tile's
 * nonzeros they serve x 1e6}. */
int cb_spmm_hub_config(cb_ctx* ctx, int enable, int cluster, int slab_bytes);
int cb_spmm_hub_info(const cb_tile* tile, int64_t info[4]);
/* Ring variant of the local multiply (K2R, same file) - OPT-IN, off by default, combinable with the hub variant.
 * The row gathers are pipelined through a per-virtual-warp ring in shared memory with cp.async (no destination registers),
 * so an SM keeps threads x depth x 16 bytes of gathers in flight instead of what its register file allows.  Results are
 * bit-identical to the default kernel.  depth: 8 on (the depth this build instantiates), 0 off, -1 follow CB_SPMM_RING. */
int cb_spmm_ring_config(cb_ctx* ctx, int depth);
/* Shape of the default local multiply (K2), for tuning: slab_bytes > 0 runs panels wider than that as column slabs of that
 * width, one after the other, so the X rows one pass gathers are narrower and more of them stay in L2 (0 = one slab as wide
 * as the layout allows; 64, 128, 256, 512); point = 0 deep / 1 wide / -1 chosen from the footprint of the X rows.  Results
 * do not depend on either (columns are independent). */
int cb_spmm_k2_config(cb_ctx* ctx, int slab_bytes, int point);
/* Variant of the local multiply, for experiments (results are identical bit for bit): 0 = K2 (gathers in groups, the default),
 * 1 = K2 with the entry prefetch, 16 = K2T (every row gather one cp.async.bulk copy into a per-warp shared-memory ring with
 * mbarrier completion, csrc/cb_spmm_tma_kernel.cuh; fp32 / int32 panels of 128-, 256- or 512-byte rows, anything else runs K2),
 * 32 = K2W (the rows of the most used columns - cb_spmm_k2_l2's budget, default 64 MB - packed into a panel that a persisting L2
 * access-policy window keeps on the chip; fp32 / fp64 PlusTimes panels of 256-byte rows and wider, anything else runs K2),
 * 64 = K2 with persistent warps (one launch fills the chip, warps take chunk groups from a counter; fp32 / fp64 / int32 panels of one
 * column slab, anything else runs K2), 4 / 8 = ring depth of K2P (register ring; only in builds with -DCB_BUILD_K2P), -1 = the build's default. */
int cb_spmm_k2_pipe(cb_ctx* ctx, int depth);
/* L2 residency hints of K2P for tiles whose X rows do not fit in L2: the rows of the tile's most used columns - as many as
 * fit `budget_mb` megabytes at the current row width - are gathered with the evict_last priority, every other row with
 * evict_first, so rows that are used once no longer push the shared ones out.  0 = off, -1 = the build's default. */
int cb_spmm_k2_l2(cb_ctx* ctx, int budget_mb);
/* The hub selection rule as pure host arithmetic (no device needed): the max_hubs most frequent columns with at least
 * two nonzeros, most frequent first, ties by ascending column; cum[r] = nonzeros in the columns of rank <= r.
 * Returns the number of hubs written, or -1 on bad arguments. */
int cb_hub_select_host(const int32_t* counts, int64_t n, int max_hubs, int32_t* hubcols, int64_t* cum);

/* number of kernels this library has launched on the ctx since creation (bench.py's gpu_launches) */
int64_t cb_launch_count(const cb_ctx* ctx);
/* per-kernel device timing: while enabled every K3 (identity fill) / K2 (multiply) / fix-up launch on the
 * compute stream is bracketed by CUDA events; read returns the summed milliseconds and launch counts per
 * kind {fill, multiply, fix-up} since enable.  Plays the role of the reference's -DTIMING accumulators
 * (include/CombBLAS/CombBLAS.h:76-102). */
int cb_profile_enable(cb_ctx* ctx, int on);
int cb_profile_read(cb_ctx* ctx, double ms[3], int64_t launches[3]);

/* ------------------------------------------------------------------ synthetic operands on the device
 * Replaces the host generators DistEdgeList::GenGraph500Data (include/CombBLAS/DistEdgeList.cpp:223-)
 * + SpParMat(const DistEdgeList&, bool) (SpParMat.cpp:3138-3254) with the recipe of
 * ReleaseTests/GenWriteMatrix.cpp:96-131: Kronecker edges with the given initiator, vertex scramble,
 * self loops removed, optional A += A^T, duplicates merged.  Counter based: edge e and entry (i,j)
 * depend only on (seed, e) / (seed, i, j), never on the grid.  Only the block of the matrix that
 * falls into rows [row0,row0+m) x cols [col0,col0+n) is kept (local indices).
 * val_dtype: CB_PATTERN, or a dtype whose values come from the counter hash with val_seed. */
int cb_gen_rmat_tile(cb_ctx* ctx, int scale, int edgefactor, uint64_t seed, const double initiator[4], int symmetric,
                     int64_t row0, int64_t m, int64_t col0, int64_t n, int val_dtype, uint64_t val_seed, cb_tile** tile);
/* The reference's OWN edge stream: the Graph500 2.1 Kronecker generator as DistEdgeList::GenGraph500Data drives it with
 * packed = true (include/CombBLAS/RefGen21.h:88-301, DistEdgeList.cpp:223-236) - the stream ReleaseTests/GenWriteMatrix.cpp
 * builds its matrices from - restated for the device and checked bit for bit against the reference's generator.
 * (userseed1, userseed2) = (0, 0) is the reference's -DDETERMINISTIC seed (RefGen21.h:274-275).
 * cb_gen_graph500_edges: edges [first, first+count) as they leave the generator (global ids, duplicates and loops included).
 * cb_gen_graph500_tile : the block of the matrix SpParMat(DistEdgeList, removeloops) [+ Symmetricize] builds from
 *   edgefactor * 2^scale such edges; val_seed == 0 gives the reference's values (duplicates summed: multiplicities),
 *   any other val_seed hashed weights as cb_gen_rmat_tile does. */
int cb_gen_graph500_edges(cb_ctx* ctx, int log_numverts, uint64_t userseed1, uint64_t userseed2, int64_t first, int64_t count,
                          int64_t* src_host, int64_t* dst_host);
int cb_gen_graph500_tile(cb_ctx* ctx, int scale, int edgefactor, uint64_t userseed1, uint64_t userseed2, int symmetric, int remove_loops,
                         int64_t row0, int64_t m, int64_t col0, int64_t n, int val_dtype, uint64_t val_seed, cb_tile** tile);
/* X[i,j] = value(seed, (row0+i)*gk + col0+j); kind 0 = uniform (0,1) / [1,100] / Bernoulli, kind 1 = MinPlus operand
 * with ~1% entries at numeric max */
int cb_gen_dense(cb_dense* d, uint64_t seed, int64_t row0, int64_t col0, int64_t gk, int kind);

#ifdef __cplusplus
}
#endif
#endif /* COMBBLAS_B200_H */
