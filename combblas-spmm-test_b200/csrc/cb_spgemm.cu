// Sparse x SPARSE tall-skinny product on the device: the literal Mult_AnXBn_Synch / PSpGEMM of the reference
// (include/CombBLAS/ParFriends.h:1004-1108) for a right-hand side that is itself sparse - the call of
// Applications/SpMMError.cpp:83 and of Applications/BetwCent.cpp:185,204 (fringe and back-propagation products).
//
// Replaces LocalHybridSpGEMM's per-column hash / heap accumulator (include/CombBLAS/mtSpGEMM.h:213-460) and MultiwayMerge
// (include/CombBLAS/MultiwayMerge.h:411-526) by expand - sort - compress on the GPU:
//   * expand   : every nonzero A(i,kk) whose row kk of B is not empty emits A(i,kk) (x) B(kk,j) for the nonzeros j of that row,
//                keyed  (j << 32) | i  - column-major, the order SpDCCols' tuple constructor wants (SpDCCols.cpp:186-195).
//                Two passes (count, exclusive scan, fill): the product list is exactly as long as the flops.
//   * sort     : stable LSD radix sort of (key, product) pairs (CUB).  Products of one output entry were generated in
//                ascending kk (A's rows are column sorted) and a stable sort keeps that order - the order in which the
//                reference's hash accumulator folds them (mtSpGEMM.h:395-423).
//   * compress : reduce-by-key with the semiring's add.  An entry exists exactly where the reference creates one - also
//                when the folded value equals SR::id() - because keys, not values, decide.
// The partial products of all SUMMA stages go through ONE sort + compress at the end, which is what MultiwayMerge does to
// the per-stage lists (equal (row, column) merged with SR::add).  Cost: what the product touches, not nnz(A) x k.
#include <algorithm>
#include <cub/cub.cuh>
#include "cb_common.cuh"
#include "cb_spgemm.cuh"

namespace {

inline int grid_for(int64_t n, int sm) {
    int64_t b = (n + 255) / 256, cap = (int64_t)sm * 32;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

enum { AK_SAME = 0, AK_PATTERN = 1, AK_BOOL = 2 };

// semiring arithmetic on scalars (Semirings.h:40-47, :191-255); the semiring is a run-time value here: these kernels are bound
// by the sort, not by arithmetic
template <typename T> struct Lim;
template <> struct Lim<float> { static __host__ __device__ float maxv() { return 3.402823466e+38f; } };
template <> struct Lim<double> { static __host__ __device__ double maxv() { return 1.7976931348623157e+308; } };
template <> struct Lim<int32_t> { static __host__ __device__ int32_t maxv() { return 0x7fffffff; } };
template <> struct Lim<int64_t> { static __host__ __device__ int64_t maxv() { return 0x7fffffffffffffffLL; } };
template <> struct Lim<uint8_t> { static __host__ __device__ uint8_t maxv() { return 255; } };

template <typename T> __host__ __device__ inline T wrap_add(T a, T b) { return a + b; }
template <> __host__ __device__ inline int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
template <> __host__ __device__ inline int64_t wrap_add(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
template <typename T> __host__ __device__ inline T wrap_mul(T a, T b) { return a * b; }
template <> __host__ __device__ inline int32_t wrap_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
template <> __host__ __device__ inline int64_t wrap_mul(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }

template <typename T>
__host__ __device__ inline T sr_mul(int sr, T a, T b) {
    switch (sr) {
        case CB_MIN_PLUS: return (a == Lim<T>::maxv() || b == Lim<T>::maxv()) ? Lim<T>::maxv() : wrap_add(a, b);
        case CB_MAX_SEL2ND: return b;
        case CB_OR_AND: return (T)((a != T(0)) && (b != T(0)));
        default: return wrap_mul(a, b);
    }
}
template <typename T>
struct SrAdd {          // SR::add(product, acc): every supported add is commutative and associative up to rounding
    int sr;
    __host__ __device__ T operator()(const T& a, const T& b) const {
        switch (sr) {
            case CB_MIN_PLUS: return b < a ? b : a;
            case CB_MAX_SEL2ND: return a < b ? b : a;
            case CB_OR_AND: return (T)((a != T(0)) || (b != T(0)));
            default: return wrap_add(a, b);
        }
    }
};

// row kk of B -> index into B's compressed row list, -1 for an empty row
__global__ void __launch_bounds__(256)
brow_index_kernel(const int32_t* __restrict__ nzrows, int64_t nzr, int32_t* __restrict__ index) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nzr; i += (int64_t)gridDim.x * blockDim.x) index[nzrows[i]] = (int32_t)i;
}

// products each nonzero of A will emit: the length of row (col + boff) of B
__global__ void __launch_bounds__(256)
count_kernel(const int32_t* __restrict__ a_colflag, int64_t a_nnz, int64_t boff, int64_t b_rows, const int32_t* __restrict__ b_index,
             const int32_t* __restrict__ b_rowptr, uint32_t* __restrict__ count) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < a_nnz; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t kk = (int64_t)(a_colflag[p] & 0x7fffffff) + boff;
        uint32_t c = 0;
        if (kk >= 0 && kk < b_rows) {
            const int32_t bi = b_index[kk];
            if (bi >= 0) c = (uint32_t)(b_rowptr[bi + 1] - b_rowptr[bi]);
        }
        count[p] = c;
    }
}

template <typename T, int AK>
__global__ void __launch_bounds__(256)
expand_kernel(const int32_t* __restrict__ a_colflag, const void* __restrict__ a_vals, const int32_t* __restrict__ a_rowptr,
              const int32_t* __restrict__ a_nzrows, int64_t a_nzr, int64_t a_nnz, int64_t boff, const int32_t* __restrict__ b_index,
              const int32_t* __restrict__ b_rowptr, const int32_t* __restrict__ b_colflag, const T* __restrict__ b_vals,
              const uint32_t* __restrict__ count, const uint64_t* __restrict__ offset, int sr, uint64_t* __restrict__ keys, T* __restrict__ prods) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < a_nnz; p += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t c = count[p];
        if (c == 0) continue;
        int64_t lo = 0, hi = a_nzr;           // row of nonzero p: largest ridx with rowptr[ridx] <= p
        while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (a_rowptr[mid] <= p) lo = mid; else hi = mid; }
        const uint64_t i = (uint64_t)(uint32_t)a_nzrows[lo];
        T a;
        if (AK == AK_PATTERN) a = T(1);
        else if (AK == AK_BOOL) a = (T)(reinterpret_cast<const uint8_t*>(a_vals)[p] != 0);
        else a = reinterpret_cast<const T*>(a_vals)[p];
        const int32_t bi = b_index[(int64_t)(a_colflag[p] & 0x7fffffff) + boff];
        const int32_t bs = b_rowptr[bi];
        uint64_t o = offset[p];
        for (uint32_t q = 0; q < c; ++q, ++o) {
            const uint64_t j = (uint64_t)(uint32_t)(b_colflag[bs + q] & 0x7fffffff);
            keys[o] = (j << 32) | i;
            prods[o] = sr_mul<T>(sr, a, b_vals ? b_vals[bs + q] : T(1));
        }
    }
}

int bits_for(int64_t n) { int b = 1; while ((int64_t(1) << b) < n) ++b; return b; }

template <typename T>
int expand_typed(cb_ctx* ctx, const cb_tile* A, const cb_tile* B, int64_t boff, int semiring, int akind, cb_spgemm_acc* acc) {
    cudaStream_t st = ctx->compute;
    const int sm = ctx->sm_count;
    if (A->nnz == 0 || B->nnz == 0) return CB_OK;
    cb_scratch sc;
    int32_t* b_index = nullptr;
    uint32_t* count = nullptr;
    uint64_t* offset = nullptr;
    CB_CUDA(ctx, sc.alloc(&b_index, (size_t)B->m));
    CB_CUDA(ctx, sc.alloc(&count, (size_t)A->nnz));
    CB_CUDA(ctx, sc.alloc(&offset, (size_t)A->nnz + 1));
    CB_CUDA(ctx, cudaMemsetAsync(b_index, 0xff, (size_t)B->m * sizeof(int32_t), st));
    brow_index_kernel<<<grid_for(B->nzr, sm), 256, 0, st>>>(B->nzrows, B->nzr, b_index);
    count_kernel<<<grid_for(A->nnz, sm), 256, 0, st>>>(A->colflag, A->nnz, boff, B->m, b_index, B->rowptr, count);
    CB_LAUNCHED(ctx); CB_LAUNCHED(ctx);
    size_t tb = 0;
    CB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tb, count, offset, (int)A->nnz, st));
    char* tmp = nullptr;
    CB_CUDA(ctx, sc.alloc(&tmp, tb));
    CB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp, tb, count, offset, (int)A->nnz, st));
    CB_LAUNCHED(ctx);
    uint64_t last_off = 0;
    uint32_t last_cnt = 0;
    CB_CUDA(ctx, cudaMemcpyAsync(&last_off, offset + (A->nnz - 1), sizeof last_off, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaMemcpyAsync(&last_cnt, count + (A->nnz - 1), sizeof last_cnt, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    const uint64_t total = last_off + last_cnt;
    if (total == 0) return CB_OK;
    if (total >= (uint64_t(1) << 31)) return cb_fail(ctx, CB_ERR_TOO_LARGE, "sparse product of one stage has %llu partial products (>= 2^31)", (unsigned long long)total);
    cb_spgemm_acc::Piece piece;
    piece.count = (int64_t)total;
    if (cudaMalloc((void**)&piece.keys, total * sizeof(uint64_t)) != cudaSuccess || cudaMalloc(&piece.vals, total * sizeof(T)) != cudaSuccess) {
        cudaFree(piece.keys);
        cudaGetLastError();
        return cb_fail(ctx, CB_ERR_ALLOC, "cudaMalloc for %llu partial products of a sparse product", (unsigned long long)total);
    }
    acc->pieces.push_back(piece);
    const void* av = A->vals;
    const T* bv = reinterpret_cast<const T*>(B->vals);
#define EXPAND(AKV) expand_kernel<T, AKV><<<grid_for(A->nnz, sm), 256, 0, st>>>(A->colflag, av, A->rowptr, A->nzrows, A->nzr, A->nnz, boff, b_index, \
                        B->rowptr, B->colflag, bv, count, offset, semiring, piece.keys, (T*)piece.vals)
    if (akind == AK_PATTERN) EXPAND(AK_PATTERN);
    else if (akind == AK_BOOL) EXPAND(AK_BOOL);
    else EXPAND(AK_SAME);
#undef EXPAND
    CB_LAUNCHED(ctx);
    CB_CUDA(ctx, cudaGetLastError());
    CB_CUDA(ctx, cudaStreamSynchronize(st));          // the scratch arrays go out of scope
    return CB_OK;
}

template <typename T>
int finish_typed(cb_ctx* ctx, cb_spgemm_acc* acc, int semiring, int64_t m, int64_t k, cb_coo* out) {
    cudaStream_t st = ctx->compute;
    int64_t total = 0;
    for (const auto& p : acc->pieces) total += p.count;
    out->nnz = 0;
    if (total == 0) return CB_OK;
    if (total >= (int64_t(1) << 31)) return cb_fail(ctx, CB_ERR_TOO_LARGE, "sparse product has %lld partial products (>= 2^31)", (long long)total);
    cb_scratch sc;
    uint64_t *keys = nullptr, *keys_sorted = nullptr;
    T *vals = nullptr, *vals_sorted = nullptr;
    if (acc->pieces.size() == 1) {                       // the usual case on one rank: sort straight out of the piece
        keys = acc->pieces[0].keys;
        vals = (T*)acc->pieces[0].vals;
    } else {
        CB_CUDA(ctx, sc.alloc(&keys, (size_t)total));
        CB_CUDA(ctx, sc.alloc(&vals, (size_t)total));
        int64_t off = 0;
        for (const auto& p : acc->pieces) {              // stage order = generation order: kept by the stable sort
            CB_CUDA(ctx, cudaMemcpyAsync(keys + off, p.keys, (size_t)p.count * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
            CB_CUDA(ctx, cudaMemcpyAsync(vals + off, p.vals, (size_t)p.count * sizeof(T), cudaMemcpyDeviceToDevice, st));
            off += p.count;
        }
    }
    CB_CUDA(ctx, sc.alloc(&keys_sorted, (size_t)total));
    CB_CUDA(ctx, sc.alloc(&vals_sorted, (size_t)total));
    size_t tb = 0;
    const int end_bit = 32 + bits_for(k > 1 ? k : 2);
    (void)m;
    CB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys_sorted, vals, vals_sorted, (int)total, 0, end_bit, st));
    char* tmp = nullptr;
    CB_CUDA(ctx, sc.alloc(&tmp, tb));
    CB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys_sorted, vals, vals_sorted, (int)total, 0, end_bit, st));
    CB_LAUNCHED(ctx);
    uint64_t* ukeys = nullptr;
    T* uvals = nullptr;
    int* d_runs = nullptr;
    if (cudaMalloc((void**)&ukeys, (size_t)total * sizeof(uint64_t)) != cudaSuccess || cudaMalloc((void**)&uvals, (size_t)total * sizeof(T)) != cudaSuccess) {
        cudaFree(ukeys);
        cudaGetLastError();
        return cb_fail(ctx, CB_ERR_ALLOC, "cudaMalloc for the %lld entries of a sparse product", (long long)total);
    }
    out->keys = ukeys;
    out->vals = uvals;
    CB_CUDA(ctx, sc.alloc(&d_runs, 1));
    SrAdd<T> add{semiring};
    size_t rb = 0;
    CB_CUDA(ctx, cub::DeviceReduce::ReduceByKey(nullptr, rb, keys_sorted, ukeys, vals_sorted, uvals, d_runs, add, (int)total, st));
    char* tmp2 = nullptr;
    CB_CUDA(ctx, sc.alloc(&tmp2, rb));
    CB_CUDA(ctx, cub::DeviceReduce::ReduceByKey(tmp2, rb, keys_sorted, ukeys, vals_sorted, uvals, d_runs, add, (int)total, st));
    CB_LAUNCHED(ctx);
    int runs = 0;
    CB_CUDA(ctx, cudaMemcpyAsync(&runs, d_runs, sizeof runs, cudaMemcpyDeviceToHost, st));
    CB_CUDA(ctx, cudaStreamSynchronize(st));
    out->nnz = runs;
    return CB_OK;
}

int akind_of(const cb_tile* A, int dtype) {
    if (A->val_dtype == CB_PATTERN) return AK_PATTERN;
    if (A->val_dtype == CB_U8 && dtype != CB_U8) return AK_BOOL;
    if (A->val_dtype == dtype) return AK_SAME;
    return -1;
}

}  // namespace

void cb_spgemm_acc_release(cb_spgemm_acc* acc) {
    for (auto& p : acc->pieces) { cudaFree(p.keys); cudaFree(p.vals); }
    acc->pieces.clear();
}

// partial products of A x B(rows shifted by boff), appended to acc.  dtype = element type of B's values and of the product.
int cb_spgemm_expand(cb_ctx* ctx, const cb_tile* A, const cb_tile* B, int64_t boff, int semiring, int dtype, cb_spgemm_acc* acc) {
    if (semiring == CB_PLUS_TIMES && dtype == CB_U8) semiring = CB_OR_AND;          // PlusTimesSRing<bool,bool>
    const int ak = akind_of(A, dtype);
    if (ak < 0) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_spgemm: A values of dtype %d with a product of dtype %d", A->val_dtype, dtype);
    if (B->val_dtype != dtype && B->val_dtype != CB_PATTERN) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_spgemm: B values of dtype %d, product dtype %d (convert B first)", B->val_dtype, dtype);
    switch (dtype) {
        case CB_F32: return expand_typed<float>(ctx, A, B, boff, semiring, ak, acc);
        case CB_F64: return expand_typed<double>(ctx, A, B, boff, semiring, ak, acc);
        case CB_I32: return expand_typed<int32_t>(ctx, A, B, boff, semiring, ak, acc);
        case CB_I64: return expand_typed<int64_t>(ctx, A, B, boff, semiring, ak, acc);
        case CB_U8: return expand_typed<uint8_t>(ctx, A, B, boff, semiring, ak, acc);
    }
    return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_spgemm: dtype %d", dtype);
}

int cb_spgemm_finish(cb_ctx* ctx, cb_spgemm_acc* acc, int semiring, int dtype, int64_t m, int64_t k, cb_coo** out) {
    if (semiring == CB_PLUS_TIMES && dtype == CB_U8) semiring = CB_OR_AND;
    cb_coo* c = new cb_coo();
    c->ctx = ctx; c->dtype = dtype; c->m = m; c->k = k;
    int st = CB_ERR_UNSUPPORTED;
    switch (dtype) {
        case CB_F32: st = finish_typed<float>(ctx, acc, semiring, m, k, c); break;
        case CB_F64: st = finish_typed<double>(ctx, acc, semiring, m, k, c); break;
        case CB_I32: st = finish_typed<int32_t>(ctx, acc, semiring, m, k, c); break;
        case CB_I64: st = finish_typed<int64_t>(ctx, acc, semiring, m, k, c); break;
        case CB_U8: st = finish_typed<uint8_t>(ctx, acc, semiring, m, k, c); break;
    }
    cb_spgemm_acc_release(acc);
    if (st != CB_OK) { cb_coo_free(c); return st; }
    *out = c;
    return CB_OK;
}

namespace {
__global__ void __launch_bounds__(256)
split_keys_kernel(const uint64_t* __restrict__ keys, int64_t n, int64_t* __restrict__ rows, int64_t* __restrict__ cols) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        rows[p] = (int64_t)(keys[p] & 0xffffffffull);
        cols[p] = (int64_t)(keys[p] >> 32);
    }
}
}  // namespace

extern "C" {

int cb_spgemm_local(cb_ctx* ctx, const cb_tile* A, const cb_tile* B, int semiring, int dtype, cb_coo** C) {
    if (!ctx || !A || !B || !C) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_spgemm_local: null argument");
    if (A->n != B->m) return cb_fail(ctx, CB_ERR_DIMMISMATCH, "cb_spgemm_local: A is %lld x %lld, B is %lld x %lld", (long long)A->m, (long long)A->n, (long long)B->m, (long long)B->n);
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cb_spgemm_acc acc;
    int st = cb_spgemm_expand(ctx, A, B, 0, semiring, dtype, &acc);
    if (st != CB_OK) { cb_spgemm_acc_release(&acc); return st; }
    return cb_spgemm_finish(ctx, &acc, semiring, dtype, A->m, B->n, C);
}

int cb_coo_info(const cb_coo* c, int64_t* nnz, int64_t* m, int64_t* k, int* dtype) {
    if (!c) return cb_fail(nullptr, CB_ERR_INVALIDPARAMS, "cb_coo_info: null argument");
    if (nnz) *nnz = c->nnz;
    if (m) *m = c->m;
    if (k) *k = c->k;
    if (dtype) *dtype = c->dtype;
    return CB_OK;
}

int cb_coo_download(cb_coo* c, int64_t* rows, int64_t* cols, void* vals) {
    if (!c) return cb_fail(nullptr, CB_ERR_INVALIDPARAMS, "cb_coo_download: null argument");
    cb_ctx* ctx = c->ctx;
    if (c->nnz == 0) return CB_OK;
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->compute;
    if (rows && cols) {
        cb_scratch sc;
        int64_t *dr = nullptr, *dc = nullptr;
        CB_CUDA(ctx, sc.alloc(&dr, (size_t)c->nnz));
        CB_CUDA(ctx, sc.alloc(&dc, (size_t)c->nnz));
        split_keys_kernel<<<grid_for(c->nnz, ctx->sm_count), 256, 0, st>>>(c->keys, c->nnz, dr, dc);
        CB_LAUNCHED(ctx);
        CB_CUDA(ctx, cudaMemcpyAsync(rows, dr, (size_t)c->nnz * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaMemcpyAsync(cols, dc, (size_t)c->nnz * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    if (vals) {
        CB_CUDA(ctx, cudaMemcpyAsync(vals, c->vals, (size_t)c->nnz * cb_dtype_size(c->dtype), cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return CB_OK;
}

int cb_coo_free(cb_coo* c) {
    if (!c) return CB_OK;
    cudaFree(c->keys);
    cudaFree(c->vals);
    delete c;
    return CB_OK;
}

}  // extern "C"
