// cb_mmio.cu - Matrix Market text parsed on the device.
//
// Replaces the parsing half of SpParMat::ParallelReadMM (reference include/CombBLAS/SpParMat.cpp:4010-4095: every process reads its
// byte range of the file, cuts it into lines and lets SpParHelper's line handler sscanf "row col value" out of each,
// SpHelper.h:75-91 adding the transpose of every off-diagonal entry of a symmetric file) for the ranks' shares: the caller reads
// the bytes of the lines that start inside its range, this file turns them into triples on the GPU (one thread per line) and
// hands them to the routing / merging of cb_summa.cu (cb_ingest_device_coo), so that between the file system and the finished
// tile the entries never exist as host arrays.
//
// Numbers (cb_mmparse.cuh): up to 19 significant decimal digits and any exponent with a normal double as the result are converted
// with the Eisel-Lemire algorithm - the correctly rounded value, bit for bit what strtod / operator>> / sscanf give (checked
// against strtod on the CPU, tests/emul/mmparse_host.cpp).  Anything else (longer mantissas, subnormal or overflowing values,
// inf / nan, hexadecimal) makes the call return CB_ERR_UNSUPPORTED on EVERY rank before anything is exchanged, and the host
// layer parses that file on the CPU.
#include "cb_common.cuh"
#include "cb_mmparse.cuh"
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

namespace {

inline int grid_for(int64_t n, int sm) {
    int64_t b = (n + 255) / 256, cap = (int64_t)sm * 32;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

__global__ void mm_mark_lines_kernel(const char* __restrict__ text, int64_t n, uint8_t* __restrict__ is_start) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) is_start[i] = (i == 0 || text[i - 1] == '\n') ? 1 : 0;
}

template <typename T> __device__ __forceinline__ T mm_cast(double v) { return (T)v; }
template <> __device__ __forceinline__ uint8_t mm_cast<uint8_t>(double v) { return v != 0.0 ? 1 : 0; }   // C++ (bool)double

// one thread per line: up to two triples (the entry and, for a symmetric file, its transpose); row -1 marks an unused slot
template <typename T>
__global__ void mm_parse_kernel(const char* __restrict__ text, int64_t nbytes, const int64_t* __restrict__ line_start, int64_t nlines, int flags,
                                int64_t* __restrict__ rows, int64_t* __restrict__ cols, T* __restrict__ vals, unsigned int* __restrict__ hard) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < nlines; l += stride) {
        const char* p = text + line_start[l];
        const char* e = text + (l + 1 < nlines ? line_start[l + 1] : nbytes);
        while (e > p && (e[-1] == '\n' || e[-1] == '\r')) --e;
        rows[2 * l] = -1; rows[2 * l + 1] = -1;
        long long ii = 0, jj = 0;
        double vv = 1.0;
        int st = mmparse::parse_int(p, e, &p, &ii);
        if (st == 0) st = mmparse::parse_int(p, e, &p, &jj);
        if (st == 0 && !(flags & 2)) {
            uint64_t bits = 0;
            st = mmparse::parse_double_bits(p, e, &bits) ? 2 : 0;       // an entry without its value column is left to the host too
            vv = __longlong_as_double((long long)bits);
        }
        if (st == 2) { atomicAdd(hard, 1u); continue; }
        if (st == 1) continue;
        if (flags & 1) { --ii; --jj; }
        rows[2 * l] = ii; cols[2 * l] = jj;
        if (vals) vals[2 * l] = mm_cast<T>(vv);
        if ((flags & 4) && ii != jj) {
            rows[2 * l + 1] = jj; cols[2 * l + 1] = ii;
            if (vals) vals[2 * l + 1] = mm_cast<T>(vv);
        }
    }
}


__global__ void mm_flags_kernel(const int64_t* __restrict__ rows, int64_t n, uint8_t* __restrict__ used) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) used[i] = rows[i] != -1;
}

template <typename T>
int compact(cb_ctx* ctx, cb_scratch& sc, const T* in, const uint8_t* used, int64_t n, T* out, int64_t* d_count) {
    size_t b = 0;
    CB_CUDA(ctx, cub::DeviceSelect::Flagged(nullptr, b, in, used, out, d_count, (int)n, ctx->compute));
    char* tmp = nullptr;
    CB_CUDA(ctx, sc.alloc(&tmp, b));
    CB_CUDA(ctx, cub::DeviceSelect::Flagged(tmp, b, in, used, out, d_count, (int)n, ctx->compute));
    return CB_OK;
}

}  // namespace

extern "C" int cb_tile_from_mm_text(cb_ctx* ctx, int64_t gm, int64_t gn, const char* text, int64_t nbytes, int flags, int val_dtype, int dup_op,
                                    cb_tile** out) {
    if (!ctx || !out || nbytes < 0 || (nbytes > 0 && !text)) return cb_fail(ctx, CB_ERR_INVALIDPARAMS, "cb_tile_from_mm_text: null argument");
    *out = nullptr;
    const size_t vs = val_dtype == CB_PATTERN ? 0 : cb_dtype_size(val_dtype);
    if (val_dtype != CB_PATTERN && !vs) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_tile_from_mm_text: value dtype %d", val_dtype);
    // a share too large for the 32-bit positions used here is not an error of this rank alone: like a number the parser refuses, it
    // sends EVERY rank to the host parser (the decision is agreed on below, before anybody enters the exchange)
    const bool too_large = nbytes >= (int64_t(1) << 30);
    CB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->compute;
    const int sm = ctx->sm_count;
    cb_scratch sc;
    int64_t nlines = 0, nz = 0;
    unsigned int hard = 0;
    int64_t *d_rows = nullptr, *d_cols = nullptr;
    char* d_vals = nullptr;
    if (nbytes > 0 && !too_large) {
        char* d_text = nullptr;
        uint8_t* d_flag = nullptr;
        int64_t *d_ls = nullptr, *d_count = nullptr;
        unsigned int* d_hard = nullptr;
        CB_CUDA(ctx, sc.alloc(&d_text, (size_t)nbytes)); CB_CUDA(ctx, sc.alloc(&d_flag, (size_t)nbytes));
        CB_CUDA(ctx, sc.alloc(&d_count, 1)); CB_CUDA(ctx, sc.alloc(&d_hard, 1));
        CB_CUDA(ctx, cudaMemcpyAsync(d_text, text, (size_t)nbytes, cudaMemcpyHostToDevice, st));
        CB_CUDA(ctx, cudaMemsetAsync(d_hard, 0, sizeof(unsigned int), st));
        mm_mark_lines_kernel<<<grid_for(nbytes, sm), 256, 0, st>>>(d_text, nbytes, d_flag);
        CB_LAUNCHED(ctx);
        {   // positions of the line starts: the flagged byte offsets out of 0..nbytes-1
            size_t b = 0;
            thrust::counting_iterator<int64_t> it(0);
            CB_CUDA(ctx, sc.alloc(&d_ls, (size_t)nbytes));
            CB_CUDA(ctx, cub::DeviceSelect::Flagged(nullptr, b, it, d_flag, d_ls, d_count, (int)nbytes, st));
            char* tmp = nullptr;
            CB_CUDA(ctx, sc.alloc(&tmp, b));
            CB_CUDA(ctx, cub::DeviceSelect::Flagged(tmp, b, it, d_flag, d_ls, d_count, (int)nbytes, st));
            ctx->launches += 1;
        }
        CB_CUDA(ctx, cudaMemcpyAsync(&nlines, d_count, sizeof nlines, cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
        int64_t *d_r2 = nullptr, *d_c2 = nullptr;
        char* d_v2 = nullptr;
        uint8_t* d_used = nullptr;
        const int64_t cap = 2 * nlines;
        CB_CUDA(ctx, sc.alloc(&d_r2, (size_t)cap)); CB_CUDA(ctx, sc.alloc(&d_c2, (size_t)cap)); CB_CUDA(ctx, sc.alloc(&d_used, (size_t)cap));
        CB_CUDA(ctx, sc.alloc(&d_rows, (size_t)cap)); CB_CUDA(ctx, sc.alloc(&d_cols, (size_t)cap));
        if (vs) { CB_CUDA(ctx, sc.alloc(&d_v2, (size_t)cap * vs)); CB_CUDA(ctx, sc.alloc(&d_vals, (size_t)cap * vs)); }
        const int g = grid_for(nlines, sm);
        switch (val_dtype) {
            case CB_F32: mm_parse_kernel<float><<<g, 256, 0, st>>>(d_text, nbytes, d_ls, nlines, flags, d_r2, d_c2, (float*)d_v2, d_hard); break;
            case CB_F64: mm_parse_kernel<double><<<g, 256, 0, st>>>(d_text, nbytes, d_ls, nlines, flags, d_r2, d_c2, (double*)d_v2, d_hard); break;
            case CB_I32: mm_parse_kernel<int32_t><<<g, 256, 0, st>>>(d_text, nbytes, d_ls, nlines, flags, d_r2, d_c2, (int32_t*)d_v2, d_hard); break;
            case CB_I64: mm_parse_kernel<int64_t><<<g, 256, 0, st>>>(d_text, nbytes, d_ls, nlines, flags, d_r2, d_c2, (int64_t*)d_v2, d_hard); break;
            default: mm_parse_kernel<uint8_t><<<g, 256, 0, st>>>(d_text, nbytes, d_ls, nlines, flags, d_r2, d_c2, (uint8_t*)d_v2, d_hard); break;
        }
        CB_LAUNCHED(ctx);
        CB_CUDA(ctx, cudaGetLastError());
        mm_flags_kernel<<<grid_for(cap, sm), 256, 0, st>>>(d_r2, cap, d_used);
        CB_LAUNCHED(ctx);
        CB_TRY(compact<int64_t>(ctx, sc, d_r2, d_used, cap, d_rows, d_count));
        CB_TRY(compact<int64_t>(ctx, sc, d_c2, d_used, cap, d_cols, d_count));
        switch (vs) {
            case 0: break;
            case 1: CB_TRY(compact<uint8_t>(ctx, sc, (const uint8_t*)d_v2, d_used, cap, (uint8_t*)d_vals, d_count)); break;
            case 4: CB_TRY(compact<uint32_t>(ctx, sc, (const uint32_t*)d_v2, d_used, cap, (uint32_t*)d_vals, d_count)); break;
            default: CB_TRY(compact<uint64_t>(ctx, sc, (const uint64_t*)d_v2, d_used, cap, (uint64_t*)d_vals, d_count)); break;
        }
        ctx->launches += 2 + (vs ? 1 : 0);
        CB_CUDA(ctx, cudaMemcpyAsync(&nz, d_count, sizeof nz, cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaMemcpyAsync(&hard, d_hard, sizeof hard, cudaMemcpyDeviceToHost, st));
        CB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    // every rank must know whether any share held a number the device parser cannot reproduce before anyone enters the exchange
    int64_t anyhard = (int64_t)hard + (too_large ? 1 : 0);
    CB_TRY(cb_comm_allreduce_i64(ctx, 0, 1, &anyhard, 1));
    if (anyhard) return cb_fail(ctx, CB_ERR_UNSUPPORTED, "cb_tile_from_mm_text: the file holds numbers outside the device parser's exact range, or a share of 1 GiB or more (parse it on the host)");
    return cb_ingest_device_coo(ctx, gm, gn, nz, d_rows, d_cols, d_vals, val_dtype, dup_op, out);
}
