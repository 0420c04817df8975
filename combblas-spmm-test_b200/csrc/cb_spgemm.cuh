// Internal declarations of the device sparse x sparse product (cb_spgemm.cu) shared with the SUMMA driver (cb_summa.cu).
#pragma once
#include "cb_common.cuh"

// the product as sorted, merged triples on the device: keys = (column << 32) | row, column-major like the tuple order
// SpDCCols' constructor wants (SpDCCols.cpp:186-195)
struct cb_coo {
    cb_ctx* ctx = nullptr;
    int64_t nnz = 0, m = 0, k = 0;
    int dtype = CB_F32;
    uint64_t* keys = nullptr;
    void* vals = nullptr;
};

// partial products of the stages of one multiply, one device buffer pair per stage
struct cb_spgemm_acc {
    struct Piece { uint64_t* keys = nullptr; void* vals = nullptr; int64_t count = 0; };
    std::vector<Piece> pieces;
};

int cb_spgemm_expand(cb_ctx* ctx, const cb_tile* A, const cb_tile* B, int64_t boff, int semiring, int dtype, cb_spgemm_acc* acc);
int cb_spgemm_finish(cb_ctx* ctx, cb_spgemm_acc* acc, int semiring, int dtype, int64_t m, int64_t k, cb_coo** out);
void cb_spgemm_acc_release(cb_spgemm_acc* acc);
extern "C" int cb_coo_free(cb_coo* c);
