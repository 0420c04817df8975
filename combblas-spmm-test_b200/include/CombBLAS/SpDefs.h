// Constants, abort codes and small helpers shared by the host layer.
// Mirrors reference include/CombBLAS/SpDefs.h:60-131 (EPSILON :64, abort codes :72-78) and the rank-0 printing of
// SpParHelper::Print (SpParHelper.cpp:836-844).
#ifndef CB_SPDEFS_H
#define CB_SPDEFS_H

#include <cstdint>
#include <cstdio>
#include <iostream>
#include <string>
#include "cb_mpi.h"
#include "combblas_b200.h"

#define EPSILON 0.01

// MPI_Abort codes (same numbers as the reference)
#define GRIDMISMATCH 3001
#define DIMMISMATCH 3002
#define NOTSQUARE 3003
#define NOFILE 3004
#define MATRIXALIAS 3005
#define UNKNOWNMPITYPE 3006
#define INVALIDPARAMS 3007

namespace combblas {

enum Dim { Column, Row };

struct SpParHelper {
    static void Print(const std::string& s) {
        int r = 0;
        MPI_Comm_rank(MPI_COMM_WORLD, &r);
        if (r == 0) std::cerr << s;
    }
};

// A failing C-ABI call ends the program the way the reference ends on its own errors: message on stderr, MPI_Abort.
inline void cb_check(int status, const cb_ctx* ctx, const char* what) {
    if (status == CB_OK) return;
    std::cerr << "COMBBLAS-B200: " << what << " failed with status " << status << " (" << cb_status_string(status) << "): "
              << cb_last_error(ctx) << std::endl;
    MPI_Abort(MPI_COMM_WORLD, status >= 3001 ? status : INVALIDPARAMS);
}

}  // namespace combblas
#endif
