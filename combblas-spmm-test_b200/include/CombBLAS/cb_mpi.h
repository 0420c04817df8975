// cb_mpi.h - the handful of MPI names the CombBLAS surface exposes (CommGrid(MPI_Comm, ...), MPI_COMM_WORLD,
// MPI_Init/Finalize, MPI_Comm_rank/size, MPI_Barrier, MPI_Abort, MPI_Wtime), for a box without MPI.
//
// One process per GPU, started by any launcher that exports RANK / WORLD_SIZE / LOCAL_RANK and
// MASTER_ADDR / MASTER_PORT (e.g. `python -m torch.distributed.run --no-python ./driver`), or a single process.
// The only out-of-band exchange the host layer needs is the 128-byte NCCL unique id; it travels through a
// rendezvous directory on the node's file system.  Everything after that (sizes, sums) uses the NCCL communicators
// inside libcombblas_b200.  If a real MPI is present, compile with -DCB_HAVE_MPI and the real <mpi.h> is used.
//
// Mirrors the role of <mpi.h> in the reference (include/CombBLAS/CommGrid.h:36-47).
#ifndef CB_MPI_H
#define CB_MPI_H

#ifdef CB_HAVE_MPI
#include <mpi.h>
#else
#include <sys/stat.h>
#include <unistd.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

typedef int MPI_Comm;
#define MPI_COMM_WORLD 1
#define MPI_COMM_NULL 0
#define MPI_SUCCESS 0

namespace cb_rt {   // tiny launcher-environment runtime

inline int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}
inline int rank() { return env_int("RANK", 0); }
inline int size() { return env_int("WORLD_SIZE", 1); }
inline int local_rank() { return env_int("LOCAL_RANK", rank()); }

inline std::string rendezvous_dir() {
    const char* d = std::getenv("CB_RENDEZVOUS_DIR");
    if (d) return d;
    const char* port = std::getenv("MASTER_PORT");
    const char* run = std::getenv("TORCHELASTIC_RUN_ID");
    return std::string("/tmp/cb_rdv_") + (port ? port : "0") + "_" + (run ? run : "none") + "_" + std::to_string((long)getppid());
}
// Row / column worlds as distinct handles: 1 = world, 1000 + r = processor row r, 2000 + c = processor column c of the most
// recently built CommGrid (its shape is kept here).  Enough for user code that reduces over GetRowWorld() / GetColWorld().
inline int& grid_cols() { static int pc = 1; return pc; }
inline bool in_comm(int comm, int rank) {
    if (comm >= 2000) return rank % grid_cols() == comm - 2000;
    if (comm >= 1000) return rank / grid_cols() == comm - 1000;
    return true;
}
inline long& seq(int comm) { static long s[3] = {0, 0, 0}; return s[comm >= 2000 ? 2 : comm >= 1000 ? 1 : 0]; }

// every member of `comm` contributes `bytes` bytes; afterwards every member holds all contributions, indexed by WORLD rank
// (the slots of non-members stay zero).  Each communicator counts its own operations, so the processor rows / columns may
// run different numbers of collectives, as they may with a real MPI.
inline void allgather_bytes(const void* mine, size_t bytes, std::vector<char>& all, int comm = 1) {
    const int p = size(), r = rank();
    all.assign(bytes * (size_t)p, 0);
    if (p == 1) { std::memcpy(all.data(), mine, bytes); return; }
    const std::string dir = rendezvous_dir();
    mkdir(dir.c_str(), 0700);
    const long s = seq(comm)++;
    const std::string base = dir + "/c" + std::to_string(comm) + "_op" + std::to_string(s) + ".r";
    {
        const std::string tmp = base + std::to_string(r) + ".tmp", fin = base + std::to_string(r);
        FILE* f = std::fopen(tmp.c_str(), "wb");
        if (!f) { std::perror("cb_mpi rendezvous"); std::exit(1); }
        std::fwrite(mine, 1, bytes, f);
        std::fclose(f);
        std::rename(tmp.c_str(), fin.c_str());
    }
    for (int q = 0; q < p; ++q) {
        if (!in_comm(comm, q)) continue;
        const std::string fin = base + std::to_string(q);
        for (int tries = 0;; ++tries) {
            FILE* f = std::fopen(fin.c_str(), "rb");
            if (f) {
                const size_t got = std::fread(all.data() + bytes * (size_t)q, 1, bytes, f);
                std::fclose(f);
                if (got == bytes) break;
            }
            if (tries > 600000) { std::fprintf(stderr, "cb_mpi: rank %d timed out waiting for rank %d\n", r, q); std::exit(1); }
            std::this_thread::sleep_for(std::chrono::microseconds(200));
        }
    }
}
inline void barrier() { char c = 0; std::vector<char> all; allgather_bytes(&c, 1, all); }
inline void bcast_bytes(void* buf, size_t bytes, int root) {
    std::vector<char> all;
    allgather_bytes(buf, bytes, all);
    std::memcpy(buf, all.data() + bytes * (size_t)root, bytes);
}

}  // namespace cb_rt

typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_INT 104
#define MPI_LONG_LONG 108
#define MPI_DOUBLE 208
#define MPI_FLOAT 204
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
namespace cb_rt {
template <class T>
inline void reduce_typed(const std::vector<char>& all, size_t bytes, int count, int op, MPI_Comm comm, T* out) {
    bool first = true;
    for (int q = 0; q < size(); ++q) {
        if (!in_comm(comm, q)) continue;
        const T* v = reinterpret_cast<const T*>(all.data() + bytes * (size_t)q);
        for (int i = 0; i < count; ++i) out[i] = first ? v[i] : op == MPI_SUM ? (T)(out[i] + v[i]) : op == MPI_MAX ? (out[i] < v[i] ? v[i] : out[i]) : (v[i] < out[i] ? v[i] : out[i]);
        first = false;
    }
}
}  // namespace cb_rt
inline int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm) {
    const size_t bytes = (size_t)count * (size_t)(dt % 100);
    std::vector<char> all;
    cb_rt::allgather_bytes(sendbuf, bytes, all, comm);
    switch (dt) {
        case MPI_INT: cb_rt::reduce_typed<int>(all, bytes, count, op, comm, (int*)recvbuf); break;
        case MPI_LONG_LONG: cb_rt::reduce_typed<long long>(all, bytes, count, op, comm, (long long*)recvbuf); break;
        case MPI_DOUBLE: cb_rt::reduce_typed<double>(all, bytes, count, op, comm, (double*)recvbuf); break;
        case MPI_FLOAT: cb_rt::reduce_typed<float>(all, bytes, count, op, comm, (float*)recvbuf); break;
        default: std::fprintf(stderr, "cb_mpi: MPI_Allreduce datatype %d\n", dt); std::_Exit(1);
    }
    return MPI_SUCCESS;
}
inline int MPI_Init(int*, char***) { return MPI_SUCCESS; }
inline int MPI_Finalize() { if (cb_rt::size() > 1) cb_rt::barrier(); return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm c, int* r) {
    const int w = cb_rt::rank(), pc = cb_rt::grid_cols();
    *r = c >= 2000 ? w / pc : c >= 1000 ? w % pc : w;            // rank in a column world = processor row, in a row world = column
    return MPI_SUCCESS;
}
inline int MPI_Comm_size(MPI_Comm c, int* s) {
    const int p = cb_rt::size(), pc = cb_rt::grid_cols();
    *s = c >= 2000 ? p / pc : c >= 1000 ? pc : p;
    return MPI_SUCCESS;
}
inline int MPI_Barrier(MPI_Comm) { cb_rt::barrier(); return MPI_SUCCESS; }
inline int MPI_Abort(MPI_Comm, int code) { std::fprintf(stderr, "MPI_Abort(%d)\n", code); std::fflush(stderr); std::_Exit(code & 0xff ? code & 0xff : 1); }
inline int MPI_Pcontrol(int, ...) { return MPI_SUCCESS; }        // profiling hook of ReleaseTests/MultTiming.cpp:67
inline double MPI_Wtime() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#endif  // CB_HAVE_MPI

#endif
