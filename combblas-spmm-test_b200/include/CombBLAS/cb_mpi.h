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
inline long& seq() { static long s = 0; return s; }

// every rank contributes `bytes` bytes; afterwards everyone holds all contributions in rank order
inline void allgather_bytes(const void* mine, size_t bytes, std::vector<char>& all) {
    const int p = size(), r = rank();
    all.assign(bytes * (size_t)p, 0);
    if (p == 1) { std::memcpy(all.data(), mine, bytes); return; }
    const std::string dir = rendezvous_dir();
    mkdir(dir.c_str(), 0700);
    const long s = seq()++;
    const std::string base = dir + "/op" + std::to_string(s) + ".r";
    {
        const std::string tmp = base + std::to_string(r) + ".tmp", fin = base + std::to_string(r);
        FILE* f = std::fopen(tmp.c_str(), "wb");
        if (!f) { std::perror("cb_mpi rendezvous"); std::exit(1); }
        std::fwrite(mine, 1, bytes, f);
        std::fclose(f);
        std::rename(tmp.c_str(), fin.c_str());
    }
    for (int q = 0; q < p; ++q) {
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

inline int MPI_Init(int*, char***) { return MPI_SUCCESS; }
inline int MPI_Finalize() { if (cb_rt::size() > 1) cb_rt::barrier(); return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = cb_rt::rank(); return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm, int* s) { *s = cb_rt::size(); return MPI_SUCCESS; }
inline int MPI_Barrier(MPI_Comm) { cb_rt::barrier(); return MPI_SUCCESS; }
inline int MPI_Abort(MPI_Comm, int code) { std::fprintf(stderr, "MPI_Abort(%d)\n", code); std::fflush(stderr); std::_Exit(code & 0xff ? code & 0xff : 1); }
inline int MPI_Pcontrol(int, ...) { return MPI_SUCCESS; }        // profiling hook of ReleaseTests/MultTiming.cpp:67
inline double MPI_Wtime() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#endif  // CB_HAVE_MPI

#endif
