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
#include <algorithm>
#include <chrono>
#include <cstdint>
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
    // the launcher's pid AND its start time (field 22 of /proc/<pid>/stat): a later launcher that reuses the pid gets another directory
    std::string started = "0";
    {
        const std::string st = "/proc/" + std::to_string((long)getppid()) + "/stat";
        if (FILE* f = std::fopen(st.c_str(), "r")) {
            char buf[1024] = {0};
            const size_t n = std::fread(buf, 1, sizeof buf - 1, f);
            std::fclose(f);
            const char* q = n ? std::strrchr(buf, ')') : nullptr;       // the command name may contain spaces
            int field = 2;
            for (const char* c = q ? q + 1 : nullptr; c && *c; ++c)
                if (*c == ' ' && ++field == 22) { started = std::to_string(std::atoll(c + 1)); break; }
        }
    }
    return std::string("/tmp/cb_rdv_") + (port ? port : "0") + "_" + (run ? run : "none") + "_" + std::to_string((long)getppid()) + "_" + started;
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

inline uint64_t fresh_token() {
    uint64_t t = (uint64_t)std::chrono::steady_clock::now().time_since_epoch().count() ^ ((uint64_t)getpid() << 32);
    if (FILE* f = std::fopen("/dev/urandom", "rb")) { uint64_t r = 0; if (std::fread(&r, sizeof r, 1, f) == 1) t ^= r; std::fclose(f); }
    return t ? t : 1;
}
inline void write_file(const std::string& fin, const void* data, size_t bytes) {       // atomic: readers see all of it or nothing
    const std::string tmp = fin + ".tmp" + std::to_string((long)getpid());
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) { std::perror("cb_mpi rendezvous"); std::exit(1); }
    std::fwrite(data, 1, bytes, f);
    std::fclose(f);
    std::rename(tmp.c_str(), fin.c_str());
}
inline bool read_file(const std::string& fin, void* data, size_t bytes) {
    FILE* f = std::fopen(fin.c_str(), "rb");
    if (!f) return false;
    const size_t got = std::fread(data, 1, bytes, f);
    std::fclose(f);
    return got == bytes;
}
inline void nap(long tries, int who) {
    if (tries > 600000) { std::fprintf(stderr, "cb_mpi: rank %d timed out in the rendezvous waiting for rank %d\n", rank(), who); std::exit(1); }
    std::this_thread::sleep_for(std::chrono::microseconds(200));
}

// Session nonce.  The directory may hold files of an earlier run that died (same launcher, same port), so before the first
// collective the ranks agree on a fresh 64-bit nonce that every later file carries in its header; files with another nonce are
// treated as not written yet.  Every rank first removes whatever files carry ITS OWN rank number, then publishes a random
// token; rank 0 publishes its nonce with the tokens it saw, generation by generation, until every rank has acknowledged a
// generation that lists the token it really wrote (an acknowledgement echoes rank 0's nonce, so a stale one cannot match).
struct SessionMsg { uint64_t nonce; uint64_t gen; uint64_t final; uint64_t tokens[256]; };
struct AckMsg { uint64_t nonce; uint64_t gen; uint64_t ok; };
inline uint64_t session_nonce() {
    static uint64_t nonce = 0;
    if (nonce) return nonce;
    const int p = size(), r = rank();
    if (p > 256) { std::fprintf(stderr, "cb_mpi: the file rendezvous supports up to 256 processes\n"); std::exit(1); }
    const std::string dir = rendezvous_dir();
    mkdir(dir.c_str(), 0700);
    const std::string me = std::to_string(r);
    std::remove((dir + "/hello." + me).c_str());
    std::remove((dir + "/ack." + me).c_str());
    std::remove((dir + "/done." + me).c_str());
    if (r == 0) std::remove((dir + "/session").c_str());
    const uint64_t mine = fresh_token();
    write_file(dir + "/hello." + me, &mine, sizeof mine);
    if (r == 0) {
        SessionMsg m;
        std::memset(&m, 0, sizeof m);
        m.nonce = fresh_token();
        for (m.gen = 1;; ++m.gen) {
            for (int q = 0; q < p; ++q)
                for (long t = 0; !read_file(dir + "/hello." + std::to_string(q), &m.tokens[q], sizeof(uint64_t)); ++t) nap(t, q);
            write_file(dir + "/session", &m, sizeof m);
            bool all_ok = true;
            for (int q = 1; q < p; ++q) {
                AckMsg a;
                for (long t = 0;; ++t) {
                    if (read_file(dir + "/ack." + std::to_string(q), &a, sizeof a) && a.nonce == m.nonce && a.gen == m.gen) break;
                    nap(t, q);
                }
                all_ok = all_ok && a.ok;
            }
            if (all_ok) break;
        }
        m.final = 1;
        write_file(dir + "/session", &m, sizeof m);
        nonce = m.nonce;
    } else {
        uint64_t seen_nonce = 0, seen_gen = 0, ok_nonce = 0;      // last generation answered; nonce I said "ok" to
        for (long t = 0;; ++t) {
            SessionMsg m;
            if (read_file(dir + "/session", &m, sizeof m)) {
                const bool lists_me = m.tokens[r] == mine;
                if (m.final && lists_me && m.nonce == ok_nonce) { nonce = m.nonce; break; }
                if (!m.final && (m.nonce != seen_nonce || m.gen != seen_gen)) {
                    AckMsg a = {m.nonce, m.gen, lists_me ? 1u : 0u};
                    write_file(dir + "/ack." + me, &a, sizeof a);
                    seen_nonce = m.nonce;
                    seen_gen = m.gen;
                    if (lists_me) ok_nonce = m.nonce;
                }
            }
            nap(t, 0);
        }
    }
    return nonce;
}

// every member of `comm` contributes `bytes` bytes; afterwards every member holds all contributions, indexed by WORLD rank
// (the slots of non-members stay zero).  Each communicator counts its own operations, so the processor rows / columns may
// run different numbers of collectives, as they may with a real MPI.  A rank removes its file of operation s-2 when it
// starts operation s of the same communicator: completing s-1 required every member to have written s-1, which each does
// only after it has read all of s-2.
inline void allgather_bytes(const void* mine, size_t bytes, std::vector<char>& all, int comm = 1) {
    const int p = size(), r = rank();
    all.assign(bytes * (size_t)p, 0);
    if (p == 1) { std::memcpy(all.data(), mine, bytes); return; }
    const uint64_t nonce = session_nonce();
    const std::string dir = rendezvous_dir();
    const long s = seq(comm)++;
    auto name = [&](long op, int q) { return dir + "/c" + std::to_string(comm) + "_op" + std::to_string(op) + ".r" + std::to_string(q); };
    if (s >= 2) std::remove(name(s - 2, r).c_str());
    {
        std::vector<char> buf(sizeof nonce + bytes);
        std::memcpy(buf.data(), &nonce, sizeof nonce);
        if (bytes) std::memcpy(buf.data() + sizeof nonce, mine, bytes);
        write_file(name(s, r), buf.data(), buf.size());
    }
    std::vector<char> buf(sizeof nonce + bytes);
    for (int q = 0; q < p; ++q) {
        if (!in_comm(comm, q)) continue;
        for (long tries = 0;; ++tries) {
            uint64_t got = 0;
            if (read_file(name(s, q), buf.data(), buf.size()) && (std::memcpy(&got, buf.data(), sizeof got), got == nonce)) break;
            nap(tries, q);
        }
        if (bytes) std::memcpy(all.data() + bytes * (size_t)q, buf.data() + sizeof nonce, bytes);
    }
}
// leave nothing behind: after a closing barrier pair every rank removes its own files; the last one out removes the directory
inline void finalize() {
    if (size() == 1) return;
    char c = 0;
    std::vector<char> all;
    allgather_bytes(&c, 1, all);
    allgather_bytes(&c, 1, all);                     // everyone has read the first barrier: all earlier files are dead
    const std::string dir = rendezvous_dir(), me = std::to_string(rank());
    const int comms[3] = {1, 1000 + rank() / grid_cols(), 2000 + rank() % grid_cols()};
    for (int ci = 0; ci < 3; ++ci)
        for (long op = std::max(0L, seq(comms[ci]) - 3); op < seq(comms[ci]) - (ci == 0 ? 1 : 0); ++op)
            std::remove((dir + "/c" + std::to_string(comms[ci]) + "_op" + std::to_string(op) + ".r" + me).c_str());
    std::remove((dir + "/hello." + me).c_str());
    std::remove((dir + "/ack." + me).c_str());
    const uint64_t one = 1;
    write_file(dir + "/done." + me, &one, sizeof one);
    bool all_done = true;
    uint64_t v;
    for (int q = 0; q < size(); ++q) all_done = all_done && read_file(dir + "/done." + std::to_string(q), &v, sizeof v);
    if (all_done) {                                  // the last barrier's files and the markers; then the directory itself
        for (int q = 0; q < size(); ++q) {
            std::remove((dir + "/c1_op" + std::to_string(seq(1) - 1) + ".r" + std::to_string(q)).c_str());
            std::remove((dir + "/done." + std::to_string(q)).c_str());
        }
        std::remove((dir + "/session").c_str());
        rmdir(dir.c_str());
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
inline int MPI_Finalize() { cb_rt::finalize(); return MPI_SUCCESS; }
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
