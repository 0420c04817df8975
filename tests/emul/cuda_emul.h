// Lock-step warp emulator for compiling the product's CUDA kernel headers with g++ (TEST INFRASTRUCTURE ONLY).
//
// compute-sanitizer is closed on the GPU pool, so memory safety and warp-convergence of the hot kernel are checked here
// instead: csrc/cb_spmm_kernel.cuh is compiled unmodified for the host, every lane of a warp runs as an OS thread, and the
// *_sync intrinsics are real rendezvous points (a lane that skips one dead-locks the warp -> the test times out).  The
// harness (kernel_emul.cpp) runs under AddressSanitizer / UBSan with exactly-sized heap buffers.
// Nothing in the product includes this file; the product still has no CPU path.
#pragma once
#include <cuda_runtime.h>
#include <pthread.h>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <vector>

#define __launch_bounds__(...)

static thread_local uint3 threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;

namespace emul {
struct Warp {
    pthread_barrier_t bar;
    uint64_t slot[32];
    unsigned pred[32];
    long long syncs = 0;
};
static thread_local Warp* t_warp = nullptr;     // the warp this OS thread is a lane of
static thread_local int t_lane = 0;
inline void rendezvous() { pthread_barrier_wait(&t_warp->bar); }
}  // namespace emul

template <typename T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    emul::t_warp->slot[emul::t_lane] = raw;
    emul::rendezvous();
    const int from = (emul::t_lane & ~(width - 1)) + (src & (width - 1));
    const uint64_t got = emul::t_warp->slot[from];
    if (emul::t_lane == 0) ++emul::t_warp->syncs;
    emul::rendezvous();
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}
inline unsigned __ballot_sync(unsigned, int p) {
    emul::t_warp->pred[emul::t_lane] = p ? 1u : 0u;
    emul::rendezvous();
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= emul::t_warp->pred[i] << i;
    emul::rendezvous();
    return m;
}
inline int __all_sync(unsigned mask, int p) { return __ballot_sync(mask, p) == 0xffffffffu; }
inline int __any_sync(unsigned mask, int p) { return __ballot_sync(mask, p) != 0; }
inline int __reduce_max_sync(unsigned, int v) {
    emul::t_warp->slot[emul::t_lane] = (uint64_t)(int64_t)v;
    emul::rendezvous();
    int m = (int)(int64_t)emul::t_warp->slot[0];
    for (int i = 1; i < 32; ++i) m = std::max(m, (int)(int64_t)emul::t_warp->slot[i]);
    emul::rendezvous();
    return m;
}
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcs(const T* p) { return *p; }
template <typename T> inline void __stcs(T* p, T v) { *p = v; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
using std::max;
using std::min;

// ---- shared memory, __syncthreads, clusters and distributed shared memory (the hub variant of K2) ----
// A "shared-window address" is an offset into the owning CTA's dynamic shared memory buffer.
#define CB_CLUSTER_INTRINSICS
// L2 eviction-priority hints of K2P: no cache on the host, the policy only has to travel
inline uint64_t cb_policy_evict_last() { return 0x1111; }
inline uint64_t cb_policy_evict_first() { return 0x2222; }
inline void cb_prefetch_l2(const void* p) { if (!p) __builtin_trap(); }          // a hint; must at least be a pointer
inline uint4 cb_ldg16_hint(const void* p, uint64_t policy) {
    if (policy != 0x1111 && policy != 0x2222) __builtin_trap();
    return *reinterpret_cast<const uint4*>(p);
}
namespace emul {
struct Cta {
    pthread_barrier_t bar;                    // __syncthreads
    std::vector<char> smem;
};
struct Cluster {
    pthread_barrier_t bar;                    // barrier.cluster
    std::vector<Cta*> ctas;
};
static thread_local Cta* t_cta = nullptr;
static thread_local Cluster* t_cluster = nullptr;
static thread_local unsigned t_ctarank = 0;
}  // namespace emul
inline void __syncthreads() { pthread_barrier_wait(&emul::t_cta->bar); }
inline char* cb_dyn_smem() { return emul::t_cta->smem.data(); }
inline uint32_t cb_smem_u32(const void* p) { return (uint32_t)((const char*)p - emul::t_cta->smem.data()); }
inline uint32_t cb_cluster_ctarank() { return emul::t_ctarank; }
inline uint32_t cb_cluster_nctarank() { return (uint32_t)emul::t_cluster->ctas.size(); }
inline void cb_cluster_sync() { pthread_barrier_wait(&emul::t_cluster->bar); }
inline uint4 cb_ld_cluster16(uint32_t addr, uint32_t cta) {
    uint4 u;
    std::memcpy(&u, emul::t_cluster->ctas.at(cta)->smem.data() + addr, 16);     // .at(): a bad CTA rank is an exception
    return u;
}
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
// cp.async: copies are DEFERRED until a wait_group retires their group, so a kernel that reads a ring slot before waiting
// for it sees the stale fill pattern (0x5a) and fails the comparison, exactly the bug class the hardware would hide at random
namespace emul {
struct AsyncCopy { uint32_t dst; const void* src; };
static thread_local std::vector<std::vector<AsyncCopy>> t_groups;      // committed, oldest first
static thread_local std::vector<AsyncCopy> t_open;
}  // namespace emul
inline void cb_cp_async16(uint32_t smem_addr, const void* gptr) { emul::t_open.push_back(emul::AsyncCopy{smem_addr, gptr}); }
inline void cb_cp_async_commit() { emul::t_groups.push_back(emul::t_open); emul::t_open.clear(); }
template <int N>
inline void cb_cp_async_wait() {
    while ((int)emul::t_groups.size() > N) {
        for (const emul::AsyncCopy& c : emul::t_groups.front()) std::memcpy(&emul::t_cta->smem.at(c.dst + 15) - 15, c.src, 16);   // .at(): bounds
        emul::t_groups.erase(emul::t_groups.begin());
    }
}
inline uint4 cb_lds16(uint32_t smem_addr) {
    uint4 u;
    std::memcpy(&u, &emul::t_cta->smem.at(smem_addr + 15) - 15, 16);
    return u;
}

namespace emul {
// run `kernel()` for every thread of a grid launched in clusters of `cs` CTAs along x with `smem_bytes` of dynamic shared
// memory per CTA: clusters one after another, all threads of a cluster concurrently (warps in lock step on their own
// barrier), so __syncthreads, cluster barriers and reads of another CTA's shared memory behave as on the device.
// Shared memory is an exactly-sized heap buffer: an out-of-bounds slot is an ASan report.
inline void launch_cluster(dim3 grid, dim3 block, unsigned cs, size_t smem_bytes, const std::function<void()>& kernel) {
    struct Job { dim3 grid, block; uint3 bidx; int warp, lane; Warp* w; Cta* cta; Cluster* cl; unsigned ctarank; const std::function<void()>* fn; };
    const unsigned nwarps = block.x / 32;
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx0 = 0; bx0 < grid.x; bx0 += cs) {
            Cluster cl;
            std::vector<Cta> ctas(cs);
            std::vector<Warp> warps((size_t)cs * nwarps);
            pthread_barrier_init(&cl.bar, nullptr, cs * block.x);
            for (unsigned c = 0; c < cs; ++c) {
                pthread_barrier_init(&ctas[c].bar, nullptr, block.x);
                ctas[c].smem.assign(smem_bytes, (char)0x5a);
                cl.ctas.push_back(&ctas[c]);
            }
            for (auto& w : warps) pthread_barrier_init(&w.bar, nullptr, 32);
            std::vector<Job> jobs((size_t)cs * block.x);
            std::vector<pthread_t> th(jobs.size());
            for (unsigned c = 0; c < cs; ++c)
                for (unsigned t = 0; t < block.x; ++t)
                    jobs[(size_t)c * block.x + t] = Job{grid, block, uint3{bx0 + c, by, 0}, (int)(t / 32), (int)(t % 32),
                                                        &warps[(size_t)c * nwarps + t / 32], &ctas[c], &cl, c, &kernel};
            for (size_t i = 0; i < jobs.size(); ++i)
                pthread_create(&th[i], nullptr, [](void* p) -> void* {
                    Job* j = (Job*)p;
                    gridDim = j->grid; blockDim = j->block; blockIdx = j->bidx;
                    threadIdx = uint3{(unsigned)(j->warp * 32 + j->lane), 0, 0};
                    t_lane = j->lane; t_warp = j->w; t_cta = j->cta; t_cluster = j->cl; t_ctarank = j->ctarank;
                    (*j->fn)();
                    return nullptr;
                }, &jobs[i]);
            for (auto& t : th) pthread_join(t, nullptr);
        }
}

// run `kernel()` for every thread of a grid; warps one after another, the 32 lanes of a warp as threads in lock step
inline long long launch(dim3 grid, dim3 block, const std::function<void()>& kernel) {
    Warp w;
    pthread_barrier_init(&w.bar, nullptr, 32);
    struct Job { dim3 grid, block; uint3 bidx; int warp; const std::function<void()>* fn; int lane; Warp* w; };
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx)
            for (unsigned wi = 0; wi < block.x / 32; ++wi) {
                pthread_t th[32];
                Job jobs[32];
                for (int l = 0; l < 32; ++l) {
                    jobs[l] = Job{grid, block, uint3{bx, by, 0}, (int)wi, &kernel, l, &w};
                    pthread_create(&th[l], nullptr, [](void* p) -> void* {
                        Job* j = (Job*)p;
                        gridDim = j->grid; blockDim = j->block; blockIdx = j->bidx;
                        threadIdx = uint3{(unsigned)(j->warp * 32 + j->lane), 0, 0};
                        t_lane = j->lane; t_warp = j->w;
                        (*j->fn)();
                        return nullptr;
                    }, &jobs[l]);
                }
                for (int l = 0; l < 32; ++l) pthread_join(th[l], nullptr);
            }
    pthread_barrier_destroy(&w.bar);
    return w.syncs;
}
}  // namespace emul
