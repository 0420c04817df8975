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
static Warp* g_warp = nullptr;                 // one warp runs at a time
static thread_local int t_lane = 0;
inline void rendezvous() { pthread_barrier_wait(&g_warp->bar); }
}  // namespace emul

template <typename T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    emul::g_warp->slot[emul::t_lane] = raw;
    emul::rendezvous();
    const int from = (emul::t_lane & ~(width - 1)) + (src & (width - 1));
    const uint64_t got = emul::g_warp->slot[from];
    if (emul::t_lane == 0) ++emul::g_warp->syncs;
    emul::rendezvous();
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}
inline unsigned __ballot_sync(unsigned, int p) {
    emul::g_warp->pred[emul::t_lane] = p ? 1u : 0u;
    emul::rendezvous();
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= emul::g_warp->pred[i] << i;
    emul::rendezvous();
    return m;
}
inline int __all_sync(unsigned mask, int p) { return __ballot_sync(mask, p) == 0xffffffffu; }
inline int __any_sync(unsigned mask, int p) { return __ballot_sync(mask, p) != 0; }
inline int __reduce_max_sync(unsigned, int v) {
    emul::g_warp->slot[emul::t_lane] = (uint64_t)(int64_t)v;
    emul::rendezvous();
    int m = (int)(int64_t)emul::g_warp->slot[0];
    for (int i = 1; i < 32; ++i) m = std::max(m, (int)(int64_t)emul::g_warp->slot[i]);
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

namespace emul {
// run `kernel()` for every thread of a grid; warps one after another, the 32 lanes of a warp as threads in lock step
inline long long launch(dim3 grid, dim3 block, const std::function<void()>& kernel) {
    Warp w;
    pthread_barrier_init(&w.bar, nullptr, 32);
    g_warp = &w;
    struct Job { dim3 grid, block; uint3 bidx; int warp; const std::function<void()>* fn; int lane; };
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx)
            for (unsigned wi = 0; wi < block.x / 32; ++wi) {
                pthread_t th[32];
                Job jobs[32];
                for (int l = 0; l < 32; ++l) {
                    jobs[l] = Job{grid, block, uint3{bx, by, 0}, (int)wi, &kernel, l};
                    pthread_create(&th[l], nullptr, [](void* p) -> void* {
                        Job* j = (Job*)p;
                        gridDim = j->grid; blockDim = j->block; blockIdx = j->bidx;
                        threadIdx = uint3{(unsigned)(j->warp * 32 + j->lane), 0, 0};
                        t_lane = j->lane;
                        (*j->fn)();
                        return nullptr;
                    }, &jobs[l]);
                }
                for (int l = 0; l < 32; ++l) pthread_join(th[l], nullptr);
            }
    pthread_barrier_destroy(&w.bar);
    g_warp = nullptr;
    return w.syncs;
}
}  // namespace emul
