// Ceiling probe for K2: how fast can this GPU gather rows of a table, with no arithmetic and no output?
//
// K2's work is "for every nonzero pull one k*sizeof(T)-byte row of X into an SM".  This program measures the rate of exactly
// that access pattern in isolation - random row indices read as a coalesced stream (as K2 reads A's column indices), rows
// gathered with 128-bit loads by virtual warps of W/16 lanes, U loads in flight per lane - for tables that sit in L2, that
// sit in DRAM, and for rows staged in shared memory (the ceiling of a hub row served on the SM).  The numbers are the
// practical denominators for K2's gather rate; the spec's roofline stays B_alg / HBM peak.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_probe tools/gather_probe.cu
//   tools/gather_probe            -> one JSON line per (source, row bytes, table bytes)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((z ^ (z >> 31)) >> 16);
}
__global__ void fill_idx(int32_t* idx, int64_t n, uint32_t mask) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) idx[i] = (int32_t)(mix((uint64_t)i) & mask);
}
__global__ void fill_tab(uint4* t, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) t[i] = make_uint4((uint32_t)i, 1u, 2u, 3u);
}

// global-memory gather: virtual warp of VW lanes walks `per` indices, U row loads in flight per lane
template <int VW, int U, int R>
__global__ void __launch_bounds__(256) gather_global(const int32_t* __restrict__ idx, int64_t nidx, int per, const char* __restrict__ tab, int row_bytes, uint4* __restrict__ sink) {
    const int lane = threadIdx.x & 31, vl = lane & (VW - 1);
    const int64_t vw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / VW;
    const int64_t s = vw * per;
    uint4 acc = make_uint4(0, 0, 0, 0);
    if (s < nidx) {
        for (int base = 0; base < per; base += VW) {
            const int32_t mine = __ldcs(idx + s + base + vl);
#pragma unroll
            for (int j0 = 0; j0 < VW; j0 += U) {
                uint4 x[U * R];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t c = (uint32_t)__shfl_sync(0xffffffffu, mine, j0 + u, VW);
#pragma unroll
                    for (int r = 0; r < R; ++r) x[u * R + r] = __ldg(reinterpret_cast<const uint4*>(tab + (uint64_t)c * (uint32_t)row_bytes + (vl + r * VW) * 16));
                }
#pragma unroll
                for (int u = 0; u < U * R; ++u) { acc.x += x[u].x; acc.y ^= x[u].y; acc.z += x[u].z; acc.w ^= x[u].w; }
            }
        }
    }
    if (acc.x == 0x12345678u && acc.y == 0x9abcdef0u) sink[0] = acc;      // never true; keeps the loads alive
}

// shared-memory gather: the table (as many rows as fit) is staged once per CTA, then gathered from with LDS.128
template <int VW, int U>
__global__ void __launch_bounds__(1024, 1) gather_shared(const int32_t* __restrict__ idx, int64_t nidx, int per, const char* __restrict__ tab, int row_bytes, int rows, uint4* __restrict__ sink) {
    extern __shared__ __align__(16) char sm[];
    for (int i = threadIdx.x; i < rows * (row_bytes / 16); i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = reinterpret_cast<const uint4*>(tab)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, vl = lane & (VW - 1);
    const int64_t vw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / VW;
    const int64_t s = vw * per;
    uint4 acc = make_uint4(0, 0, 0, 0);
    if (s < nidx) {
        for (int base = 0; base < per; base += VW) {
            const int32_t mine = __ldcs(idx + s + base + vl);
#pragma unroll
            for (int j0 = 0; j0 < VW; j0 += U) {
                uint4 x[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t c = (uint32_t)__shfl_sync(0xffffffffu, mine, j0 + u, VW) % (uint32_t)rows;
                    x[u] = *reinterpret_cast<const uint4*>(sm + c * (uint32_t)row_bytes + vl * 16);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) { acc.x += x[u].x; acc.y ^= x[u].y; acc.z += x[u].z; acc.w ^= x[u].w; }
            }
        }
    }
    if (acc.x == 0x12345678u && acc.y == 0x9abcdef0u) sink[0] = acc;
}

template <int VW, int R = 1>
static void run(const char* what, int row_bytes, int64_t rows, const int32_t* idx, int64_t nidx, const char* tab, uint4* sink, bool shared, int64_t pow2rows) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int per = 512;
    const int64_t nvw = nidx / per;
    float best = 1e30f;
    for (int it = 0; it < 6; ++it) {
        CK(cudaEventRecord(a));
        if (shared) {
            const int bt = 1024;
            const int64_t vw_per_cta = bt / VW;
            CK(cudaFuncSetAttribute(gather_shared<VW, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(rows * row_bytes)));
            gather_shared<VW, 4><<<(unsigned)((nvw + vw_per_cta - 1) / vw_per_cta), bt, (size_t)(rows * row_bytes)>>>(idx, nidx, per, tab, row_bytes, (int)rows, sink);
        } else {
            const int64_t vw_per_cta = 256 / VW;
            gather_global<VW, (R == 2 ? 4 : 8), R><<<(unsigned)((nvw + vw_per_cta - 1) / vw_per_cta), 256>>>(idx, nidx, per, tab, row_bytes, sink);
        }
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (it > 0 && ms < best) best = ms;
    }
    const double bytes = (double)nvw * per * row_bytes;
    printf("{\"source\": \"%s\", \"row_bytes\": %d, \"table_mb\": %.1f, \"gathers\": %lld, \"ms\": %.3f, \"gather_tbs\": %.2f, \"grows_per_s\": %.2f}\n", what, row_bytes,
           (double)(shared ? rows : pow2rows) * row_bytes / 1e6, (long long)(nvw * per), best, bytes / best / 1e9, (double)nvw * per / best / 1e6);
    fflush(stdout);
}

int main() {
    const int64_t nidx = 1LL << 27;                      // 134 M gathers per launch
    const size_t tab_bytes = 8ull << 30;                  // 8 GiB table
    int32_t* idx; char* tab; uint4* sink;
    CK(cudaMalloc(&idx, nidx * 4)); CK(cudaMalloc(&tab, tab_bytes)); CK(cudaMalloc(&sink, 16));
    fill_tab<<<148 * 8, 256>>>((uint4*)tab, (int64_t)(tab_bytes / 16));
    CK(cudaDeviceSynchronize());
    const int widths[4] = {128, 256, 512, 1024};
    const size_t tables[4] = {16ull << 20, 64ull << 20, 1ull << 30, 8ull << 30};       // L2, L2 (half), DRAM, DRAM
    for (int wi = 0; wi < 4; ++wi) {
        const int w = widths[wi];
        for (int ti = 0; ti < 4; ++ti) {
            const int64_t rows = (int64_t)(tables[ti] / w);
            fill_idx<<<148 * 8, 256>>>(idx, nidx, (uint32_t)(rows - 1));
            CK(cudaDeviceSynchronize());
            const int64_t n = w >= 1024 ? nidx / 2 : nidx;
            const char* what = ti < 2 ? "global_l2" : "global_dram";
            switch (w) {
                case 128: run<8>(what, w, rows, idx, n, tab, sink, false, rows); break;
                case 256: run<16>(what, w, rows, idx, n, tab, sink, false, rows); break;
                case 512: run<32>(what, w, rows, idx, n, tab, sink, false, rows); break;
                case 1024: run<32, 2>(what, w, rows, idx, n, tab, sink, false, rows); break;           // two vectors per lane, as K2's R=2 layout
            }
        }
        if (w <= 512) {
            const int64_t rows = (200 * 1024) / w;
            fill_idx<<<148 * 8, 256>>>(idx, nidx, 0x7fffffffu);
            CK(cudaDeviceSynchronize());
            switch (w) {
                case 128: run<8>("shared", w, rows, idx, nidx, tab, sink, true, rows); break;
                case 256: run<16>("shared", w, rows, idx, nidx, tab, sink, true, rows); break;
                case 512: run<32>("shared", w, rows, idx, nidx, tab, sink, true, rows); break;
            }
        }
    }
    return 0;
}
