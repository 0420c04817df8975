// tools/pcie_probe.cu - what the host-panel path (cb_spmm_summa_host) can expect from PCIe: pinned-memory copies of an
// n x 128 fp32 panel (512-byte rows) as one contiguous block and as column slabs of 64 / 128 / 256 bytes per row
// (cudaMemcpy2DAsync), each direction alone and both directions at once, plus kernel-driven (zero-copy) slab copies.
// One JSON line per measurement.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pcie_probe pcie_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

// slab copy by the SMs: every warp moves whole 16-byte vectors of `wbytes`-wide rows between a pitched host panel and a compact
// device slab (dir 0: host -> device, 1: device -> host)
__global__ void slab_copy_kernel(char* __restrict__ host, size_t pitch, char* __restrict__ dev, size_t wbytes, int64_t rows, int dir) {
    const int vec_per_row = (int)(wbytes / 16);
    const int64_t total = rows * vec_per_row;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += stride) {
        const int64_t r = v / vec_per_row;
        const int c = (int)(v % vec_per_row);
        uint4* h = (uint4*)(host + (size_t)r * pitch) + c;
        uint4* d = (uint4*)(dev + (size_t)r * wbytes) + c;
        if (dir == 0) *d = *h; else *h = *d;
    }
}

int main(int argc, char** argv) {
    const int64_t rows = argc > 1 ? atoll(argv[1]) : (1ll << 23);
    const size_t pitch = 512;
    const size_t bytes = (size_t)rows * pitch;
    char *hx, *hy, *dx, *dy;
    CK(cudaMallocHost(&hx, bytes)); CK(cudaMallocHost(&hy, bytes));
    CK(cudaMalloc(&dx, bytes)); CK(cudaMalloc(&dy, bytes));
    for (size_t i = 0; i < bytes; i += 4096) { hx[i] = 1; hy[i] = 2; }
    cudaStream_t up, down;
    CK(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, f0, f1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&f0)); CK(cudaEventCreate(&f1));
    const size_t widths[] = {512, 256, 128, 64};
    for (int mode = 0; mode < 2; ++mode)                       // 0 = copy engines, 1 = SM kernel over mapped pinned memory
        for (size_t w : widths)
            for (int dirs = 0; dirs < 3; ++dirs) {             // 0 up only, 1 down only, 2 both at once
                if (mode == 1 && w == 512 && false) continue;
                float best_up = 1e30f, best_down = 1e30f;
                for (int rep = 0; rep < 3; ++rep) {
                    CK(cudaDeviceSynchronize());
                    if (dirs != 1) {
                        CK(cudaEventRecord(e0, up));
                        for (size_t c = 0; c < pitch; c += w) {
                            if (mode == 0) CK(cudaMemcpy2DAsync(dx + c * rows, w, hx + c, pitch, w, (size_t)rows, cudaMemcpyHostToDevice, up));
                            else slab_copy_kernel<<<148 * 8, 256, 0, up>>>(hx + c, pitch, dx + c * rows, w, rows, 0);
                        }
                        CK(cudaEventRecord(e1, up));
                    }
                    if (dirs != 0) {
                        CK(cudaEventRecord(f0, down));
                        for (size_t c = 0; c < pitch; c += w) {
                            if (mode == 0) CK(cudaMemcpy2DAsync(hy + c, pitch, dy + c * rows, w, w, (size_t)rows, cudaMemcpyDeviceToHost, down));
                            else slab_copy_kernel<<<148 * 8, 256, 0, down>>>(hy + c, pitch, dy + c * rows, w, rows, 1);
                        }
                        CK(cudaEventRecord(f1, down));
                    }
                    CK(cudaDeviceSynchronize());
                    float t;
                    if (dirs != 1) { CK(cudaEventElapsedTime(&t, e0, e1)); if (t < best_up) best_up = t; }
                    if (dirs != 0) { CK(cudaEventElapsedTime(&t, f0, f1)); if (t < best_down) best_down = t; }
                }
                printf("{\"probe\": \"pcie\", \"engine\": \"%s\", \"row_bytes\": %zu, \"panel_gb\": %.2f, \"directions\": \"%s\"", mode ? "sm_zero_copy" : "copy_engine", w,
                       bytes / 1e9, dirs == 0 ? "h2d" : dirs == 1 ? "d2h" : "both");
                if (dirs != 1) printf(", \"h2d_gbs\": %.1f", bytes / 1e6 / best_up);
                if (dirs != 0) printf(", \"d2h_gbs\": %.1f", bytes / 1e6 / best_down);
                printf("}\n");
                fflush(stdout);
            }
    return 0;
}
