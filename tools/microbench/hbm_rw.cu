// HBM ceilings by direction on this device: write-only, read-only and copy streams over a buffer much larger than L2.
// conv0 is a write-only stream (3 B read, 128 B written per canvas pixel) and the tail a read-only one (128 B read, 3*s*s
// written), so the copy figure in MEASURED_PEAKS.json is not the bound either of them can reach.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hbm_rw tools/microbench/hbm_rw.cu && /tmp/hbm_rw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) write_kernel(uint4* dst, size_t n, uint32_t v) {
    for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * 256ull) dst[i] = make_uint4(v, v, v, v);
}
__global__ void __launch_bounds__(256) read_kernel(const uint4* src, size_t n, uint32_t* sink) {
    uint32_t acc = 0;
    for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * 256ull) {
        const uint4 x = src[i];
        acc ^= x.x ^ x.y ^ x.z ^ x.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}
__global__ void __launch_bounds__(256) copy_kernel(const uint4* src, uint4* dst, size_t n) {
    for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * 256ull) dst[i] = src[i];
}

template <typename F>
static float time_ms(F&& f, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    const size_t bytes = 1313507840ull;   // four 1080p canvases of 128 B pixels
    const size_t n = bytes / 16;
    uint4 *a, *b;
    uint32_t* sink;
    cudaMalloc(&a, bytes);
    cudaMalloc(&b, bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(a, 1, bytes);
    cudaMemset(b, 2, bytes);
    for (int mult : {4, 8, 16}) {
        const int grid = 148 * mult;
        const float w = time_ms([&] { write_kernel<<<grid, 256>>>(a, n, 7u); }, 10);
        const float r = time_ms([&] { read_kernel<<<grid, 256>>>(a, n, sink); }, 10);
        const float c = time_ms([&] { copy_kernel<<<grid, 256>>>(a, b, n); }, 10);
        printf("{\"grid\": %d, \"write_GBps\": %.0f, \"read_GBps\": %.0f, \"copy_GBps_read_plus_write\": %.0f}\n", grid, bytes / w / 1e6,
               bytes / r / 1e6, 2.0 * bytes / c / 1e6);
    }
    const float ms = time_ms([&] { cudaMemsetAsync(a, 0, bytes); }, 10);
    printf("{\"cudaMemsetAsync_GBps\": %.0f}\n", bytes / ms / 1e6);
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    return 0;
}
