// Write-only HBM bandwidth by store path, with the first conv's access pattern: 148 CTAs, each CTA writes rows of a
// [rows][2129 px][128 B] canvas in 16 KB pieces (128 px), 12 warps per CTA each shipping a 4 KB quarter per piece.
//   mode 0: cp.async.bulk.tensor.3d (box 64 ch x 32 px, SWIZZLE_128B)     -- what the kernels do
//   mode 1: cp.async.bulk.global.shared::cta, one 4 KB linear copy
//   mode 2: st.global.v4 from registers, a warp writes 512 contiguous bytes per instruction, 8 instructions per 4 KB
//   mode 3: as mode 0 with box 64 ch x 128 px (16 KB per store, one store per piece by one warp)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/store_paths tools/microbench/store_paths.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

constexpr int CW = 2129, CH = 1205 * 4;   // four stacked 1080p canvases: 1.3 GB
constexpr int WARPS = 12;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int MODE>
__global__ void __launch_bounds__(WARPS * 32, 1)
store_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_row, uint8_t* dst) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < WARPS * 4096 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int n_strips = (CW + 127) / 128;
    const long long total = static_cast<long long>(n_strips) * CH;
    const int lo = static_cast<int>(total * blockIdx.x / gridDim.x), hi = static_cast<int>(total * (blockIdx.x + 1) / gridDim.x);
    const int grp = warp >> 2, q = warp & 3;
    const uint32_t buf = smem_u32(smem) + (MODE == 3 ? grp * 16384 : warp * 4096);
    for (int p = lo + grp; p < hi; p += WARPS / 4) {   // three groups of four warps take rows round-robin
        const int strip = p / CH, y = p - strip * CH;
        const int x = strip * 128 + q * 32;
        if (MODE == 0) {
            if (lane == 0) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                             ::"l"(&map_q), "r"(buf), "r"(0), "r"(x), "r"(y) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else if (MODE == 3) {
            if (q == 0 && lane == 0) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                             ::"l"(&map_row), "r"(buf), "r"(0), "r"(strip * 128), "r"(y) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else if (MODE == 1) {
            if (x + 32 <= CW && lane == 0) {
                uint8_t* g = dst + (static_cast<long long>(y) * CW + x) * 128;
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(buf), "n"(4096) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            if (x + 32 <= CW) {
                uint4* g = reinterpret_cast<uint4*>(dst + (static_cast<long long>(y) * CW + x) * 128);
                const uint4 v = make_uint4(p, lane, warp, 7);
#pragma unroll
                for (int k = 0; k < 8; ++k) g[k * 32 + lane] = v;
            }
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename F>
static float time_ms(F&& f, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int i = 0; i < 2; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    const size_t bytes = static_cast<size_t>(CW) * CH * 128;
    uint8_t* d;
    cudaMalloc(&d, bytes);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    EncodeFn enc = reinterpret_cast<EncodeFn>(fn);
    CUtensorMap mq, mr;
    const cuuint64_t gdim[3] = {64, CW, CH};
    const cuuint64_t gstride[2] = {128, static_cast<cuuint64_t>(CW) * 128};
    const cuuint32_t estr[3] = {1, 1, 1};
    const cuuint32_t box_q[3] = {64, 32, 1}, box_r[3] = {64, 128, 1};
    if (enc(&mq, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, d, gdim, gstride, box_q, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        enc(&mr, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, d, gdim, gstride, box_r, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        printf("encode failed\n");
        return 1;
    }
    const int smem = WARPS * 4096 + 1024;
    cudaFuncSetAttribute(store_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(store_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(store_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(store_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int grid : {148, 296}) {
        const float t0 = time_ms([&] { store_kernel<0><<<grid, WARPS * 32, smem>>>(mq, mr, d); }, 5);
        const float t1 = time_ms([&] { store_kernel<1><<<grid, WARPS * 32, smem>>>(mq, mr, d); }, 5);
        const float t2 = time_ms([&] { store_kernel<2><<<grid, WARPS * 32, smem>>>(mq, mr, d); }, 5);
        const float t3 = time_ms([&] { store_kernel<3><<<grid, WARPS * 32, smem>>>(mq, mr, d); }, 5);
        printf("{\"grid\": %d, \"tma_tensor_4KB_GBps\": %.0f, \"bulk_1d_4KB_GBps\": %.0f, \"st_global_v4_GBps\": %.0f, \"tma_tensor_16KB_GBps\": %.0f}\n",
               grid, bytes / t0 / 1e6, bytes / t1 / 1e6, bytes / t2 / 1e6, bytes / t3 / 1e6);
    }
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
