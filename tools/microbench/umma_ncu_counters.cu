// What do ncu's two tensor-pipe counters read on a kernel whose tensor pipe is busy 100 % of the time?
// One CTA per SM issues back-to-back tcgen05.mma (M=128, K=16, fp16) of shape N into two accumulator regions; the
// micro-benchmarks of round 1 (profiles/umma_microbench_r01.txt) showed these run at exactly N/2 cycles each.  Run under
//   ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,...
// to see how `sm__pipe_tensor_cycles_active` and `..._realtime` relate (79.5 % vs 62.8 % on the
// chained kernel).  usage: umma_ncu_counters [n=192] [mmas=200000]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace reve;

__global__ void __launch_bounds__(128, 1) busy(int n, int total, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - raw);
    const uint32_t w_addr = base + 1024, ring = base + 1024 + 73728 + 1024;
    for (uint32_t i = threadIdx.x; i < (73728 + 1024 + 8 * 16384 + 1024) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(bp + 1024)[i] = 0x3c003c00u;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(base, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(base + 512, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(bp + 512);
    if (warp == 0) {
        const uint64_t proto = umma_desc_sw128(0, 0);
        const uint32_t hi = (uint32_t)(proto >> 32), lof = (uint32_t)proto;
        const uint32_t w_lo = lof | (w_addr >> 4), r_lo = lof | (ring >> 4);
        const uint32_t id = umma_idesc_f16(128, n);
        long long t0 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int i = 0; i < total; ++i) {
                const int dxk = i % 12, dx = dxk >> 2, k = dxk & 3;
                const uint32_t al = r_lo + ((i / 12) & 7) * 1024 + (dx - 1) * 8 + k * 2;
                const uint32_t bl = w_lo + dx * 1536 + k * 2;
                umma_f16(tmem + (i & 1) * n, ((uint64_t)hi << 32) | al, ((uint64_t)hi << 32) | bl, id, i >= 2 ? 1u : 0u);
            }
            umma_commit(base);
        }
        __syncwarp();
        mbar_wait(base, 0, nullptr, 0);
        if (elect_one() && blockIdx.x == 0) out[0] = clock64() - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 192, total = argc > 2 ? atoi(argv[2]) : 200000;
    const int smem = 1024 + 1024 + 73728 + 1024 + 8 * 16384 + 1024;
    cudaFuncSetAttribute(busy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* d; cudaMalloc(&d, 8);
    for (int rep = 0; rep < 3; ++rep) busy<<<148, 128, smem>>>(n, total, d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 1; }
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("N=%d: %.2f clk per MMA (ideal %d)\n", n, (double)h / total, n / 2);
    return 0;
}
