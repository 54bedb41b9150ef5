// Micro-benchmark v2: tcgen05.mma dependent-accumulate latency vs. independent accumulator rotation.
// Every MMA is M=128, K=16, fp16 in / fp32 accumulate, SW128 K-major operands in shared memory.
// A "program" is: R regions (accumulator tiles of N columns each, laid out back to back in TMEM),
// MMAs issued round-robin over the regions.  Reports SM cycles and ns per MMA and per 64-column unit.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace reve;

struct Result { long long cycles; long long ns; int mmas; };

__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

__global__ void __launch_bounds__(128, 1) bench(int n, int regions, int per_region_run, int total, int shiftA, Result* res) {
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
        long long t0 = 0, g0 = 0;
        if (elect_one()) {
            g0 = gtime(); t0 = clock64();
            int reg = 0, run = 0;
            for (int i = 0; i < total; ++i) {
                const int dxk = i % 12, dx = dxk >> 2, k = dxk & 3;
                const uint32_t al = r_lo + ((i / 12) & 7) * 1024 + (shiftA ? (dx - 1) * 8 : 0) + k * 2;
                const uint32_t bl = w_lo + dx * 1536 + k * 2;
                umma_f16(tmem + reg * n, ((uint64_t)hi << 32) | al, ((uint64_t)hi << 32) | bl, id, i >= regions ? 1u : 0u);
                if (++run == per_region_run) { run = 0; if (++reg == regions) reg = 0; }
            }
            umma_commit(base);
        }
        __syncwarp();
        mbar_wait(base, 0, nullptr, 0);
        const long long t1 = clock64(), g1 = gtime();
        if (elect_one() && blockIdx.x == 0) { res->cycles = t1 - t0; res->ns = g1 - g0; res->mmas = total; }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    const int smem = 1024 + 1024 + 73728 + 1024 + 8 * 16384 + 1024;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    Result* d; cudaMalloc(&d, sizeof(Result));
    struct Cfg { int n, regions, run; } cfgs[] = {
        {256, 1, 1}, {256, 2, 1}, {192, 1, 1}, {192, 2, 1}, {192, 2, 12}, {128, 1, 1}, {128, 2, 1}, {128, 3, 1}, {128, 4, 1},
        {64, 1, 1}, {64, 2, 1}, {64, 3, 1}, {64, 4, 1}, {64, 6, 1}, {64, 8, 1}, {96, 4, 1}, {96, 5, 1}, {160, 3, 1}, {240, 2, 1},
        {32, 8, 1}, {16, 8, 1}, {48, 8, 1}, {256, 1, 1}};
    // spin the clocks up
    for (int w = 0; w < 200; ++w) bench<<<148, 128, smem>>>(256, 2, 1, 24000, 1, d);
    cudaDeviceSynchronize();
    printf("%5s %7s %4s | %10s %10s %9s %9s %8s\n", "N", "regions", "run", "clk/MMA", "ns/MMA", "clk/unit64", "eff", "MHz");
    for (auto c : cfgs) {
        for (int shiftA = 1; shiftA >= 0; --shiftA) {
            double best_clk = 1e30, best_ns = 1e30;
            for (int rep = 0; rep < 4; ++rep) {
                Result h{};
                bench<<<148, 128, smem>>>(c.n, c.regions, c.run, 24000, shiftA, d);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 1; }
                cudaMemcpy(&h, d, sizeof h, cudaMemcpyDeviceToHost);
                if (rep == 0) continue;
                best_clk = std::min(best_clk, (double)h.cycles / h.mmas);
                best_ns = std::min(best_ns, (double)h.ns / h.mmas);
            }
            const double unit = best_clk / (c.n / 64.0);
            printf("%5d %7d %4d | %10.1f %10.1f %9.1f %8.1f%% %8.0f  %s\n", c.n, c.regions, c.run, best_clk, best_ns, unit,
                   100.0 * 32.0 / unit, best_clk / best_ns * 1000.0, shiftA ? "shifted-A" : "aligned-A");
        }
    }
    return 0;
}
