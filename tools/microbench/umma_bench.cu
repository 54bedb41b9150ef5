// Micro-benchmark: tcgen05.mma issue patterns (M=128, K=16, fp16, SW128 K-major operands in smem).
// Measures cycles per MMA for the accumulator-addressing patterns the conv kernel could use.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I reve_b200/csrc -o gpurun_out/umma_bench tools/microbench/umma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace reve;

struct Result { long long cycles; int mmas; };

__global__ void __launch_bounds__(128, 1) bench(int pattern, int rows, Result* res) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - raw);
    // layout: [0,1024) ctrl; weights 72 KB @1024; guard 1 KB; A ring 8 x 16 KB
    const uint32_t w_addr = base + 1024, ring = base + 1024 + 73728 + 1024;
    for (uint32_t i = threadIdx.x; i < (73728 + 1024 + 8 * 16384 + 1024) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(bp + 1024)[i] = 0x3c003c00u;  // fp16 1.0
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
        const uint32_t id64 = umma_idesc_f16(128, 64), id128 = umma_idesc_f16(128, 128), id192 = umma_idesc_f16(128, 192),
                       id256 = umma_idesc_f16(128, 256);
        int mmas = 0;
        long long t0 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int r = 0; r < rows; ++r) {
                const uint32_t a_lo = r_lo + (r & 7) * 1024;
                for (int dxk = 0; dxk < 12; ++dxk) {
                    const int dx = dxk >> 2, k = dxk & 3;
                    uint32_t al = a_lo + (dx - 1) * 8 + k * 2;
                    uint32_t bl = w_lo + dx * 1536 + k * 2;
                    uint32_t d = tmem, id = id192;
                    switch (pattern) {
                        case 0: break;                                    // same D, N=192, shifted A
                        case 1: al = a_lo + k * 2; break;                 // aligned A only
                        case 2: d = tmem + (5 - (r % 6)) * 64; break;     // D slides by -64 cols per row
                        case 3: d = tmem + (5 - (r % 6)) * 64; break;     // + per-group first MMAs (below)
                        case 4: id = id64; break;                         // 3 x N=64 per (dx,k)
                        case 5: id = id256; break;                        // N=256 same D
                        case 6: d = tmem + (r & 1) * 256; break;          // two disjoint D regions
                        case 7: id = id128; break;                        // N=128 same D
                        case 8: d = tmem + (r & 3) * 64; break;           // D slides +64 per row, wraps every 4
                        case 9: al = a_lo + k * 2; bl = w_lo + k * 2; break;  // aligned A, fixed B block
                        case 10: id = id256; d = tmem + 128; break;            // N=256 straddling the halves
                        case 11: d = tmem + 256; break;                        // N=192 same D in upper half
                        case 12: d = tmem + (dxk & 1) * 256; break;            // alternate halves every MMA
                        case 13: d = tmem + (dxk & 1) * 192; break;            // alternate disjoint regions, cols 0/192
                        case 14: id = id128; d = tmem + (dxk & 1) * 128; break; // N=128 alternate 0/128
                        case 15: id = id64; break;                             // N=64 same D
                        case 16: break;                                        // N=192 same D, never accumulate (below)
                        case 17: break;                                        // N=192 same D, acc=0 on first of row
                        case 18: id = umma_idesc_f16(128, 224); break;         // N=224 same D
                        case 19: id = umma_idesc_f16(128, 208); break;         // N=208 same D
                        case 20: id = umma_idesc_f16(128, 240); break;         // N=240
                        case 21: id = umma_idesc_f16(128, 160); break;         // N=160
                        case 22: id = umma_idesc_f16(128, 96); break;          // N=96
                        case 23: id = id128; d = tmem + (r & 1) * 256; break;  // N=128 alternate halves per row
                        default: break;
                    }
                    if (pattern == 4) {
                        for (int g = 0; g < 3; ++g) {
                            umma_f16(tmem + g * 64, ((uint64_t)hi << 32) | al, ((uint64_t)hi << 32) | (bl + g * 512), id64, 1u);
                            ++mmas;
                        }
                    } else if (pattern == 3 && dxk == 0) {
                        for (int g = 0; g < 3; ++g) {
                            umma_f16(d + g * 64, ((uint64_t)hi << 32) | al, ((uint64_t)hi << 32) | (bl + g * 512), id64, g ? 1u : 0u);
                            ++mmas;
                        }
                    } else {
                        const uint32_t accf = (pattern == 16) ? 0u : ((pattern == 17 && dxk == 0) ? 0u : 1u);
                        umma_f16(d, ((uint64_t)hi << 32) | al, ((uint64_t)hi << 32) | bl, id, accf);
                        ++mmas;
                    }
                }
            }
            umma_commit(base);
        }
        __syncwarp();
        mbar_wait(base, 0, nullptr, 0);
        const long long t1 = clock64();
        if (elect_one() && blockIdx.x == 0) { res->cycles = t1 - t0; res->mmas = mmas; }
        // note: t0 only valid on the elected lane; re-elect picks the same lane
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    const int smem = 1024 + 1024 + 73728 + 1024 + 8 * 16384 + 1024;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    Result* d; cudaMalloc(&d, sizeof(Result));
    const char* names[] = {"same D N=192 shifted A", "same D N=192 aligned A", "D slides -64/row (6 slots)",
                           "slide + 3 single-group first MMAs", "N=64 x3 per step", "same D N=256", "two disjoint D regions alternate",
                           "same D N=128", "D slides +64/row wrap 4", "aligned A, fixed B",
                           "N=256 at col 128", "N=192 same D col 256", "N=192 alt halves per MMA", "N=192 alt 0/192 per MMA",
                           "N=128 alt 0/128 per MMA", "N=64 same D", "N=192 same D acc=0 always", "N=192 same D acc=0 first of row",
                           "N=224 same D", "N=208 same D", "N=240 same D", "N=160 same D", "N=96 same D", "N=128 alt halves per row"};
    for (int grid : {148}) {
        for (int p = 0; p < 24; ++p) {
            Result h{};
            bench<<<grid, 128, smem>>>(p, 200, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("pattern %d: %s\n", p, cudaGetErrorString(e)); return 1; }
            bench<<<grid, 128, smem>>>(p, 200, d);
            cudaDeviceSynchronize();
            cudaMemcpy(&h, d, sizeof h, cudaMemcpyDeviceToHost);
            printf("grid %3d pattern %d %-40s: %8lld cycles, %5d MMAs, %7.1f clk/MMA, %7.1f clk/row\n", grid, p, names[p],
                   h.cycles, h.mmas, (double)h.cycles / h.mmas, (double)h.cycles / 200);
        }
    }
    return 0;
}
