// Micro-benchmark v3: the conv kernel's exact per-row MMA sequence (sliding 3-slot window over an
// 8-slot TMEM ring, split at the wrap), issued by one thread with no other traffic.
// mode bit0: add the two tcgen05.commit per row; bit1: add two (always-satisfied) mbarrier try_waits per row;
// bit2: never wrap (s0 cycles 0..5); bit3: skip the N=64/N=128 first-step split (12 x N=192)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace reve;
struct Result { long long cycles; long long ns; int rows; };
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint64_t mk(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

__global__ void __launch_bounds__(320, 1) bench(int mode, int rows, Result* res) {
    constexpr int NG = 64; constexpr uint32_t kG = NG * 8, kDx = 3 * NG * 8;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - raw);
    const uint32_t w_addr = base + 1024, ring = base + 1024 + 73728 + 1024;
    for (uint32_t i = threadIdx.x; i < (73728 + 1024 + 8 * 16384 + 1024) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(bp + 1024)[i] = 0x3c003c00u;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { for (int b = 0; b < 20; ++b) mbar_init(base + 8 * b, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(base + 512, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(bp + 512);
    if (warp == 0) {
        const uint64_t proto = umma_desc_sw128(0, 0);
        const uint32_t desc_hi = (uint32_t)(proto >> 32), lof = (uint32_t)proto;
        const uint32_t w_lo = lof | (w_addr >> 4), ring_lo = lof | (ring >> 4);
        const uint32_t idesc1 = umma_idesc_f16(128, 64), idesc2 = umma_idesc_f16(128, 128), idesc3 = umma_idesc_f16(128, 192);
        long long t0 = clock64(), g0 = gtime();
        for (int r = 0; r < rows; ++r) {
            const uint32_t stage = r & 7;
            const uint32_t a_lo = ring_lo + stage * 1024;
            const int s0 = (mode & 4) ? (5 - (r % 6)) : ((-(r + 1)) & 7);
            if (mode & 2) {
                // barriers 16,17 never complete a phase: waiting on parity 1 of a fresh barrier succeeds at once
                const bool a = mbar_try_wait(base + 8 * 16, 1), b = mbar_try_wait(base + 8 * 17, 1);
                if (!(a && b)) asm volatile("trap;");
            }
            tc_fence_after();
            if (elect_one()) {
                const uint64_t a0 = mk(desc_hi, a_lo - 8);
                if (s0 <= 5) {
                    const uint32_t d = tmem_base + s0 * NG;
                    if (mode & 8) umma_f16(d, a0, mk(desc_hi, w_lo), idesc3, 1u);
                    else { umma_f16(d, a0, mk(desc_hi, w_lo), idesc1, 0u); umma_f16(d + NG, a0, mk(desc_hi, w_lo + kG), idesc2, 1u); }
#pragma unroll
                    for (int dxk = 1; dxk < 12; ++dxk) {
                        const int dx = dxk >> 2, k = dxk & 3;
                        umma_f16(d, mk(desc_hi, a_lo + (dx - 1) * 8 + k * 2), mk(desc_hi, w_lo + dx * kDx + k * 2), idesc3, 1u);
                    }
                } else if (s0 >= 6 && (mode & 128)) {
                    // grouped order: all K-steps of the first part, then all K-steps of the second part
                    const int na = (s0 == 6) ? 2 : 1;                       // groups before the wrap
                    const uint32_t da = tmem_base + s0 * NG;
                    const uint32_t ida = na == 2 ? idesc2 : idesc1, idb = na == 2 ? idesc1 : idesc2;
                    umma_f16(da, a0, mk(desc_hi, w_lo), idesc1, 0u);
                    if (na == 2) umma_f16(da + NG, a0, mk(desc_hi, w_lo + kG), idesc1, 1u);
#pragma unroll
                    for (int dxk = 1; dxk < 12; ++dxk) {
                        const int dx = dxk >> 2, k = dxk & 3;
                        umma_f16(da, mk(desc_hi, a_lo + (dx - 1) * 8 + k * 2), mk(desc_hi, w_lo + dx * kDx + k * 2), ida, 1u);
                    }
#pragma unroll
                    for (int dxk = 0; dxk < 12; ++dxk) {
                        const int dx = dxk >> 2, k = dxk & 3;
                        umma_f16(tmem_base, mk(desc_hi, a_lo + (dx - 1) * 8 + k * 2), mk(desc_hi, w_lo + dx * kDx + k * 2 + na * kG), idb, 1u);
                    }
                } else if (s0 == 6 && (mode & 64)) {
                    const uint32_t d6 = tmem_base + 6 * NG;
                    umma_f16_a<ACollector::FILL>(d6, a0, mk(desc_hi, w_lo), idesc1, 0u);
                    umma_f16_a<ACollector::USE>(d6 + NG, a0, mk(desc_hi, w_lo + kG), idesc1, 1u);
                    umma_f16_a<ACollector::LASTUSE>(tmem_base, a0, mk(desc_hi, w_lo + 2 * kG), idesc1, 1u);
#pragma unroll
                    for (int dxk = 1; dxk < 12; ++dxk) {
                        const int dx = dxk >> 2, k = dxk & 3;
                        const uint64_t ad = mk(desc_hi, a_lo + (dx - 1) * 8 + k * 2);
                        umma_f16_a<ACollector::FILL>(d6, ad, mk(desc_hi, w_lo + dx * kDx + k * 2), idesc2, 1u);
                        umma_f16_a<ACollector::LASTUSE>(tmem_base, ad, mk(desc_hi, w_lo + dx * kDx + k * 2 + 2 * kG), idesc1, 1u);
                    }
                } else if (s0 == 7 && (mode & 64)) {
                    const uint32_t d7 = tmem_base + 7 * NG;
                    umma_f16_a<ACollector::FILL>(d7, a0, mk(desc_hi, w_lo), idesc1, 0u);
                    umma_f16_a<ACollector::LASTUSE>(tmem_base, a0, mk(desc_hi, w_lo + kG), idesc2, 1u);
#pragma unroll
                    for (int dxk = 1; dxk < 12; ++dxk) {
                        const int dx = dxk >> 2, k = dxk & 3;
                        const uint64_t ad = mk(desc_hi, a_lo + (dx - 1) * 8 + k * 2);
                        umma_f16_a<ACollector::FILL>(d7, ad, mk(desc_hi, w_lo + dx * kDx + k * 2), idesc1, 1u);
                        umma_f16_a<ACollector::LASTUSE>(tmem_base, ad, mk(desc_hi, w_lo + dx * kDx + k * 2 + kG), idesc2, 1u);
                    }
                } else if (s0 == 6) {
                    const uint32_t d6 = tmem_base + 6 * NG;
                    umma_f16(d6, a0, mk(desc_hi, w_lo), idesc1, 0u);
                    umma_f16(d6 + NG, a0, mk(desc_hi, w_lo + kG), idesc1, 1u);
                    umma_f16(tmem_base, a0, mk(desc_hi, w_lo + 2 * kG), idesc1, 1u);
#pragma unroll
                    for (int dxk = 1; dxk < 12; ++dxk) {
                        const int dx = dxk >> 2, k = dxk & 3;
                        const uint64_t ad = mk(desc_hi, a_lo + (dx - 1) * 8 + k * 2);
                        umma_f16(d6, ad, mk(desc_hi, w_lo + dx * kDx + k * 2), idesc2, 1u);
                        umma_f16(tmem_base, ad, mk(desc_hi, w_lo + dx * kDx + k * 2 + 2 * kG), idesc1, 1u);
                    }
                } else {
                    const uint32_t d7 = tmem_base + 7 * NG;
                    umma_f16(d7, a0, mk(desc_hi, w_lo), idesc1, 0u);
                    umma_f16(tmem_base, a0, mk(desc_hi, w_lo + kG), idesc2, 1u);
#pragma unroll
                    for (int dxk = 1; dxk < 12; ++dxk) {
                        const int dx = dxk >> 2, k = dxk & 3;
                        const uint64_t ad = mk(desc_hi, a_lo + (dx - 1) * 8 + k * 2);
                        umma_f16(d7, ad, mk(desc_hi, w_lo + dx * kDx + k * 2), idesc1, 1u);
                        umma_f16(tmem_base, ad, mk(desc_hi, w_lo + dx * kDx + k * 2 + kG), idesc2, 1u);
                    }
                }
                if (mode & 1) { umma_commit(base + 8 * (2 + stage)); umma_commit(base + 8 * (10 + (s0 & 3))); }
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(base);
        __syncwarp();
        mbar_wait(base, 0, nullptr, 0);
        const long long t1 = clock64(), g1 = gtime();
        if (elect_one() && blockIdx.x == 0) { res->cycles = t1 - t0; res->ns = g1 - g0; res->rows = rows; }
    }
    else if (mode & 16) {
        // spinning bystanders (like epilogue warps waiting for accumulators): bit5 = with nanosleep backoff
        if (mode & 32) { while (!mbar_try_wait(base + 8 * 18, 0)) __nanosleep(200); }
        else mbar_wait(base + 8 * 18, 0, nullptr, 0);
    }
    if (warp == 0 && elect_one()) mbar_arrive(base + 8 * 18);
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int main() {
    const int smem = 1024 + 1024 + 73728 + 1024 + 8 * 16384 + 1024;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    Result* d; cudaMalloc(&d, sizeof(Result));
    for (int w = 0; w < 100; ++w) bench<<<148, 320, smem>>>(0, 2000, d);
    cudaDeviceSynchronize();
    for (int mode : {3, 67, 131, 0, 64, 128, 4}) {
        double best = 1e30, bestns = 0;
        for (int rep = 0; rep < 3; ++rep) {
            Result h{};
            bench<<<148, 320, smem>>>(mode, 2000, d);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("mode %d failed\n", mode); return 1; }
            cudaMemcpy(&h, d, sizeof h, cudaMemcpyDeviceToHost);
            if ((double)h.cycles / h.rows < best) { best = (double)h.cycles / h.rows; bestns = (double)h.ns / h.rows; }
        }
        printf("mode %2d [%s%s%s%s%s]: %8.1f clk/row %8.1f ns/row\n", mode, mode & 1 ? "commit " : "", mode & 2 ? "trywait " : "",
               mode & 4 ? "nowrap " : "", mode & 8 ? "12xN192 " : "", mode & 64 ? "A-collector reuse in wrap rows " : (mode & 128 ? "grouped wrap order " : ""), best, bestns);
    }
    return 0;
}
