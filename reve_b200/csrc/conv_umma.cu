// 3x3 convolution over the fp16 NHWC canvas as an implicit GEMM on tcgen05 tensor cores.
//
// Replaces the 16 body `Convolution 64->64 + PReLU` layers and the final `Convolution 64->3*s*s`
// + PixelShuffle + Interp(nearest) + BinaryOp(add) + u8 post-processing that the upscaler spawned
// at reference reve-shared/src/lib.rs:134-147 runs through ncnn.
//
// Work decomposition
//   * The canvas is cut into vertical strips of 126 output columns.  One strip row is one UMMA
//     M-tile: a TMA box of 128 pixels x 64 channels (x0-1 .. x0+126, one 128-byte swizzle row per
//     pixel) lands in a 16 KB ring slot; the three horizontal taps are the same slot read through
//     descriptors shifted by -1/0/+1 pixel (128 B).  Output columns 0 and 127 of the M-tile are
//     halo garbage and are never stored.
//   * The three vertical taps are stacked along N: for input row y the B operand is
//     [W(ky=0) | W(ky=1) | W(ky=2)] (N = 3*NG), so one read of the A tile feeds output rows
//     y+1, y, y-1, whose accumulators sit in adjacent NG-column slots of an 8-slot TMEM ring
//     (slot(t) = (-t) mod 8, so the three slots are ascending and contiguous except at the
//     wrap, where the MMA is split).  Every input row is fetched from L2/HBM exactly once per
//     strip and A is read from shared memory 12 times per row instead of 36.
//   * A CTA owns a contiguous range of strip-rows (grid = #SMs, persistent), possibly spanning two
//     strips; segments restart the 3-row window with their own halo rows.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> bias/PReLU -> fp16 -> swizzled smem -> TMA store, or
// for the tail: bias -> PixelShuffle + residual -> u8 -> global).
#include "kernels.h"

#include <algorithm>
#include <cstring>

#include "model.h"

namespace reve {

namespace {

constexpr int kRowBytes = kBoxPx * 128;  // 16 KB: 128 px * 64 ch * fp16
constexpr int kCtrlBytes = 1024;
constexpr int kGuard = 1024;

__host__ __device__ constexpr int w_bytes(int ng) { return 3 * 3 * ng * 128; }
__host__ __device__ constexpr int tmem_cols(int ng) { return ng * 8 <= 128 ? 128 : (ng * 8 <= 256 ? 256 : 512); }

// control block offsets (from the 1024-aligned base)
constexpr int kBarW = 0;
constexpr int kBarAFull = 8;
constexpr int kBarAEmpty = kBarAFull + 8 * kStages;
constexpr int kBarAccFull = kBarAEmpty + 8 * kStages;
constexpr int kBarAccEmpty = kBarAccFull + 8 * 8;
constexpr int kTmemPtr = 512;

enum : uint32_t { TAG_W = 1, TAG_A_EMPTY = 2, TAG_A_FULL = 3, TAG_ACC_EMPTY = 4, TAG_ACC_FULL = 5 };

struct SegIter {
    long long lo, hi;
    int ch;
    __device__ SegIter(const ConvParams& p) : ch(p.canvas_h) {
        lo = static_cast<long long>(blockIdx.x) * p.total_rows / gridDim.x;
        hi = static_cast<long long>(blockIdx.x + 1) * p.total_rows / gridDim.x;
    }
    __device__ bool next(int& strip, int& ya, int& yb) {
        if (lo >= hi) return false;
        strip = static_cast<int>(lo / ch);
        ya = static_cast<int>(lo % ch);
        const long long n = min(static_cast<long long>(ch - ya), hi - lo);
        yb = ya + static_cast<int>(n) - 1;
        lo += n;
        return true;
    }
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

template <int NG, bool TAIL>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_umma_kernel(const __grid_constant__ CUtensorMap in_map,
                    const __grid_constant__ CUtensorMap out_map,
                    const __grid_constant__ ConvParams p) {
    constexpr int kWBytes = w_bytes(NG);
    constexpr int kTmemCols = tmem_cols(NG);
    constexpr int kOffW = kCtrlBytes;
    constexpr int kOffRing = kOffW + kWBytes + kGuard;
    constexpr int kOffStage = kOffRing + kStages * kRowBytes + kGuard;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* const base_ptr = smem_raw + (base - raw);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    DebugBlock* const dbg = p.dbg;

    if (threadIdx.x == 0) {
        mbar_init(base + kBarW, 1);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(base + kBarAFull + 8 * s, 1);
            mbar_init(base + kBarAEmpty + 8 * s, 1);
        }
        for (int s = 0; s < 8; ++s) {
            mbar_init(base + kBarAccFull + 8 * s, 1);
            mbar_init(base + kBarAccEmpty + 8 * s, 128);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(base + kTmemPtr, kTmemCols);
        tmem_relinquish();
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&in_map);
        if (!TAIL) prefetch_tmap(&out_map);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + kTmemPtr);

    const int CH = p.canvas_h;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            mbar_arrive_expect_tx(base + kBarW, kWBytes);
            bulk_load_1d(base + kOffW, p.weights, kWBytes, base + kBarW);
            SegIter it(p);
            int strip, ya, yb;
            uint32_t i = 0;
            while (it.next(strip, ya, yb)) {
                const int x0 = strip * kStripPx;
                const int y_lo = max(ya - 1, 0), y_hi = min(yb + 1, CH - 1);
                for (int y = y_lo; y <= y_hi; ++y, ++i) {
                    const uint32_t stage = i % kStages, use = i / kStages;
                    mbar_wait(base + kBarAEmpty + 8 * stage, (use & 1) ^ 1, dbg, TAG_A_EMPTY, i);
                    mbar_arrive_expect_tx(base + kBarAFull + 8 * stage, kRowBytes);
                    tma_load_3d(base + kOffRing + stage * kRowBytes, &in_map, base + kBarAFull + 8 * stage,
                                0, x0 - 1, y);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            mbar_wait(base + kBarW, 0, dbg, TAG_W);
            tc_fence_after();
            const uint32_t idesc1 = umma_idesc_f16(128, NG);
            const uint32_t idesc2 = umma_idesc_f16(128, 2 * NG);
            const uint32_t idesc3 = umma_idesc_f16(128, 3 * NG);
            const uint32_t w_addr = base + kOffW;
            SegIter it(p);
            int strip, ya, yb;
            uint32_t i = 0;
            int t_base = 0;
            while (it.next(strip, ya, yb)) {
                const int y_lo = max(ya - 1, 0), y_hi = min(yb + 1, CH - 1);
                for (int y = y_lo; y <= y_hi; ++y, ++i) {
                    const uint32_t stage = i % kStages, use = i / kStages;
                    // group g (0..2) = vertical tap ky = g: input row y feeds output row y + 1 - g
                    const int g_lo = (y + 1 <= yb) ? 0 : ((y <= yb) ? 1 : 2);
                    const int g_hi = (y - 1 >= ya) ? 2 : ((y >= ya) ? 1 : 0);
                    // groups whose output row receives its first contribution from this input row
                    const int fresh_hi = (y == 0) ? g_hi : ((g_lo == 0) ? 0 : g_lo - 1);
                    const int t0 = t_base + (y + 1 - ya);  // sequence number of output row y+1
                    const int s0 = (-t0) & 7;              // its TMEM slot; group g -> (s0 + g) & 7
                    for (int g = g_lo; g <= fresh_hi; ++g) {
                        const int tg = t0 - g;
                        mbar_wait(base + kBarAccEmpty + 8 * ((s0 + g) & 7), ((tg >> 3) & 1) ^ 1, dbg,
                                  TAG_ACC_EMPTY, tg);
                    }
                    mbar_wait(base + kBarAFull + 8 * stage, use & 1, dbg, TAG_A_FULL, i);
                    tc_fence_after();

                    const uint32_t a_row = base + kOffRing + stage * kRowBytes;
                    auto issue = [&](uint64_t adesc, uint32_t w_k, int ga, int gb, uint32_t acc) {
                        if (ga > gb) return;
                        const int sa = (s0 + ga) & 7;
                        const int n = gb - ga + 1;
                        const int n1 = min(n, 8 - sa);
                        umma_f16(tmem_base + sa * NG, adesc, umma_desc_sw128(w_k + ga * NG * 128, 0),
                                 n1 == 1 ? idesc1 : (n1 == 2 ? idesc2 : idesc3), acc);
                        if (n1 < n) {
                            const int n2 = n - n1;
                            umma_f16(tmem_base, adesc, umma_desc_sw128(w_k + (ga + n1) * NG * 128, 0),
                                     n2 == 1 ? idesc1 : idesc2, acc);
                        }
                    };
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t a_addr = a_row + (dx - 1) * 128 + k * 32;
                            // Measured on B200 (profiles/r01_notes.md): the 128B-swizzle XOR is taken
                            // from the absolute shared-memory address bits, so a start address shifted
                            // by whole 128-byte rows needs base_offset = 0 (the "(addr >> 7) & 7"
                            // formula of the PTX manual produces garbage here).
                            const uint64_t adesc = umma_desc_sw128(a_addr, 0);
                            const uint32_t w_k = w_addr + dx * (3 * NG * 128) + k * 32;
                            if (dx == 0 && k == 0 && fresh_hi >= g_lo) {
                                issue(adesc, w_k, g_lo, fresh_hi, 0u);
                                issue(adesc, w_k, fresh_hi + 1, g_hi, 1u);
                            } else {
                                issue(adesc, w_k, g_lo, g_hi, 1u);
                            }
                        }
                    }
                    umma_commit(base + kBarAEmpty + 8 * stage);  // A slot reusable once these MMAs retire
                    if (g_hi == 2) umma_commit(base + kBarAccFull + 8 * ((s0 + 2) & 7));  // row y-1 done
                    if (y == CH - 1 && yb == CH - 1)                                       // bottom edge: row y done too
                        umma_commit(base + kBarAccFull + 8 * ((s0 + 1) & 7));
                }
                t_base += yb - ya + 1;
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps
        const int q = warp & 3;            // TMEM lane quarter this warp may access
        const int m = q * 32 + lane;       // M row = pixel index inside the 128-px box
        const bool leader = (threadIdx.x == 64);
        const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        SegIter it(p);
        int strip, ya, yb;
        int t = 0;
        uint32_t rowc = 0;
        while (it.next(strip, ya, yb)) {
            const int x0 = strip * kStripPx;
            const int cx = x0 - 1 + m;
            const bool inside = (m >= 1) && (m <= kStripPx) && (cx < p.canvas_w);
            const bool colok = inside && (p.colflag[cx] != 0);
            int ox = -1, sx = 0;
            if (TAIL && inside) {
                ox = p.out_x[cx];
                sx = p.src_x[cx];
            }
            for (int r = ya; r <= yb; ++r, ++t) {
                const int s = (-t) & 7;
                mbar_wait(base + kBarAccFull + 8 * s, (t >> 3) & 1, dbg, TAG_ACC_FULL, t);
                tc_fence_after();
                uint32_t acc[NG];
#pragma unroll
                for (int c = 0; c < NG / 16; ++c) {
                    uint32_t(&dst)[16] = *reinterpret_cast<uint32_t(*)[16]>(&acc[c * 16]);
                    tmem_ld16(tmem_lane + s * NG + c * 16, dst);
                }
                tmem_wait_ld();
                tc_fence_before();
                mbar_arrive(base + kBarAccEmpty + 8 * s);

                if constexpr (!TAIL) {
                    const bool keep = colok && (p.rowflag[r] != 0);
                    const uint32_t stg = base + kOffStage + (rowc & 1) * kRowBytes;
                    if (leader) bulk_wait_read<1>();  // the store issued two rows ago has drained this buffer
                    named_bar_sync(1, 128);
                    if (m >= 1 && m <= kStripPx) {
                        const int row = m - 1;
                        const uint32_t rbase = stg + row * 128;
#pragma unroll
                        for (int c8 = 0; c8 < 8; ++c8) {
                            uint32_t pk[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int ch = c8 * 8 + j * 2;
                                float v0 = __uint_as_float(acc[ch]) + p.bias[ch];
                                float v1 = __uint_as_float(acc[ch + 1]) + p.bias[ch + 1];
                                v0 = fmaxf(v0, 0.f) + p.slope[ch] * fminf(v0, 0.f);
                                v1 = fmaxf(v1, 0.f) + p.slope[ch + 1] * fminf(v1, 0.f);
                                pk[j] = keep ? pack_half2(v0, v1) : 0u;
                            }
                            st_shared_v4(rbase + ((c8 ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
                        }
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(1, 128);
                    if (leader) {
                        tma_store_3d(&out_map, stg, 0, x0, r);
                        bulk_commit();
                    }
                    ++rowc;
                } else {
                    constexpr int S = (NG == 16) ? 2 : ((NG == 32) ? 3 : 4);
                    const int oy = p.out_y[r];
                    if (ox >= 0 && oy >= 0) {
                        const uint8_t* sp = p.src + static_cast<long long>(p.src_y[r]) * p.src_stride + sx * 3;
                        const float xin[3] = {static_cast<float>(sp[0]), static_cast<float>(sp[1]),
                                              static_cast<float>(sp[2])};
#pragma unroll
                        for (int i = 0; i < S; ++i) {
                            uint8_t* dp = p.dst + static_cast<long long>(oy * S + i) * p.dst_stride +
                                          static_cast<long long>(ox) * (S * 3);
#pragma unroll
                            for (int j = 0; j < S; ++j) {
#pragma unroll
                                for (int c = 0; c < 3; ++c) {
                                    const int idx = c * S * S + i * S + j;
                                    const float v = __uint_as_float(acc[idx]) + p.bias[idx];
                                    // y = r + x/255 ; u8 = clamp(floor(y*255 + 0.5))
                                    float o = floorf(fmaf(v, 255.f, xin[c] + 0.5f));
                                    o = fminf(fmaxf(o, 0.f), 255.f);
                                    dp[j * 3 + c] = static_cast<uint8_t>(o);
                                }
                            }
                        }
                    }
                }
            }
        }
        if (!TAIL && leader) bulk_wait<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <int NG, bool TAIL>
constexpr size_t smem_bytes_t() {
    return 1024 /*alignment slack*/ + kCtrlBytes + w_bytes(NG) + kGuard + kStages * kRowBytes + kGuard +
           (TAIL ? 0 : 2 * kRowBytes);
}

}  // namespace

size_t conv_smem_bytes(int ng, bool tail) {
    return 1024 + kCtrlBytes + w_bytes(ng) + kGuard + kStages * kRowBytes + kGuard + (tail ? 0 : 2 * kRowBytes);
}
size_t conv_weight_blob_bytes(int ng) { return w_bytes(ng); }

void pack_conv_weights(const float* w_oihw, int co, int ng, uint16_t* blob) {
    // blob[dx][n = g*ng + o][ci] fp16 with the 128-byte swizzle applied per 128-byte row:
    // 16-byte chunk c of row n is stored at chunk (c ^ (n & 7)).
    std::memset(blob, 0, w_bytes(ng));
    for (int dx = 0; dx < 3; ++dx)
        for (int g = 0; g < 3; ++g)
            for (int o = 0; o < co; ++o) {
                const int n = g * ng + o;
                for (int ci = 0; ci < 64; ++ci) {
                    const float v = w_oihw[((static_cast<size_t>(o) * 64 + ci) * 3 + g) * 3 + dx];
                    const size_t byte = static_cast<size_t>(dx) * (3 * ng * 128) + static_cast<size_t>(n) * 128 +
                                        (((ci >> 3) ^ (n & 7)) << 4) + (ci & 7) * 2;
                    blob[byte / 2] = f32_to_f16(v);
                }
            }
}

cudaError_t conv_kernels_init() {
    cudaError_t e;
    e = cudaFuncSetAttribute(conv3x3_umma_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem_bytes_t<64, false>()));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(conv3x3_umma_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem_bytes_t<16, true>()));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(conv3x3_umma_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem_bytes_t<32, true>()));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(conv3x3_umma_kernel<48, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem_bytes_t<48, true>()));
    return e;
}

cudaError_t launch_conv_body(cudaStream_t st, int grid, const CUtensorMap& in_map, const CUtensorMap& out_map,
                             const ConvParams& p) {
    conv3x3_umma_kernel<64, false><<<grid, kConvThreads, smem_bytes_t<64, false>(), st>>>(in_map, out_map, p);
    return cudaGetLastError();
}

cudaError_t launch_conv_tail(cudaStream_t st, int grid, int scale, const CUtensorMap& in_map, const ConvParams& p) {
    switch (scale) {
        case 2: conv3x3_umma_kernel<16, true><<<grid, kConvThreads, smem_bytes_t<16, true>(), st>>>(in_map, in_map, p); break;
        case 3: conv3x3_umma_kernel<32, true><<<grid, kConvThreads, smem_bytes_t<32, true>(), st>>>(in_map, in_map, p); break;
        case 4: conv3x3_umma_kernel<48, true><<<grid, kConvThreads, smem_bytes_t<48, true>(), st>>>(in_map, in_map, p); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace reve
