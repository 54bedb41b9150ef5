// First convolution (3 -> 64, 3x3) + PReLU on tcgen05 tensor cores, fused with the pre-processing
// of the upscaler spawned at reference reve-shared/src/lib.rs:134-147 (SURVEY.md section 2.3,
// K1 + K2): u8 RGB gather with the reflect-101 pre-pad, zero padding at tile borders, /255, canvas
// layout, fp16 NHWC out.
//
// The layer is a GEMM with K = 27 (padded to 32): M = 128 consecutive canvas pixels (raster order,
// no halo: the im2col row of a pixel already holds its 3x3x3 neighbourhood), N = 64.  Producer
// warps build the im2col A tile straight from the u8 frame into shared memory (exact integers
// 0..255 as fp16, canonical K-major no-swizzle core-matrix layout), one warp issues two K=16 MMAs
// per tile into a ring of TMEM accumulators, and three epilogue groups apply (acc/255 + bias), PReLU,
// gap zeroing and write the row-major fp16 tile with one TMA store.
// Warp roles (800 threads): warps 0-11 = three producer groups (tiles round-robin), warp 12 = TMEM
// allocator + MMA issuer, warps 13-24 = three epilogue groups (tiles round-robin).  Measured: the kernel is bound
// by warp-level throughput of producers AND epilogue together (3 + 3 groups: 0.121 ms per 1080p frame; 3 + 2: 0.131;
// 4 + 2: 0.133; 2 + 4: 0.133; 4 + 3: 0.125), not by HBM (2.7 TB/s of writes).  The src_x / src_y
// geometry tables are cached in shared memory so the gather needs one global round trip per tile.
#include "kernels.h"

#include <cstring>

#include "model.h"

namespace reve {

namespace {

constexpr int kProducerGroups = 3;           // groups of 4 warps take tiles round-robin
constexpr int kProducerWarps = 4 * kProducerGroups;
constexpr int kMmaWarp = kProducerWarps;
constexpr int kFirstEpiWarp = kMmaWarp + 1;  // 4 * kEpiGroups epilogue warps; TMEM lane quarter = warp % 4
constexpr int kEpiGroups = 3;                // epilogue groups of 4 warps, tiles round-robin
constexpr int kThreads = (kFirstEpiWarp + 4 * kEpiGroups) * 32;
constexpr int kMaxTableInts = 12288;         // src_x / src_y cached in shared memory when they fit (48 KB)
constexpr int kStagesA = 6;             // multiple of kProducerGroups: a group always fills the same stages
constexpr int kTileA = 128 * 64;       // 8 KB: 128 px x 32 k x fp16
constexpr int kAccBufs = 4;            // TMEM accumulator ring: 4 x 64 columns
constexpr int kWBytes0 = 64 * 64;      // 4 KB: 64 co x 32 k x fp16
constexpr int kStageOut = 128 * 128;   // 16 KB output staging per epilogue group

// control block
constexpr int kBarW = 0;
constexpr int kBarFull = 8;                          // [kStagesA]
constexpr int kBarEmpty = kBarFull + 8 * kStagesA;   // [kStagesA]
constexpr int kBarAccFull = kBarEmpty + 8 * kStagesA;  // [kAccBufs]
constexpr int kBarAccEmpty = kBarAccFull + 8 * kAccBufs;
constexpr int kTmemPtr = 512;
constexpr int kCtrl = 1024;
constexpr int kOffW = kCtrl;
constexpr int kOffA = kOffW + kWBytes0;
constexpr int kOffOut = kOffA + kStagesA * kTileA;
constexpr int kOffTab = kOffOut + kEpiGroups * kStageOut;
constexpr int kSmem = 1024 + kOffTab + kMaxTableInts * 4;

#ifndef REVE_CONV0_WAIT_SLEEP
#define REVE_CONV0_WAIT_SLEEP 0   // ns a waiting producer / epilogue warp sleeps between polls (0 = spin; see profiles/r02_notes.md)
#endif
enum : uint32_t { TAG0_W = 11, TAG0_EMPTY = 12, TAG0_FULL = 13, TAG0_ACC_EMPTY = 14, TAG0_ACC_FULL = 15 };

// K-major, no swizzle: 8-row x 16-byte core matrices; LBO = distance between core matrices that
// are adjacent in K, SBO = distance between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo >> 4) << 16;
    d |= static_cast<uint64_t>(sbo >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__global__ void __launch_bounds__(kThreads, 1)
conv0_umma_kernel(const __grid_constant__ CUtensorMap out_map, const __grid_constant__ Conv0Params p) {
    // out_map: the output canvas as a flat [pixels][64 ch] tensor with a 32-pixel box: every epilogue warp stores the
    // quarter of the tile it wrote with its own TMA store (as the chained body kernel does).
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* const base_ptr = smem_raw + (base - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    DebugBlock* const dbg = p.dbg;

    if (threadIdx.x == 0) {
        mbar_init(base + kBarW, 1);
        for (int s = 0; s < kStagesA; ++s) {
            mbar_init(base + kBarFull + 8 * s, 4);   // one arrive per producer warp
            mbar_init(base + kBarEmpty + 8 * s, 1);
        }
        for (int s = 0; s < kAccBufs; ++s) {
            mbar_init(base + kBarAccFull + 8 * s, 1);
            mbar_init(base + kBarAccEmpty + 8 * s, 4);
        }
        fence_mbar_init();
    }
    if (warp == kMmaWarp) {
        tmem_alloc(base + kTmemPtr, kAccBufs * 64);
        tmem_relinquish();
    }
    const int CW = p.canvas_w, CHh = p.canvas_h;
    const bool tab_smem = (CW + 2 * CHh) <= kMaxTableInts;
    int* const tab = reinterpret_cast<int*>(base_ptr + kOffTab);
    if (tab_smem) {
        for (int i = threadIdx.x; i < CW; i += kThreads) tab[i] = p.src_x[i];
        for (int i = threadIdx.x; i < CHh; i += kThreads) {
            tab[CW + i] = p.src_y[i];
            tab[CW + CHh + i] = p.row_frame[i];
        }
    }
    const int* const tx = tab_smem ? tab : p.src_x;
    const int* const ty = tab_smem ? tab + CW : p.src_y;
    const int* const tf = tab_smem ? tab + CW + CHh : p.row_frame;
    if (warp == kFirstEpiWarp && lane == 0) prefetch_tmap(&out_map);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + kTmemPtr);

    const unsigned npx = static_cast<unsigned>(CW) * static_cast<unsigned>(CHh);   // canvases are far below 2^31 pixels (launch_conv0 checks)
    const int n_tiles = static_cast<int>((npx + 127u) / 128u);

    if (warp < kProducerWarps) {
        // ------------------------------------------------------------------ im2col producers
        const int pg = warp >> 2;
        const int m = (warp & 3) * 32 + lane;
        uint32_t j = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
            if ((j % kProducerGroups) != static_cast<uint32_t>(pg)) continue;
            const uint32_t stage = j % kStagesA, use = j / kStagesA;
            const unsigned px = static_cast<unsigned>(tile) * 128u + m;
            // Every lane gathers its OWN canvas column for the three rows of the stencil (3 x 3 bytes); the left and
            // right columns come from the neighbouring lanes by shuffle (consecutive lanes are consecutive canvas
            // pixels), lanes 0 and 31 fetch the column beyond the warp themselves.  9 (+9 predicated) byte loads per
            // pixel instead of 27: the kernel is instruction-issue-bound in exactly this code.
            const bool in = px < npx;
            const int cy = in ? static_cast<int>(px / static_cast<unsigned>(CW)) : 0;
            const int cx = in ? static_cast<int>(px - static_cast<unsigned>(cy) * CW) : 0;
            const int xe = (lane == 0) ? cx - 1 : cx + 1;                      // the extra column of lanes 0 / 31
            const bool edge = (lane == 0) || (lane == 31);
            const int sx = in ? tx[cx] : -1;
            const int sxe = (edge && in && xe >= 0 && xe < CW) ? tx[xe] : -1;
            const uint8_t* const fsrc = p.src[in ? max(tf[cy], 0) : 0];
            uint32_t own[3], ext[3];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = cy + ky - 1;
                const int sy = (in && yy >= 0 && yy < CHh) ? ty[yy] : -1;   // -1 across a frame/tile gap
                const uint8_t* const row = fsrc + static_cast<long long>(max(sy, 0)) * p.src_stride;
                own[ky] = 0u;
                ext[ky] = 0u;
                if (sy >= 0 && sx >= 0) {
                    const uint8_t* s = row + sx * 3;
                    own[ky] = s[0] | (static_cast<uint32_t>(s[1]) << 8) | (static_cast<uint32_t>(s[2]) << 16);
                }
                if (sy >= 0 && sxe >= 0) {
                    const uint8_t* s = row + sxe * 3;
                    ext[ky] = s[0] | (static_cast<uint32_t>(s[1]) << 8) | (static_cast<uint32_t>(s[2]) << 16);
                }
            }
            // a neighbouring lane is the neighbouring column only inside one canvas row
            const bool has_l = cx > 0, has_r = cx < CW - 1;
            uint32_t L[3], R[3];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, own[ky], 1);
                const uint32_t dn = __shfl_down_sync(0xffffffffu, own[ky], 1);
                L[ky] = (lane == 0) ? ext[ky] : (has_l ? up : 0u);
                R[ky] = (lane == 31) ? ext[ky] : (has_r ? dn : 0u);
            }
            // the 27 bytes in im2col order k = (ky*3 + kx)*3 + c, packed four to a word, then two fp16 per word
            const uint32_t q[7] = {L[0] | (own[0] << 24), (own[0] >> 8) | (R[0] << 16), (R[0] >> 16) | (L[1] << 8),
                                   own[1] | (R[1] << 24), (R[1] >> 8) | (L[2] << 16), (L[2] >> 16) | (own[2] << 8), R[2]};
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                // bytes (b0, 0x64, b1, 0x64) = the fp16 pair (1024 + b0, 1024 + b1); subtracting 1024 is exact
                const uint32_t lo = __byte_perm(q[i], 0x64646464u, 0x4140), hi = __byte_perm(q[i], 0x64646464u, 0x4342);
                const __half2 k1024 = __float2half2_rn(1024.f);
                const __half2 a = __hsub2(*reinterpret_cast<const __half2*>(&lo), k1024);
                const __half2 b = __hsub2(*reinterpret_cast<const __half2*>(&hi), k1024);
                w[2 * i] = *reinterpret_cast<const uint32_t*>(&a);
                w[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&b);
            }
            w[14] = w[15] = 0u;
            mbar_wait_relaxed<REVE_CONV0_WAIT_SLEEP>(base + kBarEmpty + 8 * stage, (use & 1) ^ 1, dbg, TAG0_EMPTY, j);
            // element (row m, 16-byte k-chunk c) at (m/8)*512 + c*128 + (m%8)*16
            const uint32_t dst = base + kOffA + stage * kTileA + (m >> 3) * 512 + (m & 7) * 16;
#pragma unroll
            for (int c = 0; c < 4; ++c) st_shared_v4(dst + c * 128, w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
            fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(base + kBarFull + 8 * stage);
        }
    } else if (warp == kMmaWarp) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {   // weights: one bulk copy (after the barrier-init __syncthreads above)
            mbar_arrive_expect_tx(base + kBarW, kWBytes0);
            bulk_load_1d(base + kOffW, p.weights, kWBytes0, base + kBarW);
        }
        __syncwarp();
        mbar_wait(base + kBarW, 0, dbg, TAG0_W);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_f16(128, 64);
        constexpr uint32_t lbo = 128u, sbo = 512u;  // verified on B200: LBO = K-adjacent core matrices, SBO = 8-row groups
        uint32_t j = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
            const uint32_t stage = j % kStagesA, use = j / kStagesA;
            const uint32_t buf = j % kAccBufs, ubuf = j / kAccBufs;
            mbar_wait(base + kBarAccEmpty + 8 * buf, (ubuf & 1) ^ 1, dbg, TAG0_ACC_EMPTY, j);
            mbar_wait(base + kBarFull + 8 * stage, use & 1, dbg, TAG0_FULL, j);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a = base + kOffA + stage * kTileA;
#pragma unroll
                for (int s = 0; s < 2; ++s)
                    umma_f16(tmem_base + buf * 64, umma_desc_nosw(a + s * 256, lbo, sbo),
                             umma_desc_nosw(base + kOffW + s * 256, lbo, sbo), idesc, s ? 1u : 0u);
                umma_commit(base + kBarEmpty + 8 * stage);
                umma_commit(base + kBarAccFull + 8 * buf);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ epilogue
        const int grp = (warp - kFirstEpiWarp) >> 2;
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        uint32_t j = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
            if ((j % kEpiGroups) != static_cast<uint32_t>(grp)) continue;
            const uint32_t stg = base + kOffOut + grp * kStageOut;
            const uint32_t buf = j % kAccBufs, ubuf = j / kAccBufs;
            const unsigned px = static_cast<unsigned>(tile) * 128u + m;
            bool keep = false;
            if (px < npx) {
                const int cy = static_cast<int>(px / static_cast<unsigned>(CW)), cx = static_cast<int>(px - static_cast<unsigned>(cy) * CW);
                keep = (tx[cx] >= 0) && (ty[cy] >= 0);
            }
            mbar_wait_relaxed<REVE_CONV0_WAIT_SLEEP>(base + kBarAccFull + 8 * buf, ubuf & 1, dbg, TAG0_ACC_FULL, j);
            tc_fence_after();
            if (lane == 0) bulk_wait_read<0>();   // this warp's previous store has read its quarter of the staging buffer out
            __syncwarp();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t acc[32];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t(&dst)[16] = *reinterpret_cast<uint32_t(*)[16]>(&acc[c * 16]);
                    tmem_ld16(tmem_lane + buf * 64 + half * 32 + c * 16, dst);
                }
                tmem_wait_ld();
                if (half == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(base + kBarAccEmpty + 8 * buf);
                }
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    uint32_t pk[4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int ch = half * 32 + c8 * 8 + jj * 2;
                        // acc/255 + bias in fp32, then PReLU on the packed fp16 pair (as in the body kernel)
                        const __half2 v = __floats2half2_rn(fmaf(__uint_as_float(acc[c8 * 8 + jj * 2]), 1.0f / 255.0f, p.bias[ch]),
                                                            fmaf(__uint_as_float(acc[c8 * 8 + jj * 2 + 1]), 1.0f / 255.0f, p.bias[ch + 1]));
                        const __half2 z = __float2half2_rn(0.f);
                        const __half2 r = __hfma2(p.slope2[ch >> 1], __hmin2_nan(v, z), __hmax2_nan(v, z));
                        pk[jj] = keep ? *reinterpret_cast<const uint32_t*>(&r) : 0u;
                    }
                    st_shared_v4(stg + m * 128 + (((half * 4 + c8) ^ (m & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&out_map, stg + q * 4096, 0, tile * 128 + q * 32);
                bulk_commit();
            }
        }
        if (lane == 0) bulk_wait<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, kAccBufs * 64);
    }
}

}  // namespace

size_t conv0_weight_blob_bytes() { return kWBytes0; }

void pack_conv0_weights(const float* w_oihw, uint16_t* blob) {
    // B operand [co = 64][k = 32] fp16, k = (ky*3 + kx)*3 + c (27 used), K-major no-swizzle core-matrix
    // layout: element (co, k) at (co/8)*512 + (k/8)*128 + (co%8)*16 + (k%8)*2 bytes.
    std::memset(blob, 0, kWBytes0);
    for (int co = 0; co < 64; ++co)
        for (int c = 0; c < 3; ++c)
            for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx) {
                    const int k = (ky * 3 + kx) * 3 + c;
                    const float v = w_oihw[((static_cast<size_t>(co) * 3 + c) * 3 + ky) * 3 + kx];
                    const size_t byte = static_cast<size_t>(co / 8) * 512 + (k / 8) * 128 + (co % 8) * 16 + (k % 8) * 2;
                    blob[byte / 2] = f32_to_f16(v);
                }
}

cudaError_t conv0_kernel_init() {
    return cudaFuncSetAttribute(conv0_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
}

cudaError_t launch_conv0(cudaStream_t st, int grid, const CUtensorMap& out_map, const Conv0Params& p) {
    conv0_umma_kernel<<<grid, kThreads, kSmem, st>>>(out_map, p);
    return cudaGetLastError();
}

}  // namespace reve
