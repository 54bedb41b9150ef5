// First convolution (3 -> 64, 3x3) + PReLU, row-streaming: the same function as conv0.cu (u8 RGB gather with the
// reflect-101 pre-pad, zero padding at tile borders, /255, canvas layout, fp16 NHWC out -- the pre-processing of the
// upscaler spawned at reference reve-shared/src/lib.rs:134-147; SURVEY.md section 2.3, K1 + K2), with a third of the
// gather work.
//
// conv0.cu builds the whole 3x3x3 neighbourhood of every pixel (K = 27) and is bound by the instructions of exactly that
// gather.  Here the canvas is walked the way the body kernel walks it: strips of 128 columns, top to bottom, and every
// INPUT row is gathered ONCE.  For input row r the producers build a [128 px][K = 9 -> 16] tile -- the pixel and its left
// and right neighbours, k = kx*3 + c -- and the MMA thread multiplies it by the three vertical taps separately:
//     D[out row r+1]  = A_r x W[ky = 0]      (first contribution: overwrite)
//     D[out row r  ] += A_r x W[ky = 1]
//     D[out row r-1] += A_r x W[ky = 2]      (last contribution: the row is complete -> epilogue)
// Three N = 64 MMAs (K = 16) instead of one N = 192 so that each carries its own accumulate flag: no TMEM zero-fill,
// no rotation of B.  A strip is exactly 128 output columns (the horizontal taps sit inside K), so every epilogue warp
// stores full 32-pixel quarters with its own TMA stores.
// A STEP is two input rows: it opens one pair of output rows and completes the previous pair (six MMAs, one A-ring
// stage, one TMEM slot of 128 columns out of a ring of four, one epilogue event).  The MMAs are tiny, so the barrier
// round trips and the scalar work around them are what the kernel's time is made of (with every work stage compiled
// out the hand-shakes alone cost half of it at one row per step, profiles/r02_notes.md section 16); two rows per step
// halve them: 0.105 -> 0.088 ms per 1080p frame in the pipeline, 0.121 for conv0.cu.
// Per pixel the producers issue 3 byte loads (+3 predicated at warp edges), 2 shuffles, 5 byte->fp16 conversions and
// 2 shared stores: ~50 instructions per lane and row against ~310 in conv0.cu.
// Warp roles as in conv0.cu (800 threads): warps 0-11 = three producer groups (steps round-robin), warp 12 = TMEM
// allocator + MMA issuer, warps 13-24 = three epilogue groups (row pairs round-robin).
#include "kernels.h"

#include <cstring>

#include "model.h"

namespace reve {

namespace {

#ifndef REVE_CONV0_PG
#define REVE_CONV0_PG 3
#endif
#ifndef REVE_CONV0_EG
#define REVE_CONV0_EG 3
#endif
constexpr int kProducerGroups = REVE_CONV0_PG;
constexpr int kProducerWarps = 4 * kProducerGroups;
constexpr int kMmaWarp = kProducerWarps;
constexpr int kFirstEpiWarp = kMmaWarp + 1;
constexpr int kEpiGroups = REVE_CONV0_EG;
constexpr int kThreads = (kFirstEpiWarp + 4 * kEpiGroups) * 32;
constexpr int kMaxTableInts = 12288;
constexpr int kStagesA = 6;              // multiple of kProducerGroups
#ifndef REVE_CONV0_AHEAD
#define REVE_CONV0_AHEAD 2
#endif
constexpr int kAhead = REVE_CONV0_AHEAD; // steps a producer group gathers before it converts the first of them
constexpr int kRows = 2;                 // input rows per step = output rows per epilogue event
constexpr int kTileA = 128 * 32;         // 4 KB: 128 px x 16 k x fp16, one input row
constexpr int kStageA = kRows * kTileA;  // one step's A operands
constexpr int kSlots = 4;                // TMEM ring of row pairs (128 columns each): two being accumulated, two being drained
constexpr int kTmemCols = 512;
constexpr int kWTap = 64 * 32;           // 2 KB: 64 co x 16 k x fp16, one vertical tap
constexpr int kWBytes = 3 * kWTap;
constexpr int kRowOut = 128 * 128;       // 16 KB: one output row of the strip
constexpr int kStageOut = kRows * kRowOut;   // output staging per epilogue group: one event

constexpr int kBarW = 0;
constexpr int kBarFull = 8;                            // [kStagesA]
constexpr int kBarEmpty = kBarFull + 8 * kStagesA;     // [kStagesA]
constexpr int kBarAccFull = kBarEmpty + 8 * kStagesA;  // [kSlots]
constexpr int kBarAccEmpty = kBarAccFull + 8 * kSlots; // [kSlots]
constexpr int kTmemPtr = 512;
constexpr int kCtrl = 1024;
constexpr int kOffW = kCtrl;
constexpr int kOffA = kOffW + kWBytes;
constexpr int kOffOut = kOffA + kStagesA * kStageA;
constexpr int kOffTab = kOffOut + kEpiGroups * kStageOut;
constexpr int kSmem = 1024 + kOffTab + kMaxTableInts * 4;
static_assert(kOffA % 1024 == 0 && kOffOut % 1024 == 0, "staging buffers are swizzled in 1 KB atoms");

enum : uint32_t { TAGR_W = 21, TAGR_EMPTY = 22, TAGR_FULL = 23, TAGR_ACC_EMPTY = 24, TAGR_ACC_FULL = 25 };

__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo >> 4) << 16;
    d |= static_cast<uint64_t>(sbo >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// bytes (b0, b1) of q -> the fp16 pair (b0, b1), exactly: (b, 0x64) is the fp16 1024 + b
__device__ __forceinline__ uint32_t bytes_to_h2(uint32_t q, uint32_t sel) {
    const uint32_t v = __byte_perm(q, 0x64646464u, sel);
    const __half2 r = __hsub2(*reinterpret_cast<const __half2*>(&v), __float2half2_rn(1024.f));
    return *reinterpret_cast<const uint32_t*>(&r);
}

// A CTA owns a contiguous range of (strip, row) pairs in strip-major order and walks it as segments: runs of
// consecutive output rows [ya, ya + n) of one strip.  All roles walk the same sequence.
struct Walk {
    int p, hi, rows;
    __device__ bool next(int& strip, int& ya, int& n) {
        if (p >= hi) return false;
        strip = p / rows;
        ya = p - strip * rows;
        n = min(rows - ya, hi - p);
        p += n;
        return true;
    }
};

__global__ void __launch_bounds__(kThreads, 1)
conv0_rows_kernel(const __grid_constant__ CUtensorMap out_map_q, const __grid_constant__ Conv0Params p) {
    // out_map_q: the output canvas [rows][columns][64 ch] with a 32-pixel box
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* const base_ptr = smem_raw + (base - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    DebugBlock* const dbg = p.dbg;

    if (threadIdx.x == 0) {
        mbar_init(base + kBarW, 1);
        for (int s = 0; s < kStagesA; ++s) {
            mbar_init(base + kBarFull + 8 * s, 4);
            mbar_init(base + kBarEmpty + 8 * s, 1);
        }
        for (int s = 0; s < kSlots; ++s) {
            mbar_init(base + kBarAccFull + 8 * s, 1);
            mbar_init(base + kBarAccEmpty + 8 * s, 4);
        }
        fence_mbar_init();
    }
    if (warp == kMmaWarp) {
        tmem_alloc(base + kTmemPtr, kTmemCols);
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
    if (warp == kFirstEpiWarp && lane == 0) prefetch_tmap(&out_map_q);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + kTmemPtr);

    const int n_strips = (CW + 127) / 128;
    const long long total = static_cast<long long>(n_strips) * CHh;   // fits an int: activation canvases are capped at 6 GB (api.cu), i.e. 47 M pixels
    Walk walk{static_cast<int>(total * blockIdx.x / gridDim.x), static_cast<int>(total * (blockIdx.x + 1) / gridDim.x), CHh};
    int strip, ya, n;
    // A segment of n output rows is ceil(n/2) PAIRS of rows; step s (0 .. pairs) takes input rows ya+2s-1 and ya+2s,
    // opens pair s and completes pair s-1.  (An odd n leaves the second row of the last pair outside the segment: its
    // MMAs are not issued and the epilogue does not store it.)

    if (warp < kProducerWarps) {
        // ------------------------------------------------------------------ producers: two input rows per step
        const int pg = warp >> 2;
        const int m = (warp & 3) * 32 + lane;
        const bool edge = (lane == 0) || (lane == 31);
        uint32_t j = 0;   // step counter of the CTA
        while (walk.next(strip, ya, n)) {
            const int steps = (n + 1) / 2 + 1;
            const int cx = strip * 128 + m;
            const int xe = (lane == 0) ? cx - 1 : cx + 1;   // the column beyond the warp, fetched by lanes 0 / 31 themselves
            const int sx = (cx < CW) ? tx[cx] : -1;
            const int sxe = (edge && xe >= 0 && xe < CW) ? tx[xe] : -1;
            // This group's steps of the segment are t0, t0 + 3, ...; kAhead of them are gathered before the first is
            // converted and stored, so that their global loads are in flight together.
            const int t0 = static_cast<int>((pg + kProducerGroups - (j % kProducerGroups)) % kProducerGroups);
            for (int tb = t0; tb < steps; tb += kProducerGroups * kAhead) {
                uint32_t own[kAhead][kRows], ext[kAhead][kRows];
#pragma unroll
                for (int u = 0; u < kAhead; ++u) {
#pragma unroll
                    for (int h = 0; h < kRows; ++h) {
                        const int t = tb + u * kProducerGroups;
                        const int r = ya + 2 * t - 1 + h;
                        // -1: beyond the segment's halo, outside the canvas or a gap row -> a row of zeros
                        const int sy = (t < steps && r <= ya + n && r >= 0 && r < CHh) ? ty[r] : -1;
                        own[u][h] = 0u;
                        ext[u][h] = 0u;
                        if (sy >= 0) {
                            const uint8_t* const row = p.src[max(tf[r], 0)] + static_cast<long long>(sy) * p.src_stride;
                            if (sx >= 0) {
                                const uint8_t* s = row + sx * 3;
                                own[u][h] = s[0] | (static_cast<uint32_t>(s[1]) << 8) | (static_cast<uint32_t>(s[2]) << 16);
                            }
                            if (sxe >= 0) {
                                const uint8_t* s = row + sxe * 3;
                                ext[u][h] = s[0] | (static_cast<uint32_t>(s[1]) << 8) | (static_cast<uint32_t>(s[2]) << 16);
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < kAhead; ++u) {
                    const int t = tb + u * kProducerGroups;
                    if (t >= steps) break;
                    const uint32_t js = j + t;
                    const uint32_t stage = js % kStagesA, use = js / kStagesA;
                    uint32_t w[kRows][5];
#pragma unroll
                    for (int h = 0; h < kRows; ++h) {
                        const uint32_t up = __shfl_up_sync(0xffffffffu, own[u][h], 1);
                        const uint32_t dn = __shfl_down_sync(0xffffffffu, own[u][h], 1);
                        const uint32_t L = (lane == 0) ? ext[u][h] : up;
                        const uint32_t R = (lane == 31) ? ext[u][h] : dn;
                        // the 9 bytes in order k = kx*3 + c
                        const uint32_t q0 = L | (own[u][h] << 24), q1 = (own[u][h] >> 8) | (R << 16), q2 = R >> 16;
                        w[h][0] = bytes_to_h2(q0, 0x4140);
                        w[h][1] = bytes_to_h2(q0, 0x4342);
                        w[h][2] = bytes_to_h2(q1, 0x4140);
                        w[h][3] = bytes_to_h2(q1, 0x4342);
                        w[h][4] = bytes_to_h2(q2, 0x4140);
                    }
                    mbar_wait_relaxed<0>(base + kBarEmpty + 8 * stage, (use & 1) ^ 1, dbg, TAGR_EMPTY, js);
                    // element (row m, 16-byte k-chunk c) of input row h at h*4096 + (m/8)*256 + c*128 + (m%8)*16
                    const uint32_t dst = base + kOffA + stage * kStageA + (m >> 3) * 256 + (m & 7) * 16;
#pragma unroll
                    for (int h = 0; h < kRows; ++h) {
                        sts_v4(dst + h * kTileA, w[h][0], w[h][1], w[h][2], w[h][3]);
                        sts_v4(dst + h * kTileA + 128, w[h][4], 0u, 0u, 0u);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(base + kBarFull + 8 * stage);
                }
            }
            j += steps;
        }
    } else if (warp == kMmaWarp) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            mbar_arrive_expect_tx(base + kBarW, kWBytes);
            bulk_load_1d(base + kOffW, p.weights_rows, kWBytes, base + kBarW);
        }
        __syncwarp();
        mbar_wait(base + kBarW, 0, dbg, TAGR_W);
        tc_fence_after();
        // One thread runs the whole loop.  The MMAs of a step are tiny, so the scalar work and the barrier round trips
        // around them are the critical path of the kernel (measured: with every work stage compiled out the hand-shakes
        // alone cost half the kernel's time at one row per step): two rows per step halve them; ring positions are
        // wrapping counters and phase bits, descriptors are built once and advanced by adding to their low word.
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, 64);
            constexpr uint32_t lbo = 128u, sbo = 256u;
            const uint64_t b0 = desc_nosw(base + kOffW, lbo, sbo), b1 = desc_nosw(base + kOffW + kWTap, lbo, sbo),
                           b2 = desc_nosw(base + kOffW + 2 * kWTap, lbo, sbo);
            const uint64_t a_first = desc_nosw(base + kOffA, lbo, sbo);
            uint64_t a0 = a_first;
            uint32_t stage = 0, full_phase = 0;      // A ring position, parity of its current lap
            uint32_t slot = 0, empty_phase = 1;      // slot of the pair the next step opens, parity to wait for on its lap
            uint32_t d_open = tmem_base, d_done = 0, f_done = 0;   // accumulators of the pair being opened / completed, its AccFull barrier
            while (walk.next(strip, ya, n)) {
                const int pairs = (n + 1) / 2;
                for (int s = 0; s <= pairs; ++s) {
                    const bool open0 = 2 * s < n, open1 = 2 * s + 1 < n;       // rows of pair s inside the segment
                    const bool done1 = s >= 1 && 2 * s - 1 < n;                // second row of pair s-1 (its first always is)
                    if (open0) mbar_wait(base + kBarAccEmpty + 8 * slot, empty_phase, dbg, TAGR_ACC_EMPTY, slot);
                    mbar_wait(base + kBarFull + 8 * stage, full_phase, dbg, TAGR_FULL, stage);
                    tc_fence_after();
                    const uint64_t a1 = a0 + (kTileA >> 4);
                    // input row ya+2s-1: top tap of pair s row 0, middle tap of pair s-1 row 1, bottom tap of pair s-1 row 0
                    if (open0) umma_f16(d_open, a0, b0, idesc, 0u);
                    if (done1) umma_f16(d_done + 64, a0, b1, idesc, 1u);
                    if (s >= 1) umma_f16(d_done, a0, b2, idesc, 1u);
                    // input row ya+2s: top tap of pair s row 1, middle tap of pair s row 0, bottom tap of pair s-1 row 1
                    if (open1) umma_f16(d_open + 64, a1, b0, idesc, 0u);
                    if (open0) umma_f16(d_open, a1, b1, idesc, 1u);
                    if (done1) umma_f16(d_done + 64, a1, b2, idesc, 1u);
                    umma_commit(base + kBarEmpty + 8 * stage);
                    if (s >= 1) umma_commit(f_done);
                    d_done = d_open;
                    f_done = base + kBarAccFull + 8 * slot;
                    if (open0) {
                        if (++slot == kSlots) { slot = 0; empty_phase ^= 1u; }
                        d_open = tmem_base + slot * (64 * kRows);
                    }
                    if (++stage == kStagesA) { stage = 0; full_phase ^= 1u; a0 = a_first; } else { a0 += kStageA >> 4; }
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue: one pair of output rows per event
        const int grp = (warp - kFirstEpiWarp) >> 2;
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const uint32_t stg = base + kOffOut + grp * kStageOut;
        uint32_t e = 0;   // events of the CTA
        while (walk.next(strip, ya, n)) {
            const int cx = strip * 128 + m;
            const bool col_ok = (cx < CW) && (tx[cx] >= 0);
            const int pairs = (n + 1) / 2;
            for (int i = 0; i < pairs; ++i, ++e) {
                if ((e % kEpiGroups) != static_cast<uint32_t>(grp)) continue;
                const uint32_t slot = e % kSlots, uslot = e / kSlots;
                const int y0 = ya + 2 * i;
                const int rows = (2 * i + 1 < n) ? 2 : 1;
                mbar_wait_relaxed<0>(base + kBarAccFull + 8 * slot, uslot & 1, dbg, TAGR_ACC_FULL, e);
                tc_fence_after();
                if (lane == 0) bulk_wait_read<0>();   // this warp's previous stores have read its quarters out
                __syncwarp();
                for (int h = 0; h < rows; ++h) {
                    const bool keep = col_ok && (ty[y0 + h] >= 0);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t acc[32];
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            uint32_t(&dst)[16] = *reinterpret_cast<uint32_t(*)[16]>(&acc[c * 16]);
                            tmem_ld16(tmem_lane + slot * (64 * kRows) + h * 64 + half * 32 + c * 16, dst);
                        }
                        tmem_wait_ld();
                        if (half == 1 && h == rows - 1) {   // the slot is drained
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(base + kBarAccEmpty + 8 * slot);
                        }
#pragma unroll
                        for (int c8 = 0; c8 < 4; ++c8) {
                            uint32_t pk[4];
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const int ch = half * 32 + c8 * 8 + jj * 2;
                                const __half2 v = __floats2half2_rn(fmaf(__uint_as_float(acc[c8 * 8 + jj * 2]), 1.0f / 255.0f, p.bias[ch]),
                                                                    fmaf(__uint_as_float(acc[c8 * 8 + jj * 2 + 1]), 1.0f / 255.0f, p.bias[ch + 1]));
                                const __half2 z = __float2half2_rn(0.f);
                                const __half2 r = __hfma2(p.slope2[ch >> 1], __hmin2_nan(v, z), __hmax2_nan(v, z));
                                pk[jj] = keep ? *reinterpret_cast<const uint32_t*>(&r) : 0u;
                            }
                            sts_v4(stg + h * kRowOut + m * 128 + (((half * 4 + c8) ^ (m & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
                        }
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    for (int h = 0; h < rows; ++h) tma_store_3d(&out_map_q, stg + h * kRowOut + q * 4096, 0, strip * 128 + q * 32, y0 + h);
                    bulk_commit();
                }
            }
        }
        if (lane == 0) bulk_wait<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace

size_t conv0_rows_weight_blob_bytes() { return kWBytes; }

void pack_conv0_rows_weights(const float* w_oihw, uint16_t* blob) {
    // three B operands [co = 64][k = 16] fp16, one per vertical tap ky; k = kx*3 + c (9 used), K-major no-swizzle
    // core-matrix layout: element (co, k) at ky*2048 + (co/8)*256 + (k/8)*128 + (co%8)*16 + (k%8)*2 bytes.
    std::memset(blob, 0, kWBytes);
    for (int co = 0; co < 64; ++co)
        for (int c = 0; c < 3; ++c)
            for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx) {
                    const int k = kx * 3 + c;
                    const float v = w_oihw[((static_cast<size_t>(co) * 3 + c) * 3 + ky) * 3 + kx];
                    const size_t byte = static_cast<size_t>(ky) * kWTap + (co / 8) * 256 + (k / 8) * 128 + (co % 8) * 16 + (k % 8) * 2;
                    blob[byte / 2] = f32_to_f16(v);
                }
}

cudaError_t conv0_rows_kernel_init() {
    return cudaFuncSetAttribute(conv0_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
}

cudaError_t launch_conv0_rows(cudaStream_t st, int grid, const CUtensorMap& out_map_q, const Conv0Params& p) {
    conv0_rows_kernel<<<grid, kThreads, kSmem, st>>>(out_map_q, p);
    return cudaGetLastError();
}

}  // namespace reve
