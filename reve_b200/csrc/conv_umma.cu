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
//   * Alternate layers sweep their ranges in opposite directions (`reverse`): the rows a CTA wrote
//     last in layer l are the rows it reads first in layer l+1, while they are still in L2.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (warp-
// uniform code, one elected lane issues), warps 2..9 = two epilogue groups that take alternate
// output rows (TMEM -> registers -> bias/PReLU -> fp16 -> swizzled smem -> TMA store, or for the tail:
// bias -> PixelShuffle + residual -> u8 -> global).
#include "kernels.h"

#include <algorithm>
#include <cstring>

#include "model.h"

namespace reve {

namespace {

constexpr int kRowBytes = kBoxPx * 128;  // 16 KB: 128 px * 64 ch * fp16
constexpr int kCtrlBytes = 2048;
constexpr int kGuard = 1024;

__host__ __device__ constexpr int w_bytes(int ng) { return 3 * 3 * ng * 128; }
__host__ __device__ constexpr int tmem_cols(int ng) { return ng * 8 <= 128 ? 128 : (ng * 8 <= 256 ? 256 : 512); }

// control block offsets (from the 1024-aligned base)
static_assert(kStages % 2 == 0, "ring slots are released in pairs");
constexpr int kBarW = 0;
constexpr int kBarAFull = 8;
constexpr int kBarAEmpty = kBarAFull + 8 * kStages;
constexpr int kBarAccFull = kBarAEmpty + 8 * kStages;
constexpr int kBarAccEmpty = kBarAccFull + 8 * 8;
constexpr int kTmemPtr = 512;
constexpr int kOffBias = 1024;   // 64 floats
constexpr int kOffSlope = 1280;  // 64 floats

enum : uint32_t { TAG_W = 1, TAG_A_EMPTY = 2, TAG_A_FULL = 3, TAG_ACC_EMPTY = 4, TAG_ACC_FULL = 5 };

// Range of (virtual) strip-rows of this CTA, cut into per-strip segments.  In reverse mode the
// virtual index runs backwards over the physical one (strip' = n_strips-1-strip, y' = CH-1-y) and
// the CTA takes the mirrored block, i.e. the same physical region as in forward mode.
struct SegIter {
    long long lo, hi;
    int ch;
    __device__ SegIter(const ConvParams& p) : ch(p.canvas_h) {
        const unsigned b = p.reverse ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
        lo = static_cast<long long>(b) * p.total_rows / gridDim.x;
        hi = static_cast<long long>(b + 1) * p.total_rows / gridDim.x;
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

__device__ __forceinline__ uint64_t mk_desc(uint32_t hi, uint32_t lo) {
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int NG, bool TAIL>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_umma_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                    const __grid_constant__ ConvParams p) {
    constexpr int kWBytes = w_bytes(NG);
    constexpr int kTmemCols = tmem_cols(NG);
    constexpr int kOffW = kCtrlBytes;
    constexpr int kOffRing = kOffW + kWBytes + kGuard;
    constexpr int kOffStage = kOffRing + kStages * kRowBytes + kGuard;  // body: 2 x 16 KB output staging (one per group)

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
            mbar_init(base + kBarAccEmpty + 8 * s, 4);  // one arrive per warp of the draining group
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
    if (threadIdx.x >= 64 && threadIdx.x < 128) {
        // bias / PReLU slopes: shared-memory copies, read back as broadcast LDS.128 in the epilogue
        // (constant-bank operands turn into long-latency LDCU loads on sm_100)
        reinterpret_cast<float*>(base_ptr + kOffBias)[threadIdx.x - 64] = p.bias[threadIdx.x - 64];
        reinterpret_cast<float*>(base_ptr + kOffSlope)[threadIdx.x - 64] = p.slope[threadIdx.x - 64];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + kTmemPtr);

    const int CH = p.canvas_h;
    const bool rev = p.reverse != 0;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            mbar_arrive_expect_tx(base + kBarW, kWBytes);
            bulk_load_1d(base + kOffW, p.weights, kWBytes, base + kBarW);
            const uint64_t policy = kPolicyEvictFirst;  // activations are read once per layer
            SegIter it(p);
            int strip, ya, yb;
            uint32_t i = 0;
            while (it.next(strip, ya, yb)) {
                const int x0 = (rev ? p.n_strips - 1 - strip : strip) * kStripPx;
                const int y_lo = max(ya - 1, 0), y_hi = min(yb + 1, CH - 1);
                for (int y = y_lo; y <= y_hi; ++y, ++i) {
                    const uint32_t stage = i % kStages;
                    if ((i & 1u) == 0) {  // ring slots are handed back two at a time (one commit per two rows)
                        const uint32_t pair = (i >> 1) % (kStages / 2), usep = (i >> 1) / (kStages / 2);
                        mbar_wait(base + kBarAEmpty + 8 * pair, (usep & 1) ^ 1, dbg, TAG_A_EMPTY, i);
                    }
                    mbar_arrive_expect_tx(base + kBarAFull + 8 * stage, kRowBytes);
                    tma_load_3d_hint(base + kOffRing + stage * kRowBytes, &in_map, base + kBarAFull + 8 * stage, 0,
                                     x0 - 1, rev ? CH - 1 - y : y, policy);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // Warp-uniform control flow; one elected lane issues.  The issuing thread is the scarce
        // resource (13+ MMAs, 2 waits and 2 commits per ~1150 tensor-pipe cycles), so interior rows
        // take a straight-line path whose descriptors differ from per-row bases by constants only.
        mbar_wait(base + kBarW, 0, dbg, TAG_W);
        tc_fence_after();
        const uint32_t idesc1 = umma_idesc_f16(128, NG);
        const uint32_t idesc2 = umma_idesc_f16(128, 2 * NG);
        const uint32_t idesc3 = umma_idesc_f16(128, 3 * NG);
        const uint64_t proto = umma_desc_sw128(0, 0);
        const uint32_t desc_hi = static_cast<uint32_t>(proto >> 32);
        const uint32_t lo_flags = static_cast<uint32_t>(proto);              // LBO field
        const uint32_t w_lo = lo_flags | ((base + kOffW) >> 4);
        const uint32_t ring_lo = lo_flags | ((base + kOffRing) >> 4);
        constexpr uint32_t kG = NG * 8;        // one group of B rows, in 16-byte units
        constexpr uint32_t kDx = 3 * NG * 8;   // one dx block of B
                // Issue helpers (called by the elected lane only).  Part 1 = first K-step (the fresh group
        // overwrites, the others accumulate) plus three more K-steps; part 2 = the other eight.
        // MMAs that read the same A tile back to back keep it in the tensor core's collector buffer
        // (FILL ... LASTUSE) instead of re-reading shared memory: that makes the split MMAs of the
        // ring-wrap rows (N=128 + N=64) cost the same tensor time as one N=192 MMA.
        auto issue_part1 = [&](uint32_t a_lo, int s0, uint32_t w_lo) {
            const uint64_t a0 = mk_desc(desc_hi, a_lo - 8);
            const uint32_t d = tmem_base + s0 * NG;
            if (s0 <= 5) {
                umma_f16_a<ACollector::FILL>(d, a0, mk_desc(desc_hi, w_lo), idesc1, 0u);
                umma_f16_a<ACollector::LASTUSE>(d + NG, a0, mk_desc(desc_hi, w_lo + kG), idesc2, 1u);
#pragma unroll
                for (int dxk = 1; dxk < 4; ++dxk)
                    umma_f16(d, mk_desc(desc_hi, a_lo - 8 + dxk * 2), mk_desc(desc_hi, w_lo + dxk * 2), idesc3, 1u);
            } else if (s0 == 6) {
                umma_f16_a<ACollector::FILL>(d, a0, mk_desc(desc_hi, w_lo), idesc1, 0u);
                umma_f16_a<ACollector::USE>(d + NG, a0, mk_desc(desc_hi, w_lo + kG), idesc1, 1u);
                umma_f16_a<ACollector::LASTUSE>(tmem_base, a0, mk_desc(desc_hi, w_lo + 2 * kG), idesc1, 1u);
#pragma unroll
                for (int dxk = 1; dxk < 4; ++dxk) {
                    const uint64_t ad = mk_desc(desc_hi, a_lo - 8 + dxk * 2);
                    umma_f16_a<ACollector::FILL>(d, ad, mk_desc(desc_hi, w_lo + dxk * 2), idesc2, 1u);
                    umma_f16_a<ACollector::LASTUSE>(tmem_base, ad, mk_desc(desc_hi, w_lo + dxk * 2 + 2 * kG), idesc1, 1u);
                }
            } else {
                umma_f16_a<ACollector::FILL>(d, a0, mk_desc(desc_hi, w_lo), idesc1, 0u);
                umma_f16_a<ACollector::LASTUSE>(tmem_base, a0, mk_desc(desc_hi, w_lo + kG), idesc2, 1u);
#pragma unroll
                for (int dxk = 1; dxk < 4; ++dxk) {
                    const uint64_t ad = mk_desc(desc_hi, a_lo - 8 + dxk * 2);
                    umma_f16_a<ACollector::FILL>(d, ad, mk_desc(desc_hi, w_lo + dxk * 2), idesc1, 1u);
                    umma_f16_a<ACollector::LASTUSE>(tmem_base, ad, mk_desc(desc_hi, w_lo + dxk * 2 + kG), idesc2, 1u);
                }
            }
        };
        auto issue_part2 = [&](uint32_t a_lo, int s0, uint32_t w_lo) {
            const uint32_t d = tmem_base + s0 * NG;
            if (s0 <= 5) {
#pragma unroll
                for (int dxk = 4; dxk < 12; ++dxk) {
                    const int dx = dxk >> 2, k = dxk & 3;
                    umma_f16(d, mk_desc(desc_hi, a_lo + (dx - 1) * 8 + k * 2), mk_desc(desc_hi, w_lo + dx * kDx + k * 2),
                             idesc3, 1u);
                }
            } else if (s0 == 6) {
#pragma unroll
                for (int dxk = 4; dxk < 12; ++dxk) {
                    const int dx = dxk >> 2, k = dxk & 3;
                    const uint64_t ad = mk_desc(desc_hi, a_lo + (dx - 1) * 8 + k * 2);
                    umma_f16_a<ACollector::FILL>(d, ad, mk_desc(desc_hi, w_lo + dx * kDx + k * 2), idesc2, 1u);
                    umma_f16_a<ACollector::LASTUSE>(tmem_base, ad, mk_desc(desc_hi, w_lo + dx * kDx + k * 2 + 2 * kG), idesc1, 1u);
                }
            } else {
#pragma unroll
                for (int dxk = 4; dxk < 12; ++dxk) {
                    const int dx = dxk >> 2, k = dxk & 3;
                    const uint64_t ad = mk_desc(desc_hi, a_lo + (dx - 1) * 8 + k * 2);
                    umma_f16_a<ACollector::FILL>(d, ad, mk_desc(desc_hi, w_lo + dx * kDx + k * 2), idesc1, 1u);
                    umma_f16_a<ACollector::LASTUSE>(tmem_base, ad, mk_desc(desc_hi, w_lo + dx * kDx + k * 2 + kG), idesc2, 1u);
                }
            }
        };
        // Edge rows of a segment (first / last two input rows): one MMA per group and K-step.
        auto edge_row = [&](int y, int ya, int yb, uint32_t i, int t0) {
            const uint32_t stage = i % kStages, use = i / kStages;
            const uint32_t a_lo = ring_lo + stage * (kRowBytes >> 4);
            const int s0 = (-t0) & 7;
            // group g (0..2) = vertical tap: input row y feeds output row y + 1 - g
            const int g_lo = (y + 1 <= yb) ? 0 : ((y <= yb) ? 1 : 2);
            const int g_hi = (y - 1 >= ya) ? 2 : ((y >= ya) ? 1 : 0);
            // groups whose output row receives its first contribution from this input row
            const int fresh_hi = (y == 0) ? g_hi : ((g_lo == 0) ? 0 : g_lo - 1);
            for (int g = g_lo; g <= fresh_hi; ++g) {
                const int tg = t0 - g;
                mbar_wait(base + kBarAccEmpty + 8 * ((s0 + g) & 7), ((tg >> 3) & 1) ^ 1, dbg, TAG_ACC_EMPTY, tg);
            }
            mbar_wait(base + kBarAFull + 8 * stage, use & 1, dbg, TAG_A_FULL, i);
            tc_fence_after();
            if (elect_one()) {
                for (int dxk = 0; dxk < 12; ++dxk) {
                    const int dx = dxk >> 2, k = dxk & 3;
                    const uint64_t ad = mk_desc(desc_hi, a_lo + (dx - 1) * 8 + k * 2);
                    for (int g = g_lo; g <= g_hi; ++g)
                        umma_f16(tmem_base + ((s0 + g) & 7) * NG, ad, mk_desc(desc_hi, w_lo + dx * kDx + k * 2 + g * kG),
                                 idesc1, (dxk == 0 && g <= fresh_hi) ? 0u : 1u);
                }
                if (i & 1u) umma_commit(base + kBarAEmpty + 8 * ((i >> 1) % (kStages / 2)));  // rows i-1, i retired
                if (g_hi == 2) umma_commit(base + kBarAccFull + 8 * ((s0 + 2) & 7));  // row y-1 complete
                if (y == CH - 1 && yb == CH - 1)                                       // bottom edge: row y too
                    umma_commit(base + kBarAccFull + 8 * ((s0 + 1) & 7));
            }
            __syncwarp();
        };

        SegIter it(p);
        int strip, ya, yb;
        uint32_t i = 0;   // input rows issued so far (ring stage = i mod kStages)
        int t_base = 0;   // output rows of earlier segments
        while (it.next(strip, ya, yb)) {
            const int y_lo = max(ya - 1, 0), y_hi = min(yb + 1, CH - 1);
            int y = y_lo;
            for (; y <= min(ya, y_hi); ++y, ++i) edge_row(y, ya, yb, i, t_base + (y + 1 - ya));
            // ---- interior rows ya+1 .. yb-1: groups 0..2 enabled, only group 0 fresh, row y-1 completes.
            // The tensor pipe's instruction queue is shallow: every cycle the issuing thread spends on
            // anything else between two rows is a bubble.  So the barriers of row y+1 are polled in the
            // middle of row y, and the loop carries only (i, t0).
            int n_int = yb - ya - 1;
            if (n_int > 0) {
                int t0 = t_base + 2;  // y = ya + 1
                {
                    const int s0 = (-t0) & 7;
                    mbar_wait(base + kBarAccEmpty + 8 * s0, ((t0 >> 3) & 1) ^ 1, dbg, TAG_ACC_EMPTY, t0);
                    mbar_wait(base + kBarAFull + 8 * (i % kStages), (i / kStages) & 1, dbg, TAG_A_FULL, i);
                    tc_fence_after();
                }
                const bool leader = elect_one();
                for (; n_int > 0; --n_int, ++i, ++t0, ++y) {
                    const uint32_t stage = i % kStages;
                    const uint32_t a_lo = ring_lo + stage * (kRowBytes >> 4);
                    const int s0 = (-t0) & 7;
                    long long* const tr = (p.trace && blockIdx.x == 0 && i < 256) ? p.trace + i * 4 : nullptr;
                    if (tr && lane == 0) { tr[0] = clock64(); tr[3] = s0 << 4; }
                    // Opaque per-row copy of the weight descriptor base: keeps the compiler from hoisting 36
                    // loop-invariant B descriptors into vector registers (two R2UR moves per MMA); the
                    // descriptors become uniform-datapath adds on one per-row value instead.
                    uint32_t w_row = w_lo;
                    asm volatile("" : "+r"(w_row));
                    if (leader) issue_part1(a_lo, s0, w_row);
                    // look ahead: barriers of the next interior row
                    bool ok = true;
                    uint32_t bar_e = 0, par_e = 0, bar_f = 0, par_f = 0;
                    if (n_int > 1) {
                        bar_e = base + kBarAccEmpty + 8 * ((s0 + 7) & 7);
                        par_e = (((t0 + 1) >> 3) & 1) ^ 1;
                        bar_f = base + kBarAFull + 8 * ((i + 1) % kStages);
                        par_f = ((i + 1) / kStages) & 1;
                        const bool ok_e = mbar_try_wait(bar_e, par_e);
                        const bool ok_f = mbar_try_wait(bar_f, par_f);
                        ok = ok_e && ok_f;
                    }
                    if (tr && lane == 0) tr[1] = clock64();
                    if (leader) {
                        issue_part2(a_lo, s0, w_row);
                        if (i & 1u) umma_commit(base + kBarAEmpty + 8 * ((i >> 1) % (kStages / 2)));
                        umma_commit(base + kBarAccFull + 8 * ((s0 + 2) & 7));
                    }
                    if (!ok) {
                        if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[2040] += 1;   // look-ahead misses
                        mbar_wait(bar_e, par_e, dbg, TAG_ACC_EMPTY, t0 + 1);
                        mbar_wait(bar_f, par_f, dbg, TAG_A_FULL, i + 1);
                    }
                    tc_fence_after();
                    if (tr && lane == 0) tr[2] = clock64();
                    if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[2041] += 1;       // interior rows
                }
                __syncwarp();
            }
            for (; y <= y_hi; ++y, ++i) edge_row(y, ya, yb, i, t_base + (y + 1 - ya));
            t_base += yb - ya + 1;
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps
        const int grp = (warp - 2) >> 2;   // group 0 / 1 take even / odd output-row sequence numbers
        const int q = warp & 3;            // TMEM lane quarter this warp may access
        const int m = q * 32 + lane;       // M row = pixel index inside the 128-px box
        const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        SegIter it(p);
        int strip, ya, yb;
        int t = 0;
        while (it.next(strip, ya, yb)) {
            const int x0 = (rev ? p.n_strips - 1 - strip : strip) * kStripPx;
            const int cx = x0 - 1 + m;
            const bool inside = (m >= 1) && (m <= kStripPx) && (cx < p.canvas_w);
            const bool colok = inside && (p.colflag[cx] != 0);
            int ox = -1, sx = 0;
            if (TAIL && inside) {
                ox = p.out_x[cx];
                sx = p.src_x[cx];
            }
            for (int r = ya; r <= yb; ++r, ++t) {
                if ((t & 1) != grp) continue;
                const int pr = rev ? CH - 1 - r : r;  // physical canvas row
                const int s = (-t) & 7;
                long long* const tr = (p.trace && blockIdx.x == 0 && t < 256 && q == 0 && lane == 0)
                                          ? p.trace + 1024 + t * 4 : nullptr;
                if (tr) tr[0] = clock64();
                mbar_wait(base + kBarAccFull + 8 * s, (t >> 3) & 1, dbg, TAG_ACC_FULL, t);
                if (tr) tr[1] = clock64();
                tc_fence_after();
                uint32_t acc[NG];
#pragma unroll
                for (int c = 0; c < NG / 16; ++c) {
                    uint32_t(&dst)[16] = *reinterpret_cast<uint32_t(*)[16]>(&acc[c * 16]);
                    tmem_ld16(tmem_lane + s * NG + c * 16, dst);
                }
                tmem_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(base + kBarAccEmpty + 8 * s);
                if (tr) tr[2] = clock64();

                if constexpr (!TAIL) {
                    const bool keep = colok && (p.rowflag[pr] != 0);
                    const uint32_t stg = base + kOffStage + grp * kRowBytes;
                    const bool gleader = (q == 0 && lane == 0);
                    if (gleader) bulk_wait_read<0>();   // this group's previous row has left the staging buffer
                    named_bar_sync(1 + grp, 128);
                    if (m >= 1 && m <= kStripPx) {
                        const int row = m - 1;
                        const uint32_t rbase = stg + row * 128;
                        if (keep) {
#pragma unroll
                            for (int c8 = 0; c8 < 8; ++c8) {
                                uint32_t pk[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const int ch = c8 * 8 + j * 2;
                                    // bias in fp32, then PReLU on the packed fp16 pair: max(v,0) + a*min(v,0).
                                    // Positive values are bit-identical to the fp32 formulation; negative ones
                                    // round twice (<= 1 ulp of fp16 instead of 0.5).
                                    const __half2 v = __floats2half2_rn(__uint_as_float(acc[ch]) + p.bias[ch],
                                                                        __uint_as_float(acc[ch + 1]) + p.bias[ch + 1]);
                                    const __half2 z = __float2half2_rn(0.f);
                                    const __half2 r = __hfma2(p.slope2[ch >> 1], __hmin2(v, z), __hmax2(v, z));
                                    pk[j] = *reinterpret_cast<const uint32_t*>(&r);
                                }
                                st_shared_v4(rbase + ((c8 ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
                            }
                        } else {   // gap pixel (between tiles / frames): must read as zero in the next layer
#pragma unroll
                            for (int c8 = 0; c8 < 8; ++c8) st_shared_v4(rbase + (c8 << 4), 0u, 0u, 0u, 0u);
                        }
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(1 + grp, 128);
                    if (gleader) {
                        tma_store_3d(&out_map, stg, 0, x0, pr);
                        bulk_commit();
                    }
                } else {
                    constexpr int S = (NG == 16) ? 2 : ((NG == 32) ? 3 : 4);
                    const int oy = p.out_y[pr];
                    if (ox >= 0 && oy >= 0) {
                        const int fr = p.row_frame[pr];
                        const uint8_t* sp = p.src[fr] + static_cast<long long>(p.src_y[pr]) * p.src_stride + sx * 3;
                        const float xin[3] = {static_cast<float>(sp[0]), static_cast<float>(sp[1]),
                                              static_cast<float>(sp[2])};
                        // S*3 consecutive bytes per output row: packed into 16-/32-bit stores when aligned
                        const bool wide = (S != 3) && (((reinterpret_cast<uintptr_t>(p.dst[fr]) | static_cast<uintptr_t>(p.dst_stride)) & 3) == 0);
#pragma unroll
                        for (int i = 0; i < S; ++i) {
                            uint8_t* dp = p.dst[fr] + static_cast<long long>(oy * S + i) * p.dst_stride +
                                          static_cast<long long>(ox) * (S * 3);
                            uint32_t b[S * 3];
#pragma unroll
                            for (int j = 0; j < S; ++j) {
#pragma unroll
                                for (int c = 0; c < 3; ++c) {
                                    const int idx = c * S * S + i * S + j;
                                    const float v = __uint_as_float(acc[idx]) +
                                                    reinterpret_cast<const float*>(base_ptr + kOffBias)[idx];
                                    // y = r + x/255 ; u8 = clamp(floor(y*255 + 0.5))
                                    float o = floorf(fmaf(v, 255.f, xin[c] + 0.5f));
                                    o = fminf(fmaxf(o, 0.f), 255.f);
                                    b[j * 3 + c] = static_cast<uint32_t>(o);
                                }
                            }
                            if (wide && S == 2) {
                                uint16_t* d16 = reinterpret_cast<uint16_t*>(dp);   // 6*ox: 2-byte aligned
#pragma unroll
                                for (int k = 0; k < 3; ++k) d16[k] = static_cast<uint16_t>(b[2 * k] | (b[2 * k + 1] << 8));
                            } else if (wide && S == 4) {
                                uint32_t* d32 = reinterpret_cast<uint32_t*>(dp);   // 12*ox: 4-byte aligned
#pragma unroll
                                for (int k = 0; k < 3; ++k)
                                    d32[k] = b[4 * k] | (b[4 * k + 1] << 8) | (b[4 * k + 2] << 16) | (b[4 * k + 3] << 24);
                            } else {
#pragma unroll
                                for (int k = 0; k < S * 3; ++k) dp[k] = static_cast<uint8_t>(b[k]);
                            }
                        }
                    }
                }
            }
        }
    }

    if (!TAIL && warp >= 2 && (warp & 3) == 0 && lane == 0) bulk_wait<0>();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
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

size_t conv_weight_blob_bytes(int ng) { return w_bytes(ng); }

void pack_conv_weights(const float* w_oihw, int co, int ng, bool reverse, uint16_t* blob) {
    // blob[dx][n = g*ng + o][ci] fp16 with the 128-byte swizzle applied per 128-byte row:
    // 16-byte chunk c of row n is stored at chunk (c ^ (n & 7)).  Group g holds vertical tap
    // ky = g (forward sweep) or ky = 2 - g (reverse sweep).
    std::memset(blob, 0, w_bytes(ng));
    for (int dx = 0; dx < 3; ++dx)
        for (int g = 0; g < 3; ++g)
            for (int o = 0; o < co; ++o) {
                const int n = g * ng + o;
                const int ky = reverse ? 2 - g : g;
                for (int ci = 0; ci < 64; ++ci) {
                    const float v = w_oihw[((static_cast<size_t>(o) * 64 + ci) * 3 + ky) * 3 + dx];
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
