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
//   * The three vertical taps are stacked along N: the B operand of input row y is the three
//     weight groups W(ky=0..2) (N = 3*NG), so one read of the A tile feeds output rows y+1, y, y-1.
//   * Rotating accumulator bank: a stream of consecutive rows owns 3 TMEM slots of NG columns; the
//     accumulator of the row whose centre tap is step e lives in slot e mod 3, so every step
//     writes all three slots with ONE N = 3*NG MMA per K-step whose B operand is a cyclic rotation
//     of the three groups (the blob holds [W2|W1|W0|W2|W1], a rotation is a start offset).  No
//     MMA is ever split and none overwrites: the epilogue hands a slot back zeroed (tcgen05.st).
//   * The slot completed by step k is needed again, empty, by step k+1 of the same stream.  Each
//     CTA therefore interleaves TWO independent streams (two halves of its range of strip-rows,
//     one accumulator bank and one epilogue group each): while one stream's slot is drained the
//     tensor pipe runs the other stream's step.
//   * A stream walks its range segment by segment (a segment = consecutive rows of one strip);
//     a segment of n rows is n+2 steps (one halo row either side; rows outside the canvas are
//     TMA zero fill).  The "event" whose centre is a halo step collects garbage and is only zeroed.
//   * CTA pairs (body kernel): two CTAs of a cluster run their streams in lock-step through
//     tcgen05.mma.cta_group::2 (M = 256: each CTA's own A tile, half of B's rows from each CTA's
//     shared memory), which halves the B-operand shared-memory reads per SM; the even CTA issues.
//   * Alternate layers sweep in opposite directions (`reverse`): the rows a CTA wrote last in
//     layer l are the rows it reads first in layer l+1, while they may still be in L2.
//   * Late layers skip the canvas rows no kept pixel depends on (RowSpace below): upstream keeps only
//     the centre of every padded tile, so the work of a layer is a list of needed rows, not the canvas.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (warp-
// uniform code, one elected lane issues), warps 2..5 / 6..9 = epilogue group of stream 0 / 1
// (TMEM -> registers -> zero the slot -> bias/PReLU -> fp16 -> swizzled smem -> one TMA store per warp: every warp
// ships the quarter of the row it wrote and waits only for its own previous read-out; or for the tail: bias ->
// PixelShuffle + residual -> u8 -> global).
#include "kernels.h"

#include <algorithm>
#include <cstring>

#include "model.h"

#ifndef REVE_TAIL_WAIT_SLEEP
#define REVE_TAIL_WAIT_SLEEP 0   // tail kernel: ns its TMA producer sleeps between polls of a full ring (0 = spin)
#endif

namespace reve {

namespace {

constexpr int kRowBytes = kBoxPx * 128;  // 16 KB: 128 px * 64 ch * fp16
constexpr int kCtrlBytes = 2048;
constexpr int kGuard = 1024;
constexpr int kMaxStages = 8;
constexpr int kTailRowTab = 8192;   // canvas rows whose packed geometry the tail keeps in shared memory (32 KB)

__host__ __device__ constexpr int rows_per_dx(int ng, bool pair) { return pair ? 7 * ng / 2 : 5 * ng; }
__host__ __device__ constexpr int w_smem_bytes(int ng, bool pair) { return 3 * rows_per_dx(ng, pair) * 128; }
__host__ __device__ constexpr int tmem_cols(int ng) { return ng * 6 <= 128 ? 128 : (ng * 6 <= 256 ? 256 : 512); }
// A-row ring depth: what fits next to the weights (body, one CTA: 120 KB of weights leave room for 4 rows)
// (tail x4: 92 KB of weights + the 32 KB row table leave room for 5 rows next to the output staging below)
__host__ __device__ constexpr int ring_stages(int ng, bool tail, bool pair) { return (ng == 64 && !tail && !pair) ? 4 : ((tail && ng == 48) ? 5 : 6); }
// Tail: per epilogue warp and output sub-row, the u8 bytes of up to 32 pixels (3*S bytes each) are gathered in shared
// memory at the 16-byte phase of their global address, so that they leave as 16-byte vector stores (see the tail epilogue)
#ifndef REVE_TAIL_STAGED_MASK
// bit S set: scale S writes its u8 output through the shared-memory staging.  Default: none.  Same-box A/B on B200
// (profiles/r02_ab_tail_output_path.txt, ms per frame, direct vs staged): x2 1080p 0.098 vs 0.121, x3 540p 0.0430 vs
// 0.0443, x4 720p 0.0822 vs 0.0958 -- the direct stores win at every scale, see the tail epilogue.
#define REVE_TAIL_STAGED_MASK 0
#endif
constexpr int kTailStagedMask = REVE_TAIL_STAGED_MASK;
__host__ __device__ constexpr int tail_scale(int ng) { return ng == 16 ? 2 : (ng == 32 ? 3 : 4); }
__host__ __device__ constexpr int tail_row_stage(int ng) { return 16 + 32 * 3 * tail_scale(ng); }
__host__ __device__ constexpr int tail_out_stage_bytes(int ng) { return 8 * tail_scale(ng) * tail_row_stage(ng); }

// control block offsets (from the 1024-aligned base)
constexpr int kBarW = 0;
constexpr int kBarWPeer = 8;
constexpr int kBarAFull = 16;
constexpr int kBarAEmpty = kBarAFull + 8 * kMaxStages;
constexpr int kBarAccFull = kBarAEmpty + 8 * kMaxStages;   // [stream][slot]
constexpr int kBarAccEmpty = kBarAccFull + 8 * 6;           // [stream][slot]
constexpr int kBarStgFull = kBarAccEmpty + 8 * 6;          // chain kernel: [group][quarter] staging buffer written
constexpr int kBarStgFree = kBarStgFull + 8 * 8;           // chain kernel: [group][quarter] staging buffer read by the TMA store
constexpr int kTmemPtr = 512;
static_assert(kBarStgFree + 8 * 8 <= kTmemPtr, "control block overflow");
constexpr int kOffBias = 1024;   // 64 floats
constexpr int kOffSlope = 1280;  // 64 floats
static_assert(kBarAccEmpty + 8 * 6 <= kTmemPtr, "control block overflow");

enum : uint32_t { TAG_W = 1, TAG_A_EMPTY = 2, TAG_A_FULL = 3, TAG_ACC_EMPTY = 4, TAG_ACC_FULL = 5, TAG_W_PEER = 6 };

// Range of (virtual) strip-rows of one stream.  A worker (CTA, or CTA pair) owns a contiguous block of
// total_rows / n_workers strip-rows, cut evenly into its 2 (4) streams.  In reverse mode the virtual
// index runs backwards over the physical one (strip' = n_strips-1-strip, y' = CH-1-y) and the worker
// takes the mirrored block, i.e. the same physical region as in forward mode.
template <bool PAIR>
__device__ __forceinline__ void stream_range(const ConvParams& p, uint32_t rank, int s, long long& lo, long long& hi) {
    const unsigned n_workers = PAIR ? gridDim.x / 2 : gridDim.x;
    const unsigned w = PAIR ? blockIdx.x / 2 : blockIdx.x;
    const unsigned b = p.reverse ? (n_workers - 1 - w) : w;
    const long long wlo = static_cast<long long>(b) * p.total_rows / n_workers;
    const long long whi = static_cast<long long>(b + 1) * p.total_rows / n_workers;
    const int n_streams = PAIR ? 4 : 2;
    const int q = PAIR ? static_cast<int>(rank) * 2 + s : s;
    lo = wlo + (whi - wlo) * q / n_streams;
    hi = wlo + (whi - wlo) * (q + 1) / n_streams;
}
// Rows of a layer.  Upstream's tiles overlap by 10 px but the network's receptive field keeps shrinking towards
// the output, so late layers need fewer and fewer rows of every padded tile (a row r pixels outside the kept
// region matters only to layers at least r convolutions before the end).  A layer works on its list of needed
// canvas rows (`rowmap`, ascending; null = every row): virtual row index v -> virtual row number, and the length
// of the run of consecutive rows starting there.  In reverse mode the list is walked backwards and rows are
// numbered from the bottom (y' = CH-1-y), as everywhere in this kernel.
struct RowSpace {
    const int* rowmap;
    const int* run_fwd;
    const int* run_bwd;
    int nr, ch;
    bool rev;
    __device__ RowSpace() {}
    __device__ RowSpace(const ConvParams& p)
        : rowmap(p.rowmap), run_fwd(p.run_fwd), run_bwd(p.run_bwd), nr(p.n_rows), ch(p.canvas_h), rev(p.reverse != 0) {}
    __device__ int row(int v) const { return !rowmap ? v : (rev ? ch - 1 - rowmap[nr - 1 - v] : rowmap[v]); }
    __device__ int run(int v) const { return !rowmap ? nr - v : (rev ? run_bwd[nr - 1 - v] : run_fwd[v]); }
};

// steps of a stream: every segment (a run of consecutive rows of one strip) costs its rows plus two halo rows
// (chained layers: plus `ext` extra rows either side, see conv3x3_chain_kernel)
__device__ __noinline__ int stream_steps(const RowSpace rs, long long lo, long long hi, int ext = 0) {
    int steps = 0;
    while (lo < hi) {
        const int v = static_cast<int>(lo % rs.nr);
        const long long n = min(static_cast<long long>(rs.run(v)), hi - lo);
        steps += static_cast<int>(n) + 2 + 2 * ext;
        lo += n;
    }
    return steps;
}

// Walks the steps of a stream: (strip, virtual row y, interior?).  Past the end it yields padding steps
// (row -1 = outside the canvas, never interior) so that the two CTAs of a pair stay in lock-step.
// Start of the segment at virtual position pos: strip, first row and length, packed (20 | 20 | 24 bits) so the
// result travels in registers.  Rare (a few times per launch) and deliberately out of line: the 64-bit divisions
// would otherwise be inlined into the producer's and the epilogue's per-step loops, and the hot loops of the warps
// that share a scheduler have to stay within its instruction cache (measured: +7 % cycles per strip row).
__device__ __noinline__ unsigned long long locate_segment(const int* rowmap, const int* run_fwd, const int* run_bwd,
                                                          int nr, int ch, int rev, long long pos, long long hi) {
    RowSpace rs;
    rs.rowmap = rowmap; rs.run_fwd = run_fwd; rs.run_bwd = run_bwd; rs.nr = nr; rs.ch = ch; rs.rev = rev != 0;
    const unsigned long long strip = static_cast<unsigned long long>(pos / nr);
    const int v = static_cast<int>(pos % nr);
    const unsigned long long ya = static_cast<unsigned long long>(rs.row(v));
    const unsigned long long n = static_cast<unsigned long long>(min(static_cast<long long>(rs.run(v)), hi - pos));
    return ya | (strip << 20) | (n << 40);
}

struct Cursor {
    RowSpace rs;
    long long pos, hi;
    int strip = 0, ya = 0, yb = -2, y = 0;
    __device__ Cursor(const RowSpace& rs_, long long lo_, long long hi_) : rs(rs_), pos(lo_), hi(hi_) {}
    __device__ __forceinline__ bool next(int& strip_o, int& y_o, bool& new_segment) {
        new_segment = false;
        if (y > yb + 1) {
            if (pos >= hi) {
                strip_o = strip;
                y_o = -1;
                return false;
            }
            const unsigned long long seg = locate_segment(rs.rowmap, rs.run_fwd, rs.run_bwd, rs.nr, rs.ch, rs.rev, pos, hi);
            ya = static_cast<int>(seg & 0xFFFFFu);
            strip = static_cast<int>((seg >> 20) & 0xFFFFFu);
            const int n = static_cast<int>(seg >> 40);
            yb = ya + n - 1;
            pos += n;
            y = ya - 1;
            new_segment = true;
        }
        strip_o = strip;
        y_o = y;
        const bool interior = (y >= ya) && (y <= yb);
        ++y;
        return interior;
    }
};

// Interleaved order of the steps of the two streams: k = 1, 2, ...; stream 0 then stream 1 (each while it
// still has steps).  Producer, MMA issuer and (implicitly) the epilogue groups all follow this order.
struct Sequencer {
    int u0, u1, k = 1, s = -1;
    __device__ Sequencer(int u0_, int u1_) : u0(u0_), u1(u1_) {}
    __device__ bool next() {
        for (;;) {
            if (++s == 2) {
                s = 0;
                ++k;
            }
            if (k > u0 && k > u1) return false;
            if (k <= (s == 0 ? u0 : u1)) return true;
        }
    }
};

__device__ __forceinline__ uint64_t mk_desc(uint32_t hi, uint32_t lo) {
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

[[maybe_unused]] __device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
[[maybe_unused]] __device__ __forceinline__ void st_shared_u8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
[[maybe_unused]] __device__ __forceinline__ void st_shared_u16(uint32_t addr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<unsigned short>(v)) : "memory"); }
[[maybe_unused]] __device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_shared_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

// Tail: its MMAs are small (N = 3*NG <= 144), and one thread issuing for both streams leaves gaps in the tensor pipe
// (1030 cycles per 12-MMA step); with one issuing warp per stream (warp 1: stream 0, warp 10: stream 1; the two
// threads' MMAs target different TMEM banks and interleave in the pipe) a step takes ~940 and the tail runs 9 % faster.
constexpr int conv_threads(bool tail) { return tail ? kConvThreads + 32 : kConvThreads; }

template <int NG, bool TAIL, bool PAIR>
__global__ void __launch_bounds__(conv_threads(TAIL), 1)
conv3x3_umma_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map_q,
                    const __grid_constant__ CUtensorMap out_map_e, const __grid_constant__ ConvParams p) {
    // body: out_map_q / out_map_e = the output canvas with boxes of 32 and 31 pixels (every epilogue warp stores the
    // pixels it wrote with its own TMA store: the quarters of the row, the two end quarters one halo pixel short)
    constexpr int kStages = ring_stages(NG, TAIL, PAIR);
    constexpr int kRowsDx = rows_per_dx(NG, PAIR);
    constexpr int kWBytes = w_smem_bytes(NG, PAIR);
    constexpr int kBank = 3 * NG;               // TMEM columns of one stream's accumulator bank
    constexpr int kTmemCols = tmem_cols(NG);
    constexpr int kOffW = kCtrlBytes;
    constexpr int kOffRing = kOffW + kWBytes + kGuard;
    constexpr int kOffStage = kOffRing + kStages * kRowBytes + kGuard;  // body: 2 x 16 KB output staging (one per group)
    constexpr int kOffRowTab = kOffStage;                                // tail: packed row table (kTailRowTab entries)
    constexpr int kOffOutStage = kOffRowTab + kTailRowTab * 4;           // tail: u8 output staging, [epilogue warp][sub-row]
    static_assert(kWBytes % 1024 == 0, "the A ring must stay 1024-byte aligned");
    static_assert(kStages <= kMaxStages, "ring too deep for the control block");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* const base_ptr = smem_raw + (base - raw);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    DebugBlock* const dbg = p.dbg;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = (rank == 0);

    if (threadIdx.x == 0) {
        mbar_init(base + kBarW, 1);
        mbar_init(base + kBarWPeer, 1);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(base + kBarAFull + 8 * s, 1);
            mbar_init(base + kBarAEmpty + 8 * s, 1);
        }
        for (int s = 0; s < 6; ++s) {
            mbar_init(base + kBarAccFull + 8 * s, 1);
            mbar_init(base + kBarAccEmpty + 8 * s, PAIR ? 8 : 4);  // one arrive per warp of the draining group(s)
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        if constexpr (PAIR) {
            tmem_alloc_pair(base + kTmemPtr, kTmemCols);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(base + kTmemPtr, kTmemCols);
            tmem_relinquish();
        }
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&in_map);
        if (!TAIL) { prefetch_tmap(&out_map_q); prefetch_tmap(&out_map_e); }
    }
    const uint32_t* rowtab = p.rowpack;
    if constexpr (TAIL) {
        if (p.canvas_h <= kTailRowTab) {
            uint32_t* const t = reinterpret_cast<uint32_t*>(base_ptr + kOffRowTab);
            for (int i = threadIdx.x; i < p.canvas_h; i += conv_threads(TAIL)) t[i] = p.rowpack[i];
            rowtab = t;
        }
    }
    if (threadIdx.x >= 64 && threadIdx.x < 128) {
        reinterpret_cast<float*>(base_ptr + kOffBias)[threadIdx.x - 64] = p.bias[threadIdx.x - 64];
        reinterpret_cast<float*>(base_ptr + kOffSlope)[threadIdx.x - 64] = p.slope[threadIdx.x - 64];
    }
    tc_fence_before();
    // pair: the barriers must be initialised (and the pair-wide TMEM allocation complete) in both CTAs before
    // either sends the other anything
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + kTmemPtr);

    const int CH = p.canvas_h;
    const bool rev = p.reverse != 0;

    // steps per stream; in a pair both CTAs run the larger count (the shorter stream pads)
    const RowSpace rspace(p);
    int U[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        long long lo, hi;
        stream_range<PAIR>(p, rank, s, lo, hi);
        U[s] = stream_steps(rspace, lo, hi);
        if constexpr (PAIR) {
            stream_range<PAIR>(p, rank ^ 1u, s, lo, hi);
            U[s] = max(U[s], stream_steps(rspace, lo, hi));
        }
    }

    if (warp >= 2 && warp < 10) {
        // every accumulator slot starts out zero: this group's bank, this warp's 32 lanes
        const int grp = (warp - 2) >> 2;
        const uint32_t t = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + grp * kBank;
#pragma unroll
        for (int c = 0; c < kBank / 16; ++c) tmem_st16_fill(t + c * 16, 0u);
        tmem_wait_st();
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            if constexpr (PAIR) {
                // this CTA's half of every rotation: rows [rank * 1.5 NG, +3.5 NG) of each dx block of the blob
                const uint32_t r = rank ^ (p.flags & 1u);
                mbar_arrive_expect_tx(base + kBarW, kWBytes);
                for (int dx = 0; dx < 3; ++dx)
                    bulk_load_1d(base + kOffW + dx * kRowsDx * 128,
                                 static_cast<const uint8_t*>(p.weights) + (dx * 5 * NG + r * (3 * NG / 2)) * 128,
                                 kRowsDx * 128, base + kBarW);
            } else {
                mbar_arrive_expect_tx(base + kBarW, kWBytes);
                bulk_load_1d(base + kOffW, p.weights, kWBytes, base + kBarW);
            }
            const uint64_t policy = kPolicyEvictFirst;  // activations are read once per layer
            long long lo0, hi0, lo1, hi1;
            stream_range<PAIR>(p, rank, 0, lo0, hi0);
            stream_range<PAIR>(p, rank, 1, lo1, hi1);
            Cursor cur0(rspace, lo0, hi0), cur1(rspace, lo1, hi1);
            Sequencer seq(U[0], U[1]);
            const uint32_t full_base = PAIR ? mapa_u32(base + kBarAFull, 0) : (base + kBarAFull);   // the even CTA's barriers
            uint32_t i = 0;
            while (seq.next()) {
                int strip, y;
                bool newseg;
                if (seq.s == 0) cur0.next(strip, y, newseg); else cur1.next(strip, y, newseg);
                const int x0 = (rev ? p.n_strips - 1 - strip : strip) * kStripPx;
                const uint32_t stage = i % kStages, use = i / kStages;
                if constexpr (TAIL) mbar_wait_relaxed<REVE_TAIL_WAIT_SLEEP>(base + kBarAEmpty + 8 * stage, (use & 1) ^ 1, dbg, TAG_A_EMPTY, i);
                else mbar_wait(base + kBarAEmpty + 8 * stage, (use & 1) ^ 1, dbg, TAG_A_EMPTY, i);
                if constexpr (PAIR) {
                    if (leader) mbar_arrive_expect_tx(base + kBarAFull + 8 * stage, 2 * kRowBytes);
                    tma_load_3d_hint_pair(base + kOffRing + stage * kRowBytes, &in_map, full_base + 8 * stage, 0,
                                          x0 - 1, rev ? CH - 1 - y : y, policy);
                } else {
                    mbar_arrive_expect_tx(base + kBarAFull + 8 * stage, kRowBytes);
                    tma_load_3d_hint(base + kOffRing + stage * kRowBytes, &in_map, base + kBarAFull + 8 * stage, 0,
                                     x0 - 1, rev ? CH - 1 - y : y, policy);
                }
                ++i;
            }
        }
    } else if (warp == 1 || (TAIL && warp == 10)) {
        // ------------------------------------------------------------------ MMA issuer(s)
        // Warp-uniform control flow; one elected lane issues.  The tensor pipe's instruction queue is
        // shallow, so the barriers of the next step are polled in the middle of the current one.
        constexpr bool kSplit = TAIL;                 // one issuing warp per stream
        const int mine = (warp == 1) ? 0 : 1;
        mbar_wait(base + kBarW, 0, dbg, TAG_W);
        if constexpr (PAIR) {
            if (!leader) {
                if (lane == 0) mbar_arrive_cluster(mapa_u32(base + kBarWPeer, 0));
            } else {
                mbar_wait(base + kBarWPeer, 0, dbg, TAG_W_PEER);
            }
        }
        if (leader) {
            tc_fence_after();
            const uint32_t idesc = umma_idesc_f16(PAIR ? 256 : 128, 3 * NG);
            const uint64_t proto = umma_desc_sw128(0, 0);
            const uint32_t desc_hi = static_cast<uint32_t>(proto >> 32);
            const uint32_t lo_flags = static_cast<uint32_t>(proto);              // LBO field
            const uint32_t w_lo = lo_flags | ((base + kOffW) >> 4);
            const uint32_t ring_lo = lo_flags | ((base + kOffRing) >> 4);
            constexpr uint32_t kDx = kRowsDx * 8;   // one dx block of B, in 16-byte units
            constexpr uint32_t kRot = NG * 8;       // one weight group

            auto mma = [&](uint32_t d, uint64_t a, uint64_t b) {
                if constexpr (PAIR) umma_f16_pair(d, a, b, idesc, 1u); else umma_f16(d, a, b, idesc, 1u);
            };
            auto commit = [&](uint32_t bar) {
                if constexpr (PAIR) umma_commit_pair(bar, 3); else umma_commit(bar);
            };
            // barriers a step has to pass before its MMAs may be issued
            struct Gate { uint32_t bar_f, par_f, bar_e, par_e; bool need_e; };
            auto gate_of = [&](uint32_t i, int s, int k) {
                Gate g;
                g.bar_f = base + kBarAFull + 8 * (i % kStages);
                g.par_f = (i / kStages) & 1;
                // fresh slot of step k = slot of event k+1, last used by event k-2
                g.need_e = k >= 2;
                g.bar_e = base + kBarAccEmpty + 8 * (s * 3 + (k + 1) % 3);
                g.par_e = ((k - 2) / 3) & 1;
                return g;
            };
            Sequencer seq(U[0], U[1]);
            uint32_t i = 0, n_seen = 0;   // i = index of the current step in the interleaved order of both streams
            auto next_step = [&]() -> bool {   // the next step this warp issues
                while (seq.next()) {
                    const uint32_t idx = n_seen++;
                    if (!kSplit || seq.s == mine) {
                        i = idx;
                        return true;
                    }
                }
                return false;
            };
            bool have = next_step();
            if (have) {
                const Gate g = gate_of(i, seq.s, seq.k);
                mbar_wait(g.bar_f, g.par_f, dbg, TAG_A_FULL, 0);
                tc_fence_after();
            }
            if (p.trace && lane == 0 && warp == 1) p.trace[2048 + blockIdx.x * 4 + 1] = static_cast<long long>(globaltimer_ns());
            const bool elected = elect_one();
            uint32_t n_issued = 0;
            while (have) {
                const int s = seq.s, k = seq.k;
                const uint32_t stage = i % kStages;
                const uint32_t a_lo = ring_lo + stage * (kRowBytes >> 4);
                // event k+1-g (vertical tap g) lives in slot (k+1-g) mod 3: B starts (2 - (k+1) mod 3) groups into the blob
                uint32_t w_row = w_lo + (2 - (k + 1) % 3) * kRot;
                asm volatile("" : "+r"(w_row));   // keep the 12 B descriptors as adds on one per-step value
                const uint32_t d = tmem_base + s * kBank;
                if (p.trace && blockIdx.x == 0 && i < 1000 && lane == 0) p.trace[i] = clock64();
                if (elected) {
#pragma unroll
                    for (int dxk = 0; dxk < 4; ++dxk)
                        mma(d, mk_desc(desc_hi, a_lo - 8 + dxk * 2), mk_desc(desc_hi, w_row + dxk * 2));
                }
                // look ahead: barriers of the next step
                have = next_step();
                Gate g{};
                bool ok = true;
                if (have) {
                    g = gate_of(i, seq.s, seq.k);
                    ok = mbar_test_wait(g.bar_f, g.par_f);   // a poll: must not suspend with 8 MMAs still to issue
                    if (g.need_e) ok = mbar_test_wait(g.bar_e, g.par_e) && ok;
                }
                if (elected) {
#pragma unroll
                    for (int dxk = 4; dxk < 12; ++dxk) {
                        const int dx = dxk >> 2, kk = dxk & 3;
                        mma(d, mk_desc(desc_hi, a_lo + (dx - 1) * 8 + kk * 2), mk_desc(desc_hi, w_row + dx * kDx + kk * 2));
                    }
                    commit(base + kBarAEmpty + 8 * stage);                       // ring slot consumed
                    commit(base + kBarAccFull + 8 * (s * 3 + (k - 1) % 3));      // event k-1 complete
                }
                if (!ok) {
                    if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[1000] += 1;   // look-ahead misses
                    mbar_wait(g.bar_f, g.par_f, dbg, TAG_A_FULL, i);
                    if (g.need_e) mbar_wait(g.bar_e, g.par_e, dbg, TAG_ACC_EMPTY, seq.k);
                }
                tc_fence_after();
                ++n_issued;
            }
            if (p.trace && lane == 0 && warp == 1) {   // per-CTA wall clock (ns): first step, end of the last step, steps | smid << 32
                unsigned smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                p.trace[2048 + blockIdx.x * 4 + 2] = static_cast<long long>(globaltimer_ns());
                p.trace[2048 + blockIdx.x * 4 + 3] = static_cast<long long>(n_issued) | (static_cast<long long>(smid) << 32);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps
        const int grp = (warp - 2) >> 2;   // group = stream
        const int q = warp & 3;            // TMEM lane quarter this warp may access
        const int m = q * 32 + lane;       // M row = pixel index inside the 128-px box
        const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + grp * kBank;
        const uint32_t empty_base = PAIR ? mapa_u32(base + kBarAccEmpty + 8 * grp * 3, 0) : (base + kBarAccEmpty + 8 * grp * 3);
        long long lo, hi;
        stream_range<PAIR>(p, rank, grp, lo, hi);
        Cursor cur(rspace, lo, hi);
        const int n_events = U[grp];   // events 0 .. U-1 (event e is completed by step e+1)
        // What an event needs besides its accumulator: the geometry of its canvas row and (tail) the residual
        // pixel it points to.  Body: one table byte, loaded before the accumulator wait (the group idles ~1000
        // cycles there anyway).  Tail: its epilogue is the critical resource, so the row table sits in shared
        // memory (packed, see ConvParams::rowpack) and the residual pixel of event e+1 is prefetched into L1/L2
        // while event e is processed.  (Carrying in-flight loads across iterations does not work: the compiler
        // copies their destination registers right after issue, a ~1000-cycle stall per event, measured.)
        struct Event {
            bool valid = false, keep = false;
            int pr = 0, x0 = 0, ox = -1;
            int oy = -1, fr = 0;             // tail
            const uint8_t* pix = nullptr;    // tail: residual pixel (3 bytes)
        };
        int seg_x0 = 0, seg_ox = -1, seg_sx = 0;
        bool seg_colok = false;
        auto make_event = [&](int e) {
            Event ev;
            if (e < 1 || e >= n_events) return ev;   // event 0 has no centre step
            int strip, y;
            bool newseg;
            ev.valid = cur.next(strip, y, newseg);
            if (newseg) {
                seg_x0 = (rev ? p.n_strips - 1 - strip : strip) * kStripPx;
                const int cx = seg_x0 - 1 + m;
                const bool inside = (m >= 1) && (m <= kStripPx) && (cx < p.canvas_w);
                seg_colok = inside && (p.colflag[cx] != 0);
                seg_ox = -1;
                if (TAIL && inside) {
                    seg_ox = p.out_x[cx];
                    seg_sx = p.src_x[cx];
                }
            }
            ev.pr = rev ? CH - 1 - y : y;  // physical canvas row
            ev.x0 = seg_x0;
            ev.ox = seg_ox;
            if (ev.valid) {
                if constexpr (!TAIL) {
                    ev.keep = seg_colok && (p.rowflag[ev.pr] != 0);
                } else {
                    const uint32_t rp = rowtab[ev.pr];   // (out_y + 1) << 17 | (src_y + 1) << 2 | frame
                    ev.oy = static_cast<int>(rp >> 17) - 1;
                    const int sy = static_cast<int>((rp >> 2) & 0x7FFFu) - 1;
                    ev.fr = static_cast<int>(rp & 3u);
                    if (ev.ox >= 0 && ev.oy >= 0) {
                        ev.pix = p.src[ev.fr] + static_cast<long long>(sy) * p.src_stride + seg_sx * 3;
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(ev.pix));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(ev.pix + 2));
                    }
                }
            }
            return ev;
        };
        Event nxt;
        if constexpr (TAIL) nxt = make_event(0);
        for (int e = 0; e < n_events; ++e) {
            Event ev;
            if constexpr (TAIL) ev = nxt; else ev = make_event(e);
            const bool valid = ev.valid;
            const int pr = ev.pr, x0 = ev.x0, ox = ev.ox, oy = ev.oy, fr = ev.fr;
            const int slot = e % 3;
            long long* const tr = (p.trace && blockIdx.x == 0 && grp == 0 && e < 256 && q == 0 && lane == 0)
                                      ? p.trace + 1024 + e * 4 : nullptr;
            if (tr) tr[0] = clock64();
            mbar_wait(base + kBarAccFull + 8 * (grp * 3 + slot), (e / 3) & 1, dbg, TAG_ACC_FULL, e);
            if (tr) tr[1] = clock64();
            tc_fence_after();
            uint32_t acc[NG];
            if (valid) {
#pragma unroll
                for (int c = 0; c < NG / 16; ++c) {
                    uint32_t(&dst)[16] = *reinterpret_cast<uint32_t(*)[16]>(&acc[c * 16]);
                    tmem_ld16(tmem_lane + slot * NG + c * 16, dst);
                }
                tmem_wait_ld();
            }
#pragma unroll
            for (int c = 0; c < NG / 16; ++c) tmem_st16_fill(tmem_lane + slot * NG + c * 16, 0u);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) mbar_arrive_cluster(empty_base + 8 * slot); else mbar_arrive(empty_base + 8 * slot);
            }
            if (tr) tr[2] = clock64();
            if constexpr (TAIL) {
                // nothing of the next event may be scheduled into the drain above: its cursor depends on `after`
                int after;
                asm volatile("mov.u32 %0, 0;" : "=r"(after)::"memory");
                cur.y += after;
                nxt = make_event(e + 1);
                if (tr) tr[3] = clock64();   // tail: next event located (the body records the row store here)
            }
            if (!valid) continue;

            if constexpr (!TAIL) {
                const uint32_t stg = base + kOffStage + grp * kRowBytes;
                if (lane == 0) bulk_wait_read<0>();   // this warp's previous store has read its quarter of the staging buffer out
                __syncwarp();
                if (m >= 1 && m <= kStripPx) {
                    const int row = m;                  // staged in place: pixel m of the box is row m
                    const uint32_t rbase = stg + row * 128;
                    if (ev.keep) {
#pragma unroll
                        for (int c8 = 0; c8 < 8; ++c8) {
                            uint32_t pk[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int ch = c8 * 8 + j * 2;
                                // bias in fp32, then PReLU on the packed fp16 pair: max(v,0) + a*min(v,0).
                                // Positive values are bit-identical to the fp32 formulation; negative ones
                                // round twice (<= 1 ulp of fp16 instead of 0.5).  The _nan variants keep a NaN a
                                // NaN (plain hmin2/hmax2 return the other operand and would turn it into 0), as
                                // the `x < 0 ? x * slope : x` of ncnn's PReLU does; +-inf pass through either way.
                                const __half2 v = __floats2half2_rn(__uint_as_float(acc[ch]) + p.bias[ch],
                                                                    __uint_as_float(acc[ch + 1]) + p.bias[ch + 1]);
                                const __half2 z = __float2half2_rn(0.f);
                                const __half2 r = __hfma2(p.slope2[ch >> 1], __hmin2_nan(v, z), __hmax2_nan(v, z));
                                pk[j] = *reinterpret_cast<const uint32_t*>(&r);
                            }
                            st_shared_v4(rbase + ((c8 ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
                        }
                    } else {   // gap pixel (between tiles / frames): must read as zero in the next layer
#pragma unroll
                        for (int c8 = 0; c8 < 8; ++c8) st_shared_v4(rbase + ((c8 ^ (row & 7)) << 4), 0u, 0u, 0u, 0u);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    // pixels [32q, 32q+32) of the box without its two halo pixels 0 and 127; box pixel m is canvas column x0-1+m
                    const int m0 = (q == 0) ? 1 : 32 * q;
                    tma_store_3d((q == 0 || q == 3) ? &out_map_e : &out_map_q, stg + m0 * 128, 0, x0 - 1 + m0, pr);
                    bulk_commit();
                }
                if (tr) tr[3] = clock64();
            } else {
                // u8 output: every pixel owns 3*S consecutive bytes in each of S output rows.  Two ways to write them:
                //  * direct: 3*S/2 two-byte (x2), 3*S/4 four-byte (x4) or 3*S single-byte (x3) stores per pixel and sub-row;
                //    neighbouring lanes write neighbouring pixels, so a warp-wide store instruction covers one contiguous
                //    span (192 / 288 / 384 bytes) with interleaved pieces that L2 merges;
                //  * staged (kTailStagedMask, bit S): the valid pixels of a warp (consecutive canvas columns of one canvas
                //    row) are consecutive output pixels -- across a tile boundary too: the cropped pre-pad columns in
                //    between have ox < 0 and the kept columns of neighbouring tiles abut in the output.  So per output
                //    sub-row the warp owns ONE contiguous run of bytes; it is assembled in shared memory at the 16-byte
                //    phase of its global address and leaves as 16-byte stores (one per lane) plus at most 15 single bytes
                //    at either end.
                // Measured on B200 (REVE_TAIL_STAGED_MASK above): the staged path is slower at every scale, by 24 % (x2), 3 %
                // (x3) and 17 % (x4) of the tail kernel.  The tail's epilogue warps are its critical resource (DESIGN.md
                // section 4.2): the ballot / shuffle / shared-memory round trip adds more instructions to them than the
                // narrow stores cost, and L2 merges the pieces of a warp-wide store into full sectors anyway (the output
                // is 1/13 of the kernel's traffic).  The direct path is what ships; the staged one stays selectable.
                constexpr int S = tail_scale(NG);
                constexpr int kPx = 3 * S;
                constexpr int kRowStage = tail_row_stage(NG);
                if constexpr (((kTailStagedMask >> S) & 1) == 0) {
                if (ox >= 0 && oy >= 0) {
                    const unsigned rgb[3] = {ev.pix[0], ev.pix[1], ev.pix[2]};
                    const bool wide = (S != 3) && (((reinterpret_cast<uintptr_t>(p.dst[fr]) | static_cast<uintptr_t>(p.dst_stride)) & 3) == 0);
#pragma unroll
                    for (int i = 0; i < S; ++i) {
                        uint8_t* dp = p.dst[fr] + static_cast<long long>(oy * S + i) * p.dst_stride + static_cast<long long>(ox) * kPx;
                        uint32_t b[kPx];
#pragma unroll
                        for (int j = 0; j < S; ++j) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                const int ch = c * S * S + i * S + j;
                                const float v = __uint_as_float(acc[ch]) + reinterpret_cast<const float*>(base_ptr + kOffBias)[ch];
                                // y = r + x/255 ; u8 = clamp(floor(y*255 + 0.5)); a NaN (overflowed fp16 activations) becomes 0
                                float o = floorf(fmaf(v, 255.f, static_cast<float>(rgb[c]) + 0.5f));
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
                            for (int k = 0; k < kPx; ++k) dp[k] = static_cast<uint8_t>(b[k]);
                        }
                    }
                }
                } else {
                const bool ok = (ox >= 0) && (oy >= 0);
                const unsigned okmask = __ballot_sync(0xffffffffu, ok);
                if (okmask != 0u) {
                    const int idx = __popc(okmask & ((1u << lane) - 1u));
                    const int tot = __popc(okmask) * kPx;
                    const int first = __ffs(static_cast<int>(okmask)) - 1;
                    unsigned rgb[3] = {0u, 0u, 0u};
                    if (ok) { rgb[0] = ev.pix[0]; rgb[1] = ev.pix[1]; rgb[2] = ev.pix[2]; }
                    uint8_t* const mine = p.dst[fr] + static_cast<long long>(oy * S) * p.dst_stride + static_cast<long long>(ox) * kPx;
                    const unsigned long long g_first = __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(mine), first);
                    // vector stores into the staging need the run to start on a 2- (x2) / 4-byte (x4) boundary
                    const bool wide = (S != 3) && (((reinterpret_cast<uintptr_t>(p.dst[fr]) | static_cast<uintptr_t>(p.dst_stride)) & 3) == 0);
                    const uint32_t wst = base + kOffOutStage + static_cast<uint32_t>(warp - 2) * (S * kRowStage);
                    if (ok) {
#pragma unroll
                        for (int i = 0; i < S; ++i) {
                            const unsigned a = static_cast<unsigned>((g_first + static_cast<unsigned long long>(i) * p.dst_stride) & 15ull);
                            const uint32_t sp = wst + i * kRowStage + a + idx * kPx;
                            uint32_t b[kPx];
#pragma unroll
                            for (int j = 0; j < S; ++j) {
#pragma unroll
                                for (int c = 0; c < 3; ++c) {
                                    const int ch = c * S * S + i * S + j;
                                    const float v = __uint_as_float(acc[ch]) + reinterpret_cast<const float*>(base_ptr + kOffBias)[ch];
                                    // y = r + x/255 ; u8 = clamp(floor(y*255 + 0.5)); a NaN (overflowed fp16 activations) becomes 0
                                    float o = floorf(fmaf(v, 255.f, static_cast<float>(rgb[c]) + 0.5f));
                                    o = fminf(fmaxf(o, 0.f), 255.f);
                                    b[j * 3 + c] = static_cast<uint32_t>(o);
                                }
                            }
                            if (wide && S == 2) {
#pragma unroll
                                for (int k = 0; k < 3; ++k) st_shared_u16(sp + 2 * k, b[2 * k] | (b[2 * k + 1] << 8));
                            } else if (wide && S == 4) {
#pragma unroll
                                for (int k = 0; k < 3; ++k)
                                    st_shared_u32(sp + 4 * k, b[4 * k] | (b[4 * k + 1] << 8) | (b[4 * k + 2] << 16) | (b[4 * k + 3] << 24));
                            } else {
#pragma unroll
                                for (int k = 0; k < kPx; ++k) st_shared_u8(sp + k, b[k]);
                            }
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < S; ++i) {
                        const unsigned long long g = g_first + static_cast<unsigned long long>(i) * p.dst_stride;
                        const int a = static_cast<int>(g & 15ull);
                        uint8_t* const g0 = reinterpret_cast<uint8_t*>(g - a);            // 16-byte aligned
                        const uint32_t sp = wst + i * kRowStage;
                        const int end = a + tot;                                            // run = staging bytes [a, end)
                        const int head_end = min(end, (a + 15) & ~15);                      // bytes before the first whole chunk
                        const int tail_beg = max(end & ~15, head_end);                      // bytes after the last whole chunk
                        const int c0 = 16 * lane;
                        if (c0 >= head_end && c0 + 16 <= tail_beg) {
                            const uint4 v = ld_shared_v4(sp + c0);
                            *reinterpret_cast<uint4*>(g0 + c0) = v;
                        }
                        if (a + lane < head_end) g0[a + lane] = static_cast<uint8_t>(ld_shared_u8(sp + a + lane));
                        if (tail_beg + lane < end) g0[tail_beg + lane] = static_cast<uint8_t>(ld_shared_u8(sp + tail_beg + lane));
                    }
                }
                }   // staged
            }
        }
    }

    if (!TAIL && warp >= 2 && warp < 10 && lane == 0) bulk_wait<0>();
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_pair(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
    }
}


// ====================================================================================================
// Chained body layers (ChainParams in kernels.h): CTA b computes layer j = b % len of chain b / len.
//
// All layers of a chain walk the SAME segments (those of the chain's last layer) in the same order with the same
// two-stream interleave; layer j computes ext = len-1-j extra rows either side of every segment, so that what it
// produces is exactly what layer j+1 consumes, step for step: the rows layer j completes (its valid events) are the
// steps of layer j+1, in order, per stream.  Horizontally every layer works in the coordinates of the chain's 128-pixel
// input box: layer j's valid output pixels are box rows 1+j .. 126-j, the last layer stores 128 - 2*len columns.
//
// Hand-over of one row (16 KB, already in the swizzled layout the next layer's UMMA descriptors expect):
//   sender epilogue group: staging buffer -> mbarrier -> the group's courier warp: TMA store into slot
//     (row mod kChainSlots) of the link's scratch ring -> cp.async.bulk.wait_group 0 -> st.release.gpu published = rows
//     stored so far (the global-memory round trips stay off the epilogue warps);
//   receiver loader thread: ld.acquire.gpu published >= row -> TMA load of the slot into its A ring ->
//     (two steps later, when the load has landed) consumed = row; the courier polls `consumed` before it overwrites a
//     slot.
// The scratch rings (kChainSlots x 16 KB per link and stream) are rewritten continuously and stay in L2: of the
// 2 x len canvas passes that len separate layers cost, only one read and one write reach HBM.
// Everything that is out of the canvas, in a gap row/column or outside a layer's valid pixel range is handed on as
// zeros, which is what the next layer's SAME padding expects.
struct ChainCursor {
    RowSpace rs;
    long long pos, hi;
    int ext;
    int strip = 0, ya = 0, yb = -2, y = 0;
    __device__ ChainCursor(const RowSpace& rs_, long long lo_, long long hi_, int ext_) : rs(rs_), pos(lo_), hi(hi_), ext(ext_) {}
    // (strip, virtual row y, interior?): y may lie up to ext+1 rows outside the canvas
    __device__ __forceinline__ bool next(int& strip_o, int& y_o, bool& new_segment) {
        new_segment = false;
        if (y > yb + 1) {
            if (pos >= hi) {
                strip_o = strip;
                y_o = -1;
                return false;
            }
            const unsigned long long seg = locate_segment(rs.rowmap, rs.run_fwd, rs.run_bwd, rs.nr, rs.ch, rs.rev, pos, hi);
            const int a = static_cast<int>(seg & 0xFFFFFu);
            strip = static_cast<int>((seg >> 20) & 0xFFFFFu);
            const int n = static_cast<int>(seg >> 40);
            ya = a - ext;
            yb = a + n - 1 + ext;
            pos += n;
            y = ya - 1;
            new_segment = true;
        }
        strip_o = strip;
        y_o = y;
        const bool interior = (y >= ya) && (y <= yb);
        ++y;
        return interior;
    }
};

#ifndef REVE_CONSUMED_RELEASE
#define REVE_CONSUMED_RELEASE 0
#endif
// Hand-over staging in quarters: every epilogue warp owns the 32 pixel rows (4 KB) it writes, with its own full / free
// barriers and its own TMA store.  With ONE buffer per group a row can be written only after the previous row's 16 KB
// have been read out by the TMA store -- at ~15 B/clk, because the tensor core's operand reads have the shared memory --
// so write (~700 clk) and read-out (~1100) are strictly serial and leave 25 % slack in a stream's row period; one
// step in ten then waits ~700 clk for an accumulator slot (profiles/r02_notes.md section 12).  In quarters a warp waits
// for the read-out of its own 4 KB only, and the read-out of a row starts while the other warps still compute.
#ifndef REVE_STAGE_QUARTERS
#define REVE_STAGE_QUARTERS 1
#endif
// The chain's last layer stores its canvas rows the same way: every epilogue warp ships the valid pixels it wrote (32, or
// 32 - C at the two ends of the box) with its own TMA store and waits for ITS previous read-out only; no barrier across
// the group is left in that path.
#ifndef REVE_LAST_QUARTERS
#define REVE_LAST_QUARTERS 1
#endif
#ifndef REVE_PUBLISH_EVERY
#define REVE_PUBLISH_EVERY 2
#endif
constexpr int kChainPublishEvery = REVE_PUBLISH_EVERY;   // rows per gpu-scope release (see the courier)
enum : uint32_t { TAG_CHAIN_PUB = 7, TAG_CHAIN_CONS = 8, TAG_STG_FULL = 9, TAG_STG_FREE = 10 };
constexpr int kChainThreads = kConvThreads + 64;   // + one courier warp per epilogue group

__global__ void __launch_bounds__(kChainThreads, 1)
conv3x3_chain_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                     const __grid_constant__ CUtensorMap scratch_map, const __grid_constant__ CUtensorMap scratch_map_q,
                     const __grid_constant__ CUtensorMap out_map_q, const __grid_constant__ CUtensorMap out_map_e,
                     const __grid_constant__ ChainParams p) {
    constexpr int NG = 64;
    constexpr int kStages = ring_stages(NG, false, false);
    constexpr int kRowsDx = rows_per_dx(NG, false);
    constexpr int kWBytes = w_smem_bytes(NG, false);
    constexpr int kBank = 3 * NG;
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
    const int C = p.len;
    const int j = static_cast<int>(blockIdx.x) % C;          // layer of the chain this CTA computes
    const unsigned chain = blockIdx.x / C;
    const unsigned n_chains = gridDim.x / C;
    const bool first = (j == 0), last = (j == C - 1);
    const int ext = C - 1 - j;
    const int P = kBoxPx - 2 * C;                            // output pixels per strip of the chain
    const int CH = p.canvas_h;
    const bool rev = p.reverse != 0;

    if (threadIdx.x == 0) {
        mbar_init(base + kBarW, 1);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(base + kBarAFull + 8 * s, 1);
            mbar_init(base + kBarAEmpty + 8 * s, 1);
        }
        for (int s = 0; s < 6; ++s) {
            mbar_init(base + kBarAccFull + 8 * s, 1);
            mbar_init(base + kBarAccEmpty + 8 * s, 4);
        }
        for (int s = 0; s < 2; ++s) {
#if REVE_STAGE_QUARTERS
            for (int q = 0; q < 4; ++q) {               // [group s][quarter q]: one arrive by the warp / by the courier
                mbar_init(base + kBarStgFull + 8 * (4 * s + q), 1);
                mbar_init(base + kBarStgFree + 8 * (4 * s + q), 1);
            }
#else
            mbar_init(base + kBarStgFull + 8 * s, 4);   // one arrive per warp of the group
            mbar_init(base + kBarStgFree + 8 * s, 1);
#endif
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(base + kTmemPtr, kTmemCols);
        tmem_relinquish();
    }
    if (warp == 0 && lane == 0) {
        if (first) prefetch_tmap(&in_map);
        if (last) { prefetch_tmap(&out_map); prefetch_tmap(&out_map_q); prefetch_tmap(&out_map_e); }
        if (C > 1) { prefetch_tmap(&scratch_map); prefetch_tmap(&scratch_map_q); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + kTmemPtr);

    // the chain's block of strip-rows, cut into the two streams
    RowSpace rspace;
    rspace.rowmap = p.rowmap; rspace.run_fwd = p.run_fwd; rspace.run_bwd = p.run_bwd;
    rspace.nr = p.n_rows; rspace.ch = CH; rspace.rev = rev;
    long long lo[2], hi[2];
    {
        const unsigned b = rev ? (n_chains - 1 - chain) : chain;
        long long wlo = static_cast<long long>(b) * p.total_rows / n_chains;
        long long whi = static_cast<long long>(b + 1) * p.total_rows / n_chains;
        if (p.speed_in) {
            // block position q is worked by chain q (forward) or n_chains-1-q (reverse); every CTA runs the same
            // additions in the same order, so neighbouring chains agree on their common boundary to the bit
            float sum = 0.f, below = 0.f, upto = 0.f;
            for (unsigned q = 0; q < n_chains; ++q) {
                float wq = p.speed_in[rev ? n_chains - 1 - q : q];
                if (!(wq > 0.25f && wq < 4.f)) wq = 1.f;
                if (q == b) below = sum;
                sum += wq;
                if (q == b) upto = sum;
            }
            wlo = static_cast<long long>(static_cast<double>(p.total_rows) * static_cast<double>(below) / static_cast<double>(sum));
            whi = (b + 1 == n_chains) ? static_cast<long long>(p.total_rows)
                                      : static_cast<long long>(static_cast<double>(p.total_rows) * static_cast<double>(upto) / static_cast<double>(sum));
        }
        lo[0] = wlo;
        hi[0] = lo[1] = wlo + (whi - wlo) / 2;
        hi[1] = whi;
    }
    int U[2];
    U[0] = stream_steps(rspace, lo[0], hi[0], ext);
    U[1] = stream_steps(rspace, lo[1], hi[1], ext);

    // links: in = (j-1 -> j), out = (j -> j+1); per link and stream: kChainSlots scratch slots and two counters
    const unsigned link_in = chain * (C - 1) + (j - 1), link_out = chain * (C - 1) + j;

    if (warp >= 2 && warp < 10) {
        const int grp = (warp - 2) >> 2;
        const uint32_t t = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + grp * kBank;
#pragma unroll
        for (int c = 0; c < kBank / 16; ++c) tmem_st16_fill(t + c * 16, 0u);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp >= 10) {
        // ------------------------------------------------------------------ courier of epilogue group warp - 10
        if (lane == 0 && !last) {
            const int grp = warp - 10;
            const int n_rows_out = stream_steps(rspace, lo[grp], hi[grp], ext - 1);   // = the next layer's steps
            unsigned int* const pub_flag = p.flags + (link_out * 2 + 0) * kChainFlagStride + grp;
            const unsigned int* const cons_flag = p.flags + (link_out * 2 + 1) * kChainFlagStride + grp;
            const int slot_base = static_cast<int>((link_out * 2 + grp) * kChainSlots);
            const uint32_t stg = base + kOffStage + grp * kRowBytes;
            int cons_seen = 0, published = 0;
            long long* const tr = (p.trace && chain == static_cast<unsigned>(p.trace_chain) && grp == 0) ? p.trace + j * 512 + 500 : nullptr;
            long long t_full = 0, t_store = 0, t_pub = 0;
            // Announce rows 1..n of this stream to the next layer.  The rows were written by the async proxy (TMA);
            // wait_group 0 makes them visible to this thread, the gpu-scope RELEASE store makes them visible to whoever
            // acquires the counter.  (A relaxed store here loses the race about once in 1000 frames: the next layer then
            // reads a partly stale ring slot -- tools/race_hunt.py.)  The release costs ~1300 cycles of this thread's
            // time (a row period is ~2500, the store itself ~1000), so rows are announced kChainPublishEvery at a time,
            // never while the next row is already waiting, and always before the courier blocks on `consumed`.
            auto publish = [&](int n) {
                if (n > published) {
                    bulk_wait<0>();
                    fence_proxy_async_global();
                    st_release_gpu(pub_flag, static_cast<unsigned>(n));
                    published = n;
                }
            };
            for (int r = 0; r < n_rows_out; ++r) {
                const long long c0 = tr ? clock64() : 0;
#if REVE_STAGE_QUARTERS
                const uint32_t full0 = base + kBarStgFull + 8 * (4 * grp), free0 = base + kBarStgFree + 8 * (4 * grp);
                mbar_wait(full0, r & 1, dbg, TAG_STG_FULL, r);
#else
                mbar_wait(base + kBarStgFull + 8 * grp, r & 1, dbg, TAG_STG_FULL, r);
#endif
                const long long c1 = tr ? clock64() : 0;
                // slot (r mod kChainSlots) last held row r - kChainSlots of this stream (rows are published 1-based)
                if (cons_seen < r + 1 - kChainSlots) {
                    publish(r);                                  // everything stored so far, before waiting for the consumer
                    flag_wait_ge(cons_flag, r + 1 - kChainSlots, cons_seen, dbg, TAG_CHAIN_CONS);
                    fence_proxy_async_global();
                }
#if REVE_STAGE_QUARTERS
                // one store per quarter as soon as its warp has written it; a quarter is handed back when ITS read-out is done
                const int row0 = (slot_base + r % kChainSlots) * kBoxPx;
                tma_store_2d(&scratch_map_q, stg, 0, row0);
                bulk_commit();
#pragma unroll
                for (int q = 1; q < 4; ++q) {
                    mbar_wait(full0 + 8 * q, r & 1, dbg, TAG_STG_FULL, r);
                    tma_store_2d(&scratch_map_q, stg + q * 4096, 0, row0 + 32 * q);
                    bulk_commit();
                }
                bulk_wait_read<3>();
                mbar_arrive(free0);
                bulk_wait_read<2>();
                mbar_arrive(free0 + 8);
                bulk_wait_read<1>();
                mbar_arrive(free0 + 16);
                bulk_wait_read<0>();
                mbar_arrive(free0 + 24);
                const long long c2 = tr ? clock64() : 0;
                if (r + 1 - published >= kChainPublishEvery && !(r + 1 < n_rows_out && mbar_test_wait(full0, (r + 1) & 1)))
                    publish(r + 1);
#else
                tma_store_2d(&scratch_map, stg, 0, (slot_base + r % kChainSlots) * kBoxPx);
                bulk_commit();
                bulk_wait_read<0>();
                mbar_arrive(base + kBarStgFree + 8 * grp);
                const long long c2 = tr ? clock64() : 0;
                if (r + 1 - published >= kChainPublishEvery &&
                    !(r + 1 < n_rows_out && mbar_test_wait(base + kBarStgFull + 8 * grp, (r + 1) & 1)))
                    publish(r + 1);
#endif
                if (tr) {
                    const long long c3 = clock64();
                    t_full += c1 - c0; t_store += c2 - c1; t_pub += c3 - c2;
                }
            }
            publish(n_rows_out);
            if (tr) { tr[0] = t_full; tr[1] = t_store; tr[2] = t_pub; tr[3] = n_rows_out; }
        }
    } else if (warp == 0) {
        // ------------------------------------------------------------------ loader
        if (lane == 0) {
            mbar_arrive_expect_tx(base + kBarW, kWBytes);
            bulk_load_1d(base + kOffW, p.weights[j], kWBytes, base + kBarW);
            ChainCursor cur0(rspace, lo[0], hi[0], ext), cur1(rspace, lo[1], hi[1], ext);
            Sequencer seq(U[0], U[1]);
            int pub0 = 0, pub1 = 0;                 // last value seen of the two `published` counters
            int ps_a = 0, pk_a = 0, ps_b = 0, pk_b = 0;   // (stream, row) of the loads issued one and two steps ago
            long long* const tr = (p.trace && chain == static_cast<unsigned>(p.trace_chain)) ? p.trace + j * 512 + 504 : nullptr;
            long long t_poll = 0, t_empty = 0, t_retire = 0;
            uint32_t i = 0;
            while (seq.next()) {
                const uint32_t stage = i % kStages, use = i / kStages;
                const long long c0 = tr ? clock64() : 0;
                if (first) {
                    int strip, y;
                    bool newseg;
                    if (seq.s == 0) cur0.next(strip, y, newseg); else cur1.next(strip, y, newseg);
                    int x = (rev ? p.n_strips - 1 - strip : strip) * P - C;
                    mbar_wait(base + kBarAEmpty + 8 * stage, (use & 1) ^ 1, dbg, TAG_A_EMPTY, i);
                    mbar_arrive_expect_tx(base + kBarAFull + 8 * stage, kRowBytes);
                    tma_load_3d_hint(base + kOffRing + stage * kRowBytes, &in_map, base + kBarAFull + 8 * stage, 0,
                                     x, rev ? CH - 1 - y : y, kPolicyEvictFirst);
                } else {
                    if (i >= 2) {   // the load issued two steps ago has landed: its scratch slot may be overwritten
                        const uint32_t o = i - 2;
                        mbar_wait(base + kBarAFull + 8 * (o % kStages), (o / kStages) & 1, dbg, TAG_A_FULL, o);
                        // The slot's READ (async proxy) completed before this thread's acquire on the mbarrier returned, and
                        // this store follows that wait in program order.  REVE_CONSUMED_RELEASE makes it a gpu-scope release,
                        // which puts "read done" -> "slot may be overwritten" into the PTX causality order by the letter
                        // (DESIGN.md section 4.1c); the relaxed form relies on a completed read being immune to later writes.
#if REVE_CONSUMED_RELEASE
                        st_release_gpu(p.flags + (link_in * 2 + 1) * kChainFlagStride + ps_b, static_cast<unsigned>(pk_b));
#else
                        st_relaxed_gpu(p.flags + (link_in * 2 + 1) * kChainFlagStride + ps_b, static_cast<unsigned>(pk_b));
#endif
                    }
                    const long long c1 = tr ? clock64() : 0;
                    const int s = seq.s, k = seq.k;
                    // one 64-bit load refreshes both streams' counters
                    bool polled = false;
                    // (test hook p.fault: chain 0's second layer waits for a row that is never announced, so that the
                    // watchdog path -- diagnostic word, trap, REVE_E_CUDA at the ABI -- can be exercised on purpose)
                    const int want = (p.fault && chain == 0 && j == 1) ? k + (1 << 24) : k;
                    if ((s == 0 ? pub0 : pub1) < want) {
                        const unsigned int* const pf = p.flags + (link_in * 2 + 0) * kChainFlagStride;
                        const long long t0 = clock64();
                        do {
                            ld_acquire_gpu_v2(pf, pub0, pub1);
                            if (clock64() - t0 > (1ll << 32)) watchdog_fail(dbg, TAG_CHAIN_PUB, static_cast<uint32_t>(k), static_cast<uint32_t>(s == 0 ? pub0 : pub1));
                        } while ((s == 0 ? pub0 : pub1) < want);
                        polled = true;
                    }
                    if (polled) fence_proxy_async_global();
                    const long long c2 = tr ? clock64() : 0;
                    mbar_wait(base + kBarAEmpty + 8 * stage, (use & 1) ^ 1, dbg, TAG_A_EMPTY, i);
                    if (tr) { const long long c3 = clock64(); t_retire += c1 - c0; t_poll += c2 - c1; t_empty += c3 - c2; }
                    mbar_arrive_expect_tx(base + kBarAFull + 8 * stage, kRowBytes);
                    tma_load_2d(base + kOffRing + stage * kRowBytes, &scratch_map, base + kBarAFull + 8 * stage, 0,
                                static_cast<int>(((link_in * 2 + s) * kChainSlots + (k - 1) % kChainSlots) * kBoxPx));
                    ps_b = ps_a; pk_b = pk_a;
                    ps_a = s; pk_a = k;
                }
                ++i;
            }
            if (tr) { tr[0] = t_retire; tr[1] = t_poll; tr[2] = t_empty; tr[3] = i; }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (as in conv3x3_umma_kernel)
        mbar_wait(base + kBarW, 0, dbg, TAG_W);
        tc_fence_after();
        const uint32_t idesc = umma_idesc_f16(128, 3 * NG);
        const uint64_t proto = umma_desc_sw128(0, 0);
        const uint32_t desc_hi = static_cast<uint32_t>(proto >> 32);
        const uint32_t lo_flags = static_cast<uint32_t>(proto);
        const uint32_t w_lo = lo_flags | ((base + kOffW) >> 4);
        const uint32_t ring_lo = lo_flags | ((base + kOffRing) >> 4);
        constexpr uint32_t kDx = kRowsDx * 8;
        constexpr uint32_t kRot = NG * 8;
        struct Gate { uint32_t bar_f, par_f, bar_e, par_e; bool need_e; };
        auto gate_of = [&](uint32_t i, int s, int k) {
            Gate g;
            g.bar_f = base + kBarAFull + 8 * (i % kStages);
            g.par_f = (i / kStages) & 1;
            g.need_e = k >= 2;
            g.bar_e = base + kBarAccEmpty + 8 * (s * 3 + (k + 1) % 3);
            g.par_e = ((k - 2) / 3) & 1;
            return g;
        };
        Sequencer seq(U[0], U[1]);
        uint32_t i = 0;
        long long* const cta_tr = (p.trace && lane == 0) ? p.trace + 2048 + blockIdx.x * 4 : nullptr;   // per-CTA: ns
        if (cta_tr) cta_tr[0] = static_cast<long long>(globaltimer_ns());
        const bool clocked = last && p.speed_out != nullptr;     // the chain's last layer times the chain (see ChainParams)
        unsigned long long t_first = 0;
        bool have = seq.next();
        if (have) {
            const Gate g = gate_of(0, seq.s, seq.k);
            mbar_wait(g.bar_f, g.par_f, dbg, TAG_A_FULL, 0);
            tc_fence_after();
        }
        if (cta_tr) cta_tr[1] = static_cast<long long>(globaltimer_ns());
        if (clocked) t_first = globaltimer_ns();
        const bool elected = elect_one();
        while (have) {
            const int s = seq.s, k = seq.k;
            const uint32_t stage = i % kStages;
            const uint32_t a_lo = ring_lo + stage * (kRowBytes >> 4);
            uint32_t w_row = w_lo + (2 - (k + 1) % 3) * kRot;
            asm volatile("" : "+r"(w_row));
            const uint32_t d = tmem_base + s * kBank;
            if (p.trace && chain == static_cast<unsigned>(p.trace_chain) && i >= 200 && i < 684 && lane == 0) p.trace[j * 512 + i - 200] = clock64();
            if (elected) {
#pragma unroll
                for (int dxk = 0; dxk < 4; ++dxk)
                    umma_f16(d, mk_desc(desc_hi, a_lo - 8 + dxk * 2), mk_desc(desc_hi, w_row + dxk * 2), idesc, 1u);
            }
            have = seq.next();
            Gate g{};
            bool ok = true;
            if (have) {
                g = gate_of(i + 1, seq.s, seq.k);
                ok = mbar_test_wait(g.bar_f, g.par_f);
                if (g.need_e) ok = mbar_test_wait(g.bar_e, g.par_e) && ok;
            }
            if (elected) {
#pragma unroll
                for (int dxk = 4; dxk < 12; ++dxk) {
                    const int dx = dxk >> 2, kk = dxk & 3;
                    umma_f16(d, mk_desc(desc_hi, a_lo + (dx - 1) * 8 + kk * 2), mk_desc(desc_hi, w_row + dx * kDx + kk * 2), idesc, 1u);
                }
                umma_commit(base + kBarAEmpty + 8 * stage);
                umma_commit(base + kBarAccFull + 8 * (s * 3 + (k - 1) % 3));
            }
            if (!ok) {
                // (trace: what the next step had to wait for -- its input row, or the accumulator slot the epilogue returns)
                const bool tw = p.trace && chain == static_cast<unsigned>(p.trace_chain) && lane == 0;
                const long long w0 = tw ? clock64() : 0;
                mbar_wait(g.bar_f, g.par_f, dbg, TAG_A_FULL, i + 1);
                const long long w1 = tw ? clock64() : 0;
                if (g.need_e) mbar_wait(g.bar_e, g.par_e, dbg, TAG_ACC_EMPTY, seq.k);
                if (tw) {
                    p.trace[j * 512 + 492] += w1 - w0;            // cycles waiting for A
                    p.trace[j * 512 + 493] += clock64() - w1;     // cycles waiting for an empty accumulator slot
                    p.trace[j * 512 + 494] += 1;                  // steps that had to wait at all
                }
            }
            tc_fence_after();
            ++i;
        }
        if (cta_tr) {
            cta_tr[2] = static_cast<long long>(globaltimer_ns());
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            cta_tr[3] = static_cast<long long>(i) | (static_cast<long long>(smid) << 32);
        }
        if (clocked && lane == 0) {
            // steps per microsecond of this chain (its slowest layer paces the last one), averaged with the figure the
            // split of this launch was based on; short launches carry no information and hand the old figure on
            float old = p.speed_in ? p.speed_in[chain] : 1.f;
            if (!(old > 0.25f && old < 4.f)) old = 1.f;
            const unsigned long long dt = globaltimer_ns() - t_first;
            float now = old;
            if (i >= 256u && dt > 0ull) {
                now = 1000.f * static_cast<float>(i) / static_cast<float>(dt);
                now = fminf(fmaxf(now, 0.5f), 2.5f);
                if (p.speed_in && old != 1.f) now = 0.5f * (old + now);
            }
            p.speed_out[chain] = now;
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue warps
        const int grp = (warp - 2) >> 2;
        const int q = warp & 3;
        const int m = q * 32 + lane;       // pixel of the chain's 128-pixel box
        const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + grp * kBank;
        ChainCursor cur(rspace, lo[grp], hi[grp], ext);
        const int n_events = U[grp];
        const int mlo = 1 + j, mhi = kBoxPx - 2 - j;     // valid output pixels of this layer
        [[maybe_unused]] const bool gleader = (q == 0 && lane == 0);   // (REVE_LAST_QUARTERS = 0 path)
        const float* const bias = p.bias[j];
        const __half2* const slope2 = p.slope2[j];
        int sent = 0;
        int seg_xb = 0;
        bool seg_colok = false;
        long long ta_acc = 0, ta_drain = 0, ta_wait = 0, ta_out = 0, ta_n = 0;   // trace accumulators (registers)
        for (int e = 0; e < n_events; ++e) {
            bool valid = false, keep = false;
            int pr = 0;
            if (e >= 1) {
                int strip, y;
                bool newseg;
                valid = cur.next(strip, y, newseg);
                if (newseg) {
                    seg_xb = (rev ? p.n_strips - 1 - strip : strip) * P - C;
                    const int cx = seg_xb + m;
                    const bool inside = (m >= mlo) && (m <= mhi) && (cx >= 0) && (cx < p.canvas_w);
                    seg_colok = inside && (p.colflag[cx] != 0);
                }
                pr = rev ? CH - 1 - y : y;
                if (valid) keep = seg_colok && (pr >= 0) && (pr < CH) && (p.rowflag[pr] != 0);
            }
            const int slot = e % 3;
            // (trace: where warp 0 of group 0 of the traced chain spends an event)
            long long* const etr = (p.trace && chain == static_cast<unsigned>(p.trace_chain) && grp == 0 && q == 0 && lane == 0)
                                       ? p.trace + j * 512 + 484 : nullptr;
            const long long e0 = etr ? clock64() : 0;
            mbar_wait(base + kBarAccFull + 8 * (grp * 3 + slot), (e / 3) & 1, dbg, TAG_ACC_FULL, e);
            const long long e1 = etr ? clock64() : 0;
            tc_fence_after();
            uint32_t acc[NG];
            if (valid) {
#pragma unroll
                for (int c = 0; c < NG / 16; ++c) {
                    uint32_t(&dst)[16] = *reinterpret_cast<uint32_t(*)[16]>(&acc[c * 16]);
                    tmem_ld16(tmem_lane + slot * NG + c * 16, dst);
                }
                tmem_wait_ld();
            }
#pragma unroll
            for (int c = 0; c < NG / 16; ++c) tmem_st16_fill(tmem_lane + slot * NG + c * 16, 0u);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(base + kBarAccEmpty + 8 * (grp * 3 + slot));
            const long long e2 = etr ? clock64() : 0;
            if (etr) { ta_acc += e1 - e0; ta_drain += e2 - e1; ta_n += 1; }
            if (!valid) continue;

            const uint32_t stg = base + kOffStage + grp * kRowBytes;
            if (last) {
#if REVE_LAST_QUARTERS
                if (lane == 0) bulk_wait_read<0>();        // this warp's previous store has read its quarter out
                __syncwarp();
#else
                if (gleader) bulk_wait_read<0>();          // the previous row has left the staging buffer
                named_bar_sync(1 + grp, 128);
#endif
            } else if (sent > 0) {
#if REVE_STAGE_QUARTERS
                mbar_wait(base + kBarStgFree + 8 * (4 * grp + q), (sent - 1) & 1, dbg, TAG_STG_FREE, sent);   // this warp's quarter of row sent-1 has been read out
#else
                mbar_wait(base + kBarStgFree + 8 * grp, (sent - 1) & 1, dbg, TAG_STG_FREE, sent);   // courier has read row sent-1
#endif
            }
            // last layer: box rows mlo..mhi become staging rows 0..P-1 (the stored box); other layers hand on the
            // whole 128-row tile in place (rows outside the valid range as zeros)
            const long long e3 = etr ? clock64() : 0;
            if (etr) ta_wait += e3 - e2;                       // waiting for the staging quarter's read-out
            const bool writes = last ? (m >= mlo && m <= mhi) : true;
            if (writes) {
#if REVE_LAST_QUARTERS
                const int row = m;                         // every layer stages its pixels in place
#else
                const int row = last ? m - mlo : m;
#endif
                const uint32_t rbase = stg + row * 128;
                if (keep) {
#pragma unroll
                    for (int c8 = 0; c8 < 8; ++c8) {
                        uint32_t pk[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int ch = c8 * 8 + jj * 2;
                            const __half2 v = __floats2half2_rn(__uint_as_float(acc[ch]) + bias[ch],
                                                                __uint_as_float(acc[ch + 1]) + bias[ch + 1]);
                            const __half2 z = __float2half2_rn(0.f);
                            const __half2 r = __hfma2(slope2[ch >> 1], __hmin2_nan(v, z), __hmax2_nan(v, z));
                            pk[jj] = *reinterpret_cast<const uint32_t*>(&r);
                        }
                        st_shared_v4(rbase + ((c8 ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
                    }
                } else {   // same chunk rotation as above: 8 consecutive rows hit 8 different bank groups (a plain
                           // c8 << 4 is an 8-way bank conflict, measured +13 % per row in the right-edge strip)
#pragma unroll
                    for (int c8 = 0; c8 < 8; ++c8) st_shared_v4(rbase + ((c8 ^ (row & 7)) << 4), 0u, 0u, 0u, 0u);
                }
            }
            fence_proxy_async_smem();
            if (last) {
#if REVE_LAST_QUARTERS
                __syncwarp();
                if (lane == 0) {
                    // pixels [32q, 32q+32) of the box, clipped to the layer's valid range [mlo, mhi]: the two end quarters
                    // are 32 - C pixels wide (mlo = C, mhi = 127 - C)
                    const int m0 = (q == 0) ? mlo : 32 * q;
                    tma_store_3d((q == 0 || q == 3) ? &out_map_e : &out_map_q, stg + m0 * 128, 0, seg_xb + m0, pr);
                    bulk_commit();
                }
#else
                named_bar_sync(1 + grp, 128);
                if (gleader) {
                    tma_store_3d(&out_map, stg, 0, seg_xb + C, pr);
                    bulk_commit();
                }
#endif
            } else {
                __syncwarp();
#if REVE_STAGE_QUARTERS
                if (lane == 0) mbar_arrive(base + kBarStgFull + 8 * (4 * grp + q));
#else
                if (lane == 0) mbar_arrive(base + kBarStgFull + 8 * grp);
#endif
            }
            if (etr) ta_out += clock64() - e3;                 // bias / PReLU / pack / staging writes / fence / hand-off
            ++sent;
        }
        if (p.trace && chain == static_cast<unsigned>(p.trace_chain) && grp == 0 && q == 0 && lane == 0) {
            long long* const t = p.trace + j * 512 + 484;
            t[0] = ta_acc; t[1] = ta_drain; t[2] = ta_wait; t[3] = ta_out; t[4] = ta_n;
        }
    }

#if REVE_LAST_QUARTERS
    if (last && warp >= 2 && warp < 10 && lane == 0) bulk_wait<0>();
#else
    if (last && warp >= 2 && warp < 10 && (warp & 3) == 0 && lane == 0) bulk_wait<0>();
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <int NG, bool TAIL, bool PAIR>
constexpr size_t smem_bytes_t() {
    return 1024 /*alignment slack*/ + kCtrlBytes + w_smem_bytes(NG, PAIR) + kGuard +
           ring_stages(NG, TAIL, PAIR) * kRowBytes + kGuard + (TAIL ? kTailRowTab * 4 + tail_out_stage_bytes(NG) : 2 * kRowBytes);
}

template <int NG, bool TAIL, bool PAIR>
cudaError_t set_smem_attr() {
    return cudaFuncSetAttribute(conv3x3_umma_kernel<NG, TAIL, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem_bytes_t<NG, TAIL, PAIR>()));
}

}  // namespace

size_t conv_weight_blob_bytes(int ng) { return w_smem_bytes(ng, false); }

void pack_conv_weights(const float* w_oihw, int co, int ng, bool reverse, uint16_t* blob) {
    // blob[dx][n over 5 blocks of ng rows][ci] fp16 with the 128-byte swizzle applied per 128-byte row:
    // 16-byte chunk c of row n is stored at chunk (c ^ (n & 7)).  Block j holds weight group
    // g = (2 - j) mod 3, i.e. [W2|W1|W0|W2|W1], so the three cyclic rotations the rotating accumulator
    // bank needs are start offsets of 0, 1, 2 blocks.  Group g holds vertical tap ky = g (forward sweep)
    // or ky = 2 - g (reverse sweep).
    std::memset(blob, 0, conv_weight_blob_bytes(ng));
    for (int dx = 0; dx < 3; ++dx)
        for (int j = 0; j < 5; ++j)
            for (int o = 0; o < co; ++o) {
                const int g = ((2 - j) % 3 + 3) % 3;
                const int n = j * ng + o;
                const int ky = reverse ? 2 - g : g;
                for (int ci = 0; ci < 64; ++ci) {
                    const float v = w_oihw[((static_cast<size_t>(o) * 64 + ci) * 3 + ky) * 3 + dx];
                    const size_t byte = static_cast<size_t>(dx) * (5 * ng * 128) + static_cast<size_t>(n) * 128 +
                                        (((ci >> 3) ^ (n & 7)) << 4) + (ci & 7) * 2;
                    blob[byte / 2] = f32_to_f16(v);
                }
            }
}

cudaError_t conv_kernels_init() {
    cudaError_t e;
    if ((e = set_smem_attr<64, false, false>()) != cudaSuccess) return e;
    if ((e = set_smem_attr<64, false, true>()) != cudaSuccess) return e;
    if ((e = set_smem_attr<16, true, false>()) != cudaSuccess) return e;
    if ((e = set_smem_attr<32, true, false>()) != cudaSuccess) return e;
    if ((e = set_smem_attr<48, true, false>()) != cudaSuccess) return e;
    return cudaFuncSetAttribute(conv3x3_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem_bytes_t<64, false, false>()));
}

// The chained kernel's CTAs wait on each other (global-memory flags), so all of them must be resident at the same
// time.  That is a property of the launch, not an assumption: the kernel is only ever launched cooperatively (the
// driver starts a cooperative grid when, and only when, every CTA has its SM slot -- whatever else is running on the
// device: a second reve_ctx, the colour-conversion kernel, an MPS co-tenant), and a grid that cannot be co-resident
// at all is refused with cudaErrorCooperativeLaunchTooLarge instead of dead-locking.  chain_max_resident_ctas() is
// what fits on the current device; the context falls back to single-layer launches (no inter-CTA waits) below that.
int chain_max_resident_ctas(int sm_count) {
    int dev = 0, coop = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, conv3x3_chain_kernel, kChainThreads,
                                                      smem_bytes_t<64, false, false>()) != cudaSuccess)
        return 0;
    return per_sm * sm_count;
}

cudaError_t launch_conv_chain(cudaStream_t st, int grid, const CUtensorMap& in_map, const CUtensorMap& out_map,
                              const CUtensorMap& scratch_map, const CUtensorMap& scratch_map_q, const CUtensorMap& out_map_q,
                              const CUtensorMap& out_map_e, const ChainParams& p) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
    cfg.blockDim = dim3(kChainThreads, 1, 1);
    cfg.dynamicSmemBytes = smem_bytes_t<64, false, false>();
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, conv3x3_chain_kernel, in_map, out_map, scratch_map, scratch_map_q, out_map_q, out_map_e, p);
}

cudaError_t launch_conv_body(cudaStream_t st, int grid, bool pair, const CUtensorMap& in_map, const CUtensorMap& out_map_q,
                             const CUtensorMap& out_map_e, const ConvParams& p) {
    if (pair) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(static_cast<unsigned>(grid & ~1), 1, 1);
        cfg.blockDim = dim3(kConvThreads, 1, 1);
        cfg.dynamicSmemBytes = smem_bytes_t<64, false, true>();
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, conv3x3_umma_kernel<64, false, true>, in_map, out_map_q, out_map_e, p);
    }
    conv3x3_umma_kernel<64, false, false><<<grid, kConvThreads, smem_bytes_t<64, false, false>(), st>>>(in_map, out_map_q, out_map_e, p);
    return cudaGetLastError();
}

cudaError_t launch_conv_tail(cudaStream_t st, int grid, int scale, const CUtensorMap& in_map, const ConvParams& p) {
    switch (scale) {
        case 2: conv3x3_umma_kernel<16, true, false><<<grid, conv_threads(true), smem_bytes_t<16, true, false>(), st>>>(in_map, in_map, in_map, p); break;
        case 3: conv3x3_umma_kernel<32, true, false><<<grid, conv_threads(true), smem_bytes_t<32, true, false>(), st>>>(in_map, in_map, in_map, p); break;
        case 4: conv3x3_umma_kernel<48, true, false><<<grid, conv_threads(true), smem_bytes_t<48, true, false>(), st>>>(in_map, in_map, in_map, p); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace reve
