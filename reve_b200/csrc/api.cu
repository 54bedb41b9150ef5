// C ABI of libreve_cuda (include/reve_cuda.h): context, staging ring and the per-frame launch
// sequence.  This is what stands behind Video::upscale_segment in place of the spawned
// `realesrgan-ncnn-vulkan` process (reference reve-shared/src/lib.rs:129-155).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/reve_cuda.h"
#include "geometry.h"
#include "kernels.h"
#include "model.h"

using namespace reve;

namespace {

thread_local std::string g_last_error = "";

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Slot {
    uint8_t* d_in = nullptr;
    uint8_t* d_out = nullptr;
    cudaEvent_t ev_h2d = nullptr, ev_comp = nullptr, ev_done = nullptr;
    uint64_t tag = 0;
    uint8_t* host_out = nullptr;   // destination of the D2H copy issued when the batch is flushed
    size_t host_out_stride = 0;
};

struct ProfEvent {
    int kind;  // 0 conv0, 1 body, 2 tail, -1 frame start marker
    cudaEvent_t ev;
};

}  // namespace

struct reve_ctx {
    int device = 0;
    int sm_count = 0;
    Geometry g;
    int scale = 0;
    cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
    __half* act[2] = {nullptr, nullptr};
    size_t act_bytes = 0;
    uint8_t *d_colflag = nullptr, *d_rowflag = nullptr;
    int *d_srcx = nullptr, *d_srcy = nullptr, *d_outx = nullptr, *d_outy = nullptr, *d_rowframe = nullptr;
    uint32_t* d_rowpack = nullptr;
    // Needed-row lists per margin r (rows within r pixels of a tile's kept region), r < prepad; margins >= prepad
    // need every row.  One device array per margin: rowmap | run_fwd | run_bwd, each `cap` ints.  rows_n[r][n] =
    // entries that belong to the first n stacked frames.
    struct RowMap {
        int* d = nullptr;
        int cap = 0;
        int rows_n[kMaxBatch + 1] = {};
    };
    std::vector<RowMap> rowmaps;
    int batch = 1;        // frames stacked on the canvas per launch set (<= kMaxBatch)
    int frame_ch = 0;     // canvas rows of one frame (frames are frame_ch + 1 rows apart: one gap row)
    int n_strips = 0;
    std::vector<int> pending;  // ring slots whose H2D copy is issued but whose kernels are not yet enqueued
    void* d_wblob[kNumConv] = {};  // per layer: forward-sweep blob followed by the reverse-sweep blob
    CUtensorMap map_in[2], map_out[2], map_out_q[2];   // input boxes of 128 px; output boxes of 31 px (map_out) and 32 px (map_out_q)
    CUtensorMap map_flat0;   // act[0] as a flat [pixels][64] tensor (conv0 output tiles of 128 pixels)
    void* d_w0 = nullptr;    // conv0 B operand
    void* d_w0_rows = nullptr;   // conv0, row-streaming kernel: three per-tap B operands
    Conv0Params c0;
    ConvParams body[kNumBody];
    ConvParams tail;
    int out_format = REVE_FMT_RGB24;
    YuvCoeffs yuv = {};
    uint8_t* rgb_scratch[kMaxBatch] = {};   // tail output when the frame leaves as yuv420p10le
    int grid = 0;
    bool pair = false;    // body layers run as CTA pairs (tcgen05 cta_group::2)
    // chained body layers (ChainParams in kernels.h): chain_len layers per launch, 0 = one launch per layer
    reve_ctx_options opt = {};   // as given to reve_ctx_create_ex (zero = automatic)
    int chain_len = 0;
    bool chain_forced = false;   // layers_per_launch given: no small-batch fallback to single layers (tests)
    bool chain_refused = false;  // a cooperative launch was refused at run time: single layers from then on
    int chain_strips = 0;
    int n_chains = 0;
    __half* d_chain_scratch = nullptr;
    unsigned int* d_chain_flags = nullptr;
    size_t chain_flag_bytes = 0;
    // per-chain speed tables of the self-balancing split: one pair per chained launch of the frame (the launches differ in
    // sweep direction and needed rows, i.e. in which part of the canvas a chain works on), swapped every time it runs
    float* d_chain_speed = nullptr;   // [kNumBody / 2][2][64] floats
    unsigned chain_launches[kNumBody / 2] = {};
    CUtensorMap map_chain_scratch, map_chain_scratch_q, map_chain_out[2], map_chain_out_q[2], map_chain_out_e[2];
    ChainParams chain[kNumBody / 2];
    DebugBlock* dbg_host = nullptr;
    DebugBlock* dbg_dev = nullptr;
    long long* d_trace = nullptr;  // reve_ctx_options.trace: timeline of CTA 0 of body layer 5
    std::vector<Slot> ring;
    int head = 0, oldest = 0, inflight = 0;
    bool profiling = false;
    std::vector<ProfEvent> prof_events;
    reve_profile prof = {};
    mutable std::string err;
};

namespace {

int set_err(reve_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    g_last_error = msg;
    return code;
}

std::string cuda_msg(reve_ctx* ctx, const char* what, cudaError_t e) {
    std::string m = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    if (ctx && ctx->dbg_host && ctx->dbg_host->code != 0) {
        char buf[160];
        std::snprintf(buf, sizeof buf, " [kernel watchdog: wait tag %u timed out in block %u, aux0=%u aux1=%u]",
                      ctx->dbg_host->code, ctx->dbg_host->block, ctx->dbg_host->aux0, ctx->dbg_host->aux1);
        m += buf;
    }
    return m;
}

#define CK(ctx, call)                                                         \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return set_err(ctx, REVE_E_CUDA, cuda_msg(ctx, #call, e__)); \
    } while (0)

// Chained body layers (ChainParams in kernels.h): 4 or 2 layers per launch, whichever costs less.  A chain of L
// layers runs ~13 % (L = 4) / ~8 % (L = 2) faster per strip-row than L separate launches (measured at 1080p: the
// hand-over stays in L2 and the chip is power-bound), but its strips are 128 - 2L columns wide instead of 126, which
// can cost a whole extra strip.  reve_ctx_options.layers_per_launch overrides the choice at context creation.
int choose_chain_len(int canvas_w) {
    const double speed[3] = {1.0, 1.08, 1.13};
    const int lens[3] = {0, 2, 4};
    double best = 0;
    int len = 0;
    for (int i = 0; i < 3; ++i) {
        const int P = lens[i] ? kBoxPx - 2 * lens[i] : kStripPx;
        const double cost = ((canvas_w + P - 1) / P) / speed[i];
        if (i == 0 || cost < best) {
            best = cost;
            len = lens[i];
        }
    }
    return len;
}

int encode_map(reve_ctx* ctx, EncodeTiledFn enc, CUtensorMap* map, void* base, int cw, int ch, int box_px) {
    const cuuint64_t gdim[3] = {64, static_cast<cuuint64_t>(cw), static_cast<cuuint64_t>(ch)};
    cuuint64_t gstride[2] = {128, static_cast<cuuint64_t>(cw) * 128};
    // REVE_DBG_ALIAS_ROWS (timing experiment, results are garbage): canvas rows alias each other 16 pixels
    // apart, so a whole layer's activations stay in L2 -- the upper bound of any L2-residency scheme
    if (ctx->opt.debug_flags & REVE_DBG_ALIAS_ROWS) gstride[1] = 16 * 128;
    const cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_px), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[96];
        std::snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
        return set_err(ctx, REVE_E_CUDA, buf);
    }
    return REVE_OK;
}

template <typename T>
int upload(reve_ctx* ctx, T** dptr, const void* src, size_t bytes) {
    CK(ctx, cudaMalloc(reinterpret_cast<void**>(dptr), bytes));
    CK(ctx, cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice));
    return REVE_OK;
}

struct OutLayout {
    int w, h, cw, ch;            // output luma size, chroma plane size
    size_t rgb_row, rgb_bytes;   // packed RGB
    size_t y_row, c_row, yuv_bytes;   // packed planar yuv420p10le
};
OutLayout out_layout(const reve_ctx* ctx) {
    OutLayout o;
    o.w = ctx->g.in_w * ctx->scale;
    o.h = ctx->g.in_h * ctx->scale;
    o.cw = (o.w + 1) / 2;
    o.ch = (o.h + 1) / 2;
    o.rgb_row = static_cast<size_t>(o.w) * 3;
    o.rgb_bytes = o.rgb_row * o.h;
    o.c_row = static_cast<size_t>(o.cw) * 2;
    o.y_row = o.c_row * 2;       // >= 2*w and a multiple of 4: the chroma pitch is half the luma pitch
    o.yuv_bytes = o.y_row * o.h + 2 * o.c_row * o.ch;
    return o;
}

void prof_mark(reve_ctx* ctx, int kind) {
    if (!ctx->profiling) return;
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, ctx->s_comp);
    ctx->prof_events.push_back({kind, ev});
}

// Rows a layer has to compute when `margin` more convolutions follow it (upstream keeps
// only the centre of every padded tile, so a row `d` pixels outside the kept region matters only to layers with
// margin >= d).  n = frames stacked on the canvas, ch = its active rows.
void set_row_space(const reve_ctx* ctx, ConvParams& p, int margin, int n, int ch) {
    p.canvas_h = ch;
    if (margin < static_cast<int>(ctx->rowmaps.size()) && ctx->rowmaps[margin].d) {
        const reve_ctx::RowMap& rm = ctx->rowmaps[margin];
        p.n_rows = rm.rows_n[n];
        p.rowmap = rm.d;
        p.run_fwd = rm.d + rm.cap;
        p.run_bwd = rm.d + 2 * rm.cap;
    } else {
        p.n_rows = ch;
        p.rowmap = p.run_fwd = p.run_bwd = nullptr;
    }
    p.total_rows = ctx->n_strips * p.n_rows;
}

// Enqueue the 18 launches of one batch (n <= ctx->batch frames stacked on the canvas, separated by
// gap rows) on the compute stream.
int enqueue_batch(reve_ctx* ctx, int n, const uint8_t* const* d_in, long long in_stride, uint8_t* const* d_out,
                  long long out_stride, int stop_after_layers = kNumConv, int* last_buf = nullptr) {
    const int ch = n * (ctx->frame_ch + 1) - 1;   // active canvas rows (the trailing gap row is excluded)
    prof_mark(ctx, -1);
    Conv0Params c0 = ctx->c0;
    c0.canvas_h = ch;
    for (int f = 0; f < n; ++f) c0.src[f] = d_in[f];
    c0.src_stride = in_stride;
    {
        if (!(ctx->opt.debug_flags & REVE_DBG_CONV0_IM2COL)) {
            const long long rows0 = static_cast<long long>((c0.canvas_w + 127) / 128) * ch;
            CK(ctx, launch_conv0_rows(ctx->s_comp, static_cast<int>(rows0 < ctx->sm_count ? rows0 : ctx->sm_count), ctx->map_out_q[0], c0));
        } else {
            const long long tiles0 = (static_cast<long long>(c0.canvas_w) * ch + 127) / 128;
            CK(ctx, launch_conv0(ctx->s_comp, static_cast<int>(tiles0 < ctx->sm_count ? tiles0 : ctx->sm_count), ctx->map_flat0, c0));
        }
    }
    ctx->prof.launches_conv0++;
    prof_mark(ctx, 0);
    int cur = 0;   // canvas that holds the input of the next layer (conv0 wrote act[0])
    for (int k = 0; k < kNumBody && k + 1 < stop_after_layers;) {
        const int L = ctx->chain_len;
        // (a chain needs some rows per CTA to amortise its fill latency: tiny batches run layer by layer)
        if (L > 1 && !ctx->chain_refused && k % L == 0 && k + L < stop_after_layers &&
            (ctx->chain_forced || static_cast<long long>(ctx->chain_strips) * ch >= 32ll * ctx->n_chains)) {
            // layers k .. k+L-1 in one launch: canvas `cur` -> scratch rings (L2) -> canvas `cur ^ 1`
            ChainParams c = ctx->chain[k / L];
            ConvParams rows{};
            set_row_space(ctx, rows, kNumBody - (k + L - 1), n, ch);   // the chain's last layer decides the rows
            c.canvas_h = ch;
            c.n_rows = rows.n_rows;
            c.rowmap = rows.rowmap;
            c.run_fwd = rows.run_fwd;
            c.run_bwd = rows.run_bwd;
            c.total_rows = ctx->chain_strips * c.n_rows;
            const int chains = c.total_rows < ctx->n_chains ? c.total_rows : ctx->n_chains;
            CK(ctx, cudaMemsetAsync(ctx->d_chain_flags, 0, ctx->chain_flag_bytes, ctx->s_comp));
            // self-balancing split (ChainParams::speed_in): full grids with enough rows per chain to measure
            // (a forced launch structure -- tests, sanitizer runs -- takes the balanced path at any size: a short launch reads
            // the table and hands the old figures on)
            const bool balance = ctx->d_chain_speed && chains == ctx->n_chains && (c.total_rows >= 256 * chains || ctx->chain_forced);
            if (balance) {
                float* const pair = ctx->d_chain_speed + 128 * (k / L);
                c.speed_in = pair + 64 * (ctx->chain_launches[k / L] & 1u);
                c.speed_out = pair + 64 * ((ctx->chain_launches[k / L] + 1u) & 1u);
            }
            {
                const cudaError_t le = launch_conv_chain(ctx->s_comp, chains * L, ctx->map_in[cur], ctx->map_chain_out[cur ^ 1], ctx->map_chain_scratch, ctx->map_chain_scratch_q,
                                                         ctx->map_chain_out_q[cur ^ 1], ctx->map_chain_out_e[cur ^ 1], c);
                if (le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorLaunchOutOfResources) {
                    // the device (as partitioned right now) cannot hold the grid: nothing was launched; these layers
                    // and every later batch run as single-layer launches, which never wait for another CTA
                    (void)cudaGetLastError();
                    ctx->chain_refused = true;
                    continue;
                }
                CK(ctx, le);
            }
            if (balance) ctx->chain_launches[k / L]++;
            ctx->prof.launches_body++;
            ctx->prof.body_frames += n;
            ctx->prof.body_layer_frames += static_cast<uint64_t>(n) * L;
            prof_mark(ctx, 1);
            cur ^= 1;
            k += L;
            continue;
        }
        ConvParams b = ctx->body[k];
        // the sweep direction is baked into the layer's weight blob; the canvases alternate with every launch
        set_row_space(ctx, b, kNumBody - k, n, ch);   // body layer k is followed by 16 - k convolutions
        const int grid = b.total_rows < ctx->grid ? b.total_rows : ctx->grid;
        CK(ctx, launch_conv_body(ctx->s_comp, grid, ctx->pair && grid >= 2, ctx->map_in[cur], ctx->map_out_q[cur ^ 1], ctx->map_out[cur ^ 1], b));
        ctx->prof.launches_body++;
        ctx->prof.body_frames += n;
        ctx->prof.body_layer_frames += n;
        prof_mark(ctx, 1);
        cur ^= 1;
        ++k;
    }
    if (last_buf) *last_buf = cur;
    if (stop_after_layers >= kNumConv) {
        ConvParams t = ctx->tail;
        set_row_space(ctx, t, 0, n, ch);
        const int grid = t.total_rows < ctx->grid ? t.total_rows : ctx->grid;
        const bool yuv = ctx->out_format != REVE_FMT_RGB24;
        for (int f = 0; f < n; ++f) {
            t.src[f] = d_in[f];
            t.dst[f] = yuv ? ctx->rgb_scratch[f] : d_out[f];
        }
        t.src_stride = in_stride;
        t.dst_stride = out_stride;
        CK(ctx, launch_conv_tail(ctx->s_comp, grid, ctx->scale, ctx->map_in[cur], t));
        ctx->prof.launches_tail++;
        prof_mark(ctx, 2);
        if (yuv) {   // d_out[f] receives the packed planar frame: Y, then U, then V
            const OutLayout o = out_layout(ctx);
            for (int f = 0; f < n; ++f) {
                uint8_t* const y = d_out[f];
                uint8_t* const u = y + o.y_row * o.h;
                uint8_t* const v = u + o.c_row * o.ch;
                CK(ctx, launch_rgb_to_yuv420p10(ctx->s_comp, ctx->rgb_scratch[f], out_stride, o.w, o.h,
                                                reinterpret_cast<uint16_t*>(y), static_cast<long long>(o.y_row),
                                                reinterpret_cast<uint16_t*>(u), reinterpret_cast<uint16_t*>(v),
                                                static_cast<long long>(o.c_row), ctx->yuv));
                ctx->prof.launches_yuv++;
            }
        }
    }
    ctx->prof.frames += n;
    return REVE_OK;
}

// Kernels + D2H copies for every submitted frame whose compute is still pending.
int flush_pending(reve_ctx* ctx) {
    if (ctx->pending.empty()) return REVE_OK;
    const int n = static_cast<int>(ctx->pending.size());
    const size_t in_row = static_cast<size_t>(ctx->g.in_w) * 3, out_row = in_row * ctx->scale;
    const uint8_t* ins[kMaxBatch];
    uint8_t* outs[kMaxBatch];
    for (int f = 0; f < n; ++f) {
        Slot& s = ctx->ring[ctx->pending[f]];
        ins[f] = s.d_in;
        outs[f] = s.d_out;
        CK(ctx, cudaStreamWaitEvent(ctx->s_comp, s.ev_h2d, 0));
    }
    int rc = enqueue_batch(ctx, n, ins, static_cast<long long>(in_row), outs, static_cast<long long>(out_row));
    if (rc != REVE_OK) return rc;
    const int out_h = ctx->g.in_h * ctx->scale;
    CK(ctx, cudaEventRecord(ctx->ring[ctx->pending[0]].ev_comp, ctx->s_comp));
    CK(ctx, cudaStreamWaitEvent(ctx->s_d2h, ctx->ring[ctx->pending[0]].ev_comp, 0));
    const OutLayout o = out_layout(ctx);
    for (int f = 0; f < n; ++f) {
        Slot& s = ctx->ring[ctx->pending[f]];
        if (ctx->out_format == REVE_FMT_RGB24) {
            CK(ctx, cudaMemcpy2DAsync(s.host_out, s.host_out_stride, s.d_out, out_row, out_row, out_h, cudaMemcpyDeviceToHost, ctx->s_d2h));
        } else {   // three planes; the host chroma pitch is half the host luma pitch
            const size_t hs = s.host_out_stride, hc = hs / 2;
            uint8_t* const hu = s.host_out + hs * o.h;
            uint8_t* const hv = hu + hc * o.ch;
            const uint8_t* const du = s.d_out + o.y_row * o.h;
            const uint8_t* const dv = du + o.c_row * o.ch;
            CK(ctx, cudaMemcpy2DAsync(s.host_out, hs, s.d_out, o.y_row, static_cast<size_t>(o.w) * 2, o.h, cudaMemcpyDeviceToHost, ctx->s_d2h));
            CK(ctx, cudaMemcpy2DAsync(hu, hc, du, o.c_row, o.c_row, o.ch, cudaMemcpyDeviceToHost, ctx->s_d2h));
            CK(ctx, cudaMemcpy2DAsync(hv, hc, dv, o.c_row, o.c_row, o.ch, cudaMemcpyDeviceToHost, ctx->s_d2h));
        }
        CK(ctx, cudaEventRecord(s.ev_done, ctx->s_d2h));
    }
    ctx->pending.clear();
    return REVE_OK;
}

int collect_profile(reve_ctx* ctx) {
    CK(ctx, cudaStreamSynchronize(ctx->s_comp));
    for (size_t i = 1; i < ctx->prof_events.size(); ++i) {
        const ProfEvent& cur = ctx->prof_events[i];
        if (cur.kind < 0) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->prof_events[i - 1].ev, cur.ev) != cudaSuccess) continue;
        if (cur.kind == 0) ctx->prof.ms_conv0 += ms;
        if (cur.kind == 1) { ctx->prof.ms_body += ms; ctx->prof.timed_body++; }
        if (cur.kind == 2) { ctx->prof.ms_tail += ms; ctx->prof.timed_frames++; }
    }
    for (auto& pe : ctx->prof_events) cudaEventDestroy(pe.ev);
    ctx->prof_events.clear();
    return REVE_OK;
}

void destroy_ctx(reve_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->s_comp) cudaStreamSynchronize(ctx->s_comp);
    if (ctx->s_d2h) cudaStreamSynchronize(ctx->s_d2h);
    if (ctx->s_h2d) cudaStreamSynchronize(ctx->s_h2d);
    for (auto& pe : ctx->prof_events) cudaEventDestroy(pe.ev);
    for (auto& s : ctx->ring) {
        cudaFree(s.d_in);
        cudaFree(s.d_out);
        if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
        if (s.ev_comp) cudaEventDestroy(s.ev_comp);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
    }
    cudaFree(ctx->act[0]);
    cudaFree(ctx->act[1]);
    cudaFree(ctx->d_colflag);
    cudaFree(ctx->d_rowflag);
    cudaFree(ctx->d_srcx);
    cudaFree(ctx->d_srcy);
    cudaFree(ctx->d_outx);
    cudaFree(ctx->d_outy);
    cudaFree(ctx->d_rowframe);
    cudaFree(ctx->d_rowpack);
    for (auto& rm : ctx->rowmaps) cudaFree(rm.d);
    for (void* p : ctx->d_wblob) cudaFree(p);
    cudaFree(ctx->d_w0);
    cudaFree(ctx->d_w0_rows);
    cudaFree(ctx->d_trace);
    cudaFree(ctx->d_chain_scratch);
    cudaFree(ctx->d_chain_flags);
    cudaFree(ctx->d_chain_speed);
    for (uint8_t* p : ctx->rgb_scratch) cudaFree(p);
    if (ctx->dbg_host) cudaFreeHost(ctx->dbg_host);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_comp) cudaStreamDestroy(ctx->s_comp);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    delete ctx;
}

int create_ctx(reve_ctx* ctx, int device, const Model& m, int in_w, int in_h, int tile, int prepad, int ring_depth,
               const reve_ctx_options& opt) {
    ctx->opt = opt;
    if (opt.layers_per_launch != 0 && opt.layers_per_launch != 1 && opt.layers_per_launch != 2 && opt.layers_per_launch != 4)
        return set_err(ctx, REVE_E_INVAL, "layers_per_launch must be 0 (automatic), 1, 2 or 4");
    if (opt.max_batch < 0 || opt.max_batch > kMaxBatch) return set_err(ctx, REVE_E_INVAL, "max_batch must be within 0..4");
    if (opt.debug_grid < 0) return set_err(ctx, REVE_E_INVAL, "debug_grid must be >= 0");
    int ndev = 0;
    CK(ctx, cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return set_err(ctx, REVE_E_INVAL, "no such CUDA device");
    cudaDeviceProp prop;
    CK(ctx, cudaGetDeviceProperties(&prop, device));
    // the library embeds sm_100a SASS only (arch-specific tcgen05 code does not run on 10.1 / 10.3 either)
    if (prop.major != 10 || prop.minor != 0)
        return set_err(ctx, REVE_E_ARCH, std::string("device '") + prop.name + "' is compute capability " + std::to_string(prop.major) + "." +
                                             std::to_string(prop.minor) + ", not 10.0 (sm_100a); there is no fallback path");
    CK(ctx, cudaSetDevice(device));
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->scale = m.scale;
    std::string gerr;
    int rc = make_geometry(in_w, in_h, m.scale, tile, prepad, ctx->g, gerr);
    if (rc != REVE_OK) return set_err(ctx, rc, gerr);
    const Geometry& g = ctx->g;

    CK(ctx, conv_kernels_init());
    CK(ctx, conv0_kernel_init());
    CK(ctx, conv0_rows_kernel_init());
    CK(ctx, cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    CK(ctx, cudaStreamCreateWithFlags(&ctx->s_comp, cudaStreamNonBlocking));
    CK(ctx, cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    CK(ctx, cudaHostAlloc(reinterpret_cast<void**>(&ctx->dbg_host), sizeof(DebugBlock), cudaHostAllocMapped));
    std::memset(ctx->dbg_host, 0, sizeof(DebugBlock));
    CK(ctx, cudaHostGetDevicePointer(reinterpret_cast<void**>(&ctx->dbg_dev), ctx->dbg_host, 0));

    // batch size: frames stacked vertically on one canvas (one zero gap row between frames), bounded by
    // kMaxBatch, the ring depth and ~6 GB of activation memory
    const int fch = g.canvas_h();
    const int cw = g.canvas_w();
    ctx->frame_ch = fch;
    ctx->batch = kMaxBatch < ring_depth ? kMaxBatch : ring_depth;
    while (ctx->batch > 1 && 2.0 * cw * (static_cast<double>(ctx->batch) * (fch + 1)) * 128.0 > 6e9) --ctx->batch;
    if (opt.max_batch >= 1 && opt.max_batch < ctx->batch) ctx->batch = opt.max_batch;
    const int ch = ctx->batch * (fch + 1) - 1;   // canvas rows allocated

    // geometry tables (y tables repeated per stacked frame)
    std::vector<uint8_t> colflag(cw), rowflag(ch);
    std::vector<int> srcy(ch), outy(ch), rowframe(ch);
    for (int i = 0; i < cw; ++i) colflag[i] = g.x.src[i] >= 0;
    for (int r = 0; r < ch; ++r) {
        const int f = r / (fch + 1), i = r % (fch + 1);
        const bool gap = (i == fch);
        srcy[r] = gap ? -1 : g.y.src[i];
        outy[r] = gap ? -1 : g.y.out[i];
        rowframe[r] = (gap || g.y.src[i] < 0) ? -1 : f;
        rowflag[r] = srcy[r] >= 0;
    }
    if ((rc = upload(ctx, &ctx->d_colflag, colflag.data(), cw))) return rc;
    if ((rc = upload(ctx, &ctx->d_rowflag, rowflag.data(), ch))) return rc;
    if ((rc = upload(ctx, &ctx->d_srcx, g.x.src.data(), sizeof(int) * cw))) return rc;
    if ((rc = upload(ctx, &ctx->d_srcy, srcy.data(), sizeof(int) * ch))) return rc;
    if ((rc = upload(ctx, &ctx->d_outx, g.x.out.data(), sizeof(int) * cw))) return rc;
    if ((rc = upload(ctx, &ctx->d_outy, outy.data(), sizeof(int) * ch))) return rc;
    if ((rc = upload(ctx, &ctx->d_rowframe, rowframe.data(), sizeof(int) * ch))) return rc;
    if (in_h > 32000) return set_err(ctx, REVE_E_INVAL, "frames taller than 32000 rows are not supported");
    std::vector<uint32_t> rowpack(ch);
    for (int r = 0; r < ch; ++r)
        rowpack[r] = (static_cast<uint32_t>(outy[r] + 1) << 17) | (static_cast<uint32_t>(srcy[r] + 1) << 2) |
                     static_cast<uint32_t>(rowframe[r] < 0 ? 0 : rowframe[r]);
    if ((rc = upload(ctx, &ctx->d_rowpack, rowpack.data(), sizeof(uint32_t) * ch))) return rc;
    // needed-row lists (REVE_DBG_ALL_ROWS disables them: every layer computes every row)
    {
        const int n_margins = (opt.debug_flags & REVE_DBG_ALL_ROWS) ? 0 : prepad;
        ctx->rowmaps.resize(n_margins);
        for (int r = 0; r < n_margins; ++r) {
            std::vector<char> need(ch, 0);
            for (int b0 = 0; b0 < ch;) {            // bands = maximal runs of tile rows (srcy >= 0)
                if (srcy[b0] < 0) { ++b0; continue; }
                int b1 = b0;
                while (b1 < ch && srcy[b1] >= 0) ++b1;
                int k0 = b0, k1 = b1;               // kept rows of the band (outy >= 0) are contiguous
                while (k0 < b1 && outy[k0] < 0) ++k0;
                while (k1 > k0 && outy[k1 - 1] < 0) --k1;
                for (int i = std::max(b0, k0 - r); i < std::min(b1, k1 + r); ++i) need[i] = 1;
                b0 = b1;
            }
            std::vector<int> rows;
            for (int i = 0; i < ch; ++i) if (need[i]) rows.push_back(i);
            const int cap = static_cast<int>(rows.size());
            if (cap == 0 || cap == ch) continue;   // nothing to skip: the layer walks the whole canvas
            std::vector<int> tab(3 * static_cast<size_t>(cap));
            for (int i = cap - 1; i >= 0; --i)
                tab[cap + i] = (i + 1 < cap && rows[i + 1] == rows[i] + 1) ? tab[cap + i + 1] + 1 : 1;
            for (int i = 0; i < cap; ++i) {
                tab[i] = rows[i];
                tab[2 * cap + i] = (i > 0 && rows[i - 1] == rows[i] - 1) ? tab[2 * cap + i - 1] + 1 : 1;
            }
            reve_ctx::RowMap& rm = ctx->rowmaps[r];
            rm.cap = cap;
            for (int nf = 0; nf <= ctx->batch; ++nf) {
                const int lim = nf * (fch + 1) - 1;   // active canvas rows with nf frames stacked
                rm.rows_n[nf] = static_cast<int>(std::lower_bound(rows.begin(), rows.end(), lim) - rows.begin());
            }
            if ((rc = upload(ctx, &rm.d, tab.data(), sizeof(int) * tab.size()))) return rc;
        }
    }

    // activation canvases (ping-pong), zero-initialised
    ctx->act_bytes = static_cast<size_t>(cw) * ch * 64 * sizeof(__half);
    for (int i = 0; i < 2; ++i) {
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ctx->act[i]), ctx->act_bytes);
        if (e == cudaErrorMemoryAllocation) return set_err(ctx, REVE_E_NOMEM, "out of device memory for the activation canvas");
        CK(ctx, e);
        CK(ctx, cudaMemset(ctx->act[i], 0, ctx->act_bytes));
    }

    // tensor maps
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return set_err(ctx, REVE_E_CUDA, "driver does not export cuTensorMapEncodeTiled");
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
    for (int i = 0; i < 2; ++i) {
        if ((rc = encode_map(ctx, enc, &ctx->map_in[i], ctx->act[i], cw, ch, kBoxPx))) return rc;
        if ((rc = encode_map(ctx, enc, &ctx->map_out[i], ctx->act[i], cw, ch, 31))) return rc;
        if ((rc = encode_map(ctx, enc, &ctx->map_out_q[i], ctx->act[i], cw, ch, 32))) return rc;
    }

    {
        const cuuint64_t gdim[2] = {64, static_cast<cuuint64_t>(cw) * ch};
        const cuuint64_t gstride[1] = {128};
        const cuuint32_t box[2] = {64, 32};     // a quarter of a 128-pixel tile: one TMA store per epilogue warp
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&ctx->map_flat0, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ctx->act[0], gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return set_err(ctx, REVE_E_CUDA, "cuTensorMapEncodeTiled (flat map) failed");
    }

    // weights
    {
        std::vector<uint16_t> blob(conv0_weight_blob_bytes() / 2);
        pack_conv0_weights(m.conv[0].w.data(), blob.data());
        if ((rc = upload(ctx, &ctx->d_w0, blob.data(), blob.size() * 2))) return rc;
        std::vector<uint16_t> rows(conv0_rows_weight_blob_bytes() / 2);
        pack_conv0_rows_weights(m.conv[0].w.data(), rows.data());
        if ((rc = upload(ctx, &ctx->d_w0_rows, rows.data(), rows.size() * 2))) return rc;
    }
    const int tail_ng = m.scale == 2 ? 16 : (m.scale == 3 ? 32 : 48);
    for (int k = 1; k < kNumConv; ++k) {
        const int ng = (k == kNumConv - 1) ? tail_ng : 64;
        const size_t half = conv_weight_blob_bytes(ng) / 2;
        std::vector<uint16_t> blob(2 * half);
        pack_conv_weights(m.conv[k].w.data(), m.conv[k].out_ch, ng, false, blob.data());
        pack_conv_weights(m.conv[k].w.data(), m.conv[k].out_ch, ng, true, blob.data() + half);
        if ((rc = upload(ctx, &ctx->d_wblob[k], blob.data(), blob.size() * 2))) return rc;
    }

    // parameter blocks
    const uint32_t dflags = opt.debug_flags;
    Conv0Params& c0 = ctx->c0;
    c0 = Conv0Params{};
    c0.canvas_w = cw;
    c0.canvas_h = ch;
    c0.src_x = ctx->d_srcx;
    c0.src_y = ctx->d_srcy;
    c0.row_frame = ctx->d_rowframe;
    c0.weights = ctx->d_w0;
    c0.weights_rows = ctx->d_w0_rows;
    c0.dbg = ctx->dbg_dev;
    for (int co = 0; co < 64; ++co) {
        c0.bias[co] = m.conv[0].b[co];
        c0.slope[co] = m.conv[0].slope[co];
    }
    for (int co = 0; co < 64; co += 2) c0.slope2[co >> 1] = __floats2half2_rn(m.conv[0].slope[co], m.conv[0].slope[co + 1]);
    const int n_strips = (cw + kStripPx - 1) / kStripPx;
    ctx->n_strips = n_strips;
    const long long total = static_cast<long long>(n_strips) * ch;
    ctx->grid = ctx->sm_count;   // persistent: one CTA per SM (clamped to the work of a batch at launch)
    // Test hook: cap the persistent grid so one CTA walks many rows / strips.
    if (opt.debug_grid >= 1 && opt.debug_grid < ctx->grid) ctx->grid = opt.debug_grid;
    // REVE_DBG_NO_REVERSE: never sweep in reverse; REVE_DBG_CTA_PAIRS: CTA pairs (measured: same frames/s, the chip is
    // power-bound, see profiles/r01_notes.md); REVE_DBG_SWAP_PAIR_B: swap the pair's B halves.
    ctx->pair = (dflags & REVE_DBG_CTA_PAIRS) != 0;
    for (int k = 0; k <= kNumBody; ++k) {
        ConvParams& p = (k < kNumBody) ? ctx->body[k] : ctx->tail;
        p = ConvParams{};
        // Sweep direction per layer: it fixes the order in which the three vertical taps reach the fp32 accumulator,
        // so it must not depend on the launch structure (single layers, chains of 2 or 4) or results would differ in
        // the last bit between geometries.  Groups of four layers alternate: conv0 writes the canvas top-down, layers
        // 0..3 sweep bottom-up, 4..7 top-down, ...; the tail (k = 16) sweeps bottom-up again.
        p.reverse = (dflags & 1u) ? 0 : (((k >> 2) & 1) == 0);
        p.flags = (dflags & 4u) ? 1u : 0u;
        p.out = (k < kNumBody) ? ctx->act[(k + 1) & 1] : nullptr;
        p.canvas_w = cw;
        p.canvas_h = ch;
        p.n_strips = n_strips;
        p.total_rows = static_cast<int>(total);
        p.n_rows = ch;
        p.colflag = ctx->d_colflag;
        p.rowflag = ctx->d_rowflag;
        p.weights = static_cast<const uint8_t*>(ctx->d_wblob[k + 1]) +
                    (p.reverse ? conv_weight_blob_bytes(k < kNumBody ? 64 : tail_ng) : 0);
        p.dbg = ctx->dbg_dev;
        p.src_x = ctx->d_srcx;
        p.src_y = ctx->d_srcy;
        p.out_x = ctx->d_outx;
        p.out_y = ctx->d_outy;
        p.row_frame = ctx->d_rowframe;
        p.rowpack = ctx->d_rowpack;
        const ConvLayer& L = m.conv[k + 1];
        for (int c = 0; c < L.out_ch; ++c) {
            p.bias[c] = L.b[c];
            p.slope[c] = L.slope.empty() ? 0.f : L.slope[c];
        }
        if (!L.slope.empty())
            for (int c = 0; c < 64; c += 2) p.slope2[c >> 1] = __floats2half2_rn(L.slope[c], L.slope[c + 1]);
    }

    ctx->chain_len = choose_chain_len(cw);
    if (opt.layers_per_launch) {
        ctx->chain_len = opt.layers_per_launch == 1 ? 0 : opt.layers_per_launch;
        ctx->chain_forced = true;
    }
    if (opt.flags & REVE_CTX_SHARED_DEVICE) ctx->chain_len = 0;   // only kernels without inter-CTA waits
    if (ctx->pair || ctx->grid < ctx->chain_len) ctx->chain_len = 0;
    // the chained kernel's CTAs wait for each other: it is launched cooperatively, which needs the whole grid resident
    if (ctx->chain_len && chain_max_resident_ctas(ctx->sm_count) < (ctx->grid / ctx->chain_len) * ctx->chain_len) ctx->chain_len = 0;
    if (ctx->chain_len) {
        const int L = ctx->chain_len;
        const int P = kBoxPx - 2 * L;
        ctx->chain_strips = (cw + P - 1) / P;
        ctx->n_chains = ctx->grid / L;
        const size_t rows = chain_scratch_rows(ctx->n_chains, L);
        CK(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_chain_scratch), rows * 128));
        CK(ctx, cudaMemset(ctx->d_chain_scratch, 0, rows * 128));
        ctx->chain_flag_bytes = chain_flag_words(ctx->n_chains, L) * sizeof(unsigned int);
        CK(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_chain_flags), ctx->chain_flag_bytes));
        if (!(dflags & REVE_DBG_EQUAL_SPLIT) && ctx->n_chains <= 64) {
            const std::vector<float> ones(128 * (kNumBody / 2), 1.0f);
            if ((rc = upload(ctx, &ctx->d_chain_speed, ones.data(), ones.size() * sizeof(float)))) return rc;
        }
        {
            const cuuint64_t gdim[2] = {64, static_cast<cuuint64_t>(rows)};
            const cuuint64_t gstride[1] = {128};
            const cuuint32_t box[2] = {64, 128};
            const cuuint32_t estr[2] = {1, 1};
            const CUresult r = enc(&ctx->map_chain_scratch, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ctx->d_chain_scratch, gdim,
                                   gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return set_err(ctx, REVE_E_CUDA, "cuTensorMapEncodeTiled (chain scratch map) failed");
            const cuuint32_t box_q[2] = {64, 32};   // a quarter of a row: what one epilogue warp writes and its TMA store ships
            const CUresult rq = enc(&ctx->map_chain_scratch_q, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ctx->d_chain_scratch, gdim,
                                    gstride, box_q, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rq != CUDA_SUCCESS) return set_err(ctx, REVE_E_CUDA, "cuTensorMapEncodeTiled (chain scratch quarter map) failed");
        }
        for (int i = 0; i < 2; ++i) {
            if ((rc = encode_map(ctx, enc, &ctx->map_chain_out[i], ctx->act[i], cw, ch, P))) return rc;
            if ((rc = encode_map(ctx, enc, &ctx->map_chain_out_q[i], ctx->act[i], cw, ch, 32))) return rc;
            if ((rc = encode_map(ctx, enc, &ctx->map_chain_out_e[i], ctx->act[i], cw, ch, 32 - L))) return rc;
        }
        for (int c = 0; c < kNumBody / L; ++c) {
            ChainParams& p = ctx->chain[c];
            p = ChainParams{};
            p.canvas_w = cw;
            p.canvas_h = ch;
            p.n_strips = ctx->chain_strips;
            p.n_rows = ch;
            p.total_rows = ctx->chain_strips * ch;
            p.colflag = ctx->d_colflag;
            p.rowflag = ctx->d_rowflag;
            p.len = L;
            p.reverse = (dflags & 1u) ? 0 : ((((c * L) >> 2) & 1) == 0);   // the direction of its layers (see above)
            p.flags = ctx->d_chain_flags;
            p.dbg = ctx->dbg_dev;
            p.fault = (dflags & REVE_DBG_FAULT) ? 1u : 0u;
            for (int j = 0; j < L; ++j) {
                const int k = c * L + j;   // body layer
                const ConvLayer& Lr = m.conv[k + 1];
                p.weights[j] = static_cast<const uint8_t*>(ctx->d_wblob[k + 1]) + (p.reverse ? conv_weight_blob_bytes(64) : 0);
                for (int o = 0; o < 64; ++o) p.bias[j][o] = Lr.b[o];
                for (int o = 0; o < 64; o += 2) p.slope2[j][o >> 1] = __floats2half2_rn(Lr.slope[o], Lr.slope[o + 1]);
            }
        }
    }

    if (opt.trace) {
        CK(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_trace), 4096 * sizeof(long long)));
        CK(ctx, cudaMemset(ctx->d_trace, 0, 4096 * sizeof(long long)));
        if (opt.trace == 2) ctx->tail.trace = ctx->d_trace;
        else if (ctx->chain_len) {   // chain `trace_chain` of chained launch `trace_launch`
            int idx = opt.trace_launch;
            if (idx < 0 || idx >= kNumBody / ctx->chain_len) idx = 1;
            ctx->chain[idx].trace = ctx->d_trace;
            ctx->chain[idx].trace_chain = opt.trace_chain;
        }
        else ctx->body[5].trace = ctx->d_trace;
    }

    // staging ring
    ctx->ring.resize(ring_depth);
    const size_t in_bytes = static_cast<size_t>(in_w) * in_h * 3;
    const OutLayout ol = out_layout(ctx);
    const size_t out_bytes = ol.rgb_bytes > ol.yuv_bytes ? ol.rgb_bytes : ol.yuv_bytes;   // either output format
    for (auto& s : ctx->ring) {
        CK(ctx, cudaMalloc(reinterpret_cast<void**>(&s.d_in), in_bytes));
        CK(ctx, cudaMalloc(reinterpret_cast<void**>(&s.d_out), out_bytes));
        CK(ctx, cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming));
        CK(ctx, cudaEventCreateWithFlags(&s.ev_comp, cudaEventDisableTiming));
        CK(ctx, cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    }
    return REVE_OK;
}

}  // namespace

// ================================================================================ C ABI
extern "C" {

int reve_version(void) { return REVE_VERSION; }

const char* reve_strerror(int status) {
    switch (status) {
        case REVE_OK: return "ok";
        case REVE_E_INVAL: return "invalid argument";
        case REVE_E_NOMEM: return "out of memory";
        case REVE_E_CUDA: return "CUDA error";
        case REVE_E_IO: return "I/O error";
        case REVE_E_MODEL: return "not a realesr-animevideov3 model";
        case REVE_E_ARCH: return "device is not sm_100";
        case REVE_E_BUSY: return "submit ring full";
        case REVE_E_EMPTY: return "nothing in flight";
        default: return "unknown status";
    }
}

const char* reve_last_error(const reve_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int reve_device_count(int* n) {
    if (!n) return set_err(nullptr, REVE_E_INVAL, "n is NULL");
    *n = 0;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return set_err(nullptr, REVE_E_CUDA, cuda_msg(nullptr, "cudaGetDeviceCount", e));
    for (int d = 0; d < ndev; ++d) {
        int major = 0;
        int minor = -1;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10 &&
            cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, d) == cudaSuccess && minor == 0)
            ++*n;
    }
    return REVE_OK;
}

int reve_device_pci_bus_id(int device, char* buf, size_t cap) {
    if (!buf || cap < 16) return set_err(nullptr, REVE_E_INVAL, "buffer too small");
    buf[0] = 0;
    char tmp[32] = {};
    const cudaError_t e = cudaDeviceGetPCIBusId(tmp, sizeof tmp, device);
    if (e != cudaSuccess) return set_err(nullptr, e == cudaErrorInvalidDevice ? REVE_E_INVAL : REVE_E_CUDA, cuda_msg(nullptr, "cudaDeviceGetPCIBusId", e));
    for (size_t i = 0; tmp[i] && i + 1 < cap; ++i) {
        buf[i] = static_cast<char>(tmp[i] >= 'A' && tmp[i] <= 'F' ? tmp[i] - 'A' + 'a' : tmp[i]);   // sysfs uses lower case
        buf[i + 1] = 0;
    }
    return REVE_OK;
}

int reve_model_load_ncnn(const char* param_path, const char* bin_path, reve_model** out) {
    if (!param_path || !bin_path || !out) return set_err(nullptr, REVE_E_INVAL, "NULL argument");
    *out = nullptr;
    reve_model* m = new (std::nothrow) reve_model();
    if (!m) return set_err(nullptr, REVE_E_NOMEM, "out of host memory");
    std::string err;
    int rc;
    try {
        rc = model_load_ncnn(param_path, bin_path, m->m, err);
    } catch (const std::exception& e) {
        rc = REVE_E_NOMEM;
        err = e.what();
    }
    if (rc != REVE_OK) {
        delete m;
        return set_err(nullptr, rc, err);
    }
    *out = m;
    return REVE_OK;
}

int reve_model_random(int scale, uint64_t seed, reve_model** out) {
    if (!out) return set_err(nullptr, REVE_E_INVAL, "NULL argument");
    *out = nullptr;
    reve_model* m = new (std::nothrow) reve_model();
    if (!m) return set_err(nullptr, REVE_E_NOMEM, "out of host memory");
    std::string err;
    int rc;
    try {
        rc = model_random(scale, seed, m->m, err);
    } catch (const std::exception& e) {
        rc = REVE_E_NOMEM;
        err = e.what();
    }
    if (rc != REVE_OK) {
        delete m;
        return set_err(nullptr, rc, err);
    }
    *out = m;
    return REVE_OK;
}

int reve_model_from_arrays(int scale, const float* const* conv_w, const float* const* conv_b,
                           const float* const* prelu, reve_model** out) {
    if (!out) return set_err(nullptr, REVE_E_INVAL, "NULL argument");
    *out = nullptr;
    reve_model* m = new (std::nothrow) reve_model();
    if (!m) return set_err(nullptr, REVE_E_NOMEM, "out of host memory");
    std::string err;
    int rc;
    try {
        rc = model_from_arrays(scale, conv_w, conv_b, prelu, m->m, err);
    } catch (const std::exception& e) {
        rc = REVE_E_NOMEM;
        err = e.what();
    }
    if (rc != REVE_OK) {
        delete m;
        return set_err(nullptr, rc, err);
    }
    *out = m;
    return REVE_OK;
}

int reve_model_save_ncnn(const reve_model* m, const char* param_path, const char* bin_path, int fp16) {
    if (!m || !param_path || !bin_path) return set_err(nullptr, REVE_E_INVAL, "NULL argument");
    std::string err;
    int rc;
    try {
        rc = model_save_ncnn(m->m, param_path, bin_path, fp16 != 0, err);
    } catch (const std::exception& e) {
        rc = REVE_E_NOMEM;
        err = e.what();
    }
    return rc == REVE_OK ? rc : set_err(nullptr, rc, err);
}

int reve_model_info(const reve_model* m, int* scale, int* num_feat, int* num_conv) {
    if (!m) return set_err(nullptr, REVE_E_INVAL, "model is NULL");
    if (scale) *scale = m->m.scale;
    if (num_feat) *num_feat = kNumFeat;
    if (num_conv) *num_conv = kNumBody;
    return REVE_OK;
}

void reve_model_free(reve_model* m) { delete m; }

int reve_ctx_create(int device, const reve_model* m, int in_w, int in_h, int tile, int prepad, int ring_depth,
                    reve_ctx** out) {
    return reve_ctx_create_ex(device, m, in_w, in_h, tile, prepad, ring_depth, nullptr, out);
}

int reve_ctx_create_ex(int device, const reve_model* m, int in_w, int in_h, int tile, int prepad, int ring_depth,
                       const reve_ctx_options* opt_in, reve_ctx** out) {
    if (!out) return set_err(nullptr, REVE_E_INVAL, "out is NULL");
    *out = nullptr;
    if (!m) return set_err(nullptr, REVE_E_INVAL, "model is NULL");
    reve_ctx_options opt = {};
    if (opt_in) {
        if (opt_in->struct_size < 2 * sizeof(uint32_t) || opt_in->struct_size > sizeof(reve_ctx_options))
            return set_err(nullptr, REVE_E_INVAL, "reve_ctx_options.struct_size is not set (or from a newer header)");
        std::memcpy(&opt, opt_in, opt_in->struct_size);
    }
    opt.struct_size = sizeof(reve_ctx_options);
    if (ring_depth < 1 || ring_depth > 16) return set_err(nullptr, REVE_E_INVAL, "ring_depth must be within 1..16");
    reve_ctx* ctx = new (std::nothrow) reve_ctx();
    if (!ctx) return set_err(nullptr, REVE_E_NOMEM, "out of host memory");
    int rc;
    try {
        rc = create_ctx(ctx, device, m->m, in_w, in_h, tile, prepad, ring_depth, opt);
    } catch (const std::exception& e) {
        rc = set_err(ctx, REVE_E_NOMEM, e.what());
    }
    if (rc != REVE_OK) {
        g_last_error = ctx->err;
        destroy_ctx(ctx);
        return rc;
    }
    *out = ctx;
    return REVE_OK;
}

void reve_ctx_destroy(reve_ctx* ctx) { destroy_ctx(ctx); }

int reve_ctx_launch_info(const reve_ctx* ctx, int* layers_per_launch, int* batch, int* grid, int* cooperative) {
    if (!ctx) return set_err(nullptr, REVE_E_INVAL, "ctx is NULL");
    const bool chained = ctx->chain_len > 1 && !ctx->chain_refused;
    if (layers_per_launch) *layers_per_launch = chained ? ctx->chain_len : 1;
    if (batch) *batch = ctx->batch;
    if (grid) *grid = chained ? ctx->n_chains * ctx->chain_len : ctx->grid;
    if (cooperative) *cooperative = chained ? 1 : 0;
    return REVE_OK;
}

int reve_device_recover(int device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess && e != cudaErrorLaunchFailure) (void)cudaGetLastError();
    if (device < 0 || (ndev > 0 && device >= ndev)) return set_err(nullptr, REVE_E_INVAL, "no such CUDA device");
    (void)cudaSetDevice(device);     // fails on a dead context; cudaDeviceReset still tears it down
    (void)cudaGetLastError();
    e = cudaDeviceReset();
    if (e != cudaSuccess) return set_err(nullptr, REVE_E_CUDA, cuda_msg(nullptr, "cudaDeviceReset", e));
    (void)cudaGetLastError();
    // Whether a process may build a new primary context after a faulted one is the driver's decision: on the B200 pool
    // this was developed on (driver 580) it is refused with cudaErrorDevicesUnavailable for as long as the process lives
    // (retried for 30 s), so the only recovery there is a new process.  A few retries cover drivers that merely need a
    // moment to tear the faulted channel down.
    for (int attempt = 0;; ++attempt) {
        e = cudaSetDevice(device);
        if (e == cudaSuccess) e = cudaFree(nullptr);   // re-creates the primary context
        if (e == cudaSuccess) return REVE_OK;
        (void)cudaGetLastError();
        if (attempt >= 15) break;                      // ~3 s
        std::this_thread::sleep_for(std::chrono::milliseconds(200));
    }
    return set_err(nullptr, REVE_E_CUDA, cuda_msg(nullptr, "re-initialising the device (restart the process)", e));
}

int reve_ctx_info(const reve_ctx* ctx, int* in_w, int* in_h, int* out_w, int* out_h, int* scale) {
    if (!ctx) return set_err(nullptr, REVE_E_INVAL, "ctx is NULL");
    if (in_w) *in_w = ctx->g.in_w;
    if (in_h) *in_h = ctx->g.in_h;
    if (out_w) *out_w = ctx->g.in_w * ctx->scale;
    if (out_h) *out_h = ctx->g.in_h * ctx->scale;
    if (scale) *scale = ctx->scale;
    return REVE_OK;
}

int reve_host_alloc(size_t bytes, void** out) {
    if (!out || bytes == 0) return set_err(nullptr, REVE_E_INVAL, "bad argument");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
    if (e == cudaErrorMemoryAllocation) return set_err(nullptr, REVE_E_NOMEM, "pinned allocation failed");
    if (e != cudaSuccess) return set_err(nullptr, REVE_E_CUDA, cuda_msg(nullptr, "cudaHostAlloc", e));
    return REVE_OK;
}

void reve_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int reve_submit(reve_ctx* ctx, const uint8_t* rgb_in, size_t in_stride, uint8_t* rgb_out, size_t out_stride,
                uint64_t tag) {
    if (!ctx) return set_err(nullptr, REVE_E_INVAL, "ctx is NULL");
    if (!rgb_in || !rgb_out) return set_err(ctx, REVE_E_INVAL, "frame pointer is NULL");
    const size_t in_row = static_cast<size_t>(ctx->g.in_w) * 3;
    const bool yuv = ctx->out_format != REVE_FMT_RGB24;
    const size_t out_row = yuv ? out_layout(ctx).y_row : in_row * ctx->scale;
    if (in_stride < in_row || out_stride < out_row) return set_err(ctx, REVE_E_INVAL, "row stride smaller than the row");
    if (yuv && (out_stride % 4) != 0) return set_err(ctx, REVE_E_INVAL, "yuv420p10le: the luma row stride must be a multiple of 4 bytes");
    if (ctx->inflight == static_cast<int>(ctx->ring.size())) return set_err(ctx, REVE_E_BUSY, "submit ring full: call reve_wait");
    CK(ctx, cudaSetDevice(ctx->device));
    Slot& s = ctx->ring[ctx->head];
    CK(ctx, cudaMemcpy2DAsync(s.d_in, in_row, rgb_in, in_stride, in_row, ctx->g.in_h, cudaMemcpyHostToDevice, ctx->s_h2d));
    CK(ctx, cudaEventRecord(s.ev_h2d, ctx->s_h2d));
    s.tag = tag;
    s.host_out = rgb_out;
    s.host_out_stride = out_stride;
    ctx->pending.push_back(ctx->head);
    ctx->head = (ctx->head + 1) % static_cast<int>(ctx->ring.size());
    ctx->inflight++;
    // kernels run on whole batches: enqueue once `batch` frames are pending (reve_wait / reve_sync flush the rest)
    if (static_cast<int>(ctx->pending.size()) >= ctx->batch) return flush_pending(ctx);
    return REVE_OK;
}

int reve_wait(reve_ctx* ctx, uint64_t* tag) {
    if (!ctx) return set_err(nullptr, REVE_E_INVAL, "ctx is NULL");
    if (ctx->inflight == 0) return set_err(ctx, REVE_E_EMPTY, "nothing in flight");
    CK(ctx, cudaSetDevice(ctx->device));
    for (int slot : ctx->pending)
        if (slot == ctx->oldest) {   // the oldest frame sits in a partial batch: run it now
            int rc = flush_pending(ctx);
            if (rc != REVE_OK) return rc;
            break;
        }
    Slot& s = ctx->ring[ctx->oldest];
    ctx->oldest = (ctx->oldest + 1) % static_cast<int>(ctx->ring.size());
    ctx->inflight--;
    if (tag) *tag = s.tag;
    CK(ctx, cudaEventSynchronize(s.ev_done));
    return REVE_OK;
}

int reve_sync(reve_ctx* ctx) {
    if (!ctx) return set_err(nullptr, REVE_E_INVAL, "ctx is NULL");
    CK(ctx, cudaSetDevice(ctx->device));
    {
        int rc = flush_pending(ctx);
        if (rc != REVE_OK) return rc;
    }
    CK(ctx, cudaStreamSynchronize(ctx->s_h2d));
    CK(ctx, cudaStreamSynchronize(ctx->s_comp));
    CK(ctx, cudaStreamSynchronize(ctx->s_d2h));
    return REVE_OK;
}

int reve_upscale_device(reve_ctx* ctx, const void* d_in, void* d_out, int n_frames) {
    if (!ctx) return set_err(nullptr, REVE_E_INVAL, "ctx is NULL");
    if (!d_in || !d_out || n_frames < 0) return set_err(ctx, REVE_E_INVAL, "bad argument");
    CK(ctx, cudaSetDevice(ctx->device));
    const size_t in_row = static_cast<size_t>(ctx->g.in_w) * 3, out_row = in_row * ctx->scale;
    const size_t in_bytes = in_row * ctx->g.in_h;
    const size_t out_bytes = ctx->out_format == REVE_FMT_RGB24 ? out_layout(ctx).rgb_bytes : out_layout(ctx).yuv_bytes;
    for (int f0 = 0; f0 < n_frames; f0 += ctx->batch) {
        const int n = (n_frames - f0 < ctx->batch) ? n_frames - f0 : ctx->batch;
        const uint8_t* ins[kMaxBatch];
        uint8_t* outs[kMaxBatch];
        for (int f = 0; f < n; ++f) {
            ins[f] = static_cast<const uint8_t*>(d_in) + static_cast<size_t>(f0 + f) * in_bytes;
            outs[f] = static_cast<uint8_t*>(d_out) + static_cast<size_t>(f0 + f) * out_bytes;
        }
        int rc = enqueue_batch(ctx, n, ins, static_cast<long long>(in_row), outs, static_cast<long long>(out_row));
        if (rc != REVE_OK) return rc;
    }
    return REVE_OK;
}

int reve_ctx_set_output_format(reve_ctx* ctx, int format) {
    if (!ctx) return set_err(nullptr, REVE_E_INVAL, "ctx is NULL");
    if (format != REVE_FMT_RGB24 && format != REVE_FMT_YUV420P10LE_BT601 && format != REVE_FMT_YUV420P10LE_BT709)
        return set_err(ctx, REVE_E_INVAL, "unknown output format");
    if (ctx->inflight) return set_err(ctx, REVE_E_BUSY, "frames in flight");
    CK(ctx, cudaSetDevice(ctx->device));
    if (format != REVE_FMT_RGB24 && !ctx->rgb_scratch[0]) {
        const size_t bytes = out_layout(ctx).rgb_bytes;
        for (int f = 0; f < ctx->batch; ++f) {
            cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ctx->rgb_scratch[f]), bytes);
            if (e == cudaErrorMemoryAllocation) return set_err(ctx, REVE_E_NOMEM, "out of device memory for the RGB scratch frames");
            CK(ctx, e);
        }
    }
    ctx->yuv = colour_coeffs(format == REVE_FMT_YUV420P10LE_BT709 ? 709 : 601);
    ctx->out_format = format;
    return REVE_OK;
}

int reve_ctx_output_layout(const reve_ctx* ctx, size_t* min_stride, size_t* frame_bytes) {
    if (!ctx) return set_err(nullptr, REVE_E_INVAL, "ctx is NULL");
    const OutLayout o = out_layout(ctx);
    const bool yuv = ctx->out_format != REVE_FMT_RGB24;
    if (min_stride) *min_stride = yuv ? o.y_row : o.rgb_row;
    if (frame_bytes) *frame_bytes = yuv ? o.yuv_bytes : o.rgb_bytes;
    return REVE_OK;
}

int reve_ctx_stream(const reve_ctx* ctx, void** stream) {
    if (!ctx || !stream) return set_err(nullptr, REVE_E_INVAL, "NULL argument");
    *stream = ctx->s_comp;
    return REVE_OK;
}

int reve_ctx_set_profiling(reve_ctx* ctx, int on) {
    if (!ctx) return set_err(nullptr, REVE_E_INVAL, "ctx is NULL");
    int rc = collect_profile(ctx);
    ctx->profiling = on != 0;
    return rc;
}

int reve_ctx_get_profile(reve_ctx* ctx, reve_profile* out, int reset) {
    if (!ctx || !out) return set_err(nullptr, REVE_E_INVAL, "NULL argument");
    int rc = collect_profile(ctx);
    if (rc != REVE_OK) return rc;
    *out = ctx->prof;
    if (reset) ctx->prof = reve_profile{};
    return REVE_OK;
}

int reve_debug_features(reve_ctx* ctx, const uint8_t* rgb_in, size_t in_stride, int layer, float* out,
                        size_t cap_floats, int* canvas_w, int* canvas_h) {
    if (!ctx) return set_err(nullptr, REVE_E_INVAL, "ctx is NULL");
    const int cw = ctx->g.canvas_w(), ch = ctx->g.canvas_h();
    if (canvas_w) *canvas_w = cw;
    if (canvas_h) *canvas_h = ch;
    if (!rgb_in || !out || layer < 1 || layer > kNumBody + 1) return set_err(ctx, REVE_E_INVAL, "bad argument");
    const size_t n = static_cast<size_t>(cw) * ch * 64;
    if (cap_floats < n) return set_err(ctx, REVE_E_INVAL, "output buffer too small");
    if (ctx->inflight) return set_err(ctx, REVE_E_BUSY, "frames in flight");
    const size_t in_row = static_cast<size_t>(ctx->g.in_w) * 3;
    if (in_stride < in_row) return set_err(ctx, REVE_E_INVAL, "row stride smaller than the row");
    CK(ctx, cudaSetDevice(ctx->device));
    Slot& s = ctx->ring[0];
    CK(ctx, cudaMemcpy2DAsync(s.d_in, in_row, rgb_in, in_stride, in_row, ctx->g.in_h, cudaMemcpyHostToDevice, ctx->s_comp));
    const uint8_t* ins[1] = {s.d_in};
    uint8_t* outs[1] = {s.d_out};
    int buf = 0;
    int rc = enqueue_batch(ctx, 1, ins, static_cast<long long>(in_row), outs, static_cast<long long>(in_row) * ctx->scale, layer, &buf);
    if (rc != REVE_OK) return rc;
    std::vector<__half> h(n);
    CK(ctx, cudaMemcpyAsync(h.data(), ctx->act[buf], n * sizeof(__half), cudaMemcpyDeviceToHost, ctx->s_comp));
    CK(ctx, cudaStreamSynchronize(ctx->s_comp));
    for (size_t i = 0; i < n; ++i) {
        uint16_t bits;
        std::memcpy(&bits, &h[i], 2);
        out[i] = f16_to_f32(bits);
    }
    return REVE_OK;
}

}  // extern "C" (templates cannot have C linkage)
namespace {
// average device time of `reps` launches of fn on the compute stream
template <typename F>
int timed_launches(reve_ctx* ctx, int reps, float* ms, F fn) {
    cudaEvent_t e0, e1;
    CK(ctx, cudaEventCreate(&e0));
    CK(ctx, cudaEventCreate(&e1));
    CK(ctx, fn());                                    // warm-up (and the result the caller reads)
    CK(ctx, cudaEventRecord(e0, ctx->s_comp));
    for (int r = 0; r < reps; ++r) CK(ctx, fn());
    CK(ctx, cudaEventRecord(e1, ctx->s_comp));
    CK(ctx, cudaStreamSynchronize(ctx->s_comp));
    float t = 0.f;
    CK(ctx, cudaEventElapsedTime(&t, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms) *ms = reps > 0 ? t / reps : 0.f;
    return REVE_OK;
}
}  // namespace
extern "C" {

int reve_debug_unpack(reve_ctx* ctx, const uint8_t* rgb_in, size_t in_stride, float* out, size_t cap_floats, int reps, float* ms) {
    if (!ctx) return set_err(nullptr, REVE_E_INVAL, "ctx is NULL");
    const int cw = ctx->g.canvas_w(), ch = ctx->g.canvas_h();
    const size_t px = static_cast<size_t>(cw) * ch;
    const size_t in_row = static_cast<size_t>(ctx->g.in_w) * 3;
    if (!rgb_in || !out || in_stride < in_row || cap_floats < px * 3 || reps < 0) return set_err(ctx, REVE_E_INVAL, "bad argument");
    if (ctx->inflight) return set_err(ctx, REVE_E_BUSY, "frames in flight");
    CK(ctx, cudaSetDevice(ctx->device));
    Slot& s = ctx->ring[0];
    CK(ctx, cudaMemcpy2DAsync(s.d_in, in_row, rgb_in, in_stride, in_row, ctx->g.in_h, cudaMemcpyHostToDevice, ctx->s_comp));
    void* d = nullptr;
    CK(ctx, cudaMalloc(&d, px * 8));
    // the y table of a single frame: the first canvas_h entries of the stacked table
    int rc = timed_launches(ctx, reps, ms, [&] {
        return launch_unpack_rgb8(ctx->s_comp, s.d_in, static_cast<long long>(in_row), ctx->d_srcx, ctx->d_srcy, cw, ch, d);
    });
    std::vector<uint16_t> h(px * 4);
    if (rc == REVE_OK && cudaMemcpy(h.data(), d, px * 8, cudaMemcpyDeviceToHost) != cudaSuccess) rc = set_err(ctx, REVE_E_CUDA, "copy back failed");
    cudaFree(d);
    if (rc != REVE_OK) return rc;
    for (size_t i = 0; i < px; ++i)
        for (int c = 0; c < 3; ++c) out[i * 3 + c] = f16_to_f32(h[i * 4 + c]);
    return REVE_OK;
}

int reve_debug_pack(reve_ctx* ctx, const float* y, size_t n_floats, uint8_t* rgb_out, size_t out_stride, int reps, float* ms) {
    if (!ctx) return set_err(nullptr, REVE_E_INVAL, "ctx is NULL");
    const int cw = ctx->g.canvas_w(), ch = ctx->g.canvas_h(), sc = ctx->scale;
    const size_t px = static_cast<size_t>(cw) * sc * ch * sc;
    const int out_w = ctx->g.in_w * sc, out_h = ctx->g.in_h * sc;
    if (!y || !rgb_out || n_floats != px * 3 || out_stride < static_cast<size_t>(out_w) * 3 || reps < 0) return set_err(ctx, REVE_E_INVAL, "bad argument");
    if (ctx->inflight) return set_err(ctx, REVE_E_BUSY, "frames in flight");
    CK(ctx, cudaSetDevice(ctx->device));
    std::vector<uint16_t> h(px * 4, 0);
    for (size_t i = 0; i < px; ++i)
        for (int c = 0; c < 3; ++c) h[i * 4 + c] = f32_to_f16(y[i * 3 + c]);
    // output -> canvas tables at input resolution (every output coordinate is kept by exactly one canvas coordinate)
    std::vector<int> inv_x(ctx->g.in_w, -1), inv_y(ctx->g.in_h, -1);
    for (int i = 0; i < cw; ++i) if (ctx->g.x.out[i] >= 0) inv_x[ctx->g.x.out[i]] = i;
    for (int i = 0; i < ch; ++i) if (ctx->g.y.out[i] >= 0) inv_y[ctx->g.y.out[i]] = i;
    void* d = nullptr;
    int *dx = nullptr, *dy = nullptr;
    int rc = REVE_OK;
    if (cudaMalloc(&d, px * 8) != cudaSuccess || cudaMalloc(reinterpret_cast<void**>(&dx), inv_x.size() * 4) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&dy), inv_y.size() * 4) != cudaSuccess)
        rc = set_err(ctx, REVE_E_NOMEM, "out of device memory");
    if (rc == REVE_OK && (cudaMemcpy(d, h.data(), px * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
                          cudaMemcpy(dx, inv_x.data(), inv_x.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
                          cudaMemcpy(dy, inv_y.data(), inv_y.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess))
        rc = set_err(ctx, REVE_E_CUDA, "upload failed");
    Slot& s = ctx->ring[0];
    const size_t out_row = static_cast<size_t>(out_w) * 3;
    if (rc == REVE_OK)
        rc = timed_launches(ctx, reps, ms, [&] {
            return launch_pack_rgb8(ctx->s_comp, d, cw, sc, dx, dy, out_w, out_h, s.d_out, static_cast<long long>(out_row));
        });
    if (rc == REVE_OK && cudaMemcpy2D(rgb_out, out_stride, s.d_out, out_row, out_row, out_h, cudaMemcpyDeviceToHost) != cudaSuccess)
        rc = set_err(ctx, REVE_E_CUDA, "copy back failed");
    cudaFree(d);
    cudaFree(dx);
    cudaFree(dy);
    return rc;
}

int reve_debug_trace(reve_ctx* ctx, long long* out, size_t n) {
    if (!ctx || !out) return set_err(nullptr, REVE_E_INVAL, "NULL argument");
    if (!ctx->d_trace || n > 4096) return set_err(ctx, REVE_E_INVAL, "tracing is off (reve_ctx_options.trace) or n > 4096");
    CK(ctx, cudaStreamSynchronize(ctx->s_comp));
    CK(ctx, cudaMemcpy(out, ctx->d_trace, n * sizeof(long long), cudaMemcpyDeviceToHost));
    return REVE_OK;
}

int reve_geometry(int in_w, int in_h, int scale, int tile, int prepad, int* canvas_w, int* canvas_h, int* src_x,
                  int* out_x, int* src_y, int* out_y, size_t cap) {
    Geometry g;
    std::string err;
    int rc;
    try {
        rc = make_geometry(in_w, in_h, scale, tile, prepad, g, err);
    } catch (const std::exception& e) {
        rc = REVE_E_NOMEM;
        err = e.what();
    }
    if (rc != REVE_OK) return set_err(nullptr, rc, err);
    if (canvas_w) *canvas_w = g.canvas_w();
    if (canvas_h) *canvas_h = g.canvas_h();
    if ((src_x || out_x) && cap < static_cast<size_t>(g.canvas_w())) return set_err(nullptr, REVE_E_INVAL, "cap too small");
    if ((src_y || out_y) && cap < static_cast<size_t>(g.canvas_h())) return set_err(nullptr, REVE_E_INVAL, "cap too small");
    for (int i = 0; i < g.canvas_w(); ++i) {
        if (src_x) src_x[i] = g.x.src[i];
        if (out_x) out_x[i] = g.x.out[i];
    }
    for (int i = 0; i < g.canvas_h(); ++i) {
        if (src_y) src_y[i] = g.y.src[i];
        if (out_y) out_y[i] = g.y.out[i];
    }
    return REVE_OK;
}

int reve_launch_plan(int in_w, int in_h, int scale, int tile, int prepad, int* layers_per_launch, int* strip_px,
                     int* n_strips, int* launches_per_batch) {
    Geometry g;
    std::string err;
    int rc;
    try {
        rc = make_geometry(in_w, in_h, scale, tile, prepad, g, err);
    } catch (const std::exception& e) {
        rc = REVE_E_NOMEM;
        err = e.what();
    }
    if (rc != REVE_OK) return set_err(nullptr, rc, err);
    const int L = choose_chain_len(g.canvas_w());
    const int P = L ? kBoxPx - 2 * L : kStripPx;
    if (layers_per_launch) *layers_per_launch = L ? L : 1;
    if (strip_px) *strip_px = P;
    if (n_strips) *n_strips = (g.canvas_w() + P - 1) / P;
    if (launches_per_batch) *launches_per_batch = 2 + kNumBody / (L ? L : 1);
    return REVE_OK;
}

}  // extern "C"
