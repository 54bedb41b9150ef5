/* reve_cuda.h -- C ABI of libreve_cuda: the B200-native (sm_100a) replacement for what REVE
 * reaches through `Command::new("realesrgan-ncnn-vulkan")`.
 *
 * What it replaces in the reference (ONdraid/reve; paths relative to the reference root):
 *   - reve-shared/src/lib.rs:129-155  Video::upscale_segment: spawns the upscaler once per
 *     segment with `-i temp\tmp_frames\{i} -o temp\out_frames\{i} -n realesr-animevideov3-x2
 *     -s {scale} -f png -v` and returns its stderr.
 *   - reve-cli/src/main.rs:262-273    the caller blocks on that reader and counts "done" lines.
 *   - reve-gui/src-tauri/src/commands.rs:52-64  the GUI's spawn of the same binary.
 * The reference has a process boundary there, not an FFI; this header is the FFI a Rust crate
 * (`reve-upscale`) binds instead.  Data at the boundary: packed 8-bit RGB,
 * HWC, row stride in bytes, origin top-left (what ffmpeg's PNG / rgb24 export holds); output is
 * (W*s) x (H*s) in the same layout.  No alpha, no 16-bit.
 *
 * Conventions: every function returns 0 (REVE_OK) or a negative reve_status; the message of the
 * last failure on a context is available through reve_last_error(ctx) (or reve_last_error(NULL)
 * for failures of functions that have no context).  No C++ exception crosses this boundary.
 * There is no CPU fallback: on a machine without an sm_100 device every compute entry point
 * fails with REVE_E_ARCH / REVE_E_CUDA.
 *
 * Threading: a reve_ctx is bound to one device and is NOT thread-safe (Rust: Send, !Sync);
 * distinct contexts may be driven concurrently from distinct host threads (one per GPU).
 * A reve_model is immutable after creation and may be shared by any number of contexts.
 * The library spawns no host threads.
 */
#ifndef REVE_CUDA_H
#define REVE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REVE_VERSION 200 /* 0.2.0 */

typedef enum reve_status {
    REVE_OK = 0,
    REVE_E_INVAL = -1, /* bad argument */
    REVE_E_NOMEM = -2, /* host or device allocation failed */
    REVE_E_CUDA = -3,  /* CUDA runtime/driver error (incl. no device) */
    REVE_E_IO = -4,    /* file could not be read / written */
    REVE_E_MODEL = -5, /* .param/.bin is not a realesr-animevideov3 (SRVGGNetCompact) model */
    REVE_E_ARCH = -6,  /* device is not compute capability 10.x (sm_100) */
    REVE_E_BUSY = -7,  /* submit ring full: call reve_wait first */
    REVE_E_EMPTY = -8  /* reve_wait with nothing in flight */
} reve_status;

typedef struct reve_model reve_model;
typedef struct reve_ctx reve_ctx;

/* ---- library ------------------------------------------------------------------------------ */
int reve_version(void);
const char* reve_strerror(int status);
/* Message of the last failure on `ctx` (ctx == NULL: last failure of a context-free call on this
 * thread).  Never NULL; valid until the next failing call on the same context / thread. */
const char* reve_last_error(const reve_ctx* ctx);
/* Number of CUDA devices that can run the kernels (compute capability 10.0). */
int reve_device_count(int* n);
/* PCI address of a device as sysfs spells it ("0000:1b:00.0"), so that a host can place the worker thread and the pinned
 * buffers of a GPU on the NUMA node it hangs off (/sys/bus/pci/devices/<id>/numa_node); cap >= 16. */
int reve_device_pci_bus_id(int device, char* buf, size_t cap);

/* ---- model (replaces `-n realesr-animevideov3-x{s}` + the models\ directory) ---------------- */
/* Parse an ncnn .param/.bin pair and validate that it is SRVGGNetCompact(3->64, 16 body convs,
 * 64->3*s*s, PReLU, PixelShuffle(s) + nearest residual).  fp16-tagged and raw fp32 payloads. */
int reve_model_load_ncnn(const char* param_path, const char* bin_path, reve_model** out);
/* Seeded He-normal random init of the named architecture (used when the weight files are
 * absent offline).  Bit-identical to oracle/srvgg.py:make_weights(scale, seed). */
int reve_model_random(int scale, uint64_t seed, reve_model** out);
/* Build the model from host arrays, e.g. the tensors of a Real-ESRGAN `.pth` checkpoint
 * (SRVGGNetCompact: `body.{0,2,..,34}.weight/.bias` are the 18 convolutions, `body.{1,3,..,33}.weight`
 * the 17 PReLU slopes).  conv_w[k]: fp32 OIHW [out_k][in_k][3][3] with
 * in_0 = 3, out_17 = 3*scale*scale, 64 otherwise; conv_b[k]: [out_k]; prelu[k]: [64] for k < 17
 * (prelu[17] is ignored).  The arrays are copied.  Weights are stored as given; the device packs them to
 * fp16 exactly like an fp32 .bin (REVE_E_MODEL if a weight is NaN or outside the fp16 range). */
int reve_model_from_arrays(int scale, const float* const* conv_w, const float* const* conv_b,
                           const float* const* prelu, reve_model** out);
/* Write the model as ncnn .param/.bin (fp16 != 0: tag 0x01306B47 payload). */
int reve_model_save_ncnn(const reve_model* m, const char* param_path, const char* bin_path, int fp16);
int reve_model_info(const reve_model* m, int* scale, int* num_feat, int* num_conv);
void reve_model_free(reve_model* m);

/* ---- context (one per GPU; owns streams, device buffers, tensor maps) ----------------------- */
/* in_w x in_h: input frame size.  tile: 0 = whole frame, T > 0 = upstream tile size (the spawned
 * binary uses 200 on any device with > 1.9 GB).  prepad: upstream uses 10.  ring_depth: frames
 * that may be in flight between reve_submit and reve_wait (1..16). */
int reve_ctx_create(int device, const reve_model* m, int in_w, int in_h, int tile, int prepad,
                    int ring_depth, reve_ctx** out);

/* Same with explicit options (everything the library decides by itself in reve_ctx_create can be pinned here; the
 * library reads NO environment variables).  Zero-initialise, set struct_size = sizeof(reve_ctx_options), then the
 * fields of interest.
 *
 * REVE_CTX_SHARED_DEVICE: other GPU work of this or another process may run on the device while this context is
 * busy and the caller prefers kernels that never wait for one another: the 16 body layers then run as 16 separate
 * launches.  Without the flag the body runs as chains of 4 (or 2) layers per launch whose CTAs hand rows to each other
 * through L2 (about 10 % faster at 1080p).  That kernel is launched COOPERATIVELY, i.e. the driver starts it only once
 * all of its CTAs can be resident together, so it is safe -- merely serialised -- next to a second context, the
 * colour-conversion kernel or an MPS co-tenant; if the device cannot hold the grid at all (MIG slice, green context)
 * the context silently uses the single-layer launches.  reve_ctx_launch_info reports which structure is in use. */
#define REVE_CTX_SHARED_DEVICE 1u
/* Test hooks (results stay correct unless stated): */
#define REVE_DBG_NO_REVERSE 1u      /* every layer sweeps top-down (changes the last bit of some accumulations) */
#define REVE_DBG_CTA_PAIRS 2u       /* body layers as CTA pairs driving tcgen05.mma.cta_group::2 (no chains) */
#define REVE_DBG_SWAP_PAIR_B 4u     /* pairs load the wrong half of B: results are WRONG (negative test) */
#define REVE_DBG_ALL_ROWS 8u        /* no needed-row lists: every layer computes every canvas row */
#define REVE_DBG_ALIAS_ROWS 16u     /* canvas rows alias each other in L2: timing experiment, results are WRONG */
#define REVE_DBG_FAULT 32u          /* chained kernel waits for a row that never comes: exercises the watchdog path */
#define REVE_DBG_EQUAL_SPLIT 64u    /* chained kernel: equal blocks per chain instead of the self-balancing split */
#define REVE_DBG_CONV0_IM2COL 128u  /* first conv: the K = 27 im2col kernel (conv0.cu) instead of the row-streaming one (conv0_rows.cu) */
typedef struct reve_ctx_options {
    uint32_t struct_size;      /* sizeof(reve_ctx_options) */
    uint32_t flags;            /* REVE_CTX_* */
    int layers_per_launch;     /* 0 = automatic (reve_launch_plan), 1 = one launch per layer, 2 / 4 = chains */
    int max_batch;             /* 0 = automatic; frames stacked on the canvas per launch set (1..4) */
    /* test hooks */
    uint32_t debug_flags;      /* REVE_DBG_* */
    int debug_grid;            /* 0 = one CTA per SM; n > 0 caps the persistent grids at n CTAs */
    int trace;                 /* 0 = off; 1 = body layer 5 / the chained launch `trace_launch`; 2 = tail (reve_debug_trace) */
    int trace_launch;          /* chained launch to trace (0 .. 16/L - 1) */
    int trace_chain;           /* chain of that launch */
} reve_ctx_options;
int reve_ctx_create_ex(int device, const reve_model* m, int in_w, int in_h, int tile, int prepad,
                       int ring_depth, const reve_ctx_options* opt, reve_ctx** out);
void reve_ctx_destroy(reve_ctx* ctx);
/* Launch structure in use: body layers per launch (1, 2 or 4), frames per launch set, CTAs of the body launch,
 * whether the body launches are cooperative.  Any pointer may be NULL. */
int reve_ctx_launch_info(const reve_ctx* ctx, int* layers_per_launch, int* batch, int* grid, int* cooperative);

/* Failure semantics.  A kernel-side protocol fault (a wait that exceeds its watchdog, ~2^32 SM cycles) ends the launch
 * with a trap after writing a diagnostic word to host-mapped memory; the call that notices it (reve_wait, reve_sync,
 * reve_submit, ...) returns REVE_E_CUDA and reve_last_error(ctx) carries "kernel watchdog: wait tag T timed out in
 * block B".  As after any asynchronous CUDA fault (Xid, ECC, illegal address) the error is STICKY: the CUDA primary
 * context of that device is lost for the whole process, so every reve_ctx on the same device fails with REVE_E_CUDA
 * from then on (contexts on other devices and other processes are unaffected) and the frames in flight are lost;
 * reve_ctx_destroy stays safe on a dead context.  What a caller should do: destroy the contexts of that device, record
 * the last frame whose reve_wait succeeded, and restart the worker PROCESS for that device (the segment scheduler's
 * resume file makes that cheap: an unfinished segment is simply still listed).  reve_device_recover(device) attempts
 * the in-process alternative -- cudaDeviceReset and a fresh primary context, after which contexts can be created
 * again (pinned buffers obtained while that device was current must be dropped, not freed; reve_model handles live
 * in host memory and stay valid) -- but whether the driver allows it is not ours to decide: on the B200 / driver 580
 * pool this library was developed on it is refused (cudaErrorDevicesUnavailable, returned as REVE_E_CUDA). */
int reve_device_recover(int device);
/* Geometry of the context: output frame size and scale. */
int reve_ctx_info(const reve_ctx* ctx, int* in_w, int* in_h, int* out_w, int* out_h, int* scale);

/* Output pixel format.  The reference pipes the upscaled PNGs into
 * `ffmpeg ... -pix_fmt yuv420p10le -c:v libx265` (reve-cli/src/main.rs:306-326), so swscale converts every
 * frame on the host; with a YUV format the frames leave the GPU in the encoder's native layout instead
 * (same 3 bytes per pixel over PCIe), ready for `ffmpeg -f rawvideo -pix_fmt yuv420p10le -s WxH -i pipe:0`.
 * Planar, 16-bit little-endian samples holding 10 bits, limited range (Y 64..940, C 64..960), 2x2 box-filtered
 * chroma, integer arithmetic defined in reve_b200/csrc/yuv.cu (restated in oracle/colour.py).  BT601 is what
 * swscale applies to untagged RGB input (the reference's behaviour); BT709 is the HD matrix.
 * Buffer layout for reve_submit / reve_upscale_device: Y plane (out_h rows of out_stride bytes), then the U
 * plane and the V plane (ceil(out_h/2) rows of out_stride/2 bytes each).  Not while frames are in flight. */
typedef enum reve_format {
    REVE_FMT_RGB24 = 0,
    REVE_FMT_YUV420P10LE_BT601 = 1,
    REVE_FMT_YUV420P10LE_BT709 = 2
} reve_format;
int reve_ctx_set_output_format(reve_ctx* ctx, int format);
/* Smallest legal out_stride for the current format and the bytes of one frame stored with it. */
int reve_ctx_output_layout(const reve_ctx* ctx, size_t* min_stride, size_t* frame_bytes);

/* Pinned host memory for frame buffers (plain malloc'ed buffers work too, but copy slower and
 * do not overlap). */
int reve_host_alloc(size_t bytes, void** out);
void reve_host_free(void* p);

/* Asynchronous frame: H2D copy -> kernels -> D2H copy on the context's streams.  The caller keeps
 * rgb_in / rgb_out valid until the matching reve_wait.  Strides in bytes (>= 3*w).  Kernels run on
 * batches of up to 4 frames (stacked on one canvas): the copy starts at once, the kernels are
 * enqueued when a batch is full or when reve_wait / reve_sync needs a frame of a partial batch. */
int reve_submit(reve_ctx* ctx, const uint8_t* rgb_in, size_t in_stride, uint8_t* rgb_out,
                size_t out_stride, uint64_t tag);
/* Blocks until the oldest submitted frame is complete (FIFO) and returns its tag. */
int reve_wait(reve_ctx* ctx, uint64_t* tag);
/* Blocks until everything submitted on the context has finished. */
int reve_sync(reve_ctx* ctx);

/* Device-resident variant (kernel-only benchmarks, or callers that already hold frames on the
 * GPU): d_in = n_frames packed frames (3*in_w*in_h bytes each), d_out likewise at output size (YUV formats:
 * frames of reve_ctx_output_layout's frame_bytes with the minimal stride).
 * Enqueued on the context's compute stream; returns without synchronising. */
int reve_upscale_device(reve_ctx* ctx, const void* d_in, void* d_out, int n_frames);
/* The context's compute stream (a cudaStream_t), for callers that time with their own events. */
int reve_ctx_stream(const reve_ctx* ctx, void** stream);

/* ---- instrumentation ---------------------------------------------------------------------- */
typedef struct reve_profile {
    uint64_t launches_conv0, launches_body, launches_tail; /* kernels launched since reset */
    double ms_conv0, ms_body, ms_tail; /* summed CUDA-event time; only while profiling is on */
    uint64_t timed_body;               /* body launches covered by ms_body */
    uint64_t timed_frames;             /* tail launches (= batches) covered by the timings */
    uint64_t frames;                   /* frames enqueued since reset */
    uint64_t body_frames;              /* sum over body launches of the frames each one processed */
    uint64_t launches_yuv;             /* colour-conversion kernels (one per frame when the output is YUV) */
    uint64_t body_layer_frames;        /* sum over body launches of layers x frames (a chained launch runs several layers) */
} reve_profile;
/* on != 0: bracket every kernel launch with CUDA events (slower; for roofline measurements). */
int reve_ctx_set_profiling(reve_ctx* ctx, int on);
/* Synchronises the compute stream, then fills *out; reset != 0 clears the counters. */
int reve_ctx_get_profile(reve_ctx* ctx, reve_profile* out, int reset);

/* Test hook: run the network on one frame and copy out the fp16 NHWC feature canvas after
 * `layer` convolutions+PReLU (1..17) as float32 [canvas_h][canvas_w][64] into `out`
 * (cap_floats = capacity).  canvas_w/h may be NULL.  Used by the per-layer parity tests. */
int reve_debug_features(reve_ctx* ctx, const uint8_t* rgb_in, size_t in_stride, int layer,
                        float* out, size_t cap_floats, int* canvas_w, int* canvas_h);
/* Test hook: with reve_ctx_options.trace = 1 at context creation, CTA 0 of body layer 5
 * records clock64() timestamps (MMA warp: out[i] at the start of step i, out[1000] = look-ahead misses;
 * epilogue group 0: out[1024 + 4*e + 0..3] = wait start / accumulator full / slot released / row stored of
 * event e); this copies the first n (<= 2048) words out.  See tools/gpu_trace.py. */
int reve_debug_trace(reve_ctx* ctx, long long* out, size_t n);
/* Test hooks for the stand-alone conversion kernels (reve_b200/csrc/pack.cu; SURVEY.md section 2.3 K1 / K7 -- the
 * product path fuses both conversions into conv0 and the tail).  One frame of the context's geometry.
 * unpack: u8 RGB frame -> x/255 as fp16 on the canvas (reflect-101 pre-pad, zero gaps), returned as float32
 *         [canvas_h][canvas_w][3].
 * pack:   `y` = float32 [canvas_h*s][canvas_w*s][3], a network output at canvas geometry (rounded to fp16 on upload)
 *         -> cropped u8 RGB frame, u8 = clamp(floor(v*255 + 0.5)).
 * ms (may be NULL): device time of the kernel alone, average of `reps` launches (CUDA events). */
int reve_debug_unpack(reve_ctx* ctx, const uint8_t* rgb_in, size_t in_stride, float* out, size_t cap_floats, int reps, float* ms);
int reve_debug_pack(reve_ctx* ctx, const float* y, size_t n_floats, uint8_t* rgb_out, size_t out_stride, int reps, float* ms);

/* Canvas geometry tables (context-free, host only; test hook): the canvas is the side-by-side
 * layout of upstream's padded tiles.  For canvas column/row i: the source frame coordinate feeding
 * it (reflect-101 applied; -1 for a gap) and the output coordinate at input resolution (-1 if
 * cropped).  Arrays may be NULL; cap = capacity of each non-NULL array in ints. */
int reve_geometry(int in_w, int in_h, int scale, int tile, int prepad, int* canvas_w, int* canvas_h,
                  int* src_x, int* out_x, int* src_y, int* out_y, size_t cap);

/* Launch structure the library chooses for a frame size (context-free, host only): the 16 body layers of
 * SRVGGNetCompact run as chains of `layers_per_launch` (4, 2 or 1) layers per kernel launch whose activations are
 * handed from SM to SM through L2-resident rings; a chain of L layers works on strips of 128 - 2L output columns
 * (126 for single layers), so the choice depends on how many strips the canvas width needs.  launches_per_batch counts
 * conv0 + body launches + tail.  reve_ctx_options.layers_per_launch overrides the choice when a context is created.
 * Any pointer may be NULL. */
int reve_launch_plan(int in_w, int in_h, int scale, int tile, int prepad, int* layers_per_launch, int* strip_px,
                     int* n_strips, int* launches_per_batch);

#ifdef __cplusplus
}
#endif
#endif /* REVE_CUDA_H */
