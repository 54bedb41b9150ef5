// Kernel-side parameter blocks and launchers (internal to libreve_cuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace reve {

constexpr int kBoxPx = 128;    // pixels per TMA row box = UMMA M
constexpr int kStripPx = 126;  // valid output pixels per strip row (box minus the 2 halo columns)
constexpr int kMaxBatch = 4;    // frames stacked on one canvas per launch (gap row between frames)
constexpr int kConvThreads = 320;  // producer warp + MMA warp + 2 x 4 epilogue warps

// Parameters of the tcgen05 3x3 convolution kernels (body: 64->64 + PReLU -> fp16 canvas;
// tail: 64->3*s*s + PixelShuffle + nearest residual + u8 pack).  Passed as a __grid_constant__
// so bias/slope are read straight from the constant bank.
struct ConvParams {
    int canvas_w, canvas_h;
    int n_strips;
    int total_rows;             // n_strips * n_rows  (strip-rows of work)
    int n_rows;                 // canvas rows this layer computes (<= canvas_h)
    const int* rowmap;          // [n_rows] those rows, ascending; null = all canvas_h rows
    const int* run_fwd;         // [n_rows] length of the run of consecutive rows starting at entry i
    const int* run_bwd;         // [n_rows] length of the run of consecutive rows ending at entry i
    const uint8_t* colflag;     // [canvas_w] 1 = pixel of a tile, 0 = gap column
    const uint8_t* rowflag;     // [canvas_h]
    const void* weights;        // pre-swizzled B operand blob of this layer (global memory)
    __half* out;                // body: output canvas [canvas_h][canvas_w][64] fp16
    int reverse;                // sweep the strip-rows bottom-up (weights blob packed accordingly)
    uint32_t flags;             // debug: bit0 = CTA pair loads the other half of B
    DebugBlock* dbg;            // mapped pinned host memory, may be null
    long long* trace;           // debug timeline of CTA 0 (device memory, may be null)
    // tail only
    const uint8_t* src[kMaxBatch];  // u8 RGB input frames of the batch (device)
    uint8_t* dst[kMaxBatch];        // u8 RGB output frames (device)
    const int* row_frame;       // [canvas_h] frame of the batch a canvas row belongs to (-1 = gap)
    const uint32_t* rowpack;    // [canvas_h] the three row tables packed: (out_y + 1) << 17 | (src_y + 1) << 2 | frame
    long long src_stride, dst_stride;
    const int* src_x;           // [canvas_w] source column (or -1)
    const int* src_y;           // [canvas_h]
    const int* out_x;           // [canvas_w] output column at input resolution (or -1)
    const int* out_y;           // [canvas_h]
    float bias[64];
    float slope[64];
    __half2 slope2[32];         // body: PReLU slopes as packed fp16 pairs (epilogue half2 math)
};

// Chained body layers: `len` consecutive 64->64 layers in ONE launch.  CTA b belongs to chain b / len and computes
// layer b % len of it; a layer hands every finished row to the next one through a small scratch ring in global
// memory that never leaves L2 (TMA store -> flag -> TMA load), so only the first layer reads the canvas and only the
// last one writes it.  The strips of a chain are 128 - 2*len output pixels wide (each layer loses one halo column per
// side), the rows of layer j extend len-1-j rows beyond the last layer's segments.
constexpr int kChainMax = 4;
constexpr int kChainSlots = 8;     // scratch ring slots per (chain, link, stream)
constexpr int kChainFlagStride = 32;   // uint32 words between two flag counters (one 128-byte line each)
struct ChainParams {
    int canvas_w, canvas_h;
    int n_strips;               // strips of 128 - 2*len output pixels
    int total_rows;             // n_strips * n_rows
    int n_rows;                 // canvas rows the LAST layer computes
    const int* rowmap;          // as in ConvParams, for the last layer
    const int* run_fwd;
    const int* run_bwd;
    const uint8_t* colflag;
    const uint8_t* rowflag;
    int len;                    // layers per chain (2 or 4)
    int reverse;
    unsigned int* flags;        // [chain][link][published | consumed][stream]: one 128-byte line per (link, kind), the two
                                // streams' counters side by side (one 64-bit load reads both); zero at launch
    DebugBlock* dbg;
    long long* trace;           // debug timeline of chain `trace_chain` (device memory, may be null)
    int trace_chain;
    uint32_t fault;             // test hook: non-zero makes chain 0 wait for a row nobody announces (watchdog path)
    // Self-balancing split of the strip-rows over the chains: chains run at different speeds (their SMs sit at different
    // distances from the L2 slices that hold the hand-over rings: 841..930 ns per step measured), and a launch ends with
    // its slowest chain.  speed_in[c] = steps per microsecond chain c sustained in the previous chained launch (written by
    // that launch into ITS speed_out; two tables, swapped by the host from launch to launch), and the block of chain c is
    // sized in proportion.  Null = equal blocks.  The split moves work between chains, never the arithmetic of a pixel.
    const float* speed_in;
    float* speed_out;
    const void* weights[kChainMax];
    float bias[kChainMax][64];
    __half2 slope2[kChainMax][32];
};
inline size_t chain_flag_words(int n_chains, int len) { return static_cast<size_t>(n_chains) * (len - 1) * 2 * kChainFlagStride; }
inline size_t chain_scratch_rows(int n_chains, int len) { return static_cast<size_t>(n_chains) * (len - 1) * 2 * kChainSlots * kBoxPx; }
// grid = n_chains * len CTAs, all of which must be resident at the same time (one per SM): launched cooperatively, so
// the driver guarantees it or refuses the launch (cudaErrorCooperativeLaunchTooLarge)
int chain_max_resident_ctas(int sm_count);   // 0 when the device cannot launch cooperatively
// scratch_map: the hand-over rings as [rows][64 ch] with a 128-pixel box (loads); scratch_map_q: the same memory with a
// 32-pixel box (the senders store a row in quarters, one per epilogue warp)
// out_map_q / out_map_e: the output canvas with boxes of 32 and 32 - len pixels (the last layer stores a row in quarters too)
cudaError_t launch_conv_chain(cudaStream_t st, int grid, const CUtensorMap& in_map, const CUtensorMap& out_map,
                              const CUtensorMap& scratch_map, const CUtensorMap& scratch_map_q, const CUtensorMap& out_map_q,
                              const CUtensorMap& out_map_e, const ChainParams& p);

// First convolution (3 -> 64) + PReLU on tensor cores (K = 27 padded to 32), fused with the u8 -> fp16
// unpack, the reflect-101 pre-pad gather and the canvas layout (conv0.cu).
struct Conv0Params {
    int canvas_w, canvas_h;
    const uint8_t* src[kMaxBatch];
    long long src_stride;
    const int* src_x;
    const int* src_y;
    const int* row_frame;      // [canvas_h] frame of the batch (-1 = gap row)
    const void* weights;       // B operand blob (pack_conv0_weights)
    const void* weights_rows;  // the three per-tap B operands of the row-streaming kernel (pack_conv0_rows_weights)
    DebugBlock* dbg;
    float bias[64];
    float slope[64];
    __half2 slope2[32];        // PReLU slopes as packed fp16 pairs
};

size_t conv_weight_blob_bytes(int ng);
// Packs OIHW fp32 weights [co][64][3][3] into the pre-swizzled fp16 B-operand blob for group
// width `ng` (co <= ng; missing output channels are zero); `reverse` = blob for a bottom-up sweep.
void pack_conv_weights(const float* w_oihw, int co, int ng, bool reverse, uint16_t* blob);

cudaError_t conv_kernels_init();  // opt-in shared memory attributes; call once per device
// `pair`: launch as clusters of two CTAs driving tcgen05.mma.cta_group::2 (grid rounded down to even)
// out_map_q / out_map_e: the output canvas with boxes of 32 and 31 pixels (one TMA store per epilogue warp)
cudaError_t launch_conv_body(cudaStream_t st, int grid, bool pair, const CUtensorMap& in_map, const CUtensorMap& out_map_q,
                             const CUtensorMap& out_map_e, const ConvParams& p);
cudaError_t launch_conv_tail(cudaStream_t st, int grid, int scale, const CUtensorMap& in_map, const ConvParams& p);
// 8-bit RGB -> yuv420p10le (yuv.cu).  Integer coefficients scaled by 2^16; matrix = 601 or 709.
struct YuvCoeffs {
    int y[3], u[3], v[3];
};
YuvCoeffs colour_coeffs(int matrix);
// strides in bytes; chroma planes are ceil(w/2) x ceil(h/2)
cudaError_t launch_rgb_to_yuv420p10(cudaStream_t st, const uint8_t* rgb, long long rgb_stride, int w, int h,
                                    uint16_t* y, long long y_stride, uint16_t* u, uint16_t* v, long long c_stride,
                                    const YuvCoeffs& c);
// Stand-alone u8 RGB <-> fp16 conversions (pack.cu; the product path fuses both, these exist for tests and measurement).
// unpack: u8 RGB frame -> fp16 RGB0 canvas [ch][cw] (x / 255, reflect-101 pre-pad and gaps through the geometry tables)
cudaError_t launch_unpack_rgb8(cudaStream_t st, const uint8_t* src, long long src_stride, const int* src_x, const int* src_y,
                               int cw, int ch, void* dst_rgb0_f16);
// pack: fp16 RGB0 network output at canvas geometry [ch*s][cw*s] -> cropped u8 RGB frame, u8 = clamp(floor(v*255 + 0.5));
// inv_x / inv_y: output column / row at input resolution -> canvas column / row
cudaError_t launch_pack_rgb8(cudaStream_t st, const void* src_rgb0_f16, int cw, int scale, const int* inv_x, const int* inv_y,
                             int out_w, int out_h, uint8_t* dst, long long dst_stride);
size_t conv0_weight_blob_bytes();
void pack_conv0_weights(const float* w_oihw, uint16_t* blob);
cudaError_t conv0_kernel_init();
cudaError_t launch_conv0(cudaStream_t st, int grid, const CUtensorMap& out_map, const Conv0Params& p);
// Row-streaming variant (conv0_rows.cu): every input row gathered once, three per-tap MMAs; out_map_q = the output canvas
// with 32-pixel boxes.  grid <= strips * rows.
size_t conv0_rows_weight_blob_bytes();
void pack_conv0_rows_weights(const float* w_oihw, uint16_t* blob);
cudaError_t conv0_rows_kernel_init();
cudaError_t launch_conv0_rows(cudaStream_t st, int grid, const CUtensorMap& out_map_q, const Conv0Params& p);

}  // namespace reve
