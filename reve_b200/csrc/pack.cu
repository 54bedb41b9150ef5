// Stand-alone u8 RGB <-> fp16 unpack / pack kernels (upstream's realesrgan_preproc /
// realesrgan_postproc compute shaders, run by the upscaler spawned at reference reve-shared/src/lib.rs:134-147).
//
// In the product path both stages are FUSED -- the unpack into conv0's im2col producers (conv0.cu), the pack into the
// tail's epilogue (conv_umma.cu) -- so that neither the fp16 input canvas nor the fp32 network output ever exists in
// HBM.  These stand-alone versions compute the same two functions on their own (same geometry tables, same rounding) and
// exist so that the two conversions can be tested in isolation against the oracle (reve_debug_unpack / reve_debug_pack)
// and measured against the HBM roofline they are bound by.  Both are plain streaming kernels:
//   unpack: one thread per canvas pixel, 3 byte loads through the reflect-101 tables, one 8-byte store (RGB0 as fp16);
//           a warp writes 256 contiguous bytes.
//   pack:   one thread per 4 output pixels of one output row: four 8-byte loads, u8 = clamp(floor(v*255 + 0.5)),
//           three 4-byte stores (12 contiguous bytes; a warp writes 384 contiguous bytes), scalar at ragged row ends.
#include "kernels.h"

namespace reve {

namespace {

__global__ void __launch_bounds__(256)
unpack_rgb8_to_f16_kernel(const uint8_t* __restrict__ src, long long src_stride, const int* __restrict__ src_x,
                          const int* __restrict__ src_y, int cw, int ch, uint2* __restrict__ dst) {
    const long long n = static_cast<long long>(cw) * ch;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256ll) {
        const int cy = static_cast<int>(i / cw), cx = static_cast<int>(i - static_cast<long long>(cy) * cw);
        const int sx = src_x[cx], sy = src_y[cy];
        uint2 v = make_uint2(0u, 0u);                       // gap rows / columns between tiles read as zero
        if (sx >= 0 && sy >= 0) {
            const uint8_t* p = src + static_cast<long long>(sy) * src_stride + sx * 3;
            const __half2 rg = __floats2half2_rn(static_cast<float>(p[0]) * (1.0f / 255.0f), static_cast<float>(p[1]) * (1.0f / 255.0f));
            const __half2 b0 = __floats2half2_rn(static_cast<float>(p[2]) * (1.0f / 255.0f), 0.f);
            v.x = *reinterpret_cast<const uint32_t*>(&rg);
            v.y = *reinterpret_cast<const uint32_t*>(&b0);
        }
        dst[i] = v;
    }
}

__device__ __forceinline__ uint32_t quant(float v) {
    // u8 = clamp(floor(v*255 + 0.5), 0, 255); NaN -> 0 (oracle/srvgg.py:quantise)
    return static_cast<uint32_t>(fminf(fmaxf(floorf(fmaf(v, 255.f, 0.5f)), 0.f), 255.f));
}
__device__ __forceinline__ void quant_px(uint2 v, uint32_t (&b)[3]) {
    const float2 rg = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
    const float2 bz = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    b[0] = quant(rg.x);
    b[1] = quant(rg.y);
    b[2] = quant(bz.x);
}

// src: fp16 RGB0 at canvas resolution x scale, [ch*s][cw*s]; out_x / out_y: canvas column / row -> output column / row
// at INPUT resolution (-1 = cropped pre-pad or gap).  inv_x[ox] / inv_y[oy] (output -> canvas, at input resolution) are
// built by the host from the same tables.
__global__ void __launch_bounds__(256)
pack_f16_to_rgb8_kernel(const uint2* __restrict__ src, int cw, int s, const int* __restrict__ inv_x,
                        const int* __restrict__ inv_y, int out_w, int out_h, uint8_t* __restrict__ dst, long long dst_stride) {
    const int groups = (out_w + 3) / 4;                     // 4 output pixels per thread
    const long long n = static_cast<long long>(groups) * out_h;
    const long long src_pitch = static_cast<long long>(cw) * s;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256ll) {
        const int oy = static_cast<int>(i / groups), ox0 = static_cast<int>(i - static_cast<long long>(oy) * groups) * 4;
        const long long srow = (static_cast<long long>(inv_y[oy / s]) * s + oy % s) * src_pitch;
        uint8_t* const drow = dst + static_cast<long long>(oy) * dst_stride;
        uint32_t b[4][3];
        const int npx = min(4, out_w - ox0);
        for (int k = 0; k < npx; ++k) {
            const int ox = ox0 + k;
            quant_px(src[srow + static_cast<long long>(inv_x[ox / s]) * s + ox % s], b[k]);
        }
        uint8_t* const d = drow + ox0 * 3;
        if (npx == 4 && (reinterpret_cast<uintptr_t>(d) & 3) == 0) {
            uint32_t* const d32 = reinterpret_cast<uint32_t*>(d);
            d32[0] = b[0][0] | (b[0][1] << 8) | (b[0][2] << 16) | (b[1][0] << 24);
            d32[1] = b[1][1] | (b[1][2] << 8) | (b[2][0] << 16) | (b[2][1] << 24);
            d32[2] = b[2][2] | (b[3][0] << 8) | (b[3][1] << 16) | (b[3][2] << 24);
        } else {
            for (int k = 0; k < npx; ++k)
                for (int c = 0; c < 3; ++c) d[k * 3 + c] = static_cast<uint8_t>(b[k][c]);
        }
    }
}

int grid_for(long long items) {
    const long long blocks = (items + 255) / 256;
    return static_cast<int>(blocks < 148 * 8 ? (blocks < 1 ? 1 : blocks) : 148 * 8);   // a multiple of the SM count, grid-stride
}

}  // namespace

cudaError_t launch_unpack_rgb8(cudaStream_t st, const uint8_t* src, long long src_stride, const int* src_x, const int* src_y,
                               int cw, int ch, void* dst_rgb0_f16) {
    unpack_rgb8_to_f16_kernel<<<grid_for(static_cast<long long>(cw) * ch), 256, 0, st>>>(src, src_stride, src_x, src_y, cw, ch,
                                                                                          static_cast<uint2*>(dst_rgb0_f16));
    return cudaGetLastError();
}

cudaError_t launch_pack_rgb8(cudaStream_t st, const void* src_rgb0_f16, int cw, int scale, const int* inv_x, const int* inv_y,
                             int out_w, int out_h, uint8_t* dst, long long dst_stride) {
    pack_f16_to_rgb8_kernel<<<grid_for(static_cast<long long>((out_w + 3) / 4) * out_h), 256, 0, st>>>(
        static_cast<const uint2*>(src_rgb0_f16), cw, scale, inv_x, inv_y, out_w, out_h, dst, dst_stride);
    return cudaGetLastError();
}

}  // namespace reve
