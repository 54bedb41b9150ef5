// First convolution (3 -> 64, 3x3) + PReLU, fused with the pre-processing of the upscaler spawned
// at reference reve-shared/src/lib.rs:134-147 (SURVEY.md section 2.3, K1 + K2): u8 RGB gather
// with the reflect-101 pre-pad, /255, zero padding at tile borders, canvas layout, fp16 NHWC out.
//
// 0.29 % of the network's FLOPs: CUDA cores, one canvas pixel per thread, all 1 728 weights read
// as constant-bank operands of the FFMAs (the parameter block is a __grid_constant__).
#include "kernels.h"

namespace reve {

namespace {

constexpr int kConv0Threads = 128;

__global__ void __launch_bounds__(kConv0Threads)
conv0_kernel(const __grid_constant__ Conv0Params p) {
    const int cx = blockIdx.x * kConv0Threads + threadIdx.x;
    const int cy = blockIdx.y;
    if (cx >= p.canvas_w) return;
    uint4* dst = reinterpret_cast<uint4*>(p.dst + (static_cast<size_t>(cy) * p.canvas_w + cx) * 64);

    const int sxc = p.src_x[cx];
    const int syc = p.src_y[cy];
    if (sxc < 0 || syc < 0) {  // gap pixel: must read as zero in every later layer
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = make_uint4(0, 0, 0, 0);
        return;
    }
    float x[27];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int yy = cy + ky - 1;
        const int sy = (yy >= 0 && yy < p.canvas_h) ? p.src_y[yy] : -1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int xx = cx + kx - 1;
            const int sx = (xx >= 0 && xx < p.canvas_w) ? p.src_x[xx] : -1;
            if (sy >= 0 && sx >= 0) {
                const uint8_t* s = p.src + static_cast<long long>(sy) * p.src_stride + sx * 3;
                x[(ky * 3 + kx) * 3 + 0] = static_cast<float>(s[0]) * (1.0f / 255.0f);
                x[(ky * 3 + kx) * 3 + 1] = static_cast<float>(s[1]) * (1.0f / 255.0f);
                x[(ky * 3 + kx) * 3 + 2] = static_cast<float>(s[2]) * (1.0f / 255.0f);
            } else {
                x[(ky * 3 + kx) * 3 + 0] = 0.f;
                x[(ky * 3 + kx) * 3 + 1] = 0.f;
                x[(ky * 3 + kx) * 3 + 2] = 0.f;
            }
        }
    }
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = p.bias[c8 * 8 + j];
#pragma unroll
        for (int t = 0; t < 27; ++t) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(x[t], p.w[t][c8 * 8 + j], acc[j]);
        }
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v0 = acc[2 * j], v1 = acc[2 * j + 1];
            v0 = fmaxf(v0, 0.f) + p.slope[c8 * 8 + 2 * j] * fminf(v0, 0.f);
            v1 = fmaxf(v1, 0.f) + p.slope[c8 * 8 + 2 * j + 1] * fminf(v1, 0.f);
            const __half2 h = __floats2half2_rn(v0, v1);
            pk[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        dst[c8] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

}  // namespace

cudaError_t launch_conv0(cudaStream_t st, const Conv0Params& p) {
    const dim3 grid((p.canvas_w + kConv0Threads - 1) / kConv0Threads, p.canvas_h);
    conv0_kernel<<<grid, kConv0Threads, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace reve
