// Packed 8-bit RGB -> planar yuv420p10le on the GPU.
//
// The reference hands the upscaled PNG frames to `ffmpeg ... -pix_fmt yuv420p10le -c:v libx265`
// (reve-cli/src/main.rs:306-326), so swscale converts every 4K frame on the host.  This kernel produces the
// encoder's native input format right after the tail kernel; the D2H copy stays 3 bytes per pixel.
//
// Definition (integer, so the oracle in oracle/colour.py matches bit for bit; limited "video" range):
//   Y  = (yr*R + yg*G + yb*B + (64 << 16) + 2^15) >> 16          per pixel, 64..940
//   Cb = (ur*SR + ug*SG + ub*SB + (512 << 18) + 2^17) >> 18       per 2x2 block, S* = sum of the 4 pixels
//   Cr = (vr*SR + vg*SG + vb*SB + (512 << 18) + 2^17) >> 18       (box filter; edge pixels replicated)
// with the coefficients of BT.601 (what swscale applies to untagged RGB input) or BT.709 scaled by 2^16
// (colour_coeffs()).  HBM-bound: 3 B/px read + 3 B/px written; one thread per 4x2 pixels, 32-bit loads and
// 64-/32-bit stores when the row pitches allow.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace reve {

YuvCoeffs colour_coeffs(int matrix) {
    // Kr, Kb of the luma equation; Kg = 1 - Kr - Kb
    const double kr = (matrix == 709) ? 0.2126 : 0.299, kb = (matrix == 709) ? 0.0722 : 0.114, kg = 1.0 - kr - kb;
    const double ys = 876.0 / 255.0 * 65536.0;          // 8-bit full range -> 10-bit limited luma
    const double cs = 896.0 / 255.0 * 65536.0;          // ... chroma excursion (x 1/4 for the sum of 4 pixels via >> 18)
    auto r = [](double v) { return static_cast<int>(v < 0 ? v - 0.5 : v + 0.5); };
    YuvCoeffs c;
    c.y[0] = r(kr * ys); c.y[1] = r(kg * ys); c.y[2] = r(kb * ys);
    c.u[0] = r(-kr / (2 * (1 - kb)) * cs); c.u[1] = r(-kg / (2 * (1 - kb)) * cs); c.u[2] = r(0.5 * cs);
    c.v[0] = r(0.5 * cs); c.v[1] = r(-kg / (2 * (1 - kr)) * cs); c.v[2] = r(-kb / (2 * (1 - kr)) * cs);
    return c;
}

namespace {

__device__ __forceinline__ uint32_t luma(const YuvCoeffs& c, int r, int g, int b) {
    return static_cast<uint32_t>((c.y[0] * r + c.y[1] * g + c.y[2] * b + (64 << 16) + (1 << 15)) >> 16);
}
__device__ __forceinline__ uint32_t chroma(const int (&k)[3], int sr, int sg, int sb) {
    return static_cast<uint32_t>((k[0] * sr + k[1] * sg + k[2] * sb + (512 << 18) + (1 << 17)) >> 18);
}

// One thread: 4 pixels x 2 rows = two chroma samples.  FAST: w % 4 == 0, h % 2 == 0 and all pitches/bases aligned.
template <bool FAST>
__global__ void __launch_bounds__(256)
rgb_to_yuv420p10_kernel(const uint8_t* __restrict__ rgb, long long rgb_stride, int w, int h,
                        uint16_t* __restrict__ yp, long long y_stride, uint16_t* __restrict__ up,
                        uint16_t* __restrict__ vp, long long c_stride, const YuvCoeffs c) {
    const int bx = blockIdx.x * blockDim.x + threadIdx.x;   // block of 4 pixels
    const int by = blockIdx.y * blockDim.y + threadIdx.y;   // pair of rows
    const int x0 = bx * 4, y0 = by * 2;
    if (x0 >= w || y0 >= h) return;
    int px[2][4][3];
    if constexpr (FAST) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(rgb + (y0 + r) * rgb_stride + x0 * 3);
            const uint32_t a = __ldg(p), b = __ldg(p + 1), d = __ldg(p + 2);
            const uint32_t bytes[3] = {a, b, d};
#pragma unroll
            for (int i = 0; i < 12; ++i) px[r][i / 3][i % 3] = (bytes[i >> 2] >> ((i & 3) * 8)) & 0xFF;
        }
    } else {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int yy = min(y0 + r, h - 1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int xx = min(x0 + i, w - 1);
                const uint8_t* p = rgb + yy * rgb_stride + xx * 3;
                px[r][i][0] = p[0]; px[r][i][1] = p[1]; px[r][i][2] = p[2];
            }
        }
    }
    uint32_t yv[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) yv[r][i] = luma(c, px[r][i][0], px[r][i][1], px[r][i][2]);
    uint32_t uu[2], vv[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        int s[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) s[ch] = px[0][2 * j][ch] + px[0][2 * j + 1][ch] + px[1][2 * j][ch] + px[1][2 * j + 1][ch];
        uu[j] = chroma(c.u, s[0], s[1], s[2]);
        vv[j] = chroma(c.v, s[0], s[1], s[2]);
    }
    uint8_t* const yb = reinterpret_cast<uint8_t*>(yp);
    uint8_t* const ub = reinterpret_cast<uint8_t*>(up);
    uint8_t* const vb = reinterpret_cast<uint8_t*>(vp);
    if constexpr (FAST) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
            *reinterpret_cast<uint2*>(yb + (y0 + r) * y_stride + x0 * 2) =
                make_uint2(yv[r][0] | (yv[r][1] << 16), yv[r][2] | (yv[r][3] << 16));
        *reinterpret_cast<uint32_t*>(ub + by * c_stride + bx * 4) = uu[0] | (uu[1] << 16);
        *reinterpret_cast<uint32_t*>(vb + by * c_stride + bx * 4) = vv[0] | (vv[1] << 16);
    } else {
        for (int r = 0; r < 2 && y0 + r < h; ++r)
            for (int i = 0; i < 4 && x0 + i < w; ++i)
                *reinterpret_cast<uint16_t*>(yb + (y0 + r) * y_stride + (x0 + i) * 2) = static_cast<uint16_t>(yv[r][i]);
        for (int j = 0; j < 2 && x0 + 2 * j < w; ++j) {
            *reinterpret_cast<uint16_t*>(ub + by * c_stride + (bx * 2 + j) * 2) = static_cast<uint16_t>(uu[j]);
            *reinterpret_cast<uint16_t*>(vb + by * c_stride + (bx * 2 + j) * 2) = static_cast<uint16_t>(vv[j]);
        }
    }
}

}  // namespace

cudaError_t launch_rgb_to_yuv420p10(cudaStream_t st, const uint8_t* rgb, long long rgb_stride, int w, int h,
                                    uint16_t* y, long long y_stride, uint16_t* u, uint16_t* v, long long c_stride,
                                    const YuvCoeffs& c) {
    const dim3 block(32, 8);
    const dim3 grid((w + 4 * 32 - 1) / (4 * 32), (h + 2 * 8 - 1) / (2 * 8));
    const auto al = [](const void* p, long long s, int a) {
        return (reinterpret_cast<uintptr_t>(p) % a) == 0 && (s % a) == 0;
    };
    const bool fast = (w % 4 == 0) && (h % 2 == 0) && al(rgb, rgb_stride, 4) && al(y, y_stride, 8) &&
                      al(u, c_stride, 4) && al(v, c_stride, 4);
    if (fast)
        rgb_to_yuv420p10_kernel<true><<<grid, block, 0, st>>>(rgb, rgb_stride, w, h, y, y_stride, u, v, c_stride, c);
    else
        rgb_to_yuv420p10_kernel<false><<<grid, block, 0, st>>>(rgb, rgb_stride, w, h, y, y_stride, u, v, c_stride, c);
    return cudaGetLastError();
}

}  // namespace reve
