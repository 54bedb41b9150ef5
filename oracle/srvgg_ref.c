/* Plain-C restatement of REVE's per-segment upscale arithmetic (realesr-animevideov3).
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into libreve_cuda.so or imported by reve_b200/.
 * PARITY UNPINNED: the reference (ONdraid/reve) has no implementation of this path (it spawns
 * realesrgan-ncnn-vulkan, reference reve-shared/src/lib.rs:134-147) and no numeric test
 * (reve-cli/tests/run_test.rs:31-34 only checks that out.mp4 exists).  This file restates the
 * published upstream algorithm as recorded in SURVEY.md section 8(a):
 *   row B (tile + 10 px pre-pad, reflect-101 at the image border, zero SAME padding inside the
 *          network, crop, u8 = clamp(floor(v*255+0.5)))            -> srvgg_ref_upscale
 *   row C (SRVGGNetCompact: 18 conv3x3, 17 PReLU, PixelShuffle, nearest residual) -> net_forward
 * It is written independently of oracle/srvgg.py (scalar loops, different summation order) so
 * that the two restatements check each other.
 *
 * Packed parameter layout `params` (float32): for k = 0..17: W_k[co][ci][3][3], b_k[co], and
 * for k <= 16 additionally slope_k[co];  co/ci = (64,3), 16 x (64,64), (3*s*s,64).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NUM_FEAT 64
#define NUM_CONV 16

static int reflect101(int i, int n) {
    if (i < 0) i = -i;
    int d = i - (n - 1);
    if (d < 0) d = -d;
    return (n - 1) - d;
}

/* out[co][y][x] = b[co] + sum_ci sum_ky sum_kx W[co][ci][ky][kx] * in[ci][y+ky-1][x+kx-1],
 * zero outside the h x w plane; optional per-channel PReLU. */
static void conv3x3(const float* in, int ci_n, int co_n, int h, int w, const float* W,
                    const float* b, const float* slope, float* out) {
    const size_t plane = (size_t)h * w;
#pragma omp parallel for schedule(static)
    for (int co = 0; co < co_n; ++co) {
        float* o = out + (size_t)co * plane;
        for (size_t p = 0; p < plane; ++p) o[p] = b[co];
        for (int ci = 0; ci < ci_n; ++ci) {
            const float* ip = in + (size_t)ci * plane;
            const float* k = W + ((size_t)co * ci_n + ci) * 9;
            for (int ky = 0; ky < 3; ++ky) {
                for (int kx = 0; kx < 3; ++kx) {
                    const float wv = k[ky * 3 + kx];
                    const int dy = ky - 1, dx = kx - 1;
                    const int ya = dy < 0 ? 1 : 0, yb = dy > 0 ? h - 1 : h;
                    const int xa = dx < 0 ? 1 : 0, xb = dx > 0 ? w - 1 : w;
                    for (int y = ya; y < yb; ++y) {
                        float* orow = o + (size_t)y * w;
                        const float* irow = ip + (size_t)(y + dy) * w + dx;
                        for (int x = xa; x < xb; ++x) orow[x] += wv * irow[x];
                    }
                }
            }
        }
        if (slope) {
            const float a = slope[co];
            for (size_t p = 0; p < plane; ++p) {
                const float v = o[p];
                o[p] = (v > 0.f ? v : 0.f) + a * (v < 0.f ? v : 0.f);
            }
        }
    }
}

/* x: [3][h][w] in [0,1]; y: [3][h*s][w*s].  Returns 0, or -1 on allocation failure. */
static int net_forward(const float* x, int h, int w, int s, const float* params, float* y) {
    const size_t plane = (size_t)h * w;
    float* fa = (float*)malloc(sizeof(float) * NUM_FEAT * plane);
    float* fb = (float*)malloc(sizeof(float) * NUM_FEAT * plane);
    if (!fa || !fb) { free(fa); free(fb); return -1; }
    const float* p = params;
    const float* in = x;
    float* cur = fa;
    int ci = 3;
    for (int k = 0; k <= NUM_CONV; ++k) {
        const float* W = p;  p += (size_t)NUM_FEAT * ci * 9;
        const float* b = p;  p += NUM_FEAT;
        const float* a = p;  p += NUM_FEAT;
        conv3x3(in, ci, NUM_FEAT, h, w, W, b, a, cur);
        in = cur;
        cur = (cur == fa) ? fb : fa;
        ci = NUM_FEAT;
    }
    const int co = 3 * s * s;
    const float* W = p;  p += (size_t)co * NUM_FEAT * 9;
    const float* b = p;
    float* r = cur; /* co <= 48 < 64 planes: fits */
    conv3x3(in, NUM_FEAT, co, h, w, W, b, NULL, r);
    /* PixelShuffle (y[c][Y*s+i][X*s+j] = r[c*s*s+i*s+j][Y][X]) + nearest-upsampled input */
    const int ow = w * s;
    for (int c = 0; c < 3; ++c)
        for (int Y = 0; Y < h; ++Y)
            for (int i = 0; i < s; ++i)
                for (int X = 0; X < w; ++X)
                    for (int j = 0; j < s; ++j)
                        y[((size_t)c * h * s + (size_t)Y * s + i) * ow + (size_t)X * s + j] =
                            r[(size_t)(c * s * s + i * s + j) * plane + (size_t)Y * w + X] +
                            x[(size_t)c * plane + (size_t)Y * w + X];
    free(fa);
    free(fb);
    return 0;
}

/* frame: u8 [h][w][3]; out: u8 [h*s][w*s][3].  tile <= 0: one whole-frame tile.
 * Returns 0 on success, -1 on bad arguments / allocation failure. */
int srvgg_ref_upscale(const uint8_t* frame, int w, int h, int s, int tile, int prepad,
                      const float* params, uint8_t* out) {
    if (!frame || !out || !params || w < 1 || h < 1 || s < 2 || s > 4) return -1;
    if (prepad < 0 || prepad > (w < h ? w : h) - 1) return -1;
    if (tile <= 0) tile = w > h ? w : h;
    const int P = prepad;
    for (int y0 = 0; y0 < h; y0 += tile) {
        const int th = (y0 + tile <= h ? tile : h - y0);
        for (int x0 = 0; x0 < w; x0 += tile) {
            const int tw = (x0 + tile <= w ? tile : w - x0);
            const int pw = tw + 2 * P, ph = th + 2 * P;
            float* x = (float*)malloc(sizeof(float) * 3 * (size_t)pw * ph);
            float* y = (float*)malloc(sizeof(float) * 3 * (size_t)pw * ph * s * s);
            if (!x || !y) { free(x); free(y); return -1; }
            for (int c = 0; c < 3; ++c)
                for (int yy = 0; yy < ph; ++yy) {
                    const int sy = reflect101(y0 - P + yy, h);
                    for (int xx = 0; xx < pw; ++xx) {
                        const int sx = reflect101(x0 - P + xx, w);
                        x[((size_t)c * ph + yy) * pw + xx] =
                            (float)frame[((size_t)sy * w + sx) * 3 + c] * (1.0f / 255.0f);
                    }
                }
            if (net_forward(x, ph, pw, s, params, y) != 0) { free(x); free(y); return -1; }
            for (int c = 0; c < 3; ++c)
                for (int yy = 0; yy < th * s; ++yy)
                    for (int xx = 0; xx < tw * s; ++xx) {
                        float v = y[((size_t)c * ph * s + (size_t)(P * s + yy)) * ((size_t)pw * s) +
                                    (size_t)(P * s + xx)];
                        v = floorf(v * 255.0f + 0.5f);
                        v = v < 0.f ? 0.f : (v > 255.f ? 255.f : v);
                        out[((size_t)(y0 * s + yy) * ((size_t)w * s) + (size_t)(x0 * s + xx)) * 3 + c] =
                            (uint8_t)v;
                    }
            free(x);
            free(y);
        }
    }
    return 0;
}
