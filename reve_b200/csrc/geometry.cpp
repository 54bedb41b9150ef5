#include "geometry.h"

#include <cstdlib>

#include "../../include/reve_cuda.h"

namespace reve {

namespace {
// Upstream pre-processing border rule: x = |x|; x = (n-1) - |x - (n-1)|.
int reflect101(int i, int n) {
    i = std::abs(i);
    return (n - 1) - std::abs(i - (n - 1));
}

void make_axis(int n_in, int tile, int prepad, Axis& a) {
    a.src.clear();
    a.out.clear();
    const int t = tile > 0 ? tile : n_in;
    for (int t0 = 0; t0 < n_in; t0 += t) {
        const int tn = (t0 + t <= n_in) ? t : n_in - t0;
        if (t0 > 0) {  // gap between tiles
            a.src.push_back(-1);
            a.out.push_back(-1);
        }
        for (int i = -prepad; i < tn + prepad; ++i) {
            a.src.push_back(reflect101(t0 + i, n_in));
            a.out.push_back((i >= 0 && i < tn) ? t0 + i : -1);
        }
    }
    a.n = static_cast<int>(a.src.size());
}
}  // namespace

int make_geometry(int in_w, int in_h, int scale, int tile, int prepad, Geometry& g, std::string& err) {
    if (in_w < 1 || in_h < 1 || in_w > 16384 || in_h > 16384) {
        err = "frame size must be within 1..16384";
        return REVE_E_INVAL;
    }
    if (scale < 2 || scale > 4) {
        err = "scale must be 2, 3 or 4";
        return REVE_E_INVAL;
    }
    if (tile < 0) {
        err = "tile must be >= 0";
        return REVE_E_INVAL;
    }
    const int mn = in_w < in_h ? in_w : in_h;
    if (prepad < 0 || prepad > mn - 1 || prepad > 64) {
        err = "prepad must be within [0, min(w,h)-1] and <= 64 (reflect-101 is undefined beyond)";
        return REVE_E_INVAL;
    }
    g.in_w = in_w;
    g.in_h = in_h;
    g.scale = scale;
    g.tile = tile;
    g.prepad = prepad;
    make_axis(in_w, tile, prepad, g.x);
    make_axis(in_h, tile, prepad, g.y);
    return REVE_OK;
}

}  // namespace reve
