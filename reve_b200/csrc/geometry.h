// Canvas geometry: how upstream's tile / pre-pad border semantics
// are laid out for the GPU.
//
// Upstream runs the network once per padded tile (tile T + P px per side; real neighbours inside
// the frame, reflect-101 beyond it; zero SAME padding at the padded tile's own border in every
// conv).  Here all padded tiles of a frame are placed side by side on one "canvas", separated by
// one-pixel gap columns/rows that are forced to zero after every layer, so one strip-mined
// convolution over the whole canvas computes every tile with exactly upstream's borders:
//
//      canvas x:  [ tile0: tw0+2P px ][gap][ tile1: tw1+2P px ][gap] ...
//
// The outer canvas border needs no gap: TMA out-of-bounds reads return zero.  tile == 0 (whole
// frame) is the one-tile case.  Per canvas column/row the tables give
//   src[i]  source frame coordinate feeding that canvas pixel (reflect-101 applied), -1 = gap
//   out[i]  output coordinate at input resolution if the pixel survives the P*s crop, else -1
#pragma once
#include <string>
#include <vector>

namespace reve {

struct Axis {
    int n = 0;              // canvas extent along this axis
    std::vector<int> src;   // [n]
    std::vector<int> out;   // [n]
};

struct Geometry {
    int in_w = 0, in_h = 0, scale = 0, tile = 0, prepad = 0;
    Axis x, y;
    int canvas_w() const { return x.n; }
    int canvas_h() const { return y.n; }
};

// Returns 0 or a negative reve_status (+ message in err).
int make_geometry(int in_w, int in_h, int scale, int tile, int prepad, Geometry& g, std::string& err);

}  // namespace reve
