// Host-side model: realesr-animevideov3 = SRVGGNetCompact(3 -> 64, 16 body convs, 64 -> 3*s*s).
// Replaces the `-n realesr-animevideov3-x{s}` + models\ lookup of the spawned upscaler
// (reference reve-shared/src/lib.rs:141, reve-gui/src-tauri/src/commands.rs:60-63).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace reve {

constexpr int kNumFeat = 64;
constexpr int kNumBody = 16;             // body convolutions
constexpr int kNumConv = kNumBody + 2;   // all convolutions

struct ConvLayer {
    int out_ch = 0, in_ch = 0;
    std::vector<float> w;      // OIHW, 3x3
    std::vector<float> b;      // [out_ch]
    std::vector<float> slope;  // [out_ch] PReLU slopes; empty for the last conv
};

struct Model {
    int scale = 0;
    ConvLayer conv[kNumConv];
};

// IEEE binary16 <-> binary32, round-to-nearest-even (host).
uint16_t f32_to_f16(float f);
float f16_to_f32(uint16_t h);

// Each returns 0 or a negative reve_status and fills `err`.
int model_load_ncnn(const std::string& param_path, const std::string& bin_path, Model& m, std::string& err);
int model_save_ncnn(const Model& m, const std::string& param_path, const std::string& bin_path, bool fp16, std::string& err);
int model_random(int scale, uint64_t seed, Model& m, std::string& err);
// conv_w[k]: OIHW fp32 of convolution k (18), conv_b[k]: bias, prelu[k]: slopes of the PReLU after convolution k (17)
int model_from_arrays(int scale, const float* const* conv_w, const float* const* conv_b, const float* const* prelu,
                      Model& m, std::string& err);

}  // namespace reve

struct reve_model {
    reve::Model m;
};
