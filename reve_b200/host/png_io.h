// Minimal PNG codec (zlib only) for the frame files that cross the reference's stage boundary:
// ffmpeg writes 8-bit RGB PNGs into temp\tmp_frames\{i} (reference reve-shared/src/lib.rs:93,
// 100-119) and reads the upscaled ones back (reve-cli/src/main.rs:297-300).  Decodes 8-bit
// gray / RGB / palette / gray+alpha / RGBA, non-interlaced, to packed RGB (alpha is dropped: it never
// occurs for ffmpeg-exported frames); encodes packed RGB.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace reve_host {

struct Image {
    int w = 0, h = 0;
    std::vector<uint8_t> rgb;  // packed RGB, row stride 3*w
};

// Both return true on success; on failure `err` says why.
bool png_read(const std::string& path, Image& img, std::string& err);
bool png_write(const std::string& path, const uint8_t* rgb, int w, int h, size_t stride, std::string& err,
               int zlevel = 1);

}  // namespace reve_host
