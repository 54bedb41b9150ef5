#include "png_io.h"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace reve_host {

namespace {
const uint8_t kSig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};

uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
void put32(uint8_t* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }

bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (n < 0) { std::fclose(f); return false; }
    out.resize(static_cast<size_t>(n));
    const bool ok = n == 0 || std::fread(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok;
}

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace

bool png_read(const std::string& path, Image& img, std::string& err) {
    std::vector<uint8_t> file;
    if (!read_file(path, file)) { err = "cannot read " + path; return false; }
    if (file.size() < 8 || std::memcmp(file.data(), kSig, 8) != 0) { err = path + ": not a PNG"; return false; }
    size_t off = 8;
    int w = 0, h = 0, depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, plte;
    bool end = false;
    while (!end && off + 12 <= file.size()) {
        const uint32_t len = be32(&file[off]);
        const uint8_t* type = &file[off + 4];
        if (off + 12 + len > file.size()) { err = path + ": truncated chunk"; return false; }
        const uint8_t* data = &file[off + 8];
        if (!std::memcmp(type, "IHDR", 4) && len == 13) {
            w = static_cast<int>(be32(data)); h = static_cast<int>(be32(data + 4));
            depth = data[8]; ctype = data[9]; interlace = data[12];
        } else if (!std::memcmp(type, "PLTE", 4)) {
            plte.assign(data, data + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            end = true;
        }
        off += 12 + len;
    }
    if (w <= 0 || h <= 0 || w > 32768 || h > 32768) { err = path + ": bad IHDR"; return false; }
    if (depth != 8 || interlace != 0) { err = path + ": only 8-bit non-interlaced PNGs are supported"; return false; }
    int ch;
    switch (ctype) {
        case 0: ch = 1; break; case 2: ch = 3; break; case 3: ch = 1; break; case 4: ch = 2; break; case 6: ch = 4; break;
        default: err = path + ": unknown colour type"; return false;
    }
    const size_t stride = static_cast<size_t>(w) * ch;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rawlen = raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), idat.size()) != Z_OK || rawlen != raw.size()) {
        err = path + ": zlib inflate failed";
        return false;
    }
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    img.w = w; img.h = h;
    img.rgb.resize(static_cast<size_t>(w) * h * 3);
    for (int y = 0; y < h; ++y) {
        const uint8_t* row = &raw[(stride + 1) * y];
        const int ft = row[0];
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= static_cast<size_t>(ch) ? cur[i - ch] : 0, b = prev[i];
            const int c = i >= static_cast<size_t>(ch) ? prev[i - ch] : 0;
            int v = row[1 + i];
            switch (ft) {
                case 0: break; case 1: v += a; break; case 2: v += b; break; case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: err = path + ": bad filter type"; return false;
            }
            cur[i] = static_cast<uint8_t>(v);
        }
        uint8_t* o = &img.rgb[static_cast<size_t>(y) * w * 3];
        for (int x = 0; x < w; ++x) {
            const uint8_t* s = &cur[static_cast<size_t>(x) * ch];
            if (ctype == 2 || ctype == 6) { o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; }
            else if (ctype == 3) {
                if (static_cast<size_t>(s[0]) * 3 + 2 >= plte.size()) { err = path + ": palette index out of range"; return false; }
                o[0] = plte[s[0] * 3]; o[1] = plte[s[0] * 3 + 1]; o[2] = plte[s[0] * 3 + 2];
            } else { o[0] = o[1] = o[2] = s[0]; }  // gray -> RGB like the upstream loader
            o += 3;
        }
        prev.swap(cur);
    }
    return true;
}

bool png_write(const std::string& path, const uint8_t* rgb, int w, int h, size_t stride, std::string& err, int zlevel) {
    const size_t row = static_cast<size_t>(w) * 3;
    std::vector<uint8_t> raw((row + 1) * h);
    for (int y = 0; y < h; ++y) {
        uint8_t* o = &raw[(row + 1) * y];
        const uint8_t* s = rgb + static_cast<size_t>(y) * stride;
        o[0] = y ? 2 : 0;  // "up" filter: cheap and effective on upscaled frames
        if (y == 0) std::memcpy(o + 1, s, row);
        else {
            const uint8_t* p = s - stride;
            for (size_t i = 0; i < row; ++i) o[1 + i] = static_cast<uint8_t>(s[i] - p[i]);
        }
    }
    uLongf clen = compressBound(raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), raw.size(), zlevel) != Z_OK) { err = "zlib deflate failed"; return false; }
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) { err = "cannot write " + path; return false; }
    auto chunk = [&](const char* type, const uint8_t* data, uint32_t len) {
        uint8_t hdr[8];
        put32(hdr, len);
        std::memcpy(hdr + 4, type, 4);
        std::fwrite(hdr, 1, 8, f);
        if (len) std::fwrite(data, 1, len, f);
        uLong crc = crc32(0L, reinterpret_cast<const Bytef*>(type), 4);
        if (len) crc = crc32(crc, data, len);
        uint8_t c[4];
        put32(c, static_cast<uint32_t>(crc));
        std::fwrite(c, 1, 4, f);
    };
    std::fwrite(kSig, 1, 8, f);
    uint8_t ihdr[13];
    put32(ihdr, w); put32(ihdr + 4, h);
    ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", comp.data(), static_cast<uint32_t>(clen));
    chunk("IEND", nullptr, 0);
    const bool ok = std::ferror(f) == 0;
    std::fclose(f);
    if (!ok) err = "write failed: " + path;
    return ok;
}

}  // namespace reve_host
