// Model loader for realesr-animevideov3 (SRVGGNetCompact): ncnn .param/.bin parser + validator,
// writer, and the seeded random init used when the weight files are absent offline.
// File format as recorded in SURVEY.md section 8(a) row D (the files ship with the upscaler the
// reference spawns, reference README.md:27-29; they are not in the reference tree).
#include "model.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

#include "../../include/reve_cuda.h"

namespace reve {

// ------------------------------------------------------------------------------------ fp16
uint16_t f32_to_f16(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t mant = x & 0x7FFFFFu;
    const int exp = static_cast<int>((x >> 23) & 0xFF);
    if (exp == 0xFF) return static_cast<uint16_t>(sign | 0x7C00u | (mant ? 0x200u : 0));  // inf/nan
    int e = exp - 127 + 15;
    if (e >= 0x1F) return static_cast<uint16_t>(sign | 0x7C00u);  // overflow -> inf
    if (e <= 0) {                                                 // subnormal / zero
        if (e < -10) return static_cast<uint16_t>(sign);
        const uint32_t m = mant | 0x800000u;
        const int shift = 14 - e;  // 14..24
        uint32_t h = m >> shift;
        const uint32_t rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1))) ++h;
        return static_cast<uint16_t>(sign | h);
    }
    uint32_t h = (static_cast<uint32_t>(e) << 10) | (mant >> 13);
    const uint32_t rem = mant & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;  // carries into the exponent correctly
    return static_cast<uint16_t>(sign | h);
}

float f16_to_f32(uint16_t h) {
    const uint32_t sign = static_cast<uint32_t>(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1F, mant = h & 0x3FFu, x;
    if (exp == 0) {
        if (mant == 0) {
            x = sign;
        } else {
            int e = -1;
            do { ++e; mant <<= 1; } while (!(mant & 0x400u));
            x = sign | (static_cast<uint32_t>(127 - 15 - e) << 23) | ((mant & 0x3FFu) << 13);
        }
    } else if (exp == 0x1F) {
        x = sign | 0x7F800000u | (mant << 13);
    } else {
        x = sign | ((exp + 127 - 15) << 23) | (mant << 13);
    }
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}

// ------------------------------------------------------------------------------------ random init
// splitmix64 in counter form, Irwin-Hall(12) normals: no libm, bit-identical to
// oracle/srvgg.py (_splitmix64/_uniform24/_normal/make_weights).
namespace {
constexpr uint64_t kGamma = 0x9E3779B97F4A7C15ull;
struct Stream {
    uint64_t s0, i = 0;
    Stream(uint64_t seed, uint64_t stream) : s0(seed ^ ((stream + 1) * kGamma)) {}
    uint64_t next() {
        uint64_t z = s0 + (++i) * kGamma;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uniform24() { return static_cast<double>(next() >> 40) / 16777216.0; }
    double normal() {
        double acc = 0.0;
        for (int j = 0; j < 12; ++j) acc = acc + uniform24();
        return acc - 6.0;
    }
};
void layer_shape(int scale, int k, int& co, int& ci) {
    co = (k == kNumConv - 1) ? 3 * scale * scale : kNumFeat;
    ci = (k == 0) ? 3 : kNumFeat;
}
}  // namespace

int model_random(int scale, uint64_t seed, Model& m, std::string& err) {
    if (scale < 2 || scale > 4) {
        err = "scale must be 2, 3 or 4";
        return REVE_E_INVAL;
    }
    m.scale = scale;
    for (int k = 0; k < kNumConv; ++k) {
        ConvLayer& L = m.conv[k];
        layer_shape(scale, k, L.out_ch, L.in_ch);
        const int fan_in = L.in_ch * 9;
        double std_ = std::sqrt(2.0 / ((1.0 + 0.25 * 0.25) * fan_in));
        if (k == kNumConv - 1) std_ *= 0.3;
        Stream sw(seed, 4 * k + 0), sb(seed, 4 * k + 1), ss(seed, 4 * k + 2);
        L.w.resize(static_cast<size_t>(L.out_ch) * L.in_ch * 9);
        for (float& v : L.w) v = f16_to_f32(f32_to_f16(static_cast<float>(sw.normal() * std_)));
        L.b.resize(L.out_ch);
        for (float& v : L.b) v = static_cast<float>(sb.normal() * 0.01);
        L.slope.clear();
        if (k < kNumConv - 1) {
            L.slope.resize(L.out_ch);
            for (float& v : L.slope) v = static_cast<float>(0.25 + (ss.uniform24() - 0.5) * 0.1);
        }
    }
    return REVE_OK;
}

// ------------------------------------------------------------------------------------ ncnn parsing
namespace {
constexpr int kMagic = 7767517;
constexpr uint32_t kTagFp16 = 0x01306B47u;

struct PLayer {
    std::string type, name;
    std::vector<std::string> in, out;
    std::map<int, std::string> kv;
    int geti(int key, int def) const {
        auto it = kv.find(key);
        return it == kv.end() ? def : std::atoi(it->second.c_str());
    }
    double getf(int key, double def) const {
        auto it = kv.find(key);
        return it == kv.end() ? def : std::atof(it->second.c_str());
    }
};

int fail(std::string& err, int code, const std::string& msg) {
    err = msg;
    return code;
}
}  // namespace

int model_load_ncnn(const std::string& param_path, const std::string& bin_path, Model& m, std::string& err) {
    std::ifstream pf(param_path);
    if (!pf) return fail(err, REVE_E_IO, "cannot open " + param_path);
    int magic = 0, nlayer = 0, nblob = 0;
    if (!(pf >> magic) || magic != kMagic) return fail(err, REVE_E_MODEL, param_path + ": bad ncnn magic");
    if (!(pf >> nlayer >> nblob) || nlayer <= 0 || nlayer > 4096) return fail(err, REVE_E_MODEL, param_path + ": bad layer count");
    std::string line;
    std::getline(pf, line);
    std::vector<PLayer> layers;
    while (static_cast<int>(layers.size()) < nlayer && std::getline(pf, line)) {
        std::istringstream ls(line);
        PLayer L;
        int nin = 0, nout = 0;
        if (!(ls >> L.type)) continue;  // blank line
        if (!(ls >> L.name >> nin >> nout) || nin < 0 || nout < 0 || nin > 16 || nout > 16)
            return fail(err, REVE_E_MODEL, param_path + ": malformed layer line: " + line);
        L.in.resize(nin);
        L.out.resize(nout);
        for (auto& s : L.in) if (!(ls >> s)) return fail(err, REVE_E_MODEL, "truncated layer line: " + line);
        for (auto& s : L.out) if (!(ls >> s)) return fail(err, REVE_E_MODEL, "truncated layer line: " + line);
        std::string tok;
        while (ls >> tok) {
            const size_t eq = tok.find('=');
            if (eq == std::string::npos) return fail(err, REVE_E_MODEL, "bad key=value token '" + tok + "'");
            L.kv[std::atoi(tok.substr(0, eq).c_str())] = tok.substr(eq + 1);
        }
        layers.push_back(std::move(L));
    }
    if (static_cast<int>(layers.size()) != nlayer) return fail(err, REVE_E_MODEL, param_path + ": fewer layers than declared");

    // Structural validation: producer map, then walk back from the BinaryOp(add).
    std::map<std::string, int> producer;
    for (int i = 0; i < nlayer; ++i)
        for (const auto& o : layers[i].out) producer[o] = i;
    auto prod = [&](const std::string& blob) -> const PLayer* {
        auto it = producer.find(blob);
        return it == producer.end() ? nullptr : &layers[it->second];
    };
    auto root_of = [&](std::string blob) -> const PLayer* {  // skip Split layers
        const PLayer* p = prod(blob);
        while (p && p->type == "Split" && p->in.size() == 1) p = prod(p->in[0]);
        return p;
    };
    const PLayer* add = nullptr;
    for (const auto& L : layers) {
        static const char* known[] = {"Input", "Split", "Convolution", "PReLU", "PixelShuffle", "Interp", "BinaryOp"};
        bool ok = false;
        for (const char* k : known) ok = ok || L.type == k;
        if (!ok) return fail(err, REVE_E_MODEL, "unsupported layer type '" + L.type + "' (not an SRVGGNetCompact graph)");
        if (L.type == "BinaryOp") {
            if (add) return fail(err, REVE_E_MODEL, "more than one BinaryOp");
            add = &L;
        }
    }
    if (!add || add->in.size() != 2 || add->geti(0, 0) != 0 || add->geti(1, 0) != 0)
        return fail(err, REVE_E_MODEL, "missing residual BinaryOp(add)");
    const PLayer *ps = nullptr, *up = nullptr;
    for (int i = 0; i < 2; ++i) {
        const PLayer* p = root_of(add->in[i]);
        if (p && p->type == "PixelShuffle") ps = p;
        if (p && p->type == "Interp") up = p;
    }
    if (!ps || !up) return fail(err, REVE_E_MODEL, "residual add must join PixelShuffle and Interp");
    const int scale = ps->geti(0, 1);
    if (scale < 2 || scale > 4 || ps->geti(1, 0) != 0) return fail(err, REVE_E_MODEL, "PixelShuffle factor must be 2, 3 or 4 (mode 0)");
    if (up->geti(0, 0) != 1 || std::fabs(up->getf(1, 1.0) - scale) > 1e-6 || std::fabs(up->getf(2, 1.0) - scale) > 1e-6)
        return fail(err, REVE_E_MODEL, "Interp must be nearest with the PixelShuffle factor");
    const PLayer* in_up = up->in.size() == 1 ? root_of(up->in[0]) : nullptr;
    if (!in_up || in_up->type != "Input") return fail(err, REVE_E_MODEL, "Interp must read the network input");
    // walk the conv/PReLU chain backwards
    std::vector<const PLayer*> chain;
    const PLayer* cur = ps->in.size() == 1 ? root_of(ps->in[0]) : nullptr;
    while (cur && cur->type != "Input") {
        if ((cur->type != "Convolution" && cur->type != "PReLU") || cur->in.size() != 1)
            return fail(err, REVE_E_MODEL, "unexpected layer '" + cur->name + "' in the conv chain");
        chain.push_back(cur);
        cur = root_of(cur->in[0]);
        if (chain.size() > 64) break;
    }
    if (!cur || cur != in_up) return fail(err, REVE_E_MODEL, "conv chain does not start at the network input");
    if (chain.size() != static_cast<size_t>(2 * kNumConv - 1)) return fail(err, REVE_E_MODEL, "expected 18 Convolution + 17 PReLU layers");
    for (size_t i = 0; i < chain.size(); ++i) {
        const bool want_conv = (i % 2 == 0);  // reversed: conv17, prelu16, conv16, ...
        if ((chain[i]->type == "Convolution") != want_conv) return fail(err, REVE_E_MODEL, "Convolution/PReLU do not alternate");
    }

    // Read the .bin sequentially in layer order.
    std::ifstream bf(bin_path, std::ios::binary);
    if (!bf) return fail(err, REVE_E_IO, "cannot open " + bin_path);
    std::vector<char> data((std::istreambuf_iterator<char>(bf)), std::istreambuf_iterator<char>());
    size_t off = 0;
    auto need = [&](size_t n) { return off + n <= data.size(); };
    m = Model();
    m.scale = scale;
    int ci_conv = 0, ci_prelu = 0;
    for (const auto& L : layers) {
        if (L.type == "Convolution") {
            if (ci_conv >= kNumConv) return fail(err, REVE_E_MODEL, "too many convolutions");
            ConvLayer& C = m.conv[ci_conv];
            int co, ci;
            layer_shape(scale, ci_conv, co, ci);
            const int kw = L.geti(1, 0), kh = L.geti(11, kw);
            const int dil = L.geti(2, 1), dilh = L.geti(12, dil), st = L.geti(3, 1), sth = L.geti(13, st);
            const int pad = L.geti(4, 0), padr = L.geti(15, pad), padt = L.geti(14, pad), padb = L.geti(16, padt);
            if (kw != 3 || kh != 3 || dil != 1 || dilh != 1 || st != 1 || sth != 1 || !(pad == 1 || pad == -233) ||
                padr != pad || padt != pad || padb != pad || L.geti(8, 0) != 0 || L.geti(9, 0) != 0)
                return fail(err, REVE_E_MODEL, L.name + ": expected 3x3 stride 1 pad 1 fp convolution without fused activation");
            if (L.geti(0, 0) != co || L.geti(6, 0) != co * ci * 9 || L.geti(5, 0) != 1)
                return fail(err, REVE_E_MODEL, L.name + ": channel counts do not match SRVGGNetCompact(64 feat)");
            C.out_ch = co;
            C.in_ch = ci;
            const size_t n = static_cast<size_t>(co) * ci * 9;
            if (!need(4)) return fail(err, REVE_E_MODEL, bin_path + ": truncated");
            uint32_t tag;
            std::memcpy(&tag, &data[off], 4);
            off += 4;
            C.w.resize(n);
            if (tag == kTagFp16) {
                const size_t bytes = (2 * n + 3) / 4 * 4;
                if (!need(bytes)) return fail(err, REVE_E_MODEL, bin_path + ": truncated");
                for (size_t i = 0; i < n; ++i) {
                    uint16_t h;
                    std::memcpy(&h, &data[off + 2 * i], 2);
                    C.w[i] = f16_to_f32(h);
                }
                off += bytes;
            } else if (tag == 0) {
                if (!need(4 * n)) return fail(err, REVE_E_MODEL, bin_path + ": truncated");
                std::memcpy(C.w.data(), &data[off], 4 * n);
                off += 4 * n;
            } else {
                char buf[64];
                std::snprintf(buf, sizeof buf, "unsupported weight tag 0x%08X", tag);
                return fail(err, REVE_E_MODEL, L.name + ": " + buf);
            }
            if (!need(4 * static_cast<size_t>(co))) return fail(err, REVE_E_MODEL, bin_path + ": truncated");
            C.b.resize(co);
            std::memcpy(C.b.data(), &data[off], 4 * static_cast<size_t>(co));
            off += 4 * static_cast<size_t>(co);
            ++ci_conv;
        } else if (L.type == "PReLU") {
            if (ci_prelu >= kNumConv - 1 || ci_prelu != ci_conv - 1) return fail(err, REVE_E_MODEL, "PReLU out of order");
            ConvLayer& C = m.conv[ci_prelu];
            if (L.geti(0, 0) != C.out_ch) return fail(err, REVE_E_MODEL, L.name + ": slope count != channels");
            if (!need(4 * static_cast<size_t>(C.out_ch))) return fail(err, REVE_E_MODEL, bin_path + ": truncated");
            C.slope.resize(C.out_ch);
            std::memcpy(C.slope.data(), &data[off], 4 * static_cast<size_t>(C.out_ch));
            off += 4 * static_cast<size_t>(C.out_ch);
            ++ci_prelu;
        }
    }
    if (ci_conv != kNumConv || ci_prelu != kNumConv - 1) return fail(err, REVE_E_MODEL, "layer count mismatch");
    if (off != data.size()) return fail(err, REVE_E_MODEL, bin_path + ": trailing bytes");
    // the device stores weights as fp16: an fp32-tagged payload beyond +-65504 would silently pack to +-inf
    // (reve_model_from_arrays applies the same rule); biases and slopes stay fp32 and only need to be finite
    for (int k = 0; k < kNumConv; ++k) {
        for (float v : m.conv[k].w)
            if (!std::isfinite(v) || v > 65504.f || v < -65504.f)
                return fail(err, REVE_E_MODEL, "weight of convolution " + std::to_string(k) + " is not representable in fp16");
        for (float v : m.conv[k].b)
            if (!std::isfinite(v)) return fail(err, REVE_E_MODEL, "non-finite bias in convolution " + std::to_string(k));
        for (float v : m.conv[k].slope)
            if (!std::isfinite(v)) return fail(err, REVE_E_MODEL, "non-finite PReLU slope after convolution " + std::to_string(k));
    }
    return REVE_OK;
}

int model_save_ncnn(const Model& m, const std::string& param_path, const std::string& bin_path, bool fp16, std::string& err) {
    std::ofstream pf(param_path);
    if (!pf) return fail(err, REVE_E_IO, "cannot write " + param_path);
    const int s = m.scale;
    pf << kMagic << "\n" << 40 << " " << 41 << "\n";
    pf << "Input            data                     0 1 data\n";
    pf << "Split            splitncnn_input0         1 2 data data_splitncnn_0 data_splitncnn_1\n";
    std::string prev = "data_splitncnn_1";
    for (int k = 0; k < kNumConv; ++k) {
        const ConvLayer& C = m.conv[k];
        std::string out = "conv" + std::to_string(k);
        pf << "Convolution      Conv_" << 2 * k << " 1 1 " << prev << " " << out << " 0=" << C.out_ch
           << " 1=3 11=3 2=1 12=1 3=1 13=1 4=1 14=1 15=1 16=1 5=1 6=" << C.out_ch * C.in_ch * 9 << "\n";
        prev = out;
        if (k < kNumConv - 1) {
            out = "prelu" + std::to_string(k);
            pf << "PReLU            PRelu_" << 2 * k + 1 << " 1 1 " << prev << " " << out << " 0=" << C.out_ch << "\n";
            prev = out;
        }
    }
    char buf[64];
    std::snprintf(buf, sizeof buf, "%e", static_cast<double>(s));
    pf << "PixelShuffle     DepthToSpace_36 1 1 " << prev << " ps 0=" << s << "\n";
    pf << "Interp           Resize_38 1 1 data_splitncnn_0 up 0=1 1=" << buf << " 2=" << buf << " 3=0 4=0 6=0\n";
    pf << "BinaryOp         Add_39 2 1 ps up output 0=0\n";
    if (!pf) return fail(err, REVE_E_IO, "write failed: " + param_path);
    std::ofstream bf(bin_path, std::ios::binary);
    if (!bf) return fail(err, REVE_E_IO, "cannot write " + bin_path);
    for (int k = 0; k < kNumConv; ++k) {
        const ConvLayer& C = m.conv[k];
        if (fp16) {
            const uint32_t tag = kTagFp16;
            bf.write(reinterpret_cast<const char*>(&tag), 4);
            std::vector<uint16_t> h(C.w.size() + 1, 0);
            for (size_t i = 0; i < C.w.size(); ++i) h[i] = f32_to_f16(C.w[i]);
            bf.write(reinterpret_cast<const char*>(h.data()), (2 * C.w.size() + 3) / 4 * 4);
        } else {
            const uint32_t tag = 0;
            bf.write(reinterpret_cast<const char*>(&tag), 4);
            bf.write(reinterpret_cast<const char*>(C.w.data()), 4 * C.w.size());
        }
        bf.write(reinterpret_cast<const char*>(C.b.data()), 4 * C.b.size());
        if (k < kNumConv - 1) bf.write(reinterpret_cast<const char*>(C.slope.data()), 4 * C.slope.size());
    }
    if (!bf) return fail(err, REVE_E_IO, "write failed: " + bin_path);
    return REVE_OK;
}

int model_from_arrays(int scale, const float* const* conv_w, const float* const* conv_b, const float* const* prelu,
                      Model& m, std::string& err) {
    if (scale < 2 || scale > 4) return fail(err, REVE_E_INVAL, "scale must be 2, 3 or 4");
    if (!conv_w || !conv_b || !prelu) return fail(err, REVE_E_INVAL, "NULL array table");
    m.scale = scale;
    for (int k = 0; k < kNumConv; ++k) {
        ConvLayer& C = m.conv[k];
        C.in_ch = (k == 0) ? 3 : kNumFeat;
        C.out_ch = (k == kNumConv - 1) ? 3 * scale * scale : kNumFeat;
        if (!conv_w[k] || !conv_b[k] || (k < kNumConv - 1 && !prelu[k]))
            return fail(err, REVE_E_INVAL, "NULL tensor for convolution " + std::to_string(k));
        const size_t nw = static_cast<size_t>(C.out_ch) * C.in_ch * 9;
        C.w.assign(conv_w[k], conv_w[k] + nw);
        C.b.assign(conv_b[k], conv_b[k] + C.out_ch);
        if (k < kNumConv - 1) C.slope.assign(prelu[k], prelu[k] + C.out_ch); else C.slope.clear();
        for (size_t i = 0; i < nw; ++i)
            if (!(C.w[i] == C.w[i]) || C.w[i] > 65504.f || C.w[i] < -65504.f)
                return fail(err, REVE_E_MODEL, "weight of convolution " + std::to_string(k) + " is not representable in fp16");
    }
    return REVE_OK;
}

}  // namespace reve
