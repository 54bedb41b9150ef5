// reve-upscale: C++ host driver over the C ABI (include/reve_cuda.h).
//
// Mirrors what Video::upscale_segment reaches today (reference reve-shared/src/lib.rs:129-155):
//     realesrgan-ncnn-vulkan -i <in dir> -o <out dir> -n realesr-animevideov3-x2 -s <scale> -f png -v
// Same argv contract (plus -m model dir, -t tile, -g gpu list, -j ignored): every frame file of the
// input directory, sorted by name, is upscaled into the output directory under the same stem, and
// with -v one line "<in> -> <out> done" per finished frame goes to stderr -- which is what
// reve-cli/src/main.rs:265-273 counts for its progress bar.  Differences from the spawned upstream
// binary, all deliberate: the model matching -s is loaded, a failure ends with
// an "error: ..." line and a non-zero exit code instead of being ignored, and with several GPUs
// (-g 0,1,..) the frames of the segment are dealt round-robin to one context per GPU (host thread
// each; no collective).  The Rust crate of INTEGRATION.md does the same in-process.
#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/reve_cuda.h"
#include "png_io.h"

namespace {

struct Options {
    std::string in, out, model_dir = "models", name = "realesr-animevideov3", fmt = "png";
    int scale = 2, tile = 200, prepad = 10;
    std::vector<int> gpus;
    bool verbose = false;
    uint64_t seed = 1234;  // random-init fallback when the weight files are absent offline
};

bool ends_with(const std::string& s, const char* suf) {
    const size_t n = std::strlen(suf);
    if (s.size() < n) return false;
    std::string t = s.substr(s.size() - n);
    std::transform(t.begin(), t.end(), t.begin(), ::tolower);
    return t == suf;
}

int usage() {
    std::fprintf(stderr,
                 "usage: reve-upscale -i in_dir -o out_dir [-s 2|3|4] [-n model-name] [-m model-dir] [-t tile]\n"
                 "                    [-g gpu,gpu,...] [-f png] [-v]\n");
    return 2;
}

std::mutex g_err_mutex;
std::string g_err;
void set_error(const std::string& e) {
    std::lock_guard<std::mutex> l(g_err_mutex);
    if (g_err.empty()) g_err = e;
}

// One GPU: frames idx = first, first+step, ... of `names`
void worker(const Options& o, const reve_model* model, int device, const std::vector<std::string>& names, size_t first,
            size_t step, int w, int h, std::atomic<bool>& failed) {
    reve_ctx* ctx = nullptr;
    const int depth = 3;
    if (reve_ctx_create(device, model, w, h, o.tile, o.prepad, depth, &ctx) != REVE_OK) {
        set_error(reve_last_error(nullptr));
        failed = true;
        return;
    }
    const size_t in_bytes = size_t(w) * h * 3, out_bytes = in_bytes * o.scale * o.scale;
    std::vector<uint8_t*> hin(depth, nullptr), hout(depth, nullptr);
    for (int i = 0; i < depth; ++i) {
        if (reve_host_alloc(in_bytes, reinterpret_cast<void**>(&hin[i])) != REVE_OK ||
            reve_host_alloc(out_bytes, reinterpret_cast<void**>(&hout[i])) != REVE_OK) {
            set_error(reve_last_error(nullptr));
            failed = true;
        }
    }
    struct Pending { int slot; std::string src, dst; };
    std::vector<Pending> pending;
    auto retire = [&]() {
        uint64_t tag = 0;
        if (reve_wait(ctx, &tag) != REVE_OK) { set_error(reve_last_error(ctx)); failed = true; return; }
        const Pending pd = pending.front();
        pending.erase(pending.begin());
        std::string err;
        if (!reve_host::png_write(pd.dst, hout[pd.slot], w * o.scale, h * o.scale, size_t(w) * o.scale * 3, err)) {
            set_error(err);
            failed = true;
            return;
        }
        if (o.verbose) std::fprintf(stderr, "%s -> %s done\n", pd.src.c_str(), pd.dst.c_str());
    };
    size_t k = 0;
    for (size_t idx = first; idx < names.size() && !failed; idx += step, ++k) {
        const int slot = static_cast<int>(k % depth);
        if (static_cast<int>(pending.size()) == depth) retire();
        if (failed) break;
        const std::string src = o.in + "/" + names[idx];
        reve_host::Image img;
        std::string err;
        if (!reve_host::png_read(src, img, err)) { set_error(err); failed = true; break; }
        if (img.w != w || img.h != h) { set_error(src + ": frame size differs from the first frame of the segment"); failed = true; break; }
        std::memcpy(hin[slot], img.rgb.data(), in_bytes);
        const std::string stem = names[idx].substr(0, names[idx].find_last_of('.'));
        const std::string dst = o.out + "/" + stem + "." + o.fmt;
        if (reve_submit(ctx, hin[slot], size_t(w) * 3, hout[slot], size_t(w) * o.scale * 3, idx) != REVE_OK) {
            set_error(reve_last_error(ctx));
            failed = true;
            break;
        }
        pending.push_back({slot, src, dst});
    }
    while (!pending.empty() && !failed) retire();
    reve_sync(ctx);
    for (int i = 0; i < depth; ++i) { reve_host_free(hin[i]); reve_host_free(hout[i]); }
    reve_ctx_destroy(ctx);
}

}  // namespace

int main(int argc, char** argv) {
    // test hook (no GPU needed): decode a PNG and re-encode it with this file's codec
    if (argc == 4 && std::string(argv[1]) == "--png-roundtrip") {
        reve_host::Image img;
        std::string e;
        if (!reve_host::png_read(argv[2], img, e) ||
            !reve_host::png_write(argv[3], img.rgb.data(), img.w, img.h, size_t(img.w) * 3, e)) {
            std::fprintf(stderr, "error: %s\n", e.c_str());
            return 1;
        }
        return 0;
    }
    Options o;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> const char* { return (i + 1 < argc) ? argv[++i] : nullptr; };
        if (a == "-v") o.verbose = true;
        else if (a == "-x") { std::fprintf(stderr, "error: TTA (-x) is not supported\n"); return 2; }
        else if (a == "-i" || a == "-o" || a == "-s" || a == "-n" || a == "-m" || a == "-t" || a == "-g" || a == "-f" || a == "-j") {
            const char* v = next();
            if (!v) return usage();
            if (a == "-i") o.in = v; else if (a == "-o") o.out = v; else if (a == "-s") o.scale = std::atoi(v);
            else if (a == "-n") o.name = v; else if (a == "-m") o.model_dir = v; else if (a == "-t") o.tile = std::atoi(v);
            else if (a == "-f") o.fmt = v;
            else if (a == "-g") {
                for (const char* p = v; *p;) { o.gpus.push_back(std::atoi(p)); while (*p && *p != ',') ++p; if (*p) ++p; }
            }
        } else return usage();
    }
    if (o.in.empty() || o.out.empty()) return usage();
    if (o.scale < 2 || o.scale > 4) { std::fprintf(stderr, "error: scale must be 2, 3 or 4\n"); return 2; }
    if (o.fmt != "png") { std::fprintf(stderr, "error: only -f png is supported\n"); return 2; }
    if (o.tile < 0) o.tile = 200;

    std::vector<std::string> names;
    if (DIR* d = opendir(o.in.c_str())) {
        while (dirent* e = readdir(d)) if (ends_with(e->d_name, ".png")) names.push_back(e->d_name);
        closedir(d);
    } else { std::fprintf(stderr, "error: cannot open input directory %s\n", o.in.c_str()); return 1; }
    std::sort(names.begin(), names.end());
    for (size_t pos = 1; pos <= o.out.size(); ++pos)   // mkdir -p (the reference creates only the leaf, lib.rs:130-132)
        if (pos == o.out.size() || o.out[pos] == '/') mkdir(o.out.substr(0, pos).c_str(), 0777);
    if (names.empty()) return 0;

    // upstream appends "-x<scale>" only to the bare name; reve passes "...-x2" whatever -s is (lib.rs:141)
    std::string base = o.name;
    const size_t px = base.rfind("-x");
    if (px != std::string::npos && px + 3 == base.size()) base = base.substr(0, px);
    const std::string stem = o.model_dir + "/" + base + "-x" + std::to_string(o.scale);
    reve_model* model = nullptr;
    struct stat st;
    if (stat((stem + ".param").c_str(), &st) == 0 && stat((stem + ".bin").c_str(), &st) == 0) {
        if (reve_model_load_ncnn((stem + ".param").c_str(), (stem + ".bin").c_str(), &model) != REVE_OK) {
            std::fprintf(stderr, "error: %s\n", reve_last_error(nullptr));
            return 1;
        }
    } else {
        std::fprintf(stderr, "warning: %s.param/.bin not found, using the seeded random init of the architecture\n", stem.c_str());
        if (reve_model_random(o.scale, o.seed, &model) != REVE_OK) { std::fprintf(stderr, "error: %s\n", reve_last_error(nullptr)); return 1; }
    }
    reve_host::Image first;
    std::string err;
    if (!reve_host::png_read(o.in + "/" + names[0], first, err)) { std::fprintf(stderr, "error: %s\n", err.c_str()); return 1; }
    if (o.gpus.empty()) o.gpus.push_back(0);

    std::atomic<bool> failed{false};
    std::vector<std::thread> threads;
    for (size_t g = 0; g < o.gpus.size(); ++g)
        threads.emplace_back(worker, std::cref(o), model, o.gpus[g], std::cref(names), g, o.gpus.size(), first.w, first.h, std::ref(failed));
    for (auto& t : threads) t.join();
    reve_model_free(model);
    if (failed) { std::fprintf(stderr, "error: %s\n", g_err.c_str()); return 1; }
    return 0;
}
