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
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <cstdio>
#include <cstdlib>
#include <cerrno>
#include <chrono>
#include <cstring>
#include <memory>
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
    int out_format = REVE_FMT_RGB24;   // --pix-fmt (raw mode): rgb24 | yuv420p10le (BT.601, as swscale) | yuv420p10le-bt709
    int raw_w = 0, raw_h = 0;  // --raw WxH: rgb24 rawvideo frames on -i (file or "-" = stdin) -> -o (file or "-" = stdout)
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
                 "                    [-g gpu,gpu,...] [-f png] [-v]\n"
                 "       reve-upscale --raw WxH -i in.rgb|- -o out.rgb|- [-s 2|3|4] [-m model-dir] [-t tile] [-g gpu] [-v]\n"
                 "                    [--pix-fmt rgb24|yuv420p10le|yuv420p10le-bt709]\n"
                 "         (rgb24 rawvideo stream in, e.g. ffmpeg -f rawvideo -pix_fmt rgb24 pipes: no PNG on the path;\n"
                 "          with --pix-fmt yuv420p10le the output is what `ffmpeg -f rawvideo -pix_fmt yuv420p10le\n"
                 "          -s WxH -i pipe:0 -c:v libx265` ingests without a swscale pass)\n");
    return 2;
}

std::mutex g_err_mutex;
std::string g_err;
void set_error(const std::string& e) {
    std::lock_guard<std::mutex> l(g_err_mutex);
    if (g_err.empty()) g_err = e;
}

// Small fixed pool of host threads for PNG decode / encode (the upstream binary also runs 1 load and 2 save
// threads around its GPU loop; 4K PNG encoding is the slowest step by far).
class Pool {
public:
    explicit Pool(int n) {
        for (int i = 0; i < n; ++i) threads_.emplace_back([this] { run(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    void submit(std::function<void()> f) {
        { std::lock_guard<std::mutex> l(m_); q_.push_back(std::move(f)); ++pending_; }
        cv_.notify_one();
    }
    void wait_below(size_t n) {  // block until fewer than n tasks are queued or running
        std::unique_lock<std::mutex> l(m_);
        done_cv_.wait(l, [&] { return pending_ < n; });
    }
private:
    void run() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                f = std::move(q_.front());
                q_.pop_front();
            }
            f();
            { std::lock_guard<std::mutex> l(m_); --pending_; }
            done_cv_.notify_all();
        }
    }
    std::vector<std::thread> threads_;
    std::deque<std::function<void()>> q_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    size_t pending_ = 0;
    bool stop_ = false;
};

// One GPU: frames idx = first, first+step, ... of `names`.  Decode pool -> pinned ring -> GPU -> encode pool.
void worker(const Options& o, const reve_model* model, int device, const std::vector<std::string>& names, size_t first,
            size_t step, int w, int h, std::atomic<bool>& failed, int host_threads) {
    reve_ctx* ctx = nullptr;
    const int depth = 8;
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
    std::vector<size_t> mine;
    for (size_t idx = first; idx < names.size(); idx += step) mine.push_back(idx);
    const auto t_ready = std::chrono::steady_clock::now();   // REVE_HOST_TIMING: the segment's frames/s without start-up

    // decode ahead of the GPU: results keyed by position in `mine`
    std::mutex dm;
    std::condition_variable dcv;
    std::map<size_t, reve_host::Image> decoded;
    const size_t window = 16;
    size_t next_decode = 0;
    {
        Pool decoders(std::max(1, host_threads / 4)), encoders(std::max(1, host_threads - host_threads / 4));
        auto schedule_decodes = [&](size_t consumed) {
            while (next_decode < mine.size() && next_decode < consumed + window) {
                const size_t k = next_decode++;
                decoders.submit([&, k] {
                    reve_host::Image img;
                    std::string err;
                    if (!reve_host::png_read(o.in + "/" + names[mine[k]], img, err)) { set_error(err); failed = true; }
                    { std::lock_guard<std::mutex> l(dm); decoded[k] = std::move(img); }
                    dcv.notify_all();
                });
            }
        };
        struct Pending { int slot; std::string src, dst; };
        std::deque<Pending> pending;
        auto retire = [&]() {
            uint64_t tag = 0;
            if (reve_wait(ctx, &tag) != REVE_OK) { set_error(reve_last_error(ctx)); failed = true; return; }
            Pending pd = pending.front();
            pending.pop_front();
            // hand a copy to the encoders so the pinned slot can be reused at once
            auto buf = std::make_shared<std::vector<uint8_t>>(hout[pd.slot], hout[pd.slot] + out_bytes);
            encoders.wait_below(window);
            encoders.submit([&, buf, pd] {
                std::string err;
                if (!reve_host::png_write(pd.dst, buf->data(), w * o.scale, h * o.scale, size_t(w) * o.scale * 3, err)) {
                    set_error(err);
                    failed = true;
                    return;
                }
                if (o.verbose) std::fprintf(stderr, "%s -> %s done\n", pd.src.c_str(), pd.dst.c_str());
            });
        };
        schedule_decodes(0);
        for (size_t k = 0; k < mine.size() && !failed; ++k) {
            const int slot = static_cast<int>(k % depth);
            if (static_cast<int>(pending.size()) == depth) retire();
            if (failed) break;
            reve_host::Image img;
            {
                std::unique_lock<std::mutex> l(dm);
                dcv.wait(l, [&] { return decoded.count(k) != 0 || failed.load(); });
                if (failed) break;
                img = std::move(decoded[k]);
                decoded.erase(k);
            }
            schedule_decodes(k + 1);
            const std::string src = o.in + "/" + names[mine[k]];
            if (img.w != w || img.h != h) { set_error(src + ": frame size differs from the first frame of the segment"); failed = true; break; }
            std::memcpy(hin[slot], img.rgb.data(), in_bytes);
            const std::string stem = names[mine[k]].substr(0, names[mine[k]].find_last_of('.'));
            const std::string dst = o.out + "/" + stem + "." + o.fmt;
            if (reve_submit(ctx, hin[slot], size_t(w) * 3, hout[slot], size_t(w) * o.scale * 3, mine[k]) != REVE_OK) {
                set_error(reve_last_error(ctx));
                failed = true;
                break;
            }
            pending.push_back({slot, src, dst});
        }
        while (!pending.empty() && !failed) retire();
        // Pool destructors drain the remaining decode / encode tasks
    }
    reve_sync(ctx);
    if (std::getenv("REVE_HOST_TIMING"))
        std::fprintf(stderr, "[timing] gpu %d: %zu frames decoded, upscaled and encoded in %.1f ms after start-up\n", device, mine.size(),
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_ready).count());
    for (int i = 0; i < depth; ++i) { reve_host_free(hin[i]); reve_host_free(hout[i]); }
    reve_ctx_destroy(ctx);
}

// Blocking FIFO of ring-slot indices between the raw-mode threads (-1 = end of stream / failure).
class SlotQueue {
public:
    void push(int v) {
        { std::lock_guard<std::mutex> l(m_); q_.push_back(v); }
        cv_.notify_one();
    }
    int pop() {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return !q_.empty(); });
        const int v = q_.front();
        q_.pop_front();
        return v;
    }
    bool try_pop(int& v) {
        std::lock_guard<std::mutex> l(m_);
        if (q_.empty()) return false;
        v = q_.front();
        q_.pop_front();
        return true;
    }
private:
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<int> q_;
};

bool read_full(int fd, uint8_t* p, size_t n, size_t& got) {
    got = 0;
    while (got < n) {
        const ssize_t r = ::read(fd, p + got, n - got);
        if (r == 0) return true;                 // end of stream
        if (r < 0) { if (errno == EINTR) continue; return false; }
        got += static_cast<size_t>(r);
    }
    return true;
}
bool write_full(int fd, const uint8_t* p, size_t n) {
    while (n > 0) {
        const ssize_t r = ::write(fd, p, n);
        if (r < 0) { if (errno == EINTR) continue; return false; }
        p += r;
        n -= static_cast<size_t>(r);
    }
    return true;
}

// Raw streaming mode: rgb24 rawvideo frames in (what `ffmpeg -f rawvideo -pix_fmt rgb24 pipe:1`
// emits), upscaled frames out (rgb24 or yuv420p10le, what `ffmpeg -f rawvideo ... -s WxH -i pipe:0 -c:v libx265 ...`
// ingests) -- replacing the PNG export / image2 input of reference reve-shared/src/lib.rs:93,100-119 and
// reve-cli/src/main.rs:297-300.  Three host
// threads around one context: a reader fills pinned input buffers, this thread submits / waits (a reve_ctx is not
// thread-safe), a writer drains pinned output buffers -- so the read of frame i+k, the kernels of frame i and the write
// of frame i-k overlap, as decode and encode do in the reference's pipeline.
int run_raw(const Options& o, const reve_model* model) {
    const bool timing = std::getenv("REVE_HOST_TIMING") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto stamp = [&](const char* what) {
        if (timing) std::fprintf(stderr, "[timing] %s at %.1f ms\n", what,
                                 std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
    };
    const int fd_in = (o.in == "-") ? 0 : ::open(o.in.c_str(), O_RDONLY);
    const int fd_out = (o.out == "-") ? 1 : ::open(o.out.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd_in < 0 || fd_out < 0) { std::fprintf(stderr, "error: cannot open %s\n", fd_in < 0 ? o.in.c_str() : o.out.c_str()); return 1; }
#ifdef F_SETPIPE_SZ
    ::fcntl(fd_in, F_SETPIPE_SZ, 1 << 20);    // no-ops (errors ignored) when the descriptor is not a pipe
    ::fcntl(fd_out, F_SETPIPE_SZ, 1 << 20);
#endif
    const int w = o.raw_w, h = o.raw_h, ring = 8, depth = 12;   // 8 frames inside the library, 4 more being read / written
    reve_ctx* ctx = nullptr;
    if (reve_ctx_create(o.gpus.empty() ? 0 : o.gpus[0], model, w, h, o.tile, o.prepad, ring, &ctx) != REVE_OK) {
        std::fprintf(stderr, "error: %s\n", reve_last_error(nullptr));
        return 1;
    }
    size_t out_stride = 0, out_bytes = 0;
    if (reve_ctx_set_output_format(ctx, o.out_format) != REVE_OK || reve_ctx_output_layout(ctx, &out_stride, &out_bytes) != REVE_OK) {
        std::fprintf(stderr, "error: %s\n", reve_last_error(ctx));
        return 1;
    }
    if (o.out_format != REVE_FMT_RGB24 && ((w * o.scale) % 2 || (h * o.scale) % 2)) {
        std::fprintf(stderr, "error: yuv420p10le needs an even output size\n");   // rawvideo has no row padding
        return 1;
    }
    const size_t in_bytes = size_t(w) * h * 3;
    std::vector<uint8_t*> hin(depth, nullptr), hout(depth, nullptr);
    for (int i = 0; i < depth; ++i)
        if (reve_host_alloc(in_bytes, reinterpret_cast<void**>(&hin[i])) != REVE_OK ||
            reve_host_alloc(out_bytes, reinterpret_cast<void**>(&hout[i])) != REVE_OK) {
            std::fprintf(stderr, "error: %s\n", reve_last_error(nullptr));
            return 1;
        }
    stamp("context and pinned buffers ready");
    SlotQueue free_q, ready_q, done_q;
    for (int i = 0; i < depth; ++i) free_q.push(i);
    std::atomic<bool> failed{false};
    std::vector<uint64_t> frame_no(depth, 0);

    std::thread reader([&] {
        for (uint64_t n = 0;; ++n) {
            const int slot = free_q.pop();
            if (slot < 0) break;                  // the other side failed
            size_t got = 0;
            if (!read_full(fd_in, hin[slot], in_bytes, got) || (got != 0 && got != in_bytes)) {
                std::fprintf(stderr, "error: truncated frame %llu\n", (unsigned long long)n);
                failed = true;
                break;
            }
            if (got == 0) break;                  // end of stream
            frame_no[slot] = n;
            ready_q.push(slot);
        }
        ready_q.push(-1);
    });
    std::thread writer([&] {
        for (;;) {
            const int slot = done_q.pop();
            if (slot < 0) break;
            if (!failed && !write_full(fd_out, hout[slot], out_bytes)) {
                std::fprintf(stderr, "error: short write\n");
                failed = true;
            }
            if (o.verbose && !failed)
                std::fprintf(stderr, "frame %llu -> frame %llu done\n", (unsigned long long)frame_no[slot], (unsigned long long)frame_no[slot]);
            free_q.push(slot);
        }
    });

    int inflight = 0;
    bool eof = false;
    while (!failed && (!eof || inflight > 0)) {
        int slot = -2;
        if (!eof && inflight < ring) {
            if (inflight == 0) slot = ready_q.pop();            // nothing to wait for on the GPU: block on the reader
            else if (!ready_q.try_pop(slot)) slot = -2;         // input not ready: retire a frame instead
        }
        if (slot == -1) { eof = true; continue; }
        if (slot >= 0) {
            if (reve_submit(ctx, hin[slot], size_t(w) * 3, hout[slot], out_stride, static_cast<uint64_t>(slot)) != REVE_OK) {
                std::fprintf(stderr, "error: %s\n", reve_last_error(ctx));
                failed = true;
                break;
            }
            ++inflight;
            continue;
        }
        uint64_t tag = 0;
        if (reve_wait(ctx, &tag) != REVE_OK) {
            std::fprintf(stderr, "error: %s\n", reve_last_error(ctx));
            failed = true;
            break;
        }
        --inflight;
        done_q.push(static_cast<int>(tag));
    }
    reve_sync(ctx);
    stamp("last frame left the GPU");
    done_q.push(-1);
    writer.join();
    stamp("last frame written");
    if (failed) {
        // the reader may sit in read() on a pipe whose other end stays open: do not wait for it
        std::fflush(stderr);
        std::_Exit(1);
    }
    reader.join();
    for (int i = 0; i < depth; ++i) { reve_host_free(hin[i]); reve_host_free(hout[i]); }
    reve_ctx_destroy(ctx);
    if (fd_in > 2) ::close(fd_in);
    if (fd_out > 2) ::close(fd_out);
    return 0;
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
        else if (a == "--raw") {
            const char* v = (i + 1 < argc) ? argv[++i] : nullptr;
            if (!v || std::sscanf(v, "%dx%d", &o.raw_w, &o.raw_h) != 2 || o.raw_w < 1 || o.raw_h < 1) return usage();
        }
        else if (a == "--pix-fmt") {
            const char* v = (i + 1 < argc) ? argv[++i] : nullptr;
            const std::string f = v ? v : "";
            if (f == "rgb24") o.out_format = REVE_FMT_RGB24;
            else if (f == "yuv420p10le" || f == "yuv420p10le-bt601") o.out_format = REVE_FMT_YUV420P10LE_BT601;
            else if (f == "yuv420p10le-bt709") o.out_format = REVE_FMT_YUV420P10LE_BT709;
            else return usage();
        }
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
    if (o.out_format != REVE_FMT_RGB24 && o.raw_w == 0) { std::fprintf(stderr, "error: --pix-fmt needs --raw (PNG files hold RGB)\n"); return 2; }
    if (o.tile < 0) o.tile = 200;

    std::vector<std::string> names;
    if (o.raw_w > 0) {
        names.push_back("raw");   // stream mode: no directory
    } else if (DIR* d = opendir(o.in.c_str())) {
        while (dirent* e = readdir(d)) if (ends_with(e->d_name, ".png")) names.push_back(e->d_name);
        closedir(d);
    } else { std::fprintf(stderr, "error: cannot open input directory %s\n", o.in.c_str()); return 1; }
    std::sort(names.begin(), names.end());
    if (o.raw_w == 0)
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
    if (o.raw_w > 0) {
        const int rc = run_raw(o, model);
        reve_model_free(model);
        return rc;
    }
    reve_host::Image first;
    std::string err;
    if (!reve_host::png_read(o.in + "/" + names[0], first, err)) { std::fprintf(stderr, "error: %s\n", err.c_str()); return 1; }
    if (o.gpus.empty()) o.gpus.push_back(0);

    std::atomic<bool> failed{false};
    std::vector<std::thread> threads;
    const int host_threads = std::max(2, static_cast<int>(std::thread::hardware_concurrency()) / static_cast<int>(o.gpus.size()) - 1);
    for (size_t g = 0; g < o.gpus.size(); ++g)
        threads.emplace_back(worker, std::cref(o), model, o.gpus[g], std::cref(names), g, o.gpus.size(), first.w, first.h,
                             std::ref(failed), host_threads);
    for (auto& t : threads) t.join();
    reve_model_free(model);
    if (failed) { std::fprintf(stderr, "error: %s\n", g_err.c_str()); return 1; }
    return 0;
}
