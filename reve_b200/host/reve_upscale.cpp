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
//
// --segments <temp dir>: the in-process multi-GPU segment scheduler, i.e. the
// reference's per-segment loop (reve-cli/src/main.rs:172-350) for G GPUs: one worker thread + one reve_ctx per GPU
// pull segments from the queue in `<temp>/video.temp` and run export(k+1) || upscale(k) || encode(k-1) each, see
// run_segments() below.
#include <dirent.h>
#include <sched.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <future>
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
#include "video_state.h"

namespace {

struct Options {
    std::string in, out, model_dir = "models", name = "realesr-animevideov3", fmt = "png";
    int scale = 2, tile = 200, prepad = 10;
    std::vector<int> gpus;
    bool verbose = false;
    uint64_t seed = 1234;  // --random-weights [seed]: seeded random init instead of the weight files (tests, benches)
    bool random_weights = false;
    std::string segments_dir;          // --segments: temp directory holding video.temp, tmp_frames/, out_frames/, video_parts/
    std::string export_cmd, encode_cmd;   // --export-cmd / --encode-cmd: per-segment shell command templates
    bool schedule_only = false;        // --schedule-only: run the scheduler without the GPU stage (CPU tests of the host logic)
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
                 "                    [-g gpu,gpu,...] [-f png] [-v] [--whole-frame] [--random-weights [seed]]\n"
                 "       reve-upscale --segments temp_dir [-g gpu,gpu,...] [--export-cmd TEMPLATE] [--encode-cmd TEMPLATE] [-v]\n"
                 "         (the reference's per-segment loop for G GPUs: segments of temp_dir/video.temp are pulled from a queue,\n"
                 "          temp_dir/tmp_frames/{i} -> temp_dir/out_frames/{i}; templates may use {index} {size} {seek} {fps}\n"
                 "          {input} {in_dir} {out_dir} {part}; a segment leaves video.temp when its encode command exits 0)\n"
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

// One GPU: a context and its pinned staging buffers, kept alive from segment to segment (the reference pays the model
// load and the device set-up once per SEGMENT because it spawns a process per segment, lib.rs:134; here once per run).
struct Lane {
    const Options& o;
    const reve_model* model;
    int device;
    int host_threads;
    reve_ctx* ctx = nullptr;
    int w = 0, h = 0;
    static constexpr int kDepth = 8;
    std::vector<uint8_t*> hin, hout;

    Lane(const Options& o_, const reve_model* m, int dev, int threads) : o(o_), model(m), device(dev), host_threads(threads) {}
    ~Lane() { release(); }
    void release() {
        if (ctx) reve_sync(ctx);
        for (uint8_t* p : hin) reve_host_free(p);
        for (uint8_t* p : hout) reve_host_free(p);
        hin.clear();
        hout.clear();
        reve_ctx_destroy(ctx);
        ctx = nullptr;
    }
    // (re)creates the context when the frame size changes
    bool prepare(int fw, int fh, std::string& err) {
        if (ctx && fw == w && fh == h) return true;
        release();
        w = fw;
        h = fh;
        // frames smaller than the pre-pad: upstream's reflect-101 reads out of bounds there; pad as far as it is defined
        const int prepad = std::min(o.prepad, std::min(w, h) - 1);
        if (reve_ctx_create(device, model, w, h, o.tile, prepad, kDepth, &ctx) != REVE_OK) { err = reve_last_error(nullptr); ctx = nullptr; return false; }
        const size_t in_bytes = size_t(w) * h * 3, out_bytes = in_bytes * o.scale * o.scale;
        hin.assign(kDepth, nullptr);
        hout.assign(kDepth, nullptr);
        for (int i = 0; i < kDepth; ++i)
            if (reve_host_alloc(in_bytes, reinterpret_cast<void**>(&hin[i])) != REVE_OK ||
                reve_host_alloc(out_bytes, reinterpret_cast<void**>(&hout[i])) != REVE_OK) { err = reve_last_error(nullptr); return false; }
        return true;
    }

    // Frames idx = first, first+step, ... of `names` (files of in_dir) -> out_dir.  Decode pool -> pinned ring -> GPU ->
    // encode pool.  Returns false (and sets the global error) on the first failure.
    bool run(const std::string& in_dir, const std::string& out_dir, const std::vector<std::string>& names, size_t first, size_t step,
             std::atomic<bool>& failed) {
        const int depth = kDepth;
        const size_t in_bytes = size_t(w) * h * 3, out_bytes = in_bytes * o.scale * o.scale;
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
                        if (!reve_host::png_read(in_dir + "/" + names[mine[k]], img, err)) { set_error(err); failed = true; }
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
                const std::string src = in_dir + "/" + names[mine[k]];
                if (img.w != w || img.h != h) { set_error(src + ": frame size differs from the first frame of the segment"); failed = true; break; }
                std::memcpy(hin[slot], img.rgb.data(), in_bytes);
                const std::string stem = names[mine[k]].substr(0, names[mine[k]].find_last_of('.'));
                const std::string dst = out_dir + "/" + stem + "." + o.fmt;
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
        if (reve_sync(ctx) != REVE_OK && !failed) { set_error(reve_last_error(ctx)); failed = true; }
        if (std::getenv("REVE_HOST_TIMING"))
            std::fprintf(stderr, "[timing] gpu %d: %zu frames decoded, upscaled and encoded in %.1f ms after start-up\n", device, mine.size(),
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_ready).count());
        return !failed;
    }
};

// directory mode with several GPUs: the frames of ONE segment dealt round-robin to one lane per GPU
void worker(const Options& o, const reve_model* model, int device, const std::vector<std::string>& names, size_t first,
            size_t step, int w, int h, std::atomic<bool>& failed, int host_threads) {
    Lane lane(o, model, device, host_threads);
    std::string err;
    if (!lane.prepare(w, h, err)) {
        set_error(err);
        failed = true;
        return;
    }
    lane.run(o.in, o.out, names, first, step, failed);
}

// Host placement (what the multi-GPU path shares is host cores, host DRAM and PCIe root complexes): the
// calling thread -- and the pinned buffers it allocates next, and the pool threads it starts -- are restricted to the
// CPUs of the NUMA node the GPU hangs off.  A no-op wherever sysfs does not say (single-node hosts report -1).
int bind_thread_to_gpu_node(int device) {
    char bdf[32];
    if (reve_device_pci_bus_id(device, bdf, sizeof bdf) != REVE_OK) return 0;
    auto slurp = [](const std::string& path) {
        std::string out;
        if (FILE* f = std::fopen(path.c_str(), "r")) {
            char buf[4096];
            const size_t n = std::fread(buf, 1, sizeof buf - 1, f);
            std::fclose(f);
            out.assign(buf, n);
        }
        return out;
    };
    const std::string node = slurp(std::string("/sys/bus/pci/devices/") + bdf + "/numa_node");
    if (node.empty() || std::atoi(node.c_str()) < 0) return 0;
    const std::string list = slurp("/sys/devices/system/node/node" + std::to_string(std::atoi(node.c_str())) + "/cpulist");
    cpu_set_t set;
    CPU_ZERO(&set);
    int n = 0;
    for (const char* p = list.c_str(); *p && *p != '\n';) {
        char* e = nullptr;
        const long a = std::strtol(p, &e, 10);
        long b = a;
        if (e == p) break;
        if (*e == '-') b = std::strtol(e + 1, &e, 10);
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(static_cast<int>(c), &set); ++n; }
        p = (*e == ',') ? e + 1 : e;
    }
    if (n == 0 || sched_setaffinity(0, sizeof set, &set) != 0) return 0;
    return n;
}

std::vector<std::string> list_pngs(const std::string& dir, bool& ok) {
    std::vector<std::string> names;
    ok = false;
    if (DIR* d = opendir(dir.c_str())) {
        while (dirent* e = readdir(d)) if (ends_with(e->d_name, ".png")) names.push_back(e->d_name);
        closedir(d);
        ok = true;
    }
    std::sort(names.begin(), names.end());
    return names;
}
void mkdir_p(const std::string& path) {
    for (size_t pos = 1; pos <= path.size(); ++pos)
        if (pos == path.size() || path[pos] == '/') mkdir(path.substr(0, pos).c_str(), 0777);
}
void remove_dir(const std::string& dir) {   // one level: frame files only (what main.rs:276-285 removes)
    if (DIR* d = opendir(dir.c_str())) {
        while (dirent* e = readdir(d)) {
            const std::string n = e->d_name;
            if (n != "." && n != "..") ::unlink((dir + "/" + n).c_str());
        }
        closedir(d);
    }
    ::rmdir(dir.c_str());
}

// ------------------------------------------------------------------------------------------------------------------
// The segment scheduler: reference reve-cli/src/main.rs:172-350 for G GPUs.
//
// The reference runs export(k+1) || upscale(k) || encode(k-1) with exactly one upscale in flight, pops the head of
// `video.segments` when segment k is handed to the encoder (main.rs:340-343) and repairs the possibly-cut encode on
// resume (main.rs:142-159).  Here every GPU runs that three-stage pipeline on the segments it pulls from ONE shared
// queue (dynamic: a faster GPU simply pulls more), `video.segments` is the SET of segments whose part file is not
// complete, and four reference bugs are fixed on the way:
//   * exit codes are checked: a segment leaves the set only after its encode command exited 0 (reference: never
//     inspected, lib.rs:120-126,148-154,163-170), and any failure ends the run with "error: ..." and exit code 1 while
//     video.temp still lists everything unfinished;
//   * the export is sized by the segment itself ({size}), not by its position in the queue (lib.rs:99,117 export
//     `segments[0 or 1].size`, which is the LAST segment's size whenever two segments are left);
//   * the resume file is replaced atomically (rename), never rewritten in place;
//   * separators are '/', so the layout works outside Windows (lib.rs:90,93,130-131 hard-code '\\').
// Layout, unchanged: <temp>/video.temp, <temp>/tmp_frames/{i}/frame%08d.png -> <temp>/out_frames/{i}/frame%08d.png,
// part files <temp>/video_parts/{i}.mp4 (named by the encode template).  Templates are run with /bin/sh -c after
// substituting {index} {size} {seek} {fps} {input} {in_dir} {out_dir} {part}; without --export-cmd the frames are expected
// to be in tmp_frames/{i} already, without --encode-cmd a segment is complete when its frames are written (and
// out_frames/{i} is kept).  The reference's own commands are spelled out in INTEGRATION.md.
std::string substitute(std::string t, const std::map<std::string, std::string>& vars) {
    for (const auto& kv : vars) {
        const std::string key = "{" + kv.first + "}";
        for (size_t pos = 0; (pos = t.find(key, pos)) != std::string::npos; pos += kv.second.size()) t.replace(pos, key.size(), kv.second);
    }
    return t;
}

int run_segments(const Options& o, const reve_model* model) {
    const std::string temp = o.segments_dir, state_file = temp + "/video.temp";
    reve_host::VideoState video;
    std::string err;
    if (!reve_host::VideoState::load(state_file, video, err)) { std::fprintf(stderr, "error: %s\n", err.c_str()); return 1; }
    if (!o.schedule_only && video.upscale_ratio != o.scale) {
        std::fprintf(stderr, "error: video.temp says upscale_ratio %ld but the model is x%d\n", video.upscale_ratio, o.scale);
        return 1;
    }
    mkdir_p(temp + "/tmp_frames");
    mkdir_p(temp + "/out_frames");
    mkdir_p(temp + "/video_parts");
    // resume: whatever is still listed is redone from scratch; its half-written part file and frame directories go
    // (the reference re-inserts segment first-1 and deletes its part, main.rs:142-159; with set semantics a listed
    // segment is by definition not finished, so nothing has to be guessed)
    for (const auto& sg : video.segments) {
        ::unlink((temp + "/video_parts/" + std::to_string(sg.index) + ".mp4").c_str());
        if (!o.export_cmd.empty()) remove_dir(temp + "/tmp_frames/" + std::to_string(sg.index));
        remove_dir(temp + "/out_frames/" + std::to_string(sg.index));
    }

    // A last segment of ZERO frames exists whenever frame_count % segment_size == 1 (reference lib.rs:282-289 subtracts
    // one from the remainder): nothing to export, upscale or encode, and no part file to wait for.
    for (size_t i = 0; i < video.segments.size();) {
        if (video.segments[i].size <= 0) {
            std::fprintf(stderr, "segment %ld holds no frames (frame_count %% segment_size == 1): nothing to do\n", video.segments[i].index);
            video.segments.erase(video.segments.begin() + static_cast<long>(i));
        } else ++i;
    }
    {
        std::string e;
        if (!video.save(state_file, e)) { std::fprintf(stderr, "error: %s\n", e.c_str()); return 1; }
    }
    std::mutex qm;                                   // guards the queue, the state and its file
    std::deque<reve_host::Segment> queue(video.segments.begin(), video.segments.end());
    std::atomic<bool> failed{false};
    auto claim = [&](reve_host::Segment& sg) {
        std::lock_guard<std::mutex> l(qm);
        if (failed || queue.empty()) return false;
        sg = queue.front();
        queue.pop_front();
        return true;
    };
    auto complete = [&](const reve_host::Segment& sg, int gpu) {
        std::lock_guard<std::mutex> l(qm);
        video.mark_done(sg.index);
        std::string e;
        if (!video.save(state_file, e)) { set_error(e); failed = true; return; }
        std::fprintf(stderr, "segment %ld (%ld frames) done on gpu %d, %zu left\n", sg.index, sg.size, gpu, video.segments.size());
    };
    auto vars_of = [&](const reve_host::Segment& sg) {
        char seek[64], fps[64];
        std::snprintf(seek, sizeof seek, "%.6f", video.seek_seconds(sg.index));
        std::snprintf(fps, sizeof fps, "%.9g/1", video.frame_rate);     // main.rs:302: format!("{}/1", frame_rate)
        const std::string i = std::to_string(sg.index);
        return std::map<std::string, std::string>{{"index", i}, {"size", std::to_string(sg.size)}, {"seek", seek}, {"fps", fps},
                                                  {"input", video.path}, {"in_dir", temp + "/tmp_frames/" + i},
                                                  {"out_dir", temp + "/out_frames/" + i}, {"part", temp + "/video_parts/" + i + ".mp4"}};
    };
    auto shell = [&](const std::string& tmpl, const reve_host::Segment& sg, const char* what) {
        if (tmpl.empty()) return true;
        const std::string cmd = substitute(tmpl, vars_of(sg));
        const int rc = std::system(cmd.c_str());
        if (rc != 0) {
            set_error(std::string(what) + " of segment " + std::to_string(sg.index) + " failed with status " + std::to_string(rc) + ": " + cmd);
            failed = true;
            return false;
        }
        return true;
    };

    std::vector<int> gpus = o.gpus.empty() ? std::vector<int>{0} : o.gpus;
    const int host_threads = std::max(2, static_cast<int>(std::thread::hardware_concurrency()) / static_cast<int>(gpus.size()) - 1);
    std::atomic<long> frames_done{0};
    const auto t_start = std::chrono::steady_clock::now();
    auto gpu_worker = [&](int gpu) {
        if (!o.schedule_only) bind_thread_to_gpu_node(gpu);
        Lane lane(o, model, gpu, host_threads);
        auto do_export = [&](reve_host::Segment sg) {
            const std::string in_dir = temp + "/tmp_frames/" + std::to_string(sg.index);
            if (!o.export_cmd.empty()) mkdir_p(in_dir);
            return shell(o.export_cmd, sg, "export");
        };
        reve_host::Segment cur, nxt, prev;
        bool have_prev = false;
        std::future<bool> enc;
        auto finish_prev = [&]() {      // join encode(k-1); only then is segment k-1 complete (main.rs:280-285)
            if (!have_prev) return;
            have_prev = false;
            if (enc.valid() && !enc.get()) return;
            if (failed) return;
            if (!o.encode_cmd.empty()) remove_dir(temp + "/out_frames/" + std::to_string(prev.index));
            complete(prev, gpu);
        };
        if (!claim(cur)) return;
        if (!do_export(cur)) return;                                   // main.rs:192-216: the first export is synchronous
        for (bool have = true; have && !failed;) {
            const bool have_next = claim(nxt);
            std::future<bool> exp;
            if (have_next) exp = std::async(std::launch::async, do_export, nxt);   // export(k+1), main.rs:223-246
            const std::string in_dir = temp + "/tmp_frames/" + std::to_string(cur.index);
            const std::string out_dir = temp + "/out_frames/" + std::to_string(cur.index);
            mkdir_p(out_dir);                                          // lib.rs:130-132
            bool ok = true;
            if (!o.schedule_only) {                                     // upscale(k), main.rs:262-273
                bool listed = false;
                const std::vector<std::string> names = list_pngs(in_dir, listed);
                if (!listed) { set_error("cannot open " + in_dir); failed = true; ok = false; }
                if (ok && static_cast<long>(names.size()) != cur.size)
                    std::fprintf(stderr, "warning: segment %ld holds %zu frames, video.temp says %ld\n", cur.index, names.size(), cur.size);
                if (ok && !names.empty()) {
                    reve_host::Image first;
                    std::string e;
                    if (!reve_host::png_read(in_dir + "/" + names[0], first, e) || !lane.prepare(first.w, first.h, e)) { set_error(e); failed = true; ok = false; }
                    if (ok) ok = lane.run(in_dir, out_dir, names, 0, 1, failed);
                    if (ok) frames_done += static_cast<long>(names.size());
                }
            }
            if (ok && !o.export_cmd.empty()) remove_dir(in_dir);       // main.rs:276-278
            finish_prev();
            if (ok && !failed) {
                prev = cur;
                have_prev = true;
                const reve_host::Segment sg = cur;
                enc = std::async(std::launch::async, [&, sg] { return shell(o.encode_cmd, sg, "encode"); });   // main.rs:297-339
            }
            if (have_next && !exp.get()) ok = false;
            have = have_next && ok;
            cur = nxt;
        }
        finish_prev();
    };
    std::vector<std::thread> threads;
    for (int g : gpus) threads.emplace_back(gpu_worker, g);
    for (auto& t : threads) t.join();
    if (failed) {
        std::fprintf(stderr, "error: %s\n", g_err.c_str());
        return 1;
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    std::fprintf(stderr, "[summary] %ld frames of %ld segments on %zu lanes in %.2f s (%.1f frames/s, start-up included)\n",
                 frames_done.load(), video.segment_count, gpus.size(), secs, secs > 0 ? frames_done.load() / secs : 0.0);
    return 0;
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
    // frames smaller than the pre-pad: upstream's reflect-101 reads out of bounds there; pad as far as it is defined
    if (reve_ctx_create(o.gpus.empty() ? 0 : o.gpus[0], model, w, h, o.tile, std::min(o.prepad, std::min(w, h) - 1), ring, &ctx) != REVE_OK) {
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
        else if (a == "--whole-frame") o.tile = -1;
        else if (a == "--schedule-only") o.schedule_only = true;
        else if (a == "--random-weights") {
            o.random_weights = true;
            if (i + 1 < argc && std::isdigit(static_cast<unsigned char>(argv[i + 1][0]))) o.seed = std::strtoull(argv[++i], nullptr, 10);
        }
        else if (a == "--segments" || a == "--export-cmd" || a == "--encode-cmd") {
            const char* v = next();
            if (!v) return usage();
            if (a == "--segments") o.segments_dir = v; else if (a == "--export-cmd") o.export_cmd = v; else o.encode_cmd = v;
        }
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
    const bool seg_mode = !o.segments_dir.empty();
    if (!seg_mode && (o.in.empty() || o.out.empty())) return usage();
    if (seg_mode) {   // the scale comes from the resume file (Video.upscale_ratio, lib.rs:24)
        reve_host::VideoState v;
        std::string e;
        if (!reve_host::VideoState::load(o.segments_dir + "/video.temp", v, e)) { std::fprintf(stderr, "error: %s\n", e.c_str()); return 1; }
        o.scale = static_cast<int>(v.upscale_ratio);
    }
    if (o.scale < 2 || o.scale > 4) { std::fprintf(stderr, "error: scale must be 2, 3 or 4\n"); return 2; }
    if (o.fmt != "png") { std::fprintf(stderr, "error: only -f png is supported\n"); return 2; }
    if (o.out_format != REVE_FMT_RGB24 && o.raw_w == 0) { std::fprintf(stderr, "error: --pix-fmt needs --raw (PNG files hold RGB)\n"); return 2; }
    // upstream: -t 0 = "auto", which resolves to 200 on any device with more than 1.9 GB;
    // the library's whole-frame mode (tile = 0 at the ABI) is --whole-frame here
    if (o.tile == 0) o.tile = 200;
    else if (o.tile < 0) o.tile = 0;

    std::vector<std::string> names;
    if (o.raw_w > 0 || seg_mode) {
        names.push_back("raw");   // stream / scheduler mode: no single directory
    } else {
        bool listed = false;
        names = list_pngs(o.in, listed);
        if (!listed) { std::fprintf(stderr, "error: cannot open input directory %s\n", o.in.c_str()); return 1; }
    }
    if (o.raw_w == 0 && !seg_mode) mkdir_p(o.out);   // mkdir -p (the reference creates only the leaf, lib.rs:130-132)
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
    } else if (o.schedule_only && seg_mode) {
        model = nullptr;   // no GPU stage
    } else if (o.random_weights) {
        std::fprintf(stderr, "warning: --random-weights: seeded random init of the architecture (seed %llu), not a trained model\n",
                     static_cast<unsigned long long>(o.seed));
        if (reve_model_random(o.scale, o.seed, &model) != REVE_OK) { std::fprintf(stderr, "error: %s\n", reve_last_error(nullptr)); return 1; }
    } else {
        // as the spawned upstream binary: no model, no output (and, unlike the reference's caller, a non-zero exit code)
        std::fprintf(stderr, "error: %s.param/.bin not found (-m is resolved against the current directory); "
                             "--random-weights runs the architecture with seeded random weights\n", stem.c_str());
        return 1;
    }
    if (seg_mode) {
        const int rc = run_segments(o, model);
        reve_model_free(model);
        return rc;
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
