// The reference's resume file `temp\video.temp` (serde JSON of `Video`, reference reve-shared/src/lib.rs:9-25) read and
// written from C++ with the same field names, so a file written by reve-cli loads here and vice versa
// (SURVEY.md section 8(f) rank 2).  The reference treats `segments` as an in-order queue and removes its head when a
// segment has been HANDED to the encoder (reve-cli/src/main.rs:340-343); with G GPUs segments finish out of order, so
// here `segments` is the SET of segments whose part file is not complete yet, and an entry is removed only after its
// encode command has exited with status 0.  The same class is mirrored in Python (reve_b200/segments.py:VideoState).
#pragma once
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace reve_host {

struct Segment {
    long index = 0;
    long size = 0;
};

struct VideoState {
    std::string path, output_path;
    std::vector<Segment> segments;   // not yet encoded
    double frame_rate = 0;
    long frame_count = 0, segment_size = 0, segment_count = 0, upscale_ratio = 0;

    // reference lib.rs:282-289: the remainder minus one (compensating the one-frame-early seek at lib.rs:97), or a
    // full segment if it divides evenly
    static long last_segment_size(long frame_count, long segment_size) {
        const long last = frame_count % segment_size;
        return last == 0 ? segment_size : last - 1;
    }
    // reference lib.rs:59-73
    static VideoState create(const std::string& path, const std::string& output_path, long frame_count, double frame_rate,
                             long segment_size, long upscale_ratio) {
        VideoState v;
        v.path = path;
        v.output_path = output_path;
        v.frame_rate = frame_rate;
        v.frame_count = frame_count;
        v.segment_size = segment_size;
        v.upscale_ratio = upscale_ratio;
        const long parts = frame_count ? (frame_count + segment_size - 1) / segment_size : 0;
        for (long i = 0; i + 1 < parts; ++i) v.segments.push_back({i, segment_size});
        if (parts) v.segments.push_back({parts - 1, last_segment_size(frame_count, segment_size)});
        v.segment_count = parts;
        return v;
    }
    // reference lib.rs:94-98: `-ss` of the export; one frame early for every segment but the first
    double seek_seconds(long index) const {
        return index == 0 ? 0.0 : (static_cast<double>(index * segment_size - 1) / frame_rate);
    }
    bool mark_done(long index) {
        for (size_t i = 0; i < segments.size(); ++i)
            if (segments[i].index == index) {
                segments.erase(segments.begin() + static_cast<long>(i));
                return true;
            }
        return false;
    }

    // ---- JSON (only what serde_json emits for `Video`: objects, arrays, strings, numbers)
    static std::string quote(const std::string& s) {
        std::string o = "\"";
        for (unsigned char c : s) {
            if (c == '"' || c == '\\') { o += '\\'; o += static_cast<char>(c); }
            else if (c == '\n') o += "\\n";
            else if (c == '\t') o += "\\t";
            else if (c == '\r') o += "\\r";
            else if (c < 0x20) { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", c); o += b; }
            else o += static_cast<char>(c);
        }
        return o + "\"";
    }
    std::string to_json() const {
        std::ostringstream o;
        o << "{\"path\":" << quote(path) << ",\"output_path\":" << quote(output_path) << ",\"segments\":[";
        for (size_t i = 0; i < segments.size(); ++i)
            o << (i ? "," : "") << "{\"index\":" << segments[i].index << ",\"size\":" << segments[i].size << "}";
        char fr[64];
        std::snprintf(fr, sizeof fr, "%.9g", frame_rate);
        std::string frs = fr;
        if (frs.find_first_of(".eE") == std::string::npos) frs += ".0";   // serde writes f32 with a fraction
        o << "],\"frame_rate\":" << frs << ",\"frame_count\":" << frame_count << ",\"segment_size\":" << segment_size
          << ",\"segment_count\":" << segment_count << ",\"upscale_ratio\":" << upscale_ratio << "}";
        return o.str();
    }

    struct Parser {
        const std::string& s;
        size_t i = 0;
        std::string err;
        explicit Parser(const std::string& t) : s(t) {}
        void ws() { while (i < s.size() && std::isspace(static_cast<unsigned char>(s[i]))) ++i; }
        bool eat(char c) { ws(); if (i < s.size() && s[i] == c) { ++i; return true; } return false; }
        bool fail(const std::string& m) { if (err.empty()) err = m + " at offset " + std::to_string(i); return false; }
        bool str(std::string& out) {
            ws();
            if (i >= s.size() || s[i] != '"') return fail("expected a string");
            ++i;
            out.clear();
            while (i < s.size() && s[i] != '"') {
                if (s[i] == '\\' && i + 1 < s.size()) {
                    const char e = s[++i];
                    if (e == 'n') out += '\n'; else if (e == 't') out += '\t'; else if (e == 'r') out += '\r';
                    else if (e == 'u' && i + 4 < s.size()) {
                        const unsigned cp = static_cast<unsigned>(std::strtoul(s.substr(i + 1, 4).c_str(), nullptr, 16));
                        i += 4;
                        if (cp < 0x80) out += static_cast<char>(cp);
                        else if (cp < 0x800) { out += static_cast<char>(0xC0 | (cp >> 6)); out += static_cast<char>(0x80 | (cp & 0x3F)); }
                        else { out += static_cast<char>(0xE0 | (cp >> 12)); out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F)); out += static_cast<char>(0x80 | (cp & 0x3F)); }
                    } else out += e;
                    ++i;
                } else out += s[i++];
            }
            if (i >= s.size()) return fail("unterminated string");
            ++i;
            return true;
        }
        bool num(double& out) {
            ws();
            const char* b = s.c_str() + i;
            char* e = nullptr;
            out = std::strtod(b, &e);
            if (e == b) return fail("expected a number");
            i += static_cast<size_t>(e - b);
            return true;
        }
    };

    static bool from_json(const std::string& text, VideoState& v, std::string& err) {
        Parser p(text);
        v = VideoState();
        int seen = 0;
        if (!p.eat('{')) { err = "video.temp: expected '{'"; return false; }
        do {
            std::string key;
            if (!p.str(key) || !p.eat(':')) { err = "video.temp: " + (p.err.empty() ? std::string("expected ':'") : p.err); return false; }
            double d = 0;
            bool ok = true;
            if (key == "path") ok = p.str(v.path);
            else if (key == "output_path") ok = p.str(v.output_path);
            else if (key == "segments") {
                ok = p.eat('[');
                if (ok && !p.eat(']')) {
                    do {
                        Segment sg;
                        bool hi = false, hs = false;
                        ok = p.eat('{');
                        while (ok) {
                            std::string k2;
                            ok = p.str(k2) && p.eat(':') && p.num(d);
                            if (!ok) break;
                            if (k2 == "index") { sg.index = std::lround(d); hi = true; }
                            if (k2 == "size") { sg.size = std::lround(d); hs = true; }
                            if (!p.eat(',')) break;
                        }
                        ok = ok && p.eat('}') && hi && hs;
                        if (ok) v.segments.push_back(sg);
                    } while (ok && p.eat(','));
                    ok = ok && p.eat(']');
                }
            } else {
                ok = p.num(d);
                if (key == "frame_rate") v.frame_rate = d;
                else if (key == "frame_count") v.frame_count = std::lround(d);
                else if (key == "segment_size") v.segment_size = std::lround(d);
                else if (key == "segment_count") v.segment_count = std::lround(d);
                else if (key == "upscale_ratio") v.upscale_ratio = std::lround(d);
                else ok = false;
            }
            if (!ok) { err = "video.temp: bad value for '" + key + "'" + (p.err.empty() ? "" : " (" + p.err + ")"); return false; }
            ++seen;
        } while (p.eat(','));
        if (!p.eat('}') || seen != 8) { err = "video.temp: not the reference's Video schema (8 fields)"; return false; }
        if (v.upscale_ratio < 2 || v.upscale_ratio > 4 || v.segment_size < 1 || v.frame_rate <= 0) {
            err = "video.temp: values out of range";
            return false;
        }
        return true;
    }

    static bool load(const std::string& file, VideoState& v, std::string& err) {
        std::ifstream f(file);
        if (!f) { err = "cannot open " + file; return false; }
        std::stringstream ss;
        ss << f.rdbuf();
        return from_json(ss.str(), v, err);
    }
    // atomic: a crash never leaves a truncated resume file (the reference rewrites it in place, main.rs:340-343)
    bool save(const std::string& file, std::string& err) const {
        const std::string tmp = file + ".tmp";
        {
            std::ofstream f(tmp, std::ios::trunc);
            if (!f) { err = "cannot write " + tmp; return false; }
            f << to_json();
            f.flush();
            if (!f) { err = "write failed: " + tmp; return false; }
        }
        if (std::rename(tmp.c_str(), file.c_str()) != 0) { err = "cannot rename " + tmp; return false; }
        return true;
    }
};

}  // namespace reve_host
