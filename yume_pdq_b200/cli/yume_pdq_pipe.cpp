// yume_pdq_pipe.cpp -- `yume-pdq-cuda pipe`: the reference CLI's `pipe` subcommand (src/main.rs:139-253, 448-823,
// 1090-1300) on top of the batched CUDA hasher.  Same arguments, input formats, output wire formats, separators
// and exit codes; the per-frame 2-thread ping-pong of the reference (PairProcessor::loop_thread) becomes a
// batching reader: frames are read straight into page-locked buffers, a batch goes to the GPU in one
// yume_pdq_cuda_hash_batch_pixels call (input conversion fused into the hash kernel), and the records are
// formatted and written while the reader fills the next buffer.  A batch closes when `--batch` frames are in, or
// earlier when the input has no more bytes ready -- a slow producer (ffmpeg at camera rate) still gets every
// hash as soon as its frame is complete, a fast one gets GPU-sized batches.
//
// Links only against the C ABI (include/yume_pdq_cuda.h); there is no CPU fallback.
//
// Hidden developer subcommand `format-records` (CPU only): reads binary records
//   {u8 hash[32]; f32 quality; f32 threshold; f32 star_threshold; u32 dihedral; u32 index; f32 pdqf[256]}
// from the input and writes them in the requested wire format -- the hook tests/test_wire_formats.py uses to pin
// the formats without a GPU.
#include <fcntl.h>
#include <poll.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/yume_pdq_cuda.h"
#include "wire.hpp"

namespace {

constexpr int EXIT_IO = 5, EXIT_FATAL = 128, EXIT_USAGE = 255;   // main.rs:1036-1052, 1261, 1293

struct Args {
    std::string input = "-", output = "-", input_format = "luma8", format = "bin";
    bool dihedrals = false, stats = false, fma = false;
    int device = 0;
    size_t batch = 256;
};

const char *HELP =
    "yume-pdq-cuda pipe [-i|--input PATH|-] [-o|--output PATH|-] [--input-format luma8|rgb8|rgba8]\n"
    "                   [-f|--format FORMAT] [--dihedrals] [--stats] [--device N] [--batch N]\n\n"
    "Reads a stream of raw 512x512 frames and writes their PDQ hashes (same wire formats as `yume-pdq pipe`).\n"
    "Formats: raw/RAW, hex/HEX, bin/BIN, q+raw, q+hex, q+HEX, q+bin, diagnostic;\n"
    "modifiers: a trailing +lf / +cr / +crlf / +nul / +nl adds a separator.  Examples: q+hex+lf, HEX+crlf, q+bin\n";

bool write_all(int fd, const char *p, size_t n)
{
    while (n) {
        const ssize_t w = write(fd, p, n);
        if (w < 0) {
            if (errno == EINTR)
                continue;
            return false;
        }
        p += w;
        n -= (size_t)w;
    }
    return true;
}

int open_in(const std::string &spec) { return spec == "-" ? dup(0) : open(spec.c_str(), O_RDONLY); }
int open_out(const std::string &spec)
{
    return spec == "-" ? dup(1) : open(spec.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
}

struct Batch {
    uint8_t *frames = nullptr;   // page-locked, batch x frame_bytes
    size_t n = 0;
    bool last = false;
    int error = 0;               // EXIT_IO when the input ended in the middle of a row / failed
};

// three buffers rotate reader -> hasher/writer
struct Ring {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<Batch> bufs;
    std::vector<int> filled, free_;
    int pop(std::vector<int> &q)
    {
        std::unique_lock<std::mutex> l(mu);
        cv.wait(l, [&] { return !q.empty(); });
        const int i = q.front();
        q.erase(q.begin());
        return i;
    }
    void push(std::vector<int> &q, int i)
    {
        {
            std::lock_guard<std::mutex> l(mu);
            q.push_back(i);
        }
        cv.notify_all();
    }
};

// Fills b with up to `max_frames` frames.  Blocks for the first frame; afterwards keeps reading only while the
// descriptor has bytes ready.  End of input at a row boundary ends the stream silently -- also in the middle of a
// frame, like the reference, whose reader checks for EOF per row (main.rs:494-506); inside a row it is an error.
void read_batch(int fd, size_t frame_bytes, size_t row_bytes, size_t max_frames, Batch &b)
{
    b.n = 0;
    b.last = false;
    b.error = 0;
    size_t have = 0;    // bytes of the frame being read
    for (;;) {
        if (b.n == max_frames)
            return;
        if (b.n > 0 && have == 0) {
            struct pollfd p = {fd, POLLIN, 0};
            if (poll(&p, 1, 0) == 0)
                return;           // nothing ready: hand over what we have
        }
        const ssize_t r = read(fd, b.frames + b.n * frame_bytes + have, frame_bytes - have);
        if (r < 0) {
            if (errno == EINTR)
                continue;
            fprintf(stderr, "IO Error in reader: %s\n", strerror(errno));
            b.error = EXIT_IO;
            b.last = true;
            return;
        }
        if (r == 0) {
            if (have % row_bytes != 0) {
                fprintf(stderr, "IO Error in reader: Unexpected EOF while reading the middle of a frame\n");
                b.error = EXIT_IO;
            }
            b.last = true;
            return;
        }
        have += (size_t)r;
        if (have == frame_bytes) {
            b.n++;
            have = 0;
        }
    }
}

int run_pipe(const Args &a)
{
    ypdq_wire::Format fmt;
    if (!ypdq_wire::parse_format(a.format, &fmt)) {
        fprintf(stderr,
                "invalid output format: '%s',\n\nplease try one of the following:\n"
                "'raw', 'RAW': raw binary output\n'q+raw', 'q+RAW': raw binary output prefixed by the f32 quality\n"
                "'hex', 'HEX': hex output (case-sensitive)\n"
                "'q+hex', 'q+HEX': hex output with ASCII decimal quality score prefix separated by a colon\n"
                "'bin', 'BIN': binary output\n"
                "'q+bin', 'q+BIN': binary output with ASCII decimal quality score prefix separated by a colon\n",
                a.format.c_str());
        return EXIT_USAGE;
    }
    int pixel_format;
    if (a.input_format == "luma8" || a.input_format == "LUMA8")
        pixel_format = YUME_PDQ_CUDA_PIXEL_LUMA8;
    else if (a.input_format == "rgb8" || a.input_format == "RGB8")
        pixel_format = YUME_PDQ_CUDA_PIXEL_RGB8;
    else if (a.input_format == "rgba8" || a.input_format == "RGBA8")
        pixel_format = YUME_PDQ_CUDA_PIXEL_RGBA8;
    else {
        fprintf(stderr, "invalid input format: '%s',\n\nplease try one of the following:\nluma8, rgb8, rgba8\n",
                a.input_format.c_str());
        return EXIT_USAGE;
    }
    const size_t frame_bytes = yume_pdq_cuda_frame_bytes(pixel_format), row_bytes = frame_bytes / 512;
    if (a.fma && pixel_format != YUME_PDQ_CUDA_PIXEL_LUMA8)
        pixel_format |= YUME_PDQ_CUDA_PIXEL_FMA;
    const int in_fd = open_in(a.input), out_fd = open_out(a.output);
    if (in_fd < 0 || out_fd < 0) {
        fprintf(stderr, "cannot open %s: %s\n", in_fd < 0 ? a.input.c_str() : a.output.c_str(), strerror(errno));
        return EXIT_FATAL;
    }
    if (yume_pdq_cuda_device_count() <= a.device) {
        fprintf(stderr, "no usable CUDA device %d (%s); this build has no CPU fallback\n", a.device,
                yume_pdq_cuda_error_string(YUME_PDQ_CUDA_E_NO_DEVICE));
        return EXIT_FATAL;
    }
    Ring ring;
    ring.bufs.resize(3);
    for (int i = 0; i < 3; i++) {
        void *p = nullptr;
        const int rc = yume_pdq_cuda_host_alloc(a.device, a.batch * frame_bytes, &p);
        if (rc != 0) {
            fprintf(stderr, "pinned allocation failed: %s\n", yume_pdq_cuda_error_string(rc));
            return EXIT_FATAL;
        }
        ring.bufs[i].frames = static_cast<uint8_t *>(p);
        ring.free_.push_back(i);
    }
    std::atomic<uint64_t> frames_done{0};
    std::atomic<bool> finished{false};

    std::thread reader([&] {
        for (;;) {
            const int i = ring.pop(ring.free_);
            read_batch(in_fd, frame_bytes, row_bytes, a.batch, ring.bufs[i]);
            const bool last = ring.bufs[i].last;
            ring.push(ring.filled, i);
            if (last)
                return;
        }
    });

    std::thread stats;
    if (a.stats)
        stats = std::thread([&] {   // main.rs:1172-1196
            auto t_last = std::chrono::steady_clock::now();
            const auto t0 = t_last;
            uint64_t last = 0;
            while (!finished.load()) {
                std::this_thread::sleep_for(std::chrono::milliseconds(100));
                const auto now = std::chrono::steady_clock::now();
                const auto dt = std::chrono::duration_cast<std::chrono::microseconds>(now - t_last).count();
                if (dt <= 1000000)
                    continue;
                const uint64_t total = frames_done.load(), delta = total - last;
                const auto el = std::chrono::duration_cast<std::chrono::microseconds>(now - t0).count();
                fprintf(stderr, "%llu new frames processed (%llu fps), %llu total frames processed (%llu fps overall)\n",
                        (unsigned long long)delta, (unsigned long long)(1000000ull * delta / (uint64_t)dt),
                        (unsigned long long)total, (unsigned long long)(1000000ull * total / (uint64_t)el));
                last = total;
                t_last = now;
            }
        });

    const bool diag = (fmt.type & ~ypdq_wire::FLAG_UPPER) == ypdq_wire::DIAGNOSTIC;
    std::vector<uint8_t> hashes(a.batch * 8 * 32);
    std::vector<float> quality(a.batch * 8), thresholds(a.batch * 8), pdqf(diag ? a.batch * 256 : 0);
    std::string out;
    int exit_code = 0;
    for (;;) {
        const int i = ring.pop(ring.filled);
        Batch &b = ring.bufs[i];
        if (b.n > 0) {
            const int rc = yume_pdq_cuda_hash_batch_pixels(a.device, b.frames, pixel_format, b.n, hashes.data(), quality.data(),
                                                           thresholds.data(), diag ? pdqf.data() : nullptr, nullptr);
            if (rc != 0) {
                fprintf(stderr, "Fatal Error in hasher: %s\n", yume_pdq_cuda_error_string(rc));
                exit_code = EXIT_FATAL;
            } else {
                out.clear();
                float state[256];
                for (size_t k = 0; k < b.n; k++) {
                    const int n_rec = a.dihedrals ? 8 : 1;
                    for (int d = 0; d < n_rec; d++) {
                        ypdq_wire::Record r;
                        r.hash = hashes.data() + (k * 8 + d) * 32;
                        r.quality = quality[k * 8 + d];
                        r.threshold = thresholds[k * 8 + d];
                        r.star_threshold = thresholds[k * 8];
                        r.dihedral = d == 0 ? 0u : YUME_PDQ_CUDA_DIHEDRAL_ORDER[d];
                        r.index = (unsigned)d;
                        r.pdqf = nullptr;
                        if (diag) {
                            if (d == 0)
                                r.pdqf = pdqf.data() + k * 256;
                            else {
                                ypdq_wire::dihedral_pdqf(pdqf.data() + k * 256, d, state);
                                r.pdqf = state;
                            }
                        }
                        ypdq_wire::append_record(out, fmt, r);
                    }
                }
                if (!write_all(out_fd, out.data(), out.size())) {
                    fprintf(stderr, "IO Error in writer: %s\n", strerror(errno));
                    exit_code = EXIT_IO;
                }
                frames_done.fetch_add(b.n);
            }
        }
        const bool last = b.last;
        if (b.error && !exit_code)
            exit_code = b.error;
        if (last || exit_code)
            break;
        ring.push(ring.free_, i);
    }
    finished.store(true);
    if (exit_code && !ring.bufs.empty()) {
        // the reader may be blocked in read(); closing its descriptor and detaching is the simplest way out
        close(in_fd);
        reader.detach();
    } else {
        reader.join();
    }
    if (stats.joinable())
        stats.join();
    close(out_fd);
    if (exit_code)
        _exit(exit_code);    // do not wait for a reader stuck in read()
    for (auto &b : ring.bufs)
        yume_pdq_cuda_host_free(b.frames);
    return 0;
}

#pragma pack(push, 1)
struct RawRecord {
    uint8_t hash[32];
    float quality, threshold, star_threshold;
    uint32_t dihedral, index;
    float pdqf[256];
};
#pragma pack(pop)

int run_format_records(const Args &a)
{
    ypdq_wire::Format fmt;
    if (!ypdq_wire::parse_format(a.format, &fmt)) {
        fprintf(stderr, "invalid output format: '%s'\n", a.format.c_str());
        return EXIT_USAGE;
    }
    const int in_fd = open_in(a.input), out_fd = open_out(a.output);
    if (in_fd < 0 || out_fd < 0)
        return EXIT_FATAL;
    std::string out;
    RawRecord rr;
    for (;;) {
        size_t have = 0;
        while (have < sizeof rr) {
            const ssize_t r = read(in_fd, reinterpret_cast<char *>(&rr) + have, sizeof rr - have);
            if (r <= 0)
                break;
            have += (size_t)r;
        }
        if (have == 0)
            break;
        if (have != sizeof rr)
            return EXIT_IO;
        ypdq_wire::Record r{rr.hash, rr.quality, rr.threshold, rr.pdqf, rr.star_threshold, rr.dihedral, rr.index};
        ypdq_wire::append_record(out, fmt, r);
    }
    return write_all(out_fd, out.data(), out.size()) ? 0 : EXIT_IO;
}

}  // namespace

int main(int argc, char **argv)
{
    if (argc < 2 || !strcmp(argv[1], "--help") || !strcmp(argv[1], "-h") || !strcmp(argv[1], "help")) {
        fputs(HELP, argc < 2 ? stderr : stdout);
        return argc < 2 ? EXIT_USAGE : 0;
    }
    if (!strcmp(argv[1], "--version") || !strcmp(argv[1], "-V")) {
        printf("yume-pdq-cuda %s\n", YUME_PDQ_VERSION_STRING);
        return 0;
    }
    const std::string sub = argv[1];
    Args a;
    for (int i = 2; i < argc; i++) {
        const std::string k = argv[i];
        auto value = [&](const char *name) -> const char * {
            if (i + 1 >= argc) {
                fprintf(stderr, "error: a value is required for '%s'\n", name);
                exit(2);
            }
            return argv[++i];
        };
        if (k == "-i" || k == "--input")
            a.input = value("--input");
        else if (k == "-o" || k == "--output")
            a.output = value("--output");
        else if (k == "--input-format")
            a.input_format = value("--input-format");
        else if (k == "-f" || k == "--format")
            a.format = value("--format");
        else if (k == "--dihedrals")
            a.dihedrals = true;
        else if (k == "--stats")
            a.stats = true;
        else if (k == "--fma-luma")
            a.fma = true;     // the rounding of the AVX2 cvt_rgb8_to_luma8f override (x86.rs:686-753)
        else if (k == "--output-buffer" || k == "--core0" || k == "--core1")
            (void)value(k.c_str());    // accepted and ignored, like the reference without the hpc feature
        else if (k == "--device")
            a.device = atoi(value("--device"));
        else if (k == "--batch")
            a.batch = (size_t)std::max(1, atoi(value("--batch")));
        else if (k == "--help" || k == "-h") {
            fputs(HELP, stdout);
            return 0;
        } else {
            fprintf(stderr, "error: unexpected argument '%s' found\n", k.c_str());
            return 2;    // clap's usage error code
        }
    }
    if (sub == "pipe")
        return run_pipe(a);
    if (sub == "format-records")
        return run_format_records(a);
    fprintf(stderr, "Invalid subcommand, try --help for usage\n");   // main.rs:1503
    return EXIT_USAGE;
}
