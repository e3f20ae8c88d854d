// cvsteer-run -- batch driver with the reference CLI's contract (reference example/steer.cpp:59-173), on libcvsteer_b200.
//
//   cvsteer-run --input=<image.png | image.pgm | list.txt> --output=<dir> [--gain=<g>] [--format=png|pgm] [--verbose] [--help]
//
// Per input file, as ParallelSteerable::operator() does (example/steer.cpp:69-124): gray 8-bit image ->
// SteerableFiltersG2(gray, 4, 0.67f) -> steer(dominant angle, ...) -> findEdges / findDarkLines / findBrightLines(magnitude,
// phase) -> 8-bit (gain > 0: convertTo(CV_8UC1, gain); else normalize(0, 255, NORM_MINMAX)) -> three files
// <dir>/<basename>_{edges,lines_dark,lines_bright}.png.  All of it is ONE call, cvs_g2_lines_u8_host, per file; files are
// spread over worker threads (one handle, hence one CUDA stream, each) the way the reference spreads them over
// cv::parallel_for_ (example/steer.cpp:169), and over all visible GPUs round-robin.
//
// Image files: PNG in (any colour type / bit depth / interlacing, converted to 8-bit gray the way cv::imread + cvtColor(BGR2GRAY)
// do, png_io.h) and binary PGM in (P5, maxval 255); PNG out by default, as the reference writes (example/steer.cpp:106-121),
// or PGM with --format=pgm.  JPEG and the other formats cv::imread knows are not read (no codec with C headers in this build
// image): such files are skipped silently, like every unreadable file is (`if (image.empty()) continue;`, steer.cpp:74-77).
//
// Difference from the reference, on purpose: --gain is honoured.  The reference declares it but passes the `verbose` flag as
// the gain (example/steer.cpp:167-168), so its effective gain is 0 or 1; pass --gain=1 to reproduce `--verbose` runs.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "cvsteer_c.h"
#include "png_io.h"

namespace {

struct Args {
    std::string input, output, format = "png";
    float gain = 0.f;
    bool verbose = false, help = false;
};

bool parse(int argc, const char** argv, Args& a)
{
    for (int i = 1; i < argc; ++i) {
        std::string s = argv[i];
        auto val = [&](const char* key, std::string& dst) {
            const std::string k = std::string("--") + key;
            if (s.rfind(k + "=", 0) == 0) return dst = s.substr(k.size() + 1), true;
            if (s == k && i + 1 < argc) return dst = argv[++i], true;
            return false;
        };
        std::string g;
        if (val("input", a.input) || val("output", a.output) || val("format", a.format)) continue;
        if (val("gain", g)) {
            a.gain = (float)atof(g.c_str());
            continue;
        }
        if (s == "--verbose" || s == "--verbose=true") a.verbose = true;
        else if (s == "--help" || s == "--help=true") a.help = true;
        else if (s.rfind("--", 0) != 0 && a.input.empty()) a.input = s;
    }
    return true;
}

void usage()
{
    printf("Usage: cvsteer-run [params]\n\n"
           "\t--gain (value:0.0)\n\t\tgain for CV_8UC1 output\n"
           "\t--help (value:false)\n\t\thelp message\n"
           "\t--format (value:png)\n\t\toutput files: png (as the reference) or pgm\n"
           "\t--input\n\t\tinput image (PNG or binary PGM) or .txt list of images\n"
           "\t--output\n\t\toutput directory\n"
           "\t--verbose (value:false)\n\t\tuse verbose display\n");
}

// BASH-style basename without the extension (example/steer.cpp:51-57)
std::string stem(const std::string& name)
{
    const size_t pos = name.rfind('/');
    std::string base = pos == std::string::npos ? name : name.substr(pos + 1);
    const size_t dot = base.rfind('.');
    return dot == std::string::npos ? base : base.substr(0, dot);
}

bool read_file(const std::string& path, std::vector<unsigned char>& bytes)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    bytes.clear();
    unsigned char buf[1 << 16];
    for (size_t n; (n = fread(buf, 1, sizeof(buf), f)) > 0;) bytes.insert(bytes.end(), buf, buf + n);
    fclose(f);
    return !bytes.empty();
}

bool decode_pgm(const std::vector<unsigned char>& file, std::vector<unsigned char>& px, int& rows, int& cols)
{
    size_t pos = 2;
    auto token = [&](int& v) {
        for (;;) {
            while (pos < file.size() && (file[pos] == ' ' || file[pos] == '\t' || file[pos] == '\n' || file[pos] == '\r')) ++pos;
            if (pos >= file.size() || file[pos] != '#') break;
            while (pos < file.size() && file[pos] != '\n') ++pos;
        }
        if (pos >= file.size() || file[pos] < '0' || file[pos] > '9') return false;
        v = 0;
        while (pos < file.size() && file[pos] >= '0' && file[pos] <= '9') {
            if (v > (1 << 26)) return false;  // (a damaged header: no int overflow, no absurd allocation)
            v = v * 10 + (file[pos++] - '0');
        }
        ++pos;  // exactly one whitespace byte follows the number
        return true;
    };
    int maxv = 0;
    if (file.size() < 2 || file[0] != 'P' || file[1] != '5' || !token(cols) || !token(rows) || !token(maxv) || maxv != 255 || rows <= 0 || cols <= 0)
        return false;
    const size_t n = (size_t)rows * cols;
    if (pos > file.size() || file.size() - pos < n) return false;
    px.assign(file.begin() + pos, file.begin() + pos + n);
    return true;
}

// 8-bit gray pixels of an image file: PNG or binary PGM, told apart by their magic bytes
bool read_gray(const std::string& path, std::vector<unsigned char>& px, int& rows, int& cols)
{
    std::vector<unsigned char> file;
    if (!read_file(path, file)) return false;
    if (pngio::is_png(file.data(), file.size())) return pngio::decode_gray(file.data(), file.size(), px, rows, cols);
    return decode_pgm(file, px, rows, cols);
}

bool write_gray(const std::string& base, bool png, const unsigned char* px, int rows, int cols)
{
    std::vector<unsigned char> out;
    if (png) {
        if (!pngio::encode_gray(px, rows, cols, out)) return false;
    } else {
        char hdr[64];
        const int n = snprintf(hdr, sizeof(hdr), "P5\n%d %d\n255\n", cols, rows);
        out.assign(hdr, hdr + n);
        out.insert(out.end(), px, px + (size_t)rows * cols);
    }
    FILE* f = fopen((base + (png ? ".png" : ".pgm")).c_str(), "wb");
    if (!f) return false;
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

}  // namespace

int main(int argc, const char* argv[])
{
    Args a;
    parse(argc, argv, a);
    if (argc < 2 || a.help) {
        usage();
        return 0;
    }
    std::vector<std::string> files;
    if (a.input.rfind(".txt") != std::string::npos || a.input.rfind('.') == std::string::npos) {  // example/steer.cpp:156
        std::ifstream in(a.input.c_str());
        for (std::string line; std::getline(in, line);)
            if (!line.empty()) files.push_back(line);
    } else {
        files.push_back(a.input);
    }
    int ndev = 0;
    if (cvs_device_count(&ndev) != CVS_OK || ndev < 1) {
        fprintf(stderr, "cvsteer-run: %s\n", cvs_last_error());
        return 1;
    }
    const int nworkers = (int)std::min<size_t>(files.size(), (size_t)ndev * 3);  // 3 per GPU: overlap file I/O, copies and kernels
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0}, done{0};
    auto worker = [&](int w) {
        cvs_g2* h = nullptr;
        if (cvs_g2_create(&h, w % ndev, 4, 0.67f) != CVS_OK) {
            fprintf(stderr, "cvsteer-run: %s\n", cvs_last_error());
            ++failed;
            return;
        }
        std::vector<unsigned char> gray, out[3];
        for (size_t i = next++; i < files.size(); i = next++) {
            int rows = 0, cols = 0;
            if (!read_gray(files[i], gray, rows, cols)) continue;  // unreadable: skipped, as the reference does
            for (auto& o : out) o.resize(gray.size());
            if (cvs_g2_lines_u8_host(h, gray.data(), 1, rows, cols, (size_t)cols, gray.size(), a.gain, out[0].data(), out[1].data(), out[2].data(),
                                     (size_t)cols, gray.size()) != CVS_OK) {
                fprintf(stderr, "cvsteer-run: %s: %s\n", files[i].c_str(), cvs_last_error());
                ++failed;
                continue;
            }
            if (!a.output.empty()) {
                const std::string base = a.output + "/" + stem(files[i]);
                const bool png = a.format != "pgm";
                write_gray(base + "_edges", png, out[0].data(), rows, cols);
                write_gray(base + "_lines_dark", png, out[1].data(), rows, cols);
                write_gray(base + "_lines_bright", png, out[2].data(), rows, cols);
            }
            ++done;
            if (a.verbose) printf("%s: %dx%d\n", files[i].c_str(), cols, rows);
        }
        cvs_g2_destroy(h);
    };
    std::vector<std::thread> pool;
    for (int w = 0; w < nworkers; ++w) pool.emplace_back(worker, w);
    for (auto& t : pool) t.join();
    if (a.verbose) printf("%d of %zu files processed\n", done.load(), files.size());
    return failed ? 1 : 0;
}
