// cvsteer-run -- batch driver with the reference CLI's contract (reference example/steer.cpp:59-173), on libcvsteer_b200.
//
//   cvsteer-run --input=<image.pgm | list.txt> --output=<dir> [--gain=<g>] [--verbose] [--help]
//
// Per input file, as ParallelSteerable::operator() does (example/steer.cpp:69-124): gray 8-bit image ->
// SteerableFiltersG2(gray, 4, 0.67f) -> steer(dominant angle, ...) -> findEdges / findDarkLines / findBrightLines(magnitude,
// phase) -> 8-bit (gain > 0: convertTo(CV_8UC1, gain); else normalize(0, 255, NORM_MINMAX)) -> three files
// <dir>/<basename>_{edges,lines_dark,lines_bright}.pgm.  All of it is ONE call, cvs_g2_lines_u8_host, per file; files are
// spread over worker threads (one handle, hence one CUDA stream, each) the way the reference spreads them over
// cv::parallel_for_ (example/steer.cpp:169), and over all visible GPUs round-robin.
//
// Differences from the reference, on purpose:
//  * image files are binary PGM (P5, maxval 255): this build image has no PNG/JPEG codec with C headers.  Unreadable
//    files are skipped silently, like `if (image.empty()) continue;` (example/steer.cpp:74-77).
//  * --gain is honoured.  The reference declares it but passes the `verbose` flag as the gain
//    (example/steer.cpp:167-168), so its effective gain is 0 or 1; pass --gain=1 to reproduce `--verbose` runs.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "cvsteer_c.h"

namespace {

struct Args {
    std::string input, output;
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
        if (val("input", a.input) || val("output", a.output)) continue;
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
           "\t--input\n\t\tinput image (binary PGM) or .txt list of images\n"
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

bool read_pgm(const std::string& path, std::vector<unsigned char>& px, int& rows, int& cols)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    auto token = [&](int& v) {
        int c = fgetc(f);
        for (;;) {
            while (c == ' ' || c == '\t' || c == '\n' || c == '\r') c = fgetc(f);
            if (c != '#') break;
            while (c != '\n' && c != EOF) c = fgetc(f);
        }
        if (c < '0' || c > '9') return false;
        v = 0;
        while (c >= '0' && c <= '9') v = v * 10 + (c - '0'), c = fgetc(f);
        return true;  // exactly one whitespace byte consumed after the number
    };
    char magic[2] = {0, 0};
    int maxv = 0;
    bool ok = fread(magic, 1, 2, f) == 2 && magic[0] == 'P' && magic[1] == '5' && token(cols) && token(rows) && token(maxv) && maxv == 255 &&
              rows > 0 && cols > 0;
    if (ok) {
        px.resize((size_t)rows * cols);
        ok = fread(px.data(), 1, px.size(), f) == px.size();
    }
    fclose(f);
    return ok;
}

bool write_pgm(const std::string& path, const unsigned char* px, int rows, int cols)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    fprintf(f, "P5\n%d %d\n255\n", cols, rows);
    const bool ok = fwrite(px, 1, (size_t)rows * cols, f) == (size_t)rows * cols;
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
            if (!read_pgm(files[i], gray, rows, cols)) continue;  // unreadable: skipped, as the reference does
            for (auto& o : out) o.resize(gray.size());
            if (cvs_g2_lines_u8_host(h, gray.data(), 1, rows, cols, (size_t)cols, gray.size(), a.gain, out[0].data(), out[1].data(), out[2].data(),
                                     (size_t)cols, gray.size()) != CVS_OK) {
                fprintf(stderr, "cvsteer-run: %s: %s\n", files[i].c_str(), cvs_last_error());
                ++failed;
                continue;
            }
            if (!a.output.empty()) {
                const std::string base = a.output + "/" + stem(files[i]);
                write_pgm(base + "_edges.pgm", out[0].data(), rows, cols);
                write_pgm(base + "_lines_dark.pgm", out[1].data(), rows, cols);
                write_pgm(base + "_lines_bright.pgm", out[2].data(), rows, cols);
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
