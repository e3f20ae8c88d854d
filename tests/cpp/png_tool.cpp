// Test helper for cvsteer_b200/cli/png_io.h (CPU only, no CUDA):  png_tool decode in.png out.pgm | png_tool encode in.pgm out.png
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../cvsteer_b200/cli/png_io.h"

static bool slurp(const char* path, std::vector<unsigned char>& b)
{
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    unsigned char buf[65536];
    for (size_t n; (n = fread(buf, 1, sizeof(buf), f)) > 0;) b.insert(b.end(), buf, buf + n);
    fclose(f);
    return true;
}

int main(int argc, char** argv)
{
    if (argc != 4) return 2;
    std::vector<unsigned char> in, px, out;
    if (!slurp(argv[2], in)) return 3;
    int rows = 0, cols = 0;
    if (!strcmp(argv[1], "decode")) {
        if (!pngio::decode_gray(in.data(), in.size(), px, rows, cols)) return 4;
        char hdr[64];
        const int n = snprintf(hdr, sizeof(hdr), "P5\n%d %d\n255\n", cols, rows);
        out.assign(hdr, hdr + n);
        out.insert(out.end(), px.begin(), px.end());
    } else {
        if (sscanf(reinterpret_cast<const char*>(in.data()), "P5 %d %d 255", &cols, &rows) != 2) return 5;
        const size_t n = (size_t)rows * cols;
        if (in.size() < n) return 5;
        if (!pngio::encode_gray(in.data() + (in.size() - n), rows, cols, out)) return 6;
    }
    FILE* f = fopen(argv[3], "wb");
    if (!f) return 7;
    fwrite(out.data(), 1, out.size(), f);
    fclose(f);
    return 0;
}
