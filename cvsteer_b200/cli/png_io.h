// Minimal PNG codec for cvsteer-run, over zlib only (this image has no libpng): what `cv::imread(name)` followed by
// `cv::cvtColor(image, gray, cv::COLOR_BGR2GRAY)` yields for a PNG file (reference example/steer.cpp:73-82), and what
// `cv::imwrite(name + ".png", Mat1b)` needs (example/steer.cpp:106-121).
//
// Reader: every colour type (gray, RGB, palette, gray+alpha, RGBA), bit depths 1/2/4/8/16, Adam7 interlacing; CRCs are
// checked.  Conversion to 8-bit gray follows OpenCV's 8-bit reading path: sub-byte gray samples are expanded to 0..255,
// 16-bit samples keep their high byte (png_set_strip_16), alpha is dropped, colour goes through the fixed-point
// BGR2GRAY weights of OpenCV 4 (R 9798, G 19235, B 3735, >> 15 with rounding: bit-identical to cv2 4.13.0, the OpenCV this
// repository's oracle is pinned to; OpenCV 3.4 used 4899 / 9617 / 1868 >> 14, which differs by one grey level on 0.3 % of
// random colours), for which R = G = B gives the value itself.
// Writer: 8-bit gray, one IDAT, no filtering.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace pngio {

inline uint32_t be32(const unsigned char* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

inline bool is_png(const unsigned char* p, size_t n)
{
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    return n >= 8 && memcmp(p, sig, 8) == 0;
}

inline unsigned char bgr2gray(unsigned r, unsigned g, unsigned b) { return (unsigned char)((r * 9798u + g * 19235u + b * 3735u + 16384u) >> 15); }

// undo the per-row filters of one (sub-)image in place; `raw` holds rows of 1 filter byte + rowbytes
inline bool unfilter(unsigned char* raw, int rows, size_t rowbytes, int bpp)
{
    const unsigned char* prev = nullptr;
    for (int y = 0; y < rows; ++y) {
        unsigned char* line = raw + (size_t)y * (rowbytes + 1);
        const int ft = line[0];
        unsigned char* cur = line + 1;
        for (size_t i = 0; i < rowbytes; ++i) {
            const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0, c = (prev && i >= (size_t)bpp) ? prev[i - bpp] : 0;
            int pred = 0;
            switch (ft) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: {
                    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
                    pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: return false;
            }
            cur[i] = (unsigned char)(cur[i] + pred);
        }
        prev = cur;
    }
    return true;
}

// 8-bit gray image of a PNG file held in memory; false when the bytes are not a PNG this reader understands
inline bool decode_gray(const unsigned char* file, size_t size, std::vector<unsigned char>& gray, int& rows, int& cols)
{
    if (!is_png(file, size)) return false;
    size_t pos = 8;
    int w = 0, h = 0, depth = 0, ctype = -1, interlace = 0;
    std::vector<unsigned char> idat, plte;
    bool end = false;
    while (!end && pos + 12 <= size) {
        const uint32_t len = be32(file + pos);
        if (pos + 12 + (size_t)len > size) return false;
        const unsigned char* type = file + pos + 4;
        const unsigned char* data = file + pos + 8;
        if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), type, 4 + len) != be32(data + len)) return false;
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) return false;
            w = (int)be32(data), h = (int)be32(data + 4), depth = data[8], ctype = data[9], interlace = data[12];
            if (data[10] != 0 || data[11] != 0 || interlace > 1 || w <= 0 || h <= 0) return false;
        } else if (!memcmp(type, "PLTE", 4)) {
            plte.assign(data, data + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!memcmp(type, "IEND", 4)) {
            end = true;
        }
        pos += 12 + (size_t)len;
    }
    if (ctype < 0 || idat.empty()) return false;
    int ch;
    switch (ctype) {
        case 0: ch = 1; break;
        case 2: ch = 3; break;
        case 3: ch = 1; break;
        case 4: ch = 2; break;
        case 6: ch = 4; break;
        default: return false;
    }
    const bool depth_ok = ctype == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                        : ctype == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8)
                                     : (depth == 8 || depth == 16);
    if (!depth_ok || (ctype == 3 && plte.size() < 3)) return false;
    const int bits = ch * depth, bpp = bits >= 8 ? bits / 8 : 1;
    auto rowbytes_of = [&](int pw) { return ((size_t)pw * bits + 7) / 8; };
    // sub-images: one for a plain file, seven for Adam7
    struct Pass { int x0, y0, dx, dy; };
    static const Pass adam7[7] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    static const Pass plain = {0, 0, 1, 1};
    const int npass = interlace ? 7 : 1;
    size_t total = 0;
    for (int p = 0; p < npass; ++p) {
        const Pass& ps = interlace ? adam7[p] : plain;
        const int pw = (w - ps.x0 + ps.dx - 1) / ps.dx, ph = (h - ps.y0 + ps.dy - 1) / ps.dy;
        if (pw > 0 && ph > 0) total += (size_t)ph * (rowbytes_of(pw) + 1);
    }
    // A deflate stream expands at most ~1032 : 1, so a header that promises more than the IDAT bytes could ever inflate to is
    // damaged (or hostile): reject it before allocating what it asks for.
    if (total > idat.size() * 1040 + 64 || (size_t)w * (size_t)h > ((size_t)1 << 33)) return false;
    std::vector<unsigned char> raw(total);
    uLongf got = (uLongf)total;
    if (uncompress(raw.data(), &got, idat.data(), (uLong)idat.size()) != Z_OK || got != total) return false;
    rows = h, cols = w;
    gray.assign((size_t)w * h, 0);
    auto sample = [&](const unsigned char* line, int i) -> unsigned {  // i-th sample of a row, as an 8-bit quantity (index for palettes)
        if (depth == 8) return line[i];
        if (depth == 16) return line[2 * i];  // high byte (png_set_strip_16)
        const int per = 8 / depth, shift = (per - 1 - i % per) * depth;
        const unsigned v = (line[i / per] >> shift) & ((1u << depth) - 1u);
        return ctype == 3 ? v : v * (255u / ((1u << depth) - 1u));
    };
    size_t off = 0;
    for (int p = 0; p < npass; ++p) {
        const Pass& ps = interlace ? adam7[p] : plain;
        const int pw = (w - ps.x0 + ps.dx - 1) / ps.dx, ph = (h - ps.y0 + ps.dy - 1) / ps.dy;
        if (pw <= 0 || ph <= 0) continue;
        const size_t rb = rowbytes_of(pw);
        if (!unfilter(raw.data() + off, ph, rb, bpp)) return false;
        for (int y = 0; y < ph; ++y) {
            const unsigned char* line = raw.data() + off + (size_t)y * (rb + 1) + 1;
            unsigned char* dst = gray.data() + (size_t)(ps.y0 + y * ps.dy) * w + ps.x0;
            for (int x = 0; x < pw; ++x) {
                unsigned char g;
                if (ctype == 0 || ctype == 4) {
                    g = (unsigned char)sample(line, x * ch);
                } else if (ctype == 3) {
                    const unsigned idx = sample(line, x);
                    if (3 * (size_t)idx + 2 >= plte.size()) return false;
                    g = bgr2gray(plte[3 * idx], plte[3 * idx + 1], plte[3 * idx + 2]);
                } else {
                    g = bgr2gray(sample(line, x * ch), sample(line, x * ch + 1), sample(line, x * ch + 2));
                }
                dst[(size_t)x * ps.dx] = g;
            }
        }
        off += (size_t)ph * (rb + 1);
    }
    return true;
}

inline void put_chunk(std::vector<unsigned char>& out, const char* type, const unsigned char* data, size_t len)
{
    auto be = [&](uint32_t v) { out.push_back(v >> 24), out.push_back(v >> 16 & 255), out.push_back(v >> 8 & 255), out.push_back(v & 255); };
    be((uint32_t)len);
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (len) out.insert(out.end(), data, data + len);
    be((uint32_t)crc32(crc32(0L, Z_NULL, 0), out.data() + start, (uInt)(4 + len)));
}

// 8-bit gray PNG file image (rows x cols, dense)
inline bool encode_gray(const unsigned char* px, int rows, int cols, std::vector<unsigned char>& out)
{
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    out.assign(sig, sig + 8);
    unsigned char ihdr[13] = {0};
    ihdr[0] = cols >> 24, ihdr[1] = cols >> 16 & 255, ihdr[2] = cols >> 8 & 255, ihdr[3] = cols & 255;
    ihdr[4] = rows >> 24, ihdr[5] = rows >> 16 & 255, ihdr[6] = rows >> 8 & 255, ihdr[7] = rows & 255;
    ihdr[8] = 8;  // bit depth; colour type 0 (gray), deflate, adaptive filtering, no interlace
    put_chunk(out, "IHDR", ihdr, 13);
    std::vector<unsigned char> raw((size_t)rows * (cols + 1));
    for (int y = 0; y < rows; ++y) {
        raw[(size_t)y * (cols + 1)] = 0;  // filter type: none
        memcpy(&raw[(size_t)y * (cols + 1) + 1], px + (size_t)y * cols, cols);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<unsigned char> z(clen);
    if (compress2(z.data(), &clen, raw.data(), (uLong)raw.size(), 3) != Z_OK) return false;
    put_chunk(out, "IDAT", z.data(), clen);
    put_chunk(out, "IEND", nullptr, 0);
    return true;
}

}  // namespace pngio
