/*
 * ExrWriter.hpp -- minimal OpenEXR 2 scanline writer (host side, C++17, no OpenEXR dependency).
 *
 * Camera::saveToDisk (DG/Scene/Cameras/Camera.cpp:149-175) dumps the progressive buffer as a single-part scanline EXR:
 * Header(width, height, pixelAspectRatio 1, screenWindowCenter (0, 0), screenWindowWidth 1, DECREASING_Y), three FLOAT
 * channels R, G, B sliced out of the float4 rows, scanline y = buffer row y (so the image is stored bottom row first;
 * DeepestScatter_Train/Utils/GenerateComparisons.py flips it on read).  This writer produces the same file structure
 * from the published OpenEXR file layout -- magic 20000630, version 2, attribute list, offset table, one chunk per
 * scanline {y, byte count, B row, G row, R row} -- with NO_COMPRESSION instead of the library's default ZIP (any EXR
 * reader accepts both).  With DECREASING_Y the chunks are stored from the last scanline to the first; the offset table is
 * always indexed by increasing y.
 */
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace DeepestScatter {

namespace exrdetail {
inline void put(std::vector<uint8_t>& b, const void* p, size_t n)
{
    const uint8_t* c = (const uint8_t*)p;
    b.insert(b.end(), c, c + n);
}
inline void putStr(std::vector<uint8_t>& b, const char* s) { put(b, s, strlen(s) + 1); }
inline void putI32(std::vector<uint8_t>& b, int32_t v) { put(b, &v, 4); }
inline void putF32(std::vector<uint8_t>& b, float v) { put(b, &v, 4); }
inline void attr(std::vector<uint8_t>& b, const char* name, const char* type, const std::vector<uint8_t>& value)
{
    putStr(b, name);
    putStr(b, type);
    putI32(b, (int32_t)value.size());
    put(b, value.data(), value.size());
}
} // namespace exrdetail

/* rgba: float4 [height][width]; channels R, G, B are written, alpha is dropped (as the reference does) */
inline void writeExrRGB(const std::string& path, uint32_t width, uint32_t height, const float* rgba, bool decreasingY = true)
{
    using namespace exrdetail;
    if (width == 0 || height == 0 || !rgba) throw std::runtime_error("writeExrRGB: empty image");
    std::vector<uint8_t> head;
    const uint8_t magic[4] = {0x76, 0x2f, 0x31, 0x01}, version[4] = {2, 0, 0, 0};
    put(head, magic, 4);
    put(head, version, 4);
    {
        std::vector<uint8_t> v; /* chlist: channels in alphabetical order */
        for (const char* name : {"B", "G", "R"}) {
            putStr(v, name);
            putI32(v, 2); /* FLOAT */
            const uint8_t pLinearAndReserved[4] = {0, 0, 0, 0};
            put(v, pLinearAndReserved, 4);
            putI32(v, 1);
            putI32(v, 1);
        }
        v.push_back(0);
        attr(head, "channels", "chlist", v);
    }
    attr(head, "compression", "compression", {0}); /* NO_COMPRESSION */
    {
        std::vector<uint8_t> v;
        putI32(v, 0);
        putI32(v, 0);
        putI32(v, (int32_t)width - 1);
        putI32(v, (int32_t)height - 1);
        attr(head, "dataWindow", "box2i", v);
        attr(head, "displayWindow", "box2i", v);
    }
    attr(head, "lineOrder", "lineOrder", {(uint8_t)(decreasingY ? 1 : 0)});
    {
        /* attributes sorted by name, as the library's header map writes them */
        std::vector<uint8_t> one, centre;
        putF32(one, 1.0f);
        putF32(centre, 0.0f);
        putF32(centre, 0.0f);
        attr(head, "pixelAspectRatio", "float", one);
        attr(head, "screenWindowCenter", "v2f", centre);
        attr(head, "screenWindowWidth", "float", one);
    }
    head.push_back(0); /* end of header */

    const size_t rowBytes = (size_t)width * 4, chunkBytes = 8 + 3 * rowBytes;
    const uint64_t tableAt = head.size(), dataAt = tableAt + 8ull * height;
    std::vector<uint64_t> table(height);
    for (uint32_t k = 0; k < height; k++) {
        const uint32_t y = decreasingY ? height - 1 - k : k; /* k-th chunk in the file */
        table[y] = dataAt + (uint64_t)k * chunkBytes;
    }
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + path);
    bool ok = fwrite(head.data(), 1, head.size(), f) == head.size() && fwrite(table.data(), 8, height, f) == height;
    std::vector<float> row(3 * (size_t)width);
    for (uint32_t k = 0; ok && k < height; k++) {
        const uint32_t y = decreasingY ? height - 1 - k : k;
        const float* src = rgba + 4 * (size_t)y * width;
        for (uint32_t x = 0; x < width; x++) {
            row[x] = src[4 * x + 2];             /* B */
            row[width + x] = src[4 * x + 1];     /* G */
            row[2 * width + x] = src[4 * x];     /* R */
        }
        const int32_t hdr[2] = {(int32_t)y, (int32_t)(3 * rowBytes)};
        ok = fwrite(hdr, 4, 2, f) == 2 && fwrite(row.data(), 4, row.size(), f) == row.size();
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) throw std::runtime_error("short write to " + path);
}

} // namespace DeepestScatter
