/* TEST INFRASTRUCTURE (oracle/_ref build only): the slice of OpenEXR that Camera::saveToDisk (Camera.cpp:149-175) uses.  The
 * "file" is kept in memory (dsref::lastExr) so a test can look at what the reference would have written. */
#pragma once
#include <cstddef>
#include <map>
#include <string>
#include <vector>
namespace Imath {
struct V2f {
    float x, y;
    V2f(float a, float b) : x(a), y(b) {}
};
} // namespace Imath
namespace dsref {
struct ExrImage {
    std::string path;
    int width = 0, height = 0;
    bool decreasingY = false;
    std::vector<float> rgb; /* [row in FILE order][x][3] */
};
ExrImage& lastExr();
} // namespace dsref
namespace Imf {
enum PixelType { UINT = 0, HALF = 1, FLOAT = 2 };
enum LineOrder { INCREASING_Y = 0, DECREASING_Y = 1 };
struct Channel {
    PixelType type;
    explicit Channel(PixelType t = HALF) : type(t) {}
};
struct ChannelList {
    std::map<std::string, Channel> channels;
    void insert(const char* name, const Channel& c) { channels[name] = c; }
};
struct Header {
    int width, height;
    LineOrder order;
    ChannelList list;
    Header(int w, int h, float = 1, const Imath::V2f& = Imath::V2f(0, 0), float = 1, LineOrder o = INCREASING_Y) : width(w), height(h), order(o) {}
    ChannelList& channels() { return list; }
};
struct Slice {
    PixelType type;
    char* base;
    size_t xStride, yStride;
    Slice(PixelType t = HALF, char* b = nullptr, size_t xs = 0, size_t ys = 0) : type(t), base(b), xStride(xs), yStride(ys) {}
};
struct FrameBuffer {
    std::map<std::string, Slice> slices;
    void insert(const char* name, const Slice& s) { slices[name] = s; }
};
class OutputFile {
public:
    OutputFile(const char* path, const Header& h) : path_(path), header_(h) {}
    void setFrameBuffer(const FrameBuffer& fb) { fb_ = fb; }
    /* scan lines are written in the header's line order; pixel (x, y) is read at base + x*xStride + y*yStride */
    void writePixels(int numScanLines)
    {
        dsref::ExrImage& img = dsref::lastExr();
        img.path = path_;
        img.width = header_.width;
        img.height = numScanLines;
        img.decreasingY = header_.order == DECREASING_Y;
        img.rgb.assign((size_t)img.width * img.height * 3, 0.0f);
        const char* names[3] = {"R", "G", "B"};
        for (int row = 0; row < numScanLines; row++) {
            const int y = img.decreasingY ? header_.height - 1 - row : row;
            for (int x = 0; x < header_.width; x++)
                for (int c = 0; c < 3; c++) {
                    const Slice& s = fb_.slices[names[c]];
                    img.rgb[((size_t)row * img.width + x) * 3 + c] = *(const float*)(s.base + (size_t)x * s.xStride + (size_t)y * s.yStride);
                }
        }
    }

private:
    std::string path_;
    Header header_;
    FrameBuffer fb_;
};
} // namespace Imf
