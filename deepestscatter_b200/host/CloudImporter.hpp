/*
 * CloudImporter.hpp -- front end of the cloud importer (host side, C++17).
 *
 * The reference reads a Houdini-exported OpenVDB FloatGrid (Resources::loadVolumeBuffer, DG/Util/Resources.cpp:68-155):
 * maximum over the active voxels, active-voxel bounding box expanded by one voxel, dense u8 = uint8(v / max * 255) over
 * that box, then the box-filter mip chain.  OpenVDB does not exist in this environment; the front end accepts:
 *
 *   <file>.vdb                      an OpenVDB file whose first grid is a FloatGrid, read by host/VdbReader.hpp (the container format
 *                                   restated from the OpenVDB sources; unpinned against the library itself, see there): active =
 *                                   the grid's value masks and active tiles, exactly what the reference iterates
 *   <file>.npy                      NumPy array, C order, shape (nz, ny, nx), dtype float32 / float64 / uint8 -- e.g. the
 *                                   output of `pyopenvdb`'s copyToArray, or of Houdini's volume export
 *   synth:<n>[:<kind>[:<seed>]]     the procedural grids of include/ds_synth.h, generated on the device
 *
 * A dense array is treated like the VDB: voxels with a value > 0 are "active"; the grid is cropped to their bounding
 * box, padded by one zero voxel on every side (expandBy(1), Resources.cpp:97-101 -- which is what makes every face
 * voxel zero), and handed to ds_volume_upload_float with the maximum, where the device quantises it exactly as
 * Resources.cpp:137 does (bit-exact with the oracle, tests/test_gpu_parity.py) and builds the mip chain
 * (Resources.cpp:169-209).  uint8 arrays are taken as already quantised (no rescale), cropped and padded the same way.
 * Like Resources::volumeCache (Resources.cpp:22, 73-78) the importer keeps the last cloud: asking for the same path
 * again does not touch the device.
 */
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ds_abi.h"
#include "VdbReader.hpp"

namespace DeepestScatter {

struct DenseGrid {
    int nx = 0, ny = 0, nz = 0;
    bool quantised = false;      /* true: u8 holds the values; false: f32 does */
    std::vector<float> f32;      /* [nz][ny][nx] */
    std::vector<uint8_t> u8;
    double maxDensity = 0.0;
};

namespace detail {

inline std::string npyHeaderField(const std::string& header, const std::string& key)
{
    const size_t k = header.find("'" + key + "'");
    if (k == std::string::npos) throw std::runtime_error("npy header lacks '" + key + "'");
    size_t p = header.find(':', k);
    if (p == std::string::npos) throw std::runtime_error("malformed npy header");
    p++;
    while (p < header.size() && header[p] == ' ') p++;
    size_t e = p;
    if (header[p] == '(') {
        e = header.find(')', p);
        return header.substr(p, e - p + 1);
    }
    while (e < header.size() && header[e] != ',' && header[e] != '}') e++;
    return header.substr(p, e - p);
}

} // namespace detail

/* NumPy .npy, format versions 1.0 - 3.0, little endian, C order, 3-D */
inline DenseGrid readNpy(const std::string& path)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open cloud file " + path);
    struct Closer {
        FILE* f;
        ~Closer() { fclose(f); }
    } closer{f};
    unsigned char pre[12];
    if (fread(pre, 1, 10, f) != 10 || memcmp(pre, "\x93NUMPY", 6) != 0) throw std::runtime_error(path + " is not a .npy file");
    size_t headerLen = pre[8] | (pre[9] << 8);
    if (pre[6] >= 2) {
        if (fread(pre + 10, 1, 2, f) != 2) throw std::runtime_error("truncated npy header");
        headerLen |= ((size_t)pre[10] << 16) | ((size_t)pre[11] << 24);
    }
    std::string header(headerLen, '\0');
    if (fread(&header[0], 1, headerLen, f) != headerLen) throw std::runtime_error("truncated npy header");
    std::string descr = detail::npyHeaderField(header, "descr");
    const std::string order = detail::npyHeaderField(header, "fortran_order");
    const std::string shape = detail::npyHeaderField(header, "shape");
    if (order.find("False") == std::string::npos) throw std::runtime_error("npy array must be C-ordered (z, y, x)");
    long dims[3] = {0, 0, 0};
    if (sscanf(shape.c_str(), "(%ld, %ld, %ld", &dims[0], &dims[1], &dims[2]) != 3) throw std::runtime_error("npy array must be 3-D (nz, ny, nx), shape is " + shape);
    for (long d : dims)
        if (d <= 0 || d > 8192) throw std::runtime_error("unsupported npy shape " + shape);
    DenseGrid g;
    g.nz = (int)dims[0];
    g.ny = (int)dims[1];
    g.nx = (int)dims[2];
    const size_t count = (size_t)g.nx * g.ny * g.nz;
    auto readAll = [&](void* dst, size_t bytes) {
        if (fread(dst, 1, bytes, f) != bytes) throw std::runtime_error("truncated npy data in " + path);
    };
    if (descr.find("f4") != std::string::npos && descr.find('>') == std::string::npos) {
        g.f32.resize(count);
        readAll(g.f32.data(), count * 4);
    } else if (descr.find("f8") != std::string::npos && descr.find('>') == std::string::npos) {
        std::vector<double> tmp(count);
        readAll(tmp.data(), count * 8);
        g.f32.resize(count);
        for (size_t i = 0; i < count; i++) g.f32[i] = (float)tmp[i]; /* FloatGrid holds floats */
    } else if (descr.find("u1") != std::string::npos) {
        g.quantised = true;
        g.u8.resize(count);
        readAll(g.u8.data(), count);
    } else {
        throw std::runtime_error("unsupported npy dtype " + descr + " (float32, float64 or uint8)");
    }
    return g;
}

/* Resources.cpp:95-101: maximum, active bounding box expanded by 1; returns the cropped + padded grid */
inline DenseGrid cropToActive(const DenseGrid& g)
{
    int lo[3] = {g.nx, g.ny, g.nz}, hi[3] = {-1, -1, -1};
    double mx = 0.0;
    for (int z = 0; z < g.nz; z++)
        for (int y = 0; y < g.ny; y++) {
            const size_t row = ((size_t)z * g.ny + y) * g.nx;
            for (int x = 0; x < g.nx; x++) {
                const double v = g.quantised ? (double)g.u8[row + x] : (double)g.f32[row + x];
                if (v > 0.0) {
                    lo[0] = std::min(lo[0], x);
                    lo[1] = std::min(lo[1], y);
                    lo[2] = std::min(lo[2], z);
                    hi[0] = std::max(hi[0], x);
                    hi[1] = std::max(hi[1], y);
                    hi[2] = std::max(hi[2], z);
                    mx = std::max(mx, v);
                }
            }
        }
    if (hi[0] < 0) throw std::runtime_error("cloud grid has no active (non-zero) voxels");
    DenseGrid out;
    out.quantised = g.quantised;
    out.maxDensity = mx;
    /* boundingBox.expandBy(1); max += 1: extent + 2 per axis */
    out.nx = hi[0] - lo[0] + 3;
    out.ny = hi[1] - lo[1] + 3;
    out.nz = hi[2] - lo[2] + 3;
    const size_t count = (size_t)out.nx * out.ny * out.nz;
    if (g.quantised)
        out.u8.assign(count, 0);
    else
        out.f32.assign(count, 0.0f);
    for (int z = lo[2]; z <= hi[2]; z++)
        for (int y = lo[1]; y <= hi[1]; y++) {
            const size_t src = ((size_t)z * g.ny + y) * g.nx + lo[0];
            const size_t dst = ((size_t)(z - lo[2] + 1) * out.ny + (y - lo[1] + 1)) * out.nx + 1;
            const size_t n = (size_t)(hi[0] - lo[0] + 1);
            if (g.quantised)
                memcpy(&out.u8[dst], &g.u8[src], n);
            else
                memcpy(&out.f32[dst], &g.f32[src], n * sizeof(float));
        }
    return out;
}

class CloudImporter {
public:
    explicit CloudImporter(DsContext* ctx) : ctx(ctx) {}

    /* Resources::loadVolumeBuffer(path, createMipmaps): returns the grid size in voxels */
    void load(const std::string& path, bool createMipmaps, int sizeOut[3])
    {
        if (path == cachedPath && createMipmaps == cachedMips && generation() == cachedGeneration) { /* "Using cached." (Resources.cpp:73-78) */
            memcpy(sizeOut, cachedSize, sizeof(cachedSize));
            return;
        }
        if (path.rfind("synth:", 0) == 0) {
            int n = 0, kind = 0;
            unsigned seed = 1234;
            if (sscanf(path.c_str(), "synth:%d:%d:%u", &n, &kind, &seed) < 1 || n < 4) throw std::runtime_error("bad synthetic cloud spec " + path);
            check(ds_volume_synth(ctx, n, kind, seed, createMipmaps ? 1 : 0));
            cachedSize[0] = cachedSize[1] = cachedSize[2] = n;
        } else if (path.size() > 4 && path.compare(path.size() - 4, 4, ".npy") == 0) {
            const DenseGrid g = cropToActive(readNpy(path));
            if (g.quantised)
                check(ds_volume_upload(ctx, g.u8.data(), g.nx, g.ny, g.nz, createMipmaps ? 1 : 0));
            else
                check(ds_volume_upload_float(ctx, g.f32.data(), g.nx, g.ny, g.nz, g.maxDensity, createMipmaps ? 1 : 0));
            cachedSize[0] = g.nx;
            cachedSize[1] = g.ny;
            cachedSize[2] = g.nz;
        } else if (path.size() > 4 && path.compare(path.size() - 4, 4, ".vdb") == 0) {
            /* Resources.cpp:80-141: first grid as FloatGrid, maximum over the active values, active box + 1, accessor values */
            DenseGrid g;
            int dims[3];
            dsvdb::toDense(dsvdb::readFirstFloatGrid(dsvdb::readFile(path)), g.f32, dims, g.maxDensity);
            g.nx = dims[0];
            g.ny = dims[1];
            g.nz = dims[2];
            if (!(g.maxDensity > 0.0)) throw std::runtime_error("cloud grid has no positive active value: " + path);
            check(ds_volume_upload_float(ctx, g.f32.data(), g.nx, g.ny, g.nz, g.maxDensity, createMipmaps ? 1 : 0));
            cachedSize[0] = g.nx;
            cachedSize[1] = g.ny;
            cachedSize[2] = g.nz;
        } else {
            throw std::runtime_error("unsupported cloud file " + path + ": .vdb (OpenVDB FloatGrid), dense .npy (nz, ny, nx) or synth:<n>");
        }
        cachedPath = path;
        cachedMips = createMipmaps;
        cachedGeneration = generation();
        memcpy(sizeOut, cachedSize, sizeof(cachedSize));
    }

private:
    void check(int rc) const
    {
        if (rc != DS_OK) throw std::runtime_error(ds_last_error(ctx));
    }
    /* bumped by every volume upload of the context: a volume replaced behind the importer's back invalidates the cache */
    int generation() const
    {
        int g = 0;
        ds_get_option(ctx, "volume_generation", &g);
        return g;
    }
    DsContext* ctx;
    int cachedGeneration = -1;
    std::string cachedPath;
    bool cachedMips = false;
    int cachedSize[3] = {0, 0, 0};
};

} // namespace DeepestScatter
