/* TEST INFRASTRUCTURE (oracle/_ref build only).  OpenVDB 5.0 (Dependencies.md:8) is not vendored and not installed, and the
 * reference ships no .vdb file.  This header gives Resources::loadVolumeBuffer (Util/Resources.cpp:68-155) the calls it makes,
 * backed by a DENSE float grid read from a little container the tests write ("DSDENSE1", int32 nx ny nz, int32 origin xyz,
 * float32 values [z][y][x]).  A voxel is ACTIVE iff its value is non-zero (background 0), which is what a VDB written from a
 * dense Houdini volume with background 0 holds.  evalActiveVoxelBoundingBox / expandBy / Coord arithmetic / Extrema follow the
 * OpenVDB definitions (inclusive integer boxes). */
#pragma once
#include <cstdint>
#include <cstring>
#include <istream>
#include <limits>
#include <memory>
#include <stdexcept>
#include <vector>
namespace openvdb {
inline void initialize() {}
class Coord {
public:
    Coord() : v{0, 0, 0} {}
    Coord(int32_t x, int32_t y, int32_t z) : v{x, y, z} {}
    int32_t x() const { return v[0]; }
    int32_t y() const { return v[1]; }
    int32_t z() const { return v[2]; }
    Coord operator+(const Coord& o) const { return Coord(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    Coord operator-(const Coord& o) const { return Coord(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Coord offsetBy(int32_t n) const { return Coord(v[0] + n, v[1] + n, v[2] + n); }

private:
    int32_t v[3];
};
class CoordBBox {
public:
    CoordBBox() : mn(std::numeric_limits<int32_t>::max(), std::numeric_limits<int32_t>::max(), std::numeric_limits<int32_t>::max()),
                  mx(std::numeric_limits<int32_t>::min(), std::numeric_limits<int32_t>::min(), std::numeric_limits<int32_t>::min()) {}
    CoordBBox(const Coord& a, const Coord& b) : mn(a), mx(b) {}
    const Coord& min() const { return mn; }
    const Coord& max() const { return mx; }
    CoordBBox expandBy(int32_t padding) const { return CoordBBox(mn.offsetBy(-padding), mx.offsetBy(padding)); }

private:
    Coord mn, mx;
};
namespace math {
class Extrema {
public:
    void add(double val)
    {
        n++;
        if (val < mn) mn = val;
        if (val > mx) mx = val;
    }
    double min() const { return mn; }
    double max() const { return mx; }

private:
    uint64_t n = 0;
    double mn = std::numeric_limits<double>::max(), mx = -std::numeric_limits<double>::max();
};
} // namespace math
class GridBase {
public:
    typedef std::shared_ptr<GridBase> Ptr;
    virtual ~GridBase() = default;
};
typedef std::vector<GridBase::Ptr> GridPtrVec;
typedef std::shared_ptr<GridPtrVec> GridPtrVecPtr;
class FloatGrid : public GridBase {
public:
    typedef std::shared_ptr<FloatGrid> Ptr;
    int32_t n[3] = {0, 0, 0}, origin[3] = {0, 0, 0};
    std::vector<float> values;
    float at(int32_t x, int32_t y, int32_t z) const
    {
        x -= origin[0]; y -= origin[1]; z -= origin[2];
        if (x < 0 || y < 0 || z < 0 || x >= n[0] || y >= n[1] || z >= n[2]) return 0.0f; /* background */
        return values[((size_t)z * n[1] + y) * n[0] + x];
    }
    /* iterates the active (non-zero) voxels */
    class ValueOnCIter {
    public:
        explicit ValueOnCIter(const FloatGrid* g) : grid(g), i(0) { skip(); }
        operator bool() const { return i < grid->values.size(); }
        ValueOnCIter& operator++()
        {
            i++;
            skip();
            return *this;
        }
        const float& operator*() const { return grid->values[i]; }

    private:
        void skip()
        {
            while (i < grid->values.size() && grid->values[i] == 0.0f) i++;
        }
        const FloatGrid* grid;
        size_t i;
    };
    struct Tree {
        const FloatGrid* grid;
        ValueOnCIter cbeginValueOn() const { return ValueOnCIter(grid); }
    };
    Tree tree() const { return Tree{this}; }
    struct Accessor {
        const FloatGrid* grid;
        float getValue(const Coord& c) const { return grid->at(c.x(), c.y(), c.z()); }
    };
    Accessor getConstUnsafeAccessor() const { return Accessor{this}; }
    CoordBBox evalActiveVoxelBoundingBox() const
    {
        int32_t lo[3] = {std::numeric_limits<int32_t>::max(), std::numeric_limits<int32_t>::max(), std::numeric_limits<int32_t>::max()};
        int32_t hi[3] = {std::numeric_limits<int32_t>::min(), std::numeric_limits<int32_t>::min(), std::numeric_limits<int32_t>::min()};
        bool any = false;
        for (int32_t z = 0; z < n[2]; z++)
            for (int32_t y = 0; y < n[1]; y++)
                for (int32_t x = 0; x < n[0]; x++)
                    if (values[((size_t)z * n[1] + y) * n[0] + x] != 0.0f) {
                        const int32_t c[3] = {x + origin[0], y + origin[1], z + origin[2]};
                        for (int a = 0; a < 3; a++) {
                            if (c[a] < lo[a]) lo[a] = c[a];
                            if (c[a] > hi[a]) hi[a] = c[a];
                        }
                        any = true;
                    }
        if (!any) return CoordBBox(Coord(0, 0, 0), Coord(-1, -1, -1)); /* OpenVDB: an empty box */
        return CoordBBox(Coord(lo[0], lo[1], lo[2]), Coord(hi[0], hi[1], hi[2]));
    }
};
template <class GridType> typename GridType::Ptr gridPtrCast(const GridBase::Ptr& g) { return std::dynamic_pointer_cast<GridType>(g); }
namespace io {
class Stream {
public:
    explicit Stream(std::istream& is)
    {
        char magic[8];
        is.read(magic, 8);
        if (!is || std::memcmp(magic, "DSDENSE1", 8) != 0) throw std::runtime_error("dsref: not a DSDENSE1 grid container");
        auto g = std::make_shared<FloatGrid>();
        is.read((char*)g->n, 12);
        is.read((char*)g->origin, 12);
        g->values.resize((size_t)g->n[0] * g->n[1] * g->n[2]);
        is.read((char*)g->values.data(), (std::streamsize)(g->values.size() * sizeof(float)));
        if (!is) throw std::runtime_error("dsref: truncated DSDENSE1 grid container");
        grids = std::make_shared<GridPtrVec>();
        grids->push_back(g);
    }
    GridPtrVecPtr getGrids() { return grids; }

private:
    GridPtrVecPtr grids;
};
} // namespace io
namespace tools {
template <class IterT, class OpT> math::Extrema extrema(const IterT& iter, const OpT& op, bool /*threaded*/)
{
    math::Extrema ex;
    for (IterT it = iter; it; ++it) op(it, ex);
    return ex;
}
} // namespace tools
} // namespace openvdb
