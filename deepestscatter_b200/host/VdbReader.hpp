/*
 * VdbReader.hpp -- reader of OpenVDB .vdb files for the cloud importer (host side, C++17, zlib only).
 *
 * The reference loads its clouds with OpenVDB 5 (Resources::loadVolumeBuffer, DG/Util/Resources.cpp:80-141):
 *     grids = openvdb::io::Stream(ifile).getGrids();  grid = gridPtrCast<FloatGrid>((*grids)[0]);
 *     maxDensity = extrema over grid->tree().cbeginValueOn();  box = grid->evalActiveVoxelBoundingBox().expandBy(1);
 *     value(x, y, z) = grid->getConstUnsafeAccessor().getValue(min + (x, y, z))
 * OpenVDB is not installed here (nor are Blosc headers), so this file reads the container itself, restated from the published
 * OpenVDB file format (openvdb/io/Archive.cc, GridDescriptor.cc, Compression.h, tree/RootNode.h, InternalNode.h, LeafNode.h,
 * util/NodeMasks.h; file versions 220-224, i.e. OpenVDB 3-8 / Houdini 13+):
 *
 *   header      int64 magic 0x56444220; uint32 file version; uint32 library major, minor; char hasGridOffsets;
 *               [220, 221]: char compressed; 36-char UUID; metadata map; int32 grid count
 *   metadata    uint32 n; n x {string name; string type; uint32 size; bytes}          (string = uint32 length + characters)
 *   descriptor  string uniqueName; string gridType ("Tree_float_5_4_3", suffix "_HalfFloat" = values stored as half);
 *               string instanceParent; int64 gridPos, blockPos, endPos
 *   grid        >= 222: uint32 compression (1 = ZIP, 2 = ACTIVE_MASK, 4 = BLOSC); metadata map; transform (string map type + the
 *               map's doubles); tree topology; tree buffers
 *   topology    uint32 bufferCount; root: float background; uint32 tiles, children; tiles {int32 xyz; float value; bool active};
 *               children {int32 xyz origin; InternalNode<5>}.  InternalNode<L>: child mask, value mask (2^(3L) bits, 64-bit words);
 *               compressed values (one per table slot); then its children in child-mask order, each InternalNode<4> / LeafNode<3>;
 *               LeafNode: value mask (512 bits)
 *   buffers     same traversal; LeafNode: value mask again, compressed values (512)
 *   values      >= 222: int8 metadata (how inactive values are stored), 0-2 inactive values, optional selection mask; the payload
 *               holds all values, or only the active ones under ACTIVE_MASK; ZIP / BLOSC payloads: int64 compressed size
 *               (<= 0: -size raw bytes follow), else zlib stream / Blosc-1 frame (codecs LZ4 and zlib, byte shuffle)
 *   node offset n of local (x, y, z) in a node of dimension 2^L: (x << 2L) | (y << L) | z
 *
 * PARITY: **unpinned against OpenVDB** -- no .vdb file ships with the reference and none exists on this machine, so the reader is
 * checked against a writer of the same format in tests/vdb_writer.py (uncompressed, ZIP, ACTIVE_MASK, half floats, tiles, Blosc
 * frames with stored / LZ4-literal blocks) and against the reference's own loadVolumeBuffer running on the decoded grid
 * (oracle/_ref).  BloscLZ-coded frames (not what OpenVDB writes by default: it asks Blosc for LZ4) are rejected with a clear error.
 */
#pragma once

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace dsvdb {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

enum : uint32_t { COMPRESS_ZIP = 1, COMPRESS_ACTIVE_MASK = 2, COMPRESS_BLOSC = 4 };
enum : int8_t {
    NO_MASK_OR_INACTIVE_VALS = 0,
    NO_MASK_AND_MINUS_BG = 1,
    NO_MASK_AND_ONE_INACTIVE_VAL = 2,
    MASK_AND_NO_INACTIVE_VALS = 3,
    MASK_AND_ONE_INACTIVE_VAL = 4,
    MASK_AND_TWO_INACTIVE_VALS = 5,
    NO_MASK_AND_ALL_VALS = 6
};

/* What the importer needs of a FloatGrid: its leaves, its tiles and its background. */
struct Leaf {
    int32_t origin[3];
    uint64_t mask[8]; /* value mask, bit n = local offset n */
    float values[512];
};
struct Tile {
    int32_t origin[3];
    int32_t dim; /* edge length in voxels: 8 (level 1), 128 (level 2) or 4096 (root) */
    float value;
    bool active;
};
struct FloatGrid {
    std::string name, type;
    float background = 0.0f;
    std::vector<Leaf> leaves;
    std::vector<Tile> tiles;
};

/* ---- byte source ---- */
class Reader {
public:
    Reader(const uint8_t* p, size_t n) : base(p), p(p), end(p + n) {}
    size_t tell() const { return (size_t)(p - base); }
    void seek(size_t off)
    {
        if (off > (size_t)(end - base)) throw Error("seek beyond the end of the file");
        p = base + off;
    }
    void read(void* dst, size_t n)
    {
        if (n > (size_t)(end - p)) throw Error("truncated .vdb file");
        memcpy(dst, p, n);
        p += n;
    }
    void skip(size_t n)
    {
        if (n > (size_t)(end - p)) throw Error("truncated .vdb file");
        p += n;
    }
    template <class T> T get()
    {
        T v;
        read(&v, sizeof(T));
        return v;
    }
    std::string str()
    {
        const uint32_t n = get<uint32_t>();
        if (n > (size_t)(end - p)) throw Error("truncated .vdb file (string)");
        std::string s((const char*)p, n);
        p += n;
        return s;
    }
    const uint8_t* cursor() const { return p; }

private:
    const uint8_t *base, *p, *end;
};

inline float halfToFloat(uint16_t h)
{
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1fu, man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal half: normalise */
            int e = -1;
            uint32_t m = man;
            do {
                e++;
                m <<= 1;
            } while (!(m & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((m & 0x3ffu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

/* ---- payload decoders ---- */
inline void lz4Block(const uint8_t* src, size_t srcLen, uint8_t* dst, size_t dstLen)
{
    const uint8_t *ip = src, *iend = src + srcLen;
    uint8_t *op = dst, *oend = dst + dstLen;
    while (ip < iend) {
        const unsigned token = *ip++;
        size_t lit = token >> 4;
        if (lit == 15) {
            unsigned b;
            do {
                if (ip >= iend) throw Error("LZ4: truncated literal length");
                b = *ip++;
                lit += b;
            } while (b == 255);
        }
        if (lit > (size_t)(iend - ip) || lit > (size_t)(oend - op)) throw Error("LZ4: literal run out of bounds");
        memcpy(op, ip, lit);
        ip += lit;
        op += lit;
        if (ip >= iend) break; /* the last sequence has no match */
        if (iend - ip < 2) throw Error("LZ4: truncated offset");
        const size_t offset = (size_t)ip[0] | ((size_t)ip[1] << 8);
        ip += 2;
        size_t len = token & 15u;
        if (len == 15) {
            unsigned b;
            do {
                if (ip >= iend) throw Error("LZ4: truncated match length");
                b = *ip++;
                len += b;
            } while (b == 255);
        }
        len += 4;
        if (offset == 0 || offset > (size_t)(op - dst) || len > (size_t)(oend - op)) throw Error("LZ4: match out of bounds");
        const uint8_t* m = op - offset;
        for (size_t i = 0; i < len; i++) op[i] = m[i]; /* overlapping copies are the point */
        op += len;
    }
    if (op != oend) throw Error("LZ4: block decodes to the wrong size");
}

inline void inflateInto(const uint8_t* src, size_t srcLen, uint8_t* dst, size_t dstLen)
{
    uLongf got = (uLongf)dstLen;
    if (uncompress(dst, &got, src, (uLong)srcLen) != Z_OK || got != dstLen) throw Error("zlib: corrupt or mis-sized stream");
}

/* Blosc-1 frame: 16-byte header {version, versionlz, flags, typesize, uint32 nbytes, blocksize, cbytes}; flags: 1 = byte shuffle,
 * 2 = stored (memcpy), 4 = bit shuffle, 0x10 = blocks are not split, bits 5-7 = codec (0 BloscLZ, 1 LZ4, 3 zlib); then int32 block
 * offsets; a block is split into `typesize` streams when it is a full block with >= 128 elements per stream and the split flag
 * allows it; a stream whose stored size equals its decoded size is raw. */
inline void bloscFrame(const uint8_t* src, size_t srcLen, uint8_t* dst, size_t dstLen)
{
    if (srcLen < 16) throw Error("Blosc: truncated header");
    const unsigned flags = src[2], typesize = src[3];
    uint32_t nbytes, blocksize, cbytes;
    memcpy(&nbytes, src + 4, 4);
    memcpy(&blocksize, src + 8, 4);
    memcpy(&cbytes, src + 12, 4);
    if (nbytes != dstLen || cbytes > srcLen || blocksize == 0) throw Error("Blosc: header does not match the payload");
    if (flags & 0x2u) {
        if (srcLen < 16 + (size_t)nbytes) throw Error("Blosc: truncated stored frame");
        memcpy(dst, src + 16, nbytes);
        return;
    }
    if (flags & 0x4u) throw Error("Blosc: bit-shuffled frames are not supported");
    const unsigned codec = flags >> 5;
    if (codec != 1 && codec != 3) throw Error("Blosc: codec " + std::to_string(codec) + " is not supported (LZ4 = 1 and zlib = 3 are)");
    const uint32_t nblocks = (nbytes + blocksize - 1) / blocksize;
    if (16 + (size_t)nblocks * 4 > srcLen) throw Error("Blosc: truncated block table");
    std::vector<uint8_t> tmp(blocksize);
    for (uint32_t b = 0; b < nblocks; b++) {
        int32_t start;
        memcpy(&start, src + 16 + 4 * (size_t)b, 4);
        const uint32_t bsize = std::min(blocksize, nbytes - b * blocksize);
        const bool leftover = bsize != blocksize;
        const bool split = !(flags & 0x10u) && typesize <= 16 && !leftover && blocksize / std::max(1u, typesize) >= 128;
        const uint32_t nsplits = split ? typesize : 1;
        const uint32_t neblock = bsize / nsplits;
        const bool shuffled = (flags & 0x1u) && typesize > 1;
        uint8_t* out = shuffled ? tmp.data() : dst + (size_t)b * blocksize;
        size_t ip = (size_t)start;
        for (uint32_t s = 0; s < nsplits; s++) {
            if (ip + 4 > srcLen) throw Error("Blosc: truncated stream");
            int32_t csize;
            memcpy(&csize, src + ip, 4);
            ip += 4;
            if (csize < 0 || ip + (size_t)csize > srcLen) throw Error("Blosc: stream out of bounds");
            if ((uint32_t)csize == neblock)
                memcpy(out + (size_t)s * neblock, src + ip, neblock);
            else if (codec == 1)
                lz4Block(src + ip, (size_t)csize, out + (size_t)s * neblock, neblock);
            else
                inflateInto(src + ip, (size_t)csize, out + (size_t)s * neblock, neblock);
            ip += (size_t)csize;
        }
        if (shuffled) { /* undo the byte transpose: stream j holds byte j of every element */
            uint8_t* d = dst + (size_t)b * blocksize;
            const uint32_t nelem = bsize / typesize;
            for (uint32_t j = 0; j < typesize; j++)
                for (uint32_t i = 0; i < nelem; i++) d[(size_t)i * typesize + j] = tmp[(size_t)j * nelem + i];
            const uint32_t tail = bsize - nelem * typesize; /* bytes beyond the last whole element are copied as they are */
            memcpy(d + (size_t)nelem * typesize, tmp.data() + (size_t)nelem * typesize, tail);
        }
    }
}

/* io::readData<T>: `bytes` decoded bytes from the stream under the grid's compression flags */
inline void readPayload(Reader& r, uint32_t compression, uint8_t* dst, size_t bytes)
{
    if (compression & (COMPRESS_BLOSC | COMPRESS_ZIP)) {
        const int64_t n = r.get<int64_t>();
        if (n <= 0) { /* stored uncompressed */
            if ((uint64_t)(-n) != bytes) throw Error("stored payload has the wrong size");
            r.read(dst, bytes);
            return;
        }
        const uint8_t* src = r.cursor();
        r.skip((size_t)n);
        if (compression & COMPRESS_BLOSC)
            bloscFrame(src, (size_t)n, dst, bytes);
        else
            inflateInto(src, (size_t)n, dst, bytes);
        return;
    }
    r.read(dst, bytes);
}

struct GridState {
    uint32_t version = 0, compression = 0;
    bool half = false;
    float background = 0.0f;
};

/* io::readCompressedValues for float grids: `count` values of a node whose value mask is `mask` */
inline void readValues(Reader& r, const GridState& g, float* dst, uint32_t count, const uint64_t* mask)
{
    int8_t metadata = NO_MASK_AND_ALL_VALS;
    if (g.version >= 222) metadata = r.get<int8_t>();
    if (metadata < 0 || metadata > NO_MASK_AND_ALL_VALS) throw Error("unknown node value metadata");
    float inactive1 = g.background;
    float inactive0 = metadata == NO_MASK_OR_INACTIVE_VALS ? g.background : -g.background;
    if (metadata == NO_MASK_AND_ONE_INACTIVE_VAL || metadata == MASK_AND_ONE_INACTIVE_VAL || metadata == MASK_AND_TWO_INACTIVE_VALS) {
        inactive0 = r.get<float>();
        if (metadata == MASK_AND_TWO_INACTIVE_VALS) inactive1 = r.get<float>();
    }
    std::vector<uint64_t> selection;
    if (metadata == MASK_AND_NO_INACTIVE_VALS || metadata == MASK_AND_ONE_INACTIVE_VAL || metadata == MASK_AND_TWO_INACTIVE_VALS) {
        selection.resize(count / 64);
        r.read(selection.data(), selection.size() * 8);
    }
    uint32_t stored = count;
    const bool maskCompressed = (g.compression & COMPRESS_ACTIVE_MASK) && metadata != NO_MASK_AND_ALL_VALS && g.version >= 222;
    if (maskCompressed) {
        stored = 0;
        for (uint32_t w = 0; w < count / 64; w++) stored += (uint32_t)__builtin_popcountll(mask[w]);
    }
    std::vector<float> tmp(stored);
    if (g.half) {
        std::vector<uint16_t> h(stored);
        readPayload(r, g.compression, (uint8_t*)h.data(), (size_t)stored * 2);
        for (uint32_t i = 0; i < stored; i++) tmp[i] = halfToFloat(h[i]);
    } else {
        readPayload(r, g.compression, (uint8_t*)tmp.data(), (size_t)stored * 4);
    }
    if (!maskCompressed || stored == count) {
        memcpy(dst, tmp.data(), (size_t)count * 4);
        return;
    }
    uint32_t t = 0;
    for (uint32_t i = 0; i < count; i++) {
        if ((mask[i >> 6] >> (i & 63)) & 1u)
            dst[i] = tmp[t++];
        else
            dst[i] = (!selection.empty() && ((selection[i >> 6] >> (i & 63)) & 1u)) ? inactive1 : inactive0;
    }
}

/* ---- tree ---- */
struct Internal {
    int log2dim, childTotal; /* this node's log2 dimension; log2 of its child's total edge (voxels) */
    int32_t origin[3];
    std::vector<uint64_t> childMask, valueMask;
    std::vector<std::unique_ptr<Internal>> children; /* log2dim 5: InternalNode<4> children */
    std::vector<size_t> leafIndex;                   /* log2dim 4: indices into FloatGrid::leaves, in child-mask order */
};

inline void offsetToLocal(uint32_t n, int log2dim, int32_t out[3])
{
    out[0] = (int32_t)(n >> (2 * log2dim));
    out[1] = (int32_t)((n >> log2dim) & ((1u << log2dim) - 1u));
    out[2] = (int32_t)(n & ((1u << log2dim) - 1u));
}

inline std::unique_ptr<Internal> readInternalTopology(Reader& r, const GridState& g, FloatGrid& grid, const int32_t origin[3], int log2dim)
{
    if (g.version < 214) throw Error(".vdb files older than format 214 are not supported");
    std::unique_ptr<Internal> node(new Internal());
    node->log2dim = log2dim;
    node->childTotal = log2dim == 5 ? 7 : 3; /* InternalNode<5> holds InternalNode<4> (128 voxels); InternalNode<4> holds leaves (8) */
    memcpy(node->origin, origin, 12);
    const uint32_t slots = 1u << (3 * log2dim), words = slots / 64;
    node->childMask.resize(words);
    node->valueMask.resize(words);
    r.read(node->childMask.data(), words * 8);
    r.read(node->valueMask.data(), words * 8);
    /* tile values: one per slot (>= 222), or one per slot without a child (214-221) */
    const bool oldVersion = g.version < 222;
    uint32_t numValues = slots;
    if (oldVersion) {
        numValues = 0;
        for (uint32_t w = 0; w < words; w++) numValues += 64u - (uint32_t)__builtin_popcountll(node->childMask[w]);
    }
    std::vector<float> values(std::max(numValues, slots));
    readValues(r, g, values.data(), numValues, node->valueMask.data());
    uint32_t next = 0;
    for (uint32_t n = 0; n < slots; n++) {
        const bool child = (node->childMask[n >> 6] >> (n & 63)) & 1u;
        if (child) continue;
        const float v = oldVersion ? values[next++] : values[n];
        const bool active = (node->valueMask[n >> 6] >> (n & 63)) & 1u;
        if (active || v != g.background) { /* inactive background tiles change nothing */
            Tile t;
            int32_t l[3];
            offsetToLocal(n, log2dim, l);
            for (int a = 0; a < 3; a++) t.origin[a] = origin[a] + (l[a] << node->childTotal);
            t.dim = 1 << node->childTotal;
            t.value = v;
            t.active = active;
            grid.tiles.push_back(t);
        }
    }
    for (uint32_t n = 0; n < slots; n++) {
        if (!((node->childMask[n >> 6] >> (n & 63)) & 1u)) continue;
        int32_t l[3], co[3];
        offsetToLocal(n, log2dim, l);
        for (int a = 0; a < 3; a++) co[a] = origin[a] + (l[a] << node->childTotal);
        if (log2dim == 5) {
            node->children.push_back(readInternalTopology(r, g, grid, co, 4));
        } else {
            Leaf leaf;
            memcpy(leaf.origin, co, 12);
            r.read(leaf.mask, 64);
            for (float& v : leaf.values) v = g.background;
            node->leafIndex.push_back(grid.leaves.size());
            grid.leaves.push_back(leaf);
        }
    }
    return node;
}

inline void readInternalBuffers(Reader& r, const GridState& g, FloatGrid& grid, const Internal& node)
{
    if (node.log2dim == 5) {
        for (const auto& c : node.children) readInternalBuffers(r, g, grid, *c);
        return;
    }
    for (size_t li : node.leafIndex) {
        Leaf& leaf = grid.leaves[li];
        r.read(leaf.mask, 64); /* LeafNode::readBuffers reads the value mask again */
        if (g.version < 222) {
            r.skip(12); /* origin */
            const int8_t numBuffers = r.get<int8_t>();
            if (numBuffers != 1) throw Error("leaf nodes with auxiliary buffers are not supported");
        }
        readValues(r, g, leaf.values, 512, leaf.mask);
    }
}

inline void skipMetaMap(Reader& r)
{
    const uint32_t n = r.get<uint32_t>();
    for (uint32_t i = 0; i < n; i++) {
        r.str();
        r.str();
        r.skip(r.get<uint32_t>());
    }
}

inline void skipTransform(Reader& r)
{
    const std::string type = r.str();
    size_t doubles;
    if (type == "UniformScaleMap" || type == "ScaleMap")
        doubles = 15;
    else if (type == "UniformScaleTranslateMap" || type == "ScaleTranslateMap")
        doubles = 18;
    else if (type == "TranslationMap")
        doubles = 3;
    else if (type == "AffineMap" || type == "UnitaryMap")
        doubles = 16;
    else
        throw Error("unsupported transform map " + type);
    r.skip(doubles * 8);
}

/* The first grid of the file as a FloatGrid (Resources.cpp:87-88 takes (*grids)[0] and casts it) */
inline FloatGrid readFirstFloatGrid(const std::vector<uint8_t>& file)
{
    Reader r(file.data(), file.size());
    if (r.get<int64_t>() != 0x56444220ll) throw Error("not an OpenVDB file (bad magic)");
    GridState g;
    g.version = r.get<uint32_t>();
    if (g.version < 220 || g.version > 224) throw Error("unsupported .vdb file format version " + std::to_string(g.version) + " (220-224 are supported)");
    r.skip(8); /* library major, minor */
    const char hasOffsets = r.get<char>();
    (void)hasOffsets;
    if (g.version < 222) g.compression = r.get<char>() ? (COMPRESS_ZIP | COMPRESS_ACTIVE_MASK) : 0; /* one flag for the whole file */
    r.skip(36); /* UUID text */
    skipMetaMap(r);
    const int32_t gridCount = r.get<int32_t>();
    if (gridCount < 1) throw Error(".vdb file holds no grid");
    FloatGrid grid;
    grid.name = r.str();
    grid.type = r.str();
    const std::string halfSuffix = "_HalfFloat";
    if (grid.type.size() > halfSuffix.size() && grid.type.compare(grid.type.size() - halfSuffix.size(), halfSuffix.size(), halfSuffix) == 0) {
        g.half = true;
        grid.type.resize(grid.type.size() - halfSuffix.size());
    }
    if (grid.type != "Tree_float_5_4_3") throw Error("the first grid is a " + grid.type + ", not a FloatGrid (Tree_float_5_4_3)");
    const std::string instanceParent = r.str();
    if (!instanceParent.empty()) throw Error("instanced grids are not supported");
    r.skip(24); /* gridPos, blockPos, endPos: the grid follows its descriptor */
    if (g.version >= 222) g.compression = r.get<uint32_t>();
    skipMetaMap(r);
    skipTransform(r);
    const uint32_t bufferCount = r.get<uint32_t>();
    if (bufferCount != 1) throw Error("multi-buffer trees are not supported");
    g.background = grid.background = r.get<float>();
    const uint32_t numTiles = r.get<uint32_t>(), numChildren = r.get<uint32_t>();
    for (uint32_t i = 0; i < numTiles; i++) {
        Tile t;
        r.read(t.origin, 12);
        t.value = r.get<float>();
        t.active = r.get<uint8_t>() != 0;
        t.dim = 4096;
        if (t.active || t.value != g.background) grid.tiles.push_back(t);
    }
    std::vector<std::unique_ptr<Internal>> roots;
    for (uint32_t i = 0; i < numChildren; i++) {
        int32_t origin[3];
        r.read(origin, 12);
        roots.push_back(readInternalTopology(r, g, grid, origin, 5));
    }
    for (const auto& n : roots) readInternalBuffers(r, g, grid, *n);
    return grid;
}

inline std::vector<uint8_t> readFile(const std::string& path)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw Error("cannot open cloud file " + path);
    std::vector<uint8_t> data;
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) data.insert(data.end(), buf, buf + n);
    fclose(f);
    return data;
}

/*
 * Resources.cpp:91-141 on the decoded grid: maximum over the ACTIVE values (voxels and tiles), the active bounding box expanded
 * by one voxel, and the accessor's value at every voxel of that box (leaf voxel, else the smallest enclosing tile, else the
 * background).  out is [nz][ny][nx], x fastest, like the reference's densityBuffer.
 */
inline void toDense(const FloatGrid& grid, std::vector<float>& out, int dims[3], double& maxActive)
{
    int64_t lo[3] = {std::numeric_limits<int64_t>::max(), std::numeric_limits<int64_t>::max(), std::numeric_limits<int64_t>::max()};
    int64_t hi[3] = {std::numeric_limits<int64_t>::min(), std::numeric_limits<int64_t>::min(), std::numeric_limits<int64_t>::min()};
    bool any = false;
    maxActive = -std::numeric_limits<double>::max();
    auto grow = [&](const int64_t a[3], const int64_t b[3]) {
        for (int k = 0; k < 3; k++) {
            lo[k] = std::min(lo[k], a[k]);
            hi[k] = std::max(hi[k], b[k]);
        }
        any = true;
    };
    for (const Leaf& l : grid.leaves)
        for (uint32_t n = 0; n < 512; n++)
            if ((l.mask[n >> 6] >> (n & 63)) & 1u) {
                int32_t p[3];
                offsetToLocal(n, 3, p);
                const int64_t c[3] = {l.origin[0] + p[0], l.origin[1] + p[1], l.origin[2] + p[2]};
                grow(c, c);
                maxActive = std::max(maxActive, (double)l.values[n]);
            }
    for (const Tile& t : grid.tiles)
        if (t.active) {
            const int64_t a[3] = {t.origin[0], t.origin[1], t.origin[2]};
            const int64_t b[3] = {(int64_t)t.origin[0] + t.dim - 1, (int64_t)t.origin[1] + t.dim - 1, (int64_t)t.origin[2] + t.dim - 1};
            grow(a, b);
            maxActive = std::max(maxActive, (double)t.value);
        }
    if (!any) throw Error("the grid has no active voxels");
    int64_t n[3];
    for (int k = 0; k < 3; k++) {
        lo[k] -= 1; /* boundingBox.expandBy(1) */
        hi[k] += 1;
        n[k] = hi[k] + 1 - lo[k];
        if (n[k] > 8192) throw Error("active bounding box larger than 8192 voxels: not a cloud this importer accepts");
        dims[k] = (int)n[k];
    }
    out.assign((size_t)n[0] * n[1] * n[2], grid.background);
    auto at = [&](int64_t x, int64_t y, int64_t z) -> float* {
        if (x < lo[0] || x > hi[0] || y < lo[1] || y > hi[1] || z < lo[2] || z > hi[2]) return nullptr;
        return &out[((size_t)(z - lo[2]) * n[1] + (size_t)(y - lo[1])) * n[0] + (size_t)(x - lo[0])];
    };
    /* coarse to fine: tiles first (largest first), then leaves overwrite */
    std::vector<const Tile*> tiles;
    for (const Tile& t : grid.tiles) tiles.push_back(&t);
    std::sort(tiles.begin(), tiles.end(), [](const Tile* a, const Tile* b) { return a->dim > b->dim; });
    for (const Tile* t : tiles) {
        const int64_t x0 = std::max<int64_t>(t->origin[0], lo[0]), x1 = std::min<int64_t>((int64_t)t->origin[0] + t->dim - 1, hi[0]);
        const int64_t y0 = std::max<int64_t>(t->origin[1], lo[1]), y1 = std::min<int64_t>((int64_t)t->origin[1] + t->dim - 1, hi[1]);
        const int64_t z0 = std::max<int64_t>(t->origin[2], lo[2]), z1 = std::min<int64_t>((int64_t)t->origin[2] + t->dim - 1, hi[2]);
        for (int64_t z = z0; z <= z1; z++)
            for (int64_t y = y0; y <= y1; y++)
                for (int64_t x = x0; x <= x1; x++) *at(x, y, z) = t->value;
    }
    for (const Leaf& l : grid.leaves)
        for (uint32_t k = 0; k < 512; k++) {
            int32_t p[3];
            offsetToLocal(k, 3, p);
            if (float* v = at((int64_t)l.origin[0] + p[0], (int64_t)l.origin[1] + p[1], (int64_t)l.origin[2] + p[2])) *v = l.values[k];
        }
}

} // namespace dsvdb
