"""Test-side WRITER of OpenVDB .vdb files (format versions 222-224, FloatGrid = Tree_float_5_4_3), pure Python.

TEST INFRASTRUCTURE.  Neither OpenVDB nor a .vdb file exists on this machine, so the reader under test
(deepestscatter_b200/host/VdbReader.hpp) is exercised with files produced here from the same published container layout
(openvdb/io/Archive.cc, Compression.h, tree/RootNode.h, InternalNode.h, LeafNode.h).  The writer covers what a Houdini export can
contain: leaf nodes and tiles at every level, the seven inactive-value encodings of the node-mask compression, ZIP and ACTIVE_MASK
compression, half-float storage, and Blosc frames (stored frames and LZ4 blocks made of literals only -- no Blosc encoder exists
here either).
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

COMPRESS_ZIP, COMPRESS_ACTIVE_MASK, COMPRESS_BLOSC = 1, 2, 4
(NO_MASK_OR_INACTIVE_VALS, NO_MASK_AND_MINUS_BG, NO_MASK_AND_ONE_INACTIVE_VAL, MASK_AND_NO_INACTIVE_VALS, MASK_AND_ONE_INACTIVE_VAL,
 MASK_AND_TWO_INACTIVE_VALS, NO_MASK_AND_ALL_VALS) = range(7)


def _string(s: str) -> bytes:
    b = s.encode()
    return struct.pack("<I", len(b)) + b


def _meta_map(items) -> bytes:
    out = struct.pack("<I", len(items))
    for name, type_name, payload in items:
        out += _string(name) + _string(type_name) + struct.pack("<I", len(payload)) + payload
    return out


def _mask_words(bits: np.ndarray) -> bytes:
    """NodeMask::save: 64-bit words, bit n of the mask = bit (n & 63) of word (n >> 6)."""
    return np.packbits(bits.astype(np.uint8), bitorder="little").tobytes()


def _lz4_literals(data: bytes) -> bytes:
    """A valid LZ4 block that is one literal run (the last sequence of a block carries no match)."""
    n = len(data)
    out = bytearray()
    if n < 15:
        out.append(n << 4)
    else:
        out.append(0xF0)
        rest = n - 15
        while rest >= 255:
            out.append(255)
            rest -= 255
        out.append(rest)
    return bytes(out) + data


def _blosc_frame(data: bytes, typesize: int, mode: str) -> bytes:
    """Blosc-1 frame.  mode 'stored': the memcpy flag; 'lz4': one block per 4096 bytes, byte-shuffled, split into `typesize` streams when
    the block is full and has >= 128 elements per stream, each stream an LZ4 literal run or raw (stored size == decoded size)."""
    nbytes = len(data)
    if mode == "stored":
        return struct.pack("<BBBBIII", 2, 1, 0x2 | 0x1, typesize, nbytes, nbytes, nbytes + 16) + data
    blocksize = 4096
    nblocks = (nbytes + blocksize - 1) // blocksize
    flags = 0x1 | (1 << 5)  # byte shuffle, codec LZ4
    body = bytearray()
    starts = []
    for b in range(nblocks):
        chunk = data[b * blocksize:(b + 1) * blocksize]
        bsize = len(chunk)
        leftover = bsize != blocksize
        split = typesize <= 16 and not leftover and blocksize // typesize >= 128
        nsplits = typesize if split else 1
        if typesize > 1:  # shuffle: stream j holds byte j of every element; trailing bytes beyond whole elements stay in place
            nelem = bsize // typesize
            a = np.frombuffer(chunk[: nelem * typesize], dtype=np.uint8).reshape(nelem, typesize)
            chunk = a.T.tobytes() + chunk[nelem * typesize:]
        starts.append(16 + 4 * nblocks + len(body))
        ne = bsize // nsplits
        for s in range(nsplits):
            piece = chunk[s * ne:(s + 1) * ne]
            if (b + s) % 3 == 0:  # a raw stream: its stored size equals its decoded size
                body += struct.pack("<i", ne) + piece
            else:
                enc = _lz4_literals(piece)
                body += struct.pack("<i", len(enc)) + enc
    total = 16 + 4 * nblocks + len(body)
    return struct.pack("<BBBBIII", 2, 1, flags, typesize, nbytes, blocksize, total) + struct.pack(f"<{nblocks}i", *starts) + bytes(body)


class VdbWriter:
    def __init__(self, background=0.0, compression=0, half=False, version=224, blosc_mode="lz4", zip_threshold=0):
        self.background = np.float32(background)
        self.compression = compression
        self.half = half
        self.version = version
        self.blosc_mode = blosc_mode
        self.leaves = {}      # origin (x, y, z) multiple of 8 -> (values[512] float32, mask[512] bool, metadata or None)
        self.tiles = {}       # (level, origin) -> (value, active); level 1: 8^3 tile inside an InternalNode<4>, 2: 128^3, 3: root 4096^3

    # ---- content ----
    def set_leaf(self, origin, values, mask, metadata=None):
        assert all(o % 8 == 0 for o in origin)
        self.leaves[tuple(int(o) for o in origin)] = (np.asarray(values, np.float32).reshape(512).copy(), np.asarray(mask, bool).reshape(512).copy(), metadata)

    def set_tile(self, level, origin, value, active):
        dim = {1: 8, 2: 128, 3: 4096}[level]
        assert all(o % dim == 0 for o in origin)
        self.tiles[(level, tuple(int(o) for o in origin))] = (np.float32(value), bool(active))

    def from_dense(self, dense, origin=(0, 0, 0), inactive_policy=None):
        """Voxels != 0 become active leaf voxels (x-major leaf offsets: n = (x << 6) | (y << 3) | z); dense is [z][y][x]."""
        nz, ny, nx = dense.shape
        ox, oy, oz = origin
        for z0 in range((oz // 8) * 8, oz + nz, 8):
            for y0 in range((oy // 8) * 8, oy + ny, 8):
                for x0 in range((ox // 8) * 8, ox + nx, 8):
                    vals = np.full((8, 8, 8), self.background, np.float32)  # [x][y][z]
                    for lx in range(8):
                        gx = x0 + lx - ox
                        if not 0 <= gx < nx:
                            continue
                        for ly in range(8):
                            gy = y0 + ly - oy
                            if not 0 <= gy < ny:
                                continue
                            z_lo, z_hi = max(0, oz - z0), min(8, oz + nz - z0)
                            if z_lo < z_hi:
                                vals[lx, ly, z_lo:z_hi] = dense[z0 - oz + z_lo:z0 - oz + z_hi, gy, gx]
                    mask = vals != 0
                    if mask.any():
                        self.set_leaf((x0, y0, z0), vals.reshape(512), mask.reshape(512), inactive_policy)

    # ---- encoding ----
    def _payload(self, values: np.ndarray) -> bytes:
        raw = values.astype(np.float16).tobytes() if self.half else values.astype("<f4").tobytes()
        typesize = 2 if self.half else 4
        if self.compression & COMPRESS_BLOSC:
            if len(raw) == 0 or self.blosc_mode == "raw":
                return struct.pack("<q", -len(raw)) + raw
            frame = _blosc_frame(raw, typesize, self.blosc_mode)
            return struct.pack("<q", len(frame)) + frame
        if self.compression & COMPRESS_ZIP:
            z = zlib.compress(raw)
            if len(z) >= len(raw):  # io::zipToStream stores incompressible buffers raw, size negated
                return struct.pack("<q", -len(raw)) + raw
            return struct.pack("<q", len(z)) + z
        return raw

    def _values(self, values, mask, metadata=None, child_mask=None) -> bytes:
        """io::writeCompressedValues.  `metadata` forces one of the seven inactive-value encodings (the caller must have prepared values
        that fit it); None picks like MaskCompress does."""
        values = np.asarray(values, np.float32)
        bg = self.background
        inactive = ~mask if child_mask is None else (~mask & ~child_mask)
        out = b""
        if not (self.compression & COMPRESS_ACTIVE_MASK):
            md = NO_MASK_AND_ALL_VALS
            return struct.pack("<b", md) + self._payload(values)
        uniq = np.unique(values[inactive])
        if metadata is None:
            if len(uniq) == 0 or (len(uniq) == 1 and uniq[0] == bg):
                metadata = NO_MASK_OR_INACTIVE_VALS
            elif len(uniq) == 1 and uniq[0] == -bg:
                metadata = NO_MASK_AND_MINUS_BG
            elif len(uniq) == 1:
                metadata = NO_MASK_AND_ONE_INACTIVE_VAL
            elif len(uniq) == 2 and set(uniq.tolist()) == {float(bg), float(-bg)}:
                metadata = MASK_AND_NO_INACTIVE_VALS
            elif len(uniq) == 2 and float(bg) in uniq.tolist():
                metadata = MASK_AND_ONE_INACTIVE_VAL
            elif len(uniq) == 2:
                metadata = MASK_AND_TWO_INACTIVE_VALS
            else:
                metadata = NO_MASK_AND_ALL_VALS
        out += struct.pack("<b", metadata)
        if metadata == NO_MASK_AND_ALL_VALS:
            return out + self._payload(values)
        selection = None
        if metadata == NO_MASK_AND_ONE_INACTIVE_VAL:
            out += struct.pack("<f", uniq[0])
        elif metadata == MASK_AND_NO_INACTIVE_VALS:  # inactiveVal0 = -background, inactiveVal1 = background (selected by the mask)
            selection = inactive & (values == bg)
        elif metadata == MASK_AND_ONE_INACTIVE_VAL:
            other = [u for u in uniq.tolist() if u != float(bg)][0]
            out += struct.pack("<f", other)
            selection = inactive & (values == bg)
        elif metadata == MASK_AND_TWO_INACTIVE_VALS:
            v0, v1 = uniq.tolist()
            out += struct.pack("<ff", v0, v1)
            selection = inactive & (values == np.float32(v1))
        if selection is not None:
            out += _mask_words(selection)
        return out + self._payload(values[mask])

    def _internal(self, level, origin, topology: bool) -> bytes:
        """InternalNode<5> (level 2, children 128^3) or InternalNode<4> (level 1, children = leaves 8^3)."""
        log2dim = 5 if level == 2 else 4
        child_dim = 128 if level == 2 else 8
        n = 1 << (3 * log2dim)
        child_mask = np.zeros(n, bool)
        value_mask = np.zeros(n, bool)
        values = np.full(n, self.background, np.float32)
        children = []
        dim = 1 << log2dim
        for off in range(n):
            lx, ly, lz = off >> (2 * log2dim), (off >> log2dim) & (dim - 1), off & (dim - 1)
            co = (origin[0] + lx * child_dim, origin[1] + ly * child_dim, origin[2] + lz * child_dim)
            has_child = self._has_content(level - 1, co)
            tile = self.tiles.get((level, co))
            if has_child:
                child_mask[off] = True
                children.append(co)
            elif tile is not None:
                values[off], value_mask[off] = tile
        out = b""
        if topology:
            out += _mask_words(child_mask) + _mask_words(value_mask) + self._values(values, value_mask, None, child_mask)
        for co in children:
            if level == 2:
                out += self._internal(1, co, topology)
            else:
                vals, mask, md = self.leaves[co]
                if topology:
                    out += _mask_words(mask)
                else:
                    out += _mask_words(mask) + self._values(vals, mask, md)
        return out

    def _has_content(self, level, origin) -> bool:
        """Does the node of `level` (0 = leaf) at `origin` exist, i.e. hold a leaf or a tile below it?"""
        dim = {0: 8, 1: 128, 2: 4096}[level]
        if level == 0:
            return origin in self.leaves
        inside = lambda o: all(origin[a] <= o[a] < origin[a] + dim for a in range(3))  # noqa: E731
        return any(inside(o) for o in self.leaves) or any(lv <= level and inside(o) for (lv, o) in self.tiles)

    def tobytes(self, grid_name="density") -> bytes:
        roots = sorted({tuple((o[a] // 4096) * 4096 for a in range(3)) for o in list(self.leaves) + [o for (lv, o) in self.tiles if lv < 3]})
        root_tiles = sorted((o, t) for (lv, o), t in self.tiles.items() if lv == 3)
        head = struct.pack("<qIIIb", 0x56444220, self.version, 5, 0, 0) + b"0" * 36
        head += _meta_map([("creator", "string", b"tests/vdb_writer.py")])
        head += struct.pack("<i", 1)
        grid = struct.pack("<I", self.compression)
        grid += _meta_map([("class", "string", b"fog volume"), ("file_compression", "string", b"n/a")])
        grid += _string("UniformScaleMap") + np.array([0.1] * 3 + [0.1] * 3 + [10.0] * 3 + [100.0] * 3 + [5.0] * 3, "<f8").tobytes()
        grid += struct.pack("<I", 1) + struct.pack("<f", self.background) + struct.pack("<II", len(root_tiles), len(roots))
        for o, (value, active) in root_tiles:
            grid += struct.pack("<3if?", *o, value, active)
        for o in roots:  # std::map<Coord, ...> order: lexicographic in (x, y, z)
            grid += struct.pack("<3i", *o) + self._internal(2, o, True)
        for o in roots:
            grid += self._internal(2, o, False)
        type_name = "Tree_float_5_4_3" + ("_HalfFloat" if self.half else "")
        desc = _string(grid_name) + _string(type_name) + _string("")
        grid_pos = len(head) + len(desc) + 24
        desc += struct.pack("<qqq", grid_pos, grid_pos, grid_pos + len(grid))
        return head + desc + grid

    def dense_reference(self):
        """What Resources::loadVolumeBuffer sees: (dense [z][y][x] over the active box + 1 of accessor values, origin, max active)."""
        lo, hi = np.full(3, 2**40, np.int64), np.full(3, -(2**40), np.int64)
        mx = -np.inf
        for o, (vals, mask, _) in self.leaves.items():
            idx = np.nonzero(mask)[0]
            if len(idx):
                c = np.stack([idx >> 6, (idx >> 3) & 7, idx & 7], 1) + np.array(o)
                lo, hi = np.minimum(lo, c.min(0)), np.maximum(hi, c.max(0))
                mx = max(mx, float(vals[mask].max()))
        for (lv, o), (value, active) in self.tiles.items():
            if active:
                dim = {1: 8, 2: 128, 3: 4096}[lv]
                lo, hi = np.minimum(lo, o), np.maximum(hi, np.array(o) + dim - 1)
                mx = max(mx, float(value))
        lo, hi = lo - 1, hi + 1
        n = hi + 1 - lo
        dense = np.full((n[2], n[1], n[0]), self.background, np.float32)
        for (lv, o), (value, active) in sorted(self.tiles.items(), key=lambda kv: -kv[0][0]):
            dim = {1: 8, 2: 128, 3: 4096}[lv]
            a, b = np.maximum(np.array(o), lo) - lo, np.minimum(np.array(o) + dim - 1, hi) - lo
            if (a <= b).all():
                dense[a[2]:b[2] + 1, a[1]:b[1] + 1, a[0]:b[0] + 1] = value
        for o, (vals, mask, _) in self.leaves.items():
            v = vals.reshape(8, 8, 8)  # [x][y][z]
            for lx in range(8):
                for ly in range(8):
                    for lz in range(8):
                        p = np.array(o) + (lx, ly, lz)
                        if (p >= lo).all() and (p <= hi).all():
                            q = p - lo
                            dense[q[2], q[1], q[0]] = v[lx, ly, lz]
        return dense, lo, mx
