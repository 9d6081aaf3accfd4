"""Minimal reader of uncompressed single-part scanline OpenEXR files (test helper; the published file layout)."""
import struct

import numpy as np


def read_exr(path):
    raw = open(path, "rb").read()
    assert raw[:4] == b"\x76\x2f\x31\x01", "not an OpenEXR file"
    version, = struct.unpack_from("<I", raw, 4)
    assert version & 0xFF == 2 and version >> 8 == 0, "single-part scanline, no long names / deep data"
    pos = 8
    attrs = {}
    order = []
    while raw[pos] != 0:
        e = raw.index(b"\0", pos)
        name = raw[pos:e].decode()
        pos = e + 1
        e = raw.index(b"\0", pos)
        typ = raw[pos:e].decode()
        pos = e + 1
        size, = struct.unpack_from("<i", raw, pos)
        pos += 4
        attrs[name] = (typ, raw[pos:pos + size])
        order.append(name)
        pos += size
    pos += 1
    assert order == sorted(order)
    channels = []
    c = attrs["channels"][1]
    i = 0
    while c[i] != 0:
        e = c.index(b"\0", i)
        name = c[i:e].decode()
        ptype, plinear, xs, ys = struct.unpack_from("<iB3xii", c, e + 1)
        channels.append((name, ptype, xs, ys))
        i = e + 1 + 16
    assert attrs["compression"] == ("compression", b"\0")
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    line_order = attrs["lineOrder"][1][0]
    table = struct.unpack_from(f"<{h}Q", raw, pos)
    pos += 8 * h
    img = {name: np.zeros((h, w), np.float32) for name, *_ in channels}
    for y in range(h):
        off = table[y]
        yy, nbytes = struct.unpack_from("<ii", raw, off)
        assert yy == y0 + y and nbytes == 4 * w * len(channels)
        off += 8
        for name, ptype, xs, ys in channels:
            assert ptype == 2 and xs == ys == 1
            img[name][y] = np.frombuffer(raw, np.float32, w, off)
            off += 4 * w
    file_order = sorted(range(h), key=lambda y: table[y])
    return dict(attrs=attrs, channels=[c[0] for c in channels], width=w, height=h, line_order=line_order, image=img,
                chunk_order=file_order, size=len(raw), data_end=max(table) + 8 + 4 * w * len(channels))
