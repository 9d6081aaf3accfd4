"""ds_write_exr: the progressive buffer as the reference's Camera::saveToDisk writes it (Camera.cpp:149-175)."""
import ctypes as C
import struct

import numpy as np

from exr_reader import read_exr


def test_exr_layout_and_round_trip(built_library, tmp_path):
    ds = built_library
    lib = ds.load()
    w, h = 37, 11
    rng = np.random.default_rng(4)
    rgba = rng.uniform(0, 50, (h, w, 4)).astype(np.float32)
    rgba[..., 3] = 1.0
    path = tmp_path / "frame.exr"
    assert lib.ds_write_exr(str(path).encode(), w, h, rgba.ctypes.data_as(C.c_void_p)) == 0
    e = read_exr(path)
    assert (e["width"], e["height"]) == (w, h) and e["channels"] == ["B", "G", "R"]
    assert e["line_order"] == 1 and e["chunk_order"] == list(range(h - 1, -1, -1))  # DECREASING_Y: last scanline first in the file
    assert e["size"] == e["data_end"]
    a = e["attrs"]
    assert a["displayWindow"] == a["dataWindow"] == ("box2i", struct.pack("<4i", 0, 0, w - 1, h - 1))
    assert a["pixelAspectRatio"] == ("float", struct.pack("<f", 1.0)) and a["screenWindowWidth"] == ("float", struct.pack("<f", 1.0))
    assert a["screenWindowCenter"] == ("v2f", struct.pack("<2f", 0.0, 0.0))
    # scanline y = buffer row y (Camera.cpp:165-171: yStride = +width), alpha dropped
    for i, name in enumerate("RGB"):
        assert np.array_equal(e["image"][name], rgba[..., i])
    assert lib.ds_write_exr(str(tmp_path / "no_dir" / "x.exr").encode(), w, h, rgba.ctypes.data_as(C.c_void_p)) < 0
    assert b"cannot write" in lib.ds_cloud_last_error()
