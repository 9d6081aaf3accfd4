"""Property test of the record encoders against the REFERENCE's own protobuf modules (DeepestScatter_Train/PythonProtocols/*_pb2.py),
imported from /root/reference: only where that tree is mounted -- the build container, where the CPU suite runs; skipped elsewhere (the
GPU box never sees the reference, and the committed golden vectors of test_records.py cover it there).

Negative zero is excluded: the reference's C++ writer (protobuf 3.6.1, `if (this->x() != 0)`) omits -0.0, which the product follows
(test_negative_zero_is_omitted_like_protobuf_3_6_1), while newer python runtimes emit it."""
import os
import struct
import sys
from pathlib import Path

import pytest

REF = Path("/root/reference/DeepestScatter_Train")
if not (REF / "PythonProtocols").is_dir():
    pytest.skip("the reference tree is not mounted here", allow_module_level=True)
hypothesis = pytest.importorskip("hypothesis")
os.environ.setdefault("PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION", "python")
sys.path.insert(0, str(REF / "PythonProtocols"))
sys.path.insert(0, str(REF))

from hypothesis import given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

try:
    from PythonProtocols.DisneyDescriptor_pb2 import DisneyDescriptor  # noqa: E402
    from PythonProtocols.Result_pb2 import Result  # noqa: E402
    from PythonProtocols.ScatterSample_pb2 import ScatterSample  # noqa: E402
    from PythonProtocols.SceneSetup_pb2 import SceneSetup  # noqa: E402
except Exception as exc:  # protobuf runtime missing or incompatible
    pytest.skip(f"cannot import the reference's protobuf modules: {exc}", allow_module_level=True)


def f32(x):
    return struct.unpack("<f", struct.pack("<f", x))[0]


FLOATS = st.floats(width=32, allow_nan=False, allow_infinity=True).filter(lambda v: not (v == 0 and str(v).startswith("-"))).map(f32)
VEC = st.tuples(FLOATS, FLOATS, FLOATS)


@settings(max_examples=200, deadline=None)
@given(point=VEC, view=VEC)
def test_scatter_sample_bytes(built_library, point, view):
    m = ScatterSample()
    m.point.x, m.point.y, m.point.z = point
    m.view_direction.x, m.view_direction.y, m.view_direction.z = view
    m.point.SetInParent()  # the collectors always touch both sub-messages (ScatterSampleCollector.cpp:48-56)
    m.view_direction.SetInParent()
    assert built_library.record_scatter_sample(point, view) == m.SerializeToString()


@settings(max_examples=200, deadline=None)
@given(v=FLOATS, c=st.booleans())
def test_result_bytes(built_library, v, c):
    m = Result()
    m.light_intensity = v
    m.is_converged = c
    assert built_library.record_result(v, c) == m.SerializeToString()


@settings(max_examples=150, deadline=None)
@given(path=st.text(max_size=300), size=FLOATS, light=VEC)
def test_scene_setup_bytes(built_library, path, size, light):
    if "\x00" in path or any(0xD800 <= ord(ch) <= 0xDFFF for ch in path):
        return  # the C ABI takes a NUL-terminated UTF-8 path
    m = SceneSetup()
    m.cloud_path = path
    m.cloud_size_m = size
    m.light_direction.x, m.light_direction.y, m.light_direction.z = light
    m.light_direction.SetInParent()
    assert built_library.record_scene_setup(path, size, light) == m.SerializeToString()


@settings(max_examples=100, deadline=None)
@given(grid=st.binary(max_size=5000))
def test_disney_descriptor_bytes(built_library, grid):
    m = DisneyDescriptor()
    m.grid = grid
    assert built_library.record_disney_descriptor(grid) == m.SerializeToString()
