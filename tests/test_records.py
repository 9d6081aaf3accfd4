"""Record layout: protobuf wire bytes of the product's encoder vs golden vectors generated from the
reference's own PythonProtocols/*_pb2.py (tools/make_golden_records.py)."""
import json
from pathlib import Path

import pytest

GOLDEN = json.loads((Path(__file__).parent / "golden" / "records.json").read_text())


def test_scatter_sample_golden(built_library):
    ds = built_library
    for g in GOLDEN["scatter_sample"]:
        assert ds.record_scatter_sample(g["point"], g["view_direction"]).hex() == g["hex"]


def test_result_golden(built_library):
    ds = built_library
    for g in GOLDEN["result"]:
        assert ds.record_result(g["light_intensity"], g["is_converged"]).hex() == g["hex"]


def test_scene_setup_golden(built_library):
    ds = built_library
    for g in GOLDEN["scene_setup"]:
        assert ds.record_scene_setup(g["cloud_path"], g["cloud_size_m"], g["light_direction"]).hex() == g["hex"]


def test_disney_descriptor_golden(built_library):
    ds = built_library
    for g in GOLDEN["disney_descriptor"]:
        assert ds.record_disney_descriptor(bytes.fromhex(g["grid_hex"])).hex() == g["hex"]


def test_survey_known_answers(built_library):
    """SURVEY.md 8a known-answer vectors."""
    ds = built_library
    assert ds.record_scatter_sample((0.25, -0.5, 0.125), (0, 0, 1)).hex() == "120f0d0000803e15000000bf1d0000003e1a051d0000803f"
    assert ds.record_result(0.25, True).hex() == "0d0000803e1001"
    assert ds.record_result(0.0, True).hex() == "1001"
    assert ds.record_scene_setup("a.vdb", 7000, (-0.03, -0.25, 0.8)).hex() == "0a05612e7664621500c0da451a0f0d8fc2f5bc15000080be1dcdcc4c3f"
    d = ds.record_disney_descriptor(bytes(2250))
    assert len(d) == 2253 and d[:3].hex() == "0aca11"


def test_negative_zero_is_omitted_like_protobuf_3_6_1(built_library):
    """CppProtocols/Vector.pb.cc:286 `if (this->x() != 0)`: -0.0 is not written."""
    ds = built_library
    assert ds.record_scatter_sample((-0.0, 0.0, 0.0), (0.0, -0.0, 0.0)).hex() == "12001a00"
    assert ds.record_result(-0.0, False) == b""


def test_buffer_too_small_is_an_error(built_library):
    import ctypes as C

    lib = built_library.load()
    buf = (C.c_uint8 * 4)()
    f3 = C.c_float * 3
    assert lib.ds_record_scatter_sample(f3(1, 2, 3), f3(4, 5, 6), buf, 4) < 0
