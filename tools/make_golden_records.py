#!/usr/bin/env python3
"""Generate tests/golden/records.json from the REFERENCE's own protobuf modules.

Imports /root/reference/DeepestScatter_Train/PythonProtocols/*_pb2.py (pure-python protobuf runtime) and
serialises a fixed, seeded set of messages.  Runs in the build container only; the JSON it writes is the
committed fixture the record-encoder tests compare against (the GPU box has no /root/reference).

Negative zero is deliberately absent from the vectors: the reference's C++ writer (protobuf 3.6.1,
`if (this->x() != 0)`, CppProtocols/Vector.pb.cc:286) omits -0.0 while newer python runtimes emit it.
"""
import json
import os
import struct
import sys
from pathlib import Path

os.environ.setdefault("PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION", "python")
REF = Path("/root/reference/DeepestScatter_Train")
sys.path.insert(0, str(REF / "PythonProtocols"))
sys.path.insert(0, str(REF))

import numpy as np  # noqa: E402
from PythonProtocols.DisneyDescriptor_pb2 import DisneyDescriptor  # noqa: E402
from PythonProtocols.Result_pb2 import Result  # noqa: E402
from PythonProtocols.ScatterSample_pb2 import ScatterSample  # noqa: E402
from PythonProtocols.SceneSetup_pb2 import SceneSetup  # noqa: E402

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "records.json"


def f32(x):
    return struct.unpack("<f", struct.pack("<f", float(x)))[0]


def main():
    rng = np.random.default_rng(566)
    out = {"scatter_sample": [], "result": [], "scene_setup": [], "disney_descriptor": []}

    pts = [((0.25, -0.5, 0.125), (0.0, 0.0, 1.0)), ((0.0, 0.0, 0.0), (0.0, 0.0, 0.0)), ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0))]
    for _ in range(13):
        pts.append((tuple(f32(v) for v in rng.uniform(-0.5, 0.5, 3)), tuple(f32(v) for v in rng.normal(size=3))))
    for p, d in pts:
        m = ScatterSample()
        m.point.x, m.point.y, m.point.z = p
        m.view_direction.x, m.view_direction.y, m.view_direction.z = d
        # the collectors always touch both sub-messages (ScatterSampleCollector.cpp:48-56)
        m.point.SetInParent()
        m.view_direction.SetInParent()
        out["scatter_sample"].append({"point": [f32(v) for v in p], "view_direction": [f32(v) for v in d], "hex": m.SerializeToString().hex()})

    for v, c in [(0.25, True), (0.0, True), (1234.5, True), (3.0e-5, False), (0.0, False)] + [(f32(rng.uniform(0, 5000)), True) for _ in range(6)]:
        m = Result()
        m.light_intensity = v
        m.is_converged = c
        out["result"].append({"light_intensity": f32(v), "is_converged": c, "hex": m.SerializeToString().hex()})

    scenes = [("a.vdb", 7000.0, (-0.03, -0.25, 0.8)), ("", 0.0, (0.0, 0.0, 0.0)), ("RoundClouds/cloud 07.vdb", 1234.5, (0.0, -1.0, 0.0))]
    for i in range(5):
        scenes.append((f"Clouds/c{i}.vdb", f32(np.exp(rng.uniform(np.log(1000), np.log(12000)))), tuple(f32(v) for v in rng.normal(size=3))))
    for path, size, l in scenes:
        m = SceneSetup()
        m.cloud_path = path
        m.cloud_size_m = size
        m.light_direction.x, m.light_direction.y, m.light_direction.z = l
        m.light_direction.SetInParent()
        out["scene_setup"].append({"cloud_path": path, "cloud_size_m": f32(size), "light_direction": [f32(v) for v in l], "hex": m.SerializeToString().hex()})

    for n in (2250, 0, 1, 127, 128, 300):
        grid = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        m = DisneyDescriptor()
        m.grid = grid
        out["disney_descriptor"].append({"grid_hex": grid.hex(), "hex": m.SerializeToString().hex()})

    OUT.parent.mkdir(parents=True, exist_ok=True)
    OUT.write_text(json.dumps(out, indent=1))
    print(f"wrote {OUT}: " + ", ".join(f"{k}={len(v)}" for k, v in out.items()))


if __name__ == "__main__":
    main()
