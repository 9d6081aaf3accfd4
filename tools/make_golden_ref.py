#!/usr/bin/env python
"""Golden vectors from the reference's own source: runs oracle/_ref/libds_ref.so (the reference's DataGen code compiled
unmodified against the OptiX emulation, oracle/ref_shim/) on small seeded scenes and writes tests/golden/ref_path.npz.

Run where /root/reference is mounted:   python tools/make_golden_ref.py
tests/test_golden_ref.py holds the oracle (CPU) and the EXACT CUDA flavour (GPU) to these vectors bit for bit, so the pin
survives on machines where neither /root/reference nor oracle/_ref exists.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracle_lib as ol  # noqa: E402  (only for the synthetic grid generator, include/ds_synth.h)
import ref_lib as rl  # noqa: E402

OUT = ROOT / "tests" / "golden" / "ref_path.npz"
SUN_FRONT = (-0.586, -0.766, -0.271)
SUN_GRAZING = (0.995, -0.0998, 0.0)


def synth_grid(n, kind, seed):
    o = ol.Oracle()
    o.volume_synth(n, kind, seed)
    return o.level(0)


def rays(n, seed):
    rng = np.random.default_rng(seed)
    orig = rng.normal(size=(n, 3)).astype(np.float32)
    orig = (orig / np.linalg.norm(orig, axis=1, keepdims=True) * 2.0).astype(np.float32)
    tgt = ((rng.random((n, 3)) - 0.5) * 0.7).astype(np.float32)
    d = tgt - orig
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    orig[: n // 10] = tgt[: n // 10]
    val0 = rng.integers(0, 2**24, n).astype(np.uint32)
    stream = rng.integers(1, 5000, n).astype(np.uint32)
    return orig, d, val0, stream


def main():
    g = {}
    # scene A: 24^3 synthetic cumulus, 7 km, sun "Front", coarse step (cheap for every consumer of the file)
    n, size_m, step = 24, 7000.0, 1.0 / 48.0
    grid = synth_grid(n, 0, 1234)
    g["A_grid_n"], g["A_size_m"], g["A_step"], g["A_sun"] = np.int32(n), np.float32(size_m), np.float32(step), np.float32(SUN_FRONT)
    r = rl.Reference()
    r.volume_upload(grid)
    r.scene_init(size_m, SUN_FRONT, step, rl.MODE_ALL, 24, 12)
    g["A_inscatter"] = r.inscatter()
    d = r.derived()
    g["A_derived"] = np.concatenate([d["bbox"], d["texture_scale"], [d["density_multiplier"], d["voxel_m"], d["voxel_free_path"]], d["light"]]).astype(np.float32)
    orig, dirs, val0, stream = rays(160, 21)
    g["A_ray_orig"], g["A_ray_dir"], g["A_ray_val0"], g["A_ray_stream"] = orig, dirs, val0, stream
    for mode in (rl.MODE_ALL, rl.MODE_MULTI, rl.MODE_SINGLE):
        r.scene_init(size_m, SUN_FRONT, step, mode, 24, 12)
        g[f"A_radiance_mode{mode}"] = r.trace_paths(orig, dirs, val0, stream)
    # progressive frame: Camera::update once (10 subframes) with the all-order estimator
    r.scene_init(size_m, SUN_FRONT, step, rl.MODE_ALL, 24, 12)
    assert r.camera_update(1) == 10
    g["A_camera"] = r.camera()
    p, v, s = r.frame()
    g["A_progressive"], g["A_variance"], g["A_screen"] = p, v, s
    g["A_avg_luminance"] = np.float32(r.average_luminance())
    # dataset pass: points -> descriptors -> one RadianceCollector update
    r.scene_init(size_m, SUN_FRONT, step, rl.MODE_MULTI, 8, 8, path_tracer=False, collector=rl.COLLECT_SAMPLES, batch_size=16)
    pts, vd = r.generate_points(stream=3)
    g["A_points"], g["A_view_dirs"] = pts, vd
    r.put_samples(pts, vd, 0)
    r.scene_init(size_m, SUN_FRONT, step, rl.MODE_MULTI, 8, 8, path_tracer=False, collector=rl.COLLECT_DESCRIPTORS, batch_size=16)
    g["A_descriptors"] = r.descriptors(0, 16)

    # scene B: thin 16^3 cloud for the collector (2 M paths per update in the reference's schedule)
    n, size_m, step = 16, 600.0, 1.0 / 16.0
    grid = synth_grid(n, 0, 1234)
    g["B_grid_n"], g["B_size_m"], g["B_step"] = np.int32(n), np.float32(size_m), np.float32(step)
    r = rl.Reference()
    r.volume_upload(grid)
    r.scene_init(size_m, SUN_FRONT, step, rl.MODE_MULTI, 8, 8, path_tracer=False, collector=rl.COLLECT_SAMPLES, batch_size=5)
    pts, vd = r.generate_points(stream=5)
    g["B_points"], g["B_view_dirs"] = pts, vd
    r.put_samples(pts, vd, 0)
    r.scene_init(size_m, SUN_FRONT, step, rl.MODE_MULTI, 8, 8, path_tracer=False, collector=rl.COLLECT_RADIANCE, batch_size=5)
    done, tasks, conv, _, _ = r.radiance_update(1)
    g["B_tasks_after_1_update"] = tasks.view(np.uint8).reshape(5, 40)
    g["B_converged_after_1_update"] = conv

    # scene C: solid block, 12 km, grazing sun: long paths up to the MAX_DEPTH cap
    grid = np.zeros((20, 20, 20), np.uint8)
    grid[1:19, 1:19, 1:19] = 255
    r = rl.Reference()
    r.volume_upload(grid)
    r.scene_init(12000.0, SUN_GRAZING, 1.0 / 32.0, rl.MODE_ALL, 8, 8)
    orig, dirs, val0, stream = rays(48, 4)
    g["C_ray_orig"], g["C_ray_dir"], g["C_ray_val0"], g["C_ray_stream"] = orig, dirs, val0, stream
    g["C_radiance"] = r.trace_paths(orig, dirs, val0, stream)
    g["C_inscatter"] = r.inscatter()

    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, **g)
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes, {len(g)} arrays)")


if __name__ == "__main__":
    main()
