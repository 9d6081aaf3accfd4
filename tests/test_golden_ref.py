"""Golden vectors produced by the reference's own source (tests/golden/ref_path.npz, written by tools/make_golden_ref.py from
oracle/_ref/libds_ref.so = the reference's DataGen code compiled unmodified against the OptiX emulation in oracle/ref_shim/).

CPU: oracle/ds_oracle.cpp must reproduce them bit for bit.  GPU (-m gpu): so must the EXACT CUDA flavour, through the C ABI --
the CUDA path held directly to the reference's arithmetic, with no oracle in between.
"""
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as ol

GOLDEN = Path(__file__).resolve().parent / "golden" / "ref_path.npz"
SUN_GRAZING = (0.995, -0.0998, 0.0)


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLDEN))


def derived_vector(d):
    return np.concatenate([d["bbox"], d["texture_scale"], [d["density_multiplier"], d["voxel_m"], d["voxel_free_path"]], d["light"]]).astype(np.float32)


def oracle_scene_a(g):
    o = ol.Oracle()
    o.volume_synth(int(g["A_grid_n"]), 0, 1234)
    o.scene_set(float(g["A_size_m"]), g["A_sun"], sample_step=float(g["A_step"]))
    o.bake()
    return o


def block_grid():
    grid = np.zeros((20, 20, 20), np.uint8)
    grid[1:19, 1:19, 1:19] = 255
    return grid


# ------------------------------------------------------------------ CPU: the oracle against the reference's vectors
def test_oracle_scene_variables_and_bake(g):
    o = oracle_scene_a(g)
    assert np.array_equal(derived_vector(o.derived()), g["A_derived"])
    assert np.array_equal(o.inscatter(), g["A_inscatter"])


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_oracle_estimators(g, mode):
    o = oracle_scene_a(g)
    got = o.trace_paths(mode, g["A_ray_orig"], g["A_ray_dir"], g["A_ray_val0"], g["A_ray_stream"])
    assert np.array_equal(got, g[f"A_radiance_mode{mode}"])
    assert (got[:, 0] > 0).sum() > 20


def test_oracle_progressive_frame_and_tonemap(g):
    o = oracle_scene_a(g)
    p, v = o.render_accumulate(g["A_camera"], 24, 12, ol.MODE_ALL, 1, 10)
    assert np.array_equal(p, g["A_progressive"]) and np.array_equal(v, g["A_variance"])
    s, avg = ol.tonemap(p, 0.4)
    assert np.float32(avg) == g["A_avg_luminance"] and np.array_equal(s, g["A_screen"])


def test_oracle_points_and_descriptors(g):
    o = oracle_scene_a(g)
    p, d = o.generate_points(0, 16, 3)
    assert np.array_equal(p, g["A_points"]) and np.array_equal(d, g["A_view_dirs"])
    assert np.array_equal(o.descriptors(p, d).reshape(16, 2250), g["A_descriptors"])


def test_oracle_radiance_collector(g):
    o = ol.Oracle()
    o.volume_synth(int(g["B_grid_n"]), 0, 1234)
    o.scene_set(float(g["B_size_m"]), (-0.586, -0.766, -0.271), sample_step=float(g["B_step"]))
    o.bake()
    p, d = o.generate_points(0, 5, 5)
    assert np.array_equal(p, g["B_points"]) and np.array_equal(d, g["B_view_dirs"])
    tasks, conv, _, _ = o.point_radiance(p, d, 20480, 100, 1)
    assert np.array_equal(tasks.view(np.uint8).reshape(5, 40), g["B_tasks_after_1_update"])
    assert np.array_equal(conv, g["B_converged_after_1_update"].astype(bool))


def test_oracle_long_paths(g):
    o = ol.Oracle()
    o.volume_upload(block_grid())
    o.scene_set(12000.0, SUN_GRAZING, sample_step=1.0 / 32.0)
    o.bake()
    assert np.array_equal(o.inscatter(), g["C_inscatter"])
    assert np.array_equal(o.trace_paths(0, g["C_ray_orig"], g["C_ray_dir"], g["C_ray_val0"], g["C_ray_stream"]), g["C_radiance"])


# ------------------------------------------------------------------ GPU: the EXACT flavour against the reference's vectors
def cam_struct(ds, arr):
    from deepestscatter_b200._lib import DsCamera

    cam = DsCamera()
    for i, name in enumerate(("eye", "U", "V", "W")):
        for k in range(3):
            getattr(cam, name)[k] = float(arr[3 * i + k])
    return cam


@pytest.fixture(scope="module")
def ctx_scene_a(built_library, g):
    ds = built_library
    ctx = ds.Context(0)
    ctx.set_option("precision", ds.PRECISION_EXACT)
    ctx.volume_synth(int(g["A_grid_n"]), 0, 1234)
    ctx.scene_set(float(g["A_size_m"]), tuple(float(x) for x in g["A_sun"]), sample_step=float(g["A_step"]))
    ctx.bake()
    yield ctx
    ctx.close()


@pytest.mark.gpu
def test_gpu_exact_scene_variables_and_bake(ctx_scene_a, g):
    assert np.array_equal(derived_vector(ctx_scene_a.derived()), g["A_derived"])
    assert np.array_equal(ctx_scene_a.inscatter(), g["A_inscatter"])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_gpu_exact_estimators(ctx_scene_a, g, mode):
    got = ctx_scene_a.trace_paths(mode, g["A_ray_orig"], g["A_ray_dir"], g["A_ray_val0"], g["A_ray_stream"])
    assert np.array_equal(got, g[f"A_radiance_mode{mode}"])


@pytest.mark.gpu
def test_gpu_exact_progressive_frame_and_tonemap(built_library, ctx_scene_a, g):
    ds, ctx = built_library, ctx_scene_a
    ctx.frame_create(24, 12)
    ctx.render_subframes(cam_struct(ds, g["A_camera"]), ds.MODE_ALL_SCATTER, 1, 10)
    p, v = ctx.frame_download()
    assert np.array_equal(p, g["A_progressive"]) and np.array_equal(v, g["A_variance"])
    screen, avg = ctx.tonemap(0.4)
    assert np.float32(avg) == g["A_avg_luminance"]
    # powf is the one libm call on this surface: the CUDA and glibc results may differ by one count
    assert np.abs(screen.astype(np.int32) - g["A_screen"].astype(np.int32)).max() <= 1
    assert (screen[p[..., 0] == 0][:, :3] == 255).all()  # the reference's white background (reinhard.cu:69, 0/0 clamped to 1)


@pytest.mark.gpu
def test_gpu_exact_points_and_descriptors(ctx_scene_a, g):
    p, d = ctx_scene_a.generate_points(0, 16, 3)
    assert np.array_equal(p, g["A_points"]) and np.array_equal(d, g["A_view_dirs"])
    assert np.array_equal(ctx_scene_a.descriptors(p, d).reshape(16, 2250), g["A_descriptors"])


@pytest.mark.gpu
def test_gpu_exact_radiance_collector(built_library, g):
    ds = built_library
    with ds.Context(0) as ctx:
        ctx.set_option("precision", ds.PRECISION_EXACT)
        ctx.volume_synth(int(g["B_grid_n"]), 0, 1234)
        ctx.scene_set(float(g["B_size_m"]), (-0.586, -0.766, -0.271), sample_step=float(g["B_step"]))
        ctx.bake()
        tasks, conv, _, _ = ctx.point_radiance(g["B_points"], g["B_view_dirs"], 20480, 100, 1)
        assert np.array_equal(tasks.view(np.uint8).reshape(5, 40), g["B_tasks_after_1_update"])
        assert np.array_equal(conv, g["B_converged_after_1_update"].astype(bool))


@pytest.mark.gpu
def test_gpu_exact_long_paths(built_library, g):
    ds = built_library
    with ds.Context(0) as ctx:
        ctx.set_option("precision", ds.PRECISION_EXACT)
        ctx.volume_upload(block_grid())
        ctx.scene_set(12000.0, SUN_GRAZING, sample_step=1.0 / 32.0)
        ctx.bake()
        assert np.array_equal(ctx.inscatter(), g["C_inscatter"])
        assert np.array_equal(ctx.trace_paths(0, g["C_ray_orig"], g["C_ray_dir"], g["C_ray_val0"], g["C_ray_stream"]), g["C_radiance"])
