"""The oracle pinned to the reference's own source.

oracle/_ref/libds_ref.so is the reference's DataGen code -- the CUDA/*.cu device programs AND the host classes that drive them
(Scene, Sun, VDBCloud, CloudMaterial, Camera, PathTracingRenderer, the three collectors, Resources, Mie) -- compiled UNMODIFIED
from /root/reference by g++ against the OptiX emulation in oracle/ref_shim/.  Every test here runs the same seeded input
through oracle/ds_oracle.cpp and through that library and demands BIT-EQUAL results.

What the emulation itself defines (because the OptiX SDK is not vendored): texture fetch arithmetic, optix::Onb, the
optixu vector math, clock() -> stream id, expf/log/sin/cos -> include/ds_detmath.h.  Everything else on the path is the
reference's text.
"""
import numpy as np
import pytest

import oracle_lib as ol
import ref_lib as rl

pytestmark = pytest.mark.skipif(not rl.available(), reason="oracle/_ref is not built and /root/reference is not mounted")

SUN_FRONT = (-0.586, -0.766, -0.271)  # Tasks.cpp:56
SUN_SIDE = (-0.03, -0.25, 0.8)  # Tasks.cpp:58


def make_pair(n=32, kind=0, seed=1234, size_m=7000.0, sun=SUN_FRONT, step=1.0 / 64.0, grid=None):
    o = ol.Oracle()
    if grid is None:
        o.volume_synth(n, kind, seed)
    else:
        o.volume_upload(grid)
    o.scene_set(size_m, sun, sample_step=step)
    o.bake()
    r = rl.Reference()
    r.volume_upload(o.level(0))
    return o, r


def rays_towards_box(n, seed):
    rng = np.random.default_rng(seed)
    orig = rng.normal(size=(n, 3)).astype(np.float32)
    orig = (orig / np.linalg.norm(orig, axis=1, keepdims=True) * 2.0).astype(np.float32)
    tgt = ((rng.random((n, 3)) - 0.5) * 0.7).astype(np.float32)
    d = tgt - orig
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    # a few rays that start inside the box (t = minimalRayDistance branch of cloudBBox.cu:29-35) and a few that miss it
    orig[: n // 10] = tgt[: n // 10]
    d[-n // 10 :] = -d[-n // 10 :]
    val0 = rng.integers(0, 2**24, n).astype(np.uint32)
    stream = rng.integers(1, 5000, n).astype(np.uint32)
    return orig, d, val0, stream


@pytest.mark.parametrize("shape", [(32, 32, 32), (12, 20, 28), (1, 5, 9)])
def test_mip_chain_is_resources_generate_mipmaps(shape):
    rng = np.random.default_rng(3)
    g = rng.integers(0, 256, shape, dtype=np.uint8)
    o = ol.Oracle()
    o.volume_upload(g)
    r = rl.Reference()
    r.volume_upload(g)
    assert o.level_count() == r.level_count()
    for l in range(o.level_count()):
        assert o.level_dims(l) == r.level_dims(l)
        assert np.array_equal(o.level(l), r.level(l)), f"mip level {l}"


def test_scene_variables_mie_samplers_and_bake():
    for shape_kind, size_m, sun in [((32, 0), 7000.0, SUN_FRONT), ((24, 1), 1200.0, SUN_SIDE), ((20, 0), 12000.0, (0.995, -0.0998, 0.0))]:
        o, r = make_pair(n=shape_kind[0], kind=shape_kind[1], size_m=size_m, sun=sun, step=1.0 / 96.0)
        r.scene_init(size_m, sun, 1.0 / 96.0, rl.MODE_ALL, 8, 8)
        do, dr = o.derived(), r.derived()
        for k in do:
            assert np.array_equal(np.float32(do[k]), np.float32(dr[k])), (k, do[k], dr[k])
        a, b, c = (np.empty(4096, np.float32) for _ in range(3))
        o.L.orc_get_mie(o.h, a, b, c)
        ra, rb, rc = r.mie()  # Mie.cpp:8206-8297 through the emulated buffers
        assert np.array_equal(a, ra) and np.array_equal(b, rb) and np.array_equal(c, rc)
        assert np.array_equal(o.inscatter(), r.inscatter())  # inScatter.cu:40-66
        assert o.inscatter().min() < 255 and o.inscatter().max() == 255


def test_non_cubic_bake_and_the_skip_variant():
    rng = np.random.default_rng(11)
    g = np.zeros((14, 22, 30), np.uint8)
    g[2:12, 3:19, 4:26] = rng.integers(0, 256, (10, 16, 22), dtype=np.uint8)
    o, r = make_pair(grid=g, size_m=2500.0, sun=SUN_SIDE, step=1.0 / 48.0)
    r.scene_init(2500.0, SUN_SIDE, 1.0 / 48.0, rl.MODE_ALL, 8, 8)
    assert np.array_equal(o.inscatter(), r.inscatter())
    o.bake(skip_empty=True)  # the oracle's empty-cell skipping bake must not change a byte either
    assert np.array_equal(o.inscatter(), r.inscatter())


@pytest.mark.parametrize("mode", [rl.MODE_ALL, rl.MODE_MULTI, rl.MODE_SINGLE])
def test_estimators_bit_equal(mode):
    o, r = make_pair(n=32, size_m=7000.0, step=1.0 / 64.0)
    r.scene_init(7000.0, SUN_FRONT, 1.0 / 64.0, mode, 8, 8)
    orig, d, val0, stream = rays_towards_box(600, 7 + mode)
    o.counters_reset()
    r.counters_reset()
    a = o.trace_paths(mode, orig, d, val0, stream)
    b = r.trace_paths(orig, d, val0, stream)
    assert np.array_equal(a, b)
    assert (a[:, 0] > 0).sum() > 100
    # same number of rtTrace calls, sun-transmittance taps (scatter events) and density taps (march steps)
    assert o.counters() == r.counters()


def test_estimators_bit_equal_thick_cloud_reaches_the_depth_cap():
    """12 km cloud, grazing sun: long paths, some end at MAX_DEPTH = 2000 (cloudRadianceMaterials.cu:4,31)."""
    sun = (0.995, -0.0998, 0.0)
    g = np.zeros((20, 20, 20), np.uint8)
    g[1:19, 1:19, 1:19] = 255
    o, r = make_pair(grid=g, size_m=12000.0, sun=sun, step=1.0 / 32.0)
    r.scene_init(12000.0, sun, 1.0 / 32.0, rl.MODE_ALL, 8, 8)
    orig, d, val0, stream = rays_towards_box(120, 3)
    o.counters_reset()
    a = o.trace_paths(rl.MODE_ALL, orig, d, val0, stream)
    b = r.trace_paths(orig, d, val0, stream)
    assert np.array_equal(a, b)
    c = o.counters()
    assert c["events"] / c["paths"] > 100


def test_camera_frame_progressive_and_tonemap():
    w, h = 32, 16
    o, r = make_pair(n=32, size_m=7000.0, step=1.0 / 64.0)
    r.scene_init(7000.0, SUN_FRONT, 1.0 / 64.0, rl.MODE_ALL, w, h)
    # Camera::updatePosition (Camera.cpp:100-134) passes the eye through frame * I * frame^-1, whose round-off (the SDK's
    # Matrix4x4::inverse, emulated with a cofactor formula) moves it by an ulp; everything else is calculateCameraVariables
    cam_r, cam_o = r.camera(), ol.camera_look_at(aspect=w / h)
    assert np.abs(cam_r - cam_o).max() <= 3e-7
    assert np.array_equal(o.render_frame(cam_r, w, h, rl.MODE_ALL, 3), r.render_frame_result(3))  # pathTracingCamera.cu
    p_o = v_o = None
    for _ in range(2):  # Camera::update twice: 2 x 10 subframes, Welford (progressive.cu:17-27), then the Reinhard passes
        nsub = r.camera_update(1)
        p_o, v_o = o.render_accumulate(r.camera(), w, h, rl.MODE_ALL, nsub - 9, 10, p_o, v_o)
    assert nsub == 20
    p_r, v_r, s_r = r.frame()
    assert np.array_equal(p_o, p_r) and np.array_equal(v_o, v_r)
    s_o, avg_o = ol.tonemap(p_o, 0.4)
    assert avg_o == r.average_luminance()
    assert np.array_equal(s_o, s_r)
    # the empty background is white in the reference (0/0 -> NaN -> clamp -> 1), the cloud is not
    assert (s_r[p_r[..., 0] == 0][:, :3] == 255).all() and (s_r[p_r[..., 0] > 0][:, :3] < 255).any()
    # forked rows give the same frame as the single-process launch
    assert np.array_equal(r.render_frame_result(3, procs=3), r.render_frame_result(3))


def test_exr_rows_are_written_top_down():
    """Camera::saveToDisk (Camera.cpp:149-175) runs every 40 subframes with DECREASING_Y: file row 0 is the TOP image row."""
    w, h = 16, 8
    o, r = make_pair(n=16, size_m=3000.0, step=1.0 / 16.0)
    r.scene_init(3000.0, SUN_FRONT, 1.0 / 16.0, rl.MODE_SINGLE, w, h)
    assert r.camera_update(4) == 40
    rgb, decreasing = r.last_exr()
    p_r, _, _ = r.frame()
    assert decreasing and np.array_equal(rgb, p_r[::-1, :, :3])


def test_convergence_test_of_the_camera():
    w, h = 16, 8
    o, r = make_pair(n=16, size_m=3000.0, step=1.0 / 16.0)
    r.scene_init(3000.0, SUN_FRONT, 1.0 / 16.0, rl.MODE_SINGLE, w, h)
    assert r.camera_update(10) == 100
    p_r, v_r, _ = r.frame()
    # Camera.cpp:232-268: converged when fewer than 500 pixels remain -- always true for a 128-pixel frame after 100 subframes
    left = o.L.orc_unconverged_pixels(p_r.reshape(-1), v_r.reshape(-1), w * h, 100)
    assert r.is_converged() == (left < 500)


def test_generated_points_bit_equal():
    o, r = make_pair(n=32, size_m=3000.0, step=1.0 / 64.0)
    r.scene_init(3000.0, SUN_FRONT, 1.0 / 64.0, rl.MODE_MULTI, 8, 8, path_tracer=False, collector=rl.COLLECT_SAMPLES, batch_size=96)
    p_r, d_r = r.generate_points(stream=5)
    p_o, d_o = o.generate_points(0, 96, 5)
    assert np.array_equal(p_r, p_o) and np.array_equal(d_r, d_o)
    assert np.isfinite(p_r).all() and np.abs(p_r).max() <= 0.51


@pytest.mark.parametrize("size_m", [1000.0, 7000.0, 12000.0])
def test_descriptor_bytes_bit_equal(size_m):
    o, r = make_pair(n=32, size_m=size_m, step=1.0 / 64.0)
    p, d = o.generate_points(0, 24, 1)
    r.put_samples(p, d, 2048)
    r.scene_init(size_m, SUN_FRONT, 1.0 / 64.0, rl.MODE_MULTI, 8, 8, path_tracer=False, collector=rl.COLLECT_DESCRIPTORS, batch_start=2048,
                 batch_size=24)
    got = r.descriptors(2048, 24)  # the DisneyDescriptor records DisneyDescriptorCollector wrote
    want = o.descriptors(p, d).reshape(24, 2250)
    assert np.array_equal(got, want)
    assert 0 < want.mean() < 255


def test_network_input_pass_bit_equal():
    w, h = 32, 16
    o, r = make_pair(n=32, size_m=3000.0, step=1.0 / 64.0)
    r.scene_init(3000.0, SUN_FRONT, 1.0 / 64.0, rl.MODE_ALL, w, h)
    cam = r.camera()
    inp_r, info_r = r.network_input(4, 2, 16, 8, 9)  # disneyCamera.cu + disneyDescriptorMaterial.cu
    inp_o, info_o = o.network_input(cam, w, h, (4, 2, 16, 8), 9)
    assert np.array_equal(info_r, info_o)
    assert np.array_equal(inp_r, inp_o)
    assert info_o[..., 4].sum() > 20


def test_radiance_collector_schedule_bit_equal():
    """RadianceCollector.cpp:73-192 (the reference's host code, unmodified) driving estimateEmission: two updates of 100 launches
    x 20480 threads, merge of the repeats (PointRadianceTask.h:54-68), convergence rule, reschedule, Result records."""
    o, r = make_pair(n=16, size_m=600.0, step=1.0 / 16.0)
    p, d = o.generate_points(0, 5, 5)
    r.put_samples(p, d, 0)
    r.scene_init(600.0, SUN_FRONT, 1.0 / 16.0, rl.MODE_MULTI, 8, 8, path_tracer=False, collector=rl.COLLECT_RADIANCE, batch_size=5)
    for updates in (1, 2):
        done_r, tasks_r, conv_r, _, recorded = r.radiance_update(1)
        tasks_o, conv_o, done_o, ups = o.point_radiance(p, d, 20480, 100, updates)
        assert tasks_r.tobytes() == tasks_o.tobytes()
        assert done_r == done_o and np.array_equal(conv_r.astype(bool), conv_o)
    assert done_r == 5 and recorded == 5
    # a sample that needed the second update ran with a different repeat count: the reschedule path is covered
    assert len(set(tasks_r["experimentCount"].tolist())) > 1
    values, flags = r.results(0, 5)
    assert np.array_equal(values, tasks_o["radiance"]) and flags.all()


def test_importer_crop_quantise_mips_is_load_volume_buffer(built_library):
    """Resources::loadVolumeBuffer (Resources.cpp:68-155) on a dense float grid placed at a non-zero index-space origin:
    active box + 1, value / max * 255 in double, mip chain -- against the product's host importer and the oracle's quantiser."""
    ds = built_library
    rng = np.random.default_rng(9)
    dense = np.zeros((20, 24, 28), np.float32)
    dense[3:15, 5:20, 2:27] = rng.uniform(0, 2.5, (12, 15, 25)).astype(np.float32)
    dense[3, 5, 2] = dense[14, 19, 26] = 2.5
    r = rl.Reference()
    r.volume_import(dense, origin=(-7, 3, 100))
    cropped, mx = ds.cloud_crop_active(dense)
    assert r.level_dims(0) == (cropped.shape[2], cropped.shape[1], cropped.shape[0])
    assert np.array_equal(r.float_size(), np.float32(cropped.shape[::-1]))
    q = np.empty(cropped.size, np.uint8)
    ol.lib().orc_quantize_float_grid(np.ascontiguousarray(cropped.reshape(-1)), cropped.size, float(mx), q)
    assert np.array_equal(r.level(0), q.reshape(cropped.shape))
    o = ol.Oracle()
    o.volume_upload(q.reshape(cropped.shape))
    assert o.level_count() == r.level_count()
    for l in range(o.level_count()):
        assert np.array_equal(o.level(l), r.level(l))


def test_light_direction_is_normalised_twice():
    """installSceneSetup normalises the SceneSetup's light direction (installers.cpp:73-77) and DirectionalLight's constructor
    normalises it again (SceneDescription.h:17-19); in fp32 the second pass changes the last bit of about a third of all
    directions.  orc_scene_set / ds_scene_set are the constructor; the driver (host/DataGen.hpp makeSceneDescription) is the
    installer.  Fed the installer's output, the oracle must hold the reference's lightDirection bit for bit."""
    rng = np.random.default_rng(1)
    g = np.zeros((8, 8, 8), np.uint8)
    g[2:6, 2:6, 2:6] = 200
    o = ol.Oracle()
    o.volume_upload(g)
    r = rl.Reference()
    r.volume_upload(g)
    changed = 0
    for v in rng.normal(size=(40, 3)).astype(np.float32):
        inv = np.float32(1.0) / np.sqrt(np.float32(np.float32(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]))
        once = (v * inv).astype(np.float32)
        o.scene_set(5000.0, once, sample_step=0.25)
        r.scene_init(5000.0, v, 0.25, rl.MODE_ALL, 4, 4, path_tracer=False, collector=rl.COLLECT_SAMPLES, batch_size=1)
        assert np.array_equal(o.derived()["light"], r.derived()["light"])
        changed += int(not np.array_equal(o.derived()["light"], once))
    assert changed > 3
