"""Pins the host oracle against everything the reference gives us (Mie data, integer algorithms restated
independently in numpy) and against analytic results (Beer-Lambert slab, phase-function sampling).
The reference has no tests or golden images: the estimator itself stays 'parity unpinned'."""
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as ol


# ---------------------------------------------------------------- independent numpy restatements

def np_tea4(val0, stream):
    v0 = np.uint32(val0)
    v1 = np.uint32(stream)
    s0 = np.uint32(0)
    with np.errstate(over="ignore"):
        for _ in range(4):
            s0 = np.uint32(s0 + np.uint32(0x9E3779B9))
            v0 = np.uint32(v0 + (np.uint32((np.uint32(v1 << np.uint32(4)) + np.uint32(0xA341316C))) ^ np.uint32(v1 + s0) ^ np.uint32((v1 >> np.uint32(5)) + np.uint32(0xC8013EA4))))
            v1 = np.uint32(v1 + (np.uint32((np.uint32(v0 << np.uint32(4)) + np.uint32(0xAD90777D))) ^ np.uint32(v0 + s0) ^ np.uint32((v0 >> np.uint32(5)) + np.uint32(0x7E95761E))))
    return int(v0)


def np_mips(level0):
    levels = [level0]
    nz, ny, nx = level0.shape
    m = max(nx, ny, nz)
    count = 1
    while m // 2:
        m //= 2
        count += 1
    for l in range(1, count):
        p = levels[-1].astype(np.uint16)
        cz, cy, cx = max(1, nz >> l), max(1, ny >> l), max(1, nx >> l)
        pad = np.zeros((2 * cz + 2, 2 * cy + 2, 2 * cx + 2), dtype=np.uint16)
        pz, py, px = p.shape
        pad[:pz, :py, :px] = p
        pad = pad[: 2 * cz, : 2 * cy, : 2 * cx]
        s = pad.reshape(cz, 2, cy, 2, cx, 2).sum(axis=(1, 3, 5))
        levels.append((s // 8).astype(np.uint8))
    return levels


def np_tex3d(vol, uvw):
    nz, ny, nx = vol.shape
    out = np.empty(len(uvw), dtype=np.float64)
    for i, (u, v, w) in enumerate(uvw.astype(np.float64)):
        x, y, z = u * nx - 0.5, v * ny - 0.5, w * nz - 0.5
        x0, y0, z0 = int(np.floor(x)), int(np.floor(y)), int(np.floor(z))
        tx, ty, tz = x - x0, y - y0, z - z0
        acc = 0.0
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    xi = min(max(x0 + dx, 0), nx - 1)
                    yi = min(max(y0 + dy, 0), ny - 1)
                    zi = min(max(z0 + dz, 0), nz - 1)
                    wgt = (tx if dx else 1 - tx) * (ty if dy else 1 - ty) * (tz if dz else 1 - tz)
                    acc += wgt * vol[zi, yi, xi]
        out[i] = acc / 255.0
    return out


# ---------------------------------------------------------------- Mie data (the one reference golden we have)

def test_mie_table_values_quoted_in_survey():
    mie, chopped = ol.load_mie_tables()
    assert mie.shape == (4096,) and chopped.shape == (4096,)
    assert abs(mie.astype(np.float64).mean() - 5.2588) < 1e-3
    assert abs(chopped.astype(np.float64).mean() - 0.52687) < 1e-4
    assert int(np.argmax(mie != chopped)) == 4081
    assert abs(float(mie[4095]) - 19086.0499712) < 2e-3
    assert abs(float(mie[0]) - 0.7136052853) < 1e-7
    assert abs(float(chopped[4095]) - 9.9666332937) < 1e-6


def test_mie_samplers_follow_mie_cpp():
    """Mie.cpp:8215-8226 (table / mean) and :8254-8265 (running sum of table / sum), float32 sequential."""
    o = ol.Oracle()
    mie, chopped = ol.load_mie_tables()
    a, b, c = o.mie_tables()

    def phase(t):
        acc = np.float32(0)
        for v in t:
            acc = np.float32(acc + v)
        return (t / np.float32(acc / np.float32(4096))).astype(np.float32)

    assert np.array_equal(a, phase(mie))
    assert np.array_equal(b, phase(chopped))
    s = np.float32(0)
    for v in chopped:
        s = np.float32(s + v)
    integ = np.float32(0)
    ref = np.empty(4096, dtype=np.float32)
    for i, v in enumerate(chopped):
        integ = np.float32(integ + np.float32(v / s))
        ref[i] = integ
    assert np.array_equal(c, ref)
    assert abs(float(c[-1]) - 1.0) < 1e-3 and np.all(np.diff(c) > 0)


# ---------------------------------------------------------------- integer work

def test_tea_and_lcg_match_independent_restatement():
    rng = np.random.default_rng(1)
    val0 = rng.integers(0, 2**32, 64, dtype=np.uint32)
    stream = rng.integers(0, 2**32, 64, dtype=np.uint32)
    seeds = np.empty(64, dtype=np.uint32)
    draws = np.empty(64 * 5, dtype=np.float32)
    ol.lib().orc_rng_probe(val0, stream, 64, 5, seeds, draws)
    for i in range(64):
        s = np_tea4(val0[i], stream[i])
        assert s == int(seeds[i])
        for d in range(5):
            s = (1664525 * s + 1013904223) & 0xFFFFFFFF
            assert draws[i * 5 + d] == np.float32((s & 0xFFFFFF) / 16777216.0)
    assert draws.min() >= 0.0 and draws.max() < 1.0


@pytest.mark.parametrize("shape", [(16, 16, 16), (13, 7, 21), (1, 5, 9), (33, 32, 31)])
def test_mip_chain_matches_numpy(shape):
    rng = np.random.default_rng(7)
    g = rng.integers(0, 256, shape, dtype=np.uint8)
    o = ol.Oracle()
    o.volume_upload(g, True)
    ref = np_mips(g)
    assert o.level_count() == len(ref)
    for l, r in enumerate(ref):
        assert np.array_equal(o.level(l), r), f"level {l}"
    assert o.level(len(ref) - 1).shape == (1, 1, 1)  # Resources.cpp:208


def test_quantize_float_grid_truncates_like_narrow_cast():
    rng = np.random.default_rng(3)
    v = rng.uniform(0, 2.5, 1000).astype(np.float32)
    v[:3] = [0.0, 2.5, 1.25]
    out = np.empty(1000, dtype=np.uint8)
    ol.lib().orc_quantize_float_grid(v, 1000, 2.5, out)
    assert np.array_equal(out, np.floor(v.astype(np.float64) / 2.5 * 255).astype(np.uint8))
    assert out[1] == 255 and out[0] == 0


def test_synthetic_grids_have_zero_border_and_are_deterministic():
    o = ol.Oracle()
    for kind in (0, 1, 2):
        o.volume_synth(32, kind, 1234, False)
        g = o.level(0)
        assert g[0].max() == 0 and g[-1].max() == 0 and g[:, 0].max() == 0 and g[:, :, -1].max() == 0
        assert g.max() > 100
        o.volume_synth(32, kind, 1234, False)
        assert np.array_equal(g, o.level(0))
    o.volume_synth(32, 2, 0, False)
    g = o.level(0)
    assert np.all(g[8:24, 1:-1, 1:-1] == 255) and np.all(g[:8] == 0) and np.all(g[24:] == 0)


# ---------------------------------------------------------------- texture semantics

def test_tex3d_matches_float64_trilinear():
    rng = np.random.default_rng(11)
    g = rng.integers(0, 256, (9, 12, 17), dtype=np.uint8)
    o = ol.Oracle()
    o.volume_upload(g, False)
    o.scene_set()
    d = o.derived()
    uvw = rng.uniform(-0.1, 1.1, (500, 3))
    pos = (uvw / d["texture_scale"]).astype(np.float32)
    got = o.sample_volume(pos)
    ref = np_tex3d(g, (pos * d["texture_scale"]).astype(np.float32))
    assert np.abs(got - ref).max() < 2e-6
    # texel centres reproduce the texel exactly
    cz, cy, cx = 4, 5, 6
    centre = np.array([[(cx + 0.5) / 17, (cy + 0.5) / 12, (cz + 0.5) / 9]]) / d["texture_scale"]
    assert o.sample_volume(centre)[0] == np.float32(g[cz, cy, cx]) * np.float32(1 / 255)


def test_tex1d_clamps_and_interpolates():
    o = ol.Oracle()
    _, _, cdf = o.mie_tables()
    got = o.sample_table(2, [-1.0, 0.0, 0.5 / 4096, 1.0 / 4096, 1.0, 2.0])
    assert got[0] == cdf[0] and got[1] == cdf[0] and got[2] == cdf[0]
    assert got[3] == np.float32(cdf[0] + np.float32(0.5) * (cdf[1] - cdf[0]))
    assert got[4] == cdf[-1] and got[5] == cdf[-1]


# ---------------------------------------------------------------- analytic checks

def test_bake_matches_beer_lambert_in_a_slab():
    """Homogeneous slab, light travelling +z: T(z) = exp(-mult * depth); CU/inScatter.cu:40-66."""
    n = 32
    o = ol.Oracle()
    o.volume_synth(n, 2, 0, False)
    o.scene_set(cloud_size_m=20.0, light_dir=(0, 0, 1))  # mult = 2
    o.bake()
    t = o.inscatter()[:, n // 2, n // 2].astype(np.float64) / 255.0
    z = np.arange(n) / n  # voxel-corner sample origin (inScatter.cu:46)
    depth = np.clip(z - 0.25, 0, 0.5)
    ref = np.exp(-2.0 * depth)
    assert np.abs(t - ref).max() < 0.03  # trilinear ramp at the slab faces + u8 truncation
    assert t[0] >= 254 / 255 and abs(t[-1] - np.exp(-1.0)) < 0.03


def test_skipping_bake_is_byte_identical(oracle_small):
    import time

    ref = oracle_small.inscatter().copy()
    oracle_small.bake(skip_empty=True)
    assert np.array_equal(oracle_small.inscatter(), ref)


def test_single_scatter_mean_matches_analytic_slab_integral():
    """View +z through the slab, light +z: E[L] = I * phase(-1) * ratio * (1 - exp(-2 sigma D)) / 2."""
    n = 32
    o = ol.Oracle()
    o.volume_synth(n, 2, 0, False)
    o.scene_set(cloud_size_m=20.0, light_dir=(0, 0, 1), light_intensity=1e6)
    o.bake()
    npaths = 40000
    origins = np.tile(np.array([[0.0, 0.0, -2.0]], dtype=np.float32), (npaths, 1))
    dirs = np.tile(np.array([[0.0, 0.0, 1.0]], dtype=np.float32), (npaths, 1))
    rad = o.trace_paths(ol.MODE_SINGLE, origins, dirs, np.arange(npaths, dtype=np.uint32) * 4096, np.ones(npaths, dtype=np.uint32))
    mie, _, _ = o.mie_tables()
    phase = float(mie[0])
    ratio = 5.334615707397461e-06
    ref = 1e6 * phase * ratio * (1 - np.exp(-2 * 2.0 * 0.5)) / 2
    mean = rad[:, 0].mean()
    sem = rad[:, 0].std() / np.sqrt(npaths)
    assert abs(mean - ref) < 4 * sem + 0.03 * ref
    assert np.array_equal(rad[:, 0], rad[:, 1]) and np.array_equal(rad[:, 0], rad[:, 2])  # white light


def test_direction_sampling_follows_the_chopped_mie_pdf():
    o = ol.Oracle()
    _, chopped = ol.load_mie_tables()
    n = 200000
    prev = np.tile(np.array([[0.3, -0.5, 0.8]], dtype=np.float32), (n, 1))
    prev /= np.linalg.norm(prev, axis=1, keepdims=True)
    d = o.new_directions(np.arange(n, dtype=np.uint32), np.full(n, 9, dtype=np.uint32), prev)
    assert np.abs(np.linalg.norm(d, axis=1) - 1).max() < 1e-5
    cos = np.clip((d * prev).sum(axis=1), -1, 1)
    hist, _ = np.histogram(cos, bins=32, range=(-1, 1))
    pdf = chopped.astype(np.float64).reshape(32, 128).sum(axis=1)
    pdf /= pdf.sum()
    expected = pdf * n
    big = expected > 200
    assert np.abs(hist[big] - expected[big]).max() / np.sqrt(expected[big]).max() < 6
    assert abs(cos.mean() - (np.linspace(-1, 1, 4096) * chopped).sum() / chopped.sum()) < 5e-3


def test_paths_that_miss_the_box_return_zero(oracle_small):
    o = oracle_small
    origins = np.array([[2.5, 3.0, 0.0], [0.0, 0.0, -3.0]], dtype=np.float32)
    dirs = np.array([[-1.0, 0.0, 0.0], [0.0, 0.0, -1.0]], dtype=np.float32)  # beside the box / pointing away
    for mode in (0, 1, 2):
        rad = o.trace_paths(mode, origins, dirs, [1, 2], [1, 1])
        assert np.all(rad == 0)


# ---------------------------------------------------------------- accumulation / tone map / camera / scheduler

def test_welford_update_matches_numpy_moments():
    rng = np.random.default_rng(5)
    x = rng.gamma(2.0, 3.0, (50, 6, 4)).astype(np.float32)
    prog = np.zeros((6, 4), dtype=np.float32)
    var = np.zeros((6, 4), dtype=np.float32)
    for k in range(50):
        ol.lib().orc_update_frame_result(np.ascontiguousarray(x[k]).reshape(-1), prog.reshape(-1), var.reshape(-1), 24, k + 1)
    assert np.allclose(prog, x.mean(axis=0), rtol=1e-5)
    assert np.allclose(var, ((x - x.mean(axis=0)) ** 2).sum(axis=0), rtol=1e-3)


def test_camera_defaults():
    cam = ol.camera_look_at(aspect=2.0)
    eye, U, V, W = cam[0:3], cam[3:6], cam[6:9], cam[9:12]
    assert np.allclose(W, -eye)
    wlen = np.linalg.norm(W)
    assert abs(np.linalg.norm(U) - wlen * np.tan(np.radians(15.0))) < 1e-6
    assert abs(np.linalg.norm(V) - np.linalg.norm(U) / 2.0) < 1e-6
    assert abs(np.dot(U, V)) < 1e-6 and abs(np.dot(U, W)) < 1e-6 and V[1] > 0


def test_tonemap_basic():
    rng = np.random.default_rng(2)
    img = rng.uniform(0, 4, (8, 16, 4)).astype(np.float32)
    img[0, 0, :3] = 0
    out, avg = ol.tonemap(img, 0.4)
    lum = img[..., 0] * 0.265068 + img[..., 1] * 0.67023428 + img[..., 2] * 0.06409157
    assert abs(avg - (lum + 1e-5).mean()) < 1e-4
    # lw == 0: ld / lw = NaN, optix::clamp turns it into 1 -> white, as the reference does (pinned in test_oracle_vs_ref.py)
    assert np.all(out[..., 3] == 255) and np.all(out[0, 0, :3] == 255)


def test_render_is_deterministic_and_sky_is_black(oracle_small):
    o = oracle_small
    cam = ol.camera_look_at(aspect=32 / 18)
    a = o.render_frame(cam, 32, 18, ol.MODE_ALL, 3)
    b = o.render_frame(cam, 32, 18, ol.MODE_ALL, 3)
    c = o.render_frame(cam, 32, 18, ol.MODE_ALL, 4)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.all(a[..., 3] == 1)
    assert np.all(a[0, :, :3] == 0) or np.all(a[-1, :, :3] == 0)  # top/bottom rows look past the cloud
    assert a[..., 0].max() > 0 and np.isfinite(a).all()


def test_generated_points_lie_inside_the_cloud(oracle_small):
    o = oracle_small
    p, d = o.generate_points(0, 64)
    bbox = o.derived()["bbox"]
    assert np.all(np.abs(p) <= bbox / 2 + 0.011)
    assert np.abs(np.linalg.norm(d, axis=1) - 1).max() < 1e-5
    dens = o.sample_volume(p + bbox / 2)
    assert (dens > 0).mean() > 0.9  # collisions happen where there is density
    p2, d2 = o.generate_points(0, 64)
    assert np.array_equal(p, p2) and np.array_equal(d, d2)


def test_point_radiance_scheduler_follows_the_collector(oracle_small):
    o = oracle_small
    p, d = o.generate_points(0, 6)
    tasks, conv, nconv, updates = o.point_radiance(p, d, max_threads=48, launches_per_update=20, max_updates=3)
    assert updates >= 1 and nconv == conv.sum()
    assert list(tasks["id"]) == list(range(6))
    # every sample got taskRepeatCount * launches experiments in the first update (RadianceCollector.cpp:176-192)
    assert tasks["experimentCount"].min() >= 8 * 20
    assert np.all(tasks["radiance"] >= 0) and np.isfinite(tasks["runningVariance"]).all()


REF_VECTOR = Path("/root/reference/DeepestScatter_Train/Common/Vector.py")


@pytest.mark.skipif(not REF_VECTOR.is_file(), reason="the reference tree is not mounted here")
def test_descriptor_stencil_follows_the_references_basis(monkeypatch):
    """The stencil's tap addresses (integer work) against an independent numpy construction that takes the frame from the reference's own
    `descriptorBasis` (DeepestScatter_Train/Common/Vector.py:32-37 -- the Python twin of DisneyDescriptor.cuh:74-76): eZ against the light, eX
    across the view direction; taps z outermost / x innermost on a 5 x 5 x 9 lattice from (-2,-2,-2) to (2,2,6), spacing 0.5 / densityMultiplier
    doubling per layer, mip level max(0, -log2(voxel in free paths) - 1 + layer).  Float rounding may move a tap that sits on a voxel boundary:
    at most 0.5 % of the addresses may differ, and then by one voxel."""
    import importlib
    import sys

    monkeypatch.syspath_prepend(str(REF_VECTOR.parent))
    sys.modules.pop("Vector", None)
    V = importlib.import_module("Vector")
    try:
        n = 64
        o = ol.Oracle()
        o.volume_synth(n, 0, 1234, True)
        light = np.float32([0.3, -0.8, 0.52])
        o.scene_set(7000.0, light)
        der = o.derived()
        pos, dirs = o.generate_points(0, 40, 0)
        _, idx = o.descriptors(pos, dirs, as_float=True, want_index=True)
        levels = o.level_count()
        dims = [o.level_dims(l) for l in range(levels)]
        light_n = der["light"].astype(np.float64)
        mip0 = -np.log2(float(der["voxel_free_path"])) - 1.0
        lattice = np.array([(x, y, z) for z in range(-2, 7) for y in range(-2, 3) for x in range(-2, 3)], np.float64)  # sampleId order
        bad = total = 0
        for s in range(len(pos)):
            eX, eY, eZ = V.descriptorBasis(light_n, dirs[s].astype(np.float64))
            origin = pos[s].astype(np.float64) + 0.5 * der["bbox"].astype(np.float64)
            for layer in range(10):
                scale = 0.5 / der["density_multiplier"] * 2.0**layer
                lod = min(max(mip0 + layer, 0.0), levels - 1)
                l0 = int(np.floor(lod))
                nx, ny, nz = dims[l0]
                p = origin + (lattice[:, 0:1] * eX + lattice[:, 1:2] * eY + lattice[:, 2:3] * eZ) * scale
                uvw = p * der["texture_scale"].astype(np.float64)
                want = np.stack([np.clip(np.floor(uvw[:, 0] * nx - 0.5), -2, nx + 1), np.clip(np.floor(uvw[:, 1] * ny - 0.5), -2, ny + 1),
                                 np.clip(np.floor(uvw[:, 2] * nz - 0.5), -2, nz + 1)], axis=1).astype(np.int64)
                got = idx[s, layer].astype(np.int64)
                assert np.all(got[:, 3] == l0)
                diff = np.abs(got[:, :3] - want)
                assert diff.max() <= 1
                bad += int((diff.max(axis=1) > 0).sum())
                total += 225
        assert bad <= 0.005 * total, f"{bad} of {total} tap addresses differ"
    finally:
        sys.modules.pop("Vector", None)
