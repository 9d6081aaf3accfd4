"""Parity of the sm_100a CUDA path (called through the C ABI) against the host oracle.

EXACT arithmetic flavour: bit-for-bit equality on the same seeded inputs (integer work AND radiance).
FAST flavour (hardware texture filtering + MUFU intrinsics, what the reference itself runs): statistical
agreement -- per pixel within 3 sigma, image mean within 0.5 %.
"""
import numpy as np
import pytest

import oracle_lib as ol
from pathlib import Path

from conftest import SCENE_SMALL

ROOT = Path(__file__).resolve().parent.parent

pytestmark = pytest.mark.gpu

W, H = 40, 22  # deliberately not multiples of the 8x4 work tiles


def cam_pair(ds, w, h):
    cam = ds.camera_look_at(aspect=w / h)
    return cam, ds.camera_array(cam)


# ---------------------------------------------------------------- integer / byte work: bit exact

@pytest.mark.parametrize("kind", [0, 1, 2])
def test_synthetic_grid_and_mips_bit_exact(built_library, kind):
    ds = built_library
    o = ol.Oracle()
    o.volume_synth(48, kind, 99, True)
    with ds.Context(0) as ctx:
        ctx.volume_synth(48, kind, 99, True)
        assert ctx.level_count() == o.level_count()
        for l in range(o.level_count()):
            assert ctx.level_dims(l) == o.level_dims(l)
            assert np.array_equal(ctx.level(l), o.level(l)), f"kind {kind} level {l}"


@pytest.mark.parametrize("shape", [(13, 7, 21), (1, 5, 9), (33, 32, 31), (64, 64, 64)])
def test_uploaded_volume_mips_bit_exact(built_library, shape):
    ds = built_library
    rng = np.random.default_rng(7)
    g = rng.integers(0, 256, shape, dtype=np.uint8)
    o = ol.Oracle()
    o.volume_upload(g, True)
    with ds.Context(0) as ctx:
        ctx.volume_upload(g, True)
        assert ctx.level_count() == o.level_count()
        for l in range(o.level_count()):
            assert np.array_equal(ctx.level(l), o.level(l))
        assert ctx.level(ctx.level_count() - 1).shape == (1, 1, 1)
        ctx.volume_upload(g, False)
        assert ctx.level_count() == 1


def test_float_grid_quantisation_bit_exact(built_library):
    ds = built_library
    rng = np.random.default_rng(3)
    v = rng.uniform(0, 2.5, (9, 10, 11)).astype(np.float32)
    ref = np.empty(v.size, dtype=np.uint8)
    ol.lib().orc_quantize_float_grid(v.reshape(-1), v.size, 2.5, ref)
    with ds.Context(0) as ctx:
        ctx.volume_upload_float(v, 2.5, False)
        assert np.array_equal(ctx.level(0).reshape(-1), ref)


def test_derived_scene_variables_equal(gpu_small, oracle_small):
    a, b = gpu_small.derived(), oracle_small.derived()
    for k in a:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k


def test_bake_exact_is_bit_exact(gpu_small, oracle_small):
    assert np.array_equal(gpu_small.inscatter(), oracle_small.inscatter())


def test_bake_fast_within_one_lsb(built_library, oracle_small):
    ds = built_library
    with ds.Context(0) as ctx:
        ctx.set_option("precision", ds.PRECISION_FAST)
        ctx.volume_synth(SCENE_SMALL["n"], SCENE_SMALL["kind"], SCENE_SMALL["seed"])
        ctx.scene_set(SCENE_SMALL["cloud_size_m"], SCENE_SMALL["light_dir"])
        ctx.bake()
        diff = np.abs(ctx.inscatter().astype(int) - oracle_small.inscatter().astype(int))
        assert diff.max() <= 3 and (diff > 1).mean() < 0.01 and diff.mean() < 0.2


def test_bake_skip_empty_does_not_change_a_byte(built_library, oracle_small):
    ds = built_library
    with ds.Context(0) as ctx:
        ctx.set_option("precision", ds.PRECISION_EXACT)
        ctx.set_option("skip_empty", 0)
        ctx.volume_synth(SCENE_SMALL["n"], SCENE_SMALL["kind"], SCENE_SMALL["seed"])
        ctx.scene_set(SCENE_SMALL["cloud_size_m"], SCENE_SMALL["light_dir"])
        ctx.bake()
        assert np.array_equal(ctx.inscatter(), oracle_small.inscatter())


# ---------------------------------------------------------------- estimators, EXACT flavour: bit exact

def _random_rays(n, seed):
    rng = np.random.default_rng(seed)
    origins = rng.normal(size=(n, 3))
    origins = (origins / np.linalg.norm(origins, axis=1, keepdims=True) * 2.0).astype(np.float32)
    target = rng.uniform(-0.3, 0.3, (n, 3)).astype(np.float32)
    dirs = target - origins
    dirs = (dirs / np.linalg.norm(dirs, axis=1, keepdims=True)).astype(np.float32)
    # a few rays from inside the box and a few that miss
    origins[:8] = rng.uniform(-0.2, 0.2, (8, 3))
    dirs[8:16] = -dirs[8:16]
    val0 = rng.integers(0, 2**32, n, dtype=np.uint32)
    stream = rng.integers(1, 1000, n).astype(np.uint32)
    return origins, dirs, val0, stream


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_trace_paths_bit_exact_and_counters_equal(gpu_small, oracle_small, mode):
    origins, dirs, val0, stream = _random_rays(3000, 42 + mode)
    oracle_small.counters_reset()
    ref = oracle_small.trace_paths(mode, origins, dirs, val0, stream)
    oc = oracle_small.counters()
    gpu_small.counters_reset()
    got = gpu_small.trace_paths(mode, origins, dirs, val0, stream)
    gc = gpu_small.counters()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert ref[:, 0].max() > 0
    assert (gc["paths"], gc["events"], gc["steps"]) == (oc["paths"], oc["events"], oc["steps"])
    assert gc["density_taps"] <= gc["steps"] and gc["nonfinite"] == 0


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_render_frame_result_bit_exact(gpu_small, oracle_small, built_library, mode):
    cam, cam_np = cam_pair(built_library, W, H)
    gpu_small.frame_create(W, H)
    got = gpu_small.render_frame(cam, mode, 5)
    ref = oracle_small.render_frame(cam_np, W, H, mode, 5)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert np.all(got[..., 3] == 1)


def test_progressive_accumulation_bit_exact(gpu_small, oracle_small, built_library):
    """Camera::render loop: 10 subframes then 7 more, chunked through a 4-subframe staging buffer."""
    cam, cam_np = cam_pair(built_library, W, H)
    gpu_small.frame_create(W, H)
    gpu_small.set_option("staging_subframes", 4)
    gpu_small.render_subframes(cam, 0, 1, 10)
    gpu_small.render_subframes(cam, 0, 11, 7)
    gpu_small.set_option("staging_subframes", 16)
    p, v = gpu_small.frame_download()
    rp, rv = oracle_small.render_accumulate(cam_np, W, H, 0, 1, 17)
    assert np.array_equal(p.view(np.uint32), rp.view(np.uint32))
    assert np.array_equal(v.view(np.uint32), rv.view(np.uint32))
    assert gpu_small.unconverged(17) == int(ol.lib().orc_unconverged_pixels(rp.reshape(-1), rv.reshape(-1), W * H, 17))


def test_render_subframes_host_round_trip(gpu_small, oracle_small, built_library):
    cam, cam_np = cam_pair(built_library, W, H)
    gpu_small.frame_create(W, H)
    p = np.zeros((H, W, 4), dtype=np.float32)
    v = np.zeros((H, W, 4), dtype=np.float32)
    gpu_small.render_subframes_host(cam, 2, 1, 3, p, v)
    gpu_small.render_subframes_host(cam, 2, 4, 2, p, v)
    rp, rv = oracle_small.render_accumulate(cam_np, W, H, 2, 1, 5)
    assert np.array_equal(p, rp) and np.array_equal(v, rv)


def test_results_do_not_depend_on_scheduling_knobs(gpu_small, built_library):
    """Per-path RNG streams depend on the work item only: block size, residency, the march/event vote
    and empty-space skipping must not change a bit."""
    cam, _ = cam_pair(built_library, W, H)
    gpu_small.frame_create(W, H)
    base = gpu_small.render_frame(cam, 0, 2)
    gpu_small.counters_reset()
    gpu_small.render_frame(cam, 0, 2)
    c0 = gpu_small.counters()
    try:
        for opts in (dict(block_threads=64, blocks_per_sm=1), dict(march_keep_quarters=0), dict(march_keep_quarters=4, march_max_iters=3),
                     dict(skip_empty=0)):
            for k, val in opts.items():
                gpu_small.set_option(k, val)
            gpu_small.counters_reset()
            again = gpu_small.render_frame(cam, 0, 2)
            c = gpu_small.counters()
            assert np.array_equal(again.view(np.uint32), base.view(np.uint32)), opts
            assert (c["paths"], c["events"], c["steps"]) == (c0["paths"], c0["events"], c0["steps"])
            if opts.get("skip_empty") == 0:
                assert c["density_taps"] == c["steps"] and c0["density_taps"] < c0["steps"]
            for k, val in dict(block_threads=512, blocks_per_sm=2, march_keep_quarters=2, march_max_iters=64, skip_empty=1).items():
                gpu_small.set_option(k, val)
    finally:
        for k, val in dict(block_threads=512, blocks_per_sm=2, march_keep_quarters=2, march_max_iters=64, skip_empty=1).items():
            gpu_small.set_option(k, val)


# ---------------------------------------------------------------- estimators, FAST flavour: statistical

def test_fast_flavour_matches_oracle_statistically(built_library, oracle_small):
    ds = built_library
    w, h, spp = 48, 27, 48
    cam, cam_np = cam_pair(ds, w, h)
    rp, rv = oracle_small.render_accumulate(cam_np, w, h, 0, 1, spp)
    with ds.Context(0) as ctx:
        assert ctx.get_option("precision") == ds.PRECISION_FAST  # the default flavour
        ctx.volume_synth(SCENE_SMALL["n"], SCENE_SMALL["kind"], SCENE_SMALL["seed"])
        ctx.scene_set(SCENE_SMALL["cloud_size_m"], SCENE_SMALL["light_dir"])
        ctx.bake()
        ctx.frame_create(w, h)
        ctx.render_subframes(cam, 0, 1, spp)
        p, v = ctx.frame_download()
        assert ctx.counters()["nonfinite"] == 0
    a, b = p[..., 0].astype(np.float64), rp[..., 0].astype(np.float64)
    va, vb = v[..., 0].astype(np.float64) / (spp - 1), rv[..., 0].astype(np.float64) / (spp - 1)
    sigma = np.sqrt((va + vb) / spp)
    lit = sigma > 0
    z = np.abs(a - b)[lit] / sigma[lit]
    assert (z < 3).mean() > 0.99, f"per-pixel 3-sigma agreement only {(z < 3).mean():.4f}"
    assert np.all((a == 0) == (b == 0)) or ((a == 0) != (b == 0)).mean() < 0.01  # same silhouette
    assert abs(a.mean() - b.mean()) / b.mean() < 0.005 + 3 * np.sqrt((sigma**2).sum()) / a.size / b.mean()


def test_converged_images_agree_within_half_a_percent_rmse(built_library):
    """north_star's radiance bar at convergence: the FAST flavour (hardware trilinear taps, MUFU transcendentals) against the
    EXACT flavour (bit-identical to the oracle, see the *_bit_exact tests) at 2 M spp per pixel: every pixel within 3 sigma and
    image-wide relative RMSE below 0.5 %."""
    ds = built_library
    w, h, spp, chunk = 32, 18, 1 << 21, 1 << 13
    cam, _ = cam_pair(ds, w, h)
    res = {}
    with ds.Context(0) as ctx:
        ctx.set_option("staging_subframes", chunk)
        ctx.volume_synth(SCENE_SMALL["n"], SCENE_SMALL["kind"], SCENE_SMALL["seed"])
        ctx.scene_set(SCENE_SMALL["cloud_size_m"], SCENE_SMALL["light_dir"])
        for flavour in (ds.PRECISION_EXACT, ds.PRECISION_FAST):
            ctx.set_option("precision", flavour)
            ctx.bake()
            ctx.frame_create(w, h)
            for first in range(1, spp + 1, chunk):
                ctx.render_subframes(cam, ds.MODE_ALL_SCATTER, first, chunk)
            p, v = ctx.frame_download()
            assert ctx.counters()["nonfinite"] == 0
            res[flavour] = (p[..., 0].astype(np.float64), v[..., 0].astype(np.float64) / (spp - 1))
    (a, va), (b, vb) = res[ds.PRECISION_FAST], res[ds.PRECISION_EXACT]
    assert ((a == 0) != (b == 0)).mean() < 0.01  # silhouette: a grazing pixel may see density in one filter and not the other
    lit = (a > 0) & (b > 0)
    sigma = np.sqrt((va + vb) / spp)
    z = np.abs(a - b)[lit] / sigma[lit]
    rmse = np.sqrt(np.mean((a - b) ** 2)) / b.mean()
    print(f"converged FAST vs EXACT: relative RMSE {rmse:.5f}, max z {z.max():.2f}, mean ratio {a.mean() / b.mean():.6f}")
    assert rmse < 0.005, rmse
    assert (z < 3).mean() > 0.98 and z.max() < 5, (float((z < 3).mean()), float(z.max()))
    assert abs(a.mean() / b.mean() - 1) < 0.002


def test_fast_flavour_is_deterministic_under_scheduling_knobs(built_library):
    """k_trace_fast: a path's arithmetic never depends on which lane runs it or on the warp's phase votes."""
    ds = built_library
    w, h = 64, 36
    cam, _ = cam_pair(ds, w, h)
    with ds.Context(0) as ctx:
        ctx.volume_synth(SCENE_SMALL["n"], SCENE_SMALL["kind"], SCENE_SMALL["seed"])
        ctx.scene_set(SCENE_SMALL["cloud_size_m"], SCENE_SMALL["light_dir"])
        ctx.bake()
        ctx.frame_create(w, h)
        ctx.counters_reset()
        base = ctx.render_frame(cam, 0, 7)
        c0 = ctx.counters()
        defaults = dict(regen_min=2, skip_min=8, march_keep32=12, march_max_iters=64, skip_open_dist=1, skip_max_iters=32, march_unroll=2,
                        zero_check_min=1, block_threads=576, blocks_per_sm=2, fused_volume=1, escape_octants=1)
        # positions are a function of the step index (q0 + n * sv), so neither the phase votes, nor how leaps are cut,
        # nor the number of march steps per vote can change a single bit of a path
        for opts in (dict(regen_min=1, skip_min=1), dict(regen_min=32, skip_min=32), dict(march_keep32=0),
                     dict(march_keep32=31, march_max_iters=2), dict(block_threads=64, blocks_per_sm=1), dict(march_unroll=1), dict(zero_check_min=4), dict(zero_check_min=32, march_unroll=1),
                     dict(skip_max_iters=1, skip_open_dist=3), dict(block_threads=512), dict(block_threads=640, march_unroll=2),
                     # the two R8 arrays instead of the fused RG8 {density, sun transmittance} array: each channel filters to the same value
                     dict(fused_volume=0), dict(fused_volume=0, march_unroll=1),
                     # without the escape octants a leaving path walks the leap DDA to the grid face: same steps, same bits
                     dict(escape_octants=0), dict(escape_octants=0, skip_max_iters=2)):
            for k, val in opts.items():
                ctx.set_option(k, val)
            ctx.counters_reset()
            again = ctx.render_frame(cam, 0, 7)
            c = ctx.counters()
            assert np.array_equal(again.view(np.uint32), base.view(np.uint32)), opts
            # taps may differ (a lane past the cloud keeps tapping zeros until the end of its march round); the
            # reference-algorithm counts may not
            assert (c["paths"], c["events"], c["steps"]) == (c0["paths"], c0["events"], c0["steps"]), opts
            for k, val in defaults.items():
                ctx.set_option(k, val)
        # without the primary-ray cache every pixel is traced from the box face: same bits, same counters
        ctx.set_option("primary_cache", 0)
        ctx.counters_reset()
        nocache = ctx.render_frame(cam, 0, 7)
        c1 = ctx.counters()
        ctx.set_option("primary_cache", 1)
        assert np.array_equal(nocache.view(np.uint32), base.view(np.uint32))
        assert (c1["paths"], c1["events"], c1["steps"]) == (c0["paths"], c0["events"], c0["steps"])
        assert c1["paths"] == w * h
        # jump-free marching takes every tap: same bits as the leaping kernel
        ctx.set_option("skip_empty", 0)
        ctx.counters_reset()
        noskip = ctx.render_frame(cam, 0, 7)
        c2 = ctx.counters()
        ctx.set_option("skip_empty", 1)
        assert np.array_equal(noskip.view(np.uint32), base.view(np.uint32))
        assert (c2["paths"], c2["events"], c2["steps"]) == (c0["paths"], c0["events"], c0["steps"])
        assert c2["density_taps"] >= c2["steps"]


# ---------------------------------------------------------------- dataset generation

def test_generate_points_bit_exact(gpu_small, oracle_small):
    p, d = gpu_small.generate_points(100, 257, 0)
    rp, rd = oracle_small.generate_points(100, 257, 0)
    assert np.array_equal(p.view(np.uint32), rp.view(np.uint32))
    assert np.array_equal(d.view(np.uint32), rd.view(np.uint32))


@pytest.mark.parametrize("size_m", [1000.0, 7000.0, 12000.0])
def test_descriptors_u8_and_indices_bit_exact_floats_1e6(built_library, size_m):
    ds = built_library
    o = ol.Oracle()
    o.volume_synth(64, 0, 1234, True)
    o.scene_set(size_m, (0.3, -0.8, 0.52))
    with ds.Context(0) as ctx:
        ctx.set_option("precision", ds.PRECISION_EXACT)
        ctx.volume_synth(64, 0, 1234, True)
        ctx.scene_set(size_m, (0.3, -0.8, 0.52))
        p, d = ctx.generate_points(0, 96, 0)
        # add samples at the box faces / corners to exercise the edge fade and clamping
        p[:4] = [[0.49, 0.0, 0.0], [-0.5, -0.5, -0.5], [0.0, 0.499, 0.3], [0.2, -0.1, -0.5]]
        got_u8 = ctx.descriptors(p, d)
        got_f, got_idx = ctx.descriptors(p, d, as_float=True, want_index=True)
    ref_u8 = o.descriptors(p, d)
    ref_f, ref_idx = o.descriptors(p, d, as_float=True, want_index=True)
    assert got_u8.shape == (96, 10, 225)
    assert np.array_equal(got_idx, ref_idx)  # stencil voxel addresses + mip level
    assert np.array_equal(got_u8, ref_u8)  # record bytes
    denom = np.maximum(np.abs(ref_f), 1e-12)
    assert (np.abs(got_f - ref_f) / denom).max() <= 1e-6
    assert got_u8.max() > 0


@pytest.mark.parametrize("size_m", [1000.0, 7000.0])
def test_descriptors_on_the_texture_units_match_within_filter_precision(built_library, size_m):
    """FAST flavour of the neural renderer's descriptor gather: the same taps through a mip-mapped hardware texture (rtTex3DLod in the
    reference).  The texture unit quantises the trilinear and the mip weights to 8 fractional bits: each of the three lerps inside a
    level and the one between levels is off by at most 2^-9 of the local value range, so |hw - exact| <= 4 * 2^-9 in normalised density."""
    ds = built_library
    o = ol.Oracle()
    o.volume_synth(64, 0, 1234, True)
    o.scene_set(size_m, (0.3, -0.8, 0.52))
    with ds.Context(0) as ctx:
        ctx.volume_synth(64, 0, 1234, True)
        ctx.scene_set(size_m, (0.3, -0.8, 0.52))
        p, d = ctx.generate_points(0, 96, 0)
        p[:4] = [[0.49, 0.0, 0.0], [-0.5, -0.5, -0.5], [0.0, 0.499, 0.3], [0.2, -0.1, -0.5]]
        soft = ctx.descriptors(p, d, as_float=True)
        ctx.set_option("descriptor_hw", 1)
        hw = ctx.descriptors(p, d, as_float=True)
        assert ctx.descriptors(p, d).dtype == np.uint8  # the byte collector never takes the hardware path
    ref = o.descriptors(p, d, as_float=True)
    assert np.abs(soft - ref).max() <= 1e-6
    assert not np.array_equal(hw, soft)  # it really went through the texture units
    assert np.abs(hw - ref).max() <= 4.0 / 512.0
    assert np.abs(hw - ref).mean() <= 1e-3
    assert np.array_equal(hw == 0, ref == 0) or (np.abs(hw - ref)[(hw == 0) != (ref == 0)].max() <= 4.0 / 512.0)


def test_point_radiance_collector_bit_exact(gpu_small, oracle_small):
    p, d = oracle_small.generate_points(0, 6)
    rt, rc, rn, ru = oracle_small.point_radiance(p, d, max_threads=48, launches_per_update=20, max_updates=3)
    gt, gc, gn, gu = gpu_small.point_radiance(p, d, max_threads=48, launches_per_update=20, max_updates=3)
    assert gu == ru and gn == rn and np.array_equal(gc, rc)
    for f in ("id", "experimentCount"):
        assert np.array_equal(gt[f], rt[f])
    for f in ("radiance", "runningVariance", "position", "direction"):
        assert np.array_equal(gt[f].view(np.uint32), rt[f].view(np.uint32)), f


# ---------------------------------------------------------------- tone map, moments

def test_device_resident_collector_agrees_with_the_reference_schedule(built_library):
    """FAST flavour: the one-launch adaptive collector applies the same convergence rule as RadianceCollector's update
    loop; the two estimates of every sample agree within their confidence intervals."""
    ds = built_library
    with ds.Context(0) as ctx:
        ctx.volume_synth(SCENE_SMALL["n"], SCENE_SMALL["kind"], SCENE_SMALL["seed"])
        ctx.scene_set(SCENE_SMALL["cloud_size_m"], SCENE_SMALL["light_dir"])
        ctx.bake()
        pos, dirs = ctx.generate_points(0, 96, stream=3)
        kw = dict(max_threads=96 * 20, launches_per_update=50, max_updates=1500)
        ctx.set_option("radiance_scheduler", 0)
        ref, ref_conv, _, ref_updates = ctx.point_radiance(pos, dirs, **kw)
        ctx.set_option("radiance_scheduler", 1)
        ctx.counters_reset()
        ad, ad_conv, _, ad_updates = ctx.point_radiance(pos, dirs, **kw)
        c = ctx.counters()
    assert ad_updates == 1 and ref_updates > 1
    assert c["nonfinite"] == 0 and c["paths"] >= int(ad["experimentCount"].astype(np.int64).sum())
    assert np.all(ad["experimentCount"] >= 20 * 50)  # never tested before the reference's first test
    both = ref_conv & ad_conv
    assert both.mean() > 0.6  # the rest needs more than the 1.5 M experiments this test allows per sample
    assert np.all(ad["experimentCount"][~ad_conv] >= 1500 * 20 * 50)  # closed by the cap, not dropped

    def ci(t):
        n = t["experimentCount"].astype(np.float64)
        return 1.96 * np.sqrt(t["runningVariance"] / n) / np.sqrt(n)

    # A closed sample satisfied the rule WHEN it was closed; paths still in flight are added afterwards, and a rule that stops
    # at the first time a running interval dips below the threshold stops on under-estimated variances (the reference's
    # schedule has the same selection effect, it just never looks again) -- so afterwards most, not all, still satisfy it
    a = ad[ad_conv]
    rel = ci(a) / (a["radiance"] + np.finfo(np.float32).eps)
    ok = (rel < 0.02 * 1.3) | (ci(a) < 1e-4 * 1.3) | ((a["radiance"] < np.finfo(np.float32).eps) & (a["experimentCount"] > 100000))
    assert ok.mean() > 0.75, float(ok.mean())
    # the two schedules estimate the same radiance: differences within the combined 95 % intervals (4 sigma slack)
    diff = np.abs(ad["radiance"][both].astype(np.float64) - ref["radiance"][both])
    tol = 2.1 * (ci(ad[both]) + ci(ref[both])) + 1e-6
    assert (diff <= tol).mean() > 0.97, float((diff <= tol).mean())
    assert abs(ad["radiance"][both].mean() - ref["radiance"][both].mean()) < 0.01 * ref["radiance"][both].mean() + 1e-5


def test_network_input_pass_matches_oracle(gpu_small, oracle_small, built_library):
    """DisneyRenderer's first launch (disneyCamera.cu pinholeCamera + disneyDescriptorMaterial.cu): EXACT flavour bit-exact on the
    direct radiance / transmittance / hasScattered, descriptor densities and the angle within 1e-6."""
    ds = built_library
    w, h = 64, 36
    cam, cam_arr = cam_pair(ds, w, h)
    rect = (8, 6, 40, 24)
    ref_in, ref_info = oracle_small.network_input(cam_arr, w, h, rect, stream=5)
    got_in, got_info = gpu_small.network_input(cam, w, h, rect, stream=5)
    has = ref_info[..., 4] > 0
    assert 0.2 < has.mean() < 1.0  # the rectangle straddles the silhouette
    assert np.array_equal(got_info["hasScattered"] != 0, has)
    assert np.array_equal(got_info["radiance"].view(np.uint32), ref_info[..., :3].copy().view(np.uint32))
    assert np.array_equal(got_info["transmittance"].view(np.uint32), ref_info[..., 3].copy().view(np.uint32))
    assert np.all(got_info["transmittance"][~has] > 0)
    # descriptor: densities as the float descriptor test (<= 1e-6), zero where nothing scattered; angle (device acosf vs libm) 1e-6
    assert np.all(got_in[~has][:, :, :225] == 0) and np.all(ref_in[~has][:, :, :225] == 0)
    assert np.abs(got_in[..., :225] - ref_in[..., :225]).max() <= 1e-6
    assert np.abs(got_in[..., 225] - ref_in[..., 225]).max() <= 1e-6 and got_in[..., 225].min() > 0
    assert (got_in[has][:, 0, :225].max(axis=1) > 0).all()  # a collision point sits in the cloud: layer 0 sees density
    # the same seeds in every rectangle (rect-local launch index, as in the reference): a shifted rectangle differs
    other_in, other_info = gpu_small.network_input(cam, w, h, (9, 6, 40, 24), stream=5)
    assert not np.array_equal(other_info["radiance"][:, :-1], got_info["radiance"][:, 1:])
    # copyToFrameResult
    frame = np.zeros((h, w, 4), np.float32)
    predicted = np.full((rect[3], rect[2]), 0.25, np.float32)
    ds.blit_predicted(frame, rect, predicted, got_info)
    sub = frame[rect[1]:rect[1] + rect[3], rect[0]:rect[0] + rect[2]]
    want = (0.25 + ref_info[..., :3]) * (1 - ref_info[..., 3:4])
    assert np.array_equal(sub[..., :3][has], want[has].astype(np.float32)) and np.all(sub[~has] == 0)
    assert frame.sum() == sub.sum()


def test_network_input_pass_fast_flavour(built_library, oracle_small):
    """FAST flavour of the same pass: same silhouette, transmittance within texture-filter precision."""
    ds = built_library
    w, h = 64, 36
    cam, cam_arr = cam_pair(ds, w, h)
    rect = (0, 0, w, h)
    ref_in, ref_info = oracle_small.network_input(cam_arr, w, h, rect, stream=2)
    with ds.Context(0) as ctx:
        ctx.volume_synth(SCENE_SMALL["n"], SCENE_SMALL["kind"], SCENE_SMALL["seed"])
        ctx.scene_set(SCENE_SMALL["cloud_size_m"], SCENE_SMALL["light_dir"])
        ctx.bake()
        got_in, got_info = ctx.network_input(cam, w, h, rect, stream=2)
    has = ref_info[..., 4] > 0
    assert ((got_info["hasScattered"] != 0) != has).mean() < 0.01
    assert np.abs(got_info["transmittance"] - ref_info[..., 3]).max() < 0.02
    both = has & (got_info["hasScattered"] != 0)
    assert abs(got_info["radiance"][both].mean() - ref_info[..., :3][both].mean()) < 0.03 * ref_info[..., :3][both].mean()
    assert np.abs(got_in[..., 225] - ref_in[..., 225]).max() <= 1e-5


def test_tonemap_matches_reinhard_oracle(gpu_small, oracle_small, built_library):
    cam, cam_np = cam_pair(built_library, W, H)
    gpu_small.frame_create(W, H)
    gpu_small.render_subframes(cam, 0, 1, 4)
    p, _ = gpu_small.frame_download()
    got, avg = gpu_small.tonemap(0.4)
    ref, ravg = ol.tonemap(p, 0.4)
    assert avg == ravg  # same summation order as reinhard.cu firstPass/secondPass
    assert np.abs(got.astype(int) - ref.astype(int)).max() <= 1  # powf differs by an ulp between libm and CUDA
    assert np.all(got[..., 3] == 255)


def test_moments_export_import_round_trip_and_merge(gpu_small, built_library):
    import torch

    cam, _ = cam_pair(built_library, W, H)
    gpu_small.frame_create(W, H)
    gpu_small.render_subframes(cam, 0, 1, 6)
    full_p, full_v = gpu_small.frame_download()
    ma = torch.zeros(W * H * 8, dtype=torch.float64, device="cuda")
    mb = torch.zeros_like(ma)
    gpu_small.frame_clear()
    gpu_small.render_subframes(cam, 0, 1, 4)  # "rank 0": subframes 1..4
    gpu_small.export_moments(4, ma.data_ptr())
    gpu_small.frame_clear()
    # "rank 1": global subframes 5..6 (RNG streams 5, 6) accumulated as a fresh Welford stream (weights 1/1, 1/2)
    gpu_small.set_option("stream_offset", 4)
    gpu_small.render_subframes(cam, 0, 1, 2)
    gpu_small.set_option("stream_offset", 0)
    gpu_small.export_moments(2, mb.data_ptr())
    gpu_small.sync()
    torch.cuda.synchronize()
    total = ma + mb
    gpu_small.import_moments(6, total.data_ptr())
    p, v = gpu_small.frame_download()
    assert np.allclose(p, full_p, rtol=2e-6, atol=1e-7)
    assert np.allclose(v, full_v, rtol=1e-4, atol=1e-3 * max(1.0, float(full_v.max()) * 1e-6))


# ---------------------------------------------------------------- error behaviour and edge cases

def test_error_paths(built_library):
    ds = built_library
    with ds.Context(0) as ctx:
        cam = ds.camera_look_at()
        with pytest.raises(ds.DsError):  # no volume yet
            ctx.bake()
        ctx.volume_synth(16, 0, 1, True)
        with pytest.raises(ds.DsError):  # scene not set
            ctx.bake()
        ctx.scene_set()
        ctx.frame_create(8, 8)
        with pytest.raises(ds.DsError):  # not baked
            ctx.render_subframes(cam, 0, 1, 1)
        ctx.bake()
        with pytest.raises(ds.DsError):  # "Invalid Render Mode" (CloudMaterial.cpp:62)
            ctx.render_subframes(cam, 7, 1, 1)
        with pytest.raises(ds.DsError):  # subframe ids are 1-based
            ctx.render_subframes(cam, 0, 0, 1)
        with pytest.raises(ds.DsError):
            ctx.frame_create(5000, 10)
        with pytest.raises(ds.DsError):
            ctx.set_option("no_such_option", 1)
        # empty inputs are fine
        assert ctx.trace_paths(0, np.zeros((0, 3)), np.zeros((0, 3)), [], []).shape == (0, 3)
        assert ctx.descriptors(np.zeros((0, 3)), np.zeros((0, 3))).shape == (0, 10, 225)
        ctx.render_subframes(cam, 0, 1, 0)
        # a changed sun invalidates the baked volume
        ctx.scene_set(light_dir=(0, -1, 0))
        with pytest.raises(ds.DsError):
            ctx.render_subframes(cam, 0, 1, 1)


def test_full_size_frame_properties(built_library):
    """1920x1080 at the C2 grid size (512^3): properties that need no oracle."""
    ds = built_library
    with ds.Context(0) as ctx:
        ctx.volume_synth(512, 0, 1234, True)
        ctx.scene_set(7000.0, (-0.586, -0.766, -0.271))
        ctx.bake()
        ins = ctx.inscatter()
        assert ins.max() == 255 and ins.min() == 0
        ctx.frame_create(1920, 1080)
        cam = ds.camera_look_at(aspect=1920 / 1080)
        ctx.render_subframes(cam, 0, 1, 2)
        p, v = ctx.frame_download()
        c = ctx.counters()
        assert c["paths"] == 2 * 1920 * 1080 and c["nonfinite"] == 0
        assert np.isfinite(p).all() and np.isfinite(v).all() and p.min() >= 0
        assert np.all(p[..., 3] == 1) and np.all(v[..., 3] == 0)
        assert np.all(p[:8, :, :3] == 0) and np.all(p[-8:, :, :3] == 0)  # rows that look past the box
        assert (p[..., 0] > 0).mean() > 0.05
        # determinism: a second context renders the same bits
        again = ctx.render_frame(cam, 0, 1)
        again2 = ctx.render_frame(cam, 0, 1)
        assert np.array_equal(again, again2)


# ---------------------------------------------------------------- edge cases of the estimator


def _edge_grid(shape, seed):
    """A lumpy cloud in a non-cubic (nz, ny, nx) grid with a zero border voxel (what the importer produces, Resources.cpp:97-101)."""
    rs = np.random.RandomState(seed)
    nz, ny, nx = shape
    z, y, x = np.meshgrid(np.linspace(-1, 1, nz), np.linspace(-1, 1, ny), np.linspace(-1, 1, nx), indexing="ij")
    d = np.clip(1.3 - np.sqrt(x * x + (1.4 * y) ** 2 + z * z) * 1.5 + 0.5 * rs.uniform(-1, 1, size=shape), 0, 1)
    g = (d * 255).astype(np.uint8)
    g[0] = g[-1] = 0
    g[:, 0] = g[:, -1] = 0
    g[:, :, 0] = g[:, :, -1] = 0
    return g


EDGE_CASES = {
    # name: (grid shape (nz, ny, nx), cloud size m, sun, eye)
    "non_cubic_grid": ((40, 24, 56), 7000.0, (-0.586, -0.766, -0.271), (2.5, -0.4, 0.0)),
    "camera_inside_the_cloud": ((32, 32, 32), 3000.0, (0.3, -0.8, 0.52), (0.05, 0.02, -0.03)),  # intersect reports t = 1e-6 (cloudBBox.cu:31)
    "axis_aligned_sun_thick_cloud": ((24, 48, 24), 12000.0, (0.0, -1.0, 0.0), (0.0, 0.3, 2.2)),
    "grazing_sun": ((32, 32, 48), 7000.0, (0.995, -0.0998, 0.0), (2.5, -0.4, 0.0)),  # the C4 sun
}


@pytest.mark.parametrize("name", sorted(EDGE_CASES))
def test_edge_cases_exact_bit_exact_and_fast_consistent(built_library, name):
    """Grids that are not cubes (per-axis texture scale, occupancy cells), a camera inside the box, an axis-aligned sun (Onb branch) and a
    grazing one: EXACT equals the oracle bit for bit, FAST agrees with it statistically and keeps the silhouette."""
    ds = built_library
    shape, size_m, sun, eye = EDGE_CASES[name]
    grid = _edge_grid(shape, 5)
    w, h, spp = 36, 20, 24
    cam = ds.camera_look_at(eye=eye, aspect=w / h)
    cam_np = ds.camera_array(cam)
    o = ol.Oracle()
    o.volume_upload(grid, True)
    o.scene_set(size_m, sun)
    o.bake()
    rp, rv = o.render_accumulate(cam_np, w, h, 0, 1, spp)
    assert rp[..., 0].max() > 0
    with ds.Context(0) as ctx:
        ctx.set_option("precision", ds.PRECISION_EXACT)
        ctx.volume_upload(grid, True)
        ctx.scene_set(size_m, sun)
        ctx.bake()
        assert np.array_equal(ctx.inscatter(), o.inscatter())
        ctx.frame_create(w, h)
        ctx.render_subframes(cam, 0, 1, spp)
        p, v = ctx.frame_download()
        assert np.array_equal(p, rp) and np.array_equal(v, rv)
        ctx.set_option("precision", ds.PRECISION_FAST)
        ctx.bake()
        ctx.frame_clear()
        ctx.counters_reset()
        ctx.render_subframes(cam, 0, 1, spp)
        pf, vf = ctx.frame_download()
        c = ctx.counters()
        # the fused RG8 {density, sun transmittance} volume against the two R8 arrays on this grid: same bits, same counters -- also after the
        # sun moved (the fused array is rebuilt by the next render) and after a sun-transmittance volume uploaded from the host
        def frame_and_counters():
            ctx.frame_clear()
            ctx.counters_reset()
            ctx.render_subframes(cam, 0, 1, 4)
            cc = ctx.counters()
            return ctx.frame_download()[0].view(np.uint32), (cc["paths"], cc["events"], cc["steps"])

        for step in range(3):
            if step == 1:
                ctx.scene_set(size_m, (sun[2], sun[0], sun[1]))
                ctx.bake()
            if step == 2:
                ctx.inscatter_set(np.ascontiguousarray(o.inscatter()[::-1]))
            ctx.set_option("fused_volume", 1)
            f1, c1 = frame_and_counters()
            ctx.set_option("fused_volume", 0)
            f0, c0 = frame_and_counters()
            assert np.array_equal(f1, f0) and c1 == c0, (name, step)
    assert c["nonfinite"] == 0 and np.isfinite(pf).all() and c["paths"] == w * h * spp
    a, b = pf[..., 0].astype(np.float64), rp[..., 0].astype(np.float64)
    sigma = np.sqrt((vf[..., 0].astype(np.float64) + rv[..., 0].astype(np.float64)) / (spp - 1) / spp)
    lit = sigma > 0
    assert (np.abs(a - b)[lit] / sigma[lit] < 3).mean() > 0.98
    # same silhouette; pixels in deep shadow (the u8 sun transmittance truncates to 0 there) may flip between exactly 0 and a tiny value
    flipped = (a == 0) != (b == 0)
    assert flipped.mean() < 0.05 and (flipped.sum() == 0 or np.maximum(a, b)[flipped].max() < 0.05 * b.max())
    assert abs(a.mean() - b.mean()) < 0.01 * b.mean() + 3 * np.sqrt((sigma**2).sum()) / a.size


def test_empty_cloud_renders_black_and_terminates(built_library):
    """An all-zero grid: every tap reads 0, nothing scatters, the bake is fully transparent; both flavours, all three estimators."""
    ds = built_library
    grid = np.zeros((16, 20, 12), np.uint8)
    cam = ds.camera_look_at(aspect=2.0)
    for prec in (ds.PRECISION_EXACT, ds.PRECISION_FAST):
        with ds.Context(0) as ctx:
            ctx.set_option("precision", prec)
            ctx.volume_upload(grid, True)
            ctx.scene_set(7000.0, (-0.03, -0.25, 0.8))
            ctx.bake()
            assert ctx.inscatter().min() == 255
            ctx.frame_create(32, 16)
            for mode in (0, 1, 2):
                ctx.frame_clear()
                ctx.render_subframes(cam, mode, 1, 3)
                p, v = ctx.frame_download()
                assert np.all(p[..., :3] == 0) and np.all(v == 0) and np.all(p[..., 3] == 1)
            c = ctx.counters()
            assert c["events"] == 0 and c["nonfinite"] == 0


@pytest.mark.gpu
def test_fast_direction_sampling_inverts_the_reference_cdf(built_library):
    """k_trace_fast replaces the 16-step bisection of cloud.cuh:167-178 by the closed-form inverse of the same piecewise-linear CDF (two-level
    guide + four fixed probes).  ds_invert_phase_cdf runs exactly that device code: it must land on the float64 inverse of the CDF built
    the way Mie.cpp:8273-8282 builds it, for a dense set of variates, every knot value and the neighbours of every bucket boundary -- and
    the kernel's half-precision phase table must be the chopped sampler within a half's rounding."""
    ds = built_library
    raw = np.fromfile(ROOT / "deepestscatter_b200" / "data" / "mie_tables.f32", dtype=np.float32)
    n = raw.size // 2
    chopped = raw[n:]
    total = np.float32(0)
    for x in chopped:
        total = np.float32(total + x)
    cdf = np.zeros(n, np.float32)
    acc = np.float32(0)
    for i, x in enumerate(chopped):
        acc = np.float32(acc + np.float32(x / total))
        cdf[i] = acc
    rs = np.random.RandomState(5)
    below = np.nextafter(cdf, np.float32(0)).astype(np.float32)
    above = np.nextafter(cdf, np.float32(2)).astype(np.float32)
    edges = np.concatenate([np.arange(2048) / 2048 * 0.125, np.arange(1024) / 1024]).astype(np.float32)
    vals = np.concatenate([rs.random_sample(1 << 20).astype(np.float32), rs.random_sample(1 << 18).astype(np.float32) * np.float32(0.125), cdf, below, above,
                           edges, np.nextafter(edges, np.float32(-1)).astype(np.float32), [0.0, 1e-12, 0.99999994]]).astype(np.float32)
    vals = vals[(vals >= 0) & (vals < 1)]
    with ds.Context(0) as ctx:
        cos_t, phase = ctx.invert_phase_cdf(vals)
    # float64 inverse: first knot i with cdf[i] >= val; u = (i - 0.5 + (val - cdf[i-1]) / (cdf[i] - cdf[i-1])) / n, clamped like tex1D; i = 0 -> u = 0
    i = np.searchsorted(cdf, vals, side="left")
    assert i.max() < n
    pad = np.concatenate([[0.0], cdf.astype(np.float64)])
    a, b = pad[i], pad[i + 1]
    t = (vals.astype(np.float64) - a) / (b - a)
    u = np.where(i == 0, 0.0, (i - 0.5 + t) / n)
    want = 2 * np.minimum(u, 1.0) - 1
    # a float32 quotient of two float32 differences: a few ulp of t (<= 1) inside a knot interval of width 2 / n in cos(theta)
    assert np.abs(cos_t - want).max() <= 4e-7 + 2.0 / n * 2e-6
    # the reference's own bisection (16 halvings of [0, 1] on the tabulated CDF with linear interpolation) agrees within 2^-16 in u
    x = np.minimum(np.maximum(vals.astype(np.float64) * n - 0.5, 0.0), n - 1.0)
    k = x.astype(np.int64)
    want_phase = (chopped.astype(np.float64) / (chopped.astype(np.float64).sum() / n))
    lerp = want_phase[k] + (x - k) * (want_phase[np.minimum(k + 1, n - 1)] - want_phase[k])
    assert np.abs(phase - lerp).max() <= 6e-4 * np.abs(lerp).max() and (np.abs(phase - lerp) <= 1.5e-3 * np.abs(lerp) + 1e-6).all()


@pytest.mark.parametrize("case", ["nonzero_faces", "coarse_step"])
def test_fast_variants_with_a_box_test(built_library, case):
    """k_trace_fast drops the per-step box test and the in-box test of the collision point only when the grid's faces are zero AND the
    sampling step stays within the 0.01 slack of the reference's isInBox (cloud.cuh:40-44).  The two configurations that break a
    premise -- a cloud that fills its grid up to the faces, and a sampling step of 0.02 -- must take the variants that keep the tests:
    frames statistically equal to the oracle's, identical with and without empty-space skipping, EXACT bit-exact as everywhere."""
    ds = built_library
    rs = np.random.RandomState(11)
    if case == "nonzero_faces":
        grid = (rs.uniform(0, 1, size=(24, 28, 20)) ** 3 * 0.3 * 255).astype(np.uint8)  # haze up to the faces, with a hole
        grid[8:16, 10:18, 6:14] = 0
        step, size_m = 1.0 / 512.0, 1500.0
    else:
        grid = _edge_grid((32, 32, 32), 3)
        step, size_m = 0.02, 5000.0
    sun = (-0.586, -0.766, -0.271)
    w, h, spp = 36, 20, 32
    cam = ds.camera_look_at(eye=(2.5, -0.4, 0.3), aspect=w / h)
    o = ol.Oracle()
    o.volume_upload(grid, True)
    o.scene_set(size_m, sun, sample_step=step)
    o.bake()
    rp, rv = o.render_accumulate(ds.camera_array(cam), w, h, 0, 1, spp)
    assert rp[..., 0].max() > 0
    with ds.Context(0) as ctx:
        ctx.volume_upload(grid, True)
        ctx.scene_set(size_m, sun, sample_step=step)
        ctx.set_option("precision", ds.PRECISION_EXACT)
        ctx.bake()
        ctx.frame_create(w, h)
        ctx.render_subframes(cam, 0, 1, spp)
        p, v = ctx.frame_download()
        assert np.array_equal(p, rp) and np.array_equal(v, rv)
        ctx.set_option("precision", ds.PRECISION_FAST)
        ctx.bake()
        frames = {}
        for skip in (1, 0):
            ctx.set_option("skip_empty", skip)
            ctx.frame_clear()
            ctx.counters_reset()
            ctx.render_subframes(cam, 0, 1, spp)
            frames[skip] = (ctx.frame_download(), ctx.counters())
        (pf, vf), c = frames[1]
        assert np.array_equal(pf.view(np.uint32), frames[0][0][0].view(np.uint32))
        assert (c["paths"], c["events"], c["steps"]) == tuple(frames[0][1][k] for k in ("paths", "events", "steps"))
    assert c["nonfinite"] == 0 and np.isfinite(pf).all() and c["paths"] == w * h * spp
    a, b = pf[..., 0].astype(np.float64), rp[..., 0].astype(np.float64)
    sigma = np.sqrt((vf[..., 0].astype(np.float64) + rv[..., 0].astype(np.float64)) / (spp - 1) / spp)
    lit = sigma > 0
    assert (np.abs(a - b)[lit] / sigma[lit] < 3).mean() > 0.98
    assert abs(a.mean() - b.mean()) < 0.01 * b.mean() + 3 * np.sqrt((sigma**2).sum()) / a.size
