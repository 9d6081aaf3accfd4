"""The C++ headless DataGen driver (deepestscatter_b200/host/datagen.cpp + DataGen.hpp): same class shapes and call
order as the reference's Tasks / Scene / collectors, every launch a call into the C ABI."""
import lmdb_compat
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
DATAGEN = ROOT / "deepestscatter_b200" / "datagen"


def run(*args, check=True):
    r = subprocess.run([str(DATAGEN), *map(str, args)], capture_output=True, text=True, timeout=600)
    if check and r.returncode != 0:
        raise AssertionError(f"datagen {' '.join(map(str, args))} failed:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    return r


def test_scenes_merge_stat_are_host_only(built_library, tmp_path):
    ds = built_library
    db = tmp_path / "Train.lmdb"
    run("scenes", db, "--clouds", "a/one.npy,b/two.npy", "--scenes-per-cloud", 3, "--seed", 7)
    out = run("stat", db).stdout
    assert "SceneSetup 6" in out
    env = lmdb_compat.Environment(str(db), subdir=False, readonly=True, max_dbs=8)
    scenes = env.open_db(b"SceneSetup", integerkey=True)
    with env.begin(db=scenes) as t:
        recs = [bytes(v) for _, v in t.cursor(scenes)]
    assert len(recs) == 6 and recs[0][:2] == b"\x0a\x09" and recs[0][2:11] == b"a/one.npy" and recs[5][2:11] == b"b/two.npy"
    sizes = [struct.unpack_from("<f", r, 12)[0] for r in recs]  # field 2 follows the 9-byte path: tag 0x15
    assert all(r[11] == 0x15 for r in recs) and all(1000.0 <= s <= 12000.0 for s in sizes)
    other = tmp_path / "rank1.lmdb"
    with ds.Dataset(other) as w:
        w.append_results(2048, [0.5, 0.25], [1, 1])
    run("merge", db, other)
    assert "Result 2" in run("stat", db).stdout and lmdb_compat.check(str(db))["pages_leaked"] == 0
    assert run("collect", check=False).returncode == 1
    assert run("render", "cloud.vdb", check=False).returncode == 1  # no device here, or an unsupported cloud file: loud either way


def test_npy_importer_crops_to_the_active_box(built_library, tmp_path):
    """CloudImporter: .npy -> active bounding box expanded by one voxel (Resources.cpp:97-101); checked through datagen's error text."""
    bad = tmp_path / "flat.npy"
    np.save(bad, np.zeros((4, 4, 4), np.float32))
    db = tmp_path / "s.lmdb"
    run("scenes", db, "--clouds", str(bad), "--scenes-per-cloud", 1)
    r = run("collect", db, "--what", "samples", check=False)
    assert r.returncode == 1 and ("no active" in r.stderr or "no CUDA device" in r.stderr)


@pytest.mark.gpu
def test_collect_matches_the_python_binding(built_library, tmp_path):
    ds = built_library
    n, batch = 48, 64
    grid = np.zeros((n + 6, n + 4, n + 2), np.float32)  # (nz, ny, nx) with a zero margin the importer must crop
    with ds.Context(0) as ctx:
        ctx.volume_synth(n, 0, 1234)
        u8 = ctx.level(0)
    core = u8[1:-1, 1:-1, 1:-1].astype(np.float32) / 255.0 * 3.0  # the synthetic grids carry a zero border voxel
    assert u8[0].max() == 0 and u8[:, 0].max() == 0 and u8[:, :, -1].max() == 0 and core.max() > 0
    grid[4:4 + n - 2, 3:3 + n - 2, 1:1 + n - 2] = core
    cloud = tmp_path / "cumulus.npy"
    np.save(cloud, grid)
    db = tmp_path / "Train.lmdb"
    run("scenes", db, "--clouds", "cumulus.npy", "--scenes-per-cloud", 2, "--seed", 11)
    run("collect", db, "--what", "all", "--cloud-root", tmp_path, "--batch-size", batch, "--max-threads", 20480, "--launches", 100,
        "--opt", "radiance_scheduler=0")  # the reference update loop: deterministic, comparable bit for bit
    rep = lmdb_compat.check(str(db))
    assert {k: v["entries"] for k, v in rep["tables"].items()} == {"SceneSetup": 2, "ScatterSample": 2 * batch, "DisneyDescriptor": 2 * batch, "Result": 2 * batch,
                                                                   "BakedInterpolationSet": 0}  # the fifth table LmdbDataset.py opens, empty
    # continue mode: nothing left to do, the file does not change
    before = db.read_bytes()
    run("collect", db, "--what", "all", "--cloud-root", tmp_path, "--batch-size", batch)
    assert db.read_bytes() == before

    env = lmdb_compat.Environment(str(db), subdir=False, readonly=True, max_dbs=8)
    tabs = {k: env.open_db(k.encode(), integerkey=True) for k in rep["tables"]}

    def get(table, i):
        with env.begin(db=tabs[table]) as t:
            return t.get(i.to_bytes(4, "little"))

    with ds.Context(0) as ctx:
        ctx.set_option("radiance_scheduler", 0)
        # the importer must have reproduced the quantised grid: crop to the active box + one voxel of padding
        active = np.argwhere(core > 0)
        lo, hi = active.min(0), active.max(0)
        expect = np.zeros(tuple(hi - lo + 3), np.float32)
        expect[1:-1, 1:-1, 1:-1] = core[lo[0]:hi[0] + 1, lo[1]:hi[1] + 1, lo[2]:hi[2] + 1]
        ctx.volume_upload_float(expect, float(core.max()))
        for scene in range(2):
            rec = get("SceneSetup", scene)
            size = struct.unpack_from("<f", rec, 2 + rec[1] + 1)[0]
            lrec = rec[2 + rec[1] + 5 + 2:]
            light = [struct.unpack_from("<f", lrec, 1 + 5 * k)[0] for k in range(3)]
            ctx.scene_set(size, light)
            ctx.bake()
            pos, dirs = ctx.generate_points(0, batch, stream=scene)
            desc = ctx.descriptors(pos, dirs)
            for i in (0, 1, batch - 1):
                assert get("ScatterSample", scene * batch + i) == ds.record_scatter_sample(pos[i], dirs[i])
                assert get("DisneyDescriptor", scene * batch + i) == ds.record_disney_descriptor(desc[i].tobytes())
            tasks, conv, nconv, _ = ctx.point_radiance(pos, dirs, max_threads=20480, launches_per_update=100)
            for i in (0, batch // 2, batch - 1):
                assert get("Result", scene * batch + i) == ds.record_result(float(tasks["radiance"][i]), True)


@pytest.mark.gpu
def test_render_task_writes_the_progressive_image(built_library, tmp_path):
    ds = built_library
    run("render", "synth:48", "--width", 64, "--height", 32, "--spp", 20, "--out", tmp_path, "--size", 7000)
    files = sorted(p.name for p in tmp_path.iterdir())
    assert files == ["synth_48.Back.PathTracing.exr", "synth_48.Back.PathTracing.exr.ppm", "synth_48.Side.PathTracing.exr", "synth_48.Side.PathTracing.exr.ppm"]
    from exr_reader import read_exr

    e = read_exr(tmp_path / "synth_48.Side.PathTracing.exr")
    assert (e["width"], e["height"], e["line_order"]) == (64, 32, 1)
    img = np.stack([e["image"][c] for c in "RGB"], axis=-1)
    with ds.Context(0) as ctx:
        ctx.volume_synth(48, 0, 1234)
        ctx.scene_set(7000.0, (-0.03, -0.25, 0.8))
        ctx.bake()
        ctx.frame_create(64, 32)
        ctx.render_subframes(ds.camera_look_at(aspect=2.0), ds.MODE_ALL_SCATTER, 1, 20)
        p, _ = ctx.frame_download()
    assert np.array_equal(img, p[..., :3])


@pytest.mark.gpu
def test_render_task_with_the_neural_renderer(built_library, tmp_path):
    """`using TRenderer = DisneyRenderer` (Tasks.cpp:86): the C++ driver loads the flat weight file and accumulates neural subframes;
    the EXR equals the same subframes driven through the Python binding."""
    ds = built_library
    from deepestscatter_b200 import disney_model as dm

    w = dm.synthetic_weights(566)
    model = tmp_path / "DisneyModel.f32"
    w.tofile(model)
    run("render", "synth:48", "--width", 160, "--height", 40, "--spp", 3, "--out", tmp_path, "--size", 7000, "--renderer", "disney", "--model", model)
    assert (tmp_path / "synth_48.Side.Disney.exr").exists() and (tmp_path / "synth_48.Back.Disney.exr").exists()
    from exr_reader import read_exr

    e = read_exr(tmp_path / "synth_48.Side.Disney.exr")
    img = np.stack([e["image"][c] for c in "RGB"], axis=-1)
    cam = ds.camera_look_at(aspect=4.0)
    with ds.Context(0) as ctx:
        ctx.volume_synth(48, 0, 1234)
        ctx.scene_set(7000.0, (-0.03, -0.25, 0.8))
        ctx.bake()
        ctx.disney_model_load(w)
        ctx.frame_create(160, 40)
        ctx.render_disney_subframes(cam, 1, 3)
        p, v = ctx.frame_download()
        # the accumulation is Welford over the three neural frames (subframe s draws from stream s * 4096)
        frames = [ctx.render_disney(cam, 160, 40, stream=s * 4096) for s in (1, 2, 3)]
    assert np.array_equal(img, p[..., :3]) and img.max() > 0
    mean = np.zeros_like(frames[0])
    for k, f in enumerate(frames, start=1):
        mean = mean + (f - mean) * np.float32(1.0 / k)
    assert np.abs(mean - p).max() <= 1e-6 * max(1.0, float(np.abs(p).max()))
    assert run("render", "synth:48", "--renderer", "disney", check=False).returncode == 1  # no --model


@pytest.mark.gpu
def test_collect_commits_every_batch_and_survives_an_interrupt(built_library, tmp_path):
    """One LMDB commit per finished batch (the reference's transaction per batchAppend): SIGINT ends the run between batches with
    exit code 130 and a consistent file holding every finished batch; --mode continue then completes the dataset."""
    import signal
    import time

    batch, scenes = 64, 6
    db = tmp_path / "Train.lmdb"
    run("scenes", db, "--clouds", "synth:64", "--scenes-per-cloud", scenes, "--seed", 3)
    args = [str(DATAGEN), "collect", str(db), "--what", "results", "--batch-size", str(batch), "--opt", "radiance_scheduler=1"]
    run("collect", db, "--what", "samples", "--batch-size", batch)
    assert lmdb_compat.check(str(db))["txnid"] >= scenes  # one commit per batch, not one per phase
    p = subprocess.Popen(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    deadline = time.time() + 120
    while time.time() < deadline and p.poll() is None:  # wait until at least one Result batch is durable, then interrupt
        try:
            if lmdb_compat.check(str(db))["tables"].get("Result", {}).get("entries", 0) >= batch:
                break
        except Exception:  # the writer may be in the middle of a commit
            pass
        time.sleep(0.05)
    p.send_signal(signal.SIGINT)
    out, err = p.communicate(timeout=300)
    rep = lmdb_compat.check(str(db))
    done = rep["tables"]["Result"]["entries"]
    assert rep["pages_leaked"] == 0 and done % batch == 0 and done >= batch
    if done < scenes * batch:  # the signal arrived before the last batch finished
        assert p.returncode == 130 and "continue" in err
    run("collect", db, "--what", "results", "--batch-size", batch, "--mode", "continue")
    rep = lmdb_compat.check(str(db))
    assert rep["tables"]["Result"]["entries"] == scenes * batch and rep["pages_leaked"] == 0
