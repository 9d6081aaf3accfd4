"""Dataset store: the LMDB data file written by ds_dataset_* (deepestscatter_b200/host/LmdbFile.hpp) read back through an
independent pure-Python parser with py-lmdb's API (tests/lmdb_compat.py), the way
DeepestScatter_Train/LmdbDataset.py:24-66 reads the reference's datasets.  No GPU needed.

Neither liblmdb nor py-lmdb exists here, so byte-level format parity is unpinned against liblmdb; what IS pinned to the real library:
the table statistics (depth, branch / leaf / overflow pages, entries) that the reference's DatasetVisualisation.ipynb recorded for the
authors' dataset, reproduced exactly at full size; and to the reference: its own reader and training dataset classes, unmodified,
decode a file written here.  The rest pins the writer against the restated format, the record bytes against the golden vectors, and
the structural invariants mdb.c relies on."""
import json
import os
import struct
from pathlib import Path

import numpy as np
import pytest

import lmdb_compat

GOLDEN = json.loads((Path(__file__).parent / "golden" / "records.json").read_text())
BATCH_SIZE = 2048


class LmdbDatasetMirror:
    """DeepestScatter_Train/LmdbDataset.py, with lmdb -> lmdb_compat and protobuf messages -> raw bytes."""

    def __init__(self, ds, path):
        self.env = lmdb_compat.Environment(str(path), map_size=3e9, subdir=False, max_dbs=64, mode=0, create=False, readonly=True)
        self.dbs = {}

    def db(self, name):
        if name not in self.dbs:
            self.dbs[name] = self.env.open_db(name.encode("ascii"), integerkey=True, create=False)
        return self.dbs[name]

    def getCountOf(self, name):
        with self.env.begin() as transaction:
            return transaction.stat(self.db(name))["entries"]

    def get(self, name, id, buffers=False):
        db = self.db(name)
        with self.env.begin(db=db, buffers=buffers) as transaction:
            return transaction.get(id.to_bytes(4, "little"), db=db)


def parse_vector3(b):
    v = [0.0, 0.0, 0.0]
    i = 0
    while i < len(b):
        tag = b[i]
        i += 1
        assert tag & 7 == 5
        v[(tag >> 3) - 1] = struct.unpack_from("<f", b, i)[0]
        i += 4
    return v


def parse_scatter_sample(b):
    out = {}
    i = 0
    while i < len(b):
        tag = b[i]
        n = b[i + 1]
        out[tag >> 3] = parse_vector3(b[i + 2: i + 2 + n])
        i += 2 + n
    return out[2], out[3]


def synth(n, seed):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-0.5, 0.5, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    desc = rng.integers(0, 256, (n, 2250), dtype=np.uint8)
    rad = rng.uniform(0, 3, n).astype(np.float32)
    rad[::17] = 0.0
    return pos, d, desc, rad


def test_dataset_written_here_reads_like_the_reference_reader(built_library, tmp_path):
    ds = built_library
    path = tmp_path / "Train.lmdb"
    n = 3 * BATCH_SIZE
    pos, d, desc, rad = synth(n, 1)
    with ds.Dataset(path) as w:
        for scene in range(3):
            w.append_scene_setup(scene, f"Clouds/cloud{scene}.vdb", 7000.0 + scene, (-0.03, -0.25, 0.8))
            s = slice(scene * BATCH_SIZE, (scene + 1) * BATCH_SIZE)
            w.append_scatter_samples(scene * BATCH_SIZE, pos[s], d[s])
            w.append_descriptors(scene * BATCH_SIZE, desc[s])
            w.append_results(scene * BATCH_SIZE, rad[s], np.ones(BATCH_SIZE, np.uint8))
            w.commit()  # one transaction per batch, Dataset.h:203-232
        assert w.count("ScatterSample") == n and w.count("SceneSetup") == 3

    report = lmdb_compat.check(str(path))
    assert report["pages_leaked"] == 0
    assert report["tables"]["DisneyDescriptor"]["overflow_pages"] == n  # 2253-byte values: one overflow page each
    r = LmdbDatasetMirror(ds, path)
    assert r.getCountOf("ScatterSample") == n and r.getCountOf("DisneyDescriptor") == n and r.getCountOf("Result") == n
    assert r.getCountOf("SceneSetup") == 3
    for i in (0, 1, 2047, 2048, 4095, n - 1, 777, 5000):
        p, v = parse_scatter_sample(r.get("ScatterSample", i))
        assert np.array_equal(np.float32(p), pos[i]) and np.array_equal(np.float32(v), d[i])
        g = r.get("DisneyDescriptor", i, buffers=True)
        assert len(g) == 2253 and bytes(g[:3]) == b"\x0a\xca\x11"
        # DisneyDataset.py:26-28: bytes / 256 viewed (10, -1)
        grid = np.frombuffer(g, np.uint8, offset=3).astype(np.float32) / 256
        assert grid.reshape(10, -1).shape == (10, 225) and np.array_equal(np.frombuffer(g, np.uint8, offset=3), desc[i])
        assert r.get("Result", i) == ds.record_result(float(rad[i]), True)
        scene = r.get("SceneSetup", i // BATCH_SIZE)  # BaseDataset.py:32
        assert scene == ds.record_scene_setup(f"Clouds/cloud{i // BATCH_SIZE}.vdb", 7000.0 + i // BATCH_SIZE, (-0.03, -0.25, 0.8))
    assert r.get("Result", n) is None and r.get("Result", 2**31 - 1) is None
    # LmdbDataset.py:36-40 opens five tables with create=False: the one this library never fills exists, empty; any other name is MDB_NOTFOUND
    assert r.getCountOf("BakedInterpolationSet") == 0 and r.get("BakedInterpolationSet", 0) is None
    with pytest.raises(lmdb_compat.Error):
        r.env.open_db(b"LightProbes", integerkey=True)
    # cursor order = key order (LmdbDataset.getCountBeforeLastFlatCloud iterates SceneSetup)
    with r.env.begin() as t:
        keys = [int.from_bytes(k, "little") for k, _ in t.cursor(r.db("SceneSetup"))]
    assert keys == [0, 1, 2]


def test_golden_record_bytes_survive_the_store(built_library, tmp_path):
    ds = built_library
    path = tmp_path / "g.lmdb"
    with ds.Dataset(path) as w:
        for i, g in enumerate(GOLDEN["scatter_sample"]):
            w.append_scatter_samples(i, [g["point"]], [g["view_direction"]])
        for i, g in enumerate(GOLDEN["result"]):
            w.append_results(i, [g["light_intensity"]], [1 if g["is_converged"] else 0])
    r = LmdbDatasetMirror(ds, path)
    for i, g in enumerate(GOLDEN["scatter_sample"]):
        assert r.get("ScatterSample", i).hex() == g["hex"]
    for i, g in enumerate(GOLDEN["result"]):
        assert r.get("Result", i).hex() == g["hex"]


def test_meta_page_bytes(built_library, tmp_path):
    """mdb_env_init_meta / mdb_env_write_meta layout at fixed offsets."""
    ds = built_library
    path = tmp_path / "m.lmdb"
    ds.Dataset(path).close()  # nothing appended: both metas still txn 0
    raw = path.read_bytes()
    assert len(raw) == 2 * 4096
    for i in range(2):
        page = raw[i * 4096:(i + 1) * 4096]
        assert struct.unpack_from("<QHH", page, 0) == (i, 0, 0x08)  # pgno, pad, P_META
        magic, version, address, mapsize = struct.unpack_from("<IIQQ", page, 16)
        assert (magic, version, address, mapsize) == (0xBEEFC0DE, 1, 0, 1048576)
        psize, flags, depth = struct.unpack_from("<IHH", page, 40)
        assert psize == 4096 and flags == 0x4008 and depth == 0  # MDB_NOSUBDIR | MDB_INTEGERKEY on the free DB
        assert struct.unpack_from("<Q", page, 40 + 40)[0] == 2**64 - 1  # free root P_INVALID
        assert struct.unpack_from("<Q", page, 88 + 40)[0] == 2**64 - 1  # main root P_INVALID
        assert struct.unpack_from("<QQ", page, 136) == (1, 0)  # last_pg, txnid
    assert lmdb_compat.check(str(path))["tables"] == {}
    with ds.Dataset(path) as w:
        w.put("Result", 0, b"\x10\x01")
    raw = path.read_bytes()
    assert struct.unpack_from("<Q", raw, 4096 + 144)[0] == 1  # the first commit goes to meta page 1
    assert struct.unpack_from("<Q", raw, 144)[0] == 0


def test_continue_mode_appends_and_frees_old_tree_pages(built_library, tmp_path):
    ds = built_library
    path = tmp_path / "c.lmdb"
    pos, d, desc, rad = synth(5000, 2)
    with ds.Dataset(path) as w:
        w.append_scatter_samples(0, pos[:3000], d[:3000])
        w.append_descriptors(0, desc[:100])
    first = lmdb_compat.check(str(path))
    assert first["txnid"] == 1 and first["pages_free"] == 0
    with ds.Dataset(path) as w:  # CollectMode::Continue
        assert w.count("ScatterSample") == 3000
        w.append_scatter_samples(3000, pos[3000:], d[3000:])
        w.append_descriptors(100, desc[100:200])
        w.put("DisneyDescriptor", 5, ds.record_disney_descriptor(bytes(desc[4999])))  # replace a big value
    second = lmdb_compat.check(str(path))
    assert second["txnid"] == 2 and second["pages_free"] > 0 and second["pages_leaked"] == 0
    assert second["tables"]["ScatterSample"]["entries"] == 5000 and second["tables"]["DisneyDescriptor"]["entries"] == 200
    r = LmdbDatasetMirror(ds, path)
    assert np.array_equal(np.frombuffer(r.get("DisneyDescriptor", 5), np.uint8, offset=3), desc[4999])
    assert np.array_equal(np.frombuffer(r.get("DisneyDescriptor", 150), np.uint8, offset=3), desc[150])
    p, v = parse_scatter_sample(r.get("ScatterSample", 4999))
    assert np.array_equal(np.float32(p), pos[4999])
    with ds.Dataset(path) as w:
        w.drop("DisneyDescriptor")  # mdb_drop(dbi, 0): empty, not deleted
        assert w.count("DisneyDescriptor") == 0
    third = lmdb_compat.check(str(path))
    assert third["tables"]["DisneyDescriptor"]["entries"] == 0 and third["pages_leaked"] == 0
    assert third["pages_free"] >= second["pages_free"] + 200
    with ds.Dataset(path) as w:  # a session that writes nothing leaves the file alone
        pass
    assert lmdb_compat.check(str(path))["txnid"] == third["txnid"]


def test_three_level_tree_and_shard_merge(built_library, tmp_path):
    ds = built_library
    n = 120_000
    rng = np.random.default_rng(3)
    rad = rng.uniform(0, 2, n).astype(np.float32)
    a, b, m = tmp_path / "rank0.lmdb", tmp_path / "rank1.lmdb", tmp_path / "merged.lmdb"
    with ds.Dataset(a) as w:
        w.append_results(0, rad[: n // 2], np.ones(n // 2, np.uint8))
    with ds.Dataset(b) as w:
        w.append_results(n // 2, rad[n // 2:], np.ones(n - n // 2, np.uint8))
    with ds.Dataset(m) as w:
        w.merge(str(b))  # shards may arrive in any order
        w.merge(str(a))
    rep = lmdb_compat.check(str(m))
    assert rep["tables"]["Result"]["entries"] == n and rep["tables"]["Result"]["depth"] == 3
    r = LmdbDatasetMirror(ds, m)
    for i in list(range(0, n, 997)) + [n - 1, n // 2 - 1, n // 2]:
        assert r.get("Result", i) == ds.record_result(float(rad[i]), True), i
    with r.env.begin() as t:
        keys = np.fromiter((int.from_bytes(k, "little") for k, _ in t.cursor(r.db("Result"))), np.int64)
    assert np.array_equal(keys, np.arange(n))


def test_error_paths(built_library, tmp_path):
    ds = built_library
    bad = tmp_path / "bad.lmdb"
    bad.write_bytes(b"not an lmdb file" * 1000)
    with pytest.raises(ds.DsError) as e:
        ds.Dataset(bad)
    assert "MDB_INVALID" in str(e.value)
    with pytest.raises(lmdb_compat.Error):
        lmdb_compat.Environment(str(bad), subdir=False, readonly=True)
    with pytest.raises(ds.DsError):
        ds.Dataset(tmp_path / "no_such_dir" / "x.lmdb")
    with ds.Dataset(tmp_path / "e.lmdb") as w:
        with pytest.raises(ds.DsError) as e:
            w.get("Result", 3)
        assert "MDB_NOTFOUND" in str(e.value)
        assert w.count("Nothing") == 0
    with pytest.raises(lmdb_compat.Error):
        lmdb_compat.Environment(str(tmp_path / "e.lmdb"), subdir=False, readonly=False)


REF_TRAIN = Path("/root/reference/DeepestScatter_Train")


@pytest.mark.skipif(not (REF_TRAIN / "LmdbDataset.py").is_file(), reason="the reference tree is not mounted here")
def test_the_references_own_reader_opens_and_decodes_the_file(built_library, tmp_path, monkeypatch):
    """DeepestScatter_Train/LmdbDataset.py itself, unmodified, on a file written here: `lmdb` is the one module it needs that this image
    lacks, so the pure-Python reader with py-lmdb's API stands in for it; the protobuf messages are the reference's own."""
    import importlib
    import os
    import sys

    ds = built_library
    path = tmp_path / "Train.lmdb"
    pos, d, desc, rad = synth(5, 3)
    with ds.Dataset(path) as w:
        w.append_scene_setup(0, "RoundClouds/cloud 01.vdb", 7000.0, (-0.03, -0.25, 0.8))
        w.append_scatter_samples(0, pos, d)
        w.append_descriptors(0, desc)
        w.append_results(0, rad, np.ones(5, np.uint8))
    monkeypatch.setenv("PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION", "python")
    monkeypatch.syspath_prepend(str(REF_TRAIN))
    monkeypatch.syspath_prepend(str(REF_TRAIN / "PythonProtocols"))
    monkeypatch.setitem(sys.modules, "lmdb", lmdb_compat)
    sys.modules.pop("LmdbDataset", None)
    try:
        L = importlib.import_module("LmdbDataset")
    except Exception as exc:  # protobuf runtime incompatible with the generated modules
        pytest.skip(f"cannot import the reference's LmdbDataset.py: {exc}")
    try:
        data = L.LmdbDataset(str(path))
        assert data.getCountOf(L.ScatterSample) == 5 and data.getCountOf(L.Result) == 5 and data.getCountOf(L.SceneSetup) == 1
        assert data.getCountOf(L.DisneyDescriptor) == 5 and data.getCountOf(L.BakedInterpolationSet) == 0
        for i in range(5):
            s_ = data.get(L.ScatterSample, i)
            assert np.array_equal(np.float32([s_.point.x, s_.point.y, s_.point.z]), pos[i])
            assert np.array_equal(np.float32([s_.view_direction.x, s_.view_direction.y, s_.view_direction.z]), d[i])
            assert bytes(data.get(L.DisneyDescriptor, i).grid) == desc[i].tobytes()
            res = data.get(L.Result, i)
            assert np.float32(res.light_intensity) == rad[i] and res.is_converged
        scene = data.get(L.SceneSetup, 0)
        assert scene.cloud_path == "RoundClouds/cloud 01.vdb" and scene.cloud_size_m == 7000.0
        assert data.getCountBeforeLastFlatCloud() == 0  # iterates the SceneSetup cursor (LmdbDataset.py:70-80)
    finally:
        sys.modules.pop("LmdbDataset", None)


@pytest.mark.skipif(not (REF_TRAIN / "Disney" / "DisneyDataset.py").is_file(), reason="the reference tree is not mounted here")
def test_the_references_training_dataset_consumes_the_file(built_library, tmp_path, monkeypatch):
    """The whole consumer chain of the training side, unmodified: LmdbDataset.py -> Common/BaseDataset.py -> Disney/DisneyDataset.py turn a
    record set written here into the (10 x 226 descriptor-with-angle, light) pairs the reference's DisneyModel trains on."""
    import importlib
    import math
    import sys

    torch = pytest.importorskip("torch")
    ds = built_library
    n = BATCH_SIZE + 3  # samples of two scenes: scene id = index // 2048 (BaseDataset.py:32)
    pos, d, desc, rad = synth(n, 8)
    lights = [(-0.03, -0.25, 0.8), (0.586, -0.766, -0.271)]
    path = tmp_path / "Train.lmdb"
    with ds.Dataset(path) as w:
        for scene, light in enumerate(lights):
            w.append_scene_setup(scene, f"Clouds/c{scene}.vdb", 3000.0 * (scene + 1), light)
        w.append_scatter_samples(0, pos, d)
        w.append_descriptors(0, desc)
        w.append_results(0, rad, np.ones(n, np.uint8))
    monkeypatch.setenv("PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION", "python")
    for sub in ("", "PythonProtocols", "Common", "Disney"):
        monkeypatch.syspath_prepend(str(REF_TRAIN / sub))
    monkeypatch.setitem(sys.modules, "lmdb", lmdb_compat)
    if not hasattr(np, "math"):
        monkeypatch.setattr(np, "math", math, raising=False)  # Common/Vector.py:20 predates numpy 2
    for name in ("LmdbDataset", "BaseDataset", "DisneyDataset", "Vector"):
        sys.modules.pop(name, None)
    try:
        L = importlib.import_module("LmdbDataset")
        D = importlib.import_module("DisneyDataset")
    except Exception as exc:
        pytest.skip(f"cannot import the reference's dataset modules: {exc}")
    try:
        data = D.DisneyDataset(L.LmdbDataset(str(path)))
        assert len(data) == n
        for i in (0, 1, 2047, 2048, n - 1):
            z, light = data[i]
            assert tuple(z.shape) == (10, 226) and z.dtype == torch.float32
            assert np.array_equal(z[:, :225].numpy(), (desc[i].astype(np.float32) / 256).reshape(10, 225))  # DisneyDataset.py:26-28
            l = np.float32(lights[i // BATCH_SIZE]).astype(np.float64)
            v = d[i].astype(np.float64)
            angle = math.acos(np.dot(l / np.linalg.norm(l), v / np.linalg.norm(v)))
            assert np.allclose(z[:, 225].numpy(), angle, atol=1e-6)
            assert np.float32(light) == rad[i]
    finally:
        for name in ("LmdbDataset", "BaseDataset", "DisneyDataset", "Vector"):
            sys.modules.pop(name, None)


def test_page_counts_match_the_statistics_recorded_in_the_references_notebook(built_library, tmp_path):
    """DeepestScatter_Train/DatasetVisualisation.ipynb keeps the output of `transaction.stat(results_db)` on the authors' dataset, written by
    the real liblmdb: {'psize': 4096, 'depth': 2, 'branch_pages': 1, 'leaf_pages': 78, 'overflow_pages': 0, 'entries': 14336}.  The same
    number of Result records (7 bytes each: non-zero radiance, converged) appended in key order here must pack into the same tree.
    (The ScatterSample line of that output -- 8 663 040 entries in 103 132 leaves = 84 per leaf where 85 fit -- is not reproduced: fill
    factor is the writer's choice and no reader depends on it.)"""
    ds = built_library
    path = tmp_path / "r.lmdb"
    n = 14336
    rad = np.linspace(0.5, 3.0, n).astype(np.float32)
    with ds.Dataset(path) as w:
        for start in range(0, n, BATCH_SIZE):  # one transaction per batch, as the collectors commit
            w.append_results(start, rad[start:start + BATCH_SIZE], np.ones(BATCH_SIZE, np.uint8))
            w.commit()
    env = lmdb_compat.Environment(str(path), subdir=False, readonly=True, max_dbs=8)
    db = env.open_db(b"Result", integerkey=True)
    with env.begin() as txn:
        stat = txn.stat(db)
    assert stat == {"psize": 4096, "depth": 2, "branch_pages": 1, "leaf_pages": 78, "overflow_pages": 0, "entries": 14336}


def test_full_size_statistics_of_the_references_dataset(built_library):
    """The ScatterSample line of the same notebook output, at full size: 8 663 040 records of 34 bytes must pack into depth 4, 359 branch and
    103 132 leaf pages, as the real liblmdb left them (84 records per leaf where 85 fit: mdb_page_split moves the last node of a full page to
    the new one during ascending inserts).  ~10 s, ~3 GB of RAM, a 420 MB scratch file."""
    import importlib
    import sys

    psutil = pytest.importorskip("psutil")
    if psutil.virtual_memory().available < 8 << 30:
        pytest.skip("needs 8 GB of free memory")
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    tool = importlib.import_module("lmdb_notebook_stats")
    assert tool.main() == 0


def _table_stats(path, names=("ScatterSample", "DisneyDescriptor", "Result")):
    env = lmdb_compat.Environment(str(path), subdir=False, readonly=True, max_dbs=8)
    out = {}
    with env.begin() as txn:
        for n in names:
            out[n] = txn.stat(env.open_db(n.encode(), integerkey=True))
    env.close()
    return out


def test_commit_per_batch_is_incremental_and_equals_the_one_shot_file(built_library, tmp_path):
    """The collectors commit after every batch of 2048 records (the reference's transaction per batchAppend).  The per-batch file must
    hold the same trees as a file written in one go (same depth / branch / leaf / overflow page counts, same records), leak no page,
    and must not grow with the number of commits: the tail of an appended table is rewritten, its completed leaves are not."""
    ds = built_library
    rng = np.random.default_rng(3)
    batches, batch = 24, 2048
    pos = rng.normal(size=(batches * batch, 3)).astype(np.float32)
    dirs = rng.normal(size=(batches * batch, 3)).astype(np.float32)
    desc = rng.integers(0, 256, (batches * batch, 2250), dtype=np.uint8)
    rad = rng.random(batches * batch).astype(np.float32)
    one, many = tmp_path / "one.lmdb", tmp_path / "many.lmdb"
    with ds.Dataset(str(one)) as d:
        d.append_scatter_samples(0, pos, dirs)
        d.append_descriptors(0, desc)
        d.append_results(0, rad, np.ones(len(rad), np.uint8))
    sizes = []
    with ds.Dataset(str(many)) as d:
        for b in range(batches):
            sl = slice(b * batch, (b + 1) * batch)
            d.append_scatter_samples(b * batch, pos[sl], dirs[sl])
            d.commit()
            sizes.append(many.stat().st_size)
        for b in range(batches):
            sl = slice(b * batch, (b + 1) * batch)
            d.append_descriptors(b * batch, desc[sl])
            d.commit()
        for b in range(batches):
            sl = slice(b * batch, (b + 1) * batch)
            d.append_results(b * batch, rad[sl], np.ones(batch, np.uint8))
            d.commit()
    assert _table_stats(one) == _table_stats(many)
    rep_one, rep_many = lmdb_compat.check(str(one)), lmdb_compat.check(str(many))
    assert rep_one["pages_leaked"] == 0 and rep_many["pages_leaked"] == 0
    assert rep_many["txnid"] >= 3 * batches
    # 72 commits later the file is at most a few percent larger than the one-shot file (freed tree pages are reused)
    assert many.stat().st_size < 1.05 * one.stat().st_size + (1 << 20)
    # the ScatterSample phase: ~25 leaf pages of data per batch; a full rewrite per commit would add the whole table every time
    growth = np.diff(sizes)
    assert growth.max() < 3 * (sizes[-1] // batches)
    env = lmdb_compat.Environment(str(many), subdir=False, readonly=True, max_dbs=8)
    with env.begin() as txn:
        dbd = env.open_db(b"DisneyDescriptor", integerkey=True)
        for k in (0, 2047, 2048, batches * batch - 1):
            assert txn.get(int(k).to_bytes(4, "little"), db=dbd)[3:] == desc[k].tobytes()
    env.close()
    # reopening and appending continues incrementally (Tasks.h:65-68: Continue mode)
    with ds.Dataset(str(many)) as d:
        assert d.count("ScatterSample") == batches * batch
        d.append_scatter_samples(batches * batch, pos[:batch], dirs[:batch])
    assert lmdb_compat.check(str(many))["pages_leaked"] == 0
    assert _table_stats(many, ("ScatterSample",))["ScatterSample"]["entries"] == (batches + 1) * batch


def test_overwrite_after_commit_rebuilds_the_table(built_library, tmp_path):
    ds = built_library
    path = tmp_path / "rw.lmdb"
    with ds.Dataset(str(path)) as d:
        for k in range(500):
            d.put("Result", k, bytes([k % 251, 1, 2]))
        d.commit()
        d.put("Result", 7, b"changed")  # not an append
        d.put("Result", 500, b"tail")
        d.commit()
        assert d.get("Result", 7) == b"changed" and d.get("Result", 499) == bytes([499 % 251, 1, 2])
    rep = lmdb_compat.check(str(path))
    assert rep["pages_leaked"] == 0 and rep["tables"]["Result"]["entries"] == 501
    env = lmdb_compat.Environment(str(path), subdir=False, readonly=True, max_dbs=8)
    with env.begin() as txn:
        db = env.open_db(b"Result", integerkey=True)
        assert txn.get((7).to_bytes(4, "little"), db=db) == b"changed"
        assert txn.get((500).to_bytes(4, "little"), db=db) == b"tail"
    env.close()


def test_merge_source_is_opened_read_only_and_must_exist(built_library, tmp_path):
    ds = built_library
    shard = tmp_path / "shard.lmdb"
    with ds.Dataset(str(shard)) as d:
        d.put("Result", 3, b"abc")
    os.chmod(shard, 0o444)  # a read-only shard merges
    out = tmp_path / "out.lmdb"
    with ds.Dataset(str(out)) as d:
        d.merge(str(shard))
        assert d.get("Result", 3) == b"abc"
        with pytest.raises(ds.DsError):
            d.merge(str(tmp_path / "mistyped.lmdb"))
    assert not (tmp_path / "mistyped.lmdb").exists()  # and is not created as an empty dataset
