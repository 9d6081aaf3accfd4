"""Property test of the LMDB data-file writer (host/LmdbFile.hpp behind ds_dataset_*): arbitrary interleavings of puts, overwrites,
commits and re-opens over several integer-keyed tables must always leave a file that the independent reader (lmdb_compat, the py-lmdb
API the reference's LmdbDataset.py uses) decodes to exactly the model dictionary, with the structural audit clean (no leaked or
doubly-used pages, keys sorted, overflow chains of the right length).  Value sizes straddle the leaf / overflow-page boundary."""
import lmdb_compat
import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

TABLES = ["SceneSetup", "ScatterSample", "DisneyDescriptor", "Result"]
# page size 4096: values up to ~2 KB stay in the leaf, larger ones go to overflow pages (mdb.c: nodemax)
SIZES = st.one_of(st.integers(0, 40), st.integers(1990, 2100), st.integers(4000, 4200), st.integers(8000, 9000))
OPS = st.lists(
    st.one_of(
        st.tuples(st.just("put"), st.sampled_from(TABLES), st.integers(0, 3000), SIZES, st.integers(0, 255)),
        st.tuples(st.just("commit")),
        st.tuples(st.just("reopen")),
    ),
    min_size=1, max_size=60)


def value(size, fill, key):
    v = np.full(size, fill, np.uint8)
    if size >= 4:
        v[:4] = np.frombuffer(int(key).to_bytes(4, "little"), np.uint8)
    return v.tobytes()


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(ops=OPS)
def test_any_sequence_of_puts_commits_and_reopens_reads_back(built_library, tmp_path_factory, ops):
    ds = built_library
    path = tmp_path_factory.mktemp("lmdbprop") / "p.lmdb"
    committed, pending = {}, {}
    w = ds.Dataset(path)
    try:
        for op in ops:
            if op[0] == "put":
                _, table, key, size, fill = op
                data = value(size, fill, key)
                w.put(table, key, data)
                pending[(table, key)] = data
            elif op[0] == "commit":
                w.commit()
                committed.update(pending)
                pending.clear()
            else:  # close (commits, like Dataset's destructor) and open again
                w.close()
                committed.update(pending)
                pending.clear()
                w = ds.Dataset(path)
        w.close()
        committed.update(pending)
    finally:
        if w.h:
            w.close()

    report = lmdb_compat.check(str(path))
    assert report["pages_leaked"] == 0
    used = {t for t, _ in committed}
    for t in used:
        assert report["tables"][t]["entries"] == sum(1 for (tt, _) in committed if tt == t)
    env = lmdb_compat.Environment(str(path), subdir=False, readonly=True, max_dbs=8)
    for t in used:
        db = env.open_db(t.encode(), integerkey=True)
        with env.begin(db=db) as txn:
            items = [(int.from_bytes(k, "little"), bytes(v)) for k, v in txn.cursor(db)]
            assert [k for k, _ in items] == sorted(k for (tt, k) in committed if tt == t)  # cursor order = key order
            for k, v in items:
                assert v == committed[(t, k)]
            assert txn.get((3001).to_bytes(4, "little")) is None
    # the first commit that writes anything also creates the tables the training side's reader opens (LmdbDataset.py:36-40)
    for t in (set(TABLES) | {"BakedInterpolationSet"}) - used:
        if committed:
            assert report["tables"][t]["entries"] == 0
            env.open_db(t.encode(), integerkey=True)
        else:
            with pytest.raises(lmdb_compat.Error):
                env.open_db(t.encode(), integerkey=True)
