#!/usr/bin/env python3
"""Reproduce the table statistics the reference's notebook recorded for the authors' dataset (DeepestScatter_Train/DatasetVisualisation.ipynb,
output of cell 1, written by the real liblmdb):

    ScatterSample {'psize': 4096, 'depth': 4, 'branch_pages': 359, 'leaf_pages': 103132, 'overflow_pages': 0, 'entries': 8663040}
    Result        {'psize': 4096, 'depth': 2, 'branch_pages': 1,   'leaf_pages': 78,     'overflow_pages': 0, 'entries': 14336}

by appending the same numbers of records of the same sizes with this library's writer (needs ~3 GB of RAM and a 420 MB scratch file).
tests/test_lmdb.py checks the Result line on every run; this tool is the full-size check of the ScatterSample line."""
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

sys.path.insert(0, str(ROOT / "tests"))
import lmdb_compat  # noqa: E402  (test-side reader)
import deepestscatter_b200 as ds  # noqa: E402

WANT = {"ScatterSample": {"psize": 4096, "depth": 4, "branch_pages": 359, "leaf_pages": 103132, "overflow_pages": 0, "entries": 8663040},
        "Result": {"psize": 4096, "depth": 2, "branch_pages": 1, "leaf_pages": 78, "overflow_pages": 0, "entries": 14336}}


def main():
    path = Path(tempfile.mkdtemp()) / "Train.lmdb"
    rng = np.random.default_rng(1)
    t0 = time.time()
    with ds.Dataset(path) as w:
        n = WANT["ScatterSample"]["entries"]
        step = 2048 * 256
        for start in range(0, n, step):
            m = min(step, n - start)
            pos = rng.uniform(0.01, 0.5, (m, 3)).astype(np.float32)  # six non-zero floats: 34-byte records
            d = rng.uniform(0.01, 1.0, (m, 3)).astype(np.float32)
            w.append_scatter_samples(start, pos, d)
        m = WANT["Result"]["entries"]
        w.append_results(0, rng.uniform(0.5, 3.0, m).astype(np.float32), np.ones(m, np.uint8))
    env = lmdb_compat.Environment(str(path), subdir=False, readonly=True, max_dbs=8)
    ok = True
    with env.begin() as txn:
        for name, want in WANT.items():
            got = txn.stat(env.open_db(name.encode(), integerkey=True))
            print(name, got, "== notebook" if got == want else f"!= notebook {want}")
            ok &= got == want
    print(f"{time.time() - t0:.1f} s, file {path.stat().st_size / 1e6:.0f} MB")
    path.unlink()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
