#!/usr/bin/env python3
"""Diagnostics for the two model kernels against the oracle (prints, asserts nothing)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402

import deepestscatter_b200 as ds  # noqa: E402
import oracle_lib as ol  # noqa: E402
from deepestscatter_b200 import disney_model as dm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
w = dm.synthetic_weights(566)
x = dm.synthetic_inputs(n, 7)
ref = ol.disney_forward(w, x)
with ds.Context(0) as ctx:
    ctx.disney_model_load(w)
    for name, prec in (("exact", ds.PRECISION_EXACT), ("fast", ds.PRECISION_FAST)):
        ctx.set_option("precision", prec)
        try:
            got = ctx.disney_model_forward(x)
        except Exception as e:  # noqa: BLE001
            print(name, "FAILED:", e)
            continue
        r = np.abs(got - ref) / (np.abs(ref) + 1e-6)
        print(name, "max rel", float(r.max()), "mean rel", float(r.mean()), "finite", bool(np.isfinite(got).all()))
        print("  got", got[:6], "\n  ref", ref[:6])
        bad = np.nonzero(r > 5e-3)[0]
        print("  rows over 5e-3:", len(bad), bad[:32], "by row%8:", np.bincount(bad % 8, minlength=8), "by row//32%4:", np.bincount((bad // 32) % 4, minlength=4))
