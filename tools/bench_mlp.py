#!/usr/bin/env python3
"""Device time of the radiance-predicting network (ds_disney_model_forward) for one 128 x 128 rectangle of DisneyRenderer, both
flavours, against the dense-math roofline: 2 * 16384 * 1 331 400 MACs... printed as one JSON line per flavour."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

import deepestscatter_b200 as ds  # noqa: E402
from deepestscatter_b200 import disney_model as dm  # noqa: E402

MACS_PER_ROW = 10 * (226 * 200 + 2 * 200 * 200) - 200 * 200 + 2 * 200 * 200 + 200  # block 0 has no f1o product (o = 0)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128 * 128
    reps = 5
    w = dm.synthetic_weights(566)
    x = dm.synthetic_inputs(n, 33)
    with ds.Context(0) as ctx:
        ctx.disney_model_load(w)
        ctx.set_option("profile_events", 1)
        outs = {}
        for name, prec in (("exact_f32_fma", ds.PRECISION_EXACT), ("fast_tcgen05_tf32", ds.PRECISION_FAST), ("fast_tcgen05_bf16", ds.PRECISION_FAST),
                           ("fast_tcgen05_fp16", ds.PRECISION_FAST)):
            ctx.set_option("precision", prec)
            ctx.set_option("mlp_bf16", 1 if name.endswith("bf16") else 0)
            ctx.set_option("mlp_fp16", 1 if name.endswith("fp16") else 0)  # on by default
            us = []
            for _ in range(reps):
                outs[name] = ctx.disney_model_forward(x)
                us.append(ctx.get_option("mlp_last_us"))
            best = min(us[1:])
            print(json.dumps({"kernel": name, "rows": n, "us": us, "best_us": best, "tflops": 2 * MACS_PER_ROW * n / (best * 1e-6) / 1e12,
                              "rows_per_s": n / (best * 1e-6)}), flush=True)
            if prec == ds.PRECISION_FAST:
                ctx.set_option("profile_events", 2)  # instrumented kernel: where block 0 spends its cycles
                ctx.disney_model_forward(x)
                print(json.dumps({"block0_cycles": ctx.disney_model_profile(), "instrumented_us": ctx.get_option("mlp_last_us")}), flush=True)
                ctx.set_option("profile_events", 1)
        a, b, c, d = outs["exact_f32_fma"], outs["fast_tcgen05_tf32"], outs["fast_tcgen05_bf16"], outs["fast_tcgen05_fp16"]
        print(json.dumps({"max_rel_diff_tf32_vs_exact": float(np.max(np.abs(a - b) / (np.abs(a) + 1e-6))),
                          "max_rel_diff_bf16_vs_exact": float(np.max(np.abs(a - c) / (np.abs(a) + 1e-6))),
                          "max_rel_diff_fp16_vs_exact": float(np.max(np.abs(a - d) / (np.abs(a) + 1e-6)))}))


if __name__ == "__main__":
    main()
