#!/usr/bin/env python3
"""Neural renderer end to end (ds_render_disney = DisneyRenderer::render): one 1920 x 1080 frame on the C2 cloud, both flavours;
wall time per frame through the C ABI (host frame buffer out), one JSON line per flavour."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

import deepestscatter_b200 as ds  # noqa: E402
from deepestscatter_b200 import disney_model as dm  # noqa: E402


def main():
    w, h, grid = 1920, 1080, 512
    if len(sys.argv) > 1:
        w, h, grid = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    flavours = sys.argv[4].split(",") if len(sys.argv) > 4 else ["fast", "fast_tf32", "fast_bf16", "exact"]
    cam = ds.camera_look_at(aspect=w / h)
    with ds.Context(0) as ctx:
        ctx.volume_synth(grid, 0, 1234, True)
        ctx.scene_set(7000.0, (-0.586, -0.766, -0.271))
        ctx.disney_model_load(dm.synthetic_weights(566))
        for name in flavours:
            ctx.set_option("precision", ds.PRECISION_EXACT if name == "exact" else ds.PRECISION_FAST)
            ctx.set_option("mlp_bf16", 1 if name == "fast_bf16" else 0)
            ctx.set_option("mlp_fp16", 0 if name == "fast_tf32" else 1)  # IEEE half operands are the default
            ctx.bake()
            ctx.sync()
            times = []
            for rep in range(3):
                t0 = time.perf_counter()
                frame = ctx.render_disney(cam, w, h, stream=1 + rep * 1000)
                times.append(time.perf_counter() - t0)
            lit = frame[..., 3] != 0
            print(json.dumps({"flavour": name, "frame": [w, h], "grid": grid, "seconds": times, "best_s": min(times), "fps": 1.0 / min(times),
                              "scattering_pixels": int(lit.sum()), "mean_rgb": [float(frame[..., c].mean()) for c in range(3)],
                              "finite": bool(np.isfinite(frame).all())}), flush=True)


if __name__ == "__main__":
    main()
