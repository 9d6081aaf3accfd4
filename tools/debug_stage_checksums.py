#!/usr/bin/env python3
"""Per-stage checksums of the neural renderer in a fresh process (run several times; every line must repeat exactly)."""
import hashlib, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import deepestscatter_b200 as ds
from deepestscatter_b200 import disney_model as dm
h = lambda a: hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()[:10]
w = dm.synthetic_weights(566)
cam = ds.camera_look_at(aspect=4.0)
with ds.Context(0) as ctx:
    ctx.volume_synth(48, 0, 1234); ctx.scene_set(7000.0, (-0.03, -0.25, 0.8)); ctx.bake(); ctx.disney_model_load(w)
    out = []
    inp, info = ctx.network_input(cam, 160, 40, (0, 0, 128, 40), stream=4096)
    out.append("info " + h(info)); out.append("input " + h(inp))
    has = info["hasScattered"] != 0
    y = ctx.disney_model_forward(inp[has]); out.append("forward " + h(y))
    y2 = ctx.disney_model_forward(inp[has]); out.append("forward_again " + h(y2))
    f = ctx.render_disney(cam, 160, 40, stream=4096); out.append("frame " + h(f))
    f2 = ctx.render_disney(cam, 160, 40, stream=4096); out.append("frame_again " + h(f2))
    ctx.set_option("descriptor_hw", 0)
    f3 = ctx.render_disney(cam, 160, 40, stream=4096); out.append("frame_sw_desc " + h(f3))
    ctx.set_option("precision", ds.PRECISION_EXACT); ctx.bake()
    f4 = ctx.render_disney(cam, 160, 40, stream=4096); out.append("frame_exact " + h(f4))
    print(" | ".join(out))
