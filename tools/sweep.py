#!/usr/bin/env python3
"""Tuning sweep on one GPU: times ds_render_subframes for several option sets on the C2 workload.
Prints one JSON line per configuration (gpurun_out/sweep.jsonl when --out is given)."""
import argparse
import itertools
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import deepestscatter_b200 as ds  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=2)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--size", type=float, default=7000.0)
    ap.add_argument("--sun", default="-0.586,-0.766,-0.271")
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--kind", type=int, default=0)
    ap.add_argument("--tag", default="")
    ap.add_argument("--out", default=None)
    ap.add_argument("--pre", action="append", default=[], help="name=value set before the volume is created")
    ap.add_argument("--set", action="append", default=[], help="name=v1,v2,... (cartesian product)")
    a = ap.parse_args()
    sun = tuple(float(x) for x in a.sun.split(","))
    ctx = ds.Context(0)
    for s in a.pre:
        k, v = s.split("=")
        ctx.set_option(k, int(v))
    ctx.volume_synth(a.grid, a.kind, 1234, True)
    ctx.scene_set(a.size, sun)
    baked = {}
    ctx.frame_create(a.width, a.height)
    cam = ds.camera_look_at(aspect=a.width / a.height)
    ctx.set_option("profile_events", 1)
    names, values = [], []
    for s in a.set:
        k, v = s.split("=")
        names.append(k)
        values.append([int(x) for x in v.split(",")])
    out = open(a.out, "a") if a.out else None
    for combo in itertools.product(*values) if values else [()]:
        opts = dict(zip(names, combo))
        for k, v in opts.items():
            ctx.set_option(k, v)
        prec = ctx.get_option("precision")
        t0 = time.perf_counter()
        ctx.bake()
        ctx.sync()
        bake_s = time.perf_counter() - t0
        ctx.frame_clear()
        ctx.render_subframes(cam, a.mode, 1, a.spp)  # warm-up
        ctx.sync()
        ctx.counters_reset()
        t0 = time.perf_counter()
        for r in range(a.reps):
            ctx.render_subframes(cam, a.mode, 1 + (r + 1) * a.spp, a.spp)
        ctx.sync()
        dt = time.perf_counter() - t0
        c = ctx.counters()
        ls = ctx.launch_stats()
        p, _ = ctx.frame_download()
        rec = dict(tag=a.tag, pre=a.pre, opts=opts, nonfinite=c['nonfinite'], precision=prec, mpaths_s=c["paths"] / dt / 1e6, gevents_s=c["events"] / dt / 1e9, gsteps_s=c["steps"] / dt / 1e9,
                   gtaps_s=c["density_taps"] / dt / 1e9, events_per_path=c["events"] / c["paths"], steps_per_path=c["steps"] / c["paths"],
                   trace_ms=ls["trace_ms_total"] / max(1, ls["trace_launches_timed"]), wall_s=dt, bake_s=bake_s,
                   alg_gbs=(8 * c["steps"] + 8 * c["events"]) / (ls["trace_ms_total"] * 1e-3) / 1e9, mean=float(p[..., 0].mean()))
        line = json.dumps(rec)
        print(line, flush=True)
        if out:
            out.write(line + "\n")
            out.flush()
    ctx.close()


if __name__ == "__main__":
    main()
