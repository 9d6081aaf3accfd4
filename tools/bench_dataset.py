#!/usr/bin/env python3
"""C3 (BASELINE.json configs[2]): Deep Scattering dataset generation -- per scene 2048 (point, sun_dir, view_dir)
samples with hierarchical stencil descriptors and converged ground-truth radiance, written as LMDB records.

Per scene, the reference's three collectors in order (DG/ExecutionLoop/Tasks.cpp:116-178): ScatterSampleCollector ->
DisneyDescriptorCollector -> RadianceCollector, each followed by its records.  Scenes (cloud size log-uniform in
1..12 km, sun uniform on the sphere: DeepestScatter_Train/Utils/GenerateSceneSetups.py:11-21,48-52; numpy seed 566) are
sharded over ranks by scene id (no collective: each rank writes its own shard, shards are merged by key afterwards).

  python tools/bench_dataset.py --scenes 8                      one GPU
  torchrun --nproc-per-node N tools/bench_dataset.py --scenes 64   N GPUs (weak scaling when --scenes-per-gpu is given)
Prints one JSON line (rank 0)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

sys.path.insert(0, str(ROOT / "tests"))
import lmdb_compat  # noqa: E402  (test-side reader)
import deepestscatter_b200 as ds  # noqa: E402

BATCH = 2048


def scene_setups(count: int, seed: int = 566):
    """GenerateSceneSetups.py:44-52 with a fixed seed."""
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(count):
        size = float(np.exp(rng.uniform(np.log(1000), np.log(12000))))
        cos_t = rng.uniform(-1, 1)
        phi = rng.uniform(0, np.pi * 2)
        sin_t = np.sqrt(1 - cos_t * cos_t)
        out.append((size, (float(np.cos(phi) * sin_t), float(np.sin(phi) * sin_t), float(cos_t))))
    return out


def cpu_baseline(ctx, a, setup):
    """the host restatement (oracle port, OpenMP, all cores) on scene 0 x 64 samples: the descriptor gather
    in full, and a BOUNDED number of RadianceCollector updates (100 launches x 20480 threads each; to the CI rule a sample needs
    ~5 M experiments, i.e. hours on a CPU).  Grid and sun-transmittance volume are handed over from the GPU context (both bit-exact
    with the oracle's own, tests/test_gpu_parity.py)."""
    import oracle_lib as ol

    size, sun = setup
    cores = len(os.sched_getaffinity(0))
    ol.lib().orc_set_threads(cores)
    ctx.set_option("precision", ds.PRECISION_EXACT)
    ctx.scene_set(size, sun)
    ctx.bake()
    o = ol.Oracle()
    o.volume_upload(ctx.level(0))
    o.scene_set(size, sun)
    o.inscatter_set(ctx.inscatter())
    ctx.set_option("precision", ds.PRECISION_FAST)
    n = 64
    t0 = time.perf_counter()
    pos, dirs = o.generate_points(0, n, 0)
    t1 = time.perf_counter()
    reps = 8
    for _ in range(reps):
        o.descriptors(pos, dirs)
    t2 = time.perf_counter()
    o.counters_reset()
    tasks, conv, nconv, updates = o.point_radiance(pos, dirs, a.max_threads, a.launches, a.cpu_baseline)
    t3 = time.perf_counter()
    c = o.counters()
    return {"kind": "port", "cores": cores, "sample": f"scene 0 ({size:.0f} m), {n} samples: points + descriptors in full, {updates} RadianceCollector updates "
                                                      f"of {a.launches} x {a.max_threads} paths (a bounded part of the convergence loop)",
            "points_per_s": n / (t1 - t0), "descriptors_per_s": n * reps / (t2 - t1), "radiance_mpaths_per_s": c["paths"] / (t3 - t2) / 1e6,
            "radiance_events_per_s": c["events"] / (t3 - t2), "radiance_seconds": t3 - t2, "converged_in_sample": int(nconv)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=4, help="total scenes (all ranks)")
    ap.add_argument("--scenes-per-gpu", type=int, default=0, help="weak scaling: total = this x world size")
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--max-threads", type=int, default=20480, help="RadianceCollector MAX_THREAD_COUNT (RadianceCollector.cpp:17)")
    ap.add_argument("--launches", type=int, default=100, help="launches per update (RadianceCollector.cpp:88)")
    ap.add_argument("--out", default="/tmp/ds_c3")
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--static", action="store_true", help="round-robin scene split instead of the shared scene counter")
    ap.add_argument("--cpu-baseline", type=int, default=0, metavar="UPDATES",
                    help="rank 0, one GPU: time the CPU restatement on scene 0 x 64 samples (descriptor gather in full, UPDATES RadianceCollector updates)")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    total = a.scenes_per_gpu * world if a.scenes_per_gpu else a.scenes
    setups = scene_setups(total)
    from deepestscatter_b200.multigpu import SceneQueue, static_scenes

    if dist is not None and not a.static:
        mine = SceneQueue(dist.distributed_c10d._get_default_store(), range(total))  # dynamic: next scene from a shared counter
    else:
        mine = static_scenes(rank, world, range(total))

    out_dir = Path(a.out)
    out_dir.mkdir(parents=True, exist_ok=True)
    shard = out_dir / f"Train.rank{rank}.lmdb"
    if shard.exists():
        shard.unlink()
    ctx = ds.Context(local)
    for o in a.opt:
        k, v = o.split("=")
        ctx.set_option(k, int(v))
    ctx.volume_synth(a.grid, 0, 1234, True)  # one cloud, cached across scenes like Resources::volumeCache
    t = dict(bake=0.0, points=0.0, descriptors=0.0, radiance=0.0, records=0.0)
    stats = dict(samples=0, converged=0, updates=0, experiments=0)
    if dist is not None:
        dist.barrier()
    ctx.sync()
    ctx.counters_reset()
    t_all = time.perf_counter()
    with ds.Dataset(shard) as store:
        for sid in mine:
            size, sun = setups[sid]
            t0 = time.perf_counter()
            ctx.scene_set(size, sun)
            ctx.bake()
            ctx.sync()
            t1 = time.perf_counter()
            pos, dirs = ctx.generate_points(0, a.batch, stream=sid)
            t2 = time.perf_counter()
            desc = ctx.descriptors(pos, dirs)
            t3 = time.perf_counter()
            tasks, conv, nconv, updates = ctx.point_radiance(pos, dirs, max_threads=a.max_threads, launches_per_update=a.launches)
            t4 = time.perf_counter()
            store.append_scene_setup(sid, f"synthetic/cumulus_{a.grid}.vdb", size, sun)
            store.append_scatter_samples(sid * a.batch, pos, dirs)
            store.append_descriptors(sid * a.batch, desc)
            store.append_results(sid * a.batch, tasks["radiance"], conv)
            store.commit()
            t5 = time.perf_counter()
            for k, v in zip(t, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
                t[k] += v
            stats["samples"] += a.batch
            stats["converged"] += int(nconv)
            stats["updates"] += int(updates)
            stats["experiments"] += int(tasks["experimentCount"].astype(np.int64).sum())
    ctx.sync()
    elapsed = time.perf_counter() - t_all
    c = ctx.counters()
    if dist is not None:
        import torch

        v = torch.tensor([elapsed, stats["samples"], stats["converged"], stats["updates"], stats["experiments"], c["paths"], c["events"], c["steps"]]
                         + list(t.values()), dtype=torch.float64, device="cuda")
        mx = v.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
        elapsed = float(mx[0])
        stats = dict(samples=int(v[1]), converged=int(v[2]), updates=int(v[3]), experiments=int(v[4]))
        c = dict(paths=int(v[5]), events=int(v[6]), steps=int(v[7]))
        t = {k: float(x) / world for k, x in zip(t, v[8:])}
    if rank == 0:
        rep = lmdb_compat.check(str(shard))
        line = {
            "metric": "samples/s", "value": stats["samples"] / elapsed, "unit": "samples/s", "n_gpus": world,
            "config": {"workload": f"C3: dataset generation, {total} scenes x {a.batch} samples, {a.grid}^3 synthetic cumulus, sizes 1-12 km log-uniform, "
                                   f"sun uniform on the sphere (seed 566); radiance to the reference CI rule (2 % relative / 1e-4 absolute)",
                       "max_thread_count": a.max_threads, "launches_per_update": a.launches,
                       "scene_assignment": "static round-robin" if (a.static or world == 1) else "dynamic (shared counter in the process group's store)"},
            "seconds": elapsed, "scenes": total, **stats, "mpaths_per_s": c["paths"] / elapsed / 1e6, "events_per_s": c["events"] / elapsed,
            "steps_per_s": c["steps"] / elapsed, "seconds_per_stage_per_rank": t,
            "shard0": {"pages": rep["pages_total"], "leaked": rep["pages_leaked"], "tables": {k: v["entries"] for k, v in rep["tables"].items()}},
        }
        if a.cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(ctx, a, setups[0])
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if not a.keep and shard.exists():
        shard.unlink()


if __name__ == "__main__":
    main()
