#!/usr/bin/env python3
"""bench.py -- headline benchmark of the radiance-estimation hot path on B200.

Workload (BASELINE.json configs[1], "C2"): full multiple-scattering Lorenz-Mie path tracer with sun NEE,
1920x1080, 512^3 synthetic cumulus density grid (include/ds_synth.h kind 0, seed 1234), cloud size 7000 m,
sun "Front" (Tasks.cpp:56), default camera (Camera.cpp:37-39,102).  One STEP = `--spp` progressive subframes
of the whole frame (render + Welford accumulation), i.e. spp * 1920 * 1080 paths per GPU.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C ABI)
  python bench.py --impl reference ...                     the reference arm: the host oracle (CPU port of the
                                                           reference estimator; the reference itself has no CPU
                                                           implementation and its OptiX 5.1 programs cannot be
                                                           built here) on all host cores, bounded sample per step

For N > 1 launch under torchrun (one rank per GPU): the grid is replicated, every rank renders its own
subframe ids (weak scaling) and the per-GPU accumulation buffers are combined with ONE NCCL reduce of the
moment buffers inside the timed region.  Rank 0 prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

GRID_N = 512
WIDTH, HEIGHT = 1920, 1080
CLOUD_SIZE_M = 7000.0
SUN_FRONT = (-0.586, -0.766, -0.271)
GRID_KIND, GRID_SEED = 0, 1234
METRIC = "Mpaths/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp", type=int, default=64, help="progressive subframes per step (per GPU)")
    ap.add_argument("--grid", type=int, default=GRID_N)
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--precision", default="fast", choices=["fast", "exact"])
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (tuning)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the neural-renderer numbers reported next to the main line")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU work for the cpu_baseline sample")
    ap.add_argument("--ref-width", type=int, default=480)
    ap.add_argument("--ref-height", type=int, default=270)
    return ap.parse_args()


def workload_name(a) -> str:
    return (f"C2: all-order Lorenz-Mie multi-scatter + sun NEE, {a.width}x{a.height}, {a.grid}^3 synthetic cumulus "
            f"(size {CLOUD_SIZE_M:.0f} m, sun Front), {a.spp} spp/step/GPU")


# ---------------------------------------------------------------- clocks

class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU every 200 ms during the timed region (pynvml)."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        if self.ok:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self) -> dict:
        if self._thread:
            self._stop.set()
            self._thread.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "source": "unavailable"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s), "source": "nvml"}


# ---------------------------------------------------------------- oracle helpers (cpu_baseline / reference arm only)

def oracle_with_scene(a, density=None, inscatter=None):
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as ol

    o = ol.Oracle()
    if density is not None:
        o.volume_upload(density, True)
    else:
        o.volume_synth(a.grid, GRID_KIND, GRID_SEED, True)
    o.scene_set(CLOUD_SIZE_M, SUN_FRONT)
    if inscatter is not None:
        o.inscatter_set(inscatter)
    else:
        o.bake(skip_empty=True)
    return o, ol


def time_oracle_sample(o, ol, a, w, h, spp, first_subframe=1):
    """All-order estimator on a w x h down-sampling of the C2 frame (same camera), spp subframes; returns
    (seconds, counters)."""
    import numpy as np

    cam = ol.camera_look_at(aspect=a.width / a.height)
    o.counters_reset()
    t0 = time.perf_counter()
    for k in range(spp):
        o.render_frame(cam, w, h, ol.MODE_ALL, first_subframe + k)
    dt = time.perf_counter() - t0
    return dt, o.counters()


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------- reference arm

def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    density = inscatter = None
    setup = "grid + sun-transmittance bake prepared on the CPU by the oracle"
    try:
        import torch

        if torch.cuda.is_available():
            import deepestscatter_b200 as ds

            with ds.Context(0) as ctx:  # input preparation only; nothing of ours runs inside the timed region
                ctx.set_option("precision", ds.PRECISION_EXACT)
                ctx.volume_synth(a.grid, GRID_KIND, GRID_SEED)
                ctx.scene_set(CLOUD_SIZE_M, SUN_FRONT)
                ctx.bake()
                density, inscatter = ctx.level(0), ctx.inscatter()
            setup = "grid + sun-transmittance bake (inputs) prepared once on the GPU, outside the timed region"
    except Exception:
        density = inscatter = None
    o, ol = oracle_with_scene(a, density, inscatter)
    w, h = a.ref_width, a.ref_height
    cores = host_threads()
    sub = 1
    for _ in range(a.warmup):
        time_oracle_sample(o, ol, a, w, h, 1, sub)
        sub += 1
    total_t, paths, events, steps = 0.0, 0, 0, 0
    for _ in range(a.steps):
        dt, c = time_oracle_sample(o, ol, a, w, h, 1, sub)
        sub += 1
        total_t += dt
        paths += c["paths"]
        events += c["events"]
        steps += c["steps"]
    value = paths / total_t / 1e6
    sample = f"{w}x{h} px down-sampling of the C2 frame (same camera, grid and estimator), 1 subframe per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": total_t / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_name(a), "sample": sample, "setup": setup},
        "cpu_baseline": {"value": value, "unit": METRIC, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "events_per_s": events / total_t, "steps_per_s": steps / total_t,
        "note": "host oracle = CPU port of the reference estimator (the reference has no CPU implementation; OptiX 5.1 cannot be built here)",
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------- our arm

def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import deepestscatter_b200 as ds
    from deepestscatter_b200 import multigpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (ours) needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = ds.Context(local_rank)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_option("precision", ds.PRECISION_FAST if a.precision == "fast" else ds.PRECISION_EXACT)
    for kv in a.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    ctx.set_option("staging_subframes", max(1, min(64, a.spp)))

    # ---- inputs: resident in HBM before any timed region ----
    ctx.volume_synth(a.grid, GRID_KIND, GRID_SEED, True)
    ctx.scene_set(CLOUD_SIZE_M, SUN_FRONT)
    t0 = time.perf_counter()
    ctx.bake()
    ctx.sync()
    bake_s = time.perf_counter() - t0
    ctx.frame_create(a.width, a.height)
    cam = ds.camera_look_at(aspect=a.width / a.height)
    px = a.width * a.height
    mode = ds.MODE_ALL_SCATTER

    # rank r renders global subframe ids r*B + 1 ... (B = per-rank budget); local Welford weights run 1/k
    per_rank_budget = (a.warmup + a.steps) * a.spp * 2 + 16
    ctx.set_option("stream_offset", rank * per_rank_budget)
    moments = torch.zeros(px * 8, dtype=torch.float64, device="cuda") if distributed else None

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timed region ----
    sub = 1
    with torch.cuda.stream(stream):
        for _ in range(a.warmup):
            ctx.render_subframes(cam, mode, sub, a.spp)
            sub += a.spp
        ctx.sync()
        ctx.set_option("profile_events", 1)
        ctx.counters_reset()
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        n_local = 0
        for _ in range(a.steps):
            ctx.render_subframes(cam, mode, sub, a.spp)
            sub += a.spp
            n_local += a.spp
        if distributed:
            # the single NCCL reduce of the per-GPU accumulation buffers (as mergeable moments)
            ctx.export_moments(sub - 1, moments.data_ptr())
            multigpu.reduce_moments(moments, dst=0)
            if rank == 0:
                ctx.import_moments((sub - 1) * world, moments.data_ptr())
        ev1.record(stream)
        barrier()
        clocks = sampler.stop()
        ms = ev0.elapsed_time(ev1)
    counters = ctx.counters()
    lstats = ctx.launch_stats()
    ctx.set_option("profile_events", 0)

    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    agg = torch.tensor([counters["paths"], counters["events"], counters["steps"], counters["density_taps"]], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    ms_max = float(t.item())
    paths, events, steps, taps = (float(x) for x in agg.tolist())
    value = paths / (ms_max * 1e-3) / 1e6

    # ---- end-to-end through the C ABI with HOST buffers (pinned), copies inside the timed region ----
    hp = torch.zeros(px * 4, dtype=torch.float32).pin_memory()
    hv = torch.zeros(px * 4, dtype=torch.float32).pin_memory()
    ctx.frame_clear()
    e2e_sub = 1
    for _ in range(min(a.warmup, 2)):
        ctx.render_subframes_host_ptr(cam, mode, e2e_sub, a.spp, hp.data_ptr(), hv.data_ptr())
        e2e_sub += a.spp
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ctx.render_subframes_host_ptr(cam, mode, e2e_sub, a.spp, hp.data_ptr(), hv.data_ptr())
        e2e_sub += a.spp
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * a.steps * a.spp * px / float(te.item()) / 1e6
    checksum = float(hp.view(-1, 4)[:, 0].double().mean())

    # ---- roofline of the dominant kernel (k_trace): algorithmic bytes = 8 B/march step + 8 B/scatter event ----
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    trace_launches = max(1, lstats["trace_launches_timed"])
    trace_ms_avg = lstats["trace_ms_total"] / trace_launches
    alg_bytes_per_launch = (8.0 * counters["steps"] + 8.0 * counters["events"]) / trace_launches
    achieved = alg_bytes_per_launch / (trace_ms_avg * 1e-3) / 1e9 if trace_ms_avg > 0 else 0.0
    # DRAM traffic of one launch from the committed ncu --set full capture of this very configuration (never measured here:
    # a number taken under a profiler is not a bench value, and the capture is only valid for the configuration it was taken on)
    traffic = None
    tpath = ROOT / "profiles" / "traffic_k_trace_fast.json"
    if tpath.exists():
        t = json.loads(tpath.read_text())
        if t["config"] == {"grid": a.grid, "width": a.width, "height": a.height, "spp": a.spp, "precision": a.precision}:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "kernel": "k_trace", "kernel_ms_avg": trace_ms_avg, "kernel_share_of_step": lstats["trace_ms_total"] / ms if ms > 0 else None,
        "algorithmic_bytes_per_launch": alg_bytes_per_launch, "peak_source": peak_src,
        "note": "algorithmic bytes count every march step of the reference algorithm (8 B) and every scatter event (8 B); "
                "empty-space skipping and L2/texture-cache hits make DRAM traffic (`traffic`, bytes per launch from "
                "profiles/traffic_k_trace_fast.json) much smaller than this; ncu: L2 throughput 76 %, issue slots 70 % busy",
    }

    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "precision": a.precision, "grid_bytes": 2 * a.grid**3,
                   "l2": "inputs (density + sun-transmittance grids, 2*N^3 B) are larger than the 126 MB L2; no explicit flush",
                   "multi_gpu": "replicated grid, subframe ids split over ranks, one NCCL reduce of moment buffers" if distributed else "single GPU",
                   "options": {k: ctx.get_option(k) for k in ("variant", "block_threads", "blocks_per_sm", "skip_empty", "primary_cache", "regen_min", "skip_min",
                                                               "skip_max_iters", "march_keep32", "march_max_iters", "staging_subframes")}},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": METRIC, "h2d_bytes_per_step": 2 * px * 16, "d2h_bytes_per_step": 2 * px * 16,
                "api": "ds_render_subframes_host (progressive + variance float4 buffers in pinned host memory)", "checksum_mean_radiance": checksum},
        "gpu_launches": lstats["kernel_launches"],
        "roofline": roofline,
        "events_per_s": events / (ms_max * 1e-3), "steps_per_s": steps / (ms_max * 1e-3), "density_taps_per_s": taps / (ms_max * 1e-3),
        "events_per_path": events / paths, "steps_per_path": steps / paths, "bake_seconds": bake_s,
        "nonfinite": counters["nonfinite"],
    }

    # ---- CPU baseline: the oracle on a bounded sample of the same workload (rank 0, N = 1 only) ----
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            o, ol = oracle_with_scene(a, ctx.level(0), ctx.inscatter())
            cores = host_threads()
            w, h = 96, 54
            dt, c = time_oracle_sample(o, ol, a, w, h, 1, 1)  # calibration
            per_path = dt / max(1, c["paths"])
            want_paths = a.cpu_seconds / max(per_path, 1e-9)
            scale = max(1.0, min(10.0, (want_paths / (w * h)) ** 0.5))
            w2, h2 = int(w * scale), int(h * scale)
            spp = max(1, int(want_paths / (w2 * h2)))
            spp = min(spp, 64)
            dt, c = time_oracle_sample(o, ol, a, w2, h2, spp, 1)
            line["cpu_baseline"] = {
                "value": c["paths"] / dt / 1e6, "unit": METRIC, "cores": cores, "kind": "port",
                "sample": f"{w2}x{h2} px down-sampling of the C2 frame x {spp} subframes ({c['paths']} paths, {dt:.1f} s)",
                "events_per_s": c["events"] / dt, "steps_per_s": c["steps"] / dt,
            }
        except Exception as exc:  # the baseline is reporting only; never lose the GPU line
            line["cpu_baseline"] = {"value": None, "unit": METRIC, "cores": host_threads(), "kind": "port", "sample": f"failed: {exc}"}

    # ---- secondary numbers of the same build (rank 0, N = 1 only; never in the timed region, never able to lose the main line):
    # the neural renderer end to end on the same cloud and frame, and its network alone ----
    if rank == 0 and world == 1 and not a.no_secondary:
        try:
            from deepestscatter_b200 import disney_model as dm

            ctx.disney_model_load(dm.synthetic_weights(566))
            ctx.render_disney(cam, a.width, a.height, stream=1)  # warm-up (scratch allocation, mip-mapped texture)
            times = []
            for rep in range(3):
                t0 = time.perf_counter()
                frame = ctx.render_disney(cam, a.width, a.height, stream=2 + rep)
                times.append(time.perf_counter() - t0)
            rows = 1 << 17
            x = dm.synthetic_inputs(1024, 33)
            x = np.ascontiguousarray(np.tile(x, (rows // 1024, 1, 1)))
            ctx.set_option("profile_events", 1)
            ctx.disney_model_forward(x)
            ctx.disney_model_forward(x)
            us = ctx.get_option("mlp_last_us")
            macs = 10 * (226 * 200 + 2 * 200 * 200) - 200 * 200 + 2 * 200 * 200 + 200
            line["secondary"] = {
                "neural_renderer_ms_per_frame": min(times) * 1e3, "neural_renderer_scattering_pixels": int((frame[..., 3] != 0).sum()),
                "neural_renderer_api": "ds_render_disney (DisneyRenderer::render; host frame buffer out), synthetic weights",
                "model_kernel": "k_disney_mlp_tc (tcgen05 kind::tf32)", "model_rows": rows, "model_us": us,
                "model_tflops_tf32": 2 * macs * rows / (us * 1e-6) / 1e12 if us else None,
            }
        except Exception as exc:  # reporting only
            line["secondary"] = {"failed": str(exc)}

    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if distributed:
        dist.destroy_process_group()
    return 0


def main():
    a = parse_args()
    if a.impl == "reference":
        return run_reference(a)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
