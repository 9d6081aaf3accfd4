#!/usr/bin/env python3
"""bench.py -- headline benchmark of the radiance-estimation hot path on B200.

Workloads are BASELINE.json's configurations, all on the synthetic grids of include/ds_synth.h:
  C1  single-scatter Beer-Lambert + Mie, cloud cube 256^3, 256x256, 64 spp             (configs[0])
  C2  all-order Lorenz-Mie multi-scatter + sun NEE, cumulus 512^3, 1920x1080, 1024 spp (configs[1], the default: the
      configuration the metric is quoted on)
  C4  thick cumulus 1024^3, 12 km, grazing sun, 1920x1080                              (configs[3])
  C5  as C2 at 3840x2160, 8192 spp, split over 2/4/8 GPUs                              (configs[4])
One STEP = one progressive frame of `--spp` subframes: frame cleared, spp subframes rendered and Welford-accumulated, and -- on
more than one GPU -- the per-GPU accumulation buffers combined by ONE NCCL reduce (ds_frame_reduce, inside the library) so that
rank 0 holds the finished frame.  --scaling weak: every GPU renders --spp subframes (N*spp per frame); --scaling strong: the
--spp subframes of the frame are split over the GPUs (multigpu.subframe_range).

  python bench.py --gpus N --steps K --warmup W [--config C2] [--scaling weak]     our arm (CUDA, through the C ABI)
  python bench.py --impl reference ...     the reference arm: the reference's own estimator source compiled for the host
                                           (oracle/_ref, see oracle/ref_shim/) on all host cores, tracing a stated subset of the
                                           SAME pixels and subframes; falls back to the oracle port where _ref is not built

For N > 1 launch under torchrun (one rank per GPU).  Rank 0 prints one JSON line.
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

METRIC = "Mpaths/s"
SUN_FRONT = (-0.586, -0.766, -0.271)  # Tasks.cpp:56
SUN_SIDE = (-0.03, -0.25, 0.8)  # Tasks.cpp:58
SUN_GRAZING = (0.995, -0.0998, 0.0)
MODE_ALL, MODE_MULTI, MODE_SINGLE = 0, 1, 2

CONFIGS = {
    "C1": dict(grid=256, kind=1, size_m=7000.0, sun=SUN_SIDE, width=256, height=256, mode=MODE_SINGLE, total_spp=64,
               what="single-scatter Beer-Lambert + Mie, cloud cube"),
    "C2": dict(grid=512, kind=0, size_m=7000.0, sun=SUN_FRONT, width=1920, height=1080, mode=MODE_ALL, total_spp=1024,
               what="all-order Lorenz-Mie multi-scatter + sun NEE, cumulus"),
    "C4": dict(grid=1024, kind=0, size_m=12000.0, sun=SUN_GRAZING, width=1920, height=1080, mode=MODE_ALL, total_spp=64,
               what="all-order Mie + NEE, thick cumulus at a grazing sun (divergence stress)"),
    "C5": dict(grid=512, kind=0, size_m=7000.0, sun=SUN_FRONT, width=3840, height=2160, mode=MODE_ALL, total_spp=8192,
               what="all-order Mie + NEE, cumulus, 4K frame (multi-GPU scaling sweep)"),
}
GRID_SEED = 1234


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--spp", type=int, default=64, help="subframes per step: per GPU (weak) or per frame (strong)")
    ap.add_argument("--grid", type=int, default=None, help="override the configuration's grid size (tuning only)")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--precision", default="fast", choices=["fast", "exact"])
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (tuning)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the descriptor / bake / neural-renderer legs")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU work per reference step / cpu_baseline sample")
    ap.add_argument("--ref-stride", type=int, default=0, help="reference arm: trace every stride-th pixel in x and y (0 = auto)")
    a = ap.parse_args()
    c = dict(CONFIGS[a.config])
    for k in ("grid", "width", "height"):
        if getattr(a, k) is not None:
            c[k] = getattr(a, k)
    a.cfg = c
    return a


def config_dict(a) -> dict:
    """Identical for both arms (the driver compares it); per-arm detail goes elsewhere on the line."""
    c = a.cfg
    return {
        "workload": f"{a.config}: {c['what']} {c['grid']}^3 (ds_synth kind {c['kind']}, seed {GRID_SEED}), size {c['size_m']:.0f} m, "
                    f"{c['width']}x{c['height']}, {a.spp} spp per step" + (" per GPU" if a.scaling == "weak" else " per frame"),
        "config": a.config, "grid": c["grid"], "width": c["width"], "height": c["height"], "mode": c["mode"], "spp_per_step": a.spp,
        "scaling": a.scaling, "target_spp": c["total_spp"],
        "l2": "inputs (density + sun-transmittance grids, 2*N^3 B) exceed the 126 MB L2 from 512^3 up; no explicit flush",
    }


# ---------------------------------------------------------------- clocks

class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU every 200 ms during the timed region (pynvml)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop, self._thread, self.ok = threading.Event(), None, False
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


# ---------------------------------------------------------------- CPU side (reference arm and cpu_baseline leg only)

def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class CpuEstimator:
    """The reference's estimator on the host cores for one configuration: oracle/_ref (the reference's own source, kind
    "reference") when that library exists, else the oracle port (kind "port").  Traces explicit pixels of the configuration's
    frame -- same camera ray, same seed tea<4>(x*4096 + y, subframe) as the GPU arm -- so the sample is a subset of the same work.
    No product code is imported here."""

    def __init__(self, a):
        sys.path.insert(0, str(ROOT / "tests"))
        import numpy as np
        import oracle_lib as ol

        self.np, self.ol, self.a, self.c = np, ol, a, a.cfg
        self.cores = host_threads()
        ol.lib().orc_set_threads(self.cores)  # torch.distributed.run exports OMP_NUM_THREADS=1
        c = self.c
        self.o = ol.Oracle()
        self.o.volume_synth(c["grid"], c["kind"], GRID_SEED, True)
        self.o.scene_set(c["size_m"], c["sun"])
        self.o.bake(skip_empty=True)  # input preparation (bit-identical to the reference's bake, tests/test_oracle_vs_ref.py)
        self.cam = ol.camera_look_at(aspect=c["width"] / c["height"])
        self.ref = None
        self.kind = "port"
        try:
            import ref_lib as rl

            if rl.available():
                rl.skip_bake(True)
                r = rl.Reference()
                r.volume_upload(self.o.level(0))
                r.scene_init(c["size_m"], c["sun"], 1.0 / 512.0, c["mode"], 8, 8)
                r.inscatter_set(self.o.inscatter())
                rl.skip_bake(False)
                self.ref, self.kind = r, "reference"
        except Exception as exc:  # the port remains
            print(f"bench.py: oracle/_ref unavailable ({exc}); timing the oracle port", file=sys.stderr)

    def rays(self, stride: int, subframe: int):
        """Pixels (i*stride, j*stride) of the frame at `subframe`: origins, directions, seed values, streams."""
        np, c = self.np, self.c
        xs, ys = np.arange(0, c["width"], stride, dtype=np.uint32), np.arange(0, c["height"], stride, dtype=np.uint32)
        px, py = np.meshgrid(xs, ys)
        px, py = px.reshape(-1), py.reshape(-1)
        cam = self.cam.astype(np.float32)
        eye, U, V, W = cam[0:3], cam[3:6], cam[6:9], cam[9:12]
        # cameraCommon.cuh:22-25 in fp32: d = pixel / size * 2 - 1; dir = normalize(d.x*U + d.y*V + W)
        dx = (px.astype(np.float32) / np.float32(c["width"]) * np.float32(2) - np.float32(1)).astype(np.float32)
        dy = (py.astype(np.float32) / np.float32(c["height"]) * np.float32(2) - np.float32(1)).astype(np.float32)
        d = (dx[:, None] * U[None, :] + dy[:, None] * V[None, :] + W[None, :]).astype(np.float32)
        inv = (np.float32(1) / np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]).astype(np.float32))).astype(np.float32)
        d = (d * inv[:, None]).astype(np.float32)
        o = np.tile(eye, (len(px), 1)).astype(np.float32)
        val0 = (px * np.uint32(4096) + py).astype(np.uint32)
        stream = np.full(len(px), subframe, dtype=np.uint32)
        return o, d, val0, stream

    def trace(self, stride: int, subframe: int):
        """One sample: returns (seconds, paths, radiance)."""
        o, d, val0, stream = self.rays(stride, subframe)
        t0 = time.perf_counter()
        if self.ref is not None:
            rad = self.ref.trace_paths(o, d, val0, stream, procs=self.cores)
        else:
            rad = self.o.trace_paths(self.c["mode"], o, d, val0, stream)
        return time.perf_counter() - t0, len(o), rad

    def pick_stride(self, seconds: float) -> int:
        """Largest sample (smallest stride) whose one-subframe trace fits `seconds`, from a coarse calibration."""
        c = self.c
        stride = max(1, int(max(c["width"], c["height"]) // 64))
        while True:
            dt, n, _ = self.trace(stride, 1)
            if dt >= 1.0 or stride == 1:  # long enough that fork / start-up overhead no longer dominates the estimate
                break
            stride = max(1, stride // 2)
        per_path = dt / max(1, n)
        want = max(64.0, seconds / max(per_path, 1e-9))
        s = (c["width"] * c["height"] / want) ** 0.5
        return max(1, int(s + 0.999))

    def sample_text(self, stride: int) -> str:
        c = self.c
        nx, ny = len(range(0, c["width"], stride)), len(range(0, c["height"], stride))
        return (f"pixels (i*{stride}, j*{stride}) of the {c['width']}x{c['height']} frame ({nx}x{ny} = {nx * ny} paths per subframe), "
                f"same camera rays and RNG streams as the GPU arm, 1 subframe per step")


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    est = CpuEstimator(a)
    stride = a.ref_stride or est.pick_stride(a.cpu_seconds)
    sub = 1
    for _ in range(a.warmup):
        est.trace(stride, sub)
        sub += 1
    total_t, paths = 0.0, 0
    for _ in range(a.steps):
        dt, n, _ = est.trace(stride, sub)
        sub += 1
        total_t += dt
        paths += n
    value = paths / total_t / 1e6
    sample = est.sample_text(stride)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": total_t / a.steps * 1e3, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(a),
        "cpu_baseline": {"value": value, "unit": METRIC, "cores": est.cores, "kind": est.kind, "sample": sample},
        "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": ("oracle/_ref: the reference's own estimator sources (CUDA/cloud.cuh, cloudRadianceMaterials.cu, random.cuh, cloudBBox.cu) compiled "
                 "unmodified for the host behind an OptiX emulation, one forked worker per core" if est.kind == "reference" else
                 "oracle port (oracle/ds_oracle.cpp, OpenMP): oracle/_ref is not built on this machine"),
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------- our arm

def load_json(path: Path):
    try:
        return json.loads(path.read_text())
    except Exception:
        return None


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import deepestscatter_b200 as ds
    from deepestscatter_b200 import multigpu

    c = a.cfg
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

    # per-rank share of a step's subframes and its RNG stream offset inside the step
    if a.scaling == "strong":
        offset, n_local = multigpu.subframe_range(rank, world, a.spp)
        n_total = a.spp
    else:
        offset, n_local, n_total = rank * a.spp, a.spp, world * a.spp
    ctx.set_option("staging_subframes", max(1, min(64, max(1, n_local))))

    # the library's own NCCL communicator: rank 0 draws the id, torch.distributed only ships its 128 bytes
    if distributed:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(ds.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, src=0)
        ctx.comm_init(world, rank, bytes(idt.cpu().numpy().tobytes()))

    # ---- inputs: resident in HBM before any timed region ----
    ctx.volume_synth(c["grid"], c["kind"], GRID_SEED, True)
    ctx.scene_set(c["size_m"], c["sun"])
    ctx.set_option("profile_events", 1)
    ctx.bake()
    ctx.sync()
    bake_us = ctx.get_option("bake_last_us")
    ctx.frame_create(c["width"], c["height"])
    cam = ds.camera_look_at(aspect=c["width"] / c["height"])
    px = c["width"] * c["height"]
    mode = c["mode"]

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(step_index: int):
        """One progressive frame: clear, this rank's subframes, the NCCL reduce of the accumulation buffers."""
        ctx.frame_clear()
        ctx.set_option("stream_offset", step_index * n_total + offset)
        if n_local:
            ctx.render_subframes(cam, mode, 1, n_local)
        if distributed:
            ctx.frame_reduce(n_local, n_total, 0)

    # ---- device-resident timed region ----
    step_index = 0
    with torch.cuda.stream(stream):
        for _ in range(a.warmup):
            device_step(step_index)
            step_index += 1
        ctx.sync()
        ctx.counters_reset()
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(a.steps):
            device_step(step_index)
            step_index += 1
        ev1.record(stream)
        barrier()
        clocks = sampler.stop()
        ms = ev0.elapsed_time(ev1)
    counters = ctx.counters()
    lstats = ctx.launch_stats()
    ctx.set_option("profile_events", 0)
    frame_mean = float(ctx.frame_download()[0][..., 0].astype(np.float64).mean()) if rank == 0 else 0.0

    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    keys = ("paths", "events", "steps", "density_taps", "untraced_paths", "untraced_steps")
    agg = torch.tensor([counters[k] for k in keys], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    ms_max = float(t.item())
    paths, events, steps, taps, untraced_paths, untraced_steps = (float(x) for x in agg.tolist())
    secs = ms_max * 1e-3
    value = paths / secs / 1e6

    # ---- end to end through the C ABI with HOST buffers (pinned): every step uploads the accumulation buffers, renders,
    # reduces over NCCL and downloads the finished frame on rank 0; copies inside the timed region ----
    # No host-side memset sits in the timed region: one GPU continues ONE progressive frame across the steps (the host holds the
    # accumulation state between calls, as a progressive renderer does: subframes 1..n in the first call, n+1..2n in the next); several
    # GPUs start every step's frame from a constant pinned all-zero pair and download the reduced frame into another pair.
    hp = torch.zeros(px * 4, dtype=torch.float32).pin_memory()
    hv = torch.zeros(px * 4, dtype=torch.float32).pin_memory()
    zp = torch.zeros(px * 4, dtype=torch.float32).pin_memory() if distributed else None
    zv = torch.zeros(px * 4, dtype=torch.float32).pin_memory() if distributed else None
    lib = ctx.lib
    e2e_done = [0]  # subframes already in the single-GPU host frame

    def e2e_step(step_index: int):
        if not distributed:
            ctx.set_option("stream_offset", 0)
            ctx.render_subframes_host_ptr(cam, mode, 1 + e2e_done[0], n_local, hp.data_ptr(), hv.data_ptr())  # H2D + render + D2H
            e2e_done[0] += n_local
            return
        ctx.set_option("stream_offset", step_index * n_total + offset)
        ctx._ck(lib.ds_frame_upload(ctx.h, zp.data_ptr(), zv.data_ptr()))  # zeros: a new frame starts from host state
        if n_local:
            ctx.render_subframes(cam, mode, 1, n_local)
        ctx.frame_reduce(n_local, n_total, 0)
        if rank == 0:
            ctx._ck(lib.ds_frame_download(ctx.h, hp.data_ptr(), hv.data_ptr()))
        else:
            ctx.sync()

    for _ in range(min(a.warmup, 2)):
        e2e_step(step_index)
        step_index += 1
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        e2e_step(step_index)
        step_index += 1
    barrier()
    e2e_s = time.perf_counter() - t0
    checksum_keep = float(hp.view(-1, 4)[::97, 0].double().mean()) if distributed and rank == 0 else 0.0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_paths = a.steps * n_total * px
    e2e_value = e2e_paths / float(te.item()) / 1e6
    checksum = float(hp.view(-1, 4)[:, 0].double().mean()) if not distributed else (checksum_keep if rank == 0 else 0.0)

    # ---- roofline of the dominant kernel (k_trace_fast) ----
    peaks = load_json(ROOT / "MEASURED_PEAKS.json")
    if peaks and "hbm_gbs" in peaks:
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    l2 = load_json(ROOT / "profiles" / "l2_peaks.json") or {}
    trace_launches = max(1, lstats["trace_launches_timed"])
    kernel_s = lstats["trace_ms_total"] * 1e-3 / trace_launches  # average launch duration, CUDA events inside the library
    per = lambda x: x / trace_launches  # noqa: E731  (rank-local counters: this rank's launches)
    k_steps = counters["steps"] - counters["untraced_steps"]  # what kernel threads counted (executed + leapt inside the kernel)
    alg_kernel = 8.0 * k_steps + 8.0 * counters["events"]
    alg_reference = 8.0 * counters["steps"] + 8.0 * counters["events"]  # SURVEY 8d on every march step of the reference algorithm
    alg_executed = 8.0 * counters["density_taps"] + 8.0 * counters["events"]  # taps actually fetched
    gbs = lambda b: per(b) / kernel_s / 1e9 if kernel_s > 0 else 0.0  # noqa: E731
    tex_taps_s = per(counters["density_taps"] + counters["events"]) / kernel_s if kernel_s > 0 else 0.0
    traffic = None
    tfile = load_json(ROOT / "profiles" / "traffic_k_trace_fast.json")
    if tfile and tfile.get("config") == {"grid": c["grid"], "width": c["width"], "height": c["height"], "spp": n_local, "precision": a.precision}:
        traffic = tfile["dram_bytes_read"] + tfile["dram_bytes_write"]
    # the region-major item order keeps even the 2.1 GB of C4 volumes L2-resident in effect (ncu: L2 hit rate 91 %, DRAM at 5 % of
    # its bandwidth, profiles/r02l_*), so the L2-resident texture rate is the denominator for every configuration
    tex_peak = (l2.get("tex3d_march_gtaps") or {}).get("l2_320")
    roofline = {
        "bound": "hbm", "achieved": gbs(alg_kernel), "peak": peak, "unit": "GB/s", "frac": gbs(alg_kernel) / peak, "traffic": traffic,
        "kernel": "k_trace_fast" if a.precision == "fast" else "k_trace", "kernel_ms_avg": kernel_s * 1e3,
        "kernel_share_of_step": lstats["trace_ms_total"] / ms if ms > 0 else None, "algorithmic_bytes_per_launch": per(alg_kernel),
        "peak_source": peak_src,
        "definition": "achieved = (8 B x march steps counted by the kernel's own threads + 8 B x scatter events) per launch / launch duration",
        "frac_reference_steps": gbs(alg_reference) / peak,
        "frac_executed": gbs(alg_executed) / peak,
        "executed_gbs": gbs(alg_executed),
        "l2_frac": gbs(alg_executed) / l2["l2_sector_gather_gbs"] if l2.get("l2_sector_gather_gbs") else None,
        "tex_taps_per_s": tex_taps_s,
        "tex_frac": tex_taps_s / (tex_peak * 1e9) if tex_peak else None,
        "tex_peak_gtaps": tex_peak, "l2_peaks": "profiles/l2_peaks.json (tools/microbench/l2_gather.cu)" if l2 else None,
        "note": ("frac: SURVEY 8d bytes over the steps the kernel itself accounts for; frac_reference_steps adds the steps of pixels the "
                 "primary-ray cache settles without tracing (host-side count, DsCounters.untraced_steps); frac_executed counts only taps "
                 "actually fetched (8 B each) -- the kernel is bound by instruction issue under divergence and by the texture path together "
                 "(ncu: profiles/r04g_*), not by HBM; tex_frac = fetched taps/s over the measured trilinear tex3D rate of an R8 volume for a "
                 "march-coherent access pattern at this residency -- the kernel reads a fused RG8 {density, sun transmittance} volume whose "
                 "sun taps are mostly L1 hits, which is how the fraction can pass 1"),
    }

    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(a),
        "run": {"precision": a.precision, "grid_bytes": 2 * c["grid"] ** 3, "subframes_per_step_this_rank": n_local,
                "multi_gpu": (f"replicated grid; the {n_total} subframes of a step split over {world} ranks; one ncclReduce (float64 moments, "
                              f"{px * 64} B) per step inside ds_frame_reduce") if distributed else "single GPU",
                "options": {k: ctx.get_option(k) for k in ("variant", "block_threads", "blocks_per_sm", "skip_empty", "primary_cache", "region_pixels",
                                                            "regen_min", "skip_min", "skip_max_iters", "march_keep32", "march_max_iters",
                                                            "march_unroll", "spec_percent", "fused_volume", "staging_subframes")}},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": METRIC, "h2d_bytes_per_step": world * 2 * px * 16, "d2h_bytes_per_step": 2 * px * 16,
                "api": ("ds_render_subframes_host (progressive + variance float4 buffers in pinned host memory; the steps continue one progressive frame)" if not distributed else
                        "per rank ds_frame_upload + ds_render_subframes + ds_frame_reduce (NCCL), rank 0 ds_frame_download; pinned host buffers"),
                "checksum_mean_radiance": checksum},
        "gpu_launches": lstats["kernel_launches"],
        "roofline": roofline,
        "hit_fraction": 1.0 - untraced_paths / paths if paths else None,
        "hit_mpaths_s": (paths - untraced_paths) / secs / 1e6,
        "events_per_s": events / secs, "steps_per_s": steps / secs, "density_taps_per_s": taps / secs,
        "events_per_path": events / paths, "events_per_hit_path": events / max(1.0, paths - untraced_paths), "steps_per_path": steps / paths,
        "frame_mean_radiance": frame_mean, "nonfinite": counters["nonfinite"],
    }

    # ---- CPU baseline: the reference estimator on the host cores, a bounded subset of the same pixels (rank 0, N = 1 only) ----
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            est = CpuEstimator(a)
            stride = est.pick_stride(a.cpu_seconds)
            dt, n, rad = est.trace(stride, 1)
            line["cpu_baseline"] = {"value": n / dt / 1e6, "unit": METRIC, "cores": est.cores, "kind": est.kind,
                                    "sample": est.sample_text(stride) + f" ({dt:.1f} s)", "mean_radiance": float(rad[:, 0].astype(np.float64).mean())}
        except Exception as exc:  # the baseline is reporting only; never lose the GPU line
            line["cpu_baseline"] = {"value": None, "unit": METRIC, "cores": host_threads(), "kind": "port", "sample": f"failed: {exc}"}

    # ---- the other two kernels north_star names, with their own rooflines (rank 0, N = 1 only; outside every timed region) ----
    if rank == 0 and world == 1 and not a.no_secondary:
        sec = {}
        try:
            vox = c["grid"] ** 3
            # bake: 8 B per voxel-step (density tap) + 1 B per voxel written; the executed taps are what the counters cannot see here,
            # so the leg reports the voxel rate and the write-bound floor
            sec["bake"] = {"kernel": "k_bake", "us": bake_us, "voxels": vox, "gvoxels_per_s": vox / (bake_us * 1e-6) / 1e9 if bake_us else None,
                           "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak,
                                        "achieved": (vox * 2) / (bake_us * 1e-6) / 1e9 if bake_us else None,
                                        "frac": (vox * 2) / (bake_us * 1e-6) / 1e9 / peak if bake_us else None,
                                        "definition": "streaming floor: 1 B density read + 1 B written per voxel (taps along the sun ray are "
                                                      "cache hits of neighbouring voxels' reads); the kernel is tap/latency-bound, not HBM-bound"}}
            n_desc = 1 << 18
            pts, dirs = ctx.generate_points(0, 4096, 0)
            reps = n_desc // 4096
            pts, dirs = np.tile(pts, (reps, 1)), np.tile(dirs, (reps, 1))
            ctx.set_option("profile_events", 1)
            ctx.descriptors(pts, dirs)  # warm-up
            t0 = time.perf_counter()
            ctx.descriptors(pts, dirs)
            wall = time.perf_counter() - t0
            us = ctx.get_option("descriptors_last_us")
            ctx.set_option("profile_events", 0)
            alg = n_desc * (2250 * 16 + 2250)  # <= 16 B read per LOD tap + 2250 B written per sample
            sec["descriptors"] = {"kernel": "k_descriptors<EXACT>", "samples": n_desc, "us": us, "samples_per_s": n_desc / (us * 1e-6) if us else None,
                                  "e2e_samples_per_s": n_desc / wall, "e2e_api": "ds_collect_descriptors (host in, host out)",
                                  "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak, "achieved": alg / (us * 1e-6) / 1e9 if us else None,
                                               "frac": alg / (us * 1e-6) / 1e9 / peak if us else None,
                                               "definition": "2250 taps x 16 B (two mip levels) + 2250 B written per sample (upper bound: integral LODs read 8 B)"}}
        except Exception as exc:
            sec["failed"] = str(exc)
        if a.config == "C2":
            try:
                from deepestscatter_b200 import disney_model as dm

                ctx.disney_model_load(dm.synthetic_weights(566))
                ctx.render_disney(cam, c["width"], c["height"], stream=1)  # warm-up (scratch allocation, mip-mapped texture)
                times = []
                for rep in range(3):
                    t0 = time.perf_counter()
                    frame = ctx.render_disney(cam, c["width"], c["height"], stream=2 + rep)
                    times.append(time.perf_counter() - t0)
                rows = 1 << 17
                x = dm.synthetic_inputs(1024, 33)
                x = np.ascontiguousarray(np.tile(x, (rows // 1024, 1, 1)))
                ctx.set_option("profile_events", 1)
                ctx.disney_model_forward(x)
                ctx.disney_model_forward(x)
                us = ctx.get_option("mlp_last_us")
                macs = 10 * (226 * 200 + 2 * 200 * 200) - 200 * 200 + 2 * 200 * 200 + 200
                sec["neural_renderer"] = {
                    "ms_per_frame": min(times) * 1e3, "scattering_pixels": int((frame[..., 3] != 0).sum()),
                    "api": "ds_render_disney (DisneyRenderer::render; host frame buffer out), synthetic weights",
                    "model_kernel": "k_disney_mlp_tc (tcgen05 kind::f16 on IEEE half operands, fp32 accumulation; option mlp_fp16=0: kind::tf32)",
                    "model_rows": rows, "model_us": us, "model_tflops": 2 * macs * rows / (us * 1e-6) / 1e12 if us else None,
                }
            except Exception as exc:  # reporting only
                sec["neural_renderer"] = {"failed": str(exc)}
        line["secondary"] = sec

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
