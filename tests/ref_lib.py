"""ctypes binding of oracle/_ref/libds_ref.so: the REFERENCE'S OWN DataGen sources (device programs and host classes),
compiled unmodified from /root/reference against the OptiX emulation in oracle/ref_shim/.

TEST INFRASTRUCTURE: used by tests/ (to pin the oracle), tools/make_golden_ref.py and bench.py's reference arm only.
The library is built where /root/reference is mounted (this container); on the GPU box only the prebuilt .so exists.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
SHIM_DIR = ROOT / "oracle" / "ref_shim"
REF_SO = ROOT / "oracle" / "_ref" / "libds_ref.so"
REFERENCE_ROOT = Path(os.environ.get("DS_REFERENCE_ROOT", "/root/reference"))

MODE_ALL, MODE_MULTI, MODE_SINGLE = 0, 1, 2
COLLECT_NONE, COLLECT_SAMPLES, COLLECT_DESCRIPTORS, COLLECT_RADIANCE = 0, 1, 2, 3

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")

TASK_DTYPE = np.dtype(
    [("id", "<i4"), ("experimentCount", "<u4"), ("radiance", "<f4"), ("runningVariance", "<f4"), ("position", "<f4", 3), ("direction", "<f4", 3)]
)


def reference_mounted() -> bool:
    return (REFERENCE_ROOT / "DeepestScatter_DataGen" / "DeepestScatter_DataGen" / "src" / "CUDA" / "cloud.cuh").exists()


def build_ref(force: bool = False) -> Path | None:
    """make -C oracle/ref_shim when the reference tree is mounted; otherwise whatever prebuilt .so travelled here."""
    if reference_mounted():
        subprocess.run(["make", "-C", str(SHIM_DIR), f"REF_ROOT={REFERENCE_ROOT}"] + (["-B"] if force else []), check=True, capture_output=True)
    return REF_SO if REF_SO.exists() else None


def available() -> bool:
    try:
        return build_ref() is not None
    except subprocess.CalledProcessError:
        return False


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if build_ref() is None:
            raise RuntimeError("oracle/_ref/libds_ref.so is not built and /root/reference is not mounted")
        L = C.CDLL(str(REF_SO))
        L.ref_create.restype = C.c_void_p
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_volume_set_u8.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_volume_load.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.ref_volume_level_dims.argtypes = [C.c_int, C.POINTER(C.c_int)]
        L.ref_volume_level_get.argtypes = [C.c_int, _u8p]
        L.ref_volume_float_size.argtypes = [_f32p]
        L.ref_scene_init.argtypes = [C.c_void_p, C.c_float, C.c_float, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_scene_get_derived.argtypes = [C.c_void_p, _f32p]
        L.ref_get_mie.argtypes = [C.c_void_p, _f32p, _f32p, _f32p]
        L.ref_inscatter_get.argtypes = [C.c_void_p, _u8p]
        L.ref_inscatter_set.argtypes = [C.c_void_p, _u8p]
        L.ref_set_skip_bake.argtypes = [C.c_int]
        L.ref_trace_paths.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, _u32p, _u32p, _f32p, C.c_int]
        L.ref_camera_set.argtypes = [C.c_void_p, _f32p, _f32p, _f32p]
        L.ref_camera_get.argtypes = [C.c_void_p, _f32p]
        L.ref_render_frame_result.argtypes = [C.c_void_p, C.c_uint32, _f32p, C.c_int]
        L.ref_camera_update.argtypes = [C.c_void_p, C.c_int]
        L.ref_frame_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_average_luminance.argtypes = [C.c_void_p]
        L.ref_average_luminance.restype = C.c_float
        L.ref_camera_is_converged.argtypes = [C.c_void_p]
        L.ref_update_frame_result.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, C.c_uint32]
        L.ref_tonemap.argtypes = [C.c_void_p, _f32p, C.c_float, _u8p]
        L.ref_tonemap.restype = C.c_float
        L.ref_last_exr.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_generate_points.argtypes = [C.c_void_p, C.c_uint32, _f32p, _f32p]
        L.ref_dataset_put_samples.argtypes = [C.c_void_p, C.c_int, C.c_int, _f32p, _f32p]
        L.ref_dataset_get_descriptors.argtypes = [C.c_void_p, C.c_int, C.c_int, _u8p]
        L.ref_radiance_update.argtypes = [C.c_void_p, C.c_int, C.c_void_p, _u8p, C.c_void_p, C.POINTER(C.c_int)]
        L.ref_dataset_get_results.argtypes = [C.c_void_p, C.c_int, C.c_int, _f32p, _u8p]
        L.ref_network_input.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, _f32p, _f32p]
        L.ref_counters_get.argtypes = [C.POINTER(C.c_ulonglong)]
        _lib = L
    return _lib


def skip_bake(on: bool):
    lib().ref_set_skip_bake(int(on))


def f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def write_dense_container(path, grid: np.ndarray, origin=(0, 0, 0)) -> None:
    """The DSDENSE1 container oracle/ref_shim/include/openvdb/openvdb.h reads in place of a .vdb file."""
    g = np.ascontiguousarray(grid, dtype=np.float32)
    nz, ny, nx = g.shape
    with open(path, "wb") as f:
        f.write(b"DSDENSE1")
        f.write(struct.pack("<3i", nx, ny, nz))
        f.write(struct.pack("<3i", *origin))
        f.write(g.tobytes())


class Reference:
    """One reference scene (an emulated optix::Context plus the reference's Scene object graph)."""

    def __init__(self):
        self.L = lib()
        self.h = self.L.ref_create()
        self.w = self.h_px = 0
        self.batch = (0, 0)

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- volume ----
    def volume_upload(self, grid_u8: np.ndarray, build_mips: bool = True):
        g = np.ascontiguousarray(grid_u8, dtype=np.uint8)
        nz, ny, nx = g.shape
        assert self.L.ref_volume_set_u8(self.h, g.reshape(-1), nx, ny, nz, int(build_mips)) == 0

    def volume_import(self, grid_f32: np.ndarray, origin=(0, 0, 0), build_mips: bool = True):
        """Resources::loadVolumeBuffer on a dense float grid (crop to the active box + 1, /max*255, mips)."""
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "cloud.dsdense")
            write_dense_container(p, grid_f32, origin)
            assert self.L.ref_volume_load(self.h, p.encode(), int(build_mips)) == 0

    def level_count(self) -> int:
        return self.L.ref_volume_level_count()

    def level_dims(self, level: int):
        d = (C.c_int * 3)()
        self.L.ref_volume_level_dims(level, d)
        return d[0], d[1], d[2]

    def level(self, level: int) -> np.ndarray:
        nx, ny, nz = self.level_dims(level)
        out = np.empty(nx * ny * nz, dtype=np.uint8)
        self.L.ref_volume_level_get(level, out)
        return out.reshape(nz, ny, nx)

    def float_size(self) -> np.ndarray:
        o = np.empty(3, dtype=np.float32)
        self.L.ref_volume_float_size(o)
        return o

    # ---- scene ----
    def scene_init(self, cloud_size_m=7000.0, light_dir=(-0.03, -0.25, 0.8), sample_step=1.0 / 512.0, mode=MODE_ALL, width=8, height=8,
                   path_tracer=True, collector=COLLECT_NONE, batch_start=0, batch_size=0):
        rc = self.L.ref_scene_init(self.h, cloud_size_m, sample_step, f32(light_dir), mode, width, height, int(path_tracer), collector, batch_start,
                                   batch_size)
        assert rc == 0, "ref_scene_init failed"
        self.w, self.h_px, self.batch = width, height, (batch_start, batch_size)

    def derived(self) -> dict:
        o = np.empty(12, dtype=np.float32)
        self.L.ref_scene_get_derived(self.h, o)
        return dict(bbox=o[0:3].copy(), texture_scale=o[3:6].copy(), density_multiplier=float(o[6]), voxel_m=float(o[7]),
                    voxel_free_path=float(o[8]), light=o[9:12].copy())

    def mie(self):
        a, b, c = (np.empty(4096, dtype=np.float32) for _ in range(3))
        self.L.ref_get_mie(self.h, a, b, c)
        return a, b, c

    def inscatter(self) -> np.ndarray:
        nx, ny, nz = self.level_dims(0)
        out = np.empty(nx * ny * nz, dtype=np.uint8)
        self.L.ref_inscatter_get(self.h, out)
        return out.reshape(nz, ny, nx)

    def inscatter_set(self, vol: np.ndarray):
        """Install a sun-transmittance volume baked elsewhere (after scene_init under skip_bake(True)); benchmarks only."""
        assert self.L.ref_inscatter_set(self.h, np.ascontiguousarray(vol, dtype=np.uint8).reshape(-1)) == 0

    # ---- estimator ----
    def trace_paths(self, origins, dirs, seed_val0, stream, procs=1) -> np.ndarray:
        o, d = f32(origins).reshape(-1, 3), f32(dirs).reshape(-1, 3)
        n = len(o)
        out = np.empty((n, 3), dtype=np.float32)
        self.L.ref_trace_paths(self.h, n, o.reshape(-1), d.reshape(-1), np.ascontiguousarray(seed_val0, dtype=np.uint32),
                               np.ascontiguousarray(stream, dtype=np.uint32), out.reshape(-1), procs)
        return out

    def camera_set(self, eye, lookat=(0, 0, 0), up=(0, 1, 0)):
        self.L.ref_camera_set(self.h, f32(eye), f32(lookat), f32(up))

    def camera(self) -> np.ndarray:
        cam = np.empty(12, dtype=np.float32)
        self.L.ref_camera_get(self.h, cam)
        return cam

    def render_frame_result(self, subframe_id: int, procs=1) -> np.ndarray:
        out = np.empty((self.h_px, self.w, 4), dtype=np.float32)
        self.L.ref_render_frame_result(self.h, subframe_id, out.reshape(-1), procs)
        return out

    def camera_update(self, updates=1) -> int:
        return self.L.ref_camera_update(self.h, updates)

    def frame(self):
        p = np.empty((self.h_px, self.w, 4), dtype=np.float32)
        v = np.empty_like(p)
        s = np.empty((self.h_px, self.w, 4), dtype=np.uint8)
        self.L.ref_frame_get(self.h, p.ctypes.data, v.ctypes.data, s.ctypes.data, None)
        return p, v, s

    def average_luminance(self) -> float:
        return float(self.L.ref_average_luminance(self.h))

    def is_converged(self) -> bool:
        return bool(self.L.ref_camera_is_converged(self.h))

    def update_frame_result(self, frame_result, progressive, variance, subframe_id):
        p, v = f32(progressive).copy(), f32(variance).copy()
        self.L.ref_update_frame_result(self.h, f32(frame_result).reshape(-1), p.reshape(-1), v.reshape(-1), subframe_id)
        return p, v

    def tonemap(self, progressive, exposure=0.4):
        s = np.empty((self.h_px, self.w, 4), dtype=np.uint8)
        avg = self.L.ref_tonemap(self.h, f32(progressive).reshape(-1), exposure, s.reshape(-1))
        return s, float(avg)

    def last_exr(self):
        w, h, dec = C.c_int(), C.c_int(), C.c_int()
        if not self.L.ref_last_exr(None, C.byref(w), C.byref(h), C.byref(dec)):
            return None
        rgb = np.empty((h.value, w.value, 3), dtype=np.float32)
        self.L.ref_last_exr(rgb.ctypes.data, C.byref(w), C.byref(h), C.byref(dec))
        return rgb, bool(dec.value)

    # ---- collectors ----
    def generate_points(self, stream=0):
        n = self.batch[1]
        p, d = np.empty((n, 3), dtype=np.float32), np.empty((n, 3), dtype=np.float32)
        assert self.L.ref_generate_points(self.h, stream, p.reshape(-1), d.reshape(-1)) == 0
        return p, d

    def put_samples(self, positions, directions, start_id=0):
        p, d = f32(positions).reshape(-1, 3), f32(directions).reshape(-1, 3)
        self.L.ref_dataset_put_samples(self.h, start_id, len(p), p.reshape(-1), d.reshape(-1))

    def descriptors(self, start_id, n) -> np.ndarray:
        out = np.empty((n, 2250), dtype=np.uint8)
        assert self.L.ref_dataset_get_descriptors(self.h, start_id, n, out.reshape(-1)) == 0
        return out

    def radiance_update(self, updates=1):
        n = self.batch[1]
        tasks = np.zeros(n, dtype=TASK_DTYPE)
        conv = np.zeros(n, dtype=np.uint8)
        threads = np.zeros(20480, dtype=TASK_DTYPE)
        recorded = C.c_int()
        done = self.L.ref_radiance_update(self.h, updates, tasks.ctypes.data, conv, threads.ctypes.data, C.byref(recorded))
        assert done >= 0
        return done, tasks, conv, threads, recorded.value

    def results(self, start_id, n):
        v, c = np.empty(n, dtype=np.float32), np.empty(n, dtype=np.uint8)
        rc = self.L.ref_dataset_get_results(self.h, start_id, n, v, c)
        return (v, c) if rc == 0 else None

    def network_input(self, rect_x, rect_y, rect_w, rect_h, stream):
        inp = np.empty((rect_h, rect_w, 10, 226), dtype=np.float32)
        info = np.empty((rect_h, rect_w, 5), dtype=np.float32)
        assert self.L.ref_network_input(self.h, rect_x, rect_y, rect_w, rect_h, stream, inp.reshape(-1), info.reshape(-1)) == 0
        return inp, info

    def counters(self) -> dict:
        c = (C.c_ulonglong * 3)()
        self.L.ref_counters_get(c)
        return dict(paths=int(c[0]), events=int(c[1]), steps=int(c[2]))

    def counters_reset(self):
        self.L.ref_counters_reset()
