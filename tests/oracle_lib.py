"""ctypes binding of the host oracle (oracle/_build/libds_oracle.so).

TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs only.  The product package never imports it.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
ORACLE_SO = ORACLE_DIR / "_build" / "libds_oracle.so"
MIE_PATH = ROOT / "deepestscatter_b200" / "data" / "mie_tables.f32"

MODE_ALL, MODE_MULTI, MODE_SINGLE = 0, 1, 2

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")

TASK_DTYPE = np.dtype(
    [
        ("id", "<i4"),
        ("experimentCount", "<u4"),
        ("radiance", "<f4"),
        ("runningVariance", "<f4"),
        ("position", "<f4", 3),
        ("direction", "<f4", 3),
    ]
)
assert TASK_DTYPE.itemsize == 40


def build_oracle(force: bool = False) -> Path:
    srcs = [ORACLE_DIR / "ds_oracle.cpp", ORACLE_DIR / "ds_oracle_mlp.cpp", ROOT / "include" / "ds_detmath.h", ROOT / "include" / "ds_synth.h"]
    if force or not ORACLE_SO.exists() or any(s.stat().st_mtime > ORACLE_SO.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True, capture_output=True)
    return ORACLE_SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(str(ORACLE_SO))
        L.orc_create.restype = C.c_void_p
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_mie.argtypes = [C.c_void_p, _f32p, _f32p]
        L.orc_get_mie.argtypes = [C.c_void_p, _f32p, _f32p, _f32p]
        L.orc_volume_set.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_quantize_float_grid.argtypes = [_f32p, C.c_size_t, C.c_double, _u8p]
        L.orc_volume_synth.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_int]
        L.orc_volume_level_count.argtypes = [C.c_void_p]
        L.orc_volume_level_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.orc_volume_level_get.argtypes = [C.c_void_p, C.c_int, _u8p]
        L.orc_network_input.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, _f32p, _f32p]
        L.orc_disney_weight_count.restype = C.c_size_t
        L.orc_disney_forward.argtypes = [_f32p, _f32p, C.c_int, _f32p, C.c_void_p]
        L.orc_scene_set.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, _f32p, _f32p, C.c_float]
        L.orc_scene_get_derived.argtypes = [C.c_void_p, _f32p]
        L.orc_bake_inscatter.argtypes = [C.c_void_p]
        L.orc_bake_inscatter_skip.argtypes = [C.c_void_p]
        L.orc_inscatter_get.argtypes = [C.c_void_p, _u8p]
        L.orc_inscatter_set.argtypes = [C.c_void_p, _u8p]
        L.orc_sample_volume.argtypes = [C.c_void_p, C.c_int, _f32p, C.c_int, C.c_float, _f32p]
        L.orc_sample_table.argtypes = [C.c_void_p, C.c_int, _f32p, C.c_int, _f32p]
        L.orc_math_probe.argtypes = [C.c_int, _f32p, C.c_int, _f32p, _f32p]
        L.orc_rng_probe.argtypes = [_u32p, _u32p, C.c_int, C.c_int, _u32p, _f32p]
        L.orc_new_directions.argtypes = [C.c_void_p, _u32p, _u32p, _f32p, C.c_int, _f32p]
        L.orc_camera_look_at.argtypes = [_f32p, _f32p, _f32p, C.c_float, C.c_float, _f32p]
        L.orc_trace_paths.argtypes = [C.c_void_p, C.c_int, C.c_int, _f32p, _f32p, _u32p, _u32p, _f32p]
        L.orc_render_frame_result.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int, C.c_uint32, _f32p]
        L.orc_update_frame_result.argtypes = [_f32p, _f32p, _f32p, C.c_size_t, C.c_uint32]
        L.orc_render_accumulate.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, _f32p, _f32p]
        L.orc_unconverged_pixels.argtypes = [_f32p, _f32p, C.c_size_t, C.c_uint32]
        L.orc_unconverged_pixels.restype = C.c_uint32
        L.orc_tonemap.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, _u8p]
        L.orc_tonemap.restype = C.c_float
        L.orc_generate_points.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, _f32p, _f32p]
        L.orc_descriptors.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_point_radiance.argtypes = [
            C.c_void_p, _f32p, _f32p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, _u8p, C.POINTER(C.c_uint32),
        ]
        L.orc_point_radiance.restype = C.c_int
        L.orc_counters_get.argtypes = [C.POINTER(C.c_ulonglong)]
        L.orc_counters_reset.argtypes = []
        L.orc_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def load_mie_tables() -> tuple[np.ndarray, np.ndarray]:
    raw = np.fromfile(MIE_PATH, dtype="<f4")
    assert raw.size == 8192
    return raw[:4096].copy(), raw[4096:].copy()


def f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


class Oracle:
    """Handle-based wrapper mirroring the product's Context API (same method names)."""

    def __init__(self):
        self.L = lib()
        self.h = self.L.orc_create()
        mie, chopped = load_mie_tables()
        self.L.orc_set_mie(self.h, mie, chopped)

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
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
        self.L.orc_volume_set(self.h, g.reshape(-1), nx, ny, nz, int(build_mips))

    def volume_synth(self, n: int, kind: int = 0, seed: int = 1234, build_mips: bool = True):
        self.L.orc_volume_synth(self.h, n, kind, seed, int(build_mips))

    def level_count(self) -> int:
        return self.L.orc_volume_level_count(self.h)

    def level_dims(self, level: int) -> tuple[int, int, int]:
        d = (C.c_int * 3)()
        self.L.orc_volume_level_dims(self.h, level, d)
        return d[0], d[1], d[2]

    def level(self, level: int) -> np.ndarray:
        nx, ny, nz = self.level_dims(level)
        out = np.empty(nx * ny * nz, dtype=np.uint8)
        self.L.orc_volume_level_get(self.h, level, out)
        return out.reshape(nz, ny, nx)

    # ---- scene ----
    def scene_set(self, cloud_size_m=7000.0, light_dir=(-0.03, -0.25, 0.8), mean_free_path_m=10.0,
                  sample_step=1.0 / 512.0, light_color=(1, 1, 1), light_intensity=1e6):
        self.L.orc_scene_set(self.h, cloud_size_m, mean_free_path_m, sample_step, f32(light_dir), f32(light_color), light_intensity)

    def derived(self) -> dict:
        o = np.empty(12, dtype=np.float32)
        self.L.orc_scene_get_derived(self.h, o)
        return dict(bbox=o[0:3].copy(), texture_scale=o[3:6].copy(), density_multiplier=float(o[6]),
                    voxel_m=float(o[7]), voxel_free_path=float(o[8]), light=o[9:12].copy())

    def bake(self, skip_empty: bool = False):
        (self.L.orc_bake_inscatter_skip if skip_empty else self.L.orc_bake_inscatter)(self.h)

    def inscatter(self) -> np.ndarray:
        nx, ny, nz = self.level_dims(0)
        out = np.empty(nx * ny * nz, dtype=np.uint8)
        self.L.orc_inscatter_get(self.h, out)
        return out.reshape(nz, ny, nx)

    def inscatter_set(self, vol: np.ndarray):
        self.L.orc_inscatter_set(self.h, np.ascontiguousarray(vol, dtype=np.uint8).reshape(-1))

    def sample_volume(self, pos, which=0, lod=-1.0) -> np.ndarray:
        p = f32(pos).reshape(-1, 3)
        out = np.empty(len(p), dtype=np.float32)
        self.L.orc_sample_volume(self.h, which, p.reshape(-1), len(p), lod, out)
        return out

    def sample_table(self, which, u) -> np.ndarray:
        u = f32(u).reshape(-1)
        out = np.empty_like(u)
        self.L.orc_sample_table(self.h, which, u, len(u), out)
        return out

    def mie_tables(self):
        a, b, c = (np.empty(4096, dtype=np.float32) for _ in range(3))
        self.L.orc_get_mie(self.h, a, b, c)
        return a, b, c

    def new_directions(self, val0, stream, prev) -> np.ndarray:
        p = f32(prev).reshape(-1, 3)
        out = np.empty_like(p)
        self.L.orc_new_directions(self.h, np.ascontiguousarray(val0, dtype=np.uint32), np.ascontiguousarray(stream, dtype=np.uint32),
                                  p.reshape(-1), len(p), out.reshape(-1))
        return out

    # ---- estimators ----
    def trace_paths(self, mode, origins, dirs, seed_val0, stream) -> np.ndarray:
        o, d = f32(origins).reshape(-1, 3), f32(dirs).reshape(-1, 3)
        n = len(o)
        out = np.empty((n, 3), dtype=np.float32)
        self.L.orc_trace_paths(self.h, mode, n, o.reshape(-1), d.reshape(-1),
                               np.ascontiguousarray(seed_val0, dtype=np.uint32), np.ascontiguousarray(stream, dtype=np.uint32),
                               out.reshape(-1))
        return out

    def render_frame(self, cam, w, h, mode, subframe) -> np.ndarray:
        out = np.empty((h, w, 4), dtype=np.float32)
        self.L.orc_render_frame_result(self.h, f32(cam), w, h, mode, subframe, out.reshape(-1))
        return out

    def render_accumulate(self, cam, w, h, mode, first, n, progressive=None, variance=None):
        if progressive is None:
            progressive = np.zeros((h, w, 4), dtype=np.float32)
            variance = np.zeros((h, w, 4), dtype=np.float32)
        self.L.orc_render_accumulate(self.h, f32(cam), w, h, mode, first, n, progressive.reshape(-1), variance.reshape(-1))
        return progressive, variance

    def generate_points(self, first_index, n, stream=0):
        p = np.empty((n, 3), dtype=np.float32)
        d = np.empty((n, 3), dtype=np.float32)
        self.L.orc_generate_points(self.h, first_index, n, stream, p.reshape(-1), d.reshape(-1))
        return p, d

    def descriptors(self, pos, dirs, as_float=False, want_index=False):
        p, d = f32(pos).reshape(-1, 3), f32(dirs).reshape(-1, 3)
        n = len(p)
        out = np.empty((n, 10, 225), dtype=np.float32 if as_float else np.uint8)
        idx = np.empty((n, 10, 225, 4), dtype=np.int32) if want_index else None
        self.L.orc_descriptors(self.h, p.reshape(-1), d.reshape(-1), n, int(as_float), out.ctypes.data_as(C.c_void_p),
                               idx.ctypes.data_as(C.c_void_p) if want_index else None)
        return (out, idx) if want_index else out

    def network_input(self, cam, frame_w, frame_h, rect, stream=0):
        """orc_network_input: (input [h][w][10][226] float32, info [h][w][5] float32: r, g, b, transmittance, hasScattered)."""
        x, y, w, h = rect
        inp = np.empty((h, w, 10, 226), dtype=np.float32)
        info = np.empty((h, w, 5), dtype=np.float32)
        self.L.orc_network_input(self.h, f32(cam).reshape(-1), frame_w, frame_h, x, y, w, h, stream, inp.reshape(-1), info.reshape(-1))
        return inp, info

    def point_radiance(self, pos, dirs, max_threads=20480, launches_per_update=100, max_updates=1000):
        p, d = f32(pos).reshape(-1, 3), f32(dirs).reshape(-1, 3)
        n = len(p)
        tasks = np.zeros(n, dtype=TASK_DTYPE)
        conv = np.zeros(n, dtype=np.uint8)
        upd = C.c_uint32(0)
        nconv = self.L.orc_point_radiance(self.h, p.reshape(-1), d.reshape(-1), n, max_threads, launches_per_update, max_updates,
                                          tasks.ctypes.data_as(C.c_void_p), conv, C.byref(upd))
        return tasks, conv.astype(bool), nconv, upd.value

    def counters(self) -> dict:
        c = (C.c_ulonglong * 3)()
        self.L.orc_counters_get(c)
        return dict(paths=c[0], events=c[1], steps=c[2])

    def counters_reset(self):
        self.L.orc_counters_reset()


def camera_look_at(eye=(2.5, -0.4, 0.0), lookat=(0, 0, 0), up=(0, 1, 0), hfov=30.0, aspect=2.0) -> np.ndarray:
    """Camera::init / updatePosition defaults (Camera.cpp:37-39,102)."""
    cam = np.empty(12, dtype=np.float32)
    lib().orc_camera_look_at(f32(eye), f32(lookat), f32(up), hfov, aspect, cam)
    return cam


def tonemap(progressive: np.ndarray, exposure: float = 0.4):
    h, w, _ = progressive.shape
    out = np.empty((h, w, 4), dtype=np.uint8)
    avg = lib().orc_tonemap(f32(progressive).reshape(-1), w, h, exposure, out.reshape(-1))
    return out, avg


def disney_forward(weights: np.ndarray, inputs: np.ndarray, want_hidden: bool = False):
    """Oracle restatement of the reference's DisneyModel.forward (oracle/ds_oracle_mlp.cpp): inputs [n][10][226] -> [n]."""
    L = lib()
    w = np.ascontiguousarray(weights, np.float32)
    x = np.ascontiguousarray(inputs, np.float32)
    assert w.size == L.orc_disney_weight_count() and x.shape[1:] == (10, 226)
    n = x.shape[0]
    out = np.zeros(n, np.float32)
    hidden = np.zeros((n, 200), np.float32) if want_hidden else None
    L.orc_disney_forward(w, x, n, out, hidden.ctypes.data if want_hidden else None)
    return (out, hidden) if want_hidden else out
