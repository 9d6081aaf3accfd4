"""Thin numpy-facing wrapper over the C ABI (one `Context` per GPU).

Only plumbing lives here: argument marshalling and error translation.  All arithmetic runs in the
sm_100a kernels behind include/ds_abi.h.
"""
from __future__ import annotations

import ctypes as C
import json

import numpy as np

from . import _lib
from ._lib import DsCamera, DsCounters, DsPointRadianceTask, DsRadianceSettings, DsSceneParams

MODE_ALL_SCATTER = 0       # Cloud::Rendering::Mode::SunAndSkyAllScatter -> totalRadiance
MODE_MULTIPLE_SCATTER = 1  # SunMultipleScatter -> multipleScatterSunRadiance
MODE_SINGLE_SCATTER = 2    # SunSingleScatter -> singleScatterSunRadiance
PRECISION_EXACT = 0
PRECISION_FAST = 1

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


INFO_DTYPE = np.dtype([("radiance", "<f4", 3), ("transmittance", "<f4"), ("hasScattered", "<u4")])  # IntersectionInfo, rayData.cuh:28-33


def blit_predicted(frame_result: np.ndarray, rect, predicted: np.ndarray, info: np.ndarray):
    """copyToFrameResult (disneyCamera.cu:38-46) into a float4 [H][W] image, in place."""
    lib = _lib.load()
    hgt, wid = frame_result.shape[:2]
    x, y, w, h = rect
    pr = np.ascontiguousarray(predicted, dtype=np.float32).reshape(h, w)
    inf = np.ascontiguousarray(info, dtype=INFO_DTYPE).reshape(h, w)
    assert frame_result.dtype == np.float32 and frame_result.flags.c_contiguous and frame_result.shape[2] == 4
    rc = lib.ds_blit_predicted(wid, hgt, x, y, w, h, _ptr(pr), _ptr(inf), _ptr(frame_result))
    if rc != 0:
        raise DsError(rc, "ds_blit_predicted: bad rectangle")


class DsError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def camera_look_at(eye=(2.5, -0.4, 0.0), lookat=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), hfov=30.0, aspect=2.0) -> DsCamera:
    """sutil::calculateCameraVariables with the defaults of Camera::init (Camera.cpp:37-39,102)."""
    lib = _lib.load()
    cam = DsCamera()
    f3 = C.c_float * 3
    lib.ds_camera_look_at(f3(*eye), f3(*lookat), f3(*up), hfov, aspect, C.byref(cam))
    return cam


def comm_unique_id() -> bytes:
    """ncclGetUniqueId (rank 0); ship the 128 bytes to the other ranks by any means (torch.distributed broadcast, a file, MPI)."""
    lib = _lib.load()
    buf = (C.c_uint8 * 128)()
    rc = lib.ds_comm_unique_id(buf)
    if rc != 0:
        raise DsError(rc, (lib.ds_last_error(None) or b"").decode())
    return bytes(buf)


def camera_array(cam: DsCamera) -> np.ndarray:
    return np.array(list(cam.eye) + list(cam.U) + list(cam.V) + list(cam.W), dtype=np.float32)


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch.as_tensor can alias library-owned memory."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


class Context:
    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.ds_context_create(device, C.byref(h))
        if rc != 0:
            raise DsError(rc, (self.lib.ds_last_error(None) or b"").decode())
        self.h = h
        self.device = device
        self.width = self.height = 0

    # ---- plumbing ----
    def _ck(self, rc: int):
        if rc != 0:
            raise DsError(rc, (self.lib.ds_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.ds_context_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_stream(self, cuda_stream_ptr: int):
        self._ck(self.lib.ds_context_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def sync(self):
        self._ck(self.lib.ds_sync(self.h))

    def set_option(self, name: str, value: int):
        self._ck(self.lib.ds_set_option(self.h, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        v = C.c_int()
        self._ck(self.lib.ds_get_option(self.h, name.encode(), C.byref(v)))
        return v.value

    def describe(self) -> dict:
        return json.loads(self.lib.ds_describe(self.h).decode())

    def counters(self) -> dict:
        c = DsCounters()
        self._ck(self.lib.ds_get_counters(self.h, C.byref(c)))
        return dict(paths=c.paths, events=c.events, steps=c.steps, density_taps=c.density_taps, nonfinite=c.nonfinite,
                    untraced_paths=c.untraced_paths, untraced_steps=c.untraced_steps)

    def counters_reset(self):
        self._ck(self.lib.ds_reset_counters(self.h))

    def launch_stats(self) -> dict:
        a, b, ms = C.c_uint64(), C.c_uint64(), C.c_double()
        self._ck(self.lib.ds_get_launch_stats(self.h, C.byref(a), C.byref(b), C.byref(ms)))
        return dict(kernel_launches=a.value, trace_launches_timed=b.value, trace_ms_total=ms.value)

    # ---- volume ----
    def volume_upload(self, grid_u8: np.ndarray, build_mips: bool = True):
        g = np.ascontiguousarray(grid_u8, dtype=np.uint8)
        nz, ny, nx = g.shape
        self._ck(self.lib.ds_volume_upload(self.h, _ptr(g), nx, ny, nz, int(build_mips)))

    def volume_upload_float(self, grid_f32: np.ndarray, max_density: float, build_mips: bool = True):
        g = _f32(grid_f32)
        nz, ny, nx = g.shape
        self._ck(self.lib.ds_volume_upload_float(self.h, _ptr(g), nx, ny, nz, float(max_density), int(build_mips)))

    def volume_synth(self, n: int, kind: int = 0, seed: int = 1234, build_mips: bool = True):
        self._ck(self.lib.ds_volume_synth(self.h, n, kind, seed, int(build_mips)))

    def cloud_load(self, path: str, build_mips: bool = True):
        """Resources::loadVolumeBuffer front end: dense .npy grid or synth:<n> spec; returns (nx, ny, nz)."""
        size = (C.c_int * 3)()
        rc = self.lib.ds_cloud_load(self.h, str(path).encode(), 1 if build_mips else 0, size)
        if rc != 0:
            raise DsError(rc, (self.lib.ds_cloud_last_error() or b"").decode())
        return tuple(size)

    def level_count(self) -> int:
        v = C.c_int()
        self._ck(self.lib.ds_volume_level_count(self.h, C.byref(v)))
        return v.value

    def level_dims(self, level: int):
        d = (C.c_int * 3)()
        self._ck(self.lib.ds_volume_level_dims(self.h, level, d))
        return d[0], d[1], d[2]

    def level(self, level: int) -> np.ndarray:
        nx, ny, nz = self.level_dims(level)
        out = np.empty((nz, ny, nx), dtype=np.uint8)
        self._ck(self.lib.ds_volume_download_level(self.h, level, _ptr(out)))
        return out

    # ---- scene ----
    def scene_set(self, cloud_size_m=7000.0, light_dir=(-0.03, -0.25, 0.8), mean_free_path_m=10.0, sample_step=1.0 / 512.0,
                  light_color=(1.0, 1.0, 1.0), light_intensity=1e6):
        p = DsSceneParams()
        self.lib.ds_scene_params_default(C.byref(p))
        p.cloud_size_m = cloud_size_m
        p.mean_free_path_m = mean_free_path_m
        p.sample_step = sample_step
        p.light_direction[:] = list(map(float, light_dir))
        p.light_color[:] = list(map(float, light_color))
        p.light_intensity = light_intensity
        self._ck(self.lib.ds_scene_set(self.h, C.byref(p)))

    def derived(self) -> dict:
        o = np.empty(12, dtype=np.float32)
        self._ck(self.lib.ds_scene_get_derived(self.h, o.ctypes.data_as(C.POINTER(C.c_float))))
        return dict(bbox=o[0:3].copy(), texture_scale=o[3:6].copy(), density_multiplier=float(o[6]), voxel_m=float(o[7]),
                    voxel_free_path=float(o[8]), light=o[9:12].copy())

    def bake(self):
        self._ck(self.lib.ds_bake_sun_transmittance(self.h))

    def inscatter(self) -> np.ndarray:
        nx, ny, nz = self.level_dims(0)
        out = np.empty((nz, ny, nx), dtype=np.uint8)
        self._ck(self.lib.ds_inscatter_download(self.h, _ptr(out)))
        return out

    def inscatter_set(self, vol: np.ndarray):
        v = np.ascontiguousarray(vol, dtype=np.uint8)
        self._ck(self.lib.ds_inscatter_upload(self.h, _ptr(v)))

    # ---- progressive renderer ----
    def frame_create(self, width: int, height: int):
        self._ck(self.lib.ds_frame_create(self.h, width, height))
        self.width, self.height = width, height

    def frame_clear(self):
        self._ck(self.lib.ds_frame_clear(self.h))

    def render_frame(self, cam: DsCamera, mode: int, subframe: int) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._ck(self.lib.ds_render_frame_result(self.h, C.byref(cam), mode, subframe, _ptr(out)))
        return out

    def render_subframes(self, cam: DsCamera, mode: int, first: int, n: int):
        self._ck(self.lib.ds_render_subframes(self.h, C.byref(cam), mode, first, n))

    def render_subframes_host(self, cam: DsCamera, mode: int, first: int, n: int, progressive: np.ndarray, variance: np.ndarray):
        assert progressive.dtype == np.float32 and variance.dtype == np.float32
        self._ck(self.lib.ds_render_subframes_host(self.h, C.byref(cam), mode, first, n, _ptr(progressive), _ptr(variance)))

    def render_subframes_host_ptr(self, cam: DsCamera, mode: int, first: int, n: int, progressive_ptr: int, variance_ptr: int):
        """Same call with raw host addresses (e.g. pinned torch tensors' data_ptr())."""
        self._ck(self.lib.ds_render_subframes_host(self.h, C.byref(cam), mode, first, n, C.c_void_p(progressive_ptr), C.c_void_p(variance_ptr)))

    def frame_download(self):
        p = np.empty((self.height, self.width, 4), dtype=np.float32)
        v = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._ck(self.lib.ds_frame_download(self.h, _ptr(p), _ptr(v)))
        return p, v

    def frame_upload(self, progressive: np.ndarray, variance: np.ndarray):
        p, v = _f32(progressive), _f32(variance)
        self._ck(self.lib.ds_frame_upload(self.h, _ptr(p), _ptr(v)))
        self.sync()

    def frame_device_arrays(self):
        """(progressive, variance) as __cuda_array_interface__ objects aliasing the device buffers."""
        a, b = C.c_void_p(), C.c_void_p()
        self._ck(self.lib.ds_frame_device_ptrs(self.h, C.byref(a), C.byref(b)))
        shape = (self.height, self.width, 4)
        return _DevArray(a.value, shape, "<f4"), _DevArray(b.value, shape, "<f4")

    def tonemap(self, exposure: float = 0.4):
        out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        avg = C.c_float()
        self._ck(self.lib.ds_tonemap(self.h, exposure, _ptr(out), C.byref(avg)))
        return out, avg.value

    def unconverged(self, subframe: int) -> int:
        v = C.c_uint32()
        self._ck(self.lib.ds_frame_unconverged(self.h, subframe, C.byref(v)))
        return v.value

    def export_moments(self, n: int, device_ptr: int):
        self._ck(self.lib.ds_frame_export_moments_device(self.h, n, C.c_void_p(device_ptr)))

    def import_moments(self, n_total: int, device_ptr: int):
        self._ck(self.lib.ds_frame_import_moments_device(self.h, n_total, C.c_void_p(device_ptr)))

    # ---- multi-GPU accumulation-buffer reduce (NCCL inside the library) ----
    def comm_init(self, n_ranks: int, rank: int, unique_id: bytes):
        """ncclCommInitRank on this context's device; collective over all ranks.  `unique_id` = comm_unique_id() of rank 0."""
        assert len(unique_id) == 128
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ck(self.lib.ds_comm_init(self.h, n_ranks, rank, buf))

    def comm_destroy(self):
        self._ck(self.lib.ds_comm_destroy(self.h))

    def frame_reduce(self, n_local: int, n_total: int, root: int = 0):
        """Collective: combine the per-GPU accumulation buffers (one ncclReduce of float64 moments); root < 0 = all ranks."""
        self._ck(self.lib.ds_frame_reduce(self.h, n_local, n_total, root))

    # ---- generic paths ----
    def trace_paths(self, mode: int, origins, dirs, seed_val0, stream) -> np.ndarray:
        o, d = _f32(origins).reshape(-1, 3), _f32(dirs).reshape(-1, 3)
        s0 = np.ascontiguousarray(seed_val0, dtype=np.uint32)
        st = np.ascontiguousarray(stream, dtype=np.uint32)
        out = np.empty((len(o), 3), dtype=np.float32)
        self._ck(self.lib.ds_trace_paths(self.h, mode, len(o), _ptr(o), _ptr(d), _ptr(s0), _ptr(st), _ptr(out)))
        return out

    # ---- dataset generation ----
    def generate_points(self, first_index: int, n: int, stream: int = 0):
        p = np.empty((n, 3), dtype=np.float32)
        d = np.empty((n, 3), dtype=np.float32)
        self._ck(self.lib.ds_generate_points(self.h, first_index, n, stream, _ptr(p), _ptr(d)))
        return p, d

    def descriptors(self, pos, dirs, as_float: bool = False, want_index: bool = False):
        p, d = _f32(pos).reshape(-1, 3), _f32(dirs).reshape(-1, 3)
        n = len(p)
        if not as_float:
            out = np.empty((n, 10, 225), dtype=np.uint8)
            self._ck(self.lib.ds_collect_descriptors(self.h, _ptr(p), _ptr(d), n, _ptr(out)))
            return out
        out = np.empty((n, 10, 225), dtype=np.float32)
        idx = np.empty((n, 10, 225, 4), dtype=np.int32) if want_index else None
        self._ck(self.lib.ds_collect_descriptors_float(self.h, _ptr(p), _ptr(d), n, _ptr(out), _ptr(idx) if want_index else None))
        return (out, idx) if want_index else out

    def network_input(self, cam: DsCamera, frame_w: int, frame_h: int, rect, stream: int = 0):
        """DisneyRenderer::renderRect launch 0: returns (network input [h][w][10][226] float32, info [h][w] INFO_DTYPE)."""
        x, y, w, h = rect
        inp = np.empty((h, w, 10, 226), dtype=np.float32)
        info = np.empty((h, w), dtype=INFO_DTYPE)
        self._ck(self.lib.ds_render_network_input(self.h, C.byref(cam), frame_w, frame_h, x, y, w, h, stream, _ptr(inp), _ptr(info)))
        return inp, info

    def disney_model_load(self, weights):
        """Load DisneyModel.state_dict() as one flat float32 array (deepestscatter_b200.disney_model.flatten_state_dict)."""
        w = _f32(weights).reshape(-1)
        self._ck(self.lib.ds_disney_model_load(self.h, _ptr(w), w.size))

    def disney_model_forward(self, network_input) -> np.ndarray:
        """module->forward (DisneyRenderer.cpp:104): [n][10][226] -> [n] predicted radiance."""
        x = _f32(network_input).reshape(-1, 10, 226)
        out = np.empty(len(x), dtype=np.float32)
        self._ck(self.lib.ds_disney_model_forward(self.h, _ptr(x), len(x), _ptr(out)))
        return out

    def invert_phase_cdf(self, values: np.ndarray):
        """(cos_theta, phase): the FAST estimator's inversion of the chopped-Mie CDF and its half-precision phase sampler on `values`."""
        v = np.ascontiguousarray(values, dtype=np.float32).ravel()
        cos_t, phase = np.empty_like(v), np.empty_like(v)
        self._ck(self.lib.ds_invert_phase_cdf(self.h, _ptr(v), v.size, _ptr(cos_t), _ptr(phase)))
        return cos_t, phase

    def disney_model_profile(self) -> dict:
        """Cycle accounting of block 0 of the last tensor-core model launch (needs option profile_events = 1)."""
        c = (C.c_uint64 * 16)()
        self._ck(self.lib.ds_disney_model_profile(self.h, c))
        return {"issuer_total": c[0], "issuer_wait_weight_stage": c[1], "issuer_wait_epilogue": c[2], "issuer_wait_descriptor": c[3],
                "issuer_wait_weights": c[4], "worker_total": c[8], "worker_wait_descriptor_stage": c[9], "worker_wait_gemm": c[10],
                "worker_epilogues": c[11], "worker_staging": c[12]}

    def render_disney(self, cam: DsCamera, frame_w: int, frame_h: int, stream: int = 0) -> np.ndarray:
        """DisneyRenderer::render: the neural renderer's frameResultBuffer, float4 [h][w]."""
        out = np.empty((frame_h, frame_w, 4), dtype=np.float32)
        self._ck(self.lib.ds_render_disney(self.h, C.byref(cam), frame_w, frame_h, stream, _ptr(out)))
        return out

    def render_disney_subframes(self, cam: DsCamera, first_subframe: int, n: int):
        """Camera::render with DisneyRenderer: n neural subframes accumulated into the progressive / variance buffers."""
        self._ck(self.lib.ds_render_disney_subframes(self.h, C.byref(cam), first_subframe, n))

    def point_radiance(self, pos, dirs, max_threads: int = 20480, launches_per_update: int = 100, max_updates: int = 0):
        p, d = _f32(pos).reshape(-1, 3), _f32(dirs).reshape(-1, 3)
        n = len(p)
        s = DsRadianceSettings()
        self.lib.ds_radiance_settings_default(C.byref(s))
        s.max_thread_count = max_threads
        s.launches_per_update = launches_per_update
        s.max_updates = max_updates
        tasks = np.zeros(n, dtype=TASK_DTYPE)
        conv = np.zeros(n, dtype=np.uint8)
        upd = C.c_uint32()
        self._ck(self.lib.ds_point_radiance_run(self.h, _ptr(p), _ptr(d), n, C.byref(s), _ptr(tasks), _ptr(conv), C.byref(upd)))
        return tasks, conv.astype(bool), int(conv.sum()), upd.value


def cloud_read_vdb(path: str):
    """The .vdb front end alone (host only): (dense float32 grid [nz][ny][nx] over the active box + 1, maximum active value)."""
    lib = _lib.load()
    dims = (C.c_int * 3)()
    mx = C.c_double()
    rc = lib.ds_cloud_read_vdb(str(path).encode(), None, 0, dims, C.byref(mx))
    if rc != 0:
        raise DsError(rc, (lib.ds_cloud_last_error() or b"").decode())
    out = np.empty((dims[2], dims[1], dims[0]), dtype=np.float32)
    rc = lib.ds_cloud_read_vdb(str(path).encode(), _ptr(out), out.size, dims, C.byref(mx))
    if rc != 0:
        raise DsError(rc, (lib.ds_cloud_last_error() or b"").decode())
    return out, mx.value


def cloud_crop_active(dense: np.ndarray):
    """Active bounding box expanded by one voxel (Resources.cpp:97-101) of a dense (nz, ny, nx) float grid; host only."""
    lib = _lib.load()
    g = np.ascontiguousarray(dense, dtype=np.float32)
    nz, ny, nx = g.shape
    dims = (C.c_int * 3)()
    mx = C.c_double()
    rc = lib.ds_cloud_crop_active(_ptr(g), nx, ny, nz, None, 0, dims, C.byref(mx))
    if rc != 0:
        raise DsError(rc, (lib.ds_cloud_last_error() or b"").decode())
    out = np.empty((dims[2], dims[1], dims[0]), dtype=np.float32)
    rc = lib.ds_cloud_crop_active(_ptr(g), nx, ny, nz, _ptr(out), out.size, dims, C.byref(mx))
    if rc != 0:
        raise DsError(rc, (lib.ds_cloud_last_error() or b"").decode())
    return out, mx.value


# ---- records (host only; usable without a GPU) ----

def record_scatter_sample(point, view_direction) -> bytes:
    lib = _lib.load()
    buf = (C.c_uint8 * 64)()
    f3 = C.c_float * 3
    n = lib.ds_record_scatter_sample(f3(*map(float, point)), f3(*map(float, view_direction)), buf, 64)
    if n < 0:
        raise DsError(n, "ds_record_scatter_sample failed")
    return bytes(buf[:n])


def record_disney_descriptor(grid: bytes) -> bytes:
    lib = _lib.load()
    cap = len(grid) + 16
    buf = (C.c_uint8 * cap)()
    src = (C.c_uint8 * max(1, len(grid))).from_buffer_copy(grid if grid else b"\0")
    n = lib.ds_record_disney_descriptor(src, len(grid), buf, cap)
    if n < 0:
        raise DsError(n, "ds_record_disney_descriptor failed")
    return bytes(buf[:n])


def record_result(light_intensity: float, is_converged: bool) -> bytes:
    lib = _lib.load()
    buf = (C.c_uint8 * 16)()
    n = lib.ds_record_result(float(light_intensity), int(bool(is_converged)), buf, 16)
    if n < 0:
        raise DsError(n, "ds_record_result failed")
    return bytes(buf[:n])


def record_scene_setup(cloud_path: str, cloud_size_m: float, light_direction) -> bytes:
    lib = _lib.load()
    raw = cloud_path.encode("utf-8")
    cap = len(raw) + 64
    buf = (C.c_uint8 * cap)()
    f3 = C.c_float * 3
    n = lib.ds_record_scene_setup(raw, float(cloud_size_m), f3(*map(float, light_direction)), buf, cap)
    if n < 0:
        raise DsError(n, "ds_record_scene_setup failed")
    return bytes(buf[:n])
