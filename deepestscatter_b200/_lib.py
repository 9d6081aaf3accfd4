"""ctypes binding of libdeepestscatter_b200.so (the C ABI declared in include/ds_abi.h).

The shared library is built in-tree by `make -C deepestscatter_b200/csrc` (see __graft_entry__.build).
There is no Python or CPU fallback: if the library is missing, or no sm_100 device is present,
loading / context creation raises.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libdeepestscatter_b200.so"
CSRC_DIR = PKG_DIR / "csrc"

DS_OK = 0


class DsSceneParams(C.Structure):
    _fields_ = [
        ("cloud_size_m", C.c_float),
        ("mean_free_path_m", C.c_float),
        ("sample_step", C.c_float),
        ("light_direction", C.c_float * 3),
        ("light_color", C.c_float * 3),
        ("light_intensity", C.c_float),
        ("minimal_ray_distance", C.c_float),
    ]


class DsCamera(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("U", C.c_float * 3), ("V", C.c_float * 3), ("W", C.c_float * 3)]


class DsPointRadianceTask(C.Structure):
    _fields_ = [
        ("id", C.c_int32),
        ("experiment_count", C.c_uint32),
        ("radiance", C.c_float),
        ("running_variance", C.c_float),
        ("position", C.c_float * 3),
        ("direction", C.c_float * 3),
    ]


class DsCounters(C.Structure):
    _fields_ = [
        ("paths", C.c_uint64),
        ("events", C.c_uint64),
        ("steps", C.c_uint64),
        ("density_taps", C.c_uint64),
        ("nonfinite", C.c_uint64),
        ("untraced_paths", C.c_uint64),
        ("untraced_steps", C.c_uint64),
    ]


class DsRadianceSettings(C.Structure):
    _fields_ = [
        ("max_thread_count", C.c_uint32),
        ("launches_per_update", C.c_uint32),
        ("max_updates", C.c_uint32),
        ("relative_ci", C.c_float),
        ("absolute_ci", C.c_float),
        ("zero_radiance_min_experiments", C.c_uint32),
    ]


# name -> (restype, argtypes); every symbol include/ds_abi.h declares
_vp, _i, _u32, _f, _sz = C.c_void_p, C.c_int, C.c_uint32, C.c_float, C.c_size_t
_pf = C.POINTER(C.c_float)
_pu8 = C.POINTER(C.c_uint8)
SIGNATURES = {
    "ds_context_create": (_i, [_i, C.POINTER(_vp)]),
    "ds_context_destroy": (_i, [_vp]),
    "ds_last_error": (C.c_char_p, [_vp]),
    "ds_context_set_stream": (_i, [_vp, _vp]),
    "ds_sync": (_i, [_vp]),
    "ds_set_option": (_i, [_vp, C.c_char_p, _i]),
    "ds_get_option": (_i, [_vp, C.c_char_p, C.POINTER(_i)]),
    "ds_get_counters": (_i, [_vp, C.POINTER(DsCounters)]),
    "ds_reset_counters": (_i, [_vp]),
    "ds_get_launch_stats": (_i, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    "ds_describe": (C.c_char_p, [_vp]),
    "ds_volume_upload": (_i, [_vp, _vp, _i, _i, _i, _i]),
    "ds_volume_upload_float": (_i, [_vp, _vp, _i, _i, _i, C.c_double, _i]),
    "ds_volume_synth": (_i, [_vp, _i, _i, _u32, _i]),
    "ds_volume_level_count": (_i, [_vp, C.POINTER(_i)]),
    "ds_volume_level_dims": (_i, [_vp, _i, C.POINTER(_i)]),
    "ds_volume_download_level": (_i, [_vp, _i, _vp]),
    "ds_scene_params_default": (None, [C.POINTER(DsSceneParams)]),
    "ds_scene_set": (_i, [_vp, C.POINTER(DsSceneParams)]),
    "ds_scene_get_derived": (_i, [_vp, _pf]),
    "ds_bake_sun_transmittance": (_i, [_vp]),
    "ds_inscatter_download": (_i, [_vp, _vp]),
    "ds_inscatter_upload": (_i, [_vp, _vp]),
    "ds_camera_look_at": (None, [_pf, _pf, _pf, _f, _f, C.POINTER(DsCamera)]),
    "ds_camera_default": (None, [_i, _i, C.POINTER(DsCamera)]),
    "ds_frame_create": (_i, [_vp, _i, _i]),
    "ds_frame_clear": (_i, [_vp]),
    "ds_render_frame_result": (_i, [_vp, C.POINTER(DsCamera), _i, _u32, _vp]),
    "ds_render_subframes": (_i, [_vp, C.POINTER(DsCamera), _i, _u32, _u32]),
    "ds_render_subframes_host": (_i, [_vp, C.POINTER(DsCamera), _i, _u32, _u32, _vp, _vp]),
    "ds_frame_download": (_i, [_vp, _vp, _vp]),
    "ds_frame_upload": (_i, [_vp, _vp, _vp]),
    "ds_frame_device_ptrs": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "ds_tonemap": (_i, [_vp, _f, _vp, _pf]),
    "ds_frame_unconverged": (_i, [_vp, _u32, C.POINTER(_u32)]),
    "ds_frame_export_moments_device": (_i, [_vp, _u32, _vp]),
    "ds_frame_import_moments_device": (_i, [_vp, _u32, _vp]),
    "ds_comm_unique_id": (_i, [_vp]),
    "ds_comm_init": (_i, [_vp, _i, _i, _vp]),
    "ds_comm_destroy": (_i, [_vp]),
    "ds_frame_reduce": (_i, [_vp, _u32, _u32, _i]),
    "ds_trace_paths": (_i, [_vp, _i, _u32, _vp, _vp, _vp, _vp, _vp]),
    "ds_generate_points": (_i, [_vp, _u32, _u32, _u32, _vp, _vp]),
    "ds_collect_descriptors": (_i, [_vp, _vp, _vp, _u32, _vp]),
    "ds_collect_descriptors_float": (_i, [_vp, _vp, _vp, _u32, _vp, _vp]),
    "ds_render_network_input": (_i, [_vp, C.POINTER(DsCamera), _u32, _u32, _u32, _u32, _u32, _u32, _u32, _vp, _vp]),
    "ds_blit_predicted": (_i, [_u32, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp]),
    "ds_disney_model_weight_count": (_sz, []),
    "ds_disney_model_load": (_i, [_vp, _vp, _sz]),
    "ds_disney_model_pack": (_i, [_vp, _sz, _i, _vp, _sz, _vp, _sz, C.POINTER(_sz), C.POINTER(_sz)]),
    "ds_disney_model_profile": (_i, [_vp, C.POINTER(C.c_uint64)]),
    "ds_invert_phase_cdf": (_i, [_vp, _vp, _u32, _vp, _vp]),
    "ds_disney_model_forward": (_i, [_vp, _vp, _u32, _vp]),
    "ds_render_disney": (_i, [_vp, C.POINTER(DsCamera), _u32, _u32, _u32, _vp]),
    "ds_render_disney_subframes": (_i, [_vp, C.POINTER(DsCamera), _u32, _u32]),
    "ds_radiance_settings_default": (None, [C.POINTER(DsRadianceSettings)]),
    "ds_point_radiance_run": (_i, [_vp, _vp, _vp, _u32, C.POINTER(DsRadianceSettings), _vp, _vp, C.POINTER(_u32)]),
    "ds_record_scatter_sample": (_i, [_pf, _pf, _vp, _sz]),
    "ds_record_disney_descriptor": (_i, [_vp, _sz, _vp, _sz]),
    "ds_record_result": (_i, [_f, _i, _vp, _sz]),
    "ds_record_scene_setup": (_i, [C.c_char_p, _f, _pf, _vp, _sz]),
    "ds_cloud_load": (_i, [_vp, C.c_char_p, _i, C.POINTER(_i)]),
    "ds_cloud_forget": (None, [_vp]),
    "ds_cloud_last_error": (C.c_char_p, []),
    "ds_cloud_crop_active": (_i, [_vp, _i, _i, _i, _vp, _sz, C.POINTER(_i), C.POINTER(C.c_double)]),
    "ds_cloud_read_vdb": (_i, [C.c_char_p, _vp, _sz, C.POINTER(_i), C.POINTER(C.c_double)]),
    "ds_write_exr": (_i, [C.c_char_p, _u32, _u32, _vp]),
    "ds_dataset_open": (_i, [C.c_char_p, C.POINTER(_vp)]),
    "ds_dataset_close": (_i, [_vp]),
    "ds_dataset_last_error": (C.c_char_p, [_vp]),
    "ds_dataset_commit": (_i, [_vp]),
    "ds_dataset_put": (_i, [_vp, C.c_char_p, C.c_int32, _vp, _sz]),
    "ds_dataset_get": (C.c_longlong, [_vp, C.c_char_p, C.c_int32, _vp, _sz]),
    "ds_dataset_count": (C.c_longlong, [_vp, C.c_char_p]),
    "ds_dataset_drop": (_i, [_vp, C.c_char_p]),
    "ds_dataset_merge": (_i, [_vp, C.c_char_p]),
    "ds_dataset_append_scene_setup": (_i, [_vp, C.c_int32, C.c_char_p, _f, _pf]),
    "ds_dataset_append_scatter_samples": (_i, [_vp, C.c_int32, _u32, _vp, _vp]),
    "ds_dataset_append_descriptors": (_i, [_vp, C.c_int32, _u32, _vp, _sz]),
    "ds_dataset_append_results": (_i, [_vp, C.c_int32, _u32, _vp, _vp]),
}


def build_library(force: bool = False) -> Path:
    """Compile the CUDA library in-tree (nvcc, sm_100a).  Raises on failure."""
    cmd = ["make", "-C", str(CSRC_DIR), "-j4"]
    if force:
        subprocess.run(["make", "-C", str(CSRC_DIR), "clean"], check=True, capture_output=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libdeepestscatter_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB_PATH


_lib = None


def load() -> C.CDLL:
    """Load the shared library and set the prototypes of every exported entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or make -C deepestscatter_b200/csrc).  There is no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
