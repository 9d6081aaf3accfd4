"""deepestscatter_b200 -- B200-native radiance-estimation hot path of DeepestScatter's DataGen.

The product is the C-ABI shared library `libdeepestscatter_b200.so` (include/ds_abi.h, built from
deepestscatter_b200/csrc/ for sm_100a).  This package only loads it and marshals numpy / torch buffers.
"""
from ._lib import LIB_PATH, build_library, load
from .dataset import Dataset
from .context import (
    MODE_ALL_SCATTER,
    MODE_MULTIPLE_SCATTER,
    MODE_SINGLE_SCATTER,
    PRECISION_EXACT,
    PRECISION_FAST,
    INFO_DTYPE,
    TASK_DTYPE,
    blit_predicted,
    Context,
    DsError,
    camera_array,
    camera_look_at,
    cloud_crop_active,
    cloud_read_vdb,
    comm_unique_id,
    record_disney_descriptor,
    record_result,
    record_scatter_sample,
    record_scene_setup,
)

__all__ = [
    "LIB_PATH", "build_library", "load", "Context", "DsError", "camera_look_at", "camera_array", "comm_unique_id",
    "MODE_ALL_SCATTER", "MODE_MULTIPLE_SCATTER", "MODE_SINGLE_SCATTER", "PRECISION_EXACT", "PRECISION_FAST", "TASK_DTYPE",
    "record_scatter_sample", "record_disney_descriptor", "record_result", "record_scene_setup", "Dataset", "cloud_crop_active", "cloud_read_vdb", "INFO_DTYPE", "blit_predicted",
]
