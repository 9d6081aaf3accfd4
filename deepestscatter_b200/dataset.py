"""ctypes wrapper of the dataset store (ds_dataset_* in include/ds_abi.h): the LMDB data file the reference's
collectors append to (DG/Util/Dataset/Dataset.h) and DeepestScatter_Train/LmdbDataset.py reads."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .context import DsError

SCENE_SETUP, SCATTER_SAMPLE, DISNEY_DESCRIPTOR, RESULT = "SceneSetup", "ScatterSample", "DisneyDescriptor", "Result"
BATCH_SIZE = 2048  # DeepestScatter_Train/GlobalSettings.py:1; DG BatchSettings (Tasks.cpp:148)


class Dataset:
    def __init__(self, path: str):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.ds_dataset_open(str(path).encode(), C.byref(h))
        if rc != 0:
            raise DsError(rc, (self.lib.ds_dataset_last_error(None) or b"").decode())
        self.h = h

    def _check(self, rc):
        if rc < 0:
            raise DsError(int(rc), (self.lib.ds_dataset_last_error(self.h) or b"").decode())
        return rc

    def close(self):
        if self.h:
            h, self.h = self.h, None
            rc = self.lib.ds_dataset_close(h)
            if rc != 0:
                raise DsError(rc, (self.lib.ds_dataset_last_error(None) or b"").decode())

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def commit(self):
        self._check(self.lib.ds_dataset_commit(self.h))

    def put(self, table: str, record_id: int, data: bytes):
        buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data or b"\0")
        self._check(self.lib.ds_dataset_put(self.h, table.encode(), record_id, buf, len(data)))

    def get(self, table: str, record_id: int) -> bytes:
        n = self._check(self.lib.ds_dataset_get(self.h, table.encode(), record_id, None, 0))
        buf = (C.c_uint8 * max(1, n))()
        self._check(self.lib.ds_dataset_get(self.h, table.encode(), record_id, buf, n))
        return bytes(buf[:n])

    def count(self, table: str) -> int:
        return int(self._check(self.lib.ds_dataset_count(self.h, table.encode())))

    def drop(self, table: str):
        self._check(self.lib.ds_dataset_drop(self.h, table.encode()))

    def merge(self, other_path: str):
        self._check(self.lib.ds_dataset_merge(self.h, str(other_path).encode()))

    def append_scene_setup(self, scene_id: int, cloud_path: str, cloud_size_m: float, light_direction):
        l = (C.c_float * 3)(*[float(x) for x in light_direction])
        self._check(self.lib.ds_dataset_append_scene_setup(self.h, scene_id, cloud_path.encode(), float(cloud_size_m), l))

    def append_scatter_samples(self, start_id: int, positions, directions):
        p = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, dtype=np.float32).reshape(-1, 3)
        assert p.shape == d.shape
        self._check(self.lib.ds_dataset_append_scatter_samples(self.h, start_id, p.shape[0], p.ctypes.data, d.ctypes.data))

    def append_descriptors(self, start_id: int, descriptors):
        a = np.ascontiguousarray(descriptors, dtype=np.uint8)
        a = a.reshape(a.shape[0], -1)
        self._check(self.lib.ds_dataset_append_descriptors(self.h, start_id, a.shape[0], a.ctypes.data, a.shape[1]))

    def append_results(self, start_id: int, light_intensity, is_converged):
        r = np.ascontiguousarray(light_intensity, dtype=np.float32).reshape(-1)
        c = np.ascontiguousarray(is_converged, dtype=np.uint8).reshape(-1)
        assert r.shape == c.shape
        self._check(self.lib.ds_dataset_append_results(self.h, start_id, r.shape[0], r.ctypes.data, c.ctypes.data))
