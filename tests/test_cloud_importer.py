"""Cloud importer front end (deepestscatter_b200/host/CloudImporter.hpp through ds_cloud_*): the dense-grid restatement of
Resources::loadVolumeBuffer (DG/Util/Resources.cpp:95-141)."""
import numpy as np
import pytest


def reference_crop(dense):
    """Resources.cpp:95-101 on a dense array: max over active voxels, active bbox, expandBy(1), size = max + 1 - min."""
    active = np.argwhere(dense > 0)
    lo, hi = active.min(0), active.max(0)
    lo, hi = lo - 1, hi + 1  # expandBy(1)
    size = hi + 1 - lo
    out = np.zeros(tuple(size), dense.dtype)
    src_lo, src_hi = np.maximum(lo, 0), np.minimum(hi, np.array(dense.shape) - 1)
    dst_lo = src_lo - lo
    sl_src = tuple(slice(a, b + 1) for a, b in zip(src_lo, src_hi))
    sl_dst = tuple(slice(a, a + (b - c + 1)) for a, b, c in zip(dst_lo, src_hi, src_lo))
    out[sl_dst] = dense[sl_src]
    return out, float(dense[dense > 0].max())


@pytest.mark.parametrize("shape,box", [((9, 7, 11), ((2, 5), (1, 4), (3, 9))), ((4, 4, 4), ((0, 3), (0, 3), (0, 3))), ((6, 5, 4), ((5, 5), (0, 0), (3, 3)))])
def test_crop_to_active_box_plus_one(built_library, shape, box):
    ds = built_library
    rng = np.random.default_rng(5)
    dense = np.zeros(shape, np.float32)
    (z0, z1), (y0, y1), (x0, x1) = box
    dense[z0:z1 + 1, y0:y1 + 1, x0:x1 + 1] = rng.uniform(0.0, 4.0, (z1 - z0 + 1, y1 - y0 + 1, x1 - x0 + 1)).astype(np.float32)
    dense[z0, y0, x0] = dense[z1, y1, x1] = 1.0  # the corners are active, whatever the noise drew
    got, mx = ds.cloud_crop_active(dense)
    want, want_mx = reference_crop(dense)
    assert got.shape == want.shape == (z1 - z0 + 3, y1 - y0 + 3, x1 - x0 + 3)
    assert np.array_equal(got, want) and mx == want_mx
    # every face voxel is zero: what lets the kernels skip clamped taps outside the grid
    assert got[0].max() == 0 and got[-1].max() == 0 and got[:, 0].max() == 0 and got[:, -1].max() == 0 and got[:, :, 0].max() == 0 and got[:, :, -1].max() == 0


def test_empty_grid_is_an_error(built_library):
    ds = built_library
    with pytest.raises(ds.DsError) as e:
        ds.cloud_crop_active(np.zeros((3, 3, 3), np.float32))
    assert "no active" in str(e.value)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float64", "uint8"])
def test_cloud_load_npy_matches_the_reference_quantisation(built_library, tmp_path, dtype):
    ds = built_library
    rng = np.random.default_rng(9)
    dense = np.zeros((20, 24, 28), np.float64)
    dense[3:15, 5:20, 2:27] = rng.uniform(0, 2.5, (12, 15, 25))
    dense[3, 5, 2] = dense[14, 19, 26] = 2.5
    if dtype == "uint8":
        dense = np.floor(dense / 2.5 * 255)
    arr = dense.astype(dtype)
    path = tmp_path / f"cloud_{dtype}.npy"
    np.save(path, arr)
    want, mx = reference_crop(arr.astype(np.float32) if dtype != "uint8" else arr)
    if dtype == "uint8":
        want_u8 = want
    else:
        # Resources.cpp:137: narrow_cast<uint8_t>(value / maxDensity * 255), float / double / int -> double arithmetic
        want_u8 = (want.astype(np.float32).astype(np.float64) / np.float64(mx) * 255).astype(np.uint8)
    with ds.Context(0) as ctx:
        size = ctx.cloud_load(str(path))
        assert size == (want.shape[2], want.shape[1], want.shape[0])
        assert np.array_equal(ctx.level(0), want_u8)
        assert ctx.level_count() == 1 + int(np.floor(np.log2(max(size))))
        assert ctx.cloud_load(str(path)) == size  # cached
        assert ctx.cloud_load("synth:32:1:7") == (32, 32, 32)
        with pytest.raises(ds.DsError):
            ctx.cloud_load(str(tmp_path / "missing.npy"))
        with pytest.raises(ds.DsError) as e:
            ctx.cloud_load("cloud.vdb")  # a .vdb path goes to host/VdbReader.hpp (tests/test_vdb_reader.py); this one does not exist
        assert "cannot open" in str(e.value)
        with pytest.raises(ds.DsError) as e:
            ctx.cloud_load("cloud.xyz")
        assert "unsupported cloud file" in str(e.value)
