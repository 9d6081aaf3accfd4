"""The .vdb front end of the cloud importer (deepestscatter_b200/host/VdbReader.hpp through ds_cloud_read_vdb): OpenVDB container
-> what Resources::loadVolumeBuffer (DG/Util/Resources.cpp:80-141) takes from the grid (maximum over the active values, active
bounding box + 1, accessor value at every voxel).

PARITY: unpinned against OpenVDB itself (neither the library nor a .vdb file exists here).  The files come from tests/vdb_writer.py,
a writer of the same published format; what these tests pin is that the reader decodes every storage variant of that format to
the same grid, and that the reference's own loadVolumeBuffer (oracle/_ref, compiled from /root/reference) turns the decoded grid
into the same u8 volume as the product's importer path.
"""
import numpy as np
import pytest

import oracle_lib as ol
import ref_lib as rl
import vdb_writer as vw


def blob(seed=5, shape=(21, 30, 26), origin=(-13, 40, 5)):
    rng = np.random.default_rng(seed)
    dense = np.zeros(shape, np.float32)
    dense[2:-3, 4:-2, 3:-4] = rng.uniform(0.05, 3.0, (shape[0] - 5, shape[1] - 6, shape[2] - 7)).astype(np.float32)
    dense[rng.random(shape) < 0.3] = 0.0  # holes: inactive voxels inside the cloud
    dense[2, 4, 3] = dense[-4, -3, -5] = 2.5
    return dense, origin


VARIANTS = {
    "plain": dict(),
    "zip": dict(compression=vw.COMPRESS_ZIP),
    "active_mask": dict(compression=vw.COMPRESS_ACTIVE_MASK),
    "zip_active_mask": dict(compression=vw.COMPRESS_ZIP | vw.COMPRESS_ACTIVE_MASK),
    "blosc_lz4_active_mask": dict(compression=vw.COMPRESS_BLOSC | vw.COMPRESS_ACTIVE_MASK, blosc_mode="lz4"),
    "blosc_stored": dict(compression=vw.COMPRESS_BLOSC, blosc_mode="stored"),
    "blosc_raw": dict(compression=vw.COMPRESS_BLOSC | vw.COMPRESS_ACTIVE_MASK, blosc_mode="raw"),
    "format_222": dict(compression=vw.COMPRESS_ZIP | vw.COMPRESS_ACTIVE_MASK, version=222),
}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_every_storage_variant_decodes_to_the_same_grid(built_library, tmp_path, variant):
    ds = built_library
    dense, origin = blob()
    w = vw.VdbWriter(**VARIANTS[variant])
    w.from_dense(dense, origin)
    path = tmp_path / f"{variant}.vdb"
    path.write_bytes(w.tobytes())
    got, mx = ds.cloud_read_vdb(path)
    want, lo, want_mx = w.dense_reference()
    assert got.shape == want.shape and np.array_equal(got, want) and mx == want_mx == float(dense.max())
    # the active box of the blob + 1: the same crop the dense importer makes
    cropped, cmx = ds.cloud_crop_active(dense)
    assert np.array_equal(cropped, got) and cmx == mx
    assert tuple(lo) == (origin[0] + 3 - 1, origin[1] + 4 - 1, origin[2] + 2 - 1)


def test_half_float_grids(built_library, tmp_path):
    ds = built_library
    dense, origin = blob(seed=9)
    dense = dense.astype(np.float16).astype(np.float32)  # values a half can hold
    for comp in (0, vw.COMPRESS_ZIP | vw.COMPRESS_ACTIVE_MASK):
        w = vw.VdbWriter(half=True, compression=comp)
        w.from_dense(dense, origin)
        path = tmp_path / f"half{comp}.vdb"
        path.write_bytes(w.tobytes())
        got, mx = ds.cloud_read_vdb(path)
        want, _, want_mx = w.dense_reference()
        assert np.array_equal(got, want) and mx == want_mx


def test_inactive_value_encodings_tiles_and_background(built_library, tmp_path):
    """All seven node-mask-compression encodings of inactive values, tiles at the three levels (active ones stretch the bounding box and
    count for the maximum, inactive ones only colour the accessor), a non-zero background."""
    ds = built_library
    rng = np.random.default_rng(2)
    bg = 0.25
    w = vw.VdbWriter(background=bg, compression=vw.COMPRESS_ACTIVE_MASK | vw.COMPRESS_ZIP)
    cases = {
        vw.NO_MASK_OR_INACTIVE_VALS: lambda m: np.where(m, 1.0, bg),
        vw.NO_MASK_AND_MINUS_BG: lambda m: np.where(m, 1.0, -bg),
        vw.NO_MASK_AND_ONE_INACTIVE_VAL: lambda m: np.where(m, 1.0, 0.75),
        vw.MASK_AND_NO_INACTIVE_VALS: lambda m: np.where(m, 1.0, np.where(rng.random(512) < 0.5, bg, -bg)),
        vw.MASK_AND_ONE_INACTIVE_VAL: lambda m: np.where(m, 1.0, np.where(rng.random(512) < 0.5, bg, 0.5)),
        vw.MASK_AND_TWO_INACTIVE_VALS: lambda m: np.where(m, 1.0, np.where(rng.random(512) < 0.5, 0.5, 0.625)),
        vw.NO_MASK_AND_ALL_VALS: lambda m: np.where(m, 1.0, rng.random(512)),
    }
    for i, (md, make) in enumerate(cases.items()):
        mask = rng.random(512) < 0.4
        vals = make(mask).astype(np.float32)
        vals[mask] = rng.uniform(0.5, 2.0, int(mask.sum())).astype(np.float32)
        w.set_leaf((8 * i, 16, -8), vals, mask, None)  # the writer picks the encoding like OpenVDB's MaskCompress; check it did
        assert w._values(vals, mask)[0] == md, (md, w._values(vals, mask)[0])
    w.set_tile(1, (64, 16, -8), 3.5, True)     # active 8^3 tile: stretches the box, raises the maximum
    w.set_tile(1, (72, 16, -8), 9.0, False)    # inactive tile: visible to the accessor only
    w.set_tile(2, (0, 128, 0), 0.5, False)     # inactive 128^3 tile next to the data
    path = tmp_path / "encodings.vdb"
    path.write_bytes(w.tobytes())
    got, mx = ds.cloud_read_vdb(path)
    want, lo, want_mx = w.dense_reference()
    assert mx == want_mx == 3.5
    assert got.shape == want.shape and np.array_equal(got, want)
    assert (got == 9.0).any() and (got == bg).any()


def test_reference_load_volume_buffer_on_the_decoded_grid(built_library, tmp_path):
    """The decoded grid through the REFERENCE'S OWN Resources::loadVolumeBuffer (oracle/_ref) and through the oracle's quantiser
    gives the same u8 volume and mip chain: the .vdb path ends where the pinned dense path begins."""
    if not rl.available():
        pytest.skip("oracle/_ref is not built")
    ds = built_library
    dense, origin = blob(seed=11, shape=(18, 22, 20))
    w = vw.VdbWriter(compression=vw.COMPRESS_ZIP | vw.COMPRESS_ACTIVE_MASK)
    w.from_dense(dense, origin)
    path = tmp_path / "cloud.vdb"
    path.write_bytes(w.tobytes())
    got, mx = ds.cloud_read_vdb(path)
    r = rl.Reference()
    r.volume_import(got)  # its active box + 1 is the grid itself minus the padding the reader already added ...
    q = np.empty(got.size, np.uint8)
    ol.lib().orc_quantize_float_grid(np.ascontiguousarray(got.reshape(-1)), got.size, float(mx), q)
    assert r.level_dims(0) == (got.shape[2], got.shape[1], got.shape[0])  # ... so the size does not change
    assert np.array_equal(r.level(0), q.reshape(got.shape))


def test_errors_are_loud(built_library, tmp_path):
    ds = built_library
    bad = tmp_path / "bad.vdb"
    bad.write_bytes(b"not a vdb file at all, just bytes")
    with pytest.raises(ds.DsError) as e:
        ds.cloud_read_vdb(bad)
    assert "magic" in str(e.value) or "truncated" in str(e.value)
    w = vw.VdbWriter()
    w.from_dense(blob()[0])
    data = bytearray(w.tobytes())
    trunc = tmp_path / "trunc.vdb"
    trunc.write_bytes(bytes(data[: len(data) // 2]))
    with pytest.raises(ds.DsError):
        ds.cloud_read_vdb(trunc)
    with pytest.raises(ds.DsError):
        ds.cloud_read_vdb(tmp_path / "missing.vdb")


@pytest.mark.gpu
def test_cloud_load_vdb_on_the_device(built_library, tmp_path):
    ds = built_library
    dense, origin = blob(seed=4)
    w = vw.VdbWriter(compression=vw.COMPRESS_BLOSC | vw.COMPRESS_ACTIVE_MASK)
    w.from_dense(dense, origin)
    path = tmp_path / "cloud.vdb"
    path.write_bytes(w.tobytes())
    got, mx = ds.cloud_read_vdb(path)
    want_u8 = (got.astype(np.float64) / np.float64(mx) * 255).astype(np.uint8)  # Resources.cpp:137
    with ds.Context(0) as ctx:
        size = ctx.cloud_load(str(path))
        assert size == (got.shape[2], got.shape[1], got.shape[0])
        assert np.array_equal(ctx.level(0), want_u8)
