"""The radiance-predicting network of the neural renderer (second half).

Chain of evidence: the reference's own DisneyModel.py (torch) -> tests/golden/disney_mlp.json (tools/make_golden_disney_mlp.py)
-> oracle/ds_oracle_mlp.cpp (CPU tests below) -> the two CUDA kernels through the C ABI (GPU tests below).

Tolerances (written here as the task asks):
  oracle vs reference float64 outputs      1e-6 relative (double accumulation against torch float64)
  oracle vs reference float32 outputs      2e-6 relative (torch's own float32 rounding)
  EXACT flavour (fp32 FMA kernel) vs oracle   1e-5 relative
  FAST flavour (tcgen05 kind::f16 on IEEE half operands by default, kind::tf32 with option mlp_fp16 = 0) vs oracle 5e-3 relative: operands carry 10 mantissa bits (2^-11 relative rounding) through 23
      layers with fp32 accumulation and an fp32 residual path
"""
import json
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as ol
from conftest import SCENE_SMALL
from deepestscatter_b200 import disney_model as dm

GOLDEN = json.loads((Path(__file__).resolve().parent / "golden" / "disney_mlp.json").read_text())


@pytest.fixture(scope="module")
def weights():
    w = dm.synthetic_weights(GOLDEN["weight_seed"])
    assert [float(v) for v in w[:4]] == GOLDEN["weights_sha_head"]  # numpy's legacy stream is frozen
    return w


@pytest.fixture(scope="module")
def inputs():
    x = dm.synthetic_inputs(GOLDEN["n"], GOLDEN["input_seed"])
    assert [float(v) for v in x.reshape(-1)[:4]] == GOLDEN["inputs_head"]
    return x


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / (np.abs(np.asarray(b, np.float64)) + 1e-6)))


def test_weight_layout_matches_state_dict_order():
    assert dm.WEIGHT_COUNT == 1338601 == ol.lib().orc_disney_weight_count()
    names = [n for n, _ in dm.tensor_shapes()]
    assert names[:6] == ["blocks.0.f1z.weight", "blocks.0.f1z.bias", "blocks.0.f1o.weight", "blocks.0.f1o.bias", "blocks.0.f2.weight", "blocks.0.f2.bias"]
    assert names[-6:] == ["fullyConnected.0.weight", "fullyConnected.0.bias", "fullyConnected.2.weight", "fullyConnected.2.bias",
                          "fullyConnected.4.weight", "fullyConnected.4.bias"]
    w = dm.synthetic_weights(3)
    assert np.array_equal(dm.flatten_state_dict(dm.unflatten(w)), w)


def test_oracle_matches_reference_model_outputs(weights, inputs):
    """The oracle restatement against outputs of the reference's own DisneyModel.py (committed golden vectors)."""
    y, hidden = ol.disney_forward(weights, inputs, want_hidden=True)
    assert rel(y, GOLDEN["output_f64"]) <= 1e-6
    assert rel(y, GOLDEN["output_f32"]) <= 2e-6
    h_ref = np.array(GOLDEN["hidden_after_blocks_f64_rows0to3"])
    assert np.abs(hidden[:4] - h_ref).max() <= 1e-6 and (h_ref > 0).mean() > 0.5


def test_oracle_block_zero_ignores_previous_output(weights):
    """o = zeros entering block 0 (DisneyModel.py:34): f1o contributes its bias only; rows are independent."""
    x = dm.synthetic_inputs(5, 11)
    y = ol.disney_forward(weights, x)
    y_rev = ol.disney_forward(weights, x[::-1].copy())
    assert np.array_equal(y, y_rev[::-1])


# ---------------------------------------------------------------- GPU parity, through the C ABI


@pytest.fixture(scope="module")
def gpu_ctx(built_library, weights):
    ctx = built_library.Context(0)
    ctx.disney_model_load(weights)
    yield ctx
    ctx.close()


@pytest.mark.gpu
def test_forward_requires_a_model_and_checks_the_blob(built_library, weights):
    ds = built_library
    with ds.Context(0) as ctx:
        with pytest.raises(ds.DsError) as e:
            ctx.disney_model_forward(dm.synthetic_inputs(1))
        assert e.value.code == -3 and "no model loaded" in str(e.value)
        with pytest.raises(ds.DsError) as e:
            ctx.disney_model_load(weights[:-1])
        assert e.value.code == -1
        bad = weights.copy()
        bad[12345] = np.nan
        with pytest.raises(ds.DsError):
            ctx.disney_model_load(bad)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 63, 160, 1000])
def test_exact_flavour_matches_oracle(gpu_ctx, built_library, weights, inputs, n):
    ds = built_library
    x = inputs[:n] if n <= len(inputs) else dm.synthetic_inputs(n, 21)
    gpu_ctx.set_option("precision", ds.PRECISION_EXACT)
    got = gpu_ctx.disney_model_forward(x)
    ref = ol.disney_forward(weights, x)
    assert rel(got, ref) <= 1e-5
    if n == 160:
        assert rel(got, GOLDEN["output_f64"]) <= 1e-5  # and straight against the reference's outputs


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 127, 160, 1000])
def test_tensor_core_flavour_matches_oracle(gpu_ctx, built_library, weights, inputs, n):
    ds = built_library
    x = inputs[:n] if n <= len(inputs) else dm.synthetic_inputs(n, 21)
    gpu_ctx.set_option("precision", ds.PRECISION_FAST)
    got = gpu_ctx.disney_model_forward(x)
    ref = ol.disney_forward(weights, x)
    assert np.isfinite(got).all()
    assert rel(got, ref) <= 5e-3
    assert np.array_equal(got, gpu_ctx.disney_model_forward(x))  # deterministic
    if n == 160:
        assert rel(got, GOLDEN["output_f64"]) <= 5e-3


@pytest.mark.gpu
def test_tensor_core_flavour_full_rectangle_and_zero_input(gpu_ctx, built_library, weights):
    """128 x 128 rows (one rectangle of DisneyRenderer, 128 CTAs) and the all-zero descriptor (a pixel outside the cloud)."""
    ds = built_library
    x = dm.synthetic_inputs(128 * 128, 33)
    x[5] = 0.0
    gpu_ctx.set_option("precision", ds.PRECISION_FAST)
    got = gpu_ctx.disney_model_forward(x)
    gpu_ctx.set_option("precision", ds.PRECISION_EXACT)
    exact = gpu_ctx.disney_model_forward(x)
    assert rel(got, exact) <= 5e-3
    sel = np.r_[0:64, 5000:5064, 16320:16384]
    assert rel(exact[sel], ol.disney_forward(weights, x[sel])) <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_render_disney_composes_the_reference_frame(built_library, oracle_small, weights, precision):
    """DisneyRenderer::render: network-input launch + model on the scattering pixels + copyToFrameResult per rectangle.  The frame is
    wider than one rectangle, so the second rectangle is clipped and uses stream + 1."""
    ds = built_library
    w, h = 160, 40
    cam = ds.camera_look_at(aspect=w / h)
    cam_arr = ds.camera_array(cam)
    want = np.zeros((h, w, 4), np.float32)
    for k, (x0, rw) in enumerate([(0, 128), (128, 32)]):
        inp, info = oracle_small.network_input(cam_arr, w, h, (x0, 0, rw, h), stream=9 + k)
        has = info[..., 4] > 0
        pred = np.zeros((h, rw), np.float32)
        if has.any():
            pred[has] = ol.disney_forward(weights, inp[has])
        wgt = (1 - info[..., 3:4])
        rect = np.concatenate([(pred[..., None] + info[..., :3]) * wgt, pred[..., None] * wgt], axis=-1)
        want[:, x0:x0 + rw][has] = rect[has]
    with ds.Context(0) as ctx:
        ctx.set_option("precision", ds.PRECISION_EXACT if precision == "exact" else ds.PRECISION_FAST)
        ctx.volume_synth(SCENE_SMALL["n"], SCENE_SMALL["kind"], SCENE_SMALL["seed"])
        ctx.scene_set(SCENE_SMALL["cloud_size_m"], SCENE_SMALL["light_dir"])
        ctx.bake()
        ctx.disney_model_load(weights)
        got = ctx.render_disney(cam, w, h, stream=9)
    lit = want[..., 3] != 0
    assert 0.05 < lit.mean() < 1.0
    if precision == "exact":
        assert np.array_equal(got[..., 3] != 0, lit)
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    else:
        # FAST: hardware-filtered taps move a few silhouette pixels; compare where both scattered
        both = lit & (got[..., 3] != 0)
        assert both.sum() > 0.97 * lit.sum()
        assert abs(float(got[..., 3][both].mean()) / float(want[..., 3][both].mean()) - 1) < 0.02


# ---------------------------------------------------------------- the tensor-core kernel's program, emulated on the CPU

CHUNK_DTYPE = np.dtype([("wOffset", "<u4"), ("wBytes", "<u4"), ("k8", "<u2"), ("aK", "<u2"), ("src", "u1"), ("layer", "u1"), ("dst", "u1"),
                        ("flags", "u1"), ("epilogue", "u1"), ("gemm", "u1"), ("pad", "u1", 2)])


def packed_program(built_library, weights, bf16=False):
    import ctypes as C

    lib = built_library.load()
    nb, nc = C.c_size_t(), C.c_size_t()
    w = np.ascontiguousarray(weights, np.float32)
    assert lib.ds_disney_model_pack(w.ctypes.data, w.size, int(bf16), None, 0, None, 0, C.byref(nb), C.byref(nc)) == 0
    stream = np.zeros(nb.value, np.uint8)
    chunks = np.zeros(nc.value, CHUNK_DTYPE)
    assert lib.ds_disney_model_pack(w.ctypes.data, w.size, int(bf16), stream.ctypes.data, stream.size, chunks.ctypes.data, chunks.nbytes, C.byref(nb), C.byref(nc)) == 0
    return stream, chunks


@pytest.mark.parametrize("bf16", [0, 1, 2])
def test_tensor_core_program_emulated_on_cpu(built_library, weights, inputs, bf16):
    """Runs the chunk program of k_disney_mlp_tc in numpy (float64 accumulators): decodes every weight chunk from the canonical
    core-matrix layout (4 tf32, or 8 bfloat16 (1) / IEEE half (2) per 16-byte K group), follows the overwrite / accumulate / residual-in-accumulator / epilogue
    flags, and must land on the oracle within the rounding of the weights (activations are not rounded here)."""
    assert CHUNK_DTYPE.itemsize == 20
    G = 8 if bf16 else 4  # K values per 16-byte K group
    zpad = 240 if bf16 else 232  # descriptor layer padded to whole MMA steps
    stream, chunks = packed_program(built_library, weights, bf16)
    assert len(chunks) <= 232 and chunks["wOffset"][0] == 0 and np.all(chunks["wBytes"] == chunks["k8"].astype(np.uint32) * 2 * 26 * 128)
    assert np.all(np.diff(chunks["wOffset"].astype(np.int64)) == chunks["wBytes"][:-1])  # consumption order, no gaps
    assert chunks["wOffset"][-1] + chunks["wBytes"][-1] == stream.size and np.all(chunks["wOffset"] % 16 == 0)
    n = 24
    x = np.zeros((n, 10, zpad))
    x[:, :, :226] = inputs[:n]
    x[:, :, 226:228] = 1.0  # the kernel stages z[226] = z[227] = 1: the bias columns of a block's first GEMM
    if bf16 == 2:
        words = stream.view(np.float16).astype(np.float32)
    elif bf16:
        words = (stream.view(np.uint16).astype(np.uint32) << 16).view(np.float32)
    else:
        words = stream.view(np.float32)
    E = 2 if bf16 else 4
    act = np.zeros((n, 208))
    act[:, 200:202] = 1.0  # the constant-one columns of the activation buffer
    D = [np.zeros((n, 208)), np.zeros((n, 208))]
    unflat = dm.unflatten(weights)
    out = None
    nn, kk = np.meshgrid(np.arange(208), np.arange(64), indexing="ij")
    for ch in chunks:
        kc = int(ch["k8"]) * 2 * G
        off = (kk[:, :kc] // G) * (26 * 128) + (nn[:, :kc] // 8) * 128 + (nn[:, :kc] % 8) * 16 + (kk[:, :kc] % G) * E
        W = words[(int(ch["wOffset"]) + off) // E].astype(np.float64)  # [208][kc]
        assert np.all(W[200:] == 0)
        if ch["src"] == 1:
            k0 = int(ch["aK"])
            assert k0 + kc <= zpad
            A = x[:, int(ch["layer"]), k0 : k0 + kc]
        else:
            k0 = int(ch["aK"]) * G
            assert k0 + kc <= 208
            A = act[:, k0 : k0 + kc]
        d = int(ch["dst"])
        D[d] = A @ W.T if ch["flags"] & 1 else D[d] + A @ W.T
        if ch["flags"] & 2:
            v = np.maximum(D[d][:, :200], 0.0)  # the bias is already in the accumulator
            if ch["epilogue"] == 3:
                y = v @ unflat["fullyConnected.4.weight"][0].astype(np.float64) + float(unflat["fullyConnected.4.bias"][0])
                out = np.where(y > 0, y, 0.01 * y)
            else:
                act[:, :200] = v
                if ch["epilogue"] == 2:
                    D[d][:, :200] = v  # the residual of the next block stays in the accumulator
    assert out is not None
    ref = ol.disney_forward(weights, inputs[:n])
    assert rel(out, ref) <= (6e-3 if bf16 == 1 else 2e-3)  # a half has the 10 mantissa bits of tf32


@pytest.mark.gpu
def test_neural_frame_does_not_depend_on_the_compaction_order(built_library, weights):
    """The scattering pixels are compacted with warp-aggregated atomics, so their order -- and with it the tile and the TMEM lane a pixel
    lands in -- varies; the frame must not.  Option compact_reverse processes them in the opposite order."""
    ds = built_library
    cam = ds.camera_look_at(aspect=4.0)
    with ds.Context(0) as ctx:
        ctx.volume_synth(SCENE_SMALL["n"], SCENE_SMALL["kind"], SCENE_SMALL["seed"])
        ctx.scene_set(SCENE_SMALL["cloud_size_m"], SCENE_SMALL["light_dir"])
        ctx.bake()
        ctx.disney_model_load(weights)
        a = ctx.render_disney(cam, 160, 40, stream=3)
        ctx.set_option("compact_reverse", 1)
        b = ctx.render_disney(cam, 160, 40, stream=3)
    assert (a[..., 3] != 0).sum() > 1000 and np.array_equal(a, b)


@pytest.mark.gpu
def test_primary_ray_cache_does_not_change_the_neural_frame_beyond_rounding(built_library, weights):
    """FAST flavour: the network-input pass starts every camera ray at the first occupied cell (k_primary_prepass).  The skipped steps read
    zero density, so transmittance and collisions are the same up to the rounding of the start position; silhouette and mean must agree,
    and a changed camera must invalidate the cache."""
    ds = built_library
    with ds.Context(0) as ctx:
        ctx.volume_synth(SCENE_SMALL["n"], SCENE_SMALL["kind"], SCENE_SMALL["seed"])
        ctx.scene_set(SCENE_SMALL["cloud_size_m"], SCENE_SMALL["light_dir"])
        ctx.bake()
        ctx.disney_model_load(weights)
        cam = ds.camera_look_at(aspect=4.0)
        with_cache = ctx.render_disney(cam, 160, 40, stream=3)
        again = ctx.render_disney(cam, 160, 40, stream=3)
        ctx.set_option("primary_cache", 0)
        without = ctx.render_disney(cam, 160, 40, stream=3)
        ctx.set_option("primary_cache", 1)
        cam2 = ds.camera_look_at(eye=(2.2, 0.5, 0.4), aspect=4.0)
        moved = ctx.render_disney(cam2, 160, 40, stream=3)
        ctx.set_option("primary_cache", 0)
        moved_ref = ctx.render_disney(cam2, 160, 40, stream=3)
    assert np.array_equal(with_cache, again)
    for a, b in ((with_cache, without), (moved, moved_ref)):
        la, lb = a[..., 3] != 0, b[..., 3] != 0
        assert lb.sum() > 1000 and (la != lb).mean() < 0.002
        both = la & lb
        assert abs(float(a[both].mean()) / float(b[both].mean()) - 1) < 0.01
    assert not np.array_equal(moved, with_cache)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 127, 160, 1000])
def test_bf16_operands_match_oracle(gpu_ctx, built_library, weights, inputs, n):
    """Option mlp_bf16: the same kernel with bfloat16 operands (kind::f16; 8 mantissa bits, so 4 x the tf32 rounding: tolerance 1.5e-2),
    fp32 accumulation, residual and biases as before."""
    ds = built_library
    x = inputs[:n] if n <= len(inputs) else dm.synthetic_inputs(n, 21)
    gpu_ctx.set_option("precision", ds.PRECISION_FAST)
    gpu_ctx.set_option("mlp_bf16", 1)
    try:
        got = gpu_ctx.disney_model_forward(x)
        again = gpu_ctx.disney_model_forward(x)
    finally:
        gpu_ctx.set_option("mlp_bf16", 0)
    ref = ol.disney_forward(weights, x)
    assert np.isfinite(got).all() and np.array_equal(got, again)
    assert rel(got, ref) <= 1.5e-2
    assert not np.array_equal(got, gpu_ctx.disney_model_forward(x))  # it is not the tf32 path


def test_half_weights_are_numpy_float16(built_library, weights):
    """The host packer's float -> half rounding (nearest even, subnormals, saturation) is numpy's, element for element, on the model's own
    weights and on values around every boundary of the format."""
    weights = np.array(weights, np.float32)
    unflat0 = dm.unflatten(weights)
    probe = unflat0["fullyConnected.0.weight"]
    specials = np.array([65504.0, 65519.9, 65520.0, 1e9, -7e4, 6.103515625e-05, 6.1e-05, 6.0975552e-05, 5.9604645e-08, 8.9e-08, 8.95e-08, 2.9802322e-08,
                         2.9802326e-08, 1e-09, -1e-07, 0.33325195, 0.33337402, 1.0009766, 1.0004883, 1.0014648, 0.0], np.float32)
    # write the probes into row 0 of fullyConnected.0.weight inside the flat array
    sizes = [(k, v.size) for k, v in unflat0.items()]
    pos = 0
    for k, n_ in sizes:
        if k == "fullyConnected.0.weight":
            break
        pos += n_
    assert np.array_equal(weights[pos : pos + probe.size].reshape(probe.shape), probe)  # unflatten keeps the flat order
    weights[pos : pos + specials.size] = specials
    s16, c16 = packed_program(built_library, weights, 2)
    sbf, cbf = packed_program(built_library, weights, 1)
    assert np.array_equal(c16, cbf) and s16.size == sbf.size  # one chunk table for both 16-bit types
    # the stream holds round(w) and, in the bias columns, round(b) and round(b - round(b)): decode one GEMM's plain weights and compare
    unflat = dm.unflatten(weights)
    W = unflat["fullyConnected.0.weight"]  # [200][200]
    ch = [c for c in c16 if c["gemm"] == 20]
    half = s16.view(np.float16)
    got = np.zeros((200, 208), np.float16)
    for c in ch:
        kc = int(c["k8"]) * 16
        k0 = int(c["aK"]) * 8
        nn, kk = np.meshgrid(np.arange(200), np.arange(kc), indexing="ij")
        off = (kk // 8) * (26 * 128) + (nn // 8) * 128 + (nn % 8) * 16 + (kk % 8) * 2
        got[:, k0 : k0 + kc] = half[(int(c["wOffset"]) + off) // 2]
    with np.errstate(over="ignore"):
        want = W.astype(np.float16)
    want[np.isinf(want)] = np.sign(want[np.isinf(want)]) * np.float16(65504)  # the packer saturates (cvt.rn.satfinite), numpy overflows to inf
    assert np.array_equal(got[:, :200].view(np.uint16), want.view(np.uint16))


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 127, 160, 1000, 16384])
def test_half_operands_match_oracle_like_tf32(gpu_ctx, built_library, weights, inputs, n):
    """Option mlp_fp16: kind::f16 on IEEE half operands -- the MMA rate and operand bytes of bf16, the 10 mantissa bits of tf32: held to the
    tf32 tolerance (5e-3), not to the bf16 one."""
    ds = built_library
    x = inputs[:n] if n <= len(inputs) else dm.synthetic_inputs(n, 21)
    gpu_ctx.set_option("precision", ds.PRECISION_FAST)
    assert gpu_ctx.get_option("mlp_fp16") == 1  # the default
    got = gpu_ctx.disney_model_forward(x)
    again = gpu_ctx.disney_model_forward(x)
    gpu_ctx.set_option("mlp_fp16", 0)
    try:
        tf32 = gpu_ctx.disney_model_forward(x)
    finally:
        gpu_ctx.set_option("mlp_fp16", 1)
    ref = ol.disney_forward(weights, x)
    assert np.isfinite(got).all() and np.array_equal(got, again)
    assert rel(got, ref) <= 5e-3
    assert rel(tf32, ref) <= 5e-3  # option mlp_fp16 = 0: the tf32 kernel stays held to the same bar
    assert rel(got, ref) <= 3 * max(rel(tf32, ref), 5e-4)
    assert not np.array_equal(got, tf32)  # it is not the tf32 path


@pytest.mark.gpu
def test_render_disney_with_half_operands(built_library, weights):
    ds = built_library
    cam = ds.camera_look_at(aspect=4.0)
    with ds.Context(0) as ctx:
        ctx.volume_synth(SCENE_SMALL["n"], SCENE_SMALL["kind"], SCENE_SMALL["seed"])
        ctx.scene_set(SCENE_SMALL["cloud_size_m"], SCENE_SMALL["light_dir"])
        ctx.bake()
        ctx.disney_model_load(weights)
        ctx.set_option("mlp_fp16", 0)
        a = ctx.render_disney(cam, 160, 40, stream=3)  # tf32 operands
        ctx.set_option("mlp_fp16", 1)
        b = ctx.render_disney(cam, 160, 40, stream=3)
        b2 = ctx.render_disney(cam, 160, 40, stream=3)
    lit = a[..., 3] != 0
    assert np.array_equal(lit, b[..., 3] != 0) and np.array_equal(b, b2) and not np.array_equal(a, b)
    assert np.abs(a - b)[lit].max() <= 5e-3 * np.abs(a)[lit].max()


@pytest.mark.gpu
def test_render_disney_with_bf16_operands(built_library, weights):
    ds = built_library
    cam = ds.camera_look_at(aspect=4.0)
    with ds.Context(0) as ctx:
        ctx.volume_synth(SCENE_SMALL["n"], SCENE_SMALL["kind"], SCENE_SMALL["seed"])
        ctx.scene_set(SCENE_SMALL["cloud_size_m"], SCENE_SMALL["light_dir"])
        ctx.bake()
        ctx.disney_model_load(weights)
        a = ctx.render_disney(cam, 160, 40, stream=3)
        ctx.set_option("mlp_bf16", 1)
        b = ctx.render_disney(cam, 160, 40, stream=3)
        b2 = ctx.render_disney(cam, 160, 40, stream=3)
    lit = a[..., 3] != 0
    assert np.array_equal(lit, b[..., 3] != 0) and np.array_equal(b, b2) and not np.array_equal(a, b)
    assert np.abs(a - b)[lit].max() <= 1.5e-2 * np.abs(a)[lit].max()
