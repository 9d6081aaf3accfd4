"""Weight blob of the reference's radiance-predicting network (DeepestScatter_Train/Disney/DisneyModel.py).

The C ABI (`ds_disney_model_load`) takes the model as ONE flat float32 array: the tensors of
`DisneyModel().state_dict()` in their own order, each row-major as torch stores them --

    for i in 0..9:   blocks.i.f1z.weight [200][226], blocks.i.f1z.bias [200],
                     blocks.i.f1o.weight [200][200], blocks.i.f1o.bias [200],
                     blocks.i.f2.weight  [200][200], blocks.i.f2.bias  [200]      (DisneyBlock.py:13-15)
    fullyConnected.0.weight [200][200], .0.bias [200], .2.weight [200][200], .2.bias [200],
    fullyConnected.4.weight [1][200],   .4.bias [1]                              (DisneyModel.py:52-59)

1 338 601 floats.  `flatten_state_dict` turns a state_dict (torch tensors or numpy arrays) into that array, so a
checkpoint trained by TR/Disney/TrainDisneyModel.py is exported with

    np.asarray(flatten_state_dict(model.state_dict())).tofile("DisneyModel.f32")

`synthetic_weights` is the deterministic stand-in used by tests and benches (no trained checkpoint ships with the
reference): torch.nn.Linear's default range U(-1/sqrt(fan_in), 1/sqrt(fan_in)) from numpy's frozen legacy stream.
"""
from __future__ import annotations

import numpy as np

BLOCK_DIM = 200  # DisneyModel.BLOCK_DIMENSION
BLOCK_COUNT = 10  # DisneyModel.BLOCK_COUNT
LAYER_DIM = 226  # DisneyModel.DESCRIPTOR_LAYER_WITH_ANGLE_DIMENSION (5*5*9 + 1)


def tensor_shapes():
    """(name, shape) of every state_dict entry, in state_dict order."""
    out = []
    for i in range(BLOCK_COUNT):
        p = f"blocks.{i}."
        out += [
            (p + "f1z.weight", (BLOCK_DIM, LAYER_DIM)),
            (p + "f1z.bias", (BLOCK_DIM,)),
            (p + "f1o.weight", (BLOCK_DIM, BLOCK_DIM)),
            (p + "f1o.bias", (BLOCK_DIM,)),
            (p + "f2.weight", (BLOCK_DIM, BLOCK_DIM)),
            (p + "f2.bias", (BLOCK_DIM,)),
        ]
    out += [
        ("fullyConnected.0.weight", (BLOCK_DIM, BLOCK_DIM)),
        ("fullyConnected.0.bias", (BLOCK_DIM,)),
        ("fullyConnected.2.weight", (BLOCK_DIM, BLOCK_DIM)),
        ("fullyConnected.2.bias", (BLOCK_DIM,)),
        ("fullyConnected.4.weight", (1, BLOCK_DIM)),
        ("fullyConnected.4.bias", (1,)),
    ]
    return out


WEIGHT_COUNT = sum(int(np.prod(s)) for _, s in tensor_shapes())
assert WEIGHT_COUNT == 1338601


def flatten_state_dict(sd) -> np.ndarray:
    parts = []
    for name, shape in tensor_shapes():
        t = sd[name]
        a = t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
        if tuple(a.shape) != shape:
            raise ValueError(f"{name}: shape {tuple(a.shape)} != {shape}")
        parts.append(np.ascontiguousarray(a, dtype=np.float32).ravel())
    return np.concatenate(parts)


def unflatten(weights: np.ndarray) -> dict:
    weights = np.asarray(weights, dtype=np.float32)
    if weights.size != WEIGHT_COUNT:
        raise ValueError(f"expected {WEIGHT_COUNT} floats, got {weights.size}")
    out, off = {}, 0
    for name, shape in tensor_shapes():
        n = int(np.prod(shape))
        out[name] = weights[off : off + n].reshape(shape)
        off += n
    return out


def synthetic_weights(seed: int = 566) -> np.ndarray:
    rs = np.random.RandomState(seed)
    parts = []
    for name, shape in tensor_shapes():
        fan_in = shape[1] if len(shape) == 2 else (LAYER_DIM if "f1z" in name else BLOCK_DIM)
        bound = 1.0 / np.sqrt(fan_in)
        parts.append(rs.uniform(-bound, bound, size=shape).astype(np.float32).ravel())
    return np.concatenate(parts)


def synthetic_inputs(n: int, seed: int = 7) -> np.ndarray:
    """[n][10][226] network inputs shaped like DisneyNetworkInput: 225 densities in [0, 1] (many exactly 0, as around a
    cloud) and the light/view angle in [0, pi] repeated in every layer."""
    rs = np.random.RandomState(seed)
    x = rs.uniform(0.0, 1.0, size=(n, BLOCK_COUNT, LAYER_DIM)).astype(np.float32)
    x[rs.uniform(size=x.shape) < 0.3] = 0.0
    ang = rs.uniform(0.0, np.pi, size=(n, 1)).astype(np.float32)
    x[:, :, LAYER_DIM - 1] = ang
    return x
