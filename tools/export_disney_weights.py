#!/usr/bin/env python3
"""Convert a checkpoint of the reference's radiance-predicting network into the flat float32 file the C ABI loads.

    python tools/export_disney_weights.py <checkpoint> <out.f32>

<checkpoint> is any of what the reference's training side produces for DisneyModel (DeepestScatter_Train/Disney):
  * a TorchScript export (torch.jit.save / trace) -- the `DisneyModel.pt` that DisneyRenderer::init loads (DisneyRenderer.cpp:19-22);
  * a pickled state_dict (torch.save(model.state_dict(), ...)), bare or under the key "state_dict" / "model";
  * a pickled nn.Module.
The output is DisneyModel().state_dict() flattened in its own order (deepestscatter_b200/disney_model.py), 1 338 601 float32 values:
what ds_disney_model_load and `datagen render --renderer disney --model <out.f32>` expect.
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

from deepestscatter_b200 import disney_model as dm  # noqa: E402


def load_state_dict(path: str) -> dict:
    import torch

    try:
        module = torch.jit.load(path, map_location="cpu")
        return module.state_dict()
    except RuntimeError:
        pass  # not a TorchScript archive
    obj = torch.load(path, map_location="cpu", weights_only=False)
    if hasattr(obj, "state_dict"):
        return obj.state_dict()
    if isinstance(obj, dict):
        for key in ("state_dict", "model", "model_state_dict"):
            if key in obj and isinstance(obj[key], dict):
                obj = obj[key]
                break
        # DataParallel / Lightning prefixes
        for prefix in ("module.", "model."):
            if obj and all(k.startswith(prefix) for k in obj):
                obj = {k[len(prefix):]: v for k, v in obj.items()}
        return obj
    raise ValueError(f"{path}: neither a TorchScript module, an nn.Module nor a state_dict")


def export(checkpoint: str, out: str) -> np.ndarray:
    flat = dm.flatten_state_dict(load_state_dict(checkpoint))
    if not np.isfinite(flat).all():
        raise ValueError("the checkpoint holds non-finite weights")
    flat.tofile(out)
    return flat


def main():
    if len(sys.argv) != 3:
        print(__doc__)
        return 2
    flat = export(sys.argv[1], sys.argv[2])
    print(f"wrote {sys.argv[2]}: {flat.size} float32 values, |w| max {np.abs(flat).max():.4g}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
