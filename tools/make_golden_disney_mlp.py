#!/usr/bin/env python3
"""Generate tests/golden/disney_mlp.json from the REFERENCE's own model code.

Imports /root/reference/DeepestScatter_Train/Disney/DisneyModel.py (torch, CPU, float32), loads the deterministic
synthetic weights of deepestscatter_b200.disney_model.synthetic_weights into it through load_state_dict (so the
flat layout of the C ABI is checked against torch's own state_dict order) and runs forward() on seeded inputs.
Runs in the build container only; the JSON it writes is the committed fixture that pins the oracle restatement
(oracle/ds_oracle_mlp.cpp) and, through it, the CUDA kernels (the GPU box has no /root/reference).

The fixture stores seeds, the outputs and, for one row, the activations after every block, not the 5 MB of weights:
tests regenerate weights and inputs from the same seeds (numpy RandomState: frozen stream).
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/DeepestScatter_Train")
sys.path.insert(0, str(REF / "Disney"))
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from DisneyModel import DisneyModel  # noqa: E402  (the reference's)

from deepestscatter_b200 import disney_model as dm  # noqa: E402

OUT = ROOT / "tests" / "golden" / "disney_mlp.json"
WEIGHT_SEED, INPUT_SEED, N = 566, 7, 160


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    model = DisneyModel()
    # torch's own state_dict order must be the order the C ABI documents
    assert [k for k in model.state_dict().keys()] == [n for n, _ in dm.tensor_shapes()]
    w = dm.synthetic_weights(WEIGHT_SEED)
    sd = {k: torch.from_numpy(v.copy()) for k, v in dm.unflatten(w).items()}
    model.load_state_dict(sd, strict=True)
    model.eval()
    assert np.array_equal(dm.flatten_state_dict(model.state_dict()), w)
    x = dm.synthetic_inputs(N, INPUT_SEED)
    with torch.no_grad():
        y = model(torch.from_numpy(x)).numpy().reshape(-1)
        # activations after each block for row 0 (DisneyModel.__blocksForward), float64 copy of the same weights
        m64 = DisneyModel().double()
        m64.load_state_dict({k: v.double() for k, v in sd.items()})
        m64.eval()
        y64, hidden = [], None
        for lo in range(0, N, 32):
            xb = torch.from_numpy(x[lo : lo + 32]).double()
            out = torch.zeros((xb.shape[0], dm.BLOCK_DIM), dtype=torch.float64)
            for i, block in enumerate(m64.blocks):
                out = block(out, xb.narrow(1, i, 1).squeeze(1))
            if lo == 0:
                hidden = out[:4].numpy().copy()  # the 200 activations entering fullyConnected, rows 0..3
            y64.append(m64.fullyConnected(out).numpy().reshape(-1))
        y64 = np.concatenate(y64)
    rec = {
        "generator": "tools/make_golden_disney_mlp.py (reference DisneyModel.py, torch %s, CPU)" % torch.__version__,
        "weight_seed": WEIGHT_SEED,
        "input_seed": INPUT_SEED,
        "n": N,
        "weights_sha_head": [float(v) for v in w[:4]],
        "inputs_head": [float(v) for v in x.reshape(-1)[:4]],
        "output_f32": [float(v) for v in y],
        "output_f64": [float(v) for v in y64],
        "hidden_after_blocks_f64_rows0to3": [[float(v) for v in row] for row in hidden],
    }
    OUT.write_text(json.dumps(rec, indent=0))
    print("wrote", OUT, "max |f32 - f64| rel", float(np.max(np.abs(y - y64) / (np.abs(y64) + 1e-6))), "mean |y|", float(np.abs(y64).mean()))


if __name__ == "__main__":
    main()
