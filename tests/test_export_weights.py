"""tools/export_disney_weights.py: TorchScript exports and pickled state_dicts of the reference's network -> the flat float32 model file,
and once more the oracle against torch itself (a network with the reference's parameter names and data flow, random weights)."""
import sys
from pathlib import Path

import numpy as np
import pytest

torch = pytest.importorskip("torch")

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))

import oracle_lib as ol  # noqa: E402
from deepestscatter_b200 import disney_model as dm  # noqa: E402
from export_disney_weights import export  # noqa: E402


class Block(torch.nn.Module):
    """Same parameter names and data flow as the reference's residual block (f1z, f1o, f2)."""

    def __init__(self):
        super().__init__()
        self.f1z = torch.nn.Linear(dm.LAYER_DIM, dm.BLOCK_DIM)
        self.f1o = torch.nn.Linear(dm.BLOCK_DIM, dm.BLOCK_DIM)
        self.f2 = torch.nn.Linear(dm.BLOCK_DIM, dm.BLOCK_DIM)

    def forward(self, o, z):
        h = torch.relu(self.f1o(o) + self.f1z(z))
        return torch.relu(self.f2(h) + o)


class Model(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.blocks = torch.nn.ModuleList([Block() for _ in range(dm.BLOCK_COUNT)])
        self.fullyConnected = torch.nn.Sequential(
            torch.nn.Linear(dm.BLOCK_DIM, dm.BLOCK_DIM), torch.nn.ReLU(), torch.nn.Linear(dm.BLOCK_DIM, dm.BLOCK_DIM), torch.nn.ReLU(),
            torch.nn.Linear(dm.BLOCK_DIM, 1), torch.nn.LeakyReLU())

    def forward(self, x):
        o = torch.zeros((x.shape[0], dm.BLOCK_DIM), dtype=x.dtype)
        for i, block in enumerate(self.blocks):
            o = block(o, x[:, i])
        return self.fullyConnected(o)


@pytest.fixture(scope="module")
def model():
    torch.manual_seed(11)
    m = Model().eval()
    assert [k for k in m.state_dict()] == [n for n, _ in dm.tensor_shapes()]
    return m


def test_export_from_torchscript_state_dict_and_module(model, tmp_path):
    want = dm.flatten_state_dict(model.state_dict())
    x = torch.from_numpy(dm.synthetic_inputs(3, 5))
    traced = torch.jit.trace(model, x)
    torch.jit.save(traced, str(tmp_path / "DisneyModel.pt"))  # what DisneyRenderer::init loads
    torch.save(model.state_dict(), str(tmp_path / "sd.pth"))
    torch.save({"state_dict": {"module." + k: v for k, v in model.state_dict().items()}, "epoch": 3}, str(tmp_path / "wrapped.pth"))
    for name in ("DisneyModel.pt", "sd.pth", "wrapped.pth"):
        got = export(str(tmp_path / name), str(tmp_path / (name + ".f32")))
        assert np.array_equal(got, want)
        assert np.array_equal(np.fromfile(tmp_path / (name + ".f32"), np.float32), want)
    with pytest.raises((ValueError, KeyError)):
        torch.save({"weights": 1}, str(tmp_path / "bad.pth"))
        export(str(tmp_path / "bad.pth"), str(tmp_path / "bad.f32"))


def test_oracle_matches_torch_on_random_weights(model):
    """Independent of the committed fixture: torch float64 forward of a freshly initialised network against the oracle restatement."""
    x = dm.synthetic_inputs(40, 9)
    with torch.no_grad():
        ref = model.double()(torch.from_numpy(x).double()).numpy().reshape(-1)
    model.float()
    got = ol.disney_forward(dm.flatten_state_dict(model.state_dict()), x)
    assert np.max(np.abs(got - ref) / (np.abs(ref) + 1e-6)) <= 1e-6


REF_DISNEY = Path("/root/reference/DeepestScatter_Train/Disney")


@pytest.mark.skipif(not (REF_DISNEY / "DisneyModel.py").is_file(), reason="the reference tree is not mounted here")
def test_the_references_exported_model_converts_and_evaluates(tmp_path, monkeypatch):
    """Trainer.exportModel (Common/Trainer.py:65-67) on the reference's own DisneyModel: torch.jit.trace(model, z).save('DisneyModel.pt') -- the
    file DisneyRenderer::init loads -- converted by the tool and evaluated by the oracle against the traced module itself."""
    import importlib

    monkeypatch.syspath_prepend(str(REF_DISNEY))
    sys.modules.pop("DisneyModel", None)
    sys.modules.pop("DisneyBlock", None)
    try:
        DisneyModel = importlib.import_module("DisneyModel").DisneyModel
        torch.manual_seed(5)
        ref_model = DisneyModel().eval()
        z = torch.from_numpy(dm.synthetic_inputs(6, 13))
        traced = torch.jit.trace(ref_model, z)
        traced.save(str(tmp_path / "DisneyModel.pt"))
        flat = export(str(tmp_path / "DisneyModel.pt"), str(tmp_path / "DisneyModel.f32"))
        assert flat.size == dm.WEIGHT_COUNT
        with torch.no_grad():
            want = traced(z).numpy().reshape(-1)
        got = ol.disney_forward(flat, z.numpy())
        assert np.max(np.abs(got - want) / (np.abs(want) + 1e-6)) <= 2e-6
    finally:
        sys.modules.pop("DisneyModel", None)
        sys.modules.pop("DisneyBlock", None)
