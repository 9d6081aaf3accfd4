import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run with -m gpu on the B200 box)")


def _have_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip silently
    if HAVE_GPU:
        return
    if "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or ""):
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_library():
    import deepestscatter_b200 as ds

    if not ds.LIB_PATH.exists() or os.environ.get("DS_REBUILD"):
        ds.build_library()
    return ds


SCENE_SMALL = dict(n=64, kind=0, seed=1234, cloud_size_m=7000.0, light_dir=(-0.586, -0.766, -0.271))


@pytest.fixture(scope="session")
def oracle_small():
    """Oracle with the 64^3 synthetic cumulus, sun 'Front' (Tasks.cpp:56), baked."""
    import oracle_lib as ol

    o = ol.Oracle()
    o.volume_synth(SCENE_SMALL["n"], SCENE_SMALL["kind"], SCENE_SMALL["seed"])
    o.scene_set(SCENE_SMALL["cloud_size_m"], SCENE_SMALL["light_dir"])
    o.bake()
    return o


@pytest.fixture(scope="session")
def gpu_small(built_library):
    """Product context with the same scene (exact arithmetic by default for parity tests)."""
    ds = built_library
    ctx = ds.Context(0)
    ctx.set_option("precision", ds.PRECISION_EXACT)
    ctx.volume_synth(SCENE_SMALL["n"], SCENE_SMALL["kind"], SCENE_SMALL["seed"])
    ctx.scene_set(SCENE_SMALL["cloud_size_m"], SCENE_SMALL["light_dir"])
    ctx.bake()
    yield ctx
    ctx.close()
