"""The C-ABI library loads and exports every symbol include/ds_abi.h declares (no compute calls)."""
import ctypes as C
import re
from pathlib import Path

import pytest

from conftest import HAVE_GPU

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "ds_abi.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ds_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    assert len(syms) >= 40
    for must in ("ds_context_create", "ds_render_subframes", "ds_collect_descriptors", "ds_point_radiance_run", "ds_bake_sun_transmittance"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_library):
    lib = C.CDLL(str(built_library.LIB_PATH))
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in ds_abi.h but not exported: {missing}"


def test_python_binding_covers_every_declared_symbol(built_library):
    from deepestscatter_b200._lib import SIGNATURES

    assert sorted(SIGNATURES) == declared_symbols()


def test_struct_layouts_match_reference(built_library):
    from deepestscatter_b200._lib import DsCamera, DsPointRadianceTask, DsSceneParams

    assert C.sizeof(DsPointRadianceTask) == 40  # CU/PointRadianceTask.h:70-77
    assert C.sizeof(DsCamera) == 48
    assert C.sizeof(DsSceneParams) == 11 * 4


def test_scene_defaults_match_reference(built_library):
    from deepestscatter_b200._lib import DsSceneParams

    p = DsSceneParams()
    built_library.load().ds_scene_params_default(C.byref(p))
    assert p.mean_free_path_m == 10.0  # SceneDescription.h:80
    assert p.sample_step == 1.0 / 512.0  # installers.cpp:86
    assert p.light_intensity == 1e6  # installers.cpp:100
    assert abs(p.minimal_ray_distance - 1e-6) < 1e-12  # CloudMaterial.cpp:23


@pytest.mark.skipif(HAVE_GPU, reason="only meaningful without a GPU")
def test_no_cpu_fallback_context_creation_fails_loudly(built_library):
    ds = built_library
    with pytest.raises(ds.DsError) as e:
        ds.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_no_kernel_is_compiled_in_two_translation_units(built_library):
    """The exact and the fast translation units are built with different arithmetic flags; a kernel template instantiated in both would be
    an ODR violation (which copy a launch gets is then decided per process).  The ptxas logs of the build list every entry function."""
    obj = ROOT / "deepestscatter_b200" / "csrc" / "_obj"
    logs = sorted(obj.glob("ptxas_*.log"))
    if len(logs) < 3:
        built_library.build_library(force=True)
        logs = sorted(obj.glob("ptxas_*.log"))
    seen = {}
    for log in logs:
        for name in set(re.findall(r"Compiling entry function '(\w+)'", log.read_text())):
            assert name not in seen, f"kernel {name} is compiled in {seen[name]} and in {log.name}"
            seen[name] = log.name
    assert len(seen) > 20
