"""include/ds_detmath.h (the arithmetic contract shared by oracle and kernels) against numpy float64."""
import numpy as np

import oracle_lib as ol


def _probe(fn, x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    a, b = np.empty_like(x), np.empty_like(x)
    ol.lib().orc_math_probe(fn, x, len(x), a, b)
    return a, b


def _ulp_err(got, ref64):
    ref32 = ref64.astype(np.float32)
    return np.abs(got.astype(np.float64) - ref64) / np.spacing(np.abs(ref32)).astype(np.float64)


def test_expf_accuracy_and_exact_identities():
    x = np.linspace(-87, 0, 400001)
    got, _ = _probe(0, x)
    assert _ulp_err(got, np.exp(x.astype(np.float32).astype(np.float64))).max() <= 2.0
    z, _ = _probe(0, np.array([0.0, -0.0, -100.0, -1e30]))
    assert z[0] == 1.0 and z[1] == 1.0  # empty-space skipping relies on exp(-0) == 1 exactly
    assert z[2] == 0.0 and z[3] == 0.0


def test_logf_accuracy():
    x = np.exp(np.linspace(np.log(2.0**-20), np.log(2.0**20), 400001)).astype(np.float32)
    got, _ = _probe(1, x)
    ref = np.log(x.astype(np.float64))
    err = np.abs(got - ref)
    assert (err <= np.maximum(2.0 * np.spacing(np.abs(ref).astype(np.float32)), 1e-7)).all()
    one, _ = _probe(1, np.array([1.0]))
    assert one[0] == 0.0


def test_sincos_accuracy_on_the_used_range():
    phi = np.linspace(0, 2 * np.pi, 400001).astype(np.float32)
    s, c = _probe(2, phi)
    assert np.abs(s - np.sin(phi.astype(np.float64))).max() < 2e-7
    assert np.abs(c - np.cos(phi.astype(np.float64))).max() < 2e-7
    assert np.abs(s * s + c * c - 1).max() < 5e-7


def test_log2_exp2():
    x = np.array([0.0267, 0.5, 1.0, 2.7, 100.0], dtype=np.float32)
    a, _ = _probe(3, x)
    assert np.allclose(a, np.log2(x.astype(np.float64)), rtol=0, atol=1e-6)
    b, _ = _probe(4, np.array([-3.0, -0.5, 0.0, 1.0, 4.25], dtype=np.float32))
    assert np.allclose(b, 2.0 ** np.array([-3.0, -0.5, 0.0, 1.0, 4.25]), rtol=3e-7)
