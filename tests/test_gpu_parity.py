"""GPU parity tests: the CUDA path, called through the C ABI, against the reference's outputs
(golden fixtures made by tests/golden/make_golden.py) and against the numpy oracle.

Tolerance: BASELINE.json north_star asks for per-sample g/rho within 1e-10 relative in FP64."""
import numpy as np
import pytest

from pibronic_b200 import _cabi

pytestmark = pytest.mark.gpu

RTOL = 1e-10

VARIANTS = {
    "default": 0,
    "jacobi": _cabi.FLAG_EIG_JACOBI,
    "generic": _cabi.FLAG_FORCE_GENERIC,
    "generic_jacobi": _cabi.FLAG_FORCE_GENERIC | _cabi.FLAG_EIG_JACOBI,
}


def rel_err(got, want):
    return np.max(np.abs(got - want) / np.abs(want))


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_eval_coords_matches_reference(cuda, case, variant):
    """feed the reference's own numpy-drawn coordinates; rho, g, g+, g- and the ratios must match"""
    flags = _cabi.FLAG_PM | _cabi.QUIRK_RHO_TRUNC | VARIANTS[variant]
    plan = case.plan(flags)
    out = plan.eval_coords_host(case.R)
    want = case.expected
    for k, name in enumerate(("rho", "g", "g+", "g-")):
        assert rel_err(out[k], want[k]) < RTOL, f"{case.name}/{variant}: {name} rel err {rel_err(out[k], want[k]):.3e}"
    for k in (1, 2, 3):
        assert rel_err(out[k] / out[0], want[k] / want[0]) < RTOL
    plan.close()
