"""The oracle (oracle/pimc_oracle.py) against the reference: its own known-answer vectors
(tests/pimc/explicit_data, repacked as golden/explicit_kat.npz) and outputs of the running
reference on seeded inputs (golden/cases/*, made by golden/make_golden.py).  CPU only."""
import os
from os.path import join

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import pimc_oracle as orc

RTOL, ATOL = 1e-05, 1e-08   # the reference's own tolerances (tests/pimc/test_pimc_explicit_example.py:22-23)
TIGHT = 1e-11               # oracle vs running reference on identical coordinates


@pytest.fixture(scope="module")
def kat(tmp_path_factory):
    data = np.load(join(GOLDEN, "explicit_kat.npz"))
    root = tmp_path_factory.mktemp("kat")
    paths = {}
    for name in ("coupled_model", "sampling_model"):
        paths[name] = join(root, name + ".json")
        with open(paths[name], "w", encoding="UTF8") as fh:
            fh.write(str(data[name + "_json"]))
    vib = orc.load_vibronic_json(paths["coupled_model"])
    rho = orc.load_sampling_json(paths["sampling_model"])
    tab = orc.precompute(vib, rho, P=5, temperature=300.00)
    return data, tab


def test_beta_known_answer():
    assert np.allclose(orc.beta_of(300.00), 38.68174020133669, rtol=RTOL, atol=ATOL)


def test_kat_precompute(kat):
    data, tab = kat
    w = data["sampling_surface_weights"] / data["sampling_surface_weights"].sum()
    assert np.allclose(w, tab.weights, rtol=RTOL, atol=ATOL)
    assert np.allclose(data["coupled_Edeltas"], tab.delta_vib, rtol=RTOL, atol=ATOL)
    assert np.allclose(data["sampling_Edeltas"], tab.delta_rho, rtol=RTOL, atol=ATOL)
    assert np.allclose(data["coupled_ds"], tab.d_vib, rtol=RTOL, atol=ATOL)
    assert np.allclose(data["sampling_ds"], tab.d_rho, rtol=RTOL, atol=ATOL)
    assert np.allclose(data["coupled_means"][..., 0], tab.d_vib, rtol=RTOL, atol=ATOL)
    assert np.allclose(data["sampling_means"][..., 0], tab.d_rho, rtol=RTOL, atol=ATOL)
    assert np.allclose((data["coupled_cosh"] / data["coupled_sinh"])[0], tab.vib.coth, rtol=RTOL, atol=ATOL)
    assert np.allclose((data["sampling_cosh"] / data["sampling_sinh"])[0], tab.rho.coth, rtol=RTOL, atol=ATOL)
    assert np.allclose((data["coupled_sinh"] ** -1.)[0], tab.vib.csch, rtol=RTOL, atol=ATOL)
    assert np.allclose((data["sampling_sinh"] ** -1.)[0], tab.rho.csch, rtol=RTOL, atol=ATOL)


def test_kat_covariance(kat):
    """sampling_covariance.npy is inv(2 coth I - csch C): the normal-mode sigmas must reproduce it"""
    data, tab = kat
    for n in range(tab.N):
        cov = (tab.ring_eigvecs * tab.sigma[n][None, :] ** 2) @ tab.ring_eigvecs.T
        for a in range(tab.Ar):
            assert np.allclose(data["sampling_covariance"][a, n], cov, rtol=RTOL, atol=ATOL)


def test_kat_block_compute(kat):
    """rho_oMat, vib_oMat, vib_mMat, denominator(rho), numerator(g) for the fixed samples.npy"""
    data, tab = kat
    R = np.ascontiguousarray(data["samples"][:, 0])       # (10, N, P); the reference broadcasts over A
    details = {}
    rho, g = orc.estimate_block(tab, R, pm=False, scale=False, details=details)
    idx = np.arange(tab.A)
    assert np.allclose(details["o_rho"], data["rho_oMat"][:, :, idx, idx], rtol=RTOL, atol=ATOL)
    assert np.allclose(details["o_vib"], data["vib_oMat"][:, :, idx, idx], rtol=RTOL, atol=ATOL)
    assert np.allclose(details["M"], data["vib_mMat"], rtol=RTOL, atol=ATOL)
    assert np.allclose(rho, data["denominator(rho)"], rtol=RTOL, atol=ATOL)
    assert np.allclose(g, data["numerator(g)"], rtol=RTOL, atol=ATOL)
    # off-diagonal of the stored O matrices is exactly zero
    off = data["rho_oMat"].copy()
    off[:, :, idx, idx] = 0
    assert not off.any()


def test_oracle_matches_running_reference(case):
    """quadratic coupling, +/- path, S scaling and A_rho != A are not pinned by the reference's tests:
    compare with what the reference itself produced on the coordinates it drew"""
    tab = case.oracle_tables(rho_trunc=True)
    ref = case.ref
    assert np.allclose(tab.weights, ref["rho_weight"], rtol=1e-13)
    assert np.allclose(tab.d_vib, ref["vib_shift"], rtol=1e-14, atol=0)
    assert np.allclose(tab.d_rho, ref["rho_shift"], rtol=1e-14, atol=0)
    assert np.allclose(tab.vib.prefactor, ref["vib_pref"], rtol=1e-13)
    assert np.allclose(tab.vib_plus.prefactor, ref["vib_pref_plus"], rtol=1e-13)
    assert np.allclose(tab.vib_minus.prefactor, ref["vib_pref_minus"], rtol=1e-13)
    assert np.allclose(tab.ring_eigvals, ref["ring_eigvals"], atol=1e-13)
    for faithful in (False, True):
        out = orc.estimate_block(tab, case.R, pm=True, faithful=faithful)
        for got, want in zip(out, case.expected):
            assert np.max(np.abs(got / want - 1)) < TIGHT


def test_scalar_restatement_matches_block_form(case):
    tab = case.oracle_tables(rho_trunc=True)
    n = min(3, len(case.R))
    block = orc.estimate_block(tab, case.R[:n], pm=True, faithful=False)
    for x in range(n):
        scalar = orc.sample_scalar(tab, case.R[x], pm=True)
        for k in range(4):
            assert abs(scalar[k] / block[k][x] - 1) < 1e-12


def test_rho_trunc_quirk_only_matters_when_rho_is_larger():
    from conftest import GoldenCase
    c = GoldenCase("jt_rho4")
    full = orc.estimate_block(c.oracle_tables(rho_trunc=False), c.R, pm=False, faithful=False, scale=False)
    trunc = orc.estimate_block(c.oracle_tables(rho_trunc=True), c.R, pm=False, faithful=False, scale=False)
    assert np.all(full[0] >= trunc[0]) and np.any(full[0] > trunc[0] * (1 + 1e-6))
    assert np.allclose(full[1], trunc[1], rtol=1e-13)   # g does not depend on rho


def test_oracle_sampler_covariance():
    """draw_block reproduces the analytic ring-polymer covariance (statistical, fixed seed)"""
    from conftest import GoldenCase
    c = GoldenCase("quad_3x4")
    tab = c.oracle_tables()
    rng = np.random.RandomState(5)
    X = 40000
    R = orc.draw_block(tab, np.zeros(X, dtype=int), rng) - tab.d_rho[0][None, :, None]
    for n in range(tab.N):
        emp = R[:, n, :].T @ R[:, n, :] / X
        cov = (tab.ring_eigvecs * tab.sigma[n][None, :] ** 2) @ tab.ring_eigvecs.T
        assert np.max(np.abs(emp - cov)) < 6 * np.max(np.abs(cov)) / np.sqrt(X)


def test_oracle_matches_reference_on_the_whole_paper_family(paper_family):
    """c3: 54 (model, sampling distribution) pairs of examples/paper_1.5025058 x 2 temperatures, P=128, incl. the
    jahnteller pairs whose sampling model has 4 or 8 surfaces on a 2-surface system (reference quirk Q1).
    The oracle follows the reference's formulas and order of operations, so it reproduces the reference to 4e-12 --
    including the reference's own rounding error, which reaches 3.9e-10 against 80-bit arithmetic on this family."""
    worst, worst_ref_exact = 0.0, 0.0
    assert len(paper_family.names) == 108
    for name in paper_family.names:
        vib, rho, T, R, want = paper_family.run(name)
        tab = orc.precompute(vib, rho, paper_family.P, T, rho_trunc=True)
        got = np.stack(orc.estimate_block(tab, R, pm=True, faithful=True))
        err = np.max(np.abs(got - want) / np.abs(want))
        worst = max(worst, err)
        assert err < 1e-10, (name, err)
        exact = paper_family.exact(name)
        worst_ref_exact = max(worst_ref_exact, np.max(np.abs(want - exact) / np.abs(exact)))
    assert 1e-10 < worst_ref_exact < 1e-9       # documents the conditioning of the family
    print("c3 family: oracle vs reference %.2e, reference vs 80-bit %.2e" % (worst, worst_ref_exact))
