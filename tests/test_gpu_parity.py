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
    "fused_dmma": _cabi.FLAG_PREFER_DMMA,        # one-launch tensor-core kernel also where a register-resident one exists
    "blocked": _cabi.FLAG_NO_FUSED_DMMA,         # round-1 large-A path through HBM scratch
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


def test_paper_family_matches_reference(cuda, paper_family):
    """c3 (BASELINE configs[2]): all 54 (model, sampling distribution) pairs of examples/paper_1.5025058 at both ends
    of the temperature sweep, P=128, on the coordinates the reference drew.  The reference's own float64 values are
    off by up to 3.9e-10 on this family (its coth (q^2+q'^2) - 2 csch q q' cancels 4 digits at tau*omega = 0.008 and
    the product of 128 bead matrices amplifies it), so each of rho, g, g+, g- must be
      (i)  within 1e-10 of the 80-bit evaluation of the same formulas (tests/golden/extended_precision.py), and
      (ii) within 1e-10 + |reference - 80-bit| of the reference;
    then the fused sampler+estimator against the oracle on the coordinates the device sampler reports."""
    import sys
    from conftest import GOLDEN
    from oracle import pimc_oracle as orc
    sys.path.insert(0, GOLDEN)
    import extended_precision
    worst_exact, worst_ref, n_strict = 0.0, 0.0, 0
    for name in paper_family.names:
        vib, rho, T, R, want = paper_family.run(name)
        exact = paper_family.exact(name)
        plan = _cabi.Plan(vib["E"], vib["w"], vib["L"], vib["Q"], rho["E"], rho["w"], rho["L"], paper_family.P,
                          orc.beta_of(T), orc.DELTA_BETA, flags=_cabi.FLAG_PM | _cabi.QUIRK_RHO_TRUNC, device=0)
        assert plan.is_fast, name
        got = plan.eval_coords_host(R)
        err_exact = np.abs(got - exact) / np.abs(exact)
        err_ref = np.abs(got - want) / np.abs(want)
        ref_exact = np.abs(want - exact) / np.abs(exact)
        assert err_exact.max() < RTOL, (name, err_exact.max())
        assert np.all(err_ref < RTOL + ref_exact), (name, err_ref.max())
        worst_exact, worst_ref = max(worst_exact, err_exact.max()), max(worst_ref, err_ref.max())
        n_strict += int(err_ref.max() < RTOL)
        if name.endswith("_T350"):
            n = 96
            fused = plan.sample_eval_host(31, 7, n)
            Rd, _ = drawn_coords(cuda, plan, 31, 7, n)
            want_d = extended_precision.evaluate(vib, rho, paper_family.P, T, Rd).astype(np.float64)
            assert rel_err(fused, want_d) < RTOL, name
        plan.close()
    print("c3 family: CUDA vs 80-bit %.2e, CUDA vs reference %.2e (%d of %d runs within 1e-10 of the reference)"
          % (worst_exact, worst_ref, n_strict, len(paper_family.names)))


# ------------------------------------------------------------------ helpers
def oracle_eval(tab, R, pm=True):
    from oracle import pimc_oracle as orc
    return np.stack(orc.estimate_block(tab, R, pm=pm, faithful=False))


def drawn_coords(cuda, plan, seed, first, n):
    with cuda.cuda.device(plan.device):
        R = cuda.empty((n, plan.N, plan.P), dtype=cuda.float64, device="cuda")
        src = cuda.empty(n, dtype=cuda.int32, device="cuda")
        plan.sample_coords(seed, first, n, R, src)
        return R.cpu().numpy(), src.cpu().numpy()


# ------------------------------------------------------------------ reference KAT, step by step
def test_stage_kernels_reproduce_the_references_known_answers(cuda, tmp_path):
    """the reference's tests/pimc/test_pimc_explicit_example.py::test_block_compute, run against this
    package's step helpers (CUDA stage kernels): rho_oMat, vib_oMat, vib_mMat, rho, g"""
    from os.path import join
    from conftest import GOLDEN
    from pibronic_b200 import file_structure, pimc
    from pibronic_b200.pimc import build_o_matrix, build_denominator, diagonalize_coupling_matrix, build_numerator
    kat = np.load(join(GOLDEN, "explicit_kat.npz"))
    FS = file_structure.FileStructure(tmp_path, id_data=0, id_rho=0)
    for name, path in (("coupled_model", FS.path_vib_model), ("sampling_model", FS.path_rho_model)):
        with open(path, "w", encoding="UTF8") as fh:
            fh.write(str(kat[name + "_json"]))
    data = pimc.BoxData.from_FileStructure(FS)
    data.samples, data.beads, data.temperature, data.block_size = 10, 5, 300.00, 2
    data.blocks = data.samples // data.block_size
    data.preprocess()
    results = pimc.BoxResult(data=data)
    results.path_root, results.id_job = FS.path_rho_results, 0
    rho, vib = data.rho, data.vib
    y_rho, y_g = results.scaled_rho.view(), results.scaled_g.view()
    RTOL_, ATOL_ = 1e-05, 1e-08
    for block_index in range(data.blocks):
        view = slice(block_index * data.block_size, (block_index + 1) * data.block_size)
        data.qTensor = kat["samples"][view, ...]
        build_o_matrix(data, rho.const, rho.state_shift)
        assert np.allclose(rho.const.omatrix, kat["rho_oMat"][view, ...], rtol=RTOL_, atol=ATOL_)
        build_o_matrix(data, vib.const, vib.state_shift)
        assert np.allclose(vib.const.omatrix, kat["vib_oMat"][view, ...], rtol=RTOL_, atol=ATOL_)
        build_denominator(rho.const, y_rho, view)
        diagonalize_coupling_matrix(data)
        build_numerator(data, vib.const, y_g, view)
        assert np.allclose(data.M_matrix, kat["vib_mMat"][view, ...], rtol=RTOL_, atol=ATOL_)
    assert np.allclose(y_rho, kat["denominator(rho)"], rtol=RTOL_, atol=ATOL_)
    assert np.allclose(y_g, kat["numerator(g)"], rtol=RTOL_, atol=ATOL_)
    # and far tighter than the reference's own tolerance
    assert rel_err(y_g, kat["numerator(g)"]) < RTOL and rel_err(y_rho, kat["denominator(rho)"]) < RTOL
    data.release()


def test_stage_outputs_match_oracle(cuda, case):
    """O factors (all four table sets), scale S, V and M for every bead against the oracle's intermediates"""
    from oracle import pimc_oracle as orc
    tab = case.oracle_tables(rho_trunc=True)
    plan = case.plan()
    details = {}
    orc.estimate_block(tab, case.R, pm=True, faithful=False, details=details)
    n, P, A, Ar = len(case.R), tab.P, tab.A, tab.Ar
    t = cuda
    with t.cuda.device(0):
        R = t.from_numpy(case.R).cuda()
        o_rho = t.empty((n, P, Ar), dtype=t.float64, device="cuda")
        o_vib = t.empty((3, n, P, A), dtype=t.float64, device="cuda")
        scale = t.empty((n, P), dtype=t.float64, device="cuda")
        v_mat = t.empty((n, P, A, A), dtype=t.float64, device="cuda")
        m_mat = t.empty((n, P, A, A), dtype=t.float64, device="cuda")
        plan.eval_stages(R, o_rho=o_rho, o_vib=o_vib, scale=scale, v_mat=v_mat, m_mat=m_mat)
        o_rho, o_vib, scale, v_mat, m_mat = (x.cpu().numpy() for x in (o_rho, o_vib, scale, v_mat, m_mat))
    assert rel_err(scale, details["S"]) < 1e-11
    keep = details["o_rho"] > 1e-280
    assert np.max(np.abs(o_rho[keep] / details["o_rho"][keep] - 1)) < 1e-11
    keep = details["o_vib"] > 1e-280
    assert np.max(np.abs(o_vib[0][keep] / details["o_vib"][keep] - 1)) < 1e-11
    vscale = np.abs(details["V"]).max() + 1e-300
    assert np.max(np.abs(v_mat - details["V"])) < 1e-13 * vscale
    mscale = np.abs(details["M"]).max(axis=(2, 3), keepdims=True)
    assert np.max(np.abs(m_mat - details["M"]) / mscale) < 1e-12
    plan.close()


# ------------------------------------------------------------------ sampler
def test_device_sampler_matches_its_cpu_restatement(cuda, case):
    from oracle import philox_sampler as ps
    tab = case.oracle_tables()
    plan = case.plan()
    samp = plan.table("samp").reshape(tab.P, tab.N, 3)
    wcum = np.cumsum(plan.table("weights"))
    n, seed, first = 257, 0xDEADBEEF12345678, (1 << 33) + 5      # 64-bit seed and counter
    R, src = drawn_coords(cuda, plan, seed, first, n)
    R_cpu, src_cpu = ps.sample_coords(samp, wcum, tab.d_rho, seed, first, n)
    assert np.array_equal(src, src_cpu)
    assert np.max(np.abs(R - R_cpu)) < 1e-11 * (1 + np.abs(R_cpu).max())
    plan.close()


@pytest.mark.parametrize("variant", ["default", "generic", "fused_dmma"])
def test_fused_kernel_equals_sampler_then_estimator(cuda, case, variant):
    """the fused sampler+estimator gives what the estimator gives on the coordinates the sampler kernel
    reports for the same Philox counters -- and both match the oracle on those coordinates"""
    flags = _cabi.FLAG_PM | _cabi.QUIRK_RHO_TRUNC | VARIANTS[variant]
    plan = case.plan(flags)
    tab = case.oracle_tables(rho_trunc=True)
    n, seed, first = 300, 424242, 12345
    fused = plan.sample_eval_host(seed, first, n)
    R, _ = drawn_coords(cuda, plan, seed, first, n)
    two_step = plan.eval_coords_host(R)
    assert rel_err(fused, two_step) < 1e-12
    want = oracle_eval(tab, R)
    for k in range(4):
        assert rel_err(fused[k], want[k]) < RTOL
    plan.close()


@pytest.mark.parametrize("pm", [True, False])
def test_warp_specialised_kernel_is_bit_identical(cuda, case, pm):
    """producer/consumer kernel (default for the fused call) == one-role kernel, same Philox counters and arithmetic;
    n = 700 leaves a partly filled CTA, first_sample != 0 exercises the 64-bit counter"""
    base = (_cabi.FLAG_PM if pm else 0) | _cabi.QUIRK_RHO_TRUNC
    ws, one = case.plan(base), case.plan(base | _cabi.FLAG_NO_WARPSPEC)
    if not ws.is_fast:
        pytest.skip("shape runs on the generic kernels")
    rows = 4 if pm else 2
    for n, first in ((700, (1 << 33) + 5), (1, 0), (256, 256)):
        a = ws.sample_eval_host(99, first, n, out4=np.full((rows, n), np.nan))
        b = one.sample_eval_host(99, first, n, out4=np.full((rows, n), np.nan))
        assert np.all(np.isfinite(a)) and np.array_equal(a, b), (n, first)
    ws.close()
    one.close()


def test_pinned_host_buffer_is_written_by_the_kernel_itself(cuda, case):
    """a pinned (mapped) host buffer takes the zero-copy route -- the kernel stores the host rows directly --
    a pageable one the staged D2H copy; same bytes either way, also with a row stride and on the large-delta redo path"""
    plan = case.plan()
    n, ld = 777, 1000
    pageable = plan.sample_eval_host(5, 10, n)
    pinned = cuda.full((4, ld), float("nan"), dtype=cuda.float64).pin_memory().numpy()
    got, sums = plan.sample_eval_host(5, 10, n, out4=pinned[:, :n], block_size=100)
    assert np.array_equal(got, pageable) and np.isnan(pinned[:, n:]).all()
    _, sums_pageable = plan.sample_eval_host(5, 10, n, block_size=100)
    assert np.array_equal(sums, sums_pageable)
    plan.close()


def test_results_do_not_depend_on_how_the_index_range_is_split(cuda, case):
    plan = case.plan()
    n, seed = 1000, 77
    whole = plan.sample_eval_host(seed, 500, n)
    parts = np.concatenate([plan.sample_eval_host(seed, 500, 333), plan.sample_eval_host(seed, 833, 667)], axis=1)
    assert np.array_equal(whole, parts)
    again = plan.sample_eval_host(seed, 500, n)
    assert np.array_equal(whole, again)                                   # deterministic
    other = plan.sample_eval_host(seed + 1, 500, n)
    assert not np.array_equal(whole[1], other[1])
    plan.close()


# ------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("n", [1, 31, 129])
def test_ragged_sample_counts(cuda, n):
    from conftest import GoldenCase
    case = GoldenCase("quad_3x4")
    plan = case.plan()
    tab = case.oracle_tables()
    out = plan.sample_eval_host(5, 0, n)
    R, _ = drawn_coords(cuda, plan, 5, 0, n)
    want = oracle_eval(tab, R)
    assert out.shape == (4, n) and rel_err(out, want) < RTOL
    plan.close()


def test_empty_input_is_a_no_op(cuda):
    from conftest import GoldenCase
    case = GoldenCase("c1_2x2")
    plan = case.plan()
    out = plan.sample_eval_host(5, 0, 0, out4=np.full((4, 0), np.nan))
    assert out.shape == (4, 0)
    plan.close()


@pytest.mark.parametrize("P", [3, 4, 33])
@pytest.mark.parametrize("variant", ["default", "generic"])
def test_minimum_and_odd_bead_counts(cuda, P, variant):
    from conftest import GoldenCase
    from oracle import pimc_oracle as orc
    case = GoldenCase("quad_3x4")
    case.P = P
    plan = case.plan(_cabi.FLAG_PM | VARIANTS[variant])
    tab = orc.precompute(case.vib, case.rho, P, case.T)
    n = 64
    out = plan.sample_eval_host(9, 0, n)
    R, _ = drawn_coords(cuda, plan, 9, 0, n)
    assert rel_err(out, oracle_eval(tab, R)) < RTOL
    plan.close()


@pytest.mark.parametrize("variant", ["default", "generic", "fused_dmma"])
@pytest.mark.parametrize("name", ["c1_2x2", "quad_3x4", "c2_4x6"])
def test_strong_coupling_takes_the_squaring_path(cuda, name, variant):
    """few beads and a large inter-surface coupling: ||tau V|| is far above 1/3, so exp(-tau V) runs its squaring loop
    (after the closed form for two surfaces, after the four-product polynomial otherwise)"""
    import copy
    from conftest import GoldenCase
    from oracle import pimc_oracle as orc
    case = GoldenCase(name)
    case.P = 4
    case.vib = copy.deepcopy(case.vib)
    A = case.vib["E"].shape[0]
    case.vib["E"] = case.vib["E"] + 0.15 * (1.0 - np.eye(A))
    tau_v = orc.beta_of(case.T) / case.P * 0.15 * (A - 1)
    assert tau_v > 1.0
    plan = case.plan(_cabi.FLAG_PM | VARIANTS[variant])
    tab = orc.precompute(case.vib, case.rho, case.P, case.T)
    n = 96
    out = plan.sample_eval_host(5, 0, n)
    R, _ = drawn_coords(cuda, plan, 5, 0, n)
    assert rel_err(out, oracle_eval(tab, R)) < RTOL
    plan.close()


def test_non_pm_plan_fills_two_rows(cuda, case):
    plan = case.plan(flags=_cabi.QUIRK_RHO_TRUNC)
    out = np.full((2, len(case.R)), np.nan)
    plan.eval_coords_host(case.R, out4=out)
    assert rel_err(out, case.expected[:2]) < RTOL
    plan.close()


def test_sampling_model_with_fewer_surfaces_than_the_system(cuda):
    """A_rho < A: the reference raises IndexError (pimc.py:1110-1111); mathematically it is fine"""
    from conftest import GoldenCase
    from oracle import pimc_oracle as orc
    case = GoldenCase("quad_3x4")
    case.rho = dict(A=2, N=case.rho["N"], E=case.rho["E"][:2].copy(), w=case.rho["w"], L=case.rho["L"][:, :2].copy())
    plan = case.plan(_cabi.FLAG_PM)
    assert not plan.is_fast
    tab = orc.precompute(case.vib, case.rho, case.P, case.T)
    out = plan.sample_eval_host(3, 0, 50)
    R, src = drawn_coords(cuda, plan, 3, 0, 50)
    assert set(src) <= {0, 1} and rel_err(out, oracle_eval(tab, R)) < RTOL
    plan.close()


def test_correct_rho_when_sampling_model_is_larger(cuda):
    """default (no quirk flag): all A_rho surfaces enter rho(R) -- the mathematically correct estimator"""
    from conftest import GoldenCase
    case = GoldenCase("jt_rho4")
    plan = case.plan(_cabi.FLAG_PM)
    got = plan.eval_coords_host(case.R)
    want = oracle_eval(case.oracle_tables(rho_trunc=False), case.R)
    assert rel_err(got, want) < RTOL
    assert np.any(np.abs(got[0] / case.expected[0] - 1) > 1e-3)      # differs from the reference's truncated rho
    plan.close()


# ------------------------------------------------------------------ block sums
def test_block_sums_match_numpy(cuda, case):
    from oracle import pimc_oracle as orc
    plan = case.plan()
    n, bs = 1000, 96                                                     # ragged last block
    out, sums = plan.sample_eval_host(11, 0, n, block_size=bs)
    ratio, d1, d2 = orc.property_terms(orc.DELTA_BETA, *out)
    cols = (ratio, out[2] / out[0], out[3] / out[0], ratio ** 2, d1, d2, d1 ** 2, d2 ** 2)
    assert sums.shape == (-(-n // bs), _cabi.NSUMS)
    for b in range(sums.shape[0]):
        sl = slice(b * bs, min((b + 1) * bs, n))
        for k, col in enumerate(cols):
            assert np.isclose(sums[b, k], col[sl].sum(), rtol=1e-11, atol=1e-11 * np.abs(col[sl]).sum())
    plan.close()


# ------------------------------------------------------------------ statistics on independent draws
def test_estimators_agree_with_reference_sampler_within_error(cuda):
    """Z, E, Cv from GPU draws vs the numpy port of the reference on its own (MT19937) draws"""
    from conftest import GoldenCase
    from oracle import pimc_oracle as orc
    case = GoldenCase("quad_3x4")
    tab = case.oracle_tables()
    plan = case.plan()
    Xg, Xc = 400000, 40000
    gpu = plan.sample_eval_host(2024, 0, Xg)
    cpu = orc.run_blocks(tab, Xc, 10000, np.random.RandomState(99), pm=True, faithful=False)

    def blocks(out, nb=40):
        """Z, E, Cv per block of samples: their spread is the (jackknife-like) error estimate"""
        vals = []
        for part in np.array_split(np.arange(out.shape[1]), nb):
            r, d1, d2 = orc.property_terms(orc.DELTA_BETA, *out[:, part])
            p = orc.basic_properties(len(part), case.T, r, d1, d2)
            vals.append([p["Z"], p["E"], p["Cv"]])
        vals = np.array(vals)
        return vals.mean(axis=0), vals.std(axis=0, ddof=1) / np.sqrt(nb)
    (mg, eg), (mc, ec) = blocks(gpu), blocks(cpu)
    for k, name in enumerate(("Z", "E", "Cv")):
        assert abs(mg[k] - mc[k]) < 4.5 * np.hypot(eg[k], ec[k]), f"{name}: gpu {mg[k]} +- {eg[k]}, cpu {mc[k]} +- {ec[k]}"
    plan.close()


def test_c5_block_estimators_agree_with_the_reference_run(cuda):
    """BASELINE.json configs[4]: c2 model sampled from another rho, P=64, jackknife Z / E / Cv over 64 blocks.  The
    reference's own run (64 blocks x 100 samples, tests/golden/c5_stats.npz made by make_golden.py::pack_c5_stats) and
    this path (64 blocks x 1e4 samples, independent Philox draws) must agree within their combined jackknife errors"""
    from os.path import join
    from conftest import GOLDEN, GoldenCase
    from oracle import pimc_oracle as orc
    kat = np.load(join(GOLDEN, "c5_stats.npz"))
    ref = dict(zip([str(k) for k in kat["keys"]], kat["values"]))
    P, T, blocks = int(kat["P"]), float(kat["T"]), int(kat["blocks"])
    case = GoldenCase("c5_altrho")
    plan = _cabi.Plan(case.vib["E"], case.vib["w"], case.vib["L"], case.vib["Q"], case.rho["E"], case.rho["w"],
                      case.rho["L"], P, orc.beta_of(T), orc.DELTA_BETA, flags=_cabi.FLAG_PM, device=0)
    B = 10_000
    out, sums = plan.sample_eval_host(64, 0, blocks * B, block_size=B)
    mine = plan.stats_last()
    assert sums.shape == (blocks, _cabi.NSUMS)
    for value, error in (("Z", "Z error"), ("jk_E", "jk_E error"), ("jk_Cv", "jk_Cv error")):
        sigma = np.hypot(mine[error], ref[error])
        assert abs(mine[value] - ref[value]) < 4.5 * sigma, (value, mine[value], ref[value], sigma)
        assert mine[error] < ref[error]                      # 100x the samples: smaller error bars
    # the 64 block means scatter around the global mean as their standard error says
    block_Z = sums[:, 0] / B
    assert abs(block_Z.mean() - mine["Z"]) < 1e-12 * abs(mine["Z"]) + 1e-15
    assert 0.5 < block_Z.std(ddof=1) / (mine["Z error"] * np.sqrt(blocks)) < 2.0
    plan.close()


def test_mixture_frequencies_and_bead_covariance_on_device(cuda):
    from conftest import GoldenCase
    case = GoldenCase("jt_rho4")
    tab = case.oracle_tables()
    plan = case.plan()
    X = 200000
    R, src = drawn_coords(cuda, plan, 31337, 0, X)
    freq = np.bincount(src, minlength=tab.Ar) / X
    assert np.max(np.abs(freq - tab.weights)) < 4.5 * np.sqrt(0.25 / X)
    y = R - tab.d_rho[src][:, :, None]
    for n in range(tab.N):
        emp = y[:, n, :].T @ y[:, n, :] / X
        cov = (tab.ring_eigvecs * tab.sigma[n][None, :] ** 2) @ tab.ring_eigvecs.T
        assert np.max(np.abs(emp - cov)) < 6 * np.max(np.abs(cov)) / np.sqrt(X)
    z = y[:, 0, 0] / np.sqrt(cov[0, 0]) if tab.N == 1 else None
    plan.close()


# ------------------------------------------------------------------ BASELINE sizes: size-independent properties
def test_full_size_c2_properties(cuda):
    """A=4, N=6, P=64, X=1e6 (BASELINE.json configs[1]): the oracle cannot run this; check finiteness,
    positivity of rho, determinism, index-split invariance, agreement of the +/- variants with g to
    O(delta_beta), block-sum consistency (a checksum of checksums) and the mean against a small oracle run"""
    from oracle import pimc_oracle as orc
    from pibronic_b200 import constants, synthetic
    from pibronic_b200.model_io import VMK
    model = synthetic.model_c2()
    rho = synthetic.diagonal_of(model)
    plan = _cabi.Plan(model[VMK.E], model[VMK.w], model[VMK.G1], model[VMK.G2], rho[VMK.E], rho[VMK.w], rho[VMK.G1],
                      64, constants.beta(300.0), constants.delta_beta, flags=_cabi.FLAG_PM, device=0)
    assert plan.is_fast
    X, bs = 1_000_000, 10_000
    out, sums = plan.sample_eval_host(20260417, 0, X, block_size=bs)
    assert np.isfinite(out).all() and (out[0] > 0).all() and (out[0] <= 4 * (1 + 1e-12)).all()
    again, _ = plan.sample_eval_host(20260417, 0, X, block_size=bs)
    assert np.array_equal(out, again)
    tail = plan.sample_eval_host(20260417, X - 4097, 4097)
    assert np.array_equal(tail, out[:, X - 4097:])
    ratio = out[1] / out[0]
    assert np.isclose(sums[:, 0].sum(), ratio.sum(), rtol=1e-12)
    assert np.max(np.abs(out[2] / out[1] - 1)) < 0.05 and np.max(np.abs(out[3] / out[1] - 1)) < 0.05
    vib_d = dict(A=4, N=6, E=model[VMK.E], w=model[VMK.w], L=model[VMK.G1], Q=model[VMK.G2])
    rho_d = dict(A=4, N=6, E=rho[VMK.E], w=rho[VMK.w], L=rho[VMK.G1])
    tab = orc.precompute(vib_d, rho_d, 64, 300.0)
    cpu = orc.run_blocks(tab, 4000, 1000, np.random.RandomState(1), pm=False, faithful=False)
    rc = cpu[1] / cpu[0]
    err = np.hypot(ratio.std() / np.sqrt(X), rc.std() / np.sqrt(len(rc)))
    assert abs(ratio.mean() - rc.mean()) < 4.5 * err
    # spot-check 64 of the million samples against the oracle on the coordinates the sampler reports
    R, _ = drawn_coords(cuda, plan, 20260417, 777_000, 64)
    assert rel_err(out[:, 777_000:777_064], oracle_eval(tab, R)) < RTOL
    plan.close()


def _c4_plan(extra=0, pm=True):
    from pibronic_b200 import constants, synthetic
    from pibronic_b200.model_io import VMK
    model = synthetic.model_c4()
    rho = synthetic.diagonal_of(model)
    plan = _cabi.Plan(model[VMK.E], model[VMK.w], model[VMK.G1], model[VMK.G2], rho[VMK.E], rho[VMK.w], rho[VMK.G1],
                      256, constants.beta(300.0), constants.delta_beta, flags=(_cabi.FLAG_PM if pm else 0) | extra, device=0)
    vib_d = dict(A=12, N=24, E=model[VMK.E], w=model[VMK.w], L=model[VMK.G1], Q=model[VMK.G2])
    rho_d = dict(A=12, N=24, E=rho[VMK.E], w=rho[VMK.w], L=rho[VMK.G1])
    return plan, vib_d, rho_d


def test_c4_shape_runs_on_the_fused_tensor_core_kernel(cuda):
    """A=12, N=24, P=256 (BASELINE.json configs[3]): one launch of the fused large-A kernel (sampler on chip, no scratch)
    against the oracle on the coordinates the sampler kernel reports for the same Philox counters, and against the
    blocked kernels of round 1; n = 19 leaves most warps of the CTAs without a sample"""
    from oracle import pimc_oracle as orc
    plan, vib_d, rho_d = _c4_plan()
    assert not plan.is_fast and plan.kernel_path == _cabi.PATH_FUSED_DMMA
    n = 19
    launches = plan.launch_count
    out = plan.sample_eval_host(1, 0, n)
    assert plan.launch_count - launches == 1                      # sampler + estimator in ONE kernel
    R, _ = drawn_coords(cuda, plan, 1, 0, n)
    tab = orc.precompute(vib_d, rho_d, 256, 300.0)
    assert rel_err(out, oracle_eval(tab, R)) < RTOL
    assert np.array_equal(out, plan.eval_coords_host(R))          # same arithmetic on caller supplied coordinates
    blocked, _, _ = _c4_plan(_cabi.FLAG_NO_FUSED_DMMA)
    assert blocked.kernel_path == _cabi.PATH_BLOCKED
    assert rel_err(blocked.sample_eval_host(1, 0, n), out) < RTOL
    blocked.close()
    plan.close()


@pytest.mark.parametrize("shape", [(16, 24, 20), (14, 20, 33), (13, 5, 16), (16, 3, 7)])
def test_more_than_twelve_surfaces_run_fused(cuda, shape):
    """13..16 surfaces (two 8 x 8 tiles per matrix row, the widest the kernel is built for): the fused tensor-core kernel
    with eight warps per CTA, or four when the per-warp regions of the shape do not fit the SM eight times (16 x 24);
    against the oracle on the sampler's co-ordinates, the co-ordinate entry point and the blocked kernels; more samples
    than one pass of the resident warps"""
    from oracle import pimc_oracle as orc
    from pibronic_b200 import constants, synthetic
    from pibronic_b200.model_io import VMK
    A, N, P = shape
    model = synthetic.coupled_model(A, N, (0.1, 0.39), (14.0, 14.8), seed=A * 100 + N)
    rho = synthetic.diagonal_of(model)
    args = (model[VMK.E], model[VMK.w], model[VMK.G1], model[VMK.G2], rho[VMK.E], rho[VMK.w], rho[VMK.G1],
            P, constants.beta(300.0), constants.delta_beta)
    plan = _cabi.Plan(*args, flags=_cabi.FLAG_PM, device=0)
    assert plan.kernel_path == _cabi.PATH_FUSED_DMMA
    vib_d = dict(A=A, N=N, E=model[VMK.E], w=model[VMK.w], L=model[VMK.G1], Q=model[VMK.G2])
    rho_d = dict(A=A, N=N, E=rho[VMK.E], w=rho[VMK.w], L=rho[VMK.G1])
    tab = orc.precompute(vib_d, rho_d, P, 300.0)
    n = 40
    out = plan.sample_eval_host(3, 17, n)
    R, _ = drawn_coords(cuda, plan, 3, 17, n)
    assert rel_err(out, oracle_eval(tab, R)) < RTOL
    assert np.array_equal(out, plan.eval_coords_host(R))
    many = plan.sample_eval_host(3, 0, 148 * 8 + 57)
    assert np.array_equal(many[:, 17:17 + n], out)
    blocked = _cabi.Plan(*args, flags=_cabi.FLAG_PM | _cabi.FLAG_NO_FUSED_DMMA, device=0)
    assert blocked.kernel_path == _cabi.PATH_BLOCKED
    assert rel_err(blocked.sample_eval_host(3, 17, n), out) < RTOL
    blocked.close()
    plan.close()


def test_random_shapes_against_the_oracle(cuda):
    """tools/fuzz_shapes.py: 40 random (A <= 16, N, A_rho, P, T, PM) models, four kernel selections each (default, one-role +
    Jacobi, tensor-core preferred, generic forced) against the oracle on the sampler's co-ordinates, fused == co-ordinate entry"""
    import importlib.util
    from os.path import dirname, join
    spec = importlib.util.spec_from_file_location("fuzz_shapes", join(dirname(dirname(__file__)), "tools", "fuzz_shapes.py"))
    fuzz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fuzz)
    worst, failures, seen = fuzz.run(40, 2026, verbose=False)
    assert not failures, "\n".join(failures)
    assert worst < RTOL and {"tensor", "register", "blocked"} <= set(seen)


def test_fused_tensor_core_kernel_many_passes_and_splits(cuda):
    """more samples than resident warps (148 SMs x 8): several passes per warp, a partly filled last pass, results
    independent of how the index range is cut, non-PM variant fills two rows"""
    plan, _, _ = _c4_plan()
    n, seed, first = 148 * 8 + 77, 5, (1 << 33) + 11
    whole = plan.sample_eval_host(seed, first, n)
    assert np.all(np.isfinite(whole)) and np.all(whole[0] > 0)
    parts = np.concatenate([plan.sample_eval_host(seed, first, 300), plan.sample_eval_host(seed, first + 300, n - 300)], axis=1)
    assert np.array_equal(whole, parts)
    plan.close()
    plain, _, _ = _c4_plan(pm=False)
    two = plain.sample_eval_host(seed, first, 64, out4=np.full((2, 64), np.nan))
    assert np.array_equal(two, whole[:2, :64])
    plain.close()


@pytest.mark.parametrize("P", [3, 16, 17, 33])
def test_fused_tensor_core_kernel_bead_counts(cuda, P):
    """bead counts around the 16-bead group size of the fused kernel (ring closure inside / at the edge of a group)"""
    from conftest import GoldenCase
    from oracle import pimc_oracle as orc
    case = GoldenCase("syn_7x12")
    flags = _cabi.FLAG_PM | _cabi.QUIRK_RHO_TRUNC
    plan = _cabi.Plan(case.vib["E"], case.vib["w"], case.vib["L"], case.vib["Q"], case.rho["E"], case.rho["w"],
                      case.rho["L"], P, orc.beta_of(case.T), orc.DELTA_BETA, flags=flags, device=0)
    assert plan.kernel_path == _cabi.PATH_FUSED_DMMA
    tab = orc.precompute(case.vib, case.rho, P, case.T, rho_trunc=True)
    out = plan.sample_eval_host(9, 100, 40)
    R, _ = drawn_coords(cuda, plan, 9, 100, 40)
    assert rel_err(out, oracle_eval(tab, R)) < RTOL
    plan.close()


# ------------------------------------------------------------------ the drop-in facade end to end
@pytest.mark.parametrize("pm", [False, True])
def test_block_compute_through_the_facade(cuda, tmp_path, pm):
    """mirrors the reference's smoke tests (tests/pimc/test_pimc_general.py:74-155) but WITH numeric checks:
    FileStructure -> BoxData[PM] -> preprocess -> block_compute[_pm] -> .npz -> load_multiple_results"""
    from os.path import isfile, join
    from oracle import pimc_oracle as orc
    from pibronic_b200 import file_structure, pimc, synthetic
    FS = file_structure.FileStructure(tmp_path, 0, 0)
    synthetic.write_data_set(FS, synthetic.coupled_model(2, 2, (0.02, 0.04), (0.1, 0.2), seed=3, linear=0.05))
    FS.generate_model_hashes()
    data = (pimc.BoxDataPM if pm else pimc.BoxData).from_FileStructure(FS)
    data.samples, data.beads, data.temperature, data.block_size = 1000, 12, 300.0, 100
    data.blocks = data.samples // data.block_size
    data.hash_vib, data.hash_rho = FS.hash_vib, FS.hash_rho
    data.seed = 242351
    data.preprocess()
    result = (pimc.BoxResultPM if pm else pimc.BoxResult)(data=data)
    result.path_root, result.id_job = FS.path_rho_results, 3
    (pimc.block_compute_pm if pm else pimc.block_compute)(data, result)
    path = join(FS.path_rho_results, "P12_T300.00_J3_data_points.npz")
    assert isfile(path)
    loaded = type(result)()
    loaded.load_multiple_results([path])
    assert loaded.samples == 1000 and np.array_equal(loaded.scaled_g, result.scaled_g)
    assert not np.isnan(result.scaled_rho).any()
    # per-sample parity with the oracle on the coordinates of the first block (per-block sampler API)
    view = slice(0, data.block_size)
    data.draw_sample(view)
    data.transform_sampled_coordinates(view)
    R = np.ascontiguousarray(data.qTensor[:, 0])
    tab = orc.precompute(orc.load_vibronic_json(FS.path_vib_model), orc.load_sampling_json(FS.path_rho_model), 12, 300.0)
    want = oracle_eval(tab, R, pm=pm)
    assert rel_err(result.scaled_rho[view], want[0]) < RTOL and rel_err(result.scaled_g[view], want[1]) < RTOL
    if pm:
        assert rel_err(result.scaled_gofr_plus[view], want[2]) < RTOL
        assert rel_err(result.scaled_gofr_minus[view], want[3]) < RTOL
    assert result.block_sums.shape == (10, _cabi.NSUMS)
    assert np.isclose(result.block_sums[:, 0].sum(), (result.scaled_g / result.scaled_rho).sum(), rtol=1e-12)
    data.release()


# ------------------------------------------------------------------ "next" rows: statistics, analytic data, sweep
def test_device_statistics_match_the_reference(cuda):
    """pbx_stats_* against the reference's basic_jackknife_analysis output (golden/stats_kat.npz)"""
    from os.path import join
    from conftest import GOLDEN
    from pibronic_b200 import pimc, stats
    kat = np.load(join(GOLDEN, "stats_kat.npz"))
    res = pimc.BoxResultPM(X=3000)
    res.scaled_rho[:], res.scaled_g[:] = kat["s_rho"], kat["s_g"]
    res.scaled_gofr_plus[:], res.scaled_gofr_minus[:] = kat["s_gP"], kat["s_gM"]
    analytic_data = {"E": float(kat["E_sampling"]), "Cv": float(kat["Cv_sampling"])}
    got = stats.basic_jackknife_analysis(float(kat["T"]), res, analytic_data)
    want = dict(zip((str(k) for k in kat["keys"]), kat["values"]))
    assert set(got) == set(want)
    for key in want:
        tol = 1e-7 if key.startswith("jk_") else 1e-11     # X*E - (X-1)*mean(f) cancels log10(X) digits in both
        assert np.isclose(got[key], want[key], rtol=tol, atol=1e-300), f"{key}: {got[key]} vs {want[key]}"
    basic = stats.basic_statistical_analysis(float(kat["T"]), res, analytic_data)
    want_basic = dict(zip((str(k) for k in kat["basic_keys"]), kat["basic_values"]))
    assert set(basic) == set(want_basic)
    for key in want_basic:
        assert np.isclose(basic[key], want_basic[key], rtol=1e-11, atol=1e-300), key


@pytest.mark.parametrize("n", [2, 3, 31, 33, 591, 592, 593, 4097, 100_001])
def test_device_statistics_ragged_sizes(cuda, n):
    """sample counts around the warp, the grid (592 CTAs) and block boundaries of the statistics kernels, with a padded
    row stride, against the numpy restatement of the reference's jackknife; block sums with a ragged last block"""
    from oracle import pimc_oracle as orc
    from oracle import stats_oracle
    rng = np.random.default_rng(n)
    store = np.full((4, n + 5), np.nan)
    out = store[:, :n]
    out[0] = rng.uniform(0.5, 1.5, n)
    out[1] = out[0] * rng.uniform(2.0, 3.0, n)
    out[2] = out[1] * (1 + 1e-4 * rng.normal(size=n))
    out[3] = out[1] * (1 - 1e-4 * rng.normal(size=n))
    T = 300.0
    got = _cabi.stats_arrays_host(out, orc.beta_of(T), orc.DELTA_BETA)
    want = stats_oracle.basic_jackknife_analysis(T, *out)
    for key, value in want.items():
        if n == 2 and key.endswith("error") and key.startswith("jk"):
            continue                                  # two samples: the spread of two leave-one-out values is all rounding
        tol = 1e-5 if key.startswith("jk_") else 1e-9
        assert np.isclose(got[key], value, rtol=tol, atol=1e-12 * max(1.0, abs(value))), f"{key}: {got[key]} vs {value}"
    with cuda.cuda.device(0):
        dev = cuda.from_numpy(np.ascontiguousarray(out)).cuda()
        assert _cabi.stats_arrays_dev(dev, orc.beta_of(T), orc.DELTA_BETA) == got


def test_device_statistics_at_scale_match_oracle(cuda):
    """1e6 samples straight from the device-resident results of the last run vs the numpy restatement"""
    from oracle import stats_oracle
    from conftest import GoldenCase
    case = GoldenCase("quad_3x4")
    plan = case.plan()
    X = 1_000_000
    out = plan.sample_eval_host(5, 0, X)
    got = plan.stats_last()
    want = stats_oracle.basic_jackknife_analysis(case.T, *out)
    for key, value in want.items():
        tol = 2e-6 if key in ("jk_E", "jk_Cv") else 1e-9
        assert np.isclose(got[key], value, rtol=tol, atol=1e-300), f"{key}: {got[key]} vs {value}"
    with cuda.cuda.device(0):
        dev = cuda.from_numpy(out).cuda()
        again = plan.stats(dev)
    assert again == got
    plan.close()


def test_sweep_statistics_and_thermo_files(cuda, tmp_path):
    """temperature x bead sweep -> .npz shards -> analytic_results.json -> jackknife -> *_thermo files"""
    import json
    from os.path import isfile
    from pibronic_b200 import file_structure, stats, sweep, synthetic
    FS = file_structure.FileStructure(tmp_path, 0, 0)
    synthetic.write_data_set(FS, synthetic.coupled_model(2, 2, (0.02, 0.04), (0.1, 0.2), seed=3, linear=0.05))
    params = {"temperature_list": [250.0, 300.0], "bead_list": [8, 12], "number_of_samples": 20000,
              "block_size": 1000, "seed": 11}
    results = sweep.run_sweep(FS, params)
    assert set(results) == {(8, 250.0), (8, 300.0), (12, 250.0), (12, 300.0)}
    thermo = stats.jackknife_analysis_of_pimc(FS, method="basic")
    assert set(thermo) == set(results)
    for (P, T), out in thermo.items():
        path = FS.template_jackknife.format(P=P, T=T, X=20000)
        assert isfile(path)
        with open(path) as fh:
            stored = json.load(fh)
        assert stored["hash_vib"] == FS.hash_vib and stored["Z"] == out["Z"]
        assert set(stored) == {"Z", "Z error", "E", "E error", "Cv", "Cv error", "jk_E", "jk_E error", "jk_Cv",
                               "jk_Cv error", "hash_vib", "hash_rho"}
        assert abs(stored["jk_E"] - stored["E"]) < 5 * stored["jk_E error"] + 1e-12
        assert stored["Z"] > 0 and stored["Z error"] < 0.05 * stored["Z"]
    # colder -> lower energy
    assert thermo[(12, 250.0)]["E"] < thermo[(12, 300.0)]["E"]
    # numbers, not only files: (i) per-sample values of one point against the oracle on the co-ordinates the device
    # sampler reports for the sweep's seed; (ii) the thermo file against the numpy restatement of the reference's
    # basic_jackknife_analysis (oracle/stats_oracle.py) on the arrays of the .npz + the analytic sampling-model data
    from oracle import pimc_oracle as orc, stats_oracle
    from pibronic_b200 import pimc, postprocessing as pp
    P, T = 8, 250.0
    data = pimc.BoxDataPM.from_FileStructure(FS)
    data.samples, data.beads, data.temperature, data.block_size, data.blocks = 20000, P, T, 1000, 20
    data.seed = 11
    data.preprocess()
    data.draw_sample(slice(0, 200))
    data.transform_sampled_coordinates(slice(0, 200))
    R = np.ascontiguousarray(data.qTensor[:200, 0])
    data.release()
    tab = orc.precompute(orc.load_vibronic_json(FS.path_vib_model), orc.load_sampling_json(FS.path_rho_model), P, T)
    want = oracle_eval(tab, R)
    res = results[(P, T)]
    got = np.stack([res.scaled_rho[:200], res.scaled_g[:200], res.scaled_gofr_plus[:200], res.scaled_gofr_minus[:200]])
    assert rel_err(got, want) < RTOL
    rho_data = {}
    pp.load_analytic_data(FS, T, rho_data)
    numpy_stats = stats_oracle.basic_jackknife_analysis(T, res.scaled_rho, res.scaled_g, res.scaled_gofr_plus, res.scaled_gofr_minus,
                                                        rho_data["E"], rho_data["Cv"])
    for key in ("Z", "Z error", "E", "Cv", "jk_E", "jk_E error", "jk_Cv", "jk_Cv error"):
        assert np.isclose(thermo[(P, T)][key], numpy_stats[key], rtol=1e-7, atol=1e-14), key
    with pytest.raises(Exception, match="Invalid value for parameter method"):
        stats.statistical_analysis_of_pimc(FS, method="alpha")


def test_block_compute_gR_is_unscaled(cuda, tmp_path):
    """block_compute_gR (pimc.py:1216-1247): g without the S scaling, on the same Philox samples"""
    from oracle import pimc_oracle as orc
    from pibronic_b200 import file_structure, pimc, synthetic
    FS = file_structure.FileStructure(tmp_path, 0, 0)
    synthetic.write_data_set(FS, synthetic.coupled_model(2, 2, (0.02, 0.04), (0.1, 0.2), seed=3, linear=0.05))
    FS.generate_model_hashes()
    data = pimc.BoxData.from_FileStructure(FS)
    data.samples, data.beads, data.temperature, data.block_size, data.blocks = 200, 12, 300.0, 100, 2
    data.hash_vib, data.hash_rho, data.seed = FS.hash_vib, FS.hash_rho, 7
    data.preprocess()
    result = pimc.BoxResult(data=data)
    result.path_root, result.id_job = FS.path_rho_results, 0
    pimc.block_compute_gR(data, result)
    data.draw_sample(slice(0, 100))
    data.transform_sampled_coordinates(slice(0, 100))
    R = np.ascontiguousarray(data.qTensor[:, 0])
    tab = orc.precompute(orc.load_vibronic_json(FS.path_vib_model), orc.load_sampling_json(FS.path_rho_model), 12, 300.0)
    _, g = orc.estimate_block(tab, R, pm=False, faithful=False, scale=False)
    assert rel_err(result.scaled_g[:100], g) < RTOL
    assert np.isnan(result.scaled_rho).all()
    data.release()


def _small_job(tmp_path, pm=False, samples=200, block=100, beads=12):
    from pibronic_b200 import file_structure, pimc, synthetic
    FS = file_structure.FileStructure(tmp_path, 0, 0)
    synthetic.write_data_set(FS, synthetic.coupled_model(2, 2, (0.02, 0.04), (0.1, 0.2), seed=3, linear=0.05))
    FS.generate_model_hashes()
    data = (pimc.BoxDataPM if pm else pimc.BoxData).from_FileStructure(FS)
    data.samples, data.beads, data.temperature, data.block_size, data.blocks = samples, beads, 300.0, block, samples // block
    data.hash_vib, data.hash_rho, data.seed = FS.hash_vib, FS.hash_rho, 7
    data.preprocess()
    result = (pimc.BoxResultPM if pm else pimc.BoxResult)(data=data)
    result.path_root, result.id_job = FS.path_rho_results, 0
    return FS, data, result


def test_block_compute_rhoR_from_input_samples(cuda, tmp_path):
    """pimc.py:1273-1302: rho(R), not scaled, on caller supplied co-ordinates -- the oracle's denominator on the same R;
    with quirk_double_shift the reference's own arithmetic (q = R - d_vib[a] - d_rho[a])"""
    from oracle import pimc_oracle as orc
    from pibronic_b200 import pimc
    FS, data, result = _small_job(tmp_path)
    rng = np.random.default_rng(5)
    R = rng.normal(scale=0.15, size=(200, 2, 12))
    pimc.block_compute_rhoR_from_input_samples(data, result, R)
    tab = orc.precompute(orc.load_vibronic_json(FS.path_vib_model), orc.load_sampling_json(FS.path_rho_model), 12, 300.0)
    want = orc.denominator(orc.o_factors(R, tab.d_rho, tab.rho))
    assert rel_err(result.scaled_rho, want) < RTOL and np.isnan(result.scaled_g).all()
    pimc.block_compute_rhoR_from_input_samples(data, result, R, quirk_double_shift=True)
    quirk = orc.denominator(orc.o_factors(R, tab.d_rho + tab.d_vib, tab.rho))
    assert rel_err(result.scaled_rho, quirk) < RTOL
    assert rel_err(quirk, want) > 1e-3          # the two conventions really differ on this model
    data.release()


def test_block_compute_gR_from_raw_samples(cuda, tmp_path):
    """pimc.py:1305-1338 + save_gR_with_samples 1250-1270: unscaled g on uniform random co-ordinates, saved with them"""
    from os.path import join
    from oracle import pimc_oracle as orc
    from pibronic_b200 import pimc
    FS, data, result = _small_job(tmp_path)
    pimc.block_compute_gR_from_raw_samples(data, result)
    g_file = np.load(join(FS.path_rho_results, "P12_T300.00_J0_training_data_g_output.npz"))
    r_file = np.load(join(FS.path_rho_results, "P12_T300.00_J0_training_data_input.npz"))
    assert int(g_file["number_of_samples"]) == 200 and int(r_file["number_of_samples"]) == 200
    R = r_file["input_R_values"]
    assert R.shape == (200, 2, 12) and R.min() >= 0.0 and R.max() < 1.0
    tab = orc.precompute(orc.load_vibronic_json(FS.path_vib_model), orc.load_sampling_json(FS.path_rho_model), 12, 300.0)
    _, g = orc.estimate_block(tab, R, pm=False, faithful=False, scale=False)
    assert rel_err(g_file["g"], g) < RTOL and np.array_equal(g_file["g"], result.scaled_g)
    # same seed -> same co-ordinates, whatever the block size
    (tmp_path / "again").mkdir()
    FS2, data2, result2 = _small_job(tmp_path / "again", block=50)
    pimc.block_compute_gR_from_raw_samples(data2, result2)
    assert np.array_equal(result2.scaled_g[:50], result.scaled_g[:50])
    data.release()
    data2.release()


def test_simple_and_plus_minus_wrappers(cuda, tmp_path):
    """pimc.py:1465-1535: the two convenience drivers (100 samples, one block) write the reference's result files"""
    from os.path import isfile, join
    from pibronic_b200 import file_structure, pimc, synthetic
    FS = file_structure.FileStructure(tmp_path, 4, 0)
    synthetic.write_data_set(FS, synthetic.coupled_model(2, 2, (0.02, 0.04), (0.1, 0.2), seed=3, linear=0.05))
    res = pimc.simple_wrapper(4, path_root=str(tmp_path), beads=16)
    assert isfile(join(FS.path_rho_results, "P16_T300.00_J0_data_points.npz"))
    assert res.samples == 100 and np.all(np.isfinite(res.scaled_g / res.scaled_rho))
    FS3 = file_structure.FileStructure(tmp_path, 5, 0)
    synthetic.write_data_set(FS3, synthetic.coupled_model(3, 6, (0.02, 0.04), (0.1, 0.2), seed=4, linear=0.05))
    res = pimc.plus_minus_wrapper(5, path_root=str(tmp_path))
    loaded = pimc.BoxResultPM()
    loaded.load_multiple_results([join(FS3.path_rho_results, "P20_T300.00_J0_data_points.npz")])
    assert loaded.samples == 100 and np.array_equal(loaded.scaled_gofr_plus, res.scaled_gofr_plus)


def test_run_time_compiled_shape_uses_the_register_resident_kernel(cuda):
    """a (5, 3, 5) model is not in csrc/shapes.def: by default it runs on the fused tensor-core kernel; with jit=True
    pibronic_b200.jit compiles pbx_fast_inst.cu for the shape (nvcc, cached under pibronic_b200/_jit/), registers it and the
    plan runs the warp-specialised register-resident kernel.  Both against the oracle, and against each other."""
    import shutil
    from oracle import pimc_oracle as orc
    from pibronic_b200 import jit, synthetic
    from pibronic_b200.model_io import VMK
    if shutil.which("nvcc") is None:
        pytest.skip("no nvcc on this box")
    model = synthetic.coupled_model(5, 3, (0.05, 0.2), (1.0, 1.4), seed=535, quadratic=0.08)
    rho = synthetic.diagonal_of(model)
    P, T = 24, 300.0
    args = (model[VMK.E], model[VMK.w], model[VMK.G1], model[VMK.G2], rho[VMK.E], rho[VMK.w], rho[VMK.G1], P, orc.beta_of(T), orc.DELTA_BETA)
    default = _cabi.Plan(*args, flags=_cabi.FLAG_PM, device=0)
    assert default.kernel_path == _cabi.PATH_FUSED_DMMA
    assert jit.eligible(5, 3, 5) and not jit.eligible(7, 12, 7)
    compiled = _cabi.Plan(*args, flags=_cabi.FLAG_PM, device=0, jit=True)
    assert compiled.kernel_path == _cabi.PATH_REGISTER and compiled.is_fast
    vib_d = dict(A=5, N=3, E=model[VMK.E], w=model[VMK.w], L=model[VMK.G1], Q=model[VMK.G2])
    rho_d = dict(A=5, N=3, E=rho[VMK.E], w=rho[VMK.w], L=rho[VMK.G1])
    tab = orc.precompute(vib_d, rho_d, P, T)
    n = 700
    a, b = default.sample_eval_host(3, 50, n), compiled.sample_eval_host(3, 50, n)
    R, _ = drawn_coords(cuda, compiled, 3, 50, n)
    want = oracle_eval(tab, R)
    assert rel_err(a, want) < RTOL and rel_err(b, want) < RTOL
    assert rel_err(compiled.eval_coords_host(R), want) < RTOL
    # the Jacobi cross-check variant is not part of a run-time compiled shape: such a plan falls back to the blocked kernels
    jac = _cabi.Plan(*args, flags=_cabi.FLAG_PM | _cabi.FLAG_EIG_JACOBI, device=0)
    assert not jac.is_fast and rel_err(jac.eval_coords_host(R), want) < RTOL
    for plan in (default, compiled, jac):
        plan.close()


def test_device_math(cuda):
    """the branch-free log / sqrt / exp / sincos of pbx_device.cuh against numpy, in units of ulp"""
    rng = np.random.default_rng(0)
    n = 1 << 20
    cases = {
        0: (np.concatenate([rng.random(n), 2.0 ** -rng.integers(1, 53, 4096), [1.0, 2.0 ** -53, 0.5, np.sqrt(0.5), 0.70710678118654757,
                            1 - 2.0 ** -53, 1 - 2.0 ** -30, 0.75, np.nextafter(0.75, 0), 0.375, np.nextafter(0.5, 0)]]), np.log),
        1: (np.concatenate([rng.random(n) * 80, rng.random(4096) * 1e-6, [75.0, 1.0, 4.0]]), np.sqrt),
        2: (np.concatenate([-rng.random(n) * 60, -rng.random(4096) * 700, [0.0, -1e-300, -708.0, -745.0, -1e4, -np.inf, 1e-3]]), np.exp),
        3: (np.concatenate([rng.random(n), [0.0, 0.25, 0.5, 0.75, 0.125, 1 - 2.0 ** -53]]), lambda u: np.sin(2 * np.pi * u)),
        4: (np.concatenate([rng.random(n), [0.0, 0.25, 0.5, 0.75, 0.125, 1 - 2.0 ** -53]]), lambda u: np.cos(2 * np.pi * u)),
    }
    for kind, (x, fn) in cases.items():
        xd = cuda.from_numpy(np.ascontiguousarray(x)).cuda()
        out = cuda.empty_like(xd)
        _cabi.math_probe(kind, xd, out)
        got, want = out.cpu().numpy(), fn(x)
        if kind in (3, 4):      # absolute: the argument reduction of numpy's 2*pi*u costs it an ulp of the angle
            assert np.max(np.abs(got - want)) < 2e-15, kind
        elif kind == 2:
            tiny = want < 1e-300
            assert np.all(got[tiny] <= 1e-300) and np.all(got[tiny] >= 0)
            assert np.max(np.abs(got[~tiny] / want[~tiny] - 1)) < 1e-15, kind
        else:
            assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300)) < 1e-15, kind


@pytest.mark.parametrize("delta_beta", [0.5, 5.0])
def test_large_delta_beta_takes_the_safe_path(cuda, delta_beta):
    """delta_beta far above the default 2e-4: the tau+- factors are no longer a small correction of the tau ones,
    the fast kernel must notice and redo those samples with full exponentials (same answer as the oracle)"""
    from conftest import GoldenCase
    from oracle import pimc_oracle as orc
    for name in ("c2_4x6", "quad_3x4", "c5_altrho"):
        case = GoldenCase(name)
        beta = orc.beta_of(case.T)
        plan = _cabi.Plan(case.vib["E"], case.vib["w"], case.vib["L"], case.vib["Q"], case.rho["E"], case.rho["w"],
                          case.rho["L"], case.P, beta, delta_beta, flags=_cabi.FLAG_PM, device=0)
        assert plan.is_fast
        tab = orc.precompute(case.vib, case.rho, case.P, case.T, delta_beta=delta_beta)
        got = plan.eval_coords_host(case.R)
        want = oracle_eval(tab, case.R)
        assert rel_err(got, want) < RTOL, name
        n = 200
        fused = plan.sample_eval_host(3, 0, n)
        R, _ = drawn_coords(cuda, plan, 3, 0, n)
        assert rel_err(fused, oracle_eval(tab, R)) < RTOL, name
        plan.close()


# ------------------------------------------------------------------ the consistent estimator (PBX_FLAG_M_TAU_PM)
# not part of the reference's path: compiled only with PBX_WITH_MTAU=1 (python -m pibronic_b200.build), skipped otherwise
needs_mtau = pytest.mark.skipif(not _cabi.has_feature(_cabi.FEATURE_MTAU), reason="library built without PBX_WITH_MTAU=1")


@needs_mtau
def test_consistent_estimator_matches_oracle(cuda, case):
    """g+- with exp(-tau+- V) -- not the reference's estimator (it keeps exp(-tau V), pimc.py:1183) -- against the oracle
    with the same option, on the reference's coordinates and through the fused call; rho and g do not change"""
    from oracle import pimc_oracle as orc
    base = _cabi.FLAG_PM | _cabi.QUIRK_RHO_TRUNC
    ref_plan = case.plan(base)
    if not ref_plan.is_fast:
        ref_plan.close()
        with pytest.raises(_cabi.PbxError):
            case.plan(base | _cabi.FLAG_M_TAU_PM)
        return
    plan = case.plan(base | _cabi.FLAG_M_TAU_PM)
    tab = case.oracle_tables(rho_trunc=True)
    want = np.stack(orc.estimate_block(tab, case.R, pm=True, faithful=False, m_tau_pm=True))
    got = plan.eval_coords_host(case.R)
    assert rel_err(got, want) < RTOL
    quirk = ref_plan.eval_coords_host(case.R)
    coupled = max(np.abs(tab.E_off).max(), np.abs(tab.L_off).max(), np.abs(tab.Q).max()) > 0    # else V = 0 and M = 1
    assert np.array_equal(got[:2], quirk[:2]) and np.array_equal(got[2:], quirk[2:]) != coupled
    n = 300
    fused = plan.sample_eval_host(17, 3, n)
    Rd, _ = drawn_coords(cuda, plan, 17, 3, n)
    assert rel_err(fused, np.stack(orc.estimate_block(tab, Rd, pm=True, faithful=False, m_tau_pm=True))) < RTOL
    one_role = case.plan(base | _cabi.FLAG_M_TAU_PM | _cabi.FLAG_NO_WARPSPEC)
    assert np.array_equal(fused, one_role.sample_eval_host(17, 3, n))
    for p in (plan, ref_plan, one_role):
        p.close()


@needs_mtau
def test_consistent_estimator_with_a_large_delta_beta(cuda):
    """delta_beta so large that exp(+-kappa X) is no small correction: the flagged samples take the full-exponential path"""
    from conftest import GoldenCase
    from oracle import pimc_oracle as orc
    case = GoldenCase("c2_4x6")
    for delta_beta in (0.5, 5.0):
        plan = _cabi.Plan(case.vib["E"], case.vib["w"], case.vib["L"], case.vib["Q"], case.rho["E"], case.rho["w"],
                          case.rho["L"], case.P, orc.beta_of(case.T), delta_beta, flags=_cabi.FLAG_PM | _cabi.FLAG_M_TAU_PM, device=0)
        tab = orc.precompute(case.vib, case.rho, case.P, case.T, delta_beta=delta_beta)
        n = 200
        fused = plan.sample_eval_host(3, 0, n)
        Rd, _ = drawn_coords(cuda, plan, 3, 0, n)
        assert rel_err(fused, np.stack(orc.estimate_block(tab, Rd, pm=True, faithful=False, m_tau_pm=True))) < RTOL
        assert rel_err(plan.eval_coords_host(case.R),
                       np.stack(orc.estimate_block(tab, case.R, pm=True, faithful=False, m_tau_pm=True))) < RTOL
        plan.close()


@needs_mtau
def test_pimc_reproduces_the_sum_over_states_thermodynamics(cuda):
    """End to end against EXACT numbers: the reference's test model data_set_1 sampled from its rho_1, 2e7 samples of 64
    beads (40 ms), vs the sum-over-states Z, E, Cv its Julia dependency computed (tests/golden/sos/).
    Z = Z_rho <g/rho> is the reference's estimator.  E and Cv agree only with the consistent estimator (PBX_FLAG_M_TAU_PM,
    nothing added); the reference's (exp(-tau V) for g+-, sampling-model E and Cv added) misses both by a wide margin."""
    import json
    from os.path import join
    from conftest import GOLDEN
    from oracle import pimc_oracle as orc
    from pibronic_b200 import analytic, constants
    sos_dir = join(GOLDEN, "sos")
    vib = orc.load_vibronic_json(join(sos_dir, "coupled_model.json"))
    rho = orc.load_sampling_json(join(sos_dir, "sampling_model.json"))
    with open(join(sos_dir, "sos_B80.json")) as fh:
        sos = {k: v[0] for k, v in json.load(fh).items()}
    T, P, X = 300.0, 64, 20_000_000
    # the closed-form sampling-model numbers (analytic.py) are the Julia ones
    tilde = rho["E"] - 0.5 * (rho["L"] ** 2 / rho["w"][:, None]).sum(axis=0)
    mine = analytic.thermodynamics(tilde, rho["w"], constants.beta(T))
    # (to 1e-6: the Julia side used beta = 38.68174037, pibronic/constants.py gives 38.68174020)
    assert np.isclose(mine["Z_sampling"], sos["Z_sampling"], rtol=1e-6) and np.isclose(mine["E_sampling"], sos["E_sampling"], rtol=1e-6)
    out = cuda.empty((4, X), dtype=cuda.float64, device="cuda")
    results = {}
    for name, flags in (("consistent", _cabi.FLAG_PM | _cabi.FLAG_M_TAU_PM), ("reference", _cabi.FLAG_PM)):
        plan = _cabi.Plan(vib["E"], vib["w"], vib["L"], vib["Q"], rho["E"], rho["w"], rho["L"], P, constants.beta(T),
                          constants.delta_beta, flags=flags, device=0)
        plan.sample_eval(2090, 0, X, out)
        results[name] = plan.stats(out, X)
        plan.close()
    st = results["consistent"]
    Z, dZ = st["Z"] * mine["Z_sampling"], st["Z error"] * mine["Z_sampling"]
    # Trotter error at P = 64 is +0.65 % in Z, -4.5e-4 in E (tools/sos_check.py shows the convergence with P)
    assert abs(Z - sos["Z_coupled"] * 1.0065) < 5 * dZ + 2e-3 * sos["Z_coupled"]
    assert np.array_equal(results["reference"]["Z"], st["Z"])                         # rho and g do not depend on the flag
    assert abs(st["jk_E"] - sos["E_coupled"]) < 5 * st["jk_E error"] + 1e-3
    assert abs(st["jk_Cv"] - sos["Cv_coupled"]) < 5 * st["jk_Cv error"] + 0.05 * sos["Cv_coupled"]
    ref = results["reference"]
    assert abs(ref["jk_E"] + sos["E_sampling"] - sos["E_coupled"]) > 0.4               # +0.104 instead of -0.423
    assert ref["jk_Cv"] + sos["Cv_sampling"] > 10 * sos["Cv_coupled"]
