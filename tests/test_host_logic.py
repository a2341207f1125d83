"""Host side of the drop-in (no GPU): the facade's set-up phase against the reference's known answers,
the file formats, the C ABI library (symbols + host-only plans) and the sampler algebra."""
import ctypes
import json
import os
import re
from os.path import join

import numpy as np
import pytest

from conftest import GOLDEN, REPO, GoldenCase
from oracle import philox_sampler as ps
from oracle import pimc_oracle as orc
from pibronic_b200 import _cabi, constants, file_structure, pimc, synthetic
from pibronic_b200 import model_io as vIO
from pibronic_b200.model_io import VMK

RTOL, ATOL = 1e-05, 1e-08


# ------------------------------------------------------------------ facade set-up vs the reference's KAT
@pytest.fixture(scope="module")
def kat_data(tmp_path_factory):
    """mirrors the fixtures of the reference's tests/pimc/test_pimc_explicit_example.py:45-84"""
    kat = np.load(join(GOLDEN, "explicit_kat.npz"))
    root = tmp_path_factory.mktemp("root")
    FS = file_structure.FileStructure(root, id_data=0, id_rho=0)
    for name, path in (("coupled_model", FS.path_vib_model), ("sampling_model", FS.path_rho_model)):
        with open(path, "w", encoding="UTF8") as fh:
            fh.write(str(kat[name + "_json"]))
    data = pimc.BoxData.from_FileStructure(FS)
    assert data.id_data == 0 and data.id_rho == 0
    assert data.path_vib_model == FS.path_vib_model and data.path_rho_model == FS.path_rho_model
    model = vIO.load_model_from_JSON(FS.path_vib_model)
    assert data.states == model[VMK.number_of_surfaces] and data.modes == model[VMK.number_of_modes]
    data.samples, data.beads, data.temperature = 10, 5, 300.00
    data.block_size = 2
    data.blocks = data.samples // data.block_size
    data.preprocess()
    return kat, FS, data


def test_beta():
    assert np.allclose(constants.beta(300.00), 38.68174020133669, rtol=RTOL, atol=ATOL)
    assert constants.delta_beta == 2.0e-4
    assert np.isclose(constants.extract_T_from_beta(constants.beta(123.0)), 123.0)


def test_surface_weights(kat_data):
    kat, _, data = kat_data
    w = kat["sampling_surface_weights"] / kat["sampling_surface_weights"].sum()
    assert np.allclose(w, data.rho.state_weight, rtol=RTOL, atol=ATOL)


def test_multivariate_normal_distributions(kat_data):
    kat, _, data = kat_data
    assert np.allclose(kat["coupled_means"][..., 0], data.vib.state_shift, rtol=RTOL, atol=ATOL)
    assert np.allclose(kat["sampling_means"][..., 0], data.rho.state_shift, rtol=RTOL, atol=ATOL)
    NEW = np.newaxis
    left = 2. * data.rho.const.cothAN[..., NEW, NEW] * np.eye(5)
    right = data.rho.const.cschAN[..., NEW, NEW] * data.circulant_matrix[NEW, NEW, ...]
    cov = np.linalg.inv(left - right)
    assert np.allclose(kat["sampling_covariance"], cov, rtol=RTOL, atol=ATOL)
    # normal-mode form used by ModelSampling.compute_sampling_constants
    V = data.circulant_eigvects
    for a in range(2):
        for n in range(2):
            assert np.allclose((V / data.rho.inverse_covariance[a, n]) @ V.T, cov[a, n], rtol=RTOL, atol=ATOL)


def test_offsets_due_to_linear_terms(kat_data):
    kat, _, data = kat_data
    assert np.allclose(kat["coupled_Edeltas"], data.vib.delta_weight, rtol=RTOL, atol=ATOL)
    assert np.allclose(kat["sampling_Edeltas"], data.rho.delta_weight, rtol=RTOL, atol=ATOL)
    assert np.allclose(kat["coupled_ds"], data.vib.state_shift, rtol=RTOL, atol=ATOL)
    assert np.allclose(kat["sampling_ds"], data.rho.state_shift, rtol=RTOL, atol=ATOL)
    # the folded terms leave the coupling matrix (pimc.py:208-218)
    assert not np.diag(data.vib.energy).any()
    assert not data.vib.linear[:, np.arange(2), np.arange(2)].any()
    assert data.vib.raw["linear"][:, np.arange(2), np.arange(2)].any()


def test_precomputed_coth_csch(kat_data):
    kat, _, data = kat_data
    assert np.allclose(kat["coupled_cosh"] / kat["coupled_sinh"], data.vib.const.cothAN, rtol=RTOL, atol=ATOL)
    assert np.allclose(kat["sampling_cosh"] / kat["sampling_sinh"], data.rho.const.cothAN, rtol=RTOL, atol=ATOL)
    assert np.allclose(kat["coupled_sinh"] ** -1., data.vib.const.cschAN, rtol=RTOL, atol=ATOL)
    assert np.allclose(kat["sampling_sinh"] ** -1., data.rho.const.cschAN, rtol=RTOL, atol=ATOL)
    assert data.vib.const.omatrix_prefactor.shape == (2, 5, 2)
    assert data.vib.const.omatrix.shape == (2, 5, 2, 2)


def test_facade_tables_match_running_reference(case, tmp_path):
    """BoxDataPM.preprocess() on every golden case against the tables of the reference's BoxDataPM"""
    data = pimc.BoxDataPM()
    data.path_vib_model, data.path_rho_model = case.path_vib, case.path_rho
    data.states, data.modes = vIO.extract_dimensions_of_model(path=case.path_vib)
    data.samples, data.beads, data.temperature = 8, case.P, case.T
    data.block_size, data.blocks = 4, 2
    data.preprocess()
    ref = case.ref
    assert np.isclose(data.beta, float(ref["beta"]), rtol=1e-15) and np.isclose(data.tau, float(ref["tau"]), rtol=1e-15)
    assert np.allclose(data.vib.state_shift, ref["vib_shift"], rtol=1e-14, atol=0)
    assert np.allclose(data.vib.delta_weight, ref["vib_delta"], rtol=1e-14, atol=0)
    assert np.allclose(data.vib.const.cothAN[0], ref["vib_coth"], rtol=1e-15)
    assert np.allclose(data.vib.const.omatrix_prefactor[0, 0], ref["vib_pref"], rtol=1e-13)
    assert np.allclose(data.vib.const_plus.omatrix_prefactor[0, 0], ref["vib_pref_plus"], rtol=1e-13)
    assert np.allclose(data.vib.const_minus.omatrix_prefactor[0, 0], ref["vib_pref_minus"], rtol=1e-13)
    assert np.allclose(data.vib.const_plus.cothAN[0], ref["vib_coth_plus"], rtol=1e-15)
    assert np.allclose(data.vib.const_minus.cschAN[0], ref["vib_csch_minus"], rtol=1e-15)
    assert np.allclose(data.rho.state_shift, ref["rho_shift"], rtol=1e-14, atol=0)
    assert np.allclose(data.rho.state_weight, ref["rho_weight"], rtol=1e-13)
    assert np.allclose(data.rho.const.omatrix_prefactor[0, 0], ref["rho_pref"], rtol=1e-13)
    assert np.allclose(data.rho.inverse_covariance[0], ref["inverse_covariance"], rtol=1e-10, atol=1e-12)
    assert np.allclose(data.circulant_eigvals, ref["ring_eigvals"], atol=1e-13)
    assert data.tau_plus == data.vib.tau_plus and data.tau_plus > data.tau > data.tau_minus


def test_preprocess_requires_three_beads(case):
    data = pimc.BoxData()
    data.path_vib_model, data.path_rho_model = case.path_vib, case.path_rho
    data.states, data.modes = vIO.extract_dimensions_of_model(path=case.path_vib)
    data.samples, data.beads, data.temperature, data.block_size, data.blocks = 4, 2, 300.0, 2, 2
    with pytest.raises(AssertionError, match="3 or more beads"):
        data.preprocess()


def test_block_compute_pm_type_checks():
    with pytest.raises(AssertionError, match="incorrect object type"):
        pimc.block_compute_pm(pimc.BoxData(), pimc.BoxResultPM(X=2))
    with pytest.raises(AssertionError, match="incorrect object type"):
        pimc.block_compute_pm(pimc.BoxDataPM(), pimc.BoxResult(X=2))


# ------------------------------------------------------------------ execution parameters as JSON
def test_json_round_trip(tmp_path):
    FS = file_structure.FileStructure(tmp_path, 3, 1)
    synthetic.write_data_set(FS, synthetic.coupled_model(2, 2, (0.01, 0.02), (0.1, 0.2), seed=1))
    data = pimc.BoxDataPM()
    data.samples, data.blocks, data.states, data.beads, data.modes = 100, 10, 2, 12, 2
    data.temperature, data.block_size, data.id_data, data.id_rho = 300.0, 10, 3, 1
    data.beta, data.tau = constants.beta(300.0), constants.beta(300.0) / 12
    text = data.encode_self()
    assert "," not in text and ";" in text             # survives a SLURM env variable (pimc.py:464-465)
    params = json.loads(text.replace(";", ","))
    assert params["number_of_beads"] == 12 and params["delta_beta"] == constants.delta_beta
    params["path_root"] = str(tmp_path)
    clone = pimc.BoxDataPM.from_json_string(pimc.BoxData.json_serialize(params))
    for name in ("samples", "blocks", "states", "beads", "modes", "temperature", "block_size", "id_data", "id_rho"):
        assert getattr(clone, name) == getattr(data, name)
    assert clone.path_vib_model == FS.path_vib_model and clone.path_rho_model == FS.path_rho_model
    assert clone.hash_vib == vIO.create_model_hash(FS) and clone.hash_rho == vIO.create_diagonal_model_hash(FS)
    assert len(clone.hash_vib) == 128


# ------------------------------------------------------------------ model files
def test_model_json_round_trip(tmp_path):
    model = synthetic.model_c2()
    path = join(tmp_path, "m.json")
    vIO.save_model_to_JSON(path, model)
    back = vIO.load_model_from_JSON(path)
    for key in (VMK.E, VMK.w, VMK.G1, VMK.G2):
        assert np.array_equal(back[key], model[key])
    assert back[VMK.A] == 4 and back[VMK.N] == 6
    # zero arrays are omitted on save and come back as zeros through the in-place loader
    model[VMK.G2] = np.zeros((6, 6, 4, 4))
    vIO.save_model_to_JSON(path, model)
    with open(path) as fh:
        assert "quadratic couplings" not in json.load(fh)
    filled = {VMK.N: 6, VMK.A: 4, VMK.E: np.ones((4, 4)), VMK.w: np.ones(6), VMK.G1: np.ones((6, 4, 4)),
              VMK.G2: np.ones((6, 6, 4, 4))}
    vIO.load_model_from_JSON(path, filled)
    assert not filled[VMK.G2].any() and np.array_equal(filled[VMK.G1], model[VMK.G1])
    with pytest.raises(AssertionError, match="incorrect shape"):
        vIO.save_model_to_JSON(path, {VMK.N: 6, VMK.A: 4, VMK.E: np.ones((3, 3))})


def test_reads_the_references_model_files(case):
    """files written by the reference's vIO load identically through ours and through the oracle's loader"""
    ours = vIO.load_model_from_JSON(case.path_vib)
    assert np.array_equal(ours[VMK.E], case.vib["E"]) and np.array_equal(ours[VMK.w], case.vib["w"])
    if VMK.G1 in ours:
        assert np.array_equal(ours[VMK.G1], case.vib["L"])
    if VMK.G2 in ours:
        assert np.array_equal(ours[VMK.G2], case.vib["Q"])
    rho = vIO.load_diagonal_model_from_JSON(case.path_rho)
    assert np.array_equal(rho[VMK.E], case.rho["E"])
    assert vIO.extract_dimensions_of_diagonal_model(path=case.path_rho) == (case.rho["A"], case.rho["N"])


def test_basic_diagonal_model_and_file_structure(tmp_path):
    FS = file_structure.FileStructure(tmp_path, 7, 2)
    assert FS.path_rho_results.endswith("data_set_7/rho_2/results/") and os.path.isdir(FS.path_rho_results)
    assert FS.path_vib_model.endswith("data_set_7/parameters/coupled_model.json")
    assert FS.path_rho_model.endswith("data_set_7/rho_2/parameters/sampling_model.json")
    assert FS.directories_exist()
    model = synthetic.model_c2()
    synthetic.write_data_set(FS, model)
    rho = vIO.load_diagonal_model_from_JSON(FS.path_rho_model)
    assert np.array_equal(rho[VMK.E], np.diag(model[VMK.E]))
    assert np.array_equal(rho[VMK.G1], np.diagonal(model[VMK.G1], axis1=1, axis2=2))
    FS.generate_model_hashes()
    assert FS.valid_vib_hash({"hash_vib": FS.hash_vib}) and not FS.valid_rho_hash({"hash_rho": "x"})
    FS.change_rho(5)
    assert FS.path_rho_model.endswith("rho_5/parameters/sampling_model.json") and os.path.isdir(FS.path_rho_params)
    assert FS.template_pimc.format(P=12, T=300.0, J="*").endswith("rho_5/results/P12_T300.00_J*_data_points.npz")


# ------------------------------------------------------------------ result files
def _filled_result(X, J, root, seed, cls=pimc.BoxResultPM):
    class D:
        beads, temperature, samples, hash_vib, hash_rho = 12, 300.0, X, "hv", "hr"
    res = cls(data=D)
    res.path_root, res.id_job = str(root), J
    rng = np.random.RandomState(seed)
    for name in ("scaled_rho", "scaled_g", "scaled_gofr_plus", "scaled_gofr_minus"):
        if hasattr(res, name):
            getattr(res, name)[:] = rng.uniform(0.5, 1.5, X)
    return res


def test_result_npz_schema_and_merge(tmp_path):
    a, b = _filled_result(6, 0, tmp_path, 1), _filled_result(4, 1, tmp_path, 2)
    assert np.isnan(pimc.BoxResultPM(X=3).scaled_gofr_minus).all()       # NaN initialised (pimc.py:773-776)
    a.save_results(6)
    b.save_results(4)
    pa, pb = join(tmp_path, "P12_T300.00_J0_data_points.npz"), join(tmp_path, "P12_T300.00_J1_data_points.npz")
    assert os.path.isfile(pa) and os.path.isfile(pb)
    with np.load(pa) as f:
        assert sorted(f.keys()) == sorted(["hash_vib", "hash_rho", "number_of_samples", "s_rho", "s_g", "s_gP", "s_gM"])
        assert f["number_of_samples"] == 6 and str(f["hash_vib"]) == "hv" and f["s_gP"].dtype == np.float64
    assert pimc.BoxResultPM.read_number_of_samples(pa) == 6
    one = pimc.BoxResultPM()
    one.load_results(pa)
    assert np.array_equal(one.scaled_gofr_plus, a.scaled_gofr_plus) and one.samples == 6
    merged = pimc.BoxResultPM()
    merged.load_multiple_results([pa, pb])
    assert merged.samples == 10
    assert np.array_equal(merged.scaled_g, np.concatenate([a.scaled_g, b.scaled_g]))
    assert np.array_equal(merged.scaled_gofr_minus, np.concatenate([a.scaled_gofr_minus, b.scaled_gofr_minus]))
    capped = pimc.BoxResultPM()
    capped.load_multiple_results([pa, pb], desired_number_of_samples=8)
    assert capped.samples == 8 and np.array_equal(capped.scaled_rho[6:], b.scaled_rho[:2])
    # a non-PM shard lacks s_gP/s_gM: skipped by the PM loader, as in the reference (pimc.py:998-1001)
    plain = _filled_result(5, 2, tmp_path, 3, cls=pimc.BoxResult)
    plain.save_results(5)
    pc = join(tmp_path, "P12_T300.00_J2_data_points.npz")
    skip = pimc.BoxResultPM()
    skip.load_multiple_results([pa, pc])
    assert skip.samples == 6
    with pytest.raises(AssertionError, match="none of the provided paths were good"):
        pimc.BoxResultPM().load_multiple_results([pc])
    with pytest.raises(AssertionError, match="cannot be empty"):
        pimc.BoxResultPM().load_multiple_results([])
    zero = _filled_result(3, 3, tmp_path, 4)
    zero.scaled_rho[1] = 0.0
    zero.save_results(3)
    with pytest.raises(AssertionError, match="Zeros in the denominator"):
        pimc.BoxResultPM().load_multiple_results([join(tmp_path, "P12_T300.00_J3_data_points.npz")])
    with pytest.raises(AssertionError, match="hash values"):
        bad = _filled_result(2, 4, tmp_path, 5)
        bad.hash_vib = None
        bad.save_results(2)
    with pytest.raises(AssertionError, match="0 samples"):
        pimc.BoxResultPM().compute_path_to_file()


# ------------------------------------------------------------------ the C ABI library without a GPU
def test_library_exports_every_declared_symbol():
    header = open(join(REPO, "include", "pbx.h")).read()
    declared = set(re.findall(r"\b(pbx_[a-z0-9_]+)\s*\(", header))
    declared -= {"pbx_model", "pbx_rho", "pbx_plan", "pbx_status"}
    assert declared == set(_cabi.EXPORTED_SYMBOLS), declared ^ set(_cabi.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(_cabi.library_path())
    for name in declared:
        assert hasattr(lib, name), name
    assert _cabi.lib().pbx_abi_version() == 1


def test_no_cpu_fallback(case):
    """without a device a real plan cannot be made, and a host-only plan cannot launch anything"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    assert _cabi.device_count() == 0
    with pytest.raises(_cabi.PbxError, match="no CPU fallback"):
        case.plan(device=0)
    plan = case.plan(device=-1)
    with pytest.raises(_cabi.PbxError, match="no CPU fallback"):
        plan.eval_coords_host(case.R)
    with pytest.raises(_cabi.PbxError, match="no CPU fallback"):
        plan.sample_eval_host(1, 0, 4)
    plan.close()


def test_plan_rejects_bad_models(case):
    v, r = case.vib, case.rho
    beta = orc.beta_of(300.0)
    with pytest.raises(_cabi.PbxError, match="3 or more beads"):
        _cabi.Plan(v["E"], v["w"], v["L"], v["Q"], r["E"], r["w"], r["L"], 2, beta, 2e-4, device=-1)
    E = v["E"].copy()
    E[0, -1] += 1e-3
    with pytest.raises(_cabi.PbxError, match="not symmetric"):
        _cabi.Plan(E, v["w"], v["L"], v["Q"], r["E"], r["w"], r["L"], 8, beta, 2e-4, device=-1)
    with pytest.raises(_cabi.PbxError, match="positive"):
        _cabi.Plan(v["E"], -v["w"], v["L"], v["Q"], r["E"], r["w"], r["L"], 8, beta, 2e-4, device=-1)
    with pytest.raises(_cabi.PbxError, match="number of modes"):
        _cabi.Plan(v["E"], v["w"], v["L"], v["Q"], r["E"], r["w"][:-1], r["L"][:-1], 8, beta, 2e-4, device=-1)
    # the consistent estimator needs g+- (PBX_FLAG_PM), the default exp(-tau V) builder and a register-resident shape
    for flags in (_cabi.FLAG_M_TAU_PM, _cabi.FLAG_M_TAU_PM | _cabi.FLAG_PM | _cabi.FLAG_EIG_JACOBI,
                  _cabi.FLAG_M_TAU_PM | _cabi.FLAG_PM | _cabi.FLAG_FORCE_GENERIC):
        with pytest.raises(_cabi.PbxError, match="PBX_FLAG_M_TAU_PM"):
            _cabi.Plan(v["E"], v["w"], v["L"], v["Q"], r["E"], r["w"], r["L"], 8, beta, 2e-4, flags=flags, device=-1)


def test_plan_tables_match_oracle(case):
    """the C++ precompute (pbx_tables.hpp) against the oracle's restatement of the reference's set-up"""
    tab = case.oracle_tables()
    plan = case.plan(device=-1)
    A, Ar, N, P = tab.A, tab.Ar, tab.N, tab.P
    assert np.allclose(plan.table("d_vib").reshape(A, N), tab.d_vib, rtol=1e-15, atol=0)
    assert np.allclose(plan.table("d_rho").reshape(Ar, N), tab.d_rho, rtol=1e-15, atol=0)
    assert np.allclose(plan.table("delta_vib"), tab.delta_vib, rtol=1e-14, atol=0)
    assert np.allclose(plan.table("weights"), tab.weights, rtol=1e-12)
    coth, csch = plan.table("coth").reshape(4, N), plan.table("csch").reshape(4, N)
    for row, t in enumerate((tab.vib, tab.vib_plus, tab.vib_minus, tab.rho)):
        assert np.allclose(coth[row], t.coth, rtol=1e-15) and np.allclose(csch[row], t.csch, rtol=1e-15)
    lp = plan.table("logpref").reshape(3, A)
    for row, t in enumerate((tab.vib, tab.vib_plus, tab.vib_minus)):
        assert np.allclose(np.exp(lp[row]), t.prefactor, rtol=1e-12)
    assert np.allclose(np.exp(plan.table("logpref_rho")), tab.rho.prefactor, rtol=1e-12)
    assert np.allclose(plan.table("tau"), [tab.tau, (tab.beta + tab.delta_beta) / P, (tab.beta - tab.delta_beta) / P])
    # packed coupling tables rebuild V(R) exactly like the oracle's einsum
    AA = A * (A + 1) // 2
    e_off, l_off = plan.table("e_off"), plan.table("l_off").reshape(N, AA)
    q_pack = plan.table("q_pack").reshape(N * (N + 1) // 2, AA)
    R = case.R[:2]
    V = orc.coupling_matrices(tab, R)
    pairs = [(n, m) for n in range(N) for m in range(n, N)]
    for i in range(A):
        for j in range(i + 1):
            k = i * (i + 1) // 2 + j
            mine = e_off[k] + np.einsum("n,bnp->bp", l_off[:, k], R)
            for q, (n, m) in enumerate(pairs):
                mine = mine + q_pack[q, k] * R[:, n] * R[:, m]
            assert np.allclose(mine, V[:, :, i, j], rtol=1e-12, atol=1e-13 * (1 + np.abs(V).max()))
    assert plan.is_fast == ((A, N, Ar) in {(2, 2, 2), (2, 2, 4), (2, 2, 8), (2, 3, 2), (3, 3, 3), (3, 4, 3), (3, 6, 3),
                                            (4, 4, 4), (4, 6, 4)})
    plan.close()


# ------------------------------------------------------------------ device sampler, restated on the CPU
def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10"""
    def run(c, k):
        return [int(v) for v in ps.philox4x32_10(*[np.uint32(x) for x in c], k[0], k[1])]
    assert run((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_ring_recurrence_has_the_references_covariance(case):
    """y = W z with the plan's (a,b,e) table: W W^T must equal the covariance the reference samples from,
    V diag(sigma^2) V^T (pimc.py:400-404, 613-621) -- the two samplers draw the same Gaussian"""
    tab = case.oracle_tables()
    plan = case.plan(device=-1)
    P, N = tab.P, tab.N
    samp = plan.table("samp").reshape(P, N, 3)
    for n in range(N):
        W = np.zeros((P, P))
        for j in range(P):
            W[j, j] += samp[j, n, 0]
            if j >= 1:
                W[j] += samp[j, n, 1] * W[j - 1]
            if j >= 2:
                W[j] += samp[j, n, 2] * W[0]
        cov_ref = (tab.ring_eigvecs * tab.sigma[n][None, :] ** 2) @ tab.ring_eigvecs.T
        assert np.max(np.abs(W @ W.T - cov_ref)) < 1e-9 * np.max(np.abs(cov_ref))
        a, b, e = ps.ring_recurrence_dense(2 * tab.rho.coth[n], tab.rho.csch[n], P)
        assert np.allclose(samp[:, n, 0], a, rtol=1e-9) and np.allclose(samp[:, n, 1], b, rtol=1e-9, atol=1e-14)
        assert np.allclose(samp[:, n, 2], e, rtol=1e-7, atol=1e-13)
    plan.close()


def test_cpu_restatement_of_device_sampler_statistics():
    """moments of the Philox/Box-Muller normals, mixture frequencies and the bead covariance"""
    z = ps.standard_normals(seed=99, first=0, n=20000, N=3, P=8).ravel()
    assert abs(z.mean()) < 4 / np.sqrt(z.size) and abs(z.var() - 1) < 4 * np.sqrt(2 / z.size)
    assert abs((z ** 4).mean() - 3) < 0.1 and np.abs(z).max() < 7
    c = GoldenCase("jt_rho4")
    tab = c.oracle_tables()
    plan = c.plan(device=-1)
    samp = plan.table("samp").reshape(tab.P, tab.N, 3)
    wcum = np.cumsum(tab.weights)
    X = 40000
    R, src = ps.sample_coords(samp, wcum, tab.d_rho, seed=7, first=1000, n=X)
    freq = np.bincount(src, minlength=tab.Ar) / X
    assert np.max(np.abs(freq - tab.weights)) < 4 * np.sqrt(0.25 / X)
    y = R - tab.d_rho[src][:, :, None]
    for n in range(tab.N):
        emp = y[:, n, :].T @ y[:, n, :] / X
        cov = (tab.ring_eigvecs * tab.sigma[n][None, :] ** 2) @ tab.ring_eigvecs.T
        assert np.max(np.abs(emp - cov)) < 6 * np.max(np.abs(cov)) / np.sqrt(X)
    # counter based: a sub-range reproduces the same samples
    R2, src2 = ps.sample_coords(samp, wcum, tab.d_rho, seed=7, first=1010, n=5)
    assert np.array_equal(R2, R[10:15]) and np.array_equal(src2, src[10:15])
    plan.close()


def test_estimator_on_cpu_sampler_agrees_with_reference_sampler():
    """independent draws from the two samplers give the same <g/rho> within the statistical error"""
    c = GoldenCase("quad_3x4")
    tab = c.oracle_tables()
    plan = c.plan(device=-1)
    samp = plan.table("samp").reshape(tab.P, tab.N, 3)
    X = 6000
    R, _ = ps.sample_coords(samp, np.cumsum(tab.weights), tab.d_rho, seed=3, first=0, n=X)
    mine = orc.estimate_block(tab, R, pm=False, faithful=False)
    theirs = orc.run_blocks(tab, X, X, np.random.RandomState(8), pm=False, faithful=False)
    r1, r2 = mine[1] / mine[0], theirs[1] / theirs[0]
    err = np.sqrt(r1.var() / X + r2.var() / X)
    assert abs(r1.mean() - r2.mean()) < 4 * err
    plan.close()


def test_math_tables_header_is_what_the_generator_writes():
    """csrc/pbx_math_tables.h (log / exp lookup tables of the device math) is generated, not edited"""
    pytest.importorskip("mpmath")
    import importlib.util
    from os.path import dirname, join
    root = dirname(dirname(__file__))
    spec = importlib.util.spec_from_file_location("gen_math_tables", join(root, "tools", "gen_math_tables.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    with open(join(root, "pibronic_b200", "csrc", "pbx_math_tables.h")) as fh:
        assert fh.read() == gen.render()


def test_four_product_exp_coefficients_match_taylor():
    """the constants of namespace t12 (pbx_device.cuh) reproduce 1/k! for k = 0..12 when the four-product form is expanded"""
    import re
    from os.path import dirname, join
    from numpy.polynomial import polynomial as npoly
    with open(join(dirname(dirname(__file__)), "pibronic_b200", "csrc", "pbx_device.cuh")) as fh:
        text = fh.read()
    c = {int(k): float.fromhex(v) for k, v in re.findall(r"\bc(\d+) = (0x[0-9a-f.]+p[+-]\d+)", text)}
    assert sorted(c) == list(range(1, 11))
    x = np.array([0.0, 1.0])
    x2, x3 = npoly.polymul(x, x), npoly.polymul(npoly.polymul(x, x), x)
    y0 = npoly.polymul(x3, npoly.polyadd(c[1] * x3, npoly.polyadd(c[2] * x2, c[3] * x)))
    a = npoly.polyadd(npoly.polyadd(y0, c[4] * x3), npoly.polyadd(c[5] * x2, c[6] * x))
    b = npoly.polyadd(y0, npoly.polyadd(c[7] * x3, c[8] * x2))
    total = npoly.polyadd(npoly.polymul(a, b), npoly.polyadd(npoly.polyadd(c[9] * y0, c[10] * x3), npoly.polyadd(0.5 * x2, npoly.polyadd(x, [1.0]))))
    from math import factorial
    want = np.array([1.0 / factorial(k) for k in range(13)])
    assert len(total) == 13 and np.allclose(total, want, rtol=1e-13, atol=0)
