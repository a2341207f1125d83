"""CPU tests of the rows SURVEY.md section 8(f) marks "next": statistics oracle vs the reference's own
output, closed-form analytic results vs the reference's Julia fixtures, result-file discovery."""
import json
from os.path import join

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import stats_oracle
from pibronic_b200 import analytic, constants, file_structure, pimc, postprocessing as pp, synthetic
from pibronic_b200 import model_io as vIO
from pibronic_b200.model_io import VMK


def test_stats_oracle_matches_the_references_basic_jackknife_analysis():
    kat = np.load(join(GOLDEN, "stats_kat.npz"))
    got = stats_oracle.basic_jackknife_analysis(float(kat["T"]), kat["s_rho"], kat["s_g"], kat["s_gP"], kat["s_gM"],
                                                float(kat["E_sampling"]), float(kat["Cv_sampling"]))
    want = dict(zip((str(k) for k in kat["keys"]), kat["values"]))
    assert set(got) == set(want)
    for key in want:
        assert np.isclose(got[key], want[key], rtol=1e-9, atol=1e-300), key


# values of the reference's Julia (VibronicToolkit.jl) output, tests/test_models/data_set_0/rho_0/parameters/sos_B80.json
JULIA_DATA_SET_0 = dict(beta=38.6817403683985, Z=2.40790143798171, Zp=2.40788699664005, Zm=2.40791587954347,
                        E=0.0299876304771386, Cv=0.000178712007135159)


def test_analytic_results_reproduce_the_julia_fixture():
    E = np.array([0.09959721908298894, 0.19959721908298894])
    w = np.array([0.02, 0.04])
    L = np.array([[0.072, 0.072], [0.0, 0.0]])
    tilde, omega = analytic.sampling_model_terms({VMK.A: 2, VMK.N: 2, VMK.E: E, VMK.w: w, VMK.G1: L})
    got = analytic.thermodynamics(tilde, omega, JULIA_DATA_SET_0["beta"])
    assert np.isclose(got["Z_sampling"], JULIA_DATA_SET_0["Z"], rtol=1e-13)
    assert np.isclose(got["Z_sampling+beta"], JULIA_DATA_SET_0["Zp"], rtol=1e-13)
    assert np.isclose(got["Z_sampling-beta"], JULIA_DATA_SET_0["Zm"], rtol=1e-13)
    assert np.isclose(got["E_sampling"], JULIA_DATA_SET_0["E"], rtol=1e-12)
    # Julia's Boltzmann constant differs from pibronic/constants.py by 4.3e-9 (its beta(300 K) shows it)
    assert np.isclose(got["Cv_sampling"], JULIA_DATA_SET_0["Cv"], rtol=1e-8)


def test_analytic_derivatives_are_consistent():
    """E = -dlnZ/dbeta and Cv = kB beta^2 d2lnZ/dbeta2 by finite differences of Z"""
    tilde, omega = np.array([-0.03, 0.07, 0.01]), np.array([0.02, 0.04, 0.11])
    beta, h = 35.0, 1e-3
    lnZ = [np.log(analytic.partition_function(tilde, omega, b)) for b in (beta - h, beta, beta + h)]
    got = analytic.thermodynamics(tilde, omega, beta)
    assert np.isclose(got["E_sampling"], -(lnZ[2] - lnZ[0]) / (2 * h), rtol=1e-6)
    assert np.isclose(got["Cv_sampling"], constants.boltzman * beta ** 2 * (lnZ[2] - 2 * lnZ[1] + lnZ[0]) / h ** 2, rtol=1e-4)


def test_analytic_results_file(tmp_path):
    FS = file_structure.FileStructure(tmp_path, 0, 0)
    synthetic.write_data_set(FS, synthetic.coupled_model(2, 2, (0.02, 0.04), (0.1, 0.2), seed=3))
    for T in (250.0, 300.0):
        analytic.analytic_of_sampling_model(FS, constants.beta(T))
    with open(FS.path_analytic_rho) as fh:
        stored = json.load(fh)
    assert {"250.00", "300.00", "hash_vib", "hash_rho"} <= set(stored)
    FS.generate_model_hashes()
    loaded = {}
    pp.load_analytic_data(FS, 300.0, loaded)
    assert set(loaded) == {"Z", "E", "Cv", "alpha_plus", "alpha_minus"}
    assert loaded["alpha_plus"] != loaded["alpha_minus"] and abs(loaded["alpha_plus"] * loaded["alpha_minus"] - 1) < 1e-6
    assert loaded["Z"] == stored["300.00"]["Z_sampling"] == stored["300.00"]["Z_rho"]
    with pytest.raises(AssertionError, match="no analytical results"):
        pp.load_analytic_data(FS, 123.0, {})
    # a changed sampling model invalidates the stored entries
    model = vIO.load_diagonal_model_from_JSON(FS.path_rho_model)
    model[VMK.E] = model[VMK.E] + 0.01
    vIO.save_diagonal_model_to_JSON(FS.path_rho_model, model)
    FS.generate_model_hashes(force_flag=True)
    analytic.analytic_of_sampling_model(FS, constants.beta(300.0))
    with open(FS.path_analytic_rho) as fh:
        assert "250.00" not in json.load(fh)


def test_result_file_discovery(tmp_path):
    FS = file_structure.FileStructure(tmp_path, 0, 0)
    for P, T, J in ((12, 300.0, 0), (12, 300.0, 1), (20, 300.0, 0), (12, 250.0, 3)):
        class D:
            beads, temperature, samples, hash_vib, hash_rho = P, T, 4, "a", "b"
        res = pimc.BoxResultPM(data=D)
        res.path_root, res.id_job = FS.path_rho_results, J
        for name in ("scaled_rho", "scaled_g", "scaled_gofr_plus", "scaled_gofr_minus"):
            getattr(res, name)[:] = 1.0 + J
        res.save_results(4)
    files = pp.retrive_pimc_file_list(FS)
    assert len(files) == 4
    assert pp.extract_bead_paramater_list(files) == [12, 20]
    assert pp.extract_temperature_paramater_list(files) == [250.0, 300.0]
    merged = pimc.BoxResultPM()
    pp.load_pimc_data(FS, 12, 300.0, merged)
    assert merged.samples == 8 and sorted(set(merged.scaled_g)) == [1.0, 2.0]


# ------------------------------------------------------------------ the reference's own readers consume our files
def _staged_reference():
    import sys
    from os.path import abspath, dirname
    sys.path.insert(0, dirname(dirname(abspath(__file__))))
    from baseline import ref_runner
    if not ref_runner.available():
        pytest.skip("baseline/_ref not staged (python baseline/stage_reference.py in the build container)")
    return ref_runner.import_reference()


def test_reference_loader_and_jackknife_consume_files_written_here(tmp_path):
    """f2: .npz shards written by pibronic_b200.pimc.BoxResultPM.save_results load in the UNMODIFIED reference's
    BoxResultPM.load_multiple_results (pimc.py:975-1040), and its basic_jackknife_analysis (stats.py:271-299) on them gives
    what pibronic_b200's numpy restatement gives"""
    ref_pimc, _ = _staged_reference()
    from pibronic.stats import stats as ref_stats
    from pibronic_b200 import pimc

    class Data:
        samples, beads, temperature, hash_vib, hash_rho = 600, 12, 300.0, "a" * 128, "b" * 128
    rng = np.random.RandomState(4)
    paths, parts = [], []
    for job in range(2):
        res = pimc.BoxResultPM(data=Data)
        res.path_root, res.id_job = str(tmp_path), job
        res.scaled_rho[:] = rng.uniform(0.5, 1.5, 600)
        res.scaled_g[:] = res.scaled_rho * rng.uniform(1.0, 2.0, 600)
        res.scaled_gofr_plus[:] = res.scaled_g * (1 - 3e-3 * rng.uniform(0.9, 1.1, 600))
        res.scaled_gofr_minus[:] = res.scaled_g * (1 + 3e-3 * rng.uniform(0.9, 1.1, 600))
        res.save_results(600)
        paths.append(join(str(tmp_path), f"P12_T300.00_J{job}_data_points.npz"))
        parts.append(res)
    loaded = ref_pimc.BoxResultPM()
    loaded.hash_vib, loaded.hash_rho = Data.hash_vib, Data.hash_rho
    loaded.load_multiple_results(paths)
    assert loaded.samples == 1200
    # the reference walks list(set(paths)): the shards arrive in either order
    order = parts if loaded.scaled_rho[0] == parts[0].scaled_rho[0] else parts[::-1]
    for name in ("scaled_rho", "scaled_g", "scaled_gofr_plus", "scaled_gofr_minus"):
        assert np.array_equal(getattr(loaded, name), np.concatenate([getattr(p, name) for p in order]))
    got = ref_stats.basic_jackknife_analysis(300.0, loaded, {"E": 0.01, "Cv": 2e-5})
    want = stats_oracle.basic_jackknife_analysis(300.0, loaded.scaled_rho, loaded.scaled_g, loaded.scaled_gofr_plus,
                                                 loaded.scaled_gofr_minus, 0.01, 2e-5)
    for key, value in want.items():
        assert np.isclose(got[key], value, rtol=1e-10, atol=1e-14), key
    # and the other way round: a file written by the reference loads here
    theirs = ref_pimc.BoxResultPM(data=Data)
    theirs.path_root, theirs.id_job = str(tmp_path), 7
    for name in ("scaled_rho", "scaled_g", "scaled_gofr_plus", "scaled_gofr_minus"):
        getattr(theirs, name)[:] = getattr(parts[0], name)
    theirs.save_results(600)
    back = pimc.BoxResultPM()
    back.load_multiple_results([join(str(tmp_path), "P12_T300.00_J7_data_points.npz")])
    assert back.samples == 600 and np.array_equal(back.scaled_gofr_minus, parts[0].scaled_gofr_minus)


def test_fast_npz_writer_writes_what_numpy_writes(tmp_path):
    """pibronic_b200.npz_writer.savez == np.savez member for member (names, dtypes, shapes, values), a valid ZIP, and the
    numpy fallback for what it does not handle"""
    import zipfile
    from pibronic_b200 import npz_writer
    rng = np.random.RandomState(1)
    members = dict(hash_vib="a" * 128, hash_rho="b" * 128, number_of_samples=2_000_000, s_rho=rng.rand(2_000_000),
                   s_g=rng.rand(2_000_000), s_gP=rng.rand(300_000), empty=np.empty(0), scalar=np.float64(2.5),
                   table=np.arange(12, dtype=np.int32).reshape(3, 4))
    fast, ref = str(tmp_path / "fast.npz"), str(tmp_path / "ref.npz")
    npz_writer.savez(fast, **members)
    np.savez(ref, **members)
    assert zipfile.ZipFile(fast).testzip() is None
    with np.load(fast) as a, np.load(ref) as b:
        assert a.files == b.files
        for name in b.files:
            assert a[name].dtype == b[name].dtype and a[name].shape == b[name].shape and np.array_equal(a[name], b[name]), name
    npz_writer.savez(str(tmp_path / "noext"), x=np.arange(3))                     # numpy appends .npz: so does this
    assert np.array_equal(np.load(str(tmp_path / "noext.npz"))["x"], np.arange(3))
    strided = np.arange(10.0)[::2]
    npz_writer.savez(str(tmp_path / "strided.npz"), x=strided)                    # not C-contiguous: numpy writes it
    assert np.array_equal(np.load(str(tmp_path / "strided.npz"))["x"], strided)
