"""Z, E, Cv and their jackknife errors from the four per-sample arrays -- on the device.

Drop-in for the "basic" path of ``pibronic.stats`` (pibronic/stats/stats.py:38-55, 84-130, 148-179,
183-196, 271-299, 336-377 and jackknife.py:60-105): same function names, same output dictionary
(``Z, Z error, E, E error, Cv, Cv error, jk_E, jk_E error, jk_Cv, jk_Cv error``), same ``*_thermo``
files.  The sums and the per-sample leave-one-out estimators are evaluated by the pbx kernels
(``pbx_stats_*``); when called right after ``block_compute_pm`` the per-sample arrays are still on the
GPU and nothing is transferred.  The "alpha" variants of the reference are not provided.
"""
import json

import numpy as np

from . import _cabi, constants, postprocessing as pp
from .pimc import BoxResultPM

__all__ = ["add_harmonic_contribution", "basic_statistical_analysis", "basic_jackknife_analysis", "consistent_jackknife_analysis",
           "starmap_wrapper", "statistical_analysis_of_pimc", "jackknife_analysis_of_pimc"]

_BASIC_KEYS = ("Z", "Z error", "E", "E error", "Cv", "Cv error")


def add_harmonic_contribution(input_dict, E_sampling, Cv_sampling):
    """adds the constant contribution of the sampling distribution to the energy and the heat capacity"""
    input_dict["E"] += E_sampling
    input_dict["Cv"] += Cv_sampling


def _device_statistics(temperature, pimc_result, delta_beta=None, device=None):
    """the ten numbers of pbx_stats_arrays_host for a BoxResultPM (host arrays are uploaded once).  The finite-difference
    step is `delta_beta`, else the one the result object carries (BoxResultPM.delta_beta, set by block_compute_pm and
    stored in the .npz), else constants.delta_beta (what the reference's stats always assume, stats.py:38-55)"""
    import torch
    if device is None:
        if not torch.cuda.is_available():
            raise _cabi.PbxError("no CUDA device: pibronic_b200 has no CPU path")
        device = torch.cuda.current_device()
    if delta_beta is None:
        delta_beta = getattr(pimc_result, "delta_beta", None) or constants.delta_beta
    store = getattr(pimc_result, "_store", None)
    n = int(pimc_result.samples)
    rows = (pimc_result.scaled_rho, pimc_result.scaled_g, pimc_result.scaled_gofr_plus, pimc_result.scaled_gofr_minus)
    if store is not None and store.shape == (4, n) and all(r.ctypes.data == store[k].ctypes.data for k, r in enumerate(rows)):
        host = store
    else:
        host = np.stack([np.asarray(r, dtype=np.float64) for r in rows])
    return _cabi.stats_arrays_host(host, constants.beta(temperature), delta_beta, device)


def basic_statistical_analysis(temperature, pimc_result, analytic_data):
    """Z, E, Cv (errors of E and Cv are zero in the basic estimate) with the harmonic contribution added"""
    full = _device_statistics(temperature, pimc_result)
    basic_dict = {key: full[key] for key in _BASIC_KEYS}
    add_harmonic_contribution(basic_dict, analytic_data["E"], analytic_data["Cv"])
    return basic_dict


def basic_jackknife_analysis(temperature, pimc_result, analytic_data):
    """the basic estimates plus their per-sample leave-one-out jackknife values and errors (jk_* keys)"""
    full = _device_statistics(temperature, pimc_result)
    output_dict = {key: full[key] for key in _BASIC_KEYS}
    add_harmonic_contribution(output_dict, analytic_data["E"], analytic_data["Cv"])
    jk_dict = {"E": full["jk_E"], "E error": full["jk_E error"], "Cv": full["jk_Cv"], "Cv error": full["jk_Cv error"]}
    add_harmonic_contribution(jk_dict, analytic_data["E"], analytic_data["Cv"])
    for key, value in jk_dict.items():
        output_dict["jk_" + key] = value
    return output_dict


def consistent_jackknife_analysis(temperature, pimc_result, analytic_data=None, delta_beta=None):
    """NOT in the reference.  For results computed with ``data.m_tau_pm = True`` (PBX_FLAG_M_TAU_PM: g+- built with
    exp(-tau+- V)) the finite differences of g already carry the whole beta dependence of Z = Z_rho <g/rho>:

        E = -<d1>/<r>,   Cv = (<d2>/<r> - E^2) / (kB T^2)          -- nothing is added.

    The reference's "basic" estimator differentiates only the harmonic factors (M always uses tau, pimc.py:1183) and then
    adds E and Cv of the sampling model (stats.py:126-130); on the reference's own test model data_set_1 that gives
    E = +0.104, Cv = 3.7e-3 against the sum-over-states values -0.4233 and 1.53e-4, which this estimator reproduces
    within its jackknife error (tests/test_gpu_parity.py::test_pimc_reproduces_the_sum_over_states_thermodynamics).
    If `analytic_data` has "Z" (the sampling model's partition function) Z is returned in absolute units.
    `delta_beta` must be the step the results were computed with (default: the one recorded in the result object)."""
    out = _device_statistics(temperature, pimc_result, delta_beta=delta_beta)
    if analytic_data is not None and "Z" in analytic_data:
        out["Z"] *= analytic_data["Z"]
        out["Z error"] *= analytic_data["Z"]
    return out


def starmap_wrapper(FS, P, T, statistical_operation):
    """loads every shard of (P, T) and the analytic data, runs the operation, writes the *_thermo file"""
    pimc_results = BoxResultPM()
    rhoData = {}
    pp.load_pimc_data(FS, P, T, pimc_results)
    pp.load_analytic_data(FS, T, rhoData)
    output_dict = statistical_operation(T, pimc_results, rhoData)
    try:
        output_dict["hash_vib"] = FS.hash_vib
        output_dict["hash_rho"] = FS.hash_rho
    except AttributeError:
        raise AttributeError(f"FS {FS} doesn't have hash_vib or hash_rho attributes!")
    assert pimc_results.samples != 0, "pimc_results has 0 samples! reading the data failed?!"
    path = FS.template_jackknife.format(P=P, T=T, X=pimc_results.samples)
    with open(path, mode='w', encoding='UTF8') as target_file:
        target_file.write(json.dumps(output_dict))
    return output_dict


def _analysis_of_pimc(FS, operation):
    FS.generate_model_hashes()
    list_pimc = pp.retrive_pimc_file_list(FS)
    beads = pp.extract_bead_paramater_list(list_pimc)
    temperatures = pp.extract_temperature_paramater_list(list_pimc)
    # one GPU evaluates the (P, T) pairs one after the other: each is two short kernels
    # (the reference fans them out over a 12-process pool, stats.py:249, 362)
    return {(P, T): starmap_wrapper(FS, P, T, operation) for P in beads for T in temperatures}


def statistical_analysis_of_pimc(FS, method="basic", location="local", samples=None):
    """Z, E, Cv for every (P, T) with result files under FS; writes the *_thermo files"""
    if method != "basic":
        raise Exception(f"Invalid value for parameter method:({method})")
    return _analysis_of_pimc(FS, basic_statistical_analysis)


def jackknife_analysis_of_pimc(FS, method="basic", location="local", samples=None):
    """the same with the jackknife estimates and errors"""
    if method != "basic":
        raise Exception(f"Invalid value for parameter method:({method})")
    return _analysis_of_pimc(FS, basic_jackknife_analysis)


def statistics_of_last_run(data):
    """Z, E, Cv + jackknife of the samples the most recent block_compute_pm(data, ...) left on the GPU --
    no transfer at all (the harmonic contribution is NOT added)"""
    return data.device_plan(pm=True).stats_last()
