"""Closed-form thermodynamics of the sampling distribution rho -> ``analytic_results.json``.

The reference obtains these numbers by shelling out to the Julia package VibronicToolkit.jl
(pibronic/julia_wrapper.py:185-197, setup.py:46-48), which is not vendored; ``pibronic.stats`` cannot
run without the file (pibronic/data/postprocessing.py:313-343).  For the distribution the PIMC sampler
actually draws from -- a mixture of displaced harmonic oscillators, linear terms only, the same one
``ModelSampling`` builds (pimc.py:336-351) -- they are elementary:

    Z_rho(beta) = sum_a exp(-beta (E_a + Delta_a)) * prod_n [2 sinh(beta w_n / 2)]^-1
    E_rho  = -d ln Z / d beta  = <Etilde>_w + sum_n (w_n/2) coth(beta w_n/2)
    Cv_rho = kB beta^2 d^2 ln Z / d beta^2 = kB beta^2 [Var_w(Etilde) + sum_n (w_n/2)^2 csch^2(beta w_n/2)]

Known answers: the reference's Julia output fixtures tests/test_models/data_set_*/rho_*/parameters/sos_B*.json.
"""
import json
import os

import numpy as np

from . import constants
from . import model_io as vIO
from .model_io import VMK


def sampling_model_terms(model):
    """(Etilde(A,), omega(N,)) of a diagonal model dictionary: energies shifted by the linear displacement"""
    w = np.asarray(model[VMK.w], dtype=float)
    A = int(model[VMK.A])
    E = np.asarray(model.get(VMK.E, np.zeros(A)), dtype=float)
    L = np.asarray(model.get(VMK.G1, np.zeros((len(w), A))), dtype=float)
    delta = -0.5 * (L ** 2. / w[:, None]).sum(axis=0)
    return E + delta, w


def partition_function(tilde_energy, omega, beta):
    """Z_rho(beta)"""
    shift = tilde_energy.min()
    electronic = np.exp(-beta * (tilde_energy - shift)).sum() * np.exp(-beta * shift)
    return electronic / np.prod(2. * np.sinh(beta * omega / 2.))


def thermodynamics(tilde_energy, omega, beta, delta_beta=constants.delta_beta):
    """the five numbers stats needs for one temperature (keys of postprocessing.load_analytic_data)"""
    w = np.exp(-beta * (tilde_energy - tilde_energy.min()))
    w /= w.sum()
    mean_e = (w * tilde_energy).sum()
    var_e = (w * (tilde_energy - mean_e) ** 2).sum()
    half = omega / 2.
    E = mean_e + (half / np.tanh(beta * half)).sum()
    Cv = constants.boltzman * beta ** 2 * (var_e + ((half / np.sinh(beta * half)) ** 2).sum())
    return {
        "Z_sampling": float(partition_function(tilde_energy, omega, beta)),
        "E_sampling": float(E),
        "Cv_sampling": float(Cv),
        "Z_sampling+beta": float(partition_function(tilde_energy, omega, beta + delta_beta)),
        "Z_sampling-beta": float(partition_function(tilde_energy, omega, beta - delta_beta)),
        "beta": float(beta),
    }


def analytic_of_sampling_model(FS, beta, delta_beta=constants.delta_beta):
    """adds the entry for temperature T(beta) to FS.path_analytic_rho (same role and file as
    julia_wrapper.analytic_of_sampling_model); entries computed for other model hashes are discarded"""
    FS.generate_model_hashes()
    model = vIO.load_diagonal_model_from_JSON(FS.path_rho_model)
    results = thermodynamics(*sampling_model_terms(model), beta, delta_beta)
    # the names julia_wrapper.keyDict gives the same quantities
    results.update(Z_rho=results["Z_sampling"], E_rho=results["E_sampling"], Cv_rho=results["Cv_sampling"])
    old = {}
    if os.path.isfile(FS.path_analytic_rho):
        with open(FS.path_analytic_rho, "r") as fh:
            text = fh.read()
            if len(text) > 1:
                old = json.loads(text)
    if old.get("hash_vib") != FS.hash_vib or old.get("hash_rho") != FS.hash_rho:
        old = {}
    old["hash_vib"], old["hash_rho"] = FS.hash_vib, FS.hash_rho
    old["{:.2f}".format(constants.extract_T_from_beta(beta))] = results
    with open(FS.path_analytic_rho, "w") as fh:
        json.dump(old, fh)
    return results
