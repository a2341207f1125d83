"""numpy restatement of the reference's post-processing of the four per-sample arrays
(TEST INFRASTRUCTURE ONLY -- see oracle/pimc_oracle.py header).

Follows /root/reference/pibronic/stats/stats.py:38-55 (terms), 84-123 (Z, E, Cv), 126-130 (harmonic
contribution), 271-299 (basic_jackknife_analysis) and pibronic/stats/jackknife.py:60-105.
Pinned by tests/golden/stats_kat.npz: the reference's own ``basic_jackknife_analysis`` run on arrays the
reference produced (tests/golden/make_golden.py)."""
import numpy as np

from .pimc_oracle import BOLTZMANN_EV, DELTA_BETA


def basic_terms(delta_beta, rho, g, g_plus, g_minus):
    ratio = g / rho
    d1 = (g_plus - g_minus) / rho / (2. * delta_beta)
    d2 = (g_plus - (2. * g) + g_minus) / rho / delta_beta ** 2
    return ratio, d1, d2


def basic_properties(X, T, ratio, d1, d2):
    Z = np.mean(ratio)
    Z_err = np.std(ratio, ddof=0) / np.sqrt(X - 1)
    E = -np.mean(d1) / Z
    Cv = (np.mean(d2) / Z - E ** 2) / (BOLTZMANN_EV * T ** 2)
    return {"Z": Z, "Z error": Z_err, "E": E, "E error": 0.0, "Cv": Cv, "Cv error": 0.0}


def leave_one_out(X, array):
    return (np.sum(array) - array) / (X - 1)


def jackknife_properties(X, T, basic, jk_ratio, jk_d1, jk_d2):
    f_E = -jk_d1 / jk_ratio
    E = X * basic["E"] - (X - 1.) * np.mean(f_E)
    E_err = np.sqrt(X - 1.) * np.std(f_E, ddof=1)
    f_C = (jk_d2 / jk_ratio - f_E ** 2) / (BOLTZMANN_EV * T ** 2)
    Cv = X * basic["Cv"] - (X - 1.) * np.mean(f_C)
    Cv_err = np.sqrt(X - 1.) * np.std(f_C, ddof=1)
    return {"E": E, "E error": E_err, "Cv": Cv, "Cv error": Cv_err}


def basic_jackknife_analysis(T, rho, g, g_plus, g_minus, E_sampling=0.0, Cv_sampling=0.0, delta_beta=DELTA_BETA):
    X = len(rho)
    terms = basic_terms(delta_beta, rho, g, g_plus, g_minus)
    basic = basic_properties(X, T, *terms)
    jk = jackknife_properties(X, T, basic, *[leave_one_out(X, t) for t in terms])
    out = dict(basic)
    out["E"] += E_sampling
    out["Cv"] += Cv_sampling
    out.update({"jk_E": jk["E"] + E_sampling, "jk_E error": jk["E error"], "jk_Cv": jk["Cv"] + Cv_sampling,
                "jk_Cv error": jk["Cv error"]})
    return out
