"""Physical constants -- values must equal the reference's (pibronic/constants.py:12-28)."""
import numpy as np

# joules per electron-volt (NIST CODATA 2014, the value the reference pins)
nist_j_per_ev = np.float64(1.6021766208e-19)
# eV / K
boltzman = np.float64(1.38064852e-23) / nist_j_per_ev
hbar = 1.0
# finite-difference step in beta used for the E and Cv estimators
delta_beta = 2.0e-4


def beta(temperature):
    """1/(kB*T) in 1/eV for T in Kelvin (pibronic/constants.py:38-41)."""
    return 1. / (temperature * boltzman)


def extract_T_from_beta(beta):
    """inverse of :func:`beta` (pibronic/constants.py:44-47)."""
    return 1. / (beta * boltzman)
