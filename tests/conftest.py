"""pytest configuration: markers, paths and the golden-case loader shared by all tests."""
import os
import sys
from os.path import abspath, dirname, join

import numpy as np
import pytest

REPO = dirname(dirname(abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = join(REPO, "tests", "golden")
CASES = join(GOLDEN, "cases")
CASE_NAMES = sorted(os.listdir(CASES))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class GoldenCase:
    """one parity case: model pair + coordinates drawn by the reference + the reference's outputs"""

    def __init__(self, name):
        from oracle import pimc_oracle as orc
        self.name = name
        self.dir = join(CASES, name)
        self.path_vib = join(self.dir, "coupled_model.json")
        self.path_rho = join(self.dir, "sampling_model.json")
        self.ref = np.load(join(self.dir, "ref.npz"))
        self.vib = orc.load_vibronic_json(self.path_vib)
        self.rho = orc.load_sampling_json(self.path_rho)
        self.P, self.T = int(self.ref["P"]), float(self.ref["T"])
        self.R = np.ascontiguousarray(self.ref["R"])
        self.expected = np.stack([self.ref[k] for k in ("s_rho", "s_g", "s_gP", "s_gM")])

    def oracle_tables(self, rho_trunc=True):
        from oracle import pimc_oracle as orc
        return orc.precompute(self.vib, self.rho, self.P, self.T, rho_trunc=rho_trunc)

    def plan(self, flags=None, device=0):
        from oracle import pimc_oracle as orc
        from pibronic_b200 import _cabi
        if flags is None:
            flags = _cabi.FLAG_PM | _cabi.QUIRK_RHO_TRUNC
        beta = orc.beta_of(self.T)
        return _cabi.Plan(self.vib["E"], self.vib["w"], self.vib["L"], self.vib["Q"], self.rho["E"], self.rho["w"],
                          self.rho["L"], self.P, beta, orc.DELTA_BETA, flags=flags, device=device)


@pytest.fixture(scope="session", params=CASE_NAMES)
def case(request):
    return GoldenCase(request.param)


@pytest.fixture(scope="session")
def cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch
