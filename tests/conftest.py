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


class PaperFamily:
    """BASELINE configs[2] (c3): every model of the reference's examples/paper_1.5025058 with each of its alternate
    sampling distributions, coordinates and outputs of the UNMODIFIED reference at P=128, T=250 and 350 K
    (tests/golden/c3_paper.npz, made by tests/golden/make_c3_paper.py)"""

    def __init__(self, root):
        from oracle import pimc_oracle as orc
        self.data = np.load(join(GOLDEN, "c3_paper.npz"))
        self.names = [str(n) for n in self.data["names"]]
        self.P = int(self.data["P"])
        self.root, self.orc = str(root), orc
        self._models = {}

    def run(self, name):
        """(vib, rho, T, R, expected[4][X]) of one run; `exact(name)` has the 80-bit values"""
        key = name.rsplit("_T", 1)[0]
        if key not in self._models:
            paths = []
            for kind in ("vib", "rho"):
                path = join(self.root, f"{key}_{kind}.json")
                with open(path, "w", encoding="UTF8") as fh:
                    fh.write(str(self.data[f"{name}/{kind}"]))
                paths.append(path)
            self._models[key] = (self.orc.load_vibronic_json(paths[0]), self.orc.load_sampling_json(paths[1]))
        vib, rho = self._models[key]
        return vib, rho, float(name.rsplit("_T", 1)[1]), np.ascontiguousarray(self.data[name + "/R"]), self.data[name + "/out"]

    def exact(self, name):
        """the four numbers of every sample evaluated in 80-bit arithmetic (tests/golden/extended_precision.py)"""
        return self.data[name + "/exact"]


@pytest.fixture(scope="session")
def paper_family(tmp_path_factory):
    return PaperFamily(tmp_path_factory.mktemp("c3_paper"))
