"""End-to-end physics check on the GPU box: PIMC estimates of Z, E, Cv of the reference's test model data_set_1
(sampling distribution rho_1) against the exact sum-over-states values its Julia dependency produced
(tests/golden/sos/), for several numbers of beads.

    python tools/sos_check.py [X] [P ...]

E and Cv use PBX_FLAG_M_TAU_PM (g+- built with exp(-tau+- V)): E = -<d1>/<r>, Cv = (<d2>/<r> - E^2)/(kB T^2), nothing added.
With the reference's estimator (M(tau) for all three variants + the sampling model's E, Cv added) they come out wrong.
"""
import json
import sys
from os.path import abspath, dirname, join

ROOT = dirname(dirname(abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from oracle import pimc_oracle as orc          # model file readers only
from pibronic_b200 import _cabi, constants

SOS = join(ROOT, "tests", "golden", "sos")


def main():
    pos = [a for a in sys.argv[1:]]
    X = int(float(pos[0])) if pos else 20_000_000
    beads = [int(p) for p in pos[1:]] or [16, 32, 64, 128, 256]
    vib = orc.load_vibronic_json(join(SOS, "coupled_model.json"))
    rho = orc.load_sampling_json(join(SOS, "sampling_model.json"))
    sos = {k: v[0] for k, v in json.load(open(join(SOS, "sos_B80.json"))).items()}
    T = 300.0
    print(f"exact: Z {sos['Z_coupled']:.8e}  E {sos['E_coupled']:.8f}  Cv {sos['Cv_coupled']:.6e}   (Z_rho {sos['Z_sampling']:.8e})")
    out = torch.empty((4, X), dtype=torch.float64, device="cuda")
    for P in beads:
        for name, flags in (("consistent", _cabi.FLAG_PM | _cabi.FLAG_M_TAU_PM), ("reference", _cabi.FLAG_PM)):
            plan = _cabi.Plan(vib["E"], vib["w"], vib["L"], vib["Q"], rho["E"], rho["w"], rho["L"], P, constants.beta(T),
                              constants.delta_beta, flags=flags, device=0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.sample_eval(2026 + P, 0, X, out)
            e1.record()
            st = plan.stats(out, X)
            ms = e0.elapsed_time(e1)
            Z, dZ = st["Z"] * sos["Z_sampling"], st["Z error"] * sos["Z_sampling"]
            add_E, add_Cv = (0.0, 0.0) if name == "consistent" else (sos["E_sampling"], sos["Cv_sampling"])
            print(f"P={P:4d} {name:10s} {ms:7.1f} ms  Z {Z:.6e} +- {dZ:.1e} ({(Z / sos['Z_coupled'] - 1) / (dZ / Z):+.1f} sigma)  "
                  f"E {st['jk_E'] + add_E:+.6f} +- {st['jk_E error']:.1e}  Cv {st['jk_Cv'] + add_Cv:+.4e} +- {st['jk_Cv error']:.1e}")
            plan.close()


if __name__ == "__main__":
    main()
