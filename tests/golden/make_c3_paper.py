"""Golden fixture for BASELINE.json configs[2] (c3): EVERY model of the reference's
examples/paper_1.5025058 (4 systems x 6 parameter sets, A=2, N=2) with each of its alternate sampling
distributions (alternate_rhos/*_D{k}_R{0,1,2}.json; A_rho = 2, 4 or 8), run through the UNMODIFIED
reference at the two ends of the paper's temperature sweep with P=128 beads.

    python tests/golden/make_c3_paper.py        (build container only: needs /root/reference)

Writes tests/golden/c3_paper.npz: per run the model texts, the bead coordinates the reference drew, its
scaled_rho / scaled_g / scaled_gofr_plus / scaled_gofr_minus for exactly those coordinates, and the same
four numbers evaluated in 80-bit arithmetic (tests/golden/extended_precision.py) -- the reference's
coth (q^2+q'^2) - 2 csch q q' cancels 4 digits at tau*omega = 0.008, so its float64 values are off by up
to 3.9e-10 relative on this family.
"""
import os
import re
import sys
from os.path import join

import numpy as np

import extended_precision as xp
import make_golden as mg      # sets up the MagicMock shims and imports the reference

sys.path.insert(0, mg.REPO)
from oracle import pimc_oracle as orc  # noqa: E402

EX = join(mg.REFERENCE, "examples", "paper_1.5025058")
TEMPERATURES = (250.0, 350.0)   # ends of the sweep 250..350 K (examples/paper_1.5025058/submit_jobs_to_server.py:257)
P, X = 128, 6


def main():
    out = {}
    names = []
    rho_dir, vib_dir = join(EX, "alternate_rhos"), join(EX, "input_json")
    for fname in sorted(os.listdir(rho_dir)):
        m = re.fullmatch(r"([a-z]+)_D(\d)_R(\d)\.json", fname)
        if not m:
            continue
        system, k, r = m.group(1), int(m.group(2)), int(m.group(3))
        path_vib, path_rho = join(vib_dir, f"{system}_{k}.json"), join(rho_dir, fname)
        with open(path_vib, encoding="UTF8") as fh:
            vib_text = fh.read()
        with open(path_rho, encoding="UTF8") as fh:
            rho_text = fh.read()
        for T in TEMPERATURES:
            name = f"{system}_D{k}_R{r}_T{T:.0f}"
            ref = mg.run_reference(path_vib, path_rho, P=P, T=T, X=X, B=X, seed=1000 + len(names))
            names.append(name)
            out[name + "/vib"] = np.array(vib_text)
            out[name + "/rho"] = np.array(rho_text)
            out[name + "/R"] = ref["R"]
            out[name + "/out"] = np.stack([ref[key] for key in ("s_rho", "s_g", "s_gP", "s_gM")])
            # the same four numbers in 80-bit arithmetic, rounded once to float64: where the product of 128 bead matrices is
            # ill-conditioned the reference's own float64 value is only good to ~1e-10 and this is the yardstick
            exact = xp.evaluate(orc.load_vibronic_json(path_vib), orc.load_sampling_json(path_rho), P, T, ref["R"])
            out[name + "/exact"] = exact.astype(np.float64)
            ref_err = float(np.max(np.abs((out[name + "/out"] - exact) / exact)))
            ratio = ref["s_g"] / ref["s_rho"]
            print(f"{name:28s} A_rho={ref['rho_weight'].shape[0]}  mean g/rho = {ratio.mean():.6g}  reference vs 80-bit: {ref_err:.1e}")
    out["names"] = np.array(names)
    out["P"], out["temperatures"] = P, np.array(TEMPERATURES)
    np.savez_compressed(join(mg.HERE, "c3_paper.npz"), **out)
    print(len(names), "runs ->", join(mg.HERE, "c3_paper.npz"), os.path.getsize(join(mg.HERE, "c3_paper.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
