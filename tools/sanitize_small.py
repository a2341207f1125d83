"""Small invocations of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck) on the GPU box.

    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import sys
from os.path import abspath, dirname, join

ROOT = dirname(dirname(abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, join(ROOT, "tests"))

import numpy as np

from conftest import GoldenCase
from pibronic_b200 import _cabi

for name, n in (("c2_4x6", 600), ("quad_3x4", 300), ("jt_rho4", 300), ("c4mini_12x24", 40), ("syn_7x12", 70)):
    case = GoldenCase(name)
    for extra in (0, _cabi.FLAG_NO_WARPSPEC, _cabi.FLAG_FORCE_GENERIC, _cabi.FLAG_M_TAU_PM, _cabi.FLAG_PREFER_DMMA,
                  _cabi.FLAG_NO_FUSED_DMMA):
        try:
            plan = case.plan(_cabi.FLAG_PM | _cabi.QUIRK_RHO_TRUNC | extra)
        except _cabi.PbxError:
            continue                      # M_TAU_PM without a register-resident kernel / a library built without it
        out, sums = plan.sample_eval_host(7, 5, n, block_size=100)
        got = plan.eval_coords_host(case.R)
        assert np.all(np.isfinite(out)) and np.all(np.isfinite(got))
        st = plan.stats_host(out)
        plan.close()
    print(name, "ok")
