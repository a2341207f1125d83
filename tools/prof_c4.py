"""Profiling driver (GPU box): the fused call on the c4 shape (A=12, N=24, P=256).

    python tools/prof_c4.py [samples] [blocked]
"""
import sys
from os.path import abspath, dirname

sys.path.insert(0, dirname(dirname(abspath(__file__))))

import torch

from pibronic_b200 import _cabi, constants, synthetic
from pibronic_b200.model_io import VMK

X = int(float(sys.argv[1])) if len(sys.argv) > 1 else 148 * 8 * 2
extra = _cabi.FLAG_NO_FUSED_DMMA if "blocked" in sys.argv[2:] else 0
model = synthetic.model_c4()
rho = synthetic.diagonal_of(model)
plan = _cabi.Plan(model[VMK.E], model[VMK.w], model[VMK.G1], model[VMK.G2], rho[VMK.E], rho[VMK.w], rho[VMK.G1],
                  256, constants.beta(300.0), constants.delta_beta, flags=_cabi.FLAG_PM | extra, device=0)
out = torch.empty((4, X), dtype=torch.float64, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in range(3):
    e0.record()
    plan.sample_eval(100 + k, 0, X, out)
    e1.record()
    torch.cuda.synchronize()
    print("c4 path=%d X=%d: %.2f ms, %.3e samples*beads/s" % (plan.kernel_path, X, e0.elapsed_time(e1), X * 256 / e0.elapsed_time(e1) * 1e3))
