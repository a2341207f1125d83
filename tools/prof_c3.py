"""Profiling driver (GPU box): the fused call on the c3 shape (A=2, N=2, P=128).

    python tools/prof_c3.py [samples]
"""
import sys
from os.path import abspath, dirname

sys.path.insert(0, dirname(dirname(abspath(__file__))))

import torch

from pibronic_b200 import _cabi, constants, synthetic
from pibronic_b200.model_io import VMK

X = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
model = synthetic.coupled_model(2, 2, (0.02, 0.04), (0.1, 0.2), seed=3, linear=0.05, quadratic=0.0)
rho = synthetic.diagonal_of(model)
plan = _cabi.Plan(model[VMK.E], model[VMK.w], model[VMK.G1], model.get(VMK.G2), rho[VMK.E], rho[VMK.w], rho[VMK.G1],
                  128, constants.beta(300.0), constants.delta_beta, flags=_cabi.FLAG_PM, device=0)
out = torch.empty((4, X), dtype=torch.float64, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in range(3):
    e0.record()
    plan.sample_eval(100 + k, 0, X, out)
    e1.record()
    torch.cuda.synchronize()
    print("c3 path=%d X=%d: %.3f ms, %.3e samples*beads/s" % (plan.kernel_path, X, e0.elapsed_time(e1), X * 128 / e0.elapsed_time(e1) * 1e3))
