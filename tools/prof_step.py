"""Profiling driver (GPU box): a few fused-kernel launches on the c2 workload, for ncu.

    ncu --set full --clock-control none --import-source on -k regex:pbx_fast -s 1 -c 1 -o gpurun_out/prof python tools/prof_step.py
"""
import sys
from os.path import abspath, dirname

sys.path.insert(0, dirname(dirname(abspath(__file__))))

import torch

from pibronic_b200 import _cabi, constants, synthetic
from pibronic_b200.model_io import VMK

X = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
flags = _cabi.FLAG_PM | (_cabi.FLAG_EIG_JACOBI if "jacobi" in sys.argv else 0)
model = synthetic.model_c2()
rho = synthetic.diagonal_of(model)
plan = _cabi.Plan(model[VMK.E], model[VMK.w], model[VMK.G1], model[VMK.G2], rho[VMK.E], rho[VMK.w], rho[VMK.G1],
                  64, constants.beta(300.0), constants.delta_beta, flags=flags, device=0)
out = torch.empty((4, X), dtype=torch.float64, device="cuda")
sums = torch.empty((X // 10000, _cabi.NSUMS), dtype=torch.float64, device="cuda")
for k in range(3):
    plan.sample_eval(100 + k, 0, X, out)
    plan.block_sums(out, X, 10000, sums)
torch.cuda.synchronize()
print("done", float(out[1].mean() / out[0].mean()))
