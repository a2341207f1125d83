"""Developer probe (GPU box): quick timing of the fused kernel on the c2 workload, all variants.

    python tools/probe.py [X]
"""
import sys
from os.path import abspath, dirname

sys.path.insert(0, dirname(dirname(abspath(__file__))))

import numpy as np
import torch

from pibronic_b200 import _cabi, synthetic, constants
from pibronic_b200.model_io import VMK


def make_plan(model, P, T, flags):
    rho = synthetic.diagonal_of(model)
    return _cabi.Plan(model[VMK.E], model[VMK.w], model[VMK.G1], model[VMK.G2], rho[VMK.E], rho[VMK.w], rho[VMK.G1],
                      P, constants.beta(T), constants.delta_beta, flags=flags, device=0)


def time_plan(plan, X, reps=3):
    out = torch.empty((4, X), dtype=torch.float64, device="cuda")
    plan.sample_eval(1234, 0, X, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record()
        plan.sample_eval(1234, 0, X, out)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def main():
    pos = [a for a in sys.argv[1:] if not a.startswith("--")]
    X = int(float(pos[0])) if pos else 1_000_000
    model = synthetic.model_c2()
    if "--quick" in sys.argv:
        import os
        for name, extra in (("warp-specialised", 0), ("one-role", _cabi.FLAG_NO_WARPSPEC)):
            plan = make_plan(model, 64, 300.0, _cabi.FLAG_PM | extra)
            ms, out = time_plan(plan, X, reps=5)
            o = out.cpu().numpy()
            print(f"{os.path.basename(os.environ.get('PBX_LIB', 'default')):24s} {name:17s} c2 expm PM X={X} {ms:8.3f} ms  "
                  f"{X * 64 / ms * 1e3:.3e} samples*beads/s <g/rho>={(o[1] / o[0]).mean():.5f}")
            plan.close()
        return
    print("fp64 peak (DFMA probe): %.2f TFLOP/s" % _cabi.fp64_peak_tflops(0))
    for name, flags in (("expm", _cabi.FLAG_PM), ("jacobi", _cabi.FLAG_PM | _cabi.FLAG_EIG_JACOBI),
                        ("expm nonPM", 0)):
        plan = make_plan(model, 64, 300.0, flags)
        ms, out = time_plan(plan, X)
        o = out.cpu().numpy()
        r = o[1] / o[0]
        print(f"c2 {name:12s} fast={plan.is_fast} X={X} {ms:8.2f} ms  {X * 64 / ms * 1e3:.3e} samples*beads/s   "
              f"<g/rho>={r.mean():.5f} +- {r.std() / np.sqrt(X):.5f}")
        plan.close()
    plan = make_plan(model, 64, 300.0, _cabi.FLAG_PM | _cabi.FLAG_FORCE_GENERIC)
    Xg = min(X, 20000)
    ms, out = time_plan(plan, Xg, reps=2)
    o = out.cpu().numpy()
    r = o[1] / o[0]
    print(f"c2 generic      X={Xg} {ms:8.2f} ms  {Xg * 64 / ms * 1e3:.3e} samples*beads/s   <g/rho>={r.mean():.5f}")
    plan.close()
    m4 = synthetic.model_c4()
    plan = make_plan(m4, 256, 300.0, _cabi.FLAG_PM)
    Xg = 2048
    ms, out = time_plan(plan, Xg, reps=2)
    o = out.cpu().numpy()
    r = o[1] / o[0]
    print(f"c4 generic      X={Xg} {ms:8.2f} ms  {Xg * 256 / ms * 1e3:.3e} samples*beads/s   <g/rho>={r.mean():.5f} "
          f"+- {r.std() / np.sqrt(Xg):.5f}")


if __name__ == "__main__":
    main()
