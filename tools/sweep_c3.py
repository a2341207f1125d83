"""BASELINE.json configs[2] as a whole: every (model, sampling distribution) pair of the reference's
examples/paper_1.5025058 (54 pairs, model files from tests/golden/c3_paper.npz) x 5 temperatures 250-350 K at P=128,
1e5 samples per point -- the job list examples/paper_1.5025058/submit_jobs_to_server.py:241-279 hands to SLURM through
PimcSubmissionClass.submit_jobs (job_boss.py:572-611): 270 jobs.  Here: pibronic_b200.sweep.run_sweep per pair, the pairs
dealt over the ranks when launched under torchrun.  Prints the wall time of the whole sweep (model loading, plan creation,
kernels, .npz files, analytic files included) and of the statistics pass.

    python tools/sweep_c3.py [samples_per_point]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sweep_c3.py
"""
import os
import shutil
import sys
import tempfile
import time
from os.path import abspath, dirname, join

ROOT = dirname(dirname(abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from pibronic_b200 import file_structure, stats, sweep


def main():
    samples = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    fixture = np.load(join(ROOT, "tests", "golden", "c3_paper.npz"))
    pairs = sorted({str(n).rsplit("_T", 1)[0] for n in fixture["names"]})
    root = tempfile.mkdtemp(prefix=f"pbx_c3_r{rank}_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    params = {"temperature_list": [250.0, 275.0, 300.0, 325.0, 350.0], "bead_list": [128], "number_of_samples": samples,
              "block_size": min(samples, 10_000), "seed": 7}
    mine = pairs[rank::world]
    structures = []
    for i, pair in enumerate(mine):                  # the model files of every pair, as the reference lays them out
        FS = file_structure.FileStructure(root, i, 0)
        for kind, path in (("vib", FS.path_vib_model), ("rho", FS.path_rho_model)):
            with open(path, "w", encoding="UTF8") as fh:
                fh.write(str(fixture[f"{pair}_T250/{kind}"]))
        structures.append(FS)
    sweep.run_sweep(structures[0], dict(params, temperature_list=[300.0], number_of_samples=1000, block_size=1000))   # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    points = 0
    for FS in structures:
        points += len(sweep.run_sweep(FS, params))
    torch.cuda.synchronize()
    t_sweep = time.perf_counter() - t0
    t0 = time.perf_counter()
    for FS in structures:
        stats.jackknife_analysis_of_pimc(FS, method="basic")
    t_stats = time.perf_counter() - t0
    print(f"rank {rank}/{world}: {len(mine)} pairs, {points} (T, P) points of {samples} samples x 128 beads: sweep {t_sweep:.2f} s "
          f"({1e3 * t_sweep / max(points, 1):.1f} ms per point, {points * samples * 128 / t_sweep:.3e} samples*beads/s incl. files), "
          f"jackknife statistics {t_stats:.2f} s", flush=True)
    shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
