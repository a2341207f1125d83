"""Temperature / bead sweep on one GPU box: the local counterpart of the reference's SLURM fan-out.

``PimcSubmissionClass.submit_jobs`` (pibronic/server/job_boss.py:538-611) submits, for every
temperature and every number of beads, ``n_jobs`` identical jobs of at most 1e5 samples each.  On a
B200 a (T, P) point of 1e6 samples takes milliseconds, so the sweep is a loop in one process (or one
per GPU under torchrun: every rank takes the same (T, P) list and a share of the blocks).  File layout
and parameter names are the reference's: one ``P{P}_T{T:.2f}_J{J}_data_points.npz`` per (T, P, rank).
"""
import copy

from . import distributed, pimc

DEFAULT_PARAMETERS = {           # the keys of PimcSubmissionClass.param_dict that matter off-cluster
    "temperature_list": [300.0, ],
    "bead_list": [12, ],
    "number_of_samples": int(1e4),
    "block_size": int(1e3),
    "id_job": 0,
    "seed": None,
}


def setup_blocks(param_dict):
    """blocks per (T, P) point; same rule as setup_blocks_and_jobs without the 1e5-samples-per-job cap"""
    n_samples, block_size = param_dict["number_of_samples"], param_dict["block_size"]
    assert isinstance(n_samples, int) and isinstance(block_size, int)
    assert block_size <= n_samples, f"block size {block_size} must be less than or equal to the number of samples {n_samples}"
    return n_samples // block_size


def run_sweep(FS, input_param_dict=None, analytic=True):
    """evaluates every (temperature, beads) point; returns {(P, T): BoxResultPM}"""
    params = copy.deepcopy(DEFAULT_PARAMETERS)
    params.update(input_param_dict or {})
    blocks = setup_blocks(params)
    FS.generate_model_hashes()
    results = {}
    for T in params["temperature_list"]:
        for P in params["bead_list"]:
            data = pimc.BoxDataPM.from_FileStructure(FS)
            data.samples = blocks * params["block_size"]
            data.block_size, data.blocks = params["block_size"], blocks
            data.beads, data.temperature = P, T
            data.hash_vib, data.hash_rho = FS.hash_vib, FS.hash_rho
            data.seed = params["seed"]
            data.preprocess()
            result = pimc.BoxResultPM(data=data)
            result.path_root, result.id_job = FS.path_rho_results, params["id_job"]
            distributed.block_compute_sharded(data, result)     # one rank == plain block_compute_pm
            if analytic:
                from .analytic import analytic_of_sampling_model
                analytic_of_sampling_model(FS, data.beta, data.delta_beta)
            data.release()
            results[(P, T)] = result
    return results
