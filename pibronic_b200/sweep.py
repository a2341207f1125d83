"""Temperature / bead sweep on one GPU box: the local counterpart of the reference's SLURM fan-out.

``PimcSubmissionClass.submit_jobs`` (pibronic/server/job_boss.py:538-611) submits, for every
temperature and every number of beads, ``n_jobs`` identical jobs of at most 1e5 samples each, and the
example drivers loop that over every (model, sampling distribution) pair
(examples/paper_1.5025058/submit_jobs_to_server.py:241-279).  Here:

* the (T, P) points of a sweep are dealt out to the ranks of the process group (one process per GPU under
  torchrun): point i belongs to rank i % world -- no collective on the data path, every rank writes the files
  of its own points in the reference's naming scheme, ``P{P}_T{T:.2f}_J{J}_data_points.npz``.  With fewer
  points than ranks the BLOCKS of every point are shared out instead (``distributed.block_compute_sharded``);
* on a rank the points are software pipelined over two CUDA streams: while the kernel of point i runs, the
  host writes the ``.npz`` of point i-1 (on a B200 a point of 1e5 samples x 128 beads is 0.4 ms of kernel
  time; writing its 3.2 MB takes longer);
* model files, hashes and the analytic sampling-model data (rank 0 only, then a barrier) are read once per sweep.
"""
import copy

import numpy as np

from . import _cabi, distributed, pimc

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


def _world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def _shared_seed(seed):
    """one seed for the whole sweep: drawn on rank 0 and broadcast when the caller gave none"""
    world, rank = _world()
    if seed is not None:
        return int(seed)
    seed = pimc._fresh_seed() >> 1                     # 63 bits: fits the int64 tensor of the broadcast
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([seed if rank == 0 else 0], dtype=torch.int64)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0)
        seed = int(t.cpu()[0])
    return seed


def _new_job(FS, T, P, blocks, block_size, seed):
    data = pimc.BoxDataPM.from_FileStructure(FS)
    data.samples = blocks * block_size
    data.block_size, data.blocks = block_size, blocks
    data.beads, data.temperature = P, T
    data.hash_vib, data.hash_rho = FS.hash_vib, FS.hash_rho
    data.seed = seed
    data.preprocess()
    return data


class _Point:
    """one (T, P) point in flight: its job, its result arrays (pinned) and the CUDA event that ends its kernels"""

    def __init__(self, FS, T, P, blocks, block_size, seed, id_job, stream):
        import torch
        self.key = (P, T)
        self.data = _new_job(FS, T, P, blocks, block_size, seed)
        self.result = pimc.BoxResultPM(data=self.data)
        self.result.path_root, self.result.id_job = FS.path_rho_results, id_job
        plan = self.data.device_plan(pm=True)
        n = blocks * block_size
        with torch.cuda.device(plan.device), torch.cuda.stream(stream):
            self.out = torch.empty((4, n), dtype=torch.float64, device="cuda")
            self.sums = torch.empty((blocks, _cabi.NSUMS), dtype=torch.float64, device="cuda")
            plan.sample_eval(self.data.seed, self.data.sample_offset, n, self.out, stream=stream)
            plan.block_sums(self.out, n, block_size, self.sums, stream=stream)
            store = getattr(self.result, "_store", None)
            self.host = torch.from_numpy(store) if store is not None and store.shape == (4, n) else torch.empty((4, n), dtype=torch.float64)
            self.host.copy_(self.out, non_blocking=True)
            self.host_sums = torch.empty((blocks, _cabi.NSUMS), dtype=torch.float64, pin_memory=True)
            self.host_sums.copy_(self.sums, non_blocking=True)
            self.done = torch.cuda.Event()
            self.done.record(stream)
        self.n = n

    def finish(self):
        """waits for the kernels and copies, fills the result object, writes the .npz, frees the plan"""
        self.done.synchronize()
        names = ("scaled_rho", "scaled_g", "scaled_gofr_plus", "scaled_gofr_minus")
        host = self.host.numpy()
        for k, name in enumerate(names):
            row = getattr(self.result, name)
            if row.ctypes.data != host[k].ctypes.data:
                row[:self.n] = host[k]
        self.result.block_sums = self.host_sums.numpy().copy()
        self.result.delta_beta = self.data.delta_beta
        self.result.save_results(self.n)
        self.data.release()
        self.out = self.sums = None
        return self.result


def run_sweep(FS, input_param_dict=None, analytic=True):
    """evaluates every (temperature, beads) point; returns {(P, T): BoxResultPM} of the points this rank evaluated
    (all of them in a single process)"""
    import os
    import torch
    if "LOCAL_RANK" in os.environ and torch.cuda.is_available():     # one process per GPU under torchrun: plans AND collectives
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))         # of this rank live on its own device
    params = copy.deepcopy(DEFAULT_PARAMETERS)
    params.update(input_param_dict or {})
    blocks = setup_blocks(params)
    block_size = params["block_size"]
    world, rank = _world()
    FS.generate_model_hashes()
    seed = _shared_seed(params["seed"])
    points = [(T, P) for T in params["temperature_list"] for P in params["bead_list"]]

    if analytic:      # depends on the temperature only: once per T, written by rank 0 alone
        if rank == 0:
            from . import constants
            from .analytic import analytic_of_sampling_model
            for T in params["temperature_list"]:
                analytic_of_sampling_model(FS, constants.beta(T), constants.delta_beta)
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    results = {}
    if world > 1 and len(points) < world:
        # fewer points than GPUs: every rank takes a share of the blocks of every point
        for T, P in points:
            data = _new_job(FS, T, P, blocks, block_size, seed)
            result = pimc.BoxResultPM(data=data)
            result.path_root, result.id_job = FS.path_rho_results, params["id_job"]
            distributed.block_compute_sharded(data, result)
            data.release()
            results[(P, T)] = result
        return results

    mine = points[rank::world]
    if not mine:
        return results
    probe = _new_job(FS, mine[0][0], mine[0][1], blocks, block_size, seed)
    device = probe._device_index()
    probe.release()
    with torch.cuda.device(device):
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    in_flight = None
    for i, (T, P) in enumerate(mine):
        point = _Point(FS, T, P, blocks, block_size, seed, params["id_job"], streams[i % 2])
        if in_flight is not None:
            results[in_flight.key] = in_flight.finish()       # host I/O of point i-1 under the kernel of point i
        in_flight = point
    results[in_flight.key] = in_flight.finish()
    return results
