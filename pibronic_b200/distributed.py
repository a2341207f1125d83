"""Multi-GPU sharding of a PIMC job: one process per GPU, samples split by global index.

The reference scales out by submitting many identical SLURM jobs and merging their ``.npz`` files
(pibronic/server/job_boss.py:538-570, 606-610; pimc.py:975-1040).  Here the ranks of one
``torch.distributed`` world play the role of those jobs:

* rank r evaluates the contiguous block range ``shard_blocks(blocks, world, r)``; the Philox
  counter of a sample is its GLOBAL index, so the union over ranks is bit-identical to a
  single-GPU run with the same seed -- independent of the number of GPUs;
* the only exchange is ONE all-reduce of the per-block sums (NCCL over NVLink on GPUs, gloo in
  the CPU tests): every rank ends up with the (blocks, NSUMS) table of the whole job;
* per-sample arrays stay on their rank and are written as ``..._J{id_job+rank}_data_points.npz``,
  exactly the multi-job layout ``BoxResultPM.load_multiple_results`` merges; ``gather=True``
  additionally all-gathers them so rank 0 can write one file.
"""
import numpy as np

from . import _cabi


def shard_blocks(blocks, world, rank):
    """(first_block, n_blocks) of `rank`: contiguous, sizes differ by at most one block"""
    assert 0 <= rank < world and blocks >= 0
    base, extra = divmod(blocks, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def _world(group):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def reduce_block_sums(local_sums, first_block, total_blocks, group=None, device=None):
    """all-reduce: every rank contributes its rows of the (total_blocks, NSUMS) table, all get the sum"""
    import torch
    import torch.distributed as dist
    table = torch.zeros((total_blocks, _cabi.NSUMS), dtype=torch.float64, device=device or "cpu")
    if len(local_sums):
        table[first_block:first_block + len(local_sums)] = torch.as_tensor(local_sums, dtype=torch.float64).to(table.device)
    world, _ = _world(group)
    if world > 1:
        dist.all_reduce(table, op=dist.ReduceOp.SUM, group=group)
    return table.cpu().numpy()


def estimates_from_sums(sums, block_size, delta_beta, temperature, boltzman):
    """Z, E, Cv of the whole job from the per-block sums (pibronic/stats/stats.py:84-123)"""
    total = sums.sum(axis=0)
    X = len(sums) * block_size
    Z = total[0] / X
    Z_err = np.sqrt(max(total[3] / X - Z * Z, 0.0)) / np.sqrt(X - 1)
    E = -(total[4] / X) / Z
    Cv = ((total[5] / X) / Z - E * E) / (boltzman * temperature ** 2)
    return {"Z": Z, "Z error": Z_err, "E": E, "Cv": Cv, "number_of_samples": X}


def _compute_shard(data, first_sample, n):
    """this rank's samples on its GPU: (rows, n) results and (n/block_size, NSUMS) sums"""
    plan = data.device_plan()
    rows = 4 if plan.pm else 2
    out = np.empty((rows, n))
    _, sums = plan.sample_eval_host(data.seed, data.sample_offset + first_sample, n, out4=out,
                                    block_size=int(data.block_size))
    return out, sums


def block_compute_sharded(data, result, group=None, gather=False, save=True, compute_fn=_compute_shard):
    """block_compute[_pm] over all ranks of the process group.

    `data` / `result` are the same objects a single-GPU call would use (identical on every rank,
    same ``data.seed``).  On return ``result.block_sums`` holds the all-reduced table of the whole
    job and the rows of this rank's samples are filled in ``result.scaled_*`` (all rows with
    ``gather=True``).  Files: rank r saves ``J = (result.id_job or 0) + r`` holding only its samples
    unless ``gather`` (then rank 0 saves everything under ``J = id_job``)."""
    import torch
    import torch.distributed as dist
    from .pimc import BoxResult, BoxResultPM

    world, rank = _world(group)
    first_block, n_blocks = shard_blocks(int(data.blocks), world, rank)
    bs = int(data.block_size)
    first, n = first_block * bs, n_blocks * bs
    names = ("scaled_rho", "scaled_g", "scaled_gofr_plus", "scaled_gofr_minus")
    rows = 4 if isinstance(result, BoxResultPM) else 2       # from the result type on EVERY rank, also one without blocks
    use_cuda = world > 1 and dist.get_backend(group) == "nccl"
    device = None
    if use_cuda:
        # the collectives run on the GPU the plan computes on (LOCAL_RANK / data.device), whatever the current device is
        device = torch.device("cuda", data._device_index())
        torch.cuda.set_device(device)
    if n:
        out, sums = compute_fn(data, first, n)
        assert out.shape[0] == rows, f"compute_fn returned {out.shape[0]} rows for a {type(result).__name__}"
        for k in range(rows):
            getattr(result, names[k])[first:first + n] = out[k]
    else:
        out, sums = np.empty((rows, 0)), np.empty((0, _cabi.NSUMS))
    result.block_sums = reduce_block_sums(sums, first_block, int(data.blocks), group=group, device=device)

    if gather and world > 1:
        total = int(data.blocks) * bs
        full = torch.zeros((rows, total), dtype=torch.float64, device=device or "cpu")
        if n:
            full[:, first:first + n] = torch.as_tensor(out).to(full.device)
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)   # disjoint slices: a sum is a gather
        full = full.cpu().numpy()
        for k in range(rows):
            getattr(result, names[k])[:total] = full[k]
    if save:
        base = int(result.id_job) if result.id_job is not None else 0
        if gather:
            if rank == 0:
                result.save_results(int(data.blocks) * bs)
        elif n:
            shard_cls = BoxResultPM if isinstance(result, BoxResultPM) else BoxResult
            shard = shard_cls(X=n)
            shard.partial_name, shard.path_root = result.partial_name, result.path_root
            shard.hash_vib, shard.hash_rho, shard.id_job = result.hash_vib, result.hash_rho, base + rank
            for k in range(rows):
                getattr(shard, names[k])[:] = out[k]
            shard.save_results(n)
    return result
