"""N>1 host logic on CPU: world_size-2 gloo process group, the device compute replaced by a
deterministic stand-in (a function of the GLOBAL sample index, like the Philox-counter kernels)."""
import os
import socket
from os.path import join

import numpy as np
import pytest

from pibronic_b200 import _cabi, distributed
from pibronic_b200.distributed import shard_blocks


def test_shard_blocks_partitions_exactly():
    for blocks in (0, 1, 7, 8, 64, 100):
        for world in (1, 2, 3, 8):
            spans = [shard_blocks(blocks, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == blocks
            for (f0, n0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + n0 == f1
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1


def test_estimates_from_sums_match_stats_formulas():
    rng = np.random.RandomState(0)
    X, B, db, T = 4000, 100, 2e-4, 300.0
    rho = rng.uniform(0.5, 1.5, X)
    g = rho * rng.uniform(1.0, 2.0, X)
    gp, gm = g * (1 - 3e-3 * rng.uniform(0.9, 1.1, X)), g * (1 + 3e-3 * rng.uniform(0.9, 1.1, X))
    r, d1, d2 = g / rho, (gp - gm) / rho / (2 * db), (gp - 2 * g + gm) / rho / db ** 2
    sums = np.stack([np.stack([v[b * B:(b + 1) * B].sum() for v in (r, gp / rho, gm / rho, r * r, d1, d2, d1 * d1, d2 * d2)])
                     for b in range(X // B)])
    from oracle import pimc_oracle as orc
    want = orc.basic_properties(X, T, r, d1, d2)
    got = distributed.estimates_from_sums(sums, B, db, T, orc.BOLTZMANN_EV)
    for key in ("Z", "Z error", "E", "Cv"):
        assert np.isclose(got[key], want[key], rtol=1e-9), key


def _fake_compute(data, first, n):
    """stand-in for the GPU: values depend only on the global sample index"""
    idx = np.arange(first, first + n, dtype=np.float64) + data.sample_offset
    out = np.stack([1.0 + 0.001 * idx, 2.0 + np.sin(idx), 2.0 + np.sin(idx) + 1e-3, 2.0 + np.sin(idx) - 1e-3])
    bs = data.block_size
    r = out[1] / out[0]
    sums = np.zeros((n // bs, _cabi.NSUMS))
    sums[:, 0] = r.reshape(-1, bs).sum(axis=1)
    sums[:, 3] = (r * r).reshape(-1, bs).sum(axis=1)
    return out, sums


def _worker(rank, world, port, root, gather):
    import torch.distributed as dist
    from pibronic_b200 import pimc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class Data:
        blocks, block_size, samples, beads, temperature, sample_offset = 7, 4, 28, 12, 300.0, 100
        hash_vib, hash_rho = "hv", "hr"
    result = pimc.BoxResultPM(data=Data)
    result.path_root, result.id_job = root, 10
    distributed.block_compute_sharded(Data, result, gather=gather, compute_fn=_fake_compute)
    np.save(join(root, f"sums_{int(gather)}_{rank}.npy"), result.block_sums)
    np.save(join(root, f"g_{int(gather)}_{rank}.npy"), result.scaled_g)
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("gather", [False, True])
def test_two_rank_sharding_matches_single_process(tmp_path, gather):
    import torch.multiprocessing as mp
    from pibronic_b200 import pimc
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), gather), nprocs=world, join=True)

    class Data:
        blocks, block_size, sample_offset = 7, 4, 100
    full, full_sums = _fake_compute(Data, 0, 28)
    for rank in range(world):
        sums = np.load(join(tmp_path, f"sums_{int(gather)}_{rank}.npy"))
        assert np.allclose(sums, full_sums, rtol=1e-15)                      # all-reduced table, identical on both
        g = np.load(join(tmp_path, f"g_{int(gather)}_{rank}.npy"))
        first, n = (b * 4 for b in shard_blocks(7, world, rank))
        assert np.array_equal(g[first:first + n], full[1][first:first + n])
        if gather:
            assert np.array_equal(g, full[1])
        else:
            other = np.ones(28, dtype=bool)
            other[first:first + n] = False
            assert np.isnan(g[other]).all()
    if gather:
        merged = pimc.BoxResultPM()
        merged.load_multiple_results([join(tmp_path, "P12_T300.00_J10_data_points.npz")])
        assert np.array_equal(merged.scaled_g, full[1])
    else:   # one file per rank, J = id_job + rank: the reference's multi-job layout
        paths = [join(tmp_path, f"P12_T300.00_J{10 + r}_data_points.npz") for r in range(world)]
        merged = pimc.BoxResultPM()
        merged.load_multiple_results(paths)
        assert merged.samples == 28
        assert np.array_equal(np.sort(merged.scaled_gofr_plus), np.sort(full[2]))


def _fake_compute_two_rows(data, first, n):
    out, sums = _fake_compute(data, first, n)
    return out[:2], sums


def _worker_idle_rank(rank, world, port, root):
    import torch.distributed as dist
    from pibronic_b200 import pimc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class Data:
        blocks, block_size, samples, beads, temperature, sample_offset = 2, 4, 8, 12, 300.0, 0
        hash_vib, hash_rho = "hv", "hr"
    result = pimc.BoxResult(data=Data)              # the non-PM container: two rows
    result.path_root, result.id_job = root, 0
    distributed.block_compute_sharded(Data, result, gather=True, compute_fn=_fake_compute_two_rows)
    np.save(join(root, f"g_idle_{rank}.npy"), result.scaled_g)
    dist.destroy_process_group()


def test_rank_without_blocks_and_two_row_results(tmp_path):
    """more ranks than blocks: the idle rank takes part in the collectives with the row count of the result TYPE
    (a BoxResult has two rows, no g+-), and every rank ends up with the gathered arrays"""
    import torch.multiprocessing as mp
    world = 3
    mp.spawn(_worker_idle_rank, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)

    class Data:
        blocks, block_size, sample_offset = 2, 4, 0
    full, _ = _fake_compute(Data, 0, 8)
    for rank in range(world):
        assert np.array_equal(np.load(join(tmp_path, f"g_idle_{rank}.npy")), full[1])
