"""Two-GPU run of the sharded block loop over NCCL (skipped on a single-GPU box)."""
import os
import socket
from os.path import join

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, root):
    import torch
    import torch.distributed as dist
    from pibronic_b200 import distributed, file_structure, pimc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    FS = file_structure.FileStructure(root, 0, 0)
    FS.generate_model_hashes()
    data = pimc.BoxDataPM.from_FileStructure(FS)
    data.samples, data.beads, data.temperature, data.block_size = 7000, 16, 300.0, 1000
    data.blocks, data.seed = 7, 99
    data.hash_vib, data.hash_rho = FS.hash_vib, FS.hash_rho
    data.preprocess()
    result = pimc.BoxResultPM(data=data)
    result.path_root, result.id_job = FS.path_rho_results, 0
    distributed.block_compute_sharded(data, result)
    np.save(join(root, f"sums_{rank}.npy"), result.block_sums)
    data.release()
    dist.destroy_process_group()


def test_two_gpus_reproduce_one_gpu_bit_for_bit(cuda, tmp_path):
    if cuda.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from pibronic_b200 import file_structure, pimc, synthetic
    FS = file_structure.FileStructure(tmp_path, 0, 0)
    synthetic.write_data_set(FS, synthetic.coupled_model(3, 4, (0.05, 0.2), (1.0, 1.4), seed=7, quadratic=0.08))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    # single-GPU run of the same job
    FS.generate_model_hashes()
    data = pimc.BoxDataPM.from_FileStructure(FS)
    data.samples, data.beads, data.temperature, data.block_size = 7000, 16, 300.0, 1000
    data.blocks, data.seed = 7, 99
    data.hash_vib, data.hash_rho = FS.hash_vib, FS.hash_rho
    data.preprocess()
    single = pimc.BoxResultPM(data=data)
    single.path_root, single.id_job = str(tmp_path), 50
    pimc.block_compute_pm(data, single)
    merged = pimc.BoxResultPM()
    merged.load_multiple_results([join(FS.path_rho_results, f"P16_T300.00_J{r}_data_points.npz") for r in range(2)])
    assert merged.samples == 7000
    assert np.array_equal(np.sort(merged.scaled_g), np.sort(single.scaled_g))
    for r in range(2):
        assert np.allclose(np.load(join(tmp_path, f"sums_{r}.npy")), single.block_sums, rtol=1e-13)
    data.release()


def _sweep_worker(rank, world, port, root):
    import torch.distributed as dist
    from pibronic_b200 import file_structure, sweep
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("nccl", rank=rank, world_size=world)      # no set_device here: run_sweep binds the rank's GPU itself
    FS = file_structure.FileStructure(root, 0, 0)
    params = {"temperature_list": [250.0, 300.0], "bead_list": [8, 12], "number_of_samples": 4000, "block_size": 1000, "seed": 5}
    results = sweep.run_sweep(FS, params)
    np.save(join(root, f"points_{rank}.npy"), np.array(sorted(results)))
    dist.destroy_process_group()


def test_sweep_points_are_dealt_over_the_gpus(cuda, tmp_path):
    """four (T, P) points on two GPUs: each rank evaluates two of them and writes their files; every file equals what a
    single process computes for the same seed"""
    if cuda.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from pibronic_b200 import file_structure, pimc, sweep, synthetic
    FS = file_structure.FileStructure(tmp_path, 0, 0)
    synthetic.write_data_set(FS, synthetic.coupled_model(2, 2, (0.02, 0.04), (0.1, 0.2), seed=3, linear=0.05))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_sweep_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    mine = [tuple(p) for r in range(2) for p in np.load(join(tmp_path, f"points_{r}.npy"))]
    assert sorted(mine) == [(8.0, 250.0), (8.0, 300.0), (12.0, 250.0), (12.0, 300.0)]
    assert len(np.load(join(tmp_path, "points_0.npy"))) == 2
    two_gpu = {}
    for P in (8, 12):
        for T in (250.0, 300.0):
            res = pimc.BoxResultPM()
            res.load_multiple_results([join(FS.path_rho_results, f"P{P}_T{T:.2f}_J0_data_points.npz")])
            two_gpu[(P, T)] = res.scaled_g.copy()
    single = sweep.run_sweep(FS, {"temperature_list": [250.0, 300.0], "bead_list": [8, 12], "number_of_samples": 4000,
                                  "block_size": 1000, "seed": 5})
    for key, res in single.items():
        assert np.array_equal(res.scaled_g, two_gpu[key]), key
