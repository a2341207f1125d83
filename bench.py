"""Benchmark of the PIMC estimator hot path (BASELINE.json metric: PIMC samples*beads/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config c2 of BASELINE.json): synthetic A=4 surfaces, N=6 modes, P=64 beads, linear +
quadratic coupling, T=300 K, X = 1e6 samples per GPU per step in blocks of 1e4, PM path
(rho, g, g+, g-).  One step = one fused sampler+estimator launch over X samples (warp-specialised
kernel + its redo pass) + the per-block sums (+ one NCCL all-reduce of the block sums when N > 1).
Weak scaling: every GPU does X samples.

value   : device-resident throughput, CUDA events around each step, max over ranks.
e2e     : the same step through the host-buffer C ABI call (pbx_sample_eval_host) with pinned host
          buffers: the 32 MB of results reach the host inside the timed region (the kernel stores them
          into the mapped buffer as it goes; the block sums are copied after it).
roofline: FP64 vector pipe.  achieved = algorithmic flop/sample (with the
          sampler term counted for the O(P) recurrence actually used) x samples/s;
          peak = FP64 FMA rate measured in this run, the larger of a vector DFMA-chain probe and an
          FP64 tensor (DMMA) probe (MEASURED_PEAKS.json has no FP64 entry).
cpu_baseline / --impl reference: the numpy port of the reference's block loop (oracle/, "port")
          on a bounded sample, one process per host core with single-threaded BLAS.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from os.path import abspath, dirname

ROOT = dirname(abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

A, N, P, T_KELVIN = 4, 6, 64, 300.0
X_PER_GPU = 1_000_000
BLOCK_SIZE = 10_000
CPU_SAMPLES_PER_PROC = 4000   # bounded CPU sample: about 3 s per process and step
CPU_BLOCK = 1000
NCU_DRAM_BYTES_PER_LAUNCH = 80640 + 194816   # from the committed ncu capture of the fused kernel, 1e6 samples per launch
WORKLOAD = (f"c2: synthetic A={A} N={N} P={P} lin+quad coupling, T={T_KELVIN:.0f}K, PM path, "
            f"X={X_PER_GPU:.0e} samples/GPU/step in blocks of {BLOCK_SIZE}")


def algorithmic_flops_per_sample(A, N, P, Ar):
    """SURVEY.md section 8(d) strict count (transcendental = 1 flop).  F_x is the sequential ring
    recurrence actually used (1 mul + 2 fma per coordinate), not the reference's dense PxP product."""
    nn, aa = N * (N + 1) // 2, A * (A + 1) // 2
    F_x = 5 * N * P
    F_O = N * P * (20 * A + 8 * Ar)
    F_V = P * (nn + 2 * nn * aa + 2 * N * A * (A - 1) // 2 + aa)
    F_eig = 9 * A ** 3 * P
    F_M = P * (A ** 3 + 2 * A ** 2)
    F_chain = 3 * P * (2 * A ** 3 + A ** 2)
    F_rho = P * Ar + Ar
    N_exp = P * (3 * A + Ar) + P * A
    return F_x + F_O + F_V + F_eig + F_M + F_chain + F_rho + N_exp


# ----------------------------------------------------------------------------- CPU (reference arm)
def _cpu_worker(args):
    X, B, seed = args[:3]
    blas_threads = args[3] if len(args) > 3 else 1      # None: whatever the BLAS picks (the reference's single-job mode)
    from threadpoolctl import threadpool_limits
    from oracle import pimc_oracle as orc
    from pibronic_b200 import synthetic
    from pibronic_b200.model_io import VMK
    with threadpool_limits(limits=blas_threads):
        model = synthetic.model_c2()
        rho = synthetic.diagonal_of(model)
        vib_d = dict(A=A, N=N, E=model[VMK.E], w=model[VMK.w], L=model[VMK.G1], Q=model[VMK.G2])
        rho_d = dict(A=A, N=N, E=rho[VMK.E], w=rho[VMK.w], L=rho[VMK.G1])
        tab = orc.precompute(vib_d, rho_d, P, T_KELVIN)
        rng = np.random.RandomState(seed)
        t0 = time.perf_counter()
        out = orc.run_blocks(tab, X, B, rng, pm=True, faithful=True)
        dt = time.perf_counter() - t0
    ratio = out[1] / out[0]
    return dt, float(ratio.mean())


def cpu_port_rate(procs, X=CPU_SAMPLES_PER_PROC, B=CPU_BLOCK, seed=0):
    """aggregate samples*beads/s of `procs` independent shards of the numpy port; the elapsed time is
    the slowest shard's block-loop time (setup excluded, like the GPU arm)"""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(X, B, seed + 1000 * i) for i in range(procs)])
    slowest = max(r[0] for r in res)
    return procs * X * P / slowest, slowest, float(np.mean([r[1] for r in res]))


def run_reference(args, rank):
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    sample = (f"{procs} processes x {CPU_SAMPLES_PER_PROC} samples (blocks of {CPU_BLOCK}) of the same workload per step; "
              "numpy port of block_compute_pm, 1 BLAS thread per process")
    for _ in range(args.warmup):
        cpu_port_rate(procs, X=200, B=100)
    times, rates = [], []
    for k in range(args.steps):
        rate, dt, _ = cpu_port_rate(procs, seed=k)
        times.append(dt)
        rates.append(rate)
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": "PIMC samples*beads/sec", "value": value, "unit": "samples*beads/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "A": A, "N": N, "P": P, "cpu_sample_per_step": procs * CPU_SAMPLES_PER_PROC},
        "cpu_baseline": {"value": value, "unit": "samples*beads/s", "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples*beads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- GPU arm
def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from pibronic_b200 import _cabi, constants, synthetic
    from pibronic_b200.model_io import VMK

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    model = synthetic.model_c2()
    rho = synthetic.diagonal_of(model)
    plan = _cabi.Plan(model[VMK.E], model[VMK.w], model[VMK.G1], model[VMK.G2], rho[VMK.E], rho[VMK.w], rho[VMK.G1],
                      P, constants.beta(T_KELVIN), constants.delta_beta, flags=_cabi.FLAG_PM, device=local_rank)
    assert plan.is_fast, "the c2 shape must run on the register-resident kernel"
    X, blocks = X_PER_GPU, X_PER_GPU // BLOCK_SIZE
    out = torch.empty((4, X), dtype=torch.float64, device="cuda")
    sums = torch.empty((blocks, _cabi.NSUMS), dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")  # > 126 MB L2
    seed0 = 20260417

    def step(k):
        plan.sample_eval(seed0 + k, rank * X, X, out)       # Philox counter = global sample index
        plan.block_sums(out, X, BLOCK_SIZE, sums)
        if world > 1:
            dist.all_reduce(sums)                            # the path's one exchange step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak_vector = _cabi.fp64_peak_tflops(local_rank, 0)
    peak_tensor = _cabi.fp64_peak_tflops(local_rank, 1)
    peak = max(peak_vector, peak_tensor)   # one set of FP64 units behind both paths: the higher reading is the roofline
    for k in range(args.warmup):
        step(k)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = plan.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        flush.fill_(float(k))                                # L2 flush between timed iterations (untimed)
        ev[k][0].record()
        plan.sample_eval(seed0 + args.warmup + k, rank * X, X, out)
        ev[k][1].record()
        plan.block_sums(out, X, BLOCK_SIZE, sums)
        if world > 1:
            dist.all_reduce(sums)
        ev[k][2].record()
    barrier()
    gpu_launches = plan.launch_count - launches0
    step_ms = np.array([e[0].elapsed_time(e[2]) for e in ev])
    kern_ms = np.array([e[0].elapsed_time(e[1]) for e in ev])
    clock_info = clocks.stop() if rank == 0 else None
    total_ms = torch.tensor([step_ms.sum(), kern_ms.sum()], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms, kern_total_ms = (float(v) for v in total_ms.cpu())
    value = world * X * P * args.steps / (total_ms * 1e-3)

    # sanity of the numbers produced inside the timed region
    host = out.cpu().numpy()
    ratio = host[1] / host[0]
    assert np.all(np.isfinite(host)) and np.all(host[0] > 0), "non-finite results"

    # ---- e2e: host-buffer C ABI call, D2H of results + block sums inside the timed region
    pinned = torch.empty((4, X), dtype=torch.float64, pin_memory=True).numpy()
    e2e_steps = max(3, min(args.steps, 10))
    plan.sample_eval_host(seed0, rank * X, X, out4=pinned, block_size=BLOCK_SIZE)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        plan.sample_eval_host(seed0 + 100 + k, rank * X, X, out4=pinned, block_size=BLOCK_SIZE)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * X * P * e2e_steps / float(e2e_s.cpu())
    # the model tables travel to the device as the kernel's __grid_constant__ parameter on every launch
    table_bytes = plan.launch_param_bytes

    if rank == 0:
        flops = algorithmic_flops_per_sample(A, N, P, A)
        samples_per_s_kernel = X * args.steps / (kern_total_ms * 1e-3)  # this rank's kernel-only rate
        achieved = flops * samples_per_s_kernel / 1e12
        cpu_procs = os.cpu_count() or 1
        cpu_rate, cpu_dt, cpu_mean = (None, None, None)
        single_rate = None
        if world == 1 and not args.skip_cpu:
            cpu_rate, cpu_dt, cpu_mean = cpu_port_rate(cpu_procs)
            # the reference's actual mode: ONE process (sbatch --ntasks=1, job_boss.py:314), BLAS threads left alone
            dt1, _ = _cpu_worker((2000, 1000, 7, None))
            single_rate = 2000 * P / dt1
        line = {
            "metric": "PIMC samples*beads/sec", "value": value, "unit": "samples*beads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "A": A, "N": N, "P": P, "A_rho": A, "samples_per_gpu_per_step": X,
                       "block_size": BLOCK_SIZE, "parallelism": f"samples sharded x{world}, NCCL all-reduce of block sums",
                       "l2": "256 MB flush write between timed steps; the step has no HBM-resident inputs "
                             "(coordinates are generated on-chip), a fresh Philox seed per step"},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                         "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch "
                                           "(profiles/r01_ncu_full_pbx_fast_ws_kernel.csv); the 32 MB of results stay in L2",
                         "peak_source": "measured in this run (MEASURED_PEAKS.json has no FP64 entry): the larger of a vector "
                                        "DFMA-chain probe and an FP64 tensor (mma.sync m8n8k4) probe, which share the FP64 units "
                                        "on B200; nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2",
                         "peak_vector_dfma": peak_vector, "peak_tensor_dmma": peak_tensor,
                         "frac_of_vector_peak": achieved / peak_vector,
                         "kernel": "pbx_fast_ws_kernel<4,6,4,PM,shared-rho> (+ its MODE_REDO pass, ~10 us, inside kernel_ms)",
                         "kernel_ms": kern_total_ms / args.steps,
                         "flop_per_sample": flops, "algorithmic_bytes_per_sample": 32},
            "cpu_baseline": {"value": cpu_rate, "unit": "samples*beads/s", "cores": cpu_procs, "kind": "port",
                             "sample": f"{cpu_procs} processes x {CPU_SAMPLES_PER_PROC} samples, numpy port of "
                                       f"block_compute_pm (oracle/pimc_oracle.py), 1 BLAS thread each, slowest shard "
                                       f"{cpu_dt if cpu_dt is None else round(cpu_dt, 2)} s",
                             "single_process_value": single_rate,
                             "single_process_sample": "1 process x 2000 samples, default BLAS threads (the reference's one-job mode)"},
            "e2e": {"value": e2e_value, "unit": "samples*beads/s", "h2d_bytes_per_step": table_bytes,
                    "d2h_bytes_per_step": 4 * X * 8 + blocks * _cabi.NSUMS * 8, "steps": e2e_steps,
                    "call": "pbx_sample_eval_host (pinned, mapped host buffers: results written by the kernel, sums copied)"},
            "gpu_launches": int(gpu_launches),
            "clocks": clock_info,
            "check": {"mean_g_over_rho": float(ratio.mean()), "stderr": float(ratio.std() / np.sqrt(X)),
                      "cpu_port_mean_g_over_rho": cpu_mean},
        }
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    args.warmup = max(args.warmup, 3)
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
