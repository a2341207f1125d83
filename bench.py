"""Benchmark of the PIMC estimator hot path (BASELINE.json metric: PIMC samples*beads/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (config c2 of BASELINE.json): synthetic A=4 surfaces, N=6 modes, P=64 beads, linear +
quadratic coupling, T=300 K, X = 1e6 samples per GPU per step in blocks of 1e4, PM path (rho, g, g+, g-).
One step = one fused sampler+estimator launch over X samples (warp-specialised kernel + its redo pass) + the
per-block sums; with N > 1 the block sums of a step are all-reduced (NCCL) on a side stream while the next step
runs -- one exchange per batch, none on the data path.  Weak scaling: every GPU does X samples.

value        : device-resident throughput; CUDA events around the K steps (all-reduces drained inside), max over ranks.
e2e          : the same step through the host-buffer C ABI call (pbx_sample_eval_host) with pinned host buffers: the
               32 MB of results reach the host inside the timed region (the kernel stores them into the mapped buffer
               as it goes; the block sums are copied after it) + the all-reduce of the block sums when N > 1.
e2e_facade   : the user-facing call, pibronic_b200.pimc.block_compute_pm (plan lookup, result arrays, .npz written).
roofline     : FP64 units.  achieved = algorithmic flop/sample (with the sampler term counted for
               the O(P) recurrence actually used) x samples/s; peak = FP64 FMA rate measured in this run, the larger of
               a vector DFMA-chain probe and an FP64 tensor (DMMA) probe (MEASURED_PEAKS.json has no FP64 entry);
               traffic = DRAM bytes per launch read from the committed ncu capture of the same kernel (profiles/).
other_configs: every other BASELINE.json configuration at its stated size (c1 X=1e4, c3 P=128, c4 X=1e7 over the GPUs
               of the run, c5 = c2 sampled from another rho), each with ms, samples*beads/s and its roofline fraction
               under the same flop rule.
coords_mode  : pbx_eval_coords on 1e6 device-resident c2 samples per GPU (the entry the parity tests and the
               *_from_raw_samples block drivers use): the only configuration with inputs in HBM -- achieved GB/s of the
               co-ordinate array against MEASURED_PEAKS.json's copy bandwidth next to the FP64 fraction (no sampler term).
strong_scaling: c2 at a FIXED total of 1e6 samples and c4 at a fixed total of 1e7, split over the N GPUs.
cpu_baseline / --impl reference: the UNMODIFIED reference's block_compute_pm (staged into baseline/_ref by
               baseline/stage_reference.py; kind "reference") on a bounded sample, one process per host core with
               single-threaded BLAS; the numpy port (oracle/, kind "port") is reported beside it, and stands in when the
               staged reference is absent.
"""
import argparse
import csv
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time
from os.path import abspath, dirname, isfile, join

ROOT = dirname(abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

A, N, P, T_KELVIN = 4, 6, 64, 300.0
X_PER_GPU = 1_000_000
BLOCK_SIZE = 10_000
CPU_SAMPLES_PER_PROC = 4000   # bounded CPU sample: a few seconds per process and step
CPU_BLOCK = 1000
C4_TOTAL = 10_000_000         # BASELINE configs[3]: X = 1e7 over the GPUs of the run
C2_STRONG_TOTAL = 1_000_000
NCU_CSV = join(ROOT, "profiles", "r02_ncu_full_pbx_fast_ws_kernel.csv")
WORKLOAD = (f"c2: synthetic A={A} N={N} P={P} lin+quad coupling, T={T_KELVIN:.0f}K, PM path, "
            f"X={X_PER_GPU:.0e} samples/GPU/step in blocks of {BLOCK_SIZE}")


def algorithmic_flops_per_sample(A, N, P, Ar, sampler=True):
    """SURVEY.md section 8(d) strict count (transcendental = 1 flop).  F_x is the sequential ring
    recurrence actually used (1 mul + 2 fma per coordinate), not the reference's dense PxP product
    (sampler=False: co-ordinates supplied by the caller, no F_x).
    The ONE flop rule of this repository: every roofline fraction (bench, DESIGN.md, profiles/) uses it."""
    nn, aa = N * (N + 1) // 2, A * (A + 1) // 2
    F_x = 5 * N * P if sampler else 0
    F_O = N * P * (20 * A + 8 * Ar)
    F_V = P * (nn + 2 * nn * aa + 2 * N * A * (A - 1) // 2 + aa)
    F_eig = 9 * A ** 3 * P
    F_M = P * (A ** 3 + 2 * A ** 2)
    F_chain = 3 * P * (2 * A ** 3 + A ** 2)
    F_rho = P * Ar + Ar
    N_exp = P * (3 * A + Ar) + P * A
    return F_x + F_O + F_V + F_eig + F_M + F_chain + F_rho + N_exp


def ncu_dram_bytes_per_launch(path=NCU_CSV):
    """dram__bytes_read.sum + dram__bytes_write.sum of the bench kernel from the committed ncu summary (tools/ncu_summary.py)"""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    if not isfile(path):
        return None
    total = 0.0
    with open(path) as fh:
        for row in csv.reader(fh):
            if len(row) == 3 and row[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += float(row[2]) * scale.get(row[1], 1.0)
    return total


# ----------------------------------------------------------------------------- CPU (reference arm)
def _c2_models():
    from pibronic_b200 import synthetic
    model = synthetic.model_c2()
    return model, synthetic.diagonal_of(model)


def _port_worker(args):
    X, B, seed = args[:3]
    blas_threads = args[3] if len(args) > 3 else 1      # None: whatever the BLAS picks (the reference's single-job mode)
    from threadpoolctl import threadpool_limits
    from oracle import pimc_oracle as orc
    from pibronic_b200.model_io import VMK
    with threadpool_limits(limits=blas_threads):
        model, rho = _c2_models()
        vib_d = dict(A=A, N=N, E=model[VMK.E], w=model[VMK.w], L=model[VMK.G1], Q=model[VMK.G2])
        rho_d = dict(A=A, N=N, E=rho[VMK.E], w=rho[VMK.w], L=rho[VMK.G1])
        tab = orc.precompute(vib_d, rho_d, P, T_KELVIN)
        rng = np.random.RandomState(seed)
        t0 = time.perf_counter()
        out = orc.run_blocks(tab, X, B, rng, pm=True, faithful=True)
        dt = time.perf_counter() - t0
    ratio = out[1] / out[0]
    return dt, float(ratio.mean())


def _reference_worker(args):
    """block_compute_pm of the staged, unmodified reference on the c2 model files"""
    X, B, seed = args[:3]
    blas_threads = args[3] if len(args) > 3 else 1
    from threadpoolctl import threadpool_limits
    from baseline import ref_runner
    from pibronic_b200 import model_io as vIO
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        model, rho = _c2_models()
        path_vib, path_rho = join(tmp, "coupled_model.json"), join(tmp, "sampling_model.json")
        vIO.save_model_to_JSON(path_vib, model)
        vIO.save_diagonal_model_to_JSON(path_rho, rho)
        with threadpool_limits(limits=blas_threads):
            dt, result = ref_runner.run_block_compute_pm(path_vib, path_rho, P, T_KELVIN, X, B, seed)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return dt, float((result.scaled_g / result.scaled_rho).mean())


def reference_available():
    from baseline import ref_runner
    return ref_runner.available()


def cpu_rate(kind, procs, X=CPU_SAMPLES_PER_PROC, B=CPU_BLOCK, seed=0, pool=None):
    """aggregate samples*beads/s of `procs` independent shards of the CPU path (kind "reference" or "port"); the
    elapsed time is the slowest shard's block-loop time (setup excluded, like the GPU arm)"""
    import multiprocessing as mp
    worker = _reference_worker if kind == "reference" else _port_worker
    jobs = [(X, B, seed + 1000 * i) for i in range(procs)]
    if pool is not None:
        res = pool.map(worker, jobs, chunksize=1)
    else:
        with mp.get_context("spawn").Pool(procs) as own:
            res = own.map(worker, jobs, chunksize=1)
    slowest = max(r[0] for r in res)
    return procs * X * P / slowest, slowest, float(np.mean([r[1] for r in res]))


def run_reference(args, rank):
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    kind = "reference" if reference_available() else "port"
    what = ("block_compute_pm of the unmodified reference (baseline/_ref, pibronic/pimc/pimc.py:1388-1462)" if kind == "reference"
            else "numpy port of block_compute_pm (oracle/pimc_oracle.py; the staged reference is absent)")
    sample = (f"{procs} processes x {CPU_SAMPLES_PER_PROC} samples (blocks of {CPU_BLOCK}) of the same workload per step; "
              f"{what}, 1 BLAS thread per process")
    import multiprocessing as mp
    times, rates = [], []
    with mp.get_context("spawn").Pool(procs) as pool:      # one pool for the whole run: the imports happen once per worker
        for _ in range(args.warmup):
            cpu_rate(kind, procs, X=200, B=100, pool=pool)
        for k in range(args.steps):
            rate, dt, _ = cpu_rate(kind, procs, seed=k, pool=pool)
            times.append(dt)
            rates.append(rate)
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": "PIMC samples*beads/sec", "value": value, "unit": "samples*beads/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "A": A, "N": N, "P": P, "cpu_sample_per_step": procs * CPU_SAMPLES_PER_PROC},
        "cpu_baseline": {"value": value, "unit": "samples*beads/s", "cores": procs, "kind": kind, "sample": sample},
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
def make_plan(model, rho, beads, device, flags=None):
    from pibronic_b200 import _cabi, constants
    from pibronic_b200.model_io import VMK
    return _cabi.Plan(model[VMK.E], model[VMK.w], model[VMK.G1], model.get(VMK.G2), rho[VMK.E], rho[VMK.w], rho[VMK.G1],
                      beads, constants.beta(T_KELVIN), constants.delta_beta,
                      flags=_cabi.FLAG_PM if flags is None else flags, device=device)


class Runner:
    """timing helpers shared by the headline and the side measurements of one rank"""

    def __init__(self, torch, dist, rank, world, peak):
        self.torch, self.dist, self.rank, self.world, self.peak = torch, dist, rank, world, peak
        self.flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")   # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    def time_fused(self, plan, x_rank, reps, seed, block_size=None, first=None):
        """`reps` fused launches (+ block sums) of x_rank samples on this rank; returns (ms per launch as the max over
        ranks, device output of the last launch)"""
        torch = self.torch
        from pibronic_b200 import _cabi
        out = torch.empty((4, x_rank), dtype=torch.float64, device="cuda")
        first = self.rank * x_rank if first is None else first
        sums = None
        if block_size:
            sums = torch.empty((-(-x_rank // block_size), _cabi.NSUMS), dtype=torch.float64, device="cuda")
        plan.sample_eval(seed, first, min(x_rank, 4096), out)          # warm-up: module load, table upload
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total = 0.0
        for k in range(reps):
            self.flush.fill_(float(k))
            e0.record()
            plan.sample_eval(seed + 1 + k, first, x_rank, out)
            if sums is not None:
                plan.block_sums(out, x_rank, block_size, sums)
                if self.world > 1:
                    self.dist.all_reduce(sums)
            e1.record()
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1)
        (ms,) = self.max_over_ranks([total / reps])
        return ms, out

    def time_fused_pipelined(self, plan, x_rank, reps, seed, block_size):
        """the headline's way of timing, for any sample count: per-step events around fused launch + block sums, the
        all-reduce of step k asynchronous on NCCL's stream under step k+1, the drain after the last step added in full;
        returns ms per step (max over ranks)"""
        torch = self.torch
        from pibronic_b200 import _cabi
        out = torch.empty((4, x_rank), dtype=torch.float64, device="cuda")
        sums = [torch.empty((-(-x_rank // block_size), _cabi.NSUMS), dtype=torch.float64, device="cuda") for _ in range(2)]
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(reps)]
        t_end = torch.cuda.Event(enable_timing=True)
        plan.sample_eval(seed, self.rank * x_rank, min(x_rank, 4096), out)
        pending = []
        self.barrier()
        for k in range(reps):
            self.flush.fill_(float(k))
            ev[k][0].record()
            plan.sample_eval(seed + 1 + k, self.rank * x_rank, x_rank, out)
            plan.block_sums(out, x_rank, block_size, sums[k % 2])
            ev[k][1].record()
            if self.world > 1:
                pending.append(self.dist.all_reduce(sums[k % 2], async_op=True))
                while len(pending) > 1:
                    pending.pop(0).wait()
        for w in pending:
            w.wait()
        t_end.record()
        torch.cuda.synchronize()
        total = sum(e[0].elapsed_time(e[1]) for e in ev) + ev[-1][1].elapsed_time(t_end)
        (ms,) = self.max_over_ranks([total / reps])
        return ms

    def config_line(self, name, shape, beads, x_total, ms, extra=None):
        a, n, ar = shape
        flops = algorithmic_flops_per_sample(a, n, beads, ar)
        rate = x_total / (ms * 1e-3)
        line = {"config": name, "A": a, "N": n, "A_rho": ar, "P": beads, "samples_total": x_total, "n_gpus": self.world,
                "ms": ms, "samples_beads_per_s": rate * beads, "flop_per_sample": flops,
                "tflops": flops * rate / 1e12, "frac_of_fp64_peak": flops * rate / 1e12 / (self.peak * self.world)}
        if extra:
            line.update(extra)
        return line


def other_configs(run, device):
    """BASELINE.json configs other than the headline one, at their stated sizes, sharded over the ranks of the run"""
    from pibronic_b200 import _cabi, synthetic
    world = run.world
    lines = []

    def go(name, model, rho, beads, x_total, reps, note):
        x_rank = x_total // world
        plan = make_plan(model, rho, beads, device)
        ms, out = run.time_fused(plan, x_rank, reps, seed=7)
        host = out[:2, :min(x_rank, 100_000)].cpu().numpy()
        assert np.all(np.isfinite(host)) and np.all(host[0] > 0), name
        path = {_cabi.PATH_REGISTER: "register-resident (pbx_fast_ws_kernel)", _cabi.PATH_FUSED_DMMA: "fused tensor-core (pbx_big_kernel)",
                _cabi.PATH_BLOCKED: "blocked", _cabi.PATH_GENERIC: "generic"}[plan.kernel_path]
        shape = (plan.A, plan.N, plan.Ar)
        plan.close()
        lines.append(run.config_line(name, shape, beads, x_rank * world, ms, {"kernel": path, "launches_timed": reps, "note": note}))

    c1 = synthetic.coupled_model(2, 2, (0.01, 0.02), (0.0, 0.1), seed=1, linear=0.05, quadratic=0.0, mixing=0.0)
    go("c1", c1, synthetic.diagonal_of(c1), 12, 10_000 * world, 20,
       "examples/artificial_systems-like 2x2 model, P=12, X=1e4 per GPU: one ~30 us launch, launch-latency bound")
    c3 = synthetic.coupled_model(2, 2, (0.02, 0.04), (0.1, 0.2), seed=3, linear=0.05, quadratic=0.0)
    go("c3", c3, synthetic.diagonal_of(c3), 128, 100_000 * world, 10,
       "paper_1.5025058-like 2x2 model, P=128, X=1e5 per shard (the reference's shard size)")
    go("c3_x1e6", c3, synthetic.diagonal_of(c3), 128, 1_000_000 * world, 5,
       "the same 2x2 model and P=128 with 1e6 samples per launch (26 waves instead of 2.6: what a GPU shard would use)")
    c4 = synthetic.model_c4()
    go("c4", c4, synthetic.diagonal_of(c4), 256, C4_TOTAL, 1,
       "A=12 N=24 P=256, X=1e7 TOTAL over the GPUs of this run (strong scaling: the stated configuration at N=8)")
    c2 = synthetic.model_c2()
    free = synthetic.coupled_model(4, 6, (0.14, 0.45), (10.3, 10.9), mixing=0.0, quadratic=0.0)
    go("c5", c2, synthetic.diagonal_of(free), 64, 1_000_000 * world, 5,
       "c2 sampled from another rho (un-rotated diagonal model: no shared exponents), X=1e6 per GPU")
    return lines


def coords_mode(run, plan, x=1_000_000, reps=5):
    """the other entry of the path: pbx_eval_coords on device-resident co-ordinates R[x][N][P] (what the parity tests and
    the block drivers *_from_raw_samples use).  The only configuration with HBM-resident inputs: 8 N P + 32 bytes per sample."""
    torch = run.torch
    R = torch.empty((x, plan.N, plan.P), dtype=torch.float64, device="cuda")
    out = torch.empty((4, x), dtype=torch.float64, device="cuda")
    plan.sample_coords(11, run.rank * x, x, R)
    plan.eval_coords(R[:4096], out[:, :4096].contiguous())
    run.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total = 0.0
    for k in range(reps):
        run.flush.fill_(float(k))
        e0.record()
        plan.eval_coords(R, out)
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    (ms,) = run.max_over_ranks([total / reps])
    fused = torch.empty_like(out)
    plan.sample_eval(11, run.rank * x, x, fused)
    same = bool(torch.equal(fused, out))
    assert same, "co-ordinate mode and the fused sampler disagree on the same Philox counters"
    flops = algorithmic_flops_per_sample(plan.A, plan.N, plan.P, plan.Ar, sampler=False)
    rate = x * run.world / (ms * 1e-3)
    hbm_peak = None
    try:
        with open(join(ROOT, "MEASURED_PEAKS.json")) as fh:
            hbm_peak = float(json.load(fh)["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        pass
    gbs = (8 * plan.N * plan.P + 32) * rate / run.world / 1e9
    del R
    return {"call": "pbx_eval_coords_dev on device-resident R[x][N][P] (c2 model), read in place by the register-resident kernel",
            "samples_per_gpu": x, "n_gpus": run.world, "ms": ms, "samples_beads_per_s": rate * plan.P,
            "flop_per_sample": flops, "tflops": flops * rate / 1e12, "frac_of_fp64_peak": flops * rate / 1e12 / (run.peak * run.world),
            "algorithmic_bytes_per_sample": 8 * plan.N * plan.P + 32, "hbm_gbs_per_gpu": gbs, "hbm_peak_gbs": hbm_peak,
            "frac_of_hbm_peak": (gbs / hbm_peak) if hbm_peak else None,
            "bit_identical_to_fused_sampler": same,
            "note": "FP64 bound, not HBM bound: the estimator alone (no sampler warps competing for the FP64 units)"}


def shard_check(run, plan):
    """the union of the ranks' shards is bit-identical to one GPU evaluating the whole index range"""
    torch, dist = run.torch, run.dist
    m = 8192
    mine = torch.empty((4, m), dtype=torch.float64, device="cuda")
    plan.sample_eval(99, run.rank * m, m, mine)
    if run.world == 1:
        return "single GPU"
    gathered = [torch.empty_like(mine) for _ in range(run.world)] if run.rank == 0 else None
    dist.gather(mine, gathered, dst=0)
    if run.rank != 0:
        return None
    whole = torch.empty((4, m * run.world), dtype=torch.float64, device="cuda")
    plan.sample_eval(99, 0, m * run.world, whole)
    torch.cuda.synchronize()
    ok = bool(torch.equal(torch.cat(gathered, dim=1), whole))
    assert ok, "rank shards differ from the single-GPU evaluation of the same index range"
    return f"{run.world} shards of {m} samples bit-identical to one GPU evaluating the whole range"


def facade_e2e(run, device, reps=3):
    """pibronic_b200.pimc.block_compute_pm (the call a user of the reference makes) incl. the .npz; samples*beads/s"""
    from pibronic_b200 import file_structure, pimc, synthetic
    tmp = tempfile.mkdtemp(prefix=f"pbx_bench_r{run.rank}_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        FS = file_structure.FileStructure(tmp, 0, 0)
        model = synthetic.model_c2()
        synthetic.write_data_set(FS, model, synthetic.diagonal_of(model))
        FS.generate_model_hashes()
        data = pimc.BoxDataPM.from_FileStructure(FS)
        data.samples, data.beads, data.temperature, data.block_size = X_PER_GPU, P, T_KELVIN, BLOCK_SIZE
        data.blocks = data.samples // data.block_size
        data.hash_vib, data.hash_rho = FS.hash_vib, FS.hash_rho
        data.seed, data.sample_offset = 20260417, run.rank * X_PER_GPU
        data.preprocess()
        result = pimc.BoxResultPM(data=data)
        result.path_root, result.id_job = FS.path_rho_results, run.rank
        pimc.block_compute_pm(data, result)               # warm-up (plan creation)
        run.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            pimc.block_compute_pm(data, result)
        dt = time.perf_counter() - t0
        npz = os.path.getsize(result.compute_path_to_file() if hasattr(result, "compute_path_to_file") else tmp)
        data.release()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    (dt,) = run.max_over_ranks([dt])
    return {"value": run.world * X_PER_GPU * P * reps / dt, "unit": "samples*beads/s", "ms_per_call": 1e3 * dt / reps,
            "call": "pibronic_b200.pimc.block_compute_pm(BoxDataPM, BoxResultPM): plan lookup, fused launch into pinned result "
                    "arrays, block sums, the four arrays written as the reference's .npz (tmpfs)", "npz_bytes": int(npz), "calls": reps}


def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from pibronic_b200 import _cabi, synthetic

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    model = synthetic.model_c2()
    rho = synthetic.diagonal_of(model)
    plan = make_plan(model, rho, P, local_rank)
    assert plan.is_fast, "the c2 shape must run on the register-resident kernel"
    X, blocks = X_PER_GPU, X_PER_GPU // BLOCK_SIZE
    out = torch.empty((4, X), dtype=torch.float64, device="cuda")
    sums = [torch.empty((blocks, _cabi.NSUMS), dtype=torch.float64, device="cuda") for _ in range(2)]
    seed0 = 20260417

    peak_vector = _cabi.fp64_peak_tflops(local_rank, 0)
    peak_tensor = _cabi.fp64_peak_tflops(local_rank, 1)
    peak = max(peak_vector, peak_tensor)   # one set of FP64 units behind both paths: the higher reading is the roofline
    run = Runner(torch, dist, rank, world, peak)

    def step(k, pending):
        plan.sample_eval(seed0 + k, rank * X, X, out)       # Philox counter = global sample index
        plan.block_sums(out, X, BLOCK_SIZE, sums[k % 2])
        if world > 1:                                        # the path's one exchange, off the critical path
            pending.append(dist.all_reduce(sums[k % 2], async_op=True))
            while len(pending) > 1:                          # sums[k % 2] is reused two steps later
                pending.pop(0).wait()

    pending = []
    for k in range(args.warmup):
        run.flush.fill_(float(k))
        step(k, pending)
    for w in pending:
        w.wait()
    run.barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = plan.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_end = torch.cuda.Event(enable_timing=True)
    pending = []
    run.barrier()
    for k in range(args.steps):
        run.flush.fill_(float(k))                            # L2 flush between timed iterations (untimed)
        ev[k][0].record()
        plan.sample_eval(seed0 + args.warmup + k, rank * X, X, out)
        ev[k][1].record()
        plan.block_sums(out, X, BLOCK_SIZE, sums[k % 2])
        ev[k][2].record()
        if world > 1:
            pending.append(dist.all_reduce(sums[k % 2], async_op=True))
            while len(pending) > 1:
                pending.pop(0).wait()
    for w in pending:
        w.wait()
    t_end.record()
    run.barrier()
    gpu_launches = plan.launch_count - launches0
    step_ms = np.array([e[0].elapsed_time(e[2]) for e in ev])
    kern_ms = np.array([e[0].elapsed_time(e[1]) for e in ev])
    clock_info = clocks.stop() if rank == 0 else None
    # timed region of a step: fused kernel + block sums (events on the launching stream); the all-reduce of step k runs
    # on NCCL's stream under step k+1 (whatever it costs the kernels shows in their events), the drain of the last ones
    # after the last step is added in full
    drain_ms = ev[-1][2].elapsed_time(t_end)
    total_ms, kern_total_ms, steps_ms = run.max_over_ranks([step_ms.sum() + drain_ms, kern_ms.sum(), step_ms.sum()])
    value = world * X * P * args.steps / (total_ms * 1e-3)

    # sanity of the numbers produced inside the timed region
    host = out.cpu().numpy()
    ratio = host[1] / host[0]
    assert np.all(np.isfinite(host)) and np.all(host[0] > 0), "non-finite results"
    shard_note = shard_check(run, plan)

    # ---- e2e: host-buffer C ABI call, D2H of results + block sums (+ their all-reduce) inside the timed region
    pinned = torch.empty((4, X), dtype=torch.float64, pin_memory=True).numpy()
    e2e_steps = max(3, min(args.steps, 10))
    plan.sample_eval_host(seed0, rank * X, X, out4=pinned, block_size=BLOCK_SIZE)
    dev_sums = torch.empty((blocks, _cabi.NSUMS), dtype=torch.float64, device="cuda")
    run.barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        _, host_sums = plan.sample_eval_host(seed0 + 100 + k, rank * X, X, out4=pinned, block_size=BLOCK_SIZE)
        if world > 1:
            dev_sums.copy_(torch.from_numpy(host_sums))
            dist.all_reduce(dev_sums)
            host_sums = dev_sums.cpu().numpy()
    (e2e_s,) = run.max_over_ranks([time.perf_counter() - t0])
    e2e_value = world * X * P * e2e_steps / e2e_s
    table_bytes = plan.launch_param_bytes   # the model tables travel as the kernel's __grid_constant__ parameter on every launch
    facade = facade_e2e(run, local_rank)

    # ---- strong scaling of c2: a fixed total of 1e6 samples split over the ranks
    x_rank = C2_STRONG_TOTAL // world
    strong_ms, _ = run.time_fused(plan, x_rank, 20, seed=31, block_size=BLOCK_SIZE)
    strong_pipe_ms = run.time_fused_pipelined(plan, x_rank, 20, seed=57, block_size=BLOCK_SIZE)
    others = other_configs(run, local_rank)
    coords = coords_mode(run, plan)

    if rank == 0:
        flops = algorithmic_flops_per_sample(A, N, P, A)
        samples_per_s_kernel = X * args.steps / (kern_total_ms * 1e-3)  # this rank's kernel-only rate
        achieved = flops * samples_per_s_kernel / 1e12
        cpu_procs = os.cpu_count() or 1
        cpu = {"value": None, "unit": "samples*beads/s", "cores": cpu_procs}
        cpu_mean = None
        if world == 1 and not args.skip_cpu:
            port_rate, port_dt, cpu_mean = cpu_rate("port", cpu_procs)
            dt1, _ = _port_worker((2000, 1000, 7, None))      # the reference's actual mode: ONE process, BLAS threads left alone
            cpu.update({"port_value": port_rate, "port_sample": f"{cpu_procs} processes x {CPU_SAMPLES_PER_PROC} samples, numpy port "
                        f"(oracle/pimc_oracle.py), 1 BLAS thread each, slowest shard {port_dt:.2f} s",
                        "single_process_port_value": 2000 * P / dt1})
            if reference_available():
                ref_rate, ref_dt, _ = cpu_rate("reference", cpu_procs)
                dt1r, _ = _reference_worker((1000, 500, 7, None))
                cpu.update({"value": ref_rate, "kind": "reference",
                            "sample": f"{cpu_procs} processes x {CPU_SAMPLES_PER_PROC} samples (blocks of {CPU_BLOCK}), block_compute_pm of "
                                      f"the unmodified reference staged in baseline/_ref, 1 BLAS thread each, slowest shard {ref_dt:.2f} s",
                            "single_process_value": 1000 * P / dt1r,
                            "single_process_sample": "1 process x 1000 samples, default BLAS threads (the reference's one-job mode)"})
            else:
                cpu.update({"value": port_rate, "kind": "port", "sample": cpu["port_sample"] + " (baseline/_ref not staged)"})
        c4 = next(c for c in others if c["config"] == "c4")
        line = {
            "metric": "PIMC samples*beads/sec", "value": value, "unit": "samples*beads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "A": A, "N": N, "P": P, "A_rho": A, "samples_per_gpu_per_step": X,
                       "block_size": BLOCK_SIZE,
                       "parallelism": f"samples sharded x{world}; block sums all-reduced (NCCL) on a side stream, one exchange per step",
                       "l2": "256 MB flush write between timed steps (untimed); the step "
                             "has no HBM-resident inputs (coordinates are generated on-chip), a fresh Philox seed per step",
                       "shard_check": shard_note},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": ncu_dram_bytes_per_launch(),
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, read at run time from the "
                                           "committed ncu --set full summary " + os.path.relpath(NCU_CSV, ROOT) +
                                           " (the 32 MB of results stay in the 126 MB L2)",
                         "peak_source": "measured in this run (MEASURED_PEAKS.json has no FP64 entry): the larger of a vector "
                                        "DFMA-chain probe and an FP64 tensor (mma.sync m8n8k4) probe, which share the FP64 units "
                                        "on B200; nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2",
                         "peak_vector_dfma": peak_vector, "peak_tensor_dmma": peak_tensor,
                         "frac_of_vector_peak": achieved / peak_vector,
                         "kernel": "pbx_fast_ws_kernel<4,6,4,PM,shared-rho> (+ its MODE_REDO pass, ~10 us, inside kernel_ms)",
                         "kernel_ms": kern_total_ms / args.steps, "step_ms_kernel_plus_sums": steps_ms / args.steps,
                         "flop_per_sample": flops, "algorithmic_bytes_per_sample": 32},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "samples*beads/s", "h2d_bytes_per_step": table_bytes,
                    "d2h_bytes_per_step": 4 * X * 8 + blocks * _cabi.NSUMS * 8, "steps": e2e_steps,
                    "call": "pbx_sample_eval_host (pinned, mapped host buffers: results written by the kernel, sums copied)"
                            + (" + NCCL all-reduce of the block sums" if world > 1 else "")},
            "e2e_facade": facade,
            "gpu_launches": int(gpu_launches),
            "clocks": clock_info,
            "other_configs": others,
            "coords_mode": coords,
            "strong_scaling": {
                "c2": {"samples_total": x_rank * world, "n_gpus": world, "ms": strong_ms,
                       "samples_beads_per_s": x_rank * world * P / (strong_ms * 1e-3),
                       "step": "fused launch + block sums + all-reduce, samples split evenly over the ranks",
                       "ms_pipelined": strong_pipe_ms,
                       "samples_beads_per_s_pipelined": x_rank * world * P / (strong_pipe_ms * 1e-3),
                       "pipelined": "timed like the headline: the all-reduce of a step runs on NCCL's stream under the next "
                                    "step, the drain of the last one is inside the total"},
                "c4": {"samples_total": c4["samples_total"], "n_gpus": world, "ms": c4["ms"],
                       "samples_beads_per_s": c4["samples_beads_per_s"], "frac_of_fp64_peak": c4["frac_of_fp64_peak"]}},
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
