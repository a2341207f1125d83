"""Generates the golden fixtures under tests/golden/ by RUNNING THE REFERENCE in the build container.

Usage (build container only -- /root/reference does not exist on the GPU box):

    python tests/golden/make_golden.py

What it writes (all committed):

* ``explicit_kat.npz``  -- the reference's own known-answer vectors for this path
  (``/root/reference/tests/pimc/explicit_data/*``), repacked into one file together with the text of
  the two model files.
* ``cases/<name>/{coupled_model.json, sampling_model.json, ref.npz}`` -- for each parity case the
  model pair, the bead coordinates R (X,N,P) the reference drew (captured right after
  ``transform_sampled_coordinates``), the mixture component of each sample, and the reference's
  ``scaled_rho, scaled_g, scaled_gofr_plus, scaled_gofr_minus`` for exactly those coordinates,
  plus the precomputed tables of the reference's BoxDataPM.

The reference is imported read-only with MagicMock stand-ins for third-party modules that are
absent here and never touched by the hot path.
"""
import contextlib
import io
import os
import shutil
import sys
import tempfile
from os.path import abspath, dirname, join
from unittest.mock import MagicMock

import numpy as np

HERE = dirname(abspath(__file__))
REPO = dirname(dirname(HERE))
REFERENCE = "/root/reference"

for name in ['parse', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.lines', 'matplotlib.ticker',
             'matplotlib.gridspec', 'matplotlib.backends', 'matplotlib.backends.backend_pdf',
             'mpl_toolkits', 'mpl_toolkits.mplot3d', 'mpl_toolkits.axes_grid1', 'fortranformat',
             'julia', 'memory_profiler']:
    sys.modules.setdefault(name, MagicMock())
sys.path.insert(0, REFERENCE)
sys.path.insert(0, REPO)

import warnings  # noqa: E402
warnings.filterwarnings("ignore", category=SyntaxWarning)

from pibronic import pimc as ref_pimc  # noqa: E402
from pibronic.vibronic import vIO as ref_vIO, VMK as REF_VMK  # noqa: E402

from pibronic_b200 import synthetic  # noqa: E402
from pibronic_b200.model_io import VMK  # noqa: E402


def to_ref_keys(model):
    return {REF_VMK(k.value): v for k, v in model.items()}


def run_reference(path_vib, path_rho, P, T, X, B, seed):
    """block_compute_pm of the unmodified reference; returns dict of arrays"""
    A, N = ref_vIO.extract_dimensions_of_model(path=path_vib)
    data = ref_pimc.BoxDataPM()
    data.id_data, data.id_rho = 0, 0
    data.path_vib_model, data.path_rho_model = path_vib, path_rho
    data.states, data.modes = A, N
    data.samples, data.beads, data.temperature = X, P, T
    data.block_size, data.blocks = B, X // B
    data.hash_vib = ref_vIO.create_model_hash(path=path_vib)
    data.hash_rho = ref_vIO.create_diagonal_model_hash(path=path_rho)
    np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        data.preprocess()
    captured = []
    original = data.transform_sampled_coordinates

    def capture(view):
        original(view)
        captured.append(data.qTensor[:, 0].copy())
    data.transform_sampled_coordinates = capture

    result = ref_pimc.BoxResultPM(data=data)
    tmp = tempfile.mkdtemp()
    result.path_root, result.id_job = tmp, 0
    ref_pimc.block_compute_pm(data, result)
    shutil.rmtree(tmp)
    vib, rho = data.vib, data.rho
    return dict(
        R=np.concatenate(captured, axis=0), sources=np.asarray(rho.sample_sources),
        s_rho=result.scaled_rho, s_g=result.scaled_g, s_gP=result.scaled_gofr_plus, s_gM=result.scaled_gofr_minus,
        P=P, T=T, X=X, B=B, seed=seed, beta=data.beta, tau=data.tau,
        vib_shift=vib.state_shift, vib_delta=vib.delta_weight,
        vib_coth=vib.const.cothAN[0], vib_csch=vib.const.cschAN[0], vib_pref=vib.const.omatrix_prefactor[0, 0],
        vib_pref_plus=vib.const_plus.omatrix_prefactor[0, 0], vib_pref_minus=vib.const_minus.omatrix_prefactor[0, 0],
        vib_coth_plus=vib.const_plus.cothAN[0], vib_csch_minus=vib.const_minus.cschAN[0],
        rho_shift=rho.state_shift, rho_delta=rho.delta_weight, rho_weight=rho.state_weight,
        rho_coth=rho.const.cothAN[0], rho_csch=rho.const.cschAN[0], rho_pref=rho.const.omatrix_prefactor[0, 0],
        inverse_covariance=rho.inverse_covariance[0], ring_eigvals=data.circulant_eigvals,
    )


def write_case(name, model, rho_model, P, T, X, B, seed):
    case_dir = join(HERE, "cases", name)
    os.makedirs(case_dir, exist_ok=True)
    path_vib = join(case_dir, "coupled_model.json")
    path_rho = join(case_dir, "sampling_model.json")
    if isinstance(model, str):
        shutil.copyfile(model, path_vib)
    else:
        ref_vIO.save_model_to_JSON(path_vib, to_ref_keys(model))
    if rho_model is None:
        ref_vIO.remove_coupling_from_model(path_vib, path_rho)
    elif isinstance(rho_model, str):
        shutil.copyfile(rho_model, path_rho)
    else:
        ref_vIO.save_diagonal_model_to_JSON(path_rho, to_ref_keys(rho_model))
    out = run_reference(path_vib, path_rho, P, T, X, B, seed)
    np.savez_compressed(join(case_dir, "ref.npz"), **out)
    ratio = out["s_g"] / out["s_rho"]
    print(f"{name:14s} X={X:3d} P={P:3d} mean g/rho = {ratio.mean():.6g}  (min {ratio.min():.3g}, max {ratio.max():.3g})")


def pack_explicit_kat():
    src = join(REFERENCE, "tests", "pimc", "explicit_data")
    out = {}
    for fname in sorted(os.listdir(src)):
        stem, ext = os.path.splitext(fname)
        if ext == ".npy":
            out[stem] = np.load(join(src, fname))
        elif ext == ".json":
            with open(join(src, fname), "r", encoding="UTF8") as fh:
                out[stem + "_json"] = np.array(fh.read())
    np.savez_compressed(join(HERE, "explicit_kat.npz"), **out)
    print("explicit_kat.npz:", ", ".join(f"{k}{getattr(v, 'shape', '')}" for k, v in out.items()))


def pack_stats_kat():
    """the reference's basic_jackknife_analysis on arrays the reference itself produced (X=3000, P=7)"""
    from pibronic.stats import stats as ref_stats
    case_dir = join(HERE, "cases", "quad_3x4")
    out = run_reference(join(case_dir, "coupled_model.json"), join(case_dir, "sampling_model.json"),
                        P=7, T=250.0, X=3000, B=1000, seed=21)

    class Holder:
        pass
    res = Holder()
    res.scaled_rho, res.scaled_g, res.scaled_gofr_plus, res.scaled_gofr_minus = (out[k] for k in ("s_rho", "s_g", "s_gP", "s_gM"))
    res.samples = 3000
    analytic = {"E": 0.0123, "Cv": 4.5e-5}
    expected = ref_stats.basic_jackknife_analysis(250.0, res, dict(analytic))
    basic = ref_stats.basic_statistical_analysis(250.0, res, dict(analytic))
    np.savez_compressed(join(HERE, "stats_kat.npz"), s_rho=out["s_rho"], s_g=out["s_g"], s_gP=out["s_gP"], s_gM=out["s_gM"],
                        T=250.0, E_sampling=analytic["E"], Cv_sampling=analytic["Cv"],
                        keys=np.array(sorted(expected)), values=np.array([expected[k] for k in sorted(expected)]),
                        basic_keys=np.array(sorted(basic)), basic_values=np.array([basic[k] for k in sorted(basic)]))
    print("stats_kat.npz:", {k: float(v) for k, v in expected.items()})


def pack_c5_stats():
    """BASELINE.json configs[4]: c2 model sampled from the alternate rho (cases/c5_altrho), P=64, 64 blocks of 100 samples,
    run through the unmodified reference; Z, E, Cv and their jackknife errors from the reference's own
    basic_jackknife_analysis (harmonic contribution of the sampling model left out: E_sampling = Cv_sampling = 0)"""
    from pibronic.stats import stats as ref_stats
    case_dir = join(HERE, "cases", "c5_altrho")
    P, T, blocks, B = 64, 300.0, 64, 100
    out = run_reference(join(case_dir, "coupled_model.json"), join(case_dir, "sampling_model.json"),
                        P=P, T=T, X=blocks * B, B=B, seed=64)

    class Holder:
        pass
    res = Holder()
    res.scaled_rho, res.scaled_g, res.scaled_gofr_plus, res.scaled_gofr_minus = (out[k] for k in ("s_rho", "s_g", "s_gP", "s_gM"))
    res.samples = blocks * B
    expected = ref_stats.basic_jackknife_analysis(T, res, {"E": 0.0, "Cv": 0.0})
    np.savez_compressed(join(HERE, "c5_stats.npz"), P=P, T=T, blocks=blocks, block_size=B,
                        keys=np.array(sorted(expected)), values=np.array([expected[k] for k in sorted(expected)]))
    print("c5_stats.npz:", {k: float(v) for k, v in expected.items()})


def main():
    pack_explicit_kat()
    pack_stats_kat()
    ex = join(REFERENCE, "examples")
    # c1: 2 surfaces x 2 modes, linear diagonal coupling only, P=12 (BASELINE.json configs[0])
    write_case("c1_2x2", join(ex, "artificial_systems/input_json/model_2x2.json"), None, P=12, T=300.0, X=24, B=8, seed=11)
    # c2: synthetic A=4, N=6, P=64 with linear + quadratic coupling (configs[1])
    c2 = synthetic.model_c2()
    write_case("c2_4x6", c2, None, P=64, T=300.0, X=16, B=8, seed=12)
    # odd number of beads, 3 surfaces, another temperature
    m34 = synthetic.coupled_model(3, 4, (0.05, 0.2), (1.0, 1.4), seed=7, quadratic=0.08)
    write_case("quad_3x4", m34, None, P=7, T=250.0, X=12, B=4, seed=13)
    # c3: paper model with a 4-surface sampling distribution on a 2-surface system (quirk Q1)
    write_case("jt_rho4", join(ex, "paper_1.5025058/input_json/jahnteller_3.json"),
               join(ex, "paper_1.5025058/alternate_rhos/jahnteller_D3_R1.json"), P=16, T=275.0, X=16, B=8, seed=14)
    # c3: displaced model with its alternate rho, P=128
    write_case("displaced_p128", join(ex, "paper_1.5025058/input_json/displaced_4.json"),
               join(ex, "paper_1.5025058/alternate_rhos/displaced_D4_R1.json"), P=128, T=350.0, X=8, B=4, seed=15)
    # c4 in miniature: A=12, N=24 (large-A path), few beads
    write_case("c4mini_12x24", synthetic.model_c4(), None, P=8, T=300.0, X=4, B=2, seed=16)
    # the shape of the reference's largest example model (artificial_systems/input_json/model_7x12.json, which itself has no
    # inter-surface coupling): 7 surfaces x 12 modes with linear + quadratic coupling -> blocked kernels, odd A, padded MMA tiles
    m712 = synthetic.coupled_model(7, 12, (0.02, 0.3), (0.5, 1.2), seed=23, quadratic=0.06)
    write_case("syn_7x12", m712, None, P=20, T=300.0, X=8, B=4, seed=18)
    # c5: c2 model sampled from a different rho (the un-rotated diagonal model)
    rng_free = synthetic.coupled_model(4, 6, (0.14, 0.45), (10.3, 10.9), mixing=0.0, quadratic=0.0)
    write_case("c5_altrho", c2, synthetic.diagonal_of(rng_free), P=16, T=300.0, X=16, B=8, seed=17)
    pack_c5_stats()


if __name__ == "__main__":
    main()
