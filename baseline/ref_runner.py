"""Runs the staged, unmodified reference (``baseline/_ref/pibronic``) through its own public API:
``BoxDataPM`` / ``BoxResultPM`` / ``block_compute_pm`` (pibronic/pimc/pimc.py:1388-1462).

Used only by ``bench.py`` (reference arm and ``cpu_baseline``) and by the CPU tests that check that files written
by this repo load in the reference's readers.  Nothing under ``pibronic_b200/`` imports this module.
"""
import contextlib
import io
import os
import shutil
import sys
import tempfile
import time
import warnings
from os.path import abspath, dirname, isdir, join
from unittest.mock import MagicMock

REF_ROOT = join(dirname(abspath(__file__)), "_ref")

# third-party modules the reference imports at package level but never touches on this path
_ABSENT = ['parse', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.lines', 'matplotlib.ticker', 'matplotlib.gridspec',
           'matplotlib.backends', 'matplotlib.backends.backend_pdf', 'mpl_toolkits', 'mpl_toolkits.mplot3d',
           'mpl_toolkits.axes_grid1', 'fortranformat', 'julia', 'memory_profiler']


def available():
    return isdir(join(REF_ROOT, "pibronic"))


def import_reference():
    """the reference's ``pibronic`` package (staged copy); raises ImportError if it was not staged"""
    if not available():
        raise ImportError(f"{REF_ROOT}/pibronic is missing: run `python baseline/stage_reference.py` in the build container")
    for name in _ABSENT:
        try:
            __import__(name)
        except Exception:
            sys.modules.setdefault(name, MagicMock())
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    import pibronic  # noqa: F401
    from pibronic import pimc
    from pibronic.vibronic import vIO
    return pimc, vIO


def run_block_compute_pm(path_vib, path_rho, P, T, X, B, seed):
    """the reference's block loop on X samples in blocks of B; returns (seconds in block_compute_pm, result object)"""
    import numpy as np
    pimc, vIO = import_reference()
    A, N = vIO.extract_dimensions_of_model(path=path_vib)
    data = pimc.BoxDataPM()
    data.id_data, data.id_rho = 0, 0
    data.path_vib_model, data.path_rho_model = path_vib, path_rho
    data.states, data.modes = A, N
    data.samples, data.beads, data.temperature = X, P, T
    data.block_size, data.blocks = B, X // B
    data.hash_vib = vIO.create_model_hash(path=path_vib)
    data.hash_rho = vIO.create_diagonal_model_hash(path=path_rho)
    np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        data.preprocess()
    result = pimc.BoxResultPM(data=data)
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    result.path_root, result.id_job = tmp, 0
    try:
        t0 = time.perf_counter()
        pimc.block_compute_pm(data, result)
        dt = time.perf_counter() - t0
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return dt, result
