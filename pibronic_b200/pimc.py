"""Drop-in for ``pibronic.pimc``: the same block interface, evaluated by CUDA kernels on a B200.

Public names, attribute names, file formats and error behaviour follow the reference
(/root/reference/pibronic/pimc/pimc.py:36-47): ``BoxData[PM]``, ``BoxResult[PM]``,
``block_compute[_pm]``, the model classes, and the step helpers its golden test imports
(``build_o_matrix``, ``build_denominator``, ``diagonalize_coupling_matrix``, ``build_numerator``).

What differs, by design:

* ``block_compute[_pm]`` issue ONE fused sampler+estimator launch for all ``blocks*block_size``
  samples through the C ABI (``include/pbx.h``) instead of looping over blocks in numpy.  Per-block
  sums are reduced on the device and kept in ``result.block_sums``.
* random numbers are Philox4x32-10 streams keyed by ``data.seed`` and indexed by the global sample
  number ``data.sample_offset + i`` -- not numpy's global MT19937 (pimc.py:55, 326-334).  A run is
  reproducible from (seed, offset) and independent of how samples are split over GPUs.
* the mixture component of each sample is drawn on the device; the (X,N,P) ``standard_deviation``
  table and the print of all sample sources (pimc.py:386-404) do not exist.
* there is no CPU path: without the compiled extension or without a CUDA device the compute entry
  points raise.  ``preprocess()`` itself is host-only and works anywhere.
"""
import json
import os
from functools import partial

import numpy as np
from numpy import newaxis as NEW
from numpy import float64 as F64

from . import _cabi
from . import constants
from . import file_name
from . import file_structure
from . import model_io as vIO
from . import npz_writer
from .model_io import VMK
from .server import ServerExecutionParameters as SEP

__all__ = [
           "ModelClass",
           "ModelVibronic",
           "ModelVibronicPM",
           "ModelSampling",
           "BoxData",
           "BoxDataPM",
           "BoxResult",
           "BoxResultPM",
           "block_compute",
           "block_compute_pm",
           ]

hbar = constants.hbar


def _fresh_seed():
    """64 bits of OS entropy (the reference reseeds numpy from the OS at import, pimc.py:55)"""
    return int.from_bytes(os.urandom(8), "little")


class TemperatureDependentClass:
    """temperature dependent constants of one model at one tau (pimc.py:59-89)"""

    def __init__(self, model, tau, table_set=None):
        omega = np.broadcast_to(model.omega, model.size['AN'])
        self.cothAN = np.tanh(hbar*tau*omega)**(-1.)
        self.cschAN = np.sinh(hbar*tau*omega)**(-1.)
        self.cothANP = self.cothAN.copy().reshape(*self.cothAN.shape, 1)
        self.cschANP = self.cschAN.copy().reshape(*self.cschAN.shape, 1)
        self.cothBANP = self.cothANP[NEW, ...]
        self.cschBANP = self.cschANP[NEW, ...]

        energy = np.diag(model.energy) if len(model.energy.shape) > 1 else model.energy
        tilde_energy = energy + model.delta_weight
        per_surface = np.exp(-tau * tilde_energy) * np.prod(self.cschAN, axis=1)**0.5
        # same values for every (block, bead): kept as a read-only broadcast, not B*P copies
        self.omatrix_prefactor = np.broadcast_to(per_surface, model.size['BPA'])

        self.tau = tau
        self.table_set = table_set  # which device table this object mirrors: 'rho', 0 (tau), 1 (tau+), 2 (tau-)
        self._size = model.size
        self._omatrix = None
        self._omatrix_scaling = None

    # the (B,P,A,A) cache is only needed by the step helpers: allocate on first use
    @property
    def omatrix(self):
        if self._omatrix is None:
            self._omatrix = np.zeros(self._size['BPAA'])
        return self._omatrix

    @omatrix.setter
    def omatrix(self, value):
        self._omatrix = value

    @property
    def omatrix_scaling(self):
        if self._omatrix_scaling is None:
            self._omatrix_scaling = np.empty(self._size['BP'])
        return self._omatrix_scaling


class ModelClass:
    """information describing a quantum mechanical system (pimc.py:92-149)"""
    states = 0
    modes = 0
    omega = None
    energy = None
    linear = None
    quadratic = None
    cubic = None
    quartic = None

    def __init__(self, states=1, modes=1):
        self.states = states
        self.modes = modes
        self.state_range = range(states)
        self.mode_range = range(modes)

    def load_model(self, path):
        """fills energy, omega, linear, quadratic from a coupled_model.json (absent arrays are zeros)"""
        A, N = self.states, self.modes
        kwargs = {VMK.N: N, VMK.A: A}
        shape = vIO.model_shape_dict(A, N)
        for key in (VMK.E, VMK.w, VMK.G1, VMK.G2):
            kwargs[key] = np.zeros(shape[key], dtype=F64)
        vIO.load_model_from_JSON(path, kwargs)
        self.energy, self.omega = kwargs[VMK.E], kwargs[VMK.w]
        self.linear, self.quadratic = kwargs[VMK.G1], kwargs[VMK.G2]


class ModelVibronic(ModelClass):
    """the system of interest (pimc.py:152-235)"""

    def __init__(self, data):
        super().__init__(data.states, data.modes)
        self.size = data.size
        self.beta = data.beta
        self.tau = data.tau
        self.omega = np.zeros(self.size['N'], dtype=F64)
        self.energy = np.zeros(self.size['AA'], dtype=F64)
        self.linear = np.zeros(self.size['NAA'], dtype=F64)
        self.quadratic = np.zeros(self.size['NNAA'], dtype=F64)
        self.delta_weight = np.zeros(self.size['A'], dtype=F64)
        self.state_shift = np.zeros(self.size['AN'], dtype=F64)

    def load_model(self, path):
        super().load_model(path)
        # the device plan is built from the model as it is on disk, before anything is folded
        self.raw = dict(energy=self.energy.copy(), omega=self.omega.copy(), linear=self.linear.copy(),
                        quadratic=self.quadratic.copy())

    def compute_linear_displacement(self, data):
        """energy shift equivalent to the diagonal linear displacement"""
        idx = np.arange(data.states)
        diag_linear = self.linear[:, idx, idx]
        self.delta_weight[:] = -0.5 * (diag_linear**2. / self.omega[:, NEW]).sum(axis=0)

    def initialize_TDP_object(self):
        self.const = TemperatureDependentClass(self, self.tau, table_set=0)

    def finish_folding_in_terms(self, data):
        """the diagonal linear terms become oscillator shifts; they and diag(E) leave the coupling matrix"""
        idx = np.arange(data.states)
        self.state_shift[:] = (-self.linear[:, idx, idx] / self.omega[:, NEW]).T
        self.linear[:, idx, idx] = 0.0
        self.energy[idx, idx] = 0.0

    def precompute(self, data):
        self.compute_linear_displacement(data)
        self.initialize_TDP_object()
        self.finish_folding_in_terms(data)


class ModelVibronicPM(ModelVibronic):
    """plus minus version of ModelVibronic (pimc.py:238-269)"""
    delta_beta = 0.0
    beta_plus = 0.0
    beta_minus = 0.0
    tau_plus = 0.0
    tau_minus = 0.0

    def __init__(self, data):
        super().__init__(data)
        self.delta_beta = data.delta_beta

    def initialize_TDP_object(self):
        super().initialize_TDP_object()
        self.const_plus = TemperatureDependentClass(self, self.tau_plus, table_set=1)
        self.const_minus = TemperatureDependentClass(self, self.tau_minus, table_set=2)

    def precompute(self, data):
        self.beta_plus = self.beta + self.delta_beta
        self.beta_minus = self.beta - self.delta_beta
        self.tau_plus = self.beta_plus / data.beads
        self.tau_minus = self.beta_minus / data.beads
        super().precompute(data)


class ModelSampling(ModelClass):
    """the sampling distribution rho: a mixture of displaced harmonic oscillators (pimc.py:272-424)"""

    def __init__(self, data):
        self.param_dict = data.param_dict.copy()
        self.size_list = data.size_list.copy()
        self.tau = data.tau
        self.beta = data.beta

    def load_model(self, filePath):
        newStates, sameModes = vIO.extract_dimensions_of_diagonal_model(path=filePath)
        self.param_dict['A'] = self.states = newStates
        self.modes = sameModes
        self.size = {key: tuple(self.param_dict[letter] for letter in key) for key in self.size_list}

        kwargs = {VMK.N: self.modes, VMK.A: self.states,
                  VMK.E: np.zeros(self.size['A'], dtype=F64),
                  VMK.w: np.zeros(self.size['N'], dtype=F64),
                  VMK.G1: np.zeros(self.size['NA'], dtype=F64),
                  VMK.G2: np.zeros(self.size['NNA'], dtype=F64)}
        vIO.load_diagonal_model_from_JSON(filePath, kwargs)
        self.energy, self.omega = kwargs[VMK.E], kwargs[VMK.w]
        self.linear, self.quadratic = kwargs[VMK.G1], kwargs[VMK.G2]
        self.raw = dict(energy=self.energy.copy(), omega=self.omega.copy(), linear=self.linear.copy())

        self.delta_weight = np.zeros(self.size['A'], dtype=F64)
        self.state_weight = np.zeros(self.size['A'], dtype=F64)
        self.state_shift = np.zeros(self.size['AN'], dtype=F64)
        self.cc_samples = None  # the device sampler has no collective co-ordinates

    def compute_linear_displacement(self):
        self.delta_weight = -0.5 * (self.linear**2. / self.omega[:, NEW]).sum(axis=0)

    def compute_weight_for_each_state(self):
        """mixture weights (equation 49 of the method paper); normalised"""
        self.state_weight = np.exp(-self.beta * (self.energy + self.delta_weight))
        self.state_weight /= np.prod(np.sinh((self.beta * self.omega) / 2.))
        self.state_weight /= self.state_weight.sum()

    def initialize_TDP_object(self):
        self.const = TemperatureDependentClass(self, self.tau, table_set='rho')

    def finish_folding_in_terms(self):
        self.state_shift[:] = (-self.linear / self.omega[:, NEW]).T
        self.linear[:] = 0.0
        self.energy[:] = np.nan  # must not be used after this point

    def compute_sampling_constants(self, data):
        """covariance of the ring polymer in its normal modes (eigenvalues of the ring adjacency matrix in
        closed form, ascending like eigh returns them); the draws themselves happen on the device"""
        ring_eigvals = np.sort(2. * np.cos(2. * np.pi * np.arange(data.beads) / data.beads))
        self.inverse_covariance = (2. * self.const.cothANP - self.const.cschANP * ring_eigvals)

    def precompute(self, data):
        self.compute_linear_displacement()
        self.compute_weight_for_each_state()
        self.initialize_TDP_object()
        self.finish_folding_in_terms()
        self.compute_sampling_constants(data)


class BoxData:
    """execution parameters + models + device plan of one PIMC job (pimc.py:427-692)"""
    block_size = 0
    samples = 0
    blocks = 0

    states = 0
    modes = 0

    temperature = 0.0
    beads = 0

    beta = 0.0
    tau = 0.0
    delta_beta = 0.0

    id_data = 0
    id_rho = 0

    vib = None
    path_vib_model = ""
    rho = None
    path_rho_model = ""

    _COMMA_REPLACEMENT = ";"
    _SEPARATORS = (_COMMA_REPLACEMENT, ':')

    hash_vib = None
    hash_rho = None

    # ---- device-path knobs (no counterpart in the reference)
    seed = None              # Philox key; None -> drawn from OS entropy in preprocess()
    sample_offset = 0        # global index of this job's first sample (Philox counter)
    device = None            # CUDA device index; None -> LOCAL_RANK or the current device
    quirk_rho_trunc = False  # reproduce pimc.py:1110-1111 when A_rho > A
    eig_jacobi = False       # M from a Jacobi eigensolve instead of the scaling-and-squaring exponential
    force_generic = False    # never use the register-resident kernels
    jit = None               # True: compile a register-resident kernel for a shape that has none (pibronic_b200/jit.py: nvcc,
                             # ~1 min once, cached); None: only if PBX_JIT=1.  Default for such shapes: the fused tensor-core kernel
    m_tau_pm = False         # g+- with exp(-tau+- V) (consistent estimator, stats.consistent_jackknife_analysis); the
                             # reference uses exp(-tau V) for all three (pimc.py:1183)

    _PM = False

    @classmethod
    def from_FileStructure(cls, FS):
        data = cls()
        data.id_data = FS.id_data
        data.id_rho = FS.id_rho
        data.path_vib_model = FS.path_vib_model
        data.path_rho_model = FS.path_rho_model
        A, N = vIO.extract_dimensions_of_model(FS)
        data.states = A
        data.modes = N
        return data

    @classmethod
    def build(cls, id_data, id_rho):
        data = cls()
        data.id_data = id_data
        data.id_rho = id_rho
        return data

    @classmethod
    def from_json_file(cls, path):
        with open(path, mode='r', encoding='UTF8') as target_file:
            file_data = target_file.read()
        return cls.from_json_string(file_data)

    @classmethod
    def from_json_string(cls, json_str):
        data = cls()
        data.load_json_string(json_str)
        return data

    def __init__(self):
        self._plans = {}
        self._qTensor = None
        self._scratch = {}
        self._ring = None

    @classmethod
    def json_serialize(cls, params):
        return json.dumps(params, separators=cls._SEPARATORS)

    def encode_self(self, params=None):
        """JSON string of the execution parameters, commas replaced so it survives a SLURM env variable"""
        if params is None:
            params = {
                SEP.X: self.samples, SEP.nBlk: self.blocks, SEP.A: self.states, SEP.P: self.beads,
                SEP.N: self.modes, SEP.T: self.temperature, SEP.BlkS: self.block_size,
                SEP.dB: self.delta_beta, SEP.D: self.id_data, SEP.R: self.id_rho,
                SEP.beta: self.beta, SEP.tau: self.tau,
            }
        params = {(k.value if isinstance(k, SEP) else k): v for k, v in params.items()}
        return self.json_serialize(params)

    def load_json_string(self, json_str):
        """sets member parameters from the JSON string made by encode_self (+ "path_root")"""
        params = json.loads(json_str.replace(self._COMMA_REPLACEMENT, ","))
        self.samples = params[SEP.X.value]
        self.blocks = params[SEP.nBlk.value]
        self.states = params[SEP.A.value]
        self.beads = params[SEP.P.value]
        self.modes = params[SEP.N.value]
        self.temperature = params[SEP.T.value]
        self.block_size = params[SEP.BlkS.value]
        self.id_data = params[SEP.D.value]
        self.id_rho = params[SEP.R.value]
        self.beta = params.get(SEP.beta.value, 1.0 / (constants.boltzman * self.temperature))
        self.tau = params.get(SEP.tau.value, self.beta / self.beads)
        if type(params.get(SEP.dB.value)) is float:
            self.delta_beta = params[SEP.dB.value]
        for key in ("seed", "sample_offset"):
            if key in params:
                setattr(self, key, int(params[key]))

        FS = file_structure.FileStructure(params["path_root"], self.id_data, self.id_rho)
        self.path_vib_model = FS.path_vib_model
        self.path_rho_model = FS.path_rho_model
        FS.generate_model_hashes()
        self.hash_vib = FS.hash_vib
        self.hash_rho = FS.hash_rho

    # ------------------------------------------------------------------ set-up (host only)
    def _model_classes(self):
        return ModelVibronic, ModelSampling

    def initialize_models(self):
        vib_cls, rho_cls = self._model_classes()
        self.vib = vib_cls(self)
        self.vib.load_model(self.path_vib_model)
        self.vib.precompute(self)

        self.rho = rho_cls(self)
        self.rho.load_model(self.path_rho_model)
        self.rho.precompute(self)
        assert self.rho.modes == self.modes, "the sampling model must have the same number of modes"

    def preprocess(self):
        self.param_dict = {'X': self.samples, 'A': self.states, 'N': self.modes, 'P': self.beads,
                           'B': self.block_size, }
        self.size_list = ['X', 'P', 'N', 'A', 'B', 'XP', 'BP', 'AN', 'NA', 'AA', 'BNP', 'BPA', 'BAA', 'XNP',
                          'NAA', 'NNA', 'ANP', 'BANP', 'BPAA', 'NNAA', 'BPAN', ]
        self.size = {key: tuple(self.param_dict[letter] for letter in key) for key in self.size_list}

        self.beta = constants.beta(self.temperature)
        self.tau = self.beta / self.beads
        assert self.beads >= 3, "circulant matrix requires 3 or more beads"  # hard check
        if self.seed is None:
            self.seed = _fresh_seed()
        self._plans, self._scratch, self._qTensor, self._ring = {}, {}, None, None
        # step-helper outputs (pimc.py:675-680); filled by diagonalize_coupling_matrix / build_numerator
        self.coupling_matrix = None
        self.M_matrix = None
        self.numerator = None
        self.initialize_models()

    # ring-polymer normal modes: closed form eigenvalues; the dense matrices only on request
    @property
    def circulant_matrix(self):
        P = self.beads
        i = np.arange(P)
        C = np.zeros((P, P), dtype=int)
        C[i, (i + 1) % P] = 1
        C[i, (i - 1) % P] = 1
        return C

    def _ring_system(self):
        if self._ring is None:
            self._ring = np.linalg.eigh(self.circulant_matrix, UPLO='L')
        return self._ring

    @property
    def circulant_eigvals(self):
        return self._ring_system()[0]

    @property
    def circulant_eigvects(self):
        return self._ring_system()[1]

    # ------------------------------------------------------------------ scratch tensors of the step API
    @property
    def qTensor(self):
        if self._qTensor is None:
            self._qTensor = np.zeros(self.size['BANP'], dtype=F64)
        return self._qTensor

    @qTensor.setter
    def qTensor(self, value):
        self._qTensor = value

    @property
    def qTempTensor(self):
        if 'qTemp' not in self._scratch:
            self._scratch['qTemp'] = np.zeros([self.block_size, self.rho.states, self.modes, self.beads], dtype=F64)
        return self._scratch['qTemp']

    # ------------------------------------------------------------------ device side
    def _device_index(self):
        if self.device is not None:
            return int(self.device)
        if "LOCAL_RANK" in os.environ:
            return int(os.environ["LOCAL_RANK"])
        import torch
        if not torch.cuda.is_available():
            raise _cabi.PbxError("no CUDA device: pibronic_b200 has no CPU path")
        return torch.cuda.current_device()

    def device_plan(self, pm=None, no_scaling=False, device=None, rho_double_shift=False):
        """the pbx plan (device tables) for this job; one per (flags word, device): changing a knob gives a new plan"""
        pm = self._PM if pm is None else pm
        device = self._device_index() if device is None else device
        flags = (_cabi.FLAG_PM if pm else 0)
        flags |= _cabi.QUIRK_RHO_TRUNC if self.quirk_rho_trunc else 0
        flags |= _cabi.FLAG_EIG_JACOBI if self.eig_jacobi else 0
        flags |= _cabi.FLAG_FORCE_GENERIC if self.force_generic else 0
        # the stage-by-stage API (no_scaling) always uses tau: the consistent-estimator flag only exists on the fused path
        flags |= _cabi.FLAG_M_TAU_PM if (self.m_tau_pm and pm and not no_scaling) else 0
        flags |= _cabi.FLAG_NO_SCALING if no_scaling else 0
        flags |= _cabi.QUIRK_RHO_DOUBLE_SHIFT if rho_double_shift else 0
        key = (flags, device)
        if key not in self._plans:
            vib, rho = self.vib.raw, self.rho.raw
            self._plans[key] = _cabi.Plan(vib['energy'], vib['omega'], vib['linear'], vib['quadratic'],
                                          rho['energy'], rho['omega'], rho['linear'], self.beads, self.beta,
                                          self.delta_beta if pm else 0.0, flags=flags, device=device, jit=self.jit)
        return self._plans[key]

    def release(self):
        """frees the device plans"""
        for plan in self._plans.values():
            plan.close()
        self._plans = {}

    # ------------------------------------------------------------------ per-block sampler API
    def draw_sample(self, sample_view):
        """draws the bead co-ordinates of one block on the device (pimc.py:595-600, 326-334)"""
        import torch
        plan = self.device_plan()
        n = sample_view.stop - sample_view.start
        with torch.cuda.device(plan.device):
            R = torch.empty((n, self.modes, self.beads), dtype=torch.float64, device="cuda")
            src = torch.empty(n, dtype=torch.int32, device="cuda")
            plan.sample_coords(self.seed, self.sample_offset + sample_view.start, n, R, src)
            self._scratch['drawn'] = R.cpu().numpy()
            self.rho.sample_sources = src.cpu().numpy()

    def transform_sampled_coordinates(self, sample_view):
        """publishes the drawn bead co-ordinates in qTensor (B,A,N,P) / qTempTensor (pimc.py:613-631)"""
        R = self._scratch['drawn']
        self.qTensor[:R.shape[0]] = R[:, NEW, ...]
        self.qTempTensor[:R.shape[0]] = R[:, NEW, ...]

    def generate_random_R_values(self, result, storage_array, sample_view):
        """uniform random co-ordinates in [0, 1) with no relation to rho or g, only the right dimensions (pimc.py:602-611);
        stored in storage_array[sample_view] and published in qTensor.  The stream is numpy's PCG64 seeded with
        (data.seed, first sample of the view): reproducible, independent of how the run is cut into blocks"""
        n = sample_view.stop - sample_view.start
        rng = np.random.default_rng([int(self.seed), int(self.sample_offset) + int(sample_view.start)])
        storage_array[sample_view, :, :] = rng.random(size=(n, self.modes, self.beads))
        self.qTensor[:n] = storage_array[sample_view, NEW, ...]


class BoxDataPM(BoxData):
    """plus minus version of BoxData (pimc.py:695-738)"""
    delta_beta = 0.0
    beta_plus = 0.0
    beta_minus = 0.0
    tau_plus = 0.0
    tau_minus = 0.0
    _PM = True

    def __init__(self, delta_beta=None):
        if delta_beta is None:
            delta_beta = constants.delta_beta
        self.delta_beta = delta_beta
        super().__init__()

    def _model_classes(self):
        return ModelVibronicPM, ModelSampling

    def preprocess(self):
        super().preprocess()
        # the values actually used are the model's (pimc.py:261-266); mirror them here
        self.beta_plus = self.vib.beta_plus
        self.beta_minus = self.vib.beta_minus
        self.tau_plus = self.vib.tau_plus
        self.tau_minus = self.vib.tau_minus


def _pinned_rows(rows, samples):
    """(rows, samples) float64 filled with NaN; page-locked when a CUDA device is present so the
    device-to-host copy of the results lands in it directly"""
    if samples > 0:
        try:
            import torch
            if torch.cuda.is_available():
                store = torch.empty((rows, samples), dtype=torch.float64, pin_memory=True).numpy()
                store.fill(np.nan)
                return store
        except Exception:
            pass
    return np.full((rows, samples), np.nan, dtype=F64)


class BoxResult:
    """results of one PIMC job and their .npz files (pimc.py:741-912)"""

    id_job = None
    hash_vib = None
    hash_rho = None
    block_sums = None  # (blocks, 8) per-block sums reduced on the device (see _cabi.SUM_NAMES)

    key_list = ["number_of_samples", "s_rho", "s_g"]
    _ROWS = 2

    @classmethod
    def read_number_of_samples(cls, path_full):
        with np.load(path_full, mmap_mode='r') as data:
            return data["number_of_samples"]

    @classmethod
    def verify_result_keys_are_present(cls, path, fileObj):
        for k in cls.key_list:
            if k not in fileObj.keys():
                s = "Expected key ({:s}) not present in result file\n{:s}\n"
                raise AssertionError(s.format(k, path))

    @classmethod
    def result_keys_are_present_in(cls, iterable):
        return all(k in iterable for k in cls.key_list)

    def initialize_arrays(self):
        self._store = _pinned_rows(self._ROWS, int(self.samples))
        self.scaled_rho = self._store[0]
        self.scaled_g = self._store[1]

    def __init__(self, data=None, X=None):
        if data is not None:
            self.partial_name = partial(file_name.pimc().format, P=data.beads, T=data.temperature)
            self.samples = data.samples
            self.hash_vib = data.hash_vib
            self.hash_rho = data.hash_rho
        elif X is not None:
            self.samples = X
        else:
            self.samples = 0
        self.initialize_arrays()

    def compute_path_to_file(self):
        if self.samples == 0:
            raise AssertionError("{:s} still has 0 samples - this should not happen".format(self.__class__.__name__))
        J = int(self.id_job) if self.id_job is not None else 0
        self.name = self.partial_name(J=J)
        return os.path.join(self.path_root, self.name)

    def _arrays_to_save(self):
        return dict(s_rho=self.scaled_rho, s_g=self.scaled_g)

    def save_results(self, number_of_samples):
        path = self.compute_path_to_file()
        assert self.hash_vib is not None and self.hash_rho is not None, "we save hash values if they don't exist!"
        # same members and file format as the reference's np.savez (pimc.py:825-834, 954-963), written faster
        npz_writer.savez(path, hash_vib=self.hash_vib, hash_rho=self.hash_rho, number_of_samples=self.samples,
                         **self._arrays_to_save())

    def _take(self, data, start, length):
        self.scaled_g[start:start+length] = data["s_g"][0:length]
        self.scaled_rho[start:start+length] = data["s_rho"][0:length]

    def load_results(self, path):
        """load results from one file"""
        with np.load(path) as data:
            self.__class__.verify_result_keys_are_present(path, data)
            n = int(data["number_of_samples"])
            if self.samples == 0:
                self.samples = n
            elif n != self.samples:
                raise AssertionError("BoxResult has a different number of samples that the input file - this should not happen")
            self.initialize_arrays()
            self._take(data, 0, n)

    def load_multiple_results(self, list_of_paths, desired_number_of_samples=None):
        """concatenates the shards of a run (one file per job); files without the expected keys are skipped;
        loads UP TO desired_number_of_samples samples if that is given"""
        assert not len(list_of_paths) == 0, "list_of_paths cannot be empty"
        number_of_samples = 0
        good_paths = []
        try:
            for path in list_of_paths:
                with np.load(path, mmap_mode="r") as data:
                    if self.__class__.result_keys_are_present_in(data.keys()):
                        number_of_samples += int(data["number_of_samples"])
                        good_paths.append(path)
        except Exception as err:
            print("Did we get another mangled .npz file?\nIf numpy can't load the file because it is not a zip file then just delete the offending file and rerun the script")
            raise err
        assert len(good_paths) > 0, "none of the provided paths were good, i.e. none of them had all the required keys {}".format(self.key_list)
        assert number_of_samples != 0, "number of samples should have changed"

        if desired_number_of_samples is None:
            desired_number_of_samples = number_of_samples
        elif number_of_samples >= desired_number_of_samples:
            number_of_samples = desired_number_of_samples
        else:
            print("We found less samples than were requested!!")

        self.samples = number_of_samples
        self.initialize_arrays()

        start = 0
        for path in good_paths:
            if start >= desired_number_of_samples:
                break
            with np.load(path) as data:
                length = min(int(data["number_of_samples"]), desired_number_of_samples - start)
                assert not np.any(data["s_rho"][0:length] == 0.0), "Zeros in the denominator"
                self._take(data, start, length)
                start += length


class BoxResultPM(BoxResult):
    """plus minus version of BoxResult (pimc.py:915-1040)"""

    key_list = ["s_gP", "s_gM"] + BoxResult.key_list
    _ROWS = 4

    def initialize_arrays(self):
        super().initialize_arrays()
        self.scaled_gofr_plus = self._store[2]
        self.scaled_gofr_minus = self._store[3]

    def _arrays_to_save(self):
        return dict(s_rho=self.scaled_rho, s_g=self.scaled_g, s_gP=self.scaled_gofr_plus,
                    s_gM=self.scaled_gofr_minus)

    def _take(self, data, start, length):
        super()._take(data, start, length)
        self.scaled_gofr_plus[start:start+length] = data["s_gP"][0:length]
        self.scaled_gofr_minus[start:start+length] = data["s_gM"][0:length]


# ---------------------------------------------------------------------------------------------
# step helpers (pimc.py:1062-1213).  They keep the reference's calling convention -- numpy arrays
# attached to the data/model objects are read and written in place -- but the arithmetic of each
# step (exp of the harmonic exponents, V, exp(-tau V), the bead chain) runs in the CUDA stage
# kernels.  Only block_compute[_pm] below is the production path; these exist so that the
# reference's own golden test can be run step by step against this package.
# ---------------------------------------------------------------------------------------------
def scale_o_matrices(scalingFactor, model_one, model_two):
    """divides the O matrices of the models by the scalingFactor"""
    model_one.omatrix /= scalingFactor[..., NEW, NEW]
    model_two.omatrix /= scalingFactor[..., NEW, NEW]


def un_scale_o_matrices(scalingFactor, model_one, model_two):
    """multiples the O matrices of the models by the scalingFactor"""
    model_one.omatrix *= scalingFactor[..., NEW, NEW]
    model_two.omatrix *= scalingFactor[..., NEW, NEW]


def build_scaling_factors(S12, model_one, model_two):
    """individual and combined scaling factors of both models"""
    model_one.omatrix_scaling[:] = np.amax(model_one.omatrix, axis=(2, 3))
    model_two.omatrix_scaling[:] = np.amax(model_two.omatrix, axis=(2, 3))
    S12[:] = np.maximum(model_one.omatrix_scaling, model_two.omatrix_scaling)


def _step_coordinates(data, surfaces):
    """bead co-ordinates (B,N,P) the reference's step functions would read for a model with `surfaces` surfaces"""
    tensor = data.qTensor if surfaces == np.shape(data.qTensor)[1] else data.qTempTensor
    return np.ascontiguousarray(tensor[:, 0], dtype=F64)


def _stage(data, R, want):
    """runs the stage kernel on co-ordinates R (B,N,P); returns the requested unscaled intermediates"""
    import torch
    plan = data.device_plan(pm=True, no_scaling=True)
    B, P, A, Ar = R.shape[0], data.beads, plan.A, plan.Ar
    with torch.cuda.device(plan.device):
        dev = dict(R=torch.from_numpy(R).cuda())
        shapes = dict(o_rho=(B, P, Ar), o_vib=(3, B, P, A), v_mat=(B, P, A, A), m_mat=(B, P, A, A))
        out = {k: torch.empty(shapes[k], dtype=torch.float64, device="cuda") for k in want}
        plan.eval_stages(dev['R'], **out)
        return {k: v.cpu().numpy() for k, v in out.items()}


def build_o_matrix(data, model, state_shift):
    """O matrix of a model for the co-ordinates in data.qTensor, stored (unscaled) in model.omatrix (B,P,A,A)"""
    surfaces = state_shift.shape[0]
    R = _step_coordinates(data, surfaces)
    if model.table_set == 'rho':
        diag = _stage(data, R, ['o_rho'])['o_rho']
    else:
        diag = _stage(data, R, ['o_vib'])['o_vib'][model.table_set]
    omatrix = model.omatrix
    omatrix[...] = 0.0
    # the reference fills range(data.states) surfaces whatever the model (pimc.py:1110-1111)
    for a in range(min(data.states, omatrix.shape[2])):
        omatrix[:diag.shape[0], :, a, a] = diag[:, :, a]


def build_denominator(rho_model, outputArray, idx):
    """state trace over the bead product of the O matrices of the rho model"""
    outputArray[idx] = rho_model.omatrix.prod(axis=1).trace(axis1=1, axis2=2)


def diagonalize_coupling_matrix(data):
    """coupling matrix V(R) per bead -> data.coupling_matrix, and M = exp(-tau V) -> data.M_matrix.
    (The reference keeps eigenvalues/eigenvectors and forms M in build_numerator; M is the same matrix.)"""
    R = _step_coordinates(data, data.states)
    out = _stage(data, R, ['v_mat', 'm_mat'])
    data.coupling_matrix = out['v_mat']
    data.M_matrix = out['m_mat']


def build_numerator(data, vib, outputArray, idx):
    """g = trace over the surfaces of the bead chain prod_p M[b,p] . O[b,p] -> outputArray[idx]"""
    import torch
    plan = data.device_plan(pm=True, no_scaling=True)
    A = plan.A
    o_diag = np.ascontiguousarray(vib.omatrix[..., np.arange(A), np.arange(A)])
    with torch.cuda.device(plan.device):
        m_dev = torch.from_numpy(np.ascontiguousarray(data.M_matrix)).cuda()
        o_dev = torch.from_numpy(o_diag).cuda()
        g = torch.empty(m_dev.shape[0], dtype=torch.float64, device="cuda")
        plan.chain_trace(m_dev, o_dev, g)
        outputArray[idx] = g.cpu().numpy()


# ---------------------------------------------------------------------------------------------
# the production path
# ---------------------------------------------------------------------------------------------
def _compute_on_device(data, result, pm):
    """one fused sampler+estimator launch for samples [0, blocks*block_size) + per-block sums, through the
    host-buffer C ABI call: the results are copied device->host straight into the result arrays"""
    n = int(data.blocks) * int(data.block_size)
    assert 0 < n <= result.samples, "blocks*block_size must be in (0, samples]"
    plan = data.device_plan(pm=pm)
    rows = 4 if pm else 2
    names = ("scaled_rho", "scaled_g", "scaled_gofr_plus", "scaled_gofr_minus")[:rows]
    store = getattr(result, "_store", None)
    direct = store is not None and store.shape[0] == rows and all(
        getattr(result, name).ctypes.data == store[k].ctypes.data for k, name in enumerate(names))
    target = store if direct else np.empty((rows, n))
    _, result.block_sums = plan.sample_eval_host(data.seed, data.sample_offset, n, out4=target,
                                                 block_size=int(data.block_size))
    if not direct:
        for k, name in enumerate(names):
            getattr(result, name)[:n] = target[k]
    return n


def block_compute_gR(data, result):
    """only g(R), NOT scaled by S, for blocks*block_size sampled points (pimc.py:1216-1247: data sets for
    training ML models).  Runs on the unscaled generic kernels; saves the .npz like block_compute."""
    import torch
    n = int(data.blocks) * int(data.block_size)
    assert 0 < n <= result.samples, "blocks*block_size must be in (0, samples]"
    sampler = data.device_plan(pm=False)
    plan = data.device_plan(pm=False, no_scaling=True)
    with torch.cuda.device(plan.device):
        R = torch.empty((n, data.modes, data.beads), dtype=torch.float64, device="cuda")
        out = torch.empty((4, n), dtype=torch.float64, device="cuda")
        sampler.sample_coords(data.seed, data.sample_offset, n, R)
        plan.eval_coords(R, out)
        result.scaled_g[:n] = out[1].cpu().numpy()
    result.save_results(n)


def _eval_unscaled(data, R, rho_double_shift=False):
    """rho(R) and g(R) WITHOUT the S scaling for caller supplied co-ordinates R (n, N, P): (2, n) host array"""
    import torch
    plan = data.device_plan(pm=False, no_scaling=True, rho_double_shift=rho_double_shift)
    R = np.ascontiguousarray(R, dtype=F64)
    out = np.empty((2, R.shape[0]))
    with torch.cuda.device(plan.device):
        plan.eval_coords_host(R, out4=out)
    return out


def save_gR_with_samples(data, result, input_R_values):
    """g(R) with the R values it belongs to, as two .npz files next to the PIMC results (pimc.py:1250-1270):
    ..._training_data_g_output.npz {number_of_samples, g} and ..._training_data_input.npz {number_of_samples, input_R_values}"""
    from functools import partial
    result.partial_name = partial(file_name.training_data_g_output().format, P=data.beads, T=data.temperature)
    np.savez(result.compute_path_to_file(), number_of_samples=result.samples, g=result.scaled_g)
    result.partial_name = partial(file_name.training_data_input().format, P=data.beads, T=data.temperature)
    np.savez(result.compute_path_to_file(), number_of_samples=result.samples, input_R_values=input_R_values)


def block_compute_rhoR_from_input_samples(data, result, input_R_values, quirk_double_shift=False):
    """rho(R), not scaled, for caller supplied co-ordinates input_R_values (X, N, P) -> result.scaled_rho[0:blocks*block_size];
    the caller saves the results (pimc.py:1273-1302).  One launch on the device instead of the block loop.

    quirk_double_shift=True reproduces the reference, which subtracts the SYSTEM's shift from the co-ordinates and then
    the sampling model's (pimc.py:1293-1298): rho evaluated at R - d_vib[a] - d_rho[a]; it needs
    A_rho == A (the reference reads stale scratch otherwise).  Default: rho at the co-ordinates given."""
    n = int(data.blocks) * int(data.block_size)
    assert 0 < n <= result.samples and input_R_values.shape[0] >= n, "blocks*block_size must be in (0, samples]"
    result.scaled_rho[:n] = _eval_unscaled(data, input_R_values[:n], rho_double_shift=quirk_double_shift)[0]


def block_compute_gR_from_raw_samples(data, result):
    """g(R), not scaled, on uniform random co-ordinates that were NOT sampled from rho, saved with those co-ordinates
    (pimc.py:1305-1338: training sets for ML models)"""
    n = int(data.blocks) * int(data.block_size)
    assert 0 < n <= result.samples, "blocks*block_size must be in (0, samples]"
    input_R_values = np.empty(data.size['XNP'], dtype=F64)
    for block_index in range(int(data.blocks)):
        view = slice(block_index * data.block_size, (block_index + 1) * data.block_size)
        data.generate_random_R_values(result, input_R_values, view)
    result.scaled_g[:n] = _eval_unscaled(data, input_R_values[:n])[1]
    save_gR_with_samples(data, result, input_R_values)


def block_compute(data, result):
    """numerator g and denominator rho for blocks*block_size sampled points; saves the .npz"""
    end = _compute_on_device(data, result, pm=False)
    result.save_results(end)


def block_compute_pm(data, result):
    """rho, g, g(beta+delta_beta), g(beta-delta_beta) for blocks*block_size sampled points; saves the .npz"""
    assert isinstance(data, BoxDataPM), "incorrect object type"
    assert isinstance(result, BoxResultPM), "incorrect object type"
    end = _compute_on_device(data, result, pm=True)
    result.delta_beta = data.delta_beta      # the finite-difference step these g+- belong to (read by pibronic_b200.stats)
    result.save_results(end)


def simple_wrapper(id_data, id_rho=0, path_root='/work/ngraymon/pimc/', states=2, modes=2, beads=1000):
    """Just do simple expval(Z) calculation (pimc.py:1465-1497): 100 samples in one block at 300 K; the reference's
    hard-coded root directory, surfaces, modes and beads are the defaults of the keyword arguments"""
    samples = int(1e2)
    Bsize = int(1e2)
    data = BoxData()
    data.seed = 232942   # the reference seeds numpy with this number
    data.id_data = id_data
    data.id_rho = id_rho
    files = file_structure.FileStructure(path_root, id_data, id_rho)
    data.path_vib_model = files.path_vib_model
    data.path_rho_model = files.path_rho_model
    data.states = states
    data.modes = modes
    data.samples = samples
    data.beads = beads
    data.temperature = 300.0
    data.blocks = samples // Bsize
    data.block_size = Bsize
    files.generate_model_hashes()
    data.hash_vib, data.hash_rho = files.hash_vib, files.hash_rho
    data.preprocess()
    results = BoxResult(data=data)
    results.path_root = files.path_rho_results
    block_compute(data, results)
    data.release()
    return results


def plus_minus_wrapper(id_data, id_rho=0, path_root='/work/ngraymon/pimc/', states=3, modes=6, beads=20):
    """Calculate all the possible temp +/- approaches (pimc.py:1500-1535): 100 samples in one block at 300 K"""
    samples = int(1e2)
    Bsize = int(1e2)
    data = BoxDataPM(constants.delta_beta)
    data.seed = 232942
    data.id_data = id_data
    data.id_rho = id_rho
    files = file_structure.FileStructure(path_root, id_data, id_rho)
    data.path_vib_model = files.path_vib_model
    data.path_rho_model = files.path_rho_model
    data.states = states
    data.modes = modes
    data.samples = samples
    data.beads = beads
    data.temperature = 300.00
    data.blocks = samples // Bsize
    data.block_size = Bsize
    files.generate_model_hashes()
    data.hash_vib, data.hash_rho = files.hash_vib, files.hash_rho
    data.preprocess()
    results = BoxResultPM(data=data)
    results.path_root = files.path_rho_results
    block_compute_pm(data, results)
    data.release()
    return results
