"""Vibronic-model JSON input files: the data format on the input side of the PIMC hot path.

Keeps the on-disk contract of the reference (pibronic/vibronic/vibronic_model_io.py:42-73, 293-331,
473-650, 749-787; keys from vibronic_model_keys.py:9-26) so ``coupled_model.json`` /
``sampling_model.json`` files are interchangeable:

* arrays whose values are all zero are omitted on save and recreated as zeros on load,
* energies are always present after a load,
* the model hash is the SHA-512 hex digest of the file's text.

Only what ``BoxData.preprocess()`` and the benchmark/test model builders need is provided; model
*generation* from electronic-structure output is out of scope.
"""
import copy
import hashlib
import json
import shutil
from enum import Enum
from os.path import isfile

import numpy as np
from numpy import float64 as F64


class VibronicModelKeys(Enum):
    """keys (strings) used in the .json files"""
    number_of_modes = "number of modes"
    number_of_surfaces = "number of surfaces"
    energies = "energies"
    frequencies = "frequencies"
    linear_couplings = "linear couplings"
    quadratic_couplings = "quadratic couplings"
    cubic_couplings = "cubic couplings"
    quartic_couplings = "quartic couplings"
    # aliases
    N = number_of_modes
    A = number_of_surfaces
    E = energies
    w = frequencies
    G1 = linear_couplings
    G2 = quadratic_couplings
    G3 = cubic_couplings
    G4 = quartic_couplings


VMK = VibronicModelKeys
_ARRAY_KEYS = (VMK.E, VMK.w, VMK.G1, VMK.G2, VMK.G3, VMK.G4)


def model_shape_dict(A, N):
    """shapes of the arrays of a coupled model with A surfaces and N modes"""
    return {VMK.E: (A, A), VMK.w: (N, ), VMK.G1: (N, A, A), VMK.G2: (N, N, A, A),
            VMK.G3: (N, N, N, A, A), VMK.G4: (N, N, N, N, A, A)}


def diagonal_model_shape_dict(A, N):
    """shapes of the arrays of a model that is diagonal in the surfaces (a sampling model)"""
    return {VMK.E: (A, ), VMK.w: (N, ), VMK.G1: (N, A), VMK.G2: (N, N, A),
            VMK.G3: (N, N, N, A), VMK.G4: (N, N, N, N, A)}


def _dimensions(dictionary):
    return int(dictionary[VMK.A]), int(dictionary[VMK.N])


def _verify(dictionary, shape_fn):
    assert VMK.N in dictionary, "need the number of modes"
    assert VMK.A in dictionary, "need the number of surfaces"
    shapes = shape_fn(*_dimensions(dictionary))
    for key, value in dictionary.items():
        if key in shapes:
            assert np.shape(value) == shapes[key], f"{key} have incorrect shape"


def verify_model_parameters(kwargs):
    """the provided (coupled) model parameters follow the file conventions"""
    _verify(kwargs, model_shape_dict)


def verify_diagonal_model_parameters(kwargs):
    """the provided (diagonal) model parameters follow the file conventions"""
    _verify(kwargs, diagonal_model_shape_dict)


def _keys_to_enum(raw):
    return {VMK(key): value for key, value in raw.items()}


def _read(path):
    assert isfile(path), f"invalid path:\n{path}"
    with open(path, mode='r', encoding='UTF8') as file:
        return _keys_to_enum(json.loads(file.read()))


def _load_new(path, energy_shape_fn):
    model = _read(path)
    for key, value in model.items():
        if isinstance(value, list):
            model[key] = np.array(value, dtype=F64)
    if VMK.E not in model:
        model[VMK.E] = np.zeros(energy_shape_fn(*_dimensions(model))[VMK.E], dtype=F64)
    return model


def _load_inplace(path, dictionary):
    stored = _read(path)
    for key, value in dictionary.items():
        if isinstance(value, (np.ndarray, np.generic)):
            if key not in stored:
                dictionary[key].fill(0.0)
            else:
                dictionary[key][:] = np.array(stored[key], dtype=F64)


def load_model_from_JSON(path, dictionary=None):
    """returns a new dictionary, or fills the arrays of the provided one in place"""
    if not bool(dictionary):
        model = _load_new(path, model_shape_dict)
        verify_model_parameters(model)
        return model
    verify_model_parameters(dictionary)
    _load_inplace(path, dictionary)
    verify_model_parameters(dictionary)


def load_diagonal_model_from_JSON(path, dictionary=None):
    """returns a new dictionary, or fills the arrays of the provided one in place"""
    if not bool(dictionary):
        model = _load_new(path, diagonal_model_shape_dict)
        verify_diagonal_model_parameters(model)
        return model
    verify_diagonal_model_parameters(dictionary)
    _load_inplace(path, dictionary)
    verify_diagonal_model_parameters(dictionary)


def _save(path, dictionary):
    out = {}
    for key, value in copy.deepcopy(dictionary).items():
        name = key.value if isinstance(key, VMK) else key
        if isinstance(value, (np.ndarray, np.generic)):
            if np.count_nonzero(value) > 0:
                out[name] = value.tolist()
        else:
            out[name] = value
    with open(path, mode='w', encoding='UTF8') as target_file:
        target_file.write(json.dumps(out))


def save_model_to_JSON(path, dictionary):
    verify_model_parameters(dictionary)
    _save(path, dictionary)


def save_diagonal_model_to_JSON(path, dictionary):
    verify_diagonal_model_parameters(dictionary)
    _save(path, dictionary)


def extract_dimensions_of_model(FS=None, path=None):
    """returns A, N (in that order) of a coupled_model.json"""
    if FS is not None:
        path = FS.path_vib_model
    assert path is not None, "no arguments provided"
    return _dimensions(_read(path))


def extract_dimensions_of_diagonal_model(FS=None, path=None):
    """returns A, N (in that order) of a sampling_model.json"""
    if FS is not None:
        path = FS.path_rho_model
    assert path is not None, "no arguments provided"
    return _dimensions(_read(path))


def _hash(string):
    m = hashlib.sha512()
    m.update(string.encode('UTF-8'))
    return m.hexdigest()


def _hash_file(path):
    assert isfile(path), f"The path provided is not a valid file! Path:\n{path}"
    with open(path, mode='r', encoding='UTF8') as file:
        return _hash(file.read())


def create_model_hash(FS=None, path=None):
    """SHA-512 of the coupled_model.json text: guards result files against stale models"""
    if FS is not None:
        path = FS.path_vib_model
    assert path is not None, "no arguments provided"
    return _hash_file(path)


def create_diagonal_model_hash(FS=None, path=None):
    """SHA-512 of the sampling_model.json text"""
    if FS is not None:
        path = FS.path_rho_model
    assert path is not None, "no arguments provided"
    return _hash_file(path)


def remove_coupling_from_model(path_source, path_destination):
    """keeps only the surface-diagonal part of every array of a coupled model"""
    model = load_model_from_JSON(path_source)
    for key, value in model.items():
        if hasattr(value, 'shape') and len(value.shape) >= 2:
            model[key] = np.diagonal(value, axis1=value.ndim-2, axis2=value.ndim-1).copy()
    save_diagonal_model_to_JSON(path_destination, model)


def create_harmonic_model(FS):
    remove_coupling_from_model(FS.path_vib_model, FS.path_har_model)
    return FS.path_har_model


def create_basic_diagonal_model(FS):
    """the simplest sampling model: the diagonal of the coupled model"""
    source = create_harmonic_model(FS)
    shutil.copyfile(source, FS.path_rho_model)
    return FS.path_rho_model
