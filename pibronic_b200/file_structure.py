"""Directory contract of a Pibronic data set (pibronic/data/file_structure.py:30-140, 231-253).

``root/data_set_{D}/{parameters,results,execution_output,plots}/`` and the same four
sub-directories under ``root/data_set_{D}/rho_{R}/``.  The hot path reads the two model files from
the ``parameters`` directories and writes its ``.npz`` shards into ``rho_{R}/results``.
"""
import os
from os.path import join

from . import file_name
from . import model_io as vIO

_SUB_DIRS = ("parameters", "results", "execution_output", "plots")
_SHORT = {"parameters": "params", "results": "results", "execution_output": "output", "plots": "plots"}


class FileStructure:
    """paths of one (data set, sampling distribution) pair; creates the directories unless told not to"""

    def __init__(self, path_root, id_data, id_rho=0, no_makedir=False):
        assert os.path.exists(path_root), f"the path_root ({path_root}) does not exist"
        assert os.path.isdir(path_root), f"the path_root ({path_root}) is not a directory"
        self.id_data = id_data
        self.id_rho = id_rho
        self.path_root = os.path.abspath(str(path_root))
        self.path_data = join(self.path_root, f"data_set_{id_data:d}/")
        self.path_es = join(self.path_data, "electronic_structure")
        for sub in _SUB_DIRS:
            setattr(self, f"path_vib_{_SHORT[sub]}", join(self.path_data, sub + "/"))
        self._point_at_rho(id_rho)
        self.template_sos_vib = self.path_vib_params + "sos_B{B:d}.json"
        # attribute names of the DIRECTORIES (the path_* attributes naming files are not in this list)
        self.dir_list = ['path_root', 'path_data', 'path_es', 'path_rho'] + \
            [f'path_{kind}_{_SHORT[sub]}' for kind in ('vib', 'rho') for sub in _SUB_DIRS]

        self.path_vib_model = join(self.path_vib_params, file_name.coupled_model)
        self.path_har_model = join(self.path_vib_params, file_name.harmonic_model)
        self.path_analytic_vib = join(self.path_vib_params, file_name.analytic_results)
        if not no_makedir:
            self.make_directories()

    def _point_at_rho(self, id_rho):
        self.id_rho = id_rho
        self.path_rho = join(self.path_data, f"rho_{id_rho:d}/")
        for sub in _SUB_DIRS:
            setattr(self, f"path_rho_{_SHORT[sub]}", join(self.path_rho, sub + "/"))
        self.template_pimc = self.path_rho_results + file_name.pimc(J="{J:s}")
        self.template_jackknife = self.path_rho_results + file_name.jackknife()
        self.template_sos_rho = self.path_rho_params + "sos_B{B:d}.json"
        self.path_rho_model = join(self.path_rho_params, file_name.sampling_model)
        self.path_analytic_rho = join(self.path_rho_params, file_name.analytic_results)

    @classmethod
    def from_boxdata(cls, path_root, data):
        return cls(path_root, data.id_data, data.id_rho)

    def directories_exist(self):
        return all(os.path.isdir(getattr(self, d)) for d in self.dir_list)

    def make_directories(self):
        for name in self.dir_list:
            os.makedirs(getattr(self, name), exist_ok=True)

    def change_rho(self, id_rho):
        """point every rho path at another sampling distribution (directories are created)"""
        self._point_at_rho(id_rho)
        for sub in _SUB_DIRS:
            os.makedirs(getattr(self, f"path_rho_{_SHORT[sub]}"), exist_ok=True)

    def generate_model_hashes(self, force_flag=False):
        """attributes hash_vib / hash_rho: SHA-512 of the two model files"""
        if hasattr(self, 'hash_vib') and hasattr(self, 'hash_rho') and not force_flag:
            return
        self.hash_vib = vIO.create_model_hash(FS=self)
        self.hash_rho = vIO.create_diagonal_model_hash(FS=self)

    def valid_vib_hash(self, model_dict):
        return bool(model_dict["hash_vib"] == self.hash_vib)

    def valid_rho_hash(self, model_dict):
        return bool(model_dict["hash_rho"] == self.hash_rho)
