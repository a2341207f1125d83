"""Finding and loading the result files of a run (pibronic/data/postprocessing.py:200-343): the glue
between the ``.npz`` shards the hot path writes and the statistics that consume them."""
import glob
import json
import os
import re

from . import file_name

_PIMC_NAME = re.compile(r"P(\d+)_T(\d+\.\d+)_J(\d+)_data_points\.npz$")


def retrive_pimc_file_list(FS):
    """all P*_T*_J*_data_points.npz files of the FileStructure's rho results directory"""
    return sorted(glob.glob(os.path.join(FS.path_rho_results, file_name.pimc("*", "*", "*"))))


def _field(list_of_paths, index, cast):
    return sorted({cast(_PIMC_NAME.search(p).group(index)) for p in list_of_paths if _PIMC_NAME.search(p)})


def extract_bead_paramater_list(list_of_paths):
    return _field(list_of_paths, 1, int)


def extract_temperature_paramater_list(list_of_paths):
    return _field(list_of_paths, 2, float)


def load_pimc_data(FS, P, T, pimc_results):
    """all shards (J*) with the same P and T, concatenated"""
    pattern = FS.template_pimc.format(P=P, T=T, J="*")
    pimc_results.load_multiple_results(sorted(glob.glob(pattern)))


def load_analytic_data(FS, T, analytic):
    """Z, E, Cv, alpha+- of the sampling distribution at temperature T from analytic_results.json"""
    path = FS.path_analytic_rho
    assert os.path.isfile(path), f"This file doesn't exist:\n{path:s}"
    with open(path, "r") as file:
        in_dict = json.loads(file.read())
    assert in_dict["hash_vib"] == FS.hash_vib, "wrong vib hash"
    assert in_dict["hash_rho"] == FS.hash_rho, "wrong rho hash"
    temperature = f"{T:.2f}"
    assert temperature in in_dict.keys(), "no analytical results for temperature {:s} in file {:s}".format(temperature, path)
    analytic["Z"] = in_dict[temperature]["Z_sampling"]
    analytic["E"] = in_dict[temperature]["E_sampling"]
    analytic["Cv"] = in_dict[temperature]["Cv_sampling"]
    analytic["alpha_plus"] = analytic["Z"] / in_dict[temperature]["Z_sampling+beta"]
    analytic["alpha_minus"] = analytic["Z"] / in_dict[temperature]["Z_sampling-beta"]
