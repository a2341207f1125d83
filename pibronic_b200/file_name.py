"""File-name contract shared with pibronic.stats / pibronic.plotting (pibronic/data/file_name.py:38-60, 150-180).

Only the names the PIMC hot path reads or writes are kept."""


def pimc(P="{P:d}", T="{T:.2f}", J="{J:d}"):
    """results of one PIMC shard: P beads, temperature T (2 decimals), job number J"""
    return "P{:s}_T{:s}_J{:s}_data_points.npz".format(P, T, J)


def jackknife(P="{P:d}", T="{T:.2f}", X="{X:d}"):
    """output of postprocess + jackknife for X samples"""
    return "P{:s}_T{:s}_X{:s}_thermo".format(P, T, X)


def training_data_input(P="{P:d}", T="{T:.2f}", J="{J:d}"):
    """bead co-ordinates of a g(R) / rho(R) run (training sets for ML models, file_name.py:98-106)"""
    return "P{:s}_T{:s}_J{:s}_training_data_input.npz".format(P, T, J)


def training_data_g_output(P="{P:d}", T="{T:.2f}", J="{J:d}"):
    """g(R) of a training-set run (file_name.py:114-122)"""
    return "P{:s}_T{:s}_J{:s}_training_data_g_output.npz".format(P, T, J)


def training_data_rho_output(P="{P:d}", T="{T:.2f}", J="{J:d}"):
    """rho(R) of a training-set run (file_name.py:130-138)"""
    return "P{:s}_T{:s}_J{:s}_training_data_rho_output.npz".format(P, T, J)


coupled_model = "coupled_model.json"
harmonic_model = "harmonic_model.json"
sampling_model = "sampling_model.json"
analytic_results = "analytic_results.json"
