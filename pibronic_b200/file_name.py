"""File-name contract shared with pibronic.stats / pibronic.plotting (pibronic/data/file_name.py:38-60, 150-180).

Only the names the PIMC hot path reads or writes are kept."""


def pimc(P="{P:d}", T="{T:.2f}", J="{J:d}"):
    """results of one PIMC shard: P beads, temperature T (2 decimals), job number J"""
    return "P{:s}_T{:s}_J{:s}_data_points.npz".format(P, T, J)


def jackknife(P="{P:d}", T="{T:.2f}", X="{X:d}"):
    """output of postprocess + jackknife for X samples"""
    return "P{:s}_T{:s}_X{:s}_thermo".format(P, T, X)


coupled_model = "coupled_model.json"
harmonic_model = "harmonic_model.json"
sampling_model = "sampling_model.json"
analytic_results = "analytic_results.json"
