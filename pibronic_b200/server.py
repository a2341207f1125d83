"""Names of the execution parameters passed to a PIMC job as one JSON string
(pibronic/server/server.py:11-23); only the enum is kept, the SLURM glue is out of scope."""
from enum import Enum


class ServerExecutionParameters(Enum):
    X = "number_of_samples"
    nBlk = "number_of_blocks"
    A = "number_of_states"
    P = "number_of_beads"
    N = "number_of_modes"
    T = "temperature"
    BlkS = "block_size"
    dB = "delta_beta"
    D = "id_data"
    R = "id_rho"
    beta = "beta"
    tau = "tau"
