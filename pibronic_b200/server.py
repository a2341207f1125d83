"""The JSON contract of a PIMC job: ``BoxData.from_json_string`` / ``json_serialize`` exchange one JSON object whose keys
are the execution-parameter names of the reference (the values of its ``ServerExecutionParameters`` enum,
pibronic/server/server.py:11-23, written by job_boss.py and read back in pimc.py:543-593).  Only that naming contract is
kept here -- the SLURM glue around it is out of scope -- as one table that also says which BoxData attribute each key fills.
"""
from enum import Enum

# (enum member, JSON key, BoxData attribute)
_CONTRACT = (
    ("X", "number_of_samples", "samples"),
    ("nBlk", "number_of_blocks", "blocks"),
    ("A", "number_of_states", "states"),
    ("P", "number_of_beads", "beads"),
    ("N", "number_of_modes", "modes"),
    ("T", "temperature", "temperature"),
    ("BlkS", "block_size", "block_size"),
    ("dB", "delta_beta", "delta_beta"),
    ("D", "id_data", "id_data"),
    ("R", "id_rho", "id_rho"),
    ("beta", "beta", "beta"),
    ("tau", "tau", "tau"),
)

ServerExecutionParameters = Enum("ServerExecutionParameters", [(member, key) for member, key, _ in _CONTRACT])
ServerExecutionParameters.__doc__ = "member.value is the JSON key; same members and values as the reference's enum"

#: BoxData attribute filled by each parameter
ATTRIBUTE_OF = {ServerExecutionParameters[member]: attribute for member, _, attribute in _CONTRACT}
