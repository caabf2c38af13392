"""Search space of the descriptor supernet (hardnetNAS/supernet_functions/lookup_table_builder.py:18-45,78-110).
The GPU latency table of the reference (lookup_table.txt) is search-time only and not needed here."""
from collections import OrderedDict

from .fbnet_builder import PRIMITIVES

CANDIDATE_BLOCKS = ["skip", "ir_k3_e1", "ir_k3_e3", "ir_k3_s4", "ir_k5_e1", "ir_k5_e3", "ir_k5_s4", "ir_k3_e1_se",
                    "ir_k3_e3_se", "ir_k3_s4_se", "ir_k5_e1_se", "ir_k5_e3_se", "ir_k5_s4_se", "ir_k3_s2", "ir_k5_s2",
                    "ir_k3_s2_se", "ir_k5_s2_se"]

SEARCH_SPACE2 = OrderedDict([
    ("input_shape", [(32, 32, 32), (32, 16, 16), (32, 16, 16), (64, 8, 8), (64, 8, 8), (128, 4, 4)]),
    ("channel_size", [32, 32, 64, 64, 128, 128]),
    ("strides", [2, 1, 2, 1, 2, 1]),
])


class LookUpTable:
    """Per-layer constructor arguments (C_in, C_out, -999, stride) and the candidate op constructors."""

    def __init__(self, candidate_blocks=CANDIDATE_BLOCKS, search_space=SEARCH_SPACE2):
        self.cnt_layers = len(search_space["input_shape"])
        self.lookup_table_operations = {name: PRIMITIVES[name] for name in candidate_blocks}
        self.layers_parameters = [(search_space["input_shape"][i][0], search_space["channel_size"][i], -999,
                                   search_space["strides"][i]) for i in range(self.cnt_layers)]
        self.layers_input_shapes = search_space["input_shape"]
