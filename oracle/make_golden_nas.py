"""Golden vectors for the NAS descriptor nets, produced by the UNMODIFIED reference blocks
(hardnetNAS/fbnet_building_blocks + the stem/head definition of model_supernet.py). Run via oracle/make_golden.py."""
from __future__ import annotations

import os
import sys
from collections import OrderedDict
from pathlib import Path

import numpy as np
import torch
from torch import nn

REPO = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/hardnetNAS")
sys.path.insert(0, str(REPO))
from oracle import synth  # noqa: E402

ARCHS = {
    "wang2": ["ir_k3_e1", "ir_k5_e1", "ir_k5_s2", "ir_k3_s2", "ir_k5_e1", "skip"],
    "wang3": ["ir_k5_e1", "skip", "ir_k5_e1", "skip", "skip", "skip"],
    "wang4": ["skip", "skip", "ir_k5_s2", "ir_k3_s2", "ir_k5_e1", "ir_k5_e1"],
    # not a recorded arch: exercises expansion 3, 4-group shuffle and squeeze-excite candidates
    "mixed_se": ["ir_k3_e3_se", "ir_k5_s4", "ir_k3_s2_se", "ir_k5_e3", "ir_k3_s4_se", "ir_k3_e1_se"],
}


def _import_reference():
    cwd = os.getcwd()
    os.chdir(REF)  # LookUpTable() reads ./supernet_functions/lookup_table.txt
    sys.path.insert(0, str(REF))
    for m in [k for k in sys.modules if k.split(".")[0] in ("fbnet_building_blocks", "supernet_functions", "general_functions")]:
        del sys.modules[m]
    try:
        from fbnet_building_blocks.fbnet_builder import PRIMITIVES, ConvBNRelu, Flatten
        from supernet_functions.lookup_table_builder import LookUpTable
        table = LookUpTable()
    finally:
        os.chdir(cwd)
        sys.path.remove(str(REF))
    return PRIMITIVES, ConvBNRelu, Flatten, table


class RefSampledNet(nn.Module):
    """`first` + argmax ops + `last_stages` exactly as FBNet_Stochastic_SuperNet builds them (model_supernet.py:57-68)."""

    def __init__(self, ops, PRIMITIVES, ConvBNRelu, Flatten, table):
        super().__init__()
        self.first = ConvBNRelu(input_depth=1, output_depth=32, kernel=3, stride=1, pad=1, no_bias=1, use_relu="relu", bn_type="bn")
        self.stages = nn.ModuleList([PRIMITIVES[name](*table.layers_parameters[i]) for i, name in enumerate(ops)])
        self.last_stages = nn.Sequential(OrderedDict([
            ("conv_k1", nn.Conv2d(table.layers_parameters[-1][1], 128, kernel_size=4, bias=False)),
            ("batchnorm", nn.BatchNorm2d(128, affine=False)),
            ("flatten", Flatten()),
        ]))

    def forward(self, x):
        y = self.first(x)
        for st in self.stages:
            y = st(y)
        y = self.last_stages(y)
        return y / torch.norm(y, p=2, dim=-1, keepdim=True)


def main():
    PRIMITIVES, ConvBNRelu, Flatten, table = _import_reference()
    x = synth.make_patches(32, seed=1234, edge_cases=False)
    out = {}
    for arch, ops in ARCHS.items():
        torch.manual_seed(0)
        net = RefSampledNet(ops, PRIMITIVES, ConvBNRelu, Flatten, table)
        net.load_state_dict(synth.randomize_nas_state(net.state_dict(), 4))
        net.eval()
        with torch.no_grad():
            y = net(x)
        sd = net.state_dict()
        out[f"{arch}_desc"] = y.numpy()
        out[f"{arch}_params"] = np.array([sum(v.numel() for k, v in sd.items() if "num_batches" not in k and "running" not in k)])
        out[f"{arch}_fingerprint"] = np.array([sum(v.double().sum().item() for k, v in sd.items() if v.dtype.is_floating_point),
                                               sum(v.double().abs().sum().item() for k, v in sd.items() if v.dtype.is_floating_point)])
    np.savez_compressed(REPO / "tests" / "golden" / "nas_forward.npz", **out)
    print("nas_forward.npz", {k: (v.shape, v.ravel()[:2]) for k, v in out.items() if "desc" not in k})


if __name__ == "__main__":
    main()
