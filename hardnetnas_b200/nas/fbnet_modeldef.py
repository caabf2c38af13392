"""Sampled architectures recorded by the reference search (hardnetNAS/fbnet_building_blocks/fbnet_modeldef.py:30-95).
Only the op lists matter; the reference's `block_cfg` channel numbers are stale (real shapes come from SEARCH_SPACE2)."""

MODEL_ARCH = {
    "wang2": {"block_op_type": [["ir_k3_e1"], ["ir_k5_e1"], ["ir_k5_s2"], ["ir_k3_s2"], ["ir_k5_e1"], ["skip"]]},
    "wang3": {"block_op_type": [["ir_k5_e1"], ["skip"], ["ir_k5_e1"], ["skip"], ["skip"], ["skip"]]},
    "wang4": {"block_op_type": [["skip"], ["skip"], ["ir_k5_s2"], ["ir_k3_s2"], ["ir_k5_e1"], ["ir_k5_e1"]]},
}


def arch_ops(name: str) -> list[str]:
    return [layer[0] for layer in MODEL_ARCH[name]["block_op_type"]]
