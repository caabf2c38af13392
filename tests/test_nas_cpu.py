"""NAS descriptor nets on CPU: the mirror blocks and the oracle against goldens made with the reference's own blocks."""
import numpy as np
import pytest
import torch

from hardnetnas_b200.nas import MODEL_ARCH, SampledDescriptorNet
from hardnetnas_b200.nas.fbnet_modeldef import arch_ops
from oracle import nas_oracle, synth

MIXED = ["ir_k3_e3_se", "ir_k5_s4", "ir_k3_s2_se", "ir_k5_e3", "ir_k3_s4_se", "ir_k3_e1_se"]


def build(arch):
    ops = MIXED if arch == "mixed_se" else arch_ops(arch)
    torch.manual_seed(0)
    net = SampledDescriptorNet(ops)
    net.load_state_dict(synth.randomize_nas_state(net.state_dict(), 4))
    return net, ops


@pytest.mark.parametrize("arch", ["wang2", "wang3", "wang4", "mixed_se"])
def test_mirror_and_oracle_match_reference(arch, golden_dir):
    g = np.load(golden_dir / "nas_forward.npz")
    net, ops = build(arch)
    sd = net.state_dict()
    fp = np.array([sum(v.double().sum().item() for v in sd.values() if v.dtype.is_floating_point),
                   sum(v.double().abs().sum().item() for v in sd.values() if v.dtype.is_floating_point)])
    np.testing.assert_allclose(fp, g[f"{arch}_fingerprint"], rtol=1e-9)       # same init stream as the reference
    x = synth.make_patches(32, 1234, edge_cases=False)
    net.train(False)
    with torch.no_grad():
        y_mirror = net.forward_torch(x)
    np.testing.assert_allclose(y_mirror.numpy(), g[f"{arch}_desc"], atol=2e-6)
    y_oracle = nas_oracle.nas_forward(x, ops, sd)
    np.testing.assert_allclose(y_oracle.numpy(), g[f"{arch}_desc"], atol=2e-6)


def test_param_counts_match_survey():
    # params 290 112 / 278 112 / 321 280 for wang2 / wang3 / wang4
    for arch, n in (("wang2", 290112), ("wang3", 278112), ("wang4", 321280)):
        net = SampledDescriptorNet(arch)
        assert sum(p.numel() for p in net.parameters()) == n


def test_compiled_program_is_well_formed():
    net, _ = build("mixed_se")
    prog = net.compile_program()
    kinds = [op.kind for op in prog.ops]
    assert kinds[0] == 0 and kinds[-1] == 5 and 4 in kinds and 2 in kinds
    for op in prog.ops[1:-1]:
        assert 0 <= op.src <= 2 and 0 <= op.dst <= 2
    blob = torch.cat(prog.params)
    assert blob.numel() == prog.n and torch.isfinite(blob).all()


def test_load_from_supernet_state_dict():
    from hardnetnas_b200.nas import CANDIDATE_BLOCKS
    net, ops = build("wang2")
    sd = net.state_dict()
    fake = {}
    for k, v in sd.items():
        if k.startswith("stages."):
            i, rest = k[len("stages."):].split(".", 1)
            fake[f"module.stages_to_search.{i}.ops.{CANDIDATE_BLOCKS.index(ops[int(i)])}.{rest}"] = v
        else:
            fake["module." + k] = v
    fake["module.stages_to_search.0.thetas"] = torch.zeros(17)
    other = SampledDescriptorNet("wang2")
    other.load_from_supernet({k: v for k, v in fake.items() if "thetas" not in k}, CANDIDATE_BLOCKS)
    for k, v in other.state_dict().items():
        assert torch.equal(v, sd[k])


def test_latency_table_text_format_round_trip(tmp_path):
    """lookup_table_builder.py:160-188: op names on the first line, one line of latencies per searched layer."""
    from hardnetnas_b200.nas.lookup_table import CANDIDATE_BLOCKS, LookUpTable
    t = LookUpTable()
    assert t.lookup_table_latency is None and t.cnt_layers == 6
    t.lookup_table_latency = [{op: float(100 * i + k) + 0.25 for k, op in enumerate(CANDIDATE_BLOCKS)} for i in range(6)]
    path = tmp_path / "lookup_table.txt"
    t._write_lookup_table_to_file(path)
    lines = path.read_text().split("\n")
    assert lines[0] == " ".join(CANDIDATE_BLOCKS) and len(lines) == 7
    assert LookUpTable(path_to_file=path).lookup_table_latency == t.lookup_table_latency
    # a file in the layout the reference ships (supernet_functions/lookup_table.txt: same header, float rows)
    ref_like = tmp_path / "ref.txt"
    ref_like.write_text(" ".join(CANDIDATE_BLOCKS) + "\n" + "\n".join(" ".join(str(2.38 + j + i) for j in range(17)) for i in range(6)))
    got = LookUpTable(path_to_file=ref_like).lookup_table_latency
    assert got[0]["skip"] == 2.38 and got[5]["ir_k5_s2_se"] == 2.38 + 16 + 5
