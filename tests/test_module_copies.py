"""The drop-in modules hold a ctypes engine handle; copies (deepcopy, pickle, torch.save(model), DataParallel replicas)
must drop it and re-create their own lazily instead of failing or sharing it (CPU-only checks with a stand-in engine)."""
import copy
import ctypes
import io
import pickle

import pytest
import torch

from hardnetnas_b200.hardnet import HardNet
from hardnetnas_b200.nas import SampledDescriptorNet


class _FakeEngine:
    def __init__(self):
        self.handle = ctypes.c_void_p(1)   # ctypes pointers cannot be pickled
        self.device = None

    def close(self):
        pass


@pytest.mark.parametrize("make", [HardNet, lambda: SampledDescriptorNet("wang2")])
def test_copies_drop_the_engine(make):
    m = make()
    m._engine, m._packed_key = _FakeEngine(), ("stale",)
    c = copy.deepcopy(m)
    assert c._engine is None and c._packed_key is None
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), c.state_dict().values()))
    pickle.loads(pickle.dumps(m))
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    r = torch.load(buf, weights_only=False)
    assert r._engine is None and set(r.state_dict()) == set(m.state_dict())
    rep = m._replicate_for_data_parallel()
    assert rep._engine is None and m._engine is not None
    assert m.repack() is m and m._packed_key is None
    m._engine = None


def test_stage_table_follows_the_launch_counts():
    """bench.py attributes FLOPs per timed stage: with conv3 + conv4 fused into one kernel (the default engine) stage 2
    carries both layers' MACs and stage 3 none; with separate kernels the table is the per-layer one."""
    names, macs = HardNet.stage_table([0, 28, 28, 0, 28, 28, 1])
    assert names[2] == "conv3_conv4_fused" and macs[2] == 4718592 + 9437184 and macs[3] == 0
    assert sum(macs[1:]) == sum(HardNet.STAGE_MACS[1:])
    names, macs = HardNet.stage_table([0, 28, 28, 28, 28, 28, 1])
    assert tuple(names) == HardNet.STAGE_NAMES and tuple(macs) == HardNet.STAGE_MACS


def test_assert_unit_norm_guards_the_matching_precondition():
    from hardnetnas_b200.matching import assert_unit_norm
    d = torch.nn.functional.normalize(torch.randn(64, 128), dim=1)
    d[5] = 0                                   # a constant patch's descriptor
    assert assert_unit_norm(d) is d
    d[7] *= 1.5
    with pytest.raises(ValueError, match="row 7"):
        assert_unit_norm(d)
