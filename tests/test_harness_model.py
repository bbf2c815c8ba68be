"""The harness model (harness/dsrl_model.py) is state-dict compatible with, and numerically equal to, the reference's
``models.DSRL`` -- checked only where the reference tree is mounted (the build container); skipped on the GPU box."""
import os
import sys

import pytest
import torch

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted here")
def test_state_dict_and_forward_match_reference():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    try:
        from models import DSRL as RefDSRL
        import datasets.Cityscapes.settings as cs
    finally:
        sys.path.remove(REF)
    from harness.dsrl_model import DSRL
    torch.manual_seed(0)
    ref = RefDSRL(3, cs).eval()
    mine = DSRL(3, cs.NUM_CLASSES).eval()
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs.keys()) == list(ms.keys())
    assert all(rs[k].shape == ms[k].shape for k in rs)
    mine.load_state_dict(rs)                      # reference weights load unchanged
    x = torch.randn(2, 3, 64, 128)
    with torch.no_grad():
        a, b = ref(x), mine(x)
    for u, v in zip(a, b):
        assert u.shape == v.shape
        torch.testing.assert_close(u, v, rtol=1e-5, atol=1e-5)
    assert b[2].shape == (2, 1, 16, 32) and b[3].shape == (2, 1, 16, 32)      # FA inputs: (B, 1, Hout/8, Wout/8)
