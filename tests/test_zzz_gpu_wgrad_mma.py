"""GPU: the mma.sync weight-gradient kernel (conv_wgrad_mma_kernel: the path of stride-2 and small-channel convolutions, which
the tcgen05 kernel does not take), in a file of its own.  First hardware runs: round 2 (6 / 6 green three times)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import train_oracle as TO

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def dev(t):
    return t.cuda()


@pytest.mark.parametrize("b,h,w,cin,cout,k,stride", [
    (2, 6, 5, 8, 16, 3, 1), (2, 9, 7, 24, 40, 3, 2), (3, 8, 8, 16, 8, 1, 1), (2, 24, 24, 64, 136, 3, 1), (4, 48, 48, 192, 768, 1, 1),
    (2, 13, 11, 264, 72, 3, 1)])
def test_mma_sync_weight_gradient(b, h, w, cin, cout, k, stride):
    """conv_wgrad_mma_kernel (ldmatrix.trans + mma.sync, switched on through ftc_debug_set_wgrad_mma for this test only) against
    the oracle: bf16 operands, fp32 accumulation."""
    from findtextcenternet_b200 import _lib, _ops
    x = rnd(b, h, w, cin, seed=1).to(torch.bfloat16)
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    dy = rnd(b, ho, wo, cout, seed=3).to(torch.bfloat16)
    lib = _lib.load()
    lib.ftc_debug_set_wgrad_mma(1)
    lib.ftc_debug_set_wgrad_tc(0)            # the tcgen05 kernel would take the stride-1 shapes: this test is about the mma.sync one
    try:
        dw = _ops.conv2d_wgrad(dev(x), dev(dy), k, stride)
        torch.cuda.synchronize()
    finally:
        lib.ftc_debug_set_wgrad_mma(-1)
        lib.ftc_debug_set_wgrad_tc(-1)
    assert rel_l2(dw.cpu(), TO.conv2d_wgrad(x.float(), dy.float(), k, stride)) < 2e-5
