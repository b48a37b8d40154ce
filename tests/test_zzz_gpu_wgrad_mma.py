"""GPU: the staged mma.sync weight-gradient kernel (conv_wgrad_mma_kernel), in a file of its own and sorted after every other
suite: a fault here cannot hide or break the other GPU tests."""
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


@pytest.mark.xfail(strict=False, reason="staged mma.sync weight-gradient kernel: written and emulated without GPU time, this is its "
                                        "first hardware run (XPASS = it works and can become the default)")
@pytest.mark.parametrize("b,h,w,cin,cout,k,stride", [
    (2, 6, 5, 8, 16, 3, 1), (2, 9, 7, 24, 40, 3, 2), (3, 8, 8, 16, 8, 1, 1), (2, 24, 24, 64, 136, 3, 1), (4, 48, 48, 192, 768, 1, 1),
    (2, 13, 11, 264, 72, 3, 1)])
def test_staged_mma_weight_gradient(b, h, w, cin, cout, k, stride):
    """conv_wgrad_mma_kernel (ldmatrix.trans + mma.sync, switched on through ftc_debug_set_wgrad_mma for this test only) against
    the oracle: bf16 operands, fp32 accumulation."""
    from findtextcenternet_b200 import _lib, _ops
    x = rnd(b, h, w, cin, seed=1).to(torch.bfloat16)
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    dy = rnd(b, ho, wo, cout, seed=3).to(torch.bfloat16)
    lib = _lib.load()
    lib.ftc_debug_set_wgrad_mma(1)
    try:
        dw = _ops.conv2d_wgrad(dev(x), dev(dy), k, stride)
        torch.cuda.synchronize()
    finally:
        lib.ftc_debug_set_wgrad_mma(-1)
    assert rel_l2(dw.cpu(), TO.conv2d_wgrad(x.float(), dy.float(), k, stride)) < 2e-5
