"""GPU: the tcgen05 weight-gradient kernel (default since its first hardware run: 20 / 20 cases green, round 2) (csrc/conv_wgrad_tc.cu: dy and the input as MN-major UMMA operands from TMA tensor
boxes, split over pixel tiles, ordered fp32 reduce) against the oracle.  In a file of its own, sorted after every other suite: a
fault in a freshly written tensor-core kernel must not take the other GPU tests with it.  mode 1 = three N=64 row-tap
instructions per column shift, mode 2 = row taps fused into one N=192 instruction (overlapping LBO chunks)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import train_oracle as TO

pytestmark = pytest.mark.gpu

SHAPES = [
    # b, h, w, cin, cout, k
    (2, 24, 24, 64, 128, 3),        # one chunk, one co tile, 8-column pixel tiles
    (2, 16, 32, 64, 192, 3),        # 16-column tiles, ragged co tile (192 = 128 + 64)
    (3, 20, 24, 96, 136, 3),        # ragged ci chunk (96 = 64 + 32), ragged rows (20 = 2.5 tiles), ragged co
    (2, 48, 48, 256, 192, 3),       # head-like
    (1, 96, 96, 64, 256, 3),        # stage-2-like
    (4, 24, 24, 512, 256, 1),       # 1x1: two 256-channel blocks
    (2, 48, 48, 192, 768, 1),       # MBConv expand
    (2, 48, 48, 768, 192, 1),       # MBConv project (12 chunks -> 3 blocks)
    (1, 1, 640, 104, 2048, 1),      # decoder Linear rows: [640, 104] -> 2048, ragged ci
    (5, 7, 9, 72, 24, 1),           # small ragged everything (315 rows)
]


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("b,h,w,cin,cout,k", SHAPES)
def test_tcgen05_weight_gradient(mode, b, h, w, cin, cout, k):
    from findtextcenternet_b200 import _lib, _ops
    g = torch.Generator().manual_seed(b * 1000 + h * 10 + cin + cout + k)
    x = torch.randn(b, h, w, cin, generator=g).to(torch.bfloat16)
    dy = torch.randn(b, h, w, cout, generator=g).to(torch.bfloat16)
    lib = _lib.load()
    lib.ftc_debug_set_wgrad_tc(0)
    assert int(lib.ftc_train_conv2d_wgrad_scratch_bytes(b, h, w, cin, cout, k, 1)) == 0      # switched off: the mma.sync kernel
    lib.ftc_debug_set_wgrad_tc(mode)
    try:
        assert int(lib.ftc_train_conv2d_wgrad_scratch_bytes(b, h, w, cin, cout, k, 1)) > 0, "shape not taken by the tcgen05 kernel"
        dw = _ops.conv2d_wgrad(x.cuda(), dy.cuda(), k, 1)
        dw2 = _ops.conv2d_wgrad(x.cuda(), dy.cuda(), k, 1)
        torch.cuda.synchronize()
    finally:
        lib.ftc_debug_set_wgrad_tc(-1)
    assert torch.equal(dw, dw2)                                                               # ordered reduce: bit-reproducible
    ref = TO.conv2d_wgrad(x.float(), dy.float(), k, 1)
    assert rel_l2(dw.cpu(), ref) < 2e-5
