"""-m gpu: the memory-bound helpers on blocked bf16 tensors (conv_bf16.cu) against torch float64: bilinear x2
upsampling forward / backward (``torch.nn.Upsample(scale_factor=2, mode="bilinear")`` inside upstream neuralprocesses'
UNet resize-convolutions, coders/nn.py) with ReLU mask and accumulation, odd sizes, and both backward kernels."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from deepsensornz_b200 import _cabi
from deepsensornz_b200.engine import _Blk
from tests.test_conv_tc2_gpu import _from_blk, _pad_is_zero, _S, _to_blk
from tests.util import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("c,h,w", [(16, 19, 23), (128, 38, 38), (64, 5, 152), (8, 1, 1), (8, 3, 2)])
def test_blk_upsample_fwd(c, h, w):
    torch.manual_seed(1)
    B = 2
    x = torch.randn(B, c, h, w, device="cuda").bfloat16().float()
    xb = _to_blk(x)
    yb = _Blk(B, c // 8, 2 * h, 2 * w, x.device)
    _cabi.call("cnp_blk_upsample2x_fwd", C.byref(xb.view()), c // 8, C.byref(yb.view()), B, _S())
    ref = F.interpolate(x.double(), scale_factor=2, mode="bilinear", align_corners=False)
    assert rel_err(_from_blk(yb, c), ref) < 4e-3        # bf16 rounding of the stored result
    assert _pad_is_zero(yb, B, c // 8, 2 * h, 2 * w)


@pytest.mark.parametrize("mask,accumulate", [(False, False), (True, False), (True, True)])
@pytest.mark.parametrize("c,h,w", [(16, 19, 23), (128, 38, 38), (64, 5, 152), (8, 1, 1), (8, 3, 2), (8, 2, 1030), (8, 19, 300), (16, 17, 9), (8, 61, 40), (8, 103, 24)])
def test_blk_upsample_bwd(c, h, w, mask, accumulate):
    """dx = J^T dy (* (act > 0)) (+ dx_old).  W = 1030 takes the per-pixel kernel, the others the row kernel (8 rows per block, partial last block)."""
    torch.manual_seed(2)
    B = 2
    dy = torch.randn(B, c, 2 * h, 2 * w, device="cuda").bfloat16().float()
    act = torch.randn(B, c, h, w, device="cuda").bfloat16().float()
    old = torch.randn(B, c, h, w, device="cuda").bfloat16().float()
    xd = torch.zeros(B, c, h, w, device="cuda", dtype=torch.double, requires_grad=True)
    F.interpolate(xd, scale_factor=2, mode="bilinear", align_corners=False).backward(dy.double())
    ref = xd.grad
    if mask:
        ref = ref * (act > 0)
    if accumulate:
        ref = ref + old.double()
    dyb, actb = _to_blk(dy), _to_blk(act)
    dxb = _to_blk(old) if accumulate else _Blk(B, c // 8, h, w, dy.device)
    _cabi.call("cnp_blk_upsample2x_bwd", C.byref(dyb.view()), c // 8, C.byref(dxb.view()),
               C.byref(actb.view()) if mask else None, int(accumulate), B, _S())
    assert rel_err(_from_blk(dxb, c), ref) < 4e-3
    assert _pad_is_zero(dxb, B, c // 8, h, w)
