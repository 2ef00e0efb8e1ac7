"""-m gpu: the second tcgen05 convolution formulation (conv_tc2.cu: weights = M operand, pixels = N operand)
against torch float64 convolutions on bf16-exact operands.  Covers forward 5x5 (64/128 input channels, sizes
that exercise one and two tiles per row and ragged edges), stride 2 over the space-to-depth tensor, 1x1 with fp32
NCHW output, the input gradient in PAIR (64 outputs) and WIDE (128 outputs) mode with ReLU mask and accumulation,
and the stride-2 input gradient with its scatter epilogue."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from deepsensornz_b200 import _cabi
from deepsensornz_b200.engine import _Blk
from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _S():
    return torch.cuda.current_stream().cuda_stream


def _to_blk(x, cb_total=None):
    B, Cc, H, W = x.shape
    blk = _Blk(B, cb_total or Cc // 8, H, W, x.device)
    _cabi.call("cnp_blk_from_nchw_f32", x.data_ptr(), x.stride(0), B, Cc, H, W, C.byref(blk.view()), _S())
    return blk


def _from_blk(blk, Cc, cb_off=0):
    out = torch.empty(blk.B, Cc, blk.H, blk.W, device="cuda")
    _cabi.call("cnp_blk_to_nchw_f32", C.byref(blk.view(cb_off)), blk.B, Cc, out.data_ptr(), out.stride(0), _S())
    return out


def _pack(wt, kind, n_chunks, py=0, px=0, co_off=0, n_out=64):
    nbytes = _cabi.lib().cnp_conv_tc2_packed_bytes(kind, n_chunks, n_out)
    assert nbytes > 0
    wpk = torch.empty(nbytes // 2, dtype=torch.bfloat16, device="cuda")
    co, ci, k, _ = wt.shape
    _cabi.call("cnp_conv_tc2_pack", wt.data_ptr(), co, ci, k, kind, n_chunks, py, px, co_off, n_out, wpk.data_ptr(), _S())
    return wpk


def _out(blk_view, bias=None, relu=0, scatter=(1, 0, 1, 0), accumulate=0, mask=None):
    o = _cabi.CnpConvOut()
    o.mode, o.blk = 0, blk_view
    o.sy, o.ay, o.sx, o.ax = scatter
    o.bias, o.relu, o.accumulate = (bias.data_ptr() if bias is not None else None), relu, accumulate
    o.mask = C.pointer(mask) if mask is not None else None
    o._keep = (bias, mask)
    return o


def _pad_is_zero(blk, B, cb, h, w):
    full = blk.t[:B * blk.bstride].view(B, cb, h + 4, w + 4, 8).float()
    return (float(full[:, :, :2].abs().max()) == 0 and float(full[:, :, :, :2].abs().max()) == 0 and
            float(full[:, :, -2:].abs().max()) == 0 and float(full[:, :, :, -2:].abs().max()) == 0)


@pytest.mark.parametrize("cin,h,w", [(64, 38, 38), (128, 76, 76), (64, 152, 160), (128, 61, 45), (64, 33, 304),
                                     (128, 20, 300)])
def test_conv_tc2_k5s1(cin, h, w):
    torch.manual_seed(4)
    B = 2
    x = torch.randn(B, cin, h, w, device="cuda").bfloat16().float()
    wt = (torch.randn(64, cin, 5, 5, device="cuda") * 0.05).bfloat16().float()
    b = torch.randn(64, device="cuda")
    ref = F.relu(F.conv2d(x.double(), wt.double(), b.double(), padding=2))
    xb = _to_blk(x)
    yb = _Blk(B, 8, h, w, x.device)
    o = _out(yb.view(), bias=b, relu=1)
    _cabi.call("cnp_conv_tc2", C.byref(xb.view()), cin // 8, _pack(wt, _cabi.KIND_K5S1, cin // 8).data_ptr(),
               _cabi.KIND_K5S1, 0, 0, 64, C.byref(o), B, _S())
    y = _from_blk(yb, 64)
    torch.cuda.synchronize()
    assert rel_err(y, ref) < 1e-2   # bf16 output rounding only (operands are exact in bf16)
    assert _pad_is_zero(yb, B, 8, h, w)


def test_conv_tc2_stride2_and_1x1():
    torch.manual_seed(5)
    B, h, w = 2, 76, 80
    x = torch.randn(B, 64, h, w, device="cuda").bfloat16().float()
    wt = (torch.randn(64, 64, 5, 5, device="cuda") * 0.05).bfloat16().float()
    b = torch.randn(64, device="cuda")
    ref = F.conv2d(x.double(), wt.double(), b.double(), stride=2, padding=2)
    xb = _to_blk(x)
    ph = _Blk(B, 32, h // 2, w // 2, x.device)
    _cabi.call("cnp_blk_space_to_depth", C.byref(xb.view()), 8, C.byref(ph.view()), B, _S())
    yb = _Blk(B, 8, h // 2, w // 2, x.device)
    o = _out(yb.view(), bias=b)
    _cabi.call("cnp_conv_tc2", C.byref(ph.view()), 32, _pack(wt, _cabi.KIND_K5S2, 32).data_ptr(), _cabi.KIND_K5S2, 0, 0,
               64, C.byref(o), B, _S())
    assert rel_err(_from_blk(yb, 64), ref) < 1e-2
    assert _pad_is_zero(yb, B, 8, h // 2, w // 2)
    # 1x1 with fp32 NCHW output (aligned fast path: w = 80; ragged path: w = 77)
    for ww in (80, 77):
        x1 = x[..., :ww].contiguous()
        w1 = (torch.randn(64, 64, 1, 1, device="cuda") * 0.1).bfloat16().float()
        ref1 = F.conv2d(x1.double(), w1.double(), b.double())
        z = torch.empty(B, 64, h, ww, device="cuda")
        o = _cabi.CnpConvOut()
        o.mode, o.f32, o.f32_bstride, o.f32_ch_off = 1, z.data_ptr(), z.stride(0), 0
        o.sy, o.ay, o.sx, o.ax = 1, 0, 1, 0
        o.bias = b.data_ptr()
        x1b = _to_blk(x1)   # keep alive: a temporary would be recycled by the allocator before the launch
        _cabi.call("cnp_conv_tc2", C.byref(x1b.view()), 8, _pack(w1, _cabi.KIND_K1, 8).data_ptr(), _cabi.KIND_K1,
                   0, 0, 64, C.byref(o), B, _S())
        assert rel_err(z, ref1) < 1e-5 * 50  # fp32 accumulate of exact bf16 products


@pytest.mark.parametrize("cin,wide", [(64, False), (128, False), (128, True)])
def test_conv_tc2_dgrad_s1(cin, wide):
    torch.manual_seed(6)
    B, h, w = 2, 45, 52
    dy = torch.randn(B, 64, h, w, device="cuda").bfloat16().float()
    wt = (torch.randn(64, cin, 5, 5, device="cuda") * 0.05).bfloat16().float()
    act = torch.randn(B, cin, h, w, device="cuda").bfloat16().float()       # ReLU mask source
    prev = torch.randn(B, cin, h, w, device="cuda").bfloat16().float()      # accumulate target
    xd = torch.zeros(B, cin, h, w, device="cuda", dtype=torch.double, requires_grad=True)
    F.conv2d(xd, wt.double(), None, padding=2).backward(dy.double())
    ref = prev.double() + xd.grad * (act > 0)
    dyb, actb, dxb = _to_blk(dy), _to_blk(act), _to_blk(prev)
    if wide:
        o = _out(dxb.view(0), mask=actb.view(0), accumulate=1)
        _cabi.call("cnp_conv_tc2", C.byref(dyb.view()), 8, _pack(wt, _cabi.KIND_K5S1_DGRAD, 8, n_out=128).data_ptr(),
                   _cabi.KIND_K5S1_DGRAD, 0, 0, 128, C.byref(o), B, _S())
    else:
        for g in range(cin // 64):
            o = _out(dxb.view(8 * g), mask=actb.view(8 * g), accumulate=1)
            _cabi.call("cnp_conv_tc2", C.byref(dyb.view()), 8,
                       _pack(wt, _cabi.KIND_K5S1_DGRAD, 8, 0, 0, 64 * g).data_ptr(), _cabi.KIND_K5S1_DGRAD, 0, 0, 64,
                       C.byref(o), B, _S())
    assert rel_err(_from_blk(dxb, cin), ref) < 1e-2
    assert _pad_is_zero(dxb, B, cin // 8, h, w)


def test_conv_tc2_dgrad_s2_and_1x1():
    torch.manual_seed(7)
    B, h, w = 2, 40, 48   # input size; dy is h/2 x w/2
    dy = torch.randn(B, 64, h // 2, w // 2, device="cuda").bfloat16().float()
    wt = (torch.randn(64, 64, 5, 5, device="cuda") * 0.05).bfloat16().float()
    xd = torch.zeros(B, 64, h, w, device="cuda", dtype=torch.double, requires_grad=True)
    F.conv2d(xd, wt.double(), None, stride=2, padding=2).backward(dy.double())
    dyb = _to_blk(dy)
    dxb = _Blk(B, 8, h, w, dy.device)
    for py in (0, 1):
        for px in (0, 1):
            o = _out(dxb.view(0), scatter=(2, py, 2, px))
            _cabi.call("cnp_conv_tc2", C.byref(dyb.view()), 8,
                       _pack(wt, _cabi.KIND_K5S2_DGRAD, 8, py, px).data_ptr(), _cabi.KIND_K5S2_DGRAD, py, px, 64,
                       C.byref(o), B, _S())
    assert rel_err(_from_blk(dxb, 64), xd.grad) < 1e-2
    assert _pad_is_zero(dxb, B, 8, h, w)
    # 1x1 dgrad
    w1 = (torch.randn(64, 64, 1, 1, device="cuda") * 0.1).bfloat16().float()
    dz = torch.randn(B, 64, h, w, device="cuda").bfloat16().float()
    xd1 = torch.zeros(B, 64, h, w, device="cuda", dtype=torch.double, requires_grad=True)
    F.conv2d(xd1, w1.double(), None).backward(dz.double())
    d1 = _Blk(B, 8, h, w, dy.device)
    o = _out(d1.view(0))
    dzb = _to_blk(dz)
    _cabi.call("cnp_conv_tc2", C.byref(dzb.view()), 8, _pack(w1, _cabi.KIND_K1_DGRAD, 8).data_ptr(),
               _cabi.KIND_K1_DGRAD, 0, 0, 64, C.byref(o), B, _S())
    assert rel_err(_from_blk(d1, 64), xd1.grad) < 1e-2


def test_conv_tc2_cluster_multicast_matches_plain():
    """Clusters of two CTAs sharing the weight stream by multicast give bit-identical outputs (also with an odd number
    of tiles, where one CTA of a pair only streams weights in the last round)."""
    torch.manual_seed(9)
    B, cin, h, w = 3, 128, 70, 300
    x = torch.randn(B, cin, h, w, device="cuda").bfloat16().float()
    wt = (torch.randn(64, cin, 5, 5, device="cuda") * 0.05).bfloat16().float()
    b = torch.randn(64, device="cuda")
    xb = _to_blk(x)
    wpk = _pack(wt, _cabi.KIND_K5S1, cin // 8)
    outs = []
    try:
        for cluster in (1, 2):
            _cabi.call("cnp_conv_tc2_set_cluster", cluster)
            yb = _Blk(B, 8, h, w, x.device)
            o = _out(yb.view(), bias=b, relu=1)
            _cabi.call("cnp_conv_tc2", C.byref(xb.view()), cin // 8, wpk.data_ptr(), _cabi.KIND_K5S1, 0, 0, 64, C.byref(o), B, _S())
            outs.append(_from_blk(yb, 64))
    finally:
        _cabi.call("cnp_conv_tc2_set_cluster", 1)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    ref = torch.relu(torch.nn.functional.conv2d(x.double(), wt.double(), b.double(), padding=2))
    assert rel_err(outs[0], ref) < 1e-2


@pytest.mark.parametrize("h,w", [(40, 48), (76, 152), (22, 310)])
def test_conv_tc2_dgrad_s2_x_phase_pair(h, w):
    """Both x-phases of an output row phase in ONE launch (px = 2: lane group g = x-phase g, adjacent output pixels):
    two launches give the full stride-2 input gradient, with ReLU mask and accumulation, equal to the four-launch form."""
    torch.manual_seed(9)
    B = 2
    dy = torch.randn(B, 64, h // 2, w // 2, device="cuda").bfloat16().float()
    wt = (torch.randn(64, 64, 5, 5, device="cuda") * 0.05).bfloat16().float()
    act = torch.randn(B, 64, h, w, device="cuda").bfloat16().float()
    old = torch.randn(B, 64, h, w, device="cuda").bfloat16().float()
    xd = torch.zeros(B, 64, h, w, device="cuda", dtype=torch.double, requires_grad=True)
    F.conv2d(xd, wt.double(), None, stride=2, padding=2).backward(dy.double())
    dyb, actb = _to_blk(dy), _to_blk(act)
    plain = _Blk(B, 8, h, w, dy.device)
    for py in (0, 1):
        o = _out(plain.view(0), scatter=(2, py, 2, 0))
        _cabi.call("cnp_conv_tc2", C.byref(dyb.view()), 8, _pack(wt, _cabi.KIND_K5S2_DGRAD, 8, py, 2).data_ptr(),
                   _cabi.KIND_K5S2_DGRAD, py, 2, 64, C.byref(o), B, _S())
    assert rel_err(_from_blk(plain, 64), xd.grad) < 1e-2
    assert _pad_is_zero(plain, B, 8, h, w)
    # mask + accumulate: identical (bit for bit) to the four single-phase launches
    a4, a2 = _to_blk(old), _to_blk(old)
    mview = actb.view(0)
    for py in (0, 1):
        for px in (0, 1):
            o = _out(a4.view(0), scatter=(2, py, 2, px), accumulate=1, mask=mview)
            _cabi.call("cnp_conv_tc2", C.byref(dyb.view()), 8, _pack(wt, _cabi.KIND_K5S2_DGRAD, 8, py, px).data_ptr(),
                       _cabi.KIND_K5S2_DGRAD, py, px, 64, C.byref(o), B, _S())
        o = _out(a2.view(0), scatter=(2, py, 2, 0), accumulate=1, mask=mview)
        _cabi.call("cnp_conv_tc2", C.byref(dyb.view()), 8, _pack(wt, _cabi.KIND_K5S2_DGRAD, 8, py, 2).data_ptr(),
                   _cabi.KIND_K5S2_DGRAD, py, 2, 64, C.byref(o), B, _S())
    assert torch.equal(_from_blk(a4, 64), _from_blk(a2, 64))
    ref = xd.grad * (act > 0) + old.double()
    assert rel_err(_from_blk(a2, 64), ref) < 1e-2


@pytest.mark.parametrize("cin,h,w", [(128, 19, 23), (64, 38, 152), (16, 8, 40)])
def test_conv_tc2_up_phase_kind_matches_polyphase_reference(cin, h, w):
    """Building block for the polyphase resize-convolution (tools/polyphase_check.py): the up-phase kind
    computes one row phase (both x-phases) of conv5x5(bilinear_up2x(x)) as a 4x4 convolution of the replicate-padded
    low-resolution input.  Two launches == four phase convolutions in float64 (same bf16-rounded phase weights), and, away
    from the border, == Upsample + Conv2d itself."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import polyphase_check as P
    torch.manual_seed(13)
    B = 2
    x = torch.randn(B, cin, h, w, device="cuda").bfloat16().float()
    w5 = (torch.randn(64, cin, 5, 5, device="cuda") * 0.05)
    bias = torch.randn(64, device="cuda") * 0.1
    wp = torch.empty(2, 2, 64, cin, 4, 4, device="cuda")
    _cabi.call("cnp_up_phase_weights", w5.data_ptr(), 64, cin, wp.data_ptr(), _S())
    assert rel_err(wp, P.phase_weights(w5.double().cpu())) < 1e-6
    # blocked input whose 2-pixel pad holds the REPLICATED border
    xb = _Blk(B, cin // 8, h, w, x.device)
    xt = F.pad(x, (2, 2, 2, 2), mode="replicate")                                  # [B,C,h+4,w+4]
    xb.t[:B * xb.bstride].view(B, cin // 8, h + 4, w + 4, 8).copy_(
        xt.view(B, cin // 8, 8, h + 4, w + 4).permute(0, 1, 3, 4, 2).bfloat16())
    y = _Blk(B, 8, 2 * h, 2 * w, x.device)
    for a in (0, 1):
        nbytes = _cabi.lib().cnp_conv_tc2_packed_bytes(_cabi.KIND_UP_PHASE, cin // 8, 64)
        wpk = torch.empty(nbytes // 2, dtype=torch.bfloat16, device="cuda")
        wa = wp[a].contiguous()
        _cabi.call("cnp_conv_tc2_pack", wa.data_ptr(), 64, cin, 4, _cabi.KIND_UP_PHASE, cin // 8, a, 0, 0, 64,
                   wpk.data_ptr(), _S())
        o = _out(y.view(0), bias=bias, relu=1, scatter=(2, a, 2, 0))
        _cabi.call("cnp_conv_tc2", C.byref(xb.view()), cin // 8, wpk.data_ptr(), _cabi.KIND_UP_PHASE, a, 0, 64,
                   C.byref(o), B, _S())
    got = _from_blk(y, 64)
    # same arithmetic in float64: bf16-rounded phase weights, replicate-padded input, four 4x4 phase convolutions
    wpb = wp.bfloat16().double()
    ref = torch.zeros(B, 64, 2 * h, 2 * w, device="cuda", dtype=torch.double)
    xt64 = xt.double()
    for a in (0, 1):
        for b in (0, 1):
            ref[:, :, a::2, b::2] = F.conv2d(xt64[:, :, a:a + h + 3, b:b + w + 3], wpb[a, b])
    ref = torch.relu(ref + bias.double()[None, :, None, None])
    assert rel_err(got, ref) < 4e-3                     # bf16 rounding of the stored output only
    assert _pad_is_zero(y, B, 8, 2 * h, 2 * w)
    # and it IS the resize-convolution away from the border (where the reference zero-pads the upsampled tensor)
    true = torch.relu(F.conv2d(F.interpolate(x.double(), scale_factor=2, mode="bilinear", align_corners=False),
                               w5.double(), bias.double(), padding=2))
    assert rel_err(got[:, :, 2:-2, 2:-2], true[:, :, 2:-2, 2:-2]) < 1e-2
