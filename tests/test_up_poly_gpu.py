"""-m gpu: the polyphase resize-convolution (up_poly.cu + conv_tc2 kinds UP_PHASE / UP_PHASE_DGRAD + wgrad kinds
WG_UP_PHASE / WG_K5S1_T, orchestrated by Engine._up_poly_fwd / _up_poly_bwd) against float64 autograd of what it replaces:
``torch.nn.Upsample(scale_factor=2, mode="bilinear")`` + ``Conv2d(128, 64, 5, padding=2)`` + ReLU of upstream
neuralprocesses' UNet decoder levels.  The decomposition itself is checked in float64 on the CPU
(tests/test_polyphase_math.py, tools/polyphase_strips.py); here the kernels: every piece on its own, then the layer
forward / input gradient / weight gradient / bias gradient, including the border band, odd sizes, the ReLU mask, and
against the engine's Upsample + Conv path on the same inputs."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from deepsensornz_b200 import _cabi
from deepsensornz_b200.engine import _Blk
from tests.test_conv_tc2_gpu import _from_blk, _pack, _pad_is_zero, _out, _S, _to_blk
from tests.util import rel_err, small_model

pytestmark = pytest.mark.gpu


def _slack_is_zero(blk):
    """nothing was written past the last image of a blocked buffer (its allocation carries a zeroed over-read slack)"""
    return float(blk.t[blk.B * blk.bstride:].abs().max()) == 0


def _up(x):
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)


def _strip_views(t, B, H2, W2):
    """reference strips of a high-res NCHW tensor: rows [2B,C,6,W2] (image s*B+b), cols [2B,C,6,H2] (transposed)"""
    rows = torch.cat([t[:, :, :6], t[:, :, H2 - 6:]], 0)
    tt = t.transpose(2, 3)
    cols = torch.cat([tt[:, :, :6], tt[:, :, W2 - 6:]], 0)
    return rows, cols


@pytest.mark.parametrize("c,h,w", [(16, 8, 8), (128, 19, 23), (64, 38, 9)])
def test_strips_of_the_upsampled_tensor(c, h, w):
    torch.manual_seed(1)
    B = 2
    x = torch.randn(B, c, h, w, device="cuda").bfloat16().float()
    xb = _to_blk(x)
    rows, cols = _Blk(2 * B, c // 8, 6, 2 * w, x.device), _Blk(2 * B, c // 8, 6, 2 * h, x.device)
    _cabi.call("cnp_up_strips_fwd", C.byref(xb.view()), c // 8, C.byref(rows.view()), C.byref(cols.view()), B, _S())
    r_ref, c_ref = _strip_views(_up(x.double()), B, 2 * h, 2 * w)
    assert rel_err(_from_blk(rows, c), r_ref) < 4e-3 and rel_err(_from_blk(cols, c), c_ref) < 4e-3
    assert _pad_is_zero(rows, 2 * B, c // 8, 6, 2 * w) and _pad_is_zero(cols, 2 * B, c // 8, 6, 2 * h)
    assert _slack_is_zero(rows) and _slack_is_zero(cols)
    # the same bf16 values as the full upsampling kernel (the band must not depend on which path produced it)
    full = _Blk(B, c // 8, 2 * h, 2 * w, x.device)
    _cabi.call("cnp_blk_upsample2x_fwd", C.byref(xb.view()), c // 8, C.byref(full.view()), B, _S())
    fr, fc = _strip_views(_from_blk(full, c), B, 2 * h, 2 * w)
    assert rel_err(_from_blk(rows, c), fr) < 1e-2 and rel_err(_from_blk(cols, c), fc) < 1e-2


@pytest.mark.parametrize("h,w", [(8, 8), (19, 23)])
def test_dy_split_partitions_every_pixel_once(h, w):
    torch.manual_seed(2)
    B = 2
    dy = torch.randn(B, 64, 2 * h, 2 * w, device="cuda").bfloat16().float()
    dyb16 = _Blk(B, 16, 2 * h, 2 * w, dy.device)        # as in the engine: chunks 8..15 of a 16-chunk tensor
    _cabi.call("cnp_blk_from_nchw_f32", dy.data_ptr(), dy.stride(0), B, 64, 2 * h, 2 * w, C.byref(dyb16.view(8)), _S())
    s2d = _Blk(B, 32, h, w, dy.device)
    rows, cols = _Blk(2 * B, 8, 6, 2 * w, dy.device), _Blk(2 * B, 8, 6, 2 * h, dy.device)
    _cabi.call("cnp_up_dy_split", C.byref(dyb16.view(8)), C.byref(s2d.view()), C.byref(rows.view()), C.byref(cols.view()),
               B, _S())
    got = _from_blk(s2d, 256)
    inner = torch.zeros_like(dy)
    inner[:, :, 4:-4, 4:-4] = dy[:, :, 4:-4, 4:-4]
    for a in (0, 1):
        for b in (0, 1):
            ph = (a * 2 + b) * 64
            assert torch.equal(got[:, ph:ph + 64], inner[:, :, a::2, b::2])
    r, c = _from_blk(rows, 64), _from_blk(cols, 64)
    # put the strips back: interior + strips == dy, every pixel exactly once
    back = inner.clone()
    back[:, :, :6] += r[:B]; back[:, :, 2 * h - 6:] += r[B:]
    back[:, :, :, :6] += c[:B].transpose(2, 3); back[:, :, :, 2 * w - 6:] += c[B:].transpose(2, 3)
    assert torch.equal(back, dy)
    assert _pad_is_zero(s2d, B, 32, h, w)
    assert _slack_is_zero(s2d) and _slack_is_zero(rows) and _slack_is_zero(cols)
    assert _pad_is_zero(rows, 2 * B, 8, 6, 2 * w) and _pad_is_zero(cols, 2 * B, 8, 6, 2 * h)


@pytest.mark.parametrize("h,w", [(8, 8), (19, 23), (38, 12)])
def test_phase_dgrad_kind(h, w):
    """KIND_UP_PHASE_DGRAD == sum over phases of the transposed 4x4 phase convolutions (bf16-rounded phase weights)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    torch.manual_seed(3)
    B, cin = 2, 128
    w5 = torch.randn(64, cin, 5, 5, device="cuda") * 0.05
    wp = torch.empty(2, 2, 64, cin, 4, 4, device="cuda")
    _cabi.call("cnp_up_phase_weights", w5.data_ptr(), 64, cin, wp.data_ptr(), _S())
    dys = torch.randn(B, 256, h, w, device="cuda").bfloat16().float()       # the four phase planes of dY
    act = torch.randn(B, cin, h, w, device="cuda").bfloat16().float()
    nbytes = _cabi.lib().cnp_conv_tc2_packed_bytes(_cabi.KIND_UP_PHASE_DGRAD, 32, 128)
    wpk = torch.empty(nbytes // 2, dtype=torch.bfloat16, device="cuda")
    _cabi.call("cnp_conv_tc2_pack", wp.data_ptr(), 64, cin, 4, _cabi.KIND_UP_PHASE_DGRAD, 32, 0, 0, 0, 128,
               wpk.data_ptr(), _S())
    sb, ab = _to_blk(dys), _to_blk(act)
    dx = _Blk(B, 16, h, w, dys.device)
    o = _out(dx.view(0), mask=ab.view(0))
    _cabi.call("cnp_conv_tc2", C.byref(sb.view()), 32, wpk.data_ptr(), _cabi.KIND_UP_PHASE_DGRAD, 0, 0, 128, C.byref(o),
               B, _S())
    wpb = wp.bfloat16().double()
    full = torch.zeros(B, cin, h + 4, w + 4, device="cuda", dtype=torch.double)
    for a in (0, 1):
        for b in (0, 1):
            ph = (a * 2 + b) * 64
            full[:, :, a:a + h + 3, b:b + w + 3] += F.conv_transpose2d(dys[:, ph:ph + 64].double(), wpb[a, b])
    ref = full[:, :, 2:-2, 2:-2] * (act > 0)
    assert rel_err(_from_blk(dx, cin), ref) < 4e-3
    assert _pad_is_zero(dx, B, 16, h, w)


@pytest.mark.parametrize("h,w", [(8, 8), (19, 23), (38, 40)])
def test_phase_wgrad_kind_and_fold(h, w):
    """WG_UP_PHASE == the four low-res 4x4 correlations; cnp_up_wgrad_fold maps them (and a tap-transposed 5x5 part) back."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import polyphase_check as P
    torch.manual_seed(4)
    B, cin = 2, 128
    x = torch.randn(B, cin, h, w, device="cuda").bfloat16().float()
    dys = torch.randn(B, 256, h, w, device="cuda").bfloat16().float()
    xb, sb = _to_blk(x), _to_blk(dys)
    dwp = torch.zeros(2, 2, 64, cin, 4, 4, device="cuda")
    db = torch.zeros(64, device="cuda")
    wsb = _cabi.lib().cnp_conv_tc_wgrad_workspace_bytes()
    ws = torch.empty(wsb // 4, device="cuda")
    for workspace in (True, False):
        dwp.zero_(); db.zero_()
        _cabi.call("cnp_conv_tc_wgrad", C.byref(xb.view()), 16, C.byref(sb.view()), _cabi.WG_UP_PHASE, dwp.data_ptr(),
                   db.data_ptr(), cin, B, ws.data_ptr() if workspace else None, wsb if workspace else 0, _S())
        xz = F.pad(x.double(), (2, 2, 2, 2))
        ref = torch.zeros_like(dwp, dtype=torch.double)
        for a in (0, 1):
            for b in (0, 1):
                ph = (a * 2 + b) * 64
                win = xz[:, :, a:a + h + 3, b:b + w + 3]
                ref[a, b] = F.conv2d(win.transpose(0, 1), dys[:, ph:ph + 64].double().transpose(0, 1)).transpose(0, 1)
        assert rel_err(dwp, ref) < 1e-4, workspace
        assert rel_err(db, dys.double().view(B, 4, 64, h, w).sum((0, 1, 3, 4))) < 1e-4, workspace
    dwt = torch.randn(64, cin, 5, 5, device="cuda")
    dw5 = torch.randn(64, cin, 5, 5, device="cuda")
    fold = P.fold_matrix(torch.float64).cuda()
    want = dw5.double() + torch.einsum("aboipq,akp,blq->oikl", dwp.double(), fold, fold) + dwt.double().transpose(2, 3)
    _cabi.call("cnp_up_wgrad_fold", dwp.data_ptr(), dwt.data_ptr(), 64, cin, dw5.data_ptr(), _S())
    assert rel_err(dw5, want) < 1e-5


def test_wgrad_tap_transposed_kind():
    torch.manual_seed(5)
    B, cin, h, w = 2, 128, 6, 46
    x = torch.randn(B, cin, h, w, device="cuda").bfloat16().float()
    dy = torch.randn(B, 64, h, w, device="cuda").bfloat16().float()
    xb, db_ = _to_blk(x), _to_blk(dy)
    wsb = _cabi.lib().cnp_conv_tc_wgrad_workspace_bytes()
    ws = torch.empty(wsb // 4, device="cuda")
    g, gt = torch.zeros(64, cin, 5, 5, device="cuda"), torch.zeros(64, cin, 5, 5, device="cuda")
    _cabi.call("cnp_conv_tc_wgrad", C.byref(xb.view()), 16, C.byref(db_.view()), _cabi.WG_K5S1, g.data_ptr(), None, cin, B,
               ws.data_ptr(), wsb, _S())
    _cabi.call("cnp_conv_tc_wgrad", C.byref(xb.view()), 16, C.byref(db_.view()), _cabi.WG_K5S1_T, gt.data_ptr(), None, cin,
               B, ws.data_ptr(), wsb, _S())
    assert torch.equal(gt, g.transpose(2, 3).contiguous())
    wd = torch.zeros(64, cin, 5, 5, device="cuda", dtype=torch.double, requires_grad=True)
    F.conv2d(x.double(), wd, padding=2).backward(dy.double())
    assert rel_err(g, wd.grad) < 1e-4


@pytest.mark.parametrize("h,w", [(12, 12), (19, 23), (38, 38), (12, 76), (13, 40)])
def test_up_poly_layer_forward_backward(h, w):
    """The whole decoder level through the engine's orchestration: forward everywhere (band included), dx with the ReLU
    mask of the producer, dW5, dbias -- against float64 autograd and against the engine's Upsample + Conv path."""
    eng = small_model("bf16").engine
    torch.manual_seed(6)
    B, cin = 2, 128
    x = torch.randn(B, cin, h, w, device="cuda").relu().bfloat16().float()       # a post-ReLU tensor: it is its own mask
    w5 = torch.randn(64, cin, 5, 5, device="cuda") * 0.04
    bias = torch.randn(64, device="cuda") * 0.1
    dy_out = torch.randn(B, 64, 2 * h, 2 * w, device="cuda").bfloat16().float()
    w5b = w5.bfloat16().float()
    y_ref = F.relu(F.conv2d(_up(x.double()), w5b.double(), bias.double(), padding=2))
    xb = _to_blk(x)
    dst = _Blk(B, 16, 2 * h, 2 * w, x.device)                                    # the level writes chunks 8..15
    saved = eng._up_poly_fwd(f"t{h}x{w}", xb, w5, bias, dst.view(8), B)
    y = _from_blk(dst, 64, cb_off=8)
    assert rel_err(y[:, :, 4:-4, 4:-4], y_ref[:, :, 4:-4, 4:-4]) < 1e-2          # interior: phase weights rounded to bf16
    assert rel_err(y, y_ref) < 1e-2                                              # band: the standard kernel on strips
    assert _pad_is_zero(dst, B, 16, 2 * h, 2 * w)
    assert float(_from_blk(dst, 64).abs().max()) == 0                            # chunks 0..7 untouched
    # backward: dy = dL/d(pre-activation) = dy_out * (y > 0), as the producer's dgrad epilogue hands it over
    dpre = (dy_out * (y_ref > 0)).bfloat16().float()
    xd = x.double().requires_grad_(True)
    wd = w5b.double().requires_grad_(True)
    pre = F.conv2d(_up(xd), wd, None, padding=2)
    pre.backward(dpre.double())
    dx_ref, dw_ref, db_ref = xd.grad * (x > 0), wd.grad, dpre.double().sum((0, 2, 3))
    dyb = _Blk(B, 16, 2 * h, 2 * w, x.device)
    _cabi.call("cnp_blk_from_nchw_f32", dpre.data_ptr(), dpre.stride(0), B, 64, 2 * h, 2 * w, C.byref(dyb.view(8)), _S())
    dxb = _Blk(B, 16, h, w, x.device)
    gw, gb = torch.zeros_like(w5), torch.zeros_like(bias)
    eng._up_poly_bwd(f"t{h}x{w}", xb, w5, dyb.view(8), dxb, saved, gw, gb, B, mask=xb)
    assert rel_err(_from_blk(dxb, cin), dx_ref) < 1e-2
    assert rel_err(gw, dw_ref) < 1e-2
    assert rel_err(gb, db_ref) < 1e-3
    assert _pad_is_zero(dxb, B, 16, h, w)
    assert _slack_is_zero(dxb) and _slack_is_zero(dst)
    for k, blk in eng._ws.items():                      # every strip / space-to-depth workspace of the level
        if isinstance(blk, _Blk) and str(k[0]).startswith(f"t{h}x{w}."):
            assert _slack_is_zero(blk), k
            if ".dy" in str(k[0]) or ".dys2d" in str(k[0]):
                assert _pad_is_zero(blk, blk.B, blk.CB, blk.H, blk.W), k
    # the path it replaces, same inputs: upsample kernel + standard conv / dgrad / wgrad
    up = _Blk(B, 16, 2 * h, 2 * w, x.device)
    _cabi.call("cnp_blk_upsample2x_fwd", C.byref(xb.view()), 16, C.byref(up.view()), B, _S())
    std = _Blk(B, 8, 2 * h, 2 * w, x.device)
    o = _out(std.view(), bias=bias, relu=1)
    _cabi.call("cnp_conv_tc2", C.byref(up.view()), 16, _pack(w5, _cabi.KIND_K5S1, 16).data_ptr(), _cabi.KIND_K5S1, 0, 0, 64,
               C.byref(o), B, _S())
    y_std = _from_blk(std, 64)
    assert rel_err(y, y_std) < 1e-2
    band = torch.ones_like(y, dtype=torch.bool)
    band[:, :, 4:-4, 4:-4] = False
    assert rel_err(y[band], y_std[band]) < 5e-3         # the band IS the standard kernel on (nearly) the same bf16 values


def test_model_with_and_without_polyphase(monkeypatch):
    """Same model, same batch: the engine with the polyphase levels against CNP_NO_POLYPHASE=1 (Upsample + Conv kernels),
    both in bf16, judged against the engine's fp32 mode (itself 1e-5 from the oracle, tests/test_gpu_parity.py): the
    predictions and the loss agree to bf16 rounding, and every gradient of the polyphase path is as close to the fp32
    gradient as the standard bf16 path is (a band, transposition or fold error shows up as a multiple of that noise)."""
    from deepsensornz_b200 import concat_tasks
    from deepsensornz_b200.synthetic import make_static, make_task
    static = make_static(seed=7, n_hi=120)
    task = concat_tasks([make_task(static, 700 + i) for i in range(2)])
    out = {}
    monkeypatch.setenv("CNP_POLYPHASE_MIN_PIXELS", "0")        # small grids: below the engine's break-even threshold
    for mode in ("poly", "std", "fp32"):
        if mode == "std":
            monkeypatch.setenv("CNP_NO_POLYPHASE", "1")
        else:
            monkeypatch.delenv("CNP_NO_POLYPHASE", raising=False)
        m = small_model("fp32" if mode == "fp32" else "bf16", ppu=100)
        pred = m(task)
        loss = m.loss_fn(task, normalise=True)
        loss.backward()
        out[mode] = (torch.as_tensor(pred["mean"]).clone(), torch.as_tensor(pred["std"]).clone(), float(loss.detach()),
                     {n: p.grad.detach().clone() for n, p in m.model.named_parameters() if p.grad is not None})
    (m1, s1, l1, g1), (m0, s0, l0, g0), (_, _, lr, gr) = out["poly"], out["std"], out["fp32"]
    assert rel_err(m1, m0) < 1e-2 and rel_err(s1, s0) < 1e-2
    assert abs(l1 - l0) / abs(l0) < 5e-3 and abs(l1 - lr) / abs(lr) < 5e-3

    def dist(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-300))

    for n in gr:
        e1, e0 = dist(g1[n], gr[n]), dist(g0[n], gr[n])
        assert e1 < 1.5 * e0 + 2e-3, (n, e1, e0)


def test_wide_dgrad_epilogue_writes_the_band_zeroed_space_to_depth_copy():
    """The 128-channel input gradient (WIDE) whose chunks 8..15 are dY of the next polyphase level can write their
    space-to-depth copy itself (cnp_conv_out.s2d / s2d_c0 = 8 / s2d_band = 2): same bytes as cnp_up_dy_split."""
    torch.manual_seed(8)
    B, h, w = 2, 24, 40                       # high-res size of the gradient; the next level is 12 x 20
    dy = torch.randn(B, 64, h, w, device="cuda").bfloat16().float()
    wt = (torch.randn(64, 128, 5, 5, device="cuda") * 0.05)
    act = torch.randn(B, 128, h, w, device="cuda").bfloat16().float()
    dyb, actb = _to_blk(dy), _to_blk(act)
    dst = _Blk(B, 16, h, w, dy.device)
    s2d = _Blk(B, 32, h // 2, w // 2, dy.device)
    o = _out(dst.view(0), mask=actb.view(0))
    sview = s2d.view()
    o.s2d, o.s2d_c0, o.s2d_band = C.pointer(sview), 8, 2
    _cabi.call("cnp_conv_tc2", C.byref(dyb.view()), 8, _pack(wt, _cabi.KIND_K5S1_DGRAD, 8, n_out=128).data_ptr(),
               _cabi.KIND_K5S1_DGRAD, 0, 0, 128, C.byref(o), B, _S())
    want = _Blk(B, 32, h // 2, w // 2, dy.device)
    rows, cols = _Blk(2 * B, 8, 6, w, dy.device), _Blk(2 * B, 8, 6, h, dy.device)
    _cabi.call("cnp_up_dy_split", C.byref(dst.view(8)), C.byref(want.view()), C.byref(rows.view()), C.byref(cols.view()),
               B, _S())
    assert torch.equal(s2d.t, want.t)
    assert float(_from_blk(s2d, 256).abs().max()) > 0
    # the plain output is what it was without the second output
    dst2 = _Blk(B, 16, h, w, dy.device)
    o2 = _out(dst2.view(0), mask=actb.view(0))
    _cabi.call("cnp_conv_tc2", C.byref(dyb.view()), 8, _pack(wt, _cabi.KIND_K5S1_DGRAD, 8, n_out=128).data_ptr(),
               _cabi.KIND_K5S1_DGRAD, 0, 0, 128, C.byref(o2), B, _S())
    assert torch.equal(dst.t, dst2.t)
    # strips only (s2d = NULL): the same strips
    rows2, cols2 = _Blk(2 * B, 8, 6, w, dy.device), _Blk(2 * B, 8, 6, h, dy.device)
    _cabi.call("cnp_up_dy_split", C.byref(dst.view(8)), None, C.byref(rows2.view()), C.byref(cols2.view()), B, _S())
    assert torch.equal(rows.t, rows2.t) and torch.equal(cols.t, cols2.t)


@pytest.mark.parametrize("kind,n_out,cin", [("fwd", 64, 128), ("dgrad", 128, 64)])
def test_conv_with_two_weight_tensors_equals_two_launches(kind, n_out, cin):
    """cnp_conv_tc2_w2: images b >= w2_from_b use the second packed tensor (row strips + tap-transposed column strips of a
    square level in one launch) -- bit-identical to two launches."""
    torch.manual_seed(9)
    B, h, w = 8, 6, 88
    x = torch.randn(B, cin, h, w, device="cuda").bfloat16().float()
    wt = torch.randn(64, 128, 5, 5, device="cuda") * 0.05
    wtt = wt.transpose(2, 3).contiguous()
    K = _cabi.KIND_K5S1 if kind == "fwd" else _cabi.KIND_K5S1_DGRAD
    w1, w2 = _pack(wt, K, cin // 8, n_out=n_out), _pack(wtt, K, cin // 8, n_out=n_out)
    xb = _to_blk(x)
    one = _Blk(B, n_out // 8, h, w, x.device)
    _cabi.call("cnp_conv_tc2_w2", C.byref(xb.view()), cin // 8, w1.data_ptr(), w2.data_ptr(), B // 2, K, 0, 0, n_out,
               C.byref(_out(one.view())), B, _S())
    two = _Blk(B, n_out // 8, h, w, x.device)
    _cabi.call("cnp_conv_tc2", C.byref(xb.view(0, 0)), cin // 8, w1.data_ptr(), K, 0, 0, n_out, C.byref(_out(two.view(0, 0))),
               B // 2, _S())
    _cabi.call("cnp_conv_tc2", C.byref(xb.view(0, B // 2)), cin // 8, w2.data_ptr(), K, 0, 0, n_out,
               C.byref(_out(two.view(0, B // 2))), B // 2, _S())
    assert torch.equal(one.t, two.t)
    ref = F.conv2d(x[B // 2:].double(), wtt.bfloat16().double(), padding=2) if kind == "fwd" else \
        F.conv_transpose2d(x[B // 2:].double(), wtt.bfloat16().double(), padding=2)
    assert rel_err(_from_blk(one, n_out)[B // 2:], ref) < 1e-2
