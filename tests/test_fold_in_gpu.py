"""-m gpu: the initial 1x1 convolution folded into the first 5x5 (fold_in.cu, conv_tc2 with 2/4 source chunks,
the narrow tcgen05 wgrad) against torch float64 of the unfolded pair ``conv5x5(conv1x1(x))`` -- the head of
upstream neuralprocesses' UNet (coders/nn.py), reached from ConvNP.loss_fn (train.py:370)."""
import ctypes as C
import os

import pytest
import torch
import torch.nn.functional as F

from deepsensornz_b200 import _cabi
from deepsensornz_b200.engine import _Blk
from deepsensornz_b200.synthetic import make_static, make_task
from tests.test_conv_tc2_gpu import _from_blk, _out, _pack, _pad_is_zero, _S
from tests.util import rel_err, small_model

pytestmark = pytest.mark.gpu


def _aug(x, n_chunks):
    B, Cc, H, W = x.shape
    blk = _Blk(B, n_chunks, H, W, x.device)
    _cabi.call("cnp_blk_from_nchw_f32_ones", x.data_ptr(), x.stride(0), B, Cc, H, W, C.byref(blk.view()), n_chunks, 0, _S())
    return blk


def test_conversion_broadcasts_shared_channels():
    """Channels flagged in shared_mask are read from batch 0 for every task (a context set encoded once per batch)."""
    torch.manual_seed(4)
    B, Cc, H, W = 3, 13, 9, 11
    x = torch.randn(B, Cc, H, W, device="cuda").bfloat16().float()
    mask = (0b1111 << 2) | (1 << 12)                     # channels 2..5 and 12 are shared
    blk = _Blk(B, 2, H, W, x.device)
    _cabi.call("cnp_blk_from_nchw_f32_ones", x.data_ptr(), x.stride(0), B, Cc, H, W, C.byref(blk.view()), 2, mask, _S())
    got = _from_blk(blk, 16)
    want = x.clone()
    for c in (2, 3, 4, 5, 12):
        want[:, c] = x[0, c]
    assert torch.equal(got[:, :Cc], want) and float((got[:, Cc] - 1).abs().max()) == 0 and float(got[:, Cc + 1:].abs().max()) == 0


def _params(cin, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    w5 = torch.randn(64, 64, 5, 5, device="cuda", generator=g) * 0.05
    w1 = torch.randn(64, cin, 1, 1, device="cuda", generator=g) * 0.3
    b1 = torch.randn(64, device="cuda", generator=g) * 0.2
    b5 = torch.randn(64, device="cuda", generator=g) * 0.1
    return w5, w1, b1, b5


@pytest.mark.parametrize("cin,n_chunks,h,w", [(7, 2, 40, 52), (8, 2, 33, 170), (15, 2, 21, 30), (20, 4, 38, 44)])
def test_folded_forward_matches_conv1x1_then_conv5x5(cin, n_chunks, h, w):
    """One 5x5 convolution over [x ; 1] with the folded weights == conv5x5(pad0(conv1x1(x) + b1)) + b5, borders
    included (the constant channel is zero in the pad exactly like the 1x1's output)."""
    B = 2
    w5, w1, b1, b5 = _params(cin, 3)
    x = torch.randn(B, cin, h, w, device="cuda").bfloat16().float()
    cp = n_chunks * 8
    wf = torch.empty(64, cp, 5, 5, device="cuda")
    _cabi.call("cnp_fold_in_fwd", w5.data_ptr(), w1.data_ptr(), b1.data_ptr(), 64, 64, cin, cp, 5, wf.data_ptr(), _S())
    # the fold itself (fp32 contraction) against float64
    wf_ref = torch.einsum("omyx,mc->ocyx", w5.double(), w1.double()[:, :, 0, 0])
    assert rel_err(wf[:, :cin], wf_ref) < 1e-6
    assert rel_err(wf[:, cin], torch.einsum("omyx,m->oyx", w5.double(), b1.double())) < 1e-6
    assert float(wf[:, cin + 1:].abs().max()) == 0 if cp > cin + 1 else True
    xa = _aug(x, n_chunks)
    assert _pad_is_zero(xa, B, n_chunks, h, w)
    full = _from_blk(xa, cp)
    assert torch.equal(full[:, :cin], x) and float((full[:, cin] - 1).abs().max()) == 0
    y = _Blk(B, 8, h, w, x.device)
    _cabi.call("cnp_conv_tc2", C.byref(xa.view()), n_chunks, _pack(wf, _cabi.KIND_K5S1, n_chunks).data_ptr(),
               _cabi.KIND_K5S1, 0, 0, 64, C.byref(_out(y.view(0), bias=b5, relu=1)), B, _S())
    hd = F.conv2d(x.double(), w1.double(), b1.double())
    ref = torch.relu(F.conv2d(hd, w5.double(), b5.double(), padding=2))
    # folded weights are rounded to bf16 once (the unfolded pair rounds W5 and the 64-channel h instead)
    assert rel_err(_from_blk(y, 64), ref) < 1e-2
    assert _pad_is_zero(y, B, 8, h, w)
    # exactness of the kernel itself: same bf16-rounded folded weights in float64
    wfb = wf.bfloat16().double()
    xa64 = torch.cat([x.double(), torch.ones(B, 1, h, w, device="cuda", dtype=torch.double),
                      torch.zeros(B, cp - cin - 1, h, w, device="cuda", dtype=torch.double)], dim=1)
    ref_b = torch.relu(F.conv2d(xa64, wfb, b5.double(), padding=2))
    assert rel_err(_from_blk(y, 64), ref_b) < 4e-3      # bf16 rounding of the stored output only


@pytest.mark.parametrize("use_ws", [False, True])
@pytest.mark.parametrize("cin,n_chunks,h,w", [(7, 2, 40, 52), (8, 2, 76, 76), (20, 4, 38, 44), (3, 1, 20, 24),
                                              (40, 6, 24, 36)])
def test_narrow_wgrad_and_chain_rule(cin, n_chunks, h, w, use_ws):
    """dWf from the narrow tensor-core wgrad == torch's weight gradient of the folded convolution; fold_in_bwd maps
    it to exactly the gradients autograd gives for W5, W1, b1 of the unfolded pair."""
    if cin + 1 > n_chunks * 8:
        pytest.skip("needs room for the constant channel")
    B = 3
    w5, w1, b1, b5 = _params(cin, 5)
    x = torch.randn(B, cin, h, w, device="cuda").bfloat16().float()
    dy = torch.randn(B, 64, h, w, device="cuda").bfloat16().float()
    cp = n_chunks * 8
    ws_bytes = _cabi.lib().cnp_conv_tc_wgrad_workspace_bytes() if use_ws else 0
    ws = torch.empty(max(ws_bytes // 4, 1), device="cuda")
    xa = _aug(x, n_chunks)
    dyb = _Blk(B, 8, h, w, x.device)
    _cabi.call("cnp_blk_from_nchw_f32", dy.data_ptr(), dy.stride(0), B, 64, h, w, C.byref(dyb.view()), _S())
    dwf = torch.zeros(64, cp, 5, 5, device="cuda")
    dbf = torch.zeros(64, device="cuda")
    _cabi.call("cnp_conv_tc_wgrad", C.byref(xa.view()), n_chunks, C.byref(dyb.view()), _cabi.WG_K5S1_NARROW,
               dwf.data_ptr(), dbf.data_ptr(), cp, B, ws.data_ptr() if use_ws else None, ws_bytes, _S())
    xa64 = torch.cat([x.double(), torch.ones(B, 1, h, w, device="cuda", dtype=torch.double),
                      torch.zeros(B, cp - cin - 1, h, w, device="cuda", dtype=torch.double)], dim=1)
    wd = torch.zeros(64, cp, 5, 5, device="cuda", dtype=torch.double, requires_grad=True)
    F.conv2d(xa64, wd, None, padding=2).backward(dy.double())
    assert rel_err(dwf, wd.grad) < 1e-4
    assert rel_err(dbf, dy.double().sum(dim=(0, 2, 3))) < 1e-5
    # chain rule to the real parameters vs autograd through the unfolded pair (float64, same x / dy)
    W5, W1, B1 = (t.double().requires_grad_(True) for t in (w5, w1, b1))
    F.conv2d(F.conv2d(x.double(), W1, B1), W5, None, padding=2).backward(dy.double())
    dw5 = torch.zeros_like(w5)
    dw1 = torch.zeros_like(w1)
    db1 = torch.zeros_like(b1)
    _cabi.call("cnp_fold_in_bwd", dwf.data_ptr(), w5.data_ptr(), w1.data_ptr(), b1.data_ptr(), 64, 64, cin, cp, 5,
               dw5.data_ptr(), dw1.data_ptr(), db1.data_ptr(), _S())
    assert rel_err(dw5, W5.grad) < 1e-4
    assert rel_err(dw1, W1.grad) < 1e-4
    assert rel_err(db1, B1.grad) < 1e-4


def test_folded_model_matches_unfolded_model(monkeypatch):
    """Same model, same task: the folded first layer (default) and the unfolded 1x1 -> 5x5 path (CNP_NO_FOLD_IN=1) agree
    on the loss and on every parameter gradient to bf16 accuracy, and the folded path really ran."""
    static = make_static(seed=7, n_hi=200, with_aux_hi=True)
    task = make_task(static, 31)
    m = small_model("bf16")
    calls = []
    orig = m.engine._call
    monkeypatch.setattr(m.engine, "_call", lambda name, *a, **k: (calls.append(name), orig(name, *a, **k))[1])
    loss = m.loss_fn(task, normalise=True)
    loss.backward()
    assert "cnp_fold_in_fwd" in calls and "cnp_fold_in_bwd" in calls and "cnp_conv1x1_in_bf16" not in calls
    g_fold = {n: p.grad.detach().clone() for n, p in m.model.named_parameters() if p.grad is not None}
    for p in m.model.parameters():
        p.grad = None
    monkeypatch.setenv("CNP_NO_FOLD_IN", "1")
    calls.clear()
    loss2 = m.loss_fn(task, normalise=True)
    loss2.backward()
    assert "cnp_conv1x1_in_bf16" in calls and "cnp_fold_in_bwd" not in calls
    assert abs(float(loss) - float(loss2)) / abs(float(loss2)) < 2e-2
    for n, p in m.model.named_parameters():
        if p.grad is not None:
            a, b = g_fold[n].double().flatten(), p.grad.double().flatten()
            cos = float((a @ b) / (a.norm() * b.norm()).clamp(min=1e-300))
            assert cos > 0.99, (n, cos)
