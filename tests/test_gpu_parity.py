"""-m gpu: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Tolerances are the ones BASELINE.json's north_star states: fp32 path 1e-5 relative on mean/std/NLL,
bf16 UNet 2e-2 relative.  Per-kernel checks use the same fp32 bound against torch fp32 ops.
"""
import ctypes as C
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from deepsensornz_b200 import _cabi, concat_tasks, Task
from deepsensornz_b200.engine import _Blk
from deepsensornz_b200.synthetic import make_static, make_task
from oracle import convnp_oracle as O
from tests.util import cpu_params, oracle_inputs, rel_err, small_model

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-5
BF16_TOL = 2e-2


@pytest.fixture(scope="module")
def static():
    return make_static(seed=7, n_hi=200)


def _S():
    return torch.cuda.current_stream().cuda_stream


def test_library_loads_on_device():
    _cabi.check_device()
    assert _cabi.lib().cnp_version() >= 100


# ---------------------------------------------------------------------------------------------
# per-kernel parity (fp32)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,stride,cin,cout,h,w", [(5, 1, 64, 64, 40, 48), (5, 2, 64, 64, 40, 48), (1, 1, 15, 64, 33, 20),
                                                   (5, 1, 128, 64, 24, 24), (1, 1, 64, 64, 16, 72)])
def test_conv_f32_fwd_bwd(k, stride, cin, cout, h, w):
    torch.manual_seed(1)
    B = 2
    x = torch.randn(B, cin, h, w, device="cuda")
    wt = torch.randn(cout, cin, k, k, device="cuda") * 0.05
    b = torch.randn(cout, device="cuda")
    ref = F.conv2d(x.double(), wt.double(), b.double(), stride=stride, padding=k // 2)
    y = torch.empty(ref.shape, device="cuda")
    _cabi.call("cnp_conv2d_fwd_f32", x.data_ptr(), x.stride(0), wt.data_ptr(), b.data_ptr(), y.data_ptr(), y.stride(0),
               B, cin, h, w, cout, k, stride, 0, _S())
    assert rel_err(y, ref) < FP32_TOL
    dy = torch.randn_like(y)
    xd = x.double().requires_grad_(True)
    wd = wt.double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    F.conv2d(xd, wd, bd, stride=stride, padding=k // 2).backward(dy.double())
    dx = torch.empty_like(x)
    _cabi.call("cnp_conv2d_dgrad_f32", dy.data_ptr(), dy.stride(0), wt.data_ptr(), dx.data_ptr(), dx.stride(0), B, cin,
               h, w, cout, k, stride, 0, _S())
    assert rel_err(dx, xd.grad) < FP32_TOL
    dw = torch.zeros_like(wt)
    db = torch.zeros_like(b)
    _cabi.call("cnp_conv2d_wgrad_f32", x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0), dw.data_ptr(),
               db.data_ptr(), B, cin, h, w, cout, k, stride, _S())
    assert rel_err(dw, wd.grad) < 2e-5
    assert rel_err(db, bd.grad) < 2e-5


def test_upsample_f32():
    torch.manual_seed(2)
    x = torch.randn(2, 5, 7, 9, device="cuda")
    xd = x.double().requires_grad_(True)
    ref = F.interpolate(xd, scale_factor=2, mode="bilinear", align_corners=False)
    y = torch.empty(2, 5, 14, 18, device="cuda")
    _cabi.call("cnp_upsample2x_fwd_f32", x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), 2, 5, 7, 9, _S())
    assert rel_err(y, ref) < FP32_TOL
    dy = torch.randn_like(y)
    ref.backward(dy.double())
    dx = torch.empty_like(x)
    _cabi.call("cnp_upsample2x_bwd_f32", dy.data_ptr(), dy.stride(0), dx.data_ptr(), dx.stride(0), 2, 5, 7, 9, 0, _S())
    assert rel_err(dx, xd.grad) < FP32_TOL


def test_encoder_matches_oracle(static):
    m = small_model("fp32")
    tasks = [make_task(static, 100 + i) for i in range(2)]
    task = concat_tasks(tasks)
    batch = m._to_device(task)
    enc = m.engine.encode(batch)
    contexts, xt, yt, aux = oracle_inputs(task)
    P = cpu_params(m)
    (s1, n1), (s2, n2), res = O.discretise([c[0] for c in contexts] + [xt], m.config.points_per_unit, 0.1, 8)
    assert (n1, n2) == (batch.grid.n1, batch.grid.n2)
    ref = O.encoder(P, contexts, O.grid_points(s1, n1, res), O.grid_points(s2, n2, res))
    assert enc.shape == ref.shape
    assert rel_err(enc, ref) < FP32_TOL
    # per-channel check so that small channels are held to the same bound
    for c in range(ref.shape[1]):
        assert rel_err(enc[:, c], ref[:, c]) < 5e-5, c


def test_raw_task_equals_masked_task(static):
    """Device-side NaN->mask (raw upload) == host-side Masked path, bit for bit."""
    m = small_model("fp32")
    t = make_task(static, 5)
    e1 = m.engine.encode(m._to_device(t)).clone()
    e2 = m.engine.encode(m._to_device(m.modify_task(t))).clone()
    assert torch.equal(e1, e2)


def test_decoder_fwd_bwd():
    torch.manual_seed(3)
    B, Cz, n1, n2, Nt = 2, 64, 56, 64, 37
    res, s1, s2 = 0.02, -0.1, -0.14
    z = torch.randn(B, Cz, n1, n2, device="cuda")
    xt = torch.rand(B, 2, Nt, device="cuda") * torch.tensor([n1 * res, n2 * res], device="cuda").view(1, 2, 1) + \
        torch.tensor([s1, s2], device="cuda").view(1, 2, 1)
    ls = torch.tensor(math.log(res))
    g1, g2 = O.grid_points(s1, n1, res), O.grid_points(s2, n2, res)
    zd = z.cpu().double().requires_grad_(True)
    ref = O.decode({"decoder.set_conv.log_scale": ls.double()}, zd, g1.double(), g2.double(), xt.cpu().double())
    f = torch.empty(B, Cz, Nt, device="cuda")
    sc2 = float(np.float32(math.exp(2 * float(ls))))
    _cabi.call("cnp_setconv_dec_offgrid_fwd", z.data_ptr(), z.stride(0), xt.data_ptr(), B, Cz, Nt, s1, n1, s2, n2, res,
               sc2, f.data_ptr(), Cz, _S())
    assert rel_err(f, ref) < FP32_TOL
    df = torch.randn_like(f)
    ref.backward(df.cpu().double())
    dz = torch.empty_like(z)
    _cabi.call("cnp_setconv_dec_offgrid_bwd", df.data_ptr(), Cz, xt.data_ptr(), B, Cz, Nt, s1, n1, s2, n2, res, sc2,
               dz.data_ptr(), dz.stride(0), _S())
    assert rel_err(dz, zd.grad) < FP32_TOL


# ---------------------------------------------------------------------------------------------
# full model, fp32 mode: mean / std / NLL and gradients vs oracle autograd
# ---------------------------------------------------------------------------------------------
def _oracle_loss_and_grads(m, task):
    contexts, xt, yt, aux = oracle_inputs(task)
    P = {k: v.clone().requires_grad_(v.dim() > 0) for k, v in cpu_params(m).items()}
    mean, var = O.forward(P, contexts, xt, aux, m.config.points_per_unit)
    loss = -O.loglik(mean, var, yt, True).mean()
    loss.backward()
    return mean.detach(), var.detach(), loss.detach(), {k: v.grad for k, v in P.items() if v.grad is not None}


@pytest.mark.parametrize("nb", [1, 3])
def test_model_fp32_matches_oracle(static, nb):
    m = small_model("fp32")
    tasks = [make_task(static, 200 + i) for i in range(nb)]
    task = concat_tasks(tasks) if nb > 1 else tasks[0]
    mean_o, var_o, loss_o, grads_o = _oracle_loss_and_grads(m, task)
    pred = m(task)
    assert rel_err(pred["mean"], mean_o) < FP32_TOL
    assert rel_err(pred["std"], var_o.sqrt()) < FP32_TOL
    loss = m.loss_fn(task, normalise=True)
    assert abs(float(loss) - float(loss_o)) / abs(float(loss_o)) < FP32_TOL
    loss.backward()
    for n, p in m.model.named_parameters():
        if p.requires_grad:
            assert p.grad is not None, n
            assert rel_err(p.grad, grads_o[n]) < 1e-4, n


def test_model_bf16_matches_oracle(static):
    m = small_model("bf16")
    tasks = [make_task(static, 300 + i) for i in range(2)]
    task = concat_tasks(tasks)
    mean_o, var_o, loss_o, grads_o = _oracle_loss_and_grads(m, task)
    pred = m(task)
    assert rel_err(pred["mean"], mean_o) < BF16_TOL
    assert rel_err(pred["std"], var_o.sqrt()) < BF16_TOL
    loss = m.loss_fn(task, normalise=True)
    assert abs(float(loss) - float(loss_o)) / abs(float(loss_o)) < BF16_TOL
    loss.backward()
    for n, p in m.model.named_parameters():
        if p.requires_grad:
            g, r = p.grad.detach().cpu().double().flatten(), grads_o[n].double().flatten()
            cos = float((g @ r) / (g.norm() * r.norm()).clamp(min=1e-300))
            assert cos > 0.995, (n, cos)
            assert rel_err(p.grad, grads_o[n]) < 0.15, n


# ---------------------------------------------------------------------------------------------
# tcgen05 conv against torch on bf16-rounded operands (isolates the kernel from model effects)
# ---------------------------------------------------------------------------------------------
def _to_blk(x, cb_total=None):
    B, Cc, H, W = x.shape
    blk = _Blk(B, cb_total or Cc // 8, H, W, x.device)
    _cabi.call("cnp_blk_from_nchw_f32", x.data_ptr(), x.stride(0), B, Cc, H, W, C.byref(blk.view()), _S())
    return blk


def _from_blk(blk, Cc, cb_off=0):
    out = torch.empty(blk.B, Cc, blk.H, blk.W, device="cuda")
    _cabi.call("cnp_blk_to_nchw_f32", C.byref(blk.view(cb_off)), blk.B, Cc, out.data_ptr(), out.stride(0), _S())
    return out


def _legacy():
    return hasattr(_cabi.lib(), "cnp_conv_tc")


legacy = pytest.mark.skipif(not _legacy(), reason="first convolution formulation: only in a `make LEGACY=1` build")


@legacy
@pytest.mark.parametrize("cin,h,w", [(64, 38, 38), (128, 76, 76), (64, 152, 160), (128, 61, 45)])
def test_conv_tc_k5s1(cin, h, w):
    torch.manual_seed(4)
    B = 2
    x = torch.randn(B, cin, h, w, device="cuda").bfloat16().float()
    wt = (torch.randn(64, cin, 5, 5, device="cuda") * 0.05).bfloat16().float()
    b = torch.randn(64, device="cuda")
    ref = F.relu(F.conv2d(x.double(), wt.double(), b.double(), padding=2))
    xb = _to_blk(x)
    yb = _Blk(B, 8, h, w, x.device)
    nbytes = _cabi.lib().cnp_conv_tc_packed_bytes(_cabi.KIND_K5S1, cin // 8)
    wpk = torch.empty(nbytes // 2, dtype=torch.bfloat16, device="cuda")
    _cabi.call("cnp_conv_tc_pack", wt.data_ptr(), 64, cin, 5, _cabi.KIND_K5S1, cin // 8, 0, 0, 0, wpk.data_ptr(), _S())
    o = _cabi.CnpConvOut()
    o.mode, o.blk = 0, yb.view()
    o.sy, o.ay, o.sx, o.ax = 1, 0, 1, 0
    o.bias, o.relu = b.data_ptr(), 1
    _cabi.call("cnp_conv_tc", C.byref(xb.view()), cin // 8, wpk.data_ptr(), _cabi.KIND_K5S1, 0, 0, C.byref(o), B, _S())
    y = _from_blk(yb, 64)
    torch.cuda.synchronize()
    assert rel_err(y, ref) < 1e-2   # bf16 output rounding only (operands are exact in bf16)
    # the zero pad of the output buffer must be untouched
    full = yb.t[:B * yb.bstride].view(B, 8, h + 4, w + 4, 8).float()
    assert float(full[:, :, :2].abs().max()) == 0 and float(full[:, :, :, :2].abs().max()) == 0
    assert float(full[:, :, -2:].abs().max()) == 0 and float(full[:, :, :, -2:].abs().max()) == 0


def _pack(wt, kind, n_chunks, py=0, px=0, co_off=0):
    nbytes = _cabi.lib().cnp_conv_tc_packed_bytes(kind, n_chunks)
    wpk = torch.empty(nbytes // 2, dtype=torch.bfloat16, device="cuda")
    co, ci, k, _ = wt.shape
    _cabi.call("cnp_conv_tc_pack", wt.data_ptr(), co, ci, k, kind, n_chunks, py, px, co_off, wpk.data_ptr(), _S())
    return wpk


def _out(blk_view, bias=None, relu=0, scatter=(1, 0, 1, 0), accumulate=0):
    o = _cabi.CnpConvOut()
    o.mode, o.blk = 0, blk_view
    o.sy, o.ay, o.sx, o.ax = scatter
    o.bias, o.relu, o.accumulate = (bias.data_ptr() if bias is not None else None), relu, accumulate
    return o


@legacy
def test_conv_tc_stride2_and_1x1():
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
    _cabi.call("cnp_conv_tc", C.byref(ph.view()), 32, _pack(wt, _cabi.KIND_K5S2, 32).data_ptr(), _cabi.KIND_K5S2, 0, 0,
               C.byref(o), B, _S())
    assert rel_err(_from_blk(yb, 64), ref) < 1e-2
    # 1x1 with fp32 NCHW output
    w1 = (torch.randn(64, 64, 1, 1, device="cuda") * 0.1).bfloat16().float()
    ref1 = F.conv2d(x.double(), w1.double(), b.double())
    z = torch.empty(B, 64, h, w, device="cuda")
    o = _cabi.CnpConvOut()
    o.mode, o.f32, o.f32_bstride, o.f32_ch_off = 1, z.data_ptr(), z.stride(0), 0
    o.sy, o.ay, o.sx, o.ax = 1, 0, 1, 0
    o.bias = b.data_ptr()
    _cabi.call("cnp_conv_tc", C.byref(xb.view()), 8, _pack(w1, _cabi.KIND_K1, 8).data_ptr(), _cabi.KIND_K1, 0, 0,
               C.byref(o), B, _S())
    assert rel_err(z, ref1) < 1e-5 * 50  # fp32 accumulate of exact bf16 products


@legacy
@pytest.mark.parametrize("cin", [64, 128])
def test_conv_tc_dgrad_s1(cin):
    torch.manual_seed(6)
    B, h, w = 2, 45, 52
    dy = torch.randn(B, 64, h, w, device="cuda").bfloat16().float()
    wt = (torch.randn(64, cin, 5, 5, device="cuda") * 0.05).bfloat16().float()
    xd = torch.zeros(B, cin, h, w, device="cuda", dtype=torch.double, requires_grad=True)
    F.conv2d(xd, wt.double(), None, padding=2).backward(dy.double())
    dyb = _to_blk(dy)
    dxb = _Blk(B, cin // 8, h, w, dy.device)
    for g in range(cin // 64):
        o = _out(dxb.view(8 * g))
        _cabi.call("cnp_conv_tc", C.byref(dyb.view()), 8, _pack(wt, _cabi.KIND_K5S1_DGRAD, 8, 0, 0, 64 * g).data_ptr(),
                   _cabi.KIND_K5S1_DGRAD, 0, 0, C.byref(o), B, _S())
    assert rel_err(_from_blk(dxb, cin), xd.grad) < 1e-2


@legacy
def test_conv_tc_dgrad_s2():
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
            _cabi.call("cnp_conv_tc", C.byref(dyb.view()), 8,
                       _pack(wt, _cabi.KIND_K5S2_DGRAD, 8, py, px).data_ptr(), _cabi.KIND_K5S2_DGRAD, py, px,
                       C.byref(o), B, _S())
    assert rel_err(_from_blk(dxb, 64), xd.grad) < 1e-2


@pytest.mark.parametrize("use_ws", [False, True])
@pytest.mark.parametrize("cin,stride,k,h,w", [(128, 1, 5, 38, 44), (64, 1, 5, 76, 76), (64, 2, 5, 48, 40), (64, 1, 1, 33, 70),
                                              (64, 1, 5, 20, 24)])
def test_conv_tc_wgrad(cin, stride, k, h, w, use_ws):
    torch.manual_seed(8)
    ws_bytes = _cabi.lib().cnp_conv_tc_wgrad_workspace_bytes() if use_ws else 0
    ws = torch.empty(max(ws_bytes // 4, 1), device="cuda")
    ws_ptr = ws.data_ptr() if use_ws else None
    B = 3
    x = torch.randn(B, cin, h, w, device="cuda").bfloat16().float()
    ho, wo = h // stride, w // stride
    dy = torch.randn(B, 64, ho, wo, device="cuda").bfloat16().float()
    wd = torch.zeros(64, cin, k, k, device="cuda", dtype=torch.double, requires_grad=True)
    F.conv2d(x.double(), wd, None, stride=stride, padding=k // 2).backward(dy.double())
    xb, dyb = _to_blk(x), _to_blk(dy)
    dw = torch.zeros(64, cin, k, k, device="cuda")
    dbf = torch.zeros(64, device="cuda")
    if stride == 2:
        ph = _Blk(B, 32, ho, wo, x.device)
        _cabi.call("cnp_blk_space_to_depth", C.byref(xb.view()), 8, C.byref(ph.view()), B, _S())
        _cabi.call("cnp_conv_tc_wgrad", C.byref(ph.view()), 32, C.byref(dyb.view()), _cabi.WG_K5S2, dw.data_ptr(),
                   dbf.data_ptr(), cin, B, ws_ptr, ws_bytes, _S())
    else:
        kind = _cabi.WG_K5S1 if k == 5 else _cabi.WG_K1
        _cabi.call("cnp_conv_tc_wgrad", C.byref(xb.view()), cin // 8, C.byref(dyb.view()), kind, dw.data_ptr(),
                   dbf.data_ptr(), cin, B, ws_ptr, ws_bytes, _S())
    assert rel_err(dw, wd.grad) < 1e-4   # exact bf16 products, fp32 accumulation
    assert rel_err(dbf, dy.double().sum(dim=(0, 2, 3))) < 1e-5   # bias gradient fused into the wgrad kernel
    db = torch.zeros(64, device="cuda")
    _cabi.call("cnp_blk_channel_sum", C.byref(dyb.view()), 8, B, db.data_ptr(), _S())
    assert rel_err(db, dy.double().sum(dim=(0, 2, 3))) < 1e-5


# ---------------------------------------------------------------------------------------------
# committed golden vectors (tests/golden/*.npz, generated by tests/golden/make_golden.py from the oracle)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["g1_single", "g2_batch3", "g3_multivar"])
@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_cuda_path_matches_golden(name, precision, tol):
    import importlib.util
    import os
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(gdir, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    m, task = mg.build(name)
    gold = np.load(os.path.join(gdir, name + ".npz"))
    m2 = small_model(precision, ppu=m.config.points_per_unit, dim_yc=m.config.dim_yc, seed=1234)
    g = m2._to_device(task).grid
    assert [g.start1, g.n1, g.start2, g.n2, g.res] == list(gold["grid"])
    pred = m2(task)
    assert rel_err(pred["mean"], gold["mean"]) < tol
    assert rel_err(pred["std"], np.sqrt(gold["var"])) < tol
    loss = m2.loss_fn(task, normalise=True)
    assert abs(float(loss) - float(gold["loss"])) < tol * abs(float(gold["loss"]))
    loss.backward()
    gtol = 1e-4 if precision == "fp32" else 0.15
    assert rel_err(m2.model.decoder.unet.final_linear.bias.grad, gold["d_final_bias"]) < gtol
    assert rel_err(m2.model.decoder.mlp.layers[0].weight.grad, gold["d_mlp0"]) < gtol


def test_nan_loss_surfaces_as_nan(static):
    """train.py:371,395 filter NaN losses: a poisoned task must give NaN, not crash."""
    m = small_model("fp32")
    t = make_task(static, 77)
    t["Y_t_aux"][0, 3] = np.inf
    loss = m.loss_fn(t, normalise=True)
    assert not np.isfinite(float(loss))


def test_encoder_frozen_and_requires_grad_honoured(static):
    m = small_model("fp32")
    for p in m.model.decoder.unet.before_turn_layers[0].parameters():
        p.requires_grad = False
    loss = m.loss_fn(make_task(static, 78), normalise=True)
    loss.backward()
    assert m.model.decoder.unet.before_turn_layers[0].weight.grad is None
    assert m.model.decoder.unet.before_turn_layers[1].weight.grad is not None
    assert all(p.grad is None for p in m.model.encoder.parameters())


def test_train_epoch_decreases_loss(static):
    from deepsensornz_b200 import train_epoch
    np.random.seed(0)
    m = small_model("bf16")
    tasks = [make_task(static, 500 + i) for i in range(8)]
    opt = torch.optim.AdamW(m.model.parameters(), lr=2e-3)
    first = np.mean(train_epoch(m, tasks, batch_size=4, opt=opt))
    for _ in range(5):
        last = np.mean(train_epoch(m, tasks, batch_size=4, opt=opt))
    assert np.isfinite(first) and last < first


def test_graphed_train_step_matches_eager(static):
    """The CUDA-graph replay of (forward, NLL, backward, AdamW) follows the eager step bit for bit in its losses."""
    from deepsensornz_b200.graph import GraphedTrainStep
    tasks = [concat_tasks([make_task(static, 700 + 4 * k + i) for i in range(4)]) for k in range(3)]
    losses = {}
    for mode in ("eager", "graph"):
        m = small_model("bf16", seed=3)
        opt = torch.optim.AdamW(m.model.parameters(), lr=1e-3, fused=True, capturable=True)
        dev = [m._to_device(t) for t in tasks]
        out = []
        if mode == "graph":
            gs = GraphedTrainStep(m, opt, dev[0], warmup=1)      # = one eager step on dev[0] before the capture
            assert all(gs.matches(d) for d in dev)
            for k in range(6):
                out.append(float(gs.step(dev[k % 3])))
        else:
            for k in [0] + list(range(6)):
                opt.zero_grad(set_to_none=True)
                loss = m.loss_fn(dev[k % 3], normalise=True)
                loss.backward()
                opt.step()
                out.append(float(loss))
        losses[mode] = out
    assert losses["graph"][-1] < losses["graph"][0]          # it trains
    for a, b in zip(losses["eager"][1:], losses["graph"]):
        assert abs(a - b) <= 2e-3 * abs(a), (losses["eager"], losses["graph"])   # atomics reorder fp32 sums


def test_train_epoch_with_graph_decreases_loss(static):
    from deepsensornz_b200 import train_epoch
    import numpy as np
    np.random.seed(0)
    tasks = [make_task(static, 500 + i) for i in range(8)]
    m = small_model("bf16", seed=5)
    opt = torch.optim.AdamW(m.model.parameters(), lr=2e-3)      # not capturable: the optimiser step stays eager
    first = last = None
    for ep in range(6):
        losses = train_epoch(m, tasks, batch_size=4, opt=opt, use_graph=True)
        assert len(losses) == 2 and all(np.isfinite(losses))
        first = first if first is not None else float(np.mean(losses))
        last = float(np.mean(losses))
    assert last < first
    assert any(g not in (None, False) for g in m._train_graphs.values())   # a graph was captured and replayed


@pytest.mark.parametrize("case", ["no_context_stations", "one_target", "all_sea"])
def test_edge_cases_match_oracle(static, case):
    """Empty / degenerate inputs the domain produces: an hour with no reporting context station, a single target
    station, and a base grid that is entirely masked (all NaN)."""
    t = make_task(static, 3100, n_stations=30, context_frac=0.6)
    t = Task({k: (list(v) if isinstance(v, list) else v) for k, v in t.items()})
    if case == "no_context_stations":
        t["X_c"][3] = t["X_c"][3][:, :0]
        t["Y_c"][3] = t["Y_c"][3][:, :0]
    elif case == "one_target":
        t["X_t"][0] = t["X_t"][0][:, :1]
        t["Y_t"][0] = t["Y_t"][0][:, :1]
        t["Y_t_aux"] = t["Y_t_aux"][:, :1]
    else:
        y = t["Y_c"][0].copy()
        y[:] = np.nan
        t["Y_c"][0] = y
    m = small_model("fp32")
    loss = m.loss_fn(t, normalise=True)
    loss.backward()
    ctx, xt, yt, aux = oracle_inputs(t)
    ref = O.loss_fn(cpu_params(m), ctx, xt, yt, aux, m.config.points_per_unit)
    assert np.isfinite(float(loss))
    assert abs(float(loss) - float(ref)) / abs(float(ref)) < FP32_TOL
    assert all(torch.isfinite(p.grad).all() for p in m.model.parameters() if p.requires_grad)
