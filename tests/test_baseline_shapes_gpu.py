"""-m gpu: parity at the BASELINE shapes (VERDICT r01 "next" #1).

Model level: the CUDA path against the committed oracle vectors (tests/golden/s*.npz, made by
tests/golden/make_golden_baseline.py) at internal_density 250 (304 x 304 internal grid), 1400 x 1400 land-mask
context, 1400 x 1400 on-grid targets and the 8-channel base grid (Cin = 20), with He-scaled weights so that the UNet
matters (zeroing one up-path layer moves the loss by > 1 %, asserted from the golden file).  Tolerances are
north_star's: fp32 path 1e-5 relative on mean / std / NLL, bf16 UNet 2e-2; gradients are held per parameter
(norm, 8 fixed projections, three tensors in full).

Kernel level: the tcgen05 kernels at exactly the launches of the benchmark step (B = 16, 304 x 304): N = 160 / nacc = 3
tiles with tail splitting, masked + accumulating WIDE dgrad, K-split wgrad with workspace, narrow wgrad, the stride-2
x-phase pair; against torch float64 convolutions on bf16-exact operands.

Measured errors are appended to gpurun_out/baseline_parity.jsonl so the tolerances can be audited.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from deepsensornz_b200 import _cabi, concat_tasks
from deepsensornz_b200.engine import _Blk
from tests.test_conv_tc2_gpu import _from_blk, _out, _pack, _pad_is_zero, _S, _to_blk
from tests.util import grad_probes, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GDIR = os.path.join(ROOT, "tests", "golden")
FP32_TOL, BF16_TOL = 1e-5, 2e-2


def _mg():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_baseline", os.path.join(GDIR, "make_golden_baseline.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _log(rec):
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "baseline_parity.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass


@pytest.fixture(scope="module")
def mg():
    return _mg()


def _unet_z(m, batch):
    eng = m.engine
    g = batch.grid
    enc = eng.encode(batch)
    if m.precision == "fp32":
        z, _ = eng._unet_fwd_f32(enc, batch.B, g.n1, g.n2)
    else:
        z, _ = eng._unet_fwd_bf16(enc, batch.B, g.n1, g.n2, need_z=True)
    return enc, z


@pytest.mark.parametrize("name", ["s1_single", "s2_batch4", "s4_multivar8"])
@pytest.mark.parametrize("precision", ["fp32", "bf16", "bf16+polyphase"])
def test_training_shapes_match_golden(mg, name, precision, monkeypatch):
    """"bf16+polyphase": the polyphase resize-convolution forced on for every 128-channel decoder level (the engine only
    uses it from B x 110 x 110 low-res pixels on, i.e. at the bench batch of 16; these cases have B = 1 and 4)."""
    case = precision
    if precision.endswith("+polyphase"):
        monkeypatch.setenv("CNP_POLYPHASE_MIN_PIXELS", "0")
        precision = "bf16"
    gold = np.load(os.path.join(GDIR, name + ".npz"))
    # the golden weights make the UNet matter: zeroing one up-path layer moves the oracle's loss by far more than 1 %
    assert abs(float(gold["loss_after0_zeroed"]) - float(gold["loss"])) > 0.01 * abs(float(gold["loss"]))
    tol = FP32_TOL if precision == "fp32" else BF16_TOL
    m = mg.build_model(name, precision)
    tasks = mg.build_tasks(name)
    task = concat_tasks(tasks) if len(tasks) > 1 else tasks[0]
    batch = m._to_device(task)
    g = batch.grid
    assert [g.start1, g.n1, g.start2, g.n2, g.res] == list(gold["grid"]) and (g.n1, g.n2) == (304, 304)
    rec = dict(case=name, precision=case)
    # encoder (fp32 in both modes) and the UNet output z, an intermediate the loss alone would not pin
    enc, z = _unet_z(m, batch)
    rec["enc_sum"] = rel_err(enc.double().sum(dim=(0, 2, 3)), gold["enc_sum"])
    assert rec["enc_sum"] < 1e-5
    zs = z[mg.Z_SAMPLE]
    rec["z_sample"] = rel_err(zs, gold["z_sample"])
    rec["z_abs_mean"] = abs(float(z.abs().mean()) - float(gold["z_abs_mean"])) / float(gold["z_abs_mean"])
    assert rec["z_sample"] < (2e-5 if precision == "fp32" else BF16_TOL), rec
    assert rec["z_abs_mean"] < (1e-5 if precision == "fp32" else 5e-3), rec
    # mean / std / NLL
    pred = m(batch)
    rec["mean"] = rel_err(pred["mean"], gold["mean"])
    rec["std"] = rel_err(pred["std"], np.sqrt(gold["var"]))
    loss = m.loss_fn(batch, normalise=True)
    rec["loss"] = abs(float(loss) - float(gold["loss"])) / abs(float(gold["loss"]))
    loss.backward()
    # gradients, per parameter
    names = [str(n) for n in gold["grad_names"]]
    P = dict(m.model.named_parameters())
    worst = dict(norm=0.0, probe=0.0, full=0.0)
    detail = {}
    for i, n in enumerate(names):
        gr = P[n].grad
        assert gr is not None, n
        gn = float(gold["grad_norms"][i])
        e_norm = abs(float(gr.double().norm()) - gn) / gn
        pr = (grad_probes(n, gr.numel()).to(gr.device) @ gr.double().flatten()).cpu().numpy()
        e_probe = float(np.abs(pr - gold["grad_probes"][i]).max()) / gn
        detail[n] = (e_norm, e_probe)
        worst["norm"], worst["probe"] = max(worst["norm"], e_norm), max(worst["probe"], e_probe)
    for n in mg.FULL_GRADS:
        ref = torch.from_numpy(gold["grad." + n]).double()
        a = P[n].grad.detach().cpu().double()
        e_full = float((a - ref).norm() / ref.norm())
        cos = float((a.flatten() @ ref.flatten()) / (a.norm() * ref.norm()))
        worst["full"] = max(worst["full"], e_full)
        detail[n + ":full"] = (e_full, cos)
    rec.update(grad_norm=worst["norm"], grad_probe=worst["probe"], grad_full=worst["full"])
    _log(dict(rec, detail={k: v for k, v in detail.items() if max(v[0], v[1] if ":full" not in k else 0) > 0.2 * tol}))
    assert rec["mean"] < tol and rec["std"] < tol and rec["loss"] < tol, rec
    if precision == "fp32":
        assert worst["norm"] < 1e-4 and worst["probe"] < 1e-3 and worst["full"] < 1e-4, (rec, detail)
    else:
        # bf16 activations and weights in the UNet: every parameter's gradient within a few percent in norm and
        # direction (a projection error of e * |g| bounds the angle by about asin(e))
        assert worst["norm"] < 5e-2 and worst["probe"] < 0.15 and worst["full"] < 5e-2, (rec, detail)


@pytest.mark.parametrize("precision", ["fp32", "bf16", "bf16+polyphase"])
def test_highres_inference_matches_golden(mg, precision, monkeypatch):
    """S3: one task decoded onto the 1400 x 1400 on-grid targets (validate_ERA.py:88-92 shape)."""
    case = precision
    if precision.endswith("+polyphase"):
        monkeypatch.setenv("CNP_POLYPHASE_MIN_PIXELS", "0")
        precision = "bf16"
    gold = np.load(os.path.join(GDIR, "s3_grid1400.npz"))
    tol = FP32_TOL if precision == "fp32" else BF16_TOL
    m = mg.build_model("s3_grid1400", precision)
    task = mg.build_tasks("s3_grid1400")[0]
    batch = m._to_device(task)
    g = batch.grid
    assert [g.start1, g.n1, g.start2, g.n2, g.res] == list(gold["grid"])
    pred = m(batch)
    mean, std = pred["mean"][0, 0], pred["std"][0, 0]
    assert tuple(mean.shape) == (1400, 1400)
    s = mg.GRID_SAMPLE
    scale_m, scale_s = float(gold["mean_absmax"]), float(gold["std_max"])
    rec = dict(case="s3_grid1400", precision=case,
               mean=float((mean[::s, ::s].cpu() - torch.from_numpy(gold["mean_sample"])).abs().max()) / scale_m,
               std=float((std[::s, ::s].cpu() - torch.from_numpy(gold["std_sample"])).abs().max()) / scale_s,
               mean_rows=rel_err(mean.double().sum(dim=1), gold["mean_rowsum"]),
               mean_cols=rel_err(mean.double().sum(dim=0), gold["mean_colsum"]),
               std_rows=rel_err(std.double().sum(dim=1), gold["std_rowsum"]),
               std_cols=rel_err(std.double().sum(dim=0), gold["std_colsum"]))
    _log(rec)
    assert rec["mean"] < tol and rec["std"] < tol, rec
    # every one of the 1.96 M outputs enters a row and a column sum
    stol = 2e-5 if precision == "fp32" else 5e-3
    assert max(rec["mean_rows"], rec["mean_cols"], rec["std_rows"], rec["std_cols"]) < stol, rec


# ---------------------------------------------------------------------------------------------------------------------
# kernels at the launches of the benchmark step: B = 16, 304 x 304 (level 0) and 152 x 152 (level 1)
# ---------------------------------------------------------------------------------------------------------------------
B16, G = 16, 304


def _rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).bfloat16().float()


def _ref_conv(x, w, b=None, **kw):
    """float64 reference, batch by batch (B = 16 x 128 x 304^2 doubles at once is 1.5 GB per tensor)."""
    return torch.cat([F.conv2d(x[i:i + 1].double(), w.double(), None if b is None else b.double(), **kw)
                      for i in range(x.shape[0])])


def test_bench_launch_fwd_128_to_64_at_304():
    x, wt, b = _rnd(B16, 128, G, G, seed=1), _rnd(64, 128, 5, 5, scale=0.05, seed=2), _rnd(64, seed=3)
    ref = torch.relu(_ref_conv(x, wt, b, padding=2))
    xb, yb = _to_blk(x), _Blk(B16, 8, G, G, x.device)
    _cabi.call("cnp_conv_tc2", C.byref(xb.view()), 16, _pack(wt, _cabi.KIND_K5S1, 16).data_ptr(), _cabi.KIND_K5S1, 0, 0,
               64, C.byref(_out(yb.view(), bias=b, relu=1)), B16, _S())
    y = _from_blk(yb, 64)
    assert rel_err(y, ref) < 4e-3 and _pad_is_zero(yb, B16, 8, G, G)
    # every batch entry and both row tiles are held to the bound separately (a tile-tail bug hides in a global max)
    for i in (0, 7, 15):
        for cols in (slice(0, 160), slice(160, 304)):
            assert rel_err(y[i, :, :, cols], ref[i, :, :, cols]) < 4e-3, (i, cols)


def test_bench_launch_fwd_64_with_fused_space_to_depth_and_stride2():
    """Level-0 forward writes its space-to-depth copy in the epilogue; the stride-2 layer reads it (304 -> 152)."""
    x, w0, w1, b = _rnd(B16, 64, G, G, seed=4), _rnd(64, 64, 5, 5, scale=0.05, seed=5), \
        _rnd(64, 64, 5, 5, scale=0.05, seed=6), _rnd(64, seed=7)
    xb, yb = _to_blk(x), _Blk(B16, 8, G, G, x.device)
    ph = _Blk(B16, 32, G // 2, G // 2, x.device)
    o = _out(yb.view(), bias=b, relu=1)
    o._s2d_view = ph.view()
    o.s2d = C.pointer(o._s2d_view)
    _cabi.call("cnp_conv_tc2", C.byref(xb.view()), 8, _pack(w0, _cabi.KIND_K5S1, 8).data_ptr(), _cabi.KIND_K5S1, 0, 0, 64,
               C.byref(o), B16, _S())
    y = _from_blk(yb, 64)
    assert rel_err(y, torch.relu(_ref_conv(x, w0, b, padding=2))) < 4e-3
    y2b = _Blk(B16, 8, G // 2, G // 2, x.device)
    _cabi.call("cnp_conv_tc2", C.byref(ph.view()), 32, _pack(w1, _cabi.KIND_K5S2, 32).data_ptr(), _cabi.KIND_K5S2, 0, 0, 64,
               C.byref(_out(y2b.view(), bias=b, relu=1)), B16, _S())
    ref2 = torch.relu(_ref_conv(y, w1, b, stride=2, padding=2))      # y = the bf16 values the kernel read
    assert rel_err(_from_blk(y2b, 64), ref2) < 4e-3 and _pad_is_zero(y2b, B16, 8, G // 2, G // 2)


def test_bench_launch_wide_dgrad_masked_accumulating_at_304():
    dy, wt = _rnd(B16, 64, G, G, seed=8), _rnd(64, 128, 5, 5, scale=0.05, seed=9)
    act, prev = _rnd(B16, 128, G, G, seed=10), _rnd(B16, 128, G, G, seed=11)
    ref = torch.cat([prev[i:i + 1].double() + F.conv_transpose2d(dy[i:i + 1].double(), wt.double(), padding=2) *
                     (act[i:i + 1] > 0) for i in range(B16)])
    dyb, actb, dxb = _to_blk(dy), _to_blk(act), _to_blk(prev)
    o = _out(dxb.view(0), mask=actb.view(0), accumulate=1)
    _cabi.call("cnp_conv_tc2", C.byref(dyb.view()), 8, _pack(wt, _cabi.KIND_K5S1_DGRAD, 8, n_out=128).data_ptr(),
               _cabi.KIND_K5S1_DGRAD, 0, 0, 128, C.byref(o), B16, _S())
    got = _from_blk(dxb, 128)
    # the accumulating epilogue rounds twice (gradient to bf16, then the bf16 sum): 2^-8 of the largest value
    assert rel_err(got, ref) < 8e-3 and _pad_is_zero(dxb, B16, 16, G, G)
    assert rel_err(got[15, 64:, 300:], ref[15, 64:, 300:]) < 8e-3


@pytest.mark.parametrize("py", [0, 1])
def test_bench_launch_stride2_dgrad_x_phase_pair_at_304(py):
    dy, wt = _rnd(B16, 64, G // 2, G // 2, seed=12), _rnd(64, 64, 5, 5, scale=0.05, seed=13)
    act, prev = _rnd(B16, 64, G, G, seed=14), _rnd(B16, 64, G, G, seed=15)
    full = torch.cat([F.conv_transpose2d(dy[i:i + 1].double(), wt.double(), stride=2, padding=2, output_padding=1)
                      for i in range(B16)])
    ref = prev.double().clone()
    ref[:, :, py::2] += (full * (act > 0))[:, :, py::2]
    dyb, actb, dxb = _to_blk(dy), _to_blk(act), _to_blk(prev)
    o = _out(dxb.view(0), scatter=(2, py, 2, 0), accumulate=1, mask=actb.view(0))
    _cabi.call("cnp_conv_tc2", C.byref(dyb.view()), 8, _pack(wt, _cabi.KIND_K5S2_DGRAD, 8, py, 2).data_ptr(),
               _cabi.KIND_K5S2_DGRAD, py, 2, 64, C.byref(o), B16, _S())
    assert rel_err(_from_blk(dxb, 64), ref) < 8e-3 and _pad_is_zero(dxb, B16, 8, G, G)


@pytest.mark.parametrize("cin", [64, 128])
def test_bench_launch_wgrad_with_workspace_at_304(cin):
    x, dy = _rnd(B16, cin, G, G, seed=16), _rnd(B16, 64, G, G, seed=17)
    ref = torch.zeros(64, cin, 5, 5, device="cuda", dtype=torch.double)
    for i in range(B16):
        wd = torch.zeros(64, cin, 5, 5, device="cuda", dtype=torch.double, requires_grad=True)
        F.conv2d(x[i:i + 1].double(), wd, None, padding=2).backward(dy[i:i + 1].double())
        ref += wd.grad
    ws_bytes = _cabi.lib().cnp_conv_tc_wgrad_workspace_bytes()
    ws = torch.empty(ws_bytes // 4, device="cuda")
    xb, dyb = _to_blk(x), _to_blk(dy)
    runs = []
    for _ in range(2):
        dw, db = torch.zeros(64, cin, 5, 5, device="cuda"), torch.zeros(64, device="cuda")
        _cabi.call("cnp_conv_tc_wgrad", C.byref(xb.view()), cin // 8, C.byref(dyb.view()), _cabi.WG_K5S1, dw.data_ptr(),
                   db.data_ptr(), cin, B16, ws.data_ptr(), ws_bytes, _S())
        runs.append((dw, db))
    assert rel_err(runs[0][0], ref) < 1e-4                          # exact bf16 products, fp32 accumulation
    assert rel_err(runs[0][1], dy.double().sum(dim=(0, 2, 3))) < 1e-5
    assert torch.equal(runs[0][0], runs[1][0])                      # workspace + ordered reduce: run-to-run identical


def test_bench_launch_narrow_wgrad_at_304():
    """The folded first layer's weight gradient: 16 input channels (15 encoder channels + the constant 1)."""
    cin, cp = 15, 16
    x, dy = _rnd(B16, cin, G, G, seed=18), _rnd(B16, 64, G, G, seed=19)
    xa = _Blk(B16, 2, G, G, x.device)
    _cabi.call("cnp_blk_from_nchw_f32_ones", x.data_ptr(), x.stride(0), B16, cin, G, G, C.byref(xa.view()), 2, 0, _S())
    dyb = _to_blk(dy)
    ws_bytes = _cabi.lib().cnp_conv_tc_wgrad_workspace_bytes()
    ws = torch.empty(ws_bytes // 4, device="cuda")
    dwf, dbf = torch.zeros(64, cp, 5, 5, device="cuda"), torch.zeros(64, device="cuda")
    _cabi.call("cnp_conv_tc_wgrad", C.byref(xa.view()), 2, C.byref(dyb.view()), _cabi.WG_K5S1_NARROW, dwf.data_ptr(),
               dbf.data_ptr(), cp, B16, ws.data_ptr(), ws_bytes, _S())
    ref = torch.zeros(64, cp, 5, 5, device="cuda", dtype=torch.double)
    for i in range(B16):
        xa64 = torch.cat([x[i:i + 1].double(), torch.ones(1, 1, G, G, device="cuda", dtype=torch.double)], dim=1)
        wd = torch.zeros(64, cp, 5, 5, device="cuda", dtype=torch.double, requires_grad=True)
        F.conv2d(xa64, wd, None, padding=2).backward(dy[i:i + 1].double())
        ref += wd.grad
    assert rel_err(dwf, ref) < 1e-4
    assert rel_err(dbf, dy.double().sum(dim=(0, 2, 3))) < 1e-5
