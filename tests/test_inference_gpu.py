"""-m gpu: the inference path (on-grid targets, ``ConvNP.__call__`` / ``ConvNP.predict`` as nzdownscale's
validate_ERA.py:88-92 drives it) against the oracle, and the blocked-decoder kernels (dec_blk.cu) against the
materialised ``final 1x1 conv -> SetConv`` path they replace."""
import ctypes as C

import numpy as np
import pytest
import torch

from deepsensornz_b200 import _cabi, Task
from deepsensornz_b200.engine import _Blk
from deepsensornz_b200.synthetic import make_static, make_task
from oracle import convnp_oracle as O
from tests.util import cpu_params, oracle_inputs, rel_err, small_model

pytestmark = pytest.mark.gpu


def _S():
    return torch.cuda.current_stream().cuda_stream


@pytest.fixture(scope="module")
def static():
    return make_static(seed=7, n_hi=200, with_aux_hi=True)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_on_grid_forward_matches_oracle(static, precision, tol):
    """mean / std on a gridded target (the decoder-bound path of configs[2]) vs the dense oracle einsum."""
    task = make_task(static, 4242, grid_targets=True)
    m = small_model(precision)
    pred = m(task)
    ctx, xt, _, aux = oracle_inputs(task)
    mean_o, var_o = O.forward(cpu_params(m), ctx, xt, aux, m.config.points_per_unit)
    assert pred["mean"].shape == mean_o.shape
    assert rel_err(pred["mean"], mean_o) < tol
    assert rel_err(pred["std"], var_o.sqrt()) < tol


def test_predict_matches_forward_and_handles_many_tasks(static):
    """``predict(tasks, X_t=(x1, x2))`` = per-task forward on the target grid, aux-at-targets uploaded once."""
    tasks = [make_task(static, 900 + i, all_context=True) for i in range(3)]
    m = small_model("fp32")
    x1 = np.linspace(0.05, 0.95, 57).astype(np.float32)
    x2 = np.linspace(0.10, 0.90, 43).astype(np.float32)
    aux = np.random.default_rng(3).uniform(-1, 1, (5, 57, 43)).astype(np.float32)
    pred = m.predict(tasks, X_t=(x1, x2), X_t_is_normalised=True, aux_at_targets_override=aux)
    key = list(pred.keys())[0]
    mean, std = np.asarray(pred[key]["mean"]), np.asarray(pred[key]["std"])
    assert mean.shape == (3, 57, 43) and std.shape == (3, 57, 43)
    assert np.all(std > 0) and np.all(np.isfinite(mean))
    # oracle on the second task
    t = Task({k: v for k, v in tasks[1].items()})
    t["X_t"] = [(x1[None], x2[None])]
    t["Y_t"] = []
    t["Y_t_aux"] = aux
    ctx, xt, _, aux_t = oracle_inputs(t)
    mean_o, var_o = O.forward(cpu_params(m), ctx, xt, aux_t, m.config.points_per_unit)
    assert rel_err(mean[1], mean_o[0, 0]) < 1e-5
    assert rel_err(std[1], var_o[0, 0].sqrt()) < 1e-5
    # .where(mask) as validate_ERA.py:94-96 uses it
    mask = np.zeros((57, 43), dtype=bool)
    mask[:10] = True
    masked = np.asarray(pred[key]["mean"].where(mask))
    assert np.isnan(masked[:, 10:]).all() and np.isfinite(masked[:, :10]).all()


def test_dec_blk_kernels_match_materialised_path():
    """f, dWf, dbf and d_h from (SetConv on h, then the 1x1) == torch autograd of (1x1 conv, then dense SetConv)."""
    torch.manual_seed(11)
    B, H, W, Nt, Cz = 2, 40, 56, 9, 64
    start1, start2, res = -0.1, -0.05, 0.02
    scale2 = float(np.float32(res * res))
    h = torch.relu(torch.randn(B, 64, H, W, device="cuda")).bfloat16().float()
    Wf = (torch.randn(Cz, 64, device="cuda") * 0.1).requires_grad_(True)
    bf = torch.randn(Cz, device="cuda").requires_grad_(True)
    xt = torch.stack([torch.rand(B, Nt, device="cuda") * 0.7, torch.rand(B, Nt, device="cuda") * 1.0], dim=1).contiguous()
    df = torch.randn(B, Cz, Nt, device="cuda")
    # reference: z = conv1x1(h); f = einsum with dense fp64 weights
    hd = h.double().requires_grad_(True)
    g1 = torch.tensor([np.float32(start1 + i * res) for i in range(H)], device="cuda").double()
    g2 = torch.tensor([np.float32(start2 + j * res) for j in range(W)], device="cuda").double()
    w1 = torch.exp(-0.5 * (xt[:, 0, :, None].double() - g1[None, None]) ** 2 / scale2)
    w2 = torch.exp(-0.5 * (xt[:, 1, :, None].double() - g2[None, None]) ** 2 / scale2)
    z = torch.einsum("ck,bkij->bcij", Wf.double(), hd) + bf.double()[None, :, None, None]
    f_ref = torch.einsum("bcij,bti,btj->bct", z, w1, w2)
    f_ref.backward(df.double())
    # kernels
    hb = _Blk(B, 8, H, W, h.device)
    _cabi.call("cnp_blk_from_nchw_f32", h.data_ptr(), h.stride(0), B, 64, H, W, C.byref(hb.view()), _S())
    g = torch.empty(B, Nt, 64, device="cuda")
    sw = torch.empty(B, Nt, device="cuda")
    f = torch.empty(B, Cz, Nt, device="cuda")
    Wd, bd = Wf.detach().contiguous(), bf.detach().contiguous()
    _cabi.call("cnp_dec_blk_fwd", C.byref(hb.view()), xt.data_ptr(), B, Nt, start1, start2, res, scale2, Wd.data_ptr(),
               bd.data_ptr(), Cz, g.data_ptr(), sw.data_ptr(), f.data_ptr(), _S())
    assert rel_err(f, f_ref) < 1e-5
    dg = torch.empty(B, Nt, 64, device="cuda")
    dW, db = torch.zeros(Cz, 64, device="cuda"), torch.zeros(Cz, device="cuda")
    _cabi.call("cnp_dec_blk_bwd_params", df.data_ptr(), g.data_ptr(), sw.data_ptr(), Wd.data_ptr(), B, Nt, Cz, dg.data_ptr(),
               dW.data_ptr(), db.data_ptr(), _S())
    assert rel_err(dW, Wf.grad) < 1e-4
    assert rel_err(db, bf.grad) < 1e-4
    dh = _Blk(B, 8, H, W, h.device)
    dh.t.fill_(7.0)   # the kernel must overwrite every interior pixel (dense write)
    dh.t.view(-1)[:] = 0  # ... but pads are the caller's zeros
    _cabi.call("cnp_dec_blk_bwd_data", dg.data_ptr(), xt.data_ptr(), B, Nt, start1, start2, res, scale2, C.byref(hb.view()),
               C.byref(dh.view()), _S())
    out = torch.empty(B, 64, H, W, device="cuda")
    _cabi.call("cnp_blk_to_nchw_f32", C.byref(dh.view()), B, 64, out.data_ptr(), out.stride(0), _S())
    ref = hd.grad * (h > 0)          # ReLU mask of the producing layer fused into the kernel
    assert rel_err(out, ref) < 1e-2   # bf16 output rounding


def test_predict_off_grid_targets(static):
    """predict onto station locations (validate.py-style off-grid X_t): [2,N] normalised coordinates."""
    tasks = [make_task(static, 950 + i, all_context=True) for i in range(2)]
    m = small_model("fp32")
    rng = np.random.default_rng(5)
    X = rng.uniform(0.1, 0.9, (2, 23)).astype(np.float32)
    aux = rng.uniform(-1, 1, (5, 23)).astype(np.float32)
    pred = m.predict(tasks, X_t=X, X_t_is_normalised=True, aux_at_targets_override=aux)
    df = pred[list(pred.keys())[0]]
    assert len(df) == 2 * 23 and np.isfinite(df["mean"].values).all() and (df["std"].values > 0).all()
    t = Task({k: v for k, v in tasks[0].items()})
    t["X_t"], t["Y_t"], t["Y_t_aux"] = [X], [], aux
    ctx, xt, _, aux_t = oracle_inputs(t)
    mean_o, var_o = O.forward(cpu_params(m), ctx, xt, aux_t, m.config.points_per_unit)
    assert rel_err(df["mean"].values[:23], mean_o[0, 0]) < 1e-5
    assert rel_err(df["std"].values[:23], var_o[0, 0].sqrt()) < 1e-5


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_multivariable_ragged_batch_matches_oracle(static, precision, tol):
    """configs[3] shape: 8-channel base context (Cin = 20), tasks with different numbers of context stations padded by
    concat_tasks (masks), loss and gradients vs the oracle."""
    from deepsensornz_b200 import concat_tasks
    tasks = []
    for i, nst in enumerate((60, 52, 57)):
        t = make_task(static, 1200 + i, n_stations=nst, context_frac=1.0 - 12.0 / nst, c0_channels=8)
        tasks.append(t)
    assert len({t["X_t"][0].shape[-1] for t in tasks}) == 1          # equal targets, ragged contexts
    task = concat_tasks(tasks)
    m = small_model(precision, dim_yc=(8, 6, 1, 1))
    loss = m.loss_fn(task, normalise=True)
    loss.backward()
    ctx, xt, yt, aux = oracle_inputs(task)
    P = {k: v.clone().requires_grad_(v.dim() > 0) for k, v in cpu_params(m).items()}
    ref = O.loss_fn(P, ctx, xt, yt, aux, m.config.points_per_unit)
    ref.backward()
    assert abs(float(loss) - float(ref)) / abs(float(ref)) < tol
    for n, p in m.model.named_parameters():
        if p.requires_grad and P[n].grad is not None:
            g, r = p.grad.detach().cpu().double().flatten(), P[n].grad.double().flatten()
            cos = float((g @ r) / (g.norm() * r.norm()).clamp(min=1e-300))
            assert cos > (0.9999 if precision == "fp32" else 0.99), (n, cos)


def test_tensor_core_decoder_tail_matches_fused_and_oracle(monkeypatch):
    """On a target grid fine enough for the tcgen05 tail (128 target columns + band within 64 internal columns) the
    tensor-core decoder agrees with the CUDA-core fused decoder (same inputs, fp32 MLP) and with the oracle."""
    static = make_static(seed=11, n_hi=520, with_aux_hi=True)
    task = make_task(static, 777, grid_targets=True)
    m = small_model("bf16")
    monkeypatch.setenv("CNP_DECODE_TC", "0")
    ref = m(task)
    calls = []
    orig = m.engine._call
    monkeypatch.setattr(m.engine, "_call", lambda name, *a, **k: (calls.append(name), orig(name, *a, **k))[1])
    monkeypatch.setenv("CNP_DECODE_TC", "1")
    out = m(task)
    assert "cnp_decode_grid_tc_fwd" in calls          # the tensor-core path really ran
    assert rel_err(out["mean"], ref["mean"]) < 1e-2    # bf16 hidden activations / weights in the MLP
    assert rel_err(out["std"], ref["std"]) < 1e-2
    ctx, xt, _, aux = oracle_inputs(task)
    mean_o, var_o = O.forward(cpu_params(m), ctx, xt, aux, m.config.points_per_unit)
    assert rel_err(out["mean"], mean_o) < 2e-2
    assert rel_err(out["std"], var_o.sqrt()) < 2e-2


def test_predict_graph_replay_matches_eager(static, monkeypatch):
    """predict replays a captured forward for tasks that share a batch signature (station counts padded to a multiple of
    16 with masked points); results are bit-identical to the eager path, also when the signature changes mid-call."""
    counts = (60, 60, 57, 60, 41, 41, 60, 41, 60)  # 57 pads to 64 like 60; 41 -> 48 is a second signature
    tasks = [make_task(static, 3000 + i, all_context=True, n_stations=n) for i, n in enumerate(counts)]
    m = small_model("bf16")
    x1 = np.linspace(0.05, 0.95, 150).astype(np.float32)
    x2 = np.linspace(0.10, 0.90, 170).astype(np.float32)
    aux = np.random.default_rng(3).uniform(-1, 1, (5, 150, 170)).astype(np.float32)
    calls = []
    orig = m.engine._call
    monkeypatch.setattr(m.engine, "_call", lambda name, *a, **k: (calls.append(name), orig(name, *a, **k))[1])
    monkeypatch.setenv("CONVNP_B200_PREDICT_GRAPH", "1")
    monkeypatch.setenv("CONVNP_B200_PREDICT_BATCH", "1")       # one date per forward in both arms (launch counts below)
    pg = m.predict(tasks, X_t=(x1, x2), X_t_is_normalised=True, aux_at_targets_override=aux)
    n_graph = sum(1 for c in calls if c.startswith("cnp_decode_grid"))
    calls.clear()
    monkeypatch.setenv("CONVNP_B200_PREDICT_GRAPH", "0")
    pe = m.predict(tasks, X_t=(x1, x2), X_t_is_normalised=True, aux_at_targets_override=aux)
    n_eager = sum(1 for c in calls if c.startswith("cnp_decode_grid"))
    # eager: one decoder launch per task; graph: two eager + one captured launch per signature, the rest are replays
    assert n_eager == len(tasks) and n_graph == 2 * 3
    key = list(pg.keys())[0]
    assert np.array_equal(np.asarray(pg[key]["mean"]), np.asarray(pe[key]["mean"]))
    assert np.array_equal(np.asarray(pg[key]["std"]), np.asarray(pe[key]["std"]))


def test_predict_context_cache_is_not_fooled_by_recycled_buffers(static):
    """float64 tasks (what xarray hands over) are cast per task into temporaries whose addresses get recycled; the
    static-context cache must never serve one date's field for another: every task equals its own single-task call."""
    tasks = []
    for i in range(6):
        t = make_task(static, 5000 + i, all_context=True)
        t["Y_c"] = [np.asarray(y, dtype=np.float64) for y in t["Y_c"]]
        t["X_c"] = [tuple(np.asarray(v, dtype=np.float64) for v in x) if isinstance(x, tuple) else np.asarray(x, dtype=np.float64)
                    for x in t["X_c"]]
        tasks.append(t)
    m = small_model("fp32")
    x1 = np.linspace(0.05, 0.95, 40).astype(np.float32)
    x2 = np.linspace(0.10, 0.90, 36).astype(np.float32)
    aux = np.random.default_rng(3).uniform(-1, 1, (5, 40, 36)).astype(np.float32)
    pred = m.predict(tasks, X_t=(x1, x2), X_t_is_normalised=True, aux_at_targets_override=aux)
    key = list(pred.keys())[0]
    mean = np.asarray(pred[key]["mean"])
    for i, t in enumerate(tasks):
        single = m.predict([t], X_t=(x1, x2), X_t_is_normalised=True, aux_at_targets_override=aux)
        assert np.array_equal(np.asarray(single[key]["mean"])[0], mean[i]), i
    assert not np.array_equal(mean[0], mean[1])


class _FakeProcessor:
    """Stand-in for deepsensor's DataProcessor: mean/std normalisation of the target variable (affine), or a
    non-affine map to exercise the host fallback."""

    def __init__(self, mean=281.5, std=7.25, affine=True):
        self.mean, self.std, self.affine, self.calls = mean, std, affine, 0

    def map_array(self, data, var_ID, method=None, unnorm=False, add_offset=True):
        self.calls += 1
        data = np.asarray(data)
        assert unnorm
        if not self.affine:
            return np.exp(data * 0.1).astype(data.dtype)
        out = data * self.std
        return out + self.mean if add_offset else out


def test_predict_unnormalises_on_the_device(static):
    """An affine data-processor map is applied on the GPU before the read-back (no host pass over the [T,N1,N2]
    results); a non-affine one falls back to ``map_array`` on the host.  Both equal the normalised prediction mapped
    by hand."""
    tasks = [make_task(static, 7000 + i, all_context=True) for i in range(3)]
    m = small_model("fp32")
    x1 = np.linspace(0.05, 0.95, 41).astype(np.float32)
    x2 = np.linspace(0.10, 0.90, 37).astype(np.float32)
    aux = np.random.default_rng(3).uniform(-1, 1, (5, 41, 37)).astype(np.float32)
    kw = dict(X_t=(x1, x2), X_t_is_normalised=True, aux_at_targets_override=aux)
    raw = m.predict(tasks, **kw)
    key = list(raw.keys())[0]
    mean_n, std_n = np.asarray(raw[key]["mean"]), np.asarray(raw[key]["std"])
    dp = _FakeProcessor()
    m.data_processor = dp
    out = m.predict(tasks, **kw)
    assert dp.calls == 2                                  # only the two 3-point probes, no full-array call
    assert np.allclose(np.asarray(out[key]["mean"]), mean_n * np.float32(dp.std) + np.float32(dp.mean), rtol=1e-6, atol=1e-5)
    assert np.allclose(np.asarray(out[key]["std"]), std_n * np.float32(dp.std), rtol=1e-6)
    dp2 = _FakeProcessor(affine=False)
    m.data_processor = dp2
    out2 = m.predict(tasks, **kw)
    assert dp2.calls == 4                                 # probes rejected, then the host fallback for mean and std
    assert np.allclose(np.asarray(out2[key]["mean"]), np.exp(mean_n * 0.1), rtol=1e-6)
    unn = m.predict(tasks, unnormalise=False, **kw)
    assert np.array_equal(np.asarray(unn[key]["mean"]), mean_n)
    m.data_processor = None


def test_batched_predict_equals_one_task_per_forward(monkeypatch):
    """predict runs several dates per forward (CONVNP_B200_PREDICT_BATCH, default 4): ragged station counts are padded with
    masked-out copies of a real point, static sets are uploaded once; every date's mean / std equals its single-task
    forward (validate_ERA.py:88-92 call form)."""
    static = make_static(seed=7, n_hi=200, with_aux_hi=True)
    m = small_model("bf16", seed=8)
    rng = np.random.default_rng(5)
    tasks = [make_task(static, 8100 + i, n_stations=int(rng.integers(120, 201)), all_context=True) for i in range(7)]
    kw = dict(X_t=(static.x_hi, static.x_hi), X_t_is_normalised=True, aux_at_targets_override=static.aux_hi)
    res = {}
    for nb in ("1", "4", "3"):
        monkeypatch.setenv("CONVNP_B200_PREDICT_BATCH", nb)
        pred = m.predict(tasks, **kw)
        key = list(pred.keys())[0]
        res[nb] = (np.asarray(pred[key]["mean"]).copy(), np.asarray(pred[key]["std"]).copy())
        assert res[nb][0].shape == (7, 200, 200) and np.isfinite(res[nb][0]).all()
    for nb in ("4", "3"):
        for a, b in zip(res["1"], res[nb]):
            assert np.allclose(a, b, rtol=1e-5, atol=1e-6), float(np.abs(a - b).max())
