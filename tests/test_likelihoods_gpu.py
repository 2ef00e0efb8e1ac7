"""-m gpu: the Bernoulli-Gamma and spikes-Beta likelihood heads (csrc/head.cu) against the CPU oracle -- the heads
nzdownscale selects for precipitation and humidity (nzdownscale/dataprocess/config.py:162-169).  fp32 path 1e-5 on
loss / mean / std, gradients of every parameter against oracle autograd; bf16 UNet 2e-2."""
import numpy as np
import pytest
import torch

from deepsensornz_b200 import Task, concat_tasks
from deepsensornz_b200.synthetic import make_static, make_task
from oracle import convnp_oracle as O
from oracle.task_tensors import task_tensors
from tests.util import cpu_params, rel_err, small_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def static():
    return make_static(seed=7, n_hi=200, with_aux_hi=True)


def _tasks(static, kind, n=3, seed0=7000):
    """Targets in the likelihood's support: precipitation-like (>= 0, many exact zeros) or humidity-like ([0, 1] with
    exact 0 and 1), plus one missing observation."""
    out = []
    for i in range(n):
        t = make_task(static, seed0 + i)
        t = Task({k: (list(v) if isinstance(v, list) else v) for k, v in t.items()})
        rng = np.random.default_rng(seed0 + 100 + i)
        nt = t["Y_t"][0].shape[1]
        if kind == "bernoulli-gamma":
            y = rng.gamma(1.5, 0.8, nt).astype(np.float32)
            y[rng.random(nt) < 0.4] = 0.0
        else:
            y = rng.beta(2.0, 1.2, nt).astype(np.float32)
            u = rng.random(nt)
            y[u < 0.15] = 0.0
            y[u > 0.9] = 1.0
        y[3] = np.nan
        t["Y_t"][0] = y[None]
        out.append(t)
    return out


def _oracle(m, tasks, kind):
    ctx, xt, yt, aux = task_tensors(tasks)
    # the product drops NaN targets per task before batching (concat_tasks); the oracle skips them in the sum: same loss
    P = {k: v.clone().requires_grad_(v.dim() > 0) for k, v in cpu_params(m).items()}
    mean, var, info = O.forward(P, ctx, xt, aux, m.config.points_per_unit, return_internal=True, likelihood=kind)
    loss = -O.loglik_spike_slab(info["o"], yt, kind, True).mean()
    loss.backward()
    return mean.detach(), var.detach(), float(loss), {k: v.grad for k, v in P.items() if v.grad is not None}


@pytest.mark.parametrize("kind", ["bernoulli-gamma", "cnp-spikes-beta"])
@pytest.mark.parametrize("precision,tol,gtol", [("fp32", 1e-5, 1e-4), ("bf16", 2e-2, 0.15)])
def test_spike_slab_heads_match_oracle(static, kind, precision, tol, gtol):
    tasks = _tasks(static, kind)
    m = small_model(precision, likelihood=kind, seed=21)
    assert m.model.mlp_dims()[-1] == (4 if kind == "bernoulli-gamma" else 5)
    mean_o, var_o, loss_o, grads_o = _oracle(m, tasks, kind)
    # single task with its NaN target still in place: mean / std at every target
    t0 = tasks[0]
    pred = m(t0)
    assert rel_err(pred["mean"], mean_o[:1]) < tol
    assert rel_err(pred["std"], var_o[:1].sqrt()) < tol
    loss = m.loss_fn(concat_tasks(tasks), normalise=True)
    assert abs(float(loss) - loss_o) < tol * abs(loss_o)
    loss.backward()
    for n, p in m.model.named_parameters():
        if p.requires_grad:
            assert p.grad is not None and torch.isfinite(p.grad).all(), n
            assert rel_err(p.grad, grads_o[n]) < gtol, n
    # the loss is run-to-run identical (fixed-order float64 reduction)
    with torch.no_grad():
        a, b = float(m.loss_fn(tasks[1], normalise=True)), float(m.loss_fn(tasks[1], normalise=True))
    assert a == b


@pytest.mark.parametrize("kind", ["bernoulli-gamma", "cnp-spikes-beta"])
def test_spike_slab_on_grid_predict(static, kind):
    """predict onto a target grid with a non-Gaussian head: mean / std of the spikes-and-slab distribution."""
    m = small_model("bf16", likelihood=kind, seed=22)
    t = make_task(static, 7100, all_context=True)
    xg = static.x_hi[::4]
    aux = static.aux_hi[:, ::4, ::4].copy()
    pred = m.predict([t], X_t=(xg, xg), X_t_is_normalised=True, aux_at_targets_override=aux)
    key = list(pred.keys())[0]
    mean, std = np.asarray(pred[key]["mean"])[0], np.asarray(pred[key]["std"])[0]
    tg = Task({k: v for k, v in t.items()})
    tg["X_t"], tg["Y_t"], tg["Y_t_aux"] = [(xg[None], xg[None])], [], aux
    ctx, xt, _, auxt = task_tensors([tg])
    mean_o, var_o = O.forward(cpu_params(m), ctx, xt, auxt, m.config.points_per_unit, likelihood=kind)
    assert rel_err(mean, mean_o[0, 0]) < 2e-2 and rel_err(std, var_o[0, 0].sqrt()) < 2e-2
    if kind == "bernoulli-gamma":
        assert (mean >= 0).all()
    else:
        assert (mean >= 0).all() and (mean <= 1).all()


def test_gnp_is_refused_with_a_clear_message():
    with pytest.raises(NotImplementedError, match="gnp"):
        small_model("fp32", likelihood="gnp")
