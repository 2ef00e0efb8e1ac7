"""CPU suite: the oracle against analytic known answers and the committed golden vectors.

Parity is UNPINNED against the reference itself (no golden vectors / importable packages):
these tests pin the oracle to closed forms and to its own committed outputs."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import convnp_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_discretisation_known_sizes():
    # coords spanning [0,1]: ppu 250 -> 304 points (SURVEY A.2 variant 1), ppu 500 -> 608, ppu 50 -> 64
    assert O.discretise_1d(0.0, 1.0, 250, 0.1, 8)[1] == 304
    assert O.discretise_1d(0.0, 1.0, 500, 0.1, 8)[1] == 608
    assert O.discretise_1d(0.0, 1.0, 50, 0.1, 8)[1] == 64
    s, n, res = O.discretise_1d(0.0, 1.0, 250, 0.1, 8)
    assert n % 8 == 0 and s <= -0.1 - res + 1e-12 and s + (n - 1) * res >= 1.1
    assert abs(round(s / res) * res - s) < 1e-12  # snapped to the grid


def test_single_point_setconv_is_gaussian_bump():
    g = O.grid_points(-0.1, 32, 0.02)
    x = torch.tensor([[[0.3], [0.17]]])       # [B=1,2,N=1]
    y = torch.tensor([[[2.5]]])
    ls = torch.tensor(math.log(0.03))
    h = O.encode_set(x, y, None, g, g, ls, eps=0.0)
    w1 = torch.exp(-0.5 * (0.3 - g) ** 2 / 0.03 ** 2)
    w2 = torch.exp(-0.5 * (0.17 - g) ** 2 / 0.03 ** 2)
    dens = w1[:, None] * w2[None, :]
    assert torch.allclose(h[0, 0], dens, rtol=1e-5, atol=1e-30)
    near = dens > 1e-20
    assert torch.allclose(h[0, 1][near], torch.full_like(h[0, 1][near], 2.5), rtol=1e-5)


def test_density_normalisation_and_mask():
    torch.manual_seed(0)
    g = O.grid_points(0.0, 16, 0.05)
    x = (torch.linspace(0, 0.75, 12)[None, None], torch.linspace(0, 0.75, 12)[None, None])
    y = torch.ones(1, 1, 12, 12)
    m = torch.ones(1, 1, 12, 12)
    m[0, 0, :, 6:] = 0
    ls = torch.tensor(math.log(0.04))
    h = O.encode_set(x, y * 7.0, m, g, g, ls, eps=1e-2)
    d = h[0, 0]
    assert torch.allclose(h[0, 1], 7.0 * d / (d + 1e-2), rtol=1e-5)   # constant field -> 7 * d/(d+eps)
    full = O.encode_set(x, y * 7.0, None, g, g, ls, eps=1e-2)
    assert float(full[0, 0].sum()) > float(d.sum())                   # masking removes density


def test_gridded_equals_offgrid_on_same_points():
    torch.manual_seed(1)
    x1, x2 = torch.rand(5).sort().values, torch.rand(4).sort().values
    y = torch.randn(1, 2, 5, 4)
    g = O.grid_points(-0.1, 24, 0.05)
    ls = torch.tensor(math.log(0.07))
    a = O.encode_set((x1[None, None], x2[None, None]), y, None, g, g, ls)
    xo = torch.stack(torch.meshgrid(x1, x2, indexing="ij")).reshape(1, 2, -1)
    b = O.encode_set(xo, y.reshape(1, 2, -1), None, g, g, ls)
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)


def test_nll_closed_form_and_nan_targets():
    mean = torch.tensor([[[0.5, -1.0, 2.0]]])
    var = torch.tensor([[[0.25, 1.0, 4.0]]])
    y = torch.tensor([[[1.0, float("nan"), 0.0]]])
    lp = O.loglik(mean, var, y, normalise=False)
    exp = -0.5 * (math.log(2 * math.pi) + math.log(0.25) + 0.25 / 0.25) \
          - 0.5 * (math.log(2 * math.pi) + math.log(4.0) + 4.0 / 4.0)
    assert abs(float(lp) - exp) < 1e-12
    assert abs(float(O.loglik(mean, var, y, normalise=True)) - exp / 2) < 1e-12
    assert lp.dtype == torch.float64


def test_head_variance_is_softplus_plus_eps():
    o = torch.tensor([[[0.3], [-2.0]]])
    mean, var = O.het_gaussian(o)
    assert float(mean) == pytest.approx(0.3)
    assert float(var) == pytest.approx(1e-6 + math.log1p(math.exp(-2.0)), rel=1e-6)


def test_unet_shapes_and_skip_order():
    torch.manual_seed(2)
    from tests.util import small_model, cpu_params
    m = small_model("fp32")
    P = cpu_params(m)
    x = torch.randn(1, m.config.in_channels, 32, 40)
    z = O.unet(P, x)
    assert z.shape == (1, 64, 32, 40)
    # zeroing the weights that read the *second* half of the last concat removes the dependence on the deep path
    P2 = dict(P)
    w = P["decoder.unet.after_turn_layers.0.weight"].clone()
    w[:, 64:] = 0
    P2["decoder.unet.after_turn_layers.0.weight"] = w
    for k in list(P2):
        if "before_turn_layers.1" in k:
            P2[k] = torch.randn_like(P2[k])
    z_a = O.unet(P2, x)
    P3 = dict(P2)
    for k in list(P3):
        if "before_turn_layers.1" in k:
            P3[k] = torch.randn_like(P3[k])
    assert torch.allclose(z_a, O.unet(P3, x))  # (skip, upsampled) order: channels 64.. are the deep path


@pytest.mark.parametrize("name", ["g1_single", "g2_batch3", "g3_multivar"])
def test_oracle_reproduces_golden(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    from tests.util import cpu_params, oracle_inputs
    m, task = mg.build(name)
    ctx, xt, yt, aux = oracle_inputs(task)
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    mean, var = O.forward(cpu_params(m), ctx, xt, aux, m.config.points_per_unit)
    np.testing.assert_allclose(mean.numpy(), gold["mean"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(var.numpy(), gold["var"], rtol=2e-5, atol=1e-7)
    loss = float(-O.loglik(mean, var, yt, True).mean())
    assert abs(loss - float(gold["loss"])) < 1e-5 * abs(float(gold["loss"]))


def test_oracle_gradients_finite_difference():
    from tests.util import small_model, cpu_params, oracle_inputs
    from deepsensornz_b200.synthetic import make_static, make_task
    m = small_model("fp32", ppu=30)
    t = make_task(make_static(seed=7, n_hi=100), 11, n_stations=60)
    ctx, xt, yt, aux = oracle_inputs(t)
    P = {k: v.double() if v.dim() > 0 else v for k, v in cpu_params(m).items()}
    ctx = [(tuple(v.double() for v in x) if isinstance(x, tuple) else x.double(), y.double(),
            None if mk is None else mk.double()) for x, y, mk in ctx]
    key = "decoder.mlp.layers.0.bias"
    P[key].requires_grad_(True)
    xt, yt, aux = xt.double(), yt.double(), aux.double()
    loss = O.loss_fn(P, ctx, xt, yt, aux, 30)
    (g,) = torch.autograd.grad(loss, P[key])
    eps = 1e-6
    with torch.no_grad():
        Pp = dict(P)
        d = torch.zeros_like(P[key])
        d[3] = eps
        Pp[key] = P[key] + d
        lp = O.loss_fn(Pp, ctx, xt, yt, aux, 30)
        Pp[key] = P[key] - d
        lm = O.loss_fn(Pp, ctx, xt, yt, aux, 30)
    assert abs(float((lp - lm) / (2 * eps)) - float(g[3])) < 1e-6 * max(1.0, abs(float(g[3])))
