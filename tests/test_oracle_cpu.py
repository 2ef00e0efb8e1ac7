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


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE shapes (tests/golden/s*.npz, tests/golden/make_golden_baseline.py)
# ---------------------------------------------------------------------------------------------------------------------
def _mgb():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_baseline", os.path.join(GOLD, "make_golden_baseline.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_oracle_reproduces_baseline_shape_golden():
    """S1 (one task, 304 x 304 internal grid, 1400 x 1400 land-mask context): the oracle still gives the committed
    numbers -- inputs regenerated from the seeds, through oracle/task_tensors.py."""
    from oracle.task_tensors import task_tensors
    from tests.util import cpu_params
    mg = _mgb()
    gold = np.load(os.path.join(GOLD, "s1_single.npz"))
    m = mg.build_model("s1_single")
    ctx, xt, yt, aux = task_tensors(mg.build_tasks("s1_single"))
    with torch.no_grad():
        mean, var, info = O.forward(cpu_params(m), ctx, xt, aux, mg.PPU, return_internal=True)
        loss = -O.loglik(mean, var, yt, True).mean()
    assert info["grid"][0][1] == 304 and info["grid"][1][1] == 304
    assert np.allclose(mean.numpy(), gold["mean"], rtol=1e-4, atol=1e-5)
    assert np.allclose(info["z"][mg.Z_SAMPLE].numpy(), gold["z_sample"], rtol=1e-4, atol=1e-5)
    assert abs(float(loss) - float(gold["loss"])) < 1e-5 * abs(float(gold["loss"]))


@pytest.mark.parametrize("name", ["s1_single", "s2_batch4", "s4_multivar8"])
def test_sensitive_weights_make_the_unet_matter(name):
    """The golden weights are scaled so that the loss depends on the convolutions: zeroing after_turn_layers.0 moves
    the oracle's loss by far more than 1 % (with torch's default init it moves by 1e-3)."""
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    assert abs(float(gold["loss_after0_zeroed"]) - float(gold["loss"])) > 0.05 * abs(float(gold["loss"]))
    assert float(gold["z_abs_mean"]) > 0.1


def test_task_tensors_restatement_equals_product_batching():
    """oracle/task_tensors.py (independent restatement of modify_task + concat_tasks) and the
    product's concat_tasks -> modify_task -> convert_task_to_nps_args agree bit for bit, ragged context sets included."""
    from deepsensornz_b200 import concat_tasks
    from deepsensornz_b200.synthetic import make_static, make_task
    from oracle.task_tensors import task_tensors
    from tests.util import oracle_inputs
    static = make_static(seed=7, n_hi=200)
    tasks = [make_task(static, 900 + i, n_stations=200 - 10 * i) for i in range(3)]
    nt = min(t["X_t"][0].shape[-1] for t in tasks)
    for t in tasks:
        t["X_t"][0], t["Y_t"][0], t["Y_t_aux"] = t["X_t"][0][:, :nt], t["Y_t"][0][:, :nt], t["Y_t_aux"][:, :nt]

    def eq(u, v):
        if u is None or v is None:
            return u is None and v is None
        if isinstance(u, tuple):
            return all(eq(a, b) for a, b in zip(u, v))
        return u.shape == v.shape and u.dtype == v.dtype and torch.equal(u, v)

    for group in (tasks, tasks[:1]):
        a = oracle_inputs(concat_tasks(group) if len(group) > 1 else group[0])
        b = task_tensors(group)
        assert len(a[0]) == len(b[0])
        for (x, y, m), (x2, y2, m2) in zip(a[0], b[0]):
            assert eq(x, x2) and eq(y, y2) and eq(m, m2)
        assert eq(a[1], b[1]) and eq(a[2], b[2]) and eq(a[3], b[3])
    # the ragged set really was padded and masked
    m3 = task_tensors(tasks)[0][3][2]
    assert m3 is not None and float(m3[0].sum()) == 160 and float(m3[2].sum()) == 144


# ---------------------------------------------------------------------------------------------------------------------
# spikes-and-slab likelihoods: closed-form known answers
# ---------------------------------------------------------------------------------------------------------------------
def _inv_softplus(v):
    return math.log(math.expm1(v))


def test_bernoulli_gamma_known_answers():
    """k = 1 makes the slab an exponential: log p(y) = log p_slab - y / scale - log scale; y == 0 picks the spike."""
    scale, l0, l1 = 2.0, 0.3, -0.4
    o = torch.tensor([_inv_softplus(1.0 - 1e-6), _inv_softplus(scale - 1e-6), l0, l1], dtype=torch.float32).view(1, 4, 1).repeat(1, 1, 3)
    y = torch.tensor([[[0.0, 1.5, float("nan")]]])
    lp = O.spike_slab_logpdf(o, y, "bernoulli-gamma")
    lz = math.log(math.exp(l0) + math.exp(l1))
    assert abs(float(lp[0, 0]) - (l0 - lz)) < 1e-6
    assert abs(float(lp[0, 1]) - ((l1 - lz) - 1.5 / scale - math.log(scale))) < 1e-5
    assert math.isnan(float(lp[0, 2]))
    ll = O.loglik_spike_slab(o, y, "bernoulli-gamma", normalise=True)
    assert abs(float(ll) - 0.5 * (float(lp[0, 0]) + float(lp[0, 1]))) < 1e-12          # NaN target skipped, / 2
    mean, var = O.spike_slab_moments(o, "bernoulli-gamma")
    p1 = math.exp(l1 - lz)
    assert abs(float(mean[0, 0, 0]) - p1 * scale) < 1e-5                                  # E = p_slab * k * scale
    assert abs(float(var[0, 0, 0]) - (p1 * 2 * scale ** 2 - (p1 * scale) ** 2)) < 1e-4    # E[x^2] = p k (k+1) scale^2


def test_spikes_beta_known_answers():
    """alpha = beta = 1 makes the slab uniform on (0, 1): log p(y) = log p_slab; y in {0, 1} picks a spike."""
    l = (0.2, -0.1, 0.5)
    one = _inv_softplus(1.0 - 1e-6)
    o = torch.tensor([one, one, *l], dtype=torch.float32).view(1, 5, 1).repeat(1, 1, 3)
    y = torch.tensor([[[0.0, 1.0, 0.37]]])
    lp = O.spike_slab_logpdf(o, y, "cnp-spikes-beta")
    lz = math.log(sum(math.exp(v) for v in l))
    for i in range(3):
        assert abs(float(lp[0, i]) - (l[i] - lz)) < 1e-5
    mean, var = O.spike_slab_moments(o, "cnp-spikes-beta")
    p = [math.exp(v - lz) for v in l]
    m = p[1] + p[2] * 0.5
    assert abs(float(mean[0, 0, 0]) - m) < 1e-6
    assert abs(float(var[0, 0, 0]) - (p[1] + p[2] / 3.0 - m * m)) < 1e-6                   # E[U^2] = 1/3
    # a non-trivial Beta against scipy
    from scipy.stats import beta as sbeta
    o2 = torch.tensor([_inv_softplus(2.5), _inv_softplus(0.7), *l], dtype=torch.float32).view(1, 5, 1)
    lp2 = float(O.spike_slab_logpdf(o2, torch.tensor([[[0.8]]]), "cnp-spikes-beta"))
    a, b = 2.5 + 1e-6, 0.7 + 1e-6
    assert abs(lp2 - ((l[2] - lz) + sbeta.logpdf(0.8, a, b))) < 1e-5
