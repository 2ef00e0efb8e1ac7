"""CPU suite: host logic of the boundary -- Task ops, bit-exact batching, discretisation, C-ABI symbols."""
import os
import re

import numpy as np
import pytest
import torch

import deepsensornz_b200 as ds
from deepsensornz_b200 import Masked, Task, concat_tasks, _cabi
from deepsensornz_b200.discretisation import discretise, discretise_1d
from deepsensornz_b200.dist import shard_tasks
from deepsensornz_b200.synthetic import make_static, make_task
from oracle import convnp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def static():
    return make_static(seed=7, n_hi=120)


def test_task_layout_matches_reference(static):
    """Layout of nzdownscale/downscaler/train.py:560-637: gridded X_c tuples of [1,N], Y_c [C,N1,N2]; off-grid [2,N]."""
    t = make_task(static, 3)
    assert isinstance(t["X_c"][0], tuple) and t["X_c"][0][0].shape == (1, 140)
    assert t["Y_c"][0].shape == (3, 140, 140) and t["Y_c"][3].shape == (1, 160)
    assert t["X_c"][3].shape == (2, 160) and t["X_t"][0].shape == (2, 40) and t["Y_t_aux"].shape == (5, 40)
    assert np.all(np.diff(t["X_c"][0][0][0]) < 0)          # ERA5 latitude descending
    assert t["X_c"][3].dtype == np.float32
    # context / target are complementary station subsets (train.py:529-558)
    allx = np.concatenate([t["X_c"][3], t["X_t"][0]], axis=1)
    assert np.unique(allx, axis=1).shape[1] == 200


def test_task_ops_and_masks(static):
    t = make_task(static, 4)
    m = ds.ConvNP.modify_task(t)
    assert m["ops"] == ["batch_dim", "float32", "numpy_mask", "nps_mask"]
    y0 = m["Y_c"][0]
    assert isinstance(y0, Masked) and y0.mask.shape == (1, 1, 140, 140) and y0.y.shape == (1, 3, 140, 140)
    raw = t["Y_c"][0]
    expect = (~np.isnan(raw).any(axis=0)).astype(np.float32)
    assert np.array_equal(y0.mask[0, 0], expect)            # bit-exact mask
    assert not np.isnan(y0.y).any() and np.all(y0.y[0][np.isnan(raw)] == 0)   # only the NaN entries are zero-filled
    assert isinstance(m["Y_c"][1], np.ndarray)               # no NaNs -> stays a plain array


def test_concat_tasks_bit_exact(static):
    ts = [make_task(static, 10 + i, n_stations=200 - 10 * i, context_frac=(160 - 10 * i) / (200 - 10 * i))
          for i in range(3)]                                  # N_c = 160,150,140 ; N_t = 40 each
    assert [t["X_t"][0].shape[1] for t in ts] == [40, 40, 40]
    m = concat_tasks(ts)
    x, y = m["X_c"][3], m["Y_c"][3]
    assert x.shape == (3, 2, 160) and isinstance(y, Masked)
    for i, t in enumerate(ts):
        n = t["X_c"][3].shape[1]
        assert np.array_equal(x[i, :, :n], t["X_c"][3]) and np.all(x[i, :, n:] == 0)
        assert np.array_equal(y.y[i, :, :n], t["Y_c"][3]) and np.all(y.y[i, :, n:] == 0)
        assert np.all(y.mask[i, 0, :n] == 1) and np.all(y.mask[i, 0, n:] == 0)
        assert np.array_equal(m["X_t"][0][i], t["X_t"][0]) and np.array_equal(m["Y_t_aux"][i], t["Y_t_aux"])
    assert m["X_c"][2][0].shape == (3, 1, 120)
    with pytest.raises(ValueError):
        concat_tasks([ts[0], make_task(static, 99, n_stations=190)])   # different N_t (38 vs 40)
    with pytest.raises(ValueError):
        concat_tasks([m, m])                                            # already masked


def test_group_then_concat_like_train_py(static):
    """batch_data_by_num_stations (train.py:448-475) + train_epoch batching = one batch per station count."""
    ts = [make_task(static, 30 + i, n_stations=200 if i % 2 == 0 else 150) for i in range(6)]
    groups = {}
    for t in ts:
        groups.setdefault(t["X_t"][0].shape[1], []).append(t)
    assert sorted(groups) == [30, 40]
    for g in groups.values():
        assert concat_tasks(g)["X_t"][0].shape[0] == 3


def test_discretisation_matches_oracle(static):
    t = ds.ConvNP.modify_task(make_task(static, 5))
    xs = t["X_c"] + [t["X_t"][0]]
    for ppu in (50, 250, 500, 37.5):
        g = discretise(xs, ppu, 0.1, 8)
        xt = [tuple(torch.from_numpy(v) for v in x) if isinstance(x, tuple) else torch.from_numpy(x) for x in xs]
        (s1, n1), (s2, n2), res = O.discretise(xt, ppu, 0.1, 8)
        assert (g.start1, g.n1, g.start2, g.n2, g.res) == (s1, n1, s2, n2, res)
        assert np.array_equal(g.points(0), O.grid_points(s1, n1, res).numpy())
    assert discretise_1d(0.0, 1.0, 250)[1] == 304


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "convnp_b200.h")).read()
    hdr = re.sub(r"#ifdef CNP_LEGACY_CONV_TC.*?#endif", "", hdr, flags=re.S)     # only in a `make LEGACY=1` build
    declared = set(re.findall(r"\b(cnp_[a-z0-9_]+)\s*\(", hdr))
    lib = _cabi.lib()                                     # dlopen works without a GPU; no compute call is made
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/convnp_b200.h but not exported"
    assert declared == set(_cabi.exported_symbols())
    assert lib.cnp_version() >= 100


def test_no_cpu_fallback(static):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = ds.ConvNP(dim_yc=(3, 6, 1, 1), dim_yt=1, dim_aux_t=5, internal_density=50, verbose=False)
    with pytest.raises(_cabi.CnpError):
        m.loss_fn(make_task(static, 1), normalise=True)
    with pytest.raises(RuntimeError):
        ds.set_gpu_default_device()


def test_model_container_contract(tmp_path):
    m = ds.ConvNP(dim_yc=(8, 6, 1, 1), dim_yt=1, dim_aux_t=5, internal_density=250, verbose=False)
    assert ds.num_params(m.model) == 1_145_346 + 5                     # SURVEY Appendix B (+5 frozen log-scales)
    assert all(not p.requires_grad for p in m.model.encoder.parameters())
    for p in m.model.encoder.parameters():                              # train.py:257 freeze pattern is a no-op
        p.requires_grad = False
    sd = m.model.state_dict()
    m.save(str(tmp_path / "mdl"))
    m2 = ds.ConvNP(None, None, str(tmp_path / "mdl"), verbose=False)
    for k, v in m2.model.state_dict().items():
        assert torch.equal(v.cpu(), sd[k].cpu()), k
    assert m2.config.dim_yc == (8, 6, 1, 1)
    with pytest.raises(NotImplementedError):
        ds.ConvNP(dim_yc=(1,), dim_yt=1, dim_aux_t=5, internal_density=50, likelihood="gnp", verbose=False)


def test_shard_tasks_partition(static):
    ts = [make_task(static, 50 + i, n_stations=200 if i < 8 else 150) for i in range(13)]
    shards = [shard_tasks(ts, r, 2) for r in range(2)]
    assert [len(s) for s in shards] == [6, 6]                           # 8 -> 4+4, 5 -> 2+2 (remainder dropped)
    ids = [t["time"] for s in shards for t in s]
    assert len(set(ids)) == len(ids)


def test_upstream_style_checkpoint_loads_by_shape_sequence():
    """``model.model.load_state_dict(torch.load(path))`` (train.py:243-251, validate_ERA.py:100-111) with a
    checkpoint whose key names are not this module's: mapped by shape sequence, refused when the architecture differs."""
    from deepsensornz_b200.model import ConvNPConfig, ConvNPModule, map_upstream_state_dict
    cfg = ConvNPConfig(dim_yc=(3, 6, 1, 1), dim_aux_t=5, encoder_scales=(0.01,) * 4)
    torch.manual_seed(3)
    src = ConvNPModule(cfg)
    own = src.state_dict()
    # an upstream-looking checkpoint: other names, another interleaving of the tensors (all biases before all weights;
    # the order WITHIN a shape is the architecture's, which is what the mapping relies on), scalars shaped (1,)
    order = sorted(own, key=lambda k: own[k].dim())
    foreign = {}
    for i, k in enumerate(order):
        v = own[k].clone()
        foreign[f"model.coder_{i // 7}.links.{i}.net.{k.split('.')[-1]}"] = v.reshape(1) if v.numel() == 1 else v
    torch.manual_seed(4)
    dst = ConvNPModule(cfg)
    assert not torch.equal(dst.decoder.unet.final_linear.weight, src.decoder.unet.final_linear.weight)
    dst.load_state_dict(foreign)
    for (k, a), (_, b) in zip(dst.state_dict().items(), own.items()):
        assert torch.equal(a, b), k
    # own checkpoints keep loading by name
    dst2 = ConvNPModule(cfg)
    dst2.load_state_dict(own)
    assert torch.equal(dst2.decoder.mlp.layers[0].weight, src.decoder.mlp.layers[0].weight)
    # a checkpoint of another architecture is refused, not guessed
    other = ConvNPModule(ConvNPConfig(dim_yc=(8, 6, 1, 1), dim_aux_t=5, encoder_scales=(0.01,) * 4)).state_dict()
    with pytest.raises(RuntimeError, match="does not match"):
        dst.load_state_dict({f"x.{i}": v for i, v in enumerate(other.values())})
    missing = dict(list(foreign.items())[:-1])
    with pytest.raises(RuntimeError, match="does not match"):
        dst.load_state_dict(missing)
    table = map_upstream_state_dict(foreign, dst.state_dict())
    assert set(table) == set(own)


def test_other_unet_widths_run_the_fp32_kernels():
    """train_downscaling.py:116-117 lets the user set unet_channels: widths other than 64 are not on the tensor-core path;
    the model is built with a warning and runs in fp32."""
    with pytest.warns(UserWarning, match="64"):
        m = ds.ConvNP(dim_yc=(3, 6, 1, 1), dim_yt=1, dim_aux_t=5, internal_density=50, encoder_scales=(0.01,) * 4,
                      decoder_scale=0.02, unet_channels=(32, 32, 32, 32), precision="bf16", verbose=False)
    assert m.engine.precision == "fp32"
