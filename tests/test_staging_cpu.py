"""CPU suite: the device-resident task pipeline of train_epoch (deepsensornz_b200/staging.py) builds exactly the batch
that concat_tasks + ConvNP.stage_task build (reference: upstream concat_tasks inside train_epoch,
nzdownscale/downscaler/train.py:388-394), without copying or comparing the static fields."""
import numpy as np
import pytest
import torch

from deepsensornz_b200 import ConvNP, Masked, concat_tasks
from deepsensornz_b200.staging import BatchStager
from deepsensornz_b200.synthetic import make_static, make_task
from deepsensornz_b200.task import is_batch_broadcast
from tests.util import small_model


@pytest.fixture(scope="module")
def static():
    return make_static(seed=7, n_hi=120)


def _ragged(static, seed0, n=3):
    ts = [make_task(static, seed0 + i, n_stations=200 - 10 * i, context_frac=(160 - 10 * i) / (200 - 10 * i))
          for i in range(n)]
    assert len({t["X_t"][0].shape[1] for t in ts}) == 1
    return ts


def test_concat_keeps_shared_static_fields_as_zero_stride_views(static):
    ts = _ragged(static, 10)
    m = concat_tasks(ts)
    # static sets: same values as a stacked copy, no copy made
    for k in (1, 2):
        y = m["Y_c"][k]
        assert is_batch_broadcast(y) and y.shape == (3,) + ts[0]["Y_c"][k].shape
        assert np.array_equal(y, np.stack([t["Y_c"][k] for t in ts]))
        assert np.shares_memory(y, ts[0]["Y_c"][k])
    # per-date base grid: a real stack, masked
    y0 = m["Y_c"][0]
    assert isinstance(y0, Masked) and not is_batch_broadcast(y0.y)
    # a shared field WITH NaNs is masked once and handed out as zero-stride views
    nanfield = static.c1.copy()
    nanfield[:, :5] = np.nan
    for t in ts:
        t["Y_c"][1] = nanfield
    mm = concat_tasks(ts)["Y_c"][1]
    assert isinstance(mm, Masked) and is_batch_broadcast(mm.y) and is_batch_broadcast(mm.mask)
    assert mm.mask.shape == (3, 1, 140, 140) and float(mm.mask[2, 0, :5].sum()) == 0 and not np.isnan(mm.y).any()
    # copies of the same field (not the same buffer) are stacked as before
    for t in ts:
        t["Y_c"][1] = static.c1.copy()
    assert not is_batch_broadcast(concat_tasks(ts)["Y_c"][1])


def test_stager_equals_concat_then_stage(static):
    m = small_model("fp32")
    eng = m.engine
    st = BatchStager(eng)
    for seed0 in (100, 200, 300, 400):          # four batches: the ring (3 slots) wraps around
        ts = _ragged(static, seed0)
        assert st.fast_path_ok(ts)
        hb = st.build(ts)
        ref = m.stage_task(concat_tasks(ts), pinned=False)       # host-masked path
        assert hb.grid == ref.grid and hb.B == ref.B == 3
        assert len(hb.contexts) == len(ref.contexts)
        for a, r in zip(hb.contexts, ref.contexts):
            assert a.gridded == r.gridded and a.y_batched == r.y_batched and a.mono == r.mono
            ax, rx = (a.x if a.gridded else (a.x,)), (r.x if r.gridded else (r.x,))
            assert all(torch.equal(u, v) for u, v in zip(ax, rx))
            ay = a.y
            if r.mask is not None:       # raw path keeps NaN where the host path has (0, mask = 0)
                nan = torch.isnan(ay).any(dim=1, keepdim=True)
                assert torch.equal(nan, r.mask == 0)
                ay = torch.nan_to_num(ay, nan=0.0) * r.mask
                assert torch.equal(ay, r.y * r.mask)
            else:
                assert not torch.isnan(ay).any() and torch.equal(ay, r.y)
        assert torch.equal(hb.xt, ref.xt) and torch.equal(hb.yt, ref.yt) and torch.equal(hb.aux_t, ref.aux_t)
    # static sets: ONE device-resident copy reused by every batch; the per-date sets live in the ring
    assert len(st._static) == 2
    hb2 = st.build(_ragged(static, 500))
    assert hb2.contexts[1] is hb.contexts[1] and hb2.contexts[2] is hb.contexts[2]
    assert hb2.contexts[0].y is not hb.contexts[0].y
    assert len(st.host_ms) == 5


def test_stager_rejects_unequal_target_counts_and_falls_back_on_processed_tasks(static):
    m = small_model("fp32")
    st = BatchStager(m.engine)
    ts = [make_task(static, 1, n_stations=200), make_task(static, 2, n_stations=150)]
    with pytest.raises(ValueError, match="same number of targets"):
        st.build(ts)
    assert not st.fast_path_ok([ConvNP.modify_task(ts[0])])      # already batched / masked: generic path
    g = make_task(static, 3, grid_targets=True)
    assert not st.fast_path_ok([g])                                # on-grid targets: generic path


def test_stager_single_task_batches_cache_static_sets_from_second_sighting(static):
    m = small_model("fp32")
    st = BatchStager(m.engine)
    a = st.build([make_task(static, 11)])
    assert a.contexts[2].y_batched and len(st._static) == 0       # first sighting: could be a per-date field
    b = st.build([make_task(static, 12)])
    assert not b.contexts[2].y_batched and len(st._static) == 2
    c = st.build([make_task(static, 13)])
    assert c.contexts[2] is b.contexts[2]


def test_nan_targets_are_dropped_like_concat_tasks(static):
    m = small_model("fp32")
    st = BatchStager(m.engine)
    ts = [make_task(static, 21), make_task(static, 22)]
    for t in ts:
        t["Y_t"][0] = t["Y_t"][0].copy()
        t["Y_t"][0][0, 3] = np.nan
    hb = st.build(ts)
    ref = m.stage_task(concat_tasks(ts), pinned=False)
    assert hb.xt.shape == (2, 2, 39) and torch.equal(hb.xt, ref.xt) and torch.equal(hb.aux_t, ref.aux_t)
