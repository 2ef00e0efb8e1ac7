"""CPU tests of the ``TaskLoader`` mirror (deepsensornz_b200/loader.py) on the call forms nzdownscale uses:
construction train.py:160-166, ``task_loader(date, context_sampling, target_sampling)`` train.py:315 /
validate_ERA.py:79, pickling train.py:174-177, the ``TaskLoader_SampleStations`` override pattern train.py:525-637,
attribute swaps validate_ERA.py:117-127, and the ``ConvNP`` defaults derived from the loader (internal density,
encoder scales; validation_precip.ipynb:185)."""
import pickle

import numpy as np
import pandas as pd
import pytest

from deepsensornz_b200 import Task
from deepsensornz_b200.data import GridVar
from deepsensornz_b200.loader import InvalidSamplingStrategyError, TaskLoader

DATES = pd.date_range("2016-01-01", periods=4, freq="D")


def _grid(n1, n2, names, seed, timed=True, lo=0.0, hi=1.0):
    rng = np.random.default_rng(seed)
    x1 = np.linspace(lo, hi, n1)
    x2 = np.linspace(lo, hi, n2)
    shape = (len(DATES), n1, n2) if timed else (n1, n2)
    return GridVar({k: rng.normal(size=shape).astype(np.float32) for k in names}, x1, x2, DATES if timed else None)


def _stations(n, seed, nan_at=()):
    rng = np.random.default_rng(seed)
    x1, x2 = rng.uniform(0.05, 0.95, n), rng.uniform(0.05, 0.95, n)
    rows = [(d, a, b) for d in DATES for a, b in zip(x1, x2)]
    idx = pd.MultiIndex.from_tuples(rows, names=["time", "x1", "x2"])
    df = pd.DataFrame({"dry_bulb": rng.normal(size=len(rows)).astype(np.float32)}, index=idx)
    for k in nan_at:
        df.iloc[k] = np.nan
    return df


@pytest.fixture()
def loader():
    era = _grid(12, 15, ["t2m"], 1)
    elev = _grid(30, 40, ["elevation", "tpi"], 2, timed=False)
    aux = _grid(60, 80, ["elevation", "tpi_a", "tpi_b"], 3, timed=False)
    return TaskLoader(context=[era, elev, _stations(20, 4)], target=_stations(20, 4), aux_at_targets=aux)


def test_metadata_matches_variables(loader):
    assert loader.context_dims == (1, 2, 1)
    assert loader.target_dims == (1,)
    assert loader.aux_at_target_dims == 3
    assert loader.context_var_IDs == (("t2m",), ("elevation", "tpi"), ("dry_bulb",))
    assert loader.target_var_IDs == (("dry_bulb",),)
    assert loader.aux_at_target_var_IDs == ("elevation", "tpi_a", "tpi_b")
    assert "3 context sets" in str(loader)
    assert loader.context_delta_t == (0, 0, 0) and loader.target_delta_t == (0,)


def test_defaults_used_by_convnp(loader):
    """ppu = ceil(1 / finest gridded spacing of context+target); encoder scale = half the spacing for gridded sets and
    0.5/ppu for off-grid ones."""
    res_fine = np.mean([1 / 29, 1 / 39])
    assert loader.gen_ppu() == int(np.ceil(1 / res_fine))
    sc = loader.gen_encoder_scales(250)
    assert sc[0] == pytest.approx(0.5 * np.mean([1 / 11, 1 / 14]))
    assert sc[1] == pytest.approx(0.5 * res_fine)
    assert sc[2] == pytest.approx(0.5 / 250)
    only_stations = TaskLoader(context=_stations(5, 1), target=_stations(5, 1))
    with pytest.raises(ValueError):
        only_stations.gen_ppu()


def test_task_layout_all_sampling(loader):
    task = loader(DATES[1], context_sampling="all", target_sampling="all")
    assert isinstance(task, Task) and task["time"] == DATES[1]
    # gridded context -> tuple of [1,N] coordinate rows + [C,N1,N2] values of that date
    (x1, x2), y0 = task["X_c"][0], task["Y_c"][0]
    assert x1.shape == (1, 12) and x2.shape == (1, 15) and x1.dtype == np.float32
    assert np.array_equal(y0, loader.context[0].data_vars["t2m"][1][None])
    assert task["Y_c"][1].shape == (2, 30, 40)
    # off-grid context -> [2,N] / [C,N]
    assert task["X_c"][2].shape == (2, 20) and task["Y_c"][2].shape == (1, 20)
    assert task["X_t"][0].shape == (2, 20) and task["Y_t"][0].shape == (1, 20)
    day = loader.target[0].xs(DATES[1], level="time")
    assert np.allclose(task["Y_t"][0][0], day["dry_bulb"].values)
    assert np.allclose(task["X_t"][0][0], day.index.get_level_values("x1").values.astype(np.float32))
    # aux-at-targets: nearest high-res cell, [C_aux, N]
    aux = loader.aux_at_targets
    i = np.abs(aux.x1[None] - task["X_t"][0][0][:, None].astype(np.float64)).argmin(1)
    j = np.abs(aux.x2[None] - task["X_t"][0][1][:, None].astype(np.float64)).argmin(1)
    assert task["Y_t_aux"].shape == (3, 20) and task["Y_t_aux"].dtype == np.float32
    assert np.array_equal(task["Y_t_aux"], aux.stack()[:, i, j])


def test_list_of_dates_and_no_targets(loader):
    tasks = loader(list(DATES[:3]), "all")
    assert len(tasks) == 3 and [t["time"] for t in tasks] == list(DATES[:3])
    assert tasks[0]["X_t"] == [] and "Y_t_aux" not in tasks[0]
    assert len(loader(DATES[:2], "all", "all")) == 2           # DatetimeIndex


def test_int_and_float_sampling_are_seeded_subsets(loader):
    a = loader(DATES[0], context_sampling=["all", "all", 7], target_sampling="all", seed_override=5)
    b = loader(DATES[0], context_sampling=["all", "all", 7], target_sampling="all", seed_override=5)
    c = loader(DATES[0], context_sampling=["all", "all", 7], target_sampling="all", seed_override=6)
    assert a["X_c"][2].shape == (2, 7)
    assert np.array_equal(a["X_c"][2], b["X_c"][2]) and not np.array_equal(a["X_c"][2], c["X_c"][2])
    full = loader(DATES[0], "all")["X_c"][2]
    assert all(any(np.array_equal(col, f) for f in full.T) for col in a["X_c"][2].T)     # subset of the stations
    assert len({tuple(col) for col in a["X_c"][2].T}) == 7                                  # without replacement
    frac = loader(DATES[0], context_sampling=["all", "all", 0.5], seed_override=1)
    assert frac["X_c"][2].shape == (2, 10)
    # gridded variable sampled off-grid: [2,N] coordinates drawn from the grid axes, values = the cells hit
    g = loader(DATES[2], context_sampling=[9, "all", "all"], seed_override=3)
    X, Y = g["X_c"][0], g["Y_c"][0]
    assert X.shape == (2, 9) and Y.shape == (1, 9)
    era = loader.context[0]
    i = np.abs(era.x1[None] - X[0][:, None].astype(np.float64)).argmin(1)
    j = np.abs(era.x2[None] - X[1][:, None].astype(np.float64)).argmin(1)
    assert np.array_equal(Y[0], era.data_vars["t2m"][2][i, j])
    # datewise-deterministic seeding: same date -> same draw, other date -> other draw
    d0 = loader(DATES[0], ["all", "all", 5], datewise_deterministic=True)
    d0b = loader(DATES[0], ["all", "all", 5], datewise_deterministic=True)
    assert np.array_equal(d0["X_c"][2], d0b["X_c"][2])
    with pytest.raises(InvalidSamplingStrategyError):
        loader(DATES[0], context_sampling="every-other")


def test_nan_observations_are_dropped():
    st = _stations(10, 9, nan_at=(3, 14))           # one NaN on day 0, one on day 1
    ld = TaskLoader(context=[_grid(8, 8, ["t2m"], 1), st], target=st)
    t0, t2 = ld(DATES[0], "all", "all"), ld(DATES[2], "all", "all")
    assert t0["X_c"][1].shape == (2, 9) and t0["Y_t"][0].shape == (1, 9)
    assert t2["X_c"][1].shape == (2, 10)
    assert np.isfinite(t0["Y_c"][1]).all()


def test_on_grid_aux_lookup_and_nearest_ties():
    aux = GridVar({"e": np.arange(12, dtype=np.float32).reshape(3, 4)}, [0.0, 0.5, 1.0], [0.0, 1.0, 2.0, 3.0])
    ld = TaskLoader(context=_grid(4, 4, ["t2m"], 1), target=_stations(3, 1), aux_at_targets=aux)
    out = ld.sample_offgrid_aux((np.array([[0.1, 0.9]]), np.array([[0.2, 2.9, 1.4]])), aux)
    assert out.shape == (1, 2, 3)
    assert np.array_equal(out[0], np.array([[0, 3, 1], [8, 11, 9]], dtype=np.float32))
    pts = ld.sample_offgrid_aux(np.array([[0.1, 0.9], [2.9, 0.2]]), aux)
    assert np.array_equal(pts, np.array([[3, 8]], dtype=np.float32))
    # descending coordinate axis (latitude stored north -> south, as ERA5 files are)
    aux_d = GridVar({"e": np.arange(12, dtype=np.float32).reshape(3, 4)}, [1.0, 0.5, 0.0], [0.0, 1.0, 2.0, 3.0])
    assert np.array_equal(ld.sample_offgrid_aux(np.array([[0.1, 0.9], [2.9, 0.2]]), aux_d),
                          np.array([[11, 0]], dtype=np.float32))


def test_delta_t_shifts_the_context_date():
    era = _grid(6, 6, ["t2m"], 1)
    ld = TaskLoader(context=[era, era], context_delta_t=[0, -1], target=_stations(4, 2))
    t = ld(DATES[2], "all", "all")
    assert np.array_equal(t["Y_c"][0][0], era.data_vars["t2m"][2])
    assert np.array_equal(t["Y_c"][1][0], era.data_vars["t2m"][1])
    with pytest.raises(KeyError):
        ld(DATES[0], "all")                      # the day before the first one is not in the variable


def test_pickle_round_trip(loader):
    ld2 = pickle.loads(pickle.dumps(loader))
    a, b = loader(DATES[3], "all", "all"), ld2(DATES[3], "all", "all")
    assert np.array_equal(a["Y_c"][0], b["Y_c"][0]) and np.array_equal(a["Y_t_aux"], b["Y_t_aux"])
    assert ld2.context_dims == loader.context_dims


def test_attribute_swap_between_calls(loader):
    """validate_ERA.py:117-127 replaces ``task_loader.context`` / ``.target`` on a loaded loader."""
    new_era = _grid(9, 9, ["t2m"], 17)
    loader.context = [new_era] + loader.context[1:]
    t = loader(DATES[0], "all", "all")
    assert t["Y_c"][0].shape == (1, 9, 9)
    assert loader.gen_encoder_scales(100)[0] == pytest.approx(0.5 / 8)


class _SampleStations(TaskLoader):
    """The override pattern of train.py:525-637: the station frame is a *context* set; ``sample_df`` returns the
    sampled stations as context and the complement as targets."""

    def sample_df(self, df, sampling_strat, seed=None):
        df = df.dropna(how="any")
        if isinstance(sampling_strat, float):
            sampling_strat = int(sampling_strat * df.shape[0])
        idx = np.random.default_rng(seed).choice(df.index, sampling_strat, replace=False)
        ctx, tgt = df.loc[idx], df.drop(idx)
        xy = lambda d: d.reset_index()[["x1", "x2"]].values.T.astype(self.dtype)
        return xy(ctx), ctx.values.T, xy(tgt), tgt.values.T

    def task_generation(self, date, context_sampling="all", target_sampling=None, split_frac=0.5,
                        datewise_deterministic=False, seed_override=None):
        strat = context_sampling if isinstance(context_sampling, (list, tuple)) else [context_sampling] * len(self.context)
        date = pd.Timestamp(date)
        task = {"time": date, "ops": [], "X_c": [], "Y_c": [], "X_t": [], "Y_t": []}
        for i, (var, dt, s) in enumerate(zip(self.context, self.context_delta_t, strat)):
            v = self.time_slice_variable(var, date, dt)
            if isinstance(v, (pd.DataFrame, pd.Series)):
                X_c, Y_c, X_t, Y_t = self.sample_df(v, s, seed_override)
                task["X_t"].append(X_t)
                task["Y_t"].append(Y_t)
            else:
                X_c, Y_c = self.sample_da(v, s, seed_override)
            task["X_c"].append(X_c)
            task["Y_c"].append(Y_c)
        task["Y_t_aux"] = self.sample_offgrid_aux(task["X_t"][0], self.time_slice_variable(self.aux_at_targets, date))
        return Task(task)


def test_subclass_override_pattern(loader):
    ld = _SampleStations(context=loader.context, target=loader.target, aux_at_targets=loader.aux_at_targets)
    t = ld(DATES[1], context_sampling=["all", "all", 0.75], seed_override=2)
    assert t["X_c"][2].shape == (2, 15) and t["X_t"][0].shape == (2, 5) and t["Y_t_aux"].shape == (3, 5)
    ctx = {tuple(c) for c in t["X_c"][2].T}
    tgt = {tuple(c) for c in t["X_t"][0].T}
    assert not (ctx & tgt) and len(ctx | tgt) == 20            # context and target stations are complementary


def test_tasks_feed_the_host_staging(loader):
    """A loader task goes through the same host staging as the synthetic tasks the GPU tests use (no device needed)."""
    from deepsensornz_b200 import concat_tasks
    tasks = [loader(d, "all", "all").cast_to_float32().add_batch_dim() for d in DATES[:2]]
    batch = concat_tasks(tasks)
    assert batch["Y_c"][0].shape == (2, 1, 12, 15) and batch["X_t"][0].shape == (2, 2, 20)
    assert batch["Y_t_aux"].shape == (2, 3, 20)


def test_aux_at_contexts_appends_an_offgrid_set(loader):
    """train.py:614-626: the aux variable sampled at all off-grid context locations becomes one more context set."""
    ld = TaskLoader(context=loader.context, target=loader.target, aux_at_contexts=loader.aux_at_targets)
    t = ld(DATES[0], "all", "all")
    assert len(t["X_c"]) == 4 and t["X_c"][3].shape == (2, 20) and t["Y_c"][3].shape == (3, 20)
    assert np.array_equal(t["X_c"][3], t["X_c"][2])
    grid_only = TaskLoader(context=loader.context[:2], aux_at_contexts=loader.aux_at_targets)(DATES[0], "all")
    assert grid_only["X_c"][2].shape == (2, 0) and grid_only["Y_c"][2].shape == (3, 0)


def test_static_variables_return_identical_arrays(loader):
    """A context variable without a time axis is the same field on every date: "all" sampling hands back the same array
    objects (ConvNP.predict keys its upload-once cache on identity); dated variables and sub-sampled draws do not."""
    a, b = loader(DATES[0], "all", "all"), loader(DATES[1], "all", "all")
    assert a["Y_c"][1] is b["Y_c"][1] and a["X_c"][1][0] is b["X_c"][1][0]       # static elevation set
    assert a["Y_c"][0] is not b["Y_c"][0] and not np.array_equal(a["Y_c"][0], b["Y_c"][0])   # dated ERA field
    s1 = loader(DATES[0], [5, 5, "all"], seed_override=1)
    s2 = loader(DATES[0], [5, 5, "all"], seed_override=2)
    assert s1["Y_c"][1] is not s2["Y_c"][1]
    loader.context = [loader.context[0], _grid(30, 40, ["elevation", "tpi"], 99, timed=False), loader.context[2]]
    c = loader(DATES[0], "all", "all")
    assert c["Y_c"][1] is not a["Y_c"][1] and not np.array_equal(c["Y_c"][1], a["Y_c"][1])   # swapped variable: new field
    assert "_static_cache" not in pickle.loads(pickle.dumps(loader)).__dict__
