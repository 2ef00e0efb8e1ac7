"""CPU tests of the host-side helpers of ``ConvNP.predict`` (deepsensornz_b200/predict.py): the affine probe of the data
processor, the padding of off-grid context sets to a common size, and the batch signature a captured forward is keyed on."""
import numpy as np
import torch

from deepsensornz_b200.discretisation import GridSpec
from deepsensornz_b200.engine import DeviceContext, HostBatch
from deepsensornz_b200.predict import _affine_of, _inference_signature, _pad_offgrid


class _DP:
    def __init__(self, fn):
        self.fn = fn

    def map_array(self, data, var_ID, method=None, unnorm=False, add_offset=True):
        return self.fn(np.asarray(data), add_offset)


def test_affine_probe_recovers_scale_and_offset():
    dp = _DP(lambda d, off: d * 7.25 + (281.5 if off else 0.0))
    assert _affine_of(dp, "t", True) == (7.25, 281.5)
    assert _affine_of(dp, "t", False) == (7.25, 0.0)
    minmax = _DP(lambda d, off: (d + 1.0) / 2.0 * (40.0 - (-10.0)) + (-10.0 if off else 0.0))
    a, b = _affine_of(minmax, "t", True)
    assert abs(a - 25.0) < 1e-6 and abs(b - 15.0) < 1e-6


def test_affine_probe_rejects_non_affine_and_failing_processors():
    assert _affine_of(_DP(lambda d, off: np.exp(d)), "t", True) is None
    assert _affine_of(_DP(lambda d, off: d[:2]), "t", True) is None          # wrong shape
    assert _affine_of(_DP(lambda d, off: d * np.nan), "t", True) is None

    def boom(d, off):
        raise KeyError("unknown variable")
    assert _affine_of(_DP(boom), "t", True) is None


def _host_batch(n_stations, with_nan=False, mask=None):
    g = GridSpec(0.0, 8, 0.0, 8, 0.125)
    x = torch.rand(1, 2, n_stations)
    y = torch.randn(1, 1, n_stations)
    if with_nan:
        y[0, 0, 1] = float("nan")
    grid_ctx = DeviceContext(True, (torch.linspace(0, 1, 5)[None], torch.linspace(0, 1, 6)[None]), torch.randn(1, 2, 5, 6),
                             None, (1, 1), False, (np.linspace(0, 1, 5)[None], np.linspace(0, 1, 6)[None]))
    off = DeviceContext(False, x, y, mask)
    xt = (torch.linspace(0, 1, 9), torch.linspace(0, 1, 7))
    return HostBatch([grid_ctx, off], xt, None, None, g, 1, None)


def test_pad_offgrid_masks_the_padding_and_the_nans():
    hb = _pad_offgrid(_host_batch(21, with_nan=True), multiple=16)
    c = hb.contexts[1]
    assert c.x.shape == (1, 2, 32) and c.y.shape == (1, 1, 32) and c.mask.shape == (1, 1, 32)
    assert float(c.mask[0, 0, :21].sum()) == 20 and float(c.mask[0, 0, 1]) == 0      # the NaN point is masked out ...
    assert float(c.mask[0, 0, 21:].sum()) == 0 and float(c.y[0, 0, 21:].abs().sum()) == 0
    assert torch.isfinite(c.y).all()                                                    # ... and its value zeroed
    assert hb.contexts[0].gridded and hb.contexts[0].y.shape == (1, 2, 5, 6)           # gridded sets are untouched
    # an existing mask is kept and extended
    m = torch.ones(1, 1, 21)
    m[0, 0, 5] = 0
    c2 = _pad_offgrid(_host_batch(21, mask=m), multiple=16).contexts[1]
    assert float(c2.mask[0, 0, 5]) == 0 and float(c2.mask.sum()) == 20
    # already a multiple: same size, mask of ones
    c3 = _pad_offgrid(_host_batch(32), multiple=16).contexts[1]
    assert c3.x.shape[-1] == 32 and float(c3.mask.sum()) == 32


def test_signature_groups_padded_batches_and_separates_grids():
    s21 = _inference_signature(_pad_offgrid(_host_batch(21)))
    s30 = _inference_signature(_pad_offgrid(_host_batch(30)))
    s40 = _inference_signature(_pad_offgrid(_host_batch(40)))
    assert s21 == s30 and s21 != s40            # 21 and 30 stations pad to 32, 40 pads to 48
    hb = _pad_offgrid(_host_batch(21))
    hb.contexts[0].x_host = (np.linspace(0, 1, 5)[None] + 1e-3, np.linspace(0, 1, 6)[None])   # other grid coordinates
    assert _inference_signature(hb) != s21


def test_result_pool_recycles_only_unreferenced_arrays(monkeypatch):
    """predict's result arrays come from a pool (first-touch page faults bound predict end to end): an array is handed
    out again only when nothing -- no view, no view of a view, no tensor made from it -- references it any more."""
    import numpy as np
    import torch
    from deepsensornz_b200.predict import _ResultPool
    p = _ResultPool()
    a, b = p.take((4, 50, 60))[0], p.take((4, 50, 60))[0]
    pa, pb = a.ctypes.data, b.ctypes.data
    assert pa != pb and a.shape == (4, 50, 60) and a.dtype == np.float32
    v = a[1:3, ::2]                       # a strided view of a view keeps the owner alive
    del a
    c = p.take((4, 50, 60))[0]
    assert c.ctypes.data not in (pa, pb)
    del v
    d = p.take((2, 50, 60))[0]            # a smaller request fits the idle owner
    assert d.ctypes.data == pa and p.hits == 1
    t = torch.from_numpy(d)
    del d
    assert p.take((2, 50, 60))[0].ctypes.data != pa
    del t
    assert p.take((4, 50, 60))[0].ctypes.data == pa
    # cap: requests over the cap bypass the pool, a cap of 0 disables it
    monkeypatch.setenv("CONVNP_B200_RESULT_POOL_MB", "0")
    q = _ResultPool()
    x = q.take((8, 8))[0]
    px = x.ctypes.data
    del x
    q.take((8, 8))
    assert q.hits == 0 and q.misses == 0 and px


def test_static_device_copies_persist_across_predict_calls_only_while_unchanged():
    """The fingerprint that lets predict keep the aux-at-target tensor and the static context sets on the device between
    calls notices an in-place rewrite and a different buffer."""
    import numpy as np
    from deepsensornz_b200.predict import _fingerprint, _persistent_ctx_cache
    a = np.arange(10000, dtype=np.float32).reshape(100, 100)
    fp = _fingerprint(a)
    assert fp == _fingerprint(a) and _fingerprint(a[:, ::2]) is None          # non-contiguous: never kept
    a[0, 0] = 123.0                                                            # sample 0 is always part of the sum
    assert _fingerprint(a) != fp
    assert _fingerprint(a.copy())[0] != _fingerprint(a)[0]

    class M:                                                                   # stand-in for the model object
        pass
    m = M()
    static = np.ones((1, 8, 8), np.float32)
    mk = lambda i: {"X_c": [(np.arange(8.), np.arange(8.)), (np.arange(8.), np.arange(8.))],
                    "Y_c": [np.full((1, 8, 8), float(i), np.float32), static]}
    c1 = _persistent_ctx_cache(m, [mk(0), mk(1)])
    c1["x"] = 1
    assert _persistent_ctx_cache(m, [mk(2), mk(3)]) is c1                      # per-date grids differ: still the same cache
    static[0, 0, 0] = 5.0
    assert _persistent_ctx_cache(m, [mk(0), mk(1)]) is not c1                  # the static set changed: reset
