"""-m gpu: the fused SetConv encoder (csrc/enc_fused.cu) -- all context sets of a task in one launch -- against the CPU
oracle (upstream PrependDensityChannel + SetConv + DivideByFirstChannel), against the per-set kernels it
replaces, and its blocked bf16 output against the layout conversion of its fp32 output."""
import ctypes as C

import numpy as np
import pytest
import torch

from deepsensornz_b200 import Task, _cabi, concat_tasks
from deepsensornz_b200.synthetic import make_static, make_task
from oracle import convnp_oracle as O
from oracle.task_tensors import task_tensors
from tests.util import cpu_params, rel_err, small_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def static():
    return make_static(seed=7, n_hi=200)


def _oracle_enc(m, tasks):
    contexts, xt, yt, aux = task_tensors(tasks)
    (s1, n1), (s2, n2), res = O.discretise([c[0] for c in contexts] + [xt], m.config.points_per_unit, 0.1, 8)
    return O.encoder(cpu_params(m), contexts, O.grid_points(s1, n1, res), O.grid_points(s2, n2, res))


def _calls(m, monkeypatch):
    calls = []
    orig = m.engine._call
    monkeypatch.setattr(m.engine, "_call", lambda name, *a, **k: (calls.append(name), orig(name, *a, **k))[1])
    return calls


@pytest.mark.parametrize("nb,c0", [(1, 3), (3, 3), (2, 8)])
def test_fused_encoder_matches_oracle_and_per_set_kernels(static, monkeypatch, nb, c0):
    m = small_model("fp32", dim_yc=(c0, 6, 1, 1))
    tasks = [make_task(static, 100 + i, c0_channels=c0, n_stations=200 - 10 * i,
                       context_frac=(160 - 10 * i) / (200 - 10 * i)) for i in range(nb)]       # ragged context sets
    task = concat_tasks(tasks) if nb > 1 else tasks[0]
    calls = _calls(m, monkeypatch)
    batch = m._to_device(task)
    enc = m.engine.encode(batch).clone()
    assert [calls.count(n) for n in ("cnp_encode_hpass", "cnp_encode_vpass", "cnp_encode_fused")] == [1, 1, 1]
    assert "cnp_setconv_enc_grid_fwd" not in calls
    ref = _oracle_enc(m, tasks)
    assert enc.shape == ref.shape
    for c in range(ref.shape[1]):               # per channel, so that small channels are held to the same bound
        assert rel_err(enc[:, c], ref[:, c]) < 5e-5, c
    assert rel_err(enc, ref) < 1e-5
    monkeypatch.setenv("CNP_NO_ENC_FUSED", "1")
    calls.clear()
    old = m.engine.encode(batch)
    assert "cnp_encode_fused" not in calls and "cnp_setconv_enc_grid_fwd" in calls
    assert rel_err(enc, old) < 2e-6             # same terms, different summation order


def test_raw_nan_path_equals_host_masked_path_bitwise(static):
    m = small_model("fp32")
    t = make_task(static, 5)
    e1 = m.engine.encode(m._to_device(t)).clone()                      # NaNs travel, masks derived on the device
    e2 = m.engine.encode(m._to_device(m.modify_task(t))).clone()       # host Masked path
    assert torch.equal(e1, e2)


@pytest.mark.parametrize("nb", [1, 3])
def test_blocked_bf16_output_equals_conversion_of_fp32_output(static, nb):
    m = small_model("bf16")
    eng = m.engine
    tasks = [make_task(static, 300 + i) for i in range(nb)]
    batch = m._to_device(concat_tasks(tasks) if nb > 1 else tasks[0])
    g = batch.grid
    blk = eng.encode_blocked(batch)
    assert blk is not None and blk.CB == 2
    got = blk.t[:nb * blk.bstride].view(nb, 2, g.n1 + 4, g.n2 + 4, 8).clone()
    enc = eng.encode(batch)
    ref_blk = eng._blk("ref_aug", nb, 2, g.n1, g.n2)
    _cabi.call("cnp_blk_from_nchw_f32_ones", enc.data_ptr(), enc.stride(0), nb, m.config.in_channels, g.n1, g.n2,
               C.byref(ref_blk.view()), 2, 0, torch.cuda.current_stream().cuda_stream)
    want = ref_blk.t[:nb * ref_blk.bstride].view(nb, 2, g.n1 + 4, g.n2 + 4, 8)
    assert torch.equal(got, want)
    assert float(got[:, :, :2].abs().max()) == 0 and float(got[:, :, :, -2:].abs().max()) == 0     # pad untouched
    assert float((got[:, 1, 2:-2, 2:-2, 7].float() - 1).abs().max()) == 0                             # constant-1 channel


@pytest.mark.parametrize("case", ["no_context_stations", "all_sea", "nan_in_static_field"])
def test_fused_encoder_edge_cases(static, case):
    t = make_task(static, 3100, n_stations=30, context_frac=0.6)
    t = Task({k: (list(v) if isinstance(v, list) else v) for k, v in t.items()})
    if case == "no_context_stations":
        t["X_c"][3], t["Y_c"][3] = t["X_c"][3][:, :0], t["Y_c"][3][:, :0]
    elif case == "all_sea":
        y = t["Y_c"][0].copy()
        y[:] = np.nan
        t["Y_c"][0] = y
    else:
        y = t["Y_c"][1].copy()
        y[2, 10:40, 50:90] = np.nan
        t["Y_c"][1] = y
    m = small_model("fp32")
    enc = m.engine.encode(m._to_device(t))
    ref = _oracle_enc(m, [t])
    assert torch.isfinite(enc).all()
    assert rel_err(enc, ref) < 1e-5
