"""-m gpu: train_epoch through the device-resident task pipeline (staging.BatchStager + worker thread), CUDA-graph
replay across batch signatures, and the reference's GPU flow ``set_gpu_default_device()`` -> ``train_epoch`` ->
``predict`` (nzdownscale/downscaler/train.py:48, :388-394; validate_ERA.py:88-92)."""
import numpy as np
import pytest
import torch

from deepsensornz_b200 import _cabi, concat_tasks, set_gpu_default_device, train_epoch
from deepsensornz_b200.graph import GraphedTrainStep
from deepsensornz_b200.synthetic import make_static, make_task
from tests.util import small_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def static():
    return make_static(seed=7, n_hi=200, with_aux_hi=True)


def _manual_epoch(m, opt, groups):
    out = []
    for g in groups:
        opt.zero_grad()
        loss = m.loss_fn(concat_tasks(g), normalise=True)      # host-masked path, synchronous upload
        loss.backward()
        opt.step()
        out.append(float(loss))
    return out


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_train_epoch_pipeline_equals_concat_then_loss_fn(static, precision, monkeypatch):
    """Same tasks, same order, same initial weights: train_epoch (raw tasks -> pinned ring -> copy stream, NaNs masked
    on the device, static sets resident) and the reference's form (concat_tasks -> loss_fn) give the same losses."""
    tasks = [make_task(static, 900 + i) for i in range(12)]
    monkeypatch.setattr(np.random, "permutation", lambda n: np.arange(n))
    losses = {}
    for mode in ("pipeline", "manual"):
        m = small_model(precision, seed=11)
        opt = torch.optim.AdamW(m.model.parameters(), lr=1e-3)
        if mode == "pipeline":
            losses[mode] = train_epoch(m, tasks, batch_size=4, opt=opt)
            st = m._stager
            assert len(st._static) == 2 and len(st.host_ms) == 3      # both static sets resident, three batches built
        else:
            losses[mode] = _manual_epoch(m, opt, [tasks[i:i + 4] for i in range(0, 12, 4)])
    tol = 1e-5 if precision == "fp32" else 2e-3       # fp32 atomics reorder sums; bf16 adds rounding of the updates
    for a, b in zip(losses["pipeline"], losses["manual"]):
        assert abs(a - b) <= tol * abs(b), losses


def test_reference_gpu_flow_after_set_gpu_default_device(static):
    """train.py:48 calls set_gpu_default_device() before anything else: host-side staging must keep working when
    torch's default device is CUDA (pinned buffers, loss slots, predict's result buffers)."""
    try:
        set_gpu_default_device()
        m = small_model("bf16", seed=2)
        assert next(m.model.parameters()).is_cuda
        tasks = [make_task(static, 40 + i) for i in range(8)]
        opt = torch.optim.AdamW(m.model.parameters(), lr=1e-3)
        l1 = train_epoch(m, tasks, batch_size=4, opt=opt)
        l2 = train_epoch(m, tasks, batch_size=4, opt=opt, use_graph=True)
        l3 = train_epoch(m, tasks, batch_size=4, opt=opt, use_graph=True)
        assert len(l1) == len(l2) == len(l3) == 2 and np.all(np.isfinite(l1 + l2 + l3))
        assert float(m.loss_fn(tasks[0], normalise=True)) == float(m.loss_fn(tasks[0], normalise=True))
        gt = [make_task(static, 60 + i, all_context=True) for i in range(3)]
        pred = m.predict(gt, X_t=(static.x_hi, static.x_hi), X_t_is_normalised=True, aux_at_targets_override=static.aux_hi)
        mean = np.asarray(pred[list(pred.keys())[0]]["mean"])
        assert mean.shape == (3, 200, 200) and np.isfinite(mean).all()
    finally:
        torch.set_default_device("cpu")


def test_graph_captured_after_an_eager_forward_still_repacks_weights(static):
    """a validation forward right before the capture leaves the packed bf16 weights current; the captured
    graph must contain the packing kernels anyway, or every replay would use the weights of capture time."""
    groups = [[make_task(static, 700 + 4 * k + i) for i in range(4)] for k in range(3)]
    out = {}
    for mode in ("eager", "graph"):
        m = small_model("bf16", seed=3)
        opt = torch.optim.AdamW(m.model.parameters(), lr=2e-3, fused=True, capturable=True)
        dev = [m._to_device(concat_tasks(g)) for g in groups]
        seq = []

        def eager(b):
            opt.zero_grad(set_to_none=True)
            loss = m.loss_fn(b, normalise=True)
            loss.backward()
            opt.step()
            return float(loss)

        seq.append(eager(dev[0]))
        with torch.no_grad():
            m.loss_fn(dev[1], normalise=True)          # validation forward: re-packs, versions now current
        if mode == "graph":
            gs = GraphedTrainStep(m, opt, dev[0], warm=True)
            assert m.engine.packs_recorded > 0
            for k in range(1, 7):
                seq.append(float(gs.step(dev[k % 3])))
        else:
            for k in range(1, 7):
                seq.append(eager(dev[k % 3]))
        out[mode] = seq
    assert out["graph"][-1] < out["graph"][0]
    for a, b in zip(out["eager"], out["graph"]):
        assert abs(a - b) <= 2e-3 * abs(a), out


def test_graph_replay_interleaved_with_eager_steps_and_plain_adamw(static):
    """with a non-capturable optimiser the replayed graph writes into the gradient tensors of capture time;
    an eager step on another batch signature rebinds p.grad, so step() must bind the captured tensors back."""
    sigA = [[make_task(static, 100 + 4 * k + i) for i in range(4)] for k in range(4)]                 # 40 targets
    sigB = [[make_task(static, 300 + 4 * k + i, n_stations=150) for i in range(4)] for k in range(4)]   # 30 targets
    out = {}
    for mode in ("eager", "graph"):
        m = small_model("bf16", seed=4)
        opt = torch.optim.AdamW(m.model.parameters(), lr=2e-3)            # the reference's plain AdamW (train.py:354)
        A = [m._to_device(concat_tasks(g)) for g in sigA]
        Bb = [m._to_device(concat_tasks(g)) for g in sigB]

        def eager(b):
            opt.zero_grad()
            loss = m.loss_fn(b, normalise=True)
            loss.backward()
            opt.step()
            return float(loss)

        seq = [eager(A[0])]
        gs = GraphedTrainStep(m, opt, A[0], warm=True) if mode == "graph" else None
        for k in range(1, 4):
            seq.append(float(gs.step(A[k])) if gs is not None else eager(A[k]))
            seq.append(eager(Bb[k]))                  # other signature: eager, rebinds p.grad
        seq.append(float(gs.step(A[0])) if gs is not None else eager(A[0]))
        out[mode] = seq
    for a, b in zip(out["eager"], out["graph"]):
        assert abs(a - b) <= 2e-3 * abs(a), out


def test_second_forward_before_backward_is_refused(static):
    m = small_model("fp32")
    t1, t2 = make_task(static, 5), make_task(static, 6)
    l1 = m.loss_fn(t1, normalise=True)
    m.loss_fn(t2, normalise=True)
    with pytest.raises(_cabi.CnpError, match="no longer the engine's latest"):
        l1.backward()
    l3 = m.loss_fn(t1, normalise=True)     # and the engine is still usable
    l3.backward()
    assert all(torch.isfinite(p.grad).all() for p in m.model.parameters() if p.requires_grad)


def test_workspaces_of_old_batch_shapes_are_released(static):
    """the reference steps one batch shape per station-count group; the engine keeps the workspaces of the
    last few shapes only (a shape a CUDA graph replays from is never released)."""
    m = small_model("bf16", seed=6)
    eng = m.engine
    eng.MAX_SIGNATURES = 2
    sizes = []
    for k, nst in enumerate((200, 190, 180, 170, 160)):
        t = concat_tasks([make_task(static, 9000 + 10 * k + i, n_stations=nst) for i in range(2)])
        loss = m.loss_fn(t, normalise=True)
        loss.backward()
        sizes.append(len(eng._ws))
        assert len(eng._sig_lru) <= 2
    assert sizes[-1] <= sizes[1] + 2          # does not grow with the number of shapes seen
    # a shape that came back after eviction still gives the same loss
    t = concat_tasks([make_task(static, 9000 + i, n_stations=200) for i in range(2)])
    with torch.no_grad():
        a = float(m.loss_fn(t, normalise=True))
        b = float(m.loss_fn(t, normalise=True))
    assert a == b and np.isfinite(a)


def test_training_is_run_to_run_deterministic_and_graph_equals_eager(static):
    """Every reduction of the bf16 training step runs in a fixed order (K-split weight gradients and their bias sums
    through a workspace, MLP / final-1x1 gradients per block then in block order, the float64 log-pdf per task): two
    runs give bit-identical weights, and the CUDA-graph replay equals the eager step bit for bit."""
    groups = [[make_task(static, 1200 + 4 * k + i) for i in range(4)] for k in range(3)]

    def run(mode):
        m = small_model("bf16", seed=12)
        opt = torch.optim.AdamW(m.model.parameters(), lr=1e-3, fused=True, capturable=True)
        dev = [m._to_device(concat_tasks(g)) for g in groups]
        losses = []

        def eager(b):
            opt.zero_grad(set_to_none=True)
            loss = m.loss_fn(b, normalise=True)
            loss.backward()
            opt.step()
            return float(loss.detach())

        losses.append(eager(dev[0]))
        gs = GraphedTrainStep(m, opt, dev[0], warm=True) if mode == "graph" else None
        for k in range(1, 5):
            losses.append(float(gs.step(dev[k % 3])) if gs is not None else eager(dev[k % 3]))
        torch.cuda.synchronize()
        return losses, torch.cat([p.detach().flatten() for p in m.model.parameters()]).clone()

    l1, w1 = run("eager")
    l2, w2 = run("eager")
    l3, w3 = run("graph")
    assert l1 == l2 and torch.equal(w1, w2)
    assert l1 == l3 and torch.equal(w1, w3)
