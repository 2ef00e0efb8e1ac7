"""-m gpu, needs 2 GPUs (skipped otherwise; run with ``gpurun --gpus 2``): data-parallel training over NCCL.

SURVEY 8(e): W ranks stepping B tasks each, with the two bucketed all-reduces of Engine.backward, give the gradients of
one rank stepping the W*B batch; and the CUDA-graph replay of the data-parallel step (NCCL captured) follows the eager
data-parallel step."""
import os
import socket
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out, precision):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from deepsensornz_b200 import concat_tasks
    from deepsensornz_b200.dist import enable_data_parallel, shard_tasks
    from deepsensornz_b200.graph import GraphedTrainStep
    from deepsensornz_b200.synthetic import make_static, make_task
    from tests.util import small_model
    static = make_static(seed=7, n_hi=200)
    tasks = [make_task(static, 6000 + i) for i in range(8)]
    m = small_model(precision, seed=5)
    enable_data_parallel(m)
    mine = shard_tasks(tasks, rank, world)
    batch = m._to_device(concat_tasks(mine))
    loss = m.loss_fn(batch, normalise=True)
    loss.backward()
    grads = {n: p.grad.detach().cpu().clone() for n, p in m.model.named_parameters() if p.grad is not None}
    res = dict(grads=grads, loss=float(loss))
    if rank == 0:      # the single-rank reference on the union batch, same initial weights
        m1 = small_model(precision, seed=5)
        l1 = m1.loss_fn(concat_tasks(tasks), normalise=True)
        l1.backward()
        res["ref"] = {n: p.grad.detach().cpu().clone() for n, p in m1.model.named_parameters() if p.grad is not None}
        res["ref_loss"] = float(l1)
    # eager data-parallel steps vs graph replays of the same steps (NCCL all-reduces captured)
    seqs = {}
    for mode in ("eager", "graph"):
        mm = small_model("bf16", seed=9)
        enable_data_parallel(mm)
        opt = torch.optim.AdamW(mm.model.parameters(), lr=1e-3, fused=True, capturable=True)
        dev = [mm._to_device(concat_tasks(mine[i:i + 2])) for i in (0, 2)]

        def eager(b):
            opt.zero_grad(set_to_none=True)
            l = mm.loss_fn(b, normalise=True)
            l.backward()
            opt.step()
            return float(l.detach())

        seq = [eager(dev[0])]
        gs = GraphedTrainStep(mm, opt, dev[0], warm=True) if mode == "graph" else None
        for k in range(1, 6):
            seq.append(float(gs.step(dev[k % 2])) if gs is not None else eager(dev[k % 2]))
        seqs[mode] = seq
        del gs
        torch.cuda.synchronize()
    res["seqs"] = seqs
    torch.save(res, os.path.join(out, f"rank{rank}.pt"))
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_rank_nccl_gradients_equal_one_rank_union_batch(precision):
    import torch.multiprocessing as mp
    out = tempfile.mkdtemp(prefix="cnp_dp_")
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out, precision)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0, f"rank process exited with {p.exitcode}"
    r0, r1 = (torch.load(os.path.join(out, f"rank{r}.pt")) for r in range(2))
    tol = 1e-4 if precision == "fp32" else 2e-2       # fp32: reduction order only; bf16: per-batch rounding of activations
    for n, ref in r0["ref"].items():
        a, b = r0["grads"][n].double(), r1["grads"][n].double()
        assert torch.equal(a, b), n                    # both ranks hold the same averaged gradient
        err = float((a - ref.double()).norm() / ref.double().norm().clamp(min=1e-30))
        assert err < tol, (n, err)
    assert abs(0.5 * (r0["loss"] + r1["loss"]) - r0["ref_loss"]) < tol * abs(r0["ref_loss"])
    for r in (r0, r1):
        for a, b in zip(r["seqs"]["eager"], r["seqs"]["graph"]):
            assert abs(a - b) <= 2e-3 * abs(a), r["seqs"]
