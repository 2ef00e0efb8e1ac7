"""World-size-2 gloo tests of the data-parallel plumbing (the N>1 path of SURVEY 8(e)) on CPU.

Results travel through files in a temporary directory (a multiprocessing queue's feeder thread raced the process exit
in round 1) and the rendezvous port is one the OS just handed out."""
import os
import socket
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _init(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _run(worker, world=2, *args):
    out = tempfile.mkdtemp(prefix="cnp_dist_")
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=worker, args=(r, world, port, out) + args) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0, f"rank process exited with {p.exitcode}"
    return [torch.load(os.path.join(out, f"rank{r}.pt")) for r in range(world)]


def _worker_plumbing(rank, world, port, out):
    _init(rank, world, port)
    from deepsensornz_b200 import ConvNP
    from deepsensornz_b200.dist import allreduce_mean_, enable_data_parallel
    torch.manual_seed(100 + rank)                       # different init per rank on purpose
    m = ConvNP(dim_yc=(3, 6, 1, 1), dim_yt=1, dim_aux_t=5, internal_density=50, verbose=False)
    enable_data_parallel(m)
    w = m.model.decoder.unet.initial_linear.weight.detach().clone()
    flat = torch.full((10,), float(rank + 1))
    allreduce_mean_(flat, world)
    torch.save(dict(w=w, flat=flat, world=m.engine.world_size), os.path.join(out, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_and_allreduce_gloo():
    res = _run(_worker_plumbing)
    assert torch.equal(res[0]["w"], res[1]["w"])            # identical weights after broadcast
    assert torch.allclose(res[0]["flat"], torch.full((10,), 1.5)) and torch.allclose(res[1]["flat"], torch.full((10,), 1.5))
    assert res[0]["world"] == 2


def _worker_equivalence(rank, world, port, out):
    """Each rank: oracle gradients of the mean loss over ITS shard, then the all-reduce-mean of the flat bucket."""
    _init(rank, world, port)
    torch.set_num_threads(2)
    from deepsensornz_b200.dist import allreduce_mean_, shard_tasks
    from deepsensornz_b200.synthetic import make_static, make_task
    from oracle import convnp_oracle as O
    from oracle.task_tensors import task_tensors
    from tests.util import cpu_params, small_model
    static = make_static(seed=7, n_hi=120)
    tasks = [make_task(static, 4000 + i, n_stations=60) for i in range(4)]
    m = small_model("fp32", ppu=24)
    mine = shard_tasks(tasks, rank, world)
    assert len(mine) == 2
    P = {k: v.clone().requires_grad_(v.dim() > 0) for k, v in cpu_params(m).items()}
    ctx, xt, yt, aux = task_tensors(mine)
    O.loss_fn(P, ctx, xt, yt, aux, m.config.points_per_unit).backward()
    names = sorted(k for k, v in P.items() if v.grad is not None)
    flat = torch.cat([P[k].grad.flatten() for k in names])
    allreduce_mean_(flat, world)
    torch.save(dict(flat=flat, names=names), os.path.join(out, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradients_equal_one_rank_on_the_union_batch():
    """W ranks stepping B tasks each, gradients all-reduced to their mean, == one rank stepping the W*B
    batch (the loss is the mean over tasks).  The CPU stand-in for the model is the oracle; the same check runs on the
    CUDA path over NCCL in tests/test_dp_gpu.py."""
    res = _run(_worker_equivalence)
    assert torch.equal(res[0]["flat"], res[1]["flat"])
    from deepsensornz_b200.synthetic import make_static, make_task
    from oracle import convnp_oracle as O
    from oracle.task_tensors import task_tensors
    from tests.util import cpu_params, small_model
    static = make_static(seed=7, n_hi=120)
    tasks = [make_task(static, 4000 + i, n_stations=60) for i in range(4)]
    m = small_model("fp32", ppu=24)
    P = {k: v.clone().requires_grad_(v.dim() > 0) for k, v in cpu_params(m).items()}
    ctx, xt, yt, aux = task_tensors(tasks)
    O.loss_fn(P, ctx, xt, yt, aux, m.config.points_per_unit).backward()
    ref = torch.cat([P[k].grad.flatten() for k in res[0]["names"]])
    err = float((res[0]["flat"] - ref).norm() / ref.norm())
    assert err < 1e-5, err
