"""World-size-2 gloo test of the data-parallel plumbing (the N>1 path of SURVEY 8(e)) on CPU."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepsensornz_b200 import ConvNP
    from deepsensornz_b200.dist import allreduce_mean_, enable_data_parallel
    torch.manual_seed(100 + rank)                       # different init per rank on purpose
    m = ConvNP(dim_yc=(3, 6, 1, 1), dim_yt=1, dim_aux_t=5, internal_density=50, verbose=False)
    enable_data_parallel(m)
    w = m.model.decoder.unet.initial_linear.weight.detach().clone()
    flat = torch.full((10,), float(rank + 1))
    allreduce_mean_(flat, world)
    q.put((rank, w, flat, m.engine.world_size))
    dist.destroy_process_group()


def test_broadcast_and_allreduce_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(res[0][1], res[1][1])            # identical weights after broadcast
    assert torch.allclose(res[0][2], torch.full((10,), 1.5)) and torch.allclose(res[1][2], torch.full((10,), 1.5))
    assert res[0][3] == 2
