import os, sys, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from deepsensornz_b200 import ConvNP
from deepsensornz_b200.predict import _batch_contexts
from deepsensornz_b200.synthetic import make_task
torch.manual_seed(0)
model = ConvNP(precision="bf16", **bench.model_kwargs())
eng = model.engine
static = bench.static_fields(with_aux_hi=True)
x_hi = static.x_hi
rng = np.random.default_rng(99)
tasks = [make_task(static, 2016010100 + h, n_stations=int(rng.integers(150, 201)), all_context=True) for h in range(64)]
aux_dev = torch.from_numpy(static.aux_hi[None]).to(eng.device)
for nb in (1, 2, 4, 8):
    groups = [tasks[i:i + nb] for i in range(0, 64, nb)]
    dbs = []
    t0 = time.perf_counter()
    cache = {}
    for g in groups:
        ctxs = _batch_contexts(g)
        xt = (np.broadcast_to(x_hi[None], (len(g), 1400)), np.broadcast_to(x_hi[None], (len(g), 1400)))
        hb = eng.stage_host(ctxs, xt, None, None, pinned=False, ctx_cache=cache)
        hb.aux_t = aux_dev
        dbs.append(eng.upload(hb))
    torch.cuda.synchronize()
    t_stage = (time.perf_counter() - t0) / 64
    for db in dbs[:2]:
        model(db)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for db in dbs:
        out = model(db)
    e1.record(); torch.cuda.synchronize()
    t_host = (time.perf_counter() - t0) / 64
    print(f"nb={nb}: stage+upload {t_stage*1e3:.3f} ms/task; forward loop: device {e0.elapsed_time(e1)/64:.3f} ms/task, wall {t_host*1e3:.3f} ms/task", flush=True)
    if nb == 4:
        eng.profile_start()
        for db in dbs[:4]: model(db)
        prof = eng.profile_stop()
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]): print(f"    {k:30s} {v['launches']//4:3d} {v['ms']/4*1e3:8.1f} us per batch of 4")
