"""predict end to end, several calls in a row, with and without the result-array pool (CONVNP_B200_RESULT_POOL_MB=0)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from deepsensornz_b200 import ConvNP
from deepsensornz_b200 import predict as P
from deepsensornz_b200.synthetic import make_static, make_task

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
model = ConvNP(precision="bf16", **bench.model_kwargs())
static = bench.static_fields(with_aux_hi=True)
rng = np.random.default_rng(99)
tasks = [make_task(static, 2016010100 + h, n_stations=int(rng.integers(150, 201)), all_context=True) for h in range(n)]
kw = dict(X_t=(static.x_hi, static.x_hi), X_t_is_normalised=True, aux_at_targets_override=static.aux_hi)
model.predict(tasks[:3], **kw)
for rep in range(int(os.environ.get('CALLS', 5))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pred = model.predict(tasks, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"call {rep}: {1e3 * dt / n:.3f} ms/date   pool hits {P._result_pool.hits} misses {P._result_pool.misses}", flush=True)
    del pred

if os.environ.get("PROFILE_CALLS"):
    import cProfile, pstats, io
    for rep in range(3):
        pr = cProfile.Profile()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pr.enable()
        pred = model.predict(tasks, **kw)
        pr.disable()
        torch.cuda.synchronize()
        print(f"profiled call {rep}: {1e3 * (time.perf_counter() - t0) / n:.3f} ms/date")
        st = io.StringIO()
        pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(12)
        print("\n".join(st.getvalue().splitlines()[4:24]))
        del pred
