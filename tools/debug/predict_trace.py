"""Where does predict's wall time go?  Per batch: host time of stage_host / upload / forward (issue only), GPU time of the
forward (events), and the GPU idle gaps between consecutive forwards."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from deepsensornz_b200 import ConvNP
from deepsensornz_b200.synthetic import make_task

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
model = ConvNP(precision="bf16", **bench.model_kwargs())
static = bench.static_fields(with_aux_hi=True)
rng = np.random.default_rng(99)
tasks = [make_task(static, 2016010100 + h, n_stations=int(rng.integers(150, 201)), all_context=True) for h in range(n)]
kw = dict(X_t=(static.x_hi, static.x_hi), X_t_is_normalised=True, aux_at_targets_override=static.aux_hi)
for _ in range(3):
    model.predict(tasks, **kw)
eng = model.engine
rec = {"stage": [], "upload": [], "fwd_host": [], "ev": []}
o_stage, o_upload, o_call = eng.stage_host, eng.upload, type(model).__call__

def stage(*a, **k):
    t = time.perf_counter(); r = o_stage(*a, **k); rec["stage"].append(time.perf_counter() - t); return r
def upload(*a, **k):
    t = time.perf_counter(); r = o_upload(*a, **k); rec["upload"].append(time.perf_counter() - t); return r
def call(self, *a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter(); e0.record(); r = o_call(self, *a, **k); e1.record()
    rec["fwd_host"].append(time.perf_counter() - t); rec["ev"].append((e0, e1)); return r
eng.stage_host, eng.upload = stage, upload
type(model).__call__ = call
torch.cuda.synchronize()
t0 = time.perf_counter()
pred = model.predict(tasks, **kw)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
gpu = [a.elapsed_time(b) for a, b in rec["ev"]]
gaps = [rec["ev"][i][1].elapsed_time(rec["ev"][i + 1][0]) for i in range(len(gpu) - 1)]
ms = lambda v: " ".join(f"{1e3 * x:.2f}" for x in v)
print(f"wall {1e3 * wall:.1f} ms for {n} dates = {1e3 * wall / n:.3f} ms/date; batches {len(gpu)}")
print("stage_host ms :", ms(rec["stage"]))
print("upload ms     :", ms(rec["upload"]))
print("forward issue :", ms(rec["fwd_host"]))
print("forward GPU ms:", " ".join(f"{x:.2f}" for x in gpu))
print("GPU gaps ms   :", " ".join(f"{x:.2f}" for x in gaps))
print(f"sum GPU {sum(gpu):.1f} ms, sum gaps {sum(gaps):.1f} ms")
