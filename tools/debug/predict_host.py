import os, sys, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from deepsensornz_b200 import ConvNP
from deepsensornz_b200.synthetic import make_task
torch.manual_seed(0)
model = ConvNP(precision="bf16", **bench.model_kwargs())
static = bench.static_fields(with_aux_hi=True)
x_hi = static.x_hi
rng = np.random.default_rng(99)
tasks = [make_task(static, 2016010100 + h, n_stations=int(rng.integers(150, 201)), all_context=True) for h in range(64)]
kw = dict(X_t=(x_hi, x_hi), X_t_is_normalised=True, aux_at_targets_override=static.aux_hi)
def run(label):
    model.predict(tasks[:4], **kw); torch.cuda.synchronize()
    t0 = time.perf_counter(); model.predict(tasks, **kw); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{label:40s} {dt / len(tasks) * 1e3:.3f} ms/task", flush=True)
for thr in ("1", "3", "6", "10"):
    os.environ["CONVNP_B200_DRAIN_THREADS"] = thr
    run(f"drain threads {thr}")
os.environ["CONVNP_B200_DRAIN_THREADS"] = "6"
os.environ["CONVNP_B200_PREDICT_GRAPH"] = "1"
run("graph + 6 drain threads")
os.environ["CONVNP_B200_PREDICT_GRAPH"] = "0"
# pinned allocation speed and raw D2H / memcpy rates
t0 = time.perf_counter(); big = torch.empty(64, 2, 1400, 1400, dtype=torch.float32, device="cpu", pin_memory=True); t1 = time.perf_counter()
print(f"pinned alloc of {big.numel()*4/1e9:.2f} GB: {(t1-t0)*1e3:.1f} ms", flush=True)
del big
t0 = time.perf_counter(); big = torch.empty(64, 2, 1400, 1400, dtype=torch.float32, device="cpu", pin_memory=True); t1 = time.perf_counter()
print(f"pinned alloc again (cached): {(t1-t0)*1e3:.1f} ms", flush=True)
d = torch.empty(2, 1400, 1400, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(64): big[i].copy_(d, non_blocking=True)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"D2H 15.68 MB x 64 into pinned: {(t1-t0)/64*1e3:.3f} ms each = {15.68e-3/((t1-t0)/64):.1f} GB/s", flush=True)
a = np.empty((2, 1400, 1400), np.float32); src = big[0].numpy()
t0 = time.perf_counter()
for i in range(20): np.copyto(a, src)
t1 = time.perf_counter()
print(f"host memcpy 15.68 MB: {(t1-t0)/20*1e3:.3f} ms = {15.68e-3/((t1-t0)/20):.1f} GB/s", flush=True)
import cProfile, pstats, io
os.environ["CONVNP_B200_DRAIN_THREADS"] = "6"
pr = cProfile.Profile(); pr.enable(); model.predict(tasks, **kw); torch.cuda.synchronize(); pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(18); print(st.getvalue()[:5000])
