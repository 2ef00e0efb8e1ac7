"""Per-launch timing of the SetConv encoder at the benchmark shapes (B = 16 training batch, B = 1 inference task):
CUDA events around every launch (Engine.profile_start), algorithmic bytes / time against the measured HBM peak.
  python tools/bench_encoder.py [reps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deepsensornz_b200 import ConvNP, concat_tasks  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    torch.manual_seed(0)
    out = {}
    for name, dim_yc, nb in (("train_b16", bench.DIM_YC, 16), ("train_mv_b16", bench.DIM_YC_MV, 16), ("infer_b1", bench.DIM_YC, 1)):
        model = ConvNP(precision="bf16", **bench.model_kwargs(dim_yc=dim_yc))
        eng = model.engine
        tasks = bench.make_task_lists(1, 0, dim_yc)[0][:nb]
        batch = model._to_device(concat_tasks(tasks) if nb > 1 else tasks[0])
        for _ in range(3):
            eng.encode_blocked(batch)
        torch.cuda.synchronize()
        eng.profile_start()
        for _ in range(reps):
            assert eng.encode_blocked(batch) is not None
        prof = eng.profile_stop()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.encode_blocked(batch)
        e1.record()
        torch.cuda.synchronize()
        pk = bench.peaks()
        d = {"launches": 0, "ms": 0.0, "bytes": 0.0}
        for kname in ("cnp_encode_hpass", "cnp_encode_vpass", "cnp_encode_fused"):
            for f in d:
                d[f] += prof[kname][f]
        out[name] = {"launches_per_call": d["launches"] // reps, "us_per_call_serial": d["ms"] / reps * 1e3,
                     "us_hpass": prof["cnp_encode_hpass"]["ms"] / reps * 1e3, "us_vpass": prof["cnp_encode_vpass"]["ms"] / reps * 1e3, "us_fused": prof["cnp_encode_fused"]["ms"] / reps * 1e3,
                     "us_per_call_back_to_back": e0.elapsed_time(e1) / reps * 1e3,
                     "algorithmic_MB_per_call": d["bytes"] / reps / 1e6,
                     "GBps": d["bytes"] / (d["ms"] * 1e-3) / 1e9, "frac_of_hbm_peak": d["bytes"] / (d["ms"] * 1e-3) / 1e9 / pk["hbm"]}
        eng.release_workspaces()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
