#!/usr/bin/env python
"""bench.py -- ConvNP training step throughput on synthetic NZ-shaped tasks (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (bf16 UNet)
  python bench.py --impl reference --steps K --warmup W    # the CPU restatement of the reference path

A "step" is one ConvNP forward + Gaussian NLL + backward + AdamW update over one batch of 16 daily
NZ tasks (BASELINE configs[1]: ERA5-shaped 140x140 base grid + 6-ch aux grid + 1400x1400 land mask +
160 context stations; 40 target stations with 5 aux-at-target channels; internal_density 250).

  value : tasks/s, inputs already resident in HBM, device-timed with CUDA events (max over ranks)
  e2e   : tasks/s through ConvNP.loss_fn(host task) -- pinned H2D of every input inside the timed region,
          backward, optimiser step and the D2H read of the loss
  roofline     : the dominant kernel (tcgen05 conv) timed per launch with CUDA events
  cpu_baseline : the oracle (torch CPU restatement, "port") on a bounded sample, rank 0 / N=1 only

One process per GPU; under torchrun the gradient bucket is all-reduced with NCCL (weak scaling:
every rank steps its own 16 tasks).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 16
PPU = 250
DIM_YC = (3, 6, 1, 1)
N_STATIONS, CTX_FRAC = 200, 0.8


def model_kwargs(ppu=None):
    ppu = ppu or PPU
    n_lo = 140
    s_lo = 0.5 * (0.99643 - 0.00357) / (n_lo - 1)
    return dict(dim_yc=DIM_YC, dim_yt=1, dim_aux_t=5, internal_density=ppu,
                encoder_scales=(s_lo, s_lo, 0.5 / 1399.0, 0.5 / ppu), decoder_scale=1.0 / ppu,
                unet_channels=(64,) * 4, verbose=False)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf=d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)),
                    source="measured")
    return dict(hbm=6650.0, tf=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None and not self.rows:
            time.sleep(0.05)   # very short timed regions: give the sampler one period
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_task_lists(n_batches: int, rank: int, dim_yc=DIM_YC):
    """``n_batches`` lists of BATCH raw (numpy, un-batched) tasks -- what ``train_epoch`` receives."""
    from deepsensornz_b200.synthetic import make_static, make_task
    static = make_static(seed=7)
    return [[make_task(static, 20160101 + rank * 10000 + k * BATCH + i, n_stations=N_STATIONS, context_frac=CTX_FRAC,
                       c0_channels=dim_yc[0]) for i in range(BATCH)] for k in range(n_batches)]


def make_batches(n_batches: int, rank: int):
    from deepsensornz_b200 import concat_tasks
    return [concat_tasks(tasks) for tasks in make_task_lists(n_batches, rank)]


# ------------------------------------------------------------------------------------------------
# CPU restatement of the reference path (oracle) -- the reported baseline / --impl reference arm
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(n_tasks: int, reps: int, warmup: int):
    """fwd + NLL + bwd of the oracle on ``n_tasks`` concatenated tasks; returns (tasks/s, s/step, cores)."""
    from deepsensornz_b200 import ConvNP, concat_tasks
    from deepsensornz_b200.synthetic import make_static, make_task
    from oracle import convnp_oracle as O
    from tests.util import oracle_inputs
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    m = ConvNP(**model_kwargs())
    P = {k: v.detach().cpu().clone().requires_grad_(v.dim() > 0) for k, v in m.model.state_dict().items()}
    static = make_static(seed=7)
    tasks = [make_task(static, 20160101 + i, n_stations=N_STATIONS, context_frac=CTX_FRAC) for i in range(n_tasks)]
    task = concat_tasks(tasks) if n_tasks > 1 else tasks[0]
    ctx, xt, yt, aux = oracle_inputs(task)
    times = []
    for it in range(warmup + reps):
        t0 = time.perf_counter()
        for v in P.values():
            v.grad = None
        loss = O.loss_fn(P, ctx, xt, yt, aux, PPU)
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = float(np.median(times))
    return n_tasks / sec, sec, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_tasks = 1
    rate, sec, cores = cpu_reference_rate(n_tasks, max(1, args.steps), max(0, min(args.warmup, 1)))
    line = {
        "impl": "reference", "metric": "convnp_train_tasks_per_s", "value": rate, "unit": "tasks/s",
        "n_gpus": int(args.gpus), "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {k: v for k, v in workload_config(1, "fp32").items() if k != "static_context_dedup"},
        "cpu_baseline": {"value": rate, "unit": "tasks/s", "cores": cores, "kind": "port",
                         "sample": f"{n_tasks} task per step (fwd+NLL+bwd), torch CPU oracle, all host threads"},
        "e2e": {"value": rate, "unit": "tasks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus: int, precision: str, grid=None):
    cfg = {"workload": "configs[1]: ConvNP training step, batch of 16 synthetic daily NZ tasks "
                       "(ERA5-shaped 140x140 + 6-ch aux + 1400x1400 land mask + 160 context / 40 target stations)",
           "global_batch": BATCH * max(n_gpus, 1), "per_gpu_batch": BATCH,
           "internal_density": int(round(1.0 / grid.res)) if grid is not None else PPU,
           "unet_channels": [64, 64, 64, 64], "precision": precision,
           "parallelism": f"dp{max(n_gpus, 1)}", "l2": "per-step working set (>3 GB) exceeds the 126 MB L2",
           "static_context_dedup": "context sets that are bit-identical across the 16 tasks of a batch (topography aux, "
                                   "land mask) are staged, uploaded and encoded once per step and broadcast"}
    if grid is not None:
        cfg["internal_grid"] = [grid.n1, grid.n2]
    return cfg


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from deepsensornz_b200 import ConvNP

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (the hot path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner (and, at NCCL_DEBUG=INFO, its whole log) on stdout while the communicator is
        # created, ahead of the one JSON line the driver reads: stdout is pointed at stderr for the duration of the
        # initialisation (file-descriptor level, the banner comes from C code)
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            t = torch.zeros(1, device="cuda")
            dist.all_reduce(t)                     # forces the (possibly lazy) communicator creation now
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    torch.manual_seed(0)
    model = ConvNP(precision=args.precision, **model_kwargs(args.internal_density))
    if world > 1:
        from deepsensornz_b200.dist import enable_data_parallel
        enable_data_parallel(model)
    # the reference's optimiser (train.py:354); fused=True is torch's single-kernel implementation of the same update
    opt = torch.optim.AdamW(model.model.parameters(), lr=5e-5, weight_decay=1e-5, fused=True,
                            capturable=(world == 1 and not args.no_graph))
    eng = model.engine

    tasks = make_batches(2, rank)
    host = [model.stage_task(t, pinned=True) for t in tasks]
    dev = [eng.upload(h) for h in host]
    h2d_bytes = dev[0].h2d_bytes
    static_shared = [not c.y_batched for c in host[0].contexts]
    torch.cuda.synchronize()

    def step(batch):
        opt.zero_grad(set_to_none=True)
        loss = model.loss_fn(batch, normalise=True)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(args.warmup):
        step(dev[i % 2])
    # One-GPU runs replay the whole step (forward, NLL, backward, AdamW) as a CUDA graph (deepsensornz_b200/graph.py);
    # the data-parallel step keeps eager launches around the NCCL all-reduce.
    gs = None
    if world == 1 and not args.no_graph:
        from deepsensornz_b200.graph import GraphedTrainStep
        l0 = eng.launches
        gs = GraphedTrainStep(model, opt, dev[0], warmup=1)
        graph_launches = (eng.launches - l0) // 2      # one warm-up step + the captured step
        run = lambda batch: gs.step(batch)
    else:
        run = step
    run(dev[0])
    # ---- device-resident throughput ----
    launches0 = eng.launches
    with ClockSampler(local) as clk:
        ms = timed(lambda i: run(dev[i % 2]), args.steps)
    launches = (eng.launches - launches0) if gs is None else graph_launches * args.steps
    clocks = clk.summary()
    value = world * BATCH * args.steps / (ms * 1e-3)
    # ---- end to end through the public API with host buffers ----
    # every step: pinned-host -> device copy of that step's batch (on a copy stream, overlapping the previous step,
    # as train_epoch does), forward, backward, optimiser step and the D2H read of the loss
    last = {}
    copy_stream = torch.cuda.Stream()
    pending = {"b": eng.upload(host[0], stream=copy_stream)}

    # the loss of every step is copied to pinned memory asynchronously and read by the host one step later (while the
    # next step is already queued), exactly as deepsensornz_b200.train_epoch does
    slots = [torch.empty((), dtype=torch.float64).pin_memory() for _ in range(2)]
    evs = [None, None]

    def e2e_step(i):
        cur = pending["b"]
        loss = run(cur)
        slots[i % 2].copy_(loss.detach().to(torch.float64), non_blocking=True)   # D2H of the loss every step
        evs[i % 2] = torch.cuda.Event()
        evs[i % 2].record()
        pending["b"] = eng.upload(host[(i + 1) % 2], stream=copy_stream)   # next step's inputs
        j = (i + 1) % 2
        if evs[j] is not None:
            evs[j].synchronize()
            last["loss"] = float(slots[j])

    e2e_step(0)
    ms_e2e = timed(e2e_step, args.steps)
    evs[(args.steps - 1) % 2].synchronize()  # (the timed region ends with a device synchronise; this is the last read)
    last["loss"] = float(slots[(args.steps - 1) % 2])
    e2e = world * BATCH * args.steps / (ms_e2e * 1e-3)
    # ---- per-kernel timing (CUDA events around every launch) for the roofline ----
    # (the side-stream packing and the concurrent stride-2 dgrad phases are switched off here so that every launch is
    # timed alone on one stream; the timed regions above run with them on)
    os.environ["CNP_NO_PREPACK"] = os.environ["CNP_NO_MULTISTREAM"] = "1"
    step(dev[0])
    eng.profile_start()
    for i in range(2):
        step(dev[i % 2])
    prof = eng.profile_stop()
    del os.environ["CNP_NO_PREPACK"], os.environ["CNP_NO_MULTISTREAM"]
    pk = peaks()
    roof = None
    if "cnp_conv_tc2" in prof:
        d = prof["cnp_conv_tc2"]
        ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
        tot_ms = sum(v["ms"] for v in prof.values())
        traffic = None     # DRAM bytes per launch from the committed ncu pass (profiles/r01_traffic.json), same command
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath) and args.internal_density == PPU:
            with open(tpath) as f:
                traffic = json.load(f).get("conv_tc2_kernel", {}).get("dram_bytes_per_launch")
        roof = {"kernel": "conv_tc2_kernel (tcgen05 implicit-GEMM conv, fwd + dgrad launches)", "bound": "tensor", "achieved": ach,
                "peak": pk["tf"], "unit": "TFLOP/s", "frac": ach / pk["tf"], "traffic": traffic,
                "peak_source": f"{pk['source']} bf16_tflops_sustained", "share_of_step": d["ms"] / tot_ms,
                "avg_launch_ms": d["ms"] / d["launches"], "flops_per_launch": d["flops"] / d["launches"]}
    kernels = {k: {"launches": v["launches"] // 2, "ms_per_step": v["ms"] / 2,
                   "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["flops"] else None,
                   "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["bytes"] else None}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    line = {
        "metric": "convnp_train_tasks_per_s", "value": value, "unit": "tasks/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": workload_config(world, args.precision, dev[0].grid),
        "e2e": {"value": e2e, "unit": "tasks/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "cuda_graph": gs is not None, "clocks": clocks, "roofline": roof, "kernels": kernels,
        "loss": last.get("loss"),
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            rate, sec, cores = cpu_reference_rate(1, 2, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "tasks/s", "cores": cores, "kind": "port",
                                    "sample": "1 task per step (fwd+NLL+bwd), 2 timed reps after 1 warm-up, "
                                              "torch CPU oracle, all host threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--internal-density", type=int, default=PPU,
                    help="points per unit of the internal grid: 250 (saved models, 304^2 grid) or 500 (repo default, 608^2)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
