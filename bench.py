#!/usr/bin/env python
"""bench.py -- ConvNP throughput on synthetic NZ-shaped tasks (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W                  # configs[1]: training step, B = 16, bf16 UNet
  python bench.py --workload train_mv  ...                       # configs[3]: multi-variable base grid (8 ch, Cin = 20)
  python bench.py --workload infer     ...                       # configs[2]/[4]: ConvNP.predict onto the 1400 x 1400 grid
  python bench.py --impl reference --steps K --warmup W          # the CPU restatement of the reference path

Training ("train", "train_mv"): a "step" is one ConvNP forward + Gaussian NLL + backward + AdamW update over one batch of
16 daily NZ tasks (ERA5-shaped 140x140 base grid + 6-ch aux grid + 1400x1400 land mask + 160 context stations; 40 target
stations with 5 aux-at-target channels; internal_density 250 -> 304 x 304 internal grid).

  value : tasks/s, inputs already resident in HBM, device-timed with CUDA events (max over ranks)
  e2e   : tasks/s through the reference's call form, ``train_epoch(model, list_of_numpy_Tasks, batch_size=16, opt=opt)``
          (nzdownscale/downscaler/train.py:388-394): batching of the raw tasks, staging in pinned memory, H2D of every
          step's per-date tensors, forward, backward, optimiser step and the D2H read of every loss, all inside the timed
          region; ``host_ms_per_batch`` is the host build time of one batch (it runs on a worker thread)
  roofline     : the dominant kernel (tcgen05 conv) timed per launch with CUDA events, against the BURST cuBLAS peak
                 (a kernel timed alone); ``frac_sustained`` is the same against the sustained peak
  cpu_baseline : the oracle (torch CPU restatement, "port") on a bounded sample, rank 0 / N=1 only
  loss_check   : the first forward on the initial weights against the committed fp32 oracle value
                 (tests/golden/s2_bench16_loss0.npz) -- the run aborts when it is off by more than the bf16 tolerance
  inference, multivar, grid608 : short sub-records of the other workloads (configs[2]/[4], configs[3], and
                 configs[1] at the in-repo default internal_density=500), so that every run carries them

Inference ("infer"): a "step" is one task (one date/hour) predicted onto the 1400 x 1400 target grid (configs[2], and
configs[4] when sharded by date over N GPUs: no collective).  value = forward only with inputs resident (4 dates per
launch sequence, as ``predict`` runs them), outputs left on the device; e2e = ``ConvNP.predict(tasks, X_t=...)`` with H2D of the per-hour sets and the D2H of mean + std (15.7 MB per
task) inside the timed region; roofline = the fused tensor-core decoder (``decode_grid_tc``).

One process per GPU; under torchrun the gradient bucket is all-reduced with NCCL (weak scaling: every rank steps its own
16 tasks).
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
DIM_YC_MV = (8, 6, 1, 1)          # validation wrf.ipynb:206-210: t2m, precip, u10, v10, ... + cos/sin of the hour
N_STATIONS, CTX_FRAC = 200, 0.8


def model_kwargs(ppu=None, dim_yc=DIM_YC):
    ppu = ppu or PPU
    n_lo = 140
    s_lo = 0.5 * (0.99643 - 0.00357) / (n_lo - 1)
    return dict(dim_yc=dim_yc, dim_yt=1, dim_aux_t=5, internal_density=ppu,
                encoder_scales=(s_lo, s_lo, 0.5 / 1399.0, 0.5 / ppu), decoder_scale=1.0 / ppu,
                unet_channels=(64,) * 4, verbose=False)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        burst = d.get("bf16_tflops", 1590.0)
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf=burst, tf_sustained=d.get("bf16_tflops_sustained", burst),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md).  The sampler process
    is started EARLY (nvidia-smi takes seconds to come up when 8 ranks start one each) and runs until the arm ends; the
    summary uses the samples that arrived between ``mark_start`` and ``mark_end`` (the timed region), or -- when the region
    is shorter than the sampling period -- the nearest ones."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None
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

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def __enter__(self):
        self.t0 = time.perf_counter()
        return self

    def __exit__(self, *a):
        self.t1 = time.perf_counter()

    def close(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
            self.proc = None

    def summary(self):
        if self.proc is not None and self.t1 is not None:
            time.sleep(0.06)          # let the sample that covers the end of the region arrive
        rows = list(self.rows)
        inside = [r for t, r in rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.03]
        where = "timed region"
        if not inside and rows and self.t0 is not None:
            mid = 0.5 * (self.t0 + (self.t1 or self.t0))
            inside = [r for _, r in sorted(rows, key=lambda tr: abs(tr[0] - mid))[:3]]
            where = "nearest samples (region shorter than the sampling period)"
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "window": where}


_STATIC = {}


def static_fields(with_aux_hi=False):
    from deepsensornz_b200.synthetic import make_static
    key = bool(with_aux_hi)
    if key not in _STATIC:
        _STATIC[key] = make_static(seed=7, with_aux_hi=with_aux_hi)
    return _STATIC[key]


def make_task_lists(n_batches: int, rank: int, dim_yc=DIM_YC, first=0):
    """``n_batches`` lists of BATCH raw (numpy, un-batched) tasks -- what ``train_epoch`` receives
    (nzdownscale/downscaler/train.py:315-316 keeps the tasks as numpy)."""
    from deepsensornz_b200.synthetic import make_task
    static = static_fields()
    return [[make_task(static, 20160101 + rank * 10000 + (first + k) * BATCH + i, n_stations=N_STATIONS,
                       context_frac=CTX_FRAC, c0_channels=dim_yc[0]) for i in range(BATCH)] for k in range(n_batches)]


def make_batches(n_batches: int, rank: int, dim_yc=DIM_YC):
    from deepsensornz_b200 import concat_tasks
    return [concat_tasks(tasks) for tasks in make_task_lists(n_batches, rank, dim_yc)]


# ------------------------------------------------------------------------------------------------
# CPU restatement of the reference path (oracle) -- the reported baseline / --impl reference arm
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(n_tasks: int, reps: int, warmup: int, dim_yc=DIM_YC, infer=False):
    """Oracle on ``n_tasks`` concatenated tasks: fwd + NLL + bwd (training) or the on-grid forward (inference);
    returns (tasks/s, s/step, cores)."""
    from deepsensornz_b200 import ConvNP
    from deepsensornz_b200.synthetic import make_task
    from oracle import convnp_oracle as O
    from oracle.task_tensors import task_tensors
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    m = ConvNP(**model_kwargs(dim_yc=dim_yc))
    P = {k: v.detach().cpu().clone().requires_grad_(v.dim() > 0 and not infer) for k, v in m.model.state_dict().items()}
    if infer:
        static = static_fields(with_aux_hi=True)
        tasks = [make_task(static, 2016010100 + i, all_context=True, grid_targets=True) for i in range(n_tasks)]
    else:
        static = static_fields()
        tasks = [make_task(static, 20160101 + i, n_stations=N_STATIONS, context_frac=CTX_FRAC, c0_channels=dim_yc[0])
                 for i in range(n_tasks)]
    ctx, xt, yt, aux = task_tensors(tasks)
    times = []
    for it in range(warmup + reps):
        t0 = time.perf_counter()
        if infer:
            with torch.no_grad():
                mean, var = O.forward(P, ctx, xt, aux, PPU)
                _ = var.sqrt()
        else:
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
    infer = args.workload == "infer"
    dim_yc = DIM_YC_MV if args.workload == "train_mv" else DIM_YC
    n_tasks = 1     # a bounded sample: one task per step (a 16-task step is 16x the same CPU work, ~15 s per step)
    rate, sec, cores = cpu_reference_rate(n_tasks, max(1, args.steps), max(0, min(args.warmup, 1)), dim_yc, infer)
    cfg = {k: v for k, v in workload_config(args.workload, 1, "fp32", dim_yc=dim_yc).items() if k != "static_context_dedup"}
    cfg.update(global_batch=n_tasks, per_gpu_batch=n_tasks, parallelism="cpu",
               sample_of="the same workload at 1 task per step (the oracle's cost is linear in the batch)")
    what = "on-grid forward onto 1400x1400" if infer else "fwd+NLL+bwd"
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": rate, "unit": "tasks/s",
        "n_gpus": int(args.gpus), "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": rate, "unit": "tasks/s", "cores": cores, "kind": "port",
                         "sample": f"{n_tasks} task per step ({what}), torch CPU oracle, all host threads"},
        "e2e": {"value": rate, "unit": "tasks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def metric_name(workload):
    return "convnp_predict_tasks_per_s" if workload == "infer" else "convnp_train_tasks_per_s"


def workload_config(workload: str, n_gpus: int, precision: str, grid=None, dim_yc=DIM_YC):
    if workload == "infer":
        cfg = {"workload": "configs[2]/[4]: ConvNP.predict onto the 1400x1400 NZ target grid from an ERA5-shaped 140x140 "
                           "base grid + 6-ch aux + 1400x1400 land mask + 150-200 context stations, one task (date/hour) "
                           "per step, sharded by date across the GPUs (no collective)",
               "global_batch": max(n_gpus, 1), "per_gpu_batch": 1, "target_grid": [1400, 1400]}
    else:
        which = "configs[3]: multi-variable base grid (8 channels: t2m, precipitation, u10, v10, ... ; Cin = 20)" \
            if workload == "train_mv" else "configs[1]"
        cfg = {"workload": f"{which}: ConvNP training step, batch of 16 synthetic daily NZ tasks "
                           f"(ERA5-shaped {dim_yc[0]}x140x140 + 6-ch aux + 1400x1400 land mask + 160 context / 40 target "
                           "stations)",
               "global_batch": BATCH * max(n_gpus, 1), "per_gpu_batch": BATCH}
    cfg.update({"dim_yc": list(dim_yc), "internal_density": int(round(1.0 / grid.res)) if grid is not None else PPU,
                "unet_channels": [64, 64, 64, 64], "precision": precision, "parallelism": f"dp{max(n_gpus, 1)}",
                "l2": "per-step working set (>3 GB training, >100 MB per inference task) exceeds the 126 MB L2",
                "static_context_dedup": "context sets that every task of a batch shares as one buffer (topography aux, land "
                                        "mask) are uploaded once and encoded once per step"})
    if grid is not None:
        cfg["internal_grid"] = [grid.n1, grid.n2]
    return cfg


# ------------------------------------------------------------------------------------------------
def init_dist():
    import torch.distributed as dist
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
    return world, rank, local


def make_timer(world):
    import torch.distributed as dist

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        """barrier + synchronise, ``steps`` calls of fn(i) between two CUDA events, barrier + synchronise; max over ranks."""
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

    return barrier, timed


def kernel_table(prof, reps, pk):
    return {k: {"launches": v["launches"] // reps, "ms_per_step": v["ms"] / reps,
                "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["flops"] else None,
                "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["bytes"] else None,
                "frac_hbm": (v["bytes"] / (v["ms"] * 1e-3) / 1e9 / pk["hbm"]) if v["bytes"] else None}
            for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}


def train_arm(args, world, rank, local, dim_yc, steps, warmup, profile=True, e2e=True, check_loss=False):
    """Device-resident and end-to-end training throughput for one model shape; returns the fields of the JSON line."""
    from deepsensornz_b200 import ConvNP, concat_tasks, train_epoch
    barrier, timed = make_timer(world)
    clk = ClockSampler(local)
    torch.manual_seed(0)
    model = ConvNP(precision=args.precision, **model_kwargs(args.internal_density, dim_yc))
    if world > 1:
        from deepsensornz_b200.dist import enable_data_parallel
        enable_data_parallel(model)
    use_graph = not args.no_graph and (world == 1 or os.environ.get("CONVNP_B200_DP_GRAPH", "1") == "1")
    # the reference's optimiser (train.py:354); fused=True is torch's single-kernel implementation of the same update
    opt = torch.optim.AdamW(model.model.parameters(), lr=5e-5, weight_decay=1e-5, fused=True, capturable=use_graph)
    eng = model.engine

    lists = make_task_lists(2, rank, dim_yc)
    host = [model.stage_task(concat_tasks(t), pinned=True) for t in lists]
    dev = [eng.upload(h) for h in host]
    torch.cuda.synchronize()
    out = {}
    if check_loss and rank == 0 and args.internal_density == PPU:
        gpath = os.path.join(ROOT, "tests", "golden", "s2_bench16_loss0.npz")
        with torch.no_grad():
            got = float(model.loss_fn(dev[0], normalise=True))
        gold = float(np.load(gpath)["loss"])
        tol = 2e-2 if args.precision == "bf16" else 1e-5
        err = abs(got - gold) / abs(gold)
        out["loss_check"] = {"golden_fp32_oracle": gold, "value": got, "rel_err": err, "tol": tol,
                             "what": "first forward on the initial weights, batch 0 of rank 0"}
        if not err < tol:
            raise SystemExit(f"bench.py: loss of the first forward {got} differs from the oracle's {gold} (rel {err:.2e})")

    def step(batch):
        opt.zero_grad(set_to_none=True)
        loss = model.loss_fn(batch, normalise=True)
        loss.backward()
        opt.step()
        return loss

    for i in range(warmup):
        step(dev[i % 2])
    # The whole step (forward, NLL, backward, [bucketed NCCL all-reduces,] AdamW) is replayed as a CUDA graph
    # (deepsensornz_b200/graph.py); CONVNP_B200_DP_GRAPH=0 keeps eager launches around NCCL at N > 1.
    gs = None
    if use_graph:
        from deepsensornz_b200.graph import GraphedTrainStep
        gs = GraphedTrainStep(model, opt, dev[0], warmup=1)
        graph_launches = gs.launches
        run = lambda batch: gs.step(batch)
    else:
        run = step
    run(dev[0])
    # ---- device-resident throughput ----
    launches0 = eng.launches
    with clk:
        ms = timed(lambda i: run(dev[i % 2]), steps)
    launches = (eng.launches - launches0) if gs is None else graph_launches * steps
    out.update(value=world * BATCH * steps / (ms * 1e-3), ms_per_step=ms / steps, gpu_launches=launches,
               cuda_graph=gs is not None, clocks=clk.summary(), grid=dev[0].grid)
    clk.close()
    # ---- end to end: train_epoch over lists of raw numpy tasks (the reference's call form) ----
    if e2e:
        gs = run = None
        n_e2e = steps
        tasks = [t for lst in make_task_lists(n_e2e, rank, dim_yc, first=2) for t in lst]
        warm = [t for lst in lists for t in lst] * 2            # 4 batches: eager, capture, two replays
        np.random.seed(1234 + rank)
        train_epoch(model, warm, batch_size=BATCH, opt=opt, use_graph=use_graph)
        stager = model.__dict__.get("_stager")
        if stager is not None:
            stager.host_ms.clear()
        res = {}

        def epoch(_):
            res["losses"] = train_epoch(model, tasks, batch_size=BATCH, opt=opt, use_graph=use_graph)

        ms_e2e = timed(epoch, 1)
        assert len(res["losses"]) == n_e2e and all(np.isfinite(res["losses"]))
        out["e2e"] = {"value": world * BATCH * n_e2e / (ms_e2e * 1e-3), "unit": "tasks/s",
                      "h2d_bytes_per_step": stager.h2d_bytes if stager is not None else None, "d2h_bytes_per_step": 8,
                      "ms_per_step": ms_e2e / n_e2e,
                      "host_ms_per_batch": float(np.mean(stager.host_ms)) if stager is not None and stager.host_ms else None,
                      "api": "train_epoch(model, list[Task] (numpy), batch_size=16, opt=opt): batching + pinned staging + "
                             "H2D + fwd + bwd + AdamW + loss D2H per step, all inside the timed region",
                      "static_sets_resident": len(stager._static) if stager is not None else 0}
        out["loss"] = float(res["losses"][-1])
        model.__dict__.pop("_train_graphs", None)
    # ---- per-kernel timing (CUDA events around every launch) for the roofline ----
    # (the side-stream packing and the concurrent stride-2 dgrad phases are switched off here so that every launch is
    # timed alone on one stream; the timed regions above run with them on)
    if profile:
        os.environ["CNP_NO_PREPACK"] = os.environ["CNP_NO_MULTISTREAM"] = "1"
        step(dev[0])
        eng.profile_start()
        for i in range(2):
            step(dev[i % 2])
        prof = eng.profile_stop()
        del os.environ["CNP_NO_PREPACK"], os.environ["CNP_NO_MULTISTREAM"]
        pk = peaks()
        if "cnp_conv_tc2" in prof:
            d = prof["cnp_conv_tc2"]
            ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
            tot_ms = sum(v["ms"] for v in prof.values())
            traffic = None     # DRAM bytes per launch from the committed ncu pass of the same command
            tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
            if os.path.exists(tpath) and args.internal_density == PPU and tuple(dim_yc) == DIM_YC:
                with open(tpath) as f:
                    traffic = json.load(f).get("conv_tc2_kernel", {}).get("dram_bytes_per_launch")
            out["roofline"] = {"kernel": "conv_tc2_kernel (tcgen05 implicit-GEMM conv, fwd + dgrad launches)",
                               "bound": "tensor", "achieved": ach, "peak": pk["tf"], "unit": "TFLOP/s", "frac": ach / pk["tf"],
                               "frac_sustained": ach / pk["tf_sustained"], "peak_sustained": pk["tf_sustained"],
                               "traffic": traffic,
                               "peak_source": f"{pk['source']}: bf16_tflops (burst -- the kernel is timed per launch with "
                                              "CUDA events in a serial pass)",
                               "flops_counted": "as issued: the polyphase launches count their 4x4-tap phase convolutions, not "
                                                "the 5x5 taps on the upsampled tensor they replace (1.56x more)",
                               "share_of_step": d["ms"] / tot_ms, "avg_launch_ms": d["ms"] / d["launches"],
                               "flops_per_launch": d["flops"] / d["launches"]}
        out["kernels"] = kernel_table(prof, 2, pk)
        # the SetConv encoder as a whole, launched back to back (the per-launch events above include ~8 us of launch
        # latency per small kernel because the GPU idles between them): algorithmic bytes = every context tensor once
        # (fields shared by the batch once per step) + the blocked bf16 UNet input
        if args.precision == "bf16" and eng.encode_blocked(dev[0]) is not None:
            reps = 20
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                blk = eng.encode_blocked(dev[0])
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / reps * 1e3
            g = dev[0].grid
            by = sum(4.0 * (c.y.numel() + (c.mask.numel() if c.mask is not None else 0) +
                            sum(v.numel() for v in (c.x if isinstance(c.x, tuple) else (c.x,)))) for c in dev[0].contexts)
            by += 2.0 * BATCH * blk.CB * 8 * g.n1 * g.n2
            out["setconv_encoder"] = {"kernels": "enc_hpass + enc_vpass + enc_fused (enc_fused.cu), 3 launches back to back",
                                      "us_per_step": us, "algorithmic_mb": by / 1e6, "gbs": by / (us * 1e-6) / 1e9,
                                      "frac_hbm": by / (us * 1e-6) / 1e9 / pk["hbm"], "peak_gbs": pk["hbm"],
                                      "survey_8d_mb": 13.1 * BATCH,
                                      "frac_hbm_survey_8d": 13.1e6 * BATCH / (us * 1e-6) / 1e9 / pk["hbm"],
                                      "note": "survey_8d_mb = SURVEY 8(d)'s per-task figure (G = 320, Cin = 20, static sets "
                                              "re-read per task) x 16 tasks"}
    eng.release_workspaces()
    del model, opt, eng
    torch.cuda.empty_cache()
    return out


def infer_arm(args, world, rank, local, n_tasks, warmup, profile=True):
    """configs[2]/[4]: per-task forward onto the 1400 x 1400 grid (device-resident) and ConvNP.predict end to end."""
    from deepsensornz_b200 import ConvNP, Task
    from deepsensornz_b200.synthetic import make_task
    barrier, timed = make_timer(world)
    clk = ClockSampler(local)
    torch.manual_seed(0)
    model = ConvNP(precision=args.precision, **model_kwargs(args.internal_density))
    eng = model.engine
    static = static_fields(with_aux_hi=True)
    x_hi = static.x_hi
    # hourly tasks of this rank's date shard (8760 h / N GPUs in configs[4]); the station count varies hour to hour
    rng = np.random.default_rng(99 + rank)
    tasks = [make_task(static, 2016010100 + rank * 100000 + h, n_stations=int(rng.integers(150, 201)), all_context=True)
             for h in range(n_tasks)]
    kw = dict(X_t=(x_hi, x_hi), X_t_is_normalised=True, aux_at_targets_override=static.aux_hi)
    # ---- device-resident: inputs uploaded once, forward only, outputs stay on the device ----
    # the same forward ``predict`` runs: CONVNP_B200_PREDICT_BATCH (default 4) dates per launch sequence
    from deepsensornz_b200.predict import _batch_contexts
    nb = max(1, min(int(os.environ.get("CONVNP_B200_PREDICT_BATCH", "4")), n_tasks))
    n_groups = min(max(1, n_tasks // nb), 4)
    aux_dev = torch.from_numpy(static.aux_hi[None]).to(eng.device)
    dev = []
    for gi in range(n_groups):
        group = tasks[gi * nb:(gi + 1) * nb]
        xt = (np.broadcast_to(x_hi[None], (len(group), x_hi.size)), np.broadcast_to(x_hi[None], (len(group), x_hi.size)))
        db = eng.upload(eng.stage_host(_batch_contexts(group), xt, None, None, pinned=False))
        db.aux_t = aux_dev
        dev.append(db)
    torch.cuda.synchronize()
    n_fwd = max(1, n_tasks // nb)          # forwards in the timed region; each covers nb dates

    def fwd(i):
        with torch.no_grad():
            eng.forward(dev[i % n_groups], with_loss=False)

    for i in range(max(warmup, 3)):
        fwd(i)
    l0 = eng.launches
    with clk:
        ms = timed(fwd, n_fwd)
    launches = eng.launches - l0
    n_done = n_fwd * nb
    out = dict(value=world * n_done / (ms * 1e-3), ms_per_step=ms / n_done, gpu_launches=launches, clocks=clk.summary(),
               grid=dev[0].grid, dates_per_forward=nb)
    clk.close()
    # ---- end to end: ConvNP.predict (validate_ERA.py:88-92 / outputs/infer.py:96-103) ----
    # A first call of the same size (timed too: it pays the first touch of freshly allocated result arrays), dropped; the
    # timed call then gets its result arrays from predict's pool (deepsensornz_b200.predict._ResultPool: arrays that
    # nothing references any more are recycled) -- the steady state of a loop that predicts range after range.
    res = {}

    def run(_):
        res["pred"] = model.predict(tasks, **kw)

    from deepsensornz_b200.predict import _result_pool as _pool
    model.predict(tasks[:3], **kw)
    ms_first = timed(run, 1)
    res.clear()
    run(0)              # second call of this size: the pool page-locks the recycled arrays (once)
    res.clear()
    # three timed calls, the median reported (all three listed): on the shared host of a GPU box a single call now and
    # then stalls for tens of milliseconds in the driver or the kernel (seen as 3-5 ms per date outliers)
    calls = []
    for _ in range(3):
        res.clear()
        calls.append(timed(run, 1))
    ms_e2e = sorted(calls)[1]
    key = list(res["pred"].keys())[0]
    mean = np.asarray(res["pred"][key]["mean"])
    assert mean.shape == (n_tasks, 1400, 1400) and np.isfinite(mean).all()
    per_task_h2d = sum(int(np.asarray(a).nbytes) for a in (tasks[0]["Y_c"][0], tasks[0]["Y_c"][3], tasks[0]["X_c"][3]))
    out["e2e"] = {"value": world * n_tasks / (ms_e2e * 1e-3), "unit": "tasks/s", "h2d_bytes_per_step": per_task_h2d,
                  "d2h_bytes_per_step": 2 * 1400 * 1400 * 4, "ms_per_step": ms_e2e / n_tasks,
                  "first_call_ms_per_step": ms_first / n_tasks,
                  "timed_calls_ms_per_step": [c / n_tasks for c in calls], "reported": "median of the three timed calls",
                  "result_pool": {"hits": _pool.hits, "misses": _pool.misses, "page_locked_arrays": len(_pool._pinned)},
                  "result_arrays": "third call of this size: result arrays recycled from the previous (dropped) result and "
                                   "page-locked, so every read-back is one DMA into the array the caller receives; "
                                   "first_call_ms_per_step is the same call on freshly allocated pageable arrays",
                  "api": "ConvNP.predict(list[Task], X_t=(x1, x2)): staging + H2D of the per-hour sets + forward + D2H of "
                         "mean and std into the result array, all inside the timed region"}
    if profile:
        os.environ["CNP_NO_PREPACK"] = os.environ["CNP_NO_MULTISTREAM"] = "1"
        fwd(0)
        eng.profile_start()
        for i in range(4):
            fwd(i)
        prof = eng.profile_stop()
        del os.environ["CNP_NO_PREPACK"], os.environ["CNP_NO_MULTISTREAM"]
        for v in prof.values():            # per DATE (a launch covers nb dates)
            v["ms"] /= nb
            v["flops"] /= nb
            v["bytes"] /= nb
        pk = peaks()
        name = "cnp_decode_grid_tc_fwd"
        if name in prof:
            d = prof[name]
            tot_ms = sum(v["ms"] for v in prof.values())
            tf = d["flops"] / (d["ms"] * 1e-3) / 1e12
            gbs = d["bytes"] / (d["ms"] * 1e-3) / 1e9
            out["roofline"] = {"kernel": "decode_grid_tc (SetConv decoder + final 1x1 + aux MLP + Gaussian head, tcgen05)",
                               "bound": "tensor", "achieved": tf, "peak": pk["tf"], "unit": "TFLOP/s", "frac": tf / pk["tf"],
                               "hbm_gbs": gbs, "hbm_frac": gbs / pk["hbm"], "traffic": None,
                               "peak_source": f"{pk['source']}: bf16_tflops (burst)", "share_of_step": d["ms"] / tot_ms,
                               "avg_launch_ms": d["ms"] / d["launches"], "flops_per_launch": d["flops"] / d["launches"],
                               "bytes_per_launch": d["bytes"] / d["launches"]}
        out["kernels"] = kernel_table(prof, 4, pk)
    eng.release_workspaces()
    del model, eng
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch.distributed as dist
    world, rank, local = init_dist()
    wl = args.workload
    dim_yc = DIM_YC_MV if wl == "train_mv" else DIM_YC
    if wl == "infer":
        r = infer_arm(args, world, rank, local, max(1, args.steps), args.warmup)
    else:
        r = train_arm(args, world, rank, local, dim_yc, args.steps, args.warmup, check_loss=(wl == "train"))
    grid = r.pop("grid")
    line = {
        "metric": metric_name(wl), "value": r.pop("value"), "unit": "tasks/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r.pop("ms_per_step"), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": workload_config(wl, world, args.precision, grid, dim_yc),
    }
    line.update(r)
    # every default run also carries short records of the other two workloads (configs[3] and configs[2]/[4])
    if wl == "train" and not args.no_sub_records:
        mv = train_arm(args, world, rank, local, DIM_YC_MV, 10, 3, profile=False, e2e=False)
        line["multivar"] = {"workload": "configs[3]: dim_yc=(8,6,1,1), Cin=20, B=16 per GPU, device-resident training step",
                            "value": mv["value"], "unit": "tasks/s", "ms_per_step": mv["ms_per_step"], "steps": 10,
                            "gpu_launches": mv["gpu_launches"]}
        if args.internal_density == PPU:     # the in-repo default density (config.py:2688): 608 x 608 internal grid
            a500 = argparse.Namespace(**vars(args))
            a500.internal_density = 500
            g5 = train_arm(a500, world, rank, local, DIM_YC, 5, 3, profile=False, e2e=False)
            line["grid608"] = {"workload": "configs[1] at internal_density=500 (608 x 608 internal grid), B=16 per GPU, "
                                           "device-resident training step", "value": g5["value"], "unit": "tasks/s",
                               "ms_per_step": g5["ms_per_step"], "steps": 5, "internal_grid": [g5["grid"].n1, g5["grid"].n2]}
        inf = infer_arm(args, world, rank, local, 16, 3, profile=(world == 1))
        line["inference"] = {"workload": "configs[2]/[4]: ConvNP.predict onto 1400x1400, 16 tasks per GPU",
                             "value": inf["value"], "unit": "tasks/s", "ms_per_task": inf["ms_per_step"], "tasks": 16,
                             "dates_per_forward": inf["dates_per_forward"],
                             "e2e": inf["e2e"], "roofline": inf.get("roofline"), "gpu_launches": inf["gpu_launches"]}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            infer = wl == "infer"
            rate, sec, cores = cpu_reference_rate(1, 2, 1, dim_yc, infer)
            what = "on-grid forward onto 1400x1400" if infer else "fwd+NLL+bwd"
            line["cpu_baseline"] = {"value": rate, "unit": "tasks/s", "cores": cores, "kind": "port",
                                    "sample": f"1 task per step ({what}), 2 timed reps after 1 warm-up, "
                                              "torch CPU oracle, all host threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "train_mv", "infer"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-records", action="store_true", help="skip the short multivar / inference sub-records")
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
