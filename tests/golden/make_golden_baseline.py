"""Golden vectors at the BASELINE shapes from the CPU oracle.

    python tests/golden/make_golden_baseline.py            # all cases (~2 min on 8 cores)

PARITY UNPINNED: the reference's arithmetic packages are not importable here and it holds no vectors for this
path, so these files pin the ORACLE at the sizes the benchmark runs -- internal_density 250 (304 x 304
internal grid), 1400 x 1400 land-mask context, 160 context / 40 target stations, 1400 x 1400 on-grid targets, and the
8-channel multi-variable base grid (Cin = 20) -- with He-scaled weights (tests/util.sensitive_state) so that every
UNet layer moves the outputs.  Inputs are regenerated from seeds on the GPU box (deepsensornz_b200.synthetic); the
oracle's inputs come from oracle/task_tensors.py, not from the product's concat_tasks / modify_task.

Per case: mean, var, per-task logp, loss, the internal grid, per-channel sums of the encoder output, a strided
sample of the UNet output z (a bf16 intermediate the loss alone would not pin), for every parameter the gradient
norm and 8 fixed +-1 projections, and three gradient tensors in full.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from deepsensornz_b200.synthetic import make_static, make_task  # noqa: E402
from oracle import convnp_oracle as O  # noqa: E402
from oracle.task_tensors import task_tensors  # noqa: E402
from tests.util import baseline_model, cpu_params, grad_probes  # noqa: E402

PPU = 250
CASES = {
    # name: dict(n tasks, first seed, c0 channels, on-grid targets)
    "s1_single": dict(nb=1, seed=20160101, c0=3, grid=False),
    "s2_batch4": dict(nb=4, seed=20160101, c0=3, grid=False),
    "s3_grid1400": dict(nb=1, seed=20160101, c0=3, grid=True),
    "s4_multivar8": dict(nb=2, seed=20160101, c0=8, grid=False),
}
FULL_GRADS = ("decoder.unet.before_turn_layers.3.weight", "decoder.unet.final_linear.weight",
              "decoder.mlp.layers.0.weight")
Z_SAMPLE = (slice(None), slice(None, None, 8), slice(None, None, 16), slice(None, None, 16))
GRID_SAMPLE = 7      # s3: every 7th target row / column is stored (200 x 200 of 1400 x 1400)

_static = {}


def static_fields():
    if "s" not in _static:
        _static["s"] = make_static(seed=7, n_hi=1400)
    return _static["s"]


def build_tasks(name):
    c = CASES[name]
    st = static_fields()
    return [make_task(st, c["seed"] + i, n_stations=200, context_frac=0.8, c0_channels=c["c0"],
                      grid_targets=c["grid"]) for i in range(c["nb"])]


def build_model(name, precision="fp32"):
    return baseline_model(precision, dim_yc=(CASES[name]["c0"], 6, 1, 1), ppu=PPU, seed=1)


def bench_loss0():
    """fp32 oracle loss of bench.py's first batch (16 tasks, torch default init under manual_seed(0)): the number
    bench.py checks its own first forward against."""
    import bench
    torch.manual_seed(0)
    from deepsensornz_b200 import ConvNP
    m = ConvNP(**bench.model_kwargs())
    tasks = bench.make_task_lists(1, 0)[0]
    ctx, xt, yt, aux = task_tensors(tasks)
    with torch.no_grad():
        return float(O.loss_fn(cpu_params(m), ctx, xt, yt, aux, PPU))


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    torch.set_num_threads(os.cpu_count() or 1)
    only = sys.argv[1:]
    for name, c in CASES.items():
        if only and name not in only:
            continue
        t0 = time.time()
        m = build_model(name)
        ctx, xt, yt, aux = task_tensors(build_tasks(name))
        P = {k: v.clone().requires_grad_(v.dim() > 0 and not c["grid"]) for k, v in cpu_params(m).items()}
        with torch.set_grad_enabled(not c["grid"]):
            mean, var, info = O.forward(P, ctx, xt, aux, PPU, return_internal=True)
        g = info["grid"]
        rec = dict(grid=np.array([g[0][0], g[0][1], g[1][0], g[1][1], g[2]], dtype=np.float64),
                   enc_sum=info["enc"].double().sum(dim=(0, 2, 3)).numpy(),
                   z_sample=info["z"].detach()[Z_SAMPLE].numpy().copy(),
                   z_abs_mean=np.float64(info["z"].detach().abs().mean()))
        if c["grid"]:
            std = var.sqrt()
            rec.update(mean_sample=mean[0, 0, ::GRID_SAMPLE, ::GRID_SAMPLE].numpy().copy(),
                       std_sample=std[0, 0, ::GRID_SAMPLE, ::GRID_SAMPLE].numpy().copy(),
                       mean_rowsum=mean[0, 0].double().sum(dim=1).numpy(), mean_colsum=mean[0, 0].double().sum(dim=0).numpy(),
                       std_rowsum=std[0, 0].double().sum(dim=1).numpy(), std_colsum=std[0, 0].double().sum(dim=0).numpy(),
                       mean_absmax=np.float64(mean.abs().max()), std_max=np.float64(std.max()))
            print(name, "grid", g, "mean|.|max", float(mean.abs().max()), f"{time.time() - t0:.1f}s")
        else:
            logp = O.loglik(mean, var, yt, True)
            loss = -logp.mean()
            loss.backward()
            names = sorted(k for k, v in P.items() if v.grad is not None)
            norms = np.array([float(P[k].grad.double().norm()) for k in names])
            probes = np.stack([(grad_probes(k, P[k].numel()) @ P[k].grad.double().flatten()).numpy() for k in names])
            rec.update(mean=mean.detach().numpy(), var=var.detach().numpy(), logp=logp.detach().numpy(),
                       loss=np.float64(loss.detach()), grad_names=np.array(names), grad_norms=norms, grad_probes=probes)
            for k in FULL_GRADS:
                rec["grad." + k] = P[k].grad.numpy()
            # how much the loss depends on the UNet: zero one up-path layer and recompute
            with torch.no_grad():
                Q = {k: v.detach() for k, v in P.items()}
                Q["decoder.unet.after_turn_layers.0.weight"] = torch.zeros_like(Q["decoder.unet.after_turn_layers.0.weight"])
                rec["loss_after0_zeroed"] = np.float64(O.loss_fn(Q, ctx, xt, yt, aux, PPU))
            print(name, "loss", float(loss), "(after0 zeroed:", float(rec["loss_after0_zeroed"]), ") grid", g,
                  f"{time.time() - t0:.1f}s")
        np.savez_compressed(os.path.join(out, name + ".npz"), **rec)
    if not only or "bench" in only:
        t0 = time.time()
        l0 = bench_loss0()
        np.savez(os.path.join(out, "s2_bench16_loss0.npz"), loss=np.float64(l0))
        print("s2_bench16_loss0", l0, f"{time.time() - t0:.1f}s")


if __name__ == "__main__":
    main()
