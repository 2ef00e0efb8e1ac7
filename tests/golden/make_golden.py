"""Generate tests/golden/*.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference ships no golden vectors for this path and its arithmetic packages are not importable here
 -> these vectors pin the ORACLE (seeded synthetic tasks + seeded random-init
weights), so that the oracle cannot drift silently and the GPU box (which has no /root/reference and need
not re-run the oracle) can check the CUDA path against committed numbers.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from deepsensornz_b200 import concat_tasks  # noqa: E402
from deepsensornz_b200.synthetic import make_static, make_task  # noqa: E402
from oracle import convnp_oracle as O  # noqa: E402
from tests.util import cpu_params, oracle_inputs, small_model  # noqa: E402

CASES = {
    # name: (n tasks, first seed, ppu, n_stations, c0 channels)
    "g1_single": (1, 4100, 50, 200, 3),
    "g2_batch3": (3, 4200, 50, 200, 3),
    "g3_multivar": (2, 4300, 40, 120, 6),
}


def build(name):
    nb, seed, ppu, nst, c0 = CASES[name]
    static = make_static(seed=7, n_hi=200)
    tasks = [make_task(static, seed + i, n_stations=nst, c0_channels=c0) for i in range(nb)]
    task = concat_tasks(tasks) if nb > 1 else tasks[0]
    m = small_model("fp32", ppu=ppu, dim_yc=(c0, 6, 1, 1), seed=1234)
    return m, task


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    for name in CASES:
        m, task = build(name)
        ctx, xt, yt, aux = oracle_inputs(task)
        P = {k: v.clone().requires_grad_(v.dim() > 0) for k, v in cpu_params(m).items()}
        mean, var, info = O.forward(P, ctx, xt, aux, m.config.points_per_unit, return_internal=True)
        logp = O.loglik(mean, var, yt, True)
        loss = -logp.mean()
        loss.backward()
        gn = {k: float(v.grad.double().norm()) for k, v in P.items() if v.grad is not None}
        np.savez_compressed(
            os.path.join(out, name + ".npz"), mean=mean.detach().numpy(), var=var.detach().numpy(),
            logp=logp.detach().numpy(), loss=np.float64(loss.detach()), enc_sum=info["enc"].double().sum(dim=(0, 2, 3)).numpy(),
            z_abs_mean=np.float64(info["z"].abs().mean()), grid=np.array([info["grid"][0][0], info["grid"][0][1],
                                                                             info["grid"][1][0], info["grid"][1][1],
                                                                             info["grid"][2]], dtype=np.float64),
            grad_names=np.array(sorted(gn)), grad_norms=np.array([gn[k] for k in sorted(gn)]),
            d_final_bias=P["decoder.unet.final_linear.bias"].grad.numpy(),
            d_mlp0=P["decoder.mlp.layers.0.weight"].grad.numpy())
        print(name, "loss", float(loss), "grid", info["grid"])


if __name__ == "__main__":
    main()
