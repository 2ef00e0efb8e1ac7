"""Shared test helpers: numpy Task -> oracle inputs, small model factory."""
import numpy as np
import torch

from deepsensornz_b200 import ConvNP, Masked, Task
from deepsensornz_b200.task import convert_task_to_nps_args


def oracle_inputs(task: Task):
    """Apply the reference's host-side task ops and hand torch CPU tensors to the oracle."""
    t = ConvNP.modify_task(task)
    ctx, xt, yt, kw = convert_task_to_nps_args(t)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    contexts = []
    for x, y in ctx:
        xx = tuple(T(v) for v in x) if isinstance(x, tuple) else T(x)
        if isinstance(y, Masked):
            contexts.append((xx, T(y.y), T(y.mask)))
        else:
            contexts.append((xx, T(y), None))
    xt = tuple(T(v) for v in xt) if isinstance(xt, tuple) else T(xt)
    yt = T(np.asarray(yt)) if yt is not None else None
    aux = T(kw["aux_t"]) if "aux_t" in kw else None
    return contexts, xt, yt, aux


def small_model(precision="fp32", ppu=50, dim_yc=(3, 6, 1, 1), seed=0, **kw):
    torch.manual_seed(seed)
    n_lo = kw.pop("n_lo", 140)
    scales = (0.5 / (n_lo - 1) * 0.99286, 0.5 / (n_lo - 1) * 0.99286, 0.5 / 199.0, 0.5 / ppu)
    return ConvNP(dim_yc=dim_yc, dim_yt=1, dim_aux_t=5, internal_density=ppu, encoder_scales=scales,
                  decoder_scale=1.0 / ppu, precision=precision, verbose=False, **kw)


def cpu_params(model):
    return {k: v.detach().cpu().clone() for k, v in model.model.state_dict().items()}


def rel_err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))
