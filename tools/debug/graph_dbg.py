import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
from deepsensornz_b200 import concat_tasks
from deepsensornz_b200.graph import GraphedTrainStep
from deepsensornz_b200.synthetic import make_static, make_task
from tests.util import small_model
static = make_static(seed=7, n_hi=200, with_aux_hi=True)
groups = [[make_task(static, 700 + 4 * k + i) for i in range(4)] for k in range(3)]
def snap(m): return torch.cat([p.detach().flatten().clone() for p in m.model.parameters()])
def run(variant, lr=2e-3, fused=True):
    m = small_model("bf16", seed=3)
    opt = torch.optim.AdamW(m.model.parameters(), lr=lr, fused=fused, capturable=True)
    dev = [m._to_device(concat_tasks(g)) for g in groups]
    def eager(b):
        opt.zero_grad(set_to_none=True); loss = m.loss_fn(b, normalise=True); loss.backward(); opt.step(); return float(loss.detach())
    def val(b):
        with torch.no_grad(): return float(m.loss_fn(b, normalise=True))
    out = [eager(dev[0])]
    w1 = snap(m)
    steps = [float(s["step"]) for s in list(opt.state.values())[:2]]
    gs = GraphedTrainStep(m, opt, dev[0], warm=True)
    w1c = snap(m)
    out.append(("changed_by_capture", float((w1c - w1).abs().max()), "steps", steps, [float(s["step"]) for s in list(opt.state.values())[:2]]))
    out.append(("val", val(dev[1])))
    l = float(gs.step(dev[1])); w2 = snap(m)
    out.append(("replay1", l, "dW", float((w2 - w1).abs().max()), [float(s["step"]) for s in list(opt.state.values())[:2]]))
    out.append(("val2", val(dev[2])))
    l = float(gs.step(dev[2])); out.append(("replay2", l))
    print(variant, out, flush=True)
run("fused"); run("foreach", fused=False)
