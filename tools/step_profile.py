"""Per-launch CUDA-event timing of one bench training step (every C-ABI call, in order)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from deepsensornz_b200 import ConvNP  # noqa: E402


def main():
    torch.manual_seed(0)
    model = ConvNP(precision="bf16", **bench.model_kwargs())
    opt = torch.optim.AdamW(model.model.parameters(), lr=5e-5)
    eng = model.engine
    dev = [eng.upload(model.stage_task(t, pinned=True)) for t in bench.make_batches(1, 0)]

    def step():
        opt.zero_grad(set_to_none=True)
        loss = model.loss_fn(dev[0], normalise=True)
        loss.backward()
        opt.step()
    for _ in range(3):
        step()
    eng.profile_start()
    step()
    torch.cuda.synchronize()
    prof = eng._prof
    eng._prof = None
    tot = 0.0
    for name, e0, e1, (fl, by) in prof:
        ms = e0.elapsed_time(e1)
        tot += ms
        extra = f"{fl / ms / 1e9:8.1f} TF" if fl else (f"{by / ms / 1e6:8.1f} GB/s" if by else "")
        print(f"{name:32s} {ms * 1e3:9.1f} us {extra}")
    print(f"total {tot:.3f} ms")


if __name__ == "__main__":
    main()
