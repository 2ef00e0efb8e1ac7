"""Where does the masked WIDE dgrad (64 -> 128 channels, 304^2, B = 16) spend its time?  Per-CTA cycle counters of
conv_tc2 (cnp_conv_tc2_debug) for: no mask / mask / mask + space-to-depth copy, with the epilogue complete (flags 0),
stopped after the TMEM load (1), or without its global stores (2)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deepsensornz_b200 import _cabi  # noqa: E402
from deepsensornz_b200.engine import _Blk  # noqa: E402
from tools.bench_conv import out_blk, timeit, S  # noqa: E402

B, H = 16, 304
dev = torch.device("cuda")
dy = _Blk(B, 8, H, H, dev); dy.t.normal_()
dx = _Blk(B, 16, H, H, dev)
act = _Blk(B, 16, H, H, dev); act.t.normal_()
s2d = _Blk(B, 32, H // 2, H // 2, dev)
wt = torch.randn(64, 128, 5, 5, device=dev) * 0.05
K = _cabi.KIND_K5S1_DGRAD
nb2 = _cabi.lib().cnp_conv_tc2_packed_bytes(K, 8, 128)
wp2 = torch.empty(nb2 // 2, dtype=torch.bfloat16, device=dev)
_cabi.call("cnp_conv_tc2_pack", wt.data_ptr(), 64, 128, 5, K, 8, 0, 0, 0, 128, wp2.data_ptr(), S())
fl = 2.0 * B * H * H * 64 * 128 * 25
for name in ("plain", "mask", "mask(L2-resident: one image shared by the batch)", "mask+s2d"):
    mv = act.view(0)
    if "L2" in name:
        mv.bstride = 0
    o = out_blk(dx.view(0), mask=mv if name != "plain" else None)
    if name == "mask+s2d":
        sv = s2d.view()
        o.s2d, o.s2d_c0, o.s2d_band = C.pointer(sv), 8, 2
    run = lambda: _cabi.call("cnp_conv_tc2", C.byref(dy.view()), 8, wp2.data_ptr(), K, 0, 0, 128, C.byref(o), B, S())
    t = timeit(run)
    print(f"{name:12s}: {t * 1e3:7.1f} us  {fl / t / 1e9:7.1f} TF")
    for flags in (0, 1, 2):
        dbg = torch.zeros(148, 8, dtype=torch.int64, device=dev)
        _cabi.call("cnp_conv_tc2_debug", dbg.data_ptr(), flags)
        run()
        torch.cuda.synchronize()
        _cabi.call("cnp_conv_tc2_debug", None, 0)
        d = dbg.double().mean(0).tolist()
        print(f"   dbg[{flags}] mean/CTA: total {d[0]:.0f} cyc, MMA thread waits: epilogue {d[1]:.0f}, window {d[2]:.0f}, weights {d[3]:.0f}; "
              f"tiles {d[4]:.1f}, epilogue busy {d[5]:.0f}")
