"""Times cnp_conv_tc_wgrad at the bench shapes (B=16) with / without the bias gradient and the workspace path."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepsensornz_b200 import _cabi  # noqa: E402
from deepsensornz_b200.engine import _Blk  # noqa: E402

S = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    B = 16
    dev = torch.device("cuda")
    wsb = _cabi.lib().cnp_conv_tc_wgrad_workspace_bytes()
    ws = torch.empty(wsb // 4, device=dev)
    for cin, H in [(128, 304), (64, 304), (128, 152), (64, 76), (64, 38)]:
        x = _Blk(B, cin // 8, H, H, dev)
        x.t.normal_()
        dy = _Blk(B, 8, H, H, dev)
        dy.t.normal_()
        # zero the pads (the kernel relies on them)
        for blk, cb in ((x, cin // 8), (dy, 8)):
            v = blk.t[:B * blk.bstride].view(B, cb, H + 4, H + 4, 8)
            v[:, :, :2] = 0; v[:, :, -2:] = 0; v[:, :, :, :2] = 0; v[:, :, :, -2:] = 0
        dw = torch.zeros(64, cin, 5, 5, device=dev)
        db = torch.zeros(64, device=dev)
        fl = 2.0 * B * H * H * 64 * cin * 25
        res = []
        for bias in (False, True):
            for use_ws in (False, True):
                t = timeit(lambda: _cabi.call("cnp_conv_tc_wgrad", C.byref(x.view()), cin // 8, C.byref(dy.view()), _cabi.WG_K5S1,
                                              dw.data_ptr(), db.data_ptr() if bias else None, cin, B,
                                              ws.data_ptr() if use_ws else None, wsb if use_ws else 0, S()))
                res.append(f"bias={int(bias)} ws={int(use_ws)}: {t*1e3:7.1f} us {fl/t/1e9:7.1f} TF")
        print(f"wgrad {cin:3d}ch {H}^2 | " + " | ".join(res), flush=True)


def narrow():
    """the folded first layer: 16-channel [x ; 1] input, WG_K5S1_NARROW"""
    B, H, dev = 16, 304, torch.device("cuda")
    wsb = _cabi.lib().cnp_conv_tc_wgrad_workspace_bytes()
    ws = torch.empty(wsb // 4, device=dev)
    for nch in (2, 4):
        x = _Blk(B, nch, H, H, dev)
        x.t.normal_()
        dy = _Blk(B, 8, H, H, dev)
        dy.t.normal_()
        for blk, cb in ((x, nch), (dy, 8)):
            v = blk.t[:B * blk.bstride].view(B, cb, H + 4, H + 4, 8)
            v[:, :, :2] = 0; v[:, :, -2:] = 0; v[:, :, :, :2] = 0; v[:, :, :, -2:] = 0
        dw = torch.zeros(64, nch * 8, 5, 5, device=dev)
        db = torch.zeros(64, device=dev)
        t = timeit(lambda: _cabi.call("cnp_conv_tc_wgrad", C.byref(x.view()), nch, C.byref(dy.view()), _cabi.WG_K5S1_NARROW,
                                      dw.data_ptr(), db.data_ptr(), nch * 8, B, ws.data_ptr(), wsb, S()))
        print(f"narrow wgrad {nch * 8:3d}ch {H}^2: {t * 1e3:7.1f} us", flush=True)


if __name__ == "__main__":
    narrow()
    main()
