"""Times the tcgen05 convolution kernels at the bench shapes (B=16, internal grid 304): old (pixels = M) vs
new (pixels = N) formulation.  python tools/bench_conv.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepsensornz_b200 import _cabi  # noqa: E402
from deepsensornz_b200.engine import _Blk  # noqa: E402

S = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731


def out_blk(view, bias=None, relu=0, mask=None, accumulate=0):
    o = _cabi.CnpConvOut()
    o.mode, o.blk = 0, view
    o.sy, o.ay, o.sx, o.ax = 1, 0, 1, 0
    o.bias, o.relu, o.accumulate = (bias.data_ptr() if bias is not None else None), relu, accumulate
    o.mask = C.pointer(mask) if mask is not None else None
    o._keep = (bias, mask)
    return o


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
    have_old = hasattr(_cabi.lib(), "cnp_conv_tc")
    for (cin, H) in [(128, 304), (64, 304), (128, 152), (64, 152), (64, 76)]:
        x = _Blk(B, cin // 8, H, H, dev)
        x.t.normal_()
        y = _Blk(B, 8, H, H, dev)
        wt = torch.randn(64, cin, 5, 5, device=dev) * 0.05
        bias = torch.randn(64, device=dev)
        fl = 2.0 * B * H * H * 64 * cin * 25
        o = out_blk(y.view(), bias, 1)
        t_old = float("nan")
        if have_old:
            nb = _cabi.lib().cnp_conv_tc_packed_bytes(_cabi.KIND_K5S1, cin // 8)
            wpk = torch.empty(nb // 2, dtype=torch.bfloat16, device=dev)
            _cabi.call("cnp_conv_tc_pack", wt.data_ptr(), 64, cin, 5, _cabi.KIND_K5S1, cin // 8, 0, 0, 0, wpk.data_ptr(), S())
            t_old = timeit(lambda: _cabi.call("cnp_conv_tc", C.byref(x.view()), cin // 8, wpk.data_ptr(), _cabi.KIND_K5S1,
                                              0, 0, C.byref(o), B, S()))
        nb2 = _cabi.lib().cnp_conv_tc2_packed_bytes(_cabi.KIND_K5S1, cin // 8, 64)
        wpk2 = torch.empty(nb2 // 2, dtype=torch.bfloat16, device=dev)
        _cabi.call("cnp_conv_tc2_pack", wt.data_ptr(), 64, cin, 5, _cabi.KIND_K5S1, cin // 8, 0, 0, 0, 64, wpk2.data_ptr(), S())
        t_new = timeit(lambda: _cabi.call("cnp_conv_tc2", C.byref(x.view()), cin // 8, wpk2.data_ptr(), _cabi.KIND_K5S1, 0, 0,
                                          64, C.byref(o), B, S()))
        for flags in (0, 1, 2):
            dbg = torch.zeros(148, 8, dtype=torch.int64, device=dev)
            _cabi.call("cnp_conv_tc2_debug", dbg.data_ptr(), flags)
            _cabi.call("cnp_conv_tc2", C.byref(x.view()), cin // 8, wpk2.data_ptr(), _cabi.KIND_K5S1, 0, 0, 64, C.byref(o), B, S())
            torch.cuda.synchronize()
            _cabi.call("cnp_conv_tc2_debug", None, 0)
            d = dbg.double().mean(0).tolist()
            print(f"   dbg[{flags}] mean/CTA: total {d[0]:.0f} cyc, wait epilogue {d[1]:.0f}, window {d[2]:.0f}, weights {d[3]:.0f}, "
                  f"tiles {d[4]:.1f}, epilogue busy {d[5]:.0f}")
        print(f"fwd  {cin:3d}->64 {H}^2: old {t_old*1e3:7.1f} us {fl/t_old/1e9:7.1f} TF | new {t_new*1e3:7.1f} us "
              f"{fl/t_new/1e9:7.1f} TF", flush=True)
    # final 1x1 conv with fp32 NCHW output
    H = 304
    x = _Blk(B, 8, H, H, dev)
    x.t.normal_()
    z = torch.empty(B, 64, H, H, device=dev)
    w1 = torch.randn(64, 64, 1, 1, device=dev) * 0.1
    bias = torch.randn(64, device=dev)
    o = _cabi.CnpConvOut()
    o.mode, o.f32, o.f32_bstride, o.f32_ch_off = 1, z.data_ptr(), z.stride(0), 0
    o.sy, o.ay, o.sx, o.ax = 1, 0, 1, 0
    o.bias = bias.data_ptr()
    nb2 = _cabi.lib().cnp_conv_tc2_packed_bytes(_cabi.KIND_K1, 8, 64)
    wpk2 = torch.empty(nb2 // 2, dtype=torch.bfloat16, device=dev)
    _cabi.call("cnp_conv_tc2_pack", w1.data_ptr(), 64, 64, 1, _cabi.KIND_K1, 8, 0, 0, 0, 64, wpk2.data_ptr(), S())
    for flags in (0, 1):
        dbg = torch.zeros(148, 8, dtype=torch.int64, device=dev)
        _cabi.call("cnp_conv_tc2_debug", dbg.data_ptr(), flags)
        _cabi.call("cnp_conv_tc2", C.byref(x.view()), 8, wpk2.data_ptr(), _cabi.KIND_K1, 0, 0, 64, C.byref(o), B, S())
        torch.cuda.synchronize()
        _cabi.call("cnp_conv_tc2_debug", None, 0)
        d = dbg.double().mean(0).tolist()
        print(f"   dbg[{flags}] mean/CTA: total {d[0]:.0f} cyc, wait epilogue {d[1]:.0f}, window {d[2]:.0f}, weights {d[3]:.0f}, "
              f"tiles {d[4]:.1f}, epilogue busy {d[5]:.0f}")
    t_new = timeit(lambda: _cabi.call("cnp_conv_tc2", C.byref(x.view()), 8, wpk2.data_ptr(), _cabi.KIND_K1, 0, 0, 64, C.byref(o), B, S()))
    print(f"1x1 64->64 fp32 out {H}^2: new {t_new*1e3:7.1f} us  {(z.numel()*4 + B*64*H*H*2)/t_new/1e6:7.1f} GB/s", flush=True)
    # dgrad of a 128->64 layer: dy 64 ch -> dx 128 ch (old: two 64-channel launches; new: one WIDE launch)
    for H in (304, 152):
        dy = _Blk(B, 8, H, H, dev)
        dy.t.normal_()
        dx = _Blk(B, 16, H, H, dev)
        act = _Blk(B, 16, H, H, dev)
        act.t.normal_()
        wt = torch.randn(64, 128, 5, 5, device=dev) * 0.05
        fl = 2.0 * B * H * H * 64 * 128 * 25
        K = _cabi.KIND_K5S1_DGRAD
        t_old = float("nan")
        if have_old:
            nb = _cabi.lib().cnp_conv_tc_packed_bytes(K, 8)
            wp = [torch.empty(nb // 2, dtype=torch.bfloat16, device=dev) for _ in range(2)]
            for g in range(2):
                _cabi.call("cnp_conv_tc_pack", wt.data_ptr(), 64, 128, 5, K, 8, 0, 0, 64 * g, wp[g].data_ptr(), S())
            outs = [out_blk(dx.view(8 * g), mask=act.view(8 * g)) for g in range(2)]

            def old():
                for g in range(2):
                    _cabi.call("cnp_conv_tc", C.byref(dy.view()), 8, wp[g].data_ptr(), K, 0, 0, C.byref(outs[g]), B, S())
            t_old = timeit(old)
        nb2 = _cabi.lib().cnp_conv_tc2_packed_bytes(K, 8, 128)
        wp2 = torch.empty(nb2 // 2, dtype=torch.bfloat16, device=dev)
        _cabi.call("cnp_conv_tc2_pack", wt.data_ptr(), 64, 128, 5, K, 8, 0, 0, 0, 128, wp2.data_ptr(), S())
        ow = out_blk(dx.view(0), mask=act.view(0))
        t_new = timeit(lambda: _cabi.call("cnp_conv_tc2", C.byref(dy.view()), 8, wp2.data_ptr(), K, 0, 0, 128, C.byref(ow), B, S()))
        print(f"dgrad 64->128 {H}^2: old {t_old*1e3:7.1f} us {fl/t_old/1e9:7.1f} TF | new {t_new*1e3:7.1f} us "
              f"{fl/t_new/1e9:7.1f} TF", flush=True)


if __name__ == "__main__":
    main()
