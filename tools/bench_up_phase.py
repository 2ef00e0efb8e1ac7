"""Times the up-phase kind (polyphase resize-convolution building block) at the bench shape: 128 -> 64 channels, low-res
152^2 -> high-res 304^2, B = 16, against the regular 5x5 convolution of the materialised upsampled tensor."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepsensornz_b200 import _cabi  # noqa: E402
from deepsensornz_b200.engine import _Blk  # noqa: E402

S = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def out_blk(view, bias, scatter):
    o = _cabi.CnpConvOut()
    o.mode, o.blk = 0, view
    o.sy, o.ay, o.sx, o.ax = scatter
    o.bias, o.relu, o.accumulate = bias.data_ptr(), 1, 0
    return o


def main():
    B, cin, h, w = 16, 128, 152, 152
    dev = torch.device("cuda")
    w5 = torch.randn(64, cin, 5, 5, device=dev) * 0.05
    bias = torch.randn(64, device=dev)
    lo = _Blk(B, cin // 8, h, w, dev); lo.t.normal_()
    hi = _Blk(B, cin // 8, 2 * h, 2 * w, dev); hi.t.normal_()
    y = _Blk(B, 8, 2 * h, 2 * w, dev)
    wp = torch.empty(2, 2, 64, cin, 4, 4, device=dev)
    _cabi.call("cnp_up_phase_weights", w5.data_ptr(), 64, cin, wp.data_ptr(), S())
    packs = []
    for a in (0, 1):
        nb = _cabi.lib().cnp_conv_tc2_packed_bytes(_cabi.KIND_UP_PHASE, cin // 8, 64)
        pk = torch.empty(nb // 2, dtype=torch.bfloat16, device=dev)
        wa = wp[a].contiguous()
        _cabi.call("cnp_conv_tc2_pack", wa.data_ptr(), 64, cin, 4, _cabi.KIND_UP_PHASE, cin // 8, a, 0, 0, 64, pk.data_ptr(), S())
        packs.append(pk)
    nb = _cabi.lib().cnp_conv_tc2_packed_bytes(_cabi.KIND_K5S1, cin // 8, 64)
    pk5 = torch.empty(nb // 2, dtype=torch.bfloat16, device=dev)
    _cabi.call("cnp_conv_tc2_pack", w5.data_ptr(), 64, cin, 5, _cabi.KIND_K5S1, cin // 8, 0, 0, 0, 64, pk5.data_ptr(), S())
    outs = [out_blk(y.view(0), bias, (2, a, 2, 0)) for a in (0, 1)]
    o5 = out_blk(y.view(0), bias, (1, 0, 1, 0))

    def phases():
        for a in (0, 1):
            _cabi.call("cnp_conv_tc2", C.byref(lo.view()), cin // 8, packs[a].data_ptr(), _cabi.KIND_UP_PHASE, a, 0, 64,
                       C.byref(outs[a]), B, S())

    def regular():
        _cabi.call("cnp_conv_tc2", C.byref(hi.view()), cin // 8, pk5.data_ptr(), _cabi.KIND_K5S1, 0, 0, 64, C.byref(o5), B, S())

    def upsample():
        _cabi.call("cnp_blk_upsample2x_fwd", C.byref(lo.view()), cin // 8, C.byref(hi.view()), B, S())

    t_p, t_r, t_u = timeit(phases), timeit(regular), timeit(upsample)
    print(f"128 -> 64, 152^2 -> 304^2, B = 16: two up-phase launches {t_p:.1f} us | regular 5x5 on the upsampled tensor {t_r:.1f} us "
          f"(+ upsampling kernel {t_u:.1f} us)")


if __name__ == "__main__":
    main()
