"""Localise a weight-gradient mismatch of the polyphase level: each of the three wgrad launches against torch on the very
tensors the kernels read (strips / s2d read back from the blocked buffers)."""
import ctypes as C
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deepsensornz_b200 import _cabi
from deepsensornz_b200.engine import _Blk
from tests.test_conv_tc2_gpu import _from_blk, _S, _to_blk
from tests.util import rel_err

h, w = (int(v) for v in (sys.argv[1:3] if len(sys.argv) > 2 else (12, 76)))
B, cin = 2, 128
torch.manual_seed(6)
x = torch.randn(B, cin, h, w, device="cuda").relu().bfloat16().float()
dpre = torch.randn(B, 64, 2 * h, 2 * w, device="cuda").bfloat16().float()
xb = _to_blk(x)
u_rows, u_cols = _Blk(2 * B, 16, 6, 2 * w, x.device), _Blk(2 * B, 16, 6, 2 * h, x.device)
_cabi.call("cnp_up_strips_fwd", C.byref(xb.view()), 16, C.byref(u_rows.view()), C.byref(u_cols.view()), B, _S())
dyb = _Blk(B, 16, 2 * h, 2 * w, x.device)
_cabi.call("cnp_blk_from_nchw_f32", dpre.data_ptr(), dpre.stride(0), B, 64, 2 * h, 2 * w, C.byref(dyb.view(8)), _S())
s2d = _Blk(B, 32, h, w, x.device)
dy_rows, dy_cols = _Blk(2 * B, 8, 6, 2 * w, x.device), _Blk(2 * B, 8, 6, 2 * h, x.device)
_cabi.call("cnp_up_dy_split", C.byref(dyb.view(8)), C.byref(s2d.view()), C.byref(dy_rows.view()), C.byref(dy_cols.view()), B, _S())
wsb = _cabi.lib().cnp_conv_tc_wgrad_workspace_bytes()
ws = torch.empty(wsb // 4, device="cuda")

def wg_ref(xs, dys):
    wd = torch.zeros(64, cin, 5, 5, device="cuda", dtype=torch.double, requires_grad=True)
    F.conv2d(xs.double(), wd, padding=2).backward(dys.double())
    return wd.grad

for name, xs, dys, kind in (("rows", u_rows, dy_rows, _cabi.WG_K5S1), ("cols", u_cols, dy_cols, _cabi.WG_K5S1_T)):
    for use_bias in (False, True):
        for use_ws in (True, False):
            g = torch.zeros(64, cin, 5, 5, device="cuda")
            gb = torch.zeros(64, device="cuda")
            _cabi.call("cnp_conv_tc_wgrad", C.byref(xs.view(0)), 16, C.byref(dys.view(0)), kind, g.data_ptr(),
                       gb.data_ptr() if use_bias else None, cin, 2 * B, ws.data_ptr() if use_ws else None, wsb if use_ws else 0, _S())
            ref = wg_ref(_from_blk(xs, cin), _from_blk(dys, 64))
            if kind == _cabi.WG_K5S1_T:
                ref = ref.transpose(2, 3)
            print(name, "bias", use_bias, "ws", use_ws, "rel_err", rel_err(g, ref),
                  "bias err", rel_err(gb, _from_blk(dys, 64).double().sum((0, 2, 3))) if use_bias else None)

# interior part: phase wgrad on (x, s2d) + fold, against the 5x5 weight gradient of the band-zeroed dy on the upsampled x
up = F.interpolate(x.double(), scale_factor=2, mode="bilinear", align_corners=False)
inner = torch.zeros_like(dpre)
inner[:, :, 4:-4, 4:-4] = dpre[:, :, 4:-4, 4:-4]
ref_int = wg_ref(up, inner)
dwp = torch.zeros(2, 2, 64, cin, 4, 4, device="cuda")
gb = torch.zeros(64, device="cuda")
_cabi.call("cnp_conv_tc_wgrad", C.byref(xb.view(0)), 16, C.byref(s2d.view(0)), _cabi.WG_UP_PHASE, dwp.data_ptr(), gb.data_ptr(),
           cin, B, ws.data_ptr(), wsb, _S())
g = torch.zeros(64, cin, 5, 5, device="cuda")
_cabi.call("cnp_up_wgrad_fold", dwp.data_ptr(), None, 64, cin, g.data_ptr(), _S())
print("interior: fold(dwp) vs 5x5 wgrad of inner dy", rel_err(g, ref_int), "bias", rel_err(gb, inner.double().sum((0, 2, 3))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import polyphase_check as P
xz = F.pad(x.double(), (2, 2, 2, 2))
dwp_ref = torch.zeros_like(dwp, dtype=torch.double)
for a in (0, 1):
    for b in (0, 1):
        win = xz[:, :, a:a + h + 3, b:b + w + 3]
        dwp_ref[a, b] = F.conv2d(win.transpose(0, 1), inner[:, :, a::2, b::2].double().transpose(0, 1)).transpose(0, 1)
print("dwp vs ref", rel_err(dwp, dwp_ref))
fold = P.fold_matrix(torch.float64).cuda()
print("fold(dwp_ref) vs ref_int", rel_err(torch.einsum("aboipq,akp,blq->oikl", dwp_ref, fold, fold), ref_int))
s2 = _from_blk(s2d, 256)
for a in (0, 1):
    for b in (0, 1):
        ph = (a * 2 + b) * 64
        print("s2d phase", a, b, torch.equal(s2[:, ph:ph + 64], inner[:, :, a::2, b::2]))
