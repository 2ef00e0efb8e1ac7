"""The decomposition the engine's polyphase resize-convolution uses, stated in float64 torch on the CPU
with the same data flow as the kernels, and checked against autograd of

    y = conv5x5(zero_pad_2(bilinear_up2x(x)))      (torch.nn.Upsample(scale_factor=2, mode="bilinear") + Conv2d(k=5, padding=2)).

OUTPUT PARTITION ("mask form").  A low-resolution pixel (i, j) owns the 2x2 high-resolution outputs (2i+a, 2j+b).
* interior  i in [2, H-3], j in [2, W-3]: the four 4x4 phase convolutions of x read rows i-2 .. i+2 -- inside the image, so
  neither the clamping of the bilinear upsampling nor the zero padding of the upsampled tensor is ever seen: the
  phase kernels run on the plain blocked tensor (zero ring) and are EXACT there;
* the band of 2 low-res (= 4 high-res) pixels along every edge comes from the standard 5x5 kernels run on STRIPS of the
  upsampled tensor: two row strips (top, bottom: 6 high-res rows each, the full width) and two column strips (left,
  right) stored TRANSPOSED (strip row = high-res column) and convolved with the transposed weights, so that all four are
  ordinary row-major images of height 6.
Backward: the same partition of dy.  Every dy pixel is used exactly once: the phase kernels see dy with the 4-pixel band
zeroed (space-to-depth copy), the row strips carry the band rows, the column strips the band columns WITHOUT the corner
squares.  The strips' input gradient is a gradient w.r.t. the upsampled tensor (6 rows: 4 + the 2-pixel reach of the 5x5)
and is folded through the transposed bilinear upsampling onto the 4 outermost low-res rows / columns.

python tools/polyphase_strips.py prints the errors for a few shapes; tests/test_polyphase_math.py runs it.
"""
import torch
import torch.nn.functional as F

from polyphase_check import fold_matrix, phase_weights

SB = 2          # band width in low-res pixels
SR = 6          # strip height in high-res pixels: 2*SB outputs + the 2-pixel reach of the 5x5 kernel


def up2(x):
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)


def strips_of(u: torch.Tensor, hi_band: int, corners: bool = True):
    """u [B,C,2H,2W] -> (rows [2,B,C,SR,2W], cols [2,B,C,SR,2H]): top / bottom and left^T / right^T strips holding the
    outermost SR rows / columns, zero outside ``hi_band`` rows of the edge; ``corners=False`` zeroes the column strips'
    first and last ``hi_band`` entries (the corner squares belong to the row strips)."""
    B, C, H2, W2 = u.shape
    rows = torch.stack([u[:, :, :SR], u[:, :, H2 - SR:]]).clone()
    ut = u.transpose(2, 3)
    cols = torch.stack([ut[:, :, :SR], ut[:, :, W2 - SR:]]).clone()
    if hi_band < SR:
        rows[0][:, :, hi_band:] = 0; rows[1][:, :, :SR - hi_band] = 0
        cols[0][:, :, hi_band:] = 0; cols[1][:, :, :SR - hi_band] = 0
    if not corners:
        cols[:, :, :, :, :hi_band] = 0; cols[:, :, :, :, H2 - hi_band:] = 0
    return rows, cols


def forward(x: torch.Tensor, w5: torch.Tensor) -> torch.Tensor:
    B, C, H, W = x.shape
    assert H >= 2 * SB + 1 and W >= 2 * SB + 1
    Co = w5.shape[0]
    wp = phase_weights(w5)
    xz = F.pad(x, (2, 2, 2, 2))                                   # the blocked tensor's own ZERO ring
    y = x.new_zeros(B, Co, 2 * H, 2 * W)
    for a in range(2):
        for b in range(2):
            y[:, :, a::2, b::2] = F.conv2d(xz[:, :, a:a + H + 3, b:b + W + 3], wp[a, b])   # wrong in the band, overwritten
    rows, cols = strips_of(up2(x), SR)
    w5t = w5.transpose(2, 3)
    for s in range(2):
        o = F.conv2d(rows[s], w5, padding=2)                      # [B,Co,SR,2W]; 4 of the 6 rows are band outputs
        if s == 0: y[:, :, :2 * SB] = o[:, :, :2 * SB]
        else: y[:, :, 2 * H - 2 * SB:] = o[:, :, SR - 2 * SB:]
        o = F.conv2d(cols[s], w5t, padding=2)                     # [B,Co,SR,2H] transposed
        if s == 0: y[:, :, :, :2 * SB] = o[:, :, :2 * SB].transpose(2, 3)
        else: y[:, :, :, 2 * W - 2 * SB:] = o[:, :, SR - 2 * SB:].transpose(2, 3)
    return y


def up2_transposed_band(dU_rows, dU_cols, H, W):
    """Gather form of the transposed bilinear upsampling restricted to gradients that live in the strips:
    dU_rows [2,B,C,SR,2W] (high-res rows 0..5 / 2H-6..2H-1), dU_cols [2,B,C,SR,2H] (transposed columns).  Returns the
    low-res gradient [B,C,H,W] (non-zero only within 4 pixels of the border)."""
    B, C = dU_rows.shape[1:3]
    dU = dU_rows.new_zeros(B, C, 2 * H, 2 * W)                     # the kernel never builds this: it gathers from the strips
    dU[:, :, :SR] += dU_rows[0]; dU[:, :, 2 * H - SR:] += dU_rows[1]
    dU[:, :, :, :SR] += dU_cols[0].transpose(2, 3); dU[:, :, :, 2 * W - SR:] += dU_cols[1].transpose(2, 3)

    def taps(r, n):       # high-res indices and weights that feed low-res index r (n low-res pixels)
        out = []
        for Y, wgt in ((2 * r - 1, 0.25), (2 * r, 0.75), (2 * r + 1, 0.75), (2 * r + 2, 0.25)):
            if 0 <= Y < 2 * n:
                out.append([Y, wgt])
        if r == 0: out[0][1] = 1.0            # Y = 0 reads 0.25 x[clamp(-1)] + 0.75 x[0]
        if r == n - 1: out[-1][1] = 1.0       # Y = 2n-1 reads 0.75 x[n-1] + 0.25 x[clamp(n)]
        return out

    dx = dU.new_zeros(B, C, H, W)
    band = 4
    for r in range(H):
        for c in range(W):
            if band <= r < H - band and band <= c < W - band:
                continue
            acc = 0
            for Y, wy in taps(r, H):
                for X, wx in taps(c, W):
                    acc = acc + wy * wx * dU[:, :, Y, X]
            dx[:, :, r, c] = acc
    return dx


def backward(x: torch.Tensor, w5: torch.Tensor, dy: torch.Tensor):
    """(dx, dw5, dbias) of y = forward(x, w5) for dL/dy = dy, in the form the kernels compute them."""
    B, C, H, W = x.shape
    Co = w5.shape[0]
    fold = fold_matrix(x.dtype)
    wp = phase_weights(w5)
    hb = 2 * SB
    # ---- interior: phase kernels on the band-zeroed space-to-depth copy of dy -----------------------------------
    dyi = dy.clone()
    dyi[:, :, :hb] = 0; dyi[:, :, 2 * H - hb:] = 0; dyi[:, :, :, :hb] = 0; dyi[:, :, :, 2 * W - hb:] = 0
    xz = F.pad(x, (2, 2, 2, 2))
    dxz = torch.zeros_like(xz)
    dwp = torch.zeros_like(wp)
    for a in range(2):
        for b in range(2):
            dyp = dyi[:, :, a::2, b::2]
            win = xz[:, :, a:a + H + 3, b:b + W + 3]
            dwp[a, b] = F.conv2d(win.transpose(0, 1), dyp.transpose(0, 1)).transpose(0, 1)
            dxz[:, :, a:a + H + 3, b:b + W + 3] += F.conv_transpose2d(dyp, wp[a, b])
    dx = dxz[:, :, 2:-2, 2:-2].clone()                             # nothing reaches the ring: dyi is zero in the band
    assert float(dxz[:, :, :2].abs().max()) == 0 and float(dxz[:, :, :, :2].abs().max()) == 0
    dw5 = torch.einsum("aboipq,akp,blq->oikl", dwp, fold, fold)
    db = dyi.sum((0, 2, 3))
    # ---- band: standard kernels on strips ------------------------------------------------------------------------
    u_rows, u_cols = strips_of(up2(x), SR)
    dy_rows, dy_cols = strips_of(dy, hb, corners=False)
    w5t = w5.transpose(2, 3)
    dU_rows, dU_cols = torch.zeros_like(u_rows), torch.zeros_like(u_cols)
    dwt = torch.zeros_like(w5t)
    for s in range(2):
        dU_rows[s] = F.conv_transpose2d(dy_rows[s], w5, padding=2)
        dw5 = dw5 + F.conv2d(F.pad(u_rows[s], (2, 2, 2, 2)).transpose(0, 1), dy_rows[s].transpose(0, 1)).transpose(0, 1)
        dU_cols[s] = F.conv_transpose2d(dy_cols[s], w5t, padding=2)
        dwt = dwt + F.conv2d(F.pad(u_cols[s], (2, 2, 2, 2)).transpose(0, 1), dy_cols[s].transpose(0, 1)).transpose(0, 1)
        db = db + dy_rows[s].sum((0, 2, 3)) + dy_cols[s].sum((0, 2, 3))
    dw5 = dw5 + dwt.transpose(2, 3)
    dx = dx + up2_transposed_band(dU_rows, dU_cols, H, W)
    return dx, dw5, db


def check(B=2, C=3, Co=4, H=9, W=11, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, dtype=torch.float64, generator=g, requires_grad=True)
    w5 = torch.randn(Co, C, 5, 5, dtype=torch.float64, generator=g, requires_grad=True)
    dy = torch.randn(B, Co, 2 * H, 2 * W, dtype=torch.float64, generator=g)
    ref = F.conv2d(up2(x), w5, padding=2)
    ref.backward(dy)
    y = forward(x.detach(), w5.detach())
    dx, dw5, db = backward(x.detach(), w5.detach(), dy)
    return (float((y - ref).abs().max()), float((dx - x.grad).abs().max()), float((dw5 - w5.grad).abs().max()),
            float((db - dy.sum((0, 2, 3))).abs().max()))


if __name__ == "__main__":
    for shape in ((9, 11), (8, 8), (5, 12), (19, 38)):
        print(shape, "forward %.1e  dx %.1e  dW5 %.1e  dbias %.1e" % check(H=shape[0], W=shape[1]))
