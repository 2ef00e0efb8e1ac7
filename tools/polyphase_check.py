"""First formulation of the polyphase resize-convolution (the engine uses the strip decomposition of
tools/polyphase_strips.py, which needs no replicate padding and no frame term): the resize-convolution of upstream's
UNet decoder levels,

    y = conv5x5(zero_pad_2(bilinear_up2x(x)))          (torch.nn.Upsample(scale_factor=2, mode="bilinear") + Conv2d(k=5, padding=2))

computed WITHOUT materialising the upsampled tensor:

    y = polyphase(x~) - conv5x5(E)

* x~ = x replicate-padded by 2 low-res pixels.  Interpolating x~ with the interior formula
  (U'(2m) = 0.25 x~[m-1] + 0.75 x~[m],  U'(2m+1) = 0.75 x~[m] + 0.25 x~[m+1]) reproduces the clamped bilinear upsampling
  exactly on the image and extends it onto the 2-pixel frame where the reference has ZEROS.
* polyphase: output phase (a, b) in {0,1}^2 of y is a 4 x 4-tap convolution of x~ with weights folded from W5 and the
  interpolation weights: 4 x 16 = 64 tap-GEMMs per low-res pixel instead of 4 x 25 = 100.
* E = U' on the frame (rows / columns -2, -1, 2H, 2H+1), zero inside: the part of U' the reference does not have.  It
  only reaches the outputs within 2 pixels of the border, and E(-1, X) = E(-2, X) = U'(0, X) etc., so conv5x5(E) is a
  handful of 1-D convolutions of the border rows / columns with row-summed weights.

This script checks the identity in float64 (max abs error ~1e-15) for odd / even sizes; it is pure torch on the CPU.
"""
import torch
import torch.nn.functional as F


def phase_weights(w5: torch.Tensor) -> torch.Tensor:
    """w5 [Co,Ci,5,5] -> wp [2,2,Co,Ci,4,4]: phase (a,b) taps over low-res offsets d in {-2+a .. 1+a} (index d + 2 - a)."""
    co, ci = w5.shape[:2]
    # 1-D folding matrix: for output phase a and kernel tap k the hi-res row is R = 2i + a + k - 2 = 2m + r;
    # r = 0 reads 0.25 x[m-1] + 0.75 x[m], r = 1 reads 0.75 x[m] + 0.25 x[m+1]; m - i = floor((a + k - 2) / 2)
    fold = torch.zeros(2, 5, 4, dtype=w5.dtype)        # [phase][tap k][low-res offset index]
    for a in range(2):
        for k in range(5):
            s = a + k - 2
            m, r = s // 2, s % 2
            for dm, wgt in (((-1, 0.25), (0, 0.75)) if r == 0 else ((0, 0.75), (1, 0.25))):
                fold[a, k, m + dm + 2 - a] += wgt       # offsets m + dm in {-2+a .. 1+a}
    # separable in the two dimensions
    return torch.einsum("oikl,akp,blq->aboipq", w5, fold, fold)


def upconv_polyphase(x: torch.Tensor, w5: torch.Tensor) -> torch.Tensor:
    """conv5x5(U') where U' is the interior interpolation of the replicate-padded x, as four 4x4 phase convolutions."""
    B, C, H, W = x.shape
    xt = F.pad(x, (2, 2, 2, 2), mode="replicate")
    wp = phase_weights(w5)
    y = x.new_zeros(B, w5.shape[0], 2 * H, 2 * W)
    for a in range(2):
        for b in range(2):
            # low-res window rows i - 2 + a .. i + 1 + a  ->  padded rows i + a .. i + a + 3
            win = xt[:, :, a:a + H + 3, b:b + W + 3]
            y[:, :, a::2, b::2] = F.conv2d(win, wp[a, b])
    return y


def frame_of_extended_upsampling(x: torch.Tensor) -> torch.Tensor:
    """E on the (2H+4) x (2W+4) canvas: U' on the 2-pixel frame, zero inside."""
    B, C, H, W = x.shape
    up = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    ext = F.pad(up, (2, 2, 2, 2), mode="replicate")      # U' on the frame = the nearest border value of U (see docstring)
    ext[:, :, 2:-2, 2:-2] = 0
    return ext


def upconv_reference(x, w5):
    return F.conv2d(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False), w5, padding=2)


def check(B=2, C=3, Co=4, H=7, W=9, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, dtype=torch.float64, generator=g)
    w5 = torch.randn(Co, C, 5, 5, dtype=torch.float64, generator=g)
    ref = upconv_reference(x, w5)
    poly = upconv_polyphase(x, w5)
    corr = F.conv2d(frame_of_extended_upsampling(x), w5)  # canvas already carries the 2-pixel frame: no padding
    inner = (poly - ref)[:, :, 2:-2, 2:-2]
    err_interior = float(inner.abs().max()) if inner.numel() else 0.0
    err_total = float((poly - corr - ref).abs().max())
    touched = (corr.abs() > 0).any(dim=1).any(dim=0)
    border_only = not bool(touched[2:-2, 2:-2].any())
    return err_interior, err_total, border_only


def fold_matrix(dtype=torch.float64) -> torch.Tensor:
    """[phase a][5x5 tap k][low-res offset index p]: the 1-D map from kernel taps to phase taps used by phase_weights."""
    fold = torch.zeros(2, 5, 4, dtype=dtype)
    for a in range(2):
        for k in range(5):
            s = a + k - 2
            m, r = s // 2, s % 2
            for dm, wgt in (((-1, 0.25), (0, 0.75)) if r == 0 else ((0, 0.75), (1, 0.25))):
                fold[a, k, m + dm + 2 - a] += wgt
    return fold


def upconv_backward_explicit(x: torch.Tensor, w5: torch.Tensor, dy: torch.Tensor):
    """The gradients in the form the kernels will compute them (no autograd):

    * dWp[a,b] = correlation of the replicate-padded low-res input with the phase (a,b) sub-image of dY (a weight gradient
      at LOW resolution, 16 taps), folded back to the 5x5 weights with the same interpolation matrix:
      dW5[o,i,k,l] = sum_{a,b,p,q} dWp[a,b,o,i,p,q] fold[a,k,p] fold[b,l,q];  minus the frame term  dY (*) E.
    * dx~ = sum over phases of the transposed 4x4 convolution of the phase sub-image of dY; the replicate padding folds
      the gradient of the padded border back onto the edge pixels; minus the gradient through the frame term.
    Returns (dx, dw5) for y = upconv_reference(x, w5), dL/dy = dy."""
    B, C, H, W = x.shape
    Co = w5.shape[0]
    fold = fold_matrix(x.dtype)
    wp = phase_weights(w5)
    xt = F.pad(x, (2, 2, 2, 2), mode="replicate")
    dxt = torch.zeros_like(xt)
    dwp = torch.zeros_like(wp)
    for a in range(2):
        for b in range(2):
            dyp = dy[:, :, a::2, b::2]                                     # [B,Co,H,W] phase sub-image
            win = xt[:, :, a:a + H + 3, b:b + W + 3]
            # weight gradient of a 4x4 correlation: dwp[o,i,p,q] = sum_{n,y,x} dyp[n,o,y,x] win[n,i,y+p,x+q]
            dwp[a, b] = F.conv2d(win.transpose(0, 1), dyp.transpose(0, 1)).transpose(0, 1)
            # input gradient: full correlation with the flipped kernel
            dxt[:, :, a:a + H + 3, b:b + W + 3] += F.conv_transpose2d(dyp, wp[a, b])
    dw5 = torch.einsum("aboipq,akp,blq->oikl", dwp, fold, fold)
    # replicate padding backward: border strips accumulate onto the edge rows / columns
    dx = dxt[:, :, 2:-2, 2:-2].clone()
    dx[:, :, 0] += dxt[:, :, :2, 2:-2].sum(2); dx[:, :, -1] += dxt[:, :, -2:, 2:-2].sum(2)
    dx[:, :, :, 0] += dxt[:, :, 2:-2, :2].sum(3); dx[:, :, :, -1] += dxt[:, :, 2:-2, -2:].sum(3)
    dx[:, :, 0, 0] += dxt[:, :, :2, :2].sum((2, 3)); dx[:, :, 0, -1] += dxt[:, :, :2, -2:].sum((2, 3))
    dx[:, :, -1, 0] += dxt[:, :, -2:, :2].sum((2, 3)); dx[:, :, -1, -1] += dxt[:, :, -2:, -2:].sum((2, 3))
    # frame term  - conv5x5(E(x)):  E is linear in x (upsample, replicate-extend, zero the interior)
    E = frame_of_extended_upsampling(x)
    dw5 = dw5 - F.conv2d(E.transpose(0, 1), dy.transpose(0, 1)).transpose(0, 1)
    dE = F.conv_transpose2d(dy, w5)                                        # gradient w.r.t. the (2H+4) x (2W+4) canvas
    dE[:, :, 2:-2, 2:-2] = 0
    # back through the replicate extension of U (frame -> nearest border pixel of U) ...
    dU = torch.zeros(B, C, 2 * H, 2 * W, dtype=x.dtype)
    dU[:, :, 0] += dE[:, :, :2, 2:-2].sum(2); dU[:, :, -1] += dE[:, :, -2:, 2:-2].sum(2)
    dU[:, :, :, 0] += dE[:, :, 2:-2, :2].sum(3); dU[:, :, :, -1] += dE[:, :, 2:-2, -2:].sum(3)
    dU[:, :, 0, 0] += dE[:, :, :2, :2].sum((2, 3)); dU[:, :, 0, -1] += dE[:, :, :2, -2:].sum((2, 3))
    dU[:, :, -1, 0] += dE[:, :, -2:, :2].sum((2, 3)); dU[:, :, -1, -1] += dE[:, :, -2:, -2:].sum((2, 3))
    # ... and through the bilinear upsampling (only the border rows / columns of dU are non-zero)
    xg = x.detach().clone().requires_grad_(True)
    F.interpolate(xg, scale_factor=2, mode="bilinear", align_corners=False).backward(dU)
    dx = dx - xg.grad
    return dx, dw5


def check_backward(B=2, C=3, Co=4, H=7, W=9, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, dtype=torch.float64, generator=g, requires_grad=True)
    w5 = torch.randn(Co, C, 5, 5, dtype=torch.float64, generator=g, requires_grad=True)
    dy = torch.randn(B, Co, 2 * H, 2 * W, dtype=torch.float64, generator=g)
    upconv_reference(x, w5).backward(dy)
    dx, dw5 = upconv_backward_explicit(x.detach(), w5.detach(), dy)
    return float((dx - x.grad).abs().max()), float((dw5 - w5.grad).abs().max())


def phase_pair_plan(a: int):
    """Positions of a conv_tc2-style plan for output row phase ``a`` with the two x-phases as the two lane groups
    (the layout the stride-2 dgrad already uses): a position is a window offset (r, c) in the replicate-padded low-res
    tensor relative to the low-res pixel (i, j) -- padded index (i + r, j + c) -- and carries, per x-phase b, the index
    (p, q) of the 4x4 phase tap it multiplies, or None.  Row taps: padded rows i + a + p (p = 0..3); column taps of
    x-phase b: padded columns j + b + q (q = 0..3).  Union over b: 4 rows x 5 columns = 20 positions per 16-channel K
    block (x-phase 0 has no tap at column offset 4, x-phase 1 none at offset 0)."""
    plan = []
    for p in range(4):
        for c in range(5):
            taps = []
            for b in range(2):
                q = c - b
                taps.append((p, q) if 0 <= q < 4 else None)
            plan.append(((a + p, c), taps))
    return plan


def upconv_by_plan(x: torch.Tensor, w5: torch.Tensor) -> torch.Tensor:
    """The polyphase forward evaluated position by position exactly as the tensor-core kernel would accumulate it
    (one 'MMA' per position: all pixels of a row x all input channels x both x-phases)."""
    B, C, H, W = x.shape
    xt = F.pad(x, (2, 2, 2, 2), mode="replicate")
    wp = phase_weights(w5)
    y = x.new_zeros(B, w5.shape[0], 2 * H, 2 * W)
    n_pos = 0
    for a in range(2):
        plan = phase_pair_plan(a)
        n_pos = len(plan)
        for (r, c), taps in plan:
            win = xt[:, :, r:r + H, c:c + W]                      # the shifted window: one pixel per low-res output pixel
            for b, t in enumerate(taps):
                if t is not None:
                    y[:, :, a::2, b::2] += torch.einsum("nchw,oc->nohw", win, wp[a, b][:, :, t[0], t[1]])
    return y, n_pos


if __name__ == "__main__":
    for shape in ((7, 9), (8, 8), (1, 5), (38, 38)):
        ei, et, bo = check(H=shape[0], W=shape[1])
        print(f"H x W = {shape}: interior error {ei:.2e}, with frame correction {et:.2e}, correction confined to the border: {bo}")
        g = torch.Generator().manual_seed(1)
        xx = torch.randn(2, 3, shape[0], shape[1], dtype=torch.float64, generator=g)
        ww = torch.randn(4, 3, 5, 5, dtype=torch.float64, generator=g)
        yp, n_pos = upconv_by_plan(xx, ww)
        print(f"             position plan ({n_pos} positions per K block and row phase): "
              f"{float((yp - upconv_polyphase(xx, ww)).abs().max()):.2e} vs the phase convolutions")
        ex, ew = check_backward(H=shape[0], W=shape[1])
        print(f"             backward (explicit phase formulas vs autograd of the reference): dx {ex:.2e}, dW5 {ew:.2e}")
