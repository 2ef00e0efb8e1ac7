"""CPU checks of the identities behind the polyphase resize-convolution.  tools/polyphase_check.py: the
first formulation (replicate-padded input minus a frame term); tools/polyphase_strips.py: the decomposition the engine
uses (phase kernels inside, standard kernels on strips for the band) -- forward, dx, dW5, dbias in float64.
First formulation:
conv5x5(zero_pad(bilinear_up2x(x))) == four 4x4 phase convolutions of the replicate-padded input minus the 5x5
convolution of a two-pixel frame -- exact in float64, and the frame term only reaches outputs within 2 pixels of the border."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import polyphase_check as P  # noqa: E402


@pytest.mark.parametrize("h,w", [(7, 9), (8, 8), (1, 5), (2, 2), (19, 38)])
def test_polyphase_identity(h, w):
    err_interior, err_total, border_only = P.check(H=h, W=w, seed=h * 100 + w)
    assert err_interior < 1e-12 and err_total < 1e-12 and border_only


def test_phase_weights_cut_the_tap_count():
    import torch
    wp = P.phase_weights(torch.randn(4, 3, 5, 5, dtype=torch.float64))
    assert wp.shape == (2, 2, 4, 3, 4, 4)          # 4 phases x 16 taps = 64 tap-GEMMs per low-res pixel (vs 4 x 25)


@pytest.mark.parametrize("h,w", [(7, 9), (8, 8), (1, 5), (13, 4)])
def test_polyphase_backward_formulas(h, w):
    """Input and weight gradients in the form the kernels would compute them (phase-wise low-resolution weight gradient
    folded back through the interpolation matrix, phase-wise transposed convolutions, replicate-pad folding, frame term)
    equal autograd of upsample + conv."""
    err_dx, err_dw = P.check_backward(H=h, W=w, seed=h * 10 + w)
    assert err_dx < 1e-11 and err_dw < 1e-11


def test_polyphase_position_plan_matches_phase_convolutions():
    """The per-position accumulation a conv_tc2 phase kind would run (20 window offsets per K block and row phase, the
    two x-phases as lane groups) reproduces the four phase convolutions."""
    import torch
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 5, 6, 11, dtype=torch.float64, generator=g)
    w = torch.randn(3, 5, 5, 5, dtype=torch.float64, generator=g)
    y, n_pos = P.upconv_by_plan(x, w)
    assert n_pos == 20
    assert float((y - P.upconv_polyphase(x, w)).abs().max()) < 1e-12


@pytest.mark.parametrize("h,w", [(9, 11), (8, 8), (5, 12), (19, 38)])
def test_strip_decomposition_forward_and_backward(h, w):
    """The partition the engine runs (interior: four 4x4 phase convolutions on the zero-ringed tensor; band of 2 low-res
    pixels: the 5x5 convolution on strips of the upsampled tensor, column strips transposed; every dY pixel used once in
    the backward; strips' input gradient folded through the transposed bilinear map) equals autograd of Upsample + Conv."""
    import polyphase_strips as S
    ey, ex, ew, eb = S.check(H=h, W=w, seed=h * 31 + w)
    assert ey < 1e-12 and ex < 1e-12 and ew < 1e-11 and eb < 1e-11
