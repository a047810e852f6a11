"""Known-answer anchors for the (parity-unpinned) Fourier ILT restatement: closed-form Laplace
pairs.  The Fourier series is a modest-accuracy method; the achieved truncation error is
recorded here as loose bounds rather than asserted tight."""
import math

import pytest
import torch

from oracle import ilt

T_GRID = torch.tensor([0.05, 0.1, 0.2, 0.5, 1.0], dtype=torch.float64)


@pytest.mark.parametrize("S", [17, 33, 65, 129])
def test_closed_form_pairs(S):
    pairs = [
        (lambda s: 1.0 / (s + 1.0), lambda t: torch.exp(-t)),
        (lambda s: 1.0 / s ** 2, lambda t: t),
        (lambda s: 2.0 / (s ** 2 + 4.0), lambda t: torch.sin(2.0 * t)),
    ]
    for F, f in pairs:
        got = ilt.fourier_ilt_of(F, T_GRID, S)
        err = (got - f(T_GRID)).abs().max().item()
        assert err < 0.25 / math.sqrt(S / 17), (S, err)  # truncation error shrinks with more terms


def test_more_terms_is_more_accurate():
    F, f = (lambda s: 1.0 / (s + 1.0)), (lambda t: torch.exp(-t))
    errs = [(ilt.fourier_ilt_of(F, T_GRID, S) - f(T_GRID)).abs().max().item() for S in (17, 129, 1025)]
    assert errs[2] < errs[1] < errs[0]
    assert errs[2] < 5e-3


def test_sphere_maps_round_trip():
    g = torch.Generator().manual_seed(0)
    re = torch.randn(1000, generator=g, dtype=torch.float64) * 10
    im = torch.randn(1000, generator=g, dtype=torch.float64) * 10
    th, ph = ilt.complex_to_sphere(re, im)
    re2, im2 = ilt.sphere_to_complex(th, ph)
    assert (re - re2).abs().max() < 1e-9 and (im - im2).abs().max() < 1e-9
    assert th.abs().max() <= math.pi and ph.abs().max() <= math.pi / 2


def test_s_points_constants():
    # T = 2(t+eps), gamma = alpha - ln(tol)/T ; at the planner's normalised t = 0.125
    s_re, s_im, T = ilt.fourier_s_points(torch.tensor([0.125], dtype=torch.float64), 17)
    assert abs(T.item() - 0.250002) < 1e-12
    assert abs(s_re[0, 0].item() - (1e-3 - math.log(1e-2) / 0.250002)) < 1e-12
    assert abs(s_im[0, 16].item() - 16 * math.pi / 0.250002) < 1e-9
