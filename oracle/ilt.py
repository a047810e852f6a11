"""Fourier-series inverse Laplace transform + Riemann-sphere maps (CPU oracle).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

PARITY UNPINNED.  The reference delegates this arithmetic to the third-party
package ``torchlaplace`` (``w_nl.py:6`` import, ``w_nl.py:137-144`` call site;
``requirements.txt:17`` lists it with no version pin).  Its source is not in
``/root/reference`` and the package is not installable here, and the reference
has no test or golden vector at this boundary.  What follows is a restatement
of that package's published algorithm:

* Fourier-series ILT (Dubner-Abate / Crump family as used by torchlaplace's
  ``Fourier`` class, defaults ``alpha=1e-3``, ``tol=10*alpha``, ``scale=2``,
  ``eps=1e-6``):  ``T = scale*(t+eps)``, ``gamma = alpha - ln(tol)/T``,
  query points ``s_k = gamma + i*k*pi/T`` for ``k = 0..S-1`` and

      x(t) = exp(gamma*t)/T * [ Re F(s_0)/2 + sum_{k>=1} Re( F(s_k) e^{i k pi t / T} ) ]

* stereographic Riemann-sphere coordinates
  ``theta = atan2(Im s, Re s)``, ``phi = asin((|s|^2-1)/(|s|^2+1))`` and back
  ``s = tan(phi/2 + pi/4) * (cos theta + i sin theta)``.

* ``laplace_reconstruct(rep_func, p, t, recon_dim, "fourier", S)``: the
  representation network is fed ``[theta(s_0..s_{S-1}) | phi(s_0..s_{S-1}) | p]``
  per (trajectory, time) pair - the input width ``2*S + dim(p)`` and the
  ``(-1, 2*recon_dim, S)`` output layout are fixed by the reference's own
  ``LaplaceRepresentationFunc`` (``w_nl.py:41,45,56-63``) - and returns
  ``(batch, n_t, recon_dim)``.

It is anchored on closed-form Laplace pairs in ``tests/test_oracle_ilt.py``.
"""
from __future__ import annotations

import math

import torch

ALPHA = 1.0e-3
TOL = 10.0 * ALPHA
SCALE = 2.0
EPS = 1.0e-6


def fourier_s_points(t: torch.Tensor, n_terms: int, alpha=ALPHA, tol=TOL, scale=SCALE, eps=EPS):
    """Query points of the Fourier ILT for times ``t`` (any shape).

    Returns ``(s_re, s_im, T)`` with ``s_*`` of shape ``t.shape + (n_terms,)``.
    """
    T = scale * (t + eps)
    gamma = alpha - math.log(tol) / T
    k = torch.arange(n_terms, dtype=t.dtype, device=t.device)
    s_re = gamma.unsqueeze(-1).expand(*t.shape, n_terms).clone()
    s_im = math.pi * k / T.unsqueeze(-1)
    return s_re, s_im, T


def complex_to_sphere(s_re: torch.Tensor, s_im: torch.Tensor):
    """Complex plane -> Riemann sphere angles (theta in (-pi,pi], phi in [-pi/2,pi/2])."""
    r2 = s_re * s_re + s_im * s_im
    theta = torch.atan2(s_im, s_re)
    phi = torch.asin((r2 - 1.0) / (r2 + 1.0))
    return theta, phi


def sphere_to_complex(theta: torch.Tensor, phi: torch.Tensor):
    """Riemann sphere angles -> complex plane, returned as (re, im)."""
    r = torch.tan(phi / 2.0 + math.pi / 4.0)
    return r * torch.cos(theta), r * torch.sin(theta)


def fourier_line_integrate(f_re: torch.Tensor, f_im: torch.Tensor, t: torch.Tensor, T: torch.Tensor,
                           alpha=ALPHA, tol=TOL):
    """Sum the Fourier series.  ``f_*``: (..., S); ``t``, ``T`` broadcastable to (...)."""
    n_terms = f_re.shape[-1]
    gamma = alpha - math.log(tol) / T
    k = torch.arange(n_terms, dtype=f_re.dtype, device=f_re.device)
    ang = k * (math.pi * t / T).unsqueeze(-1)
    terms = f_re * torch.cos(ang) - f_im * torch.sin(ang)
    series = 0.5 * terms[..., 0] + terms[..., 1:].sum(-1)
    return torch.exp(gamma * t) / T * series


def laplace_reconstruct(laplace_rep_func, p, t, recon_dim=None, ilt_algorithm="fourier",
                        ilt_reconstruction_terms=33, **_unused):
    """Restatement of ``torchlaplace.laplace_reconstruct`` for the one configuration the
    reference uses (``w_nl.py:137-144``): Fourier ILT with sphere projection.

    ``p``: (K, L) latent; ``t``: (K, n_t) (or (K,) / (K,1)) times; returns (K, n_t, recon_dim).
    ``laplace_rep_func`` maps (K*n_t, 2S+L)-viewable input to ``(theta, phi)`` each
    ``(K*n_t, recon_dim, S)``.
    """
    if ilt_algorithm != "fourier":
        raise NotImplementedError("oracle restates the fourier ILT only")
    if t.dim() == 1:
        t = t.view(-1, 1)
    K, n_t = t.shape
    S = int(ilt_reconstruction_terms)
    s_re, s_im, T = fourier_s_points(t, S)
    th_s, ph_s = complex_to_sphere(s_re, s_im)
    inp = torch.cat((th_s, ph_s, p.unsqueeze(1).expand(K, n_t, p.shape[-1])), dim=-1)
    theta, phi = laplace_rep_func(inp)
    if recon_dim is None:
        recon_dim = theta.shape[-2]
    theta = theta.reshape(K, n_t, recon_dim, S)
    phi = phi.reshape(K, n_t, recon_dim, S)
    f_re, f_im = sphere_to_complex(theta, phi)
    return fourier_line_integrate(f_re, f_im, t.unsqueeze(-1), T.unsqueeze(-1))


def fourier_ilt_of(F, t: torch.Tensor, n_terms: int):
    """ILT of a closed-form ``F(s)`` (callable on complex tensors) - known-answer anchor."""
    s_re, s_im, T = fourier_s_points(t, n_terms)
    Fs = F(torch.complex(s_re, s_im))
    return fourier_line_integrate(Fs.real, Fs.imag, t, T)
