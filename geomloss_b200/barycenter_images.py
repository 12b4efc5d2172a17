"""Wasserstein barycenters of 2-D images on the separable grid softmin.

Restates ``geomloss.ImagesBarycenter`` (src/geomloss/_legacy/wasserstein_barycenter_images.py:6-93): a multiscale
(2x2 -> NxN), eps-scaled, debiased Sinkhorn barycenter iteration whose only non-elementwise operator is the
reference's ``softmin_grid`` (utils.py:190-279) — here ``csrc/b200ot_grid.cu`` behind ``b200ot_softmin_grid``.

Autograd contract of the reference: with ``backward_iterations = 0`` the whole descent is differentiated, otherwise
only ``backward_iterations`` extra iterations at the final scale.  The grid softmin is made differentiable w.r.t. its
input here with a closed-form backward that is ITSELF two grid softmins (no N^2 x N^2 matrix, no atomics):

    out_i = -eps log sum_j exp(h_j - C_ij/eps)    =>    d<go, out>/dh_j = -eps e^{h_j} sum_i (go_i e^{out_i/eps}) e^{-C_ij/eps}

and the symmetric kernel sum on the right is ``exp(-softmin_grid(log(go^+-) + out/eps)/eps)`` for the positive and
negative parts of ``go``.  Parity: unpinned against the reference (its softmin_grid needs pykeops); tests compare with
a dense CPU restatement (oracle.images_barycenter) incl. autograd gradients.
"""
from __future__ import annotations

import torch

from .sinkhorn_images import log_dens, pyramid, softmin_grid, upsample

__all__ = ["ImagesBarycenter"]


class _GridSoftmin(torch.autograd.Function):
    """softmin_grid(eps, p, h) with a gradient w.r.t. h (first order)."""

    @staticmethod
    def forward(ctx, h, eps, p):
        out = softmin_grid(eps, p, h.detach())
        ctx.save_for_backward(h.detach(), out)
        ctx.meta = (float(eps), int(p))
        return out

    @staticmethod
    def backward(ctx, go):
        h, out = ctx.saved_tensors
        eps, p = ctx.meta
        go = go.contiguous()
        grad = torch.zeros_like(h)
        for sign in (1.0, -1.0):
            part = (sign * go).clamp_min(0.0)
            if bool((part > 0).any()):
                lw = torch.where(part > 0, part.clamp_min(1e-45).log(), torch.full_like(part, -1.0e5))
                conv = softmin_grid(eps, p, lw, out, 1.0 / eps)  # -eps log sum_i go_i e^{out_i/eps} e^{-C_ij/eps}
                grad = grad - sign * eps * torch.exp(h - conv / eps)
        return grad, None, None


def _softmin(eps, p, h):
    if torch.is_grad_enabled() and h.requires_grad:
        return _GridSoftmin.apply(h, eps, p)
    return softmin_grid(eps, p, h.detach())


def barycenter_iteration(f_k, g_k, d_log, eps, p, ak_log, w_k):
    """One debiased barycenter step (wasserstein_barycenter_images.py:6-34)."""
    w = w_k[:, :, None, None]
    ft_k = _softmin(eps, p, ak_log + g_k / eps) / eps  # pseudo-step: measures -> barycenter
    bar_log = d_log - (ft_k * w).sum(1, keepdim=True)
    ft_k = _softmin(eps, p, ak_log + g_k / eps)  # symmetric Sinkhorn updates
    gt_k = _softmin(eps, p, bar_log + f_k / eps)
    f_k = (f_k + ft_k) / 2
    g_k = (g_k + gt_k) / 2
    ft_k = _softmin(eps, p, ak_log + g_k / eps) / eps
    bar_log = d_log - (ft_k * w).sum(1, keepdim=True)
    d_log = 0.5 * (d_log + bar_log + _softmin(eps, p, d_log) / eps)  # de-biasing measure
    return f_k, g_k, d_log, bar_log


def ImagesBarycenter(measures, weights, blur=0, p=2, scaling_N=10, backward_iterations=5):
    """Barycenter of ``measures`` (B, K, N, N) with barycentric ``weights`` (B, K); returns (B, 1, N, N).

    Same arguments and iteration counts as the reference (wasserstein_barycenter_images.py:37-93)."""
    a_k, w_k = measures, weights
    if a_k.dim() != 4 or a_k.shape[-1] != a_k.shape[-2]:
        raise ValueError("measures must be a (B, K, N, N) batch of square images")
    if blur == 0:
        blur = 1 / a_k.shape[-1]
    with torch.set_grad_enabled(torch.is_grad_enabled() and backward_iterations == 0):
        ak_s = pyramid(a_k)[1:]
        ak_log_s = [log_dens(t) for t in ak_s]
        sigma = 1
        eps = sigma**p
        f_k, g_k = _softmin(eps, p, ak_log_s[0]), _softmin(eps, p, ak_log_s[0])
        d_log = torch.ones_like(ak_log_s[0]).sum(dim=1, keepdim=True)
        d_log = d_log - d_log.logsumexp([2, 3], keepdim=True)
        for n, ak_log in enumerate(ak_log_s):
            for _ in range(scaling_N):
                eps = sigma**p
                f_k, g_k, d_log, bar_log = barycenter_iteration(f_k, g_k, d_log, eps, p, ak_log, w_k)
                sigma = max(sigma * (2 ** (-1 / scaling_N)), blur)
            if n + 1 < len(ak_s):
                f_k, g_k, d_log = upsample(f_k), upsample(g_k), upsample(d_log)
    if (measures.requires_grad or weights.requires_grad) and backward_iterations > 0:
        for _ in range(backward_iterations):
            f_k, g_k, d_log, bar_log = barycenter_iteration(f_k, g_k, d_log, eps, p, ak_log, w_k)
    return bar_log.exp()
