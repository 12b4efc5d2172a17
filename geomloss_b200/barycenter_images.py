"""Wasserstein barycenters of 2-D images on the separable grid softmin.

Restates ``geomloss.ImagesBarycenter`` (src/geomloss/_legacy/wasserstein_barycenter_images.py:6-93): a multiscale
(2x2 -> NxN), eps-scaled, debiased Sinkhorn barycenter iteration whose only non-elementwise operator is the
reference's ``softmin_grid`` (utils.py:190-279) — here ``csrc/b200ot_grid.cu`` behind ``b200ot_softmin_grid``.

Autograd contract of the reference: with ``backward_iterations = 0`` the whole descent is differentiated, otherwise
only ``backward_iterations`` extra iterations at the final scale.  The grid softmin is made differentiable w.r.t. its
input here with a closed-form backward that is ITSELF two grid softmins (no N^2 x N^2 matrix, no atomics):

    out_i = -eps log sum_j exp(h_j - C_ij/eps)    =>    d<go, out>/dh_j = -eps e^{h_j} sum_i (go_i e^{out_i/eps}) e^{-C_ij/eps}

and the symmetric kernel sum on the right is ``exp(-softmin_grid(log(go^+-) + out/eps)/eps)`` for the positive and
negative parts of ``go``.  Parity: pinned — tests/golden/img_bary_*.npz hold the outputs (and autograd gradients
w.r.t. weights and measures) of the unmodified reference run on tests/golden/pykeops_shim
(tests/golden/make_golden_images.py); tests/test_gpu_reference_goldens.py::test_images_barycenter_vs_reference checks
the CUDA path against them, the oracle's dense restatement (oracle.images_barycenter) is pinned to the same files.
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


class _Descent:
    """State of the debiased barycenter scheme at one resolution: the K pairs of dual potentials (`to_bar` lives
    on the barycenter side, `to_meas` on the measures' side), the log of the de-biasing density and the current
    log-barycenter.  One `step` = wasserstein_barycenter_images.py:6-34."""

    def __init__(self, log_meas0, weights, p):
        self.p, self.weights = p, weights  # (the broadcast view is taken per use: it must see the live grad mode)
        start = _softmin(1.0, p, log_meas0)
        self.to_bar, self.to_meas = start, start.clone() if start.requires_grad else start
        flat = torch.ones_like(log_meas0).sum(dim=1, keepdim=True)
        self.debias = flat - flat.logsumexp([2, 3], keepdim=True)  # uniform probability on the coarsest grid
        self.log_bar = None

    def _pull(self, eps, log_meas):
        """log-barycenter implied by the measures-side potentials: debias - sum_k lam_k softmin_k / eps."""
        transported = _softmin(eps, self.p, log_meas + self.to_meas / eps)
        return self.debias - (transported * self.weights[:, :, None, None]).sum(1, keepdim=True) / eps

    def step(self, eps, log_meas):
        p = self.p
        bar = self._pull(eps, log_meas)
        # symmetric (averaged) Sinkhorn updates between every measure and the current barycenter estimate
        new_to_bar = 0.5 * (self.to_bar + _softmin(eps, p, log_meas + self.to_meas / eps))
        new_to_meas = 0.5 * (self.to_meas + _softmin(eps, p, bar + self.to_bar / eps))
        self.to_bar, self.to_meas = new_to_bar, new_to_meas
        self.log_bar = self._pull(eps, log_meas)
        self.debias = 0.5 * (self.debias + self.log_bar + _softmin(eps, p, self.debias) / eps)

    def refine(self):
        self.to_bar, self.to_meas, self.debias = upsample(self.to_bar), upsample(self.to_meas), upsample(self.debias)


def ImagesBarycenter(measures, weights, blur=0, p=2, scaling_N=10, backward_iterations=5):
    """Barycenter of ``measures`` (B, K, N, N) with barycentric ``weights`` (B, K); returns (B, 1, N, N).

    Same arguments, schedule (``scaling_N`` steps per resolution 2x2 .. NxN, the blur scale halved per
    resolution down to ``blur``, default one pixel) and autograd contract as the reference
    (wasserstein_barycenter_images.py:37-93)."""
    if measures.dim() != 4 or measures.shape[-1] != measures.shape[-2]:
        raise ValueError("measures must be a (B, K, N, N) batch of square images")
    floor = blur if blur != 0 else 1 / measures.shape[-1]
    shrink = 2 ** (-1 / scaling_N)
    with torch.set_grad_enabled(torch.is_grad_enabled() and backward_iterations == 0):
        levels = [log_dens(t) for t in pyramid(measures)[1:]]  # 2x2, 4x4, .. NxN; the 1-pixel level is dropped
        state = _Descent(levels[0], weights, p)
        sigma = 1.0
        for depth, log_meas in enumerate(levels):
            for _ in range(scaling_N):
                eps = sigma**p
                state.step(eps, log_meas)
                sigma = max(sigma * shrink, floor)
            if depth + 1 < len(levels):
                state.refine()
    if backward_iterations > 0 and (measures.requires_grad or weights.requires_grad):
        # extra differentiable steps at the final resolution and temperature, from detached potentials
        for _ in range(backward_iterations):
            state.step(eps, log_meas)
    return state.log_bar.exp()
