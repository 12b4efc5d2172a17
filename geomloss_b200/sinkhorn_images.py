"""Sinkhorn divergence between measures on regular 2-D / 3-D grids (images, volumes).

Host-side counterpart of src/geomloss/_legacy/sinkhorn_images.py:26-202 and of the grid helpers of
src/geomloss/_legacy/utils.py:69-108: a multiscale (pyramid) epsilon-scaling Sinkhorn loop whose only
heavy operator is the separable grid softmin — here ``b200ot_softmin_grid`` (csrc/b200ot_grid.cu) instead of
three pykeops LazyTensor reductions and two permutes.  The pyramid (2x sum-pooling), the log-densities
(floor -10000) and the bi/tri-linear upsampling of the potentials at a scale jump are cheap elementwise
torch ops on the device, exactly as in the reference.

Inputs follow the reference's actual contract: ``(B, C, N, N)`` or
``(B, C, N, N, N)`` float32 CUDA tensors with equal power-of-two sides.  Autograd: like the reference,
only the weights receive gradients (the potentials), through the final value formula.
"""
from __future__ import annotations


import numpy as np
import torch
from torch.nn.functional import avg_pool2d, avg_pool3d, interpolate

from . import _lib, ops
from .sinkhorn import damping, epsilon_schedule


def _dimension(t):
    return t.dim() - 2


def pyramid(t):
    """[coarsest (1 pixel), ..., t]: repeated 2x sum-pooling.                              utils.py:77-97"""
    d = _dimension(t)
    levels = [t]
    for _ in range(int(np.log2(t.shape[2]))):
        t = 4 * avg_pool2d(t, 2) if d == 2 else 8 * avg_pool3d(t, 2)
        levels.append(t)
    levels.reverse()
    return levels


def upsample(t):
    """2x bi/tri-linear interpolation, align_corners=False.                               utils.py:100-102"""
    return interpolate(t, scale_factor=2, mode="bilinear" if _dimension(t) == 2 else "trilinear",
                       align_corners=False)


def log_dens(a):
    """log-density with empty pixels pinned to -10000.                                    utils.py:105-108"""
    return torch.where(a > 0, a.clamp_min(1e-45).log(), torch.full_like(a, -10000.0))


def softmin_grid(eps, p, h_a, h_b=None, h_scale_b=0.0, *, out_old=None, alpha_old=0.0, beta=1.0):
    """``alpha_old*out_old + beta*softmin_grid(eps, p, h_a + h_scale_b*h_b)`` on a (B, C, N, .., N) grid
    (reference operator: softmin_grid, utils.py:190-279).  One C-ABI call = one launch per axis."""
    h_a = ops._f32c(h_a, "h_a")
    h_b = ops._f32c(h_b, "h_b")
    out_old = ops._f32c(out_old, "out_old")
    d = _dimension(h_a)
    if d not in (2, 3):
        raise ValueError("grids must be (B, C, N, N) or (B, C, N, N, N)")
    N = h_a.shape[-1]
    if any(s != N for s in h_a.shape[2:]):
        raise ValueError("grid sides must be equal")
    dev = h_a.device
    out = torch.empty_like(h_a)
    L = _lib.lib()
    with torch.cuda.device(dev):
        rc = L.b200ot_softmin_grid(ops._ptr(h_a), ops._ptr(h_b), float(h_scale_b), ops._ptr(out_old), float(alpha_old),
                                   float(beta), ops._ptr(out), h_a.shape[0] * h_a.shape[1], N, d, int(p), float(eps),
                                   ops._stream(dev))
    _lib.check(rc, "b200ot_softmin_grid")
    ops.count_launches(d)
    return out


def sinkhorn_divergence(a, b, p=2, blur=None, reach=None, axes=None, scaling=0.5, cost=None, debias=True,
                        potentials=False, verbose=False, **kwargs):
    """Sinkhorn divergence between two batches of images / volumes.        sinkhorn_images.py:26-202

    Returns a ``(B,)`` tensor, or the pair of dual potentials (same shape as the inputs) if
    ``potentials=True``.
    """
    if a.shape != b.shape:
        raise ValueError("a and b must have the same shape")
    if cost is not None:
        raise NotImplementedError()
    if blur is None:
        blur = 1 / a.shape[-1]
    if scaling < 0.5:
        raise ValueError(f"Scaling value of {scaling} is too small: please use a number in [0.5, 1).")

    a_s, b_s = pyramid(a)[1:], pyramid(b)[1:]  # the 1-pixel level is dropped (:110)
    a_logs = [log_dens(t.detach()) for t in a_s]
    b_logs = [log_dens(t.detach()) for t in b_s]
    diameter = 1
    eps_final = blur**p
    rho = None if reach is None else reach**p
    eps_list = epsilon_schedule(p, diameter, blur, scaling)

    # jump to the next (finer) level as soon as its pixels are resolved by the temperature (:152-170)
    scales = [diameter / t.shape[-1] for t in a_s]
    current = scales.pop(0)
    jumps = []
    for i, eps in enumerate(eps_list[1:]):
        if current**p > eps:
            jumps.append(i + 1)
            current = scales.pop(0)
    if verbose:
        print("Pyramid scales:", [diameter / t.shape[-1] for t in a_s])
        print("Temperatures: ", eps_list)
        print("Jumps: ", jumps)
    assert len(jumps) == len(a_s) - 1, "There's a bug in the multicale pre-processing..."

    sm = softmin_grid
    last_extrapolation = True
    with torch.no_grad():
        k = 0
        eps = eps_list[0]
        lam = damping(eps, rho)
        a_log, b_log = a_logs[0], b_logs[0]
        g_ab = sm(eps, p, a_log, beta=lam)
        f_ba = sm(eps, p, b_log, beta=lam)
        if debias:
            f_aa = sm(eps, p, a_log, beta=lam)
            g_bb = sm(eps, p, b_log, beta=lam)
        for i, eps in enumerate(eps_list):
            lam = damping(eps, rho)
            inv = 1.0 / eps
            ft_ba = sm(eps, p, b_log, g_ab, inv, out_old=f_ba, alpha_old=0.5, beta=0.5 * lam)
            gt_ab = sm(eps, p, a_log, f_ba, inv, out_old=g_ab, alpha_old=0.5, beta=0.5 * lam)
            if debias:
                ft_aa = sm(eps, p, a_log, f_aa, inv, out_old=f_aa, alpha_old=0.5, beta=0.5 * lam)
                gt_bb = sm(eps, p, b_log, g_bb, inv, out_old=g_bb, alpha_old=0.5, beta=0.5 * lam)
                f_aa, g_bb = ft_aa, gt_bb
            f_ba, g_ab = ft_ba, gt_ab
            if i in jumps:
                if i == len(eps_list) - 1:
                    last_extrapolation = False  # the up-sampling below is the last step (:520-526)
                f_ba, g_ab = upsample(f_ba), upsample(g_ab)
                if debias:
                    f_aa, g_bb = upsample(f_aa), upsample(g_bb)
                k += 1
                a_log, b_log = a_logs[k], b_logs[k]
        if last_extrapolation:
            inv = 1.0 / eps
            new_f_ba = sm(eps, p, b_log, g_ab, inv, beta=lam)
            new_g_ab = sm(eps, p, a_log, f_ba, inv, beta=lam)
            f_ba, g_ab = new_f_ba, new_g_ab
            if debias:
                f_aa = sm(eps, p, a_log, f_aa, inv, beta=lam)
                g_bb = sm(eps, p, b_log, g_bb, inv, beta=lam)

    return _image_cost(eps_final, rho, a, b, f_aa if debias else None, g_bb if debias else None, g_ab, f_ba,
                       debias, potentials)


def _bdot(a, f):
    n = a.shape[0]
    return (a.reshape(n, -1) * f.reshape(n, -1)).sum(1)


def _image_cost(eps, rho, a, b, f_aa, g_bb, g_ab, f_ba, debias, potentials):
    """sinkhorn_cost with batch=True on grids (sinkhorn_divergence.py:171-250)."""
    if potentials:
        return (f_ba - f_aa, g_ab - g_bb) if debias else (f_ba, g_ab)
    if rho is None:
        if debias:
            return _bdot(a, f_ba - f_aa) + _bdot(b, g_ab - g_bb)
        return _bdot(a, f_ba) + _bdot(b, g_ab)
    w = rho + eps / 2
    if debias:
        return _bdot(a, w * ((-f_aa / rho).exp() - (-f_ba / rho).exp())) + _bdot(
            b, w * ((-g_bb / rho).exp() - (-g_ab / rho).exp()))
    return _bdot(a, w * (1 - (-f_ba / rho).exp())) + _bdot(b, w * (1 - (-g_ab / rho).exp()))
